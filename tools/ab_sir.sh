#!/bin/bash
for r in 1 2; do for f in "$@"; do echo -n "$f: "; SDDE_B200_LIB=$PWD/$f python tools/probe_sir.py 2>&1 | tail -1; done; done
