#!/bin/bash
for r in 1 2 3; do for f in "$@"; do echo -n "$f: "; SDDE_B200_LIB=$PWD/$f python tools/probe_c4.py 1e6 2>&1 | grep "lanes=1" | sed 's/.*kernel \([0-9.]* ms\).*/\1/'; done; done
