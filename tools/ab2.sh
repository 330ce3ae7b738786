#!/bin/bash
# same-box A/B, 2 rounds, best of 4 launches each
for r in 1 2; do for f in "$@"; do echo -n "$f: "; SDDE_B200_LIB=$PWD/$f python tools/profile_one.py 1e6 500 4 2>&1 | sort -n | head -1; done; done
