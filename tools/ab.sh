#!/bin/bash
# same-box A/B: alternate two library builds, 3 rounds
for r in 1 2 3; do for f in "$@"; do echo -n "$f: "; SDDE_B200_LIB=$PWD/$f python tools/profile_one.py 1e6 500 4 2>&1 | tail -1; done; done
