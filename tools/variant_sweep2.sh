#!/bin/bash
for f in variants/lib_*.so; do for ch in 32 128; do echo "== $f chunk $ch"; SDDE_B200_LIB=$PWD/$f SDDE_CHUNK=$ch python tools/profile_one.py 1e6 500 3 2>&1 | tail -1; done; done
echo "== default"; python tools/profile_one.py 1e6 500 3 2>&1 | tail -1
