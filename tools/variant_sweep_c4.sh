#!/bin/bash
for f in variants/lib_*.so; do echo "== $f"; SDDE_B200_LIB=$PWD/$f python tools/probe_c4.py 1e6 2>&1 | grep "lanes=1"; done
echo "== default"; python tools/probe_c4.py 1e6 2>&1 | grep "lanes=1"
