#!/bin/bash
# run the C2 probe for every library variant under variants/
for f in variants/lib_*.so; do
  echo "== $f"
  SDDE_B200_LIB=$PWD/$f python tools/profile_one.py 1e6 500 3 2>&1 | tail -2
done
echo "== default"; python tools/profile_one.py 1e6 500 3 2>&1 | tail -2
