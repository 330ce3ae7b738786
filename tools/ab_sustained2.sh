#!/bin/bash
# same-box A/B in the SUSTAINED (power-capped) regime: the bench's own timed loop (25 back-to-back solves), alternating builds
# usage: tools/ab_sustained2.sh [rounds] lib...   ('default' = in-tree library)
R=$1; shift
for r in $(seq 1 $R); do for f in "$@"; do echo -n "$f: "; if [ "$f" = default ]; then L=""; else L=$PWD/$f; fi
  SDDE_B200_LIB=$L python bench.py --steps 25 --warmup 5 --no-cpu-baseline --no-e2e --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step  kernel', round(d['roofline']['kernel_ms'],3), ' sm', d['clocks']['sm_mhz'], d['clocks']['power_w_max'], d['clocks']['reasons'])"; done; done
