#!/bin/bash
# same-box A/B in the SUSTAINED (power-capped) regime: bench.py's own timed loop (25 back-to-back solves), alternating
# (library, options) pairs.  usage: tools/ab_sustained3.sh rounds "lib|opts" ...   lib 'default' = in-tree; opts e.g. ring_blocked=0
R=$1; shift
for r in $(seq 1 $R); do for f in "$@"; do L=${f%%|*}; O=${f#*|}; [ "$O" = "$f" ] && O=""; echo -n "$L [$O]: "; if [ "$L" = default ]; then LP=""; else LP=$PWD/$L; fi
  SDDE_B200_LIB=$LP SDDE_B200_OPTS=$O python bench.py --steps 25 --warmup 5 --no-cpu-baseline --no-e2e --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step  kernel', round(d['roofline']['kernel_ms'],3), ' sm', d['clocks']['sm_mhz'], d['clocks']['power_w_max'], d['clocks']['reasons'])"; done; done
