#!/bin/bash
# same-box A/B in the SUSTAINED (power-capped) regime: the bench's own 20-step timed loop, alternating library builds
for r in 1 2; do for f in "$@"; do echo -n "$f: "; SDDE_B200_LIB=$PWD/$f python bench.py --steps 25 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'ms/step  kernel', round(d['roofline']['kernel_ms'],3), ' sm', d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"; done; done
