#!/bin/bash
# same-box A/B of the bench's e2e arm: zero-copy, staged, auto (SDDE_B200_OPTS), a few rounds
for i in 1 2 3; do for O in zero_copy=1 zero_copy=0 zero_copy=-1; do echo -n "$O: "
  SDDE_B200_OPTS=$O python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python tools/_e2e_line.py; done; done
