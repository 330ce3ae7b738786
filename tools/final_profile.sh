#!/bin/bash
# round-end evidence run on one B200: GPU tests, the bench line, the ncu launch list of the bench command and one
# `ncu --set full` capture of the step kernel (each ncu pass only after the same command exited 0 without ncu).
# usage: tools/final_profile.sh TAG     -> gpurun_out/TAG_*
T=$1
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || { echo bench failed; tail -5 gpurun_out/${T}_bench.err; exit 1; }
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_b.log 2>&1
python tools/profile_one.py 1e6 500 1 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sdde_solve_kernel -c 1 -o gpurun_out/${T}_bench -f \
    python tools/profile_one.py 1e6 500 1 > gpurun_out/${T}_ncu.log 2>&1
ls -la gpurun_out/${T}_*
