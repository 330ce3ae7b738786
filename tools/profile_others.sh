#!/bin/bash
# ncu --set full captures of the kernels other than the C2 bench kernel (each after the same command exited 0 without ncu),
# summarised on the box (the reports themselves are too large to bring back together)
# usage: tools/profile_others.sh TAG   -> gpurun_out/TAG_<cfg>_summary.txt
T=$1
for spec in sir:1.001e9:32 c1:1.001e9:64 eh:1.001e9:64 sra1:1.004e9:64 c3ito:2.002e9:64; do
  cfg=${spec%%:*}; rest=${spec#*:}; nts=${rest%%:*}; tpw=${rest#*:}
  python tools/profile_cfg.py $cfg 2 > gpurun_out/${T}_${cfg}_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:sdde_solve_kernel -c 1 -o /tmp/${T}_${cfg} -f \
      python tools/profile_cfg.py $cfg 1 > gpurun_out/${T}_${cfg}_ncu.log 2>&1
  python tools/ncu_summary.py /tmp/${T}_${cfg}.ncu-rep $nts $tpw > gpurun_out/${T}_${cfg}_summary.txt 2>&1
  tail -1 gpurun_out/${T}_${cfg}_plain.log
  rm -f /tmp/${T}_${cfg}.ncu-rep
done
ls -la gpurun_out/${T}_*summary.txt
