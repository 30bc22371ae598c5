#!/bin/bash
# ncu launch list (time + DRAM bytes per launch) of full deflated applies at m = $1 (default 128), graphs disabled so
# that every sweep kernel is an ordinary launch.  Output: gpurun_out/launches_apply_m$M.csv
M=${1:-128}
mkdir -p gpurun_out
HPDDM_B200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"kk_|k_fwd|k_bwd|k_perm" -c 400 --csv --log-file gpurun_out/launches_apply_m$M.csv python profiles/run_solve.py $M 0 3 > gpurun_out/apply_m$M.log 2>&1
tail -2 gpurun_out/apply_m$M.log
