#!/bin/bash
# round 2 evidence at the headline size (one 160^3 subdomain): (1) ncu --set full of the shipped single-right-hand-side sweep kernels
# (k_fwd_blk<1>, k_bwd<1,false>) of one solve, graphs off; (2) launch list (time + DRAM bytes) of whole deflated applies
M=${1:-160}
mkdir -p gpurun_out
HPDDM_B200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fwd|k_bwd" -s 34 -c 34 -o gpurun_out/r02_prof_sptrsv_m$M -f python profiles/run_solve.py $M 2 > gpurun_out/r02_prof_sptrsv_m$M.log 2>&1
tail -1 gpurun_out/r02_prof_sptrsv_m$M.log
HPDDM_B200_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"kk_|k_fwd|k_bwd|k_perm" -c 400 --csv --log-file gpurun_out/r02_launches_apply_m$M.csv python profiles/run_solve.py $M 0 2 > gpurun_out/r02_apply_m$M.log 2>&1
tail -1 gpurun_out/r02_apply_m$M.log; wc -l gpurun_out/r02_launches_apply_m$M.csv
