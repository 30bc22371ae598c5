#!/bin/bash
# ncu --set full of the SpTRSV sweep kernels of one solve at m = $1 (graphs off) -> gpurun_out/prof_sptrsv_m$M.ncu-rep
M=${1:-128}
mkdir -p gpurun_out
HPDDM_B200_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"k_fwd|k_bwd" -s 32 -c 32 -o gpurun_out/prof_sptrsv_m$M -f python profiles/run_solve.py $M 2 > gpurun_out/prof_sptrsv_m$M.log 2>&1
tail -1 gpurun_out/prof_sptrsv_m$M.log
