#!/bin/bash
# ncu --set full of the block (4 right-hand sides) SpTRSV kernels on the large levels of one m = 96 solve (graphs off)
mkdir -p gpurun_out
HPDDM_B200_NO_GRAPH=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_fwd|k_bwd" -s 40 -c 8 -o gpurun_out/prof_blk4_m96 -f python profiles/run_solve.py 96 2 0 d 4 > gpurun_out/prof_blk4_m96.log 2>&1
tail -2 gpurun_out/prof_blk4_m96.log; ls -la gpurun_out/prof_blk4_m96.ncu-rep
