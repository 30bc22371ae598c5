#!/bin/bash
# ncu --set full of the tensor-pipe block SpTRSV kernels (4 right-hand sides) on the large levels of one m = 96 solve (graphs off),
# plus the launch list (time + DRAM bytes) of the whole solve
mkdir -p gpurun_out
HPDDM_B200_NO_GRAPH=1 HPDDM_B200_MMA=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fwd_dmma|k_bwd_dmma" -s 41 -c 6 -o gpurun_out/r02_prof_dmma4_m96 -f python profiles/run_solve.py 96 2 0 d 4 > gpurun_out/r02_prof_dmma4_m96.log 2>&1
tail -2 gpurun_out/r02_prof_dmma4_m96.log; ls -la gpurun_out/r02_prof_dmma4_m96.ncu-rep
HPDDM_B200_NO_GRAPH=1 HPDDM_B200_MMA=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_fwd_dmma|k_bwd_dmma|k_perm" --csv --log-file gpurun_out/r02_launches_dmma4_m96.csv python profiles/run_solve.py 96 1 0 d 4 > /dev/null 2>&1
wc -l gpurun_out/r02_launches_dmma4_m96.csv
