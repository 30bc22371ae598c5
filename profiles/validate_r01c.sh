#!/bin/bash
# validation after the multi-RHS device GMRES / device CG: whole GPU suite, then the complex SpTRSV measurements
# (live bench at m = 96 and an ncu launch list with DRAM bytes per sweep kernel, graphs off).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --scalar z --cells 96 --steps 10 > gpurun_out/bench_z96.json 2> gpurun_out/bench_z96.err; tail -c 1200 gpurun_out/bench_z96.json; tail -2 gpurun_out/bench_z96.err
HPDDM_B200_NO_GRAPH=1 timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_fwd|k_bwd" -s 30 -c 30 --csv --log-file gpurun_out/launches_sptrsv_z96.csv python profiles/run_solve.py 96 2 0 z > gpurun_out/ncu_z96.log 2>&1
tail -2 gpurun_out/ncu_z96.log; wc -l gpurun_out/launches_sptrsv_z96.csv
timeout 200 python bench.py --rhs 4 --cells 96 --steps 10 --no-cpu-baseline --krylov > gpurun_out/bench_m96_rhs4.json 2> gpurun_out/bench_m96_rhs4.err; tail -c 700 gpurun_out/bench_m96_rhs4.json
