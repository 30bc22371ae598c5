#!/bin/bash
# validation after the device BGMRES: whole GPU suite, then full solves (GMRES vs BGMRES, 4 right-hand sides) on 2 x 1 x 1 ... one GPU
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --durations=6 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --rhs 4 --cells 96 --steps 10 --no-cpu-baseline --krylov > gpurun_out/bench_m96_rhs4_krylov.json 2> gpurun_out/bench_m96_rhs4_krylov.err; tail -c 600 gpurun_out/bench_m96_rhs4_krylov.json; tail -3 gpurun_out/bench_m96_rhs4_krylov.err
