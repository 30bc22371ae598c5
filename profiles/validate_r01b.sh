#!/bin/bash
# mid-round validation on one B200 after the complex (hpddm_b200z_*) instantiation: whole GPU suite (no -x, so that one
# failure does not hide the rest), smoke, real default bench (no CPU arm: timed by the round-end driver), complex bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/gpu.txt 2>&1
timeout 660 python -m pytest tests -m gpu -q --durations=12 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 240 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_m128.json 2> gpurun_out/bench_m128.err; tail -c 1500 gpurun_out/bench_m128.json
timeout 240 python bench.py --scalar z --cells 64 --steps 10 > gpurun_out/bench_z64.json 2> gpurun_out/bench_z64.err; tail -c 1500 gpurun_out/bench_z64.json; tail -3 gpurun_out/bench_z64.err
