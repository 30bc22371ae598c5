#!/bin/bash
# last validation of the round on one B200: whole GPU suite, smoke, default headline bench (CPU arm left to the round-end driver)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 150 python bench.py --no-cpu-baseline > gpurun_out/final_bench_m128.json 2> gpurun_out/final_bench_m128.err; cut -c1-200 gpurun_out/final_bench_m128.json; python -c "
import json; d=json.load(open('gpurun_out/final_bench_m128.json')); print('apply %.3f ms e2e %.3f ms sptrsv %.3f ms frac %.3f launches %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['ms'], d['roofline']['frac'], d['gpu_launches']), d['clocks'])"
