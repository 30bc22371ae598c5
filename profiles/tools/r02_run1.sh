#!/bin/bash
# round 2, call 1: both bench arms at the default (largest-fit) size on one box + the new host-boundary test
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pageable or apply" > gpurun_out/r02_run1_pytest.log 2>&1
tail -3 gpurun_out/r02_run1_pytest.log
( time python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02_ref_m160.json ) 2> gpurun_out/r02_ref_m160.err
tail -c 1500 gpurun_out/r02_ref_m160.json; tail -4 gpurun_out/r02_ref_m160.err
( time python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_m160.json ) 2> gpurun_out/r02_bench_m160.err
cat gpurun_out/r02_bench_m160.json; tail -5 gpurun_out/r02_bench_m160.err
