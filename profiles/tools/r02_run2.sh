#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
MASTER_ADDR=127.0.0.1 PARITY_BOOT=host PARITY_SAME_GPU=1 PARITY_M=8 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 tests/run_multi_gpu_parity.py > gpurun_out/r02_run2_dbg.log 2>&1
grep -v "^\[W\|^W1\|^\*\*\*\|Setting OMP" gpurun_out/r02_run2_dbg.log | head -60
