#!/bin/bash
# 2-GPU check after the scalar-generic refactor: NCCL halo + coarse all-gather parity (real and complex instantiation),
# then the weak-scaling bench line at 2 GPUs.
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29511 tests/run_multi_gpu_parity.py > gpurun_out/parity2_real.log 2>&1; echo "parity real rc=$?"; grep -E "^rank|rror" gpurun_out/parity2_real.log | cut -c1-300 | head -4
PARITY_SCALAR=z timeout 300 $RUN --master-port 29512 tests/run_multi_gpu_parity.py > gpurun_out/parity2_complex.log 2>&1; echo "parity complex rc=$?"; grep -E "^rank|rror" gpurun_out/parity2_complex.log | cut -c1-300 | head -4
PARITY_NONUNIFORM=1 HPDDM_B200_HALO=p2p timeout 300 $RUN --master-port 29513 tests/run_multi_gpu_parity.py > gpurun_out/parity2_p2p.log 2>&1; echo "parity p2p nonuniform rc=$?"; grep -E "^rank|rror" gpurun_out/parity2_p2p.log | cut -c1-300 | head -4
PARITY_SCALAR=z HPDDM_B200_HALO=p2p timeout 300 $RUN --master-port 29514 tests/run_multi_gpu_parity.py > gpurun_out/parity2_p2p_complex.log 2>&1; echo "parity p2p complex rc=$?"; grep -E "^rank|rror" gpurun_out/parity2_p2p_complex.log | cut -c1-300 | head -4
timeout 400 $RUN --master-port 29515 bench.py --gpus 2 --cells 96 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_m96.json 2> gpurun_out/bench_2gpu_m96.err; cut -c1-600 gpurun_out/bench_2gpu_m96.json
