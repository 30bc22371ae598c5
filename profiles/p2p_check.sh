#!/bin/bash
# 2-GPU check of the NVLink peer-memory halo against the NCCL send/recv path (parity + latency at a small size)
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
HPDDM_B200_HALO=p2p HPDDM_B200_DEBUG=1 timeout 300 $RUN --master-port 29511 tests/run_multi_gpu_parity.py > gpurun_out/p2p_parity.log 2>&1; echo "parity p2p rc=$?"; grep -E "^rank|peer-memory|rror" gpurun_out/p2p_parity.log | cut -c1-220 | head -6
HPDDM_B200_HALO=nccl timeout 300 $RUN --master-port 29512 tests/run_multi_gpu_parity.py > gpurun_out/nccl_parity.log 2>&1; echo "parity nccl rc=$?"
for mode in p2p nccl; do for m in 48 128; do
  HPDDM_B200_HALO=$mode timeout 600 $RUN --master-port 29513 bench.py --gpus 2 --cells $m --steps 20 --warmup 3 > gpurun_out/halo_${mode}_$m.json 2> gpurun_out/halo_${mode}_$m.err
done; done
python profiles/summarize.py gpurun_out/halo_*.json
