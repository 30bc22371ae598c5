#!/bin/bash
# multi-GPU: parity of every hot-path entry point + device Krylov drivers against the oracle, and the bench, over both transports
N=$1; shift
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
nvidia-smi topo -m > gpurun_out/r02_topo_$N.txt 2>&1
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) "$@"; }
par() {  # name, env...
  name=$1; shift
  env "$@" PARITY_M=8 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) tests/run_multi_gpu_parity.py > gpurun_out/r02_parity${N}_$name.log 2>&1
  echo "parity $name rc=$?: $(grep -c ' OK' gpurun_out/r02_parity${N}_$name.log) OK / $(grep -c 'FAIL' gpurun_out/r02_parity${N}_$name.log) FAIL; $(grep -m1 -o '\[[a-z -]*\]' gpurun_out/r02_parity${N}_$name.log)"
}
ben() {  # name, cells, env...   (BENCH_EXTRA: extra bench.py arguments)
  name=$1; cells=$2; shift 2
  env "$@" timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --cells $cells --steps 10 --warmup 3 --no-cpu-baseline $BENCH_EXTRA > gpurun_out/r02_scale${N}_$name.json 2> gpurun_out/r02_scale${N}_$name.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/r02_scale${N}_$name.json") if l.startswith("{")][-1]
    print("bench N=$N $name: %.3f ms/apply  value %.2f  sptrsv %.3f ms frac %.3f  phases %s" % (d["ms_per_step"], d["value"], d["roofline"]["ms"], d["roofline"]["frac"], {k: round(v, 4) for k, v in d["phases"].items()}))
except Exception as e:
    print("bench N=$N $name failed", e); print(open("gpurun_out/r02_scale${N}_$name.err").read()[-2000:])
PY
}
for job in "$@"; do
  case $job in
    parity) par fabric PARITY_EXPECT_TRANSPORT="peer-memory fabric"; par nccl HPDDM_B200_HALO=nccl PARITY_EXPECT_TRANSPORT=nccl; par hostboot PARITY_BOOT=host PARITY_EXPECT_TRANSPORT="peer-memory fabric";;
    parityz) par fabric_z PARITY_SCALAR=z; par nonuniform PARITY_NONUNIFORM=1;;
    b160) ben fabric_m160 160 HPDDM_B200_BENCH=1;;
    b160nccl) ben nccl_m160 160 HPDDM_B200_HALO=nccl;;
    b64) ben fabric_m64 64 HPDDM_B200_BENCH=1; ben nccl_m64 64 HPDDM_B200_HALO=nccl;;
    b96) ben fabric_m96 96 HPDDM_B200_BENCH=1; ben nccl_m96 96 HPDDM_B200_HALO=nccl;;
    elas) BENCH_EXTRA="--workload elasticity --rhs 4 --nvec 30" ben elasticity_q64_mu4 64 HPDDM_B200_BENCH=1;;
    z96) BENCH_EXTRA="--scalar z" ben helmholtz_z96 96 HPDDM_B200_BENCH=1;;
  esac
done
