#!/bin/bash
# tensor-pipe block SpTRSV: correctness (tests that use 3 / 4 / 8 right-hand sides) + same-box A/B against the register-tiled kernels
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -q -m gpu -x > gpurun_out/r02_run4_pytest.log 2>&1
tail -5 gpurun_out/r02_run4_pytest.log
run() {  # name, rhs, env...
  name=$1; rhs=$2; shift 2
  env "$@" python bench.py --cells 96 --rhs $rhs --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_mma_$name.json 2> gpurun_out/r02_mma_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_mma_$name.json"))
    print("$name: apply %.3f ms  sptrsv %.3f ms  frac %.3f" % (d["ms_per_step"], d["roofline"]["ms"], d["roofline"]["frac"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r02_mma_$name.err").read()[-1500:])
PY
}
run scalar_mu4 4 HPDDM_B200_MMA=0
run dmma_occ3_mu4 4 HPDDM_B200_MMA=1 HPDDM_B200_MMA_OCC=3
run dmma_occ2_mu4 4 HPDDM_B200_MMA=1 HPDDM_B200_MMA_OCC=2
run tma_mu4 4 HPDDM_B200_MMA=1 HPDDM_B200_MMA_VARIANT=tma
run dmma_occ3_mu8 8 HPDDM_B200_MMA=1 HPDDM_B200_MMA_OCC=3
run dmma_occ2_mu8 8 HPDDM_B200_MMA=1 HPDDM_B200_MMA_OCC=2
run scalar_mu2 2 HPDDM_B200_MMA=0
run dmma_occ3_mu2 2 HPDDM_B200_MMA=2 HPDDM_B200_MMA_OCC=3
run scalar_mu1 1 HPDDM_B200_MMA=0
run dmma_occ3_mu1 1 HPDDM_B200_MMA=all HPDDM_B200_MMA_OCC=3
