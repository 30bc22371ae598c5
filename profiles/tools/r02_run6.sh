#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
./profiles/tools/bin/dmma_bench | tee gpurun_out/r02_dmma_bench.txt
HPDDM_B200_PERSISTENT=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_zcomplex.py -q -m gpu -x -k "not cholesky_breakdown" > gpurun_out/r02_run6_pytest.log 2>&1
tail -4 gpurun_out/r02_run6_pytest.log
run() {  # name, workload args..., then env after --
  name=$1; shift
  args=(); while [ "$1" != "--" ]; do args+=("$1"); shift; done; shift
  env "$@" timeout 600 python bench.py "${args[@]}" --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_$name.json 2> gpurun_out/r02_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_$name.json"))
    print("$name: apply %.3f ms  sptrsv %.3f ms  frac %.3f" % (d["ms_per_step"], d["roofline"]["ms"], d["roofline"]["frac"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r02_$name.err").read()[-2500:])
PY
}
run dmma2_m96_mu8 --cells 96 --rhs 8 -- HPDDM_B200_MMA=1
run levels_m64 --cells 64 -- HPDDM_B200_PERSISTENT=0
run persist_m64 --cells 64 -- HPDDM_B200_PERSISTENT=1
run levels_m128 --cells 128 -- HPDDM_B200_PERSISTENT=0
run persist_m128 --cells 128 -- HPDDM_B200_PERSISTENT=1
