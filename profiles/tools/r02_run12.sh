#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_golden.py -q -m gpu -x > gpurun_out/r02_run12_pytest.log 2>&1
tail -3 gpurun_out/r02_run12_pytest.log | cut -c1-300
run() {  # name, workload args..., then env after --
  name=$1; shift
  args=(); while [ "$1" != "--" ]; do args+=("$1"); shift; done; shift
  env "$@" timeout 900 python bench.py "${args[@]}" --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_$name.json 2> gpurun_out/r02_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_$name.json"))
    print("$name: apply %.4f ms  sptrsv %.4f ms  frac %.3f (strict %.3f)" % (d["ms_per_step"], d["roofline"]["ms"], d["roofline"]["frac"], d["roofline"]["frac_strict_nnz"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r02_$name.err").read()[-1500:])
PY
}
run narrow1_m128 --cells 128 -- HPDDM_B200_NARROW=1
run narrow0_m128 --cells 128 -- HPDDM_B200_NARROW=0
run narrow1_m64 --cells 64 -- HPDDM_B200_NARROW=1
run narrow0_m64 --cells 64 -- HPDDM_B200_NARROW=0
run narrow1_m160 --cells 160 -- HPDDM_B200_NARROW=1
run narrow0_m160 --cells 160 -- HPDDM_B200_NARROW=0
