#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_properties.py -q -m gpu -x > gpurun_out/r02_run8_pytest.log 2>&1
tail -4 gpurun_out/r02_run8_pytest.log
run() {  # name, workload args..., then env after --
  name=$1; shift
  args=(); while [ "$1" != "--" ]; do args+=("$1"); shift; done; shift
  env "$@" timeout 900 python bench.py "${args[@]}" --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_$name.json 2> gpurun_out/r02_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_$name.json"))
    print("$name: apply %.3f ms  sptrsv %.3f ms  frac %.3f" % (d["ms_per_step"], d["roofline"]["ms"], d["roofline"]["frac"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r02_$name.err").read()[-2500:])
PY
}
run reg16_m96_mu4 --cells 96 --rhs 4 -- HPDDM_B200_MMA=1
run reg16_m96_mu8 --cells 96 --rhs 8 -- HPDDM_B200_MMA=1
run reg16_m96_mu2 --cells 96 --rhs 2 -- HPDDM_B200_MMA=2
run scalar_m96_mu4 --cells 96 --rhs 4 -- HPDDM_B200_MMA=0
run elas64_reg16 --workload elasticity --cells 64 --rhs 4 --nvec 30 -- HPDDM_B200_MMA=1
bash profiles/capture_r02.sh 160
