#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
( time python -m pytest tests -q -m gpu ) > gpurun_out/r02_run9_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^real" gpurun_out/r02_run9_pytest.log | cut -c1-250
run() {  # name, workload args..., then env after --
  name=$1; shift
  args=(); while [ "$1" != "--" ]; do args+=("$1"); shift; done; shift
  env "$@" timeout 900 python bench.py "${args[@]}" --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_$name.json 2> gpurun_out/r02_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_$name.json"))
    print("$name: apply %.3f ms  e2e %.3f / %.3f / %.3f ms  sptrsv %.3f ms  frac %.3f  launches %d  phases %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["ms_per_step"], d["e2e"]["pageable_unregistered"]["ms_per_step"], d["roofline"]["ms"], d["roofline"]["frac"], d["gpu_launches"], {k: round(v, 4) for k, v in d["phases"].items()}))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r02_$name.err").read()[-2500:])
PY
}
run final_m160 -- HPDDM_B200_BENCH=1
run m64_graph --cells 64 -- HPDDM_B200_APPLY_GRAPH=1
run m64_nograph --cells 64 -- HPDDM_B200_APPLY_GRAPH=0
run m64_spmv_direct --cells 64 -- HPDDM_B200_SPMV=direct
run m32_graph --cells 32 -- HPDDM_B200_APPLY_GRAPH=1
run m32_nograph --cells 32 -- HPDDM_B200_APPLY_GRAPH=0
