#!/bin/bash
# end-of-session validation on one B200: whole GPU suite, smoke, headline bench A/B of the MU = 1 forward sweep (same box)
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --durations=5 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -14 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for v in default l1; do
  HPDDM_B200_FWD1=$v timeout 200 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/final_${v}_m128.json 2> gpurun_out/final_${v}_m128.err
  python - <<PY
import json
d=json.load(open("gpurun_out/final_${v}_m128.json"))
print("m=128 fwd1=$v: sptrsv %.3f ms  frac %.3f  apply %.3f ms  e2e %.3f ms" % (d["roofline"]["ms"], d["roofline"]["frac"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
PY
done
