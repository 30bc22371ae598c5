#!/bin/bash
# A/B: MU = 1 forward sweep, shared-staged (default) vs right-hand side through L1 (HPDDM_B200_FWD1=l1); complex MU = 4 forward variants
mkdir -p gpurun_out
HPDDM_B200_FWD1=l1 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "local_solve or apply" 2>&1 | tail -1
for v in default l1; do for m in 128 64; do
  HPDDM_B200_FWD1=$v timeout 200 python bench.py --cells $m --steps 20 --no-cpu-baseline > gpurun_out/fwd1_${v}_m$m.json 2> gpurun_out/fwd1_${v}_m$m.err
  python - <<PY
import json
d=json.load(open("gpurun_out/fwd1_${v}_m$m.json"))
print("m=$m fwd1=$v: sptrsv %.3f ms  frac %.3f  apply %.3f ms" % (d["roofline"]["ms"], d["roofline"]["frac"], d["ms_per_step"]))
PY
done; done
for v in l1 wide; do
  HPDDM_B200_BLK=$v timeout 200 python bench.py --scalar z --rhs 4 --cells 64 --steps 10 > gpurun_out/blk_z2_${v}.json 2> gpurun_out/blk_z2_${v}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/blk_z2_${v}.json"))
print("complex m=64 mu=4 fwd=$v: sptrsv %.3f ms  frac %.3f" % (d["roofline"]["ms"], d["roofline"]["frac"]))
PY
done
