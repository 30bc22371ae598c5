#!/bin/bash
# A/B of the block (2 / 4 right-hand sides) SpTRSV kernels: first-generation (HPDDM_B200_BLK=staged) vs current
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zcomplex.py tests/test_gpu_properties.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
for v in staged new; do for mu in 2 4; do
  HPDDM_B200_BLK=$v timeout 200 python bench.py --rhs $mu --cells 96 --steps 10 --no-cpu-baseline > gpurun_out/blk_${v}_mu$mu.json 2> gpurun_out/blk_${v}_mu$mu.err
  python - <<PY
import json
d=json.load(open("gpurun_out/blk_${v}_mu$mu.json"))
print("real m=96 $v mu=$mu: sptrsv %.3f ms  frac %.3f  apply %.3f ms" % (d["roofline"]["ms"], d["roofline"]["frac"], d["ms_per_step"]))
PY
done; done
for v in staged new; do
  HPDDM_B200_BLK=$v timeout 200 python bench.py --scalar z --rhs 4 --cells 64 --steps 10 > gpurun_out/blk_z_${v}_mu4.json 2> gpurun_out/blk_z_${v}_mu4.err
  python - <<PY
import json
d=json.load(open("gpurun_out/blk_z_${v}_mu4.json"))
print("complex m=64 $v mu=4: sptrsv %.3f ms  frac %.3f  apply %.3f ms" % (d["roofline"]["ms"], d["roofline"]["frac"], d["ms_per_step"]))
PY
done
