#!/bin/bash
# A/B of block-kernel variants (forward: L1 / staged-wide; backward MU = 4: 2 or 3 CTAs per SM), parity first
mkdir -p gpurun_out
for v in "new:2" "wide:2" "new:3" "wide:3"; do
  blk=${v%%:*}; occ=${v##*:}
  export HPDDM_B200_BLK=$blk HPDDM_B200_BWD4=$occ
  timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "block_solve or config4 or apply" 2>&1 | tail -1
  for mu in 4 2; do
    timeout 200 python bench.py --rhs $mu --cells 96 --steps 10 --no-cpu-baseline > gpurun_out/blk2_${blk}_${occ}_mu$mu.json 2> gpurun_out/blk2_${blk}_${occ}_mu$mu.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/blk2_${blk}_${occ}_mu$mu.json"))
    print("real m=96 fwd=$blk bwd_occ=$occ mu=$mu: sptrsv %.3f ms  frac %.3f  apply %.3f ms" % (d["roofline"]["ms"], d["roofline"]["frac"], d["ms_per_step"]))
except Exception as e:
    print("FAILED $blk $occ $mu", e); print(open("gpurun_out/blk2_${blk}_${occ}_mu$mu.err").read()[-400:])
PY
  done
done
