#!/bin/bash
# forward-sweep kernel variants at m = 128 / 96 (HPDDM_B200_FWD_VAR = 0 baseline, 1 = 64 regs / 4 CTAs per SM, 2 = 32 accumulators per lane)
mkdir -p gpurun_out
for m in 128 96; do for v in 0 1 2; do
  HPDDM_B200_FWD_VAR=$v python bench.py --cells $m --steps 10 --no-cpu-baseline > gpurun_out/fwdvar_${m}_$v.json 2> gpurun_out/fwdvar_${m}_$v.err
done; done
python profiles/summarize.py gpurun_out/fwdvar_*.json
