#!/bin/bash
# leaf-size sweep of the nested-dissection ordering at m = 128 (HPDDM_B200_LEAF), plus a device-resident GMRES solve
mkdir -p gpurun_out
for leaf in 64 128 256 512; do
  HPDDM_B200_LEAF=$leaf python bench.py --cells 128 --steps 10 --no-cpu-baseline > gpurun_out/leaf_$leaf.json 2> gpurun_out/leaf_$leaf.err
done
python bench.py --cells 128 --steps 5 --no-cpu-baseline --krylov > gpurun_out/krylov_128.json 2> gpurun_out/krylov_128.err
python profiles/summarize.py gpurun_out/leaf_*.json gpurun_out/krylov_128.json
