#!/bin/bash
# bench.py over subdomain sizes on one GPU (device-resident apply, e2e, SpTRSV roofline)
mkdir -p gpurun_out
for m in ${SIZES:-48 64 96 128}; do
  python bench.py --cells $m --steps 10 --no-cpu-baseline > gpurun_out/size_$m.json 2> gpurun_out/size_$m.err
done
python profiles/summarize.py gpurun_out/size_*.json
