#!/bin/bash
# end-of-round validation on one B200: GPU tests, smoke, default bench (both arms), apply launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python profiles/summarize.py gpurun_out/bench_default.json gpurun_out/bench_reference.json
bash profiles/capture_apply.sh 128
