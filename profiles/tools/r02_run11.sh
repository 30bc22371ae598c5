#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q -m gpu -x -k "gmv or apply or golden or small_40x40_p4_twolevel or config1" > gpurun_out/r02_run11_pytest.log 2>&1
tail -2 gpurun_out/r02_run11_pytest.log
