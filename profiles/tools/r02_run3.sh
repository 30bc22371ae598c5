#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
( time python -m pytest tests -q -m gpu ) > gpurun_out/r02_run3_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02_run3_pytest.log | cut -c1-250
