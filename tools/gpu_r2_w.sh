#!/bin/bash
# shock capturing on curved meshes: whole GPU suite
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/w_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/w_pytest.log
tail -12 gpurun_out/w_pytest.log
