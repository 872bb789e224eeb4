#!/bin/bash
# GLM-MHD on curved meshes: whole GPU suite
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/x_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/x_pytest.log
tail -30 gpurun_out/x_pytest.log
