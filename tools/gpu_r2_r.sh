#!/bin/bash
# P4est mortars: GPU parity of the new cases, then the whole GPU suite
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "nonconforming" > gpurun_out/r_pytest_nc.log 2>&1
echo "pytest exit $?" >> gpurun_out/r_pytest_nc.log
tail -15 gpurun_out/r_pytest_nc.log
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r_pytest.log
tail -8 gpurun_out/r_pytest.log
