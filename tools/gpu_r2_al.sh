#!/bin/bash
# GLM-MHD shock capturing (tree + structured goldens), whole suite
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "mhd_ec_shockcapturing" > gpurun_out/al_pytest_mhdsc.log 2>&1
tail -30 gpurun_out/al_pytest_mhdsc.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/al_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/al_pytest.log
tail -8 gpurun_out/al_pytest.log
