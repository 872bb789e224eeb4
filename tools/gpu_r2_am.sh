#!/bin/bash
# Euler flux_hlle + further MHD goldens, whole suite
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "hlle or mhd_ec_constant or orszag or llf_naive" > gpurun_out/am_pytest_new.log 2>&1
tail -30 gpurun_out/am_pytest_new.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/am_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/am_pytest.log
tail -8 gpurun_out/am_pytest.log
