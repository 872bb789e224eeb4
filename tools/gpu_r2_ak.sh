#!/bin/bash
# P4est shock capturing across ranks (in-process multi-rank), whole suite
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "halo_exchange_matches_single_rank" -x > gpurun_out/ak_pytest_halo.log 2>&1
tail -25 gpurun_out/ak_pytest_halo.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/ak_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/ak_pytest.log
tail -8 gpurun_out/ak_pytest.log
