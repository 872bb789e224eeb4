#!/bin/bash
# MPI mortars: in-process multi-rank GPU tests (1 GPU), then the whole suite
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "halo_exchange or reproduces_reference_golden" > gpurun_out/t_pytest_halo.log 2>&1
echo "pytest exit $?" >> gpurun_out/t_pytest_halo.log
tail -25 gpurun_out/t_pytest_halo.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/t_pytest.log
tail -8 gpurun_out/t_pytest.log
