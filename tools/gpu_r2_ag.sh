#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "chandrashekar or nonconforming_shock or tuned_kernels or shima or kennedy" > gpurun_out/ag_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/ag_pytest.log
tail -12 gpurun_out/ag_pytest.log
