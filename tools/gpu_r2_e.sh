#!/bin/bash
# full GPU test-suite without -x (collect every failure and the whole parity table)
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/e_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/e_pytest.log
tail -40 gpurun_out/e_pytest.log
