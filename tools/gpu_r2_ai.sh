#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/ai_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/ai_pytest.log
tail -25 gpurun_out/ai_pytest.log
