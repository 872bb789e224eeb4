#!/bin/bash
# whole GPU suite with P4est mortars, MHD mortars (noncons) and flux_hlle
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/s_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s_pytest.log
tail -12 gpurun_out/s_pytest.log
