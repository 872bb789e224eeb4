#!/bin/bash
# 11 more reference goldens (2D vortex / KHI / structured 2D mappings / p4est 2D SC-EC incl. flux_chandrashekar along normals)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/aq_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/aq_pytest.log
tail -12 gpurun_out/aq_pytest.log
