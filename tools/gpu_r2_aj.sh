#!/bin/bash
# fused 3S*/SSP stage updates: new tests, whole suite, stage bench
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fused_3sstar or other_integrators" -x > gpurun_out/aj_pytest_stage.log 2>&1
tail -15 gpurun_out/aj_pytest_stage.log
timeout 300 python tools/bench_stages.py --level 6 > gpurun_out/aj_bench_stages.jsonl 2> gpurun_out/aj_bench_stages.err
cat gpurun_out/aj_bench_stages.jsonl; tail -3 gpurun_out/aj_bench_stages.err
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/aj_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/aj_pytest.log
tail -8 gpurun_out/aj_pytest.log
