#!/bin/bash
# GEN-templated tuned kernels: suite, then A/B against the library of commit dc60321 on the same box
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/ap_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/ap_pytest.log
tail -5 gpurun_out/ap_pytest.log
timeout 300 python tools/bench_stages.py --level 6 > gpurun_out/ap_bench_stages.jsonl 2> gpurun_out/ap_bench_stages.err
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-config5"
cp trixi.jl_b200/libtrixi_b200.so /tmp/new.so
timeout 600 $B > gpurun_out/ap_new1.json 2> gpurun_out/ap_new1.err
cp tools/ab/libtrixi_b200_old.so trixi.jl_b200/libtrixi_b200.so
timeout 600 $B > gpurun_out/ap_old1.json 2> gpurun_out/ap_old1.err
cp /tmp/new.so trixi.jl_b200/libtrixi_b200.so
timeout 600 $B > gpurun_out/ap_new2.json 2> gpurun_out/ap_new2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ap_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e9,3), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"],4), round(d["roofline"]["frac"],3), round(d["roofline_interface_kernel"]["avg_launch_ms"],4), d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"))
    except Exception as e:
        print(f, "failed", e)
for l in open("gpurun_out/ap_bench_stages.jsonl"):
    d=json.loads(l); print(d["integrator"], d["fused_stage"], round(d["ms_per_step"],3))
PY
