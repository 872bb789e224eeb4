#!/bin/bash
# LEAN instantiation of the headline kernel: parity subset, then A/B against the previous library on the same box
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tree_3d_euler_ec or tuned_kernels or halo_exchange or property or pipelined" > gpurun_out/as_pytest.log 2>&1
tail -4 gpurun_out/as_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-config5"
cp trixi.jl_b200/libtrixi_b200.so /tmp/new.so
timeout 600 $B > gpurun_out/as_new1.json 2> gpurun_out/as_new1.err
cp tools/ab/libtrixi_b200_old.so trixi.jl_b200/libtrixi_b200.so
timeout 600 $B > gpurun_out/as_old1.json 2> gpurun_out/as_old1.err
cp /tmp/new.so trixi.jl_b200/libtrixi_b200.so
timeout 600 $B > gpurun_out/as_new2.json 2> gpurun_out/as_new2.err
cp tools/ab/libtrixi_b200_old.so trixi.jl_b200/libtrixi_b200.so
timeout 600 $B > gpurun_out/as_old2.json 2> gpurun_out/as_old2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/as_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e9,3), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"],4), round(d["roofline"]["frac"],3), round(d["roofline_interface_kernel"]["avg_launch_ms"],4), d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"))
    except Exception as e:
        print(f, "failed", e)
PY
