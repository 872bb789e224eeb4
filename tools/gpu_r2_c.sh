#!/bin/bash
# round 2, GPU call C: full GPU test-suite (new parity cases, flat tolerance), bench, ncu of the headline kernel
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c_pytest.log
tail -15 gpurun_out/c_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 600 $B > gpurun_out/c_bench_v10.json 2> gpurun_out/c_bench_v10.err
timeout 600 $B --kernel-path 2 > gpurun_out/c_bench_v7.json 2> gpurun_out/c_bench_v7.err
timeout 600 $B > gpurun_out/c_bench_v10b.json 2> gpurun_out/c_bench_v10b.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_element_euler3d_ranocha_p3 -s 6 -c 2 -o gpurun_out/c_prof_v10 python bench.py --level 6 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/c_ncu.log 2>&1
tail -3 gpurun_out/c_ncu.log
python - <<'PY'
import json
for n in ("v10","v7","v10b"):
    try:
        d=json.load(open(f"gpurun_out/c_bench_{n}.json"))
        print(n, d["value"]/1e9, d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline_interface_kernel"]["avg_launch_ms"], d["clocks"]["sm_mhz"], d["e2e"]["value"]/1e9)
    except Exception as e:
        print(n, "failed", e)
PY
