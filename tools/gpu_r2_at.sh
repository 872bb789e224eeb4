#!/bin/bash
# final: whole suite (parity table), default bench line, ncu --set full of the LEAN headline kernel at level 6 and 7
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/at_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/at_pytest.log
tail -4 gpurun_out/at_pytest.log
timeout 900 python bench.py > gpurun_out/at_bench_default.json 2> gpurun_out/at_bench_default.err
B="python bench.py --no-cpu-baseline --e2e-steps 1 --no-config5 --steps 2 --warmup 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_element_euler3d_ranocha_p3 -s 6 -c 1 -o gpurun_out/at_prof_lean_l6 $B --level 6 > gpurun_out/at_ncu_l6.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_element_euler3d_ranocha_p3 -s 6 -c 1 -o gpurun_out/at_prof_lean_l7 $B --level 7 > gpurun_out/at_ncu_l7.log 2>&1
ls -la gpurun_out/at_prof_*.ncu-rep
python - <<'PY'
import json
d=json.loads(open("gpurun_out/at_bench_default.json").read().strip().splitlines()[-1])
print(round(d["value"]/1e9,3), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"],4), round(d["roofline"]["frac"],3), d["e2e"]["value"]/1e9, d["cpu_baseline"]["value"]/1e6, d["clocks"])
PY
