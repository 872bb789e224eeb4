#!/bin/bash
# shock capturing with hoisted node records: parity (SC cases + tuned-vs-generic), then A/B bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "shockcapturing or shock_capturing or sedov or blast or tuned_kernels" > gpurun_out/aa_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/aa_pytest.log
tail -8 gpurun_out/aa_pytest.log
B="python bench.py --workload euler_sc --level 6 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-config5"
timeout 600 $B > gpurun_out/aa_bench_sc_rec.json 2> gpurun_out/aa_bench_sc_rec.err
timeout 600 $B --kernel-path 2 > gpurun_out/aa_bench_sc_plain.json 2> gpurun_out/aa_bench_sc_plain.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_element_fd3d_p3 -s 6 -c 1 -o gpurun_out/aa_prof_sc $B --steps 2 --warmup 1 > gpurun_out/aa_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/aa_launches_sc.csv $B --steps 2 --warmup 1 > gpurun_out/aa_launches.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/aa_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("aa_bench_")[1], round(d["value"]/1e9,2), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"],3), round(d["roofline"]["frac"],3), round(d["roofline_interface_kernel"]["avg_launch_ms"],3), d["clocks"]["sm_mhz"], d["finite"])
    except Exception as e:
        print(f, "failed", e)
PY
