#!/bin/bash
# warp-per-element indicator kernel: whole GPU suite, SC bench, launch list, smoke
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/af_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/af_pytest.log
tail -6 gpurun_out/af_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/af_smoke.log 2>&1; tail -2 gpurun_out/af_smoke.log
B="python bench.py --level 6 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-config5"
timeout 600 $B --workload euler_sc > gpurun_out/af_bench_sc.json 2> gpurun_out/af_bench_sc.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/af_launches_sc.csv $B --workload euler_sc --steps 2 --warmup 1 > gpurun_out/af_launches.log 2>&1
python - <<'PY'
import json,glob,csv,collections
for f in sorted(glob.glob("gpurun_out/af_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("af_bench_")[1], round(d["value"]/1e9,2), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"],3), round(d["roofline"]["frac"],3), round(d["roofline_interface_kernel"]["avg_launch_ms"],3), d["clocks"]["sm_mhz"], d["finite"])
    except Exception as e:
        print(f, "failed", e)
rows=list(csv.DictReader(l for l in open('gpurun_out/af_launches_sc.csv') if l.startswith('"')))
agg=collections.defaultdict(list)
for r in rows: agg[r["Kernel Name"][:80]].append(float(r["Metric Value"])/1e6)
for k,v in agg.items(): print(len(v), round(sum(v)/len(v),3), k)
PY
