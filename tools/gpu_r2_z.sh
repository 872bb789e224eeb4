#!/bin/bash
# ncu launch list of the default bench command (per-launch durations, cold-cache and serialised), sanitizer probe,
# secondary workload lines with the final build
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-config5 > gpurun_out/z_launches_bench.log 2>&1
for w in euler_weak structured_curved p4est_curved structured_ec p4est_ec mhd_ec euler_sc; do
  timeout 600 python bench.py --workload $w --level 6 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-config5 > gpurun_out/z_bench_$w.json 2> gpurun_out/z_bench_$w.err
done
timeout 600 python bench.py --workload p4est_tgv_p5 --level 5 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-config5 > gpurun_out/z_bench_p4est_tgv_p5.json 2> gpurun_out/z_bench_p4est_tgv_p5.err
for t in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $t python tools/sanitize_probe.py > gpurun_out/z_sanitize_$t.log 2>&1
  tail -3 gpurun_out/z_sanitize_$t.log
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/z_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("z_bench_")[1], round(d["value"]/1e9,2), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"],3), round(d["roofline"]["frac"],3), round(d["roofline_interface_kernel"]["avg_launch_ms"],3), d["clocks"]["sm_mhz"], d["finite"])
    except Exception as e:
        print(f, "failed", e)
PY
