#!/bin/bash
# secondary workload lines with the final build (no regressions from the GEN instantiations)
mkdir -p gpurun_out
B="--level 6 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-config5"
for w in euler_weak structured_ec mhd_ec euler_sc euler_shima; do
  timeout 120 python bench.py --workload $w $B > gpurun_out/av_bench_$w.json 2> gpurun_out/av_bench_$w.err
done
timeout 120 python bench.py --workload p4est_tgv_p5 --level 5 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-config5 > gpurun_out/av_bench_p4est_tgv_p5.json 2> gpurun_out/av_bench_p4est_tgv_p5.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/av_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("av_bench_")[1], round(d["value"]/1e9,2), round(d["ms_per_step"],3), round(d["roofline"]["avg_launch_ms"],3), round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz"], d["finite"])
    except Exception as e:
        print(f, "failed", e)
PY
