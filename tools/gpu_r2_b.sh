#!/bin/bash
# round 2, GPU call B: streamed-tile headline kernel (16 CTAs/SM, L2 reduce-add update): tests, A/B bench, ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tree_3d_euler_ec or tree_3d_euler_source_terms or tuned or fused_cfl or pipelined or full_size or conservation_large or free_stream or step_2n or solve_2n or golden" > gpurun_out/b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/b_pytest.log
tail -5 gpurun_out/b_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 600 $B > gpurun_out/b_bench_v9.json 2> gpurun_out/b_bench_v9.err
timeout 600 $B --no-reduce-update > gpurun_out/b_bench_v9_noreduce.json 2> gpurun_out/b_bench_v9_noreduce.err
timeout 600 $B --kernel-path 2 > gpurun_out/b_bench_v7.json 2> gpurun_out/b_bench_v7.err
timeout 600 $B > gpurun_out/b_bench_v9b.json 2> gpurun_out/b_bench_v9b.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_element_euler3d_ranocha_p3 -s 6 -c 2 -o gpurun_out/b_prof_v9 python bench.py --level 6 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/b_ncu.log 2>&1
tail -3 gpurun_out/b_ncu.log
python - <<'PY'
import json
for n in ("v9","v9_noreduce","v7","v9b"):
    try:
        d=json.load(open(f"gpurun_out/b_bench_{n}.json"))
        print(n, d["value"]/1e9, d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["roofline_interface_kernel"]["avg_launch_ms"], d["clocks"]["sm_mhz"], d["e2e"]["value"]/1e9)
    except Exception as e:
        print(n, "failed", e)
PY
