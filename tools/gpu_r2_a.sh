#!/bin/bash
# round 2, GPU call A: correctness of the second-generation headline kernel, A/B bench, ncu capture
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tree_3d_euler_ec or tuned or fused_cfl or pipelined or full_size or conservation_large or free_stream" > gpurun_out/a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/a_bench_v8.json 2> gpurun_out/a_bench_v8.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --kernel-path 2 > gpurun_out/a_bench_v7.json 2> gpurun_out/a_bench_v7.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/a_bench_v8b.json 2> gpurun_out/a_bench_v8b.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_element_euler3d_ranocha_p3 -s 6 -c 2 -o gpurun_out/a_prof_v8 python bench.py --level 6 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/a_ncu.log 2>&1
tail -3 gpurun_out/a_ncu.log
python - <<'PY'
import json
for n in ("v8","v7","v8b"):
    try:
        d=json.load(open(f"gpurun_out/a_bench_{n}.json"))
        print(n, d["value"]/1e9, d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["e2e"]["value"]/1e9)
    except Exception as e:
        print(n, "failed", e)
PY
