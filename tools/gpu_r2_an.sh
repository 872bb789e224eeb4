#!/bin/bash
# final build: smoke, default bench line, reference arm, ncu launch list of the bench command, sanitizer probe
set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/an_smoke.log 2>&1; tail -2 gpurun_out/an_smoke.log
timeout 900 python bench.py > gpurun_out/an_bench_default.json 2> gpurun_out/an_bench_default.err; tail -c 600 gpurun_out/an_bench_default.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/an_bench_reference.json 2> gpurun_out/an_bench_reference.err; tail -c 400 gpurun_out/an_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/an_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-config5 > gpurun_out/an_launches_bench.log 2>&1
for t in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $t python tools/sanitize_probe.py > gpurun_out/an_sanitize_$t.log 2>&1
  tail -3 gpurun_out/an_sanitize_$t.log
done
