#!/bin/bash
# ncu of the headline interface kernel (single-copy face fluxes) at level 7
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-config5"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_interface_flux_staged -s 4 -c 1 -o gpurun_out/ab_prof_if $B > gpurun_out/ab_ncu.log 2>&1
tail -3 gpurun_out/ab_ncu.log
