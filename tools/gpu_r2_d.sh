#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/diag_headline.py > gpurun_out/d_diag.log 2>&1
cat gpurun_out/d_diag.log
