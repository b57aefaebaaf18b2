#!/bin/bash
# ncu source-level sampling of a launch that contains ONLY divergent instances (32 of them)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:anm_env_kernel -s 48 -c 1 -f -o gpurun_out/strag_real python tools/straggler_real.py > gpurun_out/strag_real.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/strag_real.log
