#!/bin/bash
# Stand-alone micro-benchmarks (run them on a GPU box with gpurun): instruction latencies, projection in isolation.
mkdir -p tools/bin
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/lat_probe.cu -o tools/bin/lat_probe
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/proj_probe.cu -o tools/bin/proj_probe
