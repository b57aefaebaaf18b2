#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4000 --warmup 20 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"; cat gpurun_out/bench_n2.json | cut -c1-700; tail -3 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 5 --impl reference --cpu-steps 100 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err; echo "n2ref rc=$?"; cut -c1-300 gpurun_out/bench_n2_ref.json
python bench.py --steps 4000 --no-cpu-baseline 2>/dev/null | cut -c1-200
