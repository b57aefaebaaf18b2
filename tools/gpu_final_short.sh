mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench.json
timeout 100 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 120 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:anm_env_kernel -c 1 -f -o gpurun_out/prof python tools/rollout_ncu.py 4096 50 > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2000 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 rc=$?"
