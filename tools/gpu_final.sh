#!/bin/bash
# last pass of the round: tests, smoke, both bench arms with the driver's arguments, ncu of three lock-step launches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2>/dev/null; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench.json")); r = json.load(open("gpurun_out/r02_bench_reference.json"))
print("value %.4g | isolated %.4g | chained %.4g | lockstep %.4g | e2e %.4g | sync %.4g | ref %.4g (%s) | clocks %s" % (
    d["value"], d["isolated_block"]["value"], d["chained"]["value"], d["lockstep"]["value"], d["e2e"]["value"],
    d["sync_every_step"]["value"], r["value"], r["cpu_baseline"]["kind"], d["clocks"]))
PY
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:anm_env_kernel -c 3 -f \
    -o gpurun_out/r02_step_prof python tools/step_ncu.py 4096 > gpurun_out/r02_step_ncu.log 2>&1; echo "ncu rc=$?"; grep "iteration cap" gpurun_out/r02_step_ncu.log
