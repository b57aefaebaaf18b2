#!/bin/bash
# compute-sanitizer (memcheck / racecheck / synccheck) over a small-size subset of the GPU parity tests.
mkdir -p gpurun_out
SEL="transition_goldens or reset_goldens or golden_trajectory and seed0 or rollout_on_pinned or synth30_env or single_env_facade or list_observation or singular_schur or tensor_level_hooks or simulator_state"
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" \
      -p no:cacheprovider > gpurun_out/sanitize_$tool.log 2>&1
  tail -4 gpurun_out/sanitize_$tool.log
done
