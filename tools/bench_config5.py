#!/usr/bin/env python
"""BASELINE config 5: ANM6Easy-v0, 16 384 instances over 1 / 2 / 4 / 8 GPUs, actions from the MPC-constant agent
(examples/mpc_constant.py: planning_steps = 10, safety_margin = 0.96).

    python tools/bench_config5.py [--envs-global 16384] [--steps 4] [--planning-steps 10] [--workers N]
    torchrun --nproc-per-node N tools/bench_config5.py ...        (instances sharded over the ranks)

A step = state D2H -> one DC-OPF LP per instance on the host (HiGHS through SciPy, spread over `workers` processes per
rank: gym_anm_b200.agents.MPCAgentConstant(workers=...)) -> actions H2D -> one batched step on the GPU.  The policy is
closed-loop, so the launches are fully ordered.  Prints one JSON line: whole-job env-steps/s, the share of the host
LP, and a sample of the same actions replayed through the C oracle (parity of the stepped batch).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs-global", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--planning-steps", type=int, default=10)
    ap.add_argument("--workers", type=int, default=0, help="LP worker processes per rank (0: usable cores / world)")
    ap.add_argument("--check", type=int, default=64, help="instances replayed through the C oracle")
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    from gym_anm_b200.agents import MPCAgentConstant
    from gym_anm_b200.anm6 import BatchedANM6Easy
    from gym_anm_b200.distributed import shard_slice

    sl = shard_slice(a.envs_global, rank, world)
    B = sl.stop - sl.start
    cores = len(os.sched_getaffinity(0))
    workers = a.workers or max(1, cores // world)
    env = BatchedANM6Easy(B, device=torch.device("cuda", local), env_offset=sl.start, validate_actions=False)
    env.reset(seed=2020)
    agent = MPCAgentConstant(env.simulator, env.action_space, env.gamma, safety_margin=0.96,
                             planning_steps=a.planning_steps, workers=workers)
    agent.act(env)  # untimed: starts the worker processes
    n_chk = min(a.check, B)
    import anm_oracle

    cpu = anm_oracle.OracleEnv(env.spec, n_chk)  # the checker: the same actions through the C restatement
    soc, aux, term = env.native.get_state()
    cpu.soc[:], cpu.aux[:], cpu.terminated[:] = soc.cpu().numpy()[:n_chk], aux.cpu().numpy()[:n_chk], term.cpu().numpy()[:n_chk]
    t_lp = t_step = 0.0
    worst = 0.0
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for t in range(a.steps):
        t1 = time.perf_counter()
        act = agent.act(env)  # D2H of the state + the LPs
        t2 = time.perf_counter()
        obs, r, d, _, _ = env.step(act)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        t_lp, t_step = t_lp + (t2 - t1), t_step + (t3 - t2)
        o_c, r_c, d_c, _ = cpu.step(act[:n_chk])
        assert np.array_equal(d.cpu().numpy()[:n_chk], d_c)
        worst = max(worst, float(np.max(np.abs(obs.cpu().numpy()[:n_chk] - o_c) / np.maximum(np.abs(o_c), 1.0))))
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    agent.close()
    tt = torch.tensor([wall, t_lp, t_step], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        wall, t_lp, t_step = (float(v) for v in tt)
        print(json.dumps({
            "config": "BASELINE config 5: ANM6Easy-v0, %d instances on %d GPU(s), MPC-constant actions (planning_steps %d, "
                      "safety_margin 0.96), one HiGHS LP per instance and step on %d worker processes per rank (%d host cores)"
                      % (a.envs_global, world, a.planning_steps, workers, cores),
            "value": a.envs_global * a.steps / wall, "unit": "env-steps/s", "n_gpus": world, "steps": a.steps,
            "s_per_step": wall / a.steps, "host_lp_s_per_step": t_lp / a.steps, "gpu_step_s_per_step": t_step / a.steps,
            "lp_per_s": a.envs_global * a.steps / t_lp,
            "oracle_check": {"instances": n_chk, "max_rel_err_obs": worst, "terminated_equal": True}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
