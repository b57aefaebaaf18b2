#!/usr/bin/env python
"""BASELINE config 5: ANM6Easy-v0, 16 384 instances over 1 / 2 / 4 / 8 GPUs, actions from the MPC-constant agent
(examples/mpc_constant.py: planning_steps = 10, safety_margin = 0.96).

    python tools/bench_config5.py [--lp gpu|host] [--envs-global 16384] [--steps 64] [--planning-steps 10]
    torchrun --nproc-per-node N tools/bench_config5.py ...        (instances sharded over the ranks, no collective)

`--lp gpu` (default): the whole closed loop stays on the device -- state tensor -> per-instance bounds -> one
warm-started solve of every instance's DC-OPF by the batched dual simplex (include/anm_lp.h, one GPU thread per
program, tableaux resident in HBM) -> action tensor -> one batched step.  One host look per step (the number of
instances whose solution failed its checks).
`--lp host`: state D2H -> one HiGHS LP per instance on the host (SciPy, `workers` processes per rank) -> actions H2D.

The policy is closed-loop, so the launches are fully ordered.  Prints one JSON line: whole-job env-steps/s, the share of
the LP, and a sample of the stepped batch replayed through the C oracle AFTER the timed region (same actions).
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
    ap.add_argument("--lp", choices=("gpu", "host"), default="gpu")
    ap.add_argument("--envs-global", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (0: 64 with --lp gpu, 4 with --lp host)")
    ap.add_argument("--planning-steps", type=int, default=10)
    ap.add_argument("--workers", type=int, default=0, help="--lp host: LP worker processes per rank (0: cores / world)")
    ap.add_argument("--refresh", type=int, default=64, help="--lp gpu: every instance's basis is rebuilt every N solves")
    ap.add_argument("--check", type=int, default=64, help="instances replayed through the C oracle")
    a = ap.parse_args()
    steps = a.steps or (64 if a.lp == "gpu" else 4)
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from gym_anm_b200.agents import MPCAgentConstant
    from gym_anm_b200.anm6 import BatchedANM6Easy
    from gym_anm_b200.distributed import shard_slice

    sl = shard_slice(a.envs_global, rank, world)
    B = sl.stop - sl.start
    cores = len(os.sched_getaffinity(0))
    workers = a.workers or max(1, cores // world)
    env = BatchedANM6Easy(B, device=dev, env_offset=sl.start, validate_actions=False)
    env.reset(seed=2020)
    kw = dict(device=dev, refresh=a.refresh) if a.lp == "gpu" else dict(workers=workers)
    agent = MPCAgentConstant(env.simulator, env.action_space, env.gamma, safety_margin=0.96,
                             planning_steps=a.planning_steps, **kw)
    n_chk = min(a.check, B)
    soc, aux, term = env.native.get_state()
    chk0 = [t[:n_chk].cpu().numpy().copy() for t in (soc, aux, term)]
    # untimed first step: starts the worker processes / builds every instance's first basis (the cold solve)
    torch.cuda.synchronize()
    t_c0 = time.perf_counter()
    act = agent.act(env)
    torch.cuda.synchronize()
    cold_s = time.perf_counter() - t_c0
    act = torch.as_tensor(act, device=dev)
    acts = torch.zeros(steps + 1, n_chk, act.shape[1], dtype=torch.float64, device=dev)
    obss = torch.zeros(steps + 1, n_chk, env.observation_N, dtype=torch.float64, device=dev)
    terms = torch.zeros(steps + 1, n_chk, dtype=torch.bool, device=dev)
    pivots = torch.zeros(steps + 1, dtype=torch.float64, device=dev)
    obs, r, d, _, _ = env.step(act)
    acts[0], obss[0], terms[0] = act[:n_chk], obs[:n_chk], d[:n_chk]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    ev[0].record()
    for t in range(steps):
        act = torch.as_tensor(agent.act(env), device=dev)  # the LPs (host: + D2H of the state, H2D of the actions)
        ev[2 * t + 1].record()
        obs, r, d, _, _ = env.step(act)
        ev[2 * t + 2].record()
        acts[t + 1], obss[t + 1], terms[t + 1] = act[:n_chk], obs[:n_chk], d[:n_chk]
        if a.lp == "gpu":
            pivots[t + 1] = agent._dev.lp.iters.double().mean()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    t_lp = sum(ev[2 * t].elapsed_time(ev[2 * t + 1]) for t in range(steps)) * 1e-3
    t_step = sum(ev[2 * t + 1].elapsed_time(ev[2 * t + 2]) for t in range(steps)) * 1e-3
    mean_reward = float(r.mean())
    # the checker, outside the timed region: the same actions through the C restatement
    import anm_oracle

    cpu = anm_oracle.OracleEnv(env.spec, n_chk)
    cpu.soc[:], cpu.aux[:], cpu.terminated[:] = chk0
    worst, acts_h, obss_h, terms_h = 0.0, acts.cpu().numpy(), obss.cpu().numpy(), terms.cpu().numpy()
    for t in range(steps + 1):
        o_c, r_c, d_c, _ = cpu.step(acts_h[t])
        assert np.array_equal(terms_h[t], d_c), "terminated flags differ from the oracle at step %d" % t
        worst = max(worst, float(np.max(np.abs(obss_h[t] - o_c) / np.maximum(np.abs(o_c), 1.0))))
    assert worst < 1e-8, worst
    stats = dict(agent.lp_stats)
    lp_bytes = agent._dev.lp.bytes if a.lp == "gpu" else 0
    lp_kernel = agent._dev.lp.kernel if a.lp == "gpu" else None
    agent.close()
    tt = torch.tensor([wall, t_lp, t_step, cold_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        wall, t_lp, t_step, cold_s = (float(v) for v in tt)
        how = ("batched dual simplex on the GPU (anm_lp_solve: one %s per program, %d x %d tableau per instance "
               "resident in HBM = %.2f GB per rank, warm-started; basis rebuilt every %d solves)"
               % ((lp_kernel,) + agent_dims(a.planning_steps) + (lp_bytes / 1e9, a.refresh))) if a.lp == "gpu" else (
               "one HiGHS LP per instance and step on %d worker processes per rank (%d host cores)" % (workers, cores))
        print(json.dumps({
            "config": "BASELINE config 5: ANM6Easy-v0, %d instances on %d GPU(s), MPC-constant actions (planning_steps %d, "
                      "safety_margin 0.96); LPs: %s" % (a.envs_global, world, a.planning_steps, how),
            "lp": a.lp, "lp_kernel": lp_kernel, "value": a.envs_global * steps / wall, "unit": "env-steps/s", "n_gpus": world, "steps": steps,
            "s_per_step": wall / steps, "lp_s_per_step": t_lp / steps, "env_step_s_per_step": t_step / steps,
            "lp_per_s": a.envs_global * steps / t_lp, "first_solve_s": cold_s,
            "mean_pivots_per_solve": float(pivots[1:].mean()) if a.lp == "gpu" else None, "lp_stats": stats,
            "mean_reward_last_step": mean_reward,
            "oracle_check": {"instances": n_chk, "steps": steps + 1, "max_rel_err_obs": worst,
                             "terminated_equal": True}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def agent_dims(N, n_load=3, n_gen=2, n_des=1, n_branch=5):
    return (N * (2 * n_branch + 2 * n_des), N * (n_load + n_gen + 2 * n_des + n_branch))


if __name__ == "__main__":
    main()
