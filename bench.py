#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched gym-anm step() path on N B200s (one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs B] [--config 2|4] [--impl ours|reference]

Workloads (BASELINE.json `configs`):
  --config 2 (default; the configuration the metric is quoted on): ANM6Easy-v0, B = 4096 environment instances PER GPU
             (weak scaling), uniform-random actions over the action box, fp64 Newton-Raphson.
  --config 4: the synthetic 30-bus feeder (gym_anm_b200.networks.synth_feeder_network), B = 8192 instances, random
             actions and caller-supplied next_vars, block-sparse Newton solver.
  (configs 3 and 5 are sizes of config 2: `--envs 8192 --gpus 8`, `--envs 2048 --gpus 8`.)
A "step" is one pass of the hot path over the whole batch: every instance advances by one timestep.  Terminated
instances are re-initialised by the kernel's next-step auto-reset from a pool of initial states known to converge, so
every instance does one full transition per step.

Keys of the line (all rates are whole-job env-steps/s over the N GPUs, timed on the device, max over ranks):
  value, ms_per_step  K steps = one block of `anm_rollout` launches (<= 1000 steps per launch; carried state on chip,
                      every step's obs / reward / terminated row written to [T, B, .] arrays in HBM).  The block is
                      repeated back to back (launches chained per instance) in 5 groups of R blocks until >= 0.3 s are
                      timed; value = median over the groups of R*K*B / t_group.  Every block reads a different window of
                      a 1000-slot action ring (197 MB > the 126 MB L2) and writes its own window of an output ring.
  isolated_block      ONE K-step block timed alone (device idle before and after; median of 9): for small K this is
                      dominated by the slowest warp of the launch (the instances whose Newton iteration ran to 100).
  chained, lockstep   the same workload as ONE kernel launch per step (CUDA graph of 100 steps): launches ordered per
                      instance (open-loop actions) / every launch fully ordered after the previous one -- `lockstep`
                      is the rate a closed-loop policy on the same stream can reach.
  e2e                 the metric through the C ABI's host calls with pinned HOST arrays: `anm_rollout_host_async`
                      (100 steps per call, H2D actions + D2H obs / reward / terminated inside the timed region).
  sync_every_step     the synchronous one-step host call `anm_step_host` (what `env.step(numpy_action)` costs).
  with_obs_allgather  (N > 1) one launch per step whose epilogue stores every [obs | reward | terminated] row into every
                      peer's gather buffer over NVLink (fused step + all-gather, no separate collective), plus the
                      per-step arrival wait; `with_obs_allgather_nccl`: the kernel writes the packed rows locally and
                      ncclAllGather moves them.  `gather_checksum*`: hashes of the gathered batch.
  roofline            hbm leg (what the metric asks for): algorithmic bytes per env-step x env-steps per
                      launch / mean launch duration over the timed region, against MEASURED_PEAKS.json `hbm_gbs`;
                      fp64 leg: algorithmic fp64 FLOP per env-step / time against the DFMA rate measured
                      in this run (`anm_debug_fp64_peak`).  Fields under `ncu_static` are constants copied from
                      profiles/ncu_summary.json (an earlier ncu capture), NOT measurements of this run.
  config5             BASELINE configs[4] beside the headline: 16 384 instances (global, sharded over the N GPUs) in closed
                      loop with the MPC-constant agent (planning_steps 10, safety_margin 0.96) whose DC-OPF programs are
                      solved on the GPU by the batched dual simplex of include/anm_lp.h (state tensor -> bounds -> LPs ->
                      action tensor -> anm_step); device-timed, max over ranks, a sample replayed through the C oracle
                      after the timed region.  `--no-config5` skips it; a failure is reported as {"error": ...} and
                      leaves the rest of the line intact.
  config4             BASELINE configs[3] beside the headline: the synthetic 30-bus feeder, 8192 instances per GPU, 20-step
                      rollout launches back to back (what `--config 4` measures as its `value`, shortened);
                      `--no-config4` skips it; failure-isolated like `config5`.
  cpu_baseline        the reference's own ANM6Easy (`oracle/_ref`, staged by oracle/build_ref.py; kind "reference")
                      on all usable host cores, one env per process; falls back to the NumPy/SciPy port (kind "port").
`--impl reference` times that CPU implementation alone and prints the same line with "impl": "reference".
"""
import argparse
import hashlib
import json
import math
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

RING = 1000         # action ring slots of config 2 (x B x 6 x 8 B = 197 MB at B = 4096)
SEED0 = 2020
FLOP_PER_ENV_STEP = 8.5e3  # ANM6Easy, fp64: SURVEY.md 8d gives 7-10 kFLOP (NR ~5.5 k at 3.2 iterations, projections, flows)
GROUPS = 5


def usable_cores():
    """Host cores this process may really use: min(affinity mask, cgroup CPU quota)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:  # cgroup v2: "max 100000" or "<quota> <period>"
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = min(n, max(1, int(float(q) / float(per))))
    except Exception:  # noqa: BLE001
        try:  # cgroup v1
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // per))
        except Exception:  # noqa: BLE001
            pass
    return max(1, n)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--envs", type=int, default=None, help="environment instances per GPU (default 4096; config 4: 8192)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 4])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="skip the closed-loop MPC leg (BASELINE configs[4])")
    ap.add_argument("--no-config4", action="store_true", help="skip the 30-bus leg (BASELINE configs[3])")
    ap.add_argument("--config5-envs", type=int, default=16384, help="global instances of the MPC leg")
    ap.add_argument("--cpu-steps", type=int, default=1500, help="CPU arms: timed env-steps per process")
    ap.add_argument("--min-timed-s", type=float, default=0.35, help="timed region of every rate (all groups together)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# CPU legs (oracle/ is used here only as the thing being timed for the baseline)
# ------------------------------------------------------------------------------------------
def cpu_rate(n_proc, n_steps, n_warm=10):
    """All host cores, one ANM6Easy env per process.  Returns (kind, total rate, per-process median, resets, what)."""
    import ref_rollout

    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    if ref_rollout.available():
        fn, kind = ref_rollout.random_agent_rollout, "reference"
        what = ("the reference's own gym_anm ANM6Easy (unmodified package staged under oracle/_ref; cvxpy = the exact-"
                "projection stand-in, so an UPPER bound on the real CVXPY->OSQP reference; NumPy/SciPy SuperLU Newton-Raphson)")
    else:
        import anm_numpy

        fn, kind = anm_numpy.random_agent_rollout, "port"
        what = "NumPy/SciPy port oracle/anm_numpy.py (reference-structured, bit-exact vs the reference goldens)"
    ctx = mp.get_context("spawn")
    with ctx.Pool(n_proc) as pool:
        res = pool.map(fn, [(SEED0 + i, n_steps, n_warm) for i in range(n_proc)])
    rates = [n / t for n, t, _ in res]
    return kind, sum(rates), statistics.median(rates), sum(r for _, _, r in res), what


def c_port_rate(B=4096, steps=100):
    os.environ["OMP_NUM_THREADS"] = str(usable_cores())  # before libgomp is loaded: no oversubscription under a CPU quota
    import numpy as np

    import anm_numpy
    import anm_oracle
    from gym_anm_b200.env_spec import anm6easy_spec

    spec = anm6easy_spec()
    env = anm_oracle.OracleEnv(spec, B)
    s0 = np.stack([anm_numpy.anm6easy_init_state(spec, np.random.Generator(np.random.PCG64(np.random.SeedSequence(SEED0 + i))))
                   for i in range(B)])  # fmt: skip
    env.reset(s0)
    rng = np.random.default_rng(1)
    acts = rng.uniform(spec.action_low, spec.action_high, size=(8, B, 6))
    for t in range(3):
        env.step(acts[t])
    t0 = time.perf_counter()
    for t in range(steps):
        _, _, term, _ = env.step(acts[t % 8])
        if term.any():
            env.reset(s0, mask=term.astype("uint8"))
    return B * steps / (time.perf_counter() - t0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = usable_cores()
    n = max(50, args.cpu_steps)
    t0 = time.perf_counter()
    kind, total, per_proc, resets, what = cpu_rate(cores, n, n_warm=min(10, max(1, args.warmup)))
    wall = time.perf_counter() - t0
    sample = "%d processes x %d timed env-steps of ANM6Easy-v0 (random agent, reset on termination)" % (cores, n)
    line = {
        "impl": "reference",
        "metric": "env_steps_per_sec",
        "value": total,
        "unit": "env-steps/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1000.0 * cores / total,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "ANM6Easy-v0 random-agent step() on the host cores, one env per core: " + what,
                   "envs_per_gpu": args.envs or 4096},
        "cpu_baseline": {"value": total, "unit": "env-steps/s", "cores": cores, "kind": kind, "sample": sample,
                         "per_core_median": per_proc, "wall_s": wall, "resets": resets,
                         "os_cpu_count": os.cpu_count()},  # fmt: skip
        "e2e": {"value": total, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, index):
        self.index, self.proc, self.samples = index, None, []
        self.windows = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)  # fmt: skip
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def mark(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        rows = []
        for ts, line in self.samples:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                rows.append((ts, float(parts[0]), float(parts[1]), parts[3:7]))
            except ValueError:
                continue
        inwin = [r for r in rows if any(a <= r[0] <= b for a, b in self.windows)] or rows
        if not inwin:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in inwin for n, v in zip(names, r[3]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(r[1] for r in inwin), "sm_max_mhz": max(r[2] for r in inwin),
                "reasons": reasons, "samples_in_timed_regions": len(inwin)}  # fmt: skip


# ------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------
def setup_config2(B, dev, rank):
    import torch

    from gym_anm_b200.anm6 import BatchedANM6Easy

    env = BatchedANM6Easy(B, device=dev, validate_actions=False, env_offset=rank * B)
    nb = env.native
    env.reset(seed=SEED0)
    pool = env.state.clone()  # reconstructed states of converged resets are valid, convergent s0 rows
    nb.set_autoreset_pool(pool)
    gen = torch.Generator(device=dev)
    gen.manual_seed(SEED0 + 10**6 + rank)
    lo = torch.as_tensor(env.spec.action_low, device=dev)
    hi = torch.as_tensor(env.spec.action_high, device=dev)
    ring = torch.rand((RING, B, 6), dtype=torch.float64, device=dev, generator=gen) * (hi - lo) + lo
    return {"nb": nb, "keep": env, "ring": ring, "ring_nv": None, "pool_rows": pool.shape[0],
            "workload": "ANM6Easy-v0, %d parallel envs per GPU, uniform-random actions, fp64 NR solve "
                        "(BASELINE.json configs[1])" % B,
            "flop_per_env_step": FLOP_PER_ENV_STEP}  # fmt: skip


def setup_config4(B, dev, rank):
    import numpy as np
    import torch

    from gym_anm_b200.env_spec import HostEnvSpec
    from gym_anm_b200.native import NativeBatch
    from gym_anm_b200.networks import synth_feeder_network

    spec = HostEnvSpec(synth_feeder_network(), "state", 1, 0.25, 0.99, 100, np.array([[0, 95]]), (1, 100))
    cn = spec.cn
    nb = NativeBatch(spec, B, dev)
    rng = np.random.default_rng(30 + rank)
    D, ns, ng = cn.N_device, cn.N_des, cn.N_non_slack_gen
    pos = {d: k for k, d in enumerate(cn.devices)}
    s0 = np.zeros((B, spec.state_N))
    for i in cn.load_ids:
        s0[:, pos[i]] = rng.uniform(cn.devices[i].p_min * 100, 0, B)
    for k, i in enumerate(cn.gen_ids):
        s0[:, pos[i]] = rng.uniform(0, cn.devices[i].p_max * 100, B)
        s0[:, 2 * D + ns + k] = rng.uniform(0, cn.devices[i].p_max * 100, B)
    for k, i in enumerate(cn.des_ids):
        s0[:, 2 * D + k] = rng.uniform(0, cn.devices[i].soc_max * 100, B)
    obs, state, conv = nb.reset(s0)
    assert bool(conv.all())
    pool = state.clone()
    nb.set_autoreset_pool(pool)
    R = 64  # 64 slots x 8192 x (18 + 17) x 8 B = 147 MB > L2
    g = torch.Generator(device=dev)
    g.manual_seed(4 + rank)
    lo, hi = torch.as_tensor(spec.action_low, device=dev), torch.as_tensor(spec.action_high, device=dev)
    ring = torch.rand((R, B, len(spec.action_low)), dtype=torch.float64, device=dev, generator=g) * (hi - lo) + lo
    nv_lo = torch.as_tensor([cn.devices[i].p_min * 100 for i in cn.load_ids] + [0.0] * ng + [0.0], device=dev)
    nv_hi = torch.as_tensor([0.0] * cn.N_load + [cn.devices[i].p_max * 100 for i in cn.gen_ids] + [95.0], device=dev)
    ring_nv = torch.rand((R, B, spec.n_next_vars), dtype=torch.float64, device=dev, generator=g) * (nv_hi - nv_lo) + nv_lo
    return {"nb": nb, "keep": spec, "ring": ring, "ring_nv": ring_nv, "pool_rows": pool.shape[0],
            "workload": "synthetic 30-bus feeder (synth_feeder_network: %d devices, %d branches), %d parallel envs per GPU, "
                        "random actions and caller-supplied next_vars, fp64 block-sparse NR solve (BASELINE.json configs[3])"
                        % (cn.N_device, len(cn.branches), B),
            "flop_per_env_step": None}  # fmt: skip


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class Config5Leg:
    """BASELINE configs[4] on this rank's shard: ANM6Easy-v0, `envs_global` instances over `world` GPUs, actions from the
    MPC-constant agent (examples/mpc_constant.py: planning_steps 10, safety_margin 0.96) whose DC-OPF programs are solved
    for the whole shard by the batched LP kernel (include/anm_lp.h); closed loop, nothing leaves the device except the
    per-step count of failed solutions.  No collective in here (the caller agrees on the step count and reduces the
    times).  A sample of the stepped shard is replayed through the C oracle after the timed region (same actions)."""

    def __init__(self, envs_global, world, rank, dev, n_check=32):
        from gym_anm_b200.agents import MPCAgentConstant
        from gym_anm_b200.anm6 import BatchedANM6Easy
        from gym_anm_b200.distributed import shard_slice

        sl = shard_slice(envs_global, rank, world)
        self.B, self.dev = sl.stop - sl.start, dev
        self.env = BatchedANM6Easy(self.B, device=dev, env_offset=sl.start, validate_actions=False)
        self.env.reset(seed=2020)
        self.agent = MPCAgentConstant(self.env.simulator, self.env.action_space, self.env.gamma, safety_margin=0.96,
                                      planning_steps=10, device=dev)
        self.n_chk = min(n_check, self.B)
        soc, aux, term = self.env.native.get_state()
        self.chk0 = [t[: self.n_chk].cpu().numpy().copy() for t in (soc, aux, term)]
        self.launches0 = self.env.native.launch_count
        self.rec = []

    def one_step(self, record):
        act = self.agent.act(self.env)  # CUDA tensor [B, A]
        obs, r, d, _, _ = self.env.step(act)
        if record:
            k = self.n_chk
            self.rec.append((act[:k].clone(), obs[:k].clone(), d[:k].clone()))
        return r

    def estimate(self):
        """Seconds per step after the cold solve (every instance's first basis) and two warm ones."""
        import torch

        for _ in range(3):
            self.one_step(True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            self.one_step(True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / 5

    def timed(self, n):
        import torch

        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        piv = torch.zeros((), dtype=torch.float64, device=self.dev)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(n):
            r = self.one_step(False)
            piv += self.agent._dev.lp.iters.double().mean()
        ev1.record()
        torch.cuda.synchronize()
        lp = self.agent._dev.lp
        self.out = {"B": self.B, "mean_pivots_per_solve": float(piv) / n, "lp_kernel": lp.kernel, "lp_bytes": lp.bytes,
                    "lp_stats": dict(self.agent.lp_stats), "mean_reward": float(r.mean()),
                    "terminated_frac": float(self.env.terminated.double().mean()),
                    "launches": self.env.native.launch_count - self.launches0 + lp.solves}
        return ev0.elapsed_time(ev1)

    def check(self):
        import numpy as np

        import anm_oracle

        cpu = anm_oracle.OracleEnv(self.env.spec, self.n_chk)
        cpu.soc[:], cpu.aux[:], cpu.terminated[:] = self.chk0
        worst, same = 0.0, True
        for a_t, o_t, d_t in self.rec:
            o_c, r_c, d_c, _ = cpu.step(a_t.cpu().numpy())
            same = same and np.array_equal(d_t.cpu().numpy().astype(bool), np.asarray(d_c).astype(bool))
            worst = max(worst, float(np.max(np.abs(o_t.cpu().numpy() - o_c) / np.maximum(np.abs(o_c), 1.0))))
        return {"instances": self.n_chk, "steps": len(self.rec), "max_rel_err_obs": worst, "terminated_equal": bool(same)}

    def close(self):
        import torch

        self.agent.close()
        self.agent = self.env = None
        torch.cuda.empty_cache()


def run_leg(make_leg, describe, min_timed_s, barrier, maxr, n_min=10, n_max=400):
    """An optional leg of the line (a dict, or {"error": ...}).  Every stage may fail on its own rank without costing the
    line its other numbers: the collectives (`barrier`, `maxr`) are outside the try blocks and a failed rank contributes
    +inf, so that all ranks take the same branches and nobody waits for a rank that has given up.
    `make_leg()` -> object with estimate() -> seconds per unit, timed(n) -> device ms (sets .out), check(), close();
    `describe(leg, n, ms, chk)` -> the dict."""
    leg, err, est, ms, chk, n = None, None, float("inf"), float("inf"), None, 0
    barrier()
    try:
        leg = make_leg()
        est = leg.estimate()
    except Exception as e:  # noqa: BLE001
        err = "%s: %s" % (type(e).__name__, e)
    est = maxr([est])[0]
    if math.isfinite(est):
        n = int(max(n_min, min(n_max, math.ceil(min_timed_s / max(est, 1e-5)))))
        barrier()
        try:
            ms = leg.timed(n)
            chk = leg.check()
        except Exception as e:  # noqa: BLE001
            ms, err = float("inf"), "%s: %s" % (type(e).__name__, e)
        ms = maxr([ms])[0]
    if math.isfinite(ms) and leg is not None:
        try:
            out = describe(leg, n, ms, chk)
        except Exception as e:  # noqa: BLE001
            out = {"error": "%s: %s" % (type(e).__name__, e)}
    else:
        out = {"error": err or "the leg failed on another rank"}
    try:
        if leg is not None:
            leg.close()
    except Exception:  # noqa: BLE001
        pass
    return out


def run_config5(envs_global, min_timed_s, world, rank, dev, barrier, maxr, leg_factory=None):
    """The `config5` key of the line (BASELINE configs[4])."""

    def describe(leg, n5, ms, chk):
        o5 = leg.out
        return {
            "value": envs_global * n5 / (ms / 1000.0), "unit": "env-steps/s", "n_gpus": world, "global_envs": envs_global,
            "envs_per_gpu": o5["B"], "steps": n5, "ms_per_step": ms / n5, "timed_s": ms / 1000.0,
            "lp_kernel": o5["lp_kernel"], "lp_bytes_per_gpu": o5["lp_bytes"],
            "mean_pivots_per_solve": o5["mean_pivots_per_solve"], "lp_stats": o5["lp_stats"],
            "mean_reward_last_step": o5["mean_reward"], "terminated_frac": o5["terminated_frac"],
            "gpu_launches": o5["launches"], "oracle_check": chk,
            "what": "BASELINE configs[4]: ANM6Easy-v0, %d instances over %d GPU(s), closed loop with the MPC-constant agent "
                    "(planning_steps 10, safety_margin 0.96): state tensor -> bounds -> batched dual simplex on the GPU "
                    "(anm_lp_solve, one %s per program, tableaux resident in HBM, warm-started) -> action tensor -> "
                    "anm_step; device-timed, max over ranks; counters and oracle check are rank 0's"
                    % (envs_global, world, o5["lp_kernel"]),
        }  # fmt: skip

    return run_leg(lambda: (leg_factory or Config5Leg)(envs_global, world, rank, dev), describe, min_timed_s, barrier, maxr)


class Config4Leg:
    """BASELINE configs[3] beside the headline: the synthetic 30-bus feeder, `B` instances on this GPU, uniform-random
    actions and caller-supplied next_vars, open-loop rollouts of 20 steps per launch, the blocks back to back (launches
    chained per instance) -- the `value` measurement of `bench.py --config 4`, shortened.  Parity of this workload is the
    test suite's business (test_synth30_env_vs_oracle, test_synth30_stressed_full_size_vs_oracle at 8192 instances)."""

    T = 20

    def __init__(self, B, world, rank, dev):
        import torch

        self.B, self.dev = B, dev
        try:
            self.wl, self.seed_rank = setup_config4(B, dev, rank), rank
        except AssertionError:  # this rank's random initial states hold one without a power-flow solution: rank 0's draw
            self.wl, self.seed_rank = setup_config4(B, dev, 0), 0
        nb = self.nb = self.wl["nb"]
        self.ring, self.ring_nv = self.wl["ring"], self.wl["ring_nv"]
        self.NR = self.ring.shape[0]
        self.outs = [(nb.empty(self.T, B, nb.O), nb.empty(self.T, B), nb.empty(self.T, B, dtype=torch.uint8))
                     for _ in range(3)]
        self.block = 0

    def run_block(self, chained):
        b, T = self.block, self.T
        self.block += 1
        o, r, d = self.outs[b % len(self.outs)]
        a0 = (b * T) % (self.NR - T + 1)
        self.nb.rollout(self.ring[a0:a0 + T], self.ring_nv[a0:a0 + T], out=(o, r, d), chained=chained)

    def estimate(self):
        import torch

        self.run_block(False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(3):
            self.run_block(i > 0)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / 3

    def timed(self, n):
        import torch

        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = self.nb.launch_count
        torch.cuda.synchronize()
        ev0.record()
        for i in range(n):
            self.run_block(i > 0)
        ev1.record()
        torch.cuda.synchronize()
        last = self.outs[(self.block - 1) % len(self.outs)][2]
        self.out = {"launches": self.nb.launch_count - launches0, "workload": self.wl["workload"],
                    "terminated_frac_last_step": float((last[self.T - 1] != 0).double().mean()),
                    "lanes_per_env": self.nb.sizes["lanes_per_env"], "seed_rank": self.seed_rank}
        return ev0.elapsed_time(ev1)

    def check(self):
        return None

    def close(self):
        import torch

        self.wl = self.nb = self.ring = self.ring_nv = self.outs = None
        torch.cuda.empty_cache()


def run_config4(B, min_timed_s, world, rank, dev, barrier, maxr, leg_factory=None):
    """The `config4` key of the line (BASELINE configs[3]); weak scaling: `B` instances per GPU."""

    def describe(leg, n, ms, chk):
        o, T = leg.out, Config4Leg.T
        return {"value": world * B * T * n / (ms / 1000.0), "unit": "env-steps/s", "n_gpus": world, "envs_per_gpu": B,
                "steps": T * n, "steps_per_launch": T, "ms_per_step": ms / (T * n), "timed_s": ms / 1000.0,
                "gpu_launches": o["launches"], "lanes_per_env": o["lanes_per_env"],
                "terminated_frac_last_step": o["terminated_frac_last_step"],
                "what": "BASELINE configs[3]: %s; %d-step anm_rollout launches back to back (chained per instance), "
                        "device-timed, max over ranks" % (o["workload"], T)}  # fmt: skip

    return run_leg(lambda: (leg_factory or Config4Leg)(B, world, rank, dev), describe, min_timed_s, barrier, maxr,
                   n_min=3, n_max=2000)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from gym_anm_b200 import _capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.envs or (4096 if args.config == 2 else 8192)
    K, W = args.steps, max(3, args.warmup)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(vals):
        """max over ranks of a list of floats (device-timed milliseconds)"""
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    wl = (setup_config2 if args.config == 2 else setup_config4)(B, dev, rank)
    nb, ring, ring_nv = wl["nb"], wl["ring"], wl["ring_nv"]
    NR, A, O, NV = ring.shape[0], nb.A, nb.O, nb.NV
    obs, rew, term = nb.empty(B, O), nb.empty(B), nb.empty(B, dtype=torch.uint8)
    nvs = (lambda t: None) if ring_nv is None else (lambda t: ring_nv[t % NR])
    in_bytes = 8 * A + (0 if ring_nv is None else 8 * NV)          # per env-step, read
    out_bytes = 8 * O + 8 + 1                                       # per env-step, written
    alg_bytes = in_bytes + out_bytes + 2 * 8 * (nb.n_des + nb.K) + 1  # + carried state in / out (234 B for ANM6Easy)

    # ---- warm-up (eager one-step launches) ------------------------------------------------------------
    for t in range(W):
        nb.step(ring[t % NR], nvs(t), out=(obs, rew, term))
    torch.cuda.synchronize()

    # ---- value: K-step blocks of anm_rollout launches, repeated back to back --------------------------------
    T = min(K, NR)                       # steps per launch
    n_full, rem = divmod(K, T)
    launches_per_block = n_full + (1 if rem else 0)
    n_out = max(1, min(12, int(math.ceil(160e6 / (T * B * out_bytes)))))  # output windows: > L2 in flight
    outs = [(nb.empty(T, B, O), nb.empty(T, B), nb.empty(T, B, dtype=torch.uint8)) for _ in range(n_out)]
    state = {"block": 0}

    def run_block(first_chained=True):
        """K steps: rollout launches over consecutive windows of the action ring, into the block's output window."""
        b = state["block"]
        state["block"] += 1
        o, r, d = outs[b % n_out]
        for i in range(launches_per_block):
            n = T if i < n_full else rem
            a0 = ((b * launches_per_block + i) * T) % (NR - n + 1)
            nb.rollout(ring[a0:a0 + n], None if ring_nv is None else ring_nv[a0:a0 + n],
                       out=(o[:n], r[:n], d[:n]), chained=(first_chained or i > 0))

    def timed_ms(fn, stream=None):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_host0 = time.perf_counter()
        ev0.record(stream)
        fn()
        ev1.record(stream)
        barrier()
        t_host1 = time.perf_counter()
        sampler.mark(t_host0, t_host1)
        return ev0.elapsed_time(ev1), 1000.0 * (t_host1 - t_host0)

    run_block(False)  # untimed: instruction cache, allocator
    torch.cuda.synchronize()
    iso = maxr([timed_ms(lambda: run_block(False))[0] for _ in range(9)])
    t_block = statistics.median(iso)
    R = int(max(1, min(20000, math.ceil(args.min_timed_s * 1000.0 / GROUPS / t_block))))

    def run_group():
        for i in range(R):
            run_block(first_chained=(i > 0))

    launches0 = nb.launch_count
    grp_ms = maxr([timed_ms(run_group)[0] for _ in range(GROUPS)])
    gpu_launches = nb.launch_count - launches0
    rates = [world * B * K * R / (ms / 1000.0) for ms in grp_ms]
    value = statistics.median(rates)
    ms_per_step = statistics.median(grp_ms) / (R * K)
    frac_reset = float((outs[(state["block"] - 1) % n_out][2][(rem or T) - 1] != 0).double().mean())

    # ---- one kernel launch per step, replayed from a CUDA graph ---------------------------------------------
    G = min(100, NR)
    side = torch.cuda.Stream()

    def capture(chained):
        g = torch.cuda.CUDAGraph()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for t in range(G):
                    nb.step(ring[t], nvs(t), out=(obs, rew, term), chained=chained)
        torch.cuda.current_stream().wait_stream(side)
        g.replay()  # untimed replay: graph upload + instruction cache
        return g

    per_step = {}
    for name, chained in (("chained", True), ("lockstep", False)):
        g = capture(chained)
        t1 = maxr([timed_ms(g.replay)[0]])[0]
        n_g = int(max(1, min(2000, math.ceil(args.min_timed_s * 1000.0 / 3 / t1))))
        ms3 = maxr([timed_ms(lambda: [g.replay() for _ in range(n_g)])[0] for _ in range(3)])
        ms_g = statistics.median(ms3)
        per_step[name] = {"value": world * B * n_g * G / (ms_g / 1000.0), "unit": "env-steps/s",
                          "ms_per_step": ms_g / (n_g * G), "steps": n_g * G, "timed_s": sum(ms3) / 1000.0}
        del g
    per_step["chained"]["what"] = ("one kernel launch per step (CUDA graph), launches ordered per instance "
                                   "(programmatic dependent launch + per-instance sequence numbers)")
    per_step["lockstep"]["what"] = ("one kernel launch per step (CUDA graph), every launch fully ordered after the "
                                    "previous one: the rate a closed-loop policy on the same stream can reach")

    # ---- e2e: host buffers through the C ABI's host calls (H2D + kernel + D2H for every step) -------------------
    hs = nb.host_stream
    Te = int(os.environ.get("BENCH_E2E_STEPS_PER_CALL", "100"))  # steps per queued rollout call; two sets of pinned
    Te = min(Te, NR)
    # output buffers; after queueing call i the host waits for call i-1 (whose results it would consume meanwhile)
    ring_host = ring.cpu().pin_memory()  # the agent's actions live in pinned host memory
    nv_host = None if ring_nv is None else ring_nv.cpu().pin_memory()
    obs_q = torch.empty(2, Te, B, O, dtype=torch.float64).pin_memory()
    rew_q = torch.empty(2, Te, B, dtype=torch.float64).pin_memory()
    term_q = torch.empty(2, Te, B, dtype=torch.uint8).pin_memory()
    act_ptr = [ring_host[t].data_ptr() for t in range(NR)]
    nv_ptr = [None] * NR if nv_host is None else [nv_host[t].data_ptr() for t in range(NR)]
    out_ptr = [(obs_q[q].data_ptr(), rew_q[q].data_ptr(), term_q[q].data_ptr()) for q in range(2)]

    def e2e_loop(n_calls, queued):
        """queued: n_calls x Te steps through anm_rollout_host_async; else n_calls synchronous one-step calls."""
        def body():
            if queued:
                for i in range(n_calls):
                    o, r, d = out_ptr[i % 2]
                    a0 = (i * Te) % (NR - Te + 1)
                    nb.rollout_host_async(Te, act_ptr[a0], nv_ptr[a0], o, r, d)
                    nb.host_sync_previous()
                nb.host_sync()
            else:
                for t in range(n_calls):
                    nb.step_host(act_ptr[t % NR], nv_ptr[t % NR], out_ptr[0][0], out_ptr[0][1], out_ptr[0][2])
        dev_ms, host_ms = timed_ms(body, hs)
        return max(dev_ms, host_ms)  # the host loop is part of the path: never report less than its wall time

    e2e_loop(2, True)
    t_call = maxr([e2e_loop(4, True)])[0] / 4
    n_calls = int(max(4, min(4000, math.ceil(args.min_timed_s * 1000.0 / t_call))))
    e2e_ms = maxr([e2e_loop(n_calls, True)])[0]
    e2e_rate = world * B * n_calls * Te / (e2e_ms / 1000.0)
    checksum = float(obs_q.sum()) + float(rew_q.sum())
    e2e_loop(3, False)
    t_s = maxr([e2e_loop(20, False)])[0] / 20
    n_s = int(max(20, min(20000, math.ceil(args.min_timed_s * 1000.0 / t_s))))
    sync_ms = maxr([e2e_loop(n_s, False)])[0]
    e2e_sync_rate = world * B * n_s / (sync_ms / 1000.0)

    # ---- the single collective of the path: all-gather of the batched [obs | reward | terminated] rows ---------
    gather = {}
    if world > 1:
        from gym_anm_b200.distributed import ObsExchange

        for mode in ("p2p", "nccl"):
            try:
                ex = ObsExchange(nb, mode=mode)
            except Exception as e:  # noqa: BLE001
                gather[mode] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
                continue
            for t in range(3):
                ex.step(ring[t % NR], nvs(t))
                ex.wait()
            t1 = maxr([timed_ms(lambda: [(ex.step(ring[t % NR], nvs(t)), ex.wait()) for t in range(50)])[0]])[0] / 50
            Kg = int(max(50, min(20000, math.ceil(args.min_timed_s * 1000.0 / t1))))
            gm = maxr([timed_ms(lambda: [(ex.step(ring[t % NR], nvs(t)), ex.wait()) for t in range(Kg)])[0]])[0]
            full = ex.wait()  # [B_global, O + 2] of the last step
            torch.cuda.synchronize()
            fb = full.cpu().numpy()
            gather[mode] = {"value": world * B * Kg / (gm / 1000.0), "unit": "env-steps/s", "steps": Kg,
                            "ms_per_step": gm / Kg, "launched": "eager (two C-ABI calls per step from Python)",
                            "checksum_all": hashlib.sha256(fb.tobytes()).hexdigest()[:16],
                            "checksum_first_shard": hashlib.sha256(fb[:B].tobytes()).hexdigest()[:16]}
            if mode == "p2p":
                # the same [step + arrival wait] pairs replayed from a CUDA graph, like `lockstep` (no Python between the
                # launches); the slot and the step count live on the device, so a graph of a multiple of `slots` steps
                # replays correctly
                gg = torch.cuda.CUDAGraph()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    with torch.cuda.graph(gg, stream=side):
                        for t in range(G):
                            ex.step(ring[t], nvs(t))
                            ex.wait()
                torch.cuda.current_stream().wait_stream(side)
                gg.replay()
                t1 = maxr([timed_ms(gg.replay)[0]])[0]
                n_g = int(max(1, min(2000, math.ceil(args.min_timed_s * 1000.0 / t1))))
                gm2 = maxr([timed_ms(lambda: [gg.replay() for _ in range(n_g)])[0]])[0]
                gather[mode].update({"eager_value": gather[mode]["value"], "value": world * B * n_g * G / (gm2 / 1000.0),
                                     "ms_per_step": gm2 / (n_g * G), "steps": n_g * G,
                                     "launched": "CUDA graph of %d [step, arrival wait] pairs" % G})
                del gg
            ex.close()
            barrier()

    # ---- the fp64 FMA rate of this GPU (denominator of the roofline's fp64 leg), before the optional leg below ----
    import ctypes as C

    fp64_peak = None
    if rank == 0:
        try:
            tf = C.c_double(0.0)
            _capi.check(nb.lib.anm_debug_fp64_peak(local, C.byref(tf)), nb.lib)
            fp64_peak = tf.value
        except Exception:  # noqa: BLE001
            pass

    # ---- BASELINE configs[3]: the 30-bus feeder, 8192 instances per GPU, open-loop rollouts -----------------------
    config4 = None
    if args.config == 2 and not args.no_config4:
        try:
            config4 = run_config4(8192, args.min_timed_s, world, rank, dev, barrier, maxr)
        except Exception as e:  # noqa: BLE001
            config4 = {"error": "%s: %s" % (type(e).__name__, e)}
    # ---- BASELINE configs[4]: 16 384 instances (global) driven by the MPC-constant agent, LPs on the GPU ----------
    config5 = None
    if args.config == 2 and not args.no_config5:
        try:
            config5 = run_config5(args.config5_envs, args.min_timed_s, world, rank, dev, barrier, maxr)
        except Exception as e:  # noqa: BLE001  (e.g. a barrier that fails: the line keeps its other numbers)
            config5 = {"error": "%s: %s" % (type(e).__name__, e)}

    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- rank 0: roofline, CPU baseline, the JSON line ----------------------------------------------------
    peaks, peak_src = {}, "fallback (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    grp_s = statistics.median(grp_ms) / 1000.0
    launches_per_group = R * launches_per_block
    launch_s = grp_s / launches_per_group                    # mean launch duration over the timed region (CUDA events)
    env_steps_per_launch = B * K / launches_per_block
    achieved = alg_bytes * env_steps_per_launch / launch_s / 1e9
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
    except Exception:  # noqa: BLE001
        pass
    per_es = ncu.get("dram_bytes_per_env_step")
    fp64_leg = None
    if wl["flop_per_env_step"] and fp64_peak:
        a = wl["flop_per_env_step"] * env_steps_per_launch / launch_s / 1e12
        fp64_leg = {"achieved": a, "peak": fp64_peak, "unit": "TFLOP/s", "frac": a / fp64_peak,
                    "algorithmic_flop_per_env_step": wl["flop_per_env_step"],
                    "peak_source": "measured in this run: anm_debug_fp64_peak (register-only DFMA chains)"}
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": None if (per_es is None or args.config != 2) else per_es * env_steps_per_launch,
        "peak_source": peak_src, "algorithmic_bytes_per_env_step": alg_bytes, "env_steps_per_launch": env_steps_per_launch,
        "kernel_us_mean": launch_s * 1e6, "kernel_us_per_step": launch_s * 1e6 * launches_per_block / K,
        "binding_resource": "dependent fp64 issue latency at ~1.7 warps per scheduler (B = 4096 is one wave), not HBM: "
                            "the hbm fraction is reported because the metric asks for it; see `fp64`",
        "fp64": fp64_leg,
        "ncu_static": {"static": True, "what": "constants copied from profiles/ncu_summary.json (an earlier ncu capture of this "
                       "kernel), NOT measurements of this run", "fp64_pipe_pct": ncu.get("fp64_pipe_pct"),
                       "issue_slot_pct": ncu.get("issue_active_pct"), "stall_samples_pct": ncu.get("stall_samples_pct"),
                       "dram_bytes_per_env_step": per_es, "profile": ncu.get("source")},
    }  # fmt: skip
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = usable_cores()
        n = 1000
        kind, total, per_core, resets, what = cpu_rate(cores, n)
        cpu_baseline = {
            "value": total, "unit": "env-steps/s", "cores": cores, "kind": kind,
            "sample": "%d processes x %d timed env-steps of ANM6Easy-v0, random agent: %s" % (cores, n, what),
            "per_core_median": per_core,
            "c_port": {"value": c_port_rate(), "unit": "env-steps/s", "threads": cores,
                       "what": "oracle/anm_oracle.c (scalar C, dense LU, OpenMP over 4096 envs)"},
        }  # fmt: skip
    line = {
        "metric": "env_steps_per_sec",
        "value": value,
        "unit": "env-steps/s",
        "n_gpus": world,
        "steps": K,
        "warmup": W,
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": wl["workload"],
            "envs_per_gpu": B, "global_envs": world * B, "parallelism": "dp%d (independent env shards)" % world,
            "autoreset": "next-step, from a pool of %d convergent initial states" % wl["pool_rows"],
            "l2": "inputs cycle through a %d-slot action ring of %.0f MB (> 126 MB L2); every block writes its own window "
                  "of a %d-window output ring (%.0f MB)" % (NR, (ring.numel() + (0 if ring_nv is None else ring_nv.numel())) * 8 / 1e6,
                                                          n_out, n_out * T * B * out_bytes / 1e6),
            "launch": "K = %d steps = %d anm_rollout launch(es) of <= %d steps; the block repeated back to back %d times per "
                      "group (launches chained per instance), %d groups, median group" % (K, launches_per_block, T, R, GROUPS),
            "repeats_per_group": R, "groups": GROUPS, "group_ms": grp_ms, "timed_s": sum(grp_ms) / 1000.0,
            "lanes_per_env": nb.sizes["lanes_per_env"], "smem_bytes_per_cta": nb.sizes["smem_bytes"],
            "frac_envs_reset_last_step": frac_reset,
        },
        "clocks": clocks,
        "isolated_block": {"value": world * B * K / (t_block / 1000.0), "unit": "env-steps/s", "ms": t_block,
                           "ms_min": min(iso), "ms_max": max(iso), "repeats": len(iso),
                           "what": "ONE K-step block with the device idle before and after (median of 9)"},
        "chained": per_step["chained"],
        "lockstep": per_step["lockstep"],
        "e2e": {"value": e2e_rate, "unit": "env-steps/s", "h2d_bytes_per_step": B * in_bytes,
                "d2h_bytes_per_step": B * out_bytes, "steps": n_calls * Te, "steps_per_call": Te, "timed_s": e2e_ms / 1000.0,
                "api": "anm_rollout_host_async (%d steps per call) + anm_host_sync_previous after every call (C ABI; pinned host "
                       "arrays, uploaded / downloaded by the copy engines through device staging buffers while the "
                       "kernels of consecutive calls run back to back)" % Te,
                "checksum": checksum},  # fmt: skip
        "sync_every_step": {"value": e2e_sync_rate, "unit": "env-steps/s", "steps": n_s, "timed_s": sync_ms / 1000.0,
                            "api": "anm_step_host (synchronous; pinned host arrays in / out every step)"},
        "gpu_launches": gpu_launches,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    if config5 is not None:
        line["config5"] = config5
    if config4 is not None:
        line["config4"] = config4
    if gather:
        if "value" in gather.get("p2p", {}):
            line["with_obs_allgather"] = dict(gather["p2p"], what="one launch per step; the kernel epilogue stores every "
                                              "[obs | reward | terminated] row into every peer's gather buffer over NVLink "
                                              "(fused step + all-gather), then the per-step arrival wait")
        else:
            line["with_obs_allgather"] = gather.get("p2p")
        line["with_obs_allgather_nccl"] = dict(gather.get("nccl", {}), what="one launch per step writing packed rows locally + "
                                               "ncclAllGather (all_gather_into_tensor) per step")
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    # exactly ONE line on stdout: libraries that print to fd 1 (e.g. NCCL's version banner) go to stderr
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_real_stdout, "w")
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
