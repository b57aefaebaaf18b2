#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched ANM6Easy step() path on N B200s (one JSON line).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs B] [--impl ours|reference]

Workload (BASELINE.json configs[1]): ANM6Easy-v0, B = 4096 environment instances PER GPU
(weak scaling), uniform-random actions over the action box, fp64 Newton-Raphson.  A "step"
is one pass of the hot path over the whole batch: every instance advances by one timestep.
Terminated instances are re-initialised by the kernel's next-step auto-reset from a pool of
initial states that are known to converge, so every instance does one full transition per step.

* `value`      : K steps as open-loop rollouts: the random agent does not look at the observations, so
                 its action sequence (a 1000-slot x B x 6 fp64 ring resident in HBM, 197 MB > the 126 MB
                 L2) is handed to `anm_rollout` 1000 steps at a time.  One kernel launch takes every
                 instance through its 1000 steps -- carried state on chip, every step's obs / reward /
                 terminated row written to HBM ([T, B, .] outputs) -- so an instance whose Newton iteration
                 diverges (100 iterations) only delays the three instances that share its warp.  Timed
                 with CUDA events, max over ranks.
                 `per_step_launches` reports the same workload as ONE kernel launch per step (CUDA graph):
                 `chained` (launches ordered per instance, ANM_STEP_CHAINED) and `lockstep` (every launch
                 fully ordered after the previous one -- what a closed-loop policy sees).
* `e2e`        : the same metric through the C-ABI host calls `anm_rollout_host_async` + `anm_host_sync`
                 -- pinned HOST action arrays in, HOST obs / reward / terminated arrays out, for every
                 step, H2D + D2H inside the timed region (copy engines, overlapped with the kernels);
                 100 steps per call, the host waits for call i-1 after queueing call i.
                 `e2e.sync_every_step` is the synchronous one-step `anm_step_host` (zero-copy).
* `roofline`   : algorithmic bytes (234 B / env-step) x B / mean kernel time
                 against the measured HBM copy bandwidth (MEASURED_PEAKS.json).  The path is NOT
                 HBM-bound (arithmetic intensity ~30 fp64 FLOP/B); the fraction is reported
                 because the metric asks for it, next to the ncu-measured fp64-pipe utilisation.
* `cpu_baseline`: the reference-structured NumPy/SciPy port (oracle/anm_numpy.py, bit-exact vs
                 the reference goldens) on all host cores, plus the scalar C port for context.

`--impl reference` times that NumPy/SciPy port (the reference itself is pure Python and does
not exist on the GPU box) on all host cores and prints the same line with "impl": "reference".
"""
import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

B_MIN_BYTES = 234  # algorithmic bytes / env-step: 8*(A + O + 1 + 2*(n_des + K)) + 2  
RING = 1000        # action ring slots (x B x 6 x 8 B = 197 MB at B = 4096)
SEED0 = 2020


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
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--envs", type=int, default=4096, help="environment instances per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=1500, help="reference arm: timed env-steps per process")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# CPU legs (oracle/ is used here only as the thing being timed for the baseline)
# ------------------------------------------------------------------------------------------
def numpy_port_rate(n_proc, n_steps, n_warm=10):
    import anm_numpy

    ctx = mp.get_context("spawn")
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    with ctx.Pool(n_proc) as pool:
        res = pool.map(anm_numpy.random_agent_rollout, [(SEED0 + i, n_steps, n_warm) for i in range(n_proc)])
    rates = [n / t for n, t, _ in res]
    return sum(rates), statistics.median(rates), sum(r for _, _, r in res)


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
    n = max(50, min(args.cpu_steps, args.steps))
    t0 = time.perf_counter()
    total, per_proc, resets = numpy_port_rate(cores, n, n_warm=min(10, max(1, args.warmup)))
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
        "config": {
            "workload": "ANM6Easy-v0 random-agent step(), reference-structured NumPy/SciPy port "
            "(oracle/anm_numpy.py; bit-exact vs reference goldens; exact projection instead of CVXPY->OSQP, "
            "so an UPPER bound on the real reference's speed), one env per host core",
            "envs_per_gpu": args.envs,
        },
        "cpu_baseline": {"value": total, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample,
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
# GPU arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from gym_anm_b200.anm6 import BatchedANM6Easy

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, K, W = args.envs, args.steps, max(3, args.warmup)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- setup: envs (global-index seeding), auto-reset pool, action rings ----------------------
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
    ring_host = ring.cpu().pin_memory()  # e2e leg: the agent's actions live in pinned host memory
    obs, rew, term = nb.empty(B, 18), nb.empty(B), nb.empty(B, dtype=torch.uint8)

    # ---- warm-up (eager launches) ---------------------------------------------------------------------
    for t in range(W):
        nb.step(ring[t % RING], None, out=(obs, rew, term))
    torch.cuda.synchronize()

    # ---- value: K steps as open-loop rollouts (anm_rollout: T steps per kernel launch) ---------------------------
    # The random agent is open-loop, so the whole action sequence is handed over at once: one launch takes every
    # instance through T = RING consecutive steps (carried state on chip) and writes every step's obs / reward /
    # terminated row; consecutive launches are chained per instance (ANM_STEP_CHAINED).
    T = min(K, RING)
    n_roll, rem = divmod(K, T)
    obs_r, rew_r, term_r = nb.empty(T, B, 18), nb.empty(T, B), nb.empty(T, B, dtype=torch.uint8)
    nb.rollout(ring[:T], out=(obs_r, rew_r, term_r))  # untimed: instruction cache, allocator
    torch.cuda.synchronize()

    def timed(fn):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_host0 = time.perf_counter()
        ev0.record()
        fn()
        ev1.record()
        barrier()
        t_host1 = time.perf_counter()
        sampler.mark(t_host0, t_host1)
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def run_rollouts():
        for i in range(n_roll):
            nb.rollout(ring[:T], out=(obs_r, rew_r, term_r), chained=(i > 0))
        if rem:
            nb.rollout(ring[:rem], out=(obs_r[:rem], rew_r[:rem], term_r[:rem]), chained=True)

    launches0 = nb.launch_count
    ms_total = timed(run_rollouts)
    gpu_launches = nb.launch_count - launches0
    frac_reset = float((term_r[-1] != 0).double().mean())

    # ---- the same K' steps as one kernel launch per step, replayed from a CUDA graph ----------------------------
    G = min(K, RING)
    side = torch.cuda.Stream()

    def capture(chained):
        g = torch.cuda.CUDAGraph()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for t in range(G):
                    nb.step(ring[t], None, out=(obs, rew, term), chained=chained)
        torch.cuda.current_stream().wait_stream(side)
        g.replay()  # untimed replay: graph upload + instruction cache
        return g

    n_g = max(1, min(K // G, 4))
    per_step = {}
    for name, chained in (("chained", True), ("lockstep", False)):
        g = capture(chained)
        ms_g = timed(lambda: [g.replay() for _ in range(n_g)])
        per_step[name] = {"value": world * B * n_g * G / (ms_g / 1000.0), "unit": "env-steps/s",
                          "ms_per_step": ms_g / (n_g * G), "steps": n_g * G}
        del g
    per_step["chained"]["what"] = ("one kernel launch per step (CUDA graph), launches ordered per instance "
                                   "(programmatic dependent launch + per-instance sequence numbers)")
    per_step["lockstep"]["what"] = ("one kernel launch per step (CUDA graph), every launch fully ordered after the "
                                    "previous one: the rate a closed-loop policy on the same stream can reach")

    # ---- e2e: host buffers through the C ABI's host calls (H2D + kernel + D2H for every step) -------------------
    hs = nb.host_stream
    Te = int(os.environ.get("BENCH_E2E_STEPS_PER_CALL", "100"))  # steps per queued rollout call; two sets of pinned
    # output buffers; after queueing call i the host waits for call i-1 (whose results it would consume meanwhile)
    obs_q = torch.empty(2, Te, B, 18, dtype=torch.float64).pin_memory()
    rew_q = torch.empty(2, Te, B, dtype=torch.float64).pin_memory()
    term_q = torch.empty(2, Te, B, dtype=torch.uint8).pin_memory()
    act_ptr = [ring_host[t].data_ptr() for t in range(RING)]
    out_ptr = [(obs_q[q].data_ptr(), rew_q[q].data_ptr(), term_q[q].data_ptr()) for q in range(2)]

    def e2e_loop(n, queued):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_host0 = time.perf_counter()
        e0.record(hs)
        if queued:
            for i in range(n // Te):
                o, r, d = out_ptr[i % 2]
                nb.rollout_host_async(Te, act_ptr[(i * Te) % (RING - Te + 1)], None, o, r, d)
                nb.host_sync_previous()
            nb.host_sync()
        else:
            for t in range(n):
                nb.step_host(act_ptr[t % RING], None, out_ptr[0][0], out_ptr[0][1], out_ptr[0][2])
        e1.record(hs)
        barrier()
        t_host1 = time.perf_counter()
        sampler.mark(t_host0, t_host1)
        ms = torch.tensor([max(e0.elapsed_time(e1), 1000.0 * (t_host1 - t_host0))], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return world * B * n / (float(ms) / 1000.0)

    Ke = max(2 * Te, min(K, 8000) // (2 * Te) * (2 * Te))
    e2e_loop(2 * Te, True)
    e2e_rate = e2e_loop(Ke, True)
    checksum = float(obs_q.sum()) + float(rew_q.sum())
    e2e_loop(3, False)
    Ks = min(K, 2000)
    e2e_sync_rate = e2e_loop(Ks, False)

    # ---- optional: the single collective of the path (all-gather of the batched observation) ------------
    gather_rate = None
    if world > 1:
        from gym_anm_b200.distributed import all_gather_rows

        Kg = min(K, 2000)
        packed = torch.empty(B, 20, dtype=torch.float64, device=dev)
        for t in range(3):
            all_gather_rows(packed)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        g0.record()
        for t in range(Kg):
            nb.step(ring[t % RING], None, out=(obs, rew, term))
            packed[:, :18] = obs
            packed[:, 18] = rew
            packed[:, 19] = term
            all_gather_rows(packed)
        g1.record()
        barrier()
        gm = torch.tensor([g0.elapsed_time(g1)], device=dev, dtype=torch.float64)
        dist.all_reduce(gm, op=dist.ReduceOp.MAX)
        gather_rate = world * B * Kg / (float(gm) / 1000.0)

    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- rank 0: roofline, CPU baseline, the JSON line ----------------------------------------------------
    value = world * B * K / (ms_total / 1000.0)
    kernel_s = (ms_total / 1000.0) / K
    peaks, peak_src = {}, "fallback (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    launch_s = (ms_total / 1000.0) / max(gpu_launches, 1)      # one anm_rollout launch = T steps of every instance
    steps_per_launch = K / max(gpu_launches, 1)
    achieved = B_MIN_BYTES * B * steps_per_launch / launch_s / 1e9
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
    except Exception:  # noqa: BLE001
        pass
    per_es = ncu.get("dram_bytes_per_env_step")
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": None if per_es is None else per_es * B * steps_per_launch, "peak_source": peak_src,
        "algorithmic_bytes_per_env_step": B_MIN_BYTES, "env_steps_per_launch": B * steps_per_launch,
        "kernel_us_mean": launch_s * 1e6, "kernel_us_per_step": kernel_s * 1e6,
        "traffic_note": "dram bytes per env-step of the ncu capture (a 50-step launch: the 153 B / env-step of outputs "
                        "were still in the 126 MB L2 when it ended) x env-steps per launch",
        "note": "compute/latency-bound path (dependent fp64 FMA chains, shuffle / shared-memory latency at 1.7 warps per "
                "scheduler), not HBM-bound: 234 B against ~1650 warp instructions per env-step; see the fp64 / issue fields",
        "fp64_pipe_pct_ncu": ncu.get("fp64_pipe_pct"), "issue_slot_pct_ncu": ncu.get("issue_active_pct"),
        "stall_samples_pct_ncu": ncu.get("stall_samples_pct"), "ncu_profile": ncu.get("source"),
    }  # fmt: skip
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = usable_cores()
        n = 800
        total, per_core, resets = numpy_port_rate(cores, n)
        cpu_baseline = {
            "value": total, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": "%d processes x %d timed env-steps of ANM6Easy-v0, random agent, NumPy/SciPy port "
                      "(oracle/anm_numpy.py, reference-structured, exact projection)" % (cores, n),
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
        "ms_per_step": ms_total / K,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": "ANM6Easy-v0, %d parallel envs per GPU, uniform-random actions, fp64 NR solve "
                        "(BASELINE.json configs[1])" % B,
            "envs_per_gpu": B, "global_envs": world * B, "parallelism": "dp%d (independent env shards)" % world,
            "autoreset": "next-step, from a pool of %d convergent initial states" % pool.shape[0],
            "l2": "inputs cycle through a %d-slot action ring of %.0f MB (> 126 MB L2)" % (RING, ring.numel() * 8 / 1e6),
            "launch": "anm_rollout: %d launches x %d steps (+%d); each instance runs its steps back to back with its "
                      "carried state on chip and writes every step's obs / reward / terminated row to HBM; launches "
                      "chained per instance" % (n_roll, T, rem),
            "lanes_per_env": nb.sizes["lanes_per_env"], "smem_bytes_per_cta": nb.sizes["smem_bytes"],
            "frac_envs_reset_last_step": frac_reset,
        },
        "clocks": clocks,
        "per_step_launches": per_step,
        "e2e": {"value": e2e_rate, "unit": "env-steps/s", "h2d_bytes_per_step": B * 6 * 8,
                "d2h_bytes_per_step": B * (18 * 8 + 8 + 1), "steps": Ke, "steps_per_call": Te,
                "api": "anm_rollout_host_async (%d steps per call) + anm_host_sync_previous after every call (C ABI; pinned host "
                       "arrays, uploaded / downloaded by the copy engines through device staging buffers while the "
                       "kernels of consecutive calls run back to back)" % Te,
                "sync_every_step": {"value": e2e_sync_rate, "unit": "env-steps/s", "steps": Ks,
                                    "api": "anm_step_host (synchronous)"},
                "checksum": checksum},  # fmt: skip
        "gpu_launches": gpu_launches,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    if gather_rate is not None:
        line["with_obs_allgather"] = {"value": gather_rate, "unit": "env-steps/s",
                                      "what": "step + NCCL all-gather of [B,20] (obs|reward|terminated) per step"}  # fmt: skip
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
