"""Timed random-agent rollout of the UNMODIFIED reference ANM6Easy (bench.py's reference arm / cpu_baseline).

MEASUREMENT INFRASTRUCTURE ONLY.  Imports the reference through oracle/ref_loader.py (staged copy under oracle/_ref on
the GPU box); cvxpy is the exact-projection stand-in of oracle/shims/ (an upper bound on the real CVXPY -> OSQP speed),
NumPy / SciPy (SuperLU) do the reference's own Newton-Raphson arithmetic.
"""
import os
import sys
import time
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)


def available():
    import ref_loader

    return ref_loader.reference_available()


def random_agent_rollout(args):
    """(seed, n_steps, n_warm) -> (n_steps, seconds, n_resets): ANM6Easy.reset(seed), then uniform-random actions over
    the action box, reset() after every termination -- examples/random_agent.py, the loop of BASELINE configs[0]."""
    import numpy as np

    import ref_loader

    seed, n_steps, n_warm = args
    warnings.simplefilter("ignore")
    ref_loader.load_reference()
    from gym_anm.envs import ANM6Easy

    env = ANM6Easy()
    env.reset(seed=seed)
    rng = np.random.default_rng(seed + 10**6)
    lo, hi = env.action_space.low, env.action_space.high
    n_reset, t0 = 0, None
    for t in range(n_warm + n_steps):
        if t == n_warm:
            t0 = time.perf_counter()
        _, _, term, _, _ = env.step(rng.uniform(lo, hi))
        if term:
            env.reset()
            n_reset += 1
    return n_steps, time.perf_counter() - t0, n_reset


if __name__ == "__main__":
    print(random_agent_rollout((2020, 300, 5)))
