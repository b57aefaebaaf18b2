"""Worker of tests/test_multi_gpu.py (one process per GPU, started by torch.distributed.run).

Shard-equivalence on real GPUs: B_global ANM6Easy instances seeded by GLOBAL index, actions drawn per
global instance, sharded over the ranks; every step's [obs | reward | terminated] rows are all-gathered by the fused
step + all-gather (mode p2p) and by the NCCL variant, and must equal, bit for bit, the rows of the same B_global
instances run as ONE batch on rank 0's GPU.  Also: device-side seeded resets by global index, and the packed rows against
the separate obs / reward / terminated outputs.
"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gym_anm_b200.anm6 import BatchedANM6Easy  # noqa: E402
from gym_anm_b200.distributed import ObsExchange, shard_slice  # noqa: E402


def actions(spec, lo, hi, t):
    return np.stack([np.random.default_rng(10**6 + g * 1000 + t).uniform(spec.action_low, spec.action_high)
                     for g in range(lo, hi)])  # fmt: skip


def main():
    B_global, T, seed = int(sys.argv[1]), int(sys.argv[2]), 2020
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sl = shard_slice(B_global, rank, world)
    assert B_global % world == 0
    B = sl.stop - sl.start
    out = {}
    for device_init in (False, True):
        env = BatchedANM6Easy(B, device=dev, validate_actions=False, env_offset=sl.start, device_init=device_init)
        obs0, _ = env.reset(seed=seed)
        want_env = None
        if rank == 0:  # the same B_global instances as one batch on this GPU
            want_env = BatchedANM6Easy(B_global, device=dev, validate_actions=False, device_init=device_init)
            want0, _ = want_env.reset(seed=seed)
            assert torch.equal(want0[sl.start:sl.stop], obs0), "reset observations differ from the single-batch run"
        snap = env.state_dict()
        snap_want = None if want_env is None else want_env.state_dict()
        for mode in ("p2p", "p2p_separate_wait", "nccl"):  # every mode replays the same steps from the same carried state
            env.load_state_dict(snap)
            if want_env is not None:
                want_env.load_state_dict(snap_want)
            ex = ObsExchange(env.native, mode=mode.split("_")[0], inline_wait=(mode == "p2p"))
            h = hashlib.sha256()
            for t in range(T):
                a = actions(env.spec, sl.start, sl.stop, t)
                o, r, d = ex.step(a)
                rows = ex.wait()
                torch.cuda.synchronize()
                mine = rows[sl.start:sl.stop]
                assert torch.equal(mine[:, :-2], o) and torch.equal(mine[:, -2], r) and torch.equal(mine[:, -1] != 0, d != 0)
                if rank == 0:
                    wo, wr, wd, _, _ = want_env.step(actions(env.spec, 0, B_global, t))
                    assert torch.equal(rows[:, :-2], wo), (mode, t, "gathered observations differ from the single batch")
                    assert torch.equal(rows[:, -2], wr) and torch.equal(rows[:, -1] != 0, wd)
                    h.update(rows.cpu().numpy().tobytes())
            ex.close()
            out[(device_init, mode)] = h.hexdigest()[:16]
        del env, want_env
    if rank == 0:
        assert len(set(out.values())) == 1 and len(out) == 6, out
        print("MGPU_OK world=%d B_global=%d T=%d hash=%s" % (world, B_global, T, out[(False, "p2p")]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
