"""Multi-GPU: environment instances are independent, so the batch shards trivially.

One process per GPU (torchrun); rank r owns the contiguous slice
``shard_slice(num_envs_global, r, world)`` and seeds its envs by GLOBAL index, so the
results do not depend on the world size (shard-equivalence).  There is no
collective on the data path.  The single optional exchange is the all-gather of the
batched observation (and reward / terminated) for a central learner.
"""
import torch
import torch.distributed as dist


def shard_slice(num_envs_global, rank, world_size):
    """Contiguous, balanced partition: the first (num_envs_global % world) ranks get one extra env."""
    base, extra = divmod(int(num_envs_global), int(world_size))
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def all_gather_rows(t, group=None):
    """All-gather row-sharded tensors [B_local, ...] -> [B_global, ...] (ncclAllGather over
    NVLink/NVSwitch with the nccl backend; gloo on CPU in the tests).  Shards may be uneven."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    world = dist.get_world_size(group)
    n_local = torch.tensor([t.shape[0]], device=t.device, dtype=torch.int64)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c) for c in counts]
    if len(set(counts)) == 1:
        out = torch.empty((world * counts[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        return out
    m = max(counts)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


class ObsGather:
    """Double-buffered all-gather of (obs, reward, terminated) on a side stream so the
    collective of step t overlaps the kernel of step t+1 (CUDA only)."""

    def __init__(self, group=None):
        self.group = group
        self.stream = torch.cuda.Stream() if torch.cuda.is_available() else None

    def __call__(self, obs, reward, terminated):
        if self.stream is None:
            return all_gather_rows(obs, self.group), all_gather_rows(reward, self.group), all_gather_rows(terminated, self.group)
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            packed = torch.cat([obs, reward.unsqueeze(1), terminated.to(obs.dtype).unsqueeze(1)], dim=1)
            g = all_gather_rows(packed, self.group)
        self.last = g
        return g

    def wait(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        g = self.last
        return g[:, :-2], g[:, -2], g[:, -1] != 0
