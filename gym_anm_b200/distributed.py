"""Multi-GPU: environment instances are independent, so the batch shards trivially.

One process per GPU (torchrun); rank r owns the contiguous slice ``shard_slice(num_envs_global, r, world)`` and seeds
its envs by GLOBAL index, so the results do not depend on the world size (shard-equivalence).
There is no collective on the data path.  The single optional exchange is the all-gather of the batched
[obs | reward | terminated] rows for a central learner:

* `ObsExchange(native, mode="p2p")` -- the exchange fused into the step: the kernel's epilogue stores every row straight
  into every rank's gather buffer over NVLink / NVSwitch (peer memory mapped with CUDA IPC: `anm_gather_*`,
  `anm_step_packed` in include/anm_b200.h); `wait()` enqueues the per-step arrival wait and returns the gathered
  [B_global, O + 2] rows.  No separate collective launch, no packing kernels, no host synchronisation.
* `ObsExchange(native, mode="nccl")` -- the kernel writes the packed rows locally and `ncclAllGather` moves them on a
  side stream (double-buffered); the fallback where CUDA IPC is not available.
* `all_gather_rows(t, counts=None)` -- plain collective for anything else (gloo on CPU in the tests).
"""
import ctypes as C

import torch
import torch.distributed as dist


def shard_slice(num_envs_global, rank, world_size):
    """Contiguous, balanced partition: the first (num_envs_global % world) ranks get one extra env."""
    base, extra = divmod(int(num_envs_global), int(world_size))
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def shard_counts(num_envs_global, world_size):
    return [shard_slice(num_envs_global, r, world_size).stop - shard_slice(num_envs_global, r, world_size).start
            for r in range(world_size)]  # fmt: skip


def all_gather_rows(t, group=None, counts=None):
    """All-gather row-sharded tensors [B_local, ...] -> [B_global, ...] (ncclAllGather over NVLink/NVSwitch with the
    nccl backend; gloo on CPU in the tests).  `counts` = rows of every rank (`shard_counts`) when the shards are
    uneven; None = every rank holds the same number of rows.  No size exchange, no host synchronisation."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    world = dist.get_world_size(group)
    if counts is None or len(set(counts)) == 1:
        out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        return out
    counts = [int(c) for c in counts]
    if len(counts) != world or counts[dist.get_rank(group)] != t.shape[0]:
        raise ValueError("counts %r do not describe this rank's %d rows" % (counts, t.shape[0]))
    m = max(counts)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


class _DevArray:
    """A device buffer owned by the native library, exposed through __cuda_array_interface__."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2,
                                         "strides": None}  # fmt: skip


class ObsExchange:
    """Step + all-gather of the batched [obs | reward | terminated] rows (see the module docstring).

        ex = ObsExchange(env.native)                  # every rank, once (collective: exchanges the IPC handles)
        ex.step(actions)                              # = native.step(...), rows also go to every rank
        rows = ex.wait()                              # [B_global, O + 2] float64; terminated = rows[:, -1] != 0

    `wait()` orders the current stream after the arrival of every rank's rows of the most recent `step`.  With
    `slots` buffers a rank may run `slots - 1` steps ahead of the slowest consumer: consume `rows` (on the current
    stream) before launching `step` number t + slots - 1.  Equal shards only for mode="nccl"; `rows_global` / `row0`
    default to equal contiguous shards in rank order."""

    def __init__(self, native, mode="p2p", slots=2, group=None, row0=None, rows_global=None, inline_wait=True):
        from . import _capi

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("ObsExchange needs an initialised torch.distributed process group")
        self.nb, self.mode, self.group, self.slots = native, mode, group, int(slots)
        # p2p: the step kernel itself waits for every rank's rows (one kernel = step + all-gather + arrival); False: the
        # arrival wait is a separate small kernel enqueued by wait(), so that work enqueued between step() and wait()
        # overlaps the slower ranks
        self.inline_wait = bool(inline_wait)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.B, self.W = native.B, native.O + 2
        self.row0 = self.rank * self.B if row0 is None else int(row0)
        self.rows = self.world * self.B if rows_global is None else int(rows_global)
        self._capi, self.lib, self.dev = _capi, native.lib, native.device
        self.obs, self.reward = native.empty(self.B, native.O), native.empty(self.B)
        self.term = native.empty(self.B, dtype=torch.uint8)
        self._n = 0
        if mode == "p2p":
            mine = (C.c_ubyte * 64)()
            _capi.check(self.lib.anm_gather_create(native.h, self.world, self.rank, self.row0, self.rows, self.slots, mine),
                        self.lib)  # fmt: skip
            h_local = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(self.dev)
            h_all = torch.empty(self.world * 64, dtype=torch.uint8, device=self.dev)
            dist.all_gather_into_tensor(h_all, h_local, group=group)
            blob = bytes(h_all.cpu().numpy().tobytes())
            _capi.check(self.lib.anm_gather_attach(native.h, blob), self.lib)
            dist.barrier(group=group)  # every rank has mapped every buffer before anyone stores into one
            self._open = True
        elif mode == "nccl":
            self.packed = [native.empty(self.B, self.W) for _ in range(self.slots)]
            self.full = [native.empty(self.world * self.B, self.W) for _ in range(self.slots)]
            self.side = torch.cuda.Stream(device=self.dev)
            self.done = [torch.cuda.Event() for _ in range(self.slots)]
            self._open = True
        else:
            raise ValueError("mode must be 'p2p' or 'nccl'")

    def step(self, action, next_vars=None, extras=None):
        nb = self.nb
        action = nb._f64(action, nb.A)
        nv = None if next_vars is None else nb._f64(next_vars, nb.NV)
        ex = None
        if extras:
            ex = self._capi.StepExtras()
            for k in ("state", "e_loss", "penalty", "n_iter", "full_state"):
                t = extras.get(k)
                setattr(ex, k, None if t is None else t.data_ptr())
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        slot = self._n % self.slots
        packed = None if self.mode == "p2p" else self.packed[slot]
        self._capi.check(self.lib.anm_step_packed(nb.h, p(action), p(nv), p(self.obs), p(self.reward), p(self.term),
                                                  p(packed), (2 if self.inline_wait else 1) if self.mode == "p2p" else 0,
                                                  None if ex is None else C.byref(ex), nb._stream()), self.lib)  # fmt: skip
        if self.mode == "nccl":
            # the pack is done by the kernel on the current stream; only the collective runs on the side stream, which
            # waits for this step's kernel (and for the consumer of this slot's previous contents: stream order of wait())
            self.side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(self.side):
                dist.all_gather_into_tensor(self.full[slot], packed, group=self.group)
                self.done[slot].record(self.side)
        self._n += 1
        return self.obs, self.reward, self.term

    def wait(self):
        """Rows [B_global, O + 2] of the most recent step, valid for work enqueued on the current stream after this."""
        if self._n < 1:
            raise RuntimeError("no step yet")
        slot = (self._n - 1) % self.slots
        if self.mode == "nccl":
            torch.cuda.current_stream(self.dev).wait_event(self.done[slot])
            return self.full[slot]
        out = C.c_void_p()
        self._capi.check(self.lib.anm_gather_wait(self.nb.h, C.byref(out), self.nb._stream()), self.lib)
        return torch.as_tensor(_DevArray(out.value, (self.rows, self.W)), device=self.dev)

    def close(self):
        if getattr(self, "_open", False):
            self._open = False
            torch.cuda.synchronize(self.dev)
            if self.mode == "p2p":
                dist.barrier(group=self.group)  # nobody is still storing into a buffer that is about to be unmapped
                self._capi.check(self.lib.anm_gather_destroy(self.nb.h), self.lib)
                dist.barrier(group=self.group)

    def __del__(self):
        try:
            if getattr(self, "_open", False) and self.mode == "p2p":
                self.lib.anm_gather_destroy(self.nb.h)
        except Exception:  # noqa: BLE001
            pass
