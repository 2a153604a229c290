"""Data-parallel plumbing: one process per GPU, clips sharded contiguously.  The forward path has no data-path
collective (clips are independent units - SURVEY.md section 8e); torch.distributed is used for the barrier and the
max-over-ranks timing reduction of the benchmark contract, and - training step, BASELINE config 5 - for the gradient
all-reduce, which `GradBuckets` issues per layer group from inside the backward pass (reference train.py:366-368
wraps the model in DistributedDataParallel, which does the same)."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when not launched by torchrun."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init(backend: str = None) -> Tuple[int, int, int]:
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n_items clips over `world` ranks; the first n_items % world ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device=None) -> float:
    """The slowest rank's time: the benchmark contract times a multi-GPU step as the max over ranks."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


class GradBuckets:
    """Gradient all-reduce overlapped with the backward pass.

    All gradients live in ONE flat fp32 buffer laid out in the order the backward pass finishes them (`groups`: lists
    of parameters, first group = finished first, i.e. the LAST layers of the forward).  Every parameter's ``.grad`` is a
    view into the buffer, autograd accumulates into it in place, and a post-accumulate hook counts a group's parameters
    down; when the last one arrives the group's contiguous slice is all-reduced asynchronously (NCCL runs it on its own
    stream, ordered behind the kernels that produced the slice) while autograd goes on with the earlier layers.
    ``finish()`` waits for all groups before the optimizer.  Works eagerly and inside a CUDA-graph capture (the
    collectives and their stream dependencies are captured with the rest of the step).

    ``comm_dtype=torch.bfloat16`` halves the bytes on NVLink: a group's slice is cast to bf16, averaged, and cast back
    (the optimizer still sees fp32 gradients).
    """

    def __init__(self, groups, comm_dtype=torch.float32, process_group=None, average=True, align: int = 8):
        """align: every parameter's slice starts at a multiple of `align` elements (16 bytes for 16-bit images of the
        buffer: gradient / weight views are TMA operands of the native GEMMs)."""
        self.groups = [[p for p in g if p.requires_grad] for g in groups]
        self.groups = [g for g in self.groups if g]
        self.pg = process_group
        self.comm_dtype = comm_dtype
        self.average = average
        params = [p for g in self.groups for p in g]
        if not params:
            raise ValueError("GradBuckets: no parameters")
        dev = params[0].device
        pad = lambda n: (n + align - 1) // align * align  # noqa: E731
        self.flat = torch.zeros(sum(pad(p.numel()) for p in params), device=dev, dtype=torch.float32)
        self.slices = []
        self.offsets = {}  # id(param) -> (start, numel) in the flat buffer
        self._group_of = {}
        off = 0
        for gi, g in enumerate(self.groups):
            start = off
            for p in g:
                p.grad = self.flat[off:off + p.numel()].view_as(p)
                self._group_of[id(p)] = gi
                self.offsets[id(p)] = (off, p.numel())
                off += pad(p.numel())
            self.slices.append((start, off))
        self.comm = None if comm_dtype == torch.float32 else torch.empty_like(self.flat, dtype=comm_dtype)
        self._pending = [len(g) for g in self.groups]
        self._works = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.pg) > 1

    @property
    def world(self):
        return dist.get_world_size(self.pg) if self.enabled else 1

    def zero(self):
        """Start of a step: clear the gradients and re-arm the group counters."""
        self.flat.zero_()
        self._pending = [len(g) for g in self.groups]
        self._works = []

    def _launch(self, gi):
        a, b = self.slices[gi]
        if not self.enabled:
            return
        if self.comm is None:
            buf = self.flat[a:b]
        else:
            buf = self.comm[a:b]
            buf.copy_(self.flat[a:b])
        op = dist.ReduceOp.AVG if (self.average and dist.get_backend(self.pg) == "nccl") else dist.ReduceOp.SUM
        self._works.append((gi, dist.all_reduce(buf, op=op, group=self.pg, async_op=True), op))

    def _on_grad(self, p):
        gi = self._group_of[id(p)]
        self._pending[gi] -= 1
        if self._pending[gi] == 0:
            self._launch(gi)

    def notify(self, p):
        """A parameter whose gradient was written into its view by a native kernel (no autograd accumulation, hence no
        hook): count it as finished."""
        self._on_grad(p)

    def finish(self):
        """After backward(): groups whose hooks never fired completely (unused parameters) are reduced now, then every
        collective is waited for and - reduced-precision transport - the averaged values are cast back."""
        for gi, n in enumerate(self._pending):
            if n > 0:
                self._pending[gi] = 0
                self._launch(gi)
        for gi, work, op in self._works:
            work.wait()
            a, b = self.slices[gi]
            if self.comm is not None:
                self.flat[a:b].copy_(self.comm[a:b])
            if self.average and op == dist.ReduceOp.SUM:
                self.flat[a:b].div_(self.world)
        self._works = []

    def bytes_per_step(self) -> int:
        return self.flat.numel() * (4 if self.comm is None else self.comm.element_size()) if self.enabled else 0

    def remove_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def shutdown():
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
