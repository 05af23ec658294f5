"""Multi-GPU data parallelism over VIEWS -- new work, the reference is single-GPU
(its batch loop renders views sequentially: gs/gaussian_splatting.py:1439-1460; the per-view parameter
gradients are summed by autograd before one optimizer.step(), trainer.py:587-599).

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Every rank holds all Gaussian
parameters; rank r renders the views {v : v mod G == r}; gradients accumulate locally into ONE flat fp32
buffer (the parameter .grad tensors are views of it, so there is no pack pass) and a single
`all_reduce(SUM)` per step makes the replicas identical -- 11+3*C*C floats per Gaussian (236 B at SH
degree 3).  SUM (not AVG) reproduces the single-GPU gradient of the same view batch, which is what the
renderer-level contract is; a trainer that normalises its loss by the local batch must divide by the
global batch instead (guidance/stable_diffusion.py:304).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

FIELDS_RGB = (("mean", 3), ("qvec", 4), ("svec", 3), ("alpha", 1), ("color", 3))


def field_layout(N: int, C: Optional[int]) -> List[tuple]:
    """[(name, shape, offset, numel)] of the flat buffers.  C=None -> RGB colour, else SH [N,3,C*C].
    Every field starts on a 16-byte boundary (offset % 4 == 0): the kernels read qvec / write g_qvec as float4 and
    stream the buffers 16 bytes at a time, so a field must not start mid-vector when N % 4 != 0.  The (at most 3)
    padding floats between fields belong to no Gaussian: zero gradient, zero moments, never read."""
    fields = [("mean", (N, 3)), ("qvec", (N, 4)), ("svec", (N, 3)), ("alpha", (N,))]
    fields.append(("color", (N, 3)) if C is None else ("sh", (N, 3, C * C)))
    out, off = [], 0
    for name, shape in fields:
        n = 1
        for s in shape:
            n *= s
        out.append((name, shape, off, n))
        off = (off + n + 3) // 4 * 4
    return out


def layout_total(layout) -> int:
    """floats of a flat buffer with this layout (fields + alignment padding, a multiple of 4)"""
    _, _, off, n = layout[-1]
    return (off + n + 3) // 4 * 4


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Round-robin view assignment: 1 view/GPU when n_views == world (BASELINE config 4 at 8 GPUs)."""
    return [v for v in range(n_views) if v % world == rank]


class ViewParallelRenderer:
    """Replicated Gaussians + view-sharded rendering + one flat gradient all-reduce per step.

    render_fn(params: dict of leaf tensors, view_index) -> loss-like scalar or (tensor, grad) pair;
    the default (GPU) implementation is `gsgen_b200.rasterizer.render_view`.  The hook exists so that the
    host-side logic (layout, sharding, collective) is testable on CPU with gloo.
    """

    def __init__(self, params: Dict[str, torch.Tensor], C: Optional[int], device, group=None,
                 register_nccl: bool = False, sparse_allreduce: bool = False, sparse_max_fraction: float = 0.6):
        self.C = C
        self.device = torch.device(device)
        self.group = group
        N = params["mean"].shape[0]
        self.N = N
        self.layout = field_layout(N, C)
        total = layout_total(self.layout)
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.grad_buffer_kind = "plain"
        self.flat_grad = None
        if register_nccl and self.world > 1 and self.device.type == "cuda":
            self.flat_grad = self._alloc_registered(total)
        if self.flat_grad is None:
            self.flat_grad = torch.zeros(total, dtype=torch.float32, device=self.device)
        # sparse all-reduce (CUDA, world > 1): the backward marks the Gaussians it sent gradients to; the ranks
        # MAX-reduce that byte mask and all-reduce only the union's rows (gsb200_rows_pack / _unpack).  T < T_thresh
        # hides most of a scene from any one view: 10 % of C3's Gaussians per view, 36 % over 8 views (oracle count).
        self.sparse = bool(sparse_allreduce) and self.world > 1 and self.device.type == "cuda"
        self.sparse_max_fraction = float(sparse_max_fraction)
        self.last_allreduce = {"mode": "dense", "rows": N}
        if self.sparse:
            self.touched = torch.zeros(N, dtype=torch.uint8, device=self.device)
            self._excl = torch.empty(N, dtype=torch.int32, device=self.device)  # ids of the union's rows
            row = sum(n // N for _, _, _, n in self.layout) if N else 0
            self._row_floats = row
            self._packed = torch.empty(int(self.sparse_max_fraction * N) * row + 64, dtype=torch.float32,
                                       device=self.device)
        self.params: Dict[str, torch.Tensor] = {}
        self.grad_views: Dict[str, torch.Tensor] = {}
        for name, shape, off, n in self.layout:
            self.flat_param[off:off + n].copy_(params[name].reshape(-1).to(self.device, torch.float32))
            p = self.flat_param[off:off + n].view(shape)
            p.requires_grad_(True)
            self.params[name] = p
            self.grad_views[name] = self.flat_grad[off:off + n].view(shape)
        if self.sparse:  # render_view(grad_sink=...) hands this to the backward (gsb200_view_grads.touched)
            self.grad_views["touched"] = self.touched

    def _alloc_registered(self, total: int):
        """The all-reduce operand allocated with ncclMemAlloc and registered with the communicator (NCCL user-buffer
        registration): on an NVSwitch box NCCL can then run the 118-236 MB gradient all-reduce through NVLS
        (in-switch reduction, zero-copy) instead of staging it through its own buffers.  Returns None when this torch /
        NCCL build does not offer the allocator -- the caller falls back to a plain tensor (`grad_buffer_kind`)."""
        try:
            pg = self.group if self.group is not None else dist.group.WORLD
            backend = pg._get_backend(self.device)
            pool = torch.cuda.MemPool(backend.mem_allocator)
            with torch.cuda.use_mem_pool(pool):
                buf = torch.zeros(total, dtype=torch.float32, device=self.device)
            backend.register_mem_pool(pool)
            self._nccl_pool = pool  # keeps the registration alive
            self.grad_buffer_kind = "ncclMemAlloc + register_mem_pool"
            return buf
        except Exception as ex:  # pragma: no cover (needs NCCL)
            self.grad_buffer_kind = f"plain (registration unavailable: {type(ex).__name__}: {str(ex)[:80]})"
            return None

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    @property
    def rank(self) -> int:
        return dist.get_rank(self.group) if dist.is_available() and dist.is_initialized() else 0

    def grad_bytes(self) -> int:
        return self.flat_grad.numel() * 4

    def zero_grad(self):
        self.flat_grad.zero_()
        if self.sparse:
            self.touched.zero_()
        for name, p in self.params.items():
            p.grad = self.grad_views[name]  # autograd accumulates in place into the flat buffer

    def all_reduce(self):
        """ONE all-reduce(SUM) of the gradients per step.  Dense: the whole flat buffer.  Sparse: MAX-reduce the
        1-byte-per-Gaussian touched mask, pack the union's rows (one host sync for the row count), all-reduce
        rows x (11 + 3C^2) floats, unpack; falls back to the dense collective when the union exceeds
        `sparse_max_fraction` of the Gaussians (the same decision on every rank: the mask is identical)."""
        if self.world == 1:
            return
        self._n_allreduce = getattr(self, "_n_allreduce", 0) + 1
        if not self.sparse or self._n_allreduce < getattr(self, "_dense_until", 0):
            # (after a step whose union was too large the next 32 steps go dense without paying for the mask reduction
            # and the row count: the touched fraction of a scene changes slowly; same decision on every rank)
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
            return
        import ctypes

        from . import _lib

        dist.all_reduce(self.touched, op=dist.ReduceOp.MAX, group=self.group)
        names = [n for n, _, _, _ in self.layout]
        offs = (ctypes.c_uint64 * len(names))(*[off for _, _, off, _ in self.layout])
        widths = (ctypes.c_uint32 * len(names))(*[(n // self.N) for _, _, _, n in self.layout])
        n_keep = ctypes.c_uint32(0)
        L, c, st = _lib.lib(), _lib.ctx(self.device), _lib.stream_ptr(self.device)
        cap = self._packed.numel()
        _lib.check(L.gsb200_rows_pack(c, _lib.fptr(self.flat_grad), _lib.fptr(self._packed), ctypes.c_uint64(cap), offs,
                                      widths, ctypes.c_int32(len(names)), ctypes.c_uint32(self.N),
                                      _lib.ptr(self.touched, torch.uint8), _lib.iptr(self._excl), ctypes.byref(n_keep),
                                      st))
        k = int(n_keep.value)
        if k * self._row_floats > cap:  # most Gaussians touched (e.g. C4's 8 views: 69 %): dense is cheaper
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
            self.last_allreduce = {"mode": "dense (union above sparse_max_fraction)", "rows": k}
            self._dense_until = self._n_allreduce + 32
            return
        if k:
            dist.all_reduce(self._packed[: k * self._row_floats], op=dist.ReduceOp.SUM, group=self.group)
            _lib.check(L.gsb200_rows_unpack(c, _lib.fptr(self.flat_grad), _lib.fptr(self._packed), offs, widths,
                                            ctypes.c_int32(len(names)), ctypes.c_uint32(self.N),
                                            _lib.iptr(self._excl), ctypes.c_uint32(k), st))
        self.last_allreduce = {"mode": "sparse", "rows": k}

    def step(self, n_views: int, render_and_backward: Callable[[Dict[str, torch.Tensor], int], None]):
        """zero grads -> local views fwd+bwd (gradients accumulate) -> one all-reduce.  Returns local view ids."""
        self.zero_grad()
        mine = shard_views(n_views, self.rank, self.world)
        for v in mine:
            render_and_backward(self.params, v)
        self.all_reduce()
        return mine
