"""Persistent capacity arena for the Gaussians and their Adam state, with densify / prune as row operations on
it.

Reference behaviour restated (gs/gaussian_splatting.py): every densify / prune step re-creates each parameter as a
new `nn.Parameter` (`torch.cat` / boolean-mask indexing), deletes and re-inserts the Adam state of its param group
with the same surgery (`densify_on_optimizer` :481-522, `prune_optimizer` :421-449), and re-slices the densification
statistics (`prune_by_mask` :528-549).  That is ~10 allocations and copies of every tensor per operation plus a
rebuild of the optimizer's state dict.

Here the parameters, their gradients and both Adam moments live in four flat fp32 buffers with the SAME field-major
layout (`field_layout(capacity, C)`): field f occupies `capacity * width_f` floats and its first `N` rows are live.
  * append (clone / split children): rows are written behind row N of every field -- nothing else moves; the
    Adam moments of new rows are zero (the reference concatenates `zeros_like`), the shared step counter is kept
    (the reference skips 0-dim state entries).
  * prune: a stable compaction of the live rows of every field, in place, by one gather per field and buffer.
  * growth: only when `N + k > capacity` (capacity grows geometrically), one copy of the live rows.
The live gradient rows are what the multi-GPU step all-reduces; `gsgen_b200.optim.FlatAdam` steps the whole arena in
one kernel (dead rows carry zero gradients and zero moments, so they do not move).

Row movers: on a CUDA device the two hand-written kernels of csrc/store.cu (`gsb200_store_compact`: prefix sum + ONE
launch that moves all fields of all four buffers into the arena's shadow buffers, which are then swapped in;
`gsb200_store_append`: one launch).  On CPU tensors (the test-suite exercises this host logic without a GPU) the same
row operations are torch indexing ops; `tests/test_store_gpu.py` checks the kernels against that path bit for bit.
Rules built on those row operations: clone / split / "official" / "scale" / "all" (:551-632, :770-810), the legacy rule the
top-level experiment configs select (`densify_legacy`, :820-946, incl. its optimizer reset), compactness-based
densification over the K-nearest-neighbour kernel (:634-743; csrc/knn.cu) and the three prune rules (:1124-1176).
Selection rules follow the reference bit for bit, including two quirks that a drop-in must keep:
  * `densify_by_clone` compares `torch.norm(grads, dim=-1)` of the 1-D per-Gaussian statistic, i.e. ONE number for the
    whole scene, with the threshold (:616-618);
  * `prune_by_mask` re-slices the statistics with the post-densify mask and falls back to zeros when their length no
    longer matches (:533-549).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from .parallel import field_layout, layout_total


def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    """kornia 0.6.0 `quaternion_to_rotation_matrix(q, WXYZ)` (utils/transforms.py:53-54): normalise (eps 1e-12), then
    the unit-quaternion matrix.  Same restatement as the kernels' `quat_to_rotmat` (csrc/gsb200_math.cuh)."""
    q = torch.nn.functional.normalize(q, p=2.0, dim=-1, eps=1e-12)
    w, x, y, z = q.unbind(-1)
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    one = torch.ones_like(w)
    return torch.stack([one - (tyy + tzz), txy - twz, txz + twy,
                        txy + twz, one - (txx + tzz), tyz - twx,
                        txz - twy, tyz + twx, one - (txx + tyy)], dim=-1).view(*q.shape[:-1], 3, 3)


class GaussianStore:
    """Raw leaves (`mean`, `qvec`, `svec` = log scale, `alpha` = logit opacity, `color` = logit RGB | `sh`) in a
    capacity arena.  `params[name]` are views of the live rows (autograd leaves whose `.grad` is a view of the flat
    gradient buffer), re-created after every operation that changes N -- as the reference re-creates its Parameters.
    """

    def __init__(self, params: Dict[str, torch.Tensor], C: Optional[int], device=None, capacity: Optional[int] = None,
                 growth: float = 1.5, group=None, knn_fn=None):
        self.C, self.group, self.growth = C, group, float(growth)
        # neighbour search of the compactness rules: gsgen_b200.knn.knn_points (CUDA kernel) unless a callable with
        # the same signature is plugged in (CPU tests: the oracle's brute-force search -- the product has no CPU path)
        self.knn_fn = knn_fn
        self.device = torch.device(device) if device is not None else params["mean"].device
        self.N = int(params["mean"].shape[0])
        self.cap = self._round_cap(max(int(capacity or 0), self.N, 1))
        self.optimizer = None
        self._alloc(self.cap)
        for name, shape, off, n in field_layout(self.N, C):
            self._rows(self.flat_param, name, self.N).copy_(params[name].detach().to(self.device, torch.float32)
                                                           .reshape(self.N, -1))
        self.reset_densify_info()
        self._make_leaves()

    # ---- arena ---------------------------------------------------------------------------------------
    @staticmethod
    def _round_cap(cap: int) -> int:
        """capacities are multiples of 4 rows: every field of every buffer then starts on a 16-byte boundary (the
        kernels access qvec / g_qvec as float4; a capacity like 6145 = int(4096*1.5)+1 used to misalign them)"""
        return (int(cap) + 3) // 4 * 4

    def _alloc(self, cap: int):
        cap = self._round_cap(cap)
        self.cap = cap
        self._shadow = None  # second set of buffers the CUDA compaction writes into (allocated at the first prune)
        self.layout = field_layout(cap, self.C)
        total = layout_total(self.layout)
        mk = lambda: torch.zeros(total, dtype=torch.float32, device=self.device)
        self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq = mk(), mk(), mk(), mk()
        self._field = {name: (shape, off) for name, shape, off, _ in self.layout}

    def _buffers(self):
        return (self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq)

    def _rows(self, buf: torch.Tensor, name: str, n: int, start: int = 0) -> torch.Tensor:
        """rows [start, start+n) of field `name` as a 2-D [n, width] view of `buf`"""
        shape, off = self._field[name]
        w = 1
        for s in shape[1:]:
            w *= s
        return buf[off + start * w: off + (start + n) * w].view(n, w)

    def _make_leaves(self):
        self.params: Dict[str, torch.Tensor] = {}
        self.grad_views: Dict[str, torch.Tensor] = {}
        for name, (shape, _) in self._field.items():
            live = (self.N,) + tuple(shape[1:])
            p = self._rows(self.flat_param, name, self.N).view(live)
            p.requires_grad_(True)
            self.params[name] = p
            self.grad_views[name] = self._rows(self.flat_grad, name, self.N).view(live)
            p.grad = self.grad_views[name]  # autograd (when no grad_sink is used) accumulates straight into the arena
        if self.optimizer is not None:
            self.optimizer.rebind(self.flat_param, self.flat_grad, self.layout, self.exp_avg, self.exp_avg_sq)

    def _grow(self, need: int):
        old = (self._buffers(), dict(self._field), self.N)
        new_cap = max(need, int(self.cap * self.growth) + 1)
        old_bufs, old_field, n = old
        old_rows = {name: [self._rows(b, name, n) for b in old_bufs] for name in old_field}
        self._alloc(new_cap)
        for name, rows in old_rows.items():
            for dst, src in zip(self._buffers(), rows):
                self._rows(dst, name, n).copy_(src)

    # ---- activated values (`self.svec` / `self.alpha` of the reference, :113-119) ---------------------------
    @property
    def svec_act(self) -> torch.Tensor:
        return torch.exp(self.params["svec"].detach())

    @property
    def alpha_act(self) -> torch.Tensor:
        return torch.sigmoid(self.params["alpha"].detach())

    def make_optimizer(self, lr, max_steps: int = 15000, betas=(0.9, 0.999), eps: float = 1e-15):
        """`set_optimizer` (:398-419): Adam over the whole arena, moments owned by the store so that they follow
        their rows through densify / prune."""
        from .optim import FlatAdam

        self.optimizer = FlatAdam(self.flat_param, self.flat_grad, self.layout, lr, max_steps, betas, eps,
                                  state=(self.exp_avg, self.exp_avg_sq))
        return self.optimizer

    def zero_grad(self):
        self.flat_grad.zero_()
        for name, p in self.params.items():
            p.grad = self.grad_views[name]

    def grad_bytes(self) -> int:
        return sum(v.numel() for v in self.grad_views.values()) * 4

    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def all_reduce(self):
        """SUM over ranks of the live gradient rows (one flat collective when the arena is full, else one per
        field -- dead capacity is not sent)."""
        if self.world_size() == 1:
            return
        if self.N == self.cap:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        else:
            for name in self._field:
                dist.all_reduce(self._rows(self.flat_grad, name, self.N), op=dist.ReduceOp.SUM, group=self.group)

    def sync_densify_info(self):
        """Replica consistency of the view-parallel mode.  The densification statistics are
        per-view side effects (`mean_2d_grad_accum += |g_mean2d|`, `cnt += 1` under each view's mask,
        gs/gaussian_splatting.py:464-469; `max_radii2d = max(...)`, :1240-1245); with the views of a batch sharded over
        ranks every rank only holds its own views' share.  Called before anything reads them (`densify_step` /
        `prune_step`): SUM of the LOCAL contributions since the last synchronisation for the two accumulators (one
        [2,N] collective), MAX for the radii -- after which all ranks hold exactly the numbers a single process
        rendering all views would, and select the same Gaussians."""
        if self.world_size() == 1:
            return
        if self._acc_base is None or self._acc_base.shape[1] != self.N:
            raise RuntimeError("densification statistics changed length outside densify / prune")
        local = torch.stack((self.mean_2d_grad_accum, self.cnt)) - self._acc_base
        dist.all_reduce(local, op=dist.ReduceOp.SUM, group=self.group)
        synced = self._acc_base + local
        self.mean_2d_grad_accum, self.cnt = synced[0].contiguous(), synced[1].contiguous()
        self._acc_base = synced.clone()
        dist.all_reduce(self.max_radii2d, op=dist.ReduceOp.MAX, group=self.group)

    def _mark_synced(self):
        """the statistics are identical on every rank right now (after a reset / a row operation on synced values)"""
        self._acc_base = torch.stack((self.mean_2d_grad_accum, self.cnt)).clone() if self.world_size() > 1 else None

    # ---- checkpoint view (gs/gaussian_splatting.py:294-339) -------------------------------------------------
    def get_params_for_save(self) -> Dict[str, torch.Tensor]:
        """The raw leaves under the reference's checkpoint keys (`Trainer.save` stores them as ckpt["params"],
        trainer.py:255-266; optimizer state is not part of a reference checkpoint).  Detached CPU copies, so a saved
        checkpoint does not alias the arena; `gsgen_b200.export` and `GaussianStore.load` take this dict."""
        return {name: p.detach().to("cpu").clone() for name, p in self.params.items()}

    @classmethod
    def load(cls, ckpt, C: Optional[int] = None, device=None, **kw) -> "GaussianStore":
        """From a reference checkpoint: a path / dict with "params" (trainer checkpoint) or the params dict itself
        (renderer-only checkpoint, :318-331).  Keys other than the Gaussian fields (cfg, bg, ...) are ignored."""
        if not isinstance(ckpt, dict):
            ckpt = torch.load(ckpt, map_location="cpu")
        if "params" in ckpt:
            ckpt = ckpt["params"]
        names = [n for n, _, _, _ in field_layout(1, C)]
        missing = [n for n in names if n not in ckpt]
        if missing:
            raise RuntimeError(f"checkpoint lacks {missing}")
        return cls({n: ckpt[n] for n in names}, C, device, **kw)

    # ---- densification statistics (gs/gaussian_splatting.py:464-479) ---------------------------------------
    def reset_densify_info(self):
        z = lambda: torch.zeros(self.N, dtype=torch.float32, device=self.device)
        self.mean_2d_grad_accum, self.cnt, self.max_radii2d = z(), z(), z()
        self._mark_synced()

    def update_densify_info(self, mask: torch.Tensor, mean2d_grad: torch.Tensor, radii2d: Optional[torch.Tensor] = None):
        """One rendered view: `mask` [N] bool, `mean2d_grad` [N,2] (aux["mean2d_grad"] of render_view, zero rows for
        culled Gaussians), optional `radii2d` [N] (aux["radii2d"]).  :464-469 and :1240-1245."""
        self.mean_2d_grad_accum[mask] += mean2d_grad[mask].norm(dim=-1)
        self.cnt[mask] += 1
        if radii2d is not None:
            self.max_radii2d[mask] = torch.max(self.max_radii2d[mask], radii2d[mask])

    # ---- row operations ----------------------------------------------------------------------------------
    def append(self, new_params: Dict[str, torch.Tensor]) -> int:
        """`densify_with_new_params` (:481-526): new rows behind the live ones, zero Adam moments."""
        k = int(new_params["mean"].shape[0])
        if k == 0:
            return 0
        if self.N + k > self.cap:
            self._grow(self.N + k)
        if self.device.type == "cuda":
            self._append_cuda(new_params, k)
        else:
            for name in self._field:
                self._rows(self.flat_param, name, k, self.N).copy_(new_params[name].detach().reshape(k, -1))
                for buf in (self.flat_grad, self.exp_avg, self.exp_avg_sq):
                    self._rows(buf, name, k, self.N).zero_()
        self.N += k
        self._make_leaves()
        return k

    def prune_by_mask(self, mask: torch.Tensor) -> int:
        """`prune_by_mask` (:528-549): rows with mask=True are removed, order of the others kept; Adam moments follow
        their rows (`prune_optimizer` :421-449)."""
        if mask.shape[0] != self.N:
            raise RuntimeError(f"prune mask has {mask.shape[0]} rows, the store {self.N}")
        mask = mask.to(self.device)
        if self.device.type == "cuda":
            n_keep = self._compact_cuda(mask)
            keep = None
        else:
            keep = (~mask).nonzero(as_tuple=True)[0]
            n_keep = int(keep.shape[0])
            if n_keep < self.N:
                for buf in self._buffers():
                    for name in self._field:
                        live = self._rows(buf, name, self.N)
                        gathered = live.index_select(0, keep)  # temporary: source and destination overlap
                        live[:n_keep].copy_(gathered)
                        live[n_keep:].zero_()  # dead rows: zero gradient / moments, so FlatAdam leaves them alone
        n_pruned = self.N - n_keep
        if keep is None:  # (the three [N] statistics are re-sliced with one boolean index each)
            keep = (~mask).nonzero(as_tuple=True)[0] if any(
                getattr(self, a).shape[0] == self.N for a in ("max_radii2d", "mean_2d_grad_accum", "cnt")) else None
        # statistics: re-sliced when their length matches the mask, zeros otherwise (the IndexError branches)
        for attr in ("max_radii2d", "mean_2d_grad_accum", "cnt"):
            t = getattr(self, attr)
            setattr(self, attr, t.index_select(0, keep) if t.shape[0] == self.N else
                    torch.zeros(n_keep, dtype=torch.float32, device=self.device))
        self.N = n_keep
        self._mark_synced()  # callers synchronise before selecting rows, so the re-sliced statistics are rank-identical
        self._make_leaves()
        return n_pruned

    # ---- CUDA row movers (csrc/store.cu) -------------------------------------------------------------------
    def _field_tables(self):
        import ctypes

        names = list(self._field)
        offs = (ctypes.c_uint64 * len(names))(*[self._field[n][1] for n in names])
        widths = []
        for n in names:
            w = 1
            for s_ in self._field[n][0][1:]:
                w *= s_
            widths.append(w)
        return names, offs, (ctypes.c_uint32 * len(names))(*widths)

    def _compact_cuda(self, mask: torch.Tensor) -> int:
        import ctypes

        from . import _lib

        if self._shadow is None:
            self._shadow = [torch.zeros_like(b) for b in self._buffers()]
            self._shadow_rows = 0  # rows of the shadow buffers that may hold stale (non-zero) data
        names, offs, widths = self._field_tables()
        bufs = list(self._buffers())
        src = (ctypes.c_void_p * 4)(*[b.data_ptr() for b in bufs])
        dst = (ctypes.c_void_p * 4)(*[b.data_ptr() for b in self._shadow])
        m8 = mask.to(torch.bool).contiguous()
        n_keep = ctypes.c_uint32(0)
        _lib.check(_lib.lib().gsb200_store_compact(
            _lib.ctx(self.device), src, dst, ctypes.c_int32(4), offs, widths, ctypes.c_int32(len(names)),
            ctypes.c_uint32(self.N), ctypes.c_uint32(max(self._shadow_rows, self.N)), _lib.ptr(m8, torch.bool, "mask"),
            ctypes.byref(n_keep), _lib.stream_ptr(self.device)))
        # swap: the compacted copies become the arena, the old buffers (stale rows < N) the next shadow
        self._shadow, old = bufs, self._shadow
        self._shadow_rows = self.N
        self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq = old
        return int(n_keep.value)

    def _append_cuda(self, new_params: Dict[str, torch.Tensor], k: int):
        import ctypes

        from . import _lib

        names, offs, widths = self._field_tables()
        rows = [new_params[n].detach().to(self.device, torch.float32).reshape(k, -1).contiguous() for n in names]
        dst = (ctypes.c_void_p * 4)(*[b.data_ptr() for b in self._buffers()])
        src = (ctypes.c_void_p * len(names))(*[r.data_ptr() for r in rows])
        _lib.check(_lib.lib().gsb200_store_append(
            _lib.ctx(self.device), dst, ctypes.c_int32(4), src, offs, widths, ctypes.c_int32(len(names)),
            ctypes.c_uint32(self.N), ctypes.c_uint32(k), _lib.stream_ptr(self.device)))

    # ---- selection rules ---------------------------------------------------------------------------------
    def densify_by_clone(self, grads: torch.Tensor, grad_thresh: float, split_thresh: float,
                         mask: Optional[torch.Tensor] = None) -> int:
        """:614-628.  NOTE the reference's test is on `torch.norm(grads, dim=-1)` of the 1-D statistic -- a scene-wide
        scalar -- and is reproduced as such."""
        if mask is None:
            sel = torch.norm(grads, dim=-1) >= grad_thresh
            sel = torch.logical_and(sel, self.svec_act.max(dim=1).values <= split_thresh)
        else:
            sel = mask
        new = {name: self.params[name].detach()[sel] for name in self._field}
        self.append(new)
        return int(torch.count_nonzero(sel))

    def densify_by_split(self, grads: Optional[torch.Tensor], grad_thresh: Optional[float], split_thresh: float,
                         n_splits: int = 2, split_shrink: float = 0.8, mask: Optional[torch.Tensor] = None,
                         noise: Optional[torch.Tensor] = None) -> int:
        """:551-612.  `noise` replaces the reference's `torch.randn`: a tensor [n_selected * n_splits, 3] or a callable
        `n_rows -> tensor` (pass it for reproducible / rank-identical splits; drawn on the store's device otherwise)."""
        if mask is not None:
            sel = mask
        else:
            padded = torch.zeros(self.N, device=self.device)
            padded[: grads.shape[0]] = grads.reshape(-1)
            sel = torch.logical_and(padded >= grad_thresh, self.svec_act.max(dim=1).values > split_thresh)
        n_sel = int(torch.count_nonzero(sel))
        rep = lambda t: t.detach()[sel].repeat(n_splits, *([1] * (t.dim() - 1)))
        new_mean, new_qvec = rep(self.params["mean"]), rep(self.params["qvec"])
        new_svec = torch.exp(rep(self.params["svec"]))
        if noise is None:
            # the reference draws torch.randn here (:576-579); replicas of the view-parallel mode must draw the SAME
            # numbers or they diverge silently (N still matches): rank 0 draws, everybody receives
            noise = torch.randn(n_sel * n_splits, 3, device=self.device)
            if self.world_size() > 1:
                src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
                dist.broadcast(noise, src=src, group=self.group)
        elif callable(noise):
            noise = noise(n_sel * n_splits)
        if noise.shape[0] != n_sel * n_splits:
            raise RuntimeError(f"split noise has {noise.shape[0]} rows, {n_sel * n_splits} are needed")
        gn = noise.to(self.device) * new_svec
        rot_t = quat_to_rotmat(new_qvec).transpose(-1, -2)  # the reference multiplies by the TRANSPOSE (:575-580)
        new = {"mean": new_mean + torch.einsum("bij,bj->bi", rot_t, gn), "qvec": new_qvec,
               "svec": torch.log(new_svec / (n_splits * split_shrink))}
        for name in self._field:
            if name not in new:
                new[name] = rep(self.params[name])
        self.append(new)
        prune_mask = torch.cat((sel, torch.zeros(n_splits * n_sel, dtype=torch.bool, device=self.device)))
        self.prune_by_mask(prune_mask)
        return n_sel

    def densify_by_scale(self, scale_max: float, split_thresh: float, n_splits: int = 2, split_shrink: float = 0.8,
                         noise=None) -> int:
        """:630-632"""
        return self.densify_by_split(None, None, split_thresh, n_splits, split_shrink,
                                     mask=(self.svec_act > scale_max).any(dim=-1), noise=noise)

    def densify_official(self, mean2d_thresh: float, split_thresh: float, n_splits: int = 2,
                         split_shrink: float = 0.8, noise=None):
        """`densify()` with type "official" (:770-778, conf/renderer/regular.yaml:29-39): clone, then split with the
        clone-time statistic zero-padded, then reset the accumulators (:816-817)."""
        grads = self.mean_2d_grad_accum / self.cnt
        grads[grads.isnan()] = 0.0
        n_clone = self.densify_by_clone(grads, mean2d_thresh, split_thresh)
        n_split = self.densify_by_split(grads, mean2d_thresh, split_thresh, n_splits, split_shrink, noise=noise)
        self.reset_densify_info()
        return n_clone, n_split

    # ---- the legacy rule (gs/gaussian_splatting.py:820-946; `use_legacy: True` in conf/base.yaml, corgi.yaml, ...) ----
    def densify_legacy(self, mean2d_thresh: float, split_thresh: float, split_shrink: float = 0.8, noise=None):
        """`densify_legacy`: ONE selection `accum / (cnt + 1e-5) > mean2d_thresh`, split where any scale exceeds
        `split_thresh` (two children, scale / split_shrink / 2), clone elsewhere.  Result order as the reference builds
        it: the rows that are not split (clone sources included), the clones, the children (all first halves, then all
        second halves -- `repeat(2, 1)`).  The reference then RE-CREATES the optimizer (`set_optimizer(opt_cfg, step)`,
        :938): every Adam moment is dropped and the step counter restarts -- reproduced by zeroing both moment buffers
        and the optimizer's update count.  Returns (num_split, num_clone)."""
        mask = self.mean_2d_grad_accum / (self.cnt + 1e-5) > mean2d_thresh
        svec = self.svec_act
        split_mask = torch.logical_and(mask, (svec > split_thresh).any(dim=-1))
        clone_mask = torch.logical_and(mask, torch.logical_not(split_mask))
        n_split, n_clone = int(torch.count_nonzero(split_mask)), int(torch.count_nonzero(clone_mask))
        n_old = self.N
        rep = lambda t: t.detach()[split_mask].repeat(2, *([1] * (t.dim() - 1)))
        split_mean, split_qvec, split_svec = rep(self.params["mean"]), rep(self.params["qvec"]), svec[split_mask].repeat(2, 1)
        if noise is None:
            noise = torch.randn(n_split * 2, 3, device=self.device)  # (:859) -- replicas must draw the SAME numbers
            if self.world_size() > 1:
                src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
                dist.broadcast(noise, src=src, group=self.group)
        elif callable(noise):
            noise = noise(n_split * 2)
        if noise.shape[0] != n_split * 2:
            raise RuntimeError(f"split noise has {noise.shape[0]} rows, {n_split * 2} are needed")
        rot_t = quat_to_rotmat(split_qvec).transpose(-1, -2)
        children = {"mean": split_mean + torch.einsum("bij,bj->bi", rot_t, noise.to(self.device) * split_svec),
                    "qvec": split_qvec, "svec": torch.log(split_svec / split_shrink / 2.0)}
        clones = {}
        for name in self._field:
            clones[name] = self.params[name].detach()[clone_mask]
            if name not in children:
                children[name] = rep(self.params[name])
        self.append(clones)
        self.append(children)
        pad = torch.zeros(self.N - n_old, dtype=torch.bool, device=self.device)
        self.prune_by_mask(torch.cat((split_mask, pad)))
        for buf in (self.exp_avg, self.exp_avg_sq):  # the re-created optimizer starts without state
            buf.zero_()
        if self.optimizer is not None:
            self.optimizer.n_steps = 0
        return n_split, n_clone

    # ---- compactness-based densification (gs/gaussian_splatting.py:634-743) ----------------------------------
    def _knn(self):
        if self.knn_fn is not None:
            return self.knn_fn
        from .knn import knn_points

        return knn_points

    def densify_by_compatness_with_idx(self, idx: torch.Tensor) -> Dict[str, torch.Tensor]:
        """:634-680.  `idx` [N]: one neighbour per Gaussian.  Where the two ellipsoid "surfaces" (utils/ops.py:137-158)
        do not reach each other along the line of centres, a new Gaussian is put in the gap: centre in the middle of
        the gap, isotropic scale gap / 6, colour / orientation / opacity copied from the Gaussian itself."""
        from .knn import distance_to_gaussian_surface

        mean = self.params["mean"].detach()
        svec, rotmat = self.svec_act, quat_to_rotmat(self.params["qvec"].detach())
        nn_svec, nn_rotmat, nn_pos = svec[idx], rotmat[idx], mean[idx]
        nn_surface = distance_to_gaussian_surface(nn_pos, nn_svec, nn_rotmat, mean)
        surface = distance_to_gaussian_surface(mean, svec, rotmat, nn_pos)
        dist_to_nn = torch.norm(nn_pos - mean, dim=-1)
        mask = (surface + nn_surface) < dist_to_nn
        direction = (nn_pos - mean) / dist_to_nn[..., None]
        new_mean = (mean + direction * (dist_to_nn + surface - nn_surface)[..., None] / 2.0)[mask]
        gap = (dist_to_nn - surface - nn_surface)[mask]
        new = {"mean": new_mean, "qvec": self.params["qvec"].detach()[mask],
               "svec": torch.log(torch.ones_like(svec[mask]) * gap[..., None] / 6.0)}
        for name in self._field:  # alpha, color | sh: the raw leaves of the Gaussian itself
            if name not in new:
                new[name] = self.params[name].detach()[mask]
        return new

    def densify_by_compatness(self, K: int = 1) -> int:
        """:682-694: the K nearest neighbours of every Gaussian (`K_nearest_neighbors(mean, K + 1)`, column 0 = the
        Gaussian itself dropped), one gap test per neighbour rank, all new Gaussians appended at once."""
        from .knn import K_nearest_neighbors

        if self.N < 2:
            return 0
        _, idx = K_nearest_neighbors(self.params["mean"].detach(), K=K + 1, knn=self._knn())
        if int(idx.min()) < 0:
            raise RuntimeError(f"densify_by_compatness: K = {K} neighbours need more than {self.N} Gaussians")
        parts = [self.densify_by_compatness_with_idx(idx[:, i]) for i in range(K)]
        new = {name: torch.cat([p[name] for p in parts], dim=0) for name in parts[0]}
        return self.append(new)

    def densify_by_shrink_then_compatness(self, shrink_factor: float, K: int = 3) -> int:
        """:741-743: `self.svec = self.svec / shrink_factor` through the activation pair (log of the shrunk scale, as
        the reference's setter :125-130 stores it), then densify_by_compatness."""
        raw = self.params["svec"].detach()
        raw.copy_(torch.log(torch.exp(raw) / shrink_factor))
        return self.densify_by_compatness(K=K)

    def prune(self, radii2d_thresh: float = 0.0, alpha_thresh: float = 0.0, radii3d_thresh: float = 0.0):
        """`prune()` (:1152-1176) with the thresholds already evaluated for the step: by screen radius, then by
        opacity, then by 3-D scale, each on the survivors of the previous one."""
        n_scale = n_alpha = n_svec = 0
        if radii2d_thresh > 0.0:
            n_scale = self.prune_by_mask(self.max_radii2d > radii2d_thresh)
        if alpha_thresh > 0.0:
            n_alpha = self.prune_by_mask(self.alpha_act.reshape(self.N) < alpha_thresh)
        if radii3d_thresh > 0.0:
            n_svec = self.prune_by_mask((self.svec_act > radii3d_thresh).all(dim=-1))
        return n_scale, n_alpha, n_svec

    # ---- the trainer-facing dispatchers (step gating as the reference) -------------------------------------
    def densify_step(self, step: int, cfg) -> Optional[tuple]:
        """`densify(step)` (gs/gaussian_splatting.py:751-817), legacy and non-legacy branches: runs when
        `cfg.enabled`, `warm_up <= step <= end` and `step % period == 0` (step_check(..., run_at_zero=True)); resets
        the accumulators afterwards (:816-817).  cfg: mapping with the keys of conf/renderer/*.yaml `densify:`.
        Returns the counts of the operation that ran, or None."""
        from .renderer import step_check

        get = cfg.get if hasattr(cfg, "get") else (lambda k, d=None: getattr(cfg, k, d))
        if not get("enabled", False):
            return None
        if step < get("warm_up") or step > get("end") or not step_check(step, get("period"), True):
            return None
        self.sync_densify_info()  # view-parallel replicas: the selection below must see the whole batch's statistics
        kind = get("type", "official")
        if get("use_legacy", False):  # (:759-768): the legacy rule, then -- by substring of the type -- a compactness pass
            res = self.densify_legacy(get("mean2d_thresh"), get("split_thresh"), get("split_shrink", 0.8),
                                      noise=get("noise"))
            if "shrink_then_compatness" in kind:
                res = res + (self.densify_by_shrink_then_compatness(get("surface_shrink", 1.5), K=get("K", 3)),)
            elif "compatness" in kind:
                res = res + (self.densify_by_compatness(K=get("K", 3)),)
            self.reset_densify_info()
            return res
        if kind == "official":
            res = self.densify_official(get("mean2d_thresh"), get("split_thresh"), get("n_splits", 2),
                                        get("split_shrink", 0.8), noise=get("noise"))
            return res  # densify_official already reset the accumulators
        if kind == "scale":
            n = self.densify_by_scale(get("scale_max"), get("split_thresh"), get("n_splits", 2),
                                      get("split_shrink", 0.8), noise=get("noise"))
        elif kind == "all":
            n = self.densify_by_split(None, None, get("split_thresh"), 2, get("split_shrink", 0.8),
                                      mask=torch.ones(self.N, dtype=torch.bool, device=self.device),
                                      noise=get("noise"))
        elif kind == "compatness":  # (:790-797)
            n = self.densify_by_compatness(K=get("K", 3))
        elif kind == "shrink_then_compatness":  # (:806-810)
            n = self.densify_by_shrink_then_compatness(get("surface_shrink", 1.5), K=get("K", 3))
        else:
            raise NotImplementedError(f"Unknown densify type: {kind}")
        self.reset_densify_info()
        return (n,)

    def prune_step(self, step: int, cfg) -> Optional[tuple]:
        """`prune(step)` (:1152-1176): runs when `cfg.enabled`, `warm_up <= step <= end` and `step != 0 and
        step % period == 0`; thresholds may be schedules in the reference (`C(value, step)`), pass numbers here."""
        from .renderer import step_check

        get = cfg.get if hasattr(cfg, "get") else (lambda k, d=None: getattr(cfg, k, d))
        if not get("enabled", False):
            return None
        if step < get("warm_up") or step > get("end") or not step_check(step, get("period")):
            return None
        self.sync_densify_info()
        return self.prune(get("radii2d_thresh", 0.0), get("alpha_thresh", 0.0), get("radii3d_thresh", 0.0) or 0.0)
