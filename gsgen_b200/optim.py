"""The optimizer step that follows the gradient all-reduce.

Reference: `GaussianSplattingRenderer.set_optimizer` builds `torch.optim.Adam(eps=1e-15)` with one param group per
field and a per-field learning-rate schedule (gs/gaussian_splatting.py:268-292, 398-419; conf/base.yaml:8-26), and
`update_lr(step)` rewrites every group's lr before each step (:451-454).  torch runs that as ~5 foreach passes per
group over separately allocated tensors.  Here the fields already live back to back in one flat fp32 buffer whose
twin is the NCCL all-reduce operand (`gsgen_b200.parallel.ViewParallelRenderer`), so the update is ONE streaming
kernel (`gsb200_adam_step`, include/gsb200.h Part 4): 16 B read + 12 B written per parameter.

`FlatAdam` mirrors the torch optimizer's interface as far as the trainer uses it (`step()`, `zero_grad()`,
per-group lr), with state kept as two flat buffers.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import math
from typing import Callable, Dict, Sequence, Union

import torch

from . import _lib
from ._lib import Gsb200AdamField, fptr

LrSpec = Union[float, Sequence]


def exp_decay(tot_steps, lr_start, lr_end, warmup_steps=0) -> Callable[[int], float]:
    """utils/schedulers.py:6-20 (linear warm-up, then log-linear interpolation, clipped to [0, 1])."""
    def _decay(step):
        if step < warmup_steps:
            return lr_start * (step / warmup_steps)
        t = min(max((step - warmup_steps) / (tot_steps - warmup_steps), 0.0), 1.0)
        return math.exp(math.log(lr_start) * (1 - t) + math.log(lr_end) * t)

    return _decay


def cosine_decay(tot_steps, lr_start, lr_end, warmup_steps=0) -> Callable[[int], float]:
    """utils/schedulers.py:23-31 (progress is NOT clipped there either)."""
    def _decay(step):
        if step < warmup_steps:
            return lr_start * (step / warmup_steps)
        progress = (step - warmup_steps) / (tot_steps - warmup_steps)
        return lr_end + (lr_start - lr_end) * (1 + math.cos(math.pi * progress)) / 2

    return _decay


def no_decay(tot_steps, lr_start, lr_end, warmup_steps=0) -> Callable[[int], float]:
    """utils/schedulers.py:34-35"""
    return lambda step: lr_start


lr_schedulers = dict(nothing=no_decay, cosine=cosine_decay, exp=exp_decay)


class CompanionAdam:
    """torch.optim.Adam for the few parameters that live outside the Gaussian arena -- the background module's, which
    the reference puts into the SAME optimizer as param group "bg" with its own scheduler
    (gs/gaussian_splatting.py:383-419, conf/base.yaml:26 `lr.bg`).  Driven by FlatAdam.step / zero_grad through
    `FlatAdam.companions`; the schedule is evaluated at the trainer's step like every other group's (:451-454)."""

    def __init__(self, params, lr: LrSpec, max_steps: int = 15000, betas=(0.9, 0.999), eps: float = 1e-15):
        self.params = [p for p in params]
        self.scheduler = make_scheduler(lr, max_steps)
        self.opt = torch.optim.Adam(self.params, lr=float(self.scheduler(0)), betas=betas, eps=eps)

    def step(self, train_step: int):
        lr = float(self.scheduler(train_step))
        for g in self.opt.param_groups:
            g["lr"] = lr
        self.opt.step()
        return lr

    def zero_grad(self):
        self.opt.zero_grad()


def make_scheduler(spec: LrSpec, max_steps: int) -> Callable[[int], float]:
    """A config entry -> scheduler, as gs/gaussian_splatting.py:268-292 reads `cfg.lr.<field>`:
    a number is a constant lr; `[start, end, steps, type]` (conf/base.yaml:13-22) selects a schedule."""
    if isinstance(spec, (int, float)):
        return no_decay(max_steps, float(spec), float(spec))
    spec = list(spec)
    if len(spec) != 4:
        raise RuntimeError(f"lr spec must be a number or [start, end, steps, type], got {spec}")
    lr_start, lr_end, steps, kind = spec
    if kind not in lr_schedulers:
        raise RuntimeError(f"unknown lr schedule '{kind}'")
    return lr_schedulers[kind](int(steps), float(lr_start), float(lr_end))


class FlatAdam:
    """Adam over `flat_param` / `flat_grad` (1-D fp32 CUDA tensors of equal length) split into named fields.

    layout: [(name, shape, offset, numel)] as `gsgen_b200.parallel.field_layout` returns it (fields tile the buffer).
    lr:     name -> number | [start, end, steps, type]  (the reference's `cfg.lr`, conf/base.yaml:12-26)
    """

    def __init__(self, flat_param: torch.Tensor, flat_grad: torch.Tensor, layout, lr: Dict[str, LrSpec],
                 max_steps: int = 15000, betas=(0.9, 0.999), eps: float = 1e-15, state=None):
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.schedulers = {}
        for name, _, _, _ in layout:
            if name not in lr:
                raise RuntimeError(f"no learning rate for field '{name}'")
            self.schedulers[name] = make_scheduler(lr[name], max_steps)
        self.n_steps = 0  # torch: state[p]["step"], shared by all rows (the reference keeps it across densify / prune)
        # moments (torch: state[p]["exp_avg"], ["exp_avg_sq"]); `state` = buffers owned by a GaussianStore
        exp_avg, exp_avg_sq = state if state is not None else (torch.zeros_like(flat_param),
                                                                torch.zeros_like(flat_param))
        self.rebind(flat_param, flat_grad, layout, exp_avg, exp_avg_sq)

    def rebind(self, flat_param, flat_grad, layout, exp_avg, exp_avg_sq):
        """Point the optimizer at (re-allocated or re-laid-out) buffers; the step counter and schedules are kept.
        Called by `GaussianStore` after densify / prune, which is all the optimizer surgery the reference's
        `densify_on_optimizer` / `prune_optimizer` (gs/gaussian_splatting.py:421-449, :481-522) amount to here."""
        if flat_param.dim() != 1 or any(t.shape != flat_param.shape for t in (flat_grad, exp_avg, exp_avg_sq)):
            raise RuntimeError("flat_param, flat_grad and the moment buffers must be 1-D tensors of the same length")
        if len(layout) > 8:
            raise RuntimeError("at most 8 fields")
        self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq = flat_param, flat_grad, exp_avg, exp_avg_sq
        self.layout = list(layout)
        for name, _, _, _ in self.layout:
            if name not in self.schedulers:
                raise RuntimeError(f"no learning rate for field '{name}'")
        self._fields = (Gsb200AdamField * len(self.layout))()
        total = flat_param.numel()
        for i, (_, _, off, n) in enumerate(self.layout):
            # the C ABI wants fields that tile [0, total): a field's count runs up to the next field's (16-byte aligned)
            # offset, i.e. includes the alignment padding -- zero gradient and zero moments there, so nothing moves
            nxt = self.layout[i + 1][2] if i + 1 < len(self.layout) else total
            if nxt < off + n:
                raise RuntimeError("layout fields overlap or exceed the buffer")
            if nxt - (off + n) >= 4:  # only 16-byte alignment padding may separate fields (parallel.field_layout)
                raise RuntimeError(f"layout leaves {nxt - (off + n)} unowned floats after field {i}: fields must tile "
                                   "the buffer up to alignment padding")
            self._fields[i].begin, self._fields[i].count = off, nxt - off

    def lr_at(self, step: int) -> Dict[str, float]:
        """what update_lr(step) writes into the param groups (gs/gaussian_splatting.py:451-454)"""
        return {name: float(self.schedulers[name](step)) for name, _, _, _ in self.layout}

    def zero_grad(self):
        self.flat_grad.zero_()
        for c in getattr(self, "companions", ()):  # parameters outside the arena (the background module's)
            c.zero_grad()

    def step(self, train_step: int = None, grad_scale: float = 1.0):
        """One Adam update.  `train_step` is the trainer's step counter the schedules are evaluated at (defaults to
        the number of updates done so far, i.e. update_lr(step) followed by optimizer.step())."""
        if train_step is None:  # the trainer's step set by GaussianSplattingRenderer.update(step), else the update count
            train_step = getattr(self, "train_step", None)
        if train_step is None:
            train_step = self.n_steps
        lrs = self.lr_at(train_step)
        for i, (name, _, _, _) in enumerate(self.layout):
            self._fields[i].lr = lrs[name]
        self.n_steps += 1
        dev = self.flat_param.device
        _lib.check(_lib.lib().gsb200_adam_step(
            _lib.ctx(dev), fptr(self.flat_param, "flat_param"), fptr(self.flat_grad, "flat_grad"),
            fptr(self.exp_avg, "exp_avg"), fptr(self.exp_avg_sq, "exp_avg_sq"),
            ctypes.c_uint64(self.flat_param.numel()), self._fields, ctypes.c_int32(len(self.layout)),
            ctypes.c_double(self.betas[0]), ctypes.c_double(self.betas[1]), ctypes.c_double(self.eps),
            ctypes.c_int64(self.n_steps), ctypes.c_float(grad_scale), _lib.stream_ptr(dev)))
        for c in getattr(self, "companions", ()):  # the reference's Adam also holds a "bg" param group (:383-396)
            c.step(train_step)
        return lrs
