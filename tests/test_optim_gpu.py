"""-m gpu parity of the flat Adam step (gsb200_adam_step through gsgen_b200.optim.FlatAdam) against
torch.optim.Adam configured as the reference does (gs/gaussian_splatting.py:398-419: one param group per field,
conf/base.yaml:8-26: eps 1e-15, per-field lr schedules rewritten by update_lr every step)."""
import pytest
import torch

from gsgen_b200.optim import FlatAdam
from gsgen_b200.parallel import field_layout

pytestmark = pytest.mark.gpu
DEV = "cuda"
LR = {"mean": [0.005, 3.0e-05, 15000, "exp"], "svec": [0.003, 0.001, 15000, "exp"], "qvec": 0.003, "color": 0.01,
      "sh": 0.01, "alpha": 0.003}


@pytest.mark.parametrize("N,C", [(1001, None), (4096, 2), (37, 4)])  # 1001 / 37: field boundaries inside a float4 + tail
def test_flat_adam_matches_torch_adam(N, C):
    layout = field_layout(N, C)
    total = layout[-1][2] + layout[-1][3]
    g = torch.Generator().manual_seed(N)
    flat_p = torch.randn(total, generator=g).to(DEV)
    flat_g = torch.zeros(total, device=DEV)
    ours = FlatAdam(flat_p, flat_g, layout, LR)
    ref_params = {name: flat_p[off:off + n].clone().view(shape).requires_grad_() for name, shape, off, n in layout}
    p0 = flat_p.clone()
    opt = torch.optim.Adam([{"params": [p], "lr": 0.0, "name": name} for name, p in ref_params.items()], lr=0.0,
                           eps=1e-15)
    for step in range(6):
        grad = torch.randn(total, generator=g).to(DEV) * (10.0 ** float(torch.randint(-5, 2, (1,), generator=g)))
        grad[::5] = 0.0
        flat_g.copy_(grad)
        train_step = step * 1500
        lrs = ours.lr_at(train_step)
        for grp in opt.param_groups:  # update_lr(step), gaussian_splatting.py:451-454
            grp["lr"] = lrs[grp["name"]]
        for name, shape, off, n in layout:
            ref_params[name].grad = grad[off:off + n].view(shape).clone()
        opt.step()
        ours.step(train_step)
        for name, shape, off, n in layout:
            st = opt.state[ref_params[name]]
            m, v = ours.exp_avg[off:off + n], ours.exp_avg_sq[off:off + n]
            rm, rv = st["exp_avg"].reshape(-1), st["exp_avg_sq"].reshape(-1)
            assert torch.allclose(m, rm, rtol=2e-6, atol=2e-7 * float(rm.abs().max())), (name, step)
            assert torch.allclose(v, rv, rtol=2e-6, atol=1e-38), (name, step)
            d = (flat_p[off:off + n] - ref_params[name].detach().reshape(-1)).abs().max()
            # an update is lr-sized: bound the drift by fp32 rounding of the accumulated step length
            assert float(d) <= 4e-6 * 0.01 * (step + 1) + 5e-7 * float(p0.abs().max()), (name, step, float(d))
    assert torch.isfinite(flat_p).all()


def test_flat_adam_rejects_bad_layouts():
    flat_p, flat_g = torch.zeros(64, device=DEV), torch.zeros(64, device=DEV)
    with pytest.raises(RuntimeError):
        FlatAdam(flat_p, flat_g, [("a", (32,), 0, 32), ("b", (16,), 40, 16)], {"a": 0.1, "b": 0.1}).step()  # gap
    with pytest.raises(RuntimeError):
        FlatAdam(flat_p, flat_g, [("a", (32,), 0, 32)], {"a": 0.1}).step()  # does not cover the buffer
    with pytest.raises(RuntimeError):
        FlatAdam(flat_p, flat_g, [("a", (64,), 0, 64)], {})  # no lr
