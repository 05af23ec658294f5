"""-m gpu: the hand-written row movers of the Gaussian arena (csrc/store.cu: gsb200_store_compact / _append) against
the torch-indexing path the CPU test-suite pins to the reference's own densify / prune trace (tests/test_store_cpu.py):
same operations on a CUDA store and on a CPU twin, all four buffers (parameters, gradients, both Adam moments) compared
bit for bit, dead rows zero."""
import pytest
import torch

from gsgen_b200.store import GaussianStore

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _twin(N, C, cap=None, seed=0):
    g = torch.Generator().manual_seed(seed)
    p = dict(mean=torch.randn(N, 3, generator=g), qvec=torch.randn(N, 4, generator=g),
             svec=torch.randn(N, 3, generator=g), alpha=torch.randn(N, generator=g))
    if C is None:
        p["color"] = torch.randn(N, 3, generator=g)
    else:
        p["sh"] = torch.randn(N, 3, C * C, generator=g)
    a, b = GaussianStore(p, C, "cpu", capacity=cap), GaussianStore(p, C, DEV, capacity=cap)
    for st in (a, b):  # give gradients and moments distinct contents so that a mixed-up buffer shows
        for k, buf in enumerate((st.flat_grad, st.exp_avg, st.exp_avg_sq)):
            for name in st._field:
                rows = st._rows(buf, name, st.N)
                rows.copy_((k + 2.0) * st._rows(st.flat_param, name, st.N) + 0.5)
    return a, b, g


def _same(a, b):
    assert a.N == b.N and a.cap == b.cap
    for x, y, what in zip(a._buffers(), b._buffers(), ("param", "grad", "exp_avg", "exp_avg_sq")):
        for name in a._field:
            assert torch.equal(a._rows(x, name, a.N), b._rows(y, name, b.N).cpu()), (what, name)
            dead = b._rows(y, name, b.cap - b.N, b.N)
            assert dead.numel() == 0 or float(dead.abs().max()) == 0.0, (what, name, "dead rows not zero")
    for attr in ("mean_2d_grad_accum", "cnt", "max_radii2d"):
        assert torch.equal(getattr(a, attr), getattr(b, attr).cpu()), attr


@pytest.mark.parametrize("C", [None, 4])
def test_compact_and_append_match_the_torch_path(C):
    a, b, g = _twin(5003, C, cap=6000)
    for rnd in range(4):
        for st in (a, b):  # statistics with content, so that their re-slicing is checked too
            z = torch.arange(st.N, dtype=torch.float32)
            st.mean_2d_grad_accum, st.cnt, st.max_radii2d = z.to(st.device) * 0.5, z.to(st.device) + 1, z.to(st.device) * 2
        mask = torch.rand(a.N, generator=g) < (0.05 + 0.3 * rnd)
        assert a.prune_by_mask(mask) == b.prune_by_mask(mask.to(DEV))
        _same(a, b)
        k = 700 + 1500 * rnd  # the later rounds exceed the capacity: growth re-allocates (and drops the shadow buffers)
        new = {name: torch.randn(k, *shape[1:], generator=g) for name, (shape, _) in a._field.items()}
        a.append(new)
        b.append({n: v.to(DEV) for n, v in new.items()})
        for st in (a, b):
            st.reset_densify_info()
        _same(a, b)


def test_prune_everything_and_nothing():
    a, b, g = _twin(257, None)
    for mask in (torch.zeros(257, dtype=torch.bool), torch.ones(257, dtype=torch.bool)):
        a.prune_by_mask(mask)
        b.prune_by_mask(mask.to(DEV))
        _same(a, b)
    assert b.N == 0


def test_official_densify_round_on_the_gpu_matches_the_cpu_twin():
    """clone + split (with the same noise) + prune: selection rules on the device, rows moved by the kernels"""
    a, b, g = _twin(4000, None, seed=3)
    for st in (a, b):
        z = torch.linspace(0, 2e-3, st.N)
        st.mean_2d_grad_accum, st.cnt = z.to(st.device), torch.ones(st.N, device=st.device)
        st.max_radii2d = torch.zeros(st.N, device=st.device)
    n_split_rows = None

    def noise(n):
        return torch.randn(n, 3, generator=torch.Generator().manual_seed(7))

    ra = a.densify_official(1e-3, 1.0, 2, 0.8, noise=noise)
    rb = b.densify_official(1e-3, 1.0, 2, 0.8, noise=lambda n: noise(n).to(DEV))
    assert ra == rb and ra[0] + ra[1] > 0
    assert a.N == b.N
    for name in a._field:  # exp / log / rotation run in fp32 on both devices: not bit-identical, 1e-5 like the CPU pin
        assert torch.allclose(a.params[name].detach(), b.params[name].detach().cpu(), rtol=1e-5, atol=1e-6), name
    assert a.prune(alpha_thresh=0.3) == b.prune(alpha_thresh=0.3)
    assert a.N == b.N
