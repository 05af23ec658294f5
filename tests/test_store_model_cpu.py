"""Model-based test of the capacity arena (gsgen_b200/store.py): random sequences of the row operations the densify /
prune rules are made of -- append (clone / split / compactness children), prune by mask, capacity growth, statistics
updates and resets, the legacy optimizer reset -- against the reference's data model, i.e. plain tensors rebuilt with
`torch.cat` and boolean indexing (what `densify_on_optimizer` / `prune_optimizer` / `prune_by_mask` do,
gs/gaussian_splatting.py:421-549).  After every operation: all four buffers' live rows equal the model's tensors bit
for bit, dead capacity rows of gradients and moments are zero, leaves are views of the arena, field offsets stay
16-byte aligned."""
import pytest
import torch

from gsgen_b200.store import GaussianStore


class Model:
    """parameters, gradients and Adam moments as independent tensors per field"""

    def __init__(self, params):
        self.t = {k: [v.clone().reshape(v.shape[0], -1), torch.zeros(v.shape[0], v[0].numel()),
                      torch.zeros(v.shape[0], v[0].numel()), torch.zeros(v.shape[0], v[0].numel())]
                  for k, v in params.items()}

    @property
    def N(self):
        return self.t["mean"][0].shape[0]

    def append(self, new):
        for k, bufs in self.t.items():
            rows = new[k].reshape(new[k].shape[0], bufs[0].shape[1])
            bufs[0] = torch.cat([bufs[0], rows])
            for b in (1, 2, 3):
                bufs[b] = torch.cat([bufs[b], torch.zeros_like(rows)])

    def prune(self, mask):
        for bufs in self.t.values():
            for b in range(4):
                bufs[b] = bufs[b][~mask]


def _check(st, m):
    assert st.N == m.N and st.cap >= st.N and st.cap % 4 == 0
    for name, shape, off, n in st.layout:
        assert off % 4 == 0, (name, off)
    for k, bufs in m.t.items():
        for b, buf in enumerate(st._buffers()):
            assert torch.equal(st._rows(buf, k, st.N), bufs[b]), (k, b)
            if b and st.cap > st.N:
                assert float(st._rows(buf, k, st.cap - st.N, st.N).abs().max()) == 0.0, (k, b, "dead rows")
        assert st.params[k].data_ptr() == st._rows(st.flat_param, k, st.N).data_ptr() and st.params[k].requires_grad
        assert st.params[k].grad is not None and st.params[k].grad.data_ptr() == st._rows(st.flat_grad, k, st.N).data_ptr()
    for a in ("cnt", "mean_2d_grad_accum", "max_radii2d"):
        assert getattr(st, a).shape[0] == st.N


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("C", [None, 2])
def test_random_operation_sequences_match_the_tensor_model(seed, C):
    g = torch.Generator().manual_seed(1000 * (C or 0) + seed)
    N = int(torch.randint(1, 40, (1,), generator=g))
    mk = lambda n: dict(mean=torch.randn(n, 3, generator=g), qvec=torch.randn(n, 4, generator=g),
                        svec=torch.randn(n, 3, generator=g), alpha=torch.randn(n, generator=g),
                        **({"color": torch.randn(n, 3, generator=g)} if C is None else
                           {"sh": torch.randn(n, 3, C * C, generator=g)}))
    p = mk(N)
    cap = [None, N + 3, 4 * N][seed % 3]
    st, m = GaussianStore(p, C, "cpu", capacity=cap), Model(p)
    _check(st, m)
    for it in range(25):
        op = int(torch.randint(0, 6, (1,), generator=g))
        if op == 0:  # append (possibly beyond the capacity -> growth)
            k = int(torch.randint(0, 30, (1,), generator=g))
            new = mk(k)
            st.append(new); m.append(new)
            st.reset_densify_info()  # (the statistics are not extended by an append: densify() resets them, :816-817)
        elif op == 1 and st.N > 0:  # prune a random subset (possibly everything but one row, possibly nothing)
            mask = torch.rand(st.N, generator=g) < float(torch.rand(1, generator=g))
            if bool(mask.all()):
                mask[0] = False
            st.prune_by_mask(mask); m.prune(mask)
        elif op == 2:  # gradients and moments receive content through the arena views (as backward / Adam would)
            for k, bufs in m.t.items():
                for b in (1, 2, 3):
                    v = torch.randn(bufs[b].shape, generator=g)
                    bufs[b] = v.clone()
                    st._rows(st._buffers()[b], k, st.N).copy_(v)
        elif op == 3:  # zero_grad
            st.zero_grad()
            for bufs in m.t.values():
                bufs[1] = torch.zeros_like(bufs[1])
        elif op == 4:  # statistics of a view, then reset
            mask = torch.rand(st.N, generator=g) < 0.5
            st.update_densify_info(mask, torch.randn(st.N, 2, generator=g), torch.rand(st.N, generator=g))
            if it % 2:
                st.reset_densify_info()
        elif op == 5 and st.N > 1:  # clone-like: append copies of selected rows (densify_by_clone with a mask)
            sel = torch.rand(st.N, generator=g) < 0.3
            n_sel = st.densify_by_clone(None, 0.0, 0.0, mask=sel)
            m.append({k: bufs[0][sel].reshape(-1, *p[k].shape[1:]) for k, bufs in m.t.items()})
            assert n_sel == int(sel.sum())
            st.reset_densify_info()
        _check(st, m)
