"""Densify / prune on the capacity arena (gsgen_b200/store.py) against a fixture produced by the
UNMODIFIED reference methods (`densify_by_clone`, `densify_by_split`, `prune_by_mask`, `prune_optimizer`,
`densify_on_optimizer` ... run by tests/golden/make_densify_golden.py in the dev container): parameters, both Adam
moments and the densification statistics after every stage of one "official" densify + prune round."""
import os

import numpy as np
import pytest
import torch

from gsgen_b200.store import GaussianStore, quat_to_rotmat

GOLD = os.path.join(os.path.dirname(__file__), "golden", "densify_official.npz")
FIELDS = ("mean", "qvec", "svec", "alpha", "color")


def _load():
    z = np.load(GOLD)
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _store_from(gold, tag, **kw):
    st = GaussianStore({f: gold[f"{tag}_{f}"] for f in FIELDS}, C=None, device="cpu", **kw)
    for f in FIELDS:
        n = st.N
        st._rows(st.exp_avg, f, n).copy_(gold[f"{tag}_{f}_exp_avg"].reshape(n, -1))
        st._rows(st.exp_avg_sq, f, n).copy_(gold[f"{tag}_{f}_exp_avg_sq"].reshape(n, -1))
    for s in ("max_radii2d", "mean_2d_grad_accum", "cnt"):
        setattr(st, s, gold[f"{tag}_{s}"].clone())
    return st


def _check(st, gold, tag, exact=True):
    assert st.N == int(gold[f"{tag}_N"])
    for f in FIELDS:
        n = st.N
        for buf, suffix in ((st.flat_param, ""), (st.exp_avg, "_exp_avg"), (st.exp_avg_sq, "_exp_avg_sq")):
            ours = st._rows(buf, f, n)
            ref = gold[f"{tag}_{f}{suffix}"].reshape(n, -1)
            if exact or suffix or f not in ("mean", "svec"):
                assert torch.equal(ours, ref), (tag, f, suffix)
            else:  # split children: exp / log / rotation evaluated by both sides in fp32
                assert torch.allclose(ours, ref, rtol=1e-5, atol=1e-6), (tag, f, float((ours - ref).abs().max()))
        assert st.params[f].shape[0] == n and st.params[f].requires_grad
        assert st.params[f].data_ptr() == st._rows(st.flat_param, f, n).data_ptr()  # leaves are views of the arena
    for s in ("max_radii2d", "mean_2d_grad_accum", "cnt"):
        assert torch.equal(getattr(st, s), gold[f"{tag}_{s}"]), (tag, s)


@pytest.mark.parametrize("capacity", [None, 2000])  # None: the arena has to grow twice; 2000: everything in place
def test_official_densify_then_prune_matches_reference(capacity):
    gold = _load()
    st = _store_from(gold, "s0", capacity=capacity)
    _check(st, gold, "s0")
    grads = st.mean_2d_grad_accum / st.cnt
    grads[grads.isnan()] = 0.0
    assert torch.equal(grads, gold["s0_grads"])
    n_clone = st.densify_by_clone(grads, 0.02, 0.02)
    _check(st, gold, "s1")
    n_split = st.densify_by_split(grads, 0.02, 0.02, n_splits=2, split_shrink=0.8, noise=gold["noise"])
    _check(st, gold, "s2", exact=False)
    assert [n_clone, n_split] == gold["s2_counts"].tolist()
    # dead capacity rows stay zero in the gradient and moment buffers (FlatAdam then leaves them untouched)
    if st.cap > st.N:
        for f in FIELDS:
            for buf in (st.flat_grad, st.exp_avg, st.exp_avg_sq):
                assert float(st._rows(buf, f, st.cap - st.N, st.N).abs().max()) == 0.0
    st.reset_densify_info()
    # prune round; the split children differ from the reference's in the last bits (exp/log), so continue from the
    # reference's own post-densify state to compare the prune stages exactly
    st = _store_from(gold, "s2", capacity=capacity)
    st.max_radii2d = gold["s2_max_radii2d_for_prune"].clone()
    st.mean_2d_grad_accum, st.cnt = torch.zeros(st.N), torch.zeros(st.N)
    counts = []
    counts.append(st.prune_by_mask(st.max_radii2d > 1.0))
    _check(st, gold, "s3")
    counts.append(st.prune_by_mask(st.alpha_act.reshape(st.N) < 0.05))
    _check(st, gold, "s4")
    counts.append(st.prune_by_mask((st.svec_act > 0.012).all(dim=-1)))
    _check(st, gold, "s5")
    assert counts == gold["prune_counts"].tolist()


def test_prune_wrapper_and_official_wrapper_agree_with_the_staged_calls():
    gold = _load()
    a = _store_from(gold, "s0")
    n_clone, n_split = a.densify_official(0.02, 0.02, 2, 0.8, noise=gold["noise"])
    assert [n_clone, n_split] == gold["s2_counts"].tolist() and a.N == int(gold["s2_N"])
    assert float(a.cnt.abs().max()) == 0.0 and a.cnt.shape[0] == a.N  # reset_densify_info
    b = _store_from(gold, "s2")
    b.max_radii2d = gold["s2_max_radii2d_for_prune"].clone()
    assert list(b.prune(1.0, 0.05, 0.012)) == gold["prune_counts"].tolist()
    assert b.N == int(gold["s5_N"])


def test_update_densify_info_and_sh_store():
    g = torch.Generator().manual_seed(1)
    N, C = 50, 3
    params = dict(mean=torch.randn(N, 3, generator=g), qvec=torch.randn(N, 4, generator=g),
                  svec=torch.randn(N, 3, generator=g), alpha=torch.randn(N, generator=g),
                  sh=torch.randn(N, 3, C * C, generator=g))
    st = GaussianStore(params, C=C, device="cpu")
    mask = torch.rand(N, generator=g) > 0.5
    g2d = torch.randn(N, 2, generator=g)
    radii = torch.rand(N, generator=g)
    st.update_densify_info(mask, g2d, radii)
    st.update_densify_info(mask, g2d, radii * 0.5)
    assert torch.allclose(st.mean_2d_grad_accum[mask], 2 * g2d[mask].norm(dim=-1))
    assert torch.equal(st.cnt, 2.0 * mask.float()) and torch.equal(st.max_radii2d, radii * mask)
    keep = torch.rand(N, generator=g) > 0.3
    st.prune_by_mask(~keep)
    assert torch.equal(st.params["sh"].detach(), params["sh"][keep]) and st.params["sh"].shape == (int(keep.sum()), 3, 9)
    st.append({k: v[:5] for k, v in params.items()})
    assert torch.equal(st.params["sh"].detach()[-5:], params["sh"][:5])
    # gradients accumulate into the flat buffer through the leaves
    st.zero_grad()
    (st.params["mean"] * 2.0).sum().backward()
    assert float(st._rows(st.flat_grad, "mean", st.N).min()) == 2.0


def test_quat_to_rotmat_is_a_rotation_and_normalises():
    q = torch.randn(64, 4, generator=torch.Generator().manual_seed(2)) * 3.0
    R = quat_to_rotmat(q)
    assert torch.allclose(R @ R.transpose(-1, -2), torch.eye(3).expand(64, 3, 3), atol=1e-5)
    assert torch.allclose(torch.linalg.det(R), torch.ones(64), atol=1e-5)
    assert torch.allclose(R, quat_to_rotmat(q / q.norm(dim=-1, keepdim=True)), atol=1e-6)


def test_checkpoint_roundtrip(tmp_path):
    gold = _load()
    st = _store_from(gold, "s0", capacity=1000)
    saved = st.get_params_for_save()
    assert set(saved) == set(FIELDS) and all(not v.requires_grad for v in saved.values())
    saved["mean"][0, 0] += 1.0  # a checkpoint does not alias the arena
    assert float(st.params["mean"].detach()[0, 0]) != float(saved["mean"][0, 0])
    path = str(tmp_path / "step_10.pt")
    torch.save({"params": {**st.get_params_for_save(), "cfg": {"x": 1}, "bg": {}}, "cfg": {}, "step": 10}, path)
    st2 = GaussianStore.load(path, None, "cpu")
    for f in FIELDS:
        assert torch.equal(st2.params[f].detach(), st.params[f].detach())
    with pytest.raises(RuntimeError):
        GaussianStore.load({"mean": saved["mean"]}, None, "cpu")


def test_step_gated_dispatchers_follow_the_reference_trace():
    """densify(step) / prune(step) of the reference (gs/gaussian_splatting.py:751-817, :1152-1176), run by the fixture
    generator over six steps around the warm-up / period / end boundaries: same N after every call, same final rows."""
    gold = _load()
    st = _store_from(gold, "s0")
    dens = dict(enabled=True, type="official", warm_up=2000, end=4999, period=500, mean2d_thresh=0.02, split_thresh=0.02,
                n_splits=2, split_shrink=0.8, use_legacy=False)
    prun = dict(enabled=True, warm_up=0, end=15000, period=100, radii2d_thresh=1.0, alpha_thresh=0.05,
                radii3d_thresh=0.012)
    pool = {"at": 0}

    def noise(n):
        out = gold["dispatch_noise"][pool["at"]: pool["at"] + n]
        pool["at"] += n
        return out

    dens["noise"] = noise
    for step, n_before, n_mid, n_after in gold["dispatch_trace"].tolist():
        assert st.N == n_before
        st.mean_2d_grad_accum = gold["s0_mean_2d_grad_accum"].clone()[: st.N]
        st.cnt = gold["s0_cnt"].clone()[: st.N]
        st.max_radii2d = torch.linspace(0.0, 1.3, st.N)
        st.densify_step(step, dens)
        assert st.N == n_mid, (step, st.N, n_mid)
        st.max_radii2d = torch.linspace(0.0, 1.3, st.N)
        st.prune_step(step, prun)
        assert st.N == n_after, (step, st.N, n_after)
    assert pool["at"] == gold["dispatch_noise"].shape[0]
    _check_rows_only(st, gold, "s9")


def _check_rows_only(st, gold, tag):
    for f in FIELDS:
        n = st.N
        for buf, suffix in ((st.flat_param, ""), (st.exp_avg, "_exp_avg"), (st.exp_avg_sq, "_exp_avg_sq")):
            ours, ref = st._rows(buf, f, n), gold[f"{tag}_{f}{suffix}"].reshape(n, -1)
            assert torch.allclose(ours, ref, rtol=1e-5, atol=1e-6), (f, suffix, float((ours - ref).abs().max()))
