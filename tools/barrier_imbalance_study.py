"""Design study (CPU, numpy; no GPU): how much can a barrier-free schedule gain in the composite kernels?

For a sample of tiles of BASELINE C3 the script replays the kernels' control flow on the CPU from the oracle's
binning output -- per warp (8x4 pixel block), per staged batch of B list entries: which entries pass the warp's
bounding-box ballot and have at least one live lane with a*G >= 1/255 ("hits", the unit of work: ~100 warp
instructions in the forward, ~240 in the backward) -- and compares three schedules of the same work:

    sync      sum over batches of max over the 8 warps          (today: block barrier(s) every batch)
    decoupled max over warps of its own sum                      (a warp never waits for another one)
    balanced  mean over warps of its own sum                     (perfect balance inside the tile; lower bound)

    python tools/barrier_imbalance_study.py [--tiles 400] [--batch 32]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402  (analysis tool: uses the oracle's CPU binning)
from gsgen_b200.scenes import make_scene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--tiles", type=int, default=400)
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args()
    oracle.build()
    sc = make_scene(args.workload)
    cam, c2w = sc.cams[0], sc.c2ws[0]
    ocam = oracle.Cam(cam.fx, cam.fy, cam.cx, cam.cy, cam.w, cam.h, cam.near_plane, cam.far_plane)
    cfg = oracle.view_cfg(ocam)
    normals, pts = oracle.get_frustum(ocam, c2w)
    mask = oracle.cull_bsphere(sc.mean, sc.svec, normals, pts, 6.0)
    m2, c2, _, dp = oracle.project_gaussians(sc.mean[mask], sc.qvec[mask], sc.svec[mask], c2w, True)
    D, tl, br = oracle.tile_culling_aabb_count(m2, c2, 16, ocam, 6.0)
    th, tw = cfg["n_tiles_h"], cfg["n_tiles_w"]
    ids, start, end = oracle.tile_culling_aabb_start_end(tl, br, dp, th, tw, D)
    alpha = torch.clamp(sc.alpha[mask].reshape(-1), max=0.99).numpy().astype(np.float64)
    mean2d, cov = m2.numpy().astype(np.float64), c2.reshape(-1, 4).numpy().astype(np.float64)
    det = cov[:, 0] * cov[:, 3] - cov[:, 1] * cov[:, 2]
    inv = np.stack([cov[:, 3], -cov[:, 1], -cov[:, 2], cov[:, 0]], 1) / det[:, None]
    # half extents of the a*G >= 1/255 box (make_splat): q <= 2 ln(255 a)
    qmax = 2.0 * np.log(np.maximum(255.0 * alpha, 1e-30))
    hx = np.sqrt(np.maximum(qmax * cov[:, 0], 0.0))
    hy = np.sqrt(np.maximum(qmax * cov[:, 3], 0.0))
    rng = np.random.default_rng(0)
    nonempty = np.nonzero((start.numpy() >= 0))[0]
    sample = rng.choice(nonempty, size=min(args.tiles, len(nonempty)), replace=False)
    B = args.batch
    tot = dict(sync=0.0, decoupled=0.0, balanced=0.0, hits=0, entries=0)
    pairings = {"opposite (i, 7-i)": [(0, 7), (1, 6), (2, 5), (3, 4)], "half-tile (i, i+4)": [(0, 4), (1, 5), (2, 6), (3, 7)],
                "neighbour (2i, 2i+1)": [(0, 1), (2, 3), (4, 5), (6, 7)]}
    pair_stat = {k: dict(busiest=0.0, mean=0.0, both=0, either=0) for k in pairings}
    lx, ly = np.meshgrid(np.arange(16), np.arange(16))
    for tile in sample:
        ty, tx = divmod(int(tile), tw)
        s, e = int(start[tile]), int(end[tile])
        g = ids[s:e].numpy()
        px = (tx * 16 + lx) / cam.fx - cam.cx / cam.fx
        py = (ty * 16 + ly) / cam.fy - cam.cy / cam.fy
        T = np.ones((16, 16))
        done = np.zeros((16, 16), bool)
        work = []  # per batch: hits per warp [8]
        for b0 in range(0, len(g), B):
            wb = np.zeros(8)
            per_entry = []
            for gi in g[b0:b0 + B]:
                dx, dy = px - mean2d[gi, 0], py - mean2d[gi, 1]
                q = inv[gi, 0] * dx * dx + (inv[gi, 1] + inv[gi, 2]) * dx * dy + inv[gi, 3] * dy * dy
                aG = alpha[gi] * np.exp(-0.5 * q)
                ok = (~done) & (aG >= 1.0 / 255.0)
                # which warps (8x4 blocks: warp = 2*(row//4) + col//8) walk this entry
                okw = ok.reshape(4, 4, 2, 8).any(axis=(1, 3)).reshape(-1)  # [row block 4][col block 2] = warp index
                wb += okw
                per_entry.append(okw)
                T = np.where(ok, T * (1.0 - aG), T)
                done |= ok & (T < 1e-4)
            work.append(wb)
            pe = np.array(per_entry)
            for name, pairs in pairings.items():
                for a_, b_ in pairs:
                    pair_stat[name]["both"] += int((pe[:, a_] & pe[:, b_]).sum())
                    pair_stat[name]["either"] += int((pe[:, a_] | pe[:, b_]).sum())
            tot["entries"] += len(g[b0:b0 + B])
            if done.all():
                break
        work = np.array(work)  # [batches, 8]
        tot["sync"] += work.max(axis=1).sum()
        tot["decoupled"] += work.sum(axis=0).max()
        tot["balanced"] += work.sum(axis=0).mean()
        tot["hits"] += work.sum()
        per_warp = work.sum(axis=0)
        for name, pairs in pairings.items():
            ps = np.array([per_warp[a_] + per_warp[b_] for a_, b_ in pairs])
            pair_stat[name]["busiest"] += ps.max()
            pair_stat[name]["mean"] += ps.mean()
    print(f"{args.workload}: {len(sample)} tiles, batch {B}: staged entries/tile {tot['entries'] / len(sample):.0f}, "
          f"warp-hits/tile {tot['hits'] / len(sample):.0f}")
    print(f"  schedule length in warp-hits per tile: sync {tot['sync'] / len(sample):.1f}  decoupled "
          f"{tot['decoupled'] / len(sample):.1f}  balanced {tot['balanced'] / len(sample):.1f}")
    print(f"  sync / decoupled = {tot['sync'] / tot['decoupled']:.3f}   sync / balanced = {tot['sync'] / tot['balanced']:.3f}"
          f"   decoupled / balanced = {tot['decoupled'] / tot['balanced']:.3f}")
    print("  two pixel blocks per warp (4 warps per tile): busiest pair / mean pair, and share of walked entries that "
          "both blocks of a pair blend (one shared-memory read serves two)")
    for name, st in pair_stat.items():
        print(f"    {name:22s} busiest/mean {st['busiest'] / st['mean']:.3f}   both/either {st['both'] / max(st['either'], 1):.3f}")


if __name__ == "__main__":
    main()
