// hostmath.cpp -- TEST INFRASTRUCTURE.  Compiles gsgen_b200/csrc/gsb200_math.cuh (the host+device
// per-Gaussian math the CUDA kernels use) with g++ so the CPU test-suite can check it against the
// oracle without a GPU.  Never loaded by the product path.
#include "../../gsgen_b200/csrc/gsb200_math.cuh"
#include "../../gsgen_b200/csrc/knn_grid.cuh"

#include <algorithm>
#include <vector>

using namespace gsb;

static Camera make_cam(const float* c2w12, int depth_detach) {
  Camera c{};
  for (int r = 0; r < 3; ++r) {
    for (int k = 0; k < 3; ++k) c.R[3 * r + k] = c2w12[4 * r + k];
    c.t[r] = c2w12[4 * r + 3];
  }
  c.depth_detach = depth_detach;
  return c;
}

extern "C" {

void hm_project_fwd(int N, const float* mean, const float* qvec, const float* svec, const float* c2w12,
                    float* mean2d, float* cov2d, float* depth) {
  Camera cam = make_cam(c2w12, 1);
  for (int i = 0; i < N; ++i) {
    Proj f;
    project_gaussian(mean + 3 * i, qvec + 4 * i, svec + 3 * i, cam, f);
    mean2d[2 * i] = f.mean2d[0]; mean2d[2 * i + 1] = f.mean2d[1];
    for (int k = 0; k < 4; ++k) cov2d[4 * i + k] = f.cov[k];
    depth[i] = f.depth;
  }
}

void hm_project_bwd(int N, const float* mean, const float* qvec, const float* svec, const float* c2w12,
                    int depth_detach, const float* g_m2, const float* g_cov, const float* g_depth, float* g_mean,
                    float* g_qvec, float* g_svec) {
  Camera cam = make_cam(c2w12, depth_detach);
  for (int i = 0; i < N; ++i) {
    Proj f;
    project_gaussian(mean + 3 * i, qvec + 4 * i, svec + 3 * i, cam, f);
    project_gaussian_bwd(svec + 3 * i, cam, f, g_m2 + 2 * i, g_cov + 4 * i, g_depth[i], g_mean + 3 * i,
                         g_qvec + 4 * i, g_svec + 3 * i);
  }
}

void hm_aabb(int N, const float* mean2d, const float* cov2d, float D, float fx, float fy, float cx, float cy, int W,
             int H, int tile, int* tl, int* br) {
  for (int i = 0; i < N; ++i) {
    int r[4];
    aabb_tiles(mean2d + 2 * i, cov2d[4 * i], cov2d[4 * i + 3], D, fx, fy, cx, cy, W, H, tile, r);
    tl[2 * i] = r[0]; tl[2 * i + 1] = r[1]; br[2 * i] = r[2]; br[2 * i + 1] = r[3];
  }
}

int hm_cull(int N, const float* mean, const float* svec, const float* fn, const float* fp, float thresh,
            unsigned char* mask) {
  int n = 0;
  for (int i = 0; i < N; ++i) {
    float r = fmaxf(fmaxf(svec[3 * i], svec[3 * i + 1]), svec[3 * i + 2]) * thresh;
    mask[i] = sphere_in_frustum(mean + 3 * i, r, fn, fp) ? 1 : 0;
    n += mask[i];
  }
  return n;
}

// a*G of splat records at query points (exp2 evaluated with exp2f here; the kernels use MUFU.EX2), plus the
// bounding-box half extents and the gradient direction v = S^-1 d used by the backward
void hm_splat_eval(int N, const float* mean2d, const float* cov2d, const float* alpha, const float* query /*[N,2]*/,
                   float* aG, float* G, float* vx, float* vy, float* hx, float* hy) {
  for (int i = 0; i < N; ++i) {
    Splat s = make_splat(mean2d + 2 * i, cov2d + 4 * i, alpha[i]);
    float dx = query[2 * i] - s.mx, dy = query[2 * i + 1] - s.my;
    float u = fmaf(s.p0, dx, s.p1 * dy), v = s.p2 * dy;
    float g = exp2f(fmaf(-u, u, -(v * v)));
    G[i] = g; aG[i] = s.a * g;
    vx[i] = kInvCholScale2 * s.p0 * u;
    vy[i] = kInvCholScale2 * fmaf(s.p1, u, s.p2 * v);
    hx[i] = s.hx; hy[i] = s.hy;
  }
}

void hm_sh_basis(int C, const float* pos2, const float* c9, float* out16) {
  float d[3];
  pixel_dir(pos2[0], pos2[1], c9, d);
  switch (C) {
    case 1: sh_basis<1>(d[0], d[1], d[2], out16); break;
    case 2: sh_basis<2>(d[0], d[1], d[2], out16); break;
    case 3: sh_basis<3>(d[0], d[1], d[2], out16); break;
    default: sh_basis<4>(d[0], d[1], d[2], out16); break;
  }
}

// Per-Gaussian slice of the fused front end with in-kernel activations, as k_preprocess /
// k_project_bwd_fused evaluate it: raw leaves -> activated values -> projection, and back to raw-leaf gradients.
void hm_front_end_raw(int N, int act, const float* mean, const float* qvec, const float* svec_raw,
                      const float* alpha_raw, const float* color_raw, const float* c2w12, float* mean2d, float* cov2d,
                      float* alpha_act, float* color_act, const float* g_m2, const float* g_cov,
                      const float* g_alpha_act, const float* g_color_act, float* g_mean, float* g_qvec,
                      float* g_svec_raw, float* g_alpha_raw, float* g_color_raw) {
  Camera cam = make_cam(c2w12, 1);
  for (int i = 0; i < N; ++i) {
    float s[3];
    for (int k = 0; k < 3; ++k) s[k] = act_svec(svec_raw[3 * i + k], act);
    Proj f;
    project_gaussian(mean + 3 * i, qvec + 4 * i, s, cam, f);
    mean2d[2 * i] = f.mean2d[0]; mean2d[2 * i + 1] = f.mean2d[1];
    for (int k = 0; k < 4; ++k) cov2d[4 * i + k] = f.cov[k];
    alpha_act[i] = act_alpha(alpha_raw[i], act);
    for (int k = 0; k < 3; ++k) color_act[3 * i + k] = act_color(color_raw[3 * i + k], act);
    float gs[3];
    project_gaussian_bwd(s, cam, f, g_m2 + 2 * i, g_cov + 4 * i, 0.f, g_mean + 3 * i, g_qvec + 4 * i, gs);
    for (int k = 0; k < 3; ++k) g_svec_raw[3 * i + k] = act_svec_bwd(gs[k], s[k], act);
    g_alpha_raw[i] = act_alpha_bwd(g_alpha_act[i], alpha_act[i], act);
    for (int k = 0; k < 3; ++k)
      g_color_raw[3 * i + k] = act_color_bwd(g_color_act[3 * i + k], color_act[3 * i + k], act);
  }
}

// One Adam update of n elements with the expression order of k_adam_flat; the scalars are derived
// exactly as gsb200_adam_step derives them.
void hm_adam(long long n, float* p, const float* g, float* m, float* v, double lr, double beta1, double beta2,
             double eps, long long step, float grad_scale) {
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  AdamScalars K;
  K.beta2 = (float)beta2;
  K.one_minus_beta1 = (float)(1.0 - beta1);
  K.one_minus_beta2 = (float)(1.0 - beta2);
  K.eps = (float)eps;
  K.bc2_sqrt = (float)sqrt(bc2);
  K.grad_scale = grad_scale;
  const float step_size = (float)(lr / bc1);
  for (long long i = 0; i < n; ++i) adam_update(p[i], g[i], m[i], v[i], step_size, K);
}

// largest eigenvalue of cov2d as k_preprocess computes it for aux["radii2d"] (gs/gaussian_splatting.py:1240-1245)
void hm_radius2d(int N, const float* cov2d, float* out) {
  for (int i = 0; i < N; ++i) out[i] = radius2d(cov2d + 4 * i);
}

// Exact KNN with the SAME grid construction and shell search the device path uses (gsgen_b200/csrc/knn_grid.cuh):
// bounding box -> knn_make_grid -> cell ids -> stable sort by cell -> cell_start -> knn_query per query.
// queries == nullptr: the points query themselves (a point is its own first neighbour).  stats[0] = cells,
// stats[1] = largest number of shells any query visited.
int hm_knn(int n, const float* pts, int nq, const float* queries, int K, unsigned max_cells, long long* idx,
           float* d2, int* stats) {
  if (K < 1 || K > kKnnMaxK) return 1;
  if (!queries) { queries = pts; nq = n; }
  float bmin[3] = {INFINITY, INFINITY, INFINITY}, bmax[3] = {-INFINITY, -INFINITY, -INFINITY}, lo[3], hi[3];
  double sum[3] = {0, 0, 0}, sumsq[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a) {
      const float v = pts[3 * i + a];
      bmin[a] = fminf(bmin[a], v); bmax[a] = fmaxf(bmax[a], v);
      sum[a] += (double)v; sumsq[a] += (double)v * (double)v;
    }
  knn_robust_box(bmin, bmax, sum, sumsq, (uint32_t)n, lo, hi);
  const KnnGrid G = knn_make_grid(lo, hi, (uint32_t)n, max_cells);
  std::vector<uint32_t> key(n);
  std::vector<int> perm(n);
  for (int i = 0; i < n; ++i) { key[i] = knn_cell_id(G, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]); perm[i] = i; }
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return key[a] < key[b]; });
  std::vector<KnnPt> sp(n > 0 ? n : 1);
  std::vector<uint32_t> skey(n);
  for (int s = 0; s < n; ++s) {
    const int i = perm[s];
    sp[s] = KnnPt{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], i};
    skey[s] = key[i];
  }
  std::vector<int32_t> cell_start(G.cells + 1);
  for (int c = 0; c <= G.cells; ++c)
    cell_start[c] = (int32_t)(std::lower_bound(skey.begin(), skey.end(), (uint32_t)c) - skey.begin());
  int shells = 0;
  auto run = [&](auto tag) {
    constexpr int KT = decltype(tag)::value;
    for (int j = 0; j < nq; ++j) {
      float bd[KT]; int32_t bi[KT];
      const int sh = knn_query<KT>(G, sp.data(), cell_start.data(), queries[3 * j], queries[3 * j + 1],
                                   queries[3 * j + 2], bd, bi);
      shells = sh > shells ? sh : shells;
      for (int k = 0; k < K; ++k) { idx[(long long)j * K + k] = bi[k]; d2[(long long)j * K + k] = bd[k]; }
    }
  };
  if (K <= 2) run(std::integral_constant<int, 2>());
  else if (K <= 4) run(std::integral_constant<int, 4>());
  else if (K <= 8) run(std::integral_constant<int, 8>());
  else if (K <= 16) run(std::integral_constant<int, 16>());
  else run(std::integral_constant<int, 32>());
  if (stats) { stats[0] = G.cells; stats[1] = shells; }
  return 0;
}
}
