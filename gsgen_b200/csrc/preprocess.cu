// preprocess.cu -- per-Gaussian kernels of the hot path (all HBM-bound streaming kernels):
//   k_cull_bsphere        _gs.culling_gaussian_bsphere                     
//   k_project_fwd / _bwd  gs.renderer.project_gaussians and its autograd   (a4)
//   k_aabb_count          gs.culling.tile_culling_aabb_count               (a6)
//   k_pack_splats         (mean2d,cov2d,alpha[,payload]) -> splat records for the composite kernels
//   k_preprocess          a2+a3+a4+a5+a6 fused, no stream compaction       (fused path, §8(f)-1 in part)
//   k_project_bwd_fused   composite gradient records -> grads of mean/qvec/svec/alpha/color
#include "gsb200_common.cuh"
#include "kernels.cuh"

namespace gsb {

constexpr int kThreads = 256;

// adds the block's sum of per-thread duplicate counts into total[0] (and, when vis >= 0, the number of Gaussians that
// passed the frustum test into total[1]): one 64-bit atomic per block and counter; must be reached by every thread
__device__ __forceinline__ void block_add_total(int v, unsigned long long* total, int vis = -1) {
  __shared__ int s_part[2][kThreads / 32];
  v = __reduce_add_sync(0xffffffffu, v);
  const int nv = (vis >= 0) ? __reduce_add_sync(0xffffffffu, vis) : 0;
  if ((threadIdx.x & 31) == 0) { s_part[0][threadIdx.x >> 5] = v; s_part[1][threadIdx.x >> 5] = nv; }
  __syncthreads();
  if (threadIdx.x == 0 && total) {
    int t = 0, u = 0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) { t += s_part[0][i]; u += s_part[1][i]; }
    if (t) atomicAdd(total, (unsigned long long)t);
    if (vis >= 0 && u) atomicAdd(total + 1, (unsigned long long)u);
  }
}

__global__ void __launch_bounds__(kThreads)
k_cull_bsphere(uint32_t N, const float* __restrict__ mean, const float* __restrict__ svec,
               const float* __restrict__ normal, const float* __restrict__ pts, uint8_t* __restrict__ mask,
               float thresh) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float m[3] = {mean[3 * i], mean[3 * i + 1], mean[3 * i + 2]};
  float r = fmaxf(fmaxf(svec[3 * i], svec[3 * i + 1]), svec[3 * i + 2]) * thresh;
  mask[i] = sphere_in_frustum(m, r, normal, pts) ? 1 : 0;  // plane data: uniform read-only loads (L1 broadcast)
}

__global__ void __launch_bounds__(kThreads)
k_project_fwd(uint32_t N, const float* __restrict__ mean, const float* __restrict__ qvec,
              const float* __restrict__ svec, Camera cam, float* __restrict__ mean2d, float* __restrict__ cov2d,
              float* __restrict__ JW, float* __restrict__ depth) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float x[3] = {mean[3 * i], mean[3 * i + 1], mean[3 * i + 2]};
  float4 q4 = reinterpret_cast<const float4*>(qvec)[i];
  float q[4] = {q4.x, q4.y, q4.z, q4.w};
  float s[3] = {svec[3 * i], svec[3 * i + 1], svec[3 * i + 2]};
  Proj f;
  project_gaussian(x, q, s, cam, f);
  reinterpret_cast<float2*>(mean2d)[i] = make_float2(f.mean2d[0], f.mean2d[1]);
  reinterpret_cast<float4*>(cov2d)[i] = make_float4(f.cov[0], f.cov[1], f.cov[2], f.cov[3]);
  depth[i] = f.depth;
  if (JW) {  // third row of J (gs/renderer.py:371-376): p / |p|
    float l = sqrtf(f.p[0] * f.p[0] + f.p[1] * f.p[1] + f.p[2] * f.p[2]);
    float j2[3] = {f.p[0] / l, f.p[1] / l, f.p[2] / l};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      JW[9 * i + k] = f.T2[k];
      JW[9 * i + 3 + k] = f.T2[3 + k];
      JW[9 * i + 6 + k] = j2[0] * cam.R[3 * k] + j2[1] * cam.R[3 * k + 1] + j2[2] * cam.R[3 * k + 2];
    }
  }
}

__global__ void __launch_bounds__(kThreads)
k_project_bwd(uint32_t N, const float* __restrict__ mean, const float* __restrict__ qvec,
              const float* __restrict__ svec, Camera cam, const float* __restrict__ g_m2,
              const float* __restrict__ g_cov, const float* __restrict__ g_depth, float* __restrict__ g_mean,
              float* __restrict__ g_qvec, float* __restrict__ g_svec) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float x[3] = {mean[3 * i], mean[3 * i + 1], mean[3 * i + 2]};
  float4 q4 = reinterpret_cast<const float4*>(qvec)[i];
  float q[4] = {q4.x, q4.y, q4.z, q4.w};
  float s[3] = {svec[3 * i], svec[3 * i + 1], svec[3 * i + 2]};
  Proj f;
  project_gaussian(x, q, s, cam, f);
  float gm2[2] = {0.f, 0.f}, gc[4] = {0.f, 0.f, 0.f, 0.f}, gd = 0.f;
  if (g_m2) { float2 t = reinterpret_cast<const float2*>(g_m2)[i]; gm2[0] = t.x; gm2[1] = t.y; }
  if (g_cov) { float4 t = reinterpret_cast<const float4*>(g_cov)[i]; gc[0] = t.x; gc[1] = t.y; gc[2] = t.z; gc[3] = t.w; }
  if (g_depth) gd = g_depth[i];
  float gx[3], gq[4], gs[3];
  project_gaussian_bwd(s, cam, f, gm2, gc, gd, gx, gq, gs);
#pragma unroll
  for (int k = 0; k < 3; ++k) { g_mean[3 * i + k] = gx[k]; g_svec[3 * i + k] = gs[k]; }
  reinterpret_cast<float4*>(g_qvec)[i] = make_float4(gq[0], gq[1], gq[2], gq[3]);
}

__global__ void __launch_bounds__(kThreads)
k_aabb_count(uint32_t N, const float* __restrict__ mean2d, const float* __restrict__ cov2d, int tile, float fx,
             float fy, float cx, float cy, int W, int H, float D, int32_t* __restrict__ tl,
             int32_t* __restrict__ br, unsigned long long* __restrict__ total) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  int cnt = 0;
  if (i < N) {
    float2 m = reinterpret_cast<const float2*>(mean2d)[i];
    float4 c = reinterpret_cast<const float4*>(cov2d)[i];
    float m2[2] = {m.x, m.y};
    int r[4];
    aabb_tiles(m2, c.x, c.w, D, fx, fy, cx, cy, W, H, tile, r);
    reinterpret_cast<int2*>(tl)[i] = make_int2(r[0], r[1]);
    reinterpret_cast<int2*>(br)[i] = make_int2(r[2], r[3]);
    cnt = (r[2] - r[0] + 1) * (r[3] - r[1] + 1);
  }
  block_add_total(cnt, total);
}

// counts from caller-provided AABBs (reference-compatible binning op)
__global__ void __launch_bounds__(kThreads)
k_count_from_aabb(uint32_t N, const int32_t* __restrict__ tl, const int32_t* __restrict__ br,
                  int32_t* __restrict__ count, ushort4* __restrict__ rect, unsigned long long* __restrict__ total) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  int cnt = 0;
  if (i < N) {
    int2 a = reinterpret_cast<const int2*>(tl)[i], b = reinterpret_cast<const int2*>(br)[i];
    int w = b.x - a.x + 1, h = b.y - a.y + 1;
    cnt = (w > 0 && h > 0) ? w * h : 0;
    count[i] = cnt;
    rect[i] = make_ushort4((unsigned short)a.x, (unsigned short)a.y, (unsigned short)b.x, (unsigned short)b.y);
  }
  block_add_total(cnt, total);
}

// payload kinds: 0 none (SH path: payload fetched from the caller's sh tensor), 1 RGB [N,3], 2 scalar [>=N]
__global__ void __launch_bounds__(kThreads)
k_pack_splats(uint32_t N, const float* __restrict__ mean2d, const float* __restrict__ cov2d,
              const float* __restrict__ alpha, const float* __restrict__ payload, int pay_kind,
              Splat* __restrict__ splat, float4* __restrict__ pay) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float2 m = reinterpret_cast<const float2*>(mean2d)[i];
  float4 c = reinterpret_cast<const float4*>(cov2d)[i];
  float m2[2] = {m.x, m.y}, cov[4] = {c.x, c.y, c.z, c.w};
  Splat s = make_splat(m2, cov, alpha[i]);
  float4* sp = reinterpret_cast<float4*>(splat + i);
  sp[0] = make_float4(s.mx, s.my, s.p0, s.p1);
  sp[1] = make_float4(s.p2, s.a, s.hx, s.hy);
  if (pay_kind == 1) pay[i] = make_float4(payload[3 * i], payload[3 * i + 1], payload[3 * i + 2], 0.f);
  else if (pay_kind == 2) pay[i] = make_float4(payload[i], 0.f, 0.f, 0.f);
}

// ---------------------------------------------------------------------------------------------------
// fused front end: cull + project + radius + AABB + count + splat record.  84 B/Gaussian algorithmic.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int preprocess_one(
    uint32_t i, const float* __restrict__ mean, const float* __restrict__ qvec, const float* __restrict__ svec,
    const float* __restrict__ alpha, const float* __restrict__ color, int act, const Camera& cam,
    float* __restrict__ mean2d, float* __restrict__ cov2d, float* __restrict__ depthg, uint8_t* __restrict__ mask,
    float* __restrict__ radii2d, Splat* __restrict__ splat, float4* __restrict__ pay, ushort4* __restrict__ rect,
    int32_t* __restrict__ count, uint32_t* __restrict__ dkeys, int32_t* __restrict__ perm0, bool* vis) {
  float x[3] = {mean[3 * i], mean[3 * i + 1], mean[3 * i + 2]};
  perm0[i] = (int32_t)i;  // input of the per-Gaussian depth sort (binning.cu step 1), written here: one launch less
  // raw leaves -> activated values in registers; act == 0: the tensors are already activated
  float s[3] = {act_svec(svec[3 * i], act), act_svec(svec[3 * i + 1], act), act_svec(svec[3 * i + 2], act)};
  bool keep = true;
  if (!cam.skip_frustum) {
    float r = fmaxf(fmaxf(s[0], s[1]), s[2]) * cam.frustum_radius;
    keep = sphere_in_frustum(x, r, cam.fn, cam.fp);
  }
  mask[i] = keep ? 1 : 0;
  *vis = keep;
  if (!keep) {
    reinterpret_cast<float2*>(mean2d)[i] = make_float2(0.f, 0.f);
    reinterpret_cast<float4*>(cov2d)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    depthg[i] = 0.f;
    dkeys[i] = 0u;
    if (radii2d) radii2d[i] = 0.f;
    count[i] = 0;
    return 0;
  }
  float4 q4 = reinterpret_cast<const float4*>(qvec)[i];
  float q[4] = {q4.x, q4.y, q4.z, q4.w};
  Proj f;
  project_gaussian(x, q, s, cam, f);
  reinterpret_cast<float2*>(mean2d)[i] = make_float2(f.mean2d[0], f.mean2d[1]);
  reinterpret_cast<float4*>(cov2d)[i] = make_float4(f.cov[0], f.cov[1], f.cov[2], f.cov[3]);
  depthg[i] = f.depth;
  dkeys[i] = __float_as_uint(f.depth);
  if (radii2d) radii2d[i] = radius2d(f.cov);
  Splat sp = make_splat(f.mean2d, f.cov, act_alpha(alpha[i], act));
  int r[4];
  aabb_tiles(f.mean2d, f.cov[0], f.cov[3], cam.tile_radius, cam.fx, cam.fy, cam.cx, cam.cy, cam.W, cam.H, 16, r);
  int cnt = (r[2] - r[0] + 1) * (r[3] - r[1] + 1);
  // a covariance that is not positive definite has undefined AABBs in the reference (NaN sqrt -> int); it owns
  // no duplicates here
  if (!(f.cov[0] > 0.f) || !(f.cov[3] > 0.f) || cnt < 0) cnt = 0;
  count[i] = cnt;
  rect[i] = make_ushort4((unsigned short)r[0], (unsigned short)r[1], (unsigned short)r[2], (unsigned short)r[3]);
  float4* spp = reinterpret_cast<float4*>(splat + i);
  spp[0] = make_float4(sp.mx, sp.my, sp.p0, sp.p1);
  spp[1] = make_float4(sp.p2, sp.a, sp.hx, sp.hy);
  if (color)
    pay[i] = make_float4(act_color(color[3 * i], act), act_color(color[3 * i + 1], act),
                         act_color(color[3 * i + 2], act), f.depth);
  return cnt;
}

__global__ void __launch_bounds__(kThreads)
k_preprocess(uint32_t N, const float* __restrict__ mean, const float* __restrict__ qvec,
             const float* __restrict__ svec, const float* __restrict__ alpha, const float* __restrict__ color,
             int act, Camera cam, float* __restrict__ mean2d, float* __restrict__ cov2d, float* __restrict__ depthg,
             uint8_t* __restrict__ mask, float* __restrict__ radii2d, Splat* __restrict__ splat,
             float4* __restrict__ pay, ushort4* __restrict__ rect, int32_t* __restrict__ count,
             uint32_t* __restrict__ dkeys, int32_t* __restrict__ perm0, unsigned long long* __restrict__ total) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool vis = false;
  const int cnt = i < N ? preprocess_one(i, mean, qvec, svec, alpha, color, act, cam, mean2d, cov2d, depthg, mask,
                                         radii2d, splat, pay, rect, count, dkeys, perm0, &vis)
                        : 0;
  block_add_total(cnt, total, vis ? 1 : 0);
}


// ---------------------------------------------------------------------------------------------------
// fused back end: gradient records of the composite backward -> parameter gradients (all written).
//   ggeom[i] = {gmx, gmy, gxx, gxy | gyy, galpha, gdepth, -}   gpay[i] = {gr, gg, gb, -}
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 4)
k_project_bwd_fused(uint32_t N, const float* __restrict__ mean, const float* __restrict__ qvec,
                    const float* __restrict__ svec, const float* __restrict__ alpha,
                    const float* __restrict__ color, int act, const uint8_t* __restrict__ mask, Camera cam,
                    const float4* __restrict__ ggeom, const float4* __restrict__ gpay, float* __restrict__ g_mean,
                    float* __restrict__ g_qvec, float* __restrict__ g_svec, float* __restrict__ g_alpha,
                    float* __restrict__ g_color, float* __restrict__ g_mean2d, int accumulate,
                    uint8_t* __restrict__ touched) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float gx[3] = {0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f}, gs[3] = {0.f, 0.f, 0.f};
  float ga = 0.f, gcol[3] = {0.f, 0.f, 0.f}, gm2[2] = {0.f, 0.f};
  const bool vis = mask[i] != 0;
  // accumulate mode: the running gradients are fetched together with the inputs (ONE round of HBM latency instead of
  // two: measured round 2, this kernel was latency bound -- 38 % issue slots, long-scoreboard stall 12.4 per issue)
  float om[3] = {0.f, 0.f, 0.f}, os[3] = {0.f, 0.f, 0.f}, oa = 0.f, oc[3] = {0.f, 0.f, 0.f};
  float4 oq = make_float4(0.f, 0.f, 0.f, 0.f);
  if (accumulate && vis) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { om[k] = g_mean[3 * i + k]; os[k] = g_svec[3 * i + k]; }
    oq = reinterpret_cast<const float4*>(g_qvec)[i];
    oa = g_alpha[i];
    if (g_color) {
#pragma unroll
      for (int k = 0; k < 3; ++k) oc[k] = g_color[3 * i + k];
    }
  }
  if (vis) {
    float4 g0 = ggeom[2 * i], g1 = ggeom[2 * i + 1];
    if (touched && g1.w != 0.f) touched[i] = 1;  // the composite backward flushed into this Gaussian (sparse all-reduce)
    float x[3] = {mean[3 * i], mean[3 * i + 1], mean[3 * i + 2]};
    float4 q4 = reinterpret_cast<const float4*>(qvec)[i];
    float q[4] = {q4.x, q4.y, q4.z, q4.w};
    float s[3] = {act_svec(svec[3 * i], act), act_svec(svec[3 * i + 1], act), act_svec(svec[3 * i + 2], act)};
    Proj f;
    project_gaussian(x, q, s, cam, f);
    gm2[0] = g0.x; gm2[1] = g0.y;
    float gc[4] = {g0.z, g0.w, g0.w, g1.x};
    project_gaussian_bwd(s, cam, f, gm2, gc, g1.z, gx, gq, gs);
    ga = g1.y;
    if (gpay) { float4 p = gpay[i]; gcol[0] = p.x; gcol[1] = p.y; gcol[2] = p.z; }
    if (act) {  // gradients w.r.t. the raw leaves (same expressions as torch's exp / sigmoid backward)
#pragma unroll
      for (int k = 0; k < 3; ++k) gs[k] = act_svec_bwd(gs[k], s[k], act);
      if (act & kActAlphaSigmoid) ga = act_alpha_bwd(ga, act_sigmoid(alpha[i]), act);
      if (gpay && (act & kActColorSigmoid)) {
#pragma unroll
        for (int k = 0; k < 3; ++k) gcol[k] = act_color_bwd(gcol[k], act_sigmoid(color[3 * i + k]), act);
      }
    }
  }
  if (g_mean2d) reinterpret_cast<float2*>(g_mean2d)[i] = make_float2(gm2[0], gm2[1]);
  if (accumulate) {  // += into the caller's running gradient (e.g. the flat all-reduce buffer); culled: no traffic
    if (!vis) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) { g_mean[3 * i + k] = om[k] + gx[k]; g_svec[3 * i + k] = os[k] + gs[k]; }
    reinterpret_cast<float4*>(g_qvec)[i] = make_float4(oq.x + gq[0], oq.y + gq[1], oq.z + gq[2], oq.w + gq[3]);
    g_alpha[i] = oa + ga;
    if (g_color) {
#pragma unroll
      for (int k = 0; k < 3; ++k) g_color[3 * i + k] = oc[k] + gcol[k];
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) { g_mean[3 * i + k] = gx[k]; g_svec[3 * i + k] = gs[k]; }
  reinterpret_cast<float4*>(g_qvec)[i] = make_float4(gq[0], gq[1], gq[2], gq[3]);
  g_alpha[i] = ga;
  if (g_color) {
#pragma unroll
    for (int k = 0; k < 3; ++k) g_color[3 * i + k] = gcol[k];
  }
}

// ---- host launchers ----------------------------------------------------------------------------------
static inline dim3 grid1d(uint32_t n) { return dim3((n + kThreads - 1) / kThreads); }

int launch_cull_bsphere(uint32_t N, const float* mean, const float* svec, const float* normal, const float* pts,
                        uint8_t* mask, float thresh, cudaStream_t st) {
  if (N == 0) return GSB200_OK;
  k_cull_bsphere<<<grid1d(N), kThreads, 0, st>>>(N, mean, svec, normal, pts, mask, thresh);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}
int launch_project_fwd(uint32_t N, const float* mean, const float* qvec, const float* svec, const Camera& cam,
                       float* mean2d, float* cov2d, float* JW, float* depth, cudaStream_t st) {
  if (N == 0) return GSB200_OK;
  k_project_fwd<<<grid1d(N), kThreads, 0, st>>>(N, mean, qvec, svec, cam, mean2d, cov2d, JW, depth);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}
int launch_project_bwd(uint32_t N, const float* mean, const float* qvec, const float* svec, const Camera& cam,
                       const float* g_m2, const float* g_cov, const float* g_depth, float* g_mean, float* g_qvec,
                       float* g_svec, cudaStream_t st) {
  if (N == 0) return GSB200_OK;
  k_project_bwd<<<grid1d(N), kThreads, 0, st>>>(N, mean, qvec, svec, cam, g_m2, g_cov, g_depth, g_mean, g_qvec,
                                               g_svec);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}
int launch_aabb_count(uint32_t N, const float* mean2d, const float* cov2d, int tile, float fx, float fy, float cx,
                      float cy, int W, int H, float D, int32_t* tl, int32_t* br, unsigned long long* total,
                      cudaStream_t st) {
  if (N == 0) return GSB200_OK;
  k_aabb_count<<<grid1d(N), kThreads, 0, st>>>(N, mean2d, cov2d, tile, fx, fy, cx, cy, W, H, D, tl, br, total);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}
int launch_count_from_aabb(uint32_t N, const int32_t* tl, const int32_t* br, int32_t* count, ushort4* rect,
                           unsigned long long* total, cudaStream_t st) {
  if (N == 0) return GSB200_OK;
  k_count_from_aabb<<<grid1d(N), kThreads, 0, st>>>(N, tl, br, count, rect, total);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}
int launch_pack_splats(uint32_t N, const float* mean2d, const float* cov2d, const float* alpha, const float* payload,
                       int pay_kind, Splat* splat, float4* pay, cudaStream_t st) {
  if (N == 0) return GSB200_OK;
  k_pack_splats<<<grid1d(N), kThreads, 0, st>>>(N, mean2d, cov2d, alpha, payload, pay_kind, splat, pay);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}
int launch_preprocess(uint32_t N, const float* mean, const float* qvec, const float* svec, const float* alpha,
                      const float* color, int act, const Camera& cam, float* mean2d, float* cov2d, float* depthg,
                      uint8_t* mask, float* radii2d, Splat* splat, float4* pay, ushort4* rect, int32_t* count,
                      uint32_t* dkeys, int32_t* perm0, unsigned long long* total, cudaStream_t st) {
  if (N == 0) return GSB200_OK;
  k_preprocess<<<grid1d(N), kThreads, 0, st>>>(N, mean, qvec, svec, alpha, color, act, cam, mean2d, cov2d, depthg, mask,
                                              radii2d, splat, pay, rect, count, dkeys, perm0, total);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}
int launch_project_bwd_fused(uint32_t N, const float* mean, const float* qvec, const float* svec,
                             const float* alpha, const float* color, int act, const uint8_t* mask, const Camera& cam, const float4* ggeom, const float4* gpay,
                             float* g_mean, float* g_qvec, float* g_svec, float* g_alpha, float* g_color,
                             float* g_mean2d, int accumulate, uint8_t* touched, cudaStream_t st) {
  if (N == 0) return GSB200_OK;
  k_project_bwd_fused<<<grid1d(N), kThreads, 0, st>>>(N, mean, qvec, svec, alpha, color, act, mask, cam, ggeom, gpay, g_mean, g_qvec,
                                                     g_svec, g_alpha, g_color, g_mean2d, accumulate, touched);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}

}  // namespace gsb
