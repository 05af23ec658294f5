/*
 * oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's tile-composite hot path (gsgen3d/gsgen `gs/src`).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product path (gsgen_b200/) never links or calls it.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 * Arithmetic types follow the reference: the RGB / scalar kernels evaluate the 2-D Gaussian in
 * fp64 (gs/src/include/kernels.h:195-224, :394-418), the SH kernels in fp32 (kernels.h:172-193).
 * Per-pixel loops are sequential over the tile list exactly as the reference's per-thread loop
 * (the reference's shared-memory batching does not change per-pixel arithmetic order).
 *
 * Parity pin: tests/golden/ holds outputs of the UNMODIFIED reference `_gs` extension
 * (built by oracle/build_ref.sh, run on a B200 by tests/golden/make_golden.py); the
 * `-m "not gpu"` suite checks this file against them.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC oracle.c -o liboracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MIN_RENDER_ALPHA 0.00392156862745098f /* common.h:89 */
#define ALPHA_CLAMP 0.99f                     /* vol_render.h:212 */
#define TILE 16
/* fp64 gradient sinks shared by OpenMP threads (order-independent to ~1e-16) */
#define ATOMIC_ADD(dst, v)                                                                        \
  do {                                                                                            \
    double v__ = (v);                                                                             \
    _Pragma("omp atomic") dst += v__;                                                             \
  } while (0)

/* ------------------------------------------------------------------------------------------
 * A.2 frustum cull: culling.h:11-20 + kernels.h:156-170
 *   r = max(svec)*thresh ; keep iff for all 6 planes dot(mean - pts_p, n_p) > -r
 * ---------------------------------------------------------------------------------------- */
void orc_cull_bsphere(int N, const float *mean, const float *svec, const float *normal,
                      const float *pts, uint8_t *mask, float thresh) {
  for (int i = 0; i < N; ++i) {
    float r = fmaxf(fmaxf(svec[3 * i], svec[3 * i + 1]), svec[3 * i + 2]) * thresh;
    int keep = 1;
    for (int p = 0; p < 6; ++p) {
      float dx = mean[3 * i] - pts[3 * p], dy = mean[3 * i + 1] - pts[3 * p + 1],
            dz = mean[3 * i + 2] - pts[3 * p + 2];
      /* helper_math.h dot(): a.x*b.x + a.y*b.y + a.z*b.z */
      float d = dx * normal[3 * p] + dy * normal[3 * p + 1] + dz * normal[3 * p + 2];
      if (!(d > -r)) { keep = 0; break; }
    }
    mask[i] = (uint8_t)keep;
  }
}

/* ------------------------------------------------------------------------------------------
 * A.5 keys / sort / ranges: aabb_culling.h:15-41 (fill_tiledepth_aabb), :235-241 (cub sort,
 * signed int64 ascending), :70-103 (fill_start/end).  Tie order among equal keys is
 * nondeterministic in the reference (atomic slot); the oracle breaks ties by Gaussian index.
 * ---------------------------------------------------------------------------------------- */
typedef struct { int64_t key; int32_t id; } kv_t;
static int kv_cmp(const void *a, const void *b) {
  const kv_t *x = (const kv_t *)a, *y = (const kv_t *)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return (x->id > y->id) - (x->id < y->id);
}
/* returns number of duplicates written (must equal D) or -1 on overflow */
int64_t orc_bin(int N, const int32_t *aabb_tl, const int32_t *aabb_br, const float *depth,
                int n_tiles_h, int n_tiles_w, int64_t D, int32_t *ids_out, int32_t *start,
                int32_t *end, int64_t *keys_out /* nullable */) {
  kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)(D > 0 ? D : 1));
  int64_t n = 0;
  for (int g = 0; g < N; ++g) {
    uint32_t dbits;
    memcpy(&dbits, &depth[g], 4);
    for (int i = aabb_tl[2 * g]; i <= aabb_br[2 * g]; ++i)
      for (int j = aabb_tl[2 * g + 1]; j <= aabb_br[2 * g + 1]; ++j) {
        if (n >= D) { free(kv); return -1; }
        int tile = j * n_tiles_w + i; /* xy2tile_id aabb_culling.h:11-13 */
        kv[n].key = (int64_t)(((uint64_t)(uint32_t)tile << 32) | (uint64_t)dbits);
        kv[n].id = g;
        ++n;
      }
  }
  qsort(kv, (size_t)n, sizeof(kv_t), kv_cmp);
  int T = n_tiles_h * n_tiles_w;
  for (int t = 0; t < T; ++t) { start[t] = -1; end[t] = -1; }
  for (int64_t s = 0; s < n; ++s) {
    int tile = (int)(kv[s].key >> 32);
    ids_out[s] = kv[s].id;
    if (keys_out) keys_out[s] = kv[s].key;
    if (s == 0 || (int)(kv[s - 1].key >> 32) != tile) start[tile] = (int32_t)s;
    if (s == n - 1 || (int)(kv[s + 1].key >> 32) != tile) end[tile] = (int32_t)(s + 1);
  }
  free(kv);
  return n;
}

/* ------------------------------------------------------------------------------------------
 * 2-D Gaussian evaluation.
 *   fp64: kernels.h:195-224 (kernel_gaussian_2d)       used by RGB / scalar kernels
 *   fp32: kernels.h:172-193 (kernel_gaussian_2d_float) used by SH kernels
 * ---------------------------------------------------------------------------------------- */
static inline float gauss2d_f64(const float *mean, const float *cov, const float *q) {
  double c0 = cov[0], c1 = cov[1], c2 = cov[2], c3 = cov[3];
  double det = c0 * c3 - c1 * c2;
  double x = q[0] - mean[0]; /* fp32 subtract, then widened (reference: double x = query[0]-mean[0]) */
  double y = q[1] - mean[1];
  double tx = x * c3 - y * c2;
  double ty = -x * c1 + y * c0;
  double radial = tx * x + ty * y;
  radial /= det;
  if (radial < 0.0) radial = 1000.0;
  return (float)exp(-0.5 * radial);
}
static inline float gauss2d_f32(const float *mean, const float *cov, const float *q) {
  float c0 = cov[0], c1 = cov[1], c2 = cov[2], c3 = cov[3];
  float det = c0 * c3 - c1 * c2;
  float x = q[0] - mean[0];
  float y = q[1] - mean[1];
  float tx = x * c3 - y * c2;
  float ty = -x * c1 + y * c0;
  float radial = tx * x + ty * y;
  radial /= det;
  if (radial < 0.0f) radial = 1000.0f;
  /* reference: (float)expf(-0.5 * radial): -0.5 is a double literal -> product in double */
  return (float)expf((float)(-0.5 * (double)radial));
}
/* kernels.h:394-418 kernel_gaussian_2d_backward (fp64 inverse).  The reference adds into shared memory with
 * float atomics and then into global memory; here each tile accumulates its list in thread-private fp64 sinks
 * (loc_* arrays indexed by list position) that are flushed once per (tile, Gaussian) instance. */
static inline void gauss2d_bwd_f64(const float *mean, const float *cov, const float *q, float grad,
                                   double *gmean, double *gcov) {
  double d_grad = (double)grad;
  double c0 = cov[0], c1 = cov[1], c2 = cov[2], c3 = cov[3];
  double det = c0 * c3 - c1 * c2;
  double x = q[0] - mean[0];
  double y = q[1] - mean[1];
  double tx = (x * c3 - y * c2) / det;
  double ty = (-x * c1 + y * c0) / det;
  gmean[0] += (double)(float)(d_grad * tx);
  gmean[1] += (double)(float)(d_grad * ty);
  gcov[0] += (double)(float)(0.5 * (float)(d_grad * tx * tx));
  gcov[1] += (double)(float)(0.5 * (float)(d_grad * tx * ty));
  gcov[2] += (double)(float)(0.5 * (float)(d_grad * ty * tx));
  gcov[3] += (double)(float)(0.5 * (float)(d_grad * ty * ty));
}

/* per-view statistics the measurement plan needs */
typedef struct {
  int64_t pairs_evaluated; /* (pixel,gaussian) Gaussian evaluations                          */
  int64_t pairs_blended;   /* evaluations that passed the 1/255 test                          */
  int64_t d_eff;           /* sum over tiles of 1+max list index any pixel reached            */
} orc_stats_t;

/* ------------------------------------------------------------------------------------------
 * A.6 RGB composite forward: vol_render.h:994-1062 (tile_based_vol_rendering_with_T) with the
 * inner loop vol_render_one_batch_v1 (:169-265).  T may be NULL (= K12, :782-847).
 * out[H,W,3], T[H,W] must be pre-initialised by the caller (zeros / ones): empty tiles are left
 * untouched (A.9-15).  margin (nullable, [H,W]) receives min over evaluated pairs of
 * |a*G*255 - 1| (how close the pixel came to a 1/255 threshold flip).
 * ---------------------------------------------------------------------------------------- */
void orc_composite_rgb_fwd(const float *mean, const float *cov, const float *color,
                           const float *alpha, const int32_t *start, const int32_t *end,
                           const int32_t *ids, float *out, float *T, const float *topleft,
                           int n_tiles_h, int n_tiles_w, float psx, float psy, int H, int W,
                           float thresh, float *margin, orc_stats_t *stats) {
  int64_t pe = 0, pb = 0, de = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : pe, pb, de)
  for (int tile = 0; tile < n_tiles_h * n_tiles_w; ++tile) {
    int s = start[tile];
    if (s == -1) continue;
    int n = end[tile] - s;
    if (n == 0) continue;
    int ty = tile / n_tiles_w, tx = tile % n_tiles_w;
    int reach = 0;
    for (int ly = 0; ly < TILE; ++ly)
      for (int lx = 0; lx < TILE; ++lx) {
        int gy = ty * TILE + ly, gx = tx * TILE + lx;
        if (gy >= H || gx >= W) continue;
        float pos[2] = {topleft[0] + gx * psx, topleft[1] + gy * psy};
        float o0 = 0.f, o1 = 0.f, o2 = 0.f, cum = 1.0f, mg = 1e30f;
        int i;
        for (i = 0; i < n; ++i) {
          if (cum < thresh) break;
          int g = ids[s + i];
          float a = fminf(alpha[g], ALPHA_CLAMP);
          float coeff = a * cum;
          float val = gauss2d_f64(mean + 2 * g, cov + 4 * g, pos);
          coeff *= val;
          ++pe;
          if (isnan(coeff)) coeff = 0.0f;
          float m_ = fabsf(a * val * 255.0f - 1.0f);
          if (m_ < mg) mg = m_;
          if (a * val < MIN_RENDER_ALPHA) continue;
          ++pb;
          o0 += color[3 * g + 0] * coeff;
          o1 += color[3 * g + 1] * coeff;
          o2 += color[3 * g + 2] * coeff;
          if (isnan(o0)) o0 = 0.f;
          if (isnan(o1)) o1 = 0.f;
          if (isnan(o2)) o2 = 0.f;
          cum *= (1 - a * val);
          if (isnan(cum) || cum < 0.0f || cum > 1.0f) cum = 0.0f;
        }
        if (i > reach) reach = i;
        out[3 * (gy * W + gx) + 0] = o0;
        out[3 * (gy * W + gx) + 1] = o1;
        out[3 * (gy * W + gx) + 2] = o2;
        if (T) T[gy * W + gx] = cum;
        if (margin) margin[gy * W + gx] = mg;
      }
    de += reach;
  }
  if (stats) { stats->pairs_evaluated = pe; stats->pairs_blended = pb; stats->d_eff = de; }
}

/* ------------------------------------------------------------------------------------------
 * A.6 RGB composite backward: vol_render.h:866-973 + vol_render_one_batch_backward (:318-418).
 * `final` = saved forward output INCLUDING any background term (renderer.py:1182,1190).
 * Gradients are accumulated in fp64 sinks (the reference uses fp32 atomics in nondeterministic
 * order) and ADDED into the caller's fp32 arrays.
 * ---------------------------------------------------------------------------------------- */
void orc_composite_rgb_bwd(int N, const float *mean, const float *cov, const float *color,
                           const float *alpha, const int32_t *start, const int32_t *end,
                           const int32_t *ids, const float *final, float *gmean, float *gcov,
                           float *gcolor, float *galpha, const float *gout, const float *topleft,
                           int n_tiles_h, int n_tiles_w, float psx, float psy, int H, int W,
                           float thresh) {
  double *dm = (double *)calloc((size_t)N * 2, 8), *dc = (double *)calloc((size_t)N * 4, 8),
         *dcol = (double *)calloc((size_t)N * 3, 8), *da = (double *)calloc((size_t)N, 8);
#pragma omp parallel for schedule(dynamic, 4)
  for (int tile = 0; tile < n_tiles_h * n_tiles_w; ++tile) {
    int s = start[tile];
    if (s == -1) continue;
    int n = end[tile] - s;
    if (n == 0) continue;
    int ty = tile / n_tiles_w, tx = tile % n_tiles_w;
    double *loc = (double *)calloc((size_t)n * 10, 8); /* [n][2 mean | 4 cov | 3 colour | 1 alpha] */
    for (int ly = 0; ly < TILE; ++ly)
      for (int lx = 0; lx < TILE; ++lx) {
        int gy = ty * TILE + ly, gx = tx * TILE + lx;
        if (gy >= H || gx >= W) continue;
        float pos[2] = {topleft[0] + gx * psx, topleft[1] + gy * psy};
        float ctt[3] = {0.f, 0.f, 0.f}, cum = 1.0f;
        const float *go = gout + 3 * (gy * W + gx);
        const float *fin = final + 3 * (gy * W + gx);
        for (int i = 0; i < n; ++i) {
          if (cum < thresh) break;
          int g = ids[s + i];
          float a = fminf(alpha[g], ALPHA_CLAMP);
          float G = gauss2d_f64(mean + 2 * g, cov + 4 * g, pos);
          if (a * G < MIN_RENDER_ALPHA) continue;
          float coeff = a * cum * G;
          ctt[0] += color[3 * g + 0] * coeff;
          ctt[1] += color[3 * g + 1] * coeff;
          ctt[2] += color[3 * g + 2] * coeff;
          loc[10 * i + 6] += (double)(coeff * go[0]);
          loc[10 * i + 7] += (double)(coeff * go[1]);
          loc[10 * i + 8] += (double)(coeff * go[2]);
          double partial_aG = 0.0;
          for (int j = 0; j < 3; ++j)
            partial_aG += (double)((color[3 * g + j] * cum - (fin[j] - ctt[j]) / (1 - a * G)) * go[j]);
          /* third arg is float in the reference signature: (float)(partial_aG * alpha_ * G) */
          gauss2d_bwd_f64(mean + 2 * g, cov + 4 * g, pos, (float)(partial_aG * a * G), loc + 10 * i,
                          loc + 10 * i + 2);
          loc[10 * i + 9] += (double)(float)(partial_aG * G);
          cum *= (1 - a * G);
        }
      }
    for (int i = 0; i < n; ++i) {
      int g = ids[s + i];
      const double *l = loc + 10 * i;
      if (l[0] == 0.0 && l[1] == 0.0 && l[6] == 0.0 && l[7] == 0.0 && l[8] == 0.0 && l[9] == 0.0) continue;
      ATOMIC_ADD(dm[2 * g], l[0]); ATOMIC_ADD(dm[2 * g + 1], l[1]);
      for (int k = 0; k < 4; ++k) ATOMIC_ADD(dc[4 * g + k], l[2 + k]);
      for (int k = 0; k < 3; ++k) ATOMIC_ADD(dcol[3 * g + k], l[6 + k]);
      ATOMIC_ADD(da[g], l[9]);
    }
    free(loc);
  }
  for (int i = 0; i < N * 2; ++i) gmean[i] += (float)dm[i];
  for (int i = 0; i < N * 4; ++i) gcov[i] += (float)dc[i];
  for (int i = 0; i < N * 3; ++i) gcolor[i] += (float)dcol[i];
  for (int i = 0; i < N; ++i) galpha[i] += (float)da[i];
  free(dm); free(dc); free(dcol); free(da);
}

/* ------------------------------------------------------------------------------------------
 * A.6 scalar composite: vol_render_scalar.h:47-102 (fwd, inner :14-45), :148-234 (bwd, inner
 * :104-146).  No NaN guards (A.6 note).
 * ---------------------------------------------------------------------------------------- */
void orc_composite_scalar_fwd(const float *mean, const float *cov, const float *scalar,
                              const float *alpha, const int32_t *start, const int32_t *end,
                              const int32_t *ids, float *out, float *T, const float *topleft,
                              int n_tiles_h, int n_tiles_w, float psx, float psy, int H, int W,
                              float thresh) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int tile = 0; tile < n_tiles_h * n_tiles_w; ++tile) {
    int s = start[tile];
    if (s == -1) continue;
    int n = end[tile] - s;
    if (n == 0) continue;
    int ty = tile / n_tiles_w, tx = tile % n_tiles_w;
    for (int ly = 0; ly < TILE; ++ly)
      for (int lx = 0; lx < TILE; ++lx) {
        int gy = ty * TILE + ly, gx = tx * TILE + lx;
        if (gy >= H || gx >= W) continue;
        float pos[2] = {topleft[0] + gx * psx, topleft[1] + gy * psy};
        float o = 0.f, cum = 1.0f;
        for (int i = 0; i < n; ++i) {
          if (cum < thresh) break;
          int g = ids[s + i];
          float a = fminf(alpha[g], ALPHA_CLAMP);
          float coeff = a * cum;
          float val = gauss2d_f64(mean + 2 * g, cov + 4 * g, pos);
          coeff *= val;
          if (a * val < MIN_RENDER_ALPHA) continue;
          o += coeff * scalar[g];
          cum *= (1 - a * val);
        }
        out[gy * W + gx] = o;
        if (T) T[gy * W + gx] = cum;
      }
  }
}

void orc_composite_scalar_bwd(int N, const float *mean, const float *cov, const float *scalar,
                              const float *alpha, const int32_t *start, const int32_t *end,
                              const int32_t *ids, const float *final, float *gmean, float *gcov,
                              float *gscalar, float *galpha, const float *gout,
                              const float *topleft, int n_tiles_h, int n_tiles_w, float psx,
                              float psy, int H, int W, float thresh) {
  double *dm = (double *)calloc((size_t)N * 2, 8), *dc = (double *)calloc((size_t)N * 4, 8),
         *ds = (double *)calloc((size_t)N, 8), *da = (double *)calloc((size_t)N, 8);
#pragma omp parallel for schedule(dynamic, 4)
  for (int tile = 0; tile < n_tiles_h * n_tiles_w; ++tile) {
    int s = start[tile];
    if (s == -1) continue;
    int n = end[tile] - s;
    if (n == 0) continue;
    int ty = tile / n_tiles_w, tx = tile % n_tiles_w;
    double *loc = (double *)calloc((size_t)n * 8, 8); /* [n][2 mean | 4 cov | 1 scalar | 1 alpha] */
    for (int ly = 0; ly < TILE; ++ly)
      for (int lx = 0; lx < TILE; ++lx) {
        int gy = ty * TILE + ly, gx = tx * TILE + lx;
        if (gy >= H || gx >= W) continue;
        float pos[2] = {topleft[0] + gx * psx, topleft[1] + gy * psy};
        float o = 0.f, cum = 1.0f;
        float go = gout[gy * W + gx], fin = final[gy * W + gx];
        for (int i = 0; i < n; ++i) {
          if (cum < thresh) break;
          int g = ids[s + i];
          float a = fminf(alpha[g], ALPHA_CLAMP);
          float G = gauss2d_f64(mean + 2 * g, cov + 4 * g, pos);
          if (a * G < MIN_RENDER_ALPHA) continue;
          float coeff = a * cum * G;
          o += scalar[g] * coeff;
          loc[8 * i + 6] += (double)(coeff * go);
          float partial_aG = 0.0f;
          partial_aG += go * (scalar[g] * cum - (fin - o) / (1 - a * G));
          gauss2d_bwd_f64(mean + 2 * g, cov + 4 * g, pos, partial_aG * a * G, loc + 8 * i, loc + 8 * i + 2);
          loc[8 * i + 7] += (double)(partial_aG * G);
          cum *= (1 - a * G);
        }
      }
    for (int i = 0; i < n; ++i) {
      int g = ids[s + i];
      const double *l = loc + 8 * i;
      if (l[0] == 0.0 && l[1] == 0.0 && l[6] == 0.0 && l[7] == 0.0) continue;
      ATOMIC_ADD(dm[2 * g], l[0]); ATOMIC_ADD(dm[2 * g + 1], l[1]);
      for (int k = 0; k < 4; ++k) ATOMIC_ADD(dc[4 * g + k], l[2 + k]);
      ATOMIC_ADD(ds[g], l[6]);
      ATOMIC_ADD(da[g], l[7]);
    }
    free(loc);
  }
  for (int i = 0; i < N * 2; ++i) gmean[i] += (float)dm[i];
  for (int i = 0; i < N * 4; ++i) gcov[i] += (float)dc[i];
  for (int i = 0; i < N; ++i) gscalar[i] += (float)ds[i];
  for (int i = 0; i < N; ++i) galpha[i] += (float)da[i];
  free(dm); free(dc); free(ds); free(da);
}

/* ------------------------------------------------------------------------------------------
 * A.7 SH basis (shencoder.h:13-56, first 16 functions; the binding dispatches C<=4 only,
 * render.cu:507-545) and pixel direction (vol_render_sh.h:48-65): the first NINE floats of the
 * c2w buffer are read as a packed 3x3 (A.9-10).
 * ---------------------------------------------------------------------------------------- */
static inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); } /* shencoder.h:4 */

void orc_sh_basis(const float *dir, float *o, int C) {
  float x = dir[0], y = dir[1], z = dir[2];
  float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  o[0] = 0.28209479177387814f;
  if (C <= 1) return;
  o[1] = -0.48860251190291987f * y;
  o[2] = 0.48860251190291987f * z;
  o[3] = -0.48860251190291987f * x;
  if (C <= 2) return;
  o[4] = 1.0925484305920792f * xy;
  o[5] = -1.0925484305920792f * yz;
  o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
  o[7] = -1.0925484305920792f * xz;
  o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
  if (C <= 3) return;
  o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
  o[10] = 2.8906114426405538f * xy * z;
  o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
  o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
  o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
  o[14] = 1.4453057213202769f * z * (x2 - y2);
  o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

void orc_pixel_dir(const float *pos3, const float *c2w9, float *dir) {
  for (int r = 0; r < 3; ++r)
    dir[r] = c2w9[3 * r] * pos3[0] + c2w9[3 * r + 1] * pos3[1] + c2w9[3 * r + 2] * pos3[2];
  float len = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
  dir[0] /= len; dir[1] /= len; dir[2] /= len;
}

static inline float sum_C(const float *a, const float *b, int CC) { /* vol_render_sh.h:18-26 */
  float s = 0.0f;
  for (int i = 0; i < CC; ++i) s += a[i] * b[i];
  return s;
}

/* SH composite forward: vol_render_sh.h:171-248 (inner :97-169); with bg: vol_render_bg.h:12-110.
 * bg_rgb == NULL -> no background; else empty tiles are written with bg (vol_render_bg.h:34-53) and
 * out += bg*T (:102-104).  T (nullable) is an oracle-only extra output. */
void orc_composite_sh_fwd(const float *mean, const float *cov, const float *sh, const float *alpha,
                          const int32_t *start, const int32_t *end, const int32_t *ids, float *out,
                          float *T, const float *topleft, const float *c2w9, int n_tiles_h,
                          int n_tiles_w, float psx, float psy, int H, int W, int C, float thresh,
                          const float *bg_rgb, float *margin, orc_stats_t *stats) {
  const int CC = C * C;
  int64_t pe = 0, pb = 0, de = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : pe, pb, de)
  for (int tile = 0; tile < n_tiles_h * n_tiles_w; ++tile) {
    int s = start[tile];
    int n = (s == -1) ? 0 : end[tile] - s;
    int ty = tile / n_tiles_w, tx = tile % n_tiles_w;
    int reach = 0;
    for (int ly = 0; ly < TILE; ++ly)
      for (int lx = 0; lx < TILE; ++lx) {
        int gy = ty * TILE + ly, gx = tx * TILE + lx;
        if (gy >= H || gx >= W) continue;
        if (n == 0) {
          if (bg_rgb)
            for (int c = 0; c < 3; ++c) out[3 * (gy * W + gx) + c] = bg_rgb[c];
          continue;
        }
        float pos[3] = {topleft[0] + gx * psx, topleft[1] + gy * psy, 1.0f};
        float dir[3], Y[16];
        orc_pixel_dir(pos, c2w9, dir);
        orc_sh_basis(dir, Y, C);
        float o[3] = {0.f, 0.f, 0.f}, cum = 1.0f, mg = 1e30f;
        int i;
        for (i = 0; i < n; ++i) {
          if (cum < thresh) break;
          int g = ids[s + i];
          float a = fminf(alpha[g], ALPHA_CLAMP);
          float coeff = a * cum;
          float val = gauss2d_f32(mean + 2 * g, cov + 4 * g, pos);
          coeff *= val;
          ++pe;
          float m_ = fabsf(a * val * 255.0f - 1.0f);
          if (m_ < mg) mg = m_;
          if (a * val < MIN_RENDER_ALPHA) continue;
          ++pb;
          if (isnan(coeff)) coeff = 0.0f;
          float y0 = sigmoidf_(sum_C(sh + (size_t)(3 * g + 0) * CC, Y, CC));
          float y1 = sigmoidf_(sum_C(sh + (size_t)(3 * g + 1) * CC, Y, CC));
          float y2 = sigmoidf_(sum_C(sh + (size_t)(3 * g + 2) * CC, Y, CC));
          if (isnan(y0 * coeff)) y0 = 0.f;
          if (isnan(y1 * coeff)) y1 = 0.f;
          if (isnan(y2 * coeff)) y2 = 0.f;
          o[0] += coeff * y0;
          o[1] += coeff * y1;
          o[2] += coeff * y2;
          cum *= (1 - a * val);
        }
        if (i > reach) reach = i;
        if (bg_rgb)
          for (int c = 0; c < 3; ++c) o[c] = o[c] + bg_rgb[c] * cum;
        for (int c = 0; c < 3; ++c) out[3 * (gy * W + gx) + c] = o[c];
        if (T) T[gy * W + gx] = cum;
        if (margin) margin[gy * W + gx] = mg;
      }
    de += reach;
  }
  if (stats) { stats->pairs_evaluated = pe; stats->pairs_blended = pb; stats->d_eff = de; }
}

/* ------------------------------------------------------------------------------------------
 * ARBITER (not a restatement of any reference kernel): the SH composite of A.6 evaluated in REAL
 * arithmetic -- every operation in fp64 on the fp32 inputs, pixel position = the fp32 value both
 * implementations use (one rounding of topleft + X*px).  Two correct fp32 implementations differ
 * from each other by rounding (the reference's fp32 `radial` formula kernels.h:172-193 cancels
 * catastrophically for thin Gaussians); where they disagree by more than the parity tolerance the
 * tests ask which of them is within tolerance of this result (tests/test_reference_gpu.py).
 * margin (nullable) = min over evaluated pairs of |a*G*255 - 1| in fp64.
 * ---------------------------------------------------------------------------------------- */
void orc_composite_sh_fwd_exact(const float *mean, const float *cov, const float *sh, const float *alpha,
                                const int32_t *start, const int32_t *end, const int32_t *ids, double *out,
                                double *T, const float *topleft, const float *c2w9, int n_tiles_h,
                                int n_tiles_w, float psx, float psy, int H, int W, int C, float thresh,
                                const float *bg_rgb, double *margin) {
  const int CC = C * C;
#pragma omp parallel for schedule(dynamic, 4)
  for (int tile = 0; tile < n_tiles_h * n_tiles_w; ++tile) {
    int s = start[tile];
    int n = (s == -1) ? 0 : end[tile] - s;
    int ty = tile / n_tiles_w, tx = tile % n_tiles_w;
    for (int ly = 0; ly < TILE; ++ly)
      for (int lx = 0; lx < TILE; ++lx) {
        int gy = ty * TILE + ly, gx = tx * TILE + lx;
        if (gy >= H || gx >= W) continue;
        double *o = out + 3 * ((size_t)gy * W + gx);
        if (n == 0) {
          for (int c = 0; c < 3; ++c) o[c] = bg_rgb ? (double)bg_rgb[c] : 0.0;
          if (T) T[(size_t)gy * W + gx] = 1.0;
          continue;
        }
        float posf[2] = {fmaf((float)gx, psx, topleft[0]), fmaf((float)gy, psy, topleft[1])};
        double px = posf[0], py = posf[1];
        double d[3], Y[16];
        for (int r = 0; r < 3; ++r)
          d[r] = (double)c2w9[3 * r] * px + (double)c2w9[3 * r + 1] * py + (double)c2w9[3 * r + 2];
        double len = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        double x = d[0] / len, y = d[1] / len, z = d[2] / len;
        {
          double xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
          Y[0] = 0.28209479177387814;
          Y[1] = -0.48860251190291987 * y; Y[2] = 0.48860251190291987 * z; Y[3] = -0.48860251190291987 * x;
          Y[4] = 1.0925484305920792 * xy; Y[5] = -1.0925484305920792 * yz;
          Y[6] = 0.94617469575755997 * z2 - 0.31539156525251999; Y[7] = -1.0925484305920792 * xz;
          Y[8] = 0.54627421529603959 * x2 - 0.54627421529603959 * y2;
          Y[9] = 0.59004358992664352 * y * (-3.0 * x2 + y2); Y[10] = 2.8906114426405538 * xy * z;
          Y[11] = 0.45704579946446572 * y * (1.0 - 5.0 * z2); Y[12] = 0.3731763325901154 * z * (5.0 * z2 - 3.0);
          Y[13] = 0.45704579946446572 * x * (1.0 - 5.0 * z2); Y[14] = 1.4453057213202769 * z * (x2 - y2);
          Y[15] = 0.59004358992664352 * x * (-x2 + 3.0 * y2);
        }
        double acc[3] = {0, 0, 0}, cum = 1.0, mg = 1e30;
        for (int i = 0; i < n; ++i) {
          if (cum < (double)thresh) break;
          int g = ids[s + i];
          double a = fmin((double)alpha[g], (double)ALPHA_CLAMP);
          double c0 = cov[4 * g], c1 = cov[4 * g + 1], c2 = cov[4 * g + 2], c3 = cov[4 * g + 3];
          double det = c0 * c3 - c1 * c2;
          double dx = px - (double)mean[2 * g], dy = py - (double)mean[2 * g + 1];
          double radial = ((dx * c3 - dy * c2) * dx + (-dx * c1 + dy * c0) * dy) / det;
          if (radial < 0.0) radial = 1000.0;
          double G = exp(-0.5 * radial);
          double aG = a * G;
          double m_ = fabs(aG * 255.0 - 1.0);
          if (m_ < mg) mg = m_;
          if (aG < (double)MIN_RENDER_ALPHA) continue;
          double w = aG * cum;
          for (int c = 0; c < 3; ++c) {
            double sdot = 0.0;
            const float *p = sh + (size_t)(3 * g + c) * CC;
            for (int k = 0; k < CC; ++k) sdot += (double)p[k] * Y[k];
            acc[c] += w / (1.0 + exp(-sdot));
          }
          cum *= (1.0 - aG);
        }
        for (int c = 0; c < 3; ++c) o[c] = acc[c] + (bg_rgb ? (double)bg_rgb[c] * cum : 0.0);
        if (T) T[(size_t)gy * W + gx] = cum;
        if (margin) margin[(size_t)gy * W + gx] = mg;
      }
  }
}

/* SH composite backward: vol_render_sh.h:353-455 (inner :268-351, backward_C :28-36); with bg:
 * vol_render_bg.h:131-242.  `final` is the saved forward output (incl. bg*T when bg is used). */
void orc_composite_sh_bwd(int N, const float *mean, const float *cov, const float *sh,
                          const float *alpha, const int32_t *start, const int32_t *end,
                          const int32_t *ids, const float *final, float *gmean, float *gcov,
                          float *gsh, float *galpha, const float *gout, const float *topleft,
                          const float *c2w9, int n_tiles_h, int n_tiles_w, float psx, float psy,
                          int H, int W, int C, float thresh) {
  const int CC = C * C;
  double *dm = (double *)calloc((size_t)N * 2, 8), *dc = (double *)calloc((size_t)N * 4, 8),
         *dsh = (double *)calloc((size_t)N * 3 * CC, 8), *da = (double *)calloc((size_t)N, 8);
#pragma omp parallel for schedule(dynamic, 4)
  for (int tile = 0; tile < n_tiles_h * n_tiles_w; ++tile) {
    int s = start[tile];
    if (s == -1) continue;
    int n = end[tile] - s;
    if (n == 0) continue;
    int ty = tile / n_tiles_w, tx = tile % n_tiles_w;
    const int LK = 7 + 3 * CC; /* [n][2 mean | 4 cov | 1 alpha | 3*CC sh] */
    double *loc = (double *)calloc((size_t)n * LK, 8);
    for (int ly = 0; ly < TILE; ++ly)
      for (int lx = 0; lx < TILE; ++lx) {
        int gy = ty * TILE + ly, gx = tx * TILE + lx;
        if (gy >= H || gx >= W) continue;
        float pos[3] = {topleft[0] + gx * psx, topleft[1] + gy * psy, 1.0f};
        float dir[3], Y[16];
        orc_pixel_dir(pos, c2w9, dir);
        orc_sh_basis(dir, Y, C);
        float o[3] = {0.f, 0.f, 0.f}, cum = 1.0f;
        const float *go = gout + 3 * (gy * W + gx);
        const float *fin = final + 3 * (gy * W + gx);
        for (int i = 0; i < n; ++i) {
          if (cum < thresh) break;
          int g = ids[s + i];
          float a = fminf(alpha[g], ALPHA_CLAMP);
          float G = gauss2d_f32(mean + 2 * g, cov + 4 * g, pos);
          if (a * G < MIN_RENDER_ALPHA) continue;
          float coeff = a * cum * G;
          if (isnan(coeff)) coeff = 0.0f;
          float y[3];
          for (int c = 0; c < 3; ++c) {
            y[c] = sigmoidf_(sum_C(sh + (size_t)(3 * g + c) * CC, Y, CC));
            if (isnan(y[c] * coeff)) y[c] = 0.f;
          }
          for (int c = 0; c < 3; ++c) o[c] += coeff * y[c];
          for (int c = 0; c < 3; ++c) {
            float gr = coeff * (y[c] * (1.0f - y[c])) * go[c];
            double *dst = loc + (size_t)LK * i + 7 + c * CC;
            for (int k = 0; k < CC; ++k) dst[k] += (double)(gr * Y[k]);
          }
          float partial_aG = 0.0f;
          for (int c = 0; c < 3; ++c)
            partial_aG += go[c] * (y[c] * cum - (fin[c] - o[c]) / (1 - a * G));
          gauss2d_bwd_f64(mean + 2 * g, cov + 4 * g, pos, partial_aG * a * G, loc + (size_t)LK * i, loc + (size_t)LK * i + 2);
          loc[(size_t)LK * i + 6] += (double)(partial_aG * G);
          cum *= (1 - a * G);
        }
      }
    for (int i = 0; i < n; ++i) {
      int g = ids[s + i];
      const double *l = loc + (size_t)LK * i;
      int any = 0;
      for (int k = 0; k < LK; ++k) any |= (l[k] != 0.0);
      if (!any) continue;
      ATOMIC_ADD(dm[2 * g], l[0]); ATOMIC_ADD(dm[2 * g + 1], l[1]);
      for (int k = 0; k < 4; ++k) ATOMIC_ADD(dc[4 * g + k], l[2 + k]);
      ATOMIC_ADD(da[g], l[6]);
      for (int k = 0; k < 3 * CC; ++k) ATOMIC_ADD(dsh[(size_t)g * 3 * CC + k], l[7 + k]);
    }
    free(loc);
  }
  for (int i = 0; i < N * 2; ++i) gmean[i] += (float)dm[i];
  for (int i = 0; i < N * 4; ++i) gcov[i] += (float)dc[i];
  for (size_t i = 0; i < (size_t)N * 3 * CC; ++i) gsh[i] += (float)dsh[i];
  for (int i = 0; i < N; ++i) galpha[i] += (float)da[i];
  free(dm); free(dc); free(dsh); free(da);
}
