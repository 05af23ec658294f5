// api.cu -- the extern "C" surface of libgsb200.so (include/gsb200.h).  Argument validation mirrors the
// reference's TORCH_CHECK contracts (gs/src/include/common.h:29-54) but reports through return codes.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "kernels.cuh"

namespace gsb {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static Camera make_camera(const float* c2w12, float fx, float fy, float cx, float cy, int W, int H) {
  Camera c;
  memset(&c, 0, sizeof(c));
  for (int r = 0; r < 3; ++r) {
    for (int k = 0; k < 3; ++k) c.R[3 * r + k] = c2w12[4 * r + k];
    c.t[r] = c2w12[4 * r + 3];
  }
  c.fx = fx; c.fy = fy; c.cx = cx; c.cy = cy;
  c.W = W; c.H = H;
  c.tiles_w = (W + 15) / 16;
  c.tiles_h = (H + 15) / 16;
  c.frustum_radius = 6.0f; c.tile_radius = 6.0f;
  c.skip_frustum = 1; c.depth_detach = 1;
  return c;
}

static Camera camera_from_abi(const gsb200_camera* in) {
  Camera c = make_camera(in->c2w, in->fx, in->fy, in->cx, in->cy, in->W, in->H);
  memcpy(c.fn, in->frustum_normals, sizeof(c.fn));
  memcpy(c.fp, in->frustum_pts, sizeof(c.fp));
  c.frustum_radius = in->frustum_radius;
  c.tile_radius = in->tile_radius;
  c.skip_frustum = in->skip_frustum_culling;
  c.depth_detach = in->depth_detach;
  return c;
}

static int check_tile(uint32_t tile_size, uint32_t th, uint32_t tw, uint32_t H, uint32_t W) {
  GSB_CHECK(tile_size == 16, GSB200_ERR_UNSUPPORTED, "tile_size %u unsupported (libgsb200 is specialised for 16)",
            tile_size);
  GSB_CHECK(th == (H + 15) / 16 && tw == (W + 15) / 16, GSB200_ERR_INVALID,
            "n_tiles (%u,%u) inconsistent with image %ux%u", th, tw, H, W);
  return GSB200_OK;
}

static int set_device(gsb200_ctx* ctx) {
  GSB_CHECK(ctx != nullptr, GSB200_ERR_INVALID, "null context");
  GSB_CUDA(cudaSetDevice(ctx->device));
  return GSB200_OK;
}

// records for the reference-compatible composite ops: pack (mean2d, cov2d, alpha, payload) into the arena
static int pack_for_compat(gsb200_ctx* ctx, uint32_t N, const float* mean, const float* cov, const float* alpha,
                           const float* payload, int pay_kind, cudaStream_t st) {
  int rc;
  if ((rc = ctx->splat.reserve((size_t)N * sizeof(Splat)))) return rc;
  if (pay_kind != 0 && (rc = ctx->pay.reserve((size_t)N * 16))) return rc;
  return launch_pack_splats(N, mean, cov, alpha, payload, pay_kind, ctx->splat.as<Splat>(), ctx->pay.as<float4>(), st);
}

static void fill_common(CompositeArgs& a, gsb200_ctx* ctx, const int32_t* start, const int32_t* end,
                        const int32_t* ids, const float* topleft, uint32_t th, uint32_t tw, float psx, float psy,
                        uint32_t H, uint32_t W, float thresh) {
  a.splat = ctx->splat.as<Splat>();
  a.pay = ctx->pay.as<float4>();
  a.ids = ids; a.start = start; a.end = end;
  a.topleft_ptr = topleft;
  a.psx = psx; a.psy = psy;
  a.H = (int)H; a.W = (int)W; a.tiles_w = (int)tw; a.tiles_h = (int)th;
  a.thresh = thresh;
  a.device = ctx->device;
}

static int next_evset(gsb200_ctx::EvSet** sets, int* cap, int* used, gsb200_ctx::EvSet** out) {
  if (*used == *cap) {
    int ncap = *cap ? *cap * 2 : 64;
    gsb200_ctx::EvSet* ns = new gsb200_ctx::EvSet[ncap];
    for (int i = 0; i < *cap; ++i) ns[i] = (*sets)[i];
    for (int i = *cap; i < ncap; ++i)
      for (int k = 0; k < 5; ++k) GSB_CUDA(cudaEventCreate(&ns[i].e[k]));
    delete[] *sets;
    *sets = ns; *cap = ncap;
  }
  *out = &(*sets)[(*used)++];
  return GSB200_OK;
}
#define GSB_EV(set, k, st)                                   \
  do {                                                       \
    if (set) GSB_CUDA(cudaEventRecord((set)->e[k], st));     \
  } while (0)

// Asynchronous-count mode.  A forward enqueues a 16-byte copy of its counters into a pinned ring slot and records an
// event; nobody waits for it.  poll_counts consumes the completed slots in order -- non-blocking (cudaEventQuery) from
// the forward / backward paths, blocking from gsb200_view_stats and the synchronous ops -- learns the capacity from
// them and remembers the first view whose lists did not fit; report_overflow turns that into GSB200_ERR_OVERFLOW once.
static int push_count(gsb200_ctx* ctx, cudaStream_t st);
static int poll_counts(gsb200_ctx* ctx, bool block) {
  while (ctx->ring_tail < ctx->ring_head) {
    const int slot = (int)(ctx->ring_tail % gsb200_ctx::kRing);
    if (block) {
      GSB_CUDA(cudaEventSynchronize(ctx->ev_ring[slot]));
    } else {
      cudaError_t q = cudaEventQuery(ctx->ev_ring[slot]);
      if (q == cudaErrorNotReady) break;
      GSB_CUDA(q);
    }
    const int64_t dup = ctx->h_ring[2 * slot], vis = ctx->h_ring[2 * slot + 1];
    if (dup > ctx->dup_seen) ctx->dup_seen = dup;
    if (ctx->profiling) ctx->sum_dup += dup;
    if (ctx->ring_gen[slot] == ctx->generation) { ctx->D = dup; ctx->N_visible = vis; }
    if (dup > ctx->ring_cap[slot] && ctx->overflow_gen == 0) {
      ctx->overflow_gen = ctx->ring_gen[slot]; ctx->overflow_dup = dup; ctx->overflow_cap = ctx->ring_cap[slot];
    }
    ++ctx->ring_tail;
  }
  return GSB200_OK;
}
static int report_overflow(gsb200_ctx* ctx) {
  if (ctx->overflow_gen == 0) return GSB200_OK;
  const long long g = ctx->overflow_gen, d = ctx->overflow_dup, c = ctx->overflow_cap;
  ctx->overflow_gen = 0;
  set_error("asynchronous-count mode: the forward with generation %lld expands to %lld duplicates but its tile sort "
            "covered %lld (capacity learnt from earlier views): its tile lists were truncated -- render that view "
            "again (the capacity has been raised)", g, d, c);
  return GSB200_ERR_OVERFLOW;
}
static int push_count(gsb200_ctx* ctx, cudaStream_t st) {
  if (!ctx->h_ring) GSB_CUDA(cudaHostAlloc((void**)&ctx->h_ring, gsb200_ctx::kRing * 2 * sizeof(int64_t), cudaHostAllocDefault));
  if (ctx->ring_head - ctx->ring_tail == (unsigned long long)gsb200_ctx::kRing) {  // ring full: consume the oldest
    const int old = (int)(ctx->ring_tail % gsb200_ctx::kRing);
    GSB_CUDA(cudaEventSynchronize(ctx->ev_ring[old]));
    int rc = poll_counts(ctx, false);
    if (rc) return rc;
  }
  const int slot = (int)(ctx->ring_head % gsb200_ctx::kRing);
  if (!ctx->ev_ring[slot]) GSB_CUDA(cudaEventCreateWithFlags(&ctx->ev_ring[slot], cudaEventDisableTiming));
  GSB_CUDA(cudaMemcpyAsync(ctx->h_ring + 2 * slot, ctx->d_total.p, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  GSB_CUDA(cudaEventRecord(ctx->ev_ring[slot], st));
  ctx->ring_gen[slot] = ctx->generation;
  ctx->ring_cap[slot] = ctx->dup_capacity;
  ++ctx->ring_head;
  return GSB200_OK;
}

static std::atomic<int64_t> g_generation{0};

}  // namespace gsb

using namespace gsb;

extern "C" {

const char* gsb200_last_error(void) { return g_err; }
int gsb200_version(void) { return 100; }

int gsb200_ctx_create(int device, gsb200_ctx** out) {
  GSB_CHECK(out != nullptr, GSB200_ERR_INVALID, "null out pointer");
  GSB_CUDA(cudaSetDevice(device));
  gsb200_ctx* c = new gsb200_ctx();
  c->device = device;
  cudaDeviceProp prop;
  GSB_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  GSB_CUDA(cudaHostAlloc((void**)&c->h_total, 4 * sizeof(int64_t), cudaHostAllocDefault));
  // tuning / A-B defaults from the environment (the options themselves: gsb200_ctx_set_option)
  if (const char* e = getenv("GSB200_BWD_SH_VARIANT")) c->bwd_sh_variant = atoi(e);
  if (const char* e = getenv("GSB200_ASYNC_COUNT")) c->async_count = atoi(e) ? 1 : 0;
  *out = c;
  return GSB200_OK;
}

int gsb200_ctx_destroy(gsb200_ctx* c) {
  if (!c) return GSB200_OK;
  cudaSetDevice(c->device);
  gsb::Buf* bufs[] = {&c->splat, &c->pay, &c->rect, &c->count, &c->incl, &c->ggeom, &c->gpay, &c->keys[0],
                      &c->keys[1], &c->vals[0], &c->vals[1], &c->cub_tmp, &c->start, &c->end, &c->d_total,
                      &c->d_overflow, &c->d_small, &c->dkeys[0], &c->dkeys[1], &c->perm[0], &c->perm[1], &c->d_stats,
                      &c->knn_keys[0], &c->knn_keys[1], &c->knn_vals[0], &c->knn_vals[1], &c->knn_pts, &c->knn_cells,
                      &c->knn_small};
  for (auto* b : bufs) b->release();
  if (c->h_total) cudaFreeHost(c->h_total);
  if (c->h_ring) cudaFreeHost(c->h_ring);
  for (auto& e : c->ev_ring) if (e) cudaEventDestroy(e);
  if (c->ev_total) cudaEventDestroy(c->ev_total);
  delete c;
  return GSB200_OK;
}

// ---- Part 1 -------------------------------------------------------------------------------------------
int gsb200_culling_gaussian_bsphere(const float* mean, const float* qvec, const float* svec, const float* normal,
                                    const float* pts, uint8_t* mask, uint32_t N, float thresh, gsb200_stream stream) {
  (void)qvec;
  GSB_CHECK(N == 0 || (mean && svec && normal && pts && mask), GSB200_ERR_INVALID, "culling_gaussian_bsphere: null tensor");
  return launch_cull_bsphere(N, mean, svec, normal, pts, mask, thresh, (cudaStream_t)stream);
}

int gsb200_tile_culling_aabb_start_end(gsb200_ctx* ctx, const int32_t* tl, const int32_t* br, int32_t* gaussian_ids,
                                       int32_t* start, int32_t* end, const float* depth, uint32_t N, uint32_t D,
                                       uint32_t th, uint32_t tw, gsb200_stream stream) {
  int rc;
  if ((rc = set_device(ctx))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = poll_counts(ctx, true))) return rc;
  GSB_CHECK(start && end, GSB200_ERR_INVALID, "tile_culling_aabb_start_end: null start/end");
  GSB_CHECK(tw <= 65535 && th <= 65535, GSB200_ERR_UNSUPPORTED, "tile grid too large");
  if (N > 0) {
    GSB_CHECK(tl && br && depth, GSB200_ERR_INVALID, "tile_culling_aabb_start_end: null tensor");
    if ((rc = ctx->count.reserve((size_t)N * 4))) return rc;
    if ((rc = ctx->rect.reserve((size_t)N * 8))) return rc;
    if ((rc = begin_total(ctx, st))) return rc;
    if ((rc = launch_count_from_aabb(N, tl, br, ctx->count.as<int32_t>(), ctx->rect.as<ushort4>(),
                                     ctx->d_total.as<unsigned long long>(), st)))
      return rc;
    if ((rc = request_total(ctx, st))) return rc;
    if ((rc = sort_depths_and_scan(ctx, N, depth, st))) return rc;
  }
  int64_t total = 0;
  if (N > 0 && (rc = wait_total(ctx, &total))) return rc;
  // the reference asserts size_h == N_with_dub on the host (aabb_culling.h:228)
  GSB_CHECK(total == (int64_t)D, GSB200_ERR_MISMATCH,
            "tile_culling_aabb_start_end: AABBs expand to %lld duplicates but gaussian_ids has %u entries",
            (long long)total, D);
  GSB_CHECK(D == 0 || gaussian_ids, GSB200_ERR_INVALID, "null gaussian_ids");
  return bin_and_sort(ctx, N, (int64_t)D, (int)th, (int)tw, gaussian_ids, start, end, st);
}

int gsb200_tile_based_vol_rendering_start_end_with_T(
    gsb200_ctx* ctx, const float* mean, const float* cov, const float* color, const float* alpha, const int32_t* start,
    const int32_t* end, const int32_t* ids, float* out, const float* topleft, uint32_t N, uint32_t D,
    uint32_t tile_size, uint32_t th, uint32_t tw, float psx, float psy, uint32_t H, uint32_t W, float thresh,
    float* T, gsb200_stream stream) {
  (void)D;
  int rc;
  if ((rc = set_device(ctx))) return rc;
  if ((rc = check_tile(tile_size, th, tw, H, W))) return rc;
  GSB_CHECK(out && start && end && topleft, GSB200_ERR_INVALID, "vol_rendering: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = pack_for_compat(ctx, N, mean, cov, alpha, color, 1, st))) return rc;
  CompositeArgs a;
  fill_common(a, ctx, start, end, ids, topleft, th, tw, psx, psy, H, W, thresh);
  a.out = out; a.T = T;
  return launch_composite_fwd(PAY_RGB, 1, false, a, st);
}

int gsb200_tile_based_vol_rendering_backward_start_end(
    gsb200_ctx* ctx, const float* mean, const float* cov, const float* color, const float* alpha, const int32_t* start,
    const int32_t* end, const int32_t* ids, const float* out, float* grad_mean, float* grad_cov, float* grad_color,
    float* grad_alpha, const float* grad_out, const float* topleft, uint32_t N, uint32_t D, uint32_t tile_size,
    uint32_t th, uint32_t tw, float psx, float psy, uint32_t H, uint32_t W, float thresh, gsb200_stream stream) {
  (void)D;
  int rc;
  if ((rc = set_device(ctx))) return rc;
  if ((rc = check_tile(tile_size, th, tw, H, W))) return rc;
  GSB_CHECK(out && grad_out && grad_mean && grad_cov && grad_color && grad_alpha && topleft, GSB200_ERR_INVALID,
            "vol_rendering_backward: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = pack_for_compat(ctx, N, mean, cov, alpha, color, 1, st))) return rc;
  CompositeArgs a;
  fill_common(a, ctx, start, end, ids, topleft, th, tw, psx, psy, H, W, thresh);
  a.fin = out; a.gout = grad_out;
  a.grad_mean = grad_mean; a.grad_cov = grad_cov; a.grad_pay = grad_color; a.grad_alpha = grad_alpha;
  return launch_composite_bwd(PAY_RGB, 1, false, false, a, st);
}

int gsb200_tile_based_vol_rendering_scalar(
    gsb200_ctx* ctx, const float* mean, const float* cov, const float* scalar, const float* alpha, const int32_t* start,
    const int32_t* end, const int32_t* ids, float* out, const float* topleft, uint32_t N, uint32_t D,
    uint32_t tile_size, uint32_t th, uint32_t tw, float psx, float psy, uint32_t H, uint32_t W, float thresh,
    float* T, gsb200_stream stream) {
  (void)D;
  int rc;
  if ((rc = set_device(ctx))) return rc;
  if ((rc = check_tile(tile_size, th, tw, H, W))) return rc;
  GSB_CHECK(out && start && end && topleft, GSB200_ERR_INVALID, "vol_rendering_scalar: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = pack_for_compat(ctx, N, mean, cov, alpha, scalar, 2, st))) return rc;
  CompositeArgs a;
  fill_common(a, ctx, start, end, ids, topleft, th, tw, psx, psy, H, W, thresh);
  a.out = out; a.T = T;
  return launch_composite_fwd(PAY_SCALAR, 1, false, a, st);
}

int gsb200_tile_based_vol_rendering_scalar_backward(
    gsb200_ctx* ctx, const float* mean, const float* cov, const float* scalar, const float* alpha, const int32_t* start,
    const int32_t* end, const int32_t* ids, const float* out, float* grad_mean, float* grad_cov, float* grad_scalar,
    float* grad_alpha, const float* grad_out, const float* topleft, uint32_t N, uint32_t D, uint32_t tile_size,
    uint32_t th, uint32_t tw, float psx, float psy, uint32_t H, uint32_t W, float thresh, gsb200_stream stream) {
  (void)D;
  int rc;
  if ((rc = set_device(ctx))) return rc;
  if ((rc = check_tile(tile_size, th, tw, H, W))) return rc;
  GSB_CHECK(out && grad_out && grad_mean && grad_cov && grad_scalar && grad_alpha && topleft, GSB200_ERR_INVALID,
            "vol_rendering_scalar_backward: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = pack_for_compat(ctx, N, mean, cov, alpha, scalar, 2, st))) return rc;
  CompositeArgs a;
  fill_common(a, ctx, start, end, ids, topleft, th, tw, psx, psy, H, W, thresh);
  a.fin = out; a.gout = grad_out;
  a.grad_mean = grad_mean; a.grad_cov = grad_cov; a.grad_pay = grad_scalar; a.grad_alpha = grad_alpha;
  return launch_composite_bwd(PAY_SCALAR, 1, false, false, a, st);
}

int gsb200_tile_based_vol_rendering_sh(
    gsb200_ctx* ctx, const float* mean, const float* cov, const float* sh, const float* alpha, const int32_t* start,
    const int32_t* end, const int32_t* ids, float* out, const float* topleft, const float* c2w, uint32_t N, uint32_t D,
    uint32_t tile_size, uint32_t th, uint32_t tw, float psx, float psy, uint32_t H, uint32_t W, uint32_t C,
    float thresh, const float* bg_rgb, gsb200_stream stream) {
  (void)D;
  int rc;
  if ((rc = set_device(ctx))) return rc;
  if ((rc = check_tile(tile_size, th, tw, H, W))) return rc;
  GSB_CHECK(C >= 1 && C <= 4, GSB200_ERR_UNSUPPORTED, "SH C=%u unsupported (reference dispatches 1..4, render.cu:507-545)", C);
  GSB_CHECK(out && start && end && topleft && c2w && (N == 0 || sh), GSB200_ERR_INVALID, "vol_rendering_sh: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = pack_for_compat(ctx, N, mean, cov, alpha, nullptr, 0, st))) return rc;
  CompositeArgs a;
  fill_common(a, ctx, start, end, ids, topleft, th, tw, psx, psy, H, W, thresh);
  a.sh = sh; a.c9_ptr = c2w; a.bg_rgb = bg_rgb;
  a.out = out;
  return launch_composite_fwd(PAY_SH, (int)C, false, a, st);
}

int gsb200_tile_based_vol_rendering_backward_sh(
    gsb200_ctx* ctx, const float* mean, const float* cov, const float* sh, const float* alpha, const int32_t* start,
    const int32_t* end, const int32_t* ids, const float* out, float* grad_mean, float* grad_cov, float* grad_sh,
    float* grad_alpha, const float* grad_out, const float* topleft, const float* c2w, uint32_t N, uint32_t D,
    uint32_t tile_size, uint32_t th, uint32_t tw, float psx, float psy, uint32_t H, uint32_t W, uint32_t C,
    float thresh, const float* bg_rgb, gsb200_stream stream) {
  (void)D;
  int rc;
  if ((rc = set_device(ctx))) return rc;
  if ((rc = check_tile(tile_size, th, tw, H, W))) return rc;
  GSB_CHECK(C >= 1 && C <= 4, GSB200_ERR_UNSUPPORTED, "SH C=%u unsupported", C);
  GSB_CHECK(out && grad_out && grad_mean && grad_cov && grad_sh && grad_alpha && topleft && c2w, GSB200_ERR_INVALID,
            "vol_rendering_backward_sh: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = pack_for_compat(ctx, N, mean, cov, alpha, nullptr, 0, st))) return rc;
  CompositeArgs a;
  fill_common(a, ctx, start, end, ids, topleft, th, tw, psx, psy, H, W, thresh);
  a.sh = sh; a.c9_ptr = c2w; a.bg_rgb = bg_rgb;
  a.fin = out; a.gout = grad_out;
  a.grad_mean = grad_mean; a.grad_cov = grad_cov; a.grad_pay = grad_sh; a.grad_alpha = grad_alpha;
  if (C >= 2 && ctx->bwd_sh_variant == 0) return launch_composite_bwd_sh((int)C, false, a, st);
  return launch_composite_bwd(PAY_SH, (int)C, false, false, a, st);
}

// ---- Part 2 -------------------------------------------------------------------------------------------
int gsb200_project_gaussians_forward(const float* mean, const float* qvec, const float* svec, const float* h_c2w,
                                     uint32_t N, float* mean2d, float* cov2d, float* JW, float* depth,
                                     gsb200_stream stream) {
  GSB_CHECK(h_c2w && (N == 0 || (mean && qvec && svec && mean2d && cov2d && depth)), GSB200_ERR_INVALID,
            "project_gaussians_forward: null tensor");
  Camera cam = make_camera(h_c2w, 1.f, 1.f, 0.f, 0.f, 16, 16);
  return launch_project_fwd(N, mean, qvec, svec, cam, mean2d, cov2d, JW, depth, (cudaStream_t)stream);
}

int gsb200_project_gaussians_backward(const float* mean, const float* qvec, const float* svec, const float* h_c2w,
                                      uint32_t N, int depth_detach, const float* g_m2, const float* g_cov,
                                      const float* g_depth, float* g_mean, float* g_qvec, float* g_svec,
                                      gsb200_stream stream) {
  GSB_CHECK(h_c2w && (N == 0 || (mean && qvec && svec && g_mean && g_qvec && g_svec)), GSB200_ERR_INVALID,
            "project_gaussians_backward: null tensor");
  Camera cam = make_camera(h_c2w, 1.f, 1.f, 0.f, 0.f, 16, 16);
  cam.depth_detach = depth_detach;
  return launch_project_bwd(N, mean, qvec, svec, cam, g_m2, g_cov, g_depth, g_mean, g_qvec, g_svec,
                            (cudaStream_t)stream);
}

int gsb200_tile_culling_aabb_count(gsb200_ctx* ctx, const float* mean2d, const float* cov2d, uint32_t N,
                                   uint32_t tile_size, float fx, float fy, float cx, float cy, uint32_t W, uint32_t H,
                                   float D, int32_t* tl, int32_t* br, int64_t* h_total, gsb200_stream stream) {
  int rc;
  if ((rc = set_device(ctx))) return rc;
  GSB_CHECK(tile_size >= 1, GSB200_ERR_INVALID, "tile_size must be positive");
  GSB_CHECK(h_total != nullptr, GSB200_ERR_INVALID, "null h_N_with_dub");
  if ((rc = poll_counts(ctx, true))) return rc;
  *h_total = 0;
  if (N == 0) return GSB200_OK;
  GSB_CHECK(mean2d && cov2d && tl && br, GSB200_ERR_INVALID, "tile_culling_aabb_count: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = begin_total(ctx, st))) return rc;
  if ((rc = launch_aabb_count(N, mean2d, cov2d, (int)tile_size, fx, fy, cx, cy, (int)W, (int)H, D, tl, br,
                              ctx->d_total.as<unsigned long long>(), st)))
    return rc;
  if ((rc = request_total(ctx, st))) return rc;
  return wait_total(ctx, h_total);
}

// ---- Part 3 -------------------------------------------------------------------------------------------
int gsb200_render_forward(gsb200_ctx* ctx, const gsb200_camera* camin, const gsb200_view_in* in,
                          const gsb200_view_out* out, gsb200_stream stream) {
  int rc;
  if ((rc = set_device(ctx))) return rc;
  GSB_CHECK(camin && in && out, GSB200_ERR_INVALID, "render_forward: null argument");
  GSB_CHECK(out->rgb && out->mean2d && out->cov2d && out->depthg && out->mask, GSB200_ERR_INVALID,
            "render_forward: rgb / mean2d / cov2d / depthg / mask outputs are required");
  const bool is_sh = in->sh != nullptr;
  GSB_CHECK(is_sh || in->color, GSB200_ERR_INVALID, "render_forward: need color or sh");
  GSB_CHECK(!is_sh || (in->C >= 1 && in->C <= 4), GSB200_ERR_UNSUPPORTED, "SH C=%d unsupported", in->C);
  GSB_CHECK((in->act & ~7) == 0, GSB200_ERR_INVALID, "render_forward: unknown activation bits 0x%x", in->act);
  const bool extras = !is_sh && out->depth && out->opacity && out->z2;
  GSB_CHECK(is_sh || extras || (!out->depth && !out->opacity && !out->z2), GSB200_ERR_INVALID,
            "render_forward: depth / opacity / z2 must be given together");
  cudaStream_t st = (cudaStream_t)stream;
  Camera cam = camera_from_abi(camin);
  GSB_CHECK(cam.tiles_w <= 65535 && cam.tiles_h <= 65535, GSB200_ERR_UNSUPPORTED, "tile grid too large");
  const uint32_t N = in->N;
  const uint32_t T = (uint32_t)cam.tiles_w * cam.tiles_h;
  if ((rc = ctx->start.reserve((size_t)T * 4))) return rc;
  if ((rc = ctx->end.reserve((size_t)T * 4))) return rc;
  int64_t D = 0;
  gsb200_ctx::EvSet* ev = nullptr;
  if (ctx->profiling && (rc = next_evset(&ctx->fwd_sets, &ctx->fwd_cap, &ctx->fwd_used, &ev))) return rc;
  // counts of earlier asynchronous forwards that have arrived are consumed (non-blocking): the capacity learns from
  // them; an overflow among them is reported here, before this view is rendered on top of a stale capacity
  if ((rc = poll_counts(ctx, false))) return rc;
  if ((rc = report_overflow(ctx))) return rc;
  if (ctx->seen_N != N || ctx->seen_W != cam.W || ctx->seen_H != cam.H) {
    if ((rc = poll_counts(ctx, true))) return rc;  // (counts of the previous shape must not teach the new one)
    ctx->overflow_gen = 0;
    ctx->seen_N = N; ctx->seen_W = cam.W; ctx->seen_H = cam.H; ctx->dup_seen = 0;
  }
  const bool padded = ctx->async_count && ctx->dup_seen > 0 && N > 0;
  ctx->generation = ++g_generation;
  GSB_EV(ev, 0, st);
  if (N > 0) {
    GSB_CHECK(in->mean && in->qvec && in->svec && in->alpha, GSB200_ERR_INVALID, "render_forward: null parameter tensor");
    GSB_CHECK((reinterpret_cast<uintptr_t>(in->qvec) & 15) == 0, GSB200_ERR_INVALID,
              "render_forward: qvec must be 16-byte aligned (it is read as float4)");
    if ((rc = ctx->splat.reserve((size_t)N * sizeof(Splat)))) return rc;
    if (!is_sh && (rc = ctx->pay.reserve((size_t)N * 16))) return rc;
    if ((rc = ctx->rect.reserve((size_t)N * 8))) return rc;
    if ((rc = ctx->count.reserve((size_t)N * 4))) return rc;
    if ((rc = reserve_depth_sort(ctx, N))) return rc;
    if ((rc = begin_total(ctx, st))) return rc;
    if ((rc = launch_preprocess(N, in->mean, in->qvec, in->svec, in->alpha, is_sh ? nullptr : in->color, in->act, cam,
                                out->mean2d, out->cov2d, out->depthg, out->mask, out->radii2d,
                                ctx->splat.as<Splat>(), ctx->pay.as<float4>(), ctx->rect.as<ushort4>(),
                                ctx->count.as<int32_t>(), ctx->dkeys[0].as<uint32_t>(), ctx->perm[0].as<int32_t>(),
                                ctx->d_total.as<unsigned long long>(), st)))
      return rc;
    if (padded) {
      // the host does not wait at all: the sort covers a capacity learnt from earlier views of this context (largest
      // count seen + 1/8), keys beyond the device-side count are padding; the exact count travels to a pinned ring
      // slot and is consumed by a later poll (next forward / backward on this context, or gsb200_view_stats)
      ctx->D = -1;
      ctx->N_visible = -1;
      ctx->dup_capacity = ctx->dup_seen + ctx->dup_seen / 8 + 4096;
      D = -1;
      if ((rc = push_count(ctx, st))) return rc;
    } else {
      if ((rc = request_total(ctx, st))) return rc;
    }
    GSB_EV(ev, 1, st);
    if ((rc = sort_depths_and_scan(ctx, N, out->depthg, st, /*keys_ready=*/true))) return rc;  // GPU keeps working ...
    if (!padded) {
      int64_t tot[2] = {0, 0};
      if ((rc = wait_total(ctx, tot))) return rc;  // ... while the host waits only for the 16-byte counters (the one
                                                   // host wait of the view; the reference blocks twice + 5 cudaMalloc/Free)
      D = tot[0];
      ctx->N_visible = ctx->h_total[1];
      ctx->dup_capacity = D;
      if (D > ctx->dup_seen) ctx->dup_seen = D;
    }
  } else {
    GSB_EV(ev, 1, st);
    ctx->N_visible = 0;
  }
  GSB_EV(ev, 2, st);
  if (out->h_num_dup) *out->h_num_dup = D;
  if ((rc = bin_and_sort(ctx, N, padded ? ctx->dup_capacity : D, cam.tiles_h, cam.tiles_w, nullptr,
                         ctx->start.as<int32_t>(), ctx->end.as<int32_t>(), st, padded)))
    return rc;
  if (out->h_generation) *out->h_generation = ctx->generation;
  GSB_EV(ev, 3, st);
  ctx->N = N; ctx->cam = cam; ctx->mode = is_sh ? PAY_SH : PAY_RGB; ctx->C = is_sh ? in->C : 1;
  if (ctx->profiling) {
    if (D >= 0) ctx->sum_dup += D;  // (asynchronous-count mode: added when the count is consumed)
    if (!ctx->d_stats.p) {
      if ((rc = ctx->d_stats.reserve(16))) return rc;
      GSB_CUDA(cudaMemsetAsync(ctx->d_stats.p, 0, 16, st));
    }
  }

  CompositeArgs a;
  a.splat = ctx->splat.as<Splat>();
  a.pay = ctx->pay.as<float4>();
  a.sh = in->sh;
  a.ids = ctx->vals[ctx->sorted_sel].as<int32_t>();
  a.start = ctx->start.as<int32_t>(); a.end = ctx->end.as<int32_t>();
  a.tlx = -cam.cx / cam.fx; a.tly = -cam.cy / cam.fy;  // gaussian_splatting.py:1274-1276
  a.psx = 1.0f / cam.fx; a.psy = 1.0f / cam.fy;
  a.H = cam.H; a.W = cam.W; a.tiles_w = cam.tiles_w; a.tiles_h = cam.tiles_h;
  a.thresh = camin->T_thresh;
  memcpy(a.c9, in->sh_c2w9, sizeof(a.c9));
  a.bg = in->bg; a.bg_rgb = in->bg_rgb;
  a.device = ctx->device;
  a.write_empty = 1;
  a.out = out->rgb; a.T = out->T;
  a.depth = out->depth; a.opacity = out->opacity; a.z2 = out->z2;
  a.stats = ctx->profiling ? ctx->d_stats.as<unsigned long long>() : nullptr;
  if ((rc = launch_composite_fwd(is_sh ? PAY_SH : PAY_RGB, is_sh ? in->C : 1, extras, a, st))) return rc;
  GSB_EV(ev, 4, st);
  return GSB200_OK;
}

int gsb200_render_backward(gsb200_ctx* ctx, const gsb200_camera* camin, const gsb200_view_in* in,
                           const gsb200_view_grads* g, gsb200_stream stream) {
  int rc;
  if ((rc = set_device(ctx))) return rc;
  GSB_CHECK(camin && in && g, GSB200_ERR_INVALID, "render_backward: null argument");
  const bool is_sh = in->sh != nullptr;
  GSB_CHECK(ctx->N == in->N && ctx->mode == (is_sh ? PAY_SH : PAY_RGB), GSB200_ERR_INVALID,
            "render_backward: context does not hold the forward state of this view (N=%u vs %u)", ctx->N, in->N);
  GSB_CHECK(g->generation == 0 || g->generation == ctx->generation, GSB200_ERR_INVALID,
            "render_backward: the context was overwritten by a later forward (this view is generation %lld, the context "
            "holds %lld): give every in-flight view its own context",
            (long long)g->generation, (long long)ctx->generation);
  // asynchronous-count mode: consume the counts that have arrived (no wait) and reject a view found truncated
  if ((rc = poll_counts(ctx, false))) return rc;
  if ((rc = report_overflow(ctx))) return rc;
  GSB_CHECK(in->N == 0 || ((reinterpret_cast<uintptr_t>(in->qvec) | reinterpret_cast<uintptr_t>(g->g_qvec)) & 15) == 0,
            GSB200_ERR_INVALID, "render_backward: qvec and g_qvec must be 16-byte aligned (float4 accesses)");
  GSB_CHECK(g->rgb && g->mask && g->g_mean && g->g_qvec && g->g_svec && g->g_alpha, GSB200_ERR_INVALID,
            "render_backward: rgb, mask and g_mean/g_qvec/g_svec/g_alpha are required");
  GSB_CHECK(is_sh ? (g->g_sh != nullptr) : (g->g_color != nullptr), GSB200_ERR_INVALID,
            "render_backward: g_sh (SH path) or g_color (RGB path) is required");
  cudaStream_t st = (cudaStream_t)stream;
  Camera cam = camera_from_abi(camin);
  const uint32_t N = in->N;
  if (N == 0) return GSB200_OK;
  const bool extras = !is_sh && (g->g_depth || g->g_opacity || g->g_z2);
  GSB_CHECK(!extras || (g->depth && g->opacity && g->z2), GSB200_ERR_INVALID,
            "render_backward: saved depth / opacity / z2 images are required with their gradients");
  if ((rc = ctx->ggeom.reserve((size_t)N * 32))) return rc;
  gsb200_ctx::EvSet* ev = nullptr;
  if (ctx->profiling && (rc = next_evset(&ctx->bwd_sets, &ctx->bwd_cap, &ctx->bwd_used, &ev))) return rc;
  GSB_EV(ev, 0, st);
  GSB_CUDA(cudaMemsetAsync(ctx->ggeom.p, 0, (size_t)N * 32, st));
  if (!is_sh) {
    if ((rc = ctx->gpay.reserve((size_t)N * 16))) return rc;
    GSB_CUDA(cudaMemsetAsync(ctx->gpay.p, 0, (size_t)N * 16, st));
  }
  CompositeArgs a;
  a.splat = ctx->splat.as<Splat>();
  a.pay = ctx->pay.as<float4>();
  a.sh = in->sh;
  a.ids = ctx->vals[ctx->sorted_sel].as<int32_t>();
  a.start = ctx->start.as<int32_t>(); a.end = ctx->end.as<int32_t>();
  a.tlx = -cam.cx / cam.fx; a.tly = -cam.cy / cam.fy;
  a.psx = 1.0f / cam.fx; a.psy = 1.0f / cam.fy;
  a.H = cam.H; a.W = cam.W; a.tiles_w = cam.tiles_w; a.tiles_h = cam.tiles_h;
  a.thresh = camin->T_thresh;
  memcpy(a.c9, in->sh_c2w9, sizeof(a.c9));
  a.bg = in->bg; a.bg_rgb = in->bg_rgb;
  a.device = ctx->device;
  a.fin = g->rgb; a.gout = g->g_rgb;
  a.fin_depth = g->depth; a.g_depth = g->g_depth;
  a.fin_opacity = g->opacity; a.g_opacity = g->g_opacity;
  a.fin_z2 = g->z2; a.g_z2 = g->g_z2;
  a.ggeom = ctx->ggeom.as<float>();
  a.gpay = ctx->gpay.as<float>();
  a.grad_pay = g->g_sh;
  a.g_bg = g->g_bg;
  if (ctx->D != 0 || g->g_bg) {  // (D < 0: count not consumed yet in asynchronous-count mode)
    if (is_sh && in->C >= 2 && ctx->bwd_sh_variant == 0) {
      if ((rc = launch_composite_bwd_sh(in->C, true, a, st))) return rc;
    } else if ((rc = launch_composite_bwd(is_sh ? PAY_SH : PAY_RGB, is_sh ? in->C : 1, extras, true, a, st))) {
      return rc;
    }
  }
  GSB_EV(ev, 1, st);
  rc = launch_project_bwd_fused(N, in->mean, in->qvec, in->svec, in->alpha, is_sh ? nullptr : in->color, in->act,
                                g->mask, cam, ctx->ggeom.as<float4>(), is_sh ? nullptr : ctx->gpay.as<float4>(),
                                g->g_mean, g->g_qvec, g->g_svec, g->g_alpha, is_sh ? nullptr : g->g_color,
                                g->g_mean2d, g->accumulate, g->touched, st);
  if (rc) return rc;
  GSB_EV(ev, 2, st);
  return GSB200_OK;
}

// ---- Part 4 -------------------------------------------------------------------------------------------
int gsb200_adam_step(gsb200_ctx* ctx, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                     uint64_t total, const gsb200_adam_field* fields, int32_t n_fields, double beta1, double beta2,
                     double eps, int64_t step, float grad_scale, gsb200_stream stream) {
  int rc;
  if ((rc = set_device(ctx))) return rc;
  if (total == 0) return GSB200_OK;
  GSB_CHECK(param && grad && exp_avg && exp_avg_sq && fields, GSB200_ERR_INVALID, "adam_step: null argument");
  GSB_CHECK(n_fields >= 1 && n_fields <= 8, GSB200_ERR_INVALID, "adam_step: n_fields=%d not in 1..8", n_fields);
  GSB_CHECK(step >= 1, GSB200_ERR_INVALID, "adam_step: step counts from 1 (torch state['step']), got %lld",
            (long long)step);
  GSB_CHECK(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
              reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0,
            GSB200_ERR_INVALID, "adam_step: buffers must be 16-byte aligned");
  AdamFields F{};
  F.n = n_fields;
  uint64_t at = 0;
  // bias corrections in double, like the Python floats torch computes them in (torch/optim/adam.py)
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  for (int f = 0; f < n_fields; ++f) {
    GSB_CHECK(fields[f].begin == at, GSB200_ERR_INVALID,
              "adam_step: fields must tile the buffer in order (field %d begins at %llu, expected %llu)", f,
              (unsigned long long)fields[f].begin, (unsigned long long)at);
    F.begin[f] = at;
    F.step_size[f] = (float)(fields[f].lr / bc1);
    at += fields[f].count;
  }
  GSB_CHECK(at == total, GSB200_ERR_INVALID, "adam_step: fields cover %llu of %llu elements",
            (unsigned long long)at, (unsigned long long)total);
  for (int f = n_fields; f <= 8; ++f) F.begin[f] = total;
  AdamScalars K;
  K.beta2 = (float)beta2;
  K.one_minus_beta1 = (float)(1.0 - beta1);
  K.one_minus_beta2 = (float)(1.0 - beta2);
  K.eps = (float)eps;
  K.bc2_sqrt = (float)sqrt(bc2);
  K.grad_scale = grad_scale;
  return launch_adam_flat(param, grad, exp_avg, exp_avg_sq, total, F, K, (cudaStream_t)stream);
}

int gsb200_ctx_set_profiling(gsb200_ctx* ctx, int enable) {
  int rc;
  if ((rc = set_device(ctx))) return rc;
  ctx->profiling = enable ? 1 : 0;
  return GSB200_OK;
}

int gsb200_ctx_get_profile(gsb200_ctx* ctx, float* h_ms, int64_t* h_counts, int reset) {
  int rc;
  if ((rc = set_device(ctx))) return rc;
  GSB_CHECK(h_ms && h_counts, GSB200_ERR_INVALID, "null output");
  GSB_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < 6; ++i) h_ms[i] = 0.f;
  for (int i = 0; i < ctx->fwd_used; ++i)
    for (int k = 0; k < 4; ++k) {
      float ms = 0.f;
      GSB_CUDA(cudaEventElapsedTime(&ms, ctx->fwd_sets[i].e[k], ctx->fwd_sets[i].e[k + 1]));
      h_ms[k] += ms;
    }
  for (int i = 0; i < ctx->bwd_used; ++i)
    for (int k = 0; k < 2; ++k) {
      float ms = 0.f;
      GSB_CUDA(cudaEventElapsedTime(&ms, ctx->bwd_sets[i].e[k], ctx->bwd_sets[i].e[k + 1]));
      h_ms[4 + k] += ms;
    }
  unsigned long long st2[2] = {0, 0};
  if (ctx->d_stats.p) GSB_CUDA(cudaMemcpy(st2, ctx->d_stats.p, 16, cudaMemcpyDeviceToHost));
  h_counts[0] = ctx->fwd_used; h_counts[1] = ctx->bwd_used; h_counts[2] = ctx->sum_dup;
  h_counts[3] = (int64_t)st2[0]; h_counts[4] = (int64_t)st2[1];
  if (reset) {
    ctx->fwd_used = 0; ctx->bwd_used = 0; ctx->sum_dup = 0;
    if (ctx->d_stats.p) GSB_CUDA(cudaMemset(ctx->d_stats.p, 0, 16));
  }
  return GSB200_OK;
}

int gsb200_view_stats(gsb200_ctx* ctx, int64_t* h_out, gsb200_stream stream) {
  int rc;
  if ((rc = set_device(ctx))) return rc;
  GSB_CHECK(h_out, GSB200_ERR_INVALID, "null h_out");
  GSB_CHECK(ctx->generation != 0, GSB200_ERR_INVALID, "view_stats: no forward has run on this context");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = poll_counts(ctx, true))) return rc;  // blocks until every count in flight has arrived
  const int rc_count = report_overflow(ctx);
  const uint32_t T = (uint32_t)ctx->cam.tiles_w * (uint32_t)ctx->cam.tiles_h;
  if ((rc = ctx->d_overflow.reserve(2 * sizeof(int32_t)))) return rc;
  int32_t* d_max = ctx->d_overflow.as<int32_t>() + 1;
  if ((rc = launch_max_list(T, ctx->start.as<int32_t>(), ctx->end.as<int32_t>(), d_max, st))) return rc;
  int32_t h_max = 0;
  GSB_CUDA(cudaMemcpyAsync(&h_max, d_max, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  GSB_CUDA(cudaStreamSynchronize(st));
  h_out[0] = ctx->D;
  h_out[1] = ctx->N_visible;
  h_out[2] = h_max;
  return rc_count;  // GSB200_ERR_OVERFLOW: the numbers are valid, the view's lists are not
}

int gsb200_ctx_set_option(gsb200_ctx* ctx, int option, int64_t value) {
  GSB_CHECK(ctx != nullptr, GSB200_ERR_INVALID, "null context");
  switch (option) {
    case GSB200_OPT_BWD_SH_VARIANT:
      GSB_CHECK(value == 0 || value == 1, GSB200_ERR_INVALID, "bwd_sh_variant must be 0 or 1");
      ctx->bwd_sh_variant = (int)value;
      return GSB200_OK;
    case GSB200_OPT_ASYNC_COUNT:
      ctx->async_count = value ? 1 : 0;
      return GSB200_OK;
    default:
      break;
  }
  set_error("ctx_set_option: unknown option %d", option);
  return GSB200_ERR_INVALID;
}

}  // extern "C"
