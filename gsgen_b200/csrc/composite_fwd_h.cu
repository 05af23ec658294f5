// composite_fwd_h.cu -- SH forward composite with HALF-WARP pixel blocks (reference kernels
// tile_based_vol_rendering_sh_entry<C> [_with_bg], vol_render_sh.h:171-248, vol_render_bg.h:12-110).
//
// Why (measured round 2): the SH composites are instruction-issue bound, and half of the issued lanes do nothing useful:
// at C3 a warp (8x4 pixels) evaluates ~99 Gaussians whose box touches its block while a pixel blends ~51 of them.  The
// two-pixels-per-thread experiment (8x8 blocks, profiles/r2_bench_call10_fwd2*.json) confirmed the direction of the
// effect: bigger blocks -> same instruction count for half the warps -> slower.  So go the other way: the two half-warps
// own a 4x4 block each and walk their OWN hit lists in lock-step -- lanes 0-15 composite the k-th Gaussian that touches
// the left 4x4 block while lanes 16-31 composite the k-th one that touches the right block.  A warp iterates
// max(hits_left, hits_right) times instead of |hits_left U hits_right|; for splats smaller than a block that is up to
// half.  Shared-memory reads become two-address (one per half-warp): still conflict-free, still 2 wavefronts per
// LDS.128.  Per-pixel semantics are those of composite_fwd.cu (same records, thresholds, early-termination rule).
#include "composite_common.cuh"

namespace gsb {

#ifndef GSB_FWDH_B
#define GSB_FWDH_B 64
#endif

template <int C, int B>
__global__ void __launch_bounds__(kCtaThreads)
k_composite_fwd_shh(const CompositeArgs a) {
  using L = StageLayout<PAY_SH, C, B, false>;
  using PT = PayTraits<PAY_SH, C>;
  constexpr int CC = PT::CC;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t s_bar[2];
  __shared__ int s_reach;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile = tile_y * a.tiles_w + tile_x;
  const float tlx = a.topleft_ptr ? a.topleft_ptr[0] : a.tlx;
  const float tly = a.topleft_ptr ? a.topleft_ptr[1] : a.tly;
  // warp: 8 wide x 4 tall (as composite_fwd.cu); half-warp h owns columns 4h .. 4h+3
  const int half = lane >> 4, l16 = lane & 15;
  const int bx0 = tile_x * kTile + (warp & 1) * 8, by0 = tile_y * kTile + (warp >> 1) * 4;
  const int gx = bx0 + 4 * half + (l16 & 3);
  const int gy = by0 + (l16 >> 2);
  const bool inside = (gx < a.W) && (gy < a.H);
  const float px = fmaf((float)gx, a.psx, tlx), py = fmaf((float)gy, a.psy, tly);
  const int pix = gy * a.W + gx;
  // camera-plane extents of the two 4x4 blocks (every lane needs both: it tests one list entry against both)
  const float XA0 = fmaf((float)bx0, a.psx, tlx), XA1 = fmaf((float)(bx0 + 3), a.psx, tlx);
  const float XB0 = fmaf((float)(bx0 + 4), a.psx, tlx), XB1 = fmaf((float)(bx0 + 7), a.psx, tlx);
  const float YB0 = fmaf((float)by0, a.psy, tly), YB1 = fmaf((float)(by0 + 3), a.psy, tly);

  const int s0 = a.start[tile];
  const int n = (s0 < 0) ? 0 : (a.end[tile] - s0);
  if (n <= 0) {  // empty tile: see composite_fwd.cu (A.9-15; sh_with_bg writes the background, vol_render_bg.h:34-53)
    if (!inside) return;
    if (a.bg_rgb) {
      a.out[3 * pix + 0] = a.bg_rgb[0]; a.out[3 * pix + 1] = a.bg_rgb[1]; a.out[3 * pix + 2] = a.bg_rgb[2];
    } else if (a.write_empty) {
      a.out[3 * pix + 0] = 0.f; a.out[3 * pix + 1] = 0.f; a.out[3 * pix + 2] = 0.f;
    }
    if (a.write_empty && a.T) a.T[pix] = 1.0f;
    return;
  }

  const bool use_bulk = PT::kBulkOk && ((reinterpret_cast<uintptr_t>(a.sh) & 15) == 0);
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    fence_mbar_init();
    s_reach = 0;
  }
  __syncthreads();

  float T = 1.0f;
  int reach = 0, staged = 0;  // statistics only (a.stats)
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
  bool done = !inside || (1.0f < a.thresh);
  float Y[CC];
  {
    float c9[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) c9[k] = a.c9_ptr ? a.c9_ptr[k] : a.c9[k];
    float d[3];
    pixel_dir(px, py, c9, d);
    sh_basis<C>(d[0], d[1], d[2], Y);
  }

  const int nb = (n + B - 1) / B;
  const int32_t* ids = a.ids + s0;
  {
    int cnt0 = min(B, n);
    int id0 = (tid < cnt0) ? ids[tid] : 0;
    if (use_bulk && tid == 0) mbar_arrive_expect_tx(&s_bar[0], (uint32_t)cnt0 * 3 * CC * 4);
    if (tid < B) stage_entry<PAY_SH, C, B, false>(a, smem, tid, id0, tid < cnt0, use_bulk, &s_bar[0]);
    else cp_async_commit();
    staged = cnt0;
  }
  int id_next = 0;
  if (nb > 1) { int j = B + tid; id_next = (tid < B && j < n) ? ids[j] : 0; }

  bool warp_done = __all_sync(kFull, done);
  cp_async_wait<0>();
  if (use_bulk) mbar_wait(&s_bar[0], 0u);
  __syncthreads();
  for (int b = 0; b < nb; ++b) {
    unsigned char* st = smem + (b & 1) * L::kBytes;
    const int cnt = min(B, n - b * B);
    const bool has_next = (b + 1 < nb);
    if (has_next) {
      const int cntn = min(B, n - (b + 1) * B);
      uint64_t* barn = &s_bar[(b + 1) & 1];
      if (use_bulk && tid == 0) mbar_arrive_expect_tx(barn, (uint32_t)cntn * 3 * CC * 4);
      if (tid < B) stage_entry<PAY_SH, C, B, false>(a, smem + ((b + 1) & 1) * L::kBytes, tid, id_next, tid < cntn,
                                                    use_bulk, barn);
      else cp_async_commit();
      staged += cntn;
      if (b + 2 < nb) { int j = (b + 2) * B + tid; id_next = (tid < B && j < n) ? ids[j] : 0; }
    }

    if (!warp_done) {
      const float4* sg0 = reinterpret_cast<const float4*>(st + L::kG0);
      const float4* sg1 = reinterpret_cast<const float4*>(st + L::kG1);
      const float* pay = reinterpret_cast<const float*>(st + L::kPay);
      for (int r = 0; r * 32 < cnt; ++r) {
        const int j = r * 32 + lane;
        bool hitA = false, hitB = false;
        if (j < cnt) {  // the a*G >= 1/255 box of entry j against both 4x4 blocks
          const float4 e0 = sg0[j], e1 = sg1[j];
          const bool yov = (e0.y - e1.w <= YB1) && (e0.y + e1.w >= YB0);
          hitA = yov && (e0.x - e1.z <= XA1) && (e0.x + e1.z >= XA0);
          hitB = yov && (e0.x - e1.z <= XB1) && (e0.x + e1.z >= XB0);
        }
        const unsigned mA = __ballot_sync(kFull, hitA), mB = __ballot_sync(kFull, hitB);
        unsigned m = half ? mB : mA;  // this half-warp's own hit list
        while (__any_sync(kFull, m != 0u)) {
          const bool has = (m != 0u);
          const int jj = r * 32 + (has ? (__ffs(m) - 1) : 0);
          m &= m - 1;  // (0 stays 0)
          const float4 g0 = sg0[jj], g1 = sg1[jj];
          float dx = px - g0.x, dy = py - g0.y;
          float u = fmaf(g0.z, dx, g0.w * dy);
          float v = g1.x * dy;
          const float aG = g1.y * ex2_approx(fmaf(-u, u, -(v * v)));
          const bool ok = has && !done && (aG >= kMinRenderAlpha);
          if (!__any_sync(kFull, ok)) continue;
          const float* shp = pay + jj * (3 * CC);
          float y[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float s = 0.f;
            if constexpr (CC % 4 == 0) {
              const float4* p4 = reinterpret_cast<const float4*>(shp + c * CC);
              float se = 0.f, so = 0.f;  // even / odd k partial sums: one FFMA2 per coefficient pair
#pragma unroll
              for (int k = 0; k < CC / 4; ++k) {
                float4 q = p4[k];
                ffma2(se, so, q.x, q.y, Y[4 * k], Y[4 * k + 1]);
                ffma2(se, so, q.z, q.w, Y[4 * k + 2], Y[4 * k + 3]);
              }
              s = se + so;
            } else {
#pragma unroll
              for (int k = 0; k < CC; ++k) s = fmaf(shp[c * CC + k], Y[k], s);
            }
            y[c] = sigmoid_fast(s);
          }
          const float w = ok ? aG * T : 0.f;
          acc0 = fmaf(w, y[0], acc0); acc1 = fmaf(w, y[1], acc1); acc2 = fmaf(w, y[2], acc2);
          if (ok) {
            T = fmaf(-aG, T, T);           // T *= (1 - a*G)
            done = T < a.thresh;           // reference tests T < thresh before the NEXT Gaussian
            if (done) reach = b * B + jj + 1;
          }
        }
        if (__all_sync(kFull, done)) { warp_done = true; break; }
      }
    }
    if (has_next) {
      cp_async_wait<0>();
      if (use_bulk) mbar_wait(&s_bar[(b + 1) & 1], (uint32_t)(((b + 1) >> 1) & 1));
    }
    if (__syncthreads_and(warp_done ? 1 : 0)) break;
  }

  if (a.stats) {  // D_eff bookkeeping for the roofline report; not on the default path
    const int rr = inside ? (done ? reach : n) : 0;
    atomicMax(&s_reach, rr);
    __syncthreads();
    if (tid == 0) {
      atomicAdd(a.stats, (unsigned long long)s_reach);
      atomicAdd(a.stats + 1, (unsigned long long)staged);
    }
  }
  if (!inside) return;
  if (a.bg_rgb) {  // vol_render_bg.h:102-104
    acc0 = fmaf(a.bg_rgb[0], T, acc0); acc1 = fmaf(a.bg_rgb[1], T, acc1); acc2 = fmaf(a.bg_rgb[2], T, acc2);
  }
  a.out[3 * pix + 0] = acc0; a.out[3 * pix + 1] = acc1; a.out[3 * pix + 2] = acc2;
  if (a.T) a.T[pix] = T;
}

template <int C, int B>
static int launch_shh(const CompositeArgs& a, cudaStream_t st) {
  using L = StageLayout<PAY_SH, C, B, false>;
  const size_t smem = 2 * (size_t)L::kBytes;
  auto kern = k_composite_fwd_shh<C, B>;
  GSB_CUDA(ensure_max_dyn_smem(reinterpret_cast<const void*>(kern), (int)smem, a.device));
  dim3 grid(a.tiles_w, a.tiles_h, 1);
  kern<<<grid, kCtaThreads, smem, st>>>(a);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}

int launch_composite_fwd_shh(int C, const CompositeArgs& a, cudaStream_t st) {
  if (a.tiles_w <= 0 || a.tiles_h <= 0) return GSB200_OK;
  switch (C) {
    case 3: return launch_shh<3, GSB_FWDH_B>(a, st);
    case 4: return launch_shh<4, GSB_FWDH_B>(a, st);
    default: break;
  }
  set_error("composite_fwd_shh: unsupported C %d (half-warp kernel covers SH degree 2 and 3)", C);
  return GSB200_ERR_UNSUPPORTED;
}

}  // namespace gsb
