// composite_fwd.cu -- per-tile front-to-back alpha compositing, forward.
//
// Reference kernels replaced (all share the same per-pixel loop, App. A.6):
//   tile_based_vol_rendering_with_T / _entry_start_end   vol_render.h:994-1062, :782-847     (RGB [+T])
//   tile_based_vol_rendering_scalar                       vol_render_scalar.h:47-102          (scalar)
//   tile_based_vol_rendering_sh_entry<C> [_with_bg]       vol_render_sh.h:171-248, vol_render_bg.h:12-110
// plus, in the fused RGB mode, the three render_scalar passes of render_one (depth, opacity, depth^2;
// gs/gaussian_splatting.py:1337-1401) and the `out + T*bg` torch op (gs/renderer.py:1182) folded into
// the same walk of the blend list.
//
// Per (pixel, Gaussian): 2 FADD + 4 FMUL/FFMA + 1 MUFU.EX2 + 1 FMUL + compare; blended pairs add
// 1 FMUL + (3..6) FFMA (RGB) or 3*C*C FFMA + 3 sigmoids (SH).  No division, no fp64: the conic is
// pre-factored once per Gaussian (make_splat).
#include "composite_common.cuh"

namespace gsb {

#ifndef GSB_FWD_B
#define GSB_FWD_B 64        // list entries per staged batch for SH degree >= 2
#endif
#ifndef GSB_FWD_MINBLOCKS
#define GSB_FWD_MINBLOCKS 0  // 0: no minBlocksPerSM hint (48 registers at SH deg 3 -> 5 CTAs/SM).  Note: a hint of 1 is
#endif                       // NOT neutral -- ptxas then spends 81 registers (3 CTAs/SM); 6 caps it at 40.
#if GSB_FWD_MINBLOCKS > 0
#define GSB_FWD_BOUNDS __launch_bounds__(kCtaThreads, GSB_FWD_MINBLOCKS)
#else
#define GSB_FWD_BOUNDS __launch_bounds__(kCtaThreads)
#endif

template <int PAY, int C, bool EXTRAS, int B>
__global__ void GSB_FWD_BOUNDS
k_composite_fwd(const CompositeArgs a) {
  using L = StageLayout<PAY, C, B, false>;
  using PT = PayTraits<PAY, C>;
  constexpr int CC = PT::CC;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t s_bar[2];
  __shared__ int s_reach;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile = tile_y * a.tiles_w + tile_x;
  const PixelGeom pg = pixel_geom(a, tile_x, tile_y, warp, lane);
  const int pix = pg.gy * a.W + pg.gx;

  const int s0 = a.start[tile];
  const int n = (s0 < 0) ? 0 : (a.end[tile] - s0);
  if (n <= 0) {
    // empty tile: the reference leaves the caller-initialised outputs untouched (A.9-15) except
    // sh_with_bg, which writes the background (vol_render_bg.h:34-53)
    if (!pg.inside) return;
    if constexpr (PAY == PAY_SH) {
      if (a.bg_rgb) {
        a.out[3 * pix + 0] = a.bg_rgb[0]; a.out[3 * pix + 1] = a.bg_rgb[1]; a.out[3 * pix + 2] = a.bg_rgb[2];
      } else if (a.write_empty) {
        a.out[3 * pix + 0] = 0.f; a.out[3 * pix + 1] = 0.f; a.out[3 * pix + 2] = 0.f;
      }
      if (a.write_empty && a.T) a.T[pix] = 1.0f;
    } else if (a.write_empty) {
      if constexpr (PAY == PAY_RGB) {
        float b0 = 0.f, b1 = 0.f, b2 = 0.f;
        if (a.bg) { b0 = a.bg[3 * pix]; b1 = a.bg[3 * pix + 1]; b2 = a.bg[3 * pix + 2]; }
        a.out[3 * pix + 0] = b0; a.out[3 * pix + 1] = b1; a.out[3 * pix + 2] = b2;
        if constexpr (EXTRAS) { a.depth[pix] = 0.f; a.opacity[pix] = 0.f; a.z2[pix] = 0.f; }
      } else {
        a.out[pix] = 0.f;
      }
      if (a.T) a.T[pix] = 1.0f;
    }
    return;
  }

  const bool use_bulk = PT::kBulkOk && ((reinterpret_cast<uintptr_t>(a.sh) & 15) == 0);
  if (PAY == PAY_SH && tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    fence_mbar_init();
  }
  if (tid == 0) s_reach = 0;
  __syncthreads();

  // per-pixel registers
  float T = 1.0f;
  int reach = 0, staged = 0;  // statistics only (a.stats)
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;          // rgb / scalar in acc0
  float accD = 0.f, accO = 0.f, accZ = 0.f;          // extras
  bool done = !pg.inside || (1.0f < a.thresh);
  float Y[(PAY == PAY_SH) ? CC : 1];
  if constexpr (PAY == PAY_SH) {
    float c9[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) c9[k] = a.c9_ptr ? a.c9_ptr[k] : a.c9[k];
    float d[3];
    pixel_dir(pg.px, pg.py, c9, d);
    sh_basis<C>(d[0], d[1], d[2], Y);
  }

  const int nb = (n + B - 1) / B;
  const int32_t* ids = a.ids + s0;
  // prologue: stage batch 0, prefetch ids of batch 1
  {
    int cnt0 = min(B, n);
    int id0 = (tid < cnt0) ? ids[tid] : 0;
    if (PAY == PAY_SH && use_bulk && tid == 0) mbar_arrive_expect_tx(&s_bar[0], (uint32_t)cnt0 * 3 * CC * 4);
    if (tid < B) stage_entry<PAY, C, B, false>(a, smem, tid, id0, tid < cnt0, use_bulk, &s_bar[0]);
    else cp_async_commit();
    staged = cnt0;
  }
  int id_next = 0;
  if (nb > 1) { int j = B + tid; id_next = (tid < B && j < n) ? ids[j] : 0; }

  bool warp_done = __all_sync(kFull, done);
  // batch 0 has landed for everyone
  cp_async_wait<0>();
  if (PAY == PAY_SH && use_bulk) mbar_wait(&s_bar[0], 0u);
  __syncthreads();
  // ONE block barrier per batch: the copies of batch b+1 are issued before batch b is composited and waited
  // for right before the barrier that also retires batch b's buffer and votes on early termination.
  for (int b = 0; b < nb; ++b) {
    unsigned char* st = smem + (b & 1) * L::kBytes;
    const int cnt = min(B, n - b * B);
    const bool has_next = (b + 1 < nb);
    if (has_next) {
      const int cntn = min(B, n - (b + 1) * B);
      uint64_t* barn = &s_bar[(b + 1) & 1];
      if (PAY == PAY_SH && use_bulk && tid == 0) mbar_arrive_expect_tx(barn, (uint32_t)cntn * 3 * CC * 4);
      if (tid < B) stage_entry<PAY, C, B, false>(a, smem + ((b + 1) & 1) * L::kBytes, tid, id_next, tid < cntn,
                                                 use_bulk, barn);
      else cp_async_commit();
      staged += cntn;
      if (b + 2 < nb) { int j = (b + 2) * B + tid; id_next = (tid < B && j < n) ? ids[j] : 0; }
    }

    if (!warp_done) {
      const float4* sg0 = reinterpret_cast<const float4*>(st + L::kG0);
      const float4* sg1 = reinterpret_cast<const float4*>(st + L::kG1);
      for (int r = 0; r * 32 < cnt; ++r) {
        const int j = r * 32 + lane;
        bool hit = false;
        if (j < cnt) hit = splat_hits_block(sg0[j], sg1[j], pg);
        unsigned m = __ballot_sync(kFull, hit);
        // (prefetching the next hit's record here was measured in round 1: +16 registers -> 4 instead of 5 CTAs/SM,
        // 0.388 -> 0.423 ms at C3; the backward, which is register-capped anyway, keeps the prefetch)
        while (m) {
          const int jj = r * 32 + (__ffs(m) - 1);
          m &= m - 1;
          const float4 g0 = sg0[jj], g1 = sg1[jj];
          float G, u, v;
          const float aG = splat_aG(g0, g1, pg.px, pg.py, &G, &u, &v);
          const bool ok = !done && (aG >= kMinRenderAlpha);
          if constexpr (PAY == PAY_SH) {
            if (!__any_sync(kFull, ok)) continue;
            const float* shp = reinterpret_cast<const float*>(st + L::kPay) + jj * (3 * CC);
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
          } else {
            const float4 p = reinterpret_cast<const float4*>(st + L::kPay)[jj];
            const float w = ok ? aG * T : 0.f;
            acc0 = fmaf(w, p.x, acc0);
            if constexpr (PAY == PAY_RGB) {
              acc1 = fmaf(w, p.y, acc1); acc2 = fmaf(w, p.z, acc2);
              if constexpr (EXTRAS) {
                const float wd = w * p.w;
                accD += wd; accZ = fmaf(wd, p.w, accZ); accO += w;
              }
            }
          }
          if (ok) {
            T = fmaf(-aG, T, T);           // T *= (1 - a*G)
            done = T < a.thresh;           // reference tests T < thresh before the NEXT Gaussian
            if (done) reach = b * B + jj + 1;
          }
        }
        if (__all_sync(kFull, done)) { warp_done = true; break; }
      }
    }
    if (has_next) {  // batch b+1 must have landed (also drains the copies before the CTA may retire)
      cp_async_wait<0>();
      if (PAY == PAY_SH && use_bulk) mbar_wait(&s_bar[(b + 1) & 1], (uint32_t)(((b + 1) >> 1) & 1));
    }
    if (__syncthreads_and(warp_done ? 1 : 0)) break;
  }

  if (a.stats) {  // D_eff bookkeeping for the roofline report; not on the default path
    const int r = pg.inside ? (done ? reach : n) : 0;
    atomicMax(&s_reach, r);
    __syncthreads();
    if (tid == 0) {
      atomicAdd(a.stats, (unsigned long long)s_reach);
      atomicAdd(a.stats + 1, (unsigned long long)staged);
    }
  }
  if (!pg.inside) return;
  if constexpr (PAY == PAY_SH) {
    if (a.bg_rgb) {  // vol_render_bg.h:102-104
      acc0 = fmaf(a.bg_rgb[0], T, acc0); acc1 = fmaf(a.bg_rgb[1], T, acc1); acc2 = fmaf(a.bg_rgb[2], T, acc2);
    }
    a.out[3 * pix + 0] = acc0; a.out[3 * pix + 1] = acc1; a.out[3 * pix + 2] = acc2;
  } else if constexpr (PAY == PAY_RGB) {
    if (a.bg) {      // gs/renderer.py:1182  out + T*bg
      acc0 = fmaf(a.bg[3 * pix], T, acc0); acc1 = fmaf(a.bg[3 * pix + 1], T, acc1);
      acc2 = fmaf(a.bg[3 * pix + 2], T, acc2);
    }
    a.out[3 * pix + 0] = acc0; a.out[3 * pix + 1] = acc1; a.out[3 * pix + 2] = acc2;
    if constexpr (EXTRAS) { a.depth[pix] = accD; a.opacity[pix] = accO; a.z2[pix] = accZ; }
  } else {
    a.out[pix] = acc0;
  }
  if (a.T) a.T[pix] = T;
}

template <int PAY, int C, bool EXTRAS, int B>
static int launch_one(const CompositeArgs& a, cudaStream_t st) {
  using L = StageLayout<PAY, C, B, false>;
  const size_t smem = 2 * (size_t)L::kBytes;
  auto kern = k_composite_fwd<PAY, C, EXTRAS, B>;
  GSB_CUDA(ensure_max_dyn_smem(reinterpret_cast<const void*>(kern), (int)smem, a.device));
  dim3 grid(a.tiles_w, a.tiles_h, 1);
  kern<<<grid, kCtaThreads, smem, st>>>(a);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}

int launch_composite_fwd(int pay_kind, int C, bool extras, const CompositeArgs& a, cudaStream_t st) {
  if (a.tiles_w <= 0 || a.tiles_h <= 0) return GSB200_OK;
  switch (pay_kind) {
    case PAY_RGB:
      return extras ? launch_one<PAY_RGB, 1, true, 256>(a, st) : launch_one<PAY_RGB, 1, false, 256>(a, st);
    case PAY_SCALAR:
      return launch_one<PAY_SCALAR, 1, false, 256>(a, st);
    case PAY_SH:
      switch (C) {
        case 1: return launch_one<PAY_SH, 1, false, 256>(a, st);
        case 2: return launch_one<PAY_SH, 2, false, 256>(a, st);
        case 3: return launch_one<PAY_SH, 3, false, GSB_FWD_B>(a, st);
        case 4: return launch_one<PAY_SH, 4, false, GSB_FWD_B>(a, st);
        default: break;
      }
    default: break;
  }
  set_error("composite_fwd: unsupported payload kind %d / C %d (reference dispatches C in 1..4, render.cu:507-545)",
            pay_kind, C);
  return GSB200_ERR_UNSUPPORTED;
}

}  // namespace gsb
