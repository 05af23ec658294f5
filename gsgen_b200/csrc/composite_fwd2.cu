// composite_fwd2.cu -- SH forward composite with TWO pixels per thread (reference kernels
// tile_based_vol_rendering_sh_entry<C> [_with_bg], vol_render_sh.h:171-248, vol_render_bg.h:12-110).
//
// Why (measured round 2, profiles/r2_ncu_full_c3_*): both SH composites are co-limited by instruction issue (~65-70 % of
// the slots) AND by the shared-memory data pipe (~66-71 % of its wavefronts): every (warp, Gaussian) hit re-reads the
// Gaussian's 3*C*C coefficients with broadcast LDS.128 (2 wavefronts each, 24 per hit at degree 3), 8 warps of a tile
// reading the same 192 bytes.  Here a warp owns an 8x8 pixel block and every lane two pixels (rows y and y+4 of the
// block): one coefficient read feeds two dot products, the box test / list walk / record fetch is shared, and the two
// independent per-pixel chains give the scheduler twice the instruction-level parallelism per warp.  128 threads per
// 16x16 tile.  Per-pixel semantics are those of composite_fwd.cu (same splat records, same thresholds, same
// early-termination rule); only the decomposition of the tile changes.
#include "composite_common.cuh"

namespace gsb {

#ifndef GSB_FWD2_B
#define GSB_FWD2_B 64
#endif
#ifndef GSB_FWD2_MINBLOCKS
#define GSB_FWD2_MINBLOCKS 6  // 128 threads: 6 CTAs/SM -> <= 80 registers
#endif

template <int C, int B>
__global__ void __launch_bounds__(128, GSB_FWD2_MINBLOCKS)
k_composite_fwd_sh2(const CompositeArgs a) {
  using L = StageLayout<PAY_SH, C, B, false>;
  using PT = PayTraits<PAY_SH, C>;
  constexpr int CC = PT::CC;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t s_bar[2];
  __shared__ int s_reach;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;  // 4 warps
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile = tile_y * a.tiles_w + tile_x;
  const float tlx = a.topleft_ptr ? a.topleft_ptr[0] : a.tlx;
  const float tly = a.topleft_ptr ? a.topleft_ptr[1] : a.tly;
  // warp block: 8 wide x 8 tall; lane -> column (lane & 7), rows (lane >> 3) and (lane >> 3) + 4
  const int bx0 = tile_x * kTile + (warp & 1) * 8, by0 = tile_y * kTile + (warp >> 1) * 8;
  const int gx = bx0 + (lane & 7);
  const int gy0 = by0 + (lane >> 3), gy1 = gy0 + 4;
  const bool in0 = (gx < a.W) && (gy0 < a.H), in1 = (gx < a.W) && (gy1 < a.H);
  const float px = fmaf((float)gx, a.psx, tlx);
  const float py0 = fmaf((float)gy0, a.psy, tly), py1 = fmaf((float)gy1, a.psy, tly);
  PixelGeom pg;  // only the block extent is used (splat_hits_block)
  pg.X0 = fmaf((float)bx0, a.psx, tlx);
  pg.X1 = fmaf((float)(bx0 + 7), a.psx, tlx);
  pg.Y0 = fmaf((float)by0, a.psy, tly);
  pg.Y1 = fmaf((float)(by0 + 7), a.psy, tly);
  const int pix0 = gy0 * a.W + gx, pix1 = gy1 * a.W + gx;

  const int s0 = a.start[tile];
  const int n = (s0 < 0) ? 0 : (a.end[tile] - s0);
  if (n <= 0) {  // empty tile: see composite_fwd.cu (A.9-15; sh_with_bg writes the background, vol_render_bg.h:34-53)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const bool in = h ? in1 : in0;
      const int pix = h ? pix1 : pix0;
      if (!in) continue;
      if (a.bg_rgb) {
        a.out[3 * pix + 0] = a.bg_rgb[0]; a.out[3 * pix + 1] = a.bg_rgb[1]; a.out[3 * pix + 2] = a.bg_rgb[2];
      } else if (a.write_empty) {
        a.out[3 * pix + 0] = 0.f; a.out[3 * pix + 1] = 0.f; a.out[3 * pix + 2] = 0.f;
      }
      if (a.write_empty && a.T) a.T[pix] = 1.0f;
    }
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

  float T0 = 1.0f, T1 = 1.0f;
  float r0 = 0.f, g0c = 0.f, b0 = 0.f, r1 = 0.f, g1c = 0.f, b1 = 0.f;
  bool done0 = !in0 || (1.0f < a.thresh), done1 = !in1 || (1.0f < a.thresh);
  int reach = 0, staged = 0;  // statistics only (a.stats)
  float Y0[CC], Y1[CC];
  {
    float c9[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) c9[k] = a.c9_ptr ? a.c9_ptr[k] : a.c9[k];
    float d[3];
    pixel_dir(px, py0, c9, d);
    sh_basis<C>(d[0], d[1], d[2], Y0);
    pixel_dir(px, py1, c9, d);
    sh_basis<C>(d[0], d[1], d[2], Y1);
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

  bool warp_done = __all_sync(kFull, done0 && done1);
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
      for (int r = 0; r * 32 < cnt; ++r) {
        const int j = r * 32 + lane;
        bool hit = false;
        if (j < cnt) hit = splat_hits_block(sg0[j], sg1[j], pg);
        unsigned m = __ballot_sync(kFull, hit);
        while (m) {
          const int jj = r * 32 + (__ffs(m) - 1);
          m &= m - 1;
          const float4 q0 = sg0[jj], q1 = sg1[jj];
          // a*G for both pixels: G = exp2(-(u^2+v^2)), u = p0*dx + p1*dy, v = p2*dy (make_splat)
          const float dx = px - q0.x;
          const float dy0 = py0 - q0.y, dy1 = py1 - q0.y;
          const float ux = q0.z * dx;
          const float u0 = fmaf(q0.w, dy0, ux), u1 = fmaf(q0.w, dy1, ux);
          const float v0 = q1.x * dy0, v1 = q1.x * dy1;
          const float aG0 = q1.y * ex2_approx(fmaf(-u0, u0, -(v0 * v0)));
          const float aG1 = q1.y * ex2_approx(fmaf(-u1, u1, -(v1 * v1)));
          const bool ok0 = !done0 && (aG0 >= kMinRenderAlpha), ok1 = !done1 && (aG1 >= kMinRenderAlpha);
          if (!__any_sync(kFull, ok0 || ok1)) continue;
          const float* shp = reinterpret_cast<const float*>(st + L::kPay) + jj * (3 * CC);
          float ya[3], yb[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float sa, sb;
            if constexpr (CC % 4 == 0) {
              const float4* p4 = reinterpret_cast<const float4*>(shp + c * CC);
              float ae = 0.f, ao = 0.f, be = 0.f, bo = 0.f;  // even / odd k partial sums per pixel (FFMA2)
#pragma unroll
              for (int k = 0; k < CC / 4; ++k) {
                const float4 q = p4[k];  // ONE coefficient read, two pixels
                ffma2(ae, ao, q.x, q.y, Y0[4 * k], Y0[4 * k + 1]);
                ffma2(be, bo, q.x, q.y, Y1[4 * k], Y1[4 * k + 1]);
                ffma2(ae, ao, q.z, q.w, Y0[4 * k + 2], Y0[4 * k + 3]);
                ffma2(be, bo, q.z, q.w, Y1[4 * k + 2], Y1[4 * k + 3]);
              }
              sa = ae + ao; sb = be + bo;
            } else {
              sa = 0.f; sb = 0.f;
#pragma unroll
              for (int k = 0; k < CC; ++k) {
                const float q = shp[c * CC + k];
                sa = fmaf(q, Y0[k], sa); sb = fmaf(q, Y1[k], sb);
              }
            }
            ya[c] = sigmoid_fast(sa); yb[c] = sigmoid_fast(sb);
          }
          const float w0 = ok0 ? aG0 * T0 : 0.f, w1 = ok1 ? aG1 * T1 : 0.f;
          r0 = fmaf(w0, ya[0], r0); g0c = fmaf(w0, ya[1], g0c); b0 = fmaf(w0, ya[2], b0);
          r1 = fmaf(w1, yb[0], r1); g1c = fmaf(w1, yb[1], g1c); b1 = fmaf(w1, yb[2], b1);
          if (ok0) {
            T0 = fmaf(-aG0, T0, T0);  // T *= (1 - a*G); the reference tests T < thresh before the NEXT Gaussian
            done0 = T0 < a.thresh;
            if (done0) reach = max(reach, b * B + jj + 1);
          }
          if (ok1) {
            T1 = fmaf(-aG1, T1, T1);
            done1 = T1 < a.thresh;
            if (done1) reach = max(reach, b * B + jj + 1);
          }
        }
        if (__all_sync(kFull, done0 && done1)) { warp_done = true; break; }
      }
    }
    if (has_next) {
      cp_async_wait<0>();
      if (use_bulk) mbar_wait(&s_bar[(b + 1) & 1], (uint32_t)(((b + 1) >> 1) & 1));
    }
    if (__syncthreads_and(warp_done ? 1 : 0)) break;
  }

  if (a.stats) {  // D_eff bookkeeping: 1 + last list index any pixel of the tile needed
    int rr = 0;
    if (in0) rr = max(rr, done0 ? reach : n);
    if (in1) rr = max(rr, done1 ? reach : n);
    // (reach holds the later of the two pixels' stopping points; a pixel that never stopped needs the whole list)
    atomicMax(&s_reach, rr);
    __syncthreads();
    if (tid == 0) {
      atomicAdd(a.stats, (unsigned long long)s_reach);
      atomicAdd(a.stats + 1, (unsigned long long)staged);
    }
  }
  if (in0) {
    if (a.bg_rgb) { r0 = fmaf(a.bg_rgb[0], T0, r0); g0c = fmaf(a.bg_rgb[1], T0, g0c); b0 = fmaf(a.bg_rgb[2], T0, b0); }
    a.out[3 * pix0 + 0] = r0; a.out[3 * pix0 + 1] = g0c; a.out[3 * pix0 + 2] = b0;
    if (a.T) a.T[pix0] = T0;
  }
  if (in1) {
    if (a.bg_rgb) { r1 = fmaf(a.bg_rgb[0], T1, r1); g1c = fmaf(a.bg_rgb[1], T1, g1c); b1 = fmaf(a.bg_rgb[2], T1, b1); }
    a.out[3 * pix1 + 0] = r1; a.out[3 * pix1 + 1] = g1c; a.out[3 * pix1 + 2] = b1;
    if (a.T) a.T[pix1] = T1;
  }
}

template <int C, int B>
static int launch_sh2(const CompositeArgs& a, cudaStream_t st) {
  using L = StageLayout<PAY_SH, C, B, false>;
  static_assert(B <= 128, "one thread stages one list entry");
  const size_t smem = 2 * (size_t)L::kBytes;
  auto kern = k_composite_fwd_sh2<C, B>;
  GSB_CUDA(ensure_max_dyn_smem(reinterpret_cast<const void*>(kern), (int)smem, a.device));
  dim3 grid(a.tiles_w, a.tiles_h, 1);
  kern<<<grid, 128, smem, st>>>(a);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}

int launch_composite_fwd_sh2(int C, const CompositeArgs& a, cudaStream_t st) {
  if (a.tiles_w <= 0 || a.tiles_h <= 0) return GSB200_OK;
  switch (C) {
    case 3: return launch_sh2<3, GSB_FWD2_B>(a, st);
    case 4: return launch_sh2<4, GSB_FWD2_B>(a, st);
    default: break;
  }
  set_error("composite_fwd_sh2: unsupported C %d (two-pixel kernel covers SH degree 2 and 3)", C);
  return GSB200_ERR_UNSUPPORTED;
}

}  // namespace gsb
