// composite_bwd.cu -- per-tile alpha compositing, backward.
//
// Reference kernels replaced (front-to-back recompute + per-pair gradients, App. A.6):
//   tile_based_vol_rendering_backward_entry_start_end   vol_render.h:866-973, inner :318-418      (RGB)
//   tile_based_vol_rendering_scalar_backward             vol_render_scalar.h:148-234, inner :104-146
//   tile_based_vol_rendering_backward_sh_entry<C>[_bg]   vol_render_sh.h:353-455 (inner :268-351), vol_render_bg.h:131-242
// and, in the fused RGB mode, the three scalar backward passes (depth, opacity, depth^2 incl. the chain
// rule of z2 = depth*depth, gs/gaussian_splatting.py:1385) in the same walk.
//
// Gradient of one blended pair (same algebra as the reference, re-associated so that only ONE running
// scalar is carried per pixel):
//   gc   = sum_ch go_ch * pay_ch                         S   = sum_ch go_ch * (F_ch - Cacc_ch)   (S -= w*gc)
//   pAG  = T*gc - S/(1 - aG)          (reference: sum_ch go_ch*(pay_ch*T - (F_ch-Cacc_ch)/(1-aG)), vol_render.h:395-398)
//   g_alpha += pAG*G ; gG = pAG*a*G ; v = S^-1 d ; g_mean2d += gG*v ; g_cov2d += 0.5*gG*v v^T
//   g_pay_ch += w*go_ch ;  g_sh[c,k] += w*y_c(1-y_c)*go_c*Y_k
//
// Reduction over the pixels of a tile.  The reference issues 10..55 shared-memory float atomics per (pixel,
// Gaussian), all 256 pixels hitting the same address.  Here a warp (32 pixels) reduces first, in one of two ways:
//   * RGB / scalar / SH degree 0 (<= 16 values): a halving butterfly (K + K/2 + ... shuffles for K values);
//   * SH degree >= 1 (3*C^2 + 6 values): measured in round 1, the butterfly was 3/4 of this kernel's instructions
//     (profiles/r1_ncu_full_c3_run2.json).  g_sh[c,k] = sum_p t_c(p) * Y_k(p) is a tiny matrix product, so each
//     lane stores only its 3 scalars t_c (+6 geometry values) per hit into a per-warp transpose buffer and every 4
//     hits the warp multiplies [12 x 32] by the block's [32 x C^2] basis with the basis column held in registers
//     (32 FFMA + 8 broadcast LDS.128 per output row) -- ~80 instructions per hit instead of ~370.
// One lane per value then adds into the per-batch shared accumulator and each (Gaussian, tile) instance is flushed
// to HBM once per batch with vector reductions (red.global.add.v4.f32).
#include "composite_common.cuh"

namespace gsb {

template <int PAY, int C, bool EXTRAS> struct BwdTraits {
  static constexpr int CC = C * C;
  static constexpr int kPayVals = (PAY == PAY_SH) ? 3 * CC : (PAY == PAY_RGB ? (EXTRAS ? 4 : 3) : 1);
  static constexpr int kVals = 6 + kPayVals;  // gmx gmy gxx gxy gyy galpha | payload grads
  static constexpr int kK = kVals <= 8 ? 8 : (kVals <= 16 ? 16 : (kVals <= 32 ? 32 : 64));
  static constexpr int kStride = (kVals + 3) / 4 * 4;  // floats per accumulator row (16 B aligned rows)
  // transpose-buffer reduction (SH with C >= 2)
  static constexpr bool kTbuf = (PAY == PAY_SH) && (CC >= 4);
  static constexpr int kG = 4;                                   // hits per flush
  static constexpr int kKL = CC <= 4 ? 4 : 16;                   // lanes along k (power of two >= CC)
  static constexpr int kJ = 32 / kKL;                            // row groups
  static constexpr int kYsmFloats = kTbuf ? kKL * 32 : 0;        // per warp: basis matrix [k][32 px], float4 groups rotated by k
  static constexpr int kTbufFloats = kTbuf ? kG * 9 * 32 + kYsmFloats : 0;  // per warp
  static constexpr int kNR = (3 * kG + kJ - 1) / kJ;             // SH rows per lane group in a flush
  // accumulator row layout: butterfly path [6 geometry | payload]; transpose-buffer path [3*CC sh | 6 geometry]
  // (keeps the SH part 16 B aligned so the batch flush streams it float4 by float4)
  static constexpr int kGeoOff = kTbuf ? kPayVals : 0;
  static constexpr int kPayOff = kTbuf ? 0 : 6;
};

#ifndef GSB_BWD_FLUSH_INLINE
#define GSB_BWD_FLUSH_INLINE 0  // 0: the flush is a real call (keeps its 20+ registers out of the hit loop's budget)
#endif
#if GSB_BWD_FLUSH_INLINE
#define GSB_FLUSH_ATTR __device__ __forceinline__
#else
#define GSB_FLUSH_ATTR __device__ __noinline__
#endif

// Flush of a warp's transpose buffer: SH rows [hit*3+c][32 px] x basis Ysm[k][px] -> g_sh partial sums; geometry rows
// -> plain row sums; both added into the per-batch shared accumulator rows of the hits' list entries (`slots`).
template <int PAY, int C, bool EXTRAS>
GSB_FLUSH_ATTR void flush_tbuf_fn(const float* my_t, const float* my_y, float* s_acc, int nslot, unsigned slots,
                                  int lane) {
  using BT = BwdTraits<PAY, C, EXTRAS>;
  constexpr int CC = C * C;
  constexpr int STR = BT::kStride;
  const int tk = lane & (BT::kKL - 1);
  const int tj = lane / BT::kKL;
  __syncwarp();
  float acc[BT::kNR], aco[BT::kNR];  // even / odd pixel partial sums (FFMA2)
#pragma unroll
  for (int i = 0; i < BT::kNR; ++i) { acc[i] = 0.f; aco[i] = 0.f; }
  const float* trow = my_t + tj * 32;
  const float* yrow = my_y + tk * 32;
#pragma unroll 2
  for (int q = 0; q < 8; ++q) {
    const float4 y4 = *reinterpret_cast<const float4*>(yrow + 4 * ((q + tk) & 7));
#pragma unroll
    for (int i = 0; i < BT::kNR; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(trow + i * (BT::kJ * 32) + 4 * q);
      ffma2(acc[i], aco[i], t.x, t.y, y4.x, y4.y);
      ffma2(acc[i], aco[i], t.z, t.w, y4.z, y4.w);
    }
  }
#pragma unroll
  for (int i = 0; i < BT::kNR; ++i) {
    const int row = tj + i * BT::kJ;
    if (row < nslot * 3 && tk < CC) {
      const int h = row / 3, c = row - 3 * h;
      const int jj = (slots >> (8 * h)) & 255u;
      atomicAdd(s_acc + jj * STR + BT::kPayOff + c * CC + tk, acc[i] + aco[i]);
    }
  }
  // geometry: plain row sums, one row per lane (rotated 16-byte reads: conflict-free)
  if (lane < nslot * 6) {
    const float* base = my_t + (BT::kG * 3 + lane) * 32;
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = *reinterpret_cast<const float4*>(base + 4 * ((q + lane) & 7));
      s += (t.x + t.y) + (t.z + t.w);
    }
    const int h = lane / 6, v = lane - 6 * h;
    const int jj = (slots >> (8 * h)) & 255u;
    atomicAdd(s_acc + jj * STR + BT::kGeoOff + v, s);
  }
  __syncwarp();
}

#ifndef GSB_BWD_B
#define GSB_BWD_B 32  // list entries per staged batch for SH degree >= 2 (shared memory: 3 CTAs/SM at 32)
#endif
#ifndef GSB_BWD_MINBLOCKS
#define GSB_BWD_MINBLOCKS 3  // CTAs/SM the register allocation is capped for (3 -> 80 registers, 56 B of spills at SH deg 3)
#endif

template <int PAY, int C, bool EXTRAS, bool FUSED, int B>
__global__ void __launch_bounds__(kCtaThreads, GSB_BWD_MINBLOCKS)
k_composite_bwd(const CompositeArgs a) {
  using L = StageLayout<PAY, C, B, true>;
  using PT = PayTraits<PAY, C>;
  using BT = BwdTraits<PAY, C, EXTRAS>;
  constexpr int CC = PT::CC;
  constexpr int K = BT::kK;
  constexpr int NV = BT::kVals;
  constexpr int STR = BT::kStride;
  constexpr bool TBUF = BT::kTbuf;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t s_bar[2];
  __shared__ unsigned s_touched[2][B / 32];

  float* s_acc = reinterpret_cast<float*>(smem + 2 * L::kBytes);  // [B][STR]
  float* s_tbuf = s_acc + B * STR;                                // [8 warps][G][9][32] (TBUF only)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile = tile_y * a.tiles_w + tile_x;
  const PixelGeom pg = pixel_geom(a, tile_x, tile_y, warp, lane);
  const int pix = pg.gy * a.W + pg.gx;

  const int s0 = a.start[tile];
  const int n = (s0 < 0) ? 0 : (a.end[tile] - s0);
  if (n <= 0) {
    if (PAY == PAY_RGB && a.g_bg && pg.inside) {  // T = 1 on empty tiles
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v = a.gout ? a.gout[3 * pix + c] : 0.f;
        a.g_bg[3 * pix + c] = (v != v) ? 0.f : fminf(fmaxf(v, -3.4028235e38f), 3.4028235e38f);
      }
    }
    return;
  }

  const bool use_bulk = PT::kBulkOk && ((reinterpret_cast<uintptr_t>(a.sh) & 15) == 0);
  if (PAY == PAY_SH && tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < B * STR; i += kCtaThreads) s_acc[i] = 0.f;
  if (tid < B / 32) { s_touched[0][tid] = 0u; s_touched[1][tid] = 0u; }
  __syncthreads();

  // per-pixel registers
  float T = 1.0f;
  bool done = !pg.inside || (1.0f < a.thresh);
  float go0 = 0.f, go1 = 0.f, go2 = 0.f, goD = 0.f, goO = 0.f, goZ = 0.f;
  float S = 0.f;  // sum_ch go_ch * (F_ch - Cacc_ch)
  if (pg.inside) {
    if constexpr (PAY == PAY_SCALAR) {
      go0 = a.gout ? a.gout[pix] : 0.f;
      S = go0 * a.fin[pix];
    } else {
      if (a.gout) { go0 = a.gout[3 * pix]; go1 = a.gout[3 * pix + 1]; go2 = a.gout[3 * pix + 2]; }
      S = go0 * a.fin[3 * pix] + go1 * a.fin[3 * pix + 1] + go2 * a.fin[3 * pix + 2];
      if constexpr (EXTRAS) {
        if (a.g_depth) { goD = a.g_depth[pix]; S = fmaf(goD, a.fin_depth[pix], S); }
        if (a.g_opacity) { goO = a.g_opacity[pix]; S = fmaf(goO, a.fin_opacity[pix], S); }
        if (a.g_z2) { goZ = a.g_z2[pix]; S = fmaf(goZ, a.fin_z2[pix], S); }
      }
    }
  }
  float Y[(PAY == PAY_SH) ? CC : 1];
  if constexpr (PAY == PAY_SH) {
    float c9[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) c9[k] = a.c9_ptr ? a.c9_ptr[k] : a.c9[k];
    float d[3];
    pixel_dir(pg.px, pg.py, c9, d);
    sh_basis<C>(d[0], d[1], d[2], Y);
  }
  // ---- TBUF: per-warp transpose buffer.  Layout (floats): SH rows [G*3][32] (row = hit*3 + channel), geometry rows
  // [G*6][32], then the block's basis matrix Ysm[k][32] (k-major, 16-byte groups rotated by k against bank conflicts)
  float* my_t = s_tbuf + warp * BT::kTbufFloats;
  float* my_y = my_t + BT::kG * 9 * 32;
  if constexpr (TBUF) {
#pragma unroll
    for (int k = 0; k < BT::kKL; ++k)  // pixel `lane` sits in float4 group lane/4, rotated by k (bank-conflict free)
      my_y[k * 32 + (((lane >> 2) + k) & 7) * 4 + (lane & 3)] = (k < CC) ? Y[k < CC ? k : 0] : 0.f;
    __syncwarp();
  }
  int nslot = 0;            // hits buffered in my_t (warp-uniform)
  unsigned slots = 0u;      // their batch entry indices, 8 bits each

  bool red_writer = false;
  const int red_e = TBUF ? 0 : red_index<K>(lane, &red_writer);

  // flush of the transpose buffer: rows [h][0..2] x basis -> SH gradients, rows [h][3..8] row sums -> geometry
  auto flush_tbuf = [&]() {
    if constexpr (TBUF) {
      flush_tbuf_fn<PAY, C, EXTRAS>(my_t, my_y, s_acc, nslot, slots, lane);
      nslot = 0;
      slots = 0u;
    }
  };

  const int nb = (n + B - 1) / B;
  const int32_t* ids = a.ids + s0;
  {
    int cnt0 = min(B, n);
    int id0 = (tid < cnt0) ? ids[tid] : 0;
    if (PAY == PAY_SH && use_bulk && tid == 0) mbar_arrive_expect_tx(&s_bar[0], (uint32_t)cnt0 * 3 * CC * 4);
    if (tid < B) stage_entry<PAY, C, B, true>(a, smem, tid, id0, tid < cnt0, use_bulk, &s_bar[0]);
    else cp_async_commit();
  }
  int id_next = 0;
  if (nb > 1) { int j = B + tid; id_next = (tid < B && j < n) ? ids[j] : 0; }

  bool warp_done = __all_sync(kFull, done);
  cp_async_wait<0>();
  if (PAY == PAY_SH && use_bulk) mbar_wait(&s_bar[0], 0u);
  __syncthreads();
  // two block barriers per batch: (B) accumulators complete -> flush; (C) batch retired + next batch landed + vote
  for (int b = 0; b < nb; ++b) {
    unsigned char* st = smem + (b & 1) * L::kBytes;
    const int cnt = min(B, n - b * B);
    const bool has_next = (b + 1 < nb);
    unsigned* touched_now = s_touched[b & 1];
    if (has_next) {
      const int cntn = min(B, n - (b + 1) * B);
      uint64_t* barn = &s_bar[(b + 1) & 1];
      if (PAY == PAY_SH && use_bulk && tid == 0) mbar_arrive_expect_tx(barn, (uint32_t)cntn * 3 * CC * 4);
      if (tid < B) stage_entry<PAY, C, B, true>(a, smem + ((b + 1) & 1) * L::kBytes, tid, id_next, tid < cntn,
                                                use_bulk, barn);
      else cp_async_commit();
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
        unsigned touched = 0u;
        // software pipeline over the hits: the next hit's record is fetched while the current one is evaluated
        int bitn = __ffs(m) - 1;
        float4 n0 = sg0[m ? r * 32 + bitn : 0], n1 = sg1[m ? r * 32 + bitn : 0];
        while (m) {
          const int bit = bitn;
          const int jj = r * 32 + bit;
          const float4 g0 = n0, g1 = n1;
          m &= m - 1;
          if (m) { bitn = __ffs(m) - 1; n0 = sg0[r * 32 + bitn]; n1 = sg1[r * 32 + bitn]; }
          float G, u, v;
          const float aG = splat_aG(g0, g1, pg.px, pg.py, &G, &u, &v);
          const bool ok = !done && (aG >= kMinRenderAlpha);
          if (!__any_sync(kFull, ok)) continue;
          touched |= 1u << bit;

          float vals[TBUF ? 1 : K];
          if constexpr (!TBUF) {
#pragma unroll
            for (int k = 0; k < K; ++k) vals[k] = 0.f;
          }
          const float w = ok ? aG * T : 0.f;
          float gc;  // sum_ch go_ch * pay_ch
          float t0 = 0.f, t1 = 0.f, t2 = 0.f;
          if constexpr (PAY == PAY_SH) {
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
            gc = go0 * y[0] + go1 * y[1] + go2 * y[2];
            t0 = w * (y[0] * (1.0f - y[0])) * go0;  // vol_render_sh.h:328-333
            t1 = w * (y[1] * (1.0f - y[1])) * go1;
            t2 = w * (y[2] * (1.0f - y[2])) * go2;
            if constexpr (!TBUF) {
#pragma unroll
              for (int k = 0; k < CC; ++k) {
                vals[6 + k] = t0 * Y[k];
                vals[6 + CC + k] = t1 * Y[k];
                vals[6 + 2 * CC + k] = t2 * Y[k];
              }
            }
          } else {
            const float4 p = reinterpret_cast<const float4*>(st + L::kPay)[jj];
            if constexpr (PAY == PAY_RGB) {
              gc = go0 * p.x + go1 * p.y + go2 * p.z;
              vals[6] = w * go0; vals[7] = w * go1; vals[8] = w * go2;
              if constexpr (EXTRAS) {
                gc = fmaf(goD, p.w, gc);
                gc += goO;
                gc = fmaf(goZ * p.w, p.w, gc);
                vals[9] = w * fmaf(2.0f * p.w, goZ, goD);  // d/d depth of (w*depth, w*depth^2)
              }
            } else {
              gc = go0 * p.x;
              vals[6] = w * go0;
            }
          }
          // pair gradient (masked by ok through w / okf)
          const float okf = ok ? 1.0f : 0.f;
          S = fmaf(-w, gc, S);
          const float rinv = rcp_approx(1.0f - aG);
          const float pAG = okf * fmaf(T, gc, -S * rinv);
          const float gG = pAG * aG;
          const float vx = kInvCholScale2 * g0.z * u;                        // (S^-1 d).x
          const float vy = kInvCholScale2 * fmaf(g0.w, u, g1.x * v);         // (S^-1 d).y
          const float hg = 0.5f * gG;
          const float e0 = gG * vx, e1 = gG * vy, e2 = hg * vx * vx, e3 = hg * vx * vy, e4 = hg * vy * vy;
          const float e5 = pAG * G;  // g_alpha (no clamp gate, vol_render.h:409)
          if (ok) {
            T = fmaf(-aG, T, T);
            done = T < a.thresh;
          }
          if constexpr (TBUF) {
            float* rs = my_t + nslot * (3 * 32) + lane;                    // SH rows of this hit
            float* rg = my_t + (BT::kG * 3 + nslot * 6) * 32 + lane;       // geometry rows of this hit
            rs[0] = t0; rs[32] = t1; rs[64] = t2;
            rg[0] = e0; rg[32] = e1; rg[64] = e2; rg[96] = e3; rg[128] = e4; rg[160] = e5;
            slots |= (unsigned)jj << (8 * nslot);
            if (++nslot == BT::kG) flush_tbuf();
          } else {
            vals[0] = e0; vals[1] = e1; vals[2] = e2; vals[3] = e3; vals[4] = e4; vals[5] = e5;
            warp_reduce_halving<K>(vals, lane);
            float* row = s_acc + jj * STR;
            if constexpr (K == 64) {
              if (red_e < NV) atomicAdd(row + red_e, vals[0]);
              if (red_e + 1 < NV) atomicAdd(row + red_e + 1, vals[1]);
            } else {
              if (red_writer && red_e < NV) atomicAdd(row + red_e, vals[0]);
            }
          }
        }
        if (touched && lane == 0) atomicOr(&touched_now[r], touched);
        if (__all_sync(kFull, done)) { warp_done = true; break; }
      }
      if (nslot) flush_tbuf();  // the batch's staging buffer (and its entry indices) is about to be recycled
    }
    __syncthreads();
    // flush this batch's accumulators: one (Gaussian, tile) instance per thread
    if (tid < B / 32) s_touched[(b + 1) & 1][tid] = 0u;  // recycled for batch b+1 (last read before barrier C of b-1)
    if (tid < cnt && ((touched_now[tid >> 5] >> (tid & 31)) & 1u)) {
      const int id = reinterpret_cast<const int*>(st + L::kIds)[tid];
      float* row = s_acc + tid * STR;
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (TBUF) {
        // [3*CC sh | 6 geometry]: streamed 16 bytes at a time (few live registers)
        float* dst = a.grad_pay + (size_t)id * (3 * CC);
        if ((3 * CC) % 4 == 0 && ((reinterpret_cast<uintptr_t>(a.grad_pay) & 15) == 0)) {
#pragma unroll 4
          for (int k = 0; k < 3 * CC; k += 4) {
            const float4 q = *reinterpret_cast<float4*>(row + k);
            *reinterpret_cast<float4*>(row + k) = z4;
            red_add_v4(dst + k, q.x, q.y, q.z, q.w);
          }
        } else {
#pragma unroll 3
          for (int k = 0; k < 3 * CC; ++k) {
            const float q = row[k];
            row[k] = 0.f;
            red_add(dst + k, q);
          }
        }
        float g[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { g[k] = row[BT::kGeoOff + k]; row[BT::kGeoOff + k] = 0.f; }
        if constexpr (FUSED) {
          red_add_v4(a.ggeom + (size_t)id * 8, g[0], g[1], g[2], g[3]);
          red_add_v4(a.ggeom + (size_t)id * 8 + 4, g[4], g[5], 0.f, 1.0f);  // .w: touch counter
        } else {
          red_add_v2(a.grad_mean + (size_t)id * 2, g[0], g[1]);
          red_add_v4(a.grad_cov + (size_t)id * 4, g[2], g[3], g[3], g[4]);
          red_add(a.grad_alpha + id, g[5]);
        }
      } else {
        float g[STR];
#pragma unroll
        for (int k = 0; k < STR; k += 4) {
          float4 q = *reinterpret_cast<float4*>(row + k);
          g[k] = q.x; g[k + 1] = q.y; g[k + 2] = q.z; g[k + 3] = q.w;
          *reinterpret_cast<float4*>(row + k) = z4;
        }
        if constexpr (FUSED) {
          float gd = 0.f;
          if constexpr (PAY == PAY_RGB && EXTRAS) gd = g[9];
          red_add_v4(a.ggeom + (size_t)id * 8, g[0], g[1], g[2], g[3]);
          red_add_v4(a.ggeom + (size_t)id * 8 + 4, g[4], g[5], gd, 1.0f);  // .w: touch counter
        } else {
          red_add_v2(a.grad_mean + (size_t)id * 2, g[0], g[1]);
          red_add_v4(a.grad_cov + (size_t)id * 4, g[2], g[3], g[3], g[4]);
          red_add(a.grad_alpha + id, g[5]);
        }
        if constexpr (PAY == PAY_SH) {  // C == 1: three coefficients
          float* dst = a.grad_pay + (size_t)id * (3 * CC);
#pragma unroll
          for (int k = 0; k < 3 * CC; ++k) red_add(dst + k, g[6 + k]);
        } else if constexpr (PAY == PAY_RGB) {
          if constexpr (FUSED) {
            red_add_v4(a.gpay + (size_t)id * 4, g[6], g[7], g[8], 0.f);
          } else {
            red_add(a.grad_pay + (size_t)id * 3, g[6]);
            red_add(a.grad_pay + (size_t)id * 3 + 1, g[7]);
            red_add(a.grad_pay + (size_t)id * 3 + 2, g[8]);
          }
        } else {
          red_add(a.grad_pay + id, g[6]);
        }
      }
    }
    if (has_next) {
      cp_async_wait<0>();
      if (PAY == PAY_SH && use_bulk) mbar_wait(&s_bar[(b + 1) & 1], (uint32_t)(((b + 1) >> 1) & 1));
    }
    if (__syncthreads_and(warp_done ? 1 : 0)) break;
  }

  if (PAY == PAY_RGB && a.g_bg && pg.inside) {  // gs/renderer.py:1282 nan_to_num(grad * T)
    const float gv[3] = {go0 * T, go1 * T, go2 * T};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = gv[c];
      a.g_bg[3 * pix + c] = (v != v) ? 0.f : fminf(fmaxf(v, -3.4028235e38f), 3.4028235e38f);
    }
  }
}

template <int PAY, int C, bool EXTRAS, bool FUSED, int B>
static int launch_one(const CompositeArgs& a, cudaStream_t st) {
  using L = StageLayout<PAY, C, B, true>;
  using BT = BwdTraits<PAY, C, EXTRAS>;
  const size_t smem = 2 * (size_t)L::kBytes + (size_t)B * BT::kStride * 4 + (size_t)8 * BT::kTbufFloats * 4;
  auto kern = k_composite_bwd<PAY, C, EXTRAS, FUSED, B>;
  GSB_CUDA(ensure_max_dyn_smem(reinterpret_cast<const void*>(kern), (int)smem, a.device));
  dim3 grid(a.tiles_w, a.tiles_h, 1);
  kern<<<grid, kCtaThreads, smem, st>>>(a);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}

int launch_composite_bwd(int pay_kind, int C, bool extras, bool fused, const CompositeArgs& a, cudaStream_t st) {
  if (a.tiles_w <= 0 || a.tiles_h <= 0) return GSB200_OK;
  switch (pay_kind) {
    case PAY_RGB:
      if (fused) return extras ? launch_one<PAY_RGB, 1, true, true, 256>(a, st)
                               : launch_one<PAY_RGB, 1, false, true, 256>(a, st);
      return launch_one<PAY_RGB, 1, false, false, 256>(a, st);
    case PAY_SCALAR:
      return launch_one<PAY_SCALAR, 1, false, false, 256>(a, st);
    case PAY_SH:
      switch (C) {
        case 1: return fused ? launch_one<PAY_SH, 1, false, true, 256>(a, st) : launch_one<PAY_SH, 1, false, false, 256>(a, st);
        case 2: return fused ? launch_one<PAY_SH, 2, false, true, 128>(a, st) : launch_one<PAY_SH, 2, false, false, 128>(a, st);
        case 3: return fused ? launch_one<PAY_SH, 3, false, true, GSB_BWD_B>(a, st) : launch_one<PAY_SH, 3, false, false, GSB_BWD_B>(a, st);
        case 4: return fused ? launch_one<PAY_SH, 4, false, true, GSB_BWD_B>(a, st) : launch_one<PAY_SH, 4, false, false, GSB_BWD_B>(a, st);
        default: break;
      }
    default: break;
  }
  set_error("composite_bwd: unsupported payload kind %d / C %d", pay_kind, C);
  return GSB200_ERR_UNSUPPORTED;
}

}  // namespace gsb
