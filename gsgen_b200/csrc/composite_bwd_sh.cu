// composite_bwd_sh.cu -- per-tile alpha compositing backward for spherical-harmonics payloads of degree >= 1
// (tile_based_vol_rendering_backward_sh_entry<C>[_bg], vol_render_sh.h:353-455 with the inner loop
// :268-351 and backward_C :28-36; vol_render_bg.h:131-242).  Round-2 rewrite of the SH specialisation of
// composite_bwd.cu (which keeps RGB / scalar / SH degree 0).
//
// Same per-pair algebra as composite_bwd.cu (one running scalar S per pixel).  What changed, and why (measured on the
// round-1 kernel, profiles/r1_ncu_full_c3_run10: 734 M warp instructions, 59 % issue slots, top stall = block barrier):
//   * the per-batch shared accumulator is gone.  Round 1 added every warp's partial sums into a [B x 54] shared array
//     with float atomics -- CAS loops in SASS (ATOMS.CAST.SPIN), 205 instructions per flush = 19 % of the kernel --
//     then needed a second block barrier per batch and a per-batch flush of that array.  Now a warp's transpose-buffer
//     product leaves 4 consecutive coefficients of one (hit, channel) row in each lane and goes straight to HBM/L2 with
//     ONE red.global.add.v4.f32 per lane (2 per flush), the six geometry sums with one v4 + one v2 reduction per hit.
//     L2 float atomics are fire-and-forget and aggregate in the 126 MB L2; the (Gaussian, tile) instance is reduced
//     there instead of in shared memory.
//   * ONE block barrier per batch (like the forward): it retires the staging buffer and votes on early termination.
//   * product mapping: lane = (pixel half, k-quad, row group): 96 units of work, exactly 3 per lane, 7 LDS.128 per
//     24 FFMA2 (round 1: 7 per 12); see flush_direct.
#include "composite_common.cuh"

namespace gsb {

template <int C> struct ShBwdTraits {
  static constexpr int CC = C * C;
  static constexpr int kG = 4;                          // hits per flush
  static constexpr int kRows = 3 * kG;                  // SH rows of a flush (hit*3 + channel) = 12
  // product mapping: lane = (pixel part p [kPH], k-quad kq [kKQ], hit h [4]); a lane owns the hit's 3 channel rows
  static constexpr int kKQ = (CC <= 4) ? 1 : 4;         // k-quads (4 coefficients each)
  static constexpr int kKL = 4 * kKQ;                   // basis rows held in shared memory (zero beyond CC)
  static constexpr int kPH = 8 / kKQ;                   // pixel parts (2 for C >= 3, 8 for C = 2)
  static constexpr int kNG = 8 / kPH;                   // 16-byte pixel groups per lane (4 or 1)
  static constexpr int kTFloats = kG * 9 * 32;          // per warp: [kRows + 6*kG][32 px]
  static constexpr int kYFloats = kKL * 32;             // per warp: basis [k][32 px]
  static constexpr int kWarpFloats = kTFloats + kYFloats;
  static constexpr bool kVec4 = (CC % 4 == 0);          // g_sh rows are 16-byte granular
};

struct FlushDst {  // by value: a reference to the kernel's argument struct would force a local-memory copy of it
  float* pay;      // g_sh [N,3,C*C]
  float* g0;       // fused: gradient records [N,8];  else grad_mean [N,2]
  float* g1;       // else grad_cov [N,4]
  float* g2;       // else grad_alpha [N]
};

// 16-byte shared-memory load with base + compile-time immediate addressing
template <int IMM> __device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];\n"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr), "n"(IMM));
  return v;
}

// Flush of a warp's transpose buffer.  SH rows T[hit*3+c][32 px] x basis Ysm[k][32 px] -> g_sh partial sums; geometry
// rows -> row sums -> the Gaussian's gradient record (fused) or the reference-layout grad_mean / grad_cov / grad_alpha.
//
// Product (C >= 3): 4 hits x 3 channels x 4 k-quads x 2 pixel halves = 96 units of (4 coefficients x 16 pixels),
// exactly 3 per lane: lane = (p, kq, h) owns hit h's three channel rows, k-quad kq, pixel half p -- no division to
// find (hit, channel), one id look-up per lane, and the three reductions of a lane differ by an immediate offset.
// Per 4 pixels a lane issues 4 LDS.128 of the basis + 3 LDS.128 of the per-pixel factors for 24 FFMA2 (round 1: 7 LDS
// for 12 FFMA2) and no FMA is wasted.  Both buffers are stored UNROTATED (the hit loop's stores are base + immediate);
// at step q a lane reads pixel group 4p + (kq ^ q): the 8 (p, kq) combinations of a quarter-warp land on 8 different
// 16-byte groups -> no bank conflict on the basis rows (same 128-byte alignment), lanes sharing a factor row broadcast,
// and because rows are 128-byte aligned the address is (base ^ (q << 4)) + immediate: 3 LOP3 per buffer and flush.
// The pixel parts are combined with one shuffle per value and ONE lane per (hit, k-quad) sends 4 consecutive
// coefficients per channel with red.global.add.v4.f32.
template <int C, bool FUSED>
__device__ __noinline__ void flush_direct(const FlushDst a, const float* my_t, const float* my_y,
                                          const int* ids_stage, int nslot, unsigned slots, int lane) {
  using ST = ShBwdTraits<C>;
  constexpr int CC = ST::CC;
  const int p = lane & (ST::kPH - 1);
  const int kq = (lane / ST::kPH) & (ST::kKQ - 1);
  const int h = lane / (ST::kPH * ST::kKQ);  // 0..3
  __syncwarp();
  float acc[3][4], aco[3][4];  // [channel][k in quad]: even / odd pixel partial sums (FFMA2)
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; aco[i][j] = 0.f; }
  const uint32_t g0 = (uint32_t)(p * ST::kNG + (ST::kNG == 4 ? kq : 0)) << 4;   // byte offset of the lane's first group
  const uint32_t at = smem_u32(my_t) + (uint32_t)h * (3 * 128) + g0;             // rows 3h, 3h+1, 3h+2 at +0/+128/+256
  const uint32_t ay = smem_u32(my_y) + (uint32_t)kq * (4 * 128) + g0;            // rows 4kq .. 4kq+3
#pragma unroll
  for (int q = 0; q < ST::kNG; ++q) {
    const uint32_t atq = at ^ (uint32_t)(q << 4), ayq = ay ^ (uint32_t)(q << 4);
    const float4 y0 = lds128<0>(ayq), y1 = lds128<128>(ayq), y2 = lds128<256>(ayq), y3 = lds128<384>(ayq);
    const float4 t0 = lds128<0>(atq), t1 = lds128<128>(atq), t2 = lds128<256>(atq);
    const float4 ts[3] = {t0, t1, t2};
    const float4 ys[4] = {y0, y1, y2, y3};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ffma2(acc[i][j], aco[i][j], ts[i].x, ts[i].y, ys[j].x, ys[j].y);
        ffma2(acc[i][j], aco[i][j], ts[i].z, ts[i].w, ys[j].z, ys[j].w);
      }
  }
  float v[3][4];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[i][j] = acc[i][j] + aco[i][j];
#pragma unroll
      for (int x = 1; x < ST::kPH; x <<= 1) v[i][j] += __shfl_xor_sync(kFull, v[i][j], x);  // the other pixel parts
    }
  if (p == 0 && h < nslot) {
    const int id = ids_stage[(slots >> (8 * h)) & 255u];
    float* dst = a.pay + (size_t)id * (3 * CC) + 4 * kq;
    if (ST::kVec4 && ((reinterpret_cast<uintptr_t>(a.pay) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 3; ++i) red_add_v4(dst + i * CC, v[i][0], v[i][1], v[i][2], v[i][3]);
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (4 * kq + j < CC) red_add(dst + i * CC + j, v[i][j]);
    }
  }
  // geometry: row sums, one row per lane (lane = hit*6 + value; at step q the lane reads 16-byte group q ^ (lane & 7):
  // conflict-free), then gathered into the hit's first lane
  float se = 0.f, so = 0.f;
  if (lane < nslot * 6) {
    const uint32_t ag = smem_u32(my_t) + (uint32_t)(ST::kRows + lane) * 128 + ((uint32_t)(lane & 7) << 4);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = lds128<0>(ag ^ (uint32_t)(q << 4));
      ffma2(se, so, t.x, t.y, 1.0f, 1.0f);
      ffma2(se, so, t.z, t.w, 1.0f, 1.0f);
    }
  }
  const float s = se + so;
  const float s1 = __shfl_down_sync(kFull, s, 1), s2 = __shfl_down_sync(kFull, s, 2);
  const float s3 = __shfl_down_sync(kFull, s, 3), s4 = __shfl_down_sync(kFull, s, 4);
  const float s5 = __shfl_down_sync(kFull, s, 5);
  if (lane < nslot * 6 && (lane % 6) == 0) {
    const int hh = lane / 6;
    const int id = ids_stage[(slots >> (8 * hh)) & 255u];
    if constexpr (FUSED) {  // {gmx, gmy, gxx, gxy | gyy, galpha, gdepth, #flushes}
      red_add_v4(a.g0 + (size_t)id * 8, s, s1, s2, s3);
      red_add_v4(a.g0 + (size_t)id * 8 + 4, s4, s5, 0.f, 1.0f);  // .w counts the flushes: != 0 <=> the Gaussian was touched
    } else {
      red_add_v2(a.g0 + (size_t)id * 2, s, s1);
      red_add_v4(a.g1 + (size_t)id * 4, s2, s3, s3, s4);
      red_add(a.g2 + id, s5);
    }
  }
  __syncwarp();
}

#ifndef GSB_BWDSH_B
#define GSB_BWDSH_B 32  // list entries per staged batch
#endif
#ifndef GSB_BWDSH_PREFETCH
#define GSB_BWDSH_PREFETCH 0  // 1: hit-loop software pipeline (next record fetched while the current one is evaluated); measured round 2: 1.028 ms with, 1.013 ms without (C3)
#endif
#ifndef GSB_BWDSH_NW
#define GSB_BWDSH_NW 8  // warps per CTA: 8 = one CTA per 16x16 tile; 4 = two CTAs per tile (16x8 pixels each, both walk
#endif                  // the tile's list; fewer warps wait at the per-batch barrier and a half tile terminates earlier)
#ifndef GSB_BWDSH_MINBLOCKS
#define GSB_BWDSH_MINBLOCKS (GSB_BWDSH_NW == 8 ? 3 : 6)
#endif

template <int C, bool FUSED, int B>
__global__ void __launch_bounds__(GSB_BWDSH_NW * 32, GSB_BWDSH_MINBLOCKS)
k_composite_bwd_sh(const CompositeArgs a) {
  using L = StageLayout<PAY_SH, C, B, true>;
  using PT = PayTraits<PAY_SH, C>;
  using ST = ShBwdTraits<C>;
  constexpr int CC = ST::CC;
  static_assert(CC >= 4, "SH degree 0 uses the butterfly kernel in composite_bwd.cu");
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t s_bar[2];

  float* s_tbuf = reinterpret_cast<float*>(smem + 2 * L::kBytes);  // [GSB_BWDSH_NW warps][kWarpFloats]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile = tile_y * a.tiles_w + tile_x;
  const PixelGeom pg = pixel_geom(a, tile_x, tile_y, warp + (int)blockIdx.z * GSB_BWDSH_NW, lane);
  const int pix = pg.gy * a.W + pg.gx;

  const int s0 = a.start[tile];
  const int n = (s0 < 0) ? 0 : (a.end[tile] - s0);
  if (n <= 0) return;

  const bool use_bulk = PT::kBulkOk && ((reinterpret_cast<uintptr_t>(a.sh) & 15) == 0);
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  // per-pixel registers
  float T = 1.0f;
  bool done = !pg.inside || (1.0f < a.thresh);
  float go0 = 0.f, go1 = 0.f, go2 = 0.f;
  float S = 0.f;  // sum_ch go_ch * (F_ch - Cacc_ch)
  if (pg.inside) {
    if (a.gout) { go0 = a.gout[3 * pix]; go1 = a.gout[3 * pix + 1]; go2 = a.gout[3 * pix + 2]; }
    S = go0 * a.fin[3 * pix] + go1 * a.fin[3 * pix + 1] + go2 * a.fin[3 * pix + 2];
  }
  float Y[CC];
  {
    float c9[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) c9[k] = a.c9_ptr ? a.c9_ptr[k] : a.c9[k];
    float d[3];
    pixel_dir(pg.px, pg.py, c9, d);
    sh_basis<C>(d[0], d[1], d[2], Y);
  }
  // per-warp transpose buffer: SH rows [G*3][32] (row = hit*3 + channel), geometry rows [G*6][32], then the block's
  // basis matrix Ysm[k][32] (k-major); everything unrotated, the reader picks conflict-free groups (flush_direct)
  float* my_t = s_tbuf + warp * ST::kWarpFloats;
  float* my_y = my_t + ST::kTFloats;
#pragma unroll
  for (int k = 0; k < ST::kKL; ++k) my_y[k * 32 + lane] = (k < CC) ? Y[k < CC ? k : 0] : 0.f;
  __syncwarp();
  int nslot = 0;        // hits buffered in my_t (warp-uniform)
  unsigned slots = 0u;  // their batch entry indices, 8 bits each
  const FlushDst dst = FUSED ? FlushDst{a.grad_pay, a.ggeom, nullptr, nullptr}
                             : FlushDst{a.grad_pay, a.grad_mean, a.grad_cov, a.grad_alpha};

  const int nb = (n + B - 1) / B;
  const int32_t* ids = a.ids + s0;
  {
    int cnt0 = min(B, n);
    int id0 = (tid < cnt0) ? ids[tid] : 0;
    if (use_bulk && tid == 0) mbar_arrive_expect_tx(&s_bar[0], (uint32_t)cnt0 * 3 * CC * 4);
    if (tid < B) stage_entry<PAY_SH, C, B, true>(a, smem, tid, id0, tid < cnt0, use_bulk, &s_bar[0]);
    else cp_async_commit();
  }
  int id_next = 0;
  if (nb > 1) { int j = B + tid; id_next = (tid < B && j < n) ? ids[j] : 0; }

  bool warp_done = __all_sync(kFull, done);
  cp_async_wait<0>();
  if (use_bulk) mbar_wait(&s_bar[0], 0u);
  __syncthreads();
  // ONE block barrier per batch: batch b+1 is staged before batch b is walked and waited for right before the barrier
  // that retires batch b's buffer and votes on early termination.
  for (int b = 0; b < nb; ++b) {
    unsigned char* st = smem + (b & 1) * L::kBytes;
    const int cnt = min(B, n - b * B);
    const bool has_next = (b + 1 < nb);
    if (has_next) {
      const int cntn = min(B, n - (b + 1) * B);
      uint64_t* barn = &s_bar[(b + 1) & 1];
      if (use_bulk && tid == 0) mbar_arrive_expect_tx(barn, (uint32_t)cntn * 3 * CC * 4);
      if (tid < B) stage_entry<PAY_SH, C, B, true>(a, smem + ((b + 1) & 1) * L::kBytes, tid, id_next, tid < cntn,
                                                   use_bulk, barn);
      else cp_async_commit();
      if (b + 2 < nb) { int j = (b + 2) * B + tid; id_next = (tid < B && j < n) ? ids[j] : 0; }
    }

    if (!warp_done) {
      const float4* sg0 = reinterpret_cast<const float4*>(st + L::kG0);
      const float4* sg1 = reinterpret_cast<const float4*>(st + L::kG1);
      const int* sids = reinterpret_cast<const int*>(st + L::kIds);
      for (int r = 0; r * 32 < cnt; ++r) {
        const int j = r * 32 + lane;
        bool hit = false;
        if (j < cnt) hit = splat_hits_block(sg0[j], sg1[j], pg);
        unsigned m = __ballot_sync(kFull, hit);
#if GSB_BWDSH_PREFETCH
        // software pipeline over the hits: the next hit's record is fetched while the current one is evaluated
        int bitn = __ffs(m) - 1;
        float4 n0 = sg0[m ? r * 32 + bitn : 0], n1 = sg1[m ? r * 32 + bitn : 0];
#endif
        while (m) {
#if GSB_BWDSH_PREFETCH
          const int jj = r * 32 + bitn;
          const float4 g0 = n0, g1 = n1;
          m &= m - 1;
          if (m) { bitn = __ffs(m) - 1; n0 = sg0[r * 32 + bitn]; n1 = sg1[r * 32 + bitn]; }
#else
          const int jj = r * 32 + (__ffs(m) - 1);
          m &= m - 1;
          const float4 g0 = sg0[jj], g1 = sg1[jj];
#endif
          float G, u, v;
          const float aG = splat_aG(g0, g1, pg.px, pg.py, &G, &u, &v);
          const bool ok = !done && (aG >= kMinRenderAlpha);
          if (!__any_sync(kFull, ok)) continue;

          const float w = ok ? aG * T : 0.f;
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
          const float gc = go0 * y[0] + go1 * y[1] + go2 * y[2];     // sum_ch go_ch * pay_ch
          const float t0 = w * (y[0] * (1.0f - y[0])) * go0;          // vol_render_sh.h:328-333
          const float t1 = w * (y[1] * (1.0f - y[1])) * go1;
          const float t2 = w * (y[2] * (1.0f - y[2])) * go2;
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
          {
            float* rs = my_t + nslot * (3 * 32) + lane;                  // SH rows of this hit
            float* rgm = my_t + (ST::kRows + nslot * 6) * 32 + lane;     // geometry rows of this hit
            rs[0] = t0; rs[32] = t1; rs[64] = t2;
            rgm[0] = e0; rgm[32] = e1; rgm[64] = e2; rgm[96] = e3; rgm[128] = e4; rgm[160] = e5;
            slots |= (unsigned)jj << (8 * nslot);
            if (++nslot == ST::kG) {
              flush_direct<C, FUSED>(dst, my_t, my_y, sids, nslot, slots, lane);
              nslot = 0; slots = 0u;
            }
          }
        }
        if (__all_sync(kFull, done)) { warp_done = true; break; }
      }
      if (nslot) {  // the batch's staging buffer (and its entry indices / ids) is about to be recycled
        flush_direct<C, FUSED>(dst, my_t, my_y, sids, nslot, slots, lane);
        nslot = 0; slots = 0u;
      }
    }
    if (has_next) {
      cp_async_wait<0>();
      if (use_bulk) mbar_wait(&s_bar[(b + 1) & 1], (uint32_t)(((b + 1) >> 1) & 1));
    }
    if (__syncthreads_and(warp_done ? 1 : 0)) break;
  }
}

template <int C, bool FUSED, int B>
static int launch_one_sh(const CompositeArgs& a, cudaStream_t st) {
  using L = StageLayout<PAY_SH, C, B, true>;
  using ST = ShBwdTraits<C>;
  static_assert(B <= 256, "entry indices are packed in 8 bits");
  static_assert(GSB_BWDSH_NW == 8 || GSB_BWDSH_NW == 4, "8 warps = whole tile, 4 = half tile");
  static_assert(B <= GSB_BWDSH_NW * 32, "one thread stages one list entry");
  const size_t smem = 2 * (size_t)L::kBytes + (size_t)GSB_BWDSH_NW * ST::kWarpFloats * 4;
  auto kern = k_composite_bwd_sh<C, FUSED, B>;
  GSB_CUDA(ensure_max_dyn_smem(reinterpret_cast<const void*>(kern), (int)smem, a.device));
  dim3 grid(a.tiles_w, a.tiles_h, 8 / GSB_BWDSH_NW);
  kern<<<grid, GSB_BWDSH_NW * 32, smem, st>>>(a);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}

int launch_composite_bwd_sh(int C, bool fused, const CompositeArgs& a, cudaStream_t st) {
  if (a.tiles_w <= 0 || a.tiles_h <= 0) return GSB200_OK;
  switch (C) {
    case 2: return fused ? launch_one_sh<2, true, 128>(a, st) : launch_one_sh<2, false, 128>(a, st);
    case 3: return fused ? launch_one_sh<3, true, GSB_BWDSH_B>(a, st) : launch_one_sh<3, false, GSB_BWDSH_B>(a, st);
    case 4: return fused ? launch_one_sh<4, true, GSB_BWDSH_B>(a, st) : launch_one_sh<4, false, GSB_BWDSH_B>(a, st);
    default: break;
  }
  set_error("composite_bwd_sh: unsupported C %d", C);
  return GSB200_ERR_UNSUPPORTED;
}

}  // namespace gsb
