// composite_common.cuh -- tile/warp geometry and the asynchronous staging pipeline shared by the forward
// and backward composite kernels.
//
// One CTA (256 threads) per 16x16 tile (same decomposition as the reference, vol_render.h:1064-1079), but:
//   * a warp owns an 8x4 pixel block, so one ballot per 32 list entries decides which Gaussians can touch
//     the block at all (bounding box of the a*G >= 1/255 region) and the warp only walks those;
//   * list entries are gathered BY ID from 32-byte AoS splat records (two float4) with cp.async (LDGSTS,
//     16 B, L2-only) and, for spherical harmonics, one cp.async.bulk (TMA engine, UBLKCP) per Gaussian
//     whose 3*C*C floats are contiguous -- completion through an mbarrier transaction count;
//   * two stages: the copies for batch b+1 are in flight while batch b is composited; the sorted ids for
//     batch b+2 are already being fetched into a register;
//   * per-pixel state lives in registers; a warp stops when all of its pixels have T < thresh and the
//     CTA stops fetching when every warp has.
#pragma once
#include "kernels.cuh"

#ifndef GSB_ELLIPSE_CULL
#define GSB_ELLIPSE_CULL 1  // exact ellipse-vs-block test behind the box test (splat_hits_block)
#endif

namespace gsb {

constexpr int kTile = 16;
constexpr int kCtaThreads = 256;
constexpr unsigned kFull = 0xffffffffu;

template <int PAY, int C> struct PayTraits {
  static constexpr int CC = C * C;
  // floats of payload staged per list entry
  static constexpr int kPayFloats = (PAY == PAY_SH) ? 3 * CC : 4;
  static constexpr bool kBulkOk = (PAY == PAY_SH) && ((3 * CC * 4) % 16 == 0);
};

template <int PAY, int C, int B, bool WITH_IDS> struct StageLayout {
  using PT = PayTraits<PAY, C>;
  static constexpr int kG0 = 0;                                   // float4[B]
  static constexpr int kG1 = kG0 + B * 16;                        // float4[B]
  static constexpr int kPay = kG1 + B * 16;                       // kPayFloats*4 bytes per entry
  static constexpr int kIds = kPay + B * PT::kPayFloats * 4;      // int32[B] (backward only)
  static constexpr int kBytes = ((kIds + (WITH_IDS ? B * 4 : 0)) + 15) / 16 * 16;
};

struct PixelGeom {
  int gx, gy;        // pixel coordinates
  bool inside;
  float px, py;      // camera-plane position of the pixel corner (vol_render.h:189-190)
  float X0, X1, Y0, Y1;  // camera-plane extent of the warp's 8x4 block
};

__device__ __forceinline__ PixelGeom pixel_geom(const CompositeArgs& a, int tile_x, int tile_y, int warp, int lane) {
  PixelGeom g;
  const float tlx = a.topleft_ptr ? a.topleft_ptr[0] : a.tlx;
  const float tly = a.topleft_ptr ? a.topleft_ptr[1] : a.tly;
  const int bx0 = tile_x * kTile + (warp & 1) * 8, by0 = tile_y * kTile + (warp >> 1) * 4;
  g.gx = bx0 + (lane & 7);
  g.gy = by0 + (lane >> 3);
  g.inside = (g.gx < a.W) && (g.gy < a.H);
  g.px = fmaf((float)g.gx, a.psx, tlx);
  g.py = fmaf((float)g.gy, a.psy, tly);
  g.X0 = fmaf((float)bx0, a.psx, tlx);
  g.X1 = fmaf((float)(bx0 + 7), a.psx, tlx);
  g.Y0 = fmaf((float)by0, a.psy, tly);
  g.Y1 = fmaf((float)(by0 + 3), a.psy, tly);
  return g;
}

// Issue the asynchronous copies of one list entry (thread `t` stages entry `t` of the batch).
template <int PAY, int C, int B, bool WITH_IDS>
__device__ __forceinline__ void stage_entry(const CompositeArgs& a, unsigned char* stage, int t, int id, bool valid,
                                            bool use_bulk, uint64_t* bar) {
  using L = StageLayout<PAY, C, B, WITH_IDS>;
  using PT = PayTraits<PAY, C>;
  if (valid) {
    const float4* src = reinterpret_cast<const float4*>(a.splat + id);
    cp_async16(stage + L::kG0 + t * 16, src);
    cp_async16(stage + L::kG1 + t * 16, src + 1);
    if constexpr (PAY == PAY_SH) {
      const float* s = a.sh + (size_t)id * (3 * PT::CC);
      float* d = reinterpret_cast<float*>(stage + L::kPay) + t * (3 * PT::CC);
      if (PT::kBulkOk && use_bulk) {
        bulk_g2s(d, s, 3 * PT::CC * 4, bar);
      } else {
#pragma unroll
        for (int k = 0; k < 3 * PT::CC; ++k) cp_async4(d + k, s + k);
      }
    } else {
      cp_async16(stage + L::kPay + t * 16, a.pay + id);
    }
    if constexpr (WITH_IDS) reinterpret_cast<int*>(stage + L::kIds)[t] = id;
  }
  cp_async_commit();
}

// Can the splat reach a*G >= 1/255 anywhere in the warp's pixel block?  Two conservative tests:
//  (1) the axis-aligned box of the a*G >= 1/255 ellipse against the block (round 1);
//  (2) round 2: the ellipse itself.  In the record's coordinates G = exp2(-(u^2+v^2)) with (u,v) = (p0*dx + p1*dy, p2*dy),
//      so a*G >= 1/255 <=> u^2+v^2 <= log2(255 a) =: L, and the block rectangle maps to a parallelogram in the (u,v)
//      plane: the splat reaches the block iff the parallelogram comes within sqrt(L) of the origin.  Measured at C3
//      (profiles/r2_ncu_full_c3_call4_*): 29 % of the box hits had no pixel above the threshold -- the box of a
//      rotated, elongated ellipse is loose.  The test costs ~45 instructions per (warp, list entry), once per 32
//      entries and lane; an evaluated hit costs 20-120 per warp.  L is inflated (1e-3 relative + 1e-4): the per-pixel
//      test stays the arbiter, this one only must never reject a splat a pixel would accept.
__device__ __forceinline__ bool splat_hits_block(const float4& g0, const float4& g1, const PixelGeom& g) {
  if (!((g0.x - g1.z <= g.X1) && (g0.x + g1.z >= g.X0) && (g0.y - g1.w <= g.Y1) && (g0.y + g1.w >= g.Y0))) return false;
#if GSB_ELLIPSE_CULL
  const float dx0 = g.X0 - g0.x, dx1 = g.X1 - g0.x, dy0 = g.Y0 - g0.y, dy1 = g.Y1 - g0.y;
  if (dx0 <= 0.f && dx1 >= 0.f && dy0 <= 0.f && dy1 >= 0.f) return true;  // centre inside the block
  const float L = fmaf(__log2f(255.0f * g1.y), 1.001f, 1e-4f);
  const float p0 = g0.z, p1 = g0.w, p2 = g1.x;
  // edges dy = const: v fixed, u spans [p0*dx0, p0*dx1] + p1*dy (p0 > 0: Cholesky diagonal)
  float best;
  {
    const float ua = p0 * dx0, ub = p0 * dx1;
    const float o0 = p1 * dy0, o1 = p1 * dy1;
    const float v0 = p2 * dy0, v1 = p2 * dy1;
    const float c0 = fminf(fmaxf(0.f, ua + o0), ub + o0);  // closest u to 0 on the edge (ua <= ub)
    const float c1 = fminf(fmaxf(0.f, ua + o1), ub + o1);
    best = fminf(fmaf(c0, c0, v0 * v0), fmaf(c1, c1, v1 * v1));
    // edges dx = const: (u,v) = (a_i + p1*t, p2*t), t in [dy0, dy1]; minimiser t* = -a_i*p1 / (p1^2 + p2^2), clamped
    const float inv = rcp_approx(fmaf(p1, p1, p2 * p2));
    const float t0 = fminf(fmaxf(-ua * p1 * inv, dy0), dy1), t1 = fminf(fmaxf(-ub * p1 * inv, dy0), dy1);
    const float e0u = fmaf(p1, t0, ua), e0v = p2 * t0, e1u = fmaf(p1, t1, ub), e1v = p2 * t1;
    best = fminf(best, fminf(fmaf(e0u, e0u, e0v * e0v), fmaf(e1u, e1u, e1v * e1v)));
  }
  return best <= L;
#else
  return true;
#endif
}

// a*G for one pixel:  G = exp2(-(u^2+v^2)),  u = p0*dx + p1*dy,  v = p2*dy   (see make_splat)
__device__ __forceinline__ float splat_aG(const float4& g0, const float4& g1, float px, float py, float* G_out,
                                          float* u_out, float* v_out) {
  float dx = px - g0.x, dy = py - g0.y;
  float u = fmaf(g0.z, dx, g0.w * dy);
  float v = g1.x * dy;
  float e = fmaf(-u, u, -(v * v));
  float G = ex2_approx(e);
  *G_out = G; *u_out = u; *v_out = v;
  return g1.y * G;
}

}  // namespace gsb
