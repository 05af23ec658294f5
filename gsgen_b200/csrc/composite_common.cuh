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

// overlap test of a splat's a*G >= 1/255 bounding box with the warp's pixel block
__device__ __forceinline__ bool splat_hits_block(const float4& g0, const float4& g1, const PixelGeom& g) {
  return (g0.x - g1.z <= g.X1) && (g0.x + g1.z >= g.X0) && (g0.y - g1.w <= g.Y1) && (g0.y + g1.w >= g.Y0);
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
