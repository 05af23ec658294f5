// binning.cu -- tile binning (reference gs/src/include/aabb_culling.h:15-103, :192-260):
//   counts --cub::DeviceScan--> offsets --k_emit_keys--> (tile<<32 | depth_bits, id)
//          --cub::DeviceRadixSort (bits [0, 32+ceil(log2 T)))--> sorted ids --k_tile_ranges--> start/end
// Differences from the reference, none of which change the result: slots come from a prefix sum instead
// of one contended global atomic (deterministic order, no memset of D keys), the sort skips the key bits
// that are always zero, start/end are produced by one kernel, scratch comes from the context arena
// instead of 5x cudaMalloc/cudaFree per call, and everything is enqueued on the caller's stream.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "gsb200_common.cuh"
#include "kernels.cuh"

namespace gsb {

// One warp expands 32 consecutive Gaussians cooperatively: for each Gaussian with duplicates the 32
// lanes write its (key, id) pairs to consecutive slots (coalesced 8 B / 4 B stores).
__global__ void __launch_bounds__(256)
k_emit_keys(uint32_t N, const int32_t* __restrict__ count, const int32_t* __restrict__ incl,
            const ushort4* __restrict__ rect, const float* __restrict__ depth, int tiles_w,
            uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t base = warp_global * 32;
  if (base >= N) return;
  const uint32_t i = base + lane;
  int c = 0, off = 0;
  uint32_t dbits = 0;
  ushort4 r = make_ushort4(0, 0, 0, 0);
  if (i < N) {
    c = count[i];
    if (c > 0) {
      off = incl[i] - c;
      r = rect[i];
      dbits = __float_as_uint(depth[i]);
    }
  }
  uint32_t have = __ballot_sync(0xffffffffu, c > 0);
  while (have) {
    int g = __ffs(have) - 1;
    have &= have - 1;
    int cg = __shfl_sync(0xffffffffu, c, g);
    int og = __shfl_sync(0xffffffffu, off, g);
    uint32_t dg = __shfl_sync(0xffffffffu, dbits, g);
    int x0 = __shfl_sync(0xffffffffu, (int)r.x, g), y0 = __shfl_sync(0xffffffffu, (int)r.y, g);
    int x1 = __shfl_sync(0xffffffffu, (int)r.z, g);
    int w = x1 - x0 + 1;
    for (int k = lane; k < cg; k += 32) {
      int ty = y0 + k / w, tx = x0 + k % w;
      uint32_t tile = (uint32_t)(ty * tiles_w + tx);
      keys[og + k] = ((uint64_t)tile << 32) | (uint64_t)dg;
      vals[og + k] = (int32_t)(base + g);
    }
  }
}

__global__ void __launch_bounds__(256)
k_fill_i32(int32_t* p, int32_t v, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// aabb_culling.h:70-103 fill_start_aabb + fill_end_aabb in one pass
__global__ void __launch_bounds__(256)
k_tile_ranges(int64_t D, const uint64_t* __restrict__ keys, int32_t* __restrict__ start, int32_t* __restrict__ end) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= D) return;
  uint32_t t = (uint32_t)(keys[s] >> 32);
  if (s == 0 || (uint32_t)(keys[s - 1] >> 32) != t) start[t] = (int32_t)s;
  if (s == D - 1 || (uint32_t)(keys[s + 1] >> 32) != t) end[t] = (int32_t)(s + 1);
}

__global__ void k_total_from_scan(uint32_t N, const int32_t* __restrict__ incl, int64_t* __restrict__ total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *total = N ? (int64_t)incl[N - 1] : 0;
}

int scan_counts(gsb200_ctx* ctx, uint32_t N, cudaStream_t st) {
  if (N == 0) return GSB200_OK;
  size_t bytes = 0;
  GSB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, ctx->count.as<int32_t>(), ctx->incl.as<int32_t>(), (int)N, st));
  int rc = ctx->cub_tmp.reserve(bytes);
  if (rc) return rc;
  GSB_CUDA(cub::DeviceScan::InclusiveSum(ctx->cub_tmp.p, bytes, ctx->count.as<int32_t>(), ctx->incl.as<int32_t>(),
                                         (int)N, st));
  return GSB200_OK;
}

// reads the duplicate count back to the host (the reference does the same: gs/culling.py:33-35 .item() and
// aabb_culling.h:227 cudaMemcpy).  One 8-byte D2H copy + stream sync per view.
int read_total(gsb200_ctx* ctx, uint32_t N, int64_t* h_total, cudaStream_t st) {
  int rc = ctx->d_total.reserve(sizeof(int64_t));
  if (rc) return rc;
  k_total_from_scan<<<1, 32, 0, st>>>(N, ctx->incl.as<int32_t>(), ctx->d_total.as<int64_t>());
  GSB_LAUNCH_CHECK();
  GSB_CUDA(cudaMemcpyAsync(ctx->h_total, ctx->d_total.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  GSB_CUDA(cudaStreamSynchronize(st));
  *h_total = *ctx->h_total;
  return GSB200_OK;
}

static int bits_for(uint32_t n) {  // ceil(log2(n)) for n >= 1
  int b = 0;
  while ((1u << b) < n) ++b;
  return b;
}

// count/incl/rect in ctx, depth from the caller.  Writes sorted ids to ids_out (or leaves them in
// ctx->vals[ctx->sorted_sel] when ids_out == nullptr) and start/end.
int bin_and_sort(gsb200_ctx* ctx, uint32_t N, int64_t D, const float* depth, int tiles_h, int tiles_w,
                 int32_t* ids_out, int32_t* start, int32_t* end, cudaStream_t st) {
  const uint32_t T = (uint32_t)tiles_h * (uint32_t)tiles_w;
  if (T) {
    k_fill_i32<<<(T + 255) / 256, 256, 0, st>>>(start, -1, T);
    k_fill_i32<<<(T + 255) / 256, 256, 0, st>>>(end, -1, T);
    GSB_LAUNCH_CHECK();
  }
  ctx->D = D;
  if (D == 0 || N == 0) return GSB200_OK;
  int rc;
  for (int k = 0; k < 2; ++k) {
    if ((rc = ctx->keys[k].reserve((size_t)D * 8))) return rc;
    if (!(k == 1 && ids_out)) {
      if ((rc = ctx->vals[k].reserve((size_t)D * 4))) return rc;
    }
  }
  uint64_t* k0 = ctx->keys[0].as<uint64_t>();
  uint64_t* k1 = ctx->keys[1].as<uint64_t>();
  int32_t* v0 = ctx->vals[0].as<int32_t>();
  int32_t* v1 = ids_out ? ids_out : ctx->vals[1].as<int32_t>();
  uint32_t warps = (N + 31) / 32;
  k_emit_keys<<<(warps * 32 + 255) / 256, 256, 0, st>>>(N, ctx->count.as<int32_t>(), ctx->incl.as<int32_t>(),
                                                       ctx->rect.as<ushort4>(), depth, tiles_w, k0, v0);
  GSB_LAUNCH_CHECK();
  GSB_CHECK(D < (int64_t)2147483647, GSB200_ERR_INVALID, "N_with_dub %lld exceeds int32", (long long)D);
  const int end_bit = 32 + bits_for(T);
  size_t bytes = 0;
  GSB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0, k1, v0, v1, (int)D, 0, end_bit, st));
  if ((rc = ctx->cub_tmp.reserve(bytes))) return rc;
  GSB_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, bytes, k0, k1, v0, v1, (int)D, 0, end_bit, st));
  ctx->sorted_sel = 1;
  k_tile_ranges<<<(unsigned)((D + 255) / 256), 256, 0, st>>>(D, k1, start, end);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}

}  // namespace gsb
