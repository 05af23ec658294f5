// binning.cu -- tile binning (reference gs/src/include/aabb_culling.h:15-103, :192-260).
//
// The reference sorts D = N_with_dub (tile<<32 | depth_bits, id) pairs with one 64-bit cub::DeviceRadixSort
// (8 passes over 12 B/pair).  An LSD radix sort may process its digits in any grouping as long as every pass is
// stable, and the 32 depth bits do not depend on the tile -- so they are sorted ONCE PER GAUSSIAN (N items)
// instead of once per duplicate (D ~ 5 N items):
//
//   1. cub::DeviceRadixSort (u32 depth bits -> Gaussian index), N items           [low 32 key bits]
//   2. cub::DeviceScan over count[perm[j]] (gather folded into the scan's input iterator) -> slot offsets
//   3. k_emit_tiles: every Gaussian, in depth order, writes (tile id, Gaussian index) for its tile rectangle
//   4. cub::DeviceRadixSort (tile id bits only, stable), D items                   [high key bits]
//   5. k_tile_ranges: start/end per tile (-1 for empty tiles)
//
// Result: exactly the order of the reference's signed 64-bit sort (depth bits compare as unsigned inside a tile,
// negative depths after positive ones), ties (equal tile and depth bits) in Gaussian-index order (the reference's
// tie order is whatever its atomic slot allocation produced).  Traffic: 8 B x 4 passes x N + 6 B x 2 passes x D
// instead of 12 B x 8 passes x D.  Slots come from a prefix sum (no contended global atomic, no memset of D keys),
// scratch from the context arena (the reference does 5x cudaMalloc/cudaFree per call), everything on the
// caller's stream.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "gsb200_common.cuh"
#include "kernels.cuh"

namespace gsb {

__global__ void __launch_bounds__(256)
k_depth_keys(uint32_t N, const float* __restrict__ depth, uint32_t* __restrict__ keys, int32_t* __restrict__ idx) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  keys[i] = __float_as_uint(depth[i]);
  idx[i] = (int32_t)i;
}

// input iterator of the slot scan: duplicate count of the j-th Gaussian in depth order
struct CountInDepthOrder {
  const int32_t* perm;
  const int32_t* count;
  __host__ __device__ __forceinline__ int32_t operator()(int j) const { return count[perm[j]]; }
};

// One warp expands 32 consecutive Gaussians (in depth order).  Their duplicate slots form ONE contiguous range of the
// output; the lanes walk that range 32 slots at a time (fully coalesced stores) and find the owner of each slot with a
// 5-step binary search over the 32 per-lane start offsets (register shuffles).
template <typename KeyT>
__global__ void __launch_bounds__(256)
k_emit_tiles(uint32_t N, const int32_t* __restrict__ perm, const int32_t* __restrict__ count,
             const int32_t* __restrict__ incl_sorted, const ushort4* __restrict__ rect, int tiles_w,
             KeyT* __restrict__ keys, int32_t* __restrict__ vals, int64_t cap) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t base = warp_global * 32;
  if (base >= N) return;
  const uint32_t j = base + lane;
  int c = 0, incl = 0, gid = 0;
  ushort4 r = make_ushort4(0, 0, 0, 0);
  if (j < N) {
    gid = perm[j];
    c = count[gid];
    incl = incl_sorted[j];
    if (c > 0) r = rect[gid];
  }
  // lanes past N inherit the last inclusive offset (zero-length ranges)
  const uint32_t valid = __ballot_sync(0xffffffffu, j < N);
  const int last = 31 - __clz(valid);
  const int incl_last = __shfl_sync(0xffffffffu, incl, last);
  if (j >= N) incl = incl_last;
  const int excl = incl - c;                                   // this Gaussian's first slot
  const int first = __shfl_sync(0xffffffffu, excl, 0);          // warp's first slot
  const int total = incl_last - first;                          // slots owned by the warp
  const int x0w = (int)r.x | ((int)r.y << 16);                  // packed for the shuffles
  const int wq = (int)r.z - (int)r.x + 1;
  for (int s0 = 0; s0 < total; s0 += 32) {  // warp-uniform trip count: every lane takes part in the shuffles
    const int slot = first + s0 + lane;
    // owner = max{ lane : excl_lane <= slot }  (a zero-count lane is shadowed by its successor with the same excl)
    int pos = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
      const int cand = pos + step;  // <= 31
      const int e = __shfl_sync(0xffffffffu, excl, cand);
      if (e <= slot) pos = cand;
    }
    const int eg = __shfl_sync(0xffffffffu, excl, pos);
    const int idg = __shfl_sync(0xffffffffu, gid, pos);
    const int xy = __shfl_sync(0xffffffffu, x0w, pos);
    const int w = __shfl_sync(0xffffffffu, wq, pos);
    if (s0 + lane < total && slot < cap) {  // cap: capacity of keys / vals (exact D on the synchronous path)
      const int k = slot - eg;
      const int ty = (xy >> 16) + k / w, tx = (xy & 0xffff) + k % w;
      keys[slot] = (KeyT)(ty * tiles_w + tx);
      vals[slot] = idg;
    }
  }
}

__global__ void __launch_bounds__(256)
k_fill_i32x2(int32_t* p, int32_t* q, int32_t v, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { p[i] = v; q[i] = v; }
}

// aabb_culling.h:70-103 fill_start_aabb + fill_end_aabb in one pass
template <typename KeyT>
__global__ void __launch_bounds__(256)
k_tile_ranges(int64_t D, const KeyT* __restrict__ keys, int32_t* __restrict__ start, int32_t* __restrict__ end,
              uint32_t T) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= D) return;
  uint32_t t = (uint32_t)keys[s];
  if (t >= T) return;  // padding key of the capacity-sized sort (asynchronous-count mode)
  if (s == 0 || (uint32_t)keys[s - 1] != t) start[t] = (int32_t)s;
  if (s == D - 1 || (uint32_t)keys[s + 1] != t) end[t] = (int32_t)(s + 1);
}

// asynchronous-count mode: keys[D .. cap) := pad (sorts behind every tile); D = incl[N-1] is read on the device.
// Also raises the overflow flag when the duplicates do not fit the capacity (the lists are then truncated and the
// host rejects the view at its next synchronisation point, gsb200_render_backward / gsb200_view_stats).
template <typename KeyT>
__global__ void __launch_bounds__(256)
k_pad_keys(const int32_t* __restrict__ incl_last, KeyT* __restrict__ keys, int64_t cap, KeyT pad,
           int32_t* __restrict__ overflow) {
  const int64_t D = (int64_t)*incl_last;
  if (blockIdx.x == 0 && threadIdx.x == 0) *overflow = (D > cap) ? 1 : 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t s = D + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += stride) keys[s] = pad;
}

// ---- duplicate count read-back ----------------------------------------------------------------------
// The per-Gaussian kernels add their block sums of `count` into ctx->d_total (zeroed by begin_total); the 8-byte
// copy is enqueued right behind them and an event marks it, so the host learns D while the GPU is already
// sorting depths.  (The reference blocks twice per view: gs/culling.py:33-35 .item(), aabb_culling.h:227.)
int begin_total(gsb200_ctx* ctx, cudaStream_t st) {
  int rc = ctx->d_total.reserve(2 * sizeof(unsigned long long));  // [0] duplicates, [1] Gaussians passing the frustum test
  if (rc) return rc;
  GSB_CUDA(cudaMemsetAsync(ctx->d_total.p, 0, 2 * sizeof(unsigned long long), st));
  return GSB200_OK;
}
int request_total(gsb200_ctx* ctx, cudaStream_t st) {
  if (!ctx->ev_total) GSB_CUDA(cudaEventCreateWithFlags(&ctx->ev_total, cudaEventDisableTiming));
  GSB_CUDA(cudaMemcpyAsync(ctx->h_total, ctx->d_total.p, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  GSB_CUDA(cudaEventRecord(ctx->ev_total, st));
  return GSB200_OK;
}
int wait_total(gsb200_ctx* ctx, int64_t* h_total) {
  GSB_CUDA(cudaEventSynchronize(ctx->ev_total));
  *h_total = *ctx->h_total;
  return GSB200_OK;
}

static int bits_for(uint32_t n) {  // ceil(log2(n)) for n >= 1
  int b = 0;
  while ((1u << b) < n) ++b;
  return b;
}

// Stage 1+2 (independent of D): depth order of the Gaussians and their slot offsets.  keys_ready: the fused front end
// already wrote dkeys[0] (depth bits) and perm[0] (identity) -- one launch less.
int sort_depths_and_scan(gsb200_ctx* ctx, uint32_t N, const float* depth, cudaStream_t st, bool keys_ready) {
  if (N == 0) return GSB200_OK;
  int rc;
  if ((rc = reserve_depth_sort(ctx, N))) return rc;
  const unsigned blocks = (N + 255) / 256;
  if (!keys_ready) {
    k_depth_keys<<<blocks, 256, 0, st>>>(N, depth, ctx->dkeys[0].as<uint32_t>(), ctx->perm[0].as<int32_t>());
    GSB_LAUNCH_CHECK();
  }
  auto counts = thrust::make_transform_iterator(
      thrust::counting_iterator<int>(0), CountInDepthOrder{ctx->perm[1].as<int32_t>(), ctx->count.as<int32_t>()});
  size_t b1 = 0, b2 = 0;
  GSB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b1, ctx->dkeys[0].as<uint32_t>(), ctx->dkeys[1].as<uint32_t>(),
                                           ctx->perm[0].as<int32_t>(), ctx->perm[1].as<int32_t>(), (int)N, 0, 32, st));
  GSB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, b2, counts, ctx->incl.as<int32_t>(), (int)N, st));
  if ((rc = ctx->cub_tmp.reserve(b1 > b2 ? b1 : b2))) return rc;
  GSB_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, b1, ctx->dkeys[0].as<uint32_t>(),
                                           ctx->dkeys[1].as<uint32_t>(), ctx->perm[0].as<int32_t>(),
                                           ctx->perm[1].as<int32_t>(), (int)N, 0, 32, st));
  GSB_CUDA(cub::DeviceScan::InclusiveSum(ctx->cub_tmp.p, b2, counts, ctx->incl.as<int32_t>(), (int)N, st));
  return GSB200_OK;
}

int reserve_depth_sort(gsb200_ctx* ctx, uint32_t N) {
  int rc;
  if ((rc = ctx->dkeys[0].reserve((size_t)N * 4))) return rc;
  if ((rc = ctx->dkeys[1].reserve((size_t)N * 4))) return rc;
  if ((rc = ctx->perm[0].reserve((size_t)N * 4))) return rc;
  if ((rc = ctx->perm[1].reserve((size_t)N * 4))) return rc;
  if ((rc = ctx->incl.reserve((size_t)N * 4))) return rc;
  return GSB200_OK;
}

// sorted = exact D (synchronous path) or the capacity (asynchronous-count mode: keys beyond the device-side D are
// padding that sorts behind every tile).
template <typename KeyT>
static int emit_sort_ranges(gsb200_ctx* ctx, uint32_t N, int64_t n_sort, bool padded, int tiles_w, uint32_t T,
                            int32_t* ids_out, int32_t* start, int32_t* end, cudaStream_t st) {
  KeyT* k0 = ctx->keys[0].as<KeyT>();
  KeyT* k1 = ctx->keys[1].as<KeyT>();
  int32_t* v0 = ctx->vals[0].as<int32_t>();
  int32_t* v1 = ids_out ? ids_out : ctx->vals[1].as<int32_t>();
  const uint32_t warps = (N + 31) / 32;
  k_emit_tiles<KeyT><<<(warps * 32 + 255) / 256, 256, 0, st>>>(N, ctx->perm[1].as<int32_t>(),
                                                              ctx->count.as<int32_t>(), ctx->incl.as<int32_t>(),
                                                              ctx->rect.as<ushort4>(), tiles_w, k0, v0, n_sort);
  GSB_LAUNCH_CHECK();
  uint32_t key_range = T;
  if (padded) {
    int rc0;
    if ((rc0 = ctx->d_overflow.reserve(sizeof(int32_t)))) return rc0;
    k_pad_keys<KeyT><<<ctx->sm_count * 2, 256, 0, st>>>(ctx->incl.as<int32_t>() + (N - 1), k0, n_sort, (KeyT)T,
                                                       ctx->d_overflow.as<int32_t>());
    GSB_LAUNCH_CHECK();
    key_range = T + 1;
  }
  const int end_bit = bits_for(key_range) < 1 ? 1 : bits_for(key_range);
  size_t bytes = 0;
  GSB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k0, k1, v0, v1, (int)n_sort, 0, end_bit, st));
  int rc;
  if ((rc = ctx->cub_tmp.reserve(bytes))) return rc;
  GSB_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, bytes, k0, k1, v0, v1, (int)n_sort, 0, end_bit, st));
  ctx->sorted_sel = 1;
  k_tile_ranges<KeyT><<<(unsigned)((n_sort + 255) / 256), 256, 0, st>>>(n_sort, k1, start, end, T);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}

// Stage 3-5.  Synchronous path: D is the exact duplicate count (known on the host).  Asynchronous-count mode
// (padded = true): D is the CAPACITY the host reserved from earlier views; the exact count stays on the device.
// Sorted ids go to ids_out, or stay in ctx->vals[ctx->sorted_sel] when ids_out == nullptr.
int bin_and_sort(gsb200_ctx* ctx, uint32_t N, int64_t D, int tiles_h, int tiles_w, int32_t* ids_out, int32_t* start,
                 int32_t* end, cudaStream_t st, bool padded) {
  const uint32_t T = (uint32_t)tiles_h * (uint32_t)tiles_w;
  if (T) {
    k_fill_i32x2<<<(T + 255) / 256, 256, 0, st>>>(start, end, -1, T);
    GSB_LAUNCH_CHECK();
  }
  if (!padded) ctx->D = D;
  if (D == 0 || N == 0) return GSB200_OK;
  GSB_CHECK(D < (int64_t)2147483647, GSB200_ERR_INVALID, "N_with_dub %lld exceeds int32", (long long)D);
  const bool k16 = (T + (padded ? 1u : 0u)) <= 65536;
  const size_t kb = k16 ? 2 : 4;
  int rc;
  for (int k = 0; k < 2; ++k) {
    if ((rc = ctx->keys[k].reserve((size_t)D * kb))) return rc;
    if (!(k == 1 && ids_out)) {
      if ((rc = ctx->vals[k].reserve((size_t)D * 4))) return rc;
    }
  }
  return k16 ? emit_sort_ranges<uint16_t>(ctx, N, D, padded, tiles_w, T, ids_out, start, end, st)
             : emit_sort_ranges<uint32_t>(ctx, N, D, padded, tiles_w, T, ids_out, start, end, st);
}

// longest tile list of the last view (gsb200_view_stats)
__global__ void __launch_bounds__(256)
k_max_list(uint32_t T, const int32_t* __restrict__ start, const int32_t* __restrict__ end, int32_t* __restrict__ out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  int len = 0;
  if (t < T && start[t] >= 0) len = end[t] - start[t];
  len = __reduce_max_sync(0xffffffffu, len);
  if ((threadIdx.x & 31) == 0 && len > 0) atomicMax(out, len);
}
int launch_max_list(uint32_t T, const int32_t* start, const int32_t* end, int32_t* out, cudaStream_t st) {
  GSB_CUDA(cudaMemsetAsync(out, 0, sizeof(int32_t), st));
  if (T == 0) return GSB200_OK;
  k_max_list<<<(T + 255) / 256, 256, 0, st>>>(T, start, end, out);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}

}  // namespace gsb
