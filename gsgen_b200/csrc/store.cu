// store.cu -- device-side row movers of the Gaussian arena.
//
// Reference behaviour replaced: every densify / prune re-creates each parameter with torch.cat / boolean-mask
// indexing and performs the same surgery on both Adam moments of its param group
// (gs/gaussian_splatting.py:421-449 prune_optimizer, :481-522 densify_on_optimizer, :528-549 prune_by_mask) --
// ~10 allocations + copies of every tensor per operation and a host sync per boolean index.
//
// Here the four flat buffers (parameters, gradients, exp_avg, exp_avg_sq; field-major, `capacity` rows per field,
// gsgen_b200/store.py) are compacted by ONE launch: a prefix sum of the keep mask gives every surviving row its
// destination, and one kernel moves all fields of all buffers (stable: order of the survivors kept) into the arena's
// shadow buffers, zero-filling the rows that die.  Appending clone / split children is one launch too: the children's
// parameter rows are copied behind row N of every field and their gradient / moment rows are zeroed (the reference
// concatenates zeros_like, :497-505).  Both kernels are pure HBM streams: 2 x 4 B per moved float.
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "gsb200_common.cuh"
#include "kernels.cuh"

namespace gsb {

constexpr int kMaxFields = 8;
constexpr int kMaxBufs = 4;

struct MoverArgs {
  const float* src[kMaxBufs];
  float* dst[kMaxBufs];
  unsigned long long off[kMaxFields];  // first float of the field in a buffer
  uint32_t width[kMaxFields];          // floats per row
  int n_bufs, n_fields;
  uint32_t N;                          // live rows before the operation
  uint32_t zero_upto;                  // rows [n_keep, zero_upto) of dst are zero-filled
};

__global__ void __launch_bounds__(256)
k_keep_flags(uint32_t N, const uint8_t* __restrict__ remove, int32_t* __restrict__ keep) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) keep[i] = remove[i] ? 0 : 1;
}

// grid: x over elements of a field (grid-stride), y = field, z = buffer
__global__ void __launch_bounds__(256)
k_compact_rows(const MoverArgs a, const uint8_t* __restrict__ remove, const int32_t* __restrict__ excl,
               const int32_t* __restrict__ n_keep_ptr) {
  const int f = blockIdx.y, b = blockIdx.z;
  const uint32_t w = a.width[f];
  const float* __restrict__ src = a.src[b] + a.off[f];
  float* __restrict__ dst = a.dst[b] + a.off[f];
  const unsigned long long n_el = (unsigned long long)a.N * w;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const unsigned long long t0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (unsigned long long j = t0; j < n_el; j += stride) {
    const uint32_t i = (uint32_t)(j / w), e = (uint32_t)(j - (unsigned long long)i * w);
    if (!remove[i]) dst[(unsigned long long)excl[i] * w + e] = src[j];
  }
  const unsigned long long z0 = (unsigned long long)(*n_keep_ptr) * w, z1 = (unsigned long long)a.zero_upto * w;
  for (unsigned long long j = z0 + t0; j < z1; j += stride) dst[j] = 0.f;
}

struct AppendArgs {
  float* dst[kMaxBufs];                // dst[0] = parameters (receives the rows), others are zero-filled
  const float* rows[kMaxFields];       // [k, width_f] new parameter rows of field f
  unsigned long long off[kMaxFields];
  uint32_t width[kMaxFields];
  int n_bufs, n_fields;
  uint32_t N, k;
};

__global__ void __launch_bounds__(256)
k_append_rows(const AppendArgs a) {
  const int f = blockIdx.y, b = blockIdx.z;
  const uint32_t w = a.width[f];
  float* __restrict__ dst = a.dst[b] + a.off[f] + (unsigned long long)a.N * w;
  const float* __restrict__ src = a.rows[f];
  const unsigned long long n_el = (unsigned long long)a.k * w;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < n_el; j += stride)
    dst[j] = (b == 0) ? src[j] : 0.f;
}

// ---- sparse gradient all-reduce: pack / unpack of the rows some view touched ---------------------
// T < T_thresh hides most of a scene from any one view (measured with the oracle: a view sends gradients to 10 % of
// C3's Gaussians, the 8 views of a step to 36 %), so the flat gradient buffer the ranks all-reduce is mostly zeros.
// pack: rows with keep[i] != 0 of every field are gathered into a TIGHT buffer -- field f starts at
// n_keep * (sum of the widths before f) -- so the ranks (which hold the same OR-reduced mask) all-reduce n_keep * 59
// floats with ONE collective; unpack scatters the reduced rows back.
struct KeepFlag {
  const uint8_t* m;
  __host__ __device__ __forceinline__ int32_t operator()(int i) const { return m[i] ? 1 : 0; }
};

struct PackArgs {
  unsigned long long off[kMaxFields];  // field offsets in the flat buffer
  unsigned long long poff[kMaxFields]; // field offsets in the packed buffer (n_keep based)
  uint32_t width[kMaxFields];
  uint32_t N;
};

__global__ void __launch_bounds__(256)
k_kept_ids(uint32_t N, const uint8_t* __restrict__ keep, const int32_t* __restrict__ excl, int32_t* __restrict__ idx) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N && keep[i]) idx[excl[i]] = (int32_t)i;
}

// grid: x over the elements of the KEPT rows of a field (grid-stride), y = field.  Packed side coalesced; flat side
// whole rows (contiguous width_f floats) at the kept row's position.
template <bool PACK>
__global__ void __launch_bounds__(256)
k_pack_rows(const PackArgs a, float* __restrict__ flat, float* __restrict__ packed, const int32_t* __restrict__ idx,
            uint32_t n_keep) {
  const int f = blockIdx.y;
  const uint32_t w = a.width[f];
  float* __restrict__ fl = flat + a.off[f];
  float* __restrict__ pk = packed + a.poff[f];
  const unsigned long long n_el = (unsigned long long)n_keep * w;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < n_el; j += stride) {
    const uint32_t r = (uint32_t)(j / w), e = (uint32_t)(j - (unsigned long long)r * w);
    const unsigned long long src = (unsigned long long)idx[r] * w + e;
    if (PACK) pk[j] = fl[src]; else fl[src] = pk[j];
  }
}

}  // namespace gsb

using namespace gsb;

extern "C" {

int gsb200_rows_pack(gsb200_ctx* ctx, const float* flat, float* packed, uint64_t packed_capacity,
                     const uint64_t* h_field_off, const uint32_t* h_field_width, int32_t n_fields, uint32_t N,
                     const uint8_t* keep, int32_t* idx, uint32_t* h_n_keep, gsb200_stream stream) {
  GSB_CHECK(ctx != nullptr, GSB200_ERR_INVALID, "null context");
  GSB_CUDA(cudaSetDevice(ctx->device));
  GSB_CHECK(flat && packed && h_field_off && h_field_width && keep && idx && h_n_keep, GSB200_ERR_INVALID,
            "rows_pack: null argument");
  GSB_CHECK(n_fields >= 1 && n_fields <= kMaxFields, GSB200_ERR_INVALID, "rows_pack: n_fields=%d (1..8)", n_fields);
  GSB_CHECK(N < 2147483647u, GSB200_ERR_INVALID, "rows_pack: N exceeds int32");
  cudaStream_t st = (cudaStream_t)stream;
  *h_n_keep = 0;
  if (N == 0) return GSB200_OK;
  int rc;
  if ((rc = ctx->incl.reserve((size_t)N * 4))) return rc;
  int32_t* excl = ctx->incl.as<int32_t>();
  auto flags = thrust::make_transform_iterator(thrust::counting_iterator<int>(0), KeepFlag{keep});
  size_t bytes = 0;
  GSB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, flags, excl, (int)N, st));
  if ((rc = ctx->cub_tmp.reserve(bytes))) return rc;
  GSB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, bytes, flags, excl, (int)N, st));
  k_kept_ids<<<(N + 255) / 256, 256, 0, st>>>(N, keep, excl, idx);
  GSB_LAUNCH_CHECK();
  int32_t h_last = 0;
  uint8_t h_flag = 0;
  GSB_CUDA(cudaMemcpyAsync(&h_last, excl + (N - 1), 4, cudaMemcpyDeviceToHost, st));
  GSB_CUDA(cudaMemcpyAsync(&h_flag, keep + (N - 1), 1, cudaMemcpyDeviceToHost, st));
  GSB_CUDA(cudaStreamSynchronize(st));  // the collective's element count is a host quantity: the one sync
  const uint32_t n_keep = (uint32_t)h_last + (h_flag ? 1u : 0u);
  *h_n_keep = n_keep;
  PackArgs a;
  memset(&a, 0, sizeof(a));
  unsigned long long row = 0, widest = 1;
  for (int f = 0; f < n_fields; ++f) {
    a.off[f] = h_field_off[f]; a.width[f] = h_field_width[f];
    a.poff[f] = (unsigned long long)n_keep * row;
    row += h_field_width[f];
    widest = a.width[f] > widest ? a.width[f] : widest;
  }
  a.N = N;
  if ((unsigned long long)n_keep * row > packed_capacity) return GSB200_OK;  // does not fit: the caller reduces densely
  if (n_keep == 0) return GSB200_OK;
  unsigned long long blocks = ((unsigned long long)n_keep * widest + 255) / 256;
  const unsigned long long cap_blocks = (unsigned long long)ctx->sm_count * 8;
  if (blocks > cap_blocks) blocks = cap_blocks;
  dim3 grid((unsigned)blocks, (unsigned)n_fields, 1);
  k_pack_rows<true><<<grid, 256, 0, st>>>(a, const_cast<float*>(flat), packed, idx, n_keep);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}

int gsb200_rows_unpack(gsb200_ctx* ctx, float* flat, const float* packed, const uint64_t* h_field_off,
                       const uint32_t* h_field_width, int32_t n_fields, uint32_t N, const int32_t* idx,
                       uint32_t n_keep, gsb200_stream stream) {
  GSB_CHECK(ctx != nullptr, GSB200_ERR_INVALID, "null context");
  GSB_CUDA(cudaSetDevice(ctx->device));
  GSB_CHECK(flat && packed && h_field_off && h_field_width && idx, GSB200_ERR_INVALID, "rows_unpack: null argument");
  GSB_CHECK(n_fields >= 1 && n_fields <= kMaxFields, GSB200_ERR_INVALID, "rows_unpack: n_fields=%d (1..8)", n_fields);
  if (N == 0 || n_keep == 0) return GSB200_OK;
  PackArgs a;
  memset(&a, 0, sizeof(a));
  unsigned long long row = 0, widest = 1;
  for (int f = 0; f < n_fields; ++f) {
    a.off[f] = h_field_off[f]; a.width[f] = h_field_width[f];
    a.poff[f] = (unsigned long long)n_keep * row;
    row += h_field_width[f];
    widest = a.width[f] > widest ? a.width[f] : widest;
  }
  a.N = N;
  unsigned long long blocks = ((unsigned long long)n_keep * widest + 255) / 256;
  const unsigned long long cap_blocks = (unsigned long long)ctx->sm_count * 8;
  if (blocks > cap_blocks) blocks = cap_blocks;
  dim3 grid((unsigned)blocks, (unsigned)n_fields, 1);
  k_pack_rows<false><<<grid, 256, 0, (cudaStream_t)stream>>>(a, flat, const_cast<float*>(packed), idx, n_keep);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}

int gsb200_store_compact(gsb200_ctx* ctx, const float* const* h_src, float* const* h_dst, int32_t n_bufs,
                         const uint64_t* h_field_off, const uint32_t* h_field_width, int32_t n_fields, uint32_t N,
                         uint32_t zero_upto, const uint8_t* remove_mask, uint32_t* h_n_keep, gsb200_stream stream) {
  GSB_CHECK(ctx != nullptr, GSB200_ERR_INVALID, "null context");
  GSB_CUDA(cudaSetDevice(ctx->device));
  GSB_CHECK(h_src && h_dst && h_field_off && h_field_width && h_n_keep, GSB200_ERR_INVALID, "store_compact: null argument");
  GSB_CHECK(n_bufs >= 1 && n_bufs <= kMaxBufs && n_fields >= 1 && n_fields <= kMaxFields, GSB200_ERR_INVALID,
            "store_compact: n_bufs=%d (1..4), n_fields=%d (1..8)", n_bufs, n_fields);
  GSB_CHECK(N < 2147483647u, GSB200_ERR_INVALID, "store_compact: N exceeds int32");
  cudaStream_t st = (cudaStream_t)stream;
  *h_n_keep = 0;
  if (N == 0) return GSB200_OK;
  GSB_CHECK(remove_mask != nullptr, GSB200_ERR_INVALID, "store_compact: null mask");
  MoverArgs a;
  memset(&a, 0, sizeof(a));
  for (int b = 0; b < n_bufs; ++b) {
    GSB_CHECK(h_src[b] && h_dst[b] && h_src[b] != h_dst[b], GSB200_ERR_INVALID,
              "store_compact: buffer %d needs distinct source and destination (the compaction is out of place)", b);
    a.src[b] = h_src[b]; a.dst[b] = h_dst[b];
  }
  for (int f = 0; f < n_fields; ++f) { a.off[f] = h_field_off[f]; a.width[f] = h_field_width[f]; }
  a.n_bufs = n_bufs; a.n_fields = n_fields; a.N = N; a.zero_upto = zero_upto > N ? zero_upto : N;
  // keep flags -> exclusive prefix sum (destination rows); the total is the new row count
  int rc;
  if ((rc = ctx->incl.reserve((size_t)(N + 1) * 4))) return rc;       // excl[N] + total
  if ((rc = ctx->count.reserve((size_t)N * 4))) return rc;              // keep flags
  int32_t* keep = ctx->count.as<int32_t>();
  int32_t* excl = ctx->incl.as<int32_t>();
  k_keep_flags<<<(N + 255) / 256, 256, 0, st>>>(N, remove_mask, keep);
  GSB_LAUNCH_CHECK();
  size_t bytes = 0;
  GSB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, keep, excl, (int)N, st));
  if ((rc = ctx->cub_tmp.reserve(bytes))) return rc;
  GSB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, bytes, keep, excl, (int)N, st));
  // n_keep = excl[N-1] + keep[N-1], formed on the device for the zero-fill and copied to the host for the new N
  if ((rc = ctx->d_small.reserve(4 * sizeof(int32_t)))) return rc;
  int32_t* d_n = ctx->d_small.as<int32_t>();
  GSB_CUDA(cudaMemcpyAsync(d_n, excl + (N - 1), 4, cudaMemcpyDeviceToDevice, st));
  GSB_CUDA(cudaMemcpyAsync(d_n + 1, keep + (N - 1), 4, cudaMemcpyDeviceToDevice, st));
  int32_t h2[2] = {0, 0};
  GSB_CUDA(cudaMemcpyAsync(h2, d_n, 8, cudaMemcpyDeviceToHost, st));
  GSB_CUDA(cudaStreamSynchronize(st));  // the new row count is a host quantity (tensor shapes): the one sync
  const uint32_t n_keep = (uint32_t)(h2[0] + h2[1]);
  *h_n_keep = n_keep;
  int32_t hn = (int32_t)n_keep;
  GSB_CUDA(cudaMemcpyAsync(d_n, &hn, 4, cudaMemcpyHostToDevice, st));
  unsigned long long widest = 1;
  for (int f = 0; f < n_fields; ++f) widest = a.width[f] > widest ? a.width[f] : widest;
  unsigned long long blocks = ((unsigned long long)a.zero_upto * widest + 255) / 256;
  const unsigned long long cap_blocks = (unsigned long long)ctx->sm_count * 8;
  if (blocks > cap_blocks) blocks = cap_blocks;
  if (blocks < 1) blocks = 1;
  dim3 grid((unsigned)blocks, (unsigned)n_fields, (unsigned)n_bufs);
  k_compact_rows<<<grid, 256, 0, st>>>(a, remove_mask, excl, d_n);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}

int gsb200_store_append(gsb200_ctx* ctx, float* const* h_dst, int32_t n_bufs, const float* const* h_rows,
                        const uint64_t* h_field_off, const uint32_t* h_field_width, int32_t n_fields, uint32_t N,
                        uint32_t k, gsb200_stream stream) {
  GSB_CHECK(ctx != nullptr, GSB200_ERR_INVALID, "null context");
  GSB_CUDA(cudaSetDevice(ctx->device));
  GSB_CHECK(h_dst && h_rows && h_field_off && h_field_width, GSB200_ERR_INVALID, "store_append: null argument");
  GSB_CHECK(n_bufs >= 1 && n_bufs <= kMaxBufs && n_fields >= 1 && n_fields <= kMaxFields, GSB200_ERR_INVALID,
            "store_append: n_bufs=%d (1..4), n_fields=%d (1..8)", n_bufs, n_fields);
  if (k == 0) return GSB200_OK;
  AppendArgs a;
  memset(&a, 0, sizeof(a));
  for (int b = 0; b < n_bufs; ++b) { GSB_CHECK(h_dst[b], GSB200_ERR_INVALID, "store_append: null buffer"); a.dst[b] = h_dst[b]; }
  unsigned long long widest = 1;
  for (int f = 0; f < n_fields; ++f) {
    GSB_CHECK(h_rows[f], GSB200_ERR_INVALID, "store_append: null rows for field %d", f);
    a.rows[f] = h_rows[f]; a.off[f] = h_field_off[f]; a.width[f] = h_field_width[f];
    widest = a.width[f] > widest ? a.width[f] : widest;
  }
  a.n_bufs = n_bufs; a.n_fields = n_fields; a.N = N; a.k = k;
  unsigned long long blocks = ((unsigned long long)k * widest + 255) / 256;
  const unsigned long long cap_blocks = (unsigned long long)ctx->sm_count * 8;
  if (blocks > cap_blocks) blocks = cap_blocks;
  dim3 grid((unsigned)blocks, (unsigned)n_fields, (unsigned)n_bufs);
  k_append_rows<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}

}  // extern "C"
