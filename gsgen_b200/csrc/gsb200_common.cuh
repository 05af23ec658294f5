// gsb200_common.cuh -- error handling, the per-view context (scratch arena) and the sm_100a PTX
// wrappers (cp.async / cp.async.bulk + mbarrier staging, vector reductions) shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gsb200.h"
#include "gsb200_math.cuh"

namespace gsb {

// ---- error plumbing (C ABI returns int, message via gsb200_last_error) ---------------------------
void set_error(const char* fmt, ...);
#define GSB_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      gsb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));   \
      return GSB200_ERR_CUDA;                                                                   \
    }                                                                                           \
  } while (0)
#define GSB_CHECK(cond, code, ...)     \
  do {                                 \
    if (!(cond)) {                     \
      gsb::set_error(__VA_ARGS__);     \
      return code;                     \
    }                                  \
  } while (0)
#define GSB_LAUNCH_CHECK() GSB_CUDA(cudaGetLastError())

// ---- grow-only device buffer ---------------------------------------------------------------------
struct Buf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {  // returns GSB200_* code
    if (bytes <= cap) return GSB200_OK;
    if (p) {
      GSB_CUDA(cudaFree(p));  // cudaFree synchronises: growth is rare (arena grows geometrically)
      p = nullptr; cap = 0;
    }
    size_t want = bytes + bytes / 4 + 256;
    GSB_CUDA(cudaMalloc(&p, want));
    cap = want;
    return GSB200_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace gsb

// One context = one device + the saved state of one view (records, sorted lists) between the
// forward and backward calls, plus scratch for the reference-compatible ops.
struct gsb200_ctx {
  int device = 0;
  int sm_count = 148;
  // per-Gaussian
  gsb::Buf splat;     // Splat[N]            32 B  (geometry record)
  gsb::Buf pay;       // float4[N]           16 B  (r,g,b,depth) or (scalar,-,-,-)
  gsb::Buf rect;      // ushort4[N]           8 B  tile rectangle
  gsb::Buf count;     // int32[N]
  gsb::Buf incl;      // int32[N]            inclusive scan of count (in depth order)
  gsb::Buf dkeys[2];  // uint32[N]           depth bits (radix sort double buffer)
  gsb::Buf perm[2];   // int32[N]            Gaussian indices in depth order
  gsb::Buf ggeom;     // float[N*8]          gradient record (gmx,gmy,gxx,gxy | gyy,galpha,gdepth,-)
  gsb::Buf gpay;      // float[N*4]          (gr,gg,gb,-)
  // per-duplicate
  gsb::Buf keys[2];   // uint64[D]
  gsb::Buf vals[2];   // int32[D]
  gsb::Buf cub_tmp;
  // per-tile
  gsb::Buf start, end;  // int32[T]
  // host-visible scalars
  int64_t* h_total = nullptr;  // pinned [4]: duplicates, visible Gaussians, overflow flag, max list length
  gsb::Buf d_total;            // unsigned long long[2] device-side counters (duplicates, visible Gaussians)
  gsb::Buf d_small;            // int32[4] scratch scalars of the store movers (store.cu)
  gsb::Buf d_overflow;         // int32[2]: [0] tile-list capacity overflow flag (async-count mode), [1] max list length
  cudaEvent_t ev_total = nullptr;
  // options (gsb200_ctx_set_option)
  int bwd_sh_variant = 0;      // 0: direct vector-reduction flush (composite_bwd_sh.cu); 1: round-1 shared accumulator
  int async_count = 0;         // 1: render_forward does not wait for N_with_dub (capacity from earlier views)
  // saved view state
  uint32_t N = 0;
  int64_t D = 0;               // exact duplicate count of the last view, -1 while unresolved (async-count mode)
  int64_t dup_capacity = 0;    // async-count mode: entries the tile sort covered (>= D unless overflow)
  int64_t dup_seen = 0;        // largest N_with_dub this context has seen for the current (N, image size)
  uint32_t seen_N = 0; int seen_W = 0, seen_H = 0;
  // asynchronous-count mode: counts in flight (FIFO ring of pinned slots + events; consumed by non-blocking polls)
  static constexpr int kRing = 64;
  int64_t* h_ring = nullptr;           // pinned [kRing][2]: duplicates, visible Gaussians
  cudaEvent_t ev_ring[kRing] = {};
  int64_t ring_gen[kRing] = {};        // generation of the forward that owns the slot
  int64_t ring_cap[kRing] = {};        // tile-list capacity that forward's sort covered
  unsigned long long ring_head = 0, ring_tail = 0;
  int64_t overflow_gen = 0;            // first generation found to have overflowed and not reported yet (0 = none)
  int64_t overflow_dup = 0, overflow_cap = 0;
  int64_t generation = 0;      // stamp of the forward whose state the context holds (0 = none)
  int64_t N_visible = -1;
  int sorted_sel = 0;          // which vals[] holds the sorted ids
  gsb::Camera cam;
  int mode = 0, C = 0;
  // stage profiling (bench only)
  int profiling = 0;
  struct EvSet { cudaEvent_t e[5]; };
  EvSet* fwd_sets = nullptr; int fwd_cap = 0, fwd_used = 0;
  EvSet* bwd_sets = nullptr; int bwd_cap = 0, bwd_used = 0;
  gsb::Buf d_stats;            // unsigned long long[2]: D_eff, staged
  // K-nearest-neighbour search (knn.cu): cell keys / point order (radix sort double buffers), points in cell order,
  // cell table, statistics + grid descriptor
  gsb::Buf knn_keys[2], knn_vals[2], knn_pts, knn_cells, knn_small;
  int64_t sum_dup = 0;
};

namespace gsb {

// ---- PTX wrappers -----------------------------------------------------------------------------------
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Ampere-style async copy global->shared (LDGSTS), 16 B, L2-only (.cg)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// mbarrier (transaction barrier for the bulk/TMA copies)
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity));
}
// 1-D bulk async copy global->shared through the TMA engine (UBLKCP), completion on an mbarrier.
// src / dst 16 B aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem)),
               "l"(gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// vector float reductions to global memory (sm_90+): one L2 atomic transaction for 2 / 4 floats
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};\n" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add(float* addr, float a) {  // (atomicAdd on a pointer of unknown address space
  asm volatile("red.global.add.f32 [%0], %1;\n" ::"l"(addr), "f"(a) : "memory");  // compiles to generic ATOM + QSPC)
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
// Packed FP32x2 FMA (Blackwell FFMA2, PTX fma.rn.f32x2): (dx,dy) += (ax,ay)*(bx,by) in ONE issue slot.  Measured on
// B200 (tools/microbench/ffma2.cu): the FMA pipe still retires 128 lane-FMAs/clk/SM, but an FFMA2 costs one issue
// slot for two FMAs -- and these kernels are issue-slot bound (LDS / MUFU / ALU compete with the FMAs).
__device__ __forceinline__ void ffma2(float& dx, float& dy, float ax, float ay, float bx, float by) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(ax), "f"(ay));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(bx), "f"(by));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(dx), "f"(dy));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(dx), "=f"(dy) : "l"(rd));
}

// sigmoid as the reference's 1/(1+expf(-x)) (shencoder.h:4), MUFU ex2 + rcp
__device__ __forceinline__ float sigmoid_fast(float x) {
  return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
}

// Halving butterfly: reduces K per-lane values across the 32 lanes with K-ish shuffles instead of 5K.
// On return lane l holds, in v[0..max(K/32,1)), the warp-wide sum of element index red_index<K>(l)+i.
template <int N, int STEP> struct HalvingReduce {
  static __device__ __forceinline__ void run(float* v, int lane) {
    if constexpr (STEP >= 1) {
      if constexpr (N > 1) {
        constexpr int H = N / 2;
        const bool up = (lane & STEP) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
          float send = up ? v[i] : v[i + H];
          float keep = up ? v[i + H] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, STEP);
        }
        HalvingReduce<H, STEP / 2>::run(v, lane);
      } else {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], STEP);
        HalvingReduce<1, STEP / 2>::run(v, lane);
      }
    }
  }
};
template <int K> __device__ __forceinline__ void warp_reduce_halving(float (&v)[K], int lane) {
  static_assert(K >= 2 && K <= 64 && (K & (K - 1)) == 0, "K must be a power of two in [2,64]");
  HalvingReduce<K, 16>::run(v, lane);
}
// element index held by `lane` after warp_reduce_halving<K>; *writer tells whether this lane is the
// designated writer of its element(s) (elements are replicated over the lanes that differ in the low bits)
template <int K> __device__ __forceinline__ int red_index(int lane, bool* writer) {
  int e = 0, n = K, step = 16;
  uint32_t low = 0;
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    if (n > 1) {
      n >>= 1;
      if (lane & step) e += n;
    } else {
      low |= (uint32_t)step;
    }
    step >>= 1;
  }
  *writer = (lane & low) == 0;
  return e;  // for K == 64, n == 2 remains: lane holds e and e+1
}
#endif  // __CUDACC__

}  // namespace gsb
