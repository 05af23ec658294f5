// optimizer.cu -- Adam on the flat parameter / gradient buffers.
//
// Reference: torch.optim.Adam built by GaussianSplattingRenderer.set_optimizer (gs/gaussian_splatting.py:398-419)
// with conf/base.yaml:8-11 (eps 1e-15) -- one param group per field, five foreach passes per group.  Here the fields
// live back to back in ONE fp32 buffer (the all-reduce operand's twin, gsgen_b200/parallel.py), so the whole update is
// one HBM-bound streaming kernel: 16 B read + 12 B written per parameter, 16-byte accesses, grid-stride over
// 148 SMs x 8 CTAs.
#include "gsb200_common.cuh"
#include "kernels.cuh"

namespace gsb {

__device__ __forceinline__ int field_of(const AdamFields& F, unsigned long long i) {
  int k = 0;
#pragma unroll
  for (int f = 1; f < 8; ++f)
    if (f < F.n && i >= F.begin[f]) k = f;
  return k;
}
__device__ __forceinline__ float field_step_size(const AdamFields& F, unsigned long long i) {
  return F.step_size[field_of(F, i)];
}

__global__ void __launch_bounds__(256)
k_adam_flat(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ exp_avg,
            float* __restrict__ exp_avg_sq, unsigned long long total, const AdamFields F, const AdamScalars K) {
  const unsigned long long n4 = total >> 2;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
    const unsigned long long i = q << 2;
    float4 p = reinterpret_cast<float4*>(param)[q];
    const float4 g = reinterpret_cast<const float4*>(grad)[q];
    float4 m = reinterpret_cast<float4*>(exp_avg)[q];
    float4 v = reinterpret_cast<float4*>(exp_avg_sq)[q];
    const int f0 = field_of(F, i), f3 = field_of(F, i + 3);
    const float s0 = F.step_size[f0], s3 = F.step_size[f3];
    float s1 = s0, s2 = s0;
    // straddles a field boundary: compare FIELD INDICES (equal step sizes of the outer fields say nothing about a
    // one- or two-element field in between; round-1 advice)
    if (f0 != f3) { s1 = field_step_size(F, i + 1); s2 = field_step_size(F, i + 2); }
    adam_update(p.x, g.x, m.x, v.x, s0, K);
    adam_update(p.y, g.y, m.y, v.y, s1, K);
    adam_update(p.z, g.z, m.z, v.z, s2, K);
    adam_update(p.w, g.w, m.w, v.w, s3, K);
    reinterpret_cast<float4*>(param)[q] = p;
    reinterpret_cast<float4*>(exp_avg)[q] = m;
    reinterpret_cast<float4*>(exp_avg_sq)[q] = v;
  }
  // tail (total % 4 elements): the first threads of block 0
  const unsigned long long t0 = n4 << 2;
  if (blockIdx.x == 0 && t0 + threadIdx.x < total) {
    const unsigned long long i = t0 + threadIdx.x;
    float p = param[i], m = exp_avg[i], v = exp_avg_sq[i];
    adam_update(p, grad[i], m, v, field_step_size(F, i), K);
    param[i] = p; exp_avg[i] = m; exp_avg_sq[i] = v;
  }
}

int launch_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, unsigned long long total,
                     const AdamFields& fields, const AdamScalars& k, cudaStream_t st) {
  if (total == 0) return GSB200_OK;
  const unsigned long long n4 = total >> 2;
  unsigned long long blocks = (n4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148ull * 8ull) blocks = 148ull * 8ull;  // persistent-style grid-stride: 8 CTAs of 256 per SM
  k_adam_flat<<<(unsigned)blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, total, fields, k);
  GSB_LAUNCH_CHECK();
  return GSB200_OK;
}

}  // namespace gsb
