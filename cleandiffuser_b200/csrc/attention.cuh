// CDS_OP_ATTN, fp32 CUDA-core path: softmax(Q K^T / sqrt(hd)) V for one (trajectory, head) per CTA.
// K and V of the head (L x hd each, L <= 128) are staged once in shared memory; each thread owns one query
// row and runs an online softmax over the keys, so the L x L score matrix never exists in memory.
// Algorithmic HBM bytes per (trajectory, head): 4*L*hd*4 (q, k, v read; out written).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace cds {

template <int HD>
__global__ void __launch_bounds__(128) attention_f32_kernel(const cds_attn_op p) {
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;                 // [L][HD]
  float* Vs = smem + p.L * HD;      // [L][HD]
  const int b = blockIdx.x / p.heads, h = blockIdx.x - b * p.heads;
  const int ld = 3 * p.C;
  const float* base = p.qkv + (int64_t)b * p.L * ld + h * HD;
  for (int idx = threadIdx.x; idx < p.L * HD; idx += blockDim.x) {
    int j = idx / HD, d = idx - j * HD;
    Ks[idx] = base[(int64_t)j * ld + p.C + d];
    Vs[idx] = base[(int64_t)j * ld + 2 * p.C + d];
  }
  __syncthreads();
  const float scale = rsqrtf((float)HD);
  for (int i = threadIdx.x; i < p.L; i += blockDim.x) {
    float q[HD], o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) { q[d] = base[(int64_t)i * ld + d] * scale; o[d] = 0.f; }
    float mx = -INFINITY, den = 0.f;
    for (int j = 0; j < p.L; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) s = fmaf(q[d], Ks[j * HD + d], s);
      float nm = fmaxf(mx, s);
      float corr = expf(mx - nm), w = expf(s - nm);
      den = den * corr + w;
#pragma unroll
      for (int d = 0; d < HD; ++d) o[d] = fmaf(o[d], corr, w * Vs[j * HD + d]);
      mx = nm;
    }
    float inv = 1.f / den;
    const int64_t o_off = ((int64_t)b * p.L + i) * p.C + h * HD;
    if (p.out_dtype == CDS_BF16) {
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + o_off;
#pragma unroll
      for (int d = 0; d < HD; ++d) dst[d] = __float2bfloat16_rn(o[d] * inv);
    } else {
      float* dst = reinterpret_cast<float*>(p.out) + o_off;
#pragma unroll
      for (int d = 0; d < HD; ++d) dst[d] = o[d] * inv;
    }
  }
}

inline cudaError_t attention_launch(const cds_attn_op& p, cudaStream_t st) {
  int hd = p.C / p.heads;
  size_t smem = sizeof(float) * 2 * (size_t)p.L * hd;
  dim3 grid(p.batch * p.heads);
#define CDS_ATTN_CASE(HD)                                                                                   \
  case HD: {                                                                                                 \
    static bool set = false;                                                                                 \
    if (!set) { cudaFuncSetAttribute(attention_f32_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     200 * 1024); set = true; }                                              \
    attention_f32_kernel<HD><<<grid, 128, smem, st>>>(p);                                                  \
    break; }
  switch (hd) {
    CDS_ATTN_CASE(16)
    CDS_ATTN_CASE(32)
    CDS_ATTN_CASE(64)
    default: return cudaErrorInvalidValue;
  }
#undef CDS_ATTN_CASE
  return cudaGetLastError();
}

}  // namespace cds
