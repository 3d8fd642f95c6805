// Elementwise operators of the reverse loop: solver update (CDS_OP_UPDATE), consistency-model
// pre-scale (CDS_OP_PREP) and LayerNorm+modulate (CDS_OP_LNMOD).  All HBM-bound streaming kernels:
// algorithmic bytes per element are listed with each kernel.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace cds {

// IEEE fp32 ops that the compiler may not contract into FMAs: the update mirrors the reference's tensor
// expression operation by operation (diffusionsde.py:539-592), so results agree to fp32 rounding of
// identical operations rather than to "some tolerance".
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_(float a, float b) { return __fdiv_rn(a, b); }

// reads x, pred (+pred_uncond, noise, prior, xhat_prev), writes x (+xhat_prev): 12..28 B / element
// `advance` (or NULL): [0] = the iteration counter, [1] = blocks finished; the last block to finish bumps the counter, which
// saves the one-thread advance kernel at the end of every iteration (every block has read the counter before it signals).
__global__ void __launch_bounds__(256) solver_update_kernel(const cds_update_op p, const int* iter_ptr, int* advance) {
  const int iter = *iter_ptr;
  const float* row = p.coef + (int64_t)iter * CDS_ROW_FLOATS;
  const float alpha = row[CDS_ROW_ALPHA], sigma = row[CDS_ROW_SIGMA];
  const float k0 = row[CDS_ROW_K0], k1 = row[CDS_ROW_K1], k2 = row[CDS_ROW_K2], k3 = row[CDS_ROW_K3], k4 = row[CDS_ROW_K4];
  const int kind = (int)row[CDS_ROW_KIND];
  const int slot = (int)row[CDS_ROW_NOISE] - 1;
  const int64_t total = (int64_t)p.batch * p.row;
  const float* noise = slot >= 0 ? p.noise + (int64_t)slot * total : nullptr;

  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(i % p.row);
    const float x = p.x[i];
    float pr = p.pred[i];
    if (p.pred_uncond) pr = add_(mul_(p.w_cfg, pr), mul_(p.w_uncond, p.pred_uncond[i]));

    float out;
    if (kind == CDS_UPD_CM) {
      // f = c_skip*x + c_out*net ; clip (consistency_model.py:257-261)
      out = add_(mul_(k0, x), mul_(k1, pr));
      if (p.final_clip) {
        if (p.x_min) out = fmaxf(out, p.x_min[e]);
        if (p.x_max) out = fminf(out, p.x_max[e]);
      }
    } else {
      // clip_prediction (diffusionsde.py:208-223)
      if (p.predict_noise) {
        if (p.x_max) pr = fmaxf(pr, div_(sub_(x, mul_(alpha, p.x_max[e])), sigma));
        if (p.x_min) pr = fminf(pr, div_(sub_(x, mul_(alpha, p.x_min[e])), sigma));
      } else {
        if (p.x_min) pr = fmaxf(pr, p.x_min[e]);
        if (p.x_max) pr = fminf(pr, p.x_max[e]);
      }
      float eps, xhat;
      if (p.predict_noise) { eps = pr; xhat = div_(sub_(x, mul_(sigma, pr)), alpha); }
      else { xhat = pr; eps = div_(sub_(x, mul_(alpha, pr)), sigma); }

      if (kind == CDS_UPD_DDPM) {
        out = add_(mul_(k0, sub_(x, mul_(sigma, eps))), mul_(k1, eps));
        if (noise) out = add_(out, mul_(k2, noise[i]));
      } else if (kind == CDS_UPD_DDIM) {
        out = add_(mul_(k0, div_(sub_(x, mul_(sigma, eps)), alpha)), mul_(k1, eps));
      } else {
        float target;
        if (kind == CDS_UPD_EPS) target = eps;
        else if (kind == CDS_UPD_X2M) target = sub_(mul_(k3, xhat), mul_(k4, p.xhat_prev[i]));
        else target = xhat;
        out = sub_(mul_(k0, x), mul_(k1, target));
        if (noise) out = add_(out, mul_(k2, noise[i]));
        if (p.xhat_prev) p.xhat_prev[i] = xhat;
      }
    }
    if (p.mask) { const float m = p.mask[e]; out = add_(mul_(out, 1.f - m), mul_(p.prior[i], m)); }
    p.x[i] = out;
    if (p.x_cast) {
      const int64_t r = i / p.cast_C_in;
      reinterpret_cast<__nv_bfloat16*>(p.x_cast)[r * p.cast_C_out + (i - r * p.cast_C_in)] = __float2bfloat16_rn(out);
    }
  }
  if (advance) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(&advance[1], 1) == (int)gridDim.x - 1) { advance[1] = 0; advance[0] = iter + 1; }
    }
  }
}

// x += K2*z (re-noise, only when the row has a noise slot); xin = K3*x.   8..16 B / element
__global__ void __launch_bounds__(256) cm_prep_kernel(const cds_prep_op p, const int* __restrict__ iter_ptr) {
  const int iter = *iter_ptr;
  const float* row = p.coef + (int64_t)iter * CDS_ROW_FLOATS;
  const float k2 = row[CDS_ROW_K2], k3 = row[CDS_ROW_K3];
  const int slot = (int)row[CDS_ROW_NOISE] - 1;
  const int64_t total = (int64_t)p.batch * p.row;
  const float* noise = slot >= 0 ? p.noise + (int64_t)slot * total : nullptr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float x = p.x[i];
    if (noise) { x = add_(x, mul_(k2, noise[i])); p.x[i] = x; }
    p.xin[i] = mul_(k3, x);
  }
}

// one warp per token row: LayerNorm (biased variance, no affine) then x*(1+scale)+shift.  8 B / element
__global__ void __launch_bounds__(256) ln_modulate_kernel(const cds_lnmod_op p) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t n_rows = (int64_t)p.batch * p.L;
  for (int64_t r = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_rows;
       r += (int64_t)gridDim.x * warps_per_block) {
    const float* src = p.in + r * p.C;
    float s = 0.f;
    for (int c = lane; c < p.C; c += 32) s += src[c];
    const float mean = warp_sum(s) / (float)p.C;
    float q = 0.f;
    for (int c = lane; c < p.C; c += 32) { float d = src[c] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) / (float)p.C + p.eps);
    const int b = (int)(r / p.L);
    const float* sh = p.shift + (int64_t)b * p.mod_bstride;
    const float* sc = p.scale + (int64_t)b * p.mod_bstride;
    float* dst = p.out + r * p.C;
    for (int c = lane; c < p.C; c += 32) dst[c] = fmaf((src[c] - mean) * rstd, 1.f + sc[c], sh[c]);
  }
}

// fp32 (rows, C_in) -> bf16 (rows, C_out) zero padded; one thread per output pair.  4*C_in + 2*C_out B / row
__global__ void __launch_bounds__(256) cast_pad_kernel(const cds_cast_op p) {
  const int64_t rows = (int64_t)p.batch * p.L;
  const int pairs = p.C_out >> 1;
  const int64_t total = rows * pairs;
  __nv_bfloat162* out = reinterpret_cast<__nv_bfloat162*>(p.out);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / pairs;
    const int c = (int)(i - r * pairs) * 2;
    const float a = c < p.C_in ? p.in[r * p.C_in + c] : 0.f;
    const float b = c + 1 < p.C_in ? p.in[r * p.C_in + c + 1] : 0.f;
    out[i] = __floats2bfloat162_rn(a, b);
  }
}

__global__ void set_iter_kernel(int* iter_ptr, int value) { *iter_ptr = value; }
__global__ void advance_iter_kernel(int* iter_ptr) { *iter_ptr += 1; }

}  // namespace cds
