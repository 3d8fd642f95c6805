// Micro-experiment: can a tcgen05 K-major SWIZZLE_128B operand descriptor START at a row that is not a multiple of 8 (i.e. not
// aligned to the 1024-byte swizzle atom)?  That is what a "halo" activation tile needs: ONE TMA box of L + 2*pad positions per
// chunk, read by the MMA of tap j through a descriptor whose start address is shifted by j rows.
//
//   X : [R][32] fp32 in global memory (R = 272 rows), loaded by ONE TMA box {32, 272} -> smem, 128-byte rows, SWIZZLE_128B
//   W : [32][32] fp32, TMA box {32, 32}
//   for j in 0..4:  D_j[128 x 32] = A_j * W^T   with A_j = rows (g*S + j + r), g = 0..15 groups, r = 0..7, S = slot rows
//   variants: S = 8 (dense rows m + j), S = 16 (16-row slots); descriptor base_offset field = 0 or = j
// Prints the max |D - reference| per (variant, j).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -I../../include -I../../cleandiffuser_b200/csrc
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <vector>

#include "conv_tc.cuh"

namespace cds { long long* conv_tc_trace_hook(int) { return nullptr; } }
using namespace cds;

constexpr int kRows = 272, kN = 32, kK = 32;

__global__ void __launch_bounds__(128) rowshift_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                                                       float* out, int slot_rows, int use_base_offset) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar, done;
  __shared__ uint32_t tmem_holder;
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - ptx::smem_u32(smem_raw));
  uint8_t* sx = sm;                       // 272 x 128 B
  uint8_t* sw = sm + 35 * 1024;           // 32 x 128 B
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::mbar_init(&done, 1); ptx::fence_barrier_init(); }
  if (warp == 0) ptx::tmem_alloc<256>(&tmem_holder);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_holder;
  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(&bar, 256 * 128 + kN * 128);       // (one 256-row box of X: the box limit)
    // box rows are limited to 256: two loads (rows 0..255 and 256..271)
    ptx::tma_load_2d(sx, &tm_x, &bar, 0, 0);
    ptx::tma_load_2d(sw, &tm_w, &bar, 0, 0);
    ptx::mbar_wait(&bar, 0);
    ptx::tc_fence_after_sync();
    constexpr uint32_t idesc = ptx::make_idesc_tf32(128, kN);
    for (int j = 0; j < 5; ++j) {
      const uint32_t a_addr = base + (uint32_t)j * 128u;
      uint64_t da = (uint64_t)((a_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((slot_rows * 128) >> 4) << 32) | ((uint64_t)1 << 46) |
                    ((uint64_t)2 << 61);
      if (use_base_offset) da |= (uint64_t)(j & 7) << 49;
      const uint64_t db = ptx::make_kmajor_desc<128>(base + 35 * 1024);
      for (int k = 0; k < 4; ++k) ptx::umma_tf32(tmem + (uint32_t)(j * kN), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
    }
    ptx::umma_commit(&done);
  }
  ptx::mbar_wait(&done, 0);
  ptx::tc_fence_after_sync();
  for (int j = 0; j < 5; ++j) {
    float v[32];
    ptx::tmem_ld<32>(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(j * kN), v);
    const int m = threadIdx.x;
    for (int n = 0; n < kN; ++n) out[(j * 128 + m) * kN + n] = v[n];
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc<256>(tmem); }
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }

int main() {
  std::vector<float> X(kRows * kK), W(kN * kK);
  for (int i = 0; i < kRows * kK; ++i) X[i] = tf32_trunc(sinf(0.37f * i) + 0.01f * (i % 97));
  for (int i = 0; i < kN * kK; ++i) W[i] = tf32_trunc(cosf(0.11f * i));
  float *dX, *dW, *dO;
  cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dO, 5 * 128 * kN * 4);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap tx, tw;
  { uint64_t dims[2] = {(uint64_t)kK, (uint64_t)kRows}; uint64_t str[1] = {(uint64_t)kK}; uint32_t box[2] = {32u, 256u};
    if (!encode_act_map(&tx, dX, 2, dims, str, box, 64, true)) { printf("encode x failed\n"); return 1; } }
  { uint64_t dims[2] = {(uint64_t)kK, (uint64_t)kN}; uint64_t str[1] = {(uint64_t)kK}; uint32_t box[2] = {32u, 32u};
    if (!encode_act_map(&tw, dW, 2, dims, str, box, 64, true)) { printf("encode w failed\n"); return 1; } }
  cudaFuncSetAttribute(rowshift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  std::vector<float> O(5 * 128 * kN);
  for (int slot = 8; slot <= 16; slot += 8) {
    for (int ubo = 0; ubo < 2; ++ubo) {
      cudaMemset(dO, 0, O.size() * 4);
      rowshift_kernel<<<1, 128, 48 * 1024>>>(tx, tw, dO, slot, ubo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("slot %d base_offset %d: CUDA error %s\n", slot, ubo, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      for (int j = 0; j < 5; ++j) {
        double worst = 0;
        for (int m = 0; m < 128; ++m) {
          const int row = (m / 8) * slot + (m % 8) + j;
          if (row >= 256) continue;                                  // only rows 0..255 were loaded (one 256-row box)
          for (int n = 0; n < kN; ++n) {
            double ref = 0;
            for (int c = 0; c < kK; ++c) ref += (double)X[row * kK + c] * (double)W[n * kK + c];
            worst = fmax(worst, fabs(ref - (double)O[(j * 128 + m) * kN + n]));
          }
        }
        printf("slot_rows %2d base_offset_field %d tap-shift j=%d : max |D - ref| = %.3e %s\n", slot, ubo, j, worst, worst < 1e-3 ? "OK" : "MISMATCH");
      }
    }
  }
  return 0;
}
