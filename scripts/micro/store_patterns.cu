// Micro-benchmark: how should an epilogue whose threads each own ONE ROW of a 128-row x 320-column fp32 tile (the TMEM accumulator
// layout: lane = row) get that tile to global memory?  148 persistent CTAs x 8 warps (TMEM lane quarter q = warp & 3, column half
// hh = warp >> 2, like the conv / linear_ln epilogues), T tiles per CTA, values synthesised in registers (no loads, no MMA).
//
//   v0  per warp: 32 rows x 16 cols staged (2 KB, SWIZZLE_64B), fence.proxy.async, one TMA store per 16 columns, one buffer
//   v1  v0 with two staging buffers per warp (wait_group.read 1)
//   v2  per warp: 32 rows x 32 cols staged (4 KB, SWIZZLE_128B), one TMA store per 32 columns, two buffers
//   v3  per column half: all four quarter-warps stage 128 rows x 32 cols (16 KB, SWIZZLE_128B), named barrier, ONE thread
//       stores the 16 KB box; two buffers per half
//   v4  per warp: 32 x 32 staged (XOR-swizzled by hand), then COALESCED st.global.v4: 8 lanes cover one 128-byte row segment
//   v5  direct st.global.v4 from registers, thread = row (what the round-1 epilogue did)
//   v6  v4 without staging the data twice: lanes exchange through shuffles? -- not built
// Prints microseconds per variant and the implied write bandwidth.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../include -I../../cleandiffuser_b200/csrc store_patterns.cu -o store_patterns -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

#include "conv_tc.cuh"

namespace cds { long long* conv_tc_trace_hook(int) { return nullptr; } }
using namespace cds;

constexpr int kC = 320, kNH = 160, kTilesPerCta = 20, kCtas = 148;
constexpr int kRows = kCtas * kTilesPerCta * 128;

__device__ __forceinline__ float synth(int row, int col) { return (float)(row & 1023) * 0.5f + (float)col * 0.0078125f; }

template <int V>
__global__ void __launch_bounds__(256, 1) store_kernel(const __grid_constant__ CUtensorMap tm16, const __grid_constant__ CUtensorMap tm32,
                                                       const __grid_constant__ CUtensorMap tm32x128, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - ptx::smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, hh = warp >> 2;
  uint32_t nst = 0;
  for (int t = 0; t < kTilesPerCta; ++t) {
    const int tile = blockIdx.x + t * gridDim.x;
    const int row = tile * 128 + 32 * q + lane;
    if constexpr (V == 0 || V == 1) {
      uint8_t* stg = sm + warp * 4096;
      const int sw = (lane >> 1) & 3;
      for (int ch = 0; ch < kNH / 16; ++ch) {
        const int c0 = hh * kNH + 16 * ch;
        uint8_t* buf = stg + (V == 1 ? (nst & 1u) * 2048 : 0);
        if (lane == 0) { if (V == 1) ptx::bulk_wait_group_read<1>(); else ptx::bulk_wait_group_read<0>(); }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          *reinterpret_cast<float4*>(buf + lane * 64 + ((k ^ sw) << 4)) =
              make_float4(synth(row, c0 + 4 * k), synth(row, c0 + 4 * k + 1), synth(row, c0 + 4 * k + 2), synth(row, c0 + 4 * k + 3));
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) { ptx::tma_store_2d(&tm16, buf, c0, tile * 128 + 32 * q); ptx::bulk_commit_group(); }
        ++nst;
      }
    } else if constexpr (V == 2) {
      uint8_t* stg = sm + warp * 8192;
      const int sw = lane & 7;
      for (int ch = 0; ch < kNH / 32; ++ch) {
        const int c0 = hh * kNH + 32 * ch;
        uint8_t* buf = stg + (nst & 1u) * 4096;
        if (lane == 0) ptx::bulk_wait_group_read<1>();
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<float4*>(buf + lane * 128 + ((k ^ sw) << 4)) =
              make_float4(synth(row, c0 + 4 * k), synth(row, c0 + 4 * k + 1), synth(row, c0 + 4 * k + 2), synth(row, c0 + 4 * k + 3));
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) { ptx::tma_store_2d(&tm32, buf, c0, tile * 128 + 32 * q); ptx::bulk_commit_group(); }
        ++nst;
      }
    } else if constexpr (V == 3) {
      uint8_t* stg = sm + hh * 32768;                                   // two 16 KB buffers per column half
      const int r = 32 * q + lane, sw = r & 7;
      for (int ch = 0; ch < kNH / 32; ++ch) {
        const int c0 = hh * kNH + 32 * ch;
        uint8_t* buf = stg + (nst & 1u) * 16384;
        if (q == 0 && lane == 0) ptx::bulk_wait_group_read<1>();
        ptx::named_bar_sync(1 + hh, 128);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<float4*>(buf + r * 128 + ((k ^ sw) << 4)) =
              make_float4(synth(row, c0 + 4 * k), synth(row, c0 + 4 * k + 1), synth(row, c0 + 4 * k + 2), synth(row, c0 + 4 * k + 3));
        ptx::fence_proxy_async();
        ptx::named_bar_sync(1 + hh, 128);
        if (q == 0 && lane == 0) { ptx::tma_store_2d(&tm32x128, buf, c0, tile * 128); ptx::bulk_commit_group(); }
        ++nst;
      }
    } else if constexpr (V == 4) {
      uint8_t* buf = sm + warp * 4096;
      const int sw = lane & 7;
      const int rr = lane >> 3, cc = lane & 7;                          // read-back: 4 rows per instruction, 8 lanes x 16 B per row
      for (int ch = 0; ch < kNH / 32; ++ch) {
        const int c0 = hh * kNH + 32 * ch;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<float4*>(buf + lane * 128 + ((k ^ sw) << 4)) =
              make_float4(synth(row, c0 + 4 * k), synth(row, c0 + 4 * k + 1), synth(row, c0 + 4 * k + 2), synth(row, c0 + 4 * k + 3));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = 4 * i + rr;
          const float4 v = *reinterpret_cast<const float4*>(buf + r * 128 + ((cc ^ (r & 7)) << 4));
          *reinterpret_cast<float4*>(out + (int64_t)(tile * 128 + 32 * q + r) * kC + c0 + 4 * cc) = v;
        }
      }
    } else {
      for (int ch = 0; ch < kNH / 16; ++ch) {
        const int c0 = hh * kNH + 16 * ch;
        float4* dst = reinterpret_cast<float4*>(out + (int64_t)row * kC + c0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          dst[k] = make_float4(synth(row, c0 + 4 * k), synth(row, c0 + 4 * k + 1), synth(row, c0 + 4 * k + 2), synth(row, c0 + 4 * k + 3));
      }
    }
  }
  if (lane == 0) ptx::bulk_wait_group<0>();
}

static bool enc2d(CUtensorMap* m, void* base, uint32_t bi, uint32_t bo, CUtensorMapSwizzle sw) {
  PFN_encodeTiled enc = get_encode_tiled();
  cuuint64_t gdim[2] = {(cuuint64_t)kC, (cuuint64_t)kRows};
  cuuint64_t gstr[1] = {(cuuint64_t)kC * 4};
  cuuint32_t bx[2] = {bi, bo}, es[2] = {1u, 1u};
  return enc && enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int V>
static void run(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, float* out, const std::vector<float>& want_rows) {
  cudaFuncSetAttribute(store_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  cudaMemset(out, 0, (size_t)kRows * kC * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  store_kernel<V><<<kCtas, 256, 80 * 1024>>>(a, b, c, out);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) store_kernel<V><<<kCtas, 256, 80 * 1024>>>(a, b, c, out);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  // spot check: rows 0, 777, last
  std::vector<float> got(kC);
  int bad = 0;
  for (int r : {0, 777, 128 * 147 + 5, kRows - 1}) {
    cudaMemcpy(got.data(), out + (size_t)r * kC, kC * 4, cudaMemcpyDeviceToHost);
    for (int c2 = 0; c2 < kC; ++c2) if (got[c2] != (float)(r & 1023) * 0.5f + (float)c2 * 0.0078125f) ++bad;
  }
  const double us = ms / 5 * 1e3, gb = (double)kRows * kC * 4 / 1e9;
  printf("v%d: %8.1f us per %0.0f MB  = %6.0f GB/s  (%.2f us per 128x320 tile per SM)  %s %s\n", V, us, gb * 1e3, gb / (us * 1e-6), us / kTilesPerCta,
         bad ? "MISMATCH" : "ok", e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  float* out;
  cudaMalloc(&out, (size_t)kRows * kC * 4);
  CUtensorMap t16, t32, t32x128;
  if (!enc2d(&t16, out, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B) || !enc2d(&t32, out, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B) ||
      !enc2d(&t32x128, out, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("encode failed\n"); return 1; }
  std::vector<float> dummy;
  run<0>(t16, t32, t32x128, out, dummy);
  run<1>(t16, t32, t32x128, out, dummy);
  run<2>(t16, t32, t32x128, out, dummy);
  run<3>(t16, t32, t32x128, out, dummy);
  run<4>(t16, t32, t32x128, out, dummy);
  run<5>(t16, t32, t32x128, out, dummy);
  return 0;
}
