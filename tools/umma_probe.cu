// Hardware probe (B200): which rows does tcgen05.mma read when the shared-memory matrix descriptor
// of a K-major SWIZZLE_128B operand starts at a row that is NOT a multiple of 8 (start address not
// 1024-byte aligned)?  The 3x3 convolutions load the SAME activation rows three times (taps dx = -1,
// 0, +1 are the HALO matrix shifted by one row): if a descriptor may start at any row, one TMA box
// of 128 + 2 rows serves three taps.
//
// Method: A = [ROWS, 64] fp16 with A[r][c] = r + c/64 (exact in fp16 for r < 512); B = 64x64
// identity, so D[m][n] = A[row(m)][n] shows which row the tensor core fetched for accumulator lane
// m.  For shift s in 0..17 and the variants of the descriptor's "matrix base offset" field
// (bits 49-51): 0, (addr >> 7) & 7, s & 7 — report the number of mismatching elements against
// A[s + m].
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../mivos_b200/csrc \
//        -o build/umma_probe umma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "tc05.cuh"

constexpr int ROWS = 160;     // rows of A staged in shared memory (20 groups of 8)
constexpr int KC = 64;        // fp16 elements per 128-byte row
constexpr int NSHIFT = 18;
constexpr int NVAR = 3;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int* err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;                     // ROWS * 128 B = 20 KB
  uint8_t* sB = smem + ROWS * 128;        // 64 * 128 B = 8 KB (1024-aligned: ROWS*128 = 20480)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 64 * 128);
  uint64_t* mma_bar = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tc05::mbar_init(bar, 1);
    tc05::mbar_init(mma_bar, 1);
    tc05::fence_barrier_init();
  }
  if (warp == 0) tc05::tmem_alloc<64>(tmem_slot);
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    tc05::mbar_arrive_expect_tx(bar, ROWS * 128 + 64 * 128);
    tc05::tma_load_2d(sA, &tmA, bar, 0, 0);        // box {64, 80} twice
    tc05::tma_load_2d(sA + 80 * 128, &tmA, bar, 0, 80);
    tc05::tma_load_2d(sB, &tmB, bar, 0, 0);
  }
  tc05::mbar_wait(bar, 0, err, 901);
  tc05::fence_after_sync();
  uint32_t parity = 0;
  for (int var = 0; var < NVAR; ++var) {
    for (int s = 0; s < NSHIFT; ++s) {
      if (threadIdx.x == 0) {
        const uint32_t a_addr = tc05::smem_u32(sA) + s * 128;
        uint64_t da = tc05::make_desc_sw128(a_addr);
        uint32_t bo = 0;
        if (var == 1) bo = (a_addr >> 7) & 7;
        if (var == 2) bo = (8 - (s & 7)) & 7;
        da |= static_cast<uint64_t>(bo) << 49;
        const uint64_t db = tc05::make_desc_sw128(tc05::smem_u32(sB));
        const uint32_t idesc = tc05::make_idesc_f16(128, 64);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc05::umma_f16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, k != 0 ? 1u : 0u);
        tc05::umma_commit(mma_bar);
      }
      tc05::mbar_wait(mma_bar, parity, err, 902);
      parity ^= 1;
      tc05::fence_after_sync();
      // thread t reads accumulator lane t (row m = t), 64 columns
      float* dst = out + ((static_cast<size_t>(var) * NSHIFT + s) * 128 + threadIdx.x) * 64;
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tc05::tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
        tc05::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) dst[c0 + j] = __uint_as_float(v[j]);
      }
      tc05::fence_before_sync();
      __syncthreads();
      tc05::fence_after_sync();
    }
  }
  __syncthreads();
  if (warp == 0) tc05::tmem_dealloc<64>(tmem_base);
}

static int encode(CUtensorMap* m, void* base, uint64_t rows, uint32_t box_rows) {
  cuuint64_t dims[2] = {KC, rows};
  cuuint64_t strides[1] = {KC * 2};
  cuuint32_t box[2] = {KC, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

int main() {
  cudaFree(0);
  std::vector<__half> hA(ROWS * KC), hB(64 * KC);
  for (int r = 0; r < ROWS; ++r)
    for (int c = 0; c < KC; ++c) hA[r * KC + c] = __float2half(static_cast<float>(r) + c / 64.0f);
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < KC; ++k) hB[n * KC + k] = __float2half(n == k ? 1.0f : 0.0f);
  __half *dA, *dB;
  float* dOut;
  int* dErr;
  const size_t out_n = static_cast<size_t>(NVAR) * NSHIFT * 128 * 64;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dOut, out_n * 4);
  cudaMalloc(&dErr, 4);
  cudaMemset(dErr, 0, 4);
  cudaMemset(dOut, 0, out_n * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  if (encode(&tmA, dA, ROWS, 80) || encode(&tmB, dB, 64, 64)) {
    printf("tensor map encode failed\n");
    return 2;
  }
  const int smem = ROWS * 128 + 64 * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(tmA, tmB, dOut, dErr);
  cudaError_t e = cudaDeviceSynchronize();
  int herr = 0;
  cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost);
  printf("probe: cuda=%s err_flag=%d\n", cudaGetErrorString(e), herr);
  if (e != cudaSuccess) return 1;
  std::vector<float> o(out_n);
  cudaMemcpy(o.data(), dOut, out_n * 4, cudaMemcpyDeviceToHost);
  const char* names[NVAR] = {"base_offset=0", "base_offset=(addr>>7)&7", "base_offset=(8-s)&7"};
  for (int var = 0; var < NVAR; ++var) {
    printf("variant %d (%s): mismatches per shift:", var, names[var]);
    for (int s = 0; s < NSHIFT; ++s) {
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int c = 0; c < 64; ++c) {
          const float want = __half2float(hA[(s + m) * KC + c]);
          if (o[((static_cast<size_t>(var) * NSHIFT + s) * 128 + m) * 64 + c] != want) ++bad;
        }
      printf(" %d", bad);
    }
    printf("\n");
    // show what rows were fetched for shift 1 and 3 (first 12 lanes, column 0 = row index)
    for (int s : {1, 3, 9}) {
      printf("   shift %d rows seen by lanes 0..15:", s);
      for (int m = 0; m < 16; ++m) printf(" %.3f", o[((static_cast<size_t>(var) * NSHIFT + s) * 128 + m) * 64 + 0]);
      printf("  | lane0 cols 0,8,16,63: %.4f %.4f %.4f %.4f\n", o[((static_cast<size_t>(var) * NSHIFT + s) * 128) * 64 + 0],
             o[((static_cast<size_t>(var) * NSHIFT + s) * 128) * 64 + 8], o[((static_cast<size_t>(var) * NSHIFT + s) * 128) * 64 + 16],
             o[((static_cast<size_t>(var) * NSHIFT + s) * 128) * 64 + 63]);
    }
  }
  return 0;
}
