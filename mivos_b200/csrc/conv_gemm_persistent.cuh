// Persistent variant of the implicit-GEMM convolution (included by conv_gemm.cu).
//
// Same math, operands, pipeline stages and epilogue as conv_gemm_kernel, but
//   * the grid is one CTA per SM and every CTA loops over output tiles
//     (tile = blockIdx.x + i * gridDim.x; M-tile fastest so that neighbouring CTAs share the
//     weight tile in L2), so barrier init, TMEM allocation and descriptor prefetch are paid once
//     per SM instead of once per tile;
//   * the accumulator is double-buffered in TMEM (2 x BN columns): while the epilogue warps drain
//     tile i, the MMA warp already accumulates tile i+1 and the TMA producer runs further ahead —
//     the smem ring never drains at a tile boundary.
// Barrier protocol per accumulator buffer b: tmem_full[b] (MMA commit -> epilogue),
// tmem_empty[b] (4 epilogue warps -> MMA).
#pragma once

template <int BN, int STAGES>
struct SmemLayoutP {
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int STG_OFF = (BAR_OFF + (2 * STAGES + 4) * 8 + 16 + 127) & ~127;  // 16B-aligned staging
  static constexpr int TOTAL = STG_OFF + 4 * kStgBytesPerWarp;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
conv_gemm_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                            const ConvParams p, const int m_tiles, const int num_tiles) {
  using L = SmemLayoutP<BN, STAGES>;
  constexpr uint32_t TMEM_COLS = (2 * BN) < 32 ? 32 : (2 * BN);
  static_assert(2 * BN <= 512, "double-buffered accumulator must fit the 512 TMEM columns");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int iters = p.taps * p.kblocks;

  if (warp == 0 && lane == 0) {
    tc05::prefetch_tmap(&tmA);
    tc05::prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      tc05::mbar_init(&full_bar[s], 1);
      tc05::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc05::mbar_init(&tmem_full[b], 1);
      tc05::mbar_init(&tmem_empty[b], 4);
    }
    tc05::fence_barrier_init();
  }
  if (warp == 1) tc05::tmem_alloc<TMEM_COLS>(tmem_slot);
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (tc05::elect_one()) {
      const int wp = p.w + 2;
      int it = 0;  // global k-iteration counter: the smem ring is continuous across tiles
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile % m_tiles, nt = tile / m_tiles;
        const int64_t m0 = static_cast<int64_t>(mt) * BM;
        const int n0 = nt * BN;
        for (int j = 0; j < iters; ++j, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          const int t = j / p.kblocks;
          const int kb = j - t * p.kblocks;
          int64_t row = m0;
          if (p.taps == 9) row += static_cast<int64_t>(t / 3 - 1) * wp + (t % 3 - 1);
          tc05::mbar_wait(&empty_bar[s], ph ^ 1, p.err, 111);
          tc05::mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          tc05::tma_load_2d(sa, &tmA, &full_bar[s], p.in_coff + kb * BK, static_cast<int32_t>(row));
          tc05::tma_load_2d(sa + A_BYTES, &tmB, &full_bar[s], kb * BK, t * p.cout_pad + n0);
        }
      }
    }
  } else if (warp == 1) {
    if (tc05::elect_one()) {
      constexpr uint32_t idesc = tc05::make_idesc_tf32(BM, BN);
      int it = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const int buf = local & 1;
        const uint32_t use = static_cast<uint32_t>(local >> 1);
        tc05::mbar_wait(&tmem_empty[buf], (use & 1) ^ 1, p.err, 112);
        tc05::fence_after_sync();
        const uint32_t d = tmem_base + buf * BN;
        for (int j = 0; j < iters; ++j, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          tc05::mbar_wait(&full_bar[s], ph, p.err, 113);
          tc05::fence_after_sync();
          const uint32_t sa = tc05::smem_u32(smem + s * L::STAGE_BYTES);
          const uint64_t da = tc05::make_desc_sw128(sa);
          const uint64_t db = tc05::make_desc_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k)
            tc05::umma_tf32_ss(d, da + 2 * k, db + 2 * k, idesc, (j | k) != 0 ? 1u : 0u);
          tc05::umma_commit(&empty_bar[s]);
        }
        tc05::umma_commit(&tmem_full[buf]);
      }
    }
  } else {
    // ---------------- epilogue: warps 2..5 own TMEM lane quarters (warp % 4)
    const int q = warp & 3;
    float* stg = reinterpret_cast<float*>(smem + L::STG_OFF) + q * (kStgBytesPerWarp / 4);
    const int wp = p.w + 2;
    const int64_t per_img = static_cast<int64_t>(p.h + 2) * wp;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int mt = tile % m_tiles, nt = tile / m_tiles;
      const int64_t r = static_cast<int64_t>(mt) * BM + q * 32 + lane;
      const int n0 = nt * BN;
      bool interior = false;
      if (r < p.rows) {
        const int64_t rem = r % per_img;
        const int y = static_cast<int>(rem / wp);
        const int x = static_cast<int>(rem - static_cast<int64_t>(y) * wp);
        interior = (y >= 1) && (y <= p.h) && (x >= 1) && (x <= p.w);
      }
      const int buf = local & 1;
      const uint32_t use = static_cast<uint32_t>(local >> 1);
      tc05::mbar_wait(&tmem_full[buf], use & 1, p.err, 114);
      tc05::fence_after_sync();
      const uint32_t interior_mask = __ballot_sync(0xffffffffu, interior);
      const int64_t row0 = static_cast<int64_t>(mt) * BM + q * 32;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tc05::tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + c0, v);
        tc05::tmem_ld_wait();
        if (c0 + 32 >= BN) {
          // all TMEM reads of this tile are done: hand the buffer back before the global stores
          tc05::fence_before_sync();
          __syncwarp();
          if (lane == 0) tc05::mbar_arrive(&tmem_empty[buf]);
        }
        conv_epilogue_block(v, stg, lane, row0, interior_mask, n0 + c0, p);
      }
    }
  }

  tc05::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc05::fence_after_sync();
    tc05::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int BN, int STAGES>
int launch_persistent(const mivos_conv_args* a, const CUtensorMap& tmA, const CUtensorMap& tmB,
                      const ConvParams& p, cudaStream_t stream) {
  using L = SmemLayoutP<BN, STAGES>;
  constexpr int smem_bytes = L::TOTAL + 1024;
  static bool configured = false;
  if (!configured) {
    MIVOS_CUDA_OK(cudaFuncSetAttribute(conv_gemm_persistent_kernel<BN, STAGES>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  const int m_tiles = static_cast<int>(ceil_div64(p.rows, BM));
  const int num_tiles = m_tiles * (a->cout_pad / BN);
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  conv_gemm_persistent_kernel<BN, STAGES><<<grid, 192, smem_bytes, stream>>>(tmA, tmB, p, m_tiles, num_tiles);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MIVOS_CUDA_OK(cudaGetLastError());
  return MIVOS_OK;
}
