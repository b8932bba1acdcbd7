// Persistent, optionally clustered variant of the implicit-GEMM convolution (included by
// conv_gemm.cu).  Same math, operands and epilogue as conv_gemm_kernel, plus
//
//   * persistence: one CTA per SM loops over output tiles, so barrier init, TMEM allocation and
//     descriptor prefetch are paid once per SM, and the accumulator is double-buffered in TMEM
//     (2 x BN columns): the epilogue of tile i overlaps the MMAs of tile i+1 while the TMA ring
//     keeps running across the tile boundary;
//   * operand multicast in a thread-block cluster of CL = 2 or 4 CTAs.  Measured on B200
//     (profiles/): one SM ingests at most ~45 B/clk through TMA, and a 128 x BN TF32 tile needs
//     16 KB (A) + BN*128 B (B) per 32-channel k-block against 4 MMAs of BN/2 cycles — every
//     layer was ingest-bound.  CTAs of a cluster work on tiles that share one operand:
//       SHARE_A: same 128 pixel rows, CL consecutive channel tiles  -> A is loaded once per cluster
//       SHARE_B: CL consecutive row tiles, same channel tile        -> B is loaded once per cluster
//     Each CTA fetches 1/CL of the shared tile with cp.async.bulk.tensor ... .multicast::cluster,
//     which writes the slice to the same shared-memory offset of all CL CTAs and completes the
//     transaction bytes on each CTA's `full` barrier.  A stage may only be refilled when EVERY
//     CTA of the cluster has consumed it, so tcgen05.commit multicasts its arrival to the `empty`
//     barrier of all CL CTAs (count = CL).
// Barrier protocol per accumulator buffer b: tmem_full[b] (MMA commit -> epilogue),
// tmem_empty[b] (kEpiWarps epilogue warps -> MMA), both CTA-local.
#pragma once

enum { SHARE_NONE = 0, SHARE_A = 1, SHARE_B = 2 };

// Epilogue warps: a warp may only read the TMEM lane quarter (warp % 4).  kEpiWarps = 8 puts two
// warps on every quarter (they split the 32-column blocks even / odd); measured on B200
// (profiles/r01_conv_epilogue_ab.md) that is a loss for the TF32 path — the 168-register cap of a
// 320-thread CTA spills in the epilogue and the staging tiles cost a pipeline stage — so 4 it is.
constexpr int kEpiWarps = 4;
constexpr int kPersistentThreads = 64 + 32 * kEpiWarps;

template <int BN, int STAGES, bool F16 = false>
struct SmemLayoutP {
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int STG_OFF = (BAR_OFF + (2 * STAGES + 4) * 8 + 16 + 127) & ~127;  // 16B-aligned staging
  static constexpr int RES_OFF = STG_OFF + kEpiWarps * (F16 ? kStg64BytesPerWarp : kStgBytesPerWarp);
  static constexpr int TOTAL = RES_OFF + (F16 ? kEpiWarps * 2 * kRes64BytesPerBuf : 0);  // residual tiles (cp.async)
};

// Tile owned by this CTA in super-tile `st`.  SHARE_A: super-tile = (row tile, group of CL channel
// tiles); SHARE_B / none: super-tile = (group of CL row tiles, channel tile), row groups fastest so
// neighbouring clusters share the weight tile in L2.
template <int CL>
__device__ __forceinline__ void tile_of(int st, int rank, int share, int m_tiles, int n_tiles, int& mt, int& nt) {
  if (share == SHARE_A) {
    const int ngroups = n_tiles / CL;
    mt = st / ngroups;
    nt = (st - mt * ngroups) * CL + rank;
  } else {
    const int mgroups = (m_tiles + CL - 1) / CL;
    nt = st / mgroups;
    mt = (st - nt * mgroups) * CL + rank;  // may be >= m_tiles in the ragged last group: all-zero A, nothing stored
  }
}

template <int BN, int STAGES, int CL, bool F16>
__global__ void __launch_bounds__(kPersistentThreads, 1)
conv_gemm_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                            const ConvParams p, const int m_tiles, const int n_tiles, const int num_super,
                            const int share) {
  using L = SmemLayoutP<BN, STAGES, F16>;
  constexpr uint32_t TMEM_COLS = (2 * BN) < 32 ? 32 : (2 * BN);
  static_assert(2 * BN <= 512, "double-buffered accumulator must fit the 512 TMEM columns");
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CL) - 1);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int S = CL > 1 ? 1 : p.splits;                   // K ranges per tile (host: 1 for clustered launches)
  const int num_items = num_super * S;

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int iters = p.taps * p.kblocks;
  const int rank = CL > 1 ? static_cast<int>(tc05::cluster_ctarank()) : 0;
  const int cluster_id = blockIdx.x / CL;
  const int num_clusters = gridDim.x / CL;

  if (warp == 0 && lane == 0) {
    tc05::prefetch_tmap(&tmA);
    tc05::prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      tc05::mbar_init(&full_bar[s], 1);
      tc05::mbar_init(&empty_bar[s], CL);  // one tcgen05.commit arrival from every CTA of the cluster
    }
    for (int b = 0; b < 2; ++b) {
      tc05::mbar_init(&tmem_full[b], 1);
      tc05::mbar_init(&tmem_empty[b], kEpiWarps);
    }
    tc05::fence_barrier_init();
  }
  if (warp == 1) tc05::tmem_alloc<TMEM_COLS>(tmem_slot);
  tc05::fence_before_sync();
  __syncthreads();
  if (CL > 1) tc05::cluster_sync();  // peers' barriers exist before anyone multicasts into them
  tc05::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // everything above touched only shared memory, TMEM and the kernel parameters: it ran under the
  // previous kernel's tail (pdl.cuh); from here on global memory is read and written
  pdl_wait();

  if (warp == 0) {
    if (tc05::elect_one()) {
      const int wp = p.w + 2;
      int it = 0;  // global k-iteration counter: the smem ring is continuous across tiles
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        const int st = item / S, split = item - st * S;
        int mt, nt;
        tile_of<CL>(st, rank, share, m_tiles, n_tiles, mt, nt);
        const int64_t m0 = static_cast<int64_t>(mt) * BM;
        const int n0 = nt * BN;
        const int j1 = static_cast<int>(static_cast<int64_t>(split + 1) * iters / S);
        for (int j = static_cast<int>(static_cast<int64_t>(split) * iters / S); j < j1; ++j, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          const int t = j / p.kblocks;
          const int kb = j - t * p.kblocks;
          int64_t row = m0;
          if (p.taps == 9) row += static_cast<int64_t>(t / 3 - 1) * wp + (t % 3 - 1);
          tc05::mbar_wait(&empty_bar[s], ph ^ 1, p.err, 111);
          tc05::mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          const int32_t ca = p.in_coff + kb * p.bk, cb = kb * p.bk;  // element coordinates
          const int32_t rb = t * p.cout_pad + n0;
          if (CL > 1 && share == SHARE_A) {
            // my 1/CL slice of the shared A tile goes to every CTA of the cluster
            tc05::tma_load_2d_multicast(sa + rank * (A_BYTES / CL), &tmA, &full_bar[s], ca,
                                        static_cast<int32_t>(row) + rank * (BM / CL), kMask);
            tc05::tma_load_2d(sa + A_BYTES, &tmB, &full_bar[s], cb, rb);
          } else if (CL > 1 && share == SHARE_B) {
            tc05::tma_load_2d(sa, &tmA, &full_bar[s], ca, static_cast<int32_t>(row));
            tc05::tma_load_2d_multicast(sa + A_BYTES + rank * (L::B_BYTES / CL), &tmB, &full_bar[s], cb,
                                        rb + rank * (BN / CL), kMask);
          } else {
            tc05::tma_load_2d(sa, &tmA, &full_bar[s], ca, static_cast<int32_t>(row));
            tc05::tma_load_2d(sa + A_BYTES, &tmB, &full_bar[s], cb, rb);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (tc05::elect_one()) {
      const uint32_t idesc = F16 ? tc05::make_idesc_f16(BM, BN) : tc05::make_idesc_tf32(BM, BN);
      int it = 0;
      int local = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters, ++local) {
        const int split = item % S;
        const int j0 = static_cast<int>(static_cast<int64_t>(split) * iters / S);
        const int j1 = static_cast<int>(static_cast<int64_t>(split + 1) * iters / S);
        const int buf = local & 1;
        const uint32_t use = static_cast<uint32_t>(local >> 1);
        tc05::mbar_wait(&tmem_empty[buf], (use & 1) ^ 1, p.err, 112);
        tc05::fence_after_sync();
        const uint32_t d = tmem_base + buf * BN;
        for (int j = j0; j < j1; ++j, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          tc05::mbar_wait(&full_bar[s], ph, p.err, 113);
          tc05::fence_after_sync();
          const uint32_t sa = tc05::smem_u32(smem + s * L::STAGE_BYTES);
          const uint64_t da = tc05::make_desc_sw128(sa);
          const uint64_t db = tc05::make_desc_sw128(sa + A_BYTES);
          // a 128-byte k-block row is 4 MMA K-steps of 32 bytes in either type (8 x fp32 / 16 x fp16)
          if (F16) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc05::umma_f16_ss(d, da + 2 * k, db + 2 * k, idesc, (j != j0 || k != 0) ? 1u : 0u);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc05::umma_tf32_ss(d, da + 2 * k, db + 2 * k, idesc, (j != j0 || k != 0) ? 1u : 0u);
          }
          if (CL > 1) tc05::umma_commit_multicast(&empty_bar[s], kMask);
          else tc05::umma_commit(&empty_bar[s]);
        }
        tc05::umma_commit(&tmem_full[buf]);
      }
    }
  } else {
    // ---------------- epilogue: warps 2..9; warp w owns TMEM lane quarter (w % 4) and the 32-column
    // blocks of parity (w - 2) / 4
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    float* stg = reinterpret_cast<float*>(smem + L::STG_OFF) +
                 (warp - 2) * ((F16 ? kStg64BytesPerWarp : kStgBytesPerWarp) / 4);
    const int wp = p.w + 2;
    const int64_t per_img = static_cast<int64_t>(p.h + 2) * wp;
    int local = 0;
    for (int item = cluster_id; item < num_items; item += num_clusters, ++local) {
      const int st = item / S, split = item - st * S;
      int mt, nt;
      tile_of<CL>(st, rank, share, m_tiles, n_tiles, mt, nt);
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
      const uint32_t interior_mask = __ballot_sync(0xffffffffu, interior);
      const int64_t row0 = static_cast<int64_t>(mt) * BM + q * 32;
      // 64-column fp16 steps with the residual through shared memory: whole tiles of real channels
      const bool res_pipe = F16 && kEpiWarps == 4 && BN >= 64 && S == 1 && p.f16_out && (p.cout - n0) >= BN;
      uint8_t* const resb = smem + L::RES_OFF + (warp - 2) * 2 * kRes64BytesPerBuf;
      if (res_pipe) {  // the first two steps' residual rows are requested before the MMAs are waited for
        conv_epilogue_prefetch64_smem(resb, lane, row0, interior_mask, n0, p);
        if (BN > 64) conv_epilogue_prefetch64_smem(resb + kRes64BytesPerBuf, lane, row0, interior_mask, n0 + 64, p);
      }
      tc05::mbar_wait(&tmem_full[buf], use & 1, p.err, 114);
      tc05::fence_after_sync();
      if (BN == 32 && half == 1) {  // a single column block: the second warp of the quarter has nothing to read
        if (lane == 0) tc05::mbar_arrive(&tmem_empty[buf]);
        continue;
      }
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;

      // ---- split-K: this CTA accumulated K range `split` of tile `st`: park the fp32 partial tile
      // in the workspace [tile][split][128 rows][BN]; splitk_epilogue_kernel (next launch) sums the
      // S partials in split order and applies bias / residual / ReLU / conversion.  (A first version
      // let the CTA that finished last reduce in place: one SM re-reading S x 128 KB with 128
      // threads is latency-bound at ~16 GB/s — 2x slower than not splitting at all.)
      if (S > 1) {
        float* mine = p.sk_ws + ((static_cast<int64_t>(st) * S + split) * BM + q * 32 + lane) * BN;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tc05::tmem_ld32(tacc + c0, v);
          tc05::tmem_ld_wait();
          if (c0 + 32 >= BN) {
            tc05::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(&tmem_empty[buf]);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(mine + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        continue;
      }
      auto load_acc = [&](const int c0, uint32_t(&v)[32]) {
        tc05::tmem_ld32(tacc + c0, v);
        tc05::tmem_ld_wait();
      };
      auto release_acc = [&]() {
        // all TMEM reads of this tile are done: hand the buffer back before the global stores
        tc05::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc05::mbar_arrive(&tmem_empty[buf]);
      };
      if (res_pipe) {
        constexpr int NS = BN / 64;
#pragma unroll 1
        for (int sidx = 0; sidx < NS; ++sidx) {
          const int c0 = sidx * 64;
          uint32_t v[32];
          load_acc(c0, v);
          conv_epilogue_stage64(v, stg, lane, 0);
          load_acc(c0 + 32, v);
          if (sidx + 1 == NS) release_acc();
          conv_epilogue_stage64(v, stg, lane, 32);
          if (sidx + 1 < NS) cp_async_wait<1>();  // this step's rows have landed; the next step's may be in flight
          else cp_async_wait<0>();
          uint8_t* rb = resb + (sidx & 1) * kRes64BytesPerBuf;
          conv_epilogue_store64_smem(stg, rb, lane, row0, interior_mask, n0 + c0, p);
          if (sidx + 2 < NS) conv_epilogue_prefetch64_smem(rb, lane, row0, interior_mask, n0 + c0 + 128, p);
        }
        continue;
      }
      if (F16 && kEpiWarps == 4 && BN >= 64 && p.f16_out) {
        // fp16 maps: 64 output channels per step, so that every row segment a warp reads (residual)
        // or writes is a full 128-byte line — with 32-column steps the 64-byte segments need the
        // same number of LSU wavefronts as the fp32 path for half the bytes
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 64) {
          const bool full64 = p.cout - (n0 + c0) >= 64;
          uint4 res[8];
          if (full64) conv_epilogue_prefetch64(res, lane, row0, interior_mask, n0 + c0, p);
          uint32_t v[32];
          load_acc(c0, v);
          if (full64) conv_epilogue_stage64(v, stg, lane, 0);
          else conv_epilogue_block(v, stg, lane, row0, interior_mask, n0 + c0, p);  // ragged channel tail: generic path
          load_acc(c0 + 32, v);
          if (c0 + 64 >= BN) release_acc();
          if (full64) {
            conv_epilogue_stage64(v, stg, lane, 32);
            conv_epilogue_store64(stg, lane, row0, interior_mask, n0 + c0, p, res);
          } else {
            conv_epilogue_block(v, stg, lane, row0, interior_mask, n0 + c0 + 32, p);
          }
        }
        continue;
      }
#pragma unroll 1
      for (int c0 = half * 32; c0 < BN; c0 += (kEpiWarps / 4) * 32) {
        uint32_t v[32];
        load_acc(c0, v);
        if (c0 + (kEpiWarps / 4) * 32 >= BN) release_acc();
        conv_epilogue_block(v, stg, lane, row0, interior_mask, n0 + c0, p);
      }
    }
  }

  tc05::fence_before_sync();
  __syncthreads();
  if (CL > 1) tc05::cluster_sync();  // nobody leaves while a peer may still multicast / arrive here
  if (warp == 1) {
    __syncwarp();
    tc05::fence_after_sync();
    tc05::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// Launch geometry chosen on the host: which operand a cluster shares and how many CTAs it has.
struct ClusterChoice {
  int share;
  int cl;
};

inline ClusterChoice choose_cluster(int m_tiles, int n_tiles) {
  static const int max_cl = [] {
    // MIVOS_CONV_CLUSTER = 2 | 4 enables operand multicast.  Default 1 (off): measured on B200
    // (profiles/r01_conv_cluster_ab.md) multicast does not help — the limit is the bytes an SM can
    // take INTO its shared memory per clock, and a multicast slice still lands in every CTA's smem.
    const char* e = getenv("MIVOS_CONV_CLUSTER");
    const int v = e ? atoi(e) : 1;
    return v < 1 ? 1 : (v > 4 ? 4 : v);
  }();
  if (max_cl == 1) return {SHARE_NONE, 1};
  // several channel tiles read the same pixels -> share A (exact divisor only: a phantom channel
  // tile would read the next tap's weights); otherwise share the weight tile across row tiles
  if (n_tiles >= 2) {
    for (int c = max_cl; c >= 2; c >>= 1)
      if (n_tiles % c == 0) return {SHARE_A, c};
  }
  if (m_tiles >= 2) {
    const int c = (max_cl >= 4 && m_tiles >= 4) ? 4 : 2;
    return {SHARE_B, c};
  }
  return {SHARE_NONE, 1};
}

template <int BN, int STAGES, int CL, bool F16>
int launch_persistent_cl(const mivos_conv_args* a, const ConvParams& p, int m_tiles, int n_tiles, int share,
                         cudaStream_t stream) {
  using L = SmemLayoutP<BN, STAGES, F16>;
  constexpr int smem_bytes = L::TOTAL + 1024;
  static_assert(smem_bytes <= 232448, "stage ring + epilogue staging exceed the 227 KB a CTA may use");
  auto kernel = conv_gemm_persistent_kernel<BN, STAGES, CL, F16>;
  static int max_clusters = 0;  // resident clusters of this configuration (queried once)
  if (max_clusters == 0) {
    MIVOS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    if (CL > 1) {
      cudaLaunchConfig_t q{};
      q.gridDim = dim3(num_sms() / CL * CL);
      q.blockDim = dim3(kPersistentThreads);
      q.dynamicSmemBytes = smem_bytes;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      q.attrs = at; q.numAttrs = 1;
      int n = 0;
      MIVOS_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, kernel, &q));
      MIVOS_REQUIRE(n > 0, "conv_gemm: no resident cluster of %d CTAs with %d B shared memory", CL, smem_bytes);
      max_clusters = n;
    } else {
      max_clusters = num_sms();
    }
  }
  // operand tensor maps: the shared operand is fetched in 1/CL slices
  CUtensorMap tmA, tmB;
  const int eb = a->in_f16 ? 2 : 4;
  int rc = encode_tmap_2d(&tmA, a->in, static_cast<uint64_t>(a->in_rows), static_cast<uint64_t>(a->in_cstride),
                          static_cast<uint64_t>(a->in_cstride), p.bk, share == SHARE_A ? BM / CL : BM, eb);
  if (rc != MIVOS_OK) return rc;
  rc = encode_tmap_2d(&tmB, a->weight, static_cast<uint64_t>(a->taps) * a->cout_pad, static_cast<uint64_t>(a->cin_pad),
                      static_cast<uint64_t>(a->cin_pad), p.bk, share == SHARE_B ? BN / CL : BN, eb);
  if (rc != MIVOS_OK) return rc;
  const int num_super = share == SHARE_A ? m_tiles * (n_tiles / CL) : ((m_tiles + CL - 1) / CL) * n_tiles;
  const int num_items = num_super * (CL > 1 ? 1 : p.splits);
  const int clusters = num_items < max_clusters ? num_items : max_clusters;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(kPersistentThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (CL > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = CL; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  MIVOS_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, p, m_tiles, n_tiles, num_super, share));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return MIVOS_OK;
}

template <int BN, int STAGES, bool F16>
int launch_persistent(const mivos_conv_args* a, const ConvParams& p, cudaStream_t stream) {
  const int m_tiles = static_cast<int>(ceil_div64(p.rows, BM));
  const int n_tiles = a->cout_pad / BN;
  if constexpr (F16) {  // operand multicast (measured: no gain) is only instantiated for the TF32 kernels
    return launch_persistent_cl<BN, STAGES, 1, F16>(a, p, m_tiles, n_tiles, SHARE_NONE, stream);
  } else {
    const ClusterChoice c = choose_cluster(m_tiles, n_tiles);
    switch (c.cl) {
      case 4: return launch_persistent_cl<BN, STAGES, 4, false>(a, p, m_tiles, n_tiles, c.share, stream);
      case 2: return launch_persistent_cl<BN, STAGES, 2, false>(a, p, m_tiles, n_tiles, c.share, stream);
      default: return launch_persistent_cl<BN, STAGES, 1, false>(a, p, m_tiles, n_tiles, SHARE_NONE, stream);
    }
  }
}
