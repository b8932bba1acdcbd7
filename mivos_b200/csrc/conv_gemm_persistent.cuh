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

// Epilogue warps (template parameter EW): a warp may only read the TMEM lane quarter (warp % 4).  EW = 8 puts
// two warps on every quarter; they split a tile's column steps even / odd.  Measured on B200: with the
// register epilogue of round 1 that was a loss for the TF32 path (profiles/r01_conv_epilogue_ab.md: the
// 168-register cap of a 320-thread CTA spilled, the staging tiles cost a pipeline stage), so fp32 maps keep 4.
// The TMA epilogue of the fp16 maps needs ~60 registers and 12 KB per warp, and the output-bound layers (1x1
// expansions with residual: 1 k-block of MMAs per 128 x 128 outputs) are bound by how many epilogue
// instructions an SM issues with ONE warp per scheduler (profiles/r02c2: 0.6 IPC, stalls barrier / wait /
// long_scoreboard): BN = 128 fp16 tiles run 8 epilogue warps.
constexpr int persistent_threads(int ew) { return 64 + 32 * ew; }

// fp16 OUTPUT maps leave the SM through TMA (conv_epilogue_tma_step below): per epilogue warp two
// 32-row x 64-channel output tiles (4 KB each, 128B-swizzled: exactly the layout tcgen05.ld's
// row-per-thread ownership produces, so nothing is transposed), two residual tiles filled by TMA loads
// two steps ahead, and one tile for the optional ReLU copy.  The generic register path (ragged channel
// tiles, fp32 outputs, split-K) stages through the same bytes.
constexpr int kEpiTile = 32 * 128;                 // 32 rows x 64 fp16

// 3x3 convolutions, tap-reuse mode (T9): the taps (dy, dx = -1, 0, +1) of one dy read the SAME activation rows
// shifted by one HALO row, and a K-major SWIZZLE_128B UMMA descriptor may start at any 128-byte row of a
// TMA-written tile (measured on B200: profiles/r02c1_umma_probe.log).  So ONE box of 128 + 2 rows (136 for the
// 8-row swizzle groups) per (k-block, dy) serves three taps: A' tiles and weight tiles flow through two rings,
// and the operand bytes an SM ingests per tap drop from 16 KB + B to 6 KB + B (the small-N 3x3 layers and the
// big layers at batch 1 are bound by exactly that ingest).
constexpr int A3_ROWS = 136;
constexpr int A3_BYTES = 18 * 1024;  // 136 rows x 128 B = 17408, padded to the 1024-byte swizzle alignment

template <int BN, int STAGES, bool F16 = false, int EW = 4, bool T9 = false>
struct SmemLayoutP {
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NA = T9 ? (BN == 256 ? 2 : 3) : 0;                                  // A' ring
  static constexpr int NB = T9 ? (BN == 256 ? 3 : BN == 128 ? 4 : BN == 64 ? 6 : 7) : 0;  // weight ring
  static constexpr int NRING = T9 ? NA + NB : STAGES;
  static constexpr int BAR_OFF = T9 ? NA * A3_BYTES + NB * B_BYTES : STAGES * STAGE_BYTES;
  static constexpr int NBAR = 2 * NRING + 4 + (F16 ? 3 * EW : 0);  // + 3 residual barriers per epilogue warp
  static constexpr int STG_OFF = F16 ? ((BAR_OFF + NBAR * 8 + 16 + 1023) & ~1023)   // swizzled TMA tiles: 1024-byte aligned
                                     : ((BAR_OFF + NBAR * 8 + 16 + 127) & ~127);    // 16B-aligned staging
  // TMA epilogue bytes of a warp: io[NBUF] | out_relu.  An io tile receives the residual rows of a 64-channel
  // step by TMA, is transformed IN PLACE into the output rows and leaves by TMA; the residual of the step
  // NBUF - 1 steps ahead (across tile boundaries: one whole tile ahead for the 8-warp kernels) is in flight
  // meanwhile, so its L2 / HBM latency is never waited for.
  static constexpr int NBUF = EW == 8 ? 2 : 3;
  static constexpr int PER_WARP = F16 ? (NBUF + 1) * kEpiTile : kStgBytesPerWarp;
  static_assert(!F16 || (EW == 4 ? kStg64BytesPerWarp : kStgBytesPerWarp) <= PER_WARP, "generic staging tile must fit the warp's epilogue bytes");
  static constexpr int TOTAL = STG_OFF + EW * PER_WARP;
};

// One 64-channel step of the TMA epilogue for the warp's 32 rows: thread = accumulator row.
//   acc (TMEM) + bias (+ residual tile in shared memory) (ReLU) -> fp16 -> swizzled output tile.
// Rows of the HALO border are written as zeros (the border stays zero), so the whole box can be stored.
// The 8 bias vectors (32 channels) of one half-step: requested BEFORE the accumulator / residual waits, so their
// L1 / L2 latency is off the dependent chain of the single warp a scheduler has for this work.
struct BiasHalf {
  float4 b[8];
};
__device__ __forceinline__ void conv_epilogue_load_bias(BiasHalf& bh, const int half, const int ncol0, const ConvParams& p) {
#pragma unroll
  for (int j = 0; j < 8; ++j) bh.b[j] = __ldg(reinterpret_cast<const float4*>(p.bias + ncol0 + half * 32 + j * 4));
}

template <bool kRelu2>
__device__ __forceinline__ void conv_epilogue_tma_half(const uint32_t (&v)[32], const BiasHalf& bh, const int half, uint8_t* io,
                                                       uint8_t* orl, const bool has_res, const bool interior,
                                                       const int lane, const int ncol0, const ConvParams& p) {
  const int sw = lane & 7;
  uint8_t* row = io + lane * 128;  // residual in, output out: the same 16-byte pieces, read then overwritten by this lane
  // all four residual pieces first (the compiler cannot prove that the in-place stores below do not alias them,
  // and would otherwise serialise load -> math -> store per piece)
  uint4 r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int pos = ((half * 4 + j) ^ sw) * 16;  // 128B swizzle: piece index XOR (row mod 8)
    r[j] = has_res ? *reinterpret_cast<const uint4*>(row + pos) : make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int pos = ((half * 4 + j) ^ sw) * 16;
    const float4 b0 = bh.b[2 * j], b1 = bh.b[2 * j + 1];
    float o[8] = {__uint_as_float(v[j * 8 + 0]) + b0.x, __uint_as_float(v[j * 8 + 1]) + b0.y,
                  __uint_as_float(v[j * 8 + 2]) + b0.z, __uint_as_float(v[j * 8 + 3]) + b0.w,
                  __uint_as_float(v[j * 8 + 4]) + b1.x, __uint_as_float(v[j * 8 + 5]) + b1.y,
                  __uint_as_float(v[j * 8 + 6]) + b1.z, __uint_as_float(v[j * 8 + 7]) + b1.w};
    if (has_res) {
      const uint32_t w[4] = {r[j].x, r[j].y, r[j].z, r[j].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
        o[2 * e] += t.x;
        o[2 * e + 1] += t.y;
      }
    }
    if (p.relu) {
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaxf(o[e], 0.f);
    }
    if (!interior) {
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = 0.f;
    }
    *reinterpret_cast<uint4*>(row + pos) = pack8_half(o);
    if (kRelu2) {
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaxf(o[e], 0.f);
      *reinterpret_cast<uint4*>(orl + lane * 128 + pos) = pack8_half(o);
    }
  }
}

// Tile owned by this CTA in super-tile `st`.  SHARE_A: super-tile = (row tile, group of CL channel tiles).
// SHARE_B / none: super-tile = (group of CL row tiles, channel tile).  Without clusters the CHANNEL tiles of one
// row tile are consecutive (share = SHARE_NONE): CTAs that run at the same time then read the same activation
// rows (L2 hits) and write / read-as-residual the adjacent column blocks of the same output rows, so that L2
// merges them into whole rows before they reach DRAM — with row tiles fastest, an output-bound layer wrote the
// left half of EVERY row long before the right half (256-byte bursts 512 bytes apart: ~2.5 TB/s).  SHARE_ROWS
// (MIVOS_CONV_TILE_ORDER=m) keeps the round-1 order for A/B measurements.
enum { SHARE_ROWS = 3 };
template <int CL>
__device__ __forceinline__ void tile_of(int st, int rank, int share, int m_tiles, int n_tiles, int& mt, int& nt) {
  if (share == SHARE_A) {
    const int ngroups = n_tiles / CL;
    mt = st / ngroups;
    nt = (st - mt * ngroups) * CL + rank;
  } else if (CL == 1 && share == SHARE_NONE) {
    mt = st / n_tiles;
    nt = st - mt * n_tiles;
  } else {
    const int mgroups = (m_tiles + CL - 1) / CL;
    nt = st / mgroups;
    mt = (st - nt * mgroups) * CL + rank;  // may be >= m_tiles in the ragged last group: all-zero A, nothing stored
  }
}

template <int BN, int STAGES, int CL, bool F16, int EW, bool T9>
__global__ void __launch_bounds__(persistent_threads(EW), 1)
conv_gemm_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                            const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR,
                            const __grid_constant__ CUtensorMap tmOR,
                            const ConvParams p, const int m_tiles, const int n_tiles, const int num_super,
                            const int share, const int epi_tma) {
  using L = SmemLayoutP<BN, STAGES, F16, EW, T9>;
  static_assert(!T9 || CL == 1, "tap-reuse mode has no cluster variant");
  constexpr int NRING = L::NRING;
  constexpr int kEpiWarps = EW;
  constexpr int EWH = EW / 4;  // epilogue warps per TMEM lane quarter
  constexpr int NBUF = L::NBUF;
  constexpr uint32_t TMEM_COLS = (2 * BN) < 32 ? 32 : (2 * BN);
  static_assert(2 * BN <= 512, "double-buffered accumulator must fit the 512 TMEM columns");
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CL) - 1);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + NRING;    // T9: [0, NA) the A' ring, [NA, NA + NB) the weight ring
  uint64_t* tmem_full = empty_bar + NRING;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint64_t* res_bar = tmem_empty + 2;        // [kEpiWarps][3] (F16 only): residual tiles landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + (F16 ? 3 * kEpiWarps : 0));
  const int S = CL > 1 ? 1 : p.splits;                   // K ranges per tile (host: 1 for clustered launches)
  const int num_items = num_super * S;

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int iters = T9 ? 3 * p.kblocks : p.taps * p.kblocks;  // T9: (dy, k-block) macro-iterations of three taps
  const int rank = CL > 1 ? static_cast<int>(tc05::cluster_ctarank()) : 0;
  const int cluster_id = blockIdx.x / CL;
  const int num_clusters = gridDim.x / CL;

  if (warp == 0 && lane == 0) {
    tc05::prefetch_tmap(&tmA);
    tc05::prefetch_tmap(&tmB);
    for (int s = 0; s < NRING; ++s) {
      tc05::mbar_init(&full_bar[s], 1);
      tc05::mbar_init(&empty_bar[s], CL);  // one tcgen05.commit arrival from every CTA of the cluster
    }
    for (int b = 0; b < 2; ++b) {
      tc05::mbar_init(&tmem_full[b], 1);
      tc05::mbar_init(&tmem_empty[b], kEpiWarps);
    }
    if (F16) {
      for (int b = 0; b < 3 * kEpiWarps; ++b) tc05::mbar_init(&res_bar[b], 1);
      if (epi_tma) {
        tc05::prefetch_tmap(&tmO);
        if (p.residual) tc05::prefetch_tmap(&tmR);
        if (p.out_relu) tc05::prefetch_tmap(&tmOR);
      }
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
        if constexpr (T9) {
          // macro-iteration j = dy * kblocks + kb: one A' box (rows m0 + (dy-1)*wp - 1 ... + 135), then the three
          // weight tiles of taps (dy, dx = -1, 0, +1); `it` counts A' boxes, 3 * it + dx weight tiles
          for (int j = static_cast<int>(static_cast<int64_t>(split) * iters / S); j < j1; ++j, ++it) {
            const int dy = j / p.kblocks;
            const int kb = j - dy * p.kblocks;
            const int sa = it % L::NA;
            tc05::mbar_wait(&empty_bar[sa], ((it / L::NA) & 1) ^ 1, p.err, 116);
            tc05::mbar_arrive_expect_tx(&full_bar[sa], A3_ROWS * 128);
            tc05::tma_load_2d(smem + sa * A3_BYTES, &tmA, &full_bar[sa], p.in_coff + kb * p.bk,
                              static_cast<int32_t>(m0 + static_cast<int64_t>(dy - 1) * wp - 1));
#pragma unroll 1
            for (int dx = 0; dx < 3; ++dx) {
              const int ib = 3 * it + dx;
              const int sb = L::NA + ib % L::NB;
              tc05::mbar_wait(&empty_bar[sb], ((ib / L::NB) & 1) ^ 1, p.err, 117);
              tc05::mbar_arrive_expect_tx(&full_bar[sb], L::B_BYTES);
              tc05::tma_load_2d(smem + L::NA * A3_BYTES + (ib % L::NB) * L::B_BYTES, &tmB, &full_bar[sb], kb * p.bk,
                                (dy * 3 + dx) * p.cout_pad + n0);
            }
          }
          continue;
        }
        for (int j = static_cast<int>(static_cast<int64_t>(split) * iters / S); j < j1; ++j, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          const int t = j / p.kblocks;
          const int kb = j - t * p.kblocks;
          int64_t row = m0;
          if (p.taps == 9) row += static_cast<int64_t>(t / 3 - 1) * wp + (t % 3 - 1);
          else if (p.taps == 4) row += static_cast<int64_t>(t - 2) * wp;  // vertical taps dy = -2..1 (space-to-depth stem)
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
        if constexpr (T9) {
          for (int j = j0; j < j1; ++j, ++it) {
            const int sa = it % L::NA;
            tc05::mbar_wait(&full_bar[sa], (it / L::NA) & 1, p.err, 118);
            const uint32_t a_addr = tc05::smem_u32(smem + sa * A3_BYTES);
#pragma unroll 1
            for (int dx = 0; dx < 3; ++dx) {
              const int ib = 3 * it + dx;
              const int sb = L::NA + ib % L::NB;
              tc05::mbar_wait(&full_bar[sb], (ib / L::NB) & 1, p.err, 119);
              tc05::fence_after_sync();
              // tap dx reads rows dx .. dx + 127 of the A' tile: the descriptor starts dx rows (128 B each) in
              const uint64_t da = tc05::make_desc_sw128(a_addr + dx * 128);
              const uint64_t db = tc05::make_desc_sw128(tc05::smem_u32(smem + L::NA * A3_BYTES + (ib % L::NB) * L::B_BYTES));
              if (F16) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  tc05::umma_f16_ss(d, da + 2 * k, db + 2 * k, idesc, (j != j0 || dx != 0 || k != 0) ? 1u : 0u);
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  tc05::umma_tf32_ss(d, da + 2 * k, db + 2 * k, idesc, (j != j0 || dx != 0 || k != 0) ? 1u : 0u);
              }
              tc05::umma_commit(&empty_bar[sb]);
            }
            tc05::umma_commit(&empty_bar[sa]);
          }
          tc05::umma_commit(&tmem_full[buf]);
          continue;
        }
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
    uint8_t* const wbytes = smem + L::STG_OFF + (warp - 2) * L::PER_WARP;  // this warp's epilogue bytes
    float* stg = reinterpret_cast<float*>(wbytes);
    uint64_t* const rbar = res_bar + (F16 ? (warp - 2) * 3 : 0);
    uint32_t rphase = 0;        // bit b: parity the next wait on rbar[b] expects
    bool tma_dirty = false;     // bulk stores of this warp may still read its io tiles
    uint32_t gstep = 0;         // 64-channel TMA steps of this warp so far: step g uses io tile / barrier (g % NBUF)
    // Look-ahead over this warp's TMA steps (tiles of real channels only, in the order the loop below visits them):
    // the residual rows of step g + NBUF - 1 are requested when step g begins.
    constexpr int NSW = F16 ? (BN / 64 / EWH > 0 ? BN / 64 / EWH : 1) : 1;  // column steps of this warp per tile
    int la_item = cluster_id, la_mine = 0;
    uint32_t la_g = 0;          // step index the look-ahead position will get
    const bool res_tma = F16 && epi_tma != 0 && p.residual != nullptr && BN >= 64 * EWH;
    auto la_issue = [&]() {     // issue the residual load of the look-ahead position (if any), then advance it
      while (la_item < num_items) {
        int mt2, nt2;
        tile_of<CL>(la_item, rank, share, m_tiles, n_tiles, mt2, nt2);  // (S == 1 whenever epi_tma is set)
        if ((p.cout - nt2 * BN) < BN) {  // not a TMA tile: no step, nothing to request
          la_item += num_clusters;
          la_mine = 0;
          continue;
        }
        if (lane == 0) {
          const int bb = la_g % NBUF;
          tc05::mbar_arrive_expect_tx(&rbar[bb], kEpiTile);
          tc05::tma_load_2d(wbytes + bb * kEpiTile, &tmR, &rbar[bb], p.res_coff + nt2 * BN + (half + la_mine * EWH) * 64,
                            static_cast<int32_t>(static_cast<int64_t>(mt2) * BM + q * 32));
        }
        ++la_g;
        if (++la_mine == NSW) {
          la_mine = 0;
          la_item += num_clusters;
        }
        return;
      }
    };
    if (res_tma) {
#pragma unroll 1
      for (int d = 0; d < NBUF - 1; ++d) la_issue();  // the first NBUF - 1 steps' rows, before any MMA is waited for
    }
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
      // fp16 maps, whole tiles of real channels: 64-channel steps through TMA (residual in, output out)
      const bool tma_tile = F16 && BN >= 64 * EWH && epi_tma != 0 && (p.cout - n0) >= BN;
      const bool has_res = p.residual != nullptr;
      if (tma_tile) {
        // (residual rows were requested NBUF - 1 steps ago by the look-ahead)
      } else if (tma_dirty) {  // the generic path stages through the same bytes: drain the bulk stores first
        if (lane == 0) tc05::bulk_wait_group_read<0>();
        __syncwarp();
        tma_dirty = false;
      }
      // (a generic tile inside a TMA launch is a ragged channel tile; it has no residual in flight of its own, and the
      //  look-ahead's requests for LATER tiles land in io tiles the generic staging tile overlaps: the host only
      //  enables the TMA epilogue with a residual when every channel tile is whole — see launch_persistent_cl)
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
        // partial tiles are indexed by (channel tile, row tile) whatever order the tiles are visited in
        float* mine = p.sk_ws + (((static_cast<int64_t>(nt) * m_tiles + mt) * S + split) * BM + q * 32 + lane) * BN;
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN; c0 += EWH * 32) {
          uint32_t v[32];
          tc05::tmem_ld32(tacc + c0, v);
          tc05::tmem_ld_wait();
          if (c0 + EWH * 32 >= BN) {
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
      if (tma_tile) {
        constexpr int NS = BN / 64 / EWH;  // column steps of THIS warp per tile
        const bool relu2 = p.out_relu != nullptr;
#pragma unroll 1
        for (int mine = 0; mine < NS; ++mine, ++gstep) {
          const int sidx = half + mine * EWH;
          const int b = gstep % NBUF;
          uint8_t* io = wbytes + b * kEpiTile;
          uint8_t* orl = wbytes + NBUF * kEpiTile;
          // every bulk store this warp has issued must have finished READING shared memory: the io tile of step
          // g - 1 is about to receive the residual rows of step g + NBUF - 1, the single orl tile is rewritten below,
          // and io[b] itself was last stored from NBUF steps ago
          if (lane == 0) tc05::bulk_wait_group_read<0>();
          __syncwarp();
          if (res_tma) la_issue();
          const int ncol0 = n0 + sidx * 64;
          BiasHalf bh0, bh1;
          conv_epilogue_load_bias(bh0, 0, ncol0, p);  // in flight under the waits below
          conv_epilogue_load_bias(bh1, 1, ncol0, p);
          if (has_res) {
            tc05::mbar_wait(&rbar[b], (rphase >> b) & 1u, p.err, 115);
            rphase ^= 1u << b;
          }
          uint32_t v[32];
          load_acc(sidx * 64, v);
          if (relu2) conv_epilogue_tma_half<true>(v, bh0, 0, io, orl, has_res, interior, lane, ncol0, p);
          else conv_epilogue_tma_half<false>(v, bh0, 0, io, orl, has_res, interior, lane, ncol0, p);
          load_acc(sidx * 64 + 32, v);
          if (mine + 1 == NS) release_acc();
          if (relu2) conv_epilogue_tma_half<true>(v, bh1, 1, io, orl, has_res, interior, lane, ncol0, p);
          else conv_epilogue_tma_half<false>(v, bh1, 1, io, orl, has_res, interior, lane, ncol0, p);
          tc05::fence_proxy_async();  // generic-proxy writes of every lane -> visible to the bulk copy
          __syncwarp();
          if (lane == 0) {
            tc05::tma_store_2d(&tmO, io, p.out_coff + ncol0, static_cast<int32_t>(row0));
            if (relu2) tc05::tma_store_2d(&tmOR, orl, p.out_relu_coff + ncol0, static_cast<int32_t>(row0));
            tc05::bulk_commit_group();
          }
        }
        tma_dirty = true;
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

  // bulk stores issued by the epilogue warps' lane 0 must be COMPLETE (global writes performed, shared
  // memory no longer read) before the CTA exits
  if (F16 && warp >= 2 && lane == 0) tc05::bulk_wait_group<0>();
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

template <int BN, int STAGES, int CL, bool F16, int EW = 4, bool T9 = false>
int launch_persistent_cl(const mivos_conv_args* a, const ConvParams& p, int m_tiles, int n_tiles, int share,
                         cudaStream_t stream) {
  using L = SmemLayoutP<BN, STAGES, F16, EW, T9>;
  constexpr int smem_bytes = L::TOTAL + 1024;
  static_assert(smem_bytes <= 232448, "stage ring + epilogue staging exceed the 227 KB a CTA may use");
  static_assert(EW == 4 || (F16 && CL == 1), "8 epilogue warps: fp16 operand kernels without clusters only");
  constexpr int kPersistentThreads = persistent_threads(EW);
  auto kernel = conv_gemm_persistent_kernel<BN, STAGES, CL, F16, EW, T9>;
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
                          static_cast<uint64_t>(a->in_cstride), p.bk, T9 ? A3_ROWS : (share == SHARE_A ? BM / CL : BM), eb);
  if (rc != MIVOS_OK) return rc;
  rc = encode_tmap_2d(&tmB, a->weight, static_cast<uint64_t>(a->taps) * a->cout_pad, static_cast<uint64_t>(a->cin_pad),
                      static_cast<uint64_t>(a->cin_pad), p.bk, share == SHARE_B ? BN / CL : BN, eb);
  if (rc != MIVOS_OK) return rc;
  // fp16 output maps go through the TMA epilogue: 32-row x 64-channel boxes of the output map (and of the
  // residual / ReLU-copy maps), 128B-swizzled.  MIVOS_CONV_TMA_EPILOGUE=0: register path (A/B measurements).
  static const bool allow_tma_epi = [] {
    const char* e = getenv("MIVOS_CONV_TMA_EPILOGUE");
    return !(e && e[0] == '0');
  }();
  CUtensorMap tmO{}, tmR{}, tmOR{};
  int epi_tma = 0;
  // (with a residual every channel tile must be whole: the look-ahead's residual loads land in the bytes a ragged
  //  tile's register path would stage through)
  if (F16 && CL == 1 && allow_tma_epi && a->out_f16 && p.splits == 1 && BN >= (EW == 8 ? 128 : 64) && a->cout >= BN &&
      (!a->residual || a->cout % BN == 0)) {
    epi_tma = 1;
    rc = encode_tmap_2d(&tmO, a->out, static_cast<uint64_t>(p.rows), static_cast<uint64_t>(a->out_cstride),
                        static_cast<uint64_t>(a->out_cstride), 64, 32, 2);
    if (rc != MIVOS_OK) return rc;
    if (a->residual) {
      rc = encode_tmap_2d(&tmR, a->residual, static_cast<uint64_t>(p.rows), static_cast<uint64_t>(a->res_cstride),
                          static_cast<uint64_t>(a->res_cstride), 64, 32, 2);
      if (rc != MIVOS_OK) return rc;
    }
    if (a->out_relu) {
      rc = encode_tmap_2d(&tmOR, a->out_relu, static_cast<uint64_t>(p.rows), static_cast<uint64_t>(a->out_relu_cstride),
                          static_cast<uint64_t>(a->out_relu_cstride), 64, 32, 2);
      if (rc != MIVOS_OK) return rc;
    }
  }
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
  MIVOS_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, tmO, tmR, tmOR, p, m_tiles, n_tiles, num_super, share, epi_tma));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return MIVOS_OK;
}

inline int tile_order_share() {  // SHARE_NONE: channel tiles of a row tile consecutive (default); SHARE_ROWS: row tiles fastest
  static const int v = [] {
    const char* e = getenv("MIVOS_CONV_TILE_ORDER");
    return (e && (e[0] == 'm' || e[0] == 'M')) ? static_cast<int>(SHARE_ROWS) : static_cast<int>(SHARE_NONE);
  }();
  return v;
}

inline bool tap3_enabled() {
  static const bool on = [] {
    const char* e = getenv("MIVOS_CONV_TAP3");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <int BN, int STAGES, bool F16>
int launch_persistent(const mivos_conv_args* a, const ConvParams& p, cudaStream_t stream) {
  const int m_tiles = static_cast<int>(ceil_div64(p.rows, BM));
  const int n_tiles = a->cout_pad / BN;
  if constexpr (F16) {  // operand multicast (measured: no gain) is only instantiated for the TF32 kernels
    // 8 epilogue warps for the 128-wide fp16 tiles (the output-bound 1x1 expansions land here); MIVOS_CONV_EPI8=0: A/B
    static const bool epi8 = [] {
      const char* e = getenv("MIVOS_CONV_EPI8");
      return !(e && e[0] == '0');
    }();
    // 3x3 layers: one activation box per (k-block, dy) serves three taps (MIVOS_CONV_TAP3=0: one box per tap, A/B)
    const bool t9 = a->taps == 9 && tap3_enabled();
    if constexpr (BN == 128) {
      if (epi8) {
        if (t9) return launch_persistent_cl<BN, STAGES, 1, F16, 8, true>(a, p, m_tiles, n_tiles, tile_order_share(), stream);
        return launch_persistent_cl<BN, STAGES, 1, F16, 8, false>(a, p, m_tiles, n_tiles, tile_order_share(), stream);
      }
    }
    if (t9) return launch_persistent_cl<BN, STAGES, 1, F16, 4, true>(a, p, m_tiles, n_tiles, tile_order_share(), stream);
    return launch_persistent_cl<BN, STAGES, 1, F16, 4, false>(a, p, m_tiles, n_tiles, tile_order_share(), stream);
  } else {
    const ClusterChoice c = choose_cluster(m_tiles, n_tiles);
    if (c.cl == 1 && a->taps == 9 && tap3_enabled())
      return launch_persistent_cl<BN, STAGES, 1, false, 4, true>(a, p, m_tiles, n_tiles, tile_order_share(), stream);
    switch (c.cl) {
      case 4: return launch_persistent_cl<BN, STAGES, 4, false>(a, p, m_tiles, n_tiles, c.share, stream);
      case 2: return launch_persistent_cl<BN, STAGES, 2, false>(a, p, m_tiles, n_tiles, c.share, stream);
      default: return launch_persistent_cl<BN, STAGES, 1, false>(a, p, m_tiles, n_tiles, tile_order_share(), stream);
    }
  }
}
