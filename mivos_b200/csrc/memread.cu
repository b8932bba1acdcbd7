// Space-time memory read: EvalMemoryReader.forward + softmax_w_g_top of the reference
// (model/propagation/prop_net.py:47-73, 81-108) without ever materialising the [slots, hw]
// affinity matrix.
//
// Two stages:
//   A. candidate generation over the bank, split over CTAs along the memory axis.  Each
//      (object, query, split) produces a short list of (score, slot) candidates:
//        A1 `memread_exact_kernel`  : fp32 FMA scores on CUDA cores, exact per-split top-k
//                                     (this file; also the overflow fallback of A2)
//        A2 `memread_tc_kernel`     : TF32 tcgen05 scores + conservative threshold with an error
//                                     margin (memread_tc.cu); survivors are re-scored exactly in B
//   B. `memread_select_kernel`: per (object, query) merge candidates, (re-score,) exact top-k with
//      a total order (score desc, slot asc), softmax over the k survivors
//      (exp(s - s_max) / sum, prop_net.py:55-58), value read-out as a k-row gather of the
//      slot-major value bank (algorithmically 2*k*512 flops/query instead of the reference's
//      dense 2*slots*512).
//
// Score definition (both paths, bit-identical): s = sum_{c=0..127} key[c] * (q[c] / sqrt(128)),
// accumulated with fmaf in ascending channel order; q/sqrt(128) is an IEEE fp32 division exactly
// as `qk / math.sqrt(CK)` at prop_net.py:86.
#include "host_util.h"
#include "pdl.cuh"
#include "memread.h"

#include <cuda_fp16.h>
#include <atomic>

namespace mivos {
extern std::atomic<int64_t> g_launches;

MemreadPlan memread_plan(int k_objects, int64_t slots, int hw, int top_k, int algo) {
  MemreadPlan pl;
  pl.algo = algo;
  if (algo == MIVOS_MEMREAD_TCGEN05) {
    pl.qtile = 128;
    pl.slot_tile = 256;
    pl.kcap = kTcCandCap;
  } else {
    pl.qtile = 32;
    pl.slot_tile = 128;
    pl.kcap = top_k;
  }
  pl.qtiles = ceil_div(hw, pl.qtile);
  const int64_t tiles = ceil_div64(slots, pl.slot_tile);
  // exact path: ~2 CTAs per SM; tcgen05 path: ONE wave of 148 CTAs (192 KB smem = 1 CTA/SM);
  // every split owns at least 4 slot tiles
  int64_t want = algo == MIVOS_MEMREAD_TCGEN05 ? 148 / (static_cast<int64_t>(pl.qtiles) * k_objects)
                                               : ceil_div64(296, static_cast<int64_t>(pl.qtiles) * k_objects);
  int64_t max_by_tiles = tiles / 4 > 0 ? tiles / 4 : 1;
  int64_t s = want < max_by_tiles ? want : max_by_tiles;
  if (s < 1) s = 1;
  if (s > kMaxSplits) s = kMaxSplits;
  pl.tiles_per_split = static_cast<int>(ceil_div64(tiles, s));
  pl.splits = static_cast<int>(ceil_div64(tiles, pl.tiles_per_split));
  pl.nlists = algo == MIVOS_MEMREAD_TCGEN05 ? pl.splits * kTcHalves : pl.splits;
  const int64_t lists = static_cast<int64_t>(k_objects) * hw * pl.nlists;
  pl.off_list = 0;
  pl.off_cnt = lists * pl.kcap * 8;
  pl.off_flag = pl.off_cnt + lists * 4;
  pl.bytes = pl.off_flag + static_cast<int64_t>(k_objects) * hw * 4;
  pl.bytes = (pl.bytes + 255) & ~255ll;
  return pl;
}

namespace {

constexpr float kSqrtCK = 11.313708498984761f;  // sqrt(128) rounded to fp32, as the reference's divisor

constexpr int A1_THREADS = 256;
constexpr int A1_Q = 32;
constexpr int A1_S = 128;
constexpr int KS_STRIDE = 132;  // padded key-row stride (floats): conflict-free LDS.128 by row
constexpr int MAXK = 64;

struct A1Smem {
  float qs[A1_Q][128];
  float ks[A1_S][KS_STRIDE];
  float sc[A1_Q][A1_S + 1];
  float list_s[A1_Q][MAXK];
  int list_i[A1_Q][MAXK];
  float pend_s[A1_Q][A1_S];
  int pend_i[A1_Q][A1_S];
  int pend_cnt[A1_Q];
  int list_cnt[A1_Q];
};

// (score desc, slot asc) total order: true if a ranks before b
__device__ __forceinline__ bool before(float sa, int ia, float sb, int ib) {
  return (sa > sb) || (sa == sb && ia < ib);
}

__global__ void __launch_bounds__(A1_THREADS, 1)
memread_exact_kernel(const float* __restrict__ bank_k, int64_t slots_cap, int64_t slots,
                     const float* __restrict__ qk, int hw, int top_k, int tiles_per_split,
                     int splits, int2* __restrict__ cand,
                     int* __restrict__ cand_cnt, const int* __restrict__ flags,
                     const int* __restrict__ dyn_slots) {
  mivos::pdl_prologue();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  A1Smem& sm = *reinterpret_cast<A1Smem*>(smem_raw);
  const int tid = threadIdx.x;
  const int q0 = blockIdx.x * A1_Q;
  const int split = blockIdx.y;
  const int obj = blockIdx.z;
  if (flags) {  // fallback mode: only tiles that own an overflowed query do any work
    const int mine = (tid < A1_Q && q0 + tid < hw) ? flags[static_cast<int64_t>(obj) * hw + q0 + tid] : 0;
    if (!__syncthreads_or(mine)) return;
  }
  const float* keys = bank_k + static_cast<int64_t>(obj) * slots_cap * 128;

  // queries, pre-divided by sqrt(CK) (prop_net.py:86)
  for (int i = tid; i < A1_Q * 128; i += A1_THREADS) {
    const int q = i >> 7, c = i & 127;
    sm.qs[q][c] = (q0 + q < hw) ? qk[static_cast<int64_t>(q0 + q) * 128 + c] / kSqrtCK : 0.f;
  }
  if (tid < A1_Q) {
    sm.pend_cnt[tid] = 0;
    sm.list_cnt[tid] = 0;
  }
  __syncthreads();

  if (dyn_slots) {  // device-side slot count (CUDA-graph replay): fixed grid, ranges derived here
    slots = *dyn_slots;
    const int tiles = static_cast<int>((slots + A1_S - 1) / A1_S);
    tiles_per_split = (tiles + splits - 1) / splits;
  }
  const int64_t s_begin = static_cast<int64_t>(split) * tiles_per_split * A1_S;
  int64_t s_end = s_begin + static_cast<int64_t>(tiles_per_split) * A1_S;
  if (s_end > slots) s_end = slots;

  for (int64_t s0 = s_begin; s0 < s_end; s0 += A1_S) {
    // ---- stage 128 key rows (coalesced float4), zero rows past the end
    for (int i = tid; i < A1_S * 32; i += A1_THREADS) {
      const int r = i >> 5, c4 = i & 31;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s0 + r < s_end) v = *reinterpret_cast<const float4*>(keys + (s0 + r) * 128 + c4 * 4);
      *reinterpret_cast<float4*>(&sm.ks[r][c4 * 4]) = v;
    }
    __syncthreads();
    // ---- 128 slots x 32 queries of fp32 dot products; thread = (slot, 16-query half)
    {
      const int sl = tid & 127, qh = (tid >> 7) * 16;
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll 4
      for (int c4 = 0; c4 < 32; ++c4) {
        const float4 kv = *reinterpret_cast<const float4*>(&sm.ks[sl][c4 * 4]);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 qv = *reinterpret_cast<const float4*>(&sm.qs[qh + j][c4 * 4]);
          acc[j] = fmaf(kv.x, qv.x, acc[j]);
          acc[j] = fmaf(kv.y, qv.y, acc[j]);
          acc[j] = fmaf(kv.z, qv.z, acc[j]);
          acc[j] = fmaf(kv.w, qv.w, acc[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) sm.sc[qh + j][sl] = acc[j];
    }
    __syncthreads();
    // ---- threshold filter: 8 threads per query, 16 slots each
    {
      const int q = tid >> 3, sub = tid & 7;
      const int cnt = sm.list_cnt[q];
      const float thr = (cnt >= top_k) ? sm.list_s[q][top_k - 1] : -INFINITY;
#pragma unroll 4
      for (int j = 0; j < 16; ++j) {
        const int sl = sub * 16 + j;
        const float s = sm.sc[q][sl];
        if (s0 + sl < s_end && (cnt < top_k || s > thr)) {
          const int pos = atomicAdd(&sm.pend_cnt[q], 1);
          sm.pend_s[q][pos] = s;
          sm.pend_i[q][pos] = static_cast<int>(s0 + sl);
        }
      }
    }
    __syncthreads();
    // ---- merge pending candidates into the sorted per-query list (one thread per query)
    if (tid < A1_Q) {
      const int q = tid;
      int cnt = sm.list_cnt[q];
      const int np = sm.pend_cnt[q];
      for (int p = 0; p < np; ++p) {
        const float s = sm.pend_s[q][p];
        const int id = sm.pend_i[q][p];
        if (cnt == top_k && !before(s, id, sm.list_s[q][cnt - 1], sm.list_i[q][cnt - 1])) continue;
        int pos = (cnt < top_k) ? cnt : cnt - 1;  // slot that will be overwritten / appended
        while (pos > 0 && before(s, id, sm.list_s[q][pos - 1], sm.list_i[q][pos - 1])) {
          sm.list_s[q][pos] = sm.list_s[q][pos - 1];
          sm.list_i[q][pos] = sm.list_i[q][pos - 1];
          --pos;
        }
        sm.list_s[q][pos] = s;
        sm.list_i[q][pos] = id;
        if (cnt < top_k) ++cnt;
      }
      sm.list_cnt[q] = cnt;
      sm.pend_cnt[q] = 0;
    }
    __syncthreads();
  }

  // ---- publish the per-split lists
  for (int i = tid; i < A1_Q * top_k; i += A1_THREADS) {
    const int q = i / top_k, j = i - q * top_k;
    if (q0 + q < hw && j < sm.list_cnt[q]) {
      const int64_t base = ((static_cast<int64_t>(obj) * hw + q0 + q) * splits + split) * top_k;
      cand[base + j] = make_int2(__float_as_int(sm.list_s[q][j]), sm.list_i[q][j]);
    }
  }
  if (tid < A1_Q && q0 + tid < hw)
    cand_cnt[(static_cast<int64_t>(obj) * hw + q0 + tid) * splits + split] = sm.list_cnt[tid];
}

// ------------------------------------------------------------------------------------------
// Stage B.  One CTA (128 threads) per (object, query).
constexpr int B_THREADS = 128;
constexpr int B_MAXSURV = 1024;  // candidates within the TF32 margin of the top-k that get re-scored

__device__ __forceinline__ float exact_score(const float* __restrict__ key, const float* qs) {
  float acc = 0.f;
#pragma unroll 16
  for (int c4 = 0; c4 < 32; ++c4) {
    const float4 kv = *reinterpret_cast<const float4*>(key + c4 * 4);
    acc = fmaf(kv.x, qs[c4 * 4 + 0], acc);
    acc = fmaf(kv.y, qs[c4 * 4 + 1], acc);
    acc = fmaf(kv.z, qs[c4 * 4 + 2], acc);
    acc = fmaf(kv.w, qs[c4 * 4 + 3], acc);
  }
  return acc;
}

struct SelectLists {
  const int2* e;  // {score bits, slot}
  const int* cnt;
  int splits;
  int kcap;
};

__global__ void __launch_bounds__(B_THREADS)
memread_select_kernel(const float* __restrict__ bank_k, const float* __restrict__ bank_v,
                      int64_t slots_cap, const float* __restrict__ qk, int hw, int top_k,
                      const SelectLists prim, const int prim_rescore, const SelectLists fb,
                      const int* __restrict__ flags, const float* __restrict__ qnorm,
                      const float* __restrict__ kmax2, void* __restrict__ out, int out_cstride,
                      int out_coff, int halo_h, int halo_w, int out_f16, int* __restrict__ topk_idx,
                      float* __restrict__ topk_val, int* err, const int max_cand) {
  mivos::pdl_prologue();
  extern __shared__ __align__(16) uint8_t sel_smem[];
  float* cs = reinterpret_cast<float*>(sel_smem);          // [max_cand] candidate scores
  int* ci = reinterpret_cast<int*>(cs + max_cand);         // [max_cand] candidate slots
  int* ei = ci + max_cand;                                 // [B_MAXSURV] slots of the survivors
  __shared__ float red_lo[4], red_hi[4];
  __shared__ int red_cnt[4];
  __shared__ float qs[128];
  __shared__ float top_s[MAXK];
  __shared__ int top_i[MAXK];
  __shared__ float top_w[MAXK];
  __shared__ int order[MAXK];
  __shared__ int offs[kMaxLists + 1];
  __shared__ int m_sh;

  const int tid = threadIdx.x;
  const int q = blockIdx.x, obj = blockIdx.y;
  const int64_t lq = static_cast<int64_t>(obj) * hw + q;
  const bool use_fb = flags != nullptr && flags[lq] != 0;  // overflowed on the tcgen05 path
  const SelectLists& L = use_fb ? fb : prim;
  const int rescore = use_fb ? 0 : prim_rescore;

  // ---- gather the candidates of all splits
  // list lengths -> exclusive prefix (one warp: the count loads are independent)
  if (tid < 32) {
    const int c = tid < L.splits ? L.cnt[lq * L.splits + tid] : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += up;
    }
    if (tid < L.splits) offs[tid] = incl - c;
    if (tid == L.splits - 1) offs[L.splits] = incl;
    if (tid == 0) m_sh = 0;
  }
  qs[tid] = qk[static_cast<int64_t>(q) * 128 + tid] / kSqrtCK;
  __syncthreads();
  int n = offs[L.splits];
  if (n > max_cand) {  // cannot happen with the capacities chosen by memread_plan
    if (tid == 0 && err) atomicExch(err, 201);
    n = max_cand;
  }
  // flat gather: thread i takes candidates i, i+128, ... of the concatenated lists, so all of a
  // thread's loads are independent (a per-list loop would be one dependent L2 round trip per list)
  for (int i = tid; i < n; i += B_THREADS) {
    int sp = 0;
    while (offs[sp + 1] <= i) ++sp;
    const int64_t src = (lq * L.splits + sp) * L.kcap + (i - offs[sp]);
    const int2 e = L.e[src];
    cs[i] = __int_as_float(e.x);
    ci[i] = e.y;
  }

  if (rescore) {
    // The candidate scores are TF32 approximations.  Find a value t whose rank is in
    // [top_k, 2*top_k+8] by bisection with block-wide counts (any t with count(>= t) >= top_k is a
    // lower bound of the top_k-th largest approximate score), keep the candidates that can still
    // belong to the exact top-k (approx >= t - 2*eps, same margin as the generator) and re-score
    // them with the exact fp32 FMA chain.
    __syncthreads();
    const int kk0 = top_k < n ? top_k : n;
    float lo = INFINITY, hi = -INFINITY;
    for (int i = tid; i < n; i += B_THREADS) {
      lo = fminf(lo, cs[i]);
      hi = fmaxf(hi, cs[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((tid & 31) == 0) {
      red_lo[tid >> 5] = lo;
      red_hi[tid >> 5] = hi;
    }
    __syncthreads();
    lo = fminf(fminf(red_lo[0], red_lo[1]), fminf(red_lo[2], red_lo[3]));
    hi = fmaxf(fmaxf(red_hi[0], red_hi[1]), fmaxf(red_hi[2], red_hi[3]));
    const int limit = 2 * kk0 + 8;
    for (int iter = 0; iter < 26 && n > limit; ++iter) {
      const float mid = 0.5f * lo + 0.5f * hi;
      if (!(mid > lo) || !(mid < hi)) break;  // interval exhausted (ties)
      int c = 0;
      for (int i = tid; i < n; i += B_THREADS) c += (cs[i] >= mid) ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      __syncthreads();  // previous iteration's readers of red_cnt are done
      if ((tid & 31) == 0) red_cnt[tid >> 5] = c;
      __syncthreads();
      c = red_cnt[0] + red_cnt[1] + red_cnt[2] + red_cnt[3];
      if (c >= kk0) {
        lo = mid;
        if (c <= limit) break;
      } else {
        hi = mid;
      }
    }
    const float cut = lo - kTcMarginFactor * qnorm[q] * sqrtf(kmax2[obj]);
    // survivors -> ei, in arbitrary order (the final ranking is a total order).  Compaction and
    // re-scoring are separate passes: survivors are scattered over the candidate list, and a warp
    // that re-scores inside the filter loop pays one exact_score latency (4 dependent L2 round
    // trips) per loop iteration that contains ANY survivor — ~17 of them per query at cfg-2.
    for (int i = tid; i < n; i += B_THREADS) {
      if (cs[i] >= cut) {
        const int pos = atomicAdd(&m_sh, 1);
        if (pos < B_MAXSURV) ei[pos] = ci[i];
      }
    }
    __syncthreads();
    n = m_sh;
    if (n > B_MAXSURV) {  // more than 1024 candidates within the TF32 margin of the top-k
      if (tid == 0 && err) atomicExch(err, 202);
      n = B_MAXSURV;
    }
    // exact fp32 scores, one survivor per thread (normally a single pass: n ~ 2k + margin hits);
    // the ranking below works on cs/ci (n <= B_MAXSURV <= capacity)
    for (int i = tid; i < n; i += B_THREADS) {
      const int slot = ei[i];
      cs[i] = exact_score(bank_k + (static_cast<int64_t>(obj) * slots_cap + slot) * 128, qs);
      ci[i] = slot;
    }
  }
  __syncthreads();

  // ---- exact top-k by rank counting under the (score desc, slot asc) total order
  const int kk = top_k < n ? top_k : n;
  for (int i = tid; i < n; i += B_THREADS) {
    const float si = cs[i];
    const int ii = ci[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += before(cs[j], ci[j], si, ii) ? 1 : 0;
    if (rank < kk) {
      top_s[rank] = si;
      top_i[rank] = ii;
    }
  }
  __syncthreads();
  // ---- softmax over the survivors: exp(s - s_0) / sum (prop_net.py:55-58)
  if (tid < kk) top_w[tid] = expf(top_s[tid] - top_s[0]);
  __syncthreads();
  float sum = 0.f;
  for (int j = 0; j < kk; ++j) sum += top_w[j];  // same sequential sum in every thread
  // accumulation order of the read-out: ascending slot index
  if (tid < kk) {
    int r = 0;
    for (int j = 0; j < kk; ++j) r += (top_i[j] < top_i[tid]) ? 1 : 0;
    order[r] = tid;
  }
  __syncthreads();
  if (topk_idx && tid < top_k) topk_idx[lq * top_k + tid] = tid < kk ? top_i[tid] : -1;
  if (topk_val && tid < top_k) topk_val[lq * top_k + tid] = tid < kk ? top_s[tid] : 0.f;

  // ---- read-out: thread owns 4 of the 512 value channels
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* vals = reinterpret_cast<const float4*>(bank_v + static_cast<int64_t>(obj) * slots_cap * 512);
  // loads in batches of 8 independent rows (one L2/HBM round trip per batch instead of per row);
  // the accumulation order stays ascending slot index
  for (int j0 = 0; j0 < kk; j0 += 8) {
    float4 v[8];
    float w[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int jj = j0 + u;
      const int j = order[jj < kk ? jj : kk - 1];
      w[u] = top_w[j] / sum;
      v[u] = vals[static_cast<int64_t>(top_i[j]) * 128 + tid];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (j0 + u < kk) {
        acc.x = fmaf(w[u], v[u].x, acc.x);
        acc.y = fmaf(w[u], v[u].y, acc.y);
        acc.z = fmaf(w[u], v[u].z, acc.z);
        acc.w = fmaf(w[u], v[u].w, acc.w);
      }
    }
  }
  int64_t row;
  if (halo_w > 0) {
    const int y = q / halo_w, x = q - y * halo_w;
    row = (static_cast<int64_t>(obj) * (halo_h + 2) + y + 1) * (halo_w + 2) + x + 1;
  } else {
    row = lq;
  }
  if (out_f16) {
    const __half2 h01 = __floats2half2_rn(acc.x, acc.y), h23 = __floats2half2_rn(acc.z, acc.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&h01);
    u.y = *reinterpret_cast<const uint32_t*>(&h23);
    *reinterpret_cast<uint2*>(static_cast<__half*>(out) + row * out_cstride + out_coff + tid * 4) = u;
  } else {
    *reinterpret_cast<float4*>(static_cast<float*>(out) + row * out_cstride + out_coff + tid * 4) = acc;
  }
}

}  // namespace

int launch_exact_candidates(const float* bank_k, int64_t slots_cap, int k_objects, int64_t slots,
                            const float* qk, int hw, int top_k, const MemreadPlan& pl, void* ws,
                            const int* flags, const int* dyn_slots, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    MIVOS_CUDA_OK(cudaFuncSetAttribute(memread_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(sizeof(A1Smem))));
    configured = true;
  }
  uint8_t* w = static_cast<uint8_t*>(ws);
  dim3 grid(pl.qtiles, pl.splits, k_objects);
  launch_pdl(memread_exact_kernel, grid, A1_THREADS, sizeof(A1Smem), stream, 
      bank_k, slots_cap, slots, qk, hw, top_k, pl.tiles_per_split, pl.splits,
      reinterpret_cast<int2*>(w + pl.off_list),
      reinterpret_cast<int*>(w + pl.off_cnt), flags, dyn_slots);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MIVOS_CUDA_OK(cudaGetLastError());
  return MIVOS_OK;
}

int launch_select(const float* bank_k, const float* bank_v, int64_t slots_cap, int k_objects,
                  const float* qk, int hw, int top_k, const MemreadPlan& pl, void* ws,
                  const MemreadPlan* fbp, void* fb_ws, const int* flags, const float* qnorm,
                  const float* kmax2, void* out, int out_cstride, int out_coff, int halo_h,
                  int halo_w, int out_f16, int32_t* topk_idx, float* topk_val, cudaStream_t stream) {
  uint8_t* w = static_cast<uint8_t*>(ws);
  SelectLists prim{reinterpret_cast<const int2*>(w + pl.off_list), reinterpret_cast<const int*>(w + pl.off_cnt),
                   pl.nlists, pl.kcap};
  SelectLists fb = prim;
  if (fbp) {
    uint8_t* f = static_cast<uint8_t*>(fb_ws);
    fb = SelectLists{reinterpret_cast<const int2*>(f + fbp->off_list), reinterpret_cast<const int*>(f + fbp->off_cnt),
                     fbp->nlists, fbp->kcap};
  }
  const int rescore = pl.algo == MIVOS_MEMREAD_TCGEN05 ? 1 : 0;
  int max_cand = rescore ? pl.nlists * kTcFinalCap : pl.nlists * pl.kcap;
  if (fbp && fbp->nlists * fbp->kcap > max_cand) max_cand = fbp->nlists * fbp->kcap;
  if (max_cand < B_MAXSURV) max_cand = B_MAXSURV;
  const int smem = max_cand * 8 + B_MAXSURV * 4;
  static int configured_smem = 0;
  if (smem > configured_smem) {
    MIVOS_CUDA_OK(cudaFuncSetAttribute(memread_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  dim3 grid(hw, k_objects);
  launch_pdl(memread_select_kernel, grid, B_THREADS, smem, stream, bank_k, bank_v, slots_cap, qk, hw, top_k, prim, rescore, fb,
                                                           fbp ? flags : nullptr, qnorm, kmax2, out, out_cstride,
                                                           out_coff, halo_h, halo_w, out_f16, topk_idx, topk_val,
                                                           device_error_flag(), max_cand);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MIVOS_CUDA_OK(cudaGetLastError());
  return MIVOS_OK;
}

}  // namespace mivos

using namespace mivos;

extern "C" MIVOS_API int64_t mivos_memory_read_workspace(int k_objects, int64_t slots, int hw, int top_k) {
  if (k_objects < 1 || slots < 1 || hw < 1 || top_k < 1 || top_k > MAXK) return -1;
  const MemreadPlan a = memread_plan(k_objects, slots, hw, top_k, MIVOS_MEMREAD_EXACT_SIMT);
  const MemreadPlan b = memread_plan(k_objects, slots, hw, top_k, MIVOS_MEMREAD_TCGEN05);
  // tcgen05 lists | exact lists (its overflow fallback) | scaled queries | their norms | key norms |
  // shared per-(object, query) threshold
  return b.bytes + a.bytes +
         (static_cast<int64_t>(hw) * 128 + ((hw + 63) & ~63) + 64 + static_cast<int64_t>(k_objects) * hw) * 4 + 1024;
}

extern "C" MIVOS_API int mivos_memory_read(const float* bank_k, const float* bank_v, int64_t slots_cap,
                                           int k_objects, int64_t slots, const float* qk, int hw,
                                           int top_k, void* out, int out_cstride, int out_coff,
                                           int out_halo_h, int out_halo_w, int32_t* topk_idx,
                                           float* topk_val, void* workspace, int64_t workspace_bytes,
                                           int algo, const int32_t* dyn_slots, int out_f16, mivos_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MIVOS_REQUIRE(bank_k && bank_v && qk && out && workspace, "memory_read: null pointer");
  MIVOS_REQUIRE(top_k >= 1 && top_k <= MAXK, "memory_read: top_k %d outside [1,%d]", top_k, MAXK);
  MIVOS_REQUIRE(slots >= top_k && slots <= slots_cap && slots < (1ll << 31) - 65536,
                "memory_read: bad slot count %lld (cap %lld, k %d)", (long long)slots, (long long)slots_cap, top_k);
  MIVOS_REQUIRE(k_objects >= 1 && hw >= 1, "memory_read: bad object/query count");
  MIVOS_REQUIRE(out_cstride % 4 == 0 && out_coff % 4 == 0 && out_coff + 512 <= out_cstride,
                "memory_read: output channel window does not fit");
  MIVOS_REQUIRE(out_halo_w == 0 || out_halo_h * out_halo_w == hw, "memory_read: halo dims do not match hw");
  MIVOS_REQUIRE((reinterpret_cast<uintptr_t>(bank_k) & 15) == 0 && (reinterpret_cast<uintptr_t>(bank_v) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                "memory_read: pointers must be 16-byte aligned (workspace 256)");
  const int64_t need = mivos_memory_read_workspace(k_objects, slots, hw, top_k);
  MIVOS_REQUIRE(workspace_bytes >= need, "memory_read: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
  if (algo == MIVOS_MEMREAD_AUTO) algo = memread_tc_available() ? MIVOS_MEMREAD_TCGEN05 : MIVOS_MEMREAD_EXACT_SIMT;

  if (algo == MIVOS_MEMREAD_EXACT_SIMT) {
    const MemreadPlan pl = memread_plan(k_objects, slots, hw, top_k, MIVOS_MEMREAD_EXACT_SIMT);
    int rc = launch_exact_candidates(bank_k, slots_cap, k_objects, slots, qk, hw, top_k, pl, workspace, nullptr, dyn_slots, stream);
    if (rc != MIVOS_OK) return rc;
    return launch_select(bank_k, bank_v, slots_cap, k_objects, qk, hw, top_k, pl, workspace, nullptr, nullptr,
                         nullptr, nullptr, nullptr, out, out_cstride, out_coff, out_halo_h, out_halo_w, out_f16, topk_idx,
                         topk_val, stream);
  }
  if (algo == MIVOS_MEMREAD_TCGEN05) {
    return memread_tc_run(bank_k, bank_v, slots_cap, k_objects, slots, qk, hw, top_k, out, out_cstride,
                          out_coff, out_halo_h, out_halo_w, out_f16, topk_idx, topk_val, workspace, dyn_slots, stream);
  }
  set_last_error("memory_read: unknown algo %d", algo);
  return MIVOS_ERR_INVALID;
}

// Debug/diagnostic (synchronises): candidate statistics of the last tcgen05-path read that used
// `workspace`: out[0] = total candidates after compaction, out[1] = max per (object, query),
// out[2] = queries flagged for the exact fallback, out[3] = splits.
extern "C" MIVOS_API int mivos_memory_read_stats(const void* workspace, int k_objects, int64_t slots, int hw,
                                                 int top_k, int64_t* out) {
  MIVOS_REQUIRE(workspace && out, "memory_read_stats: null pointer");
  const MemreadPlan tc = memread_plan(k_objects, slots, hw, top_k, MIVOS_MEMREAD_TCGEN05);
  const int64_t nq = static_cast<int64_t>(k_objects) * hw;
  int* cnt = new int[nq * tc.nlists];
  int* flg = new int[nq];
  const uint8_t* w = static_cast<const uint8_t*>(workspace);
  cudaError_t e1 = cudaMemcpy(cnt, w + tc.off_cnt, nq * tc.nlists * 4, cudaMemcpyDeviceToHost);
  cudaError_t e2 = cudaMemcpy(flg, w + tc.off_flag, nq * 4, cudaMemcpyDeviceToHost);
  int64_t total = 0, mx = 0, flagged = 0;
  if (e1 == cudaSuccess && e2 == cudaSuccess) {
    for (int64_t q = 0; q < nq; ++q) {
      int64_t s = 0;
      if (flg[q]) { ++flagged; continue; }
      for (int sp = 0; sp < tc.nlists; ++sp) s += cnt[q * tc.nlists + sp];
      total += s;
      if (s > mx) mx = s;
    }
  }
  delete[] cnt;
  delete[] flg;
  MIVOS_CUDA_OK(e1);
  MIVOS_CUDA_OK(e2);
  out[0] = total; out[1] = mx; out[2] = flagged; out[3] = tc.splits;
  return MIVOS_OK;
}
