// Internal interface between the memory-read translation units (memread.cu, memread_tc.cu).
#pragma once
#include "host_util.h"

namespace mivos {

constexpr int kMaxSplits = 16;
constexpr int kTcCandCap = 256;  // candidates one (query, split) may emit on the tcgen05 path

struct MemreadPlan {
  int algo;
  int qtile;      // queries per CTA
  int slot_tile;  // bank slots per inner tile
  int qtiles;
  int splits;  // CTAs along the memory axis
  int tiles_per_split;
  int kcap;  // list capacity per (object, query, split)
  int64_t off_score, off_idx, off_cnt, off_flag, bytes;
};

MemreadPlan memread_plan(int k_objects, int64_t slots, int hw, int top_k, int algo);

int launch_exact_candidates(const float* bank_k, int64_t slots_cap, int k_objects, int64_t slots,
                            const float* qk, int hw, int top_k, const MemreadPlan& pl, void* ws,
                            cudaStream_t stream);
int launch_select(const float* bank_k, const float* bank_v, int64_t slots_cap, int k_objects,
                  const float* qk, int hw, int top_k, const MemreadPlan& pl, void* ws, int rescore,
                  const float* margin, float* out, int out_cstride, int out_coff, int halo_h,
                  int halo_w, int32_t* topk_idx, float* topk_val, cudaStream_t stream);

bool memread_tc_available();
int memread_tc_run(const float* bank_k, const float* bank_v, int64_t slots_cap, int k_objects,
                   int64_t slots, const float* qk, int hw, int top_k, float* out, int out_cstride,
                   int out_coff, int halo_h, int halo_w, int32_t* topk_idx, float* topk_val,
                   void* workspace, cudaStream_t stream);

}  // namespace mivos
