// HBM-bound kernels of the propagation path: gathers that feed the implicit GEMM, pooling,
// bilinear resampling, layout conversion at the API boundary, soft aggregation and argmax.
// All are coalesced / 128-bit vectorised streaming kernels; none has data reuse worth staging
// beyond the transposes (which go through a padded shared-memory tile).
#include "host_util.h"

#include <atomic>

namespace mivos {
extern std::atomic<int64_t> g_launches;
namespace {

constexpr int kThreads = 256;

inline unsigned grid_for(int64_t work, int per_block = kThreads) {
  int64_t g = (work + per_block - 1) / per_block;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

#define MIVOS_LAUNCHED()                                   \
  do {                                                     \
    g_launches.fetch_add(1, std::memory_order_relaxed);    \
    MIVOS_CUDA_OK(cudaGetLastError());                     \
  } while (0)

// ------------------------------------------------------------------------------------------
// stem gather: 7x7 / stride 2 / pad 3 window of cat(frame, mask, others) -> im2col matrix.
// One thread per (row, k) with k fastest: writes coalesced, reads served by L1/L2.
template <int CIN>
__global__ void stem_gather_kernel(const float* __restrict__ frame, const float* __restrict__ masks,
                                   int kobj, int h, int w, float* __restrict__ out, int kpad) {
  // One warp per output row (HALO row of the half-resolution map): the row -> (image, y, x)
  // decomposition is done once per row, lanes sweep k so every store instruction writes 128
  // contiguous bytes; the divisions by CIN / 7 are by compile-time constants.
  const int ho = h / 2, wo = w / 2;
  const int wp = wo + 2;
  const int64_t per_img = static_cast<int64_t>(ho + 2) * wp;
  const int64_t rows = static_cast<int64_t>(kobj) * per_img;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int64_t plane = static_cast<int64_t>(h) * w;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    const int obj = static_cast<int>(r / per_img);
    const int rem = static_cast<int>(r - obj * per_img);
    const int yo = rem / wp - 1, xo = rem - (rem / wp) * wp - 1;
    const bool inside = yo >= 0 && yo < ho && xo >= 0 && xo < wo;
    float* orow = out + r * kpad;
    const float* fr = frame + (CIN == 3 ? static_cast<int64_t>(obj) * 3 * plane : 0);
    for (int k = lane; k < kpad; k += 32) {
      float v = 0.f;
      if (inside && k < 49 * CIN) {
        const int tap = k / CIN, c = k - tap * CIN;
        const int ky = tap / 7, kx = tap - ky * 7;
        const int y = 2 * yo + ky - 3, x = 2 * xo + kx - 3;
        if (y >= 0 && y < h && x >= 0 && x < w) {
          const int64_t pix = static_cast<int64_t>(y) * w + x;
          if (c < 3) {
            // CIN == 3: `obj` indexes a BATCH of frames; CIN == 5: one frame shared by all objects
            v = fr[c * plane + pix];
          } else if (c == 3) {
            v = masks[obj * plane + pix];
          } else {
            // "others": sum of the other objects' masks, in object order (prop_net.py:150-157)
            float s = 0.f;
            for (int j = 0; j < kobj; ++j)
              if (j != obj) s += masks[j * plane + pix];
            v = s;
          }
        }
      }
      orow[k] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// stride-2 gather from a HALO map into an im2col matrix (rows = HALO rows of the output map).
__global__ void gather_s2_kernel(const float4* __restrict__ in, int n, int h, int w, int c4,
                                 int in_cstride4, int ks, float4* __restrict__ out,
                                 int out_cstride4) {
  const int ho = h / 2, wo = w / 2;
  const int wpo = wo + 2, wpi = w + 2;
  const int kk = ks * ks;
  const int64_t rows = static_cast<int64_t>(n) * (ho + 2) * wpo;
  const int64_t total = rows * kk * c4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % c4);
    const int64_t t1 = i / c4;
    const int tap = static_cast<int>(t1 % kk);
    const int64_t r = t1 / kk;
    const int64_t per_img = static_cast<int64_t>(ho + 2) * wpo;
    const int img = static_cast<int>(r / per_img);
    const int rem = static_cast<int>(r - img * per_img);
    const int yo = rem / wpo - 1, xo = rem % wpo - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yo >= 0 && yo < ho && xo >= 0 && xo < wo) {
      const int ky = tap / ks, kx = tap - ky * ks;
      // input pixel (2*yo + ky - ks/2, 2*xo + kx - ks/2); +1 for the halo offset
      const int yi = 2 * yo + ky - ks / 2 + 1, xi = 2 * xo + kx - ks / 2 + 1;
      const int64_t rin = (static_cast<int64_t>(img) * (h + 2) + yi) * wpi + xi;
      v = in[rin * in_cstride4 + ci];
    }
    out[r * out_cstride4 + tap * c4 + ci] = v;
  }
}

// ------------------------------------------------------------------------------------------
__global__ void maxpool3x3s2_kernel(const float4* __restrict__ in, int n, int h, int w, int c4,
                                    float4* __restrict__ out) {
  const int ho = h / 2, wo = w / 2;
  const int wpo = wo + 2, wpi = w + 2;
  const int64_t total = static_cast<int64_t>(n) * ho * wo * c4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % c4);
    int64_t t1 = i / c4;
    const int xo = static_cast<int>(t1 % wo);
    t1 /= wo;
    const int yo = static_cast<int>(t1 % ho);
    const int img = static_cast<int>(t1 / ho);
    // inputs are post-ReLU (>= 0), so the zero halo is equivalent to the -inf pad of MaxPool2d
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yi = 2 * yo + ky, xi = 2 * xo + kx;  // halo coords of (2yo+ky-1, 2xo+kx-1)
        const float4 v = in[((static_cast<int64_t>(img) * (h + 2) + yi) * wpi + xi) * c4 + ci];
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    out[((static_cast<int64_t>(img) * (ho + 2) + yo + 1) * wpo + xo + 1) * c4 + ci] = m;
  }
}

// ------------------------------------------------------------------------------------------
// bilinear source index, align_corners=False (ATen area_pixel_compute_source_index)
__device__ __forceinline__ void bilin(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = static_cast<int>(src);
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - static_cast<float>(i0);
}

__global__ void upsample2x_add_kernel(float4* __restrict__ x, const float4* __restrict__ up, int n,
                                      int h, int w, int c4, float4* __restrict__ x_relu,
                                      const float4* __restrict__ skip) {
  const int hs = h / 2, ws = w / 2;
  const int64_t total = static_cast<int64_t>(n) * h * w * c4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % c4);
    int64_t t1 = i / c4;
    const int xo = static_cast<int>(t1 % w);
    t1 /= w;
    const int yo = static_cast<int>(t1 % h);
    const int img = static_cast<int>(t1 / h);
    int y0, y1, x0, x1;
    float ly, lx;
    bilin(yo, 0.5f, hs, y0, y1, ly);
    bilin(xo, 0.5f, ws, x0, x1, lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const int64_t base = static_cast<int64_t>(img) * (hs + 2);
    const float4 v00 = up[((base + y0 + 1) * (ws + 2) + x0 + 1) * c4 + ci];
    const float4 v01 = up[((base + y0 + 1) * (ws + 2) + x1 + 1) * c4 + ci];
    const float4 v10 = up[((base + y1 + 1) * (ws + 2) + x0 + 1) * c4 + ci];
    const float4 v11 = up[((base + y1 + 1) * (ws + 2) + x1 + 1) * c4 + ci];
    const int64_t o = ((static_cast<int64_t>(img) * (h + 2) + yo + 1) * (w + 2) + xo + 1) * c4 + ci;
    // `skip` (batch 1, broadcast over images) replaces x as the addend: x = skip + up2x(up)
    float4 xv = skip ? skip[(static_cast<int64_t>(yo + 1) * (w + 2) + xo + 1) * c4 + ci] : x[o];
    xv.x += hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
    xv.y += hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
    xv.z += hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z);
    xv.w += hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w);
    x[o] = xv;
    if (x_relu) {
      x_relu[o] = make_float4(fmaxf(xv.x, 0.f), fmaxf(xv.y, 0.f), fmaxf(xv.z, 0.f), fmaxf(xv.w, 0.f));
    }
  }
}

// ------------------------------------------------------------------------------------------
// [channels x pixels] <-> [pixels x channels] transposes through a padded 32x32 smem tile.
// grid: (pixel tiles, channel tiles, planes)
__global__ void halo_to_nchw_kernel(const float* __restrict__ halo, int h, int w, int cstride,
                                    int coff, int c, float* __restrict__ nchw) {
  __shared__ float tile[32][33];
  const int img = blockIdx.z;
  const int hw = h * w;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int p = p0 + j, ch = c0 + tx;
    float v = 0.f;
    if (p < hw && ch < c) {
      const int y = p / w, x = p - y * w;
      v = halo[((static_cast<int64_t>(img) * (h + 2) + y + 1) * (w + 2) + x + 1) * cstride + coff + ch];
    }
    tile[j][tx] = v;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int ch = c0 + j, p = p0 + tx;
    if (p < hw && ch < c) nchw[(static_cast<int64_t>(img) * c + ch) * hw + p] = tile[tx][j];
  }
}

__global__ void nchw_to_halo_kernel(const float* __restrict__ nchw, int h, int w, int c,
                                    float* __restrict__ halo, int cstride, int coff, int relu) {
  __shared__ float tile[32][33];
  const int img = blockIdx.z;
  const int hw = h * w;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int ch = c0 + j, p = p0 + tx;
    float v = 0.f;
    if (p < hw && ch < c) v = nchw[(static_cast<int64_t>(img) * c + ch) * hw + p];
    tile[j][tx] = relu ? fmaxf(v, 0.f) : v;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int p = p0 + j, ch = c0 + tx;
    if (p < hw && ch < c) {
      const int y = p / w, x = p - y * w;
      halo[((static_cast<int64_t>(img) * (h + 2) + y + 1) * (w + 2) + x + 1) * cstride + coff + ch] = tile[tx][j];
    }
  }
}

// ------------------------------------------------------------------------------------------
__global__ void bank_write_kernel(const float4* __restrict__ halo, int kobj, int h, int w,
                                  int cstride4, int coffk4, int coffv4, float4* __restrict__ bank_k,
                                  float4* __restrict__ bank_v, int64_t slots_cap, int t,
                                  const int* __restrict__ dyn_t) {
  if (dyn_t) t = *dyn_t;  // bank slot from device memory (CUDA-graph replay)
  const int hw = h * w;
  const int per_pix = 32 + 128;  // float4s of key + value
  const int64_t total = static_cast<int64_t>(kobj) * hw * per_pix;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(i % per_pix);
    const int64_t t1 = i / per_pix;
    const int p = static_cast<int>(t1 % hw);
    const int obj = static_cast<int>(t1 / hw);
    const int y = p / w, x = p - y * w;
    const int64_t r = (static_cast<int64_t>(obj) * (h + 2) + y + 1) * (w + 2) + x + 1;
    const int64_t slot = static_cast<int64_t>(t) * hw + p;
    if (e < 32) {
      bank_k[(static_cast<int64_t>(obj) * slots_cap + slot) * 32 + e] = halo[r * cstride4 + coffk4 + e];
    } else {
      bank_v[(static_cast<int64_t>(obj) * slots_cap + slot) * 128 + (e - 32)] = halo[r * cstride4 + coffv4 + (e - 32)];
    }
  }
}

// src [obj][C][slots] -> dst [obj][slots_cap][C]; grid (slot tiles, C tiles, obj)
__global__ void bank_transpose_kernel(const float* __restrict__ src, int c, int64_t slots,
                                      float* __restrict__ dst, int64_t slots_cap) {
  __shared__ float tile[32][33];
  const int obj = blockIdx.z;
  const int64_t s0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int ch = c0 + j;
    const int64_t s = s0 + tx;
    tile[j][tx] = (s < slots && ch < c) ? src[(static_cast<int64_t>(obj) * c + ch) * slots + s] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int64_t s = s0 + j;
    const int ch = c0 + tx;
    if (s < slots && ch < c) dst[(static_cast<int64_t>(obj) * slots_cap + s) * c + ch] = tile[tx][j];
  }
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_exact(float x) { return 1.f / (1.f + expf(-x)); }

constexpr int kMaxObjects = 16;

// aggregate_wbg (aggregate.py:22-37) for one pixel: p[0..k) object probabilities in, softmax of
// logit(clamp([prod(1-p), p...])) out (o[0] = background).
__device__ __forceinline__ void aggregate_pixel(const float* p, int k, bool hard, float* o,
                                                bool const_bg = false) {
  float bg = 1.f;
  for (int j = 0; j < k; ++j) bg *= (1.f - p[j]);
  if (const_bg) bg = 0.5f;  // aggregate_sbg (aggregate.py:8)
  float mx = -INFINITY;
  for (int j = 0; j <= k; ++j) {
    float v = (j == 0) ? bg : p[j - 1];
    v = fminf(fmaxf(v, 1e-7f), 1.f - 1e-7f);
    float l = logf(v / (1.f - v));
    if (hard) l *= 1000.f;
    o[j] = l;
    mx = fmaxf(mx, l);
  }
  float s = 0.f;
  for (int j = 0; j <= k; ++j) {
    o[j] = expf(o[j] - mx);
    s += o[j];
  }
  for (int j = 0; j <= k; ++j) o[j] = o[j] / s;
}

__global__ void upsample4x_sigmoid_aggregate_kernel(const float* __restrict__ logits, int kobj,
                                                    int h4, int w4, int cstride, int coff,
                                                    float* __restrict__ raw_out,
                                                    float* __restrict__ prob_out) {
  const int H = 4 * h4, W = 4 * w4;
  const int64_t plane = static_cast<int64_t>(H) * W;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < plane;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int y = static_cast<int>(i / W), x = static_cast<int>(i - static_cast<int64_t>(y) * W);
    int y0, y1, x0, x1;
    float ly, lx;
    bilin(y, 0.25f, h4, y0, y1, ly);
    bilin(x, 0.25f, w4, x0, x1, lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    float p[kMaxObjects], o[kMaxObjects + 1];
    for (int j = 0; j < kobj; ++j) {
      const int64_t base = static_cast<int64_t>(j) * (h4 + 2);
      const float v00 = logits[((base + y0 + 1) * (w4 + 2) + x0 + 1) * cstride + coff];
      const float v01 = logits[((base + y0 + 1) * (w4 + 2) + x1 + 1) * cstride + coff];
      const float v10 = logits[((base + y1 + 1) * (w4 + 2) + x0 + 1) * cstride + coff];
      const float v11 = logits[((base + y1 + 1) * (w4 + 2) + x1 + 1) * cstride + coff];
      const float v = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
      p[j] = sigmoidf_exact(v);
      if (raw_out) raw_out[j * plane + i] = p[j];
    }
    if (prob_out) {
      aggregate_pixel(p, kobj, false, o);
      for (int j = 0; j <= kobj; ++j) prob_out[j * plane + i] = o[j];
    }
  }
}

__global__ void aggregate_wbg_kernel(const float* __restrict__ prob, int kobj, int64_t hw,
                                     int keep_bg, int hard, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < hw;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float p[kMaxObjects], o[kMaxObjects + 1];
    for (int j = 0; j < kobj; ++j) p[j] = prob[j * hw + i];
    aggregate_pixel(p, kobj, (hard & 1) != 0, o, (hard & 2) != 0);
    if (keep_bg) {
      for (int j = 0; j <= kobj; ++j) out[j * hw + i] = o[j];
    } else {
      for (int j = 1; j <= kobj; ++j) out[(j - 1) * hw + i] = o[j];
    }
  }
}

// prob [(K+1)][T][nh*nw] -> masks_padded [T][nh*nw] u8, masks_out [T][h][w] u8 (first max wins)
__global__ void argmax_unpad_kernel(const float* __restrict__ prob, int k1, int t, int nh, int nw,
                                    int pad_l, int pad_t, int h, int w,
                                    uint8_t* __restrict__ masks_padded,
                                    uint8_t* __restrict__ masks_out) {
  const int64_t plane = static_cast<int64_t>(nh) * nw;
  const int64_t total = static_cast<int64_t>(t) * plane;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float best = prob[i];
    int arg = 0;
    for (int j = 1; j < k1; ++j) {
      const float v = prob[static_cast<int64_t>(j) * total + i];
      if (v > best) { best = v; arg = j; }
    }
    masks_padded[i] = static_cast<uint8_t>(arg);
    if (masks_out) {
      const int ti = static_cast<int>(i / plane);
      const int rem = static_cast<int>(i - ti * plane);
      const int y = rem / nw - pad_t, x = rem % nw - pad_l;
      if (y >= 0 && y < h && x >= 0 && x < w)
        masks_out[(static_cast<int64_t>(ti) * h + y) * w + x] = static_cast<uint8_t>(arg);
    }
  }
}

__global__ void pad2d_kernel(const float* __restrict__ in, int planes, int h, int w, int pad_l,
                             int pad_t, int nh, int nw, float* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(planes) * nh * nw;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % nw) - pad_l;
    const int64_t t1 = i / nw;
    const int y = static_cast<int>(t1 % nh) - pad_t;
    const int64_t pl = t1 / nh;
    float v = 0.f;
    if (y >= 0 && y < h && x >= 0 && x < w) v = in[(pl * h + y) * w + x];
    out[i] = v;
  }
}

__global__ void fusion_gather_kernel(const float* __restrict__ im, const float* __restrict__ seg1,
                                     const float* __restrict__ seg2, const float* __restrict__ attn,
                                     float nc, float nr, int h, int w, float4* __restrict__ out) {
  const int64_t plane = static_cast<int64_t>(h) * w;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < plane;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int y = static_cast<int>(i / w), x = static_cast<int>(i - static_cast<int64_t>(y) * w);
    float4* o = out + (static_cast<int64_t>(y + 1) * (w + 2) + x + 1) * 8;
    o[0] = make_float4(im[i], im[plane + i], im[2 * plane + i], seg1[i]);
    o[1] = make_float4(seg2[i], attn[i], attn[plane + i], nc);
    o[2] = make_float4(nr, 0.f, 0.f, 0.f);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    o[3] = z; o[4] = z; o[5] = z; o[6] = z; o[7] = z;
  }
}

__global__ void halo_sigmoid_to_plane_kernel(const float* __restrict__ halo, int h, int w,
                                             int cstride, int coff, float* __restrict__ plane) {
  const int64_t total = static_cast<int64_t>(h) * w;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int y = static_cast<int>(i / w), x = static_cast<int>(i - static_cast<int64_t>(y) * w);
    plane[i] = sigmoidf_exact(halo[(static_cast<int64_t>(y + 1) * (w + 2) + x + 1) * cstride + coff]);
  }
}

__global__ void halo_copy_kernel(const float4* __restrict__ src, int src_n, int src_cs4, int src_co4,
                                 float4* __restrict__ dst, int dst_cs4, int dst_co4, int n, int h,
                                 int w, int c4, int relu) {
  const int64_t total = static_cast<int64_t>(n) * h * w * c4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % c4);
    int64_t t1 = i / c4;
    const int x = static_cast<int>(t1 % w);
    t1 /= w;
    const int y = static_cast<int>(t1 % h);
    const int img = static_cast<int>(t1 / h);
    const int simg = src_n == 1 ? 0 : img;
    const int64_t rs = (static_cast<int64_t>(simg) * (h + 2) + y + 1) * (w + 2) + x + 1;
    const int64_t rd = (static_cast<int64_t>(img) * (h + 2) + y + 1) * (w + 2) + x + 1;
    float4 v = src[rs * src_cs4 + src_co4 + ci];
    if (relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    dst[rd * dst_cs4 + dst_co4 + ci] = v;
  }
}

__global__ void halo_to_pixels_kernel(const float4* __restrict__ halo, int n, int h, int w, int cs4,
                                      int co4, int c4, float4* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(n) * h * w * c4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % c4);
    int64_t t1 = i / c4;
    const int x = static_cast<int>(t1 % w);
    t1 /= w;
    const int y = static_cast<int>(t1 % h);
    const int img = static_cast<int>(t1 / h);
    out[i] = halo[((static_cast<int64_t>(img) * (h + 2) + y + 1) * (w + 2) + x + 1) * cs4 + co4 + ci];
  }
}

__global__ void store_i32_kernel(int* dst, int n, int v0, int v1, int v2, int v3) {
  const int v[4] = {v0, v1, v2, v3};
  if (threadIdx.x < n) dst[threadIdx.x] = v[threadIdx.x];
}

inline unsigned capped_grid(int64_t work) {
  // grid-stride kernels: a few waves of 148 SMs x 8 resident 256-thread CTAs is plenty
  const int64_t cap = 148ll * 16;
  int64_t g = (work + kThreads - 1) / kThreads;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

}  // namespace
}  // namespace mivos

using namespace mivos;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define AL16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

extern "C" MIVOS_API int mivos_stem_gather(const float* frame, const float* masks, int k_objects,
                                           int h, int w, float* out, int kpad, mivos_stream_t s) {
  MIVOS_REQUIRE(frame && out, "stem_gather: null pointer");
  MIVOS_REQUIRE(h % 2 == 0 && w % 2 == 0 && h > 0 && w > 0, "stem_gather: h,w must be even");
  const int cin = masks ? 5 : 3;
  MIVOS_REQUIRE(k_objects >= 1, "stem_gather: bad object / frame count");
  MIVOS_REQUIRE(kpad >= 49 * cin && kpad % 32 == 0, "stem_gather: kpad %d too small for cin %d", kpad, cin);
  const int64_t total = static_cast<int64_t>(k_objects) * (h / 2 + 2) * (w / 2 + 2) * kpad;
  if (masks)
    stem_gather_kernel<5><<<capped_grid(total), kThreads, 0, ST(s)>>>(frame, masks, k_objects, h, w, out, kpad);
  else
    stem_gather_kernel<3><<<capped_grid(total), kThreads, 0, ST(s)>>>(frame, nullptr, k_objects, h, w, out, kpad);
  MIVOS_LAUNCHED();
  return MIVOS_OK;
}

extern "C" MIVOS_API int mivos_gather_s2(const float* in, int n, int h, int w, int c, int in_cstride,
                                         int ks, float* out, int out_cstride, mivos_stream_t s) {
  MIVOS_REQUIRE(in && out && AL16(in) && AL16(out), "gather_s2: null/unaligned pointer");
  MIVOS_REQUIRE((ks == 1 || ks == 3) && c % 4 == 0 && in_cstride % 4 == 0 && out_cstride % 4 == 0 &&
                    out_cstride >= ks * ks * c && h % 2 == 0 && w % 2 == 0,
                "gather_s2: bad shape (ks=%d c=%d)", ks, c);
  const int64_t total = static_cast<int64_t>(n) * (h / 2 + 2) * (w / 2 + 2) * ks * ks * (c / 4);
  gather_s2_kernel<<<capped_grid(total), kThreads, 0, ST(s)>>>(
      reinterpret_cast<const float4*>(in), n, h, w, c / 4, in_cstride / 4, ks,
      reinterpret_cast<float4*>(out), out_cstride / 4);
  MIVOS_LAUNCHED();
  return MIVOS_OK;
}

extern "C" MIVOS_API int mivos_maxpool3x3s2(const float* in, int n, int h, int w, int c, float* out,
                                            mivos_stream_t s) {
  MIVOS_REQUIRE(in && out && AL16(in) && AL16(out) && c % 4 == 0 && h % 2 == 0 && w % 2 == 0,
                "maxpool: bad arguments");
  const int64_t total = static_cast<int64_t>(n) * (h / 2) * (w / 2) * (c / 4);
  maxpool3x3s2_kernel<<<capped_grid(total), kThreads, 0, ST(s)>>>(
      reinterpret_cast<const float4*>(in), n, h, w, c / 4, reinterpret_cast<float4*>(out));
  MIVOS_LAUNCHED();
  return MIVOS_OK;
}

extern "C" MIVOS_API int mivos_upsample2x_add(float* x, const float* up, int n, int h, int w, int c,
                                              float* x_relu, const float* skip, mivos_stream_t s) {
  MIVOS_REQUIRE(x && up && AL16(x) && AL16(up) && (!x_relu || AL16(x_relu)) && c % 4 == 0 &&
                    h % 2 == 0 && w % 2 == 0,
                "upsample2x_add: bad arguments");
  const int64_t total = static_cast<int64_t>(n) * h * w * (c / 4);
  upsample2x_add_kernel<<<capped_grid(total), kThreads, 0, ST(s)>>>(
      reinterpret_cast<float4*>(x), reinterpret_cast<const float4*>(up), n, h, w, c / 4,
      reinterpret_cast<float4*>(x_relu), reinterpret_cast<const float4*>(skip));
  MIVOS_LAUNCHED();
  return MIVOS_OK;
}

extern "C" MIVOS_API int mivos_halo_to_nchw(const float* halo, int n, int h, int w, int cstride,
                                            int coff, int c, float* nchw, mivos_stream_t s) {
  MIVOS_REQUIRE(halo && nchw && n > 0 && c > 0 && coff + c <= cstride, "halo_to_nchw: bad arguments");
  dim3 grid(ceil_div(h * w, 32), ceil_div(c, 32), n);
  halo_to_nchw_kernel<<<grid, 256, 0, ST(s)>>>(halo, h, w, cstride, coff, c, nchw);
  MIVOS_LAUNCHED();
  return MIVOS_OK;
}

extern "C" MIVOS_API int mivos_nchw_to_halo(const float* nchw, int n, int h, int w, int c, float* halo,
                                            int cstride, int coff, int relu, mivos_stream_t s) {
  MIVOS_REQUIRE(halo && nchw && n > 0 && c > 0 && coff + c <= cstride, "nchw_to_halo: bad arguments");
  dim3 grid(ceil_div(h * w, 32), ceil_div(c, 32), n);
  nchw_to_halo_kernel<<<grid, 256, 0, ST(s)>>>(nchw, h, w, c, halo, cstride, coff, relu);
  MIVOS_LAUNCHED();
  return MIVOS_OK;
}

extern "C" MIVOS_API int mivos_bank_write(const float* halo, int k_objects, int h, int w, int cstride,
                                          int coff_k, int coff_v, float* bank_k, float* bank_v,
                                          int64_t slots_cap, int t, const int32_t* dyn_t, mivos_stream_t s) {
  MIVOS_REQUIRE(halo && bank_k && bank_v && AL16(halo) && AL16(bank_k) && AL16(bank_v), "bank_write: null/unaligned pointer");
  MIVOS_REQUIRE(cstride % 4 == 0 && coff_k % 4 == 0 && coff_v % 4 == 0 && t >= 0 &&
                    static_cast<int64_t>(t + 1) * h * w <= slots_cap,
                "bank_write: slot %d does not fit capacity %lld", t, (long long)slots_cap);
  const int64_t total = static_cast<int64_t>(k_objects) * h * w * 160;
  bank_write_kernel<<<capped_grid(total), kThreads, 0, ST(s)>>>(
      reinterpret_cast<const float4*>(halo), k_objects, h, w, cstride / 4, coff_k / 4, coff_v / 4,
      reinterpret_cast<float4*>(bank_k), reinterpret_cast<float4*>(bank_v), slots_cap, t, dyn_t);
  MIVOS_LAUNCHED();
  return MIVOS_OK;
}

extern "C" MIVOS_API int mivos_bank_from_nchw(const float* keys, const float* values, int k_objects,
                                              int t, int hw, float* bank_k, float* bank_v,
                                              int64_t slots_cap, mivos_stream_t s) {
  MIVOS_REQUIRE(keys && values && bank_k && bank_v, "bank_from_nchw: null pointer");
  const int64_t slots = static_cast<int64_t>(t) * hw;
  MIVOS_REQUIRE(slots <= slots_cap && slots > 0, "bank_from_nchw: %lld slots exceed capacity %lld",
                (long long)slots, (long long)slots_cap);
  dim3 gk(static_cast<unsigned>(ceil_div64(slots, 32)), 4, k_objects);
  bank_transpose_kernel<<<gk, 256, 0, ST(s)>>>(keys, 128, slots, bank_k, slots_cap);
  MIVOS_LAUNCHED();
  dim3 gv(static_cast<unsigned>(ceil_div64(slots, 32)), 16, k_objects);
  bank_transpose_kernel<<<gv, 256, 0, ST(s)>>>(values, 512, slots, bank_v, slots_cap);
  MIVOS_LAUNCHED();
  return MIVOS_OK;
}

extern "C" MIVOS_API int mivos_upsample4x_sigmoid_aggregate(const float* logits, int k_objects, int h4,
                                                            int w4, int cstride, int coff,
                                                            float* raw_out, float* prob_out,
                                                            mivos_stream_t s) {
  MIVOS_REQUIRE(logits && (raw_out || prob_out), "upsample4x: null pointer");
  MIVOS_REQUIRE(k_objects >= 1 && k_objects <= kMaxObjects, "upsample4x: %d objects (max %d)", k_objects, kMaxObjects);
  const int64_t plane = 16ll * h4 * w4;
  upsample4x_sigmoid_aggregate_kernel<<<capped_grid(plane), kThreads, 0, ST(s)>>>(
      logits, k_objects, h4, w4, cstride, coff, raw_out, prob_out);
  MIVOS_LAUNCHED();
  return MIVOS_OK;
}

extern "C" MIVOS_API int mivos_aggregate_wbg(const float* prob, int k_objects, int64_t hw, int keep_bg,
                                             int hard, float* out, mivos_stream_t s) {
  MIVOS_REQUIRE(prob && out, "aggregate_wbg: null pointer");
  MIVOS_REQUIRE(k_objects >= 1 && k_objects <= kMaxObjects, "aggregate_wbg: %d objects (max %d)", k_objects, kMaxObjects);
  aggregate_wbg_kernel<<<capped_grid(hw), kThreads, 0, ST(s)>>>(prob, k_objects, hw, keep_bg, hard, out);
  MIVOS_LAUNCHED();
  return MIVOS_OK;
}

extern "C" MIVOS_API int mivos_argmax_unpad(const float* prob, int k_plus_1, int t, int nh, int nw,
                                            int pad_l, int pad_t, int h, int w, uint8_t* masks_padded,
                                            uint8_t* masks_out, mivos_stream_t s) {
  MIVOS_REQUIRE(prob && masks_padded && k_plus_1 >= 1 && k_plus_1 <= 255, "argmax_unpad: bad arguments");
  const int64_t total = static_cast<int64_t>(t) * nh * nw;
  argmax_unpad_kernel<<<capped_grid(total), kThreads, 0, ST(s)>>>(prob, k_plus_1, t, nh, nw, pad_l, pad_t,
                                                                  h, w, masks_padded, masks_out);
  MIVOS_LAUNCHED();
  return MIVOS_OK;
}

extern "C" MIVOS_API int mivos_pad2d(const float* in, int planes, int h, int w, int pad_l, int pad_r,
                                     int pad_t, int pad_b, float* out, mivos_stream_t s) {
  MIVOS_REQUIRE(in && out && pad_l >= 0 && pad_r >= 0 && pad_t >= 0 && pad_b >= 0, "pad2d: bad arguments");
  const int nh = h + pad_t + pad_b, nw = w + pad_l + pad_r;
  const int64_t total = static_cast<int64_t>(planes) * nh * nw;
  pad2d_kernel<<<capped_grid(total), kThreads, 0, ST(s)>>>(in, planes, h, w, pad_l, pad_t, nh, nw, out);
  MIVOS_LAUNCHED();
  return MIVOS_OK;
}

extern "C" MIVOS_API int mivos_fusion_gather(const float* im, const float* seg1, const float* seg2,
                                             const float* attn, float nc, float nr, int h, int w,
                                             float* out_halo, mivos_stream_t s) {
  MIVOS_REQUIRE(im && seg1 && seg2 && attn && out_halo && AL16(out_halo), "fusion_gather: null/unaligned pointer");
  const int64_t plane = static_cast<int64_t>(h) * w;
  fusion_gather_kernel<<<capped_grid(plane), kThreads, 0, ST(s)>>>(im, seg1, seg2, attn, nc, nr, h, w,
                                                                   reinterpret_cast<float4*>(out_halo));
  MIVOS_LAUNCHED();
  return MIVOS_OK;
}

extern "C" MIVOS_API int mivos_halo_sigmoid_to_plane(const float* halo, int h, int w, int cstride,
                                                     int coff, float* plane, mivos_stream_t s) {
  MIVOS_REQUIRE(halo && plane, "halo_sigmoid_to_plane: null pointer");
  halo_sigmoid_to_plane_kernel<<<capped_grid(static_cast<int64_t>(h) * w), kThreads, 0, ST(s)>>>(
      halo, h, w, cstride, coff, plane);
  MIVOS_LAUNCHED();
  return MIVOS_OK;
}

extern "C" MIVOS_API int mivos_halo_copy(const float* src, int src_n, int src_cstride, int src_coff,
                                         float* dst, int dst_cstride, int dst_coff, int n, int h, int w,
                                         int c, int relu, mivos_stream_t s) {
  MIVOS_REQUIRE(src && dst && AL16(src) && AL16(dst), "halo_copy: null/unaligned pointer");
  MIVOS_REQUIRE(c % 4 == 0 && src_cstride % 4 == 0 && dst_cstride % 4 == 0 && src_coff % 4 == 0 &&
                    dst_coff % 4 == 0 && src_coff + c <= src_cstride && dst_coff + c <= dst_cstride &&
                    (src_n == 1 || src_n == n),
                "halo_copy: bad channel window");
  const int64_t total = static_cast<int64_t>(n) * h * w * (c / 4);
  halo_copy_kernel<<<capped_grid(total), kThreads, 0, ST(s)>>>(
      reinterpret_cast<const float4*>(src), src_n, src_cstride / 4, src_coff / 4,
      reinterpret_cast<float4*>(dst), dst_cstride / 4, dst_coff / 4, n, h, w, c / 4, relu);
  MIVOS_LAUNCHED();
  return MIVOS_OK;
}

extern "C" MIVOS_API int mivos_halo_to_pixels(const float* halo, int n, int h, int w, int cstride, int coff,
                                              int c, float* out, mivos_stream_t s) {
  MIVOS_REQUIRE(halo && out && AL16(halo) && AL16(out) && c % 4 == 0 && cstride % 4 == 0 && coff % 4 == 0 &&
                    coff + c <= cstride,
                "halo_to_pixels: bad arguments");
  const int64_t total = static_cast<int64_t>(n) * h * w * (c / 4);
  halo_to_pixels_kernel<<<capped_grid(total), kThreads, 0, ST(s)>>>(
      reinterpret_cast<const float4*>(halo), n, h, w, cstride / 4, coff / 4, c / 4, reinterpret_cast<float4*>(out));
  MIVOS_LAUNCHED();
  return MIVOS_OK;
}

extern "C" MIVOS_API int mivos_store_i32(int32_t* dst, int n, int v0, int v1, int v2, int v3, mivos_stream_t s) {
  MIVOS_REQUIRE(dst && n >= 1 && n <= 4, "store_i32: bad arguments");
  store_i32_kernel<<<1, 32, 0, ST(s)>>>(dst, n, v0, v1, v2, v3);
  MIVOS_LAUNCHED();
  return MIVOS_OK;
}
