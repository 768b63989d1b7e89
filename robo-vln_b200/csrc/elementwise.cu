// Bandwidth-bound kernels around the tensor-core contractions: stems, pooling, GroupNorm,
// LayerNorm, embeddings, small heads.  All activations are NHWC / row-major h16 with fp32
// statistics; every kernel moves 16-byte vectors along the contiguous (channel) dimension.
#include "common.cuh"
#include "rvb.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace rvb {

namespace {

RVB_DEVICE void load8(const h16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 t;
  t = unpack_h2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_h2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_h2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_h2(u.w); f[6] = t.x; f[7] = t.y;
}
RVB_DEVICE void store8(h16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_h2(f[0], f[1]);
  u.y = pack_h2(f[2], f[3]);
  u.z = pack_h2(f[4], f[5]);
  u.w = pack_h2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ---------------------------------------------------------------------------------------
// RGB stem: [NB,H,W,3] fp32 0..255 -> im2col rows [NB*Ho*Wo, Kpitch] h16 for the 7x7 s2 p3
// conv (k = r*21 + s*3 + c), with the reference's /255 (resnet_encoders.py:213) folded in.
// ---------------------------------------------------------------------------------------
constexpr int STEM_PIX = 32;
__global__ void __launch_bounds__(256) rgb_stem_im2col_kernel(const float* __restrict__ rgb, h16* __restrict__ out,
                                                              int NB, int H, int W, int Ho, int Wo, int Kpitch) {
  RVB_PDL_PROLOGUE();
  extern __shared__ __align__(16) uint8_t sm_raw[];
  h16* tile = reinterpret_cast<h16*>(sm_raw);  // [STEM_PIX][Kpitch]
  const long long M = static_cast<long long>(NB) * Ho * Wo;
  const long long m0 = static_cast<long long>(blockIdx.x) * STEM_PIX;
  const int total = STEM_PIX * Kpitch;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int px = e / Kpitch;
    const int k = e - px * Kpitch;
    float v = 0.0f;
    const long long m = m0 + px;
    if (k < 147 && m < M) {
      const int wo = static_cast<int>(m % Wo);
      const int ho = static_cast<int>((m / Wo) % Ho);
      const int img = static_cast<int>(m / (static_cast<long long>(Wo) * Ho));
      const int r = k / 21;
      const int rem = k - r * 21;
      const int s = rem / 3;
      const int c = rem - s * 3;
      const int h = 2 * ho + r - 3;
      const int w = 2 * wo + s - 3;
      if (h >= 0 && h < H && w >= 0 && w < W)
        v = __ldg(rgb + ((static_cast<long long>(img) * H + h) * W + w) * 3 + c) / 255.0f;
    }
    tile[e] = to_h16(v);
  }
  __syncthreads();
  const long long rows = std::min<long long>(STEM_PIX, M - m0);
  const int nvec = static_cast<int>(rows * Kpitch / 8);
  uint4* dst = reinterpret_cast<uint4*>(out + m0 * Kpitch);
  const uint4* src = reinterpret_cast<const uint4*>(tile);
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------
// RGB stem pre-pass: [NB,H,W,3] fp32 0..255 -> zero-padded [NB, H+6, Wp, 8] 16-bit image
// (3 px of padding on every side, channels 3..7 zero, /255 folded in, resnet_encoders.py:213).
// The 7x7 stride-2 conv then runs on the tensor cores in "window" mode (gemm_tc.cu) with no
// im2col buffer.  One thread per padded pixel: 12 B read, 16 B written.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rgb_pad_convert_kernel(const float* __restrict__ rgb, h16* __restrict__ out,
                                                              int NB, int H, int W, int Hp, int Wp) {
  RVB_PDL_PROLOGUE();
  const long long total = static_cast<long long>(NB) * Hp * Wp;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xp = static_cast<int>(i % Wp);
    const int yp = static_cast<int>((i / Wp) % Hp);
    const int img = static_cast<int>(i / (static_cast<long long>(Wp) * Hp));
    const int x = xp - 3, y = yp - 3;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (x >= 0 && x < W && y >= 0 && y < H) {
      const float* src = rgb + ((static_cast<long long>(img) * H + y) * W + x) * 3;
      q.x = pack_h2(__ldg(src) / 255.0f, __ldg(src + 1) / 255.0f);   // true division, as the reference
      q.y = pack_h2(__ldg(src + 2) / 255.0f, 0.0f);
    }
    *reinterpret_cast<uint4*>(out + i * 8) = q;
  }
}

// Packed variant for the two-filter-rows-per-K-block stem: row-pair interleaved [NB, (H+6)/2, Wp, 2, 4]
// (for every pixel column: row 2t then row 2t+1, 4 channels each with channel 3 zero).  One thread per
// (row pair, pixel): 24 B read, 16 B written.
template <typename T>   // float (0..255, the reference's batch_obs output) or uint8_t (the sensor's own format)
__global__ void __launch_bounds__(256) rgb_pad_convert4_kernel(const T* __restrict__ rgb, h16* __restrict__ out,
                                                               int NB, int H, int W, int Hh, int Wp) {
  RVB_PDL_PROLOGUE();
  const long long total = static_cast<long long>(NB) * Hh * Wp;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xp = static_cast<int>(i % Wp);
    const int tp = static_cast<int>((i / Wp) % Hh);
    const int img = static_cast<int>(i / (static_cast<long long>(Wp) * Hh));
    const int x = xp - 3;
    uint32_t q[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int r2 = 0; r2 < 2; ++r2) {
      const int y = tp * 2 + r2 - 3;
      if (x >= 0 && x < W && y >= 0 && y < H) {
        const T* src = rgb + ((static_cast<long long>(img) * H + y) * W + x) * 3;
        q[2 * r2] = pack_h2(static_cast<float>(__ldg(src)) / 255.0f, static_cast<float>(__ldg(src + 1)) / 255.0f);   // true division, as the reference
        q[2 * r2 + 1] = pack_h2(static_cast<float>(__ldg(src + 2)) / 255.0f, 0.0f);
      }
    }
    *reinterpret_cast<uint4*>(out + i * 8) = make_uint4(q[0], q[1], q[2], q[3]);
  }
}

// ---------------------------------------------------------------------------------------
// 3x3 stride-2 pad-1 max pooling, NHWC h16
// ---------------------------------------------------------------------------------------
RVB_DEVICE uint32_t hmax2_u32(uint32_t a, uint32_t b) {
#if RVB_BF16
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
#else
  __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
#endif
  return *reinterpret_cast<uint32_t*>(&r);
}

// One thread per (output pixel, 8-channel vector): the nine 16-byte taps are loaded up front (taps that fall
// outside the image are redirected to the centre tap, which is always inside, so no -inf and no branches) and
// reduced with packed 16-bit max.
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const h16* __restrict__ in, h16* __restrict__ out, int NB, int H,
                                                           int W, int C, int Ho, int Wo) {
  RVB_PDL_PROLOGUE();
  const int cv = C / 8;
  const long long total = static_cast<long long>(NB) * Ho * Wo * cv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % cv);
    const long long pix = i / cv;
    const int wo = static_cast<int>(pix % Wo);
    const int ho = static_cast<int>((pix / Wo) % Ho);
    const int img = static_cast<int>(pix / (static_cast<long long>(Wo) * Ho));
    const h16* base = in + static_cast<long long>(img) * H * W * C + v * 8;
    uint4 t[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      int h = 2 * ho + r - 1;
      h = (h < 0 || h >= H) ? 2 * ho : h;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int w = 2 * wo + s - 1;
        w = (w < 0 || w >= W) ? 2 * wo : w;
        t[r * 3 + s] = __ldg(reinterpret_cast<const uint4*>(base + (static_cast<long long>(h) * W + w) * C));
      }
    }
    uint4 m = t[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) {
      m.x = hmax2_u32(m.x, t[k].x); m.y = hmax2_u32(m.y, t[k].y);
      m.z = hmax2_u32(m.z, t[k].z); m.w = hmax2_u32(m.w, t[k].w);
    }
    *reinterpret_cast<uint4*>(out + pix * C + v * 8) = m;
  }
}

// ---------------------------------------------------------------------------------------
// Depth stem: [NB,H,W,1] fp32 -> avg_pool2d(2) -> conv7x7 s2 p3 (1->32, no bias) -> raw h16
// [NB,Ho,Wo,32]  (resnet_policy.py:184-186, resnet.py:185-196).  fp32 math on CUDA cores:
// 12.8 MFLOP per image.
// ---------------------------------------------------------------------------------------
constexpr int DS_ROWS = 8;   // output rows per CTA: the 49 x 32 weights and the overlapping input rows are staged once per 8 rows
__global__ void __launch_bounds__(256) depth_stem_kernel(const float* __restrict__ depth, const float* __restrict__ w,
                                                         h16* __restrict__ out, int H, int W, int Hp, int Wp, int Ho,
                                                         int Wo) {
  RVB_PDL_PROLOGUE();
  extern __shared__ __align__(16) uint8_t sm_raw[];
  float* ws = reinterpret_cast<float*>(sm_raw);  // [49][32]
  float* rows = ws + 49 * 32;                    // [2*DS_ROWS + 5][Wp + 6] pooled input rows (zero padded)
  const int img = blockIdx.y;
  const int ho0 = blockIdx.x * DS_ROWS;
  const int RW = Wp + 6;
  constexpr int NR = 2 * DS_ROWS + 5;
  for (int i = threadIdx.x; i < 49 * 32; i += blockDim.x) {
    const int k = i / 32, ch = i % 32;
    ws[i] = w[ch * 49 + k];
  }
  const float* base = depth + static_cast<long long>(img) * H * W;
  for (int i = threadIdx.x; i < NR * RW; i += blockDim.x) {
    const int r = i / RW;
    const int x = i - r * RW - 3;
    const int hp = 2 * ho0 + r - 3;
    float v = 0.0f;
    if (hp >= 0 && hp < Hp && x >= 0 && x < Wp) {
      const float2 a = *reinterpret_cast<const float2*>(base + static_cast<long long>(2 * hp) * W + 2 * x);
      const float2 b = *reinterpret_cast<const float2*>(base + static_cast<long long>(2 * hp + 1) * W + 2 * x);
      v = 0.25f * (a.x + a.y + b.x + b.y);
    }
    rows[i] = v;
  }
  __syncthreads();
  // thread = (channel, group of 4 consecutive output columns) x output row: per filter row the 13 inputs the four
  // sliding windows share are read once and each weight feeds 4 FMAs (0.7 shared loads per FMA instead of 2)
  const int ch = threadIdx.x & 31;
  const int wgroups = (Wo + 3) / 4;
  for (int item = threadIdx.x >> 5; item < DS_ROWS * wgroups; item += blockDim.x >> 5) {
    const int ro = item / wgroups, wo0 = (item - ro * wgroups) * 4;
    const int ho = ho0 + ro;
    if (ho >= Ho) continue;
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      const float* rr = rows + (2 * ro + r) * RW + 2 * wo0;
      float x[13];
#pragma unroll
      for (int i = 0; i < 13; ++i) x[i] = (2 * wo0 + i < RW) ? rr[i] : 0.0f;
#pragma unroll
      for (int s = 0; s < 7; ++s) {
        const float wv = ws[(r * 7 + s) * 32 + ch];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaf(x[2 * j + s], wv, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (wo0 + j < Wo) out[((static_cast<long long>(img) * Ho + ho) * Wo + wo0 + j) * 32 + ch] = to_h16(acc[j]);
  }
}

// ---------------------------------------------------------------------------------------
// GroupNorm statistics: stats[img][g] = (sum, sumsq) over HW x (C/G) elements.
// One CTA per sample and a fixed-order reduction (registers -> smem tree over pixel rows ->
// per-group serial sum): bit-reproducible from run to run, no atomics, no zero-init needed.
// Thread t owns the 8-channel slot t % (C/8) and pixel rows t / (C/8), + k * 256/(C/8).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_stats_kernel(const h16* __restrict__ x, float* __restrict__ stats, int HW,
                                                       int C, int G) {
  RVB_PDL_PROLOGUE();
  __shared__ float ps[256][8];
  __shared__ float pq[256][8];
  const int img = blockIdx.x;
  const int cv = C / 8;
  const int slot = threadIdx.x % cv;
  const int prow = threadIdx.x / cv;
  const int rows_per_iter = blockDim.x / cv;   // power of two
  const int cpg = C / G;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.0f;
  const h16* base = x + static_cast<long long>(img) * HW * C + slot * 8;
  for (int p = prow; p < HW; p += rows_per_iter) {
    float f[8];
    load8(base + static_cast<long long>(p) * C, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += f[j];
      q[j] = fmaf(f[j], f[j], q[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ps[threadIdx.x][j] = s[j];
    pq[threadIdx.x][j] = q[j];
  }
  __syncthreads();
  for (int stride = rows_per_iter >> 1; stride > 0; stride >>= 1) {
    if (prow < stride) {
      const int other = threadIdx.x + stride * cv;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ps[threadIdx.x][j] += ps[other][j];
        pq[threadIdx.x][j] += pq[other][j];
      }
    }
    __syncthreads();
  }
  // rows 0..cv-1 of ps/pq now hold the per-channel totals (channel c -> [c / 8][c % 8])
  if (threadIdx.x < G) {
    float ss = 0.0f, qq = 0.0f;
    const int c0 = threadIdx.x * cpg;
    for (int c = c0; c < c0 + cpg; ++c) {
      ss += ps[c >> 3][c & 7];
      qq += pq[c >> 3][c & 7];
    }
    stats[(static_cast<long long>(img) * G + threadIdx.x) * 2 + 0] = ss;
    stats[(static_cast<long long>(img) * G + threadIdx.x) * 2 + 1] = qq;
  }
}

struct GnApplyDev {
  const h16* x; const float* stats; const float* gamma; const float* beta;
  int NB, HW, C, G, relu, res_mode;
  const h16* res; const float* res_stats; const float* res_gamma; const float* res_beta;
  h16* out; long long out_pitch;
};

RVB_DEVICE void gn_scale_shift(const float* stats, int img, int G, int g, float cnt, float gamma, float beta,
                               float& sc, float& sh) {
  const float sum = stats[(static_cast<long long>(img) * G + g) * 2 + 0];
  const float sq = stats[(static_cast<long long>(img) * G + g) * 2 + 1];
  const float mean = sum / cnt;
  const float var = fmaxf(sq / cnt - mean * mean, 0.0f);
  const float rstd = rsqrtf(var + 1e-5f);
  sc = rstd * gamma;
  sh = beta - mean * sc;
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const GnApplyDev a) {
  RVB_PDL_PROLOGUE();
  const int cv = a.C / 8;
  const int cpg = a.C / a.G;
  const float cnt = static_cast<float>(a.HW) * static_cast<float>(cpg);
  const long long total = static_cast<long long>(a.NB) * a.HW * cv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % cv);
    const long long pix = i / cv;
    const int img = static_cast<int>(pix / a.HW);
    const int c0 = v * 8;
    float f[8];
    load8(a.x + pix * a.C + c0, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float sc, sh;
      gn_scale_shift(a.stats, img, a.G, (c0 + j) / cpg, cnt, __ldg(a.gamma + c0 + j), __ldg(a.beta + c0 + j), sc, sh);
      f[j] = fmaf(f[j], sc, sh);
    }
    if (a.res_mode == 1) {
      float r[8];
      load8(a.res + pix * a.C + c0, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += r[j];
    } else if (a.res_mode == 2) {
      float r[8];
      load8(a.res + pix * a.C + c0, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float sc, sh;
        gn_scale_shift(a.res_stats, img, a.G, (c0 + j) / cpg, cnt, __ldg(a.res_gamma + c0 + j),
                       __ldg(a.res_beta + c0 + j), sc, sh);
        f[j] += fmaf(r[j], sc, sh);
      }
    }
    if (a.relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.0f);
    }
    store8(a.out + pix * a.out_pitch + c0, f);
  }
}

// ---------------------------------------------------------------------------------------
// Fused GroupNorm: statistics + normalise + affine (+ residual | + second normalised branch)
// + ReLU in ONE launch, one CTA per sample.  Pass 1 reduces (sum, sumsq) per channel in a fixed
// order (bit-reproducible), turns them into per-channel scale/shift tables in shared memory;
// pass 2 re-reads the sample (L2-resident: <= 256 KB) and writes the result.  Halves the number
// of launches of the latency-bound depth trunk compared with separate stats / apply kernels.
// ---------------------------------------------------------------------------------------
constexpr int GNF_THREADS = 512;
__global__ void __launch_bounds__(GNF_THREADS) gn_fused_kernel(const GnApplyDev a) {
  RVB_PDL_PROLOGUE();
  extern __shared__ __align__(16) uint8_t gnf_dyn[];
  float (*ps)[8] = reinterpret_cast<float (*)[8]>(gnf_dyn);
  float (*pq)[8] = ps + GNF_THREADS;
  float* sc_x = reinterpret_cast<float*>(pq + GNF_THREADS);   // [C] scale, [C] shift for x; then for res (mode 2)
  float* sh_x = sc_x + a.C;
  float* sc_r = sh_x + a.C;
  float* sh_r = sc_r + a.C;
  const int img = blockIdx.x;
  const int cv = a.C / 8;
  const int slot = threadIdx.x % cv;
  const int prow = threadIdx.x / cv;
  const int rows_per_iter = blockDim.x / cv;   // power of two
  const int cpg = a.C / a.G;
  const float cnt = static_cast<float>(a.HW) * static_cast<float>(cpg);
  const int n_src = (a.res_mode == 2) ? 2 : 1;
  for (int src = 0; src < n_src; ++src) {
    const h16* base = (src == 0 ? a.x : a.res) + static_cast<long long>(img) * a.HW * a.C + slot * 8;
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.0f;
    // four 16-byte loads in flight per thread: one CTA streams up to 256 KB here, and with a single
    // outstanding load per thread the pass ran at a few GB/s per SM
    int p = prow;
    for (; p + 3 * rows_per_iter < a.HW; p += 4 * rows_per_iter) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(p + u * rows_per_iter) * a.C));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = unpack_h2(w[j]);
          s[2 * j] += t.x; q[2 * j] = fmaf(t.x, t.x, q[2 * j]);
          s[2 * j + 1] += t.y; q[2 * j + 1] = fmaf(t.y, t.y, q[2 * j + 1]);
        }
      }
    }
    for (; p < a.HW; p += rows_per_iter) {
      float f[8];
      load8(base + static_cast<long long>(p) * a.C, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        q[j] = fmaf(f[j], f[j], q[j]);
      }
    }
    __syncthreads();   // previous use of ps/pq finished
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ps[threadIdx.x][j] = s[j];
      pq[threadIdx.x][j] = q[j];
    }
    __syncthreads();
    for (int stride = rows_per_iter >> 1; stride > 0; stride >>= 1) {
      if (prow < stride) {
        const int other = threadIdx.x + stride * cv;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          ps[threadIdx.x][j] += ps[other][j];
          pq[threadIdx.x][j] += pq[other][j];
        }
      }
      __syncthreads();
    }
    // per-channel scale / shift (thread c: group of channel c, fixed-order sum over its cpg channels)
    const float* gam = (src == 0) ? a.gamma : a.res_gamma;
    const float* bet = (src == 0) ? a.beta : a.res_beta;
    float* sc = (src == 0) ? sc_x : sc_r;
    float* sh = (src == 0) ? sh_x : sh_r;
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
      const int c0 = (c / cpg) * cpg;
      float ss = 0.0f, qq = 0.0f;
      for (int k = c0; k < c0 + cpg; ++k) {
        ss += ps[k >> 3][k & 7];
        qq += pq[k >> 3][k & 7];
      }
      const float mean = ss / cnt;
      const float var = fmaxf(qq / cnt - mean * mean, 0.0f);
      const float scale = rsqrtf(var + 1e-5f) * __ldg(gam + c);
      sc[c] = scale;
      sh[c] = __ldg(bet + c) - mean * scale;
    }
    __syncthreads();
  }
  // pass 2
  const long long pix0 = static_cast<long long>(img) * a.HW;
  // (slot, prow) mapping as in pass 1: this thread's 8 channels are fixed, so its scale/shift live in
  // registers; two pixels per iteration keep 2-4 loads in flight
  const int c0 = slot * 8;
  float scx[8], shx[8], scr[8], shr[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    scx[j] = sc_x[c0 + j]; shx[j] = sh_x[c0 + j];
    scr[j] = (a.res_mode == 2) ? sc_r[c0 + j] : 1.0f;
    shr[j] = (a.res_mode == 2) ? sh_r[c0 + j] : 0.0f;
  }
  for (int p0 = prow; p0 < a.HW; p0 += 2 * rows_per_iter) {
    float f[2][8], r[2][8];
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int p = p0 + u * rows_per_iter;
      ok[u] = p < a.HW;
      if (ok[u]) {
        load8(a.x + (pix0 + p) * a.C + c0, f[u]);
        if (a.res_mode != 0) load8(a.res + (pix0 + p) * a.C + c0, r[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!ok[u]) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = fmaf(f[u][j], scx[j], shx[j]);
        if (a.res_mode != 0) v += fmaf(r[u][j], scr[j], shr[j]);
        f[u][j] = a.relu ? fmaxf(v, 0.0f) : v;
      }
      store8(a.out + (pix0 + p0 + u * rows_per_iter) * a.out_pitch + c0, f[u]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// RGB head pooling: layer4 output [NB,H,W,C] ->
//   tokens[img][cell][0..C)  = adaptive_avg_pool2d(4,4)   (resnet_encoders.py:162-166)
//   cellmean[img][0..C)      = mean over the 16 cells     (rgb_linear's AdaptiveAvgPool1d(1))
//   gmean[img][0..C)         = global mean                (torchvision avgpool, lo path)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rgb_pool_kernel(const h16* __restrict__ feat, int H, int W, int C,
                                                       h16* __restrict__ tokens, long long tok_pitch,
                                                       h16* __restrict__ cellmean, long long cm_pitch,
                                                       h16* __restrict__ gmean) {
  RVB_PDL_PROLOGUE();
  const int img = blockIdx.x;
  const h16* base = feat + static_cast<long long>(img) * H * W * C;
  for (int v = threadIdx.x; v < C / 8; v += blockDim.x) {
    float cm[8], gm[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cm[j] = gm[j] = 0.0f;
    for (int ch = 0; ch < 4; ++ch) {
      const int h0 = (ch * H) / 4, h1 = ((ch + 1) * H + 3) / 4;
      for (int cw = 0; cw < 4; ++cw) {
        const int w0 = (cw * W) / 4, w1 = ((cw + 1) * W + 3) / 4;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
        for (int h = h0; h < h1; ++h)
          for (int w = w0; w < w1; ++w) {
            float f[8];
            load8(base + (static_cast<long long>(h) * W + w) * C + v * 8, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
          }
        const float inv = 1.0f / static_cast<float>((h1 - h0) * (w1 - w0));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[j] *= inv;
          cm[j] += acc[j];
        }
        store8(tokens + (static_cast<long long>(img) * 16 + ch * 4 + cw) * tok_pitch + v * 8, acc);
      }
    }
    for (int p = 0; p < H * W; ++p) {
      float f[8];
      load8(base + static_cast<long long>(p) * C + v * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) gm[j] += f[j];
    }
    const float invp = 1.0f / static_cast<float>(H * W);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      cm[j] *= (1.0f / 16.0f);
      gm[j] *= invp;
    }
    store8(cellmean + static_cast<long long>(img) * cm_pitch + v * 8, cm);
    store8(gmean + static_cast<long long>(img) * C + v * 8, gm);
  }
}

// Spatial-embedding channels: the reference views the [16,64] table as [64,4,4]
// (resnet_encoders.py:91-102): channel c of cell k reads flat[c*16 + k].
__global__ void fill_spatial_embedding_kernel(const float* __restrict__ flat, h16* __restrict__ tokens, int NB,
                                              long long tok_pitch, int col0, h16* __restrict__ cellmean,
                                              long long cm_pitch) {
  RVB_PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NB * 64) return;
  const int img = i / 64, c = i % 64;
  float m = 0.0f;
  for (int k = 0; k < 16; ++k) {
    const float v = flat[c * 16 + k];
    m += v;
    tokens[(static_cast<long long>(img) * 16 + k) * tok_pitch + col0 + c] = to_h16(v);
  }
  if (cellmean != nullptr) cellmean[static_cast<long long>(img) * cm_pitch + col0 + c] = to_h16(m / 16.0f);
}

// ---------------------------------------------------------------------------------------
// BERT embeddings + LayerNorm(eps 1e-12): one warp per token, D = 768.
// ---------------------------------------------------------------------------------------
template <int NV>  // float4 per lane: D = NV * 128
__global__ void __launch_bounds__(256) bert_embed_ln_kernel(const long long* __restrict__ ids_i64,
                                                            const float* __restrict__ ids_f32, int id_rows, int R,
                                                            int L, const float* __restrict__ word,
                                                            const float* __restrict__ pos,
                                                            const float* __restrict__ type0,
                                                            const float* __restrict__ g, const float* __restrict__ b,
                                                            h16* __restrict__ out) {
  RVB_PDL_PROLOGUE();
  constexpr int D = NV * 128;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= R * L) return;
  const int row = warp / L, l = warp - row * L;
  const int src_row = (id_rows == 1) ? 0 : row;
  long long id = ids_i64 != nullptr ? ids_i64[static_cast<long long>(src_row) * L + l]
                                    : static_cast<long long>(ids_f32[static_cast<long long>(src_row) * L + l]);
  float4 v[NV];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 a = __ldg(reinterpret_cast<const float4*>(word + id * D + c));
    const float4 p4 = __ldg(reinterpret_cast<const float4*>(pos + static_cast<long long>(l) * D + c));
    const float4 t = __ldg(reinterpret_cast<const float4*>(type0 + c));
    v[i] = make_float4(a.x + t.x + p4.x, a.y + t.y + p4.y, a.z + t.z + p4.z, a.w + t.w + p4.w);
    sum += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(sum) / D;
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
    sq += dx * dx + dy * dy + dz * dz + dw * dw;
  }
  const float rstd = rsqrtf(warp_sum(sq) / D + 1e-12f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
    uint2 o;
    o.x = pack_h2((v[i].x - mean) * rstd * gg.x + bb.x, (v[i].y - mean) * rstd * gg.y + bb.y);
    o.y = pack_h2((v[i].z - mean) * rstd * gg.z + bb.z, (v[i].w - mean) * rstd * gg.w + bb.w);
    *reinterpret_cast<uint2*>(out + static_cast<long long>(warp) * D + c) = o;
  }
}

// LayerNorm over fp32 rows (the producing GEMM already added bias and residual), optional
// additive table (sinusoid PE, transformer.py:271-274) AFTER the norm; h16 out.
template <int NV>
__global__ void __launch_bounds__(128) layernorm_rows_kernel(const float* __restrict__ x, int M,
                                                             const float* __restrict__ g,
                                                             const float* __restrict__ b, float eps,
                                                             const float* __restrict__ pe, int pe_rows,
                                                             h16* __restrict__ out) {
  RVB_PDL_PROLOGUE();
  constexpr int D = NV * 128;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float4 v[NV];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = *reinterpret_cast<const float4*>(x + static_cast<long long>(row) * D + (i * 32 + lane) * 4);
    sum += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(sum) / D;
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
    sq += dx * dx + dy * dy + dz * dz + dw * dw;
  }
  const float rstd = rsqrtf(warp_sum(sq) / D + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
    float4 y = make_float4((v[i].x - mean) * rstd * gg.x + bb.x, (v[i].y - mean) * rstd * gg.y + bb.y,
                           (v[i].z - mean) * rstd * gg.z + bb.z, (v[i].w - mean) * rstd * gg.w + bb.w);
    if (pe != nullptr) {
      const float4 pp = __ldg(reinterpret_cast<const float4*>(pe + static_cast<long long>(row % pe_rows) * D + c));
      y.x += pp.x; y.y += pp.y; y.z += pp.z; y.w += pp.w;
    }
    uint2 o;
    o.x = pack_h2(y.x, y.y);
    o.y = pack_h2(y.z, y.w);
    *reinterpret_cast<uint2*>(out + static_cast<long long>(row) * D + c) = o;
  }
}

// mean over the L tokens of each (modality, batch row) group: x [n_mod*B*L, D] ->
// out[b*out_pitch + mod*mod_stride + d]   (cross_pooler, seq2seq_highlevel_cma.py:114-115,209-210)
__global__ void __launch_bounds__(256) token_mean_kernel(const h16* __restrict__ x, int B, int L, int D,
                                                         h16* __restrict__ out, long long out_pitch,
                                                         long long mod_stride) {
  RVB_PDL_PROLOGUE();
  // one CTA per (modality, batch row); D == 256: lane v owns channels [8v, 8v+8), warp w sums
  // rows w, w+8, ...; partials are folded in warp order (deterministic)
  __shared__ float part[8][256];
  const int g = blockIdx.x;
  const int mod = g / B, b = g % B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
  const h16* p = x + static_cast<long long>(g) * L * D + lane * 8;
  for (int l = warp; l < L; l += 8) {
    float f[8];
    load8(p + static_cast<long long>(l) * D, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += f[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[warp][lane * 8 + e] = acc[e];
  __syncthreads();
  const int d = threadIdx.x;
  float s = 0.0f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += part[w][d];
  out[b * out_pitch + mod * mod_stride + d] = to_h16(s / static_cast<float>(L));
}

__global__ void sub_task_embed_kernel(const long long* __restrict__ ids, const float* __restrict__ table, int B,
                                      h16* __restrict__ out, long long out_pitch) {
  RVB_PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 32) return;
  const int b = i / 32, c = i % 32;
  long long id = ids[b];
  id = id < 0 ? 0 : (id > 4 ? 4 : id);
  out[b * out_pitch + c] = to_h16(table[id * 32 + c]);
}

// out[m][o] = dot(y[m], w[o]) + b[o], one warp per output (tiny heads: 512 -> 4 / 2 / 1)
__global__ void heads_linear_kernel(const float* __restrict__ y, int M, int K, const float* __restrict__ w,
                                    const float* __restrict__ b, int n_out, float* __restrict__ out) {
  RVB_PDL_PROLOGUE();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M * n_out) return;
  const int m = warp / n_out, o = warp % n_out;
  float acc = 0.0f;
  for (int k = lane; k < K; k += 32) acc = fmaf(y[static_cast<long long>(m) * K + k], w[static_cast<long long>(o) * K + k], acc);
  acc = warp_sum(acc);
  if (lane == 0) out[static_cast<long long>(m) * n_out + o] = acc + b[o];
}

// Policy heads, one warp per row m (K == 512):
//   outA[m][0..nA) = y[m] . wA^T + bA          (hi: sub-goal logits; lo: linear/angular velocity)
//   outB[m][0..nB) = y[m] . wB^T + bB          (lo: stop logit; optional)
//   amax[m]        = argmax over group A        (hi -> lo hand-off, hierarchical_trainer.py:1098; optional)
//   emb_out[m][0..32) = emb[amax[m]]            (lo's sub_task_embedding of that sub-goal; optional)
constexpr int HEADS_MAX = 5;
__global__ void __launch_bounds__(256) heads_fused_kernel(const float* __restrict__ y, int M, int K,
                                                          const float* __restrict__ wA, const float* __restrict__ bA, int nA,
                                                          float* __restrict__ outA, const float* __restrict__ wB,
                                                          const float* __restrict__ bB, int nB, float* __restrict__ outB,
                                                          long long* __restrict__ amax, const float* __restrict__ emb,
                                                          h16* __restrict__ emb_out, long long emb_pitch) {
  RVB_PDL_PROLOGUE();
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  // K == 512: lane owns k = 4 lane + 128 i.  All loads of the row and of every weight row are
  // issued before the first reduction (one memory round trip instead of one per output).
  const float4* yr = reinterpret_cast<const float4*>(y + static_cast<long long>(m) * K) + lane;
  float4 yv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) yv[i] = yr[32 * i];
  const int n_tot = nA + nB;
  float acc[HEADS_MAX];
#pragma unroll
  for (int o = 0; o < HEADS_MAX; ++o) {
    acc[o] = 0.0f;
    if (o < n_tot) {
      const float4* w = reinterpret_cast<const float4*>(o < nA ? wA + static_cast<long long>(o) * K
                                                               : wB + static_cast<long long>(o - nA) * K) + lane;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 wv = __ldg(w + 32 * i);
        acc[o] = fmaf(yv[i].x, wv.x, acc[o]); acc[o] = fmaf(yv[i].y, wv.y, acc[o]);
        acc[o] = fmaf(yv[i].z, wv.z, acc[o]); acc[o] = fmaf(yv[i].w, wv.w, acc[o]);
      }
    }
  }
  int best = 0;
  float bv = 0.0f;
#pragma unroll
  for (int o = 0; o < HEADS_MAX; ++o) {
    if (o < n_tot) {
      const bool a = o < nA;
      const float v = warp_sum(acc[o]) + (a ? bA[o] : bB[o - nA]);
      if (a) {
        if (lane == 0) outA[static_cast<long long>(m) * nA + o] = v;
        if (o == 0 || v > bv) { bv = v; best = o; }
      } else if (lane == 0) {
        outB[static_cast<long long>(m) * nB + (o - nA)] = v;
      }
    }
  }
  if (amax != nullptr && lane == 0) amax[m] = best;
  if (emb != nullptr) emb_out[m * emb_pitch + lane] = to_h16(emb[best * 32 + lane]);
}

__global__ void argmax_rows_kernel(const float* __restrict__ x, int M, int n, long long* __restrict__ out) {
  RVB_PDL_PROLOGUE();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  int best = 0;
  float bv = x[static_cast<long long>(m) * n];
  for (int j = 1; j < n; ++j) {
    const float v = x[static_cast<long long>(m) * n + j];
    if (v > bv) {
      bv = v;
      best = j;
    }
  }
  out[m] = best;
}

// PE[p,2i] = sin(p / 10000^(2i/D)), PE[p,2i+1] = cos(same)   (common/utils.py:167-185)
__global__ void sinusoid_table_kernel(float* __restrict__ pe, int L, int D) {
  RVB_PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L * (D / 2)) return;
  const int p = i / (D / 2), k = i % (D / 2);
  const float denom = powf(10000.0f, 2.0f * static_cast<float>(k) / static_cast<float>(D));
  const float ang = static_cast<float>(p) / denom;
  pe[static_cast<long long>(p) * D + 2 * k] = sinf(ang);
  pe[static_cast<long long>(p) * D + 2 * k + 1] = cosf(ang);
}

int grid_for(long long total, int block, int max_blocks) {
  const long long g = (total + block - 1) / block;
  return static_cast<int>(std::max<long long>(1, std::min<long long>(g, max_blocks)));
}

}  // namespace

void rgb_stem_im2col(const float* rgb, h16* out, int NB, int H, int W, int Kpitch, cudaStream_t s) {
  RVB_CHECK(Kpitch >= 147 && Kpitch % 8 == 0, "stem: bad K pitch");
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  const long long M = static_cast<long long>(NB) * Ho * Wo;
  const int blocks = static_cast<int>((M + STEM_PIX - 1) / STEM_PIX);
  launch_k(rgb_stem_im2col_kernel, dim3(blocks), dim3(256), STEM_PIX * Kpitch * 2, s, rgb, out, NB, H, W, Ho, Wo, Kpitch);
  RVB_CUDA(cudaGetLastError());
}

void rgb_pad_convert(const float* rgb, h16* out, int NB, int H, int W, int Wp, cudaStream_t s) {
  RVB_CHECK(Wp >= W + 6, "rgb_pad_convert: padded width too small");
  const long long total = static_cast<long long>(NB) * (H + 6) * Wp;
  launch_k(rgb_pad_convert_kernel, dim3(grid_for(total, 256, 148 * 16)), dim3(256), 0, s, rgb, out, NB, H, W, H + 6, Wp);
  RVB_CUDA(cudaGetLastError());
}

void rgb_pad_convert4(const float* rgb, h16* out, int NB, int H, int W, int Wp, cudaStream_t s) {
  RVB_CHECK(Wp >= W + 6 && H % 2 == 0, "rgb_pad_convert4: padded width too small / odd height");
  const long long total = static_cast<long long>(NB) * ((H + 6) / 2) * Wp;
  launch_k(rgb_pad_convert4_kernel<float>, dim3(grid_for(total, 256, 148 * 16)), dim3(256), 0, s, rgb, out, NB, H, W, (H + 6) / 2, Wp);
  RVB_CUDA(cudaGetLastError());
}

void rgb_pad_convert4_u8(const uint8_t* rgb, h16* out, int NB, int H, int W, int Wp, cudaStream_t s) {
  RVB_CHECK(Wp >= W + 6 && H % 2 == 0, "rgb_pad_convert4_u8: padded width too small / odd height");
  const long long total = static_cast<long long>(NB) * ((H + 6) / 2) * Wp;
  launch_k(rgb_pad_convert4_kernel<uint8_t>, dim3(grid_for(total, 256, 148 * 16)), dim3(256), 0, s, rgb, out, NB, H, W, (H + 6) / 2, Wp);
  RVB_CUDA(cudaGetLastError());
}

void maxpool3x3s2(const h16* in, h16* out, int NB, int H, int W, int C, cudaStream_t s) {
  RVB_CHECK(C % 8 == 0, "maxpool: C % 8");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = static_cast<long long>(NB) * Ho * Wo * (C / 8);
  launch_k(maxpool3x3s2_kernel, dim3(grid_for(total, 256, 148 * 16)), dim3(256), 0, s, in, out, NB, H, W, C, Ho, Wo);
  RVB_CUDA(cudaGetLastError());
}

void depth_stem_conv(const float* depth, const float* w, h16* out, int NB, int H, int W, cudaStream_t s) {
  const int Hp = H / 2, Wp = W / 2;
  const int Ho = (Hp + 6 - 7) / 2 + 1, Wo = (Wp + 6 - 7) / 2 + 1;
  RVB_CHECK(W % 2 == 0, "depth stem: even width");
  const size_t smem = (49 * 32 + (2 * DS_ROWS + 5) * (Wp + 6)) * sizeof(float);
  dim3 grid((Ho + DS_ROWS - 1) / DS_ROWS, NB);
  launch_k(depth_stem_kernel, dim3(grid), dim3(256), smem, s, depth, w, out, H, W, Hp, Wp, Ho, Wo);
  RVB_CUDA(cudaGetLastError());
}

void gn_stats(const h16* x, float* stats, int NB, int HW, int C, int G, cudaStream_t s) {
  const int cv = C / 8;
  RVB_CHECK(C % 8 == 0 && cv <= 256 && 256 % cv == 0 && G <= 64 && C % G == 0, "gn_stats: unsupported C/G");
  launch_k(gn_stats_kernel, dim3(NB), dim3(256), 0, s, x, stats, HW, C, G);
  RVB_CUDA(cudaGetLastError());
}

void gn_fused(const GnApply& a, cudaStream_t s) {
  const int cv = a.C / 8;
  RVB_CHECK(a.C % 8 == 0 && a.C % a.G == 0 && a.out_pitch % 8 == 0 && cv <= GNF_THREADS && GNF_THREADS % cv == 0 &&
                a.C <= 2048 && a.G <= 16, "gn_fused: unsupported shape");
  GnApplyDev d;
  d.x = a.x; d.stats = nullptr; d.gamma = a.gamma; d.beta = a.beta;
  d.NB = a.NB; d.HW = a.HW; d.C = a.C; d.G = a.G; d.relu = a.relu; d.res_mode = a.res_mode;
  d.res = a.res; d.res_stats = nullptr; d.res_gamma = a.res_gamma; d.res_beta = a.res_beta;
  d.out = a.out; d.out_pitch = a.out_pitch;
  const size_t smem = static_cast<size_t>(2) * GNF_THREADS * 8 * sizeof(float) + static_cast<size_t>(4) * a.C * sizeof(float);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    RVB_CUDA(cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * GNF_THREADS * 8 * 4 + 4 * 2048 * 4));
  }
  launch_k(gn_fused_kernel, dim3(a.NB), dim3(GNF_THREADS), smem, s, d);
  RVB_CUDA(cudaGetLastError());
}

void gn_apply(const GnApply& a, cudaStream_t s) {
  RVB_CHECK(a.C % 8 == 0 && a.C % a.G == 0 && a.out_pitch % 8 == 0, "gn_apply: bad shape");
  GnApplyDev d;
  d.x = a.x; d.stats = a.stats; d.gamma = a.gamma; d.beta = a.beta;
  d.NB = a.NB; d.HW = a.HW; d.C = a.C; d.G = a.G; d.relu = a.relu; d.res_mode = a.res_mode;
  d.res = a.res; d.res_stats = a.res_stats; d.res_gamma = a.res_gamma; d.res_beta = a.res_beta;
  d.out = a.out; d.out_pitch = a.out_pitch;
  const long long total = static_cast<long long>(a.NB) * a.HW * (a.C / 8);
  launch_k(gn_apply_kernel, dim3(grid_for(total, 256, 148 * 16)), dim3(256), 0, s, d);
  RVB_CUDA(cudaGetLastError());
}

void rgb_pool(const h16* feat, int NB, int H, int W, int C, h16* tokens, int64_t tok_pitch, h16* cellmean,
              int64_t cm_pitch, h16* gmean, cudaStream_t s) {
  RVB_CHECK(C % 8 == 0 && H >= 4 && W >= 4, "rgb_pool: bad shape");
  launch_k(rgb_pool_kernel, dim3(NB), dim3(256), 0, s, feat, H, W, C, tokens, tok_pitch, cellmean, cm_pitch, gmean);
  RVB_CUDA(cudaGetLastError());
}

void fill_spatial_embedding(const float* emb_flat, h16* tokens, int NB, int64_t tok_pitch, int col0, h16* cellmean,
                            int64_t cm_pitch, cudaStream_t s) {
  launch_k(fill_spatial_embedding_kernel, dim3((NB * 64 + 127) / 128), dim3(128), 0, s, emb_flat, tokens, NB, tok_pitch, col0, cellmean,
                                                                      cm_pitch);
  RVB_CUDA(cudaGetLastError());
}

void bert_embed_ln(const int64_t* ids_i64, const float* ids_f32, int id_rows, int R, int L, const float* word,
                   const float* pos, const float* type0, const float* g, const float* b, h16* out, cudaStream_t s) {
  const long long warps = static_cast<long long>(R) * L;
  launch_k(bert_embed_ln_kernel<6>, dim3(static_cast<int>((warps * 32 + 255) / 256)), dim3(256), 0, s, 
      reinterpret_cast<const long long*>(ids_i64), ids_f32, id_rows, R, L, word, pos, type0, g, b, out);
  RVB_CUDA(cudaGetLastError());
}

void layernorm_rows(const float* x, int M, int D, const float* g, const float* b, float eps, const float* pe,
                    int pe_rows, h16* out, cudaStream_t s) {
  // 128-thread CTAs without shared memory: 8 K registers each, so they fit NEXT TO a resident GEMM CTA
  // (320 threads x 160 registers, ~226 KB of shared memory) instead of waiting for a whole free SM
  const int blocks = static_cast<int>((static_cast<long long>(M) * 32 + 127) / 128);
  if (D == 768) launch_k(layernorm_rows_kernel<6>, dim3(blocks), dim3(128), 0, s, x, M, g, b, eps, pe, pe_rows, out);
  else if (D == 256) launch_k(layernorm_rows_kernel<2>, dim3(blocks), dim3(128), 0, s, x, M, g, b, eps, pe, pe_rows, out);
  else RVB_CHECK(false, "layernorm: D must be 256 or 768");
  RVB_CUDA(cudaGetLastError());
}

void token_mean(const h16* x, int n_mod, int B, int L, int D, h16* out, int64_t out_pitch, int64_t mod_stride,
                cudaStream_t s) {
  RVB_CHECK(D == 256, "token_mean: D must be 256");
  launch_k(token_mean_kernel, dim3(n_mod * B), dim3(256), 0, s, x, B, L, D, out, out_pitch, mod_stride);
  RVB_CUDA(cudaGetLastError());
}

void sub_task_embed(const int64_t* ids, const float* table, int B, h16* out, int64_t out_pitch, cudaStream_t s) {
  launch_k(sub_task_embed_kernel, dim3((B * 32 + 127) / 128), dim3(128), 0, s, reinterpret_cast<const long long*>(ids), table, B, out,
                                                             out_pitch);
  RVB_CUDA(cudaGetLastError());
}

void heads_linear(const float* y, int M, int K, const float* w, const float* b, int n_out, float* out,
                  cudaStream_t s) {
  const long long warps = static_cast<long long>(M) * n_out;
  launch_k(heads_linear_kernel, dim3(static_cast<int>((warps * 32 + 255) / 256)), dim3(256), 0, s, y, M, K, w, b, n_out, out);
  RVB_CUDA(cudaGetLastError());
}

void heads_fused(const float* y, int M, int K, const float* wA, const float* bA, int nA, float* outA, const float* wB,
                 const float* bB, int nB, float* outB, int64_t* amax, const float* emb, h16* emb_out, int64_t emb_pitch,
                 cudaStream_t s) {
  RVB_CHECK(nA >= 1 && nB >= 0 && nA + nB <= HEADS_MAX && K == 512 && (nB == 0 || (wB != nullptr && outB != nullptr)),
            "heads_fused: bad groups");
  RVB_CHECK(emb == nullptr || (nA <= 5 && emb_out != nullptr), "heads_fused: embedding table has 5 rows");
  launch_k(heads_fused_kernel, dim3((M * 32 + 255) / 256), dim3(256), 0, s, y, M, K, wA, bA, nA, outA, wB, bB, nB, outB,
           reinterpret_cast<long long*>(amax), emb, emb_out, static_cast<long long>(emb_pitch));
  RVB_CUDA(cudaGetLastError());
}

void argmax_rows(const float* x, int M, int n, int64_t* out, cudaStream_t s) {
  launch_k(argmax_rows_kernel, dim3((M + 127) / 128), dim3(128), 0, s, x, M, n, reinterpret_cast<long long*>(out));
  RVB_CUDA(cudaGetLastError());
}

void sinusoid_table(float* pe, int L, int D, cudaStream_t s) {
  const int total = L * (D / 2);
  launch_k(sinusoid_table_kernel, dim3((total + 255) / 256), dim3(256), 0, s, pe, L, D);
  RVB_CUDA(cudaGetLastError());
}

}  // namespace rvb
