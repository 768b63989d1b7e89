// One-time / per-call plumbing kernels that used to be chains of ATen launches:
//  * pack_weight: fp32 OIHW (or [O, K]) parameter -> the K-major 16-bit layout gemm_tc.cu walks
//    (k = (r*KW + s)*I + c), with eval-mode BatchNorm folded in (scale into the weights, shift as an fp32
//    bias; SURVEY.md A.6, torchvision ResNet-50 behind resnet_encoders.py:151).  One launch per parameter, so a
//    freshly constructed policy reaches its first engine kernel after ~250 launches instead of > 1000.
//  * compare_many: are N pairs of device buffers bit-identical?  (hi's and lo's frozen trunks: the
//    de-duplication of SURVEY.md 7.2 is legal only when they are.)  One launch for all pairs.
//  * checksum: 128-bit order-independent content checksum of a buffer.  lo reuses the trunk features hi just
//    computed only when the observation CONTENT is the same (hierarchical_trainer.py:1096-1100 passes the
//    same batch to both models); pointer / version heuristics cannot see writes through numpy views or
//    interop pointers.
#include "common.cuh"
#include "rvb.h"

#include <algorithm>

namespace rvb {

namespace {

__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ mean,
                                                          const float* __restrict__ var, float eps, h16* __restrict__ out,
                                                          float* __restrict__ bias_out, int O, int I, int KH, int KW,
                                                          long long out_pitch) {
  const long long K = static_cast<long long>(KH) * KW * I;
  const long long total = static_cast<long long>(O) * K;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(idx / K);
    const long long k = idx - static_cast<long long>(o) * K;
    const int tap = static_cast<int>(k / I);
    const int c = static_cast<int>(k - static_cast<long long>(tap) * I);
    const int r = tap / KW, s = tap - r * KW;
    float scale = 1.0f;
    if (gamma != nullptr) scale = gamma[o] / sqrtf(var[o] + eps);   // same operation order as the fp32 reference fold
    const float v = w[((static_cast<long long>(o) * I + c) * KH + r) * KW + s] * scale;
    out[static_cast<long long>(o) * out_pitch + k] = to_h16(v);
    if (bias_out != nullptr && k == 0) bias_out[o] = beta[o] - mean[o] * scale;
  }
}

struct PairList {
  const uint32_t* const* a;
  const uint32_t* const* b;
  const long long* words;
};

__global__ void __launch_bounds__(256) compare_many_kernel(PairList pl, int* __restrict__ mismatch) {
  const int t = blockIdx.y;
  const uint32_t* a = pl.a[t];
  const uint32_t* b = pl.b[t];
  const long long n = pl.words[t];
  if (a == b) return;
  int bad = 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    bad |= (a[i] != b[i]);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(mismatch, 1);
}

RVB_DEVICE unsigned long long mix64(unsigned long long x) {   // splitmix64 finaliser
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

// out[0] += sum_i mix(word_i, i) ; out[1] += sum_i word_i  (both mod 2^64: order independent -> deterministic)
__global__ void __launch_bounds__(256) checksum_kernel(const uint4* __restrict__ p, long long n16, const uint8_t* __restrict__ tail,
                                                       int tail_bytes, unsigned long long* __restrict__ out) {
  unsigned long long h = 0, s = 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n16;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 v = __ldg(p + i);
    const unsigned long long lo = (static_cast<unsigned long long>(v.y) << 32) | v.x;
    const unsigned long long hi = (static_cast<unsigned long long>(v.w) << 32) | v.z;
    h += mix64(lo ^ (static_cast<unsigned long long>(2 * i) * 0x9e3779b97f4a7c15ull));
    h += mix64(hi ^ (static_cast<unsigned long long>(2 * i + 1) * 0x9e3779b97f4a7c15ull));
    s += lo + hi;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int j = 0; j < tail_bytes; ++j) {
      h += mix64(static_cast<unsigned long long>(tail[j]) ^ (static_cast<unsigned long long>(2 * n16 + j) * 0xd6e8feb86659fd93ull));
      s += tail[j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    h += __shfl_xor_sync(0xffffffffu, h, o);
    s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out, h);
    atomicAdd(out + 1, s);
  }
}

}  // namespace

void pack_weight(const float* w, const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                 h16* out, float* bias_out, int O, int I, int KH, int KW, int64_t out_pitch, cudaStream_t s) {
  RVB_CHECK(w != nullptr && out != nullptr && O > 0 && I > 0 && KH > 0 && KW > 0, "pack_weight: bad arguments");
  RVB_CHECK((gamma == nullptr) == (bias_out == nullptr), "pack_weight: BatchNorm fold needs a bias output (and vice versa)");
  RVB_CHECK(gamma == nullptr || (beta != nullptr && mean != nullptr && var != nullptr), "pack_weight: incomplete BatchNorm");
  const long long total = static_cast<long long>(O) * KH * KW * I;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8));
  launch_k(pack_weight_kernel, dim3(grid), dim3(256), 0, s, w, gamma, beta, mean, var, eps, out, bias_out, O, I, KH, KW,
           static_cast<long long>(out_pitch > 0 ? out_pitch : static_cast<int64_t>(KH) * KW * I));
  RVB_CUDA(cudaGetLastError());
}

void compare_many(const void* const* a_dev, const void* const* b_dev, const long long* words_dev, int n, int* mismatch_dev,
                  cudaStream_t s) {
  RVB_CHECK(n > 0, "compare_many: empty list");
  RVB_CUDA(cudaMemsetAsync(mismatch_dev, 0, sizeof(int), s));
  PairList pl{reinterpret_cast<const uint32_t* const*>(a_dev), reinterpret_cast<const uint32_t* const*>(b_dev), words_dev};
  launch_k(compare_many_kernel, dim3(32, n), dim3(256), 0, s, pl, mismatch_dev);
  RVB_CUDA(cudaGetLastError());
}

void checksum(const void* p, size_t bytes, unsigned long long* out2_dev, cudaStream_t s) {
  RVB_CHECK((reinterpret_cast<uintptr_t>(p) & 15) == 0, "checksum: pointer must be 16-byte aligned");
  RVB_CUDA(cudaMemsetAsync(out2_dev, 0, 16, s));
  const long long n16 = static_cast<long long>(bytes / 16);
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((n16 + 255) / 256, 148 * 8)));
  launch_k(checksum_kernel, dim3(grid), dim3(256), 0, s, reinterpret_cast<const uint4*>(p), n16,
           reinterpret_cast<const uint8_t*>(p) + n16 * 16, static_cast<int>(bytes - static_cast<size_t>(n16) * 16), out2_dev);
  RVB_CUDA(cudaGetLastError());
}

}  // namespace rvb
