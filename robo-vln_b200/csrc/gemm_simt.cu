// Naive CUDA-core implementation of the ConvGemm contract (rvb.h).  It exists to validate the
// tcgen05 kernel on the GPU (tests compare the two on identical inputs) and to bisect
// failures (ROBOVLN_GEMM=simt routes every contraction through it).  It is a GPU kernel; the
// library has no CPU path.
#include "common.cuh"
#include "rvb.h"

namespace rvb {

namespace {

struct SimtParams {
  const h16* in; const h16* w; const float* bias; const h16* res; void* out;
  int NB, H, W, Cin, Cout, KH, KW, stride, pad, Ho, Wo;
  long long in_pitch, ldr, ldc, M;
  int res_rows, act, out_f32;
  int window; long long win_row_pitch;
};

__global__ void gemm_simt_kernel(const SimtParams p) {
  RVB_PDL_PROLOGUE();
  const int n = blockIdx.y * blockDim.x + threadIdx.x;
  const long long m = static_cast<long long>(blockIdx.x) * blockDim.y + threadIdx.y;
  if (n >= p.Cout || m >= p.M) return;
  const int wo = static_cast<int>(m % p.Wo);
  const int ho = static_cast<int>((m / p.Wo) % p.Ho);
  const int img = static_cast<int>(m / (static_cast<long long>(p.Wo) * p.Ho));
  const long long Ktot = static_cast<long long>(p.KH) * p.KW * p.Cin;
  float acc = 0.0f;
  if (p.window == 2) {   // packed stem: K block kb = row pair ho + kb, 64 contiguous elements (8 px x 2 rows x 4 ch)
    for (int kb = 0; kb < p.KH; ++kb) {
      const h16* a = p.in + (static_cast<long long>(img) * p.H + ho + kb) * p.win_row_pitch +
                     static_cast<long long>(wo) * p.stride * 8;
      const h16* b = p.w + n * Ktot + static_cast<long long>(kb) * p.Cin;
      for (int c = 0; c < p.Cin; ++c) acc = fmaf(from_h16(a[c]), from_h16(b[c]), acc);
    }
  } else if (p.window) {
    for (int r = 0; r < p.KH; ++r) {
      const h16* a = p.in + (static_cast<long long>(img) * p.H + ho * p.stride + r) * p.win_row_pitch +
                     static_cast<long long>(wo) * p.stride * 8;
      const h16* b = p.w + n * Ktot + static_cast<long long>(r) * p.Cin;
      for (int c = 0; c < p.Cin; ++c) acc = fmaf(from_h16(a[c]), from_h16(b[c]), acc);
    }
  }
  for (int r = 0; r < (p.window ? 0 : p.KH); ++r) {
    const int h = ho * p.stride + r - p.pad;
    if (h < 0 || h >= p.H) continue;
    for (int s = 0; s < p.KW; ++s) {
      const int w = wo * p.stride + s - p.pad;
      if (w < 0 || w >= p.W) continue;
      const h16* a = p.in + ((static_cast<long long>(img) * p.H + h) * p.W + w) * p.in_pitch;
      const h16* b = p.w + n * Ktot + static_cast<long long>(r * p.KW + s) * p.Cin;
      for (int c = 0; c < p.Cin; ++c) acc = fmaf(from_h16(a[c]), from_h16(b[c]), acc);
    }
  }
  if (p.bias != nullptr) acc += p.bias[n];
  if (p.res != nullptr) {
    const long long rr = p.res_rows > 0 ? (m % p.res_rows) : m;
    acc += from_h16(p.res[rr * p.ldr + n]);
  }
  if (p.act == ACT_RELU) acc = fmaxf(acc, 0.0f);
  else if (p.act == ACT_GELU) acc = gelu_erf(acc);
  if (p.out_f32) reinterpret_cast<float*>(p.out)[m * p.ldc + n] = acc;
  else reinterpret_cast<h16*>(p.out)[m * p.ldc + n] = to_h16(acc);
}

}  // namespace

void gemm_simt_launch(const ConvGemm& g, cudaStream_t stream) {
  SimtParams p;
  p.in = g.in; p.w = g.w; p.bias = g.bias; p.res = g.res; p.out = g.out;
  p.NB = g.NB; p.H = g.H; p.W = g.W; p.Cin = g.Cin; p.Cout = g.Cout; p.KH = g.KH; p.KW = g.KW;
  p.stride = g.stride; p.pad = g.pad; p.Ho = g.Ho(); p.Wo = g.Wo();
  p.in_pitch = g.in_pitch; p.ldr = g.ldr; p.ldc = g.ldc; p.M = g.M();
  p.res_rows = g.res_rows; p.act = g.act; p.out_f32 = g.out_f32;
  p.window = g.window; p.win_row_pitch = g.win_row_pitch;
  dim3 block(32, 8);
  const long long gx = (p.M + block.y - 1) / block.y;
  RVB_CHECK(gx < (1ll << 31), "simt gemm: M too large");
  dim3 grid(static_cast<unsigned>(gx), (p.Cout + block.x - 1) / block.x);
  launch_k(gemm_simt_kernel, dim3(grid), dim3(block), 0, stream, p);
  RVB_CUDA(cudaGetLastError());
}

}  // namespace rvb
