// The 16-bit storage / tensor-core operand type of the library, fixed at compile time:
//   -DRVB_BF16=0 (default)  IEEE fp16  (11-bit significand) -> librobovln_b200.so
//   -DRVB_BF16=1            bfloat16   ( 8-bit significand) -> librobovln_b200_bf16.so
// Both run tcgen05 kind::f16 at the same rate with fp32 accumulation.  fp16 is the default
// because every sub-network on this path except the RGB trunk is normalised layer by layer
// (GroupNorm / LayerNorm), so range is not a concern while bf16's 8-bit rounding, amplified
// through the 54-layer GroupNorm trunk and 12 BERT layers, misses the 1e-2 tolerance on the
// LSTM state.  fp16 conversions saturate instead of producing inf.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#ifndef RVB_BF16
#define RVB_BF16 0
#endif

namespace rvb {

#if RVB_BF16
using h16 = __nv_bfloat16;
#define RVB_H16_NAME "bf16"
#define RVB_H16_CODE 1   /* HCM_BF16 */
#else
using h16 = __half;
#define RVB_H16_NAME "fp16"
#define RVB_H16_CODE 3   /* HCM_F16 */
#endif

}  // namespace rvb
