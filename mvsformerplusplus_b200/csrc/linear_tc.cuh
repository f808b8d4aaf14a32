// Host-side interface of the tcgen05 linear layers (linear_tc.cu).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace mvsf {

struct TcLinArgs {
  const __half* A2; int lda2;   // activations, fp16 hi|lo split: row m = [hi(0..K) | lo(0..K)], lda2 >= 2K (elements)
  const __half* B2;             // weights [N][2K] fp16 hi|lo split (nn.Linear weight [N][K])
  int M, N, K;
  const float* bias;            // [N] or nullptr
  const float* res; int ldres;  // residual (LIN_RES / LIN_RES_LN)
  const float* gamma;           // [N]
  const float* ln_w; const float* ln_b; float ln_eps;
  int elu_cols;
  float* C; int ldc;            // fp32 result (may be nullptr when only the split result is needed)
  __half* C2; int ldc2;         // optional fp16 hi|lo split of the result: row m = [hi(0..N) | lo(0..N)]
};

int launch_linear_tc(const TcLinArgs& a, int epi, cudaStream_t s);
int launch_split_f16(const float* x, int ldx, __half* out, int ldo, int M, int K, cudaStream_t s);

}  // namespace mvsf
