// Token-wise linear layers  C[M,N] = epi(A[M,K] * W[N,K]^T)  with the element-wise / LayerNorm work that
// follows them in the reference fused into the epilogue.  fp32 SIMT (parity mode): 128x64 CTA tile, 256 threads,
// 8x4 register tile, K streamed 16 at a time through shared memory.
// Used by the stage-1 transformer regulariser (models/module.py:507-646) and FMT (models/FMT.py, block.py:336-346).
#pragma once
#include "common.cuh"

namespace mvsf {


struct LinArgs {
  const float* A; int lda;
  const float* W;            // [N][K] row-major (nn.Linear weight)
  const float* bias;         // [N] or nullptr
  float* C; int ldc;
  int M, N, K;
  const float* res; int ldres;
  const float* gamma;        // [N]
  const float* ln_w; const float* ln_b; float ln_eps;
  int elu_cols;
};

constexpr int LBM = 128, LBN = 64, LBK = 16;

template <int EPI>
__global__ void __launch_bounds__(256)
linear_kernel(LinArgs a) {
  __shared__ __align__(16) float As[LBK][LBM + 4];
  __shared__ __align__(16) float Bs[LBK][LBN + 4];
  __shared__ float Cs[(EPI == LIN_RES_LN || EPI == LIN_LN) ? LBM : 1][(EPI == LIN_RES_LN || EPI == LIN_LN) ? LBN + 1 : 1];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.x * LBM, n0 = blockIdx.y * LBN;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < a.K; k0 += LBK) {
    // A tile: 128 rows x 16 k = 512 float4, two per thread
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      int idx = tid + r * 256;
      int m = idx >> 2, kq = idx & 3;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + m < a.M) v = ldg4(a.A + (size_t)(m0 + m) * a.lda + k0 + kq * 4);
      As[kq * 4 + 0][m] = v.x; As[kq * 4 + 1][m] = v.y; As[kq * 4 + 2][m] = v.z; As[kq * 4 + 3][m] = v.w;
    }
    {  // W tile: 64 rows(n) x 16 k = 256 float4, one per thread
      int n = tid >> 2, kq = tid & 3;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + n < a.N) v = ldg4(a.W + (size_t)(n0 + n) * a.K + k0 + kq * 4);
      Bs[kq * 4 + 0][n] = v.x; Bs[kq * 4 + 1][n] = v.y; Bs[kq * 4 + 2][n] = v.z; Bs[kq * 4 + 3][n] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < LBK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = fmaf(av[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(av[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
      }
    }
    __syncthreads();
  }

  const int ncol = n0 + tx * 4;
  float bv[4] = {0.f, 0.f, 0.f, 0.f}, gv[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (ncol + j < a.N) {
      if (a.bias) bv[j] = __ldg(a.bias + ncol + j);
      if (EPI == LIN_RES || EPI == LIN_RES_LN) gv[j] = __ldg(a.gamma + ncol + j);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ml = ty * 8 + i, m = m0 + ml;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x = acc[i][j] + bv[j];
      if (EPI == LIN_GELU) x = gelu_erf(x);
      if (EPI == LIN_ELU1) x = (ncol + j < a.elu_cols) ? ((x > 0.f ? x : expm1f(x)) + 1.0f) : x;
      if (EPI == LIN_RES || EPI == LIN_RES_LN) {
        float r = (m < a.M && ncol + j < a.N) ? a.res[(size_t)m * a.ldres + ncol + j] : 0.f;  // plain load: C may alias res
        x = r + gv[j] * x;
      }
      v[j] = x;
    }
    if (EPI == LIN_RES_LN || EPI == LIN_LN) {
#pragma unroll
      for (int j = 0; j < 4; ++j) Cs[ml][tx * 4 + j] = v[j];
    } else if (m < a.M) {
      float* c = a.C + (size_t)m * a.ldc + ncol;
      if (ncol + 3 < a.N && ((a.ldc & 3) == 0)) {
        *reinterpret_cast<float4*>(c) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ncol + j < a.N) c[j] = v[j];
      }
    }
  }
  if (EPI == LIN_RES_LN || EPI == LIN_LN) {
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    const float w0 = __ldg(a.ln_w + lane), w1 = __ldg(a.ln_w + lane + 32);
    const float b0 = __ldg(a.ln_b + lane), b1 = __ldg(a.ln_b + lane + 32);
    for (int r = warp * 16; r < warp * 16 + 16; ++r) {
      const int m = m0 + r;
      float x0 = Cs[r][lane], x1 = Cs[r][lane + 32];
      float s = x0 + x1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * (1.0f / 64.0f);
      float d0 = x0 - mean, d1 = x1 - mean;
      float q = d0 * d0 + d1 * d1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float sd = sqrtf(q * (1.0f / 64.0f) + a.ln_eps);
      if (m < a.M) {
        a.C[(size_t)m * a.ldc + lane] = __fdiv_rn(d0, sd) * w0 + b0;
        a.C[(size_t)m * a.ldc + lane + 32] = __fdiv_rn(d1, sd) * w1 + b1;
      }
    }
  }
}

static inline int launch_linear(const LinArgs& a, int epi, cudaStream_t s) {
  MVSF_REQUIRE(a.A && a.W && a.C && a.M > 0 && a.N > 0 && a.K > 0, "linear: bad arguments");
  MVSF_REQUIRE(a.K % LBK == 0 && (a.lda % 4) == 0, "linear: K must be a multiple of 16 and lda of 4");
  MVSF_REQUIRE(((uintptr_t)a.A & 15) == 0 && ((uintptr_t)a.W & 15) == 0, "linear: A and W must be 16-byte aligned");
  if (epi == LIN_RES_LN || epi == LIN_LN) MVSF_REQUIRE(a.N == 64 && a.ln_w && a.ln_b, "linear: LayerNorm epilogue needs N == 64");
  if (epi == LIN_RES || epi == LIN_RES_LN) MVSF_REQUIRE(a.res && a.gamma, "linear: residual epilogue needs res and gamma");
  dim3 grid(cdiv(a.M, LBM), cdiv(a.N, LBN));
  switch (epi) {
    case LIN_BIAS: linear_kernel<LIN_BIAS><<<grid, 256, 0, s>>>(a); break;
    case LIN_GELU: linear_kernel<LIN_GELU><<<grid, 256, 0, s>>>(a); break;
    case LIN_ELU1: linear_kernel<LIN_ELU1><<<grid, 256, 0, s>>>(a); break;
    case LIN_RES: linear_kernel<LIN_RES><<<grid, 256, 0, s>>>(a); break;
    case LIN_RES_LN: linear_kernel<LIN_RES_LN><<<grid, 256, 0, s>>>(a); break;
    case LIN_LN: linear_kernel<LIN_LN><<<grid, 256, 0, s>>>(a); break;
    default: return fail(MVSF_ERR_INVALID, "linear: unknown epilogue %d", epi);
  }
  MVSF_LAUNCH_CHECK("linear");
  return MVSF_OK;
}

// Row LayerNorm over 64 channels (nn.LayerNorm(64), eps 1e-5): one warp per row.  block.py:341-345 pre-norm.
static __global__ void __launch_bounds__(256)
layernorm64_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                   float* __restrict__ y, int M, float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  float2 v = ldg2(x + (size_t)row * 64 + lane * 2);
  float s = v.x + v.y;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / 64.0f);
  float d0 = v.x - mean, d1 = v.y - mean;
  float q = d0 * d0 + d1 * d1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float sd = sqrtf(q * (1.0f / 64.0f) + eps);
  float2 ww = ldg2(w + lane * 2), bb = ldg2(b + lane * 2);
  float2 o2 = make_float2(__fdiv_rn(d0, sd) * ww.x + bb.x, __fdiv_rn(d1, sd) * ww.y + bb.y);
  *reinterpret_cast<float2*>(y + (size_t)row * 64 + lane * 2) = o2;
}

}  // namespace mvsf
