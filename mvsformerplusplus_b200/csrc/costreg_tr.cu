// R1: stage-1 transformer cost regulariser, models/module.py:602-646 (PureTransformerCostReg) with
// FlashAttnBlock post-norm blocks (:569-582), LayerNorm3D (:586-599), PositionEncoding3D
// (models/position_encoding.py:164-189) and the entropy-invariance softmax scale (attention.py:158-161).
//
// volume [D][H][W][8] --(+pe_proj(PE3D))--> patchify (2,4,4) GEMM K=256 + LN3D --> tokens [N][64]
//   6 x { qkv GEMM -> softmax attention (4 heads x 16) -> proj GEMM + gamma1/residual/LN -> FFN GEMMs + gamma2/residual/LN }
//   --> un-patchify GEMM (64 -> 2*4*4*8) --> per-voxel LN3D(8) + 1x1x1 prob --> logits [D][H][W]
// Softmax attention is order-free over tokens, so tokens are kept in (d', h', w') raster order instead of the
// reference's "(h w d)" (module.py:573) - the result is identical.
#include <cuda_fp16.h>
#include <float.h>

#include "linear.cuh"
#include <stdlib.h>

#include "linear_tc.cuh"
#include "umma.cuh"

namespace mvsf {

// packed weights (floats), produced by packing.pack_costreg_tr:
//   pe_w[8][24] | down_w[64][256] (k = ((kd*4+kh)*4+kw)*8+ci) | down_b[64] | down_ln_w[64] | down_ln_b[64]
//   per layer: qkv_w[192][64] | proj_w[64][64] | proj_b[64] | gamma1[64] | n1_w[64] | n1_b[64]
//              | f1_w[256][64] | f1_b[256] | f2_w[64][256] | f2_b[64] | gamma2[64] | n2_w[64] | n2_b[64]
//   up_w[256][64] (n = ((kd*4+kh)*4+kw)*8+co) | up_b[256] | up_ln_w[8] | up_ln_b[8] | prob_w[8] | prob_b[1] (+3 pad)
constexpr int TR_PE = 0, TR_DOWN_W = 192, TR_DOWN_B = TR_DOWN_W + 64 * 256, TR_DOWN_LNW = TR_DOWN_B + 64,
              TR_DOWN_LNB = TR_DOWN_LNW + 64, TR_LAYER0 = TR_DOWN_LNB + 64;
constexpr int L_QKV = 0, L_PROJ_W = 192 * 64, L_PROJ_B = L_PROJ_W + 64 * 64, L_G1 = L_PROJ_B + 64, L_N1W = L_G1 + 64,
              L_N1B = L_N1W + 64, L_F1W = L_N1B + 64, L_F1B = L_F1W + 256 * 64, L_F2W = L_F1B + 256,
              L_F2B = L_F2W + 64 * 256, L_G2 = L_F2B + 64, L_N2W = L_G2 + 64, L_N2B = L_N2W + 64, TR_LAYER = L_N2B + 64;
constexpr int U_W = 0, U_B = 256 * 64, U_LNW = U_B + 256, U_LNB = U_LNW + 8, U_PW = U_LNB + 8, U_PB = U_PW + 8;

// volume[d,y,x,:] += pe_w (8x24) * PE3D(pos[:,d,y,x])
__global__ void pe3d_add_kernel(float* __restrict__ vol, const float* __restrict__ pos, const float* __restrict__ pe_w,
                                size_t nvox) {
  __shared__ float w[8 * 24];
  for (int i = threadIdx.x; i < 192; i += blockDim.x) w[i] = __ldg(pe_w + i);
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvox) return;
  // div_term = exp(arange(0,8,2) * (-ln(1e4)/8))  (position_encoding.py:169)
  const float div[4] = {1.0f, 0.1f, 0.01f, 0.001f};
  float pe[24];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float p = __ldg(pos + a * nvox + i);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float arg = __fmul_rn(__fmul_rn(p, 4.0f), div[k]);
      pe[a * 8 + 2 * k] = sinf(arg);
      pe[a * 8 + 2 * k + 1] = cosf(arg);
    }
  }
  float4* v = reinterpret_cast<float4*>(vol + i * 8);
  float4 v0 = v[0], v1 = v[1];
  float o[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 24; ++k) s = fmaf(w[c * 24 + k], pe[k], s);
    o[c] += s;
  }
  v[0] = make_float4(o[0], o[1], o[2], o[3]);
  v[1] = make_float4(o[4], o[5], o[6], o[7]);
}

// im2col for the (2,4,4)/(2,4,4) patchify conv, emitted as the fp16 hi|lo split the tensor-core GEMM consumes:
// patches2[t] = [hi(256) | lo(256)], k = ((kd*4+kh)*4+kw)*8 + ci
__global__ void patch_gather_kernel(const float* __restrict__ vol, __half* __restrict__ patches2, int D, int H, int W) {
  const int Hp = H / 4, Wp = W / 4;
  size_t total = (size_t)(D / 2) * Hp * Wp * 32;  // one thread per (token, voxel): 8 channels
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int vox = (int)(i & 31);
  size_t t = i >> 5;
  int wp = (int)(t % Wp), hp = (int)((t / Wp) % Hp), dp = (int)(t / ((size_t)Wp * Hp));
  int kw = vox & 3, kh = (vox >> 2) & 3, kd = vox >> 4;
  const float* src = vol + (((size_t)(dp * 2 + kd) * H + hp * 4 + kh) * W + wp * 4 + kw) * 8;
  float4 a = ldg4(src), b = ldg4(src + 4);
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  __half* row = patches2 + t * 512 + vox * 8;
  split_store8(row, row + 256, v);
}

__device__ __forceinline__ uint32_t pack_f16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t u) {
  __half2 h = *reinterpret_cast<__half2*>(&u);
  return __half22float2(h);  // .x = low half, .y = high half
}
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}
// split layout: 6 planes of 4*16*Np halves (Np = N rounded up to 8): 0 Qh, 1 Ql, 2 Kh, 3 Kl as [4 heads][Np][16];
// 4 Vh, 5 Vl TRANSPOSED per head as [4 heads][16 dims][Np] (keys contiguous: K-major B operand of the tcgen05 P*V
// product).  Pad keys [N, Np) of V are zero.
__global__ void qkv_split_kernel(const float* __restrict__ qkv, __half* __restrict__ split, int N, int Np, float qscale) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;  // (token, which(q/k/v), head, quad of 4 dims)
  int total = Np * 3 * 4 * 4;
  if (i >= total) return;
  int quad = i & 3, h = (i >> 2) & 3, which = (i >> 4) % 3, tok = i / 48;
  const size_t plane = (size_t)4 * 16 * Np;
  if (tok >= N) {
    if (which == 2) {
      __half* th = split + (size_t)4 * plane + ((size_t)h * 16 + quad * 4) * Np + tok;
      __half* tl = split + (size_t)5 * plane + ((size_t)h * 16 + quad * 4) * Np + tok;
      for (int e = 0; e < 4; ++e) { th[(size_t)e * Np] = __float2half_rn(0.f); tl[(size_t)e * Np] = __float2half_rn(0.f); }
    }
    return;
  }
  float4 v = ldg4(qkv + (size_t)tok * 192 + which * 64 + h * 16 + quad * 4);
  if (which == 0) { v.x *= qscale; v.y *= qscale; v.z *= qscale; v.w *= qscale; }
  __half hi[4], lo[4];
  split_f16(v.x, hi[0], lo[0]); split_f16(v.y, hi[1], lo[1]); split_f16(v.z, hi[2], lo[2]); split_f16(v.w, hi[3], lo[3]);
  if (which == 2) {
    __half* th = split + (size_t)4 * plane + ((size_t)h * 16 + quad * 4) * Np + tok;
    __half* tl = split + (size_t)5 * plane + ((size_t)h * 16 + quad * 4) * Np + tok;
#pragma unroll
    for (int e = 0; e < 4; ++e) { th[(size_t)e * Np] = hi[e]; tl[(size_t)e * Np] = lo[e]; }
    return;
  }
  size_t off = ((size_t)h * Np + tok) * 16 + quad * 4;
  __half2* ph = reinterpret_cast<__half2*>(split + (size_t)(which * 2) * plane + off);
  __half2* pl = reinterpret_cast<__half2*>(split + (size_t)(which * 2 + 1) * plane + off);
  ph[0] = __halves2half2(hi[0], hi[1]); ph[1] = __halves2half2(hi[2], hi[3]);
  pl[0] = __halves2half2(lo[0], lo[1]); pl[1] = __halves2half2(lo[2], lo[3]);
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#include "attention_fa.cuh"   // second-generation attention kernel (inside namespace mvsf)

// ------------------------------------------------------------------------------------------------------------------
// tcgen05 softmax attention, first generation (kept for reference measurements: MVSF_ATTENTION_V3=1).  CTA = 128 queries x one head, 128 threads, thread t owns query row t
// (TMEM lane t), so row max / row sum need no cross-thread traffic.  Per 128-key tile:
//   S[128x128] = Q K^T      : 3 split-fp16 tcgen05.mma (lo*hi, hi*lo, hi*hi), K = 16, fp32 accumulators in TMEM
//   softmax                  : tcgen05.ld the row, online max, P = exp2(S - m), split P into fp16 hi/lo, store both in the
//                              canonical K-major A-operand layout in shared memory
//   Otile[128x16] = P V      : 8 k-steps x 3 split products = 24 tcgen05.mma (N = 16), V^T tiles are K-major B operands
//   O = O * corr + Otile     : round-to-nearest FMA in registers (tensor-core accumulation truncates; chaining all
//                              tiles on one accumulator biases the result, see profiles/)
// Two CTAs per SM (<= 113 KB smem, 256 TMEM columns each): one CTA's softmax overlaps the other's MMAs.
// ------------------------------------------------------------------------------------------------------------------
// Debug build (-DMVSF_FA_TRACE, tools/fa_trace.py): clock64 stamps of thread 0's phases for 64 key tiles of one CTA.
#ifdef MVSF_FA_TRACE
__device__ long long g_fa_trace[64 * 8];
#define FA_TRACE(slot) do { if (blockIdx.x == 100 && blockIdx.y == 1 && tid == 0 && j >= 100 && j < 164) g_fa_trace[(j - 100) * 8 + (slot)] = clock64(); } while (0)
#else
#define FA_TRACE(slot) do {} while (0)
#endif
namespace fa5 {
using namespace umma;
constexpr int BM = 128, BN = 128, THREADS = 256;
constexpr uint32_t LBO_QK = 16 * 128 + 16;   // 2 chunks (hd = 16) x 128 rows
constexpr uint32_t QK_TILE = 2 * LBO_QK;
constexpr uint32_t LBO_P = 16 * 128;         // 16 chunks (128 keys) x 128 rows
constexpr uint32_t P_TILE = 16 * LBO_P;      // 32 KB
constexpr uint32_t LBO_V = 2 * 128 + 16;     // 16 chunks (128 keys) x 16 rows (head dims)
constexpr uint32_t V_TILE = 16 * LBO_V;
// Q (hi,lo) | K ring 2 x (hi,lo) | V^T ring 2 x (hi,lo) | P (hi,lo) | row-max exchange [2][128] | barriers
constexpr uint32_t OFF_Q = 0, OFF_K = 2 * QK_TILE, OFF_V = OFF_K + 4 * QK_TILE, OFF_P = OFF_V + 4 * V_TILE,
                   OFF_X = OFF_P + 2 * P_TILE, OFF_BAR = OFF_X + 1024;
constexpr uint32_t SMEM = OFF_BAR + 64;
}  // namespace fa5

// 256 threads: thread (warp w, lane) owns query row (w%4)*32+lane and the key columns [64*(w/4), 64*(w/4)+64) of each
// tile; the two threads of a row exchange their partial row maxima through shared memory.  The normaliser l is summed
// per tile and folded like the outputs (l = l*corr + tile_sum): a single running fp32 sum over 27k keys is 5x noisier.
__global__ void __launch_bounds__(256, 2)
attention_tc_kernel(const __half* __restrict__ split, float* __restrict__ out, __half* __restrict__ out2, int N, int Np) {
  using namespace fa5;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = warp >> 2;                       // which 64 key columns of a tile
  const int row = (warp & 3) * 32 + lane;           // query row == TMEM lane
  const int h = blockIdx.y;
  const int q0 = blockIdx.x * BM;
  const size_t plane = (size_t)4 * 16 * Np;
  const __half* Qg[2] = {split + 0 * plane + (size_t)h * Np * 16, split + 1 * plane + (size_t)h * Np * 16};
  const __half* Kg[2] = {split + 2 * plane + (size_t)h * Np * 16, split + 3 * plane + (size_t)h * Np * 16};
  const __half* Vg[2] = {split + 4 * plane + (size_t)h * 16 * Np, split + 5 * plane + (size_t)h * 16 * Np};  // [16][Np]
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar_s = sb + OFF_BAR, bar_o = sb + OFF_BAR + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 32);
  volatile float* xchg = reinterpret_cast<volatile float*>(smem + OFF_X);  // [2][128]

  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(sb + OFF_BAR + 32, 256);

  auto load_qk_tile = [&](uint32_t dst, const __half* g, int row0) {  // 128 rows x 2 chunks, zero fill beyond N
    int r = tid >> 1, c = tid & 1;
    bool ok = row0 + r < N;
    cp_async16_zfill(dst + c * LBO_QK + (r >> 3) * 128 + (r & 7) * 16, g + (size_t)(ok ? row0 + r : 0) * 16 + c * 8, ok);
  };
  auto load_v_tile = [&](uint32_t dst, const __half* g, int key0) {   // 16 rows (dims) x 16 chunks of 8 keys
    {
      int r = tid >> 4, c = tid & 15;
      bool ok = key0 + c * 8 < Np;  // rows are padded to Np (multiple of 8) with zeros: chunks are whole
      cp_async16_zfill(dst + c * LBO_V + (r >> 3) * 128 + (r & 7) * 16, g + (size_t)r * Np + (ok ? key0 + c * 8 : 0), ok);
    }
  };
  const int ntiles = (N + BN - 1) / BN;
  auto load_k = [&](int tile) {  // always commits a group (possibly empty) to keep the group accounting uniform
    if (tile < ntiles) {
      const uint32_t s0 = sb + OFF_K + (tile & 1) * 2 * QK_TILE;
      load_qk_tile(s0, Kg[0], tile * BN);
      load_qk_tile(s0 + QK_TILE, Kg[1], tile * BN);
    }
    cp_async_commit_group();
  };
  auto load_v = [&](int tile) {
    if (tile < ntiles) {
      const uint32_t s0 = sb + OFF_V + (tile & 1) * 2 * V_TILE;
      load_v_tile(s0, Vg[0], tile * BN);
      load_v_tile(s0 + V_TILE, Vg[1], tile * BN);
    }
    cp_async_commit_group();
  };
  const uint32_t idesc_s = make_idesc_f16(128, 128), idesc_o = make_idesc_f16(128, 16);
  auto issue_s = [&](uint32_t tS, int tile) {  // thread 0 only
    const uint32_t sK = sb + OFF_K + (tile & 1) * 2 * QK_TILE;
    const uint64_t qh = make_desc(sb + OFF_Q, LBO_QK, 128), ql = make_desc(sb + OFF_Q + QK_TILE, LBO_QK, 128);
    const uint64_t kh = make_desc(sK, LBO_QK, 128), kl = make_desc(sK + QK_TILE, LBO_QK, 128);
    mma_f16_ss(tS, ql, kh, idesc_s, 0u);
    mma_f16_ss(tS, qh, kl, idesc_s, 1u);
    mma_f16_ss(tS, qh, kh, idesc_s, 1u);
    commit(bar_s);
  };

  load_qk_tile(sb + OFF_Q, Qg[0], q0);
  load_qk_tile(sb + OFF_Q + QK_TILE, Qg[1], q0);
  load_k(0);   // group: Q + K(0)
  load_v(0);
  load_k(1);
  cp_async_wait_group<0>();
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 128;
  const uint32_t trow = ((uint32_t)((warp & 3) * 32)) << 16;
  const uint32_t prow = sb + OFF_P + (row >> 3) * 128 + (row & 7) * 16;  // this thread's row inside every P chunk
  if (tid == 0) issue_s(tS, 0);

  float o[8];   // this thread's 8 of the 16 head dims: dims [8*half, 8*half+8)
#pragma unroll
  for (int d = 0; d < 8; ++d) o[d] = 0.f;
  float m = -1e30f, l = 0.f, corr_prev = 1.0f;

  for (int j = 0; j < ntiles; ++j) {
    // ---- 1. this thread's 64 columns of S(j) -> registers (single wait), partial row maximum -> exchange buffer
    FA_TRACE(0);
    mbar_wait(bar_s, (uint32_t)(j & 1));
    tc_fence_after_sync();
    FA_TRACE(1);
    uint32_t sr[2][32];
    tmem_ld32_nowait(tS + trow + half * 64, sr[0]);
    tmem_ld32_nowait(tS + trow + half * 64 + 32, sr[1]);
    tmem_ld_wait();
    if (j * BN + BN > N) {  // last, partial tile only (uniform branch): keys >= N never win the max and get P = 0
      const int kbase = j * BN + half * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (kbase + c * 32 + e >= N) sr[c][e] = 0xf149f2caU;  // -1e30f
    }
    float pmax = -1e30f;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int e = 0; e < 32; ++e) pmax = fmaxf(pmax, __uint_as_float(sr[c][e]));
    xchg[half * 128 + row] = pmax;
    // ---- 2. K(j+2) prefetch; S(j+1) starts as soon as everybody has S(j) in registers
    load_k(j + 2);
    cp_async_wait_group<1>();   // K(j+1) and V(j) have landed
    fence_proxy_async();
    tc_fence_before_sync();
    FA_TRACE(2);
    __syncthreads();
    FA_TRACE(3);
    if (tid == 0 && j + 1 < ntiles) { tc_fence_after_sync(); issue_s(tS, j + 1); }
    // ---- 3. row maximum, fold O_tile(j-1)
    const float mx = fmaxf(m, fmaxf(pmax, xchg[(half ^ 1) * 128 + row]));
    const float corr = ex2f(m - mx);
    m = mx;
    FA_TRACE(4);
    if (j > 0) {
      mbar_wait(bar_o, (uint32_t)((j - 1) & 1));
      tc_fence_after_sync();
      float ot[16];
      tmem_ld16(tO + trow, ot);
#pragma unroll
      for (int d = 0; d < 8; ++d) o[d] = fmaf(o[d], corr_prev, half ? ot[8 + d] : ot[d]);
    }
    corr_prev = corr;
    FA_TRACE(5);
    float tsum = 0.f;
    load_v(j + 1);   // its stage held V(j-1), released by the P*V product we just waited for
    // ---- 4. P = exp2(S - m), hi/lo split -> shared memory (A operand of P*V); chunk = 8 keys
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = c8 * 8 + 2 * e;
        const float p0 = ex2f(__uint_as_float(sr[col >> 5][col & 31]) - m);
        const float p1 = ex2f(__uint_as_float(sr[(col + 1) >> 5][(col + 1) & 31]) - m);
        tsum += p0 + p1;
        const __half2 hh = __floats2half2_rn(p0, p1);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(p0 - hf.x, p1 - hf.y);
        ph[e] = *reinterpret_cast<const uint32_t*>(&hh);
        pl[e] = *reinterpret_cast<const uint32_t*>(&ll);
      }
      const uint32_t dst = prow + (half * 8 + c8) * LBO_P;
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(ph[0]), "r"(ph[1]), "r"(ph[2]), "r"(ph[3]) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + P_TILE), "r"(pl[0]), "r"(pl[1]), "r"(pl[2]), "r"(pl[3]) : "memory");
    }
    l = fmaf(l, corr, tsum);
    // ---- 5. O_tile(j) = P(j) V(j)
    fence_proxy_async();
    tc_fence_before_sync();
    FA_TRACE(6);
    __syncthreads();
    FA_TRACE(7);
    if (tid == 0) {
      tc_fence_after_sync();
      const uint32_t sV = sb + OFF_V + (j & 1) * 2 * V_TILE;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint64_t pH = make_desc(sb + OFF_P + 2 * i * LBO_P, LBO_P, 128);
        const uint64_t pL = make_desc(sb + OFF_P + P_TILE + 2 * i * LBO_P, LBO_P, 128);
        const uint64_t vH = make_desc(sV + 2 * i * LBO_V, LBO_V, 128);
        const uint64_t vL = make_desc(sV + V_TILE + 2 * i * LBO_V, LBO_V, 128);
        mma_f16_ss(tO, pL, vH, idesc_o, i > 0 ? 1u : 0u);
        mma_f16_ss(tO, pH, vL, idesc_o, 1u);
        mma_f16_ss(tO, pH, vH, idesc_o, 1u);
      }
      commit(bar_o);
    }
  }
  {  // fold the last tile
    mbar_wait(bar_o, (uint32_t)((ntiles - 1) & 1));
    tc_fence_after_sync();
    float ot[16];
    tmem_ld16(tO + trow, ot);
#pragma unroll
    for (int d = 0; d < 8; ++d) o[d] = fmaf(o[d], corr_prev, half ? ot[8 + d] : ot[d]);
  }
  cp_async_wait_group<0>();
  // the two threads of a row summed disjoint key columns: combine the normalisers
  xchg[half * 128 + row] = l;
  __syncthreads();
  l += xchg[(half ^ 1) * 128 + row];

  const int r = q0 + row;
  if (r < N) {
    const float inv = __fdiv_rn(1.0f, l);
    float res[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) res[d] = o[d] * inv;
    const int col = h * 16 + half * 8;
    if (out) {
      *reinterpret_cast<float4*>(out + (size_t)r * 64 + col) = make_float4(res[0], res[1], res[2], res[3]);
      *reinterpret_cast<float4*>(out + (size_t)r * 64 + col + 4) = make_float4(res[4], res[5], res[6], res[7]);
    }
    if (out2) split_store8(out2 + (size_t)r * 128 + col, out2 + (size_t)r * 128 + 64 + col, res);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// un-patchify epilogue: u [N][256] (n = vox*8+co) -> LayerNorm3D over the 8 channels of each voxel (eps 1e-6)
// -> prob 1x1x1 (8 -> 1) + bias -> logits [D][H][W]
__global__ void unpatch_ln_prob_kernel(const float* __restrict__ u, const float* __restrict__ tail,
                                       float* __restrict__ logits, int D, int H, int W) {
  const int Hp = H / 4, Wp = W / 4;
  size_t total = (size_t)D * H * W;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // thread order follows u's memory order (token-major, voxel-minor) so the 32-byte reads are coalesced
  int vox = (int)(i & 31);
  size_t t = i >> 5;
  int wp = (int)(t % Wp), hp = (int)((t / Wp) % Hp), dp = (int)(t / ((size_t)Wp * Hp));
  int kw = vox & 3, kh = (vox >> 2) & 3, kd = vox >> 4;
  float4 a = ldg4(u + i * 8), b = ldg4(u + i * 8 + 4);
  float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float mean = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) mean += x[c];
  mean *= 0.125f;
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) { float dd = x[c] - mean; var = fmaf(dd, dd, var); }
  var *= 0.125f;
  const float sd = sqrtf(var + 1e-6f);
  float s = __ldg(tail + (U_PB - U_LNW));
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float y = __ldg(tail + c) * __fdiv_rn(x[c] - mean, sd) + __ldg(tail + 8 + c);
    s = fmaf(y, __ldg(tail + 16 + c), s);
  }
  logits[((size_t)(dp * 2 + kd) * H + hp * 4 + kh) * W + wp * 4 + kw] = s;
}


static int run_attention(const float* qkv, float* o, __half* o2, __half* split, int N, float scale_log2e, cudaStream_t s) {
  static int use_v3 = -1;
  if (use_v3 < 0) { const char* e = getenv("MVSF_ATTENTION_V3"); use_v3 = (e && e[0] == '1') ? 1 : 0; }
  cudaEvent_t kt = nullptr;
  if (use_v3) {
    const int Np = (N + 7) & ~7;
    qkv_split_kernel<<<cdiv((long long)Np * 48, 256), 256, 0, s>>>(qkv, split, N, Np, scale_log2e);
    MVSF_LAUNCH_CHECK("qkv_split");
    static bool configured = false;
    if (!configured) {
      MVSF_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fa5::SMEM));
      configured = true;
    }
    kt = ktimer_enabled() ? ktimer_begin("attention_tc", s) : nullptr;
    attention_tc_kernel<<<dim3(cdiv(N, fa5::BM), 4), fa5::THREADS, fa5::SMEM, s>>>(split, o, o2, N, Np);
  } else {
    const int ntiles = cdiv(N, 128);
    qkv_tile_kernel<<<cdiv((long long)ntiles * 128 * 24, 256), 256, 0, s>>>(qkv, split, N, ntiles, scale_log2e);
    MVSF_LAUNCH_CHECK("qkv_tile");
    static bool configured = false;
    if (!configured) {
      MVSF_CUDA_OK(cudaFuncSetAttribute(attention_fa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fa6::SMEM));
      configured = true;
    }
    kt = ktimer_enabled() ? ktimer_begin("attention_tc", s) : nullptr;
    attention_fa_kernel<<<dim3(cdiv(ntiles, 2), 4), fa6::THREADS, fa6::SMEM, s>>>(split, o, o2, N, ntiles);
  }
  if (kt) ktimer_end(kt, s);
  MVSF_LAUNCH_CHECK("attention_tc");
  return MVSF_OK;
}

}  // namespace mvsf

using namespace mvsf;

#ifdef MVSF_FA_TRACE
extern "C" int mvsf_debug_fa_trace(long long* out) {
  MVSF_CUDA_OK(cudaDeviceSynchronize());
  MVSF_CUDA_OK(cudaMemcpyFromSymbol(out, mvsf::g_fa_trace, sizeof(long long) * 64 * 8));
  return MVSF_OK;
}
#endif

extern "C" {

int mvsf_costreg_tr_workspace_bytes(int C, int D, int H, int W, size_t* bytes) {
  MVSF_REQUIRE(bytes && C == 8, "costreg_tr: base channel must be 8");
  MVSF_REQUIRE(D % 2 == 0 && H % 4 == 0 && W % 4 == 0 && D > 0 && H > 0 && W > 0,
               "costreg_tr: D %% 2, H %% 4, W %% 4 must be 0 (down_rate (2,4,4))");
  size_t N = (size_t)(D / 2) * (H / 4) * (W / 4);
  // per token (in floats): big 256 (patches2 / ffn hidden split / un-patchify out), x 64, y 64, x2 64, y2 64, o2 64,
  // qkv 192, attention operand split 192
  *bytes = (N * (256 + 64 + 64 + 64 + 64 + 64 + 192) + (N + 128) * 192) * sizeof(float);
  return MVSF_OK;
}

int mvsf_costreg_tr_forward(float* volume, const float* pos, const float* wts, const void* wts16, size_t n_wts,
                            float* logits, void* workspace, size_t workspace_bytes, int C, int D, int H, int W,
                            int layers, float softmax_scale, mvsf_stream_t stream) {
  MVSF_REQUIRE(volume && wts && wts16 && logits && workspace && layers >= 0, "costreg_tr: bad arguments");
  size_t need = 0;
  int rc = mvsf_costreg_tr_workspace_bytes(C, D, H, W, &need);
  if (rc) return rc;
  if (workspace_bytes < need) return fail(MVSF_ERR_WORKSPACE, "costreg_tr: workspace %zu < %zu bytes", workspace_bytes, need);
  MVSF_REQUIRE(((uintptr_t)workspace & 15) == 0 && ((uintptr_t)wts & 15) == 0 && ((uintptr_t)volume & 15) == 0 &&
                   ((uintptr_t)wts16 & 15) == 0 && (n_wts % 8) == 0,
               "costreg_tr: pointers must be 16-byte aligned");
  MVSF_REQUIRE(n_wts >= (size_t)TR_LAYER0 + (size_t)layers * TR_LAYER + U_PB + 1, "costreg_tr: weight blob too small");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t nvox = (size_t)D * H * W;
  const int N = (int)((size_t)(D / 2) * (H / 4) * (W / 4));
  const __half* wh = reinterpret_cast<const __half*>(wts16);  // fp16 hi parts, same indexing as wts
  const __half* wl = wh + n_wts;                              // fp16 lo parts
  float* big = (float*)workspace;                             // [N][256] floats
  __half* big2 = reinterpret_cast<__half*>(big);             // [N][512] halves: row = [hi(256) | lo(256)]
  float* x = big + (size_t)N * 256;                           // [N][64]
  float* y = x + (size_t)N * 64;                              // [N][64]
  __half* x2 = reinterpret_cast<__half*>(y + (size_t)N * 64);    // [N][128] = [hi(64) | lo(64)]
  __half* y2 = x2 + (size_t)N * 128;
  __half* o2 = y2 + (size_t)N * 128;
  float* qkv = reinterpret_cast<float*>(o2 + (size_t)N * 128);   // [N][192]
  __half* split = reinterpret_cast<__half*>(qkv + (size_t)N * 192);  // [6][4][N][16] fp16

  if (pos) {
    pe3d_add_kernel<<<cdiv((long long)nvox, 256), 256, 0, s>>>(volume, pos, wts + TR_PE, nvox);
    MVSF_LAUNCH_CHECK("pe3d_add");
  }
  patch_gather_kernel<<<cdiv((long long)N * 32, 256), 256, 0, s>>>(volume, big2, D, H, W);
  MVSF_LAUNCH_CHECK("patch_gather");

  TcLinArgs a{};
  a.Ah = big2; a.Al = big2 + 256; a.lda = 512; a.Bh = wh + TR_DOWN_W; a.Bl = wl + TR_DOWN_W; a.ldb = 256;
  a.M = N; a.N = 64; a.K = 256; a.bias = wts + TR_DOWN_B; a.ln_w = wts + TR_DOWN_LNW; a.ln_b = wts + TR_DOWN_LNB; a.ln_eps = 1e-6f;
  a.C = x; a.ldc = 64; a.C2 = x2; a.ldc2 = 128;
  if ((rc = launch_linear_tc(a, LIN_LN, s))) return rc;

  const float scale_log2e = softmax_scale * 1.4426950408889634f;
  for (int l = 0; l < layers; ++l) {
    const size_t lo = (size_t)TR_LAYER0 + (size_t)l * TR_LAYER;
    const float* lw = wts + lo;
    TcLinArgs q{};
    q.Ah = x2; q.Al = x2 + 64; q.lda = 128; q.Bh = wh + lo + L_QKV; q.Bl = wl + lo + L_QKV; q.ldb = 64;
    q.M = N; q.N = 192; q.K = 64; q.C = qkv; q.ldc = 192;
    if ((rc = launch_linear_tc(q, LIN_BIAS, s))) return rc;
    if ((rc = run_attention(qkv, nullptr, o2, split, N, scale_log2e, s))) return rc;
    TcLinArgs p{};
    p.Ah = o2; p.Al = o2 + 64; p.lda = 128; p.Bh = wh + lo + L_PROJ_W; p.Bl = wl + lo + L_PROJ_W; p.ldb = 64;
    p.M = N; p.N = 64; p.K = 64; p.bias = lw + L_PROJ_B; p.res = x; p.ldres = 64; p.gamma = lw + L_G1;
    p.ln_w = lw + L_N1W; p.ln_b = lw + L_N1B; p.ln_eps = 1e-5f; p.C = y; p.ldc = 64; p.C2 = y2; p.ldc2 = 128;
    if ((rc = launch_linear_tc(p, LIN_RES_LN, s))) return rc;
    TcLinArgs f1{};
    f1.Ah = y2; f1.Al = y2 + 64; f1.lda = 128; f1.Bh = wh + lo + L_F1W; f1.Bl = wl + lo + L_F1W; f1.ldb = 64;
    f1.M = N; f1.N = 256; f1.K = 64; f1.bias = lw + L_F1B; f1.C2 = big2; f1.ldc2 = 512;
    if ((rc = launch_linear_tc(f1, LIN_GELU, s))) return rc;
    TcLinArgs f2{};
    f2.Ah = big2; f2.Al = big2 + 256; f2.lda = 512; f2.Bh = wh + lo + L_F2W; f2.Bl = wl + lo + L_F2W; f2.ldb = 256;
    f2.M = N; f2.N = 64; f2.K = 256; f2.bias = lw + L_F2B; f2.res = y; f2.ldres = 64; f2.gamma = lw + L_G2;
    f2.ln_w = lw + L_N2W; f2.ln_b = lw + L_N2B; f2.ln_eps = 1e-5f; f2.C = x; f2.ldc = 64; f2.C2 = x2; f2.ldc2 = 128;
    if ((rc = launch_linear_tc(f2, LIN_RES_LN, s))) return rc;
  }
  const size_t uo = (size_t)TR_LAYER0 + (size_t)layers * TR_LAYER;
  TcLinArgs u{};
  u.Ah = x2; u.Al = x2 + 64; u.lda = 128; u.Bh = wh + uo + U_W; u.Bl = wl + uo + U_W; u.ldb = 64;
  u.M = N; u.N = 256; u.K = 64; u.bias = wts + uo + U_B; u.C = big; u.ldc = 256;
  if ((rc = launch_linear_tc(u, LIN_BIAS, s))) return rc;
  unpatch_ln_prob_kernel<<<cdiv((long long)nvox, 256), 256, 0, s>>>(big, wts + uo + U_LNW, logits, D, H, W);
  MVSF_LAUNCH_CHECK("unpatch_ln_prob");
  return MVSF_OK;
}

/* Softmax attention alone (attention.py:141-170): qkv [N][3][4][16] fp32 -> out [N][64].  workspace >= N*768 bytes. */
int mvsf_attention_forward(const float* qkv, float* out, void* workspace, size_t workspace_bytes, int N,
                           float softmax_scale, mvsf_stream_t stream) {
  MVSF_REQUIRE(qkv && out && workspace && N > 0, "attention_forward: bad arguments");
  if (workspace_bytes < (size_t)(N + 128) * 768) return fail(MVSF_ERR_WORKSPACE, "attention_forward: workspace %zu < %zu bytes", workspace_bytes, (size_t)(N + 128) * 768);
  return run_attention(qkv, out, nullptr, reinterpret_cast<__half*>(workspace), N, softmax_scale * 1.4426950408889634f, (cudaStream_t)stream);
}
}
