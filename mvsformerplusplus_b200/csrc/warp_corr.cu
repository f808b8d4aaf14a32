// W2+W3+W4: homography warp fused with group-wise correlation (models/warping.py:69-109,
// models/cost_volume.py:72-101).  The (V-1,C,D,H,W) warped volume is never written to HBM.
//
// Data layout: feature maps channels-last [V][H][W][C] so one bilinear corner is C contiguous floats.
// Thread mapping: LPP = C/4 lanes cooperate on one reference pixel, each lane owning 4 consecutive channels
// and loading them with one 128-bit LDG per corner; a warp therefore covers 32/LPP consecutive pixels and each
// warp-wide load touches 512 contiguous bytes of the source map when neighbouring pixels map to neighbouring
// source texels (the common case).  Partial dot products are combined across the LPP lanes with xor-shuffles.
//
//   pass A (warp_corr_entropy):   sim[d] = sum_g mean_{c in g} ref[c]*warp[c,d]  ->  softmax_D -> entropy
//   pass B (warp_corr_aggregate): vol[g,d] = sum_v w_v * mean_{c in g} ref*warp_v / (sum_v w_v + 1e-6)
//
// Two passes because the visibility weight w_v is a 7x7-receptive-field CNN of the entropy map
// (cost_volume.py:89-93); recomputing the gather is cheaper than spilling (V-1) x G x D x H x W floats.
#include <float.h>

#include "common.cuh"

namespace mvsf {

constexpr int kMaxGenericD = 512;

template <int LPP>
__device__ __forceinline__ float lanes_sum(float v) {
#pragma unroll
  for (int o = LPP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------- pass A
template <int C, int DT>  // DT > 0: D == DT known at compile time (sims stay in registers); DT == 0: runtime D
__global__ void __launch_bounds__(256)
warp_corr_entropy_kernel(const float* __restrict__ feat, const float* __restrict__ homs,
                         const float* __restrict__ depth, float* __restrict__ entropy, int G, int D, int H, int W) {
  constexpr int LPP = C / 4;
  const int HW = H * W;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int lip = gtid % LPP;
  const int pix = gtid / LPP;
  const int v = blockIdx.y;  // source view index (0-based among sources)
  const bool active = pix < HW;
  const int p = active ? pix : HW - 1;  // keep every lane alive for the shuffles
  const int y = p / W, x = p - y * W;

  const Hom m = load_hom(homs + (size_t)v * 12);
  const float4 r = ldg4(feat + (size_t)p * C + lip * 4);
  const float* __restrict__ src = feat + (size_t)(v + 1) * HW * C + lip * 4;

  const float fx = (float)x, fy = (float)y;
  const float rx = __fadd_rn(fmaf(m.r01, fy, __fmul_rn(m.r00, fx)), m.r02);
  const float ry = __fadd_rn(fmaf(m.r11, fy, __fmul_rn(m.r10, fx)), m.r12);
  const float rz = __fadd_rn(fmaf(m.r21, fy, __fmul_rn(m.r20, fx)), m.r22);
  const float half_w = (float)(W - 1) * 0.5f, half_h = (float)(H - 1) * 0.5f;
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  const float gscale = (float)G / (float)C;  // 1/(C/G): mean over the channels of a group (power of two)

  float sims[DT > 0 ? DT : kMaxGenericD];
  float mx = -FLT_MAX;
  const int Dn = DT > 0 ? DT : D;
#pragma unroll
  for (int d = 0; d < Dn; ++d) {
    float dv = __ldg(depth + (size_t)d * HW + p);
    float ix, iy, z;
    warp_coord(rx, ry, rz, m, dv, half_w, half_h, wm1, hm1, ix, iy, z);
    Tap t = make_tap(ix, iy, W, H, C);
    float4 s = tap4(src, t);
    float part = fmaf(r.w, s.w, fmaf(r.z, s.z, fmaf(r.y, s.y, r.x * s.x)));
    float sim = lanes_sum<LPP>(part) * gscale;
    sims[d] = sim;
    mx = fmaxf(mx, sim);
  }
  float Z = 0.f;
#pragma unroll
  for (int d = 0; d < Dn; ++d) { sims[d] = expf(sims[d] - mx); Z += sims[d]; }
  float ent = 0.f;
#pragma unroll
  for (int d = 0; d < Dn; ++d) {
    float pr = __fdiv_rn(sims[d], Z);
    ent -= pr * logf(pr + 1e-7f);
  }
  if (active && lip == 0) entropy[(size_t)v * HW + p] = ent;
}

// ---------------------------------------------------------------------------------------------- pass B
// CPG = C/G channels per group.  NGL = groups owned by one lane = max(1, 4/CPG).
template <int C, int CPG, int DC>
__global__ void __launch_bounds__(256)
warp_corr_aggregate_kernel(const float* __restrict__ feat, const float* __restrict__ homs,
                           const float* __restrict__ depth, const float* __restrict__ vis,
                           float* __restrict__ volume, int V, int D, int H, int W) {
  constexpr int LPP = C / 4;
  constexpr int G = C / CPG;
  constexpr int NGL = (CPG >= 4) ? 1 : 4 / CPG;
  constexpr int LPG = (CPG >= 4) ? CPG / 4 : 1;  // lanes per group
  const int HW = H * W;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int lip = gtid % LPP;
  const int pix = gtid / LPP;
  const bool active = pix < HW;
  const int p = active ? pix : HW - 1;
  const int y = p / W, x = p - y * W;
  const float4 r = ldg4(feat + (size_t)p * C + lip * 4);
  const float fx = (float)x, fy = (float)y;
  const float half_w = (float)(W - 1) * 0.5f, half_h = (float)(H - 1) * 0.5f;
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  constexpr float inv_cpg = 1.0f / (float)CPG;

  for (int d0 = 0; d0 < D; d0 += DC) {
    float acc[DC][NGL];
#pragma unroll
    for (int i = 0; i < DC; ++i)
#pragma unroll
      for (int g = 0; g < NGL; ++g) acc[i][g] = 0.f;
    float dvs[DC];
#pragma unroll
    for (int i = 0; i < DC; ++i) dvs[i] = (d0 + i < D) ? __ldg(depth + (size_t)(d0 + i) * HW + p) : 1.0f;
    float wsum = 0.f;
    for (int v = 0; v < V - 1; ++v) {
      const Hom m = load_hom(homs + (size_t)v * 12);
      const float rx = __fadd_rn(fmaf(m.r01, fy, __fmul_rn(m.r00, fx)), m.r02);
      const float ry = __fadd_rn(fmaf(m.r11, fy, __fmul_rn(m.r10, fx)), m.r12);
      const float rz = __fadd_rn(fmaf(m.r21, fy, __fmul_rn(m.r20, fx)), m.r22);
      const float* __restrict__ src = feat + (size_t)(v + 1) * HW * C + lip * 4;
      const float w = __ldg(vis + (size_t)v * HW + p);
      wsum = __fadd_rn(wsum, w);
#pragma unroll
      for (int i = 0; i < DC; ++i) {
        float ix, iy, z;
        warp_coord(rx, ry, rz, m, dvs[i], half_w, half_h, wm1, hm1, ix, iy, z);
        Tap t = make_tap(ix, iy, W, H, C);
        float4 s = tap4(src, t);
        if (CPG >= 4) {
          float part = fmaf(r.w, s.w, fmaf(r.z, s.z, fmaf(r.y, s.y, r.x * s.x)));
          part = lanes_sum<LPG>(part) * inv_cpg;
          acc[i][0] = fmaf(part, w, acc[i][0]);
        } else if (CPG == 2) {
          float p0 = fmaf(r.y, s.y, r.x * s.x) * inv_cpg;
          float p1 = fmaf(r.w, s.w, r.z * s.z) * inv_cpg;
          acc[i][0] = fmaf(p0, w, acc[i][0]);
          acc[i][NGL > 1 ? 1 : 0] = fmaf(p1, w, acc[i][NGL > 1 ? 1 : 0]);
        } else {  // CPG == 1: plain product (cost_volume.py:84-85)
          acc[i][0] = fmaf(r.x * s.x, w, acc[i][0]);
          acc[i][NGL > 1 ? 1 : 0] = fmaf(r.y * s.y, w, acc[i][NGL > 1 ? 1 : 0]);
          acc[i][NGL > 2 ? 2 : 0] = fmaf(r.z * s.z, w, acc[i][NGL > 2 ? 2 : 0]);
          acc[i][NGL > 3 ? 3 : 0] = fmaf(r.w * s.w, w, acc[i][NGL > 3 ? 3 : 0]);
        }
      }
    }
    const float den = __fadd_rn(wsum, 1e-6f);
    if (active) {
#pragma unroll
      for (int i = 0; i < DC; ++i) {
        if (d0 + i < D) {
          float* o = volume + ((size_t)(d0 + i) * HW + p) * G;
          if (CPG >= 4) {
            if (lip % LPG == 0) o[lip / LPG] = __fdiv_rn(acc[i][0], den);
          } else if (CPG == 2) {
            float2 v2 = make_float2(__fdiv_rn(acc[i][0], den), __fdiv_rn(acc[i][NGL > 1 ? 1 : 0], den));
            *reinterpret_cast<float2*>(o + lip * 2) = v2;
          } else {
            float4 v4 = make_float4(__fdiv_rn(acc[i][0], den), __fdiv_rn(acc[i][NGL > 1 ? 1 : 0], den),
                                    __fdiv_rn(acc[i][NGL > 2 ? 2 : 0], den), __fdiv_rn(acc[i][NGL > 3 ? 3 : 0], den));
            *reinterpret_cast<float4*>(o + lip * 4) = v4;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- finest seam
// models/warping.py:69-109 as a standalone op (materialises [C][D][H][W]; used for seam parity, not by the hot path)
__global__ void homo_warp_kernel(const float* __restrict__ src, const float* __restrict__ hom,
                                 const float* __restrict__ depth, float* __restrict__ warped,
                                 uint8_t* __restrict__ mask, int C, int D, int H, int W) {
  const int HW = H * W;
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  int d = blockIdx.y;
  if (p >= HW) return;
  int y = p / W, x = p - y * W;
  const Hom m = load_hom(hom);
  const float fx = (float)x, fy = (float)y;
  const float rx = __fadd_rn(fmaf(m.r01, fy, __fmul_rn(m.r00, fx)), m.r02);
  const float ry = __fadd_rn(fmaf(m.r11, fy, __fmul_rn(m.r10, fx)), m.r12);
  const float rz = __fadd_rn(fmaf(m.r21, fy, __fmul_rn(m.r20, fx)), m.r22);
  const float half_w = (float)(W - 1) * 0.5f, half_h = (float)(H - 1) * 0.5f;
  float ix, iy, z;
  warp_coord(rx, ry, rz, m, __ldg(depth + (size_t)d * HW + p), half_w, half_h, (float)(W - 1), (float)(H - 1), ix, iy, z);
  Tap t = make_tap(ix, iy, W, H, C);
  for (int c = 0; c < C; ++c) {
    float a = __ldg(src + t.o00 + c), b = __ldg(src + t.o01 + c), cc = __ldg(src + t.o10 + c), dd = __ldg(src + t.o11 + c);
    warped[((size_t)c * D + d) * HW + p] = fmaf(dd, t.w11, fmaf(cc, t.w10, fmaf(b, t.w01, a * t.w00)));
  }
  if (mask) {
    // warping.py:98-103: |normalised coordinate| > 1 or z <= 0.  ix = (g+1)/2*(W-1)  =>  g>1 <=> ix > W-1 etc.
    bool out = (ix > (float)(W - 1)) || (ix < 0.0f) || (iy > (float)(H - 1)) || (iy < 0.0f) || (z <= 0.0f);
    mask[(size_t)d * HW + p] = out ? 1 : 0;
  }
}

template <int C>
static int launch_entropy(const float* feat, const float* homs, const float* depth, float* entropy, int V, int G, int D,
                          int H, int W, cudaStream_t s) {
  constexpr int LPP = C / 4;
  long long threads = (long long)H * W * LPP;
  dim3 grid(cdiv(threads, 256), V - 1);
  switch (D) {
    case 4: warp_corr_entropy_kernel<C, 4><<<grid, 256, 0, s>>>(feat, homs, depth, entropy, G, D, H, W); break;
    case 8: warp_corr_entropy_kernel<C, 8><<<grid, 256, 0, s>>>(feat, homs, depth, entropy, G, D, H, W); break;
    case 16: warp_corr_entropy_kernel<C, 16><<<grid, 256, 0, s>>>(feat, homs, depth, entropy, G, D, H, W); break;
    case 32: warp_corr_entropy_kernel<C, 32><<<grid, 256, 0, s>>>(feat, homs, depth, entropy, G, D, H, W); break;
    default: warp_corr_entropy_kernel<C, 0><<<grid, 256, 0, s>>>(feat, homs, depth, entropy, G, D, H, W); break;
  }
  return 0;
}

template <int C, int CPG>
static int launch_aggregate(const float* feat, const float* homs, const float* depth, const float* vis, float* volume,
                            int V, int D, int H, int W, cudaStream_t s) {
  constexpr int LPP = C / 4;
  long long threads = (long long)H * W * LPP;
  dim3 grid(cdiv(threads, 256));
  warp_corr_aggregate_kernel<C, CPG, 4><<<grid, 256, 0, s>>>(feat, homs, depth, vis, volume, V, D, H, W);
  return 0;
}

}  // namespace mvsf

using namespace mvsf;

extern "C" {

int mvsf_warp_corr_entropy(const float* feat, const float* homs, const float* depth, float* entropy, int V, int C,
                           int G, int D, int H, int W, mvsf_stream_t stream) {
  MVSF_REQUIRE(feat && homs && depth && entropy, "warp_corr_entropy: null pointer");
  MVSF_REQUIRE(V >= 2 && H > 0 && W > 0 && D >= 1, "warp_corr_entropy: bad shape");
  MVSF_REQUIRE(G <= C, "G must <= C!");  // models/cost_volume.py:87
  MVSF_REQUIRE(C % G == 0 && (C == 8 || C == 16 || C == 32 || C == 64), "warp_corr_entropy: C must be 8/16/32/64, C %% G == 0");
  MVSF_REQUIRE(D <= kMaxGenericD, "warp_corr_entropy: D <= %d", kMaxGenericD);
  cudaStream_t s = (cudaStream_t)stream;
  switch (C) {
    case 8: launch_entropy<8>(feat, homs, depth, entropy, V, G, D, H, W, s); break;
    case 16: launch_entropy<16>(feat, homs, depth, entropy, V, G, D, H, W, s); break;
    case 32: launch_entropy<32>(feat, homs, depth, entropy, V, G, D, H, W, s); break;
    default: launch_entropy<64>(feat, homs, depth, entropy, V, G, D, H, W, s); break;
  }
  MVSF_LAUNCH_CHECK("warp_corr_entropy");
  return MVSF_OK;
}

int mvsf_warp_corr_aggregate(const float* feat, const float* homs, const float* depth, const float* vis,
                             float* volume, int V, int C, int G, int D, int H, int W, mvsf_stream_t stream) {
  MVSF_REQUIRE(feat && homs && depth && vis && volume, "warp_corr_aggregate: null pointer");
  MVSF_REQUIRE(V >= 2 && H > 0 && W > 0 && D >= 1, "warp_corr_aggregate: bad shape");
  MVSF_REQUIRE(G <= C, "G must <= C!");
  MVSF_REQUIRE(G == 8 && (C == 8 || C == 16 || C == 32 || C == 64), "warp_corr_aggregate: G must be 8 and C in 8/16/32/64");
  cudaStream_t s = (cudaStream_t)stream;
  switch (C) {
    case 8: launch_aggregate<8, 1>(feat, homs, depth, vis, volume, V, D, H, W, s); break;
    case 16: launch_aggregate<16, 2>(feat, homs, depth, vis, volume, V, D, H, W, s); break;
    case 32: launch_aggregate<32, 4>(feat, homs, depth, vis, volume, V, D, H, W, s); break;
    default: launch_aggregate<64, 8>(feat, homs, depth, vis, volume, V, D, H, W, s); break;
  }
  MVSF_LAUNCH_CHECK("warp_corr_aggregate");
  return MVSF_OK;
}

int mvsf_homo_warp(const float* src_nhwc, const float* hom, const float* depth, float* warped, uint8_t* mask, int C,
                   int D, int H, int W, mvsf_stream_t stream) {
  MVSF_REQUIRE(src_nhwc && hom && depth && warped && C > 0 && D > 0 && H > 0 && W > 0 && D <= 65535, "homo_warp: bad arguments");
  dim3 grid(cdiv(H * W, 128), D);
  homo_warp_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(src_nhwc, hom, depth, warped, mask, C, D, H, W);
  MVSF_LAUNCH_CHECK("homo_warp");
  return MVSF_OK;
}
}
