// W4 visibility CNN: models/cost_volume.py:37,93
//   ConvBnReLU(1,16) -> ConvBnReLU(16,16) -> ConvBnReLU(16,8) -> Conv1x1(8,1) -> Sigmoid      (fp32, BN folded)
// One fused kernel: a CTA owns a 30x30 output tile, keeps the 16-channel intermediates of all three 3x3
// layers in shared memory (34x34 and 32x32 halo regions) and never writes them to HBM.
// HBM traffic = 4 B in + 4 B out per pixel; the kernel is fp32-FMA bound (7216 FLOP / pixel).
// Register tile for the 16->16 and 16->8 layers: 4 consecutive x-pixels x HALF of the output channels per thread (512
// threads = 16 warps: the kernel runs one CTA per SM, so latency has to be hidden inside the CTA), inputs fetched with one
// LDS.128 + one LDS.64 per (ic, ky), weights with broadcast LDS.128.
#include "common.cuh"

namespace mvsf {

constexpr int VT = 30;         // output tile edge
constexpr int VP = 36;         // shared-memory row pitch (floats), multiple of 4 for 128-bit loads
constexpr int V_IN = VT + 6;   // 36
constexpr int V_A1 = VT + 4;   // 34
constexpr int V_A2 = VT + 2;   // 32
// packed weights (floats): w1[9][16] b1[16] w2[16][9][16] b2[16] w3[16][9][8] b3[8] w4[8] b4[1]
constexpr int OFF_W1 = 0, OFF_B1 = 144, OFF_W2 = 160, OFF_B2 = 160 + 2304, OFF_W3 = OFF_B2 + 16,
              OFF_B3 = OFF_W3 + 1152, OFF_W4 = OFF_B3 + 8, OFF_B4 = OFF_W4 + 8, VIS_WTS = OFF_B4 + 1;

struct VisSmem {
  float a1[16][V_A1][VP];
  float a2[16][V_A2][VP];
  float in[V_IN][VP];
  float w[VIS_WTS + 3];
};

__global__ void __launch_bounds__(512, 1)
vis_cnn_kernel(const float* __restrict__ entropy, const float* __restrict__ wts, float* __restrict__ vis, int H, int W) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  VisSmem& S = *reinterpret_cast<VisSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int n = blockIdx.z;
  const int gx0 = blockIdx.x * VT, gy0 = blockIdx.y * VT;
  const float* __restrict__ E = entropy + (size_t)n * H * W;

  for (int i = tid; i < VIS_WTS; i += 512) S.w[i] = __ldg(wts + i);
  for (int i = tid; i < V_IN * V_IN; i += 512) {
    int yy = i / V_IN, xx = i - yy * V_IN;
    int gy = gy0 - 3 + yy, gx = gx0 - 3 + xx;
    S.in[yy][xx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(E + (size_t)gy * W + gx) : 0.0f;
  }
  __syncthreads();

  // ---- layer 1: 1 -> 16 on the 34x34 region (zero outside the image: that is layer 2's zero padding)
  for (int i = tid; i < V_A1 * V_A1; i += 512) {
    int yy = i / V_A1, xx = i - yy * V_A1;
    int gy = gy0 - 2 + yy, gx = gx0 - 2 + xx;
    bool inside = (gy >= 0 && gy < H && gx >= 0 && gx < W);
    float acc[16];
#pragma unroll
    for (int oc = 0; oc < 16; ++oc) acc[oc] = S.w[OFF_B1 + oc];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float v = S.in[yy + ky][xx + kx];
        const float* wp = &S.w[OFF_W1 + (ky * 3 + kx) * 16];
#pragma unroll
        for (int oc = 0; oc < 16; ++oc) acc[oc] = fmaf(v, wp[oc], acc[oc]);
      }
#pragma unroll
    for (int oc = 0; oc < 16; ++oc) S.a1[oc][yy][xx] = inside ? fmaxf(acc[oc], 0.0f) : 0.0f;
  }
  __syncthreads();

  // ---- layer 2: 16 -> 16 on the 32x32 region; thread = (row, 4-pixel strip) x 8 output channels (oh = channel half)
  {
    const int oh = tid >> 8, t = tid & 255;
    const int r = t >> 3, j = t & 7;
    const int xs = j * 4;
    float acc[4][8];
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
      for (int oc = 0; oc < 8; ++oc) acc[px][oc] = S.w[OFF_B2 + oh * 8 + oc];
    for (int ic = 0; ic < 16; ++ic) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const float* row = &S.a1[ic][r + ky][xs];
        float4 lo = *reinterpret_cast<const float4*>(row);
        float2 hi = *reinterpret_cast<const float2*>(row + 4);
        float in[6] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y};
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4* wp = reinterpret_cast<const float4*>(&S.w[OFF_W2 + (ic * 9 + ky * 3 + kx) * 16 + oh * 8]);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            float4 w4 = wp[q];
#pragma unroll
            for (int px = 0; px < 4; ++px) {
              float v = in[px + kx];
              acc[px][q * 4 + 0] = fmaf(v, w4.x, acc[px][q * 4 + 0]);
              acc[px][q * 4 + 1] = fmaf(v, w4.y, acc[px][q * 4 + 1]);
              acc[px][q * 4 + 2] = fmaf(v, w4.z, acc[px][q * 4 + 2]);
              acc[px][q * 4 + 3] = fmaf(v, w4.w, acc[px][q * 4 + 3]);
            }
          }
        }
      }
    }
    const int gy = gy0 - 1 + r;
    const bool rowin = (gy >= 0 && gy < H);
    bool colin[4];
#pragma unroll
    for (int px = 0; px < 4; ++px) { int gx = gx0 - 1 + xs + px; colin[px] = rowin && gx >= 0 && gx < W; }
#pragma unroll
    for (int oc = 0; oc < 8; ++oc) {
      float4 o;
      o.x = colin[0] ? fmaxf(acc[0][oc], 0.f) : 0.f;
      o.y = colin[1] ? fmaxf(acc[1][oc], 0.f) : 0.f;
      o.z = colin[2] ? fmaxf(acc[2][oc], 0.f) : 0.f;
      o.w = colin[3] ? fmaxf(acc[3][oc], 0.f) : 0.f;
      *reinterpret_cast<float4*>(&S.a2[oh * 8 + oc][r][xs]) = o;
    }
  }
  __syncthreads();

  // ---- layer 3 (16 -> 8) + layer 4 (1x1, 8 -> 1) + sigmoid on the 30x30 tile; thread = (row, strip) x 4 output
  //      channels; the two channel halves meet through shared memory (the input tile is dead by now)
  {
    const int oh = tid >> 8, t = tid & 255;
    const int r = t >> 3, j = t & 7;
    const int xs = j * 4;
    float part[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < VT) {
      float acc[4][4];
#pragma unroll
      for (int px = 0; px < 4; ++px)
#pragma unroll
        for (int oc = 0; oc < 4; ++oc) acc[px][oc] = S.w[OFF_B3 + oh * 4 + oc];
      for (int ic = 0; ic < 16; ++ic) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const float* row = &S.a2[ic][r + ky][xs];
          float4 lo = *reinterpret_cast<const float4*>(row);
          float2 hi = *reinterpret_cast<const float2*>(row + 4);
          float in[6] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y};
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float4 w4 = *reinterpret_cast<const float4*>(&S.w[OFF_W3 + (ic * 9 + ky * 3 + kx) * 8 + oh * 4]);
#pragma unroll
            for (int px = 0; px < 4; ++px) {
              float v = in[px + kx];
              acc[px][0] = fmaf(v, w4.x, acc[px][0]);
              acc[px][1] = fmaf(v, w4.y, acc[px][1]);
              acc[px][2] = fmaf(v, w4.z, acc[px][2]);
              acc[px][3] = fmaf(v, w4.w, acc[px][3]);
            }
          }
        }
      }
      // layer 4 in the reference's channel order: s = b4 + sum_{oc = 0..7} relu(x_oc) * w_oc  (half 0 first, then half 1)
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        float sacc = oh == 0 ? S.w[OFF_B4] : 0.f;
#pragma unroll
        for (int oc = 0; oc < 4; ++oc) sacc = fmaf(fmaxf(acc[px][oc], 0.f), S.w[OFF_W4 + oh * 4 + oc], sacc);
        part[px] = sacc;
      }
      if (oh == 1) *reinterpret_cast<float4*>(&S.in[r][xs]) = make_float4(part[0], part[1], part[2], part[3]);
    }
    __syncthreads();
    if (oh == 0 && r < VT) {
      const int gy = gy0 + r;
      if (gy < H) {
        const float4 other = *reinterpret_cast<const float4*>(&S.in[r][xs]);
        const float o4[4] = {other.x, other.y, other.z, other.w};
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          int lx = xs + px, gx = gx0 + lx;
          if (lx < VT && gx < W) {
            const float sv = part[px] + o4[px];
            vis[((size_t)n * H + gy) * W + gx] = __fdiv_rn(1.0f, 1.0f + expf(-sv));
          }
        }
      }
    }
  }
}

}  // namespace mvsf

using namespace mvsf;

extern "C" int mvsf_vis_cnn(const float* entropy, const float* wts, float* vis, int N, int H, int W,
                            mvsf_stream_t stream) {
  MVSF_REQUIRE(entropy && wts && vis && N > 0 && N <= 65535 && H > 0 && W > 0, "vis_cnn: bad arguments");
  static_assert(OFF_W2 % 4 == 0 && OFF_W3 % 4 == 0, "weight blocks must be 16-byte aligned");
  size_t smem = sizeof(VisSmem);
  static bool configured = false;
  if (!configured) {
    MVSF_CUDA_OK(cudaFuncSetAttribute(vis_cnn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid(cdiv(W, VT), cdiv(H, VT), N);
  vis_cnn_kernel<<<grid, 512, smem, (cudaStream_t)stream>>>(entropy, wts, vis, H, W);
  MVSF_LAUNCH_CHECK("vis_cnn");
  return MVSF_OK;
}
