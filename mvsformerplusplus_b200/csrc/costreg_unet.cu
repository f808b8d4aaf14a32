// R2-R4: 3-D conv U-Net cost regularisers (models/module.py:367-408 CostRegNet, :453-504 CostRegNet3D).
// Activations are NDHWC fp32; BatchNorm (eval) is folded into the packed weights [27][Cin][Cout] + bias[Cout];
// ReLU, the U-Net skip additions (added AFTER the ReLU: `conv4 + self.conv7(x)`, module.py:403-405) and, for
// CostRegNet3D, the final 1x1x1 `prob` conv are fused into the producing kernel's epilogue.
//
// conv3d_k3_kernel  : out tile 32(w) x TH(h) x 1(d); 128 threads; a thread owns 4 voxels (along h) x CT<=16 output
//                     channels; input channels are streamed 4 at a time through shared memory (one float4 per
//                     voxel, so a warp reads 512 contiguous bytes per tap) with the matching weight slab
//                     [27][4][Cout]; weights are warp-broadcast LDS.128.
// deconv3d_k3_kernel: gather form of ConvTranspose3d(k=3, s=(SD,2,2), p=1, output_padding=s-1): a thread owns 4 input
//                     cells (along h) and produces their 2x2 output quads for 8 output channels; every (kh,kw) tap
//                     contributes to exactly one of the 4 output parities.
#include "common.cuh"
#include "conv3d_tc.cuh"

namespace mvsf {

// ------------------------------------------------------------------------------------------------------ conv
template <int CIN, int COUT, int SD, int SH, int SW>
struct ConvCfg {
  static constexpr int CT = COUT < 16 ? COUT : 16;
  static constexpr int NCG = COUT / CT;
  static constexpr int WARPS_H = 4 / NCG;
  static constexpr int TH = 4 * WARPS_H;
  static constexpr int TW = 32;
  static constexpr int IH_T = (TH - 1) * SH + 3;
  static constexpr int IW_T = (TW - 1) * SW + 3;
  static constexpr int IN_F4 = 3 * IH_T * IW_T;           // float4 elements
  static constexpr int WT_F = 27 * 4 * COUT;              // floats
  static constexpr size_t SMEM = (size_t)IN_F4 * 16 + (size_t)WT_F * 4;
};

template <int CIN, int COUT, int SD, int SH, int SW>
__global__ void __launch_bounds__(128)
conv3d_k3_kernel(const float* __restrict__ in, const float* __restrict__ wts, const float* __restrict__ bias,
                 float* __restrict__ out, int ID, int IH, int IW, int OD, int OH, int OW) {
  using Cfg = ConvCfg<CIN, COUT, SD, SH, SW>;
  constexpr int CT = Cfg::CT, NCG = Cfg::NCG, TH = Cfg::TH, IH_T = Cfg::IH_T, IW_T = Cfg::IW_T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* in_s = reinterpret_cast<float4*>(smem_raw);
  float* wt_s = reinterpret_cast<float*>(smem_raw + (size_t)Cfg::IN_F4 * 16);

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int cg = wid % NCG, hg = wid / NCG;
  const int od = blockIdx.z, oh0 = blockIdx.y * TH, ow0 = blockIdx.x * 32;
  const int id0 = od * SD - 1, ih0 = oh0 * SH - 1, iw0 = ow0 * SW - 1;

  float acc[4][CT];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < CT; ++c) acc[j][c] = 0.f;

  for (int pc = 0; pc < CIN / 4; ++pc) {
    __syncthreads();
    for (int i = tid; i < Cfg::IN_F4; i += 128) {
      int kd = i / (IH_T * IW_T), rem = i - kd * (IH_T * IW_T);
      int iy = rem / IW_T, ix = rem - iy * IW_T;
      int id = id0 + kd, ih = ih0 + iy, iw = iw0 + ix;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (id >= 0 && id < ID && ih >= 0 && ih < IH && iw >= 0 && iw < IW)
        v = ldg4(in + (((size_t)id * IH + ih) * IW + iw) * CIN + pc * 4);
      in_s[i] = v;
    }
    for (int i = tid; i < 27 * 4 * COUT / 4; i += 128) {
      int e = i * 4;
      int tap = e / (4 * COUT), rem = e - tap * (4 * COUT);
      int ci = rem / COUT, co = rem - ci * COUT;
      *reinterpret_cast<float4*>(wt_s + e) = ldg4(wts + ((size_t)tap * CIN + pc * 4 + ci) * COUT + co);
    }
    __syncthreads();
#pragma unroll 1
    for (int kdh = 0; kdh < 9; ++kdh) {
      const int kd = kdh / 3, kh = kdh - kd * 3;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        float4 a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          a[j] = in_s[(kd * IH_T + (hg * 4 + j) * SH + kh) * IW_T + lane * SW + kw];
        const float* wp = wt_s + (size_t)((kdh * 3 + kw) * 4) * COUT + cg * CT;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
#pragma unroll
          for (int q = 0; q < CT / 4; ++q) {
            float4 w4 = *reinterpret_cast<const float4*>(wp + ci * COUT + q * 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float v = (ci == 0) ? a[j].x : (ci == 1) ? a[j].y : (ci == 2) ? a[j].z : a[j].w;
              acc[j][q * 4 + 0] = fmaf(v, w4.x, acc[j][q * 4 + 0]);
              acc[j][q * 4 + 1] = fmaf(v, w4.y, acc[j][q * 4 + 1]);
              acc[j][q * 4 + 2] = fmaf(v, w4.z, acc[j][q * 4 + 2]);
              acc[j][q * 4 + 3] = fmaf(v, w4.w, acc[j][q * 4 + 3]);
            }
          }
        }
      }
    }
  }
  const int ow = ow0 + lane;
  if (ow < OW) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int oh = oh0 + hg * 4 + j;
      if (oh < OH) {
        float* o = out + (((size_t)od * OH + oh) * OW + ow) * COUT + cg * CT;
#pragma unroll
        for (int q = 0; q < CT / 4; ++q) {
          float4 b4 = ldg4(bias + cg * CT + q * 4);
          float4 r;
          r.x = fmaxf(acc[j][q * 4 + 0] + b4.x, 0.f);
          r.y = fmaxf(acc[j][q * 4 + 1] + b4.y, 0.f);
          r.z = fmaxf(acc[j][q * 4 + 2] + b4.z, 0.f);
          r.w = fmaxf(acc[j][q * 4 + 3] + b4.w, 0.f);
          *reinterpret_cast<float4*>(o + q * 4) = r;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------- deconv
template <int CIN, int COUT>
struct DeconvCfg {
  static constexpr int CT = 8;
  static constexpr int NCG = COUT / CT;
  static constexpr int WARPS_H = 4 / NCG;
  static constexpr int TH = 4 * WARPS_H;       // input cells along h per CTA
  static constexpr int IH_T = TH + 1, IW_T = 33;
  static constexpr int IN_F4 = 3 * IH_T * IW_T;
  static constexpr int WT_F = 27 * 4 * COUT;
  static constexpr size_t SMEM = (size_t)IN_F4 * 16 + (size_t)WT_F * 4;
};

// out[od][2ih+ph][2iw+pw][co] = skip + relu(bias + sum ...) ; FUSE_PROB: logits = pb + sum_co pw[co]*that (COUT == 8)
template <int CIN, int COUT, int SD, bool FUSE_PROB>
__global__ void __launch_bounds__(128)
deconv3d_k3_kernel(const float* __restrict__ in, const float* __restrict__ wts, const float* __restrict__ bias,
                   const float* __restrict__ skip, float* __restrict__ out, const float* __restrict__ probw,
                   int ID, int IH, int IW) {
  using Cfg = DeconvCfg<CIN, COUT>;
  constexpr int CT = 8, NCG = Cfg::NCG, TH = Cfg::TH, IH_T = Cfg::IH_T, IW_T = Cfg::IW_T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* in_s = reinterpret_cast<float4*>(smem_raw);
  float* wt_s = reinterpret_cast<float*>(smem_raw + (size_t)Cfg::IN_F4 * 16);

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int cg = wid % NCG, hg = wid / NCG;
  const int od = blockIdx.z, ih0 = blockIdx.y * TH, iw0 = blockIdx.x * 32;
  const int OD = ID * SD, OH = IH * 2, OW = IW * 2;
  (void)OD;

  // depth taps of this output slice: od = id*SD - 1 + kd
  int ntap = 0, tkd[3], tid_[3];
  if (SD == 1) {
    for (int kd = 0; kd < 3; ++kd) {
      int id = od + 1 - kd;
      if (id >= 0 && id < ID) { tkd[ntap] = kd; tid_[ntap] = id; ++ntap; }
    }
  } else {
    if ((od & 1) == 0) { tkd[0] = 1; tid_[0] = od >> 1; ntap = 1; }
    else {
      int id = (od + 1) >> 1;
      if (id < ID) { tkd[ntap] = 0; tid_[ntap] = id; ++ntap; }
      id = (od - 1) >> 1;
      if (id >= 0) { tkd[ntap] = 2; tid_[ntap] = id; ++ntap; }
    }
  }

  float acc[4][4][CT];  // [class ph*2+pw][cell j][co]
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < CT; ++k) acc[c][j][k] = 0.f;

  for (int pc = 0; pc < CIN / 4; ++pc) {
    __syncthreads();
    for (int i = tid; i < ntap * IH_T * IW_T; i += 128) {
      int t = i / (IH_T * IW_T), rem = i - t * (IH_T * IW_T);
      int iy = rem / IW_T, ix = rem - iy * IW_T;
      int ih = ih0 + iy, iw = iw0 + ix;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ih < IH && iw < IW) v = ldg4(in + (((size_t)tid_[t] * IH + ih) * IW + iw) * CIN + pc * 4);
      in_s[i] = v;
    }
    for (int i = tid; i < 27 * 4 * COUT / 4; i += 128) {
      int e = i * 4;
      int tap = e / (4 * COUT), rem = e - tap * (4 * COUT);
      int ci = rem / COUT, co = rem - ci * COUT;
      *reinterpret_cast<float4*>(wt_s + e) = ldg4(wts + ((size_t)tap * CIN + pc * 4 + ci) * COUT + co);
    }
    __syncthreads();
    for (int t = 0; t < ntap; ++t) {
      float4 a[5][2];
#pragma unroll
      for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c) a[r][c] = in_s[(t * IH_T + hg * 4 + r) * IW_T + lane + c];
      const float* wbase = wt_s + (size_t)(tkd[t] * 9 * 4) * COUT + cg * CT;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          constexpr int dummy = 0; (void)dummy;
          const int ph = (kh == 1) ? 0 : 1, dih = (kh == 0) ? 1 : 0;
          const int pw = (kw == 1) ? 0 : 1, diw = (kw == 0) ? 1 : 0;
          const int cls = ph * 2 + pw;
          const float* wp = wbase + (size_t)((kh * 3 + kw) * 4) * COUT;
#pragma unroll
          for (int ci = 0; ci < 4; ++ci) {
            float4 w0 = *reinterpret_cast<const float4*>(wp + ci * COUT);
            float4 w1 = *reinterpret_cast<const float4*>(wp + ci * COUT + 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 av = a[j + dih][diw];
              float v = (ci == 0) ? av.x : (ci == 1) ? av.y : (ci == 2) ? av.z : av.w;
              acc[cls][j][0] = fmaf(v, w0.x, acc[cls][j][0]);
              acc[cls][j][1] = fmaf(v, w0.y, acc[cls][j][1]);
              acc[cls][j][2] = fmaf(v, w0.z, acc[cls][j][2]);
              acc[cls][j][3] = fmaf(v, w0.w, acc[cls][j][3]);
              acc[cls][j][4] = fmaf(v, w1.x, acc[cls][j][4]);
              acc[cls][j][5] = fmaf(v, w1.y, acc[cls][j][5]);
              acc[cls][j][6] = fmaf(v, w1.z, acc[cls][j][6]);
              acc[cls][j][7] = fmaf(v, w1.w, acc[cls][j][7]);
            }
          }
        }
      }
    }
  }

  const int iw = iw0 + lane;
  if (iw >= IW) return;
  const float4 b0 = ldg4(bias + cg * CT), b1 = ldg4(bias + cg * CT + 4);
  float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
  float pb = 0.f;
  if (FUSE_PROB) { p0 = ldg4(probw); p1 = ldg4(probw + 4); pb = __ldg(probw + 8); }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int ih = ih0 + hg * 4 + j;
    if (ih >= IH) continue;
#pragma unroll
    for (int cls = 0; cls < 4; ++cls) {
      int oh = 2 * ih + (cls >> 1), ow = 2 * iw + (cls & 1);
      size_t vox = ((size_t)od * OH + oh) * OW + ow;
      const float* sk = skip + vox * COUT + cg * CT;
      float4 s0 = ldg4(sk), s1 = ldg4(sk + 4);
      float4 r0, r1;
      r0.x = s0.x + fmaxf(acc[cls][j][0] + b0.x, 0.f);
      r0.y = s0.y + fmaxf(acc[cls][j][1] + b0.y, 0.f);
      r0.z = s0.z + fmaxf(acc[cls][j][2] + b0.z, 0.f);
      r0.w = s0.w + fmaxf(acc[cls][j][3] + b0.w, 0.f);
      r1.x = s1.x + fmaxf(acc[cls][j][4] + b1.x, 0.f);
      r1.y = s1.y + fmaxf(acc[cls][j][5] + b1.y, 0.f);
      r1.z = s1.z + fmaxf(acc[cls][j][6] + b1.z, 0.f);
      r1.w = s1.w + fmaxf(acc[cls][j][7] + b1.w, 0.f);
      if (FUSE_PROB) {
        float l = pb;
        l = fmaf(r0.x, p0.x, l); l = fmaf(r0.y, p0.y, l); l = fmaf(r0.z, p0.z, l); l = fmaf(r0.w, p0.w, l);
        l = fmaf(r1.x, p1.x, l); l = fmaf(r1.y, p1.y, l); l = fmaf(r1.z, p1.z, l); l = fmaf(r1.w, p1.w, l);
        out[vox] = l;
      } else {
        float* o = out + vox * COUT + cg * CT;
        *reinterpret_cast<float4*>(o) = r0;
        *reinterpret_cast<float4*>(o + 4) = r1;
      }
    }
  }
}

// CostRegNet `prob`: Conv3d(8,1,3,padding=1,bias=False) (module.py:392).  in [D][H][W][8], w [27][8] -> out [D][H][W]
__global__ void prob3_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out, int D,
                             int H, int W) {
  __shared__ float ws[27 * 8];
  for (int i = threadIdx.x; i < 27 * 8; i += blockDim.x) ws[i] = __ldg(w + i);
  __syncthreads();
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, d = blockIdx.z;
  if (x >= W) return;
  float acc = 0.f;
  for (int kd = 0; kd < 3; ++kd) {
    int zd = d + kd - 1;
    if (zd < 0 || zd >= D) continue;
    for (int kh = 0; kh < 3; ++kh) {
      int yh = y + kh - 1;
      if (yh < 0 || yh >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int xw = x + kw - 1;
        if (xw < 0 || xw >= W) continue;
        const float* p = in + (((size_t)zd * H + yh) * W + xw) * 8;
        float4 a = ldg4(p), b = ldg4(p + 4);
        const float* wp = ws + ((kd * 3 + kh) * 3 + kw) * 8;
        acc = fmaf(a.x, wp[0], acc); acc = fmaf(a.y, wp[1], acc); acc = fmaf(a.z, wp[2], acc); acc = fmaf(a.w, wp[3], acc);
        acc = fmaf(b.x, wp[4], acc); acc = fmaf(b.y, wp[5], acc); acc = fmaf(b.z, wp[6], acc); acc = fmaf(b.w, wp[7], acc);
      }
    }
  }
  out[((size_t)d * H + y) * W + x] = acc;
}

// ---------------------------------------------------------------------------------------------------- host
template <int CIN, int COUT, int SD, int SH, int SW>
static int run_conv(const float* in, const float* w, float* out, int ID, int IH, int IW, cudaStream_t s) {
  using Cfg = ConvCfg<CIN, COUT, SD, SH, SW>;
  int OD = (ID - 1) / SD + 1, OH = (IH - 1) / SH + 1, OW = (IW - 1) / SW + 1;
  auto kern = conv3d_k3_kernel<CIN, COUT, SD, SH, SW>;
  static bool configured = false;
  if (!configured) {
    MVSF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    configured = true;
  }
  dim3 grid(cdiv(OW, 32), cdiv(OH, Cfg::TH), OD);
  MVSF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv3d: volume too large");
  kern<<<grid, 128, Cfg::SMEM, s>>>(in, w, w + (size_t)27 * CIN * COUT, out, ID, IH, IW, OD, OH, OW);
  MVSF_LAUNCH_CHECK("conv3d_k3");
  return MVSF_OK;
}

template <int CIN, int COUT, int SD, bool FUSE>
static int run_deconv(const float* in, const float* w, const float* skip, float* out, const float* probw, int ID,
                      int IH, int IW, cudaStream_t s) {
  using Cfg = DeconvCfg<CIN, COUT>;
  auto kern = deconv3d_k3_kernel<CIN, COUT, SD, FUSE>;
  static bool configured = false;
  if (!configured) {
    MVSF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    configured = true;
  }
  dim3 grid(cdiv(IW, 32), cdiv(IH, Cfg::TH), ID * SD);
  MVSF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "deconv3d: volume too large");
  kern<<<grid, 128, Cfg::SMEM, s>>>(in, w, w + (size_t)27 * CIN * COUT, skip, out, probw, ID, IH, IW);
  MVSF_LAUNCH_CHECK("deconv3d_k3");
  return MVSF_OK;
}

static size_t layer_floats(int cin, int cout) { return (size_t)27 * cin * cout + cout; }

template <int SD>
static int unet_forward(int kind, const float* vol, const float* wts, float* logits, float* ws, int D, int H, int W,
                        cudaStream_t s) {
  const int D1 = (D - 1) / SD + 1, H1 = H / 2, W1 = W / 2;
  const int D2 = (D1 - 1) / SD + 1, H2 = H1 / 2, W2 = W1 / 2;
  const int D3 = (D2 - 1) / SD + 1, H3 = H2 / 2, W3 = W2 / 2;
  const size_t n1 = (size_t)D1 * H1 * W1 * 16, n2 = (size_t)D2 * H2 * W2 * 32, n3 = (size_t)D3 * H3 * W3 * 64;
  float* t1 = ws;            float* c2 = t1 + n1;
  float* t3 = c2 + n1;       float* c4 = t3 + n2;
  float* t5 = c4 + n2;       float* c6 = t5 + n3;
  float* x11 = c6 + n3;      // kind 0 only: [D][H][W][8]
  const float* w1 = wts;
  const float* w2 = w1 + layer_floats(8, 16);
  const float* w3 = w2 + layer_floats(16, 16);
  const float* w4 = w3 + layer_floats(16, 32);
  const float* w5 = w4 + layer_floats(32, 32);
  const float* w6 = w5 + layer_floats(32, 64);
  const float* w7 = w6 + layer_floats(64, 64);
  const float* w9 = w7 + layer_floats(64, 32);
  const float* w11 = w9 + layer_floats(32, 16);
  const float* wp = w11 + layer_floats(16, 8);
  int rc;
  if ((rc = run_conv<8, 16, SD, 2, 2>(vol, w1, t1, D, H, W, s))) return rc;
  if ((rc = run_conv<16, 16, 1, 1, 1>(t1, w2, c2, D1, H1, W1, s))) return rc;
  if ((rc = run_conv<16, 32, SD, 2, 2>(c2, w3, t3, D1, H1, W1, s))) return rc;
  if ((rc = run_conv<32, 32, 1, 1, 1>(t3, w4, c4, D2, H2, W2, s))) return rc;
  if ((rc = run_conv<32, 64, SD, 2, 2>(c4, w5, t5, D2, H2, W2, s))) return rc;
  if ((rc = run_conv<64, 64, 1, 1, 1>(t5, w6, c6, D3, H3, W3, s))) return rc;
  // x = conv4 + conv7(x) -> t3 ; x = conv2 + conv9(x) -> t1 ; x = conv0 + conv11(x)
  if ((rc = run_deconv<64, 32, SD, false>(c6, w7, c4, t3, nullptr, D3, H3, W3, s))) return rc;
  if ((rc = run_deconv<32, 16, SD, false>(t3, w9, c2, t1, nullptr, D2, H2, W2, s))) return rc;
  if (kind == 1) {
    if ((rc = run_deconv<16, 8, SD, true>(t1, w11, vol, logits, wp, D1, H1, W1, s))) return rc;
  } else {
    if ((rc = run_deconv<16, 8, SD, false>(t1, w11, vol, x11, nullptr, D1, H1, W1, s))) return rc;
    dim3 grid(cdiv(W, 128), H, D);
    prob3_kernel<<<grid, 128, 0, s>>>(x11, wp, logits, D, H, W);
    MVSF_LAUNCH_CHECK("prob3");
  }
  return MVSF_OK;
}

// ------------------------------------------------------------------------------------ tensor-core path (conv3d_tc.cu)
static const int kLayerCh[9][2] = {{8, 16}, {16, 16}, {16, 32}, {32, 32}, {32, 64}, {64, 64}, {64, 32}, {32, 16}, {16, 8}};
static const int kLayerMode[9] = {CONV_S2, CONV_S1, CONV_S2, CONV_S1, CONV_S2, CONV_S1, DECONV_S2, DECONV_S2, DECONV_S2};

static size_t tc_total_halves() {
  size_t n = 0;
  for (int l = 0; l < 9; ++l) n += conv3d_tc_packed_halves(kLayerMode[l], kLayerCh[l][0], kLayerCh[l][1]);
  return n;
}

static int unet_forward_tc(int kind, const float* vol, const float* wts, const __half* wtc, float* logits, void* ws, int D,
                           int H, int W, cudaStream_t s) {
  const int SD = kind == 0 ? 2 : 1;
  const int D1 = (D - 1) / SD + 1, H1 = H / 2, W1 = W / 2;
  const int D2 = (D1 - 1) / SD + 1, H2 = H1 / 2, W2 = W1 / 2;
  const int D3 = (D2 - 1) / SD + 1, H3 = H2 / 2, W3 = W2 / 2;
  const size_t n0 = (size_t)D * H * W * 8;
  const size_t n1 = (size_t)D1 * H1 * W1 * 16, n2 = (size_t)D2 * H2 * W2 * 32, n3 = (size_t)D3 * H3 * W3 * 64;
  // every activation buffer is [hi (n halves) | lo (n halves)]
  __half* v0 = reinterpret_cast<__half*>(ws);
  __half* t1 = v0 + 2 * n0;   __half* c2 = t1 + 2 * n1;
  __half* t3 = c2 + 2 * n1;   __half* c4 = t3 + 2 * n2;
  __half* t5 = c4 + 2 * n2;   __half* c6 = t5 + 2 * n3;
  float* x11 = reinterpret_cast<float*>(c6 + 2 * n3);   // kind 0 only: [D][H][W][8] fp32
  const float* w32[9];
  const __half* w16[9];
  {
    const float* p = wts;
    const __half* q = wtc;
    for (int l = 0; l < 9; ++l) {
      w32[l] = p; w16[l] = q;
      p += layer_floats(kLayerCh[l][0], kLayerCh[l][1]);
      q += conv3d_tc_packed_halves(kLayerMode[l], kLayerCh[l][0], kLayerCh[l][1]);
    }
  }
  const float* wp = w32[8] + layer_floats(16, 8);
  int rc;
  if ((rc = launch_split_vec8(vol, v0, v0 + n0, n0, s))) return rc;
  auto conv = [&](int l, const __half* in, size_t nin, __half* out, size_t nout, const __half* skip, size_t nskip,
                  int ID, int IH, int IW) {
    ConvTcArgs a{};
    a.in_hi = in; a.in_lo = in + nin;
    a.wtc = w16[l]; a.bias = w32[l] + (size_t)27 * kLayerCh[l][0] * kLayerCh[l][1];
    a.skip_hi = skip; a.skip_lo = skip ? skip + nskip : nullptr;
    a.out_hi = out; a.out_lo = out + nout;
    a.CIN = kLayerCh[l][0]; a.COUT = kLayerCh[l][1]; a.SD = SD; a.ID = ID; a.IH = IH; a.IW = IW;
    a.KG = conv3d_tc_kg(kLayerMode[l], a.CIN);
    a.col = conv3d_tc_col(kLayerMode[l], SD, a.COUT);
    return launch_conv3d_tc(a, kLayerMode[l], OUT_SPLIT, s);
  };
  if ((rc = conv(0, v0, n0, t1, n1, nullptr, 0, D, H, W))) return rc;
  if ((rc = conv(1, t1, n1, c2, n1, nullptr, 0, D1, H1, W1))) return rc;
  if ((rc = conv(2, c2, n1, t3, n2, nullptr, 0, D1, H1, W1))) return rc;
  if ((rc = conv(3, t3, n2, c4, n2, nullptr, 0, D2, H2, W2))) return rc;
  if ((rc = conv(4, c4, n2, t5, n3, nullptr, 0, D2, H2, W2))) return rc;
  if ((rc = conv(5, t5, n3, c6, n3, nullptr, 0, D3, H3, W3))) return rc;
  // x = conv4 + conv7(x) -> t3 ; x = conv2 + conv9(x) -> t1 ; x = conv0 + conv11(x)
  if ((rc = conv(6, c6, n3, t3, n2, c4, n2, D3, H3, W3))) return rc;
  if ((rc = conv(7, t3, n2, t1, n1, c2, n1, D2, H2, W2))) return rc;
  {
    ConvTcArgs a{};
    a.in_hi = t1; a.in_lo = t1 + n1;
    a.wtc = w16[8]; a.bias = w32[8] + (size_t)27 * 16 * 8;
    a.skip32 = vol;
    a.CIN = 16; a.COUT = 8; a.SD = SD; a.ID = D1; a.IH = H1; a.IW = W1; a.KG = conv3d_tc_kg(DECONV_S2, 16);
    if (kind == 1) {
      a.out32 = logits; a.probw = wp;
      if ((rc = launch_conv3d_tc(a, DECONV_S2, OUT_PROB, s))) return rc;
    } else {
      a.out32 = x11;
      if ((rc = launch_conv3d_tc(a, DECONV_S2, OUT_F32, s))) return rc;
      dim3 grid(cdiv(W, 128), H, D);
      prob3_kernel<<<grid, 128, 0, s>>>(x11, wp, logits, D, H, W);
      MVSF_LAUNCH_CHECK("prob3");
    }
  }
  return MVSF_OK;
}

}  // namespace mvsf

using namespace mvsf;

extern "C" {

int mvsf_costreg_unet_workspace_bytes(int kind, int C, int D, int H, int W, size_t* bytes) {
  MVSF_REQUIRE(bytes && (kind == 0 || kind == 1) && C == 8, "costreg_unet: kind in {0,1}, C == 8");
  MVSF_REQUIRE(H % 8 == 0 && W % 8 == 0 && D >= 1 && (kind == 1 || D % 8 == 0),
               "costreg_unet: H, W (and D for CostRegNet) must be multiples of 8");
  const int SD = kind == 0 ? 2 : 1;
  size_t D1 = (D - 1) / SD + 1, D2 = (D1 - 1) / SD + 1, D3 = (D2 - 1) / SD + 1;
  size_t n1 = D1 * (H / 2) * (W / 2) * 16, n2 = D2 * (H / 4) * (W / 4) * 32, n3 = D3 * (H / 8) * (W / 8) * 64;
  // fp16 hi|lo activation buffers (4 bytes per element, like fp32): input split + two per level; kind 0: fp32 x11
  size_t n = (size_t)D * H * W * 8 + 2 * (n1 + n2 + n3) + (kind == 0 ? (size_t)D * H * W * 8 : 0);
  *bytes = n * sizeof(float);
  return MVSF_OK;
}

int mvsf_costreg_unet_tc_bytes(size_t* bytes) {
  MVSF_REQUIRE(bytes, "costreg_unet_tc_bytes: null pointer");
  *bytes = tc_total_halves() * sizeof(__half);
  return MVSF_OK;
}

int mvsf_costreg_unet_pack_tc(int kind, const float* wts, void* wts_tc, size_t wts_tc_bytes, mvsf_stream_t stream) {
  MVSF_REQUIRE(wts && wts_tc && ((uintptr_t)wts_tc & 15) == 0 && (kind == 0 || kind == 1), "costreg_unet_pack_tc: null or unaligned pointer, or bad kind");
  if (wts_tc_bytes < tc_total_halves() * sizeof(__half))
    return fail(MVSF_ERR_WORKSPACE, "costreg_unet_pack_tc: buffer %zu < %zu bytes", wts_tc_bytes, tc_total_halves() * sizeof(__half));
  const float* p = wts;
  __half* q = reinterpret_cast<__half*>(wts_tc);
  for (int l = 0; l < 9; ++l) {
    int rc = conv3d_tc_pack(p, q, kLayerMode[l], kind == 0 ? 2 : 1, kLayerCh[l][0], kLayerCh[l][1], (cudaStream_t)stream);
    if (rc) return rc;
    p += layer_floats(kLayerCh[l][0], kLayerCh[l][1]);
    q += conv3d_tc_packed_halves(kLayerMode[l], kLayerCh[l][0], kLayerCh[l][1]);
  }
  return MVSF_OK;
}

int mvsf_costreg_unet_forward(int kind, const float* volume, const float* wts, const void* wts_tc, float* logits,
                              void* workspace, size_t workspace_bytes, int C, int D, int H, int W, mvsf_stream_t stream) {
  MVSF_REQUIRE(volume && wts && logits && workspace, "costreg_unet: null pointer");
  size_t need = 0;
  int rc = mvsf_costreg_unet_workspace_bytes(kind, C, D, H, W, &need);
  if (rc) return rc;
  if (workspace_bytes < need) return fail(MVSF_ERR_WORKSPACE, "costreg_unet: workspace %zu < %zu bytes", workspace_bytes, need);
  MVSF_REQUIRE(((uintptr_t)workspace & 15) == 0 && ((uintptr_t)wts & 15) == 0 && ((uintptr_t)volume & 15) == 0,
               "costreg_unet: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (wts_tc) {
    MVSF_REQUIRE(((uintptr_t)wts_tc & 15) == 0, "costreg_unet: wts_tc must be 16-byte aligned");
    return unet_forward_tc(kind, volume, wts, reinterpret_cast<const __half*>(wts_tc), logits, workspace, D, H, W, s);
  }
  if (kind == 0) return unet_forward<2>(kind, volume, wts, logits, (float*)workspace, D, H, W, s);
  return unet_forward<1>(kind, volume, wts, logits, (float*)workspace, D, H, W, s);
}
}
