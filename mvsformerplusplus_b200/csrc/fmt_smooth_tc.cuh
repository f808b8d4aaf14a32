// FMT pathway level (models/FMT.py:154-162,195-197):  out = smooth( bilinear_up2(red) + lateral )  with smooth = Conv2d(C, C, 3,
// padding=1, bias=False), fused into one persistent kernel.  Included by fmt.cu (inside namespace mvsf).
//   phase 1 (SIMT)   pre = lateral (NCHW) + bilinear_up2(red) (NHWC, F.interpolate align_corners=False) on the 18 x 34 halo
//                    of a 16 x 32 output tile, written as fp16 hi|lo voxel-octet planes in shared memory (zero outside the
//                    image = the conv padding) - the upsampled + added tensor never goes to HBM
//   phase 2 (tcgen05) 3x3 conv as implicit GEMM: four 16 x 8 M-tiles, a tap = a descriptor start address (conv3d_tc.cu);
//                    C = 8: one MMA per tap  [x_hi | x_lo] x [[w_hi;w_hi] | [w_lo;0]]  (N = 32);
//                    C >= 16: per 16 channels  x_hi x [w_hi | w_lo] (N = 2C) and x_lo x w_hi (N = C) onto the first half
//   phase 3          TMEM -> registers: add the two halves, fp32 NHWC store
#pragma once

namespace sm2 {
using namespace umma;
constexpr int PR = 18, PC = 34;
constexpr uint32_t PLANE = PR * PC * 16, PITCH = PC * 16;
template <int C>
struct Cfg {
  static constexpr int NO = C / 8;                       // channel octets
  static constexpr int NPAD = C < 16 ? 16 : C;           // rows of one weight part
  static constexpr int NG = C < 16 ? 1 : C / 16;         // K = 16 groups
  static constexpr uint32_t BT = 2 * 2 * NPAD * 16;      // one (tap, group) weight tile: 2 k-chunks x 2 NPAD rows x 16 B
  static constexpr uint32_t OFF_PL = 0, OFF_BT = 2 * NO * PLANE, OFF_BAR = OFF_BT + 9 * NG * BT, SMEM = OFF_BAR + 32;
  static constexpr uint32_t TCOLS = 4 * 2 * NPAD;        // accumulator columns (4 M-tiles x [first | second] part)
};
}  // namespace sm2

template <int C>
__global__ void __launch_bounds__(256)
fmt_smooth_tc_kernel(const float* __restrict__ red, const float* __restrict__ lat, const float* __restrict__ wts,
                     float* __restrict__ out, int h, int w, int tiles_x, int tiles_y, int ntiles) {
  using namespace sm2;
  using K = Cfg<C>;
  constexpr int NO = K::NO, NPAD = K::NPAD, NG = K::NG;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = 2 * h, W = 2 * w;
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar = sb + K::OFF_BAR;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + K::OFF_BAR + 16);
  uint32_t ncols = 32;
  while (ncols < K::TCOLS) ncols <<= 1;

  // ---- once per CTA: weight tiles.  wts = [tap][ci][co] fp32.  Tile (tap, g) = [2 k-chunks][2 NPAD rows][8 halves]:
  //      C >= 16: rows [0, NPAD) = w_hi, [NPAD, 2 NPAD) = w_lo, k-chunk kc = input channels 16 g + 8 kc + e
  //      C == 8 : K = [x_hi | x_lo] of the single octet: rows [0,16) = w_hi in BOTH k-chunks; rows [16,32) = w_lo in k-chunk 0, 0 in 1
  for (int i = tid; i < 9 * NG * 2 * 2 * NPAD * 8; i += 256) {
    const int e = i & 7;
    int q = i >> 3;
    const int row = q % (2 * NPAD); q /= 2 * NPAD;
    const int kc = q & 1; q >>= 1;
    const int g = q % NG, tap = q / NG;
    const int part = row / NPAD, n = row % NPAD;
    const int ci = C < 16 ? e : g * 16 + kc * 8 + e;
    float wv = 0.f;
    if (n < C) wv = __ldg(wts + ((size_t)tap * C + ci) * C + n);
    const __half hi = __float2half_rn(wv), lo = __float2half_rn(wv - __half2float(hi));
    __half v;
    if (C < 16) v = part == 0 ? hi : (kc == 0 ? lo : __float2half_rn(0.f));
    else v = part == 0 ? hi : lo;
    reinterpret_cast<__half*>(smem + K::OFF_BT)[i] = v;
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(sb + K::OFF_BAR + 16, ncols);
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t el = elect_one();
  constexpr uint32_t a_hi = desc_hi(PITCH), b_hi = desc_hi(128);
  const uint32_t idesc_full = make_idesc_f16(128, 2 * NPAD), idesc_half = make_idesc_f16(128, NPAD);
  uint32_t phase = 0;

  const int quarter = warp & 3, m = quarter * 32 + lane;
  const int er = m >> 3, ec0 = (warp >> 2) * 16 + (m & 7);     // warps 0-3: M-tiles 0, 1; warps 4-7: M-tiles 2, 3
  const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((warp >> 2) * 2 * 2 * NPAD);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, v = tile / (tiles_x * tiles_y);
    const int x0 = tx * 32, y0 = ty * 16;
    const float* lv = lat + (size_t)v * C * H * W;
    const float* rv = red + (size_t)v * h * w * C;
    // ---- phase 1: pre = lateral + bilinear_up2(red) on the halo region -> planes [octet][hi|lo]
    for (int i = tid; i < NO * PR * PC; i += 256) {
      const int o = i / (PR * PC), pix = i - o * (PR * PC);
      const int r = pix / PC, c = pix - r * PC;
      const int y = y0 - 1 + r, x = x0 - 1 + c;
      float pre[8];
      if (y >= 0 && y < H && x >= 0 && x < W) {
        // ATen area_pixel_compute_source_index(scale=0.5, align_corners=False): src = 0.5*(dst+0.5)-0.5, clamped at 0
        const float sy = fmaxf(0.5f * ((float)y + 0.5f) - 0.5f, 0.0f);
        const int ya = (int)sy, yb = ya + ((ya < h - 1) ? 1 : 0);
        const float ly1 = sy - (float)ya, ly0 = 1.0f - ly1;
        const float sx = fmaxf(0.5f * ((float)x + 0.5f) - 0.5f, 0.0f);
        const int xa = (int)sx, xb = xa + ((xa < w - 1) ? 1 : 0);
        const float lx1 = sx - (float)xa, lx0 = 1.0f - lx1;
        const float* p00 = rv + ((size_t)ya * w + xa) * C + o * 8;
        const float* p01 = rv + ((size_t)ya * w + xb) * C + o * 8;
        const float* p10 = rv + ((size_t)yb * w + xa) * C + o * 8;
        const float* p11 = rv + ((size_t)yb * w + xb) * C + o * 8;
#pragma unroll
        for (int q4 = 0; q4 < 2; ++q4) {
          const float4 v00 = ldg4(p00 + q4 * 4), v01 = ldg4(p01 + q4 * 4), v10 = ldg4(p10 + q4 * 4), v11 = ldg4(p11 + q4 * 4);
          const float a00[4] = {v00.x, v00.y, v00.z, v00.w}, a01[4] = {v01.x, v01.y, v01.z, v01.w};
          const float a10[4] = {v10.x, v10.y, v10.z, v10.w}, a11[4] = {v11.x, v11.y, v11.z, v11.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float up = ly0 * (lx0 * a00[e] + lx1 * a01[e]) + ly1 * (lx0 * a10[e] + lx1 * a11[e]);
            pre[q4 * 4 + e] = up + __ldg(lv + ((size_t)(o * 8 + q4 * 4 + e) * H + y) * W + x);
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) pre[e] = 0.f;
      }
      __half* p = reinterpret_cast<__half*>(smem + K::OFF_PL + (uint32_t)(2 * o) * PLANE + (uint32_t)pix * 16u);
      split_store8(p, p + PLANE / 2, pre);     // hi plane of octet o, then its lo plane (+ PLANE bytes)
    }
    fence_proxy_async();
    tc_fence_before_sync();
    __syncthreads();
    // ---- phase 2: MMAs (converged warp 0, elected lane)
    if (warp == 0) {
      tc_fence_after_sync();
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const uint32_t aoff = (uint32_t)(kh * PC + kw) * 16u;
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            const uint32_t wb = desc_lo(sb + K::OFF_BT + (uint32_t)((kh * 3 + kw) * NG + g) * K::BT, 2 * NPAD * 16);
            const uint32_t acc = (kh | kw | g) ? 1u : 0u;
            if (C < 16) {
              const uint32_t ad = desc_lo(sb + K::OFF_PL + aoff, PLANE);                     // K = [hi | lo] planes of the octet
#pragma unroll
              for (int ct = 0; ct < 4; ++ct) mma_f16_ss_lh(el, tmem_base + ct * 2 * NPAD, ad + ct * 8, a_hi, wb, b_hi, idesc_full, acc);
            } else {
              // planes of group g: [hi o(2g) | lo o(2g) | hi o(2g+1) | lo o(2g+1)]: K chunks = the two hi (or lo) planes
              const uint32_t ah = desc_lo(sb + K::OFF_PL + (uint32_t)(4 * g) * PLANE + aoff, 2 * PLANE), al = ah + (PLANE >> 4);
#pragma unroll
              for (int ct = 0; ct < 4; ++ct) mma_f16_ss_lh(el, tmem_base + ct * 2 * NPAD, ah + ct * 8, a_hi, wb, b_hi, idesc_full, acc);
#pragma unroll
              for (int ct = 0; ct < 4; ++ct) mma_f16_ss_lh(el, tmem_base + ct * 2 * NPAD, al + ct * 8, a_hi, wb, b_hi, idesc_half, 1u);
            }
          }
        }
      }
      commit_el(el, bar);
    }
    // ---- phase 3: epilogue
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after_sync();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int y = y0 + er, x = x0 + ec0 + k * 8;
      const bool valid = y < H && x < W;
      float* op = out + (((size_t)v * H + (valid ? y : 0)) * W + (valid ? x : 0)) * C;
#pragma unroll
      for (int c16 = 0; c16 < NPAD / 16; ++c16) {
        float a[16], b[16];
        tmem_ld16(trow + k * 2 * NPAD + c16 * 16, a);
        tmem_ld16(trow + k * 2 * NPAD + NPAD + c16 * 16, b);
        if (valid) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            if (c16 * 16 + q4 * 4 < C)
              *reinterpret_cast<float4*>(op + c16 * 16 + q4 * 4) =
                  make_float4(a[q4 * 4] + b[q4 * 4], a[q4 * 4 + 1] + b[q4 * 4 + 1], a[q4 * 4 + 2] + b[q4 * 4 + 2], a[q4 * 4 + 3] + b[q4 * 4 + 3]);
          }
        }
      }
    }
    tc_fence_before_sync();
    __syncthreads();   // planes and accumulators are free again
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

template <int C>
static int launch_fmt_smooth_tc(const float* red, const float* lat, const float* wts, float* out, int V, int h, int w,
                                cudaStream_t s) {
  using K = sm2::Cfg<C>;
  static DeviceOnce once;
  static int per_sm = 1;   // a function of the kernel's compile-time footprint only
  const int dev = current_device();
  const int num_sms = device_sm_count(dev);
  if (once.need(dev)) {
    MVSF_CUDA_OK(cudaFuncSetAttribute(fmt_smooth_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM));
    // resident CTAs per SM: registers (<= 80 x 256 threads -> 3), shared memory, tensor memory (512 columns per SM)
    uint32_t ncols = 32;
    while (ncols < K::TCOLS) ncols <<= 1;
    per_sm = 3;
    if (per_sm > (int)(512 / ncols)) per_sm = 512 / ncols;
    if (per_sm > (int)((220 * 1024) / K::SMEM)) per_sm = (int)((220 * 1024) / K::SMEM);
    if (per_sm < 1) per_sm = 1;
    once.done(dev);
  }
  const int H = 2 * h, W = 2 * w;
  const int tiles_x = cdiv(W, 32), tiles_y = cdiv(H, 16);
  const long long ntiles = (long long)tiles_x * tiles_y * V;
  MVSF_REQUIRE(ntiles < (1ll << 30), "fmt pathway: image too large");
  const long long cap = (long long)per_sm * num_sms;
  fmt_smooth_tc_kernel<C><<<(int)(ntiles < cap ? ntiles : cap), 256, K::SMEM, s>>>(red, lat, wts, out, h, w, tiles_x, tiles_y,
                                                                                       (int)ntiles);
  MVSF_LAUNCH_CHECK("fmt_smooth_tc");
  return MVSF_OK;
}
