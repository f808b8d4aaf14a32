// W2+W3+W4 for the fine stages (C = 8 / 16 feature channels): homography warp fused with group-wise correlation
// (models/warping.py:69-109, models/cost_volume.py:72-101) around TMA-staged source windows in shared memory.
//
// Why a second organisation next to warp_corr.cu: with 8 or 16 channels one bilinear corner is only 32 / 64 bytes, so a
// warp-wide global gather touches 16 / 8 different 128-byte lines per instruction and runs at 25-50 % of the L1
// wavefront rate, and the taps of neighbouring pixels are NOT coherent at stages 2-4 (the hypothesis planes follow the
// previous stage's per-pixel depth).  Here
//   * a CTA owns a tile of reference pixels; per (source view, chunk of <= 8 hypotheses) it computes its tap
//     coordinates, reduces their bounding box, and one warp issues the source footprint as bulk asynchronous copies
//     (cp.async.bulk, the TMA engine's linear mode: one 2-4 KB row segment per request, completion on an mbarrier); the
//     window lands in shared memory as rows [y][x][C] with a row pitch of (row bytes + 64), parts of the window outside
//     the image are zero-filled (= grid_sample's padding_mode='zeros', no per-corner masking).  (A first version used ONE
//     5-D cp.async.bulk.tensor box with the two row parities interleaved; its 32-byte inner rows made the TMA unit the
//     bottleneck - ~4 clk per box row, 2 us per 32 KB window - see DESIGN.md.)
//   * every lane then owns whole taps: the 4 corners x 8 channels of a tap are 8 pieces of 16 bytes.  With the odd
//     pitch the 16-byte bank group of a piece is  (4y + 2x + quad) mod 8  (C = 8), i.e. bit0 = channel quad,
//     bit1 = x parity, bit2 = (x>>1 parity) XOR (y parity): the 8 pieces of ANY tap fall on 8 different bank groups.
//     Lane l reads them in the order (round i) bank group = i XOR (l mod 8), so the 8 lanes of every LDS.128 phase hit 8
//     different bank groups: conflict-free shared-memory gathers at 128 B/clk/SM for arbitrary (incoherent) tap
//     positions.  Which corner a round delivers depends on the tap's x/y parities; that is folded into per-tap swaps of
//     the x / y weights and base addresses (one-bit XORs), so the 8 rounds themselves are straight-line code;
//   * taps outside the staged window (depth outliers) fall back to global loads for that lane only.
// Pass A (entropy) and pass B (view aggregation) both gather (no spill of the per-view correlations: at these stages
// the spill costs more HBM time than the second gather).
#include <float.h>
#include <limits.h>

#include "common.cuh"

#ifndef MVSF_WT_PASSB_BLOCKS
#define MVSF_WT_PASSB_BLOCKS 3   // resident CTAs per SM of the C = 8, D <= 4 aggregation kernel (3 costs ~20 spilled registers)
#endif

namespace mvsf {
namespace wt {

constexpr int TW = 32, TH = 8, THREADS = 256;   // reference-pixel tile: one warp per tile row
constexpr int DCH = 8;                          // hypotheses per window (register-resident tap coordinates)
constexpr int kMaxD = 512;

template <int C>
struct Cfg {
  static constexpr int LPX = C / 8;                  // lanes per pixel: every lane owns 8 channels of a tap
  static constexpr int PIX = THREADS / LPX;          // pixels per CTA
  static constexpr int TROWS = PIX / TW;             // tile rows
  static constexpr int WX = 64, WY = 16;             // staged window (source texels)
  static constexpr int TEX = C * 4;                  // bytes of one texel
  static constexpr int ROWB = WX * TEX;              // bytes of one window row (2 KB / 4 KB)
  static constexpr int PITCH = ROWB + 64;            // odd multiple of 64 B: consecutive rows sit in opposite bank halves
  static constexpr uint32_t BYTES = WY * PITCH;      // 33 KB (C = 8), 65 KB (C = 16)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// one lane polls, the warp reconverges; bounded so that a mis-programmed pipeline traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) {
    uint32_t it = 0;
    while (!mbar_try_wait(bar, parity))
      if (++it > (1u << 26)) __trap();
  }
  __syncwarp();
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// acc (4 channels) += w * t, as two packed fp32x2 FMAs (FFMA2: one issue slot per two FMAs)
__device__ __forceinline__ void fma4(float4& acc, float w, const float4& t) {
  const float2 ww = make_float2(w, w);
  float2 lo = __ffma2_rn(make_float2(t.x, t.y), ww, make_float2(acc.x, acc.y));
  float2 hi = __ffma2_rn(make_float2(t.z, t.w), ww, make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x))); }

// per-lane constants of the rotated read order
struct Lane {
  bool b0, b1, b2;      // bits of (lane & 7) that are XORed into the bank-group index of a round
  uint32_t base[2];     // window base + 16-byte piece offset for rounds with i0 = 0 / 1
};

// One tap (4 corners x 8 channels of this lane) from the staged window: sA / sB receive the bilinear sample of the lane's
// two channel quads (quad A = the one read in rounds with i0 = 0).  (lx, ly) = integer tap position relative to the window
// origin (both rows / columns inside the window), (fx, fy) = fractional parts.
// C = 8 : bank group of piece (x, y, quad) = (4y + 2x + quad) mod 8: bit0 = quad, bit1 = x&1, bit2 = ((x>>1)&1) ^ (y&1).
//         Round (i2,i1,i0) of lane (b2,b1,b0) reads group (i2^b2, i1^b1, i0^b0):
//           dx = i1 ^ b1 ^ xodd,  dy = i2 ^ b2 ^ ((lx>>1)&1) ^ yodd ^ (dx & xodd),  quad = i0 ^ b0 (folded into base[]).
// C = 16: bank group of piece (x, y, quarter) = (4y + 4x + quarter) mod 8: bits 0-1 = quarter, bit2 = (x&1) ^ (y&1).  A lane
//         owns quarters {2k, 2k+1} (k = lane & 1).  Round (ia,ib,ic): dx = ic, dy = ia ^ ic ^ b2 ^ xodd ^ yodd, low quarter
//         bit = ib ^ b1 (folded into base[]).
template <int C>
__device__ __forceinline__ void gather_window(const Lane& L, int lx, int ly, float fx, float fy, float4& sA, float4& sB) {
  using K = Cfg<C>;
  const float gx = 1.0f - fx, gy = 1.0f - fy;
  const bool xodd = lx & 1, yodd = ly & 1;
  const uint32_t Y0 = (uint32_t)ly * K::PITCH + (uint32_t)lx * K::TEX, Y1 = Y0 + K::PITCH;   // tap origin in row ly / ly + 1
  if (C == 8) {
    const bool a = (lx >> 1) & 1;
    const bool b = L.b1 != xodd;                               // dx of the rounds with i1 = 0
    const bool e0 = ((L.b2 != a) != yodd) != (b && xodd);      // dy of round (i2 = 0, i1 = 0)
    const bool e1 = ((L.b2 != a) != yodd) != (!b && xodd);     // dy of round (i2 = 0, i1 = 1)
    const uint32_t X0 = b ? K::TEX : 0, X1 = b ? 0 : K::TEX;
    const float wx0 = b ? fx : gx, wx1 = b ? gx : fx;
    const uint32_t R00 = (e0 ? Y1 : Y0) + X0, R10 = (e0 ? Y0 : Y1) + X0;   // R[i2][i1]
    const uint32_t R01 = (e1 ? Y1 : Y0) + X1, R11 = (e1 ? Y0 : Y1) + X1;
    const float w00 = wx0 * (e0 ? fy : gy), w10 = wx0 * (e0 ? gy : fy);
    const float w01 = wx1 * (e1 ? fy : gy), w11 = wx1 * (e1 ? gy : fy);
    const float4 a0 = lds128(R00 + L.base[0]), b0 = lds128(R00 + L.base[1]);
    const float4 a1 = lds128(R01 + L.base[0]), b1 = lds128(R01 + L.base[1]);
    const float4 a2 = lds128(R10 + L.base[0]), b2 = lds128(R10 + L.base[1]);
    const float4 a3 = lds128(R11 + L.base[0]), b3 = lds128(R11 + L.base[1]);
    fma4(sA, w00, a0); fma4(sB, w00, b0);
    fma4(sA, w01, a1); fma4(sB, w01, b1);
    fma4(sA, w10, a2); fma4(sB, w10, b2);
    fma4(sA, w11, a3); fma4(sB, w11, b3);
  } else {
    const bool e = (L.b2 != xodd) != yodd;                     // dy of the rounds with ia ^ ic = 0
    const uint32_t U0 = e ? Y1 : Y0, U1 = e ? Y0 : Y1;         // row of rounds with ia ^ ic = 0 / 1
    const float v0 = e ? fy : gy, v1 = e ? gy : fy;
    // (ia, ic): (0,0) -> U0, dx 0 | (0,1) -> U1, dx 1 | (1,0) -> U1, dx 0 | (1,1) -> U0, dx 1
    const float w00 = gx * v0, w01 = fx * v1, w10 = gx * v1, w11 = fx * v0;
    const float4 a0 = lds128(U0 + L.base[0]), b0 = lds128(U0 + L.base[1]);
    const float4 a1 = lds128(U1 + L.base[0] + K::TEX), b1 = lds128(U1 + L.base[1] + K::TEX);
    const float4 a2 = lds128(U1 + L.base[0]), b2 = lds128(U1 + L.base[1]);
    const float4 a3 = lds128(U0 + L.base[0] + K::TEX), b3 = lds128(U0 + L.base[1] + K::TEX);
    fma4(sA, w00, a0); fma4(sB, w00, b0);
    fma4(sA, w01, a1); fma4(sB, w01, b1);
    fma4(sA, w10, a2); fma4(sB, w10, b2);
    fma4(sA, w11, a3); fma4(sB, w11, b3);
  }
}

// same tap through global memory (zero padding by masked weights, as warp_corr.cu): the rare out-of-window taps
template <int C>
__device__ __forceinline__ void gather_global(const float* __restrict__ srcA, const float* __restrict__ srcB, float ix, float iy,
                                              int W, int H, float4& sA, float4& sB) {
  int4 off;
  float4 wt;
  make_tap_fast(ix, iy, W, H, C, off, wt);
  fma4(sA, wt.x, ldg4(srcA + off.x)); fma4(sB, wt.x, ldg4(srcB + off.x));
  fma4(sA, wt.y, ldg4(srcA + off.y)); fma4(sB, wt.y, ldg4(srcB + off.y));
  fma4(sA, wt.z, ldg4(srcA + off.z)); fma4(sB, wt.z, ldg4(srcB + off.z));
  fma4(sA, wt.w, ldg4(srcA + off.w)); fma4(sB, wt.w, ldg4(srcB + off.w));
}

struct TapCoord {
  float fx, fy;
  int x0, y0;
  bool inb;   // sample position inside (-1, W) x (-1, H): otherwise every corner is outside the image -> contributes 0
};
__device__ __forceinline__ TapCoord split_coord(float ix, float iy, int W, int H) {
  TapCoord t;
  t.inb = (ix > -1.0f) && (ix < (float)W) && (iy > -1.0f) && (iy < (float)H);   // false for NaN / Inf
  const float sx = t.inb ? ix : 0.0f, sy = t.inb ? iy : 0.0f;
  const float MAGIC = 12582912.0f;   // 1.5 * 2^23: round-down add = floor (|s| < 2^22)
  const float tx = __fadd_rd(sx, MAGIC), ty = __fadd_rd(sy, MAGIC);
  t.x0 = __float_as_int(tx) - 0x4B400000;
  t.y0 = __float_as_int(ty) - 0x4B400000;
  t.fx = sx - (tx - MAGIC);
  t.fy = sy - (ty - MAGIC);
  return t;
}

struct Shared {
  unsigned long long bar;
  int bbox[2][4];   // double-buffered {min x0, max x0, min y0, max y0} of the current window's taps
};

// Stages one window: reduces the bounding box of the CTA's taps, centres the WX x WY window on it, issues its rows as bulk
// asynchronous copies (warp 0: one row per lane), zero-fills whatever lies outside the image and waits for the data.
// Returns the window origin to every thread.  `slot` alternates per call.  src = feature map of the view, [H][W][C].
template <int C>
__device__ __forceinline__ void stage_window(const float* __restrict__ src, Shared& sh, uint32_t win, int slot, uint32_t& phase,
                                             int H, int W, int mnx, int mxx, int mny, int mxy, int& ox, int& oy) {
  using K = Cfg<C>;
  mnx = __reduce_min_sync(0xffffffffu, mnx);
  mxx = __reduce_max_sync(0xffffffffu, mxx);
  mny = __reduce_min_sync(0xffffffffu, mny);
  mxy = __reduce_max_sync(0xffffffffu, mxy);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&sh.bbox[slot][0], mnx);
    atomicMax(&sh.bbox[slot][1], mxx);
    atomicMin(&sh.bbox[slot][2], mny);
    atomicMax(&sh.bbox[slot][3], mxy);
  }
  __syncthreads();   // bbox complete; every lane has also finished reading the previous window
  const int bx0 = sh.bbox[slot][0], bx1 = sh.bbox[slot][1], by0 = sh.bbox[slot][2], by1 = sh.bbox[slot][3];
  if (bx0 > bx1) { ox = 0; oy = 0; }
  else {
    const int slack_x = K::WX - (bx1 + 2 - bx0), slack_y = K::WY - (by1 + 2 - by0);
    ox = bx0 - (slack_x > 0 ? slack_x / 2 : 0);
    oy = by0 - (slack_y > 0 ? slack_y / 2 : 0);
  }
  // the part of the window that lies inside the image (columns [x_lo, x_hi), rows [y_lo, y_hi))
  const int x_lo = max(ox, 0), x_hi = min(ox + K::WX, W), y_lo = max(oy, 0), y_hi = min(oy + K::WY, H);
  const bool any = x_hi > x_lo && y_hi > y_lo;
  const uint32_t bar = smem_u32(&sh.bar);
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;          // window row
    if (r == 0) {
      sh.bbox[slot ^ 1][0] = INT_MAX; sh.bbox[slot ^ 1][1] = INT_MIN;   // reset the other slot for the next window
      sh.bbox[slot ^ 1][2] = INT_MAX; sh.bbox[slot ^ 1][3] = INT_MIN;
      const uint32_t total = any ? (uint32_t)(y_hi - y_lo) * (uint32_t)(x_hi - x_lo) * K::TEX : 0u;
      expect_tx(bar, total);
    }
    __syncwarp();
    const int y = oy + r;
    if (any && r < K::WY && y >= y_lo && y < y_hi)
      bulk_load(win + (uint32_t)r * K::PITCH + (uint32_t)(x_lo - ox) * K::TEX, src + ((size_t)y * W + x_lo) * C,
                (uint32_t)(x_hi - x_lo) * K::TEX, bar);
  }
  if (!(any && x_lo == ox && x_hi == ox + K::WX && y_lo == oy && y_hi == oy + K::WY)) {
    // border window (CTA-uniform, rare): zero everything the copies do not write (disjoint from what they write)
    for (int i = threadIdx.x; i < K::WY * (K::ROWB / 16); i += THREADS) {
      const int r = i / (K::ROWB / 16), u = i % (K::ROWB / 16);
      const int x = ox + (u * 16) / K::TEX, y = oy + r;
      if (!any || y < y_lo || y >= y_hi || x < x_lo || x >= x_hi)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(win + (uint32_t)r * K::PITCH + (uint32_t)u * 16), "r"(0) : "memory");
    }
    __syncthreads();
  }
  mbar_wait_warp(bar, phase);
  phase ^= 1;
}

// Everything a lane needs to sample source view `view` for its pixel.
template <int C>
struct ViewCtx {
  Hom m;
  float rx, ry, rz;
  const float* src;    // feature map of the source view [H][W][C]
  const float* srcA;   // + first channel of the lane's quad A / quad B (global fallback)
  const float* srcB;
};

// One (view, hypothesis chunk [d0, d0 + n)) of this lane's pixel: tap coordinates, window staging, gather.
// consume(k, sA, sB) receives the bilinear sample of hypothesis d0 + k (quad A / quad B channels of the lane).
// The bounding box is taken from the first and last hypothesis of the chunk: taps of one pixel lie on its epipolar line
// and move monotonically with the (monotone) hypotheses, so those two bound the rest; a tap that still falls outside the
// window (non-monotone caller-supplied hypotheses, depth outliers of neighbours) goes through global memory.
template <int C, int DCHT, typename F>
__device__ __forceinline__ void process_chunk(Shared& sh, uint32_t win, int& slot, uint32_t& phase,
                                              const Lane& L, const ViewCtx<C>& vc, const float* __restrict__ depth_p,
                                              int HW, int d0, int n, bool active, const CoordConst& cc, int W, int H, F&& consume) {
  using K = Cfg<C>;
  float ix, iy;
  warp_coord_lean(vc.rx, vc.ry, vc.rz, vc.m, __ldg(depth_p + (size_t)d0 * HW), cc, ix, iy);
  TapCoord tF = split_coord(ix, iy, W, H);
  tF.inb = tF.inb && active;
  TapCoord tL = tF;
  if (n > 1) {
    warp_coord_lean(vc.rx, vc.ry, vc.rz, vc.m, __ldg(depth_p + (size_t)(d0 + n - 1) * HW), cc, ix, iy);
    tL = split_coord(ix, iy, W, H);
    tL.inb = tL.inb && active;
  }
  int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
  if (tF.inb) { mnx = mxx = tF.x0; mny = mxy = tF.y0; }
  if (tL.inb) { mnx = min(mnx, tL.x0); mxx = max(mxx, tL.x0); mny = min(mny, tL.y0); mxy = max(mxy, tL.y0); }
  int ox, oy;
  stage_window<C>(vc.src, sh, win, slot, phase, H, W, mnx, mxx, mny, mxy, ox, oy);
  slot ^= 1;
#pragma unroll
  for (int k = 0; k < DCHT; ++k) {
    if (k < n) {
      TapCoord t;
      if (k == 0) t = tF;
      else if (k == n - 1) t = tL;
      else {
        warp_coord_lean(vc.rx, vc.ry, vc.rz, vc.m, __ldg(depth_p + (size_t)(d0 + k) * HW), cc, ix, iy);
        t = split_coord(ix, iy, W, H);
        t.inb = t.inb && active;
      }
      float4 sA = make_float4(0.f, 0.f, 0.f, 0.f), sB = sA;
      const int lx = t.x0 - ox, ly = t.y0 - oy;
      const bool inwin = t.inb && (unsigned)lx <= (unsigned)(K::WX - 2) && (unsigned)ly <= (unsigned)(K::WY - 2);
      gather_window<C>(L, inwin ? lx : 0, inwin ? ly : 0, t.fx, t.fy, sA, sB);
      if (!inwin) {
        sA = make_float4(0.f, 0.f, 0.f, 0.f); sB = sA;
        if (t.inb) {   // rare: sample through global memory (coordinates recomputed: they are not kept in registers)
          warp_coord_lean(vc.rx, vc.ry, vc.rz, vc.m, __ldg(depth_p + (size_t)(d0 + k) * HW), cc, ix, iy);
          gather_global<C>(vc.srcA, vc.srcB, ix, iy, W, H, sA, sB);
        }
      }
      consume(k, sA, sB);
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------------
// MODE 0: pass A  -> entropy[v][pixel]      (cost_volume.py:89-92)
// MODE 1: pass B  -> volume[d][pixel][g]    (cost_volume.py:95-101), G = 8 groups
// DCHT: hypotheses per window (4 or 8).  GENERIC (pass A only): D > DCHT, similarities parked in a per-thread array.
// ----------------------------------------------------------------------------------------------------------------------
template <int C, int MODE, int DCHT, bool GENERIC>
__global__ void __launch_bounds__(THREADS, (C == 8 && (MODE == 0 || (DCHT == 4 && MVSF_WT_PASSB_BLOCKS == 3))) ? 3 : 2)
warp_tile_kernel(const float* __restrict__ feat, const float* __restrict__ homs,
                 const float* __restrict__ depth, const float* __restrict__ vis, float* __restrict__ out, int V, int D, int H,
                 int W, int dch) {
  using K = Cfg<C>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ Shared sh;
  const uint32_t win = (smem_u32(smem_raw) + 127u) & ~127u;
  const int tid = threadIdx.x, lane = tid & 31;
  const int HW = H * W;
  // pixel of this lane: LPX adjacent lanes share a pixel
  const int pix_in_cta = tid / K::LPX;
  const int px = blockIdx.x * TW + (pix_in_cta % TW), py = blockIdx.y * K::TROWS + (pix_in_cta / TW);
  const bool active = (px < W) && (py < H);
  const int p = min(py, H - 1) * W + min(px, W - 1);
  const int sub = tid % K::LPX;           // which 8-channel slice of the pixel
  Lane L;
  int chA, chB;                            // first channel of the lane's quad A / quad B
  if (C == 8) {
    L.b0 = lane & 1; L.b1 = (lane >> 1) & 1; L.b2 = (lane >> 2) & 1;
    L.base[0] = win + (L.b0 ? 16 : 0);
    L.base[1] = win + (L.b0 ? 0 : 16);
    chA = L.b0 ? 4 : 0; chB = L.b0 ? 0 : 4;
  } else {
    // lanes 2j, 2j+1 = the two channel halves of one pixel; bits 1-2 of the lane = tap slot inside an LDS.128 phase
    L.b0 = false; L.b1 = (lane >> 1) & 1; L.b2 = (lane >> 2) & 1;
    const int q0 = 2 * sub + (L.b1 ? 1 : 0), q1 = 2 * sub + (L.b1 ? 0 : 1);
    L.base[0] = win + q0 * 16;
    L.base[1] = win + q1 * 16;
    chA = q0 * 4; chB = q1 * 4;
  }
  if (tid == 0) {
    mbar_init(smem_u32(&sh.bar), 1);
    fence_barrier_init();
    sh.bbox[0][0] = INT_MAX; sh.bbox[0][1] = INT_MIN; sh.bbox[0][2] = INT_MAX; sh.bbox[0][3] = INT_MIN;
    sh.bbox[1][0] = INT_MAX; sh.bbox[1][1] = INT_MIN; sh.bbox[1][2] = INT_MAX; sh.bbox[1][3] = INT_MIN;
  }
  __syncthreads();
  uint32_t phase = 0;
  int slot = 0;
  const CoordConst cc = make_coord_const(W, H);
  const float fxp = (float)min(px, W - 1), fyp = (float)min(py, H - 1);
  const float4 rA = ldg4(feat + (size_t)p * C + chA), rB = ldg4(feat + (size_t)p * C + chB);
  constexpr float inv_cpg = 8.0f / (float)C;   // G / C with G = 8 groups: mean over the channels of a group
  const float* __restrict__ depth_p = depth + p;

  auto make_view = [&](int v) {
    ViewCtx<C> vc;
    vc.m = load_hom(homs + (size_t)v * 12);
    vc.rx = __fadd_rn(fmaf(vc.m.r01, fyp, __fmul_rn(vc.m.r00, fxp)), vc.m.r02);
    vc.ry = __fadd_rn(fmaf(vc.m.r11, fyp, __fmul_rn(vc.m.r10, fxp)), vc.m.r12);
    vc.rz = __fadd_rn(fmaf(vc.m.r21, fyp, __fmul_rn(vc.m.r20, fxp)), vc.m.r22);
    vc.src = feat + (size_t)(v + 1) * HW * C;
    vc.srcA = vc.src + chA;
    vc.srcB = vc.src + chB;
    return vc;
  };

  if (MODE == 0) {
    // ------------------------------------------------------------------ pass A: views outer, hypothesis chunks inner
    float sims[GENERIC ? kMaxD : DCHT];
    for (int v = 0; v < V - 1; ++v) {
      const ViewCtx<C> vc = make_view(v);
      float mx = -FLT_MAX;
      for (int d0 = 0; d0 < D; d0 += dch) {
        const int n = min(dch, D - d0);
        process_chunk<C, DCHT>(sh, win, slot, phase, L, vc, depth_p, HW, d0, n, active, cc, W, H,
                               [&](int k, const float4& sA, const float4& sB) {
                                 float s = dot4(rA, sA) + dot4(rB, sB);
                                 if (K::LPX == 2) s += __shfl_xor_sync(0xffffffffu, s, 1);
                                 s *= inv_cpg;
                                 sims[GENERIC ? d0 + k : k] = s;
                                 mx = fmaxf(mx, s);
                               });
        if (!GENERIC) break;
      }
      // softmax over D -> entropy (cost_volume.py:90-92): p = exp(s - max) / Z ; H = -sum p * log(p + 1e-7)
      float Z = 0.f, ent = 0.f;
      if (GENERIC) {
        for (int d = 0; d < D; ++d) { sims[d] = expf(sims[d] - mx); Z += sims[d]; }
        for (int d = 0; d < D; ++d) { const float pr = __fdiv_rn(sims[d], Z); ent -= pr * logf(pr + 1e-7f); }
      } else {
#pragma unroll
        for (int k = 0; k < DCHT; ++k)
          if (k < D) { sims[k] = expf(sims[k] - mx); Z += sims[k]; }
#pragma unroll
        for (int k = 0; k < DCHT; ++k)
          if (k < D) { const float pr = __fdiv_rn(sims[k], Z); ent -= pr * logf(pr + 1e-7f); }
      }
      if (active && sub == 0) out[(size_t)v * HW + p] = ent;
    }
  } else {
    // ------------------------------------------------------------------ pass B: hypothesis chunks outer, views inner
    float wsum = 0.f;
    for (int v = 0; v < V - 1; ++v) wsum = __fadd_rn(wsum, __ldg(vis + (size_t)v * HW + p));
    const float den = __fadd_rn(wsum, 1e-6f);
    // ref * (1 / channels per group): what every warped channel is multiplied with before the visibility weight
    const float4 qA = make_float4(rA.x * inv_cpg, rA.y * inv_cpg, rA.z * inv_cpg, rA.w * inv_cpg);
    const float4 qB = make_float4(rB.x * inv_cpg, rB.y * inv_cpg, rB.z * inv_cpg, rB.w * inv_cpg);
    for (int d0 = 0; d0 < D; d0 += dch) {
      const int n = min(dch, D - d0);
      float4 accA[DCHT], accB[DCHT];
#pragma unroll
      for (int k = 0; k < DCHT; ++k) accA[k] = accB[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int v = 0; v < V - 1; ++v) {
        const ViewCtx<C> vc = make_view(v);
        const float w = __ldg(vis + (size_t)v * HW + p);
        process_chunk<C, DCHT>(sh, win, slot, phase, L, vc, depth_p, HW, d0, n, active, cc, W, H,
                               [&](int k, const float4& sA, const float4& sB) {
                                 // group correlation of this view (cost_volume.py:78-85) times its weight (:97)
                                 accA[k].x = fmaf(qA.x * sA.x, w, accA[k].x); accA[k].y = fmaf(qA.y * sA.y, w, accA[k].y);
                                 accA[k].z = fmaf(qA.z * sA.z, w, accA[k].z); accA[k].w = fmaf(qA.w * sA.w, w, accA[k].w);
                                 accB[k].x = fmaf(qB.x * sB.x, w, accB[k].x); accB[k].y = fmaf(qB.y * sB.y, w, accB[k].y);
                                 accB[k].z = fmaf(qB.z * sB.z, w, accB[k].z); accB[k].w = fmaf(qB.w * sB.w, w, accB[k].w);
                               });
      }
      if (active) {
#pragma unroll
        for (int k = 0; k < DCHT; ++k) {
          if (k < n) {
            float* o = out + ((size_t)(d0 + k) * HW + p) * 8;
            if (C == 8) {   // channels == groups
              *reinterpret_cast<float4*>(o + chA) = make_float4(__fdiv_rn(accA[k].x, den), __fdiv_rn(accA[k].y, den),
                                                                __fdiv_rn(accA[k].z, den), __fdiv_rn(accA[k].w, den));
              *reinterpret_cast<float4*>(o + chB) = make_float4(__fdiv_rn(accB[k].x, den), __fdiv_rn(accB[k].y, den),
                                                                __fdiv_rn(accB[k].z, den), __fdiv_rn(accB[k].w, den));
            } else {        // 2 channels per group: a quad is 2 groups
              *reinterpret_cast<float2*>(o + chA / 2) = make_float2(__fdiv_rn(accA[k].x + accA[k].y, den), __fdiv_rn(accA[k].z + accA[k].w, den));
              *reinterpret_cast<float2*>(o + chB / 2) = make_float2(__fdiv_rn(accB[k].x + accB[k].y, den), __fdiv_rn(accB[k].z + accB[k].w, den));
            }
          }
        }
      }
    }
  }
}

template <int C, int MODE, int DCHT, bool GENERIC>
static int launch(const float* feat, const float* homs, const float* depth, const float* vis, float* out, int V, int D, int H,
                  int W, int dch, cudaStream_t s) {
  using K = Cfg<C>;
  auto kern = warp_tile_kernel<C, MODE, DCHT, GENERIC>;
  static DeviceOnce once;
  const int dev = current_device();
  const size_t smem = K::BYTES + 128;
  if (once.need(dev)) {
    MVSF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    once.done(dev);
  }
  dim3 grid(cdiv(W, TW), cdiv(H, K::TROWS));
  kern<<<grid, THREADS, smem, s>>>(feat, homs, depth, vis, out, V, D, H, W, dch);
  return MVSF_OK;
}

template <int C>
static int dispatch(int mode, const float* feat, const float* homs, const float* depth, const float* vis, float* out, int V,
                    int D, int H, int W, cudaStream_t s) {
  if (mode == 0) {
    if (D <= 4) return launch<C, 0, 4, false>(feat, homs, depth, vis, out, V, D, H, W, D, s);
    if (D <= 8) return launch<C, 0, 8, false>(feat, homs, depth, vis, out, V, D, H, W, D, s);
    int c = D / 24;   // plane sweeps: the epipolar span of a chunk has to stay inside the window
    c = c < 2 ? 2 : (c > 8 ? 8 : c);
    return launch<C, 0, 8, true>(feat, homs, depth, vis, out, V, D, H, W, c, s);
  }
  if (D <= 4) return launch<C, 1, 4, false>(feat, homs, depth, vis, out, V, D, H, W, D, s);
  int c = D <= 8 ? D : D / 24;
  c = c < 2 ? 2 : (c > 8 ? 8 : c);
  return launch<C, 1, 8, false>(feat, homs, depth, vis, out, V, D, H, W, c, s);
}

}  // namespace wt

// Used by warp_corr.cu's entry points.  Returns false when this organisation does not apply (other channel counts,
// misaligned pointers: the bulk copies need 16-byte aligned row segments).
bool warp_tile_supported(const float* feat, int C, int G, int D, int H, int W) {
  return (C == 8 || C == 16) && G == 8 && H >= 2 && W >= 2 && D >= 1 && D <= wt::kMaxD && ((uintptr_t)feat & 15) == 0;
}
// pass A: entropy [V-1][H][W]
int warp_tile_entropy(const float* feat, const float* homs, const float* depth, float* entropy, int V, int C, int D, int H,
                      int W, cudaStream_t s) {
  return C == 8 ? wt::dispatch<8>(0, feat, homs, depth, nullptr, entropy, V, D, H, W, s)
                : wt::dispatch<16>(0, feat, homs, depth, nullptr, entropy, V, D, H, W, s);
}
// pass B: volume [D][H][W][8]
int warp_tile_aggregate(const float* feat, const float* homs, const float* depth, const float* vis, float* volume, int V, int C,
                        int D, int H, int W, cudaStream_t s) {
  return C == 8 ? wt::dispatch<8>(1, feat, homs, depth, vis, volume, V, D, H, W, s)
                : wt::dispatch<16>(1, feat, homs, depth, vis, volume, V, D, H, W, s);
}

}  // namespace mvsf
