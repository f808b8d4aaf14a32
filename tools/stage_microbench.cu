// Micro-benchmark: how fast can an SM stage a 64 x 16 texel window (C = 8 fp32 channels, 32 KB) of a channels-last feature
// map [H][W][8] into shared memory, as a function of the mechanism?  (Input to the warp + correlation design)
//   mode 0  cp.async.bulk.tensor, 4-D box (c, y&1, x, y/2): the two row parities interleaved, inner rows of 32 B (1024 rows)
//   mode 1  cp.async.bulk.tensor, 2-D boxes over the merged (x*8 + c) axis: 2 boxes of 256 floats x 16 rows (inner rows 1 KB)
//   mode 2  cp.async.bulk (linear), one 2 KB row per request, 16 requests, destination pitch 2048 + 64 B
//   mode 3  256 threads: LDG.128 -> STS.128 (8 of each per thread), pitch 2048 + 64 B
// Every CTA stages `iters` windows at pseudo-random positions back to back (single buffer, one mbarrier) and touches one word
// of each.  Prints ns per window per CTA and aggregate GB/s for 1, 2, 3 CTAs per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/stage_microbench.cu -o gpurun_out/stage_microbench
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

constexpr int WX = 64, WY = 16, C = 8, H = 1152, W = 1536;
constexpr uint32_t WIN = WX * WY * C * 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma4(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma2(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256) bench(const __grid_constant__ CUtensorMap m4, const __grid_constant__ CUtensorMap m2,
                                             const float* __restrict__ feat, int iters, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ unsigned long long barv;
  const uint32_t win = (smem_u32(smem) + 127u) & ~127u, bar = smem_u32(&barv);
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  uint32_t rng = blockIdx.x * 2654435761u + 12345u, phase = 0;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    rng = rng * 1664525u + 1013904223u;
    const int ox = (int)((rng >> 8) % (W - WX)), oy = (int)(((rng >> 4) * 7919u >> 8) % (H - WY)) & ~1;
    if (MODE == 3) {
      for (int i = threadIdx.x; i < WY * (WX * C / 4); i += 256) {
        const int r = i / (WX * C / 4), u = i % (WX * C / 4);
        const float4 v = __ldg(reinterpret_cast<const float4*>(feat + ((size_t)(oy + r) * W + ox) * C) + u);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(win + r * (WX * C * 4 + 64) + u * 16), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      }
      __syncthreads();
    } else {
      if (threadIdx.x < 32) {
        if (threadIdx.x == 0) expect_tx(bar, WIN);
        __syncwarp();
        if (MODE == 0) { if (threadIdx.x == 0) tma4(win, &m4, 0, 0, ox, oy >> 1, bar); }
        if (MODE == 1) { if (threadIdx.x < 2) tma2(win + threadIdx.x * (256 * 4 * WY), &m2, ox * C + threadIdx.x * 256, oy, bar); }
        if (MODE == 2) { if (threadIdx.x < WY) bulk(win + threadIdx.x * (WX * C * 4 + 64), feat + ((size_t)(oy + threadIdx.x) * W + ox) * C, WX * C * 4, bar); }
      }
      mbar_wait(bar, phase);
      phase ^= 1;
    }
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(win + (threadIdx.x & 63) * 64));
    acc += v;
    __syncthreads();
  }
  if (acc == 123.456f) sink[0] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE>
static void run(const CUtensorMap& m4, const CUtensorMap& m2, const float* feat, float* sink, int ctas_per_sm, int iters) {
  const size_t smem = WY * (WX * C * 4 + 64) + 256;
  cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * ctas_per_sm;
  bench<MODE><<<grid, 256, smem>>>(m4, m2, feat, 8, sink);
  cudaEventRecord(e0);
  bench<MODE><<<grid, 256, smem>>>(m4, m2, feat, iters, sink);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("{\"mode\": %d, \"ctas_per_sm\": %d, \"ns_per_window_per_cta\": %.1f, \"aggregate_GBs\": %.1f, \"bytes_per_clk_per_sm\": %.2f, \"status\": \"%s\"}\n",
         MODE, ctas_per_sm, ms * 1e6 / iters, (double)grid * iters * WIN / (ms * 1e6), (double)ctas_per_sm * iters * WIN / (ms * 1e-3 * 1.93e9),
         cudaGetErrorString(err));
}

int main() {
  float *feat, *sink;
  cudaMalloc(&feat, (size_t)H * W * C * 4 + 65536);
  cudaMalloc(&sink, 64);
  cudaMemset(feat, 0, (size_t)H * W * C * 4 + 65536);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(p);
  CUtensorMap m4, m2;
  const cuuint64_t row = (cuuint64_t)W * C * 4;
  {
    const cuuint64_t dims[4] = {C, 2, W, H / 2};
    const cuuint64_t strides[3] = {row, C * 4, 2 * row};
    const cuuint32_t box[4] = {C, 2, WX, WY / 2}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&m4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, feat, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) printf("m4 encode failed %d\n", (int)r);
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)W * C, H};
    const cuuint64_t strides[1] = {row};
    const cuuint32_t box[2] = {256, WY}, es[2] = {1, 1};
    CUresult r = enc(&m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, feat, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) printf("m2 encode failed %d\n", (int)r);
  }
  for (int c = 1; c <= 3; ++c) {
    run<0>(m4, m2, feat, sink, c, 400);
    run<1>(m4, m2, feat, sink, c, 400);
    run<2>(m4, m2, feat, sink, c, 400);
    run<3>(m4, m2, feat, sink, c, 400);
  }
  return 0;
}
