// Micro-benchmark: tcgen05.ld throughput (4 warps, 32 columns per instruction) alone and while another warp issues
// tcgen05.mma (M = 128, N = 16, K = 16) back to back.  Build like tools/mma_microbench.cu.
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace mvsf::umma;

// mode 0: loads only | 1: MMAs only | 2: both concurrently
template <int MODE, int NLDW>
__global__ void __launch_bounds__(32 * (NLDW + 1), 1) bench(int iters, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar = sb + 65536;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + 65536 + 16);
  for (int i = tid; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(sb + 65536 + 16, 512);
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = *slot;
  const uint32_t idesc = make_idesc_f16(128, 16);
  const uint64_t ad = make_desc(sb, 2048, 128), bd = make_desc(sb + 8192, 256, 128);
  if (warp < NLDW) {
    if (MODE != 1) {
      const uint32_t t = tm + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
      uint32_t acc = 0;
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        uint32_t r[32];
        tmem_ld32_nowait(t + (i & 3) * 32, r);
        tmem_ld_wait();
        acc += r[0] ^ r[31];
      }
      const long long t1 = clock64();
      if (lane == 0) out[warp] = t1 - t0 + (acc == 0x12345u);
    }
  } else if (lane == 0 && MODE != 0) {
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) mma_f16_ss(tm + 256 + (uint32_t)((i & 3) * 16), ad, bd, idesc, 1u);
    commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    out[15] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int MODE, int NLDW>
void run(int iters, long long* d) {
  cudaFuncSetAttribute(bench<MODE, NLDW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  long long h[16];
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(d, 0, sizeof(h));
    bench<MODE, NLDW><<<1, 32 * (NLDW + 1), 66 * 1024>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", MODE, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("mode %d, %d load warps: ld x32 = %.1f clk per instruction per warp (%.1f B/clk/SM) | mma = %.1f clk\n", MODE, NLDW,
         (double)h[0] / iters, h[0] ? (double)NLDW * 32 * 32 * 4 * iters / h[0] : 0.0, (double)h[15] / iters);
}

int main() {
  long long* d;
  cudaMalloc(&d, 128);
  const int iters = 4000;
  run<0, 1>(iters, d);
  run<0, 4>(iters, d);
  run<0, 8>(iters, d);
  run<1, 4>(iters, d);
  run<2, 4>(iters, d);
  run<2, 8>(iters, d);
  return 0;
}
