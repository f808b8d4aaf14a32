// Micro-benchmark: cost of back-to-back tcgen05.mma.kind::f16 (M = 128, K = 16) issued by one thread, as a function of
// N, operand source (A from shared memory / tensor memory), accumulator reuse and issue style.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mvsformerplusplus_b200/csrc tools/mma_microbench.cu -o gpurun_out/mma_microbench
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace mvsf::umma;

// mode 0: same accumulator, smem A | 1: 4 accumulators round robin, smem A | 2: same accumulator, TMEM A
// mode 3: 4 accumulators, TMEM A | 4: same as 1 but issued by a converged warp with elect.sync
template <int MODE>
__global__ void __launch_bounds__(128, 1) bench(int n, int iters, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar = sb + 65536;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + 65536 + 16);
  for (int i = tid; i < 16384; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 ones
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(sb + 65536 + 16, 512);
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = *slot;
  const uint32_t idesc = make_idesc_f16(128, n);
  const uint64_t ad = make_desc(sb, 2048, 128), bd = make_desc(sb + 8192, (uint32_t)n * 16u, 128);
  long long t0 = 0, t1 = 0;
  if (MODE == 4) {
    if (warp == 1) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
        if (pred) mma_f16_ss(tm + (uint32_t)((i & 3) * 128), ad, bd, idesc, 1u);
        __syncwarp();
      }
      if ((tid & 31) == 0) { commit(bar); mbar_wait(bar, 0); }
      t1 = clock64();
      if ((tid & 31) == 0) out[0] = t1 - t0;
    }
  } else if (tid == 32) {
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
      const uint32_t acc = tm + (uint32_t)(((MODE & 1) ? (i & 3) : 0) * 128);
      if (MODE >= 2) mma_f16_ts(acc, tm + 480, bd, idesc, 1u);
      else mma_f16_ss(acc, ad, bd, idesc, 1u);
    }
    const long long tissue = clock64();
    commit(bar);
    mbar_wait(bar, 0);
    t1 = clock64();
    out[0] = t1 - t0;
    out[1] = tissue - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int MODE>
void run(int n, int iters, long long* d) {
  cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  long long h[2] = {0, 0};
  for (int rep = 0; rep < 2; ++rep) {
    bench<MODE><<<1, 128, 66 * 1024>>>(n, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d n %d: %s\n", MODE, n, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("mode %d N=%3d: %.1f clk/MMA total, %.1f clk/MMA issue\n", MODE, n, (double)h[0] / iters, (double)h[1] / iters);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const int iters = 2000;
  for (int n : {16, 48, 64, 128, 256}) {
    run<0>(n, iters, d);
    if (n <= 128) run<1>(n, iters, d);
    run<2>(n, iters, d);
    if (n <= 128) run<3>(n, iters, d);
    if (n <= 128) run<4>(n, iters, d);
  }
  return 0;
}
