// Micro-benchmark: does MUFU.EX2 issue overlap with FMA-pipe / ALU issue on the same scheduler?  W warps per scheduler, each
// running a loop of 8 independent MUFU.EX2 + NF independent FFMA (+ NA FMNMX).  Prints clk per loop iteration per scheduler.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/mufu_microbench.cu -o tools/mufu_microbench.bin
#include <cuda_runtime.h>
#include <cstdio>
template <int NM, int NF, int NA>
__global__ void k(int iters, float* sink, unsigned* out) {
  float x[8], y[16], z[8];
  for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
  for (int i = 0; i < 16; ++i) y[i] = 0.5f + 0.001f * i;
  for (int i = 0; i < 8; ++i) z[i] = 0.25f * i;
  const float a = 0.999f, b = 1e-6f;
  __syncthreads();
  unsigned t0 = clock();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NM; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i & 7]));
#pragma unroll
    for (int i = 0; i < NF; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y[i & 15]) : "f"(a), "f"(b));
#pragma unroll
    for (int i = 0; i < NA; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(z[i & 7]) : "f"(b));
  }
  unsigned t1 = clock();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += x[i] + z[i];
  for (int i = 0; i < 16; ++i) s += y[i];
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0) out[0] = (t1 - t0) / iters;
}
template <int NM, int NF, int NA>
static void run(int warps_per_sched) {
  float* sink; unsigned* out;
  cudaMalloc(&sink, 16); cudaMalloc(&out, 16);
  k<NM, NF, NA><<<1, 128 * warps_per_sched>>>(2000, sink, out);
  unsigned h = 0;
  cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
  printf("{\"mufu\": %d, \"ffma\": %d, \"fmnmx\": %d, \"warps_per_scheduler\": %d, \"clk_per_iteration_of_one_warp\": %u, \"clk_per_warp_iteration_per_scheduler\": %.1f}\n",
         NM, NF, NA, warps_per_sched, h, (double)h / warps_per_sched);
}
int main() {
  for (int w : {1, 4}) {
    run<8, 0, 0>(w); run<0, 16, 0>(w); run<0, 0, 16>(w);
    run<8, 8, 0>(w); run<8, 16, 0>(w); run<8, 32, 0>(w); run<8, 64, 0>(w);
    run<8, 0, 16>(w); run<8, 16, 16>(w); run<8, 32, 32>(w);
  }
  return 0;
}
