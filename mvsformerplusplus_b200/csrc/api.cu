// Error reporting, launch accounting and boundary layout helpers of libmvsf_b200.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mvsf {
static thread_local char g_err[1024] = "";
static thread_local long long g_launches = 0;

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(int n) { g_launches += n; }

// [N][C][HW] -> [N][HW][C] through a 32x32 shared-memory tile (coalesced on both sides)
__global__ void transpose_chw_hwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const float* s = src + (size_t)n * C * HW;
  float* d = dst + (size_t)n * C * HW;
  int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) tile[i][threadIdx.x] = s[(size_t)c * HW + p];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, c = c0 + threadIdx.x;
    if (c < C && p < HW) d[(size_t)p * C + c] = tile[threadIdx.x][i];
  }
}
// [N][HW][C] -> [N][C][HW]
__global__ void transpose_hwc_chw_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const float* s = src + (size_t)n * C * HW;
  float* d = dst + (size_t)n * C * HW;
  int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, c = c0 + threadIdx.x;
    if (c < C && p < HW) tile[i][threadIdx.x] = s[(size_t)p * C + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) d[(size_t)c * HW + p] = tile[threadIdx.x][i];
  }
}
}  // namespace mvsf

extern "C" {
const char* mvsf_last_error(void) { return mvsf::g_err; }
int mvsf_abi_version(void) { return 1; }
long long mvsf_launch_count(int reset) {
  long long v = mvsf::g_launches;
  if (reset) mvsf::g_launches = 0;
  return v;
}

int mvsf_nchw_to_nhwc(const float* src, float* dst, int N, int C, int HW, mvsf_stream_t stream) {
  MVSF_REQUIRE(src && dst && N > 0 && C > 0 && HW > 0 && N <= 65535, "nchw_to_nhwc: bad arguments");
  dim3 grid(mvsf::cdiv(HW, 32), mvsf::cdiv(C, 32), N), block(32, 8);
  mvsf::transpose_chw_hwc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, dst, C, HW);
  MVSF_LAUNCH_CHECK("nchw_to_nhwc");
  return MVSF_OK;
}
int mvsf_nhwc_to_nchw(const float* src, float* dst, int N, int C, int HW, mvsf_stream_t stream) {
  MVSF_REQUIRE(src && dst && N > 0 && C > 0 && HW > 0 && N <= 65535, "nhwc_to_nchw: bad arguments");
  dim3 grid(mvsf::cdiv(HW, 32), mvsf::cdiv(C, 32), N), block(32, 8);
  mvsf::transpose_hwc_chw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, dst, C, HW);
  MVSF_LAUNCH_CHECK("nhwc_to_nchw");
  return MVSF_OK;
}
}
