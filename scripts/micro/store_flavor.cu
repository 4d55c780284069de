// Micro-benchmark: write bandwidth of 128-bit stores by cache hint, one or three plane streams.
#include <cstdio>
#include <cuda_runtime.h>
template <int F> __device__ __forceinline__ void st4(float* p, float4 v) {
  if (F == 0) *reinterpret_cast<float4*>(p) = v;
  else if (F == 1) asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  else if (F == 2) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  else if (F == 3) asm volatile("st.global.wt.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  else __stcs(reinterpret_cast<float4*>(p), v);
}
template <int F, int P>
__global__ void wr(float* __restrict__ y, long long n4, long long plane) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = make_float4(1.f, 2.f, 3.f, (float)i);
#pragma unroll
    for (int p = 0; p < P; ++p) st4<F>(y + p * plane + 4 * i, v);
  }
}
// 256-bit stores (sm_100): thread i writes 32 B
template <int P>
__global__ void wr8(float* __restrict__ y, long long n8, long long plane) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float a = 1.f, b = (float)i;
#pragma unroll
    for (int p = 0; p < P; ++p)
      asm volatile("st.global.L2::evict_first.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(y + p * plane + 8 * i), "f"(a), "f"(b), "f"(a), "f"(b), "f"(a), "f"(b), "f"(a), "f"(b) : "memory");
  }
}
template <class F> float timeit(F f) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); for (int i = 0; i < 3; ++i) f(); cudaEventRecord(a); for (int i = 0; i < 20; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms / 20; }
int main() {
  const long long plane = 4ll * 3000 * 4000;
  float* y; cudaMalloc(&y, 3 * plane * 4);
  const double bytes = 12.0 * plane;
#define RUN(F, G, T) { float ms1 = timeit([&] { wr<F, 1><<<G, T>>>(y, 3 * plane / 4, 0); }); float ms3 = timeit([&] { wr<F, 3><<<G, T>>>(y, plane / 4, plane); }); \
  printf("flavor %d grid %d x %d: 1 stream %.4f ms %.0f GB/s | 3 planes %.4f ms %.0f GB/s\n", F, G, T, ms1, bytes / ms1 / 1e6, ms3, bytes / ms3 / 1e6); }
  RUN(0, 148 * 8, 256) RUN(1, 148 * 8, 256) RUN(2, 148 * 8, 256) RUN(3, 148 * 8, 256) RUN(4, 148 * 8, 256)
  RUN(0, 148 * 16, 128) RUN(0, 148 * 4, 512) RUN(0, 148 * 2, 1024) RUN(0, 148 * 64, 256)
  { float ms1 = timeit([&] { wr8<1><<<148 * 8, 256>>>(y, 3 * plane / 8, 0); }); float ms3 = timeit([&] { wr8<3><<<148 * 8, 256>>>(y, plane / 8, plane); });
    printf("256-bit stores: 1 stream %.4f ms %.0f GB/s | 3 planes %.4f ms %.0f GB/s\n", ms1, bytes / ms1 / 1e6, ms3, bytes / ms3 / 1e6); }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
