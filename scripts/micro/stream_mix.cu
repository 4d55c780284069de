// Micro-benchmark: ceiling of a 4 B/px read + 12 B/px write kernel (the fused inference's traffic) for different
// access orders.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/stream_mix scripts/micro/stream_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void st4(float* p, float4 v) { asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
__device__ __forceinline__ float4 ld4(const float* p) { float4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v; }
// (a) linear: thread i handles float4 i of the plane, grid-stride
__global__ void linear(const float* __restrict__ raw, float* __restrict__ y, long long plane4, long long plane, int doread, int dowrite) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < plane4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = doread ? ld4(raw + 4 * i) : make_float4(1.f, 2.f, 3.f, (float)i);
    if (dowrite) { st4(y + 4 * i, v); v.x += 1.f; st4(y + plane + 4 * i, v); v.y += 1.f; st4(y + 2 * plane + 4 * i, v); }
    else if (v.x == 123.4567f) st4(y + 4 * i, v);
  }
}
// (b) fronts: warp = 128-column strip marching down a chunk of rows (the fused kernel's order); U rows in flight per warp
template <int U>
__global__ void fronts(const float* __restrict__ raw, float* __restrict__ y, int H, int W, int rows, int chunks, int strips, int doread, int dowrite) {
  const int lane = threadIdx.x & 31;
  const long long plane = (long long)H * W;
  const int wpb = blockDim.x / 32;
  const int items = chunks * strips;
  for (int item = blockIdx.x * wpb + (threadIdx.x >> 5); item < items; item += gridDim.x * wpb) {
    const int chunk = item / strips, strip = item - chunk * strips;
    const int c0 = strip * 128 + lane * 4;
    if (c0 >= W) continue;
    const int ra = chunk * rows, rb = min(H, ra + rows);
    for (int r = ra; r < rb; r += U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = (doread && r + u < rb) ? ld4(raw + (long long)(r + u) * W + c0) : make_float4(1.f, 2.f, 3.f, (float)r);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r + u >= rb) break;
        float* p = y + (long long)(r + u) * W + c0;
        if (dowrite) { st4(p, v[u]); v[u].x += 1.f; st4(p + plane, v[u]); v[u].y += 1.f; st4(p + 2 * plane, v[u]); }
        else if (v[u].x == 123.4567f) st4(p, v[u]);
      }
    }
  }
}
template <class F> float timeit(F f) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); for (int i = 0; i < 3; ++i) f(); cudaEventRecord(a); for (int i = 0; i < 20; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms / 20; }
int main() {
  const int N = 4, H = 3000, W = 4000;
  const long long plane = (long long)H * W;
  float *raw, *y;
  cudaMalloc(&raw, N * plane * 4); cudaMalloc(&y, N * 3 * plane * 4);
  cudaMemset(raw, 0, N * plane * 4);
  // treat the N frames as one tall frame of N*H rows for simplicity (same traffic)
  const int HH = N * H;
  const long long pl = (long long)HH * W;
  for (int rw = 3; rw >= 1; --rw) {
    const int rd = rw & 1, wr = (rw >> 1) & 1;
    const double bytes = (rd ? 4.0 : 0.0) * pl + (wr ? 12.0 : 0.0) * pl;
    float ms = timeit([&] { linear<<<148 * 8, 256>>>(raw, y, pl / 4, pl, rd, wr); });
    printf("linear  read=%d write=%d: %.4f ms  %.0f GB/s\n", rd, wr, ms, bytes / ms / 1e6);
    for (int rows : {82, 16}) {
      const int chunks = (HH + rows - 1) / rows, strips = (W + 127) / 128;
      ms = timeit([&] { fronts<1><<<148 * 8, 128>>>(raw, y, HH, W, rows, chunks, strips, rd, wr); });
      printf("fronts<1> rows=%d read=%d write=%d: %.4f ms  %.0f GB/s\n", rows, rd, wr, ms, bytes / ms / 1e6);
      ms = timeit([&] { fronts<4><<<148 * 8, 128>>>(raw, y, HH, W, rows, chunks, strips, rd, wr); });
      printf("fronts<4> rows=%d read=%d write=%d: %.4f ms  %.0f GB/s\n", rows, rd, wr, ms, bytes / ms / 1e6);
      ms = timeit([&] { fronts<4><<<148 * 16, 32>>>(raw, y, HH, W, rows, chunks, strips, rd, wr); });
      printf("fronts<4> 1-warp CTAs rows=%d read=%d write=%d: %.4f ms  %.0f GB/s\n", rows, rd, wr, ms, bytes / ms / 1e6);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
