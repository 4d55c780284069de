// Latency-bound gradient exchange over NVLink peer memory: the data-parallel tuning / search steps average a few hundred
// floats per backward pass (module-parameter and architecture-weight gradients: 37 .. 216 floats; reference: the DDP bucket of
// darts_model.py:31,173 and isp_model.py).  An NCCL all-reduce of that size costs ~15 us of pure latency on a 0.36 ms step and
// keeps the step from being one CUDA graph.  Here every rank owns a small buffer that all ranks of the node map through CUDA
// IPC; one single-CTA kernel per rank
//   1. stores its vector into slot[parity][rank] of EVERY rank's buffer (peer stores over NVLink / NVSwitch),
//   2. publishes an epoch flag on every rank (release, system scope),
//   3. waits until the flags of all ranks have arrived in its own buffer (acquire, system scope; bounded by wall clock),
//   4. sums the world slots IN RANK ORDER -- every rank computes bit-identical averages, run to run.
// Slots and flags are double-buffered by epoch parity: a rank can only start epoch e+2 after every rank has finished reading
// epoch e (it needed their e+1 flags), so nothing is overwritten while it may still be read.  The epoch counter lives on the
// device, so the kernel is CUDA-graph capturable (no host argument changes between replays).
#include "risp_common.cuh"

namespace risp {

constexpr int kP2PMaxWorld = 8;
constexpr int kP2PThreads = 128;

struct P2PArgs {
  float* base[kP2PMaxWorld];     // base[r]: rank r's buffer as mapped into THIS process
  int rank, world, cap;
};

// buffer of one rank: float slots[2][world][cap] | uint32 flags[2][world] | uint32 epoch | uint32 timeouts
__host__ __device__ inline size_t p2p_flag_offset_floats(int world, int cap) { return (size_t)2 * world * cap; }

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(kP2PThreads)
p2p_allreduce_mean_kernel(float* __restrict__ data, int n, P2PArgs a, unsigned long long timeout_ns) {
  __shared__ unsigned s_epoch;
  const int tid = threadIdx.x, world = a.world, rank = a.rank, cap = a.cap;
  float* self = a.base[rank];
  unsigned* ctl = reinterpret_cast<unsigned*>(self + p2p_flag_offset_floats(world, cap));     // flags[2][world], epoch, timeouts
  if (tid == 0) { s_epoch = ctl[2 * world] + 1u; ctl[2 * world] = s_epoch; }
  __syncthreads();
  const unsigned e = s_epoch, par = e & 1u;
  // 1. my vector -> slot[par][rank] of every rank
  for (int p = 0; p < world; ++p) {
    float* dst = a.base[p] + ((size_t)par * world + rank) * cap;
    for (int i = tid; i < n; i += kP2PThreads) dst[i] = data[i];
  }
  __threadfence_system();
  __syncthreads();
  // 2. + 3. flags out, flags in
  if (tid < world) {
    unsigned* out = reinterpret_cast<unsigned*>(a.base[tid] + p2p_flag_offset_floats(world, cap)) + par * world + rank;
    st_release_sys(out, e);
    const unsigned* in = ctl + par * world + tid;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(in) != e) {
      __nanosleep(64);
      if (global_ns() - t0 > timeout_ns) { atomicAdd(ctl + 2 * world + 1, 1u); break; }       // a lost peer must not hang the device
    }
  }
  __syncthreads();
  // 4. rank-ordered sum (volatile loads: the lines were written by peers, never trust a cached copy)
  const float inv = 1.f / (float)world;
  for (int i = tid; i < n; i += kP2PThreads) {
    float s = 0.f;
    for (int r = 0; r < world; ++r) s += __ldcv(self + ((size_t)par * world + r) * cap + i);
    data[i] = s * inv;
  }
}

}  // namespace risp

using namespace risp;

extern "C" size_t risp_p2p_buffer_bytes(int world, int cap) {
  if (world < 1 || world > kP2PMaxWorld || cap < 1) return 0;
  return sizeof(float) * p2p_flag_offset_floats(world, cap) + sizeof(unsigned) * (2 * (size_t)world + 2);
}

extern "C" int risp_p2p_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  RISP_REQUIRE(bytes > 0 && ptr && handle64, RISP_E_INVALID, "risp_p2p_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess || cudaMemset(p, 0, bytes) != cudaSuccess) { set_error("risp_p2p_alloc: cudaMalloc/cudaMemset of %zu bytes failed", bytes); cudaGetLastError(); return RISP_E_CUDA; }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { set_error("risp_p2p_alloc: cudaIpcGetMemHandle failed"); cudaGetLastError(); cudaFree(p); return RISP_E_CUDA; }
  memcpy(handle64, &h, 64);
  cudaDeviceSynchronize();
  *ptr = p;
  return RISP_OK;
}

extern "C" int risp_p2p_open(const unsigned char* handle64, void** peer_ptr) {
  RISP_REQUIRE(handle64 && peer_ptr, RISP_E_INVALID, "risp_p2p_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { set_error("risp_p2p_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return RISP_E_CUDA; }
  *peer_ptr = p;
  return RISP_OK;
}

extern "C" int risp_p2p_close(void* peer_ptr) {
  if (peer_ptr && cudaIpcCloseMemHandle(peer_ptr) != cudaSuccess) { cudaGetLastError(); return RISP_E_CUDA; }
  return RISP_OK;
}

extern "C" int risp_p2p_free(void* ptr) {
  if (ptr && cudaFree(ptr) != cudaSuccess) { cudaGetLastError(); return RISP_E_CUDA; }
  return RISP_OK;
}

// number of waits that ran into the timeout so far (0 = healthy); buffer = this rank's own buffer
extern "C" int risp_p2p_timeouts(const void* buffer, int world, int cap, unsigned* out_host) {
  RISP_REQUIRE(buffer && out_host && world >= 1 && world <= kP2PMaxWorld && cap >= 1, RISP_E_INVALID, "risp_p2p_timeouts: bad arguments");
  const unsigned* ctl = reinterpret_cast<const unsigned*>(static_cast<const float*>(buffer) + p2p_flag_offset_floats(world, cap));
  if (cudaMemcpy(out_host, ctl + 2 * world + 1, sizeof(unsigned), cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("risp_p2p_timeouts: copy failed"); return RISP_E_CUDA; }
  return RISP_OK;
}

// data (n floats, device, this rank) <- mean over ranks, in place.  bases: HOST array of `world` device pointers, bases[r] =
// rank r's buffer as mapped into this process (bases[rank] = the own allocation).  Every rank must call it the same number of
// times (it is a collective).  timeout_ms bounds the wait for the peers' flags.
extern "C" int risp_p2p_allreduce_mean(float* data, int n, void* const* bases, int rank, int world, int cap, int timeout_ms,
                                       risp_stream_t stream) {
  RISP_REQUIRE(data && bases && world >= 1 && world <= kP2PMaxWorld && rank >= 0 && rank < world && n >= 1 && n <= cap && timeout_ms > 0,
               RISP_E_INVALID, "risp_p2p_allreduce_mean: bad arguments (n %d, cap %d, world %d)", n, cap, world);
  P2PArgs a;
  for (int r = 0; r < kP2PMaxWorld; ++r) a.base[r] = r < world ? static_cast<float*>(bases[r]) : nullptr;
  a.rank = rank; a.world = world; a.cap = cap;
  for (int r = 0; r < world; ++r) RISP_REQUIRE(a.base[r], RISP_E_INVALID, "risp_p2p_allreduce_mean: null buffer of rank %d", r);
  p2p_allreduce_mean_kernel<<<1, kP2PThreads, 0, as_stream(stream)>>>(data, n, a, (unsigned long long)timeout_ms * 1000000ull);
  return check_launch("p2p_allreduce_mean_kernel");
}
