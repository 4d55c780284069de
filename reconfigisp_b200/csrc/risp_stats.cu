// Per-image statistics: plane min/mean/max, log-average luminance, torch.histc-compatible histograms
// and the exact k-th largest value per plane (radix select).
// References: grayworld tools_origin.py:35-41; SRCNNRes global features srcnn_res_arch.py:36-40;
// whiteworld :655-662; reinhard :535-546; conditional-module histogram :120-129 (which the reference
// computes on the CPU with a device->host sync per channel per image).
#include "risp_common.cuh"

namespace risp {

constexpr int kT = 256;

static int stat_blocks(int planes, long long HW) {
  long long g = cdiv(HW, (long long)kT * 16);
  long long cap = (long long)sm_count() * 8 / (planes > 0 ? planes : 1);
  if (cap < 4) cap = 4;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

__global__ void __launch_bounds__(kT)
plane_stats_kernel(const float* __restrict__ x, float* __restrict__ partial, long long HW, int vec) {
  const int plane = blockIdx.y;
  const float* p = x + (long long)plane * HW;
  float mn = INFINITY, mx = -INFINITY, sm = 0.f;
  if (vec) {
    const long long nv = HW / 4;
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < nv; i += (long long)gridDim.x * kT) {
      float4 v = ld_stream4(p + 4 * i);
      mn = fminf(fminf(mn, fminf(v.x, v.y)), fminf(v.z, v.w));
      mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
      sm += (v.x + v.y) + (v.z + v.w);
    }
  } else {
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < HW; i += (long long)gridDim.x * kT) {
      float v = p[i];
      mn = fminf(mn, v); mx = fmaxf(mx, v); sm += v;
    }
  }
  __shared__ float red[kT / 32][3];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  mn = warp_min(mn); mx = warp_max(mx); sm = warp_sum(sm);
  if (lane == 0) { red[wid][0] = mn; red[wid][1] = sm; red[wid][2] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kT / 32; ++w) {
      mn = fminf(mn, red[w][0]); sm += red[w][1]; mx = fmaxf(mx, red[w][2]);
    }
    float* o = partial + ((long long)plane * gridDim.x + blockIdx.x) * 4;
    o[0] = mn; o[1] = sm; o[2] = mx;
  }
}

__global__ void plane_stats_final(const float* __restrict__ partial, float* __restrict__ out, int B, float inv_hw) {
  const int plane = blockIdx.x, lane = threadIdx.x;
  float mn = INFINITY, mx = -INFINITY, sm = 0.f;
  for (int b = lane; b < B; b += 32) {
    const float* o = partial + ((long long)plane * B + b) * 4;
    mn = fminf(mn, o[0]); sm += o[1]; mx = fmaxf(mx, o[2]);
  }
  mn = warp_min(mn); mx = warp_max(mx); sm = warp_sum(sm);
  if (lane == 0) { out[plane * 3] = mn; out[plane * 3 + 1] = sm * inv_hw; out[plane * 3 + 2] = mx; }
}

// first position (row-major) of the plane's minimum and maximum: where torch.min(x, dim=3) -> torch.min(.., dim=2) of
// srcnn_res_arch.py:36-40 sends its gradient on the CPU (first occurrence among ties).  idx (planes, 2) starts at INT_MAX.
__global__ void __launch_bounds__(kT)
plane_argfirst_kernel(const float* __restrict__ x, const float* __restrict__ stats, int* __restrict__ idx, long long HW) {
  const int plane = blockIdx.y;
  const float* p = x + (long long)plane * HW;
  const float mn = stats[plane * 3], mx = stats[plane * 3 + 2];
  int imn = 0x7fffffff, imx = 0x7fffffff;
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < HW; i += (long long)gridDim.x * kT) {
    const float v = p[i];
    if (v == mn && (int)i < imn) imn = (int)i;
    if (v == mx && (int)i < imx) imx = (int)i;
  }
  imn = __reduce_min_sync(0xffffffffu, imn);
  imx = __reduce_min_sync(0xffffffffu, imx);
  if ((threadIdx.x & 31) == 0) {
    if (imn != 0x7fffffff) atomicMin(idx + plane * 2, imn);
    if (imx != 0x7fffffff) atomicMin(idx + plane * 2 + 1, imx);
  }
}

// dx = g_mean / HW everywhere, + g_min at the arg-min pixel, + g_max at the arg-max pixel;  g (planes, 3) = d/d[min, mean, max]
__global__ void __launch_bounds__(kT)
plane_stats_bwd_kernel(const float* __restrict__ g, const int* __restrict__ idx, float* __restrict__ dx, long long HW, float inv_hw) {
  const int plane = blockIdx.y;
  const float gm = g[plane * 3 + 1] * inv_hw, gmn = g[plane * 3], gmx = g[plane * 3 + 2];
  const int imn = idx[plane * 2], imx = idx[plane * 2 + 1];
  float* p = dx + (long long)plane * HW;
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < HW; i += (long long)gridDim.x * kT)
    p[i] = gm + ((int)i == imn ? gmn : 0.f) + ((int)i == imx ? gmx : 0.f);
}

__global__ void __launch_bounds__(kT)
loglum_kernel(const float* __restrict__ x, float* __restrict__ partial, long long HW, float scale) {
  const int n = blockIdx.y;
  const float* p = x + (long long)n * 3 * HW;
  float sm = 0.f;
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < HW; i += (long long)gridDim.x * kT) {
    float lum = 0.114f * (p[i] * scale) + 0.587f * (p[HW + i] * scale) + 0.299f * (p[2 * HW + i] * scale);
    sm += logf(lum + 1e-6f);
  }
  __shared__ float red[kT / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  sm = warp_sum(sm);
  if (lane == 0) red[wid] = sm;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kT / 32; ++w) sm += red[w];
    partial[(long long)n * gridDim.x + blockIdx.x] = sm;
  }
}

__global__ void loglum_final(const float* __restrict__ partial, float* __restrict__ out, int B, float inv_hw) {
  const int n = blockIdx.x, lane = threadIdx.x;
  float sm = 0.f;
  for (int b = lane; b < B; b += 32) sm += partial[(long long)n * B + b];
  sm = warp_sum(sm);
  if (lane == 0) out[n] = sm * inv_hw;
}

// torch.histc(x, bins, min=0, max=1): pos = (int)(x * bins), x == 1 -> last bin, outside ignored.
// Counts are integers < 2^24 per bin for frames up to 16 MP, so float atomics are exact and
// order-independent.
__global__ void __launch_bounds__(kT)
histc_kernel(const float* __restrict__ x, float* __restrict__ out, long long HW, int bins) {
  extern __shared__ unsigned int sh[];
  for (int i = threadIdx.x; i < bins; i += kT) sh[i] = 0;
  __syncthreads();
  const int plane = blockIdx.y;
  const float* p = x + (long long)plane * HW;
  const float fb = (float)bins;
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < HW; i += (long long)gridDim.x * kT) {
    float v = p[i];
    if (v >= 0.f && v <= 1.f) {
      int pos = (int)(__fmul_rn(v, fb));
      pos = pos < bins - 1 ? pos : bins - 1;
      atomicAdd(&sh[pos], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += kT)
    if (sh[i]) atomicAdd(out + (long long)plane * bins + i, (float)sh[i]);
}

// ---- radix select (k-th largest) -----------------------------------------------------------------------
struct SelState {
  unsigned int prefix;        // selected high digits so far
  unsigned int pad;
  long long k;                // remaining 1-based rank among elements matching the prefix (descending)
  unsigned int hist[256];
};

__device__ __forceinline__ unsigned int order_key(float f) {
  unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);   // ascending unsigned order == ascending float order
}
__device__ __forceinline__ float key_to_float(unsigned int k) {
  unsigned int b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}

__global__ void sel_init_kernel(SelState* st, const long long* __restrict__ k, long long HW) {
  const int plane = blockIdx.x;
  if (threadIdx.x == 0) {
    long long kk = k[plane];
    kk = kk < 1 ? 1 : (kk > HW ? HW : kk);
    st[plane].prefix = 0; st[plane].k = kk;
  }
  st[plane].hist[threadIdx.x] = 0;
}

__global__ void __launch_bounds__(kT)
sel_hist_kernel(const float* __restrict__ x, SelState* st, long long HW, int pass) {
  __shared__ unsigned int sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const int plane = blockIdx.y;
  const float* p = x + (long long)plane * HW;
  const int shift = 24 - 8 * pass;
  const unsigned int prefix = st[plane].prefix;
  const unsigned int himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < HW; i += (long long)gridDim.x * kT) {
    unsigned int key = order_key(p[i]);
    if ((key & himask) == prefix) atomicAdd(&sh[(key >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&st[plane].hist[threadIdx.x], sh[threadIdx.x]);
}

__global__ void sel_pick_kernel(SelState* st, float* __restrict__ out, int pass) {
  const int plane = blockIdx.x;
  __shared__ unsigned int h[256];
  h[threadIdx.x] = st[plane].hist[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    long long k = st[plane].k;
    int d = 255;
    for (; d > 0; --d) {            // descending digits: the k-th LARGEST
      if (k <= (long long)h[d]) break;
      k -= h[d];
    }
    const int shift = 24 - 8 * pass;
    unsigned int prefix = st[plane].prefix | ((unsigned int)d << shift);
    st[plane].prefix = prefix; st[plane].k = k;
    if (pass == 3) out[plane] = key_to_float(prefix);
  }
  __syncthreads();
  st[plane].hist[threadIdx.x] = 0;
}

}  // namespace risp

using namespace risp;

extern "C" size_t risp_plane_stats_workspace(int planes, long long HW) {
  if (planes <= 0 || HW <= 0) return 0;
  return (size_t)planes * stat_blocks(planes, HW) * 4 * sizeof(float);
}

extern "C" int risp_plane_stats(const float* x, float* out, int planes, long long HW, void* workspace,
                                size_t workspace_bytes, risp_stream_t stream) {
  RISP_REQUIRE(x && out && planes > 0 && HW > 0 && planes <= 65535, RISP_E_INVALID, "risp_plane_stats: bad arguments");
  RISP_REQUIRE(workspace && workspace_bytes >= risp_plane_stats_workspace(planes, HW), RISP_E_WORKSPACE,
               "risp_plane_stats: workspace too small");
  cudaStream_t st = as_stream(stream);
  int B = stat_blocks(planes, HW);
  float* partial = static_cast<float*>(workspace);
  int vec = (HW % 4 == 0) && aligned16(x);
  plane_stats_kernel<<<dim3(B, planes), kT, 0, st>>>(x, partial, HW, vec);
  plane_stats_final<<<planes, 32, 0, st>>>(partial, out, B, (float)(1.0 / (double)HW));
  return check_launch("plane_stats");
}

// idx: (planes, 2) int32 = row-major position of the first minimum / first maximum of every plane (stats from risp_plane_stats)
extern "C" int risp_plane_argfirst(const float* x, const float* stats, int* idx, int planes, long long HW, risp_stream_t stream) {
  RISP_REQUIRE(x && stats && idx && planes > 0 && planes <= 65535 && HW > 0 && HW < 0x7f000000ll, RISP_E_INVALID,
               "risp_plane_argfirst: bad arguments");          // idx starts at 0x7f7f7f7f (byte memset)
  cudaStream_t st = as_stream(stream);
  if (cudaMemsetAsync(idx, 0x7f, sizeof(int) * 2 * (size_t)planes, st) != cudaSuccess) { set_error("risp_plane_argfirst: memset failed"); return RISP_E_CUDA; }
  plane_argfirst_kernel<<<dim3(stat_blocks(planes, HW), planes), kT, 0, st>>>(x, stats, idx, HW);
  return check_launch("plane_argfirst");
}

// dx (planes, HW) = gradient of [min, mean, max] per plane: g (planes, 3), idx from risp_plane_argfirst
extern "C" int risp_plane_stats_bwd(const float* g, const int* idx, float* dx, int planes, long long HW, risp_stream_t stream) {
  RISP_REQUIRE(g && idx && dx && planes > 0 && planes <= 65535 && HW > 0 && HW < 0x7fffffffll, RISP_E_INVALID,
               "risp_plane_stats_bwd: bad arguments");
  plane_stats_bwd_kernel<<<dim3(stat_blocks(planes, HW), planes), kT, 0, as_stream(stream)>>>(g, idx, dx, HW, (float)(1.0 / (double)HW));
  return check_launch("plane_stats_bwd");
}

extern "C" int risp_loglum_mean(const float* x, float* out, int N, long long HW, float scale, void* workspace,
                                size_t workspace_bytes, risp_stream_t stream) {
  RISP_REQUIRE(x && out && N > 0 && HW > 0 && N <= 65535, RISP_E_INVALID, "risp_loglum_mean: bad arguments");
  RISP_REQUIRE(workspace && workspace_bytes >= risp_plane_stats_workspace(N, HW), RISP_E_WORKSPACE,
               "risp_loglum_mean: workspace too small (size it with risp_plane_stats_workspace(N, HW))");
  cudaStream_t st = as_stream(stream);
  int B = stat_blocks(N, HW);
  float* partial = static_cast<float*>(workspace);
  loglum_kernel<<<dim3(B, N), kT, 0, st>>>(x, partial, HW, scale);
  loglum_final<<<N, 32, 0, st>>>(partial, out, B, (float)(1.0 / (double)HW));
  return check_launch("loglum");
}

extern "C" int risp_histc01(const float* x, float* out, int planes, long long HW, int bins, risp_stream_t stream) {
  RISP_REQUIRE(x && out && planes > 0 && HW > 0 && planes <= 65535, RISP_E_INVALID, "risp_histc01: bad arguments");
  RISP_REQUIRE(bins >= 1 && bins <= 8192, RISP_E_INVALID, "risp_histc01: bins %d not in [1,8192]", bins);
  RISP_REQUIRE(HW < (1ll << 24), RISP_E_UNSUPPORTED, "risp_histc01: planes above 2^24 pixels are not supported");
  cudaStream_t st = as_stream(stream);
  if (cudaMemsetAsync(out, 0, sizeof(float) * (size_t)planes * bins, st) != cudaSuccess) {
    set_error("risp_histc01: memset failed");
    return RISP_E_CUDA;
  }
  int B = stat_blocks(planes, HW);
  histc_kernel<<<dim3(B, planes), kT, bins * sizeof(unsigned int), st>>>(x, out, HW, bins);
  return check_launch("histc_kernel");
}

extern "C" size_t risp_kth_largest_workspace(int planes) { return planes > 0 ? (size_t)planes * sizeof(SelState) : 0; }

extern "C" int risp_kth_largest(const float* x, const long long* k, float* out, int planes, long long HW,
                                void* workspace, size_t workspace_bytes, risp_stream_t stream) {
  RISP_REQUIRE(x && k && out && planes > 0 && HW > 0 && planes <= 65535, RISP_E_INVALID, "risp_kth_largest: bad arguments");
  RISP_REQUIRE(HW < (1ll << 32), RISP_E_UNSUPPORTED, "risp_kth_largest: plane too large");
  RISP_REQUIRE(workspace && workspace_bytes >= risp_kth_largest_workspace(planes), RISP_E_WORKSPACE,
               "risp_kth_largest: workspace too small");
  cudaStream_t st = as_stream(stream);
  SelState* state = static_cast<SelState*>(workspace);
  int B = stat_blocks(planes, HW);
  sel_init_kernel<<<planes, 256, 0, st>>>(state, k, HW);
  for (int pass = 0; pass < 4; ++pass) {
    sel_hist_kernel<<<dim3(B, planes), kT, 0, st>>>(x, state, HW, pass);
    sel_pick_kernel<<<planes, 256, 0, st>>>(state, out, pass);
  }
  return check_launch("kth_largest");
}
