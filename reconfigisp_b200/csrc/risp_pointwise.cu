// Per-pixel stage chain on BGR planes: risp_chain_fwd / risp_chain_bwd.
//
// HBM-bound (24 B/px forward, 36 B/px backward for any chain length): every thread streams
// 128-bit vectors of the three planes through registers, two vectors in flight per thread; the
// grid is a multiple of the SM count and grid-strides over the image.  Parameter gradients are
// accumulated in registers, reduced warp-shuffle -> shared memory -> one partial row per block, and
// finished by the deterministic finaliser (no float atomics).
#include "risp_common.cuh"
#include "risp_stage.cuh"

namespace risp {

constexpr int kThreads = 256;

template <int VEC> struct Vec;
template <> struct Vec<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) { float4 t = ld_stream4(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  __device__ __forceinline__ void store(float* p) const { st_stream4(p, make_float4(v[0], v[1], v[2], v[3])); }
};
template <> struct Vec<1> {
  float v[1];
  __device__ __forceinline__ void load(const float* p) { v[0] = ld_stream1(p); }
  __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
};

template <int VEC, int SMAX>
__global__ void __launch_bounds__(kThreads, 2)
chain_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long HW, ChainDesc d,
                 const float* __restrict__ params, int pstride, float in_scale, float out_scale) {
  const int n = blockIdx.y;
  const float* __restrict__ prow = params + (long long)n * pstride;
  const float* xb = x + (long long)n * 3 * HW;
  float* yb = y + (long long)n * 3 * HW;
  const long long nvec = HW / VEC;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += 2 * stride) {
    Vec<VEC> B[2], G[2], R[2];
    bool has2 = (i + stride) < nvec;
    B[0].load(xb + i * VEC); G[0].load(xb + HW + i * VEC); R[0].load(xb + 2 * HW + i * VEC);
    if (has2) {
      long long j = i + stride;
      B[1].load(xb + j * VEC); G[1].load(xb + HW + j * VEC); R[1].load(xb + 2 * HW + j * VEC);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !has2) break;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        float b = B[u].v[k] * in_scale, g = G[u].v[k] * in_scale, r = R[u].v[k] * in_scale;
#pragma unroll
        for (int s = 0; s < SMAX; ++s)
          if (s < d.S) stage_fwd(d.op[s], d.iarg[s], prow + d.off[s], b, g, r);
        B[u].v[k] = b * out_scale; G[u].v[k] = g * out_scale; R[u].v[k] = r * out_scale;
      }
      long long j = i + u * stride;
      B[u].store(yb + j * VEC); G[u].store(yb + HW + j * VEC); R[u].store(yb + 2 * HW + j * VEC);
    }
  }
}

// block-wide reduction of the per-thread accumulators into partial[(n*B + blockIdx.x)*NSLOT + slot]
template <bool BIG, int SMAX>
__device__ __forceinline__ void flush_accumulators(float (&accS)[SMAX][RISP_SMALL_ACC],
                                                   float (&accB)[RISP_BIG_ACC], float extra,
                                                   float* __restrict__ prow_out, int S, bool has_extra) {
  __shared__ float red[kThreads / 32][RISP_NSLOT];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < SMAX; ++s) {
    if (s < S) {
#pragma unroll
      for (int j = 0; j < RISP_SMALL_ACC; ++j) {
        float v = warp_sum(accS[s][j]);
        if (lane == 0) red[wid][s * RISP_SMALL_ACC + j] = v;
      }
    }
  }
  if (BIG) {
#pragma unroll
    for (int k = 0; k < RISP_BIG_ACC; ++k) {
      float v = warp_sum(accB[k]);
      if (lane == 0) red[wid][RISP_SLOT_BIG + k] = v;
    }
  }
  if (has_extra) {
    float v = warp_sum(extra);
    if (lane == 0) red[wid][RISP_SLOT_LOSS] = v;
  }
  __syncthreads();
  for (int slot = threadIdx.x; slot < RISP_NSLOT; slot += kThreads) {
    bool live = (slot < RISP_SLOT_BIG) ? (slot / RISP_SMALL_ACC < S)
                                       : (slot < RISP_SLOT_LOSS ? BIG : (slot == RISP_SLOT_LOSS && has_extra));
    float v = 0.f;
    if (live)
      for (int w = 0; w < kThreads / 32; ++w) v += red[w][slot];
    prow_out[slot] = v;
  }
}

template <int VEC, bool BIG, int SMAX>
__global__ void __launch_bounds__(kThreads, (BIG && SMAX > 3) ? 1 : 2)
chain_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                 float* __restrict__ partial, long long HW, ChainDesc d, const float* __restrict__ params,
                 int pstride) {
  const int n = blockIdx.y;
  const float* __restrict__ prow = params + (long long)n * pstride;
  const float* xb = x + (long long)n * 3 * HW;
  const float* gb = dy + (long long)n * 3 * HW;
  float* ob = dx ? dx + (long long)n * 3 * HW : nullptr;
  float accS[SMAX][RISP_SMALL_ACC];
  float accB[RISP_BIG_ACC];
#pragma unroll
  for (int s = 0; s < SMAX; ++s)
#pragma unroll
    for (int j = 0; j < RISP_SMALL_ACC; ++j) accS[s][j] = 0.f;
#pragma unroll
  for (int k = 0; k < RISP_BIG_ACC; ++k) accB[k] = 0.f;

  const long long nvec = HW / VEC;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += stride) {
    Vec<VEC> B, G, R, DB, DG, DR;
    B.load(xb + i * VEC); G.load(xb + HW + i * VEC); R.load(xb + 2 * HW + i * VEC);
    DB.load(gb + i * VEC); DG.load(gb + HW + i * VEC); DR.load(gb + 2 * HW + i * VEC);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float sb[SMAX], sg[SMAX], sr[SMAX];
      float b = B.v[k], g = G.v[k], r = R.v[k];
#pragma unroll
      for (int s = 0; s < SMAX; ++s) {
        if (s < d.S) {
          sb[s] = b; sg[s] = g; sr[s] = r;
          if (s + 1 < d.S) stage_fwd(d.op[s], d.iarg[s], prow + d.off[s], b, g, r);
        }
      }
      float db = DB.v[k], dg = DG.v[k], dr = DR.v[k];
#pragma unroll
      for (int s = SMAX - 1; s >= 0; --s)
        if (s < d.S)
          stage_bwd<BIG>(d.op[s], d.iarg[s], prow + d.off[s], sb[s], sg[s], sr[s], db, dg, dr, accS[s], accB);
      DB.v[k] = db; DG.v[k] = dg; DR.v[k] = dr;
    }
    if (ob) { DB.store(ob + i * VEC); DG.store(ob + HW + i * VEC); DR.store(ob + 2 * HW + i * VEC); }
  }
  flush_accumulators<BIG, SMAX>(accS, accB, 0.f, partial + ((long long)n * gridDim.x + blockIdx.x) * RISP_NSLOT, d.S,
                          false);
}

static int pick_grid(long long nvec, int N, int per_thread) {
  long long want = cdiv(nvec, (long long)kThreads * per_thread);
  long long cap = (long long)sm_count() * 8 / (N > 0 ? N : 1);
  if (cap < sm_count() / 2) cap = sm_count() / 2;   // large batches: still give every image several blocks
  if (cap < 1) cap = 1;
  long long g = want < cap ? want : cap;
  return (int)(g < 1 ? 1 : g);
}

static int chain_bwd_blocks(int N, long long HW) {
  long long nvec = (HW % 4 == 0) ? HW / 4 : HW;
  return pick_grid(nvec, N, 1);
}

}  // namespace risp

using namespace risp;

extern "C" int risp_chain_fwd(const float* x, float* y, int N, long long HW, const int* ops, const int* param_off,
                              const int* iarg, int S, const float* params, int param_stride, float in_scale,
                              float out_scale, risp_stream_t stream) {
  RISP_REQUIRE(x && y && N > 0 && HW > 0, RISP_E_INVALID, "risp_chain_fwd: bad tensor arguments");
  ChainDesc d;
  int P = 0;
  int rc = make_chain(&d, ops, param_off, iarg, S, &P);
  if (rc != RISP_OK) return rc;
  RISP_REQUIRE(P == 0 || params, RISP_E_INVALID, "risp_chain_fwd: chain needs %d parameters but params is null", P);
  RISP_REQUIRE(param_stride == 0 || param_stride >= P, RISP_E_INVALID, "risp_chain_fwd: param_stride %d < %d", param_stride, P);
  RISP_REQUIRE(N <= 65535, RISP_E_INVALID, "risp_chain_fwd: batch %d > 65535", N);
  bool vec = (HW % 4 == 0) && aligned16(x) && aligned16(y);
  cudaStream_t st = as_stream(stream);
#define LAUNCH(V, SM) chain_fwd_kernel<V, SM><<<grid, kThreads, 0, st>>>(x, y, HW, d, params, param_stride, in_scale, out_scale)
  if (vec) {
    dim3 grid(pick_grid(HW / 4, N, 2), N);
    if (S <= 1) LAUNCH(4, 1); else if (S <= 3) LAUNCH(4, 3); else LAUNCH(4, RISP_MAX_STAGES);
  } else {
    dim3 grid(pick_grid(HW, N, 2), N);
    if (S <= 1) LAUNCH(1, 1); else LAUNCH(1, RISP_MAX_STAGES);
  }
#undef LAUNCH
  return check_launch("chain_fwd_kernel");
}

extern "C" size_t risp_chain_bwd_workspace(int N, long long HW, int P) {
  (void)P;
  if (N <= 0 || HW <= 0) return 0;
  return (size_t)N * chain_bwd_blocks(N, HW) * RISP_NSLOT * sizeof(float);
}

extern "C" int risp_chain_bwd(const float* x, const float* dy, float* dx, float* dparams, int N, long long HW,
                              const int* ops, const int* param_off, const int* iarg, int S, const float* params,
                              int param_stride, int P, void* workspace, size_t workspace_bytes,
                              risp_stream_t stream) {
  RISP_REQUIRE(x && dy && N > 0 && HW > 0, RISP_E_INVALID, "risp_chain_bwd: bad tensor arguments");
  ChainDesc d;
  int Pn = 0;
  int rc = make_chain(&d, ops, param_off, iarg, S, &Pn);
  if (rc != RISP_OK) return rc;
  bool big = false;
  for (int s = 0; s < S; ++s) {
    RISP_REQUIRE(op_has_bwd(ops[s]), RISP_E_UNSUPPORTED, "risp_chain_bwd: op %d is forward-only (non-differentiable Origin* stage)", ops[s]);
    big = big || op_is_big(ops[s]);
  }
  RISP_REQUIRE(Pn <= P, RISP_E_INVALID, "risp_chain_bwd: chain needs %d parameters, P=%d", Pn, P);
  RISP_REQUIRE(Pn == 0 || (params && dparams), RISP_E_INVALID, "risp_chain_bwd: null params/dparams");
  RISP_REQUIRE(param_stride == 0 || param_stride >= P, RISP_E_INVALID, "risp_chain_bwd: param_stride %d < P %d", param_stride, P);
  RISP_REQUIRE(N <= 65535, RISP_E_INVALID, "risp_chain_bwd: batch %d > 65535", N);
  RISP_REQUIRE(workspace && workspace_bytes >= risp_chain_bwd_workspace(N, HW, P), RISP_E_WORKSPACE,
               "risp_chain_bwd: workspace %zu < %zu", workspace_bytes, risp_chain_bwd_workspace(N, HW, P));
  cudaStream_t st = as_stream(stream);
  bool vec = (HW % 4 == 0) && aligned16(x) && aligned16(dy) && (!dx || aligned16(dx));
  int B = chain_bwd_blocks(N, HW);
  dim3 grid(B, N);
  float* partial = static_cast<float*>(workspace);
#define LAUNCH(V, BG, SM) chain_bwd_kernel<V, BG, SM><<<grid, kThreads, 0, st>>>(x, dy, dx, partial, HW, d, params, param_stride)
#define LAUNCH_S(V, BG) do { if (S <= 1) LAUNCH(V, BG, 1); else if (S <= 3) LAUNCH(V, BG, 3); else LAUNCH(V, BG, RISP_MAX_STAGES); } while (0)
  if (vec) { if (big) LAUNCH_S(4, true); else LAUNCH_S(4, false); }
  else     { if (big) LAUNCH(1, true, RISP_MAX_STAGES); else LAUNCH(1, false, RISP_MAX_STAGES); }
#undef LAUNCH_S
#undef LAUNCH
  rc = check_launch("chain_bwd_kernel");
  if (rc != RISP_OK) return rc;
  if (P > 0) {
    bool shared_row = (param_stride == 0);
    rc = cudaMemsetAsync(dparams, 0, sizeof(float) * (size_t)P * (shared_row ? 1 : N), st) == cudaSuccess ? RISP_OK : RISP_E_CUDA;
    if (rc != RISP_OK) { set_error("risp_chain_bwd: memset failed"); return rc; }
    SlotList m;
    chain_slot_list(d, &m);
    rc = finalize_partials(partial, dparams, N, B, RISP_NSLOT, P, m.dst, m.slot, m.n, 1.f, shared_row, st);
  }
  return rc;
}
