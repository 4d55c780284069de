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

// A thread's pixel group: G vectors of VEC pixels from each plane (NPX = G*VEC, kept even for the
// packed fp32 math; the scalar path pairs neighbouring pixels and masks the odd tail).
template <int VEC, int G>
struct Group {
  static constexpr int NPX = (VEC == 1) ? 2 * G : VEC * G;
  Px<NPX> px;
  // element index of pixel k for vector slot i0 (+ g*stride)
  __device__ __forceinline__ void load(const float* __restrict__ base, long long HW, long long i0, long long stride,
                                       long long nvec, float scale) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const long long i = i0 + g * stride;
      if (VEC == 4) {
        if (i < nvec) {
          const float4 b = ld_stream4(base + 4 * i), gg = ld_stream4(base + HW + 4 * i), r = ld_stream4(base + 2 * HW + 4 * i);
          px.b[4 * g] = b.x * scale; px.b[4 * g + 1] = b.y * scale; px.b[4 * g + 2] = b.z * scale; px.b[4 * g + 3] = b.w * scale;
          px.g[4 * g] = gg.x * scale; px.g[4 * g + 1] = gg.y * scale; px.g[4 * g + 2] = gg.z * scale; px.g[4 * g + 3] = gg.w * scale;
          px.r[4 * g] = r.x * scale; px.r[4 * g + 1] = r.y * scale; px.r[4 * g + 2] = r.z * scale; px.r[4 * g + 3] = r.w * scale;
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) { px.b[4 * g + k] = 0.f; px.g[4 * g + k] = 0.f; px.r[4 * g + k] = 0.f; }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const long long e = 2 * i + k;
          const bool ok = (i < nvec) && (e < HW);
          px.b[2 * g + k] = ok ? base[e] * scale : 0.f;
          px.g[2 * g + k] = ok ? base[HW + e] * scale : 0.f;
          px.r[2 * g + k] = ok ? base[2 * HW + e] * scale : 0.f;
        }
      }
    }
  }
  __device__ __forceinline__ void store(float* __restrict__ base, long long HW, long long i0, long long stride,
                                        long long nvec, float scale) const {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const long long i = i0 + g * stride;
      if (i >= nvec) continue;
      if (VEC == 4) {
        st_stream4(base + 4 * i, make_float4(px.b[4 * g] * scale, px.b[4 * g + 1] * scale, px.b[4 * g + 2] * scale, px.b[4 * g + 3] * scale));
        st_stream4(base + HW + 4 * i, make_float4(px.g[4 * g] * scale, px.g[4 * g + 1] * scale, px.g[4 * g + 2] * scale, px.g[4 * g + 3] * scale));
        st_stream4(base + 2 * HW + 4 * i, make_float4(px.r[4 * g] * scale, px.r[4 * g + 1] * scale, px.r[4 * g + 2] * scale, px.r[4 * g + 3] * scale));
      } else {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const long long e = 2 * i + k;
          if (e < HW) { base[e] = px.b[2 * g + k] * scale; base[HW + e] = px.g[2 * g + k] * scale; base[2 * HW + e] = px.r[2 * g + k] * scale; }
        }
      }
    }
  }
};

// number of vector slots of a plane: VEC==4 -> HW/4 float4s ; VEC==1 -> ceil(HW/2) pixel pairs
template <int VEC> __device__ __host__ __forceinline__ long long vec_slots(long long HW) { return VEC == 4 ? HW / 4 : (HW + 1) / 2; }

template <int VEC, unsigned SIG>
__global__ void __launch_bounds__(kThreads, 2)
chain_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long HW, ChainDesc d,
                 const float* __restrict__ params, int pstride, float in_scale, float out_scale) {
  using SG = Sig<SIG>;
  constexpr int SMAX = SG::S;
  const int n = blockIdx.y;
  const float* __restrict__ prow = params + (long long)n * pstride;
  const float* xb = x + (long long)n * 3 * HW;
  float* yb = y + (long long)n * 3 * HW;
  const long long nvec = vec_slots<VEC>(HW);
  const long long stride = (long long)gridDim.x * kThreads;
  constexpr int G = 2;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += G * stride) {
    Group<VEC, G> grp;
    grp.load(xb, HW, i, stride, nvec, in_scale);
#pragma unroll
    for (int s = 0; s < SMAX; ++s)
      if (SG::live(d, s)) stage_fwd(SG::op(d, s), SG::iarg(d, s), prow + d.off[s], grp.px);
    grp.store(yb, HW, i, stride, nvec, out_scale);
  }
}

// block-wide reduction of the per-thread accumulators into partial[(n*B + blockIdx.x)*NSLOT + slot]
template <bool BIG, int SMAX>
__device__ __forceinline__ void flush_accumulators(float (&accS)[SMAX][RISP_SMALL_ACC],
                                                   float2 (&accB)[RISP_BIG_ACC], float extra,
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
      float v0 = warp_sum(accB[k].x), v1 = warp_sum(accB[k].y);
      if (lane == 0) { red[wid][RISP_SLOT_BIG + 2 * k] = v0; red[wid][RISP_SLOT_BIG + 2 * k + 1] = v1; }
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

template <int VEC, unsigned SIG, bool BIGG>
__global__ void __launch_bounds__(kThreads, (SIG == 0 && BIGG) ? 1 : 2)
chain_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                 float* __restrict__ partial, long long HW, ChainDesc d, const float* __restrict__ params,
                 int pstride) {
  using SG = Sig<SIG>;
  constexpr int SMAX = SG::S;
  constexpr bool BIG = SG::generic ? BIGG : SG::big_c();
  const int n = blockIdx.y;
  const float* __restrict__ prow = params + (long long)n * pstride;
  const float* xb = x + (long long)n * 3 * HW;
  const float* gb = dy + (long long)n * 3 * HW;
  float* ob = dx ? dx + (long long)n * 3 * HW : nullptr;
  float accS[SMAX][RISP_SMALL_ACC];
  float2 accB[RISP_BIG_ACC];
#pragma unroll
  for (int s = 0; s < SMAX; ++s)
#pragma unroll
    for (int j = 0; j < RISP_SMALL_ACC; ++j) accS[s][j] = 0.f;
#pragma unroll
  for (int k = 0; k < RISP_BIG_ACC; ++k) accB[k] = make_float2(0.f, 0.f);

  const long long nvec = vec_slots<VEC>(HW);
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < nvec; i += stride) {
    Group<VEC, 1> X, D;
    X.load(xb, HW, i, stride, nvec, 1.f);
    D.load(gb, HW, i, stride, nvec, 1.f);      // out-of-range tail pixels get d = 0
    constexpr int NPX = Group<VEC, 1>::NPX;
    Px<NPX> saved[SMAX + 1];        // saved[s] = input of stage s, saved[s+1] = its output
    saved[0] = X.px;
#pragma unroll
    for (int s = 0; s < SMAX; ++s) {
      if (SG::live(d, s)) {
        stage_fwd(SG::op(d, s), SG::iarg(d, s), prow + d.off[s], X.px);
        saved[s + 1] = X.px;
      }
    }
#pragma unroll
    for (int s = SMAX - 1; s >= 0; --s)
      if (SG::live(d, s))
        stage_bwd<NPX, BIG>(SG::op(d, s), SG::iarg(d, s), prow + d.off[s], saved[s], saved[s + 1], D.px, accS[s], accB);
    if (ob) D.store(ob, HW, i, stride, nvec, 1.f);
  }
  flush_accumulators<BIG, SMAX>(accS, accB, 0.f, partial + ((long long)n * gridDim.x + blockIdx.x) * RISP_NSLOT,
                                SG::generic ? d.S : SMAX, false);
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
  long long nvec = (HW % 4 == 0) ? HW / 4 : (HW + 1) / 2;
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
  const unsigned sig = chain_signature(d);
  bool done = false;
  if (vec) {
    dim3 grid(pick_grid(HW / 4, N, 2), N);
#define RISP_TRY(SG)                                                                                                 \
    if (!done && sig == (SG)) {                                                                                      \
      chain_fwd_kernel<4, (SG)><<<grid, kThreads, 0, st>>>(x, y, HW, d, params, param_stride, in_scale, out_scale); \
      done = true;                                                                                                   \
    }
    RISP_FOR_EACH_SINGLE_SIG(RISP_TRY)
    RISP_FOR_EACH_CHAIN_SIG(RISP_TRY)
#undef RISP_TRY
    if (!done) chain_fwd_kernel<4, 0u><<<grid, kThreads, 0, st>>>(x, y, HW, d, params, param_stride, in_scale, out_scale);
  } else {
    dim3 grid(pick_grid((HW + 1) / 2, N, 2), N);
    chain_fwd_kernel<1, 0u><<<grid, kThreads, 0, st>>>(x, y, HW, d, params, param_stride, in_scale, out_scale);
  }
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
  const unsigned sig = chain_signature(d);
  bool done = false;
  if (vec) {
#define RISP_TRY(SG)                                                                                          \
    if (!done && sig == (SG)) {                                                                               \
      chain_bwd_kernel<4, (SG), false><<<grid, kThreads, 0, st>>>(x, dy, dx, partial, HW, d, params, param_stride); \
      done = true;                                                                                            \
    }
    RISP_FOR_EACH_SINGLE_SIG(RISP_TRY)
    RISP_FOR_EACH_CHAIN_SIG(RISP_TRY)
#undef RISP_TRY
    if (!done) {
      if (big) chain_bwd_kernel<4, 0u, true><<<grid, kThreads, 0, st>>>(x, dy, dx, partial, HW, d, params, param_stride);
      else chain_bwd_kernel<4, 0u, false><<<grid, kThreads, 0, st>>>(x, dy, dx, partial, HW, d, params, param_stride);
    }
  } else {
    if (big) chain_bwd_kernel<1, 0u, true><<<grid, kThreads, 0, st>>>(x, dy, dx, partial, HW, d, params, param_stride);
    else chain_bwd_kernel<1, 0u, false><<<grid, kThreads, 0, st>>>(x, dy, dx, partial, HW, d, params, param_stride);
  }
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

// ---- parameter table of a fixed pipeline: kernel-level values from the trainable logits, and back -----------------------
// The wrappers map sigmoid(logit) affinely into each stage's range (gain = p*5 tools_origin.py:214, P = p*10-5 :326, gamma
// and tone-curve knots = p).  One launch builds the whole (1,P) table, one launch turns d loss / d table into d loss / d logits:
// the proxy-tuning step then needs no autograd graph at all (isp_model.py:128-142 as 7 launches instead of ~25).
namespace risp {
__global__ void param_table_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ a, const float* __restrict__ b,
                                       float* __restrict__ table, int P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) table[i] = fmaf(a[i], 1.f / (1.f + expf(-logits[i])), b[i]);
}
__global__ void param_table_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ a,
                                       const float* __restrict__ dtable, float* __restrict__ dlogits, int P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) {
    const float s = 1.f / (1.f + expf(-logits[i]));
    dlogits[i] = dtable[i] * a[i] * s * (1.f - s);
  }
}
}  // namespace risp

extern "C" int risp_param_table_fwd(const float* logits, const float* a, const float* b, float* table, int P,
                                    risp_stream_t stream) {
  RISP_REQUIRE(logits && a && b && table && P > 0, RISP_E_INVALID, "risp_param_table_fwd: bad arguments");
  param_table_fwd_kernel<<<(int)cdiv(P, 128), 128, 0, as_stream(stream)>>>(logits, a, b, table, P);
  return check_launch("param_table_fwd_kernel");
}

extern "C" int risp_param_table_bwd(const float* logits, const float* a, const float* dtable, float* dlogits, int P,
                                    risp_stream_t stream) {
  RISP_REQUIRE(logits && a && dtable && dlogits && P > 0, RISP_E_INVALID, "risp_param_table_bwd: bad arguments");
  param_table_bwd_kernel<<<(int)cdiv(P, 128), 128, 0, as_stream(stream)>>>(logits, a, dtable, dlogits, P);
  return check_launch("param_table_bwd_kernel");
}
