// Fused raw -> BGR pipeline on packed fp32 with constant-bank parameters: the kernels behind
// risp_pipeline_fwd / risp_pipeline_mse_step / risp_pipeline_bwd for the pre-instantiated chain signatures.
//
// Reference: isp_universal.py:210-232 / origin_universal.py:143-161 (one full-image pass per stage) and the
// proxy-tuning step isp_model.py:128-142 (forward, MSELoss, backward).  Here: one pass, 16 B/px.
//
// Shape of the kernel:
//   * persistent CTAs (one wave: SMs x resident CTAs), each walking a static list of (row chunk, 512-column block)
//     items of ONE frame -> accumulators live in registers for the whole kernel, one reduction per warp, and the
//     summation order is fixed (bit-reproducible gradients);
//   * a warp owns a 128-column strip and marches down the rows with the demosaic window in registers (horizontal halo
//     by shuffle), two rows per loop trip so the CFA row parity is a compile-time constant;
//   * the chain runs on pixel PAIRS with FFMA2/FMUL2/FADD2, coefficients from uniform registers (risp_fused.cuh).
#include "risp_fused.cuh"

#include <mutex>

namespace risp {
namespace fused {

__constant__ float g_cpar[kCSlots][kCRows][kCRowFloats];

enum { MODE_FWD = 0, MODE_STEP = 1, MODE_BWD = 2 };
constexpr int kWarps = 4;
constexpr int kStrip = 128;

struct FusedArgs {
  const float* raw;     // (N,1,H,W)
  const float* gt;      // (N,3,H,W): target (STEP) or dL/dy (BWD)
  float* y;             // (N,3,H,W), nullable except in FWD
  float* partial;       // [N][cpf*kWarps][RISP_NSLOT]
  const float* params;  // raw parameter rows (epilogue of the folded gain)
  int pstride;
  int H, W;
  int rows_per_chunk, chunks, strip_blocks, cpf;   // cpf = CTAs per frame
  float clip_hi;
  int slot;
};

// ---- preparation: derived constants of every parameter row ---------------------------------------------------------
__global__ void fused_prep_kernel(const float* __restrict__ params, int pstride, ChainDesc d, float* __restrict__ cslot) {
  if (threadIdx.x != 0) return;
  const float* p = params + (long long)blockIdx.x * pstride;
  float* c = cslot + (long long)blockIdx.x * kCRowFloats;
  const EffChain e = eff_from_ops(d.op, d.iarg, d.S);
  for (int k = 0; k < e.n; ++k) {
    float* o = c + e.coff[k];
    const float* q = p + d.off[e.src[k]];
    switch (e.op[k]) {
      case RISP_OP_GAMMA: o[0] = q[0]; o[1] = 0.f; break;
      case RISP_OP_GAIN: o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = 0.f; break;
      case RISP_OP_POLY10:
      case FOP_POLYG: {
        float gm[10];
        if (e.op[k] == FOP_POLYG) {
          const float* g = p + d.off[e.src[k] - 1];
          const float gb = g[0], gg = g[1], gr = g[2];
          gm[0] = gb * gb; gm[1] = gg * gg; gm[2] = gr * gr; gm[3] = gb * gg; gm[4] = gb * gr; gm[5] = gg * gr;
          gm[6] = gb; gm[7] = gg; gm[8] = gr; gm[9] = 1.f;
        } else {
          for (int i = 0; i < 10; ++i) gm[i] = 1.f;
        }
        for (int cc = 0; cc < 3; ++cc)
          for (int i = 0; i < 10; ++i) o[cc * 10 + i] = q[cc * 10 + i] * gm[i];
        for (int cc = 0; cc < 3; ++cc)
          for (int i = 0; i < 3; ++i) o[30 + cc * 3 + i] = 2.f * o[cc * 10 + i];
        o[39] = 0.f;
      } break;
      case RISP_OP_GTM: {
        // knots y_0 = 0, y_k = q[k-1], y_4 = 1;  s_k = (y_{k+1} - y_k) * 4
        const float y1 = q[0], y2 = q[1], y3 = q[2];
        const float s0 = (y1 - 0.f) * 4.f, s1 = (y2 - y1) * 4.f, s2 = (y3 - y2) * 4.f, s3 = (1.f - y3) * 4.f;
        o[0] = s0; o[1] = s1 - s0; o[2] = s2 - s1; o[3] = s3 - s2;
        const float s3acc = ((s0 + o[1]) + o[2]) + o[3];     // the slope the kernel accumulates on the last segment
        o[4] = 1.f - s3acc;
        const bool in_range = (y1 >= 0.f && y1 <= 1.f) && (y2 >= 0.f && y2 <= 1.f) && (y3 >= 0.f && y3 <= 1.f);
        o[5] = in_range ? 0.f : 1.f;
        o[6] = 0.f; o[7] = 0.f;
      } break;
      default: break;
    }
  }
}

// ---- raw rows: issue / finish with the horizontal halo by shuffle ------------------------------------------------------
// A lane loads its 4 columns with one 128-bit load.  The horizontal halo comes from the neighbouring lanes by shuffle;
// only the lanes at the two ends of the strip load it from memory, and the reflect-101 border is folded into the ADDRESS of
// that load (column -1 -> 1, column W -> W-2), so finishing a row needs no border logic at all.
template <int HL> struct RawRow;
template <> struct RawRow<1> { float4 v; float l, r; };
template <> struct RawRow<2> { float4 v; float2 l, r; };

struct LaneGeom {       // per item, per lane
  bool active, endL, endR;   // inside the frame; loads the left / right halo itself
  int offL, offR;            // element offset of that load relative to the lane's first column
};

template <int HL>
__device__ __forceinline__ LaneGeom lane_geom(int c0, int W, int lane) {
  LaneGeom g;
  g.active = c0 < W;
  const bool first = (c0 == 0), last = g.active && (c0 + 4 >= W);
  g.endL = g.active && lane == 0;
  g.endR = g.active && (lane == 31 || last);
  if (HL == 1) { g.offL = first ? 1 : -1; g.offR = last ? 2 : 4; }
  else         { g.offL = first ? 1 : -2; g.offR = last ? 1 : 4; }   // float2 loads: cols (1,2) reversed / (W-3,W-2) reversed
  return g;
}

__device__ __forceinline__ int reflect101(int r, int H) { return r < 0 ? -r : (r >= H ? 2 * H - 2 - r : r); }

// issue the loads of one raw row (rp = the lane's first column of that row) -- no dependent instruction here
template <int HL>
__device__ __forceinline__ void row_issue(RawRow<HL>& q, const float* __restrict__ rp, const LaneGeom& g) {
  q.v = g.active ? __ldg(reinterpret_cast<const float4*>(rp)) : make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (HL == 1) {
    q.l = 0.f; q.r = 0.f;
    if (g.endL) q.l = __ldg(rp + g.offL);
    if (g.endR) q.r = __ldg(rp + g.offR);
  } else {
    q.l = make_float2(0.f, 0.f); q.r = make_float2(0.f, 0.f);
    if (g.endL) { q.l.x = __ldg(rp + g.offL); q.l.y = __ldg(rp + g.offL + 1); }
    if (g.endR) { q.r.x = __ldg(rp + g.offR); q.r.y = __ldg(rp + g.offR + 1); }
  }
}

// halo exchange: dst[0 .. 4+2*HL) = columns c0-HL .. c0+3+HL of the row
template <int HL>
__device__ __forceinline__ void row_finish(float (&dst)[4 + 2 * HL], const RawRow<HL>& q, int W, int c0, const LaneGeom& g) {
  const unsigned full = 0xffffffffu;
  const float4 v = q.v;
  const bool first = (c0 == 0), last = g.active && (c0 + 4 >= W);
  if constexpr (HL == 1) {
    float l = __shfl_up_sync(full, v.w, 1), r = __shfl_down_sync(full, v.x, 1);
    if (g.endL) l = q.l;
    if (g.endR) r = q.r;
    dst[0] = l; dst[1] = v.x; dst[2] = v.y; dst[3] = v.z; dst[4] = v.w; dst[5] = r;
  } else {
    float l0 = __shfl_up_sync(full, v.z, 1), l1 = __shfl_up_sync(full, v.w, 1);
    float r0 = __shfl_down_sync(full, v.x, 1), r1 = __shfl_down_sync(full, v.y, 1);
    // interior strip ends load (c0-2, c0-1) / (c0+4, c0+5); at the frame border the pair is the reflected one, reversed:
    // cols (-2,-1) = cols (2,1), loaded as (1,2);  cols (W, W+1) = cols (W-2, W-3), loaded as (W-3, W-2)
    if (g.endL) { l0 = first ? q.l.y : q.l.x; l1 = first ? q.l.x : q.l.y; }
    if (g.endR) { r0 = last ? q.r.y : q.r.x; r1 = last ? q.r.x : q.r.y; }
    dst[0] = l0; dst[1] = l1; dst[2] = v.x; dst[3] = v.y; dst[4] = v.z; dst[5] = v.w; dst[6] = r0; dst[7] = r1;
  }
}

// ---- demosaic of the lane's 4 pixels; row parity is a template parameter (no selects) ---------------------------------
// w[RO + j][i]: raw row r-HL+j, column c0-HL+i.  Sites: (even,even)=R (even,odd)=G1 (odd,even)=G2 (odd,odd)=B.
template <int DM, int HL, int RO, bool ODD, int WRT>
__device__ __forceinline__ void demosaic4(const float (&w)[WRT][4 + 2 * HL], float clip_hi, P2& lo, P2& hi) {
  float B[4], G[4], R[4];
  constexpr int M = RO + HL;     // window row of the pixel's own raw row
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = HL + k;
    const bool xo = (k & 1);
    const float c = w[M][x];
    if constexpr (DM == RISP_DM_NEAREST) {
      if (!ODD) { R[k] = xo ? w[M][x - 1] : c; G[k] = xo ? c : w[M][x + 1]; B[k] = xo ? w[M + 1][x] : w[M + 1][x + 1]; }
      else      { R[k] = xo ? w[M - 1][x - 1] : w[M - 1][x]; G[k] = xo ? w[M][x - 1] : c; B[k] = xo ? c : w[M][x + 1]; }
    } else if constexpr (DM == RISP_DM_BILINEAR) {
      const bool at_g = (ODD != xo);
      if (at_g) {
        const float hor = 0.5f * (w[M][x - 1] + w[M][x + 1]);
        const float ver = 0.5f * (w[M - 1][x] + w[M + 1][x]);
        G[k] = c;
        if (!ODD) { R[k] = hor; B[k] = ver; } else { R[k] = ver; B[k] = hor; }
      } else {
        const float cross = (w[M - 1][x] + w[M + 1][x]) + (w[M][x - 1] + w[M][x + 1]);
        const float diag = (w[M - 1][x - 1] + w[M - 1][x + 1]) + (w[M + 1][x - 1] + w[M + 1][x + 1]);
        G[k] = 0.25f * cross;
        if (!ODD) { R[k] = c; B[k] = 0.25f * diag; } else { R[k] = 0.25f * diag; B[k] = c; }
      }
    } else {  // Malvar-He-Cutler 5x5
      const float n1 = w[M - 1][x], s1 = w[M + 1][x], e1 = w[M][x + 1], w1 = w[M][x - 1];
      const float n2 = w[M - 2][x], s2 = w[M + 2][x], e2 = w[M][x + 2], w2 = w[M][x - 2];
      const float dg = (w[M - 1][x - 1] + w[M - 1][x + 1]) + (w[M + 1][x - 1] + w[M + 1][x + 1]);
      float rr, gg, bb;
      const bool at_g = (ODD != xo);
      if (at_g) {
        const float f_row = 0.125f * (5.f * c + 4.f * (e1 + w1) - dg - (e2 + w2) + 0.5f * (n2 + s2));
        const float f_col = 0.125f * (5.f * c + 4.f * (n1 + s1) - dg - (n2 + s2) + 0.5f * (e2 + w2));
        gg = c;
        if (!ODD) { rr = f_row; bb = f_col; } else { rr = f_col; bb = f_row; }
      } else {
        const float f_g = 0.125f * (4.f * c + 2.f * ((n1 + s1) + (e1 + w1)) - ((n2 + s2) + (e2 + w2)));
        const float f_dg = 0.125f * (6.f * c + 2.f * dg - 1.5f * ((n2 + s2) + (e2 + w2)));
        gg = f_g;
        if (!ODD) { rr = c; bb = f_dg; } else { rr = f_dg; bb = c; }
      }
      R[k] = fminf(fmaxf(rr, 0.f), clip_hi); G[k] = fminf(fmaxf(gg, 0.f), clip_hi); B[k] = fminf(fmaxf(bb, 0.f), clip_hi);
    }
  }
  lo.b = make_float2(B[0], B[1]); lo.g = make_float2(G[0], G[1]); lo.r = make_float2(R[0], R[1]);
  hi.b = make_float2(B[2], B[3]); hi.g = make_float2(G[2], G[3]); hi.r = make_float2(R[2], R[3]);
}

// ---- the chain on one pixel pair: forward, loss, backward (compile-time recursion over the effective stages) ----------
template <unsigned SIG>
__host__ __device__ constexpr bool per_channel_from(int k) {     // stages k.. never mix channels
  constexpr EffChain E = Eff<SIG>::e;
  for (int i = k; i < E.n; ++i)
    if (E.op[i] == RISP_OP_POLY10 || E.op[i] == FOP_POLYG) return false;
  return true;
}

// stages K.. on ONE channel C of the pair (gamma / gain / tone curve): forward, loss term, backward.  Keeping a channel's
// whole tail together makes its saved state transient (a few registers) instead of live across all three channels.
template <unsigned SIG, int K, int MODE, int C>
struct Tail {
  static constexpr EffChain E = Eff<SIG>::e;
  static __device__ __forceinline__ float2 go(float2 x, float2 tgt, const float* __restrict__ cp, float2* acc, float2& loss,
                                              float2& yout, bool slow, float lane_w) {
    if constexpr (K == E.n) {
      yout = x;
      if constexpr (MODE == MODE_BWD) return tgt;
      // d loss / d y up to the constant 2/numel (applied by the finaliser).  lane_w = 1 for lanes inside the frame (1*x - t is
      // exactly x - t) and 0 for the lanes of the last strip that hang over the right edge (their t is 0): no divergent branch
      const float2 dd = __ffma2_rn(make_float2(lane_w, lane_w), x, make_float2(-tgt.x, -tgt.y));
      loss = fma2(dd, dd, loss);
      return dd;
    } else {
      constexpr int OP = E.op[K];
      constexpr bool IN01 = (K > 0) && fop_out01(E.op[K > 0 ? K - 1 : 0]);
      constexpr bool NEED_DX = (K > 0);
      const float* c = cp + E.coff[K];
      float2* a = acc + E.aoff[K];
      if constexpr (OP == RISP_OP_GAMMA) {
        float2 l2;
        const float gm = c[0];
        const float2 y = gamma_fwd2<IN01>(x, gm, l2);
        const float2 d = Tail<SIG, K + 1, MODE, C>::go(y, tgt, cp, acc, loss, yout, slow, lane_w);
        return gamma_bwd2<IN01, NEED_DX, true>(x, y, l2, d, gm, a[0]);
      } else if constexpr (OP == RISP_OP_GAIN) {
        const float2 y = mul2s(x, c[C]);
        const float2 d = Tail<SIG, K + 1, MODE, C>::go(y, tgt, cp, acc, loss, yout, slow, lane_w);
        a[C] = fma2(d, x, a[C]);
        return NEED_DX ? mul2s(d, c[C]) : zero2();
      } else {
        static_assert(OP == RISP_OP_GTM && IN01, "packed path: unsupported per-channel op");
        float2 hd[3], m = zero2();
        const float2 y = gtm_fwd2(x, c, hd, slow, m);
        float2 d = Tail<SIG, K + 1, MODE, C>::go(y, tgt, cp, acc, loss, yout, slow, lane_w);
        if (slow) d = mul2(d, m);
        return gtm_bwd2<NEED_DX>(x, hd, d, c, a);
      }
    }
  }
};

template <unsigned SIG, int K, int MODE>
struct Run {
  static constexpr EffChain E = Eff<SIG>::e;
  static __device__ __forceinline__ P2 go(const P2& x, const P2& tgt, const float* __restrict__ cp, float2* acc, float2& loss,
                                          P2& yout, bool slow, float lane_w) {
    if constexpr (per_channel_from<SIG>(K)) {
      P2 d;
      d.b = Tail<SIG, K, MODE, 0>::go(x.b, tgt.b, cp, acc, loss, yout.b, slow, lane_w);
      d.g = Tail<SIG, K, MODE, 1>::go(x.g, tgt.g, cp, acc, loss, yout.g, slow, lane_w);
      d.r = Tail<SIG, K, MODE, 2>::go(x.r, tgt.r, cp, acc, loss, yout.r, slow, lane_w);
      return d;
    } else {
      constexpr int OP = E.op[K];
      constexpr bool IN01 = (K > 0) && fop_out01(E.op[K > 0 ? K - 1 : 0]);
      constexpr bool NEED_DX = (K > 0);
      const float* c = cp + E.coff[K];
      float2* a = acc + E.aoff[K];
      if constexpr (OP == RISP_OP_GAMMA) {
        GammaSaved sv;
        const float gm = c[0];
        const P2 y = gamma_fwd<IN01>(x, gm, sv);
        const P2 d = Run<SIG, K + 1, MODE>::go(y, tgt, cp, acc, loss, yout, slow, lane_w);
        return gamma_bwd<IN01, NEED_DX, true>(sv, d, gm, a[0]);
      } else if constexpr (OP == RISP_OP_GAIN) {
        GainSaved sv;
        const P2 y = gain_fwd(x, c, sv);
        const P2 d = Run<SIG, K + 1, MODE>::go(y, tgt, cp, acc, loss, yout, slow, lane_w);
        return gain_bwd<NEED_DX>(sv, d, c, a);
      } else if constexpr ((OP == RISP_OP_POLY10 || OP == FOP_POLYG) && !NEED_DX && per_channel_from<SIG>(K + 1)) {
        // channel-major: the monomials are the only cross-channel state; each output channel then runs its polynomial
        // row, its whole tail, the loss and the way back, and accumulates e_c * phi straight away
        float2 phi[9];
        phi[0] = mul2(x.b, x.b); phi[1] = mul2(x.g, x.g); phi[2] = mul2(x.r, x.r);
        phi[3] = mul2(x.b, x.g); phi[4] = mul2(x.b, x.r); phi[5] = mul2(x.g, x.r);
        phi[6] = x.b; phi[7] = x.g; phi[8] = x.r;
        auto channel = [&](auto ctag, float2 tg, float2& yo) {
          constexpr int C = decltype(ctag)::value;
          float2 u = fma2ss(c[C * 10], phi[0], c[C * 10 + 9]);
#pragma unroll
          for (int i = 1; i < 9; ++i) u = fma2s(c[C * 10 + i], phi[i], u);
          const float2 y = sat2(u);
          const float2 d = Tail<SIG, K + 1, MODE, C>::go(y, tg, cp, acc, loss, yo, slow, lane_w);
          const float2 e = sel2(u.x == y.x, u.y == y.y, d);     // clamp mask, inclusive: u in [0,1] <=> sat(u) == u
#pragma unroll
          for (int i = 0; i < 9; ++i) a[C * 10 + i] = fma2(e, phi[i], a[C * 10 + i]);
          a[C * 10 + 9] = add2(a[C * 10 + 9], e);
        };
        channel(std::integral_constant<int, 0>{}, tgt.b, yout.b);
        channel(std::integral_constant<int, 1>{}, tgt.g, yout.g);
        channel(std::integral_constant<int, 2>{}, tgt.r, yout.r);
        P2 z; z.b = zero2(); z.g = zero2(); z.r = zero2();
        return z;
      } else if constexpr (OP == RISP_OP_POLY10 || OP == FOP_POLYG) {
        PolySaved sv;
        const P2 y = poly_fwd(x, c, sv);
        const P2 d = Run<SIG, K + 1, MODE>::go(y, tgt, cp, acc, loss, yout, slow, lane_w);
        return poly_bwd<NEED_DX>(sv, d, c, a);
      } else {
        static_assert(OP == RISP_OP_GTM && IN01, "packed path: unsupported effective op");
        GtmSaved sv;
        const P2 y = gtm_fwd(x, c, sv, slow);
        const P2 d = Run<SIG, K + 1, MODE>::go(y, tgt, cp, acc, loss, yout, slow, lane_w);
        return gtm_bwd<NEED_DX>(sv, d, c, a, slow);
      }
    }
  }
};

template <unsigned SIG, int K>
struct Fwd {
  static constexpr EffChain E = Eff<SIG>::e;
  static __device__ __forceinline__ P2 go(const P2& x, const float* __restrict__ cp, bool slow) {
    if constexpr (K == E.n) {
      return x;
    } else {
      constexpr int OP = E.op[K];
      constexpr bool IN01 = (K > 0) && fop_out01(E.op[K > 0 ? K - 1 : 0]);
      const float* c = cp + E.coff[K];
      P2 y;
      if constexpr (OP == RISP_OP_GAMMA) { GammaSaved sv; y = gamma_fwd<IN01>(x, c[0], sv); }
      else if constexpr (OP == RISP_OP_GAIN) { GainSaved sv; y = gain_fwd(x, c, sv); }
      else if constexpr (OP == RISP_OP_POLY10 || OP == FOP_POLYG) { PolySaved sv; y = poly_fwd(x, c, sv); }
      else { GtmSaved sv; y = gtm_fwd(x, c, sv, slow); }
      return Fwd<SIG, K + 1>::go(y, cp, slow);
    }
  }
};

template <unsigned SIG>
__device__ __forceinline__ bool chain_slow(const float* __restrict__ cp) {
  constexpr EffChain E = Eff<SIG>::e;
  bool slow = false;
#pragma unroll
  for (int k = 0; k < E.n; ++k)
    if (E.op[k] == RISP_OP_GTM) slow = slow || (cp[E.coff[k] + 5] != 0.f);
  return slow;
}

#ifndef RISP_FUSED_STEP_MINB
#define RISP_FUSED_STEP_MINB 2
#endif
#ifndef RISP_FUSED_FWD_MINB
#define RISP_FUSED_FWD_MINB 4
#endif

template <int DM, int MODE, unsigned SIG>
__global__ void __launch_bounds__(kWarps * 32, (MODE == MODE_FWD) ? RISP_FUSED_FWD_MINB : RISP_FUSED_STEP_MINB)
fused_kernel(FusedArgs a, ChainDesc d) {
  constexpr EffChain E = Eff<SIG>::e;
  constexpr int HL = (DM == RISP_DM_MALVAR) ? 2 : 1;
  constexpr int WR = 2 * HL + 1, WC = 4 + 2 * HL;
  constexpr int NACC = (MODE == MODE_FWD) ? 1 : E.nacc;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n = blockIdx.y, j0 = blockIdx.x;       // frame, CTA within the frame (both uniform: constants go to URs)
  const int H = a.H, W = a.W;
  const long long plane = (long long)H * W;
  const float* __restrict__ img = a.raw + (long long)n * plane;
  const float* __restrict__ gtb = (MODE != MODE_FWD) ? a.gt + (long long)n * 3 * plane : nullptr;
  float* __restrict__ yb = a.y ? a.y + (long long)n * 3 * plane : nullptr;
  // derived constants of this frame's parameter row (written by fused_prep_kernel earlier on the stream): loaded once,
  // they stay in registers for the whole kernel and enter the packed instructions as broadcast scalars (R.F32)
  float cst[E.ncst];
  {
    const float* __restrict__ crow = &g_cpar[a.slot][a.pstride ? n : 0][0];
#pragma unroll
    for (int i = 0; i < E.ncst; ++i) cst[i] = crow[i];
  }
  const float* cp = cst;
  const bool slow = chain_slow<SIG>(cp);

  float2 acc[NACC];
  float2 loss = zero2();
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = zero2();

  const int items = a.chunks * a.strip_blocks;
  for (int item = j0; item < items; item += a.cpf) {
    const int chunk = item / a.strip_blocks, sb = item - chunk * a.strip_blocks;
    const int strip = sb * kWarps + wid;
    const int c0 = strip * kStrip + lane * 4;
    const LaneGeom lg = lane_geom<HL>(c0, W, lane);
    const bool active = lg.active;
    const float lane_w = active ? 1.f : 0.f;
    const int ra = chunk * a.rows_per_chunk;
    const int rb = min(H, ra + a.rows_per_chunk);
    const float* __restrict__ rawl = img + c0;                          // the lane's first column, row 0
    const float* __restrict__ gtl = (MODE != MODE_FWD) ? gtb + c0 : nullptr;
    float* __restrict__ yl = yb ? yb + c0 : nullptr;

    float w[WR + 1][WC];
    RawRow<HL> q;
#pragma unroll
    for (int j = 0; j < WR - 1; ++j) {
      row_issue<HL>(q, rawl + (size_t)reflect101(ra - HL + j, H) * W, lg);
      row_finish<HL>(w[j], q, W, c0, lg);
    }
    row_issue<HL>(q, rawl + (size_t)reflect101(ra + HL, H) * W, lg);
    // GT rows ping-pong between two register sets (even rows in g0, odd rows in g1): the next row's loads never overwrite
    // the row being consumed, so no copies
    float4 g0[3], g1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { g0[c] = make_float4(0.f, 0.f, 0.f, 0.f); g1[c] = g0[c]; }
    auto gt_issue = [&](float4 (&g)[3], int row) {
      if (MODE != MODE_FWD && active) {
        const float* pr = gtl + (size_t)row * W;
        g[0] = ld_stream4(pr); g[1] = ld_stream4(pr + plane); g[2] = ld_stream4(pr + 2 * plane);
      }
    };
    gt_issue(g0, ra);

    // one row: finish the in-flight raw row into the window, prefetch the next raw / GT row, demosaic, chain, store
    auto do_row = [&](auto ro_tag, auto odd_tag, int r, bool more, const float4 (&tg)[3], float4 (&gnext)[3]) {
      constexpr int RO = decltype(ro_tag)::value;
      constexpr bool ODD = decltype(odd_tag)::value;
      row_finish<HL>(w[RO + WR - 1], q, W, c0, lg);
      if (more) {
        row_issue<HL>(q, rawl + (size_t)reflect101(r + 1 + HL, H) * W, lg);
        gt_issue(gnext, r + 1);
      }
      P2 lo, hi;
      demosaic4<DM, HL, RO, ODD, WR + 1>(w, a.clip_hi, lo, hi);
      P2 ylo, yhi;
      if constexpr (MODE == MODE_FWD) {
        ylo = Fwd<SIG, 0>::go(lo, cp, slow);
        yhi = Fwd<SIG, 0>::go(hi, cp, slow);
      } else {
        P2 tlo, thi;
        tlo.b = make_float2(tg[0].x, tg[0].y); tlo.g = make_float2(tg[1].x, tg[1].y); tlo.r = make_float2(tg[2].x, tg[2].y);
        thi.b = make_float2(tg[0].z, tg[0].w); thi.g = make_float2(tg[1].z, tg[1].w); thi.r = make_float2(tg[2].z, tg[2].w);
        Run<SIG, 0, MODE>::go(lo, tlo, cp, acc, loss, ylo, slow, lane_w);
        Run<SIG, 0, MODE>::go(hi, thi, cp, acc, loss, yhi, slow, lane_w);
      }
      if ((MODE == MODE_FWD || yb) && active) {
        float* po = yl + (size_t)r * W;
        st_stream4(po, make_float4(ylo.b.x, ylo.b.y, yhi.b.x, yhi.b.y));
        st_stream4(po + plane, make_float4(ylo.g.x, ylo.g.y, yhi.g.x, yhi.g.y));
        st_stream4(po + 2 * plane, make_float4(ylo.r.x, ylo.r.y, yhi.r.x, yhi.r.y));
      }
    };

    for (int r = ra; r < rb; r += 2) {      // ra, rb even
      do_row(std::integral_constant<int, 0>{}, std::false_type{}, r, true, g0, g1);
      do_row(std::integral_constant<int, 1>{}, std::true_type{}, r + 1, r + 2 < rb, g1, g0);
#pragma unroll
      for (int j = 0; j < WR - 1; ++j)
#pragma unroll
        for (int i = 0; i < WC; ++i) w[j][i] = w[j + 2][i];
    }
  }

  if constexpr (MODE != MODE_FWD) {
    // ---- epilogue: one partial row per warp, slot layout of risp_common.cuh ------------------------------------------
    float tot[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) tot[i] = warp_sum(acc[i].x + acc[i].y);
    const float lsum = warp_sum(loss.x + loss.y);
    if (lane == 0) {
      float* __restrict__ out = a.partial + (((long long)n * a.cpf + j0) * kWarps + wid) * RISP_NSLOT;
#pragma unroll
      for (int i = 0; i < RISP_NSLOT; ++i) out[i] = 0.f;
      out[RISP_SLOT_LOSS] = lsum;
      const float* __restrict__ prow = a.params + (long long)n * a.pstride;
      float scale = 1.f;        // product of the uniform factors deferred by the stages downstream
#pragma unroll
      for (int k = E.n - 1; k >= 0; --k) {
        const float* t = tot + E.aoff[k];
        const int src = E.src[k];
        if (E.op[k] == RISP_OP_GAMMA) {
          out[src * RISP_SMALL_ACC] = t[0] * RISP_LN2 * scale;
          scale *= cp[E.coff[k]];
        } else if (E.op[k] == RISP_OP_GAIN) {
#pragma unroll
          for (int i = 0; i < 3; ++i) out[src * RISP_SMALL_ACC + i] = t[i] * scale;
        } else if (E.op[k] == RISP_OP_GTM) {
          out[src * RISP_SMALL_ACC + 0] = 4.f * ((t[0] - 2.f * t[1]) + t[2]) * scale;
          out[src * RISP_SMALL_ACC + 1] = 4.f * ((t[1] - 2.f * t[2]) + t[3]) * scale;
          out[src * RISP_SMALL_ACC + 2] = 4.f * (t[2] - 2.f * t[3]) * scale;
        } else if (E.op[k] == RISP_OP_POLY10) {
#pragma unroll
          for (int i = 0; i < 30; ++i) out[RISP_SLOT_BIG + i] = t[i] * scale;
        } else if (E.op[k] == FOP_POLYG) {
          // S[c][i] = sum e_c phi_i(x0)  ->  d/dq[c][i] = gmon_i S ,  d/dg = sum q (d gmon / d g) S
          const float* g = prow + d.off[src - 1];
          const float* qq = prow + d.off[src];
          const float gb = g[0], gg = g[1], gr = g[2];
          const float gmon[10] = {gb * gb, gg * gg, gr * gr, gb * gg, gb * gr, gg * gr, gb, gg, gr, 1.f};
          const float dgb[10] = {2.f * gb, 0.f, 0.f, gg, gr, 0.f, 1.f, 0.f, 0.f, 0.f};
          const float dgg[10] = {0.f, 2.f * gg, 0.f, gb, 0.f, gr, 0.f, 1.f, 0.f, 0.f};
          const float dgr[10] = {0.f, 0.f, 2.f * gr, 0.f, gb, gg, 0.f, 0.f, 1.f, 0.f};
          float sb = 0.f, sg = 0.f, sr = 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int i = 0; i < 10; ++i) {
              const float s = t[c * 10 + i] * scale;
              out[RISP_SLOT_BIG + c * 10 + i] = gmon[i] * s;
              const float qs = qq[c * 10 + i] * s;
              sb = fmaf(qs, dgb[i], sb); sg = fmaf(qs, dgg[i], sg); sr = fmaf(qs, dgr[i], sr);
            }
          out[(src - 1) * RISP_SMALL_ACC + 0] = sb; out[(src - 1) * RISP_SMALL_ACC + 1] = sg; out[(src - 1) * RISP_SMALL_ACC + 2] = sr;
        }
      }
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
struct Geometry { int rows_per_chunk, chunks, strip_blocks, cpf, grid; };

static Geometry geometry(int N, int H, int W, int ctas_per_sm) {
  Geometry g;
  g.strip_blocks = (int)cdiv(W, kStrip * kWarps);
  const int slots = sm_count() * ctas_per_sm;
  g.cpf = slots / N < 1 ? 1 : slots / N;
  // rows per chunk: minimise rounds * (rows + per-item overhead); an item costs ~3 extra rows (window fill, exposed latency)
  long long best = -1;
  int best_rows = 2;
  for (int rows = 2; rows <= 256; rows += 2) {
    if (rows > H && rows > 2) break;
    const int chunks = (int)cdiv(H, rows);
    const long long items = (long long)chunks * g.strip_blocks;
    const long long rounds = cdiv(items, g.cpf);
    const long long cost = rounds * (rows + 3);
    if (best < 0 || cost < best || (cost == best && rows > best_rows)) { best = cost; best_rows = rows; }
  }
  g.rows_per_chunk = best_rows;
  g.chunks = (int)cdiv(H, best_rows);
  const long long items = (long long)g.chunks * g.strip_blocks;
  if (items < g.cpf) g.cpf = (int)items;
  g.grid = N * g.cpf;
  return g;
}

static std::mutex g_slot_mu;
static cudaStream_t g_slot_stream[kCSlots];
static int g_slot_used = 0;

// one constant slot per stream: prep kernel and consumer are ordered by the stream, so a slot is never rewritten while
// an earlier launch on that stream still reads it.  More than kCSlots distinct streams share the last slot (documented).
static int slot_of(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_slot_mu);
  for (int i = 0; i < g_slot_used; ++i)
    if (g_slot_stream[i] == st) return i;
  if (g_slot_used < kCSlots) { g_slot_stream[g_slot_used] = st; return g_slot_used++; }
  return kCSlots - 1;
}

static float* cpar_device_ptr() {
  static float* p = nullptr;
  if (!p) {
    void* q = nullptr;
    if (cudaGetSymbolAddress(&q, g_cpar) == cudaSuccess) p = static_cast<float*>(q);
  }
  return p;
}

template <int DM, int MODE, unsigned SIG>
static int resident_ctas() {
  static int n = 0;
  if (n == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fused_kernel<DM, MODE, SIG>, kWarps * 32, 0) != cudaSuccess || n < 1) n = 1;
  }
  return n;
}

template <int DM, int MODE, unsigned SIG>
static int launch_one(FusedArgs a, const ChainDesc& d, int N, cudaStream_t st, Geometry* gout) {
  const Geometry g = geometry(N, a.H, a.W, resident_ctas<DM, MODE, SIG>());
  if (gout) { *gout = g; return RISP_OK; }
  a.rows_per_chunk = g.rows_per_chunk; a.chunks = g.chunks; a.strip_blocks = g.strip_blocks; a.cpf = g.cpf;
  fused_kernel<DM, MODE, SIG><<<dim3(g.cpf, N), kWarps * 32, 0, st>>>(a, d);
  return check_launch("fused_kernel");
}

template <int MODE>
static int dispatch(const FusedArgs& a, const ChainDesc& d, unsigned sig, int N, int dm_kind, cudaStream_t st, Geometry* gout) {
#define RISP_F_SIG(DMK, SG) if (sig == (SG)) return launch_one<DMK, MODE, (SG)>(a, d, N, st, gout);
#define RISP_F_DM(DMK) RISP_F_SIG(DMK, RISP_SIG_A) RISP_F_SIG(DMK, RISP_SIG_B) RISP_F_SIG(DMK, RISP_SIG_C) RISP_F_SIG(DMK, RISP_SIG_D)
  switch (dm_kind) {
    case RISP_DM_NEAREST: RISP_F_DM(RISP_DM_NEAREST) break;
    case RISP_DM_BILINEAR: RISP_F_DM(RISP_DM_BILINEAR) break;
    case RISP_DM_MALVAR: RISP_F_DM(RISP_DM_MALVAR) break;
    default: break;
  }
#undef RISP_F_DM
#undef RISP_F_SIG
  return 1;   // not handled
}

}  // namespace fused

// true if the packed kernels take this chain (a pre-instantiated signature, parameter rows fit a constant slot)
bool fused_handles(const ChainDesc& d, int N, int param_stride, int H, int W) {
  if ((long long)H * W * 3 >= (1ll << 31)) return false;     // 32-bit element offsets inside a frame
  const unsigned sig = chain_signature(d);
  if (sig != RISP_SIG_A && sig != RISP_SIG_B && sig != RISP_SIG_C && sig != RISP_SIG_D) return false;
  if (param_stride != 0 && N > fused::kCRows) return false;
  return fused::eff_from_ops(d.op, d.iarg, d.S).ok;
}

// partial rows the step / backward kernels write: [N][rows_per_frame][RISP_NSLOT]
int fused_partial_rows_per_frame(int N, int H, int W) {
  // the largest grid any instantiation may use: 8 resident CTAs per SM is an upper bound for 128-thread CTAs here
  int slots = sm_count() * 8;
  int cpf = slots / N < 1 ? 1 : slots / N;
  (void)H; (void)W;
  return cpf * fused::kWarps;
}

// mode: 0 forward, 1 MSE step, 2 backward with upstream dy.  Returns RISP_OK, an error, or 1 when the chain is not handled.
// *rows_per_frame receives the number of partial rows per frame actually written (modes 1, 2).
int fused_launch(int mode, const float* raw, const float* gt, float* y, float* partial, const float* params, int pstride,
                 int N, int H, int W, int dm_kind, float clip_hi, const ChainDesc& d, cudaStream_t st, int* rows_per_frame) {
  using namespace fused;
  if (!fused_handles(d, N, pstride, H, W)) return 1;
  float* cbase = cpar_device_ptr();
  RISP_REQUIRE(cbase, RISP_E_CUDA, "fused pipeline: cannot resolve the constant parameter block");
  const int slot = slot_of(st);
  const int rows = pstride ? N : 1;
  fused_prep_kernel<<<rows, 32, 0, st>>>(params, pstride, d, cbase + (size_t)slot * kCRows * kCRowFloats);
  int rc = check_launch("fused_prep_kernel");
  if (rc != RISP_OK) return rc;
  FusedArgs a{raw, gt, y, partial, params, pstride, H, W, 0, 0, 0, 0, clip_hi, slot};
  const unsigned sig = chain_signature(d);
  Geometry g;
  if (mode == 0) return dispatch<MODE_FWD>(a, d, sig, N, dm_kind, st, nullptr);
  rc = (mode == 1) ? dispatch<MODE_STEP>(a, d, sig, N, dm_kind, st, &g) : dispatch<MODE_BWD>(a, d, sig, N, dm_kind, st, &g);
  if (rc != RISP_OK) return rc;
  if (rows_per_frame) *rows_per_frame = g.cpf * kWarps;
  return (mode == 1) ? dispatch<MODE_STEP>(a, d, sig, N, dm_kind, st, nullptr) : dispatch<MODE_BWD>(a, d, sig, N, dm_kind, st, nullptr);
}

}  // namespace risp
