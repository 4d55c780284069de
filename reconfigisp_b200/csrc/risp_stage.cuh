// Per-pixel stage arithmetic (forward and backward) shared by every kernel that evaluates
// classical sRGB candidates in registers: the stage chain, the fused raw->BGR pipeline and the
// DARTS mixed-op.  A stage works on a GROUP of NPX pixels (NPX even) so that its parameters are
// fetched once per group and the polynomial / matrix stages can use Blackwell's packed fp32
// instructions (FFMA2 / FMUL2 / FADD2: two pixels per issue slot).
//
// Reference semantics: tools_origin.py:48-73 (gamma), :200-225 (wb manual), :313-359 (WbQuadratic),
// :409-440 (GtmManual), :513-630 (tone operators); definitions of the un-shipped kernels are the
// ones written down in oracle/SPEC.md.  Parameters are kernel-level values (enum risp_op).
#pragma once
#include "risp_common.cuh"

namespace risp {

#define RISP_GAMMA_EPS 1e-8f
#define RISP_LN2 0.6931471805599453f
#define RISP_SMALL_ACC 4   // per-stage small accumulator slots (gamma 1, gain 3, gtm <= 4 knots)
#define RISP_BIG_ACC 15    // float2 slots of the one "big" op per chain (POLY10: 30 floats, CCM: 9)

template <int NPX>
struct Px {
  float b[NPX], g[NPX], r[NPX];
};

__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float hable(float v) {
  const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
  return __fdividef(v * (A * v + C * B) + D * E, v * (A * v + B) + D * F) - E / F;
}

__device__ __forceinline__ float gamma_px(float x, float gm) {
  float xc = fminf(fmaxf(x, RISP_GAMMA_EPS), 1.f);
  return ex2_ftz(gm * lg2_ftz(xc));
}

// ---- piecewise-linear tone curve in hinge form ---------------------------------------------------------
// knots y_0 = 0, y_k = p[k-1], y_n = 1 on the uniform grid x_k = k/n.  On [0,1) the reference's segment
// formula (tools_origin.py:429-435) is the continuous PWL function
//     f(x) = s_0 x + sum_{k>=1} (s_k - s_{k-1}) max(x - x_k, 0),    s_k = (y_{k+1} - y_k) n,
// which needs no segment search.  Pixels outside [0,1) pass through and are clamped (:438), i.e. 0 / 1.
struct GtmCoef {
  float s0;
  float ds[4];   // slope increments at the interior knots x_1..x_{n-1} (unused entries are 0)
  float xk[4];
  float fn;
  bool knots_in_range;   // all y_k in [0,1]  =>  f(x) in [0,1] on [0,1) and the clamp mask is always 1
};

__device__ __forceinline__ GtmCoef gtm_prepare(const float* __restrict__ p, int n) {
  GtmCoef c;
  const float fn = (float)n;
  c.fn = fn;
  c.knots_in_range = true;
  float prev_y = 0.f, prev_s = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    float s = 0.f;
    if (k < n) {
      const float ey = (k < n - 1) ? p[k] : 1.f;
      if (k < n - 1) c.knots_in_range = c.knots_in_range && (ey >= 0.f) && (ey <= 1.f);
      s = (ey - prev_y) * fn;
      prev_y = ey;
    }
    if (k == 0) c.s0 = s;
    else { c.ds[k - 1] = (k < n) ? s - prev_s : 0.f; c.xk[k - 1] = (float)k / fn; }
    if (k < n) prev_s = s;
  }
  return c;
}

__device__ __forceinline__ float gtm_px(float x, const GtmCoef& c, int n) {
  const float xc = __saturatef(x);
  float f = c.s0 * xc;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < n - 1) f = fmaf(c.ds[k], fmaxf(xc - c.xk[k], 0.f), f);
  return __saturatef(f);
}

// backward: d (in: dL/dout, out: dL/dx) and the knot gradients via the hat basis
//   d out / d y_j = max(0, 1 - n |x - x_j|)  for x in [0,1)
template <bool CHECK_RANGE>
__device__ __forceinline__ void gtm_bwd_px(float x, const GtmCoef& c, int n, float& d,
                                           float (&accS)[RISP_SMALL_ACC]) {
  const bool inside = (x >= 0.f) && (x < 1.f);
  float slope = c.s0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < n - 1) slope += (x >= c.xk[k]) ? c.ds[k] : 0.f;
  float dm = inside ? d : 0.f;
  if (CHECK_RANGE) {                 // only when some knot lies outside [0,1] (never for sigmoid-generated knots)
    float f = c.s0 * x;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < n - 1) f = fmaf(c.ds[k], fmaxf(x - c.xk[k], 0.f), f);
    dm = (f >= 0.f && f <= 1.f) ? dm : 0.f;
  }
#pragma unroll
  for (int j = 0; j < RISP_SMALL_ACC; ++j) {
    if (j < n - 1) {
      const float hat = fmaxf(fmaf(-c.fn, fabsf(x - c.xk[j]), 1.f), 0.f);
      accS[j] = fmaf(dm, hat, accS[j]);
    }
  }
  // pass-through region: out = x, clamp mask inclusive -> only x == 1 keeps a gradient (slope 1)
  d = inside ? dm * slope : ((x == 1.f) ? d : 0.f);
}

// ---- packed helpers ------------------------------------------------------------------------------------
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 splat(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 sat01_2(float2 v) { return make_float2(sat01(v.x), sat01(v.y)); }
__device__ __forceinline__ float2 in01_2(float2 v) { return make_float2(in01(v.x), in01(v.y)); }

// monomial basis of WbQuadratic packed in pairs: [(b2,g2) (r2,bg) (br,gr) (b,g) (r,1)]  (3 FMUL2)
__device__ __forceinline__ void poly10_basis(float b, float g, float r, float2 (&phi)[5]) {
  const float2 bg = f2(b, g);
  phi[0] = __fmul2_rn(bg, bg);
  phi[1] = __fmul2_rn(f2(r, b), f2(r, g));
  phi[2] = __fmul2_rn(bg, splat(r));
  phi[3] = bg;
  phi[4] = f2(r, 1.f);
}

// ------------------------------------------------------------------------------------------------
// forward of one stage on a group of NPX pixels
// ------------------------------------------------------------------------------------------------
template <int NPX>
__device__ __forceinline__ void stage_fwd(int op, int iarg, const float* __restrict__ p, Px<NPX>& x) {
  switch (op) {
    case RISP_OP_GAMMA: {
      const float gm = p[0];
#pragma unroll
      for (int k = 0; k < NPX; ++k) { x.b[k] = gamma_px(x.b[k], gm); x.g[k] = gamma_px(x.g[k], gm); x.r[k] = gamma_px(x.r[k], gm); }
    } break;
    case RISP_OP_GAIN: {
      const float g0 = p[0], g1 = p[1], g2 = p[2];
#pragma unroll
      for (int k = 0; k < NPX; ++k) { x.b[k] *= g0; x.g[k] *= g1; x.r[k] *= g2; }
    } break;
    case RISP_OP_GAIN_CLIP: {
      const float g0 = p[0], g1 = p[1], g2 = p[2];
#pragma unroll
      for (int k = 0; k < NPX; ++k) { x.b[k] = sat01(x.b[k] * g0); x.g[k] = sat01(x.g[k] * g1); x.r[k] = sat01(x.r[k] * g2); }
    } break;
    case RISP_OP_POLY10: {
      // coefficients packed in pairs along k: q2[c][j] = (P[c][2j], P[c][2j+1]); one FFMA2 = two monomials
      float2 q2[15];
#pragma unroll
      for (int i = 0; i < 15; ++i) q2[i] = make_float2(p[2 * i], p[2 * i + 1]);
#pragma unroll
      for (int k = 0; k < NPX; ++k) {
        float2 phi[5];
        poly10_basis(x.b[k], x.g[k], x.r[k], phi);
        float o[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float2 u = __fmul2_rn(q2[c * 5 + 4], phi[4]);
#pragma unroll
          for (int i = 3; i >= 0; --i) u = __ffma2_rn(q2[c * 5 + i], phi[i], u);
          o[c] = sat01(u.x + u.y);
        }
        x.b[k] = o[0]; x.g[k] = o[1]; x.r[k] = o[2];
      }
    } break;
    case RISP_OP_GTM: {
      const GtmCoef c = gtm_prepare(p, iarg);
#pragma unroll
      for (int k = 0; k < NPX; ++k) { x.b[k] = gtm_px(x.b[k], c, iarg); x.g[k] = gtm_px(x.g[k], c, iarg); x.r[k] = gtm_px(x.r[k], c, iarg); }
    } break;
    case RISP_OP_CCM: {
      float q[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) q[i] = p[i];
#pragma unroll
      for (int k = 0; k < NPX; ++k) {
        const float b = x.b[k], g = x.g[k], r = x.r[k];
        x.b[k] = sat01(fmaf(q[0], b, fmaf(q[1], g, q[2] * r)));
        x.g[k] = sat01(fmaf(q[3], b, fmaf(q[4], g, q[5] * r)));
        x.r[k] = sat01(fmaf(q[6], b, fmaf(q[7], g, q[8] * r)));
      }
    } break;
    case RISP_OP_REINHARD: {
      const float s = p[0], iw2 = p[1];
      auto f = [&](float v0) { float v = s * v0; return sat01(__fdividef(v * fmaf(v, iw2, 1.f), 1.f + v)); };
#pragma unroll
      for (int k = 0; k < NPX; ++k) { x.b[k] = f(x.b[k]); x.g[k] = f(x.g[k]); x.r[k] = f(x.r[k]); }
    } break;
    case RISP_OP_CRYSIS: {
      const float il = p[0] * 1.4426950408889634f;
      auto f = [&](float v) { return sat01(1.f - ex2_ftz(-v * il)); };
#pragma unroll
      for (int k = 0; k < NPX; ++k) { x.b[k] = f(x.b[k]); x.g[k] = f(x.g[k]); x.r[k] = f(x.r[k]); }
    } break;
    case RISP_OP_FILMIC: {
      const float e = p[0], ifw = p[1];
      auto f = [&](float v) { return sat01(hable(e * v) * ifw); };
#pragma unroll
      for (int k = 0; k < NPX; ++k) { x.b[k] = f(x.b[k]); x.g[k] = f(x.g[k]); x.r[k] = f(x.r[k]); }
    } break;
    default:
      break;  // RISP_OP_SKIP
  }
}

// ------------------------------------------------------------------------------------------------
// backward of one stage on a group.  x = the stage's INPUT; d in: grad wrt output, out: grad wrt input.
// accS: this stage's small accumulators; accB: the chain's big accumulator, 30 floats viewed as 15
// float2 (POLY10 coefficient pairs (P[c][2j], P[c][2j+1]) -> one FFMA2 per pair).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gamma_bwd_px(float x, float y, float gm, float& d, float& dgm) {
  const float xc = fminf(fmaxf(x, RISP_GAMMA_EPS), 1.f);
  const float l2 = lg2_ftz(xc);
  dgm = fmaf(d * y, l2, dgm);                         // * ln2 applied once when the accumulator is flushed
  const bool inside = (x >= RISP_GAMMA_EPS) && (x <= 1.f);
  d = inside ? d * gm * y * rcp_ftz(xc) : 0.f;
}

// y = the stage's OUTPUT (already computed by the forward sweep; saves recomputing it here)
template <int NPX, bool BIG>
__device__ __forceinline__ void stage_bwd(int op, int iarg, const float* __restrict__ p, const Px<NPX>& x,
                                          const Px<NPX>& y, Px<NPX>& d, float (&accS)[RISP_SMALL_ACC],
                                          float2 (&accB)[RISP_BIG_ACC]) {
  switch (op) {
    case RISP_OP_GAMMA: {
      const float gm = p[0];
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < NPX; ++k) { gamma_bwd_px(x.b[k], y.b[k], gm, d.b[k], a); gamma_bwd_px(x.g[k], y.g[k], gm, d.g[k], a); gamma_bwd_px(x.r[k], y.r[k], gm, d.r[k], a); }
      accS[0] = fmaf(a, RISP_LN2, accS[0]);
    } break;
    case RISP_OP_GAIN: {
      const float g0 = p[0], g1 = p[1], g2 = p[2];
#pragma unroll
      for (int k = 0; k < NPX; ++k) {
        accS[0] = fmaf(d.b[k], x.b[k], accS[0]); accS[1] = fmaf(d.g[k], x.g[k], accS[1]); accS[2] = fmaf(d.r[k], x.r[k], accS[2]);
        d.b[k] *= g0; d.g[k] *= g1; d.r[k] *= g2;
      }
    } break;
    case RISP_OP_GAIN_CLIP: {
      const float g0 = p[0], g1 = p[1], g2 = p[2];
#pragma unroll
      for (int k = 0; k < NPX; ++k) {
        const float u0 = x.b[k] * g0, u1 = x.g[k] * g1, u2 = x.r[k] * g2;
        const float e0 = (u0 >= 0.f && u0 <= 1.f) ? d.b[k] : 0.f, e1 = (u1 >= 0.f && u1 <= 1.f) ? d.g[k] : 0.f,
                    e2 = (u2 >= 0.f && u2 <= 1.f) ? d.r[k] : 0.f;
        accS[0] = fmaf(e0, x.b[k], accS[0]); accS[1] = fmaf(e1, x.g[k], accS[1]); accS[2] = fmaf(e2, x.r[k], accS[2]);
        d.b[k] = e0 * g0; d.g[k] = e1 * g1; d.r[k] = e2 * g2;
      }
    } break;
    case RISP_OP_POLY10:
      if (BIG) {
        float2 q2[15];
#pragma unroll
        for (int i = 0; i < 15; ++i) q2[i] = make_float2(p[2 * i], p[2 * i + 1]);
#pragma unroll
        for (int k = 0; k < NPX; ++k) {
          const float b = x.b[k], g = x.g[k], r = x.r[k];
          float2 phi[5];
          poly10_basis(b, g, r, phi);
          const float din[3] = {d.b[k], d.g[k], d.r[k]};
          float2 dphi[5];
#pragma unroll
          for (int i = 0; i < 5; ++i) dphi[i] = splat(0.f);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float2 u = __fmul2_rn(q2[c * 5 + 4], phi[4]);
#pragma unroll
            for (int i = 3; i >= 0; --i) u = __ffma2_rn(q2[c * 5 + i], phi[i], u);
            const float us = u.x + u.y;
            const float2 e = splat((us >= 0.f && us <= 1.f) ? din[c] : 0.f);
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              dphi[i] = __ffma2_rn(e, q2[c * 5 + i], dphi[i]);
              accB[c * 5 + i] = __ffma2_rn(e, phi[i], accB[c * 5 + i]);
            }
          }
          // phi = [(b2,g2) (r2,bg) (br,gr) (b,g) (r,1)]
          d.b[k] = fmaf(2.f * b, dphi[0].x, fmaf(g, dphi[1].y, fmaf(r, dphi[2].x, dphi[3].x)));
          d.g[k] = fmaf(2.f * g, dphi[0].y, fmaf(b, dphi[1].y, fmaf(r, dphi[2].y, dphi[3].y)));
          d.r[k] = fmaf(2.f * r, dphi[1].x, fmaf(b, dphi[2].x, fmaf(g, dphi[2].y, dphi[4].x)));
        }
      }
      break;
    case RISP_OP_GTM: {
      const GtmCoef c = gtm_prepare(p, iarg);
      if (c.knots_in_range) {      // block-uniform
#pragma unroll
        for (int k = 0; k < NPX; ++k) {
          gtm_bwd_px<false>(x.b[k], c, iarg, d.b[k], accS); gtm_bwd_px<false>(x.g[k], c, iarg, d.g[k], accS);
          gtm_bwd_px<false>(x.r[k], c, iarg, d.r[k], accS);
        }
      } else {
#pragma unroll
        for (int k = 0; k < NPX; ++k) {
          gtm_bwd_px<true>(x.b[k], c, iarg, d.b[k], accS); gtm_bwd_px<true>(x.g[k], c, iarg, d.g[k], accS);
          gtm_bwd_px<true>(x.r[k], c, iarg, d.r[k], accS);
        }
      }
    } break;
    case RISP_OP_CCM:
      if (BIG) {
        float q[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) q[i] = p[i];
        float* acc = reinterpret_cast<float*>(accB);   // linear view, static indices only
#pragma unroll
        for (int k = 0; k < NPX; ++k) {
          const float xin[3] = {x.b[k], x.g[k], x.r[k]};
          const float din[3] = {d.b[k], d.g[k], d.r[k]};
          float dx[3] = {0.f, 0.f, 0.f};
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float u = fmaf(q[c * 3], xin[0], fmaf(q[c * 3 + 1], xin[1], q[c * 3 + 2] * xin[2]));
            const float e = din[c] * in01(u);
#pragma unroll
            for (int j2 = 0; j2 < 3; ++j2) {
              dx[j2] = fmaf(e, q[c * 3 + j2], dx[j2]);
              acc[c * 3 + j2] = fmaf(e, xin[j2], acc[c * 3 + j2]);
            }
          }
          d.b[k] = dx[0]; d.g[k] = dx[1]; d.r[k] = dx[2];
        }
      }
      break;
    default:
      break;  // SKIP: identity.  Forward-only ops are rejected on the host.
  }
}

__host__ __device__ __forceinline__ constexpr bool op_is_big(int op) { return op == RISP_OP_POLY10 || op == RISP_OP_CCM; }

// ---- compile-time chain signatures ---------------------------------------------------------------------
// A signature packs (op+1) in 4-bit nibbles, stage 0 in the low nibble; 0 = the generic interpreter that
// reads the ops from the ChainDesc at run time.  With a signature every `switch (op)` folds away, unused
// accumulators die and the register allocator sees the real live set; the host picks a pre-instantiated
// signature when the requested chain matches one (GTM stages must have n_seg == 4) and falls back to
// the interpreter otherwise.
constexpr unsigned sig1(int a) { return (unsigned)(a + 1); }
constexpr unsigned sig_cat(unsigned lo, int op, int pos) { return lo | ((unsigned)(op + 1) << (4 * pos)); }
constexpr unsigned make_sig(int a, int b = -1, int c = -1, int d = -1, int e = -1, int f = -1) {
  return (unsigned)(a + 1) | ((unsigned)(b + 1) << 4) | ((unsigned)(c + 1) << 8) | ((unsigned)(d + 1) << 12) |
         ((unsigned)(e + 1) << 16) | ((unsigned)(f + 1) << 20);
}

template <unsigned SIG>
struct Sig {
  static constexpr bool generic = (SIG == 0);
  static constexpr int count() {
    int n = 0;
    for (int s = 0; s < 8; ++s) n += ((SIG >> (4 * s)) & 15u) ? 1 : 0;
    return n;
  }
  static constexpr int S = generic ? RISP_MAX_STAGES : count();
  static constexpr int op_c(int s) { return (int)((SIG >> (4 * s)) & 15u) - 1; }
  static constexpr bool big_c() {
    for (int s = 0; s < 8; ++s)
      if (((SIG >> (4 * s)) & 15u) && op_is_big(op_c(s))) return true;
    return false;
  }
  __device__ __forceinline__ static bool live(const ChainDesc& d, int s) { return generic ? (s < d.S) : true; }
  __device__ __forceinline__ static int op(const ChainDesc& d, int s) { return generic ? d.op[s] : op_c(s); }
  __device__ __forceinline__ static int iarg(const ChainDesc& d, int s) { return generic ? d.iarg[s] : 4; }
};

// Pre-instantiated signatures.  Fused fixed pipelines (sRGB tails of the shipped architectures and the
// all-classical pipeline of SURVEY.md §8d):
//   11_13_01_14 (wbmanual, wbquadratic, gamma, gtmmanual)   01_13_11 (SID_isp.yml)   01_13 (S7ISP tail)   01_14 (yolo head)
#define RISP_SIG_A ::risp::make_sig(RISP_OP_GAIN, RISP_OP_POLY10, RISP_OP_GAMMA, RISP_OP_GTM)
#define RISP_SIG_B ::risp::make_sig(RISP_OP_GAMMA, RISP_OP_POLY10, RISP_OP_GAIN)
#define RISP_SIG_C ::risp::make_sig(RISP_OP_GAMMA, RISP_OP_POLY10)
#define RISP_SIG_D ::risp::make_sig(RISP_OP_GAMMA, RISP_OP_GTM)
#define RISP_SIG_SKIP ::risp::make_sig(RISP_OP_SKIP)     // demosaic only (fused kernels)
#define RISP_FOR_EACH_CHAIN_SIG(X) X(RISP_SIG_A) X(RISP_SIG_B) X(RISP_SIG_C) X(RISP_SIG_D)
#define RISP_FOR_EACH_SINGLE_SIG(X)                                                                          \
  X(::risp::make_sig(RISP_OP_GAMMA)) X(::risp::make_sig(RISP_OP_GAIN)) X(::risp::make_sig(RISP_OP_GAIN_CLIP)) \
  X(::risp::make_sig(RISP_OP_POLY10)) X(::risp::make_sig(RISP_OP_GTM)) X(::risp::make_sig(RISP_OP_CCM))

// host: signature of a chain description (0 if it contains something a signature cannot express)
inline unsigned chain_signature(const ChainDesc& d) {
  unsigned sig = 0;
  if (d.S == 0 || d.S > 6) return 0;
  for (int s = 0; s < d.S; ++s) {
    if (d.op[s] == RISP_OP_GTM && d.iarg[s] != 4) return 0;
    sig |= (unsigned)(d.op[s] + 1) << (4 * s);
  }
  return sig;
}

}  // namespace risp
