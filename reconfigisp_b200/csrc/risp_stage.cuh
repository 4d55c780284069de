// Per-pixel stage arithmetic (forward and backward) shared by every kernel that evaluates
// classical sRGB candidates in registers: the stage chain, the fused raw->BGR pipeline and the
// DARTS mixed-op.  One pixel = (b, g, r) floats; parameters are kernel-level values read from a
// block-uniform row `p` (see include/reconfigisp_b200.h, enum risp_op).
//
// Reference semantics: tools_origin.py:48-73 (gamma), :200-225 (wb manual), :313-359 (WbQuadratic),
// :409-440 (GtmManual), :513-630 (tone operators); definitions of the un-shipped kernels are the
// ones written down in oracle/SPEC.md.
#pragma once
#include "risp_common.cuh"

namespace risp {

#define RISP_GAMMA_EPS 1e-8f
#define RISP_LN2 0.6931471805599453f
#define RISP_SMALL_ACC 4   // per-stage small accumulator slots (gamma 1, gain 3, gtm <= 4 knots)
#define RISP_BIG_ACC 30    // one "big" op per chain (POLY10: 30, CCM: 9)

__device__ __forceinline__ float hable(float v) {
  const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
  return __fdividef(v * (A * v + C * B) + D * E, v * (A * v + B) + D * F) - E / F;
}

__device__ __forceinline__ float gamma_px(float x, float gm) {
  float xc = fminf(fmaxf(x, RISP_GAMMA_EPS), 1.f);
  return exp2f(gm * __log2f(xc));
}

__device__ __forceinline__ float gtm_px(float x, const float* __restrict__ p, int n) {
  // knots y_0 = 0, y_k = p[k-1], y_n = 1 ; x-bounds k/n ; half-open tests ; outside [0,1) keeps x
  float out = x;
  const float fn = (float)n;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    if (k < n) {
      float sx = (float)k / fn, ex = (float)(k + 1) / fn;
      float sy = (k > 0) ? p[k - 1] : 0.f;
      float ey = (k < n - 1) ? p[k] : 1.f;
      float slope = __fdiv_rn(ey - sy, ex - sx);
      if (x >= sx && x < ex) out = __fmaf_rn(x - sx, slope, sy);
    }
  }
  return sat01(out);
}

// ------------------------------------------------------------------------------------------------
// forward of one stage on one pixel
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_fwd(int op, int iarg, const float* __restrict__ p, float& b, float& g,
                                          float& r) {
  switch (op) {
    case RISP_OP_GAMMA: {
      float gm = p[0];
      b = gamma_px(b, gm); g = gamma_px(g, gm); r = gamma_px(r, gm);
    } break;
    case RISP_OP_GAIN:
      b *= p[0]; g *= p[1]; r *= p[2];
      break;
    case RISP_OP_GAIN_CLIP:
      b = sat01(b * p[0]); g = sat01(g * p[1]); r = sat01(r * p[2]);
      break;
    case RISP_OP_POLY10: {
      float bb = b * b, gg = g * g, rr = r * r, bg = b * g, br = b * r, gr = g * r;
      float o[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* q = p + c * 10;
        float u = q[9];
        u = fmaf(q[8], r, u); u = fmaf(q[7], g, u); u = fmaf(q[6], b, u);
        u = fmaf(q[5], gr, u); u = fmaf(q[4], br, u); u = fmaf(q[3], bg, u);
        u = fmaf(q[2], rr, u); u = fmaf(q[1], gg, u); u = fmaf(q[0], bb, u);
        o[c] = sat01(u);
      }
      b = o[0]; g = o[1]; r = o[2];
    } break;
    case RISP_OP_GTM:
      b = gtm_px(b, p, iarg); g = gtm_px(g, p, iarg); r = gtm_px(r, p, iarg);
      break;
    case RISP_OP_CCM: {
      float o0 = sat01(p[0] * b + p[1] * g + p[2] * r);
      float o1 = sat01(p[3] * b + p[4] * g + p[5] * r);
      float o2 = sat01(p[6] * b + p[7] * g + p[8] * r);
      b = o0; g = o1; r = o2;
    } break;
    case RISP_OP_REINHARD: {
      float s = p[0], iw2 = p[1];
      float v;
      v = s * b; b = sat01(__fdividef(v * fmaf(v, iw2, 1.f), 1.f + v));
      v = s * g; g = sat01(__fdividef(v * fmaf(v, iw2, 1.f), 1.f + v));
      v = s * r; r = sat01(__fdividef(v * fmaf(v, iw2, 1.f), 1.f + v));
    } break;
    case RISP_OP_CRYSIS: {
      float il = p[0] * 1.4426950408889634f;
      b = sat01(1.f - exp2f(-b * il)); g = sat01(1.f - exp2f(-g * il)); r = sat01(1.f - exp2f(-r * il));
    } break;
    case RISP_OP_FILMIC: {
      float e = p[0], ifw = p[1];
      b = sat01(hable(e * b) * ifw); g = sat01(hable(e * g) * ifw); r = sat01(hable(e * r) * ifw);
    } break;
    default:
      break;  // RISP_OP_SKIP
  }
}

// ------------------------------------------------------------------------------------------------
// backward of one stage on one pixel.  (b,g,r) = the stage's INPUT; (db,dg,dr) in: grad wrt output,
// out: grad wrt input.  accS: this stage's small accumulators, accB: the chain's big accumulator.
// `wgt` scales parameter gradients only (mixed-op branch weight); pass 1 for chains.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gamma_bwd_px(float x, float gm, float& d, float& dgm) {
  float xc = fminf(fmaxf(x, RISP_GAMMA_EPS), 1.f);
  float l2 = __log2f(xc);
  float y = exp2f(gm * l2);
  dgm = fmaf(d * y, l2 * RISP_LN2, dgm);
  bool inside = (x >= RISP_GAMMA_EPS) && (x <= 1.f);
  d = inside ? d * gm * __fdividef(y, xc) : 0.f;
}

__device__ __forceinline__ void gtm_bwd_px(float x, const float* __restrict__ p, int n, float& d,
                                           float (&accS)[RISP_SMALL_ACC]) {
  const float fn = (float)n;
  float dxv = d;      // pass-through branch: out = x
  float outv = x;
  float w_lo = 0.f, w_hi = 0.f;
  int kk = -1;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    if (k < n) {
      float sx = (float)k / fn, ex = (float)(k + 1) / fn;
      float sy = (k > 0) ? p[k - 1] : 0.f;
      float ey = (k < n - 1) ? p[k] : 1.f;
      float slope = __fdiv_rn(ey - sy, ex - sx);
      if (x >= sx && x < ex) {
        outv = __fmaf_rn(x - sx, slope, sy);
        dxv = d * slope;
        float t = __fdiv_rn(x - sx, ex - sx);
        w_lo = d * (1.f - t); w_hi = d * t; kk = k;
      }
    }
  }
  float m = in01(outv);
  d = dxv * m;
#pragma unroll
  for (int j = 0; j < RISP_SMALL_ACC; ++j) {
    // knot j is y_{j+1}: the low end of segment j+1 and the high end of segment j
    float add = (kk == j + 1 ? w_lo : 0.f) + (kk == j ? w_hi : 0.f);
    accS[j] = fmaf(add, m, accS[j]);
  }
}

template <bool BIG>
__device__ __forceinline__ void stage_bwd(int op, int iarg, const float* __restrict__ p, float b, float g,
                                          float r, float& db, float& dg, float& dr,
                                          float (&accS)[RISP_SMALL_ACC], float (&accB)[RISP_BIG_ACC]) {
  switch (op) {
    case RISP_OP_GAMMA: {
      float gm = p[0];
      gamma_bwd_px(b, gm, db, accS[0]); gamma_bwd_px(g, gm, dg, accS[0]); gamma_bwd_px(r, gm, dr, accS[0]);
    } break;
    case RISP_OP_GAIN:
      accS[0] = fmaf(db, b, accS[0]); accS[1] = fmaf(dg, g, accS[1]); accS[2] = fmaf(dr, r, accS[2]);
      db *= p[0]; dg *= p[1]; dr *= p[2];
      break;
    case RISP_OP_GAIN_CLIP: {
      float e0 = db * in01(b * p[0]), e1 = dg * in01(g * p[1]), e2 = dr * in01(r * p[2]);
      accS[0] = fmaf(e0, b, accS[0]); accS[1] = fmaf(e1, g, accS[1]); accS[2] = fmaf(e2, r, accS[2]);
      db = e0 * p[0]; dg = e1 * p[1]; dr = e2 * p[2];
    } break;
    case RISP_OP_POLY10:
      if (BIG) {
        float phi[9] = {b * b, g * g, r * r, b * g, b * r, g * r, b, g, r};
        float dphi[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float din[3] = {db, dg, dr};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float* q = p + c * 10;
          float u = q[9];
#pragma unroll
          for (int k = 8; k >= 0; --k) u = fmaf(q[k], phi[k], u);
          float e = din[c] * in01(u);
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            dphi[k] = fmaf(e, q[k], dphi[k]);
            accB[c * 10 + k] = fmaf(e, phi[k], accB[c * 10 + k]);
          }
          accB[c * 10 + 9] += e;
        }
        db = fmaf(2.f * b, dphi[0], fmaf(g, dphi[3], fmaf(r, dphi[4], dphi[6])));
        dg = fmaf(2.f * g, dphi[1], fmaf(b, dphi[3], fmaf(r, dphi[5], dphi[7])));
        dr = fmaf(2.f * r, dphi[2], fmaf(b, dphi[4], fmaf(g, dphi[5], dphi[8])));
      }
      break;
    case RISP_OP_GTM:
      gtm_bwd_px(b, p, iarg, db, accS); gtm_bwd_px(g, p, iarg, dg, accS); gtm_bwd_px(r, p, iarg, dr, accS);
      break;
    case RISP_OP_CCM:
      if (BIG) {
        float xin[3] = {b, g, r};
        float din[3] = {db, dg, dr};
        float dx[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float u = p[c * 3] * b + p[c * 3 + 1] * g + p[c * 3 + 2] * r;
          float e = din[c] * in01(u);
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            dx[d] = fmaf(e, p[c * 3 + d], dx[d]);
            accB[c * 3 + d] = fmaf(e, xin[d], accB[c * 3 + d]);
          }
        }
        db = dx[0]; dg = dx[1]; dr = dx[2];
      }
      break;
    default:
      break;  // SKIP: identity.  Forward-only ops are rejected on the host.
  }
}

__host__ __device__ __forceinline__ bool op_is_big(int op) { return op == RISP_OP_POLY10 || op == RISP_OP_CCM; }

}  // namespace risp
