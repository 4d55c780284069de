// Windowed stages on BGR images with shared-memory halo tiles:
//   bilateral / median     spatialnoisereduction.run  (tools_origin.py:696-710, 742-751; OpenCV semantics)
//   guided filter, sharpen north_star extensions (oracle/SPEC.md)
// One CTA owns a 32x16 output tile of one image and stages the tile plus its halo for all three
// planes in shared memory (coalesced row segments, border rule applied while loading), so every input
// pixel is fetched from L2/HBM ~once and the window loops run out of shared memory.
#include "risp_common.cuh"
#include "risp_march.cuh"

namespace risp {

constexpr int TW = 32, TH = 16, kT = TW * TH;
constexpr int kMaxR = 8;

enum { BORDER_REFLECT101 = 0, BORDER_REPLICATE = 1 };

__device__ __forceinline__ int border_idx(int i, int n, int mode) {
  if (mode == BORDER_REPLICATE) return i < 0 ? 0 : (i >= n ? n - 1 : i);
  // reflect-101, applied repeatedly for tiny images
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

// loads planes [0,C) of image n: tile origin (y0-R, x0-R), (TH+2R) x (TW+2R) floats per plane.  A warp takes whole tile
// rows (lanes = consecutive columns: coalesced row segments, border rule per element, no division in the loop).
template <int C>
__device__ __forceinline__ void load_tile(float* sh, const float* __restrict__ img, int H, int W, int y0, int x0, int R,
                                          int mode, float scale) {
  const int tw = TW + 2 * R, th = TH + 2 * R;
  const long long plane = (long long)H * W;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int ly = wid; ly < th; ly += kT / 32) {
    const long long ro = (long long)border_idx(y0 - R + ly, H, mode) * W;
    for (int lx = lane; lx < tw; lx += 32) {
      const int gx = border_idx(x0 - R + lx, W, mode);
#pragma unroll
      for (int c = 0; c < C; ++c) sh[c * tw * th + ly * tw + lx] = __ldg(img + c * plane + ro + gx) * scale;
    }
  }
}

// ---- bilateral ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT)
bilateral_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, const int* __restrict__ window,
                 const float* __restrict__ sigma_color, const float* __restrict__ sigma_space, int Rmax) {
  extern __shared__ float sh[];
  const int n = blockIdx.z;
  int R = window[n] / 2;
  R = R < 0 ? 0 : (R > Rmax ? Rmax : R);
  const float sc = sigma_color[n], ss = sigma_space[n];
  const float kc = -0.5f / (sc * sc), ks = -0.5f / (ss * ss);
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const long long plane = (long long)H * W;
  const float* img = x + (long long)n * 3 * plane;
  load_tile<3>(sh, img, H, W, y0, x0, R, BORDER_REFLECT101, 1.f);
  __syncthreads();
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const int gx = x0 + tx, gy = y0 + ty;
  if (gx >= W || gy >= H) return;
  const int tw = TW + 2 * R, th = TH + 2 * R, ps = tw * th;
  const float* c0 = sh + (ty + R) * tw + tx + R;
  const float cb = c0[0], cg = c0[ps], cr = c0[2 * ps];
  float nb = 0.f, ng = 0.f, nr = 0.f, den = 0.f;
  const int R2 = R * R;
  for (int dy = -R; dy <= R; ++dy) {
    for (int dx = -R; dx <= R; ++dx) {
      const int r2 = dy * dy + dx * dx;
      if (r2 > R2) continue;   // circular support (cv2.bilateralFilter)
      const float* q = c0 + dy * tw + dx;
      const float vb = q[0], vg = q[ps], vr = q[2 * ps];
      const float dist = fabsf(vb - cb) + fabsf(vg - cg) + fabsf(vr - cr);
      const float wgt = __expf(dist * dist * kc + (float)r2 * ks);
      nb = fmaf(wgt, vb, nb); ng = fmaf(wgt, vg, ng); nr = fmaf(wgt, vr, nr); den += wgt;
    }
  }
  float* o = y + (long long)n * 3 * plane + (long long)gy * W + gx;
  o[0] = nb / den; o[plane] = ng / den; o[2 * plane] = nr / den;
}

// ---- non-local means (oracle/SPEC.md 'fastnlm') --------------------------------------------------------
// y_c(p) = sum_q w(p,q) x_c(q) / sum_q w(p,q),  q over the s x s search window,
// w = exp(-d2/h^2),  d2 = mean over the b x b patch and the 3 channels of (x(p+o) - x(q+o))^2.
// Halo = s/2 + b/2 (reflect-101 applied while staging), everything runs out of the shared tile.
__global__ void __launch_bounds__(kT)
fastnlm_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, const int* __restrict__ block_size,
               const int* __restrict__ search_block, const float* __restrict__ decay, int Rmax) {
  extern __shared__ float sh[];
  const int n = blockIdx.z;
  int rb = block_size[n] / 2, rs = search_block[n] / 2;
  rb = rb < 0 ? 0 : rb; rs = rs < 0 ? 0 : rs;
  if (rb + rs > Rmax) { rs = rs > Rmax ? Rmax : rs; rb = Rmax - rs; }
  const int R = rb + rs;
  const float h = decay[n];
  const float kexp = -1.f / (h * h * 3.f * (float)((2 * rb + 1) * (2 * rb + 1)));
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const long long plane = (long long)H * W;
  load_tile<3>(sh, x + (long long)n * 3 * plane, H, W, y0, x0, R, BORDER_REFLECT101, 1.f);
  __syncthreads();
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const int gx = x0 + tx, gy = y0 + ty;
  if (gx >= W || gy >= H) return;
  const int tw = TW + 2 * R, ps = tw * (TH + 2 * R);
  const float* c0 = sh + (ty + R) * tw + tx + R;
  float nb = 0.f, ng = 0.f, nr = 0.f, den = 0.f;
  for (int dy = -rs; dy <= rs; ++dy) {
    for (int dx = -rs; dx <= rs; ++dx) {
      const float* q0 = c0 + dy * tw + dx;
      float d2 = 0.f;
      for (int oy = -rb; oy <= rb; ++oy) {
        for (int ox = -rb; ox <= rb; ++ox) {
          const int o = oy * tw + ox;
          const float eb = c0[o] - q0[o], eg = c0[o + ps] - q0[o + ps], er = c0[o + 2 * ps] - q0[o + 2 * ps];
          d2 = fmaf(eb, eb, d2); d2 = fmaf(eg, eg, d2); d2 = fmaf(er, er, d2);
        }
      }
      const float wgt = __expf(d2 * kexp);
      nb = fmaf(wgt, q0[0], nb); ng = fmaf(wgt, q0[ps], ng); nr = fmaf(wgt, q0[2 * ps], nr); den += wgt;
    }
  }
  float* o = y + (long long)n * 3 * plane + (long long)gy * W + gx;
  o[0] = nb / den; o[plane] = ng / den; o[2 * plane] = nr / den;
}

// ---- median ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cswap(float& a, float& b) { float t = fminf(a, b); b = fmaxf(a, b); a = t; }

__device__ __forceinline__ float median9(float* v) {
  // classic 19-exchange network (Paeth / Smith)
  cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]); cswap(v[0], v[1]); cswap(v[3], v[4]); cswap(v[6], v[7]);
  cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]); cswap(v[0], v[3]); cswap(v[5], v[8]); cswap(v[4], v[7]);
  cswap(v[3], v[6]); cswap(v[1], v[4]); cswap(v[2], v[5]); cswap(v[4], v[7]); cswap(v[4], v[2]); cswap(v[6], v[4]);
  cswap(v[4], v[2]);
  return v[4];
}

// Exact median of K*K window elements by forgetful selection: keep t = n/2 + 2 candidates in registers; the minimum and the
// maximum of any t elements cannot be the median, so each round drops both and takes in one unseen element; 3 remain at the
// end.  ~0.75 t^2 compare-exchanges (K = 5: 147, 7: 507, 9: 1323) instead of 32 counting passes over the window.
template <int CUR, int T>
__device__ __forceinline__ void minmax_to_ends(float (&a)[T]) {      // a[0] <- min, a[CUR-1] <- max of a[0..CUR)
#pragma unroll
  for (int i = 0; i < CUR / 2; ++i) cswap(a[i], a[CUR - 1 - i]);
#pragma unroll
  for (int i = 1; i < (CUR + 1) / 2; ++i) cswap(a[0], a[i]);
#pragma unroll
  for (int i = CUR / 2; i < CUR - 1; ++i) cswap(a[i], a[CUR - 1]);
}
template <int K, int CUR, int NEXT, int T>
struct Forget {
  static __device__ __forceinline__ float run(float (&a)[T], const float* __restrict__ c0, int tw) {
    minmax_to_ends<CUR, T>(a);
    if constexpr (CUR == 3) {
      return a[1];
    } else {
      constexpr int n = K * K, R = K / 2;
      if constexpr (NEXT < n) {           // the minimum's slot takes the next unseen element; the maximum (last) is dropped
        a[0] = c0[(NEXT / K - R) * tw + (NEXT % K - R)];
        return Forget<K, CUR - 1, NEXT + 1, T>::run(a, c0, tw);
      } else {                            // nothing unseen: drop both ends (move the last kept element into slot 0)
        a[0] = a[CUR - 2];
        return Forget<K, CUR - 2, NEXT, T>::run(a, c0, tw);
      }
    }
  }
};
template <int K>
__device__ __forceinline__ float median_select(const float* __restrict__ c0, int tw) {
  constexpr int n = K * K, T = n / 2 + 2, R = K / 2;
  float a[T];
#pragma unroll
  for (int i = 0; i < T; ++i) a[i] = c0[(i / K - R) * tw + (i % K - R)];
  return Forget<K, T, T, T>::run(a, c0, tw);
}

__device__ __forceinline__ unsigned int okey(float f) {
  unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float okey_inv(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(kT)
median_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int R) {
  extern __shared__ float sh[];
  const long long plane = (long long)H * W;
  const float* img = x + (long long)blockIdx.z * plane;     // one plane per z
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  load_tile<1>(sh, img, H, W, y0, x0, R, BORDER_REPLICATE, 1.f);
  __syncthreads();
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const int gx = x0 + tx, gy = y0 + ty;
  if (gx >= W || gy >= H) return;
  const int tw = TW + 2 * R;
  const float* c0 = sh + (ty + R) * tw + tx + R;
  float out;
  if (R == 1) {
    float v[9];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) v[(dy + 1) * 3 + dx + 1] = c0[dy * tw + dx];
    out = median9(v);
  } else if (R == 2) {
    out = median_select<5>(c0, tw);
  } else if (R == 3) {
    out = median_select<7>(c0, tw);
  } else if (R == 4) {
    out = median_select<9>(c0, tw);
  } else {
    // exact selection by bit-wise binary search on order-preserving keys: find the largest key K such
    // that #(elements >= K) >= rank_from_top, i.e. the median element itself
    const int cnt = (2 * R + 1) * (2 * R + 1);
    const int need = cnt - cnt / 2;        // elements >= median (median is the (cnt/2)-th smallest, 0-based)
    unsigned int key = 0;
    for (int bit = 31; bit >= 0; --bit) {
      const unsigned int cand = key | (1u << bit);
      int ge = 0;
      for (int dy = -R; dy <= R; ++dy)
        for (int dx = -R; dx <= R; ++dx) ge += (okey(c0[dy * tw + dx]) >= cand) ? 1 : 0;
      if (ge >= need) key = cand;
    }
    out = okey_inv(key);
  }
  y[(long long)blockIdx.z * plane + (long long)gy * W + gx] = out;
}

// ---- guided filter (self-guided), two passes, separable box sums ------------------------------------------------
// Each pass stages the tile + halo, forms the horizontal window sums of every staged row once (2R+1 adds per element
// instead of (2R+1)^2 per output), then every thread adds 2R+1 of them vertically.
__global__ void __launch_bounds__(kT)
guided_ab_kernel(const float* __restrict__ x, float* __restrict__ a_out, float* __restrict__ b_out, int H, int W, int R,
                 float eps) {
  extern __shared__ float sh[];
  const long long plane = (long long)H * W;
  const float* img = x + (long long)blockIdx.z * plane;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int tw = TW + 2 * R, th = TH + 2 * R;
  float* hs = sh + tw * th;            // [th][TW] horizontal sums of x
  float* hs2 = hs + th * TW;           // ... of x^2
  load_tile<1>(sh, img, H, W, y0, x0, R, BORDER_REFLECT101, 1.f);
  __syncthreads();
  for (int i = threadIdx.x; i < th * TW; i += kT) {
    const int ly = i / TW, lx = i % TW;
    const float* row = sh + ly * tw + lx;
    float s = 0.f, s2 = 0.f;
    for (int dx = 0; dx <= 2 * R; ++dx) { const float v = row[dx]; s += v; s2 = fmaf(v, v, s2); }
    hs[i] = s; hs2[i] = s2;
  }
  __syncthreads();
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const int gx = x0 + tx, gy = y0 + ty;
  if (gx >= W || gy >= H) return;
  float s = 0.f, s2 = 0.f;
  for (int dy = 0; dy <= 2 * R; ++dy) { s += hs[(ty + dy) * TW + tx]; s2 += hs2[(ty + dy) * TW + tx]; }
  const float inv = 1.f / (float)((2 * R + 1) * (2 * R + 1));
  const float mean = s * inv, var = s2 * inv - mean * mean;
  const float a = var / (var + eps);
  const long long o = (long long)blockIdx.z * plane + (long long)gy * W + gx;
  a_out[o] = a; b_out[o] = mean - a * mean;
}

__global__ void __launch_bounds__(kT)
guided_out_kernel(const float* __restrict__ x, const float* __restrict__ a_in, const float* __restrict__ b_in,
                  float* __restrict__ y, int H, int W, int R) {
  extern __shared__ float sh[];
  const long long plane = (long long)H * W;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int tw = TW + 2 * R, th = TH + 2 * R;
  float* hsa = sh + 2 * tw * th;
  float* hsb = hsa + th * TW;
  load_tile<1>(sh, a_in + (long long)blockIdx.z * plane, H, W, y0, x0, R, BORDER_REFLECT101, 1.f);
  load_tile<1>(sh + tw * th, b_in + (long long)blockIdx.z * plane, H, W, y0, x0, R, BORDER_REFLECT101, 1.f);
  __syncthreads();
  for (int i = threadIdx.x; i < th * TW; i += kT) {
    const int ly = i / TW, lx = i % TW;
    const float* ra = sh + ly * tw + lx;
    const float* rb = ra + tw * th;
    float sa = 0.f, sb = 0.f;
    for (int dx = 0; dx <= 2 * R; ++dx) { sa += ra[dx]; sb += rb[dx]; }
    hsa[i] = sa; hsb[i] = sb;
  }
  __syncthreads();
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const int gx = x0 + tx, gy = y0 + ty;
  if (gx >= W || gy >= H) return;
  float sa = 0.f, sb = 0.f;
  for (int dy = 0; dy <= 2 * R; ++dy) { sa += hsa[(ty + dy) * TW + tx]; sb += hsb[(ty + dy) * TW + tx]; }
  const float inv = 1.f / (float)((2 * R + 1) * (2 * R + 1));
  const long long o = (long long)blockIdx.z * plane + (long long)gy * W + gx;
  y[o] = sa * inv * x[o] + sb * inv;
}

// ---- guided filter, marching version (W % 4 == 0): running column sums in registers, horizontal window by shuffles ----
// A warp owns a strip of 128 loaded columns (112 output columns + an 8-column halo on both sides = kMaxR) and walks down a
// chunk of rows.  Every lane keeps the VERTICAL running sums of its 4 columns over the 2R+1 rows of the window (one row
// enters, one row leaves per step: two coalesced 128-bit loads per lane, the leaving row comes from L2), and the HORIZONTAL
// window sum of those column sums is gathered from the neighbouring lanes by shuffles -- no shared memory, no barrier, and
// (2R+1) + 2 adds per pixel and quantity instead of the tile kernels' 2(2R+1) shared-memory reads.  Two passes as before:
// (a, b) from x, then y = mean(a) x + mean(b).
namespace gmarch {
constexpr int kHalo = kMaxR, kOutCols = 128 - 2 * kHalo, kWarps = 4;

__device__ __forceinline__ int refl(int i, int n) {
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}
// 4 consecutive columns [c, c+4) of row q (both reflect-101); c is a multiple of 4
__device__ __forceinline__ float4 load4(const float* __restrict__ plane, int H, int W, int q, int c) {
  const float* row = plane + (long long)refl(q, H) * W;
  if (c >= 0 && c + 4 <= W) return __ldg(reinterpret_cast<const float4*>(row + c));
  return make_float4(row[refl(c, W)], row[refl(c + 1, W)], row[refl(c + 2, W)], row[refl(c + 3, W)]);
}
// window sums over columns k-R .. k+R (k = the lane's 4 columns) of a quantity whose per-column values are v (this lane)
template <int R>
__device__ __forceinline__ void hbox(const float (&v)[4], float (&out)[4]) {
  constexpr int NL = (R + 3) / 4;                    // neighbour lanes needed on each side
  float e[4 * (2 * NL + 1)];
#pragma unroll
  for (int dl = -NL; dl <= NL; ++dl)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // lanes at the ends of the warp receive their own value for out-of-range sources; those lanes are halo-only
      e[(dl + NL) * 4 + i] = (dl == 0) ? v[i] : (dl < 0 ? __shfl_up_sync(0xffffffffu, v[i], -dl) : __shfl_down_sync(0xffffffffu, v[i], dl));
    }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float s = 0.f;
#pragma unroll
    for (int d = -R; d <= R; ++d) s += e[NL * 4 + k + d];
    out[k] = s;
  }
}

template <int R, int PASS>      // PASS 0: x -> (a, b);  PASS 1: (a, b, x) -> y
__global__ void __launch_bounds__(kWarps * 32)
guided_march_kernel(const float* __restrict__ x, float* __restrict__ a_buf, float* __restrict__ b_buf, float* __restrict__ y,
                    int H, int W, int rows_per_chunk, float eps) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int strip = blockIdx.x * kWarps + wid;
  if (strip * kOutCols >= W) return;
  const int c = strip * kOutCols - kHalo + lane * 4;              // first of this lane's 4 columns (may lie outside the frame)
  const long long plane = (long long)H * W;
  const float* __restrict__ p0 = (PASS == 0 ? x : a_buf) + (long long)blockIdx.z * plane;
  const float* __restrict__ p1 = (PASS == 0 ? x : b_buf) + (long long)blockIdx.z * plane;
  const int ra = blockIdx.y * rows_per_chunk, rb = min(H, ra + rows_per_chunk);
  if (ra >= rb) return;
  const bool writer = lane >= kHalo / 4 && lane < 32 - kHalo / 4 && c < W;
  float v0[4] = {0.f, 0.f, 0.f, 0.f}, v1[4] = {0.f, 0.f, 0.f, 0.f};   // running column sums: PASS 0: x, x^2 ; PASS 1: a, b
  auto add_row = [&](int q, float sgn) {
    const float4 t = load4(p0, H, W, q, c);
    const float tv[4] = {t.x, t.y, t.z, t.w};
    if (PASS == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { v0[i] = fmaf(sgn, tv[i], v0[i]); v1[i] = fmaf(sgn * tv[i], tv[i], v1[i]); }
    } else {
      const float4 u = load4(p1, H, W, q, c);
      const float uv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) { v0[i] = fmaf(sgn, tv[i], v0[i]); v1[i] = fmaf(sgn, uv[i], v1[i]); }
    }
  };
  for (int d = -R; d <= R; ++d) add_row(ra + d, 1.f);
  const float inv = 1.f / (float)((2 * R + 1) * (2 * R + 1));
  for (int r = ra; r < rb; ++r) {
    float s0[4], s1[4];
    hbox<R>(v0, s0);
    hbox<R>(v1, s1);
    if (writer) {
      const long long o = (long long)blockIdx.z * plane + (long long)r * W + c;
      if (PASS == 0) {
        float av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float mean = s0[i] * inv, var = s1[i] * inv - mean * mean;
          av[i] = var / (var + eps);
          bv[i] = mean - av[i] * mean;
        }
        *reinterpret_cast<float4*>(a_buf + o) = make_float4(av[0], av[1], av[2], av[3]);
        *reinterpret_cast<float4*>(b_buf + o) = make_float4(bv[0], bv[1], bv[2], bv[3]);
      } else {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + o));
        st_stream4(y + o, make_float4(fmaf(s0[0] * inv, xv.x, s1[0] * inv), fmaf(s0[1] * inv, xv.y, s1[1] * inv),
                                      fmaf(s0[2] * inv, xv.z, s1[2] * inv), fmaf(s0[3] * inv, xv.w, s1[3] * inv)));
      }
    }
    if (r + 1 < rb) { add_row(r + 1 + R, 1.f); add_row(r - R, -1.f); }
  }
}

template <int R>
static void launch(const float* x, float* a, float* b, float* y, int N, int H, int W, float eps, cudaStream_t st) {
  const int strips = (int)cdiv(W, kOutCols), bx = (int)cdiv(strips, kWarps);
  long long chunks = cdiv((long long)sm_count() * 8, (long long)bx * N * 3);
  int rows = (int)cdiv(H, chunks < 1 ? 1 : chunks);
  rows = rows < 48 ? 48 : rows;                                      // a chunk start costs 2R+1 rows of loads
  dim3 grid((unsigned)bx, (unsigned)cdiv(H, rows), (unsigned)(N * 3));
  guided_march_kernel<R, 0><<<grid, kWarps * 32, 0, st>>>(x, a, b, y, H, W, rows, eps);
  guided_march_kernel<R, 1><<<grid, kWarps * 32, 0, st>>>(x, a, b, y, H, W, rows, eps);
}
}  // namespace gmarch

// ---- sharpen (unsharp mask, 5x5 binomial) ------------------------------------------------------------------
__constant__ float kBinom[5] = {1.f / 16, 4.f / 16, 6.f / 16, 4.f / 16, 1.f / 16};

__device__ __forceinline__ float blur5(const float* c0, int tw) {
  float acc = 0.f;
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy) {
    float row = 0.f;
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) row = fmaf(kBinom[dx + 2], c0[dy * tw + dx], row);
    acc = fmaf(kBinom[dy + 2], row, acc);
  }
  return acc;
}

__global__ void __launch_bounds__(kT)
sharpen_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, const float* __restrict__ amount) {
  extern __shared__ float sh[];
  const long long plane = (long long)H * W;
  const int pl = blockIdx.z;                      // n*3 + c
  const float a = amount[pl / 3];
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  load_tile<1>(sh, x + (long long)pl * plane, H, W, y0, x0, 2, BORDER_REFLECT101, 1.f);
  __syncthreads();
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const int gx = x0 + tx, gy = y0 + ty;
  if (gx >= W || gy >= H) return;
  const int tw = TW + 4;
  const float* c0 = sh + (ty + 2) * tw + tx + 2;
  const float v = c0[0];
  y[(long long)pl * plane + (long long)gy * W + gx] = sat01(fmaf(a, v - blur5(c0, tw), v));
}

// pass 1 of the backward: e = dy * [0 <= u <= 1], d amount partials = sum e * (x - blur(x))
__global__ void __launch_bounds__(kT)
sharpen_bwd_e_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ e_out,
                     float* __restrict__ partial, int H, int W, const float* __restrict__ amount) {
  extern __shared__ float sh[];
  const long long plane = (long long)H * W;
  const int pl = blockIdx.z;
  const float a = amount[pl / 3];
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  load_tile<1>(sh, x + (long long)pl * plane, H, W, y0, x0, 2, BORDER_REFLECT101, 1.f);
  __syncthreads();
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const int gx = x0 + tx, gy = y0 + ty;
  float contrib = 0.f;
  if (gx < W && gy < H) {
    const int tw = TW + 4;
    const float* c0 = sh + (ty + 2) * tw + tx + 2;
    const float v = c0[0], hp = v - blur5(c0, tw);
    const long long o = (long long)pl * plane + (long long)gy * W + gx;
    const float e = dy[o] * in01(fmaf(a, hp, v));
    e_out[o] = e;
    contrib = e * hp;
  }
  __shared__ float red[kT / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  contrib = warp_sum(contrib);
  if (lane == 0) red[wid] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kT / 32; ++w) s += red[w];
    partial[((long long)pl * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
  }
}

// z(py,px) = sum_d k[d] * E[p - d] on the zero-extended e (used only for the folded border terms)
__device__ float blur_zero_ext(const float* __restrict__ e, int H, int W, int py, int px) {
  float acc = 0.f;
  for (int dy = -2; dy <= 2; ++dy) {
    const int yy = py - dy;
    if (yy < 0 || yy >= H) continue;
    for (int dx = -2; dx <= 2; ++dx) {
      const int xx = px - dx;
      if (xx < 0 || xx >= W) continue;
      acc = fmaf(kBinom[dy + 2] * kBinom[dx + 2], e[(long long)yy * W + xx], acc);
    }
  }
  return acc;
}

// pass 2: dx = (1+a) e - a * blur^T(e);  blur^T folds the reflect-101 padding back into the frame
__global__ void __launch_bounds__(kT)
sharpen_bwd_dx_kernel(const float* __restrict__ e, float* __restrict__ dx, int H, int W, const float* __restrict__ amount) {
  const long long plane = (long long)H * W;
  const int pl = blockIdx.z;
  const float a = amount[pl / 3];
  const float* ep = e + (long long)pl * plane;
  const int gx = blockIdx.x * TW + threadIdx.x % TW, gy = blockIdx.y * TH + threadIdx.x / TW;
  if (gx >= W || gy >= H) return;
  // candidate pre-images of (gy,gx) under reflect-101 inside the padded domain [-2, n+2)
  int ys[3], xs[3], ny = 0, nx = 0;
  ys[ny++] = gy; xs[nx++] = gx;
  if (gy >= 1 && gy <= 2) ys[ny++] = -gy;
  if (gy <= H - 2 && gy >= H - 3) ys[ny++] = 2 * H - 2 - gy;
  if (gx >= 1 && gx <= 2) xs[nx++] = -gx;
  if (gx <= W - 2 && gx >= W - 3) xs[nx++] = 2 * W - 2 - gx;
  float bt = 0.f;
  for (int i = 0; i < ny; ++i)
    for (int j = 0; j < nx; ++j) bt += blur_zero_ext(ep, H, W, ys[i], xs[j]);
  const long long o = (long long)pl * plane + (long long)gy * W + gx;
  dx[o] = fmaf(1.f + a, ep[(long long)gy * W + gx], -a * bt);
}

// ---- register-marching fast paths (risp_march.cuh) -----------------------------------------------------------
struct Median3F {
  // exact median of 9 with the columns shared by the 4 pixels of a lane: sort every window column (3 compare-exchanges),
  // then  median9 = med3( max of the column minima, med3 of the column medians, min of the column maxima )  -- 21 min/max per
  // pixel instead of the 38 of a 19-exchange network per pixel
  __device__ __forceinline__ static float med3(float a, float b, float c) { return fmaxf(fminf(a, b), fminf(fmaxf(a, b), c)); }
  __device__ __forceinline__ void operator()(const float (&w)[1][3][6], float (&o)[1][4], int) const {
    float lo[6], mid[6], hi[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float a = w[0][0][i], b = w[0][1][i], c = w[0][2][i];
      const float mn = fminf(a, b), mx = fmaxf(a, b);
      lo[i] = fminf(mn, c);
      hi[i] = fmaxf(mx, c);
      mid[i] = fmaxf(mn, fminf(mx, c));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      o[0][k] = med3(fmaxf(fmaxf(lo[k], lo[k + 1]), lo[k + 2]), med3(mid[k], mid[k + 1], mid[k + 2]),
                     fminf(fminf(hi[k], hi[k + 1]), hi[k + 2]));
  }
};
struct SharpenF {
  const float* amount;
  __device__ __forceinline__ void operator()(const float (&w)[1][5][8], float (&o)[1][4], int zp) const {
    const float a = __ldg(amount + zp / 3);
    const float kb[5] = {1.f / 16, 4.f / 16, 6.f / 16, 4.f / 16, 1.f / 16};
    float col[8];                                   // vertical pass once per window column, shared by the 4 pixels
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float c = 0.f;
#pragma unroll
      for (int j = 0; j < 5; ++j) c = fmaf(kb[j], w[0][j][i], c);
      col[i] = c;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float b = 0.f;
#pragma unroll
      for (int i = 0; i < 5; ++i) b = fmaf(kb[i], col[k + i], b);
      const float v = w[0][2][k + 2];
      o[0][k] = sat01(fmaf(a, v - b, v));
    }
  }
};
// window 3 (the only size the reference wrapper can produce for p in [0,1): tools_origin.py:698) or 1, per image
struct Bilateral3F {
  const int* window; const float* sigma_color; const float* sigma_space;
  __device__ __forceinline__ void operator()(const float (&w)[3][3][6], float (&o)[3][4], int n) const {
    const int R = __ldg(window + n) / 2;
    const float sc = __ldg(sigma_color + n), ss = __ldg(sigma_space + n);
    const float kc = -0.5f / (sc * sc), ks = -0.5f / (ss * ss);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float cb = w[0][1][k + 1], cg = w[1][1][k + 1], cr = w[2][1][k + 1];
      float nb = cb, ng = cg, nr = cr, den = 1.f;                 // centre tap: weight exp(0)
      if (R >= 1) {                                             // circular support r^2 <= R^2: the 4-neighbour cross
        const int dys[4] = {-1, 1, 0, 0}, dxs[4] = {0, 0, -1, 1};
        const float ws1 = __expf(ks);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float vb = w[0][1 + dys[t]][k + 1 + dxs[t]], vg = w[1][1 + dys[t]][k + 1 + dxs[t]], vr = w[2][1 + dys[t]][k + 1 + dxs[t]];
          const float dist = fabsf(vb - cb) + fabsf(vg - cg) + fabsf(vr - cr);
          const float wgt = __expf(dist * dist * kc) * ws1;
          nb = fmaf(wgt, vb, nb); ng = fmaf(wgt, vg, ng); nr = fmaf(wgt, vr, nr); den += wgt;
        }
      }
      const float inv = 1.f / den;
      o[0][k] = nb * inv; o[1][k] = ng * inv; o[2][k] = nr * inv;
    }
  }
};

static dim3 tile_grid(int H, int W, int Z) { return dim3((unsigned)cdiv(W, TW), (unsigned)cdiv(H, TH), (unsigned)Z); }
static size_t tile_smem(int R, int planes) { return sizeof(float) * (size_t)planes * (TW + 2 * R) * (TH + 2 * R); }

}  // namespace risp

using namespace risp;

extern "C" int risp_bilateral_fwd(const float* x, float* y, int N, int H, int W, const int* window,
                                  const float* sigma_color, const float* sigma_space, int max_window,
                                  risp_stream_t stream) {
  RISP_REQUIRE(x && y && window && sigma_color && sigma_space && N > 0 && H > 1 && W > 1, RISP_E_INVALID,
               "risp_bilateral_fwd: bad arguments");
  RISP_REQUIRE(max_window >= 1 && max_window <= 2 * kMaxR + 1 && (max_window & 1), RISP_E_INVALID,
               "risp_bilateral_fwd: max_window %d must be odd and <= %d", max_window, 2 * kMaxR + 1);
  RISP_REQUIRE(N <= 65535, RISP_E_INVALID, "risp_bilateral_fwd: batch too large");
  const int R = max_window / 2;
  if (R <= 1 && march::usable(x, y, H, W, 1)) {
    const march::Geom g = march::geometry(H, W, N);
    march::march_kernel<1, 3, march::REFLECT101, Bilateral3F><<<g.grid, march::kWarps * 32, 0, as_stream(stream)>>>(
        x, y, H, W, g.rows_per_chunk, Bilateral3F{window, sigma_color, sigma_space});
    return check_launch("march_kernel<bilateral3>");
  }
  bilateral_kernel<<<tile_grid(H, W, N), kT, tile_smem(R, 3), as_stream(stream)>>>(x, y, H, W, window, sigma_color,
                                                                                sigma_space, R);
  return check_launch("bilateral_kernel");
}

extern "C" int risp_fastnlm_fwd(const float* x, float* y, int N, int H, int W, const int* block_size,
                                const int* search_block, const float* decay_factor, int max_halo, risp_stream_t stream) {
  RISP_REQUIRE(x && y && block_size && search_block && decay_factor && N > 0 && H > 1 && W > 1, RISP_E_INVALID,
               "risp_fastnlm_fwd: bad arguments");
  RISP_REQUIRE(max_halo >= 0 && max_halo <= 2 * kMaxR, RISP_E_INVALID, "risp_fastnlm_fwd: max_halo %d not in [0,%d]", max_halo,
               2 * kMaxR);
  RISP_REQUIRE(N <= 65535, RISP_E_INVALID, "risp_fastnlm_fwd: batch too large");
  fastnlm_kernel<<<tile_grid(H, W, N), kT, tile_smem(max_halo, 3), as_stream(stream)>>>(x, y, H, W, block_size, search_block,
                                                                                     decay_factor, max_halo);
  return check_launch("fastnlm_kernel");
}

extern "C" int risp_median_fwd(const float* x, float* y, int N, int H, int W, int size, risp_stream_t stream) {
  RISP_REQUIRE(x && y && N > 0 && H > 0 && W > 0, RISP_E_INVALID, "risp_median_fwd: bad arguments");
  RISP_REQUIRE(size >= 3 && size <= 2 * kMaxR + 1 && (size & 1), RISP_E_INVALID,
               "risp_median_fwd: size %d must be odd and in [3,%d]", size, 2 * kMaxR + 1);
  RISP_REQUIRE((long long)N * 3 <= 65535, RISP_E_INVALID, "risp_median_fwd: batch too large");
  const int R = size / 2;
  if (R == 1 && march::usable(x, y, H, W, 1)) {
    const march::Geom g = march::geometry(H, W, N * 3);
    march::march_kernel<1, 1, march::REPLICATE, Median3F><<<g.grid, march::kWarps * 32, 0, as_stream(stream)>>>(
        x, y, H, W, g.rows_per_chunk, Median3F{});
    return check_launch("march_kernel<median3>");
  }
  median_kernel<<<tile_grid(H, W, N * 3), kT, tile_smem(R, 1), as_stream(stream)>>>(x, y, H, W, R);
  return check_launch("median_kernel");
}

extern "C" size_t risp_guided_workspace(int N, int H, int W) {
  return (N > 0 && H > 0 && W > 0) ? (size_t)2 * N * 3 * H * W * sizeof(float) : 0;
}

extern "C" int risp_guided_fwd(const float* x, float* y, int N, int H, int W, int radius, float eps, void* workspace,
                               size_t workspace_bytes, risp_stream_t stream) {
  RISP_REQUIRE(x && y && N > 0 && H > 1 && W > 1, RISP_E_INVALID, "risp_guided_fwd: bad arguments");
  RISP_REQUIRE(radius >= 1 && radius <= kMaxR, RISP_E_INVALID, "risp_guided_fwd: radius %d not in [1,%d]", radius, kMaxR);
  RISP_REQUIRE(workspace && workspace_bytes >= risp_guided_workspace(N, H, W), RISP_E_WORKSPACE,
               "risp_guided_fwd: workspace too small");
  RISP_REQUIRE((long long)N * 3 <= 65535, RISP_E_INVALID, "risp_guided_fwd: batch too large");
  cudaStream_t st = as_stream(stream);
  float* a = static_cast<float*>(workspace);
  float* b = a + (size_t)N * 3 * H * W;
  if (W % 4 == 0 && aligned16(x) && aligned16(y) && aligned16(a) && aligned16(b) && H > radius && W > radius && cdiv(H, 48) <= 65535) {
    switch (radius) {
      case 1: gmarch::launch<1>(x, a, b, y, N, H, W, eps, st); break;
      case 2: gmarch::launch<2>(x, a, b, y, N, H, W, eps, st); break;
      case 3: gmarch::launch<3>(x, a, b, y, N, H, W, eps, st); break;
      case 4: gmarch::launch<4>(x, a, b, y, N, H, W, eps, st); break;
      case 5: gmarch::launch<5>(x, a, b, y, N, H, W, eps, st); break;
      case 6: gmarch::launch<6>(x, a, b, y, N, H, W, eps, st); break;
      case 7: gmarch::launch<7>(x, a, b, y, N, H, W, eps, st); break;
      default: gmarch::launch<8>(x, a, b, y, N, H, W, eps, st); break;
    }
    return check_launch("guided_march_kernel");
  }
  guided_ab_kernel<<<tile_grid(H, W, N * 3), kT, tile_smem(radius, 1) + 2 * sizeof(float) * (TH + 2 * radius) * TW, st>>>(x, a, b, H, W, radius, eps);
  guided_out_kernel<<<tile_grid(H, W, N * 3), kT, tile_smem(radius, 2) + 2 * sizeof(float) * (TH + 2 * radius) * TW, st>>>(x, a, b, y, H, W, radius);
  return check_launch("guided kernels");
}

extern "C" int risp_sharpen_fwd(const float* x, float* y, int N, int H, int W, const float* amount,
                                risp_stream_t stream) {
  RISP_REQUIRE(x && y && amount && N > 0 && H > 2 && W > 2, RISP_E_INVALID, "risp_sharpen_fwd: bad arguments");
  RISP_REQUIRE((long long)N * 3 <= 65535, RISP_E_INVALID, "risp_sharpen_fwd: batch too large");
  if (march::usable(x, y, H, W, 2)) {
    const march::Geom g = march::geometry(H, W, N * 3);
    march::march_kernel<2, 1, march::REFLECT101, SharpenF, 2><<<g.grid, march::kWarps * 32, 0, as_stream(stream)>>>(
        x, y, H, W, g.rows_per_chunk, SharpenF{amount});
    return check_launch("march_kernel<sharpen>");
  }
  sharpen_fwd_kernel<<<tile_grid(H, W, N * 3), kT, tile_smem(2, 1), as_stream(stream)>>>(x, y, H, W, amount);
  return check_launch("sharpen_fwd_kernel");
}

extern "C" size_t risp_sharpen_bwd_workspace(int N, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  size_t tiles = (size_t)cdiv(W, TW) * cdiv(H, TH);
  return ((size_t)N * 3 * H * W + (size_t)N * 3 * tiles) * sizeof(float);
}

extern "C" int risp_sharpen_bwd(const float* x, const float* dy, float* dx, float* damount, int N, int H, int W,
                                const float* amount, void* workspace, size_t workspace_bytes, risp_stream_t stream) {
  RISP_REQUIRE(x && dy && dx && damount && amount && N > 0 && H > 2 && W > 2, RISP_E_INVALID, "risp_sharpen_bwd: bad arguments");
  RISP_REQUIRE(workspace && workspace_bytes >= risp_sharpen_bwd_workspace(N, H, W), RISP_E_WORKSPACE,
               "risp_sharpen_bwd: workspace too small");
  RISP_REQUIRE((long long)N * 3 <= 65535, RISP_E_INVALID, "risp_sharpen_bwd: batch too large");
  cudaStream_t st = as_stream(stream);
  float* e = static_cast<float*>(workspace);
  float* partial = e + (size_t)N * 3 * H * W;
  dim3 grid = tile_grid(H, W, N * 3);
  const int tiles = grid.x * grid.y;
  sharpen_bwd_e_kernel<<<grid, kT, tile_smem(2, 1), st>>>(x, dy, e, partial, H, W, amount);
  sharpen_bwd_dx_kernel<<<grid, kT, 0, st>>>(e, dx, H, W, amount);
  int rc = check_launch("sharpen_bwd kernels");
  if (rc != RISP_OK) return rc;
  // damount[n] = sum over the 3 planes x tiles of image n: rows = N, B = 3*tiles, one slot
  const short zero = 0;
  return finalize_partials(partial, damount, N, 3 * tiles, 1, 1, &zero, &zero, 1, 1.f, false, st);
}
