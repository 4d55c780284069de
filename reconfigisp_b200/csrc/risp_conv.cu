// Dense convolution of the CNN candidates (srcnn_res_arch.py, srcnn_demosaic_arch.py, path_14l_*_arch.py):
// NCHW fp32, stride 1, "same" zero padding, K in {1,3,5,9}, with the fusions those architectures need.
//
// Round-1 kernel: direct convolution on the FP32 pipe (exact fp32 accumulation, which the 1e-4 parity bar
// needs; cuDNN's fp32 engines on B200 are ~3e-3 off, see modules/conv.py).  A CTA owns a 16x32 pixel tile x 16
// output channels; every thread accumulates 4 pixels x 16 channels in registers; input channels stream through
// shared memory in chunks (tile + halo, zero padding and the input-side ReLU / mask applied while staging),
// weights in the [ci][tap][co] layout so a warp reads them as broadcast 128-bit loads.  FMA : LDS = ~10 : 1.
// The data gradient is the same kernel on flipped / transposed weights.
#include "risp_common.cuh"

namespace risp {

constexpr int CT_Y = 16, CT_X = 32, CT_THREADS = 128, CT_CO = 16, CT_PX = 4;

template <int K> struct ConvCfg {
  static constexpr int CI = (K == 1) ? 8 : (K == 3) ? 8 : (K == 5) ? 4 : 2;   // input channels per smem chunk
  static constexpr int PITCH = ((CT_X + K - 1) + 3) / 4 * 4;                  // floats per staged row (16 B aligned)
  static constexpr int ROWS = CT_Y + K - 1;
  static constexpr int IN_FLOATS = CI * ROWS * PITCH;
  static constexpr int W_FLOATS = CI * K * K * CT_CO;
};

struct ConvArgs {
  const float* x;        // (N,Cin,H,W)
  const float* mask_in;  // nullable: x is multiplied by [mask_in > 0] while staging (backward through an output ReLU)
  const float* wk;       // (Cin, K*K, CoutPad) prepared weights, CoutPad multiple of 16
  const float* bias;     // nullable (Cout)
  const float* res;      // nullable (N,Cout,H,W)
  const float* mask_out; // nullable: the result is multiplied by [mask_out > 0] (backward through an input ReLU)
  float* y;              // (N,Cout,H,W)
  int Cin, Cout, CoutPad, H, W, flags;
};

template <int K>
__global__ void __launch_bounds__(CT_THREADS)
conv2d_kernel(ConvArgs a) {
  using C = ConvCfg<K>;
  constexpr int PAD = K / 2;
  extern __shared__ __align__(16) float smem[];
  float* s_in = smem;
  float* s_w = smem + C::IN_FLOATS;
  const int tiles_x = (a.W + CT_X - 1) / CT_X;
  const int x0 = (blockIdx.x % tiles_x) * CT_X, y0 = (blockIdx.x / tiles_x) * CT_Y;
  const int co0 = blockIdx.y * CT_CO;
  const int n = blockIdx.z;
  const int tx = threadIdx.x % (CT_X / CT_PX), ty = threadIdx.x / (CT_X / CT_PX);
  const long long plane = (long long)a.H * a.W;
  const float* xin = a.x + (long long)n * a.Cin * plane;
  const float* min_ = a.mask_in ? a.mask_in + (long long)n * a.Cin * plane : nullptr;
  const bool relu_in = (a.flags & RISP_CONV_RELU_IN) != 0;

  float acc[CT_PX][CT_CO];
#pragma unroll
  for (int p = 0; p < CT_PX; ++p)
#pragma unroll
    for (int c = 0; c < CT_CO; ++c) acc[p][c] = 0.f;

  for (int c0 = 0; c0 < a.Cin; c0 += C::CI) {
    __syncthreads();
    // stage the input chunk: CI planes x ROWS x (CT_X + K - 1) with zero padding outside the frame
    for (int i = threadIdx.x; i < C::CI * C::ROWS * C::PITCH; i += CT_THREADS) {
      const int col = i % C::PITCH;
      const int row = (i / C::PITCH) % C::ROWS;
      const int ci = i / (C::PITCH * C::ROWS);
      const int gx = x0 - PAD + col, gy = y0 - PAD + row, gc = c0 + ci;
      float v = 0.f;
      if (gc < a.Cin && col < CT_X + K - 1 && gx >= 0 && gx < a.W && gy >= 0 && gy < a.H) {
        const long long o = (long long)gc * plane + (long long)gy * a.W + gx;
        v = __ldg(xin + o);
        if (relu_in) v = fmaxf(v, 0.f);
        if (min_) v = (__ldg(min_ + o) > 0.f) ? v : 0.f;
      }
      s_in[i] = v;
    }
    // stage the weight chunk: [ci][tap][16 co]
    for (int i = threadIdx.x; i < C::W_FLOATS; i += CT_THREADS) {
      const int co = i % CT_CO;
      const int tap = (i / CT_CO) % (K * K);
      const int ci = i / (CT_CO * K * K);
      const int gc = c0 + ci;
      s_w[i] = (gc < a.Cin) ? __ldg(a.wk + ((long long)gc * K * K + tap) * a.CoutPad + co0 + co) : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < C::CI; ++ci) {
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        constexpr int NV = (CT_PX + K - 1 + 3) / 4;      // 128-bit loads covering the 4+K-1 inputs of this row
        float xv[NV * 4];
        const float4* rowp = reinterpret_cast<const float4*>(s_in + (ci * C::ROWS + ty + ky) * C::PITCH + tx * CT_PX);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          // the last vector of the last thread may reach past the staged row; PITCH is padded for it
          const float4 t = rowp[v];
          xv[4 * v] = t.x; xv[4 * v + 1] = t.y; xv[4 * v + 2] = t.z; xv[4 * v + 3] = t.w;
        }
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const float4* wp = reinterpret_cast<const float4*>(s_w + (ci * K * K + ky * K + kx) * CT_CO);
          float wv[CT_CO];
#pragma unroll
          for (int v = 0; v < CT_CO / 4; ++v) {
            const float4 t = wp[v];
            wv[4 * v] = t.x; wv[4 * v + 1] = t.y; wv[4 * v + 2] = t.z; wv[4 * v + 3] = t.w;
          }
#pragma unroll
          for (int p = 0; p < CT_PX; ++p)
#pragma unroll
            for (int c = 0; c < CT_CO; ++c) acc[p][c] = fmaf(xv[p + kx], wv[c], acc[p][c]);
        }
      }
    }
  }

  // epilogue: bias, output ReLU, residual (optionally relu'd), output mask; 128-bit stores along x
  const int gy = y0 + ty, gx = x0 + tx * CT_PX;
  if (gy >= a.H || gx >= a.W) return;
  const bool relu_out = (a.flags & RISP_CONV_RELU_OUT) != 0, add_res = (a.flags & RISP_CONV_ADD_RES) != 0,
             res_relu = (a.flags & RISP_CONV_RES_RELU) != 0;
  const bool full = (gx + CT_PX <= a.W) && ((a.W & 3) == 0);
#pragma unroll
  for (int c = 0; c < CT_CO; ++c) {
    const int co = co0 + c;
    if (co >= a.Cout) break;
    const long long o = ((long long)n * a.Cout + co) * plane + (long long)gy * a.W + gx;
    const float b = a.bias ? __ldg(a.bias + co) : 0.f;
    float v[CT_PX];
#pragma unroll
    for (int p = 0; p < CT_PX; ++p) {
      v[p] = acc[p][c] + b;
      if (relu_out) v[p] = fmaxf(v[p], 0.f);
    }
    if (full) {
      if (add_res) {
        const float4 r = *reinterpret_cast<const float4*>(a.res + o);
        v[0] += res_relu ? fmaxf(r.x, 0.f) : r.x; v[1] += res_relu ? fmaxf(r.y, 0.f) : r.y;
        v[2] += res_relu ? fmaxf(r.z, 0.f) : r.z; v[3] += res_relu ? fmaxf(r.w, 0.f) : r.w;
      }
      if (a.mask_out) {
        const float4 m = *reinterpret_cast<const float4*>(a.mask_out + o);
        v[0] = m.x > 0.f ? v[0] : 0.f; v[1] = m.y > 0.f ? v[1] : 0.f; v[2] = m.z > 0.f ? v[2] : 0.f; v[3] = m.w > 0.f ? v[3] : 0.f;
      }
      *reinterpret_cast<float4*>(a.y + o) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int p = 0; p < CT_PX; ++p) {
        if (gx + p < a.W) {
          float t = v[p];
          if (add_res) { const float r = a.res[o + p]; t += res_relu ? fmaxf(r, 0.f) : r; }
          if (a.mask_out) t = a.mask_out[o + p] > 0.f ? t : 0.f;
          a.y[o + p] = t;
        }
      }
    }
  }
}

// weights (Cout,Cin,K,K) -> wk (Cin', K*K, CoutPad') ; transpose_flip builds the data-gradient operator
__global__ void conv_prepare_weights_kernel(const float* __restrict__ w, float* __restrict__ wk, int Cin, int Cout, int K,
                                            int transpose_flip, int rows, int colsPad) {
  const long long total = (long long)rows * K * K * colsPad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % colsPad);
    const int tap = (int)((i / colsPad) % (K * K));
    const int r = (int)(i / ((long long)colsPad * K * K));
    float v = 0.f;
    if (!transpose_flip) {                 // rows = Cin, cols = Cout
      if (c < Cout) v = w[((long long)c * Cin + r) * K * K + tap];
    } else {                               // rows = Cout (input of the gradient), cols = Cin, taps flipped
      if (c < Cin) v = w[((long long)r * Cin + c) * K * K + (K * K - 1 - tap)];
    }
    wk[i] = v;
  }
}

template <int K>
static int launch_conv(const ConvArgs& a, int N, cudaStream_t st) {
  using C = ConvCfg<K>;
  const size_t smem = sizeof(float) * (C::IN_FLOATS + C::W_FLOATS);
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    cudaFuncSetAttribute(conv2d_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  dim3 grid((unsigned)(cdiv(a.W, CT_X) * cdiv(a.H, CT_Y)), (unsigned)(a.CoutPad / CT_CO), (unsigned)N);
  conv2d_kernel<K><<<grid, CT_THREADS, smem, st>>>(a);
  return check_launch("conv2d_kernel");
}

}  // namespace risp

using namespace risp;

extern "C" size_t risp_conv2d_prepared_weight_floats(int Cin, int Cout, int K, int transpose_flip) {
  const int rows = transpose_flip ? Cout : Cin, cols = transpose_flip ? Cin : Cout;
  return (size_t)rows * K * K * ((cols + CT_CO - 1) / CT_CO * CT_CO);
}

extern "C" int risp_conv2d_prepare_weights(const float* weight, float* wk, int Cin, int Cout, int K, int transpose_flip,
                                           risp_stream_t stream) {
  RISP_REQUIRE(weight && wk && Cin > 0 && Cout > 0 && (K == 1 || K == 3 || K == 5 || K == 9), RISP_E_INVALID,
               "risp_conv2d_prepare_weights: bad arguments (K must be 1, 3, 5 or 9)");
  const int rows = transpose_flip ? Cout : Cin, cols = transpose_flip ? Cin : Cout;
  const int colsPad = (cols + CT_CO - 1) / CT_CO * CT_CO;
  const long long total = (long long)rows * K * K * colsPad;
  conv_prepare_weights_kernel<<<(int)cdiv(total, 256), 256, 0, as_stream(stream)>>>(weight, wk, Cin, Cout, K, transpose_flip,
                                                                                   rows, colsPad);
  return check_launch("conv_prepare_weights_kernel");
}

extern "C" int risp_conv2d_fwd(const float* x, const float* mask_in, const float* wk, const float* bias, const float* res,
                               const float* mask_out, float* y, int N, int Cin, int Cout, int H, int W, int K, int flags,
                               risp_stream_t stream) {
  RISP_REQUIRE(x && wk && y && N > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, RISP_E_INVALID, "risp_conv2d_fwd: bad arguments");
  RISP_REQUIRE(N <= 65535, RISP_E_INVALID, "risp_conv2d_fwd: batch too large");
  RISP_REQUIRE(!(flags & RISP_CONV_ADD_RES) || res, RISP_E_INVALID, "risp_conv2d_fwd: ADD_RES without a residual tensor");
  RISP_REQUIRE(aligned16(y) && (!res || aligned16(res)) && (!mask_out || aligned16(mask_out)), RISP_E_ALIGN,
               "risp_conv2d_fwd: y / res / mask_out must be 16-byte aligned");
  ConvArgs a{x, mask_in, wk, bias, res, mask_out, y, Cin, Cout, (Cout + CT_CO - 1) / CT_CO * CT_CO, H, W, flags};
  cudaStream_t st = as_stream(stream);
  switch (K) {
    case 1: return launch_conv<1>(a, N, st);
    case 3: return launch_conv<3>(a, N, st);
    case 5: return launch_conv<5>(a, N, st);
    case 9: return launch_conv<9>(a, N, st);
    default: set_error("risp_conv2d_fwd: kernel size %d not in {1,3,5,9}", K); return RISP_E_UNSUPPORTED;
  }
}

// ---- weight gradient (proxy fine-tuning, darts_ft_model.py:206-246) -------------------------------------
//   dW[co][ci][ky][kx] = sum_{n,y,x} dy'[n][co][y][x] * x'[n][ci][y+ky-P][x+kx-P],   dy' = dy*[mask_dy>0], x' = relu?(x)
// A CTA owns one input channel, a block of 16 output channels and a strided set of 8x64 pixel tiles; thread
// (co, ky) keeps the K partial sums over kx in registers.  Partials go to [tile_group][Cout][Cin][K*K] and are
// finished by the deterministic finaliser (no float atomics).
namespace risp {
constexpr int WG_TY = 8, WG_TX = 64, WG_CO = 16, WG_GROUPS = 16;

template <int K>
__global__ void __launch_bounds__(WG_CO * K)
conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mask_dy,
                  float* __restrict__ partial, int N, int Cin, int Cout, int H, int W, int relu_in) {
  constexpr int PAD = K / 2, XW = WG_TX + K - 1, XH = WG_TY + K - 1;
  __shared__ float s_x[XH][XW];
  __shared__ float s_dy[WG_CO][WG_TY][WG_TX];
  const int ci = blockIdx.y, co0 = blockIdx.z * WG_CO;
  const int co_l = threadIdx.x / K, ky = threadIdx.x % K;
  const int tiles_x = (W + WG_TX - 1) / WG_TX, tiles_y = (H + WG_TY - 1) / WG_TY;
  const int n_tiles = N * tiles_y * tiles_x;
  const long long plane = (long long)H * W;
  float acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.f;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int n = t / (tiles_y * tiles_x);
    const int y0 = ((t / tiles_x) % tiles_y) * WG_TY, x0 = (t % tiles_x) * WG_TX;
    __syncthreads();
    for (int i = threadIdx.x; i < XH * XW; i += blockDim.x) {
      const int r = i / XW, c = i % XW;
      const int gy = y0 - PAD + r, gx = x0 - PAD + c;
      float v = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(x + ((long long)n * Cin + ci) * plane + (long long)gy * W + gx);
      s_x[r][c] = relu_in ? fmaxf(v, 0.f) : v;
    }
    for (int i = threadIdx.x; i < WG_CO * WG_TY * WG_TX; i += blockDim.x) {
      const int c = i % WG_TX, r = (i / WG_TX) % WG_TY, co = i / (WG_TX * WG_TY);
      const int gy = y0 + r, gx = x0 + c;
      float v = 0.f;
      if (co0 + co < Cout && gy < H && gx < W) {
        const long long o = ((long long)n * Cout + co0 + co) * plane + (long long)gy * W + gx;
        v = __ldg(dy + o);
        if (mask_dy) v = (__ldg(mask_dy + o) > 0.f) ? v : 0.f;
      }
      s_dy[co][r][c] = v;
    }
    __syncthreads();
    for (int r = 0; r < WG_TY; ++r) {
      float win[K];
#pragma unroll
      for (int k = 0; k < K - 1; ++k) win[k + 1] = s_x[r + ky][k];
      for (int c = 0; c < WG_TX; ++c) {
#pragma unroll
        for (int k = 0; k < K - 1; ++k) win[k] = win[k + 1];
        win[K - 1] = s_x[r + ky][c + K - 1];
        const float g = s_dy[co_l][r][c];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = fmaf(g, win[k], acc[k]);
      }
    }
  }
  if (co0 + co_l < Cout) {
    float* o = partial + (((long long)blockIdx.x * Cout + co0 + co_l) * Cin + ci) * K * K + ky * K;
#pragma unroll
    for (int k = 0; k < K; ++k) o[k] = acc[k];
  }
}

__global__ void wgrad_final_kernel(const float* __restrict__ partial, float* __restrict__ dw, long long n, int groups) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int g = 0; g < groups; ++g) s += partial[(long long)g * n + i];
    dw[i] = s;
  }
}

template <int K>
static int launch_wgrad(const float* x, const float* dy, const float* mask_dy, float* partial, float* dw, int N, int Cin,
                        int Cout, int H, int W, int relu_in, cudaStream_t st) {
  dim3 grid(WG_GROUPS, Cin, (Cout + WG_CO - 1) / WG_CO);
  conv_wgrad_kernel<K><<<grid, WG_CO * K, 0, st>>>(x, dy, mask_dy, partial, N, Cin, Cout, H, W, relu_in);
  const long long n = (long long)Cout * Cin * K * K;
  wgrad_final_kernel<<<(int)cdiv(n, 256), 256, 0, st>>>(partial, dw, n, WG_GROUPS);
  return check_launch("conv_wgrad_kernel");
}
}  // namespace risp

extern "C" size_t risp_conv2d_bwd_weight_workspace(int Cin, int Cout, int K) {
  return (size_t)risp::WG_GROUPS * Cout * Cin * K * K * sizeof(float);
}

extern "C" int risp_conv2d_bwd_weight(const float* x, const float* dy, const float* mask_dy, float* dweight, int N, int Cin,
                                      int Cout, int H, int W, int K, int relu_in, void* workspace, size_t workspace_bytes,
                                      risp_stream_t stream) {
  RISP_REQUIRE(x && dy && dweight && N > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0, RISP_E_INVALID, "risp_conv2d_bwd_weight: bad arguments");
  RISP_REQUIRE(Cin <= 65535 && Cout <= 16 * 65535, RISP_E_INVALID, "risp_conv2d_bwd_weight: too many channels");
  RISP_REQUIRE(workspace && workspace_bytes >= risp_conv2d_bwd_weight_workspace(Cin, Cout, K), RISP_E_WORKSPACE,
               "risp_conv2d_bwd_weight: workspace too small");
  cudaStream_t st = as_stream(stream);
  float* partial = static_cast<float*>(workspace);
  switch (K) {
    case 1: return launch_wgrad<1>(x, dy, mask_dy, partial, dweight, N, Cin, Cout, H, W, relu_in, st);
    case 3: return launch_wgrad<3>(x, dy, mask_dy, partial, dweight, N, Cin, Cout, H, W, relu_in, st);
    case 5: return launch_wgrad<5>(x, dy, mask_dy, partial, dweight, N, Cin, Cout, H, W, relu_in, st);
    case 9: return launch_wgrad<9>(x, dy, mask_dy, partial, dweight, N, Cin, Cout, H, W, relu_in, st);
    default: set_error("risp_conv2d_bwd_weight: kernel size %d not in {1,3,5,9}", K); return RISP_E_UNSUPPORTED;
  }
}
