// Patch split / merge on the device (utils/util_path_restore.py:47-134, test_split.py:82-108).
// The reference tiles a frame on the host in numpy, pushes the tiles through the model one by one with
// a H2D + D2H per tile, and blends on the host.  Here both directions are single launches on NCHW
// device tensors; the merge is a GATHER (each output pixel sums the tiles that cover it, in the
// reference's tile order) so it needs no atomics and reproduces the reference's fp32 rounding.
#include "risp_common.cuh"

namespace risp {

constexpr int kT = 256;
constexpr int kMaxOrigins = 256;

struct Origins {
  int ny, nx;
  int ys[kMaxOrigins];
  int xs[kMaxOrigins];
};

__global__ void __launch_bounds__(kT)
whole2patch_kernel(const float* __restrict__ frame, float* __restrict__ tiles, int C, int H, int W, int h, int w,
                   Origins o, long long total) {
  for (long long t = (long long)blockIdx.x * kT + threadIdx.x; t < total; t += (long long)gridDim.x * kT) {
    const int tx = (int)(t % w);
    long long r = t / w;
    const int ty = (int)(r % h); r /= h;
    const int c = (int)(r % C);
    const int tile = (int)(r / C);
    const int iy = tile / o.nx, ix = tile % o.nx;
    tiles[t] = frame[((long long)c * H + o.ys[iy] + ty) * W + o.xs[ix] + tx];
  }
}

// ramp of create_patch_mask: (i+1)/(e+1) near the low edge, mirrored at the high edge, 1 inside
__device__ __forceinline__ float ramp(int i, int n, int e) {
  int d = i < n - 1 - i ? i : n - 1 - i;   // distance to the nearer edge
  return d < e ? __fdiv_rn((float)(d + 1), (float)(e + 1)) : 1.f;
}

__global__ void __launch_bounds__(kT)
patch2whole_kernel(const float* __restrict__ tiles, float* __restrict__ frame, unsigned char* __restrict__ frame_u8, int C, int H,
                   int W, int h, int w, int eh, int ew, Origins o, int clip01) {
  const long long plane = (long long)H * W;
  for (long long p = (long long)blockIdx.x * kT + threadIdx.x; p < plane; p += (long long)gridDim.x * kT) {
    const int y = (int)(p / W), x = (int)(p % W);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float cnt = 0.f;
    for (int iy = 0; iy < o.ny; ++iy) {
      const int ty = y - o.ys[iy];
      if (ty < 0 || ty >= h) continue;
      const float my = ramp(ty, h, eh);
      for (int ix = 0; ix < o.nx; ++ix) {
        const int tx = x - o.xs[ix];
        if (tx < 0 || tx >= w) continue;
        const float m = fminf(my, ramp(tx, w, ew));
        cnt = __fadd_rn(cnt, m);
        const long long base = (((long long)(iy * o.nx + ix) * C) * h + ty) * w + tx;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < C) acc[c] = __fadd_rn(acc[c], __fmul_rn(tiles[base + (long long)c * h * w], m));
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < C) {
        float v = __fdiv_rn(acc[c], cnt);
        if (clip01) v = sat01(v);
        if (frame) frame[(long long)c * plane + p] = v;
        // (np.clip(x, 0, 1) * 255.).astype(np.uint8): truncation, HWC (test_split.py:107, utils/util.py:118-135)
        if (frame_u8) frame_u8[p * C + c] = (unsigned char)(int)__fmul_rn(sat01(v), 255.f);
      }
    }
  }
}

static int fill_origins(Origins* o, const char* who, const int* ys, int ny, const int* xs, int nx, int H, int W, int h,
                        int w) {
  RISP_REQUIRE(ys && xs && ny > 0 && nx > 0 && ny <= kMaxOrigins && nx <= kMaxOrigins, RISP_E_INVALID,
               "%s: origin lists must hold 1..%d entries", who, kMaxOrigins);
  o->ny = ny; o->nx = nx;
  for (int i = 0; i < ny; ++i) {
    RISP_REQUIRE(ys[i] >= 0 && ys[i] + h <= H, RISP_E_INVALID, "%s: tile row origin %d out of the frame", who, ys[i]);
    o->ys[i] = ys[i];
  }
  for (int i = 0; i < nx; ++i) {
    RISP_REQUIRE(xs[i] >= 0 && xs[i] + w <= W, RISP_E_INVALID, "%s: tile column origin %d out of the frame", who, xs[i]);
    o->xs[i] = xs[i];
  }
  return RISP_OK;
}

}  // namespace risp

using namespace risp;

extern "C" int risp_whole2patch(const float* frame, float* tiles, int C, int H, int W, int h, int w, const int* ys,
                                int ny, const int* xs, int nx, risp_stream_t stream) {
  RISP_REQUIRE(frame && tiles && C >= 1 && h >= 1 && w >= 1 && h <= H && w <= W, RISP_E_INVALID,
               "risp_whole2patch: bad arguments (the reference asserts sh<=h<=H and sw<=w<=W, util_path_restore.py:78)");
  Origins o;
  int rc = fill_origins(&o, "risp_whole2patch", ys, ny, xs, nx, H, W, h, w);
  if (rc != RISP_OK) return rc;
  long long total = (long long)ny * nx * C * h * w;
  long long g = cdiv(total, kT);
  long long cap = (long long)sm_count() * 16;
  whole2patch_kernel<<<(int)(g > cap ? cap : g), kT, 0, as_stream(stream)>>>(frame, tiles, C, H, W, h, w, o, total);
  return check_launch("whole2patch_kernel");
}

extern "C" int risp_patch2whole(const float* tiles, float* frame, int C, int H, int W, int h, int w, int sh, int sw,
                                const int* ys, int ny, const int* xs, int nx, int clip01, risp_stream_t stream) {
  RISP_REQUIRE(frame && tiles && C >= 1 && C <= 4 && h >= 1 && w >= 1 && h <= H && w <= W && sh >= 1 && sw >= 1 &&
                   sh <= h && sw <= w,
               RISP_E_INVALID, "risp_patch2whole: bad arguments");
  Origins o;
  int rc = fill_origins(&o, "risp_patch2whole", ys, ny, xs, nx, H, W, h, w);
  if (rc != RISP_OK) return rc;
  long long plane = (long long)H * W;
  long long g = cdiv(plane, kT);
  long long cap = (long long)sm_count() * 16;
  patch2whole_kernel<<<(int)(g > cap ? cap : g), kT, 0, as_stream(stream)>>>(tiles, frame, nullptr, C, H, W, h, w, (h - sh) / 2,
                                                                           (w - sw) / 2, o, clip01);
  return check_launch("patch2whole_kernel");
}

extern "C" int risp_patch2whole_u8(const float* tiles, float* frame, unsigned char* frame_u8, int C, int H, int W, int h, int w,
                                   int sh, int sw, const int* ys, int ny, const int* xs, int nx, risp_stream_t stream) {
  RISP_REQUIRE(frame_u8 && tiles && C >= 1 && C <= 4 && h >= 1 && w >= 1 && h <= H && w <= W && sh >= 1 && sw >= 1 &&
                   sh <= h && sw <= w,
               RISP_E_INVALID, "risp_patch2whole_u8: bad arguments");
  Origins o;
  int rc = fill_origins(&o, "risp_patch2whole_u8", ys, ny, xs, nx, H, W, h, w);
  if (rc != RISP_OK) return rc;
  long long plane = (long long)H * W;
  long long g = cdiv(plane, kT);
  long long cap = (long long)sm_count() * 16;
  patch2whole_kernel<<<(int)(g > cap ? cap : g), kT, 0, as_stream(stream)>>>(tiles, frame, frame_u8, C, H, W, h, w, (h - sh) / 2,
                                                                           (w - sw) / 2, o, 1);
  return check_launch("patch2whole_kernel");
}

namespace risp {
// tensor2bgr on the device (utils/util.py:118-135): (C,H,W) float -> (H,W,C) uint8, clip(x*255, 0, 255) truncated
__global__ void __launch_bounds__(kT)
to_u8_hwc_kernel(const float* __restrict__ x, unsigned char* __restrict__ out, int C, long long HW) {
  for (long long p = (long long)blockIdx.x * kT + threadIdx.x; p < HW; p += (long long)gridDim.x * kT) {
    for (int c = 0; c < C; ++c) {
      const float v = fminf(fmaxf(__fmul_rn(x[(long long)c * HW + p], 255.f), 0.f), 255.f);
      out[p * C + c] = (unsigned char)(int)v;
    }
  }
}
}  // namespace risp

extern "C" int risp_to_u8_hwc(const float* x, unsigned char* out, int C, int H, int W, risp_stream_t stream) {
  RISP_REQUIRE(x && out && C >= 1 && C <= 4 && H > 0 && W > 0, RISP_E_INVALID, "risp_to_u8_hwc: bad arguments");
  const long long HW = (long long)H * W;
  long long g = cdiv(HW, kT), cap = (long long)sm_count() * 16;
  to_u8_hwc_kernel<<<(int)(g > cap ? cap : g), kT, 0, as_stream(stream)>>>(x, out, C, HW);
  return check_launch("to_u8_hwc_kernel");
}
