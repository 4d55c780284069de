// Tensor-core convolution for the CNN candidates: implicit GEMM on tcgen05 (UMMA) with TMEM accumulators,
// fp32-accurate through a 3-term split  a*b ~ a_hi*b_hi + (a_lo*b + a*b_lo)  (a_hi, b_hi = the TF32 truncations; the
// dropped a_lo*b_lo term is ~2^-22 relative), so the 1e-4 parity bar of this path holds.  The main term is ONE
// kind::tf32 MMA (K = 8 channels).  The two cross terms are ~2^-11 of the main term, so 8 mantissa bits are enough for
// them: they are ONE kind::f16 MMA on bf16 operands whose K = 16 is [a_lo(4) | a(4)] x [b(4) ; b_lo(4)] per 4-channel
// group -- two instructions per 8 channels instead of three, and the bf16 tile has exactly the geometry (16 bytes per
// pixel and channel group) of the fp32 tile, so both use the same descriptors.  Error of the cross terms: 2^-9 * 2^-11.
//
// GEMM view: M = 128 consecutive pixels of one output row, N = output channels (16..64), K = taps x input
// channels.  Activations live in a channel-blocked layout [N][H][C/4][W][4] so that, for a group of 4 input
// channels, the pixels of a row are a contiguous run of 16-byte elements.  Staged into shared memory as
// [row][kgroup][pixel (tile + halo)][4] this IS the canonical no-swizzle K-major UMMA operand layout (8 x 16 B
// core matrices, SBO = 128 B between 8-pixel groups, LBO = the kgroup stride), and a horizontal tap shift is
// nothing but +16 B on the descriptor start address: no im2col, no data duplication across the K_w taps.
// A CTA owns R output rows x 128 columns x all output channels: the R accumulators sit in TMEM (R*N <= 512
// columns), every staged input row feeds up to K_h of them, and the weights of one tap row stream through
// shared memory once per CTA and input-channel chunk.
//
// One elected thread issues the MMAs; completion is tracked with tcgen05.commit -> mbarrier; the epilogue
// reads TMEM with tcgen05.ld (32x32b), applies bias / ReLU / residual and writes the blocked layout (and/or
// planar NCHW) with 128-bit stores.
#include <cuda_bf16.h>
#include "risp_common.cuh"

namespace risp {

constexpr int TC_M = 128;           // pixels per row tile
constexpr int TC_THREADS = 256;

struct ConvTcArgs {
  const float* x;        // blocked (N, H, CinG, W, 4)
  const float* wprep;    // [chunk][dy][dx][kg][CoutPad][4] hi, then the same again for lo
  const float* bias;     // nullable (Cout)
  const float* bias_tab; // nullable (N, K*K, CoutPad): bias by border class of the output pixel (SRCNNRes' folded constant channels)
  const float* res;      // nullable, blocked (N, H, CoutG, W, 4)
  const float* mask_in;  // nullable, blocked like x: x is multiplied by [mask_in > 0] while staging (backward of an output ReLU)
  const float* mask_out; // nullable, blocked like y: the result is multiplied by [mask_out > 0] (backward of an input ReLU)
  float* y_blk;          // nullable, blocked (N, H, CoutG, W, 4)
  float* y_pln;          // (unused by the kernel: a planar copy is a from_blocked pass after it)
  int CinG, Cout, CoutPad, H, W, flags;
  int Cin;               // real input channels: only ceil(Cin/8) chunks are multiplied (the rest of the layout is zero padding)
  int w_chunks;          // chunks the weights were prepared with (the lo block follows the hi block of ALL chunks)
  // grouped launch (a bank of networks with identical layer shapes, e.g. the eight SRCNNRes proxies of a supernet step):
  // blockIdx.y = g * n_per_group + n; group g uses weight / bias slot gslot[g]; x and res may be shared by all groups
  int n_per_group;
  long long w_group_floats;   // floats between the prepared weights of consecutive slots
  int bias_group_floats;      // floats between the bias vectors of consecutive slots
  int x_shared, res_shared;   // 1: the input / residual has n_per_group images, read by every group
  unsigned char gslot[8];
};

// ---- PTX wrappers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48)
  // | layout_type SWIZZLE_NONE=0 [61,64)
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// nearest TF32 value (10 explicit mantissa bits, low 13 bits zero): the residual a - tf32_rn(a) is at most 2^-12 |a|, half of
// what truncation leaves, which halves every error term of the bf16 cross-term MMA
__device__ __forceinline__ float tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo_elem, float hi_elem) {      // lo_elem at the lower address
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t mbar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
               : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must trap (and fail the launch) instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try(mbar, parity); ++spin)
    if (spin > (1u << 24)) __trap();
}
// one elected lane of a fully converged warp (ptxas keeps the MMA operands in uniform registers and emits a
// plain predicated UTCHMMA; under an ordinary `tid == 0` branch it wraps every MMA in an ELECT / BRA.U.ANY loop)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[N]);
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- kernel ---------------------------------------------------------------------------------------------
// K: filter size; CI_C: input channels per chunk (multiple of 8); R: output rows per CTA; NP: padded Cout (16/32/48/64);
// NBUF: weight-stage buffers; MINB: CTAs per SM the configuration is sized for; ABUF: input-tile buffers (1 or 2);
// PF: input chunks prefetched in registers (1 or 2); DXS: horizontal taps per weight stage (divides K)
//
// Fat MMAs (measured on B200, scripts/probe_mma_rate.py): a 128 x N x 8 tf32 MMA with both operands in shared memory
// costs max(N/2, (4 KB + N*32 B)/128 B per clk) + ~5 cycles, with a floor of ~46 -- N = 64 runs at 60 % of the tensor
// rate, N = 16 at 17 %.  So the vertical taps are stacked ON N: for one staged input row i and one horizontal tap dx,
// ONE MMA multiplies the row by [W(dy=K-1,dx) | ... | W(dy=0,dx)] (reversed so that consecutive column blocks belong
// to consecutive output rows i-K+1 .. i) and lands directly in the TMEM accumulators of all the output rows that
// input row feeds: N = min(rows fed, R) * NP <= 256, and the accumulators are laid out [row][NP] as the epilogue wants.
//
// Accuracy note (measured on B200): every tcgen05.mma rounds its fp32 accumulator once, toward zero, so the error
// of one accumulator grows linearly with the number of MMAs chained into it (~6e-8 relative each).  The two
// cross terms (a_lo*b_hi, a_hi*b_lo) are ~2^-11 of the main term, so they get their OWN accumulator: the main
// accumulator then sees one rounding per 8 input channels per tap instead of three, and the roundings of the
// small accumulator are 2^-11 times less significant.  The two accumulators are added in the epilogue.
//
// Pipeline of one CTA:
//   weights  : stages of one horizontal tap (all K vertical taps, one chunk), NBUF-deep ring filled by 1-D bulk async
//              copies (TMA engine, no registers, no thread work) that signal bfull[buf] by complete_tx; the issuing
//              lane refills a buffer as soon as the MMAs that read it have committed (mdone[buf])
//   inputs   : the rows of the NEXT input-channel chunk are loaded into registers while the MMAs of the current
//              chunk run, then split (hi/lo) + stored once the MMAs that read the target buffer have committed
//              (afree[buf]).  With ABUF = 2 that is the chunk before the current one, so the MMA stream never drains
//              at a chunk boundary; with ABUF = 1 the second resident CTA fills the gap.  With PF = 2 a second
//              register set keeps the loads TWO chunks ahead: a small-K chunk's MMAs are shorter than one L2 round trip
//   MMAs     : one elected lane of warp 0; accumulators are zeroed with tcgen05.st up front so every MMA accumulates
template <int K, int CI_C, int R, int NP, int NBUF, int MINB, int ABUF, int PF, int DXS>
struct TcCfg {
  static constexpr int PAD = K / 2;
  static constexpr int PW = TC_M + K - 1;                 // staged pixels per row
  static constexpr int KG = CI_C / 4;                     // 16-byte k-groups per chunk
  static constexpr int ROWS = R + K - 1;
  static constexpr int A_ROW_FLOATS = KG * PW * 4;        // one staged row, one precision
  static constexpr int A_FLOATS = ROWS * A_ROW_FLOATS;    // hi (lo follows)
  static constexpr int NSTACK = K * NP;                   // GEMM-N rows of one weight stage: [dy reversed][co]
  static constexpr int B_TAP_FLOATS = KG * NSTACK * 4;    // one horizontal tap, one precision
  static constexpr int B_STAGE_FLOATS = DXS * B_TAP_FLOATS;
  static constexpr int B_CHUNK_FLOATS = K * B_TAP_FLOATS;
  static constexpr int NST = K / DXS;                     // weight stages per chunk
  static_assert(K % DXS == 0, "taps per stage must divide K");
  static constexpr size_t SMEM = sizeof(float) * (2 * ABUF * A_FLOATS + 2 * NBUF * B_STAGE_FLOATS) + 128 + 256;   // + barriers + bias
  static constexpr int ACC_COLS = 2 * R * NP;             // main + cross-term accumulators
  static constexpr int TMEM_COLS = (ACC_COLS <= 32) ? 32 : (ACC_COLS <= 64) ? 64 : (ACC_COLS <= 128) ? 128 : (ACC_COLS <= 256) ? 256 : 512;
  static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
  static_assert(TMEM_COLS * MINB <= 512, "TMEM oversubscribed");
  static_assert((R < K ? R : K) * NP <= 256, "MMA N exceeds 256");
  static_assert(NBUF >= 2 && NBUF <= 6 && (ABUF == 1 || ABUF == 2), "ring depths");
  static constexpr int A_TOTAL = ROWS * KG * PW;          // float4 elements staged per chunk
  static constexpr int STAGERS = TC_THREADS - 32;         // warp 0 only issues MMAs / weight fetches; warps 1..7 stage inputs
  static constexpr int A_ITER = (A_TOTAL + STAGERS - 1) / STAGERS;
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
// zero 16 consecutive TMEM columns of this warp's 32 lanes
__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
               ::"r"(taddr), "r"(0u) : "memory");
}

// optional in-kernel timeline of one CTA (build with -DRISP_TC_TRACE; read back with risp_debug_tc_trace)
__device__ long long g_tc_trace[16];
#ifdef RISP_TC_TRACE
#define TC_TRACE(slot) do { if (trace_cta && tid == 0) g_tc_trace[slot] = clock64(); } while (0)
#define TC_TRACE_ADD(slot, t_begin) do { tacc[slot - 8] += clock64() - (t_begin); } while (0)
#else
#define TC_TRACE(slot) do { } while (0)
#define TC_TRACE_ADD(slot, t_begin) do { } while (0)
#endif

template <int K, int CI_C, int R, int NP, int NBUF, int MINB, int ABUF, int PF, int DXS>
__global__ void __launch_bounds__(TC_THREADS, MINB)
conv_tc_kernel(ConvTcArgs a) {
  using C = TcCfg<K, CI_C, R, NP, NBUF, MINB, ABUF, PF, DXS>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* sA = reinterpret_cast<float*>(smem_raw);          // [abuf][hi|lo][A_FLOATS]
  float* sB = sA + 2 * ABUF * C::A_FLOATS;                 // [buf][hi|lo][B_STAGE_FLOATS]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sB + 2 * NBUF * C::B_STAGE_FLOATS);   // bfull[NBUF], mdone[NBUF], afree[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2 * NBUF + 2);
  float* s_bias = reinterpret_cast<float*>(mbar + 16);     // 64 floats, zero beyond Cout / without a bias
  const uint32_t bfull = smem_u32(mbar), mdone = smem_u32(mbar + NBUF), afree = smem_u32(mbar + 2 * NBUF);   // afree[2]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int strips = (a.W + TC_M - 1) / TC_M;
  const int x0 = (blockIdx.x % strips) * TC_M;
  const int y0 = (blockIdx.x / strips) * R;
  const int n = blockIdx.y;                                  // image of the (grouped) output, mask and bias-table tensors
  const int grp = n / a.n_per_group, nl = n - grp * a.n_per_group;
  const int slot = a.gslot[grp & 7];
  const int n_x = a.x_shared ? nl : n, n_res = a.res_shared ? nl : n;
  const float* __restrict__ wprep = a.wprep + (long long)slot * a.w_group_floats;
  const bool relu_in = (a.flags & RISP_CONV_RELU_IN) != 0;
#ifdef RISP_TC_TRACE
  const bool trace_cta = blockIdx.x == 5 && blockIdx.y == 0;
  if (trace_cta && tid == 0) { for (int i = 0; i < 16; ++i) g_tc_trace[i] = 0; }
  long long tq = 0; (void)tq;
  long long tacc[6] = {0, 0, 0, 0, 0, 0};      // per-thread accumulators (registers): slots 8..13
#endif
  TC_TRACE(0);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 2 * NBUF + 2; ++i) mbar_init(smem_u32(mbar + i), 1);
  }
  if (tid >= 64 && tid < 128) s_bias[tid - 64] = (a.bias && tid - 64 < a.Cout) ? __ldg(a.bias + slot * a.bias_group_floats + tid - 64) : 0.f;
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_chunks = (a.Cin + CI_C - 1) / CI_C;
  const int n_stages = n_chunks * C::NST;
  const long long row_stride = (long long)a.CinG * a.W * 4;              // floats per image row (blocked layout)
  const float* xin = a.x + (long long)n_x * a.H * row_stride;
  const float* min_ = a.mask_in ? a.mask_in + (long long)n * a.H * row_stride : nullptr;
  const long long lo_off = (long long)a.w_chunks * C::B_CHUNK_FLOATS;
  constexpr uint32_t kStageBytes = C::B_STAGE_FLOATS * 4;

  // weight stage s -> buffer s % NBUF (hi block then lo block); issued by the MMA lane only
  auto fetch_weights = [&](int s) {
    const long long off = (long long)s * C::B_STAGE_FLOATS;              // stages are contiguous: [chunk][dx]
    const uint32_t b = (uint32_t)(s % NBUF);
    const uint32_t bar = bfull + 8u * b;
    const uint32_t dst = smem_u32(sB) + b * 2u * kStageBytes;
    mbar_expect_tx(bar, 2u * kStageBytes);
    bulk_g2s(dst, wprep + off, kStageBytes, bar);
    bulk_g2s(dst + kStageBytes, wprep + lo_off + off, kStageBytes, bar);
  };

  // input rows of chunk c -> registers (zero padding; the optional mask is applied here so only one array stays live).
  // The unmasked path must not contain any use of the loaded values: a (predicated-off) select right behind a load
  // still waits for it and would serialise the L2 latency of every element.
  float4 va[C::A_ITER], vb[PF == 2 ? C::A_ITER : 1];
  auto input_offset = [&](int it, int c, long long& o) -> bool {
    const int i = (tid - 32) + it * C::STAGERS;
    const int px = i % C::PW;
    const int kg = (i / C::PW) % C::KG;
    const int row = i / (C::PW * C::KG);
    const int gy = y0 - C::PAD + row, gx = x0 - C::PAD + px, gkg = c * C::KG + kg;
    o = (long long)gy * row_stride + ((long long)gkg * a.W + gx) * 4;
    return i < C::A_TOTAL && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W && gkg < a.CinG;
  };
  auto load_set = [&](int c, float4* v) {
    if (warp == 0) return;
    if (!min_) {
#pragma unroll
      for (int it = 0; it < C::A_ITER; ++it) {
        long long o;
        const bool ok = input_offset(it, c, o);
        v[it] = ok ? __ldg(reinterpret_cast<const float4*>(xin + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      // masked (data-gradient) path: groups of four elements, eight loads in flight per group
#pragma unroll
      for (int g = 0; g < C::A_ITER; g += 4) {
        float4 m[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (g + u < C::A_ITER) {
            long long o;
            const bool ok = input_offset(g + u, c, o);
            v[g + u] = ok ? __ldg(reinterpret_cast<const float4*>(xin + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
            m[u] = ok ? __ldg(reinterpret_cast<const float4*>(min_ + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (g + u < C::A_ITER) {
            float4& t = v[g + u];
            t.x = m[u].x > 0.f ? t.x : 0.f; t.y = m[u].y > 0.f ? t.y : 0.f; t.z = m[u].z > 0.f ? t.z : 0.f; t.w = m[u].w > 0.f ? t.w : 0.f;
          }
        }
      }
    }
  };
  auto store_set = [&](int c, const float4* v) {
    if (warp == 0) return;
    float* sA_hi = sA + (size_t)(c % ABUF) * 2 * C::A_FLOATS;
    float* sA_lo = sA_hi + C::A_FLOATS;
#pragma unroll
    for (int it = 0; it < C::A_ITER; ++it) {
      const int i = (tid - 32) + it * C::STAGERS;
      float4 t = v[it];
      if (relu_in) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
      float4 hi, lo;
      hi.x = tf32_rn(t.x); lo.x = t.x - hi.x;
      hi.y = tf32_rn(t.y); lo.y = t.y - hi.y;
      hi.z = tf32_rn(t.z); lo.z = t.z - hi.z;
      hi.w = tf32_rn(t.w); lo.w = t.w - hi.w;
      if (i < C::A_TOTAL) {
        reinterpret_cast<float4*>(sA_hi)[i] = hi;
        // cross-term operand, 8 bf16 per pixel and channel group: [a_lo(4) | a(4)]
        reinterpret_cast<uint4*>(sA_lo)[i] = make_uint4(pack_bf16(lo.x, lo.y), pack_bf16(lo.z, lo.w), pack_bf16(t.x, t.y), pack_bf16(t.z, t.w));
      }
    }
  };

  // register set of chunk c: alternating with PF = 2
  auto load_inputs = [&](int c) { if (PF == 2 && (c & 1)) load_set(c, vb); else load_set(c, va); };
  auto store_inputs = [&](int c) { if (PF == 2 && (c & 1)) store_set(c, vb); else store_set(c, va); };

  // ---- prologue: first weight stages in flight, accumulators zeroed, inputs of chunk 0 staged ----
  if (warp == 0) {
    if (elect_one()) {
      for (int s = 0; s < NBUF - 1 && s < n_stages; ++s) fetch_weights(s);
    }
    __syncwarp();
  }
  TC_TRACE(5);
  load_inputs(0);
  if (PF == 2 && n_chunks > 1) load_inputs(1);
  TC_TRACE(6);
  {
    // warp w owns TMEM lanes 32*(w%4)..; the two warp sets split the 16-column groups
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll
    for (int g = 0; g < C::ACC_COLS / 16; ++g)
      if ((g & 1) == (warp >> 2)) tmem_zero16(lane_base + (uint32_t)(g * 16));
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  TC_TRACE(7);
  store_inputs(0);
  TC_TRACE(14);
  fence_async_smem();        // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();
  TC_TRACE(1);

  for (int s = 0; s < n_stages; ++s) {
    const int c = s / C::NST, dx = s % C::NST;      // dx: stage within the chunk (DXS horizontal taps each)
    if (dx == 0 && s > 0) {
      // chunk boundary
#ifdef RISP_TC_TRACE
      tq = clock64();
#endif
      // the previous reader of this chunk's input buffer is chunk c - ABUF
      if (c >= ABUF) mbar_wait(afree + 8u * (uint32_t)(c % ABUF), (uint32_t)(((c - ABUF) / ABUF) & 1));
      TC_TRACE_ADD(8, tq);
#ifdef RISP_TC_TRACE
      tq = clock64();
#endif
      store_inputs(c);         // registers were loaded while the MMAs ran
      TC_TRACE_ADD(12, tq);
#ifdef RISP_TC_TRACE
      tq = clock64();
#endif
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      TC_TRACE_ADD(13, tq);
    }
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t b = (uint32_t)(s % NBUF);
#ifdef RISP_TC_TRACE
        tq = clock64();
#endif
        mbar_wait(bfull + 8u * b, (uint32_t)((s / NBUF) & 1));
        TC_TRACE_ADD(9, tq);
#ifdef RISP_TC_TRACE
        tq = clock64();
#endif
        tc_fence_after();
        const uint32_t sBh = smem_u32(sB) + b * 2u * kStageBytes, sBl = sBh + kStageBytes;
        const uint32_t sAh = smem_u32(sA) + (uint32_t)(c % ABUF) * (uint32_t)(2 * C::A_FLOATS * 4), sAl = sAh + (uint32_t)(C::A_FLOATS * 4);
        // One descriptor per operand tile and stage; every MMA's descriptor is that base plus a COMPILE-TIME offset in the
        // start-address field (16-byte units; shared memory < 256 KB, so the 14-bit field never carries).  Building each
        // descriptor from its address cost ~30 uniform-datapath instructions per MMA -- as long as the MMA itself runs.
        const uint64_t dAh0 = make_desc(sAh + (uint32_t)(dx * DXS * 16), C::PW * 16, 128);
        const uint64_t dAl0 = make_desc(sAl + (uint32_t)(dx * DXS * 16), C::PW * 16, 128);
        const uint64_t dBh0 = make_desc(sBh, C::NSTACK * 16, 128);
        const uint64_t dBl0 = make_desc(sBl, C::NSTACK * 16, 128);
#pragma unroll
        for (int dxl = 0; dxl < DXS; ++dxl) {
        const uint32_t a_dx = (uint32_t)(dxl * 16);
        const uint32_t b_tap = (uint32_t)(dxl * C::B_TAP_FLOATS * 4);
#pragma unroll
        for (int i = 0; i < C::ROWS; ++i) {
          // input row i feeds output rows r_lo..r_hi; their weights are the column blocks j0.. of the tap
          const int r_lo = (i - K + 1 > 0) ? i - K + 1 : 0;
          const int r_hi = (i < R - 1) ? i : R - 1;
          const int nrows = r_hi - r_lo + 1;
          const int j0 = r_lo - (i - K + 1);
          const uint32_t shape = ((uint32_t)((nrows * NP) >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
          const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | shape;       // D fp32, A/B tf32, K-major
          const uint32_t idesc_x = (1u << 4) | (1u << 7) | (1u << 10) | shape;     // D fp32, A/B bf16, K-major
          const uint32_t d_main = tmem_base + (uint32_t)(r_lo * NP);
          const uint32_t d_cross = tmem_base + (uint32_t)((R + r_lo) * NP);
#pragma unroll
          for (int ks = 0; ks < C::KG / 2; ++ks) {
            const uint32_t a_off = (uint32_t)(i * C::A_ROW_FLOATS * 4 + (2 * ks) * C::PW * 16) + a_dx;
            const uint32_t b_off = b_tap + (uint32_t)((2 * ks) * C::NSTACK * 16 + j0 * NP * 16);
            const uint64_t a16 = (uint64_t)(a_off >> 4), b16 = (uint64_t)(b_off >> 4);
            umma_tf32(d_main, dAh0 + a16, dBh0 + b16, idesc, 1u);
            umma_bf16(d_cross, dAl0 + a16, dBl0 + b16, idesc_x, 1u);          // a_lo*b + a*b_lo in one K = 16 instruction
          }
        }
        }
        umma_commit(mdone + 8u * b);
        if (dx == C::NST - 1) umma_commit(afree + 8u * (uint32_t)(c % ABUF));   // one phase per chunk and input buffer
        TC_TRACE_ADD(10, tq);
#ifdef RISP_TC_TRACE
        tq = clock64();
#endif
        const int sn = s + NBUF - 1;                          // refill the buffer stage s-1 used
        if (sn < n_stages) {
          if (s >= 1) mbar_wait(mdone + 8u * (uint32_t)(sn % NBUF), (uint32_t)(((s - 1) / NBUF) & 1));
          fetch_weights(sn);
        }
        TC_TRACE_ADD(11, tq);
      }
      __syncwarp();
    }
    if (dx == 0 && c + PF < n_chunks) load_inputs(c + PF);   // rows PF chunks ahead: in flight while the MMAs run
  }
  // ---- epilogue: 8 warps; warp w reads TMEM lanes 32*(w%4).., and the (w/4)-th half of the 16-column groups ----
  TC_TRACE(2);
#ifdef RISP_TC_TRACE
  if (trace_cta && warp == 0 && tacc[2] != 0) { g_tc_trace[9] = tacc[1]; g_tc_trace[10] = tacc[2]; g_tc_trace[11] = tacc[3]; }
  if (trace_cta && tid == 32) { g_tc_trace[8] = tacc[0]; g_tc_trace[12] = tacc[4]; g_tc_trace[13] = tacc[5]; }
#endif
  mbar_wait(afree + 8u * (uint32_t)((n_chunks - 1) % ABUF), (uint32_t)(((n_chunks - 1) / ABUF) & 1));
  tc_fence_after();
  TC_TRACE(3);
#ifdef RISP_TC_TRACE
  tacc[0] = 0;
#endif
  // With 8 warps per CTA (2 per scheduler) this straight-line code is bound by dependent-instruction latency, so it is kept
  // short (~30 instructions per 16-byte store; a first version needed ~140 and took as long as the MMAs of a 9x9 layer):
  // row base pointers are hoisted, offsets inside a row are 32-bit, the ReLU / residual-ReLU switches are clamps against
  // 0 or -inf instead of branches, and the bias / residual / mask values of a 16-channel group are all requested before the
  // TMEM loads are waited for.
  const bool relu_out = (a.flags & RISP_CONV_RELU_OUT) != 0, add_res = (a.flags & RISP_CONV_ADD_RES) != 0,
             res_relu = (a.flags & RISP_CONV_RES_RELU) != 0;
  const float out_floor = relu_out ? 0.f : -INFINITY, res_floor = res_relu ? 0.f : -INFINITY;
  const int lq = warp & 3, half = warp >> 2;
  const int gx = x0 + lq * 32 + lane;
  const int cout4 = (a.Cout + 3) & ~3;                  // channels of the blocked output / residual / mask (layout: ceil(C/4) groups)
  const int gstride = a.W * 4;                          // floats between channel groups of one row
  const int orow = (cout4 >> 2) * gstride;              // floats per row
  // border class of the output pixel (bias table): taps that fall outside the frame see zero padding, so the folded
  // contribution of a spatially constant channel depends on which of the K x K border classes the pixel is in
  const int gxc = gx < a.W ? gx : a.W - 1;              // lanes beyond the frame load (and discard) the last pixel's bias
  const int xcls = gxc < C::PAD ? gxc : (gxc >= a.W - C::PAD ? K - (a.W - gxc) : C::PAD);
  const long long img_base = (long long)n * a.H * orow + (long long)gx * 4;
  const uint32_t t_lane = tmem_base + ((uint32_t)(lq * 32) << 16);
  const bool has_mask = a.mask_out != nullptr;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int gy = y0 + r;                              // warp-uniform
    const bool ok = gy < a.H && gx < a.W;
    const long long rb = img_base + (long long)gy * orow;
    float* yrow = a.y_blk + rb;
    const float* rrow = a.res + rb + (long long)(n_res - n) * a.H * orow;      // only dereferenced when the pointer is set
    const float* mrow = a.mask_out + rb;
    const int gyc = gy < a.H ? gy : a.H - 1;
    const int ycls = gyc < C::PAD ? gyc : (gyc >= a.H - C::PAD ? K - (a.H - gyc) : C::PAD);
    // generic pointer: the per-class row of the bias table (global) or the staged per-channel bias (shared)
    const float* bsrc = a.bias_tab ? a.bias_tab + ((long long)n * (K * K) + ycls * K + xcls) * a.CoutPad : s_bias;
#pragma unroll
    for (int cb = 0; cb < NP / 16; ++cb) {
      if ((cb & 1) != half && NP > 16) continue;        // split the column groups between the two warp sets
      if (NP == 16 && half != 0) continue;
      float v[16], vc[16];
#ifdef RISP_TC_TRACE
      tq = clock64();
#endif
      tmem_ld<16>(t_lane + (uint32_t)(r * NP + cb * 16), v);
      tmem_ld<16>(t_lane + (uint32_t)((R + r) * NP + cb * 16), vc);
      float4 bb[4], rr[4], mm[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int co = cb * 16 + q * 4;
        const bool live = ok && co < cout4;
        bb[q] = *reinterpret_cast<const float4*>(bsrc + co);
        rr[q] = (live && add_res) ? __ldg(reinterpret_cast<const float4*>(rrow + (cb * 4 + q) * gstride)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_mask) mm[q] = live ? __ldg(reinterpret_cast<const float4*>(mrow + (cb * 4 + q) * gstride)) : make_float4(1.f, 1.f, 1.f, 1.f);
      }
      tmem_ld_wait();
      TC_TRACE_ADD(8, tq);          // slot 8 is re-used in the epilogue: cycles in TMEM loads (thread 0)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int co = cb * 16 + q * 4;
        float4 o;
        o.x = fmaxf(v[q * 4] + vc[q * 4] + bb[q].x, out_floor) + fmaxf(rr[q].x, res_floor);
        o.y = fmaxf(v[q * 4 + 1] + vc[q * 4 + 1] + bb[q].y, out_floor) + fmaxf(rr[q].y, res_floor);
        o.z = fmaxf(v[q * 4 + 2] + vc[q * 4 + 2] + bb[q].z, out_floor) + fmaxf(rr[q].z, res_floor);
        o.w = fmaxf(v[q * 4 + 3] + vc[q * 4 + 3] + bb[q].w, out_floor) + fmaxf(rr[q].w, res_floor);
        if (has_mask) {
          o.x = mm[q].x > 0.f ? o.x : 0.f; o.y = mm[q].y > 0.f ? o.y : 0.f; o.z = mm[q].z > 0.f ? o.z : 0.f; o.w = mm[q].w > 0.f ? o.w : 0.f;
        }
        if (ok && co < cout4) *reinterpret_cast<float4*>(yrow + (cb * 4 + q) * gstride) = o;
      }
    }
  }
#ifdef RISP_TC_TRACE
  if (trace_cta && tid == 0) { g_tc_trace[15] = clock64(); g_tc_trace[7] = tacc[0]; }     // own epilogue work done; [7] = boundary waits + TMEM-load cycles
#endif
  tc_fence_before();
  __syncthreads();
  TC_TRACE(4);
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

// weights (Cout,Cin,K,K) -> [chunk][dx][kg][j = K-1-dy][NP][4] hi block (tf32 truncations), then the cross-term block of the
// same geometry (16 bytes per element, bf16).
// transpose_flip: data-gradient operator.
__global__ void conv_tc_prepare_kernel(const float* __restrict__ w, float* __restrict__ out, int Cin, int Cout, int K, int CI_C,
                                       int NP, int transpose_flip, long long total) {
  const int rows = transpose_flip ? Cout : Cin, cols = transpose_flip ? Cin : Cout;   // rows = GEMM-K channels, cols = GEMM-N
  const int KG = CI_C / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int e = (int)(t % 4); t /= 4;
    const int co = (int)(t % NP); t /= NP;
    const int jr = (int)(t % K); t /= K;
    const int kg = (int)(t % KG); t /= KG;
    const int dx = (int)(t % K); t /= K;
    const int chunk = (int)t;
    const int dy = K - 1 - jr;
    const int ci = chunk * CI_C + kg * 4 + e;
    float v = 0.f;
    if (ci < rows && co < cols) {
      if (!transpose_flip) v = w[(((long long)co * Cin + ci) * K + dy) * K + dx];
      else v = w[(((long long)ci * Cin + co) * K + (K - 1 - dy)) * K + (K - 1 - dx)];
    }
    const float hi = tf32_rn(v);
    out[i] = hi;
    // cross-term operand: per (k-group, n) 8 bf16 = [b(4) ; b_lo(4)], pairing with the activations' [a_lo(4) | a(4)]
    __nv_bfloat16* xo = reinterpret_cast<__nv_bfloat16*>(out + total) + (i / 4) * 8;
    xo[e] = __float2bfloat16_rn(v);
    xo[4 + e] = __float2bfloat16_rn(v - hi);
  }
}

// planar (N,C,H,W) <-> blocked (N,H,CG,W,4), channels zero-padded to 4*CG
__global__ void to_blocked_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int CG, int H, int W, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    long long t = i / W;
    const int g = (int)(t % CG); t /= CG;
    const int y = (int)(t % H);
    const long long n = t / H;
    const long long plane = (long long)H * W;
    float4 o;
    const float* p = src + (n * C + 4 * g) * plane + (long long)y * W + x;
    o.x = (4 * g < C) ? p[0] : 0.f; o.y = (4 * g + 1 < C) ? p[plane] : 0.f;
    o.z = (4 * g + 2 < C) ? p[2 * plane] : 0.f; o.w = (4 * g + 3 < C) ? p[3 * plane] : 0.f;
    reinterpret_cast<float4*>(dst)[i] = o;
  }
}
__global__ void from_blocked_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int CG, int H, int W, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    long long t = i / W;
    const int g = (int)(t % CG); t /= CG;
    const int y = (int)(t % H);
    const long long n = t / H;
    const long long plane = (long long)H * W;
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    float* p = dst + (n * C + 4 * g) * plane + (long long)y * W + x;
    if (4 * g < C) p[0] = v.x;
    if (4 * g + 1 < C) p[plane] = v.y;
    if (4 * g + 2 < C) p[2 * plane] = v.z;
    if (4 * g + 3 < C) p[3 * plane] = v.w;
  }
}

// ---- gradient of the bias table: sums of a blocked gradient over the K x K border classes of its pixels -----------------
// Pass 1: one CTA per (row, image): for every channel group the interior columns are summed by a warp, the 2*PAD border
// columns are copied -> partial (N, H, K, CG*4).  Pass 2: rows folded into their K classes -> out (N, K, K, CP) with
// CP >= CG*4 (padding channels zero).  Fixed summation order (bit-reproducible).
__global__ void __launch_bounds__(256)
class_sums_rows_kernel(const float* __restrict__ g, const float* __restrict__ mask, float* __restrict__ partial, int CG, int H, int W,
                       int K) {
  const int y = blockIdx.x, n = blockIdx.y, PAD = K / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = ((long long)n * H + y) * CG * W;                // float4 elements
  const float4* gr = reinterpret_cast<const float4*>(g) + row;
  const float4* mr = mask ? reinterpret_cast<const float4*>(mask) + row : nullptr;
  float4* out = reinterpret_cast<float4*>(partial) + ((long long)n * H + y) * K * CG;
  auto load = [&](int cg, int x) {
    float4 v = __ldg(gr + (long long)cg * W + x);
    if (mr) {
      const float4 m = __ldg(mr + (long long)cg * W + x);
      v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
    }
    return v;
  };
  for (int cg = warp; cg < CG; cg += 8) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int x = PAD + lane;
    for (; x + 96 < W - PAD; x += 128) {                                // four independent loads (eight with a mask) in flight
      const float4 v0 = load(cg, x), v1 = load(cg, x + 32), v2 = load(cg, x + 64), v3 = load(cg, x + 96);
      acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
      acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; x < W - PAD; x += 32) {
      const float4 v = load(cg, x);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z); acc.w = warp_sum(acc.w);
    if (lane == 0) out[PAD * CG + cg] = acc;
    if (lane < 2 * PAD) {
      const int xb = lane < PAD ? lane : W - 2 * PAD + lane;              // classes 0..PAD-1 and PAD+1..K-1
      const int cls = lane < PAD ? lane : lane + 1;
      out[cls * CG + cg] = load(cg, xb);
    }
  }
}

// grid (K x-classes, N); 256 threads = C4 channels x (256 / C4) row lanes
__global__ void __launch_bounds__(256)
class_sums_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int CG, int H, int K, int CP) {
  __shared__ float red[256];
  const int xcls = blockIdx.x, n = blockIdx.y, PAD = K / 2, C4 = CG * 4;
  const int nlan = 256 / C4, c = threadIdx.x % C4, rl = threadIdx.x / C4;
  const float* base = partial + ((long long)n * H * K + xcls) * C4 + c;
  const long long ystride = (long long)K * C4;
  float acc = 0.f;
  if (rl < nlan) {
#pragma unroll 8
    for (int y = PAD + rl; y < H - PAD; y += nlan) acc += base[y * ystride];
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (rl == 0) {
    for (int j = 1; j < nlan; ++j) acc += red[j * C4 + c];
    out[(((long long)n * K + PAD) * K + xcls) * CP + c] = acc;
  }
  if (rl < nlan) {
    for (int j = rl; j < 2 * PAD; j += nlan) {                          // the single-row classes
      const int ycls = j < PAD ? j : j + 1, y = j < PAD ? j : H - 2 * PAD + j;
      out[(((long long)n * K + ycls) * K + xcls) * CP + c] = base[y * ystride];
    }
  }
  // zero the padding channels
  for (int i = threadIdx.x; i < K * (CP - C4); i += 256) {
    const int ycls = i / (CP - C4), cc = C4 + i % (CP - C4);
    out[(((long long)n * K + ycls) * K + xcls) * CP + cc] = 0.f;
  }
}

template <int K, int CI_C, int R, int NP, int NBUF, int MINB, int ABUF, int PF, int DXS>
static int launch_tc(const ConvTcArgs& a, int N, cudaStream_t st) {
  using C = TcCfg<K, CI_C, R, NP, NBUF, MINB, ABUF, PF, DXS>;
  static_assert(C::SMEM <= 227 * 1024 / MINB - 1024 * (MINB > 1), "shared memory budget");
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(conv_tc_kernel<K, CI_C, R, NP, NBUF, MINB, ABUF, PF, DXS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) != cudaSuccess) {
      set_error("conv_tc: cannot opt in to %zu bytes of shared memory", C::SMEM);
      return RISP_E_CUDA;
    }
    attr = true;
  }
  dim3 grid((unsigned)(cdiv(a.W, TC_M) * cdiv(a.H, R)), (unsigned)N);
  conv_tc_kernel<K, CI_C, R, NP, NBUF, MINB, ABUF, PF, DXS><<<grid, TC_THREADS, C::SMEM, st>>>(a);
  return check_launch("conv_tc_kernel");
}

static int tc_chunk(int K) { (void)K; return 8; }
static int g_tc_variant = 0;        // tuning knob (risp_debug_tc_variant): alternative tile configurations of the same kernel
}  // namespace risp
static int conv_tc_dispatch(const risp::ConvTcArgs& a, int N, int K, int NP, cudaStream_t st);
namespace risp {

}  // namespace risp

using namespace risp;

extern "C" int risp_conv_tc_supported(int Cin, int Cout, int K) {
  return (K == 1 || K == 3 || K == 5 || K == 9) && Cout <= 64 && Cin >= 1;
}

extern "C" int risp_conv_tc_padded_channels(int C) { return (C + 3) / 4 * 4; }

extern "C" size_t risp_conv_tc_weight_floats(int Cin, int Cout, int K, int transpose_flip) {
  const int rows = transpose_flip ? Cout : Cin, cols = transpose_flip ? Cin : Cout;
  const int CI_C = tc_chunk(K), NP = (cols + 15) / 16 * 16;
  const int chunks = (((rows + 15) / 16 * 16) + CI_C - 1) / CI_C;
  return (size_t)2 * chunks * K * K * (CI_C / 4) * NP * 4;
}

extern "C" int risp_conv_tc_prepare_weights(const float* weight, float* out, int Cin, int Cout, int K, int transpose_flip,
                                            risp_stream_t stream) {
  RISP_REQUIRE(weight && out && risp_conv_tc_supported(transpose_flip ? Cout : Cin, transpose_flip ? Cin : Cout, K), RISP_E_INVALID,
               "risp_conv_tc_prepare_weights: unsupported shape");
  const long long total = (long long)risp_conv_tc_weight_floats(Cin, Cout, K, transpose_flip) / 2;
  const int cols = transpose_flip ? Cin : Cout;
  conv_tc_prepare_kernel<<<(int)cdiv(total, 256), 256, 0, as_stream(stream)>>>(weight, out, Cin, Cout, K, tc_chunk(K),
                                                                             (cols + 15) / 16 * 16, transpose_flip, total);
  return check_launch("conv_tc_prepare_kernel");
}

extern "C" int risp_to_blocked(const float* planar, float* blocked, int N, int C, int CG, int H, int W, risp_stream_t stream) {
  RISP_REQUIRE(planar && blocked && N > 0 && C > 0 && CG * 4 >= C && H > 0 && W > 0, RISP_E_INVALID, "risp_to_blocked: bad arguments");
  const long long total = (long long)N * H * CG * W;
  to_blocked_kernel<<<(int)(cdiv(total, 256) > 148 * 32 ? 148 * 32 : cdiv(total, 256)), 256, 0, as_stream(stream)>>>(planar, blocked, C, CG, H, W, total);
  return check_launch("to_blocked_kernel");
}

extern "C" int risp_from_blocked(const float* blocked, float* planar, int N, int C, int CG, int H, int W, risp_stream_t stream) {
  RISP_REQUIRE(planar && blocked && N > 0 && C > 0 && CG * 4 >= C && H > 0 && W > 0, RISP_E_INVALID, "risp_from_blocked: bad arguments");
  const long long total = (long long)N * H * CG * W;
  from_blocked_kernel<<<(int)(cdiv(total, 256) > 148 * 32 ? 148 * 32 : cdiv(total, 256)), 256, 0, as_stream(stream)>>>(blocked, planar, C, CG, H, W, total);
  return check_launch("from_blocked_kernel");
}

// ---- the bias table itself: tab (N, J) = b (J) + feat (N, F) x S (F, J), and d feat = d tab x S^T -----------------------
// F = 9 + P <= 32 features, J = K*K*CoutPad (5184 for SRCNNRes); S is constant.  Two trivial kernels instead of cuBLAS's
// gemv (which needs ~85 us for this shape) / an elementwise product + row sums.
namespace risp {
struct GroupSlots { unsigned char s[8]; };
__global__ void __launch_bounds__(256)
bias_table_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ S, const float* __restrict__ b, float* __restrict__ tab,
                      int F, int J, int n_per_group, GroupSlots gs) {
  const int n = blockIdx.y;
  const int slot = gs.s[(n / n_per_group) & 7];
  S += (long long)slot * F * J;
  b += (long long)slot * J;
  __shared__ float f_s[32];
  if (threadIdx.x < F) f_s[threadIdx.x] = feat[n * F + threadIdx.x];
  __syncthreads();
  for (int j = blockIdx.x * 256 + threadIdx.x; j < J; j += gridDim.x * 256) {
    float acc = b[j];
    for (int f = 0; f < F; ++f) acc = fmaf(f_s[f], __ldg(S + (long long)f * J + j), acc);
    tab[(long long)n * J + j] = acc;
  }
}
__global__ void __launch_bounds__(256)
bias_table_bwd_kernel(const float* __restrict__ dtab, const float* __restrict__ S, float* __restrict__ dfeat, int F, int J,
                      int n_per_group, GroupSlots gs) {
  const int f = blockIdx.x, n = blockIdx.y;                  // one CTA per output element, fixed summation order
  S += (long long)gs.s[(n / n_per_group) & 7] * F * J;
  float acc = 0.f;
  for (int j = threadIdx.x; j < J; j += 256) acc = fmaf(dtab[(long long)n * J + j], __ldg(S + (long long)f * J + j), acc);
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    dfeat[n * F + f] = t;
  }
}
}  // namespace risp

static int make_group_slots(risp::GroupSlots* gs, int G, const int* slots, const char* who) {
  RISP_REQUIRE(G >= 1 && G <= 8 && slots, RISP_E_INVALID, "%s: 1..8 groups", who);
  for (int g = 0; g < 8; ++g) gs->s[g] = 0;
  for (int g = 0; g < G; ++g) {
    RISP_REQUIRE(slots[g] >= 0 && slots[g] < 256, RISP_E_INVALID, "%s: slot out of range", who);
    gs->s[g] = (unsigned char)slots[g];
  }
  return RISP_OK;
}

// grouped: feat / tab / dtab / dfeat have G*N rows; group g uses S + slots[g]*F*J and b + slots[g]*J
extern "C" int risp_bias_table_fwd_grouped(const float* feat, const float* S, const float* b, float* tab, int G, int N_per_group,
                                           const int* slots, int F, int J, risp_stream_t stream) {
  risp::GroupSlots gs;
  int rc = make_group_slots(&gs, G, slots, "risp_bias_table_fwd");
  if (rc != RISP_OK) return rc;
  const int N = G * N_per_group;
  RISP_REQUIRE(feat && S && b && tab && N_per_group > 0 && N <= 65535 && F >= 1 && F <= 32 && J >= 1, RISP_E_INVALID, "risp_bias_table_fwd: bad arguments");
  const int bx = (int)(cdiv(J, 256) > 64 ? 64 : cdiv(J, 256));
  risp::bias_table_fwd_kernel<<<dim3((unsigned)bx, (unsigned)N), 256, 0, as_stream(stream)>>>(feat, S, b, tab, F, J, N_per_group, gs);
  return check_launch("bias_table_fwd_kernel");
}

extern "C" int risp_bias_table_bwd_grouped(const float* dtab, const float* S, float* dfeat, int G, int N_per_group, const int* slots,
                                           int F, int J, risp_stream_t stream) {
  risp::GroupSlots gs;
  int rc = make_group_slots(&gs, G, slots, "risp_bias_table_bwd");
  if (rc != RISP_OK) return rc;
  const int N = G * N_per_group;
  RISP_REQUIRE(dtab && S && dfeat && N_per_group > 0 && N <= 65535 && F >= 1 && F <= 32 && J >= 1, RISP_E_INVALID, "risp_bias_table_bwd: bad arguments");
  risp::bias_table_bwd_kernel<<<dim3((unsigned)F, (unsigned)N), 256, 0, as_stream(stream)>>>(dtab, S, dfeat, F, J, N_per_group, gs);
  return check_launch("bias_table_bwd_kernel");
}

extern "C" int risp_bias_table_fwd(const float* feat, const float* S, const float* b, float* tab, int N, int F, int J,
                                   risp_stream_t stream) {
  const int slot0 = 0;
  return risp_bias_table_fwd_grouped(feat, S, b, tab, 1, N, &slot0, F, J, stream);
}

extern "C" int risp_bias_table_bwd(const float* dtab, const float* S, float* dfeat, int N, int F, int J, risp_stream_t stream) {
  const int slot0 = 0;
  return risp_bias_table_bwd_grouped(dtab, S, dfeat, 1, N, &slot0, F, J, stream);
}

extern "C" size_t risp_blocked_class_sums_workspace(int N, int C, int H, int K) {
  return (N > 0 && C > 0 && H > 0 && K > 0) ? sizeof(float) * (size_t)N * H * K * ((C + 3) / 4 * 4) : 0;
}

// out (N, K, K, CP): sums of g (blocked, ceil(C/4) groups; optionally masked by [mask > 0]) over the pixels of every border class
extern "C" int risp_blocked_class_sums(const float* g_blk, const float* mask_blk, float* out, int N, int C, int CP, int H, int W, int K,
                                       void* workspace, size_t workspace_bytes, risp_stream_t stream) {
  RISP_REQUIRE(g_blk && out && N > 0 && N <= 65535 && C > 0 && CP >= (C + 3) / 4 * 4 && (K & 1) && K >= 1 && K <= 15 && H >= K - 1 &&
                   W >= K - 1 && H <= 65535, RISP_E_INVALID, "risp_blocked_class_sums: bad arguments");
  RISP_REQUIRE(workspace && workspace_bytes >= risp_blocked_class_sums_workspace(N, C, H, K), RISP_E_WORKSPACE,
               "risp_blocked_class_sums: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int CG = (C + 3) / 4;
  class_sums_rows_kernel<<<dim3((unsigned)H, (unsigned)N), 256, 0, st>>>(g_blk, mask_blk, static_cast<float*>(workspace), CG, H, W, K);
  RISP_REQUIRE(CG * 4 <= 256, RISP_E_UNSUPPORTED, "risp_blocked_class_sums: more than 256 channels");
  class_sums_final_kernel<<<dim3((unsigned)K, (unsigned)N), 256, 0, st>>>(static_cast<const float*>(workspace), out, CG, H, K, CP);
  return check_launch("class_sums");
}

// x, res, y_blk in the blocked layout with channel groups CinG = pad16(Cin)/4, CoutG = pad16(Cout)/4
extern "C" int risp_conv_tc_fwd(const float* x_blk, const float* mask_in_blk, const float* wprep, const float* bias,
                                const float* res_blk, const float* mask_out_blk, float* y_blk, float* y_planar, int N, int Cin,
                                int Cout, int H, int W, int K, int flags, risp_stream_t stream) {
  return risp_conv_tc_fwd_tab(x_blk, mask_in_blk, wprep, bias, nullptr, res_blk, mask_out_blk, y_blk, y_planar, N, Cin, Cout, H, W, K,
                              flags, stream);
}

extern "C" int risp_conv_tc_fwd_tab(const float* x_blk, const float* mask_in_blk, const float* wprep, const float* bias,
                                    const float* bias_tab, const float* res_blk, const float* mask_out_blk, float* y_blk,
                                    float* y_planar, int N, int Cin, int Cout, int H, int W, int K, int flags, risp_stream_t stream) {
  const int slot0 = 0;
  return risp_conv_tc_fwd_grouped(x_blk, mask_in_blk, wprep, bias, bias_tab, res_blk, mask_out_blk, y_blk, y_planar, 1, N, &slot0, 0, 0,
                                  0, 0, Cin, Cout, H, W, K, flags, stream);
}

// G groups x N images per group in ONE launch: the output, the masks and the bias table have G*N images; group g multiplies
// by the prepared weights at wprep + slots[g] * w_group_floats (bias at bias + slots[g] * bias_group_floats); the input /
// residual have N images when x_shared / res_shared is set (every group reads the same images), else G*N.
extern "C" int risp_conv_tc_fwd_grouped(const float* x_blk, const float* mask_in_blk, const float* wprep, const float* bias,
                                        const float* bias_tab, const float* res_blk, const float* mask_out_blk, float* y_blk,
                                        float* y_planar, int G, int N_per_group, const int* slots, long long w_group_floats,
                                        int bias_group_floats, int x_shared, int res_shared, int Cin, int Cout, int H, int W,
                                        int K, int flags, risp_stream_t stream) {
  RISP_REQUIRE(G >= 1 && G <= 8 && N_per_group >= 1 && slots, RISP_E_INVALID, "risp_conv_tc_fwd_grouped: 1..8 groups");
  const int N = G * N_per_group;
  RISP_REQUIRE(x_blk && wprep && y_blk && N > 0 && H > 0 && W > 0 && N <= 65535, RISP_E_INVALID,
               "risp_conv_tc_fwd: bad arguments (the blocked output is required; the planar one is an optional copy)");
  RISP_REQUIRE(!bias_tab || (H >= K - 1 && W >= K - 1 && aligned16(bias_tab)), RISP_E_INVALID,
               "risp_conv_tc_fwd: a bias table needs H, W >= K-1 (disjoint border classes) and 16-byte alignment");
  RISP_REQUIRE(risp_conv_tc_supported(Cin, Cout, K), RISP_E_UNSUPPORTED, "risp_conv_tc_fwd: unsupported shape %d->%d k%d", Cin, Cout, K);
  RISP_REQUIRE(!(flags & RISP_CONV_ADD_RES) || res_blk, RISP_E_INVALID, "risp_conv_tc_fwd: ADD_RES without a residual");
  const int NP = (Cout + 15) / 16 * 16;
  ConvTcArgs a{x_blk, wprep, bias, bias_tab, res_blk, mask_in_blk, mask_out_blk, y_blk, y_planar, risp_conv_tc_padded_channels(Cin) / 4,
               Cout, NP, H, W, flags, Cin, ((Cin + 15) / 16 * 16 + tc_chunk(K) - 1) / tc_chunk(K),
               N_per_group, w_group_floats, bias_group_floats, x_shared, res_shared, {0, 0, 0, 0, 0, 0, 0, 0}};
  for (int g = 0; g < G; ++g) {
    RISP_REQUIRE(slots[g] >= 0 && slots[g] < 256, RISP_E_INVALID, "risp_conv_tc_fwd_grouped: slot out of range");
    a.gslot[g] = (unsigned char)slots[g];
  }
  cudaStream_t st = as_stream(stream);
  const int rc = conv_tc_dispatch(a, N, K, NP, st);
  if (rc != RISP_OK || !y_planar) return rc;
  return risp_from_blocked(y_blk, y_planar, N, Cout, (Cout + 3) / 4, H, W, stream);      // optional planar copy of the result
}

static int conv_tc_dispatch(const ConvTcArgs& a, int N, int K, int NP, cudaStream_t st) {
#define RISP_TC(KK, RR, NN, BB, MM, AA, PP, DD) return launch_tc<KK, 8, RR, NN, BB, MM, AA, PP, DD>(a, N, st)
  // R rows per CTA: the fattest MMA has N = min(R,K)*NP <= 256; TMEM columns = 2*R*NP per CTA.  MM = 2 CTAs per SM
  // where 2*R*NP <= 256 and shared memory <= 112.5 KB, otherwise one CTA with more rows.  BB = weight-stage ring depth,
  // AA = input-tile buffers, PP = chunks prefetched in registers, DD = horizontal taps per weight stage.
  switch (K) {
    case 1:
      switch (NP) { case 16: RISP_TC(1, 8, 16, 4, 2, 1, 1, 1); case 32: RISP_TC(1, 4, 32, 4, 2, 2, 2, 1); case 48: RISP_TC(1, 2, 48, 4, 2, 2, 2, 1); default: RISP_TC(1, 2, 64, 4, 2, 2, 2, 1); }
    case 3:
      // 64 -> 64 (the Path-Restore body), measured at 64 x 256^2 (scripts/probe_tc_variants.py): one weight stage per chunk with
      // a single input buffer 1262 us; one tap per stage + double-buffered inputs (still two CTAs per SM) 1228 us; four rows per
      // CTA with one resident CTA 1500 us
      if (NP == 64 && g_tc_variant == 1) RISP_TC(3, 4, 64, 2, 1, 2, 2, 3);
      if (NP == 64 && g_tc_variant == 2) RISP_TC(3, 2, 64, 2, 2, 1, 2, 3);
      if (NP == 64 && g_tc_variant == 3) RISP_TC(3, 2, 64, 3, 2, 2, 2, 1);
      switch (NP) { case 16: RISP_TC(3, 8, 16, 2, 2, 1, 1, 3); case 32: RISP_TC(3, 4, 32, 2, 2, 1, 2, 3); case 48: RISP_TC(3, 2, 48, 2, 2, 1, 2, 3); default: RISP_TC(3, 2, 64, 3, 2, 2, 1, 1); }
    case 5:
      switch (NP) { case 16: RISP_TC(5, 4, 16, 4, 2, 1, 1, 1); case 32: RISP_TC(5, 4, 32, 4, 2, 1, 1, 1); case 48: RISP_TC(5, 2, 48, 2, 2, 1, 1, 1); default: RISP_TC(5, 2, 64, 2, 2, 1, 1, 1); }
    default:
      switch (NP) { case 16: RISP_TC(9, 8, 16, 4, 1, 1, 1, 1); case 32: RISP_TC(9, 4, 32, 4, 1, 1, 1, 1); case 48: RISP_TC(9, 4, 48, 3, 1, 1, 1, 1); default: RISP_TC(9, 4, 64, 3, 1, 1, 1, 1); }
  }
#undef RISP_TC
}

extern "C" int risp_debug_tc_variant(int variant) { risp::g_tc_variant = variant; return RISP_OK; }

// Timeline of the traced CTA of the last conv_tc launch (only meaningful in a -DRISP_TC_TRACE build): out = HOST
// long long[16]: [0..4] clock64 at start / prologue done / all stages issued / MMAs done / end, [8] cycles waiting at
// chunk boundaries, [9] waiting for weights, [10] issuing MMAs, [11] waiting to refill + issuing the weight fetch.
extern "C" int risp_debug_tc_trace(long long* out_host) {
  RISP_REQUIRE(out_host, RISP_E_INVALID, "risp_debug_tc_trace: null output");
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out_host, risp::g_tc_trace, sizeof(long long) * 16) != cudaSuccess) { set_error("risp_debug_tc_trace: copy failed"); return RISP_E_CUDA; }
  return RISP_OK;
}

// ---- micro-benchmark: raw tcgen05 tf32 issue / execution rate for the operand layout used above ---------------
namespace risp {
template <int NP>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out, int iters, int n_acc, int split3, int a_pw) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* sA = reinterpret_cast<float*>(smem_raw);          // 2 precisions x 2 kgroups x a_pw px x 4
  float* sB = sA + 2 * 2 * 160 * 4;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sB + 2 * 2 * NP * 4);
  uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);
  for (int i = threadIdx.x; i < 2 * 2 * 160 * 4 + 2 * 2 * NP * 4; i += 128) sA[i] = 0.001f * (i & 63);
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) mbar_init(smem_u32(mbar), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (threadIdx.x < 32) {
    if (elect_one()) {
      const uint64_t dA = make_desc(smem_u32(sA), a_pw * 16, 128), dA2 = make_desc(smem_u32(sA + 2 * 160 * 4), a_pw * 16, 128);
      const uint64_t dB = make_desc(smem_u32(sB), NP * 16, 128), dB2 = make_desc(smem_u32(sB + 2 * NP * 4), NP * 16, 128);
      const long long t0 = clock64();
      for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t d = tb + (uint32_t)(((i + u) % n_acc) * NP);
          umma_tf32(d, dA + (uint64_t)u, dB, idesc, (i + u) >= n_acc ? 1u : 0u);
          if (split3) {
            umma_tf32(d, dA2 + (uint64_t)u, dB, idesc, 1u);
            umma_tf32(d, dA + (uint64_t)u, dB2, idesc, 1u);
          }
        }
      }
      const long long t1 = clock64();
      umma_commit(smem_u32(mbar));
      mbar_wait(smem_u32(mbar), 0);
      const long long t2 = clock64();
      out[0] = t1 - t0; out[1] = t2 - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}
}  // namespace risp

// out: DEVICE long long[2] = {issue cycles, total cycles}.  Diagnostic entry (scripts/probe_mma_rate.py).
// a_pw: pixels per k-group of the A operand (its LBO is a_pw*16 bytes), <= 160.
extern "C" int risp_debug_mma_rate(long long* out, int NP, int iters, int n_acc, int split3, int a_pw, risp_stream_t stream) {
  RISP_REQUIRE(out && (NP == 16 || NP == 32 || NP == 64 || NP == 128 || NP == 256) && iters > 0 && (iters % 4) == 0 && n_acc >= 1 &&
                   n_acc * NP <= 512 && a_pw >= 128 && a_pw <= 160, RISP_E_INVALID, "risp_debug_mma_rate: bad arguments");
  const size_t smem = 64 * 1024;
  cudaStream_t st = as_stream(stream);
#define RISP_RATE(NN) { cudaFuncSetAttribute(mma_rate_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
                        mma_rate_kernel<NN><<<1, 128, smem, st>>>(out, iters, n_acc, split3, a_pw); }
  if (NP == 16) RISP_RATE(16) else if (NP == 32) RISP_RATE(32) else if (NP == 64) RISP_RATE(64) else if (NP == 128) RISP_RATE(128) else RISP_RATE(256)
#undef RISP_RATE
  return check_launch("mma_rate_kernel");
}
