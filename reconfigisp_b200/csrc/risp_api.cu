// Error plumbing, chain validation and the deterministic partial-sum finaliser.
#include "risp_common.cuh"
#include "risp_stage.cuh"

#include <mutex>
#include <string.h>

namespace risp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;

int check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return RISP_E_CUDA;
  }
  return RISP_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;  // B200
  }
  return n;
}

int op_param_count(int op, int iarg) {
  switch (op) {
    case RISP_OP_SKIP: return 0;
    case RISP_OP_GAMMA: return 1;
    case RISP_OP_GAIN: return 3;
    case RISP_OP_GAIN_CLIP: return 3;
    case RISP_OP_POLY10: return 30;
    case RISP_OP_GTM: return iarg - 1;
    case RISP_OP_CCM: return 9;
    case RISP_OP_REINHARD: return 2;
    case RISP_OP_CRYSIS: return 1;
    case RISP_OP_FILMIC: return 2;
    default: return -1;
  }
}

bool op_has_bwd(int op) { return op >= RISP_OP_SKIP && op <= RISP_OP_CCM; }

int make_chain(ChainDesc* d, const int* ops, const int* off, const int* iarg, int S, int* P_needed) {
  RISP_REQUIRE(S >= 0 && S <= RISP_MAX_STAGES, RISP_E_INVALID, "chain length %d not in [0,%d]", S, RISP_MAX_STAGES);
  RISP_REQUIRE(S == 0 || (ops && off && iarg), RISP_E_INVALID, "null chain description");
  memset(d, 0, sizeof(*d));
  d->S = S;
  int need = 0, big = 0;
  for (int s = 0; s < S; ++s) {
    int cnt = op_param_count(ops[s], iarg[s]);
    RISP_REQUIRE(cnt >= 0, RISP_E_INVALID, "stage %d: unknown op %d", s, ops[s]);
    if (ops[s] == RISP_OP_GTM)
      RISP_REQUIRE(iarg[s] >= 2 && iarg[s] <= 5, RISP_E_INVALID, "GTM n_seg %d not in [2,5]", iarg[s]);
    RISP_REQUIRE(off[s] >= 0, RISP_E_INVALID, "stage %d: negative parameter offset", s);
    big += op_is_big(ops[s]) ? 1 : 0;
    d->op[s] = ops[s]; d->off[s] = off[s]; d->iarg[s] = iarg[s];
    if (cnt > 0 && off[s] + cnt > need) need = off[s] + cnt;
  }
  RISP_REQUIRE(big <= 1, RISP_E_INVALID, "at most one POLY10/CCM stage per fused chain (got %d); split the chain", big);
  if (P_needed) *P_needed = need;
  return RISP_OK;
}

void chain_slot_list(const ChainDesc& d, SlotList* m) {
  m->n = 0;
  for (int s = 0; s < d.S; ++s) {
    int cnt = op_param_count(d.op[s], d.iarg[s]);
    for (int j = 0; j < cnt; ++j) {
      m->dst[m->n] = (short)(d.off[s] + j);
      m->slot[m->n] = (short)(op_is_big(d.op[s]) ? RISP_SLOT_BIG + j : s * RISP_SMALL_ACC + j);
      m->n++;
    }
  }
}

// out[r][map.dst[e]] = scale * sum_b partial[(r*B + b)*NS + map.slot[e]]   (one warp per entry)
struct SlotMap {
  int n;
  short dst[96];
  short slot[96];
};

__global__ void __launch_bounds__(256)
finalize_kernel(const float* __restrict__ partial, float* __restrict__ out, int R, int B, int NS,
                int P, SlotMap map, float scale, int sum_rows) {
  // one CTA per entry: 256 rows in flight (each thread walks its own rows: fixed assignment), then a fixed-order
  // tree over warps -- deterministic, and the row loads overlap instead of queueing behind one warp
  __shared__ float s_w[8];
  const int e = blockIdx.x, r = blockIdx.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long rows = sum_rows ? (long long)R * B : B;
  const float* p = partial + (sum_rows ? 0 : (long long)r * B) * NS + map.slot[e];
  float acc = 0.f;
  for (long long b = threadIdx.x; b < rows; b += 256) acc += p[b * NS];
  acc = warp_sum(acc);
  if (lane == 0) s_w[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_w[w];
    out[(long long)(sum_rows ? 0 : r) * P + map.dst[e]] = t * scale;
  }
}

int finalize_partials(const float* partial, float* out, int R, int B, int NS, int P, const short* dst,
                      const short* slot, int n, float scale, bool sum_rows, cudaStream_t st) {
  RISP_REQUIRE(n <= 96, RISP_E_INVALID, "too many reduced entries (%d)", n);
  if (n == 0) return RISP_OK;
  SlotMap m;
  m.n = n;
  for (int i = 0; i < n; ++i) { m.dst[i] = dst[i]; m.slot[i] = slot[i]; }
  dim3 grid(n, sum_rows ? 1 : R);
  finalize_kernel<<<grid, 256, 0, st>>>(partial, out, R, B, NS, P, m, scale, sum_rows ? 1 : 0);
  return check_launch("finalize_kernel");
}

// ---- two-level reduction in ONE launch (the tuning step's finaliser) ---------------------------------------
// Level 1: kFinBlocks CTAs per output row each reduce a contiguous slice of partial rows for ALL slots at once (a warp
// reads whole partial rows: coalesced), fixed warp order.  Level 2: the last CTA to finish (ticket counter) adds the
// kFinBlocks slice sums in index order, applies the slot -> parameter map and the scales, zero-fills unmapped
// parameters and (optionally) writes the loss.  Fixed summation order => bit-reproducible.
__global__ void __launch_bounds__(256)
finalize_rows_kernel(const float* __restrict__ partial, float* __restrict__ scratch, unsigned int* __restrict__ counter,
                     float* __restrict__ out, float* __restrict__ loss_out, int R, int B, int NS, int P, SlotMap map,
                     float scale, float loss_scale, int loss_slot, int sum_rows) {
  __shared__ float s_part[8][64];
  __shared__ float s_tot[64];
  __shared__ int s_last;
  const int rp = blockIdx.y, G = gridDim.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long rows_total = sum_rows ? (long long)R * B : B;
  const long long base = sum_rows ? 0 : (long long)rp * B;
  const long long per = (rows_total + G - 1) / G;
  const long long lo = (long long)blockIdx.x * per, hi = (lo + per < rows_total) ? lo + per : rows_total;
  float acc0 = 0.f, acc1 = 0.f;
  for (long long row = lo + wid; row < hi; row += 8) {
    const float* p = partial + (base + row) * NS;
    acc0 += p[lane];
    if (lane + 32 < NS) acc1 += p[lane + 32];
  }
  s_part[wid][lane] = acc0;
  s_part[wid][lane + 32] = acc1;
  __syncthreads();
  if (threadIdx.x < NS) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_part[w][threadIdx.x];
    scratch[((long long)rp * G + blockIdx.x) * NS + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter + rp, 1u) == (unsigned int)(G - 1));
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < NS) {
    float t = 0.f;
    for (int g = 0; g < G; ++g) t += __ldcg(scratch + ((long long)rp * G + g) * NS + threadIdx.x);
    s_tot[threadIdx.x] = t;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < P; q += blockDim.x) {
    float v = 0.f;
    for (int e = 0; e < map.n; ++e)
      if (map.dst[e] == q) v = s_tot[map.slot[e]] * scale;
    out[(long long)rp * P + q] = v;
  }
  if (loss_out && rp == 0 && threadIdx.x == 0) loss_out[0] = s_tot[loss_slot] * loss_scale;
  if (threadIdx.x == 0) counter[rp] = 0;      // ready for the next launch on this workspace
}

size_t finalize_rows_workspace(int R, int NS) {
  return ((size_t)R * kFinBlocks * NS + 64) * sizeof(float);     // slice sums + ticket counters
}

int finalize_rows(const float* partial, void* scratch_ws, float* out, float* loss_out, int R, int B, int NS, int P,
                  const short* dst, const short* slot, int n, float scale, float loss_scale, int loss_slot, bool sum_rows,
                  cudaStream_t st) {
  RISP_REQUIRE(n <= 96 && NS <= 64 && R <= 64, RISP_E_INVALID, "finalize_rows: %d entries / %d slots / %d rows", n, NS, R);
  SlotMap m;
  m.n = n;
  for (int i = 0; i < n; ++i) { m.dst[i] = dst[i]; m.slot[i] = slot[i]; }
  float* scratch = static_cast<float*>(scratch_ws);
  unsigned int* counter = reinterpret_cast<unsigned int*>(scratch + (size_t)R * kFinBlocks * NS);
  const int Rp = sum_rows ? 1 : R;
  // the ticket counters must be zero: the kernel re-arms them, but the workspace is caller-owned scratch
  if (cudaMemsetAsync(counter, 0, sizeof(unsigned int) * Rp, st) != cudaSuccess) { set_error("finalize_rows: memset failed"); return RISP_E_CUDA; }
  finalize_rows_kernel<<<dim3(kFinBlocks, Rp), 256, 0, st>>>(partial, scratch, counter, out, loss_out, R, B, NS, P, m, scale,
                                                            loss_scale, loss_slot, sum_rows ? 1 : 0);
  return check_launch("finalize_rows_kernel");
}

}  // namespace risp

extern "C" {
int risp_abi_version(void) { return RISP_ABI_VERSION; }
const char* risp_last_error(void) { return risp::g_err; }
int risp_sm_count(void) { return risp::sm_count(); }
long long risp_launch_count(void) { return (long long)risp::g_launches; }
}
