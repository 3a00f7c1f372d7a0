// fa_tc_bwd.cu -- tcgen05 / TMEM / TMA flash-attention backward for sm_100a (dense + circulant).
//
// Replaces the hot loop of dense_fa_backward (reference src/dense.jl:104-167, broken) ==
// OneDFastBack (src_cpp/FlashAttention.cpp:194-252) and its band-restricted circulant form
//  for 16-bit inputs with d == dv in {64, 128}:
//     P = exp(tau Q K^T - m) / l,  dV = P^T dO,  dP = dO V^T,  D = rowsum(dO o O),
//     dS = P o (dP - D),  dQ = tau dS K,  dK = tau dS^T Q.
// The reference accumulates dK/dV from all row blocks into shared arrays (a race when run in
// parallel, src_cpp/FlashAttention.cpp:299-312).  Here the pass is split into two deterministic
// kernels that share one template -- no atomics, bitwise reproducible:
//     KIND 0  (key-owner)    CTA owns 128 keys,    streams 64-query tiles:  dV, dK
//     KIND 1  (query-owner)  CTA owns 128 queries, streams 64-key tiles:    dQ
// With "row" = owner token (TMEM lane) and "col" = streamed token, each step g computes
//     T1[b] = X1 Y1(g)^T     (S^T or S;   X1/Y1 = K/Q or Q/K;   MN-major smem operands, N = 64)
//     T2[b] = X2 Y2(g)^T     (dP^T or dP; X2/Y2 = V/dO or dO/V)
//     KIND 0: accV += P^T(g) Y2(g)   (A = 16-bit P^T in TMEM over T1[b], B = dO tile, K-major)
//             accK += dS^T(g) Y1(g)  (A = 16-bit dS^T in TMEM over T2[b], B = Q tile, K-major)
//     KIND 1: accQ += dS(g) Y1(g)    (B = K tile, K-major)
// The [B][d][N] token-contiguous layout (src/dense.jl:6-8) serves BOTH uses of a streamed tile:
// a 64-token x d-channel SWIZZLE_128B TMA box is the MN-major operand of T1/T2 and, read with a
// K-major descriptor, the B operand of the accumulating MMAs.  Nothing is transposed.
//
// The softmax statistics are folded on the device by bwd_prep_kernel into
//     nlse_i = -(m_i + ln l_i) log2(e)      so that  P_ij = exp2(s_ij tau log2(e) + nlse_i)
//     ndelta_i = -sum_c dO_ic O_ic
// padded to a multiple of 128 tokens with (-inf, 0), which makes out-of-range queries vanish
// (P = 0) without masks; out-of-range keys are zero-filled by TMA and contribute exact zeros.
//
// CTA = 384 threads, one CTA per SM:  warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM
// allocator, warps 4-7 / 8-11 two element-wise warpgroups that ALTERNATE steps (the backward has
// no running max, so steps are independent): warpgroup w owns TMEM buffers T1[w], T2[w].
// TMEM columns: T1[b] at 64 b, T2[b] at 128 + 64 b, accumulators at 256 (+ D).
#include <cuda.h>
#include <algorithm>
#include "fa_common.cuh"
#include "fa_ptx.cuh"

namespace fa {

int make_tmap_public(CUtensorMap* tm, const void* base, int dtype, long long N, int D, long long B, long long stride_c = 0, long long stride_b = 0);
bool tc_fwd_supported(const Geo& g, int dtype);

namespace {

using namespace ptx;

constexpr int BW_THREADS = 384;
constexpr int BT = 64;                      // streamed tokens per step (one TMA box)
constexpr float LOG2E = 1.4426950408889634f;

template <int D, int KIND, int OT = 0>
struct BCfg {
  static constexpr int NS = (D == 128) ? 4 : 6;          // streamed stages
  static constexpr int BOX_BYTES = 64 * D * 2;           // 64 tokens x D channels, 16-bit
  static constexpr int OWN_BYTES = 2 * BOX_BYTES;        // 128 owner tokens of one tensor
  static constexpr int STAGE_BYTES = 2 * BOX_BYTES;      // Y1 box + Y2 box
  static constexpr int STAT_BYTES = 512;                 // nlse[64], ndelta[64] (KIND 0 only)
  // KIND 1 keeps its owner tiles (Q, dO) in TMEM as K-major A operands: an SS-mode M=128, N=64 MMA
  // reads 6 KB of shared memory per 32 clk of math (192 B/clk > the 128 B/clk port), a TS-mode one 2 KB.
  // KIND 0 has no TMEM left for that (T1, T2 double-buffered + dV + dK = 512 columns at d = 128).
  // (Dense only: with the few-step loops of a circulant band the synchronous owner load costs more
  // than the SS-mode penalty, measured 1.34 vs 1.07 ms at C4.)
  static constexpr bool OWN_TMEM = (KIND == 1) && (OT == 1);
  static constexpr int OFF_X1 = 0;
  static constexpr int OFF_X2 = OWN_BYTES;
  static constexpr int OFF_Y = OWN_TMEM ? 0 : 2 * OWN_BYTES;
  static constexpr int OFF_STAT = OFF_Y + NS * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_STAT + NS * STAT_BYTES;
  static constexpr int BAR_OWN = 0;
  static constexpr int BAR_FULL = 1;                     // [NS]
  static constexpr int BAR_EMPTY = BAR_FULL + NS;        // [NS]
  static constexpr int BAR_T1 = BAR_EMPTY + NS;          // [2]
  static constexpr int BAR_T2 = BAR_T1 + 2;              // [2]
  static constexpr int BAR_P = BAR_T2 + 2;               // [2]
  static constexpr int BAR_DS = BAR_P + 2;               // [2]
  static constexpr int BAR_ACC = BAR_DS + 2;             // [1]
  static constexpr int NUM_BARS = BAR_ACC + 1;
  static constexpr int OFF_TMEM_SLOT = OFF_BAR + NUM_BARS * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM_SLOT + 16 + 1024;
  static constexpr int COL_T1 = 0, COL_T2 = 128, COL_ACC = 256;
  static constexpr int COL_X = 384;                      // KIND 1: X1 at 384, X2 at 384 + D/2 (16-bit, K-major)
};

struct BwdParams {
  void* out0;            // KIND 0: dV    KIND 1: dQ
  void* out1;            // KIND 0: dK
  const float* nlse;     // [B][Npad]
  const float* ndelta;   // [B][Npad]  (scaled by sv*sg when the inputs were re-encoded)
  const float* amax;     // device: max|q|, max|k|, max|v|, max|dO| of the re-encoded inputs, or NULL
  int out_f32;           // gradients stored as float32 whatever the input dtype (partials of the ring pass)
  const void *x1, *x2;   // KIND 1: owner tensors (Q, dO) as raw [B][D][N] pointers (loaded straight into TMEM)
  int N, Npad, W, p, mode;
  int X, Y;              // 2-D periodic neighbourhood (mode circulant, X > 0): image extents, N = X * Y, x fastest
  float scale_log2;      // tau * log2(e)
  float tau;
  int dbg;               // FA_TRACE builds: timing knock-outs with WRONG results (1 = half the ex2, 2 = half the T MMAs, 4 = half the accumulating MMAs)
};

__host__ __device__ inline int fdiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// Power of two s with amax * s in [4, 8): bf16 inputs are re-encoded as x * s in fp16 (exact for every
// element within 2^-17 of the tensor's max; smaller ones keep an absolute error below 2^-28 max), so
// that P and dS can be rounded to fp16 (11 significand bits) instead of bf16 (8) for the MMAs.
__host__ __device__ inline float norm_scale(float amax) {
  if (!(amax > 0.f) || !(amax < 3.0e38f)) return 1.f;
  int e;
  frexpf(amax, &e);                       // amax in [2^(e-1), 2^e)
  int k = 3 - e;
  k = k > 100 ? 100 : (k < -100 ? -100 : k);
  return ldexpf(1.f, k);
}

// max |x| of up to four tensors (blockIdx.y selects), as uint bit patterns of non-negative floats
__global__ void bwd_amax_kernel(const __nv_bfloat16* x0, const __nv_bfloat16* x1, const __nv_bfloat16* x2,
                                const __nv_bfloat16* x3, size_t n, float* amax) {
  const __nv_bfloat16* x = blockIdx.y == 0 ? x0 : blockIdx.y == 1 ? x1 : blockIdx.y == 2 ? x2 : x3;
  const uint4* xv = reinterpret_cast<const uint4*>(x);
  uint32_t mx = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 8; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 u = xv[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t lo = (w[j] << 16) & 0x7fffffffu, hi = w[j] & 0x7fff0000u;
      mx = max(mx, max(lo, hi));
    }
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(reinterpret_cast<unsigned int*>(amax) + blockIdx.y, mx);
}

// y = fp16(x * norm_scale(amax)) for four tensors
__global__ void bwd_reencode_kernel(const __nv_bfloat16* x0, const __nv_bfloat16* x1, const __nv_bfloat16* x2,
                                    const __nv_bfloat16* x3, __half* y0, __half* y1, __half* y2, __half* y3,
                                    size_t n, const float* amax) {
  const __nv_bfloat16* x = blockIdx.y == 0 ? x0 : blockIdx.y == 1 ? x1 : blockIdx.y == 2 ? x2 : x3;
  __half* y = blockIdx.y == 0 ? y0 : blockIdx.y == 1 ? y1 : blockIdx.y == 2 ? y2 : y3;
  const float s = norm_scale(amax[blockIdx.y]);
  const uint4* xv = reinterpret_cast<const uint4*>(x);
  uint4* yv = reinterpret_cast<uint4*>(y);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 8; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 u = xv[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2 h = __floats2half2_rn(__uint_as_float(w[j] << 16) * s, __uint_as_float(w[j] & 0xffff0000u) * s);
      r[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    yv[i] = make_uint4(r[0], r[1], r[2], r[3]);
  }
}

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}

template <int FMT>
__device__ __forceinline__ uint32_t pack16(float a, float b) {
  if (FMT == 1) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <int FMT>
__device__ __forceinline__ float2 unpack16(uint32_t u) {
  if (FMT == 1) return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}
template <int FMT> struct OutT { using type = __half; };
template <> struct OutT<1> { using type = __nv_bfloat16; };

// nlse = -(m + ln l) log2 e, ndelta = -sum_c dO o O; padded entries (-inf, 0)
template <typename T>
__global__ void bwd_prep_kernel(const T* __restrict__ o, const T* __restrict__ d_o, const float* __restrict__ l,
                                const float* __restrict__ m, float* __restrict__ nlse, float* __restrict__ ndelta,
                                int N, int Npad, int dv, const float* __restrict__ amax) {
  const int b = blockIdx.y;
  const float dscale = amax ? norm_scale(amax[2]) * norm_scale(amax[3]) : 1.f;   // dP is formed from v*sv, dO*sg
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Npad; i += gridDim.x * blockDim.x) {
    float nl = -INFINITY, nd = 0.f;
    if (i < N) {
      const T* po = o + (size_t)b * dv * N + i;
      const T* pg = d_o + (size_t)b * dv * N + i;
      float acc = 0.f;
#pragma unroll 4
      for (int c = 0; c < dv; ++c) acc = fmaf(to_f32(po[(size_t)c * N]), to_f32(pg[(size_t)c * N]), acc);
      nd = -acc * dscale;
      nl = -(m[(size_t)b * N + i] + logf(l[(size_t)b * N + i])) * LOG2E;
    }
    nlse[(size_t)b * Npad + i] = nl;
    ndelta[(size_t)b * Npad + i] = nd;
  }
}

template <int D, int FMT, int KIND, int OBF, int OT>
__global__ void __launch_bounds__(BW_THREADS, 1)
tc_bwd_kernel(const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
              const __grid_constant__ CUtensorMap tmY1, const __grid_constant__ CUtensorMap tmY2,
              const BwdParams prm) {
  using C = BCfg<D, KIND, OT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sX1 = sbase + C::OFF_X1, sX2 = sbase + C::OFF_X2, sY = sbase + C::OFF_Y;
  const uint32_t sStat = sbase + C::OFF_STAT;
  const uint32_t bars = sbase + C::OFF_BAR;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  const uint32_t tmem_slot = sbase + C::OFF_TMEM_SLOT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const bool circ = prm.mode == MODE_CIRCULANT;
  // 2-D (td): the owner tile is up to 128 consecutive x of ONE image row yo; blockIdx.x = row * tiles_per_row + tile
  const bool td = circ && prm.X > 0;
  const int tpr = td ? (prm.X + 127) / 128 : 1;
  const int yo = td ? (int)(blockIdx.x / tpr) : 0;
  const int x0 = td ? (int)(blockIdx.x % tpr) * 128 : 0;
  const int tq = td ? (prm.X - x0 < 128 ? prm.X - x0 : 128) : 128;
  const int t0 = td ? yo * prm.X + x0 : blockIdx.x * 128;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmX1); prefetch_tensormap(&tmX2); prefetch_tensormap(&tmY1); prefetch_tensormap(&tmY2);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar(C::BAR_OWN), C::OWN_TMEM ? 256 : 1);
    for (int i = 0; i < C::NS; ++i) { mbar_init(bar(C::BAR_FULL + i), 1); mbar_init(bar(C::BAR_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(C::BAR_T1 + i), 1); mbar_init(bar(C::BAR_T2 + i), 1);
      mbar_init(bar(C::BAR_P + i), 128); mbar_init(bar(C::BAR_DS + i), 128);
    }
    mbar_init(bar(C::BAR_ACC), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // streamed-tile range (unwrapped token coordinates; circulant wraps tile-aligned, N % 64 == 0)
  int cbase = 0, ns = (prm.N + BT - 1) / BT;
  int nx = 1;
  if (td) {
    // streamed tiles = W image rows x the nx 64-token tiles of a row that meet the x band of the owner tile
    // (keys of (x, y): (mod(x - p + s, X), mod(y - p + t, Y)); queries of a key: the mirror image)
    if (KIND == 0) { cbase = fdiv(x0 + prm.p - prm.W + 1, BT) * BT; nx = fdiv(x0 + tq - 1 + prm.p - cbase, BT) + 1; }
    else { cbase = fdiv(x0 - prm.p, BT) * BT; nx = fdiv(x0 + tq - 1 - prm.p + prm.W - 1 - cbase, BT) + 1; }
    if (nx * BT >= prm.X) { nx = prm.X / BT; cbase = 0; }
    ns = prm.W * nx;
  } else if (circ) {
    if (KIND == 0) {   // rows = keys j, cols = queries i with j - (W-1-p) <= i <= j + p
      cbase = fdiv(t0 + prm.p - prm.W + 1, BT) * BT;
      ns = fdiv(t0 + 127 + prm.p - cbase, BT) + 1;
    } else {           // rows = queries i, cols = keys j with i - p <= j <= i - p + W - 1
      cbase = fdiv(t0 - prm.p, BT) * BT;
      ns = fdiv(t0 + 127 - prm.p + prm.W - 1 - cbase, BT) + 1;
    }
  }

  // first token of streamed tile g; xs = its image column (td)
  auto stream_tile = [&](int g, int& xs) {
    if (!td) { xs = 0; return circ ? (int)pmod(cbase + BT * g, prm.N) : BT * g; }
    const int t = g / nx, jx = g - t * nx;
    xs = (int)pmod(cbase + BT * jx, prm.X);
    const int yy = (KIND == 0) ? (int)pmod(yo + prm.p - prm.W + 1 + t, prm.Y) : (int)pmod(yo - prm.p + t, prm.Y);
    return yy * prm.X + xs;
  };

  if (warp < 4) {
    setmaxnreg_dec<64>();
    if (warp == 0 && lane == 0) {
      // ------------------------------------------------------------ TMA producer
      if (!C::OWN_TMEM) {
        mbar_arrive_expect_tx(bar(C::BAR_OWN), 2 * C::OWN_BYTES);
        tma_load_3d(sX1, &tmX1, bar(C::BAR_OWN), t0, 0, b);
        tma_load_3d(sX1 + C::BOX_BYTES, &tmX1, bar(C::BAR_OWN), t0 + 64, 0, b);
        tma_load_3d(sX2, &tmX2, bar(C::BAR_OWN), t0, 0, b);
        tma_load_3d(sX2 + C::BOX_BYTES, &tmX2, bar(C::BAR_OWN), t0 + 64, 0, b);
      }
      for (int g = 0; g < ns; ++g) {
        const int s = g % C::NS;
        int xs_unused;
        const int tok = stream_tile(g, xs_unused);
        mbar_wait(bar(C::BAR_EMPTY + s), ((uint32_t)(g / C::NS) & 1u) ^ 1u);
        mbar_arrive_expect_tx(bar(C::BAR_FULL + s), C::STAGE_BYTES + (KIND == 0 ? C::STAT_BYTES : 0));
        tma_load_3d(sY + s * C::STAGE_BYTES, &tmY1, bar(C::BAR_FULL + s), tok, 0, b);
        tma_load_3d(sY + s * C::STAGE_BYTES + C::BOX_BYTES, &tmY2, bar(C::BAR_FULL + s), tok, 0, b);
        if (KIND == 0) {
          bulk_load_1d(sStat + s * C::STAT_BYTES, prm.nlse + (size_t)b * prm.Npad + tok, 256, bar(C::BAR_FULL + s));
          bulk_load_1d(sStat + s * C::STAT_BYTES + 256, prm.ndelta + (size_t)b * prm.Npad + tok, 256, bar(C::BAR_FULL + s));
        }
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------ MMA issuer (warp-uniform loop)
      constexpr uint32_t idesc_t = make_idesc_f16(FMT, FMT, C::OWN_TMEM ? 0 : 1, 1, 128, BT);   // A MN-major smem (or TMEM), B MN-major
      constexpr uint32_t idesc_acc = make_idesc_f16(FMT, FMT, 0, 0, 128, D);    // A in TMEM, B K-major
      const uint64_t x1d = make_smem_desc_sw128(sX1, C::BOX_BYTES, 1024);
      const uint64_t x2d = make_smem_desc_sw128(sX2, C::BOX_BYTES, 1024);
      const uint64_t ymn = make_smem_desc_sw128(sY, C::BOX_BYTES, 1024);         // MN-major view of a box
      const uint64_t ykm = make_smem_desc_sw128(sY, 16, 1024);                   // K-major view of a box
      const uint32_t tT1 = tmem_base + C::COL_T1, tT2 = tmem_base + C::COL_T2;
      const uint32_t tA0 = tmem_base + C::COL_ACC, tA1 = tmem_base + C::COL_ACC + D;
      mbar_wait(bar(C::BAR_OWN), 0);
      auto issue_T = [&](int g) {
        const int s = g % C::NS, bb = g & 1;
        mbar_wait(bar(C::BAR_FULL + s), (uint32_t)(g / C::NS) & 1u);
        tc_fence_after();
        const uint64_t y1 = ymn + (uint64_t)(s * (C::STAGE_BYTES >> 4));
        const uint64_t y2 = y1 + (uint64_t)(C::BOX_BYTES >> 4);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
#ifdef FA_TRACE
            if ((prm.dbg & 2) && ks >= D / 32) continue;
#endif
            if (C::OWN_TMEM) mma_ts(tT1 + 64 * bb, tmem_base + C::COL_X + ks * 8, y1 + (uint64_t)(ks * 128), idesc_t, ks > 0 ? 1u : 0u);
            else mma_ss(tT1 + 64 * bb, x1d + (uint64_t)(ks * 128), y1 + (uint64_t)(ks * 128), idesc_t, ks > 0 ? 1u : 0u);
          }
          tc_commit(bar(C::BAR_T1 + bb));
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
#ifdef FA_TRACE
            if ((prm.dbg & 2) && ks >= D / 32) continue;
#endif
            if (C::OWN_TMEM) mma_ts(tT2 + 64 * bb, tmem_base + C::COL_X + D / 2 + ks * 8, y2 + (uint64_t)(ks * 128), idesc_t, ks > 0 ? 1u : 0u);
            else mma_ss(tT2 + 64 * bb, x2d + (uint64_t)(ks * 128), y2 + (uint64_t)(ks * 128), idesc_t, ks > 0 ? 1u : 0u);
          }
          tc_commit(bar(C::BAR_T2 + bb));
        }
        __syncwarp();
      };
      issue_T(0);
      if (ns > 1) issue_T(1);
      for (int j = 0; j < ns; ++j) {
        const int s = j % C::NS, bb = j & 1;
        const uint32_t par = (uint32_t)(j >> 1) & 1u;
        const uint64_t y1 = ykm + (uint64_t)(s * (C::STAGE_BYTES >> 4));
        const uint64_t y2 = y1 + (uint64_t)(C::BOX_BYTES >> 4);
        if (KIND == 0) {
          mbar_wait(bar(C::BAR_P + bb), par);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < BT / 16; ++ks) {
#ifdef FA_TRACE
              if ((prm.dbg & 4) && ks >= BT / 32) continue;
#endif
              mma_ts(tA0, tT1 + 64 * bb + ks * 8, y2 + (uint64_t)(ks * 2), idesc_acc, (j > 0 || ks > 0) ? 1u : 0u);
            }
          }
          __syncwarp();
        }
        mbar_wait(bar(C::BAR_DS + bb), par);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < BT / 16; ++ks) {
#ifdef FA_TRACE
            if ((prm.dbg & 4) && ks >= BT / 32) continue;
#endif
            mma_ts(KIND == 0 ? tA1 : tA0, tT2 + 64 * bb + ks * 8, y1 + (uint64_t)(ks * 2), idesc_acc, (j > 0 || ks > 0) ? 1u : 0u);
          }
          tc_commit(bar(C::BAR_EMPTY + s));
        }
        __syncwarp();
        if (j + 2 < ns) issue_T(j + 2);
      }
      if (elect_one()) tc_commit(bar(C::BAR_ACC));
      __syncwarp();
    }
  } else {
    // -------------------------------------------------------------- element-wise warpgroups
    setmaxnreg_inc<216>();
    const int wg = (warp - 4) >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tT1 = tmem_base + lane_addr + C::COL_T1 + 64 * wg;
    const uint32_t tT2 = tmem_base + lane_addr + C::COL_T2 + 64 * wg;
    const int row_tok = t0 + row;
    // scale factors of re-encoded inputs (powers of two; 1 when the inputs are used as they are)
    float sq = 1.f, sk = 1.f, sv = 1.f, sg = 1.f;
    if (prm.amax) { sq = norm_scale(prm.amax[0]); sk = norm_scale(prm.amax[1]); sv = norm_scale(prm.amax[2]); sg = norm_scale(prm.amax[3]); }
    const float sl2 = prm.scale_log2 / (sq * sk);
    const float2 scale2 = make_float2(sl2, sl2);
    float2 rnl = make_float2(0.f, 0.f), rnd = make_float2(0.f, 0.f);
    if (KIND == 1) {
      const bool own = !td || row < tq;                    // td: rows past the image row belong to another tile
      const float a = own ? prm.nlse[(size_t)b * prm.Npad + row_tok] : -INFINITY, d = own ? prm.ndelta[(size_t)b * prm.Npad + row_tok] : 0.f;
      rnl = make_float2(a, a); rnd = make_float2(d, d);
    }
    const int WW = prm.W, pp = prm.p;

    if (C::OWN_TMEM) {
      // owner rows straight from global memory into TMEM (thread == row; a warp reads 32 consecutive
      // tokens per channel): warpgroup 0 loads X1 (Q), warpgroup 1 loads X2 (dO); two channels per column
      const unsigned short* src = static_cast<const unsigned short*>(wg == 0 ? prm.x1 : prm.x2) + (size_t)b * D * prm.N + row_tok;
      const uint32_t tX = tmem_base + lane_addr + C::COL_X + wg * (D / 2);
      const bool inr = row_tok < prm.N;
#pragma unroll 1
      for (int c0 = 0; c0 < D; c0 += 64) {
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          uint32_t lo = 0, hi = 0;
          if (inr) {
            lo = __ldg(src + (size_t)(c0 + 2 * i) * prm.N);
            hi = __ldg(src + (size_t)(c0 + 2 * i + 1) * prm.N);
          }
          r[i] = lo | (hi << 16);
        }
        tmem_st32(tX + c0 / 2, r);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(bar(C::BAR_OWN));
    }

    for (int j = wg; j < ns; j += 2) {
      const uint32_t par = (uint32_t)(j >> 1) & 1u;
      const int s = j % C::NS;
      const uint32_t stat = sStat + s * C::STAT_BYTES;
      if (KIND == 0) mbar_wait(bar(C::BAR_FULL + s), (uint32_t)(j / C::NS) & 1u);   // nlse / ndelta landed
      // band limits of this step in column units: lo <= col < hi is inside the window
      int lo = 0, hi = BT, lo2 = 0, hi2 = 0;                 // td: second interval = the band wrapped round the image row
      if (td) {
        int xs;
        stream_tile(j, xs);
        const int rx = x0 + row;
        if (KIND == 0) { hi = (int)pmod(rx + pp - xs, prm.X) + 1; lo = hi - WW; lo2 = lo + prm.X; hi2 = hi + prm.X; }   // (row + p - col) mod X < W
        else { lo = (int)pmod(rx - pp - xs, prm.X); hi = lo + WW; lo2 = lo - prm.X; hi2 = hi - prm.X; }               // (col - row + p) mod X < W
      } else if (circ) {
        const int c0 = cbase + BT * j;                       // unwrapped token of column 0
        if (KIND == 0) { hi = row_tok + pp - c0 + 1; lo = hi - WW; }      // 0 <= row + p - col < W
        else { lo = row_tok - pp - c0; hi = lo + WW; }                     // 0 <= col - row + p < W
      }
      const bool edge = lo > 0 || hi < BT;

      // P and dS are formed in fp32 and rounded to the 16-bit MMA format exactly once each
      uint32_t pk[32], dk[32];
      mbar_wait(bar(C::BAR_T1 + wg), par);
      mbar_wait(bar(C::BAR_T2 + wg), par);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        // band kernels: a 32-column chunk outside the band of every row of this warp has P = dS = 0 -- nothing to read
        // or exponentiate (2-D neighbourhood: W of the 64 columns of a step are inside; 1-D: the edge tiles)
        if (circ) {
          const bool outside = (hi <= 32 * c || lo >= 32 * c + 32) && (hi2 <= 32 * c || lo2 >= 32 * c + 32);
          if (__all_sync(0xffffffffu, outside)) {
#pragma unroll
            for (int e = 0; e < 16; ++e) { pk[16 * c + e] = 0u; dk[16 * c + e] = 0u; }
            continue;
          }
        }
        uint32_t sc[32], dp[32];
        tmem_ld32(tT1 + 32 * c, sc);
        tmem_ld32(tT2 + 32 * c, dp);
        tmem_wait_ld();
        if (edge) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int col = 32 * c + e;
            if ((col < lo || col >= hi) && (col < lo2 || col >= hi2)) sc[e] = 0xff800000u;   // -inf -> P = 0, dS = 0
          }
        }
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          float4 nl4, nd4;
          if (KIND == 0) {
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(nl4.x), "=f"(nl4.y), "=f"(nl4.z), "=f"(nl4.w) : "r"(stat + 4u * (uint32_t)(32 * c + e)));
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(nd4.x), "=f"(nd4.y), "=f"(nd4.z), "=f"(nd4.w) : "r"(stat + 256u + 4u * (uint32_t)(32 * c + e)));
          } else {
            nl4 = make_float4(rnl.x, rnl.x, rnl.x, rnl.x);
            nd4 = make_float4(rnd.x, rnd.x, rnd.x, rnd.x);
          }
          const float2 xa = __ffma2_rn(make_float2(__uint_as_float(sc[e]), __uint_as_float(sc[e + 1])), scale2, make_float2(nl4.x, nl4.y));
          const float2 xb = __ffma2_rn(make_float2(__uint_as_float(sc[e + 2]), __uint_as_float(sc[e + 3])), scale2, make_float2(nl4.z, nl4.w));
#ifdef FA_TRACE
          const bool noex = (prm.dbg & 1) && c == 1;
          const float2 pa = noex ? xa : make_float2(ex2(xa.x), ex2(xa.y)), pb = noex ? xb : make_float2(ex2(xb.x), ex2(xb.y));
#else
          const float2 pa = make_float2(ex2(xa.x), ex2(xa.y)), pb = make_float2(ex2(xb.x), ex2(xb.y));
#endif
          const float2 ta = __fadd2_rn(make_float2(__uint_as_float(dp[e]), __uint_as_float(dp[e + 1])), make_float2(nd4.x, nd4.y));
          const float2 tb = __fadd2_rn(make_float2(__uint_as_float(dp[e + 2]), __uint_as_float(dp[e + 3])), make_float2(nd4.z, nd4.w));
          const float2 da = __fmul2_rn(pa, ta), db = __fmul2_rn(pb, tb);
          if (KIND == 0) {
            pk[16 * c + (e >> 1)] = pack16<FMT>(pa.x, pa.y);
            pk[16 * c + (e >> 1) + 1] = pack16<FMT>(pb.x, pb.y);
          }
          dk[16 * c + (e >> 1)] = pack16<FMT>(da.x, da.y);
          dk[16 * c + (e >> 1) + 1] = pack16<FMT>(db.x, db.y);
        }
      }
      if (KIND == 0) tmem_st32(tT1, pk);    // P^T (16-bit) over the first 32 columns of T1[wg]
      tmem_st32(tT2, dk);                   // dS  (16-bit) over the first 32 columns of T2[wg]
      tmem_wait_st();
      tc_fence_before();
      if (KIND == 0) mbar_arrive(bar(C::BAR_P + wg));
      mbar_arrive(bar(C::BAR_DS + wg));
    }

    // ---- epilogue: accumulators -> global (token-contiguous rows: a warp writes 32 consecutive tokens)
    mbar_wait(bar(C::BAR_ACC), 0);
    tc_fence_after();
    using OT = typename OutT<OBF>::type;
    const bool in_range = td ? row < tq : row_tok < prm.N;
    // KIND 0: warpgroup 0 stores dV (acc0), warpgroup 1 stores tau * dK (acc1);
    // KIND 1: each warpgroup stores half the channels of tau * dQ (acc0)
    const uint32_t tacc = tmem_base + lane_addr + C::COL_ACC + ((KIND == 0 && wg == 1) ? D : 0);
    OT* out = static_cast<OT*>((KIND == 0 && wg == 1) ? prm.out1 : prm.out0) + (size_t)b * D * prm.N + row_tok;
    // dV = P^T (dO sg) / sg;  dK = tau dS'^T (Q sq) / (sv sg sq);  dQ = tau dS' (K sk) / (sv sg sk),  dS' = dS sv sg
    const float mul = (KIND == 0 && wg == 0) ? 1.f / sg : prm.tau / (sv * sg * (KIND == 0 ? sq : sk));
    const int c_lo = (KIND == 1) ? wg * (D / 64) : 0, c_hi = (KIND == 1) ? (wg + 1) * (D / 64) : D / 32;
    float* outf = static_cast<float*>((KIND == 0 && wg == 1) ? prm.out1 : prm.out0) + (size_t)b * D * prm.N + row_tok;
#pragma unroll 1
    for (int c = c_lo; c < c_hi; ++c) {
      uint32_t o[32];
      tmem_ld32(tacc + 32 * c, o);
      tmem_wait_ld();
      if (in_range) {
        if (prm.out_f32) {
#pragma unroll
          for (int e = 0; e < 32; ++e) outf[(size_t)(32 * c + e) * prm.N] = __uint_as_float(o[e]) * mul;
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            out[(size_t)(32 * c + e) * prm.N] = from_f32<OT>(__uint_as_float(o[e]) * mul);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int D, int FMT, int OBF>
int launch_tc_bwd(const Geo& g, const BwdArgs& a, const void* q, const void* k, const void* v, const void* d_o,
                  const float* amax, float* nlse, float* ndelta, int Npad, int out_f32, cudaStream_t st) {
  const int mma_dtype = FMT ? FA_BF16 : FA_F16;
  CUtensorMap tq, tk, tv, tg;
  int rc;
  if ((rc = make_tmap_public(&tq, q, mma_dtype, g.N, D, g.B))) return rc;
  if ((rc = make_tmap_public(&tk, k, mma_dtype, g.N, D, g.B))) return rc;
  if ((rc = make_tmap_public(&tv, v, mma_dtype, g.N, D, g.B))) return rc;
  if ((rc = make_tmap_public(&tg, d_o, mma_dtype, g.N, D, g.B))) return rc;
  using IT = typename OutT<OBF>::type;      // element type of the caller's tensors
  {
    const dim3 grid((unsigned)((Npad + 255) / 256), (unsigned)g.B);
    bwd_prep_kernel<IT><<<grid, 256, 0, st>>>(static_cast<const IT*>(a.o), static_cast<const IT*>(a.d_o), a.l, a.m,
                                              nlse, ndelta, (int)g.N, Npad, g.dv, amax);
    FA_CUDA_TRY(cudaGetLastError());
  }
  BwdParams prm;
  prm.nlse = nlse; prm.ndelta = ndelta; prm.amax = amax;
  prm.x1 = q; prm.x2 = d_o; prm.out_f32 = out_f32;
  prm.N = (int)g.N; prm.Npad = Npad; prm.W = g.W; prm.p = g.p; prm.mode = g.mode;
  // 2-D periodic neighbourhood: a circulant geometry that carries the image extents (nd = 2, s[0] = X, s[1] = Y)
  const bool td = g.mode == MODE_CIRCULANT && g.nd == 2;
  prm.X = td ? g.s[0] : 0; prm.Y = td ? g.s[1] : 0;
  prm.scale_log2 = g.tau * LOG2E; prm.tau = g.tau;
  prm.dbg = 0;
#ifdef FA_TRACE
  { const char* e = getenv("FA_BWD_DBG"); prm.dbg = e ? atoi(e) : 0; }
#endif
  const dim3 grid(td ? (unsigned)(((g.s[0] + 127) / 128) * g.s[1]) : (unsigned)((g.N + 127) / 128), (unsigned)g.B);
  {
    auto kern = tc_bwd_kernel<D, FMT, 0, OBF, 0>;
    FA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BCfg<D, 0>::SMEM_BYTES));
    prm.out0 = a.dv; prm.out1 = a.dk;
    kern<<<grid, BW_THREADS, BCfg<D, 0>::SMEM_BYTES, st>>>(tk, tv, tq, tg, prm);
    FA_CUDA_TRY(cudaGetLastError());
  }
  prm.out0 = a.dq; prm.out1 = nullptr;
  if (g.mode == MODE_DENSE) {
    auto kern = tc_bwd_kernel<D, FMT, 1, OBF, 1>;
    FA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BCfg<D, 1, 1>::SMEM_BYTES));
    kern<<<grid, BW_THREADS, BCfg<D, 1, 1>::SMEM_BYTES, st>>>(tq, tg, tk, tv, prm);
    FA_CUDA_TRY(cudaGetLastError());
  } else {
    auto kern = tc_bwd_kernel<D, FMT, 1, OBF, 0>;
    FA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BCfg<D, 1, 0>::SMEM_BYTES));
    kern<<<grid, BW_THREADS, BCfg<D, 1, 0>::SMEM_BYTES, st>>>(tq, tg, tk, tv, prm);
    FA_CUDA_TRY(cudaGetLastError());
  }
  return FA_OK;
}

size_t align256b(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// the backward kernels are instantiated for d in {64, 128}; d = 32 (forward: band kernel) takes the exact-fp32 backward
bool tc_bwd_supported(const Geo& g, int dtype) { return (g.d == 64 || g.d == 128) && tc_fwd_supported(g, dtype); }

// workspace: nlse | ndelta | amax[4] | (bf16 inputs, default mode) fp16 re-encodings of q, k, v, dO
size_t tc_bwd_workspace_bytes(const Geo& g, int dtype, int flags) {
  const size_t Npad = ((size_t)g.N + 127) / 128 * 128;
  size_t bytes = 2 * align256b(Npad * (size_t)g.B * sizeof(float)) + 256;
  if (dtype == FA_BF16 && !(flags & FA_FLAG_BF16_INTERNALS)) bytes += 4 * align256b((size_t)g.N * g.d * g.B * 2);
  return bytes;
}

int tc_bwd(const Geo& g, const BwdArgs& a, int dtype, int flags, void* workspace, cudaStream_t st, int out_f32) {
  if (!tc_bwd_supported(g, dtype)) { set_error("tc_bwd: unsupported configuration"); return FA_ERR_UNSUPPORTED; }
  if ((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v) |
       reinterpret_cast<uintptr_t>(a.d_o) | reinterpret_cast<uintptr_t>(workspace)) & 15) {
    set_error("tc_bwd: q/k/v/dO/workspace must be 16-byte aligned"); return FA_ERR_INVALID;
  }
  const int Npad = (int)(((size_t)g.N + 127) / 128 * 128);
  char* ws = static_cast<char*>(workspace);
  const size_t stat_bytes = align256b((size_t)Npad * g.B * sizeof(float));
  float* nlse = reinterpret_cast<float*>(ws);
  float* ndelta = reinterpret_cast<float*>(ws + stat_bytes);
  float* amax = reinterpret_cast<float*>(ws + 2 * stat_bytes);
  const bool reencode = dtype == FA_BF16 && !(flags & FA_FLAG_BF16_INTERNALS);
  const void *q = a.q, *k = a.k, *v = a.v, *d_o = a.d_o;
  if (reencode) {
    // bf16 inputs: P and dS rounded to bf16 (8 significand bits) put the gradients at 2-3e-3 of their
    // max, above the 2e-3 parity tolerance; re-encoding the inputs as scaled fp16 is exact and lets
    // the same kernels carry P and dS with 11 bits.
    const size_t n = (size_t)g.N * g.d * g.B, tb = align256b(n * 2);
    __half* c[4];
    for (int i = 0; i < 4; ++i) c[i] = reinterpret_cast<__half*>(ws + 2 * stat_bytes + 256 + i * tb);
    FA_CUDA_TRY(cudaMemsetAsync(amax, 0, 16, st));
    const unsigned blocks = (unsigned)std::min<size_t>((n / 8 + 255) / 256, 148 * 16);
    const __nv_bfloat16 *xq = static_cast<const __nv_bfloat16*>(a.q), *xk = static_cast<const __nv_bfloat16*>(a.k),
                        *xv = static_cast<const __nv_bfloat16*>(a.v), *xg = static_cast<const __nv_bfloat16*>(a.d_o);
    bwd_amax_kernel<<<dim3(blocks, 4), 256, 0, st>>>(xq, xk, xv, xg, n, amax);
    FA_CUDA_TRY(cudaGetLastError());
    bwd_reencode_kernel<<<dim3(blocks, 4), 256, 0, st>>>(xq, xk, xv, xg, c[0], c[1], c[2], c[3], n, amax);
    FA_CUDA_TRY(cudaGetLastError());
    q = c[0]; k = c[1]; v = c[2]; d_o = c[3];
  }
  const float* am = reencode ? amax : nullptr;
  const int sel = (g.d == 128 ? 4 : 0) | (dtype == FA_BF16 ? 2 : 0) | ((dtype == FA_BF16 && !reencode) ? 1 : 0);
  switch (sel) {   // D, caller dtype, MMA format
    case 4 | 2 | 1: return launch_tc_bwd<128, 1, 1>(g, a, q, k, v, d_o, am, nlse, ndelta, Npad, out_f32, st);
    case 4 | 2:     return launch_tc_bwd<128, 0, 1>(g, a, q, k, v, d_o, am, nlse, ndelta, Npad, out_f32, st);
    case 4:         return launch_tc_bwd<128, 0, 0>(g, a, q, k, v, d_o, am, nlse, ndelta, Npad, out_f32, st);
    case 2 | 1:     return launch_tc_bwd<64, 1, 1>(g, a, q, k, v, d_o, am, nlse, ndelta, Npad, out_f32, st);
    case 2:         return launch_tc_bwd<64, 0, 1>(g, a, q, k, v, d_o, am, nlse, ndelta, Npad, out_f32, st);
    default:        return launch_tc_bwd<64, 0, 0>(g, a, q, k, v, d_o, am, nlse, ndelta, Npad, out_f32, st);
  }
}

}  // namespace fa
