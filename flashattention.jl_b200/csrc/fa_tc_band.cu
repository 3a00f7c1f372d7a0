// fa_tc_band.cu -- compact tcgen05 / TMEM / TMA forward for SHORT key loops: circulant_fa! (reference
// src/circulant.jl:9-118), described for d == dv == 64, where a 128-query tile meets only (128 + W) / 64 key tiles.
//
// Why a second forward kernel: the pair kernel of fa_tc_fwd.cu is built for long key loops (S double
// buffered, QK two steps ahead, 200-register softmax threads) and fits 2 CTAs = 2 query tiles per SM.  With
// 6-7 key tiles per query tile the time goes to latencies -- prologue (barrier init, TMEM alloc, first TMA
// round trip), mbarrier wake-ups between the roles, epilogue -- and not to any pipe (ncu, config 4: tensor
// 16 %, MUFU 27 %, issue 40 %: profiles/r1s_ncu_summary.md).  The cure is more query tiles in flight per SM,
// so this kernel is sized for FOUR (FA_BAND_CTAS=3: three) CTAs per SM:
//   * one S buffer (64 TMEM columns; P aliases its first 32) + O (64 columns) = 128 columns per CTA;
//   * 2-deep (3-deep) K and V rings + the Q tile = 48 (64) KB of shared memory per CTA;
//   * 256 threads at 64 (80) launch registers: warps 0-3 (TMA producer, MMA issuer, TMEM allocator, idle) drop to
//     32 (40), the four softmax warps (thread == query row) rise to 96 (120).  At 96 registers a softmax thread
//     holds one 32-column chunk of its S row at a time and reads S twice (row max, then exponentials).
//   * each 32-column chunk is classified per warp against the band (inside / outside / mixed): only mixed chunks
//     pay the per-element select, chunks outside the band of all 32 rows are neither read nor exponentiated.
// Measured at config 4 (N = 16384, W = 255, B = 512, bf16; same box): two-CTA pair kernel 2.317 ms, this kernel
// with three CTAs 1.811 ms, with four 1.645 ms; one-sided masks written back once 1.512 ms; chunk 0 kept in
// registers between the passes and PV(j) + QK(j+1) handed to the pipe by one elected lane 1.425 ms
// (profiles/r1t_band_kernel.md).
// Per step the CTA is serial (QK(j) -> softmax(j) -> PV(j) -> QK(j+1)); the other CTAs of the SM fill the gaps.
// Layout, descriptors and masking are those of fa_tc_fwd.cu (token-contiguous [B][d][N], SWIZZLE_128B boxes of
// 64 tokens x d channels, MN-major Q/K for S = Q K^T, K-major V for O = P V).
//
// The same kernel also serves (template parameters TD, D, EMU; all measured in profiles/r1t_band_kernel.md):
//   * d = 128 (D): two CTAs per SM, S 64 + O 128 = 192 TMEM columns, one-pass softmax at 192 registers -- circulant
//     d = 128: 1.34 vs 2.19 ms for the pair kernel;
//   * the 2-D periodic neighbourhood of fa_circulant2d_fwd (TD = 1): a query tile is up to 128 consecutive x of one
//     image row, the steps are the W key rows x the 64-key tiles of a row that meet the x band, the mask
//     (c - a) mod X < W is one or two column intervals;
//   * dense attention at d = 64 (TD = 2): every key tile once, the last one masked at N; MUFU-bound there, so a
//     quarter of the exponentials go to the FMA pipe (EMU = 1): 2.60 vs 3.28 ms at N = 8192, B = 128.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "fa_common.cuh"
#include "fa_ptx.cuh"

namespace fa {

int make_tmap_public(CUtensorMap* tm, const void* base, int dtype, long long N, int D, long long B, long long stride_c = 0, long long stride_b = 0);

namespace {

using namespace ptx;

constexpr int BN = 64;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float RESCALE_THRESHOLD = 8.0f;

// D = 64: CTAS = 4 (default) or 3.  D = 128 (2-D neighbourhood only): CTAS = 2 -- S 64 + O 128 = 192 -> 256 TMEM columns,
// Q 32 KB + 2-deep K/V rings of 16 KB tiles = 96 KB, 128 launch registers (misc 64, softmax 192, one-pass softmax).
template <int CTAS, int D> struct BandCfg {
  static constexpr int THREADS = 256, CTAS_PER_SM = CTAS;
  static constexpr int STAGES = CTAS == 3 ? 3 : 2;
  static constexpr int BOX_BYTES = 64 * D * 2, QTILE_BYTES = 2 * BOX_BYTES;
  static constexpr int OFF_Q = 0, OFF_K = QTILE_BYTES, OFF_V = OFF_K + STAGES * BOX_BYTES, OFF_BAR = OFF_V + STAGES * BOX_BYTES;
  static constexpr int BAR_QFULL = 0, BAR_KFULL = 1, BAR_KEMPTY = BAR_KFULL + STAGES, BAR_VFULL = BAR_KEMPTY + STAGES,
                       BAR_VEMPTY = BAR_VFULL + STAGES, BAR_SFULL = BAR_VEMPTY + STAGES, BAR_PFULL = BAR_SFULL + 1,
                       BAR_OFINAL = BAR_PFULL + 1, NUM_BARS = BAR_OFINAL + 1;
  static constexpr int OFF_TMEM_SLOT = OFF_BAR + NUM_BARS * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM_SLOT + 16 + 1024;
  static constexpr int TMEM_COLS = D <= 64 ? 128 : 256, COL_S = 0, COL_O = 64;
  static_assert(CTAS_PER_SM * (SMEM_BYTES + 1024) <= 228 * 1024, "shared memory for CTAS CTAs per SM");
  static_assert(CTAS_PER_SM * TMEM_COLS <= 512, "TMEM");
};

struct BandParams {
  void* o;
  float *l, *m;
  int N, W, p;
  int X, Y;              // TD (2-D periodic neighbourhood): image extents, N = X * Y, tokens x-fastest
  float scale_log2;
  long long* trace;      // FA_TRACE builds: one CTA in the middle of the grid records clock64() per event
};

#ifdef FA_TRACE
// slot = role * 128 + step * 8 + event; role 0 = issuer, 1 = softmax warp 4; step 15 = per-CTA events
#define BTRACE(role, step, ev)                                                                              \
  do {                                                                                                      \
    if (prm.trace && blockIdx.x == gridDim.x / 2 && blockIdx.y == gridDim.y / 2 && lane == 0 && (step) < 16) \
      prm.trace[(role) * 128 + (step) * 8 + (ev)] = clock64();                                               \
  } while (0)
#define BTRACE_E(role, step, ev)                                                                            \
  do {                                                                                                      \
    if (prm.trace && blockIdx.x == gridDim.x / 2 && blockIdx.y == gridDim.y / 2 && (step) < 16)              \
      prm.trace[(role) * 128 + (step) * 8 + (ev)] = clock64();                                               \
  } while (0)
#else
#define BTRACE(role, step, ev) do {} while (0)
#define BTRACE_E(role, step, ev) do {} while (0)
#endif

// FA_BAND_SPIN: poll the two hand-off barriers of the per-step chain (S ready -> softmax, P ready -> issuer) with
// the non-suspending test_wait instead of try_wait (which may park the thread for a system-dependent time)
// EMU (template): column pairs (of the 4 per block of 8) whose exponentials are evaluated on the FMA pipe instead of
// MUFU.  Dense d = 64 is MUFU-bound (512 clk of ex2 vs 492 clk of MMA per 128 x 64 tile): EMU = 1 gives 2.74 -> 2.63 ms at
// N = 8192, B = 128; EMU = 2 is slower (2.87); the latency-bound band configurations do not gain (1.426 vs 1.4215 ms).
#ifndef FA_BAND_SPIN
#define FA_BAND_SPIN 0
#endif
__device__ __forceinline__ void chain_wait(uint32_t bar, uint32_t parity) {
#if FA_BAND_SPIN
  if (mbar_test_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_test_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) { printf("fa_sm100a: band kernel hand-off timeout\n"); __trap(); }
  }
#else
  mbar_wait(bar, parity);
#endif
}

__host__ __device__ inline int fdiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

template <int FMT>
__device__ __forceinline__ uint32_t pack16(float a, float b) {
  if (FMT == 1) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// TD: 0 = 1-D periodic band (circulant_fa!), 1 = 2-D periodic neighbourhood, 2 = dense (every key tile once)
template <int FMT, int CTAS, int TD, int D, int EMU>
__global__ void __launch_bounds__(BandCfg<CTAS, D>::THREADS, BandCfg<CTAS, D>::CTAS_PER_SM)
tc_band_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
               const __grid_constant__ CUtensorMap tmv, const BandParams prm) {
  using C = BandCfg<CTAS, D>;
  constexpr bool TWO_PASS = CTAS == 4;   // 96-register softmax threads hold one 32-column chunk at a time
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = sbase + C::OFF_Q, sK = sbase + C::OFF_K, sV = sbase + C::OFF_V;
  auto bar = [&](int i) { return sbase + C::OFF_BAR + 8u * (uint32_t)i; };
  const uint32_t tmem_slot = sbase + C::OFF_TMEM_SLOT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1-D: query tile = 128 consecutive tokens.  TD: query tile = up to 128 consecutive x of ONE image row yq
  // (tokens are x-fastest, X % 64 == 0), blockIdx.x = row * tiles_per_row + tile.
  const int b = blockIdx.y;
  const int tpr = TD == 1 ? (prm.X + 127) / 128 : 1;
  const int yq = TD == 1 ? (int)(blockIdx.x / tpr) : 0;
  const int x0 = TD == 1 ? (int)(blockIdx.x % tpr) * 128 : 0;
  const int tq = TD == 1 ? (prm.X - x0 < 128 ? prm.X - x0 : 128) : 128;
  const int q0 = TD == 1 ? yq * prm.X + x0 : blockIdx.x * 128;
  if (warp == 1) BTRACE(0, 15, 0);                         // kernel entry

  // key tiles of this query tile: keys (q - p) .. (q - p + W - 1) for q in [q0, q0 + 128), from the 64-aligned
  // tile at or below q0 - p (src/circulant.jl:61-67: the window of query j starts p keys before it, periodic)
  // TD: the keys of (x, y) are (mod(x - p + s, X), mod(y - p + t, Y)), s, t in [0, W) -- the direct product of the
  // 1-D key set (src/utils.jl:6-17).  Steps = W key rows x the nx 64-key tiles of a row that meet the x band of the
  // query tile (all X / 64 tiles of the row, once each, when the band wraps round the whole row).
  const int kbase = TD == 2 ? 0 : (TD == 1 ? fdiv(x0 - prm.p, BN) * BN : fdiv(q0 - prm.p, BN) * BN);
  int nx = 1, kxb = kbase;
  if (TD == 1) {
    nx = fdiv(x0 + tq - 1 - prm.p + prm.W - 1 - kbase, BN) + 1;
    if (nx * BN >= prm.X) { nx = prm.X / BN; kxb = 0; }
  }
  const int nj = TD == 2 ? (prm.N + BN - 1) / BN : (TD == 1 ? prm.W * nx : fdiv(q0 + 127 - prm.p + prm.W - 1 - kbase, BN) + 1);
  // token of the first key of step j (TD: row (yq - p + t) mod Y, x tile (kxb + 64 jx) mod X)
  auto key_tile = [&](int j, int& kxs) {
    if (TD != 1) { kxs = 0; return (int)pmod(kbase + BN * j, prm.N); }
    const int t = j / nx, jx = j - t * nx;
    kxs = (int)pmod(kxb + BN * jx, prm.X);
    return (int)pmod(yq - prm.p + t, prm.Y) * prm.X + kxs;
  };

  // The producer initialises its own "full" barriers and sends the first round trip (Q, K(0), V(0)) on its way BEFORE the
  // CTA-wide set-up barrier: TMEM allocation and the other barrier inits overlap the ~1800-clk TMA latency.
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmq); prefetch_tensormap(&tmk); prefetch_tensormap(&tmv);
    mbar_init(bar(C::BAR_QFULL), 1);
    for (int i = 0; i < C::STAGES; ++i) { mbar_init(bar(C::BAR_KFULL + i), 1); mbar_init(bar(C::BAR_VFULL + i), 1); }
    fence_barrier_init();
    mbar_arrive_expect_tx(bar(C::BAR_QFULL), C::QTILE_BYTES);
    tma_load_3d(sQ, &tmq, bar(C::BAR_QFULL), q0, 0, b);
    tma_load_3d(sQ + C::BOX_BYTES, &tmq, bar(C::BAR_QFULL), q0 + 64, 0, b);
    int kxs0;
    const int tok0 = key_tile(0, kxs0);
    mbar_arrive_expect_tx(bar(C::BAR_KFULL), C::BOX_BYTES);
    tma_load_3d(sK, &tmk, bar(C::BAR_KFULL), tok0, 0, b);
    mbar_arrive_expect_tx(bar(C::BAR_VFULL), C::BOX_BYTES);
    tma_load_3d(sV, &tmv, bar(C::BAR_VFULL), tok0, 0, b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) { mbar_init(bar(C::BAR_KEMPTY + i), 1); mbar_init(bar(C::BAR_VEMPTY + i), 1); }
    mbar_init(bar(C::BAR_SFULL), 1); mbar_init(bar(C::BAR_PFULL), 128); mbar_init(bar(C::BAR_OFINAL), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (warp == 1) BTRACE(0, 15, 1);                         // set-up done

  if (warp < 4) {
    if (CTAS == 4) setmaxnreg_dec<32>(); else if (CTAS == 3) setmaxnreg_dec<40>(); else setmaxnreg_dec<64>();
    if (warp == 0 && lane == 0) {
      // ------------------------------------------------------------ TMA producer
      for (int j = 1; j < nj; ++j) {                       // step 0 went out before the set-up barrier
        const int s = j % C::STAGES;
        const uint32_t par = ((uint32_t)(j / C::STAGES) & 1u) ^ 1u;
        int kxs_unused;
        const int tok = key_tile(j, kxs_unused);
        mbar_wait(bar(C::BAR_KEMPTY + s), par);
        mbar_arrive_expect_tx(bar(C::BAR_KFULL + s), C::BOX_BYTES);
        tma_load_3d(sK + s * C::BOX_BYTES, &tmk, bar(C::BAR_KFULL + s), tok, 0, b);
        mbar_wait(bar(C::BAR_VEMPTY + s), par);
        mbar_arrive_expect_tx(bar(C::BAR_VFULL + s), C::BOX_BYTES);
        tma_load_3d(sV + s * C::BOX_BYTES, &tmv, bar(C::BAR_VFULL + s), tok, 0, b);
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------ MMA issuer (warp-uniform loop, one elected lane issues)
      constexpr uint32_t idesc_qk = make_idesc_f16(FMT, FMT, 1, 1, 128, BN);
      constexpr uint32_t idesc_pv = make_idesc_f16(FMT, FMT, 0, 0, 128, D);
      const uint64_t qdesc = make_smem_desc_sw128(sQ, C::BOX_BYTES, 1024);
      const uint64_t kdesc = make_smem_desc_sw128(sK, C::BOX_BYTES, 1024);
      const uint64_t vdesc = make_smem_desc_sw128(sV, 16, 1024);
      const uint32_t tS = tmem_base + C::COL_S, tO = tmem_base + C::COL_O;
      mbar_wait(bar(C::BAR_QFULL), 0);
      BTRACE(0, 15, 2);                                     // Q landed
      // S(0) = Q K(0)^T
      mbar_wait(bar(C::BAR_KFULL), 0);
      BTRACE(0, 0, 0);                                      // K(0) ready
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks)
          mma_ss(tS, qdesc + (uint64_t)(ks * 128), kdesc + (uint64_t)(ks * 128), idesc_qk, ks > 0 ? 1u : 0u);
        BTRACE_E(0, 0, 4);
        tc_commit(bar(C::BAR_SFULL));
        tc_commit(bar(C::BAR_KEMPTY));
      }
      __syncwarp();
      BTRACE(0, 0, 1);                                      // QK(0) issued
      for (int j = 0; j < nj; ++j) {
        const int s = j % C::STAGES, s1 = (j + 1) % C::STAGES;
        const uint32_t par = (uint32_t)(j / C::STAGES) & 1u, par1 = (uint32_t)((j + 1) / C::STAGES) & 1u;
        const bool more = j + 1 < nj;
        // operands of this step's PV and of the next step's QK arrive long before P(j): wait for them first, so
        // that once P(j) is published one elected lane hands PV(j) AND QK(j+1) to the pipe back to back.
        // (QK(j+1) overwrites the S buffer PV(j) reads P from: the tensor pipe runs in order.)
        mbar_wait(bar(C::BAR_VFULL + s), par);
        if (more) mbar_wait(bar(C::BAR_KFULL + s1), par1);
        BTRACE(0, j, 5);                                    // V(j), K(j+1) ready
        chain_wait(bar(C::BAR_PFULL), (uint32_t)j & 1u);
        BTRACE(0, j, 2);                                    // P(j) seen
        tc_fence_after();
        if (elect_one()) {
          const uint64_t vd = vdesc + (uint64_t)(s * (C::BOX_BYTES >> 4));
#pragma unroll
          for (int ks = 0; ks < BN / 16; ++ks)
            mma_ts(tO, tS + ks * 8, vd + (uint64_t)(ks * 2), idesc_pv, (j > 0 || ks > 0) ? 1u : 0u);
          BTRACE_E(0, j, 6);                                // PV MMAs handed to the pipe
          tc_commit(bar(C::BAR_VEMPTY + s));
          if (more) {
            const uint64_t kd = kdesc + (uint64_t)(s1 * (C::BOX_BYTES >> 4));
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks)
              mma_ss(tS, qdesc + (uint64_t)(ks * 128), kd + (uint64_t)(ks * 128), idesc_qk, ks > 0 ? 1u : 0u);
            BTRACE_E(0, j + 1, 4);                          // QK(j+1) MMAs handed to the pipe
            tc_commit(bar(C::BAR_SFULL));        // also: every earlier MMA (PV(j)) has completed
            tc_commit(bar(C::BAR_KEMPTY + s1));
          }
        }
        __syncwarp();
        BTRACE(0, j, 3);                                    // PV(j) (+ QK(j+1)) issued
      }
      if (elect_one()) tc_commit(bar(C::BAR_OFINAL));
      __syncwarp();
    }
  } else {
    // -------------------------------------------------------------- softmax: thread == query row
    if (CTAS == 4) setmaxnreg_inc<96>(); else if (CTAS == 3) setmaxnreg_inc<120>(); else setmaxnreg_inc<192>();
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + C::COL_S, tO = tmem_base + lane_addr + C::COL_O;
    const int qi = q0 + row;
    const float scale = prm.scale_log2;
    const float2 scale2 = make_float2(scale, scale);
    float m_true = -INFINITY, m_used = -INFINITY;
    float2 l2a = make_float2(0.f, 0.f), l2b = make_float2(0.f, 0.f);
    const int lo0 = (qi - prm.p) - kbase;                       // 1-D: first in-band column of key tile 0 for this row
    const int qx = x0 + row;                                    // TD: image column of this row's query

    // x16 TMEM load (rescale of O in 16-column pieces keeps the register peak low)
    auto tmem_ld16 = [](uint32_t taddr, uint32_t (&o)[16]) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]),
            "=r"(o[8]), "=r"(o[9]), "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15])
          : "r"(taddr) : "memory");
    };
    // mask of one 32-column chunk held in registers: columns outside [lo, hi) (chunk-relative) become -inf
    // (W >= 32 means a chunk meets at most one band edge for most warps: the one-sided forms save a compare per
    // element; `side` is warp-uniform: 1 = left edge only, 2 = right edge only, 0 = both)
    auto mask32 = [](uint32_t (&sc)[32], int lo, int hi) {
      const uint32_t side = __all_sync(0xffffffffu, hi >= 32) ? 1u : (__all_sync(0xffffffffu, lo <= 0) ? 2u : 0u);
      if (side == 1u) {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (e < lo) sc[e] = 0xff800000u;
      } else if (side == 2u) {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (e >= hi) sc[e] = 0xff800000u;
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (e < lo || e >= hi) sc[e] = 0xff800000u;
      }
    };
    // TD, band wrapping round the image row inside one key tile: a second interval [lo2, hi2)
    auto mask32_2 = [](uint32_t (&sc)[32], int lo, int hi, int lo2, int hi2) {
#pragma unroll
      for (int e = 0; e < 32; ++e)
        if (!((e >= lo && e < hi) || (e >= lo2 && e < hi2))) sc[e] = 0xff800000u;
    };
    auto max32 = [](const uint32_t (&sc)[32]) {
      float a = -INFINITY, bq = -INFINITY;
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        a = fmaxf(a, fmaxf(__uint_as_float(sc[e]), __uint_as_float(sc[e + 1])));
        bq = fmaxf(bq, fmaxf(__uint_as_float(sc[e + 2]), __uint_as_float(sc[e + 3])));
      }
      return fmaxf(a, bq);
    };
    // P chunk = exp2(s * scale - m) -> 16 bit (blocks of 8: 4 FFMA2, 8 ex2, 4 FADD2 + 4 packs), stored over S
    auto exp32 = [&](const uint32_t (&sc)[32], float2 negm2, uint32_t dst) {
      uint32_t pk[16];
#pragma unroll
      for (int e0 = 0; e0 < 32; e0 += 8) {
        float2 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          x[u] = __ffma2_rn(make_float2(__uint_as_float(sc[e0 + 2 * u]), __uint_as_float(sc[e0 + 2 * u + 1])), scale2, negm2);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (u >= 4 - EMU) {
            // exponential on the FMA pipe (Cody-Waite split + cubic, rel. error 9e-5 << 16-bit rounding of P)
            float2 t = x[u];
            t.x = fmaxf(t.x, -126.f); t.y = fmaxf(t.y, -126.f);
            const float2 xf = __fadd2_rd(t, make_float2(12582912.f, 12582912.f));
            const float2 xr = __fadd2_rn(xf, make_float2(-12582912.f, -12582912.f));
            const float2 fr = __fadd2_rn(t, make_float2(-xr.x, -xr.y));
            float2 q = __ffma2_rn(fr, make_float2(0.077119089663028717f, 0.077119089663028717f),
                                  make_float2(0.227564394474029541f, 0.227564394474029541f));
            q = __ffma2_rn(q, fr, make_float2(0.695146143436431885f, 0.695146143436431885f));
            q = __ffma2_rn(q, fr, make_float2(1.f, 1.f));
            x[u].x = __uint_as_float(__float_as_uint(q.x) + (__float_as_uint(xf.x) << 23));
            x[u].y = __uint_as_float(__float_as_uint(q.y) + (__float_as_uint(xf.y) << 23));
          } else {
            x[u].x = ex2(x[u].x); x[u].y = ex2(x[u].y);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (u & 1) l2b = __fadd2_rn(l2b, x[u]); else l2a = __fadd2_rn(l2a, x[u]);
          pk[(e0 >> 1) + u] = pack16<FMT>(x[u].x, x[u].y);
        }
      }
      tmem_st16(dst, pk);
    };
    auto zero16 = [](uint32_t dst) {
      uint32_t z[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) z[e] = 0u;
      tmem_st16(dst, z);
    };

#pragma unroll 1
    for (int j = 0; j < nj; ++j) {
      // valid columns of this key tile for this row: [lo, hi), and for TD also [lo2, hi2) = the same band one image
      // row length to the left (periodic in x): (c - a) mod X < W with a = (qx - p - tile start) mod X
      int lo, hi, lo2 = 0, hi2 = 0;
      if (TD == 1) {
        int kxs;
        key_tile(j, kxs);
        lo = (int)pmod(qx - prm.p - kxs, prm.X); hi = lo + prm.W;
        lo2 = lo - prm.X; hi2 = lo2 + prm.W;
      } else {
        lo = lo0 - BN * j; hi = lo + prm.W;
        if (TD == 2) { lo = 0; hi = prm.N - BN * j; }
      }
      const bool two = TD == 1 && __any_sync(0xffffffffu, hi2 > 0);       // warp-uniform: some lane's band wraps into this tile
      // Each 32-column chunk is classified per warp (the band edge is a diagonal: it crosses ~32 columns over a
      // warp's 32 rows): 1 = inside the band of every lane (no predicates), 2 = outside for every lane (P = 0: S is
      // not read, no exponentials), 0 = mixed (per-element select).
      uint32_t kc[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        bool in = lo <= 32 * c && 32 * c + 32 <= hi, out = hi <= 32 * c || lo >= 32 * c + 32;
        if (TD == 1) {
          const int c1 = min(hi, 32 * c + 32) - max(lo, 32 * c), c2 = min(hi2, 32 * c + 32) - max(lo2, 32 * c);
          const int cnt = (c1 > 0 ? c1 : 0) + (c2 > 0 ? c2 : 0);
          in = cnt == 32; out = cnt == 0;
        }
        kc[c] = __all_sync(0xffffffffu, in) ? 1u : (__all_sync(0xffffffffu, out) ? 2u : 0u);
      }
      chain_wait(bar(C::BAR_SFULL), (uint32_t)j & 1u);
      if (warp == 4) BTRACE(1, j, 0);                       // S(j) seen
      tc_fence_after();
      if (kc[0] == 2u && kc[1] == 2u) {
        zero16(tS); zero16(tS + 16);
      } else {
        uint32_t s0[32], s1[TWO_PASS ? 1 : 32];
        float mx = -INFINITY;
        if (TWO_PASS) {
          // ---- pass 1: row max, one chunk in registers at a time; chunk 0 last, so that it is still in registers
          // when pass 2 starts with it
#pragma unroll
          for (int c = 1; c >= 0; --c) {
            if (kc[c] != 2u) {
              tmem_ld32(tS + 32 * c, s0);
              tmem_wait_ld();
              if (kc[c] == 0u) {                    // masked once: pass 2 reads the masked chunk back
                if (two) mask32_2(s0, lo - 32 * c, hi - 32 * c, lo2 - 32 * c, hi2 - 32 * c);
                else mask32(s0, lo - 32 * c, hi - 32 * c);
                tmem_st32(tS + 32 * c, s0);
              }
              mx = fmaxf(mx, max32(s0));
            }
          }
          if (kc[0] == 0u || kc[1] == 0u) tmem_wait_st();
        } else {
          if (kc[0] != 2u) tmem_ld32(tS, s0);
          if (kc[1] != 2u) tmem_ld32(tS + 32, (uint32_t(&)[32])s1);
          tmem_wait_ld();
          if (kc[0] == 0u) { if (two) mask32_2(s0, lo, hi, lo2, hi2); else mask32(s0, lo, hi); }
          if (kc[1] == 0u) {
            if (two) mask32_2((uint32_t(&)[32])s1, lo - 32, hi - 32, lo2 - 32, hi2 - 32);
            else mask32((uint32_t(&)[32])s1, lo - 32, hi - 32);
          }
          if (kc[0] != 2u) mx = max32(s0);
          if (kc[1] != 2u) mx = fmaxf(mx, max32((uint32_t(&)[32])s1));
        }
        m_true = fmaxf(m_true, mx * scale);
        if (warp == 4) BTRACE(1, j, 1);                     // max done
        // lazy rescale (warp-uniform decision).  S(j) complete implies PV(j-1) complete: O may be touched.
        const bool want = (m_true - m_used) > RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, want)) {
          const float alpha = (m_used == -INFINITY) ? 0.f : ex2(m_used - m_true);
          if (j > 0) {
#pragma unroll 1
            for (int c = 0; c < D / 16; ++c) {
              uint32_t o[16];
              tmem_ld16(tO + 16 * c, o);
              tmem_wait_ld();
#pragma unroll
              for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              tmem_st16(tO + 16 * c, o);
            }
          }
          l2a.x *= alpha; l2a.y *= alpha; l2b.x *= alpha; l2b.y *= alpha;
          m_used = m_true;
        }
        if (warp == 4) BTRACE(1, j, 2);                     // rescale decided / done
        const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used;
        const float2 negm2 = make_float2(neg_m, neg_m);
        // ---- P over the first 32 columns of S.  Chunk 0 is in registers before its own columns are overwritten,
        // chunk 1 (columns 32-63) is read before P lands in columns 16-31.
        if (TWO_PASS) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (kc[c] == 2u) { zero16(tS + 16 * c); continue; }
            if (c == 1) {                          // chunk 0 is still in registers from pass 1
              tmem_ld32(tS + 32, s0);
              tmem_wait_ld();
            }
            exp32(s0, negm2, tS + 16 * c);
          }
        } else {
          if (kc[0] == 2u) zero16(tS); else exp32(s0, negm2, tS);
          if (kc[1] == 2u) zero16(tS + 16); else exp32((uint32_t(&)[32])s1, negm2, tS + 16);
        }
      }
      if (warp == 4) BTRACE(1, j, 3);                       // exps done, P stores issued
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(bar(C::BAR_PFULL));
      if (warp == 4) BTRACE(1, j, 4);                       // P published
    }
    const float l_run = (l2a.x + l2a.y) + (l2b.x + l2b.y);

    // ---- epilogue: O / l -> global (token-contiguous rows: a warp writes 32 consecutive tokens per channel)
    mbar_wait(bar(C::BAR_OFINAL), 0);
    if (warp == 4) BTRACE(1, 15, 0);                        // all MMAs done
    tc_fence_after();
    const float inv_l = 1.f / l_run;
    const bool in_range = TD == 1 ? row < tq : qi < prm.N;
    unsigned short* ob = static_cast<unsigned short*>(prm.o) + (size_t)b * D * prm.N + qi;
#pragma unroll 1
    for (int c = 0; c < D / 32; ++c) {
      uint32_t o[32];
      tmem_ld32(tO + 32 * c, o);
      tmem_wait_ld();
      if (in_range) {
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const uint32_t pr = pack16<FMT>(__uint_as_float(o[e]) * inv_l, __uint_as_float(o[e + 1]) * inv_l);
          ob[(size_t)(32 * c + e) * prm.N] = (unsigned short)(pr & 0xffffu);
          ob[(size_t)(32 * c + e + 1) * prm.N] = (unsigned short)(pr >> 16);
        }
      }
    }
    if (in_range) {
      // l = sum exp(s - m_true), m = max s (natural-log domain), reference src/circulant.jl:110-116
      prm.l[(size_t)b * prm.N + qi] = l_run * ex2(m_used - m_true);
      prm.m[(size_t)b * prm.N + qi] = m_true * LN2;
    }
    if (warp == 4) BTRACE(1, 15, 1);                        // epilogue stores issued
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int FMT, int CTAS, int TD, int D = 64, int EMU = 0>
int launch_band(const Geo& g, const FwdArgs& a, int dtype, cudaStream_t st, int X = 0, int Y = 0) {
  using C = BandCfg<CTAS, D>;
  CUtensorMap tmq, tmk, tmv;
  int rc;
  if ((rc = make_tmap_public(&tmq, a.q, dtype, g.N, D, g.B, a.in_stride_c, a.in_stride_b))) return rc;
  if ((rc = make_tmap_public(&tmk, a.k, dtype, g.N, D, g.B, a.in_stride_c, a.in_stride_b))) return rc;
  if ((rc = make_tmap_public(&tmv, a.v, dtype, g.N, D, g.B, a.in_stride_c, a.in_stride_b))) return rc;
  BandParams prm;
  prm.o = a.o; prm.l = a.l; prm.m = a.m;
  prm.N = (int)g.N; prm.W = g.W; prm.p = g.p;
  prm.X = X; prm.Y = Y;
  prm.scale_log2 = g.tau * LOG2E;
  prm.trace = nullptr;
#ifdef FA_TRACE
  { const char* e = getenv("FA_TRACE_PTR"); prm.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr; }
#endif
  auto kern = tc_band_kernel<FMT, CTAS, TD, D, EMU>;
  FA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  const dim3 grid(TD == 1 ? (unsigned)(((X + 127) / 128) * Y) : (unsigned)((g.N + 127) / 128), (unsigned)g.B);
  kern<<<grid, C::THREADS, C::SMEM_BYTES, st>>>(tmq, tmk, tmv, prm);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

}  // namespace

// circulant, d = dv in {64, 128}, 16-bit output, tile-aligned wrap-around (N % 64 == 0) -- checked by the caller (tc_fwd)
int tc_band_fwd(const Geo& g, const FwdArgs& a, int dtype, cudaStream_t st) {
  static const int ctas = [] { const char* e = getenv("FA_BAND_CTAS"); return e ? atoi(e) : 4; }();
  if (g.d == 32) {      // the reference's benchmark head dim (logs/circ_t*.txt, logs/wind_t*.txt): same kernel, two K steps per QK
    if (g.mode == MODE_DENSE) return dtype == FA_BF16 ? launch_band<1, 4, 2, 32, 1>(g, a, dtype, st) : launch_band<0, 4, 2, 32, 1>(g, a, dtype, st);
    return dtype == FA_BF16 ? launch_band<1, 4, 0, 32>(g, a, dtype, st) : launch_band<0, 4, 0, 32>(g, a, dtype, st);
  }
  if (g.d == 128) {
    if (g.mode == MODE_DENSE) return dtype == FA_BF16 ? launch_band<1, 2, 2, 128>(g, a, dtype, st) : launch_band<0, 2, 2, 128>(g, a, dtype, st);
    return dtype == FA_BF16 ? launch_band<1, 2, 0, 128>(g, a, dtype, st) : launch_band<0, 2, 0, 128>(g, a, dtype, st);
  }
  // dense (TD = 2): every key tile once, no band, the last tile masked at N; a quarter of the exponentials on the FMA pipe
  if (g.mode == MODE_DENSE) {
    if (ctas == 3) return dtype == FA_BF16 ? launch_band<1, 3, 2, 64, 1>(g, a, dtype, st) : launch_band<0, 3, 2, 64, 1>(g, a, dtype, st);
    return dtype == FA_BF16 ? launch_band<1, 4, 2, 64, 1>(g, a, dtype, st) : launch_band<0, 4, 2, 64, 1>(g, a, dtype, st);
  }
  if (ctas != 3) return dtype == FA_BF16 ? launch_band<1, 4, 0>(g, a, dtype, st) : launch_band<0, 4, 0>(g, a, dtype, st);
  return dtype == FA_BF16 ? launch_band<1, 3, 0>(g, a, dtype, st) : launch_band<0, 3, 0>(g, a, dtype, st);
}

// 2-D periodic neighbourhood attention on (X, Y, 64, B), 16-bit: the same kernel walking W key rows
bool tc_band2d_supported(long long X, long long Y, long long d, long long dv, long long B, long long W, int dtype) {
  if (dtype != FA_BF16 && dtype != FA_F16) return false;
  if (d != dv || (d != 64 && d != 128) || X % 64 != 0 || X <= 0 || Y <= 0 || W <= 0 || W > X || W > Y || W > 64) return false;
  if (B > 65535 || X * Y > 0x3fffffffLL || ((X + 127) / 128) * Y > 0x7fffffffLL) return false;
  return true;
}

int tc_band2d_fwd(const void* q, const void* k, const void* v, void* o, float* l, float* m,
                  long long X, long long Y, long long d, long long B, long long W, int dtype, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) {
    set_error("tc_band2d_fwd: q/k/v must be 16-byte aligned"); return FA_ERR_INVALID;
  }
  Geo g;
  memset(&g, 0, sizeof(g));
  g.mode = MODE_CIRCULANT; g.d = (int)d; g.dv = (int)d; g.N = X * Y; g.B = B; g.W = (int)W; g.p = (int)((W - 1) / 2);
  g.tau = 1.0f / sqrtf((float)d);
  FwdArgs a{q, k, v, o, nullptr, l, m};
  if (d == 128)
    return dtype == FA_BF16 ? launch_band<1, 2, 1, 128>(g, a, dtype, st, (int)X, (int)Y) : launch_band<0, 2, 1, 128>(g, a, dtype, st, (int)X, (int)Y);
  return dtype == FA_BF16 ? launch_band<1, 4, 1>(g, a, dtype, st, (int)X, (int)Y) : launch_band<0, 4, 1>(g, a, dtype, st, (int)X, (int)Y);
}

}  // namespace fa
