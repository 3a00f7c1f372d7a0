// fa_tc_fwd.cu -- tcgen05 / TMEM / TMA flash-attention forward for sm_100a.
//
// Replaces the hot loop of dense_fa! (reference src/dense.jl:70-92) and circulant_fa!
// (src/circulant.jl:61-108) for 16-bit inputs with d == dv in {64, 128}.  Written from scratch
// for Blackwell; nothing here derives from src/cuda/flash.jl or src_cpp/FlashAttention.cu.
//
// Data layout (reference src/dense.jl:6-8): q,k,v,o are [B][d][N] with the TOKEN index
// contiguous.  A TMA box of 64 tokens x d channels therefore lands in shared memory as d rows
// of 128 B (SWIZZLE_128B), which is exactly
//   * the MN-major SW128 canonical UMMA layout for Q and K in S = Q K^T
//       (MN = token contiguous, K = channel rows at 128 B pitch, SBO = 1024 B per 8 channels,
//        LBO = one box = 128*d B per 64 tokens), and
//   * the K-major SW128 canonical layout for V in O = P V
//       (rows = channel, K = key contiguous; SBO = 1024 B per 8 channels).
// No transposes, no staging copies: the transposed Julia layout maps 1:1 onto UMMA descriptors.
//
// CTA = 384 threads, one CTA per SM (512 TMEM columns), a pair of 128-query tiles t = 0,1:
//   warp 0      TMA producer (Q pair once; 64-key K and V tiles through two mbarrier rings)
//   warps 1, 3  MMA issuers  (one thread per Q tile t; the tiles are independent, and one thread
//               cannot issue 24 small MMAs per 1024-cycle step): at global step g, for its t:
//                   O_t += P_t(g-2) V(g-2)      (A = P from TMEM, B = V K-major from smem)
//                   S_t[g&1] = Q_t K(g)^T       (A, B MN-major from smem, N = 64)
//               i.e. QK runs TWO steps ahead of PV into a double-buffered S, so the softmax
//               warps never wait for the tensor pipe and the tensor pipe never waits for more
//               than one softmax (v1 had one S buffer per tile: softmax(j) and [PV(j), QK(j+1)]
//               of a tile were serialized, tensor pipe 53 % -- profiles/r1a_*).
//   warp 2      TMEM allocator
//   warps 4-7   softmax for Q tile 0   (thread == row; S row read from TMEM with tcgen05.ld,
//   warps 8-11  softmax for Q tile 1    running max/sum in registers, exp2, P written back to
//                                       TMEM as 16-bit over its own S buffer, lazy O rescale,
//                                       final O/l, m epilogue)
// TMEM columns: S_t[b] at 128 t + 64 b (64 fp32 columns each), O_t at 256 + d t; P_t(j) aliases
// the first 32 columns of S_t[j&1].
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include "fa_common.cuh"
#include "fa_ptx.cuh"

namespace fa {
namespace {

using namespace ptx;

constexpr int TC_THREADS = 384;
constexpr int BN = 64;                      // keys per K/V tile (one TMA box)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float RESCALE_THRESHOLD = 8.0f;   // lazy rescale: P may grow to 2^8 before O is rescaled
// Tuning knobs, measured on B200 at the C3 shape over 40 back-to-back steps (power-capped, ~1650 MHz;
// profiles/r1b_ab_variants.txt): packed f32x2 math +2 %; the TMEM prefetch of S(j+1) -2 %; emulating
// 25 % of the exponentials on the FMA pipe -5 % (more instructions -> more power -> lower clocks).
#ifndef FA_PREFETCH
#define FA_PREFETCH 0
#endif
#ifndef FA_PACKED
#define FA_PACKED 1
#endif
// Clock traces (tools/trace_fwd.py, profiles/r1o_fwd_step_trace.md) show the softmax warps idle ~390 clk
// per step between publishing P(j) and holding S(j+1) in registers (mbarrier wake-up + tcgen05.ld), although
// S(j+1) is complete by the middle of the exp phase of step j: fetch it there, between the two 32-column
// halves, when the barrier already flipped.
#ifndef FA_PREFETCH_MID
#define FA_PREFETCH_MID 1
#endif
#ifndef FA_EMU_PAIRS
#define FA_EMU_PAIRS 0      // column pairs per exp block whose exponentials are evaluated on the FMA pipe
#endif
#ifndef FA_EXP_SWP
#define FA_EXP_SWP 0
#endif
#ifndef FA_EXP_BLOCK
#define FA_EXP_BLOCK 8      // elements per exp block (0 = pairwise loop); 8: +1.5 % over pairwise, same-box A/B
#endif
#ifndef FA_EMU_PER_32
#define FA_EMU_PER_32 0
#endif
constexpr int EMU_PER_32 = FA_EMU_PER_32;   // exponentials per 32 emulated on the FMA pipe (8 = 25 %)

// NQT = 128-query tiles per CTA.  2: one CTA per SM, 384 threads, all 512 TMEM columns -- the
// tensor-bound dense configuration.  1: 256 threads, 256 TMEM columns and ~100 KB of smem, so TWO CTAs
// share an SM and overlap each other's prologue (Q/K/V fetch latency) and epilogue (O store) -- the
// short-loop circulant configuration, where a CTA lives for only (128 + W) / 64 key tiles.
// SPLIT = softmax threads per query row.  2: two warpgroups per Q tile, each thread owns 32 of the 64
// columns of a step (row max exchanged through shared memory + a 64-thread named barrier, row sums kept
// per thread): four softmax warps per scheduler instead of two hide the TMEM / MUFU / barrier latencies
// of the serial softmax chain better.
template <int D, int NQT = 2, int SPLIT = 1>
struct Cfg {
  // SPLIT == 3: one softmax warpgroup per Q tile plus one HELPER warpgroup per Q tile that computes the
  // row maxima of S(j+1) off the softmax warps' serial chain (and publishes them through shared memory)
  static constexpr bool HELP = (SPLIT == 3);
  static constexpr int THREADS = 128 + 128 * NQT * (HELP ? 2 : SPLIT);
  static constexpr int K_STAGES = (NQT == 1) ? ((D == 128) ? 2 : 4) : ((D == 128) ? 3 : 4);
  static constexpr int V_STAGES = (NQT == 1) ? ((D == 128) ? 3 : 6) : ((D == 128) ? 4 : 6);
  static constexpr int BOX_BYTES = 64 * D * 2;          // 64 tokens x D channels, 16-bit
  static constexpr int QTILE_BYTES = 2 * BOX_BYTES;     // 128 queries
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = NQT * QTILE_BYTES;
  static constexpr int OFF_V = OFF_K + K_STAGES * BOX_BYTES;
  static constexpr int OFF_BAR = OFF_V + V_STAGES * BOX_BYTES;
  // barrier slots (8 B each)
  static constexpr int BAR_QFULL = 0;                          // [2]
  static constexpr int BAR_KFULL = 2;                          // [K_STAGES]
  static constexpr int BAR_KEMPTY = BAR_KFULL + K_STAGES;
  static constexpr int BAR_VFULL = BAR_KEMPTY + K_STAGES;      // [V_STAGES]
  static constexpr int BAR_VEMPTY = BAR_VFULL + V_STAGES;
  static constexpr int BAR_SFULL = BAR_VEMPTY + V_STAGES;      // [2 tiles][2 buffers]
  static constexpr int BAR_PFULL = BAR_SFULL + 4;              // [2][2]
  static constexpr int BAR_ODONE = BAR_PFULL + 4;              // [2]  one completion per PV_t
  static constexpr int BAR_OFINAL = BAR_ODONE + 2;             // [2]  single use: all MMAs of tile t done
  static constexpr int BAR_MAX = BAR_OFINAL + 2;               // [2][2] HELP: row maxima of S_t[b] published
  static constexpr int BAR_MXX = BAR_MAX + 4;                  // [2 tiles][4 quarters][2 halves][2 parities] SPLIT = 2
  static constexpr int NUM_BARS = BAR_MXX + (SPLIT == 2 ? 32 : 0);
  static constexpr int OFF_TMEM_SLOT = OFF_BAR + NUM_BARS * 8;
  static constexpr int OFF_XCH = OFF_TMEM_SLOT + 16;             // SPLIT = 2: float[NQT][3][2][128] max / sum exchange
  static constexpr int SMEM_BYTES = OFF_XCH + (SPLIT >= 2 ? NQT * 3 * 2 * 128 * 4 : 0) + 1024;   // + alignment slack
  static constexpr int TMEM_COLS = (NQT == 2) ? 512 : 256;
  static constexpr int COL_S = 0, COL_O = 128 * NQT;             // S_t[b] = 128 t + 64 b; O_t = COL_O + D t
  static constexpr int CTAS_PER_SM = (NQT == 2) ? 1 : 2;
};

struct TcParams {
  void* o;
  float* l;
  float* m;
  int N, B, W, p, mode;
  float scale_log2;      // tau * log2(e)
  int o_f32;             // store O as float32
  int stagger;           // clocks by which the softmax warpgroup of Q tile 1 starts late (experiment)
  int pingpong;          // the two softmax warpgroups take turns on the MUFU-heavy exp phase (named barriers): 1 = on the
                         // whole phase, 2 = on its first half only (the second half overlaps the other tile's first half)
  long long* trace;      // FA_TRACE builds: CTA (0,0) records (clock) per (role, step, event); else NULL
};

#ifdef FA_TRACE
// role: 0/1 = issuer of tile t, 2/3 = softmax warp 0 of tile t.  8 events per step, 64 steps.
#define TRACE(role, step, ev)                                                                         \
  do {                                                                                                \
    if (prm.trace && blockIdx.x == 1 && blockIdx.y == 0 && lane == 0 && (step) >= 32 && (step) < 96)  \
      prm.trace[((role) * 64 + ((step) - 32)) * 8 + (ev)] = clock64();                                 \
  } while (0)
#else
#define TRACE(role, step, ev) do {} while (0)
#endif

__host__ __device__ inline int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

struct TileRange { int jlo, jhi; };

// key-tile stream of one CTA (a pair of 128-query tiles starting at q0), in 64-key tiles
template <int NQT>
__device__ __forceinline__ void tile_ranges(const TcParams& prm, int q0, int& kbase, int& nj, TileRange (&tr)[2]) {
  tr[1].jlo = 0; tr[1].jhi = 0;
  if (prm.mode == MODE_CIRCULANT) {
    kbase = floor_div(q0 - prm.p, BN) * BN;
    nj = 0;
#pragma unroll
    for (int t = 0; t < NQT; ++t) {
      const int i0 = q0 + 128 * t;
      if (i0 < prm.N) {
        tr[t].jlo = floor_div(i0 - prm.p - kbase, BN);
        tr[t].jhi = floor_div(i0 + 127 - prm.p + prm.W - 1 - kbase, BN) + 1;
        nj = tr[t].jhi > nj ? tr[t].jhi : nj;
      } else { tr[t].jlo = 0; tr[t].jhi = 0; }
    }
  } else {
    kbase = 0;
    nj = (prm.N + BN - 1) / BN;
#pragma unroll
    for (int t = 0; t < NQT; ++t) {
      tr[t].jlo = 0;
      tr[t].jhi = (q0 + 128 * t < prm.N) ? nj : 0;
    }
  }
}

template <int FMT>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (FMT == 1) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int D, int FMT, int NQT, int SPLIT>
__global__ void __launch_bounds__(Cfg<D, NQT, SPLIT>::THREADS, Cfg<D, NQT, SPLIT>::CTAS_PER_SM)
tc_fwd_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
              const __grid_constant__ CUtensorMap tmv, const TcParams prm) {
  using C = Cfg<D, NQT, SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = sbase + C::OFF_Q, sK = sbase + C::OFF_K, sV = sbase + C::OFF_V;
  const uint32_t bars = sbase + C::OFF_BAR;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  const uint32_t tmem_slot = sbase + C::OFF_TMEM_SLOT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (128 * NQT), b = blockIdx.y;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmq); prefetch_tensormap(&tmk); prefetch_tensormap(&tmv);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) mbar_init(bar(C::BAR_QFULL + i), 1);
    // K/V slots are released by every MMA issuer (count NQT)
    for (int i = 0; i < C::K_STAGES; ++i) { mbar_init(bar(C::BAR_KFULL + i), 1); mbar_init(bar(C::BAR_KEMPTY + i), NQT); }
    for (int i = 0; i < C::V_STAGES; ++i) { mbar_init(bar(C::BAR_VFULL + i), 1); mbar_init(bar(C::BAR_VEMPTY + i), NQT); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar(C::BAR_SFULL + i), 1);
      mbar_init(bar(C::BAR_PFULL + i), SPLIT == 2 ? 256 : 128);
      mbar_init(bar(C::BAR_MAX + i), 128);
    }
    if (SPLIT == 2) for (int i = 0; i < 32; ++i) mbar_init(bar(C::BAR_MXX + i), 32);
    for (int i = 0; i < 2; ++i) { mbar_init(bar(C::BAR_ODONE + i), 1); mbar_init(bar(C::BAR_OFINAL + i), 1); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  int kbase, nj;
  TileRange tr[2];
  tile_ranges<NQT>(prm, q0, kbase, nj, tr);

  if (warp < 4) {
    // register pool of the CTA = THREADS x launch registers (168 for NQT = 2, 128 for NQT = 1):
    // 4 x 32 x 64 + 8 x 32 x 216 = 63488 <= 384 x 168;  4 x 32 x 56 + 4 x 32 x 200 = 32768 = 256 x 128
    // SPLIT = 2: 640 x 96 launch registers:  4 x 32 x 56 + 16 x 32 x 104 = 60416 <= 61440
    // HELP: 4 x 32 x 56 + 8 x 32 x 64 (helpers) + 8 x 32 x 144 (softmax) = 60416 <= 640 x 96
    if (SPLIT >= 2) setmaxnreg_dec<56>(); else if (NQT == 2) setmaxnreg_dec<64>(); else setmaxnreg_dec<56>();
    if (warp == 0 && lane == 0) {
      // ------------------------------------------------------------ TMA producer
#pragma unroll
      for (int t = 0; t < NQT; ++t) {
        if (tr[t].jhi > tr[t].jlo) {
          mbar_arrive_expect_tx(bar(C::BAR_QFULL + t), C::QTILE_BYTES);
          tma_load_3d(sQ + t * C::QTILE_BYTES, &tmq, bar(C::BAR_QFULL + t), q0 + 128 * t, 0, b);
          tma_load_3d(sQ + t * C::QTILE_BYTES + C::BOX_BYTES, &tmq, bar(C::BAR_QFULL + t), q0 + 128 * t + 64, 0, b);
        }
      }
      for (int j = 0; j < nj; ++j) {
        const int sk = j % C::K_STAGES, sv = j % C::V_STAGES;
        const int tok = (prm.mode == MODE_CIRCULANT) ? (int)pmod(kbase + BN * j, prm.N) : BN * j;
        mbar_wait(bar(C::BAR_KEMPTY + sk), ((uint32_t)(j / C::K_STAGES) & 1u) ^ 1u);
        mbar_arrive_expect_tx(bar(C::BAR_KFULL + sk), C::BOX_BYTES);
        tma_load_3d(sK + sk * C::BOX_BYTES, &tmk, bar(C::BAR_KFULL + sk), tok, 0, b);
        mbar_wait(bar(C::BAR_VEMPTY + sv), ((uint32_t)(j / C::V_STAGES) & 1u) ^ 1u);
        mbar_arrive_expect_tx(bar(C::BAR_VFULL + sv), C::BOX_BYTES);
        tma_load_3d(sV + sv * C::BOX_BYTES, &tmv, bar(C::BAR_VFULL + sv), tok, 0, b);
      }
    } else if (warp == 1 || (warp == 3 && NQT == 2)) {
      // ------------------------------------------------------------ MMA issuer of Q tile t
      // The whole warp runs this loop (all values warp-uniform -> uniform registers); only the
      // tcgen05 instructions themselves are issued by one elected lane.
      const int t = warp >> 1;
      constexpr uint32_t idesc_qk = make_idesc_f16(FMT, FMT, 1, 1, 128, BN);   // A, B MN-major
      // P takes the input format: mixing A = f16 with B = bf16 in one kind::f16 MMA raises
      // "illegal instruction" on sm_100a (measured, round 1), so bf16 inputs mean bf16 P.
      constexpr uint32_t idesc_pv = make_idesc_f16(FMT, FMT, 0, 0, 128, D);    // A in TMEM, B K-major
      const int jlo = tr[t].jlo, jhi = tr[t].jhi;
      const uint64_t qdesc = make_smem_desc_sw128(sQ + t * C::QTILE_BYTES, C::BOX_BYTES, 1024);
      const uint64_t kdesc = make_smem_desc_sw128(sK, C::BOX_BYTES, 1024);
      const uint64_t vdesc = make_smem_desc_sw128(sV, 16, 1024);
      const uint32_t tSt = tmem_base + C::COL_S + 128 * t;
      const uint32_t tOt = tmem_base + C::COL_O + D * t;
      for (int g = 0; g <= nj + 1; ++g) {
        const int jp = g - 2;
        // Both issuers wait for every K/V tile, used or not: the empty barriers count 2 arrivals
        // per phase, so an issuer must never run a whole ring ahead of the producer.
        if (jp >= 0 && jp < nj) mbar_wait(bar(C::BAR_VFULL + jp % C::V_STAGES), (uint32_t)(jp / C::V_STAGES) & 1u);
        if (jp >= jlo && jp < jhi) {                                         // O_t += P_t(jp) V(jp)
          const int sv = jp % C::V_STAGES, i = jp - jlo, bb = i & 1;
          TRACE(t, jp, 0);
          mbar_wait(bar(C::BAR_PFULL + 2 * t + bb), (uint32_t)(i >> 1) & 1u);
          TRACE(t, jp, 1);
          tc_fence_after();
          const uint64_t vd = vdesc + (uint64_t)(sv * (C::BOX_BYTES >> 4));
          const uint32_t tP = tSt + 64 * bb;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < BN / 16; ++ks)
              mma_ts(tOt, tP + ks * 8, vd + (uint64_t)(ks * 2), idesc_pv, (i > 0 || ks > 0) ? 1u : 0u);
            tc_commit(bar(C::BAR_ODONE + t));
          }
          __syncwarp();
        }
        TRACE(t, g, 2);
        if (g < nj) mbar_wait(bar(C::BAR_KFULL + g % C::K_STAGES), (uint32_t)(g / C::K_STAGES) & 1u);
        TRACE(t, g, 3);
        if (g >= jlo && g < jhi) {                                           // S_t[b] = Q_t K(g)^T
          const int sk = g % C::K_STAGES, i = g - jlo, bb = i & 1;
          if (i == 0) mbar_wait(bar(C::BAR_QFULL + t), 0);
          tc_fence_after();
          const uint64_t kd = kdesc + (uint64_t)(sk * (C::BOX_BYTES >> 4));
          const uint32_t tS = tSt + 64 * bb;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks)
              mma_ss(tS, qdesc + (uint64_t)(ks * 128), kd + (uint64_t)(ks * 128), idesc_qk, ks > 0 ? 1u : 0u);
            tc_commit(bar(C::BAR_SFULL + 2 * t + bb));
          }
          __syncwarp();
          TRACE(t, g, 4);
        }
        // release the K/V slots of this step (every issuer arrives, whether or not it used them)
        if (elect_one()) {
          if (jp >= 0 && jp < nj) tc_commit(bar(C::BAR_VEMPTY + jp % C::V_STAGES));
          if (g < nj) tc_commit(bar(C::BAR_KEMPTY + g % C::K_STAGES));
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(bar(C::BAR_OFINAL + t));
      __syncwarp();
    }
  } else if (C::HELP && warp >= 4 + 4 * NQT) {
    // -------------------------------------------------------------- helper warpgroups: row max of S(j)
    setmaxnreg_dec<64>();
    const int t = (warp - 4 - 4 * NQT) >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS0 = tmem_base + lane_addr + C::COL_S + 128 * t;
    const int qi = q0 + 128 * t + row;
    const int jlo = tr[t].jlo, jhi = tr[t].jhi;
    const bool circ = prm.mode == MODE_CIRCULANT;
    const uint32_t hm = sbase + C::OFF_XCH + (uint32_t)(t * 2 * 128 * 4);      // float[2 buffers][128 rows]
    for (int j = jlo; j < jhi; ++j) {
      const int i = j - jlo, bb = i & 1;
      mbar_wait(bar(C::BAR_SFULL + 2 * t + bb), (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      int lo = 0, hi = BN;
      if (circ) { lo = (qi - prm.p) - (kbase + BN * j); hi = lo + prm.W; }
      else { hi = prm.N - BN * j; }
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t sv[32];
        tmem_ld32(tS0 + 64 * bb + 32 * c, sv);
        tmem_wait_ld();
        float m0 = -INFINITY, m1 = -INFINITY;
        if (lo > 32 * c || hi < 32 * c + 32) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int col = 32 * c + e;
            if (col < lo || col >= hi) sv[e] = 0xff800000u;
          }
        }
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          m0 = fmaxf(m0, fmaxf(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])));
          m1 = fmaxf(m1, fmaxf(__uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3])));
        }
        mx = fmaxf(mx, fmaxf(m0, m1));
      }
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(hm + (uint32_t)((bb * 128 + row) * 4)), "f"(mx) : "memory");
      mbar_arrive(bar(C::BAR_MAX + 2 * t + bb));      // release: the store above is visible to whoever sees the phase
    }
  } else if (SPLIT == 2) {
    // -------------------------------------------------------------- softmax, two threads per row (v2)
    // A lone warp's exp phase is the sum of its dispatch intervals (profiles/r1o_fwd_step_trace.md), so the
    // 64 columns of a step are split over two threads (two warps of the same lane quarter).  The row max of
    // S(j+1) is exchanged EARLY: each thread reduces its half right after fetching it at the end of step j
    // and publishes it through shared memory + an mbarrier, so the partner's value is there when step j+1
    // starts and no named barrier sits in the chain.
    setmaxnreg_inc<104>();
    const int sw = warp - 4, t = (sw >> 2) & 1, h = sw >> 3, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t tS0 = tmem_base + lane_addr + C::COL_S + 128 * t;
    const uint32_t tO = tmem_base + lane_addr + C::COL_O + D * t;
    const int qi = q0 + 128 * t + row;
    const float scale = prm.scale_log2;
    const int jlo = tr[t].jlo, jhi = tr[t].jhi;
    const uint32_t xch = sbase + C::OFF_XCH + (uint32_t)(t * 3 * 2 * 128 * 4);      // float [3 slots][2 halves][128 rows]
    auto xaddr = [&](int slot, int hh) { return xch + (uint32_t)(((slot * 2 + hh) * 128 + row) * 4); };
    auto xst = [&](int slot, int hh, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(xaddr(slot, hh)), "f"(v) : "memory"); };
    auto xld = [&](int slot, int hh) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(xaddr(slot, hh)) : "memory"); return v; };
    // max-exchange barriers: [tile][quarter][half][step parity], 32 arrivals each
    auto mxbar = [&](int hh, int par) { return bar(C::BAR_MXX + ((t * 4 + quarter) * 2 + hh) * 2 + par); };
    const bool circ = prm.mode == MODE_CIRCULANT;
    const int NN = prm.N, WW = prm.W, pp = prm.p;

    if (jhi > jlo) {
      float m_true = -INFINITY, m_used = -INFINITY;
      float2 l2a = make_float2(0.f, 0.f), l2b = make_float2(0.f, 0.f);
      const float2 scale2 = make_float2(scale, scale);
      // mask + local max of one fetched half-row, published for the partner
      auto reduce_publish = [&](uint32_t (&sv)[32], int j) -> float {
        int lo = 0, hi = BN;
        if (circ) { lo = (qi - pp) - (kbase + BN * j); hi = lo + WW; }
        else { hi = NN - BN * j; }
        if (lo > 32 * h || hi < 32 * h + 32) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int col = 32 * h + e;
            if (col < lo || col >= hi) sv[e] = 0xff800000u;
          }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3])));
        }
        const float ml = fmaxf(mx0, mx1) * scale;
        const int i = j - jlo;
        xst(i & 1, h, ml);
        mbar_arrive(mxbar(h, i & 1));          // release: the store is visible to whoever observes the phase
        return ml;
      };
      uint32_t sc[32], sn[32];
      mbar_wait(bar(C::BAR_SFULL + 2 * t), 0);
      tc_fence_after();
      tmem_ld32(tS0 + 32 * h, sc);
      tmem_wait_ld();
      float mloc = reduce_publish(sc, jlo);

      auto step = [&](const int j, uint32_t (&cur)[32], uint32_t (&nxt)[32]) {
        const int i = j - jlo, bb = i & 1;
        const uint32_t tS = tS0 + 64 * bb;
        const bool more = j + 1 < jhi;
        const uint32_t nbar = bar(C::BAR_SFULL + 2 * t + (bb ^ 1)), npar = (uint32_t)((i + 1) >> 1) & 1u;
        // partner's half of the row max (published at the end of its previous step)
        mbar_wait(mxbar(h ^ 1, i & 1), (uint32_t)(i >> 1) & 1u);
        m_true = fmaxf(m_true, fmaxf(mloc, xld(i & 1, h ^ 1)));
        const bool want = (m_true - m_used) > RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, want)) {      // identical in both warps of the pair: same rows, same m
          const float alpha = (m_used == -INFINITY) ? 0.f : ex2(m_used - m_true);
          if (i > 0) {
            mbar_wait(bar(C::BAR_ODONE + t), (uint32_t)(i - 1) & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int c = h * (D / 64); c < (h + 1) * (D / 64); ++c) {      // each thread rescales half the channels
              uint32_t o[32];
              tmem_ld32(tO + 32 * c, o);
              tmem_wait_ld();
#pragma unroll
              for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              tmem_st32(tO + 32 * c, o);
            }
          }
          l2a.x *= alpha; l2a.y *= alpha; l2b.x *= alpha; l2b.y *= alpha;
          m_used = m_true;
        }
        const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used;
        const float2 negm2 = make_float2(neg_m, neg_m);
        uint32_t pk[16];
        bool fetched = false;
#pragma unroll
        for (int e0 = 0; e0 < 32; e0 += 8) {
          float2 x[4], p[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            x[u] = __ffma2_rn(make_float2(__uint_as_float(cur[e0 + 2 * u]), __uint_as_float(cur[e0 + 2 * u + 1])), scale2, negm2);
#pragma unroll
          for (int u = 0; u < 4; ++u) { p[u].x = ex2(x[u].x); p[u].y = ex2(x[u].y); }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (u & 1) l2b = __fadd2_rn(l2b, p[u]); else l2a = __fadd2_rn(l2a, p[u]);
            pk[(e0 >> 1) + u] = pack2<FMT>(p[u].x, p[u].y);
          }
          if (e0 == 8 && more && mbar_test_wait(nbar, npar)) {      // S(j+1) usually complete by now
            tc_fence_after();
            tmem_ld32(tS0 + 64 * (bb ^ 1) + 32 * h, nxt);
            fetched = true;
          }
        }
        tmem_st16(tS + 16 * h, pk);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bar(C::BAR_PFULL + 2 * t + bb));
        if (more) {
          if (!fetched) {
            mbar_wait(nbar, npar);
            tc_fence_after();
            tmem_ld32(tS0 + 64 * (bb ^ 1) + 32 * h, nxt);
          }
          tmem_wait_ld();
          mloc = reduce_publish(nxt, j + 1);
        }
      };
      for (int j = jlo; j < jhi; j += 2) {
        step(j, sc, sn);
        if (j + 1 < jhi) step(j + 1, sn, sc);
      }
      const float l_half = (l2a.x + l2a.y) + (l2b.x + l2b.y);
      xst(2, h, l_half);
      asm volatile("bar.sync %0, 64;" ::"r"(1u + (uint32_t)(t * 4 + quarter)) : "memory");
      const float l_run = l_half + xld(2, h ^ 1);

      mbar_wait(bar(C::BAR_OFINAL + t), 0);
      tc_fence_after();
      const float inv_l = 1.f / l_run;
      const bool in_range = qi < prm.N;
#pragma unroll 1
      for (int c = h * (D / 64); c < (h + 1) * (D / 64); ++c) {
        uint32_t o[32];
        tmem_ld32(tO + 32 * c, o);
        tmem_wait_ld();
        if (in_range) {
          if (prm.o_f32) {
            float* ob = static_cast<float*>(prm.o) + (size_t)b * D * prm.N + qi;
#pragma unroll
            for (int e = 0; e < 32; ++e) ob[(size_t)(32 * c + e) * prm.N] = __uint_as_float(o[e]) * inv_l;
          } else if (FMT == 1) {
            __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(prm.o) + (size_t)b * D * prm.N + qi;
#pragma unroll
            for (int e = 0; e < 32; ++e) ob[(size_t)(32 * c + e) * prm.N] = __float2bfloat16_rn(__uint_as_float(o[e]) * inv_l);
          } else {
            __half* ob = static_cast<__half*>(prm.o) + (size_t)b * D * prm.N + qi;
#pragma unroll
            for (int e = 0; e < 32; ++e) ob[(size_t)(32 * c + e) * prm.N] = __float2half_rn(__uint_as_float(o[e]) * inv_l);
          }
        }
      }
      if (in_range && h == 0) {
        prm.l[(size_t)b * prm.N + qi] = l_run * ex2(m_used - m_true);
        prm.m[(size_t)b * prm.N + qi] = m_true * LN2;
      }
    }
  } else {
    // -------------------------------------------------------------- softmax warpgroups
    if (C::HELP) setmaxnreg_inc<144>(); else if (NQT == 2) setmaxnreg_inc<216>(); else setmaxnreg_inc<200>();
    const int t = (warp - 4) >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS0 = tmem_base + lane_addr + C::COL_S + 128 * t;
    const uint32_t tO = tmem_base + lane_addr + C::COL_O + D * t;
    const int qi = q0 + 128 * t + row;                 // query token of this thread
    const float scale = prm.scale_log2;
    const int jlo = tr[t].jlo, jhi = tr[t].jhi;

    if (jhi > jlo) {
      float m_true = -INFINITY, m_used = -INFINITY;
      const bool circ = prm.mode == MODE_CIRCULANT;
      const int NN = prm.N, WW = prm.W, pp = prm.p;
      float2 l2a = make_float2(0.f, 0.f), l2b = make_float2(0.f, 0.f);      // packed row-sum accumulators
      const float2 scale2 = make_float2(scale, scale);

      // both tiles active with the same number of steps (dense, not the last partial CTA): safe to alternate
      const bool pp_on = NQT == 2 && prm.pingpong && prm.mode == MODE_DENSE && tr[0].jhi == tr[1].jhi && tr[1].jhi > tr[1].jlo;
      // One 64-key step on the S row held in `sc`; S(j+1) is prefetched from TMEM into `sn`
      // BEFORE the exp phase (QK runs two steps ahead, so it is normally ready), which hides
      // the mbarrier + tcgen05.ld latency behind the MUFU-bound part of the step.
      auto softmax_step = [&](const int j, uint32_t (&sc)[2][32], uint32_t (&sn)[2][32]) {
        const int i = j - jlo, bb = i & 1;
        const uint32_t tS = tS0 + 64 * bb;
        if ((warp & 3) == 0) TRACE(2 + t, j, 0);          // step start (S(j) already in registers)
        // opportunistic prefetch of S(j+1): taken now only if the tensor pipe already delivered it
        const bool more = j + 1 < jhi;
        const uint32_t nbar = bar(C::BAR_SFULL + 2 * t + (bb ^ 1)), npar = (uint32_t)((i + 1) >> 1) & 1u;
        bool fetched = false;
        if (FA_PREFETCH && more && mbar_test_wait(nbar, npar)) {
          tc_fence_after();
          tmem_ld32(tS0 + 64 * (bb ^ 1), sn[0]);
          tmem_ld32(tS0 + 64 * (bb ^ 1) + 32, sn[1]);
          fetched = true;
        }
        // ---- mask (last dense tile / circulant band edges); lo <= col < hi stays
        int lo = 0, hi = BN;
        if (circ) { lo = (qi - pp) - (kbase + BN * j); hi = lo + WW; }
        else { hi = NN - BN * j; }
        if (lo > 0 || hi < BN) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int col = 32 * c + e;
              if (col < lo || col >= hi) sc[c][e] = 0xff800000u;   // -inf
            }
        }
        // A 64-key tile that lies outside the band of EVERY row of this warp (circulant: the CTA's key range
        // covers 128 + W - 1 keys, a warp's rows only 32 + W - 1) contributes P = 0: skip max, rescale and the
        // 64 exponentials, publish zeros.  One warp-uniform branch per step.
        // (Only in the one-Q-tile instantiation the band kernels use: in the NQT = 2 kernel the extra branch costs
        // the dense path 4 %, measured.)
        const bool allmasked = NQT == 1 && __all_sync(0xffffffffu, hi <= 0 || lo >= BN);
        if (allmasked) {
          uint32_t z[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) z[e] = 0u;
          tmem_st16(tS, z);
          tmem_st16(tS + 16, z);
        } else {
        // ---- running max (thread-local: one thread owns one row)
        if (C::HELP) {
          // the helper warpgroup of this tile has already reduced S(j) (it saw s_full long before we got here)
          mbar_wait(bar(C::BAR_MAX + 2 * t + bb), (uint32_t)(i >> 1) & 1u);
          float hmx;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(hmx) : "r"(sbase + C::OFF_XCH + (uint32_t)(((t * 2 + bb) * 128 + row) * 4)) : "memory");
          m_true = fmaxf(m_true, hmx * scale);
        } else {
          float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sc[0][e]), __uint_as_float(sc[0][e + 1])));
            mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sc[1][e]), __uint_as_float(sc[1][e + 1])));
          }
          m_true = fmaxf(m_true, fmaxf(mx0, mx1) * scale);
        }
        if ((warp & 3) == 0) TRACE(2 + t, j, 1);          // max done
        // ---- lazy rescale of O and l (warp-uniform decision; this warp owns its 32 TMEM lanes)
        const bool want = (m_true - m_used) > RESCALE_THRESHOLD;   // inf on first use
        if (__any_sync(0xffffffffu, want)) {
          const float alpha = (m_used == -INFINITY) ? 0.f : ex2(m_used - m_true);
          if (i > 0) {
            // O_t holds PV(jlo .. j-1); PV(j-1) was issued after our previous arrive and must have
            // completed before O is touched.  o_done counts one completion per PV_t; at this point
            // it is at most one phase behind (s_full(j) already implies PV(j-2) completed).
            mbar_wait(bar(C::BAR_ODONE + t), (uint32_t)(i - 1) & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < D / 32; ++c) {
              uint32_t o[32];
              tmem_ld32(tO + 32 * c, o);
              tmem_wait_ld();
#pragma unroll
              for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              tmem_st32(tO + 32 * c, o);
            }
          }
          l2a.x *= alpha; l2a.y *= alpha; l2b.x *= alpha; l2b.y *= alpha;
          m_used = m_true;
        }
        const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used;   // guard (-inf)-(-inf)
        const float2 negm2 = make_float2(neg_m, neg_m);
        // ping-pong: the exp phase (64 MUFU per thread) of one tile runs while the other tile's warpgroup is in
        // its MUFU-free part (row max, publish, S fetch); the token is a pair of 256-thread named barriers
        if (pp_on) asm volatile("bar.sync %0, 256;" ::"r"(2 + t) : "memory");
        if ((warp & 3) == 0) TRACE(2 + t, j, 4);          // exp phase entered (token held)
        // ---- P = exp2(s*scale - m) -> 16-bit, written over the first 32 columns of this S buffer.
        // The last EMU elements of every 32 go through the FMA pipe (Cody-Waite split + degree-3
        // minimax polynomial, rel. error 9e-5 << 16-bit P rounding) to unload the MUFU.
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t pk[16];
#if FA_EXP_SWP
          // software pipeline: the ex2 pair of column pair u is issued FA_EXP_SWP pairs before its sum / pack,
          // interleaved with them (MUFU accepts one warp instruction per 8 clk: the 7 slots in between are
          // filled with the packed-math instructions of older pairs instead of stalling on the pipe)
          {
            float2 pb[FA_EXP_SWP];
#pragma unroll
            for (int u = 0; u < 16 + FA_EXP_SWP; ++u) {
              if (u >= FA_EXP_SWP) {
                const int w = u - FA_EXP_SWP;
                if (w & 1) l2b = __fadd2_rn(l2b, pb[w % FA_EXP_SWP]); else l2a = __fadd2_rn(l2a, pb[w % FA_EXP_SWP]);
                pk[w] = pack2<FMT>(pb[w % FA_EXP_SWP].x, pb[w % FA_EXP_SWP].y);
              }
              if (u < 16) {
                const float2 x = __ffma2_rn(make_float2(__uint_as_float(sc[c][2 * u]), __uint_as_float(sc[c][2 * u + 1])), scale2, negm2);
                pb[u % FA_EXP_SWP] = make_float2(ex2(x.x), ex2(x.y));
              }
            }
          }
#elif FA_EXP_BLOCK
          // blocks of 8: all eight ex2 are issued before the first result is consumed, so the sum / pack
          // instructions never wait on a MUFU that was issued two instructions earlier
#pragma unroll
          for (int e0 = 0; e0 < 32; e0 += FA_EXP_BLOCK) {
            float2 x[FA_EXP_BLOCK / 2], p[FA_EXP_BLOCK / 2];
#pragma unroll
            for (int u = 0; u < FA_EXP_BLOCK / 2; ++u)
              x[u] = __ffma2_rn(make_float2(__uint_as_float(sc[c][e0 + 2 * u]), __uint_as_float(sc[c][e0 + 2 * u + 1])), scale2, negm2);
#pragma unroll
            for (int u = 0; u < FA_EXP_BLOCK / 2; ++u) {
              if (u >= FA_EXP_BLOCK / 2 - FA_EMU_PAIRS) {
                // this pair on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial, rel. error 9e-5,
                // far below the 16-bit rounding of P): fills issue slots the MUFU pipe leaves idle
                float2 xx = x[u];
                xx.x = fmaxf(xx.x, -126.f); xx.y = fmaxf(xx.y, -126.f);
                const float2 xf = __fadd2_rd(xx, make_float2(12582912.f, 12582912.f));
                const float2 xr = __fadd2_rn(xf, make_float2(-12582912.f, -12582912.f));
                const float2 fr = __fadd2_rn(xx, make_float2(-xr.x, -xr.y));
                float2 qq = __ffma2_rn(fr, make_float2(0.077119089663028717f, 0.077119089663028717f),
                                       make_float2(0.227564394474029541f, 0.227564394474029541f));
                qq = __ffma2_rn(qq, fr, make_float2(0.695146143436431885f, 0.695146143436431885f));
                qq = __ffma2_rn(qq, fr, make_float2(1.f, 1.f));
                p[u].x = __uint_as_float(__float_as_uint(qq.x) + (__float_as_uint(xf.x) << 23));
                p[u].y = __uint_as_float(__float_as_uint(qq.y) + (__float_as_uint(xf.y) << 23));
              } else {
                p[u].x = ex2(x[u].x); p[u].y = ex2(x[u].y);
              }
            }
#pragma unroll
            for (int u = 0; u < FA_EXP_BLOCK / 2; ++u) {
              if (u & 1) l2b = __fadd2_rn(l2b, p[u]); else l2a = __fadd2_rn(l2a, p[u]);
              pk[(e0 >> 1) + u] = pack2<FMT>(p[u].x, p[u].y);
            }
          }
#else
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
#if FA_PACKED
            float2 x = __ffma2_rn(make_float2(__uint_as_float(sc[c][e]), __uint_as_float(sc[c][e + 1])), scale2, negm2);
#else
            float2 x = make_float2(fmaf(__uint_as_float(sc[c][e]), scale, neg_m), fmaf(__uint_as_float(sc[c][e + 1]), scale, neg_m));
#endif
            float2 p;
            if (e >= 32 - EMU_PER_32) {
              x.x = fmaxf(x.x, -126.f); x.y = fmaxf(x.y, -126.f);
              const float2 xf = __fadd2_rd(x, make_float2(12582912.f, 12582912.f));    // floor(x) in the low mantissa bits
              const float2 xr = __fadd2_rn(xf, make_float2(-12582912.f, -12582912.f)); // floor(x)
              const float2 fr = __fadd2_rn(x, make_float2(-xr.x, -xr.y));              // frac in [0,1)
              float2 q = __ffma2_rn(fr, make_float2(0.077119089663028717f, 0.077119089663028717f),
                                    make_float2(0.227564394474029541f, 0.227564394474029541f));
              q = __ffma2_rn(q, fr, make_float2(0.695146143436431885f, 0.695146143436431885f));
              q = __ffma2_rn(q, fr, make_float2(1.f, 1.f));
              p.x = __uint_as_float(__float_as_uint(q.x) + (__float_as_uint(xf.x) << 23));
              p.y = __uint_as_float(__float_as_uint(q.y) + (__float_as_uint(xf.y) << 23));
            } else {
              p.x = ex2(x.x); p.y = ex2(x.y);
            }
#if FA_PACKED
            if (e & 2) l2b = __fadd2_rn(l2b, p); else l2a = __fadd2_rn(l2a, p);
#else
            l2a.x += p.x; l2a.y += p.y;
#endif
            pk[e >> 1] = pack2<FMT>(p.x, p.y);
          }
#endif
          tmem_st16(tS + 16 * c, pk);
          if (c == 0 && (warp & 3) == 0) TRACE(2 + t, j, 6);          // first half of the exps issued
          if (FA_PREFETCH_MID && c == 0 && more && !fetched && mbar_test_wait(nbar, npar)) {
            tc_fence_after();
            tmem_ld32(tS0 + 64 * (bb ^ 1), sn[0]);
            tmem_ld32(tS0 + 64 * (bb ^ 1) + 32, sn[1]);
            fetched = true;
          }
          if (c == 0 && (warp & 3) == 0) TRACE(2 + t, j, 7);          // mid-phase fetch issued
          // soft token (pingpong = 2): hand over after the first half, so that the second half of this tile and the first
          // half of the other one share the MUFU pipe (two warps per scheduler reach the pipe rate, a lone warp does not)
          if (c == 0 && pp_on && prm.pingpong == 2) asm volatile("bar.arrive %0, 256;" ::"r"(2 + (t ^ 1)) : "memory");
        }
        }
        if (pp_on && prm.pingpong != 2) asm volatile("bar.arrive %0, 256;" ::"r"(2 + (t ^ 1)) : "memory");
        if ((warp & 3) == 0) TRACE(2 + t, j, 2);          // exps + pack + st issued
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bar(C::BAR_PFULL + 2 * t + bb));
        if ((warp & 3) == 0) TRACE(2 + t, j, 3);          // P published
        if (more && !fetched) {                // late path: S(j+1) was not ready before the exp phase
          mbar_wait(nbar, npar);
          tc_fence_after();
          tmem_ld32(tS0 + 64 * (bb ^ 1), sn[0]);
          tmem_ld32(tS0 + 64 * (bb ^ 1) + 32, sn[1]);
        }
        tmem_wait_ld();                        // S(j+1) registers are valid from here on
        if ((warp & 3) == 0) TRACE(2 + t, j, 5);          // S(j+1) in registers
      };

      if (pp_on && t == 1) asm volatile("bar.arrive %0, 256;" ::"r"(2) : "memory");     // tile 0 goes first
      uint32_t sA[2][32], sB[2][32];
      if (t == 1 && prm.stagger > 0) {           // de-phase the two warpgroups: see launch_tc
        const long long t0 = clock64();
        while (clock64() - t0 < prm.stagger) {}
      }
      mbar_wait(bar(C::BAR_SFULL + 2 * t), 0);
      tc_fence_after();
      tmem_ld32(tS0, sA[0]);
      tmem_ld32(tS0 + 32, sA[1]);
      tmem_wait_ld();
      for (int j = jlo; j < jhi; j += 2) {
        softmax_step(j, sA, sB);
        if (j + 1 < jhi) softmax_step(j + 1, sB, sA);
      }
      const float l_run = (l2a.x + l2a.y) + (l2b.x + l2b.y);
      if (pp_on && t == 0) asm volatile("bar.sync %0, 256;" ::"r"(2) : "memory");       // consume tile 1's last hand-off

      // ---- epilogue: O / l -> global (token-contiguous rows: a warp writes 32 consecutive tokens)
      mbar_wait(bar(C::BAR_OFINAL + t), 0);
      tc_fence_after();
      const float inv_l = 1.f / l_run;
      const bool in_range = qi < prm.N;
      if (prm.o_f32 == 2) {
        // ring pass, steps after the first: `o`, `l`, `m` hold the running (normalised O, l, m) of the blocks seen so
        // far; this block's partial is merged in place with the update rule of src/dense.jl:82-91 -- no separate
        // partial buffer and merge kernel (3 extra passes over N x dv x 4 bytes per ring step in round 1)
        float* ob = static_cast<float*>(prm.o) + (size_t)b * D * prm.N + qi;
        float c1 = 0.f, c2 = 0.f, ln = 0.f, mn = 0.f;
        if (in_range) {
          const float m2 = m_true * LN2, l2 = l_run * ex2(m_used - m_true);
          const float m1 = prm.m[(size_t)b * prm.N + qi], l1 = prm.l[(size_t)b * prm.N + qi];
          mn = fmaxf(m1, m2);
          const float w1 = (m1 == -INFINITY) ? 0.f : l1 * __expf(m1 - mn);
          const float w2 = (m2 == -INFINITY) ? 0.f : l2 * __expf(m2 - mn);
          ln = w1 + w2;
          const float inv = ln > 0.f ? 1.f / ln : 0.f;
          c1 = w1 * inv; c2 = w2 * inv * inv_l;
        }
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          uint32_t o[32];
          tmem_ld32(tO + 32 * c, o);
          tmem_wait_ld();
          if (in_range) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              float* po = ob + (size_t)(32 * c + e) * prm.N;
              *po = *po * c1 + __uint_as_float(o[e]) * c2;
            }
          }
        }
        if (in_range) { prm.l[(size_t)b * prm.N + qi] = ln; prm.m[(size_t)b * prm.N + qi] = mn; }
      } else if (prm.o_f32) {
        float* ob = static_cast<float*>(prm.o) + (size_t)b * D * prm.N + qi;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          uint32_t o[32];
          tmem_ld32(tO + 32 * c, o);
          tmem_wait_ld();
          if (in_range) {
#pragma unroll
            for (int e = 0; e < 32; ++e) ob[(size_t)(32 * c + e) * prm.N] = __uint_as_float(o[e]) * inv_l;
          }
        }
      } else if (FMT == 1) {
        __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(prm.o) + (size_t)b * D * prm.N + qi;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          uint32_t o[32];
          tmem_ld32(tO + 32 * c, o);
          tmem_wait_ld();
          if (in_range) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              ob[(size_t)(32 * c + e) * prm.N] = __float2bfloat16_rn(__uint_as_float(o[e]) * inv_l);
          }
        }
      } else {
        __half* ob = static_cast<__half*>(prm.o) + (size_t)b * D * prm.N + qi;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          uint32_t o[32];
          tmem_ld32(tO + 32 * c, o);
          tmem_wait_ld();
          if (in_range) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              ob[(size_t)(32 * c + e) * prm.N] = __float2half_rn(__uint_as_float(o[e]) * inv_l);
          }
        }
      }
      if (in_range && prm.o_f32 != 2) {
        // l = sum exp(s - m_true), m = max s (natural-log domain), reference src/dense.jl:12-18
        prm.l[(size_t)b * prm.N + qi] = l_run * ex2(m_used - m_true);
        prm.m[(size_t)b * prm.N + qi] = m_true * LN2;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// tensor map over one [B][D][N] tensor: box = 64 tokens x D channels x 1 batch, SWIZZLE_128B
int make_tmap(CUtensorMap* tm, const void* base, int dtype, long long N, int D, long long B, long long stride_c = 0, long long stride_b = 0) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return FA_ERR_CUDA; }
  const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)D, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)(stride_c ? stride_c : N * 2), (cuuint64_t)(stride_b ? stride_b : N * D * 2)};
  const cuuint32_t box[3] = {64, (cuuint32_t)D, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = dtype == FA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(tm, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return FA_ERR_CUDA; }
  return FA_OK;
}

template <int D, int FMT, int NQT, int SPLIT = 1>
int launch_tc(const Geo& g, const FwdArgs& a, int dtype, cudaStream_t st) {
  CUtensorMap tmq, tmk, tmv;
  int rc;
  if ((rc = make_tmap(&tmq, a.q, dtype, g.N, D, g.B, a.in_stride_c, a.in_stride_b))) return rc;
  if ((rc = make_tmap(&tmk, a.k, dtype, g.N, D, g.B, a.in_stride_c, a.in_stride_b))) return rc;
  if ((rc = make_tmap(&tmv, a.v, dtype, g.N, D, g.B, a.in_stride_c, a.in_stride_b))) return rc;
  TcParams prm;
  prm.o = a.o; prm.l = a.l; prm.m = a.m;
  prm.N = (int)g.N; prm.B = (int)g.B; prm.W = g.W; prm.p = g.p; prm.mode = g.mode;
  prm.scale_log2 = g.tau * LOG2E;
  prm.o_f32 = a.o_f32;
  static const int stagger = [] { const char* e = getenv("FA_FWD_STAGGER"); return e ? atoi(e) : 0; }();
  prm.stagger = stagger;
  // 1 = the token covers the whole exp phase: 1.826 vs 1.860 ms at N=8192, d=128, B=64 (same box, three interleaved runs);
  // 2 (default) = handed over after the first half: 1.752 vs 1.815 ms (profiles/r2x_dense_fwd_experiments.md);
  // FA_FWD_PINGPONG=0 disables
  static const int pingpong = [] { const char* e = getenv("FA_FWD_PINGPONG"); return e ? atoi(e) : 2; }();
  prm.pingpong = pingpong;
  prm.trace = nullptr;
#ifdef FA_TRACE
  { const char* e = getenv("FA_TRACE_PTR"); prm.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr; }
#endif
  using C = Cfg<D, NQT, SPLIT>;
  auto kern = tc_fwd_kernel<D, FMT, NQT, SPLIT>;
  FA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  const dim3 grid((unsigned)((g.N + 128 * NQT - 1) / (128 * NQT)), (unsigned)g.B);
  kern<<<grid, C::THREADS, C::SMEM_BYTES, st>>>(tmq, tmk, tmv, prm);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

}  // namespace

int make_tmap_public(CUtensorMap* tm, const void* base, int dtype, long long N, int D, long long B, long long stride_c, long long stride_b) {
  return make_tmap(tm, base, dtype, N, D, B, stride_c, stride_b);
}

bool tc_fwd_supported(const Geo& g, int dtype) {
  if (dtype != FA_BF16 && dtype != FA_F16) return false;
  if (g.mode != MODE_DENSE && g.mode != MODE_CIRCULANT) return false;
  if (g.d != g.dv || (g.d != 32 && g.d != 64 && g.d != 128)) return false;
  if (g.N % 8 != 0 || g.N < 8 || g.N > 0x3fffffff) return false;    // TMA: 16-byte global strides
  if (g.B > 65535) return false;                                     // gridDim.y
  if (g.mode == MODE_CIRCULANT && (g.N % 64 != 0)) return false;     // tile-aligned wrap-around
  return true;
}

int tc_fwd(const Geo& g, const FwdArgs& a, int dtype, cudaStream_t st) {
  if (!tc_fwd_supported(g, dtype)) { set_error("tc_fwd: unsupported configuration"); return FA_ERR_UNSUPPORTED; }
  if ((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v)) & 15) {
    set_error("tc_fwd: q/k/v must be 16-byte aligned"); return FA_ERR_INVALID;
  }
  const int fmt = dtype == FA_BF16 ? 1 : 0;
  if (g.d == 32) {      // d = 32 (the reference's logged benchmark head dim) exists on the band kernel only
    if (a.o_f32) { set_error("FA_FLAG_OUT_F32 needs a shape the tcgen05 kernels cover with float32 output (d in {64,128})"); return FA_ERR_UNSUPPORTED; }
    return tc_band_fwd(g, a, dtype, st);
  }
  // short key loops (circulant with a band of a few tiles): one Q tile per CTA, two CTAs per SM
  const bool short_loop = g.mode == MODE_CIRCULANT && (256 + g.W) / 64 <= 24;
  if (short_loop) {
    // d = 64: the compact band kernel (three CTAs per SM, fa_tc_band.cu); FA_BAND=0 keeps the two-CTA variant below
    static const int band = [] { const char* e = getenv("FA_BAND"); return e ? atoi(e) : 1; }();
    if (band && g.d == 64 && !a.o_f32) return tc_band_fwd(g, a, dtype, st);
    // d = 128: the band kernel's two-CTA configuration (N = 16384, W = 255, B = 256: 1.34 vs 2.19 ms); FA_BAND128=0 disables
    static const int band128 = [] { const char* e = getenv("FA_BAND128"); return e ? atoi(e) : 1; }();
    if (band && band128 && g.d == 128 && !a.o_f32) return tc_band_fwd(g, a, dtype, st);
    if (g.d == 128) return fmt ? launch_tc<128, 1, 1>(g, a, dtype, st) : launch_tc<128, 0, 1>(g, a, dtype, st);
    return fmt ? launch_tc<64, 1, 1>(g, a, dtype, st) : launch_tc<64, 0, 1>(g, a, dtype, st);
  }
  // dense, d = 64: the band kernel (serial CTAs, four per SM) beats the pair kernel -- N = 8192, B = 128: 2.63 vs 3.28 ms
  // (836 vs 670 TFLOP/s; a quarter of the exponentials on the FMA pipe); at d = 128 the pair kernel wins (1.79 vs 2.09 ms at B = 64).  FA_DENSE_BAND=0 / 2 = never / always.
  static const int dense_band = [] { const char* e = getenv("FA_DENSE_BAND"); return e ? atoi(e) : 1; }();
  if (g.mode == MODE_DENSE && !a.o_f32 && ((dense_band == 1 && g.d == 64) || dense_band == 2)) return tc_band_fwd(g, a, dtype, st);
#ifdef FA_TRACE
  // experiment variants (two softmax threads per row / helper warpgroup), both measured slower (profiles/r1o): only
  // instantiated in the -DFA_TRACE build (`make trace`), not in the product library
  static const int split = [] { const char* e = getenv("FA_FWD_SPLIT"); return e ? atoi(e) : 1; }();
  if (split == 2 && g.d == 128) return fmt ? launch_tc<128, 1, 2, 2>(g, a, dtype, st) : launch_tc<128, 0, 2, 2>(g, a, dtype, st);
  if (split == 3 && g.d == 128) return fmt ? launch_tc<128, 1, 2, 3>(g, a, dtype, st) : launch_tc<128, 0, 2, 3>(g, a, dtype, st);
#endif
  if (g.d == 128) return fmt ? launch_tc<128, 1, 2>(g, a, dtype, st) : launch_tc<128, 0, 2>(g, a, dtype, st);
  return fmt ? launch_tc<64, 1, 2>(g, a, dtype, st) : launch_tc<64, 0, 2>(g, a, dtype, st);
}

}  // namespace fa
