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
// CTA = 384 threads, one CTA per SM (512 TMEM columns):
//   warp 0      TMA producer (Q pair once; K and V tiles through two mbarrier rings)
//   warp 1      MMA issuer   (single thread; S_t = Q_t K_j^T, O_t += P_t V_j, ping-pong t = 0,1)
//   warp 2      TMEM allocator
//   warps 4-7   softmax for Q tile 0   (thread == row; S row read from TMEM with tcgen05.ld,
//   warps 8-11  softmax for Q tile 1    running max/sum in registers, exp2, P written back to
//                                       TMEM as 16-bit, lazy O rescale, final O/l, m epilogue)
// TMEM columns: S0 [0,128) S1 [128,256) O0 [256,256+d) O1 [256+d,256+2d); P_t aliases S_t.
#include <cuda.h>
#include "fa_common.cuh"
#include "fa_ptx.cuh"

namespace fa {
namespace {

using namespace ptx;

constexpr int TC_THREADS = 384;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float RESCALE_THRESHOLD = 8.0f;   // lazy rescale: P may grow to 2^8 before O is rescaled

template <int D>
struct Cfg {
  static constexpr int KV_STAGES = (D == 128) ? 2 : 4;
  static constexpr int BOX_BYTES = 64 * D * 2;          // 64 tokens x D channels, 16-bit
  static constexpr int TILE_BYTES = 2 * BOX_BYTES;      // 128 tokens
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = 2 * TILE_BYTES;
  static constexpr int OFF_V = OFF_K + KV_STAGES * TILE_BYTES;
  static constexpr int OFF_BAR = OFF_V + KV_STAGES * TILE_BYTES;
  // barrier slots (8 B each)
  static constexpr int BAR_QFULL = 0;                    // [2]
  static constexpr int BAR_KFULL = 2;                    // [KV_STAGES]
  static constexpr int BAR_KEMPTY = BAR_KFULL + KV_STAGES;
  static constexpr int BAR_VFULL = BAR_KEMPTY + KV_STAGES;
  static constexpr int BAR_VEMPTY = BAR_VFULL + KV_STAGES;
  static constexpr int BAR_SFULL = BAR_VEMPTY + KV_STAGES;   // [2]
  static constexpr int BAR_PFULL = BAR_SFULL + 2;            // [2]
  static constexpr int BAR_OFINAL = BAR_PFULL + 2;           // [2]
  static constexpr int NUM_BARS = BAR_OFINAL + 2;
  static constexpr int OFF_TMEM_SLOT = OFF_BAR + NUM_BARS * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM_SLOT + 16 + 1024;   // + alignment slack
  static constexpr int COL_S0 = 0, COL_S1 = 128, COL_O0 = 256, COL_O1 = 256 + D;
};

struct TcParams {
  void* o;
  float* l;
  float* m;
  int N, B, W, p, mode;
  float scale_log2;      // tau * log2(e)
};

__host__ __device__ inline int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

struct TileRange { int jlo, jhi; };

// key-tile stream of one CTA (a pair of 128-query tiles starting at q0)
__device__ __forceinline__ void tile_ranges(const TcParams& prm, int q0, int& kbase, int& nj, TileRange (&tr)[2]) {
  if (prm.mode == MODE_CIRCULANT) {
    kbase = floor_div(q0 - prm.p, 128) * 128;
    nj = 0;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int i0 = q0 + 128 * t;
      if (i0 < prm.N) {
        tr[t].jlo = floor_div(i0 - prm.p - kbase, 128);
        tr[t].jhi = floor_div(i0 + 127 - prm.p + prm.W - 1 - kbase, 128) + 1;
        nj = tr[t].jhi > nj ? tr[t].jhi : nj;
      } else { tr[t].jlo = 0; tr[t].jhi = 0; }
    }
  } else {
    kbase = 0;
    nj = (prm.N + 127) / 128;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
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

template <int D, int FMT>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_fwd_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
              const __grid_constant__ CUtensorMap tmv, const TcParams prm) {
  using C = Cfg<D>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = sbase + C::OFF_Q, sK = sbase + C::OFF_K, sV = sbase + C::OFF_V;
  const uint32_t bars = sbase + C::OFF_BAR;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  const uint32_t tmem_slot = sbase + C::OFF_TMEM_SLOT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256, b = blockIdx.y;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmq); prefetch_tensormap(&tmk); prefetch_tensormap(&tmv);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) mbar_init(bar(C::BAR_QFULL + i), 1);
    for (int i = 0; i < C::KV_STAGES; ++i) {
      mbar_init(bar(C::BAR_KFULL + i), 1); mbar_init(bar(C::BAR_KEMPTY + i), 1);
      mbar_init(bar(C::BAR_VFULL + i), 1); mbar_init(bar(C::BAR_VEMPTY + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(C::BAR_SFULL + i), 1);
      mbar_init(bar(C::BAR_PFULL + i), 128);
      mbar_init(bar(C::BAR_OFINAL + i), 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  int kbase, nj;
  TileRange tr[2];
  tile_ranges(prm, q0, kbase, nj, tr);

  if (warp < 4) {
    setmaxnreg_dec<48>();
    if (warp == 0 && lane == 0) {
      // ------------------------------------------------------------ TMA producer
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (tr[t].jhi > tr[t].jlo) {
          mbar_arrive_expect_tx(bar(C::BAR_QFULL + t), C::TILE_BYTES);
          tma_load_3d(sQ + t * C::TILE_BYTES, &tmq, bar(C::BAR_QFULL + t), q0 + 128 * t, 0, b);
          tma_load_3d(sQ + t * C::TILE_BYTES + C::BOX_BYTES, &tmq, bar(C::BAR_QFULL + t), q0 + 128 * t + 64, 0, b);
        }
      }
      for (int j = 0; j < nj; ++j) {
        const int s = j % C::KV_STAGES;
        const uint32_t ph = (uint32_t)(j / C::KV_STAGES) & 1u;
        const int tok = (prm.mode == MODE_CIRCULANT) ? (int)pmod(kbase + 128 * j, prm.N) : 128 * j;
        mbar_wait(bar(C::BAR_KEMPTY + s), ph ^ 1u);
        mbar_arrive_expect_tx(bar(C::BAR_KFULL + s), C::TILE_BYTES);
        tma_load_3d(sK + s * C::TILE_BYTES, &tmk, bar(C::BAR_KFULL + s), tok, 0, b);
        tma_load_3d(sK + s * C::TILE_BYTES + C::BOX_BYTES, &tmk, bar(C::BAR_KFULL + s), tok + 64, 0, b);
        mbar_wait(bar(C::BAR_VEMPTY + s), ph ^ 1u);
        mbar_arrive_expect_tx(bar(C::BAR_VFULL + s), C::TILE_BYTES);
        tma_load_3d(sV + s * C::TILE_BYTES, &tmv, bar(C::BAR_VFULL + s), tok, 0, b);
        tma_load_3d(sV + s * C::TILE_BYTES + C::BOX_BYTES, &tmv, bar(C::BAR_VFULL + s), tok + 64, 0, b);
      }
    } else if (warp == 1 && lane == 0) {
      // ------------------------------------------------------------ MMA issuer
      constexpr uint32_t idesc_qk = make_idesc_f16(FMT, FMT, 1, 1, 128, 128);   // A, B MN-major
      // P takes the input format: mixing A = f16 with B = bf16 in one kind::f16 MMA raises
      // "illegal instruction" on sm_100a (measured, round 1), so bf16 inputs mean bf16 P.
      constexpr uint32_t idesc_pv = make_idesc_f16(FMT, FMT, 0, 0, 128, D);     // A in TMEM, B K-major
      const uint32_t colS[2] = {tmem_base + C::COL_S0, tmem_base + C::COL_S1};
      const uint32_t colO[2] = {tmem_base + C::COL_O0, tmem_base + C::COL_O1};
      uint32_t pphase[2] = {0, 0};
      for (int j = 0; j <= nj; ++j) {
        bool k_waited = false, v_waited = false;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (j >= 1 && tr[t].jlo <= j - 1 && j - 1 < tr[t].jhi) {        // O_t += P_t(j-1) V(j-1)
            const int sv = (j - 1) % C::KV_STAGES;
            if (!v_waited) { mbar_wait(bar(C::BAR_VFULL + sv), (uint32_t)((j - 1) / C::KV_STAGES) & 1u); v_waited = true; }
            mbar_wait(bar(C::BAR_PFULL + t), pphase[t]); pphase[t] ^= 1u;
            tc_fence_after();
            const uint32_t vb = sV + sv * C::TILE_BYTES;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t bd = make_smem_desc_sw128(vb + (ks >> 2) * C::BOX_BYTES + (ks & 3) * 32, 16, 1024);
              mma_ts(colO[t], colS[t] + ks * 8, bd, idesc_pv, (j - 1 > tr[t].jlo || ks > 0) ? 1u : 0u);
            }
          }
          if (j < nj && tr[t].jlo <= j && j < tr[t].jhi) {                // S_t = Q_t K(j)^T
            const int sk = j % C::KV_STAGES;
            if (j == tr[t].jlo) mbar_wait(bar(C::BAR_QFULL + t), 0);
            if (!k_waited) { mbar_wait(bar(C::BAR_KFULL + sk), (uint32_t)(j / C::KV_STAGES) & 1u); k_waited = true; }
            tc_fence_after();
            const uint32_t qb = sQ + t * C::TILE_BYTES, kb = sK + sk * C::TILE_BYTES;
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks) {
              const uint64_t ad = make_smem_desc_sw128(qb + ks * 2048, C::BOX_BYTES, 1024);
              const uint64_t bd = make_smem_desc_sw128(kb + ks * 2048, C::BOX_BYTES, 1024);
              mma_ss(colS[t], ad, bd, idesc_qk, ks > 0 ? 1u : 0u);
            }
            tc_commit(bar(C::BAR_SFULL + t));
          }
        }
        if (j >= 1) tc_commit(bar(C::BAR_VEMPTY + (j - 1) % C::KV_STAGES));
        if (j < nj) tc_commit(bar(C::BAR_KEMPTY + j % C::KV_STAGES));
      }
      tc_commit(bar(C::BAR_OFINAL + 0));
      tc_commit(bar(C::BAR_OFINAL + 1));
    }
  } else {
    // -------------------------------------------------------------- softmax warpgroups
    setmaxnreg_inc<224>();
    const int t = (warp - 4) >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + (t == 0 ? C::COL_S0 : C::COL_S1);
    const uint32_t tO = tmem_base + lane_addr + (t == 0 ? C::COL_O0 : C::COL_O1);
    const int qi = q0 + 128 * t + row;                 // query token of this thread
    const float scale = prm.scale_log2;

    if (tr[t].jhi > tr[t].jlo) {
      float m_true = -INFINITY, m_used = -INFINITY, l_run = 0.f;
      uint32_t sphase = 0;
      for (int j = tr[t].jlo; j < tr[t].jhi; ++j) {
        mbar_wait(bar(C::BAR_SFULL + t), sphase); sphase ^= 1u;
        tc_fence_after();
        uint32_t s[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(tS + 32 * c, s[c]);
        tmem_wait_ld();

        // ---- mask (last dense tile / circulant band edges); lo <= col < hi stays
        int lo = 0, hi = 128;
        if (prm.mode == MODE_CIRCULANT) { lo = (qi - prm.p) - (kbase + 128 * j); hi = lo + prm.W; }
        else { hi = prm.N - 128 * j; }
        if (lo > 0 || hi < 128) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int col = 32 * c + i;
              if (col < lo || col >= hi) s[c][i] = 0xff800000u;   // -inf
            }
        }
        // ---- running max (thread-local: one thread owns one row)
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            mx = fmaxf(mx, fmaxf(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])));
        m_true = fmaxf(m_true, mx * scale);
        // ---- lazy rescale of O and l (warp-uniform decision; this warp owns its 32 TMEM lanes)
        const bool want = (m_true - m_used) > RESCALE_THRESHOLD;   // inf on first use
        if (__any_sync(0xffffffffu, want)) {
          const float alpha = (m_used == -INFINITY) ? 0.f : ex2(m_used - m_true);
          if (j > tr[t].jlo) {                        // O_t holds PV results only after the first tile
#pragma unroll 1
            for (int c = 0; c < D / 32; ++c) {
              uint32_t o[32];
              tmem_ld32(tO + 32 * c, o);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st32(tO + 32 * c, o);
            }
          }
          l_run *= alpha;
          m_used = m_true;
        }
        const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used;   // guard (-inf)-(-inf)
        // ---- P = exp2(s*scale - m) -> 16-bit, written over the first 64 columns of S
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2(fmaf(__uint_as_float(s[c][i]), scale, neg_m));
            const float p1 = ex2(fmaf(__uint_as_float(s[c][i + 1]), scale, neg_m));
            sum0 += p0; sum1 += p1;
            pk[i >> 1] = pack2<FMT>(p0, p1);
          }
          tmem_st16(tS + 16 * c, pk);
        }
        l_run += sum0 + sum1;
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bar(C::BAR_PFULL + t));
      }

      // ---- epilogue: O / l -> global (token-contiguous rows: a warp writes 32 consecutive tokens)
      mbar_wait(bar(C::BAR_OFINAL + t), 0);
      tc_fence_after();
      const float inv_l = 1.f / l_run;
      const bool in_range = qi < prm.N;
      if (FMT == 1) {
        __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(prm.o) + (size_t)b * D * prm.N + qi;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          uint32_t o[32];
          tmem_ld32(tO + 32 * c, o);
          tmem_wait_ld();
          if (in_range) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              ob[(size_t)(32 * c + i) * prm.N] = __float2bfloat16_rn(__uint_as_float(o[i]) * inv_l);
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
            for (int i = 0; i < 32; ++i)
              ob[(size_t)(32 * c + i) * prm.N] = __float2half_rn(__uint_as_float(o[i]) * inv_l);
          }
        }
      }
      if (in_range) {
        // l = sum exp(s - m_true), m = max s (natural-log domain), reference src/dense.jl:12-18
        prm.l[(size_t)b * prm.N + qi] = l_run * ex2(m_used - m_true);
        prm.m[(size_t)b * prm.N + qi] = m_true * LN2;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// tensor map over one [B][D][N] tensor: box = 64 tokens x D channels x 1 batch, SWIZZLE_128B
int make_tmap(CUtensorMap* tm, const void* base, int dtype, long long N, int D, long long B) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return FA_ERR_CUDA; }
  const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)D, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)N * 2, (cuuint64_t)N * D * 2};
  const cuuint32_t box[3] = {64, (cuuint32_t)D, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = dtype == FA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(tm, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return FA_ERR_CUDA; }
  return FA_OK;
}

template <int D, int FMT>
int launch_tc(const Geo& g, const FwdArgs& a, int dtype, cudaStream_t st) {
  CUtensorMap tmq, tmk, tmv;
  int rc;
  if ((rc = make_tmap(&tmq, a.q, dtype, g.N, D, g.B))) return rc;
  if ((rc = make_tmap(&tmk, a.k, dtype, g.N, D, g.B))) return rc;
  if ((rc = make_tmap(&tmv, a.v, dtype, g.N, D, g.B))) return rc;
  TcParams prm;
  prm.o = a.o; prm.l = a.l; prm.m = a.m;
  prm.N = (int)g.N; prm.B = (int)g.B; prm.W = g.W; prm.p = g.p; prm.mode = g.mode;
  prm.scale_log2 = g.tau * LOG2E;
  auto kern = tc_fwd_kernel<D, FMT>;
  FA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<D>::SMEM_BYTES));
  const dim3 grid((unsigned)((g.N + 255) / 256), (unsigned)g.B);
  kern<<<grid, TC_THREADS, Cfg<D>::SMEM_BYTES, st>>>(tmq, tmk, tmv, prm);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

}  // namespace

int make_tmap_public(CUtensorMap* tm, const void* base, int dtype, long long N, int D, long long B) {
  return make_tmap(tm, base, dtype, N, D, B);
}

bool tc_fwd_supported(const Geo& g, int dtype) {
  if (dtype != FA_BF16 && dtype != FA_F16) return false;
  if (g.mode != MODE_DENSE && g.mode != MODE_CIRCULANT) return false;
  if (g.d != g.dv || (g.d != 64 && g.d != 128)) return false;
  if (g.N % 8 != 0 || g.N < 8 || g.N > 0x3fffffff) return false;    // TMA: 16-byte global strides
  if (g.B > 65535) return false;                                     // gridDim.y
  if (g.mode == MODE_CIRCULANT && (g.N % 128 != 0)) return false;    // tile-aligned wrap-around
  return true;
}

int tc_fwd(const Geo& g, const FwdArgs& a, int dtype, cudaStream_t st) {
  if (!tc_fwd_supported(g, dtype)) { set_error("tc_fwd: unsupported configuration"); return FA_ERR_UNSUPPORTED; }
  if ((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v)) & 15) {
    set_error("tc_fwd: q/k/v must be 16-byte aligned"); return FA_ERR_INVALID;
  }
  const int fmt = dtype == FA_BF16 ? 1 : 0;
  if (g.d == 128) return fmt ? launch_tc<128, 1>(g, a, dtype, st) : launch_tc<128, 0>(g, a, dtype, st);
  return fmt ? launch_tc<64, 1>(g, a, dtype, st) : launch_tc<64, 0>(g, a, dtype, st);
}

}  // namespace fa
