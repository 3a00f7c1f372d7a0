// fa_simt.cu -- exact-fp32 (FFMA, no TF32) flash-attention kernels for dense, circulant and
// windowed attention, forward and backward, any dtype in/out (fp32 math).
//
// This is the "TF32-off fp32 path" of BASELINE.json:north_star (parity 1e-5) and the
// generic fallback for shapes the tcgen05 kernels (fa_tc_fwd.cu) do not cover.  One kernel
// template serves all three attention patterns through the slot/token geometry of
// fa_common.cuh: tiles of 64 query slots x 64 key slots, 256 threads as a 16x16 grid with a
// 4x4 register micro-tile for S = Q K^T and a 4 x (dv/16) micro-tile for O += P V.
//
// Reference semantics restated (never copied):
//   dense     src/dense.jl:21-102        (online softmax; tau = 1/sqrt(d); returns l, m)
//   circulant src/circulant.jl:9-118     (keys of query i: mod(i-p+t, N), t = 0..W-1)
//   windowed  src/windowed.jl:3-23 + src/utils.jl:36-54 (zero-padded unfold, fold-sum / count)
//   backward  src_cpp/FlashAttention.cpp:194-252 (P recomputed from l, m)
#include "fa_common.cuh"

namespace fa {
namespace {

constexpr int BM = 64;        // owner-side slots per CTA
constexpr int BN = 64;        // other-side slots per inner tile
constexpr int NT = 256;       // threads per CTA (16 x 16)
constexpr int LD = BN + 4;    // padded smem row (floats): float4-aligned, conflict-free

struct Tile {
  long long prob;   // problem index (batch element, or window + L*batch)
  long long b;      // batch element
  long long win;    // window index inside the batch element (windowed)
  int t;            // owner tile index inside the problem
  long long nslots; // slots per problem
};

__device__ __forceinline__ Tile decode_tile(const Geo& g) {
  Tile tl;
  tl.nslots = geo_slots(g);
  const long long tpp = (tl.nslots + BM - 1) / BM;
  const long long bid = blockIdx.x;
  tl.prob = bid / tpp;
  tl.t = (int)(bid % tpp);
  if (g.mode == MODE_WINDOWED) { tl.b = tl.prob / g.L; tl.win = tl.prob % g.L; }
  else { tl.b = tl.prob; tl.win = 0; }
  return tl;
}

// owner-side slot -> token (or -1: zero pad / beyond the problem)
__device__ __forceinline__ int owner_token(const Geo& g, const Tile& tl, int r) {
  const long long slot = (long long)tl.t * BM + r;
  if (slot >= tl.nslots) return -1;
  if (g.mode == MODE_WINDOWED) return (int)window_slot_token(g, tl.win, (int)slot);
  return (int)slot;
}

// number of inner tiles and the unwrapped base index of the other side
template <bool OWNER_IS_QUERY>
__device__ __forceinline__ void other_range(const Geo& g, const Tile& tl, long long& base, int& ntiles) {
  if (g.mode == MODE_CIRCULANT) {
    const long long own0 = (long long)tl.t * BM;
    base = OWNER_IS_QUERY ? own0 - g.p : own0 - (g.W - 1 - g.p);
    ntiles = (BM + g.W - 1 + BN - 1) / BN;
  } else {
    base = 0;
    ntiles = (int)((tl.nslots + BN - 1) / BN);
  }
}

// other-side slot -> token (or -1)
__device__ __forceinline__ int other_token(const Geo& g, const Tile& tl, long long base, int it, int c) {
  const long long u = base + (long long)it * BN + c;
  if (g.mode == MODE_CIRCULANT) {
    if ((long long)it * BN + c >= BM + g.W - 1) return -1;
    return (int)pmod(u, g.N);
  }
  if (u >= tl.nslots) return -1;
  if (g.mode == MODE_WINDOWED) return (int)window_slot_token(g, tl.win, (int)u);
  return (int)u;
}

// is the (owner row, other col) pair inside the attention pattern?
template <bool OWNER_IS_QUERY>
__device__ __forceinline__ bool pair_valid(const Geo& g, const Tile& tl, long long base, int it, int r, int c) {
  const long long own = (long long)tl.t * BM + r;
  const long long u = base + (long long)it * BN + c;
  if (g.mode == MODE_CIRCULANT) {
    const long long dj = OWNER_IS_QUERY ? (u - (own - g.p)) : (own - (u - g.p));
    return dj >= 0 && dj < g.W && own < g.N;
  }
  return u < tl.nslots;
}

// cooperative gather of a [ch][64] tile (token-contiguous global rows) into padded smem, zero
// for token -1.  `scale_cnt` divides by the fold count (dYw = window(dY ./ count), A.5.2).
template <typename T, bool SCALE_CNT>
__device__ __forceinline__ void load_tile(float* dst, const T* __restrict__ base, const int* tok,
                                          int ch, int ch_pad, long long N, const Geo& g) {
  for (int idx = threadIdx.x; idx < ch_pad * 64; idx += NT) {
    const int c = idx >> 6, s = idx & 63;
    const int t = tok[s];
    float v = 0.f;
    if (c < ch && t >= 0) {
      v = to_f32<T>(base[(long long)c * N + t]);
      if (SCALE_CNT) v /= (float)window_count_at(g, t);
    }
    dst[c * LD + s] = v;
  }
}

__device__ __forceinline__ float half_warp_max(float x) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}
__device__ __forceinline__ float half_warp_sum(float x) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// s[a][b] += sum_k A[k][ty*4+a] * Bm[k][tx*4+b]
__device__ __forceinline__ void tile_outer(float (&s)[4][4], const float* A, const float* Bm, int kdim, int ty, int tx) {
#pragma unroll 4
  for (int kk = 0; kk < kdim; ++kk) {
    const float4 av = *reinterpret_cast<const float4*>(&A[kk * LD + ty * 4]);
    const float4 bv = *reinterpret_cast<const float4*>(&Bm[kk * LD + tx * 4]);
    const float aa[4] = {av.x, av.y, av.z, av.w};
    const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) s[a][b] = fmaf(aa[a], bb[b], s[a][b]);
  }
}

// acc[a][cc] += sum_j P[ty*4+a][j] * Cm[tx+16cc][j]
template <int CPT>
__device__ __forceinline__ void tile_accum(float (&acc)[4][CPT], const float* P, const float* Cm, int ty, int tx) {
#pragma unroll 2
  for (int j = 0; j < BN; j += 4) {
    float4 pv[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) pv[a] = *reinterpret_cast<const float4*>(&P[(ty * 4 + a) * LD + j]);
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) {
      const float4 vv = *reinterpret_cast<const float4*>(&Cm[(tx + 16 * cc) * LD + j]);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        acc[a][cc] = fmaf(pv[a].x, vv.x, acc[a][cc]);
        acc[a][cc] = fmaf(pv[a].y, vv.y, acc[a][cc]);
        acc[a][cc] = fmaf(pv[a].z, vv.z, acc[a][cc]);
        acc[a][cc] = fmaf(pv[a].w, vv.w, acc[a][cc]);
      }
    }
  }
}

// stage acc (owner rows x channels) through smem and store/fold it token-contiguously
template <typename T, int CPT>
__device__ __forceinline__ void store_owner_tile(float (&acc)[4][CPT], float* stage, const int* otok,
                                                 int ch, long long N, T* out_base, float* acc_base,
                                                 bool use_atomic, int ty, int tx) {
  __syncthreads();
#pragma unroll
  for (int cc = 0; cc < CPT; ++cc)
    *reinterpret_cast<float4*>(&stage[(tx + 16 * cc) * LD + ty * 4]) =
        make_float4(acc[0][cc], acc[1][cc], acc[2][cc], acc[3][cc]);
  __syncthreads();
  for (int idx = threadIdx.x; idx < ch * 64; idx += NT) {
    const int c = idx >> 6, r = idx & 63;
    const int t = otok[r];
    if (t < 0) continue;
    const float v = stage[c * LD + r];
    if (use_atomic) atomicAdd(&acc_base[(long long)c * N + t], v);
    else out_base[(long long)c * N + t] = from_f32<T>(v);
  }
}

// =========================================================================================
// forward
// =========================================================================================
template <typename T, int CPT>
__global__ void __launch_bounds__(NT) simt_fwd_kernel(const Geo g, const FwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int d = g.d, dv = g.dv, dvp = 16 * CPT;
  float* Qs = smem;                 // [d][LD]
  float* Ks = Qs + d * LD;          // [d][LD]
  float* Vs = Ks + d * LD;          // [dvp][LD]
  float* Ps = Vs + dvp * LD;        // [BM][LD]
  int* qtok = reinterpret_cast<int*>(Ps + BM * LD);
  int* ktok = qtok + BM;

  const Tile tl = decode_tile(g);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const T* qb = static_cast<const T*>(a.q) + tl.b * (long long)d * g.N;
  const T* kb = static_cast<const T*>(a.k) + tl.b * (long long)d * g.N;
  const T* vb = static_cast<const T*>(a.v) + tl.b * (long long)dv * g.N;

  if (tid < BM) qtok[tid] = owner_token(g, tl, tid);
  __syncthreads();
  load_tile<T, false>(Qs, qb, qtok, d, d, g.N, g);

  long long base; int ntiles;
  other_range<true>(g, tl, base, ntiles);

  float m_run[4], l_run[4], acc[4][CPT];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    m_run[r] = -INFINITY; l_run[r] = 0.f;
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) acc[r][cc] = 0.f;
  }

  for (int it = 0; it < ntiles; ++it) {
    __syncthreads();                                   // previous tile fully consumed
    if (tid < BN) ktok[tid] = other_token(g, tl, base, it, tid);
    __syncthreads();
    load_tile<T, false>(Ks, kb, ktok, d, d, g.N, g);
    load_tile<T, false>(Vs, vb, ktok, dv, dvp, g.N, g);
    __syncthreads();

    float s[4][4] = {};
    tile_outer(s, Qs, Ks, d, ty, tx);

#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s[r][c] = pair_valid<true>(g, tl, base, it, ty * 4 + r, tx * 4 + c) ? s[r][c] * g.tau : -INFINITY;
        mx = fmaxf(mx, s[r][c]);
      }
      mx = half_warp_max(mx);
      const float m_new = fmaxf(m_run[r], mx);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;   // guard (-inf)-(-inf), SURVEY A.1
      const float alpha = expf(m_run[r] - m_safe);
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) { s[r][c] = expf(s[r][c] - m_safe); rs += s[r][c]; }
      rs = half_warp_sum(rs);
      l_run[r] = l_run[r] * alpha + rs;
      m_run[r] = m_new;
#pragma unroll
      for (int cc = 0; cc < CPT; ++cc) acc[r][cc] *= alpha;
      *reinterpret_cast<float4*>(&Ps[(ty * 4 + r) * LD + tx * 4]) = make_float4(s[r][0], s[r][1], s[r][2], s[r][3]);
    }
    __syncthreads();
    tile_accum<CPT>(acc, Ps, Vs, ty, tx);
  }

  // ---- epilogue: normalise, write l/m, store or fold O
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float inv = 1.f / l_run[r];
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) acc[r][cc] *= inv;
    const long long slot = (long long)tl.t * BM + ty * 4 + r;
    if (tx == 0 && slot < tl.nslots) {
      a.l[tl.prob * tl.nslots + slot] = l_run[r];
      a.m[tl.prob * tl.nslots + slot] = m_run[r];
    }
  }
  const bool atomic = (g.mode == MODE_WINDOWED) && g.overlap;
  T* ob = static_cast<T*>(a.o) + tl.b * (long long)dv * g.N;
  float* ab = a.acc ? a.acc + tl.b * (long long)dv * g.N : nullptr;
  store_owner_tile<T, CPT>(acc, Vs, qtok, dv, g.N, ob, ab, atomic, ty, tx);
}

// =========================================================================================
// backward.  OWNER_IS_QUERY: CTA owns 64 queries, loops over keys, produces dQ (or, with
// DELTA_ONLY, D_i = sum_j P_ij dP_ij for the windowed case where the per-window O is not
// materialised).  !OWNER_IS_QUERY: CTA owns 64 keys, loops over queries, produces dK and dV.
// Two deterministic passes instead of the reference's racy shared accumulation
// (src_cpp/FlashAttention.cpp:299-312).
// =========================================================================================
template <typename T, int CPT, bool OWNER_IS_QUERY, bool DELTA_ONLY>
__global__ void __launch_bounds__(NT) simt_bwd_kernel(const Geo g, const BwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int d = g.d, dv = g.dv, cp = 16 * CPT;
  float* Rs  = smem;                  // owner  [d][LD]    Q (query owner) | K (key owner)
  float* Rvs = Rs + cp * LD;          // owner  [cp][LD]   dO             | V
  float* Cs  = Rvs + cp * LD;         // other  [cp][LD]   K              | Q
  float* Cvs = Cs + cp * LD;          // other  [cp][LD]   V              | dO
  float* Ps  = Cvs + cp * LD;         // [BM][LD]
  float* dSs = Ps + BM * LD;          // [BM][LD]
  float* st_m = dSs + BM * LD;        // other-side stats when the owner is the key
  float* st_l = st_m + BN;
  float* st_d = st_l + BN;
  int* otok = reinterpret_cast<int*>(st_d + BN);
  int* ctok = otok + BM;

  const Tile tl = decode_tile(g);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const bool win = g.mode == MODE_WINDOWED;
  const long long offd = tl.b * (long long)d * g.N, offv = tl.b * (long long)dv * g.N;
  const T* qb = static_cast<const T*>(a.q) + offd;
  const T* kb = static_cast<const T*>(a.k) + offd;
  const T* vb = static_cast<const T*>(a.v) + offv;
  const T* gb = static_cast<const T*>(a.d_o) + offv;
  const long long sbase = tl.prob * tl.nslots;        // stats base of this problem

  if (tid < BM) otok[tid] = owner_token(g, tl, tid);
  __syncthreads();
  if (OWNER_IS_QUERY) {
    load_tile<T, false>(Rs, qb, otok, d, cp, g.N, g);
    if (win) load_tile<T, true>(Rvs, gb, otok, dv, cp, g.N, g);
    else     load_tile<T, false>(Rvs, gb, otok, dv, cp, g.N, g);
  } else {
    load_tile<T, false>(Rs, kb, otok, d, cp, g.N, g);
    load_tile<T, false>(Rvs, vb, otok, dv, cp, g.N, g);
  }

  long long base; int ntiles;
  other_range<OWNER_IS_QUERY>(g, tl, base, ntiles);

  float row_m[4], row_l[4], row_d[4], dacc[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    row_m[r] = 0.f; row_l[r] = 1.f; row_d[r] = 0.f; dacc[r] = 0.f;
    if (OWNER_IS_QUERY) {
      const long long slot = (long long)tl.t * BM + ty * 4 + r;
      if (slot < tl.nslots) {
        row_m[r] = a.m[sbase + slot];
        row_l[r] = a.l[sbase + slot];
        if (!DELTA_ONLY) row_d[r] = a.delta[sbase + slot];
      }
    }
  }
  float acc1[4][CPT], acc2[4][CPT];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) { acc1[r][cc] = 0.f; acc2[r][cc] = 0.f; }

  for (int it = 0; it < ntiles; ++it) {
    __syncthreads();
    if (tid < BN) {
      ctok[tid] = other_token(g, tl, base, it, tid);
      if (!OWNER_IS_QUERY) {
        // stats of the other-side (query) slots
        const long long u = base + (long long)it * BN + tid;
        long long slot = -1;
        if (g.mode == MODE_CIRCULANT) { if ((long long)it * BN + tid < BM + g.W - 1) slot = pmod(u, g.N); }
        else if (u < tl.nslots) slot = u;
        st_m[tid] = slot >= 0 ? a.m[sbase + slot] : 0.f;
        st_l[tid] = slot >= 0 ? a.l[sbase + slot] : 1.f;
        st_d[tid] = slot >= 0 ? a.delta[sbase + slot] : 0.f;
      }
    }
    __syncthreads();
    if (OWNER_IS_QUERY) {
      load_tile<T, false>(Cs, kb, ctok, d, cp, g.N, g);
      load_tile<T, false>(Cvs, vb, ctok, dv, cp, g.N, g);
    } else {
      load_tile<T, false>(Cs, qb, ctok, d, cp, g.N, g);
      if (win) load_tile<T, true>(Cvs, gb, ctok, dv, cp, g.N, g);
      else     load_tile<T, false>(Cvs, gb, ctok, dv, cp, g.N, g);
    }
    __syncthreads();

    float s[4][4] = {}, dp[4][4] = {};
    tile_outer(s, Rs, Cs, d, ty, tx);        // S  (owner rows x other cols)
    tile_outer(dp, Rvs, Cvs, dv, ty, tx);    // dP = dO V^T (same orientation)

#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float pr[4], ds[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = tx * 4 + c;
        const bool ok = pair_valid<OWNER_IS_QUERY>(g, tl, base, it, ty * 4 + r, col);
        const float mm = OWNER_IS_QUERY ? row_m[r] : st_m[col];
        const float ll = OWNER_IS_QUERY ? row_l[r] : st_l[col];
        const float dd = OWNER_IS_QUERY ? row_d[r] : st_d[col];
        const float p = ok ? expf(s[r][c] * g.tau - mm) / ll : 0.f;    // cpp:239-240
        pr[c] = p;
        ds[c] = p * (dp[r][c] - dd) * g.tau;                           // cpp:244 (tau folded in)
        if (DELTA_ONLY) dacc[r] += p * dp[r][c];
      }
      if (!DELTA_ONLY) {
        *reinterpret_cast<float4*>(&dSs[(ty * 4 + r) * LD + tx * 4]) = make_float4(ds[0], ds[1], ds[2], ds[3]);
        if (!OWNER_IS_QUERY)
          *reinterpret_cast<float4*>(&Ps[(ty * 4 + r) * LD + tx * 4]) = make_float4(pr[0], pr[1], pr[2], pr[3]);
      }
    }
    if (!DELTA_ONLY) {
      __syncthreads();
      tile_accum<CPT>(acc1, dSs, Cs, ty, tx);                    // dQ += dS K   | dK += dS^T Q
      if (!OWNER_IS_QUERY) tile_accum<CPT>(acc2, Ps, Cvs, ty, tx);   // dV += P^T dO
    }
  }

  if (DELTA_ONLY) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float t = half_warp_sum(dacc[r]);
      const long long slot = (long long)tl.t * BM + ty * 4 + r;
      if (tx == 0 && slot < tl.nslots) a.delta[sbase + slot] = t;
    }
    return;
  }
  const bool atomic = win && g.overlap;
  if (OWNER_IS_QUERY) {
    store_owner_tile<T, CPT>(acc1, Cs, otok, d, g.N, static_cast<T*>(a.dq) + offd,
                             a.aq ? a.aq + offd : nullptr, atomic, ty, tx);
  } else {
    store_owner_tile<T, CPT>(acc1, Cs, otok, d, g.N, static_cast<T*>(a.dk) + offd,
                             a.ak ? a.ak + offd : nullptr, atomic, ty, tx);
    store_owner_tile<T, CPT>(acc2, Cvs, otok, dv, g.N, static_cast<T*>(a.dv) + offv,
                             a.av ? a.av + offv : nullptr, atomic, ty, tx);
  }
}

// D_i = sum_c dO[i][c] * O[i][c]   (src_cpp/FlashAttention.cpp:243), dense / circulant
template <typename T>
__global__ void delta_kernel(const T* __restrict__ o, const T* __restrict__ d_o, float* __restrict__ delta,
                             long long N, int dv, long long B) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N * B) return;
  const long long b = i / N, n = i % N;
  const T* ob = o + b * dv * N + n;
  const T* gb = d_o + b * dv * N + n;
  float acc = 0.f;
  for (int c = 0; c < dv; ++c) acc = fmaf(to_f32<T>(ob[(long long)c * N]), to_f32<T>(gb[(long long)c * N]), acc);
  delta[i] = acc;
}

// y = acc (/ count) -> T, NaN where count == 0 and divide != 0 (0/0 of src/windowed.jl:19)
template <typename T>
__global__ void fold_finalize_kernel(const Geo g, const float* __restrict__ acc, T* __restrict__ y,
                                     int channels, int divide) {
  const long long total = g.N * channels * g.B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long tok = i % g.N;
    float v = acc[i];
    if (divide) v = v / (float)window_count_at(g, tok);
    y[i] = from_f32<T>(v);
  }
}

// slab of a larger volume (planes [plane_lo, plane_lo + nplanes) of the slowest spatial dim): y = acc ./ count with the
// count of the WHOLE volume `g` (src/windowed.jl:16-19); acc, y :: (slab tokens, channels, B)
template <typename T>
__global__ void slab_divide_kernel(const Geo g, const float* __restrict__ acc, T* __restrict__ y, int channels,
                                   long long plane_lo, long long slab_tokens) {
  const long long total = slab_tokens * channels * g.B;
  long long plane_tokens = 1;
  for (int k = 0; k < g.nd - 1; ++k) plane_tokens *= g.s[k];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long tok = i % slab_tokens;
    y[i] = from_f32<T>(acc[i] / (float)window_count_at(g, tok + plane_lo * plane_tokens));
  }
}

template <typename T>
__global__ void fill_uncovered_kernel(const Geo g, T* __restrict__ y, int channels) {
  const long long total = g.N * channels * g.B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    if (window_count_at(g, i % g.N) == 0) y[i] = from_f32<T>(__int_as_float(0x7fc00000));
  }
}

// standalone unfold: xw[(kappa, c, w, b)] = x[(token(kappa, w), c, b)] or 0   (src/utils.jl:36-44)
template <typename T>
__global__ void window_gather_kernel(const Geo g, const T* __restrict__ x, T* __restrict__ xw) {
  const long long total = (long long)g.WD * g.d * g.L * g.B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int kap = (int)(r % g.WD); r /= g.WD;
    const int c = (int)(r % g.d);    r /= g.d;
    const long long w = r % g.L;     const long long b = r / g.L;
    const long long t = window_slot_token(g, w, kap);
    xw[i] = t >= 0 ? x[(b * g.d + c) * g.N + t] : from_f32<T>(0.f);
  }
}

// standalone fold: x[(tok, c, b)] = sum over (kappa, w) reading tok of xw   (src/utils.jl:46-54)
// gather form (deterministic, no atomics): enumerate the windows covering tok.
template <typename T>
__global__ void window_scatter_kernel(const Geo g, const T* __restrict__ xw, T* __restrict__ x, int divide) {
  const long long total = g.N * g.d * g.B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    long long tok = r % g.N; r /= g.N;
    const int c = (int)(r % g.d); const long long b = r / g.d;
    int pos[3] = {0, 0, 0}, wlo[3] = {0, 0, 0}, wn[3] = {1, 1, 1};
    bool any = true;
    for (int k = 0; k < g.nd; ++k) {
      pos[k] = (int)(tok % g.s[k]); tok /= g.s[k];
      const int av = pos[k] + g.padv[k];
      int wmax = av / g.stride; if (wmax > g.o[k] - 1) wmax = g.o[k] - 1;
      const int lo = av - g.W + 1;
      wlo[k] = lo <= 0 ? 0 : (lo + g.stride - 1) / g.stride;
      wn[k] = wmax - wlo[k] + 1;
      if (wn[k] <= 0) any = false;
    }
    float sum = 0.f;
    if (any) {
      for (int i2 = 0; i2 < wn[2]; ++i2)
        for (int i1 = 0; i1 < wn[1]; ++i1)
          for (int i0 = 0; i0 < wn[0]; ++i0) {
            const int wi[3] = {wlo[0] + i0, wlo[1] + i1, wlo[2] + i2};
            long long w = 0, wm = 1; int kap = 0, km = 1;
            for (int k = 0; k < g.nd; ++k) {
              w += wi[k] * wm; wm *= g.o[k];
              kap += (pos[k] + g.padv[k] - wi[k] * g.stride) * km; km *= g.W;
            }
            sum += to_f32<T>(xw[((b * g.L + w) * g.d + c) * g.WD + kap]);
          }
    }
    if (divide) {                                   // windowed_fa: y = fold(yw) ./ count (src/windowed.jl:16-19); 0/0 = NaN
      int cnt = 1;
      for (int k = 0; k < g.nd; ++k) cnt *= wn[k] > 0 ? wn[k] : 0;
      sum = sum / (float)cnt;
    }
    x[i] = from_f32<T>(sum);
  }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) { set_error("simt kernel needs %zu B shared memory (> 227 KB): d/dv too large", bytes); return FA_ERR_UNSUPPORTED; }
  FA_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return FA_OK;
}

inline int pick_cpt(int ch) { return ch <= 16 ? 1 : ch <= 64 ? 4 : ch <= 128 ? 8 : 0; }

template <typename T, int CPT>
int launch_fwd(const Geo& g, const FwdArgs& a, cudaStream_t st) {
  const size_t smem = ((size_t)2 * g.d * LD + 16 * CPT * LD + BM * LD) * 4 + 2 * 64 * 4;
  int rc = set_smem(simt_fwd_kernel<T, CPT>, smem);
  if (rc) return rc;
  const long long tpp = (geo_slots(g) + BM - 1) / BM;
  const long long grid = tpp * geo_problems(g);
  if (grid <= 0 || grid > 0x7fffffffLL) { set_error("grid size %lld out of range", grid); return FA_ERR_INVALID; }
  simt_fwd_kernel<T, CPT><<<(unsigned)grid, NT, smem, st>>>(g, a);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

template <typename T>
int dispatch_fwd(const Geo& g, const FwdArgs& a, cudaStream_t st) {
  switch (pick_cpt(g.dv)) {
    case 1: return launch_fwd<T, 1>(g, a, st);
    case 4: return launch_fwd<T, 4>(g, a, st);
    case 8: return launch_fwd<T, 8>(g, a, st);
  }
  set_error("simt forward supports dv <= 128 (got %d)", g.dv);
  return FA_ERR_UNSUPPORTED;
}

template <typename T, int CPT, bool OQ, bool DO>
int launch_bwd_pass(const Geo& g, const BwdArgs& a, cudaStream_t st) {
  const size_t smem = ((size_t)4 * 16 * CPT * LD + 2 * BM * LD + 3 * BN) * 4 + 2 * 64 * 4;
  int rc = set_smem(simt_bwd_kernel<T, CPT, OQ, DO>, smem);
  if (rc) return rc;
  const long long tpp = (geo_slots(g) + BM - 1) / BM;
  const long long grid = tpp * geo_problems(g);
  if (grid <= 0 || grid > 0x7fffffffLL) { set_error("grid size %lld out of range", grid); return FA_ERR_INVALID; }
  simt_bwd_kernel<T, CPT, OQ, DO><<<(unsigned)grid, NT, smem, st>>>(g, a);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

template <typename T, int CPT>
int launch_bwd(const Geo& g, const BwdArgs& a, cudaStream_t st) {
  int rc;
  if (g.mode == MODE_WINDOWED) {
    if ((rc = launch_bwd_pass<T, CPT, true, true>(g, a, st))) return rc;      // delta = rowsum(P o dP)
  } else {
    const long long n = g.N * g.B;
    delta_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        static_cast<const T*>(a.o), static_cast<const T*>(a.d_o), a.delta, g.N, g.dv, g.B);
    FA_CUDA_TRY(cudaGetLastError());
  }
  if ((rc = launch_bwd_pass<T, CPT, true, false>(g, a, st))) return rc;       // dQ
  return launch_bwd_pass<T, CPT, false, false>(g, a, st);                     // dK, dV
}

template <typename T>
int dispatch_bwd(const Geo& g, const BwdArgs& a, cudaStream_t st) {
  const int ch = g.d > g.dv ? g.d : g.dv;
  switch (pick_cpt(ch)) {
    case 1: return launch_bwd<T, 1>(g, a, st);
    case 4: return launch_bwd<T, 4>(g, a, st);
    case 8: return launch_bwd<T, 8>(g, a, st);
  }
  set_error("simt backward supports d, dv <= 128 (got %d, %d)", g.d, g.dv);
  return FA_ERR_UNSUPPORTED;
}

inline unsigned ew_grid(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = 148LL * 16;      // grid-stride loops: a few waves over the 148 SMs
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

#define FA_DISPATCH_DTYPE(dtype, CALL)                                   \
  switch (dtype) {                                                       \
    case FA_F32:  { using T = float;          return CALL; }             \
    case FA_F16:  { using T = __half;         return CALL; }             \
    case FA_BF16: { using T = __nv_bfloat16;  return CALL; }             \
    default: set_error("unknown dtype %d", dtype); return FA_ERR_INVALID; \
  }

int simt_fwd(const Geo& g, const FwdArgs& a, int dtype, cudaStream_t st) {
  FA_DISPATCH_DTYPE(dtype, dispatch_fwd<T>(g, a, st));
}
int simt_bwd(const Geo& g, const BwdArgs& a, int dtype, cudaStream_t st) {
  FA_DISPATCH_DTYPE(dtype, dispatch_bwd<T>(g, a, st));
}

template <typename T>
static int fold_finalize_t(const Geo& g, const float* acc, void* y, int ch, int divide, cudaStream_t st) {
  fold_finalize_kernel<T><<<ew_grid(g.N * ch * g.B), 256, 0, st>>>(g, acc, static_cast<T*>(y), ch, divide);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}
int fold_finalize(const Geo& g, const float* acc, void* y, int ch, int dtype, int divide, cudaStream_t st) {
  FA_DISPATCH_DTYPE(dtype, fold_finalize_t<T>(g, acc, y, ch, divide, st));
}

template <typename T>
static int fill_uncovered_t(const Geo& g, void* y, int ch, cudaStream_t st) {
  fill_uncovered_kernel<T><<<ew_grid(g.N * ch * g.B), 256, 0, st>>>(g, static_cast<T*>(y), ch);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}
template <typename T>
static int slab_divide_t(const Geo& g, const float* acc, void* y, int ch, long long plane_lo, long long slab_tokens, cudaStream_t st) {
  slab_divide_kernel<T><<<ew_grid(slab_tokens * ch * g.B), 256, 0, st>>>(g, acc, static_cast<T*>(y), ch, plane_lo, slab_tokens);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}
int slab_divide(const Geo& g, const float* acc, void* y, int ch, int dtype, long long plane_lo, long long slab_tokens, cudaStream_t st) {
  FA_DISPATCH_DTYPE(dtype, slab_divide_t<T>(g, acc, y, ch, plane_lo, slab_tokens, st));
}
int fill_uncovered_nan(const Geo& g, void* y, int ch, int dtype, cudaStream_t st) {
  FA_DISPATCH_DTYPE(dtype, fill_uncovered_t<T>(g, y, ch, st));
}

template <typename T>
static int window_gather_t(const Geo& g, const void* x, void* xw, cudaStream_t st) {
  window_gather_kernel<T><<<ew_grid((long long)g.WD * g.d * g.L * g.B), 256, 0, st>>>(
      g, static_cast<const T*>(x), static_cast<T*>(xw));
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}
int window_gather(const Geo& g, const void* x, void* xw, int dtype, cudaStream_t st) {
  FA_DISPATCH_DTYPE(dtype, window_gather_t<T>(g, x, xw, st));
}

template <typename T>
static int window_scatter_t(const Geo& g, const void* xw, void* x, int divide, cudaStream_t st) {
  window_scatter_kernel<T><<<ew_grid(g.N * g.d * g.B), 256, 0, st>>>(
      g, static_cast<const T*>(xw), static_cast<T*>(x), divide);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}
int window_scatter(const Geo& g, const void* xw, void* x, int dtype, cudaStream_t st, int divide) {
  FA_DISPATCH_DTYPE(dtype, window_scatter_t<T>(g, xw, x, divide, st));
}


// element-wise dtype conversion (round-to-nearest-even), grid-stride, 128-bit accesses where aligned
template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = from_f32<TO>(to_f32<TI>(in[i]));
}
template <typename TI, typename TO>
static int cast_t(const void* in, void* out, long long n, cudaStream_t st) {
  cast_kernel<TI, TO><<<ew_grid(n), 256, 0, st>>>(static_cast<const TI*>(in), static_cast<TO*>(out), n);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}
int cast_launch(const void* in, void* out, long long n, int from, int to, cudaStream_t st) {
  if (from == FA_F32 && to == FA_BF16) return cast_t<float, __nv_bfloat16>(in, out, n, st);
  if (from == FA_F32 && to == FA_F16) return cast_t<float, __half>(in, out, n, st);
  if (from == FA_BF16 && to == FA_F32) return cast_t<__nv_bfloat16, float>(in, out, n, st);
  if (from == FA_F16 && to == FA_F32) return cast_t<__half, float>(in, out, n, st);
  set_error("fa_cast: unsupported dtype pair (%d -> %d)", from, to);
  return FA_ERR_UNSUPPORTED;
}

}  // namespace fa
