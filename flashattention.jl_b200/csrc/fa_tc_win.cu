// fa_tc_win.cu -- tcgen05 windowed attention for sm_100a: windowed_fa / block_fa
// (reference src/windowed.jl:1-23) with window / unwindow (src/utils.jl:36-54) fused in.
//
// The reference materialises three unfolds, a permutedims, a dense attention over (W^D, d, L*B),
// a fold and a ones-unfold-fold for the divisor.  Here a window is (part of) one tensor-core tile:
//   * G = floor(128 / W^D) consecutive windows are packed into one 128-row tile (C2: 2 x 49 slots,
//     C5: 1 x 125 slots); row = (window, slot).  A CTA works on NT tiles at a time, i.e. NT*G
//     consecutive (x-adjacent) windows.
//   * gather (the fused `window`): the CTA builds a small table {entry -> (global element offset,
//     tile row)} whose warp-lanes run along x ACROSS the windows of the CTA, so one warp load
//     touches as few 128-byte lines as the geometry allows (a window alone uses 10-14 bytes of a
//     line).  Zero padding is a real token with q = k = v = 0, exactly as NNlib.unfold produces
//     (SURVEY A.3).  Elements go to shared memory in the canonical SWIZZLE_128B [channel][token]
//     layout (the layout a TMA box has in the dense kernels), so S = Q K^T is one M=128, N=128
//     UMMA chain with MN-major operands and O = P V takes P from TMEM and V as a K-major operand.
//   * softmax is row-local (thread == row) with a block-diagonal mask (columns of the row's window).
//   * epilogue (the fused `unwindow`/fold): O rows are staged in fp32 over the dead Q/K tiles and
//     scattered with the same x-contiguous lane mapping: 16-bit stores when windows do not overlap
//     (stride >= W), fp32 atomicAdd into the fold accumulator when they do (fold_finalize then
//     divides by the count, src/windowed.jl:16-19).
// HBM-bound by design: q/k/v are read once per covering window, y written once.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "fa_common.cuh"
#include "fa_ptx.cuh"

namespace fa {
namespace {

using namespace ptx;

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int MAXE = 640;      // table entries per CTA iteration (20 warp-items); sized so that 2 CTAs fit one SM
constexpr int CBMAX = 64;      // channels per gather / scatter unit (64 loads in flight per lane); d = 32: one unit of 32

template <int FMT> struct El { using type = __half; };
template <> struct El<1> { using type = __nv_bfloat16; };

template <int FMT>
__device__ __forceinline__ uint32_t pack16(float a, float b) {
  if (FMT == 1) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 16-bit loads/stores through 32-bit registers (zero-extended / truncated): no PRMT repacking
__device__ __forceinline__ uint32_t ldg_nc_u16(const void* p) {
  uint32_t v;
  asm volatile("ld.global.nc.u16 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// one gather unit: NCH channels of one token (element offset `so`, -1 = zero padding) into the
// SWIZZLE_128B [channel][token] tile; `dst` = smem address of (row, first channel) before swizzling
template <int NCH>
__device__ __forceinline__ uint32_t gather_unit(const unsigned short* base, long long so, long long N, uint32_t dst, uint32_t chunk) {
  uint32_t sw[8];
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) sw[kk] = dst + ((chunk ^ (uint32_t)kk) << 4);
  uint32_t v[NCH];
  if (so >= 0) {
    const char* p = reinterpret_cast<const char*>(base + so);
    const long long sb = N * 2;
#pragma unroll
    for (int j = 0; j < NCH; ++j) v[j] = ldg_nc_u16(p + j * sb);
  } else {
#pragma unroll
    for (int j = 0; j < NCH; ++j) v[j] = 0;
  }
  uint32_t mx = 0;
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    sts_u16(sw[j & 7] + (uint32_t)(j * 128), v[j]);
    mx = max(mx, v[j] & 0x7fffu);
  }
  return mx;     // max |x| bit pattern (used by the bf16 re-encoding of the backward)
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }

// lane packing of the gather / scatter table (computed on the host)
struct WinMap {
  int G;          // windows per 128-row tile
  int nwc;        // windows per CTA iteration = NT * G
  int TW;         // tokens of one run group = nwc * W (contiguous in x when stride == W)
  int RG;         // run groups = W^D / W
  int gpw;        // TW <= 32: run groups per warp-item
  int wpg;        // TW  > 32: warp-items per run group
  int nwi;        // warp-items per CTA iteration (<= 32)
};

struct WinParams {
  const void *q, *k, *v;
  void* y;          // non-overlapping windows: output
  float* acc;       // overlapping windows: fp32 fold accumulator (zero-initialised)
  float *l, *m;     // (W^D, 1, L, B)
  Geo g;
  WinMap mp;
  long long ngroups, nwin;
  float scale_log2;
  long long* trace;  // FA_TRACE builds: CTA 1 records clock64() per phase of its first 16 iterations
};

#ifdef FA_TRACE
#define WTRACE(ev) do { if (prm.trace && blockIdx.x == 1 && tid == 0 && it < 16) prm.trace[it * 8 + (ev)] = clock64(); } while (0)
#else
#define WTRACE(ev) do {} while (0)
#endif

// TMAG: the gather is done by TMA instead of per-thread loads (exact-cover windows, stride == W): one 5-D box
// (x: all windows of the CTA, y: W, z: W, all channels, 1 batch) per tensor lands in a staging buffer --
// full sectors, zero fill outside the volume = the zero padding of `window`, no registers, no load
// instructions, and the box of the next tensor / next group is in flight while the CTA works -- and is then
// repacked shared -> shared into the SWIZZLE_128B [channel][token] operand tiles.
template <int D, int NT, int TMAG = 0> struct WCfg {
  static constexpr int THREADS = 128 * NT;
  static constexpr int BOX_BYTES = 64 * D * 2;
  static constexpr int TILE_BYTES = 2 * BOX_BYTES;
  static constexpr int OFF_TILES = 0;                        // [NT][q,k,v]
  static constexpr int STG_BYTES = TMAG ? 40 * 1024 : 0;     // one staged box (all or a slice of the channels of one tensor)
  static constexpr int OFF_STG = NT * 3 * TILE_BYTES;        // [2] staging buffers
  static constexpr int RPK_MAX = 1024;                       // repack table entries (run groups x tokens per run group)
  static constexpr int OFF_RPK = OFF_STG + 2 * STG_BYTES;    // uint2[RPK_MAX]
  static constexpr int OFF_SRC = OFF_RPK + (TMAG ? RPK_MAX * 8 : 0);   // long long[MAXE]
  static constexpr int OFF_ROW = OFF_SRC + MAXE * 8;         // int[MAXE]
  static constexpr int OFF_WIN = OFF_ROW + MAXE * 4;         // int4[256] window origins
  static constexpr int OFF_BAR = OFF_WIN + 256 * 16;
  static constexpr int SMEM_BYTES = OFF_BAR + 64 + 1024;
  static constexpr int COLS_PER_TILE = (D <= 64) ? 128 : 256;
  static constexpr int TMEM_COLS = COLS_PER_TILE * NT;
  static constexpr int COL_O = (D <= 64) ? 64 : 128;         // P (64 cols) aliases S; O beside / after it
  // 228 KB of shared memory per SM, 1 KB of it reserved per resident CTA
  static constexpr int BY_SMEM = 228 * 1024 / (SMEM_BYTES + 1024);
  static constexpr int CTAS_PER_SM = 512 / TMEM_COLS < BY_SMEM ? 512 / TMEM_COLS : BY_SMEM;
  static_assert(TMAG || NT != 2 || D != 64 || CTAS_PER_SM == 2, "d = 64 pair kernel must fit twice per SM");
  static_assert(D == 32 || D == 64 || D == 128, "head dims of the tcgen05 windowed kernels");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct TmaGeo {       // TMAG only
  int BX, by, bz;     // box extents (tokens) in x, y, z
  int RGN, TW;        // run groups per window (W^(D-1)), tokens per run group of the CTA (nwc * W)
  int gpr;            // groups per window row = ceil(Lx / nwc)
  int CH, ncs;        // channels per box, boxes per tensor (CH * ncs = d)
  unsigned box_bytes;
};

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---- table: entry e = warp-item * 32 + lane  ->  (source element offset | -1, (tile << 8) | row | -1)
// wininfo[win] = {x0, y0, z0, batch | -1}: origin (first token per dim, may be negative = padding) of
// every window of this CTA iteration, computed once per window; entries then need 32-bit math only.
__device__ __forceinline__ void build_wininfo(const Geo& g, const WinMap& mp, long long gw0, long long nwin, int4* wininfo, int tid,
                                              int nvalid = 1 << 30) {
  if (tid < mp.nwc) {
    const long long gw = gw0 + tid;
    int4 wi4 = make_int4(0, 0, 0, -1);
    if (gw < nwin && tid < nvalid) {
      const long long b = gw / g.L;
      int w = (int)(gw - b * g.L);
      int o3[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int wk = w % g.o[k];
        w /= g.o[k];
        o3[k] = (k < g.nd) ? wk * g.stride - g.padv[k] : 0;
      }
      wi4 = make_int4(o3[0], o3[1], o3[2], (int)b);
    }
    wininfo[tid] = wi4;
  }
}

__device__ __forceinline__ void build_table(const Geo& g, const WinMap& mp, const int4* wininfo, int D,
                                            long long* src, int* rowinfo, int tid, int nthreads, int* rowtok = nullptr) {
  const int W = g.W, WD = g.WD;
  for (int e = tid; e < mp.nwi * 32; e += nthreads) {
    const int witem = e >> 5, lane = e & 31;
    int rg, t;
    bool has;
    if (mp.TW <= 32) { const int gs = lane / mp.TW; t = lane - gs * mp.TW; rg = witem * mp.gpw + gs; has = gs < mp.gpw && rg < mp.RG; }
    else { rg = witem / mp.wpg; t = (witem - rg * mp.wpg) * 32 + lane; has = t < mp.TW; }
    long long so = -1;
    int ri = -1;
    if (has) {
      const int win = t / W, kx = t - win * W;
      const int kz = rg / W, ky = rg - kz * W;             // rg = ky + W kz (0 beyond the spatial rank)
      const int slot = rg * W + kx;
      const int ti = win / mp.G;
      ri = (ti << 8) | ((win - ti * mp.G) * WD + slot);
      const int4 wi4 = wininfo[win];
      int tok = -1;
      if (wi4.w >= 0) {
        const int x = wi4.x + kx, y = wi4.y + ky, z = wi4.z + kz;
        if (x >= 0 && x < g.s[0] && y >= 0 && y < g.s[1] && z >= 0 && z < g.s[2]) {
          tok = (z * g.s[1] + y) * g.s[0] + x;
          so = (long long)wi4.w * D * g.N + tok;
        }
      }
      if (rowtok) rowtok[(ri >> 8) * 128 + (ri & 255)] = tok;
    }
    src[e] = so;
    rowinfo[e] = ri;
  }
}

template <int D, int NT, int FMT, int TMAG>
__global__ void __launch_bounds__(WCfg<D, NT, TMAG>::THREADS, WCfg<D, NT, TMAG>::CTAS_PER_SM)
tc_win_fwd_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                  const __grid_constant__ CUtensorMap tmv, const WinParams prm, const TmaGeo tg) {
  using C = WCfg<D, NT, TMAG>;
  using T = typename El<FMT>::type;
  constexpr int CB = D < CBMAX ? D : CBMAX;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sptr = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sptr);
  long long* tsrc = reinterpret_cast<long long*>(sptr + C::OFF_SRC);
  int* trow = reinterpret_cast<int*>(sptr + C::OFF_ROW);
  int4* wininfo = reinterpret_cast<int4*>(sptr + C::OFF_WIN);
  const uint32_t bar_s = sbase + C::OFF_BAR, bar_o = bar_s + 8, tmem_slot = bar_s + 16, bar_stg = bar_s + 24;   // bar_stg[2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NWARPS = C::THREADS / 32;
  const Geo& g = prm.g;
  const WinMap mp = prm.mp;
  const long long N = g.N;
  const int WD = g.WD;
  uint2* rpk = reinterpret_cast<uint2*>(sptr + C::OFF_RPK);

  if (tid == 0) {
    mbar_init(bar_s, 1); mbar_init(bar_o, 1); mbar_init(bar_stg, 1); mbar_init(bar_stg + 8, 1); fence_barrier_init();
    if (TMAG) { prefetch_tensormap(&tmq); prefetch_tensormap(&tmk); prefetch_tensormap(&tmv); }
  }
  if (TMAG) {
    // static repack table: entry (run group rg, token t of the CTA's x-range) -> operand-tile position
    for (int e = tid; e < tg.RGN * tg.TW; e += C::THREADS) {
      const int rg = e / tg.TW, t = e - rg * tg.TW;
      const int w = t / g.W, kx = t - w * g.W;
      const int til = w / mp.G, rr = (w - til * mp.G) * WD + rg * g.W + kx;
      const uint32_t dst_lo = (uint32_t)(til * 3 * C::TILE_BYTES + (rr >> 6) * C::BOX_BYTES + (rr & 7) * 2);
      rpk[e] = make_uint2(dst_lo | ((uint32_t)((rr & 63) >> 3) << 24), (uint32_t)(rg * tg.BX + t) | ((uint32_t)w << 16));
    }
    for (int e = tg.RGN * tg.TW + tid; e < (tg.RGN * tg.TW + 255) / 256 * 256; e += C::THREADS) rpk[e] = make_uint2(0u, 0xffff0000u);
  }
  if (warp == 0) tmem_alloc(tmem_slot, C::TMEM_COLS);
  // rows no window maps to (>= G * W^D) stay zero for the whole kernel: K/V pad rows must be finite
  for (int i = tid; i < NT * 3 * C::TILE_BYTES / 16; i += C::THREADS)
    reinterpret_cast<uint4*>(sptr)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // softmax role: thread == row `row` of tile `ti`
  const int ti = tid >> 7, row = tid & 127;
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tS = tmem_base + lane_addr + ti * C::COLS_PER_TILE, tO = tS + C::COL_O;
  const int wi = row / WD, slot = row - wi * WD;
  const int c_lo = wi * WD, c_hi = c_lo + WD;           // columns of this row's own window
  const unsigned short* tens[3] = {static_cast<const unsigned short*>(prm.q), static_cast<const unsigned short*>(prm.k),
                                   static_cast<const unsigned short*>(prm.v)};
  constexpr uint32_t idesc_qk = make_idesc_f16(FMT, FMT, 1, 1, 128, 128);
  constexpr uint32_t idesc_pv = make_idesc_f16(FMT, FMT, 0, 0, 128, D);
  const float2 scale2 = make_float2(prm.scale_log2, prm.scale_log2);

  // TMAG: a group is nwc x-adjacent windows of ONE window row; decode (first window, valid count, box origin)
  auto decode = [&](long long grp, long long& gw0, int& nvalid, int& x0, int& y0, int& z0, int& b) {
    const long long rowi = grp / tg.gpr;
    const int gi = (int)(grp - rowi * tg.gpr);
    const int wx0 = gi * mp.nwc;
    nvalid = g.o[0] - wx0 < mp.nwc ? g.o[0] - wx0 : mp.nwc;
    const int wy = (int)(rowi % g.o[1]);
    const long long r2 = rowi / g.o[1];
    const int wz = (int)(r2 % g.o[2]);
    b = (int)(r2 / g.o[2]);
    gw0 = (((long long)b * g.o[2] + wz) * g.o[1] + wy) * g.o[0] + wx0;
    x0 = wx0 * g.stride - g.padv[0];                     // the box itself starts at the 16-byte boundary below x0
    y0 = g.nd >= 2 ? wy * g.stride - g.padv[1] : 0;
    z0 = g.nd >= 3 ? wz * g.stride - g.padv[2] : 0;
  };
  // load number `pos` of this CTA (lpg = 3 * ncs per group: q, k, v in channel slices) goes to staging buffer pos & 1.
  // TMA wants the innermost coordinate 16-byte aligned: the box starts at the multiple of 8 below x0.
  const int lpg = 3 * tg.ncs;
  auto issue_load = [&](long long grp, int j, uint32_t pos) {
    long long gw0; int nvalid, x0, y0, z0, b;
    decode(grp, gw0, nvalid, x0, y0, z0, b);
    const int x = j / tg.ncs, cs = j - x * tg.ncs;
    const uint32_t bar = bar_stg + 8u * (pos & 1u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic reads of the buffer precede this async write
    mbar_arrive_expect_tx(bar, tg.box_bytes);
    tma_load_5d(sbase + C::OFF_STG + (pos & 1u) * C::STG_BYTES, x == 0 ? &tmq : x == 1 ? &tmk : &tmv, bar,
                (x0 + 1024) / 8 * 8 - 1024, y0, z0, cs * tg.CH, b);
  };
  if (TMAG && tid == 0 && (long long)blockIdx.x < prm.ngroups) { issue_load(blockIdx.x, 0, 0); issue_load(blockIdx.x, 1, 1); }

  uint32_t it = 0;
  for (long long grp = blockIdx.x; grp < prm.ngroups; grp += gridDim.x, ++it) {
    long long gw0 = grp * mp.nwc;
    int nvalid = mp.nwc, xshift = 0;
    if (TMAG) { int x0, y0, z0, b; decode(grp, gw0, nvalid, x0, y0, z0, b); xshift = (x0 + 1024) & 7; }
    WTRACE(0);                                           // iteration start
    build_wininfo(g, mp, gw0, prm.nwin, wininfo, tid, nvalid);
    __syncthreads();
    build_table(g, mp, wininfo, D, tsrc, trow, tid, C::THREADS);
    __syncthreads();
    WTRACE(1);                                           // tables built

    if (TMAG) {
      // ---- gather by TMA + shared -> shared repack (8 table entries per lane held in registers per pass)
      const int nent = tg.RGN * tg.TW;
      const uint32_t cstride = (uint32_t)(tg.RGN * tg.BX * 2);
#pragma unroll 1
      for (int j = 0; j < lpg; ++j) {
        const uint32_t pos = (uint32_t)lpg * it + (uint32_t)j;
        const int x = j / tg.ncs, cs = j - x * tg.ncs;
        mbar_wait(bar_stg + 8u * (pos & 1u), (pos >> 1) & 1u);
        const uint32_t stg = sbase + C::OFF_STG + (pos & 1u) * C::STG_BYTES + (uint32_t)xshift * 2u;
        for (int e0 = 0; e0 < nent; e0 += 256) {
          uint32_t so[8], dlo[8], dch[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint2 en = rpk[e0 + k * 32 + lane];
            const bool on = (int)(en.y >> 16) < nvalid;
            so[k] = on ? (en.y & 0xffffu) * 2u : 0xffffffffu;
            dlo[k] = en.x & 0xffffffu;
            dch[k] = en.x >> 24;
          }
          for (int cl = warp; cl < tg.CH; cl += NWARPS) {
            const int c = cs * tg.CH + cl;
            const uint32_t srow = stg + (uint32_t)cl * cstride;
            const uint32_t drow = sbase + (uint32_t)(x * C::TILE_BYTES + c * 128);
            uint32_t val[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (so[k] != 0xffffffffu) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(val[k]) : "r"(srow + so[k]));
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (so[k] != 0xffffffffu) sts_u16(drow + dlo[k] + (((dch[k] ^ (uint32_t)c) & 7u) << 4), val[k]);
          }
        }
        __syncthreads();                                 // staging buffer fully consumed
        if (tid == 0) {                                  // refill it with the load two positions ahead
          const int nj = j + 2 < lpg ? j + 2 : j + 2 - lpg;
          const long long ngrp = j + 2 < lpg ? grp : grp + gridDim.x;
          if (ngrp < prm.ngroups) issue_load(ngrp, nj, pos + 2);
        }
      }
    } else {
    // ---- gather (fused `window`): unit = (warp-item, tensor, 32-channel block)
    {
      const int units = mp.nwi * 3 * (D / CB);
      for (int u = warp; u < units; u += NWARPS) {
        const int witem = u % mp.nwi, rest = u / mp.nwi;
        const int x = rest % 3, c0 = (rest / 3) * CB;
        const int e = witem * 32 + lane;
        const int ri = trow[e];
        if (ri < 0) continue;
        const long long so = tsrc[e];
        const int t_i = ri >> 8, rr = ri & 255;
        const uint32_t dst = sbase + (uint32_t)((t_i * 3 + x) * C::TILE_BYTES + (rr >> 6) * C::BOX_BYTES + (rr & 7) * 2 + c0 * 128);
        gather_unit<CB>(tens[x] + (long long)c0 * N, so, N, dst, (uint32_t)((rr & 63) >> 3));
      }
    }
    }
    fence_proxy_async();
    __syncthreads();
    WTRACE(2);                                           // gather done

    // ---- S = Q K^T per tile (M = 128 rows, N = 128 columns, K = D channels)
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const uint64_t qdesc = make_smem_desc_sw128(sbase + (t * 3 + 0) * C::TILE_BYTES, C::BOX_BYTES, 1024);
          const uint64_t kdesc = make_smem_desc_sw128(sbase + (t * 3 + 1) * C::TILE_BYTES, C::BOX_BYTES, 1024);
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks)
            mma_ss(tmem_base + t * C::COLS_PER_TILE, qdesc + (uint64_t)(ks * 128), kdesc + (uint64_t)(ks * 128), idesc_qk, ks > 0 ? 1u : 0u);
        }
        tc_commit(bar_s);
      }
      __syncwarp();
    }
    const long long gw = gw0 + (long long)ti * mp.G + wi;
    const bool valid = wi < mp.G && gw < prm.nwin && ti * mp.G + wi < nvalid;
    mbar_wait(bar_s, it & 1u);
    WTRACE(3);                                           // S ready
    tc_fence_after();

    // ---- softmax over the columns of this row's window (block-diagonal mask).  Each 32-column chunk
    // is classified per warp: entirely inside every lane's window (no predicates), entirely outside
    // (P = 0, S not even read), or mixed (per-element select).  Rows no window maps to take the
    // "inside" path with m = +inf, i.e. P = exp2(-inf) = 0.
    const int r_lo = valid ? c_lo : 0, r_hi = valid ? c_hi : 128;
    uint32_t cls = 0;                                    // 2 bits per chunk: 1 = inside, 2 = outside, 0 = mixed
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const bool in = r_lo <= 32 * ch && 32 * ch + 32 <= r_hi, out = r_hi <= 32 * ch || r_lo >= 32 * ch + 32;
      cls |= (__all_sync(0xffffffffu, in) ? 1u : (__all_sync(0xffffffffu, out) ? 2u : 0u)) << (2 * ch);
    }
    float mx = -INFINITY;
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      const uint32_t k = (cls >> (2 * ch)) & 3u;
      if (k == 2u) continue;
      uint32_t s[32];
      tmem_ld32(tS + 32 * ch, s);
      tmem_wait_ld();
      if (k == 1u) {
        float m0 = mx, m1 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          m0 = fmaxf(m0, fmaxf(__uint_as_float(s[e]), __uint_as_float(s[e + 1])));
          m1 = fmaxf(m1, fmaxf(__uint_as_float(s[e + 2]), __uint_as_float(s[e + 3])));
        }
        mx = fmaxf(m0, m1);
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int col = 32 * ch + e;
          if (col >= r_lo && col < r_hi) mx = fmaxf(mx, __uint_as_float(s[e]));
        }
      }
    }
    const float m2 = valid ? mx * prm.scale_log2 : INFINITY;
    const float2 negm2 = make_float2(-m2, -m2);
    float2 ls2 = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      const uint32_t k = (cls >> (2 * ch)) & 3u;
      uint32_t pk[16];
      if (k == 2u) {
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = 0u;
      } else {
        uint32_t s[32];
        tmem_ld32(tS + 32 * ch, s);
        tmem_wait_ld();
        if (k == 1u) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), scale2, negm2);
            const float2 p = make_float2(ex2(x.x), ex2(x.y));
            ls2 = __fadd2_rn(ls2, p);
            pk[e >> 1] = pack16<FMT>(p.x, p.y);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const int col = 32 * ch + e;
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), scale2, negm2);
            const float p0 = (col >= r_lo && col < r_hi) ? ex2(x.x) : 0.f;
            const float p1 = (col + 1 >= r_lo && col + 1 < r_hi) ? ex2(x.y) : 0.f;
            ls2 = __fadd2_rn(ls2, make_float2(p0, p1));
            pk[e >> 1] = pack16<FMT>(p0, p1);
          }
        }
      }
      tmem_st16(tS + 16 * ch, pk);     // P (16-bit) over the S columns already consumed
    }
    const float lsum = ls2.x + ls2.y;
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    WTRACE(4);                                           // softmax done (all warps)

    // ---- O = P V per tile (A = P from TMEM, B = V K-major, K = 128 keys)
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const uint64_t vdesc = make_smem_desc_sw128(sbase + (t * 3 + 2) * C::TILE_BYTES, 16, 1024);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            mma_ts(tmem_base + t * C::COLS_PER_TILE + C::COL_O, tmem_base + t * C::COLS_PER_TILE + ks * 8,
                   vdesc + (uint64_t)((ks >> 2) * (C::BOX_BYTES >> 4) + (ks & 3) * 2), idesc_pv, ks > 0 ? 1u : 0u);
        }
        tc_commit(bar_o);
      }
      __syncwarp();
    }
    if (valid) {      // l, m for every slot, padded ones included
      prm.l[gw * WD + slot] = lsum;
      prm.m[gw * WD + slot] = m2 * LN2;
    }
    const float inv_l = valid ? 1.f / lsum : 0.f;      // rows no window maps to: O row is never scattered
    mbar_wait(bar_o, it & 1u);
    WTRACE(5);                                           // O ready
    tc_fence_after();

    // ---- O rows -> fp32 staging [channel][128 rows] over the dead Q and K tiles of this tile
    const uint32_t stage = sbase + (uint32_t)(ti * 3 * C::TILE_BYTES) + (uint32_t)row * 4u;
#pragma unroll 1
    for (int ch = 0; ch < D / 32; ++ch) {
      uint32_t o[32];
      tmem_ld32(tO + 32 * ch, o);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; ++e) sts_f32(stage + (uint32_t)((32 * ch + e) * 512), __uint_as_float(o[e]) * inv_l);
    }
    tc_fence_before();
    __syncthreads();
    WTRACE(6);                                           // O staged

    // ---- scatter (fused `unwindow` / fold) with the gather's lane mapping
    {
      const int units = mp.nwi * (D / CB);
      for (int u = warp; u < units; u += NWARPS) {
        const int witem = u % mp.nwi, c0 = (u / mp.nwi) * CB;
        const int e = witem * 32 + lane;
        const int ri = trow[e];
        const long long so = tsrc[e];
        if (ri < 0 || so < 0) continue;
        const int t_i = ri >> 8, rr = ri & 255;
        const uint32_t sa = sbase + (uint32_t)(t_i * 3 * C::TILE_BYTES) + (uint32_t)(rr * 4 + c0 * 512);
        float v[CB];
#pragma unroll
        for (int j = 0; j < CB; ++j) v[j] = lds_f32(sa + (uint32_t)(j * 512));
        if (prm.acc) {
          float* a = prm.acc + so + (long long)c0 * N;
#pragma unroll
          for (int j = 0; j < CB; ++j) atomicAdd(a + (long long)j * N, v[j]);
        } else {
          T* y = static_cast<T*>(prm.y) + so + (long long)c0 * N;
#pragma unroll
          for (int j = 0; j < CB; ++j) y[(long long)j * N] = from_f32<T>(v[j]);
        }
      }
    }
    __syncthreads();      // staging (= Q/K tiles), table and TMEM are free for the next iteration
    WTRACE(7);                                           // scatter issued
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, C::TMEM_COLS);
}


// =====================================================================================================
// Backward (the reference has none): dYw = window(dY ./ count), per-window flash backward
// with P recomputed from the saved (l, m), dq = unwindow(dQw) etc.  One 128-row tile per CTA iteration:
//   A  gather q, k, v, dY tiles (bf16 inputs: per-tile power-of-two scales + in-place fp16 re-encoding,
//      so P and dS carry 11 significand bits in the MMAs -- see fa_tc_bwd.cu)
//   B  S = Q K^T, dP = dY V^T                        (thread == query row)
//      D_i = sum_j P_ij dP_ij,  dS = P o (dP - D) / count_i  -> TMEM;  dQ = tau dS K
//   C  S^T = K Q^T, dP^T = V dY^T                    (thread == key row; per-column stats from smem)
//      P^T / count, dS^T -> TMEM;  dV = P^T dY,  dK = tau dS^T Q
//   D  dQ, dV, dK rows staged in fp32 and scattered (16-bit stores, or fp32 atomics when windows overlap)
// TMEM per tile: T1 = S / S^T at 0, T2 = dP / dP^T at 128, dQ over T1, P^T over T1, dS(^T) over T2,
// dV and dK accumulators in the dead halves (d = 64) or in columns 256..511 (d = 128).
template <int D> struct WBCfg {
  static constexpr int THREADS = 128;
  static constexpr int BOX_BYTES = 64 * D * 2;
  static constexpr int TILE_BYTES = 2 * BOX_BYTES;
  static constexpr int OFF_TILES = 0;                        // q, k, v, dy
  static constexpr int BMAXE = 512;
  static constexpr int OFF_SRC = 4 * TILE_BYTES;             // long long[BMAXE]
  static constexpr int OFF_ROW = OFF_SRC + BMAXE * 8;        // int[BMAXE]
  static constexpr int OFF_TOK = OFF_ROW + BMAXE * 4;        // int[128] token of every tile row
  static constexpr int OFF_STAT = OFF_TOK + 128 * 4;         // float nlse[128], a[128], b[128]
  static constexpr int OFF_AMAX = OFF_STAT + 3 * 128 * 4;    // uint[4]
  static constexpr int OFF_WIN = OFF_AMAX + 16;              // int4[128] window origins
  static constexpr int OFF_BAR = OFF_WIN + 128 * 16;
  static constexpr int SMEM_BYTES = OFF_BAR + 64 + 1024;
  static constexpr int TMEM_COLS = (D <= 64) ? 256 : 512;
  static constexpr int COL_T1 = 0, COL_T2 = 128, COL_DQ = 0;
  static constexpr int COL_DV = (D <= 64) ? 64 : 256, COL_DK = (D <= 64) ? 192 : 384;
  static constexpr int CTAS_PER_SM = (D <= 64) ? 2 : 1;
};

struct WinBwdParams {
  const void *q, *k, *v, *dy;
  const float *l, *m;
  void *dq, *dk, *dv;        // non-overlapping windows: outputs
  float *aq, *ak, *av;       // overlapping windows: fp32 fold accumulators (zero-initialised)
  Geo g;
  WinMap mp;
  long long ngroups, nwin;
  float scale_log2, tau;
  long long* trace;          // FA_TRACE builds: CTA 1 records clock64() per phase of its first 16 iterations
};

#ifdef FA_TRACE
#define BTR(ev) do { if (prm.trace && blockIdx.x == 1 && tid == 0 && bit < 16) prm.trace[bit * 8 + (ev)] = clock64(); } while (0)
#else
#define BTR(ev) do {} while (0)
#endif

__device__ __forceinline__ float pow2_norm_scale(uint32_t amax_bits) {   // s = 2^k with amax * s in [4, 8)
  const int e = (int)((amax_bits >> 7) & 0xff);      // bf16 biased exponent (bits are |x| of a bf16 value)
  if (e == 0 || e == 255) return 1.f;
  int k = 129 - e;                                   // amax in [2^(e-127), 2^(e-126))  ->  [4, 8)
  k = k > 100 ? 100 : (k < -100 ? -100 : k);
  return __uint_as_float((uint32_t)(k + 127) << 23);
}

template <int D, int INBF>
__global__ void __launch_bounds__(128, WBCfg<D>::CTAS_PER_SM)
tc_win_bwd_kernel(const WinBwdParams prm) {
  using C = WBCfg<D>;
  using T = typename El<INBF>::type;
  constexpr int CB = D < CBMAX ? D : CBMAX;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sptr = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sptr);
  long long* tsrc = reinterpret_cast<long long*>(sptr + C::OFF_SRC);
  int* trow = reinterpret_cast<int*>(sptr + C::OFF_ROW);
  int* rowtok = reinterpret_cast<int*>(sptr + C::OFF_TOK);
  float* snl = reinterpret_cast<float*>(sptr + C::OFF_STAT);
  float* sa = snl + 128;
  float* sb = snl + 256;
  unsigned int* samax = reinterpret_cast<unsigned int*>(sptr + C::OFF_AMAX);
  int4* wininfo = reinterpret_cast<int4*>(sptr + C::OFF_WIN);
  const uint32_t bar_a = sbase + C::OFF_BAR, bar_b = bar_a + 8, tmem_slot = bar_a + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Geo& g = prm.g;
  const WinMap mp = prm.mp;
  const long long N = g.N;
  const int WD = g.WD;

  if (tid == 0) { mbar_init(bar_a, 1); mbar_init(bar_b, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, C::TMEM_COLS);
  for (int i = tid; i < 4 * C::TILE_BYTES / 16; i += 128) reinterpret_cast<uint4*>(sptr)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 128; i += 128) rowtok[i] = -1;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int row = tid;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  const uint32_t tT1 = tmem_base + lane_addr + C::COL_T1, tT2 = tmem_base + lane_addr + C::COL_T2;
  const int wi = row / WD, slot = row - wi * WD;
  const int c_lo = wi * WD, c_hi = c_lo + WD;
  const unsigned short* tens[4] = {static_cast<const unsigned short*>(prm.q), static_cast<const unsigned short*>(prm.k),
                                   static_cast<const unsigned short*>(prm.v), static_cast<const unsigned short*>(prm.dy)};
  // MMAs always run in fp16: bf16 inputs are re-encoded tile by tile
  constexpr uint32_t idesc_t = make_idesc_f16(0, 0, 1, 1, 128, 128);      // A, B MN-major, N = 128
  constexpr uint32_t idesc_acc = make_idesc_f16(0, 0, 0, 0, 128, D);      // A in TMEM, B K-major, N = D
  const uint32_t sQ = sbase, sK = sbase + C::TILE_BYTES, sV = sbase + 2 * C::TILE_BYTES, sG = sbase + 3 * C::TILE_BYTES;
  auto mn = [&](uint32_t a) { return make_smem_desc_sw128(a, C::BOX_BYTES, 1024); };
  auto km = [&](uint32_t a) { return make_smem_desc_sw128(a, 16, 1024); };
  auto issue_T = [&](uint32_t x1, uint32_t y1, uint32_t x2, uint32_t y2, uint32_t bar) {   // T1 = X1 Y1^T, T2 = X2 Y2^T
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks)
          mma_ss(tmem_base + C::COL_T1, mn(x1) + (uint64_t)(ks * 128), mn(y1) + (uint64_t)(ks * 128), idesc_t, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks)
          mma_ss(tmem_base + C::COL_T2, mn(x2) + (uint64_t)(ks * 128), mn(y2) + (uint64_t)(ks * 128), idesc_t, ks > 0 ? 1u : 0u);
        tc_commit(bar);
      }
      __syncwarp();
    }
  };
  auto issue_acc = [&](uint32_t col_acc, uint32_t col_a, uint32_t ytile) {   // acc = A(TMEM, 128 x 128) * Y (K-major)
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
      mma_ts(tmem_base + col_acc, tmem_base + col_a + ks * 8,
             km(ytile) + (uint64_t)((ks >> 2) * (C::BOX_BYTES >> 4) + (ks & 3) * 2), idesc_acc, ks > 0 ? 1u : 0u);
  };

  uint32_t pa = 0, pb = 0;
  int bit = -1;                                          // iteration counter (trace builds)
  for (long long grp = blockIdx.x; grp < prm.ngroups; grp += gridDim.x) {
    ++bit;
    BTR(6);                                              // iteration start
    const long long gw0 = grp * mp.nwc;
    if (tid < 4) samax[tid] = 0;
    build_wininfo(g, mp, gw0, prm.nwin, wininfo, tid);
    __syncthreads();
    build_table(g, mp, wininfo, D, tsrc, trow, tid, 128, rowtok);
    __syncthreads();
    BTR(0);                                              // tables built (iteration start + tables)

    // ---- A: gather q, k, v, dy
    {
      const int units = mp.nwi * 4 * (D / CB);
      for (int u = warp; u < units; u += 4) {
        const int witem = u % mp.nwi, rest = u / mp.nwi;
        const int x = rest & 3, c0 = (rest >> 2) * CB;
        const int e = witem * 32 + lane;
        const int ri = trow[e];
        uint32_t mx = 0;
        if (ri >= 0) {
          const int rr = ri & 255;
          const uint32_t dst = sbase + (uint32_t)(x * C::TILE_BYTES + (rr >> 6) * C::BOX_BYTES + (rr & 7) * 2 + c0 * 128);
          mx = gather_unit<CB>(tens[x] + (long long)c0 * N, tsrc[e], N, dst, (uint32_t)((rr & 63) >> 3));
        }
        if (INBF) {
          mx = __reduce_max_sync(0xffffffffu, mx);
          if (lane == 0 && mx) atomicMax(&samax[x], mx);
        }
      }
    }
    __syncthreads();
    BTR(1);                                              // gather done
    float sq = 1.f, sk = 1.f, sv = 1.f, sg = 1.f;
    if (INBF) {
      // per-tile power-of-two scales, then bf16 -> fp16 in place (exact for everything within 2^-17 of the max)
      sq = pow2_norm_scale(samax[0]); sk = pow2_norm_scale(samax[1]); sv = pow2_norm_scale(samax[2]); sg = pow2_norm_scale(samax[3]);
      constexpr int VPT = C::TILE_BYTES / 16;      // 16-byte vectors per tensor tile
      for (int i = tid; i < 4 * VPT; i += 128) {
        const int x = i / VPT;
        const float sc = x == 0 ? sq : x == 1 ? sk : x == 2 ? sv : sg;
        uint4 u = reinterpret_cast<uint4*>(sptr)[i];
        uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __half2 h = __floats2half2_rn(__uint_as_float(w[j] << 16) * sc, __uint_as_float(w[j] & 0xffff0000u) * sc);
          w[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        reinterpret_cast<uint4*>(sptr)[i] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    fence_proxy_async();
    __syncthreads();
    BTR(2);                                              // re-encode done

    // ---- B: S = Q K^T, dP = dY V^T;  thread == query row
    issue_T(sQ, sK, sG, sV, bar_a);
    const long long gw = gw0 + wi;
    const bool valid = wi < mp.G && gw < prm.nwin;
    float nl = -INFINITY, rc = 1.f;
    if (valid) {
      nl = -(prm.m[gw * WD + slot] + logf(prm.l[gw * WD + slot])) * LOG2E;
      const int tok = rowtok[row];
      if (g.overlap && tok >= 0) rc = 1.f / (float)window_count_at(g, tok);     // dYw = window(dY ./ count)
    }
    const float sl2 = prm.scale_log2 / (sq * sk);
    const float2 scale2 = make_float2(sl2, sl2), nl2 = make_float2(nl, nl);
    mbar_wait(bar_a, pa & 1u); ++pa;
    tc_fence_after();
    float Dsum = 0.f;
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t s[32], dp[32];
      tmem_ld32(tT1 + 32 * ch, s);
      tmem_ld32(tT2 + 32 * ch, dp);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        const int col = 32 * ch + e;
        const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), scale2, nl2);
        const float p0 = (col >= c_lo && col < c_hi) ? ex2(x.x) : 0.f;
        const float p1 = (col + 1 >= c_lo && col + 1 < c_hi) ? ex2(x.y) : 0.f;
        Dsum = fmaf(p0, __uint_as_float(dp[e]), Dsum);
        Dsum = fmaf(p1, __uint_as_float(dp[e + 1]), Dsum);
      }
    }
    if (!valid) Dsum = 0.f;
    snl[row] = nl; sa[row] = rc; sb[row] = -Dsum * rc;
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t s[32], dp[32], pk[16];
      tmem_ld32(tT1 + 32 * ch, s);
      tmem_ld32(tT2 + 32 * ch, dp);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        const int col = 32 * ch + e;
        const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), scale2, nl2);
        const float p0 = (col >= c_lo && col < c_hi) ? ex2(x.x) : 0.f;
        const float p1 = (col + 1 >= c_lo && col + 1 < c_hi) ? ex2(x.y) : 0.f;
        pk[e >> 1] = pack16<0>(p0 * (__uint_as_float(dp[e]) - Dsum) * rc, p1 * (__uint_as_float(dp[e + 1]) - Dsum) * rc);
      }
      tmem_st16(tT2 + 16 * ch, pk);      // dS (fp16) over the dP columns already consumed
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    BTR(3);                                              // pass B element-wise done
    if (warp == 0) {                     // dQ = dS K  (accumulator over the dead S columns)
      tc_fence_after();
      if (elect_one()) { issue_acc(C::COL_DQ, C::COL_T2, sK); tc_commit(bar_b); }
      __syncwarp();
    }
    mbar_wait(bar_b, pb & 1u); ++pb;
    tc_fence_after();
    float dq[D];
    {
      const float mul = prm.tau / (sv * sg * sk);
#pragma unroll
      for (int ch = 0; ch < D / 32; ++ch) {
        uint32_t o[32];
        tmem_ld32(tmem_base + lane_addr + C::COL_DQ + 32 * ch, o);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; ++e) dq[32 * ch + e] = __uint_as_float(o[e]) * mul;
      }
    }
    tc_fence_before();
    __syncthreads();
    BTR(4);                                              // dQ MMA issued / pass B done

    // ---- C: S^T = K Q^T, dP^T = V dY^T;  thread == key row, per-column (query) stats from smem
    issue_T(sK, sQ, sV, sG, bar_a);
    mbar_wait(bar_a, pa & 1u); ++pa;
    tc_fence_after();
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t s[32], dp[32], pk[16], dk[16];
      tmem_ld32(tT1 + 32 * ch, s);
      tmem_ld32(tT2 + 32 * ch, dp);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        const int col = 32 * ch + e;
        const float2 nlc = *reinterpret_cast<const float2*>(snl + col);
        const float2 ac = *reinterpret_cast<const float2*>(sa + col);
        const float2 bc = *reinterpret_cast<const float2*>(sb + col);
        const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), scale2, nlc);
        const float p0 = (col >= c_lo && col < c_hi) ? ex2(x.x) : 0.f;
        const float p1 = (col + 1 >= c_lo && col + 1 < c_hi) ? ex2(x.y) : 0.f;
        pk[e >> 1] = pack16<0>(p0 * ac.x, p1 * ac.y);                                        // P^T / count_i
        dk[e >> 1] = pack16<0>(p0 * fmaf(__uint_as_float(dp[e]), ac.x, bc.x), p1 * fmaf(__uint_as_float(dp[e + 1]), ac.y, bc.y));
      }
      tmem_st16(tT1 + 16 * ch, pk);
      tmem_st16(tT2 + 16 * ch, dk);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    BTR(5);                                              // pass C element-wise done
    if (warp == 0) {                     // dV = P^T dY,  dK = dS^T Q
      tc_fence_after();
      if (elect_one()) { issue_acc(C::COL_DV, C::COL_T1, sG); issue_acc(C::COL_DK, C::COL_T2, sQ); tc_commit(bar_b); }
      __syncwarp();
    }
    mbar_wait(bar_b, pb & 1u); ++pb;
    tc_fence_after();

    // ---- D: stage one gradient at a time in fp32 [channel][128 rows] over the dead q/k tiles, scatter
#pragma unroll 1
    for (int which = 0; which < 3; ++which) {
      const uint32_t stage = sbase + (uint32_t)row * 4u;
      if (which == 0) {
#pragma unroll
        for (int c = 0; c < D; ++c) sts_f32(stage + (uint32_t)(c * 512), dq[c]);
      } else {
        const float mul = which == 1 ? 1.f / sg : prm.tau / (sv * sg * sq);
        const uint32_t tacc = tmem_base + lane_addr + (which == 1 ? C::COL_DV : C::COL_DK);
#pragma unroll 1
        for (int ch = 0; ch < D / 32; ++ch) {
          uint32_t o[32];
          tmem_ld32(tacc + 32 * ch, o);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) sts_f32(stage + (uint32_t)((32 * ch + e) * 512), __uint_as_float(o[e]) * mul);
        }
      }
      __syncthreads();
      float* accp = which == 0 ? prm.aq : which == 1 ? prm.av : prm.ak;
      T* outp = static_cast<T*>(which == 0 ? prm.dq : which == 1 ? prm.dv : prm.dk);
      const int units = mp.nwi * (D / CB);
      for (int u = warp; u < units; u += 4) {
        const int witem = u % mp.nwi, c0 = (u / mp.nwi) * CB;
        const int e = witem * 32 + lane;
        const int ri = trow[e];
        const long long so = tsrc[e];
        if (ri < 0 || so < 0) continue;
        const uint32_t sadr = sbase + (uint32_t)((ri & 255) * 4 + c0 * 512);
        float v[CB];
#pragma unroll
        for (int j = 0; j < CB; ++j) v[j] = lds_f32(sadr + (uint32_t)(j * 512));
        if (accp) {
          float* a = accp + so + (long long)c0 * N;
#pragma unroll
          for (int j = 0; j < CB; ++j) atomicAdd(a + (long long)j * N, v[j]);
        } else {
          T* y = outp + so + (long long)c0 * N;
#pragma unroll
          for (int j = 0; j < CB; ++j) y[(long long)j * N] = from_f32<T>(v[j]);
        }
      }
      __syncthreads();
    }
    // the staging left fp32 bytes in the q and k tiles: rows no window maps to must read as zero
    // (their dS / P^T entries are exact zeros, and 0 x garbage could be NaN in the accumulating MMAs)
    {
      const int r0 = mp.G * WD, npad = (128 - r0) * D * 2;
      for (int i = tid; i < npad; i += 128) {
        const int x = i / ((128 - r0) * D), rem = i - x * (128 - r0) * D;
        const int c = rem / (128 - r0), r = r0 + rem - c * (128 - r0);
        sts_u16(sbase + (uint32_t)(x * C::TILE_BYTES + (r >> 6) * C::BOX_BYTES + c * 128 + ((((r & 63) >> 3) ^ (c & 7)) << 4) + (r & 7) * 2), 0);
      }
    }
    tc_fence_before();
    __syncthreads();
    BTR(7);                                              // scatter of dQ, dV, dK done
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

bool make_map(const Geo& g, int NT, WinMap& mp) {
  mp.G = 128 / g.WD;
  mp.nwc = NT * mp.G;
  mp.TW = mp.nwc * g.W;
  mp.RG = g.WD / g.W;
  mp.gpw = 1; mp.wpg = 1;
  if (mp.TW <= 32) { mp.gpw = 32 / mp.TW; mp.nwi = (mp.RG + mp.gpw - 1) / mp.gpw; }
  else { mp.wpg = (mp.TW + 31) / 32; mp.nwi = mp.RG * mp.wpg; }
  return mp.nwi * 32 <= MAXE;
}

typedef CUresult (*EncodeTiledFn5)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 5-D map over one (X, Y, Z, d, B) tensor (missing spatial dims = 1): box (BX, by, bz, CH, 1), no swizzle,
// zero fill out of bounds (= the zero padding of `window`, src/utils.jl:40)
}  // namespace
int make_win_tmap_box(CUtensorMap* tm, const void* base, int dtype, const Geo& g, int D, int BX, int by, int bz, int CH) {
  static EncodeTiledFn5 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn5>(p);
  });
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return FA_ERR_CUDA; }
  const cuuint64_t dims[5] = {(cuuint64_t)g.s[0], (cuuint64_t)g.s[1], (cuuint64_t)g.s[2], (cuuint64_t)D, (cuuint64_t)g.B};
  const cuuint64_t strides[4] = {(cuuint64_t)g.s[0] * 2, (cuuint64_t)g.s[0] * g.s[1] * 2, (cuuint64_t)g.N * 2, (cuuint64_t)g.N * D * 2};
  const cuuint32_t box[5] = {(cuuint32_t)BX, (cuuint32_t)by, (cuuint32_t)bz, (cuuint32_t)CH, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapDataType dt = dtype == FA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(tm, dt, 5, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (windowed 5-D) failed (CUresult %d)", (int)r); return FA_ERR_CUDA; }
  return FA_OK;
}
namespace {
int make_win_tmap(CUtensorMap* tm, const void* base, int dtype, const Geo& g, int D, const TmaGeo& tg) {
  return make_win_tmap_box(tm, base, dtype, g, D, tg.BX, tg.by, tg.bz, tg.CH);
}

// can the TMA-gather variant take this geometry?  (exact-cover windows, 16-byte aligned rows, box fits the staging)
template <int D, int NT>
bool tma_gather_geo(const Geo& g, const WinMap& mp, TmaGeo& tg) {
  using C = WCfg<D, NT, 1>;
  if (g.stride != g.W || g.s[0] % 8 != 0) return false;
  tg.TW = mp.nwc * g.W;
  // the box starts at the 16-byte boundary at or below the first token: up to 7 extra tokens in front
  const bool always_aligned = (mp.nwc * g.W) % 8 == 0 && g.padv[0] % 8 == 0;
  tg.BX = (tg.TW + (always_aligned ? 0 : 7) + 7) / 8 * 8;
  tg.by = g.nd >= 2 ? g.W : 1;
  tg.bz = g.nd >= 3 ? g.W : 1;
  tg.RGN = tg.by * tg.bz;
  tg.gpr = (g.o[0] + mp.nwc - 1) / mp.nwc;
  tg.CH = D;
  while (tg.CH > 8 && (long long)tg.BX * tg.by * tg.bz * tg.CH * 2 > C::STG_BYTES) tg.CH /= 2;
  tg.ncs = D / tg.CH;
  const long long bytes = (long long)tg.BX * tg.by * tg.bz * tg.CH * 2;
  if (tg.BX > 256 || g.padv[0] > 1024 || bytes > C::STG_BYTES || tg.RGN * tg.TW > C::RPK_MAX || tg.RGN * tg.BX + 8 > 65535) return false;
  tg.box_bytes = (unsigned)bytes;
  return true;
}

template <int D, int NT, int FMT>
int launch_win_fwd(const Geo& g, const FwdArgs& a, cudaStream_t st) {
  WinParams prm;
  prm.q = a.q; prm.k = a.k; prm.v = a.v; prm.y = a.o; prm.acc = a.acc; prm.l = a.l; prm.m = a.m;
  prm.g = g;
  if (!make_map(g, NT, prm.mp)) { set_error("tc_win_fwd: window table too large"); return FA_ERR_UNSUPPORTED; }
  prm.nwin = g.L * g.B;
  prm.ngroups = (prm.nwin + prm.mp.nwc - 1) / prm.mp.nwc;
  prm.scale_log2 = g.tau * LOG2E;
  prm.trace = nullptr;
#ifdef FA_TRACE
  { const char* e = getenv("FA_TRACE_PTR"); prm.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr; }
#endif
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  CUtensorMap tq, tk, tv;
  memset(&tq, 0, sizeof(tq)); memset(&tk, 0, sizeof(tk)); memset(&tv, 0, sizeof(tv));
  TmaGeo tg;
  memset(&tg, 0, sizeof(tg));
  static const int use_tma = [] { const char* e = getenv("FA_WIN_TMA"); return e ? atoi(e) : 0; }();
  const bool aligned = ((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v)) & 15) == 0;
  if (D == 64 && use_tma && aligned && tma_gather_geo<D, NT>(g, prm.mp, tg)) {
    using C = WCfg<D, NT, 1>;
    const int dtype = FMT ? FA_BF16 : FA_F16;
    int rc;
    if ((rc = make_win_tmap(&tq, a.q, dtype, g, D, tg))) return rc;
    if ((rc = make_win_tmap(&tk, a.k, dtype, g, D, tg))) return rc;
    if ((rc = make_win_tmap(&tv, a.v, dtype, g, D, tg))) return rc;
    prm.ngroups = (long long)tg.gpr * g.o[1] * g.o[2] * g.B;      // groups never straddle a window row
    auto kern = tc_win_fwd_kernel<D, NT, FMT, 1>;
    FA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    const long long cap = (long long)sms * C::CTAS_PER_SM;
    const unsigned grid = (unsigned)(prm.ngroups < cap ? prm.ngroups : cap);
    kern<<<grid, C::THREADS, C::SMEM_BYTES, st>>>(tq, tk, tv, prm, tg);
    FA_CUDA_TRY(cudaGetLastError());
    return FA_OK;
  }
  using C = WCfg<D, NT, 0>;
  auto kern = tc_win_fwd_kernel<D, NT, FMT, 0>;
  FA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  const long long cap = (long long)sms * C::CTAS_PER_SM;
  const unsigned grid = (unsigned)(prm.ngroups < cap ? prm.ngroups : cap);
  kern<<<grid, C::THREADS, C::SMEM_BYTES, st>>>(tq, tk, tv, prm, tg);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

template <int D, int INBF>
int launch_win_bwd(const Geo& g, const BwdArgs& a, cudaStream_t st) {
  using C = WBCfg<D>;
  WinBwdParams prm;
  prm.q = a.q; prm.k = a.k; prm.v = a.v; prm.dy = a.d_o; prm.l = a.l; prm.m = a.m;
  prm.dq = a.dq; prm.dk = a.dk; prm.dv = a.dv; prm.aq = a.aq; prm.ak = a.ak; prm.av = a.av;
  prm.g = g;
  if (!make_map(g, 1, prm.mp) || prm.mp.nwi * 32 > C::BMAXE) { set_error("tc_win_bwd: window table too large"); return FA_ERR_UNSUPPORTED; }
  prm.nwin = g.L * g.B;
  prm.ngroups = (prm.nwin + prm.mp.nwc - 1) / prm.mp.nwc;
  prm.scale_log2 = g.tau * LOG2E; prm.tau = g.tau;
  prm.trace = nullptr;
#ifdef FA_TRACE
  { const char* e = getenv("FA_TRACE_PTR"); prm.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr; }
#endif
  auto kern = tc_win_bwd_kernel<D, INBF>;
  FA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long cap = (long long)sms * C::CTAS_PER_SM;
  const unsigned grid = (unsigned)(prm.ngroups < cap ? prm.ngroups : cap);
  kern<<<grid, 128, C::SMEM_BYTES, st>>>(prm);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

}  // namespace

bool tc_win_supported(const Geo& g, int dtype);
bool tc_win_bwd_supported(const Geo& g, int dtype) {
  if (!tc_win_supported(g, dtype)) return false;
  WinMap mp;
  return make_map(g, 1, mp) && mp.nwi * 32 <= 512;
}

int tc_win_bwd(const Geo& g, const BwdArgs& a, int dtype, cudaStream_t st) {
  if (!tc_win_bwd_supported(g, dtype)) { set_error("tc_win_bwd: unsupported configuration"); return FA_ERR_UNSUPPORTED; }
  const int bf = dtype == FA_BF16 ? 1 : 0;
  if (g.d == 128) return bf ? launch_win_bwd<128, 1>(g, a, st) : launch_win_bwd<128, 0>(g, a, st);
  if (g.d == 32) return bf ? launch_win_bwd<32, 1>(g, a, st) : launch_win_bwd<32, 0>(g, a, st);
  return bf ? launch_win_bwd<64, 1>(g, a, st) : launch_win_bwd<64, 0>(g, a, st);
}

bool tc_win_supported(const Geo& g, int dtype) {
  if (dtype != FA_BF16 && dtype != FA_F16) return false;
  if (g.mode != MODE_WINDOWED) return false;
  if (g.d != g.dv || (g.d != 32 && g.d != 64 && g.d != 128)) return false;
  if (g.WD > 128 || g.WD < 1) return false;
  WinMap mp;
  return make_map(g, 2, mp);
}

bool tc_winx_supported(const Geo& g, const FwdArgs& a, int dtype);
int tc_winx_fwd(const Geo& g, const FwdArgs& a, int dtype, cudaStream_t st);
int tc_winx_groups_per_window_row(const Geo& g);

int tc_win_fwd(const Geo& g, const FwdArgs& a, int dtype, cudaStream_t st) {
  if (!tc_win_supported(g, dtype)) { set_error("tc_win_fwd: unsupported configuration"); return FA_ERR_UNSUPPORTED; }
  const int fmt = dtype == FA_BF16 ? 1 : 0;
  // large exact-cover 3-D problems: the streamed kernel of fa_tc_winx.cu (TMA boxes + 16-byte repack, four tiles per SM).
  // Measured on B200 (profiles/r2f_winx.md): config 5 at B = 64 6.45 vs 7.33 ms, one 256^3 volume 4.6 vs 9.8 ms; 2-D
  // images (short (y) columns: 7 rows per box) are faster on the per-thread gather below (0.75 vs 1.35 ms at B = 512).
  // FA_WINX = 0 never, 1 whenever it applies, unset: 3-D volumes with at least four groups per SM
  {
    static const int winx = [] { const char* e = getenv("FA_WINX"); return e ? atoi(e) : -1; }();
    if (winx != 0 && tc_winx_supported(g, a, dtype)) {
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const long long groups = (long long)tc_winx_groups_per_window_row(g) * g.o[1] * g.o[2] * g.B;
      if (winx == 1 || (g.nd == 3 && groups >= 8LL * sms)) return tc_winx_fwd(g, a, dtype, st);
    }
  }
  if (g.d == 128) return fmt ? launch_win_fwd<128, 1, 1>(g, a, st) : launch_win_fwd<128, 1, 0>(g, a, st);
  if (g.d == 32) return fmt ? launch_win_fwd<32, 2, 1>(g, a, st) : launch_win_fwd<32, 2, 0>(g, a, st);     // the reference's benchmark head dim (logs/wind_t*.txt)
  // d = 64: two 128-row tiles per CTA (2 CTAs / SM; longer x-runs per gather) or one tile per CTA (3 CTAs / SM).
  // Small problems -- fewer than two pair-groups per resident CTA -- are latency bound and take the finer split
  // (config 2: 19.0 -> 15.0 us); large ones keep the pair kernel (config 5: 0.93 vs 0.99 ms).  FA_WIN_NT overrides.
  static const int nt_env = [] { const char* e = getenv("FA_WIN_NT"); return e ? atoi(e) : 0; }();
  int nt = nt_env;
  if (nt != 1 && nt != 2) {
    WinMap mp2;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    make_map(g, 2, mp2);
    const long long groups2 = (g.L * g.B + mp2.nwc - 1) / mp2.nwc;
    nt = groups2 < 4LL * sms ? 1 : 2;
  }
  if (nt == 1) return fmt ? launch_win_fwd<64, 1, 1>(g, a, st) : launch_win_fwd<64, 1, 0>(g, a, st);
  return fmt ? launch_win_fwd<64, 2, 1>(g, a, st) : launch_win_fwd<64, 2, 0>(g, a, st);
}

}  // namespace fa
