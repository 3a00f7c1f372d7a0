// fa_tc_winx.cu -- streamed tcgen05 windowed attention forward for sm_100a (round 2): windowed_fa / block_fa
// (reference src/windowed.jl:1-23, window / unwindow of src/utils.jl:36-54 fused in) for exact-cover windows
// (stride == W), d = dv = 64, 16-bit inputs.  Replaces the per-thread 2-byte gather of fa_tc_win.cu, which the
// round-1 profile showed bound by the L1 data pipe (every (channel, y, z) line of the volume is touched for 10-14
// useful bytes), on the large problems; fa_tc_win.cu stays for everything this kernel does not take.
//
// One persistent 512-thread CTA per SM works on GROUPS of nwc = 4 * G x-adjacent windows (G = floor(128 / W^D)
// windows per 128-row tile, four tiles, all 512 TMEM columns):
//   * q, k, v arrive as 5-D TMA boxes (x: the group's tokens rounded out to 16-byte boundaries, y: W, z: W,
//     16 channels, 1 batch element) through a ring of three 25 KB staging slots -- full 32-byte sectors, no load
//     instructions, zero fill outside the volume = the zero padding of `window`.  The boxes of a group are a
//     stream Q0 K0 Q1 K1 Q2 K2 Q3 K3 V0 V1 V2 V3 (16-channel slices); the stream runs up to three boxes ahead of the
//     consumer, i.e. into the next group while this one is in its softmax / output phases.
//   * repack staging -> operand tiles: one thread produces one 16-byte chunk (8 consecutive tile rows of one
//     channel) of the canonical SWIZZLE_128B [channel][token] layout with 8 two-byte shared loads and ONE 16-byte
//     conflict-free shared store (round 1: one 2-byte global load + one 2-byte shared store per element).
//   * S = Q K^T is issued slice by slice (K = 16 channels per tcgen05.mma) as soon as the Q and K slices are
//     repacked, so the tensor work hides under the repack; V reuses the Q tiles once S is complete.
//   * softmax as in fa_tc_win.cu (thread == row, block-diagonal mask, 32-column chunks classified per warp).
//   * O rows are converted and staged in the volume's own order [channel][y, z row][x] over the dead K tiles and
//     written with 4-byte stores along x across all windows of the group (40-byte runs instead of 10-byte ones).
// HBM-bound by design: q, k, v read once, y written once; L2 -> SM traffic is the box overfetch (1.2-1.6x).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "fa_common.cuh"
#include "fa_ptx.cuh"

namespace fa {
int make_win_tmap_box(CUtensorMap* tm, const void* base, int dtype, const Geo& g, int D, int BX, int by, int bz, int CH);

namespace {
using namespace ptx;

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

template <int FMT> struct El { using type = __half; };
template <> struct El<1> { using type = __nv_bfloat16; };

template <int FMT>
__device__ __forceinline__ uint32_t pack16(float a, float b) {
  if (FMT == 1) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <int FMT>
__device__ __forceinline__ uint32_t cvt16(float a) { return pack16<FMT>(a, 0.f) & 0xffffu; }

__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__device__ __forceinline__ void tma_prefetch_5d(const void* tmap, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

// y[box] += staged box (element-wise add in L2, 16-bit type of the tensor map); coordinates outside the tensor are dropped
__device__ __forceinline__ void tma_reduce_add_5d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// NTILES = 4: one CTA per SM, four 128-row tiles = all 512 TMEM columns, 40-byte output runs (config 5: groups of four
// windows).  NTILES = 2: two CTAs per SM with two tiles each -- smaller boxes and 20-byte runs, but the two CTAs' phases
// (TMA stream + repack | softmax | output) overlap each other, which one CTA's phases cannot.
template <int NTILES> struct XCfg {
  static constexpr int D = 64;
  static constexpr int NT = NTILES;                          // 128-row tiles per CTA
  static constexpr int CONSUMERS = 128 * NT;                 // warps 0..15: repack / softmax / output (thread == tile row)
  static constexpr int THREADS = CONSUMERS + 32;             // warp 16: TMA producer
  static constexpr int CH = 16;                              // channels per TMA box = one tcgen05.mma K step
  static constexpr int NCS = D / CH;
  static constexpr int BOXES = 3 * NCS;                      // boxes per group: Q0 K0 .. Q3 K3 V0 .. V3
  static constexpr int BOX_BYTES = 64 * D * 2;               // one 64-token block of an operand tile
  static constexpr int TILE_BYTES = 2 * BOX_BYTES;           // 16 KB
  static constexpr int OFF_TQ = 0;                           // [NT] Q tiles, later V tiles
  static constexpr int OFF_TK = NT * TILE_BYTES;             // [NT] K tiles, later the output staging
  static constexpr int NSLOT = NT == 4 ? 3 : 2;
  static constexpr int SLOT_BYTES = NT == 4 ? 25600 : 19200; // >= BX * RG * CH * 2
  static constexpr int CTAS_PER_SM = NT == 4 ? 1 : 2;
  static constexpr int TMEM_COLS = 128 * NT;
  static constexpr int OFF_ST = 2 * NT * TILE_BYTES;
  static constexpr int OFF_ROWTAB = OFF_ST + NSLOT * SLOT_BYTES;   // uint32[128]: static (rowidx << 16 | column) of a tile row
  static constexpr int OFF_ROWOFF = OFF_ROWTAB + 128 * 4;          // int[64]: global token offset of every (y, z) row of the group, -1 = padding
  static constexpr int OFF_BAR = OFF_ROWOFF + 64 * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
  static_assert(CTAS_PER_SM * (SMEM_BYTES + 1024) <= 228 * 1024, "shared memory budget");
};

struct XParams {
  const void *q, *k, *v;
  void* y;
  float *l, *m;
  Geo g;
  int G, nwc, TW, RG;          // windows per tile / per group, tokens per (y, z) row of a group, rows per window
  int gpr;                     // groups per window row
  int BXs, BXl;                // box extents along x (tokens): small when shift + TW fits, large otherwise
  int NP, TWP;                 // output: 4-byte pairs per run (TW / 2 + 1), tokens per staged row (2 * NP)
  unsigned magicNP;            // ceil(2^32 / NP): exact i / NP for i < RG * NP
  long long ngroups;
  float scale_log2;
  int out_tma;                 // output by ONE TMA reduce-add box per group into a zero-initialised y (FA_WINX_OUT=0: per-thread 4-byte stores)
  int l2_prefetch;             // producer prefetches the next group's boxes into L2 (FA_WINX_PF=0 disables)
  int dbg;                     // FA_TRACE builds: FA_WINX_DBG experiment switch (0 = product behaviour)
  long long* trace;            // FA_TRACE builds: CTA 1 records clock64() per phase of its first 16 groups (20 events each)
};

#ifdef FA_TRACE
#define XTRACE(ev) do { if (prm.trace && blockIdx.x == 1 && tid == 0 && it < 16) prm.trace[it * 20 + (ev)] = clock64(); } while (0)
#else
#define XTRACE(ev) do {} while (0)
#endif

struct GroupInfo {
  long long gw0;               // linear index (batch included) of the group's first window
  int nvalid;                  // windows of the group that exist (the last group of a window row may be short)
  int x0, y0, z0, b;           // first token of the group per dim (may be negative = padding), batch element
  int shift, BX;               // x0 - box start (0..7), box extent used
};

__device__ __forceinline__ GroupInfo decode_group(const XParams& prm, unsigned grp) {   // ngroups < 2^31: 32-bit divisions
  const Geo& g = prm.g;
  GroupInfo gi;
  const unsigned rowi = grp / (unsigned)prm.gpr;
  const int gix = (int)(grp - rowi * (unsigned)prm.gpr);
  const int wx0 = gix * prm.nwc;
  gi.nvalid = g.o[0] - wx0 < prm.nwc ? g.o[0] - wx0 : prm.nwc;
  const unsigned r2 = rowi / (unsigned)g.o[1];
  const int wy = (int)(rowi - r2 * (unsigned)g.o[1]);
  const unsigned r3 = r2 / (unsigned)g.o[2];
  const int wz = (int)(r2 - r3 * (unsigned)g.o[2]);
  gi.b = (int)r3;
  gi.gw0 = (((long long)gi.b * g.o[2] + wz) * g.o[1] + wy) * g.o[0] + wx0;
  gi.x0 = wx0 * g.stride - g.padv[0];
  gi.y0 = g.nd >= 2 ? wy * g.stride - g.padv[1] : 0;
  gi.z0 = g.nd >= 3 ? wz * g.stride - g.padv[2] : 0;
  gi.shift = (gi.x0 + 1024) & 7;
  gi.BX = (gi.shift + gi.nvalid * g.W <= prm.BXs) ? prm.BXs : prm.BXl;
  return gi;
}

template <int FMT, int NTILES>
__global__ void __launch_bounds__(XCfg<NTILES>::THREADS, XCfg<NTILES>::CTAS_PER_SM)
tc_winx_fwd_kernel(const __grid_constant__ CUtensorMap tq_s, const __grid_constant__ CUtensorMap tk_s,
                   const __grid_constant__ CUtensorMap tv_s, const __grid_constant__ CUtensorMap tq_l,
                   const __grid_constant__ CUtensorMap tk_l, const __grid_constant__ CUtensorMap tv_l,
                   const __grid_constant__ CUtensorMap ty_s, const __grid_constant__ CUtensorMap ty_l,
                   const XParams prm) {
  using C = XCfg<NTILES>;
  using T = typename El<FMT>::type;
  constexpr int D = C::D;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sptr = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sptr);
  uint32_t* rowtab = reinterpret_cast<uint32_t*>(sptr + C::OFF_ROWTAB);
  int* rowoff = reinterpret_cast<int*>(sptr + C::OFF_ROWOFF);
  const uint32_t bar_s = sbase + C::OFF_BAR, bar_o = bar_s + 8, tmem_slot = bar_s + 16;
  const uint32_t bar_full = bar_s + 32, bar_empty = bar_s + 64;      // [NSLOT] each: box landed / every consumer warp has read it
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Geo& g = prm.g;
  const long long N = g.N;
  const int W = g.W, WD = g.WD, RG = prm.RG, TW = prm.TW;

  if (tid == 0) {
    mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    for (int i = 0; i < C::NSLOT; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, C::CONSUMERS / 32); }
    fence_barrier_init();
    prefetch_tensormap(&tq_s); prefetch_tensormap(&tk_s); prefetch_tensormap(&tv_s);
    prefetch_tensormap(&tq_l); prefetch_tensormap(&tk_l); prefetch_tensormap(&tv_l);
    if (prm.out_tma) { prefetch_tensormap(&ty_s); prefetch_tensormap(&ty_l); }
  }
  // static table of a tile row: which (y, z) row of the window and which column of the group's x-range it reads
  if (tid < 128) {
    const int r = tid;
    uint32_t e = 0xffffffffu;
    if (r < prm.G * WD) {
      const int wi = r / WD, slot = r - wi * WD;
      const int rowidx = slot / W, kx = slot - rowidx * W;
      e = ((uint32_t)rowidx << 16) | (uint32_t)(wi * W + kx);
    }
    rowtab[r] = e;
  }
  if (warp == 0) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // ---- warp 16: the TMA stream.  Box number `pos` of this CTA = box (pos % BOXES) of its (pos / BOXES)-th group, into
  //      slot pos % NSLOT as soon as every consumer warp has released that slot's previous box.
  if (warp == C::CONSUMERS / 32) {
    if (lane == 0) {
      uint32_t pos = 0;
      for (unsigned grp = blockIdx.x; grp < (unsigned)prm.ngroups; grp += gridDim.x) {
        const GroupInfo gi = decode_group(prm, grp);
        const bool small = gi.BX == prm.BXs;
        const uint32_t bytes = (uint32_t)(gi.BX * RG * C::CH * 2);
        // The stream can only run NSLOT boxes (77 KB) ahead, which at DRAM latency is ~30 B/clk: pull the NEXT group's
        // boxes into L2 now, a whole group time ahead, so that its loads are L2 hits.
        if (grp + gridDim.x < (unsigned)prm.ngroups && prm.l2_prefetch) {
          const GroupInfo gn = decode_group(prm, grp + gridDim.x);
          const bool sm2 = gn.BX == prm.BXs;
#pragma unroll 1
          for (int j = 0; j < C::BOXES; ++j) {
            const int x = j < 2 * C::NCS ? (j & 1) : 2, cs = j < 2 * C::NCS ? (j >> 1) : j - 2 * C::NCS;
            const CUtensorMap* tm = x == 0 ? (sm2 ? &tq_s : &tq_l) : x == 1 ? (sm2 ? &tk_s : &tk_l) : (sm2 ? &tv_s : &tv_l);
            tma_prefetch_5d(tm, gn.x0 - gn.shift, gn.y0, gn.z0, cs * C::CH, gn.b);
          }
        }
#pragma unroll 1
        for (int j = 0; j < C::BOXES; ++j, ++pos) {
          const int x = j < 2 * C::NCS ? (j & 1) : 2, cs = j < 2 * C::NCS ? (j >> 1) : j - 2 * C::NCS;
          const CUtensorMap* tm = x == 0 ? (small ? &tq_s : &tq_l) : x == 1 ? (small ? &tk_s : &tk_l) : (small ? &tv_s : &tv_l);
          const uint32_t slot = pos % C::NSLOT, fill = pos / C::NSLOT;
          if (fill > 0) mbar_wait(bar_empty + 8u * slot, (fill - 1) & 1u);
          fence_proxy_async();                             // the consumers' generic reads of the slot precede this async write
          mbar_arrive_expect_tx(bar_full + 8u * slot, bytes);
          tma_load_5d(sbase + C::OFF_ST + slot * C::SLOT_BYTES, tm, bar_full + 8u * slot, gi.x0 - gi.shift, gi.y0, gi.z0, cs * C::CH, gi.b);
        }
      }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();                                       // matches the barrier in front of the TMEM deallocation
    return;
  }
  auto csync = [] { asm volatile("bar.sync 1, %0;" ::"n"(C::CONSUMERS) : "memory"); };    // barrier of the 16 consumer warps

  // ---- roles
  const int ti = tid >> 7, row = tid & 127;                 // softmax / output: thread == row `row` of tile `ti`
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tS = tmem_base + lane_addr + ti * 128, tO = tS + 64;
  const int wi = row / WD, slot_in_win = row - wi * WD;
  const int c_lo = wi * WD, c_hi = c_lo + WD;               // columns of this row's own window
  // repack: this thread produces the 16-byte chunk rp_j (tile rows 8 rp_j .. 8 rp_j + 7) of channels rp_c and rp_c + 8
  // of tile rp_t.  Lane bits = (c0, c1, j2, t0, t1): the 8 lanes of a quarter warp store to 8 different 16-byte bank
  // groups ((j ^ c) & 7 distinct) and the 32 two-byte loads of a warp spread over the four windows of the group --
  // 2.2 wavefronts per load instead of 6.5 with lanes along j (64-byte staging rows alias to two bank groups).
  // (two tiles: lane bits = (c0, c1, c2, j2, t0), warp bits = (j0, j1, j3): 2.0 wavefronts per load)
  const int rp_t = C::NT == 4 ? (lane >> 3) & 3 : lane >> 4;
  const int rp_j = (warp & 3) | ((C::NT == 4 ? (lane >> 2) & 1 : (lane >> 3) & 1) << 2) | (((warp >> 2) & 1) << 3);
  const int rp_c = C::NT == 4 ? (lane & 3) | (((warp >> 3) & 1) << 2) : lane & 7;
  constexpr uint32_t idesc_qk = make_idesc_f16(FMT, FMT, 1, 1, 128, 128);
  constexpr uint32_t idesc_pv = make_idesc_f16(FMT, FMT, 0, 0, 128, D);
  const float2 scale2 = make_float2(prm.scale_log2, prm.scale_log2);

  uint32_t it = 0;
  for (unsigned grp = blockIdx.x; grp < (unsigned)prm.ngroups; grp += gridDim.x, ++it) {
    const GroupInfo gi = decode_group(prm, grp);
    const uint32_t pitch = (uint32_t)(RG * gi.BX * 2);      // channel pitch of a staged box
    // per-group repack recipe of this thread: byte offsets of its 8 source tokens inside one channel of a box
    uint32_t off[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t e = rowtab[rp_j * 8 + i];
      const int col = (int)(e & 0xffffu) + rp_t * prm.G * W;  // column inside the group's x-range
      const bool ok = e != 0xffffffffu && col < gi.nvalid * W;
      off[i] = ok ? (uint32_t)(((int)(e >> 16) * gi.BX + gi.shift + col) * 2) : 0xffffffffu;
    }
    if (tid < RG) {                                         // global token offset of every (y, z) row of the group
      const int kz = tid / W, ky = tid - kz * W;
      const int y = gi.y0 + (g.nd >= 2 ? ky : 0), z = gi.z0 + (g.nd >= 3 ? kz : 0);
      rowoff[tid] = (y >= 0 && y < g.s[1] && z >= 0 && z < g.s[2]) ? (z * g.s[1] + y) * g.s[0] : -1;
    }

    auto repack = [&](uint32_t pos, uint32_t tile_base, int cs) {
      const uint32_t slot = pos % C::NSLOT;
      mbar_wait(bar_full + 8u * slot, (pos / C::NSLOT) & 1u);
      const uint32_t stg = sbase + C::OFF_ST + slot * C::SLOT_BYTES;
#ifdef FA_TRACE
      if (prm.dbg == 4) { __syncwarp(); if (lane == 0) mbar_arrive(bar_empty + 8u * slot); return; }   // TMA stream alone
#endif
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cc = rp_c + 8 * h, c = cs * C::CH + cc;
        const uint32_t src = stg + (uint32_t)cc * pitch;
        uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = off[i] != 0xffffffffu ? lds_u16(src + off[i]) : 0u;
        const uint32_t dst = tile_base + (uint32_t)(rp_t * C::TILE_BYTES + (rp_j >> 3) * C::BOX_BYTES + c * 128 + (((rp_j & 7) ^ (c & 7)) << 4));
        sts_v4(dst, v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
      }
      __syncwarp();                                         // the warp's loads of the slot have completed (their values were stored)
      if (lane == 0) mbar_arrive(bar_empty + 8u * slot);
    };

    const uint32_t pos0 = it * C::BOXES;
    XTRACE(0);                                              // group start (recipe built)
    // ---- Q / K slices: repack, then one K = 16 step of S = Q K^T per tile
#pragma unroll 1
    for (int cs = 0; cs < C::NCS; ++cs) {
      repack(pos0 + 2 * cs, sbase + C::OFF_TQ, cs);
      repack(pos0 + 2 * cs + 1, sbase + C::OFF_TK, cs);
      fence_proxy_async();
      csync();
      XTRACE(1 + cs);                                       // Q/K slice cs repacked
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < C::NT; ++t) {
            const uint64_t qdesc = make_smem_desc_sw128(sbase + C::OFF_TQ + t * C::TILE_BYTES, C::BOX_BYTES, 1024);
            const uint64_t kdesc = make_smem_desc_sw128(sbase + C::OFF_TK + t * C::TILE_BYTES, C::BOX_BYTES, 1024);
            mma_ss(tmem_base + t * 128, qdesc + (uint64_t)(cs * 128), kdesc + (uint64_t)(cs * 128), idesc_qk, cs > 0 ? 1u : 0u);
          }
          if (cs == C::NCS - 1) tc_commit(bar_s);
        }
        __syncwarp();
      }
    }
    mbar_wait(bar_s, it & 1u);                              // S complete: the Q and K tiles are dead
    tc_fence_after();
    XTRACE(5);                                              // S complete
    // ---- V slices into the Q tiles
#pragma unroll 1
    for (int cs = 0; cs < C::NCS; ++cs) {
      repack(pos0 + 2 * C::NCS + cs, sbase + C::OFF_TQ, cs);
      XTRACE(6 + cs);                                       // V slice cs repacked
    }
    fence_proxy_async();

    // ---- softmax over the columns of this row's window (block-diagonal mask), as in fa_tc_win.cu
    const long long gw = gi.gw0 + (long long)ti * prm.G + wi;
    const bool valid = wi < prm.G && ti * prm.G + wi < gi.nvalid;
    const int r_lo = valid ? c_lo : 0, r_hi = valid ? c_hi : 128;
    uint32_t cls = 0;                                       // 2 bits per 32-column chunk: 1 = inside, 2 = outside, 0 = mixed
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
      tmem_st16(tS + 16 * ch, pk);                          // P (16-bit) over the S columns already consumed
    }
    const float lsum = ls2.x + ls2.y;
    tmem_wait_st();
    tc_fence_before();
    csync();
    XTRACE(10);                                             // softmax done (all warps)

    // ---- O = P V per tile (A = P from TMEM, B = V K-major in the Q tiles, K = 128 keys)
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < C::NT; ++t) {
          const uint64_t vdesc = make_smem_desc_sw128(sbase + C::OFF_TQ + t * C::TILE_BYTES, 16, 1024);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            mma_ts(tmem_base + t * 128 + 64, tmem_base + t * 128 + ks * 8,
                   vdesc + (uint64_t)((ks >> 2) * (C::BOX_BYTES >> 4) + (ks & 3) * 2), idesc_pv, ks > 0 ? 1u : 0u);
        }
        tc_commit(bar_o);
      }
      __syncwarp();
    }
    if (valid) {                                            // l, m for every slot, padded ones included
      prm.l[gw * WD + slot_in_win] = lsum;
      prm.m[gw * WD + slot_in_win] = m2 * LN2;
    }
    const float inv_l = valid ? 1.f / lsum : 0.f;
    mbar_wait(bar_o, it & 1u);
    tc_fence_after();
    XTRACE(11);                                             // O ready

    if (prm.out_tma) {
      // ---- output by TMA: the group's O rows are staged as ONE box [64 channels][(y, z) row][BX tokens] in the geometry of
      //      the input boxes (same 16-byte aligned start x0 - shift), zeros in the columns that belong to the neighbouring
      //      groups, and added to the zero-initialised y by a single cp.reduce.async.bulk.tensor: whole sectors, no store
      //      instructions, out-of-volume rows / columns dropped by the TMA unit.  (Every token belongs to exactly one window:
      //      the other groups add +0.)
      const uint32_t ostg = sbase + C::OFF_TQ;
      const uint32_t cpitch = pitch * 1u;                     // RG * BX * 2 bytes per channel
      {
        const uint32_t nvec = (uint32_t)D * cpitch / 16u;     // BX is a multiple of 8 tokens
        for (uint32_t i = tid; i < nvec; i += C::CONSUMERS) sts_v4(ostg + 16u * i, 0u, 0u, 0u, 0u);
      }
      csync();
      // TMA stores take no negative start coordinates (tools/probes/tma_reduce_probe.cu: illegal instruction, whereas
      // coordinates past the upper bound are clipped): the box starts at the clipped origin (xs, ys, zs) and the tokens
      // are staged relative to it; padding tokens in front of the volume are not staged, the rows / columns the shifted
      // box gains at its far end stay zero.
      const int xs = gi.x0 - gi.shift < 0 ? 0 : gi.x0 - gi.shift;       // multiple of 8 tokens
      const int dy = gi.y0 < 0 ? -gi.y0 : 0, dz = gi.z0 < 0 ? -gi.z0 : 0;
      {
        const uint32_t e = valid ? rowtab[row] : 0u;
        const int rowidx = (int)(e >> 16), kz = rowidx / W, ky = rowidx - kz * W;
        const int col = gi.x0 + ti * prm.G * W + (int)(e & 0xffffu) - xs;
        const bool st = valid && col >= 0 && ky >= dy && kz >= dz;
        const uint32_t o0 = ostg + (uint32_t)((((kz - dz) * W + (ky - dy)) * gi.BX + col) * 2);
#pragma unroll 1
        for (int ch = 0; ch < D / 32; ++ch) {
          uint32_t o[32];
          tmem_ld32(tO + 32 * ch, o);
          tmem_wait_ld();
          if (st) {
#pragma unroll
            for (int e2 = 0; e2 < 32; ++e2) sts_u16(o0 + (uint32_t)(32 * ch + e2) * cpitch, cvt16<FMT>(__uint_as_float(o[e2]) * inv_l));
          }
        }
      }
      tc_fence_before();
      fence_proxy_async();
      csync();
      XTRACE(12);                                             // O staged
      if (tid == 0) {
        tma_reduce_add_5d(gi.BX == prm.BXs ? &ty_s : &ty_l, ostg, xs, gi.y0 + dy, gi.z0 + dz, 0, gi.b);
        bulk_commit();
        bulk_wait_read0();                                    // the staging (= V and K tiles) may be overwritten
      }
    } else {
    // ---- O rows -> 16-bit output staging [channel][(y, z) row][x across the group] over the dead V and K tiles.  Token x
    //      of a row sits at index x - x0 + par (par = x0 & 1; rows start at even global element indices because the x
    //      extent is even), so that a pair of tokens at an even global index is one aligned 32-bit word here.
    const uint32_t ostg = sbase + C::OFF_TQ;
    const int par = gi.x0 & 1;
    const uint32_t cpitch = (uint32_t)(RG * prm.TWP * 2);
    {
      // (tcgen05.ld is warp-collective: every thread reads its row, only rows of existing windows are stored)
      const uint32_t e = valid ? rowtab[row] : 0u;
      const uint32_t o0 = ostg + (uint32_t)(((int)(e >> 16) * prm.TWP + ti * prm.G * W + (int)(e & 0xffffu) + par) * 2);
#pragma unroll 1
      for (int ch = 0; ch < D / 32; ++ch) {
        uint32_t o[32];
        tmem_ld32(tO + 32 * ch, o);
        tmem_wait_ld();
        if (valid) {
#pragma unroll
          for (int e2 = 0; e2 < 32; ++e2) sts_u16(o0 + (uint32_t)(32 * ch + e2) * cpitch, cvt16<FMT>(__uint_as_float(o[e2]) * inv_l));
        }
      }
    }
    tc_fence_before();
    csync();
    XTRACE(12);                                             // O staged

    // ---- write the runs: item = (row, pair) of one channel, pair = 2 tokens at an even global element index = one
    //      32-bit shared load + one 4-byte store (2-byte stores at the clipped ends of a run).  Lanes run along the pairs
    //      of a row, then the rows; warp w writes the channels w, w + 16, w + 32, w + 48.
    {
      const int xa = gi.x0 < 0 ? 0 : gi.x0;
      const int xe_ = gi.x0 + gi.nvalid * W;
      const int xe = xe_ > g.s[0] ? g.s[0] : xe_;
      T* ybase = static_cast<T*>(prm.y) + ((long long)gi.b * D + warp) * N;
      const uint32_t sw = ostg + (uint32_t)warp * cpitch;
      const long long cstep = (long long)(C::CONSUMERS / 32) * N;
      const uint32_t sstep = (uint32_t)(C::CONSUMERS / 32) * cpitch;
      int nitems = RG * prm.NP;
#ifdef FA_TRACE
      if (prm.dbg == 3) nitems = 0;                         // experiment: no output loop at all
#endif
      for (int i = lane; i < nitems; i += 32) {
        const int rowidx = (int)__umulhi((unsigned)i, prm.magicNP), p = i - rowidx * prm.NP;
        const int ro = rowoff[rowidx];
        const int t0 = gi.x0 - par + 2 * p;                 // x of the pair's first token
        const bool h0 = t0 >= xa && t0 < xe, h1 = t0 + 1 >= xa && t0 + 1 < xe;
        if (ro < 0 || (!h0 && !h1)) continue;
        uint32_t sa = sw + (uint32_t)((rowidx * prm.TWP + 2 * p) * 2);
        T* yp = ybase + ro + t0;
        if (h0 && h1) {
#pragma unroll
          for (int c4 = 0; c4 < D / (C::CONSUMERS / 32); ++c4, sa += sstep, yp += cstep) {
            uint32_t v;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(sa));
#ifdef FA_TRACE
            if (prm.dbg == 1 && v != 0x7fc12345u) continue;   // experiment: shared loads only
            if (prm.dbg == 5) { yp = static_cast<T*>(prm.y) + (long long)(((blockIdx.x * 16 + warp) * 4 + c4) * 1024 + lane * 2); }   // experiment: dense, fully coalesced stores
#endif
            *reinterpret_cast<uint32_t*>(yp) = v;
          }
        } else {
          const uint32_t sa1 = sa + (h0 ? 0u : 2u);
          T* yp1 = yp + (h0 ? 0 : 1);
#pragma unroll
          for (int c4 = 0; c4 < D / (C::CONSUMERS / 32); ++c4)
            *reinterpret_cast<unsigned short*>(yp1 + c4 * cstep) = (unsigned short)lds_u16(sa1 + c4 * sstep);
        }
      }
    }
    }
    csync();              // output staging (= K tiles), row table and TMEM are free for the next group
    XTRACE(13);                                             // runs written
  }

  if (prm.out_tma && tid == 0) bulk_wait0();               // every reduce of this CTA has been performed
  tc_fence_before();
  __syncthreads();                                         // all 17 warps (the producer arrives from its own branch)
  if (warp == 0) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int FMT, int NTILES>
int launch_winx(const Geo& g, const FwdArgs& a, XParams& prm, cudaStream_t st) {
  using C = XCfg<NTILES>;
  const int dtype = FMT ? FA_BF16 : FA_F16;
  CUtensorMap tm[8];
  memset(tm, 0, sizeof(tm));
  const void* src[3] = {a.q, a.k, a.v};
  const int by = g.nd >= 2 ? g.W : 1, bz = g.nd >= 3 ? g.W : 1;
  for (int x = 0; x < 3; ++x) {
    int rc;
    if ((rc = make_win_tmap_box(&tm[x], src[x], dtype, g, C::D, prm.BXs, by, bz, C::CH))) return rc;
    if ((rc = make_win_tmap_box(&tm[3 + x], src[x], dtype, g, C::D, prm.BXl, by, bz, C::CH))) return rc;
  }
  if (prm.out_tma) {
    int rc;
    if ((rc = make_win_tmap_box(&tm[6], a.o, dtype, g, C::D, prm.BXs, by, bz, C::D))) return rc;
    if ((rc = make_win_tmap_box(&tm[7], a.o, dtype, g, C::D, prm.BXl, by, bz, C::D))) return rc;
    FA_CUDA_TRY(cudaMemsetAsync(a.o, 0, (size_t)g.N * C::D * g.B * 2, st));
  }
  auto kern = tc_winx_fwd_kernel<FMT, NTILES>;
  FA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long cap = (long long)sms * C::CTAS_PER_SM;
  const unsigned grid = (unsigned)(prm.ngroups < cap ? prm.ngroups : cap);
  kern<<<grid, C::THREADS, C::SMEM_BYTES, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], tm[7], prm);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

}  // namespace

// Does the streamed kernel take this call?  Exact-cover windows (stride == W), d = dv = 64, 16-bit, TMA-legal rows
// (x extent a multiple of 8 tokens, 16-byte aligned bases), no fold accumulator, boxes that fit a staging slot.
template <int NTILES>
static bool winx_geo_ok(const Geo& g) {
  using C = XCfg<NTILES>;
  const int G = 128 / g.WD, nwc = C::NT * G, TW = nwc * g.W, RG = g.WD / g.W;
  const int BXl = (TW + 7 + 7) / 8 * 8;
  if (RG > 64 || BXl > 256 || (long long)BXl * RG * C::CH * 2 > C::SLOT_BYTES) return false;   // TMA box extents <= 256
  if ((long long)C::D * RG * (TW + 2) * 2 > 2LL * C::NT * C::TILE_BYTES) return false;      // output staging over the V and K tiles
  if ((long long)((g.o[0] + nwc - 1) / nwc) * g.o[1] * g.o[2] * g.B > 0x7fffffffLL) return false;      // 32-bit group index
  return true;
}
int winx_tiles() { static const int nt = [] { const char* e = getenv("FA_WINX_NT"); const int v = e ? atoi(e) : 4; return v == 2 ? 2 : 4; }(); return nt; }

bool tc_winx_supported(const Geo& g, const FwdArgs& a, int dtype) {
  if (dtype != FA_BF16 && dtype != FA_F16) return false;
  if (g.mode != MODE_WINDOWED || g.d != 64 || g.dv != 64 || a.acc) return false;
  if (g.stride != g.W || g.s[0] % 8 != 0 || g.WD > 128 || g.WD < 1) return false;
  if ((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v)) & 15) return false;
  if (reinterpret_cast<uintptr_t>(a.o) & 3) return false;
  if ((g.N & 1) || g.N * 64 > 0x7fffffffLL || g.padv[0] > 1000) return false;
  return winx_tiles() == 4 ? winx_geo_ok<4>(g) : winx_geo_ok<2>(g);
}

int tc_winx_fwd(const Geo& g, const FwdArgs& a, int dtype, cudaStream_t st) {
  if (!tc_winx_supported(g, a, dtype)) { set_error("tc_winx_fwd: unsupported configuration"); return FA_ERR_UNSUPPORTED; }
  const int NT = winx_tiles();
  XParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.q = a.q; prm.k = a.k; prm.v = a.v; prm.y = a.o; prm.l = a.l; prm.m = a.m;
  prm.g = g;
  prm.G = 128 / g.WD; prm.nwc = NT * prm.G; prm.TW = prm.nwc * g.W; prm.RG = g.WD / g.W;
  prm.gpr = (g.o[0] + prm.nwc - 1) / prm.nwc;
  prm.BXs = (prm.TW + 7) / 8 * 8;
  prm.BXl = (prm.TW + 7 + 7) / 8 * 8;
  prm.NP = prm.TW / 2 + 1;
  prm.TWP = 2 * prm.NP;
  prm.magicNP = (unsigned)((0x100000000ULL + prm.NP - 1) / prm.NP);
  prm.ngroups = (long long)prm.gpr * g.o[1] * g.o[2] * g.B;
  prm.scale_log2 = g.tau * LOG2E;
  prm.trace = nullptr;
  prm.dbg = 0;
  { static const int pf = [] { const char* e = getenv("FA_WINX_PF"); return e ? atoi(e) : 1; }(); prm.l2_prefetch = pf; }
  {
    // FA_WINX_OUT=1: one TMA reduce-add box per group instead of the 4-byte store loop (needs a 16-byte aligned y and a box
    // of all 64 channels inside the dead V + K tiles).  Measured (profiles/r2f_winx.md): the output phase of a group drops
    // from 10 000 to 5 200 clk, but the zero-fill pass over y and the reduction's sector fills cost more at config 5, B = 64
    // (7.77 vs 6.58 ms); one 256^3 volume gains 5 % (4.44 vs 4.66 ms).  Off by default.
    static const int ot = [] { const char* e = getenv("FA_WINX_OUT"); return e ? atoi(e) : 0; }();
    const int by = g.nd >= 2 ? g.W : 1, bz = g.nd >= 3 ? g.W : 1;
    const long long box = (long long)prm.BXl * by * bz * 64 * 2;
    prm.out_tma = ot && NT == 4 && (reinterpret_cast<uintptr_t>(a.o) & 15) == 0 && box <= 2LL * NT * 16384;
  }
#ifdef FA_TRACE
  { const char* e = getenv("FA_WINX_DBG"); prm.dbg = e ? atoi(e) : 0; }
  { const char* e = getenv("FA_TRACE_PTR"); prm.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr; }
#endif
  if (NT == 4) return dtype == FA_BF16 ? launch_winx<1, 4>(g, a, prm, st) : launch_winx<0, 4>(g, a, prm, st);
  return dtype == FA_BF16 ? launch_winx<1, 2>(g, a, prm, st) : launch_winx<0, 2>(g, a, prm, st);
}

int tc_winx_groups_per_window_row(const Geo& g) { const int nwc = winx_tiles() * (128 / g.WD); return (g.o[0] + nwc - 1) / nwc; }

}  // namespace fa
