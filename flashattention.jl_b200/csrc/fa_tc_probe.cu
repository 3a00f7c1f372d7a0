// fa_tc_probe.cu -- diagnostics: one 128x128 UMMA tile with caller-supplied descriptor fields.
// Used by tests/test_gpu_probe.py to pin the shared-memory descriptor conventions (MN-major
// SWIZZLE_128B for Q/K, K-major SWIZZLE_128B for V, P operand in TMEM) on real hardware,
// independently of the full attention pipeline.  Not part of the product path.
#include <cuda.h>
#include "fa_common.cuh"
#include "fa_ptx.cuh"

namespace fa {
int make_tmap_public(CUtensorMap* tm, const void* base, int dtype, long long N, int D, long long B, long long stride_c = 0, long long stride_b = 0);

namespace {
using namespace ptx;

struct ProbeParams {
  int mode;         // 0: S = A^T B (both MN-major from smem);  1: O = P V (P in TMEM, V K-major smem)
  int D;            // channels (64 or 128)
  int fmt;          // 0 f16, 1 bf16 (smem operands)
  int afmt;         // mode 1: format of P in TMEM (0 f16, 1 bf16)
  int lbo, sbo;     // descriptor byte offsets for the smem operands
  int kstep;        // byte advance of the start address per K=16 step (within a 64-wide box)
  int kbox;         // k-steps per box before jumping by box_bytes (mode 1); 0 = never
  const float* p;   // mode 1: P [128][128] fp32 row-major
  float* out;       // mode 0: [128][128]; mode 1: [128][D]   (row-major fp32)
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const ProbeParams prm) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int D = prm.D;
  const uint32_t box = 64u * D * 2u, tile = 2u * box;
  const uint32_t sA = sbase, sB = sbase + tile, bars = sbase + 2 * tile, slot = bars + 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bars, 1); mbar_init(bars + 8, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;

  if (prm.mode == 1) {
    // P row of this thread -> 16-bit pairs -> TMEM columns [0,64)
    const float* pr = prm.p + (size_t)(warp * 32 + lane) * 128;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = pr[32 * c + 2 * i], b = pr[32 * c + 2 * i + 1];
        if (prm.afmt == 1) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); pk[i] = *reinterpret_cast<uint32_t*>(&h); }
        else { __half2 h = __floats2half2_rn(a, b); pk[i] = *reinterpret_cast<uint32_t*>(&h); }
      }
      tmem_st16(tmem_base + lane_addr + 16 * c, pk);
    }
    tmem_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  if (threadIdx.x == 0) {
    const uint32_t bytes = prm.mode == 0 ? 2 * tile : tile;
    mbar_arrive_expect_tx(bars, bytes);
    if (prm.mode == 0) {
      tma_load_3d(sA, &tma_a, bars, 0, 0, 0);
      tma_load_3d(sA + box, &tma_a, bars, 64, 0, 0);
    }
    tma_load_3d(sB, &tma_b, bars, 0, 0, 0);
    tma_load_3d(sB + box, &tma_b, bars, 64, 0, 0);
    mbar_wait(bars, 0);
    tc_fence_after();
    if (prm.mode == 0) {
      const uint32_t idesc = make_idesc_f16(prm.fmt, prm.fmt, 1, 1, 128, 128);
      for (int ks = 0; ks < D / 16; ++ks) {
        const uint64_t ad = make_smem_desc_sw128(sA + ks * prm.kstep, prm.lbo, prm.sbo);
        const uint64_t bd = make_smem_desc_sw128(sB + ks * prm.kstep, prm.lbo, prm.sbo);
        mma_ss(tmem_base + 128, ad, bd, idesc, ks > 0);
      }
    } else {
      const uint32_t idesc = make_idesc_f16(prm.afmt, prm.fmt, 0, 0, 128, D);
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t off = prm.kbox ? (ks / prm.kbox) * box + (ks % prm.kbox) * prm.kstep : ks * prm.kstep;
        const uint64_t bd = make_smem_desc_sw128(sB + off, prm.lbo, prm.sbo);
        mma_ts(tmem_base + 128, tmem_base + ks * 8, bd, idesc, ks > 0);
      }
    }
    tc_commit(bars + 8);
  }
  __syncwarp();
  mbar_wait(bars + 8, 0);
  tc_fence_after();
  const int ncol = prm.mode == 0 ? 128 : D;
  float* orow = prm.out + (size_t)(warp * 32 + lane) * ncol;
#pragma unroll 1
  for (int c = 0; c < ncol / 32; ++c) {
    uint32_t r[32];
    tmem_ld32(tmem_base + lane_addr + 128 + 32 * c, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) orow[32 * c + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}
// TMEM bandwidth probe: `nwarps` warps (4 or 8; warp w owns lane quarter w % 4) each issue `iters`
// rounds of 4 x tcgen05.ld.32x32b.x32 (mode 0) or 4 x tcgen05.st.32x32b.x32 (mode 1) over 128 columns and
// report elapsed SM clocks.  bytes moved = nwarps * iters * 4 * 4096.
__global__ void tmem_bw_kernel(int mode, int iters, long long* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld32(base + 32 * c, r);
        tmem_wait_ld();
        acc += r[0] ^ r[31];
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st32(base + 32 * c, r);
      tmem_wait_st();
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x % 32 == 0) out[warp] = t1 - t0 + (acc == 0xdeadbeef ? 1 : 0);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}
// UMMA issue-rate / operand-bandwidth probe: `iters` rounds of 8 kind::f16 MMAs (M = 128, K = 16 each, one
// 128-channel K loop) with N = n_cols, A from shared memory (mode 0, MN-major SW128) or from TMEM (mode 1),
// B from shared memory (MN-major SW128).  Data is whatever the buffers hold; only the clocks matter.
__global__ void umma_rate_kernel(int mode, int n_cols, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = sbase, sB = sbase + 32768, bars = sbase + 32768 + 65536, slot = bars + 16;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bars, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 512);
  for (int i = threadIdx.x; i < (32768 + 65536) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_raw + (sbase - smem_u32(smem_raw)))[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  if (warp == 1) {
    // mode bit 0: A from TMEM; bit 1: A K-major in smem; bit 2: B K-major in smem (else MN-major)
    const bool a_tmem = mode & 1, a_km = mode & 2, b_km = mode & 4;
    const uint32_t idesc = make_idesc_f16(0, 0, (a_tmem || a_km) ? 0 : 1, b_km ? 0 : 1, 128, n_cols);
    const uint64_t ad = a_km ? make_smem_desc_sw128(sA, 16, 1024) : make_smem_desc_sw128(sA, 16384, 1024);
    const uint64_t bd = b_km ? make_smem_desc_sw128(sB, 16, 1024) : make_smem_desc_sw128(sB, 16384, 1024);
    // K-major: 64 K-elements per 128-byte row -> k-steps 0..3 advance 32 B inside a tile, 4..7 sit in the next tile
    auto a_off = [&](int ks) { return a_km ? (uint64_t)((ks >> 2) * (16384 >> 4) + (ks & 3) * 2) : (uint64_t)(ks * 128); };
    auto b_off = [&](int ks) { return b_km ? (uint64_t)((ks >> 2) * (32768 >> 4) + (ks & 3) * 2) : (uint64_t)(ks * 128); };
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          if (a_tmem) mma_ts(tmem_base, tmem_base + 256 + ks * 8, bd + b_off(ks), idesc, ks > 0 ? 1u : 0u);
          else mma_ss(tmem_base, ad + a_off(ks), bd + b_off(ks), idesc, ks > 0 ? 1u : 0u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(bars);
    __syncwarp();
    mbar_wait(bars, 0);
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}
// 5-D TMA tile-mode probe: one box load at (c0..c4) into shared memory, copied out to `out` (bytes).
__global__ void tma5d_probe_kernel(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3, int c4,
                                   unsigned bytes, unsigned char* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sp = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar = sbase + 60 * 1024;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, bytes);
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(sbase), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
  }
  mbar_wait(bar, 0);
  for (unsigned i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sp[i];
}
}  // namespace

int tma5d_probe(const void* base, const long long* dims, const long long* strides_bytes, const int* box, const int* coord,
                unsigned char* out_dev, cudaStream_t st) {
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                         const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) { set_error("no encode fn"); return FA_ERR_CUDA; }
  CUtensorMap tm;
  cuuint64_t d[5], sb[4];
  cuuint32_t bx[5], es[5] = {1, 1, 1, 1, 1};
  unsigned bytes = 2;
  for (int i = 0; i < 5; ++i) { d[i] = (cuuint64_t)dims[i]; bx[i] = (cuuint32_t)box[i]; bytes *= (unsigned)box[i]; }
  for (int i = 0; i < 4; ++i) sb[i] = (cuuint64_t)strides_bytes[i];
  CUresult r = reinterpret_cast<Fn>(p)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), d, sb, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("encode failed %d", (int)r); return FA_ERR_CUDA; }
  const int smem = 62 * 1024 + 1024;
  FA_CUDA_TRY(cudaFuncSetAttribute(tma5d_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tma5d_probe_kernel<<<1, 128, smem, st>>>(tm, coord[0], coord[1], coord[2], coord[3], coord[4], bytes, out_dev);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

int umma_rate_probe(int mode, int n_cols, int iters, int blocks, long long* out_dev, cudaStream_t st) {
  const int smem = 32768 + 65536 + 64 + 1024;
  FA_CUDA_TRY(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_rate_kernel<<<blocks, 64, smem, st>>>(mode, n_cols, iters, out_dev);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

int tmem_bw_probe(int mode, int nwarps, int iters, long long* out_dev, cudaStream_t st) {
  tmem_bw_kernel<<<1, 32 * nwarps, 0, st>>>(mode, iters, out_dev);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

int tc_probe(int mode, const void* a, const void* b, const float* p, float* out, int D, int dtype,
             int lbo, int sbo, int kstep, int kbox, int afmt, cudaStream_t st) {
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_tmap_public(&ta, a ? a : b, dtype, 128, D, 1))) return rc;
  if ((rc = make_tmap_public(&tb, b, dtype, 128, D, 1))) return rc;
  ProbeParams prm;
  prm.mode = mode; prm.D = D; prm.fmt = dtype == FA_BF16 ? 1 : 0; prm.afmt = afmt < 0 ? prm.fmt : afmt;
  prm.lbo = lbo; prm.sbo = sbo; prm.kstep = kstep; prm.kbox = kbox; prm.p = p; prm.out = out;
  const int smem = 4 * 64 * D * 2 + 64 + 1024;
  FA_CUDA_TRY(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<1, 128, smem, st>>>(ta, tb, prm);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

}  // namespace fa
