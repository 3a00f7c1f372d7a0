// fa_softmax.cu -- standalone safe softmax along Julia dim 1 (columns, contiguous) or dim 2
// (rows, strided) of a column-major (M, N, B) array.  Replaces fused_softmax! / col_softmax! /
// row_softmax! (reference src/fused_softmax.jl:11-39).  HBM-bound: one read + one write when a column fits the
// registers of its thread group (dim 1, M up to 32768 Float32), two reads + one write otherwise.
#include <cooperative_groups.h>
#include "fa_common.cuh"

namespace fa {
namespace {

// (mx, sum) <- merge with a partial (om, os).  -inf entries (masked scores) contribute 0 and never produce
// exp(-inf - -inf) = NaN: the reference's fused_softmax! (src/fused_softmax.jl:20-24) gives 0 for them.
__device__ __forceinline__ void online_merge(float& mx, float& sum, float om, float os) {
  const float mn = fmaxf(mx, om);
  const float a = (mx == -INFINITY) ? 0.f : expf(mx - mn), b = (om == -INFINITY) ? 0.f : expf(om - mn);
  sum = sum * a + os * b;
  mx = mn;
}

// dim 1: one warp per (column n, batch b); lanes stride the contiguous M axis.
template <typename T>
__global__ void softmax_dim1_kernel(T* __restrict__ out, const T* __restrict__ in, long long M, long long cols) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long col = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); col < cols; col += warps) {
    const T* x = in + col * M;
    T* y = out + col * M;
    float mx = -INFINITY, sum = 0.f;
    for (long long i = lane; i < M; i += 32) {            // online max / sum
      online_merge(mx, sum, to_f32<T>(x[i]), 1.f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) online_merge(mx, sum, __shfl_xor_sync(0xffffffffu, mx, o), __shfl_xor_sync(0xffffffffu, sum, o));
    const float inv = 1.f / sum;
    for (long long i = lane; i < M; i += 32) y[i] = from_f32<T>(expf(to_f32<T>(x[i]) - mx) * inv);
  }
}

// dim 2: one thread per (row i, batch b); consecutive threads read consecutive rows (coalesced).
template <typename T>
__global__ void softmax_dim2_kernel(T* __restrict__ out, const T* __restrict__ in, long long M, long long N, long long B) {
  const long long total = M * B;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < total; r += (long long)gridDim.x * blockDim.x) {
    const long long i = r % M, b = r / M;
    const T* x = in + b * M * N + i;
    T* y = out + b * M * N + i;
    float mx = -INFINITY, sum = 0.f;
    for (long long n = 0; n < N; ++n) {
      online_merge(mx, sum, to_f32<T>(x[n * M]), 1.f);
    }
    const float inv = 1.f / sum;
    for (long long n = 0; n < N; ++n) y[n * M] = from_f32<T>(expf(to_f32<T>(x[n * M]) - mx) * inv);
  }
}

// dim 1, few long columns (the vector case of bench/softmax.jl:8-35): one 1024-thread block per column.
template <typename T>
__global__ void softmax_dim1_block_kernel(T* __restrict__ out, const T* __restrict__ in, long long M) {
  __shared__ float smx[32], ssum[32];
  const T* x = in + (long long)blockIdx.x * M;
  T* y = out + (long long)blockIdx.x * M;
  float mx = -INFINITY, sum = 0.f;
  for (long long i = threadIdx.x; i < M; i += blockDim.x) online_merge(mx, sum, to_f32<T>(x[i]), 1.f);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) online_merge(mx, sum, __shfl_xor_sync(0xffffffffu, mx, o), __shfl_xor_sync(0xffffffffu, sum, o));
  if ((threadIdx.x & 31) == 0) { smx[threadIdx.x >> 5] = mx; ssum[threadIdx.x >> 5] = sum; }
  __syncthreads();
  mx = -INFINITY; sum = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) online_merge(mx, sum, smx[w], ssum[w]);
  const float inv = 1.f / sum;
  for (long long i = threadIdx.x; i < M; i += blockDim.x) y[i] = from_f32<T>(expf(to_f32<T>(x[i]) - mx) * inv);
}

// dim 2, long rows: block = 8 consecutive rows x 128 slices of the strided N axis (a warp = 8 rows x 4
// slices reads four full 32-byte sectors per load); (max, sum) reduced over the slices through smem.
template <typename T>
__global__ void softmax_dim2_sliced_kernel(T* __restrict__ out, const T* __restrict__ in, long long M, long long N, long long B) {
  __shared__ float smx[128][8], ssum[128][8];
  const int rx = threadIdx.x & 7, sy = threadIdx.x >> 3;
  const long long groups = (M + 7) / 8;
  for (long long gb = blockIdx.x; gb < groups * B; gb += gridDim.x) {
    const long long b = gb / groups, i = (gb - b * groups) * 8 + rx;
    const bool ok = i < M;
    const T* x = in + b * M * N + i;
    T* y = out + b * M * N + i;
    float mx = -INFINITY, sum = 0.f;
    if (ok) for (long long n = sy; n < N; n += 128) online_merge(mx, sum, to_f32<T>(x[n * M]), 1.f);
    smx[sy][rx] = mx; ssum[sy][rx] = sum;
    __syncthreads();
    mx = -INFINITY; sum = 0.f;
    for (int w = 0; w < 128; ++w) online_merge(mx, sum, smx[w][rx], ssum[w][rx]);
    const float inv = 1.f / sum;
    if (ok) for (long long n = sy; n < N; n += 128) y[n * M] = from_f32<T>(expf(to_f32<T>(x[n * M]) - mx) * inv);
    __syncthreads();
  }
}

// dim 1, column held in REGISTERS: G threads per column (a warp, 256 or 1024 threads), each thread keeps up to 8 16-byte
// vectors of the column -- one read and one write of the array (the two-pass kernels above read it twice).  Covers
// the reference's logged column-softmax shapes (logs/sm_cuda.txt: M = 256 .. 8192, Float32).
template <typename T, int G>
__global__ void __launch_bounds__(G >= 256 ? G : 256)
softmax_dim1_cached_kernel(T* __restrict__ out, const T* __restrict__ in, long long M, long long cols) {
  constexpr int EPV = 16 / (int)sizeof(T);              // elements per 16-byte vector
  constexpr int CPB = G >= 256 ? 1 : 256 / G;           // columns per block
  __shared__ float smx[32], ssum[32];
  const int t = threadIdx.x % G, sub = threadIdx.x / G;
  const long long nvec = M / EPV;
  for (long long col = (long long)blockIdx.x * CPB + sub; col < cols; col += (long long)gridDim.x * CPB) {
    const uint4* x = reinterpret_cast<const uint4*>(in + col * M);
    uint4* y = reinterpret_cast<uint4*>(out + col * M);
    uint4 v[8];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long i = t + (long long)k * G;
      if (i < nvec) {
        v[k] = x[i];
        const T* e = reinterpret_cast<const T*>(&v[k]);
#pragma unroll
        for (int j = 0; j < EPV; ++j) mx = fmaxf(mx, to_f32<T>(e[j]));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (G > 32) {
      if ((threadIdx.x & 31) == 0) smx[threadIdx.x >> 5] = mx;
      __syncthreads();
      mx = -INFINITY;
      for (int w = 0; w < G / 32; ++w) mx = fmaxf(mx, smx[w]);
    }
    float sum = 0.f;                                     // (the exponentials are recomputed below instead of kept: registers)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long i = t + (long long)k * G;
      if (i < nvec) {
        const T* e = reinterpret_cast<const T*>(&v[k]);
#pragma unroll
        for (int j = 0; j < EPV; ++j) sum += expf(to_f32<T>(e[j]) - mx);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (G > 32) {
      if ((threadIdx.x & 31) == 0) ssum[threadIdx.x >> 5] = sum;
      __syncthreads();
      sum = 0.f;
      for (int w = 0; w < G / 32; ++w) sum += ssum[w];
      __syncthreads();                                   // smx / ssum are reused by the next column of this block
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long i = t + (long long)k * G;
      if (i < nvec) {
        uint4 o4;
        T* e = reinterpret_cast<T*>(&o4);
        const T* xi = reinterpret_cast<const T*>(&v[k]);
#pragma unroll
        for (int j = 0; j < EPV; ++j) e[j] = from_f32<T>(expf(to_f32<T>(xi[j]) - mx) * inv);
        y[i] = o4;
      }
    }
  }
}

template <typename T, int G>
static int launch_cached(void* out, const void* in, long long M, long long cols, cudaStream_t st) {
  constexpr int CPB = G >= 256 ? 1 : 256 / G;
  long long blocks = (cols + CPB - 1) / CPB;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  softmax_dim1_cached_kernel<T, G><<<(unsigned)blocks, G >= 256 ? G : 256, 0, st>>>(static_cast<T*>(out), static_cast<const T*>(in), M, cols);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

// dim 2, very long rows (N >= 4096): a thread-block CLUSTER of 8 CTAs shares one group of 8 consecutive rows, each CTA
// takes an eighth of the columns; the per-row (max, sum) partials are exchanged through distributed shared memory, so
// a row group gets 8 x 256 threads instead of 1024 and the grid fills the GPU (M = 1024: 1024 CTAs instead of 128).  The
// normalise pass re-reads the CTA's own 8 x N/8 slice, which the resident CTAs keep in L2 (148 SMs x ~0.5 MB).
template <typename T>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(256)
softmax_dim2_cluster_kernel(T* __restrict__ out, const T* __restrict__ in, long long M, long long N, long long B) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float smx[32][8], ssum[32][8];
  __shared__ float pmx[8], psum[8];                       // this CTA's partial per row, read by the 7 other CTAs
  const int rx = threadIdx.x & 7, sy = threadIdx.x >> 3;  // 8 rows x 32 column lanes
  const unsigned cr = cluster.block_rank();
  const long long groups = (M + 7) / 8;
  const long long gb = blockIdx.x / 8;                    // (row group, batch); grid.x = 8 * groups * B
  const long long b = gb / groups, i = (gb - b * groups) * 8 + rx;
  const long long Nc = (N + 7) / 8, n0 = cr * Nc, n1 = n0 + Nc < N ? n0 + Nc : N;
  const bool ok = i < M;
  const T* x = in + b * M * N + i;
  T* y = out + b * M * N + i;
  float mx = -INFINITY, sum = 0.f;
  if (ok) for (long long n = n0 + sy; n < n1; n += 32) online_merge(mx, sum, to_f32<T>(x[n * M]), 1.f);
  smx[sy][rx] = mx; ssum[sy][rx] = sum;
  __syncthreads();
  if (threadIdx.x < 8) {
    float a = -INFINITY, s2 = 0.f;
    for (int w = 0; w < 32; ++w) online_merge(a, s2, smx[w][threadIdx.x], ssum[w][threadIdx.x]);
    pmx[threadIdx.x] = a; psum[threadIdx.x] = s2;
  }
  cluster.sync();
  mx = -INFINITY; sum = 0.f;
  for (unsigned r = 0; r < 8; ++r) {
    const float* rm = cluster.map_shared_rank(pmx, r);
    const float* rs = cluster.map_shared_rank(psum, r);
    online_merge(mx, sum, rm[rx], rs[rx]);
  }
  cluster.sync();                                          // nobody leaves while its partials may still be read
  const float inv = 1.f / sum;
  if (ok) for (long long n = n0 + sy; n < n1; n += 32) y[n * M] = from_f32<T>(expf(to_f32<T>(x[n * M]) - mx) * inv);
}

template <typename T>
int launch(void* out, const void* in, long long M, long long N, long long B, int dim, cudaStream_t st) {
  if (dim == 1) {
    const long long cols = N * B;
    constexpr long long EPV = 16 / (long long)sizeof(T);
    const bool vec_ok = M % EPV == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (vec_ok && cols >= 148) {                       // single-read register-cached form
      const long long nvec = M / EPV;
      if (nvec <= 32 * 8) return launch_cached<T, 32>(out, in, M, cols, st);
      if (nvec <= 256 * 8) return launch_cached<T, 256>(out, in, M, cols, st);
      if (nvec <= 1024 * 8) return launch_cached<T, 1024>(out, in, M, cols, st);
    }
    if (cols < 148 * 2 && M >= 4096) {      // few long columns: a block per column
      softmax_dim1_block_kernel<T><<<(unsigned)cols, 1024, 0, st>>>(static_cast<T*>(out), static_cast<const T*>(in), M);
    } else {
      long long blocks = (cols + 7) / 8;
      if (blocks > 148 * 32) blocks = 148 * 32;
      softmax_dim1_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(static_cast<T*>(out), static_cast<const T*>(in), M, cols);
    }
  } else if (N >= 4096 && ((M + 7) / 8) * B * 8 <= 0x7fffffffLL) {      // very long rows: 8-CTA clusters
    softmax_dim2_cluster_kernel<T><<<(unsigned)(((M + 7) / 8) * B * 8), 256, 0, st>>>(static_cast<T*>(out), static_cast<const T*>(in), M, N, B);
  } else if (N >= 256) {                    // long strided rows: slice N across the block
    long long blocks = ((M + 7) / 8) * B;
    if (blocks > 148 * 8) blocks = 148 * 8;
    softmax_dim2_sliced_kernel<T><<<(unsigned)blocks, 1024, 0, st>>>(static_cast<T*>(out), static_cast<const T*>(in), M, N, B);
  } else {
    long long blocks = (M * B + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    softmax_dim2_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(static_cast<T*>(out), static_cast<const T*>(in), M, N, B);
  }
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}
}  // namespace

int softmax_launch(void* out, const void* in, long long M, long long N, long long B, int dim, int dtype, cudaStream_t st) {
  switch (dtype) {
    case FA_F32: return launch<float>(out, in, M, N, B, dim, st);
    case FA_F16: return launch<__half>(out, in, M, N, B, dim, st);
    case FA_BF16: return launch<__nv_bfloat16>(out, in, M, N, B, dim, st);
  }
  set_error("unknown dtype %d", dtype);
  return FA_ERR_INVALID;
}
}  // namespace fa
