// fa_softmax.cu -- standalone safe softmax along Julia dim 1 (columns, contiguous) or dim 2
// (rows, strided) of a column-major (M, N, B) array.  Replaces fused_softmax! / col_softmax! /
// row_softmax! (reference src/fused_softmax.jl:11-39).  HBM-bound: two reads + one write.
#include "fa_common.cuh"

namespace fa {
namespace {

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
      const float v = to_f32<T>(x[i]);
      const float mn = fmaxf(mx, v);
      sum = sum * expf(mx - mn) + expf(v - mn);
      mx = mn;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o), os = __shfl_xor_sync(0xffffffffu, sum, o);
      const float mn = fmaxf(mx, om);
      const float a = (mx == -INFINITY) ? 0.f : expf(mx - mn), b = (om == -INFINITY) ? 0.f : expf(om - mn);
      sum = sum * a + os * b;
      mx = mn;
    }
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
      const float v = to_f32<T>(x[n * M]);
      const float mn = fmaxf(mx, v);
      sum = sum * expf(mx - mn) + expf(v - mn);
      mx = mn;
    }
    const float inv = 1.f / sum;
    for (long long n = 0; n < N; ++n) y[n * M] = from_f32<T>(expf(to_f32<T>(x[n * M]) - mx) * inv);
  }
}

template <typename T>
int launch(void* out, const void* in, long long M, long long N, long long B, int dim, cudaStream_t st) {
  if (dim == 1) {
    const long long cols = N * B;
    long long blocks = (cols + 7) / 8;
    if (blocks > 148 * 32) blocks = 148 * 32;
    softmax_dim1_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(static_cast<T*>(out), static_cast<const T*>(in), M, cols);
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
