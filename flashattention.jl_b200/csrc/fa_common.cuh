// fa_common.cuh -- shared host/device definitions for libfa_sm100a.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>
#include <math.h>

#include "../../include/fa_sm100a.h"

namespace fa {

// ---------------------------------------------------------------------------------------
// Geometry of one attention call.  "Slots" are rows/columns of one independent attention
// problem:  dense/circulant -> one problem per batch element, slot == token;
//           windowed        -> one problem per (window, batch), slot == position inside the
//                              window (kappa, first-dim-fastest), gathered on the fly from the
//                              (spatial..., d, B) tensor with zero fill outside (the fused
//                              `window`/`unwindow` of reference src/utils.jl:36-54).
// ---------------------------------------------------------------------------------------
enum Mode : int { MODE_DENSE = 0, MODE_CIRCULANT = 1, MODE_WINDOWED = 2 };

struct Geo {
  int mode;
  int d, dv;
  int nd;            // spatial dims (windowed)
  int s[3];          // spatial extents (windowed), s[0] fastest
  int o[3];          // windows per dim (windowed)
  int W;             // window size (circulant: band width; windowed: per-dim window)
  int stride;        // windowed
  int padv[3];       // windowed: zero padding in front of each spatial dim (equal unless the volume is a slab)
  int WD;            // W^nd slots per window (windowed)
  int p;             // (W-1)/2 (circulant)
  int overlap;       // windowed: 1 if a position can be covered by >1 window (fold needs +=)
  long long N;       // tokens per batch element
  long long L;       // windows per batch element (windowed)
  long long B;       // batch elements
  float tau;         // 1/sqrt(d)
};

__host__ __device__ inline long long geo_problems(const Geo& g) {
  return g.mode == MODE_WINDOWED ? g.L * g.B : g.B;
}
__host__ __device__ inline long long geo_slots(const Geo& g) {
  return g.mode == MODE_WINDOWED ? (long long)g.WD : g.N;
}

// token (0-based linear spatial index) read by slot `slot` of window `win`, or -1 for padding
__host__ __device__ inline long long window_slot_token(const Geo& g, long long win, int slot) {
  long long tok = 0, mult = 1;
  for (int k = 0; k < g.nd; ++k) {
    const int wk = (int)(win % g.o[k]);  win /= g.o[k];
    const int kk = slot % g.W;           slot /= g.W;
    const int pos = wk * g.stride - g.padv[k] + kk;
    if (pos < 0 || pos >= g.s[k]) return -1;
    tok += pos * mult;
    mult *= g.s[k];
  }
  return tok;
}

// number of windows covering spatial token `tok` (the divisor of src/windowed.jl:16-17)
__host__ __device__ inline int window_count_at(const Geo& g, long long tok) {
  int cnt = 1;
  for (int k = 0; k < g.nd; ++k) {
    const int pos = (int)(tok % g.s[k]);  tok /= g.s[k];
    const int a = pos + g.padv[k];                      // 0 <= a - w*stride < W
    int wmax = a / g.stride;
    if (wmax > g.o[k] - 1) wmax = g.o[k] - 1;
    int lo = a - g.W + 1;
    int wmin = lo <= 0 ? 0 : (lo + g.stride - 1) / g.stride;
    const int c = wmax - wmin + 1;
    cnt *= (c > 0 ? c : 0);
  }
  return cnt;
}

__host__ __device__ inline long long pmod(long long a, long long n) {
  long long r = a % n;
  return r < 0 ? r + n : r;
}

// ---------------------------------------------------------------------------------------
// dtype helpers
// ---------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T x);
template <> __device__ __forceinline__ float to_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f32<__half>(__half x) { return __half2float(x); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }

template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ __half from_f32<__half>(float x) { return __float2half_rn(x); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

inline size_t dtype_size(int dtype) { return dtype == FA_F32 ? 4 : 2; }

// ---------------------------------------------------------------------------------------
// error plumbing (thread-local message, status codes across the C ABI)
// ---------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void set_path(const char* name);
int cuda_fail(cudaError_t e, const char* what);

#define FA_CUDA_TRY(expr)                                   \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) return fa::cuda_fail(_e, #expr); \
  } while (0)

// ---------------------------------------------------------------------------------------
// launchers implemented in the .cu files
// ---------------------------------------------------------------------------------------
struct FwdArgs {
  const void *q, *k, *v;
  void* o;          // dense/circulant: output; windowed non-overlap: y; windowed overlap: unused
  float* acc;       // windowed overlap: fp32 fold accumulator (N*dv*B), zero-initialised
  float *l, *m;
  int o_f32;        // tcgen05 dense/circulant forward only: 1 = `o` is float32 whatever the input dtype; 2 (pair kernel) =
                    // `o`, `l`, `m` hold a running (O, l, m) and this call's result is MERGED into them (ring pass)
                    // (block partials of the ring pass must not be rounded to 16 bits before merging)
  // tcgen05 dense forward only: q, k, v are VIEWS with these byte strides between channels / batch elements instead of
  // the dense N * 2 and N * d * 2 (0 = dense).  Used for 1-D windows larger than a tile: "batch element" w = window w,
  // stride_b = window stride, so that overlapping windows are read in place (no unfold copy).  Multiples of 16.
  long long in_stride_c, in_stride_b;
};
struct BwdArgs {
  const void *q, *k, *v, *o, *d_o;
  const float *l, *m;
  void *dq, *dk, *dv;      // dense/circulant/windowed non-overlap: final outputs
  float *aq, *ak, *av;     // windowed overlap: fp32 fold accumulators, zero-initialised
  float* delta;            // D_i = rowsum(dO o O) per slot, (slots * problems) floats
};

int simt_fwd(const Geo& g, const FwdArgs& a, int dtype, cudaStream_t st);
int simt_bwd(const Geo& g, const BwdArgs& a, int dtype, cudaStream_t st);
int fold_finalize(const Geo& g, const float* acc, void* y, int channels, int dtype, int divide,
                  cudaStream_t st);
int slab_divide(const Geo& g, const float* acc, void* y, int channels, int dtype, long long plane_lo, long long slab_tokens, cudaStream_t st);
int fill_uncovered_nan(const Geo& g, void* y, int channels, int dtype, cudaStream_t st);
int window_gather(const Geo& g, const void* x, void* xw, int dtype, cudaStream_t st);
// divide != 0: y = fold(xw) ./ count (windowed_fa, src/windowed.jl:16-19); channels = g.d
int window_scatter(const Geo& g, const void* xw, void* x, int dtype, cudaStream_t st, int divide = 0);
int cast_launch(const void* in, void* out, long long n, int from, int to, cudaStream_t st);
int softmax_launch(void* out, const void* in, long long M, long long N, long long B, int dim,
                   int dtype, cudaStream_t st);

// tensor-core (tcgen05) forward: dense + circulant, 16-bit dtypes, d == dv in {64,128}
bool tc_fwd_supported(const Geo& g, int dtype);
int tc_fwd(const Geo& g, const FwdArgs& a, int dtype, cudaStream_t st);
// compact three-CTAs-per-SM forward for short key loops (circulant, d = dv = 64); dispatched from tc_fwd
int tc_band_fwd(const Geo& g, const FwdArgs& a, int dtype, cudaStream_t st);
// the same kernel for the 2-D periodic neighbourhood (fa_circulant2d_fwd): 16-bit, d = dv in {64, 128}, X % 64 == 0
bool tc_band2d_supported(long long X, long long Y, long long d, long long dv, long long B, long long W, int dtype);
int tc_band2d_fwd(const void* q, const void* k, const void* v, void* o, float* l, float* m,
                  long long X, long long Y, long long d, long long B, long long W, int dtype, cudaStream_t st);
// tensor-core backward (two deterministic kernels: key-owner dK/dV, query-owner dQ); same coverage
bool tc_bwd_supported(const Geo& g, int dtype);
size_t tc_bwd_workspace_bytes(const Geo& g, int dtype, int flags);
// out_f32: dq/dk/dv are float32 buffers whatever the input dtype (partials of the ring backward)
int tc_bwd(const Geo& g, const BwdArgs& a, int dtype, int flags, void* workspace, cudaStream_t st, int out_f32 = 0);
int merge_partials(float* oa, float* la, float* ma, const void* ob, const float* lb, const float* mb, void* out,
                   long long N, int dv, long long B, int dtype, int blk_f32, int first, cudaStream_t st);

// tensor-core windowed attention (one 128-row tile = floor(128 / W^D) windows): 16-bit dtypes,
// d == dv in {64,128}, W^D <= 128
bool tc_win_supported(const Geo& g, int dtype);
int tc_win_fwd(const Geo& g, const FwdArgs& a, int dtype, cudaStream_t st);
bool tc_win_bwd_supported(const Geo& g, int dtype);
int tc_win_bwd(const Geo& g, const BwdArgs& a, int dtype, cudaStream_t st);

}  // namespace fa
