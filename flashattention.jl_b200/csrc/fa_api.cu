// fa_api.cu -- the extern "C" surface of libfa_sm100a.so (see include/fa_sm100a.h).
// Argument validation mirrors the reference's Julia signatures (same shapes, same defaults,
// errors instead of MethodError/assertion); dispatch picks the tcgen05 path for 16-bit dense /
// circulant forward and the exact-fp32 SIMT path otherwise.  There is no CPU fallback.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <sched.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <vector>

#include "fa_common.cuh"

namespace fa {

static thread_local char g_err[512] = "";
static thread_local char g_path[64] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void set_path(const char* name) { snprintf(g_path, sizeof(g_path), "%s", name); }
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return FA_ERR_CUDA;
}

namespace {

bool valid_dtype(int dt) { return dt == FA_F32 || dt == FA_F16 || dt == FA_BF16; }

int check_common(int64_t N, int64_t d, int64_t dv, int64_t B, int dtype) {
  if (!valid_dtype(dtype)) { set_error("dtype must be FA_F32, FA_F16 or FA_BF16 (got %d)", dtype); return FA_ERR_INVALID; }
  if (N <= 0 || d <= 0 || dv <= 0 || B <= 0) { set_error("N, d, dv, B must be positive (got %lld, %lld, %lld, %lld)", (long long)N, (long long)d, (long long)dv, (long long)B); return FA_ERR_INVALID; }
  if (N > 0x7fffffffLL || d > 4096 || dv > 4096) { set_error("dimension out of range"); return FA_ERR_INVALID; }
  return FA_OK;
}

int need_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_error("no CUDA device available (libfa_sm100a has no CPU fallback)");
    return FA_ERR_CUDA;
  }
  return FA_OK;
}

Geo dense_geo(int64_t N, int64_t d, int64_t dv, int64_t B) {
  Geo g;
  memset(&g, 0, sizeof(g));
  g.mode = MODE_DENSE; g.d = (int)d; g.dv = (int)dv; g.N = N; g.B = B;
  g.tau = 1.0f / sqrtf((float)d);                     // reference src/dense.jl:43
  return g;
}

// slab_pad_lo >= 0 makes the geometry a SLAB of a larger volume along the slowest spatial dim: `dims` are the
// slab's own extents, the first window of that dim starts slab_pad_lo planes in front of the slab and
// slab_nwin windows are taken (instead of the NNlib output size); the other dims keep `pad`.
int windowed_geo(Geo& g, int ndim, const int64_t* dims, int64_t d, int64_t dv, int64_t B, int64_t W,
                 int64_t stride, int64_t pad, int64_t slab_pad_lo = -1, int64_t slab_nwin = 0) {
  if (ndim < 1 || ndim > 3 || !dims) { set_error("ndim must be 1, 2 or 3"); return FA_ERR_INVALID; }
  if (W <= 0 || stride <= 0 || pad < 0) { set_error("need W > 0, stride > 0, pad >= 0"); return FA_ERR_INVALID; }
  memset(&g, 0, sizeof(g));
  g.mode = MODE_WINDOWED; g.d = (int)d; g.dv = (int)dv; g.B = B; g.nd = ndim;
  g.W = (int)W; g.stride = (int)stride;
  for (int k = 0; k < 3; ++k) g.padv[k] = (int)pad;
  g.tau = 1.0f / sqrtf((float)d);
  long long N = 1, L = 1, WD = 1;
  for (int k = 0; k < 3; ++k) { g.s[k] = 1; g.o[k] = 1; }
  for (int k = 0; k < ndim; ++k) {
    if (dims[k] <= 0 || dims[k] > 0x7fffffffLL) { set_error("bad spatial extent"); return FA_ERR_INVALID; }
    long long o = (dims[k] + 2 * pad - W) / stride + 1;            // NNlib output size
    if (slab_pad_lo >= 0 && k == ndim - 1) {
      if (slab_nwin <= 0 || slab_pad_lo >= W || (slab_nwin - 1) * stride - slab_pad_lo >= dims[k]) {
        set_error("slab: need nwin > 0, pad_lo < W and the last window starting inside the slab"); return FA_ERR_INVALID;
      }
      g.padv[k] = (int)slab_pad_lo; o = slab_nwin;
    } else if (dims[k] + 2 * pad < W || o <= 0) { set_error("window (%lld) larger than padded extent (%lld + 2*%lld)", (long long)W, (long long)dims[k], (long long)pad); return FA_ERR_INVALID; }
    g.s[k] = (int)dims[k]; g.o[k] = (int)o;
    N *= dims[k]; L *= o; WD *= W;
    if (WD > 0x7fffffffLL || N > 0x7fffffffLL) { set_error("window or volume too large"); return FA_ERR_INVALID; }
  }
  g.N = N; g.L = L; g.WD = (int)WD;
  g.overlap = stride < W ? 1 : 0;
  return FA_OK;
}

// does every spatial position belong to at least one window?
bool full_cover(const Geo& g) {
  for (int k = 0; k < g.nd; ++k)
    for (int pos = 0; pos < g.s[k]; ++pos) {
      Geo g1 = g; g1.nd = 1; g1.s[0] = g.s[k]; g1.o[0] = g.o[k]; g1.padv[0] = g.padv[k];
      if (window_count_at(g1, pos) == 0) return false;
    }
  return true;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// FA_FLAG_OUT_F32: 16-bit inputs, float32 outputs (the accumulators are stored unrounded).  Only the tcgen05
// kernels implement it; for Float32 inputs the flag is a no-op.
bool want_f32_out(int dtype, int flags) { return (flags & FA_FLAG_OUT_F32) && dtype != FA_F32; }
int no_f32_out() { set_error("FA_FLAG_OUT_F32 needs a shape the tcgen05 kernels cover (16-bit, d == dv in {64,128})"); return FA_ERR_UNSUPPORTED; }

}  // namespace
}  // namespace fa

using namespace fa;

extern "C" {

int fa_version(void) { return FA_VERSION_MAJOR * 100 + FA_VERSION_MINOR; }
const char* fa_last_error_string(void) { return g_err; }
const char* fa_last_path(void) { return g_path; }
int fa_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// ------------------------------------------------------------------------------ index sets
int fa_circulant_index(int64_t N, int64_t W, int64_t* keys) {
  if (N <= 0 || W <= 0 || W > N || !keys) { set_error("need 0 < W <= N and a keys buffer"); return FA_ERR_INVALID; }
  const int64_t p = (W - 1) / 2;                                       // src/utils.jl:8
  for (int64_t j = 1; j <= N; ++j)
    for (int64_t w = 1; w <= W; ++w) {
      int64_t m = w;                                                   // :10
      if (j <= p) m = ((m - 1 - (j - p - 1)) % W + W) % W + 1;         // :11-12
      else if (j > N - p) m = ((m - 1 - (p - N + j)) % W + W) % W + 1; // :13
      const int64_t i = (((m - 1) + (j - 1) - p) % N + N) % N + 1;     // :15
      keys[(j - 1) * W + (w - 1)] = i - 1;
    }
  return FA_OK;
}

int fa_window_index(int ndim, const int64_t* dims, int64_t W, int64_t stride, int64_t pad,
                    int64_t* n_windows, int64_t* idx) {
  Geo g;
  int rc = windowed_geo(g, ndim, dims, 1, 1, 1, W, stride, pad);
  if (rc) return rc;
  if (n_windows) for (int k = 0; k < ndim; ++k) n_windows[k] = g.o[k];
  if (idx)
    for (long long w = 0; w < g.L; ++w)
      for (int s = 0; s < g.WD; ++s) idx[w * g.WD + s] = window_slot_token(g, w, s);
  return FA_OK;
}

int fa_window_count(int ndim, const int64_t* dims, int64_t W, int64_t stride, int64_t pad, int64_t* cnt) {
  Geo g;
  int rc = windowed_geo(g, ndim, dims, 1, 1, 1, W, stride, pad);
  if (rc) return rc;
  if (!cnt) { set_error("cnt is NULL"); return FA_ERR_INVALID; }
  for (long long t = 0; t < g.N; ++t) cnt[t] = window_count_at(g, t);
  return FA_OK;
}

// ------------------------------------------------------------------------------ dense
int fa_dense_fwd(const void* q, const void* k, const void* v, void* o, float* l, float* m,
                 int64_t N, int64_t d, int64_t dv, int64_t B, int dtype, int flags, void* stream) {
  int rc = check_common(N, d, dv, B, dtype);
  if (rc) return rc;
  if (!q || !k || !v || !o || !l || !m) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  if ((rc = need_device())) return rc;
  const Geo g = dense_geo(N, d, dv, B);
  FwdArgs a{q, k, v, o, nullptr, l, m, want_f32_out(dtype, flags) ? 1 : 0};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!(flags & FA_FLAG_FORCE_SIMT) && tc_fwd_supported(g, dtype)) { set_path("tc"); return tc_fwd(g, a, dtype, st); }
  if (a.o_f32) return no_f32_out();
  set_path("simt");
  return simt_fwd(g, a, dtype, st);
}

size_t fa_workspace_bytes_dense_bwd(int64_t N, int64_t d, int64_t dv, int64_t B, int dtype, int flags) {
  if (N <= 0 || B <= 0 || d <= 0 || dv <= 0) return 0;
  const size_t simt = align256((size_t)N * B * sizeof(float));          // delta
  const Geo g = dense_geo(N, d, dv, B);
  if ((flags & FA_FLAG_FORCE_SIMT) || !tc_bwd_supported(g, dtype)) return simt;
  const size_t tc = tc_bwd_workspace_bytes(g, dtype, flags);            // nlse, ndelta (+ fp16 re-encodings)
  return simt > tc ? simt : tc;
}

int fa_dense_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                 const float* l, const float* m, void* dq, void* dk, void* dv_out,
                 int64_t N, int64_t d, int64_t dv, int64_t B, int dtype, int flags,
                 void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(N, d, dv, B, dtype);
  if (rc) return rc;
  if (!q || !k || !v || !o || !d_o || !l || !m || !dq || !dk || !dv_out) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  if (!workspace || workspace_bytes < fa_workspace_bytes_dense_bwd(N, d, dv, B, dtype, flags)) { set_error("workspace too small"); return FA_ERR_WORKSPACE; }
  if ((rc = need_device())) return rc;
  const Geo g = dense_geo(N, d, dv, B);
  BwdArgs a{q, k, v, o, d_o, l, m, dq, dk, dv_out, nullptr, nullptr, nullptr, static_cast<float*>(workspace)};
  if (!(flags & FA_FLAG_FORCE_SIMT) && tc_bwd_supported(g, dtype)) { set_path("tc"); return tc_bwd(g, a, dtype, flags, workspace, static_cast<cudaStream_t>(stream), want_f32_out(dtype, flags) ? 1 : 0); }
  if (want_f32_out(dtype, flags)) return no_f32_out();
  set_path("simt");
  return simt_bwd(g, a, dtype, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------ circulant
static int circ_geo(Geo& g, int64_t N, int64_t d, int64_t dv, int64_t B, int64_t W) {
  if (W <= 0 || W > N) { set_error("circulant window must satisfy 0 < W <= N (got W=%lld, N=%lld)", (long long)W, (long long)N); return FA_ERR_INVALID; }
  g = dense_geo(N, d, dv, B);
  g.mode = MODE_CIRCULANT; g.W = (int)W; g.p = (int)((W - 1) / 2);     // src/utils.jl:8
  return FA_OK;
}

int fa_circulant_fwd(const void* q, const void* k, const void* v, void* o, float* l, float* m,
                     int64_t N, int64_t d, int64_t dv, int64_t B, int64_t W, int dtype, int flags, void* stream) {
  int rc = check_common(N, d, dv, B, dtype);
  if (rc) return rc;
  if (!q || !k || !v || !o || !l || !m) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  Geo g;
  if ((rc = circ_geo(g, N, d, dv, B, W))) return rc;
  if ((rc = need_device())) return rc;
  FwdArgs a{q, k, v, o, nullptr, l, m, want_f32_out(dtype, flags) ? 1 : 0};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!(flags & FA_FLAG_FORCE_SIMT) && tc_fwd_supported(g, dtype)) { set_path("tc"); return tc_fwd(g, a, dtype, st); }
  if (a.o_f32) return no_f32_out();
  set_path("simt");
  return simt_fwd(g, a, dtype, st);
}

size_t fa_workspace_bytes_circulant_bwd(int64_t N, int64_t d, int64_t dv, int64_t B, int64_t W, int dtype, int flags) {
  (void)W;
  return fa_workspace_bytes_dense_bwd(N, d, dv, B, dtype, flags);
}

int fa_circulant_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                     const float* l, const float* m, void* dq, void* dk, void* dv_out,
                     int64_t N, int64_t d, int64_t dv, int64_t B, int64_t W, int dtype, int flags,
                     void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(N, d, dv, B, dtype);
  if (rc) return rc;
  if (!q || !k || !v || !o || !d_o || !l || !m || !dq || !dk || !dv_out) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  Geo g;
  if ((rc = circ_geo(g, N, d, dv, B, W))) return rc;
  if (!workspace || workspace_bytes < fa_workspace_bytes_circulant_bwd(N, d, dv, B, W, dtype, flags)) { set_error("workspace too small"); return FA_ERR_WORKSPACE; }
  if ((rc = need_device())) return rc;
  BwdArgs a{q, k, v, o, d_o, l, m, dq, dk, dv_out, nullptr, nullptr, nullptr, static_cast<float*>(workspace)};
  if (!(flags & FA_FLAG_FORCE_SIMT) && tc_bwd_supported(g, dtype)) { set_path("tc"); return tc_bwd(g, a, dtype, flags, workspace, static_cast<cudaStream_t>(stream), want_f32_out(dtype, flags) ? 1 : 0); }
  if (want_f32_out(dtype, flags)) return no_f32_out();
  set_path("simt");
  return simt_bwd(g, a, dtype, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------ windowed
// 1-D windows LARGER than one 128-row tile (W > 128; the reference benchmarks W up to 512, logs/wind_t16.txt:7-8): a
// window is a dense attention problem of W tokens, and with pad = 0 window w is the token range [w stride, w stride + W)
// of the array itself.  So the dense tcgen05 kernels run on a VIEW (W, d, L) of q, k, v -- "batch element" w = window w,
// batch stride = the window stride, read in place through the tensor map, no unfold copy -- into a (W, dv, L, B)
// workspace (the `yw` of src/windowed.jl:8-11), and the deterministic gather-form fold finishes y = fold(yw) ./ count.
static bool bigwin_supported(const Geo& g, int dtype, int flags) {
  if (flags & (FA_FLAG_FORCE_SIMT | FA_FLAG_OUT_F32)) return false;
  if (dtype != FA_BF16 && dtype != FA_F16) return false;
  if (g.mode != MODE_WINDOWED || g.nd != 1 || g.WD <= 128 || g.padv[0] != 0) return false;
  if (g.d != g.dv || (g.d != 32 && g.d != 64 && g.d != 128)) return false;
  if (g.stride % 8 != 0 || g.N % 8 != 0 || g.W % 8 != 0 || g.L > 65535) return false;     // 16-byte TMA strides; gridDim.y
  return true;
}
// f32out (FA_FLAG_OUT_F32): the fold accumulators are used even without overlap and finalised into float32 outputs
static size_t windowed_fwd_ws(const Geo& g, bool f32out = false, int dtype = FA_F32, int flags = FA_FLAG_FORCE_SIMT) {
  if (bigwin_supported(g, dtype, flags)) return align256((size_t)g.WD * g.dv * g.L * g.B * dtype_size(dtype));
  return (g.overlap || f32out) ? align256((size_t)g.N * g.dv * g.B * sizeof(float)) : 256;
}
static size_t windowed_bwd_ws(const Geo& g, bool f32out = false) {
  size_t bytes = align256((size_t)g.WD * g.L * g.B * sizeof(float));                 // delta per window slot
  if (g.overlap || f32out) bytes += 2 * align256((size_t)g.N * g.d * g.B * sizeof(float)) + align256((size_t)g.N * g.dv * g.B * sizeof(float));
  return bytes;
}

size_t fa_workspace_bytes_windowed_fwd(int ndim, const int64_t* dims, int64_t d, int64_t dv, int64_t B,
                                       int64_t W, int64_t stride, int64_t pad, int dtype, int flags) {
  Geo g;
  if (windowed_geo(g, ndim, dims, d, dv, B, W, stride, pad)) return 0;
  return windowed_fwd_ws(g, want_f32_out(dtype, flags), dtype, flags);
}

static int windowed_fwd_impl(const Geo& g, const void* q, const void* k, const void* v, void* y, float* l, float* m,
                             int dtype, int flags, void* workspace, size_t workspace_bytes, void* stream) {
  const int64_t d = g.d, dv = g.dv, B = g.B;
  int rc;
  if ((rc = check_common(g.N, d, dv, B, dtype))) return rc;
  if (!q || !k || !v || !y || !l || !m) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  const bool f32out = want_f32_out(dtype, flags);
  const bool bigwin = bigwin_supported(g, dtype, flags) && !(workspace == nullptr && workspace_bytes == 0);   // slab calls carry no workspace
  if ((g.overlap || f32out || bigwin) && (!workspace || workspace_bytes < windowed_fwd_ws(g, f32out, dtype, bigwin ? flags : FA_FLAG_FORCE_SIMT))) {
    set_error("workspace too small"); return FA_ERR_WORKSPACE;
  }
  if ((rc = need_device())) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bigwin) {
    set_path("tc");
    Geo gd = dense_geo(g.WD, d, dv, g.L);                     // one window = one dense problem; tau = 1/sqrt(d) as src/windowed.jl:8-11
    const size_t esz = dtype_size(dtype);
    for (int64_t b = 0; b < B; ++b) {
      const char* qb = static_cast<const char*>(q) + (size_t)b * g.N * d * esz;
      const char* kb = static_cast<const char*>(k) + (size_t)b * g.N * d * esz;
      const char* vb = static_cast<const char*>(v) + (size_t)b * g.N * dv * esz;
      char* ow = static_cast<char*>(workspace) + (size_t)b * g.WD * dv * g.L * esz;
      FwdArgs aw{qb, kb, vb, ow, nullptr, l + (size_t)b * g.WD * g.L, m + (size_t)b * g.WD * g.L, 0,
                 (long long)g.N * (long long)esz, (long long)g.stride * (long long)esz};
      if ((rc = tc_fwd(gd, aw, dtype, st))) return rc;
    }
    Geo gf = g; gf.d = (int)dv;                               // fold over dv channels
    return window_scatter(gf, workspace, y, dtype, st, /*divide=*/1);      // y = fold(yw) ./ count; uncovered -> 0/0 = NaN
  }
  FwdArgs a{q, k, v, y, nullptr, l, m};
  const bool tc = !(flags & FA_FLAG_FORCE_SIMT) && tc_win_supported(g, dtype);
  if (f32out && !tc) return no_f32_out();
  set_path(tc ? "tc" : "simt");
  if (g.overlap || f32out) {
    a.acc = static_cast<float*>(workspace);
    FA_CUDA_TRY(cudaMemsetAsync(a.acc, 0, (size_t)g.N * dv * B * sizeof(float), st));
    if ((rc = tc ? tc_win_fwd(g, a, dtype, st) : simt_fwd(g, a, dtype, st))) return rc;
    return fold_finalize(g, a.acc, y, (int)dv, f32out ? FA_F32 : dtype, /*divide=*/1, st);     // src/windowed.jl:19
  }
  if ((rc = tc ? tc_win_fwd(g, a, dtype, st) : simt_fwd(g, a, dtype, st))) return rc;
  if (!full_cover(g)) return fill_uncovered_nan(g, y, (int)dv, dtype, st);   // 0/0 = NaN
  return FA_OK;
}

int fa_windowed_fwd(const void* q, const void* k, const void* v, void* y, float* l, float* m,
                    int ndim, const int64_t* dims, int64_t d, int64_t dv, int64_t B,
                    int64_t W, int64_t stride, int64_t pad, int dtype, int flags,
                    void* workspace, size_t workspace_bytes, void* stream) {
  Geo g;
  int rc = windowed_geo(g, ndim, dims, d, dv, B, W, stride, pad);
  if (rc) return rc;
  return windowed_fwd_impl(g, q, k, v, y, l, m, dtype, flags, workspace, workspace_bytes, stream);
}

size_t fa_workspace_bytes_windowed_bwd(int ndim, const int64_t* dims, int64_t d, int64_t dv, int64_t B,
                                       int64_t W, int64_t stride, int64_t pad, int dtype, int flags) {
  Geo g;
  if (windowed_geo(g, ndim, dims, d, dv, B, W, stride, pad)) return 0;
  return windowed_bwd_ws(g, want_f32_out(dtype, flags));
}

static int windowed_bwd_impl(const Geo& g, const void* q, const void* k, const void* v, const void* d_y,
                             const float* l, const float* m, void* dq, void* dk, void* dv_out,
                             int dtype, int flags, void* workspace, size_t workspace_bytes, void* stream) {
  const int64_t d = g.d, dv = g.dv, B = g.B;
  int rc;
  if ((rc = check_common(g.N, d, dv, B, dtype))) return rc;
  if (!q || !k || !v || !d_y || !l || !m || !dq || !dk || !dv_out) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  const bool f32out = want_f32_out(dtype, flags);
  if (!workspace || workspace_bytes < windowed_bwd_ws(g, f32out)) {
    set_error("workspace too small"); return FA_ERR_WORKSPACE;
  }
  if ((rc = need_device())) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  BwdArgs a{q, k, v, nullptr, d_y, l, m, dq, dk, dv_out, nullptr, nullptr, nullptr, reinterpret_cast<float*>(ws)};
  ws += align256((size_t)g.WD * g.L * B * sizeof(float));
  const size_t esz = dtype_size(dtype);
  const bool tc = !(flags & FA_FLAG_FORCE_SIMT) && tc_win_bwd_supported(g, dtype);
  if (f32out && !tc) return no_f32_out();
  set_path(tc ? "tc" : "simt");
  if (g.overlap || f32out) {
    const int odt = f32out ? FA_F32 : dtype;
    const size_t nq = (size_t)g.N * d * B, nv = (size_t)g.N * dv * B;
    a.aq = reinterpret_cast<float*>(ws); ws += align256(nq * 4);
    a.ak = reinterpret_cast<float*>(ws); ws += align256(nq * 4);
    a.av = reinterpret_cast<float*>(ws);
    FA_CUDA_TRY(cudaMemsetAsync(a.aq, 0, nq * 4, st));
    FA_CUDA_TRY(cudaMemsetAsync(a.ak, 0, nq * 4, st));
    FA_CUDA_TRY(cudaMemsetAsync(a.av, 0, nv * 4, st));
    if ((rc = tc ? tc_win_bwd(g, a, dtype, st) : simt_bwd(g, a, dtype, st))) return rc;
    if ((rc = fold_finalize(g, a.aq, dq, (int)d, odt, 0, st))) return rc;   // adjoint of unfold: fold, no division
    if ((rc = fold_finalize(g, a.ak, dk, (int)d, odt, 0, st))) return rc;
    return fold_finalize(g, a.av, dv_out, (int)dv, odt, 0, st);
  }
  if (!full_cover(g)) {   // uncovered positions receive no gradient
    FA_CUDA_TRY(cudaMemsetAsync(dq, 0, (size_t)g.N * d * B * esz, st));
    FA_CUDA_TRY(cudaMemsetAsync(dk, 0, (size_t)g.N * d * B * esz, st));
    FA_CUDA_TRY(cudaMemsetAsync(dv_out, 0, (size_t)g.N * dv * B * esz, st));
  }
  return tc ? tc_win_bwd(g, a, dtype, st) : simt_bwd(g, a, dtype, st);
}

int fa_windowed_bwd(const void* q, const void* k, const void* v, const void* d_y,
                    const float* l, const float* m, void* dq, void* dk, void* dv_out,
                    int ndim, const int64_t* dims, int64_t d, int64_t dv, int64_t B,
                    int64_t W, int64_t stride, int64_t pad, int dtype, int flags,
                    void* workspace, size_t workspace_bytes, void* stream) {
  Geo g;
  int rc = windowed_geo(g, ndim, dims, d, dv, B, W, stride, pad);
  if (rc) return rc;
  return windowed_bwd_impl(g, q, k, v, d_y, l, m, dq, dk, dv_out, dtype, flags, workspace, workspace_bytes, stream);
}

// ------------------------------------------------------------------------------ windowed, one volume over several GPUs
// Non-overlapping windows (stride >= W) are independent, so a single volume splits into SLABS along its
// slowest spatial dim on window boundaries with no exchange at all: rank r takes the window
// planes [win_lo, win_hi) of that dim and holds exactly the token planes [plane_lo, plane_hi) they read.
// plan[5] = {plane_lo, plane_hi, win_lo, win_hi, pad_lo}; pad_lo = zero planes in front of the slab (the
// volume's own padding, non-zero only for the rank that owns the first window).
int fa_windowed_slab_plan(int ndim, const int64_t* dims, int64_t W, int64_t stride, int64_t pad,
                          int rank, int nranks, int64_t* plan) {
  Geo g;
  int rc = windowed_geo(g, ndim, dims, 1, 1, 1, W, stride, pad);
  if (rc) return rc;
  if (!plan || nranks <= 0 || rank < 0 || rank >= nranks) { set_error("bad rank / nranks / plan"); return FA_ERR_INVALID; }
  if (stride < W) { set_error("slab split needs non-overlapping windows (stride >= W); overlapping windows need a halo reduce"); return FA_ERR_UNSUPPORTED; }
  if (pad >= W) { set_error("slab split needs pad < W"); return FA_ERR_UNSUPPORTED; }
  const int ks = ndim - 1;
  const int64_t S = dims[ks], nw = g.o[ks];
  int64_t lo = 0, cnt = 0;
  if ((rc = fa_shard_batch(nw, nranks, rank, &lo, &cnt))) return rc;
  const int64_t hi = lo + cnt;
  plan[2] = lo; plan[3] = hi;
  if (cnt <= 0) { plan[0] = plan[1] = 0; plan[4] = 0; return FA_OK; }          // more ranks than window planes
  // slab boundary in front of window w = its first plane (clipped): the slabs tile the volume, and planes in
  // the gaps between windows (stride > W) or behind the last window stay with the slab in front of them --
  // uncovered there (y = NaN, zero gradient) exactly as in the one-GPU call
  auto bnd = [&](int64_t w) { const int64_t x = w * stride - pad; return x < 0 ? (int64_t)0 : (x > S ? S : x); };
  plan[0] = lo == 0 ? 0 : bnd(lo);
  plan[1] = hi == nw ? S : bnd(hi);
  plan[4] = plan[0] - (lo * stride - pad);
  return FA_OK;
}

static int slab_geo(Geo& g, int ndim, const int64_t* slab_dims, int64_t d, int64_t dv, int64_t B, int64_t W,
                    int64_t stride, int64_t pad, int64_t pad_lo, int64_t nwin) {
  if (stride < W) { set_error("slab split needs non-overlapping windows (stride >= W)"); return FA_ERR_UNSUPPORTED; }
  if (pad_lo < 0) { set_error("slab: pad_lo must be >= 0"); return FA_ERR_INVALID; }
  return windowed_geo(g, ndim, slab_dims, d, dv, B, W, stride, pad, pad_lo, nwin);
}

size_t fa_workspace_bytes_windowed_slab_bwd(int ndim, const int64_t* slab_dims, int64_t d, int64_t dv, int64_t B,
                                            int64_t W, int64_t stride, int64_t pad, int64_t pad_lo, int64_t nwin) {
  Geo g;
  if (slab_geo(g, ndim, slab_dims, d, dv, B, W, stride, pad, pad_lo, nwin)) return 0;
  return windowed_bwd_ws(g);
}

int fa_windowed_slab_fwd(const void* q, const void* k, const void* v, void* y, float* l, float* m,
                         int ndim, const int64_t* slab_dims, int64_t d, int64_t dv, int64_t B,
                         int64_t W, int64_t stride, int64_t pad, int64_t pad_lo, int64_t nwin,
                         int dtype, int flags, void* stream) {
  Geo g;
  int rc = slab_geo(g, ndim, slab_dims, d, dv, B, W, stride, pad, pad_lo, nwin);
  if (rc) return rc;
  return windowed_fwd_impl(g, q, k, v, y, l, m, dtype, flags, nullptr, 0, stream);
}

int fa_windowed_slab_bwd(const void* q, const void* k, const void* v, const void* d_y,
                         const float* l, const float* m, void* dq, void* dk, void* dv_out,
                         int ndim, const int64_t* slab_dims, int64_t d, int64_t dv, int64_t B,
                         int64_t W, int64_t stride, int64_t pad, int64_t pad_lo, int64_t nwin,
                         int dtype, int flags, void* workspace, size_t workspace_bytes, void* stream) {
  Geo g;
  int rc = slab_geo(g, ndim, slab_dims, d, dv, B, W, stride, pad, pad_lo, nwin);
  if (rc) return rc;
  return windowed_bwd_impl(g, q, k, v, d_y, l, m, dq, dk, dv_out, dtype, flags, workspace, workspace_bytes, stream);
}

// ------------------------------------------------------------------------------ windowed, one volume, OVERLAPPING windows
// "overlapping windows -> halo of W - stride planes of K/V in, partial-y halo reduce out".  Window planes
// are dealt out by their START plane, so a rank's windows read its own planes plus up to W - stride planes of the NEXT
// rank (the halo in); it then holds fold SUMS for all those planes, the halo part of which belongs to the next rank
// (the halo reduce out); the owner divides by the window count of the whole volume.  The two exchanges are the host
// layer's (fa_sm100a/halo.py over torch.distributed, julia/../multigpu.jl over NCCL.jl); this file supplies the plan,
// the un-normalised slab forward and the final division.
int fa_windowed_halo_plan(int ndim, const int64_t* dims, int64_t W, int64_t stride, int64_t pad,
                          int rank, int nranks, int64_t* plan) {
  Geo g;
  int rc = windowed_geo(g, ndim, dims, 1, 1, 1, W, stride, pad);
  if (rc) return rc;
  if (!plan || nranks <= 0 || rank < 0 || rank >= nranks) { set_error("bad rank / nranks / plan"); return FA_ERR_INVALID; }
  if (pad >= W) { set_error("halo split needs pad < W"); return FA_ERR_UNSUPPORTED; }
  const int ks = ndim - 1;
  const int64_t S = dims[ks], nw = g.o[ks];
  int64_t lo = 0, cnt = 0;
  if ((rc = fa_shard_batch(nw, nranks, rank, &lo, &cnt))) return rc;
  const int64_t hi = lo + cnt;
  auto clip = [&](int64_t x) { return x < 0 ? (int64_t)0 : (x > S ? S : x); };
  // plan = {own_lo, own_hi, ext_hi, win_lo, win_hi, pad_lo}: owned planes [own_lo, own_hi) tile the volume; the windows
  // [win_lo, win_hi) read planes [own_lo, ext_hi) (ext_hi - own_hi halo planes of the next rank), the first of them
  // starting pad_lo planes in front of own_lo
  plan[3] = lo; plan[4] = hi;
  if (cnt <= 0) { plan[0] = plan[1] = plan[2] = (lo >= nw ? S : clip(lo * stride - pad)); plan[5] = 0; return FA_OK; }
  plan[0] = lo == 0 ? 0 : clip(lo * stride - pad);
  plan[1] = hi == nw ? S : clip(hi * stride - pad);
  plan[2] = clip((hi - 1) * stride - pad + W);
  if (plan[2] < plan[1]) plan[2] = plan[1];
  plan[5] = plan[0] - (lo * stride - pad);
  return FA_OK;
}

int fa_windowed_slab_fwd_sums(const void* q, const void* k, const void* v, float* acc, float* l, float* m,
                              int ndim, const int64_t* slab_dims, int64_t d, int64_t dv, int64_t B,
                              int64_t W, int64_t stride, int64_t pad, int64_t pad_lo, int64_t nwin,
                              int dtype, int flags, void* stream) {
  Geo g;
  if (pad_lo < 0) { set_error("slab: pad_lo must be >= 0"); return FA_ERR_INVALID; }
  int rc = windowed_geo(g, ndim, slab_dims, d, dv, B, W, stride, pad, pad_lo, nwin);
  if (rc) return rc;
  if ((rc = check_common(g.N, d, dv, B, dtype))) return rc;
  if (!q || !k || !v || !acc || !l || !m) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  if ((rc = need_device())) return rc;
  g.overlap = 1;                                            // always the accumulating (fold-sum) form
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FwdArgs a{q, k, v, nullptr, acc, l, m};
  const bool tc = !(flags & FA_FLAG_FORCE_SIMT) && tc_win_supported(g, dtype);
  set_path(tc ? "tc" : "simt");
  FA_CUDA_TRY(cudaMemsetAsync(acc, 0, (size_t)g.N * dv * B * sizeof(float), st));
  return tc ? tc_win_fwd(g, a, dtype, st) : simt_fwd(g, a, dtype, st);
}

int fa_window_divide(const float* acc, void* y, int ndim, const int64_t* dims, int64_t W, int64_t stride, int64_t pad,
                     int64_t plane_lo, int64_t nplanes, int64_t channels, int64_t B, int dtype, void* stream) {
  Geo g;
  int rc = windowed_geo(g, ndim, dims, channels, channels, B, W, stride, pad);
  if (rc) return rc;
  if (!acc || !y || !valid_dtype(dtype) || channels <= 0 || B <= 0 || plane_lo < 0 || nplanes <= 0 || plane_lo + nplanes > dims[ndim - 1]) {
    set_error("bad fa_window_divide arguments"); return FA_ERR_INVALID;
  }
  if ((rc = need_device())) return rc;
  long long plane_tokens = 1;
  for (int k = 0; k < ndim - 1; ++k) plane_tokens *= dims[k];
  return slab_divide(g, acc, y, (int)channels, dtype, plane_lo, plane_tokens * nplanes, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------ unfold / fold
int fa_window(const void* x, void* xw, int ndim, const int64_t* dims, int64_t d, int64_t B,
              int64_t W, int64_t stride, int64_t pad, int dtype, void* stream) {
  Geo g;
  int rc = windowed_geo(g, ndim, dims, d, d, B, W, stride, pad);
  if (rc) return rc;
  if ((rc = check_common(g.N, d, d, B, dtype))) return rc;
  if (!x || !xw) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  if ((rc = need_device())) return rc;
  return window_gather(g, x, xw, dtype, static_cast<cudaStream_t>(stream));
}

int fa_unwindow(const void* xw, void* x, int ndim, const int64_t* dims, int64_t d, int64_t B,
                int64_t W, int64_t stride, int64_t pad, int dtype, void* stream) {
  Geo g;
  int rc = windowed_geo(g, ndim, dims, d, d, B, W, stride, pad);
  if (rc) return rc;
  if ((rc = check_common(g.N, d, d, B, dtype))) return rc;
  if (!x || !xw) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  if ((rc = need_device())) return rc;
  return window_scatter(g, xw, x, dtype, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------ dtype conversion
int fa_cast(const void* in, void* out, int64_t n, int from_dtype, int to_dtype, void* stream) {
  if (!in || !out || n <= 0 || !valid_dtype(from_dtype) || !valid_dtype(to_dtype)) { set_error("bad fa_cast arguments"); return FA_ERR_INVALID; }
  int rc = need_device();
  if (rc) return rc;
  return cast_launch(in, out, n, from_dtype, to_dtype, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------ softmax
int fa_softmax(void* out, const void* in, int64_t M, int64_t N, int64_t B, int dim, int dtype, void* stream) {
  if (dim != 1 && dim != 2) { set_error("only softmax in dims 1 or 2 supported"); return FA_ERR_INVALID; }   // src/fused_softmax.jl:12
  if (!valid_dtype(dtype) || M <= 0 || N <= 0 || B <= 0 || !out || !in) { set_error("bad softmax arguments"); return FA_ERR_INVALID; }
  int rc = need_device();
  if (rc) return rc;
  return softmax_launch(out, in, M, N, B, dim, dtype, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// ------------------------------------------------------------------------------ host buffers
namespace {
// Grow-only per-device arena for the host-buffer entry points: cudaMalloc / cudaFree of ~0.5 GB per
// call cost far more than the copies themselves, so the staging memory is kept between calls
// (fa_release_host_staging() frees it).  Calls on one device are serialised by the arena lock.
struct Arena {
  std::mutex mu;
  void* p = nullptr;
  size_t cap = 0;
};
constexpr int kMaxDev = 16;
Arena g_arena[kMaxDev];

// The *_host calls run on `device` and put the caller's current device back on every exit path
// (Julia's CUDA.jl and torch both track the current device per thread).
struct DeviceGuard {
  int prev = -1;
  int rc = FA_OK;
  explicit DeviceGuard(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); set_error("no CUDA device available (libfa_sm100a has no CPU fallback)"); rc = FA_ERR_CUDA; return; }
    if (device < 0 || device >= n || device >= kMaxDev) { set_error("device index %d out of range (0..%d)", device, (n < kMaxDev ? n : kMaxDev) - 1); rc = FA_ERR_INVALID; return; }
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaSetDevice");
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Pageable caller buffers (a plain Julia Array): an async copy from pageable memory is staged by the driver and blocks
// the host, so nothing overlaps.  Unless FA_FLAG_HOST_NO_REGISTER is set, big pageable buffers are page-locked for the
// duration of the call (cudaHostRegister) and released on exit.  Already-pinned buffers (cudaHostAlloc, CUDA.pin,
// torch pin_memory, fa_host_alloc) are used as they are.
struct HostPins {
  void* reg[8];
  int n = 0;
  void pin(const void* p, size_t bytes) {
    if (!p || bytes < ((size_t)4 << 20) || n >= 8) return;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return; }
    if (at.type != cudaMemoryTypeUnregistered) return;
    if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterPortable) == cudaSuccess) reg[n++] = const_cast<void*>(p);
    else cudaGetLastError();            // e.g. overlaps a registered range: fall back to the driver's staged copy
  }
  ~HostPins() { for (int i = 0; i < n; ++i) cudaHostUnregister(reg[i]); }
};

// Batch-chunked pipeline on three streams -- host->device copies, kernels, device->host copies -- over three
// staging buffer sets, so the H2D engine never waits for a kernel or a D2H copy: chunk i+1 (and i+2) stream in while
// chunk i computes and chunk i-1 streams out.  Every tensor is (.., B) column-major, so a batch chunk is a contiguous
// byte range of `bytes_per_b * nb`.  `run(nb, din, dout, stream)` enqueues the kernels of one chunk on device copies.
struct HostIn { const void* h; size_t bytes_per_b; };
struct HostOut { void* h; size_t bytes_per_b; };
constexpr int kMaxIO = 8;

template <typename Run>
int host_pipeline_n(const HostIn* in, int nin, const HostOut* out, int nout, int64_t B, int flags, int device, Run run) {
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  int rc;
  size_t per_b = 0;
  for (int t = 0; t < nin; ++t) per_b += in[t].bytes_per_b;
  for (int t = 0; t < nout; ++t) per_b += out[t].bytes_per_b;
  // The pipeline is bound by the host->device copies (PCIe); what it adds on top is the head (first H2D) and the tail
  // (kernel + D2H of the last chunk), so big jobs are cut into ~24 chunks (16 MiB .. 256 MiB of tensors each).
  int64_t chunk = B;
  const size_t total = per_b * (size_t)B;
  size_t target = total / 24;
  if (target < ((size_t)16 << 20)) target = (size_t)16 << 20;
  if (target > ((size_t)256 << 20)) target = (size_t)256 << 20;
  if (total > 2 * target) { chunk = (int64_t)(target / per_b); if (chunk < 1) chunk = 1; }
  constexpr int NB = 3;
  const int64_t nchunks = (B + chunk - 1) / chunk;
  const int nbuf = nchunks < NB ? (int)nchunks : NB;
  size_t off_in[kMaxIO], off_out[kMaxIO], per_set = 0;
  for (int t = 0; t < nin; ++t) { off_in[t] = per_set; per_set += align256(chunk * in[t].bytes_per_b); }
  for (int t = 0; t < nout; ++t) { off_out[t] = per_set; per_set += align256(chunk * out[t].bytes_per_b); }
  HostPins pins;
  if (!(flags & FA_FLAG_HOST_NO_REGISTER) && nchunks > 1) {
    for (int t = 0; t < nin; ++t) pins.pin(in[t].h, (size_t)B * in[t].bytes_per_b);
    for (int t = 0; t < nout; ++t) pins.pin(out[t].h, (size_t)B * out[t].bytes_per_b);
  }
  Arena& ar = g_arena[device];
  std::lock_guard<std::mutex> lock(ar.mu);
  if (ar.cap < per_set * nbuf) {
    if (ar.p) { cudaFree(ar.p); ar.p = nullptr; ar.cap = 0; }
    void* np = nullptr;
    FA_CUDA_TRY(cudaMalloc(&np, per_set * nbuf));
    ar.p = np; ar.cap = per_set * nbuf;
  }
  struct Lanes {           // streams and events; drained and released on every exit path
    cudaStream_t s[3] = {nullptr, nullptr, nullptr};                // 0: H2D, 1: kernels, 2: D2H
    cudaEvent_t in_done[NB] = {}, run_done[NB] = {}, out_done[NB] = {};
    ~Lanes() {
      for (int i = 0; i < 3; ++i) if (s[i]) { cudaStreamSynchronize(s[i]); cudaStreamDestroy(s[i]); }
      for (int i = 0; i < NB; ++i) { if (in_done[i]) cudaEventDestroy(in_done[i]); if (run_done[i]) cudaEventDestroy(run_done[i]); if (out_done[i]) cudaEventDestroy(out_done[i]); }
    }
  } ln;
  for (int i = 0; i < 3; ++i) FA_CUDA_TRY(cudaStreamCreateWithFlags(&ln.s[i], cudaStreamNonBlocking));
  for (int i = 0; i < nbuf; ++i) {
    FA_CUDA_TRY(cudaEventCreateWithFlags(&ln.in_done[i], cudaEventDisableTiming));
    FA_CUDA_TRY(cudaEventCreateWithFlags(&ln.run_done[i], cudaEventDisableTiming));
    FA_CUDA_TRY(cudaEventCreateWithFlags(&ln.out_done[i], cudaEventDisableTiming));
  }
  int64_t it = 0;
  for (int64_t b0 = 0; b0 < B; b0 += chunk, ++it) {
    const int i = (int)(it % nbuf);
    const int64_t nb = (B - b0 < chunk) ? (B - b0) : chunk;
    char* base = static_cast<char*>(ar.p) + i * per_set;
    void *din[kMaxIO], *dout[kMaxIO];
    for (int t = 0; t < nin; ++t) din[t] = base + off_in[t];
    for (int t = 0; t < nout; ++t) dout[t] = base + off_out[t];
    if (it >= nbuf) FA_CUDA_TRY(cudaStreamWaitEvent(ln.s[0], ln.run_done[i], 0));      // inputs of set i consumed
    for (int t = 0; t < nin; ++t)
      FA_CUDA_TRY(cudaMemcpyAsync(din[t], static_cast<const char*>(in[t].h) + b0 * in[t].bytes_per_b, nb * in[t].bytes_per_b, cudaMemcpyHostToDevice, ln.s[0]));
    FA_CUDA_TRY(cudaEventRecord(ln.in_done[i], ln.s[0]));
    FA_CUDA_TRY(cudaStreamWaitEvent(ln.s[1], ln.in_done[i], 0));
    if (it >= nbuf) FA_CUDA_TRY(cudaStreamWaitEvent(ln.s[1], ln.out_done[i], 0));     // outputs of set i copied out
    if ((rc = run(nb, din, dout, ln.s[1]))) return rc;
    FA_CUDA_TRY(cudaEventRecord(ln.run_done[i], ln.s[1]));
    FA_CUDA_TRY(cudaStreamWaitEvent(ln.s[2], ln.run_done[i], 0));
    for (int t = 0; t < nout; ++t)
      FA_CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(out[t].h) + b0 * out[t].bytes_per_b, dout[t], nb * out[t].bytes_per_b, cudaMemcpyDeviceToHost, ln.s[2]));
    FA_CUDA_TRY(cudaEventRecord(ln.out_done[i], ln.s[2]));
  }
  for (int i = 0; i < 3; ++i) FA_CUDA_TRY(cudaStreamSynchronize(ln.s[i]));
  return FA_OK;
}

// forward form: inputs q, k, v; outputs o, l, m
template <typename Run>
int host_pipeline(const void* q, const void* k, const void* v, void* o, float* l, float* m,
                  size_t q_elems_per_b, size_t v_elems_per_b, size_t stat_per_b, int64_t B, int dtype,
                  int flags, int device, Run run) {
  const size_t esz = dtype_size(dtype);
  const HostIn in[3] = {{q, q_elems_per_b * esz}, {k, q_elems_per_b * esz}, {v, v_elems_per_b * esz}};
  const HostOut out[3] = {{o, v_elems_per_b * esz}, {l, stat_per_b * 4}, {m, stat_per_b * 4}};
  return host_pipeline_n(in, 3, out, 3, B, flags, device, [&](int64_t nb, void** di, void** dout, cudaStream_t st) {
    return run(nb, di[0], di[1], di[2], dout[0], static_cast<float*>(dout[1]), static_cast<float*>(dout[2]), st);
  });
}

// backward form: inputs q, k, v, (o,) dO, l, m; outputs dq, dk, dv.  `ws_bytes(nb)` sizes the per-chunk workspace.
template <typename Run>
int host_pipeline_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* l, const float* m,
                      void* dq, void* dk, void* dvo, size_t q_elems_per_b, size_t v_elems_per_b, size_t stat_per_b,
                      int64_t B, int dtype, int flags, int device, Run run) {
  const size_t esz = dtype_size(dtype);
  HostIn in[7]; int nin = 0;
  in[nin++] = {q, q_elems_per_b * esz}; in[nin++] = {k, q_elems_per_b * esz}; in[nin++] = {v, v_elems_per_b * esz};
  const int io = o ? nin : -1;
  if (o) in[nin++] = {o, v_elems_per_b * esz};
  const int ig = nin; in[nin++] = {d_o, v_elems_per_b * esz};
  const int il = nin; in[nin++] = {l, stat_per_b * 4};
  const int im = nin; in[nin++] = {m, stat_per_b * 4};
  const HostOut out[3] = {{dq, q_elems_per_b * esz}, {dk, q_elems_per_b * esz}, {dvo, v_elems_per_b * esz}};
  return host_pipeline_n(in, nin, out, 3, B, flags, device, [&](int64_t nb, void** di, void** dout, cudaStream_t st) {
    return run(nb, di[0], di[1], di[2], io >= 0 ? di[io] : nullptr, di[ig], static_cast<const float*>(di[il]), static_cast<const float*>(di[im]),
               dout[0], dout[1], dout[2], st);
  });
}

// Per-device workspace for the host entry points (kernels of consecutive chunks are serialised on the pipeline's
// compute stream, so ONE workspace sized for the largest chunk serves them all); kept like the staging arena.
struct WsArena { std::mutex mu; void* p = nullptr; size_t cap = 0; };
WsArena g_ws[kMaxDev];
int ws_reserve(WsArena& wa, size_t bytes) {
  if (wa.cap >= bytes) return FA_OK;
  if (wa.p) { cudaFree(wa.p); wa.p = nullptr; wa.cap = 0; }
  void* np = nullptr;
  FA_CUDA_TRY(cudaMalloc(&np, bytes ? bytes : 256));
  wa.p = np; wa.cap = bytes ? bytes : 256;
  return FA_OK;
}
}  // namespace

extern "C" {

int fa_dense_fwd_host(const void* q, const void* k, const void* v, void* o, float* l, float* m,
                      int64_t N, int64_t d, int64_t dv, int64_t B, int dtype, int flags, int device) {
  int rc = check_common(N, d, dv, B, dtype);
  if (rc) return rc;
  if (!q || !k || !v || !o || !l || !m) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  return host_pipeline(q, k, v, o, l, m, (size_t)N * d, (size_t)N * dv, (size_t)N, B, dtype, flags, device,
      [&](int64_t nb, void* dq, void* dk, void* dvp, void* dop, float* dl, float* dm, cudaStream_t st) {
        return fa_dense_fwd(dq, dk, dvp, dop, dl, dm, N, d, dv, nb, dtype, flags, st);
      });
}

int fa_circulant_fwd_host(const void* q, const void* k, const void* v, void* o, float* l, float* m,
                          int64_t N, int64_t d, int64_t dv, int64_t B, int64_t W, int dtype, int flags, int device) {
  int rc = check_common(N, d, dv, B, dtype);
  if (rc) return rc;
  if (!q || !k || !v || !o || !l || !m) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  return host_pipeline(q, k, v, o, l, m, (size_t)N * d, (size_t)N * dv, (size_t)N, B, dtype, flags, device,
      [&](int64_t nb, void* dq, void* dk, void* dvp, void* dop, float* dl, float* dm, cudaStream_t st) {
        return fa_circulant_fwd(dq, dk, dvp, dop, dl, dm, N, d, dv, nb, W, dtype, flags, st);
      });
}

int fa_windowed_fwd_host(const void* q, const void* k, const void* v, void* y, float* l, float* m,
                         int ndim, const int64_t* dims, int64_t d, int64_t dv, int64_t B,
                         int64_t W, int64_t stride, int64_t pad, int dtype, int flags, int device) {
  Geo g;
  int rc = windowed_geo(g, ndim, dims, d, dv, B, W, stride, pad);
  if (rc) return rc;
  if ((rc = check_common(g.N, d, dv, B, dtype))) return rc;
  if (!q || !k || !v || !y || !l || !m) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  if (flags & FA_FLAG_OUT_F32) { set_error("FA_FLAG_OUT_F32 is a device-pointer option"); return FA_ERR_UNSUPPORTED; }
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  const size_t wsb = fa_workspace_bytes_windowed_fwd(ndim, dims, d, dv, B, W, stride, pad, dtype, flags);   // largest chunk <= B
  WsArena& wa = g_ws[device];
  std::lock_guard<std::mutex> lock(wa.mu);
  if ((rc = ws_reserve(wa, wsb))) return rc;
  void* w = wa.p;
  return host_pipeline(q, k, v, y, l, m, (size_t)g.N * d, (size_t)g.N * dv, (size_t)g.WD * g.L, B, dtype, flags, device,
      [&](int64_t nb, void* dq, void* dk, void* dvp, void* dop, float* dl, float* dm, cudaStream_t st) {
        return fa_windowed_fwd(dq, dk, dvp, dop, dl, dm, ndim, dims, d, dv, nb, W, stride, pad, dtype, flags, w, wsb, st);
      });
}

int fa_dense_bwd_host(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                      const float* l, const float* m, void* dq, void* dk, void* dv_out,
                      int64_t N, int64_t d, int64_t dv, int64_t B, int dtype, int flags, int device) {
  int rc = check_common(N, d, dv, B, dtype);
  if (rc) return rc;
  if (!q || !k || !v || !o || !d_o || !l || !m || !dq || !dk || !dv_out) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  if (flags & FA_FLAG_OUT_F32) { set_error("FA_FLAG_OUT_F32 is a device-pointer option"); return FA_ERR_UNSUPPORTED; }
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  const size_t wsb = fa_workspace_bytes_dense_bwd(N, d, dv, B, dtype, flags);
  WsArena& wa = g_ws[device];
  std::lock_guard<std::mutex> lock(wa.mu);
  if ((rc = ws_reserve(wa, wsb))) return rc;
  void* w = wa.p;
  return host_pipeline_bwd(q, k, v, o, d_o, l, m, dq, dk, dv_out, (size_t)N * d, (size_t)N * dv, (size_t)N, B, dtype, flags, device,
      [&](int64_t nb, void* a, void* b, void* c, void* oo, void* g, const float* ll, const float* mm, void* x, void* y, void* z, cudaStream_t st) {
        return fa_dense_bwd(a, b, c, oo, g, ll, mm, x, y, z, N, d, dv, nb, dtype, flags, w, wsb, st);
      });
}

int fa_circulant_bwd_host(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                          const float* l, const float* m, void* dq, void* dk, void* dv_out,
                          int64_t N, int64_t d, int64_t dv, int64_t B, int64_t W, int dtype, int flags, int device) {
  int rc = check_common(N, d, dv, B, dtype);
  if (rc) return rc;
  if (!q || !k || !v || !o || !d_o || !l || !m || !dq || !dk || !dv_out) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  if (flags & FA_FLAG_OUT_F32) { set_error("FA_FLAG_OUT_F32 is a device-pointer option"); return FA_ERR_UNSUPPORTED; }
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  const size_t wsb = fa_workspace_bytes_circulant_bwd(N, d, dv, B, W, dtype, flags);
  WsArena& wa = g_ws[device];
  std::lock_guard<std::mutex> lock(wa.mu);
  if ((rc = ws_reserve(wa, wsb))) return rc;
  void* w = wa.p;
  return host_pipeline_bwd(q, k, v, o, d_o, l, m, dq, dk, dv_out, (size_t)N * d, (size_t)N * dv, (size_t)N, B, dtype, flags, device,
      [&](int64_t nb, void* a, void* b, void* c, void* oo, void* g, const float* ll, const float* mm, void* x, void* y, void* z, cudaStream_t st) {
        return fa_circulant_bwd(a, b, c, oo, g, ll, mm, x, y, z, N, d, dv, nb, W, dtype, flags, w, wsb, st);
      });
}

int fa_windowed_bwd_host(const void* q, const void* k, const void* v, const void* d_y,
                         const float* l, const float* m, void* dq, void* dk, void* dv_out,
                         int ndim, const int64_t* dims, int64_t d, int64_t dv, int64_t B,
                         int64_t W, int64_t stride, int64_t pad, int dtype, int flags, int device) {
  Geo g;
  int rc = windowed_geo(g, ndim, dims, d, dv, B, W, stride, pad);
  if (rc) return rc;
  if ((rc = check_common(g.N, d, dv, B, dtype))) return rc;
  if (!q || !k || !v || !d_y || !l || !m || !dq || !dk || !dv_out) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  if (flags & FA_FLAG_OUT_F32) { set_error("FA_FLAG_OUT_F32 is a device-pointer option"); return FA_ERR_UNSUPPORTED; }
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  const size_t wsb = fa_workspace_bytes_windowed_bwd(ndim, dims, d, dv, B, W, stride, pad, dtype, flags);
  WsArena& wa = g_ws[device];
  std::lock_guard<std::mutex> lock(wa.mu);
  if ((rc = ws_reserve(wa, wsb))) return rc;
  void* w = wa.p;
  return host_pipeline_bwd(q, k, v, nullptr, d_y, l, m, dq, dk, dv_out, (size_t)g.N * d, (size_t)g.N * dv, (size_t)g.WD * g.L, B, dtype, flags, device,
      [&](int64_t nb, void* a, void* b, void* c, void*, void* gy, const float* ll, const float* mm, void* x, void* y, void* z, cudaStream_t st) {
        return fa_windowed_bwd(a, b, c, gy, ll, mm, x, y, z, ndim, dims, d, dv, nb, W, stride, pad, dtype, flags, w, wsb, st);
      });
}

// Page-locked host memory placed on the NUMA node of `device` (its PCIe root): the thread is bound to the device's
// local CPUs (sysfs local_cpulist of the PCI function) while the pages are allocated and first touched, then its
// affinity is restored.  Host buffers allocated this way keep every H2D / D2H copy off the inter-socket link, which
// is what caps the aggregate rate when all 8 GPUs of a box stream from host memory at once.
int fa_host_alloc(void** out, size_t bytes, int device) {
  if (!out || bytes == 0) { set_error("fa_host_alloc: bad arguments"); return FA_ERR_INVALID; }
  *out = nullptr;
  DeviceGuard guard(device);
  if (guard.rc) return guard.rc;
  cpu_set_t old_set, new_set;
  bool bound = false;
  char bus[32] = "";
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) == cudaSuccess && sched_getaffinity(0, sizeof(old_set), &old_set) == 0) {
    for (char* c = bus; *c; ++c) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
    if (FILE* f = fopen(path, "r")) {
      char line[1024] = "";
      if (fgets(line, sizeof(line), f)) {
        CPU_ZERO(&new_set);
        int n_set = 0;
        for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
          int a = 0, b = 0;
          const int got = sscanf(tok, "%d-%d", &a, &b);
          if (got == 1) b = a;
          if (got >= 1) for (int c = a; c <= b && c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &old_set)) { CPU_SET(c, &new_set); ++n_set; }
        }
        if (n_set > 0 && sched_setaffinity(0, sizeof(new_set), &new_set) == 0) bound = true;
      }
      fclose(f);
    }
  } else cudaGetLastError();
  void* p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
  if (e == cudaSuccess) {                                   // first touch under the binding (cudaHostAlloc usually has, be sure)
    volatile char* c = static_cast<volatile char*>(p);
    for (size_t off = 0; off < bytes; off += 4096) c[off] = 0;
  }
  if (bound) sched_setaffinity(0, sizeof(old_set), &old_set);
  if (e != cudaSuccess) return cuda_fail(e, "cudaHostAlloc");
  *out = p;
  return FA_OK;
}

int fa_host_free(void* p) {
  if (!p) return FA_OK;
  FA_CUDA_TRY(cudaFreeHost(p));
  return FA_OK;
}

int fa_release_host_staging(void) {
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
  for (int d = 0; d < kMaxDev; ++d) {
    std::lock_guard<std::mutex> lock(g_arena[d].mu);
    std::lock_guard<std::mutex> lock2(g_ws[d].mu);
    if (g_arena[d].p || g_ws[d].p) {
      cudaSetDevice(d);
      if (g_arena[d].p) cudaFree(g_arena[d].p);
      if (g_ws[d].p) cudaFree(g_ws[d].p);
      g_arena[d].p = nullptr; g_arena[d].cap = 0;
      g_ws[d].p = nullptr; g_ws[d].cap = 0;
    }
  }
  if (prev >= 0) cudaSetDevice(prev);
  return FA_OK;
}

}  // extern "C"
