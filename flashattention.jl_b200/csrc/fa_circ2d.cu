// fa_circ2d.cu -- 2-D circulant (periodic neighbourhood) attention, forward and backward.
//
// SURVEY 8f-2 / reference README.md:38-41,53: `circulant_fa` exists only in 1-D in the reference
// (src/circulant.jl:9-118); the 2-D form is its stated todo.  Definition used here, the direct product of
// the 1-D one (src/utils.jl:6-17, p = (W-1) / 2, integer division):
//     keys of query (x, y) = { (mod(x - p + s, X), mod(y - p + t, Y)) : s, t = 0..W-1 },   W <= min(X, Y)
//     O = softmax_over_the_W^2_keys(tau q.k) V,  tau = 1/sqrt(d);  l, m as in dense_fa (src/dense.jl:12-18).
// Arrays are (X, Y, d, B) column-major = [B][d][Y][X], so consecutive threads = consecutive x read
// consecutive addresses for every (channel, neighbour offset): all loads coalesce (up to the wrap).
//
// Exact fp32 math (FFMA, no tensor cores) for every dtype: this is the parity-first implementation of a
// "next" row; one thread per query (forward, dQ) or per key (dK, dV), scores of the W^2 neighbours kept
// in a per-thread local array.  Deterministic: no atomics.
#include "fa_common.cuh"

namespace fa {
namespace {

constexpr int MAXW2 = 256;     // W <= 16

struct C2Params {
  int X, Y, d, dv, W, p;
  long long N, B;
  float tau;
};

// token of the (s, t)-th neighbour of (x, y); sign = +1: keys of a query, -1: queries of a key
__device__ __forceinline__ int nb_token(const C2Params& g, int x, int y, int s, int t, int sign) {
  int xx = sign > 0 ? x - g.p + s : x + g.p - s;
  int yy = sign > 0 ? y - g.p + t : y + g.p - t;
  xx %= g.X; if (xx < 0) xx += g.X;
  yy %= g.Y; if (yy < 0) yy += g.Y;
  return yy * g.X + xx;
}

template <typename T>
__global__ void c2_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                              T* __restrict__ o, float* __restrict__ l, float* __restrict__ m, const C2Params g) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.N * g.B) return;
  const long long b = idx / g.N;
  const int tok = (int)(idx - b * g.N), x = tok % g.X, y = tok / g.X;
  const int W2 = g.W * g.W;
  int kt[MAXW2];
  float sc[MAXW2];
  for (int t = 0; t < g.W; ++t)
    for (int s = 0; s < g.W; ++s) kt[t * g.W + s] = nb_token(g, x, y, s, t, +1);
  const T* qb = q + b * g.d * g.N + tok;
  const T* kb = k + b * g.d * g.N;
  float mx = -INFINITY;
  for (int j = 0; j < W2; ++j) {
    float acc = 0.f;
    const T* kj = kb + kt[j];
    for (int c = 0; c < g.d; ++c) acc = fmaf(to_f32(qb[(long long)c * g.N]), to_f32(kj[(long long)c * g.N]), acc);
    acc *= g.tau;
    sc[j] = acc;
    mx = fmaxf(mx, acc);
  }
  float sum = 0.f;
  for (int j = 0; j < W2; ++j) { const float pj = __expf(sc[j] - mx); sc[j] = pj; sum += pj; }
  const float inv = 1.f / sum;
  const T* vb = v + b * g.dv * g.N;
  T* ob = o + b * g.dv * g.N + tok;
  for (int c = 0; c < g.dv; ++c) {
    float acc = 0.f;
    const T* vc = vb + (long long)c * g.N;
    for (int j = 0; j < W2; ++j) acc = fmaf(sc[j], to_f32(vc[kt[j]]), acc);
    ob[(long long)c * g.N] = from_f32<T>(acc * inv);
  }
  l[idx] = sum;
  m[idx] = mx;
}

// delta_i = sum_c dO_ic O_ic
template <typename T>
__global__ void c2_delta_kernel(const T* __restrict__ o, const T* __restrict__ d_o, float* __restrict__ delta, const C2Params g) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.N * g.B) return;
  const long long b = idx / g.N, tok = idx - b * g.N;
  const T* po = o + b * g.dv * g.N + tok;
  const T* pg = d_o + b * g.dv * g.N + tok;
  float acc = 0.f;
  for (int c = 0; c < g.dv; ++c) acc = fmaf(to_f32(po[(long long)c * g.N]), to_f32(pg[(long long)c * g.N]), acc);
  delta[idx] = acc;
}

// OWNER_Q: thread == query i, neighbours are its keys      -> dQ_i = tau sum_j dS_ij K_j
// !OWNER_Q: thread == key j, neighbours are its queries     -> dK_j = tau sum_i dS_ij Q_i, dV_j = sum_i P_ij dO_i
template <typename T, bool OWNER_Q>
__global__ void c2_bwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                              const T* __restrict__ d_o, const float* __restrict__ l, const float* __restrict__ m,
                              const float* __restrict__ delta, T* __restrict__ dq, T* __restrict__ dk, T* __restrict__ dvo,
                              const C2Params g) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.N * g.B) return;
  const long long b = idx / g.N;
  const int tok = (int)(idx - b * g.N), x = tok % g.X, y = tok / g.X;
  const int W2 = g.W * g.W;
  int nt[MAXW2];
  float pp[MAXW2], ds[MAXW2];
  for (int t = 0; t < g.W; ++t)
    for (int s = 0; s < g.W; ++s) nt[t * g.W + s] = nb_token(g, x, y, s, t, OWNER_Q ? +1 : -1);
  const T* qb = q + b * g.d * g.N;
  const T* kb = k + b * g.d * g.N;
  const T* vb = v + b * g.dv * g.N;
  const T* gb = d_o + b * g.dv * g.N;
  const float* lb = l + b * g.N;
  const float* mb = m + b * g.N;
  const float* db = delta + b * g.N;
  for (int j = 0; j < W2; ++j) {
    const int qi = OWNER_Q ? tok : nt[j], kj = OWNER_Q ? nt[j] : tok;
    float s = 0.f, dp = 0.f;
    for (int c = 0; c < g.d; ++c) s = fmaf(to_f32(qb[(long long)c * g.N + qi]), to_f32(kb[(long long)c * g.N + kj]), s);
    for (int c = 0; c < g.dv; ++c) dp = fmaf(to_f32(gb[(long long)c * g.N + qi]), to_f32(vb[(long long)c * g.N + kj]), dp);
    const float p = __expf(s * g.tau - mb[qi]) / lb[qi];
    pp[j] = p;
    ds[j] = p * (dp - db[qi]);
  }
  if (OWNER_Q) {
    T* out = dq + b * g.d * g.N + tok;
    for (int c = 0; c < g.d; ++c) {
      float acc = 0.f;
      const T* kc = kb + (long long)c * g.N;
      for (int j = 0; j < W2; ++j) acc = fmaf(ds[j], to_f32(kc[nt[j]]), acc);
      out[(long long)c * g.N] = from_f32<T>(acc * g.tau);
    }
  } else {
    T* outk = dk + b * g.d * g.N + tok;
    for (int c = 0; c < g.d; ++c) {
      float acc = 0.f;
      const T* qc = qb + (long long)c * g.N;
      for (int j = 0; j < W2; ++j) acc = fmaf(ds[j], to_f32(qc[nt[j]]), acc);
      outk[(long long)c * g.N] = from_f32<T>(acc * g.tau);
    }
    T* outv = dvo + b * g.dv * g.N + tok;
    for (int c = 0; c < g.dv; ++c) {
      float acc = 0.f;
      const T* gc = gb + (long long)c * g.N;
      for (int j = 0; j < W2; ++j) acc = fmaf(pp[j], to_f32(gc[nt[j]]), acc);
      outv[(long long)c * g.N] = from_f32<T>(acc);
    }
  }
}

int check(int64_t X, int64_t Y, int64_t d, int64_t dv, int64_t B, int64_t W, int dtype, C2Params& g) {
  if (dtype != FA_F32 && dtype != FA_F16 && dtype != FA_BF16) { set_error("bad dtype"); return FA_ERR_INVALID; }
  if (X <= 0 || Y <= 0 || d <= 0 || dv <= 0 || B <= 0 || X * Y > 0x7fffffffLL || d > 4096 || dv > 4096) { set_error("bad 2-D circulant shape"); return FA_ERR_INVALID; }
  if (W <= 0 || W > X || W > Y) { set_error("2-D circulant window must satisfy 0 < W <= min(X, Y) (got W=%lld, X=%lld, Y=%lld)", (long long)W, (long long)X, (long long)Y); return FA_ERR_INVALID; }
  if (W * W > MAXW2) { set_error("2-D circulant window too large (W^2 <= %d)", MAXW2); return FA_ERR_UNSUPPORTED; }
  g.X = (int)X; g.Y = (int)Y; g.d = (int)d; g.dv = (int)dv; g.W = (int)W; g.p = (int)((W - 1) / 2);
  g.N = X * Y; g.B = B; g.tau = 1.0f / sqrtf((float)d);
  return FA_OK;
}

template <typename T>
int fwd_t(const void* q, const void* k, const void* v, void* o, float* l, float* m, const C2Params& g, cudaStream_t st) {
  const long long n = g.N * g.B;
  c2_fwd_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(static_cast<const T*>(q), static_cast<const T*>(k), static_cast<const T*>(v),
                                                                 static_cast<T*>(o), l, m, g);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}
template <typename T>
int bwd_t(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* l, const float* m,
          void* dq, void* dk, void* dvo, float* delta, const C2Params& g, cudaStream_t st) {
  const long long n = g.N * g.B;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  c2_delta_kernel<T><<<blocks, 128, 0, st>>>(static_cast<const T*>(o), static_cast<const T*>(d_o), delta, g);
  FA_CUDA_TRY(cudaGetLastError());
  c2_bwd_kernel<T, true><<<blocks, 128, 0, st>>>(static_cast<const T*>(q), static_cast<const T*>(k), static_cast<const T*>(v), static_cast<const T*>(d_o),
                                                 l, m, delta, static_cast<T*>(dq), nullptr, nullptr, g);
  FA_CUDA_TRY(cudaGetLastError());
  c2_bwd_kernel<T, false><<<blocks, 128, 0, st>>>(static_cast<const T*>(q), static_cast<const T*>(k), static_cast<const T*>(v), static_cast<const T*>(d_o),
                                                  l, m, delta, nullptr, static_cast<T*>(dk), static_cast<T*>(dvo), g);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

int device_ok() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); set_error("no CUDA device available (libfa_sm100a has no CPU fallback)"); return FA_ERR_CUDA; }
  return FA_OK;
}

}  // namespace
}  // namespace fa

using namespace fa;

extern "C" {

int fa_circulant2d_index(int64_t X, int64_t Y, int64_t W, int64_t* keys) {
  if (X <= 0 || Y <= 0 || W <= 0 || W > X || W > Y || !keys) { set_error("need 0 < W <= min(X, Y) and a keys buffer"); return FA_ERR_INVALID; }
  const int64_t p = (W - 1) / 2;
  for (int64_t y = 0; y < Y; ++y)
    for (int64_t x = 0; x < X; ++x)
      for (int64_t t = 0; t < W; ++t)
        for (int64_t s = 0; s < W; ++s) {
          const int64_t xx = ((x - p + s) % X + X) % X, yy = ((y - p + t) % Y + Y) % Y;
          keys[(y * X + x) * W * W + t * W + s] = yy * X + xx;
        }
  return FA_OK;
}

int fa_circulant2d_fwd(const void* q, const void* k, const void* v, void* o, float* l, float* m,
                       int64_t X, int64_t Y, int64_t d, int64_t dv, int64_t B, int64_t W, int dtype, int flags, void* stream) {
  C2Params g;
  int rc = check(X, Y, d, dv, B, W, dtype, g);
  if (rc) return rc;
  if (!q || !k || !v || !o || !l || !m) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  if ((rc = device_ok())) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // 16-bit, d = dv in {64, 128}, X % 64 == 0: the tcgen05 band kernel walking W key rows (fa_tc_band.cu)
  const bool aligned = ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (!(flags & FA_FLAG_FORCE_SIMT) && aligned && tc_band2d_supported(X, Y, d, dv, B, W, dtype)) {
    set_path("tc");
    return tc_band2d_fwd(q, k, v, o, l, m, X, Y, d, B, W, dtype, st);
  }
  set_path("simt");
  if (dtype == FA_F32) return fwd_t<float>(q, k, v, o, l, m, g, st);
  if (dtype == FA_F16) return fwd_t<__half>(q, k, v, o, l, m, g, st);
  return fwd_t<__nv_bfloat16>(q, k, v, o, l, m, g, st);
}

size_t fa_workspace_bytes_circulant2d_bwd(int64_t X, int64_t Y, int64_t B) {
  if (X <= 0 || Y <= 0 || B <= 0) return 0;
  return ((size_t)X * Y * B * sizeof(float) + 255) & ~(size_t)255;
}

// geometry of the tcgen05 backward for the 2-D neighbourhood: a circulant Geo that carries the image extents
static bool tc2d_bwd_geo(Geo& g, int64_t X, int64_t Y, int64_t d, int64_t dv, int64_t B, int64_t W, int dtype) {
  if (dtype != FA_BF16 && dtype != FA_F16) return false;
  if (d != dv || (d != 64 && d != 128) || X % 64 != 0 || W > 64) return false;
  if (((X + 127) / 128) * Y > 0x7fffffffLL) return false;
  memset(&g, 0, sizeof(g));
  g.mode = MODE_CIRCULANT; g.d = (int)d; g.dv = (int)dv; g.N = X * Y; g.B = B; g.W = (int)W; g.p = (int)((W - 1) / 2);
  g.nd = 2; g.s[0] = (int)X; g.s[1] = (int)Y; g.s[2] = 1;
  g.tau = 1.0f / sqrtf((float)d);
  return tc_bwd_supported(g, dtype);
}

// workspace that also admits the tcgen05 backward (16-bit, d = dv in {64, 128}, X % 64 == 0); with only the
// fa_workspace_bytes_circulant2d_bwd(X, Y, B) amount the exact fp32 kernels run
size_t fa_workspace_bytes_circulant2d_bwd_ex(int64_t X, int64_t Y, int64_t d, int64_t dv, int64_t B, int64_t W, int dtype, int flags) {
  const size_t base = fa_workspace_bytes_circulant2d_bwd(X, Y, B);
  Geo g;
  if (base == 0 || (flags & FA_FLAG_FORCE_SIMT) || !tc2d_bwd_geo(g, X, Y, d, dv, B, W, dtype)) return base;
  const size_t tc = tc_bwd_workspace_bytes(g, dtype, flags);
  return tc > base ? tc : base;
}

int fa_circulant2d_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                       const float* l, const float* m, void* dq, void* dk, void* dv_out,
                       int64_t X, int64_t Y, int64_t d, int64_t dv, int64_t B, int64_t W, int dtype, int flags,
                       void* workspace, size_t workspace_bytes, void* stream) {
  C2Params g;
  int rc = check(X, Y, d, dv, B, W, dtype, g);
  if (rc) return rc;
  if (!q || !k || !v || !o || !d_o || !l || !m || !dq || !dk || !dv_out) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  if (!workspace || workspace_bytes < fa_workspace_bytes_circulant2d_bwd(X, Y, B)) { set_error("workspace too small"); return FA_ERR_WORKSPACE; }
  if ((rc = device_ok())) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    // tcgen05 path: the 1-D circulant backward kernels of fa_tc_bwd.cu walking W image rows (needs the _ex workspace)
    Geo tg;
    const bool aligned = ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                           reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(d_o)) & 15) == 0;
    if (!(flags & FA_FLAG_FORCE_SIMT) && aligned && tc2d_bwd_geo(tg, X, Y, d, dv, B, W, dtype) &&
        workspace_bytes >= tc_bwd_workspace_bytes(tg, dtype, flags)) {
      BwdArgs a{q, k, v, o, d_o, l, m, dq, dk, dv_out, nullptr, nullptr, nullptr, static_cast<float*>(workspace)};
      set_path("tc");
      return tc_bwd(tg, a, dtype, flags, workspace, st);
    }
  }
  set_path("simt");
  float* delta = static_cast<float*>(workspace);
  if (dtype == FA_F32) return bwd_t<float>(q, k, v, o, d_o, l, m, dq, dk, dv_out, delta, g, st);
  if (dtype == FA_F16) return bwd_t<__half>(q, k, v, o, d_o, l, m, dq, dk, dv_out, delta, g, st);
  return bwd_t<__nv_bfloat16>(q, k, v, o, d_o, l, m, dq, dk, dv_out, delta, g, st);
}

}  // extern "C"
