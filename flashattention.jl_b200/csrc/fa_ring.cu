// fa_ring.cu -- multi-GPU pieces of the hot path.
//
// Every (batch*head) slice and every window is independent, so the normal multi-GPU mode is a
// contiguous split of the trailing batch dim with NO collective (fa_shard_batch).  The one real
// exchange step is a SINGLE long sequence that is sharded over the ranks by tokens: ring attention.
//   rank r holds tokens [r*Nl, (r+1)*Nl) of q, k, v.  For s = 0..G-1 it runs the dense forward of
//   its queries against the K/V block it currently holds (block of rank (r-s) mod G), merges the
//   partial (O, l, m) with the online-softmax update of the reference (src/dense.jl:82-91), and
//   -- on a second stream, overlapped with that compute -- passes the block to rank r+1 and
//   receives the next one from rank r-1 with ncclSend/ncclRecv over NVLink.
// NCCL is not linked: the symbols are resolved at run time from the libnccl the caller's
// communicator was created with (torch's bundled one), so the library still loads without NCCL.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>

#include "fa_common.cuh"

namespace fa {
namespace {

struct NcclApi {
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

const NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy already in the process
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW);
    if (h) {
      api.Send = reinterpret_cast<decltype(api.Send)>(dlsym(h, "ncclSend"));
      api.Recv = reinterpret_cast<decltype(api.Recv)>(dlsym(h, "ncclRecv"));
      api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(h, "ncclGroupStart"));
      api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
      api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
      api.ok = api.Send && api.Recv && api.GroupStart && api.GroupEnd;
    }
  });
  return api;
}

#define FA_NCCL_TRY(expr)                                                                     \
  do {                                                                                        \
    ncclResult_t _r = (expr);                                                                 \
    if (_r != ncclSuccess) {                                                                  \
      set_error("NCCL error %d (%s) at %s", (int)_r, nc.GetErrorString ? nc.GetErrorString(_r) : "?", #expr); \
      return FA_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

// (Oa, la, ma) <- merge((Oa, la, ma), (Ob, lb, mb)); Oa is fp32 and normalised after every merge
// exactly like O in src/dense.jl:82-91.  `out` (optional) receives Oa in the caller's dtype.
template <typename TB, typename T>
__global__ void merge_partials_kernel(float* __restrict__ oa, float* __restrict__ la, float* __restrict__ ma,
                                      const TB* __restrict__ ob, const float* __restrict__ lb, const float* __restrict__ mb,
                                      T* __restrict__ out, long long N, int dv, int first) {
  const long long b = blockIdx.y;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const long long si = b * N + i;
    const float m1 = first ? -INFINITY : ma[si], l1 = first ? 0.f : la[si];
    const float m2 = mb[si], l2 = lb[si];
    const float mn = fmaxf(m1, m2);
    const float w1 = (m1 == -INFINITY) ? 0.f : l1 * __expf(m1 - mn);
    const float w2 = (m2 == -INFINITY) ? 0.f : l2 * __expf(m2 - mn);
    const float ln = w1 + w2;
    const float inv = ln > 0.f ? 1.f / ln : 0.f;
    const float c1 = w1 * inv, c2 = w2 * inv;
    float* po = oa + b * dv * N + i;
    const TB* pb = ob + b * dv * N + i;
    T* pout = out ? out + b * dv * N + i : nullptr;
    for (int c = 0; c < dv; ++c) {
      const float prev = first ? 0.f : po[(long long)c * N];
      const float val = prev * c1 + to_f32(pb[(long long)c * N]) * c2;
      po[(long long)c * N] = val;
      if (pout) pout[(long long)c * N] = from_f32<T>(val);
    }
    la[si] = ln;
    ma[si] = mn;
  }
}

template <typename TB, typename T>
int merge_t(float* oa, float* la, float* ma, const void* ob, const float* lb, const float* mb, void* out,
            long long N, int dv, long long B, int first, cudaStream_t st) {
  const dim3 grid((unsigned)((N + 255) / 256 < 1024 ? (N + 255) / 256 : 1024), (unsigned)B);
  merge_partials_kernel<TB, T><<<grid, 256, 0, st>>>(oa, la, ma, static_cast<const TB*>(ob), lb, mb, static_cast<T*>(out), N, dv, first);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }

// Exchange stream + events of one ring call, released on EVERY exit path (round 1 leaked them on an early error return).
// On destruction the exchange stream is drained first, so no NCCL operation of this call is still in flight when
// its buffers are reused; an NCCL group left open by an error between GroupStart and GroupEnd is closed.
struct RingLanes {
  cudaStream_t xs = nullptr;
  cudaEvent_t ev[8] = {};
  int nev = 0;
  bool group_open = false;
  bool done = false;           // success: the caller's stream already waits on the last exchange event, no host sync needed
  const NcclApi* nc = nullptr;
  int make(int n_events) {
    FA_CUDA_TRY(cudaStreamCreateWithFlags(&xs, cudaStreamNonBlocking));
    for (int i = 0; i < n_events && i < 8; ++i) { FA_CUDA_TRY(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)); ++nev; }
    return FA_OK;
  }
  ~RingLanes() {
    if (group_open && nc && nc->GroupEnd) nc->GroupEnd();
    if (xs) { if (!done) cudaStreamSynchronize(xs); cudaStreamDestroy(xs); }      // destruction itself is deferred by the runtime
    for (int i = 0; i < nev; ++i) cudaEventDestroy(ev[i]);
  }
};
#define FA_NCCL_GROUP_BEGIN(L) do { FA_NCCL_TRY(nc.GroupStart()); (L).group_open = true; } while (0)
#define FA_NCCL_GROUP_END(L) do { (L).group_open = false; FA_NCCL_TRY(nc.GroupEnd()); } while (0)

// acc = (first ? 0 : acc) + part      (fp32, grid-stride)
__global__ void accumulate_kernel(float* __restrict__ acc, const float* __restrict__ part, size_t n, int first) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc[i] = (first ? 0.f : acc[i]) + part[i];
}
// part (in `dtype`) added into acc (fp32): the exact-fp32 SIMT fallback writes its gradients in `dtype`
template <typename T>
__global__ void accumulate_typed_kernel(float* __restrict__ acc, const T* __restrict__ part, size_t n, int first) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc[i] = (first ? 0.f : acc[i]) + to_f32<T>(part[i]);
}
template <typename T>
__global__ void cast_out_kernel(T* __restrict__ out, const float* __restrict__ acc, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = from_f32<T>(acc[i]);
}
unsigned ew_blocks(size_t n) { const size_t b = (n + 255) / 256; return (unsigned)(b < 148 * 16 ? (b ? b : 1) : 148 * 16); }

int accumulate(float* acc, const void* part, size_t n, int part_f32, int dtype, int first, cudaStream_t st) {
  if (part_f32 || dtype == FA_F32) accumulate_kernel<<<ew_blocks(n), 256, 0, st>>>(acc, static_cast<const float*>(part), n, first);
  else if (dtype == FA_F16) accumulate_typed_kernel<__half><<<ew_blocks(n), 256, 0, st>>>(acc, static_cast<const __half*>(part), n, first);
  else accumulate_typed_kernel<__nv_bfloat16><<<ew_blocks(n), 256, 0, st>>>(acc, static_cast<const __nv_bfloat16*>(part), n, first);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}
int cast_out(void* out, const float* acc, size_t n, int dtype, cudaStream_t st) {
  if (dtype == FA_F32) { FA_CUDA_TRY(cudaMemcpyAsync(out, acc, n * 4, cudaMemcpyDeviceToDevice, st)); return FA_OK; }
  if (dtype == FA_F16) cast_out_kernel<__half><<<ew_blocks(n), 256, 0, st>>>(static_cast<__half*>(out), acc, n);
  else cast_out_kernel<__nv_bfloat16><<<ew_blocks(n), 256, 0, st>>>(static_cast<__nv_bfloat16*>(out), acc, n);
  FA_CUDA_TRY(cudaGetLastError());
  return FA_OK;
}

}  // namespace

// blk_f32: the block partial is float32 although `dtype` (the type of `out`) is 16-bit
int merge_partials(float* oa, float* la, float* ma, const void* ob, const float* lb, const float* mb, void* out,
                   long long N, int dv, long long B, int dtype, int blk_f32, int first, cudaStream_t st) {
  if (dtype == FA_F32) return merge_t<float, float>(oa, la, ma, ob, lb, mb, out, N, dv, B, first, st);
  if (dtype == FA_F16) return blk_f32 ? merge_t<float, __half>(oa, la, ma, ob, lb, mb, out, N, dv, B, first, st)
                                      : merge_t<__half, __half>(oa, la, ma, ob, lb, mb, out, N, dv, B, first, st);
  return blk_f32 ? merge_t<float, __nv_bfloat16>(oa, la, ma, ob, lb, mb, out, N, dv, B, first, st)
                 : merge_t<__nv_bfloat16, __nv_bfloat16>(oa, la, ma, ob, lb, mb, out, N, dv, B, first, st);
}

}  // namespace fa

using namespace fa;

extern "C" {

int fa_shard_batch(int64_t B, int nranks, int rank, int64_t* begin, int64_t* count) {
  if (B < 0 || nranks <= 0 || rank < 0 || rank >= nranks || !begin || !count) { set_error("bad shard arguments"); return FA_ERR_INVALID; }
  const int64_t base = B / nranks, rem = B % nranks;       // the first `rem` ranks take one extra element
  *count = base + (rank < rem ? 1 : 0);
  *begin = rank * base + (rank < rem ? rank : rem);
  return FA_OK;
}

int fa_merge_partials(float* o_acc, float* l_acc, float* m_acc, const void* o_blk, const float* l_blk,
                      const float* m_blk, void* out, int64_t N, int64_t dv, int64_t B, int dtype, int first, void* stream) {
  if (!o_acc || !l_acc || !m_acc || !o_blk || !l_blk || !m_blk || N <= 0 || dv <= 0 || B <= 0 || B > 65535) { set_error("bad merge arguments"); return FA_ERR_INVALID; }
  if (dtype != FA_F32 && dtype != FA_F16 && dtype != FA_BF16) { set_error("bad dtype"); return FA_ERR_INVALID; }
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); set_error("no CUDA device available (libfa_sm100a has no CPU fallback)"); return FA_ERR_CUDA; }
  return merge_partials(o_acc, l_acc, m_acc, o_blk, l_blk, m_blk, out, N, (int)dv, B, dtype, 0, first, static_cast<cudaStream_t>(stream));
}

size_t fa_workspace_bytes_ring_dense_fwd(int64_t Nl, int64_t d, int64_t dv, int64_t B, int dtype) {
  if (Nl <= 0 || d <= 0 || dv <= 0 || B <= 0) return 0;
  const size_t esz = dtype_size(dtype);
  return 2 * a256((size_t)Nl * d * B * esz) + 2 * a256((size_t)Nl * dv * B * esz)      // K, V receive buffers (double-buffered)
         + a256((size_t)Nl * dv * B * 4) + 2 * a256((size_t)Nl * B * 4)                 // block partial O (fp32), l, m
         + a256((size_t)Nl * dv * B * 4);                                               // fp32 O accumulator
}

int fa_ring_dense_fwd(const void* q, const void* k, const void* v, void* o, float* l, float* m,
                      int64_t Nl, int64_t d, int64_t dv, int64_t B, int dtype, int flags,
                      void* nccl_comm, int rank, int nranks, void* workspace, size_t workspace_bytes, void* stream) {
  if (!q || !k || !v || !o || !l || !m) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  if (Nl <= 0 || d <= 0 || dv <= 0 || B <= 0 || B > 65535 || nranks <= 0 || rank < 0 || rank >= nranks) { set_error("bad ring arguments"); return FA_ERR_INVALID; }
  if (dtype != FA_F32 && dtype != FA_F16 && dtype != FA_BF16) { set_error("bad dtype"); return FA_ERR_INVALID; }
  if (nranks > 1 && !nccl_comm) { set_error("ring attention over %d ranks needs an NCCL communicator", nranks); return FA_ERR_INVALID; }
  if (!workspace || workspace_bytes < fa_workspace_bytes_ring_dense_fwd(Nl, d, dv, B, dtype)) { set_error("workspace too small"); return FA_ERR_WORKSPACE; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); set_error("no CUDA device available (libfa_sm100a has no CPU fallback)"); return FA_ERR_CUDA; }
  const NcclApi& nc = nccl_api();
  if (nranks > 1 && !nc.ok) { set_error("libnccl.so.2 (ncclSend/ncclRecv) could not be resolved"); return FA_ERR_CUDA; }

  const size_t esz = dtype_size(dtype);
  const size_t kb = a256((size_t)Nl * d * B * esz), vb = a256((size_t)Nl * dv * B * esz), sb = a256((size_t)Nl * B * 4);
  char* ws = static_cast<char*>(workspace);
  void* kbuf[2] = {ws, ws + kb};
  void* vbuf[2] = {ws + 2 * kb, ws + 2 * kb + vb};
  const size_t ob4 = a256((size_t)Nl * dv * B * 4);
  void* oblk = ws + 2 * kb + 2 * vb;
  float* lblk = reinterpret_cast<float*>(ws + 2 * kb + 2 * vb + ob4);
  float* mblk = reinterpret_cast<float*>(ws + 2 * kb + 2 * vb + ob4 + sb);
  float* oacc = reinterpret_cast<float*>(ws + 2 * kb + 2 * vb + ob4 + 2 * sb);
  // block partials stay float32 when the tcgen05 forward runs (16-bit rounding of every partial
  // would add up over the ring); the exact-fp32 SIMT fallback writes them in `dtype`
  Geo gd;
  memset(&gd, 0, sizeof(gd));
  gd.mode = MODE_DENSE; gd.d = (int)d; gd.dv = (int)dv; gd.N = Nl; gd.B = B; gd.tau = 1.0f / sqrtf((float)d);
  // d = 32 runs on the band kernel, which has no float32-partial output: the ring keeps the exact-fp32 kernels there
  const bool tc = !(flags & FA_FLAG_FORCE_SIMT) && gd.d != 32 && tc_fwd_supported(gd, dtype);

  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  RingLanes lanes;
  lanes.nc = &nc;
  int rc = FA_OK;
  if (nranks > 1) {
    if ((rc = lanes.make(5))) return rc;
    FA_CUDA_TRY(cudaEventRecord(lanes.ev[4], cs));          // inputs are ready on the caller's stream
    FA_CUDA_TRY(cudaStreamWaitEvent(lanes.xs, lanes.ev[4], 0));
  }
  cudaStream_t xs = lanes.xs;
  cudaEvent_t* ev_compute = &lanes.ev[0];
  cudaEvent_t* ev_comm = &lanes.ev[2];
  const int next = (rank + 1) % nranks, prev = (rank + nranks - 1) % nranks;
  const ncclComm_t comm = static_cast<ncclComm_t>(nccl_comm);
  const void* cur_k = k;
  const void* cur_v = v;
  // Two ways to fold a block's partial into the running (oacc, l, m) on the tcgen05 path:
  //   FA_RING_FUSED=1: the forward kernel merges in its epilogue (FwdArgs::o_f32 = 2): no partial buffer, no merge pass;
  //   default:         float32 partial + merge_partials_kernel.
  // Measured on 2 x B200 (2 x 16384 tokens, d = 128, 8 heads; profiles/r2h_ring.md): fused 2.44 ms, unfused 2.32 ms.  The
  // pair kernel runs one CTA per SM, so the read-modify-write of oacc sits on each CTA's critical path, while the
  // separate merge is a full-bandwidth streaming pass (0.04 ms): fusing removes 3 passes over N x dv x 4 bytes but
  // costs more than it saves, hence opt-in.
  static const int fused = [] { const char* e = getenv("FA_RING_FUSED"); return e ? atoi(e) : 0; }();
  for (int s = 0; s < nranks; ++s) {
    const int bi = s & 1;
    if (s + 1 < nranks) {
      // exchange on the second stream: send the block we hold, receive the next one into buf[bi];
      // buf[bi] was the compute input of step s-1, so wait for that compute first.
      if (s >= 1) FA_CUDA_TRY(cudaStreamWaitEvent(xs, ev_compute[(s - 1) & 1], 0));
      FA_NCCL_GROUP_BEGIN(lanes);
      FA_NCCL_TRY(nc.Send(cur_k, (size_t)Nl * d * B * esz, ncclUint8, next, comm, xs));
      FA_NCCL_TRY(nc.Send(cur_v, (size_t)Nl * dv * B * esz, ncclUint8, next, comm, xs));
      FA_NCCL_TRY(nc.Recv(kbuf[bi], (size_t)Nl * d * B * esz, ncclUint8, prev, comm, xs));
      FA_NCCL_TRY(nc.Recv(vbuf[bi], (size_t)Nl * dv * B * esz, ncclUint8, prev, comm, xs));
      FA_NCCL_GROUP_END(lanes);
      FA_CUDA_TRY(cudaEventRecord(ev_comm[bi], xs));
    }
    // compute on the caller's stream: partial attention against the resident block, merged into the running result
    if (tc && fused) {
      FwdArgs fa_args{q, cur_k, cur_v, oacc, nullptr, l, m, /*o_f32=*/s == 0 ? 1 : 2};
      set_path("tc");
      if ((rc = tc_fwd(gd, fa_args, dtype, cs))) return rc;
      if (s + 1 == nranks && (rc = cast_out(o, oacc, (size_t)Nl * dv * B, dtype, cs))) return rc;
    } else if (tc) {
      FwdArgs fa_args{q, cur_k, cur_v, oblk, nullptr, lblk, mblk, /*o_f32=*/1};
      set_path("tc");
      if ((rc = tc_fwd(gd, fa_args, dtype, cs))) return rc;
      if ((rc = merge_partials(oacc, l, m, oblk, lblk, mblk, s + 1 == nranks ? o : nullptr, Nl, (int)dv, B, dtype, 1, s == 0, cs))) return rc;
    } else {
      if ((rc = fa_dense_fwd(q, cur_k, cur_v, oblk, lblk, mblk, Nl, d, dv, B, dtype, flags, cs))) return rc;
      if ((rc = merge_partials(oacc, l, m, oblk, lblk, mblk, s + 1 == nranks ? o : nullptr, Nl, (int)dv, B, dtype, 0, s == 0, cs))) return rc;
    }
    if (s + 1 < nranks) {
      FA_CUDA_TRY(cudaEventRecord(ev_compute[bi], cs));
      FA_CUDA_TRY(cudaStreamWaitEvent(cs, ev_comm[bi], 0));          // next block must have arrived
      cur_k = kbuf[bi];
      cur_v = vbuf[bi];
    }
  }
  lanes.done = true;   // the last exchange event was waited on by the caller's stream
  return FA_OK;
}

// ------------------------------------------------------------------------------ ring backward
// Layout of the workspace (all 256-byte aligned):
//   kbuf[2], vbuf[2]        K/V blocks in flight (dtype)
//   kacc[2], vacc[2]        fp32 dK/dV accumulators that TRAVEL WITH their K/V block round the ring
//   pq, pk, pv              fp32 partial gradients of the current step
//   qacc                    fp32 dQ accumulator of the local queries
//   bws                     workspace of the per-block backward (fa_workspace_bytes_dense_bwd)
size_t fa_workspace_bytes_ring_dense_bwd(int64_t Nl, int64_t d, int64_t dv, int64_t B, int dtype, int flags) {
  if (Nl <= 0 || d <= 0 || dv <= 0 || B <= 0) return 0;
  const size_t esz = dtype_size(dtype);
  const size_t kb = a256((size_t)Nl * d * B * esz), vb = a256((size_t)Nl * dv * B * esz);
  const size_t k4 = a256((size_t)Nl * d * B * 4), v4 = a256((size_t)Nl * dv * B * 4);
  return 2 * kb + 2 * vb + 2 * k4 + 2 * v4 + (2 * k4 + v4) + k4 + a256(fa_workspace_bytes_dense_bwd(Nl, d, dv, B, dtype, flags));
}

/* Ring attention backward ("backward rotates dK/dV accumulators with K/V").  Rank r holds the
 * (Nl, ., B) shards of q, k, v, o, dO and the GLOBAL statistics l, m of its queries (from
 * fa_ring_dense_fwd).  Step s: flash backward of the local queries against the resident block (rank
 * (r - s) mod G): dQ += partial; the block's dK/dV accumulators (which arrived with it) += partials; then
 * the accumulators are sent on.  After G steps every block's accumulators are back at their owner. */
int fa_ring_dense_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                      const float* l, const float* m, void* dq, void* dk, void* dv_out,
                      int64_t Nl, int64_t d, int64_t dv, int64_t B, int dtype, int flags,
                      void* nccl_comm, int rank, int nranks, void* workspace, size_t workspace_bytes, void* stream) {
  if (!q || !k || !v || !o || !d_o || !l || !m || !dq || !dk || !dv_out) { set_error("NULL tensor pointer"); return FA_ERR_INVALID; }
  if (Nl <= 0 || d <= 0 || dv <= 0 || B <= 0 || B > 65535 || nranks <= 0 || rank < 0 || rank >= nranks) { set_error("bad ring arguments"); return FA_ERR_INVALID; }
  if (dtype != FA_F32 && dtype != FA_F16 && dtype != FA_BF16) { set_error("bad dtype"); return FA_ERR_INVALID; }
  if (nranks > 1 && !nccl_comm) { set_error("ring attention over %d ranks needs an NCCL communicator", nranks); return FA_ERR_INVALID; }
  if (!workspace || workspace_bytes < fa_workspace_bytes_ring_dense_bwd(Nl, d, dv, B, dtype, flags)) { set_error("workspace too small"); return FA_ERR_WORKSPACE; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); set_error("no CUDA device available (libfa_sm100a has no CPU fallback)"); return FA_ERR_CUDA; }
  const NcclApi& nc = nccl_api();
  if (nranks > 1 && !nc.ok) { set_error("libnccl.so.2 (ncclSend/ncclRecv) could not be resolved"); return FA_ERR_CUDA; }

  const size_t esz = dtype_size(dtype);
  const size_t nk = (size_t)Nl * d * B, nv = (size_t)Nl * dv * B;
  const size_t kb = a256(nk * esz), vb = a256(nv * esz), k4 = a256(nk * 4), v4 = a256(nv * 4);
  char* ws = static_cast<char*>(workspace);
  void* kbuf[2] = {ws, ws + kb};                      ws += 2 * kb;
  void* vbuf[2] = {ws, ws + vb};                      ws += 2 * vb;
  float* kacc[2] = {reinterpret_cast<float*>(ws), reinterpret_cast<float*>(ws + k4)};  ws += 2 * k4;
  float* vacc[2] = {reinterpret_cast<float*>(ws), reinterpret_cast<float*>(ws + v4)};  ws += 2 * v4;
  float* pq = reinterpret_cast<float*>(ws);           ws += k4;
  float* pk = reinterpret_cast<float*>(ws);           ws += k4;
  float* pv = reinterpret_cast<float*>(ws);           ws += v4;
  float* qacc = reinterpret_cast<float*>(ws);         ws += k4;
  void* bws = ws;
  const size_t bws_bytes = fa_workspace_bytes_dense_bwd(Nl, d, dv, B, dtype, flags);

  Geo gd;
  memset(&gd, 0, sizeof(gd));
  gd.mode = MODE_DENSE; gd.d = (int)d; gd.dv = (int)dv; gd.N = Nl; gd.B = B; gd.tau = 1.0f / sqrtf((float)d);
  const bool tc = !(flags & FA_FLAG_FORCE_SIMT) && tc_bwd_supported(gd, dtype);

  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  RingLanes lanes;
  lanes.nc = &nc;
  if (nranks > 1) {
    int rc0 = lanes.make(7);
    if (rc0) return rc0;
    FA_CUDA_TRY(cudaEventRecord(lanes.ev[6], cs));
    FA_CUDA_TRY(cudaStreamWaitEvent(lanes.xs, lanes.ev[6], 0));
  }
  cudaStream_t xs = lanes.xs;
  cudaEvent_t* ev_compute = &lanes.ev[0];
  cudaEvent_t* ev_kv = &lanes.ev[2];
  cudaEvent_t ev_acc_ready = lanes.ev[4], ev_acc_recv = lanes.ev[5];
  const int next = (rank + 1) % nranks, prev = (rank + nranks - 1) % nranks;
  const ncclComm_t comm = static_cast<ncclComm_t>(nccl_comm);
  const void* cur_k = k;
  const void* cur_v = v;
  int rc = FA_OK;
  for (int s = 0; s < nranks && rc == FA_OK; ++s) {
    const int bi = s & 1;
    if (s + 1 < nranks) {        // K/V of the next step: overlapped with this step's compute
      if (s >= 1) FA_CUDA_TRY(cudaStreamWaitEvent(xs, ev_compute[(s - 1) & 1], 0));
      FA_NCCL_GROUP_BEGIN(lanes);
      FA_NCCL_TRY(nc.Send(cur_k, nk * esz, ncclUint8, next, comm, xs));
      FA_NCCL_TRY(nc.Send(cur_v, nv * esz, ncclUint8, next, comm, xs));
      FA_NCCL_TRY(nc.Recv(kbuf[bi], nk * esz, ncclUint8, prev, comm, xs));
      FA_NCCL_TRY(nc.Recv(vbuf[bi], nv * esz, ncclUint8, prev, comm, xs));
      FA_NCCL_GROUP_END(lanes);
      FA_CUDA_TRY(cudaEventRecord(ev_kv[bi], xs));
    }
    // partial gradients of (local queries) x (resident block)
    BwdArgs a{q, cur_k, cur_v, o, d_o, l, m, pq, pk, pv, nullptr, nullptr, nullptr, static_cast<float*>(bws)};
    if (tc) { set_path("tc"); rc = tc_bwd(gd, a, dtype, flags, bws, cs, /*out_f32=*/1); }
    else rc = fa_dense_bwd(q, cur_k, cur_v, o, d_o, l, m, pq, pk, pv, Nl, d, dv, B, dtype, flags, bws, bws_bytes, cs);
    if (rc) break;
    const int pf32 = tc ? 1 : 0;
    if ((rc = accumulate(qacc, pq, nk, pf32, dtype, s == 0, cs))) break;
    // this block's travelling accumulators: fresh at step 0, otherwise they arrived from the previous rank
    if (s >= 1) FA_CUDA_TRY(cudaStreamWaitEvent(cs, ev_acc_recv, 0));
    if ((rc = accumulate(kacc[bi], pk, nk, pf32, dtype, s == 0, cs))) break;
    if ((rc = accumulate(vacc[bi], pv, nv, pf32, dtype, s == 0, cs))) break;
    if (nranks > 1) {
      FA_CUDA_TRY(cudaEventRecord(ev_compute[bi], cs));
      FA_CUDA_TRY(cudaEventRecord(ev_acc_ready, cs));
      FA_CUDA_TRY(cudaStreamWaitEvent(xs, ev_acc_ready, 0));
      FA_NCCL_GROUP_BEGIN(lanes);
      FA_NCCL_TRY(nc.Send(kacc[bi], nk * 4, ncclUint8, next, comm, xs));
      FA_NCCL_TRY(nc.Send(vacc[bi], nv * 4, ncclUint8, next, comm, xs));
      FA_NCCL_TRY(nc.Recv(kacc[bi ^ 1], nk * 4, ncclUint8, prev, comm, xs));
      FA_NCCL_TRY(nc.Recv(vacc[bi ^ 1], nv * 4, ncclUint8, prev, comm, xs));
      FA_NCCL_GROUP_END(lanes);
      FA_CUDA_TRY(cudaEventRecord(ev_acc_recv, xs));
      if (s + 1 < nranks) {
        FA_CUDA_TRY(cudaStreamWaitEvent(cs, ev_kv[bi], 0));
        cur_k = kbuf[bi];
        cur_v = vbuf[bi];
      }
    }
  }
  if (rc == FA_OK) {
    // after the last exchange the accumulators of MY block are in buffer (nranks & 1); one rank: buffer 0
    const int fin = nranks > 1 ? (nranks & 1) : 0;
    if (nranks > 1) FA_CUDA_TRY(cudaStreamWaitEvent(cs, ev_acc_recv, 0));
    if (!(rc = cast_out(dq, qacc, nk, dtype, cs)) && !(rc = cast_out(dk, kacc[fin], nk, dtype, cs)))
      rc = cast_out(dv_out, vacc[fin], nv, dtype, cs);
  }
  lanes.done = rc == FA_OK;      // success: the caller's stream waited on the last exchange event; failure: drain first
  return rc;
}

}  // extern "C"
