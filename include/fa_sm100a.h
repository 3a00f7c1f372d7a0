/*
 * fa_sm100a.h -- C ABI of libfa_sm100a.so, the B200 (sm_100a) backend behind the
 * FlashAttention.jl API.
 *
 * The reference has no FFI: its boundary is the exported Julia API
 * (reference src/FlashAttention.jl:13,20-21,26-27).  Each entry point below replaces the
 * host-algorithm body of one exported function; the Julia wrapper keeps the reshape/allocate
 * prologue (src/dense.jl:1-19 etc.) and does one `ccall`.
 *
 * Conventions
 *  - Layout: Julia column-major `(spatial..., d, B)`, i.e. C `[B][d][N]` with the token index
 *    contiguous (src/dense.jl:6-8).  All tensors dense/contiguous, 16-byte aligned.
 *  - Device entry points take raw DEVICE pointers and a `cudaStream_t` (as `void*`; NULL = the
 *    legacy default stream); they enqueue work and return without synchronising.
 *  - `*_host` entry points take HOST pointers (the reference's `Array` arguments); they copy
 *    in, run the same kernels, copy out and synchronise before returning.
 *  - Ownership: the caller owns every buffer (inputs, outputs, workspace).  Nothing is retained.
 *  - Softmax statistics `l`, `m` are always float32 (reference allocates them with eltype(Q),
 *    src/dense.jl:12-13; for 16-bit inputs float32 is kept so backward can reuse them).
 *  - Every function returns an `fa_status`; 0 = ok.  `fa_last_error_string()` returns a
 *    thread-local message.  No exceptions or aborts cross the boundary.
 *  - There is NO CPU fallback: without a CUDA device every compute entry point returns
 *    FA_ERR_CUDA.
 */
#ifndef FA_SM100A_H_
#define FA_SM100A_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FA_VERSION_MAJOR 0
#define FA_VERSION_MINOR 1

typedef enum fa_status {
  FA_OK = 0,
  FA_ERR_INVALID = 1,      /* bad argument (shape, dtype, NULL pointer, W > N, ...) */
  FA_ERR_UNSUPPORTED = 2,  /* valid request that no kernel covers                */
  FA_ERR_CUDA = 3,         /* CUDA runtime/driver error, or no device            */
  FA_ERR_WORKSPACE = 4     /* workspace missing or too small                     */
} fa_status;

typedef enum fa_dtype {
  FA_F32 = 0,   /* Float32: exact FFMA path (no TF32), parity 1e-5 */
  FA_F16 = 1,   /* Float16: tcgen05 kind::f16, fp32 accumulate     */
  FA_BF16 = 2   /* BFloat16: tcgen05 kind::f16, fp32 accumulate    */
} fa_dtype;

/* flags (bit set) */
#define FA_FLAG_NONE 0
#define FA_FLAG_FORCE_SIMT 1   /* never take the tensor-core path (debug / exact arithmetic) */
#define FA_FLAG_BF16_INTERNALS 2 /* backward, bf16 inputs: keep P and dS in bf16 for the MMAs (as FlashAttention-2/3
                                    do) instead of re-encoding the inputs as scaled fp16 first.  Skips one pass over
                                    q,k,v,dO and 4*N*d*B*2 bytes of workspace; gradients then carry the 2^-9 rounding
                                    of bf16 P/dS (max-abs error 2-3e-3 of max instead of < 2e-3). */

#define FA_FLAG_OUT_F32 4        /* 16-bit inputs on the tcgen05 kernels: outputs (o / y, dq, dk, dv) are float32 buffers and
                                    receive the fp32 accumulators unrounded.  Separates the COMPUTE error (north_star:
                                    2e-3) from the 2^-8 / 2^-11 rounding of a bf16 / fp16 result; FA_ERR_UNSUPPORTED when
                                    the shape falls to the exact-fp32 kernels, no-op for Float32 inputs.  Windowed calls
                                    then need the workspace fa_workspace_bytes_windowed_* report for these flags. */

#define FA_FLAG_HOST_NO_REGISTER 8 /* *_host entry points: do not page-lock pageable caller buffers for the call
                                    (cudaHostRegister); the driver's staged, host-blocking copies are used instead */

/* ---- host helpers -------------------------------------------------------------------- */
int fa_version(void);                       /* FA_VERSION_MAJOR*100 + FA_VERSION_MINOR     */
const char* fa_last_error_string(void);     /* thread-local, never NULL                    */
const char* fa_last_path(void);             /* thread-local name of the kernel family the last
                                               compute call dispatched to ("simt", "tc")   */
int fa_device_count(void);                  /* 0 when no CUDA device is usable             */

/* ---- index sets (host, integer, bit-exact) --------------------------------------------- */
/* reference src/utils.jl:6-17 cartesian_circulant: fills keys[W*N] (column-major (W,N)) with
 * the 0-based key index of nz-entry w of query j, in the reference's storage order. */
int fa_circulant_index(int64_t N, int64_t W, int64_t* keys);
/* reference src/utils.jl:36-44 window (NNlib.unfold): fills idx[WD*L] (column-major (W^D,L))
 * with the 0-based linear spatial index read by slot kappa of window w, -1 for zero padding;
 * n_windows[ndim] receives the windows per dim.  Pass idx = NULL to query n_windows only. */
int fa_window_index(int ndim, const int64_t* dims, int64_t W, int64_t stride, int64_t pad,
                    int64_t* n_windows, int64_t* idx);
/* per-position window count (the `divisor` of src/windowed.jl:16-17): cnt[prod(dims)] */
int fa_window_count(int ndim, const int64_t* dims, int64_t W, int64_t stride, int64_t pad,
                    int64_t* cnt);

/* ---- dense: replaces dense_fa! (src/dense.jl:21-102) -------------------------------- */
int fa_dense_fwd(const void* q, const void* k, const void* v, void* o, float* l, float* m,
                 int64_t N, int64_t d, int64_t dv, int64_t B, int dtype, int flags, void* stream);
/* replaces dense_fa_backward (src/dense.jl:104-167, broken) == OneDFastBack
 * (src_cpp/FlashAttention.cpp:194-252) */
size_t fa_workspace_bytes_dense_bwd(int64_t N, int64_t d, int64_t dv, int64_t B, int dtype, int flags);
int fa_dense_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                 const float* l, const float* m, void* dq, void* dk, void* dv_out,
                 int64_t N, int64_t d, int64_t dv, int64_t B, int dtype, int flags,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ---- circulant: replaces circulant_fa! (src/circulant.jl:9-118), 1-D, W <= N ----------- */
int fa_circulant_fwd(const void* q, const void* k, const void* v, void* o, float* l, float* m,
                     int64_t N, int64_t d, int64_t dv, int64_t B, int64_t W,
                     int dtype, int flags, void* stream);
size_t fa_workspace_bytes_circulant_bwd(int64_t N, int64_t d, int64_t dv, int64_t B, int64_t W,
                                        int dtype, int flags);
int fa_circulant_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                     const float* l, const float* m, void* dq, void* dk, void* dv_out,
                     int64_t N, int64_t d, int64_t dv, int64_t B, int64_t W, int dtype, int flags,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- 2-D circulant (periodic neighbourhood) attention: the reference's stated todo (README.md:38-41,53),
 *      the direct product of the 1-D definition.  q,k,v,o :: (X, Y, d|dv, B); l,m :: (X*Y, 1, B);
 *      keys of query (x,y): (mod(x-p+s, X), mod(y-p+t, Y)), s,t = 0..W-1, p = (W-1)/2; W <= min(X,Y), W <= 16.
 *      Exact fp32 math for every dtype (parity-first implementation). */
int fa_circulant2d_index(int64_t X, int64_t Y, int64_t W, int64_t* keys);   /* keys[(y*X+x)*W*W + t*W + s] */
int fa_circulant2d_fwd(const void* q, const void* k, const void* v, void* o, float* l, float* m,
                       int64_t X, int64_t Y, int64_t d, int64_t dv, int64_t B, int64_t W,
                       int dtype, int flags, void* stream);
size_t fa_workspace_bytes_circulant2d_bwd(int64_t X, int64_t Y, int64_t B);
/* workspace that also admits the tcgen05 backward (16-bit, d = dv in {64,128}, X % 64 == 0); with the smaller amount
 * above fa_circulant2d_bwd runs its exact fp32 kernels */
size_t fa_workspace_bytes_circulant2d_bwd_ex(int64_t X, int64_t Y, int64_t d, int64_t dv, int64_t B, int64_t W, int dtype, int flags);
int fa_circulant2d_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                       const float* l, const float* m, void* dq, void* dk, void* dv_out,
                       int64_t X, int64_t Y, int64_t d, int64_t dv, int64_t B, int64_t W, int dtype, int flags,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- windowed: replaces windowed_fa / block_fa (src/windowed.jl:1-23) with window/unwindow
 *      (src/utils.jl:36-54) fused in.  q,k,v,y :: (dims[0..ndim), d|dv, B); l,m :: (W^D,1,L,B). */
size_t fa_workspace_bytes_windowed_fwd(int ndim, const int64_t* dims, int64_t d, int64_t dv, int64_t B,
                                       int64_t W, int64_t stride, int64_t pad, int dtype, int flags);
int fa_windowed_fwd(const void* q, const void* k, const void* v, void* y, float* l, float* m,
                    int ndim, const int64_t* dims, int64_t d, int64_t dv, int64_t B,
                    int64_t W, int64_t stride, int64_t pad, int dtype, int flags,
                    void* workspace, size_t workspace_bytes, void* stream);
size_t fa_workspace_bytes_windowed_bwd(int ndim, const int64_t* dims, int64_t d, int64_t dv, int64_t B,
                                       int64_t W, int64_t stride, int64_t pad, int dtype, int flags);
/* backward per SURVEY A.5.2 (no reference code): dYw = window(dY ./ count), per-window dense
 * backward with P recomputed from (l, m), dq = unwindow(dQw) etc. */
int fa_windowed_bwd(const void* q, const void* k, const void* v, const void* d_y,
                    const float* l, const float* m, void* dq, void* dk, void* dv_out,
                    int ndim, const int64_t* dims, int64_t d, int64_t dv, int64_t B,
                    int64_t W, int64_t stride, int64_t pad, int dtype, int flags,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---- windowed, ONE volume over several GPUs (slab split on window boundaries, no exchange).
 *      Non-overlapping windows only (stride >= W, pad < W).  fa_windowed_slab_plan (host): rank r of nranks takes
 *      the windows [win_lo, win_hi) of the slowest spatial dim and holds the token planes [plane_lo, plane_hi)
 *      of that dim; plan[5] = {plane_lo, plane_hi, win_lo, win_hi, pad_lo}.  The slab calls then see a volume
 *      of extents slab_dims (= dims with the slowest one replaced by plane_hi - plane_lo), `pad_lo` zero planes
 *      in front of it and `nwin` = win_hi - win_lo windows along it; the other dims keep `pad`.  Outputs:
 *      y :: (slab_dims.., dv, B), l,m :: (W^D, 1, L_slab, B) -- the windows of a slab are a contiguous range of
 *      the volume's window index (src/utils.jl:36-44 enumerates windows x-fastest).  Mirrors what a caller of
 *      windowed_fa (src/windowed.jl:3-23) would get on the whole volume, bit for bit. */
int fa_windowed_slab_plan(int ndim, const int64_t* dims, int64_t W, int64_t stride, int64_t pad,
                          int rank, int nranks, int64_t* plan);
int fa_windowed_slab_fwd(const void* q, const void* k, const void* v, void* y, float* l, float* m,
                         int ndim, const int64_t* slab_dims, int64_t d, int64_t dv, int64_t B,
                         int64_t W, int64_t stride, int64_t pad, int64_t pad_lo, int64_t nwin,
                         int dtype, int flags, void* stream);
size_t fa_workspace_bytes_windowed_slab_bwd(int ndim, const int64_t* slab_dims, int64_t d, int64_t dv, int64_t B,
                                            int64_t W, int64_t stride, int64_t pad, int64_t pad_lo, int64_t nwin);
int fa_windowed_slab_bwd(const void* q, const void* k, const void* v, const void* d_y,
                         const float* l, const float* m, void* dq, void* dk, void* dv_out,
                         int ndim, const int64_t* slab_dims, int64_t d, int64_t dv, int64_t B,
                         int64_t W, int64_t stride, int64_t pad, int64_t pad_lo, int64_t nwin,
                         int dtype, int flags, void* workspace, size_t workspace_bytes, void* stream);

/* ---- windowed, ONE volume over several GPUs, OVERLAPPING windows (halo of W - stride planes in, partial-y
 *      halo reduce out).  Window planes of the slowest spatial dim are dealt out by their start plane:
 *      plan[6] = {own_lo, own_hi, ext_hi, win_lo, win_hi, pad_lo}.  Rank r owns planes [own_lo, own_hi) (a tiling of the
 *      volume), computes the windows [win_lo, win_hi), which read planes [own_lo, ext_hi): its own plus ext_hi - own_hi
 *      planes of rank r+1 (the halo the host layer fetches).  fa_windowed_slab_fwd_sums runs the windowed kernels on
 *      that extended slab and returns the FOLD SUMS (float32, (slab_dims.., dv, B), not yet divided) and l, m of its
 *      windows; the sums of the halo planes are sent on to rank r+1 and added there; fa_window_divide then divides the
 *      owned planes by the window count of the WHOLE volume (src/windowed.jl:16-19; 0/0 = NaN for uncovered planes).
 *      Together: exactly windowed_fa on the whole volume.  Forward only. */
int fa_windowed_halo_plan(int ndim, const int64_t* dims, int64_t W, int64_t stride, int64_t pad,
                          int rank, int nranks, int64_t* plan);
int fa_windowed_slab_fwd_sums(const void* q, const void* k, const void* v, float* acc, float* l, float* m,
                              int ndim, const int64_t* slab_dims, int64_t d, int64_t dv, int64_t B,
                              int64_t W, int64_t stride, int64_t pad, int64_t pad_lo, int64_t nwin,
                              int dtype, int flags, void* stream);
int fa_window_divide(const float* acc, void* y, int ndim, const int64_t* dims, int64_t W, int64_t stride, int64_t pad,
                     int64_t plane_lo, int64_t nplanes, int64_t channels, int64_t B, int dtype, void* stream);

/* ---- standalone unfold/fold: window / unwindow (src/utils.jl:36-54) --------------------- */
int fa_window(const void* x, void* xw, int ndim, const int64_t* dims, int64_t d, int64_t B,
              int64_t W, int64_t stride, int64_t pad, int dtype, void* stream);
int fa_unwindow(const void* xw, void* x, int ndim, const int64_t* dims, int64_t d, int64_t B,
                int64_t W, int64_t stride, int64_t pad, int dtype, void* stream);

/* ---- element-wise dtype conversion (round to nearest even): F32 <-> F16 / BF16.  The host layer uses it for the opt-in
 *      "Float32 arrays on the tensor cores" mode (Julia: dense_fa(q, k, v; via = BFloat16); Python: via=torch.bfloat16):
 *      cast q, k, v, run the tcgen05 kernels with FA_FLAG_OUT_F32, i.e. bf16/fp16 compute class (2e-3) for callers
 *      that hold Float32 arrays as bench/compare.jl:8-10 does. */
int fa_cast(const void* in, void* out, int64_t n, int from_dtype, int to_dtype, void* stream);

/* ---- softmax: replaces fused_softmax! (src/fused_softmax.jl:11-39); in :: (M,N,B) column-major,
 *      dim = 1 (columns) or 2 (rows); any other dim -> FA_ERR_INVALID (assertion at :12) */
int fa_softmax(void* out, const void* in, int64_t M, int64_t N, int64_t B, int dim, int dtype,
               void* stream);

/* ---- host-buffer entry points (the reference's Array arguments; H2D + kernels + D2H) ---- */
int fa_dense_fwd_host(const void* q, const void* k, const void* v, void* o, float* l, float* m,
                      int64_t N, int64_t d, int64_t dv, int64_t B, int dtype, int flags, int device);
int fa_circulant_fwd_host(const void* q, const void* k, const void* v, void* o, float* l, float* m,
                          int64_t N, int64_t d, int64_t dv, int64_t B, int64_t W,
                          int dtype, int flags, int device);
int fa_windowed_fwd_host(const void* q, const void* k, const void* v, void* y, float* l, float* m,
                         int ndim, const int64_t* dims, int64_t d, int64_t dv, int64_t B,
                         int64_t W, int64_t stride, int64_t pad, int dtype, int flags, int device);

/* backward with HOST buffers: the reference's dense_fa_backward(Q,K,V,O,dO,l,m) takes Arrays (src/dense.jl:104-111).
 * Same batch-chunked three-stream pipeline; the workspace is kept inside the library per device. */
int fa_dense_bwd_host(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                      const float* l, const float* m, void* dq, void* dk, void* dv_out,
                      int64_t N, int64_t d, int64_t dv, int64_t B, int dtype, int flags, int device);
int fa_circulant_bwd_host(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                          const float* l, const float* m, void* dq, void* dk, void* dv_out,
                          int64_t N, int64_t d, int64_t dv, int64_t B, int64_t W, int dtype, int flags, int device);
int fa_windowed_bwd_host(const void* q, const void* k, const void* v, const void* d_y,
                         const float* l, const float* m, void* dq, void* dk, void* dv_out,
                         int ndim, const int64_t* dims, int64_t d, int64_t dv, int64_t B,
                         int64_t W, int64_t stride, int64_t pad, int dtype, int flags, int device);
/* The *_host calls run on `device` and restore the caller's current device before returning.  Pageable caller
 * buffers are page-locked for the duration of a multi-chunk call (cudaHostRegister; FA_FLAG_HOST_NO_REGISTER skips
 * that); buffers from fa_host_alloc -- page-locked AND placed on the NUMA node of `device` -- avoid both the
 * registration cost and inter-socket traffic. */
int fa_host_alloc(void** out, size_t bytes, int device);
int fa_host_free(void* p);

/* ---- multi-GPU ------------------------------------------------------------------
 * The path shards over the trailing batch*head dim (contiguous in memory, src/dense.jl:6-8,45) with no
 * collective: fa_shard_batch gives rank `rank` of `nranks` its [begin, begin+count) range. */
int fa_shard_batch(int64_t B, int nranks, int rank, int64_t* begin, int64_t* count);
/* Online-softmax merge of a block partial into fp32 accumulators, the update rule of
 * src/dense.jl:82-91: (o_acc, l_acc, m_acc) <- merge with (o_blk [dtype], l_blk, m_blk); o_acc stays
 * normalised.  first != 0 ignores the accumulators' previous content.  out (optional) receives o_acc
 * in `dtype`.  All (N, dv|1, B) column-major. */
int fa_merge_partials(float* o_acc, float* l_acc, float* m_acc, const void* o_blk, const float* l_blk,
                      const float* m_blk, void* out, int64_t N, int64_t dv, int64_t B, int dtype, int first,
                      void* stream);
/* Ring attention forward for ONE long sequence sharded by tokens: rank r holds tokens
 * [r*Nl, (r+1)*Nl) of q, k, v ((Nl, d|dv, B) each) and receives its (Nl, dv, B) slice of
 * dense_fa(q, k, v) over the full N = nranks*Nl sequence, plus l, m.  nranks steps of
 * [dense forward against the resident K/V block, merge] with the blocks passed round the ring by
 * ncclSend/ncclRecv on a second stream (overlapped).  nccl_comm is the caller's ncclComm_t (e.g.
 * torch's ProcessGroupNCCL communicator); nranks == 1 needs none. */
size_t fa_workspace_bytes_ring_dense_fwd(int64_t Nl, int64_t d, int64_t dv, int64_t B, int dtype);
int fa_ring_dense_fwd(const void* q, const void* k, const void* v, void* o, float* l, float* m,
                      int64_t Nl, int64_t d, int64_t dv, int64_t B, int dtype, int flags,
                      void* nccl_comm, int rank, int nranks, void* workspace, size_t workspace_bytes,
                      void* stream);

/* Ring attention backward: rank r passes its shards of q, k, v, o, dO and the GLOBAL l, m of its queries
 * (outputs of fa_ring_dense_fwd) and receives dq, dk, dv of its shard.  Every step runs the flash backward
 * of the local queries against the resident K/V block; the block's fp32 dK/dV accumulators travel round
 * the ring with it (ncclSend/ncclRecv), the K/V exchange of the next step overlaps the compute. */
size_t fa_workspace_bytes_ring_dense_bwd(int64_t Nl, int64_t d, int64_t dv, int64_t B, int dtype, int flags);
int fa_ring_dense_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                      const float* l, const float* m, void* dq, void* dk, void* dv_out,
                      int64_t Nl, int64_t d, int64_t dv, int64_t B, int dtype, int flags,
                      void* nccl_comm, int rank, int nranks, void* workspace, size_t workspace_bytes,
                      void* stream);

/* The *_host entry points keep their device staging buffers between calls (grow-only, per device);
 * this frees them.  Returns FA_OK. */
int fa_release_host_staging(void);

#ifdef __cplusplus
}
#endif
#endif /* FA_SM100A_H_ */
