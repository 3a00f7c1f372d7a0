// fa_probe_api.cu -- extern "C" entry points of the hardware probes in fa_tc_probe.cu.  NOT part of the product
// library: built only into lib/libfa_sm100a_probe.so (`make probe`) and the -DFA_TRACE library (`make trace`), which
// tests/test_gpu_probe.py and tools/probe_*.py load explicitly.  libfa_sm100a.so exports exactly include/fa_sm100a.h.
#include "fa_common.cuh"

namespace fa {
int tc_probe(int mode, const void* a, const void* b, const float* p, float* out, int D, int dtype,
             int lbo, int sbo, int kstep, int kbox, int afmt, cudaStream_t st);
int tmem_bw_probe(int mode, int nwarps, int iters, long long* out_dev, cudaStream_t st);
int umma_rate_probe(int mode, int n_cols, int iters, int blocks, long long* out_dev, cudaStream_t st);
int tma5d_probe(const void* base, const long long* dims, const long long* strides_bytes, const int* box, const int* coord,
                unsigned char* out_dev, cudaStream_t st);
}  // namespace fa
using namespace fa;

static int need_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); set_error("no CUDA device available"); return FA_ERR_CUDA; }
  return FA_OK;
}

extern "C" {

// ------------------------------------------------------------------------------ diagnostics
// One UMMA tile with caller-supplied descriptor fields (see fa_tc_probe.cu).  Not a product API.
int fa_debug_umma_probe(int mode, const void* a, const void* b, const float* p, float* out, int D, int dtype,
                        int lbo, int sbo, int kstep, int kbox, int afmt, void* stream) {
  int rc = need_device();
  if (rc) return rc;
  return tc_probe(mode, a, b, p, out, D, dtype, lbo, sbo, kstep, kbox, afmt, static_cast<cudaStream_t>(stream));
}

// TMEM read/write bandwidth microbenchmark (see fa_tc_probe.cu).  Not a product API.
int fa_debug_tmem_bw(int mode, int nwarps, int iters, long long* out_dev, void* stream) {
  int rc = need_device();
  if (rc) return rc;
  if ((nwarps != 1 && nwarps != 4 && nwarps != 8) || iters <= 0 || !out_dev) { set_error("bad probe arguments"); return FA_ERR_INVALID; }
  return tmem_bw_probe(mode, nwarps, iters, out_dev, static_cast<cudaStream_t>(stream));
}

// tcgen05.mma throughput by operand source and N (see fa_tc_probe.cu).  Not a product API.
int fa_debug_umma_rate(int mode, int n_cols, int iters, int blocks, long long* out_dev, void* stream) {
  int rc = need_device();
  if (rc) return rc;
  if (mode < 0 || mode > 7 || n_cols < 16 || n_cols > 256 || n_cols % 16 || iters <= 0 || blocks <= 0 || !out_dev) { set_error("bad probe arguments"); return FA_ERR_INVALID; }
  return umma_rate_probe(mode, n_cols, iters, blocks, out_dev, static_cast<cudaStream_t>(stream));
}

// One 5-D TMA box load (bf16 elements) copied back out (see fa_tc_probe.cu).  Not a product API.
int fa_debug_tma5d(const void* base, const long long* dims, const long long* strides_bytes, const int* box, const int* coord,
                   unsigned char* out_dev, void* stream) {
  int rc = need_device();
  if (rc) return rc;
  return tma5d_probe(base, dims, strides_bytes, box, coord, out_dev, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
