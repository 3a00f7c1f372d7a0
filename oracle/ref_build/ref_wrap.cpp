// C entry points around the UNMODIFIED reference source src_cpp/FlashAttention.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The reference file is compiled where it lies (the Makefile passes -I$(REFERENCE), nothing is copied into this
// repo) against the stand-in <eigen3/Eigen/Dense> of this directory; its benchmark main() is renamed by the
// preprocessor so that the translation unit can live in a shared library.  Every function below converts raw
// column-major double buffers -- the memory of a Julia (N, d) slice -- to Eigen matrices, calls the reference
// function with the same argument order as the reference's own main() (src_cpp/FlashAttention.cpp:358-470), and
// copies the results back.  `lambda` is the score scale (the reference defaults it to 1.0; the Julia package
// uses 1/sqrt(d), src/dense.jl:43).
#define main fa_reference_cpp_main
#include "src_cpp/FlashAttention.cpp"
#undef main

namespace {
Eigen::MatrixXd mat(const double* p, long r, long c) {
  Eigen::MatrixXd m(r, c);
  for (long t = 0; t < r * c; ++t) m.data()[t] = p[t];
  return m;
}
Eigen::VectorXd vec(const double* p, long n) {
  Eigen::VectorXd v(n);
  for (long t = 0; t < n; ++t) v(t) = p[t];
  return v;
}
void out(const Eigen::MatrixXd& m, double* p) {
  for (long t = 0; t < m.size(); ++t) p[t] = m.data()[t];
}
}  // namespace

extern "C" {

// OneDNaive, src_cpp/FlashAttention.cpp:15-47 (wsize != 0: independent blocks of wsize tokens, :36-45)
void fa_ref_OneDNaive(const double* Q, const double* K, const double* V, double* O, long N, long d, long wsize, double lambda) {
  Eigen::MatrixXd o(N, d);
  o.setZero();
  OneDNaive(mat(Q, N, d), mat(K, N, d), mat(V, N, d), o, wsize, lambda);
  out(o, O);
}

// OneDFast, :49-100 -- O must start at zero (main() does O.setZero(), :381)
void fa_ref_OneDFast(const double* Q, const double* K, const double* V, double* O, long N, long d, long cache, long wsize, double lambda) {
  Eigen::MatrixXd o(N, d);
  o.setZero();
  OneDFast(mat(Q, N, d), mat(K, N, d), mat(V, N, d), o, cache, wsize, lambda);
  out(o, O);
}

// OneDParallelCPU, :102-159
void fa_ref_OneDParallelCPU(const double* Q, const double* K, const double* V, double* O, long N, long d, long cache, long wsize, double lambda, int threads) {
  if (threads > 0) omp_set_num_threads(threads);
  Eigen::MatrixXd o(N, d);
  o.setZero();
  OneDParallelCPU(mat(Q, N, d), mat(K, N, d), mat(V, N, d), o, cache, wsize, lambda);
  out(o, O);
}

// OneDNaiveBack, :161-192 (dense branch; P is the N x N softmax matrix)
void fa_ref_OneDNaiveBack(const double* Q, const double* K, const double* V, const double* P, const double* dO,
                          double* dQ, double* dK, double* dV, long N, long d, double lambda) {
  Eigen::MatrixXd dq(N, d), dk(N, d), dv(N, d);
  dq.setZero(); dk.setZero(); dv.setZero();
  OneDNaiveBack(mat(Q, N, d), mat(K, N, d), mat(V, N, d), mat(P, N, N), mat(dO, N, d), dq, dk, dv, 0, lambda);
  out(dq, dQ); out(dk, dK); out(dv, dV);
}

// OneDFastBack, :194-252 (dense branch) -- the only runnable statement of dense_fa_backward (src/dense.jl:104-167 is
// broken, SURVEY B-5).  dQ, dK, dV start at zero as in main() (:432).
void fa_ref_OneDFastBack(const double* Q, const double* K, const double* V, const double* O, const double* dO,
                         const double* l, const double* m, double* dQ, double* dK, double* dV,
                         long N, long d, long cache, double lambda) {
  Eigen::MatrixXd dq(N, d), dk(N, d), dv(N, d);
  dq.setZero(); dk.setZero(); dv.setZero();
  OneDFastBack(mat(Q, N, d), mat(K, N, d), mat(V, N, d), mat(O, N, d), mat(dO, N, d), dq, dk, dv, vec(l, N), vec(m, N), cache, 0, lambda);
  out(dq, dQ); out(dk, dK); out(dv, dV);
}

// OneDParallelCPUBack, :253-317.  Its omp-for over query blocks accumulates into shared dK/dV blocks
// (:299-312, a race, SURVEY B-6); threads = 1 makes it deterministic.
void fa_ref_OneDParallelCPUBack(const double* Q, const double* K, const double* V, const double* O, const double* dO,
                                const double* l, const double* m, double* dQ, double* dK, double* dV,
                                long N, long d, long cache, double lambda, int threads) {
  if (threads > 0) omp_set_num_threads(threads);
  Eigen::MatrixXd dq(N, d), dk(N, d), dv(N, d);
  dq.setZero(); dk.setZero(); dv.setZero();
  OneDParallelCPUBack(mat(Q, N, d), mat(K, N, d), mat(V, N, d), mat(O, N, d), mat(dO, N, d), dq, dk, dv, vec(l, N), vec(m, N), cache, 0, lambda);
  out(dq, dQ); out(dk, dK); out(dv, dV);
}

}  // extern "C"
