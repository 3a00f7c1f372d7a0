/*
 * fa_oracle.c -- C/OpenMP restatement of the reference's CPU flash-attention loops.
 * TEST INFRASTRUCTURE ONLY: used by tests/ (checked against the numpy oracle fa_oracle.py) and as
 * the timed CPU baseline of bench.py (`cpu_baseline`, `--impl reference`).  Never linked into or
 * called from the product library.
 *
 * PARITY PINNING STATUS: pinned against oracle/fa_oracle.py (tests/test_c_oracle.py), which since round 2 is itself
 * pinned to the reference's own src_cpp/FlashAttention.cpp compiled into oracle/_ref (dense forward / backward, 1-D
 * block attention; tests/test_ref_pin.py).  Parity unpinned for the parts no runnable reference code covers here (no
 * Julia runtime): NNlib unfold/fold with padding or overlap, circulant -- see the header of fa_oracle.py.
 *
 * What is restated (citations relative to /root/reference):
 *   fa_oracle_dense_fwd      dense_fa!      src/dense.jl:21-102   same task decomposition
 *                            (@threads over (batch, row-block), :45), same Br/Bc from M=32_000
 *                            (:28-36), same per-tile sequence gemm -> max/exp/sum -> gemm ->
 *                            normalise-every-step update (:77-91).  The two BLAS gemms of the
 *                            reference are plain register-blocked loops here (single-threaded
 *                            inside a task, like BLAS threads = 1 under Julia tasks).
 *   fa_oracle_circulant_fwd  circulant_fa!  src/circulant.jl:9-118 scalar loops incl. the
 *                            cartesian_circulant index (src/utils.jl:6-17) per score.
 *   fa_oracle_windowed_fwd   windowed_fa    src/windowed.jl:3-23  window -> dense_fa ->
 *                            unwindow ./ count with NNlib unfold/fold semantics.
 * Layout: Julia column-major (N, d, B) == C [B][d][N].  Float32.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline long cld(long a, long b) { return (a + b - 1) / b; }
static inline long clampl(long x, long lo, long hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline long pmodl(long a, long n) { long r = a % n; return r < 0 ? r + n : r; }

int fa_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 into every rank; the timed CPU baseline wants all host cores
 * (the reference's Threads.@threads loop, src/dense.jl:45, uses every Julia thread). */
void fa_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* P[i][j] = tau * sum_k Q[k][i] K[k][j];  Q, K are [d][ld] (token contiguous); P is [br][bc] */
static void gemm_qkt(float* restrict P, const float* restrict Q, const float* restrict K, long br, long bc,
                     long d, long ldq, long ldk, float tau) {
  for (long i0 = 0; i0 < br; i0 += 4) {
    const long ni = br - i0 < 4 ? br - i0 : 4;
    for (long j0 = 0; j0 < bc; j0 += 32) {
      const long nj = bc - j0 < 32 ? bc - j0 : 32;
      float acc[4][32];
      memset(acc, 0, sizeof(acc));
      for (long k = 0; k < d; ++k) {
        const float* kr = K + k * ldk + j0;
        float qv[4] = {0, 0, 0, 0};
        for (long a = 0; a < ni; ++a) qv[a] = Q[k * ldq + i0 + a];
        if (nj == 32) {
#pragma omp simd
          for (long j = 0; j < 32; ++j) {
            const float kv = kr[j];
            acc[0][j] += qv[0] * kv; acc[1][j] += qv[1] * kv; acc[2][j] += qv[2] * kv; acc[3][j] += qv[3] * kv;
          }
        } else {
          for (long j = 0; j < nj; ++j) {
            const float kv = kr[j];
            acc[0][j] += qv[0] * kv; acc[1][j] += qv[1] * kv; acc[2][j] += qv[2] * kv; acc[3][j] += qv[3] * kv;
          }
        }
      }
      for (long a = 0; a < ni; ++a)
        for (long j = 0; j < nj; ++j) P[(i0 + a) * bc + j0 + j] = tau * acc[a][j];
    }
  }
}

/* On[i][c] = sum_j P[i][j] V[c][j];  V is [dv][ldv]; On is [br][dv] */
static void gemm_pv(float* restrict On, const float* restrict P, const float* restrict V, long br, long bc,
                    long dv, long ldv) {
  for (long i = 0; i < br; ++i) {
    const float* pr = P + i * bc;
    for (long c = 0; c < dv; ++c) {
      const float* vr = V + c * ldv;
      float s = 0.f;
#pragma omp simd reduction(+ : s)
      for (long j = 0; j < bc; ++j) s += pr[j] * vr[j];
      On[i * dv + c] = s;
    }
  }
}

/* p[c] = exp(p[c] - mij), returns the row sum (src/dense.jl:79-80).  All arguments are finite here,
 * so this one helper may use the vector math library (libmvec) through fast-math codegen. */
static float __attribute__((optimize("-ffast-math"), noinline)) exp_row(float* restrict p, long n, float mij) {
  float s = 0.f;
#pragma omp simd reduction(+ : s)
  for (long c = 0; c < n; ++c) { p[c] = expf(p[c] - mij); s += p[c]; }
  return s;
}

/* dense_fa!(O,l,m,Q,K,V)  src/dense.jl:21-102.  ldn = token stride between channels (N for a
 * plain (N,d,B) array), bstride_* = elements between batch elements. */
static void dense_fwd_core(const float* Q, const float* K, const float* V, float* O, float* l, float* m,
                           long N, long d, long dv, long B) {
  const long M = 32000;                                            /* :28 */
  const long Bc = clampl(cld(M, d), 1, N);                         /* :34 */
  const long Br = clampl(cld(M, d) < d ? cld(M, d) : d, 1, N);     /* :35 */
  const long Tr = cld(N, Br), Tc = cld(N, Bc);                     /* :39-41 */
  const float tau = 1.0f / sqrtf((float)d);                        /* :43 */
#pragma omp parallel
  {
    float* Pij = (float*)malloc(sizeof(float) * Br * Bc);          /* :75 (allocated per tile in the reference) */
    float* On = (float*)malloc(sizeof(float) * Br * dv);           /* Oi_new :62 */
    float* Oi = (float*)malloc(sizeof(float) * Br * dv);
    float* li = (float*)malloc(sizeof(float) * Br);
    float* mi = (float*)malloc(sizeof(float) * Br);
    float* f_old = (float*)malloc(sizeof(float) * Br);
    float* f_new = (float*)malloc(sizeof(float) * Br);
#pragma omp for collapse(2) schedule(dynamic, 1)
    for (long b = 0; b < B; ++b)
      for (long i = 0; i < Tr; ++i) {                              /* :45 */
        const long r0 = i * Br, br = (N - r0 < Br) ? N - r0 : Br;
        for (long x = 0; x < br * dv; ++x) Oi[x] = 0.f;            /* :58 */
        for (long x = 0; x < br; ++x) { li[x] = 0.f; mi[x] = -INFINITY; }   /* :59-60 */
        for (long j = 0; j < Tc; ++j) {                            /* :70 */
          const long c0 = j * Bc, bc = (N - c0 < Bc) ? N - c0 : Bc;
          gemm_qkt(Pij, Q + b * d * N + r0, K + b * d * N + c0, br, bc, d, N, N, tau);   /* :77 */
          for (long r = 0; r < br; ++r) {
            float* p = Pij + r * bc;
            float mij = -INFINITY;
            for (long c = 0; c < bc; ++c) mij = p[c] > mij ? p[c] : mij;               /* :78 */
            const float lij = exp_row(p, bc, mij);                                     /* :79-80 */
            const float mnew = mi[r] > mij ? mi[r] : mij;                              /* :82 */
            const float ei = expf(mi[r] - mnew), eij = expf(mij - mnew);               /* :83-84 */
            const float lnew = ei * li[r] + eij * lij;                                 /* :85 */
            f_old[r] = li[r] * ei / lnew;                                              /* :89 */
            f_new[r] = eij / lnew;
            li[r] = lnew; mi[r] = mnew;                                                /* :90-91 */
          }
          gemm_pv(On, Pij, V + b * dv * N + c0, br, bc, dv, N);                        /* :88 */
          for (long r = 0; r < br; ++r)
            for (long c = 0; c < dv; ++c) Oi[r * dv + c] = f_old[r] * Oi[r * dv + c] + f_new[r] * On[r * dv + c];
        }
        for (long r = 0; r < br; ++r) {
          for (long c = 0; c < dv; ++c) O[(b * dv + c) * N + r0 + r] = Oi[r * dv + c];
          l[b * N + r0 + r] = li[r];
          m[b * N + r0 + r] = mi[r];
        }
      }
    free(Pij); free(On); free(Oi); free(li); free(mi); free(f_old); free(f_new);
  }
}

void fa_oracle_dense_fwd(const float* Q, const float* K, const float* V, float* O, float* l, float* m,
                         long N, long d, long dv, long B) {
  dense_fwd_core(Q, K, V, O, l, m, N, d, dv, B);
}

/* cartesian_circulant(n, N, M)[1] - 1  (src/utils.jl:6-17), n 1-based */
static inline long circ_key(long n, long N, long M) {
  const long p = (M - 1) / 2;
  const long j = cld(n, M);
  long mm = (n - 1) % M + 1;
  if (j <= p) mm = pmodl(mm - 1 - (j - p - 1), M) + 1;
  else if (j > N - p) mm = pmodl(mm - 1 - (p - N + j), M) + 1;
  return pmodl((mm - 1) + (j - 1) - p, N);
}

/* circulant_fa!(O,l,m,Q,K,V,W)  src/circulant.jl:9-118 (scalar loops, index recomputed per use) */
void fa_oracle_circulant_fwd(const float* Q, const float* K, const float* V, float* O, float* l, float* m,
                             long N, long d, long dv, long B, long W) {
  const long M = 32000;
  const long Bw = clampl(cld(M, d), 1, W);                         /* :24 */
  const long Br = clampl(cld(M, d) < d ? cld(M, d) : d, 1, N);     /* :25 */
  const long Tr = cld(N, Br), Tw = cld(W, Bw);
  const float tau = 1.0f / sqrtf((float)d);                        /* :33 */
#pragma omp parallel
  {
    float* Piw = (float*)malloc(sizeof(float) * Br * Bw);
    float* li = (float*)malloc(sizeof(float) * Br);
    float* mi = (float*)malloc(sizeof(float) * Br);
    float* f_old = (float*)malloc(sizeof(float) * Br);
    float* f_new = (float*)malloc(sizeof(float) * Br);
#pragma omp for collapse(2) schedule(dynamic, 1)
    for (long b = 0; b < B; ++b)
      for (long i = 0; i < Tr; ++i) {                              /* :35 */
        const long r0 = i * Br, br = (N - r0 < Br) ? N - r0 : Br;
        const float *Qb = Q + b * d * N, *Kb = K + b * d * N, *Vb = V + b * dv * N;
        float* Ob = O + b * dv * N;
        for (long r = 0; r < br; ++r) {
          li[r] = 0.f; mi[r] = -INFINITY;
          for (long c = 0; c < dv; ++c) Ob[c * N + r0 + r] = 0.f;  /* :50-52 */
        }
        for (long w = 0; w < Tw; ++w) {                            /* :61 */
          const long w0 = w * Bw, bw = (W - w0 < Bw) ? W - w0 : Bw;
          for (long r = 0; r < br; ++r)                            /* :68-79 */
            for (long ww = 0; ww < bw; ++ww) {
              const long jj = circ_key((r0 + r) * W + w0 + ww + 1, N, W);
              float t = 0.f;
              for (long kk = 0; kk < d; ++kk) t += Qb[kk * N + r0 + r] * Kb[kk * N + jj];
              Piw[r * bw + ww] = tau * t;
            }
          for (long r = 0; r < br; ++r) {                          /* :80-87 */
            float* p = Piw + r * bw;
            float miw = -INFINITY;
            for (long c = 0; c < bw; ++c) miw = p[c] > miw ? p[c] : miw;
            float liw = 0.f;
            for (long c = 0; c < bw; ++c) { p[c] = expf(p[c] - miw); liw += p[c]; }
            const float mnew = mi[r] > miw ? mi[r] : miw;
            const float ei = expf(mi[r] - mnew), eiw = expf(miw - mnew);
            const float lnew = ei * li[r] + eiw * liw;
            f_old[r] = li[r] * ei / lnew; f_new[r] = eiw / lnew;
            li[r] = lnew; mi[r] = mnew;
          }
          for (long r = 0; r < br; ++r)                            /* :90-102 */
            for (long c = 0; c < dv; ++c) {
              float t = 0.f;
              for (long ww = 0; ww < bw; ++ww) {
                const long jj = circ_key((r0 + r) * W + w0 + ww + 1, N, W);
                t += Piw[r * bw + ww] * Vb[c * N + jj];
              }
              Ob[c * N + r0 + r] = f_old[r] * Ob[c * N + r0 + r] + f_new[r] * t;
            }
        }
        for (long r = 0; r < br; ++r) { l[b * N + r0 + r] = li[r]; m[b * N + r0 + r] = mi[r]; }
      }
    free(Piw); free(li); free(mi); free(f_old); free(f_new);
  }
}

/* window geometry (NNlib.unfold as used at src/utils.jl:40): token of slot `slot` of window `win` */
static long win_token(int nd, const long* s, const long* o, long W, long stride, long pad, long win, long slot) {
  long tok = 0, mult = 1;
  for (int k = 0; k < nd; ++k) {
    const long wk = win % o[k]; win /= o[k];
    const long kk = slot % W;   slot /= W;
    const long pos = wk * stride - pad + kk;
    if (pos < 0 || pos >= s[k]) return -1;
    tok += pos * mult; mult *= s[k];
  }
  return tok;
}

/* windowed_fa(q,k,v,W;stride,pad)  src/windowed.jl:3-23: window x3 -> dense_fa -> unwindow ./ count.
 * y :: (spatial.., dv, B); l, m :: (W^D, 1, L, B). */
void fa_oracle_windowed_fwd(const float* q, const float* k, const float* v, float* y, float* l, float* m,
                            int nd, const long* s, long d, long dv, long B, long W, long stride, long pad) {
  long o[3] = {1, 1, 1}, N = 1, L = 1, WD = 1;
  for (int i = 0; i < nd; ++i) { o[i] = (s[i] + 2 * pad - W) / stride + 1; N *= s[i]; L *= o[i]; WD *= W; }
  const long LB = L * B;
  float* qw = (float*)calloc((size_t)WD * d * LB, sizeof(float));      /* window(q) :4  (W^D, d, L, B) */
  float* kw = (float*)calloc((size_t)WD * d * LB, sizeof(float));      /* :5 */
  float* vw = (float*)calloc((size_t)WD * dv * LB, sizeof(float));     /* :6 */
  float* yw = (float*)malloc(sizeof(float) * WD * dv * LB);
  long* tok = (long*)malloc(sizeof(long) * WD * L);
  for (long w = 0; w < L; ++w)
    for (long sl = 0; sl < WD; ++sl) tok[w * WD + sl] = win_token(nd, s, o, W, stride, pad, w, sl);
#pragma omp parallel for collapse(2) schedule(static)
  for (long b = 0; b < B; ++b)
    for (long w = 0; w < L; ++w)
      for (long sl = 0; sl < WD; ++sl) {
        const long t = tok[w * WD + sl];
        if (t < 0) continue;
        const long base = (b * L + w);
        for (long c = 0; c < d; ++c) {
          qw[(base * d + c) * WD + sl] = q[(b * d + c) * N + t];
          kw[(base * d + c) * WD + sl] = k[(b * d + c) * N + t];
        }
        for (long c = 0; c < dv; ++c) vw[(base * dv + c) * WD + sl] = v[(b * dv + c) * N + t];
      }
  dense_fwd_core(qw, kw, vw, yw, l, m, WD, d, dv, LB);                  /* :8-11 */
  float* cnt = (float*)calloc((size_t)N, sizeof(float));                /* divisor :16-17 */
  for (long i = 0; i < WD * L; ++i) if (tok[i] >= 0) cnt[tok[i]] += 1.f;
  memset(y, 0, sizeof(float) * N * dv * B);
#pragma omp parallel for schedule(static)
  for (long b = 0; b < B; ++b) {                                        /* unwindow (fold, +=) :19 */
    for (long w = 0; w < L; ++w)
      for (long c = 0; c < dv; ++c)
        for (long sl = 0; sl < WD; ++sl) {
          const long t = tok[w * WD + sl];
          if (t >= 0) y[(b * dv + c) * N + t] += yw[((b * L + w) * dv + c) * WD + sl];
        }
    for (long c = 0; c < dv; ++c)
      for (long t = 0; t < N; ++t) y[(b * dv + c) * N + t] /= cnt[t];   /* ./ divisor; 0/0 = NaN */
  }
  free(qw); free(kw); free(vw); free(yw); free(tok); free(cnt);
}
