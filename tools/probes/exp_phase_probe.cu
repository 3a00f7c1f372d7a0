// exp_phase_probe.cu -- how fast can ONE warp per scheduler (and two) run the exp phase of a softmax step?
// Standalone micro-benchmark (nvcc -arch=sm_100a); prints clocks per 64-element row step for variants of the
// inner loop of tc_fwd_kernel (fa_tc_fwd.cu): knock-outs tell which instruction class paces a lone warp.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_cvt(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h);
}
// round-half-up on the magnitude, integer pipe: (bits + 0x8000) >> 16, two of them merged by one PRMT
__device__ __forceinline__ uint32_t pack_int(float a, float b) {
  const uint32_t ua = __float_as_uint(a) + 0x8000u, ub = __float_as_uint(b) + 0x8000u;
  return __byte_perm(ua, ub, 0x7632);
}
__device__ __forceinline__ uint32_t pack_trunc(float a, float b) { return __byte_perm(__float_as_uint(a), __float_as_uint(b), 0x7632); }

// MODE bits: 1 = no pack, 2 = integer pack, 4 = truncating pack, 8 = no row sum, 16 = no ffma (x = raw), 32 = no ex2
// BLK = elements per block (all ex2 of a block issued before their consumers)
template <int MODE, int BLK>
__global__ void __launch_bounds__(256) probe(long long* clk, float* sink, int iters) {
  __shared__ uint4 sm[256 * 2];
  float s[64];
#pragma unroll
  for (int e = 0; e < 64; ++e) s[e] = -0.01f * (float)((threadIdx.x * 7 + e * 13) & 63);
  float2 l2a = make_float2(0.f, 0.f), l2b = make_float2(0.f, 0.f);
  const float2 scale2 = make_float2(0.125f, 0.125f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float nm = -1e-3f * (float)it;
    const float2 negm2 = make_float2(nm, nm);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t pk[16];
#pragma unroll
      for (int e0 = 0; e0 < 32; e0 += BLK) {
        float2 x[BLK / 2], p[BLK / 2];
#pragma unroll
        for (int u = 0; u < BLK / 2; ++u) {
          const float2 raw = make_float2(s[32 * c + e0 + 2 * u], s[32 * c + e0 + 2 * u + 1]);
          x[u] = (MODE & 16) ? make_float2(raw.x + nm, raw.y + nm) : __ffma2_rn(raw, scale2, negm2);
        }
#pragma unroll
        for (int u = 0; u < BLK / 2; ++u) {
          if (MODE & 32) p[u] = x[u]; else { p[u].x = ex2(x[u].x); p[u].y = ex2(x[u].y); }
        }
#pragma unroll
        for (int u = 0; u < BLK / 2; ++u) {
          if (!(MODE & 8)) { if (u & 1) l2b = __fadd2_rn(l2b, p[u]); else l2a = __fadd2_rn(l2a, p[u]); }
          if (MODE & 1) pk[(e0 >> 1) + u] = __float_as_uint(p[u].x) ^ __float_as_uint(p[u].y);
          else if (MODE & 2) pk[(e0 >> 1) + u] = pack_int(p[u].x, p[u].y);
          else if (MODE & 4) pk[(e0 >> 1) + u] = pack_trunc(p[u].x, p[u].y);
          else pk[(e0 >> 1) + u] = pack_cvt(p[u].x, p[u].y);
        }
      }
#pragma unroll
      for (int v = 0; v < 4; ++v)
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"((uint32_t)__cvta_generic_to_shared(&sm[threadIdx.x * 2 + (v & 1)])),
                     "r"(pk[4 * v]), "r"(pk[4 * v + 1]), "r"(pk[4 * v + 2]), "r"(pk[4 * v + 3]) : "memory");
    }
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) clk[blockIdx.x * 8 + (threadIdx.x >> 5)] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = l2a.x + l2a.y + l2b.x + l2b.y;
}


// Software-pipelined step: the consumers (row sum, pack) of block b-1 and the scale/shift FFMA2 of block b+1 are
// placed after the ex2 of block b, and the row max of the NEXT step's S row (sn) rides in the second half.
// PMODE bits: 1 = also reduce the max of sn, 2 = integer pack
template <int PMODE, int BLK>
__global__ void __launch_bounds__(256) probe_swp(long long* clk, float* sink, int iters) {
  __shared__ uint4 sm[256 * 2];
  float s[64], sn[64];
#pragma unroll
  for (int e = 0; e < 64; ++e) { s[e] = -0.01f * (float)((threadIdx.x * 7 + e * 13) & 63); sn[e] = s[e] * 1.5f + 0.25f; }
  float2 l2a = make_float2(0.f, 0.f), l2b = make_float2(0.f, 0.f);
  const float2 scale2 = make_float2(0.125f, 0.125f);
  float macc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float nm = -1e-3f * (float)it;
    const float2 negm2 = make_float2(nm, nm);
    constexpr int NB = 64 / BLK, H = BLK / 2;
    uint32_t pk[32];
    float2 x[2][H], p[2][H];
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int u = 0; u < H; ++u) x[0][u] = __ffma2_rn(make_float2(s[2 * u], s[2 * u + 1]), scale2, negm2);
#pragma unroll
    for (int b = 0; b <= NB; ++b) {
      if (b < NB) {
#pragma unroll
        for (int u = 0; u < H; ++u) { p[b & 1][u].x = ex2(x[b & 1][u].x); p[b & 1][u].y = ex2(x[b & 1][u].y); }
      }
      if (b + 1 < NB) {
#pragma unroll
        for (int u = 0; u < H; ++u)
          x[(b + 1) & 1][u] = __ffma2_rn(make_float2(s[BLK * (b + 1) + 2 * u], s[BLK * (b + 1) + 2 * u + 1]), scale2, negm2);
      }
      if (b > 0) {
        const int bb = b - 1;
#pragma unroll
        for (int u = 0; u < H; ++u) {
          if (u & 1) l2b = __fadd2_rn(l2b, p[bb & 1][u]); else l2a = __fadd2_rn(l2a, p[bb & 1][u]);
          pk[bb * H + u] = (PMODE & 2) ? pack_int(p[bb & 1][u].x, p[bb & 1][u].y) : pack_cvt(p[bb & 1][u].x, p[bb & 1][u].y);
        }
        if ((bb * BLK) % 32 == 32 - BLK) {       // a 32-column half is complete: store it
          const int c = (bb * BLK) / 32;
#pragma unroll
          for (int v = 0; v < 4; ++v)
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"((uint32_t)__cvta_generic_to_shared(&sm[threadIdx.x * 2 + (v & 1)])),
                         "r"(pk[16 * c + 4 * v]), "r"(pk[16 * c + 4 * v + 1]), "r"(pk[16 * c + 4 * v + 2]), "r"(pk[16 * c + 4 * v + 3]) : "memory");
        }
      }
      if ((PMODE & 1) && b >= NB / 2 && b < NB) {     // max of the next row in the second half
        const int per = 64 / (NB / 2), o = (b - NB / 2) * per;
#pragma unroll
        for (int e = 0; e < per; e += 4) {
          mx0 = fmaxf(mx0, fmaxf(sn[o + e] + nm, sn[o + e + 1]));
          mx1 = fmaxf(mx1, fmaxf(sn[o + e + 2], sn[o + e + 3]));
        }
      }
    }
    macc += fmaxf(mx0, mx1);
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) clk[blockIdx.x * 8 + (threadIdx.x >> 5)] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = l2a.x + l2a.y + l2b.x + l2b.y + macc;
}

template <int PMODE, int BLK>
void run_swp(const char* name, long long* d_clk, float* d_sink) {
  const int iters = 2000;
  for (int threads : {128, 256}) {
    probe_swp<PMODE, BLK><<<148, threads>>>(d_clk, d_sink, iters);
    probe_swp<PMODE, BLK><<<148, threads>>>(d_clk, d_sink, iters);
    cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-34s blk=%2d warps/scheduler=%d  clk per 64-element step: %.1f\n", name, BLK, threads / 128, (double)h[0] / iters);
  }
}

template <int MODE, int BLK>
void run(const char* name, long long* d_clk, float* d_sink) {
  const int iters = 2000;
  for (int threads : {128, 256}) {
    probe<MODE, BLK><<<148, threads>>>(d_clk, d_sink, iters);
    probe<MODE, BLK><<<148, threads>>>(d_clk, d_sink, iters);
    cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-34s blk=%2d warps/scheduler=%d  clk per 64-element step: %.1f\n", name, BLK, threads / 128, (double)h[0] / iters);
  }
}

int main() {
  long long* d_clk; float* d_sink;
  cudaMalloc(&d_clk, 148 * 8 * sizeof(long long));
  cudaMalloc(&d_sink, 148 * 256 * sizeof(float));
  run<0, 8>("as in kernel (cvt pack)", d_clk, d_sink);
  run<0, 16>("as in kernel (cvt pack)", d_clk, d_sink);
  run<0, 2>("pairwise", d_clk, d_sink);
  run<2, 8>("integer round + prmt pack", d_clk, d_sink);
  run<2, 16>("integer round + prmt pack", d_clk, d_sink);
  run<4, 8>("truncating prmt pack", d_clk, d_sink);
  run<1, 8>("no pack (xor)", d_clk, d_sink);
  run<8, 8>("no row sum", d_clk, d_sink);
  run<8 | 2, 8>("no row sum, integer pack", d_clk, d_sink);
  run<8 | 1, 8>("no row sum, no pack", d_clk, d_sink);
  run<16, 8>("no ffma", d_clk, d_sink);
  run<16 | 8 | 1, 8>("ex2 only", d_clk, d_sink);
  run<32, 8>("no ex2 (ffma + sum + cvt pack)", d_clk, d_sink);
  run<32 | 2, 8>("no ex2 (ffma + sum + int pack)", d_clk, d_sink);
  run_swp<0, 8>("swp", d_clk, d_sink);
  run_swp<0, 4>("swp", d_clk, d_sink);
  run_swp<0, 2>("swp", d_clk, d_sink);
  run_swp<0, 16>("swp", d_clk, d_sink);
  run_swp<1, 8>("swp + next-row max", d_clk, d_sink);
  run_swp<1, 4>("swp + next-row max", d_clk, d_sink);
  run_swp<1, 16>("swp + next-row max", d_clk, d_sink);
  run_swp<2, 8>("swp, integer pack", d_clk, d_sink);
  run_swp<3, 8>("swp, integer pack + next-row max", d_clk, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
