// tma_reduce_probe.cu -- which 5-D TMA stores / reductions does sm_100a accept for a 16-bit tensor?
// usage: tma_reduce_probe MODE BX CH X0 [DT [Y0]]   MODE 0 = plain store, 1 = reduce add; DT 0 = bf16, 1 = f16, 2 = f32
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE>
__global__ void k(const __grid_constant__ CUtensorMap tm, int bytes, int x0, int esz, int dt, int y0) {
  extern __shared__ __align__(1024) uint8_t sm[];
  for (int i = threadIdx.x; i < bytes / esz; i += blockDim.x) {
    if (dt == 0) reinterpret_cast<__nv_bfloat16*>(sm)[i] = __float2bfloat16(1.f);
    else if (dt == 1) reinterpret_cast<__half*>(sm)[i] = __float2half(1.f);
    else reinterpret_cast<float*>(sm)[i] = 1.f;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(sm);
    if (MODE == 0)
      asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                   ::"l"(reinterpret_cast<uint64_t>(&tm)), "r"(src), "r"(x0), "r"(y0), "r"(y0), "r"(0), "r"(0) : "memory");
    else
      asm volatile("cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                   ::"l"(reinterpret_cast<uint64_t>(&tm)), "r"(src), "r"(x0), "r"(y0), "r"(y0), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main(int argc, char** argv) {
  const int mode = atoi(argv[1]), BX = atoi(argv[2]), CH = atoi(argv[3]), x0 = atoi(argv[4]), dt = argc > 5 ? atoi(argv[5]) : 0, y0 = argc > 6 ? atoi(argv[6]) : 0;
  const int esz = dt == 2 ? 4 : 2;
  const int X = 64, Y = 8, Z = 8, C = 64, B = 2;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeFn fn = (EncodeFn)p;
  void* y;
  const size_t n = (size_t)X * Y * Z * C * B;
  cudaMalloc(&y, n * esz);
  cudaMemset(y, 0, n * esz);
  CUtensorMap tm;
  const cuuint64_t dims[5] = {X, Y, Z, C, B};
  const cuuint64_t strides[4] = {(cuuint64_t)X * esz, (cuuint64_t)X * Y * esz, (cuuint64_t)X * Y * Z * esz, (cuuint64_t)X * Y * Z * C * esz};
  const cuuint32_t box[5] = {(cuuint32_t)BX, 5, 5, (cuuint32_t)CH, 1};
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  const CUtensorMapDataType cdt = dt == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : dt == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(&tm, cdt, 5, y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("mode %d BX %d CH %d x0 %d dt %d: encode failed %d\n", mode, BX, CH, x0, dt, (int)r); return 1; }
  const int bytes = BX * 25 * CH * esz;
  cudaError_t e;
  if (mode == 0) { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); k<0><<<1, 256, bytes>>>(tm, bytes, x0, esz, dt, y0); }
  else { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); k<1><<<2, 256, bytes>>>(tm, bytes, x0, esz, dt, y0); }
  e = cudaDeviceSynchronize();
  double sum = 0;
  if (e == cudaSuccess) {
    uint8_t* h = (uint8_t*)malloc(n * esz);
    cudaMemcpy(h, y, n * esz, cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < n; ++i)
      sum += dt == 0 ? __bfloat162float(reinterpret_cast<__nv_bfloat16*>(h)[i]) : dt == 1 ? __half2float(reinterpret_cast<__half*>(h)[i]) : reinterpret_cast<float*>(h)[i];
  }
  const int xin = (x0 + BX > X ? X : x0 + BX) - (x0 < 0 ? 0 : x0);
  const int yin = (y0 + 5 > Y ? Y : y0 + 5) - (y0 < 0 ? 0 : y0);
  printf("mode %d BX %d CH %d x0 %d dt %d y0=z0 %d (%d B box): %s, sum %.0f (expect %d)\n", mode, BX, CH, x0, dt, y0, bytes, cudaGetErrorString(e), sum,
         xin * yin * yin * CH * (mode == 1 ? 2 : 1));
  return 0;
}
