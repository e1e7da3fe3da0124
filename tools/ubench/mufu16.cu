// Micro-benchmark: is ex2.approx.f16x2 (two MUFU.EX2.F16 in SASS) faster per element than ex2.approx.ftz.f32?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, int iters) {
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0xb800b400u + threadIdx.x + i;  // small negative halves / some float
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[i]));
      else if (MODE == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(*reinterpret_cast<float*>(&v[i])));
      else asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(v[i]));
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  uint32_t* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 400;
  const char* names[3] = {"ex2.approx.f16x2 (2 elements/instr)", "ex2.approx.ftz.f32", "ex2.approx.ftz.bf16x2 (2 elements/instr)"};
  for (int mode = 0; mode < 3; ++mode)
    for (int warps = 4; warps <= 8; warps *= 2) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 1) k<1><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 2) k<2><<<148, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-44s warps/scheduler %d: %.2f cycles per PTX instruction per scheduler\n", names[mode], warps / 4, (double)h / iters / 32.0 / (warps / 4));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
