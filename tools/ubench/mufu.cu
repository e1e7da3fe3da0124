// Micro-benchmark: MUFU.EX2 / softmax inner-loop throughput per scheduler on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint32_t packbf(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = -0.001f * (threadIdx.x + i);
  float acc = 0.f;
  uint64_t c2 = pk(0.18f, 0.18f), n2 = pk(-0.3f, -0.3f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // pure MUFU
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = ex2(v[i]);
    } else if (MODE == 1) {  // softmax inner loop: ffma2 + 2 ex2 + add2 + pack
      uint64_t ls0 = 0, ls1 = 0;
      uint32_t p[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        uint64_t t2 = fma2(pk(v[2 * i], v[2 * i + 1]), c2, n2);
        float a, b; upk(t2, a, b);
        float p0 = ex2(a), p1 = ex2(b);
        if (i & 1) ls1 = add2(ls1, pk(p0, p1)); else ls0 = add2(ls0, pk(p0, p1));
        p[i] = packbf(p0, p1);
      }
      float a, b; upk(add2(ls0, ls1), a, b); acc += a + b;
#pragma unroll
      for (int i = 0; i < 32; ++i) { v[2 * i] = __uint_as_float(p[i]) * 1e-30f - 0.5f; v[2 * i + 1] -= 0.25f; }
    } else if (MODE == 2) {  // polynomial exp2 on the FMA pipe (Cody-Waite + degree-3), packed
      uint32_t p[32];
      const uint64_t magic = pk(12582912.f, 12582912.f), nmagic = pk(-12582912.f, -12582912.f);
      const uint64_t c3 = pk(0.0555054f, 0.0555054f), c2p = pk(0.2402265f, 0.2402265f), c1 = pk(0.6931472f, 0.6931472f), one = pk(1.f, 1.f);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        uint64_t x = fma2(pk(v[2 * i], v[2 * i + 1]), c2, n2);
        uint64_t xf = add2(x, magic);
        uint64_t xr = add2(xf, nmagic);
        uint64_t nxr; { float a, b; upk(xr, a, b); nxr = pk(-a, -b); }
        uint64_t f = add2(x, nxr);
        uint64_t q = fma2(c3, f, c2p);
        q = fma2(q, f, c1);
        q = fma2(q, f, one);
        float q0, q1, e0, e1; upk(q, q0, q1); upk(xf, e0, e1);
        float p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(e0) << 23));
        float p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(e1) << 23));
        p[i] = packbf(p0, p1);
        acc += p0 + p1;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) { v[2 * i] = __uint_as_float(p[i]) * 1e-30f - 0.5f; v[2 * i + 1] -= 0.25f; }
    }
  }
  long long t1 = clock64();
  float s = acc;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 200;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps = 4; warps <= 16; warps *= 2) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 1) k<1><<<148, warps * 32>>>(out, cyc, iters);
        if (mode == 2) k<2><<<148, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      double per_elem = (double)h / iters / 64.0;  // cycles per element per warp
      printf("mode %d warps/SM %2d (per scheduler %d): %.2f cycles per element per warp, %.2f per scheduler-element\n", mode, warps, warps / 4, per_elem, per_elem / (warps / 4));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
