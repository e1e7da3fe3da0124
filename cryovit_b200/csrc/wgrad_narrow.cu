// Weight gradient of the two 8-channel, full-resolution convolutions of the CryoVIT head (output_layer.0 / .2,
// models/cryovit.py:30-34; head training, BASELINE config 5) straight from the channels-last volumes:
//
//   dW[tap][co][ci] += sum over voxels v of  dZ[v, co] * X[v + off(tap), ci]        (27 taps x 8 x 8 numbers)
//
// 116 GFLOP of work whose result is 1728 numbers: a REDUCTION over 33.5 M voxels. The split-K tcgen05 path
// (csrc/wgrad.cu) needs channels-first zero-padded operand copies (three of X, one of dZ: 2.2 GB written and re-read)
// and, with 8 x 8 outputs, keeps 1/16 of every MMA: 2.6 ms per launch, three launches per layer, plus the copies --
// 23 of the training step's 65 ms. Here the voxels are the K dimension of warp-level mma.sync.m16n8k16 (bf16, fp32
// accumulate) and NOTHING is re-laid-out: a channels-last voxel is 16 bytes = one row of an 8x8 ldmatrix tile, and
// ldmatrix.trans hands out exactly the fragments the MMA wants:
//     A (16 x 16, rows = co, cols = 16 consecutive voxels of a row)  <- trans of the [voxel][co] tile of dZ (rows 8-15 zero)
//     B (16 x 8,  rows = the same voxels shifted by the tap, cols = ci) <- trans of the [voxel][ci] tile of X
// A CTA stages the (8 + 2) x (128 + 2) halo tile of the three depth planes of X and the 8 x 128 tile of dZ with
// cp.async (zero fill = "same" padding), each warp walks one row in 16-voxel blocks (27 ldmatrix + 27 mma per block)
// and keeps all 27 x 8 x 8 partial sums in 54 registers per thread across its tiles; one shared-memory and one global
// red.add pass at the very end.
#include "ptx.cuh"
#include "tmap.h"

namespace cvit {

constexpr int WN_TH = 8, WN_TW = 128, WN_THREADS = 256;
constexpr int WN_XH = WN_TH + 2, WN_XW = WN_TW + 2;
constexpr int WN_X_BYTES = 3 * WN_XH * WN_XW * 16;
constexpr int WN_Z_BYTES = WN_TH * WN_TW * 16;
constexpr int WN_SMEM = WN_X_BYTES + WN_Z_BYTES;

__device__ __forceinline__ void wn_cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void wn_ldmatrix_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void wn_mma_bf16(float& d0, float& d1, float& d2, float& d3, uint32_t a0, uint32_t a1, uint32_t a2,
                                            uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(WN_THREADS, 2)
wgrad_narrow8_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dz, float* __restrict__ dw,
                     int D, int H, int W, int dil, int num_tiles) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t sX = smem_u32(smem_raw), sZ = sX + WN_X_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_w = (W + WN_TW - 1) / WN_TW, tiles_h = (H + WN_TH - 1) / WN_TH;
  const int per_plane = tiles_w * tiles_h;

  float acc[27][2];
#pragma unroll
  for (int t = 0; t < 27; ++t) acc[t][0] = acc[t][1] = 0.f;
  float zero2 = 0.f, zero3 = 0.f;  // rows 8-15 of every product are zero (A rows 8-15 are zero): shared dummies

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int d = tile / per_plane;
    const int rt = tile - d * per_plane;
    const int th = rt / tiles_w;
    const int h0 = th * WN_TH, w0 = (rt - th * tiles_w) * WN_TW;
    __syncthreads();  // the previous tile's fragments have been read
    for (int q = threadIdx.x; q < 3 * WN_XH * WN_XW; q += WN_THREADS) {
      const int kd = q / (WN_XH * WN_XW), r = q - kd * (WN_XH * WN_XW);
      const int hh = r / WN_XW, c = r - hh * WN_XW;
      const int pz = d + (kd - 1) * dil, h = h0 - 1 + hh, w = w0 - 1 + c;
      const bool ok = pz >= 0 && pz < D && h >= 0 && h < H && w >= 0 && w < W;
      const __nv_bfloat16* src = ok ? x + (((int64_t)pz * H + h) * W + w) * 8 : x;
      wn_cp_async16(sX + q * 16, src, ok ? 16u : 0u);
    }
    for (int q = threadIdx.x; q < WN_TH * WN_TW; q += WN_THREADS) {
      const int hl = q / WN_TW, vl = q - hl * WN_TW;
      const int h = h0 + hl, w = w0 + vl;
      const bool ok = h < H && w < W;
      const __nv_bfloat16* src = ok ? dz + (((int64_t)d * H + h) * W + w) * 8 : dz;
      wn_cp_async16(sZ + q * 16, src, ok ? 16u : 0u);
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // warp == tile row; lanes 0-15 supply the 16 row addresses of an ldmatrix.x2 (matrix = lane / 8, row = lane % 8)
    const int lrow = (lane & 15);
    const uint32_t zrow = sZ + (warp * WN_TW + lrow) * 16;
#pragma unroll 1
    for (int b = 0; b < WN_TW / 16; ++b) {
      uint32_t a0, a2;
      wn_ldmatrix_x2_trans(zrow + b * 256, a0, a2);
#pragma unroll
      for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const uint32_t xrow = sX + ((kd * WN_XH + warp + kh) * WN_XW + b * 16 + lrow) * 16;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            uint32_t b0, b1;
            wn_ldmatrix_x2_trans(xrow + kw * 16, b0, b1);
            const int t = (kd * 3 + kh) * 3 + kw;
            wn_mma_bf16(acc[t][0], acc[t][1], zero2, zero3, a0, 0u, a2, 0u, b0, b1);
          }
        }
      }
    }
  }

  // reduce the CTA's 8 warps in shared memory, then one red.add per output from the CTA
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem_raw);
  for (int i = threadIdx.x; i < 27 * 64; i += WN_THREADS) red[i] = 0.f;
  __syncthreads();
  const int co = lane >> 2, ci = (lane & 3) * 2;
#pragma unroll
  for (int t = 0; t < 27; ++t) {
    atomicAdd(&red[t * 64 + co * 8 + ci], acc[t][0]);
    atomicAdd(&red[t * 64 + co * 8 + ci + 1], acc[t][1]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * 64; i += WN_THREADS) atomicAdd(dw + i, red[i]);
}

}  // namespace cvit

using namespace cvit;

// dw (fp32 [27][8][8], tap = (kd*3+kh)*3+kw, [co][ci]) += weight gradient of a 3x3x3 depth-dilated "same" convolution
// with 8 input and 8 output channels; x (the forward input) and dz (the output gradient) are bf16 [D,H,W,8].
extern "C" int cvit_wgrad_narrow8_ndhwc(const void* x, const void* dz, float* dw, int64_t D, int64_t H, int64_t W, int64_t dil,
                                        void* stream) {
  if (!x || !dz || !dw || D <= 0 || H <= 0 || W <= 0 || dil <= 0) {
    set_error("wgrad_narrow8: bad arguments (D=%lld H=%lld W=%lld dil=%lld)", (long long)D, (long long)H, (long long)W,
              (long long)dil);
    return CVIT_ERR_INVALID;
  }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dz)) & 15u) {
    set_error("wgrad_narrow8: x and dz must be 16-byte aligned");
    return CVIT_ERR_INVALID;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_narrow8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WN_SMEM);
    if (e != cudaSuccess) {
      set_error("wgrad_narrow8: cudaFuncSetAttribute(smem=%d): %s", WN_SMEM, cudaGetErrorString(e));
      return CVIT_ERR_CUDA;
    }
    configured = true;
  }
  const int64_t tiles = D * ((H + WN_TH - 1) / WN_TH) * ((W + WN_TW - 1) / WN_TW);
  int grid = 2 * num_sms();
  if (grid > tiles) grid = (int)tiles;
  wgrad_narrow8_kernel<<<grid, WN_THREADS, WN_SMEM, (cudaStream_t)stream>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dz), dw, (int)D, (int)H, (int)W, (int)dil, (int)tiles);
  return check_launch("wgrad_narrow8_kernel");
}
