// Weight gradients of the NARROW 3x3x3 convolutions of the CryoVIT head (8 / 16 / 32 channels: SynthesisBlocks 3-4 and
// output_layer, models/cryovit.py:26-34; head training, BASELINE config 5) straight from the channels-last volumes:
//
//   dW[tap][co][ci] += sum over voxels v of  dZ[v, co] * X[v + off(tap), ci]        (27 x Cout x Cin numbers)
//
// A few GFLOP whose result is a few thousand numbers: a REDUCTION over up to 33.5 M voxels. The split-K tcgen05 path
// (csrc/wgrad.cu) needs channels-first zero-padded operand copies (three of X, one of dZ, written and re-read) and keeps
// a small fraction of every 128-row MMA: 1.5-2.6 ms per launch, three launches per layer, plus the copies -- a third of
// the training step. Here the voxels are the K dimension of warp-level mma.sync.m16n8k16 (bf16, fp32 accumulate) and
// NOTHING is re-laid-out: 8 channels of a channels-last voxel are 16 bytes = one row of an 8x8 ldmatrix tile, and
// ldmatrix.trans hands out exactly the fragments the MMA wants:
//     A (16 x 16, rows = co, cols = 16 consecutive voxels of a row)  <- trans of [voxel][co] tiles of dZ
//     B (16 x 8,  rows = the same voxels shifted by the tap, cols = ci) <- trans of [voxel][ci] tiles of X
// A CTA stages the (8 + 2) x (TW + 2) halo tile of the three depth planes of X and the 8 x TW tile of dZ with cp.async
// (zero fill = "same" padding; 16-byte chunks XOR-swizzled by voxel so that the 8 rows of every ldmatrix tile fall in
// distinct banks). Nine warps, one per (kd, kh): a warp walks the whole tile in 16-voxel blocks and owns the three kw
// taps of its (kd, kh) for all Cout x Cin -- 3 x Cout/16 x Cin/8 accumulator fragments that stay in registers across
// all of the CTA's tiles; one red.global.add pass at the very end.
#include "ptx.cuh"
#include "tmap.h"

namespace cvit {

constexpr int WN_TH = 8, WN_XH = WN_TH + 2, WN_THREADS = 288;

template <int CIN, int COUT, int TW>
struct WnCfg {
  static constexpr int XW = TW + 2;
  static constexpr int NCX = CIN / 8, NCZ = COUT / 8;           // 16-byte chunks per voxel
  static constexpr int X_BYTES = 3 * WN_XH * XW * CIN * 2;
  static constexpr int Z_BYTES = WN_TH * TW * COUT * 2;
  static constexpr int SMEM = X_BYTES + Z_BYTES;
  static constexpr int MB = (COUT + 15) / 16, NB = CIN / 8;       // MMA row blocks (co), column blocks (ci)
  static_assert(SMEM <= 232448 && TW % 16 == 0, "tile does not fit");
};

// chunk j of the voxel with in-row index v sits at chunk position j ^ swz(v): 8 consecutive voxels (any start) then
// cover all eight 16-byte bank groups of a 128-byte line for a fixed j
template <int NCH>
__device__ __forceinline__ int wn_swz(int v) {
  return NCH == 4 ? (v >> 1) & 3 : NCH == 2 ? (v >> 2) & 1 : 0;
}
__device__ __forceinline__ void wn_cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void wn_ldmatrix_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void wn_ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void wn_mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                            uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int CIN, int COUT, int TW>
__global__ void __launch_bounds__(WN_THREADS, 1)
wgrad_narrow_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dz, float* __restrict__ dw, int D,
                    int H, int W, int dil, int num_tiles) {
  using Cfg = WnCfg<CIN, COUT, TW>;
  constexpr int XW = Cfg::XW, NCX = Cfg::NCX, NCZ = Cfg::NCZ, MB = Cfg::MB, NB = Cfg::NB;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t sX = smem_u32(smem_raw), sZ = sX + Cfg::X_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kd = warp / 3, kh = warp - 3 * kd;  // this warp's taps: (kd, kh, kw = 0..2)
  const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + WN_TH - 1) / WN_TH;
  const int per_plane = tiles_w * tiles_h;

  float acc[3][MB][NB][4];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < MB; ++b)
#pragma unroll
      for (int c = 0; c < NB; ++c)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[a][b][c][e] = 0.f;

  // ldmatrix row addressing of this lane: matrix m = lane / 8, row r = lane % 8
  const int lm = lane >> 3, lr = lane & 7;

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int d = tile / per_plane;
    const int rt = tile - d * per_plane;
    const int th = rt / tiles_w;
    const int h0 = th * WN_TH, w0 = (rt - th * tiles_w) * TW;
    __syncthreads();  // the previous tile's fragments have been read
    for (int q = threadIdx.x; q < 3 * WN_XH * XW * NCX; q += WN_THREADS) {
      const int j = q % NCX, vq = q / NCX;  // chunk fastest: a voxel's channels are contiguous in global memory
      const int pk = vq / (WN_XH * XW), r = vq - pk * (WN_XH * XW);
      const int hh = r / XW, c = r - hh * XW;
      const int pz = d + (pk - 1) * dil, h = h0 - 1 + hh, w = w0 - 1 + c;
      const bool ok = pz >= 0 && pz < D && h >= 0 && h < H && w >= 0 && w < W;
      const __nv_bfloat16* src = ok ? x + (((int64_t)pz * H + h) * W + w) * CIN + j * 8 : x;
      wn_cp_async16(sX + ((pk * WN_XH + hh) * XW + c) * (CIN * 2) + ((j ^ wn_swz<NCX>(c)) << 4), src, ok ? 16u : 0u);
    }
    for (int q = threadIdx.x; q < WN_TH * TW * NCZ; q += WN_THREADS) {
      const int j = q % NCZ, vq = q / NCZ;
      const int hl = vq / TW, vl = vq - hl * TW;
      const int h = h0 + hl, w = w0 + vl;
      const bool ok = h < H && w < W;
      const __nv_bfloat16* src = ok ? dz + (((int64_t)d * H + h) * W + w) * COUT + j * 8 : dz;
      wn_cp_async16(sZ + (hl * TW + vl) * (COUT * 2) + ((j ^ wn_swz<NCZ>(vl)) << 4), src, ok ? 16u : 0u);
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();

#pragma unroll 1
    for (int hl = 0; hl < WN_TH; ++hl) {
      const uint32_t zrow = sZ + hl * TW * (COUT * 2);
      const uint32_t xrow = sX + ((kd * WN_XH + hl + kh) * XW) * (CIN * 2);
#pragma unroll 1
      for (int b = 0; b < TW / 16; ++b) {
        // A fragments: rows = co, cols = voxels b*16 .. b*16+15 of row hl
        uint32_t a[MB][4];
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          if (COUT >= 16) {
            // matrices: (voxels 0-7, chunk 2mb), (voxels 0-7, chunk 2mb+1), (voxels 8-15, chunk 2mb), (voxels 8-15, chunk 2mb+1)
            const int v = b * 16 + (lm >> 1) * 8 + lr, j = 2 * mb + (lm & 1);
            wn_ldmatrix_x4_trans(zrow + v * (COUT * 2) + ((j ^ wn_swz<NCZ>(v)) << 4), a[mb][0], a[mb][1], a[mb][2], a[mb][3]);
          } else {
            const int v = b * 16 + (lm & 1) * 8 + lr;  // lanes 0-15 supply the addresses of an x2
            wn_ldmatrix_x2_trans(zrow + v * (COUT * 2), a[mb][0], a[mb][2]);
            a[mb][1] = a[mb][3] = 0u;  // Cout = 8: rows 8-15 of the product are zero
          }
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
          for (int nb = 0; nb < NB; nb += 2) {
            uint32_t b0, b1, b2 = 0u, b3 = 0u;
            if (NB >= 2) {
              // matrices: (voxels 0-7, chunk nb), (voxels 8-15, chunk nb), (voxels 0-7, chunk nb+1), (voxels 8-15, chunk nb+1)
              const int c = b * 16 + (lm & 1) * 8 + lr + kw, j = nb + (lm >> 1);
              wn_ldmatrix_x4_trans(xrow + c * (CIN * 2) + ((j ^ wn_swz<NCX>(c)) << 4), b0, b1, b2, b3);
            } else {
              const int c = b * 16 + (lm & 1) * 8 + lr + kw;
              wn_ldmatrix_x2_trans(xrow + c * (CIN * 2), b0, b1);
            }
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
              wn_mma_bf16(acc[kw][mb][nb], a[mb][0], a[mb][1], a[mb][2], a[mb][3], b0, b1);
              if (NB >= 2) wn_mma_bf16(acc[kw][mb][nb + 1], a[mb][0], a[mb][1], a[mb][2], a[mb][3], b2, b3);
            }
          }
        }
      }
    }
  }

  // every (tap, co, ci) of the CTA lives in exactly one thread: one red.add each
  const int r0 = lane >> 2, c0 = (lane & 3) * 2;
#pragma unroll
  for (int kw = 0; kw < 3; ++kw) {
    float* out = dw + (int64_t)((kd * 3 + kh) * 3 + kw) * COUT * CIN;
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const int co = mb * 16 + r0, ci = nb * 8 + c0;
        atomicAdd(out + co * CIN + ci, acc[kw][mb][nb][0]);
        atomicAdd(out + co * CIN + ci + 1, acc[kw][mb][nb][1]);
        if (COUT >= 16) {
          atomicAdd(out + (co + 8) * CIN + ci, acc[kw][mb][nb][2]);
          atomicAdd(out + (co + 8) * CIN + ci + 1, acc[kw][mb][nb][3]);
        }
      }
  }
}

template <int CIN, int COUT, int TW>
static int launch_wgrad_narrow(const void* x, const void* dz, float* dw, int64_t D, int64_t H, int64_t W, int64_t dil,
                               cudaStream_t stream) {
  using Cfg = WnCfg<CIN, COUT, TW>;
  auto kern = wgrad_narrow_kernel<CIN, COUT, TW>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) {
      set_error("wgrad_narrow: cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM, cudaGetErrorString(e));
      return CVIT_ERR_CUDA;
    }
    configured = true;
  }
  const int64_t tiles = D * ((H + WN_TH - 1) / WN_TH) * ((W + TW - 1) / TW);
  int grid = (Cfg::SMEM <= 110 * 1024 ? 2 : 1) * num_sms();
  if (grid > tiles) grid = (int)tiles;
  kern<<<grid, WN_THREADS, Cfg::SMEM, stream>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dz), dw,
                                                (int)D, (int)H, (int)W, (int)dil, (int)tiles);
  return check_launch("wgrad_narrow_kernel");
}

}  // namespace cvit

using namespace cvit;

// dw (fp32 [27][Cout][Cin], tap = (kd*3+kh)*3+kw) += weight gradient of a 3x3x3 depth-dilated "same" convolution with
// narrow channel counts; x (the forward input, bf16 [D,H,W,Cin]) and dz (the output gradient, bf16 [D,H,W,Cout]).
// Supported (Cin, Cout): (8, 8), (16, 16), (32, 16), (32, 32).
extern "C" int cvit_wgrad_narrow_ndhwc(const void* x, const void* dz, float* dw, int64_t D, int64_t H, int64_t W, int64_t Cin,
                                       int64_t Cout, int64_t dil, void* stream) {
  if (!x || !dz || !dw || D <= 0 || H <= 0 || W <= 0 || dil <= 0) {
    set_error("wgrad_narrow: bad arguments (D=%lld H=%lld W=%lld dil=%lld)", (long long)D, (long long)H, (long long)W,
              (long long)dil);
    return CVIT_ERR_INVALID;
  }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dz)) & 15u) {
    set_error("wgrad_narrow: x and dz must be 16-byte aligned");
    return CVIT_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 8 && Cout == 8) return launch_wgrad_narrow<8, 8, 128>(x, dz, dw, D, H, W, dil, st);
  if (Cin == 16 && Cout == 16) return launch_wgrad_narrow<16, 16, 128>(x, dz, dw, D, H, W, dil, st);
  if (Cin == 32 && Cout == 16) return launch_wgrad_narrow<32, 16, 64>(x, dz, dw, D, H, W, dil, st);
  if (Cin == 32 && Cout == 32) return launch_wgrad_narrow<32, 32, 64>(x, dz, dw, D, H, W, dil, st);
  set_error("wgrad_narrow: (Cin, Cout) = (%lld, %lld) unsupported: (8,8), (16,16), (32,16), (32,32)", (long long)Cin,
            (long long)Cout);
  return CVIT_ERR_UNSUPPORTED;
}
