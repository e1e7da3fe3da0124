// 3x3x3 depth-dilated "same" convolution + bias (table) + GELU for the 16- and 32-channel layers of the CryoVIT head
// (SynthesisBlocks 3 and 4, models/cryovit.py:26-27,68-78) on tcgen05 with P CONSECUTIVE OUTPUT VOXELS ALONG W PACKED
// INTO ONE MMA ROW.
//
// Why. conv_halo.cu spends one MMA (M = 128, N = Cout, K = 16) per (tap, 16 channels) and 128 voxels: 54 MMAs per tile
// for 32 -> 16, and an M = 128 MMA costs ~65-73 cycles however small N is (the A tile is streamed from shared memory at
// ~64 B/clk): 27 cycles per voxel, tensor pipe 13 % busy, 0.77 ms for SB4's first convolution (profiles/r01_ncu_full_v14).
// The work has to go into N. Here an MMA row is a GROUP of P voxels, K runs over the group's input window -- (P + 2)
// voxels x Cin channels -- and N over its P x Cout outputs; the column taps kw are folded into a banded weight matrix
//     B[(j_in, ci)][(j_out, co)] = w[co][ci][kd][kh][kw = j_in - j_out]      (zero outside 0..2):
//     32 -> 16, P = 2:  8 MMAs (N = 32) per (kd, kh) and 256 voxels   = 0.28 MMAs / voxel   (was 0.42): 0.77 -> 0.65 ms
//     16 -> 16, P = 4:  6 MMAs (N = 64) per (kd, kh) and 512 voxels   = 0.11 MMAs / voxel   (was 0.21): 0.39 -> 0.31 ms
//     32 -> 32, P = 2:  8 MMAs (N = 64) per (kd, kh) and 256 voxels   = 0.28 MMAs / voxel   (was 0.42)
// with the banded weights of all 27 taps resident in shared memory (72 / 108 / 144 KB).
//
// Staging. With P * Cin * 2 = 128 bytes per group (P = 2 at 32 channels, P = 4 at 16) a row of the volume is a sequence
// of 128-byte LINES, line l = voxels P l .. P l + P - 1, and the window of the MMA row that produces voxels
// P g + 1 .. P g + P is exactly lines g and g + 1 (K = 256 bytes; 192 at 16 channels). A plane staged PLAINLY by TMA with
// the 128-byte swizzle -- [h 18][8 lines][128 B] -- therefore already IS a K-major SWIZZLE_128B A operand: rows
// g = 0..7 are consecutive lines, K steps 0..3 walk line g (+32 B each), the row tap kh moves one staged row (1024 B,
// which keeps the swizzle phase) on; K steps 4..7 read the same tile of the lines one further right, staged as a second
// box. A tile starts at group 8 t - 1, and TMA's out-of-bounds zero fill is the "same" padding on all three axes (line
// -1, line W / P, rows -1 and H; depth taps outside [0, D) are skipped outright). No loader warps, no re-layout.
// (Measured alternatives, profiles/r02_head_notes.md: per-piece cp.async staging of a duplicated, un-swizzled window
// layout was L1TEX-bound at 0.92 ms; ONE box of 9 lines per row with descriptor base offsets returns wrong data: the
// 8-row-group stride of a swizzled operand must keep the 1024-byte phase.)
//
// One CTA per SM, persistent over (d, tile) work items, 192 threads: warp 0 TMA producer (two boxes per plane, a plane is
// one pipeline stage), warp 1 MMA issuer (warp-uniform, elect.sync) and TMEM owner, warps 2-5 epilogue (thread = MMA row
// = P voxels x Cout channels: bias-table row per voxel + GELU -> bf16). Accumulators double-buffered in TMEM.
#include <stdlib.h>

#include "ptx.cuh"
#include "tmap.h"

namespace cvit {

constexpr int WN_G = 8, WN_TH = 16;           // MMA row = (h, g): 16 x 8 = 128
constexpr int WN_ROWS = WN_TH + 2;            // staged rows (1-voxel halo)

struct WnArgs {
  const __nv_bfloat16* x;      // [D, H, W, CIN]
  const __nv_bfloat16* w_img;  // WnCfg::W_BYTES, host-arranged (cryovit_b200.head.wpackn_weight_image)
  const float* table;          // fp32 [64][COUT]: bias row by in-bounds tap masks (a plain bias = 64 equal rows)
  __nv_bfloat16* out;          // [D, H, W, n_valid]
  __nv_bfloat16* aux;          // act = ACT_DUAL: gelu(out); ACT_GELU_GRAD: the pre-activation whose gelu' scales the result
  int D, H, W, dil, n_valid, act;
};

__device__ __forceinline__ uint64_t wn_desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return static_cast<uint64_t>((addr & 0x3ffffu) >> 4) | (static_cast<uint64_t>(lbo >> 4) << 16) |
         (static_cast<uint64_t>(sbo >> 4) << 32) | (1ull << 46);
}
template <int CIN, int COUT, int P>
struct WtCfg {
  static constexpr int K_BYTES = (P + 2) * CIN * 2;          // bytes of K per row: 256 or 192
  static constexpr int KSTEPS = K_BYTES / 32;
  static constexpr int N = P * COUT;
  static constexpr int LINES = 8;
  static constexpr int ROW_PITCH = LINES * 128;
  static constexpr int BOX_BYTES = WN_ROWS * ROW_PITCH;      // one TMA box
  static constexpr int PLANE_RAW = 2 * BOX_BYTES;
  static constexpr int BOX_STRIDE = (BOX_BYTES + 1023) / 1024 * 1024;
  static constexpr int PLANE = 2 * BOX_STRIDE;
  static constexpr int W_BYTES = 9 * KSTEPS * 2 * N * 16;
  static constexpr int TAB_BYTES = 64 * COUT * 4;
  static constexpr int STAGES_RAW = (232448 - 1024 - 512 - W_BYTES - TAB_BYTES) / PLANE;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM = STAGES * PLANE + W_BYTES + TAB_BYTES + 512 + 1024;
  static constexpr int TMEM_COLS = 2 * N <= 32 ? 32 : 2 * N <= 64 ? 64 : 2 * N <= 128 ? 128 : 256;
  static_assert(P * CIN * 2 == 128, "a group of P voxels must be one 128-byte line");
  static_assert(STAGES >= 2, "needs at least two plane slots");
};

template <int CIN, int COUT, int P>
__global__ void __launch_bounds__(192, 1) conv3d_wpackt_kernel(const __grid_constant__ CUtensorMap tmX, const WnArgs args) {
  using Cfg = WtCfg<CIN, COUT, P>;
  constexpr int STAGES = Cfg::STAGES, N = Cfg::N;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sIn = smem_base;
  const uint32_t sW = smem_base + STAGES * Cfg::PLANE;
  const uint32_t sTab = sW + Cfg::W_BYTES;
  const uint32_t sBar = sTab + Cfg::TAB_BYTES;
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * STAGES;
  const uint32_t bar_tfull = sBar + 16 * STAGES, bar_tempty = bar_tfull + 16, tmem_slot = bar_tempty + 16;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_w = (args.W / P + 1 + WN_G - 1) / WN_G, tiles_h = (args.H + WN_TH - 1) / WN_TH;
  const int per_plane = tiles_w * tiles_h;
  const int num_tiles = args.D * per_plane;

  {
    const uint4* src = reinterpret_cast<const uint4*>(args.w_img);
    uint4* dst = reinterpret_cast<uint4*>(smem_gen + (sW - smem_base));
    for (int i = threadIdx.x; i < Cfg::W_BYTES / 16; i += 192) dst[i] = __ldg(src + i);
    const uint4* tsrc = reinterpret_cast<const uint4*>(args.table);
    uint4* tdst = reinterpret_cast<uint4*>(smem_gen + (sTab - smem_base));
    for (int i = threadIdx.x; i < Cfg::TAB_BYTES / 16; i += 192) tdst[i] = __ldg(tsrc + i);
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  auto tile_of = [&](int tile, int& d, int& h0, int& g0) {
    d = tile / per_plane;
    const int r = tile - d * per_plane;
    const int th = r / tiles_w;
    h0 = th * WN_TH;
    g0 = (r - th * tiles_w) * WN_G - 1;  // first group (= first line) of the tile
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int n = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int d, h0, g0;
        tile_of(tile, d, h0, g0);
        for (int kd = 0; kd < 3; ++kd) {
          const int dz = d + (kd - 1) * args.dil;
          if (dz < 0 || dz >= args.D) continue;
          const int slot = n % STAGES;
          mbar_wait(bar_empty + 8 * slot, (((n / STAGES) & 1) ^ 1) & 1);
          mbar_arrive_expect_tx(bar_full + 8 * slot, Cfg::PLANE_RAW);
          tma_load_4d(sIn + slot * Cfg::PLANE, &tmX, bar_full + 8 * slot, 0, g0, h0 - 1, dz);
          tma_load_4d(sIn + slot * Cfg::PLANE + Cfg::BOX_STRIDE, &tmX, bar_full + 8 * slot, 0, g0 + 1, h0 - 1, dz);
          ++n;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16_f32(128, N);
    constexpr uint32_t B_SBO = 128, B_LBO = N * 16, B_MMA = 2 * N * 16;
    int n = 0, acc = 0;
    uint32_t acc_ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int d = tile / per_plane;
      mbar_wait(bar_tempty + 8 * acc, acc_ph ^ 1u);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + acc * N;
      uint32_t accumulate = 0;
      for (int kd = 0; kd < 3; ++kd) {
        const int dz = d + (kd - 1) * args.dil;
        if (dz < 0 || dz >= args.D) continue;
        const int slot = n % STAGES;
        mbar_wait(bar_full + 8 * slot, (n / STAGES) & 1);
        tcgen05_fence_after();
        if (elect_one_sync()) {
          const uint32_t plane = sIn + slot * Cfg::PLANE;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t wrow = sW + ((kd * 3 + kh) * Cfg::KSTEPS) * B_MMA;
#pragma unroll
            for (int st = 0; st < Cfg::KSTEPS; ++st) {
              // K steps 0..3: box 0 (lines g0 ..), 4..7: box 1 (lines g0 + 1 ..); +32 B per step inside the line, the row
              // tap kh moves one staged row (1024 B: keeps the swizzle phase) on
              const uint64_t adesc = umma_smem_desc_kmajor<128>(plane + (st >> 2) * Cfg::BOX_STRIDE + kh * Cfg::ROW_PITCH + (st & 3) * 32);
              umma_bf16(d_tmem, adesc, wn_desc_nosw(wrow + st * B_MMA, B_LBO, B_SBO), idesc, accumulate);
              accumulate = 1;
            }
          }
          umma_commit(bar_empty + 8 * slot);
        }
        __syncwarp();
        ++n;
      }
      if (elect_one_sync()) umma_commit(bar_tfull + 8 * acc);
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1u;
    }
  } else {
    // ------------------------------------------------------------------ epilogue: thread == MMA row == P voxels
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int hl = r / WN_G, g = r - hl * WN_G;
    const float4* tab4 = reinterpret_cast<const float4*>(smem_gen + (sTab - smem_base));
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int d, h0, g0;
      tile_of(tile, d, h0, g0);
      mbar_wait(bar_tfull + 8 * acc, acc_ph);
      tcgen05_fence_after();
      const uint32_t t_acc = tmem_base + acc * N + (static_cast<uint32_t>(q * 32) << 16);
      uint32_t v[N];
#pragma unroll
      for (int c = 0; c < N; c += 32) tmem_ld_32x32(t_acc + c, *reinterpret_cast<uint32_t(*)[32]>(&v[c]));
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      const int h = h0 + hl, wv0 = P * (g0 + g) + 1;  // this row's first output voxel (may be -P + 1 .. W)
      if (h < args.H) {
        const int dm = (d >= args.dil ? 1 : 0) | (d + args.dil < args.D ? 2 : 0);
        const int hm = (h >= 1 ? 1 : 0) | (h + 1 < args.H ? 2 : 0);
        const int64_t off0 = (((int64_t)d * args.H + h) * args.W + wv0) * args.n_valid;
        __nv_bfloat16* o = args.out + off0;
#pragma unroll
        for (int j = 0; j < P; ++j) {
          const int w = wv0 + j;
          if (w < 0 || w >= args.W) continue;
          const int wm = (w >= 1 ? 1 : 0) | (w + 1 < args.W ? 2 : 0);
          const float4* row = tab4 + ((dm * 4 + hm) * 4 + wm) * (COUT / 4);
          uint32_t pk[COUT / 2];
          if (args.act == ACT_GELU_GRAD) {  // input gradient times gelu'(z) of the layer below
            uint32_t z[COUT / 2];
#pragma unroll
            for (int c = 0; c < COUT; c += 8) {
              uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
              if (c < args.n_valid) z4 = __ldg(reinterpret_cast<const uint4*>(args.aux + off0 + (int64_t)j * args.n_valid + c));
              z[c / 2] = z4.x; z[c / 2 + 1] = z4.y; z[c / 2 + 2] = z4.z; z[c / 2 + 3] = z4.w;
            }
#pragma unroll
            for (int c = 0; c < COUT / 4; ++c) {
              const float4 tb = row[c];
              pk[2 * c] = act_gelu_grad_pair(__uint_as_float(v[j * COUT + 4 * c]) + tb.x, __uint_as_float(v[j * COUT + 4 * c + 1]) + tb.y, z[2 * c]);
              pk[2 * c + 1] = act_gelu_grad_pair(__uint_as_float(v[j * COUT + 4 * c + 2]) + tb.z, __uint_as_float(v[j * COUT + 4 * c + 3]) + tb.w, z[2 * c + 1]);
            }
          } else {
            uint32_t pz[COUT / 2];
#pragma unroll
            for (int c = 0; c < COUT / 4; ++c) {
              const float4 tb = row[c];
              float a0 = __uint_as_float(v[j * COUT + 4 * c]) + tb.x, a1 = __uint_as_float(v[j * COUT + 4 * c + 1]) + tb.y;
              float a2 = __uint_as_float(v[j * COUT + 4 * c + 2]) + tb.z, a3 = __uint_as_float(v[j * COUT + 4 * c + 3]) + tb.w;
              if (args.act == ACT_DUAL) {
                pz[2 * c] = pack_bf16x2(a0, a1);
                pz[2 * c + 1] = pack_bf16x2(a2, a3);
              }
              if (args.act) {
                gelu_erf2(a0, a1);
                gelu_erf2(a2, a3);
              }
              pk[2 * c] = pack_bf16x2(a0, a1);
              pk[2 * c + 1] = pack_bf16x2(a2, a3);
            }
            if (args.act == ACT_DUAL) {  // pre-activation -> out, activation -> aux (stored below through `o`)
#pragma unroll
              for (int c = 0; c < COUT; c += 8)
                if (c < args.n_valid)
                  *reinterpret_cast<uint4*>(o + (int64_t)j * args.n_valid + c) = make_uint4(pz[c / 2], pz[c / 2 + 1], pz[c / 2 + 2], pz[c / 2 + 3]);
            }
          }
          __nv_bfloat16* o2 = args.act == ACT_DUAL ? args.aux + off0 : o;
#pragma unroll
          for (int c = 0; c < COUT; c += 8)
            if (c < args.n_valid)
              *reinterpret_cast<uint4*>(o2 + (int64_t)j * args.n_valid + c) = make_uint4(pk[c / 2], pk[c / 2 + 1], pk[c / 2 + 2], pk[c / 2 + 3]);
        }
      }
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1u;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int CIN, int COUT, int P>
static int launch_wpackt(const WnArgs& a, cudaStream_t stream) {
  using Cfg = WtCfg<CIN, COUT, P>;
  CUtensorMap tm;
  // (64 elements = one 128-byte line, lines per row, H, D)
  uint64_t dims[4] = {64, (uint64_t)(a.W / P), (uint64_t)a.H, (uint64_t)a.D};
  uint64_t strides[4] = {0, 128, (uint64_t)a.W * CIN * 2, (uint64_t)a.H * a.W * CIN * 2};
  uint32_t box[4] = {64, (uint32_t)Cfg::LINES, WN_ROWS, 1};
  int rc = encode_tmap(&tm, TmapDtype::BF16, 4, a.x, dims, strides, box, 128);
  if (rc) return rc;
  auto kern = conv3d_wpackt_kernel<CIN, COUT, P>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) {
      set_error("conv3d_wpackt: cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM, cudaGetErrorString(e));
      return CVIT_ERR_CUDA;
    }
    configured = true;
  }
  const int num_tiles = a.D * ((a.H + WN_TH - 1) / WN_TH) * ((a.W / P + 1 + WN_G - 1) / WN_G);
  int grid = num_sms();
  if (grid > num_tiles) grid = num_tiles;
  kern<<<grid, 192, Cfg::SMEM, stream>>>(tm, a);
  return check_launch("conv3d_wpackt_kernel");
}

}  // namespace cvit

using namespace cvit;

// Voxels per MMA row chosen for (Cin, Cout_pad): 0 = the layer has no W-packed kernel.
extern "C" int64_t cvit_conv3d_wpackn_group(int64_t Cin, int64_t Cout_pad) {
  if (Cin == 32 && Cout_pad == 16) return 2;
  if (Cin == 16 && Cout_pad == 16) return 4;
  if (Cin == 32 && Cout_pad == 32) return 2;
  return 0;
}

extern "C" int64_t cvit_conv3d_wpackn_weight_bytes(int64_t Cin, int64_t Cout_pad) {
  if (Cin == 32 && Cout_pad == 16) return WtCfg<32, 16, 2>::W_BYTES;
  if (Cin == 16 && Cout_pad == 16) return WtCfg<16, 16, 4>::W_BYTES;
  if (Cin == 32 && Cout_pad == 32) return WtCfg<32, 32, 2>::W_BYTES;
  return -1;
}

extern "C" int cvit_conv3d_wpackn_ndhwc_aux(const void* x, const void* w_img, const float* bias_table, void* out, int64_t D, int64_t H,
                                            int64_t W, int64_t Cin, int64_t Cout_pad, int64_t Cout_valid, int64_t dil, int act,
                                            void* aux, void* stream);

extern "C" int cvit_conv3d_wpackn_ndhwc(const void* x, const void* w_img, const float* bias_table, void* out, int64_t D, int64_t H,
                                        int64_t W, int64_t Cin, int64_t Cout_pad, int64_t Cout_valid, int64_t dil, int act,
                                        void* stream) {
  return cvit_conv3d_wpackn_ndhwc_aux(x, w_img, bias_table, out, D, H, W, Cin, Cout_pad, Cout_valid, dil, act ? 1 : 0, nullptr, stream);
}

// act: 0 none, 1 GELU, 2 out = pre-activation and aux = GELU of it, 3 out = result * gelu'(aux) (ptx.cuh ACT_*).
extern "C" int cvit_conv3d_wpackn_ndhwc_aux(const void* x, const void* w_img, const float* bias_table, void* out, int64_t D, int64_t H,
                                            int64_t W, int64_t Cin, int64_t Cout_pad, int64_t Cout_valid, int64_t dil, int act,
                                            void* aux, void* stream) {
  if (act < 0 || act > 3 || (act >= 2 && (!aux || (reinterpret_cast<uintptr_t>(aux) & 15u)))) {
    set_error("conv3d_wpackn: act=%d needs a 16-byte aligned aux (0 none, 1 GELU, 2 out=z aux=gelu(z), 3 out=y*gelu'(aux))", act);
    return CVIT_ERR_INVALID;
  }
  const int64_t P = cvit_conv3d_wpackn_group(Cin, Cout_pad);
  if (!x || !w_img || !bias_table || !out || D <= 0 || H <= 0 || W <= 0 || dil <= 0 || Cout_valid <= 0 || Cout_valid > Cout_pad ||
      (Cout_valid % 8) != 0) {
    set_error("conv3d_wpackn: bad arguments (D=%lld H=%lld W=%lld Cin=%lld Cout=%lld/%lld dil=%lld)", (long long)D, (long long)H,
              (long long)W, (long long)Cin, (long long)Cout_valid, (long long)Cout_pad, (long long)dil);
    return CVIT_ERR_INVALID;
  }
  if (P == 0 || (W % P) != 0) {
    set_error("conv3d_wpackn: no W-packed kernel for Cin=%lld Cout_pad=%lld W=%lld (W must be a multiple of the group; use "
              "cvit_conv3d_halo_ndhwc otherwise)", (long long)Cin, (long long)Cout_pad, (long long)W);
    return CVIT_ERR_UNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_img) | reinterpret_cast<uintptr_t>(out) |
       reinterpret_cast<uintptr_t>(bias_table)) & 15u) {
    set_error("conv3d_wpackn: x, w_img, bias_table and out must be 16-byte aligned");
    return CVIT_ERR_INVALID;
  }
  WnArgs a;
  a.x = static_cast<const __nv_bfloat16*>(x);
  a.w_img = static_cast<const __nv_bfloat16*>(w_img);
  a.table = bias_table;
  a.out = static_cast<__nv_bfloat16*>(out);
  a.D = (int)D;
  a.H = (int)H;
  a.W = (int)W;
  a.dil = (int)dil;
  a.n_valid = (int)Cout_valid;
  a.act = act;
  a.aux = static_cast<__nv_bfloat16*>(aux);
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 32 && Cout_pad == 16) return launch_wpackt<32, 16, 2>(a, st);
  if (Cin == 16 && Cout_pad == 16) return launch_wpackt<16, 16, 4>(a, st);
  return launch_wpackt<32, 32, 2>(a, st);
}
