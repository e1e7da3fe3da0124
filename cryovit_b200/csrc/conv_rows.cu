// 3x3x3 depth-dilated "same" convolution of the 16- and 32-channel layers of the CryoVIT head (SynthesisBlocks 3-4,
// models/cryovit.py:26-27,68-78) with ONE VOXEL PER MMA ROW: csrc/conv_rows8.cu generalised to more than one 8-channel chunk
// per voxel, to depth dilation and to the border-aware bias table of a folded GroupNorm.
//
//   * The input is staged as 8-channel CHUNK ARRAYS [row][voxel][8] (TMA boxes {8 channels, one chunk, 132 voxels, HT + 2
//     rows}): each is a K-major SWIZZLE_NONE operand whose rows (voxels) are 16 bytes apart and whose K chunks -- the
//     column taps -- are the same bytes one voxel further on (LBO = 16). K = 32 per chunk (taps 0..3, the fourth against
//     zero weights); the chunks accumulate into the same window: 2 CIN/8 MMAs per input row and 128 voxels.
//   * N = (output row yr, output plane slot p, 16 output channels) = 144 columns per input row: a window that slides by 48
//     columns per input row through (HT row slots) x 48 accumulator columns of tensor memory; rows at the edge of the tile
//     use the sub-windows N = 48 / 96 from image rows 96 / 48 / 0. 32 output channels run as two passes of 16.
//   * Depth dilation d: the planes of one residue class z = r + d t behave like a dilation-1 stack, so a unit walks t; the
//     output plane of step t lives in plane slot t mod 3 and the weights exist in three rotations (by t mod 3).
//   * When input plane t is done, output plane t - 1 is complete: the eight epilogue warps read its slot (thread = voxel, 32
//     contiguous bytes per row), add the voxel's bias-table row, apply the activation, store, and zero the slot.
//   * A unit = (16 steps of one residue class, two tiles of HT rows x 128 voxels): the two tiles own half of tensor memory
//     each and alternate plane by plane (one's drain under the other's MMAs).
// Against conv3d_wpackn (P voxels per row, banded weights): 32 -> 16 at 256^2 0.65 ms, 16 -> 16 0.31 ms, 32 -> 32 at 128^2
// 0.20 ms; measured numbers of this kernel in profiles/r02_head_notes.md.
#include "ptx.cuh"
#include "tmap.h"

namespace cvit {

constexpr int RN_THREADS = 320;
constexpr int RN_ZC = 16;     // steps (output planes of one residue class) per unit
constexpr int RN_SEG = 128;   // voxels per tile row
constexpr int RN_XV = 132;    // staged voxels per row: x0 - 1 .. x0 + 130
constexpr int RN_CO = 16;     // output channels per pass
constexpr int RN_SLOT = 3 * RN_CO;        // accumulator columns per output row slot
constexpr int RN_IMG_ROWS = 9 * RN_CO;    // 144
constexpr int RN_SEG_COLS = 256;

template <int CIN, int HT>
struct RnCfg {
  static constexpr int NCX = CIN / 8;
  static constexpr int ROWS = HT + 2;
  static constexpr int ROW_BYTES = RN_XV * 16;
  static constexpr int BOX_BYTES = ROWS * ROW_BYTES;           // one TMA box: a chunk array of a (plane, tile)
  static constexpr int ARR = (BOX_BYTES + 127) / 128 * 128;    // TMA destinations are 128-byte aligned
  static constexpr int STAGE = NCX * ARR;
  static constexpr int IMG_STEP = 2 * RN_IMG_ROWS * 16;        // [chunk c][144 rows][8]: one K step
  static constexpr int IMG_VAR = NCX * 2 * IMG_STEP;           // [j][K step]
  static constexpr int IMG_BYTES = 3 * IMG_VAR;                // [t mod 3]
  static constexpr int TAB_BYTES = 64 * RN_CO * 4;
  static constexpr int STAGES_RAW = (232448 - 1024 - 512 - IMG_BYTES - TAB_BYTES) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 4 ? 4 : STAGES_RAW;
  static constexpr int SMEM = STAGES * STAGE + IMG_BYTES + TAB_BYTES + 512 + 1024;
  static_assert(HT * RN_SLOT <= RN_SEG_COLS && STAGES >= 2 && HT >= 3, "tile does not fit");
};

struct RnArgs {
  const __nv_bfloat16* w_img;  // RnCfg::IMG_BYTES of this pass, host-arranged (cryovit_b200.head.rowsn_weight_image)
  const float* table;          // fp32 [64][tab_stride]: bias row by in-bounds tap masks; this pass's 16 columns start at table
  __nv_bfloat16* out;          // [D, H, W, out_stride], this pass's 16 channels start at out
  __nv_bfloat16* aux;          // act = ACT_DUAL: gelu(out); ACT_GELU_GRAD: the pre-activation z whose gelu' scales the result
  float* db;                   // ACT_GELU_GRAD: fp32 [16] column sums of what is stored are added here (may be null)
  int D, H, W, dil, act, out_stride, tab_stride;
};

__device__ __forceinline__ uint64_t rn_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {  // K-major SWIZZLE_NONE
  return static_cast<uint64_t>((addr & 0x3ffffu) >> 4) | (static_cast<uint64_t>(lbo >> 4) << 16) |
         (static_cast<uint64_t>(sbo >> 4) << 32) | (1ull << 46);
}

template <int CIN, int HT>
__global__ void __launch_bounds__(RN_THREADS, 1) conv3d_rows_kernel(const __grid_constant__ CUtensorMap tmX, const RnArgs args) {
  using Cfg = RnCfg<CIN, HT>;
  constexpr int NCX = Cfg::NCX, ROWS = Cfg::ROWS, STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sX = smem_base, sW = smem_base + STAGES * Cfg::STAGE, sTab = sW + Cfg::IMG_BYTES, sBar = sTab + Cfg::TAB_BYTES;
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * STAGES;
  const uint32_t bar_done = sBar + 16 * STAGES, bar_free = bar_done + 16, tmem_slot = bar_free + 16;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // units: (pair of tiles) x (residue class r < dil, chunk of RN_ZC steps t; plane z = r + dil t), the z part fastest
  const int nseg = (args.W + RN_SEG - 1) / RN_SEG, ntile = nseg * ((args.H + HT - 1) / HT), npair = (ntile + 1) / 2;
  const int d = args.dil;
  int zunits = 0;
  for (int r = 0; r < d && r < args.D; ++r) zunits += ((args.D - r + d - 1) / d + RN_ZC - 1) / RN_ZC;
  const int num_units = npair * zunits;
  auto unit_of = [&](int u, int& pair, int& r, int& ta, int& tb, int& T) {
    int zu = u % zunits;
    pair = u / zunits;
    for (r = 0;; ++r) {
      T = (args.D - r + d - 1) / d;  // planes in residue class r
      const int c = (T + RN_ZC - 1) / RN_ZC;
      if (zu < c) break;
      zu -= c;
    }
    ta = zu * RN_ZC;
    tb = min(ta + RN_ZC, T);
  };
  auto tile_of = [&](int pair, int g, int& x0, int& yt0) {  // a tile past the end lies outside the volume: zero input, no stores
    const int tile = 2 * pair + g;
    x0 = tile < ntile ? (tile % nseg) * RN_SEG : 0x3fff0000;
    yt0 = tile < ntile ? (tile / nseg) * HT : 0;
  };

  for (int i = threadIdx.x; i < Cfg::IMG_BYTES / 16; i += RN_THREADS)
    reinterpret_cast<uint4*>(smem_gen + (sW - smem_base))[i] = __ldg(reinterpret_cast<const uint4*>(args.w_img) + i);
  for (int i = threadIdx.x; i < 64 * RN_CO; i += RN_THREADS)
    reinterpret_cast<float*>(smem_gen + (sTab - smem_base))[i] = __ldg(args.table + (i / RN_CO) * args.tab_stride + (i % RN_CO));
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(bar_done + 8 * g, 1);
      mbar_init(bar_free + 8 * g, 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: one box per chunk and (plane, tile)
    if (lane == 0) {
      uint32_t it = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        int pair, r, ta, tb, T;
        unit_of(u, pair, r, ta, tb, T);
        for (int t = ta - 1; t <= tb; ++t) {
          if (t < 0 || t >= T) continue;
          for (int g = 0; g < 2; ++g, ++it) {
            int x0, yt0;
            tile_of(pair, g, x0, yt0);
            const uint32_t s = it % STAGES;
            mbar_wait(bar_empty + 8 * s, ((it / STAGES) & 1) ^ 1u);
            mbar_arrive_expect_tx(bar_full + 8 * s, NCX * Cfg::BOX_BYTES);
#pragma unroll
            for (int j = 0; j < NCX; ++j)
              tma_load_5d(sX + s * Cfg::STAGE + j * Cfg::ARR, &tmX, bar_full + 8 * s, 0, j, x0 - 1, yt0 - 1, r + d * t);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (row / chunk loops unrolled: constant windows)
    constexpr uint32_t id48 = umma_idesc_bf16_f32(128, 48), id96 = umma_idesc_bf16_f32(128, 96), id144 = umma_idesc_bf16_f32(128, 144);
    uint32_t it = 0, np[2] = {0, 0};
    const uint64_t a_base = rn_desc(sX + 0, 16, 128);  // array voxel 0 = x0 - 1; LBO = 16: the next column tap = the next voxel
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      int pair, r, ta, tb, T;
      unit_of(u, pair, r, ta, tb, T);
      for (int t = ta - 1; t <= tb; ++t) {
        const bool real = t >= 0 && t < T;
        const uint64_t b_base = rn_desc(sW + ((t + 3) % 3) * Cfg::IMG_VAR, RN_IMG_ROWS * 16, 128);
        for (int g = 0; g < 2; ++g) {
          mbar_wait(bar_free + 8 * g, np[g] & 1);
          tcgen05_fence_after();
          ++np[g];
          if (real) {
            const uint32_t d_seg = tmem_base + g * RN_SEG_COLS;
            const uint32_t s = it % STAGES;
            mbar_wait(bar_full + 8 * s, (it / STAGES) & 1);
            tcgen05_fence_after();
            ++it;
            if (elect_one_sync()) {
              const uint64_t a_stage = a_base + s * (Cfg::STAGE >> 4);
#pragma unroll
              for (int k = 0; k < ROWS; ++k) {
                const int i = k - 1;  // input row y' = yt0 + i feeds output rows y' - 1 + yr
                const int r0 = i == -1 ? 2 * RN_SLOT : i == 0 ? RN_SLOT : 0;
                const int slot = i <= 0 ? 0 : i - 1;
                const uint32_t idesc = (i == -1 || i == HT) ? id48 : (i == 0 || i == HT - 1) ? id96 : id144;
#pragma unroll
                for (int j = 0; j < NCX; ++j)
#pragma unroll
                  for (int ks = 0; ks < 2; ++ks)
                    umma_bf16(d_seg + slot * RN_SLOT, a_stage + ((j * Cfg::ARR + k * Cfg::ROW_BYTES) >> 4) + 2 * ks,
                              b_base + (j * 2 + ks) * (Cfg::IMG_STEP >> 4) + r0, idesc, 1u);
              }
              umma_commit(bar_empty + 8 * s);
            }
            __syncwarp();
          }
          if (elect_one_sync()) umma_commit(bar_done + 8 * g);
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: warps 2-5 the upper rows of a tile, 6-9 the lower
    const int half = (warp - 2) >> 2, q = warp & 3;
    constexpr int R0 = (HT + 1) / 2;                 // rows [0, R0) / [R0, HT)
    const int rlo = half ? R0 : 0, rhi = half ? HT : R0;
    const float4* tab4 = reinterpret_cast<const float4*>(smem_gen + (sTab - smem_base));
    const uint32_t zeros[16] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    auto seg_addr = [&](int g) { return tmem_base + g * RN_SEG_COLS + (static_cast<uint32_t>(q * 32) << 16); };
    auto zero_all = [&](int g) {
#pragma unroll 1
      for (int c = rlo * RN_SLOT; c < rhi * RN_SLOT; c += 16) tmem_st_32x16(seg_addr(g) + c, zeros);
    };
    auto hand_back = [&](int g) {
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free + 8 * g);
    };
    for (int g = 0; g < 2; ++g) {
      zero_all(g);
      hand_back(g);
    }
    uint32_t nd = 0;
    float bsum[RN_CO];  // ACT_GELU_GRAD: column sums of everything this thread stores (the bias gradient of the layer below)
#pragma unroll
    for (int c = 0; c < RN_CO; ++c) bsum[c] = 0.f;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      int pair, r, ta, tb, T;
      unit_of(u, pair, r, ta, tb, T);
      for (int t = ta - 1; t <= tb; ++t, ++nd) {
        const int to = t - 1, p = (to + 3) % 3, zo = r + d * to;
        for (int g = 0; g < 2; ++g) {
          int x0, yt0;
          tile_of(pair, g, x0, yt0);
          const uint32_t t_seg = seg_addr(g);
          const int x = x0 + q * 32 + lane;
          const bool live = to >= ta && to < tb;
          constexpr int RMAX = (HT + 1) / 2;
          // ACT_GELU_GRAD: this thread's z values of the plane about to be drained are requested BEFORE the wait for its MMAs,
          // so the round trip to L2 / DRAM hides behind them (a load inside the drain is what made the W-packed kernels'
          // gelu' epilogue slower than a separate pass, profiles/r02_train_notes.md)
          uint4 zz[RMAX][2];
          if (args.act == ACT_GELU_GRAD && live) {
#pragma unroll
            for (int k = 0; k < RMAX; ++k) {
              const int y = yt0 + rlo + k;
              zz[k][0] = zz[k][1] = make_uint4(0u, 0u, 0u, 0u);
              if (rlo + k < rhi && y < args.H && x < args.W) {
                const uint4* zp = reinterpret_cast<const uint4*>(args.aux + (((size_t)zo * args.H + y) * args.W + x) * args.out_stride);
                zz[k][0] = __ldg(zp);
                zz[k][1] = __ldg(zp + 1);
              }
            }
          }
          mbar_wait(bar_done + 8 * g, nd & 1);
          tcgen05_fence_after();
          // Read this half's rows of the finished plane, zero the slot and hand it back BEFORE the math and the stores: the
          // other tile's MMAs cover the drain's TMEM round trips, not its GELUs.
          uint32_t v[RMAX][16];
          if (live) {
#pragma unroll
            for (int k = 0; k < RMAX; ++k)
              if (rlo + k < rhi) tmem_ld_32x16(t_seg + (rlo + k) * RN_SLOT + p * RN_CO, v[k]);
            tmem_ld_wait();
          }
          if (t == tb) {
            zero_all(g);  // the unit's last step: the halo planes' leftovers go too
          } else {
#pragma unroll
            for (int k = 0; k < RMAX; ++k)
              if (rlo + k < rhi) tmem_st_32x16(t_seg + (rlo + k) * RN_SLOT + p * RN_CO, zeros);
          }
          hand_back(g);
          if (live) {
            const int dm = (zo >= d ? 1 : 0) | (zo + d < args.D ? 2 : 0);
            const int wm = (x >= 1 ? 1 : 0) | (x + 1 < args.W ? 2 : 0);
#pragma unroll
            for (int k = 0; k < RMAX; ++k) {
              const int y = yt0 + rlo + k;
              if (rlo + k < rhi && y < args.H && x < args.W) {
                const int hm = (y >= 1 ? 1 : 0) | (y + 1 < args.H ? 2 : 0);
                const float4* row = tab4 + ((dm * 4 + hm) * 4 + wm) * (RN_CO / 4);
                const size_t off = (((size_t)zo * args.H + y) * args.W + x) * args.out_stride;
                float o[16];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const float4 tb4 = row[c];
                  o[4 * c] = __uint_as_float(v[k][4 * c]) + tb4.x;
                  o[4 * c + 1] = __uint_as_float(v[k][4 * c + 1]) + tb4.y;
                  o[4 * c + 2] = __uint_as_float(v[k][4 * c + 2]) + tb4.z;
                  o[4 * c + 3] = __uint_as_float(v[k][4 * c + 3]) + tb4.w;
                }
                uint32_t pk[8];
                if (args.act == ACT_GELU_GRAD) {
                  const uint32_t zw[8] = {zz[k][0].x, zz[k][0].y, zz[k][0].z, zz[k][0].w, zz[k][1].x, zz[k][1].y, zz[k][1].z, zz[k][1].w};
#pragma unroll
                  for (int c = 0; c < 8; ++c) {
                    pk[c] = act_gelu_grad_pair(o[2 * c], o[2 * c + 1], zw[c]);
                    bsum[2 * c] += __uint_as_float(pk[c] << 16);  // sums over the STORED gradient
                    bsum[2 * c + 1] += __uint_as_float(pk[c] & 0xffff0000u);
                  }
                  st_global_v8(args.out + off, pk);
                  continue;
                }
                if (args.act == ACT_DUAL) {
#pragma unroll
                  for (int c = 0; c < 8; ++c) pk[c] = pack_bf16x2(o[2 * c], o[2 * c + 1]);
                  st_global_v8(args.out + off, pk);
                }
                if (args.act) {
#pragma unroll
                  for (int c = 0; c < 8; ++c) gelu_erf2(o[2 * c], o[2 * c + 1]);
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) pk[c] = pack_bf16x2(o[2 * c], o[2 * c + 1]);
                st_global_v8((args.act == ACT_DUAL ? args.aux : args.out) + off, pk);  // 16 channels = one 32-byte sector
              }
            }
          }
        }
      }
    }
    if (args.act == ACT_GELU_GRAD && args.db) {  // bias gradient: one shuffle tree and 16 atomics per warp
#pragma unroll
      for (int c = 0; c < RN_CO; ++c) {
        const float tot = warp_sum(bsum[c]);
        if (lane == 0) atomicAdd(args.db + c, tot);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int CIN, int HT>
static int launch_rows(const void* x, const RnArgs& a, cudaStream_t stream) {
  using Cfg = RnCfg<CIN, HT>;
  CUtensorMap tmX;
  uint64_t dims[5] = {8, (uint64_t)Cfg::NCX, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.D};
  uint64_t strides[5] = {0, 16, (uint64_t)CIN * 2, (uint64_t)a.W * CIN * 2, (uint64_t)a.H * a.W * CIN * 2};
  uint32_t box[5] = {8, 1, RN_XV, (uint32_t)Cfg::ROWS, 1};
  int rc = encode_tmap(&tmX, TmapDtype::BF16, 5, x, dims, strides, box, 0);
  if (rc) return rc;
  auto kern = conv3d_rows_kernel<CIN, HT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) {
      set_error("conv3d_rows: cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM, cudaGetErrorString(e));
      return CVIT_ERR_CUDA;
    }
    configured = true;
  }
  const int nseg = (a.W + RN_SEG - 1) / RN_SEG, ntile = nseg * ((a.H + HT - 1) / HT);
  int64_t zunits = 0;
  for (int r = 0; r < a.dil && r < a.D; ++r) zunits += ((a.D - r + a.dil - 1) / a.dil + RN_ZC - 1) / RN_ZC;
  const int64_t units = (int64_t)((ntile + 1) / 2) * zunits;
  int grid = num_sms();
  if (grid > units) grid = (int)units;
  kern<<<grid, RN_THREADS, Cfg::SMEM, stream>>>(tmX, a);
  return check_launch("conv3d_rows_kernel");
}

}  // namespace cvit

using namespace cvit;

// bytes of the weight image of ONE 16-channel output pass (cryovit_b200.head.rowsn_weight_image); -1: no such kernel
extern "C" int64_t cvit_conv3d_rows_weight_bytes(int64_t Cin) {
  if (Cin == 16) return RnCfg<16, 5>::IMG_BYTES;
  if (Cin == 32) return RnCfg<32, 4>::IMG_BYTES;
  return -1;
}

// out (bf16 [D,H,W,Cout]) = conv3d(x bf16 [D,H,W,Cin], 3x3x3, "same", dilation (dil,1,1)) + bias_table row of the voxel, act as
// the other *_aux entry points (0 none, 1 GELU, 2 out = pre-activation and aux = GELU of it, 3 out = result * gelu'(aux), with
// db (fp32 [Cout], may be null) += the column sums of what is stored: the bias gradient of the layer below). Cin, Cout in {16, 32};
// w_img: Cout / 16 images of cvit_conv3d_rows_weight_bytes(Cin) bytes each (output channels 16 h .. 16 h + 15);
// bias_table fp32 [64][Cout] (the *_tab layout: a plain bias = 64 equal rows).
extern "C" int cvit_conv3d_rows_ndhwc(const void* x, const void* w_img, const float* bias_table, void* out, int64_t D, int64_t H,
                                      int64_t W, int64_t Cin, int64_t Cout, int64_t dil, int act, void* aux, float* db, void* stream) {
  if (!x || !w_img || !bias_table || !out || D <= 0 || H <= 0 || W <= 0 || dil <= 0) {
    set_error("conv3d_rows: bad arguments (D=%lld H=%lld W=%lld Cin=%lld Cout=%lld dil=%lld)", (long long)D, (long long)H,
              (long long)W, (long long)Cin, (long long)Cout, (long long)dil);
    return CVIT_ERR_INVALID;
  }
  if ((Cin != 16 && Cin != 32) || (Cout != 16 && Cout != 32)) {
    set_error("conv3d_rows: Cin=%lld Cout=%lld unsupported (16 or 32 each; cvit_conv3d_wpackn_ndhwc / cvit_conv3d_halo_ndhwc otherwise)",
              (long long)Cin, (long long)Cout);
    return CVIT_ERR_UNSUPPORTED;
  }
  if (act < 0 || act > 3 || (act >= 2 && !aux) ||
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_img) | reinterpret_cast<uintptr_t>(out) |
        reinterpret_cast<uintptr_t>(aux) | reinterpret_cast<uintptr_t>(bias_table)) & 15u)) {
    set_error("conv3d_rows: act must be 0, 1, 2 or 3 (2, 3 with aux); x, w_img, bias_table, out and aux 16-byte aligned");
    return CVIT_ERR_INVALID;
  }
  const int64_t img = cvit_conv3d_rows_weight_bytes(Cin);
  for (int h = 0; h < Cout / RN_CO; ++h) {
    RnArgs a;
    a.w_img = reinterpret_cast<const __nv_bfloat16*>(static_cast<const uint8_t*>(w_img) + h * img);
    a.table = bias_table + h * RN_CO;
    a.out = static_cast<__nv_bfloat16*>(out) + h * RN_CO;
    a.aux = aux ? static_cast<__nv_bfloat16*>(aux) + h * RN_CO : nullptr;
    a.db = db ? db + h * RN_CO : nullptr;
    a.D = (int)D;
    a.H = (int)H;
    a.W = (int)W;
    a.dil = (int)dil;
    a.act = act;
    a.out_stride = (int)Cout;
    a.tab_stride = (int)Cout;
    int rc = Cin == 16 ? launch_rows<16, 5>(x, a, (cudaStream_t)stream) : launch_rows<32, 4>(x, a, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return CVIT_OK;
}
