// 3x3x3 depth-dilated "same" convolution + bias + GELU for the NARROW layers of the CryoVIT head (Cin 8 / 16 / 32:
// SynthesisBlocks 3 and 4 and the first output convolution, models/cryovit.py:26-33,68-78) on tcgen05, from a
// shared-memory halo tile.
//
// Why not the per-tap TMA boxes of gemm_tcgen05.cuh (AMODE_CONV3)? With few channels that scheme re-reads every
// input voxel 27 times from L2 (14.5 GB for the 256x256 layers) and is L2-bandwidth-bound at 3 ms per layer.
// Here a tile's input is staged ONCE per depth tap -- (16+2) x (8+2) voxels of the planes d-dil, d, d+dil -- and the
// 9 in-plane taps are nothing but different START ADDRESSES of the same staged block:
//
//   smem block of one plane:  [c_hi = Cin/8][18 rows h][10 cols w][8 channels = 16 B]      (no swizzle)
//   A operand of tap (kh, kw): row r = (h, w) of the 16x8 output tile reads voxel (h+kh, w+kw), i.e.
//       core matrix  = 8 consecutive w  x 16 B                     (contiguous 128 B)
//       SBO (next 8 rows = next h)      = 10 * 16 B                (the staged row pitch)
//       LBO (next 8 channels)           = 18 * 10 * 16 B           (the staged channel-chunk pitch)
//       start                           = block + (kh * 10 + kw) * 16 B
//   which is exactly the canonical K-major SWIZZLE_NONE layout ((8,m),(8,2)):((16B,SBO),(2B,LBO)) of a UMMA
//   descriptor, so no data is moved or re-laid-out per tap. For Cin = 8 one MMA (K = 16) covers TWO taps: the second
//   K chunk is the next voxel (LBO = 16 B); the odd third tap is paired with a zero-weight copy of the second.
//   B operand: all 27 taps' weights, pre-arranged by the host into their smem image [tap][c_hi][Cout][8], are copied
//   once per CTA and stay resident.
//
// One CTA per SM, persistent over (d, 16x8) tiles, 192 threads: warp 0 TMA producer (one 5-D box per depth tap;
// out-of-bounds zero fill is the "same" padding; depth taps outside [0, D) are skipped outright), warp 1 MMA issuer
// (warp-uniform, elect.sync), warps 2-5 epilogue (thread = output voxel; bias + exact-erf GELU -> bf16, 32-64 B per
// voxel, 8 neighbouring voxels contiguous). Accumulators double-buffered in TMEM.
#include "ptx.cuh"
#include "tmap.h"

namespace cvit {

constexpr int CH_TH = 16, CH_TW = 8;             // output tile (rows h, cols w) = 128 voxels = MMA M
constexpr int CH_SH = CH_TH + 2, CH_SW = CH_TW + 2;  // staged plane with its 1-voxel halo
constexpr int CH_THREADS = 192;

template <int CIN, int COUT>
struct HaloCfg {
  static constexpr int KCH = CIN / 8;                       // 16-byte channel chunks
  static constexpr int CH_PITCH = CH_SH * CH_SW * 16;       // bytes between channel chunks of one plane (LBO of A)
  static constexpr int PLANE_RAW = KCH * CH_PITCH;
  static constexpr int PLANE = (PLANE_RAW + 127) / 128 * 128;  // TMA destinations stay 128-byte aligned
  static constexpr int STAGE = 3 * PLANE;
  // MMAs (K = 16) per in-plane row of taps: Cin >= 16 -> 3 taps x Cin/16; Cin == 8 -> 2 (tap pairs)
  static constexpr int MMA_PER_KH = CIN >= 16 ? 3 * (CIN / 16) : 2;
  static constexpr int W_BYTES = 9 * MMA_PER_KH * 2 * COUT * 16;  // [kd*3+kh][mma][2 K chunks][Cout][16 B]
  static constexpr int TAB_BYTES = 64 * COUT * 4;           // bias table of a folded GroupNorm (HaloArgs::bias_table)
  static constexpr int STAGES_RAW = (200 * 1024 - W_BYTES - TAB_BYTES) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM = STAGES * STAGE + W_BYTES + TAB_BYTES + 256 + 1024;
  static constexpr int TMEM_COLS = COUT <= 16 ? 32 : 2 * COUT;   // two accumulators
  static_assert(CIN == 8 || CIN == 16 || CIN == 32, "narrow layers only");
  static_assert(COUT == 16 || COUT == 32, "MMA N (zero-padded output channels)");
  static_assert(STAGES >= 2, "needs at least two stages");
};

struct HaloArgs {
  const __nv_bfloat16* w_img;  // host-arranged smem image of the weights, HaloCfg::W_BYTES
  const float* bias;           // [COUT] (zero padded)
  const float* bias_table;     // optional fp32 [64][COUT]: per-voxel bias row by in-bounds tap masks (folded GroupNorm,
                               // see gn_fold.cu / GemmArgs::bias_table); replaces `bias` when non-null
  __nv_bfloat16* out;          // [D, H, W, n_valid]
  int D, H, W, dil, n_valid;
  int act;  // ptx.cuh ACT_*: 1 = GELU (inference), 0 = none, 2 = out = z and aux = gelu(z), 3 = out = y * gelu'(aux)
  __nv_bfloat16* aux;  // [D, H, W, n_valid], act 2 / 3 only
};

// K-major, SWIZZLE_NONE shared-memory descriptor (cute::UMMA::SmemDescriptor): start>>4 @[0,14), LBO>>4 @[16,30)
// (stride between 8-element K chunks), SBO>>4 @[32,46) (stride between 8-row groups), version 1 @[46,48), layout 0.
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return static_cast<uint64_t>((addr & 0x3ffffu) >> 4) | (static_cast<uint64_t>(lbo >> 4) << 16) |
         (static_cast<uint64_t>(sbo >> 4) << 32) | (1ull << 46);
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(CH_THREADS, 1)
conv3d_halo_kernel(const __grid_constant__ CUtensorMap tmX, const HaloArgs args) {
  using Cfg = HaloCfg<CIN, COUT>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sIn = smem_base;
  const uint32_t sW = smem_base + STAGES * Cfg::STAGE;
  const uint32_t sTab = sW + Cfg::W_BYTES;
  const uint32_t sBar = sTab + Cfg::TAB_BYTES;
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * STAGES;
  const uint32_t bar_tfull = sBar + 16 * STAGES, bar_tempty = bar_tfull + 16, tmem_slot = bar_tempty + 16;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_w = (args.W + CH_TW - 1) / CH_TW, tiles_h = (args.H + CH_TH - 1) / CH_TH;
  const int per_plane = tiles_w * tiles_h;
  const int num_tiles = args.D * per_plane;

  // weights: one cooperative copy of the host-arranged image, resident for the whole kernel
  {
    const uint4* src = reinterpret_cast<const uint4*>(args.w_img);
    uint4* dst = reinterpret_cast<uint4*>(smem_gen + (sW - smem_base));
    for (int i = threadIdx.x; i < Cfg::W_BYTES / 16; i += CH_THREADS) dst[i] = __ldg(src + i);
    if (args.bias_table) {
      const uint4* tsrc = reinterpret_cast<const uint4*>(args.bias_table);
      uint4* tdst = reinterpret_cast<uint4*>(smem_gen + (sTab - smem_base));
      for (int i = threadIdx.x; i < Cfg::TAB_BYTES / 16; i += CH_THREADS) tdst[i] = __ldg(tsrc + i);
    }
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
  fence_proxy_async_smem();  // the generic-proxy weight stores must be visible to the tensor core (async proxy)
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  auto tile_of = [&](int tile, int& d, int& h0, int& w0) {
    d = tile / per_plane;
    const int r = tile - d * per_plane;
    const int th = r / tiles_w;
    h0 = th * CH_TH;
    w0 = (r - th * tiles_w) * CH_TW;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int d, h0, w0;
        tile_of(tile, d, h0, w0);
        int n_planes = 0;
        for (int kd = 0; kd < 3; ++kd) {
          const int dz = d + (kd - 1) * args.dil;
          n_planes += (dz >= 0 && dz < args.D);
        }
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        mbar_arrive_expect_tx(bar_full + 8 * s, n_planes * Cfg::PLANE_RAW);
        for (int kd = 0; kd < 3; ++kd) {
          const int dz = d + (kd - 1) * args.dil;
          if (dz < 0 || dz >= args.D) continue;  // the whole depth tap is zero padding
          tma_load_5d(sIn + s * Cfg::STAGE + kd * Cfg::PLANE, &tmX, bar_full + 8 * s, 0, w0 - 1, h0 - 1, dz, 0);
        }
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform, one elected lane)
    constexpr uint32_t idesc = umma_idesc_bf16_f32(128, COUT);
    constexpr uint32_t A_SBO = CH_SW * 16, A_LBO = CIN >= 16 ? Cfg::CH_PITCH : 16;
    constexpr uint32_t B_SBO = 128, B_LBO = COUT * 16, B_MMA = 2 * COUT * 16;
    int s = 0, acc = 0;
    uint32_t ph = 0, acc_ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int d = tile / per_plane;
      mbar_wait(bar_tempty + 8 * acc, acc_ph ^ 1u);
      mbar_wait(bar_full + 8 * s, ph);
      tcgen05_fence_after();
      if (elect_one_sync()) {
        const uint32_t d_tmem = tmem_base + acc * COUT;
        uint32_t accumulate = 0;
        for (int kd = 0; kd < 3; ++kd) {
          const int dz = d + (kd - 1) * args.dil;
          if (dz < 0 || dz >= args.D) continue;
          const uint32_t plane = sIn + s * Cfg::STAGE + kd * Cfg::PLANE;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t wrow = sW + ((kd * 3 + kh) * Cfg::MMA_PER_KH) * B_MMA;
#pragma unroll
            for (int i = 0; i < Cfg::MMA_PER_KH; ++i) {
              uint32_t a_addr;
              if (CIN >= 16) {
                const int kw = i / (CIN / 16), c2 = i % (CIN / 16);  // tap, then its 16-channel K step
                a_addr = plane + (kh * CH_SW + kw) * 16 + c2 * 2 * Cfg::CH_PITCH;
              } else {
                a_addr = plane + (kh * CH_SW + i) * 16;  // taps (0,1) then (1,2): chunk 1 is the next voxel
              }
              umma_bf16(d_tmem, umma_desc_nosw(a_addr, A_LBO, A_SBO), umma_desc_nosw(wrow + i * B_MMA, B_LBO, B_SBO), idesc,
                        accumulate);
              accumulate = 1;
            }
          }
        }
        umma_commit(bar_empty + 8 * s);      // staged planes reusable
        umma_commit(bar_tfull + 8 * acc);    // accumulator complete
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1u; }
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1u;
    }
  } else {
    // ------------------------------------------------------------------ epilogue: thread == output voxel
    const int q = warp & 3;  // TMEM lane quarter (hardware rule: warp id % 4)
    const int r = q * 32 + lane;
    const int hl = r / CH_TW, wl = r % CH_TW;
    int acc = 0;
    uint32_t acc_ph = 0;
    float bias_r[COUT];  // in registers for the whole kernel, not one __ldg round trip per tile
#pragma unroll
    for (int c = 0; c < COUT; ++c) bias_r[c] = __ldg(args.bias + c);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int d, h0, w0;
      tile_of(tile, d, h0, w0);
      mbar_wait(bar_tfull + 8 * acc, acc_ph);
      tcgen05_fence_after();
      const uint32_t t_acc = tmem_base + acc * COUT + (static_cast<uint32_t>(q * 32) << 16);
      uint32_t v[COUT];
      if (COUT == 32) {
        tmem_ld_32x32(t_acc, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      } else {
        tmem_ld_32x16(t_acc, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
      }
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);  // the accumulator is in registers
      const int h = h0 + hl, w = w0 + wl;
      if (h < args.H && w < args.W) {
        __nv_bfloat16* o = args.out + (((int64_t)d * args.H + h) * args.W + w) * args.n_valid;
        if (args.bias_table) {
          // folded GroupNorm: this voxel's bias row by which taps are inside the volume (mostly row 63: a broadcast read)
          const int dm = (d >= args.dil ? 1 : 0) | (d + args.dil < args.D ? 2 : 0);
          const int hm = (h >= 1 ? 1 : 0) | (h + 1 < args.H ? 2 : 0), wm = (w >= 1 ? 1 : 0) | (w + 1 < args.W ? 2 : 0);
          const float4* row = reinterpret_cast<const float4*>(smem_gen + (sTab - smem_base)) + ((dm * 4 + hm) * 4 + wm) * (COUT / 4);
#pragma unroll
          for (int c = 0; c < COUT / 4; ++c) {
            const float4 tb = row[c];
            bias_r[4 * c] = tb.x; bias_r[4 * c + 1] = tb.y; bias_r[4 * c + 2] = tb.z; bias_r[4 * c + 3] = tb.w;
          }
        }
        // math for all COUT channels without branches (activation switch hoisted), then predicated 16-byte stores
        uint32_t pk[COUT / 2];
        if (args.act == ACT_GELU_GRAD) {  // input gradient times gelu'(z) of the layer below (training)
          const __nv_bfloat16* zp = args.aux + (o - args.out);
          uint32_t z[COUT / 2];
#pragma unroll
          for (int c = 0; c < COUT; c += 8) {
            uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
            if (c < args.n_valid) z4 = __ldg(reinterpret_cast<const uint4*>(zp + c));
            z[c / 2] = z4.x; z[c / 2 + 1] = z4.y; z[c / 2 + 2] = z4.z; z[c / 2 + 3] = z4.w;
          }
#pragma unroll
          for (int i = 0; i < COUT / 2; ++i)
            pk[i] = act_gelu_grad_pair(__uint_as_float(v[2 * i]) + bias_r[2 * i], __uint_as_float(v[2 * i + 1]) + bias_r[2 * i + 1], z[i]);
        } else if (args.act == ACT_DUAL) {  // pre-activation -> out here, activation -> aux below
          uint32_t pz[COUT / 2];
#pragma unroll
          for (int i = 0; i < COUT / 2; ++i) {
            float a = __uint_as_float(v[2 * i]) + bias_r[2 * i];
            float b = __uint_as_float(v[2 * i + 1]) + bias_r[2 * i + 1];
            pz[i] = pack_bf16x2(a, b);
            gelu_erf2(a, b);
            pk[i] = pack_bf16x2(a, b);
          }
#pragma unroll
          for (int c = 0; c < COUT; c += 8)
            if (c < args.n_valid)
              *reinterpret_cast<uint4*>(o + c) = make_uint4(pz[c / 2], pz[c / 2 + 1], pz[c / 2 + 2], pz[c / 2 + 3]);
          o = args.aux + (o - args.out);
        } else if (args.act) {
#pragma unroll
          for (int i = 0; i < COUT / 2; ++i) {
            float a = __uint_as_float(v[2 * i]) + bias_r[2 * i];
            float b = __uint_as_float(v[2 * i + 1]) + bias_r[2 * i + 1];
            gelu_erf2(a, b);
            pk[i] = pack_bf16x2(a, b);
          }
        } else {
#pragma unroll
          for (int i = 0; i < COUT / 2; ++i)
            pk[i] = pack_bf16x2(__uint_as_float(v[2 * i]) + bias_r[2 * i], __uint_as_float(v[2 * i + 1]) + bias_r[2 * i + 1]);
        }
#pragma unroll
        for (int c = 0; c < COUT; c += 8)
          if (c < args.n_valid)  // n_valid is a multiple of 8
            *reinterpret_cast<uint4*>(o + c) = make_uint4(pk[c / 2], pk[c / 2 + 1], pk[c / 2 + 2], pk[c / 2 + 3]);
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

template <int CIN, int COUT>
static int launch_halo(const void* x, const HaloArgs& a, cudaStream_t stream) {
  using Cfg = HaloCfg<CIN, COUT>;
  CUtensorMap tm;
  // (8 channels, W, H, D, Cin/8): the channel chunk index is the OUTERMOST box dimension so that a box lands as
  // [c_hi][h][w][8 ch] in shared memory
  uint64_t dims[5] = {8, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.D, (uint64_t)(CIN / 8)};
  uint64_t strides[5] = {0, (uint64_t)CIN * 2, (uint64_t)a.W * CIN * 2, (uint64_t)a.H * a.W * CIN * 2, 16};
  uint32_t box[5] = {8, CH_SW, CH_SH, 1, (uint32_t)(CIN / 8)};
  int rc = encode_tmap(&tm, TmapDtype::BF16, 5, x, dims, strides, box, 0);
  if (rc) return rc;
  auto kern = conv3d_halo_kernel<CIN, COUT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) {
      set_error("conv3d_halo: cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM, cudaGetErrorString(e));
      return CVIT_ERR_CUDA;
    }
    configured = true;
  }
  const int num_tiles = a.D * ((a.H + CH_TH - 1) / CH_TH) * ((a.W + CH_TW - 1) / CH_TW);
  int grid = num_sms();
  if (grid > num_tiles) grid = num_tiles;
  kern<<<grid, CH_THREADS, Cfg::SMEM, stream>>>(tm, a);
  return check_launch("conv3d_halo_kernel");
}

}  // namespace cvit

using namespace cvit;

extern "C" int64_t cvit_conv3d_halo_weight_bytes(int64_t Cin, int64_t Cout_pad) {
  if (Cout_pad != 16 && Cout_pad != 32) return -1;
  if (Cin == 8) return HaloCfg<8, 16>::W_BYTES * (Cout_pad / 16);
  if (Cin == 16) return HaloCfg<16, 16>::W_BYTES * (Cout_pad / 16);
  if (Cin == 32) return HaloCfg<32, 16>::W_BYTES * (Cout_pad / 16);
  return -1;
}

extern "C" int cvit_conv3d_halo_ndhwc_act(const void* x, const void* w_img, const float* bias, void* out, int64_t D, int64_t H,
                                          int64_t W, int64_t Cin, int64_t Cout_pad, int64_t Cout_valid, int64_t dil, int act,
                                          void* stream);

extern "C" int cvit_conv3d_halo_ndhwc(const void* x, const void* w_img, const float* bias, void* out, int64_t D, int64_t H,
                                      int64_t W, int64_t Cin, int64_t Cout_pad, int64_t Cout_valid, int64_t dil,
                                      void* stream) {
  return cvit_conv3d_halo_ndhwc_act(x, w_img, bias, out, D, H, W, Cin, Cout_pad, Cout_valid, dil, 1, stream);
}

static int conv3d_halo_impl(const void* x, const void* w_img, const float* bias, const float* bias_table, void* out, int64_t D,
                           int64_t H, int64_t W, int64_t Cin, int64_t Cout_pad, int64_t Cout_valid, int64_t dil, int act,
                           void* aux, void* stream);

extern "C" int cvit_conv3d_halo_ndhwc_act(const void* x, const void* w_img, const float* bias, void* out, int64_t D, int64_t H,
                                          int64_t W, int64_t Cin, int64_t Cout_pad, int64_t Cout_valid, int64_t dil, int act,
                                          void* stream) {
  return conv3d_halo_impl(x, w_img, bias, nullptr, out, D, H, W, Cin, Cout_pad, Cout_valid, dil, act ? 1 : 0, nullptr, stream);
}

// act: 0 none, 1 GELU, 2 out = pre-activation and aux = GELU of it, 3 out = result * gelu'(aux) (ptx.cuh ACT_*).
extern "C" int cvit_conv3d_halo_ndhwc_aux(const void* x, const void* w_img, const float* bias, void* out, int64_t D, int64_t H,
                                          int64_t W, int64_t Cin, int64_t Cout_pad, int64_t Cout_valid, int64_t dil, int act,
                                          void* aux, void* stream) {
  if (act < 0 || act > 3 || (act >= 2 && (!aux || (reinterpret_cast<uintptr_t>(aux) & 15u)))) {
    set_error("conv3d_halo: act=%d needs a 16-byte aligned aux (0 none, 1 GELU, 2 out=z aux=gelu(z), 3 out=y*gelu'(aux))", act);
    return CVIT_ERR_INVALID;
  }
  return conv3d_halo_impl(x, w_img, bias, nullptr, out, D, H, W, Cin, Cout_pad, Cout_valid, dil, act, aux, stream);
}

// The convolution after a folded GroupNorm: per-voxel bias rows from the 64-row table (see gn_fold.cu), + GELU.
extern "C" int cvit_conv3d_halo_ndhwc_tab(const void* x, const void* w_img, const float* bias_table, void* out, int64_t D, int64_t H,
                                          int64_t W, int64_t Cin, int64_t Cout_pad, int64_t Cout_valid, int64_t dil, void* stream) {
  if (!bias_table || (reinterpret_cast<uintptr_t>(bias_table) & 15u)) {
    set_error("conv3d_halo_tab: a 16-byte aligned bias table is required");
    return CVIT_ERR_INVALID;
  }
  return conv3d_halo_impl(x, w_img, bias_table + 63 * Cout_pad, bias_table, out, D, H, W, Cin, Cout_pad, Cout_valid, dil, 1, nullptr, stream);
}

static int conv3d_halo_impl(const void* x, const void* w_img, const float* bias, const float* bias_table, void* out, int64_t D,
                           int64_t H, int64_t W, int64_t Cin, int64_t Cout_pad, int64_t Cout_valid, int64_t dil, int act,
                           void* aux, void* stream) {
  if (!x || !w_img || !bias || !out || D <= 0 || H <= 0 || W <= 0 || dil <= 0 || Cout_valid <= 0 || Cout_valid > Cout_pad ||
      (Cout_valid % 8) != 0) {
    set_error("conv3d_halo: bad arguments (D=%lld H=%lld W=%lld Cin=%lld Cout=%lld/%lld dil=%lld)", (long long)D, (long long)H,
              (long long)W, (long long)Cin, (long long)Cout_valid, (long long)Cout_pad, (long long)dil);
    return CVIT_ERR_INVALID;
  }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_img) | reinterpret_cast<uintptr_t>(out)) & 15u) {
    set_error("conv3d_halo: x, w_img and out must be 16-byte aligned");
    return CVIT_ERR_INVALID;
  }
  HaloArgs a;
  a.w_img = static_cast<const __nv_bfloat16*>(w_img);
  a.bias = bias;
  a.out = static_cast<__nv_bfloat16*>(out);
  a.D = (int)D;
  a.H = (int)H;
  a.W = (int)W;
  a.dil = (int)dil;
  a.n_valid = (int)Cout_valid;
  a.act = act;
  a.aux = static_cast<__nv_bfloat16*>(aux);
  a.bias_table = bias_table;
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 32 && Cout_pad == 32) return launch_halo<32, 32>(x, a, st);
  if (Cin == 32 && Cout_pad == 16) return launch_halo<32, 16>(x, a, st);
  if (Cin == 16 && Cout_pad == 16) return launch_halo<16, 16>(x, a, st);
  if (Cin == 16 && Cout_pad == 32) return launch_halo<16, 32>(x, a, st);
  if (Cin == 8 && Cout_pad == 16) return launch_halo<8, 16>(x, a, st);
  set_error("conv3d_halo: no kernel for Cin=%lld Cout_pad=%lld (narrow layers: Cin 8/16/32, Cout 16/32)", (long long)Cin,
            (long long)Cout_pad);
  return CVIT_ERR_UNSUPPORTED;
}
