// 3x3x3 "same" convolutions of the 8-channel, full-resolution end of the CryoVIT head (output_layer.0: 8 -> 8 + GELU,
// output_layer.2: 8 -> 1 + clip + sigmoid; models/cryovit.py:30-34,39,49) on tcgen05 with the OUTPUT VOXELS OF A ROW
// PACKED INTO THE MMA N DIMENSION.
//
// Why. With 8 channels the per-tap formulation of conv_halo.cu needs 18 MMAs (M = 128, N = 16, K = 16) per 128 voxels,
// and an M = 128 MMA costs ~65-70 cycles however small N is (the A tile is re-read from shared memory every time):
// 1.2 ms per 128 x 512 x 512 volume, tensor pipe 12 % busy (profiles/r01_ncu_full_v8.md). Here one MMA row is a GROUP of
// P consecutive voxels along W, K runs over the group's input window ((P + 2) voxels x 8 channels) and N over the
// group's P x Cout outputs; the in-plane column taps are folded into a banded weight matrix
//
//     B[(j_out, co)][(j_in, ci)] = w[kd][kh][kw = j_in - j_out][co][ci]   if 0 <= kw <= 2, else 0,
//
// so a (kd, kh) pair costs (P + 2) / 2 MMAs for 128 x P voxels: 45 MMAs per 1024 voxels for 8 -> 8 (P = 8, N = 64) and
// 81 per 2048 voxels for 8 -> 1 (P = 16, N = 16) instead of 144 / 288. The banded matrix multiplies (P + 2) / 3 times
// more zeros than the convolution needs -- irrelevant, these layers are nowhere near the tensor pipe's rate.
//
// Shared-memory image of one staged depth plane (tile = 16 rows x 8 groups, one halo row above and below):
//     slab j (window voxel j - 1 of every group):  [18 rows h][8 groups g][8 channels = 16 B]      (+16 B pad)
// A operand of (kh, K step s): rows m = (h, g) are 16 B apart (canonical K-major SWIZZLE_NONE: 8 rows = one 128 B core
// matrix, SBO = 128 B), the two 8-channel K chunks are slabs 2s and 2s + 1 (LBO = slab pitch), the row tap kh is the
// start address (+ kh * 8 * 16 B). Nothing is re-laid-out per tap.
// Staging: four loader warps copy each input voxel (16 B) of the plane from global memory straight to its slab
// position(s) with cp.async (zero fill = "same" padding), enumerated so that a warp reads consecutive voxels; the two
// voxels shared by neighbouring groups are copied twice (10 / 8 of the plane). A plane is one pipeline stage; the
// banded weights of all 27 (kd, kh, kw) taps (92 KB / 41 KB) stay resident.
//
// One CTA per SM, persistent over tiles, 288 threads: warps 0-3 loaders, warps 4-7 epilogue (thread = MMA row: P voxels
// x Cout values -> bias + GELU -> bf16, or + clip / sigmoid -> fp32; 128 / 64 contiguous bytes per thread), warp 8 MMA
// issuer (warp-uniform, elect.sync) and TMEM owner. Accumulators double-buffered in TMEM.
#include "ptx.cuh"
#include "tmap.h"

namespace cvit {

constexpr int WP_G = 8, WP_TH = 16;                 // MMA row = (h, g): 16 x 8 = 128
constexpr int WP_ROWS = WP_TH + 2;                  // staged rows (1-voxel halo)
constexpr int WP_SLAB = WP_ROWS * WP_G * 16 + 16;   // bytes; +16 so the 8 slabs a quarter-warp writes hit distinct banks
constexpr int WP_THREADS = 288;
constexpr int WP_LOADERS = 128;
constexpr int WP_LAG = 2;                           // cp.async groups kept in flight per loader thread

template <int P, int COUT>
struct WpCfg {
  static constexpr int N = P * COUT;                // MMA N
  static constexpr int NS = P + 2;                  // 16-byte K chunks (window voxels) per row
  static constexpr int KSTEPS = NS / 2;             // MMAs (K = 16) per (kd, kh)
  static constexpr int PLANE = NS * WP_SLAB;
  static constexpr int W_BYTES = 9 * KSTEPS * 2 * N * 16;   // [kd*3+kh][K step][2 chunks][N][16 B]
  static constexpr int STAGES_RAW = (232448 - 1024 - 256 - W_BYTES) / PLANE;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM = STAGES * PLANE + W_BYTES + 256 + 1024;
  static constexpr int PIECES = NS * WP_ROWS * WP_G;  // 16-byte copies per plane
  static constexpr int PER_THREAD = (PIECES + WP_LOADERS - 1) / WP_LOADERS;
  static constexpr int TMEM_COLS = 2 * N <= 32 ? 32 : 2 * N <= 64 ? 64 : 2 * N <= 128 ? 128 : 256;
  static constexpr int TW = WP_G * P;               // tile width in voxels
  static_assert(NS % 2 == 0 && N % 16 == 0 && N <= 128, "P + 2 even; MMA N a multiple of 16");
  static_assert(STAGES > WP_LAG, "the loaders signal a plane LAG planes late: needs more stages than that");
  static_assert(PLANE < 65536, "slab offsets are packed into 16 bits");
};

struct WpArgs {
  const __nv_bfloat16* x;      // [D, H, W, 8]
  const __nv_bfloat16* w_img;  // WpCfg::W_BYTES, host-arranged (cryovit_b200.head.wpack_weight_image)
  const float* bias;           // [N]: bias[j_out * COUT + co] = b[co]
  __nv_bfloat16* out;          // GELU mode: [D, H, W, 8]
  float* logits;               // final mode: [D, H, W] clipped logits (may be null)
  float* probs;                // final mode: [D, H, W] sigmoid of the clipped logits (may be null)
  int D, H, W, act;
};

__device__ __forceinline__ uint64_t wp_desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return static_cast<uint64_t>((addr & 0x3ffffu) >> 4) | (static_cast<uint64_t>(lbo >> 4) << 16) |
         (static_cast<uint64_t>(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int P, int COUT, bool FINAL>
__global__ void __launch_bounds__(WP_THREADS, 1) conv3d_wpack_kernel(const WpArgs args) {
  using Cfg = WpCfg<P, COUT>;
  constexpr int STAGES = Cfg::STAGES, N = Cfg::N;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sIn = smem_base;
  const uint32_t sW = smem_base + STAGES * Cfg::PLANE;
  const uint32_t sBar = (sW + Cfg::W_BYTES + 15u) & ~15u;
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * STAGES;
  const uint32_t bar_tfull = sBar + 16 * STAGES, bar_tempty = bar_tfull + 16, tmem_slot = bar_tempty + 16;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_w = (args.W + Cfg::TW - 1) / Cfg::TW, tiles_h = (args.H + WP_TH - 1) / WP_TH;
  const int per_plane = tiles_w * tiles_h;
  const int num_tiles = args.D * per_plane;

  {  // resident banded weights: one cooperative copy of the host-arranged image
    const uint4* src = reinterpret_cast<const uint4*>(args.w_img);
    uint4* dst = reinterpret_cast<uint4*>(smem_gen + (sW - smem_base));
    for (int i = threadIdx.x; i < Cfg::W_BYTES / 16; i += WP_THREADS) dst[i] = __ldg(src + i);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, WP_LOADERS / 32);  // one arrival per loader warp
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 4);
    }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  fence_proxy_async_smem();  // generic-proxy weight stores -> visible to the tensor core (async proxy)
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  auto tile_of = [&](int tile, int& d, int& h0, int& w0) {
    d = tile / per_plane;
    const int r = tile - d * per_plane;
    const int th = r / tiles_w;
    h0 = th * WP_TH;
    w0 = (r - th * tiles_w) * Cfg::TW;
  };

  if (warp < 4) {
    // ------------------------------------------------------------------ loaders: global -> slab positions (cp.async)
    // piece q = (row = h * 8 + g, window chunk j), j fastest: consecutive lanes read consecutive voxels.
    // Packed per piece, tile-invariant: slab byte offset [0,16) | staged row h [16,21) | tile column + 1 [21,30).
    uint32_t tab[Cfg::PER_THREAD];
#pragma unroll
    for (int i = 0; i < Cfg::PER_THREAD; ++i) {
      const int q = threadIdx.x + i * WP_LOADERS;
      if (q < Cfg::PIECES) {
        const int row = q / Cfg::NS, j = q - row * Cfg::NS;
        const int h = row / WP_G, g = row - h * WP_G;
        tab[i] = static_cast<uint32_t>(j * WP_SLAB + row * 16) | (static_cast<uint32_t>(h) << 16) |
                 (static_cast<uint32_t>(P * g + j) << 21);
      } else {
        tab[i] = 0xffffffffu;
      }
    }
    int s = 0, sig_s = 0, in_flight = 0;
    uint32_t ph = 0;
    auto signal_oldest = [&]() {
      fence_proxy_async_smem();  // this thread's completed cp.async writes -> async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * sig_s);
      if (++sig_s == STAGES) sig_s = 0;
      --in_flight;
    };
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int d, h0, w0;
      tile_of(tile, d, h0, w0);
      for (int kd = 0; kd < 3; ++kd) {
        const int dz = d + kd - 1;
        if (dz < 0 || dz >= args.D) continue;  // the whole depth tap is zero padding
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        const uint32_t dst0 = sIn + s * Cfg::PLANE;
        const __nv_bfloat16* src_plane = args.x + (int64_t)dz * args.H * args.W * 8;
#pragma unroll
        for (int i = 0; i < Cfg::PER_THREAD; ++i) {
          const uint32_t e = tab[i];
          if (e == 0xffffffffu) continue;
          const int hh = h0 - 1 + static_cast<int>((e >> 16) & 31u);
          const int ww = w0 - 1 + static_cast<int>(e >> 21);
          const bool ok = hh >= 0 && hh < args.H && ww >= 0 && ww < args.W;
          const __nv_bfloat16* src = ok ? src_plane + ((int64_t)hh * args.W + ww) * 8 : args.x;
          cp_async16_zfill(dst0 + (e & 0xffffu), src, ok ? 16u : 0u);
        }
        cp_async_commit();
        if (++in_flight > WP_LAG) {
          cp_async_wait<WP_LAG>();
          signal_oldest();
        }
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
    cp_async_wait<0>();
    while (in_flight > 0) signal_oldest();
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform, one elected lane)
    constexpr uint32_t idesc = umma_idesc_bf16_f32(128, N);
    constexpr uint32_t B_MMA = 2 * N * 16;
    int s = 0, acc = 0;
    uint32_t ph = 0, acc_ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int d = tile / per_plane;
      mbar_wait(bar_tempty + 8 * acc, acc_ph ^ 1u);
      const uint32_t d_tmem = tmem_base + acc * N;
      uint32_t accumulate = 0;
      for (int kd = 0; kd < 3; ++kd) {
        const int dz = d + kd - 1;
        if (dz < 0 || dz >= args.D) continue;
        mbar_wait(bar_full + 8 * s, ph);
        tcgen05_fence_after();
        if (elect_one_sync()) {
          const uint32_t plane = sIn + s * Cfg::PLANE;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int st = 0; st < Cfg::KSTEPS; ++st) {
              const uint32_t a = plane + 2 * st * WP_SLAB + kh * WP_G * 16;
              const uint32_t b = sW + ((kd * 3 + kh) * Cfg::KSTEPS + st) * B_MMA;
              umma_bf16(d_tmem, wp_desc_nosw(a, WP_SLAB, 128), wp_desc_nosw(b, N * 16, 128), idesc, accumulate);
              accumulate = 1;
            }
          }
          umma_commit(bar_empty + 8 * s);  // staged plane reusable
        }
        __syncwarp();
        accumulate = 1;
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      if (elect_one_sync()) umma_commit(bar_tfull + 8 * acc);  // accumulator complete
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1u;
    }
  } else {
    // ------------------------------------------------------------------ epilogue: thread == MMA row == P voxels
    const int q = warp & 3;  // TMEM lane quarter (hardware rule: warp id % 4)
    const int r = q * 32 + lane;
    const int hl = r / WP_G, g = r - hl * WP_G;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int d, h0, w0;
      tile_of(tile, d, h0, w0);
      mbar_wait(bar_tfull + 8 * acc, acc_ph);
      tcgen05_fence_after();
      const uint32_t t_acc = tmem_base + acc * N + (static_cast<uint32_t>(q * 32) << 16);
      uint32_t v[N];
      if (N == 16) {
        tmem_ld_32x16(t_acc, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
      } else {
#pragma unroll
        for (int c = 0; c < N; c += 32) tmem_ld_32x32(t_acc + c, *reinterpret_cast<uint32_t(*)[32]>(&v[c]));
      }
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);  // the accumulator is in registers
      const int h = h0 + hl, w = w0 + P * g;
      if (h < args.H && w < args.W) {  // W is a multiple of P: a group is inside or outside as a whole
        const int64_t vox = ((int64_t)d * args.H + h) * args.W + w;
        if (FINAL) {
          // COUT == 1: P clipped logits (and their sigmoid), fp32, contiguous along W
          float lg[P];
#pragma unroll
          for (int j = 0; j < P; ++j) lg[j] = fminf(fmaxf(__uint_as_float(v[j]) + __ldg(args.bias + j), -5.0f), 5.0f);
          if (args.logits) {
#pragma unroll
            for (int j = 0; j < P; j += 4)
              *reinterpret_cast<float4*>(args.logits + vox + j) = make_float4(lg[j], lg[j + 1], lg[j + 2], lg[j + 3]);
          }
          if (args.probs) {
#pragma unroll
            for (int j = 0; j < P; j += 4) {
              float4 p;
              p.x = 1.0f / (1.0f + __expf(-lg[j]));
              p.y = 1.0f / (1.0f + __expf(-lg[j + 1]));
              p.z = 1.0f / (1.0f + __expf(-lg[j + 2]));
              p.w = 1.0f / (1.0f + __expf(-lg[j + 3]));
              *reinterpret_cast<float4*>(args.probs + vox + j) = p;
            }
          }
        } else {
          // COUT == 8: P voxels x 8 channels, bias + GELU -> bf16, 16 B per voxel, P * 16 contiguous bytes
          __nv_bfloat16* o = args.out + vox * 8;
          float bb[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) bb[c] = __ldg(args.bias + c);
#pragma unroll
          for (int j = 0; j < P; ++j) {
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float a = __uint_as_float(v[j * 8 + 2 * i]) + bb[2 * i];
              float b = __uint_as_float(v[j * 8 + 2 * i + 1]) + bb[2 * i + 1];
              if (args.act) { a = gelu_erf(a); b = gelu_erf(b); }
              pk[i] = pack_bf16x2(a, b);
            }
            *reinterpret_cast<uint4*>(o + j * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1u;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int P, int COUT, bool FINAL>
static int launch_wpack(const WpArgs& a, cudaStream_t stream) {
  using Cfg = WpCfg<P, COUT>;
  auto kern = conv3d_wpack_kernel<P, COUT, FINAL>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) {
      set_error("conv3d_wpack: cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM, cudaGetErrorString(e));
      return CVIT_ERR_CUDA;
    }
    configured = true;
  }
  const int num_tiles = a.D * ((a.H + WP_TH - 1) / WP_TH) * ((a.W + Cfg::TW - 1) / Cfg::TW);
  int grid = num_sms();
  if (grid > num_tiles) grid = num_tiles;
  kern<<<grid, WP_THREADS, Cfg::SMEM, stream>>>(a);
  return check_launch("conv3d_wpack_kernel");
}

}  // namespace cvit

using namespace cvit;

// Bytes of the banded weight image of the two supported layers: (P = 8, Cout = 8) and (P = 16, Cout = 1).
extern "C" int64_t cvit_conv3d_wpack_weight_bytes(int64_t P, int64_t Cout) {
  if (P == 8 && Cout == 8) return WpCfg<8, 8>::W_BYTES;
  if (P == 16 && Cout == 1) return WpCfg<16, 1>::W_BYTES;
  return -1;
}

static int wpack_check(const void* x, const void* w_img, const float* bias, int64_t D, int64_t H, int64_t W, int64_t P) {
  if (!x || !w_img || !bias || D <= 0 || H <= 0 || W <= 0) {
    set_error("conv3d_wpack: bad arguments (D=%lld H=%lld W=%lld)", (long long)D, (long long)H, (long long)W);
    return CVIT_ERR_INVALID;
  }
  if (W % P != 0) {
    set_error("conv3d_wpack: W=%lld must be a multiple of %lld (use cvit_conv3d_halo_ndhwc / cvit_head_out_conv otherwise)",
              (long long)W, (long long)P);
    return CVIT_ERR_UNSUPPORTED;
  }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_img)) & 15u) {
    set_error("conv3d_wpack: x and w_img must be 16-byte aligned");
    return CVIT_ERR_INVALID;
  }
  return CVIT_OK;
}

extern "C" int cvit_conv3d_wpack8_gelu(const void* x, const void* w_img, const float* bias_n, void* out, int64_t D, int64_t H,
                                       int64_t W, int act, void* stream) {
  int rc = wpack_check(x, w_img, bias_n, D, H, W, 8);
  if (rc) return rc;
  if (!out || (reinterpret_cast<uintptr_t>(out) & 15u)) {
    set_error("conv3d_wpack8_gelu: out must be a 16-byte aligned bf16 [D,H,W,8] buffer");
    return CVIT_ERR_INVALID;
  }
  WpArgs a;
  a.x = static_cast<const __nv_bfloat16*>(x);
  a.w_img = static_cast<const __nv_bfloat16*>(w_img);
  a.bias = bias_n;
  a.out = static_cast<__nv_bfloat16*>(out);
  a.logits = nullptr;
  a.probs = nullptr;
  a.D = (int)D;
  a.H = (int)H;
  a.W = (int)W;
  a.act = act;
  return launch_wpack<8, 8, false>(a, (cudaStream_t)stream);
}

extern "C" int cvit_conv3d_wpack8_final(const void* x, const void* w_img, const float* bias_n, float* logits, float* probs,
                                        int64_t D, int64_t H, int64_t W, void* stream) {
  int rc = wpack_check(x, w_img, bias_n, D, H, W, 16);
  if (rc) return rc;
  if ((!logits && !probs) || ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(probs)) & 15u)) {
    set_error("conv3d_wpack8_final: need logits and/or probs, 16-byte aligned fp32 [D,H,W]");
    return CVIT_ERR_INVALID;
  }
  WpArgs a;
  a.x = static_cast<const __nv_bfloat16*>(x);
  a.w_img = static_cast<const __nv_bfloat16*>(w_img);
  a.bias = bias_n;
  a.out = nullptr;
  a.logits = logits;
  a.probs = probs;
  a.D = (int)D;
  a.H = (int)H;
  a.W = (int)W;
  a.act = 0;
  return launch_wpack<16, 1, true>(a, (cudaStream_t)stream);
}
