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
// Shared-memory image of one staged depth plane (tile = 16 rows x 8 groups, one halo row above and below; staged row
// index m' = h * 8 + g, 144 of them): the window's 16-byte chunks j = 0 .. P+1 (window voxel j - 1 of every group) are
//     chunks 0 .. P-1 : P / 8 atoms of the K-major SWIZZLE_128B layout -- row m' is 128 B, chunk c of the atom sits at
//                       16-byte slot c ^ (m' & 7) (exactly what TMA's 128B swizzle would produce, and what the ViT
//                       GEMMs feed the tensor core), so the 8 consecutive voxels a quarter-warp copies fill one 128 B
//                       row conflict-free AND every core matrix the MMA reads is one aligned 128 B line;
//     chunks P, P+1   : two SWIZZLE_NONE slabs [144 rows][16 B] (LBO = slab pitch, SBO = 128 B).
// The row tap kh is the descriptor start address (+ kh * 8 rows = 1024 B resp. 128 B, which keeps the swizzle phase),
// a K step inside an atom is +32 B. Nothing is re-laid-out per tap. (A first version used P + 2 un-swizzled slabs
// with a 16 B skew against the loaders' bank conflicts: the skew mis-aligned every core matrix read and doubled the
// cost of an MMA -- 0.76 ms instead of the per-tap kernel's 1.16 ms, where ~0.4 was expected.)
// Staging: loader warps copy each input voxel (16 B) of a plane from global memory straight to its position(s) with
// cp.async (zero fill = "same" padding), enumerated so that a warp reads consecutive voxels; the two voxels shared by
// neighbouring groups are copied twice (10 / 8 of the plane). A plane is one pipeline stage, and a CTA walks ALONG
// DEPTH through its tiles, so every plane is staged once per CTA, not once per depth tap. The banded weights of all
// 27 (kd, kh, kw) taps (92 KB / 41 KB) stay resident.
//
// INPUT-STATIONARY along depth. A staged input plane p feeds three outputs: depth p+1 through the kd = 0 weights, p
// through kd = 1 and p-1 through kd = 2. The banded weights of the three depth taps are stacked along N in that order
// ([kd=2 | kd=1 | kd=0] blocks of N rows) and the accumulators of consecutive output depths sit in consecutive N-column
// slots of a ring of 8 in TMEM, so ONE MMA (N_total = 3 N) adds a plane's contribution to all three outputs: every
// plane is multiplied once instead of three times (15 resp. 27 MMAs per plane-tile plus a few for ring wrap-around and
// for the first tap, where a fresh accumulator needs its own accumulate = 0 MMA). An output is complete after the
// plane above it; its slot is handed to the epilogue with tcgen05.commit and comes back on a per-slot barrier.
//
// One CTA per SM, persistent over tiles, 416 threads: warps 0-3 loaders, warps 4-11 epilogue (thread = MMA row; the two
// warps of a TMEM lane quarter split the row's P voxels: bias + GELU -> bf16, 64 contiguous bytes per thread -- the
// 8 -> 8 layer is bound by exactly this epilogue math; the 8 -> 1 layer's clip / sigmoid -> fp32 needs one warp per
// quarter only), warp 12 MMA issuer (warp-uniform, elect.sync) and TMEM owner.
#include "ptx.cuh"
#include "tmap.h"

namespace cvit {

constexpr int WP_G = 8, WP_TH = 16;                 // MMA row = (h, g): 16 x 8 = 128
constexpr int WP_ROWS = WP_TH + 2;                  // staged rows (1-voxel halo)
constexpr int WP_MROWS = WP_ROWS * WP_G;            // staged MMA rows per plane (144)
constexpr int WP_ATOM = WP_MROWS * 128;             // one SWIZZLE_128B atom: 8 K chunks of every staged row
constexpr int WP_SLAB = WP_MROWS * 16;              // one un-swizzled slab: one K chunk of every staged row
constexpr int WP_THREADS = 416;                     // warps 0-3 loaders, 4-11 epilogue, 12 MMA issuer
constexpr int WP_LOADERS = 128;

template <int P, int COUT>
struct WpCfg {
  static constexpr int N = P * COUT;                // MMA N
  static constexpr int NS = P + 2;                  // 16-byte K chunks (window voxels) per row
  static constexpr int KSTEPS = NS / 2;             // MMAs (K = 16) per (kd, kh)
  static constexpr int ATOMS = P / 8;              // swizzled atoms; the last two chunks are un-swizzled slabs
  static constexpr int TAIL = ATOMS * WP_ATOM;     // byte offset of the two slabs
  static constexpr int PLANE = (TAIL + 2 * WP_SLAB + 1023) / 1024 * 1024;  // planes stay 1024-aligned (swizzle phase)
  static constexpr int W_BYTES = 9 * KSTEPS * 2 * N * 16;   // [kh][K step][2 chunks][3 blocks: kd = 2, 1, 0][N][16 B]
  static constexpr int RING = 8;                    // accumulator slots (consecutive output depths) in TMEM
  static constexpr int STAGES_RAW = (232448 - 1024 - 256 - W_BYTES) / PLANE;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM = STAGES * PLANE + W_BYTES + 256 + 1024;
  static constexpr int PIECES = NS * WP_ROWS * WP_G;  // 16-byte copies per plane
  static constexpr int PER_THREAD = (PIECES + WP_LOADERS - 1) / WP_LOADERS;
  static constexpr int TMEM_COLS = RING * N <= 32 ? 32 : RING * N <= 64 ? 64 : RING * N <= 128 ? 128 : RING * N <= 256 ? 256 : 512;
  static_assert(RING * N <= 512, "accumulator ring exceeds TMEM");
  static constexpr int TW = WP_G * P;               // tile width in voxels
  static_assert(P % 8 == 0 && N % 16 == 0 && N <= 128, "whole swizzle atoms; MMA N a multiple of 16");
  static_assert(STAGES >= 3, "one plane under the MMAs, one landing, one being issued");
  static_assert(PLANE < 65536, "slab offsets are packed into 16 bits");
};

struct WpArgs {
  const __nv_bfloat16* x;      // [D, H, W, 8]
  const __nv_bfloat16* w_img;  // WpCfg::W_BYTES, host-arranged (cryovit_b200.head.wpack_weight_image)
  const float* bias;           // [N]: bias[j_out * COUT + co] = b[co]
  __nv_bfloat16* out;          // GELU mode: [D, H, W, 8]
  __nv_bfloat16* aux;          // act = ACT_DUAL: gelu(out); ACT_GELU_GRAD: the pre-activation whose gelu' scales the result
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
  const uint32_t bar_ofull = sBar + 16 * STAGES, bar_oempty = bar_ofull + 8 * Cfg::RING, tmem_slot = bar_oempty + 8 * Cfg::RING;
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
    for (int i = 0; i < Cfg::RING; ++i) {
      mbar_init(bar_ofull + 8 * i, 1);    // MMA -> epilogue: the output of this slot is complete
      mbar_init(bar_oempty + 8 * i, FINAL ? 4 : 8);  // epilogue warps -> MMA: the slot is in registers, may be re-initialised
    }
    fence_mbar_init();
  }
  if (warp == 12) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  fence_proxy_async_smem();  // generic-proxy weight stores -> visible to the tensor core (async proxy)
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  // Work = (spatial tile "column", depth) pairs, depth fastest; every CTA owns one contiguous range of them, cut into
  // SEGMENTS at column changes. Inside a segment the CTA walks along depth: planes d_a-1 .. d_b+1 are staged once each,
  // in order, and each is consumed once (input-stationary MMAs, see the header). Planes are numbered in load order
  // (n = 0, 1, ...) by every role; plane n lives in ring slot n % STAGES with barrier parity (n / STAGES) & 1. Outputs
  // are numbered in processing order too (seq); output seq accumulates in TMEM slot seq % RING.
  const int64_t total = (int64_t)per_plane * args.D;
  const int i_begin = (int)(total * blockIdx.x / gridDim.x), i_end = (int)(total * (blockIdx.x + 1) / gridDim.x);
  auto col_origin = [&](int col, int& h0, int& w0) {
    const int th = col / tiles_w;
    h0 = th * WP_TH;
    w0 = (col - th * tiles_w) * Cfg::TW;
  };
  (void)num_tiles;

  if (warp < 4) {
    // ------------------------------------------------------------------ loaders: global -> staged positions (cp.async)
    // All four warps copy every plane; a thread signals plane n-1 (cp.async.wait_group 1, proxy fence, one arrival per
    // warp) right after issuing plane n, so two plane loads are in flight. Plane n reuses the slot of plane
    // n - STAGES, which the MMAs released at least one output earlier, so the one-plane lag cannot dead-lock.
    // piece q = (row = h * 8 + g, window chunk j), j fastest: consecutive lanes read consecutive voxels.
    // Packed per piece, tile-invariant: staged byte offset [0,16) | staged row h [16,21) | tile column + 1 [21,30).
    uint32_t tab[Cfg::PER_THREAD];
#pragma unroll
    for (int i = 0; i < Cfg::PER_THREAD; ++i) {
      const int q = threadIdx.x + i * WP_LOADERS;
      if (q < Cfg::PIECES) {
        const int row = q / Cfg::NS, j = q - row * Cfg::NS;
        const int h = row / WP_G, g = row - h * WP_G;
        const int dst = j < P ? (j >> 3) * WP_ATOM + row * 128 + (((j & 7) ^ (row & 7)) << 4)
                              : Cfg::TAIL + (j - P) * WP_SLAB + row * 16;
        tab[i] = static_cast<uint32_t>(dst) | (static_cast<uint32_t>(h) << 16) | (static_cast<uint32_t>(P * g + j) << 21);
      } else {
        tab[i] = 0xffffffffu;
      }
    }
    int n = 0;
    for (int i = i_begin; i < i_end;) {
      const int col = i / args.D, d_a = i - col * args.D;
      const int d_b = min(args.D - 1, d_a + (i_end - i) - 1);
      int h0, w0;
      col_origin(col, h0, w0);
      const int p_a = max(d_a - 1, 0), p_b = min(d_b + 1, args.D - 1);
      for (int pz = p_a; pz <= p_b; ++pz, ++n) {
        const int slot = n % STAGES;
        mbar_wait(bar_empty + 8 * slot, (((n / STAGES) & 1) ^ 1) & 1);
        const uint32_t dst0 = sIn + slot * Cfg::PLANE;
        const __nv_bfloat16* src_plane = args.x + (int64_t)pz * args.H * args.W * 8;
#pragma unroll
        for (int k = 0; k < Cfg::PER_THREAD; ++k) {
          const uint32_t e = tab[k];
          if (e == 0xffffffffu) continue;
          const int hh = h0 - 1 + static_cast<int>((e >> 16) & 31u);
          const int ww = w0 - 1 + static_cast<int>(e >> 21);
          const bool ok = hh >= 0 && hh < args.H && ww >= 0 && ww < args.W;
          const __nv_bfloat16* src = ok ? src_plane + ((int64_t)hh * args.W + ww) * 8 : args.x;
          cp_async16_zfill(dst0 + (e & 0xffffu), src, ok ? 16u : 0u);
        }
        cp_async_commit();
        if (n > 0) {
          cp_async_wait<1>();        // plane n-1 has landed
          fence_proxy_async_smem();  // this thread's completed cp.async writes -> async proxy
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_full + 8 * ((n - 1) % STAGES));
        }
      }
      i += d_b - d_a + 1;
    }
    if (n > 0) {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * ((n - 1) % STAGES));
    }
  } else if (warp == 12) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform, one elected lane)
    constexpr int R = Cfg::RING;
    constexpr uint32_t idesc1 = umma_idesc_bf16_f32(128, N), idesc2 = umma_idesc_bf16_f32(128, 2 * N),
                       idesc3 = umma_idesc_bf16_f32(128, 3 * N);
    constexpr uint32_t B_MMA = 2 * 3 * N * 16, B_LBO = 3 * N * 16;
    int n = 0, seq_base = 0;  // n: planes consumed (ring slot n % STAGES); seq_base: outputs before this segment
    for (int i = i_begin; i < i_end;) {
      const int col = i / args.D, d_a = i - col * args.D;
      const int d_b = min(args.D - 1, d_a + (i_end - i) - 1);
      const int p_a = max(d_a - 1, 0), p_b = min(d_b + 1, args.D - 1);
      for (int pz = p_a; pz <= p_b; ++pz, ++n) {
        // wanted outputs of this plane (consecutive depths), their ring slots and whether this is their first plane
        const int o_lo = max(pz - 1, d_a), o_hi = min(pz + 1, d_b);
        for (int o = o_lo; o <= o_hi; ++o) {
          if (pz == max(o - 1, p_a)) {  // fresh accumulator: the epilogue must have drained the slot's previous output
            const int seq = seq_base + o - d_a;
            mbar_wait(bar_oempty + 8 * (seq % R), (((seq / R) & 1) ^ 1) & 1);
          }
        }
        mbar_wait(bar_full + 8 * (n % STAGES), (n / STAGES) & 1);
        tcgen05_fence_after();
        if (elect_one_sync()) {
          const uint32_t plane = sIn + (n % STAGES) * Cfg::PLANE;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int st = 0; st < Cfg::KSTEPS; ++st) {
              uint64_t adesc;
              if (st < 4 * Cfg::ATOMS)  // chunks 2 st, 2 st + 1 of atom st / 4: +32 B per K step inside the swizzle span
                adesc = umma_smem_desc_kmajor<128>(plane + (st >> 2) * WP_ATOM + kh * WP_G * 128) + 2 * (st & 3);
              else
                adesc = wp_desc_nosw(plane + Cfg::TAIL + kh * WP_G * 16, WP_SLAB, 128);
              const uint32_t b0 = sW + (kh * Cfg::KSTEPS + st) * B_MMA;
              if (kh == 0 && st == 0) {
                // first tap of the plane: one MMA per output so that fresh accumulators start from zero
                for (int o = o_lo; o <= o_hi; ++o) {
                  const int slot = (seq_base + o - d_a) % R, blk = o - (pz - 1);
                  umma_bf16(tmem_base + slot * N, adesc, wp_desc_nosw(b0 + blk * N * 16, B_LBO, 128), idesc1,
                            pz == max(o - 1, p_a) ? 0u : 1u);
                }
              } else {
                // one MMA per run of outputs whose slots are consecutive (a second one only where the ring wraps)
                int o = o_lo;
                while (o <= o_hi) {
                  const int slot = (seq_base + o - d_a) % R, blk = o - (pz - 1);
                  int len = 1;
                  while (o + len <= o_hi && slot + len < R) ++len;
                  umma_bf16(tmem_base + slot * N, adesc, wp_desc_nosw(b0 + blk * N * 16, B_LBO, 128),
                            len == 3 ? idesc3 : len == 2 ? idesc2 : idesc1, 1u);
                  o += len;
                }
              }
            }
          }
          umma_commit(bar_empty + 8 * (n % STAGES));  // the staged plane has served all three of its outputs
          // the output below this plane is complete; at the end of the segment so is the plane's own depth
          if (pz - 1 >= d_a) umma_commit(bar_ofull + 8 * ((seq_base + pz - 1 - d_a) % R));
          if (pz == p_b && pz <= d_b) umma_commit(bar_ofull + 8 * ((seq_base + pz - d_a) % R));
        }
        __syncwarp();
      }
      seq_base += d_b - d_a + 1;
      i += d_b - d_a + 1;
    }
  } else if (!(FINAL && warp >= 8)) {
    // ------------------------------------------------------------------ epilogue: thread == MMA row == P voxels
    const int q = warp & 3;  // TMEM lane quarter (hardware rule: warp id % 4)
    const int half = (warp - 4) >> 2;  // GELU mode: which half of the row's voxels (columns) this warp takes
    const int r = q * 32 + lane;
    const int hl = r / WP_G, g = r - hl * WP_G;
    // the bias lives in registers for the whole kernel (a per-tile __ldg put a global-load round trip in front of
    // every tile's first FADD: ncu source page, stall_long_sb)
    float bb[FINAL ? P : 8];
#pragma unroll
    for (int c = 0; c < (FINAL ? P : 8); ++c) bb[c] = __ldg(args.bias + c);
    for (int i = i_begin; i < i_end; ++i) {
      const int col = i / args.D, d = i - col * args.D;
      int h0, w0;
      col_origin(col, h0, w0);
      const int seq = i - i_begin, slot = seq % Cfg::RING;  // outputs take the ring slots in order
      mbar_wait(bar_ofull + 8 * slot, (seq / Cfg::RING) & 1);
      tcgen05_fence_after();
      const uint32_t t_acc = tmem_base + slot * N + (static_cast<uint32_t>(q * 32) << 16);
      constexpr int NV = FINAL ? N : N / 2;  // columns per epilogue warp
      uint32_t v[NV];
      if (FINAL) {
        tmem_ld_32x16(t_acc, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
      } else {
#pragma unroll
        for (int c = 0; c < NV; c += 32) tmem_ld_32x32(t_acc + half * NV + c, *reinterpret_cast<uint32_t(*)[32]>(&v[c]));
      }
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_oempty + 8 * slot);  // the accumulator is in registers
      const int h = h0 + hl, w = w0 + P * g;
      if (h < args.H && w < args.W) {  // W is a multiple of P: a group is inside or outside as a whole
        const int64_t vox = ((int64_t)d * args.H + h) * args.W + w;
        if (FINAL) {
          // COUT == 1: P clipped logits (and their sigmoid), fp32, contiguous along W
          float lg[P];
#pragma unroll
          for (int j = 0; j < P; ++j) lg[j] = fminf(fmaxf(__uint_as_float(v[j]) + bb[j], -5.0f), 5.0f);
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
          // COUT == 8: this warp's P / 2 voxels x 8 channels, bias + GELU -> bf16, 16 B per voxel
          __nv_bfloat16* o = args.out + (vox + half * (P / 2)) * 8;
          // the activation switch is hoisted: a branch inside the unrolled loops keeps the voxels' GELU chains from
          // interleaving (see the transposed-convolution epilogue in gemm_tcgen05.cuh)
          if (args.act == ACT_GELU_GRAD) {  // input gradient times gelu'(z) of the layer below (training)
            const __nv_bfloat16* zp = args.aux + (vox + half * (P / 2)) * 8;
            uint4 z4[P / 2];
#pragma unroll
            for (int j = 0; j < P / 2; ++j) z4[j] = __ldg(reinterpret_cast<const uint4*>(zp + j * 8));
#pragma unroll
            for (int j = 0; j < P / 2; ++j) {
              const uint32_t zz[4] = {z4[j].x, z4[j].y, z4[j].z, z4[j].w};
              uint32_t pk[4];
#pragma unroll
              for (int k = 0; k < 4; ++k)
                pk[k] = act_gelu_grad_pair(__uint_as_float(v[j * 8 + 2 * k]) + bb[2 * k], __uint_as_float(v[j * 8 + 2 * k + 1]) + bb[2 * k + 1], zz[k]);
              *reinterpret_cast<uint4*>(o + j * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          } else if (args.act == ACT_DUAL) {  // pre-activation -> out, activation -> aux (training forward)
            __nv_bfloat16* o2 = args.aux + (vox + half * (P / 2)) * 8;
#pragma unroll
            for (int j = 0; j < P / 2; ++j) {
              uint32_t pz[4], pk[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float a = __uint_as_float(v[j * 8 + 2 * k]) + bb[2 * k];
                float b = __uint_as_float(v[j * 8 + 2 * k + 1]) + bb[2 * k + 1];
                pz[k] = pack_bf16x2(a, b);
                gelu_erf2(a, b);
                pk[k] = pack_bf16x2(a, b);
              }
              *reinterpret_cast<uint4*>(o + j * 8) = make_uint4(pz[0], pz[1], pz[2], pz[3]);
              *reinterpret_cast<uint4*>(o2 + j * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          } else if (args.act) {
#pragma unroll
            for (int j = 0; j < P / 2; ++j) {
              uint32_t pk[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float a = __uint_as_float(v[j * 8 + 2 * k]) + bb[2 * k];
                float b = __uint_as_float(v[j * 8 + 2 * k + 1]) + bb[2 * k + 1];
                gelu_erf2(a, b);
                pk[k] = pack_bf16x2(a, b);
              }
              *reinterpret_cast<uint4*>(o + j * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < P / 2; ++j) {
              uint32_t pk[4];
#pragma unroll
              for (int k = 0; k < 4; ++k)
                pk[k] = pack_bf16x2(__uint_as_float(v[j * 8 + 2 * k]) + bb[2 * k], __uint_as_float(v[j * 8 + 2 * k + 1]) + bb[2 * k + 1]);
              *reinterpret_cast<uint4*>(o + j * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 12) {
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

extern "C" int cvit_conv3d_wpack8_aux(const void* x, const void* w_img, const float* bias_n, void* out, int64_t D, int64_t H,
                                      int64_t W, int act, void* aux, void* stream);

extern "C" int cvit_conv3d_wpack8_gelu(const void* x, const void* w_img, const float* bias_n, void* out, int64_t D, int64_t H,
                                       int64_t W, int act, void* stream) {
  return cvit_conv3d_wpack8_aux(x, w_img, bias_n, out, D, H, W, act ? 1 : 0, nullptr, stream);
}

// act: 0 none, 1 GELU, 2 out = pre-activation and aux = GELU of it, 3 out = result * gelu'(aux) (ptx.cuh ACT_*).
extern "C" int cvit_conv3d_wpack8_aux(const void* x, const void* w_img, const float* bias_n, void* out, int64_t D, int64_t H,
                                      int64_t W, int act, void* aux, void* stream) {
  int rc = wpack_check(x, w_img, bias_n, D, H, W, 8);
  if (rc) return rc;
  if (act < 0 || act > 3 || (act >= 2 && (!aux || (reinterpret_cast<uintptr_t>(aux) & 15u)))) {
    set_error("conv3d_wpack8: act=%d needs a 16-byte aligned aux (0 none, 1 GELU, 2 out=z aux=gelu(z), 3 out=y*gelu'(aux))", act);
    return CVIT_ERR_INVALID;
  }
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
  a.aux = static_cast<__nv_bfloat16*>(aux);
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
  a.aux = nullptr;
  return launch_wpack<16, 1, true>(a, (cudaStream_t)stream);
}
