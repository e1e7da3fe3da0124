// Flash-style self-attention of one ViT slice on tcgen05 / TMEM (head_dim 64, bf16 in/out, fp32 softmax).
// Restates upstream MemEffAttention: softmax(q k^T / 8) v, q,k,v = qkv.reshape(B, N, 3, H, 64).
//
// Persistent kernel: 2 CTAs per SM (80 KB smem, 256 TMEM columns each), each looping over work items
// (slice, head, 128-query tile), query tile fastest so co-running CTAs share K/V in L2. The two co-resident CTAs
// overlap one's tensor-core phases with the other's softmax; barrier phases, the K/V ring and the Q buffer run
// across items, so the next item's Q and first K/V tiles are prefetched during the current item's tail. 192 threads:
//   warp 0    TMA producer: Q tile once, then (K, V) tiles of 128 keys through a 2-stage ring. The tensor map is
//             3-D (column, token, slice) so tokens past the end of a slice are zero-filled, never the next slice.
//   warp 1    MMA issuer:  S = Q K^T     (SS: both operands K-major in 128B-swizzled smem, N = 128 keys)
//                          O += P V      (TS: P read from TMEM as packed bf16, V tile as an MN-major B operand)
//   warps 2-5 softmax, one thread per query row (TMEM lane): pass A row max (with the lazy-rescale rule: O and l
//             are only rescaled when the max grows by more than 2^8), pass B p = exp2(s*c - m*c) -> bf16 -> TMEM.
// TMEM columns: S fp32 [0,128) | P bf16x2 [128,192) | O fp32 [192,256).
// Ordering: the MMA warp issues PV(j) then S(j+1) and commits ONE barrier, so when the softmax warps see S(j+1)
// they also know PV(j) has retired: O may be rescaled and P overwritten without further synchronisation.
// The ragged tail (1029 = 8*128 + 5 keys) runs as an N=16 MMA with the 11 padding keys masked to -inf.
#include "ptx.cuh"
#include "tmap.h"

namespace cvit {

constexpr int FA_BQ = 128, FA_BK = 128, FA_D = 64;
constexpr int FA_THREADS = 192;
constexpr int FA_TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16
constexpr int FA_SMEM = 5 * FA_TILE_BYTES + 128 + 1024;
constexpr int FA_TMEM_COLS = 256;
constexpr uint32_t FA_COL_S = 0, FA_COL_P = 128, FA_COL_O = 192;

struct FaArgs {
  __nv_bfloat16* out;  // [B*T, C]
  int T, heads, C, slices;
  float scale_log2e;   // head_dim^-0.5 * log2(e)
};

__global__ void __launch_bounds__(FA_THREADS, 2)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, const FaArgs args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sK = smem_base + FA_TILE_BYTES;      // 2 stages
  const uint32_t sV = smem_base + 3 * FA_TILE_BYTES;  // 2 stages
  const uint32_t sBar = smem_base + 5 * FA_TILE_BYTES;
  const uint32_t bar_q = sBar, bar_kv_full = sBar + 8, bar_kv_empty = sBar + 24;
  const uint32_t bar_s = sBar + 40, bar_p = sBar + 48, bar_o = sBar + 56, bar_q_empty = sBar + 64;
  const uint32_t bar_o_empty = sBar + 72, tmem_slot = sBar + 80;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = args.T;
  const int n_tiles = (T + FA_BK - 1) / FA_BK;   // K/V tiles per item
  const int n_qt = (T + FA_BQ - 1) / FA_BQ;      // query tiles per (slice, head)
  const int total_items = n_qt * args.heads * args.slices;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(bar_q, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_kv_full + 8 * s, 1);
      mbar_init(bar_kv_empty + 8 * s, 1);
    }
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 4);
    mbar_init(bar_o, 1);
    mbar_init(bar_q_empty, 1);
    mbar_init(bar_o_empty, 4);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<FA_TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp == 0) {
    if (lane == 0) {
      uint32_t kv_it = 0, k = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++k) {
        const int qt = item % n_qt, bh = item / n_qt;
        const int head = bh % args.heads, slice = bh / args.heads;
        const int cq = head * FA_D, ck = args.C + head * FA_D, cv = 2 * args.C + head * FA_D;
        mbar_wait(bar_q_empty, (k & 1) ^ 1u);  // every S MMA of the previous item has retired
        mbar_arrive_expect_tx(bar_q, FA_TILE_BYTES);
        tma_load_3d(sQ, &tmQKV, bar_q, cq, qt * FA_BQ, slice);
        for (int j = 0; j < n_tiles; ++j, ++kv_it) {
          const uint32_t s = kv_it & 1;
          mbar_wait(bar_kv_empty + 8 * s, ((kv_it >> 1) & 1) ^ 1u);
          mbar_arrive_expect_tx(bar_kv_full + 8 * s, 2 * FA_TILE_BYTES);
          tma_load_3d(sK + s * FA_TILE_BYTES, &tmQKV, bar_kv_full + 8 * s, ck, j * FA_BK, slice);
          tma_load_3d(sV + s * FA_TILE_BYTES, &tmQKV, bar_kv_full + 8 * s, cv, j * FA_BK, slice);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t tS = tmem_base + FA_COL_S, tP = tmem_base + FA_COL_P, tO = tmem_base + FA_COL_O;
      constexpr uint32_t idesc_pv = umma_idesc_bf16_f32(FA_BQ, FA_D) | (1u << 16);  // B is MN-major
      auto n_mma_of = [&](int j) {  // keys of tile j rounded up to the MMA granularity (16)
        const int valid = min(FA_BK, T - j * FA_BK);
        return (valid + 15) & ~15;
      };
      uint32_t kv_it = 0, tile_it = 0, k = 0;
      auto issue_s = [&](int j, uint32_t kv) {
        const uint32_t s = kv & 1;
        mbar_wait(bar_kv_full + 8 * s, (kv >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t idesc_s = umma_idesc_bf16_f32(FA_BQ, n_mma_of(j));
        const uint64_t qd = umma_smem_desc_kmajor<128>(sQ), kd = umma_smem_desc_kmajor<128>(sK + s * FA_TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < FA_D / 16; ++kk) umma_bf16(tS, qd + 2 * kk, kd + 2 * kk, idesc_s, kk > 0);
      };
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++k) {
        mbar_wait(bar_q, k & 1);
        issue_s(0, kv_it);  // overlaps the previous item's epilogue: S was fully read when its last P was published
        umma_commit(bar_s);
        for (int j = 0; j < n_tiles; ++j, ++kv_it, ++tile_it) {
          const uint32_t s = kv_it & 1;
          mbar_wait(bar_p, tile_it & 1);  // P(j) written, S(j) fully read
          if (j == 0) mbar_wait(bar_o_empty, (k & 1) ^ 1u);  // previous item's O has been read out of TMEM
          tcgen05_fence_after();
          const int ksteps = n_mma_of(j) / 16;
          for (int kk = 0; kk < ksteps; ++kk) {
            // 16 keys per step: 8 packed TMEM columns of P, 16 rows (2 KB) of the V tile
            const uint64_t vd = umma_smem_desc_mnmajor_sw128(sV + s * FA_TILE_BYTES + kk * 2048, FA_TILE_BYTES);
            umma_bf16_ts(tO, tP + 8 * kk, vd, idesc_pv, (j | kk) != 0);
          }
          umma_commit(bar_kv_empty + 8 * s);  // K(j), V(j) free once these retire
          if (j + 1 < n_tiles) {
            issue_s(j + 1, kv_it + 1);
            umma_commit(bar_s);  // fires after PV(j) AND S(j+1)
          } else {
            umma_commit(bar_q_empty);
            umma_commit(bar_o);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps, thread == query row
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float c = args.scale_log2e;
    uint32_t tile_it = 0, k = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++k) {
    const int qt = item % n_qt, bh = item / n_qt;
    const bool warp_active = qt * FA_BQ + q * 32 < T;  // warp-uniform: all-padding warps only keep the barriers moving
    float m = 0.f, l = 0.f;
    for (int j = 0; j < n_tiles; ++j, ++tile_it) {
      mbar_wait(bar_s, tile_it & 1);
      tcgen05_fence_after();
      if (warp_active) {
        const int valid = min(FA_BK, T - j * FA_BK);
        if (valid == FA_BK) {
          // The whole 128-wide S row lives in registers: one TMEM read per tile, four loads in flight.
          uint32_t v[128];
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            tmem_ld_32x32(t_row + FA_COL_S + ch * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32 * ch]));
          tmem_ld_wait();
          // ---- row max (4 independent chains)
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 128; i += 4) {
            mx[0] = fmaxf(mx[0], __uint_as_float(v[i]));
            mx[1] = fmaxf(mx[1], __uint_as_float(v[i + 1]));
            mx[2] = fmaxf(mx[2], __uint_as_float(v[i + 2]));
            mx[3] = fmaxf(mx[3], __uint_as_float(v[i + 3]));
          }
          const float mt = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          if (j == 0) {
            m = mt;
          } else {
            const bool need = (mt - m) * c > 8.0f;
            if (__any_sync(0xffffffffu, need)) {
              const float f = need ? ex2_approx((m - mt) * c) : 1.0f;
#pragma unroll 1
              for (int hh = 0; hh < 4; ++hh) {  // 16 columns at a time: the S row keeps 128 registers busy
                uint32_t o[16];
                tmem_ld_32x16(t_row + FA_COL_O + hh * 16, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
                tmem_st_32x16(t_row + FA_COL_O + hh * 16, o);
              }
              l *= f;
              if (need) m = mt;
            }
          }
          // ---- probabilities, packed in place (element pair i -> register i), 4 partial row sums
          const float nmc = -m * c;
          float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), c, nmc));
            const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), c, nmc));
            ls[i & 3] += p0 + p1;
            v[i] = pack_bf16x2(p0, p1);
          }
          l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            tmem_st_32x16(t_row + FA_COL_P + ch * 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16 * ch]));
        } else {
          // ---- ragged last tile: `valid` keys inside an N = roundup16(valid) MMA, 16-column chunks
          const int nch = (valid + 15) >> 4;
          float mt = -INFINITY;
          for (int ch = 0; ch < nch; ++ch) {
            uint32_t v[16];
            tmem_ld_32x16(t_row + FA_COL_S + ch * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (ch * 16 + i < valid) mt = fmaxf(mt, __uint_as_float(v[i]));
          }
          if (j == 0) {
            m = mt;
          } else {
            const bool need = (mt - m) * c > 8.0f;
            if (__any_sync(0xffffffffu, need)) {
              const float f = need ? ex2_approx((m - mt) * c) : 1.0f;
#pragma unroll 1
              for (int hh = 0; hh < 2; ++hh) {
                uint32_t o[32];
                tmem_ld_32x32(t_row + FA_COL_O + hh * 32, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
                tmem_st_32x32(t_row + FA_COL_O + hh * 32, o);
              }
              l *= f;
              if (need) m = mt;
            }
          }
          const float nmc = -m * c;
          for (int ch = 0; ch < nch; ++ch) {
            uint32_t v[16];
            tmem_ld_32x16(t_row + FA_COL_S + ch * 16, v);
            tmem_ld_wait();
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int k0 = ch * 16 + 2 * i;
              const float p0 = k0 < valid ? ex2_approx(fmaf(__uint_as_float(v[2 * i]), c, nmc)) : 0.f;
              const float p1 = k0 + 1 < valid ? ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), c, nmc)) : 0.f;
              l += p0 + p1;
              pk[i] = pack_bf16x2(p0, p1);
            }
            tmem_st_32x8(t_row + FA_COL_P + ch * 8, pk);
          }
        }
        tmem_st_wait();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }
    // ------------------------------------------------------------------ epilogue: O / l -> bf16 -> global
    mbar_wait(bar_o, k & 1);
    tcgen05_fence_after();
    uint32_t ob[32];  // the 64 outputs of this row, normalised and packed to bf16 pairs
    if (warp_active) {
      const float inv = 1.0f / l;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t o[32];
        tmem_ld_32x32(t_row + FA_COL_O + hh * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          ob[hh * 16 + i] = pack_bf16x2(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv);
      }
    }
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_o_empty);  // O has left TMEM: the next item's first PV may overwrite it
    if (warp_active) {
      const int tok = qt * FA_BQ + row;
      if (tok < T) {
        uint4* dst = reinterpret_cast<uint4*>(args.out + ((size_t)(bh / args.heads) * T + tok) * args.C + (bh % args.heads) * FA_D);
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = make_uint4(ob[4 * i], ob[4 * i + 1], ob[4 * i + 2], ob[4 * i + 3]);
      }
    }
    }  // item loop
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc<FA_TMEM_COLS>(tmem_base);
  }
}

}  // namespace cvit

using namespace cvit;

extern "C" int cvit_attention_fwd_bf16(const void* qkv, void* out, int64_t n_slices, int64_t tokens, int64_t heads,
                                       int64_t head_dim, void* stream) {
  if (!qkv || !out || n_slices <= 0 || tokens <= 0 || heads <= 0) {
    set_error("attention: bad arguments");
    return CVIT_ERR_INVALID;
  }
  if (head_dim != FA_D) {
    set_error("attention: head_dim=%lld unsupported (64 only: every DINOv2 variant)", (long long)head_dim);
    return CVIT_ERR_UNSUPPORTED;
  }
  if (((tokens + FA_BQ - 1) / FA_BQ) * heads * n_slices > 0x7fffffffll) {
    set_error("attention: too many work items");
    return CVIT_ERR_UNSUPPORTED;
  }
  const int64_t C = heads * FA_D;
  CUtensorMap tm;
  uint64_t dims[3] = {(uint64_t)(3 * C), (uint64_t)tokens, (uint64_t)n_slices};
  uint64_t strides[3] = {0, (uint64_t)(3 * C) * 2, (uint64_t)tokens * 3 * C * 2};
  uint32_t box[3] = {FA_D, FA_BK, 1};
  int rc = encode_tmap(&tm, TmapDtype::BF16, 3, qkv, dims, strides, box, 128);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM);
    if (e != cudaSuccess) {
      set_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return CVIT_ERR_CUDA;
    }
    configured = true;
  }
  FaArgs a;
  a.out = static_cast<__nv_bfloat16*>(out);
  a.T = (int)tokens;
  a.heads = (int)heads;
  a.C = (int)C;
  a.slices = (int)n_slices;
  a.scale_log2e = 0.125f * 1.4426950408889634f;
  const int64_t items = ((tokens + FA_BQ - 1) / FA_BQ) * heads * n_slices;
  int grid = 2 * num_sms();
  if (grid > items) grid = (int)items;
  attention_tcgen05_kernel<<<grid, FA_THREADS, FA_SMEM, (cudaStream_t)stream>>>(tm, a);
  return check_launch("attention_tcgen05_kernel");
}
