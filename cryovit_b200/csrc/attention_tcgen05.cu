// Flash-style self-attention of one ViT slice on tcgen05 / TMEM (head_dim 64, bf16 in/out, fp32 softmax).
// Restates upstream MemEffAttention: softmax(q k^T / 8) v, q,k,v = qkv.reshape(B, N, 3, H, 64).
//
// Persistent kernel, one CTA per SM, 512 threads = four warpgroups; registers are re-balanced with setmaxnreg
// (softmax warps hold a whole 128-wide S row). A work item is (slice, head, 256 query rows) = two 128-row query
// tiles A and B that share every K/V tile; items are walked query-pair fastest so co-running CTAs share K/V in L2.
// Barrier phases, the K/V ring and the S/P/O buffers run across items: the tile stream never drains at an item end.
//   warps 0-3   softmax group A, warps 4-7 softmax group B; one thread per query row (TMEM lane = row). The whole
//               128-wide S row is held in registers: row max (3-input FMNMX), lazy rescale (O and l are only rescaled
//               when the max grows by more than 2^8), p = exp2(s*c - m*c) with packed f32x2 FMA/ADD -> bf16 pairs
//               -> TMEM. P(j) is published to the MMA warp only after the loads of S(j+1) were issued, so the TMEM
//               store drain (300-600 cycles when the tensor pipe is busy) is off the critical path.
//   warps 8-11  epilogue: O_g / l -> bf16 -> global, while the softmax warps already run the next item.
//   warp 12     TMA producer: Q_A|Q_B once per item, (K, V) tiles of 128 keys through a 4-stage ring. The tensor map
//               is 3-D (column, token, slice): tokens past the end of a slice are zero-filled, never the next slice.
//   warps 13,14 MMA issuers, one per query tile (whole warp walks the schedule, one elected lane issues):
//                   S_g = Q_g K^T  (SS: K-major operands in 128B-swizzled smem, N = 128 keys)
//                   O_g += P_g V   (TS: P read from TMEM as packed bf16, V tile as an MN-major B operand)
//               Each issues S_g(t+1) as soon as its group has pulled S_g(t) into registers (so the next S is ready
//               before the softmax of the current one ends), then PV_g(t) when P_g(t) is published.
// TMEM (512 columns): S_A [0,128) S_B [128,256) | P_A [256,320) P_B [320,384) | O_A [384,448) O_B [448,512).
// The ragged tail (1029 = 8*128 + 5 keys) runs as an N=16 MMA with the 11 padding keys masked to -inf.
// A slice whose query-tile count is odd ends with an A-only item (group B idles through it).
#include "ptx.cuh"
#include "tmap.h"

namespace cvit {

constexpr int FA_BQ = 128, FA_BK = 128, FA_D = 64;
constexpr int FA_THREADS = 512;
// setmaxnreg budget per scheduler (one warp of each warpgroup): 2 x softmax + epilogue + control <= 512
constexpr int FA_REGS_SOFTMAX = 184, FA_REGS_EPILOGUE = 72, FA_REGS_CONTROL = 72;
constexpr int FA_STAGES = 4;
constexpr int FA_TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16
constexpr int FA_SMEM = (2 + 2 * FA_STAGES) * FA_TILE_BYTES + 2048 /* row sums */ + 512 /* barriers */ + 1024;
constexpr int FA_TMEM_COLS = 512;
constexpr uint32_t FA_COL_S = 0, FA_COL_P = 256, FA_COL_O = 384;

// Optional per-phase cycle accounting (tools/fa_trace.py builds with -DCVIT_FA_TRACE): every warp accumulates the
// cycles it spends in each phase in registers and dumps the totals once at the end (a timeline of global stores
// proved too intrusive: it moved the phases it was measuring).
#ifdef CVIT_FA_TRACE
#define FA_PROF_DECL uint32_t prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; uint32_t prof_last = clock()
#define FA_PROF(slot) do { const uint32_t now_ = clock(); prof_acc[slot] += now_ - prof_last; prof_last = now_; } while (0)
#define FA_PROF_DUMP(tiles) do { if (args.trace && blockIdx.x < 4 && lane == 0) {                      \
    long long* d_ = args.trace + (blockIdx.x * 16 + warp) * 10;                                        \
    for (int i_ = 0; i_ < 8; ++i_) d_[i_] = prof_acc[i_];                                              \
    d_[8] = (tiles); } } while (0)
#else
#define FA_PROF_DECL do { } while (0)
#define FA_PROF(slot) do { } while (0)
#define FA_PROF_DUMP(tiles) do { } while (0)
#endif
struct FaArgs {
  long long* trace;    // debug accounting (tools/fa_trace.py), normally null
  __nv_bfloat16* out;  // [B*T, C]
  int T, heads, C, slices;
  float scale_log2e;   // head_dim^-0.5 * log2(e)
};

// packed fp32 pairs (Blackwell FFMA2 / FADD2): halves the issue slots of the softmax inner loop
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__global__ void __launch_bounds__(FA_THREADS, 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, const FaArgs args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;                                        // Q_A | Q_B
  const uint32_t sK = smem_base + 2 * FA_TILE_BYTES;                    // FA_STAGES tiles
  const uint32_t sV = smem_base + (2 + FA_STAGES) * FA_TILE_BYTES;      // FA_STAGES tiles
  const uint32_t sL = smem_base + (2 + 2 * FA_STAGES) * FA_TILE_BYTES;  // row sums: [item parity][group][128] fp32
  const uint32_t sBar = sL + 2048;
  const uint32_t bar_q = sBar, bar_q_empty = sBar + 8;
  const uint32_t bar_kv_full = sBar + 16;                 // FA_STAGES x 8
  const uint32_t bar_kv_empty = sBar + 16 + 8 * FA_STAGES;
  const uint32_t bar_grp = sBar + 16 + 16 * FA_STAGES;    // per group (64 B apart): the seven barriers below
  constexpr uint32_t B_S = 0, B_SFREE = 8, B_P = 16, B_PV = 24, B_OFULL = 32, B_OEMPTY = 40, B_LFULL = 48;
  const uint32_t tmem_slot = bar_grp + 128;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = args.T;
  const int n_tiles = (T + FA_BK - 1) / FA_BK;   // K/V tiles per item
  const int n_qt = (T + FA_BQ - 1) / FA_BQ;      // query tiles per (slice, head)
  const int n_pairs = (n_qt + 1) >> 1;
  const int total_items = n_pairs * args.heads * args.slices;
  // K/V tiles are walked ragged-tile-first: the short tile's P is published at once and its PV retires under the
  // first full tile (walked last, the single P buffer would stall the softmax behind PV of the tile before it).
  const bool ragged = (T % FA_BK) != 0;
  auto kv_tile = [&](int j) { return ragged ? (j == 0 ? n_tiles - 1 : j - 1) : j; };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(bar_q, 1);
    mbar_init(bar_q_empty, 2);  // one commit per group (group A commits twice on A-only items)
    for (int s = 0; s < FA_STAGES; ++s) {
      mbar_init(bar_kv_full + 8 * s, 1);
      mbar_init(bar_kv_empty + 8 * s, 2);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(bar_grp + 64 * g + B_S, 1);       // MMA -> softmax: S_g(t) complete
      mbar_init(bar_grp + 64 * g + B_SFREE, 4);   // softmax -> MMA: S_g(t) is in registers, the TMEM tile may be overwritten
      mbar_init(bar_grp + 64 * g + B_P, 4);       // softmax -> MMA: P_g(t) stored
      mbar_init(bar_grp + 64 * g + B_PV, 1);      // MMA -> softmax: PV_g(t) retired (P reusable, O consistent)
      mbar_init(bar_grp + 64 * g + B_OFULL, 1);   // MMA -> epilogue: the item's last PV_g retired, O_g complete
      mbar_init(bar_grp + 64 * g + B_OEMPTY, 4);  // epilogue -> MMA: O_g read out, the next item may overwrite it
      mbar_init(bar_grp + 64 * g + B_LFULL, 4);   // softmax -> epilogue: row sums of the item are in smem
    }
    fence_mbar_init();
  }
  if (warp == 13) tmem_alloc<FA_TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FA_REGS_CONTROL));
    if (warp == 12) {
      // ------------------------------------------------------------------ TMA producer
      if (lane == 0) {
        uint32_t kv_it = 0, k = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++k) {
          const int pair = item % n_pairs, bh = item / n_pairs;
          const int head = bh % args.heads, slice = bh / args.heads;
          const int cq = head * FA_D, ck = args.C + head * FA_D, cv = 2 * args.C + head * FA_D;
          mbar_wait(bar_q_empty, (k & 1) ^ 1u);  // every S MMA of the previous item has retired
          mbar_arrive_expect_tx(bar_q, 2 * FA_TILE_BYTES);
          tma_load_3d(sQ, &tmQKV, bar_q, cq, (2 * pair) * FA_BQ, slice);
          tma_load_3d(sQ + FA_TILE_BYTES, &tmQKV, bar_q, cq, (2 * pair + 1) * FA_BQ, slice);  // all-OOB box = zeros
          for (int j = 0; j < n_tiles; ++j, ++kv_it) {
            const uint32_t s = kv_it % FA_STAGES;
            mbar_wait(bar_kv_empty + 8 * s, ((kv_it / FA_STAGES) & 1) ^ 1u);
            mbar_arrive_expect_tx(bar_kv_full + 8 * s, 2 * FA_TILE_BYTES);
            tma_load_3d(sK + s * FA_TILE_BYTES, &tmQKV, bar_kv_full + 8 * s, ck, kv_tile(j) * FA_BK, slice);
            tma_load_3d(sV + s * FA_TILE_BYTES, &tmQKV, bar_kv_full + 8 * s, cv, kv_tile(j) * FA_BK, slice);
          }
        }
      }
    } else if (warp <= 14) {
      // ------------------------------------------------------------------ MMA issuers: warp 13 group A, warp 14 group B
      // One issuer warp per query tile: each walks its own stream  S_g(t+1), PV_g(t), S_g(t+2), PV_g(t+1) ...  and
      // never waits on the other group's softmax (a single issuer serving both in a fixed order coupled the two
      // groups through its blocking waits). tcgen05.commit only tracks the issuing thread's own MMAs, which is
      // exactly the per-group completion the softmax warps need. The whole warp walks the schedule (warp-uniform
      // control flow keeps the descriptors in uniform registers); one elected lane issues.
      // Shared barriers take one arrival per group: kv_empty (count 2) and q_empty (count 2); on A-only items
      // warp 13 commits twice, and warp 14 still waits every q / kv_full phase in order (a waiter that skipped
      // phases would alias on the parity bit).
      const int g = warp - 13;
      const uint32_t bg = bar_grp + 64 * g;
      constexpr uint32_t idesc_pv = umma_idesc_bf16_f32(FA_BQ, FA_D) | (1u << 16);  // B is MN-major
      auto n_mma_of = [&](int j) {  // keys of the j-th walked tile rounded up to the MMA granularity (16)
        const int valid = min(FA_BK, T - kv_tile(j) * FA_BK);
        return (valid + 15) & ~15;
      };
      auto paired = [&](int item) { return 2 * (item % n_pairs) + 1 < n_qt; };  // group B has a real query tile
      FA_PROF_DECL;
      // S_g for the tile with per-group index tg: walked tile j of the item whose ring position is kv
      auto issue_s = [&](uint32_t tg, int j, uint32_t kv, bool last_s_of_item, bool solo) {
        FA_PROF(7);
        if (tg > 0) mbar_wait(bg + B_SFREE, (tg - 1) & 1);
        FA_PROF(0);  // wait: S tile free
        const uint32_t s = kv % FA_STAGES;
        mbar_wait(bar_kv_full + 8 * s, (kv / FA_STAGES) & 1);
        tcgen05_fence_after();
        FA_PROF(1);  // wait: K/V landed
        if (elect_one_sync()) {
          const uint32_t idesc_s = umma_idesc_bf16_f32(FA_BQ, n_mma_of(j));
          const uint64_t qd = umma_smem_desc_kmajor<128>(sQ + g * FA_TILE_BYTES);
          const uint64_t kd = umma_smem_desc_kmajor<128>(sK + s * FA_TILE_BYTES);
          const uint32_t tS = tmem_base + FA_COL_S + g * 128;
#pragma unroll
          for (int kk = 0; kk < FA_D / 16; ++kk) umma_bf16(tS, qd + 2 * kk, kd + 2 * kk, idesc_s, kk > 0);
          if (last_s_of_item) {
            umma_commit(bar_q_empty);
            if (solo) umma_commit(bar_q_empty);
          }
          umma_commit(bg + B_S);
        }
        __syncwarp();
        FA_PROF(2);  // issue S
      };
      auto issue_pv = [&](uint32_t tg, int j, uint32_t kv, uint32_t kg, bool solo) {
        FA_PROF(7);
        mbar_wait(bg + B_P, tg & 1);
        FA_PROF(3);  // wait: P stored
        if (j == 0) mbar_wait(bg + B_OEMPTY, (kg & 1) ^ 1u);  // previous item's O_g has been read out of TMEM
        tcgen05_fence_after();
        FA_PROF(4);  // wait: O drained
        const uint32_t s = kv % FA_STAGES;
        if (elect_one_sync()) {
          const uint32_t tP = tmem_base + FA_COL_P + g * 64, tO = tmem_base + FA_COL_O + g * 64;
          const uint64_t vd = umma_smem_desc_mnmajor_sw128(sV + s * FA_TILE_BYTES, FA_TILE_BYTES);
          if (n_mma_of(j) == FA_BK) {
#pragma unroll
            for (int kk = 0; kk < FA_BK / 16; ++kk)  // 16 keys per step: 8 packed TMEM columns of P, 2 KB of the V tile
              umma_bf16_ts(tO, tP + 8 * kk, vd + 128 * kk, idesc_pv, (j | kk) != 0);
          } else {
            for (int kk = 0; kk < n_mma_of(j) / 16; ++kk)
              umma_bf16_ts(tO, tP + 8 * kk, vd + 128 * kk, idesc_pv, (j | kk) != 0);
          }
          umma_commit(bar_kv_empty + 8 * s);  // K(j), V(j) free once both groups' MMAs on them retired
          if (solo) umma_commit(bar_kv_empty + 8 * s);
          umma_commit(bg + B_PV);
          if (j + 1 == n_tiles) umma_commit(bg + B_OFULL);
        }
        __syncwarp();
        FA_PROF(5);  // issue PV
      };
      // skipped (A-only) items: group B's issuer still observes every phase of the shared barriers
      auto skip_item = [&](uint32_t k) {
        mbar_wait(bar_q, k & 1);
        for (int j = 0; j < n_tiles; ++j) {
          const uint32_t kv = k * n_tiles + j;
          mbar_wait(bar_kv_full + 8 * (kv % FA_STAGES), (kv / FA_STAGES) & 1);
        }
      };
      int item = blockIdx.x;
      uint32_t k = 0, tg = 0, kg = 0;
      while (item < total_items && g == 1 && !paired(item)) {
        skip_item(k);
        item += gridDim.x;
        ++k;
      }
      if (item < total_items) {
        mbar_wait(bar_q, k & 1);
        issue_s(0, 0, k * n_tiles, n_tiles == 1, g == 0 && !paired(item));
      }
      while (item < total_items) {
        const bool solo = g == 0 && !paired(item);
        const int next = item + gridDim.x;  // group A: always the next item; group B: only if it is a paired one
        const bool next_direct = next < total_items && (g == 0 || paired(next));
        for (int j = 0; j < n_tiles; ++j) {
          const uint32_t kv = k * n_tiles + j;
          // the next S tile first: it only needs the group's current S to be in registers
          bool s_after = false;
          if (j + 1 < n_tiles) {
            issue_s(tg + 1, j + 1, kv + 1, j + 2 == n_tiles, solo);
          } else if (next_direct) {
            // the next item's Q tiles were requested when this item's last S retired; if the other group is late
            // and they have not landed yet, do not hold this group's last PV back
            if (__all_sync(0xffffffffu, mbar_try_wait(bar_q, (k + 1) & 1)))
              issue_s(tg + 1, 0, kv + 1, n_tiles == 1, g == 0 && !paired(next));
            else
              s_after = true;
          }
          issue_pv(tg, j, kv, kg, solo);
          if (s_after) {
            mbar_wait(bar_q, (k + 1) & 1);
            issue_s(tg + 1, 0, kv + 1, n_tiles == 1, g == 0 && !paired(next));
          }
          ++tg;
        }
        ++kg;
        item = next;
        ++k;
        if (!next_direct && item < total_items) {  // group B steps over A-only items, then starts its next stream
          while (item < total_items && !paired(item)) {
            skip_item(k);
            item += gridDim.x;
            ++k;
          }
          if (item < total_items) {
            mbar_wait(bar_q, k & 1);
            issue_s(tg, 0, k * n_tiles, n_tiles == 1, false);
          }
        }
      }
      FA_PROF(7);
      FA_PROF_DUMP(tg);
    }  // warp 15 idle
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ epilogue: O / l -> bf16 -> global
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FA_REGS_EPILOGUE));
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float* sL_gen = reinterpret_cast<const float*>(smem_gen + (sL - smem_base));
    uint32_t kg[2] = {0, 0};
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int pair = item % n_pairs, bh = item / n_pairs;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int qt = 2 * pair + g;
        if (qt >= n_qt) continue;
        const uint32_t bg = bar_grp + 64 * g;
        const uint32_t par = kg[g] & 1;
        ++kg[g];
        mbar_wait(bg + B_LFULL, par);
        const float l = sL_gen[(par * 2 + g) * 128 + row];
        mbar_wait(bg + B_OFULL, par);
        tcgen05_fence_after();
        const bool warp_active = qt * FA_BQ + q * 32 < T;  // warp-uniform
        uint32_t ob[32];  // the 64 outputs of this row, normalised and packed to bf16 pairs
        if (warp_active) {
          const float inv = 1.0f / l;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t o[32];
            tmem_ld_32x32(t_row + FA_COL_O + g * 64 + hh * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              ob[hh * 16 + i] = pack_bf16x2(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bg + B_OEMPTY);  // O has left TMEM: the next item's first PV may overwrite it
        if (warp_active) {
          const int tok = qt * FA_BQ + row;
          if (tok < T) {
            uint4* dst = reinterpret_cast<uint4*>(args.out + ((size_t)(bh / args.heads) * T + tok) * args.C + (bh % args.heads) * FA_D);
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = make_uint4(ob[4 * i], ob[4 * i + 1], ob[4 * i + 2], ob[4 * i + 3]);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps, thread == query row
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FA_REGS_SOFTMAX));
    const int g = warp >> 2;        // 0 = group A, 1 = group B
    const int q = warp & 3;         // TMEM lane quarter this warp may access (hardware rule: warp id % 4)
    const int row = q * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t tS = t_row + FA_COL_S + g * 128, tP = t_row + FA_COL_P + g * 64, tO = t_row + FA_COL_O + g * 64;
    const uint32_t bg = bar_grp + 64 * g;
    const float c = args.scale_log2e;
    const uint64_t c2 = pack_f32x2(c, c);
    float* sL_gen = reinterpret_cast<float*>(smem_gen + (sL - smem_base));
    uint32_t tile_it = 0, kg = 0;
    bool pending = false;  // P of the previous tile is stored but not yet published to the MMA warp
    FA_PROF_DECL;
    auto publish = [&]() {
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bg + B_P);
      pending = false;
    };
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int pair = item % n_pairs;
      const int qt = 2 * pair + g;
      if (qt >= n_qt) continue;  // A-only item: the MMA warp skips this group too
      const bool warp_active = qt * FA_BQ + q * 32 < T;  // warp-uniform: all-padding warps only keep the barriers moving
      float m = 0.f, l = 0.f;
      for (int j = 0; j < n_tiles; ++j, ++tile_it) {
        FA_PROF(7);  // loop
        mbar_wait(bg + B_S, tile_it & 1);
        tcgen05_fence_after();
        FA_PROF(0);  // wait for S
        const int valid = min(FA_BK, T - kv_tile(j) * FA_BK);
        bool pv_seen = tile_it == 0;  // PV(t-1) known retired: O may be rescaled, P overwritten
        if (warp_active && valid == FA_BK) {
          // The whole 128-wide S row lives in registers: one TMEM read per tile, four loads in flight.
          uint32_t v[128];
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            tmem_ld_32x32(tS + ch * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32 * ch]));
          if (pending) publish();  // P(j-1): its TMEM store drained under the loads above
          tmem_ld_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bg + B_SFREE);  // the MMA warp may start S(j+1) under this tile's softmax
          FA_PROF(1);  // TMEM -> registers
          // ---- row max: 4 independent chains of 3-input max
          float mx[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) mx[u] = fmaxf(__uint_as_float(v[2 * u]), __uint_as_float(v[2 * u + 1]));
#pragma unroll
          for (int i = 8; i < 128; i += 8) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              mx[u] = fmax3(mx[u], __uint_as_float(v[i + 2 * u]), __uint_as_float(v[i + 2 * u + 1]));
          }
          const float mt = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          if (j == 0) {
            m = mt;
          } else {
            const bool need = (mt - m) * c > 8.0f;
            if (__any_sync(0xffffffffu, need)) {
              mbar_wait(bg + B_PV, (tile_it - 1) & 1);
              tcgen05_fence_after();
              pv_seen = true;
              const float f = need ? ex2_approx((m - mt) * c) : 1.0f;
#pragma unroll 1
              for (int hh = 0; hh < 4; ++hh) {  // 16 columns at a time: the S row keeps 128 registers busy
                uint32_t o[16];
                tmem_ld_32x16(tO + hh * 16, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
                tmem_st_32x16(tO + hh * 16, o);
              }
              l *= f;
              if (need) m = mt;
            }
          }
          FA_PROF(2);  // row max (+ rescale)
          // ---- probabilities, packed in place (element pair i -> register i); packed FMA + packed row sums
          const float nmc = -m * c;
          const uint64_t nmc2 = pack_f32x2(nmc, nmc);
          uint64_t ls[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            const uint64_t t2 = fma_f32x2(pack_f32x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), c2, nmc2);
            float t0, t1;
            unpack_f32x2(t2, t0, t1);
            const float p0 = ex2_approx(t0), p1 = ex2_approx(t1);
            ls[i & 3] = add_f32x2(ls[i & 3], pack_f32x2(p0, p1));
            v[i] = pack_bf16x2(p0, p1);
          }
          {
            float a0, a1;
            unpack_f32x2(add_f32x2(add_f32x2(ls[0], ls[1]), add_f32x2(ls[2], ls[3])), a0, a1);
            l += a0 + a1;
          }
          FA_PROF(4);  // exp
          if (!pv_seen) {
            mbar_wait(bg + B_PV, (tile_it - 1) & 1);
            tcgen05_fence_after();
          }
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            tmem_st_32x16(tP + ch * 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16 * ch]));
          pending = true;
          if (j + 1 == n_tiles) publish();  // the epilogue is waiting for this one
          FA_PROF(5);  // wait PV(j-1), store P
        } else {
          if (pending) publish();
          if (!pv_seen) {
            mbar_wait(bg + B_PV, (tile_it - 1) & 1);
            tcgen05_fence_after();
          }
          if (warp_active) {
            // ---- ragged last tile: `valid` keys inside an N = roundup16(valid) MMA, 16-column chunks
            const int nch = (valid + 15) >> 4;
            float mt = -INFINITY;
            for (int ch = 0; ch < nch; ++ch) {
              uint32_t v[16];
              tmem_ld_32x16(tS + ch * 16, v);
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
                  tmem_ld_32x32(tO + hh * 32, o);
                  tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
                  tmem_st_32x32(tO + hh * 32, o);
                }
                l *= f;
                if (need) m = mt;
              }
            }
            const float nmc = -m * c;
            for (int ch = 0; ch < nch; ++ch) {
              uint32_t v[16];
              tmem_ld_32x16(tS + ch * 16, v);
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
              tmem_st_32x8(tP + ch * 8, pk);
            }
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bg + B_SFREE);
          publish();
          FA_PROF(6);  // ragged / padding tile
        }
      }
      // hand the row sums to the epilogue warps and move on to the next item
      sL_gen[((kg & 1) * 2 + g) * 128 + row] = l;
      ++kg;
      __syncwarp();
      if (lane == 0) mbar_arrive(bg + B_LFULL);
    }  // item loop
    FA_PROF(7);
    FA_PROF_DUMP(tile_it);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 13) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc<FA_TMEM_COLS>(tmem_base);
  }
}

}  // namespace cvit

using namespace cvit;

#ifdef CVIT_FA_TRACE
static long long* g_fa_trace = nullptr;
extern "C" void cvit_fa_set_trace(long long* p) { g_fa_trace = p; }
#endif

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
#ifdef CVIT_FA_TRACE
  a.trace = g_fa_trace;
#else
  a.trace = nullptr;
#endif
  a.out = static_cast<__nv_bfloat16*>(out);
  a.T = (int)tokens;
  a.heads = (int)heads;
  a.C = (int)C;
  a.slices = (int)n_slices;
  a.scale_log2e = 0.125f * 1.4426950408889634f;
  const int n_qt = (int)((tokens + FA_BQ - 1) / FA_BQ);
  const int64_t items = (int64_t)((n_qt + 1) / 2) * heads * n_slices;
  int grid = num_sms();
  if (grid > items) grid = (int)items;
  attention_tcgen05_kernel<<<grid, FA_THREADS, FA_SMEM, (cudaStream_t)stream>>>(tm, a);
  return check_launch("attention_tcgen05_kernel");
}
