// Flash-style multi-head self-attention over one ViT slice (no mask), head_dim 64, bf16 in / bf16 out,
// fp32 online softmax. Restates upstream MemEffAttention: softmax(q k^T / sqrt(64)) v with
// q, k, v = qkv.reshape(B, N, 3, H, 64) (HF: modeling_dinov2_with_registers.py:202-256).
//
// Legacy comparison kernel (the first round-1 version; the product path is attention_tcgen05.cu): warp-level mma.sync (m16n8k16) tiles, cp.async double-buffered K/V, warp-level online
// softmax (quad shuffles). Each CTA owns BQ query rows of one (slice, head); each warp owns 16 rows.
// The 5-token ragged tail (1029 = 16*64 + 5) is handled by masking keys >= T to -inf and clamping loads.
#include "ptx.cuh"
#include "tmap.h"

namespace cvit {

constexpr int ATT_D = 64;
constexpr int ATT_BK = 64;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// smem tile: rows of 64 bf16 (128 B), 16-byte chunk index XOR-swizzled with (row & 7)
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int col /*element, multiple of 8*/) {
  return base + row * 128 + ((((col >> 3) ^ (row & 7))) << 4);
}

template <int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) attention_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                 __nv_bfloat16* __restrict__ out, int T, int heads,
                                                                 float scale_log2e) {
  constexpr int BQ = NWARPS * 16;
  constexpr int NTHREADS = NWARPS * 32;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + BQ * 128;            // 2 buffers x 64 rows x 128 B
  const uint32_t sV = sK + 2 * ATT_BK * 128;    // 2 buffers

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * BQ;
  const int head = blockIdx.y, b = blockIdx.z;
  const int C = heads * ATT_D;
  const int64_t row_stride = 3 * (int64_t)C;
  const __nv_bfloat16* base = qkv + (int64_t)b * T * row_stride + head * ATT_D;
  const __nv_bfloat16* gQ = base;
  const __nv_bfloat16* gK = base + C;
  const __nv_bfloat16* gV = base + 2 * C;

  // ---- async loads: Q tile (group 0 together with KV tile 0)
  for (int i = tid; i < BQ * 8; i += NTHREADS) {
    const int r = i >> 3, ch = i & 7;
    const int tok = q0 + r;
    cp_async16(tile_addr(sQ, r, ch * 8), gQ + (int64_t)min(tok, T - 1) * row_stride + ch * 8, tok < T);
  }
  auto load_kv = [&](int kt, int buf) {
    const int k0 = kt * ATT_BK;
    for (int i = tid; i < ATT_BK * 8; i += NTHREADS) {
      const int r = i >> 3, ch = i & 7;
      const int tok = k0 + r;
      const bool ok = tok < T;
      const int64_t off = (int64_t)min(tok, T - 1) * row_stride + ch * 8;
      cp_async16(tile_addr(sK + buf * ATT_BK * 128, r, ch * 8), gK + off, ok);
      cp_async16(tile_addr(sV + buf * ATT_BK * 128, r, ch * 8), gV + off, ok);
    }
  };
  const int num_kt = (T + ATT_BK - 1) / ATT_BK;
  load_kv(0, 0);
  cp_async_commit();

  uint32_t qf[4][4];  // A fragments of this warp's 16 query rows, 4 k-steps over d
  float o[8][4];
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;

  for (int kt = 0; kt < num_kt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < num_kt) load_kv(kt + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (kt == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = ks * 16 + (lane >> 4) * 8;
        ldsm_x4(tile_addr(sQ, r, c), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    const uint32_t kb = sK + buf * ATT_BK * 128, vb = sV + buf * ATT_BK * 128;
    // ---- S = Q K^T (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-key blocks
        uint32_t b0, b1, b2, b3;
        const int key = np * 16 + (lane & 7) + (lane >> 4) * 8;
        const int dcol = ks * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(tile_addr(kb, key, dcol), b0, b1, b2, b3);
        mma_bf16_16816(s[2 * np], qf[ks], b0, b1);
        mma_bf16_16816(s[2 * np + 1], qf[ks], b2, b3);
      }
    }
    // ---- mask ragged tail, online softmax (rows lane/4 and lane/4 + 8 of the warp's 16)
    const int kbase = kt * ATT_BK;
    if (kbase + ATT_BK > T) {
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const int key = kbase + nb * 8 + (lane & 3) * 2;
        if (key >= T) s[nb][0] = s[nb][2] = -INFINITY;
        if (key + 1 >= T) s[nb][1] = s[nb][3] = -INFINITY;
      }
    }
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nb][0], s[nb][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nb][2], s[nb][3]));
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
    }
    float corr[2], msc[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      corr[h] = exp2f((m_run[h] - mx[h]) * scale_log2e);  // m_run = -inf on the first tile -> 0
      msc[h] = mx[h] * scale_log2e;
      m_run[h] = mx[h];
      l_run[h] *= corr[h];
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];  // P as A fragments for the PV product (4 k-steps of 16 keys)
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const float p0 = exp2f(s[nb][0] * scale_log2e - msc[0]);
      const float p1 = exp2f(s[nb][1] * scale_log2e - msc[0]);
      const float p2 = exp2f(s[nb][2] * scale_log2e - msc[1]);
      const float p3 = exp2f(s[nb][3] * scale_log2e - msc[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pf[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    l_run[0] += rs[0];
    l_run[1] += rs[1];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      o[nb][0] *= corr[0];
      o[nb][1] *= corr[0];
      o[nb][2] *= corr[1];
      o[nb][3] *= corr[1];
    }
    // ---- O += P V
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {     // 16 keys per step
#pragma unroll
      for (int np = 0; np < 4; ++np) {   // pairs of 8-wide d blocks
        uint32_t b0, b1, b2, b3;
        const int key = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int dcol = np * 16 + (lane >> 4) * 8;
        ldsm_x4_t(tile_addr(vb, key, dcol), b0, b1, b2, b3);
        mma_bf16_16816(o[2 * np], pf[ks], b0, b1);
        mma_bf16_16816(o[2 * np + 1], pf[ks], b2, b3);
      }
    }
    __syncthreads();  // everyone done with buf before the next iteration's prefetch overwrites it
  }
  cp_async_wait<0>();

  // ---- normalise and store (each quad owns a row; lanes hold cols (lane&3)*2, +1 of each 8-wide block)
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float l = l_run[h];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const float inv = 1.0f / l;
    const int tok = q0 + warp * 16 + (lane >> 2) + h * 8;
    if (tok < T) {
      __nv_bfloat16* orow = out + ((int64_t)b * T + tok) * C + head * ATT_D + (lane & 3) * 2;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        *reinterpret_cast<uint32_t*>(orow + nb * 8) = pack_bf16x2(o[nb][2 * h] * inv, o[nb][2 * h + 1] * inv);
      }
    }
  }
}

}  // namespace cvit

using namespace cvit;

extern "C" int cvit_attention_fwd_bf16_mma_sync(const void* qkv, void* out, int64_t n_slices, int64_t tokens, int64_t heads,
                                       int64_t head_dim, void* stream) {
  if (!qkv || !out || n_slices <= 0 || tokens <= 0 || heads <= 0) {
    set_error("attention: bad arguments");
    return CVIT_ERR_INVALID;
  }
  if (head_dim != ATT_D) {
    set_error("attention: head_dim=%lld unsupported (64 only: every DINOv2 variant)", (long long)head_dim);
    return CVIT_ERR_UNSUPPORTED;
  }
  constexpr int NW = 4;
  constexpr int BQ = NW * 16;
  const int smem = BQ * 128 + 4 * ATT_BK * 128;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return CVIT_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid((unsigned)((tokens + BQ - 1) / BQ), (unsigned)heads, (unsigned)n_slices);
  const float scale_log2e = 0.125f * 1.4426950408889634f;
  attention_kernel<NW><<<grid, NW * 32, smem, (cudaStream_t)stream>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), (int)tokens, (int)heads, scale_log2e);
  return check_launch("attention_kernel");
}
