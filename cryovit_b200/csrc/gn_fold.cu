// GroupNorm of the CryoVIT head folded into its neighbours (inference; reference models/cryovit.py:56-62: every
// SynthesisBlock starts with GroupNorm(max(8, c1/8), c1, eps=1e-3) followed by Conv3d(c1, c2, 3, dilation=(d1,1,1))).
//
// A GroupNorm between a producer P and a convolution is   conv_w(a_c * x_c + b_c),   a_c = gamma_c * rstd_g,
// b_c = beta_c - mean_g * a_c.  Instead of two more passes over the volume (statistics, then normalise: 3.0 GB of the
// head's 10.3 GB of DRAM traffic in round 1),
//   1. P's epilogue writes per-(32-row block, group) partial sums of what it stores (gemm_tcgen05.cuh, gn_stats),
//   2. gn_finalize_kernel reduces them in a fixed order (deterministic) to a_c, b_c,
//   3. gn_fold_kernel writes the bf16 weights w' = w * a_c in the consumer's own operand layout and the 64-row bias
//      table  T[mask][co] = bias[co] + sum over the taps that `mask` says are inside the volume of sum_c w[tap][co][c] b_c
//      ("same" padding pads the NORMALISED tensor with zeros, so a border voxel must not receive b_c through the taps
//      that fall outside),
//   4. the consumer convolves the un-normalised tensor with w' and adds the table row of each voxel.
// Both kernels here are a few microseconds; the weights are re-folded for every volume (the statistics are per volume).
#include "ptx.cuh"
#include "tmap.h"

namespace cvit {

// partials: [R][Pc] float2 (sum, sum of squares), Pc = reps * G (reps = 4 sub-pixels for a transposed-conv producer).
// Two levels, both in a fixed order (bit-reproducible): every block sums a contiguous slice of the R row blocks for
// all Pc columns at once (coalesced: a block row is Pc consecutive float2), folds the reps and parks [G] double pairs
// in `slices`; the block that finishes last (atomic ticket) adds the slices up in slice order and writes
// ab: [2][C] = scale a_c = gamma_c * rstd_g, shift b_c = beta_c - mean_g * a_c. The ticket counter returns to zero.
constexpr int GN_MAX_SLICES = 64;
__global__ void __launch_bounds__(256) gn_finalize_kernel(const float2* __restrict__ partials, int64_t R, int Pc, int G, int cpg,
                                                          double n_per_group, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float eps, float* __restrict__ ab,
                                                          unsigned int* __restrict__ ticket, double2* __restrict__ slices) {
  const int S = gridDim.x, sidx = blockIdx.x;
  const int64_t r0 = R * sidx / S, r1 = R * (sidx + 1) / S;
  const int col = threadIdx.x % Pc, lane_row = threadIdx.x / Pc, row_step = blockDim.x / Pc;  // Pc divides 256
  double s = 0.0, q = 0.0;
  for (int64_t r = r0 + lane_row; r < r1; r += row_step) {
    const float2 v = partials[r * Pc + col];
    s += v.x;
    q += v.y;
  }
  __shared__ double rs[256], rq[256];
  rs[threadIdx.x] = s;
  rq[threadIdx.x] = q;
  __syncthreads();
  if ((int)threadIdx.x < G) {  // group g of this slice: its reps columns x the row lanes, in a fixed order
    double ts = 0.0, tq = 0.0;
    const int reps = Pc / G;
    for (int rep = 0; rep < reps; ++rep)
      for (int lr = 0; lr < row_step; ++lr) {
        ts += rs[lr * Pc + rep * G + threadIdx.x];
        tq += rq[lr * Pc + rep * G + threadIdx.x];
      }
    slices[(size_t)sidx * G + threadIdx.x] = make_double2(ts, tq);
  }
  __threadfence();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == (unsigned)S - 1u;
  __syncthreads();
  if (!last) return;
  __threadfence();
  {  // slice sums -> group sums with every thread: thread = (slice lane, group), then the lanes in a fixed order
    const int g = threadIdx.x % G, sl = threadIdx.x / G, nsl = blockDim.x / G;  // G divides 256 (checked by the launcher)
    double ts = 0.0, tq = 0.0;
    for (int i = sl; i < S; i += nsl) {
      const double2 v = slices[(size_t)i * G + g];
      ts += v.x;
      tq += v.y;
    }
    __syncthreads();
    rs[threadIdx.x] = ts;
    rq[threadIdx.x] = tq;
    __syncthreads();
  }
  if ((int)threadIdx.x < G) {
    double ts = 0.0, tq = 0.0;
    for (int l = 0; l < (int)blockDim.x / G; ++l) {
      ts += rs[l * G + threadIdx.x];
      tq += rq[l * G + threadIdx.x];
    }
    const double mean = ts / n_per_group;
    double var = tq / n_per_group - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const int C = G * cpg;
    for (int i = 0; i < cpg; ++i) {
      const int c = threadIdx.x * cpg + i;
      const float a = __ldg(gamma + c) * rstd;
      ab[c] = a;
      ab[C + c] = __ldg(beta + c) - (float)mean * a;
    }
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// Operand layouts of the two consumers.
//   LAYOUT_TAPS: conv3d_dilated (gemm.cu conv3_rows): [tap = (kd*3+kh)*3+kw][coutp][cin]
//   LAYOUT_HALO: conv3d_halo (conv_halo.cu, cin >= 16): [kd*3+kh][mma i = kw * (cin/16) + c2][k chunk][coutp][8],
//                channel = c2*16 + k*8 + e
//   LAYOUT_WPACKN: conv3d_wpackn (conv_wpackn.cu): [kd*3+kh][K step][2 chunks][n = j_out * coutp + co][8], chunk k = 2 step + c
//                = (window voxel j_in, c_hi) = divmod(k, cin / 8), channel = c_hi*8 + e, tap kw = j_in - j_out; `p` voxels per row
//   LAYOUT_ROWS: conv3d_rows (conv_rows.cu): [pass h = co / 16][rotation v][chunk j = ci / 8][K step][2 chunks c][row 144][8],
//                row = ((yr * 3 + slot) * 16 + co % 16), kw = 2 step + c (the fourth tap is zero), kh = 2 - yr,
//                kd = (v + 1 - slot) mod 3, channel = j*8 + e; `p` = cin / 8
enum { LAYOUT_TAPS = 0, LAYOUT_HALO = 1, LAYOUT_WPACKN = 2, LAYOUT_ROWS = 3 };

template <int LAYOUT>
__device__ __forceinline__ void decode(int64_t idx, int cin, int coutp, int& tap, int& co, int& ci, int p = 0) {
  if (LAYOUT == LAYOUT_ROWS) {
    ci = (int)((idx / (8 * 144 * 4)) % (cin / 8)) * 8 + (int)(idx & 7);
    tap = co = 0;
  } else if (LAYOUT == LAYOUT_WPACKN) {
    const int ch = cin / 8, ksteps = (p + 2) * ch / 2, n_cols = p * coutp;
    const int e = (int)(idx & 7);
    int64_t r = idx >> 3;
    r /= n_cols;  // the column (j_out, co) does not matter for the channel
    const int c = (int)(r & 1);
    const int st = (int)((r >> 1) % ksteps);
    ci = ((2 * st + c) % ch) * 8 + e;
    tap = co = 0;
  } else if (LAYOUT == LAYOUT_TAPS) {
    ci = (int)(idx % cin);
    const int64_t r = idx / cin;
    co = (int)(r % coutp);
    tap = (int)(r / coutp);
  } else {
    const int steps = cin / 16;
    const int e = (int)(idx & 7);
    int64_t r = idx >> 3;
    co = (int)(r % coutp);
    r /= coutp;
    const int k = (int)(r & 1);
    r >>= 1;
    const int i = (int)(r % (3 * steps));
    const int t9 = (int)(r / (3 * steps));
    const int kw = i / steps, c2 = i - kw * steps;
    tap = t9 * 3 + kw;
    ci = c2 * 16 + k * 8 + e;
  }
}

// blocks [0, coutp): one output channel each: B[tap] = sum_ci w32 * b[ci], then the 64 table rows (first in the grid:
//                    they are the longer-running blocks)
// blocks [coutp, gridDim.x): w_out = bf16(w32 * a[ci]) over the whole image (grid stride)
template <int LAYOUT>
__global__ void __launch_bounds__(256) gn_fold_kernel(const float* __restrict__ w32, __nv_bfloat16* __restrict__ w_out, int64_t n_elems,
                                                      int cin, int coutp, const float* __restrict__ ab, int C,
                                                      const float* __restrict__ bias, float* __restrict__ table, int p) {
  if ((int)blockIdx.x >= coutp) {
    const unsigned n_scale_blocks = gridDim.x - (unsigned)coutp, b0 = blockIdx.x - (unsigned)coutp;
    // n_elems < 2^31 (checked by the launcher): 32-bit index arithmetic, one divide per element
    for (unsigned idx = b0 * blockDim.x + threadIdx.x; idx < (unsigned)n_elems; idx += n_scale_blocks * blockDim.x) {
      int ci;
      if (LAYOUT == LAYOUT_TAPS) {
        ci = (int)(idx % (unsigned)cin);
      } else {
        int tap, co;
        decode<LAYOUT>((int64_t)idx, cin, coutp, tap, co, ci, p);
      }
      w_out[idx] = __float2bfloat16(w32[idx] * ab[ci]);
    }
    return;
  }
  const int co = (int)blockIdx.x;
  __shared__ float B[27];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // warp w takes taps w, w + 8, ...; its lanes stride over the input channels (coalesced), fixed-order shuffle tree
  for (int tap = warp; tap < 27; tap += 8) {
    float acc = 0.f;
    for (int ci = lane; ci < cin; ci += 32) {
      int64_t idx;
      if (LAYOUT == LAYOUT_TAPS) {
        idx = ((int64_t)tap * coutp + co) * cin + ci;
      } else if (LAYOUT == LAYOUT_ROWS) {
        // the tap's weight as it sits in rotation v = 0: slot = (1 - kd) mod 3
        const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3, slot = (4 - kd) % 3, yr = 2 - kh;
        const int row = (yr * 3 + slot) * 16 + (co & 15);
        idx = ((((((int64_t)(co >> 4) * 3 + 0) * (cin / 8) + (ci >> 3)) * 2 + (kw >> 1)) * 2 + (kw & 1)) * 144 + row) * 8 + (ci & 7);
      } else if (LAYOUT == LAYOUT_WPACKN) {
        // the tap's weights as they sit in the column of output voxel j_out = 0: window voxel j_in = kw
        const int ch = cin / 8, ksteps = (p + 2) * ch / 2, n_cols = p * coutp, t9 = tap / 3, kw = tap - t9 * 3;
        const int k = kw * ch + (ci >> 3);
        idx = ((((int64_t)t9 * ksteps + (k >> 1)) * 2 + (k & 1)) * n_cols + co) * 8 + (ci & 7);
      } else {
        const int steps = cin / 16, t9 = tap / 3, kw = tap - t9 * 3;
        const int c2 = ci >> 4, k = (ci >> 3) & 1, e = ci & 7;
        idx = ((((int64_t)t9 * (3 * steps) + kw * steps + c2) * 2 + k) * coutp + co) * 8 + e;
      }
      acc += w32[idx] * ab[C + ci];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) B[tap] = acc;
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int m = threadIdx.x, dm = m >> 4, hm = (m >> 2) & 3, wm = m & 3;
    float v = bias[co];
    for (int kd = 0; kd < 3; ++kd) {
      if ((kd == 0 && !(dm & 1)) || (kd == 2 && !(dm & 2))) continue;
      for (int kh = 0; kh < 3; ++kh) {
        if ((kh == 0 && !(hm & 1)) || (kh == 2 && !(hm & 2))) continue;
        for (int kw = 0; kw < 3; ++kw) {
          if ((kw == 0 && !(wm & 1)) || (kw == 2 && !(wm & 2))) continue;
          v += B[(kd * 3 + kh) * 3 + kw];
        }
      }
    }
    table[m * coutp + co] = v;
  }
}

}  // namespace cvit

using namespace cvit;

extern "C" int64_t cvit_conv3d_wpackn_group(int64_t Cin, int64_t Cout_pad);

// See include/cryovit_b200.h.
// fp32 elements the `ab` argument of cvit_groupnorm_fold must hold: scale / shift, the ticket, the slice sums.
extern "C" int64_t cvit_groupnorm_fold_ab_elems(int64_t channels, int64_t groups) {
  return ((2 * channels + 3) / 4 * 4) + 4 + (int64_t)GN_MAX_SLICES * groups * 4;
}

extern "C" int cvit_groupnorm_fold(const float* partials, int64_t rows32, int64_t partial_cols, int64_t groups, int64_t channels,
                                   double n_per_group, const float* gamma, const float* beta, float eps, float* ab,
                                   const float* w32, void* w_out, int64_t n_elems, int64_t cin, int64_t cout_pad, int layout,
                                   const float* bias, float* table, void* stream) {
  if (!partials || !gamma || !beta || !ab || !w32 || !w_out || !bias || !table || rows32 <= 0 || groups <= 0 ||
      channels % groups != 0 || partial_cols % groups != 0 || cin != channels || cout_pad <= 0 || n_per_group <= 0.0 ||
      n_elems >= (1ll << 31) || layout < LAYOUT_TAPS || layout > LAYOUT_ROWS || (layout == LAYOUT_HALO && (cin % 16) != 0)) {
    set_error("groupnorm_fold: bad arguments (rows32=%lld cols=%lld G=%lld C=%lld cin=%lld coutp=%lld n=%lld layout=%d)", (long long)rows32,
              (long long)partial_cols, (long long)groups, (long long)channels, (long long)cin, (long long)cout_pad, (long long)n_elems,
              layout);
    return CVIT_ERR_INVALID;
  }
  int p = 0;
  if (layout == LAYOUT_WPACKN) {
    p = (int)cvit_conv3d_wpackn_group(cin, cout_pad);
    if (p == 0 || n_elems != 9 * (p + 2) * cin * p * cout_pad) {
      set_error("groupnorm_fold: no W-packed layout for cin=%lld cout_pad=%lld (or n_elems=%lld does not match it)", (long long)cin,
                (long long)cout_pad, (long long)n_elems);
      return CVIT_ERR_INVALID;
    }
  } else if (layout == LAYOUT_ROWS) {
    if ((cin != 16 && cin != 32) || (cout_pad != 16 && cout_pad != 32) || n_elems != (cout_pad / 16) * 3 * (cin / 8) * 4 * 144 * 8) {
      set_error("groupnorm_fold: no rows layout for cin=%lld cout=%lld (or n_elems=%lld does not match it)", (long long)cin,
                (long long)cout_pad, (long long)n_elems);
      return CVIT_ERR_INVALID;
    }
  } else if (n_elems != 27 * cin * cout_pad) {
    set_error("groupnorm_fold: n_elems=%lld does not match 27 x %lld x %lld", (long long)n_elems, (long long)cin, (long long)cout_pad);
    return CVIT_ERR_INVALID;
  }
  const int cpg = (int)(channels / groups);
  if (groups > 256 || (256 % groups) != 0 || partial_cols > 256 || (256 % partial_cols) != 0 || (reinterpret_cast<uintptr_t>(ab) & 15u)) {
    set_error("groupnorm_fold: groups=%lld / partial columns=%lld unsupported (columns must divide 256; ab 16-byte aligned)",
              (long long)groups, (long long)partial_cols);
    return CVIT_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // ab: [2 * channels] scale / shift | pad to 16 B | ticket (zero before the first call, returned to zero) | slices
  const size_t off = ((size_t)2 * channels + 3) / 4 * 4;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(ab + off);
  double2* slices = reinterpret_cast<double2*>(ab + off + 4);
  int64_t n_slices = (rows32 + 7) / 8;
  if (n_slices > GN_MAX_SLICES) n_slices = GN_MAX_SLICES;
  gn_finalize_kernel<<<(unsigned)n_slices, 256, 0, st>>>(reinterpret_cast<const float2*>(partials), rows32, (int)partial_cols,
                                                        (int)groups, cpg, n_per_group, gamma, beta, eps, ab, ticket, slices);
  int rc = check_launch("gn_finalize_kernel");
  if (rc) return rc;
  int64_t scale_blocks = (n_elems + 256 * 8 - 1) / (256 * 8);
  const int64_t cap = (int64_t)num_sms() * 8;
  if (scale_blocks > cap) scale_blocks = cap;
  const unsigned grid = (unsigned)(scale_blocks + cout_pad);
  if (layout == LAYOUT_TAPS)
    gn_fold_kernel<LAYOUT_TAPS><<<grid, 256, 0, st>>>(w32, static_cast<__nv_bfloat16*>(w_out), n_elems, (int)cin, (int)cout_pad, ab,
                                                     (int)channels, bias, table, 0);
  else if (layout == LAYOUT_HALO)
    gn_fold_kernel<LAYOUT_HALO><<<grid, 256, 0, st>>>(w32, static_cast<__nv_bfloat16*>(w_out), n_elems, (int)cin, (int)cout_pad, ab,
                                                     (int)channels, bias, table, 0);
  else if (layout == LAYOUT_ROWS)
    gn_fold_kernel<LAYOUT_ROWS><<<grid, 256, 0, st>>>(w32, static_cast<__nv_bfloat16*>(w_out), n_elems, (int)cin, (int)cout_pad, ab,
                                                     (int)channels, bias, table, 0);
  else
    gn_fold_kernel<LAYOUT_WPACKN><<<grid, 256, 0, st>>>(w32, static_cast<__nv_bfloat16*>(w_out), n_elems, (int)cin, (int)cout_pad, ab,
                                                       (int)channels, bias, table, p);
  return check_launch("gn_fold_kernel");
}
