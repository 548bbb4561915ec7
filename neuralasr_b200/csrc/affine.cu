// The model tails' affine projection and its backward: logits = H.W + b on the reshaped recurrent
// outputs H [rows, K] (reference networks/bilstm_ctc_net.py:33-45, lstm_ctc_net.py:28-40: K = num_hidden = 500,
// W [K, C] Xavier, b [C] zero; rows = B*T, or 2*B*T for the BiLSTM tail, whose reshape stacks both directions).
//
// Why its own kernels instead of a producer fused into the CTC kernel: H is 13 times the logits in bytes and the
// loss reads every frame in two phases half a recursion apart, so a fused producer would read H twice.
// Here H is read ONCE, the logits are written once (39 MB at cfg3: they stay in the 126 MB L2 for the loss kernel
// that follows on the same stream), and the roofline is HBM: 4*(rows*K + rows*C) bytes per call.
//
// Arithmetic: float32 in, float32 out, on the tensor cores as 3xTF32 -- every operand is split into a TF32 high
// part and a TF32 low part, and hi*hi + hi*lo + lo*hi is accumulated in float32 (mma.sync.m16n8k8; the dropped
// lo*lo term is 2^-22 relative), which keeps the result within float32 rounding of tf.matmul's float32 product.
// A single-pass TF32 product (10-bit mantissa) would put ~1e-3 absolute error on the logits and break the 1e-4
// gradient tolerance of the path; tcgen05 has no float32-accurate mode and its operands come from shared memory, so
// the split would need a conversion pass through registers anyway, for a kernel that is bound by reading H from HBM.
//
// Kernels:
//   affine_rows_kernel   out[r, n] (+)= sum_k A[r, k] * Wt[n, k] (+ bias[n]);  Wt (<= 40 columns x <= 672 k, hi and lo
//                        parts) lives in shared memory for the whole CTA, each warp walks its own contiguous range of
//                        rows and reads A straight from global memory into mma fragments (16-byte loads along k: the
//                        k order inside a tile is permuted identically for A and Wt, so no transposition is needed).
//                        Forward: A = H, Wt = W^T.  Backward dH = dL.W^T: A = dL, Wt = W itself, 13 column blocks.
//   affine_dw_kernel     dW[k, c] = sum_r H[r, k] * dL[r, c], db[c] = sum_r dL[r, c]: each CTA streams its range of
//                        rows through a three-stage cp.async ring (32 rows of H and dL per stage), every warp owns 32
//                        of the 512 k of a block as accumulators; per-CTA partials go to the workspace and
//                        affine_reduce_kernel adds them in a fixed order (deterministic, no atomics).
#include <algorithm>
#include <cstdlib>

#include "nasr_common.cuh"

namespace nasr {
namespace affine {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kNT = 5;            // n8 tiles of a column block
constexpr int kNB = 8 * kNT;      // 40 output columns per CTA
constexpr int kMaxSeg = 672;      // longest contraction segment whose Wt (hi + lo) fits shared memory

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// hi*hi + hi*lo + lo*hi, small terms first
__device__ __forceinline__ void mma_3x(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                       uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_tf32(d, al[0], al[1], al[2], al[3], bh0, bh1);
  mma_tf32(d, ah[0], ah[1], ah[2], ah[3], bl0, bl1);
  mma_tf32(d, ah[0], ah[1], ah[2], ah[3], bh0, bh1);
}

struct RowsParams {
  const float* A;        // [rows, Kc] at stride lda
  long long lda;
  long long rows;
  int Kc;                // contraction length of this call (<= kMaxSeg)
  const float* W;        // element (n, k) of Wt at W[n * w_sn + k * w_sk]
  long long w_sn, w_sk;
  int N;                 // output columns
  const float* bias;     // [N] or NULL
  float* out;            // [rows, N] at stride ldo
  long long ldo;
  int accumulate;        // out += instead of out =
  int kstride;           // floats per staged Wt row (see launch_rows)
  int vec_a;             // rows of A allow the mode's vector loads (16-byte in MODE 16, 8-byte in MODE 8)
  int vec_o;             // rows of out start 8-byte aligned
};

// Split of an operand that is about to be fed to an mma: hi is x rounded to TF32's 10 mantissa bits (add half an ulp,
// clear the low 13 bits: 2 instructions where cvt.rna.tf32.f32 compiles to 4), lo = x - hi exactly; lo goes to the
// tensor core as it is, which ignores its low 13 bits (error < 2^-21 |x|, either sign).
__device__ __forceinline__ void split_fast(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

__device__ __forceinline__ float4 load_a4_checked(const float* row, int k, int Kc) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k < Kc) v.x = __ldg(row + k);
  if (k + 1 < Kc) v.y = __ldg(row + k + 1);
  if (k + 2 < Kc) v.z = __ldg(row + k + 2);
  if (k + 3 < Kc) v.w = __ldg(row + k + 3);
  return v;
}

__device__ __forceinline__ float2 load_a2_checked(const float* row, int k, int Kc) {
  float2 v = make_float2(0.f, 0.f);
  if (k < Kc) v.x = __ldg(row + k);
  if (k + 1 < Kc) v.y = __ldg(row + k + 1);
  return v;
}

// accumulator (c0, c1) = (row g, columns 2t, 2t+1), (c2, c3) = row g + 8
template <int MT>
__device__ __forceinline__ void store_tile(const RowsParams& p, const float (&acc)[MT][kNT][4], long long r0,
                                           long long r_end, int n_base, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int m = 0; m < MT; m++) {
#pragma unroll
    for (int half = 0; half < 2; half++) {
      const long long r = r0 + 16 * m + 8 * half + g;
      if (r >= r_end) continue;
      float* orow = p.out + r * p.ldo;
#pragma unroll
      for (int j = 0; j < kNT; j++) {
        const int n = n_base + j * 8 + 2 * t;
        if (n >= p.N) continue;
        float v0 = acc[m][j][2 * half], v1 = acc[m][j][2 * half + 1];
        const bool two = n + 1 < p.N;
        if (p.bias) {
          v0 += __ldg(p.bias + n);
          if (two) v1 += __ldg(p.bias + n + 1);
        }
        if (two && p.vec_o) {
          float2* dst = reinterpret_cast<float2*>(orow + n);
          if (p.accumulate) {
            const float2 old = *dst;
            v0 += old.x;
            v1 += old.y;
          }
          *dst = make_float2(v0, v1);
        } else {
          if (p.accumulate) {
            v0 += orow[n];
            if (two) v1 += orow[n + 1];
          }
          orow[n] = v0;
          if (two) orow[n + 1] = v1;
        }
      }
    }
  }
}

// One warp, MT m16 tiles of rows starting at r0 (rows >= r_end are computed on a clamped row and not stored), all kNT
// column tiles of the CTA's block.  The k order inside a tile is permuted identically for A and Wt, so that a lane
// reads consecutive k of one row with one vector load and the sum over k is unchanged:
//   MODE 16 (rows of A 16-byte aligned): lane (g = lane / 4, t = lane % 4) loads A[row g + 8i][16*ch + 4t .. +3]; the
//            four values are the k slots (t, t+4) of two m16n8k8 products;
//   MODE 8  (anything else; 8-byte loads where the rows allow): A[row][8*st + 2t, +1] are the slots (t, t+4) of one.
template <int MT, int MODE>
__device__ __forceinline__ void rows_tile(const RowsParams& p, const uint32_t* __restrict__ Wh,
                                          const uint32_t* __restrict__ Wl, long long r0, long long r_end,
                                          int n_base, int lane) {
  const int g = lane >> 2, t = lane & 3;
  float acc[MT][kNT][4];
#pragma unroll
  for (int m = 0; m < MT; m++)
#pragma unroll
    for (int j = 0; j < kNT; j++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[m][j][e] = 0.f;
  const float* arow[2 * MT];
#pragma unroll
  for (int i = 0; i < 2 * MT; i++) arow[i] = p.A + min(r0 + g + 8 * i, p.rows - 1) * p.lda;
  const int kstride = p.kstride;
  const int Kc = p.Kc;

  if (MODE == 16) {
    const int nch = (Kc + 15) >> 4, nfull = Kc >> 4;   // chunks, and chunks whose 16 k all exist
    const uint32_t* wh = Wh + g * kstride + 4 * t;
    const uint32_t* wl = Wl + g * kstride + 4 * t;
    float4 cur[2 * MT], nxt[2 * MT];
#pragma unroll
    for (int i = 0; i < 2 * MT; i++) {
      cur[i] = nfull > 0 ? __ldg(reinterpret_cast<const float4*>(arow[i] + 4 * t))
                         : load_a4_checked(arow[i], 4 * t, Kc);
      nxt[i] = cur[i];
    }
    for (int ch = 0; ch < nch; ch++) {
      const int kn = (ch + 1) * 16 + 4 * t;
      if (ch + 1 < nfull) {
#pragma unroll
        for (int i = 0; i < 2 * MT; i++) nxt[i] = __ldg(reinterpret_cast<const float4*>(arow[i] + kn));
      } else if (ch + 1 < nch) {
#pragma unroll
        for (int i = 0; i < 2 * MT; i++) nxt[i] = load_a4_checked(arow[i], kn, Kc);
      }
      // the rows' bytes five chunks ahead go to L2 now: the register prefetch above then waits an L2 hit, not DRAM
      if (t == 0 && (ch + 5) * 16 < Kc) {
#pragma unroll
        for (int i = 0; i < 2 * MT; i++) asm volatile("prefetch.global.L2 [%0];" ::"l"(arow[i] + (ch + 5) * 16));
      }
      uint32_t ah[MT][2][4], al[MT][2][4];   // [m tile][product 0/1][a0..a3]
#pragma unroll
      for (int m = 0; m < MT; m++) {
        const float4 ra = cur[2 * m], rb = cur[2 * m + 1];   // rows g and g + 8 of the tile
        split_fast(ra.x, ah[m][0][0], al[m][0][0]);
        split_fast(rb.x, ah[m][0][1], al[m][0][1]);
        split_fast(ra.y, ah[m][0][2], al[m][0][2]);
        split_fast(rb.y, ah[m][0][3], al[m][0][3]);
        split_fast(ra.z, ah[m][1][0], al[m][1][0]);
        split_fast(rb.z, ah[m][1][1], al[m][1][1]);
        split_fast(ra.w, ah[m][1][2], al[m][1][2]);
        split_fast(rb.w, ah[m][1][3], al[m][1][3]);
      }
#pragma unroll
      for (int j = 0; j < kNT; j++) {
        const uint4 bh = *reinterpret_cast<const uint4*>(wh + j * 8 * kstride + ch * 16);
        const uint4 bl = *reinterpret_cast<const uint4*>(wl + j * 8 * kstride + ch * 16);
#pragma unroll
        for (int m = 0; m < MT; m++) {
          mma_3x(acc[m][j], ah[m][0], al[m][0], bh.x, bh.y, bl.x, bl.y);
          mma_3x(acc[m][j], ah[m][1], al[m][1], bh.z, bh.w, bl.z, bl.w);
        }
      }
#pragma unroll
      for (int i = 0; i < 2 * MT; i++) cur[i] = nxt[i];
    }
  } else {
    const int nst = (Kc + 7) >> 3, nfull = p.vec_a ? (Kc >> 3) : 0;   // vec_a: rows 8-byte aligned in this mode
    const uint32_t* wh = Wh + g * kstride + 2 * t;
    const uint32_t* wl = Wl + g * kstride + 2 * t;
    float2 cur[2 * MT], nxt[2 * MT];
#pragma unroll
    for (int i = 0; i < 2 * MT; i++) {
      cur[i] = nfull > 0 ? __ldg(reinterpret_cast<const float2*>(arow[i] + 2 * t))
                         : load_a2_checked(arow[i], 2 * t, Kc);
      nxt[i] = cur[i];
    }
    for (int st = 0; st < nst; st++) {
      const int kn = (st + 1) * 8 + 2 * t;
      if (st + 1 < nfull) {
#pragma unroll
        for (int i = 0; i < 2 * MT; i++) nxt[i] = __ldg(reinterpret_cast<const float2*>(arow[i] + kn));
      } else if (st + 1 < nst) {
#pragma unroll
        for (int i = 0; i < 2 * MT; i++) nxt[i] = load_a2_checked(arow[i], kn, Kc);
      }
      uint32_t ah[MT][4], al[MT][4];
#pragma unroll
      for (int m = 0; m < MT; m++) {
        split_fast(cur[2 * m].x, ah[m][0], al[m][0]);
        split_fast(cur[2 * m + 1].x, ah[m][1], al[m][1]);
        split_fast(cur[2 * m].y, ah[m][2], al[m][2]);
        split_fast(cur[2 * m + 1].y, ah[m][3], al[m][3]);
      }
#pragma unroll
      for (int j = 0; j < kNT; j++) {
        const uint2 bh = *reinterpret_cast<const uint2*>(wh + j * 8 * kstride + st * 8);
        const uint2 bl = *reinterpret_cast<const uint2*>(wl + j * 8 * kstride + st * 8);
#pragma unroll
        for (int m = 0; m < MT; m++) mma_3x(acc[m][j], ah[m], al[m], bh.x, bh.y, bl.x, bl.y);
      }
#pragma unroll
      for (int i = 0; i < 2 * MT; i++) cur[i] = nxt[i];
    }
  }
  store_tile<MT>(p, acc, r0, r_end, n_base, lane);
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) affine_rows_kernel(const RowsParams p) {
  extern __shared__ __align__(16) uint32_t smem_u[];
  uint32_t* Wh = smem_u;
  uint32_t* Wl = smem_u + kNB * p.kstride;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_base = blockIdx.y * kNB;
  // stage this block's 40 columns of Wt, split once for every warp; the index runs along whichever axis is
  // contiguous in global memory
  const int total = kNB * p.kstride;
  for (int idx = tid; idx < total; idx += kThreads) {
    int n, k;
    if (p.w_sn == 1) {
      k = idx / kNB;
      n = idx - k * kNB;
    } else {
      n = idx / p.kstride;
      k = idx - n * p.kstride;
    }
    float w = 0.f;
    if (n_base + n < p.N && k < p.Kc) w = __ldg(p.W + (long long)(n_base + n) * p.w_sn + (long long)k * p.w_sk);
    uint32_t hi, lo;
    split_tf32(w, hi, lo);
    Wh[n * p.kstride + k] = hi;
    Wl[n * p.kstride + k] = lo;
  }
  __syncthreads();
  // every warp of the grid takes one contiguous range of rows (a multiple of 16), walked in tiles of 32 and 16
  const long long warps = (long long)gridDim.x * kWarps;
  const long long gw = (long long)blockIdx.x * kWarps + warp;
  long long per = (p.rows + warps - 1) / warps;
  per = (per + 15) & ~15LL;
  long long r = gw * per;
  const long long r_end = min(p.rows, r + per);
  while (r < r_end) {
    if (r_end - r > 16) {
      rows_tile<2, MODE>(p, Wh, Wl, r, r_end, n_base, lane);
      r += 32;
    } else {
      rows_tile<1, MODE>(p, Wh, Wl, r, r_end, n_base, lane);
      r += 16;
    }
  }
}

// Contractions of at most 40 (dH = dL.W^T at the reference's C = 38 / 41): a warp keeps its rows of A in registers
// (five 8-byte steps), Wt for up to 17 column blocks lives in shared memory, and the warp walks the column blocks
// itself -- A is read from global memory once, the prologue runs once per row tile.
constexpr int kSmallK = 40;
constexpr int kSmallSteps = kSmallK / 8;
constexpr int kSmallBlocks = 17;      // 17 * 40 rows of Wt, hi + lo, 40 floats each: 217.6 KB

template <int MT>
__device__ __forceinline__ void smallk_tile(const RowsParams& p, const uint32_t* __restrict__ Wh,
                                            const uint32_t* __restrict__ Wl, long long r0, long long r_end,
                                            int n_first, int nblocks, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int Kc = p.Kc, nst = (Kc + 7) >> 3, nfull = p.vec_a ? (Kc >> 3) : 0;
  float2 a[kSmallSteps][2 * MT];
#pragma unroll
  for (int i = 0; i < 2 * MT; i++) {
    const float* row = p.A + min(r0 + g + 8 * i, p.rows - 1) * p.lda;
#pragma unroll
    for (int st = 0; st < kSmallSteps; st++) {
      if (st < nfull) a[st][i] = __ldg(reinterpret_cast<const float2*>(row + st * 8 + 2 * t));
      else a[st][i] = load_a2_checked(row, st * 8 + 2 * t, Kc);
    }
  }
  const uint32_t* wh = Wh + g * kSmallK + 2 * t;
  const uint32_t* wl = Wl + g * kSmallK + 2 * t;
  for (int nb = 0; nb < nblocks; nb++) {
    float acc[MT][kNT][4];
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
      for (int j = 0; j < kNT; j++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[m][j][e] = 0.f;
#pragma unroll
    for (int st = 0; st < kSmallSteps; st++) {
      if (st < nst) {
        uint32_t ah[MT][4], al[MT][4];
#pragma unroll
        for (int m = 0; m < MT; m++) {
          split_fast(a[st][2 * m].x, ah[m][0], al[m][0]);
          split_fast(a[st][2 * m + 1].x, ah[m][1], al[m][1]);
          split_fast(a[st][2 * m].y, ah[m][2], al[m][2]);
          split_fast(a[st][2 * m + 1].y, ah[m][3], al[m][3]);
        }
#pragma unroll
        for (int j = 0; j < kNT; j++) {
          const int off = (nb * kNB + j * 8) * kSmallK + st * 8;
          const uint2 bh = *reinterpret_cast<const uint2*>(wh + off);
          const uint2 bl = *reinterpret_cast<const uint2*>(wl + off);
#pragma unroll
          for (int m = 0; m < MT; m++) mma_3x(acc[m][j], ah[m], al[m], bh.x, bh.y, bl.x, bl.y);
        }
      }
    }
    store_tile<MT>(p, acc, r0, r_end, n_first + nb * kNB, lane);
  }
}

__global__ void __launch_bounds__(kThreads, 1) affine_smallk_kernel(const RowsParams p) {
  extern __shared__ __align__(16) uint32_t smem_u[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_first = blockIdx.y * kSmallBlocks * kNB;
  const int nblocks = min(kSmallBlocks, (p.N - n_first + kNB - 1) / kNB);
  uint32_t* Wh = smem_u;
  uint32_t* Wl = smem_u + nblocks * kNB * kSmallK;
  const int total = nblocks * kNB * kSmallK;
  for (int idx = tid; idx < total; idx += kThreads) {
    int n, k;
    if (p.w_sn == 1) {
      k = idx / (nblocks * kNB);
      n = idx - k * (nblocks * kNB);
    } else {
      n = idx / kSmallK;
      k = idx - n * kSmallK;
    }
    float w = 0.f;
    if (n_first + n < p.N && k < p.Kc) w = __ldg(p.W + (long long)(n_first + n) * p.w_sn + (long long)k * p.w_sk);
    uint32_t hi, lo;
    split_tf32(w, hi, lo);
    Wh[n * kSmallK + k] = hi;
    Wl[n * kSmallK + k] = lo;
  }
  __syncthreads();
  const long long warps = (long long)gridDim.x * kWarps;
  const long long gw = (long long)blockIdx.x * kWarps + warp;
  long long per = (p.rows + warps - 1) / warps;
  per = (per + 15) & ~15LL;
  long long r = gw * per;
  const long long r_end = min(p.rows, r + per);
  while (r < r_end) {
    if (r_end - r > 16) {
      smallk_tile<2>(p, Wh, Wl, r, r_end, n_first, nblocks, lane);
      r += 32;
    } else {
      smallk_tile<1>(p, Wh, Wl, r, r_end, n_first, nblocks, lane);
      r += 16;
    }
  }
}

// ---- dW / db ---------------------------------------------------------------------------------------------------
constexpr int kSlab = 32;            // rows per stage
constexpr int kKB = 512;             // k per CTA (32 per warp)
constexpr int kHS = kKB + 8;         // floats per staged H row (8 mod 32: the fragment loads hit 32 banks once)
constexpr int kDS = kNB;             // floats per staged dL row (40 = 8 mod 32)
constexpr int kStages = 3;
constexpr size_t kDwSmem = (size_t)kStages * kSlab * (kHS + kDS) * sizeof(float);

struct DwParams {
  const float* H;
  long long ldh;
  const float* dL;
  long long ldd;
  long long rows;
  int K, C;
  float* part;     // [gridDim.x][gridDim.y * kKB + 1][gridDim.z * kNB]; the last row of a slice holds db's partial
  int vec_h;       // 16-byte cp.async legal for H rows
  int vec_d;       // 8-byte cp.async legal for dL rows
};

__device__ __forceinline__ void cp_async(void* dst, const void* src, int bytes, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  if (bytes == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
  else if (bytes == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) affine_dw_kernel(const DwParams p) {
  extern __shared__ __align__(16) float smem_f[];
  float* Hs = smem_f;                                        // [stage][kSlab][kHS]
  float* Ds = smem_f + (size_t)kStages * kSlab * kHS;        // [stage][kSlab][kDS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int k0 = blockIdx.y * kKB, c0 = blockIdx.z * kNB;
  const int kv = min(kKB, p.K - k0), cv = min(kNB, p.C - c0);   // valid k / c of this block
  // columns the copies never write stay zero for the whole kernel
  for (int i = tid; i < kStages * kSlab * kHS; i += kThreads) Hs[i] = 0.f;
  for (int i = tid; i < kStages * kSlab * kDS; i += kThreads) Ds[i] = 0.f;
  __syncthreads();

  long long per = (p.rows + gridDim.x - 1) / gridDim.x;
  per = (per + kSlab - 1) / kSlab * kSlab;
  const long long r_begin = (long long)blockIdx.x * per;
  const long long r_end = min(p.rows, r_begin + per);
  const int slabs = r_begin < r_end ? (int)((r_end - r_begin + kSlab - 1) / kSlab) : 0;

  auto issue = [&](int s) {   // copies of slab s into stage s % kStages (rows past r_end arrive as zeros)
    float* hs = Hs + (size_t)(s % kStages) * kSlab * kHS;
    float* ds = Ds + (size_t)(s % kStages) * kSlab * kDS;
    const long long rs = r_begin + (long long)s * kSlab;
    if (p.vec_h) {
      const int per_row = (kv + 3) >> 2;          // kv is a multiple of 4 whenever vec_h is set
      for (int i = tid; i < kSlab * per_row; i += kThreads) {
        const int rr = i / per_row, q = i - rr * per_row;
        const bool in = rs + rr < r_end;
        cp_async(hs + rr * kHS + 4 * q, p.H + (in ? rs + rr : r_begin) * p.ldh + k0 + 4 * q, 16, in ? 16 : 0);
      }
    } else {
      for (int i = tid; i < kSlab * kv; i += kThreads) {
        const int rr = i / kv, q = i - rr * kv;
        const bool in = rs + rr < r_end;
        cp_async(hs + rr * kHS + q, p.H + (in ? rs + rr : r_begin) * p.ldh + k0 + q, 4, in ? 4 : 0);
      }
    }
    if (p.vec_d) {
      const int per_row = (cv + 1) >> 1;          // cv is even whenever vec_d is set
      for (int i = tid; i < kSlab * per_row; i += kThreads) {
        const int rr = i / per_row, q = i - rr * per_row;
        const bool in = rs + rr < r_end;
        cp_async(ds + rr * kDS + 2 * q, p.dL + (in ? rs + rr : r_begin) * p.ldd + c0 + 2 * q, 8, in ? 8 : 0);
      }
    } else {
      for (int i = tid; i < kSlab * cv; i += kThreads) {
        const int rr = i / cv, q = i - rr * cv;
        const bool in = rs + rr < r_end;
        cp_async(ds + rr * kDS + q, p.dL + (in ? rs + rr : r_begin) * p.ldd + c0 + q, 4, in ? 4 : 0);
      }
    }
  };

  float acc[2][kNT][4];
#pragma unroll
  for (int m = 0; m < 2; m++)
#pragma unroll
    for (int j = 0; j < kNT; j++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[m][j][e] = 0.f;
  float db_acc = 0.f;

  if (slabs > 0) issue(0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  if (slabs > 1) issue(1);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int s = 0; s < slabs; s++) {
    asm volatile("cp.async.wait_group 1;" ::: "memory");   // slab s has landed (one younger group may be in flight)
    __syncthreads();                                        // ... for every thread; and stage (s+2)%3 is free again
    if (s + 2 < slabs) issue(s + 2);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const float* hs = Hs + (size_t)(s % kStages) * kSlab * kHS;
    const float* ds = Ds + (size_t)(s % kStages) * kSlab * kDS;
    if (blockIdx.y == 0 && tid < kNB) {
      float sum = 0.f;
#pragma unroll 8
      for (int rr = 0; rr < kSlab; rr++) sum += ds[rr * kDS + tid];
      db_acc += sum;
    }
#pragma unroll
    for (int ks = 0; ks < kSlab / 8; ks++) {
      // A = H^T: a0 (m = g, k slot t) = H[row t][k g], a1 = k g+8, a2/a3 = row t+4
      const float* h0 = hs + (ks * 8 + t) * kHS + warp * 32 + g;
      const float* h1 = h0 + 4 * kHS;
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int m = 0; m < 2; m++) {
        split_fast(h0[16 * m], ah[m][0], al[m][0]);
        split_fast(h0[16 * m + 8], ah[m][1], al[m][1]);
        split_fast(h1[16 * m], ah[m][2], al[m][2]);
        split_fast(h1[16 * m + 8], ah[m][3], al[m][3]);
      }
      const float* d0 = ds + (ks * 8 + t) * kDS + g;
      const float* d1 = d0 + 4 * kDS;
#pragma unroll
      for (int j = 0; j < kNT; j++) {
        uint32_t bh0, bl0, bh1, bl1;
        split_fast(d0[8 * j], bh0, bl0);
        split_fast(d1[8 * j], bh1, bl1);
#pragma unroll
        for (int m = 0; m < 2; m++) mma_3x(acc[m][j], ah[m], al[m], bh0, bh1, bl0, bl1);
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // partials: slice of this CTA, rows = k of the whole grid (+ one row for db), columns = c of the whole grid
  const long long ncols = (long long)gridDim.z * kNB;
  const long long nrows = (long long)gridDim.y * kKB + 1;
  float* slice = p.part + (long long)blockIdx.x * nrows * ncols;
#pragma unroll
  for (int m = 0; m < 2; m++)
#pragma unroll
    for (int half = 0; half < 2; half++) {
      const long long kk = k0 + warp * 32 + 16 * m + 8 * half + g;
#pragma unroll
      for (int j = 0; j < kNT; j++) {
        float2* dst = reinterpret_cast<float2*>(slice + kk * ncols + c0 + j * 8 + 2 * t);
        *dst = make_float2(acc[m][j][2 * half], acc[m][j][2 * half + 1]);
      }
    }
  if (blockIdx.y == 0 && tid < kNB) slice[(nrows - 1) * ncols + c0 + tid] = db_acc;
}

__global__ void affine_reduce_kernel(const float* __restrict__ part, int slices, long long nrows, long long ncols,
                                     int K, int C, float* __restrict__ dW, float* __restrict__ db) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)(K + 1) * C;
  if (i >= n) return;
  const int kk = (int)(i / C), c = (int)(i - (long long)kk * C);
  const long long row = kk < K ? kk : nrows - 1;
  if (kk < K ? dW == nullptr : db == nullptr) return;
  float s = 0.f;
  for (int x = 0; x < slices; x++) s += part[((long long)x * nrows + row) * ncols + c];
  if (kk < K) dW[(long long)kk * C + c] = s;
  else db[c] = s;
}

static int launch_rows(const float* A, long long lda, long long rows, int Kc_total, const float* W, long long w_sn,
                       long long w_sk, int N, const float* bias, float* out, long long ldo, cudaStream_t stream) {
  int sms = 0;
  NASR_CUDA(device_sm_count(&sms));
  const long long tiles16 = (rows + 15) / 16;
  const int gx = (int)max(1LL, min((long long)sms, (tiles16 + kWarps - 1) / kWarps));
  const int gy = (N + kNB - 1) / kNB;
  if (Kc_total <= kSmallK) {
    RowsParams p;
    p.Kc = Kc_total;
    p.A = A;
    p.lda = lda;
    p.rows = rows;
    p.W = W;
    p.w_sn = w_sn;
    p.w_sk = w_sk;
    p.N = N;
    p.bias = bias;
    p.out = out;
    p.ldo = ldo;
    p.accumulate = 0;
    p.kstride = kSmallK;
    p.vec_a = (((uintptr_t)A & 7) == 0) && (lda % 2 == 0);
    p.vec_o = (((uintptr_t)out & 7) == 0) && (ldo % 2 == 0);
    const int nb_total = (N + kNB - 1) / kNB;
    const int gy_s = (nb_total + kSmallBlocks - 1) / kSmallBlocks;
    const size_t smem = (size_t)2 * min(nb_total, kSmallBlocks) * kNB * kSmallK * sizeof(uint32_t);
    NASR_CUDA((ensure_max_dynamic_smem<affine_smallk_kernel>(227 * 1024)));
    affine_smallk_kernel<<<dim3(gx, gy_s), kThreads, smem, stream>>>(p);
    count_launch();
    NASR_CUDA(cudaGetLastError());
    return NASR_OK;
  }
  NASR_CUDA((ensure_max_dynamic_smem<affine_rows_kernel<16>>(227 * 1024)));
  NASR_CUDA((ensure_max_dynamic_smem<affine_rows_kernel<8>>(227 * 1024)));
  for (int kseg = 0; kseg < Kc_total; kseg += kMaxSeg) {
    RowsParams p;
    p.Kc = min(kMaxSeg, Kc_total - kseg);
    p.A = A + kseg;
    p.lda = lda;
    p.rows = rows;
    p.W = W + (long long)kseg * w_sk;
    p.w_sn = w_sn;
    p.w_sk = w_sk;
    p.N = N;
    p.bias = kseg == 0 ? bias : nullptr;
    p.out = out;
    p.ldo = ldo;
    p.accumulate = kseg > 0;
    // 16-byte loads of A where its rows allow them and are long enough to pay; otherwise 8-byte steps
    const bool mode16 = (((uintptr_t)p.A & 15) == 0) && (lda % 4 == 0) && p.Kc >= 64;
    p.vec_a = mode16 ? 1 : ((((uintptr_t)p.A & 7) == 0) && (lda % 2 == 0));
    // floats per staged Wt row: the fragment loads of 8 rows must hit every bank once (16 mod 32 for 16-byte loads,
    // 8 mod 32 for 8-byte loads)
    p.kstride = ((p.Kc + 31) & ~31) + (mode16 ? 16 : 8);
    p.vec_o = (((uintptr_t)out & 7) == 0) && (ldo % 2 == 0);
    const size_t smem = (size_t)2 * kNB * p.kstride * sizeof(uint32_t);
    if (mode16) affine_rows_kernel<16><<<dim3(gx, gy), kThreads, smem, stream>>>(p);
    else affine_rows_kernel<8><<<dim3(gx, gy), kThreads, smem, stream>>>(p);
    count_launch();
    NASR_CUDA(cudaGetLastError());
  }
  return NASR_OK;
}

static void dw_grid(long long rows, int K, int C, int sms, int* gx, int* gy, int* gz) {
  *gy = (K + kKB - 1) / kKB;
  *gz = (C + kNB - 1) / kNB;
  const long long slabs = (rows + kSlab - 1) / kSlab;
  const int want = max(1, sms / ((*gy) * (*gz)));
  *gx = (int)max(1LL, min((long long)want, slabs));
}

int workspace_bytes(long long rows, int K, int C, size_t* out) {
  int sms = 0;
  NASR_CUDA(device_sm_count(&sms));
  int gx, gy, gz;
  dw_grid(rows, K, C, sms, &gx, &gy, &gz);
  *out = sizeof(float) * (size_t)gx * ((size_t)gy * kKB + 1) * ((size_t)gz * kNB);
  return NASR_OK;
}

int forward(const float* H, long long rows, int K, long long ldh, const float* W, const float* bias, int C,
            float* logits, long long ldl, cudaStream_t stream) {
  if (rows == 0) return NASR_OK;
  // Wt[n = c][k] = W[k*C + c]
  return launch_rows(H, ldh, rows, K, W, 1, C, C, bias, logits, ldl, stream);
}

int backward(const float* H, long long rows, int K, long long ldh, const float* W, int C, const float* dL,
             long long ldd, float* dH, long long lddh, float* dW, float* db, void* ws, size_t ws_bytes,
             cudaStream_t stream) {
  if (dH && rows > 0) {
    // dH[r, k] = sum_c dL[r, c] * W[k, c]: Wt[n = k][contraction c] = W[k*C + c]
    const int rc = launch_rows(dL, ldd, rows, C, W, C, 1, K, nullptr, dH, lddh, stream);
    if (rc != NASR_OK) return rc;
  }
  if (dW || db) {
    if (rows == 0) {
      if (dW) NASR_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)K * C, stream));
      if (db) NASR_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)C, stream));
      return NASR_OK;
    }
    int sms = 0;
    NASR_CUDA(device_sm_count(&sms));
    int gx, gy, gz;
    dw_grid(rows, K, C, sms, &gx, &gy, &gz);
    const size_t need = sizeof(float) * (size_t)gx * ((size_t)gy * kKB + 1) * ((size_t)gz * kNB);
    if (!ws || ws_bytes < need) {
      set_error("nasr_affine_backward_f32: workspace of %zu bytes, %zu needed (nasr_affine_workspace_bytes)",
                ws_bytes, need);
      return NASR_ERR_WORKSPACE_TOO_SMALL;
    }
    DwParams p;
    p.H = H;
    p.ldh = ldh;
    p.dL = dL;
    p.ldd = ldd;
    p.rows = rows;
    p.K = K;
    p.C = C;
    p.part = static_cast<float*>(ws);
    p.vec_h = (((uintptr_t)H & 15) == 0) && (ldh % 4 == 0) && (K % 4 == 0);
    p.vec_d = (((uintptr_t)dL & 7) == 0) && (ldd % 2 == 0) && (C % 2 == 0);
    NASR_CUDA((ensure_max_dynamic_smem<affine_dw_kernel>((int)kDwSmem)));
    affine_dw_kernel<<<dim3(gx, gy, gz), kThreads, kDwSmem, stream>>>(p);
    count_launch();
    NASR_CUDA(cudaGetLastError());
    const long long n = (long long)(K + 1) * C;
    affine_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
        p.part, gx, (long long)gy * kKB + 1, (long long)gz * kNB, K, C, dW, db);
    count_launch();
    NASR_CUDA(cudaGetLastError());
  }
  return NASR_OK;
}

}  // namespace affine
}  // namespace nasr

namespace nasr {
namespace affine_tc {   // csrc/affine_tc.cu: the tcgen05 / TMA / TMEM forward
bool eligible(const float* H, long long rows, int K, long long ldh, int C);
int forward(const float* H, long long rows, int K, long long ldh, const float* W, const float* bias, int C,
            float* logits, long long ldl, cudaStream_t stream);
}  // namespace affine_tc
namespace affine_tc_dh {   // csrc/affine_tc_dh.cu: dH = dL.W^T on tcgen05 (both parts of dL in TMEM)
bool eligible(const float* dL, long long rows, int K, long long ldd, int C);
int dh(const float* dL, long long rows, int K, const float* W, int C, float* dH, long long ldh, cudaStream_t stream);
}  // namespace affine_tc_dh
namespace affine_tc_dw {   // csrc/affine_tc_dw.cu: dW = H^T.dL and db on tcgen05 (H turned on its way into TMEM)
bool eligible(const float* H, long long rows, int K, long long ldh, long long ldd, int C);
size_t workspace_bytes(long long rows, int sms);
int dw(const float* H, long long rows, int K, long long ldh, const float* dL, int C, float* dW, float* db, void* ws,
       size_t ws_bytes, cudaStream_t stream);
}  // namespace affine_tc_dw
}  // namespace nasr

using namespace nasr;

// The forward runs on the tcgen05 kernel (csrc/affine_tc.cu) wherever its shape rules allow; NASR_AFFINE_TC=0 in the
// environment (read at every call: tests switch it) keeps everything on the mma.sync kernels.
static bool use_tcgen05() {
  const char* e = getenv("NASR_AFFINE_TC");
  return !(e && e[0] == '0');
}

int nasr_affine_workspace_bytes(long long rows, int K, int C, size_t* out_bytes) {
  NASR_CHECK_ARG(out_bytes, "nasr_affine_workspace_bytes: out_bytes is NULL");
  NASR_CHECK_ARG(rows >= 0 && K >= 1 && C >= 1, "nasr_affine_workspace_bytes: bad shape rows=%lld K=%d C=%d", rows, K,
                 C);
  const int rc = affine::workspace_bytes(rows, K, C, out_bytes);
  if (rc != NASR_OK) return rc;
  int sms = 0;
  NASR_CUDA(device_sm_count(&sms));
  *out_bytes = std::max(*out_bytes, affine_tc_dw::workspace_bytes(rows, sms));   // whichever kernel takes the call
  return NASR_OK;
}

int nasr_affine_logits_f32(const float* H, long long rows, int K, long long ldh, const float* W, const float* bias,
                           int C, float* logits, long long ldl, void* stream) {
  NASR_CHECK_ARG(rows >= 0 && K >= 1 && C >= 1, "nasr_affine_logits_f32: bad shape rows=%lld K=%d C=%d", rows, K, C);
  NASR_CHECK_ARG(ldh >= K && ldl >= C, "nasr_affine_logits_f32: row strides ldh=%lld ldl=%lld shorter than a row",
                 ldh, ldl);
  NASR_CHECK_ARG(W && (rows == 0 || (H && logits)), "nasr_affine_logits_f32: NULL argument");
  if (use_tcgen05() && affine_tc::eligible(H, rows, K, ldh, C))
    return affine_tc::forward(H, rows, K, ldh, W, bias, C, logits, ldl, static_cast<cudaStream_t>(stream));
  return affine::forward(H, rows, K, ldh, W, bias, C, logits, ldl, static_cast<cudaStream_t>(stream));
}

int nasr_affine_backward_f32(const float* H, long long rows, int K, long long ldh, const float* W, int C,
                             const float* dlogits, long long ldd, float* dH, long long lddh, float* dW, float* db,
                             void* workspace, size_t workspace_bytes, void* stream) {
  NASR_CHECK_ARG(rows >= 0 && K >= 1 && C >= 1, "nasr_affine_backward_f32: bad shape rows=%lld K=%d C=%d", rows, K,
                 C);
  NASR_CHECK_ARG(ldd >= C && (!dH || lddh >= K) && (!(dW || db) || ldh >= K),
                 "nasr_affine_backward_f32: a row stride is shorter than its row");
  NASR_CHECK_ARG(rows == 0 || dlogits, "nasr_affine_backward_f32: dlogits is NULL");
  NASR_CHECK_ARG(!dH || W, "nasr_affine_backward_f32: dH needs W");
  NASR_CHECK_ARG(!(dW || db) || rows == 0 || H, "nasr_affine_backward_f32: dW / db need H");
  if (dH && use_tcgen05() && affine_tc_dh::eligible(dlogits, rows, K, ldd, C)) {
    const int rc = affine_tc_dh::dh(dlogits, rows, K, W, C, dH, lddh, static_cast<cudaStream_t>(stream));
    if (rc != NASR_OK) return rc;
    dH = nullptr;   // done
  }
  if ((dW || db) && use_tcgen05() && affine_tc_dw::eligible(H, rows, K, ldh, ldd, C)) {
    const int rc = affine_tc_dw::dw(H, rows, K, ldh, dlogits, C, dW, db, workspace, workspace_bytes,
                                    static_cast<cudaStream_t>(stream));
    if (rc != NASR_OK) return rc;
    dW = nullptr;
    db = nullptr;
  }
  return affine::backward(H, rows, K, ldh, W, C, dlogits, ldd, dH, lddh, dW, db, workspace, workspace_bytes,
                          static_cast<cudaStream_t>(stream));
}
