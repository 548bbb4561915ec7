// Greedy CTC decode, Levenshtein label error rate and the small batch reductions, for sm_100a.
//
// Replaces tf.nn.ctc_greedy_decoder (reference networks/tfnetwork.py:63) and
// tf.edit_distance(tf.cast(model, tf.int32), labels) + reduce_mean (networks/tfnetwork.py:66-70);
// semantics per SURVEY.md Appendix A.2 / A.3.  Integer results are bit-exact by construction
// (first-index argmax on the raw logits, exact integer DP); neg_sum_logits is accumulated in frame
// order in fp32 like TF's scalar loop, so it is bit-exact too.
#include <type_traits>

#include "nasr_common.cuh"

namespace nasr {
namespace {

constexpr int kDecodeThreads = 256;
constexpr int kLerPending = -2147483647 - 1;  // dist[b] of an utterance edit_distance_lanes_kernel left to the warp kernel
constexpr int kDecodeChunk = 2048;  // frames staged per pass (argmax ids + max logits in shared memory)

// One CTA per utterance.  Phase 1: one warp per frame finds the first-index argmax of the raw logits
// (coalesced row read).  Phase 2: warp 0 collapses (drop blank, merge repeats) with ballot compaction
// while lane 0 of warp 1 accumulates -max in frame order.
// (Measured and dropped, end of round 2: the arg-max of narrow rows as its own kernel over tiles of 16 frames x 16
// utterances -- contiguous 2.4 KB pieces instead of rows a frame apart -- with this kernel as the collapse: 14 + 10 us
// against 26 us for this kernel alone under ncu, 0.0897 against 0.0905 ms for decode + label error rate at cfg3.)
// packed != 0: phase 1 was done by greedy_argmax_wide_kernel, which left (max logit bits << 32 | class id) of frame
// t in hyp[b][t]; the collapse then compacts the row in place (it never writes past what it has read).
__global__ void __launch_bounds__(kDecodeThreads)
greedy_decode_kernel(const float* __restrict__ logits, int T, int B, int C, long long st_t, long long st_b,
                     const int32_t* __restrict__ seq_len, int blank, int merge_repeated,
                     int64_t* hyp, int32_t* __restrict__ hyp_len, float* __restrict__ neg_sum_logits, int packed) {
  __shared__ int s_id[kDecodeChunk];
  __shared__ float s_mx[kDecodeChunk];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  int Tb = seq_len[b];
  Tb = max(0, min(T, Tb));
  int count = 0;       // warp 0: symbols emitted so far
  int carry_prev = -1; // warp 0: id of the previous frame
  float acc = 0.f;     // warp 1 lane 0
  for (int base = 0; base < Tb; base += kDecodeChunk) {
    const int n = min(kDecodeChunk, Tb - base);
    if (packed) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long v = (unsigned long long)__ldcg(hyp + (size_t)b * T + base + i);
        s_id[i] = (int)(unsigned)(v & 0xffffffffull);
        s_mx[i] = __uint_as_float((unsigned)(v >> 32));
      }
    } else if (C <= 64) {
      // narrow rows (a row is 152 B at C=38): 8 lanes per frame, 4 frames per warp instruction and 4 such
      // groups in flight, so that 32 independent loads per lane cover the DRAM latency
      const int sub = lane & 7, rl = lane >> 3;
      for (int g0 = warp; g0 * 4 < n; g0 += nw * 4) {
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = min((g0 + u * nw) * 4 + rl, n - 1);
          const float* x = logits + (size_t)(base + i) * st_t + (size_t)b * st_b;
#pragma unroll
          for (int e = 0; e < 8; e++) v[u][e] = __ldg(x + min(sub + 8 * e, C - 1));
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = (g0 + u * nw) * 4 + rl;
          float m = v[u][0];
          int am = sub;
#pragma unroll
          for (int e = 1; e < 8; e++)
            if (sub + 8 * e < C && v[u][e] > m) {  // strict '>' keeps the earliest index within a lane
              m = v[u][e];
              am = sub + 8 * e;
            }
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, m, o);
            const int oa = __shfl_xor_sync(0xffffffffu, am, o);
            if (om > m || (om == m && oa < am)) {
              m = om;
              am = oa;
            }
          }
          if (sub == 0 && i < n) {
            s_id[i] = am;
            s_mx[i] = m;
          }
        }
      }
    } else {
      for (int i = warp; i < n; i += nw) {
        const float* x = logits + (size_t)(base + i) * st_t + (size_t)b * st_b;
        float m = -INFINITY;
        int am = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
          const float v = __ldg(x + c);
          if (am == 0x7fffffff || v > m) {  // strict '>' keeps the earliest index within a lane
            m = v;
            am = c;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float om = __shfl_xor_sync(0xffffffffu, m, o);
          const int oa = __shfl_xor_sync(0xffffffffu, am, o);
          if (om > m || (om == m && oa < am)) {
            m = om;
            am = oa;
          }
        }
        if (lane == 0) {
          s_id[i] = am;
          s_mx[i] = m;
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      for (int g = 0; g < n; g += 32) {
        const int i = g + lane;
        const bool valid = i < n;
        const int id = valid ? s_id[i] : -1;
        int prev = __shfl_up_sync(0xffffffffu, id, 1);
        if (lane == 0) prev = carry_prev;
        const bool emit = valid && id != blank && !(merge_repeated && id == prev);
        const unsigned mask = __ballot_sync(0xffffffffu, emit);
        if (emit) hyp[(size_t)b * T + count + __popc(mask & ((1u << lane) - 1))] = (int64_t)id;
        count += __popc(mask);
        const int last = min(31, n - g - 1);
        carry_prev = __shfl_sync(0xffffffffu, id, last);
      }
    } else if (warp == 1 && lane == 0) {
      // frame order (TF's scalar loop), eight shared-memory reads ahead of the chain of subtractions
      int i = 0;
      for (; i + 8 <= n; i += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = s_mx[i + u];
#pragma unroll
        for (int u = 0; u < 8; u++) acc -= v[u];
      }
      for (; i < n; i++) acc -= s_mx[i];
    }
    __syncthreads();
  }
  if (warp == 0 && lane == 0) hyp_len[b] = count;
  if (warp == 1 && lane == 0 && neg_sum_logits) neg_sum_logits[b] = acc;
}

// Wide rows (C > 64): the arg-max phase of the greedy decoder spread over gridDim.x CTAs per utterance, because one
// CTA per utterance cannot pull a batch of few, wide utterances through one SM each (B=32, C=3000: 192 MB).
// One warp per frame, 16-byte loads when the rows are 16-byte aligned (vec4), four loads in flight per lane;
// first-index arg-max like Eigen's maxCoeff; the result goes to hyp[b][t] as (max bits << 32 | id).
__global__ void __launch_bounds__(kDecodeThreads)
greedy_argmax_wide_kernel(const float* __restrict__ logits, int T, int C, long long st_t, long long st_b,
                          const int32_t* __restrict__ seq_len, int frames_per_cta, int vec4, int64_t* hyp) {
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  int Tb = seq_len[b];
  Tb = max(0, min(T, Tb));
  const int t0 = blockIdx.x * frames_per_cta, t1 = min(Tb, t0 + frames_per_cta);
  for (int t = t0 + warp; t < t1; t += nw) {
    const float* x = logits + (size_t)t * st_t + (size_t)b * st_b;
    float m = -INFINITY;
    int am = 0x7fffffff;
    if (vec4) {
      for (int c0 = 4 * lane; c0 < C; c0 += 4 * 128) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int c = c0 + 128 * u;
          v[u] = c < C ? __ldg(reinterpret_cast<const float4*>(x + c))
                       : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int c = c0 + 128 * u;
          const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
          for (int q = 0; q < 4; q++)
            if (c + q < C && (am == 0x7fffffff || e[q] > m)) {  // strict '>' keeps the earliest index within a lane
              m = e[q];
              am = c + q;
            }
        }
      }
    } else {
      for (int c0 = lane; c0 < C; c0 += 4 * 32) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = c0 + 32 * u < C ? __ldg(x + c0 + 32 * u) : -INFINITY;
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (c0 + 32 * u < C && (am == 0x7fffffff || v[u] > m)) {
            m = v[u];
            am = c0 + 32 * u;
          }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const int oa = __shfl_xor_sync(0xffffffffu, am, o);
      if (om > m || (om == m && oa < am)) {
        m = om;
        am = oa;
      }
    }
    if (lane == 0)
      hyp[(size_t)b * T + t] = (int64_t)(((unsigned long long)__float_as_uint(m) << 32) | (unsigned)am);
  }
}

// One warp per utterance.  Fast path: Myers' bit-vector algorithm in its block form -- lane w owns bits
// [32w, 32w+32) of the truth (pattern) axis: vertical deltas Pv/Mv of one column of the DP lattice as two
// words; per hypothesis symbol each lane does ~20 integer operations and hands the horizontal delta at the
// top of its block (-1, 0, +1) to the next lane.  The lanes run skewed by one symbol (lane w works on symbol
// s-w at step s), so the hand-off is one shuffle per step and the whole utterance takes |hyp| + W - 1 steps
// instead of the |hyp| + |truth| barrier-separated anti-diagonals of the plain DP.  Match masks Peq[sym][w]
// live in shared memory, indexed by symbol value.  Exact integers (tests compare with the oracle bit for bit).
// Truths of up to 512 symbols run TWO-ENDED: half the warp takes the first half of the hypothesis, the other half the
// reversed second half against the reversed truth, and the two last columns are joined (see below) -- the chain of
// dependent steps, which is what the kernel's time is (one warp per utterance, ~300 cycles per step), halves.
// (Measured and dropped: staging the hypothesis in shared memory, 0.128 -> 0.27 ms; a skew of two symbols per lane
// to take the shuffle off the chain, 0.128 -> 0.140 ms.)
// Slow path (truth longer than 1024 symbols, or symbol values too large for the table): anti-diagonal
// wavefront, three diagonals and both strings in shared memory.
template <typename HypT>
__global__ void __launch_bounds__(32)
edit_distance_kernel(const HypT* __restrict__ hyp, long hyp_stride, const int32_t* __restrict__ hyp_len,
                     const int32_t* __restrict__ hyp_offsets,
                     const int32_t* __restrict__ truth_values, const int32_t* __restrict__ truth_offsets,
                     int max_truth_len, int max_hyp_len, int table_words, int normalize, int only_pending,
                     int32_t* __restrict__ dist, float* __restrict__ ler) {
  extern __shared__ int sm[];
  const int b = blockIdx.x, lane = threadIdx.x;
  // second launch behind edit_distance_lanes_kernel: only the utterances that kernel declined
  if (only_pending && dist[b] != kLerPending) return;
  const int t0 = truth_offsets[b];
  const int m = truth_offsets[b + 1] - t0;  // truth length
  int n;                                     // hyp length
  const HypT* h;
  if (hyp_offsets) {
    h = hyp + hyp_offsets[b];
    n = hyp_offsets[b + 1] - hyp_offsets[b];
  } else {
    h = hyp + (size_t)b * hyp_stride;
    n = hyp_len[b];
  }
  int d = 0;
  if (m > max_truth_len || n > max_hyp_len) {
    d = -1;  // caller's maxima were wrong: refuse rather than overrun shared memory
  } else if (n == 0 || m == 0) {
    d = n + m;
  } else {
    // does the bit-vector path apply?  W words of 32 truth positions, one per lane; table of (max symbol+1) x W
    const int W = (m + 31) >> 5;
    int mx = -1, mn = 0;
    for (int j = lane; j < m; j += 32) {
      const int v = truth_values[t0 + j];
      mx = max(mx, v);
      mn = min(mn, v);
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    mn = __reduce_min_sync(0xffffffffu, mn);
    const long long need = (long long)(mx + 1) * W;
    if (W <= 32 && mn >= 0 && need <= table_words) {
      unsigned* peq = reinterpret_cast<unsigned*>(sm);  // [direction][sym][W]
      const int nsym = mx + 1;
      // Two-ended run (truths up to 512 symbols): lanes 0-15 take the first half of the hypothesis against the truth,
      // lanes 16-31 the REVERSED second half against the reversed truth, in the same instruction stream; the distance
      // is min over j of D_fwd[j] + D_bwd[m-j], the last columns of the two halves, which the vertical deltas of the
      // bit vectors spell out.  Halves the chain of |hyp| dependent steps (cfg3: 128 -> 70 us).
      const bool two = W <= 16 && n >= 32 && 2 * need + 2 * ((long long)m + 1) <= table_words;
      const int ntab = two ? 2 : 1;
      for (int i = lane; i < ntab * nsym * W; i += 32) peq[i] = 0u;
      __syncwarp();
      for (int j = lane; j < m; j += 32) {
        const int v = truth_values[t0 + j];
        atomicOr(&peq[v * W + (j >> 5)], 1u << (j & 31));
        if (two) {
          const int jr = m - 1 - j;
          atomicOr(&peq[(nsym + v) * W + (jr >> 5)], 1u << (jr & 31));
        }
      }
      __syncwarp();
      const int GW = two ? 16 : 32;                 // lanes per direction
      const int dir = two ? (lane >> 4) : 0, w = lane & (GW - 1);
      const int n1 = two ? n - n / 2 : n;           // symbols of the forward half (>= the backward half's)
      const int nd = dir ? n - n1 : n1;
      const unsigned* tab = peq + dir * nsym * W;
      unsigned Pv = 0xffffffffu, Mv = 0u;
      int score = m;
      const int lastw = W - 1;
      const int topbit = w == lastw ? ((m - 1) & 31) : 31;
      int hout = 0;
      // symbol i of this lane's direction: forward h[i], backward h[n-1-i]
      long long cnext = (w == 0 && nd > 0) ? (long long)h[dir ? n - 1 : 0] : 0;
      const int steps = n1 + W - 1;
      for (int s = 0; s < steps; s++) {
        const int hin_up = __shfl_up_sync(0xffffffffu, hout, 1, GW);
        const int i = s - w;
        const bool active = w < W && i >= 0 && i < nd;
        const long long c = cnext;
        // prefetch the symbol of the next step (lane w reads symbol s+1-w)
        const int inext = i + 1;
        if (w < W && inext >= 0 && inext < nd) cnext = (long long)h[dir ? n - 1 - inext : inext];
        if (active) {
          const int hin = w == 0 ? 1 : hin_up;
          unsigned Eq = (c >= 0 && c < nsym) ? tab[(int)c * W + w] : 0u;
          const unsigned Xv = Eq | Mv;
          if (hin < 0) Eq |= 1u;
          const unsigned Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
          unsigned Ph = Mv | ~(Xh | Pv);
          unsigned Mh = Pv & Xh;
          hout = (int)((Ph >> topbit) & 1u) - (int)((Mh >> topbit) & 1u);
          if (w == lastw) score += hout;
          Ph <<= 1;
          Mh <<= 1;
          if (hin < 0) Mh |= 1u;
          else if (hin > 0) Ph |= 1u;
          Pv = Mh | ~(Xv | Ph);
          Mv = Ph & Xv;
        }
      }
      if (!two) {
        d = __shfl_sync(0xffffffffu, score, lastw);
      } else {
        // last column of each half: D[0] = symbols consumed, D[j+1] - D[j] = (Pv bit j) - (Mv bit j)
        int* Df = sm + 2 * nsym * W;    // [m + 1]
        int* Db = Df + m + 1;           // [m + 1]
        const int nb = w < W ? min(32, m - 32 * w) : 0;
        const unsigned valid = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
        const int delta = __popc(Pv & valid) - __popc(Mv & valid);
        int incl = delta;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, incl, o, 16);
          if (w >= o) incl += up;
        }
        int* D = dir ? Db : Df;
        int v = nd + incl - delta;      // D[32 w]
        for (int bit = 0; bit < nb; bit++) {
          D[32 * w + bit] = v;
          v += (int)((Pv >> bit) & 1u) - (int)((Mv >> bit) & 1u);
        }
        if (w == lastw) D[m] = v;
        __syncwarp();
        int best = 0x7fffffff;
        for (int j = lane; j <= m; j += 32) best = min(best, Df[j] + Db[m - j]);
        d = __reduce_min_sync(0xffffffffu, best);
      }
    } else {
      int* tr = sm;                       // [max_truth_len]
      int* hy = tr + max_truth_len;       // [max_hyp_len]
      int* d0 = hy + max_hyp_len;         // three diagonals, indexed by truth position j in [0, m]
      int* d1 = d0 + max_truth_len + 1;
      int* d2 = d1 + max_truth_len + 1;
      for (int j = lane; j < m; j += 32) tr[j] = truth_values[t0 + j];
      for (int i = lane; i < n; i += 32) hy[i] = (int)h[i];  // the reference casts int64 -> int32 first
      __syncwarp();
      // diagonal k holds cells (i, j) with i + j = k, i over hyp [0,n], j over truth [0,m]
      int* pp = d0;  // diagonal k-2
      int* pv = d1;  // diagonal k-1
      int* cu = d2;  // diagonal k
      for (int k = 0; k <= n + m; k++) {
        const int jlo = max(0, k - n), jhi = min(m, k);
        for (int j = jlo + lane; j <= jhi; j += 32) {
          const int i = k - j;
          int v;
          if (i == 0) {
            v = j;
          } else if (j == 0) {
            v = i;
          } else {
            v = min(min(pv[j] + 1, pv[j - 1] + 1), pp[j - 1] + (hy[i - 1] != tr[j - 1]));
          }
          cu[j] = v;
        }
        __syncwarp();
        int* tmp = pp;
        pp = pv;
        pv = cu;
        cu = tmp;
      }
      d = pv[m];
    }
  }
  if (lane == 0) {
    dist[b] = d;
    float r;
    if (!normalize) {
      r = (float)d;
    } else if (m == 0) {
      r = d ? INFINITY : 0.f;
    } else {
      r = (float)d / (float)m;
    }
    ler[b] = r;
  }
}

// ---- narrow symbol ranges, truths of up to 224 symbols: one LANE per (utterance, direction) -----------------------
// The warp-per-utterance kernel above spends its time on the chain of |hyp|/2 dependent steps, each a shuffle plus ~15
// dependent integer operations in a warp of which W lanes work (~300 cycles per step).  Here a lane keeps the WHOLE
// column of vertical deltas of its direction as one WT-word integer in registers (Myers' algorithm as published, the
// only thing that crosses words being the carry of one addition -- an add.cc / addc.cc chain -- and two funnel shifts),
// so a step needs no shuffle, and 32 lanes = 16 utterances x (forward half of the hypothesis | reversed second half
// against the reversed truth) fill the warp.  Match masks: table[symbol][word][lane] in shared memory, so that lane l
// reads bank l whatever its symbol.  No running score: D[n][m] = n + popc(Pv) - popc(Mv) of the last column, and the two
// halves are joined as in the kernel above (min over j of D_fwd[j] + D_bwd[m-j], spelt out by the vertical deltas).
// Declines (dist = kLerPending, picked up by the kernel above in the same stream) when the CTA's symbol values or
// hypothesis lengths do not fit shared memory (96 KB: the table, and every hypothesis as 16-bit table rows).  Exact
// integers.  Eight warps share the set-up (each of its loops waits on loads), warp 0 runs the recursion.
// Measured at cfg3 (B=256, hypotheses of ~950 symbols, truths of 100-200): 0.058 ms per call including the second
// launch, against 0.077 ms for the warp-per-utterance kernel alone; 87 instructions per symbol, of which ~72 on the
// integer pipe (LOP3 / SHF / IADD3, 16 lanes a cycle: two cycles each) -- the loop is bound by that pipe
// (profiles/r2_ler_lanes.md has the steps that did not pay: per-lane global loads of the symbols, cp.async rings).
template <int WT>
__device__ __forceinline__ void add_words(const unsigned (&a)[WT], const unsigned (&b)[WT], unsigned (&s)[WT]) {
  static_assert(WT == 2 || WT == 4 || WT == 7, "word counts of the lanes kernel");
  if constexpr (WT == 2) {
    asm("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, %5;"
        : "=r"(s[0]), "=r"(s[1])
        : "r"(a[0]), "r"(a[1]), "r"(b[0]), "r"(b[1]));
  } else if constexpr (WT == 4) {
    asm("add.cc.u32 %0, %4, %8;\n\taddc.cc.u32 %1, %5, %9;\n\taddc.cc.u32 %2, %6, %10;\n\taddc.u32 %3, %7, %11;"
        : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
  } else {
    asm("add.cc.u32 %0, %7, %14;\n\taddc.cc.u32 %1, %8, %15;\n\taddc.cc.u32 %2, %9, %16;\n\t"
        "addc.cc.u32 %3, %10, %17;\n\taddc.cc.u32 %4, %11, %18;\n\taddc.cc.u32 %5, %12, %19;\n\t"
        "addc.u32 %6, %13, %20;"
        : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(b[0]), "r"(b[1]),
          "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]));
  }
}

constexpr int kLanesWarps = 8;  // warps that set a CTA up (tables, symbol rows); warp 0 alone runs the recursion
template <typename HypT, int WT, int UPW>
__global__ void __launch_bounds__(32 * kLanesWarps)
edit_distance_lanes_kernel(const HypT* __restrict__ hyp, long hyp_stride, const int32_t* __restrict__ hyp_len,
                           const int32_t* __restrict__ hyp_offsets, const int32_t* __restrict__ truth_values,
                           const int32_t* __restrict__ truth_offsets, int max_truth_len, int max_hyp_len, int B,
                           int table_words, int normalize, int32_t* __restrict__ dist, float* __restrict__ ler) {
  extern __shared__ unsigned smu[];
  __shared__ int s_range[2];
  // every warp holds the same per-lane view (lane -> utterance, direction); the eight warps share the set-up work --
  // each of its loops waits on loads, and one warp alone has too few in flight -- then warp 0 runs the recursion
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, dir = lane & 1;
  const int b = blockIdx.x * UPW + (lane >> 1);
  const bool have = lane < 2 * UPW && b < B;
  int t0 = 0, m = 0, n = 0;
  const HypT* h = hyp;
  if (have) {
    t0 = truth_offsets[b];
    m = truth_offsets[b + 1] - t0;
    if (hyp_offsets) {
      h = hyp + hyp_offsets[b];
      n = hyp_offsets[b + 1] - hyp_offsets[b];
    } else {
      h = hyp + (size_t)b * hyp_stride;
      n = hyp_len[b];
    }
  }
  const bool refuse = have && (m > max_truth_len || n > max_hyp_len);  // as the kernel above: d = -1
  const bool work = have && !refuse && n > 0 && m > 0;
  // symbol range of the CTA's truths (each utterance once: its even lane; utterances dealt to the warps)
  if (threadIdx.x == 0) {
    s_range[0] = -1;
    s_range[1] = 0;
  }
  __syncthreads();
  int mx = -1, mn = 0;
  for (int t = 2 * wid; t < 2 * UPW; t += 2 * kLanesWarps) {
    const int mt = __shfl_sync(0xffffffffu, work ? m : 0, t), tt0 = __shfl_sync(0xffffffffu, t0, t);
    for (int j = lane; j < mt; j += 32) {
      const int v = truth_values[tt0 + j];
      mx = max(mx, v);
      mn = min(mn, v);
    }
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  mn = __reduce_min_sync(0xffffffffu, mn);
  if (lane == 0) {
    atomicMax(&s_range[0], mx);
    atomicMin(&s_range[1], mn);
  }
  __syncthreads();
  mx = s_range[0];
  mn = s_range[1];
  const int nsym = mx + 1;                       // rows 0..nsym-1 of the table, row nsym stays zero (any other symbol)
  const int tab_words = (nsym + 1) * WT * 32;
  int pw = (max_hyp_len + 1) / 2 + 1;            // words per utterance row of 16-bit symbol rows; odd: the forward
  pw |= 1;                                       // lanes of a warp, all at the same position, hit different banks
  const long long need = (long long)tab_words + 2 * WT * 32 + (long long)UPW * pw + 1;  // table, last columns, rows
  const bool decline = mn < 0 || nsym >= 65535 || need > table_words;
  int d = refuse ? -1 : n + m;                   // n == 0 or m == 0: n + m
  if (decline) {
    if (work) d = kLerPending;
  } else if (nsym > 0) {
    {
      uint4* z = reinterpret_cast<uint4*>(smu);
      for (int i = threadIdx.x; i < tab_words / 4; i += 32 * kLanesWarps) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    {
      // every lane fills its own table (its truth, reversed for the backward direction), an eighth of it in each
      // warp: bank = lane, and the reductions return nothing, so they queue behind each other without waiting
      const int mw = work ? m : 0;
      const int per = (mw + kLanesWarps - 1) / kLanesWarps;
      const int j1 = min(mw, (wid + 1) * per);
#pragma unroll 4
      for (int j = wid * per; j < j1; j++) {
        const int v = truth_values[t0 + j];
        const int pos = dir ? mw - 1 - j : j;
        atomicOr(&smu[(v * WT + (pos >> 5)) * 32 + lane], 1u << (pos & 31));
      }
    }
    // Hypotheses: the warps turn each utterance's symbols into 16-bit table rows in shared memory with coalesced
    // loads, eight in flight per lane (row nsym = the zero row for a symbol no truth of the CTA holds); a lane of the
    // recursion then needs one 2-byte read per symbol.  (A lane fetching its own row from global memory touches 32
    // lines per instruction, and the loop is bound by the integer pipe -- 16 lanes a cycle: every instruction beside
    // the bit-vector arithmetic counts.)
    unsigned short* soff = reinterpret_cast<unsigned short*>(smu + tab_words + 2 * WT * 32);
    const int P16 = 2 * pw;
    for (int u = wid; u < UPW; u += kLanesWarps) {
      const int nu = __shfl_sync(0xffffffffu, work ? n : 0, 2 * u);
      const unsigned long long hp = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)h, 2 * u);
      const HypT* hu = reinterpret_cast<const HypT*>((uintptr_t)hp);
      for (int i = lane; i < nu; i += 32 * 8) {
        long long c[8];
#pragma unroll
        for (int q = 0; q < 8; q++) c[q] = i + 32 * q < nu ? (long long)__ldg(hu + i + 32 * q) : 0ll;
#pragma unroll
        for (int q = 0; q < 8; q++)
          if (i + 32 * q < nu)
            soff[u * P16 + i + 32 * q] =
                (unsigned short)((unsigned long long)c[q] < (unsigned long long)nsym ? (int)c[q] : nsym);
      }
    }
    if (threadIdx.x == 0) soff[UPW * P16] = (unsigned short)nsym;   // what a lane without work reads
    __syncthreads();
    if (wid != 0) return;
    const int n1 = n - n / 2;                    // symbols of the forward half (>= the backward half's)
    const int nd = work ? (dir ? n - n1 : n1) : 0;
    const int steps = __reduce_max_sync(0xffffffffu, nd);
    const int smin = min(steps, __reduce_min_sync(0xffffffffu, work ? nd : 0x7fffffff));  // every lane with work has this many
    const unsigned* tl = smu + lane;
    unsigned Pv[WT], Mv[WT];
#pragma unroll
    for (int k = 0; k < WT; k++) {
      Pv[k] = 0xffffffffu;
      Mv[k] = 0u;
    }
    // symbol i of this lane's direction: forward h[i], backward h[n-1-i]; a lane without work reads the zero row
    const unsigned short* sp = soff + (work ? (lane >> 1) * P16 + (dir ? n - 1 : 0) : UPW * P16);
    const int sgn = work ? (dir ? -1 : 1) : 0;
    const int ilast = nd > 0 ? nd - 1 : 0;
    auto row_of = [&](int i) -> int { return (int)sp[sgn * min(i, ilast)] * (WT * 32); };
    constexpr int U = 4;
    int off[U];
#pragma unroll
    for (int u = 0; u < U; u++) off[u] = row_of(u);
    unsigned EqN[WT];
#pragma unroll
    for (int k = 0; k < WT; k++) EqN[k] = tl[off[0] + k * 32];
    // one round = U symbols; TAIL: lanes past their last symbol keep their column (a select per word), before that
    // (every lane with work has symbols left; lanes without work compute on the zero row and nobody reads them) none
    auto round = [&](auto tail, int i0) {
      constexpr bool TAIL = decltype(tail)::value;
      int offn[U];
#pragma unroll
      for (int u = 0; u < U; u++) offn[u] = row_of(i0 + U + u);
#pragma unroll
      for (int u = 0; u < U; u++) {
        unsigned Eq[WT], Xv[WT], A[WT], Sm[WT], Ph[WT], Mh[WT];
        const int onext = u + 1 < U ? off[u + 1 < U ? u + 1 : 0] : offn[0];
#pragma unroll
        for (int k = 0; k < WT; k++) {
          Eq[k] = EqN[k];
          EqN[k] = tl[onext + k * 32];
        }
#pragma unroll
        for (int k = 0; k < WT; k++) {
          Xv[k] = Eq[k] | Mv[k];
          A[k] = Eq[k] & Pv[k];
        }
        add_words<WT>(A, Pv, Sm);
#pragma unroll
        for (int k = 0; k < WT; k++) {
          const unsigned Xh = (Sm[k] ^ Pv[k]) | Eq[k];
          Ph[k] = Mv[k] | ~(Xh | Pv[k]);
          Mh[k] = Pv[k] & Xh;
        }
        const bool active = !TAIL || i0 + u < nd;
#pragma unroll
        for (int k = WT - 1; k >= 0; k--) {
          // one row down: the horizontal delta entering row 0 is +1 (D[i][0] = i)
          const unsigned ph = __funnelshift_l(k ? Ph[k - 1] : 0x80000000u, Ph[k], 1);
          const unsigned mh = __funnelshift_l(k ? Mh[k - 1] : 0u, Mh[k], 1);
          if (active) {
            Pv[k] = mh | ~(Xv[k] | ph);
            Mv[k] = ph & Xv[k];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) off[u] = offn[u];
    };
    int i0 = 0;
    for (; i0 + U <= smin; i0 += U) round(std::false_type{}, i0);
    for (; i0 < steps; i0 += U) round(std::true_type{}, i0);
    // last columns to shared memory; join: S[j] = D_fwd[j] + D_bwd[m-j], S[0] = n1 + D_bwd[m],
    // S[j+1] - S[j] = (Pf_j - Mf_j) - (Pb_(m-1-j) - Mb_(m-1-j))
    unsigned* xv = smu + tab_words;              // [2 * WT][32]
#pragma unroll
    for (int k = 0; k < WT; k++) {
      xv[k * 32 + lane] = Pv[k];
      xv[(WT + k) * 32 + lane] = Mv[k];
    }
    __syncwarp();
    // S[j] = n + (sum of the forward deltas below j) + (sum of the backward deltas below m - j); the two lanes of an
    // utterance take the halves of the range of j (split at a multiple of 32), each from its own S[j0]
    int best = 0x7fffffff;
    if (work) {
      const int lf = lane & ~1, lb = lane | 1;
      auto below = [&](int l, int x) -> int {      // sum of lane l's vertical deltas at positions < x
        int acc = 0;
#pragma unroll
        for (int k = 0; k < WT; k++) {
          const int nb = min(32, max(0, x - 32 * k));
          const unsigned valid = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
          acc += __popc(xv[k * 32 + l] & valid) - __popc(xv[(WT + k) * 32 + l] & valid);
        }
        return acc;
      };
      const int mh = (m / 2) & ~31;
      const int j0 = dir ? mh : 0, j1 = dir ? m : mh;
      int S = n + below(lf, j0) + below(lb, m - j0);
      best = S;
      int wq = (m - 1 - j0) >> 5, bq = (m - 1 - j0) & 31;
      unsigned Pb = xv[wq * 32 + lb], Mb = xv[(WT + wq) * 32 + lb], Pf = 0u, Mf = 0u;
      for (int j = j0; j < j1; j++) {
        if ((j & 31) == 0) {
          Pf = xv[(j >> 5) * 32 + lf];
          Mf = xv[(WT + (j >> 5)) * 32 + lf];
        }
        S += (int)(Pf & 1u) - (int)(Mf & 1u) - (int)((Pb >> bq) & 1u) + (int)((Mb >> bq) & 1u);
        Pf >>= 1;
        Mf >>= 1;
        best = min(best, S);
        if (--bq < 0 && wq > 0) {
          bq = 31;
          --wq;
          Pb = xv[wq * 32 + lb];
          Mb = xv[(WT + wq) * 32 + lb];
        }
      }
    }
    best = min(best, __shfl_xor_sync(0xffffffffu, best, 1));
    if (work) {
      d = best;
    }
  }
  if (wid == 0 && have && dir == 0) {
    dist[b] = d;
    if (d != kLerPending) {
      float r;
      if (!normalize) {
        r = (float)d;
      } else if (m == 0) {
        r = d ? INFINITY : 0.f;
      } else {
        r = (float)d / (float)m;
      }
      ler[b] = r;
    }
  }
}

__global__ void hyp_to_sparse_kernel(const int64_t* __restrict__ hyp, long hyp_stride,
                                     const int32_t* __restrict__ offs, int B,
                                     int64_t* __restrict__ indices, int64_t* __restrict__ values) {
  const int b = blockIdx.x;
  const int o = offs[b], n = offs[b + 1] - o;
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    indices[2 * (size_t)(o + p)] = b;
    indices[2 * (size_t)(o + p) + 1] = p;
    values[o + p] = hyp[(size_t)b * hyp_stride + p];
  }
}

__global__ void dense_shape_kernel(const int32_t* __restrict__ offs, int B, int64_t* __restrict__ shape) {
  __shared__ int red[32];
  int mx = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) mx = max(mx, offs[b + 1] - offs[b]);
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    mx = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0;
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (threadIdx.x == 0) {
      shape[0] = B;
      shape[1] = mx;
    }
  }
}

// Fixed-order reduction (thread-strided partials, then a shared-memory tree): deterministic.
__global__ void __launch_bounds__(256)
batch_sums_kernel(const float* __restrict__ loss, const float* __restrict__ ler,
                  const int32_t* __restrict__ dist, int B, double* __restrict__ sums) {
  __shared__ double s[3][256];
  double a = 0, c = 0, d = 0;
  for (int b = threadIdx.x; b < B; b += 256) {
    if (loss) a += (double)loss[b];
    if (ler) c += (double)ler[b];
    if (dist) d += (double)dist[b];
  }
  s[0][threadIdx.x] = a;
  s[1][threadIdx.x] = c;
  s[2][threadIdx.x] = d;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s[0][threadIdx.x] += s[0][threadIdx.x + o];
      s[1][threadIdx.x] += s[1][threadIdx.x + o];
      s[2][threadIdx.x] += s[2][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    sums[0] = s[0][0];
    sums[1] = s[1][0];
    sums[2] = s[2][0];
    sums[3] = (double)B;
  }
}

template <typename HypT>
int launch_edit_distance(const HypT* hyp, long hyp_stride, const int32_t* hyp_len,
                         const int32_t* hyp_offsets, const int32_t* truth_values,
                         const int32_t* truth_offsets, int max_truth_len, int max_hyp_len, int B,
                         int normalize, int32_t* dist, float* ler, cudaStream_t stream) {
  const size_t dp = sizeof(int) * ((size_t)max_truth_len + max_hyp_len + 3 * ((size_t)max_truth_len + 1));
  const size_t table = 64 * 1024;  // match masks of the bit-vector path: (max symbol + 1) * ceil(|truth|/32) words
  const size_t smem = dp > table ? dp : table;
  if (smem > 200 * 1024) {
    set_error("nasr_edit_distance: max_truth_len=%d max_hyp_len=%d exceed shared memory", max_truth_len,
              max_hyp_len);
    return NASR_ERR_UNSUPPORTED;
  }
  // truths of up to 224 symbols: a lane per (utterance, direction) first; what it declines (symbol values that do not
  // fit its table) is left marked in dist[] for the warp-per-utterance kernel.  NASR_LER_LANES=0 switches it off
  // (read at every call, so tests and tools can compare the two).
  int only_pending = 0;
  const int WT = max_truth_len <= 64 ? 2 : (max_truth_len <= 128 ? 4 : (max_truth_len <= 224 ? 7 : 0));
  const char* e = getenv("NASR_LER_LANES");
  if (WT && !(e && e[0] == '0')) {
    constexpr int kLanesSmem = 96 * 1024;
    // 16 utterances per CTA (half-filled warps, 8 per CTA, were measured: no faster -- the integer pipe takes two
    // cycles for a warp instruction whatever the number of active lanes)
    const int grid = (B + 15) / 16;
#define NASR_LANES(W_)                                                                                          \
  do {                                                                                                          \
    NASR_CUDA((ensure_max_dynamic_smem<edit_distance_lanes_kernel<HypT, W_, 16>>(kLanesSmem)));                 \
    edit_distance_lanes_kernel<HypT, W_, 16><<<grid, 32 * kLanesWarps, kLanesSmem, stream>>>(                   \
        hyp, hyp_stride, hyp_len, hyp_offsets, truth_values, truth_offsets, max_truth_len, max_hyp_len, B,      \
        kLanesSmem / 4, normalize, dist, ler);                                                                  \
  } while (0)
    if (WT == 2) {
      NASR_LANES(2);
    } else if (WT == 4) {
      NASR_LANES(4);
    } else {
      NASR_LANES(7);
    }
#undef NASR_LANES
    count_launch();
    NASR_CUDA(cudaGetLastError());
    only_pending = 1;
  }
  NASR_CUDA((ensure_max_dynamic_smem<edit_distance_kernel<HypT>>(200 * 1024)));
  edit_distance_kernel<HypT><<<B, 32, smem, stream>>>(hyp, hyp_stride, hyp_len, hyp_offsets,
                                                     truth_values, truth_offsets, max_truth_len,
                                                     max_hyp_len, (int)(table / 4), normalize, only_pending, dist,
                                                     ler);
  count_launch();
  NASR_CUDA(cudaGetLastError());
  return NASR_OK;
}

}  // namespace

int greedy_decode(const float* logits, int T, int B, int C, long long st_t, long long st_b,
                  const int32_t* seq_len, int blank, int merge_repeated, int64_t* hyp, int32_t* hyp_len,
                  float* neg_sum_logits, cudaStream_t stream) {
  NASR_CHECK_ARG(st_t >= 0 && st_b >= 0, "nasr_ctc_greedy_decode: negative stride");
  NASR_CHECK_ARG(T >= 0 && B >= 0 && C >= 1, "nasr_ctc_greedy_decode: bad shape T=%d B=%d C=%d", T, B, C);
  if (B == 0) return NASR_OK;
  NASR_CHECK_ARG((logits || T == 0) && seq_len && hyp_len && (hyp || T == 0),
                 "nasr_ctc_greedy_decode: NULL argument");
  int packed = 0;
  if (C > 64 && T > 0) {
    // wide rows: arg-max over about four CTAs per SM (148 SMs), at least 32 frames each; then the collapse
    const int splits = max(1, min((592 + B - 1) / B, (T + 31) / 32));
    const int fpc = (T + splits - 1) / splits;
    const int vec4 = ((uintptr_t)logits & 15) == 0 && (C & 3) == 0 && (st_t & 3) == 0 && (st_b & 3) == 0;
    greedy_argmax_wide_kernel<<<dim3((T + fpc - 1) / fpc, B), kDecodeThreads, 0, stream>>>(logits, T, C, st_t, st_b,
                                                                                         seq_len, fpc, vec4, hyp);
    count_launch();
    packed = 1;
  }
  greedy_decode_kernel<<<B, kDecodeThreads, 0, stream>>>(logits, T, B, C, st_t, st_b, seq_len, blank,
                                                        merge_repeated, hyp, hyp_len, neg_sum_logits, packed);
  count_launch();
  NASR_CUDA(cudaGetLastError());
  return NASR_OK;
}

int edit_distance_dense(const int64_t* hyp, int hyp_stride, const int32_t* hyp_len,
                        const int32_t* truth_values, const int32_t* truth_offsets, int max_truth_len,
                        int B, int normalize, int32_t* dist, float* ler, cudaStream_t stream) {
  NASR_CHECK_ARG(B >= 0 && hyp_stride >= 0 && max_truth_len >= 0, "nasr_edit_distance: bad sizes");
  if (B == 0) return NASR_OK;
  NASR_CHECK_ARG(hyp_len && truth_offsets && dist && ler, "nasr_edit_distance: NULL argument");
  return launch_edit_distance<int64_t>(hyp, hyp_stride, hyp_len, nullptr, truth_values, truth_offsets,
                                       max_truth_len, hyp_stride, B, normalize, dist, ler, stream);
}

int edit_distance_csr(const int64_t* hyp_values, const int32_t* hyp_offsets, int max_hyp_len,
                      const int32_t* truth_values, const int32_t* truth_offsets, int max_truth_len,
                      int B, int normalize, int32_t* dist, float* ler, cudaStream_t stream) {
  NASR_CHECK_ARG(B >= 0 && max_truth_len >= 0 && max_hyp_len >= 0, "nasr_edit_distance_csr: bad sizes");
  if (B == 0) return NASR_OK;
  NASR_CHECK_ARG(hyp_offsets && truth_offsets && dist && ler, "nasr_edit_distance_csr: NULL argument");
  return launch_edit_distance<int64_t>(hyp_values, 0, nullptr, hyp_offsets, truth_values, truth_offsets,
                                       max_truth_len, max_hyp_len, B, normalize, dist, ler, stream);
}

// ---- device-side label ingest: the COO triple of sparse_tuple_from (reference utils.py:44-58) -> CSR ------------
// One block.  indices int64 [N,2] row-major (row, position), rows non-decreasing (tf.SparseTensor's canonical order,
// which sparse_tuple_from produces).  Rows [row0, row0 + B) are taken -- tf.sparse_split(axis=0)'s block of a tower
// (tfnetwork.py:97-99) is a row window -- and re-based to 0.  offsets[B+1] = prefix sum of the row lengths,
// values_out[n - first] = values[n] for the window's entries, info[0] = 1 if the rows are not ordered or a position is
// not the running index within its row, info[1] = longest row, info[2] = first entry of the window.
__global__ void coo_to_csr_kernel(const int64_t* __restrict__ indices, const int32_t* __restrict__ values, int N,
                                  int row0, int B, int32_t* __restrict__ offsets, int32_t* __restrict__ values_out,
                                  int32_t* __restrict__ info) {
  extern __shared__ int s_cnt[];   // [B + 1]
  __shared__ int s_bad, s_max, s_first;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i <= B; i += nt) s_cnt[i] = 0;
  if (tid == 0) {
    s_bad = 0;
    s_max = 0;
    s_first = N;
  }
  __syncthreads();
  for (int n = tid; n < N; n += nt) {
    const long long r = indices[2 * n], c = indices[2 * n + 1];
    if (n > 0) {
      const long long rp = indices[2 * n - 2], cp = indices[2 * n - 1];
      if (r < rp || (r == rp && c != cp + 1) || (r > rp && c != 0)) s_bad = 1;
    } else if (c != 0) {
      s_bad = 1;
    }
    if (r >= row0 && r < (long long)row0 + B) {
      atomicAdd(&s_cnt[(int)(r - row0) + 1], 1);
      atomicMin(&s_first, n);
    }
  }
  __syncthreads();
  // inclusive scan of the B row counts by one warp (B is a batch size: a few hundred)
  if (tid < 32) {
    int run = 0;
    for (int base = 1; base <= B; base += 32) {
      const int i = base + tid;
      const int v = i <= B ? s_cnt[i] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (tid >= o) inc += t;
      }
      if (i <= B) {
        s_cnt[i] = run + inc;
        atomicMax(&s_max, v);
      }
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncthreads();
  for (int i = tid; i <= B; i += nt) offsets[i] = s_cnt[i];
  const int first = s_first == N ? 0 : s_first, total = s_cnt[B];
  if (values_out)
    for (int n = tid; n < total; n += nt) values_out[n] = values[first + n];
  if (tid == 0) {
    info[0] = s_bad;
    info[1] = s_max;
    info[2] = first;
  }
}

int labels_coo_to_csr(const int64_t* indices, const int32_t* values, int N, int row0, int B, int32_t* offsets,
                      int32_t* values_out, int32_t* info, cudaStream_t stream) {
  NASR_CHECK_ARG(N >= 0 && B >= 0 && row0 >= 0 && offsets && info && (N == 0 || indices),
                 "nasr_labels_coo_to_csr: bad arguments");
  NASR_CHECK_ARG((size_t)(B + 1) * 4 <= 160 * 1024, "nasr_labels_coo_to_csr: batch %d too large", B);
  NASR_CUDA((ensure_max_dynamic_smem<coo_to_csr_kernel>(160 * 1024)));
  coo_to_csr_kernel<<<1, 1024, (size_t)(B + 1) * 4, stream>>>(indices, values, N, row0, B, offsets, values_out, info);
  count_launch();
  NASR_CUDA(cudaGetLastError());
  return NASR_OK;
}

int hyp_to_sparse(const int64_t* hyp, int hyp_stride, const int32_t* hyp_offsets, int B,
                  int64_t* indices, int64_t* values, int64_t* dense_shape, cudaStream_t stream) {
  NASR_CHECK_ARG(B >= 0, "nasr_hyp_to_sparse: bad B");
  NASR_CHECK_ARG(hyp_offsets && dense_shape, "nasr_hyp_to_sparse: NULL argument");
  if (B > 0 && indices && values) {
    hyp_to_sparse_kernel<<<B, 128, 0, stream>>>(hyp, hyp_stride, hyp_offsets, B, indices, values);
    count_launch();
  }
  dense_shape_kernel<<<1, 256, 0, stream>>>(hyp_offsets, B, dense_shape);
  count_launch();
  NASR_CUDA(cudaGetLastError());
  return NASR_OK;
}

int batch_sums(const float* loss, const float* ler, const int32_t* dist, int B, double* sums,
               cudaStream_t stream) {
  NASR_CHECK_ARG(B >= 0 && sums, "nasr_batch_sums: bad argument");
  batch_sums_kernel<<<1, 256, 0, stream>>>(loss, ler, dist, B, sums);
  count_launch();
  NASR_CUDA(cudaGetLastError());
  return NASR_OK;
}

}  // namespace nasr
