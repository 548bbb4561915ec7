// CTC beam search decoder for sm_100a.
//
// Replaces tf.nn.ctc_beam_search_decoder(logits, seq_len) — beam_width 100, top_paths 1, merge_repeated True —
// the decoder the reference's create_model actually runs in every train step (networks/tfnetwork.py:62,64);
// semantics per oracle/beam_oracle.py (TF 1.x core/util/ctc/ctc_beam_search.h restated).
//
// One CTA per utterance walks the frames.  Per frame:
//   1. log-softmax of the logits row in fp64 (next row prefetched into registers meanwhile);
//   2. every active prefix keeps itself: label' = (LSE(label, parent term) if its parent prefix is active) + lp[last],
//      blank' = total + lp[blank];
//   3. candidates = those W updated prefixes + every extension (prefix b, label l) that is not already an active
//      prefix: W*(C-1) scores  lp[l] + (l == last(b) ? blank(b) : total(b)).  They are never materialised: each
//      selection pass recomputes them from two shared-memory tables (3800 candidates at W=100, C=38);
//   4. the best W by (score, existing-before-new, prefix hash) are found with an MSB-first radix select over the
//      order-preserving 64-bit image of the fp64 score (11-bit digits, histogram in shared memory, starting at
//      the first bit where the candidates differ; exact ties go on through the hash), after TF's own pruning
//      (a full beam admits only extensions that beat its worst kept prefix, so prefixes whose best possible
//      extension cannot are skipped wholesale);
//   5. survivors are compacted into the other beam buffer; new prefixes get a node (parent node, label) in a
//      per-utterance trie in the workspace; each prefix finds its parent's slot by hash, and each parent gets
//      the bit set of labels whose extension is already active.
// Prefix identity is a 64-bit hash chain (root constant, child = mix(parent + K*(label+1))) plus the prefix length:
// two different prefixes of equal length colliding inside one beam is a 2^-64-per-pair event and would merge them.
// After the last frame the top_paths best prefixes are read back through the trie, repeats collapsed when
// merge_repeated (TF merges in the OUTPUT), and their log probabilities returned.
// Scores are fp64 (TF: fp32): the label sequences are what the reference consumes, and fp64 keeps the kernel
// and the oracle on the same side of every comparison that is not an exact tie.
#include <math.h>

#include "nasr_common.cuh"

namespace nasr {
namespace {

typedef unsigned long long u64;

constexpr int kBeamThreads = 512;
constexpr int kBeamWarps = kBeamThreads / 32;
constexpr int kBins = 2048;  // 11-bit digits
constexpr int kBinsPerThread = kBins / kBeamThreads;
constexpr u64 kRootHash = 0x243f6a8885a308d3ull;
constexpr int kPrefetch = 2;  // logits of the next frame held in registers: C <= 2*512

__host__ __device__ inline u64 mix64(u64 x) {
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
__host__ __device__ inline u64 child_hash(u64 h, int label) {
  return mix64(h + 0x9e3779b97f4a7c15ull * (u64)(label + 1));
}

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000ll); }

// order-preserving image of a double: a < b  <=>  okey(a) < okey(b); okey(v) > 0 for every v > -inf
__device__ __forceinline__ u64 okey(double v) {
  const long long b = __double_as_longlong(v);
  return (u64)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
}

__device__ __forceinline__ double lse2(double a, double b) {
  const double ninf = neg_inf();
  if (a == ninf) return b;
  if (b == ninf) return a;
  const double m = fmax(a, b), n = fmin(a, b);
  return m + log1p(exp(n - m));
}

__host__ __device__ inline size_t al16(size_t n) { return (n + 15) & ~(size_t)15; }

__host__ __device__ inline size_t beam_smem_bytes(int W, int C) {
  const int CW = (C + 31) / 32;
  size_t s = 0;
  s += al16(sizeof(double) * C) + al16(sizeof(float) * C);
  s += 3 * al16(sizeof(double) * 2 * W) + 2 * al16(sizeof(u64) * 2 * W) + 5 * al16(sizeof(int) * 2 * W);
  s += 3 * al16(sizeof(double) * W);
  s += al16(sizeof(int) * W);
  s += al16(sizeof(uint32_t) * (size_t)W * CW);
  s += al16(sizeof(int) * kBins);
  s += al16(sizeof(double) * 64) + al16(sizeof(u64) * 64) + al16(sizeof(int) * 64);
  return s;
}

template <typename T>
__device__ __forceinline__ T* carve(char*& p, size_t n) {
  T* r = reinterpret_cast<T*>(p);
  p += al16(n * sizeof(T));
  return r;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  return v;
}

struct Ctx {
  // frame constants every candidate evaluation needs
  const double *lp, *pb, *pt, *ut;
  const u64* hash;
  const int *last, *live;
  const uint32_t* mask;
  int n, C, CW, blank;
  u64 divM;  // floor(2^40 / C) + 1: j / C == (j * divM) >> 40 for j < 2^20
  double tau0;
};

// Candidate i of the frame: i < n is active prefix i itself (l = -1); otherwise extension (live[q], l) with
// i - n = q*C + l.
__device__ __forceinline__ void item_of(const Ctx& c, int i, int& b, int& l) {
  if (i < c.n) {
    b = i;
    l = -1;
    return;
  }
  const unsigned j = (unsigned)(i - c.n);
  const unsigned q = (unsigned)(((u64)j * c.divM) >> 40);
  l = (int)(j - q * (unsigned)c.C);
  b = c.live[q];
}

// Its score; false if it is not offered to the beam.
__device__ __forceinline__ bool eval_item(const Ctx& c, int i, double& v, int& b, int& l) {
  item_of(c, i, b, l);
  if (l < 0) {
    v = c.ut[b];
    return v > neg_inf();
  }
  if (l == c.blank) return false;
  if ((c.mask[b * c.CW + (l >> 5)] >> (l & 31)) & 1u) return false;
  v = c.lp[l] + (l == c.last[b] ? c.pb[b] : c.pt[b]);
  return v > c.tau0;
}

__device__ __forceinline__ u64 item_k2(const Ctx& c, int b, int l) {
  // active prefixes win ties against extensions (TF admits an extension only if it is strictly better than the
  // worst kept entry); then the smaller hash wins
  if (l < 0) return 0x8000000000000000ull | ((~c.hash[b]) >> 1);
  return (~child_hash(c.hash[b], l)) >> 1;
}

// Find, from the top bin down, the bin in which the running count reaches `need`.  hist is left zeroed.
// res[0] = bin, res[1] = candidates in the bins above it, res[2] = candidates in it.
__device__ __forceinline__ void find_bin(int* hist, int nb, int need, int* wsum, int* res, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  int loc[kBinsPerThread];
  int sum = 0;
#pragma unroll
  for (int k = 0; k < kBinsPerThread; k++) {
    const int r = tid * kBinsPerThread + k;
    loc[k] = 0;
    if (r < nb) {
      loc[k] = hist[nb - 1 - r];
      hist[nb - 1 - r] = 0;
    }
    sum += loc[k];
  }
  int incl = warp_incl_scan(sum, lane);
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  for (int w = 0; w < warp; w++) incl += wsum[w];
  const int excl = incl - sum;
  if (excl < need && need <= incl) {
    int cum = excl;
#pragma unroll
    for (int k = 0; k < kBinsPerThread; k++) {
      if (cum + loc[k] >= need) {
        res[0] = nb - 1 - (tid * kBinsPerThread + k);
        res[1] = cum;
        res[2] = loc[k];
        break;
      }
      cum += loc[k];
    }
  }
  __syncthreads();
}

// IPT > 0: every thread keeps the keys of its IPT candidates in registers between the selection passes
// (W*(C+1) <= IPT*512); IPT == 0: candidates are recomputed in every pass (any W*C the tables fit for).
template <int IPT>
__global__ void __launch_bounds__(kBeamThreads, 2)
ctc_beam_kernel(const float* __restrict__ logits, int T, int B, int C, long long st_t, long long st_b,
                const int32_t* __restrict__ seq_len, int blank, int W, int P, int merge_repeated,
                int64_t* hyp, int32_t* __restrict__ hyp_len, float* __restrict__ log_prob,
                int2* nodes_all, u64 divM) {
  extern __shared__ __align__(16) char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b_utt = blockIdx.x;
  const int CW = (C + 31) / 32;
  const double ninf = neg_inf();

  char* sp = smem_raw;
  double* s_lp = carve<double>(sp, C);
  float* s_xs = carve<float>(sp, C);
  // the two beam buffers are the halves [0,W) and [W,2W) of each array
  double* g_pb = carve<double>(sp, 2 * W);   // log P(prefix, ends in blank)
  double* g_pl = carve<double>(sp, 2 * W);   //                ends in its last label
  double* g_pt = carve<double>(sp, 2 * W);   //                either
  u64* g_hash = carve<u64>(sp, 2 * W);       // prefix identity
  u64* g_phash = carve<u64>(sp, 2 * W);      // parent prefix identity
  int* g_node = carve<int>(sp, 2 * W);
  int* g_len = carve<int>(sp, 2 * W);
  int* g_last = carve<int>(sp, 2 * W);
  int* g_plast = carve<int>(sp, 2 * W);
  int* g_pslot = carve<int>(sp, 2 * W);
  double* s_ub = carve<double>(sp, W);       // this frame's update of the active prefixes
  double* s_ul = carve<double>(sp, W);
  double* s_ut = carve<double>(sp, W);
  int* s_live = carve<int>(sp, W);           // prefixes whose extensions can still enter the beam
  uint32_t* s_mask = carve<uint32_t>(sp, (size_t)W * CW);  // [W][CW] labels whose extension is already active
  int* s_hist = carve<int>(sp, kBins);
  double* s_redd = carve<double>(sp, 64);
  u64* s_redu = carve<u64>(sp, 64);
  int* s_redi = carve<int>(sp, 64);
  // s_redi: [0..15] warp partials, [16..18] find_bin result, [20] new-beam counter, [21] node counter,
  //         [22] n_live, [32..47] find_bin warp sums;  s_redd: [0..15] max, [16..31] sum, [32] tau0

  int Tb = seq_len[b_utt];
  Tb = max(0, min(T, Tb));
  const float* xrow = logits + (size_t)b_utt * st_b;
  int2* nodes = nodes_all + (size_t)b_utt * ((size_t)T * W + 1);

  for (int i = tid; i < kBins; i += kBeamThreads) s_hist[i] = 0;
  for (int i = tid; i < CW; i += kBeamThreads) s_mask[i] = 0;
  if (tid == 0) {
    g_pb[0] = 0.0;
    g_pl[0] = ninf;
    g_pt[0] = 0.0;
    g_hash[0] = kRootHash;
    g_phash[0] = 0;
    g_node[0] = 0;
    g_len[0] = 0;
    g_last[0] = -1;
    g_plast[0] = -1;
    g_pslot[0] = -1;
    s_redi[21] = 1;
    nodes[0] = make_int2(-1, -1);
  }
  if (Tb > 0)
    for (int c = tid; c < C; c += kBeamThreads) s_xs[c] = __ldg(xrow + c);
  __syncthreads();

  int n = 1, cur = 0;
  for (int t = 0; t < Tb; t++) {
    const int ao = cur * W, no = (cur ^ 1) * W;  // offsets of the active and of the next beam buffer
    // next frame's row: in flight while this frame is searched
    float xr[kPrefetch];
    const bool pre = (t + 1 < Tb) && C <= kPrefetch * kBeamThreads;
    if (pre) {
      const float* xn = xrow + (size_t)(t + 1) * st_t;
#pragma unroll
      for (int k = 0; k < kPrefetch; k++) {
        const int c = tid + k * kBeamThreads;
        xr[k] = c < C ? __ldg(xn + c) : 0.f;
      }
    }
    // ---- 1. log-softmax of the row (fp64)
    float mloc = -INFINITY;
    for (int c = tid; c < C; c += kBeamThreads) mloc = fmaxf(mloc, s_xs[c]);
    mloc = warp_max(mloc);
    if (lane == 0) s_redd[warp] = (double)mloc;
    __syncthreads();
    double m = s_redd[0];
#pragma unroll
    for (int w = 1; w < kBeamWarps; w++) m = fmax(m, s_redd[w]);
    double sloc = 0.0;
    for (int c = tid; c < C; c += kBeamThreads) sloc += exp((double)s_xs[c] - m);
    sloc = warp_sum(sloc);
    if (lane == 0) s_redd[16 + warp] = sloc;
    __syncthreads();
    double ssum = 0.0;
#pragma unroll
    for (int w = 0; w < kBeamWarps; w++) ssum += s_redd[16 + w];
    const double lse = log(ssum);
    for (int c = tid; c < C; c += kBeamThreads) s_lp[c] = ((double)s_xs[c] - m) - lse;
    const double lpmax = -lse;
    __syncthreads();
    // ---- 2. active prefixes keep themselves
    for (int e = tid; e < n; e += kBeamThreads) {
      double nl = ninf;
      const int last = g_last[ao + e];
      if (g_len[ao + e] > 0) {
        nl = g_pl[ao + e];
        const int ps = g_pslot[ao + e];
        if (ps >= 0) nl = lse2(nl, last == g_plast[ao + e] ? g_pb[ao + ps] : g_pt[ao + ps]);
        nl += s_lp[last];
      }
      const double nb = g_pt[ao + e] + s_lp[blank];
      s_ub[e] = nb;
      s_ul[e] = nl;
      s_ut[e] = lse2(nb, nl);
    }
    __syncthreads();
    // ---- is the beam full, what is its worst kept score, which prefixes can still place an extension
    if (warp == 0) {
      int cntv = 0;
      double mn = __longlong_as_double(0x7ff0000000000000ll);
      for (int e = lane; e < n; e += 32) {
        const double v = s_ut[e];
        if (v > ninf) {
          cntv++;
          mn = fmin(mn, v);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        cntv += __shfl_xor_sync(0xffffffffu, cntv, o);
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      }
      const double tau0 = cntv >= W ? mn : ninf;
      int nl = 0;
      for (int e0 = 0; e0 < n; e0 += 32) {
        const int e = e0 + lane;
        bool lv = false;
        if (e < n) {
          const double pt = g_pt[ao + e];
          lv = pt > ninf && (pt + lpmax > tau0);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, lv);
        if (lv) s_live[nl + __popc(bal & ((1u << lane) - 1u))] = e;
        nl += __popc(bal);
      }
      if (lane == 0) {
        s_redi[22] = nl;
        s_redi[20] = 0;
        s_redd[32] = tau0;
      }
    }
    __syncthreads();
    Ctx c;
    c.lp = s_lp; c.pb = g_pb + ao; c.pt = g_pt + ao; c.ut = s_ut; c.hash = g_hash + ao; c.last = g_last + ao;
    c.live = s_live; c.mask = s_mask; c.n = n; c.C = C; c.CW = CW; c.blank = blank; c.divM = divM;
    c.tau0 = s_redd[32];
    const int nitems = n + s_redi[22] * C;
    // ---- 3. the candidates' scores, how many they are, and their range
    u64 key[IPT > 0 ? IPT : 1];
    int cnt = 0;
    u64 kmin = ~0ull, kmax = 0;
    if (IPT > 0) {
#pragma unroll
      for (int r = 0; r < IPT; r++) {
        const int i = tid + r * kBeamThreads;
        double v;
        int b, l;
        key[r] = 0;
        if (i < nitems && eval_item(c, i, v, b, l)) key[r] = okey(v);
        if (key[r]) {
          cnt++;
          kmin = key[r] < kmin ? key[r] : kmin;
          kmax = key[r] > kmax ? key[r] : kmax;
        }
      }
    } else {
      for (int i = tid; i < nitems; i += kBeamThreads) {
        double v;
        int b, l;
        if (eval_item(c, i, v, b, l)) {
          const u64 k = okey(v);
          cnt++;
          kmin = k < kmin ? k : kmin;
          kmax = k > kmax ? k : kmax;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      const u64 a = __shfl_xor_sync(0xffffffffu, kmin, o), z = __shfl_xor_sync(0xffffffffu, kmax, o);
      kmin = a < kmin ? a : kmin;
      kmax = z > kmax ? z : kmax;
    }
    if (lane == 0) {
      s_redi[warp] = cnt;
      s_redu[warp] = kmin;
      s_redu[16 + warp] = kmax;
    }
    __syncthreads();
    cnt = 0;
    kmin = ~0ull;
    kmax = 0;
#pragma unroll
    for (int w = 0; w < kBeamWarps; w++) {
      cnt += s_redi[w];
      kmin = s_redu[w] < kmin ? s_redu[w] : kmin;
      kmax = s_redu[16 + w] > kmax ? s_redu[16 + w] : kmax;
    }
    // ---- 4. threshold of the best W: admit k >= F1, and among k == F1 (tie_mode) those with k2 >= F2
    u64 F1 = 0, F2 = 0;
    bool tie_mode = false;
    if (cnt > W) {
      int need = W;
      const u64 diff = kmin ^ kmax;
      if (diff == 0) {
        F1 = kmax;
        tie_mode = true;
      } else {
        u64 prefix = kmax;
        int top = 64 - __clzll((long long)diff);  // bits [top, 64) are common to every candidate
        for (;;) {
          const int width = min(11, top), shift = top - width;
          const unsigned dmask = (1u << width) - 1u;
          if (IPT > 0) {
#pragma unroll
            for (int r = 0; r < IPT; r++) {
              const u64 k = key[r];
              if (k && (top >= 64 || ((k ^ prefix) >> top) == 0)) atomicAdd(&s_hist[(int)((k >> shift) & dmask)], 1);
            }
          } else {
            for (int i = tid; i < nitems; i += kBeamThreads) {
              double v;
              int b, l;
              if (eval_item(c, i, v, b, l)) {
                const u64 k = okey(v);
                if (top >= 64 || ((k ^ prefix) >> top) == 0) atomicAdd(&s_hist[(int)((k >> shift) & dmask)], 1);
              }
            }
          }
          __syncthreads();
          find_bin(s_hist, 1 << width, need, s_redi + 32, s_redi + 16, tid);
          need -= s_redi[17];
          prefix = (prefix & ~((u64)dmask << shift)) | ((u64)s_redi[16] << shift);
          top = shift;
          if (s_redi[18] == need) {  // the whole bin is in
            F1 = prefix & ~(((u64)1 << shift) - 1);
            break;
          }
          if (top == 0) {  // more candidates with exactly this score than places left
            F1 = prefix;
            tie_mode = true;
            break;
          }
        }
      }
      if (tie_mode) {
        // order the candidates that share the threshold score by (active first, hash)
        u64 prefix = 0;
        int top = 64;
        for (;;) {
          const int width = min(11, top), shift = top - width;
          const unsigned dmask = (1u << width) - 1u;
#pragma unroll
          for (int r = 0; r < (IPT > 0 ? IPT : (nitems + kBeamThreads - 1) / kBeamThreads); r++) {
            const int i = tid + r * kBeamThreads;
            if (i >= nitems) break;
            double v;
            int b, l;
            u64 k1;
            if (IPT > 0) {
              k1 = key[r];  // r is a compile-time index after unrolling only when IPT > 0
              if (k1 != F1) continue;
              item_of(c, i, b, l);
            } else {
              if (!eval_item(c, i, v, b, l) || okey(v) != F1) continue;
            }
            const u64 k = item_k2(c, b, l);
            if (top >= 64 || ((k ^ prefix) >> top) == 0) atomicAdd(&s_hist[(int)((k >> shift) & dmask)], 1);
          }
          __syncthreads();
          find_bin(s_hist, 1 << width, need, s_redi + 32, s_redi + 16, tid);
          need -= s_redi[17];
          prefix = (prefix & ~((u64)dmask << shift)) | ((u64)s_redi[16] << shift);
          top = shift;
          if (s_redi[18] == need || top == 0) {
            F2 = prefix & ~(((u64)1 << shift) - 1);
            break;
          }
        }
      }
    }
    // ---- 5. survivors into the other buffer
#pragma unroll
    for (int r = 0; r < (IPT > 0 ? IPT : (nitems + kBeamThreads - 1) / kBeamThreads); r++) {
      const int i = tid + r * kBeamThreads;
      if (i >= nitems) break;
      double v;
      int b, l;
      u64 k;
      if (IPT > 0) {
        k = key[r];
        if (k == 0 || k < F1) continue;
        item_of(c, i, b, l);
        const long long bits = (long long)k;  // invert okey
        v = __longlong_as_double(bits < 0 ? (bits ^ (long long)0x8000000000000000ull) : ~bits);
      } else {
        if (!eval_item(c, i, v, b, l)) continue;
        k = okey(v);
        if (k < F1) continue;
      }
      if (tie_mode && k == F1 && item_k2(c, b, l) < F2) continue;
      const int slot = atomicAdd(&s_redi[20], 1);
      if (slot >= W) continue;  // only reachable through a hash collision among exact ties
      const int d = no + slot, a = ao + b;
      if (l < 0) {
        g_pb[d] = s_ub[b];
        g_pl[d] = s_ul[b];
        g_pt[d] = v;
        g_hash[d] = g_hash[a];
        g_phash[d] = g_phash[a];
        g_node[d] = g_node[a];
        g_len[d] = g_len[a];
        g_last[d] = g_last[a];
        g_plast[d] = g_plast[a];
      } else {
        const int id = atomicAdd(&s_redi[21], 1);
        nodes[id] = make_int2(g_node[a], l);
        g_pb[d] = ninf;
        g_pl[d] = v;
        g_pt[d] = v;
        g_hash[d] = child_hash(g_hash[a], l);
        g_phash[d] = g_hash[a];
        g_node[d] = id;
        g_len[d] = g_len[a] + 1;
        g_last[d] = l;
        g_plast[d] = g_last[a];
      }
    }
    __syncthreads();
    const int n_new = min(s_redi[20], W);
    // ---- 6. parents' slots, then the per-parent sets of active extensions
    for (int i = tid; i < n_new * CW; i += kBeamThreads) s_mask[i] = 0;
    for (int e0 = 0; e0 < n_new; e0 += kBeamThreads / 4) {  // warp-uniform trip count: the shuffles need every lane
      const int e = e0 + (tid >> 2);
      int ps = -1, ln = -1;
      if (e < n_new) {
        const u64 ph = g_phash[no + e];
        ln = g_len[no + e] - 1;
        for (int j = tid & 3; j < n_new; j += 4)
          if (g_hash[no + j] == ph && g_len[no + j] == ln) ps = j;
      }
      ps = max(ps, __shfl_xor_sync(0xffffffffu, ps, 1));
      ps = max(ps, __shfl_xor_sync(0xffffffffu, ps, 2));
      if (e < n_new && (tid & 3) == 0) g_pslot[no + e] = ln >= 0 ? ps : -1;
    }
    if (pre) {
#pragma unroll
      for (int k = 0; k < kPrefetch; k++) {
        const int cc = tid + k * kBeamThreads;
        if (cc < C) s_xs[cc] = xr[k];
      }
    } else if (t + 1 < Tb) {
      const float* xn = xrow + (size_t)(t + 1) * st_t;
      for (int cc = tid; cc < C; cc += kBeamThreads) s_xs[cc] = __ldg(xn + cc);
    }
    __syncthreads();
    for (int e = tid; e < n_new; e += kBeamThreads) {
      const int ps = g_pslot[no + e];
      if (ps >= 0) atomicOr(&s_mask[ps * CW + (g_last[no + e] >> 5)], 1u << (g_last[no + e] & 31));
    }
    __syncthreads();
    n = n_new;
    cur ^= 1;
  }

  // ---- read the best P prefixes back through the trie
  const int ao = cur * W;
  for (int e = tid; e < n; e += kBeamThreads) {
    const double v = g_pt[ao + e];
    const u64 h = g_hash[ao + e];
    int r = 0;
    for (int j = 0; j < n; j++) {
      const double vj = g_pt[ao + j];
      const u64 hj = g_hash[ao + j];
      r += (vj > v) || (vj == v && (hj < h || (hj == h && j < e)));
    }
    if (r < P) {
      const int len = g_len[ao + e];
      int64_t* row = hyp + ((size_t)b_utt * P + r) * T;
      int id = g_node[ao + e];
      for (int pos = len - 1; pos >= 0; pos--) {
        const int2 nd = nodes[id];
        row[pos] = nd.y;
        id = nd.x;
      }
      s_live[r] = len;  // s_live is free now (P <= W)
      log_prob[(size_t)b_utt * P + r] = (float)v;
    }
  }
  for (int r = n + tid; r < P; r += kBeamThreads) {
    s_live[r] = 0;
    log_prob[(size_t)b_utt * P + r] = -INFINITY;
  }
  __syncthreads();
  for (int r = warp; r < P; r += kBeamWarps) {
    const int len = s_live[r];
    int64_t* row = hyp + ((size_t)b_utt * P + r) * T;
    int out = len;
    if (merge_repeated) {
      out = 0;
      int64_t carry = -1;
      for (int base = 0; base < len; base += 32) {
        const int i = base + lane;
        const int64_t x = i < len ? row[i] : -1;
        int64_t xp = __shfl_up_sync(0xffffffffu, x, 1);
        if (lane == 0) xp = carry;
        carry = __shfl_sync(0xffffffffu, x, 31);
        const bool keep = i < len && x != xp;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) row[out + __popc(bal & ((1u << lane) - 1u))] = x;
        out += __popc(bal);
      }
    }
    if (lane == 0) hyp_len[(size_t)b_utt * P + r] = out;
  }
}

}  // namespace

int ctc_beam_workspace_bytes(int T, int B, int C, int W, size_t* out) {
  (void)C;
  *out = sizeof(int2) * (size_t)B * ((size_t)T * W + 1) + 256;
  return NASR_OK;
}

int ctc_beam_search(const float* logits, int T, int B, int C, long long st_t, long long st_b,
                    const int32_t* seq_len, int blank, int W, int P, int merge_repeated, int64_t* hyp,
                    int32_t* hyp_len, float* log_prob, void* workspace, size_t workspace_bytes,
                    cudaStream_t stream) {
  NASR_CHECK_ARG(T >= 0 && B >= 0 && C >= 1, "ctc_beam_search: bad shape T=%d B=%d C=%d", T, B, C);
  NASR_CHECK_ARG(blank >= 0 && blank < C, "ctc_beam_search: blank=%d outside [0,%d)", blank, C);
  NASR_CHECK_ARG(W >= 1 && P >= 1 && P <= W, "ctc_beam_search: need 1 <= top_paths <= beam_width (got %d, %d)", P, W);
  if (B == 0) return NASR_OK;
  NASR_CHECK_ARG(logits && seq_len && hyp && hyp_len && log_prob && workspace, "ctc_beam_search: NULL argument");
  size_t need = 0;
  ctc_beam_workspace_bytes(T, B, C, W, &need);
  if (workspace_bytes < need) {
    set_error("ctc_beam_search: workspace of %zu bytes, %zu needed", workspace_bytes, need);
    return NASR_ERR_WORKSPACE_TOO_SMALL;
  }
  const size_t smem = beam_smem_bytes(W, C);
  if (W > 1024 || C > 8192 || (size_t)W * C >= ((size_t)1 << 20) || smem > 200 * 1024) {
    set_error("ctc_beam_search: beam_width=%d with C=%d is not supported (needs %zu bytes of shared memory)", W, C,
              smem);
    return NASR_ERR_UNSUPPORTED;
  }
  int2* nodes = reinterpret_cast<int2*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  const u64 divM = (((u64)1 << 40) / (u64)C) + 1;
  const size_t cand = (size_t)W * ((size_t)C + 1);  // most candidates a frame can have
#define NASR_BEAM_LAUNCH(IPT)                                                                                     \
  do {                                                                                                            \
    NASR_CUDA(cudaFuncSetAttribute(ctc_beam_kernel<IPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
    ctc_beam_kernel<IPT><<<B, kBeamThreads, smem, stream>>>(logits, T, B, C, st_t, st_b, seq_len, blank, W, P,    \
                                                            merge_repeated, hyp, hyp_len, log_prob, nodes, divM); \
  } while (0)
  if (cand <= 8 * (size_t)kBeamThreads)
    NASR_BEAM_LAUNCH(8);
  else if (cand <= 16 * (size_t)kBeamThreads)
    NASR_BEAM_LAUNCH(16);
  else
    NASR_BEAM_LAUNCH(0);
#undef NASR_BEAM_LAUNCH
  count_launch();
  NASR_CUDA(cudaGetLastError());
  return NASR_OK;
}

}  // namespace nasr
