// CTC beam search decoder for sm_100a.
//
// Replaces tf.nn.ctc_beam_search_decoder(logits, seq_len) — beam_width 100, top_paths 1, merge_repeated True —
// the decoder the reference's create_model actually runs in every train step (networks/tfnetwork.py:62,64);
// semantics per oracle/beam_oracle.py (TF 1.x core/util/ctc/ctc_beam_search.h restated).
//
// One CTA per utterance walks the frames: 16 search warps and one producer warp that has the log-softmax (fp64) of
// the next frame's row ready when the search of the current one ends.  Per frame, phases separated by a named
// barrier of the search warps:
//   1. update: one thread per active prefix: blank' = total + lp[blank], label' = LSE(label + lp[last], parent term
//      + lp[last]) if its parent prefix is active (slot kept per prefix), total' as one three-way log-sum-exp;
//   2. lists: a full beam admits only scores above its worst kept total tau0 (TF's candidate test), so prefixes with
//      total + max lp <= tau0 and labels with lp + max total <= tau0 are dropped and the rest compacted by ballot;
//      wide vocabularies first tighten tau0 to the W-th best of {kept prefixes, extensions by the two best labels};
//   3. keys: candidates = kept prefixes + live prefixes x live labels that are not already active prefixes (bit set
//      per prefix): lp[l] + (l == last(b) ? blank(b) : total(b)).  Each search thread evaluates up to 8 and keeps the
//      order-preserving 64-bit image of the fp64 score in registers (more than 4096: recomputed in every pass);
//   4. select: MSB-first radix select of the W-th best key (11-bit digits, shared-memory histogram, from the first
//      bit in which the known bounds differ); exact ties go on through (kept-before-new, prefix hash);
//   5. admit: keys at or above the threshold -> a list (one shared atomic per warp);
//   6. build: one thread per admitted candidate writes its entry of the other beam buffer; a new prefix gets node
//      1 + t*W + slot = (parent node, label) of a per-utterance trie in the workspace and enters a small hash table;
//   7. parents: each entry's parent slot in the new beam (old->new slot map, or the hash table when the parent has
//      just re-entered the beam) and the per-parent bit sets of active extensions.
// Prefix identity is a 64-bit hash chain (root constant, child = mix(parent + K*(label+1))) plus the prefix length:
// two different prefixes of equal length colliding inside one beam is a 2^-64-per-pair event and would merge them.
// After the last frame the top_paths best prefixes are read back through the trie, repeats collapsed when
// merge_repeated (TF merges in the OUTPUT), and their log probabilities returned.
// Scores are fp64 (TF: fp32): the label sequences are what the reference consumes, and fp64 keeps the kernel
// and the oracle on the same side of every comparison that is not an exact tie.
#include <math.h>

#include "nasr_common.cuh"

// The phase ticks behind nasr_debug_profile are compiled in only with -DNASR_TUNING=1 (see ctc_fast.cu).
#ifndef NASR_TUNING
#define NASR_TUNING 0
#endif

namespace nasr {
extern long long* g_debug_prof;  // ctc_fast.cu (nasr_debug_profile)
namespace {

typedef unsigned long long u64;

#ifndef NASR_BEAM_SEARCH_THREADS
#define NASR_BEAM_SEARCH_THREADS 512
#endif
constexpr int kSearchThreads = NASR_BEAM_SEARCH_THREADS;  // the warps that search (a power of two, 128..512)
constexpr int kSearchWarps = kSearchThreads / 32;
constexpr int kOneProducerMaxC = 1536;              // wider rows get four producer warps
constexpr int kIPT = 4096 / kSearchThreads;         // candidate keys a search thread keeps in registers
constexpr int kBins = 2048;                         // 11-bit digits
constexpr int kBitSetMaxC = 4095;                   // widest vocabulary whose active extensions are kept as bit sets
#ifndef NASR_BEAM_STAGE2_MINC
#define NASR_BEAM_STAGE2_MINC 64
#endif
constexpr int kStage2MinC = NASR_BEAM_STAGE2_MINC;  // vocabularies wider than this get the second-stage bound
constexpr int kBinsPerThread = kBins / kSearchThreads;
constexpr u64 kRootHash = 0x243f6a8885a308d3ull;

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
__device__ __forceinline__ double okey_inv(u64 k) {
  const long long b = (long long)k;
  return __longlong_as_double(b < 0 ? (b ^ (long long)0x8000000000000000ull) : ~b);
}

__device__ __forceinline__ double lse2(double a, double b) {
  const double ninf = neg_inf();
  if (a == ninf) return b;
  if (b == ninf) return a;
  const double m = fmax(a, b), n = fmin(a, b);
  return m + log1p(exp(n - m));
}

// barrier of the search warps only (the producer warp meets them once per frame at barrier 0)
__device__ __forceinline__ void bar_search() { asm volatile("bar.sync 1, %0;" ::"n"(kSearchThreads) : "memory"); }

__host__ __device__ inline size_t al16(size_t n) { return (n + 15) & ~(size_t)15; }
__host__ __device__ inline int tab_size(int W) {
  int s = 64;
  while (s < 2 * W) s <<= 1;
  return s;
}

// Shared-memory layout, computed once on the host and handed to the kernel as a __grid_constant__ parameter:
// every array base is then a constant-bank operand instead of arithmetic on W and C redone inside the frame loop.
struct BeamLayout {
  int x2, lp2, pb, pl, pt, hash, phash, node, len, last, plast, pslot, ub, ul, ut, liveP, newslot, adm_i, liveL, adm_k,
      tab_key, tab_slot, mtab, hist, redd, redu, redi, c1, c2;
  int has_bits, CW;  // bit sets of active extensions ([W][CW] words) while they fit (C <= kBitSetMaxC), else a hash set
  int total, TS, tshift;  // TS: slots of the two hash tables (a power of two >= 2W); tshift = 32 - log2(TS)
};

inline BeamLayout beam_layout(int W, int C) {
  BeamLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t r = o;
    o += al16(bytes);
    return (int)r;
  };
  L.TS = tab_size(W);
  L.tshift = 32;
  for (int t = L.TS; t > 1; t >>= 1) L.tshift--;
  L.has_bits = C <= kBitSetMaxC;
  L.x2 = take(L.has_bits ? 0 : sizeof(float) * 2 * C);        // BIG: raw rows, lp on demand
  L.lp2 = take(L.has_bits ? sizeof(double) * 2 * C : 0);      // else: fp64 log-probability rows
  L.pb = take(sizeof(double) * 2 * W);
  L.pl = take(sizeof(double) * 2 * W);
  L.pt = take(sizeof(double) * 2 * W);
  L.hash = take(sizeof(u64) * 2 * W);
  L.phash = take(sizeof(u64) * 2 * W);
  L.node = take(sizeof(int) * 2 * W);
  L.len = take(sizeof(int) * 2 * W);
  L.last = take(sizeof(int) * 2 * W);
  L.plast = take(sizeof(int) * 2 * W);
  L.pslot = take(sizeof(int) * 2 * W);
  L.ub = take(sizeof(double) * W);
  L.ul = take(sizeof(double) * W);
  L.ut = take(sizeof(double) * W);
  L.liveP = take(sizeof(int) * W);
  L.newslot = take(sizeof(int) * W);
  L.adm_i = take(sizeof(int) * W);
  L.liveL = take(sizeof(int) * C);
  L.adm_k = take(sizeof(u64) * W);
  L.tab_key = take(sizeof(u64) * L.TS);
  L.tab_slot = take(sizeof(int) * L.TS);
  L.CW = (C + 31) / 32;
  L.mtab = take(sizeof(uint32_t) * (L.has_bits ? (size_t)W * L.CW : (size_t)L.TS));
  L.hist = take(sizeof(int) * kBins);
  L.redd = take(sizeof(double) * 128);  // indices used: 0-59 by the search warps, 96-119 by the producer warps
  L.redu = take(sizeof(u64) * 64);
  L.redi = take(sizeof(int) * 64);
  L.c1 = take(sizeof(double) * W);
  L.c2 = take(sizeof(double) * W);
  L.total = (int)o;
  return L;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  return v;
}

// Extensions that are already active prefixes: a hash set of (parent slot << 16 | label), at most one entry per
// active prefix, open addressing in a table of TS >= 2W slots (0xffffffff = empty).
constexpr uint32_t kEmpty = 0xffffffffu;
__device__ __forceinline__ uint32_t mtab_key(int slot, int label) { return ((uint32_t)slot << 16) | (uint32_t)label; }
__device__ __forceinline__ unsigned mtab_home(uint32_t key, int tshift) { return (key * 2654435761u) >> tshift; }
__device__ __forceinline__ bool mtab_has(const uint32_t* tab, int TS, int tshift, int slot, int label) {
  const uint32_t key = mtab_key(slot, label);
  unsigned idx = mtab_home(key, tshift);
  for (;;) {
    const uint32_t k = tab[idx];
    if (k == key) return true;
    if (k == kEmpty) return false;
    idx = (idx + 1) & (unsigned)(TS - 1);
  }
}
__device__ __forceinline__ void mtab_insert(uint32_t* tab, int TS, int tshift, int slot, int label) {
  const uint32_t key = mtab_key(slot, label);
  unsigned idx = mtab_home(key, tshift);
  for (;;) {
    const uint32_t old = atomicCAS(&tab[idx], kEmpty, key);
    if (old == kEmpty || old == key) return;
    idx = (idx + 1) & (unsigned)(TS - 1);
  }
}

template <bool BIG>
__device__ __forceinline__ bool active_ext(const uint32_t* tab, int CW, int TS, int tshift, int slot, int label) {
  return BIG ? mtab_has(tab, TS, tshift, slot, label) : (tab[slot * CW + (label >> 5)] >> (label & 31)) & 1u;
}

struct Ctx {
  // frame constants every candidate evaluation needs
  const float* x;        // BIG: the frame's logits row, lp[l] = ((double)x[l] - m) - lse
  const double* lp;      // else: the frame's log-probability row
  double m, lse;
  const double *pb, *pt, *ut;
  const u64* hash;
  const int *last, *liveP, *liveL;
  const uint32_t* mtab;  // extensions that are already active prefixes: bit sets [W][CW] (CW > 0) or a hash set
  int n, nL, TS, tshift, CW;
  bool div32;            // q = j / nL by one 32-bit multiply (exact while W * nL^2 < 2^32)
  unsigned divM;  // floor((2^32-1) / nL) + 1: j / nL == umulhi(j, divM) for j < 2^20 (nL >= 2)
  u64 kkeep, kext;  // smallest key a kept prefix / an extension must have to be a candidate
};

// Candidate i of the frame: i < n is active prefix i itself (l = -1); otherwise the extension of live prefix
// liveP[q] by live label liveL[r], i - n = q*nL + r.
template <bool BIG>
__device__ __forceinline__ void item_of(const Ctx& c, int i, int& b, int& l) {
  if (i < c.n) {
    b = i;
    l = -1;
    return;
  }
  const unsigned j = (unsigned)(i - c.n);
  // (beam_width * C < 2^20 and C <= kBitSetMaxC keep the multiply exact; wider vocabularies check per frame)
  const unsigned q = c.nL == 1 ? j : ((!BIG || c.div32) ? __umulhi(j, c.divM) : j / (unsigned)c.nL);
  b = c.liveP[q];
  l = c.liveL[j - q * (unsigned)c.nL];
}

// Its key (order-preserving image of its score), 0 if it is not offered to the beam.
template <bool BIG>
__device__ __forceinline__ u64 eval_key(const Ctx& c, int i) {
  int b, l;
  item_of<BIG>(c, i, b, l);
  double v;
  bool ok;
  u64 kmin;
  if (l < 0) {
    v = c.ut[b];
    ok = true;
    kmin = c.kkeep;
  } else {
    const bool masked = BIG ? mtab_has(c.mtab, c.TS, c.tshift, b, l) : (c.mtab[b * c.CW + (l >> 5)] >> (l & 31)) & 1u;
    v = (BIG ? ((double)c.x[l] - c.m) - c.lse : c.lp[l]) + (l == c.last[b] ? c.pb[b] : c.pt[b]);
    ok = !masked;
    kmin = c.kext;
  }
  const u64 k = okey(v);
  return ok && k >= kmin ? k : 0ull;
}

template <bool BIG>
__device__ __forceinline__ u64 item_k2(const Ctx& c, int i) {
  // active prefixes win ties against extensions (TF admits an extension only if it is strictly better than the
  // worst kept entry); then the smaller hash wins
  int b, l;
  item_of<BIG>(c, i, b, l);
  if (l < 0) return 0x8000000000000000ull | ((~c.hash[b]) >> 1);
  return (~child_hash(c.hash[b], l)) >> 1;
}

// Find, from the top bin down, the bin in which the running count reaches `need`.  hist is left zeroed.
// res[0] = bin, res[1] = candidates in the bins above it, res[2] = candidates in it.  Search warps only.
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
  bar_search();
  {  // + the totals of the warps before this one
    const int ws = warp_incl_scan(lane < kSearchWarps ? wsum[lane] : 0, lane);
    const int before = __shfl_sync(0xffffffffu, ws, (warp + 31) & 31);
    if (warp) incl += before;
  }
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
  bar_search();
}

// STAGE2: second-stage bound on the frame's threshold (phase 2); pays for wide vocabularies, where it shortens the
// live-label list by one to two orders of magnitude, and for narrow ones while a CTA has an SM to itself.
// BIG: vocabularies wider than kBitSetMaxC.  Their fp64 log-probability rows (2 x C doubles) and bit sets of active
// extensions (W x C bits) would not fit shared memory: the rows stay float and lp is computed where it is used, the
// active extensions become a hash set of (prefix slot, label) pairs.  Narrower ones keep both (6-13 % faster).
// NP: producer warps (1, or 4 for rows of more than kOneProducerMaxC classes, whose fp64 softmax one warp cannot
// finish within a frame's search; then one CTA per SM).
// MINB: CTAs per SM the kernel is compiled for: 2 (56 registers, a few spills) when the batch needs two per SM,
// 1 (96 registers, none) otherwise -- 10 % faster at B=64 in an A/B on one box.
// CO ("cached only"): beam_width * (C + 1) candidates always fit the register-cached path, so the passes that
// recompute candidates are compiled out -- the kernel spends 2.7 cycles per issued instruction waiting for
// instruction fetch at two CTAs per SM (ncu), and a third of its loop body is those passes.
template <bool STAGE2, bool BIG, int NP, int MINB, bool CO>
__global__ void __launch_bounds__(kSearchThreads + 32 * NP, MINB)
ctc_beam_kernel(const float* __restrict__ logits, int T, int B, int C, long long st_t, long long st_b,
                const int32_t* __restrict__ seq_len, int blank, int W, int P, int merge_repeated,
                int64_t* hyp, int32_t* __restrict__ hyp_len, float* __restrict__ log_prob,
                int2* nodes_all, long long* prof, const __grid_constant__ BeamLayout L) {
  extern __shared__ __align__(16) char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b_utt = blockIdx.x;
  const int TS = L.TS, tshift = L.tshift;
  const int CW = BIG ? 0 : L.CW;  // 0: the active extensions are a hash set
  const double ninf = neg_inf();

#define BEAM_ARR(T, name, field) T* const name = reinterpret_cast<T*>(smem_raw + L.field)
  BEAM_ARR(float, s_x2, x2);          // BIG: logits rows of frame t (t&1) and t+1; their max / log-sum in s_redd
  BEAM_ARR(double, s_lp2, lp2);       // else: their log-softmax in fp64
  // the two beam buffers are the halves [0,W) and [W,2W) of each array
  BEAM_ARR(double, g_pb, pb);         // log P(prefix, ends in blank)
  BEAM_ARR(double, g_pl, pl);         //                ends in its last label
  BEAM_ARR(double, g_pt, pt);         //                either
  BEAM_ARR(u64, g_hash, hash);        // prefix identity
  BEAM_ARR(u64, g_phash, phash);      // parent prefix identity
  BEAM_ARR(int, g_node, node);
  BEAM_ARR(int, g_len, len);
  BEAM_ARR(int, g_last, last);
  BEAM_ARR(int, g_plast, plast);
  BEAM_ARR(int, g_pslot, pslot);
  BEAM_ARR(double, s_ub, ub);         // this frame's update of the active prefixes
  BEAM_ARR(double, s_ul, ul);
  BEAM_ARR(double, s_ut, ut);
  BEAM_ARR(int, s_liveP, liveP);      // prefixes whose extensions can still enter the beam
  BEAM_ARR(int, s_newslot, newslot);  // active prefix -> its slot in the next beam (-1: dropped)
  BEAM_ARR(int, s_adm_i, adm_i);      // admitted candidates (item index, key)
  BEAM_ARR(int, s_liveL, liveL);      // labels whose extension of the best prefix could enter
  BEAM_ARR(u64, s_adm_k, adm_k);
  BEAM_ARR(u64, s_tab_key, tab_key);  // hash -> slot of the prefixes that enter the beam this frame
  BEAM_ARR(int, s_tab_slot, tab_slot);
  BEAM_ARR(uint32_t, s_mtab, mtab);   // (prefix slot, label) pairs whose extension is already an active prefix
  BEAM_ARR(int, s_hist, hist);
  BEAM_ARR(double, s_redd, redd);
  BEAM_ARR(u64, s_redu, redu);
  BEAM_ARR(int, s_redi, redi);
  BEAM_ARR(double, s_c1, c1);         // score of each prefix's extension by the frame's best / second label
  BEAM_ARR(double, s_c2, c2);
#undef BEAM_ARR
  // s_redi: [0..15] warp partials, [16..18] find_bin result, [20] admitted counter,
  //         [22] live prefixes, [23] live labels, [32..47] find_bin warp sums, [56..59] the two best labels of rows t&1
  // s_redd: [0..15] min of the updated totals, [16..31] max of the old totals, [40..41] max of lp rows t&1,
  //         [44..59] max of the updated totals, [36..39] max / log-sum of rows t&1, [96..119] the producer warps' partial maxima,
  //         sums and best labels
  // s_redu: [32..39] profile accumulators

  int Tb = seq_len[b_utt];
  Tb = max(0, min(T, Tb));
  const float* xrow = logits + (size_t)b_utt * st_b;
  int2* nodes = nodes_all + (size_t)b_utt * ((size_t)T * W + 1);

  constexpr int kBeamThreads = kSearchThreads + 32 * NP;
  for (int i = tid; i < kBins; i += kBeamThreads) s_hist[i] = 0;
  for (int i = tid; i < (BIG ? TS : CW); i += kBeamThreads) s_mtab[i] = BIG ? kEmpty : 0u;
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
    nodes[0] = make_int2(-1, -1);
    for (int k = 32; k < 40; k++) s_redu[k] = 0;
  }

  if (warp >= kSearchWarps) {
    // ---- producer warps: log-softmax (fp64) of row t+1 while the others search frame t.  Warp pw of NP takes the
    //      classes pw*32 + lane, pw*32 + lane + 32*NP, ...; with NP > 1 the maxima, sums and best labels of the
    //      warps meet in shared memory at a named barrier of the producers.
    const int pw = warp - kSearchWarps, pstep = 32 * NP;
    auto bar_producers = []() {
      if (NP > 1) asm volatile("bar.sync 2, %0;" ::"n"(32 * NP) : "memory");
    };
    for (int t = 0; t <= Tb; t++) {  // Tb + 1 barriers: before frame 0 and after every frame
      if (t == Tb) {
        __syncthreads();
        break;
      }
      const float* x = xrow + (size_t)t * st_t;
      if (t + 1 < Tb) {  // pull the row after this one towards L2
        const char* nx = reinterpret_cast<const char*>(x + st_t);
        for (int o = (pw * 32 + lane) * 128; o < C * 4; o += pstep * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + o));
      }
      float mx = -INFINITY;
      for (int c = pw * 32 + lane; c < C; c += pstep) mx = fmaxf(mx, __ldg(x + c));
      mx = warp_max(mx);
      if (NP > 1) {
        if (lane == 0) s_redd[96 + pw] = (double)mx;
        bar_producers();
#pragma unroll
        for (int w = 0; w < NP; w++) mx = fmaxf(mx, (float)s_redd[96 + w]);
      }
      const double m = (double)mx;
      double sum = 0.0;
      for (int c = pw * 32 + lane; c < C; c += pstep) sum += exp((double)__ldg(x + c) - m);
      sum = warp_sum(sum);
      // the two best labels (blank aside; ties: the smaller index), for the second-stage bound of phase 2
      float v1 = -INFINITY, v2 = -INFINITY;
      int i1 = -1, i2 = -1;
      for (int c = pw * 32 + lane; STAGE2 && c < C; c += pstep) {
        const float v = __ldg(x + c);
        if (c == blank) continue;
        if (v > v1) {
          v2 = v1; i2 = i1; v1 = v; i1 = c;
        } else if (v > v2) {
          v2 = v; i2 = c;
        }
      }
      auto merge2 = [&](float w1, int j1, float w2, int j2) {  // fold another (best, second) pair into ours
        const bool first = w1 > v1 || (w1 == v1 && (unsigned)j1 < (unsigned)i1);  // the other's best is the best
        const float a1 = first ? w1 : v1, b1 = first ? v1 : w1;  // b1: the loser of the two bests
        const int ai = first ? j1 : i1, bi = first ? i1 : j1;
        const float c2v = first ? w2 : v2;                       // the winner's own second
        const int c2i = first ? j2 : i2;
        const bool bsecond = b1 > c2v || (b1 == c2v && (unsigned)bi < (unsigned)c2i);
        v1 = a1; i1 = ai;
        v2 = bsecond ? b1 : c2v; i2 = bsecond ? bi : c2i;
      };
#pragma unroll
      for (int o = 16; STAGE2 && o > 0; o >>= 1) {
        const float w1 = __shfl_xor_sync(0xffffffffu, v1, o), w2 = __shfl_xor_sync(0xffffffffu, v2, o);
        const int j1 = __shfl_xor_sync(0xffffffffu, i1, o), j2 = __shfl_xor_sync(0xffffffffu, i2, o);
        merge2(w1, j1, w2, j2);
      }
      if (NP > 1) {
        if (lane == 0) {
          s_redd[100 + pw] = sum;
          s_redd[104 + 4 * pw] = (double)v1;
          s_redd[105 + 4 * pw] = (double)i1;
          s_redd[106 + 4 * pw] = (double)v2;
          s_redd[107 + 4 * pw] = (double)i2;
        }
        bar_producers();
        sum = 0.0;
#pragma unroll
        for (int w = 0; w < NP; w++) sum += s_redd[100 + w];  // the same order in every warp
        if (STAGE2) {
          v1 = v2 = -INFINITY;
          i1 = i2 = -1;
#pragma unroll
          for (int w = 0; w < NP; w++)
            merge2((float)s_redd[104 + 4 * w], (int)s_redd[105 + 4 * w], (float)s_redd[106 + 4 * w],
                   (int)s_redd[107 + 4 * w]);
        }
      }
      const double lse = log(sum);
      if (BIG) {
        float* xs = s_x2 + (t & 1) * C;
        for (int c = pw * 32 + lane; c < C; c += pstep) xs[c] = __ldg(x + c);
      } else {
        double* lp = s_lp2 + (t & 1) * C;
        for (int c = pw * 32 + lane; c < C; c += pstep) lp[c] = ((double)__ldg(x + c) - m) - lse;
      }
      if (pw == 0 && lane == 0) {
        s_redd[40 + (t & 1)] = -lse;  // the row's best log-probability
        s_redd[36 + (t & 1)] = m;
        s_redd[38 + (t & 1)] = lse;
        s_redi[56 + 2 * (t & 1)] = i1;
        s_redi[57 + 2 * (t & 1)] = i2;
      }
      __syncthreads();  // row t is ready (and the search of frame t-1 is over)
    }
    return;
  }

  // tuning hook (nasr_debug_profile): thread 0 of CTA 0 accumulates the cycles of each phase of the frame loop
  long long tk = 0;  // the accumulators live in s_redu[32..39]
  const bool profiling = NASR_TUNING && prof != nullptr && blockIdx.x == 0 && tid == 0;
#define BEAM_TICK(k)                    \
  if (profiling) {                      \
    const long long now_ = clock64();   \
    s_redu[32 + k] += (u64)(now_ - tk); \
    tk = now_;                          \
  }
  __syncthreads();  // row 0 ready, root entry written
  int n = 1, cur = 0;
  for (int t = 0; t < Tb; t++) {
    if (profiling) tk = clock64();
    const int ao = cur * W, no = (cur ^ 1) * W;  // offsets of the active and of the next beam buffer
    const float* xr = s_x2 + (t & 1) * C;
    const double rm = s_redd[36 + (t & 1)], rlse = s_redd[38 + (t & 1)];
    const double* lpd = s_lp2 + (t & 1) * C;
#define BEAM_LP(i) (BIG ? ((double)xr[i] - rm) - rlse : lpd[i])
    const double lpmax = s_redd[40 + (t & 1)];
    const int top1 = s_redi[56 + 2 * (t & 1)], top2 = s_redi[57 + 2 * (t & 1)];
    // ---- 1. active prefixes keep themselves; worst kept score and best old score by warp
    {
      double ut = ninf, pt_old = ninf;
      if (tid < n) {
        const int e = tid;
        const int last = g_last[ao + e];
        pt_old = g_pt[ao + e];
        double b1 = ninf, b2 = ninf;  // stay on the last label; arrive from the parent prefix
        if (g_len[ao + e] > 0) {
          const double lpl = BEAM_LP(last);
          b1 = g_pl[ao + e] + lpl;
          const int ps = g_pslot[ao + e];
          if (ps >= 0) b2 = (last == g_plast[ao + e] ? g_pb[ao + ps] : g_pt[ao + ps]) + lpl;
        }
        const double nb = pt_old + BEAM_LP(blank);
        const double nl = lse2(b1, b2);
        // total' as ONE three-way log-sum-exp, independent of nl's chain (oracle/beam_oracle.py does the same)
        const double mm = fmax(nb, fmax(b1, b2));
        ut = mm == ninf ? ninf : mm + log(exp(nb - mm) + exp(b1 - mm) + exp(b2 - mm));
        s_ub[e] = nb;
        s_ul[e] = nl;
        s_ut[e] = ut;
        s_newslot[e] = -1;
        // this prefix's extensions by the frame's two best labels: true candidates of the frame (phase 2)
        double c1 = ninf, c2 = ninf;
        if (STAGE2 && top1 >= 0 && !active_ext<BIG>(s_mtab, CW, TS, tshift, e, top1))
          c1 = BEAM_LP(top1) + (top1 == last ? g_pb[ao + e] : pt_old);
        if (STAGE2 && top2 >= 0 && !active_ext<BIG>(s_mtab, CW, TS, tshift, e, top2))
          c2 = BEAM_LP(top2) + (top2 == last ? g_pb[ao + e] : pt_old);
        if (STAGE2) {
          s_c1[e] = c1;
          s_c2[e] = c2;
        }
      }
      if (warp * 32 < n) {
        const unsigned okb = __ballot_sync(0xffffffffu, ut > ninf);
        double mn = ut > ninf ? ut : __longlong_as_double(0x7ff0000000000000ll), mxo = pt_old, mxu = ut;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
          mxo = fmax(mxo, __shfl_xor_sync(0xffffffffu, mxo, o));
          mxu = fmax(mxu, __shfl_xor_sync(0xffffffffu, mxu, o));
        }
        if (lane == 0) {
          s_redi[warp] = __popc(okb);
          s_redd[warp] = mn;
          s_redd[16 + warp] = mxo;
          s_redd[44 + warp] = mxu;
        }
      }
      for (int i = tid; i < TS; i += kSearchThreads) s_tab_key[i] = 0;
      if (tid == 0) {
        s_redi[20] = 0;
        s_redi[22] = 0;
        s_redi[23] = 0;
      }
    }
    bar_search();
    BEAM_TICK(0);
    // ---- 2. is the beam full (then only scores above its worst kept one, tau0, matter); wide vocabularies: a
    //         second-stage bound; live prefixes and labels
    double tau0, vtop;  // every candidate's score lies in (tau0, vtop] (kept prefixes: [tau0, vtop])
    u64 kfloor = 0;     // second-stage bound (key domain), 0 = none
    if (STAGE2) {
      // every thread combines the warps' partials itself
      double ptmax = ninf;
      {
        int cntv = 0;
        double mn = __longlong_as_double(0x7ff0000000000000ll), utmax = ninf;
        for (int w = 0; w * 32 < n; w++) {
          cntv += s_redi[w];
          mn = fmin(mn, s_redd[w]);
          ptmax = fmax(ptmax, s_redd[16 + w]);
          utmax = fmax(utmax, s_redd[44 + w]);
        }
        tau0 = cntv >= W ? mn : ninf;
        vtop = fmax(utmax, ptmax + lpmax);
      }
      // Second stage.  S = the W updated kept prefixes + every prefix's extensions by the frame's two best labels:
      // all true candidates, so the W-th best of S is a lower bound of the frame's threshold, and usually a tight one
      // (measured: live labels 441 -> 5 at C = 1024, 16 -> 2 at C = 38 on N(0,1)*3 logits; 425 -> 20 and 18 -> 6 on
      // planted alignments).  One histogram pass over S gives the bin of its W-th best; the bin's lower edge, kfloor,
      // is the bound used from here on: a candidate below it is below at least W candidates.
      const u64 kt0 = okey(tau0), ktop = okey(vtop);
      if (tau0 > ninf && kt0 != ktop) {
        const int top0 = 64 - __clzll((long long)(kt0 ^ ktop));
        const int width0 = min(11, top0), shift0 = top0 - width0;
        const unsigned dmask0 = (1u << width0) - 1u;
        if (tid < n) {
          atomicAdd(&s_hist[(int)((okey(s_ut[tid]) >> shift0) & dmask0)], 1);
          const double c1 = s_c1[tid], c2 = s_c2[tid];
          if (c1 > tau0) atomicAdd(&s_hist[(int)((okey(c1) >> shift0) & dmask0)], 1);
          if (c2 > tau0) atomicAdd(&s_hist[(int)((okey(c2) >> shift0) & dmask0)], 1);
        }
        bar_search();
        find_bin(s_hist, 1 << width0, W, s_redi + 32, s_redi + 16, tid);
        kfloor = (ktop & ~(((u64)1 << top0) - 1)) | ((u64)s_redi[16] << shift0);
      }
      const u64 kext = kfloor > kt0 + 1ull ? kfloor : kt0 + 1ull;
      {
        const double tl = okey_inv(kext);  // lists are supersets: non-strict comparisons against the bound
        if (warp * 32 < n) {
          bool lv = false;
          if (tid < n) {
            const double pt = g_pt[ao + tid];
            lv = pt > ninf && pt + lpmax >= tl;
          }
          const unsigned bal = __ballot_sync(0xffffffffu, lv);
          int base = 0;
          if (lane == 0 && bal) base = atomicAdd(&s_redi[22], __popc(bal));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (lv) s_liveP[base + __popc(bal & ((1u << lane) - 1u))] = tid;
        }
        for (int c0 = warp * 32; c0 < C; c0 += kSearchThreads) {
          const int cc = c0 + lane;
          const bool lv = cc < C && cc != blank && BEAM_LP(cc) + ptmax >= tl;
          const unsigned bal = __ballot_sync(0xffffffffu, lv);
          int base = 0;
          if (lane == 0 && bal) base = atomicAdd(&s_redi[23], __popc(bal));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (lv) s_liveL[base + __popc(bal & ((1u << lane) - 1u))] = cc;
        }
      }
    } else {
      // only the warps that have a list to build combine the partials; thread 0 publishes tau0 and vtop
      if (warp * 32 < n || warp * 32 < C) {
        int cntv = 0;
        double mn = __longlong_as_double(0x7ff0000000000000ll), ptmax = ninf, utmax = ninf;
        for (int w = 0; w * 32 < n; w++) {
          cntv += s_redi[w];
          mn = fmin(mn, s_redd[w]);
          ptmax = fmax(ptmax, s_redd[16 + w]);
          utmax = fmax(utmax, s_redd[44 + w]);
        }
        const double tau0 = cntv >= W ? mn : ninf;
        if (tid == 0) {  // every candidate's score lies in (tau0, vtop] (kept prefixes: [tau0, vtop])
          s_redd[32] = tau0;
          s_redd[33] = fmax(utmax, ptmax + lpmax);
        }
        if (warp * 32 < n) {
          bool lv = false;
          if (tid < n) {
            const double pt = g_pt[ao + tid];
            lv = pt > ninf && pt + lpmax > tau0;
          }
          const unsigned bal = __ballot_sync(0xffffffffu, lv);
          int base = 0;
          if (lane == 0 && bal) base = atomicAdd(&s_redi[22], __popc(bal));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (lv) s_liveP[base + __popc(bal & ((1u << lane) - 1u))] = tid;
        }
        for (int c0 = warp * 32; c0 < C; c0 += kSearchThreads) {
          const int cc = c0 + lane;
          const bool lv = cc < C && cc != blank && BEAM_LP(cc) + ptmax > tau0;
          const unsigned bal = __ballot_sync(0xffffffffu, lv);
          int base = 0;
          if (lane == 0 && bal) base = atomicAdd(&s_redi[23], __popc(bal));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (lv) s_liveL[base + __popc(bal & ((1u << lane) - 1u))] = cc;
        }
      }
    }
    bar_search();
    if (!STAGE2) {
      tau0 = s_redd[32];
      vtop = s_redd[33];
    }
    BEAM_TICK(1);
    // okey(-inf) + 1 is the key of the smallest finite score
    const u64 kfin = okey(ninf) + 1ull, kt0 = okey(tau0), ktop = okey(vtop);
    const u64 kkeep = kfloor > kfin ? kfloor : kfin;             // kept prefixes: any finite score, or the bound
    const u64 kext = kfloor > kt0 + 1ull ? kfloor : kt0 + 1ull;  // extensions: strictly above tau0, and the bound
    Ctx c;
    if (BIG) {
      c.x = xr; c.m = rm; c.lse = rlse; c.lp = nullptr;
    } else {
      c.x = nullptr; c.m = 0.0; c.lse = 0.0; c.lp = lpd;
    }
    c.pb = g_pb + ao; c.pt = g_pt + ao; c.ut = s_ut; c.hash = g_hash + ao; c.last = g_last + ao;
    c.liveP = s_liveP; c.liveL = s_liveL; c.mtab = s_mtab; c.n = n; c.nL = s_redi[23]; c.TS = TS; c.tshift = tshift;
    c.CW = CW;
    c.div32 = !BIG || (u64)W * (u64)c.nL * (u64)c.nL < ((u64)1 << 32);
    c.divM = c.nL >= 2 ? 0xffffffffu / (unsigned)c.nL + 1u : 0u;
    c.kkeep = kkeep;
    c.kext = kext;
    const int nitems = n + s_redi[22] * c.nL;
    const bool cached = CO || nitems <= kIPT * kSearchThreads;  // else: recompute the candidates in every pass
    // ---- 3. the candidates' keys and how many they are
    u64 key[kIPT];
    int cnt = 0;
    if (cached) {
#pragma unroll
      for (int r = 0; r < kIPT; r++) {
        key[r] = 0;
        if (r * kSearchThreads < nitems) {
          const int i = tid + r * kSearchThreads;
          if (i < nitems) key[r] = eval_key<BIG>(c, i);
          cnt += key[r] != 0;
        }
      }
    } else {
      for (int i = tid; i < nitems; i += kSearchThreads) cnt += eval_key<BIG>(c, i) != 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) s_redi[warp] = cnt;
    bar_search();
    cnt = s_redi[lane & (kSearchWarps - 1)];
#pragma unroll
    for (int o = kSearchWarps / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    // bounds of the keys, known without looking at them: (key of tau0 .. key of vtop]
    // with a full beam no candidate is below the worst kept total (key kt0), nor below the second-stage bound
    const u64 kmax = ktop, kmin = tau0 > ninf ? (kfloor > kt0 ? kfloor : kt0) : kfin;
    BEAM_TICK(2);
    // ---- 4. threshold of the best W: admit k >= F1, and among k == F1 (tie_mode) those with k2 >= F2
    u64 F1 = 1, F2 = 0;
    bool tie_mode = false;
    if (cnt > W) {
      int need = W;
      const u64 diff = kmin ^ kmax;
      if (__builtin_expect(diff == 0, 0)) {
        F1 = kmax;
        tie_mode = true;
      } else {
        u64 prefix = kmax;
        int top = 64 - __clzll((long long)diff);  // bits [top, 64) are common to every candidate
        for (;;) {
          const int width = min(11, top), shift = top - width;
          const unsigned dmask = (1u << width) - 1u;
          if (cached) {
#pragma unroll
            for (int r = 0; r < kIPT; r++) {
              const u64 k = key[r];
              if (r * kSearchThreads < nitems && k && (top >= 64 || ((k ^ prefix) >> top) == 0))
                atomicAdd(&s_hist[(int)((k >> shift) & dmask)], 1);
            }
          } else {
            for (int i = tid; i < nitems; i += kSearchThreads) {
              const u64 k = eval_key<BIG>(c, i);
              if (k && (top >= 64 || ((k ^ prefix) >> top) == 0)) atomicAdd(&s_hist[(int)((k >> shift) & dmask)], 1);
            }
          }
          bar_search();
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
      if (__builtin_expect(tie_mode, 0)) {  // rare: keep it out of the hot path's instruction stream
        // order the candidates that share the threshold score by (active first, hash)
        u64 prefix = 0;
        int top = 64;
        for (;;) {
          const int width = min(11, top), shift = top - width;
          const unsigned dmask = (1u << width) - 1u;
          if (cached) {
#pragma unroll
            for (int r = 0; r < kIPT; r++) {
              if (key[r] == F1) {
                const u64 k = item_k2<BIG>(c, tid + r * kSearchThreads);
                if (top >= 64 || ((k ^ prefix) >> top) == 0) atomicAdd(&s_hist[(int)((k >> shift) & dmask)], 1);
              }
            }
          } else {
            for (int i = tid; i < nitems; i += kSearchThreads) {
              if (eval_key<BIG>(c, i) == F1) {
                const u64 k = item_k2<BIG>(c, i);
                if (top >= 64 || ((k ^ prefix) >> top) == 0) atomicAdd(&s_hist[(int)((k >> shift) & dmask)], 1);
              }
            }
          }
          bar_search();
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
    BEAM_TICK(3);
    // ---- 5. the admitted candidates, as a list
    if (cached) {
      // one shared-memory atomic per warp: the lanes' counts are scanned, lane 31 reserves the warp's slots
      unsigned adm = 0;
#pragma unroll
      for (int r = 0; r < kIPT; r++) {
        const u64 k = key[r];
        if (r * kSearchThreads < nitems && k >= F1 &&
            !(tie_mode && k == F1 && item_k2<BIG>(c, tid + r * kSearchThreads) < F2))
          adm |= 1u << r;
      }
      const int mine = __popc(adm);
      const int incl = warp_incl_scan(mine, lane);
      int base = 0;
      if (lane == 31 && incl) base = atomicAdd(&s_redi[20], incl);
      int slot = __shfl_sync(0xffffffffu, base, 31) + incl - mine;
#pragma unroll
      for (int r = 0; r < kIPT; r++) {
        if ((adm >> r) & 1u) {
          if (slot < W) {  // beyond W only through a hash collision among exact ties
            s_adm_i[slot] = tid + r * kSearchThreads;
            s_adm_k[slot] = key[r];
          }
          slot++;
        }
      }
    } else {
      for (int i = tid; i < nitems; i += kSearchThreads) {
        const u64 k = eval_key<BIG>(c, i);
        if (k >= F1 && !(tie_mode && k == F1 && item_k2<BIG>(c, i) < F2)) {
          const int slot = atomicAdd(&s_redi[20], 1);
          if (slot < W) {
            s_adm_i[slot] = i;
            s_adm_k[slot] = k;
          }
        }
      }
    }
    bar_search();
    BEAM_TICK(4);
    // ---- 6. one thread per admitted candidate builds its entry of the next beam
    const int n_new = min(s_redi[20], W);
    bool orphan = false;
    if (tid < n_new) {
      int b, l;
      item_of<BIG>(c, s_adm_i[tid], b, l);
      const double v = okey_inv(s_adm_k[tid]);
      const int d = no + tid, a = ao + b;
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
        const int po = g_pslot[a];  // the parent's slot in the active beam; mapped to the next beam below
        g_pslot[d] = po;
        orphan = po < 0 && g_len[a] > 0;
        s_newslot[b] = tid;
      } else {
        const int id = 1 + t * W + tid;  // frame t owns nodes [1 + t*W, 1 + (t+1)*W)
        nodes[id] = make_int2(g_node[a], l);
        const u64 h = child_hash(g_hash[a], l);
        g_pb[d] = ninf;
        g_pl[d] = v;
        g_pt[d] = v;
        g_hash[d] = h;
        g_phash[d] = g_hash[a];
        g_node[d] = id;
        g_len[d] = g_len[a] + 1;
        g_last[d] = l;
        g_plast[d] = g_last[a];
        g_pslot[d] = b;
        // a prefix that enters the beam may be the parent of prefixes that stayed in it without it
        unsigned idx = (unsigned)h & (unsigned)(TS - 1);
        for (;;) {
          const u64 old = atomicCAS(&s_tab_key[idx], 0ull, h);
          if (old == 0ull || old == h) {
            s_tab_slot[idx] = tid;
            break;
          }
          idx = (idx + 1) & (unsigned)(TS - 1);
        }
      }
    }
    // the set of active extensions is rebuilt in phase 7 for the next frame
    if (!BIG) {
      if (tid < n_new)
        for (int w = 0; w < CW; w++) s_mtab[tid * CW + w] = 0;
    } else {
      for (int i = tid; i < TS; i += kSearchThreads) s_mtab[i] = kEmpty;
    }
    bar_search();
    BEAM_TICK(5);
    // ---- 7. parents' slots in the next beam, then the per-parent sets of active extensions
    if (tid < n_new) {
      const int d = no + tid;
      int ps = g_pslot[d];
      if (ps >= 0) {
        ps = s_newslot[ps];
      } else if (orphan) {
        const u64 ph = g_phash[d];
        const int ln = g_len[d] - 1;
        unsigned idx = (unsigned)ph & (unsigned)(TS - 1);
        for (;;) {
          const u64 k = s_tab_key[idx];
          if (k == 0ull) break;
          if (k == ph && g_len[no + s_tab_slot[idx]] == ln) {
            ps = s_tab_slot[idx];
            break;
          }
          idx = (idx + 1) & (unsigned)(TS - 1);
        }
      }
      g_pslot[d] = ps;
      if (ps >= 0) {
        if (!BIG)
          atomicOr(&s_mtab[ps * CW + (g_last[d] >> 5)], 1u << (g_last[d] & 31));
        else
          mtab_insert(s_mtab, TS, tshift, ps, g_last[d]);
      }
    }
    BEAM_TICK(6);
    __syncthreads();  // frame over; the producer warp has row t+1 ready
    n = n_new;
    cur ^= 1;
#undef BEAM_LP
  }
  if (profiling)
    for (int k = 0; k < 8; k++) prof[k] = (long long)s_redu[32 + k];
#undef BEAM_TICK

  // ---- read the best P prefixes back through the trie (search warps only from here on)
  const int ao = cur * W;
  for (int e = tid; e < n; e += kSearchThreads) {
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
      s_liveP[r] = len;  // s_liveP is free now (P <= W)
      log_prob[(size_t)b_utt * P + r] = (float)v;
    }
  }
  for (int r = n + tid; r < P; r += kSearchThreads) {
    s_liveP[r] = 0;
    log_prob[(size_t)b_utt * P + r] = -INFINITY;
  }
  bar_search();
  for (int r = warp; r < P; r += kSearchWarps) {
    const int len = s_liveP[r];
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
  NASR_CHECK_ARG((logits || T == 0) && seq_len && (hyp || T == 0) && hyp_len && log_prob && workspace,
                 "ctc_beam_search: NULL argument");
  size_t need = 0;
  ctc_beam_workspace_bytes(T, B, C, W, &need);
  if (workspace_bytes < need) {
    set_error("ctc_beam_search: workspace of %zu bytes, %zu needed", workspace_bytes, need);
    return NASR_ERR_WORKSPACE_TOO_SMALL;
  }
  const BeamLayout L = beam_layout(W, C);
  const size_t smem = (size_t)L.total;
  // (beam_width * C < 2^20 keeps the candidate index arithmetic of the narrow-vocabulary build exact; the build for
  // more than kBitSetMaxC classes checks per frame instead)
  if (W > kSearchThreads || C > 8192 || (C <= kBitSetMaxC && (size_t)W * C >= ((size_t)1 << 20)) || smem > 200 * 1024) {
    set_error("ctc_beam_search: beam_width=%d with C=%d is not supported (limits: beam_width <= %d, C <= 8192, "
              "beam_width * C < 2^20 for C <= %d, shared memory %zu of 204800 bytes)", W, C, kSearchThreads,
              kBitSetMaxC, smem);
    return NASR_ERR_UNSUPPORTED;
  }
  const bool big = C > kBitSetMaxC;
  // four producer warps (then one CTA per SM): always for very wide rows; for moderately wide ones while the batch
  // does not need a second CTA per SM anyway
  int num_sms = 0;
  NASR_CUDA(device_sm_count(&num_sms));
  const bool four = C > kOneProducerMaxC || (C > kStage2MinC * 4 && B <= num_sms);
  // second-stage bound: always for wide vocabularies; for narrow ones while there is one CTA per SM (the frame is
  // then latency bound and the shorter lists pay: B=64, C=38 6.1 -> 5.3 ms on N(0,1)*3, 5.4 -> 5.1 ms on planted
  // alignments; at B=256 it is -5 % / +5 %, so it stays off there)
  const bool stage2 = C > kStage2MinC || B <= num_sms;
  const bool one_cta = B <= num_sms;  // (implies stage2)
  const bool co = (size_t)W * ((size_t)C + 1) <= (size_t)kIPT * kSearchThreads;  // every frame fits the cached path
  auto kernel = big ? ctc_beam_kernel<true, true, 4, 1, false>
                    : (four ? ctc_beam_kernel<true, false, 4, 1, false>   /* four implies C > 64: second-stage bound on */
                            : (one_cta ? (co ? ctc_beam_kernel<true, false, 1, 1, true>
                                             : ctc_beam_kernel<true, false, 1, 1, false>)
                                       : (stage2 ? ctc_beam_kernel<true, false, 1, 2, false>
                                                 : (co ? ctc_beam_kernel<false, false, 1, 2, true>
                                                       : ctc_beam_kernel<false, false, 1, 2, false>))));
  NASR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int2* nodes = reinterpret_cast<int2*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  kernel<<<B, kSearchThreads + 32 * (four ? 4 : 1), smem, stream>>>(logits, T, B, C, st_t, st_b, seq_len, blank, W, P, merge_repeated, hyp,
                                            hyp_len, log_prob, nodes, g_debug_prof, L);
  count_launch();
  NASR_CUDA(cudaGetLastError());
  return NASR_OK;
}

}  // namespace nasr
