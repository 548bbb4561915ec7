// Fast CTC loss + d(loss)/d(logits) for sm_100a: one CTA per utterance, eight specialised warps, one launch.
//
// Replaces tf.nn.ctc_loss + _CTCLossGrad behind create_loss (reference networks/tfnetwork.py:58-59);
// semantics per SURVEY.md Appendix A.1.  This is the throughput path; ctc_loss.cu holds the robust
// kernel (per-state exponents) that redoes any utterance this kernel flags in retry[].
//
// Arithmetic (checked on the CPU by tests/model_bfp.py against the oracle):
//   * linear-domain recursion in "ratio units": every emission is divided by the frame's blank
//     probability, R[t][c] = y[t][c] / y[t][blank], so blank states need no multiply;
//     log p gets sum_t log y_blank(t) added back at the end;
//   * pairs: slot i holds (blank state 2(i-1), label state 2(i-1)+1); lane l of a recursion warp owns the
//     NL consecutive slots [l*NL, (l+1)*NL) in registers, the s-1/s-2 neighbours of a lane's first slot
//     arrive by one warp shuffle per frame;
//   * the backward recursion is the same code on the reversed label string over descending frames;
//   * block floating point: fp64 values with one int exponent per lane, renormalised at every chunk of
//     KC frames (fp64 is used for its exponent range, not its mantissa: tools/range_study.py shows states
//     that carry posterior mass sit up to 2^-170 below their lane maximum on random logits);
//   * meet in the middle: the forward warp covers frames [0,M) while the backward warp covers [M,Tb);
//     then each continues through the other half, multiplying its pre-emission sums with the other
//     direction's rows, which a recompute warp regenerates chunk by chunk from checkpoints
//     (the [T,U] lattice never goes to memory; a checkpoint is one row per KC frames);
//   * posterior of a label state = pre-emission sum * other direction's value / p; the blank's occupancy
//     is 1 - sum of the label occupancies; grad = grad_loss * (softmax - occupancy).
//   Any event that could invalidate the result (a value flushed below 2^-1022 of its lane scale that may
//   carry mass, overflow, p = 0, shapes outside the compiled range) raises retry[b] instead.
//
// Schedule: all eight warps run one loop of "iterations" separated by __syncthreads; in iteration I
//   producer warps   turn the logits rows of the chunk used in iteration I+2 into R (double) and softmax
//                    (float) rows in shared memory, and issue the global loads for iteration I+3;
//   recompute warps  regenerate the other direction's rows for the chunk consumed in iteration I+1;
//   recursion warps  advance alpha / beta over the chunk of iteration I (phase 2: emit posteriors);
//   gradient warps   reduce the posteriors of iteration I-1 by class and write grad rows.
#include <math.h>
#include <stdlib.h>

#include "nasr_common.cuh"

// Tuning hooks (nasr_debug_profile, the ablation bits of nasr_debug_config) are compiled in only with -DNASR_TUNING=1:
// the kernel is ~130 KB of SASS and waits on instruction fetch, so production builds do not carry them.
#ifndef NASR_TUNING
#define NASR_TUNING 0
#endif
#define NASR_PROF_PTR(p) (NASR_TUNING ? (p).prof : (long long*)nullptr)
#define NASR_ABLATE(p) (NASR_TUNING ? (p).ablate : 0)

namespace nasr {
namespace fast {

constexpr int KC = 8;            // frames per chunk (= rescale and checkpoint interval)
constexpr int IFIRST = -5;       // first iteration: the producers' copies run five iterations ahead of the recursion
constexpr int NTHREADS = 256;    // 8 warps
constexpr int NTHREADS_WIDE = 512;  // wide-vocabulary variant: 2 recursion + 2 recompute + 4 gradient + 8 producer warps
constexpr int GCAP = 200;        // a lane with mass sits at most this far below the nearest lane with mass beneath it
constexpr int GROWTH = 550;      // bits a lane maximum may grow inside one chunk (GCAP of inflow + 350 of emissions)
// Certificate: a value flushed in a lane of exponent Ea is < 2^(Ea-1022); its partner in the other direction is
// < 2^(Eb+1+GROWTH); so whatever a flush removes from any posterior (or from p) is below 2^(Ea+Eb-e_p-471),
// and with at most 2^20 flush events per utterance ZALARM = 400 keeps the total under 2^-50.
constexpr int ZALARM = 400;
constexpr int ENEG = -(1 << 24); // exponent tag of a lane that can never receive mass

enum Role { H_F = 0, H_B = 1, RC_F = 2, RC_B = 3, P_F = 4, P_B = 5, G_F = 6, G_B = 7 };

struct Params {
  const float* logits;
  int T, B, C;
  long long st_t, st_b;  // element strides of logits and grad between frames / between utterances
  const int32_t* lab_vals;
  const int32_t* lab_offs;
  const int32_t* seq_len;
  int blank;
  float* loss;
  float* grad;
  const float* grad_loss;
  int32_t* status;
  int32_t* retry;      // [B] out: 1 = redo this utterance with the robust kernel
  uint32_t* ckpt;      // [B][2][maxch][(2*NL+1)*32]
  int maxch;
  int num_sms;
  int split;           // debug: frames of the forward half (multiple of KC), 0 = automatic
  int ablate;          // debug: bit0 gradient warps idle, bit1 producers idle in phase 2, bit2 recompute warps idle, bit3 no posterior store
  long long* prof;     // debug: [B][16 warps][4] cycles of work in phase 1 / phase 2, total, role (or NULL)
};

struct Smem {
  size_t rows, raw, rowbuf, stage, dtab, obuf, gbuf, oexp, meet_nb, meet_pre, meet_e, lab, pos, cls_off, psum, scal, total;
  int gstride;  // floats per posterior row: NL*32 cells + one zero cell (+ padding)
};

__host__ __device__ constexpr size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

// Row record of one frame, sized for the largest class count the instantiation takes (CMAX = 8*EPL):
// uint32 R_hi[CMAX+1] -- the HIGH WORDS of the double ratio emissions, rounded to nearest at the 20 mantissa
// bits they keep (entry CMAX is 0, the "emission" of dead slots) -- then float y[CMAX] (softmax, for the
// gradient).  4-byte entries because a C=38 row of doubles puts three classes on every bank pair and the
// recursion's gathers then cost 3.75 wavefronts each (profiles/r1_v1_summary.md); a row of words is almost
// conflict free.  The rounding error (2.4e-7 relative, unbiased) is far inside the gradient tolerance.
// Compile-time sizes so that frame offsets are immediates.
__host__ __device__ constexpr int row_bytes(int cmax) { return (int)al16((size_t)(2 * cmax + 1) * 4); }
__host__ __device__ constexpr int row_yoff(int cmax) { return (cmax + 1) * 4; }

// wideC > 0 selects the wide-vocabulary layout (C > 64): a row record is then uint32 R_hi[NL*32] in the SLOT
// order of the side it belongs to (the producer gathers the emissions of the transcript's classes, so the
// record does not grow with C), there is no raw-logits ring, every producer warp owns one row of C floats
// (rowbuf), and the gradient warps get a table of the distinct classes of the transcript (dtab).
constexpr int kWideProducers = 8;
constexpr int kWideRegC = 1024;   // widest row a producer warp holds in registers (32 floats per lane)
constexpr int kWideMaxC = 8192;   // widest row of the wide variant: beyond kWideRegC the producers stream the row
__host__ __device__ inline Smem smem_layout(int NL, int cmax, int wideC = 0, bool stream = false) {
  Smem s;
  const int Lcap = NL * 32;
  const int rowbytes = wideC ? NL * 32 * 4 : row_bytes(cmax);
  if (wideC) cmax = wideC;
  s.gstride = NL * 32 + 4;
  size_t o = 0;
  s.rows = o;      o = al16(o + (size_t)2 * 4 * KC * rowbytes);              // [side][4 slots][KC] records
  s.raw = o;       o = al16(o + (wideC ? 0 : (size_t)2 * 4 * KC * cmax * 4)); // [side][4 slots][KC][cmax] raw logits
  const int stagedC = stream ? 0 : wideC;  // streamed rows are never staged
  s.rowbuf = o;    o = al16(o + (size_t)kWideProducers * stagedC * 4);
  s.stage = o;     o = al16(o + (size_t)kWideProducers * 2 * stagedC * 4);  // raw rows in flight
  s.dtab = o;      o = al16(o + (wideC ? (size_t)3 * Lcap * 4 : 0));
  s.obuf = o;      o = al16(o + (size_t)2 * 2 * KC * NL * 32 * 4);           // [side][2][KC][NL][32] high words
  s.gbuf = o;      o = al16(o + (size_t)2 * 2 * KC * s.gstride * 4);         // [side][2][KC][gstride] posteriors
  s.oexp = o;      o = al16(o + (size_t)2 * 2 * 32 * 4);
  s.meet_nb = o;   o = al16(o + (size_t)NL * 32 * 8);
  s.meet_pre = o;  o = al16(o + (size_t)NL * 32 * 8);
  s.meet_e = o;    o = al16(o + 32 * 4);
  s.lab = o;       o = al16(o + (size_t)Lcap * 4);
  s.pos = o;       o = al16(o + (size_t)Lcap * 2);                           // class-sorted rank of label j
  s.cls_off = o;   o = al16(o + (size_t)(cmax + 2) * 4);
  s.psum = o;      o = al16(o + 16 * 8);
  s.scal = o;      o = al16(o + 64);
  s.total = o;
  return s;
}

// scalars block: [0] alarm (int) [1] e_p (int) [2] rep count (int) [4..5] 1/m_p (double)
struct Sched {
  int Tb;
  int n1F, n1B;      // frames each direction covers in phase 1
  int nch1F, nch1B;  // chunks of phase 1
  int offB;          // iteration at which the backward warp starts phase 1
  int P1;            // iteration of the meeting; phase 2 starts at P1 + 1
  int last;          // last iteration (gradient of the last phase-2 chunk)
};

struct Chunk {
  int phase;  // 0 none, 1, 2
  int idx;    // phase 1: own chunk index; phase 2: the other direction's chunk index
  int len;    // frames
  int base;   // frame of position 0; position f is frame base + f (forward) or base - f (backward)
};

__device__ __forceinline__ Chunk chunk_at(const Sched& S, int d, int J) {
  Chunk c;
  c.phase = 0; c.idx = 0; c.len = 0; c.base = 0;
  const int k = J - (d ? S.offB : 0);
  const int own_n = d ? S.n1B : S.n1F, own_ch = d ? S.nch1B : S.nch1F;
  const int oth_n = d ? S.n1F : S.n1B, oth_ch = d ? S.nch1F : S.nch1B;
  if (J >= 0 && k >= 0 && k < own_ch) {
    const int tau0 = k * KC;
    c.phase = 1;
    c.idx = k;
    c.len = min(KC, own_n - tau0);
    c.base = d ? S.Tb - 1 - tau0 : tau0;
  } else if (J > S.P1) {
    const int q = J - S.P1 - 1;
    if (q < oth_ch) {
      const int j = oth_ch - 1 - q;
      const int tau0 = j * KC;
      const int e = min(oth_n, tau0 + KC);
      c.phase = 2;
      c.idx = j;
      c.len = e - tau0;
      c.base = d ? e - 1 : S.Tb - e;
    }
  }
  return c;
}

// CTA-wide barrier that may be reached from different code locations (every role loop has its own)
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 0;" ::: "memory"); }

__device__ __forceinline__ double pow2d(int e) {  // 2^e, e clamped to the normal range
  e = max(-1022, min(1023, e));
  return __hiloint2double((1023 + e) << 20, 0);
}
__device__ __forceinline__ double hi2d(uint32_t hi) { return __hiloint2double((int)hi, 0); }

template <int NL>
struct Dir {
  double Ab[NL], Al[NL];
  int E;
  uint32_t coloff[NL];  // byte offset of R[class of slot k] inside a row record (dead slot: the zero entry)
  double skip[NL];      // 1.0 if slot k may take the skip transition (label differs from the previous one), else 0.0
};

// Slot tables and the virtual row before the first frame, for direction d (0 forward, 1 mirrored).
// wide: 0 = narrow records indexed by class; 1 = slot-ordered records read in this direction's own order
// (recursion warps); 2 = slot-ordered records of the OTHER side, read mirrored (recompute warps).
template <int NL>
__device__ __forceinline__ void dir_setup(Dir<NL>& s, int d, int lane, const int* lab, int L, int CZ, int wide = 0) {
  const int N = NL * 32;
  const int pad = N - L - 1;
#pragma unroll
  for (int k = 0; k < NL; k++) {
    const int i = lane * NL + k;
    int col = CZ;  // zero entry of the row record
    bool skip = false;
    if (d == 0) {
      const int j = i - 1;
      if (j >= 0 && j < L) {
        col = lab[j];
        skip = j >= 1 && lab[j] != lab[j - 1];
      }
    } else {
      const int m = i - pad;
      if (m >= 0 && m < L) {
        col = lab[L - 1 - m];
        skip = m >= 1 && lab[L - 1 - m] != lab[L - m];
      }
    }
    s.coloff[k] = (uint32_t)col * 4u;
    if (wide == 1) s.coloff[k] = (uint32_t)(k * 32 + lane) * 4u;
    if (wide == 2) s.coloff[k] = (uint32_t)((NL - 1 - k) * 32 + 31 - lane) * 4u;
    s.skip[k] = skip ? 1.0 : 0.0;
    s.Ab[k] = (i == (d == 0 ? 1 : pad)) ? 1.0 : 0.0;
    s.Al[k] = 0.0;
  }
  s.E = 0;
}

// Alarm reasons (bit mask written to retry[b]; any non-zero value sends the utterance to the robust kernel)
enum Alarm { AL_SHAPE = 1, AL_EMISSION = 2, AL_NONFINITE = 4, AL_GROWTH = 8, AL_TAG = 16, AL_P = 32, AL_RANGE = 64 };

// Bring the lane maximum back to [1,2) and fold the shift into E.  Two rules keep the factor that carries
// a value from lane l-1 into lane l representable: a lane with mass sits at most GCAP below the nearest
// lane with mass beneath it, and a lane without mass adopts that lane's exponent unchanged.
template <int NL>
__device__ __forceinline__ void rescale(Dir<NL>& s, int lane, int& alarm) {
  int mh = 0;
#pragma unroll
  for (int k = 0; k < NL; k++) {
    mh = max(mh, __double2hiint(s.Ab[k]));
    mh = max(mh, __double2hiint(s.Al[k]));
  }
  const int ex = (mh >> 20) & 0x7ff;
  if (ex == 0x7ff) alarm |= AL_NONFINITE;           // inf / nan: arithmetic broke
  const bool nz = ex != 0;
  const int e = ex - 1023;
  if (nz && e > GROWTH) alarm |= AL_GROWTH;
  if (nz && s.E == ENEG) alarm |= AL_TAG;           // mass in a lane tagged unreachable: cannot happen
  const unsigned below = __ballot_sync(0xffffffffu, nz) & (0xffffffffu >> (31 - lane));
  const int hops = GCAP * __popc(below);            // GCAP per lane with mass at or below this one
  int v = (nz && s.E != ENEG) ? s.E + e + hops : INT_MIN / 2;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v = max(v, u);
  }
  int En = v - hops;
  if (v < -(1 << 28)) En = ENEG;                    // no mass at or below this lane, ever
  const double f = (En == ENEG || s.E == ENEG) ? 1.0 : pow2d(s.E - En);
#pragma unroll
  for (int k = 0; k < NL; k++) {
    s.Ab[k] *= f;
    s.Al[k] *= f;
  }
  s.E = En;
}

// 2^(E[lane-1] - E[lane]): the factor that brings the lower neighbour's last label value into this
// lane's scale (0 for lane 0 and for neighbours that hold no mass).
__device__ __forceinline__ double inflow_factor(int E, int lane) {
  const int Eb = __shfl_up_sync(0xffffffffu, E, 1);
  if (lane == 0 || Eb == ENEG || E == ENEG) return 0.0;
  return pow2d(Eb - E);
}

enum Mode { PLAIN = 0, STORE_O = 1, COMBINE = 2 };

// One frame of one direction.  erow: the frame's row record; orow: the frame's row of high words
// ([k][lane]); grow: the frame's posterior row.  All loads of the frame are issued before any store so
// that the NL slots form NL independent dependency chains.
//   STORE_O : write the high words of the new label values to orow[k][lane]
//   COMBINE : posterior = pre-emission sum * (other direction's high word at the mirrored slot, shifted by
//             kshift in the exponent field) -> grow[gphys[k]]
// (A hand software-pipelined variant -- next frame's emissions prefetched, the shuffle issued a frame early
// -- was measured: 135 instead of 172 cycles per frame for a lone warp, but slower inside the full kernel,
// where the schedulers are shared by four warps; tools/microbench_step.cu, tools/microbench_mix.cu.)
template <int NL, int MODE>
__device__ __forceinline__ void step(Dir<NL>& s, const unsigned char* erow, uint32_t* orow, unsigned char* grow,
                                     double fin, int kshift, const uint32_t (&gphys)[NL], int lane) {
  uint32_t rh[NL];
  int oh[NL];
#pragma unroll
  for (int k = 0; k < NL; k++) rh[k] = *reinterpret_cast<const uint32_t*>(erow + s.coloff[k]);
  if (MODE == COMBINE) {
#pragma unroll
    for (int k = 0; k < NL; k++) oh[k] = (int)orow[(NL - 1 - k) * 32 + 31 - lane];
  }
  double a_in = __shfl_up_sync(0xffffffffu, s.Al[NL - 1], 1);
  a_in = lane ? a_in * fin : 0.0;
#pragma unroll
  for (int k = NL - 1; k >= 0; k--) {
    const double alp = k > 0 ? s.Al[k - 1] : a_in;
    const double nb = s.Ab[k] + alp;
    // q = Al + Ab + [skip allowed] * alp: one DFMA with a {0,1} constant instead of a bitwise select of
    // nb / Ab (2 LOP3 on the half-rate ALU pipe cost more dispatch than the extra FP64 operation)
    const double q = fma(s.skip[k], alp, s.Al[k] + s.Ab[k]);
    if (MODE == COMBINE) {
      const double od = __hiloint2double(max(oh[k] + kshift, 0), 0);
      *reinterpret_cast<float*>(grow + gphys[k]) = (float)(q * od);
    }
    s.Al[k] = q * hi2d(rh[k]);
    if (MODE == STORE_O) orow[k * 32 + lane] = (uint32_t)__double2hiint(s.Al[k]);
    s.Ab[k] = nb;
  }
}

// Advance one direction over a chunk of `len` frames whose row records start at erows (consumer order).
// reverse: walk the records backwards (recompute warps run against the consumer's order).
// UNR: unroll factor of the eight-frame chunk loop.  The kernel waits on instruction fetch, so the right factor is
// not the largest: measured (same-box A/B) 8 is best for the narrow kernel at small batches, 2 for the wide variant
// (cfg5 0.317 -> 0.309 ms) and within 1 % of 8 for the narrow kernel at cfg3.
template <int NL, int MODE, int ROWB, int UNR = KC>
__device__ __forceinline__ void run_chunk(Dir<NL>& s, const unsigned char* erows, int len,
                                          bool reverse, uint32_t* obuf, float* gbuf, double fin,
                                          int kshift, const uint32_t (&gphys)[NL], int lane) {
  constexpr int rowbytes = ROWB;
  constexpr int gstride = NL * 32 + 4;
  if (len == KC) {
#pragma unroll(UNR)
    for (int g = 0; g < KC; g++) {
      const int f = reverse ? KC - 1 - g : g;
      step<NL, MODE>(s, erows + f * rowbytes, obuf + f * (NL * 32), reinterpret_cast<unsigned char*>(gbuf + f * gstride),
                     fin, kshift, gphys, lane);
    }
  } else {
#pragma unroll 1
    for (int g = 0; g < len; g++) {
      const int f = reverse ? len - 1 - g : g;
      step<NL, MODE>(s, erows + f * rowbytes, obuf + f * (NL * 32), reinterpret_cast<unsigned char*>(gbuf + f * gstride),
                     fin, kshift, gphys, lane);
    }
  }
}

// Producer, one group of four rows (8 lanes per row, lane sub handles classes sub, sub+8, ...):
// raw logits of row f (staged in shared memory by cp.async; columns C..CMAX-1 hold -inf, written once at
// kernel start, so no class-count predicates are needed here) -> row record f of the chunk: high words of
// the ratio emissions, softmax as floats.  Returns log y_blank of the row in the lane that holds the blank
// class (0 elsewhere and for rows past the end of a short chunk).  One copy of this code serves the
// producer and, in phase 1, the gradient warps.
template <int EPL, int ROWB, int YOFF>
__device__ __noinline__ float rows_to_smem(const float* raw, int f, int len, unsigned char* rows, int C, int blank,
                                           int* alarm_word) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & 7;
  const float* xrow = raw + f * (8 * EPL);
  float x[EPL];
#pragma unroll
  for (int e = 0; e < EPL; e++) x[e] = xrow[sub + 8 * e];
  float m = x[0];
#pragma unroll
  for (int e = 1; e < EPL; e++) m = fmaxf(m, x[e]);
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  float n[EPL];
  float ssum = 0.f, nmin = 1.f;
  const float ml2 = m * 1.4426950408889634f;
#pragma unroll
  for (int e = 0; e < EPL; e++) {
    n[e] = exp2f(fmaf(x[e], 1.4426950408889634f, -ml2));   // e^(x-m); 0 for the -inf padding columns
    ssum += n[e];
    nmin = fminf(nmin, sub + 8 * e < C ? n[e] : 1.f);
  }
  ssum += __shfl_xor_sync(0xffffffffu, ssum, 4);
  ssum += __shfl_xor_sync(0xffffffffu, ssum, 2);
  ssum += __shfl_xor_sync(0xffffffffu, ssum, 1);
  nmin = fminf(nmin, __shfl_xor_sync(0xffffffffu, nmin, 4));
  nmin = fminf(nmin, __shfl_xor_sync(0xffffffffu, nmin, 2));
  nmin = fminf(nmin, __shfl_xor_sync(0xffffffffu, nmin, 1));
  // the blank's numerator and logit, from the lane and register that hold class `blank`
  float nbl = 0.f, xbl = 0.f;
#pragma unroll
  for (int e = 0; e < EPL; e++)
    if (e == (blank >> 3)) {
      nbl = n[e];
      xbl = x[e];
    }
  const int src = (lane & ~7) | (blank & 7);
  nbl = __shfl_sync(0xffffffffu, nbl, src);
  const float inv_s = __fdividef(1.0f, ssum);
  const float inv_nb = __fdividef(1.0f, nbl);
  float logyb = 0.f;
  if (f < len) {
    unsigned char* row = rows + (size_t)f * ROWB;
    uint32_t* Rrow = reinterpret_cast<uint32_t*>(row);
    float* yrow = reinterpret_cast<float*>(row + YOFF);
#pragma unroll
    for (int e = 0; e < EPL; e++) {
      const int c = sub + 8 * e;
      // high word of (double)(n/n_blank) from the bits of the float, rounded to nearest (no F2F here);
      // the padding columns get garbage that nothing reads
      Rrow[c] = ((__float_as_uint(n[e] * inv_nb) + 4u) >> 3) + (896u << 20);
      yrow[c] = n[e] * inv_s;
    }
    // every ratio n/n_blank must be a normal float: n >= FLT_MIN (then n/n_blank >= FLT_MIN because
    // n_blank <= 1) and n_blank >= 1e-38 (then n/n_blank <= 1e38).  Anything else -- a class or blank
    // probability that underflowed, inf, nan -- is the robust kernel's business.
    if (!(nmin >= 1.1754944e-38f && nbl >= 1.0e-38f && ssum <= 3.0e38f)) atomicOr(alarm_word, (int)AL_EMISSION);
    if (sub == (blank & 7)) logyb = (xbl - m) - __logf(ssum);
  }
  return logyb;
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- wide-vocabulary producer (64 < C <= 1024, C % 4 == 0): one warp per row, the row in registers -------
// lane l holds classes 4l + 128i + {0,1,2,3}, i = 0..7 (eight 16-byte loads, coalesced)
struct WideRow {
  float4 v[8];
};

// global -> this warp's staging row, asynchronously (16-byte LDGSTS; the row is consumed an iteration later)
__device__ __forceinline__ void wide_issue(float* stage, const float* xrow, int C, int lane) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int c = 4 * lane + 128 * i;
    if (c < C)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(stage + c)),
                   "l"(xrow + c)
                   : "memory");
  }
}
__device__ __forceinline__ void wide_fetch(WideRow& r, const float* stage, int C, int lane) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int c = 4 * lane + 128 * i;
    r.v[i] = c < C ? *reinterpret_cast<const float4*>(stage + c)
                   : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  }
}

// One row: softmax pieces, the slot-ordered record of ratio emissions (high words) for the recursion, and --
// in phase 2 -- the provisional gradient row grad_loss * softmax, from which the gradient warps later subtract
// the occupancies of the transcript's classes.  pcls[k] = class of slot (lane, k) of this side, -1 if dead.
// Returns log y_blank (all lanes).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int NL>
__device__ __forceinline__ float wide_row(const WideRow& r, float* rowbuf, uint32_t* rec, const int (&pcls)[NL],
                                          float* grow, float gs, int C, int blank, int lane, int& alarm) {
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; i++) m = fmaxf(m, fmaxf(fmaxf(r.v[i].x, r.v[i].y), fmaxf(r.v[i].z, r.v[i].w)));
  m = warp_max(m);
  const float ml2 = m * 1.4426950408889634f;
  float ssum = 0.f;
  // e^(x - m) straight into the warp's row buffer (kept out of registers: two rows of logits are in flight)
#pragma unroll
  for (int i = 0; i < 8; i++) {
    float4 n;
    n.x = ex2_approx(fmaf(r.v[i].x, 1.4426950408889634f, -ml2));
    n.y = ex2_approx(fmaf(r.v[i].y, 1.4426950408889634f, -ml2));
    n.z = ex2_approx(fmaf(r.v[i].z, 1.4426950408889634f, -ml2));
    n.w = ex2_approx(fmaf(r.v[i].w, 1.4426950408889634f, -ml2));
    ssum += (n.x + n.y) + (n.z + n.w);
    const int c = 4 * lane + 128 * i;
    if (c < C) *reinterpret_cast<float4*>(rowbuf + c) = n;
  }
  ssum = warp_sum(ssum);
  __syncwarp();
  const float nbl = rowbuf[blank];
  const float inv_nb = __fdividef(1.0f, nbl);
  bool bad = !(nbl >= 1.0e-38f && ssum <= 3.0e38f);
#pragma unroll
  for (int k = 0; k < NL; k++) {
    uint32_t w = 0u;
    if (pcls[k] >= 0) {
      const float rr = rowbuf[pcls[k]] * inv_nb;
      bad |= !(rr >= 1.1754944e-38f && rr <= 1.0e38f);   // only the transcript's classes have to stay normal
      w = ((__float_as_uint(rr) + 4u) >> 3) + (896u << 20);
    }
    rec[k * 32 + lane] = w;
  }
  if (bad) alarm |= AL_EMISSION;
  if (grow) {
    const float sc = gs * __fdividef(1.0f, ssum);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int c = 4 * lane + 128 * i;
      if (c < C) {
        const float4 n = *reinterpret_cast<const float4*>(rowbuf + c);
        __stcg(reinterpret_cast<float4*>(grow + c), make_float4(n.x * sc, n.y * sc, n.z * sc, n.w * sc));
      }
    }
  }
  __syncwarp();
  return __logf(nbl) - __logf(ssum);
}

// Streamed rows (wide variant, STREAM): rows wider than kWideRegC, or whose length is not a multiple of four, or that
// are not 16-byte aligned.  The producer warp streams the row from global memory instead of holding it -- one pass
// for the maximum and the sum of e^(x-m) together (running maximum per lane, sum rescaled when it rises) and, in
// phase 2, one that writes the provisional gradient and finds the row in L2 (a row is at most 32 KB).  Up to three scalars are peeled off at either end so
// that the body moves in aligned 16-byte accesses whatever the row's address (the gradient row has the same
// misalignment: the host checks that the two base pointers agree modulo 16).  The emissions of the transcript's
// classes are gathered straight from the row.  ex2.approx on fma(x, log2 e, -m log2 e) like wide_row.
constexpr int kHV = 16;  // 16-byte loads a lane keeps in flight per slice of a streamed row
template <int NL>
__device__ __forceinline__ float huge_row(const float* __restrict__ x, uint32_t* rec, const int (&pcls)[NL],
                                          float* grow, float gs, int C, int blank, int lane, int& alarm) {
  const float kL2E = 1.4426950408889634f;
  const float4 ninf4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  const int head = min(C, (int)((4u - (unsigned)(((uintptr_t)x >> 2) & 3u)) & 3u));  // scalars before the boundary
  const int nv = (C - head) >> 2;                                                   // aligned float4 of the body
  const int tail0 = head + 4 * nv, ntail = C - tail0;                               // up to three scalars after it
  const float4* xv = reinterpret_cast<const float4*>(x + head);
  const bool edge = lane < head || (lane >= 4 && lane - 4 < ntail);                 // lanes 0-2: head, 4-6: tail
  const int ec = lane < 4 ? lane : tail0 + lane - 4;
  const float xe = edge ? __ldg(x + ec) : -INFINITY;
  // one pass for both the maximum and the sum: every lane keeps a running maximum ml and the sum of e^(x - ml) of
  // what it has seen, rescaled whenever a slice raises the maximum; the lanes are combined at the end
  float ml = xe, sl = edge ? 1.f : 0.f;
  for (int k0 = 0; k0 < nv; k0 += 32 * kHV) {
    float4 v[kHV];
#pragma unroll
    for (int i = 0; i < kHV; i++) {
      const int k = k0 + lane + 32 * i;
      v[i] = k < nv ? __ldg(xv + k) : ninf4;
    }
    float cm = ml;
#pragma unroll
    for (int i = 0; i < kHV; i++) cm = fmaxf(cm, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
    if (cm > -INFINITY) {  // (a lane that has seen nothing finite yet keeps sl = 0)
      const float cl2 = cm * kL2E;
      float acc = sl * ex2_approx(fmaf(ml, kL2E, -cl2));  // ml = -inf: e^(-inf) = 0 times sl = 0
#pragma unroll
      for (int i = 0; i < kHV; i++)
        acc += (ex2_approx(fmaf(v[i].x, kL2E, -cl2)) + ex2_approx(fmaf(v[i].y, kL2E, -cl2))) +
               (ex2_approx(fmaf(v[i].z, kL2E, -cl2)) + ex2_approx(fmaf(v[i].w, kL2E, -cl2)));
      sl = acc;
      ml = cm;
    }
  }
  const float m = warp_max(ml);
  const float ml2 = m * kL2E;
  float ssum = warp_sum(ml > -INFINITY ? sl * ex2_approx(fmaf(ml, kL2E, -ml2)) : 0.f);
  const float nbl = ex2_approx(fmaf(__ldg(x + blank), kL2E, -ml2));
  const float inv_nb = __fdividef(1.0f, nbl);
  bool bad = !(nbl >= 1.0e-38f && ssum <= 3.0e38f);
#pragma unroll
  for (int k = 0; k < NL; k++) {
    uint32_t w = 0u;
    if (pcls[k] >= 0) {
      const float rr = ex2_approx(fmaf(__ldg(x + pcls[k]), kL2E, -ml2)) * inv_nb;
      bad |= !(rr >= 1.1754944e-38f && rr <= 1.0e38f);   // only the transcript's classes have to stay normal
      w = ((__float_as_uint(rr) + 4u) >> 3) + (896u << 20);
    }
    rec[k * 32 + lane] = w;
  }
  if (bad) alarm |= AL_EMISSION;
  if (grow) {
    const float sc = gs * __fdividef(1.0f, ssum);
    if (edge) __stcg(grow + ec, ex2_approx(fmaf(xe, kL2E, -ml2)) * sc);
    float4* gv = reinterpret_cast<float4*>(grow + head);
    for (int k0 = 0; k0 < nv; k0 += 32 * kHV) {
      float4 v[kHV];
#pragma unroll
      for (int i = 0; i < kHV; i++) {
        const int k = k0 + lane + 32 * i;
        v[i] = k < nv ? __ldg(xv + k) : ninf4;
      }
#pragma unroll
      for (int i = 0; i < kHV; i++) {
        const int k = k0 + lane + 32 * i;
        if (k < nv)
          __stcg(gv + k,
                 make_float4(ex2_approx(fmaf(v[i].x, kL2E, -ml2)) * sc, ex2_approx(fmaf(v[i].y, kL2E, -ml2)) * sc,
                             ex2_approx(fmaf(v[i].z, kL2E, -ml2)) * sc, ex2_approx(fmaf(v[i].w, kL2E, -ml2)) * sc));
      }
    }
  }
  __syncwarp();
  return __logf(nbl) - __logf(ssum);
}

// pull a row towards L2 (one 128-byte line per lane and step)
__device__ __forceinline__ void l2_prefetch_row(const float* x, int C, int lane) {
  const char* pch = reinterpret_cast<const char*>(x);
  for (int o = lane * 128; o < C * 4; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pch + o));
}

template <int NL>
__device__ __forceinline__ uint32_t* ckpt_ptr(const Params& p, int b, int d, int c) {
  return p.ckpt + (((size_t)b * 2 + d) * p.maxch + c) * (size_t)((2 * NL + 1) * 32);
}

// physical cell of class-sorted position pos in a posterior row: [pos % NL][pos / NL], so that the gradient
// warp's lane l finds its NL consecutive positions l*NL .. l*NL+NL-1 at stride 32 (no bank conflicts)
template <int NL>
__device__ __forceinline__ uint32_t gcell(int pos) {
  return (uint32_t)((pos % NL) * 32 + pos / NL);
}

// How many CTAs of this kernel currently sit on each SM (incremented on arrival, decremented on exit, so it
// is zero between launches).  A CTA uses its arrival rank to pick its warp->role permutation: the block
// scheduler's pairing of CTAs on an SM is not a function of blockIdx (measured: only 30 of 108 pairs are
// (b, b+148)), and two CTAs with the same permutation put all four recursion warps of phase 1 on two of the
// four schedulers.
__device__ int g_sm_arrivals[1024];

// EPL = 0 instantiates the wide-vocabulary variant (64 < C <= 8192): 16 warps, one CTA per SM.
// STREAM (wide variant only): rows wider than kWideRegC, streamed by the producers instead of held in registers.
template <int NL, int EPL, bool STREAM = false>
__global__ void __launch_bounds__((EPL ? NTHREADS : NTHREADS_WIDE), ((EPL && NL <= 8) ? 2 : 1))
ctc_fast_kernel(const Params p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr bool WIDE = EPL == 0;
  constexpr int NT = WIDE ? NTHREADS_WIDE : NTHREADS;
  constexpr int CMAX = WIDE ? kWideMaxC : 8 * EPL;
  constexpr int ROWB = WIDE ? NL * 32 * 4 : row_bytes(CMAX);
  constexpr int YOFF = row_yoff(CMAX);
  const Smem sl = smem_layout(NL, CMAX, WIDE ? p.C : 0, STREAM);
  unsigned char* s_rows = smem + sl.rows;
  float* s_raw = reinterpret_cast<float*>(smem + sl.raw);
  float* s_rowbuf = reinterpret_cast<float*>(smem + sl.rowbuf);
  float* s_stage = reinterpret_cast<float*>(smem + sl.stage);
  int* s_dtab = reinterpret_cast<int*>(smem + sl.dtab);
  uint32_t* s_obuf = reinterpret_cast<uint32_t*>(smem + sl.obuf);
  float* s_gbuf = reinterpret_cast<float*>(smem + sl.gbuf);
  int* s_oexp = reinterpret_cast<int*>(smem + sl.oexp);
  double* s_meet_nb = reinterpret_cast<double*>(smem + sl.meet_nb);
  double* s_meet_pre = reinterpret_cast<double*>(smem + sl.meet_pre);
  int* s_meet_e = reinterpret_cast<int*>(smem + sl.meet_e);
  int* s_lab = reinterpret_cast<int*>(smem + sl.lab);
  uint16_t* s_pos = reinterpret_cast<uint16_t*>(smem + sl.pos);
  int* s_cls_off = reinterpret_cast<int*>(smem + sl.cls_off);
  double* s_psum = reinterpret_cast<double*>(smem + sl.psum);
  int* s_scal = reinterpret_cast<int*>(smem + sl.scal);
  double* s_inv_mp = reinterpret_cast<double*>(smem + sl.scal + 16);

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = p.T, C = p.C, blank = p.blank;
  constexpr int rowbytes = ROWB;
  constexpr int gstride = NL * 32 + 4;
  constexpr int N = NL * 32;
  constexpr int Lcap = N - 2;
  constexpr int OBUF = KC * NL * 32;  // words per buffer of high words
  constexpr int GBUF = KC * gstride;  // floats per posterior buffer

  const int Tb = p.seq_len[b];
  const int l0 = p.lab_offs[b];
  const int L = p.lab_offs[b + 1] - l0;

  // ---- can this kernel take the utterance? everything unusual goes to the robust kernel -----------
  int bad = (Tb < 2 * KC) | (Tb > T) | (L < 0) | (L > Lcap) | (C > CMAX) |
            (WIDE && !STREAM && ((C & 3) || C > kWideRegC));  // (the launcher also checks the rows' alignment)
  if (tid < 16) {
    if (tid < 8) s_scal[tid] = 0;
    s_psum[tid] = 0.0;
  }
  for (int c = tid; c < C + 2; c += NT) s_cls_off[c] = 0;
  __syncthreads();
  int rep = 0;
  if (!bad) {
    for (int i = tid; i < L; i += NT) {
      const int v = p.lab_vals[l0 + i];
      s_lab[i] = v;
      if (v < 0 || v >= C || v == blank) {
        bad = 1;
      } else {
        atomicAdd(&s_cls_off[v + 1], 1);
        if (i > 0 && v == p.lab_vals[l0 + i - 1]) rep++;
      }
    }
  }
  if (rep) atomicAdd(&s_scal[2], rep);
  bad = __syncthreads_or(bad);
  if (!bad && Tb < L + s_scal[2]) bad = 1;
  if (bad) {
    if (tid == 0) p.retry[b] = AL_SHAPE;
    return;
  }

  // ---- class-sorted rank of every label (the order the gradient warps reduce in) -----------------
  if (warp == 0) {
    const int per = (C + 1 + 31) / 32;
    const int lo = min(C + 1, lane * per), hi = min(C + 1, lo + per);
    int s = 0;
    for (int i = lo; i < hi; i++) s += s_cls_off[i];
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    int run = incl - s;
    for (int i = lo; i < hi; i++) {
      run += s_cls_off[i];
      s_cls_off[i] = run;
    }
  }
  __syncthreads();
  // s_cls_off[c] = number of labels with class < c
  for (int j = tid; j < L; j += NT) {
    const int v = s_lab[j];
    int r = 0;
    for (int i = 0; i < j; i++) r += (s_lab[i] == v);
    s_pos[j] = (uint16_t)(s_cls_off[v] + r);
  }
  if (!WIDE) {
    // raw rows: the columns past C stay -inf for the whole kernel; row records: the zero entry of dead slots
    for (int i = tid; i < 2 * 4 * KC * CMAX; i += NT)
      if (i % CMAX >= C) s_raw[i] = -INFINITY;
    for (int i = tid; i < 2 * 4 * KC; i += NT)
      *reinterpret_cast<uint32_t*>(s_rows + (size_t)i * rowbytes + CMAX * 4) = 0u;
  } else {
    // distinct classes of the transcript, with the posterior-row cells that hold the inclusive prefix just
    // before and at the end of each class: [0,Lcap) class, [Lcap,2Lcap) cell "lo", [2Lcap,3Lcap) cell "hi"
    for (int c = tid; c < C; c += NT) {
      const int lo = s_cls_off[c], hi = s_cls_off[c + 1];
      if (hi > lo && c != blank) {
        const int u = atomicAdd(&s_scal[6], 1);
        s_dtab[u] = c;
        s_dtab[(Lcap + 2) + u] = lo > 0 ? (int)gcell<NL>(lo - 1) : N;
        s_dtab[2 * (Lcap + 2) + u] = (int)gcell<NL>(hi - 1);
      }
    }
  }
  // the zero cell of every posterior row (prefix "before position 0")
  for (int i = tid; i < 2 * 2 * KC; i += NT) {
    float* row = s_gbuf + (size_t)i * gstride;
    row[N] = 0.f;
  }

  // ---- gradient rows of padded frames are exactly zero -----------------------------------------
  const size_t rstride = (size_t)p.st_t;
  float* gbase = p.grad ? p.grad + (size_t)b * p.st_b : nullptr;
  if (gbase) {
    const size_t n = (size_t)(T - Tb) * C;
    for (size_t i = tid; i < n; i += NT) {
      const size_t t = Tb + i / C;
      gbase[t * rstride + (i % C)] = 0.f;
    }
  }

  // ---- schedule -----------------------------------------------------------------------------
  Sched S;
  S.Tb = Tb;
  {
    int nf = KC * ((Tb + 2 * KC - 1) / (2 * KC));
    if (p.split > 0) nf = min(max(KC, (p.split / KC) * KC), ((Tb - 1) / KC) * KC);
    S.n1F = nf;
    S.n1B = Tb - nf;
    S.nch1F = (S.n1F + KC - 1) / KC;
    S.nch1B = (S.n1B + KC - 1) / KC;
    S.P1 = max(S.nch1F, S.nch1B);
    S.offB = S.P1 - S.nch1B;
    S.last = gbase ? S.P1 + 1 + S.P1 : S.P1;
  }
  const bool want_grad = gbase != nullptr;
  const float gs = p.grad_loss ? p.grad_loss[b] : 1.0f;

  // roles: the second CTA that lands on an SM swaps recursion and recompute warps (and producer and
  // gradient warps) so that the four heavy warps of phase 1 sit on four different schedulers
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (tid == 0) s_scal[3] = atomicAdd(&g_sm_arrivals[smid & 1023], 1);
  __syncthreads();
  const int perm = WIDE ? 0 : (s_scal[3] & 1);
  // wide: warps 0-3 recursion / recompute, 4-7 gradient (4,5 frames 0-3 of a chunk, 6,7 frames 4-7),
  // 8-15 producers; side = parity of the warp index
  int role = WIDE ? (warp < 4 ? warp : (warp < 8 ? G_F + (warp & 1) : P_F + (warp & 1))) : (warp ^ (perm << 1));
  if (!WIDE) {
    // A warp issues from scheduler (hardware warp slot mod 4), and the slots of an SM's second CTA need not start at
    // a multiple of four (measured with four-warp CTAs: 5, 6, 7, 4).  Deal the roles by the scheduler actually got:
    // the first warp of the CTA on a scheduler takes recursion / recompute, the second producer / gradient, the second
    // CTA the other way round, so that every scheduler of the SM carries one warp of each kind.
    __shared__ int s_hw[8];
    unsigned hwid;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(hwid));
    hwid = __shfl_sync(0xffffffffu, hwid, 0);
    if (lane == 0) s_hw[warp] = (int)(hwid & 3u);
    __syncthreads();
    int cnt = 0, mine = 0;   // four 4-bit counters: warps of this CTA per scheduler
#pragma unroll
    for (int v = 0; v < 8; v++) {
      const int q = s_hw[v];
      if (v == warp) mine = (cnt >> (4 * q)) & 15;
      cnt += 1 << (4 * q);
    }
    // (which of a scheduler's two warps takes the heavy role makes no difference: measured 0.2546 / 0.2536 ms)
    if (cnt == 0x2222) role = ((int)(hwid & 3u) ^ (perm << 1)) + 4 * mine;
  }
  const int d = role & 1;  // direction / side this warp works for
  __syncthreads();

  Dir<NL> st;
  uint32_t gphys[NL];  // recursion warps: byte offset of slot k's posterior inside a posterior row
  if (role <= RC_B) {
    dir_setup<NL>(st, d, lane, s_lab, L, CMAX, WIDE ? (role <= H_B ? 1 : 2) : 0);
    const int pad = N - L - 1;
#pragma unroll
    for (int k = 0; k < NL; k++) {
      const int i = lane * NL + k;
      int pos;
      if (d == 0) {
        const int j = i - 1;
        pos = (j >= 0 && j < L) ? (int)s_pos[j] : (i == 0 ? L : i);
      } else {
        const int m = i - pad;
        pos = (m >= 0 && m < L) ? (int)s_pos[L - 1 - m] : (i < pad ? L + i : N - 1);
      }
      gphys[k] = gcell<NL>(pos) * 4u;
    }
  }
  // gradient warps: cells holding the inclusive prefix at the end of each of this lane's two classes
  uint32_t pc_lo0 = N * 4u, pc_hi0 = N * 4u, pc_lo1 = N * 4u, pc_hi1 = N * 4u, pc_tot = N * 4u;
  if (role >= G_F) {
    const int c0 = lane, c1 = lane + 32;
    if (c0 < C && c0 != blank) {
      const int lo = s_cls_off[c0], hi = s_cls_off[c0 + 1];
      if (lo > 0) pc_lo0 = gcell<NL>(lo - 1) * 4u;
      if (hi > 0) pc_hi0 = gcell<NL>(hi - 1) * 4u;
    }
    if (c1 < C && c1 != blank) {
      const int lo = s_cls_off[c1], hi = s_cls_off[c1 + 1];
      if (lo > 0) pc_lo1 = gcell<NL>(lo - 1) * 4u;
      if (hi > 0) pc_hi1 = gcell<NL>(hi - 1) * 4u;
    }
    if (L > 0) pc_tot = gcell<NL>(L - 1) * 4u;
  }
  // wide gradient warps: this lane's share of the distinct classes of the transcript (at most NL of them)
  constexpr int WD = WIDE ? NL : 1;
  int wd_cls[WD], wd_lo[WD], wd_hi[WD];
  if (WIDE && role >= G_F) {
    const int U = s_scal[6];
#pragma unroll
    for (int i = 0; i < WD; i++) {
      const int u = lane + 32 * i;
      wd_cls[i] = u < U ? s_dtab[u] : -1;
      wd_lo[i] = u < U ? s_dtab[N + u] : N;
      wd_hi[i] = u < U ? s_dtab[2 * N + u] : N;
    }
  }
  int alarm = 0;
  bool scaled = false;     // recursion warps: state already divided by the mantissa of p

  long long prof_w1 = 0, prof_w2 = 0;
  const long long prof_t0 = clock64();
  // Each role runs its own copy of the iteration loop (same trip count, one CTA barrier per iteration) so
  // that a warp keeps only its own role's state in registers.
#define NASR_PROF_BEGIN() const long long prof_a = NASR_PROF_PTR(p) ? clock64() : 0
#define NASR_PROF_END()                                    \
  if (NASR_PROF_PTR(p)) {                                            \
    const long long pe = clock64();                        \
    const long long dt = pe - prof_a;                      \
    if (I < S.P1) prof_w1 += dt; else prof_w2 += dt;       \
    if (b < 4 && lane == 0 && warp < 8 && I - IFIRST < 200) {  \
      long long* tr = NASR_PROF_PTR(p) + (size_t)p.B * 64 + (((size_t)b * 200 + (I - IFIRST)) * 8 + warp) * 2; \
      tr[0] = prof_a - prof_t0;                            \
      tr[1] = pe - prof_t0;                                \
    }                                                      \
  }
  if (role == H_F || role == H_B) {
#pragma unroll 1
    for (int I = IFIRST; I <= S.last; I++) {
      NASR_PROF_BEGIN();
      // ======================= recursion warps =======================
      const Chunk ci = chunk_at(S, d, I);
      const unsigned char* erows = s_rows + (size_t)(d * 4 + (I & 3)) * KC * rowbytes;
      if (ci.phase == 1) {
        rescale<NL>(st, lane, alarm);
        uint32_t* ck = ckpt_ptr<NL>(p, b, d, ci.idx);
#pragma unroll
        for (int k = 0; k < NL; k++) {
          ck[k * 32 + lane] = (uint32_t)__double2hiint(st.Ab[k]);
          ck[(NL + k) * 32 + lane] = (uint32_t)__double2hiint(st.Al[k]);
        }
        ck[2 * NL * 32 + lane] = (uint32_t)st.E;
        const double fin = inflow_factor(st.E, lane);
        run_chunk<NL, PLAIN, ROWB, (WIDE ? 2 : KC)>(st, erows, ci.len, false, s_obuf, s_gbuf, fin, 0, gphys, lane);
        if (d == 1 && I == S.P1 - 1) {
          // pre-emission sums of the frame below the meeting point, for the forward warp
          rescale<NL>(st, lane, alarm);
          const double fin2 = inflow_factor(st.E, lane);
          double a_in = __shfl_up_sync(0xffffffffu, st.Al[NL - 1], 1);
          a_in = lane ? a_in * fin2 : 0.0;
#pragma unroll
          for (int k = NL - 1; k >= 0; k--) {
            const double alp = k > 0 ? st.Al[k - 1] : a_in;
            const double nb = st.Ab[k] + alp;
            const double pre = fma(st.skip[k], alp, st.Al[k] + st.Ab[k]);
            s_meet_nb[k * 32 + lane] = nb;
            s_meet_pre[k * 32 + lane] = pre;
          }
          s_meet_e[lane] = st.E;
        }
      } else if (I == S.P1 && d == 0) {
        // ---- meeting: p = sum over states of alpha(M-1) * beta(M-1) ----
        rescale<NL>(st, lane, alarm);
        const int lm = 31 - lane;
        const int Eb1 = s_meet_e[lm];
        double S1 = 0.0;
#pragma unroll
        for (int k = 0; k < NL; k++) S1 += st.Al[k] * s_meet_pre[(NL - 1 - k) * 32 + lm];
#pragma unroll
        for (int k = 1; k < NL; k++) S1 += st.Ab[k] * s_meet_nb[(NL - k) * 32 + lm];
        double S2 = 0.0;
        int Eb2 = ENEG;
        if (lane >= 1) {
          S2 = st.Ab[0] * s_meet_nb[32 - lane];
          Eb2 = s_meet_e[32 - lane];
        }
        const bool ok1 = S1 > 0.0 && st.E != ENEG && Eb1 != ENEG;
        const bool ok2 = S2 > 0.0 && st.E != ENEG && Eb2 != ENEG;
        const int K1 = st.E + Eb1, K2 = st.E + Eb2;
        const int X1 = ok1 ? K1 + (((__double2hiint(S1) >> 20) & 0x7ff) - 1023) : INT_MIN / 2;
        const int X2 = ok2 ? K2 + (((__double2hiint(S2) >> 20) & 0x7ff) - 1023) : INT_MIN / 2;
        int Xm = max(X1, X2);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) Xm = max(Xm, __shfl_xor_sync(0xffffffffu, Xm, o));
        double tot = 0.0;
        if (ok1) tot += S1 * pow2d(K1 - Xm);
        if (ok2) tot += S2 * pow2d(K2 - Xm);
        tot = warp_sum(tot);
        if (!(tot > 0.0) || !(tot < 1e300) || Xm < -(1 << 28)) {
          alarm |= AL_P;
          tot = 1.0;
          Xm = 0;
        }
        const int et = ((__double2hiint(tot) >> 20) & 0x7ff) - 1023;
        const double mp = tot * pow2d(-et);
        double ls = 0.0;
#pragma unroll
        for (int i = 0; i < 16; i++) ls += s_psum[i];
        if (lane == 0) {
          s_scal[1] = Xm + et;
          *s_inv_mp = 1.0 / mp;
          p.loss[b] = (float)(-((double)Xm * 0.6931471805599453 + log(tot) + ls));
          p.status[b] = 0;
        }
      } else if (ci.phase == 2 && want_grad) {
        if (!scaled) {
          // the recursion is linear: dividing the state by the mantissa of p once makes every later
          // pre-emission sum carry the 1/m_p factor of the posterior
          const double im = *s_inv_mp;
#pragma unroll
          for (int k = 0; k < NL; k++) {
            st.Ab[k] *= im;
            st.Al[k] *= im;
          }
          scaled = true;
        }
        rescale<NL>(st, lane, alarm);
        const double fin = inflow_factor(st.E, lane);
        const int buf = I & 1;
        const int Eo = s_oexp[(d * 2 + buf) * 32 + (31 - lane)];
        int ks = -2047;
        if (Eo != ENEG && st.E != ENEG) {
          ks = st.E + Eo - s_scal[1];
          if (ks > ZALARM) alarm |= AL_RANGE;
          ks = max(-2047, min(ks, 600));
        }
        run_chunk<NL, COMBINE, ROWB, (WIDE ? 2 : KC)>(st, erows, ci.len, false, s_obuf + (size_t)(d * 2 + buf) * OBUF,
                                     s_gbuf + (size_t)(d * 2 + buf) * GBUF, fin, ks * (1 << 20), gphys, lane);
      }
      NASR_PROF_END();
      cta_sync();
    }
  } else if (role == RC_F || role == RC_B) {
#pragma unroll 1
    for (int I = IFIRST; I <= S.last; I++) {
      NASR_PROF_BEGIN();
      // ======================= recompute warps =======================
      // this warp computes direction d's rows; they are consumed by the other direction (side d^1)
      const int side = d ^ 1;
      const Chunk ci = chunk_at(S, side, I + 1);
      if (ci.phase == 2 && want_grad && !(NASR_ABLATE(p) & 4)) {
        const uint32_t* ck = ckpt_ptr<NL>(p, b, d, ci.idx);
#pragma unroll
        for (int k = 0; k < NL; k++) {
          st.Ab[k] = hi2d(__ldcg(ck + k * 32 + lane));
          st.Al[k] = hi2d(__ldcg(ck + (NL + k) * 32 + lane));
        }
        st.E = (int)__ldcg(ck + 2 * NL * 32 + lane);
        const double fin = inflow_factor(st.E, lane);
        const int buf = (I + 1) & 1;
        s_oexp[(side * 2 + buf) * 32 + lane] = st.E;
        const unsigned char* erows = s_rows + (size_t)(side * 4 + ((I + 1) & 3)) * KC * rowbytes;
        run_chunk<NL, STORE_O, ROWB, (WIDE ? 2 : KC)>(st, erows, ci.len, true, s_obuf + (size_t)(side * 2 + buf) * OBUF, s_gbuf, fin, 0,
                                     gphys, lane);
      }
      NASR_PROF_END();
      cta_sync();
    }
  } else if (role == P_F || role == P_B) {
    if constexpr (WIDE) {
      // ======================= producer warps, wide vocabulary =======================
      // Warp j of a side converts rows j and j+4 of every chunk of that side, two iterations before the
      // recursion needs it; the row that follows is requested while the current one is processed, and the
      // rows of two chunks later are prefetched into L2.
      const int pj = (warp - 8) >> 1;
      float* rowbuf = s_rowbuf + (size_t)(warp - 8) * (STREAM ? 0 : C);
      int pcls[NL];
      {
        const int pad = N - L - 1;
#pragma unroll
        for (int k = 0; k < NL; k++) {
          const int i = lane * NL + k;
          if (d == 0) {
            pcls[k] = (i >= 1 && i <= L) ? s_lab[i - 1] : -1;
          } else {
            const int m = i - pad;
            pcls[k] = (m >= 0 && m < L) ? s_lab[L - 1 - m] : -1;
          }
        }
      }
      double lsum = 0.0;
      float* stA = s_stage + (size_t)(warp - 8) * 2 * (STREAM ? 0 : C);   // staging rows of this warp: row pj and row pj+4
      float* stB = stA + (STREAM ? 0 : C);
      auto row_ptr = [&](const Chunk& ci, int f) {
        const int ff = min(f, ci.len - 1);
        const int t = d ? ci.base - ff : ci.base + ff;
        return p.logits + (size_t)t * p.st_t + (size_t)b * p.st_b;
      };
      auto wanted = [&](const Chunk& ci) { return ci.phase != 0 && (ci.phase == 1 || want_grad); };
      constexpr bool streamed = STREAM;  // rows too wide for registers: streamed from global / L2 (huge_row)
      if (!streamed) {
        const Chunk c0 = chunk_at(S, d, IFIRST + 2);
        if (wanted(c0)) {
          wide_issue(stA, row_ptr(c0, pj), C, lane);
          wide_issue(stB, row_ptr(c0, pj + 4), C, lane);
        }
        cp_async_commit();
      }
#pragma unroll 1
      for (int I = IFIRST; I <= S.last; I++) {
        NASR_PROF_BEGIN();
        const Chunk ci = chunk_at(S, d, I + 2);
        const Chunk cn = chunk_at(S, d, I + 3);
        const bool on = wanted(ci), onn = wanted(cn);
        uint32_t* recs = reinterpret_cast<uint32_t*>(s_rows + (size_t)(d * 4 + ((I + 2) & 3)) * KC * rowbytes);
        if (streamed) {
          if (onn) {  // next iteration's two rows on their way to L2 while this iteration's are processed
            l2_prefetch_row(row_ptr(cn, pj), C, lane);
            l2_prefetch_row(row_ptr(cn, pj + 4), C, lane);
          }
#pragma unroll 1
          for (int h = 0; h < 2; h++) {
            const int f = pj + 4 * h;
            if (on && f < ci.len) {
              const int t = d ? ci.base - f : ci.base + f;
              float* grow = (ci.phase == 2) ? gbase + (size_t)t * rstride : nullptr;
              const float ly = huge_row<NL>(row_ptr(ci, f), recs + f * N, pcls, grow, gs, C, blank, lane, alarm);
              if (ci.phase == 1) lsum += (double)ly;
            }
          }
          if (lane == 0) s_psum[warp - 8] = lsum;
          NASR_PROF_END();
          cta_sync();
          continue;
        }
        cp_async_wait<0>();   // the rows of this iteration were requested one iteration ago
        __syncwarp();
        WideRow r;
        if (on) wide_fetch(r, stA, C, lane);
        __syncwarp();
        if (onn) wide_issue(stA, row_ptr(cn, pj), C, lane);
        if (on && pj < ci.len) {
          const int t = d ? ci.base - pj : ci.base + pj;
          float* grow = (ci.phase == 2) ? gbase + (size_t)t * rstride : nullptr;
          const float ly = wide_row<NL>(r, rowbuf, recs + pj * N, pcls, grow, gs, C, blank, lane, alarm);
          if (ci.phase == 1) lsum += (double)ly;
        }
        if (on) wide_fetch(r, stB, C, lane);
        __syncwarp();
        if (onn) wide_issue(stB, row_ptr(cn, pj + 4), C, lane);
        cp_async_commit();
        if (on && pj + 4 < ci.len) {
          const int t = d ? ci.base - (pj + 4) : ci.base + (pj + 4);
          float* grow = (ci.phase == 2) ? gbase + (size_t)t * rstride : nullptr;
          const float ly = wide_row<NL>(r, rowbuf, recs + (pj + 4) * N, pcls, grow, gs, C, blank, lane, alarm);
          if (ci.phase == 1) lsum += (double)ly;
        }
        if (lane == 0) s_psum[warp - 8] = lsum;
        NASR_PROF_END();
        cta_sync();
      }
    } else {
    // ======================= producer warps =======================
    // Raw logits rows travel global -> shared memory by cp.async, issued five iterations before the recursion
    // needs the chunk (an iteration is about as long as a DRAM round trip), and are turned into row records
    // three iterations later.  In phase 1 the gradient warps have nothing to reduce and convert rows 4..7.
    double lsum = 0.0;        // sum of log y_blank over the phase-1 rows this lane group owned
    const int rl = lane >> 3, sub = lane & 7;
#pragma unroll 1
    for (int I = IFIRST; I <= S.last; I++) {
      NASR_PROF_BEGIN();
      {  // raw rows of the chunk of iteration I+2 (complete and visible since the last barrier) -> row records
        const Chunk ci = chunk_at(S, d, I + 2);
        if (ci.phase != 0 && (ci.phase == 1 || (want_grad && !(NASR_ABLATE(p) & 2)))) {
          unsigned char* rows = s_rows + (size_t)(d * 4 + ((I + 2) & 3)) * KC * rowbytes;
          const float* raw = s_raw + (size_t)(d * 4 + ((I + 2) & 3)) * KC * CMAX;
          const float l0v = rows_to_smem<EPL, ROWB, YOFF>(raw, rl, ci.len, rows, C, blank, s_scal);
          if (ci.phase == 1) {
            lsum += (double)l0v;
          } else {
            rows_to_smem<EPL, ROWB, YOFF>(raw, 4 + rl, ci.len, rows, C, blank, s_scal);
          }
          if (sub == (blank & 7)) s_psum[d * 4 + rl] = lsum;
        }
      }
      {  // issue the copies for the chunk of iteration I+5 (one commit group per iteration, empty or not)
        const Chunk ci = chunk_at(S, d, I + 5);
        if (ci.phase != 0 && (ci.phase == 1 || (want_grad && !(NASR_ABLATE(p) & 2)))) {
          float* raw = s_raw + (size_t)(d * 4 + ((I + 5) & 3)) * KC * CMAX;
#pragma unroll
          for (int g = 0; g < KC / 4; g++) {
            const int f = g * 4 + rl;
            const int ff = min(f, ci.len - 1);
            const int t = d ? ci.base - ff : ci.base + ff;
            const float* xrow = p.logits + (size_t)t * p.st_t + (size_t)b * p.st_b;
#pragma unroll
            for (int e = 0; e < EPL; e++)
              if (sub + 8 * e < C) cp_async4(raw + f * CMAX + sub + 8 * e, xrow + sub + 8 * e);
          }
        }
        cp_async_commit();
        cp_async_wait<2>();  // everything up to the chunk of iteration I+3 has landed; the barrier publishes it
      }
      NASR_PROF_END();
      cta_sync();
    }
    }
  } else {
    double lsumg = 0.0;
    const int rl = lane >> 3;
#pragma unroll 1
    for (int I = IFIRST; I <= S.last; I++) {
      NASR_PROF_BEGIN();
      if constexpr (!WIDE) {  // phase 1: second half of the producer's job
        const Chunk c2 = chunk_at(S, d, I + 2);
        if (c2.phase == 1) {
          unsigned char* rows = s_rows + (size_t)(d * 4 + ((I + 2) & 3)) * KC * rowbytes;
          const float* raw = s_raw + (size_t)(d * 4 + ((I + 2) & 3)) * KC * CMAX;
          lsumg += (double)rows_to_smem<EPL, ROWB, YOFF>(raw, 4 + rl, c2.len, rows, C, blank, s_scal);
          if ((lane & 7) == (blank & 7)) s_psum[8 + d * 4 + rl] = lsumg;
        }
      }
      // ======================= gradient warps =======================
      // posterior rows are in class-sorted order: an inclusive prefix sum turns "occupancy of class c" into
      // the difference of two prefixes
      const Chunk ci = chunk_at(S, d, I - 1);
      if (ci.phase == 2 && want_grad && !(NASR_ABLATE(p) & 1)) {
        const int buf = (I - 1) & 1;
        float* G = s_gbuf + (size_t)(d * 2 + buf) * GBUF;
        const unsigned char* rows = s_rows + (size_t)(d * 4 + ((I - 1) & 3)) * KC * rowbytes;
        const int c0 = lane, c1 = lane + 32;
        constexpr int NF = 4;  // frames in flight (8 was measured: the extra shared-memory traffic in flight slows the recursion warps): their dependency chains interleave
        // wide: two gradient warps per side, each takes one half of the chunk's frames
        const int f_first = WIDE ? ((warp >> 1) & 1) * NF : 0;
        const int f_step = WIDE ? 2 * NF : NF;
#pragma unroll 1
        for (int f0 = f_first; f0 < ci.len; f0 += f_step) {
          float v[NF][NL];
#pragma unroll
          for (int j = 0; j < NF; j++)
#pragma unroll
            for (int i = 0; i < NL; i++) v[j][i] = G[(f0 + j) * gstride + i * 32 + lane];
          float inc[NF], base[NF];
#pragma unroll
          for (int j = 0; j < NF; j++) {
#pragma unroll
            for (int i = 1; i < NL; i++) v[j][i] += v[j][i - 1];
            inc[j] = v[j][NL - 1];
          }
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int j = 0; j < NF; j++) {
              const float u = __shfl_up_sync(0xffffffffu, inc[j], o);
              if (lane >= o) inc[j] += u;
            }
          }
#pragma unroll
          for (int j = 0; j < NF; j++) {
            base[j] = inc[j] - v[j][NL - 1];
#pragma unroll
            for (int i = 0; i < NL; i++) G[(f0 + j) * gstride + i * 32 + lane] = v[j][i] + base[j];
          }
          __syncwarp();
          if constexpr (WIDE) {
            // the producers already wrote grad_loss * softmax for these frames; subtract the occupancy of
            // every distinct class of the transcript (and of the blank) in place
#pragma unroll
            for (int j = 0; j < NF; j++) {
              const int f = f0 + j;
              if (f < ci.len) {
                const int t = d ? ci.base - f : ci.base + f;
                float* g = gbase + (size_t)t * rstride;
                const float* Gr = G + (f0 + j) * gstride;
                // one contribution per cell, so the order-free reduction in L2 is deterministic; red
                // does not wait for a round trip the way a load-modify-store would
#pragma unroll
                for (int i = 0; i < WD; i++)
                  if (wd_cls[i] >= 0) atomicAdd(g + wd_cls[i], -(gs * (Gr[wd_hi[i]] - Gr[wd_lo[i]])));
                if (lane == 0) atomicAdd(g + blank, -(gs * (1.0f - Gr[pc_tot >> 2])));
              }
            }
            continue;
          }
          float o0[NF], o1[NF], tot[NF], y0[NF], y1[NF];
#pragma unroll
          for (int j = 0; j < NF; j++) {
            const unsigned char* Grb = reinterpret_cast<const unsigned char*>(G + (f0 + j) * gstride);
            const float* yrow = reinterpret_cast<const float*>(rows + (size_t)(f0 + j) * rowbytes + YOFF);
            o0[j] = *reinterpret_cast<const float*>(Grb + pc_hi0) - *reinterpret_cast<const float*>(Grb + pc_lo0);
            o1[j] = *reinterpret_cast<const float*>(Grb + pc_hi1) - *reinterpret_cast<const float*>(Grb + pc_lo1);
            tot[j] = *reinterpret_cast<const float*>(Grb + pc_tot);
            y0[j] = yrow[min(c0, C - 1)];
            y1[j] = yrow[min(c1, C - 1)];
          }
#pragma unroll
          for (int j = 0; j < NF; j++) {
            const int f = f0 + j;
            if (f < ci.len) {
              const int t = d ? ci.base - f : ci.base + f;
              float* g = gbase + (size_t)t * rstride;
              if (c0 < C) g[c0] = gs * (y0[j] - (c0 == blank ? 1.0f - tot[j] : o0[j]));
              if (c1 < C) g[c1] = gs * (y1[j] - (c1 == blank ? 1.0f - tot[j] : o1[j]));
            }
          }
        }
      }
      NASR_PROF_END();
      cta_sync();
    }
  }
#undef NASR_PROF_BEGIN
#undef NASR_PROF_END
  if (NASR_PROF_PTR(p) && lane == 0) {
    long long* q = NASR_PROF_PTR(p) + ((size_t)b * 16 + warp) * 4;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    q[0] = prof_w1; q[1] = prof_w2; q[2] = clock64() - prof_t0; q[3] = role | ((long long)smid << 8);
  }
  if (alarm) atomicOr(&s_scal[0], alarm);
  __syncthreads();
  if (tid == 0) {
    p.retry[b] = s_scal[0];
    atomicSub(&g_sm_arrivals[smid & 1023], 1);
  }
}

}  // namespace fast

// ---- host side -------------------------------------------------------------------------------

namespace {

constexpr int kMaxSmem = 226 * 1024;  // of the 227 KB a CTA may have on sm_100

int pick_nl(int Lmax) {
  static const int kNL[] = {2, 4, 5, 7, 10, 20};
  for (int nl : kNL)
    if (Lmax <= nl * 32 - 2) return nl;
  return 0;
}

bool is_wide(int C) { return C > 64 && C <= fast::kWideMaxC; }
// shapes only the streaming producers take; aligned narrower rows are held in registers
bool stream_only(int C) { return C > fast::kWideRegC || (C & 3) != 0; }

// the wide-vocabulary variant is instantiated for these slot counts only (L <= 126, 158, 222; 318 with streamed
// rows, whose producers need no row buffers)
int pick_nl_wide(int Lmax) {
  static const int kNL[] = {4, 5, 7, 10};
  for (int nl : kNL)
    if (Lmax <= nl * 32 - 2) return nl;
  return 0;
}

template <int NL, int EPL, bool STREAM = false>
int launch_fast(const fast::Params& p, cudaStream_t stream) {
  const fast::Smem sl = fast::smem_layout(NL, 8 * EPL, EPL ? 0 : p.C, STREAM);
  NASR_CUDA((ensure_max_dynamic_smem<fast::ctc_fast_kernel<NL, EPL, STREAM>>(kMaxSmem)));
  if (sl.total > (size_t)kMaxSmem) {
    set_error("nasr_ctc: fast kernel shared memory %zu too large", sl.total);
    return NASR_ERR_UNSUPPORTED;
  }
  fast::ctc_fast_kernel<NL, EPL, STREAM><<<p.B, EPL ? fast::NTHREADS : fast::NTHREADS_WIDE, sl.total, stream>>>(p);
  count_launch();
  NASR_CUDA(cudaGetLastError());
  return NASR_OK;
}

template <int NL>
int launch_fast_c(const fast::Params& p, cudaStream_t stream) {
  if (p.C <= 40) return launch_fast<NL, 5>(p, stream);
  return launch_fast<NL, 8>(p, stream);
}

}  // namespace

// implemented in ctc_narrow.cu: the float32 kernel for C <= 64 and transcripts up to 254 labels
bool ctc_narrow_supported(int T, int C, int Lmax);
bool ctc_narrow_layout_ok(const float* logits, long long st_t, long long st_b);
size_t ctc_narrow_workspace_bytes(int T, int B, int C, int Lmax);
int ctc_narrow_launch(const float* logits, int T, int B, int C, long long st_t, long long st_b,
                      const int32_t* label_values, const int32_t* label_offsets, int Lmax, const int32_t* seq_len,
                      int blank, float* loss, float* grad, const float* grad_loss, int32_t* status, int32_t* retry,
                      void* ckpt, cudaStream_t stream);
// Which kernel takes narrow vocabularies (C <= 64): the fp64 kernel of this file by default; the float32 kernel of
// ctc_narrow.cu where NASR_NARROW_F32=1 is set in the environment (read at every call, so tests can switch).  Measured on
// B200 (profiles/r2_*): 0.256 ms against 0.298 ms at cfg3, 0.137 against 0.143 at cfg2, 0.089 against 0.095 at cfg1.
static bool use_narrow_f32() {
  const char* e = getenv("NASR_NARROW_F32");
  return e && e[0] == '1';
}

int g_debug_split = 0;  // test hook (nasr_debug_config): frames of the forward half, 0 = automatic
int g_debug_ablate = 0;
long long* g_debug_prof = nullptr;  // test hook (nasr_debug_profile): device buffer for per-warp cycle counts

static int max_chunks(int T) {
  // chunks one direction can own in phase 1: half the frames normally, all of them under a split override
  const int half = (T + 2 * fast::KC - 1) / (2 * fast::KC) + 2;
  return g_debug_split ? (T + fast::KC - 1) / fast::KC + 2 : half;
}

bool ctc_fast_is_wide(int C) { return is_wide(C); }

bool ctc_fast_supported(int T, int C, int Lmax) {
  if (T < 2 * fast::KC) return false;
  if (is_wide(C)) {
    const int NL = pick_nl_wide(Lmax);
    return NL != 0 && fast::smem_layout(NL, 0, C, stream_only(C) || NL == 10).total <= (size_t)kMaxSmem;
  }
  const int NL = pick_nl(Lmax);
  if (C > 64 || NL == 0) return false;
  return fast::smem_layout(NL, C <= 40 ? 40 : 64).total <= (size_t)kMaxSmem;
}

size_t ctc_fast_workspace_bytes(int T, int B, int C, int Lmax) {
  const size_t nw = ctc_narrow_workspace_bytes(T, B, C, Lmax);
  if (!ctc_fast_supported(T, C, Lmax)) return nw;
  const int NL = is_wide(C) ? pick_nl_wide(Lmax) : pick_nl(Lmax);
  const size_t ow = (size_t)B * 2 * max_chunks(T) * (2 * NL + 1) * 32 * sizeof(uint32_t);
  return ow > nw ? ow : nw;
}

int ctc_fast_launch(const float* logits, int T, int B, int C, long long st_t, long long st_b,
                    const int32_t* label_values,
                    const int32_t* label_offsets, int Lmax, const int32_t* seq_len, int blank, float* loss,
                    float* grad, const float* grad_loss, int32_t* status, int32_t* retry, void* ckpt,
                    cudaStream_t stream) {
  if (use_narrow_f32() && ctc_narrow_supported(T, C, Lmax) && ctc_narrow_layout_ok(logits, st_t, st_b))
    return ctc_narrow_launch(logits, T, B, C, st_t, st_b, label_values, label_offsets, Lmax, seq_len, blank, loss, grad,
                             grad_loss, status, retry, ckpt, stream);
  fast::Params p;
  p.logits = logits; p.T = T; p.B = B; p.C = C; p.st_t = st_t; p.st_b = st_b;
  p.lab_vals = label_values; p.lab_offs = label_offsets; p.seq_len = seq_len;
  p.blank = blank; p.loss = loss; p.grad = grad; p.grad_loss = grad_loss; p.status = status;
  p.retry = retry;
  p.ckpt = static_cast<uint32_t*>(ckpt);
  p.maxch = max_chunks(T);
  p.split = g_debug_split;
  p.prof = g_debug_prof;
  p.ablate = g_debug_ablate;
  int num_sms = 0;
  NASR_CUDA(device_sm_count(&num_sms));
  p.num_sms = num_sms;
  if (is_wide(C)) {
    // rows the register-held path cannot take (too wide, not a multiple of four, not 16-byte aligned) are streamed
    const bool streamed = stream_only(C) || ((uintptr_t)logits & 15) != 0 || (st_t & 3) != 0 || (st_b & 3) != 0;
    switch (pick_nl_wide(Lmax)) {
      case 4: return streamed ? launch_fast<4, 0, true>(p, stream) : launch_fast<4, 0>(p, stream);
      case 5: return streamed ? launch_fast<5, 0, true>(p, stream) : launch_fast<5, 0>(p, stream);
      case 7: return streamed ? launch_fast<7, 0, true>(p, stream) : launch_fast<7, 0>(p, stream);
      case 10: return launch_fast<10, 0, true>(p, stream);  // ten slots per lane leave no room for row buffers
    }
  }
  switch (pick_nl(Lmax)) {
    case 2: return launch_fast_c<2>(p, stream);
    case 4: return launch_fast_c<4>(p, stream);
    case 5: return launch_fast_c<5>(p, stream);
    case 7: return launch_fast_c<7>(p, stream);
    case 10: return launch_fast_c<10>(p, stream);
    case 20: return launch_fast_c<20>(p, stream);
  }
  set_error("nasr_ctc: fast kernel does not support max_label_len=%d", Lmax);
  return NASR_ERR_UNSUPPORTED;
}

}  // namespace nasr
