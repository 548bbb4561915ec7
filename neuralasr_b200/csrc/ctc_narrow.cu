// CTC loss + d(loss)/d(logits) for narrow vocabularies (C <= 64, transcripts up to 254 labels) on sm_100a:
// one CTA of four specialised warps per utterance, one launch per batch, float32 arithmetic.
//
// Replaces tf.nn.ctc_loss + _CTCLossGrad behind create_loss (reference networks/tfnetwork.py:58-59); semantics per
// SURVEY.md Appendix A.1.  ctc_loss.cu holds the robust kernel that redoes any utterance flagged in retry[].
//
// The kernel is bound by instruction issue and the shared-memory pipe, not by HBM (12*C bytes per frame are nothing
// next to ~200 warp instructions per frame), so everything here is arranged to execute few instructions:
//
//   * linear-domain recursion on emissions in units of u(t) = max(y_blank(t), y_max(t) / 32): R[t][c] = y[t][c] / u(t)
//     <= 32.  Units of the blank's probability (one multiply per slot cheaper) let the band of a peaked alignment grow
//     by 2^29 per label and ran into the cap below; units of the largest class cost the band of a flat frame ~9 bits
//     and ran out of float32 at the other end.  log p gets sum_t log u(t) added back;
//   * slot i = (blank state 2(i-1), label state 2(i-1)+1); lane l of a recursion warp owns the NL consecutive slots
//     [l*NL, (l+1)*NL) in registers; a frame is, per slot, one LDS (the label's emission), five float operations
//     (two FFMA, FADD, two FMUL) and the cap, plus ONE shuffle for the neighbour lane's last label value;
//   * the backward recursion is the same code on the reversed label string over descending frames;
//   * float32 values with one integer exponent PER SLOT: true = stored * 2^E[k].  What crosses from slot k-1 into
//     slot k is multiplied by F[k] = 2^(E[k-1]-E[k]) inside the fused multiply-add that consumes it.  Every KC = 8
//     frames each slot is renormalised (larger state -> 2^(TB-127)) subject to E[k] >= E[k-1] - GCAP, so F <= 2^GCAP;
//     label values saturate at 2^100, hence nothing can reach inf and no NaN can arise from 0 * inf;
//   * meet in the middle: the forward warp covers frames [0,M) while the backward warp covers [M,Tb), each leaving a
//     checkpoint (state + exponents) per chunk; p comes from the two at the meeting point; then each continues through
//     the other half.  There the SAME warp first regenerates the other direction's eight label rows of the chunk from
//     that direction's checkpoint -- in its own lane layout, i.e. the other direction's recursion written with the
//     neighbours mirrored -- keeping them in registers, then advances its own recursion over the chunk and multiplies
//     its pre-emission sums with them: nothing but checkpoints leaves the warp, the [T,U] lattice never exists;
//   * posterior of a label state = pre-emission sum * other direction's value / p, computed as (q*c1)*(o*c2) with the
//     power of two 2^(E+Eo-e_p-64) split over both factors; the blank's occupancy is 1 - sum of label occupancies;
//   * posteriors are scattered into a row of cells sorted by class; classes are dealt to gradient lanes in groups of
//     four in order of falling count, so that a lane (class, frame) sums a contiguous run and the run lengths inside
//     a warp instruction are equal up to what the sort leaves over;
//   * a-posteriori certificate: saturation and underflow only ever REMOVE mass, so alpha(T-1)[final states]/p and beta(0)[first
//     states]/p fall short of 1 by exactly the posterior mass that was lost.  Both must be within 3e-5 of 1, no frame's
//     label occupancy may exceed 1 and no state may turn non-finite, else retry[b] hands the utterance to the robust
//     kernel.
//
// Data movement: a frame's logits row (4*C bytes, 8-byte aligned at C = 38) travels global -> shared memory as the
// 16-byte cp.async pieces of its 16-byte aligned superset, 32 rows per instruction; the producer warp converts 32
// rows at a time, one row per lane, in place.  Warps hand chunks to each other through
// monotonic counters in shared memory (st.release / ld.acquire): there is no CTA-wide barrier in the main loop.
//
//   warp 0,1  recursion forward / backward (phase 2: recompute of the other direction + combine)
//   warp 2    producer (cp.async rows, softmax pieces)      warp 3    gradient of both sides (class sums, row stores)
#include <limits.h>
#include <math.h>

#include <type_traits>

#include "nasr_common.cuh"

#ifndef NASR_TUNING
#define NASR_TUNING 0
#endif

namespace nasr {
namespace lean {

constexpr int KC = 8;             // frames per chunk (= renormalisation and checkpoint interval)
constexpr int NST = 8;            // chunks of row records per side in the ring
constexpr int NTHREADS = 128;     // 4 warps: two CTAs per SM leave every thread 255 registers
constexpr int TB = 107;           // biased exponent a slot's larger state is brought to (2^-20)
constexpr int GCAP = 20;          // a slot with mass sits at most this far below the slot with mass beneath it
constexpr int PSHIFT = 64;        // the posterior buffer holds posterior * 2^-PSHIFT
constexpr int ENEG = -(1 << 28);  // exponent tag of a slot that can never receive mass
constexpr int ROWW = 76;          // words per row record: odd number of 16-byte vectors (one row per lane, no bank
                                  // conflicts), room for the 16-byte aligned superset of a 64-class row + y_blank + zero
constexpr int ROWB = ROWW * 4;
constexpr int GS = 356;           // words per posterior row: = 4 (mod 32), so that 16 bytes of 8 frames hit 32 banks
constexpr int GSB = GS * 4;
constexpr int GCELLS = GS - 8;    // cells that may hold posteriors; then four cells nobody writes (zeros), then the
                                  // four dump cells of label-less slots
constexpr int NGRP = 16;          // class groups of four (63 label classes at most)
constexpr float kTol = 3e-5f;
constexpr float kUnitGap = 3.4657359f;   // 5 ln 2: see the producer
// Label values saturate here (2^100).  Mass that climbs a run of slots each GCAP bits below the one beneath it gains
// 2^GCAP in stored units per slot and can cross eight slots in a chunk: without the cap it reaches inf in slots ahead
// of the band that carries the posterior (measured: a quarter of the cfg3 utterances).  With F <= 2^20 eight frames of
// inflow stay below 2^124, so nothing ever reaches inf; what saturation removes the certificate accounts for.
#define NASR_BIG 1.2676506e30f

enum Role { R_F = 0, R_B = 1, PROD = 2, GRAD = 3 };
enum Alarm { AL_SHAPE = 1, AL_EMISSION = 2, AL_NONFINITE = 4, AL_SAT = 8, AL_P = 32, AL_CERT = 128, AL_OCC = 256 };
// counters in shared memory (each written by one warp at a time, only ever increasing)
enum Flag { FL_READY = 0, FL_FREED = 2, FL_GFULL = 4, FL_GFREE = 6, FL_MEETB = 8, FL_MEETP = 9, NFLAGS = 16 };

__device__ int g_sm_arrivals[1024];

struct Params {
  const float* logits;
  int T, B, C;
  long long st_t, st_b;
  const int32_t* lab_vals;
  const int32_t* lab_offs;
  const int32_t* seq_len;
  int blank;
  float* loss;
  float* grad;
  const float* grad_loss;
  int32_t* status;
  int32_t* retry;
  float4* ckpt;            // [B][2][maxch][3*NV][32]
  int maxch;
  int split;               // debug: frames of the forward half (multiple of KC), 0 = automatic
  const char* lo;          // lowest byte of the logits view
  const char* hi;          // one past its highest byte
  long long* prof;         // tuning builds: time stamps [role 5][chunk 160][4] of CTA 0
};

// ---- shared-memory accessors, counters, mbarrier, bulk copy ----------------------------------------------------
__device__ __forceinline__ uint32_t sptr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void flag_set(uint32_t a, int v) {
  asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ int flag_get(uint32_t a) {
  int v;
  asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
// a waiting warp must not take issue slots from the warps it waits for: back off between polls
template <int NS>
__device__ __forceinline__ void flag_wait(uint32_t a, int need) {
  while (flag_get(a) < need) __nanosleep(NS);
}
__device__ __forceinline__ void mbar_init(uint32_t a, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(a), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^e for e in (-inf, 127]; 0 at and below 2^-127
__device__ __forceinline__ float pow2f(int e) { return __int_as_float(max(e + 127, 0) << 23); }

// ---- shared memory layout -----------------------------------------------------------------------------------------
__host__ __device__ constexpr size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }
struct Smem {
  size_t bars, flags, rows, gbuf, obuf, meet, lab, cell, cnt, clsoff, gtab, lsum, scal, total;
};
__host__ __device__ inline Smem smem_layout(int NL) {
  Smem s;
  const int N = NL * 32;
  size_t o = 0;
  s.bars = o;    o = al16(o + 2 * 8);
  s.flags = o;   o = al16(o + NFLAGS * 4);
  s.rows = o;    o = al16(o + (size_t)2 * NST * KC * ROWB);
  s.gbuf = o;    o = al16(o + (size_t)2 * 2 * KC * GSB);
  s.obuf = o;    o = al16(o + (size_t)2 * KC * 32 * 32);          // [side][frame][lane][8]: the other direction's rows
  s.meet = o;    o = al16(o + (size_t)3 * N * 4);
  s.lab = o;     o = al16(o + (size_t)N * 4);
  s.cell = o;    o = al16(o + (size_t)N * 2);
  s.cnt = o;     o = al16(o + 64 * 4);
  s.clsoff = o;  o = al16(o + 64 * 4);
  s.gtab = o;    o = al16(o + (NGRP + 1) * 4 * 16);
  s.lsum = o;    o = al16(o + 32 * 4);
  s.scal = o;    o = al16(o + 64);
  s.total = o;
  return s;
}

// ---- schedule -----------------------------------------------------------------------------------------------------
struct Sched {
  int Tb;
  int n10, n11;    // frames each direction covers in phase 1 (scalars: arrays indexed by the direction would live
  int nc10, nc11;  // in local memory)            ... chunks of phase 1
  __device__ __forceinline__ int n1(int d) const { return d ? n11 : n10; }
  __device__ __forceinline__ int nc1(int d) const { return d ? nc11 : nc10; }
  int nch;      // chunks per side (phase 1 + phase 2)
  int grad;     // 0: loss only, no phase 2
};
struct Chunk {
  int phase;  // 0 none, 1, 2
  int len;    // frames
  int jo;     // phase 1: own chunk index; phase 2: the other direction's chunk index
  int t0;     // frame of consumer position 0
  int dt;     // +1 / -1: frame of position f is t0 + f * dt
};
__device__ __forceinline__ Chunk chunk_at(const Sched& S, int d, int i) {
  Chunk c;
  c.phase = 0; c.len = 0; c.jo = 0; c.t0 = 0; c.dt = 1;
  if (i < 0 || i >= S.nch) return c;
  if (i < S.nc1(d)) {
    const int tau0 = i * KC;
    c.phase = 1;
    c.jo = i;
    c.len = min(KC, S.n1(d) - tau0);
    c.t0 = d ? S.Tb - 1 - tau0 : tau0;
    c.dt = d ? -1 : 1;
  } else if (S.grad) {
    const int o = d ^ 1;
    const int jo = S.nc1(o) - 1 - (i - S.nc1(d));
    const int hi_tau = min(jo * KC + KC, S.n1(o));
    c.phase = 2;
    c.jo = jo;
    c.len = hi_tau - jo * KC;
    // consumer position f <-> other direction's position hi_tau-1-f
    c.t0 = o ? S.Tb - 1 - (hi_tau - 1) : hi_tau - 1;
    c.dt = o ? 1 : -1;
  }
  return c;
}

// ---- recursion ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lds_f(const unsigned char* smem, uint32_t off) {
  return *reinterpret_cast<const float*>(smem + off);
}
__device__ __forceinline__ void sts_f(unsigned char* smem, uint32_t off, float v) {
  *reinterpret_cast<float*>(smem + off) = v;
}

// The emissions of a frame for this lane's slots (r[NL] = the blank's).  They are loaded a frame ahead of their use:
// a single warp carries the recursion, so nothing else hides the latency of shared memory.
template <int NL>
__device__ __forceinline__ void load_em(float (&r)[NL + 1], const uint32_t (&coloff)[NL], uint32_t boff,
                                        const unsigned char* smem, uint32_t fo) {
#pragma unroll
  for (int k = 0; k < NL; k++) r[k] = lds_f(smem, coloff[k] + fo);
  r[NL] = lds_f(smem, boff + fo);
}

// One frame of this warp's own recursion (slots descending, in place).  r: the frame's emissions on entry, those of
// the frame at fnext on exit (pref).  COMBINE: also the posteriors of the frame, from the other direction's label
// values o[] of the same frame.
template <int NL, bool COMBINE>
__device__ __forceinline__ void own_frame(float (&A)[NL], float (&B)[NL], const float (&F)[NL], const float (&SF)[NL],
                                          float (&r)[NL + 1], const uint32_t (&coloff)[NL], uint32_t boff,
                                          unsigned char* smem, bool pref, uint32_t fnext, const float (&o)[NL],
                                          const float (&c1)[NL], const float (&c2)[NL], const uint32_t (&gph)[NL],
                                          uint32_t go) {
  const float a_in = __shfl_up_sync(0xffffffffu, A[NL - 1], 1);   // lane 0 multiplies it by F = 0
  float rn[NL + 1];
  if (pref) load_em<NL>(rn, coloff, boff, smem, fnext);
#pragma unroll
  for (int k = NL - 1; k >= 0; k--) {
    const float alp = k ? A[k - 1] : a_in;
    const float nb = fmaf(F[k], alp, B[k]);
    const float q = fmaf(SF[k], alp, A[k] + B[k]);
    if (COMBINE) sts_f(smem, gph[k] + go, (q * c1[k]) * (o[k] * c2[k]));
    A[k] = fminf(q * r[k], NASR_BIG);
    B[k] = nb * r[NL];
  }
  if (pref) {
#pragma unroll
    for (int k = 0; k <= NL; k++) r[k] = rn[k];
  }
}

// One frame of the OTHER direction's recursion in this warp's lane layout: position i holds that direction's label
// state of label i-1 and the blank AFTER it; mass arrives from position i+1 (slots ascending, in place).
template <int NL>
__device__ __forceinline__ void other_frame(float (&A)[NL], float (&B)[NL], const float (&F)[NL], const float (&SF)[NL],
                                            float (&r)[NL + 1], const uint32_t (&coloff)[NL], uint32_t boff,
                                            const unsigned char* smem, bool pref, uint32_t fnext, float (&o)[NL]) {
  const float a_in = __shfl_down_sync(0xffffffffu, A[0], 1);      // lane 31 multiplies it by F = 0
  float rn[NL + 1];
  if (pref) load_em<NL>(rn, coloff, boff, smem, fnext);
#pragma unroll
  for (int k = 0; k < NL; k++) {
    const float alp = (k < NL - 1) ? A[k + 1] : a_in;
    const float nb = fmaf(F[k], alp, B[k]);
    const float q = fmaf(SF[k], alp, A[k] + B[k]);
    A[k] = fminf(q * r[k], NASR_BIG);
    B[k] = nb * r[NL];
    o[k] = A[k];
  }
  if (pref) {
#pragma unroll
    for (int k = 0; k <= NL; k++) r[k] = rn[k];
  }
}

// Transfer factors of the own recursion from the slot exponents: F[k] = 2^(E[k-1]-E[k]) (<= 2^GCAP).  A dead slot
// (exponent ENEG) only ever sits below a live one, so 2^(ENEG - E) = 0 switches its outflow off by itself.
template <int NL>
__device__ __forceinline__ void set_F(float (&F)[NL], float (&SF)[NL], const int (&E)[NL], uint32_t skipmask, int lane) {
  int Ep = __shfl_up_sync(0xffffffffu, E[NL - 1], 1);
  if (lane == 0) Ep = ENEG;
#pragma unroll
  for (int k = 0; k < NL; k++) {
    const int lower = k ? E[k - 1] : Ep;
    F[k] = pow2f(min(lower - E[k], GCAP));
    SF[k] = ((skipmask >> k) & 1u) ? F[k] : 0.f;
  }
}
// ... of the other direction in this warp's layout: F[k] = 2^(E[k+1]-E[k]); there the dead slots sit above
template <int NL>
__device__ __forceinline__ void set_F_other(float (&F)[NL], float (&SF)[NL], const int (&E)[NL], uint32_t skipmask,
                                            int lane) {
  int En = __shfl_down_sync(0xffffffffu, E[0], 1);
  if (lane == 31) En = ENEG;
#pragma unroll
  for (int k = 0; k < NL; k++) {
    const int upper = (k < NL - 1) ? E[k + 1] : En;
    F[k] = pow2f(min(upper - E[k], GCAP));
    SF[k] = ((skipmask >> k) & 1u) ? F[k] : 0.f;
  }
}

// Renormalise every slot (larger state -> biased exponent TB), keeping E[k] >= E[k-1] - GCAP over slots with mass;
// empty slots adopt the exponent below.  The chain through a lane is x -> max(a, x - b) (a: what the lane yields on its
// own, b: GCAP per slot with mass), composed over the lanes by a scan.  A single warp runs this between two chunks of
// its recursion, so the dependent chain is kept short: one pass over the slots, the scan, every slot then finished
// independently of the others.
template <int NL>
__device__ __forceinline__ void rescale(float (&A)[NL], float (&B)[NL], float (&F)[NL], float (&SF)[NL], int (&E)[NL],
                                        uint32_t skipmask, int lane, int& alarm) {
  int En[NL], loc[NL], cnt[NL];
  uint32_t mall = 0;
  int prev = ENEG, c = 0;
#pragma unroll
  for (int k = 0; k < NL; k++) {
    const uint32_t m = max(__float_as_uint(A[k]), __float_as_uint(B[k]));
    mall = max(mall, m);
    const bool nz = m != 0u;                    // dead slots hold zeros
    En[k] = E[k] + (int)(m >> 23) - TB;
    prev = nz ? max(En[k], prev - GCAP) : prev;
    c += nz ? GCAP : 0;
    loc[k] = prev;
    cnt[k] = c;
  }
  if (mall >= 0x7f800000u) alarm |= AL_NONFINITE;   // inf or nan (sign bit set or exponent 255)
  // exponent below this lane's first slot: an inclusive scan of the lanes' maps x -> max(a, x - b), shifted by one lane
  // (exact; following only the two lanes below was measured to lose mass at steep fronts with two slots per lane)
  int ca = prev, cb = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int la = __shfl_up_sync(0xffffffffu, ca, o);
    const int lb = __shfl_up_sync(0xffffffffu, cb, o);
    if (lane >= o) {
      ca = max(ca, max(la, ENEG + (1 << 20)) - cb);
      cb += lb;
    }
  }
  int pin = __shfl_up_sync(0xffffffffu, ca, 1);
  if (lane == 0) pin = ENEG;
#pragma unroll
  for (int k = 0; k < NL; k++) {
    const int e = max(loc[k], max(pin, ENEG + (1 << 20)) - cnt[k]);
    const int d = min(max(E[k] - e, -127), 100);   // slots without mass hold zeros: any finite factor will do
    const float sc = __int_as_float((d + 127) << 23);
    A[k] *= sc;
    B[k] *= sc;
    E[k] = e > ENEG / 2 ? e : ENEG;
  }
  set_F<NL>(F, SF, E, skipmask, lane);
}

template <int NL>
__device__ __forceinline__ float4* ckpt_ptr(const Params& p, int b, int d, int c) {
  constexpr int NV = (NL + 3) / 4;
  return p.ckpt + (((size_t)b * 2 + d) * p.maxch + c) * (size_t)(3 * NV * 32);
}
__device__ __forceinline__ float4 ldcg4(const float4* p) { return __ldcg(p); }

template <int NL>
__global__ void __launch_bounds__(NTHREADS, 2) ctc_lean_kernel(const Params p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int N = NL * 32;
  constexpr int NV = (NL + 3) / 4;
  constexpr int Lcap = N - 2;
  const Smem sl = smem_layout(NL);
  float* s_gbuf = reinterpret_cast<float*>(smem + sl.gbuf);
  float* s_meet_nb = reinterpret_cast<float*>(smem + sl.meet);
  float* s_meet_pre = s_meet_nb + N;
  int* s_meet_e = reinterpret_cast<int*>(s_meet_pre + N);
  int* s_lab = reinterpret_cast<int*>(smem + sl.lab);
  uint16_t* s_cell = reinterpret_cast<uint16_t*>(smem + sl.cell);
  int* s_cnt = reinterpret_cast<int*>(smem + sl.cnt);
  int* s_clsoff = reinterpret_cast<int*>(smem + sl.clsoff);
  int4* s_gtab = reinterpret_cast<int4*>(smem + sl.gtab);   // [group][class in group]: class, first cell, run length
  float* s_lsum = reinterpret_cast<float*>(smem + sl.lsum);
  int* s_scal = reinterpret_cast<int*>(smem + sl.scal);   // [0] alarm [1] e_p [2] repeats [3] 1/m_p [4] SM arrival
  int* s_flag = reinterpret_cast<int*>(smem + sl.flags);

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably uniform: no reconvergence code around shuffles
  const int T = p.T, C = p.C, blank = p.blank;
  const uint32_t bar0 = sptr(smem + sl.bars);
  const uint32_t flag0 = sptr(s_flag);
#define FLAG(i) (flag0 + 4u * (uint32_t)(i))
#if NASR_TUNING
#define TS(role_, chunk_, slot_)                                                                    \
  do {                                                                                              \
    if (p.prof && b == 0 && lane == 0 && (chunk_) < 160)                                            \
      p.prof[((role_) * 160 + (chunk_)) * 4 + (slot_)] = clock64();                                 \
  } while (0)
#else
#define TS(role_, chunk_, slot_) do {} while (0)
#endif

#if NASR_TUNING
  if (p.prof && tid == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.prof[3700 + 2 * b] = (long long)gt;
  }
#endif
  const int Tb = p.seq_len[b];
  const int l0 = p.lab_offs[b];
  const int L = p.lab_offs[b + 1] - l0;

  // ---- can this kernel take the utterance? everything unusual goes to the robust kernel -----------------------
  int bad = (Tb < 2 * KC) | (Tb > T) | (L < 0) | (L > Lcap) | (C > 64);
  if (tid < 8) s_scal[tid] = 0;
  if (tid < NFLAGS) s_flag[tid] = 0;
  if (tid < 64) s_cnt[tid] = 0;
  if (tid < 32) s_lsum[tid] = 0.f;
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  __syncthreads();
  if (tid == 0) s_scal[4] = atomicAdd(&g_sm_arrivals[smid & 1023], 1);
  int rep = 0;
  if (!bad) {
    for (int i = tid; i < L; i += NTHREADS) {
      const int v = p.lab_vals[l0 + i];
      s_lab[i] = v;
      if (v < 0 || v >= C || v == blank) {
        bad = 1;
      } else {
        atomicAdd(&s_cnt[v], 1);
        if (i > 0 && v == p.lab_vals[l0 + i - 1]) rep++;
      }
    }
  }
  if (rep) atomicAdd(&s_scal[2], rep);
  bad = __syncthreads_or(bad);
  if (!bad && Tb < L + s_scal[2]) bad = 1;
  // ---- class table of the gradient warps: label classes in order of falling count, dealt in groups of four; the
  // cells of a group's classes are four runs of the group's run length (a multiple of four cells: a lane (class,
  // frame) reads its run in 16-byte pieces, and the eight frames of a class cover all 32 banks)
  if (!bad && warp == 0) {
    for (int i = lane; i < (NGRP + 1) * 4; i += 32) s_gtab[i] = make_int4(-1, GCELLS, 0, 0);
    __syncwarp();
    for (int c = lane; c < C; c += 32)
      if (c != blank) {
        const int n = s_cnt[c];
        int rank = 0;
        for (int o = 0; o < C; o++) {
          const int m = s_cnt[o];
          rank += (o != blank && (m > n || (m == n && o < c))) ? 1 : 0;
        }
        s_gtab[rank].x = c;
      }
    __syncwarp();
    if (lane == 0) {
      int base = 0;
      const int ngrp = (C - 1 + 3) / 4;
      for (int g = 0; g < ngrp; g++) {
        const int c0 = s_gtab[4 * g].x;
        const int trip = c0 >= 0 ? (s_cnt[c0] + 3) & ~3 : 0;
        for (int j = 0; j < 4; j++) {
          s_gtab[4 * g + j].y = trip ? base + j * trip : GCELLS;
          s_gtab[4 * g + j].z = trip;
        }
        base += 4 * trip;
      }
      if (base > GCELLS) s_scal[0] = AL_SHAPE;
    }
    __syncwarp();
    for (int r = lane; r < NGRP * 4; r += 32) {
      const int4 e = s_gtab[r];
      if (e.x >= 0) s_clsoff[e.x] = e.y;
    }
  }
  __syncthreads();
  if (bad || s_scal[0]) {
    if (tid == 0) {
      p.retry[b] = AL_SHAPE;
      atomicSub(&g_sm_arrivals[smid & 1023], 1);
    }
    return;
  }
  for (int j = tid; j < L; j += NTHREADS) {
    const int v = s_lab[j];
    int r = 0;
    for (int i = 0; i < j; i++) r += (s_lab[i] == v);
    s_cell[j] = (uint16_t)(s_clsoff[v] + r);
  }
  // posterior rows: padding cells are never written and must read as zero
  for (int i = tid; i < 2 * 2 * KC * GS; i += NTHREADS) s_gbuf[i] = 0.f;
  if (tid == 0) {
    mbar_init(bar0, 32);
    mbar_init(bar0 + 8, 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // ---- gradient rows of padded frames are exactly zero ------------------------------------------------------------
  const size_t rstride = (size_t)p.st_t;
  float* gbase = p.grad ? p.grad + (size_t)b * p.st_b : nullptr;
  if (gbase) {
    const size_t n = (size_t)(T - Tb) * C;
    for (size_t i = tid; i < n; i += NTHREADS) {
      const size_t t = Tb + i / C;
      gbase[t * rstride + (i % C)] = 0.f;
    }
  }
  const bool want_grad = gbase != nullptr;
  const float gs = p.grad_loss ? p.grad_loss[b] : 1.0f;

  // ---- schedule -----------------------------------------------------------------------------------------------------
  Sched S;
  S.Tb = Tb;
  {
    int nf = KC * ((Tb + KC) / (2 * KC));
    if (p.split > 0) nf = (p.split / KC) * KC;
    nf = max(KC, min(nf, ((Tb - 1) / KC) * KC));
    S.n10 = nf;
    S.n11 = Tb - nf;
    S.nc10 = (S.n10 + KC - 1) / KC;
    S.nc11 = (S.n11 + KC - 1) / KC;
    S.nch = want_grad ? S.nc10 + S.nc11 : max(S.nc10, S.nc11);
    S.grad = want_grad ? 1 : 0;
  }
  // every row of this utterance starts at the same offset inside its 16-byte aligned superset (the launcher checked)
  const float* xbase = p.logits + (size_t)b * p.st_b;
  const int shift = (int)(((uintptr_t)xbase & 15) >> 2);
  __syncthreads();

  // Roles go by the scheduler a warp issues from (hardware warp slot mod 4; measured: the second CTA of an SM gets
  // slots 5, 6, 7, 4 for its warps 0..3), and the second CTA that lands on an SM deals them the other way round, so
  // that the four recursion warps of an SM sit on four different schedulers, each beside one of the helper warps.
  unsigned hwid;
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(hwid));
  hwid = __shfl_sync(0xffffffffu, hwid, 0);
  if (lane == 0) s_flag[NFLAGS - 4 + warp] = (int)(hwid & 3u);
  __syncthreads();
  int role = warp;
  {
    // (fall back to the warp's index if the four slots do not cover the four schedulers)
    const int m = (1 << s_flag[NFLAGS - 4]) | (1 << s_flag[NFLAGS - 3]) | (1 << s_flag[NFLAGS - 2]) | (1 << s_flag[NFLAGS - 1]);
    const int q = m == 15 ? (int)(hwid & 3u) : warp;
    role = (s_scal[4] & 1) ? (q ^ 2) : q;
  }
  const int d = role == R_B ? 1 : 0;   // direction / side of a recursion warp
#if NASR_TUNING
  if (p.prof && lane == 0 && b >= 148 && b < 212) {
    unsigned hwid;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(hwid));
    p.prof[3400 + (b - 148) * 4 + warp] = ((long long)smid << 32) | (hwid << 8) | (unsigned)role;
  }
#endif
  int alarm = 0;

  if (role == R_F || role == R_B) {
    // ================================ recursion warps =================================
    float A[NL], B[NL], F[NL], SF[NL];
    int E[NL];
    uint32_t coloff[NL], gph[NL];
    uint32_t skipmask = 0, skipo = 0;
    uint32_t boff = (uint32_t)sl.rows + (uint32_t)(d * NST) * KC * ROWB + (uint32_t)(shift + blank) * 4u;
    const uint32_t rows0 = (uint32_t)sl.rows + (uint32_t)(d * NST) * KC * ROWB;   // stage 0 of this side
    const uint32_t gbuf0 = (uint32_t)sl.gbuf + (uint32_t)(d * 2) * KC * GSB;      // buffer 0 of this side
    {
      // slot tables for direction d (0 forward, 1 mirrored) and the virtual row before the first frame
      const int pad = N - L - 1;
      const int first = d == 0 ? 1 : pad;
#pragma unroll
      for (int k = 0; k < NL; k++) {
        const int i = lane * NL + k;
        int j = -1;   // label index (in the direction's own order) of slot i
        if (d == 0) {
          if (i - 1 >= 0 && i - 1 < L) j = i - 1;
        } else {
          if (i - pad >= 0 && i - pad < L) j = i - pad;
        }
        int col = -1, cell = GCELLS + 4 + (k & 3);
        bool skip = false, skipn = false;
        if (j >= 0) {
          const int jj = d == 0 ? j : L - 1 - j;          // index into the transcript
          const int jp = d == 0 ? j - 1 : L - j;          // the label before it in the direction's order
          const int jn = d == 0 ? j + 1 : L - 2 - j;      // the label after it in the direction's order
          col = s_lab[jj];
          cell = s_cell[jj];
          skip = j >= 1 && s_lab[jj] != s_lab[jp];
          skipn = j + 1 < L && s_lab[jj] != s_lab[jn];
        }
        coloff[k] = rows0 + (col >= 0 ? (uint32_t)(shift + col) * 4u : (uint32_t)(ROWW - 1) * 4u);
        gph[k] = gbuf0 + (uint32_t)cell * 4u;
        skipmask |= skip ? (1u << k) : 0u;
        skipo |= skipn ? (1u << k) : 0u;
        B[k] = (i == first) ? __int_as_float(TB << 23) : 0.f;
        A[k] = 0.f;
        E[k] = i >= first ? 127 - TB : ENEG;
      }
      set_F<NL>(F, SF, E, skipmask, lane);
    }
    float c1[NL], c2[NL], Ao[NL], Bo[NL], Fo[NL], SFo[NL], r[NL + 1];
    int Eo[NL];
#pragma unroll
    for (int k = 0; k < NL; k++) {
      c1[k] = c2[k] = Ao[k] = Bo[k] = Fo[k] = SFo[k] = 0.f;
      Eo[k] = ENEG;
    }
    int e_p = 0;
    int stg = 0;   // ring stage of chunk i
#pragma unroll 1
    for (int i = 0; i < S.nch; i++) {
      const Chunk ci = chunk_at(S, d, i);
      if (ci.phase == 0) break;   // loss only: this side has fewer phase-1 chunks than the other
      if (ci.phase == 1) {
        if (i) rescale<NL>(A, B, F, SF, E, skipmask, lane, alarm);
        if (want_grad) {
          float4* ck = ckpt_ptr<NL>(p, b, d, i);
          float a[4 * NV], bb[4 * NV];
          int e[4 * NV];
#pragma unroll
          for (int k = 0; k < 4 * NV; k++) {
            a[k] = k < NL ? A[k] : 0.f;
            bb[k] = k < NL ? B[k] : 0.f;
            e[k] = k < NL ? E[k] : ENEG;
          }
#pragma unroll
          for (int v = 0; v < NV; v++) {
            ck[v * 32 + lane] = make_float4(a[4 * v], a[4 * v + 1], a[4 * v + 2], a[4 * v + 3]);
            ck[(NV + v) * 32 + lane] = make_float4(bb[4 * v], bb[4 * v + 1], bb[4 * v + 2], bb[4 * v + 3]);
            reinterpret_cast<int4*>(ck)[(2 * NV + v) * 32 + lane] = make_int4(e[4 * v], e[4 * v + 1], e[4 * v + 2], e[4 * v + 3]);
          }
        }
        TS(role, i, 0);
        flag_wait<20>(FLAG(FL_READY + d), i + 1);
        TS(role, i, 1);
        load_em<NL>(r, coloff, boff, smem, 0);
        if (ci.len == KC) {
#pragma unroll
          for (int f = 0; f < KC; f++)
            own_frame<NL, false>(A, B, F, SF, r, coloff, boff, smem, f < KC - 1, (f + 1) * ROWB, c1, c1, c2, gph, 0);
        } else {
#pragma unroll 1
          for (int f = 0; f < ci.len; f++)
            own_frame<NL, false>(A, B, F, SF, r, coloff, boff, smem, true, (f + 1) * ROWB, c1, c1, c2, gph, 0);
        }
        __syncwarp();
        TS(role, i, 2);
        if (lane == 0) flag_set(FLAG(FL_FREED + d), i + 1);
        TS(role, i, 3);
        if (i == S.nc1(d) - 1) {
          // ---- end of phase 1: checkpoints visible to the other recursion warp, then the meeting ----
          __threadfence_block();
          rescale<NL>(A, B, F, SF, E, skipmask, lane, alarm);
          if (d == 1) {
            // pre-emission sums of the frame below the meeting point, for the forward warp
            const float a_in = __shfl_up_sync(0xffffffffu, A[NL - 1], 1);
#pragma unroll
            for (int k = 0; k < NL; k++) {
              const float alp = k ? A[k - 1] : a_in;
              s_meet_nb[k * 32 + lane] = fmaf(F[k], alp, B[k]);
              s_meet_pre[k * 32 + lane] = fmaf(SF[k], alp, A[k] + B[k]);
              s_meet_e[k * 32 + lane] = E[k];
            }
            __syncwarp();
            if (lane == 0) flag_set(FLAG(FL_MEETB), 1);
            flag_wait<100>(FLAG(FL_MEETP), 1);
          } else {
            flag_wait<100>(FLAG(FL_MEETB), 1);
            // ---- p = sum over states of alpha(M-1) * beta(M-1); label of slot i pairs with mirrored slot N-1-i,
            // the blank of slot i (state 2(i-1)) with the blank of mirrored slot N-i ----
            const float S64 = 1.0f;           // both factors sit near 2^0 after the rescale
            const int lm = 31 - lane;
            float term[2 * NL];
            int kt[2 * NL];
            int kmax = INT_MIN / 2;
#pragma unroll
            for (int k = 0; k < NL; k++) {
              const int mk = NL - 1 - k;
              const int Em = s_meet_e[mk * 32 + lm];
              const float tl = (A[k] * S64) * (s_meet_pre[mk * 32 + lm] * S64);
              const bool okl = tl > 0.f && E[k] > ENEG / 2 && Em > ENEG / 2;
              term[k] = okl ? tl : 0.f;
              kt[k] = E[k] + Em;
              // blank partner: global mirrored slot N - i = (lane', slot') one above the label partner
              const int gi = N - (lane * NL + k);
              const int bl = gi / NL, bk = gi % NL;
              float tb = 0.f;
              int Eb = ENEG;
              if (gi < N) {
                Eb = s_meet_e[bk * 32 + bl];
                tb = (B[k] * S64) * (s_meet_nb[bk * 32 + bl] * S64);
              }
              const bool okb = tb > 0.f && E[k] > ENEG / 2 && Eb > ENEG / 2;
              term[NL + k] = okb ? tb : 0.f;
              kt[NL + k] = E[k] + Eb;
              if (okl) kmax = max(kmax, kt[k] + (int)((__float_as_uint(term[k]) >> 23) & 255u) - 127);
              if (okb) kmax = max(kmax, kt[NL + k] + (int)((__float_as_uint(term[NL + k]) >> 23) & 255u) - 127);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
            float tot = 0.f;
#pragma unroll
            for (int k = 0; k < 2 * NL; k++)
              if (term[k] > 0.f) {
                const int sh = max(-250, kt[k] - kmax);   // <= -(exponent of the term) <= ~+140
                const int h = sh >> 1;
                tot += (term[k] * pow2f(h)) * pow2f(sh - h);
              }
            tot = warp_sum(tot);
            if (!(tot > 0.f) || !(tot < 3e38f) || kmax < -(1 << 27)) {
              alarm |= AL_P;
              tot = 1.f;
              kmax = 0;
            }
            const int et = (int)((__float_as_uint(tot) >> 23) & 255u) - 127;
            const float mp = tot * pow2f(-et);
            double ls = 0.0;
#pragma unroll 1
            for (int q = 0; q < 32; q++) ls += (double)*reinterpret_cast<volatile float*>(&s_lsum[q]);
            if (lane == 0) {
              s_scal[1] = kmax + et;
              s_scal[3] = __float_as_int(1.0f / mp);
              // log p = kmax ln 2 + ln(tot) - ln 2 * sum_t log2(sum_c R(t))   (y = R / sum_c R)
              p.loss[b] = (float)(-((double)kmax * 0.6931471805599453 + log((double)tot) - ls * 0.6931471805599453));
              p.status[b] = 0;
            }
            __syncwarp();
            if (lane == 0) flag_set(FLAG(FL_MEETP), 1);
          }
          __syncwarp();
          e_p = *reinterpret_cast<volatile int*>(&s_scal[1]);
          const float im = __int_as_float(*reinterpret_cast<volatile int*>(&s_scal[3]));
#pragma unroll
          for (int k = 0; k < NL; k++) {
            A[k] *= im;
            B[k] *= im;
          }
          if (want_grad && S.nch > S.nc1(d)) {
            // the other direction's checkpoint of the first chunk of phase 2
            const Chunk cn = chunk_at(S, d, i + 1);
            const float4* ck = ckpt_ptr<NL>(p, b, d ^ 1, cn.jo);
            float a[4 * NV], bb[4 * NV];
            int e[4 * NV];
#pragma unroll
            for (int v = 0; v < NV; v++) {
              const float4 va = ldcg4(ck + v * 32 + (31 - lane));
              const float4 vb = ldcg4(ck + (NV + v) * 32 + (31 - lane));
              const int4 ve = __ldcg(reinterpret_cast<const int4*>(ck) + (2 * NV + v) * 32 + (31 - lane));
              a[4 * v] = va.x; a[4 * v + 1] = va.y; a[4 * v + 2] = va.z; a[4 * v + 3] = va.w;
              bb[4 * v] = vb.x; bb[4 * v + 1] = vb.y; bb[4 * v + 2] = vb.z; bb[4 * v + 3] = vb.w;
              e[4 * v] = ve.x; e[4 * v + 1] = ve.y; e[4 * v + 2] = ve.z; e[4 * v + 3] = ve.w;
            }
#pragma unroll
            for (int k = 0; k < NL; k++) {
              Ao[k] = a[NL - 1 - k];
              Bo[k] = bb[NL - 1 - k];
              Eo[k] = e[NL - 1 - k];
            }
          }
        }
      } else {
        // ---- phase 2: continue through the other half against the other direction's values, regenerated here ----
        const int j2 = i - S.nc1(d);
        rescale<NL>(A, B, F, SF, E, skipmask, lane, alarm);
        set_F_other<NL>(Fo, SFo, Eo, skipo, lane);
#pragma unroll
        for (int k = 0; k < NL; k++) {
          int kk = E[k] + Eo[k] - e_p - PSHIFT;
          kk = max(-250, min(250, kk));
          const int h = kk >> 1;
          c1[k] = __int_as_float((h + 127) << 23);
          c2[k] = __int_as_float((kk - h + 127) << 23);
        }
        TS(role, i, 0);
        flag_wait<20>(FLAG(FL_READY + d), i + 1);
        TS(role, i, 1);
        // The chunk's frames run through two-frame loop bodies (a chunk fully unrolled is ~27 KB of code per warp: the
        // profile of that version had "no instruction" -- instruction-cache misses -- as its first stall reason), and
        // the other direction's rows go through shared memory in between: [frame][lane][8], two 16-byte stores per frame.
        const uint32_t obase = (uint32_t)sl.obuf + (uint32_t)d * KC * 32 * 32 + (uint32_t)lane * 32;
        {
          uint32_t wc[NL], wb;
          const uint32_t top = (uint32_t)(ci.len - 1) * ROWB;
#pragma unroll
          for (int k = 0; k < NL; k++) wc[k] = coloff[k] + top;
          wb = boff + top;
          load_em<NL>(r, wc, wb, smem, 0);
          auto put = [&](int f, const float (&o)[NL]) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = k < NL ? o[k] : 0.f;
            float4* q = reinterpret_cast<float4*>(smem + obase + (uint32_t)f * (32 * 32));
            q[0] = make_float4(v[0], v[1], v[2], v[3]);
            if (NL > 4) q[1] = make_float4(v[4], v[5], v[6], v[7]);
          };
          int f = ci.len - 1;
          float o[NL];
          if (ci.len & 1) {
            other_frame<NL>(Ao, Bo, Fo, SFo, r, wc, wb, smem, f > 0, (uint32_t)(-ROWB), o);
            put(f, o);
            f--;
#pragma unroll
            for (int k = 0; k < NL; k++) wc[k] -= ROWB;
            wb -= ROWB;
          }
#pragma unroll 1
          for (; f >= 1; f -= 2) {
            other_frame<NL>(Ao, Bo, Fo, SFo, r, wc, wb, smem, true, (uint32_t)(-ROWB), o);
            put(f, o);
            other_frame<NL>(Ao, Bo, Fo, SFo, r, wc, wb, smem, f > 1, (uint32_t)(-2 * ROWB), o);
            put(f - 1, o);
#pragma unroll
            for (int k = 0; k < NL; k++) wc[k] -= 2 * ROWB;
            wb -= 2 * ROWB;
          }
        }
        // r now holds the emissions of the chunk's first frame again: where the own recursion starts
        TS(role, i, 2);
        flag_wait<20>(FLAG(FL_GFREE + d), j2 - 1);
        TS(role, i, 3);
        if (i + 1 < S.nch) {
          // the other direction's checkpoint of the next chunk, on its way while this chunk is combined
          const Chunk cn = chunk_at(S, d, i + 1);
          const float4* ck = ckpt_ptr<NL>(p, b, d ^ 1, cn.jo);
          float a[4 * NV], bb[4 * NV];
          int e[4 * NV];
#pragma unroll
          for (int v = 0; v < NV; v++) {
            const float4 va = ldcg4(ck + v * 32 + (31 - lane));
            const float4 vb = ldcg4(ck + (NV + v) * 32 + (31 - lane));
            const int4 ve = __ldcg(reinterpret_cast<const int4*>(ck) + (2 * NV + v) * 32 + (31 - lane));
            a[4 * v] = va.x; a[4 * v + 1] = va.y; a[4 * v + 2] = va.z; a[4 * v + 3] = va.w;
            bb[4 * v] = vb.x; bb[4 * v + 1] = vb.y; bb[4 * v + 2] = vb.z; bb[4 * v + 3] = vb.w;
            e[4 * v] = ve.x; e[4 * v + 1] = ve.y; e[4 * v + 2] = ve.z; e[4 * v + 3] = ve.w;
          }
#pragma unroll
          for (int k = 0; k < NL; k++) {
            Ao[k] = a[NL - 1 - k];
            Bo[k] = bb[NL - 1 - k];
            Eo[k] = e[NL - 1 - k];
          }
        }
        {
          uint32_t wc[NL], wg[NL], wb = boff;
#pragma unroll
          for (int k = 0; k < NL; k++) {
            wc[k] = coloff[k];
            wg[k] = gph[k];
          }
          auto get = [&](int f, float (&o)[NL]) {
            const float4* q = reinterpret_cast<const float4*>(smem + obase + (uint32_t)f * (32 * 32));
            const float4 a = q[0];
            float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
            if (NL > 4) bq = q[1];
            const float v[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int k = 0; k < NL; k++) o[k] = v[k];
          };
          __syncwarp();
          float o[NL];
          int f = 0;
#pragma unroll 1
          for (; f + 1 < ci.len; f += 2) {
            get(f, o);
            own_frame<NL, true>(A, B, F, SF, r, wc, wb, smem, true, ROWB, o, c1, c2, wg, 0);
            get(f + 1, o);
            own_frame<NL, true>(A, B, F, SF, r, wc, wb, smem, f + 2 < ci.len, 2 * ROWB, o, c1, c2, wg, GSB);
#pragma unroll
            for (int k = 0; k < NL; k++) {
              wc[k] += 2 * ROWB;
              wg[k] += 2 * GSB;
            }
            wb += 2 * ROWB;
          }
          if (f < ci.len) {
            get(f, o);
            own_frame<NL, true>(A, B, F, SF, r, wc, wb, smem, false, 0, o, c1, c2, wg, 0);
          }
        }
        __syncwarp();
        if (lane == 0) flag_set(FLAG(FL_GFULL + d), j2 + 1);
        const uint32_t gd = (j2 & 1) ? (uint32_t)(-KC * GSB) : (uint32_t)(KC * GSB);
#pragma unroll
        for (int k = 0; k < NL; k++) {
          gph[k] += gd;
          asm volatile("" : "+r"(gph[k]));   // keep the updated address in its register: the stores take [reg + imm]
        }
      }
      // the row records of the next chunk are in the next ring stage
      const uint32_t cd = stg + 1 == NST ? (uint32_t)(-(NST - 1) * KC * ROWB) : (uint32_t)(KC * ROWB);
#pragma unroll
      for (int k = 0; k < NL; k++) {
        coloff[k] += cd;
        asm volatile("" : "+r"(coloff[k]));
      }
      boff += cd;
      asm volatile("" : "+r"(boff));
      stg = stg + 1 == NST ? 0 : stg + 1;
    }
    if (want_grad) {
      // ---- certificate: what reached the end of this direction, over p ----
      const int sb = d == 0 ? L + 1 : N - 1;   // slot whose blank is the last state of this direction
      const int sl2 = d == 0 ? L : N - 2;      // slot whose label is the state before it
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < NL; k++) {
        const int i = lane * NL + k;
        const int sh = max(-126, min(126, E[k] - e_p));
        if (i == sb && E[k] > ENEG / 2) v += B[k] * pow2f(sh);
        if (i == sl2 && L > 0 && E[k] > ENEG / 2) v += A[k] * pow2f(sh);
      }
      v = warp_sum(v);
      if (!(fabsf(v - 1.0f) < kTol)) alarm |= AL_CERT;
    }
  } else if (role == PROD) {
    // ================================ producer warp =================================
    // A round brings the next two chunks of both sides: 32 rows, one per lane (side, chunk in round, frame), as
    // 16-byte cp.async pieces of the 16-byte aligned superset of the row (rows are only 8-byte aligned at C = 38, and
    // 32 scattered 152-byte rows per instruction are exactly what LDGSTS is for; a bulk copy per row would cost an
    // elected-lane loop of 32 issues).  Three rounds are in flight while the oldest is converted in place: raw logits
    // -> ratio emissions, y_blank, the zero entry.  The sides advance independently: a side whose ring is full (its
    // recursion warp waits at the meeting point) sits the round out and must not hold back the other one.
    const int side = lane >> 4, cip = (lane >> 3) & 1, f = lane & 7;
    const size_t rstr = rstride;
    float lsum = 0.f;
    const int nvu = (shift + C + 3) >> 2;   // 16-byte vectors that hold the row
    int rs = 0;                             // rounds this lane's side has issued
    int q0 = -1, q1 = -1, q2 = -1;          // side rounds in flight, oldest first (-1: the side sat that round out)
    int it = 0;
#pragma unroll 1
    while (true) {
      TS(role, it, 0);
      // ---- issue ----
      int nq = -1;
      {
        const int ilast = min(2 * rs + 1, S.nch - 1);
        // the stages must have been released by their readers of NST chunks ago (the same answer in all lanes of a side)
        const bool can = 2 * rs < S.nch && flag_get(FLAG(FL_FREED + side)) >= ilast - NST + 1;
        if (can) {
          const int i = 2 * rs + cip;
          const Chunk ci = chunk_at(S, side, i);
          if (ci.phase != 0 && f < ci.len) {
            unsigned char* slot = smem + sl.rows + ((size_t)(side * NST + i % NST) * KC + f) * ROWB;
            const int t = ci.t0 + f * ci.dt;
            const char* row = reinterpret_cast<const char*>(xbase + (size_t)t * rstr);
            const char* a0 = reinterpret_cast<const char*>((uintptr_t)row & ~(uintptr_t)15);
            if (a0 >= p.lo && a0 + 16 * nvu <= p.hi) {
              const uint32_t dst = sptr(slot);
#pragma unroll 1
              for (int v = 0; v < nvu; v++)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (uint32_t)v), "l"(a0 + 16 * v)
                             : "memory");
            } else {
              // the superset of this row would leave the tensor: element-wise copy
              float* dstf = reinterpret_cast<float*>(slot) + shift;
              const float* src = reinterpret_cast<const float*>(row);
              for (int c = 0; c < C; c++) dstf[c] = __ldg(src + c);
            }
          }
          nq = rs;
          rs++;
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      TS(role, it, 1);
      asm volatile("cp.async.wait_group 3;" ::: "memory");
      TS(role, it, 2);
      // ---- convert the oldest round in flight ----
      if (q0 >= 0) {
        const int i = 2 * q0 + cip;
        const Chunk cc = chunk_at(S, side, i);
        if (cc.phase != 0 && f < cc.len) {
          float* slot = reinterpret_cast<float*>(smem + sl.rows + ((size_t)(side * NST + i % NST) * KC + f) * ROWB);
          // emissions in units of u = max(blank's, largest class's / 32): R[c] = exp(x[c] - log u).  Flat frames
          // (the label on the path is a typical class, far below the largest) then cost a state ~4 bits instead of
          // ~9, peaked frames (the label IS the largest class and the blank is tiny) let it grow by at most 5 bits
          // instead of by the blank's 11..33: eight frames stay inside float32 either way.
          // The words of the 16-byte pieces that are not this row's become -3e38 first (exp -> 0, no part in the
          // maximum), so that the pieces are processed whole; a row of up to 41 classes (11 pieces) is held in
          // registers: one warp converts all rows, the loads must not wait for each other.
          for (int c = 0; c < shift; c++) slot[c] = -3.0e38f;
          for (int c = shift + C; c < 4 * nvu; c++) slot[c] = -3.0e38f;
          const float xbl = slot[shift + blank];
          float rsum;
          if (nvu <= 11) {
            float4 x[11];
#pragma unroll
            for (int v = 0; v < 11; v++) x[v] = *reinterpret_cast<const float4*>(slot + 4 * v);
            float m[11];
#pragma unroll
            for (int v = 0; v < 11; v++) m[v] = v < nvu ? fmaxf(fmaxf(x[v].x, x[v].y), fmaxf(x[v].z, x[v].w)) : -3.0e38f;
            const float mx = fmaxf(fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7]))),
                                   fmaxf(fmaxf(m[8], m[9]), m[10]));
            const float nxb = -fmaxf(xbl, mx - kUnitGap) * 1.4426950408889634f;
            float sv[11];
#pragma unroll
            for (int v = 0; v < 11; v++) {
              float4 rr;
              rr.x = ex2a(fmaf(x[v].x, 1.4426950408889634f, nxb));
              rr.y = ex2a(fmaf(x[v].y, 1.4426950408889634f, nxb));
              rr.z = ex2a(fmaf(x[v].z, 1.4426950408889634f, nxb));
              rr.w = ex2a(fmaf(x[v].w, 1.4426950408889634f, nxb));
              sv[v] = v < nvu ? (rr.x + rr.y) + (rr.z + rr.w) : 0.f;
              if (v < nvu) *reinterpret_cast<float4*>(slot + 4 * v) = rr;
            }
            rsum = (((sv[0] + sv[1]) + (sv[2] + sv[3])) + ((sv[4] + sv[5]) + (sv[6] + sv[7]))) + ((sv[8] + sv[9]) + sv[10]);
          } else {
            float m0 = -3.0e38f, m1 = -3.0e38f;
#pragma unroll 2
            for (int v = 0; v < nvu; v++) {
              const float4 x = *reinterpret_cast<const float4*>(slot + 4 * v);
              m0 = fmaxf(m0, fmaxf(x.x, x.y));
              m1 = fmaxf(m1, fmaxf(x.z, x.w));
            }
            const float nxb = -fmaxf(xbl, fmaxf(m0, m1) - kUnitGap) * 1.4426950408889634f;
            float rs0 = 0.f, rs1 = 0.f;
#pragma unroll 2
            for (int v = 0; v < nvu; v++) {
              const float4 x = *reinterpret_cast<const float4*>(slot + 4 * v);
              float4 rr;
              rr.x = ex2a(fmaf(x.x, 1.4426950408889634f, nxb));
              rr.y = ex2a(fmaf(x.y, 1.4426950408889634f, nxb));
              rr.z = ex2a(fmaf(x.z, 1.4426950408889634f, nxb));
              rr.w = ex2a(fmaf(x.w, 1.4426950408889634f, nxb));
              rs0 += rr.x + rr.y;
              rs1 += rr.z + rr.w;
              *reinterpret_cast<float4*>(slot + 4 * v) = rr;
            }
            rsum = rs0 + rs1;
          }
          // a class ratio that overflowed (or a non-finite logit) is the robust kernel's business; ratios that
          // underflow only remove mass and are covered by the certificate
          if (!(rsum < 1e37f) || !(rsum > 0.f)) alarm |= AL_EMISSION;
          slot[ROWW - 2] = __fdividef(1.0f, rsum);   // y[c] = R[c] / sum
          slot[ROWW - 1] = 0.f;                      // "emission" of slots without a label
          if (cc.phase == 1) lsum += __log2f(rsum);
        }
        s_lsum[lane] = lsum;
      }
      __syncwarp();
      if (q0 >= 0 && (lane & 15) == 0) flag_set(FLAG(FL_READY + side), min(2 * q0 + 2, S.nch));
      TS(role, it, 3);
      if (NASR_TUNING && p.prof && b == 0 && lane == 0 && it < 160) p.prof[3200 + it] = ((long long)q0 << 32) | (unsigned)nq;
      it++;
      const bool idle = q0 < 0 && nq < 0;
      q0 = q1; q1 = q2; q2 = nq;
      const bool more = 2 * rs < S.nch || q0 >= 0 || q1 >= 0 || q2 >= 0;
      if (!__any_sync(0xffffffffu, more)) break;
      // nothing to request and nothing to convert: the producer is up to NST chunks (tens of microseconds) ahead of
      // its readers and shares a scheduler with a recursion warp -- sleep, do not poll
      if (__all_sync(0xffffffffu, idle)) __nanosleep(2000);
    }
  } else {
    // ================================ gradient warp =================================
    // Serves both sides, whichever has a chunk of posteriors ready.  lane = (class of the group, frame): sums the
    // run of cells of its class in 16-byte pieces, finishes the class in place in the frame's row record (emission ->
    // gradient); the warp then stores the rows to global memory, a row per instruction.  One warp does this for the
    // whole CTA, so the code is laid out for latency: the class table lives in registers, the first three pieces of
    // EVERY run are loaded before the first add (a run shorter than that reads cells that hold zeros), only what a
    // run has beyond twelve cells goes through a loop, and the rows leave as eight loads followed by eight stores.
    if (want_grad) {
      const int f = lane & 7, cg = lane >> 3;
      const float PS = 1.8446744e19f;  // 2^PSHIFT
      const float gps = gs * PS;
      const int ngrp = (C - 1 + 3) / 4;
      const bool vec2 = ((shift | C) & 1) == 0 && ((((uintptr_t)gbase) | (rstride * 4)) & 7) == 0;
      constexpr int NG = NGRP;
      int cls[NG], o0[NG], o1[NG], o2[NG];
      int nlong = 0;   // groups whose runs are longer than twelve cells (falling counts: the first nlong)
#pragma unroll
      for (int g = 0; g < NG; g++) {
        const int4 e = s_gtab[4 * g + cg];
        cls[g] = e.x;
        o0[g] = e.z > 0 ? e.y : GCELLS;
        o1[g] = e.z > 4 ? e.y + 4 : GCELLS;
        o2[g] = e.z > 8 ? e.y + 8 : GCELLS;
        nlong += (e.z > 12) ? 1 : 0;
      }
      int done0 = 0, done1 = 0;
      const int todo0 = S.nch - S.nc10, todo1 = S.nch - S.nc11;
#pragma unroll 1
      while (done0 < todo0 || done1 < todo1) {
        int d2 = -1;
        if (done0 < todo0 && flag_get(FLAG(FL_GFULL + 0)) > done0) d2 = 0;
        else if (done1 < todo1 && flag_get(FLAG(FL_GFULL + 1)) > done1) d2 = 1;
        d2 = __shfl_sync(0xffffffffu, d2, 0);
        if (d2 < 0) {
          __nanosleep(100);
          continue;
        }
        const int j2 = d2 ? done1 : done0;
        const int i = S.nc1(d2) + j2;
        const Chunk ci = chunk_at(S, d2, i);
        const int stg = i % NST;
        TS(3 + d2, i, 1);
        const float* G = s_gbuf + (size_t)((d2 * 2 + (j2 & 1)) * KC + f) * GS;
        float* rec = reinterpret_cast<float*>(smem + sl.rows + ((size_t)(d2 * NST + stg) * KC + f) * ROWB);
        const float inv = rec[ROWW - 2];
        const float gy = gs * inv;
        rec += shift;
        auto sums = [&](auto ngc) {
          constexpr int N = decltype(ngc)::value;
          float4 u[N], w[N], x3[N];
          float y[N], acc[N];
#pragma unroll
          for (int g = 0; g < N; g++) {
            u[g] = *reinterpret_cast<const float4*>(G + o0[g]);
            w[g] = *reinterpret_cast<const float4*>(G + o1[g]);
            x3[g] = *reinterpret_cast<const float4*>(G + o2[g]);
            y[g] = rec[max(cls[g], 0)];
          }
#pragma unroll
          for (int g = 0; g < N; g++)
            acc[g] = (((u[g].x + w[g].x) + (u[g].y + w[g].y)) + ((u[g].z + w[g].z) + (u[g].w + w[g].w))) +
                     ((x3[g].x + x3[g].y) + (x3[g].z + x3[g].w));
          for (int g = 0; g < nlong; g++) {   // the same in all lanes
            const int4 e = s_gtab[4 * g + cg];
            const float4* q = reinterpret_cast<const float4*>(G + e.y);
            float x = 0.f;
            for (int n = 3; n < (e.z >> 2); n++) {
              const float4 v = q[n];
              x += (v.x + v.y) + (v.z + v.w);
            }
#pragma unroll
            for (int h = 0; h < N; h++) acc[h] += (h == g) ? x : 0.f;
          }
          float t = 0.f;
#pragma unroll
          for (int g = 0; g < N; g++) {
            t += acc[g];
            if (cls[g] >= 0) rec[cls[g]] = fmaf(y[g], gy, -gps * acc[g]);
          }
          return t;
        };
        float tot = ngrp <= 10 ? sums(std::integral_constant<int, 10>()) : sums(std::integral_constant<int, NG>());
        tot += __shfl_xor_sync(0xffffffffu, tot, 8);
        tot += __shfl_xor_sync(0xffffffffu, tot, 16);
        tot *= PS;
        if (f < ci.len && !(tot < 1.0f + 1e-4f)) alarm |= AL_OCC;
        if (cg == 0) rec[blank] = gs * (rec[blank] * inv - (1.0f - tot));
        __syncwarp();
        TS(3 + d2, i, 2);
        if (lane == 0) flag_set(FLAG(FL_GFREE + d2), j2 + 1);
        // rows -> global, one row per instruction (152 bytes at C = 38: 19 lanes of 8 bytes where the alignment allows)
        const float* src = reinterpret_cast<const float*>(smem + sl.rows + (size_t)(d2 * NST + stg) * KC * ROWB) + shift;
        float* dst = gbase + (size_t)ci.t0 * rstride;
        const long long dstep = (long long)ci.dt * (long long)rstride;
        if (vec2) {
          // the rows go into registers, the stage is released, then the stores drain on their own (the release is a
          // fence: after the stores it would wait for every one of them to be acknowledged)
          float2 v[KC];
          if (2 * lane < C) {
#pragma unroll
            for (int ff = 0; ff < KC; ff++) v[ff] = *reinterpret_cast<const float2*>(src + ff * ROWW + 2 * lane);
          }
          __syncwarp();
          if (lane == 0) flag_set(FLAG(FL_FREED + d2), i + 1);
          if (2 * lane < C) {
#pragma unroll
            for (int ff = 0; ff < KC; ff++)
              if (ff < ci.len) *reinterpret_cast<float2*>(dst + ff * dstep + 2 * lane) = v[ff];
          }
        } else {
#pragma unroll 1
          for (int ff = 0; ff < ci.len; ff++) {
            if (lane < C) dst[lane] = src[lane];
            if (lane + 32 < C) dst[lane + 32] = src[lane + 32];
            src += ROWW;
            dst += dstep;
          }
          __syncwarp();
          if (lane == 0) flag_set(FLAG(FL_FREED + d2), i + 1);
        }
        TS(3 + d2, i, 3);
        if (d2) done1++; else done0++;
      }
    }
  }
  if (alarm) atomicOr(&s_scal[0], alarm);
  __syncthreads();
#if NASR_TUNING
  if (p.prof && tid == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.prof[3700 + 2 * b + 1] = (long long)gt | ((long long)(s_scal[4] & 1) << 62);
  }
#endif
  if (tid == 0) {
    p.retry[b] = s_scal[0];
    atomicSub(&g_sm_arrivals[smid & 1023], 1);
  }
#undef FLAG
}

}  // namespace lean

// ---- host side ------------------------------------------------------------------------------------------------------

namespace {

constexpr int kLeanMaxSmem = 113 * 1024;   // two CTAs per SM

int lean_nl(int Lmax) {
  if (Lmax <= 2 * 32 - 2) return 2;
  if (Lmax <= 4 * 32 - 2) return 4;
  if (Lmax <= 5 * 32 - 2) return 5;
  if (Lmax <= 7 * 32 - 2) return 7;
  if (Lmax <= 8 * 32 - 2) return 8;
  return 0;
}

template <int NL>
int launch_lean(const lean::Params& p, cudaStream_t stream) {
  const lean::Smem sl = lean::smem_layout(NL);
  int dev = 0;
  NASR_CUDA(cudaGetDevice(&dev));
  static std::atomic<bool> attr_set[64];
  if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
    NASR_CUDA(cudaFuncSetAttribute(lean::ctc_lean_kernel<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kLeanMaxSmem));
    // two CTAs of ~100 KB per SM: ask for the largest shared-memory carve-out (the default picks one that fits one)
    NASR_CUDA(cudaFuncSetAttribute(lean::ctc_lean_kernel<NL>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   cudaSharedmemCarveoutMaxShared));
    if (dev >= 0 && dev < 64) attr_set[dev].store(true, std::memory_order_release);
  }
  lean::ctc_lean_kernel<NL><<<p.B, lean::NTHREADS, sl.total, stream>>>(p);
  count_launch();
  NASR_CUDA(cudaGetLastError());
  return NASR_OK;
}

}  // namespace

extern int g_debug_split;
extern long long* g_debug_prof;

bool ctc_narrow_supported(int T, int C, int Lmax) {
  if (T < 2 * lean::KC || C > 64 || C < 2) return false;
  const int NL = lean_nl(Lmax);
  if (!NL) return false;
  return lean::smem_layout(NL).total <= (size_t)kLeanMaxSmem;
}

size_t ctc_narrow_workspace_bytes(int T, int B, int C, int Lmax) {
  if (!ctc_narrow_supported(T, C, Lmax)) return 0;
  const int NL = lean_nl(Lmax);
  // under a split override one direction may own every chunk
  const int maxch = (T + lean::KC - 1) / lean::KC + 2;
  return (size_t)B * 2 * maxch * (3 * ((NL + 3) / 4) * 32) * sizeof(float4);
}

int ctc_narrow_launch(const float* logits, int T, int B, int C, long long st_t, long long st_b,
                      const int32_t* label_values, const int32_t* label_offsets, int Lmax, const int32_t* seq_len,
                      int blank, float* loss, float* grad, const float* grad_loss, int32_t* status, int32_t* retry,
                      void* ckpt, cudaStream_t stream) {
  lean::Params p;
  p.logits = logits; p.T = T; p.B = B; p.C = C; p.st_t = st_t; p.st_b = st_b;
  p.lab_vals = label_values; p.lab_offs = label_offsets; p.seq_len = seq_len;
  p.blank = blank; p.loss = loss; p.grad = grad; p.grad_loss = grad_loss; p.status = status;
  p.retry = retry;
  p.ckpt = static_cast<float4*>(ckpt);
  p.maxch = (T + lean::KC - 1) / lean::KC + 2;
  p.split = g_debug_split;
  p.lo = reinterpret_cast<const char*>(logits);
  p.hi = reinterpret_cast<const char*>(logits) +
         4 * ((size_t)(T - 1) * st_t + (size_t)(B - 1) * st_b + (size_t)C);
  p.prof = g_debug_prof;
  switch (lean_nl(Lmax)) {
    case 2: return launch_lean<2>(p, stream);
    case 4: return launch_lean<4>(p, stream);
    case 5: return launch_lean<5>(p, stream);
    case 7: return launch_lean<7>(p, stream);
    case 8: return launch_lean<8>(p, stream);
  }
  set_error("nasr_ctc: narrow kernel does not support max_label_len=%d", Lmax);
  return NASR_ERR_UNSUPPORTED;
}

// every row of an utterance must start at the same offset inside its 16-byte aligned superset (the recursion's gather
// offsets are per utterance): the frame stride has to be a multiple of 16 bytes, rows 4-byte aligned (always)
bool ctc_narrow_layout_ok(const float* logits, long long st_t, long long st_b) {
  (void)logits;
  (void)st_b;
  return ((st_t * 4) & 15) == 0;
}

}  // namespace nasr
