// CTC loss + d(loss)/d(logits) for narrow vocabularies (C <= 64, transcripts up to 254 labels) on sm_100a:
// one CTA of seven specialised warps per utterance, one launch per batch, float32 arithmetic.
//
// Replaces tf.nn.ctc_loss + _CTCLossGrad behind create_loss (reference networks/tfnetwork.py:58-59); semantics per
// SURVEY.md Appendix A.1.  ctc_loss.cu holds the robust kernel that redoes any utterance flagged in retry[].
//
// Arithmetic (restated on the CPU by tests/model_f32.py and checked there against the oracle):
//   * linear-domain recursion in "ratio units", R[t][c] = exp(x[t][c] - x[t][blank]): blank states need no multiply,
//     log p gets sum_t log y_blank(t) added back;
//   * slot i = (blank state 2(i-1), label state 2(i-1)+1); lane l of a recursion warp owns the NL consecutive slots
//     [l*NL, (l+1)*NL) as NL/2 packed pairs (slot j, slot j+NL/2), so that one add/mul/fma.rn.f32x2 advances two slots;
//     the s-1/s-2 neighbours of a lane's first slot arrive by ONE 32-bit shuffle per frame;
//   * the backward recursion is the same code on the reversed label string over descending frames;
//   * float32 values with one integer exponent PER SLOT: true = stored * 2^E[k].  What crosses from slot k-1 into
//     slot k is multiplied by F[k] = 2^(E[k-1]-E[k]) inside the fused multiply-add that consumes it.  Every KC = 8
//     frames each slot is renormalised (larger state -> 2^(TB-127)) subject to E[k] >= E[k-1] - GCAP, so F <= 2^GCAP;
//     label values saturate at 2^90, hence nothing can reach inf and no NaN can arise from 0 * inf;
//   * meet in the middle: the forward warp covers frames [0,M) while the backward warp covers [M,Tb); p comes from
//     the two at the meeting point; then each continues through the other half and multiplies its pre-emission sums
//     with the other direction's values, which a recompute warp regenerates chunk by chunk from checkpoints
//     (the [T,U] lattice never exists in memory);
//   * posterior of a label state = pre-emission sum * other direction's value / p, computed as (q*c1)*(o*c2) with the
//     power of two 2^(E+Eo-e_p-64) split over both factors; the blank's occupancy is 1 - sum of label occupancies;
//   * a-posteriori certificate: saturation and underflow only ever REMOVE mass, so alpha(T-1)[final states]/p and
//     beta(0)[first states]/p fall short of 1 by exactly the posterior mass that was lost.  Both must be within 3e-5
//     of 1, no frame's label occupancy may exceed 1, else retry[b] hands the utterance to the robust kernel.
//
// Data movement: a frame's logits row (4*C bytes, 8-byte aligned at C = 38) travels global -> shared memory as ONE
// bulk copy (cp.async.bulk, the TMA's 1-D form) of the 16-byte aligned superset of the row, completing on an mbarrier;
// the producer warp converts 32 rows at a time, one row per lane, in place.  All hand-offs between warps are
// mbarrier producer/consumer rings -- there is no CTA-wide barrier in the main loop.
//
//   warp 0,1  recursion (forward / backward)      warp 4    producer (bulk copies, softmax pieces)
//   warp 2,3  recompute of the other direction     warp 5,6  gradient (class reduction, coalesced row stores)
#include <math.h>

#include "nasr_common.cuh"

#ifndef NASR_TUNING
#define NASR_TUNING 0
#endif

namespace nasr {
namespace narrow {

typedef unsigned long long u64;

constexpr int KC = 8;             // frames per chunk (= renormalisation and checkpoint interval)
constexpr int NST = 8;            // chunks of row records per side in the ring
constexpr int NTHREADS = 224;     // 7 warps
constexpr int TB = 60;            // biased exponent a slot's larger state is brought to (2^-67)
constexpr int GCAP = 30;          // a slot with mass sits at most this far below the slot with mass beneath it
constexpr int PSHIFT = 64;        // the posterior buffer holds posterior * 2^-PSHIFT
constexpr int ENEG = -(1 << 28);  // exponent tag of a slot that can never receive mass
// Posterior rows.  The labels of a class are cut into quads of four; the quads of all classes, in class order, are dealt
// column-major onto a table of nrows x 4 groups, and element e of the quad in (row r, group g) lives in cell
// 4 * (4 r + e) + g of the frame's posterior row.  A gradient lane (frame, group) therefore reads one word per element
// at a stride of four words, the 32 lanes of a warp hit 32 banks (row length / 4 is odd), control flow is the same in
// every lane, and a class is finished by at most a few read-modify-writes of its gradient cell.
__host__ __device__ constexpr int kMaxRows(int NL) { return NL <= 4 ? 18 : 23; }
__host__ __device__ constexpr int kGrow(int NL) { return 16 * kMaxRows(NL) + 4; }   // words per posterior row (+4: dump cells)
constexpr int MAXROWS = 24;
#define NASR_BIG 1.2379400e27f    // 2^90: with F <= 2^30 eight frames of inflow stay below 2^127, so nothing ever reaches inf
constexpr float kTol = 3e-5f;

enum Role { R_F = 0, R_B = 1, RC_F = 2, RC_B = 3, PROD = 4, G_F = 5, G_B = 6 };
enum Alarm { AL_SHAPE = 1, AL_EMISSION = 2, AL_NONFINITE = 4, AL_P = 32, AL_CERT = 128, AL_OCC = 256 };

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
  float4* ckpt;            // [B][2][maxch][3*NL/4][32]
  int maxch;
  int roww;                // words per row record
  int split;               // debug: frames of the forward half (multiple of KC), 0 = automatic
  const char* lo;          // lowest byte of the logits view
  const char* hi;          // one past its highest byte
  int* dbg;                // tuning builds: progress markers [B][8 warps][4] (may be mapped host memory)
};

// ---- mbarrier / bulk copy ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sptr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t a, int n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(n) : "memory");
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
__device__ __forceinline__ bool mbar_test(uint32_t a, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(a), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}

// ---- packed float32 pairs -------------------------------------------------------------------------------------
__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^e for e in (-inf, 127]; 0 below 2^-126
__device__ __forceinline__ float pow2f(int e) { return __int_as_float(max(e + 127, 0) << 23); }

// ---- shared memory ----------------------------------------------------------------------------------------------
__host__ __device__ constexpr size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

struct Smem {
  size_t bars, rows, obuf, oexp, gbuf, meet, lab, cell, cnt, rowtab, clstab, lsum, scal, total;
};
// barrier indices (each an 8-byte mbarrier)
constexpr int B_RAW = 0;                   // [2][NST] raw rows landed (tx count)
constexpr int B_ROW = B_RAW + 2 * NST;     // [2][NST] row records ready
constexpr int B_EMP = B_ROW + 2 * NST;     // [2][NST] row records released by their three readers
constexpr int B_OFULL = B_EMP + 2 * NST;   // [2][2]
constexpr int B_OEMP = B_OFULL + 4;
constexpr int B_GFULL = B_OEMP + 4;
constexpr int B_GEMP = B_GFULL + 4;
constexpr int B_PH1 = B_GEMP + 4;          // [2] direction d finished phase 1 (its checkpoints are visible)
constexpr int B_MEETB = B_PH1 + 2;
constexpr int B_MEETP = B_MEETB + 1;
constexpr int NBARS = B_MEETP + 1;

__host__ __device__ inline Smem smem_layout(int NL, int roww) {
  Smem s;
  const int N = NL * 32;
  size_t o = 0;
  s.bars = o;    o = al16(o + (size_t)NBARS * 8);
  s.rows = o;    o = al16(o + (size_t)2 * NST * KC * roww * 4);
  s.obuf = o;    o = al16(o + (size_t)2 * 2 * KC * N * 4);
  s.oexp = o;    o = al16(o + (size_t)2 * 2 * N * 4);
  s.gbuf = o;    o = al16(o + (size_t)2 * 2 * KC * kGrow(NL) * 4);
  s.meet = o;    o = al16(o + (size_t)3 * N * 4);
  s.lab = o;     o = al16(o + (size_t)N * 4);
  s.cell = o;    o = al16(o + (size_t)N * 2);
  s.cnt = o;     o = al16(o + (size_t)(64 + 2) * 4 * 3);     // per class: count, first quad, (spare)
  s.rowtab = o;  o = al16(o + (size_t)4 * 4);
  s.clstab = o;  o = al16(o + (size_t)MAXROWS * 4 * 4);
  s.lsum = o;    o = al16(o + 32 * 4);
  s.scal = o;    o = al16(o + 64);
  s.total = o;
  return s;
}

// ---- schedule -----------------------------------------------------------------------------------------------------
struct Sched {
  int Tb;
  int n1[2];    // frames each direction covers in phase 1
  int nc1[2];   // chunks of phase 1
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
  if (i < S.nc1[d]) {
    const int tau0 = i * KC;
    c.phase = 1;
    c.jo = i;
    c.len = min(KC, S.n1[d] - tau0);
    c.t0 = d ? S.Tb - 1 - tau0 : tau0;
    c.dt = d ? -1 : 1;
  } else if (S.grad) {
    const int o = d ^ 1;
    const int jo = S.nc1[o] - 1 - (i - S.nc1[d]);
    const int hi_tau = min(jo * KC + KC, S.n1[o]);
    c.phase = 2;
    c.jo = jo;
    c.len = hi_tau - jo * KC;
    // consumer position f <-> other direction's position hi_tau-1-f
    c.t0 = o ? S.Tb - 1 - (hi_tau - 1) : hi_tau - 1;
    c.dt = o ? 1 : -1;
  }
  return c;
}

// ---- recursion state ----------------------------------------------------------------------------------------------
// SW: the halves of a pair are swapped (lo = slot j+H, hi = slot j): the recompute warps run this layout so that the
// 64-bit word they store for pair H-1-j holds exactly the two values the consumer's pair j needs, in its order.
template <int NL>
struct St {
  static constexpr int H = NL / 2;
  u64 A[H], B[H], F[H], SF[H];
  int E[NL];
  uint32_t coloff[NL];   // byte offset of R[class of slot k] in a row record (dead slot: the zero entry)
  uint32_t skipmask;     // bit k: slot k may take the skip transition
};
template <int NL, bool SW>
__device__ __forceinline__ constexpr int slot_of(int j, int h) { return (h == (SW ? 1 : 0)) ? j : j + NL / 2; }
template <int NL, bool SW>
__device__ __forceinline__ float get(const u64 (&V)[NL / 2], int k) {
  constexpr int H = NL / 2;
  const int j = k < H ? k : k - H;
  const bool is_lo = SW ? (k >= H) : (k < H);
  return is_lo ? lo(V[j]) : hi(V[j]);
}
template <int NL, bool SW>
__device__ __forceinline__ void set(u64 (&V)[NL / 2], int k, float x) {
  constexpr int H = NL / 2;
  const int j = k < H ? k : k - H;
  const bool is_lo = SW ? (k >= H) : (k < H);
  V[j] = is_lo ? pk(x, hi(V[j])) : pk(lo(V[j]), x);
}

// Transfer factors from the slot exponents: F[k] = 2^(E[k-1]-E[k]) (<= 2^GCAP), 0 where either slot is dead.
template <int NL, bool SW>
__device__ __forceinline__ void set_F(St<NL>& s, int lane) {
  int Ep = __shfl_up_sync(0xffffffffu, s.E[NL - 1], 1);
  if (lane == 0) Ep = ENEG;
  float f[NL], sf[NL];
#pragma unroll
  for (int k = 0; k < NL; k++) {
    const int lower = k ? s.E[k - 1] : Ep;
    const bool dead = lower <= ENEG / 2 || s.E[k] <= ENEG / 2;
    f[k] = dead ? 0.f : pow2f(min(lower - s.E[k], GCAP));
    sf[k] = ((s.skipmask >> k) & 1u) ? f[k] : 0.f;
  }
#pragma unroll
  for (int j = 0; j < NL / 2; j++) {
    s.F[j] = pk(f[slot_of<NL, SW>(j, 0)], f[slot_of<NL, SW>(j, 1)]);
    s.SF[j] = pk(sf[slot_of<NL, SW>(j, 0)], sf[slot_of<NL, SW>(j, 1)]);
  }
}

// Renormalise every slot (larger state -> biased exponent TB), keeping E[k] >= E[k-1] - GCAP over slots with mass;
// empty slots adopt the exponent below.  Across lanes the chain starts from the neighbour's exponent of the previous
// rescale and is raised twice against the neighbour's fresh value (one shuffle each, no scan over lanes); what a
// longer-range raise would have changed is left to the clamp of F and, if it mattered, to the certificate.
template <int NL, bool SW>
__device__ __forceinline__ void rescale(St<NL>& s, int lane, int& alarm) {
  constexpr int H = NL / 2;
  int En[NL];
  uint32_t nzmask = 0;
#pragma unroll
  for (int k = 0; k < NL; k++) {
    const uint32_t m = max(__float_as_uint(get<NL, SW>(s.A, k)), __float_as_uint(get<NL, SW>(s.B, k)));
    int ex = (int)(m >> 23);
    if (ex >= 255) alarm |= AL_NONFINITE;
    // a slot whose values have decayed into the denormals still holds mass: take its exponent from m * 2^64
    // (adopting the neighbour's exponent instead would scale the residue up by an arbitrary power of two)
    const int exd = (int)(__float_as_uint(__uint_as_float(m) * 1.8446744e19f) >> 23) - 64;
    ex = ex ? ex : exd;
    const bool nz = m != 0u && s.E[k] > ENEG / 2;
    nzmask |= nz ? (1u << k) : 0u;
    En[k] = s.E[k] + ex - TB;
  }
  int prev = __shfl_up_sync(0xffffffffu, s.E[NL - 1], 1);
#pragma unroll
  for (int ps = 0; ps < 3; ps++) {
    if (lane == 0) prev = ENEG;
#pragma unroll
    for (int k = 0; k < NL; k++) {
      prev = ((nzmask >> k) & 1u) ? max(En[k], prev - GCAP) : prev;
      En[k] = prev;
    }
    prev = __shfl_up_sync(0xffffffffu, En[NL - 1], 1);
  }
  float s1[NL], s2[NL];
#pragma unroll
  for (int k = 0; k < NL; k++) {
    const bool alive = En[k] > ENEG / 2 && s.E[k] > ENEG / 2;
    const int d = alive ? max(-252, min(252, s.E[k] - En[k])) : 0;
    const int h = d >> 1;
    s1[k] = pow2f(h);
    s2[k] = pow2f(d - h);
    s.E[k] = En[k] > ENEG / 2 ? En[k] : ENEG;
  }
#pragma unroll
  for (int j = 0; j < H; j++) {
    const u64 S1 = pk(s1[slot_of<NL, SW>(j, 0)], s1[slot_of<NL, SW>(j, 1)]);
    const u64 S2 = pk(s2[slot_of<NL, SW>(j, 0)], s2[slot_of<NL, SW>(j, 1)]);
    s.A[j] = mul2(mul2(s.A[j], S1), S2);
    s.B[j] = mul2(mul2(s.B[j], S1), S2);
  }
  set_F<NL, SW>(s, lane);
}

enum Mode { PLAIN = 0, STORE_O = 1, COMBINE = 2 };

// One frame.  erow: the frame's row record; orow: the frame's 64-bit words of the other direction ([pair][lane]);
// grow: the frame's posterior row.  All loads are issued before any store.
template <int NL, bool SW, int MODE>
__device__ __forceinline__ void step(St<NL>& s, const unsigned char* erow, u64* orow, unsigned char* grow,
                                     const u64 (&C1)[NL / 2], const u64 (&C2)[NL / 2], const uint32_t (&gph)[NL],
                                     int lane) {
  constexpr int H = NL / 2;
  u64 rh[H], oh[H];
#pragma unroll
  for (int j = 0; j < H; j++)
    rh[j] = pk(*reinterpret_cast<const float*>(erow + s.coloff[slot_of<NL, SW>(j, 0)]),
               *reinterpret_cast<const float*>(erow + s.coloff[slot_of<NL, SW>(j, 1)]));
  if (MODE == COMBINE) {
#pragma unroll
    for (int j = 0; j < H; j++) oh[j] = orow[(H - 1 - j) * 32 + (31 - lane)];
  }
  // label value of the neighbour lane's last slot (lane 0 multiplies it by F = 0)
  const float a_in = __shfl_up_sync(0xffffffffu, SW ? lo(s.A[H - 1]) : hi(s.A[H - 1]), 1);
  u64 alp[H];
  alp[0] = SW ? pk(hi(s.A[H - 1]), a_in) : pk(a_in, lo(s.A[H - 1]));
#pragma unroll
  for (int j = 1; j < H; j++) alp[j] = s.A[j - 1];
#pragma unroll
  for (int j = 0; j < H; j++) {
    const u64 nb = fma2(s.F[j], alp[j], s.B[j]);
    const u64 q = fma2(s.SF[j], alp[j], add2(s.A[j], s.B[j]));
    if (MODE == COMBINE) {
      const u64 po = mul2(mul2(q, C1[j]), mul2(oh[j], C2[j]));
      *reinterpret_cast<float*>(grow + gph[slot_of<NL, SW>(j, 0)]) = lo(po);
      *reinterpret_cast<float*>(grow + gph[slot_of<NL, SW>(j, 1)]) = hi(po);
    }
    const u64 an = mul2(q, rh[j]);
    s.A[j] = pk(fminf(lo(an), NASR_BIG), fminf(hi(an), NASR_BIG));
    s.B[j] = nb;
    if (MODE == STORE_O) orow[j * 32 + lane] = s.A[j];
  }
}

template <int NL, bool SW, int MODE>
__device__ __forceinline__ void run_chunk(St<NL>& s, const unsigned char* erows, int rowbytes, int len, bool reverse,
                                          u64* obuf, float* gbuf, const u64 (&C1)[NL / 2], const u64 (&C2)[NL / 2],
                                          const uint32_t (&gph)[NL], int lane) {
  constexpr int OROW = NL * 16;  // 64-bit words per frame of the other direction's values
  constexpr int GROW = kGrow(NL);
  if (len == KC) {
#pragma unroll
    for (int g = 0; g < KC; g++) {
      const int f = reverse ? KC - 1 - g : g;
      step<NL, SW, MODE>(s, erows + f * rowbytes, obuf + f * OROW, reinterpret_cast<unsigned char*>(gbuf + f * GROW), C1, C2,
                         gph, lane);
    }
  } else {
#pragma unroll 1
    for (int g = 0; g < len; g++) {
      const int f = reverse ? len - 1 - g : g;
      step<NL, SW, MODE>(s, erows + f * rowbytes, obuf + f * OROW, reinterpret_cast<unsigned char*>(gbuf + f * GROW), C1, C2,
                         gph, lane);
    }
  }
}

// Slot tables for direction d (0 forward, 1 mirrored) and the virtual row before the first frame.
template <int NL, bool SW>
__device__ __forceinline__ void dir_setup(St<NL>& s, int d, int lane, const int* lab, int L, int zero_off, int shift,
                                          bool init_state) {
  constexpr int N = NL * 32;
  const int pad = N - L - 1;
  const int first = d == 0 ? 1 : pad;
  float ab[NL];
  s.skipmask = 0;
#pragma unroll
  for (int k = 0; k < NL; k++) {
    const int i = lane * NL + k;
    int col = -1;
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
    s.coloff[k] = col >= 0 ? (uint32_t)(shift + col) * 4u : (uint32_t)zero_off;
    s.skipmask |= skip ? (1u << k) : 0u;
    ab[k] = (i == first) ? __int_as_float(TB << 23) : 0.f;
    s.E[k] = i >= first ? 127 - TB : ENEG;
  }
  if (init_state) {
#pragma unroll
    for (int j = 0; j < NL / 2; j++) {
      s.B[j] = pk(ab[slot_of<NL, SW>(j, 0)], ab[slot_of<NL, SW>(j, 1)]);
      s.A[j] = pk(0.f, 0.f);
    }
    set_F<NL, SW>(s, lane);
  }
}

template <int NL>
__device__ __forceinline__ float4* ckpt_ptr(const Params& p, int b, int d, int c) {
  return p.ckpt + (((size_t)b * 2 + d) * p.maxch + c) * (size_t)(3 * NL / 4 * 32);
}

template <int NL>
__global__ void __launch_bounds__(NTHREADS, (NL <= 8 ? 2 : 1)) ctc_narrow_kernel(const Params p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int N = NL * 32;
  constexpr int H = NL / 2;
  constexpr int Lcap = N - 2;
  constexpr int GROW = kGrow(NL);
  constexpr int MAXR = kMaxRows(NL);
  const Smem sl = smem_layout(NL, p.roww);
  u64* s_bar = reinterpret_cast<u64*>(smem + sl.bars);
  unsigned char* s_rows = smem + sl.rows;
  u64* s_obuf = reinterpret_cast<u64*>(smem + sl.obuf);
  int* s_oexp = reinterpret_cast<int*>(smem + sl.oexp);
  float* s_gbuf = reinterpret_cast<float*>(smem + sl.gbuf);
  float* s_meet_nb = reinterpret_cast<float*>(smem + sl.meet);
  float* s_meet_pre = s_meet_nb + N;
  int* s_meet_e = reinterpret_cast<int*>(s_meet_pre + N);
  int* s_lab = reinterpret_cast<int*>(smem + sl.lab);
  uint16_t* s_cell = reinterpret_cast<uint16_t*>(smem + sl.cell);
  int* s_cnt = reinterpret_cast<int*>(smem + sl.cnt);
  int* s_qbase = s_cnt + 66;                                   // first quad of every class
  int* s_rowtab = reinterpret_cast<int*>(smem + sl.rowtab);   // [0]: number of table rows
  int* s_clstab = reinterpret_cast<int*>(smem + sl.clstab);   // [r][g]: class of group g in row r, -1 if none
  float* s_lsum = reinterpret_cast<float*>(smem + sl.lsum);
  int* s_scal = reinterpret_cast<int*>(smem + sl.scal);       // [0] alarm [1] e_p [2] repeats [3] 1/m_p (float bits)

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = p.T, C = p.C, blank = p.blank;
  const int roww = p.roww, rowbytes = roww * 4;
  const uint32_t bar0 = sptr(s_bar);
#define BAR(i) (bar0 + 8u * (uint32_t)(i))
#define MARK(slot, val)                                                                                  \
  do {                                                                                                   \
    if (NASR_TUNING && p.dbg && lane == 0) {                                                             \
      *reinterpret_cast<volatile int*>(p.dbg + ((size_t)b * 8 + warp) * 4 + (slot)) = (val);             \
      __threadfence_system();                                                                            \
    }                                                                                                    \
  } while (0)

  const int Tb = p.seq_len[b];
  const int l0 = p.lab_offs[b];
  const int L = p.lab_offs[b + 1] - l0;

  // ---- can this kernel take the utterance? everything unusual goes to the robust kernel -----------------------
  int bad = (Tb < 2 * KC) | (Tb > T) | (L < 0) | (L > Lcap) | (C > 64);
  if (tid < 8) s_scal[tid] = 0;
  for (int c = tid; c < 66 * 3; c += NTHREADS) s_cnt[c] = 0;
  if (tid < 32) s_lsum[tid] = 0.f;
  __syncthreads();
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
  // ---- class table of the gradient warps (see kGrow): quads of every class in class order, dealt column-major
  if (!bad && warp == 0) {
    for (int i = lane; i < MAXROWS * 4; i += 32) s_clstab[i] = -1;
    // exclusive prefix of the quad counts over classes (one class per lane and pass)
    int run = 0;
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int c = c0 + lane;
      const int nq = (c < C && c != blank) ? (s_cnt[c] + 3) >> 2 : 0;
      int inc = nq;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      if (c < C) s_qbase[c] = run + inc - nq;
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
    const int nrows = (run + 3) >> 2;
    __syncwarp();
    // a class whose quads wrap around a whole column would meet itself in one table row (two lanes updating one
    // gradient cell in the same instruction): such transcripts, and ones with more quads than the table has room
    // for, go to the robust kernel
    int over = nrows > MAXR ? 1 : 0;
    for (int c = lane; c < C; c += 32)
      if (c != blank) {
        const int nq = (s_cnt[c] + 3) >> 2;
        if (nq > nrows) over = 1;
        for (int q = 0; q < nq && !over; q++) {
          const int Q = s_qbase[c] + q;
          s_clstab[(Q % nrows) * 4 + Q / nrows] = c;
        }
      }
    over = __any_sync(0xffffffffu, over);
    if (lane == 0) {
      s_rowtab[0] = nrows;
      if (over) s_scal[0] = AL_SHAPE;
    }
  }
  __syncthreads();
  if (bad || s_scal[0]) {
    if (tid == 0) p.retry[b] = AL_SHAPE;
    return;
  }
  {
    const int nrows = max(1, s_rowtab[0]);
    for (int j = tid; j < L; j += NTHREADS) {
      const int v = s_lab[j];
      int r = 0;
      for (int i = 0; i < j; i++) r += (s_lab[i] == v);
      const int Q = s_qbase[v] + (r >> 2);
      s_cell[j] = (uint16_t)(4 * (4 * (Q % nrows) + (r & 3)) + Q / nrows);
    }
  }
  // posterior rows: padding cells are never written and must read as zero
  for (int i = tid; i < 2 * 2 * KC * GROW; i += NTHREADS) s_gbuf[i] = 0.f;
  if (tid == 0) {
    for (int i = 0; i < 2 * NST; i++) {
      mbar_init(BAR(B_RAW + i), KC);
      mbar_init(BAR(B_ROW + i), 1);
      mbar_init(BAR(B_EMP + i), 3);
    }
    for (int i = 0; i < 4; i++) {
      mbar_init(BAR(B_OFULL + i), 1);
      mbar_init(BAR(B_OEMP + i), 1);
      mbar_init(BAR(B_GFULL + i), 1);
      mbar_init(BAR(B_GEMP + i), 1);
    }
    mbar_init(BAR(B_PH1), 1);
    mbar_init(BAR(B_PH1 + 1), 1);
    mbar_init(BAR(B_MEETB), 1);
    mbar_init(BAR(B_MEETP), 1);
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
    S.n1[0] = nf;
    S.n1[1] = Tb - nf;
    S.nc1[0] = (S.n1[0] + KC - 1) / KC;
    S.nc1[1] = (S.n1[1] + KC - 1) / KC;
    S.nch = want_grad ? S.nc1[0] + S.nc1[1] : max(S.nc1[0], S.nc1[1]);
    S.grad = want_grad ? 1 : 0;
  }
  // every row of this utterance starts at the same offset inside its 16-byte aligned superset (the launcher checked)
  const float* xbase = p.logits + (size_t)b * p.st_b;
  const int shift = (int)(((uintptr_t)xbase & 15) >> 2);
  const int zero_off = (roww - 1) * 4;
  __syncthreads();

  const int role = warp;
  const int d = role & 1;  // direction / side this warp works for (producer: unused)
  int alarm = 0;

  if (role == R_F || role == R_B) {
    // ================================ recursion warps =================================
    St<NL> st;
    dir_setup<NL, false>(st, d, lane, s_lab, L, zero_off, shift, true);
    uint32_t gph[NL];
    {
      const int pad = N - L - 1;
#pragma unroll
      for (int k = 0; k < NL; k++) {
        const int i = lane * NL + k;
        int cell = 16 * MAXR + (k & 3);  // dump cell of slots without a label
        if (d == 0) {
          const int j = i - 1;
          if (j >= 0 && j < L) cell = s_cell[j];
        } else {
          const int m = i - pad;
          if (m >= 0 && m < L) cell = s_cell[L - 1 - m];
        }
        gph[k] = (uint32_t)cell * 4u;
      }
    }
    u64 C1[H], C2[H];
#pragma unroll
    for (int j = 0; j < H; j++) C1[j] = C2[j] = 0ull;
    int e_p = 0;
#pragma unroll 1
    for (int i = 0; i < S.nch; i++) {
      const Chunk ci = chunk_at(S, d, i);
      if (ci.phase == 0) break;   // loss only: this side has fewer phase-1 chunks than the other
      const int stg = i % NST;
      MARK(0, i); MARK(1, 1);
      mbar_wait(BAR(B_ROW + d * NST + stg), (i / NST) & 1);
      MARK(1, 2);
      const unsigned char* erows = s_rows + (size_t)(d * NST + stg) * KC * rowbytes;
      if (ci.phase == 1) {
        if (i) rescale<NL, false>(st, lane, alarm);
        if (want_grad) {
          float4* ck = ckpt_ptr<NL>(p, b, d, i);
#pragma unroll
          for (int v = 0; v < NL / 4; v++) {
            ck[v * 32 + lane] = make_float4(get<NL, false>(st.B, 4 * v), get<NL, false>(st.B, 4 * v + 1),
                                            get<NL, false>(st.B, 4 * v + 2), get<NL, false>(st.B, 4 * v + 3));
            ck[(NL / 4 + v) * 32 + lane] = make_float4(get<NL, false>(st.A, 4 * v), get<NL, false>(st.A, 4 * v + 1),
                                                       get<NL, false>(st.A, 4 * v + 2), get<NL, false>(st.A, 4 * v + 3));
            reinterpret_cast<int4*>(ck)[(2 * (NL / 4) + v) * 32 + lane] =
                make_int4(st.E[4 * v], st.E[4 * v + 1], st.E[4 * v + 2], st.E[4 * v + 3]);
          }
        }
        run_chunk<NL, false, PLAIN>(st, erows, rowbytes, ci.len, false, s_obuf, s_gbuf, C1, C2, gph, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive_n(BAR(B_EMP + d * NST + stg), 3);
        if (i == S.nc1[d] - 1) {
          // ---- end of phase 1: checkpoints visible to the recompute warp, then the meeting ----
          __threadfence_block();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_PH1 + d));
          rescale<NL, false>(st, lane, alarm);
          if (d == 1) {
            // pre-emission sums of the frame below the meeting point, for the forward warp
            const float a_in = __shfl_up_sync(0xffffffffu, hi(st.A[H - 1]), 1);
#pragma unroll
            for (int k = 0; k < NL; k++) {
              const float alp = k ? get<NL, false>(st.A, k - 1) : a_in;
              const float f = get<NL, false>(st.F, k), sf = get<NL, false>(st.SF, k);
              const float ab = get<NL, false>(st.B, k), al = get<NL, false>(st.A, k);
              s_meet_nb[k * 32 + lane] = fmaf(f, alp, ab);
              s_meet_pre[k * 32 + lane] = fmaf(sf, alp, al + ab);
              s_meet_e[k * 32 + lane] = st.E[k];
            }
            if (NASR_TUNING && p.dbg && b == 0) {
#pragma unroll
              for (int k = 0; k < NL; k++) {
                p.dbg[3000 + k * 32 + lane] = __float_as_int(get<NL, false>(st.A, k));
                p.dbg[3000 + 256 + k * 32 + lane] = __float_as_int(get<NL, false>(st.B, k));
                p.dbg[3000 + 512 + k * 32 + lane] = st.E[k];
                p.dbg[3000 + 768 + k * 32 + lane] = __float_as_int(get<NL, false>(st.F, k));
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B_MEETB));
            MARK(1, 3);
            mbar_wait(BAR(B_MEETP), 0);
            MARK(1, 4);
          } else {
            MARK(1, 5);
            mbar_wait(BAR(B_MEETB), 0);
            MARK(1, 6);
            // ---- p = sum over states of alpha(M-1) * beta(M-1); label of slot i pairs with mirrored slot N-1-i,
            // the blank of slot i (state 2(i-1)) with the blank of mirrored slot N-i ----
            const float S64 = 1.8446744e19f;  // 2^64: both factors sit near 2^-67
            const int lm = 31 - lane;
            float term[2 * NL];
            int kt[2 * NL];
            int kmax = INT_MIN / 2;
#pragma unroll
            for (int k = 0; k < NL; k++) {
              const int mk = NL - 1 - k;
              const int Eo = s_meet_e[mk * 32 + lm];
              const float tl = (get<NL, false>(st.A, k) * S64) * (s_meet_pre[mk * 32 + lm] * S64);
              const bool okl = tl > 0.f && st.E[k] > ENEG / 2 && Eo > ENEG / 2;
              term[k] = okl ? tl : 0.f;
              kt[k] = st.E[k] + Eo;
              // blank partner: global mirrored slot N - i = (lane', slot') one above the label partner
              const int gi = N - (lane * NL + k);
              const int bl = gi / NL, bk = gi % NL;
              float tb = 0.f;
              int Eb = ENEG;
              if (gi < N) {
                Eb = s_meet_e[bk * 32 + bl];
                tb = (get<NL, false>(st.B, k) * S64) * (s_meet_nb[bk * 32 + bl] * S64);
              }
              const bool okb = tb > 0.f && st.E[k] > ENEG / 2 && Eb > ENEG / 2;
              term[NL + k] = okb ? tb : 0.f;
              kt[NL + k] = st.E[k] + Eb;
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
            if (NASR_TUNING && p.dbg && b == 0) {
#pragma unroll
              for (int k = 0; k < 2 * NL; k++) {
                p.dbg[64 + k * 32 + lane] = __float_as_int(term[k]);
                p.dbg[64 + 512 + k * 32 + lane] = kt[k];
              }
              p.dbg[64 + 1024 + lane] = __float_as_int(tot);
#pragma unroll
              for (int k = 0; k < NL; k++) {
                p.dbg[64 + 1088 + k * 32 + lane] = __float_as_int(get<NL, false>(st.A, k));
                p.dbg[64 + 1088 + 256 + k * 32 + lane] = __float_as_int(s_meet_pre[(NL - 1 - k) * 32 + lm]);
                p.dbg[64 + 1088 + 512 + k * 32 + lane] = st.E[k];
                p.dbg[64 + 1088 + 768 + k * 32 + lane] = s_meet_e[(NL - 1 - k) * 32 + lm];
              }
            }
            tot = warp_sum(tot);
            MARK(2, kmax); MARK(3, __float_as_int(tot));
            if (!(tot > 0.f) || !(tot < 3e38f) || kmax < -(1 << 27)) {
              alarm |= AL_P;
              tot = 1.f;
              kmax = 128;
            }
            const int et = (int)((__float_as_uint(tot) >> 23) & 255u) - 127;
            const float mp = tot * pow2f(-et);
            double ls = 0.0;
#pragma unroll 1
            for (int q = 0; q < 32; q++) ls += (double)s_lsum[q];
            if (lane == 0) {
              s_scal[1] = kmax - 128 + et;
              s_scal[3] = __float_as_int(1.0f / mp);
              // log p = (kmax - 128) ln 2 + ln(tot); sum_t log y_blank(t) = -ln 2 * sum_t log2(sum_c R)
              p.loss[b] = (float)(-((double)(kmax - 128) * 0.6931471805599453 + log((double)tot) - ls * 0.6931471805599453));
              p.status[b] = 0;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B_MEETP));
          }
          __syncwarp();
          e_p = s_scal[1];
          const float im = __int_as_float(s_scal[3]);
          const u64 IM = pk(im, im);
#pragma unroll
          for (int j = 0; j < H; j++) {
            st.A[j] = mul2(st.A[j], IM);
            st.B[j] = mul2(st.B[j], IM);
          }
        }
      } else {
        // ---- phase 2: continue through the other half against the other direction's recomputed values ----
        const int j2 = i - S.nc1[d];
        const int buf = j2 & 1;
        rescale<NL, false>(st, lane, alarm);
        MARK(1, 7);
        mbar_wait(BAR(B_OFULL + d * 2 + buf), (j2 >> 1) & 1);
        MARK(1, 8);
        mbar_wait(BAR(B_GEMP + d * 2 + buf), ((j2 >> 1) & 1) ^ 1);
        MARK(1, 9);
        {
          const int* oe = s_oexp + (size_t)(d * 2 + buf) * N;
          float c1[NL], c2[NL];
#pragma unroll
          for (int k = 0; k < NL; k++) {
            const int Eo = oe[(NL - 1 - k) * 32 + (31 - lane)];
            const bool live = st.E[k] > ENEG / 2 && Eo > ENEG / 2;
            int kk = st.E[k] + Eo - e_p - PSHIFT;
            const bool zero = !live || kk < -252;
            kk = max(-252, min(252, kk));
            const int h = kk >> 1;
            c1[k] = zero ? 0.f : pow2f(h);
            c2[k] = pow2f(kk - h);
          }
#pragma unroll
          for (int j = 0; j < H; j++) {
            C1[j] = pk(c1[j], c1[j + H]);
            C2[j] = pk(c2[j], c2[j + H]);
          }
        }
        run_chunk<NL, false, COMBINE>(st, erows, rowbytes, ci.len, false, s_obuf + (size_t)(d * 2 + buf) * KC * (NL * 16),
                                      s_gbuf + (size_t)(d * 2 + buf) * KC * GROW, C1, C2, gph, lane);
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(BAR(B_GFULL + d * 2 + buf));
          mbar_arrive(BAR(B_OEMP + d * 2 + buf));
          mbar_arrive(BAR(B_EMP + d * NST + stg));
        }
      }
    }
    if (want_grad) {
      // ---- certificate: what reached the end of this direction, over p ----
      const int pad = N - L - 1;
      const int sb = d == 0 ? L + 1 : N - 1;   // slot whose blank is the last state of this direction
      const int sl2 = d == 0 ? L : N - 2;      // slot whose label is the state before it
      (void)pad;
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < NL; k++) {
        const int i = lane * NL + k;
        const int sh = max(-126, min(126, st.E[k] - e_p));
        if (i == sb && st.E[k] > ENEG / 2) v += get<NL, false>(st.B, k) * pow2f(sh);
        if (i == sl2 && L > 0 && st.E[k] > ENEG / 2) v += get<NL, false>(st.A, k) * pow2f(sh);
      }
      v = warp_sum(v);
      if (!(fabsf(v - 1.0f) < kTol)) alarm |= AL_CERT;
    }
  } else if (role == RC_F || role == RC_B) {
    // ================================ recompute warps =================================
    // warp RC_x serves side d = x: it recomputes direction d^1 over the frames side d consumes in phase 2
    if (want_grad) {
      const int o = d ^ 1;
      St<NL> st;
      dir_setup<NL, true>(st, o, lane, s_lab, L, zero_off, shift, false);
      u64 C1[H], C2[H];
      uint32_t gph[NL];
#pragma unroll
      for (int j = 0; j < H; j++) C1[j] = C2[j] = 0ull;
#pragma unroll
      for (int k = 0; k < NL; k++) gph[k] = 0;
      // Start only when BOTH directions have finished phase 1: the other one because its checkpoints are read here,
      // this side's own because a parity wait on a ring barrier is only meaningful within one use of the stage -- a
      // recompute warp that ran ahead of its recursion warp by two uses of a stage would see the barrier's parity
      // match one use too early.
      MARK(1, 20);
      mbar_wait(BAR(B_PH1 + o), 0);
      mbar_wait(BAR(B_PH1 + d), 0);
      MARK(1, 21);
#pragma unroll 1
      for (int i = S.nc1[d]; i < S.nch; i++) {
        const Chunk ci = chunk_at(S, d, i);
        const int j2 = i - S.nc1[d];
        const int buf = j2 & 1;
        const int stg = i % NST;
        const float4* ck = ckpt_ptr<NL>(p, b, o, ci.jo);
        float4 vb[NL / 4], va[NL / 4];
        int4 ve[NL / 4];
#pragma unroll
        for (int v = 0; v < NL / 4; v++) {
          vb[v] = __ldcg(ck + v * 32 + lane);
          va[v] = __ldcg(ck + (NL / 4 + v) * 32 + lane);
          ve[v] = __ldcg(reinterpret_cast<const int4*>(ck) + (2 * (NL / 4) + v) * 32 + lane);
        }
        MARK(0, i); MARK(1, 22);
        mbar_wait(BAR(B_OEMP + d * 2 + buf), ((j2 >> 1) & 1) ^ 1);
        MARK(1, 23);
        int* oe = s_oexp + (size_t)(d * 2 + buf) * N;
#pragma unroll
        for (int v = 0; v < NL / 4; v++) {
          set<NL, true>(st.B, 4 * v, vb[v].x); set<NL, true>(st.B, 4 * v + 1, vb[v].y);
          set<NL, true>(st.B, 4 * v + 2, vb[v].z); set<NL, true>(st.B, 4 * v + 3, vb[v].w);
          set<NL, true>(st.A, 4 * v, va[v].x); set<NL, true>(st.A, 4 * v + 1, va[v].y);
          set<NL, true>(st.A, 4 * v + 2, va[v].z); set<NL, true>(st.A, 4 * v + 3, va[v].w);
          st.E[4 * v] = ve[v].x; st.E[4 * v + 1] = ve[v].y; st.E[4 * v + 2] = ve[v].z; st.E[4 * v + 3] = ve[v].w;
        }
#pragma unroll
        for (int k = 0; k < NL; k++) oe[k * 32 + lane] = st.E[k];
        set_F<NL, true>(st, lane);
        MARK(1, 24);
        mbar_wait(BAR(B_ROW + d * NST + stg), (i / NST) & 1);
        MARK(1, 25);
        const unsigned char* erows = s_rows + (size_t)(d * NST + stg) * KC * rowbytes;
        run_chunk<NL, true, STORE_O>(st, erows, rowbytes, ci.len, true, s_obuf + (size_t)(d * 2 + buf) * KC * (NL * 16),
                                     s_gbuf, C1, C2, gph, lane);
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(BAR(B_OFULL + d * 2 + buf));
          mbar_arrive(BAR(B_EMP + d * NST + stg));
        }
      }
    }
  } else if (role == PROD) {
    // ================================ producer warp =================================
    // Four groups of eight lanes: group (side, cip) owns the chunks i = cip, cip + 2, ... of its side, one row per
    // lane.  The groups advance independently -- a side whose ring is full (its recursion warp waits for the other
    // direction at the meeting point) must not hold back the other side -- so the warp runs rounds in which every group
    // does what it can without blocking: issue the bulk copies of its next chunk if the ring stage is free, convert a
    // chunk whose rows have landed (in place: raw logits -> ratio emissions, y_blank, the zero entry).
    const int side = lane >> 4, cip = (lane >> 3) & 1, f = lane & 7;
    const int nv = roww / 4;
    float lsum = 0.f;
    int ni = cip, nc = cip;   // next chunk of this group to issue / to convert
#pragma unroll 1
    while (true) {
      const Chunk cc = chunk_at(S, side, nc);
      if (!__any_sync(0xffffffffu, cc.phase != 0)) break;
      bool worked = false;
      MARK(0, nc); MARK(2, ni);
      // ---- issue (at most two chunks of this group in flight) ----
      if (ni <= nc + 2) {
        const Chunk ci = chunk_at(S, side, ni);
        const int stg = ni % NST;
        if (ci.phase != 0 && mbar_test(BAR(B_EMP + side * NST + stg), ((ni / NST) & 1) ^ 1)) {
          const uint32_t raw = BAR(B_RAW + side * NST + stg);
          unsigned char* slot = s_rows + ((size_t)(side * NST + stg) * KC + f) * rowbytes;
          if (f < ci.len) {
            const int t = ci.t0 + f * ci.dt;
            const char* row = reinterpret_cast<const char*>(xbase + (size_t)t * rstride);
            const char* a0 = reinterpret_cast<const char*>((uintptr_t)row & ~(uintptr_t)15);
            const char* a1 = reinterpret_cast<const char*>(((uintptr_t)row + 4 * C + 15) & ~(uintptr_t)15);
            if (a0 >= p.lo && a1 <= p.hi) {
              const uint32_t bytes = (uint32_t)(a1 - a0);
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              mbar_expect_tx(raw, bytes);
              bulk_g2s(sptr(slot), a0, bytes, raw);
            } else {
              // the superset of this row would leave the tensor: element-wise copy
              float* dst = reinterpret_cast<float*>(slot) + shift;
              const float* src = reinterpret_cast<const float*>(row);
              for (int c = 0; c < C; c++) dst[c] = __ldg(src + c);
              mbar_arrive(raw);
            }
          } else {
            mbar_arrive(raw);
          }
          ni += 2;
          worked = true;
        }
      }
      // ---- convert ----
      bool conv = false;
      const int cstg = nc % NST;
      if (cc.phase != 0 && nc < ni && mbar_test(BAR(B_RAW + side * NST + cstg), (nc / NST) & 1)) {
        conv = true;
        if (f < cc.len) {
          float* slot = reinterpret_cast<float*>(s_rows + ((size_t)(side * NST + cstg) * KC + f) * rowbytes);
          const float xb = slot[shift + blank];
          const float nxb = -xb * 1.4426950408889634f;
          float rs = 0.f;
#pragma unroll 1
          for (int v = 0; v < nv; v++) {
            const float4 x = *reinterpret_cast<const float4*>(slot + 4 * v);
            const int w = 4 * v - shift;   // class of x.x
            float4 r;
            r.x = (w >= 0 && w < C) ? ex2a(fmaf(x.x, 1.4426950408889634f, nxb)) : 0.f;
            r.y = (w + 1 >= 0 && w + 1 < C) ? ex2a(fmaf(x.y, 1.4426950408889634f, nxb)) : 0.f;
            r.z = (w + 2 >= 0 && w + 2 < C) ? ex2a(fmaf(x.z, 1.4426950408889634f, nxb)) : 0.f;
            r.w = (w + 3 >= 0 && w + 3 < C) ? ex2a(fmaf(x.w, 1.4426950408889634f, nxb)) : 0.f;
            rs += (r.x + r.y) + (r.z + r.w);
            *reinterpret_cast<float4*>(slot + 4 * v) = r;
          }
          // a class ratio that overflowed (or a non-finite logit) is the robust kernel's business; ratios that
          // underflow only remove mass and are covered by the certificate
          if (!(rs < 1e37f) || !(rs > 0.f)) alarm |= AL_EMISSION;
          slot[roww - 2] = __fdividef(1.0f, rs);   // y_blank
          slot[roww - 1] = 0.f;                    // "emission" of slots without a label
          if (cc.phase == 1) lsum += __log2f(rs);
        }
        if (cc.phase == 1) s_lsum[lane] = lsum;
        worked = true;
      }
      __syncwarp();
      if (conv) {
        if (f == 0) mbar_arrive(BAR(B_ROW + side * NST + cstg));
        nc += 2;
      }
      if (!__any_sync(0xffffffffu, worked)) __nanosleep(64);
    }
  } else {
    // ================================ gradient warps =================================
    // lane = frame * 4 + group; a lane sums the cells of its group's classes row by row (table rows have the same
    // length in every group, so control flow is uniform), finishes each class in place in the frame's row record
    // (ratio emission -> gradient) and the warp then stores the rows to global memory
    if (want_grad) {
      const int f = lane >> 2, g = lane & 3;
      const int nrows = s_rowtab[0];
      const float PS = 1.8446744e19f;  // 2^PSHIFT
      const float gps = gs * PS;
#pragma unroll 1
      for (int i = S.nc1[d]; i < S.nch; i++) {
        const Chunk ci = chunk_at(S, d, i);
        const int j2 = i - S.nc1[d];
        const int buf = j2 & 1;
        const int stg = i % NST;
        MARK(0, i); MARK(1, 30);
        mbar_wait(BAR(B_GFULL + d * 2 + buf), (j2 >> 1) & 1);
        MARK(1, 31);
        const float* G = s_gbuf + (size_t)((d * 2 + buf) * KC + f) * GROW + g;
        float* rec = reinterpret_cast<float*>(s_rows + ((size_t)(d * NST + stg) * KC + f) * rowbytes) + shift;
        const float yb = rec[roww - 2 - shift];
        // ratio emissions -> grad_loss * softmax, in place (group g takes classes g, g+4, ...)
        const float gy = gs * yb;
        for (int c = g; c < C; c += 4) rec[c] *= gy;
        __syncwarp();
        float tot = 0.f;
#pragma unroll 1
        for (int r = 0; r < nrows; r++) {
          const float* q = G + 16 * r;
          const float acc = (q[0] + q[4]) + (q[8] + q[12]);
          tot += acc;
          const int c = s_clstab[r * 4 + g];
          if (c >= 0) rec[c] = fmaf(-gps, acc, rec[c]);
          __syncwarp();
        }
        tot += __shfl_xor_sync(0xffffffffu, tot, 1);
        tot += __shfl_xor_sync(0xffffffffu, tot, 2);
        tot *= PS;
        if (f < ci.len && !(tot < 1.0f + 1e-4f)) alarm |= AL_OCC;
        if (g == 0) rec[blank] = gs * (yb - (1.0f - tot));
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_GEMP + d * 2 + buf));
        // rows -> global, one row at a time (152 bytes at C = 38: 19 lanes of 8 bytes where the alignment allows)
        const unsigned char* rows = s_rows + (size_t)(d * NST + stg) * KC * rowbytes;
        const bool vec2 = ((shift | C) & 1) == 0 && ((((uintptr_t)gbase) | (rstride * 4)) & 7) == 0;
#pragma unroll 1
        for (int ff = 0; ff < ci.len; ff++) {
          const int t = ci.t0 + ff * ci.dt;
          const float* src = reinterpret_cast<const float*>(rows + (size_t)ff * rowbytes) + shift;
          float* dst = gbase + (size_t)t * rstride;
          if (vec2) {
            for (int c = 2 * lane; c < C; c += 64)
              *reinterpret_cast<float2*>(dst + c) = *reinterpret_cast<const float2*>(src + c);
          } else {
            for (int c = lane; c < C; c += 32) dst[c] = src[c];
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_EMP + d * NST + stg));
      }
    }
  }
  MARK(1, 99);
  if (alarm) atomicOr(&s_scal[0], alarm);
  __syncthreads();
  if (tid == 0) p.retry[b] = s_scal[0];
#undef BAR
#undef MARK
}

}  // namespace narrow

// ---- host side ------------------------------------------------------------------------------------------------------

namespace {

constexpr int kNarrowMaxSmem = 113 * 1024;   // two CTAs per SM

int narrow_nl(int Lmax) {
  if (Lmax <= 4 * 32 - 2) return 4;
  if (Lmax <= 8 * 32 - 2) return 8;
  return 0;
}

// words per row record: a multiple of four with an odd quotient (one row per lane, 16-byte accesses, no bank
// conflicts), holding the 16-byte aligned superset of any row plus y_blank and the zero entry
int narrow_roww(int C) {
  int w = ((C + 3 + 3) & ~3);   // superset of a row that starts up to 3 words into its first 16 bytes
  if (w - (C + 3) < 2) w += 4;
  if (((w / 4) & 1) == 0) w += 4;
  return w;
}

template <int NL>
int launch_narrow(const narrow::Params& p, cudaStream_t stream) {
  const narrow::Smem sl = narrow::smem_layout(NL, p.roww);
  int dev = 0;
  NASR_CUDA(cudaGetDevice(&dev));
  static std::atomic<bool> attr_set[64];
  if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
    NASR_CUDA(cudaFuncSetAttribute(narrow::ctc_narrow_kernel<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kNarrowMaxSmem));
    if (dev >= 0 && dev < 64) attr_set[dev].store(true, std::memory_order_release);
  }
  narrow::ctc_narrow_kernel<NL><<<p.B, narrow::NTHREADS, sl.total, stream>>>(p);
  count_launch();
  NASR_CUDA(cudaGetLastError());
  return NASR_OK;
}

int narrow_maxch(int T) { return (T / 2 + narrow::KC) / narrow::KC + 2; }

}  // namespace

extern int g_debug_split;
extern long long* g_debug_prof;

bool ctc_narrow_supported(int T, int C, int Lmax) {
  if (T < 2 * narrow::KC || C > 64 || C < 2) return false;
  const int NL = narrow_nl(Lmax);
  if (!NL) return false;
  return narrow::smem_layout(NL, narrow_roww(C)).total <= (size_t)kNarrowMaxSmem;
}

size_t ctc_narrow_workspace_bytes(int T, int B, int C, int Lmax) {
  if (!ctc_narrow_supported(T, C, Lmax)) return 0;
  const int NL = narrow_nl(Lmax);
  // under a split override one direction may own every chunk
  const int maxch = (T + narrow::KC - 1) / narrow::KC + 2;
  (void)narrow_maxch;
  return (size_t)B * 2 * maxch * (3 * NL / 4 * 32) * sizeof(float4);
}

int ctc_narrow_launch(const float* logits, int T, int B, int C, long long st_t, long long st_b,
                      const int32_t* label_values, const int32_t* label_offsets, int Lmax, const int32_t* seq_len,
                      int blank, float* loss, float* grad, const float* grad_loss, int32_t* status, int32_t* retry,
                      void* ckpt, cudaStream_t stream) {
  narrow::Params p;
  p.logits = logits; p.T = T; p.B = B; p.C = C; p.st_t = st_t; p.st_b = st_b;
  p.lab_vals = label_values; p.lab_offs = label_offsets; p.seq_len = seq_len;
  p.blank = blank; p.loss = loss; p.grad = grad; p.grad_loss = grad_loss; p.status = status;
  p.retry = retry;
  p.ckpt = static_cast<float4*>(ckpt);
  p.maxch = (T + narrow::KC - 1) / narrow::KC + 2;
  p.roww = narrow_roww(C);
  p.split = g_debug_split;
  p.lo = reinterpret_cast<const char*>(logits);
  p.dbg = reinterpret_cast<int*>(g_debug_prof);
  p.hi = reinterpret_cast<const char*>(logits) +
         4 * ((size_t)(T - 1) * st_t + (size_t)(B - 1) * st_b + (size_t)C);
  switch (narrow_nl(Lmax)) {
    case 4: return launch_narrow<4>(p, stream);
    case 8: return launch_narrow<8>(p, stream);
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
