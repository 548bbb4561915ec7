// Forward affine projection on Blackwell's tcgen05 tensor cores (logits = H.W + b, reference
// networks/bilstm_ctc_net.py:33-45) -- the fast path of nasr_affine_logits_f32 for C <= 40 and 16-byte aligned rows of H;
// csrc/affine.cu (mma.sync) takes every other shape and the backward.
//
// float32-accurate 3xTF32 with operands in shared memory:
//   * `tcgen05.mma.kind::tf32` reads float32 words and ignores their low 13 bits, so the RAW tile of H that TMA lands
//     is its own high part (H_hi = trunc(H)); four converter warps write H_lo = H - trunc(H) (rounded to TF32) into a
//     second tile of the stage -- an elementwise pass, so the 128-byte swizzle TMA applied needs no decoding;
//   * W is staged once per CTA as W_hi (rounded to TF32) and W_lo = W - W_hi, K-major, 128-byte swizzled by hand;
//   * per 8 k:  D += H_lo.W_hi;  D += H_raw.W_lo;  D += H_raw.W_hi   (M = 64 rows, N = 40 classes, accumulator in TMEM).
//     Dropped: H_lo.W_lo and the roundings of the low parts, ~2^-21 |h||w| per product.
// Shared memory decides the shape of the kernel: W_hi + W_lo for all K = 500 would take 160 KB and leave four 16 KB
// stages (24 KB of H in flight per SM: not enough to keep HBM busy), so the contraction runs in PASSES of 256 k -- one
// launch per pass, 80 KB of W resident, eight stages (56 KB in flight); the second pass adds to the logits the first
// wrote (39 MB of extra read + write on 551 MB).  H is still read once.
//
// Roles (320 threads, one CTA per SM, persistent over 64-row tiles):
//   warp 0      TMA producer: one lane, `cp.async.bulk.tensor.2d` of the [64 rows x 32 k] box per stage, mbarrier tx
//   warp 1      TMEM allocation; one lane issues the tcgen05.mma's, `tcgen05.commit` frees the stage / publishes the tile
//   warps 2-5   converters: raw tile -> low tile, `fence.proxy.async`, arrive
//   warps 6-9   epilogue: `tcgen05.ld` (M = 64 puts row r in TMEM lane (r % 16) + 32 (r / 16)), bias, (+ old logits), store
// Two accumulators in TMEM (2 x 64 columns) let the epilogue of a tile run under the products of the next.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "nasr_common.cuh"

namespace nasr {
namespace affine_tc {

constexpr int kThreads = 320;
#ifndef TC_BM
#define TC_BM 128   // 128: the shipped kernel (low tile of H in TMEM); 64: both tiles in shared memory (measured: 0.166 ms)
#endif
constexpr int kBM = TC_BM;              // rows per tile (UMMA M)
constexpr int kBN = TC_BM == 128 ? 48 : 40;   // classes (UMMA N: a multiple of 8 at M = 64, of 16 at M = 128)
constexpr int kBK = 32;                 // k per stage: 128 bytes, one swizzle span
constexpr int kPassBlocks = 8;          // k-blocks per pass (256 k)
constexpr bool kLoInTmem = TC_BM == 128;  // the low tile of H in TMEM (A operand from TMEM) instead of shared memory
constexpr int kStages = kLoInTmem ? 6 : 8;
constexpr int kLoSlots = 4;             // TMEM ring of low tiles (32 columns each)
constexpr int kATile = kBM * kBK * 4;   // 8192
constexpr int kBTile = kBN * kBK * 4;   // 5120 (five 8-row groups of 1024 bytes)
constexpr int kAccCols = 128;           // columns of one accumulator slot: [H.W_hi | H_raw.W_lo], 2 * kBN <= 128
constexpr int kTmemCols = kLoInTmem ? 512 : 2 * kAccCols;   // two accumulators (+ the low tiles at column 256)
constexpr int kLoCol0 = 2 * kAccCols;

// shared memory map (offsets from a 1024-byte aligned base)
constexpr int kOffBhi = 0;
constexpr int kOffA = kOffBhi + kPassBlocks * 2 * kBTile;          // W: [k-block][hi tile | lo tile]; then H: [stage][raw | lo]
constexpr int kStageBytes = kLoInTmem ? kATile : 2 * kATile;       // raw tile (+ low tile)
constexpr int kOffBar = kOffA + kStages * kStageBytes;
constexpr int kNumBars = 3 * kStages + 4 + kLoSlots;                // full_raw, full_lo, empty per stage; tmem full/empty x2; lo_empty
constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr int kOffBias = kOffTmemPtr + 16;                          // [kBN] floats (zeros without a bias)
constexpr int kOffOut = (kOffBias + 64 * 4 + 15) & ~15;              // [64][C] floats: the tile's logits before they leave
constexpr int kSmemBytes = kOffOut + kBM * kBN * 4 + 1024;          // + slack for the manual alignment

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// K-major, 128-byte swizzle: 8-row groups of 1024 bytes (SBO), LBO unused, descriptor version 1 (sm_100).  The high
// word is a constant; the low word is the start address >> 4 (advancing 8 floats along K adds 2).
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return (addr & 0x3FFFFu) >> 4; }
__device__ __forceinline__ uint64_t make_desc(uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; }

// kind::tf32, D = F32, A and B = TF32, both K-major, N = 40, M = 64
constexpr uint32_t idesc(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24); }
constexpr uint32_t kIdescN = idesc(kBN), kIdesc2N = idesc(2 * kBN);

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc_v,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc_v), "r"(accumulate), "r"(0u)
      : "memory");
}
// A operand from TMEM
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc_v,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc_v), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ long long g_prof[16];   // tuning (debug bit 8): cycles of CTA 0 per role: waiting / working

// one lane of a converged warp (the compiler keeps the operands of what it guards in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

struct Params {
  long long rows;
  int K, C;
  const float* W;         // [K, C]
  const float* bias;      // [C] or NULL
  float* out;
  long long ldo;
  int vec16;              // out 16-byte aligned and 64 rows of C floats a multiple of 16 bytes
  int debug;              // tuning: 4 no stores (results wrong), 8 role cycle counters of CTA 0
};

__global__ void __launch_bounds__(kThreads, 1) affine_tc_kernel(const __grid_constant__ CUtensorMap tmap_h, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = sbase + kOffBar;
  auto full_raw = [&](int s) { return bar0 + 8u * s; };
  auto full_lo = [&](int s) { return bar0 + 8u * (kStages + s); };
  auto empty = [&](int s) { return bar0 + 8u * (2 * kStages + s); };
  auto tmem_full = [&](int a) { return bar0 + 8u * (3 * kStages + a); };
  auto tmem_empty = [&](int a) { return bar0 + 8u * (3 * kStages + 2 + a); };
  auto lo_empty = [&](int t) { return bar0 + 8u * (3 * kStages + 4 + t); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);

  // ---- one-time setup ------------------------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_h)) : "memory");
    for (int s = 0; s < kStages; s++) {
      mbar_init(full_raw(s), 1);
      mbar_init(full_lo(s), 128);
      mbar_init(empty(s), 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tmem_full(a), 1);
      mbar_init(tmem_empty(a), 128);
    }
    for (int t = 0; t < kLoSlots; t++) mbar_init(lo_empty(t), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  const long long ntiles = (p.rows + kBM - 1) / kBM;
  // ring / accumulator state of this thread's role: it runs on across the passes
  int s = 0, acc = 0, lt = 0;
  uint32_t ph = 0, acc_ph[2] = {0, 0}, lph = 0;
  long long pw = 0, mw = 0, mi = 0, me = 0, cw = 0, cc = 0, ew = 0, ec = 0;
  const long long pstart = clock64();
  uint32_t tmem_base = 0;

  for (int k0 = 0; k0 < p.K; k0 += kPassBlocks * kBK) {
  const int nkb = min(kPassBlocks, (p.K - k0 + kBK - 1) / kBK);
  const bool accumulate = k0 > 0;          // later passes add to the logits the first one wrote
  if (k0 > 0) {
    // every product that reads the previous pass's W has completed: the epilogue warps get here only after the last
    // accumulator, committed behind all of them, has been read
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  // W of this pass: element (n, k) of k-block kb at kb*2*5120 + n*128 + ((k/4 ^ (n & 7)) * 16) + (k % 4) * 4
  {
    const int total = nkb * kBN * kBK;
    for (int base = tid; base < total; base += 8 * kThreads) {
      float w[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {           // eight loads in flight per thread
        const int idx = base + u * kThreads;
        const int n = idx % kBN;               // n fastest: consecutive threads read consecutive classes of one k
        const int kg = k0 + idx / kBN;
        w[u] = (idx < total && n < p.C && kg < p.K) ? __ldg(p.W + (long long)kg * p.C + n) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int idx = base + u * kThreads;
        if (idx >= total) break;
        const int n = idx % kBN, kk = idx / kBN;
        const int kb = kk >> 5, k = kk & 31;
        const uint32_t hi = (__float_as_uint(w[u]) + 0x1000u) & 0xffffe000u;
        const float lo_f = w[u] - __uint_as_float(hi);
        const uint32_t lo = (__float_as_uint(lo_f) + 0x1000u) & 0xffffe000u;
        const int off = kb * 2 * kBTile + n * 128 + ((((k >> 2) ^ (n & 7)) & 7) << 4) + (k & 3) * 4;
        *reinterpret_cast<uint32_t*>(smem + kOffBhi + off) = hi;
        *reinterpret_cast<uint32_t*>(smem + kOffBhi + kBTile + off) = lo;
      }
    }
  }
  if (tid < kBN) reinterpret_cast<float*>(smem + kOffBias)[tid] = (p.bias && !accumulate && tid < p.C) ? __ldg(p.bias + tid) : 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tensor core reads W through the async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer (the whole warp walks the ring; one elected lane issues) =====
    {
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int kb = 0; kb < nkb; kb++) {
          const long long t0 = clock64();
          mbar_wait(empty(s), ph ^ 1);
          pw += clock64() - t0;
          if (elect_one()) {
            mbar_expect_tx(full_raw(s), kATile);
            const uint32_t dst = sbase + kOffA + s * kStageBytes;
            const int c0 = k0 + kb * kBK, c1 = (int)(tile * kBM);
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                "l"(reinterpret_cast<uint64_t>(&tmap_h)), "r"(full_raw(s)), "r"(c0), "r"(c1)
                : "memory");
          }
          __syncwarp();
          if (++s == kStages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
      if ((p.debug & 8) && blockIdx.x == 0 && lane == 0) {
        g_prof[0] = pw;
        g_prof[1] = clock64() - pstart;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (the whole warp walks the ring; one elected lane issues) =====
    {
      const uint32_t a_desc0 = desc_lo(sbase + kOffA), b_desc0 = desc_lo(sbase + kOffBhi);
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long te = clock64();
        mbar_wait(tmem_empty(acc), acc_ph[acc] ^ 1);
        me += clock64() - te;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols);
        for (int kb = 0; kb < nkb; kb++) {
          const long long t0 = clock64();
          mbar_wait(full_lo(kLoInTmem ? lt : s), kLoInTmem ? lph : ph);
          const long long t1 = clock64();
          mw += t1 - t0;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_raw = a_desc0 + (uint32_t)s * (kStageBytes >> 4), a_lo = a_raw + (kATile >> 4);
          const uint32_t a_lo_tmem = tmem_base + (uint32_t)(kLoCol0 + 32 * lt);
          const uint32_t b_hi = b_desc0 + (uint32_t)kb * (2 * kBTile >> 4);   // the lo tile follows the hi tile
          if (elect_one()) {
#pragma unroll
            for (int k8 = 0; k8 < kBK / 8; k8++) {
              const uint32_t ko = k8 * 2;   // 8 floats along K inside the swizzle span, in 16-byte units
              // columns [0, N): H_raw.W_hi, columns [N, 2N): H_raw.W_lo -- ONE product over the stacked W tile
              umma_tf32(d_tmem, make_desc(a_raw + ko), make_desc(b_hi + ko), kIdesc2N, (kb | k8) ? 1u : 0u);
              // columns [0, N) += H_lo.W_hi
              if (kLoInTmem) umma_tf32_ts(d_tmem, a_lo_tmem + 8 * k8, make_desc(b_hi + ko), kIdescN, 1u);
              else umma_tf32(d_tmem, make_desc(a_lo + ko), make_desc(b_hi + ko), kIdescN, 1u);
            }
            umma_commit(empty(s));            // the stage is free once these products have read it
            if (kLoInTmem) umma_commit(lo_empty(lt));
            if (kb == nkb - 1) umma_commit(tmem_full(acc));
          }
          if (++lt == kLoSlots) {
            lt = 0;
            lph ^= 1;
          }
          __syncwarp();
          mi += clock64() - t1;
          if (++s == kStages) {
            s = 0;
            ph ^= 1;
          }
        }
        acc_ph[acc] ^= 1;
        acc ^= 1;
      }
      if ((p.debug & 8) && blockIdx.x == 0 && lane == 0) {
        g_prof[2] = mw;
        g_prof[3] = mi;
        g_prof[4] = me;
      }
    }
  } else if (warp < 6) {
    // ===== converters: low part of the raw tile =====
    const int ct = tid - 64;               // 0..127
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int kb = 0; kb < nkb; kb++) {
        const long long t0 = clock64();
        mbar_wait(full_raw(s), ph);
        const long long t1 = clock64();
        cw += t1 - t0;
        if (kLoInTmem) {
          // thread = row of the tile = TMEM lane (a warp reaches the 32 lanes of its own sub-partition): the row's 32 k
          // from the swizzled raw tile, their low parts into 32 TMEM columns of the ring slot
          mbar_wait(lo_empty(lt), lph ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const int r = 32 * (warp & 3) + lane;
          const uint8_t* rowp = smem + kOffA + s * kStageBytes + r * 128;
          uint32_t lo[32];
#pragma unroll
          for (int c = 0; c < 8; c++) {
            const float4 v = *reinterpret_cast<const float4*>(rowp + (((c ^ (r & 7)) & 7) << 4));
            lo[4 * c + 0] = (__float_as_uint(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u)) + 0x1000u) & 0xffffe000u;
            lo[4 * c + 1] = (__float_as_uint(v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u)) + 0x1000u) & 0xffffe000u;
            lo[4 * c + 2] = (__float_as_uint(v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u)) + 0x1000u) & 0xffffe000u;
            lo[4 * c + 3] = (__float_as_uint(v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u)) + 0x1000u) & 0xffffe000u;
          }
          const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(kLoCol0 + 32 * lt);
          asm volatile(
              "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
              "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
              "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]), "r"(lo[8]),
              "r"(lo[9]), "r"(lo[10]), "r"(lo[11]), "r"(lo[12]), "r"(lo[13]), "r"(lo[14]), "r"(lo[15]), "r"(lo[16]),
              "r"(lo[17]), "r"(lo[18]), "r"(lo[19]), "r"(lo[20]), "r"(lo[21]), "r"(lo[22]), "r"(lo[23]), "r"(lo[24]),
              "r"(lo[25]), "r"(lo[26]), "r"(lo[27]), "r"(lo[28]), "r"(lo[29]), "r"(lo[30]), "r"(lo[31])
              : "memory");
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(full_lo(lt));
          if (++lt == kLoSlots) {
            lt = 0;
            lph ^= 1;
          }
        } else {
        const float4* src = reinterpret_cast<const float4*>(smem + kOffA + s * 2 * kATile);
        float4* dst = reinterpret_cast<float4*>(smem + kOffA + s * 2 * kATile + kATile);
#pragma unroll
        for (int i = 0; i < kATile / 16 / 128; i++) {
          const float4 v = src[ct + 128 * i];
          float4 o;
          // lo = x - trunc_tf32(x), then rounded to TF32 (what the tensor core would otherwise truncate)
          o.x = __uint_as_float((__float_as_uint(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u)) + 0x1000u) & 0xffffe000u);
          o.y = __uint_as_float((__float_as_uint(v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u)) + 0x1000u) & 0xffffe000u);
          o.z = __uint_as_float((__float_as_uint(v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u)) + 0x1000u) & 0xffffe000u);
          o.w = __uint_as_float((__float_as_uint(v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u)) + 0x1000u) & 0xffffe000u);
          dst[ct + 128 * i] = o;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(full_lo(s));
        }
        cc += clock64() - t1;
        if (++s == kStages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
    if ((p.debug & 8) && blockIdx.x == 0 && ct == 0) {
      g_prof[5] = cw;
      g_prof[6] = cc;
    }
  } else {
    // ===== epilogue =====
    const int q = warp & 3;                 // TMEM sub-partition of this warp: lanes 32q .. 32q+31
    const int et = tid - 192;               // 0..127
    const float* bias_s = reinterpret_cast<const float*>(smem + kOffBias);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      // the tile's block of the logits: [nrows][C] floats, contiguous when ldo == C
      const long long r0 = tile * kBM;
      const int count = (int)min((long long)kBM, p.rows - r0) * p.C;
      const bool fastpath = p.ldo == p.C && p.vec16;
      float* obase = p.out + r0 * p.ldo;
      const int n4 = count >> 2;
      constexpr int kOld = (kBM * kBN / 4 + 127) / 128;   // 16-byte pieces per thread
      float4 old[kOld];
      if (accumulate && fastpath) {
        // later passes add to what the first wrote: those loads are issued BEFORE the wait for this tile's products
#pragma unroll
        for (int u = 0; u < kOld; u++) {
          const int i = et + 128 * u;
          old[u] = i < n4 ? __ldcs(reinterpret_cast<const float4*>(obase) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      const long long t0 = clock64();
      mbar_wait(tmem_full(acc), acc_ph[acc]);
      const long long t1 = clock64();
      ew += t1 - t0;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v[kBN], v2[kBN];
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * kAccCols);
#pragma unroll
      for (int c8 = 0; c8 < kBN / 8; c8++) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[c8 * 8 + 0]), "=r"(v[c8 * 8 + 1]), "=r"(v[c8 * 8 + 2]), "=r"(v[c8 * 8 + 3]),
                       "=r"(v[c8 * 8 + 4]), "=r"(v[c8 * 8 + 5]), "=r"(v[c8 * 8 + 6]), "=r"(v[c8 * 8 + 7])
                     : "r"(taddr + (uint32_t)(c8 * 8)));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v2[c8 * 8 + 0]), "=r"(v2[c8 * 8 + 1]), "=r"(v2[c8 * 8 + 2]), "=r"(v2[c8 * 8 + 3]),
                       "=r"(v2[c8 * 8 + 4]), "=r"(v2[c8 * 8 + 5]), "=r"(v2[c8 * 8 + 6]), "=r"(v2[c8 * 8 + 7])
                     : "r"(taddr + (uint32_t)(kBN + c8 * 8)));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(tmem_empty(acc));         // the accumulator is in registers: the next tile may overwrite it
      // M = 64: row r of the tile sits in TMEM lane (r % 16) + 32 (r / 16), i.e. lanes 0-15 of every sub-partition.
      // The rows go to shared memory as the [64][C] block they are in global memory (rows of the logits are
      // contiguous when ldo == C), then the four warps copy the block out with coalesced accesses.
      float* stage = reinterpret_cast<float*>(smem + kOffOut);
      asm volatile("bar.sync 1, 128;" ::: "memory");      // the previous tile's copy-out has read the staging block
      if (kBM == 128 || lane < 16) {
        float* srow = stage + (kBM == 128 ? 32 * q + lane : 16 * q + lane) * p.C;
#pragma unroll
        for (int c = 0; c < kBN; c++)
          if (c < p.C) srow[c] = (__uint_as_float(v[c]) + __uint_as_float(v2[c])) + bias_s[c];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (!(p.debug & 4)) {
        if (fastpath) {
#pragma unroll
          for (int u = 0; u < kOld; u++) {
            const int i = et + 128 * u;
            if (i < n4) {
              float4 o = reinterpret_cast<const float4*>(stage)[i];
              if (accumulate) {
                o.x += old[u].x; o.y += old[u].y; o.z += old[u].z; o.w += old[u].w;
              }
              reinterpret_cast<float4*>(obase)[i] = o;
            }
          }
          for (int i = 4 * n4 + et; i < count; i += 128) obase[i] = stage[i] + (accumulate ? obase[i] : 0.f);
        } else {
          for (int i = et; i < count; i += 128) {
            const int rr = i / p.C, c = i - rr * p.C;
            float* dst = p.out + (r0 + rr) * p.ldo + c;
            *dst = stage[i] + (accumulate ? *dst : 0.f);
          }
        }
      }
      ec += clock64() - t1;
      acc_ph[acc] ^= 1;
      acc ^= 1;
    }
    if ((p.debug & 8) && blockIdx.x == 0 && tid == 192) {
      g_prof[7] = ew;
      g_prof[8] = ec;
    }
  }

  }   // passes

  // ---- teardown ------------------------------------------------------------------------------------------------
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
  }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn encode_tiled() {
  static std::atomic<encode_tiled_fn> cached{nullptr};
  encode_tiled_fn f = cached.load(std::memory_order_acquire);
  if (!f) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      f = reinterpret_cast<encode_tiled_fn>(sym);
    cached.store(f, std::memory_order_release);
  }
  return f;
}

// Can this call take the tcgen05 path?
bool eligible(const float* H, long long rows, int K, long long ldh, int C) {
  return rows >= kBM && C <= 40 && K >= kBK && (((uintptr_t)H & 15) == 0) && (ldh % 4 == 0) &&
         rows < (1LL << 31) && encode_tiled() != nullptr;
}

int forward(const float* H, long long rows, int K, long long ldh, const float* W, const float* bias, int C,
            float* logits, long long ldl, cudaStream_t stream) {
  encode_tiled_fn enc = encode_tiled();
  if (!enc) {
    set_error("nasr_affine_logits_f32: cuTensorMapEncodeTiled is not available in this driver");
    return NASR_ERR_UNSUPPORTED;
  }
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ldh * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)kBM};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(H), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("nasr_affine_logits_f32: cuTensorMapEncodeTiled failed with code %d", (int)r);
    return NASR_ERR_CUDA;
  }
  int sms = 0;
  NASR_CUDA(device_sm_count(&sms));
  const long long ntiles = (rows + kBM - 1) / kBM;
  const int grid = (int)std::min<long long>(sms, ntiles);
  NASR_CUDA((ensure_max_dynamic_smem<affine_tc_kernel>(kSmemBytes)));
  Params p;
  p.rows = rows;
  p.K = K;
  p.C = C;
  p.W = W;
  p.bias = bias;
  p.out = logits;
  p.ldo = ldl;
  p.vec16 = (((uintptr_t)logits & 15) == 0) && ((kBM * C) % 4 == 0);
  {
    const char* e = getenv("NASR_AFFINE_TC_DEBUG");
    p.debug = e ? atoi(e) : 0;
  }
  affine_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tmap, p);
  count_launch();
  NASR_CUDA(cudaGetLastError());
  if (p.debug & 8) {
    long long h[16];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(h, g_prof, sizeof(h));
    fprintf(stderr, "affine_tc (CTA 0, cycles, all passes): producer wait %lld of %lld | mma wait full %lld issue %lld wait tmem %lld | "
            "converter wait %lld work %lld | epilogue wait %lld work %lld\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8]);
  }
  return NASR_OK;
}

}  // namespace affine_tc
}  // namespace nasr
