// dH = dL . W^T of the affine projection's backward on tcgen05 (reference networks/bilstm_ctc_net.py:33-45, the MatMul
// gradient w.r.t. the recurrent outputs) -- the fast path of nasr_affine_backward_f32's dH for C <= 40 and contiguous dL;
// csrc/affine.cu (mma.sync) takes every other shape.
//
//   dH[r, k] = sum_c dL[r, c] * W[k, c]        M = 128 rows per tile, N = 128 columns (k) per product, contraction c <= 40
//
// float32-accurate 3xTF32 as in csrc/affine_tc.cu, but here BOTH parts of the A operand live in TMEM: a converter thread
// owns one row of the tile (= one TMEM lane), reads its 38 floats of dL straight from global memory (dL is 39 MB and
// L2-resident: no TMA, no shared-memory stage), and writes dL_hi and dL_lo with two `tcgen05.st` into 40 + 40 columns.
// W (hi and lo, K-major core-matrix layout without swizzle: rows of 40 floats do not fill a 128-byte swizzle span) is
// resident in shared memory for 256 output columns at a time (passes over k inside one launch; dL is re-read per pass).
// Per 128 x 128 output tile: 5 x 3 `tcgen05.mma` (A from TMEM), then the epilogue -- which is the kernel: 512 MB of dH
// leave through `tcgen05.ld` -> registers -> a padded shared-memory block -> coalesced 16-byte row pieces.  Two
// accumulators in TMEM, one epilogue group of four warps per accumulator.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "nasr_common.cuh"

namespace nasr {
namespace affine_tc_dh {

constexpr int kThreads = 800;            // 8 converter warps (two sets), the issuing warp, 16 epilogue warps
constexpr int kMmaWarp = 8, kEpiWarp0 = 9;
constexpr int kBM = 128;                 // rows per tile
constexpr int kBN = 128;                 // output columns per product
constexpr int kCP = 40;                  // contraction, padded (5 steps of 8)
constexpr int kPassN = 256;              // output columns whose W is resident
constexpr int kACols = 2 * kCP;          // TMEM columns of one A slot: hi | lo
constexpr int kDCol0 = 2 * kACols;       // accumulators behind the two A slots
constexpr int kTmemCols = 512;
constexpr int kSBO = (kCP / 4) * 128;    // bytes between 8-row groups of the W layout
constexpr int kWBytes = (kPassN / 8) * kSBO;          // one part (hi or lo) of a pass: 40960
constexpr int kStageWords = 36;          // padded row of the epilogue's 32-column block: 16-byte pieces, conflict-free
constexpr int kOffWhi = 0;
constexpr int kOffWlo = kOffWhi + kWBytes;
constexpr int kOffStage = kOffWlo + kWBytes;                       // [group][128][33] floats
constexpr int kOffBar = kOffStage + 4 * kBM * kStageWords * 4;
constexpr int kOffTmemPtr = kOffBar + 8 * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16 + 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// K-major without swizzle: core matrices of 8 rows x 16 bytes; LBO = bytes between core matrices along K (128),
// SBO = bytes between 8-row groups; descriptor version 1
constexpr uint32_t kDescHi = (uint32_t)(kSBO >> 4) | (1u << 14);
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  const uint32_t lo = ((addr & 0x3FFFFu) >> 4) | ((128u >> 4) << 16);
  return ((uint64_t)kDescHi << 32) | lo;
}
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(kIdesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t tf32_round(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

#define NASR_TMEM_ST8(taddr, v, o)                                                                                   \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),         \
               "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), \
               "r"(v[o + 7])                                                                                          \
               : "memory")

__device__ long long g_prof_dh[8];   // tuning: CTA 0 cycles: converter load / wait, mma wait a / wait d, epilogue wait / work

struct Params {
  const float* dL;   // [rows, C] contiguous
  const float* W;    // [K, C]
  float* dH;         // [rows, K] at stride ldh
  long long rows, ldh;
  int K, C;
  int vec16;         // rows of dH 16-byte aligned
  int debug;         // 8: role cycle counters of CTA 0
};

__global__ void __launch_bounds__(kThreads, 1) affine_dh_tc_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = sbase + kOffBar;
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (2 + s); };
  auto d_full = [&](int s) { return bar0 + 8u * (4 + s); };
  auto d_empty = [&](int s) { return bar0 + 8u * (6 + s); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);

  if (tid == 0) {
    for (int s = 0; s < 2; s++) {
      mbar_init(a_full(s), 128);
      mbar_init(a_empty(s), 1);
      mbar_init(d_full(s), 1);
      mbar_init(d_empty(s), 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }

  const long long ntiles = (p.rows + kBM - 1) / kBM;
  int acnt = 0, dcnt = 0;          // A slots / accumulators used so far (every role counts the same sequence)
  long long c0 = 0, c1 = 0;
  const long long kstart = clock64();
  uint32_t tmem_base = 0;

  for (int k0 = 0; k0 < p.K; k0 += kPassN) {
    const int ncols = min(kPassN, p.K - k0);
    const int nnt = (ncols + kBN - 1) / kBN;          // 128-column products per row tile in this pass
    if (k0 > 0) {
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();                                 // every product on the previous pass's W has completed
    }
    // W rows [k0, k0 + 256): element (n, c) at (n/8)*SBO + (c/4)*128 + (n%8)*16 + (c%4)*4, hi and lo parts
    for (int base = tid; base < kPassN * kCP; base += 8 * kThreads) {
      float w[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int idx = base + u * kThreads;
        const int c = idx % kCP, n = idx / kCP;
        w[u] = (idx < kPassN * kCP && c < p.C && k0 + n < p.K) ? __ldg(p.W + (long long)(k0 + n) * p.C + c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int idx = base + u * kThreads;
        if (idx >= kPassN * kCP) break;
        const int c = idx % kCP, n = idx / kCP;
        const uint32_t hi = tf32_round(w[u]);
        const uint32_t lo = tf32_round(w[u] - __uint_as_float(hi));
        const int off = (n >> 3) * kSBO + (c >> 2) * 128 + (n & 7) * 16 + (c & 3) * 4;
        *reinterpret_cast<uint32_t*>(smem + kOffWhi + off) = hi;
        *reinterpret_cast<uint32_t*>(smem + kOffWlo + off) = lo;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem_base = *tmem_ptr_smem;

    if (warp < 8) {
      // ===== converters: one row of dL per thread -> hi | lo in TMEM.  Two sets of four warps take alternate tiles
      //       (set = A slot): under the epilogue's store traffic a tile's loads take longer than a tile lasts =====
      const int cs = warp >> 2, sub = warp & 3;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, acnt++) {
        if ((acnt & 1) != cs) continue;
        const int as = cs;
        const long long r = tile * kBM + 32 * sub + lane;
        const long long t0 = clock64();
        float x[kCP];
#pragma unroll
        for (int c = 0; c < kCP; c++) x[c] = 0.f;
        if (r < p.rows) {
          const float* row = p.dL + r * p.C;
          if ((p.C & 1) == 0 && (((uintptr_t)p.dL & 7) == 0)) {
#pragma unroll
            for (int c = 0; c < kCP; c += 2) {
              if (c + 1 < p.C) {
                const float2 v = __ldg(reinterpret_cast<const float2*>(row + c));
                x[c] = v.x;
                x[c + 1] = v.y;
              }
            }
          } else {
#pragma unroll
            for (int c = 0; c < kCP; c++)
              if (c < p.C) x[c] = __ldg(row + c);
          }
        }
        const long long t1 = clock64();
        mbar_wait(a_empty(as), ((acnt >> 1) & 1) ^ 1);
        c0 += t1 - t0;
        c1 += clock64() - t1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * sub) << 16) + (uint32_t)(as * kACols);
#pragma unroll
        for (int o = 0; o < kCP; o += 8) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int c = 0; c < 8; c++) {
            hi[c] = tf32_round(x[o + c]);
            lo[c] = tf32_round(x[o + c] - __uint_as_float(hi[c]));
          }
          NASR_TMEM_ST8(taddr + (uint32_t)o, hi, 0);
          NASR_TMEM_ST8(taddr + (uint32_t)(kCP + o), lo, 0);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(a_full(as));
      }
    } else if (warp == kMmaWarp) {
      // ===== MMA issuer =====
      const uint32_t whi = sbase + kOffWhi, wlo = sbase + kOffWlo;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int as = acnt & 1;
        { const long long t0 = clock64(); mbar_wait(a_full(as), (acnt >> 1) & 1); c0 += clock64() - t0; }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = tmem_base + (uint32_t)(as * kACols), a_lo = a_hi + kCP;
        for (int j = 0; j < nnt; j++) {
          const int ds = dcnt & 1;
          { const long long t0 = clock64(); mbar_wait(d_empty(ds), ((dcnt >> 1) & 1) ^ 1); c1 += clock64() - t0; }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d_tmem = tmem_base + (uint32_t)(kDCol0 + ds * kBN);
          const uint32_t boff = (uint32_t)(j * (kBN / 8) * kSBO);
          if (elect_one()) {
#pragma unroll
            for (int s8 = 0; s8 < kCP / 8; s8++) {
              const uint32_t ko = boff + s8 * 256;   // two core matrices along the contraction
              umma_ts(d_tmem, a_lo + 8 * s8, make_desc(whi + ko), s8 ? 1u : 0u);
              umma_ts(d_tmem, a_hi + 8 * s8, make_desc(wlo + ko), 1u);
              umma_ts(d_tmem, a_hi + 8 * s8, make_desc(whi + ko), 1u);
            }
            umma_commit(d_full(ds));
            if (j == nnt - 1) umma_commit(a_empty(as));
          }
          __syncwarp();
          dcnt++;
        }
        acnt++;
      }
    } else {
      // ===== epilogue: four groups of four warps; accumulator g is read by the groups (g, half 0) and (g, half 1),
      //       64 columns each, so that one group's loads and staging run under another's stores =====
      const int gi = (warp - kEpiWarp0) >> 2, g = gi >> 1, hf = gi & 1, q = warp & 3;
      const int et = tid - 32 * kEpiWarp0 - 128 * gi;   // 0..127 inside the group
      float* stage = reinterpret_cast<float*>(smem + kOffStage) + gi * kBM * kStageWords;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int j = 0; j < nnt; j++, dcnt++) {
          if ((dcnt & 1) != g) continue;
          const long long t0 = clock64();
          mbar_wait(d_full(g), (dcnt >> 1) & 1);
          const long long t1 = clock64();
          c0 += t1 - t0;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(kDCol0 + g * kBN + 64 * hf);
          const long long r0 = tile * kBM;
          const int kc0 = k0 + j * kBN + 64 * hf;         // first output column of this group's half
#pragma unroll 1
          for (int ch = 0; ch < 2; ch++) {
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr + (uint32_t)(32 * ch)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (ch == 1) {
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              mbar_arrive(d_empty(g));       // this half of the accumulator is read
            }
            // rows of 32 columns into the padded block (row stride 36 words: 16-byte pieces of eight rows hit every
            // bank once), then 128-byte row pieces out
            asm volatile("bar.sync %0, 128;" ::"r"(1 + gi) : "memory");
            float4* srow = reinterpret_cast<float4*>(stage + (32 * q + lane) * kStageWords);
#pragma unroll
            for (int c = 0; c < 8; c++)
              srow[c] = make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                                    __uint_as_float(v[4 * c + 3]));
            asm volatile("bar.sync %0, 128;" ::"r"(1 + gi) : "memory");
            const int col = (et & 7) * 4;
            const int kc = kc0 + 32 * ch + col;
#pragma unroll
            for (int it = 0; it < kBM / 16; it++) {
              const int rr = (et >> 3) + 16 * it;
              const long long r = r0 + rr;
              if (r < p.rows && kc < p.K && !(p.debug & 4)) {
                const float4 o = *reinterpret_cast<const float4*>(stage + rr * kStageWords + col);
                float* dst = p.dH + r * p.ldh + kc;
                if (p.vec16 && kc + 3 < p.K) {
                  *reinterpret_cast<float4*>(dst) = o;
                } else {
                  dst[0] = o.x;
                  if (kc + 1 < p.K) dst[1] = o.y;
                  if (kc + 2 < p.K) dst[2] = o.z;
                  if (kc + 3 < p.K) dst[3] = o.w;
                }
              }
            }
          }
          c1 += clock64() - t1;
        }
      }
    }
  }

  if ((p.debug & 8) && blockIdx.x == 0) {
    if (tid == 0) { g_prof_dh[0] = c0; g_prof_dh[1] = clock64() - kstart; }
    if (tid == 32 * kMmaWarp) { g_prof_dh[2] = c0; g_prof_dh[3] = c1; }
    if (tid == 32 * kEpiWarp0) { g_prof_dh[4] = c0; g_prof_dh[5] = c1; }
    if (tid == 32 * kEpiWarp0 + 256) { g_prof_dh[6] = c0; g_prof_dh[7] = c1; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
  }
}

bool eligible(const float* dL, long long rows, int K, long long ldd, int C) {
  return rows >= kBM && C <= kCP && ldd == C && K >= 8;
}

int dh(const float* dL, long long rows, int K, const float* W, int C, float* dH, long long ldh, cudaStream_t stream) {
  int sms = 0;
  NASR_CUDA(device_sm_count(&sms));
  const long long ntiles = (rows + kBM - 1) / kBM;
  const int grid = (int)std::min<long long>(sms, ntiles);
  Params p;
  p.dL = dL;
  p.W = W;
  p.dH = dH;
  p.rows = rows;
  p.ldh = ldh;
  p.K = K;
  p.C = C;
  p.vec16 = (((uintptr_t)dH & 15) == 0) && (ldh % 4 == 0);
  {
    const char* e = getenv("NASR_AFFINE_TC_DEBUG");
    p.debug = e ? atoi(e) : 0;
  }
  NASR_CUDA((ensure_max_dynamic_smem<affine_dh_tc_kernel>(kSmemBytes)));
  affine_dh_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(p);
  count_launch();
  NASR_CUDA(cudaGetLastError());
  if (p.debug & 8) {
    long long h[8];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(h, g_prof_dh, sizeof(h));
    fprintf(stderr, "affine_dh_tc (CTA 0, cycles): converter load %lld, CTA lifetime so far %lld | mma wait A %lld wait D %lld | epilogue (0,0) wait %lld work %lld | "
            "epilogue (1,0) wait %lld work %lld\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
  }
  return NASR_OK;
}

}  // namespace affine_tc_dh
}  // namespace nasr
