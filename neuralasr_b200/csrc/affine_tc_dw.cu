// dW = H^T . dL and db = column sums of dL on tcgen05 (the MatMul / BiasAdd gradients of the reference's affine
// projection, networks/bilstm_ctc_net.py:33-45) -- the fast path of nasr_affine_backward_f32's dW / db for K <= 511,
// C <= 40, 16-byte aligned rows of H and contiguous dL; csrc/affine.cu (mma.sync) takes every other shape.
//
//   dW[k, c] = sum_r H[r, k] * dL[r, c]     M = k (four tiles of 128), N = c (48), contraction over the ROWS
//
// The contraction runs over rows, so H -- k contiguous -- is the wrong way round for a K-major operand.  It is turned
// on the way into TMEM: TMA lands [32 rows x 32 k] boxes of H (128-byte swizzle) in shared memory; a converter thread
// owns one k (= one TMEM lane of the A operand) and reads its 32 rows column-wise -- the swizzle spreads a column over
// all banks, so the 32 lanes of a warp read the 32 words of one row: conflict-free -- splits them and writes H^T_hi and
// H^T_lo with two `tcgen05.st.x32`.  dL^T (hi | lo stacked along N, K-major, swizzled by hand) is staged per 32 rows
// by four stager warps.  Per 8 rows and k tile: D[:, 0:96] += H^T_hi . [dL_hi ; dL_lo], D[:, 0:48] += H^T_lo . dL_hi,
// both with the A operand from TMEM.  The four accumulators (4 x 96 TMEM columns) live for the whole kernel; at the end
// each CTA writes its partial sums and `affine_dw_tc_reduce_kernel` adds the CTAs in a fixed order (deterministic).
// db rides along: lane k = K of the last tile feeds ones, so row K of the partial is sum_r dL[r, :].
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "nasr_common.cuh"

namespace nasr {
namespace affine_tc_dw {

constexpr int kThreads = 512;
constexpr int kR = 32;                  // rows per stage (four products of 8 along the contraction)
constexpr int kKT = 4;                  // k tiles of 128
constexpr int kBN = 48;                 // classes, padded
constexpr int kBoxBytes = kR * 128;     // one [32 rows x 32 k] box
constexpr int kBoxes = 16;              // 512 k
constexpr int kHStage = kBoxes * kBoxBytes;     // 65536
constexpr int kHStages = 3;
constexpr int kBTile = kBN * 128;       // 6144: [48 c][32 r] floats, hi; lo follows
constexpr int kBStages = 2;
constexpr int kAccCols = 96;            // per k tile: [. dL_hi | . dL_lo]
constexpr int kACol0 = kKT * kAccCols;  // 384: ring of two A slots (32 hi + 32 lo columns each)
constexpr int kTmemCols = 512;

constexpr int kOffH = 0;
constexpr int kOffB = kOffH + kHStages * kHStage;              // 196608
constexpr int kOffBar = kOffB + kBStages * 2 * kBTile;         // + 24576
constexpr int kNumBars = 2 * kHStages + 4 + 4 + 1;
constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16 + 1024;

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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t tf32_round(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

// K-major, 128-byte swizzle, 8-row groups of 1024 bytes, descriptor version 1
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) { return ((uint64_t)kDescHi << 32) | ((addr & 0x3FFFFu) >> 4); }
constexpr uint32_t idesc(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }
constexpr uint32_t kIdesc2N = idesc(2 * 48), kIdescN = idesc(48);

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc_v,
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

#define NASR_TMEM_ST32(taddr, v)                                                                                       \
  asm volatile(                                                                                                        \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, " \
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),                 \
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),    \
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),      \
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),      \
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])                                                                   \
      : "memory")

struct Params {
  const float* dL;    // [rows, C] contiguous
  long long rows;
  long long per_cta;  // rows per CTA, a multiple of kR
  int K, C;
  float* part;        // [gridDim.x][512][48]
};

__global__ void __launch_bounds__(kThreads, 1) affine_dw_tc_kernel(const __grid_constant__ CUtensorMap tmap_h, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = sbase + kOffBar;
  auto h_full = [&](int s) { return bar0 + 8u * s; };
  auto h_empty = [&](int s) { return bar0 + 8u * (kHStages + s); };
  auto a_full = [&](int s) { return bar0 + 8u * (2 * kHStages + s); };
  auto a_empty = [&](int s) { return bar0 + 8u * (2 * kHStages + 2 + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (2 * kHStages + 4 + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (2 * kHStages + 6 + s); };
  const uint32_t done_bar = bar0 + 8u * (2 * kHStages + 8);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_h)) : "memory");
    for (int s = 0; s < kHStages; s++) {
      mbar_init(h_full(s), 1);
      mbar_init(h_empty(s), 256);
    }
    for (int s = 0; s < 2; s++) {
      mbar_init(a_full(s), 128);
      mbar_init(a_empty(s), 1);
      mbar_init(b_full(s), 128);
      mbar_init(b_empty(s), 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the dL^T tiles: rows c >= C stay zero for the whole kernel
  for (int i = tid; i < kBStages * 2 * kBTile / 4; i += kThreads) reinterpret_cast<uint32_t*>(smem + kOffB)[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;

  const long long r_begin = (long long)blockIdx.x * p.per_cta;
  const long long r_end = min(p.rows, r_begin + p.per_cta);
  const int nst = r_begin < r_end ? (int)((r_end - r_begin + kR - 1) / kR) : 0;

  if (warp == 0) {
    // ===== TMA producer: 16 boxes of [32 rows x 32 k] per stage =====
    for (int st = 0; st < nst; st++) {
      const int hs = st % kHStages;
      mbar_wait(h_empty(hs), ((st / kHStages) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(h_full(hs), kHStage);
        const int c1 = (int)(r_begin + (long long)st * kR);
#pragma unroll 1
        for (int bx = 0; bx < kBoxes; bx++) {
          const uint32_t dst = sbase + kOffH + hs * kHStage + bx * kBoxBytes;
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
              "l"(reinterpret_cast<uint64_t>(&tmap_h)), "r"(h_full(hs)), "r"(bx * 32), "r"(c1)
              : "memory");
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    for (int st = 0; st < nst; st++) {
      const int bs = st & 1;
      mbar_wait(b_full(bs), (st >> 1) & 1);
      const uint32_t bt = sbase + kOffB + bs * 2 * kBTile;
      for (int mt = 0; mt < kKT; mt++) {
        const int mc = st * kKT + mt, as = mc & 1;
        mbar_wait(a_full(as), (mc >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = tmem_base + (uint32_t)(kACol0 + 64 * as), a_lo = a_hi + 32;
        const uint32_t d_tmem = tmem_base + (uint32_t)(mt * kAccCols);
        if (elect_one()) {
#pragma unroll
          for (int k8 = 0; k8 < kR / 8; k8++) {
            const uint64_t bd = make_desc(bt + k8 * 32);
            umma_ts(d_tmem, a_hi + 8 * k8, bd, kIdesc2N, (st | k8) ? 1u : 0u);   // [. dL_hi | . dL_lo]
            umma_ts(d_tmem, a_lo + 8 * k8, bd, kIdescN, 1u);                        // += H_lo . dL_hi
          }
          umma_commit(a_empty(as));
          if (mt == kKT - 1) umma_commit(b_empty(bs));
        }
        __syncwarp();
      }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ===== stagers: dL rows of the stage -> dL^T, hi | lo, K-major swizzled: (c, r) at c*128 + ((r/4 ^ (c&7))*16) + (r%4)*4 =====
    const int t = tid - 128;
    for (int st = 0; st < nst; st++) {
      const int bs = st & 1;
      const long long rs = r_begin + (long long)st * kR;
      float x[10];
#pragma unroll
      for (int u = 0; u < 10; u++) {
        const int idx = t + 128 * u;
        const int r = idx / p.C;
        x[u] = (idx < kR * p.C && rs + r < r_end) ? __ldg(p.dL + rs * p.C + idx) : 0.f;
      }
      mbar_wait(b_empty(bs), ((st >> 1) & 1) ^ 1);
      uint8_t* bt = smem + kOffB + bs * 2 * kBTile;
#pragma unroll
      for (int u = 0; u < 10; u++) {
        const int idx = t + 128 * u;
        if (idx < kR * p.C) {
          const int r = idx / p.C, c = idx - r * p.C;
          const uint32_t hi = tf32_round(x[u]);
          const uint32_t lo = tf32_round(x[u] - __uint_as_float(hi));
          const int off = c * 128 + ((((r >> 2) ^ (c & 7)) & 7) << 4) + (r & 3) * 4;
          *reinterpret_cast<uint32_t*>(bt + off) = hi;
          *reinterpret_cast<uint32_t*>(bt + kBTile + off) = lo;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(b_full(bs));
    }
  } else if (warp >= 8) {
    // ===== converters: thread = one k of the tile = one TMEM lane; set cs takes the k tiles whose count is cs mod 2 =====
    const int cs = (warp - 8) >> 2, q = warp & 3;
    const int kl = 32 * q + lane;                  // k inside the tile (= lane of the A operand)
    for (int st = 0; st < nst; st++) {
      const int hs = st % kHStages;
      mbar_wait(h_full(hs), (st / kHStages) & 1);
      for (int mt = cs; mt < kKT; mt += 2) {       // st*4 + mt has the parity of mt
        const int mc = st * kKT + mt;
        const int k = 128 * mt + kl;
        // column k of the stage: box k/32, word k%32 of every 128-byte row, at chunk ((k%32)/4) ^ (r & 7)
        const uint8_t* box = smem + kOffH + hs * kHStage + (k >> 5) * kBoxBytes;
        const int w4 = (k & 31) >> 2, wi = (k & 3) * 4;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int r = 0; r < kR; r++) {
          float x = *reinterpret_cast<const float*>(box + r * 128 + (((w4 ^ (r & 7)) & 7) << 4) + wi);
          if (k == p.K) x = 1.f;                   // the row of ones: partial row K collects db
          hi[r] = tf32_round(x);
          lo[r] = tf32_round(x - __uint_as_float(hi[r]));
        }
        mbar_wait(a_empty(cs), ((mc >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(kACol0 + 64 * cs);
        NASR_TMEM_ST32(taddr, hi);
        NASR_TMEM_ST32(taddr + 32u, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(a_full(cs));
      }
      mbar_arrive(h_empty(hs));                    // this thread has read its columns of the stage
    }
    // ===== epilogue (set 0): the four accumulators -> this CTA's partial sums =====
    if (cs == 0) {
      mbar_wait(done_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float* slice = p.part + (long long)blockIdx.x * (kKT * 128) * kBN;
      for (int mt = 0; mt < kKT; mt++) {
        uint32_t v[kBN], v2[kBN];
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(mt * kAccCols);
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
        float* row = slice + (long long)(128 * mt + kl) * kBN;
#pragma unroll
        for (int c = 0; c < kBN; c += 4)
          *reinterpret_cast<float4*>(row + c) =
              make_float4(nst ? __uint_as_float(v[c]) + __uint_as_float(v2[c]) : 0.f,
                          nst ? __uint_as_float(v[c + 1]) + __uint_as_float(v2[c + 1]) : 0.f,
                          nst ? __uint_as_float(v[c + 2]) + __uint_as_float(v2[c + 2]) : 0.f,
                          nst ? __uint_as_float(v[c + 3]) + __uint_as_float(v2[c + 3]) : 0.f);
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
  }
}

__global__ void affine_dw_tc_reduce_kernel(const float* __restrict__ part, int slices, int K, int C,
                                           float* __restrict__ dW, float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (K + 1) * C) return;
  const int kk = i / C, c = i - kk * C;
  if (kk < K ? dW == nullptr : db == nullptr) return;
  // eight loads in flight, added in a fixed order: the result does not depend on scheduling
  const float* src = part + (long long)kk * kBN + c;
  const long long step = (long long)(kKT * 128) * kBN;
  float s = 0.f;
  int x = 0;
  for (; x + 8 <= slices; x += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = __ldg(src + (x + u) * step);
#pragma unroll
    for (int u = 0; u < 8; u++) s += v[u];
  }
  for (; x < slices; x++) s += __ldg(src + x * step);
  if (kk < K) dW[(long long)kk * C + c] = s;
  else db[c] = s;
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

static int grid_for(long long rows, int sms) { return (int)std::max<long long>(1, std::min<long long>(sms, (rows + kR - 1) / kR)); }

bool eligible(const float* H, long long rows, int K, long long ldh, long long ldd, int C) {
  return rows >= 1024 && K >= 32 && K <= kKT * 128 - 1 && C <= 40 && ldd == C && (((uintptr_t)H & 15) == 0) &&
         (ldh % 4 == 0) && rows < (1LL << 31) && encode_tiled() != nullptr;
}

size_t workspace_bytes(long long rows, int sms) { return sizeof(float) * (size_t)grid_for(rows, sms) * (kKT * 128) * kBN; }

int dw(const float* H, long long rows, int K, long long ldh, const float* dL, int C, float* dW, float* db, void* ws,
       size_t ws_bytes, cudaStream_t stream) {
  int sms = 0;
  NASR_CUDA(device_sm_count(&sms));
  const int grid = grid_for(rows, sms);
  const size_t need = workspace_bytes(rows, sms);
  if (!ws || ws_bytes < need) {
    set_error("nasr_affine_backward_f32: workspace of %zu bytes, %zu needed (nasr_affine_workspace_bytes)", ws_bytes, need);
    return NASR_ERR_WORKSPACE_TOO_SMALL;
  }
  encode_tiled_fn enc = encode_tiled();
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ldh * sizeof(float)};
  const cuuint32_t box[2] = {32, (cuuint32_t)kR};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(H), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("nasr_affine_backward_f32: cuTensorMapEncodeTiled failed with code %d", (int)r);
    return NASR_ERR_CUDA;
  }
  Params p;
  p.dL = dL;
  p.rows = rows;
  long long per = (rows + grid - 1) / grid;
  p.per_cta = (per + kR - 1) / kR * kR;
  p.K = K;
  p.C = C;
  p.part = static_cast<float*>(ws);
  NASR_CUDA((ensure_max_dynamic_smem<affine_dw_tc_kernel>(kSmemBytes)));
  affine_dw_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tmap, p);
  count_launch();
  NASR_CUDA(cudaGetLastError());
  const int n = (K + 1) * C;
  affine_dw_tc_reduce_kernel<<<(n + 63) / 64, 64, 0, stream>>>(p.part, grid, K, C, dW, db);
  count_launch();
  NASR_CUDA(cudaGetLastError());
  return NASR_OK;
}

}  // namespace affine_tc_dw
}  // namespace nasr
