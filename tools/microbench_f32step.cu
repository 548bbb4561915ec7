// Microbenchmark of the fp32 CTC recursion step (csrc/ctc_narrow.cu) on B200: cycles per frame of one warp and of
// 1..4 warps per scheduler, for the three modes a frame is visited in (phase 1, recompute with stores, phase 2 with
// the other direction's values and posterior stores), packed f32x2 against scalar arithmetic; plus the issue rates
// the design leans on (LDS.32 gathers, SHFL and LDS on one SM, packed add/mul, mbarrier round trips).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --cudart=shared -o tools/microbench_f32step tools/microbench_f32step.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

constexpr int ROWW = 44;  // words per emission row
constexpr int KC = 8;

// MODE 0 phase 1, 1 recompute (stores), 2 phase 2 (loads of the other direction, posterior stores)
template <int MODE, bool PACKED>
__global__ void __launch_bounds__(512) step_kernel(float* out, int frames, long long* cycles) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* rows = reinterpret_cast<float*>(smem);                 // [KC][ROWW], shared by all warps
  float* obuf = rows + KC * ROWW + warp * (KC * 256);           // per warp [KC][4][32] float2
  float* gbuf = rows + KC * ROWW + (blockDim.x >> 5) * (KC * 256) + warp * (KC * 260);
  for (int i = threadIdx.x; i < KC * ROWW; i += blockDim.x) rows[i] = 0.5f + 0.001f * (i % 37);
  for (int i = lane; i < KC * 256; i += 32) obuf[i] = 1.0f;
  __syncthreads();
  uint32_t coloff[8], gph[8];
  for (int k = 0; k < 8; k++) {
    coloff[k] = ((lane * 8 + k) * 7 % 38) * 4;
    gph[k] = ((lane * 8 + k) * 37 % 256) * 4;
  }
  u64 A[4], B[4], F[4], SF[4];
  float a[8], b[8], f[8], sf[8];
  for (int j = 0; j < 4; j++) {
    A[j] = pk(1e-3f * (lane + j), 1e-3f * (lane + j + 4));
    B[j] = pk(1.f, 1.f);
    F[j] = pk(0.5f, 0.5f);
    SF[j] = pk(0.25f, 0.25f);
  }
  for (int k = 0; k < 8; k++) { a[k] = 1e-3f * (lane + k); b[k] = 1.f; f[k] = 0.5f; sf[k] = 0.25f; }
  const long long t0 = clock64();
  for (int t0f = 0; t0f < frames; t0f += KC) {
#pragma unroll
    for (int g = 0; g < KC; g++) {
      const unsigned char* erow = reinterpret_cast<const unsigned char*>(rows + g * ROWW);
      if (PACKED) {
        u64 rh[4], oh[4];
#pragma unroll
        for (int j = 0; j < 4; j++)
          rh[j] = pk(*reinterpret_cast<const float*>(erow + coloff[j]), *reinterpret_cast<const float*>(erow + coloff[j + 4]));
        if (MODE == 2) {
#pragma unroll
          for (int j = 0; j < 4; j++) oh[j] = *reinterpret_cast<const u64*>(obuf + (g * 4 + (3 - j)) * 64 + (31 - lane) * 2);
        }
        const float a_in = __shfl_up_sync(0xffffffffu, hi(A[3]), 1);
        u64 alp[4];
        alp[0] = pk(a_in, lo(A[3]));
#pragma unroll
        for (int j = 1; j < 4; j++) alp[j] = A[j - 1];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const u64 nb = fma2(F[j], alp[j], B[j]);
          const u64 q = fma2(SF[j], alp[j], add2(A[j], B[j]));
          if (MODE == 2) {
            const u64 po = mul2(q, oh[j]);
            *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(gbuf + g * 260) + gph[j]) = lo(po);
            *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(gbuf + g * 260) + gph[j + 4]) = hi(po);
          }
          A[j] = mul2(q, rh[j]);
          B[j] = nb;
          if (MODE == 1) *reinterpret_cast<u64*>(obuf + (g * 4 + j) * 64 + lane * 2) = A[j];
        }
      } else {
        float rh[8], oh[8];
#pragma unroll
        for (int k = 0; k < 8; k++) rh[k] = *reinterpret_cast<const float*>(erow + coloff[k]);
        if (MODE == 2) {
#pragma unroll
          for (int k = 0; k < 8; k++) oh[k] = obuf[(g * 8 + (7 - k)) * 32 + 31 - lane];
        }
        const float a_in = __shfl_up_sync(0xffffffffu, a[7], 1);
#pragma unroll
        for (int k = 7; k >= 0; k--) {
          const float alp = k > 0 ? a[k - 1] : a_in;
          const float nb = fmaf(f[k], alp, b[k]);
          const float q = fmaf(sf[k], alp, a[k] + b[k]);
          if (MODE == 2) *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(gbuf + g * 260) + gph[k]) = q * oh[k];
          a[k] = q * rh[k];
          b[k] = nb;
          if (MODE == 1) obuf[(g * 8 + k) * 32 + lane] = a[k];
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < 4; j++) s += lo(A[j]) + hi(A[j]) + lo(B[j]) + hi(B[j]);
  for (int k = 0; k < 8; k++) s += a[k] + b[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + gbuf[lane];
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// OP 0: conflict-free LDS.32; 1: SHFL; 2: LDS and SHFL alternating; 3: packed add; 4: packed mul; 5: scalar FADD;
// 6: LDS.64 conflict free; 7: gather LDS.32 with a random class pattern (38 classes)
template <int OP>
__global__ void rate_kernel(float* out, int iters, long long* cycles) {
  __shared__ float sm[2048];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = (float)(i & 31);
  __syncthreads();
  float v[8];
  u64 p[8];
  int idx[8];
  for (int i = 0; i < 8; i++) { v[i] = 1.f + lane + i; p[i] = pk(v[i], v[i]); idx[i] = (OP == 7) ? ((lane * 8 + i) * 7 % 38) : lane + 32 * i; }
  const u64 c2 = pk(1e-6f, 1e-6f), m2 = pk(1.000001f, 1.000001f);
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (OP == 0 || OP == 7) v[i] += sm[(idx[i] + it) & 2047];
      if (OP == 1) v[i] = __shfl_up_sync(0xffffffffu, v[i], 1);
      if (OP == 2) { if (i & 1) v[i] = __shfl_up_sync(0xffffffffu, v[i], 1); else v[i] += sm[(idx[i] + it) & 2047]; }
      if (OP == 3) p[i] = add2(p[i], c2);
      if (OP == 4) p[i] = mul2(p[i], m2);
      if (OP == 5) v[i] = v[i] + 1e-6f;
      if (OP == 6) { const float2 w = *reinterpret_cast<const float2*>(&sm[((lane * 2 + 64 * i) + 2 * it) & 2046]); v[i] += w.x + w.y; }
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; i++) s += v[i] + lo(p[i]) + hi(p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// two warps ping-pong over a pair of mbarriers: cycles per round trip
__global__ void mbar_kernel(int iters, long long* cycles) {
  __shared__ __align__(8) u64 bar[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; i++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar[i])));
  }
  __syncthreads();
  const uint32_t mine = (uint32_t)__cvta_generic_to_shared(&bar[warp]), other = (uint32_t)__cvta_generic_to_shared(&bar[warp ^ 1]);
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const uint32_t parity = it & 1;
    if (warp == 0) {
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(other) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(mine), "r"(parity) : "memory");
    }
    if (warp == 1) {
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(other) : "memory");
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
  const int frames = 4000;
  auto report = [&](const char* name, int warps, int blocks) {
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps/SM=%2d x%d : %.1f cycles per frame per warp; SM: %.2f cycles per warp-frame\n", name, warps, blocks,
           (double)h / frames, (double)h / frames / warps);
  };
#define RUN(MODE, PACKED, name)                                                                            \
  for (int warps : {1, 4, 8, 16}) {                                                                        \
    const size_t sh = (KC * ROWW + warps * KC * 256 + warps * KC * 260) * 4;                               \
    cudaFuncSetAttribute(step_kernel<MODE, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); \
    for (int rep = 0; rep < 2; rep++) { step_kernel<MODE, PACKED><<<1, 32 * warps, sh>>>(out, frames, cyc); cudaDeviceSynchronize(); } \
    report(name, warps, 1);                                                                                \
  }
  RUN(0, true, "phase1 packed") RUN(0, false, "phase1 scalar")
  RUN(1, true, "recompute packed") RUN(1, false, "recompute scalar")
  RUN(2, true, "phase2 packed") RUN(2, false, "phase2 scalar")
  const int iters = 4000;
  auto rate = [&](const char* name, auto kern) {
    for (int warps : {1, 4, 8, 16}) {
      for (int rep = 0; rep < 2; rep++) { kern<<<1, 32 * warps>>>(out, iters, cyc); cudaDeviceSynchronize(); }
      long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-28s warps/SM=%2d : %.2f cycles per op per warp; SM rate %.2f warp-ops/cycle\n", name, warps,
             (double)h / iters / 8, (double)warps * 8 * iters / h);
    }
  };
  rate("LDS.32 conflict free", rate_kernel<0>);
  rate("SHFL.32", rate_kernel<1>);
  rate("LDS.32 + SHFL alternating", rate_kernel<2>);
  rate("add.f32x2", rate_kernel<3>);
  rate("mul.f32x2", rate_kernel<4>);
  rate("FADD", rate_kernel<5>);
  rate("LDS.64 conflict free", rate_kernel<6>);
  rate("LDS.32 gather of 38 classes", rate_kernel<7>);
  for (int rep = 0; rep < 2; rep++) { mbar_kernel<<<1, 64>>>(2000, cyc); cudaDeviceSynchronize(); }
  { long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("mbarrier ping-pong: %.1f cycles per round trip (two arrive + two wait)\n", (double)h / 2000); }
  printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
