// Microbenchmarks that decide the arithmetic of the CTC recursion on B200: issue rate and latency of
// DFMA/DADD/DMUL, FFMA, packed FFMA2, F2F, shuffles and LDS, per warp, for 1..8 warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int CHAINS>
__global__ void k(double* out, int iters, long long* cycles) {
  double a[CHAINS];
  float f[CHAINS];
  float2 p[CHAINS];
  for (int i = 0; i < CHAINS; i++) { a[i] = 1.0 + threadIdx.x * 1e-9 + i; f[i] = 1.0f + i + threadIdx.x * 1e-6f; p[i] = make_float2(f[i], f[i] + 1); }
  const double m = 1.0000000001, c = 1e-12;
  const float mf = 1.000001f, cf = 1e-7f;
  __shared__ double sm[1024];
  sm[threadIdx.x % 1024] = threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      if (OP == 0) a[i] = fma(a[i], m, c);
      if (OP == 1) a[i] = a[i] + c;
      if (OP == 2) a[i] = a[i] * m;
      if (OP == 3) f[i] = fmaf(f[i], mf, cf);
      if (OP == 4) {
        unsigned long long r, x = *reinterpret_cast<unsigned long long*>(&p[i]);
        unsigned long long mm, cc; float2 m2 = make_float2(mf, mf), c2 = make_float2(cf, cf);
        mm = *reinterpret_cast<unsigned long long*>(&m2); cc = *reinterpret_cast<unsigned long long*>(&c2);
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(mm), "l"(cc));
        p[i] = *reinterpret_cast<float2*>(&r);
      }
      if (OP == 5) a[i] = (double)(float)a[i] + c;            // F2F both ways
      if (OP == 6) a[i] = __shfl_up_sync(0xffffffffu, a[i], 1);  // 2 SHFL
      if (OP == 7) a[i] = sm[((int)a[i] + i) & 1023];          // dependent LDS.64
      if (OP == 8) f[i] = __shfl_up_sync(0xffffffffu, f[i], 1);
      if (OP == 9) a[i] = fmax(a[i], c) + c;                   // double max + add
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CHAINS; i++) s += a[i] + f[i] + p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP, int CHAINS>
void run(const char* name, double* out, long long* cyc) {
  const int iters = 2000;
  for (int warps : {1, 4, 8, 16}) {
    k<OP, CHAINS><<<1, 32 * warps>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    k<OP, CHAINS><<<1, 32 * warps>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-8s chains=%d warps/SM=%2d : %.2f cycles per op per warp (%.2f cycles per iteration); SM rate %.2f warp-ops/cycle\n", name, CHAINS, warps,
           (double)h / iters / CHAINS, (double)h / iters, (double)warps * CHAINS * iters / h);
  }
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  run<0, 1>("DFMA", out, cyc); run<0, 8>("DFMA", out, cyc);
  run<1, 1>("DADD", out, cyc); run<1, 8>("DADD", out, cyc);
  run<2, 8>("DMUL", out, cyc);
  run<3, 1>("FFMA", out, cyc); run<3, 8>("FFMA", out, cyc);
  run<4, 1>("FFMA2", out, cyc); run<4, 8>("FFMA2", out, cyc);
  run<5, 1>("F2Fx2+DADD", out, cyc); run<5, 8>("F2Fx2+DADD", out, cyc);
  run<6, 1>("SHFL64", out, cyc); run<6, 8>("SHFL64", out, cyc);
  run<8, 1>("SHFL32", out, cyc); run<8, 8>("SHFL32", out, cyc);
  run<7, 1>("LDS64dep", out, cyc);
  run<9, 8>("DMAX+DADD", out, cyc);
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
