// Which resource saturates the recursion step?  The production step's instruction mix with pieces
// switched off: FLAGS bit0 = emission loads from shared memory, bit1 = shuffle, bit2 = LOP3 select.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int NL = 7;
template <int FLAGS>
__global__ void k(double* out, const uint32_t* maskin, long long* cycles, int frames) {
  __shared__ double rows[8 * 64];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8 * 64; i += blockDim.x) rows[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double Ab[NL], Al[NL];
  uint32_t mask[NL], off[NL];
  for (int kk = 0; kk < NL; kk++) { Ab[kk] = 1.0 + lane; Al[kk] = 0.5 + kk; mask[kk] = maskin[lane * NL + kk]; off[kk] = (lane * 5 + kk * 3) % 40; }
  const double fin = 1.0000001;
  long long t0 = clock64();
  double a_raw = 0.25;
  for (int f = 0; f < frames; f++) {
    double r[NL];
#pragma unroll
    for (int kk = 0; kk < NL; kk++) r[kk] = (FLAGS & 1) ? rows[(f & 7) * 64 + off[kk]] : 1.0000001;
    double a_next = a_raw;
#pragma unroll
    for (int kk = NL - 1; kk >= 0; kk--) {
      const double alp = kk > 0 ? Al[kk - 1] : (lane ? a_raw * fin : 0.0);
      const double nb = Ab[kk] + alp;
      double w;
      if (FLAGS & 4) {
        const uint32_t m = mask[kk];
        w = __hiloint2double((int)(((uint32_t)__double2hiint(nb) & m) | ((uint32_t)__double2hiint(Ab[kk]) & ~m)),
                             (int)(((uint32_t)__double2loint(nb) & m) | ((uint32_t)__double2loint(Ab[kk]) & ~m)));
      } else {
        w = nb;
      }
      const double q = Al[kk] + w;
      Al[kk] = q * r[kk];
      if (kk == NL - 1) a_next = (FLAGS & 2) ? __shfl_up_sync(0xffffffffu, Al[NL - 1], 1) : Al[NL - 1];
      Ab[kk] = nb;
    }
    a_raw = a_next;
    if ((f & 7) == 7) {  // keep values bounded
#pragma unroll
      for (int kk = 0; kk < NL; kk++) { Ab[kk] *= 1e-3; Al[kk] *= 1e-3; }
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int kk = 0; kk < NL; kk++) s += Ab[kk] + Al[kk];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int FLAGS>
void run(double* out, uint32_t* m, long long* cyc) {
  const int frames = 4000;
  for (int warps : {1, 4, 8, 16}) {
    k<FLAGS><<<148, 32 * warps>>>(out, m, cyc, frames);
    cudaDeviceSynchronize();
    k<FLAGS><<<148, 32 * warps>>>(out, m, cyc, frames);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("lds=%d shfl=%d lop=%d warps/SM=%2d: %.1f cycles per frame per warp\n", FLAGS & 1, (FLAGS >> 1) & 1, (FLAGS >> 2) & 1, warps, (double)h / frames);
  }
}
int main() {
  double* out; long long* cyc; uint32_t* m;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8); cudaMalloc(&m, 4 * 32 * NL);
  uint32_t hm[32 * NL]; for (int i = 0; i < 32 * NL; i++) hm[i] = (i % 10) ? 0xffffffffu : 0u;
  cudaMemcpy(m, hm, sizeof(hm), cudaMemcpyHostToDevice);
  run<0>(out, m, cyc); run<4>(out, m, cyc); run<1>(out, m, cyc); run<2>(out, m, cyc); run<7>(out, m, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
