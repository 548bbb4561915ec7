// How fast can one recursion warp go?  Runs the production step code (ctc_fast.cu) in isolation:
// W warps per CTA, one CTA per SM, every warp advancing its own direction over synthetic row records.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iinclude -Ineuralasr_b200/csrc \
//        -o tools/microbench_step tools/microbench_step.cu neuralasr_b200/csrc/nasr_api.cu neuralasr_b200/csrc/ctc_loss.cu neuralasr_b200/csrc/ctc_decode.cu
#include "../neuralasr_b200/csrc/ctc_fast.cu"
#include <cstdio>

using namespace nasr::fast;

template <int NL, int MODE>
__global__ void bench(double* out, long long* cycles, int chunks) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int ROWB = row_bytes(40);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // rows: KC records of doubles ~1.0; obuf/gbuf per warp
  unsigned char* rows = smem;
  uint32_t* obuf = reinterpret_cast<uint32_t*>(smem + KC * ROWB) + warp * KC * NL * 32;
  float* gbuf = reinterpret_cast<float*>(smem + KC * ROWB + 16 * KC * NL * 32 * 4) + warp * KC * (NL * 32 + 4);
  for (int i = threadIdx.x; i < KC * ROWB / 8; i += blockDim.x) reinterpret_cast<double*>(rows)[i] = 1.0 + 1e-3 * (i % 7);
  for (int i = lane; i < KC * NL * 32; i += 32) obuf[i] = 0x3ff00000u;
  __syncthreads();
  Dir<NL> st;
  uint32_t gphys[NL];
  for (int k = 0; k < NL; k++) {
    st.Ab[k] = 1.0 + lane; st.Al[k] = 0.5 + k; st.coloff[k] = ((lane * NL + k) % 38) * 8; st.skip[k] = (maskin[lane * NL + k]) ? 1.0 : 0.0;
    gphys[k] = ((k * 32 + (lane * 5 + k) % 32)) * 4;
  }
  st.E = 0;
  int alarm = 0;
  long long t0 = clock64();
  for (int c = 0; c < chunks; c++) {
    rescale<NL>(st, lane, alarm);
    const double fin = inflow_factor(st.E, lane);
    run_chunk<NL, MODE, ROWB>(st, rows, KC, MODE == STORE_O, obuf, gbuf, fin, -(5 << 20), gphys, lane);
  }
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < NL; k++) s += st.Ab[k] + st.Al[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + alarm;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int NL, int MODE>
void run(const char* name, double* out, long long* cyc) {
  const int chunks = 200;
  cudaFuncSetAttribute(bench<NL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int warps : {1, 4, 8}) {
    bench<NL, MODE><<<148, 32 * warps, 200 * 1024>>>(out, cyc, chunks);
    cudaDeviceSynchronize();
    bench<NL, MODE><<<148, 32 * warps, 200 * 1024>>>(out, cyc, chunks);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-8s NL=%d warps/SM=%2d: %.1f cycles per frame per warp  (%s)\n", name, NL, warps, (double)h / (chunks * KC), cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
  run<7, PLAIN>("plain", out, cyc);
  run<7, STORE_O>("store_o", out, cyc);
  run<7, COMBINE>("combine", out, cyc);
  run<4, PLAIN>("plain", out, cyc);
  return 0;
}
