// Measures the FP64 roofline denominators on the box (MEASURED_PEAKS.json has no FP64 entry):
//   cuBLAS DGEMM / ZGEMM (burst = best of 10, sustained = back to back for ~3 s),
//   a register-resident DMMA.8x8x4 issue-rate loop and a DFMA loop (pipe peaks).
// Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp64_peak.cu -lcublas
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_dmma(double* out, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0 + blockIdx.x * 1e-6;
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
                   : "+d"(c[2 * i]), "+d"(c[2 * i + 1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// other FP64 MMA shapes (register-resident issue loops, 4 independent accumulator sets per warp) and the latency of a
// dependent accumulator chain of m8n8k4 (one warp per CTA, one chain)
template <int SHAPE>
__global__ void k_dmma_shape(double* out, int iters) {
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = 1.0 + blockIdx.x * 1e-6 + i;
  double c[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (SHAPE == 4)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5},{%6},{%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
      else if (SHAPE == 8)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7},{%8,%9},{%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7,%8,%9,%10,%11},{%12,%13,%14,%15},{%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
                       "d"(b[2]), "d"(b[3]));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma_chain(double* out, int iters, long long* cycles) {
  double a = threadIdx.x * 1e-3, b = 1.0;
  double c[2] = {0.0, 0.0};
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
  }
  const long long t1 = clock64();
  out[threadIdx.x] = c[0] + c[1];
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

__global__ void k_dfma(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * blockIdx.x;
  double c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cublasHandle_t h;
  if (cublasCreate(&h) != CUBLAS_STATUS_SUCCESS) { printf("{\"error\": \"cublasCreate\"}\n"); return 1; }
  const int N = 8192;
  double *A, *B, *Cm;
  CK(cudaMalloc(&A, sizeof(double) * N * N)); CK(cudaMalloc(&B, sizeof(double) * N * N)); CK(cudaMalloc(&Cm, sizeof(double) * N * N));
  std::vector<double> hst((size_t)N * N);
  for (size_t i = 0; i < hst.size(); ++i) hst[i] = (double)rand() / RAND_MAX - 0.5;
  CK(cudaMemcpy(A, hst.data(), sizeof(double) * N * N, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(B, hst.data(), sizeof(double) * N * N, cudaMemcpyHostToDevice));
  double one = 1.0, zero = 0.0;
  double dgemm_burst = 0, dgemm_sust = 0, zgemm_burst = 0, zgemm_sust = 0;
  for (int i = 0; i < 3; ++i) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, N, N, N, &one, A, N, B, N, &zero, Cm, N);
  cudaDeviceSynchronize();
  for (int i = 0; i < 10; ++i) {
    cudaEventRecord(e0); cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, N, N, N, &one, A, N, B, N, &zero, Cm, N); cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    double tf = 2.0 * N * N * N / (time_ms(e0, e1) * 1e-3) / 1e12;
    if (tf > dgemm_burst) dgemm_burst = tf;
  }
  { int reps = 0; float tot = 0; cudaEventRecord(e0);
    while (true) { for (int i = 0; i < 10; ++i) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, N, N, N, &one, A, N, B, N, &zero, Cm, N);
      reps += 10; cudaEventRecord(e1); cudaEventSynchronize(e1); tot = time_ms(e0, e1); if (tot > 3000) break; }
    dgemm_sust = 2.0 * N * N * N * reps / (tot * 1e-3) / 1e12; }
  const int NZ = 4096;  // complex: N*N*16 bytes
  cuDoubleComplex zone = make_cuDoubleComplex(1, 0), zzero = make_cuDoubleComplex(0, 0);
  cuDoubleComplex *ZA = (cuDoubleComplex*)A, *ZB = (cuDoubleComplex*)B, *ZC = (cuDoubleComplex*)Cm;
  for (int i = 0; i < 3; ++i) cublasZgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, NZ, NZ, NZ, &zone, ZA, NZ, ZB, NZ, &zzero, ZC, NZ);
  cudaDeviceSynchronize();
  for (int i = 0; i < 10; ++i) {
    cudaEventRecord(e0); cublasZgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, NZ, NZ, NZ, &zone, ZA, NZ, ZB, NZ, &zzero, ZC, NZ); cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    double tf = 8.0 * NZ * NZ * NZ / (time_ms(e0, e1) * 1e-3) / 1e12;
    if (tf > zgemm_burst) zgemm_burst = tf;
  }
  { int reps = 0; float tot = 0; cudaEventRecord(e0);
    while (true) { for (int i = 0; i < 10; ++i) cublasZgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, NZ, NZ, NZ, &zone, ZA, NZ, ZB, NZ, &zzero, ZC, NZ);
      reps += 10; cudaEventRecord(e1); cudaEventSynchronize(e1); tot = time_ms(e0, e1); if (tot > 3000) break; }
    zgemm_sust = 8.0 * NZ * NZ * NZ * reps / (tot * 1e-3) / 1e12; }
  // pipe microbenchmarks
  double* out;
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
  double dmma_tf = 0, dfma_tf = 0;
  const int iters = 20000;
  k_dmma<<<blocks, threads>>>(out, 100); cudaDeviceSynchronize();
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); k_dmma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    double flops = (double)blocks * (threads / 32) * iters * 8.0 * 512.0;
    double tf = flops / (time_ms(e0, e1) * 1e-3) / 1e12;
    if (tf > dmma_tf) dmma_tf = tf;
  }
  k_dfma<<<blocks, threads>>>(out, 100); cudaDeviceSynchronize();
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); k_dfma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    double flops = (double)blocks * threads * iters * 8.0 * 2.0;
    double tf = flops / (time_ms(e0, e1) * 1e-3) / 1e12;
    if (tf > dfma_tf) dfma_tf = tf;
  }
  // MMA shapes and the dependent-chain latency of m8n8k4
  double shape_tf[3] = {0, 0, 0};
  for (int sh = 0; sh < 3; ++sh) {
    const double fma_per = sh == 0 ? 512.0 : sh == 1 ? 1024.0 : 2048.0;
    for (int r = 0; r < 4; ++r) {
      cudaEventRecord(e0);
      if (sh == 0) k_dmma_shape<4><<<blocks, threads>>>(out, r == 0 ? 100 : iters / 4);
      else if (sh == 1) k_dmma_shape<8><<<blocks, threads>>>(out, r == 0 ? 100 : iters / 4);
      else k_dmma_shape<16><<<blocks, threads>>>(out, r == 0 ? 100 : iters / 8);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      if (r == 0) continue;
      const double n = sh == 2 ? iters / 8 : iters / 4;
      double tf = (double)blocks * (threads / 32) * n * 4.0 * fma_per * 2.0 / (time_ms(e0, e1) * 1e-3) / 1e12;
      if (tf > shape_tf[sh]) shape_tf[sh] = tf;
    }
  }
  long long* dcyc; long long hcyc = 0;
  CK(cudaMalloc(&dcyc, sizeof(long long)));
  k_dmma_chain<<<1, 32>>>(out, 1000, dcyc); cudaDeviceSynchronize();
  k_dmma_chain<<<1, 32>>>(out, 1000, dcyc);
  CK(cudaMemcpy(&hcyc, dcyc, sizeof(long long), cudaMemcpyDeviceToHost));
  const double chain_cycles = (double)hcyc / 8000.0;
  CK(cudaGetLastError());
  printf("{\"dmma_m16n8k4_tflops\": %.2f, \"dmma_m16n8k8_tflops\": %.2f, \"dmma_m16n8k16_tflops\": %.2f, \"dmma_m8n8k4_dependent_chain_cycles\": %.1f}\n",
         shape_tf[0], shape_tf[1], shape_tf[2], chain_cycles);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dgemm_tflops_burst\": %.2f, \"dgemm_tflops_sustained\": %.2f, "
         "\"zgemm_tflops_burst\": %.2f, \"zgemm_tflops_sustained\": %.2f, \"dmma_pipe_tflops\": %.2f, \"dfma_pipe_tflops\": %.2f, "
         "\"how\": \"cuBLAS DGEMM 8192^3 (2N^3) and ZGEMM 4096^3 (8N^3): best of 10 and back-to-back for 3 s; DMMA.8x8x4 / DFMA register loops, 8 CTAs x 256 thr per SM\"}\n",
         prop.name, prop.multiProcessorCount, dgemm_burst, dgemm_sust, zgemm_burst, zgemm_sust, dmma_tf, dfma_tf);
  return 0;
}
