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
  CK(cudaGetLastError());
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dgemm_tflops_burst\": %.2f, \"dgemm_tflops_sustained\": %.2f, "
         "\"zgemm_tflops_burst\": %.2f, \"zgemm_tflops_sustained\": %.2f, \"dmma_pipe_tflops\": %.2f, \"dfma_pipe_tflops\": %.2f, "
         "\"how\": \"cuBLAS DGEMM 8192^3 (2N^3) and ZGEMM 4096^3 (8N^3): best of 10 and back-to-back for 3 s; DMMA.8x8x4 / DFMA register loops, 8 CTAs x 256 thr per SM\"}\n",
         prop.name, prop.multiProcessorCount, dgemm_burst, dgemm_sust, zgemm_burst, zgemm_sust, dmma_tf, dfma_tf);
  return 0;
}
