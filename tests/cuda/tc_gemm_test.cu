// tc_gemm_test.cu -- standalone check of the tcgen05 building blocks in csrc/tc.cuh:
// D[128 x N] = A[128 x 64] * W[N x 64]^T via 3xTF32 (N = 64 and N = 48), vs fp64 on the host.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I<pkg>/csrc -o tc_gemm_test tc_gemm_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "tc.cuh"

using namespace pinn;

template <int N>
__global__ void __launch_bounds__(128) gemm_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ D) {
  constexpr int M = 128, K = 64;
  constexpr uint32_t LBO_A = M * 16, LBO_B = N * 16;
  extern __shared__ __align__(128) float smem[];
  float* a_hi = smem;                      // M*K floats
  float* a_lo = a_hi + M * K;
  float* b_hi = a_lo + M * K;              // N*K floats
  float* b_lo = b_hi + N * K;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  if (warp == 0) { tc::tmem_alloc(&tmem_base, 64); tc::tmem_relinquish(); }
  // stage operands: thread = row of A; B rows split over the threads
  for (int kc = 0; kc < K / 4; ++kc)
    tc::store_split4(a_hi, a_lo, LBO_A, tid, kc, *reinterpret_cast<const float4*>(A + tid * K + 4 * kc));
  for (int idx = tid; idx < N * (K / 4); idx += blockDim.x) {
    const int n = idx % N, kc = idx / N;
    tc::store_split4(b_hi, b_lo, LBO_B, n, kc, *reinterpret_cast<const float4*>(W + n * K + 4 * kc));
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = tmem_base;
  if (tid == 0) {
    tc::issue_3xtf32<K>(taddr, tc::make_desc(tc::smem_u32(a_hi), LBO_A, 128), tc::make_desc(tc::smem_u32(a_lo), LBO_A, 128), LBO_A,
                        tc::make_desc(tc::smem_u32(b_hi), LBO_B, 128), tc::make_desc(tc::smem_u32(b_lo), LBO_B, 128), LBO_B,
                        tc::make_idesc_tf32(M, N));
    tc::umma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  float v[N];
#pragma unroll
  for (int c = 0; c < N; c += 16) tc::tmem_ld16(taddr + (static_cast<uint32_t>(warp * 32) << 16) + c, v + c);
  tc::tmem_wait_ld();
#pragma unroll
  for (int c = 0; c < N; ++c) D[tid * N + c] = v[c];
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(taddr, 64);
}

template <int N>
double run(const std::vector<float>& A, const std::vector<float>& W) {
  float *dA, *dW, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, N * 64 * 4); cudaMalloc(&dD, 128 * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), N * 64 * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, 128 * N * 4);
  size_t smem = (2 * 128 * 64 + 2 * N * 64) * sizeof(float);
  cudaFuncSetAttribute(gemm_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  gemm_kernel<N><<<1, 128, smem>>>(dA, dW, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d CUDA error: %s\n", N, cudaGetErrorString(e)); return 1e9; }
  std::vector<float> D(128 * N);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < 64; ++k) ref += (double)A[r * 64 + k] * (double)W[n * 64 + k];
      maxerr = fmax(maxerr, fabs(ref - D[r * N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("N=%d  max|err| = %.3e  max|ref| = %.3e  norm-rel = %.3e   D[0][0..3] = %g %g %g %g\n", N, maxerr, maxref,
         maxerr / maxref, D[0], D[1], D[2], D[3]);
  cudaFree(dA); cudaFree(dW); cudaFree(dD);
  return maxerr / maxref;
}

int main() {
  std::vector<float> A(128 * 64), W(64 * 64);
  srand(1);
  for (auto& v : A) v = (rand() / (float)RAND_MAX) * 2.f - 1.f;
  for (auto& v : W) v = ((rand() / (float)RAND_MAX) * 2.f - 1.f) * 0.125f;
  double e64 = run<64>(A, W);
  double e48 = run<48>(A, W);
  bool ok = e64 < 2e-6 && e48 < 2e-6;
  printf(ok ? "TC_GEMM_TEST PASS\n" : "TC_GEMM_TEST FAIL\n");
  return ok ? 0 : 1;
}
