// tc_m64_test.cu -- where does an M = 64 tcgen05.mma (cta_group::1, kind::tf32) put its accumulator rows in
// tensor memory, and where does it expect the rows of a TMEM-resident A operand?  Probes:
//   (1) SS form, D at lane offset 0 and 16: dump all 128 lanes, match every lane against the 64 reference rows;
//   (2) TS form: A rows written to TMEM in the lane pattern found by (1).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I<pkg>/csrc -o tc_m64_test tc_m64_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "tc.cuh"

using namespace pinn;

// mode 0: SS, D lane offset `lane_off`.  mode 1: TS, A placed at lanes map(row) + lane_off, D at lane_off.
__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ D,
                                             int mode, int lane_off, int pattern) {
  constexpr int M = 64, N = 64, K = 64;
  constexpr uint32_t LBO_A = M * 16, LBO_B = N * 16;
  extern __shared__ __align__(128) float smem[];
  float* a_hi = smem;
  float* b_hi = a_hi + M * K;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  if (warp == 0) { tc::tmem_alloc(&tmem_base, 256); tc::tmem_relinquish(); }
  for (int idx = tid; idx < M * (K / 4); idx += blockDim.x) {
    const int r = idx % M, kc = idx / M;
    const float4 v = *reinterpret_cast<const float4*>(A + r * K + 4 * kc);
    *reinterpret_cast<float4*>(reinterpret_cast<char*>(a_hi) + kc * LBO_A + r * 16) =
        make_float4(tc::tf32_hi(v.x), tc::tf32_hi(v.y), tc::tf32_hi(v.z), tc::tf32_hi(v.w));
  }
  for (int idx = tid; idx < N * (K / 4); idx += blockDim.x) {
    const int n = idx % N, kc = idx / N;
    const float4 v = *reinterpret_cast<const float4*>(W + n * K + 4 * kc);
    *reinterpret_cast<float4*>(reinterpret_cast<char*>(b_hi) + kc * LBO_B + n * 16) =
        make_float4(tc::tf32_hi(v.x), tc::tf32_hi(v.y), tc::tf32_hi(v.z), tc::tf32_hi(v.w));
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = tmem_base;
  // poison the accumulator region (all 128 lanes, 64 columns) and the A region (columns 128..191)
  {
    float z[16];
    for (int q = 0; q < 16; ++q) z[q] = -777.0f;
    for (int c = 0; c < 64; c += 16) tc::tmem_st16(tb + (static_cast<uint32_t>(warp * 32) << 16) + c, z);
    for (int q = 0; q < 16; ++q) z[q] = 0.0f;
    for (int c = 0; c < 64; c += 16) tc::tmem_st16(tb + 128 + (static_cast<uint32_t>(warp * 32) << 16) + c, z);
    tc::tmem_wait_st();
  }
  if (mode == 1) {
    // place A row r in TMEM lane L(r): pattern 0: L = r (lanes 0..63); pattern 1: L = 32 (r / 16) + r % 16 (16 per quarter)
    // this thread owns TMEM lane `tid`; find the row that maps to it
    int r = -1;
    const int l = tid - lane_off;
    if (pattern == 0) { if (l >= 0 && l < 64) r = l; }
    else { if (l >= 0 && (l & 31) < 16) r = 16 * (l >> 5) + (l & 31); }
    float v[16];
    for (int c = 0; c < 64; c += 16) {
      for (int q = 0; q < 16; ++q) v[q] = r >= 0 ? tc::tf32_hi(A[r * K + c + q]) : 0.f;
      tc::tmem_st16(tb + 128 + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    }
    tc::tmem_wait_st();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_tf32(M, N);
    const uint32_t d = tb + (static_cast<uint32_t>(lane_off) << 16);
    const uint64_t bd = tc::make_desc(tc::smem_u32(b_hi), LBO_B, 128), bstep = (2u * LBO_B) >> 4;
    if (mode == 0) {
      const uint64_t ad = tc::make_desc(tc::smem_u32(a_hi), LBO_A, 128), astep = (2u * LBO_A) >> 4;
      for (int k8 = 0; k8 < K / 8; ++k8) tc::umma_tf32(d, ad + k8 * astep, bd + k8 * bstep, idesc, k8 ? 1u : 0u);
    } else {
      const uint32_t at = tb + 128 + (static_cast<uint32_t>(lane_off) << 16);
      for (int k8 = 0; k8 < K / 8; ++k8) tc::umma_tf32_ts(d, at + 8u * k8, bd + k8 * bstep, idesc, k8 ? 1u : 0u);
    }
    tc::umma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  float v[64];
  for (int c = 0; c < 64; c += 16) tc::tmem_ld16(tb + (static_cast<uint32_t>(warp * 32) << 16) + c, v + c);
  tc::tmem_wait_ld();
  for (int c = 0; c < 64; ++c) D[tid * 64 + c] = v[c];
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 256);
}

int main() {
  std::vector<float> A(64 * 64), W(64 * 64), D(128 * 64);
  srand(3);
  for (auto& v : A) v = (rand() / (float)RAND_MAX) * 2.f - 1.f;
  for (auto& v : W) v = ((rand() / (float)RAND_MAX) * 2.f - 1.f) * 0.125f;
  std::vector<double> ref(64 * 64);
  for (int r = 0; r < 64; ++r)
    for (int n = 0; n < 64; ++n) {
      double s = 0;
      for (int k = 0; k < 64; ++k) s += (double)A[r * 64 + k] * (double)W[n * 64 + k];
      ref[r * 64 + n] = s;
    }
  float *dA, *dW, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  size_t smem = 2 * 64 * 64 * sizeof(float);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct Cfg { int mode, lane_off, pattern; const char* name; };
  Cfg cfgs[] = {{0, 0, 0, "SS  D@lane0"}, {0, 16, 0, "SS  D@lane16"}, {1, 0, 0, "TS  A rows->lanes 0..63, D@lane0"},
                {1, 0, 1, "TS  A 16 rows per quarter, D@lane0"}, {1, 16, 1, "TS  A 16 per quarter @16, D@lane16"}};
  for (auto& c : cfgs) {
    cudaMemset(dD, 0, D.size() * 4);
    probe<<<1, 128, smem>>>(dA, dW, dD, c.mode, c.lane_off, c.pattern);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    printf("%s\n   lane->row: ", c.name);
    int matched = 0;
    for (int l = 0; l < 128; ++l) {
      int best = -1;
      for (int r = 0; r < 64 && best < 0; ++r) {
        double err = 0;
        for (int n = 0; n < 64; ++n) err = fmax(err, fabs(ref[r * 64 + n] - D[l * 64 + n]));
        if (err < 2e-3) best = r;
      }
      if (best >= 0) { ++matched; printf("%d:%d ", l, best); }
    }
    printf("\n   %d lanes hold a reference row; lane 0 col 0 = %g, lane 16 col 0 = %g, lane 64 col 0 = %g\n", matched, D[0], D[16 * 64],
           D[64 * 64]);
  }
  return 0;
}
