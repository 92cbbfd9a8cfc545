// tc_mn32_test.cu -- standalone probe: MN-major TF32 UMMA operands in the SWIZZLE_128B_BASE32B layout (the only
// shared-memory layout CUTLASS lists for MN-major tf32), i.e. the [sample][feature] arrays an epilogue thread writes
// with 128-bit stores, used directly for a contraction over SAMPLES (the weight-gradient product):
//     D[m][n] = sum_s A[s][m] * B[s][n],   A: [128 samples][128 features], B: [128 samples][64 features]
// Layout per 32-feature block: rows of 128 B (one sample each), 4-row atoms of 512 B, the 32-byte chunk index of a row
// XOR-ed with (sample % 4)  (Swizzle<2,5,2> on the byte address); blocks LBO apart, 4-sample groups SBO apart.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I ../../physics-informed-neural-network-for-explainable-fault-diagnosis-in-fuel-cells_b200/csrc -o tc_mn32_test tc_mn32_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "tc.cuh"
using namespace pinn;

__device__ __forceinline__ uint64_t make_desc_sw(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= static_cast<uint64_t>(layout_type & 7u) << 61;
  return d;
}

// variant 0: LBO = block stride, SBO = k-group stride (CUTLASS's make_umma_desc);  variant 1: swapped
__global__ void __launch_bounds__(128) mn32_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out, int variant) {
  constexpr int S = 128;                       // samples (K)
  constexpr uint32_t BLK = S * 128;            // bytes of one 32-feature block: 128 samples x 128 B
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* a_s = sm;                     // 4 blocks (M = 128)
  unsigned char* b_s = sm + 4 * BLK;           // 2 blocks (N = 64)
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tc::uniform_warp_idx();
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  __syncwarp();
  if (warp == 0) { tc::tmem_alloc(&tmem_base, 64); tc::tmem_relinquish(); }
  // thread = sample: writes its row of every block as 32-byte chunks, chunk index XOR (sample & 3)
  const int s = tid;
  for (int blk = 0; blk < 6; ++blk) {
    const float* src = blk < 4 ? A + s * 128 + 32 * blk : B + s * 64 + 32 * (blk - 4);
    unsigned char* dst = (blk < 4 ? a_s + blk * BLK : b_s + (blk - 4) * BLK) + s * 128;
    for (int c = 0; c < 4; ++c) {
      const float4 v0 = *reinterpret_cast<const float4*>(src + 8 * c), v1 = *reinterpret_cast<const float4*>(src + 8 * c + 4);
      unsigned char* p = dst + ((c ^ (s & 3)) << 5);
      *reinterpret_cast<float4*>(p) = v0;
      *reinterpret_cast<float4*>(p + 16) = v1;
    }
  }
  tc::fence_proxy_async(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  if (warp == 0) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_tf32(128, 64, true, true);
      const uint32_t lbo = variant == 0 ? BLK : 512u, sbo = variant == 0 ? 512u : BLK;
      const uint64_t ad = make_desc_sw(tc::smem_u32(a_s), lbo, sbo, 1u), bd = make_desc_sw(tc::smem_u32(b_s), lbo, sbo, 1u);
      for (int k8 = 0; k8 < S / 8; ++k8)      // one MMA = 8 samples = two 4-sample groups = 1024 B further down every block
        tc::umma_tf32(tmem_base, ad + k8 * (1024 >> 4), bd + k8 * (1024 >> 4), idesc, k8 ? 1u : 0u);
      tc::umma_commit(&bar);
    }
    __syncwarp();
  }
  tc::mbar_wait(&bar, 0);
  __syncwarp();
  tc::fence_after_sync();
  float v[64];
  const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
  for (int c = 0; c < 64; c += 16) tc::tmem_ld16(tmem_base + lane_off + c, v + c);
  tc::tmem_wait_ld();
  for (int c = 0; c < 64; ++c) out[tid * 64 + c] = v[c];
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 64);
}

int main() {
  std::vector<float> A(128 * 128), B(128 * 64);
  srand(3);
  for (auto& v : A) v = (rand() / (float)RAND_MAX) * 2.f - 1.f;
  for (auto& v : B) v = (rand() / (float)RAND_MAX) * 2.f - 1.f;
  float *dA, *dB, *dO;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 6 * 128 * 128;
  cudaFuncSetAttribute(mn32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  bool any = false;
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dO, 0, 128 * 64 * 4);
    mn32_kernel<<<1, 128, smem>>>(dA, dB, dO, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error: %s\n", variant, cudaGetErrorString(e)); return 1; }
    std::vector<float> O(128 * 64);
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0, ref_max = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 64; ++n) {
        double ref = 0;
        for (int s = 0; s < 128; ++s) ref += (double)A[s * 128 + m] * (double)B[s * 64 + n];
        err = fmax(err, fabs(ref - O[m * 64 + n])); ref_max = fmax(ref_max, fabs(ref));
      }
    printf("variant %d (%s): norm-rel error %.3e   D[0][0..3] = %g %g %g %g\n", variant,
           variant == 0 ? "LBO = block stride, SBO = k-group stride" : "swapped", err / ref_max, O[0], O[1], O[2], O[3]);
    if (err / ref_max < 2e-3) { any = true; printf("MN32 layout WORKS with variant %d\n", variant); }
  }
  printf(any ? "TC_MN32_TEST PASS\n" : "TC_MN32_TEST FAIL (no variant matched)\n");
  return 0;
}
