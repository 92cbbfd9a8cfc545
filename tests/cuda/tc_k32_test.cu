// tc_k32_test.cu -- probe: does a K-MAJOR TF32 UMMA operand accept the SWIZZLE_128B_BASE32B layout (CUTLASS lists it as
// MN-major only)?  If it does, ONE [sample][feature] image (128-byte rows, 32-byte chunk index XOR (row & 3)) can serve both
// the forward / dgrad products (K = feature: K-major) and the weight-gradient product (K = sample: MN-major, tc_mn32_test.cu).
//     D[m][n] = sum_k A[m][k] * B[n][k],  A: [128][64] and B: [64][64], rows of 64 features = two 128-byte blocks
// Descriptor: SBO = 512 B (the 4-row swizzle atom), LBO = block stride (not used by the hardware for a swizzled K-major operand),
// the K = 8 slab of an MMA = one 32-byte chunk: start address + 32 B per slab.  Measured on B200 / CUDA 12.9: works
// (norm-rel 6.0e-4 = one TF32 product).  (SBO = 1024 reads past the operand: illegal address.)
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

__global__ void __launch_bounds__(128) k32_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out, int variant) {
  constexpr uint32_t BLKA = 128 * 128, BLKB = 64 * 128;      // one 32-feature block: rows x 128 B
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* a_s = sm;                   // 2 blocks of A (features 0..31, 32..63)
  unsigned char* b_s = sm + 2 * BLKA;        // 2 blocks of B
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tc::uniform_warp_idx();
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  __syncwarp();
  if (warp == 0) { tc::tmem_alloc(&tmem_base, 64); tc::tmem_relinquish(); }
  for (int blk = 0; blk < 2; ++blk) {
    {
      const int r = tid;
      const float* src = A + r * 64 + 32 * blk;
      unsigned char* dst = a_s + blk * BLKA + r * 128;
      for (int c = 0; c < 4; ++c) {
        unsigned char* p = dst + ((c ^ (r & 3)) << 5);
        *reinterpret_cast<float4*>(p) = *reinterpret_cast<const float4*>(src + 8 * c);
        *reinterpret_cast<float4*>(p + 16) = *reinterpret_cast<const float4*>(src + 8 * c + 4);
      }
    }
    if (tid < 64) {
      const int r = tid;
      const float* src = B + r * 64 + 32 * blk;
      unsigned char* dst = b_s + blk * BLKB + r * 128;
      for (int c = 0; c < 4; ++c) {
        unsigned char* p = dst + ((c ^ (r & 3)) << 5);
        *reinterpret_cast<float4*>(p) = *reinterpret_cast<const float4*>(src + 8 * c);
        *reinterpret_cast<float4*>(p + 16) = *reinterpret_cast<const float4*>(src + 8 * c + 4);
      }
    }
  }
  tc::fence_proxy_async(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  if (warp == 0) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_tf32(128, 64, false, false);
      uint32_t lbo_a, sbo_a, lbo_b, sbo_b;
      if (variant == 0) { lbo_a = BLKA; sbo_a = 512; lbo_b = BLKB; sbo_b = 512; }
      else if (variant == 1) { lbo_a = BLKA; sbo_a = 1024; lbo_b = BLKB; sbo_b = 1024; }
      else if (variant == 2) { lbo_a = 512; sbo_a = BLKA; lbo_b = 512; sbo_b = BLKB; }
      else { lbo_a = 16; sbo_a = 1024; lbo_b = 16; sbo_b = 1024; }
      // K = 8 per MMA = one 32-byte chunk of a row; 4 MMAs per 32-feature block, then the next block
      uint32_t acc = 0;
      for (int blk = 0; blk < 2; ++blk)
        for (int k8 = 0; k8 < 4; ++k8) {
          const uint64_t ad = make_desc_sw(tc::smem_u32(a_s) + blk * BLKA + k8 * 32, lbo_a, sbo_a, 1u);
          const uint64_t bd = make_desc_sw(tc::smem_u32(b_s) + blk * BLKB + k8 * 32, lbo_b, sbo_b, 1u);
          tc::umma_tf32(tmem_base, ad, bd, idesc, acc);
          acc = 1;
        }
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
  std::vector<float> A(128 * 64), B(64 * 64);
  srand(5);
  for (auto& v : A) v = (rand() / (float)RAND_MAX) * 2.f - 1.f;
  for (auto& v : B) v = (rand() / (float)RAND_MAX) * 2.f - 1.f;
  float *dA, *dB, *dO;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 2 * 128 * 128 + 2 * 64 * 128;
  cudaFuncSetAttribute(k32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  bool any = false;
  for (int variant = 0; variant < 1; ++variant) {
    cudaMemset(dO, 0, 128 * 64 * 4);
    k32_kernel<<<1, 128, smem>>>(dA, dB, dO, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error: %s\n", variant, cudaGetErrorString(e)); return 1; }
    std::vector<float> O(128 * 64);
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0, ref_max = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 64; ++n) {
        double ref = 0;
        for (int k = 0; k < 64; ++k) ref += (double)A[m * 64 + k] * (double)B[n * 64 + k];
        err = fmax(err, fabs(ref - O[m * 64 + n])); ref_max = fmax(ref_max, fabs(ref));
      }
    printf("variant %d: norm-rel error %.3e   D[0][0..3] = %g %g %g %g\n", variant, err / ref_max, O[0], O[1], O[2], O[3]);
    if (err / ref_max < 2e-3) { any = true; printf("K-major SWIZZLE_128B_BASE32B WORKS with variant %d\n", variant); }
  }
  printf(any ? "TC_K32_TEST: supported\n" : "TC_K32_TEST: not supported (no variant matched)\n");
  return 0;
}
