// ubench.cu -- design-grounding microbenchmarks for the MLP kernels (not part of the product).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench ubench.cu && ./ubench
// Prints per-SM per-clock rates for: FFMA, FFMA2, uniform LDS.128 (+FFMA2 mix), tanhf, Philox4x32-10.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <math.h>

#define ITERS 4096

__global__ void k_ffma(float* out, float a, float b) {
  float acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  float s = 0; for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
  float2 acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i);
  const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(acc[i], a2, b2);
  float s = 0; for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// uniform LDS.128 feeding FFMA2: R loads per 2*R FFMA2 (the thread-per-sample inner loop)
template <int FMA_PER_LDS>
__global__ void k_lds_mix(float* out) {
  __shared__ float4 w[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) w[i] = make_float4(i * 1e-4f, 1.f, 0.5f, 0.25f);
  __syncthreads();
  float2 acc[8];
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f, i);
  float2 h = make_float2(1.0001f, 0.9999f);
  for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      float4 v = w[(it * 64 + j) & 1023];           // warp-uniform address
#pragma unroll
      for (int q = 0; q < FMA_PER_LDS; ++q) {
        acc[(2 * q) & 7] = __ffma2_rn(make_float2(v.x, v.y), h, acc[(2 * q) & 7]);
        acc[(2 * q + 1) & 7] = __ffma2_rn(make_float2(v.z, v.w), h, acc[(2 * q + 1) & 7]);
      }
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_tanh(float* out, float a) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = (threadIdx.x % 64) * 0.05f - 1.6f + i * 0.01f;
  for (int it = 0; it < ITERS / 4; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = tanhf(v[i] * a + 0.3f);
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ uint4 philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
__global__ void k_philox(float* out) {
  uint32_t s = 0;
  for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 r = philox(1, 2, threadIdx.x, blockIdx.x, it, i);
      s += (r.x < 0x66666666u) + (r.y < 0x66666666u) + (r.z < 0x66666666u) + (r.w < 0x66666666u);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  printf("SMs %d, max clock %.0f MHz (rates below assume the max clock; real clock may be lower)\n", sms, khz / 1e3);
  float* out; cudaMalloc(&out, sizeof(float) * 148 * 8 * 1024);
  const int threads = 512, blocks = sms * 2;
  const double clk = khz * 1e3;
  auto rep = [&](const char* name, float ms, double ops_per_thread) {
    double ops = ops_per_thread * threads * blocks;
    printf("%-34s %8.3f ms  %8.2f ops/clk/SM  %8.2f Tops/s\n", name, ms, ops / (ms * 1e-3) / clk / sms, ops / (ms * 1e-3) / 1e12);
  };
  rep("FFMA   (fma lanes)", time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 16.0 * ITERS);
  rep("FFMA2  (fma lanes = 2/instr)", time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 32.0 * ITERS);
  rep("LDS.128u + 2 FFMA2 (fma lanes)", time_ms([&] { k_lds_mix<1><<<blocks, threads>>>(out); }), (ITERS / 8) * 64.0 * 4);
  rep("LDS.128u + 4 FFMA2 (fma lanes)", time_ms([&] { k_lds_mix<2><<<blocks, threads>>>(out); }), (ITERS / 8) * 64.0 * 8);
  rep("LDS.128u + 8 FFMA2 (fma lanes)", time_ms([&] { k_lds_mix<4><<<blocks, threads>>>(out); }), (ITERS / 8) * 64.0 * 16);
  rep("tanhf (calls)", time_ms([&] { k_tanh<<<blocks, threads>>>(out, 1.01f); }), 8.0 * (ITERS / 4));
  rep("Philox4x32-10 (calls)", time_ms([&] { k_philox<<<blocks, threads>>>(out); }), 4.0 * (ITERS / 4));
  cudaFree(out);
  return 0;
}
