// net.cuh -- the stack-voltage DNN (01:389-438) as a per-sample device program.
//
// Mapping ("thread-per-sample"): one thread owns one sample; its activation vector
// lives in a private column of shared memory (SMALL nets, H <= 64: conflict-free,
// stride = block size) or of a global scratch (LARGE nets).  Weights are read with
// warp-uniform 128-bit loads -- from a shared-memory arena holding ALL layers
// (SMALL) or through the read-only path from L2 (LARGE) -- and consumed by packed
// fp32x2 FMAs (FFMA2, sm_100) paired along the contraction index so neither operand
// needs duplicating.  All arithmetic is fp32 (parity 1e-5).
#pragma once
#include "common.cuh"

namespace pinn {

PINN_HD float2 ffma2(float2 a, float2 b, float2 c) {
#ifdef __CUDA_ARCH__
  return __ffma2_rn(a, b, c);
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}

template <bool WG>
PINN_HD float4 ldw4(const float* p) {
#ifdef __CUDA_ARCH__
  if constexpr (WG) return __ldg(reinterpret_cast<const float4*>(p));
#endif
  return *reinterpret_cast<const float4*>(p);
}
template <bool WG>
PINN_HD float ldw1(const float* p) {
#ifdef __CUDA_ARCH__
  if constexpr (WG) return __ldg(p);
#endif
  return *p;
}

// A sample's private column: element k at p[k * stride].
struct Col {
  float* p;
  int stride;
  PINN_HD float get(int k) const { return p[static_cast<size_t>(k) * stride]; }
  PINN_HD void set(int k, float v) const { p[static_cast<size_t>(k) * stride] = v; }
  PINN_HD Col at(int k0) const { return Col{p + static_cast<size_t>(k0) * stride, stride}; }
};

// Weight view: WG=false -> shared arena + layout offsets, WG=true -> global pointers.
template <bool WG>
struct Weights {
  const pinn_net_t* net;   // kernel-parameter copy (global pointers)
  const float* arena;      // shared-memory arena (padded flat layout), WG=false
  const ParamLayout* lay;
  PINN_HD const float* W(int l) const { return WG ? net->W[l] : arena + lay->offW[l]; }
  PINN_HD const float* b(int l) const { return WG ? net->b[l] : arena + lay->offb[l]; }
  PINN_HD const float* Wp() const { return WG ? net->Wp : arena + lay->offWp; }
  PINN_HD const float* bp() const { return WG ? net->bp : arena + lay->offbp; }
  PINN_HD const float* Wv0() const { return WG ? net->Wv0 : arena + lay->offWv0; }
  PINN_HD const float* bv0() const { return WG ? net->bv0 : arena + lay->offbv0; }
  PINN_HD const float* Wv1() const { return WG ? net->Wv1 : arena + lay->offWv1; }
  PINN_HD const float* bv1() const { return WG ? net->bv1 : arena + lay->offbv1; }
  PINN_HD const float* Wv2() const { return WG ? net->Wv2 : arena + lay->offWv2; }
  PINN_HD const float* bv2() const { return WG ? net->bv2 : arena + lay->offbv2; }
};

#ifdef __CUDACC__
// Cooperative copy of every tensor into the shared arena (once per CTA).
__device__ inline void stage_tensor(float* dst, const float* src, int count) {
  for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = __ldg(src + i);
}
__device__ inline void stage_tensor_scaled(float* dst, const float* src, int count, float c) {
  for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = __ldg(src + i) * c;
}
__device__ inline void stage_weights(float* arena, const pinn_net_t& net, const ParamLayout& lay) {
  const int H = lay.H;
  for (int l = 0; l < lay.L; ++l) {
    stage_tensor(arena + lay.offW[l], net.W[l], H * (l == 0 ? PINN_N_IN : H));
    stage_tensor(arena + lay.offb[l], net.b[l], H);
  }
  stage_tensor(arena + lay.offWp, net.Wp, H);
  stage_tensor(arena + lay.offbp, net.bp, 1);
  stage_tensor(arena + lay.offWv0, net.Wv0, (H / 2) * H);
  stage_tensor(arena + lay.offbv0, net.bv0, H / 2);
  stage_tensor(arena + lay.offWv1, net.Wv1, (H / 4) * (H / 2));
  stage_tensor(arena + lay.offbv1, net.bv1, H / 4);
  stage_tensor(arena + lay.offWv2, net.Wv2, H / 4);
  stage_tensor(arena + lay.offbv2, net.bv2, 1);
}
#endif

// acc{A,B}[jj] += sum over a KC-chunk of W[j0+jj][k] * h[k]; h given as KC/2 pairs.
// Two accumulator pairs per output keep 2*JB independent FFMA2 chains in flight.
template <int KC, int JB, bool WG>
PINN_HD void dot_chunk(const float2* h, const float* Wrow0, int row_stride, float2* accA, float2* accB) {
#pragma unroll
  for (int k4 = 0; k4 < KC / 4; ++k4) {
#pragma unroll
    for (int jj = 0; jj < JB; ++jj) {
      float4 w = ldw4<WG>(Wrow0 + static_cast<size_t>(jj) * row_stride + 4 * k4);
      accA[jj] = ffma2(make_float2(w.x, w.y), h[2 * k4], accA[jj]);
      accB[jj] = ffma2(make_float2(w.z, w.w), h[2 * k4 + 1], accB[jj]);
    }
  }
}

// z[j] = bias[j] + sum_k W[j][k] * in[k] for j in [0,NOUT), handed to epi in blocks
// of JB.  K <= 64: inputs are preloaded into registers, so `epi` may overwrite the
// input column (in-place).  K > 64: inputs are re-read per chunk; epi must write
// elsewhere.
template <int K, int NOUT, int JB, bool WG, class Epi>
PINN_HD void tps_dense(Col in, const float* W, const float* bias, Epi epi) {
  static_assert(NOUT % JB == 0 && K % 4 == 0, "tile");
  constexpr int KC = K <= 64 ? K : 64;
  if constexpr (K <= 64) {
    float2 h[KC / 2];
#pragma unroll
    for (int k2 = 0; k2 < KC / 2; ++k2) h[k2] = make_float2(in.get(2 * k2), in.get(2 * k2 + 1));
#pragma unroll 1
    for (int j0 = 0; j0 < NOUT; j0 += JB) {
      float2 accA[JB], accB[JB];
#pragma unroll
      for (int jj = 0; jj < JB; ++jj) { accA[jj] = make_float2(0.f, 0.f); accB[jj] = make_float2(0.f, 0.f); }
      dot_chunk<KC, JB, WG>(h, W + static_cast<size_t>(j0) * K, K, accA, accB);
      float z[JB];
#pragma unroll
      for (int jj = 0; jj < JB; ++jj)
        z[jj] = ((accA[jj].x + accA[jj].y) + (accB[jj].x + accB[jj].y)) + ldw1<WG>(bias + j0 + jj);
      epi(j0, z);
    }
  } else {
#pragma unroll 1
    for (int j0 = 0; j0 < NOUT; j0 += JB) {
      float2 accA[JB], accB[JB];
#pragma unroll
      for (int jj = 0; jj < JB; ++jj) { accA[jj] = make_float2(0.f, 0.f); accB[jj] = make_float2(0.f, 0.f); }
#pragma unroll 1
      for (int kc = 0; kc < K; kc += KC) {
        float2 h[KC / 2];
#pragma unroll
        for (int k2 = 0; k2 < KC / 2; ++k2) h[k2] = make_float2(in.get(kc + 2 * k2), in.get(kc + 2 * k2 + 1));
        dot_chunk<KC, JB, WG>(h, W + static_cast<size_t>(j0) * K + kc, K, accA, accB);
      }
      float z[JB];
#pragma unroll
      for (int jj = 0; jj < JB; ++jj)
        z[jj] = ((accA[jj].x + accA[jj].y) + (accB[jj].x + accB[jj].y)) + ldw1<WG>(bias + j0 + jj);
      epi(j0, z);
    }
  }
}

// Single-output dot (predict head, last variance layer): 4 independent partial sums.
template <int K, bool WG>
PINN_HD float tps_dot1(Col in, const float* w, const float* bias) {
  float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll 4
  for (int k4 = 0; k4 < K / 4; ++k4) {
    float4 ww = ldw4<WG>(w + 4 * k4);
    a0 = ffma2(make_float2(ww.x, ww.y), make_float2(in.get(4 * k4), in.get(4 * k4 + 1)), a0);
    a1 = ffma2(make_float2(ww.z, ww.w), make_float2(in.get(4 * k4 + 2), in.get(4 * k4 + 3)), a1);
  }
  return ((a0.x + a0.y) + (a1.x + a1.y)) + ldw1<WG>(bias);
}

// Hidden-layer epilogue: a = tanh(z); out[j] = a * keep_mask(layer, j).
template <int JB>
struct TanhDropStore {
  Col out;
  const DropCtx* dc;
  uint32_t layer, unit_base;
  PINN_HD void operator()(int j0, float (&z)[JB]) const {
    static_assert(JB % 8 == 0, "mask groups of 8");
#pragma unroll
    for (int g = 0; g < JB; g += 8) {
      float m[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
      if (dc->active) drop8(*dc, layer, j0 + g, unit_base, m);
#pragma unroll
      for (int q = 0; q < 8; ++q) out.set(j0 + g + q, tanh_act(z[g + q]) * m[q]);
    }
  }
};
template <int JB>
struct TanhStore {  // no dropout (layer-0 activation kept pass-invariant; var_layers.4)
  Col out;
  PINN_HD void operator()(int j0, float (&z)[JB]) const {
#pragma unroll
    for (int q = 0; q < JB; ++q) out.set(j0 + q, tanh_act(z[q]));
  }
};

// Layer 0: x[8] (registers) -> tanh(W0 x + b0) into `out`, optionally masked.
template <int H, bool WG, class Epi>
PINN_HD void layer0(const float (&x)[PINN_N_IN], const float* W0, const float* b0, Epi epi) {
  float xs[PINN_N_IN];
#pragma unroll
  for (int i = 0; i < PINN_N_IN; ++i) xs[i] = x[i];
  Col in{xs, 1};
  tps_dense<PINN_N_IN, H, 8, WG>(in, W0, b0, epi);
}

// Everything after the (masked) layer-0 activation sitting in `cur`:
// hidden layers 1..L-1, predict head, variance head.  SMALL nets run in place
// (nxt.p == cur.p); LARGE nets ping-pong between cur and nxt.
// Dropout layer ids: trunk layer l -> l, variance head -> L.
template <int H, bool WG>
PINN_HD void forward_tail(const Weights<WG>& w, int L, Col cur, Col nxt, const DropCtx& dc,
                          float& u, float& v_raw) {
  constexpr int JB = 8;
  for (int l = 1; l < L; ++l) {
    TanhDropStore<JB> epi{nxt, &dc, static_cast<uint32_t>(l), static_cast<uint32_t>(l * H)};
    tps_dense<H, H, JB, WG>(cur, w.W(l), w.b(l), epi);
    Col t = cur; cur = nxt; nxt = t;
  }
  u = tps_dot1<H, WG>(cur, w.Wp(), w.bp());
  {
    TanhDropStore<JB> epi{nxt, &dc, static_cast<uint32_t>(L), static_cast<uint32_t>(L * H)};
    tps_dense<H, H / 2, JB, WG>(cur, w.Wv0(), w.bv0(), epi);
  }
  {
    TanhStore<JB> epi{cur};
    tps_dense<H / 2, H / 4, JB, WG>(nxt, w.Wv1(), w.bv1(), epi);
  }
  v_raw = tps_dot1<H / 4, WG>(cur, w.Wv2(), w.bv2());
}

}  // namespace pinn
