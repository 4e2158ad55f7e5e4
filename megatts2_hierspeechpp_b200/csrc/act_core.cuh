// Core of the fused anti-aliased activation (UpSample1d x2 -> SnakeBeta -> DownSample1d x2), shared by the
// stand-alone kernel (act1d.cu) and the activation-producing variant of the tcgen05 conv (conv_umma.cu).
// Closed form and mapping: see act1d.cu.
#pragma once
#include "hsv_common.cuh"

namespace hsv_act {

constexpr int ROWS = 8;
constexpr int RUNS = 16;
constexpr int NT = ROWS * RUNS;  // 128 threads

template <int R>
struct Cfg {
  static constexpr int TILE = RUNS * R;
  static constexpr int XOFF = 8;             // staged halo (>= 5), multiple of 4 for 16-byte TMA alignment
  static constexpr int XW = TILE + 2 * XOFF;
  // pitch == 4 (mod 32) words -> (row*PITCH + R*run + j) hits 32 distinct banks per warp
  static constexpr int PITCH = ((XW + 27) / 32) * 32 + 4;
  static constexpr int OPITCH = ((TILE + 27) / 32) * 32 + 4;
};

__device__ __forceinline__ float snake(float y, float a, float ib) {
  // activations.py:119  x + 1/(beta+eps) * sin(x*alpha)^2 ; MUFU sine of the fp32 product
  const float s = __sinf(y * a);
  return fmaf(ib * s, s, y);
}

struct ZPair {
  float o, e;  // z[2m-1], z[2m]
};

__device__ __forceinline__ ZPair up_snake(const float w0, const float w1, const float w2, const float w3,
                                          const float w4, const float w5, float a, float ib) {
  // taps pre-doubled (ratio*conv_transpose, resample.py:29; x2 is exact)
  constexpr float G0 = 2.f * HSV_F0, G1 = 2.f * HSV_F1, G2 = 2.f * HSV_F2, G3 = 2.f * HSV_F3,
                  G4 = 2.f * HSV_F4, G5 = 2.f * HSV_F5;
  float yo = G0 * w5;
  float ye = G0 * w0;
  yo = fmaf(G1, w0, yo);
  ye = fmaf(G1, w5, ye);
  yo = fmaf(G2, w4, yo);
  ye = fmaf(G2, w1, ye);
  yo = fmaf(G3, w1, yo);
  ye = fmaf(G3, w4, ye);
  yo = fmaf(G4, w3, yo);
  ye = fmaf(G4, w2, ye);
  yo = fmaf(G5, w2, yo);
  ye = fmaf(G5, w3, ye);
  ZPair z;
  z.o = snake(yo, a, ib);
  z.e = snake(ye, a, ib);
  return z;
}

// ---- packed pairs: Blackwell's FFMA2/FMUL2 (fma.rn.f32x2) process the (odd, even) 2x samples of
// one step in one instruction; a scalar operand broadcasts for free, tap pairs live in uniform registers.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(u64 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

__device__ __forceinline__ u64 up_snake2(const float w0, const float w1, const float w2, const float w3,
                                         const float w4, const float w5, u64 a2, u64 ib2) {
  constexpr float G0 = 2.f * HSV_F0, G1 = 2.f * HSV_F1, G2 = 2.f * HSV_F2, G3 = 2.f * HSV_F3,
                  G4 = 2.f * HSV_F4, G5 = 2.f * HSV_F5;
  // lane lo = odd sample z[2m-1], lane hi = even sample z[2m]
  u64 y = mul2(pk(w0, w0), pk(G1, G0));
  y = fma2(pk(w5, w5), pk(G0, G1), y);
  y = fma2(pk(w1, w1), pk(G3, G2), y);
  y = fma2(pk(w4, w4), pk(G2, G3), y);
  y = fma2(pk(w2, w2), pk(G5, G4), y);
  y = fma2(pk(w3, w3), pk(G4, G5), y);
  float to, te;
  upk(mul2(y, a2), to, te);
  const u64 s = pk(__sinf(to), __sinf(te));
  return fma2(mul2(s, ib2), s, y);
}

__device__ __forceinline__ float down6_2(u64 p0, u64 p1, u64 p2, u64 p3, u64 p4, u64 p5) {
  u64 acc = mul2(p0, pk(HSV_F0, HSV_F1));
  acc = fma2(p5, pk(HSV_F1, HSV_F0), acc);
  acc = fma2(p1, pk(HSV_F2, HSV_F3), acc);
  acc = fma2(p4, pk(HSV_F3, HSV_F2), acc);
  acc = fma2(p2, pk(HSV_F4, HSV_F5), acc);
  acc = fma2(p3, pk(HSV_F5, HSV_F4), acc);
  float lo, hi;
  upk(acc, lo, hi);
  return lo + hi;
}

template <int R, bool EDGE>
__device__ __forceinline__ void walk2(const float *__restrict__ xw, float *__restrict__ outv, float a, float ib,
                                      int64_t ta, int64_t L, float zL, float zR, float sc) {
  u64 ring[6];
  const u64 a2 = pk(a, a), ib2 = pk(ib, ib);
  float w0 = xw[0] * sc, w1 = xw[1] * sc, w2 = xw[2] * sc, w3 = xw[3] * sc, w4 = xw[4] * sc;
  const int64_t n_last = 2 * L - 1;
#pragma unroll
  for (int s = 0; s < R + 5; ++s) {
    const float w5 = xw[s + 5] * sc;
    u64 z = up_snake2(w0, w1, w2, w3, w4, w5, a2, ib2);
    if (EDGE) {
      const int64_t m = ta - 2 + s;
      const int64_t no = 2 * m - 1, ne = 2 * m;
      float zo, ze;
      upk(z, zo, ze);
      zo = no < 0 ? zL : (no > n_last ? zR : zo);
      ze = ne < 0 ? zL : (ne > n_last ? zR : ze);
      z = pk(zo, ze);
    }
    ring[s % 6] = z;
    if (s >= 5) {
      outv[s - 5] = down6_2(ring[(s - 5) % 6], ring[(s - 4) % 6], ring[(s - 3) % 6], ring[(s - 2) % 6],
                            ring[(s - 1) % 6], ring[s % 6]);
    }
    w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5;
  }
}

// One run: outv[0..R-1] = Activation1d(x * sc)[ta .. ta+R-1] for one (b, channel) row.
//   xw[0 .. R+9] = x[ta-5 .. ta+R+4] (replicate-clamped to [0, L-1]); xr = the row in global memory (used only
//   by runs that touch either end of the sequence); al, be = log-scale alpha / beta of the channel.
// Edge semantics: the *activated* 2x signal is replicate-clamped on the 2x grid -- runs whose
// window touches an end substitute z[0] / z[2L-1]; interior runs take the branch-free path.
template <int R>
__device__ __forceinline__ void act_run(const float *__restrict__ xw, float *__restrict__ outv, float al, float be,
                                        int64_t ta, int64_t L, const float *__restrict__ xr, float sc) {
  const float a = expf(al);
  const float ib = 1.0f / (expf(be) + 0.000000001f);
  const bool edge = (2 * ta - 5 < 0) || (2 * (ta + R - 1) + 6 > 2 * L - 1);
  if (!edge) {
    walk2<R, false>(xw, outv, a, ib, ta, L, 0.f, 0.f, sc);
  } else {
    // z[0] (m=0, even) and z[2L-1] (m=L, odd) from clamped global x
    float wl[6], wr[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      int64_t tl = -3 + q, tr = L - 3 + q;
      tl = tl < 0 ? 0 : (tl > L - 1 ? L - 1 : tl);
      tr = tr < 0 ? 0 : (tr > L - 1 ? L - 1 : tr);
      wl[q] = __ldg(xr + tl) * sc;
      wr[q] = __ldg(xr + tr) * sc;
    }
    const float zL = up_snake(wl[0], wl[1], wl[2], wl[3], wl[4], wl[5], a, ib).e;
    const float zR = up_snake(wr[0], wr[1], wr[2], wr[3], wr[4], wr[5], a, ib).o;
    walk2<R, true>(xw, outv, a, ib, ta, L, zL, zR, sc);
  }
}

}  // namespace hsv_act
