// Fused anti-aliased activation: UpSample1d(x2 kaiser-sinc) -> SnakeBeta -> DownSample1d(x2).
//
// Replaces alias_free_torch/act.py:23-27 (9 ATen kernels and three 2x-rate
// temporaries per call in the reference) with ONE kernel that reads x once and
// writes the result once.  Closed form, f = 12 symmetric taps:
//   y[2m-1] = 2*sum_i f[2i]  *x[clamp(m+2-i)]      (odd 2x sample)
//   y[2m]   = 2*sum_i f[2i+1]*x[clamp(m+2-i)]      (even 2x sample)
//   z[n]    = y[n] + 1/(exp(beta)+1e-9) * sin(y[n]*exp(alpha))^2
//   out[t]  = sum_j f[j]*z[clamp2L(2t+j-5)]
// x is edge-replicated on the 1x grid, z on the 2x grid (the *activated* edge
// sample is replicated, resample.py:28 / filter.py:90).
//
// Mapping: a CTA owns ROWS=8 channel rows x TILE=RUNS*R time steps.  The x tile
// (+-5 halo, clamped) is staged in shared memory with coalesced loads; each
// thread then walks a run of R consecutive outputs with the 2x signal living
// only in registers (a ring of six (odd,even) z pairs), so no 2x-rate value
// ever touches shared or global memory.  R is odd and the row pitch is
// 4 (mod 32) words, which makes every shared access of the walk conflict-free.
// Results are staged in shared memory and written back coalesced, either as
// fp32 [B,C,L] or as the fp16 "blk16" tensor-core operand layout.
#include "hsv_common.cuh"

namespace {

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

__device__ __forceinline__ float down6(const ZPair &p0, const ZPair &p1, const ZPair &p2, const ZPair &p3,
                                       const ZPair &p4, const ZPair &p5) {
  // out[t] = f0 zo(t-2) + f1 ze(t-2) + f2 zo(t-1) + f3 ze(t-1) + f4 zo(t) + f5 ze(t)
  //        + f5 zo(t+1) + f4 ze(t+1) + f3 zo(t+2) + f2 ze(t+2) + f1 zo(t+3) + f0 ze(t+3)
  float s0 = HSV_F0 * (p0.o + p5.e);
  float s1 = HSV_F1 * (p0.e + p5.o);
  s0 = fmaf(HSV_F2, p1.o + p4.e, s0);
  s1 = fmaf(HSV_F3, p1.e + p4.o, s1);
  s0 = fmaf(HSV_F4, p2.o + p3.e, s0);
  s1 = fmaf(HSV_F5, p2.e + p3.o, s1);
  return s0 + s1;
}

template <int R, bool EDGE>
__device__ __forceinline__ void walk(const float *__restrict__ xw, float *__restrict__ outv, float a, float ib,
                                     int64_t ta, int64_t L, float zL, float zR, float sc) {
  // xw[0 .. R+9] = x[ta-5 .. ta+R+4] (already clamped);  outv[0..R-1] = out[ta .. ta+R-1]
  ZPair ring[6];
  float w0 = xw[0] * sc, w1 = xw[1] * sc, w2 = xw[2] * sc, w3 = xw[3] * sc, w4 = xw[4] * sc;
  const int64_t n_last = 2 * L - 1;
#pragma unroll
  for (int s = 0; s < R + 5; ++s) {
    const float w5 = xw[s + 5] * sc;
    ZPair z = up_snake(w0, w1, w2, w3, w4, w5, a, ib);
    if (EDGE) {
      const int64_t m = ta - 2 + s;
      const int64_t no = 2 * m - 1, ne = 2 * m;
      z.o = no < 0 ? zL : (no > n_last ? zR : z.o);
      z.e = ne < 0 ? zL : (ne > n_last ? zR : z.e);
    }
    ring[s % 6] = z;
    if (s >= 5) {
      outv[s - 5] = down6(ring[(s - 5) % 6], ring[(s - 4) % 6], ring[(s - 3) % 6], ring[(s - 2) % 6],
                          ring[(s - 1) % 6], ring[s % 6]);
    }
    w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5;
  }
}

// ---- packed-pair variant: Blackwell's FFMA2/FMUL2 (fma.rn.f32x2) process the (odd, even) 2x samples of
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

template <int R, int OUT_MODE, bool PACKED>
__global__ void __launch_bounds__(NT) act1d_kernel(const float *__restrict__ x, void *__restrict__ outp,
                                                   const float *__restrict__ alpha,
                                                   const float *__restrict__ beta, int C, int64_t L,
                                                   int64_t nrows, int ntiles, int64_t Lp, float sc, int cw) {
  using K = Cfg<R>;
  // x_s[c][p] holds x[row0+c][t0 - XOFF + p]; XOFF = 8 keeps the tile start 16-byte aligned for TMA
  __shared__ __align__(16) float x_s[ROWS * K::PITCH];
  __shared__ __align__(16) float o_s[OUT_MODE == 0 ? ROWS * K::OPITCH : K::TILE * 4];
  __shared__ __align__(8) uint64_t bar;

  const int tid = threadIdx.x;
  int tile, ch;        // time tile; channel of this thread's row
  int64_t row0;        // first of the CTA's 8 (b, channel) rows
  const int c = tid % ROWS, run = tid / ROWS;
  if (OUT_MODE == 1) {
    // swizzled operand output: grid = (tiles x units-per-chunk, chunks, batch), no divisions.  The cw/8 CTAs that
    // fill the 16-byte units of the same operand rows are adjacent in launch order, so their partial-sector
    // writes meet in L2.
    const int lgu = cw == 64 ? 3 : (cw == 32 ? 2 : 1);  // log2(units per chunk)
    tile = (int)(blockIdx.x >> lgu);
    const int c0 = (int)blockIdx.y * cw + (int)(blockIdx.x & ((1u << lgu) - 1u)) * 8;
    row0 = (int64_t)blockIdx.z * C + c0;
    ch = c0 + c;
  } else {
    tile = blockIdx.x % ntiles;
    row0 = (int64_t)(blockIdx.x / ntiles) * ROWS;
    ch = (int)((row0 + c) % C);
  }
  const int64_t t0 = (int64_t)tile * K::TILE;
  const int64_t row = row0 + c;
  // PDL: let the next kernel start its prologue now; parameters are static, so load them before waiting
  // for the producer of x
  hsv::pdl_launch_dependents();
  float al = 0.f, be = 0.f;
  if (row < nrows) {
    al = __ldg(alpha + ch);
    be = __ldg(beta + ch);
  }
  hsv::pdl_wait();

  // ---- stage the x tile ----
  // interior tiles of 16-byte-aligned rows: one 1-D bulk TMA copy per row (no issue slots, no registers);
  // edge tiles / unaligned rows: plain loads with the replicate clamp of the 1x grid.
  const bool fast = ((L & 3) == 0) && t0 >= K::XOFF && t0 + K::TILE + K::XOFF <= L && row0 + ROWS <= nrows &&
                    ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (fast) {
    const uint32_t bar_a = static_cast<uint32_t>(__cvta_generic_to_shared(&bar));
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    constexpr uint32_t ROW_BYTES = (K::TILE + 2 * K::XOFF) * 4;
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(ROW_BYTES * ROWS)
                   : "memory");
    if (tid < ROWS) {
      const float *src = x + (row0 + tid) * L + (t0 - K::XOFF);
      const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(x_s + tid * K::PITCH));
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "l"(src), "r"(ROW_BYTES), "r"(bar_a)
                   : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t"
          "}"
          : "=r"(ok)
          : "r"(bar_a)
          : "memory");
    }
  } else {
    // all loads of a thread are issued before the first store (one exposed memory latency, not one per row)
    constexpr int NLD = (ROWS * K::XW + NT - 1) / NT;
    const int Lm1 = (int)(L - 1);
    float v[NLD];
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int idx = tid + NT * i;
      const int cc = idx / K::XW, p = idx - cc * K::XW;
      const int64_t rw = row0 + cc;
      const int64_t t = t0 - K::XOFF + p;
      const int tc = t < 0 ? 0 : (t > Lm1 ? Lm1 : (int)t);
      v[i] = (idx < ROWS * K::XW && rw < nrows) ? __ldg(x + rw * L + tc) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int idx = tid + NT * i;
      const int cc = idx / K::XW, p = idx - cc * K::XW;
      if (idx < ROWS * K::XW) x_s[cc * K::PITCH + p] = v[i];
    }
    __syncthreads();
  }

  const int64_t ta = t0 + (int64_t)run * R;
  float outv[R];
  const bool active = row < nrows && ta < L;
  if (active) {
    const float a = expf(al);
    const float ib = 1.0f / (expf(be) + 0.000000001f);
    const float *xw = x_s + c * K::PITCH + run * R + (K::XOFF - 5);
    const bool edge = (2 * ta - 5 < 0) || (2 * (ta + R - 1) + 6 > 2 * L - 1);
    if (!edge) {
      if (PACKED) walk2<R, false>(xw, outv, a, ib, ta, L, 0.f, 0.f, sc);
      else walk<R, false>(xw, outv, a, ib, ta, L, 0.f, 0.f, sc);
    } else {
      // z[0] (m=0, even) and z[2L-1] (m=L, odd) from clamped global x
      const float *xr = x + row * L;
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
      if (PACKED) walk2<R, true>(xw, outv, a, ib, ta, L, zL, zR, sc);
      else walk<R, true>(xw, outv, a, ib, ta, L, zL, zR, sc);
    }
  }

  // ---- stage results, write back coalesced ----
  if (OUT_MODE == 0) {
    if (active) {
      float *o = o_s + c * K::OPITCH + run * R;
#pragma unroll
      for (int j = 0; j < R; ++j) o[j] = outv[j];
    }
    __syncthreads();
    float *out = reinterpret_cast<float *>(outp);
    for (int idx = tid; idx < ROWS * K::TILE; idx += NT) {
      const int cc = idx / K::TILE, p = idx - cc * K::TILE;
      const int64_t r = row0 + cc, t = t0 + p;
      if (r < nrows && t < L) out[r * L + t] = o_s[cc * K::OPITCH + p];
    }
  } else {
    __half *o_h = reinterpret_cast<__half *>(o_s);  // [TILE][8]
    if (active) {
#pragma unroll
      for (int j = 0; j < R; ++j) o_h[(run * R + j) * 8 + c] = __float2half_rn(outv[j]);
    }
    __syncthreads();
    // rows row0..row0+7 are one 8-channel unit of one batch item (C % 8 == 0)
    const uint4 *src = reinterpret_cast<const uint4 *>(o_s);
    // the swizzled offset is shifts and xors (cw is a power of two)
    const int lg = cw == 64 ? 7 : (cw == 32 ? 6 : 5);              // log2(row bytes)
    const uint32_t mask = (uint32_t)(cw >> 3) - 1u;
    const uint32_t ub = (blockIdx.x & mask) << 4;                  // byte offset of this CTA's unit in a row
    uint8_t *base = reinterpret_cast<uint8_t *>(outp) + (((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * Lp << lg);
    const uint64_t r0 = (uint64_t)(HSV_BLK_PAD + t0);
    for (int p = tid; p < K::TILE; p += NT) {
      if (t0 + p < L) {
        const uint64_t lin = ((r0 + (uint64_t)p) << lg) + ub;
        *reinterpret_cast<uint4 *>(base + (lin ^ (((lin >> 7) & mask) << 4))) = src[p];
      }
    }
  }
}

int g_act_variant = 1;  // 1: packed f32x2 math (default), 0: scalar math
int g_act_run = 0;      // bring-up aid: force the run length (17 or 33); 0 = automatic

template <int R, int OUT_MODE>
int launch(const float *x, void *out, const float *alpha, const float *beta, int B, int C, int64_t L, float sc,
           cudaStream_t st) {
  using K = Cfg<R>;
  const int64_t nrows = (int64_t)B * C;
  const int64_t ntiles = (L + K::TILE - 1) / K::TILE;
  const int64_t ngrp = (nrows + ROWS - 1) / ROWS;
  const int64_t nblk = ntiles * ngrp;
  HSV_REQUIRE(nblk < (1ll << 31) && ntiles < (1ll << 31), "act1d: grid too large");
  cudaLaunchConfig_t cfg = {};
  if (OUT_MODE == 1) {
    const int cwl = hsv::blk_cw(C);
    HSV_REQUIRE(ntiles * (cwl / 8) < (1ll << 31) && C / cwl <= 65535 && B <= 65535, "act1d: grid too large");
    cfg.gridDim = dim3((unsigned)(ntiles * (cwl / 8)), (unsigned)(C / cwl), (unsigned)B);
  } else {
    cfg.gridDim = dim3((unsigned)nblk);
  }
  cfg.blockDim = dim3(NT);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = hsv::g_pdl ? 1 : 0;
  const int nt_i = (int)ntiles;
  const int64_t Lp = hsv::blk16_rows(L);
  const int cw = OUT_MODE == 1 ? hsv::blk_cw(C) : 0;
  cudaError_t e;
  if (g_act_variant)
    e = cudaLaunchKernelEx(&cfg, act1d_kernel<R, OUT_MODE, true>, x, out, alpha, beta, C, L, nrows, nt_i, Lp, sc, cw);
  else
    e = cudaLaunchKernelEx(&cfg, act1d_kernel<R, OUT_MODE, false>, x, out, alpha, beta, C, L, nrows, nt_i, Lp, sc, cw);
  if (e != cudaSuccess) {
    cudaGetLastError();
    hsv::set_error("act1d_snakebeta: launch failed: %s", cudaGetErrorString(e));
    return HSV_ERR_CUDA;
  }
  return hsv::check_launch("act1d_snakebeta");
}

}  // namespace

// bring-up aid (A/B of the packed-math variant); not part of the drop-in contract
extern "C" int hsv_set_act_variant(int v) {
  g_act_variant = v & 1;
  g_act_run = v >> 8;  // bits 8..: forced run length
  return HSV_OK;
}

extern "C" int hsv_act1d_snakebeta(const float *x, void *out, const float *alpha, const float *beta, int B,
                                   int C, int64_t L, int out_mode, float in_scale, void *stream) {
  if (B == 0 || L == 0) return HSV_OK;  // empty batch / sequence
  HSV_REQUIRE(x && out && alpha && beta, "act1d: null pointer");
  HSV_REQUIRE(B >= 0 && C > 0 && L >= 0, "act1d: bad shape B=%d C=%d L=%lld", B, C, (long long)L);
  HSV_REQUIRE(out_mode == 0 || out_mode == 1, "act1d: out_mode must be 0 (fp32 NCL) or 1 (fp16 blk16)");
  if (B == 0 || L == 0) return HSV_OK;
  cudaStream_t st = hsv::as_stream(stream);
  // run length per thread: 33 outputs (10 halo steps per 33: 1.15x redundant 2x-rate work) when the tensor is
  // large enough to still fill the GPU with the bigger tiles, else 17 (1.29x) for more CTAs
  const int64_t groups = ((int64_t)B * C + ROWS - 1) / ROWS;
  const bool big = g_act_run ? g_act_run == 33 : (groups * ((L + 33 * RUNS - 1) / (33 * RUNS)) >= 16 * 148);
  if (out_mode == 0)
    return big ? launch<33, 0>(x, out, alpha, beta, B, C, L, in_scale, st)
               : launch<17, 0>(x, out, alpha, beta, B, C, L, in_scale, st);
  HSV_REQUIRE(C % 16 == 0, "act1d: blk16 output needs C %% 16 == 0 (C=%d)", C);
  return big ? launch<33, 1>(x, out, alpha, beta, B, C, L, in_scale, st)
             : launch<17, 1>(x, out, alpha, beta, B, C, L, in_scale, st);
}
