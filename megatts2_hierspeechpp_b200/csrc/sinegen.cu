// f0-driven harmonic sine source with a drift-free phase accumulator (BASELINE.json north_star item 3).
//
// The reference repository has no SineGen on this path (its HierSpeech++ generator takes the SourceNetwork's hidden
// state): this is the NSF / HiFTNet-style source the north_star names, offered as a utility for
// vocoder variants that want it.  phase[n] = sum_{k<=n} f0_up[k] / sr (in turns), out[h][n] = amp * sin(2 pi (h+1)
// phase[n]) * voiced[n].
//
// A float32 running sum of ~5e5 increments of ~1e-2 turns loses ~1e-2 rad over 30 s.  Here the phase is a 64-bit
// FIXED-POINT number of turns (2^64 = one cycle): every increment is rounded once to 2^-64 turns (from an fp64
// quotient), the running sum is integer arithmetic -- exact, associative (so the parallel scan is deterministic and
// equal to the sequential sum bit for bit) and its wrap-around IS the phase wrap.  Harmonic h is (h+1) * phase mod
// 2^64, again exact.  Only the final conversion to an fp32 angle rounds (2^-24 turns = 3.7e-7 rad).
#include "hsv_common.cuh"

namespace {

constexpr int SG_THREADS = 256, SG_PER = 16, SG_BLOCK = SG_THREADS * SG_PER;  // 4096 samples per CTA

__device__ __forceinline__ unsigned long long inc_fixed(float f0, double inv_sr) {
  // turns per sample as a 0.64 fixed-point number; unvoiced (f0 <= 0) frames do not advance the phase
  if (!(f0 > 0.f)) return 0ull;
  const double turns = (double)f0 * inv_sr;                 // < 0.5 for f0 below Nyquist
  return (unsigned long long)__double2ull_rn(turns * 18446744073709551616.0);
}

// pass 1: per-CTA totals of the increments
__global__ void sg_block_sums(const float *__restrict__ f0, unsigned long long *__restrict__ sums, int64_t T, int hop,
                              int64_t L, double inv_sr, int nblk) {
  const int row = blockIdx.y, blk = blockIdx.x;
  const float *fr = f0 + (int64_t)row * T;
  const int64_t n0 = (int64_t)blk * SG_BLOCK + (int64_t)threadIdx.x * SG_PER;
  unsigned long long s = 0ull;
#pragma unroll
  for (int i = 0; i < SG_PER; ++i) {
    const int64_t n = n0 + i;
    if (n < L) s += inc_fixed(__ldg(fr + n / hop), inv_sr);
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ unsigned long long w[SG_THREADS / 32];
  if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0ull;
    for (int i = 0; i < SG_THREADS / 32; ++i) t += w[i];
    sums[(int64_t)row * nblk + blk] = t;
  }
}

// pass 2: exclusive scan of the CTA totals of each row (a few hundred values: one thread per row)
__global__ void sg_scan_rows(unsigned long long *__restrict__ sums, int rows, int nblk) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  unsigned long long acc = 0ull;
  for (int i = 0; i < nblk; ++i) {
    const unsigned long long v = sums[(int64_t)row * nblk + i];
    sums[(int64_t)row * nblk + i] = acc;
    acc += v;
  }
}

// pass 3: inclusive scan inside the CTA (+ the CTA's offset), harmonics, sines
__global__ void sg_emit(const float *__restrict__ f0, const unsigned long long *__restrict__ offs,
                        float *__restrict__ out, float *__restrict__ uv, int64_t T, int hop, int64_t L, double inv_sr,
                        int nblk, int H, float amp) {
  const int row = blockIdx.y, blk = blockIdx.x;
  const float *fr = f0 + (int64_t)row * T;
  const int64_t n0 = (int64_t)blk * SG_BLOCK + (int64_t)threadIdx.x * SG_PER;
  unsigned long long ph[SG_PER];
  float voiced[SG_PER];
  unsigned long long s = 0ull;
#pragma unroll
  for (int i = 0; i < SG_PER; ++i) {
    const int64_t n = n0 + i;
    const float f = n < L ? __ldg(fr + n / hop) : 0.f;
    voiced[i] = f > 0.f ? 1.f : 0.f;
    s += inc_fixed(f, inv_sr);
    ph[i] = s;                                   // inclusive within the thread
  }
  // exclusive scan of the per-thread totals across the CTA
  unsigned long long incl = s;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __shared__ unsigned long long wsum[SG_THREADS / 32];
  if (lane == 31) wsum[wp] = incl;
  __syncthreads();
  unsigned long long base = offs[(int64_t)row * nblk + blk] + (incl - s);
  for (int i = 0; i < wp; ++i) base += wsum[i];
#pragma unroll
  for (int i = 0; i < SG_PER; ++i) {
    const int64_t n = n0 + i;
    if (n >= L) break;
    const unsigned long long p = base + ph[i];
    if (uv) uv[(int64_t)row * L + n] = voiced[i];
    for (int h = 0; h < H; ++h) {
      const unsigned long long q = p * (unsigned long long)(h + 1);           // harmonic phase, wraps exactly
      const float turns = (float)(unsigned int)(q >> 40) * (1.0f / 16777216.0f);   // top 24 bits -> [0, 1)
      out[((int64_t)row * H + h) * L + n] = amp * voiced[i] * sinpif(2.0f * turns);
    }
  }
}

}  // namespace

extern "C" int64_t hsv_sinegen_workspace(int B, int64_t T, int hop) {
  const int64_t L = T * hop;
  return (int64_t)B * ((L + SG_BLOCK - 1) / SG_BLOCK) * 8;
}

extern "C" int hsv_sinegen(const float *f0, float *out, float *uv, void *workspace, int B, int64_t T, int hop,
                           float sample_rate, int harmonics, float amp, void *stream) {
  if (B == 0 || T == 0) return HSV_OK;
  HSV_REQUIRE(f0 && out && workspace, "sinegen: null pointer");
  HSV_REQUIRE(B > 0 && B <= 65535 && T > 0 && hop >= 1 && harmonics >= 1 && sample_rate > 0.f,
              "sinegen: bad arguments B=%d T=%lld hop=%d H=%d sr=%g", B, (long long)T, hop, harmonics, sample_rate);
  const int64_t L = T * hop;
  const int64_t nb = (L + SG_BLOCK - 1) / SG_BLOCK;
  HSV_REQUIRE(nb < (1ll << 31), "sinegen: sequence too long");
  cudaStream_t st = hsv::as_stream(stream);
  unsigned long long *ws = reinterpret_cast<unsigned long long *>(workspace);
  const double inv_sr = 1.0 / (double)sample_rate;
  sg_block_sums<<<dim3((unsigned)nb, B), SG_THREADS, 0, st>>>(f0, ws, T, hop, L, inv_sr, (int)nb);
  sg_scan_rows<<<(B + 63) / 64, 64, 0, st>>>(ws, B, (int)nb);
  sg_emit<<<dim3((unsigned)nb, B), SG_THREADS, 0, st>>>(f0, ws, out, uv, T, hop, L, inv_sr, (int)nb, harmonics, amp);
  return hsv::check_launch("sinegen");
}
