// LEGACY (layout 0, SWIZZLE_NONE operands) variant of the tcgen05 conv, kept only as the bring-up reference for
// conv_umma.cu (swizzled operands); selected with hsv_set_layout(0).
// Dense Conv1d / ConvTranspose1d as a tcgen05 / TMEM implicit GEMM (sm_100a).
//
// Replaces the weight-normed convolutions of the reference's waveform path
//   * AMPBlock convs    hierspeechpp_speechsynthesizer.py:349-364,380-384; speechsr24k/speechsr.py:21-36,52-56
//   * conv_pre / proj / DBlock convs   hierspeechpp_speechsynthesizer.py:401,426,321-325
//   * ups[i] ConvTranspose1d           hierspeechpp_speechsynthesizer.py:404-408,434
// with one kernel.  Both are "sum over taps of a row-shifted [rows x Cin] x [Cin x Cout] product":
//   conv1d      out[t]       = b + sum_j  W[:, :, j]        a[t + (j-(k-1)/2) d]
//   convT phase out[u q + r] = b + sum_i  W[:, :, r' + i u] a[q + c - i]      (r' = (r+p) mod u, c = (r+p) div u)
// GEMM view per CTA:  D[M = 128 rows, N = n_tile out channels] += A_tap[128 x 16] * W_tap[16 x n_tile]
// over all taps and 16-channel K-steps; fp16 operands, fp32 accumulation in TMEM.
//
// Operand staging.  Activations arrive in the "blk16" layout written by the fused activation kernel
// (or hsv_pack_blk16): fp16 [B][Cin/8][Lp][8] -- per 8-channel chunk the time rows are consecutive
// 16-byte records with zero rows around every sequence.  A time tile (+halo) of one chunk is ONE
// contiguous span, fetched with a 1-D bulk TMA copy into shared memory as [chunk][row][8 halves].
// That is tcgen05's K-major SWIZZLE_NONE canonical layout with SBO = 128 B (8 rows x 16 B) and
// LBO = rows*16 B, in which rows of one K-chunk are uniformly 16 B apart -- so the operand of a tap is
// the SAME shared tile with the descriptor start address advanced by the tap's row offset.  The halo is
// loaded once and reused by all taps; zero padding comes from the zero rows of the blk16 layout.
// Weights are pre-packed as [phase][n-tile][K-step][2][n_tile][8] fp16 so a group of K-steps is one
// contiguous span, streamed through a ring of shared stages by bulk TMA copies (mbarrier full/empty).
//
// Roles (128 threads): warp0/lane0 TMA producer, warp1/lane0 MMA issuer, warp2 TMEM alloc/free, then all
// four warps run the epilogue: tcgen05.ld (lane = row), + bias, + residual, store fp32 [B,C,L] (for
// stride-1 outputs a warp stores 32 consecutive time steps of one channel = 128 B coalesced) and
// optionally accumulate the mean over resblocks.  Residual loads are batched per 16-column chunk and
// prefetched one chunk ahead (out may alias residual, so the compiler cannot do this itself).
#include "hsv_common.cuh"

namespace {

constexpr int TILE_M = HSV_UMMA_TILE_M;  // 128
constexpr int MAX_STAGES = 4;
constexpr int MAX_PHASES = 8;
constexpr int MAX_TAPS = 16;

struct TapTable {
  int nphase;
  int out_stride;
  int ntaps[MAX_PHASES];
  int out_off[MAX_PHASES];
  int8_t row_off[MAX_PHASES][MAX_TAPS];  // input row offset of the tap relative to the output row
  int8_t wj[MAX_PHASES][MAX_TAPS];       // which tap of the weight tensor
  int h_lo, h_hi;                        // rows needed before / after the tile
};

struct Params {
  const uint4 *a;   // blk16 activations, 16-byte records
  const uint4 *w;   // packed weights
  const float *bias;
  const float *residual;
  float *out;
  float *acc;
  int acc_mode;
  float acc_div;
  int Cin, Cout;
  int64_t L, Lp, Lout;
  int n_tile, nco_tiles;
  int ntiles;   // real CTA tiles (gridDim.x is rounded up to the cluster size; the extra CTAs only stream weights)
  int msub;     // 128-row sub-tiles per CTA (1, 2 or 4): every weight K-step feeds msub MMAs, which divides the
                // per-SM weight ingest (the B200 L2->SM port delivers ~42 B/clk, less than one N=128 MMA eats)
  int R;        // rows per chunk in the shared A tile = 128 + h_lo + h_hi
  int Rs;       // row spacing of the chunks in shared memory (>= R; LBO_A = Rs*16 B)
  int NB;       // rows per K-chunk of a packed weight K-step (>= n_tile; LBO_B = NB*16 B)
  int G;        // K-steps per weight block
  int stages;
  uint32_t tmem_cols;
  int debug;
  TapTable tt;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // try_wait suspends in hardware for a bounded time; the iteration bound turns a protocol bug into a
  // trap instead of a hang
  for (uint32_t it = 0; !mbar_try(bar, parity); ++it) {
    if (it > (1u << 26)) {
      printf("hsv conv_umma: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar,
                                            uint16_t cta_mask) {
  // multicast: the bytes land at the same shared offset of every CTA in cta_mask and complete_tx is
  // signalled on the mbarrier at the same offset of each of them
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  // one lane of the (converged) warp; lets the surrounding address arithmetic stay warp-uniform so the
  // compiler keeps descriptors in uniform registers instead of moving them per MMA (R2UR)
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
  // base_offset 0, layout SWIZZLE_NONE (0) [61,64)
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                              uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int MSUB, int MINB, bool SMALLN>
__global__ void __launch_bounds__(128, MINB) conv_umma_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 + 2 * MAX_STAGES];  // a_full, acc_full, w_full[S], w_empty[S]
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, b = blockIdx.z;
  const int ph = blockIdx.y / p.nco_tiles, nt = blockIdx.y - ph * p.nco_tiles;
  const int nchunks = p.Cin >> 3;
  const int KC = p.Cin >> 4;
  const int ntaps = p.tt.ntaps[ph];
  const int ksteps = ntaps * KC;
  const int nblocks = (ksteps + p.G - 1) / p.G;
  const uint32_t a_bytes_chunk = (uint32_t)p.R * 16u;     // bytes copied per chunk
  const uint32_t a_pitch = (uint32_t)p.Rs * 16u;          // chunk spacing in shared memory
  const uint32_t a_bytes = a_bytes_chunk * nchunks;
  const uint32_t kstep_bytes = 32u * p.NB;
  const uint32_t wblk_bytes = kstep_bytes * p.G;

  const uint32_t a_s = smem_u32(smem);
  const uint32_t w_s = a_s + ((a_pitch * nchunks + 127u) & ~127u);
  const uint32_t bar_a = smem_u32(&bars[0]), bar_acc = smem_u32(&bars[1]);
  const uint32_t bar_wf = smem_u32(&bars[2]), bar_we = smem_u32(&bars[2 + MAX_STAGES]);

  // Cluster of CL CTAs = CL consecutive M-tiles of the same (phase, n-tile, batch): they need the same
  // weights, so every CTA fetches 1/CL of each weight block and multicasts it to the whole cluster
  // (L2 -> SM weight traffic per SM drops by CL).  A stage is free again when all CL consumers released it.
  hsv::pdl_launch_dependents();  // PDL: the next kernel may begin its prologue
  const uint32_t CL = cluster_nctarank();
  const uint32_t crank = cluster_ctarank();
  const uint16_t cmask = (uint16_t)((1u << CL) - 1u);
  const bool real_tile = tile < p.ntiles;

  if (threadIdx.x == 0) {
    mbar_init(bar_a, 1);
    mbar_init(bar_acc, 1);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_wf + 8 * s, 1);
      mbar_init(bar_we + 8 * s, CL);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_s)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // every CTA's barriers are initialised before any remote arrive / multicast
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (warp == 0 && lane == 0) {
    // ---------------- TMA producer ----------------
    // K-step offset of (phase, n-tile) in the packed weight stream
    int64_t ks0 = 0;
    for (int q = 0; q < ph; ++q) ks0 += (int64_t)p.tt.ntaps[q] * KC * p.nco_tiles;
    ks0 += (int64_t)nt * ksteps;
    const uint4 *wsrc = p.w + ks0 * (kstep_bytes >> 4);
    auto load_w = [&](int blk) {
      const int s = blk % p.stages;
      if (blk >= p.stages) mbar_wait(bar_we + 8 * s, ((blk / p.stages) - 1) & 1);
      const int nk = min(p.G, ksteps - blk * p.G);
      const uint32_t bytes = kstep_bytes * nk;
      mbar_expect_tx(bar_wf + 8 * s, bytes);  // the whole block lands here: own slice + the peers' multicasts
      if (CL == 1) {
        bulk_g2s(w_s + s * wblk_bytes, wsrc + (int64_t)blk * (wblk_bytes >> 4), bytes, bar_wf + 8 * s);
      } else {
        const uint32_t slice = bytes / CL;  // multiple of 16: kstep_bytes >= 2048 whenever CL > 1
        bulk_g2s_mc(w_s + s * wblk_bytes + crank * slice,
                    wsrc + (int64_t)blk * (wblk_bytes >> 4) + ((crank * slice) >> 4), slice, bar_wf + 8 * s, cmask);
      }
    };
    // weights are static: fill the ring before waiting for the kernel that produces the activations
    const int npre = nblocks < p.stages ? nblocks : p.stages;
    for (int blk = 0; blk < npre; ++blk) load_w(blk);
    hsv::pdl_wait();
    if (real_tile) {
      const int64_t row0 = (int64_t)HSV_BLK_PAD + (int64_t)tile * TILE_M * MSUB - p.tt.h_lo;
      mbar_expect_tx(bar_a, a_bytes);
      for (int q = 0; q < nchunks; ++q) {
        const uint4 *src = p.a + ((int64_t)b * nchunks + q) * p.Lp + row0;
        bulk_g2s(a_s + q * a_pitch, src, a_bytes_chunk, bar_a);
      }
    }
    for (int blk = npre; blk < nblocks; ++blk) load_w(blk);
  } else if (warp == 1 && lane == 0) {
    // ---------------- MMA issuer ----------------
    // This loop runs on ONE thread, so every dependent scalar instruction per MMA is exposed latency.
    // Weight blocks are tap-aligned (host picks G = whole taps, or a divisor of the K-steps of one tap),
    // so the inner loop only bumps the two 14-bit address fields of the descriptors.
    // InstrDescriptor: D=F32 (1<<4), A=B=F16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
    const bool swap = p.debug & 1;
    if (real_tile) mbar_wait(bar_a, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t sbo16 = 128u >> 4, a_lbo16 = p.Rs, b_lbo16 = p.NB;  // 16-byte units
    // SmemDescriptor hi word: SBO>>4 [0,14), version=1 at bit 14; lo word: addr>>4 [0,14), LBO>>4 [16,30)
    const uint32_t a_hi = (swap ? a_lbo16 : sbo16) | (1u << 14);
    const uint32_t b_hi = (swap ? b_lbo16 : sbo16) | (1u << 14);
    // (inside a cluster the shared-window address carries the CTA rank in its upper bits: keep the
    //  18-bit CTA-local offset only)
    const uint32_t a_lo0 = ((swap ? sbo16 : a_lbo16) << 16) | ((a_s & 0x3FFFFu) >> 4);   // + row + kc*2R
    const uint32_t b_lo0 = ((swap ? sbo16 : b_lbo16) << 16) | ((w_s & 0x3FFFFu) >> 4);   // + stage*blk + g*kstep
    const uint32_t a_kstep16 = 2u * (uint32_t)p.Rs, b_kstep16 = kstep_bytes >> 4, wblk16 = wblk_bytes >> 4;
    const int G = p.G;
    const bool whole_taps = KC <= G;            // block = m whole taps, else a tap = bpt blocks
    const int m = whole_taps ? G / KC : 1;
    const int bpt = whole_taps ? 1 : KC / G;
    uint32_t acc_flag = 0;
    int stage = 0;
    uint32_t parity = 0;
    for (int blk = 0; blk < nblocks; ++blk) {
      mbar_wait(bar_wf + 8 * stage, parity);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (real_tile) {
        uint32_t b_lo = b_lo0 + (uint32_t)stage * wblk16;
        const int j0 = whole_taps ? blk * m : blk / bpt;
        const int nt_blk = whole_taps ? min(m, ntaps - j0) : 1;
        const int kc0 = whole_taps ? 0 : (blk - j0 * bpt) * G;
        const int nkc = whole_taps ? KC : G;
        for (int tp = 0; tp < nt_blk; ++tp) {
          uint32_t a_lo = a_lo0 + (uint32_t)(p.tt.row_off[ph][j0 + tp] + p.tt.h_lo) + (uint32_t)kc0 * a_kstep16;
#pragma unroll 4
          for (int kc = 0; kc < nkc; ++kc) {
#pragma unroll
            for (int sub = 0; sub < MSUB; ++sub)  // the same weight K-step feeds every 128-row sub-tile
              umma_f16_lohi(tmem + (uint32_t)(sub * p.n_tile), a_lo + (uint32_t)(sub * TILE_M), a_hi, b_lo, b_hi, idesc,
                            acc_flag);
            acc_flag = 1u;
            a_lo += a_kstep16;
            b_lo += b_kstep16;
          }
        }
      }
      if (CL == 1) umma_commit(bar_we + 8 * stage);
      else umma_commit_mc(bar_we + 8 * stage, cmask);  // release this stage in every CTA of the cluster
      if (++stage == p.stages) {
        stage = 0;
        parity ^= 1u;
      }
    }
    umma_commit(bar_acc);
  }

  // ---------------- epilogue: all 4 warps ----------------
  // A compact ROLLED loop over (sub-tile, 16-column chunk) units: the CTA has only four warps, so a fully
  // unrolled epilogue is fetch-bound straight-line code (measured: +7 us per launch at 64 KB of SASS).
  // Residual loads run two units ahead of their use (the first two are issued before the accumulator wait,
  // so they overlap the MMAs); out may alias residual, hence the explicit ordering.
  hsv::pdl_wait();  // residual / out / acc belong to predecessor kernels
  __syncwarp();
  const int co0 = nt * p.n_tile;
  const int64_t cs = p.Lout;  // channel stride
  const int64_t chan_base = ((int64_t)b * p.Cout + co0) * p.Lout + p.tt.out_off[ph];
  const int64_t row_base = (int64_t)tile * MSUB * TILE_M + warp * 32 + lane;  // GEMM row of sub-tile 0
  const int nchk = p.n_tile >> 4;
  const int nunits = MSUB * nchk;
  const bool has_res = p.residual != nullptr;
  auto unit_ptr = [&](int u, int64_t &off, bool &ok) {
    const int sub = u / nchk, c0 = (u - sub * nchk) << 4;
    const int64_t t = row_base + (int64_t)sub * TILE_M;
    ok = real_tile && t < p.L;
    off = chan_base + (int64_t)p.tt.out_stride * t + (int64_t)c0 * cs;
  };
  if constexpr (SMALLN) {
    // n_tile <= 32 (one or two units, MSUB == 1): the streaming layers.  Everything is preloaded before the
    // accumulator wait and the code is short enough to unroll; fits 64 registers -> 8 CTAs per SM.
    int64_t off; bool valid;
    unit_ptr(0, off, valid);
    float res[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) res[c] = (valid && has_res && c < p.n_tile) ? p.residual[off + c * cs] : 0.f;
    mbar_wait(bar_acc, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncwarp();
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += 16) {
      if (c0 < p.n_tile) {
        uint32_t r[16];
        tmem_ld16(trow + c0, r);
        if (valid) {
          float v[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] = __uint_as_float(r[c]);
          if (p.bias) {
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] += __ldg(p.bias + co0 + c0 + c);
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] += res[c0 + c];
          if (p.acc_mode == 1) {
#pragma unroll
            for (int c = 0; c < 16; ++c) p.acc[off + (c0 + c) * cs] = v[c];
          } else if (p.acc_mode == 2) {
#pragma unroll
            for (int c = 0; c < 16; ++c) atomicAdd(p.acc + off + (c0 + c) * cs, v[c]);
          }
          if (p.out) {
#pragma unroll
            for (int c = 0; c < 16; ++c) p.out[off + (c0 + c) * cs] = v[c];
          }
        }
      }
    }
  } else {
  float ra[16], rb[16];  // residual ring: ra = unit u, rb = unit u+1
  {
    int64_t off; bool ok;
    unit_ptr(0, off, ok);
#pragma unroll
    for (int c = 0; c < 16; ++c) ra[c] = (ok && has_res) ? p.residual[off + c * cs] : 0.f;
    unit_ptr(nunits > 1 ? 1 : 0, off, ok);
    ok = ok && nunits > 1;
#pragma unroll
    for (int c = 0; c < 16; ++c) rb[c] = (ok && has_res) ? p.residual[off + c * cs] : 0.f;
  }
  mbar_wait(bar_acc, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncwarp();
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
  for (int u = 0; u < nunits; ++u) {
    const int sub = u / nchk, c0 = (u - sub * nchk) << 4;
    int64_t off; bool valid;
    unit_ptr(u, off, valid);
    uint32_t r[16];
    tmem_ld16(trow + (uint32_t)(sub * p.n_tile + c0), r);
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = __uint_as_float(r[c]);
    if (p.bias) {
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] += __ldg(p.bias + co0 + c0 + c);
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] += ra[c];  // (conv + bias) + residual: the reference's order
    // rotate the ring and fetch unit u+2 before this unit's stores
#pragma unroll
    for (int c = 0; c < 16; ++c) ra[c] = rb[c];
    {
      int64_t offn; bool okn;
      unit_ptr(u + 2 < nunits ? u + 2 : u, offn, okn);
      okn = okn && has_res && u + 2 < nunits;
#pragma unroll
      for (int c = 0; c < 16; ++c) rb[c] = okn ? p.residual[offn + c * cs] : 0.f;
    }
    if (valid) {
      if (p.acc_mode == 1) {
#pragma unroll
        for (int c = 0; c < 16; ++c) p.acc[off + c * cs] = v[c];
      } else if (p.acc_mode == 2) {
        // red.global.add: no read, one add per element per kernel -> deterministic given stream order
#pragma unroll
        for (int c = 0; c < 16; ++c) atomicAdd(p.acc + off + c * cs, v[c]);
      }
      if (p.out) {
#pragma unroll
        for (int c = 0; c < 16; ++c) p.out[off + c * cs] = v[c];
      }
    }
  }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols)
                 : "memory");
  }
  if (CL > 1) cluster_sync_all();  // no CTA exits while peers may still multicast into it / arrive on its barriers
}

// out[ph][nt][s][c2][n][e] = W(co = nt*n_tile + n, ci = 16*kc + 8*c2 + e, tap wj[ph][i]),  s = i*(Cin/16) + kc
// element strides (s_co, s_ci) select Conv1d [Cout,Cin,k] or ConvTranspose1d [Cin,Cout,k] weights
__global__ void pack_weight_kernel(const float *__restrict__ w, __half *__restrict__ out, int Cout, int Cin,
                                   int k, int n_tile, int NB, int64_t s_co, int64_t s_ci, const TapTable tt) {
  const int KC = Cin >> 4;
  const int nco = Cout / n_tile;
  int64_t ph_base = 0;
  for (int ph = 0; ph < tt.nphase; ++ph) {
    const int64_t cnt = (int64_t)tt.ntaps[ph] * KC * nco * 2 * NB * 8;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * blockDim.x) {
      int64_t r = i;
      const int e = r % 8; r /= 8;
      const int n = r % NB; r /= NB;
      const int c2 = r % 2; r /= 2;
      const int s = r % (tt.ntaps[ph] * KC); r /= (tt.ntaps[ph] * KC);
      const int nt = (int)r;
      const int ti = s / KC, kc = s % KC;
      const int co = nt * n_tile + n, ci = 16 * kc + 8 * c2 + e;
      out[ph_base + i] = n < n_tile ? __float2half_rn(w[co * s_co + ci * s_ci + tt.wj[ph][ti]]) : __float2half_rn(0.f);
    }
    ph_base += cnt;
  }
}

TapTable conv_taps(int k, int d) {
  TapTable tt = {};
  tt.nphase = 1;
  tt.out_stride = 1;
  tt.ntaps[0] = k;
  const int h = ((k - 1) / 2) * d;
  for (int j = 0; j < k; ++j) {
    tt.row_off[0][j] = (int8_t)((j - (k - 1) / 2) * d);
    tt.wj[0][j] = (int8_t)j;
  }
  tt.h_lo = h;
  tt.h_hi = h;
  return tt;
}

// ConvTranspose1d(k, stride u, padding p=(k-u)/2): output phase rho = o mod u reads input rows q + c - i
// through taps j = r + i*u, r = (rho+p) mod u, c = (rho+p) div u   
TapTable convT_taps(int k, int u) {
  TapTable tt = {};
  tt.nphase = u;
  tt.out_stride = u;
  const int p = (k - u) / 2;
  int lo = 0, hi = 0;
  for (int rho = 0; rho < u; ++rho) {
    const int r = (rho + p) % u, c = (rho + p) / u;
    int n = 0;
    for (int j = r; j < k; j += u, ++n) {
      tt.row_off[rho][n] = (int8_t)(c - n);
      tt.wj[rho][n] = (int8_t)j;
      if (c - n < lo) lo = c - n;
      if (c - n > hi) hi = c - n;
    }
    tt.ntaps[rho] = n;
    tt.out_off[rho] = rho;
  }
  tt.h_lo = -lo;
  tt.h_hi = hi;
  return tt;
}

int g_host_debug = 0;
int g_cluster_override = 0;  // bring-up aid: force the cluster size (0 = automatic)
int g_msub_override = 0;     // bring-up aid: force the sub-tiles per CTA (0 = automatic)
int g_apad = 0, g_bpad = 0;  // experiment: extra rows between the K-chunks of the A tile / packed weights

template <int MSUB, int MINB, bool SMALLN>
int launch_variant(const Params &p, dim3 grid, int cluster, size_t smem, cudaStream_t st, const char *what) {
  // opt-in dynamic shared memory: 227 KB per block minus the kernel's static shared memory
  static int max_dyn[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (max_dyn[dev] == 0) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, conv_umma_kernel<MSUB, MINB, SMALLN>);
    int want = 227 * 1024 - (e == cudaSuccess ? (int)fa.sharedSizeBytes : 1024);
    want &= ~1023;
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_umma_kernel<MSUB, MINB, SMALLN>, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    if (e != cudaSuccess) {
      cudaGetLastError();  // clear
      hsv::set_error("%s: cudaFuncSetAttribute(%d): %s", what, want, cudaGetErrorString(e));
      return HSV_ERR_CUDA;
    }
    max_dyn[dev] = want;
  }
  HSV_REQUIRE(smem <= (size_t)max_dyn[dev], "%s: shared memory %zu B exceeds %d B (Cin=%d)", what, smem,
              max_dyn[dev], p.Cin);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = hsv::g_pdl ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_umma_kernel<MSUB, MINB, SMALLN>, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    hsv::set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return HSV_ERR_CUDA;
  }
  return hsv::check_launch(what);
}

int launch(const TapTable &tt, const void *a_blk16, const void *w_packed, const float *bias,
           const float *residual, float *out, float *acc, int acc_mode, float acc_div, int B, int Cin, int Cout,
           int64_t L, int64_t Lout, int n_tile, cudaStream_t st, const char *what) {
  Params p;
  p.a = reinterpret_cast<const uint4 *>(a_blk16);
  p.w = reinterpret_cast<const uint4 *>(w_packed);
  p.bias = bias; p.residual = residual; p.out = out; p.acc = acc;
  p.acc_mode = acc_mode; p.acc_div = acc_div;
  p.Cin = Cin; p.Cout = Cout; p.L = L; p.Lp = hsv::blk16_rows(L); p.Lout = Lout;
  p.n_tile = n_tile; p.nco_tiles = Cout / n_tile;
  const int ntiles128 = (int)((L + TILE_M - 1) / TILE_M);
  // Measured on B200: with the SWIZZLE_NONE operand layout one 128x128x16 MMA occupies the tensor pipe for
  // ~150 cycles (vs 64 at peak), so the kernel is MMA-bound, not weight-ingest-bound, and extra sub-tiles
  // per CTA only lengthen the serial MMA chain of each CTA.  Default 1; 2/4 stay available (override) for
  // the throughput regime once the operands move to a swizzled layout.
  int msub = 1;
  if (g_msub_override > 0) msub = g_msub_override;
  if (msub == 3) msub = 2;
  // must fit: TMEM columns, tiles available, and the A tile (+ a 2-stage weight ring) in shared memory
  while (msub > 1 && (msub * n_tile > 512 || msub > ntiles128 ||
                      (size_t)(TILE_M * msub + tt.h_lo + tt.h_hi + g_apad) * 16 * (Cin / 8) + 2 * 16384 > 200 * 1024))
    msub >>= 1;
  p.msub = msub;
  p.ntiles = (ntiles128 + msub - 1) / msub;
  p.R = TILE_M * msub + tt.h_lo + tt.h_hi;
  p.Rs = p.R + g_apad;
  p.NB = n_tile + g_bpad;
  p.tt = tt;
  int max_ksteps = 0;
  for (int q = 0; q < tt.nphase; ++q) max_ksteps = tt.ntaps[q] * (Cin / 16) > max_ksteps ? tt.ntaps[q] * (Cin / 16) : max_ksteps;
  const int kstep_bytes = 32 * p.NB;
  // weight block = G K-steps (<= 16 KB), tap-aligned: whole taps if one tap fits, else a divisor of a tap
  const int KC = Cin / 16;
  const int64_t total_ctas = (int64_t)p.ntiles * p.nco_tiles * tt.nphase * B;
  int gmax = (total_ctas <= 148 ? 32768 : 16384) / kstep_bytes;
  if (gmax < 1) gmax = 1;
  int G;
  if (KC <= gmax) {
    G = KC * (gmax / KC);
    if (G > max_ksteps) G = max_ksteps;  // max_ksteps is a multiple of KC
  } else {
    G = 1;
    for (int q = gmax; q >= 1; --q)
      if (KC % q == 0) {
        G = q;
        break;
      }
  }
  p.G = G;
  const int nblocks = (max_ksteps + G - 1) / G;
  p.stages = nblocks < MAX_STAGES ? nblocks : MAX_STAGES;
  uint32_t cols = 32;
  while ((int)cols < n_tile * msub) cols <<= 1;
  p.tmem_cols = cols;
  p.debug = g_host_debug;

  const size_t a_bytes = ((size_t)p.Rs * 16 * (Cin / 8) + 127) & ~(size_t)127;
  // shared-memory budget: leave room for as many co-resident CTAs per SM as the grid can use (they hide
  // each other's prologue / epilogue latency), down to a 2-stage weight ring
  const int minb = n_tile <= 32 ? 8 : (n_tile <= 64 ? 4 : 2);
  int want = (int)((total_ctas + 147) / 148);
  want = want < 1 ? 1 : (want > minb ? minb : want);
  const size_t budget = (size_t)(226 * 1024) / want - 1024;
  size_t smem = a_bytes + (size_t)p.stages * G * kstep_bytes;
  while (smem > budget && p.stages > 2) {
    p.stages--;
    smem = a_bytes + (size_t)p.stages * G * kstep_bytes;
  }
  HSV_REQUIRE(B <= 65535 && (int64_t)p.nco_tiles * tt.nphase <= 65535, "%s: grid too large", what);
  // weight multicast across a cluster of consecutive M-tiles pays when the weights are large
  // (n_tile >= 64 -> K-step >= 2 KB) and there are at least two tiles to share them
  int cluster = 1;
  if (g_cluster_override > 0) cluster = g_cluster_override;
  // (measured on B200: clusters cost more in launch/co-scheduling than the multicast saves at these
  //  sizes, so the automatic choice is 1; the path stays available through the bring-up override)
  if (n_tile < 64) cluster = 1;
  while (cluster > 1 && (32 * p.NB) % (16 * cluster) != 0) cluster >>= 1;
  const int gx = ((p.ntiles + cluster - 1) / cluster) * cluster;
  dim3 grid((unsigned)gx, (unsigned)(p.nco_tiles * tt.nphase), (unsigned)B);
  // small n_tile = HBM/latency-bound streaming layers: they want many co-resident CTAs (register cap 80);
  // large n_tile = few fat CTAs per SM anyway
  if (p.msub == 1) {
    if (n_tile <= 32) return launch_variant<1, 8, true>(p, grid, cluster, smem, st, what);
    if (n_tile <= 64) return launch_variant<1, 4, false>(p, grid, cluster, smem, st, what);
    return launch_variant<1, 2, false>(p, grid, cluster, smem, st, what);
  }
  if (p.msub == 2) return launch_variant<2, 3, false>(p, grid, cluster, smem, st, what);
  return launch_variant<4, 2, false>(p, grid, cluster, smem, st, what);
}

int check_common(const char *what, const void *a, const void *w, int Cin, int Cout, int n_tile) {
  HSV_REQUIRE(a && w, "%s: null operand", what);
  HSV_REQUIRE(Cin > 0 && Cin % 16 == 0, "%s: Cin %% 16 != 0 (Cin=%d)", what, Cin);
  HSV_REQUIRE(n_tile >= 16 && n_tile <= 128 && n_tile % 16 == 0 && Cout % n_tile == 0,
              "%s: bad n_tile=%d for Cout=%d", what, n_tile, Cout);
  return HSV_OK;
}

int pack(const float *w, void *packed, int Cout, int Cin, int k, int n_tile, int64_t s_co, int64_t s_ci,
         const TapTable &tt, cudaStream_t st, const char *what) {
  const int64_t total = (int64_t)Cout * Cin * k;
  const int blocks = (int)((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
  pack_weight_kernel<<<blocks, 256, 0, st>>>(w, reinterpret_cast<__half *>(packed), Cout, Cin, k, n_tile,
                                             n_tile + g_bpad, s_co, s_ci, tt);
  return hsv::check_launch(what);
}

}  // namespace

// bring-up aid only (bit0: swap LBO/SBO roles in the smem descriptors); not part of the drop-in contract
int hsv_v1::set_umma_debug(int flags) {
  g_host_debug = flags & 0xff;
  g_cluster_override = (flags >> 8) & 0xff;  // bits 8..15: forced cluster size
  g_apad = (flags >> 16) & 0xf;              // bits 16..19: A chunk padding rows
  g_bpad = (flags >> 20) & 0xf;              // bits 20..23: weight chunk padding rows
  g_msub_override = (flags >> 24) & 0x7;     // bits 24..26: forced sub-tiles per CTA
  return HSV_OK;
}

int hsv_v1::pack_conv_weight(const float *w, void *packed, int Cout, int Cin, int k, int n_tile,
                                    void *stream) {
  if (int rc = check_common("pack_conv_weight", w, packed, Cin, Cout, n_tile)) return rc;
  HSV_REQUIRE(k >= 1 && k <= MAX_TAPS && (k & 1), "pack_conv_weight: k=%d (odd, <= %d)", k, MAX_TAPS);
  return pack(w, packed, Cout, Cin, k, n_tile, (int64_t)Cin * k, k, conv_taps(k, 1), hsv::as_stream(stream),
              "pack_conv_weight");
}

int hsv_v1::pack_convT_weight(const float *w, void *packed, int Cin, int Cout, int k, int u, int n_tile,
                                     void *stream) {
  if (int rc = check_common("pack_convT_weight", w, packed, Cin, Cout, n_tile)) return rc;
  HSV_REQUIRE(u >= 1 && u <= MAX_PHASES && k >= u && k <= MAX_TAPS && k - 2 * ((k - u) / 2) == u,
              "pack_convT_weight: unsupported (k,u)=(%d,%d)", k, u);
  return pack(w, packed, Cout, Cin, k, n_tile, k, (int64_t)Cout * k, convT_taps(k, u), hsv::as_stream(stream),
              "pack_convT_weight");
}

int hsv_v1::conv1d_umma(const void *a_blk16, const void *w_packed, const float *bias,
                               const float *residual, float *out, float *acc, int acc_mode, float acc_div,
                               int B, int Cin, int Cout, int64_t L, int k, int d, int n_tile, void *stream) {
  if (B == 0 || L == 0) return HSV_OK;  // empty batch / sequence
  if (int rc = check_common("conv1d_umma", a_blk16, w_packed, Cin, Cout, n_tile)) return rc;
  HSV_REQUIRE(k >= 1 && k <= MAX_TAPS && (k & 1) && d >= 1, "conv1d_umma: k must be odd and <= %d (k=%d d=%d)",
              MAX_TAPS, k, d);
  HSV_REQUIRE(((k - 1) / 2) * d <= HSV_BLK_PAD, "conv1d_umma: halo %d exceeds blk16 padding %d",
              ((k - 1) / 2) * d, HSV_BLK_PAD);
  HSV_REQUIRE(acc_mode >= 0 && acc_mode <= 2 && (acc_mode == 0 || acc), "conv1d_umma: bad acc_mode/acc");
  HSV_REQUIRE(out || acc_mode, "conv1d_umma: no output");
  if (B == 0 || L == 0) return HSV_OK;
  return launch(conv_taps(k, d), a_blk16, w_packed, bias, residual, out, acc, acc_mode, acc_div, B, Cin, Cout, L,
                L, n_tile, hsv::as_stream(stream), "conv1d_umma");
}

int hsv_v1::conv_transpose1d_umma(const void *a_blk16, const void *w_packed, const float *bias,
                                         const float *add, float *out, int B, int Cin, int Cout, int64_t Lin,
                                         int k, int u, int n_tile, void *stream) {
  if (B == 0 || Lin == 0) return HSV_OK;
  if (int rc = check_common("conv_transpose1d_umma", a_blk16, w_packed, Cin, Cout, n_tile)) return rc;
  HSV_REQUIRE(u >= 1 && u <= MAX_PHASES && k >= u && k <= MAX_TAPS && k - 2 * ((k - u) / 2) == u,
              "conv_transpose1d_umma: unsupported (k,u)=(%d,%d)", k, u);
  HSV_REQUIRE(out, "conv_transpose1d_umma: no output");
  if (B == 0 || Lin == 0) return HSV_OK;
  return launch(convT_taps(k, u), a_blk16, w_packed, bias, add, out, nullptr, 0, 1.f, B, Cin, Cout, Lin,
                (int64_t)u * Lin, n_tile, hsv::as_stream(stream), "conv_transpose1d_umma");
}
