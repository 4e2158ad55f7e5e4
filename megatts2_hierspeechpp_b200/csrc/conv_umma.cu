// Dilated 'same' Conv1d as a tcgen05 / TMEM implicit GEMM (sm_100a).
//
// Replaces the weight-normed AMPBlock convs of the reference
// (hierspeechpp_speechsynthesizer.py:349-364,380-384; speechsr24k/speechsr.py:21-36,52-56):
//   out[b,co,t] = bias[co] + sum_{ci,j} W[co,ci,j] * a[b,ci,t + (j-(k-1)/2)*d]      (zero padded)
// GEMM view per CTA:  D[M=128 time rows, N=n_tile out channels] = sum over taps j and 16-channel
// K-steps of  A_j[128 x 16] * W_j[16 x n_tile],  fp16 operands, fp32 accumulation in TMEM.
//
// Operand staging.  The activation operand arrives in the "blk16" layout written by the fused
// activation kernel: fp16 [B][Cin/8][Lp][8], i.e. for each 8-channel chunk the time rows are
// contiguous 16-byte records.  A time tile (+halo) of one chunk is therefore ONE contiguous span,
// fetched with a 1-D bulk TMA copy (cp.async.bulk) straight into shared memory as
// [chunk][row][8 halves].  That is exactly the tcgen05 K-major SWIZZLE_NONE canonical layout with
// SBO = 128 B (8 rows x 16 B) and LBO = rows*16 B, in which consecutive rows of one K-chunk are
// uniformly 16 B apart -- so the operand of tap j is the SAME shared tile with the descriptor start
// address advanced by j*d rows.  The halo is loaded once and reused by all k taps; zero padding
// comes from the zero rows the blk16 layout keeps around every sequence.
// Weights are pre-packed (hsv_pack_conv_weight) as [n_tile block][K-step][2][n_tile][8] fp16 so a
// group of K-steps is again one contiguous span, streamed through a ring of shared stages by bulk
// TMA copies with mbarrier completion.
//
// Roles (128 threads): warp0/lane0 TMA producer, warp1/lane0 MMA issuer, warp2 TMEM alloc/free,
// then all four warps run the epilogue: tcgen05.ld (lane = time row), + bias, + residual, store
// fp32 [B,C,L] (a warp stores 32 consecutive time steps of one channel: 128 B coalesced) and
// optionally accumulate the mean over resblocks.
#include "hsv_common.cuh"

namespace {

constexpr int TILE_M = HSV_UMMA_TILE_M;  // 128
constexpr int MAX_STAGES = 4;

struct Params {
  const uint4 *a;      // blk16 activations, 16-byte records
  const uint4 *w;      // packed weights
  const float *bias;
  const float *residual;
  float *out;
  float *acc;
  int acc_mode;
  float acc_div;
  int Cin, Cout;
  int64_t L, Lp;
  int k, d, n_tile;
  int R;             // rows per chunk in the shared A tile = 128 + (k-1)*d
  int ksteps;        // k * Cin/16
  int G;             // K-steps per weight block
  int nblocks;       // ceil(ksteps / G)
  int stages;
  uint32_t tmem_cols;
  int debug;
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // try_wait suspends in hardware; the clock bound turns a protocol bug into a trap instead of a hang
  uint32_t ok;
  const long long t_start = clock64();
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && clock64() - t_start > 4000000000ll) {
      printf("hsv conv1d_umma: mbarrier wait timed out (block %d,%d,%d thread %d bar %u parity %u)\n", blockIdx.x,
             blockIdx.y, blockIdx.z, threadIdx.x, bar, parity);
      __trap();
    }
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
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

__global__ void __launch_bounds__(128) conv1d_umma_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 + 2 * MAX_STAGES];  // a_full, acc_full, w_full[S], w_empty[S]
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, nt = blockIdx.y, b = blockIdx.z;
  const int nchunks = p.Cin >> 3;
  const int KC = p.Cin >> 4;
  const int h = ((p.k - 1) >> 1) * p.d;
  const uint32_t a_bytes_chunk = (uint32_t)p.R * 16u;
  const uint32_t a_bytes = a_bytes_chunk * nchunks;
  const uint32_t kstep_bytes = 32u * p.n_tile;
  const uint32_t wblk_bytes = kstep_bytes * p.G;

  const uint32_t a_s = smem_u32(smem);
  const uint32_t w_s = a_s + ((a_bytes + 127u) & ~127u);
  const uint32_t bar_a = smem_u32(&bars[0]), bar_acc = smem_u32(&bars[1]);
  const uint32_t bar_wf = smem_u32(&bars[2]), bar_we = smem_u32(&bars[2 + MAX_STAGES]);

  if (threadIdx.x == 0) {
    mbar_init(bar_a, 1);
    mbar_init(bar_acc, 1);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_wf + 8 * s, 1);
      mbar_init(bar_we + 8 * s, 1);
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
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (warp == 0 && lane == 0) {
    // ---------------- TMA producer ----------------
    const int64_t row0 = (int64_t)HSV_BLK_PAD + (int64_t)tile * TILE_M - h;
    mbar_expect_tx(bar_a, a_bytes);
    for (int q = 0; q < nchunks; ++q) {
      const uint4 *src = p.a + ((int64_t)b * nchunks + q) * p.Lp + row0;
      bulk_g2s(a_s + q * a_bytes_chunk, src, a_bytes_chunk, bar_a);
    }
    const uint4 *wsrc = p.w + (int64_t)nt * p.ksteps * (kstep_bytes >> 4);
    for (int blk = 0; blk < p.nblocks; ++blk) {
      const int s = blk % p.stages;
      if (blk >= p.stages) mbar_wait(bar_we + 8 * s, ((blk / p.stages) - 1) & 1);
      const int nk = min(p.G, p.ksteps - blk * p.G);
      const uint32_t bytes = kstep_bytes * nk;
      mbar_expect_tx(bar_wf + 8 * s, bytes);
      bulk_g2s(w_s + s * wblk_bytes, wsrc + (int64_t)blk * (wblk_bytes >> 4), bytes, bar_wf + 8 * s);
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------- MMA issuer ----------------
    // InstrDescriptor: D=F32 (1<<4), A=B=F16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
    const bool swap = p.debug & 1;
    mbar_wait(bar_a, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int blk = 0; blk < p.nblocks; ++blk) {
      const int s = blk % p.stages;
      mbar_wait(bar_wf + 8 * s, (blk / p.stages) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int nk = min(p.G, p.ksteps - blk * p.G);
      for (int g = 0; g < nk; ++g) {
        const int ks = blk * p.G + g;
        const int j = ks / KC, kc = ks - j * KC;
        const uint32_t a_addr = a_s + (uint32_t)(2 * kc * p.R + j * p.d) * 16u;
        const uint32_t b_addr = w_s + s * wblk_bytes + g * kstep_bytes;
        const uint32_t a_lbo = a_bytes_chunk, b_lbo = 16u * p.n_tile, sbo = 128u;
        const uint64_t ad = swap ? make_desc(a_addr, sbo, a_lbo) : make_desc(a_addr, a_lbo, sbo);
        const uint64_t bd = swap ? make_desc(b_addr, sbo, b_lbo) : make_desc(b_addr, b_lbo, sbo);
        umma_f16(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
      }
      umma_commit(bar_we + 8 * s);
    }
    umma_commit(bar_acc);
  }

  // ---------------- epilogue: all 4 warps ----------------
  __syncwarp();
  mbar_wait(bar_acc, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncwarp();
  const int64_t t = (int64_t)tile * TILE_M + warp * 32 + lane;
  const bool valid = t < p.L;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(trow + c0, r);
    if (valid) {
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int co = nt * p.n_tile + c0 + c;
        const int64_t off = ((int64_t)b * p.Cout + co) * p.L + t;
        float v = __uint_as_float(r[c]);
        if (p.bias) v += __ldg(p.bias + co);
        if (p.residual) v += p.residual[off];
        if (p.out) p.out[off] = v;
        if (p.acc_mode == 1) p.acc[off] = v;
        else if (p.acc_mode == 2) p.acc[off] += v;
        else if (p.acc_mode == 3) p.acc[off] = (p.acc[off] + v) / p.acc_div;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols)
                 : "memory");
  }
}

__global__ void pack_weight_kernel(const float *__restrict__ w, __half *__restrict__ out, int Cout, int Cin,
                                   int k, int n_tile) {
  // out[nt][s][c2][n][e] = w[co = nt*n_tile + n][ci = 16*kc + 8*c2 + e][j],  s = j*(Cin/16) + kc
  const int64_t total = (int64_t)Cout * Cin * k;
  const int KC = Cin >> 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int e = r % 8; r /= 8;
    const int n = r % n_tile; r /= n_tile;
    const int c2 = r % 2; r /= 2;
    const int s = r % (k * KC); r /= (k * KC);
    const int nt = (int)r;
    const int j = s / KC, kc = s % KC;
    const int co = nt * n_tile + n, ci = 16 * kc + 8 * c2 + e;
    out[i] = __float2half_rn(w[((int64_t)co * Cin + ci) * k + j]);
  }
}

}  // namespace

// bring-up aid only (bit0: swap LBO/SBO roles in the smem descriptors); not part of the drop-in contract
static int g_host_debug = 0;
extern "C" int hsv_set_umma_debug(int flags) {
  g_host_debug = flags;
  return HSV_OK;
}

extern "C" int hsv_pack_conv_weight(const float *w, void *packed, int Cout, int Cin, int k, int n_tile,
                                    void *stream) {
  HSV_REQUIRE(w && packed, "pack_conv_weight: null pointer");
  HSV_REQUIRE(Cin > 0 && Cin % 16 == 0, "pack_conv_weight: Cin %% 16 != 0 (Cin=%d)", Cin);
  HSV_REQUIRE(n_tile >= 16 && n_tile <= 256 && n_tile % 16 == 0 && Cout % n_tile == 0,
              "pack_conv_weight: bad n_tile=%d for Cout=%d", n_tile, Cout);
  HSV_REQUIRE(k >= 1, "pack_conv_weight: k=%d", k);
  const int64_t total = (int64_t)Cout * Cin * k;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pack_weight_kernel<<<blocks, 256, 0, hsv::as_stream(stream)>>>(w, reinterpret_cast<__half *>(packed), Cout,
                                                                 Cin, k, n_tile);
  return hsv::check_launch("pack_conv_weight");
}

extern "C" int hsv_conv1d_umma(const void *a_blk16, const void *w_packed, const float *bias,
                               const float *residual, float *out, float *acc, int acc_mode, float acc_div,
                               int B, int Cin, int Cout, int64_t L, int k, int d, int n_tile, void *stream) {
  HSV_REQUIRE(a_blk16 && w_packed, "conv1d_umma: null operand");
  HSV_REQUIRE(Cin > 0 && Cin % 16 == 0, "conv1d_umma: Cin %% 16 != 0 (Cin=%d)", Cin);
  HSV_REQUIRE(n_tile >= 16 && n_tile <= 256 && n_tile % 16 == 0 && Cout % n_tile == 0,
              "conv1d_umma: bad n_tile=%d for Cout=%d", n_tile, Cout);
  HSV_REQUIRE(k >= 1 && (k & 1) && d >= 1, "conv1d_umma: k must be odd (k=%d d=%d)", k, d);
  HSV_REQUIRE(((k - 1) / 2) * d <= HSV_BLK_PAD, "conv1d_umma: halo %d exceeds blk16 padding %d",
              ((k - 1) / 2) * d, HSV_BLK_PAD);
  HSV_REQUIRE(acc_mode >= 0 && acc_mode <= 3 && (acc_mode == 0 || acc), "conv1d_umma: bad acc_mode/acc");
  HSV_REQUIRE(out || acc_mode, "conv1d_umma: no output");
  if (B == 0 || L == 0) return HSV_OK;
  HSV_REQUIRE(B <= 65535 && Cout / n_tile <= 65535, "conv1d_umma: grid too large");

  Params p;
  p.a = reinterpret_cast<const uint4 *>(a_blk16);
  p.w = reinterpret_cast<const uint4 *>(w_packed);
  p.bias = bias; p.residual = residual; p.out = out; p.acc = acc;
  p.acc_mode = acc_mode; p.acc_div = acc_div;
  p.Cin = Cin; p.Cout = Cout; p.L = L; p.Lp = hsv::blk16_rows(L);
  p.k = k; p.d = d; p.n_tile = n_tile;
  p.R = TILE_M + (k - 1) * d;
  p.ksteps = k * (Cin / 16);
  const int kstep_bytes = 32 * n_tile;
  int G = 32768 / kstep_bytes;
  if (G < 1) G = 1;
  if (G > p.ksteps) G = p.ksteps;
  p.G = G;
  p.nblocks = (p.ksteps + G - 1) / G;
  p.stages = p.nblocks < MAX_STAGES ? p.nblocks : MAX_STAGES;
  uint32_t cols = 32;
  while ((int)cols < n_tile) cols <<= 1;
  p.tmem_cols = cols;
  p.debug = g_host_debug;

  const size_t a_bytes = ((size_t)p.R * 16 * (Cin / 8) + 127) & ~(size_t)127;
  size_t smem = a_bytes + (size_t)p.stages * G * kstep_bytes;
  // shrink the weight ring if the A tile is large
  while (smem > 200 * 1024 && p.stages > 2) {
    p.stages--;
    smem = a_bytes + (size_t)p.stages * G * kstep_bytes;
  }
  // opt-in dynamic shared memory: 227 KB per block minus the kernel's static shared memory
  static int max_dyn[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (max_dyn[dev] == 0) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, conv1d_umma_kernel);
    int want = 227 * 1024 - (e == cudaSuccess ? (int)fa.sharedSizeBytes : 1024);
    want &= ~1023;
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv1d_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    if (e != cudaSuccess) {
      cudaGetLastError();  // clear
      hsv::set_error("conv1d_umma: cudaFuncSetAttribute(%d): %s", want, cudaGetErrorString(e));
      return HSV_ERR_CUDA;
    }
    max_dyn[dev] = want;
  }
  HSV_REQUIRE(smem <= (size_t)max_dyn[dev], "conv1d_umma: shared memory %zu B exceeds %d B (Cin=%d k=%d d=%d)",
              smem, max_dyn[dev], Cin, k, d);
  dim3 grid((unsigned)((L + TILE_M - 1) / TILE_M), (unsigned)(Cout / n_tile), (unsigned)B);
  conv1d_umma_kernel<<<grid, 128, smem, hsv::as_stream(stream)>>>(p);
  return hsv::check_launch("conv1d_umma");
}
