// Fused anti-aliased activation with BOTH FIRs on the tensor cores (tcgen05 / TMEM), sm_100a.
//
// Same operator as act1d.cu (alias_free_torch/act.py:23-27: UpSample1d x2 -> SnakeBeta -> DownSample1d x2, fp32
// [B,C,L] in, fp16 "blk16" tensor-core operand out), different machine mapping.  The CUDA-core kernel spends ~31
// issue slots and 33-38 FP32 lane-operations per output element, 24 of them the two 12-tap FIRs, and is
// FP32-pipe-bound at ~60 % of the HBM roofline while the tensor pipe idles.  Here the FIRs are banded-Toeplitz
// MMAs and the CUDA cores only do what the tensor cores cannot:
//
//   P1  x (fp32, TMA-staged)  -> split into fp16 hi + fp16 lo (x == hi + lo to ~2^-22), interleaved along K, into
//       the swizzled A tile XA                                           (3 ops / element)
//   P2  Y = XA * U         6 K-blocks x {U_hi, U_lo}: 12 MMAs 128 x 48 x 16 into TMEM     (tensor pipe)
//   P3  z = y + 1/(e^beta+1e-9) sin^2(y e^alpha) on packed f32x2, z -> fp16 into the A tile Z   (~9 ops / element)
//   P4  O = Z * D          6 K-blocks: 6 MMAs 128 x 32 x 16 into TMEM (fp16 taps, unit DC gain)  (tensor pipe)
//   P5  O -> fp16 -> staged -> 16-byte units of the blk16 layout                           (~2 ops / element)
//
// Rows.  One MMA row = one (channel, run) pair: 8 channels x 16 runs of 32 time steps = a [8 x 512] window per
// tile, row m = run * 8 + channel; two threads per row (256 threads), each on half of the row's columns.  A FIR output of run r needs the last few samples of run r-1 and the first few
// of run r+1 of the same channel: that is the SAME A tile with the descriptor start address moved by -/+ 8 rows
// (one swizzle atom), so every sample is split / activated exactly once -- no halo recomputation between runs.
// Between tiles the first and the last run are recomputed (outputs of the inner 14 runs are stored: 448 of 512).
// K-blocks and their output windows are shift-invariant, so one [N x 16] Toeplitz matrix per FIR (and window
// alignment) serves every block: 5 KB of constants (tools/gen_act_tables.py, pinned by tests/test_act_mma_model.py).
//
// Precision.  x and the up-sampling taps are hi/lo-split (three-term product), so y -- the input of the
// non-linearity -- keeps ~22 bits; z and the low-pass taps are fp16 (the result is rounded to fp16 for the conv operand
// anyway): emulated on the bundled SpeechSR48 checkpoint, end to end: 53.1 dB SNR / 7.9e-4 max-abs vs 55.7 dB /
// 6.2e-4 for act1d.cu (bar: 40 dB / 2e-3).  Edge semantics are those of act1d.cu: x replicate-clamped on the 1x grid when the tile is
// staged, the activated signal replicate-clamped on the 2x grid (z[0] / z[2L-1] substituted outside [0, 2L)).
#include "act_core.cuh"
#include "act_toeplitz_tables.h"
#include "umma_common.cuh"

namespace {

using namespace hsv_umma;

constexpr int A_RUNS = 16, A_CH = 8, A_RT = 32;
constexpr int A_WIN = A_RUNS * A_RT;             // 512 steps per tile window
constexpr int A_VALID = (A_RUNS - 2) * A_RT;     // 448 outputs stored per tile
constexpr int A_XP = A_WIN + 4;                  // x stage pitch (floats): == 4 (mod 32) -> conflict-free LDS.128
constexpr int A_NT = 256;                        // threads per CTA: two per MMA row
constexpr uint32_t A_TILE = (8 + 128 + 8) * 128; // A tile: 128 rows of 128 bytes + 8 zero rows either side
constexpr uint32_t A_OFF_XA = 0;                 // the x stage aliases the start of the XA tile (see the kernel)
constexpr uint32_t A_OFF_Z = A_TILE;
constexpr uint32_t A_OFF_TAB = 2 * A_TILE;                       // 36864
constexpr uint32_t A_TAB_BYTES = (HSV_TOEP_BYTES + 1023u) & ~1023u;
constexpr uint32_t A_OFF_STG = A_OFF_TAB + A_TAB_BYTES;
constexpr uint32_t A_STG_RUN = 32 * 16 + 32;                     // staging pitch per run: 544 B (8-bank rotation)
constexpr uint32_t A_SMEM = A_OFF_STG + A_RUNS * A_STG_RUN;      // ~50 KB -> four CTAs per SM
constexpr uint32_t A_TMEM_COLS = 128;                            // Y: cols 0..127 (Y[c] at 32 + c); O aliased at 16 + c
constexpr int A_CTAS_PER_SM = 4;
static_assert(A_OFF_STG % 16 == 0 && HSV_TOEP_BYTES % 256 == 0 && A_CH * A_XP * 4 <= A_TILE - 1024, "layout");

struct ActParams {
  const float *x;
  uint8_t *out;          // blk16
  const float *alpha, *beta;
  int C, cw, nchunk, nun, lg;
  int64_t L, Lp;
  int ntiles;
  float sc;
  int aligned;           // x rows 16-byte aligned (L % 4 == 0 and base aligned): TMA staging allowed
};

// one fp16-kind MMA, M = 128, issued by the elected lane of a converged warp
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe, pa;\n\t"
      ".reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, pa;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}

// K-block j = -1 .. 4 of a row: (row shift of the A tile, K-step inside the 128-byte row)
__device__ __forceinline__ void block_of(int j, int &shift_rows, int &ks) {
  shift_rows = j < 0 ? -8 : (j > 3 ? 8 : 0);
  ks = j < 0 ? 3 : (j > 3 ? 0 : j);
}

// mbarrier wait that lets the hardware park the warp (suspend-time hint) instead of spinning through the issue slots
// the other warps need; the iteration bound turns a protocol bug into a trap
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t it = 0; !ok; ++it) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(100000u)
        : "memory");
    if (it > (1u << 22)) __trap();
  }
}

__device__ __forceinline__ uint32_t split_pack(float v) {
  // hi = the top 11 significant bits (exact in fp16), lo = the remainder rounded to fp16; K order (hi, lo)
  const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  const __half2 h = __floats2half2_rn(hi, v - hi);
  return *reinterpret_cast<const uint32_t *>(&h);
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// wait for the outstanding TMEM loads; the registers are in/out operands so that no consumer of r[] can be scheduled
// above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// SnakeBeta on 16 consecutive 2x-rate samples (8 (odd, even) pairs) -> 8 packed fp16x2 words.
// EDGE: samples with row index < lo take zL, samples with index >= hi take zR (base = index of r[0]).
template <bool EDGE>
__device__ __forceinline__ void snake16(const uint32_t (&r)[16], uint32_t (&zw)[8], hsv_act::u64 a2, hsv_act::u64 ib2,
                                        int base, int lo, int hi, float zL, float zR) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const hsv_act::u64 y2 = hsv_act::pk(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
    float to, te;
    hsv_act::upk(hsv_act::mul2(y2, a2), to, te);
    const hsv_act::u64 s = hsv_act::pk(__sinf(to), __sinf(te));
    float zo, ze;
    hsv_act::upk(hsv_act::fma2(hsv_act::mul2(s, ib2), s, y2), zo, ze);
    if (EDGE) {
      const int io = base + 2 * i, ie = io + 1;
      zo = io < lo ? zL : (io >= hi ? zR : zo);
      ze = ie < lo ? zL : (ie >= hi ? zR : ze);
    }
    const __half2 h = __floats2half2_rn(zo, ze);
    zw[i] = *reinterpret_cast<const uint32_t *>(&h);
  }
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// P3 of one half row: 32 samples Y (TMEM cols 32 + 32*half ..) -> SnakeBeta -> fp16 -> 64 bytes of the row in the
// swizzled Z tile.  Both TMEM loads are in flight before the arithmetic starts.
template <bool EDGE>
__device__ __forceinline__ void snake_half_row(uint32_t tcol, uint32_t z_row, uint32_t swz, int half, hsv_act::u64 a2,
                                               hsv_act::u64 ib2, int lo, int hi, float zL, float zR) {
  uint32_t ra[16], rb[16], zw[8];
  tmem_ld16_nowait(tcol, ra);
  tmem_ld16_nowait(tcol + 16u, rb);
  tmem_ld_wait(ra);
  snake16<EDGE>(ra, zw, a2, ib2, 32 * half, lo, hi, zL, zR);
  sts128(z_row + (((uint32_t)(4 * half) ^ swz) << 4), zw[0], zw[1], zw[2], zw[3]);
  sts128(z_row + (((uint32_t)(4 * half + 1) ^ swz) << 4), zw[4], zw[5], zw[6], zw[7]);
  tmem_ld_wait(rb);
  snake16<EDGE>(rb, zw, a2, ib2, 32 * half + 16, lo, hi, zL, zR);
  sts128(z_row + (((uint32_t)(4 * half + 2) ^ swz) << 4), zw[0], zw[1], zw[2], zw[3]);
  sts128(z_row + (((uint32_t)(4 * half + 3) ^ swz) << 4), zw[4], zw[5], zw[6], zw[7]);
}

// grid = (tile CTAs, channel groups): a CTA owns 8 channels of one batch item and walks tiles blockIdx.x,
// blockIdx.x + gridDim.x, ...
template <bool SC1>
__global__ void __launch_bounds__(A_NT, A_CTAS_PER_SM) act1d_mma_kernel(const __grid_constant__ ActParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int m = tid & 127, half = tid >> 7;     // MMA row, and which half of its columns this thread works on
  const int c = m & 7, run = m >> 3;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_x = smem_u32(&bars[0]), bar_y = smem_u32(&bars[1]), bar_o = smem_u32(&bars[2]);

  if (tid == 0) {
    mbar_init(bar_x, 1);
    mbar_init(bar_y, 1);
    mbar_init(bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(A_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero rows around both A tiles (read by the row-shifted MMAs of the first / last run; must be finite).  The front
  // rows of the XA tile are re-zeroed per tile (the x stage aliases them).
  {
    const int i = tid;  // 256 x 16 bytes = 4 x 1 KB
    const uint32_t off = ((i & 64) ? A_OFF_Z : A_OFF_XA) + ((i & 128) ? (A_TILE - 1024u) : 0u) + (uint32_t)(i & 63) * 16u;
    *reinterpret_cast<uint4 *>(gbase + off) = make_uint4(0u, 0u, 0u, 0u);
  }
  // Toeplitz tables (static data)
  for (int i = tid; i < (int)(HSV_TOEP_BYTES / 16); i += A_NT)
    reinterpret_cast<uint4 *>(gbase + A_OFF_TAB)[i] = reinterpret_cast<const uint4 *>(g_act_toeplitz)[i];

  // this CTA's channel group (decoded once)
  const int grp = blockIdx.y;
  const int unit = grp % p.nun, bc = grp / p.nun;
  const int chunk = bc % p.nchunk, b = bc / p.nchunk;
  const int c0 = chunk * p.cw + unit * 8;
  const int64_t row0 = (int64_t)b * p.C + c0;
  const float al = __ldg(p.alpha + c0 + c), be = __ldg(p.beta + c0 + c);
  const float a = expf(al);
  const float ib = 1.0f / (expf(be) + 0.000000001f);
  const hsv_act::u64 a2 = hsv_act::pk(a, a), ib2 = hsv_act::pk(ib, ib);

  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  hsv::pdl_launch_dependents();  // after the TMEM allocation: a dependent can never hold columns this CTA waits for
  hsv::pdl_wait();               // x belongs to the predecessor kernel

  // descriptors (hi words): A tiles = K-major SWIZZLE_128B (8-row groups of 1024 B), tables = K-major SWIZZLE_32B
  const uint32_t a_hi = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);
  const uint32_t b_hi = ((8u * 32u) >> 4) | (1u << 14) | (6u << 29);
  const uint32_t idesc_up = (1u << 4) | ((48u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t idesc_dn = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
  auto desc_lo = [](uint32_t addr) { return (1u << 16) | ((addr & 0x3FFFFu) >> 4); };
  const uint32_t xa0 = base + A_OFF_XA + 8u * 128u, z0 = base + A_OFF_Z + 8u * 128u, tab = base + A_OFF_TAB;

  float *x_s = reinterpret_cast<float *>(gbase + A_OFF_XA);   // aliases the XA tile: dead between P2 and the next P1
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t swz = (uint32_t)(m & 7);
  const uint32_t xa_row = base + A_OFF_XA + (uint32_t)(8 + m) * 128u;
  const uint32_t z_row = base + A_OFF_Z + (uint32_t)(8 + m) * 128u;
  const uint32_t mask = (uint32_t)(p.cw >> 3) - 1u;
  uint8_t *ob = p.out + ((((int64_t)b * p.nchunk + chunk) * p.Lp) << p.lg);
  const uint32_t ub = (uint32_t)unit << 4;
  const float *xg = p.x + row0 * p.L;

  auto is_fast = [&](int tile) {
    const int64_t tw = -A_RT + (int64_t)A_VALID * tile;
    return p.aligned && tw >= 0 && tw + A_WIN <= p.L;
  };
  auto issue_x = [&](int tile) {   // one thread: TMA staging of the x window of a tile
    const int64_t tw = -A_RT + (int64_t)A_VALID * tile;
    mbar_expect_tx(bar_x, A_CH * A_WIN * 4);
#pragma unroll
    for (int r = 0; r < A_CH; ++r) bulk_g2s(smem_u32(x_s + r * A_XP), xg + r * p.L + tw, A_WIN * 4, bar_x);
  };

  // ---- tile-invariant addresses of P5 / copy-out (a warp copies out exactly the staging rows it wrote) ----
  // staging write: lane pairs (channels c, c^1) exchange packed halves so that every lane stores (even channel, odd
  // channel) words; even lanes take the even time steps, odd lanes the odd ones
  const bool odd = (c & 1) != 0;
  const uint32_t perm_sel = odd ? 0x3276u : 0x5410u;
  const uint32_t stg_w = base + A_OFF_STG + (uint32_t)run * A_STG_RUN + (uint32_t)(half * 16) * 16u +
                         (uint32_t)(c >> 1) * 4u + (odd ? 16u : 0u);
  // copy-out: idx = lane + 32 k -> run r_k, step tl_k of the window.  The swizzle term of the blk16 address only
  // depends on (row mod 8) and the tile stride (448 rows) is a multiple of 8: it is a per-thread constant.
  uint32_t stg_r[2];
  uint8_t *gp[2];
  int64_t t_rel[2];      // t - 448 * tile
  bool keep[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = lane + 32 * k;
    const int r = 4 * (warp & 3) + (idx >> 4), tl = r * A_RT + half * 16 + (idx & 15);
    keep[k] = r != 0 && r != A_RUNS - 1;                 // runs 0 and 15 are the recomputed halo of the tile
    stg_r[k] = A_OFF_STG + (uint32_t)r * A_STG_RUN + (uint32_t)(tl & 31) * 16u;
    t_rel[k] = (int64_t)tl - A_RT;
    const uint64_t row = (uint64_t)(HSV_BLK_PAD + t_rel[k] + (int64_t)A_VALID * blockIdx.x);
    const uint64_t lin = (row << p.lg) + ub;
    gp[k] = ob + (lin ^ (((lin >> 7) & mask) << 4));
  }
  const int64_t g_step = ((int64_t)A_VALID * gridDim.x) << p.lg;

  uint32_t ph = 0, xph = 0;
  bool have_x = false;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int64_t tw = -A_RT + (int64_t)A_VALID * tile;
    const bool interior = tw >= 0 && tw + A_WIN <= p.L;
    const bool fast = p.aligned && interior;

    // ---- x window -> shared (the stage aliases the XA tile, free since the previous tile's P2 completed) ----
    if (fast) {
      if (!have_x && tid == A_NT - 32) issue_x(tile);
      mbar_wait_sleep(bar_x, xph);
      xph ^= 1u;
    } else {
      // tiles touching either end of the sequence (or unaligned tensors): replicate-clamped loads, all issued
      // before the first store
      constexpr int NLD = A_CH * A_WIN / A_NT;
      const int64_t Lm1 = p.L - 1;
      float v[NLD];
#pragma unroll
      for (int i = 0; i < NLD; ++i) {
        const int idx = tid + A_NT * i;
        const int r = idx >> 9, pp = idx & (A_WIN - 1);
        int64_t t = tw + pp;
        t = t < 0 ? 0 : (t > Lm1 ? Lm1 : t);
        v[i] = __ldg(xg + r * p.L + t);
      }
#pragma unroll
      for (int i = 0; i < NLD; ++i) {
        const int idx = tid + A_NT * i;
        x_s[(idx >> 9) * A_XP + (idx & (A_WIN - 1))] = v[i];
      }
      __syncthreads();
    }

    // ---- P1: hi/lo split of this thread's 16 steps into its half of the A-tile row (in place: the x stage and
    //      the XA tile share memory, so every thread reads first, then all write) ----
    {
      const float4 *xs4 = reinterpret_cast<const float4 *>(x_s + c * A_XP + run * A_RT + half * 16);
      float4 v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = xs4[q];
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (!SC1) {
          v[q].x *= p.sc; v[q].y *= p.sc; v[q].z *= p.sc; v[q].w *= p.sc;
        }
        sts128(xa_row + (((uint32_t)(4 * half + q) ^ swz) << 4), split_pack(v[q].x), split_pack(v[q].y),
               split_pack(v[q].z), split_pack(v[q].w));
      }
      if (tid < 64) *reinterpret_cast<uint4 *>(gbase + A_OFF_XA + (uint32_t)tid * 16u) = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // XA complete; every thread is done with the previous tile's TMEM reads
    if (warp == A_NT / 32 - 1) {
      // ---- P2: Y = XA * U.  Blocks j = 0 and j = 3 first (disjoint windows covering every used column) ----
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t bh = desc_lo(tab + HSV_TOEP_UP_HI), bl = desc_lo(tab + HSV_TOEP_UP_LO);
      const int order[6] = {0, 3, -1, 1, 2, 4};
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        int sh, ks;
        block_of(order[q], sh, ks);
        const uint32_t a_lo = desc_lo(xa0 + (uint32_t)(sh * 128 + ks * 32));
        const uint32_t d = tmem + (uint32_t)(16 * order[q] + 16);
        mma_f16(d, a_lo, a_hi, bh, b_hi, idesc_up, q >= 2);
        mma_f16(d, a_lo, a_hi, bl, b_hi, idesc_up, 1u);
      }
      umma_commit_elect(bar_y);
    }

    // ---- P3: SnakeBeta on the 2x-rate samples of this half row, fp16 into the Z tile ----
    float zL = 0.f, zR = 0.f;
    int lo = 0, hi = 64;
    if (!interior) {
      // z[0] and z[2L-1] from replicate-clamped global x (as act1d.cu); samples of this row with 2x-grid index < 0
      // take z[0], those beyond 2L-1 take z[2L-1]
      const float *xr = xg + c * p.L;
      float wl[6], wr[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        int64_t tl = -3 + q, tr = p.L - 3 + q;
        tl = tl < 0 ? 0 : (tl > p.L - 1 ? p.L - 1 : tl);
        tr = tr < 0 ? 0 : (tr > p.L - 1 ? p.L - 1 : tr);
        wl[q] = __ldg(xr + tl) * p.sc;
        wr[q] = __ldg(xr + tr) * p.sc;
      }
      zL = hsv_act::up_snake(wl[0], wl[1], wl[2], wl[3], wl[4], wl[5], a, ib).e;
      zR = hsv_act::up_snake(wr[0], wr[1], wr[2], wr[3], wr[4], wr[5], a, ib).o;
      const int64_t n_row = 2 * (tw + (int64_t)A_RT * run) - 1;  // 2x-grid index of this row's first sample
      const int64_t lo64 = -n_row, hi64 = 2 * p.L - n_row;       // first / one-past-last sample index inside [0, 2L)
      lo = lo64 < 0 ? 0 : (lo64 > 64 ? 64 : (int)lo64);
      hi = hi64 < 0 ? 0 : (hi64 > 64 ? 64 : (int)hi64);
    }
    mbar_wait_sleep(bar_y, ph);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // the XA tile is free: stage the next tile's x window behind the rest of this tile
    {
      const int tn = tile + (int)gridDim.x;
      have_x = tn < p.ntiles && is_fast(tn);
      if (have_x && tid == A_NT - 32) issue_x(tn);
    }
    if (interior) snake_half_row<false>(trow + 32u + 32u * (uint32_t)half, z_row, swz, half, a2, ib2, 0, 64, 0.f, 0.f);
    else snake_half_row<true>(trow + 32u + 32u * (uint32_t)half, z_row, swz, half, a2, ib2, lo, hi, zL, zR);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // Z complete; Y fully read (O aliases its columns)
    if (warp == A_NT / 32 - 1) {
      // ---- P4: O = Z * D.  Blocks j = -1 and j = 3 first (disjoint windows covering every used column) ----
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t bo = desc_lo(tab + HSV_TOEP_DN_ODD), bev = desc_lo(tab + HSV_TOEP_DN_EVEN);
      const int order[6] = {-1, 3, 0, 1, 2, 4};
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const int j = order[q];
        int sh, ks;
        block_of(j, sh, ks);
        const uint32_t a_lo = desc_lo(z0 + (uint32_t)(sh * 128 + ks * 32));
        const bool even = (j & 1) == 0;
        const uint32_t d = tmem + (uint32_t)(even ? 8 * j : 8 * j + 8);
        mma_f16(d, a_lo, a_hi, even ? bev : bo, b_hi, idesc_dn, q >= 2);
      }
      umma_commit_elect(bar_o);
    }

    // ---- P5: 16 outputs of this half row -> fp16 -> staging [t][8 channels] -> 16-byte units of the blk16 layout ----
    mbar_wait_sleep(bar_o, ph);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      uint32_t ra[16];
      tmem_ld16_nowait(trow + 16u + 16u * (uint32_t)half, ra);
      tmem_ld_wait(ra);
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const __half2 mine = __floats2half2_rn(__uint_as_float(ra[i]), __uint_as_float(ra[i + 1]));   // steps i, i+1
        const uint32_t mb = *reinterpret_cast<const uint32_t *>(&mine);
        const uint32_t th = __shfl_xor_sync(0xffffffffu, mb, 1);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(stg_w + (uint32_t)i * 16u), "r"(__byte_perm(mb, th, perm_sel))
                     : "memory");
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    {
      const int64_t t0 = (int64_t)A_VALID * tile;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (keep[k] && t0 + t_rel[k] < p.L)
          *reinterpret_cast<uint4 *>(gp[k]) = *reinterpret_cast<const uint4 *>(gbase + stg_r[k]);
        gp[k] += g_step;
      }
    }
    __syncwarp();
    ph ^= 1u;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(A_TMEM_COLS) : "memory");
  }
}

}  // namespace

namespace hsv {

// fp32 [B,C,L] -> fp16 blk16 through the tensor-core kernel.  Returns 1 when the shape is not eligible (caller falls
// back to act1d.cu), 0 on success, negative on error.
int act1d_mma_launch(const float *x, void *out, const float *alpha, const float *beta, int B, int C, int64_t L,
                     float sc, cudaStream_t st) {
  if (C % 16 != 0 || L < 1) return 1;
  static int attr_set[64] = {0};
  const bool sc1 = sc == 1.0f;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  const int smem = (int)A_SMEM + 1024;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(act1d_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(act1d_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("act1d_mma: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
      return HSV_ERR_CUDA;
    }
    attr_set[dev] = 1;
  }
  ActParams p;
  p.x = x;
  p.out = reinterpret_cast<uint8_t *>(out);
  p.alpha = alpha;
  p.beta = beta;
  p.C = C;
  p.cw = blk_cw(C);
  p.nchunk = C / p.cw;
  p.nun = p.cw / 8;
  p.lg = p.cw == 64 ? 7 : (p.cw == 32 ? 6 : 5);
  p.L = L;
  p.Lp = blk16_rows(L);
  p.ntiles = (int)((L + A_VALID - 1) / A_VALID);
  p.sc = sc;
  p.aligned = ((L & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  // grid = (tile CTAs, channel groups): about one resident wave (148 SMs x 4 CTAs), every CTA of a group walking
  // the same number of tiles (+-1)
  const long long groups = (long long)B * p.nchunk * p.nun;
  if (groups > 65535 || (L + A_VALID - 1) / A_VALID > (1ll << 30)) return 1;
  long long per = (148ll * A_CTAS_PER_SM + groups - 1) / groups;      // tile CTAs per group for one wave
  per = per < 1 ? 1 : (per > p.ntiles ? p.ntiles : per);
  const long long iters = (p.ntiles + per - 1) / per;                 // tiles per CTA
  per = (p.ntiles + iters - 1) / iters;                               // same depth, fewest CTAs
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)per, (unsigned)groups);
  cfg.blockDim = dim3(A_NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  cudaError_t e = sc1 ? cudaLaunchKernelEx(&cfg, act1d_mma_kernel<true>, p)
                      : cudaLaunchKernelEx(&cfg, act1d_mma_kernel<false>, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("act1d_mma: launch failed: %s", cudaGetErrorString(e));
    return HSV_ERR_CUDA;
  }
  return check_launch("act1d_mma");
}

}  // namespace hsv
