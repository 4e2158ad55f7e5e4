/*
 * hsv.h — C-ABI of the B200 (sm_100a) waveform-generation kernels.
 *
 * The reference (liuhuang31/Megatts2_HierSpeechpp) is pure Python/PyTorch and
 * has no FFI for this path; its "operator interface" is the set of ATen calls
 * made by the nn.Modules on the hot path.  Each entry point below replaces one
 * such call group and cites it (paths relative to the reference root).  The
 * Python host side (megatts2_hierspeechpp_b200/) binds these with ctypes and
 * mirrors the reference's nn.Module classes on top.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless
 *     named host_*; the caller (PyTorch) owns every buffer, nothing is
 *     allocated or freed here;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on
 *     that stream, does no host synchronisation and is CUDA-graph capturable;
 *   - return value 0 = success, negative = error (hsv_last_error() gives the
 *     message, thread-local);
 *   - activations are fp32, layout [B, C, L] contiguous ("NCL", the
 *     reference's layout) unless stated otherwise;
 *   - "blk16" is the tensor-core operand layout produced by the activation
 *     kernel and consumed by hsv_conv1d_umma: fp16 [B][C/CW][Lp][CW], CW = 64
 *     (C % 64 == 0), else 32 (C % 32 == 0), else 16 channels per row, with
 *     Lp = HSV_BLK_PAD + roundup(L,HSV_BLK_ROUND) + HSV_BLK_PAD rows per (b, chunk);
 *     within the buffer of one (b, chunk) the 16-byte unit at linear byte offset
 *     o is stored at o ^ (((o >> 7) & (CW/8 - 1)) << 4)  -- tcgen05's K-major
 *     SWIZZLE_128B/64B/32B pattern, so a linear (1-D bulk TMA) copy of a row span
 *     into 1024-byte aligned shared memory is a ready MMA operand tile.  Rows
 *     outside [HSV_BLK_PAD, HSV_BLK_PAD+L) must be zero (they implement the
 *     conv's zero padding) and are never written.  Treat the buffer as opaque:
 *     hsv_pack_blk16 / hsv_unpack_blk16 convert from / to fp32 [B,C,L].
 */
#ifndef HSV_H_
#define HSV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSV_VERSION 100
#define HSV_BLK_PAD 32          /* zero rows before/after each blk16 sequence */
#define HSV_UMMA_TILE_M 128     /* rows of one tcgen05 accumulator tile */
#define HSV_BLK_ROUND 512       /* blk16 sequences are padded to a multiple of this many rows (max CTA tile) */

/* error codes */
#define HSV_OK 0
#define HSV_ERR_ARG (-1)        /* bad argument / unsupported shape */
#define HSV_ERR_CUDA (-2)       /* CUDA runtime error at launch */

int hsv_version(void);
const char *hsv_last_error(void);
/* 1 if the running device is compute capability 10.x (tcgen05 capable). */
int hsv_device_supported(void);

/* Rows per (b, chunk) of a blk16 buffer for sequence length L. */
int64_t hsv_blk16_rows(int64_t L);

/* ---- Activation1d(SnakeBeta): alias_free_torch/act.py:23-27 =
 * UpSample1d (resample.py:25-32) o SnakeBeta (activations.py:107-119, log-scale
 * alpha/beta) o DownSample1d (resample.py:46-48 -> filter.py:86-94), with the
 * fixed 12-tap kaiser-sinc filter (filter.py:28-57) — ONE fused kernel.
 *   x      [B,C,L] fp32
 *   alpha, beta [C] fp32 (log scale, as stored in the state_dict)
 *   out_mode 0: out = fp32 [B,C,L]
 *   out_mode 1: out = fp16 blk16 (C % 16 == 0), rows per chunk = hsv_blk16_rows(L)
 *   in_scale    x is multiplied by this first (1/num_kernels: the "xs / self.num_kernels" of
 *               hierspeechpp_speechsynthesizer.py:446 when x is the un-normalised sum over resblocks)
 */
int hsv_act1d_snakebeta(const float *x, void *out, const float *alpha, const float *beta,
                        int B, int C, int64_t L, int out_mode, float in_scale, void *stream);

/* ---- weight norm fold: torch._weight_norm(v, g, dim=0) as applied by the
 * forward pre-hook of torch.nn.utils.weight_norm on every conv of the path
 * (hierspeechpp_speechsynthesizer.py:401,406,349-364; speechsr.py:21-36,73).
 *   v [n0, inner], g [n0]  ->  w [n0, inner] = v * g / ||v||_2(row)
 */
int hsv_weight_norm_fold(const float *v, const float *g, float *w, int n0, int inner, void *stream);

/* ---- pack a folded Conv1d weight [Cout,Cin,k] fp32 into the tcgen05 B-operand
 * stream: fp16 [Cout/n_tile][Cin/CW][k] blocks of n_tile rows x CW channels
 * (K-major, swizzled like blk16).  Cin % 16 == 0, Cout % n_tile == 0,
 * n_tile % 16 == 0, n_tile <= 256.  Size: Cout*Cin*k halves.
 */
int hsv_pack_conv_weight(const float *w, void *packed, int Cout, int Cin, int k, int n_tile, void *stream);

/* ---- dilated 'same' Conv1d as a tcgen05/TMEM implicit GEMM:
 * the AMPBlock convs (hierspeechpp_speechsynthesizer.py:349-364,380-384;
 * speechsr24k/speechsr.py:21-36,52-56): F.conv1d(x, w, b, dilation=d,
 * padding=(k*d-d)/2), fp16 operands, fp32 accumulation, fused epilogue.
 *   a_blk16   fp16 blk16 activations (B*Cin*Lp halves)
 *   w_packed  from hsv_pack_conv_weight (same n_tile)
 *   bias      [Cout] fp32 or NULL
 *   residual  [B,Cout,L] fp32 or NULL     (v = acc + bias + residual)
 *   out       [B,Cout,L] fp32 or NULL     (out = v; may alias residual)
 *   acc       [B,Cout,L] fp32 or NULL, acc_mode: 0 none, 1 acc = v, 2 acc += v (sum over
 *             resblocks, hierspeechpp_speechsynthesizer.py:440-445; the division by
 *             num_kernels (:446) is the consumer's in_scale).  acc_div is reserved (pass 1).
 */
int hsv_conv1d_umma(const void *a_blk16, const void *w_packed, const float *bias,
                    const float *residual, float *out, float *acc, int acc_mode, float acc_div,
                    int B, int Cin, int Cout, int64_t L, int k, int d, int n_tile, void *stream);

/* hsv_conv1d_umma_blk16: the same convolution with an OPERAND-WRITING epilogue: (conv + bias + bc[b]) -> activation ->
 * * mask -> fp16 blk16 buffer that the next hsv_conv1d_umma reads (no fp32 round trip, no pack launch).
 *   mode 0: none;  2: gelu_tanh (FFN_Conv, modules.py:382-388);  3: leaky_relu(0.1);
 *   mode 1: the WN gate tanh(a) * sigmoid(b) (commons.py:108-114): the caller packs the weights / bias / bc with the
 *           output channels ordered as groups of [8 tanh | 8 sigmoid] (channels 8j..8j+7 then H+8j..H+8j+7), the
 *           output buffer has Cout / 2 channels.
 *   bc: [B][Cout] (batch stride bc_stride) added before the activation, or NULL; mask: [B][L] or NULL. */
int hsv_conv1d_umma_blk16(const void *a_blk16, const void *w_packed, const float *bias, void *out_blk16, int mode,
                          const float *bc, int64_t bc_stride, const float *mask, int B, int Cin, int Cout, int64_t L, int k,
                          int d, int n_tile, void *stream);

/* hsv_conv1d_umma_wn_tail: a WN layer's res_skip 1x1 conv (Cin -> 2H) with the layer tail in its epilogue (modules.py:167-174):
 * x = (x + rs[:, :H]) * mask in place, output += rs[:, H:] in place, and the new x written as the next in_layer's fp16
 * blk16 operand.  H must be a multiple of n_tile. */
int hsv_conv1d_umma_wn_tail(const void *a_blk16, const void *w_packed, const float *bias, float *x, float *output,
                            const float *mask, void *next_blk16, int B, int Cin, int H, int64_t L, int n_tile, void *stream);

/* ---- whole AMP half-layer: Activation1d(SnakeBeta) followed by the dilated Conv1d,
 *   xt = conv(Activation1d(x * in_scale))            (hierspeechpp_speechsynthesizer.py:380-384)
 * in ONE kernel: the CTA evaluates the fused activation on the fp32 input and writes the fp16 operand
 * straight into the shared-memory MMA tile, so the operand never exists in HBM.  Same arithmetic as
 * hsv_act1d_snakebeta(out_mode 1) + hsv_conv1d_umma (bit-identical results).
 *   x [B,Cin,L] fp32, Cin in {16, 32, 64}; w_packed from hsv_pack_conv_weight with n_tile = Cout (<= 256);
 *   bias / residual / out / acc / acc_mode as hsv_conv1d_umma; out and acc must not alias x
 *   (other CTAs still read x's halo); residual may alias out.
 */
int hsv_act_conv1d_umma(const float *x, const float *alpha, const float *beta, float in_scale,
                        const void *w_packed, const float *bias, const float *residual, float *out,
                        float *acc, int acc_mode, int B, int Cin, int Cout, int64_t L, int k, int d,
                        void *stream);

/* ---- ConvTranspose1d (ups[i], hierspeechpp_speechsynthesizer.py:404-408,434) on the same
 * tcgen05 kernel: the u output phases are u stride-1 convolutions over the input rows
 * (SURVEY.md §A.3), each an N-tile group of one launch.
 *   hsv_pack_convT_weight: w [Cin,Cout,k] fp32 (folded) -> fp16 [phase][Cout/n_tile][Cin/CW][taps] blocks
 *   hsv_conv_transpose1d_umma: a_blk16 (rows per chunk = hsv_blk16_rows(Lin)) -> out [B,Cout,u*Lin] fp32,
 *   out = convT + bias (+ add, same shape as out: proj(pitch) at stage 0, :436-438).
 *   stride u <= 8, padding (k-u)/2, k - 2*((k-u)/2) == u (true for every (k,u) on the path).
 */
int hsv_pack_convT_weight(const float *w, void *packed, int Cin, int Cout, int k, int u, int n_tile, void *stream);
int hsv_conv_transpose1d_umma(const void *a_blk16, const void *w_packed, const float *bias, const float *add,
                              float *out, int B, int Cin, int Cout, int64_t Lin, int k, int u, int n_tile,
                              void *stream);

/* ---- generic fp32 Conv1d (stride 1, zero padding `pad`, dilation d) on CUDA
 * cores, for the small/odd-shaped convs of the path: conv_pre, cond, proj,
 * DBlock convs, conv_post (+tanh) (hierspeechpp_speechsynthesizer.py:401,
 * 421-426,430,437,449-450; speechsr.py:106-107).
 *   x [B,Cin,Lin], w [Cout,Cin,k], bias [Cout] or NULL -> out [B,Cout,Lout],
 *   Lout = Lin + 2*pad - d*(k-1).
 *   flags: HSV_CONV_LRELU_IN  apply leaky_relu(0.1) to x first (DBlock :336-337)
 *          HSV_CONV_TANH      tanh on the result (Generator :450)
 *          HSV_CONV_ADD_OUT   out += result instead of out = result
 */
#define HSV_CONV_LRELU_IN 1
#define HSV_CONV_TANH 2
#define HSV_CONV_ADD_OUT 4
#define HSV_CONV_SILU_IN 8      /* apply SiLU to x first (adaLN / cond_block, modules.py:402, hierspeechpp_speechsynthesizer.py:70) */
#define HSV_CONV_LRELU001_IN 16 /* apply leaky_relu(0.01) to x first (PitchPredictor, ttv_v1/t2w2v_transformer.py:456) */
int hsv_conv1d_direct(const float *x, const float *w, const float *bias, float *out,
                      int B, int Cin, int Cout, int64_t Lin, int64_t Lout, int k, int d, int pad,
                      int flags, void *stream);

/* ---- ConvTranspose1d (ups[i], hierspeechpp_speechsynthesizer.py:404-408,434):
 *   x [B,Cin,Lin], w [Cin,Cout,k] (folded), bias [Cout] -> out [B,Cout,u*Lin],
 *   stride u, padding (k-u)/2, output_padding 0.  add [B,Cout,u*Lin] or NULL is
 *   added to the result (proj(pitch) at stage 0, :436-438).
 */
int hsv_conv_transpose1d_direct(const float *x, const float *w, const float *bias, const float *add,
                                float *out, int B, int Cin, int Cout, int64_t Lin, int k, int u,
                                void *stream);

/* ---- SpeechSR front end: conv_pre (Cin=1, k=7, pad 3; speechsr.py:90) followed
 * by F.interpolate(mode='linear', align_corners=False) to Lout (speechsr.py:96),
 * fused.  x [B,1,Lin], w [C,1,7], bias [C] -> out [B,C,Lout].
 * Source index math in fp32 exactly as ATen's CUDA kernel.
 */
int hsv_sr_pre_interp(const float *x, const float *w, const float *bias, float *out,
                      int B, int C, int64_t Lin, int64_t Lout, void *stream);

/* Index probe for the bit-exact part of the parity contract: writes, for each
 * dst in [0,Lout): i0, i1 (int32) and lambda (fp32) of the linear interpolation. */
int hsv_interp_linear_table(int64_t Lin, int64_t Lout, int32_t *i0, int32_t *i1, float *lam, void *stream);

/* ---- nearest-neighbour down-sampling gather of DBlock
 * (F.interpolate(x, size=L//factor), hierspeechpp_speechsynthesizer.py:329-334):
 * out[b,c,t] = x[b,c,min(floor(t*scale), Lin-1)], scale = (float)Lin/Lout. */
int hsv_nearest_gather(const float *x, float *out, int rows, int64_t Lin, int64_t Lout, void *stream);

/* out[i] = a[i] + b[i] + c[row(i)] ... small fused adds used at the stage-0 join
 * (x = conv_pre(x) + downs(pitch) + cond(g), :430): out = a + b + bc[b,c,0]. */
int hsv_add3_bcast(const float *a, const float *b, const float *bc, float *out,
                   int rows, int64_t L, void *stream);

/* fp32 [B,C,L] * in_scale -> fp16 blk16 (optional leaky_relu(0.1) first); C % 16 == 0. */
int hsv_pack_blk16(const float *x, void *out, int B, int C, int64_t L, int lrelu, float in_scale, void *stream);
/* the same with the operand = ((x1 + x2) + x3) * in_scale (x2, x3 nullable): the sum over the resblocks of a stage
 * (hierspeechpp_speechsynthesizer.py:440-446) taken by the consumer, each resblock writing its own tensor. */
int hsv_pack_blk16_sum3(const float *x1, const float *x2, const float *x3, void *out, int B, int C, int64_t L,
                        int lrelu, float in_scale, void *stream);
/* inverse (tests / debugging): fp16 blk16 -> fp32 [B,C,L]. */
int hsv_unpack_blk16(const void *in, float *x, int B, int C, int64_t L, void *stream);
/* operand health (debugging aid; the reference keeps fp32 everywhere, the tensor-core operands here are fp16):
 * stats[0] (uint32) += number of non-finite values in the valid rows of a blk16 buffer (an activation beyond
 * +-65504 saturates to inf when the operand is produced), stats[1] (float bits, non-negative) = max(stats[1],
 * max |finite value|).  The caller zeroes stats[2] first. */
int hsv_blk16_stats(const void *in, void *stats, int B, int C, int64_t L, void *stream);

/* ---- the step after the path: peak-normalise + int16 quantise on the device.
 * Replaces inference_plm.py:183-188 (audio / abs(audio).max() * 32767.0 * s) and
 * inference_speechsr.py:39-41 (audio / abs(audio).max() * 0.999 * 32767.0), each followed by
 * .cpu().numpy().astype('int16'):  out = trunc(((x / peak) * s1) * s2), fp32, operations rounded
 * separately in that order (bit-exact with the reference), saturated to int16.
 *   x [rows, L] fp32, out [rows, L] int16, peak_ws [rows] fp32 workspace (receives the peaks)
 *   per_row 0: peak = max |x| over the whole tensor (the reference's semantics, one utterance);
 *           1: one peak per row (batched utterances are normalised independently).
 */
int hsv_peak_norm_pcm16(const float *x, int16_t *out, float *peak_ws, int rows, int64_t L, float s1, float s2,
                        int per_row, void *stream);

/* ---- frame-rate operators of the step BEFORE the vocoder (csrc/frame_ops.cu).  fp32 [B,C,T] tensors,
 * mask = [B,T] (1 = valid frame, may be NULL), blk16 outputs feed hsv_conv1d_umma.
 *
 * hsv_pack_blk16_act: operand packer with a fused activation.
 *   mode 0: x * mask;  1: tanh(x[c] + bc[c]) * sigmoid(x[C+c] + bc[C+c]) (x [B,2C,T], bc [B,2C] or NULL: WN gate,
 *   commons.py:108-114);  2: gelu_tanh(x) * mask (FFN_Conv, modules.py:382-388);  3: mish(x) * mask (styleencoder.py:6-10)
 *   x_channels = channels per batch item of x (a channel prefix of a wider tensor can be packed).
 * hsv_ln_mod_blk16: LayerNorm over channels (eps, no affine) [* mask] -> x*(1+scale[b,c]) + shift[b,c] (modules.py:346,
 *   405-410); inmask: x*mask first; premask: normalised*mask before the modulation; mod_stride = batch stride of shift/scale.
 * hsv_frame_op: element-wise ops (op codes in csrc/frame_ops.cu): 1 wn_res, 2 wn_last, 3 gate_add, 4 couple, 5 sample,
 *   6 mask, 7 add, 8 glu_res, 9 mish, 10 flip, 11 add_bcast.
 * hsv_mha: softmax(q k^T * scale) v per (batch, head); q/k/v channel-major [heads*D, T] with batch strides (so the fused
 *   qkv tensor of timm's Attention can be addressed in place); lens [B] int32 or NULL = masked_fill(-1e4) of
 *   attentions.py:174-175 for prefix masks; prescale_q: scale q before the product (attentions.py:164) or the scores
 *   after it (timm).
 * hsv_conv1d_c1_strided: Conv1d(1, Cout, k, stride, pad) * mask (PosteriorSFEncoder.pre_filter, :187,196).
 * hsv_masked_mean: out[b,c] = sum_t x[b,c,t] / sum_t mask[b,t] (styleencoder.py:91-99: the sum runs over all frames). */
int hsv_pack_blk16_act(const float *x, const float *bcast, const float *mask, void *out, int B, int C, int64_t L,
                       int mode, int x_channels, void *stream);
/* WN layer tail + the next layer's operand pack in one pass (modules.py:167-174): x = (x + rs[:, :C]) * mask in place,
 * output += rs[:, C:], blk = fp16 blk16 operand of the new x.  rs is [B, 2C, T]. */
int hsv_wn_res_pack(float *x, const float *rs, const float *mask, float *output, void *blk, int B, int C, int64_t L,
                    void *stream);
int hsv_ln_mod_blk16(const float *x, const float *shift, const float *scale, const float *mask, void *out, int B, int C,
                     int64_t L, float eps, int inmask, int premask, int64_t mod_stride, void *stream);
/* hsv_gate_ln_mod_blk16: x = x + gate[b,c] * y * mask (in place, modules.py:408-409) fused with hsv_ln_mod_blk16 of the new x:
 * the gated residual update of a DiT block and the LayerNorm/modulate/pack that follows it, one pass. */
int hsv_gate_ln_mod_blk16(float *x, const float *y, const float *gate, int64_t gate_stride, const float *shift,
                          const float *scale, const float *mask, void *out, int B, int C, int64_t L, float eps, int premask,
                          int64_t mod_stride, void *stream);
int hsv_frame_op(int op, const float *a, const float *b, const float *c, const float *mask, float *out, float *out2,
                 int B, int C, int64_t L, float s, int64_t cstride, void *stream);
int hsv_mha(const float *q, const float *k, const float *v, float *out, const int *lens, int B, int heads, int D, int Tq,
            int Tk, int64_t q_bstride, int64_t k_bstride, int64_t v_bstride, float scale, int prescale_q, void *stream);
/* the same (tensor-core kernel only), the result written as the fp16 blk16 operand [heads*D channels, Tq rows] of the
 * projection conv that follows (timm Attention.proj): saves the fp32 round trip and the pack launch. */
int hsv_mha_blk16(const float *q, const float *k, const float *v, void *out_blk16, const int *lens, int B, int heads, int D,
                  int Tq, int Tk, int64_t q_bstride, int64_t k_bstride, int64_t v_bstride, float scale, int prescale_q,
                  void *stream);
/* test hook: 0 = tensor-core attention (csrc/mha_mma.cu, default), 1 = the fp32 CUDA-core kernel. Process-global. */
int hsv_set_mha_variant(int v);
int hsv_conv1d_c1_strided(const float *x, const float *w, const float *bias, const float *mask, float *out, int B,
                          int Cout, int64_t Lin, int64_t Lout, int k, int stride, int pad, void *stream);
int hsv_masked_mean(const float *x, const float *mask, float *out, int B, int C, int64_t L, void *stream);

/* ---- f0-driven harmonic sine source (BASELINE.json north_star item 3; NSF / HiFTNet-style SineGen -- the
 * reference repository has no counterpart on this path).  The phase accumulator is 64-bit fixed point
 * (2^64 = one cycle): the cumulative sum is exact integer arithmetic, so a 30 s utterance does not drift.
 *   f0        [B, T] fp32, Hz per frame (<= 0: unvoiced, the phase holds)
 *   out       [B, harmonics, T*hop] fp32 = amp * voiced * sin(2 pi (h+1) phase[n]),  phase[n] = sum_{k<=n} f0[k/hop]/sr
 *   uv        [B, T*hop] fp32 voiced mask (may be NULL)
 *   workspace hsv_sinegen_workspace(B, T, hop) bytes
 */
int64_t hsv_sinegen_workspace(int B, int64_t T, int hop);
int hsv_sinegen(const float *f0, float *out, float *uv, void *workspace, int B, int64_t T, int hop,
                float sample_rate, int harmonics, float amp, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HSV_H_ */
