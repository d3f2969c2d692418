/*
 * sg2im_b200 — C-ABI of libsg2im_b200.so (CUDA, sm_100a only).
 *
 * The reference (google/sg2im) has no FFI layer: its hot path is Python over
 * torch ops.  Each entry point below replaces the torch op sequence at the
 * cited reference location (paths relative to the reference tree).  The host
 * mirror in sg2im_b200/*.py binds these with ctypes; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless stated; fp32 values, int64 indices
 *    (the reference's LongTensors), int32 for library-built CSR tables;
 *  - activations are NHWC ("pixels x channels" row-major); a tensor argument
 *    followed by (cstride, coff) is a channel slice [coff, coff+C) of a wider
 *    NHWC buffer whose pixel stride is cstride floats (virtual concat);
 *  - all work is enqueued on `stream` (a cudaStream_t); nothing synchronises,
 *    nothing allocates; buffers (incl. zero-initialised accumulators where
 *    stated) are owned by the caller;
 *  - return 0 on success, <0 invalid argument / unsupported shape, >0 a
 *    cudaError_t from the launch; sg2im_last_error_string() describes the last
 *    failure on the calling thread.  There is no CPU fallback.
 */
#ifndef SG2IM_B200_H_
#define SG2IM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sg2im_stream_t;           /* cudaStream_t */

int sg2im_abi_version(void);
const char* sg2im_last_error_string(void);
/* 1 if the running device is compute capability 10.x, else 0 (no dispatch: the
 * library refuses to run elsewhere). */
int sg2im_device_ok(void);

/* ---------------------------------------------------------------- graph --
 * CSR of "which (triple, role) entries touch row r", entries ordered
 * role 0 ascending t, then role 1 ascending t — the order the reference's two
 * scatter_add calls impose (sg2im/graph.py:98-99).  idx[t*idx_stride + role]
 * is the destination row.  nroles in {1,2}.  Also used with nroles=1 for
 * obj_to_img (sg2im/layout.py:146-148).  row_ptr: int32[R+1]; entries:
 * int32[nroles*T], entry = t*2 + role.  Out-of-range indices are dropped. */
int sg2im_csr_build(const int64_t* idx, int64_t T, int64_t idx_stride, int nroles,
                    int64_t R, int32_t* row_ptr, int32_t* entries,
                    sg2im_stream_t stream);

/* out[t] = [ rows[idx[t,0]] * f0 | mid[t] | rows[idx[t,1]] * f1 ], widths
 * (Wr, Wm, Wr); f = 1/max(count,1) per row when row_ptr != NULL else 1.
 * Forward use: cur_t_vecs = cat(obj[s], pred, obj[o]) (sg2im/graph.py:77-82).
 * Backward use: gradient of the avg-pool scatter (graph.py:98-114), mid = grad
 * of the predicate slice (NULL => zeros). */
int sg2im_triple_gather(const float* rows, const float* mid, const int64_t* edges,
                        int64_t T, int64_t Wr, int64_t Wm, const int32_t* row_ptr,
                        float* out, sg2im_stream_t stream);

/* out[r, 0:W] = sum over CSR entries (t, role) of row r, in CSR order, of
 * src[t, off_role : off_role+W]   (/ max(count,1) if avg).  Sequential fp32
 * adds in CSR order => bit-exact vs the CPU scatter_add (graph.py:92-114).
 * Backward use: gradient of the gather (graph.py:77-78). */
int sg2im_segment_sum(const float* src, int64_t src_stride, int64_t off0, int64_t off1,
                      int64_t W, const int32_t* row_ptr, const int32_t* entries,
                      int64_t R, int avg, float* out, sg2im_stream_t stream);

/* ----------------------------------------------------------------- conv --
 * Implicit-GEMM convolution, NHWC, fp32 FFMA (exact fp32) — replaces
 * nn.Conv2d / nn.Linear (sg2im/crn.py:41-45,80-82, model.py:100,105,
 * layers.py:178,221, discriminators.py:62-66).
 *
 * mode 0 (forward):  y[n,oy,ox,co] = b[co] + sum_{ky,kx,ci}
 *      x[n, oy*S-P+ky, ox*S-P+kx, ci] * w[(ky*KW+kx)*Cin + ci][co]
 * mode 1 (data gradient): x plays dY (N,Hin,Win,Cin=Cout_fwd), y plays dX
 *      (N,Hout,Wout,Cout=Cin_fwd): y[n,iy,ix,c] = sum_{ky,kx,co: (iy+P-ky)%S==0..}
 *      x[n,(iy+P-ky)/S,(ix+P-kx)/S,co] * w[(ky*KW+kx)*Cin + co][c]
 * w is the packed (KH*KW*Cin) x Cout row-major matrix for the given mode.
 * x is addressed with element strides (sxn,sxh,sxw,sxc) so NCHW callers need
 * no copy.  Epilogue: + bias (may be NULL), act: 0 none, 1 leaky(slope)
 * (slope 0 = ReLU).  Output written to the slice (y, y_cstride, y_coff).
 * A Linear is KH=KW=1, Hin=Win=1, N=rows. */
int sg2im_conv_igemm(int mode, const float* x, int64_t sxn, int64_t sxh, int64_t sxw,
                     int64_t sxc, int64_t N, int64_t Hin, int64_t Win, int64_t Cin,
                     const float* w, const float* bias, int KH, int KW, int S, int P,
                     int64_t Hout, int64_t Wout, int64_t Cout, int act, float slope,
                     float* y, int64_t y_cstride, int64_t y_coff,
                     sg2im_stream_t stream);

/* Tensor-core path (tcgen05.mma kind::tf32, TMA-staged NHWC tiles, TMEM
 * accumulators): stride-1 convolution / Linear, TF32 multiply, fp32 accumulate.
 *   y[n,oy,ox,co] = act(b[co] + sum_{ky,kx,ci} x[n,oy-P+ky,ox-P+kx,ci] *
 *                                               w_tc[ky*KW+kx][co][ci])
 * for oy < Hout, ox < Wout (Hout/Wout may be smaller than the natural output
 * size).  x is NHWC with pixel stride x_cstride floats (first Cin channels
 * used); w_tc is packed [KH*KW][Cout][Cin].  The data gradient of a stride-1
 * conv is the same call with spatially flipped, channel-transposed weights
 * and P' = K-1-P.  A KxK stride-2 'valid' conv (the discriminators) is the
 * same call on the space-to-depth input (sg2im_s2d_fwd) with K/2 taps.
 * stats (optional, act == 0, Cout <= 1024): caller-zeroed double[2*Cout]; the
 * epilogue adds the per-channel sum and sum of squares of the outputs — the
 * batch statistics of the BatchNorm that follows, without a second pass.
 * round_out: write the outputs rounded to nearest TF32 (for outputs that feed another
 * tensor-core op: the hardware truncates its fp32 operands, which biases products toward
 * zero; RN-rounded operands are consumed exactly).
 * math: the tensor-core arithmetic (same HBM traffic in all three; activations and weights stay
 * fp32 in HBM):
 *   SG2IM_MATH_TF32    kind::tf32 on the fp32 words as they are (2^-11 operand precision);
 *   SG2IM_MATH_BF16X3  converter warps split every landed shared-memory tile into bf16 hi / mid
 *                      halves and every fp32 product is issued as hi*hi + mid*hi + hi*mid on
 *                      kind::f16 MMAs with fp32 accumulation (2^-17 operand precision: the mode
 *                      that meets the 1e-3 bar against the fp32 reference,
 *                      scripts/train.py:423), 1.5x the tensor-pipe time of TF32;
 *   SG2IM_MATH_BF16    the same kernels issuing hi*hi only (plain bf16 operands, BASELINE.json
 *                      configs[3]); round_out is ignored outside SG2IM_MATH_TF32.
 * sg2im_conv_tc_supported(): S == 1, Cin % 4 == 0, Cout % 4 == 0, 16-byte
 * aligned slices; otherwise sg2im_conv_tc returns -2 (use sg2im_conv_igemm). */
#define SG2IM_MATH_TF32 0
#define SG2IM_MATH_BF16X3 1
#define SG2IM_MATH_BF16 2
int sg2im_conv_tc_supported(int64_t N, int64_t Hin, int64_t Win, int64_t Cin,
                            int64_t x_cstride, int KH, int KW, int S, int P,
                            int64_t Hout, int64_t Wout, int64_t Cout, int64_t y_cstride,
                            int64_t y_coff);
int sg2im_conv_tc(const float* x, int64_t x_cstride, int64_t N, int64_t Hin, int64_t Win,
                  int64_t Cin, const float* w_tc, const float* bias, int KH, int KW, int P,
                  int64_t Hout, int64_t Wout, int64_t Cout, int act, float slope, float* y,
                  int64_t y_cstride, int64_t y_coff, double* stats, int round_out, int math,
                  sg2im_stream_t stream);

/* The same convolution with the weights taken straight from the WEIGHT-GRADIENT layout
 * w_kcc[tap][row][col] (rows = the conv weight's input channels, cols = its output channels —
 * what sg2im_conv_wgrad_tc writes), so that master weights kept in that layout need no
 * pack / unpack pass:
 *   dgrad == 0  forward: Cin rows used of w_rows_full, Cout == cols (B operand MN-major);
 *   dgrad != 0  data gradient of that conv: call with x = dY, Cin = cols, Cout = rows used,
 *               P' = K-1-P; the kernel flips the tap index itself.
 * w_rows_full = row pitch of one tap (>= rows used: the CRN's first stage uses a channel prefix).
 * Validated on the B200 (round 2): what TrainStep(weights='kcc') runs. */
int sg2im_conv_tc_kcc(const float* x, int64_t x_cstride, int64_t N, int64_t Hin, int64_t Win,
                      int64_t Cin, const float* w_kcc, int64_t w_rows_full, int dgrad,
                      const float* bias, int KH, int KW, int P, int64_t Hout, int64_t Wout,
                      int64_t Cout, int act, float slope, float* y, int64_t y_cstride,
                      int64_t y_coff, double* stats, int round_out, int math,
                      sg2im_stream_t stream);

/* dw[(ky*KW+kx)*Cin + ci][co] += sum_{n,oy,ox} dy[n,oy,ox,co] *
 *      x[n, oy*S-P+ky, ox*S-P+kx, ci]      (dw must be zero-initialised: the
 * reduction over pixels is split across CTAs and combined with fp32 atomics). */
int sg2im_conv_wgrad(const float* x, int64_t sxn, int64_t sxh, int64_t sxw, int64_t sxc,
                     int64_t N, int64_t Hin, int64_t Win, int64_t Cin,
                     const float* dy, int KH, int KW, int S, int P,
                     int64_t Hout, int64_t Wout, int64_t Cout,
                     float* dw, sg2im_stream_t stream);

/* Tensor-core weight gradient of a stride-1 convolution or a Linear
 * (tcgen05.mma kind::tf32 or kind::f16 on in-kernel bf16 pairs — `math` as in sg2im_conv_tc —,
 * MN-major operands straight from NHWC, one smem halo
 * tile serves all taps): dw[(ky*KW+kx)*Cin + ci][co] += sum_pix
 * x[pix+tap-P, ci] * dy[pix, co] over the Hout x Wout outputs; dw
 * zero-initialised by the caller (partial tiles are combined with vector
 * atomics).  Supported when S == 1, K <= 3, Cin % 4 == 0, Cout % 32 == 0 (and
 * N*H*W % 32 == 0 for K == 1); else -2 -> use sg2im_conv_wgrad. */
int sg2im_conv_wgrad_tc_supported(int64_t N, int64_t Hin, int64_t Win, int64_t Cin,
                                  int64_t x_cstride, int KH, int KW, int S, int P,
                                  int64_t Hout, int64_t Wout, int64_t Cout);
int sg2im_conv_wgrad_tc(const float* x, int64_t x_cstride, int64_t N, int64_t Hin, int64_t Win,
                        int64_t Cin, const float* dy, int KH, int KW, int P, int64_t Hout,
                        int64_t Wout, int64_t Cout, float* dw, int math, int64_t s2d_channels,
                        sg2im_stream_t stream);
/* s2d_channels = C > 0 (2x2 taps, Cin = 4C, x = the space-to-depth form of the input of a 4x4 stride-2
 * convolution): dw is THAT filter's gradient [16][C][Cout] — the kernel maps tap (ty, tx), channel
 * (py, px, c) to filter tap (2 ty + py, 2 tx + px), channel c — so that the discriminators' weight
 * gradients land directly in the parameter's slot of the gradient bucket.  0 = plain. */

/* Pre-split weight operands of the bf16 arithmetic (SG2IM_MATH_BF16X3 / SG2IM_MATH_BF16): all
 * convolution / Linear weights of a network in ONE launch.  `table`: n_entries x 8 int64 in device
 * memory — {src, fwd, dgrad, taps, Cin, Cout, first_tile, 0} per weight — with src the fp32 master
 * stored [taps][Cin][Cout] (the weight-gradient layout), fwd / dgrad (either may be 0) the operand
 * copies [taps][Cout][cin_pad] and [taps, flipped][Cin][cout_pad] (pads = channels rounded up to 32)
 * whose rows are 32-channel blocks of [32 x bf16 hi | 32 x bf16 mid], and first_tile the running sum
 * of taps * ceil(Cin/32) * ceil(Cout/32) over the preceding entries (total_tiles = the full sum).
 * The 8th field (0 above) = C > 0 marks a 4x4 stride-2 filter stored [16][C][Cout]: its copies are
 * those of the equivalent 2x2 stride-1 filter on the space-to-depth input (taps = 4, Cin = 4C).
 * sg2im_conv_tc_presplit consumes them: the kernels then split only the activation tiles. */
int sg2im_split_weights(const int64_t* table, int64_t n_entries, int64_t total_tiles,
                        sg2im_stream_t stream);
/* sg2im_conv_tc with a pre-split B operand: w_split = rows of w_pitch floats (w_pitch % 32 == 0,
 * >= Cin rounded up to 32), w_rows_per_tap rows per tap (>= Cout: the data gradient of a channel
 * prefix uses the first Cout rows).  math: SG2IM_MATH_BF16X3 or SG2IM_MATH_BF16. */
int sg2im_conv_tc_presplit(const float* x, int64_t x_cstride, int64_t N, int64_t Hin, int64_t Win,
                           int64_t Cin, const float* w_split, int64_t w_pitch, int64_t w_rows_per_tap,
                           const float* bias, int KH, int KW, int P, int64_t Hout, int64_t Wout,
                           int64_t Cout, int act, float slope, float* y, int64_t y_cstride,
                           int64_t y_coff, double* stats, int math, sg2im_stream_t stream);

/* Mean BCE-with-logits against a constant target t (0 or 1): the GAN losses bce_loss(scores, ones /
 * zeros) of sg2im/losses.py:39-57,60-103 in one pass forward (scratch: one caller-zeroed double) and
 * one backward (gout: the upstream scalar gradient on the device), instead of ~20 elementwise ATen
 * kernels per evaluation. */
int sg2im_bce_logits_mean_fwd(const float* x, int64_t n, float target, double* scratch, float* out,
                              sg2im_stream_t stream);
int sg2im_bce_logits_mean_bwd(const float* x, int64_t n, float target, const float* gout, float* dx,
                              sg2im_stream_t stream);

/* Space-to-depth by 2 (and its adjoint): out[n, y/2, x/2, ((y&1)*2+(x&1))*C + c]
 * = x[n,y,x,c], zero padded to even H, W.  x addressed with element strides.
 * Turns the discriminators' 4x4 stride-2 'valid' convolutions
 * (sg2im/layers.py:164-181, scripts/train.py:122-130) into 2x2 stride-1 ones. */
int sg2im_s2d_fwd(const float* x, int64_t sxn, int64_t sxh, int64_t sxw, int64_t sxc,
                  int64_t N, int64_t H, int64_t W, int64_t C, float* out, sg2im_stream_t stream);
int sg2im_s2d_bwd(const float* dout, int64_t N, int64_t H, int64_t W, int64_t C, float* dx,
                  sg2im_stream_t stream);

/* OIHW master weights <-> kernel layouts, smem-tiled (both sides coalesced).
 * pack:   w[co][ci][t] (ci < cin_use of Cin) -> w_fwd[t][co][ci] (sg2im_conv_tc
 *         forward) and/or w_dgrad[T-1-t][ci][co] (sg2im_conv_tc as data
 *         gradient: flipped taps, channels swapped); either may be NULL.
 * unpack: dw[t][ci][co] (sg2im_conv_wgrad[_tc] output) -> grad_oihw[co][ci][t]
 *         (= or += when accumulate). */
int sg2im_pack_weights(const float* w, int64_t Cout, int64_t Cin, int64_t cin_use, int64_t taps,
                       float* w_fwd, float* w_dgrad, int round_tf32, sg2im_stream_t stream);
int sg2im_unpack_wgrad(const float* dw, int64_t Cout, int64_t Cin, int64_t cin_use, int64_t taps,
                       float* grad_oihw, int accumulate, sg2im_stream_t stream);

/* out[c] = sum_m x[m, c]  (bias gradient), fp64 accumulation; out zeroed by
 * the call. */
int sg2im_colsum(const float* x, int64_t M, int64_t C, float* out,
                 double* scratch /* C doubles */, sg2im_stream_t stream);

/* dx = dy * (y > 0 ? 1 : slope)   (in place allowed) — backward of a fused
 * epilogue activation, keyed on the activation OUTPUT. */
int sg2im_act_bwd(const float* dy, const float* y, float slope, int64_t n,
                  float* dx, sg2im_stream_t stream);

/* -------------------------------------------------------- batch norm etc --
 * Train-mode nn.BatchNorm2d (+ LeakyReLU, + nearest x2 upsample, + channel
 * concat) — sg2im/crn.py:43-47,63,107; layers.py:22-46; model.py:98-99.
 *
 * bn_stats:    sums[c] += sum_m x[m,c], sums[C+c] += sum_m x[m,c]^2 (fp64,
 *              caller zeroes `sums`, 2C doubles).
 * bn_finalize: mean/var from sums over `count` rows; scale = gamma*invstd,
 *              shift = beta - mean*scale; running stats updated with momentum
 *              and the unbiased factor n/(n-1), n = count*unbias_mult
 *              (unbias_mult = 4 when BN follows a x2 upsample: same mean/var,
 *              4x the elements).  training=0: scale/shift from running stats.
 *              save[0:C]=mean, save[C:2C]=invstd.  gamma/beta NULL => 1/0.
 * scale_act_fwd: y[n,Y,X,coff+c] = leaky(x[n,Y/up,X/up,c]*scale[c]+shift[c])
 *              (rounded to nearest TF32 when round_tf32: the consumer is a tensor-core conv);
 *              scale NULL => identity affine; slope 1 => no activation.
 */
int sg2im_bn_stats(const float* x, int64_t M, int64_t C, double* sums,
                   sg2im_stream_t stream);
int sg2im_bn_finalize(const double* sums, int64_t count, int64_t unbias_mult, int64_t C,
                      const float* gamma, const float* beta, float eps, float momentum,
                      int training, float* running_mean, float* running_var,
                      float* scale, float* shift, float* save, int64_t* num_batches_tracked,
                      sg2im_stream_t stream);   /* num_batches_tracked (may be NULL): +1 in training mode */
int sg2im_scale_act_fwd(const float* x, int64_t N, int64_t H, int64_t W, int64_t C,
                        const float* scale, const float* shift, float slope, int up,
                        float* y, int64_t y_cstride, int64_t y_coff, int round_tf32,
                        sg2im_stream_t stream);
/* Backward of scale_act_fwd given dy on the (upsampled, sliced) output.
 * reduce: with g = leaky'(x*scale+shift) * sum_{up x up} dy,
 *         sums[c] += sum g, sums[C+c] += sum g*xhat, xhat = (x-mean)*invstd
 *         (fp64; caller zeroes).
 * apply:  training: dx = scale*(g - sums[c]/M - xhat*sums[C+c]/M)
 *         else:     dx = scale*g ;   dgamma = sums[C+c], dbeta = sums[c]
 *         (dgamma/dbeta NULL allowed; written, not accumulated). */
int sg2im_scale_act_bwd_reduce(const float* dy, int64_t dy_cstride, int64_t dy_coff,
                               const float* x, int64_t N, int64_t H, int64_t W, int64_t C,
                               const float* scale, const float* shift, const float* save,
                               float slope, int up, double* sums, sg2im_stream_t stream);
/* (sg2im_scale_act_bwd_apply: `training` bit 0 = batch statistics were used; bit 1 = ADD the
 * parameter gradients into dgamma / dbeta — the parameters' slots of a zeroed gradient bucket —
 * instead of overwriting them.) */
int sg2im_scale_act_bwd_apply(const float* dy, int64_t dy_cstride, int64_t dy_coff,
                              const float* x, int64_t N, int64_t H, int64_t W, int64_t C,
                              const float* scale, const float* shift, const float* save,
                              float slope, int up, int training, const double* sums,
                              float* dx, float* dgamma, float* dbeta,
                              sg2im_stream_t stream);

/* 2x2 average pool between channel slices (F.avg_pool2d, sg2im/crn.py:58-62,
 * applied as a factor-2 cascade) and its backward:
 * dfine[n,Y,X,c] (+)= dcoarse[n,Y/2,X/2,c] / 4  (accumulate=1 adds). */
int sg2im_avgpool2_fwd(const float* x, int64_t x_cstride, int64_t x_coff,
                       int64_t N, int64_t H, int64_t W, int64_t C,
                       float* y, int64_t y_cstride, int64_t y_coff, sg2im_stream_t stream);
int sg2im_avgpool2_bwd(const float* dcoarse, int64_t dc_cstride, int64_t dc_coff,
                       int64_t N, int64_t H, int64_t W, int64_t C,
                       float* dfine, int64_t df_cstride, int64_t df_coff, int accumulate,
                       sg2im_stream_t stream);

/* Non-overlapping pooling, NHWC contiguous: nn.MaxPool2d / nn.AvgPool2d(kernel_size = stride =
 * factor) as built by build_cnn's 'PX' token (sg2im/layers.py:195-201).  mode 0 = average,
 * 1 = max (ties: first element in row-major window order, like ATen).  y is
 * (N, H/factor, W/factor, C) (floor: trailing rows / columns are dropped); the backward writes
 * every element of the covered region of dx (N,H,W,C) — the caller zeroes dx when H or W is not a
 * multiple of factor.  x is read only for mode 1 (arg-max re-derived). */
int sg2im_pool2d_fwd(const float* x, int64_t N, int64_t H, int64_t W, int64_t C, int factor,
                     int mode, float* y, sg2im_stream_t stream);
int sg2im_pool2d_bwd(const float* dy, const float* x, int64_t N, int64_t H, int64_t W, int64_t C,
                     int factor, int mode, float* dx, sg2im_stream_t stream);

/* --------------------------------------------------------------- layout --
 * Fused masks_to_layout / boxes_to_layout (sg2im/layout.py:30-162):
 *   out[n,h,w,d] = sum_{o in image n, ascending o} vecs[o,d] * S_o(h,w),
 *   S_o = bilinear sample (zeros padding, align_corners flag) of masks[o]
 *   (MxM) at the box-local coordinate of pixel (h,w).  masks NULL => the
 *   constant 8x8 ones "mask" of boxes_to_layout.  img_row_ptr/img_entries:
 *   CSR image -> objects from sg2im_csr_build(obj_to_img, nroles=1).
 * The (O,D,H,W) temporary of the reference never exists.  Channels
 * [D, D+noise_c) of the output slice are filled from `noise` addressed with
 * element strides (n,c,h,w) (model.py:164-169); noise NULL => untouched. */
int sg2im_layout_fwd(const float* vecs, const float* boxes, const float* masks, int64_t M,
                     const int32_t* img_row_ptr, const int32_t* img_entries,
                     int64_t N, int64_t O, int64_t D, int64_t H, int64_t W,
                     int align_corners,
                     const float* noise, int64_t noise_c, int64_t nsn, int64_t nsc,
                     int64_t nsh, int64_t nsw,
                     float* out, int64_t out_cstride, int round_tf32, sg2im_stream_t stream);
/* dvecs[o,d] += sum_hw dout[n,h,w,d]*S_o(h,w);  dmasks[o,my,mx] += bilinear
 * scatter of dS_o(h,w) = sum_d dout[n,h,w,d]*vecs[o,d].  Both zero-initialised
 * by the caller; dmasks may be NULL. */
int sg2im_layout_bwd(const float* dout, int64_t dout_cstride,
                     const float* vecs, const float* boxes, const float* masks, int64_t M,
                     const int64_t* obj_to_img, int64_t N, int64_t O, int64_t D,
                     int64_t H, int64_t W, int align_corners,
                     float* dvecs, float* dmasks, sg2im_stream_t stream);

/* dboxes[o,0:4] = d(loss)/d(x0,y0,x1,y1) of object o through the sampling grid
 * (sg2im/layout.py:94-128 + F.grid_sample's grid gradient); only reached when the generator
 * trains on its predicted boxes (sg2im/model.py:151-160, no boxes_gt).  Every entry of dboxes
 * is written.  masks NULL => the constant 8x8 ones image of boxes_to_layout. */
int sg2im_layout_bwd_boxes(const float* dout, int64_t dout_cstride,
                           const float* vecs, const float* boxes, const float* masks, int64_t M,
                           const int64_t* obj_to_img, int64_t N, int64_t O, int64_t D,
                           int64_t H, int64_t W, int align_corners, float* dboxes,
                           sg2im_stream_t stream);

/* ----------------------------------------------------------------- crop --
 * crop_bbox_batch (sg2im/bilinear.py:28-132,249-278): out[b,i,j,c] = bilinear
 * sample (zeros padding) of feats[idx[b]] at X_j = (1-a_j)(2x0-1)+a_j(2x1-1),
 * a = linspace(0,1,WW), Y_i likewise.  feats addressed with element strides;
 * out NHWC (B,HH,WW,C).  No host sync, any object order. */
int sg2im_crop_fwd(const float* feats, int64_t sfn, int64_t sfh, int64_t sfw, int64_t sfc,
                   int64_t N, int64_t H, int64_t W, int64_t C,
                   const float* boxes, const int64_t* idx, int64_t B, int64_t HH, int64_t WW,
                   int align_corners, float* out, sg2im_stream_t stream);
/* dfeats (N,H,W,C NHWC contiguous, zero-initialised) += scatter of dout. */
int sg2im_crop_bwd(const float* dout, const float* boxes, const int64_t* idx,
                   int64_t N, int64_t H, int64_t W, int64_t C, int64_t B, int64_t HH,
                   int64_t WW, int align_corners, float* dfeats, sg2im_stream_t stream);

/* ------------------------------------------------------------ deprocess --
 * imagenet_deprocess_batch (sg2im/data/utils.py:32-67; callers
 * scripts/train.py:365-366, scripts/run_model.py:70): per element
 * v = x/inv_std[c] - neg_mean[c]; if rescale, v = (v-lo)/(hi-lo) with lo/hi the
 * min/max of v over the whole image n; byte = trunc(clamp(255 v, 0, 255)).
 * imgs and out are addressed with ELEMENT strides (n, c, h, w), so the NCHW
 * view of an NHWC buffer and either output order work.  inv_std, neg_mean:
 * device float[C].  minmax: device scratch uint32[2N] (rescale only; written
 * by the call).  fp32 IEEE arithmetic in the reference's order: identical bytes. */
int sg2im_deprocess(const float* imgs, int64_t sn, int64_t sc, int64_t sh, int64_t sw,
                    int64_t N, int64_t C, int64_t H, int64_t W,
                    const float* inv_std, const float* neg_mean, int rescale,
                    uint32_t* minmax, uint8_t* out, int64_t on, int64_t oc, int64_t oh,
                    int64_t ow, sg2im_stream_t stream);

/* ------------------------------------------------------------ relations --
 * COCO scene-graph synthesis for a collated batch (sg2im/data/coco.py:294-356 per sample +
 * the index offsets of coco_collate_fn, coco.py:400-407).  Objects are grouped by image with the
 * __image__ object last in every image: obj_off[N+1] / trip_off[N+1] are the per-image object and
 * triple offsets, obj_to_img[O] the image of every object.  Per real object i of an image with at
 * least two real objects (trip_off says so: the image owns 2*n_real triples, else n_real):
 * partner[i] = the GLOBAL index of the randomly chosen other object, swap[i] = 0 -> (i, p, partner),
 * 1 -> (partner, p, i) — the host draws both from Python's `random` in the reference's call order.
 * p from the boxes (surrounding / inside, strict inequalities) else from the sector of the
 * difference of the masked centroids (mask value 1; empty mask -> box centre).
 * pred_ids: HOST int64[7] = ids of left of, right of, above, below, inside, surrounding,
 * __in_image__.  centers: device scratch float[2*O] (written).  Outputs triples (T,3) int64 with
 * global object indices and triple_to_img (T): per image first the geometric triples in object
 * order, then (i, __in_image__, image object) for every real object. */
int sg2im_coco_relations(const float* boxes, const int64_t* masks, int64_t MH, int64_t MW,
                         const int64_t* obj_off, const int64_t* trip_off, const int64_t* obj_to_img,
                         const int64_t* partner, const uint8_t* swap, int64_t O,
                         const int64_t* pred_ids, float* centers, int64_t* triples,
                         int64_t* triple_to_img, sg2im_stream_t stream);

/* ------------------------------------------------------------ optimiser --
 * torch.optim.Adam (scripts/train.py:426,436,443; steps at :560,579,592) over
 * one flat fp32 bucket: params/grads/exp_avg/exp_avg_sq are four arrays of n
 * floats (16-byte aligned).  `step`: device float holding the step count,
 * incremented by the call; `found_inf` (device float, may be NULL): nonzero
 * skips the update AND the increment (the collective non-finite-loss skip of
 * train.py:552-555 inside a CUDA graph).  amsgrad=False, maximize=False,
 * weight_decay = L2 added to the gradient.  rounded_out (may be NULL): also write the updated
 * parameters rounded to nearest TF32 — the copy the tensor-core kernels read when the weights
 * live in the weight-gradient layout (the hardware would truncate the fp32 master). */
int sg2im_adam_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                    int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                    float* step, const float* found_inf, float* rounded_out,
                    float grad_scale, sg2im_stream_t stream);
/* dx = dy * leaky'(y) and db[c] += sum_m dx[m,c] in one pass (the activation backward of a
 * conv+bias+LeakyReLU epilogue and its bias gradient, layers.py:39 / crn.py:43-47; autograd
 * derives both in the reference).  db is accumulated into: zeroed buffer or gradient slot.
 * C % 4 == 0, 16-byte aligned. */
int sg2im_act_bwd_colsum(const float* dy, const float* y, float slope, int64_t M, int64_t C,
                         float* dx, float* db, sg2im_stream_t stream);
/* y[i] = x[i] rounded to nearest TF32 (initialises that copy). */
int sg2im_round_tf32(const float* x, int64_t n, float* y, sg2im_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif  /* SG2IM_B200_H_ */
