/* dlsg_b200 — C ABI of the B200-native hot path of dl-cs-dynamic (tjtiger86/dl-swin-gan).
 *
 * The reference has no FFI: its seam is Python duck-typing.  These
 * entry points are what a ctypes/cffi binding of that seam calls; the Python mirror
 * of the reference classes lives in dl_swin_gan_b200/mri/transforms.py (SenseModel,
 * FFT) and dl_swin_gan_b200/swin/window_attention.py (WindowAttention3D).
 *
 * Conventions for every function
 *   - all pointers are DEVICE pointers owned by the caller (outputs and workspace too);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream) and the call returns without synchronising;
 *   - tensors are contiguous, row-major, X fastest; complex64 = interleaved fp32;
 *   - return value: DLSG_OK (0) or a negative DLSG_E* code; nothing is written on error;
 *   - the library keeps no mutable global state: it is re-entrant and honours the
 *     calling thread's current CUDA device (nn.DataParallel threads, reconstruct.py:199).
 */
#ifndef DLSG_B200_H
#define DLSG_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLSG_OK 0
#define DLSG_EINVAL (-1)      /* null pointer, non-positive extent, ndims not in {1,2} */
#define DLSG_EWORKSPACE (-2)  /* workspace smaller than dlsg_sense_workspace_bytes(..., 0) */
#define DLSG_EUNSUPPORTED (-3)/* axis length beyond the direct-DFT limits (X > 4096 or Y > 1536 when not 32*R) */
#define DLSG_ELAUNCH (-4)     /* cudaGetLastError() != cudaSuccess after a launch */

/* Library version (major*10000 + minor*100 + patch). */
int dlsg_version(void);

/* Kernels this library has launched so far in this process (monotonic atomic counter; statistics only —
 * bench.py reports the difference over its timed region as gpu_launches). */
unsigned long long dlsg_launch_count(void);

/* Human-readable text for a status code (static storage). */
const char* dlsg_status_string(int status);

/* Scratch the SENSE entry points need for a problem of this shape.
 *   budget_bytes == 0 : the minimum (one (slice, frame) set of coil images:  C*Y*X*8; 0 when ndims == 1
 *                       except for dlsg_sense_normal_c64 / dlsg_sense_dc_step_c64, which need it in 1-D too);
 *   budget_bytes  > 0 : the largest useful size not above max(budget, minimum).  The row-transformed
 *                       intermediate of one chunk lives here; keeping it a fraction of the 126 MB L2
 *                       keeps it on chip between the row and the column pass.
 */
size_t dlsg_sense_workspace_bytes(int B, int E, int C, int T, int Y, int X, int ndims, size_t budget_bytes);

/* SENSE forward operator, replaces SenseModel._forward_op (dl_cs/mri/transforms.py:92-98):
 *     k[b,c,t] = W[b,c,t] * FFT_ortho( sum_e maps[b,e,c] * x[b,e,t] )
 *   x     [B,E,T,Y,X] complex64 (in)      maps [B,E,C,Y,X] complex64 (in; the reference's
 *   k_out [B,C,T,Y,X] complex64 (out)          [B,E,C,1,Y,X] with the singleton dropped)
 *   weights : real fp32 k-space weights or NULL (the reference's `weights=None -> 1.0`,
 *             transforms.py:79-82); addressed as weights[b*ws[0]+c*ws[1]+t*ws[2]+y*ws[3]+x*ws[4]]
 *             with ws = w_strides (element strides, 0 on broadcast axes; any expand()ed view works).
 *   ndims   : 2 = FFT over (Y,X); 1 = FFT over X only (what the reference derives from a
 *             maps tensor whose squeezed image is 1-D, transforms.py:76).
 *   maps    : may be NULL = unit sensitivities; with E = C = 1 the call is then the plain ortho
 *             (i)FFT of FFT.forward (transforms.py:28-46) over a batch of B*T images.
 */
int dlsg_sense_forward_c64(const void* x, const void* maps, const float* weights, const long long* w_strides,
                           void* k_out, void* workspace, size_t workspace_bytes,
                           int B, int E, int C, int T, int Y, int X, int ndims, void* stream);

/* SENSE adjoint operator, replaces SenseModel._adjoint_op (dl_cs/mri/transforms.py:84-90):
 *     x[b,e,t] = sum_c conj(maps[b,e,c]) * IFFT_ortho( W[b,c,t] * k[b,c,t] )
 */
int dlsg_sense_adjoint_c64(const void* k, const void* maps, const float* weights, const long long* w_strides,
                           void* x_out, void* workspace, size_t workspace_bytes,
                           int B, int E, int C, int T, int Y, int X, int ndims, void* stream);

/* Fused normal operator  x_out = A^H(A(x))  — the core of the data-consistency update
 * A(A(x), adjoint=True) at dl_cs/models/unrolledswin.py:109 and unrolled.py:104.
 * k-space never leaves the chip's L2: the weights are applied squared between the two column passes.
 * x_out must not alias x. */
int dlsg_sense_normal_c64(const void* x, const void* maps, const float* weights, const long long* w_strides,
                          void* x_out, void* workspace, size_t workspace_bytes,
                          int B, int E, int C, int T, int Y, int X, int ndims, void* stream);

/* Regularised normal operator  x_out = A^H(A(x)) + lamda * x  — `model_normal` of the MoDL / half-quadratic-splitting solver
 * (dl_cs/models/unrolledswin.py:154, dl_cs/models/unrolled.py:151), the operator ConjugateGradient (dl_cs/mri/algorithms.py:11-73)
 * inverts.  lamda from the host float or, if lamda_dev is non-NULL, from that device address.  x_out must not alias x. */
int dlsg_sense_normal_reg_c64(const void* x, const void* maps, const float* weights, const long long* w_strides,
                              float lamda, const float* lamda_dev, void* x_out, void* workspace, size_t workspace_bytes,
                              int B, int E, int C, int T, int Y, int X, int ndims, void* stream);

/* Fused data-consistency update (dl_cs/models/unrolledswin.py:109):
 *     x_out = x + step * ( A^H(A(x)) - aty )
 * aty [B,E,T,Y,X] complex64 is A^H y precomputed once per reconstruction (unrolledswin.py:96).
 * step: the PGD step size (nn.Parameter init -2.0, unrolledswin.py:88-89).  If step_dev is non-NULL the
 * kernel reads the scalar from that device address instead (no host sync on a CUDA parameter, graph-safe).
 * x_out must not alias x or aty. */
int dlsg_sense_dc_step_c64(const void* x, const void* aty, const void* maps, const float* weights,
                           const long long* w_strides, float step, const float* step_dev, void* x_out,
                           void* workspace, size_t workspace_bytes,
                           int B, int E, int C, int T, int Y, int X, int ndims, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sampled-line path of the same operators.  The reference's k-space weights are k-t sampling masks
 * (dl_cs/mri/utils.py:69-79 get_mask; scripts/reconstruct.py:151): whole ky lines of W are zero.
 * With the Y transform done first, only the lines that carry a non-zero weight need the X
 * transform and the HBM intermediate, so the operators below compute exactly what the dense entry
 * points above compute (for finite data: 0 * k contributes 0 in both), with an intermediate of
 * lmax lines per image instead of Y.  2-D only (ndims == 2), Y and X in {32..320 step 32*R} as
 * accepted by dlsg_sense_lines_supported; everything else stays on the dense entry points.
 * For Y in {128, 160, 192, 256, 320} and E <= 2 the column passes are the second-generation kernels
 * (packed FP32x2 butterflies, pruned radix stage, coil maps staged by TMA: csrc/sense_lines_p.cuh); the environment
 * variable DLSG_LINES_KERNELS=v1 selects the first generation (csrc/sense_lines.cuh) for A/B runs.
 * ------------------------------------------------------------------------------------------------ */

/* Bytes of the device-resident line plan for B slices of T frames with Y lines:
 * int32 counts[B*T] | list[B*T*Y] | pos[B*T*Y] | order[B*T*Y] | cls[B*T*36] | pad to 16 bytes | float wq[B*T*Y*8]
 *   n(b,t) = counts; list = sampled ky ascending; pos[ky] = index of ky in list or -1;
 *   order / cls / wq = the same lines grouped by ky mod 16 with their unit roots exp(-2 pi i ky / Y) and squares:
 *   tables of the pruned column transforms (csrc/sense_lines_p.cuh), built here once instead of in every CTA.
 * Only counts[] is meant to be read by the caller (to obtain lmax). */
size_t dlsg_sense_line_plan_bytes(int B, int T, int Y);

/* Builds the plan from the weights (same addressing as above; one pass over W, run once per
 * SenseModel, i.e. where the reference stores `self.weights`, transforms.py:79-82).  A line is
 * sampled when any weight on it over (c, kx) is non-zero or NaN.  The caller reads counts[] back
 * and passes lmax >= max n(b,t) (>= 1) to the operators. */
int dlsg_sense_line_plan_build(const float* weights, const long long* w_strides, int B, int C, int T, int Y, int X,
                               void* plan, void* stream);

/* The same with a promise instead of the read-back: the caller knows (from the sampling pattern) that no frame has more than
 * max_lines sampled ky lines, passes lmax = max_lines to the operators and never touches counts[].  Lines beyond the promise are
 * dropped - the operators then behave as if their weights were zero - and cls[(b*T+t)*36 + 35] is set to 1 for that frame, so a
 * broken promise can be detected later; nothing is written out of bounds.  max_lines = 0: no cap (the function above). */
int dlsg_sense_line_plan_build_capped(const float* weights, const long long* w_strides, int B, int C, int T, int Y, int X,
                                      int max_lines, void* plan, void* stream);

/* 1 when the sampled-line kernels cover this problem (compiled radices and shared-memory budget). */
int dlsg_sense_lines_supported(int E, int Y, int X, int lmax);

/* Workspace for the sampled-line operators: B*C*T*lmax*X complex64 holds everything at once; any
 * multiple of one slice (C*T*lmax*X*8 bytes, the minimum) is accepted and walked in slice chunks. */
size_t dlsg_sense_lines_workspace_bytes(int B, int C, int T, int X, int lmax, size_t budget_bytes);

/* SenseModel._forward_op (transforms.py:92-98) / _adjoint_op (:84-90) through the sampled lines. */
int dlsg_sense_forward_lines_c64(const void* x, const void* maps, const float* weights, const long long* w_strides,
                                 const void* plan, int lmax, void* k_out, void* workspace, size_t workspace_bytes,
                                 int B, int E, int C, int T, int Y, int X, void* stream);
int dlsg_sense_adjoint_lines_c64(const void* k, const void* maps, const float* weights, const long long* w_strides,
                                 const void* plan, int lmax, void* x_out, void* workspace, size_t workspace_bytes,
                                 int B, int E, int C, int T, int Y, int X, void* stream);

/* Same adjoint for a k-space that need not be in device memory: `k` is any address the device can read, in
 * particular pinned (mapped) host memory, i.e. the zero-filled k-space as the reference's loader leaves it on the
 * host (scripts/reconstruct.py:132-171).  Its sampled lines are first copied into the workspace (16-byte loads
 * straight from `k`), so only those lines cross PCIe and the dense k-space never needs a device copy.
 * LIFETIME: the call only enqueues work; `k` is read when the gather kernel runs on `stream`.  The caller must keep
 * the buffer alive and unmodified until work enqueued on `stream` after this call has completed (record an event after
 * the call and release / recycle the buffer only when it has fired).  `k` must be a device-readable contiguous
 * [B,C,T,Y,X] array (pinned + mapped host memory or device memory); pageable memory faults. */
int dlsg_sense_adjoint_lines_gather_c64(const void* k, const void* maps, const float* weights, const long long* w_strides,
                                        const void* plan, int lmax, void* x_out, void* workspace, size_t workspace_bytes,
                                        int B, int E, int C, int T, int Y, int X, void* stream);

/* A^H A x with the epilogues of the dense entry points, selected by `epilogue`:
 *   0  x_out = A^H A x                                   (dlsg_sense_normal_c64)
 *   1  x_out = x + step * (A^H A x - aty)                (dlsg_sense_dc_step_c64, unrolledswin.py:109)
 *   2  x_out = A^H A x + step * x                        (dlsg_sense_normal_reg_c64, unrolledswin.py:154)
 * step_dev (device pointer) overrides step when non-NULL; aty is only read for epilogue 1. */
int dlsg_sense_normal_lines_c64(const void* x, const void* aty, const void* maps, const float* weights,
                                const long long* w_strides, const void* plan, int lmax,
                                float step, const float* step_dev, int epilogue, void* x_out,
                                void* workspace, size_t workspace_bytes,
                                int B, int E, int C, int T, int Y, int X, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Shifted-window attention of the Swin denoiser block
 * (dl_cs/models/video_swin_transformer_mri_downsample.py).  All tensors fp32, contiguous.
 * `window` / `shift` are the values actually used for this volume, i.e. after get_window_size (:72-85).
 * ------------------------------------------------------------------------------------------------ */

/* LayerNorm (norm1) + zero pad to window multiples + cyclic shift + window_partition, the first half of
 * SwinTransformerBlock3D.forward_part1 (:215-236) in one pass:
 *   x [B,D,H,W,C]  ->  xw [B*nW, wd*wh*ww, C];  gamma / beta may be NULL (1 / 0);  padded tokens are exact zeros
 *   (the reference pads after the norm, :219-225). */
int dlsg_swin_ln_partition_f32(const float* x, const float* gamma, const float* beta, float eps, float* xw,
                               int B, int D, int H, int W, int C, const int* window, const int* shift, void* stream);

/* window_reverse + reverse cyclic shift + crop, the tail of forward_part1 (:238-248):
 *   yw [B*nW, N, C]  ->  y [B,D,H,W,C] */
int dlsg_swin_window_reverse_f32(const float* yw, float* y, int B, int D, int H, int W, int C,
                                 const int* window, const int* shift, void* stream);

/* Attention core of WindowAttention3D.forward (:148-167), between the qkv and proj Linear layers:
 *   out[b_, i, h*hd + d] = sum_j softmax_j( scale * q_i . k_j + table[index[i,j], h] + mask[b_ % nW, i, j] ) * v_j[d]
 *   qkv        [B_, N, 3, heads, head_dim]  exactly the output of the qkv Linear (:146), no permute
 *   bias_table [(2Wd-1)(2Wh-1)(2Ww-1), heads] the module parameter; index[i,j] is regenerated in the kernel
 *              from module_window = (Wd, Wh, Ww) (:115-129); N may be smaller than Wd*Wh*Ww (index[:N,:N], :152)
 *   mask_mode  0 none; 1 `mask` = region ids, uint8 [nW, N]: additive -100 where ids differ (what compute_mask
 *              :342-355 encodes); 2 `mask` = dense additive fp32 [nW, N, N] (any mask a caller passes)
 *   out        [B_, N, heads*head_dim]  == (attn @ v).transpose(1, 2).reshape(B_, N, C)  (:167)
 * head_dim <= 32.  fp16 tensor-core operands, fp32 softmax and accumulation. */
int dlsg_swin_window_attention_f32(const float* qkv, const float* bias_table, float* out,
                                   int B_, int N, int heads, int head_dim, const int* module_window, float scale,
                                   int mask_mode, const void* mask, int nW, void* stream);

/* Attention core on fp16 q / k / v as written by dlsg_swin_qkv_fused_f32:
 *   q16 (already multiplied by scale*log2(e)), k16, v16: each [B_, heads, N, 32], zero beyond head_dim except v16[..., head_dim] = 1
 *   when head_dim < 32 (the softmax row sums then come out of the PV MMA); N % 8 == 0.
 * Everything else as dlsg_swin_window_attention_f32. */
int dlsg_swin_window_attention_f16(const void* q16, const void* k16, const void* v16, const float* bias_table, float* out,
                                   int B_, int N, int heads, int head_dim, const int* module_window,
                                   int mask_mode, const void* mask, int nW, void* stream);

/* norm1 + pad + roll + window_partition + qkv Linear (:218-236 and :146) as ONE tcgen05 GEMM: the A rows are window tokens
 * gathered from x [B,D,H,W,C] through the cyclic shift and LayerNorm'ed on the fly; the epilogue writes the fp16 triple above
 * (q pre-scaled by scale*log2(e)).  weight [3C, C], bias [3C] or NULL, gamma / beta [C] or NULL.  head_dim = C/heads <= 32,
 * head_dim % 4 == 0, tokens per window % 8 == 0, C % 16 == 0. */
int dlsg_swin_qkv_fused_f32(const float* x, const float* gamma, const float* beta, float eps,
                            const float* weight, const float* bias, void* q16, void* k16, void* v16,
                            int B, int D, int H, int W, int C, const int* window, const int* shift,
                            int heads, float scale, void* stream);

/* The same operation with the qkv weight streamed in three chunks (q, k, v) instead of held whole in shared memory, so that two
 * CTAs share an SM: wqkv_f16 is qkv.weight [3C, C] as fp16 in the chunked core-matrix order of dlsg_swin_mlp_f32's w1_f16
 * (wqkv_f16[nc][k/8][n/8][n%8][k%8] = weight[nc*C + n, k]), prepared once by the caller.  gamma / beta required.  C % 32 == 0, C <= 256. */
int dlsg_swin_qkv_stream_f32(const float* x, const float* gamma, const float* beta, float eps,
                             const void* wqkv_f16, const float* bias, void* q16, void* k16, void* v16,
                             int B, int D, int H, int W, int C, const int* window, const int* shift,
                             int heads, float scale, void* stream);

/* proj Linear (:168) + window_reverse + reverse shift + crop (:238-248) as ONE tcgen05 GEMM: rows of attn_out [B*nW*N, C]
 * are projected and written to their place in y [B,D,H,W,C]; rows of padded tokens are dropped. */
int dlsg_swin_proj_scatter_f32(const float* attn_out, const float* weight, const float* bias, float* y,
                               int B, int D, int H, int W, int C, const int* window, const int* shift, void* stream);

/* The same with the block's shortcut folded in (`x = shortcut + drop_path(part1)`, :262, DropPath = identity at inference):
 * y[token] = residual[token] + proj(attn_out row); residual [B,D,H,W,C] or NULL. */
int dlsg_swin_proj_scatter_residual_f32(const float* attn_out, const float* weight, const float* bias, const float* residual,
                                        float* y, int B, int D, int H, int W, int C, const int* window, const int* shift, void* stream);

/* forward_part2 of the Swin block (:251-252; Mlp :20-38, norm2 :208) as two tcgen05 GEMM kernels:
 *     y = [x +] fc2( GELU( fc1( LayerNorm(x) ) ) )        (residual != 0 gives the block's `x + forward_part2(x)`, :271)
 * x, y [M, C] fp32 (tokens of [B,D,H,W,C]); gamma / beta [C]; b1 [H], b2 [C] fp32 or NULL.  The weights are passed as fp16 in
 * the tensor core's core-matrix order, prepared once by the caller (constants at inference): with nk = H / C chunks,
 *   w1_f16[nc][k/8][n/8][n%8][k%8] = fc1.weight[nc*C + n, k]        (chunk = C output rows,    n, k < C)
 *   w2_f16[kc][k/8][n/8][n%8][k%8] = fc2.weight[n, kc*C + k]        (chunk = C input columns,  n, k < C)
 * hidden_f16 is a caller-owned fp16 workspace of ceil(M/128)*128*H elements.  GELU is the exact erf form (nn.GELU default).
 * C % 32 == 0, C <= 256, H % C == 0. */
int dlsg_swin_mlp_f32(const float* x, const float* gamma, const float* beta, float eps,
                      const void* w1_f16, const float* b1, const void* w2_f16, const float* b2,
                      void* hidden_f16, float* y, long long M, int C, int H, int residual, void* stream);

/* Linear layer on the tcgen05 tensor cores (qkv :146 and proj :168 of WindowAttention3D):
 *     y[M, N] = x[M, K] @ weight[N, K]^T + bias[N]       (torch.nn.functional.linear semantics)
 * fp32 in/out, fp16 operands, fp32 accumulation in tensor memory.  K % 16 == 0, N % 32 == 0, N <= 512,
 * N*K*2 + 128*K*2 + 4N + 64 bytes of shared memory must fit 227 KB; x and weight 16-byte aligned. */
int dlsg_swin_linear_f32(const float* x, const float* weight, const float* bias, float* y,
                         int M, int N, int K, void* stream);

/* ------------------------------------------------------------------------------------------------
 * k-t pre-processing either side of the operator (dl_cs/mri/utils.py), complex64, contiguous.
 * ------------------------------------------------------------------------------------------------ */

/* fftmod (:7-19), in place on n_images images of Y x X: data[..., y, x] *= -(-1)^(x+y) (the three sign flips combined). */
int dlsg_fftmod_c64(void* data, long long n_images, int Y, int X, void* stream);

/* get_mask (:69-79): mask[i] = |data[i]| > eps ? 1 : 0. */
int dlsg_get_mask_c64(const void* data, float* mask, long long n, float eps, void* stream);

/* sliding_window (:37-49) / time_average (:29-34) along the middle axis of data [outer, T, inner]:
 *   out[o, t, p] = sum_{j < window} data[o, (t - window/2 + j) mod T, p] / (#{|data| > mask_eps in that window} + eps)
 * keep_all_frames != 0: out [outer, T, inner];  keep_all_frames == 0: out [outer, 1, inner], frame t = 0 only —
 * with window == T that is time_average(dim, keepdim=True) (a full-ring sum does not depend on the phase). */
int dlsg_sliding_window_c64(const void* data, void* out, long long outer, int T, long long inner, int window,
                            int keep_all_frames, float mask_eps, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Denoiser convolutions: the layers of ConvBlock / PatchEmbed3D / PatchUnembed3D
 * (dl_cs/models/swin3D.py:120-134, 225-273 == resnet3d.py; video_swin_transformer_mri_downsample.py:455-520), fp32 storage,
 * channels-last, as implicit GEMMs on the tcgen05 tensor cores with split operands: every fp32 operand is a bf16 head (round to
 * nearest) plus a bf16 remainder, and a product is the three (Cout = 64: four) tensor-core products hi*hi + lo*hi + hi*lo with fp32
 * accumulation - 16 mantissa bits per operand, 1e-5 relative on a layer's output (one TF32 pass: 5e-4).  The entry points keep the
 * "tf32x3" of the precision mode they implement (models.set_conv_precision) in their names.
 * ------------------------------------------------------------------------------------------------ */

/* K elements per pipeline stage (32) when the tensor-core kernel takes this channel pair (Cin % 32 == 0 or Cin in {2, 4}; Cout % 32 == 0;
 * shared memory), else 0.  Bit 0x100 is set when the 3x3x3 layer runs on the operand-swapped kernel (Cout = 64), whose weight image
 * stacks heads and remainders as rows (see below). */
int dlsg_conv3d_k3_supported(int Cin, int Cout);

/* out[n,t,y,x,:] = bias + sum over the 27 taps and Cin of x[n,t+dt,y+dy,x+dx,ci] * weight[co,ci,dt+1,dy+1,dx+1]  (cross-correlation,
 * zeros outside the volume).  x [NB,T,Y,X,Cin] and out [NB,T,Y,X,Cout] are channels-last (torch.channels_last_3d of [NB,C,T,Y,X]).
 * wprep is the bf16 weight image the kernel stages with one bulk copy per pipeline stage, prepared once per weight version by the caller
 * (h = 0: bf16 head of weight[co,ci,tap], h = 1: bf16 of the remainder; tap = (dt+1)*9+(dy+1)*3+(dx+1)):
 *   Cin % 32 == 0:            wprep[tap][ci/32][h][(ci%32)/8][co/8][co%8][ci%8]
 *   stacked (code & 0x100):   wprep[tap][ci/32][(ci%32)/8][h][co/8][co%8][ci%8]
 *   Cin = 2 or 4 (first layer of a denoiser), K index k = tap*Cin + ci zero-padded to a multiple of 32:
 *                             wprep[k/32][h][(k%32)/8][co/8][co%8][k%8]
 * relu_in != 0 rectifies x on the fly (ConvBlock's pre-activation, :262-268).  resid (NULL or [NB,T,Y,X,Cout]) is a skip connection added
 * in the epilogue: out = conv + bias + resid_scale * resid (ResBlock :300, ResSwinTransformer3DBlock :340, DeepFeatureExtraction :369);
 * a NEGATIVE resid_scale means |resid_scale| * relu(resid): ResBlock's first ConvBlock rectifies, in place, the very tensor its skip
 * connection re-reads (resnet3d.py:240) - with the ReLU folded into both reads that tensor is never rewritten.
 * bias may be NULL.  NB*T*Y*X*Cin < 2^31.
 * t_lo, t_n: only the output frames t_lo .. t_lo + t_n - 1 of every sample are computed and written (t_n = 0: all T).  The denoisers pad the
 * time axis circularly and crop it at the end (swin3D.py:401-417), so the outer frames of a layer that no later layer reads need not exist;
 * the frames that are computed are bit-identical to the full convolution (x is still read one frame beyond the window). */
int dlsg_conv3d_k3_tf32x3_f32(const float* x, const void* wprep, const float* bias, const float* resid, float resid_scale, float* out,
                              int NB, int T, int Y, int X, int Cin, int Cout, int relu_in, int t_lo, int t_n, void* stream);

/* PatchEmbed3D.proj (dl_cs/models/video_swin_transformer_mri_downsample.py:455-476): nn.Conv3d(kernel = stride = patch, no padding) on
 * the same kernel.  x [NB,T,Y,X,Cin] (T, Y, X multiples of patch) -> out [NB,T/patch,Y/patch,X/patch,Cout]; wprep as above (first form) with
 * patch^3 taps, tap = (dt*patch + dy)*patch + dx.  Cin % 32 == 0. */
int dlsg_conv3d_patch_tf32x3_f32(const float* x, const void* wprep, const float* bias, float* out,
                                 int NB, int T, int Y, int X, int Cin, int Cout, int patch, void* stream);

/* PatchUnembed3D.proj (:503-520): nn.ConvTranspose3d(kernel = stride = patch): out[n, t*patch+dt, y*patch+dy, x*patch+dx, :] =
 * bias + W_tap x[n,t,y,x,:].  x [NB,T,Y,X,Cin] -> out [NB,T*patch,Y*patch,X*patch,Cout]; wprep as above built from
 * weight[ci,co,dt,dy,dx] transposed to [co,ci,...].  Cin % 32 == 0. */
int dlsg_conv3d_unpatch_tf32x3_f32(const float* x, const void* wprep, const float* bias, float* out,
                                   int NB, int T, int Y, int X, int Cin, int Cout, int patch, void* stream);

/* The denoisers' last layer (final_layer, dl_cs/models/swin3D.py:395) in two steps, each one pass over its input:
 *   z[v, tap*Cout + co] = sum_ci weight[co,ci,tap] * relu?(x[v,ci])       dlsg_conv3d_pointwise_tf32x3_f32 (the tensor-core kernel with one tap;
 *                                                                         its "Cout" is the padded 27*Cout, wprep the image of that [ZC, Cin] matrix)
 *   out[v, co] = bias[co] + sum over the taps inside the volume of z[v + tap, tap*Cout + co]      dlsg_conv3d_tapsum_f32 (ZC = channel stride of z) */
int dlsg_conv3d_pointwise_tf32x3_f32(const float* x, const void* wprep, const float* bias, float* out,
                                     int NB, int T, int Y, int X, int Cin, int Cout, int relu_in, int t_lo, int t_n, void* stream);
int dlsg_conv3d_tapsum_f32(const float* z, const float* bias, float* out, int NB, int T, int Y, int X, int ZC, int Cout,
                           int t_lo, int t_n, void* stream);

/* The same layer as ONE direct kernel: 3 x 3 x 3 'same' convolution down to Cout = 2 or 4 channels in
 * exact fp32 on the CUDA cores (packed FMAs).  x, out channels-last as above; w_tcc[tap][ci][co] = weight[co,ci,tap].  Cin % 32 == 0. */
int dlsg_conv3d_k3_fewout_f32(const float* x, const float* w_tcc, const float* bias, float* out,
                              int NB, int T, int Y, int X, int Cin, int Cout, int relu_in, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DLSG_B200_H */
