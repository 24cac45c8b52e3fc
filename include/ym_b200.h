/* libym_b200 — C ABI of the B200 (sm_100a) detection-forward hot path of YOLO-Master.
 *
 * The reference has no FFI on this path: every operator is an ATen call behind the `ultralytics.nn.modules`
 * Python classes.  This header is therefore the boundary a maintainer would bind with
 * ctypes: each entry point names the reference operator(s) it replaces (file:line under
 * /root/reference/ultralytics).  Conventions:
 *   - every function returns 0 on success, non-zero on error; `ym_last_error()` gives the thread-local message;
 *   - all pointers are DEVICE pointers unless the name says `host`; no torch types cross the boundary;
 *   - activations are NHWC fp16; `ld*` = row pitch in elements (lets a conv read/write a channel slice of a
 *     wider concat buffer in place); `stream` is a cudaStream_t passed as void*;
 *   - kernels never synchronise, allocate, or keep global mutable state (re-entrant, graph-capturable).
 */
#ifndef YM_B200_H
#define YM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

const char* ym_last_error(void);
int ym_version(void);
int ym_device_info(int* sm_major, int* sm_minor, int* sm_count, long long* l2_bytes);

/* Conv (k x k, stride s, pad p, groups=1) + folded-BN bias + optional SiLU (+ residual add), implicit GEMM.
 * Replaces Conv.forward / forward_fuse  nn/modules/conv.py:69-89 (+ fuse_conv_and_bn utils/torch_utils.py:315-349),
 * Bottleneck's `x + cv2(cv1(x))` add  nn/modules/block.py:484-486, and bare nn.Conv2d(+bias) heads  nn/modules/head.py:104-119.
 * w: packed fp16 [Cout][Kpad], k index = (ky*KW + kx)*Cin + ci, Kpad = ceil32(KH*KW*Cin) zero padded.
 * out: fp16 (out_f32=0) or fp32 (out_f32=1) [B*Ho*Wo][ldo]; res: optional fp16 residual (same rows), act: 0 none, 1 SiLU. */
int ym_conv2d_nhwc(const void* x, int ldx, int B, int H, int W, int Cin, const void* w, int Kpad, const float* bias,
                   int Cout, int KH, int KW, int stride, int pad, void* out, int ldo, int out_f32, const void* res,
                   int ldr, int act, void* stream);
/* Kernel behind ym_conv2d_nhwc for 3x3 / pad 1 layers with Cin in {8, 16}, Cout in {8, 16, 32}, fp16 output: 1 = patch-staged kernel
 * (csrc/small_conv.cu, default), 0 = the implicit GEMM every other shape takes.  Returns the previous setting. */
int ym_set_small_conv_impl(int impl);
/* Epilogue organisation of the persistent tcgen05 convolution kernel behind ym_conv2d_tc: 2 = two groups of four warps take alternate
 * tiles (default), 1 = all eight warps work on every tile.  Bit-identical results.  Returns the previous setting. */
int ym_set_conv2_epi_groups(int n);
/* Timing experiments on the persistent tcgen05 convolution kernel (tools/conv2_probe.py): bit 0 no TMA store, bit 1 no epilogue
 * arithmetic / staging, bit 2 no tensor-memory read, bit 3 weights loaded for a CTA's first tile only, bit 4 no proxy fence, bit 5 no epilogue barrier, bit 6 no staging writes, bits 8-11 operand ring depth.
 * Outputs are INVALID while the value is non-zero; the package never sets it.  Returns the previous value. */
int ym_set_conv2_debug(int flags);

/* model.0 stem: Conv(Cin<=4 -> Cout in {16,32,64}, k3 s2 p1) + bias + SiLU reading the NCHW image (conv.py:69-89).
 * in_dtype: 0 fp16, 1 fp32, 2 uint8 (x/255, engine/predictor.py:175).  out NHWC fp16.
 * wgt_host fp32 [Cin*9][Cout] and bias_host fp32 [Cout] are HOST pointers: the (<= 9.3 KB of) weights ride in the kernel's
 * parameter block so that every FFMA reads its weight from the constant bank (baked in at capture time under a CUDA graph). */
int ym_stem_conv_nchw(const void* img, int in_dtype, int B, int Cin, int H, int W, const float* wgt_host,
                      const float* bias_host, int Cout, void* out, int ldo, void* stream);
/* Kernel behind ym_stem_conv_nchw for Cin <= 3, Cout = 16: 1 = mma.sync implicit GEMM with fp16 operands / fp32 accumulation (default;
 * inputs rounded to fp16 like the reference's fp16 predictor path), 0 = fp32 FFMA kernel.  Returns the previous setting. */
int ym_set_stem_impl(int impl);

/* Depthwise k x k (k in 3/5/7/9, stride 1, pad k/2) + bias (+SiLU) (+add).  Replaces DWConv conv.py:185-199,
 * AAttn.pe block.py:1688,1731 and Attention.pe block.py:1311,1331 (reads V in place from the head-interleaved qkv:
 * source channel of c = (c/grp_w)*grp_stride + grp_off + c%grp_w).  w: fp16 tap-major [k*k][C]. */
int ym_dwconv_nhwc(const void* x, int ldx, int grp_w, int grp_stride, int grp_off, const void* w, const float* bias,
                   int B, int H, int W, int C, int ksize, int act, const void* add, int ldadd, void* out, int ldo,
                   void* stream);
/* Depthwise 7x7 layers with C % 32 == 0 behind ym_dwconv_nhwc: 1 = Toeplitz-GEMM kernel on mma.sync (csrc/dwconv_tc.cu: half the
 * instructions, the same time at bs32), 0 = the FFMA kernel every other depthwise shape takes (default).  Returns the previous setting. */
int ym_set_dwconv_tc(int on);

/* SPPF pooling: slots 1..3 of the [.., 4C] concat buffer = MaxPool(k) chained 1..3 times of slot 0 (block.py:237-242). */
int ym_sppf_pool_nhwc(void* buf, int ld, int B, int H, int W, int C, int k, void* stream);

/* nn.Upsample(nearest, up) + Concat(dim=1) of two tensors (conv.py:616-640); up=1 is a plain concat. */
int ym_concat2_nhwc(const void* a, int lda, int Ca, int up, const void* b, int ldb, int Cb, void* out, int ldo, int B,
                    int H, int W, void* stream);

/* Fused multi-head attention O = softmax((Q*scale)^T K) V over token rows of the qkv conv output.
 * Replaces AAttn.forward block.py:1708-1722 (batch = B*area, N = H*W/area) and Attention.forward block.py:1324-1331.
 * d_qk must be 32; d_v 32 or 64.  Output channel = head*d_v + d. */
int ym_attention_fwd(const void* qkv, int ld, int batch, int N, int heads, int head_stride, int q_off, int k_off,
                     int v_off, int d_qk, int d_v, float scale, void* out, int ldo, void* stream);
/* Same contract on the tcgen05 path: S = QK^T and O += PV as tcgen05.mma with S/O in tensor memory, one query row per
 * thread (shuffle-free softmax), lazy O rescaling.  ym_attention_fwd dispatches here by default;
 * ym_set_attention_impl(0) selects the mma.sync kernel (returns the previous setting). */
int ym_attention_fwd_tc(const void* qkv, int ld, int batch, int N, int heads, int head_stride, int q_off, int k_off,
                        int v_off, int d_qk, int d_v, float scale, void* out, int ldo, void* stream);
/* tcgen05 kernel only: compute every `every`-th softmax exponential with an FMA-pipe polynomial (2^f, |rel err| < 7.5e-5)
 * instead of the MUFU (0 = MUFU only, the default).  Measured on B200 (P3 shape, bs32): 1.005 ms at 0, 1.001 / 1.015 / 1.018 /
 * 1.099 ms at 6 / 4 / 3 / 2 - the kernel's per-tile dependency chain, not the MUFU rate, sets its time.  Returns the previous
 * setting. */
int ym_set_attention_poly(int every);
/* tcgen05 kernel only, schedule selector (returns the previous setting):
 *   0  whole-row softmax; P through shared memory (SS-mode MMAs, row sums as a P x ones MMA)
 *   1  chunked softmax: each S row is read from tensor memory in four 16-column chunks and the exponentials of one chunk run
 *      while the next is in flight (running maximum per chunk, already-packed P values rescaled on growth)
 *   2  whole-row softmax; P written to tensor memory (tcgen05.st) and consumed by a TS-mode tcgen05.mma, row sums in registers
 *      (default: removes 50 of the 74 KB per KV tile that crossed the 128 B/clk shared-memory pipe; 1.005 -> 0.922 ms)
 *   3  1 + 2 */
int ym_set_attention_chunked(int mode);
/* Same contract, warp-specialised (tc_attention2.cu; the default of ym_attention_fwd): one CTA = two 128-row query tiles of one
 * (image, head) sharing every K / V tile, K / V / Q by TMA (4-D tensor maps over the packed qkv buffer: d, token, head, image) into a
 * 6-stage ring, one thread issuing all TMA + tcgen05.mma, eight softmax warps (thread = query row) that hand S / P / O over through
 * mbarriers only - no __syncthreads in the key loop, QK(t+1) issued while softmax(t) runs.  Same arithmetic as ym_attention_fwd_tc
 * (lazy rescale, P in tensor memory, TS-mode PV).  `_supported`: tensor-map strides must be multiples of 16 bytes. */
int ym_attention_fwd_tc2(const void* qkv, int ld, int batch, int N, int heads, int head_stride, int q_off, int k_off,
                         int v_off, int d_qk, int d_v, float scale, void* out, int ldo, void* stream);
int ym_attention_fwd_tc2_supported(int heads, int head_stride, int ld);
/* ym_attention_fwd_tc2, d_v = 32: every `every`-th exponential of a score row is computed on the FMA pipe (degree-3 polynomial,
 * |rel err| < 7.5e-5) instead of the MUFU, whose 16 ex2 / clk / SM is the kernel's ceiling (0 = MUFU only; 4).  Returns the
 * previous setting; ym_attention2_poly reads it. */
int ym_set_attention2_poly(int every);
/* Query tiles (128 rows) per CTA of ym_attention_fwd_tc2: 0 = chosen by wave fit (two tiles share every K / V tile; one tile per CTA when
 * the grid is below two waves and splits better), 1 / 2 = forced.  Results are bit-identical either way.  Returns the previous setting. */
int ym_set_attention2_qtiles(int n);
int ym_attention2_poly(void);
/* Scheduling variants of ym_attention_fwd_tc2 (bit 0: row maximum as four independent chains, bit 1: K / V ring refilled three key tiles
 * behind the consumers, bit 2: issuing threads sleep between barrier polls).  Results are bit-identical for every value.  Returns the
 * previous setting. */
int ym_set_attention2_variant(int bits);
/* Kernel behind ym_attention_fwd: 2 = ym_attention_fwd_tc2 (default), 1 = ym_attention_fwd_tc, 0 = mma.sync kernel.  Returns the
 * previous setting (A/B baselines for tests and profiles; nothing in the package changes it). */
int ym_set_attention_impl(int impl);

/* ---- routed expert FFN on tcgen05 with the hidden activation kept on chip (csrc/tc_moe.cu) --------------------------------------
 * SimpleExpert (moe/experts.py:73-88: 1x1 -> GroupNorm -> SiLU -> 1x1 -> GroupNorm) of the top-k experts of every image inside
 * OptimizedMOEImproved.forward (moe/modules.py:1128-1157).  Problem p = (image p / topk, rank p % topk), expert route_idx[p].
 *   x  fp16 [B][HW][ldx] (C channels used), w1 fp16 [E][HID][C], w2 fp16 [E][C][HID], (C, HID) = (64, 128) or (128, 256)
 *   stage 1: stats[P][strips][HID/8][2] = GroupNorm-1 partial sums (sum, sum of squares per 8-channel slice) of h = x W1[e]^T, h
 *            rounded to fp16 as the reference's fp16 execution stores it; h itself is never written
 *   ym_gn_finalize_tiles(stats, P, strips, groups, HID, HW * HID / groups, eps, gamma, beta, route_idx, NULL, a_scale, a_shift)
 *   stage 2: out[P][HW][C] = SiLU(h * a_scale + a_shift) W2[e]^T in fp16 (h recomputed, normalised and fed to the second tcgen05.mma from
 *            tensor memory) and stats[P][strips][C/8][2] = GroupNorm-2 partial sums of out
 * `strips` = ym_moe_ffn_strips(HW, P) row strips per problem (one CTA each); a dropped route (route_idx < 0) writes zero statistics. */
int ym_moe_ffn_supported(int C, int HID, int ldx);
int ym_moe_ffn_strips(int HW, int P);
long long ym_moe_ffn_stats_floats(int P, int strips, int N);
int ym_moe_ffn(int stage, const void* x, int ldx, int B, int HW, int C, int HID, int topk, const void* w1, const void* w2, int E,
               const int* route_idx, const float* a_scale, const float* a_shift, void* out, float* stats, int strips, void* stream);
/* Combine of OptimizedMOEImproved (moe/modules.py:1144-1157) on tcgen05: y[b] = SiLU(x[b] Ws^T + bs) + sum_j (o[b*topk+j] * o_scale +
 * o_shift) (+ x[b]); ws fp16 [C][C] (BatchNorm folded), o fp16 [B*topk][HW][C], o_scale / o_shift fp32 [B*topk][C] (GroupNorm-2 affine
 * times the routing weight), C = 64 or 128.  Same arithmetic order as ym_moe_combine (mma.sync), which serves the other widths. */
int ym_moe_combine_tc_supported(int C, int ldx, int ldo);
int ym_moe_combine_tc(const void* x, int ldx, int B, int HW, int C, const void* ws, const float* bias_s, const void* o,
                      const float* o_scale, const float* o_shift, int topk, void* out, int ldo, int add_residual, void* stream);
/* The same two kernels finalising their GroupNorm THEMSELVES from the partial sums of the kernel before (no ym_gn_finalize_tiles launch in
 * between; bit-identical to the two-launch form): pass 2 of ym_moe_ffn with GroupNorm-1 given as (partial sums of pass 1, groups, element
 * count per group, eps, gamma / beta fp32 [E][HID]); the combine with GroupNorm-2 given the same way (partial sums of pass 2 over
 * `gn2_tiles` = strips, gamma / beta fp32 [E][C]) plus the routing table and weights (route_w [B*topk] is folded into the affine). */
/* Pass 1 of ym_moe_ffn with the router's finish in its prologue: idx_out / w_out (/ probs_out) are OUTPUTS (same values as
 * ym_router_topk writes), used by this kernel for its weight selection and by the kernels after it. */
int ym_moe_ffn_routed(const void* x, int ldx, int B, int HW, int C, int HID, int topk, const void* w1, int E, const float* partial, int nblk,
                      int Cr, int npix, const float* rw2, const float* rscale2, const float* rshift2, int* idx_out, float* w_out,
                      float* probs_out, float* stats, int strips, void* stream);
int ym_moe_ffn_gn(const void* x, int ldx, int B, int HW, int C, int HID, int topk, const void* w1, const void* w2, int E, const int* route_idx,
                  const float* gn1_stats, int gn1_groups, float gn1_count, float gn1_eps, const float* gamma1, const float* beta1, void* out,
                  float* stats, int strips, void* stream);
int ym_moe_combine_tc_gn(const void* x, int ldx, int B, int HW, int C, const void* ws, const float* bias_s, const void* o,
                         const float* gn2_stats, int gn2_tiles, int gn2_groups, float gn2_count, float gn2_eps, const float* gamma2,
                         const float* beta2, const int* route_idx, const float* route_w, int topk, void* out, int ldo, int add_residual,
                         void* stream);
/* ym_gn_finalize with an explicit number of partial-sum tiles per problem. */
int ym_gn_finalize_tiles(const float* stats, int P, int tiles, int groups, int C, float count, float eps, const float* gamma,
                         const float* beta, const int* route_idx, const float* route_w, float* scale, float* shift, void* stream);
/* Programmatic dependent launch of the forward-path kernels (default on; YM_PDL=0 in the environment or ym_set_pdl(0) turns the launch
 * attribute off - results are identical, only launch overlap changes).  ym_set_pdl returns the previous setting. */
int ym_pdl_enabled(void);
int ym_set_pdl(int on);
/* Launch priority of every forward-path kernel EXCEPT the attention kernels (those stay at 0, the lowest): 0 = off, -1 .. -8 are passed to
 * cudaLaunchAttributePriority.  With several graph instances in flight the slots a long attention kernel frees then go to the other
 * instances' short kernels first.  Takes effect at launch (capture) time.  Returns the previous value. */
int ym_kernel_priority(void);
int ym_set_kernel_priority(int prio);

/* EfficientSpatialRouter.forward + BaseRouter._process_logits (eval)  moe/routers.py:283-304, :185-265.
 * w1: fp32 [9][C/4][Cr][4] (tap-major, float4 over channels), scale1/shift1: folded BN1 [Cr]; w2: fp32 [E][Cr], scale2/shift2: folded BN2 [E].
 * Writes idx int32 [B,topk] (descending prob), w fp32 [B,topk] (renormalised), probs fp32 [B,E] (nullable). */
long long ym_router_scratch_floats(int B, int H, int W, int C, int Cr, int pool);
int ym_router_topk(const void* x, int ldx, int B, int H, int W, int C, int pool, const float* w1, int Cr,
                   const float* scale1, const float* shift1, const float* w2, const float* scale2, const float* shift2,
                   int E, int topk, float* scratch, int* idx_out, float* w_out, float* probs_out, void* stream);
/* The two halves of ym_router_topk: ym_router_partial runs the fused pool + conv3x3 + BN + SiLU pass and leaves per-tile partial sums
 * [B][nblk][Cr] at the start of `scratch` (nblk = ym_router_blocks(H, W, pool, &npix)); the finish (mean -> logits -> softmax -> top-k)
 * then runs either as its own launch (ym_router_topk) or in the prologue of ym_moe_ffn_routed. */
int ym_router_blocks(int H, int W, int pool, int* npix);
int ym_router_partial(const void* x, int ldx, int B, int H, int W, int C, int pool, const float* w1, int Cr, const float* scale1,
                      const float* shift1, float* scratch, void* stream);

/* Routed expert GEMM, one problem per (image, k): replaces the per-expert Python loop + x[batch_idx] gather of
 * OptimizedMOEImproved.forward moe/modules.py:1128-1142 for SimpleExpert's two 1x1 convs (moe/experts.py:79-85).
 * Problem p: expert e = route_idx[p]; A = a + (p/a_div)*HW*lda [HW x K]; out + p*HW*ldo [HW x N] fp16.
 * a_scale/a_shift (nullable, [P][K]): A := SiLU(A*scale + shift) (GroupNorm+SiLU of the hidden, fused on load).
 * stats (nullable, ym_moe_stats_floats(P,HW,N) floats): per-(problem, M tile, 8-channel tile) partial sum / sum-of-squares
 * of the stored output, written without atomics so the statistics are bit-reproducible. */
long long ym_moe_stats_floats(int P, int HW, int N);
int ym_moe_expert_gemm(const void* a, int lda, int a_div, int P, int HW, int K, const void* w, int Kpad,
                       long long w_expert_stride, const int* route_idx, int N, void* out, int ldo, const float* a_scale,
                       const float* a_shift, float* stats, int groups, void* stream);

/* GroupNorm partial statistics (fixed-order reduction) -> per-(problem, channel) affine: scale = rw*rstd*gamma[e], shift = rw*(beta[e] - mean*rstd*gamma[e])
 * (nn.GroupNorm inside SimpleExpert, experts.py:81,84; rw = routing weight folds modules.py:1139-1142 into GN2). */
int ym_gn_finalize(const float* stats, int P, int HW, int groups, int C, float count, float eps, const float* gamma,
                   const float* beta, const int* route_idx, const float* route_w, float* scale, float* shift, void* stream);

/* out = [x +] SiLU(BN(shared 1x1(x))) + sum_j (o_j*o_scale_j + o_shift_j): shared expert + weighted expert sum in fp32
 * + ABlockMoE residual (moe/modules.py:1085,1147-1157,1256-1258). */
int ym_moe_combine(const void* x, int ldx, int B, int HW, int C, const void* ws, int Kpad, const float* bias_s,
                   const void* o, int ldo_o, const float* o_scale, const float* o_shift, int topk, void* out, int ldo,
                   int add_residual, void* stream);

/* Detect post-processing.  box[l]: fp32 [B, h_l*w_l, 4*reg_max] ltrb distances (reg_max == 1) or DFL bin logits,
 * channel = side*reg_max + bin (block.py:63-85); cls[l]: fp32 [B, h_l*w_l, nc] logits.
 * ym_detect_topk: Detect._inference + postprocess + get_topk_index (end2end)  head.py:173-258, tal.py:398-423.
 *   out fp32 [B, k, 6] = (x1,y1,x2,y2,score,cls), k = min(max_det, A), score-descending; out_anchor int32 [B,k] nullable;
 *   scratch: B*A uint32 (per-anchor max-logit keys written by the first of the two kernels).
 * ym_detect_dense: Detect._inference (+ DFL when reg_max > 1) -> y fp32 [B, 4+nc, A] (xyxy!=0: corner boxes, else xywh)
 *   head.py:173-194. */
int ym_detect_topk(int nl, const void* const* box, const void* const* cls, const int* hs, const int* ws,
                   const float* strides, int B, int nc, int max_det, float* out, int* out_anchor, void* scratch,
                   void* stream);
int ym_detect_dense(int nl, const void* const* box, const void* const* cls, const int* hs, const int* ws,
                    const float* strides, int B, int nc, int reg_max, int xyxy, float* y, void* stream);

/* Element-wise glue of the residual / mixture blocks on NHWC fp16 rows (C % 8 == 0); a or b may be NULL (= 0) where allowed.
 *   op 0  out = a + chan[c]*b            A2C2f layer-scale `x + gamma*y` block.py:1879; ls1/ls2 mot/experts.py:165,170;
 *                                        ls_attn/ls_ffn moa/block.py:271-273                        (p0 = chan fp32[C])
 *   op 1  out = a + tok[row*ldt+toff]*b  per-token routed accumulation mot/block.py:347-364, moa/block.py:232-244
 *   op 2  out = sigmoid(a)*b             GLU gate mot/experts.py:168
 *   op 3  out = gelu(a)                  exact (erf) nn.GELU mot/experts.py:225,365
 *   op 4  out = tok * [silu](a*sc[img,c] + sh[img,c]) + b      GroupNorm apply (+ routed weight + accumulate);
 *                                        img = row / rows_per_img; p0 = sc, p1 = sh fp32[imgs*C]; tok NULL = 1
 *   op 5  out = t*a + (1-t)*b            scalar blend, t = p0[0] (moa/heads.py:371-375)
 *   op 6  out = sigmoid(a)               gate heads of moe/gated.py:1165-1172,1206-1208
 *   op 7  out = a * (1 + t*b)            VisualDetailGate gated.py:1176-1178, t = p0[0] = tanh(detail_scale)
 *   op 8  out = a * b                    context * gate gated.py:1219-1221
 * tok: fp32 per-token weights [rows*ldt] read at column toff. */
int ym_ew_nhwc(int op, const void* a, int lda, const void* b, int ldb, const float* p0, const float* p1, const float* tok,
               int ldt, int toff, int rows_per_img, int act, void* out, int ldo, long long rows, int C, void* stream);

/* ---- Mixture-of-Transformers / Mixture-of-Attention neck blocks (nn/modules/mot/*, nn/modules/moa/*) -------------------
 * GroupNorm statistics -> per-(image, channel) affine (scale = rstd*gamma, shift = beta - mean*rstd*gamma) consumed by
 * ym_ew_nhwc op 4.  Replaces nn.GroupNorm in mot/experts.py:99-100, mot/block.py:145, mot/router.py:109, moa/heads.py:137,203,284,
 * moa/router.py:40.  x: fp16 (x_f32=0) or fp32 (x_f32=1) [B][HW][ld]; scale/shift fp32 [B][C]. */
int ym_groupnorm_stats(const void* x, int x_f32, int ld, int B, int HW, int C, int G, float eps, const float* gamma,
                       const float* beta, float* scale, float* shift, void* stream);

/* nn.LayerNorm(C) per token row (mot/experts.py:215-216,355-356). */
int ym_layernorm_nhwc(const void* x, int ldx, const float* gamma, const float* beta, float eps, void* out, int ldo,
                      long long rows, int C, void* stream);

/* softmax(q k^T * scale) v for small heads (padded head_dim hdp in 8/16/24/32), one query row per thread.
 * q rows [batch*Nq][ldq], k/v rows [batch*Nkv][ld*]; head h = channels [h*hdp, (h+1)*hdp) of each pointer.
 * Replaces F.scaled_dot_product_attention in _LocalConvTransformerExpert (global branch) mot/experts.py:159,
 * _RegionalAttnHead moa/heads.py:243 (Nkv = pooled tokens), _GlobalAttnHead exact branch moa/heads.py:369. */
int ym_attn_small(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int batch, int heads, int hdp,
                  int Nq, int Nkv, float scale, void* out, int ldo, void* stream);

/* Window attention over win x win tokens (win <= 8) with the reference's pad -> roll(-shift) -> partition token mapping done
 * by index arithmetic (no copies); padding tokens carry padk/padv (NULL = zeros; the Window expert pads BEFORE LayerNorm so
 * its padding tokens are qkv(LayerNorm(0)) = W_qkv . beta).  Outputs of padding queries are cropped by the reference and are
 * never computed.  Replaces _WindowTransformerExpert.forward mot/experts.py:270-303, the windowed branch of the LocalConv
 * expert :137-157 and _window_flash_attn moa/heads.py:88-121. */
int ym_attn_window(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int B, int H, int W, int heads,
                   int hdp, int win, int shift, const void* padq, const void* padk, const void* padv, float scale, void* out,
                   int ldo, void* stream);

/* Deformable sampling + point softmax (_DeformableTransformerExpert._deform_attn mot/experts.py:416-475).
 * oa fp32 [rows][ldoa] = [offset logits (head, point, xy) | attention logits (head, point)]; v fp16 [B,H,W,heads*hd]. */
int ym_deform_sample(const float* oa, int ldoa, const void* v, int ldv, int B, int H, int W, int heads, int hd, int np,
                     int align_corners, void* out, int ldo, void* stream);

/* Per-token router: 1x1 (C->HID) -> GroupNorm(G) -> SiLU -> 1x1 (HID->E)+b -> softmax(/T) -> [top-k, renormalise, scatter].
 * _MoTRouter mot/router.py:211-291 (topk < E, temperature = device buffer temp_dev) and _MoARouter moa/router.py:50-62
 * (topk == E, host temperature).  fp32 throughout.  weights fp32 [B*HW][E] dense (zeros off the top-k); idx int32 [B*HW][topk]
 * (nullable).  scratch: ym_token_router_scratch_floats() floats. */
long long ym_token_router_scratch_floats(int B, int HW, int HID);
int ym_token_router(const void* x, int ldx, int B, int HW, int C, const float* w1, int HID, int G, const float* gn_w,
                    const float* gn_b, float gn_eps, const float* w2, const float* b2, int E, int topk, const float* temp_dev,
                    float temp, float* weights, int* idx, float* scratch, void* stream);

/* Performer ReLU-feature linear attention of _GlobalAttnHead._linear_attn moa/heads.py:318-352 (fp32): rf fp32 [nb][hd]. */
long long ym_linear_attn_scratch_floats(int batch, int heads, int hdp, int N);
int ym_linear_attn(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int batch, int heads, int hdp, int hd,
                   int nb, int N, const float* rf, float eps, float limit, float* scratch, void* out, int ldo, void* stream);

/* F.adaptive_avg_pool2d (moa/heads.py:224). */
int ym_adaptive_avgpool_nhwc(const void* x, int ldx, int B, int H, int W, int C, int h, int w, void* out, int ldo, void* stream);

/* nn.AdaptiveAvgPool2d(1) of an NHWC fp16 map x [B][HW][ldx] -> fp16 [B][ldo] (SE / feature / cross gates of the gated MoE family, the
 * latent tokens, Classify): one CTA per (image, 64-channel slab), fixed-order reduction. */
int ym_gap_nhwc(const void* x, int ldx, int B, int HW, int C, void* out, int ldo, void* stream);

/* LatentRouter.forward nn/modules/latent_mixture.py:219-241 (per_token = False) for LatentMixture (:721-734): T pooled token vectors
 * fp16 [B][ld_t] -> mean over tokens of (token + emb[t]) -> LayerNorm -> Linear(C, hid) + SiLU -> Linear(hid, C) + SiLU ->
 * Linear(C, E) -> nan_to_num / clamp(+-30) -> softmax(/ max(temperature, 0.1)).  logits, probs fp32 [B][E]. */
int ym_latent_router(int T, const void* const* tokens, const int* lds, int B, int C, const float* emb, const float* ln_w,
                     const float* ln_b, float ln_eps, const float* w1, const float* b1, int hid, const float* w2, const float* b2,
                     const float* wh, const float* bh, int E, float temperature, float* logits, float* probs, void* stream);

/* Classify tail nn/modules/head.py:823-832 on the pooled feature vector v fp16 [B][ldv]: logits = w . v + b (w fp32 [nc][Cin]) and
 * probs = softmax(logits), both fp32 [B][nc]. */
int ym_classify_head(const void* v, int ldv, int B, int Cin, const float* w, const float* b, int nc, float* logits, float* probs,
                     void* stream);

/* Pose.kpts_decode head.py:644-664: per level kpt fp32 [B][h][w][nk] (the pose tower's output) ->
 * y fp32 [B][nk][A], A = sum h*w: x, y = (v*2 + grid coordinate) * stride, visibility (ndim 3) = sigmoid.  ndim 1 copies the values
 * unchanged: the mask-coefficient rows of Segment._inference head.py:332-344. */
int ym_kpts_decode(int nl, const void* const* kpt, const int* hs, const int* ws, const float* strides, int B, int nk, int ndim,
                   float* y, void* stream);

/* OBB head head.py:477-500 + dist2rbox utils/tal.py:447-453 on top of ym_detect_dense's xywh decode: angle fp32 [B][h][w][1] per level
 * (raw tower output), yin fp32 [B][4+nc][A] -> yout fp32 [B][4+nc+1][A] = rotated centre, w, h, class scores, angle = (sigmoid - 0.25)*pi. */
int ym_obb_finish(int nl, const void* const* angle, const int* hs, const int* ws, const float* strides, int B, int nc,
                  const float* yin, float* yout, void* stream);

/* ops.scale_coords + clip_coords utils/ops.py:596-631,204-225, in place, for the points of one image (PosePredictor.construct_result
 * models/yolo/pose/predict.py:62-66): coords fp32 rows of pitch ld >= 2, (x, y) first; params_host = (gain, pad_x, pad_y, w0, h0). */
int ym_scale_coords(float* coords, int ld, long long n, const float* params_host, int padding, int normalize, void* stream);

/* ---- DiversifiedExpertGroup (moe/gated.py:2214-2333, v0_14 zoo; csrc/mix.cu) ---------------------------------------------------
 * ym_dwconv3_routed_nhwc: dw_layers[e][0] for e = route[b * route_stride]: depthwise 3x3 whose taps (w fp16 [E][9][C], tap-major) AND
 *   dilation (dil int32 [E] on the device; padding = dilation) are chosen per image by the router, without a host round trip
 *   (the reference loops over torch.unique(indices).tolist()).  x fp16 [B][H][W][ldx] -> out fp16 [B][H][W][ldo]; fp32 accumulation.
 * ym_route_affine: in place on (scale, shift) fp32 [B][C] = (rstd, -mean * rstd) from ym_groupnorm_stats with unit gamma:
 *   scale *= gamma[e], shift = shift * gamma[e] + beta[e]  (gamma, beta fp32 [E][C]): the per-expert GroupNorm dw_layers[e][1];
 *   optional route_w fp32 [B]: both are then multiplied by the image's routing weight (the routed projection of the gated family when
 *   its GroupNorm groups are narrower than the 8-channel statistics granule of ym_moe_expert_gemm's epilogue). */
int ym_dwconv3_routed_nhwc(const void* x, int ldx, const void* w, const int* route, int route_stride, const int* dil, int E, int B, int H,
                           int W, int C, void* out, int ldo, void* stream);
int ym_route_affine(float* scale, float* shift, const float* gamma, const float* beta, const int* route, int route_stride, int E, int B,
                    int C, const float* route_w, void* stream);

/* ---- Segment / OBB post-processing (csrc/postproc.cu) ------------------------------------------------
 * ym_process_mask: ops.process_mask(protos, masks_in, bboxes, shape, upsample) ultralytics/utils/ops.py:500-528 (+ crop_mask :477-497)
 *   for the detections of one image.  protos [nm][mh][mw] fp16 (proto_dtype 1) or fp32 (2); dets fp32 rows of pitch ld with the xyxy
 *   box (in `shape` = (in_h, in_w) coordinates) at columns 0..3 and the nm mask coefficients from column coef_col (NMS rows: 6).
 *   logits = coefficients @ prototypes (fp32), bilinear upsampling to (in_h, in_w) as F.interpolate(align_corners=False) when
 *   upsample, crop x1 <= col < x2, y1 <= row < y2 (boxes scaled by (mw/in_w, mh/in_h) when not upsample), > 0.
 *   out uint8 [n][in_h][in_w] (upsample) or [n][mh][mw].  scratch: ym_process_mask_scratch_bytes(n, mh, mw).  n == 0 is a no-op. */
long long ym_process_mask_scratch_bytes(int n, int mh, int mw);
int ym_process_mask(const void* protos, int proto_dtype, int nm, int mh, int mw, const float* dets, int ld, int n, int coef_col,
                    int in_h, int in_w, int upsample, unsigned char* out, void* scratch, void* stream);

/* ym_nms_rotated: non_max_suppression(..., rotated=True) ultralytics/utils/nms.py:13-171 - TorchNMS.fast_nms :193-242 with
 *   batch_probiou utils/metrics.py:293-326: candidate j (score order, ties towards the lower anchor) survives iff no candidate i < j
 *   has ProbIoU >= iou_thres; class offset cls * max_wh on the centre; the first max_nms candidates by score enter, the first max_det
 *   survivors leave.  pred fp32 [B][4+nc+1][A] = xywh, class scores, angle (ym_obb_finish's output).
 *   out fp32 [B][max_det][7] = x, y, w, h, conf, cls, angle; out_count int32 [B]; out_idx int32 [B][max_det] (-1 past the count).
 *   scratch: ym_nms_rotated_scratch_bytes(B, A). */
long long ym_nms_rotated_scratch_bytes(int B, int A);
int ym_nms_rotated(const float* pred, int B, int nc, int A, float conf_thres, float iou_thres, int max_det, int max_nms, float max_wh,
                   float* out, int* out_count, int* out_idx, void* scratch, void* stream);

/* ---- Gated MoE family (VisualEnhancedAdaptiveGateMoE, nn/modules/moe/gated.py) ----------------------
 * ym_gate_router: DualStreamGateRouter.forward gated.py:129-151 (fp32 throughout, as the reference's FP32RouterMixin) followed by
 *   AdaptiveGateMoE._safe_complexity / _apply_complexity_gate gated.py:455-490.  x: fp16 [B][H*W][ldx] (the dynamic channel half).
 *   global stream: [mean | population std] over H*W per channel -> global_fc fp32 [E][2C];
 *   local stream : avg_pool(pool) when H,W > pool -> dw3x3 (dw fp32 [C][9]) -> GN(G1) -> SiLU -> 1x1 (pw1 [R][C]) -> GN(G2) -> SiLU
 *                  -> 1x1 (pw2 [E][R]) + b2 -> spatial mean;
 *   logits = clamp(alpha*global + (1-alpha)*local, +-30) (alpha = sigmoid(self.alpha), passed as a value), softmax(/temperature),
 *   top-k, w / (sum + 1e-6); complexity c = clamp(mean over the BATCH of sigmoid(cx_w . mean_c + cx_b), 0.3, 1.5) keeps the
 *   round(c*topk) best ranks and renormalises.  DualStreamGateRouterV2 (gated.py:181-260, v0_11 / v0_12 zoos): ln_w / ln_b fp32 [2C]
 *   (nullable pair) = LayerNorm over the statistics in front of global_fc, prior fp32 [E] (nullable) added to the blended logits
 *   before the clamp.  Outputs: w fp32 [B][topk], idx int32 [B][topk], probs fp32 [B][E] (nullable).
 *   scratch: ym_gate_router_scratch_floats() floats.  Six small kernels (slab statistics and their merge; the local stream as depthwise + GN partials,
 *   GN + 1x1 + GN partials on pixel slabs, then its head per image; finish), no host synchronisation.
 * ym_fc_gate: out[b][o] = offset + scale * sigmoid(b2[o] + w2[o] . silu(w1 . v[b]))   v fp16 [B][ldv] (a 1x1 adaptive average pool):
 *   se_gate gated.py:325-332 (scale 1), feature_gate moe/hooks.py:50-57 (scale = tanh(refine_scale)), CrossPathGate gated.py:2396-2412
 *   (offset 0.5, scale 0.5*tanh(gate_scale), v0_15 zoo); consumed by ym_ew_nhwc op 4.
 * ym_gated_select: FusedExpertGroup.forward gated.py:1061-1081 after the all-expert grouped conv: fo fp16 [B][HW][ldf] holds
 *   expert e in channels [e*oc, (e+1)*oc); for the routed experts: GroupNorm(G, no affine) over the slice, gamma/beta fp32 [E][oc],
 *   SiLU, sum_j w[b][j] * (.) -> out fp16 [B][HW][ldo].  scratch: ym_gated_select_scratch_floats() floats.  The statistics run on a
 *   (pixel slabs x routes) grid and are merged per route in slab order; three kernels.
 * ym_ctx_mean3: PyramidContextMixer gated.py:1213-1219: (a + nearest_up(b) + nearest_up(c)) / 3 with b (h2,w2), c (h4,w4). */
long long ym_gate_router_scratch_floats(int B, int H, int W, int C, int R, int E, int pool);
int ym_gate_router(const void* x, int ldx, int B, int H, int W, int C, int pool, const float* global_fc, const float* dw,
                   const float* gn1_w, const float* gn1_b, int G1, const float* pw1, int R, const float* gn2_w,
                   const float* gn2_b, int G2, const float* pw2, const float* b2, int E, float gn_eps, float alpha,
                   float temperature, const float* cx_w, float cx_b, int topk, const float* ln_w, const float* ln_b, float ln_eps,
                   const float* prior, float* scratch, float* w_out, int* idx_out, float* probs_out, void* stream);
/* ZeroCostRouter.forward gated.py:953-968 (UltraLightRouter of UltimateOptimizedMoE, v0_3 zoo) + the complexity SCALE of
 * UltimateOptimizedMoE.forward modules.py:1663-1670: probs = softmax(clamp(softmax(fc . [mean | std]) / T, +-30)), top-k,
 * w / (sum + 1e-6), then w *= clamp(mean over the batch of sigmoid(cx_w . mean_c + cx_b), 0.3, 1.5).  fc fp32 [E][2C].
 * scratch: ym_zero_cost_router_scratch_floats() floats. */
/* UltraEfficientRouter.forward moe/routers.py:96-118 (UltraOptimizedMoE, v0_1 uomoe / v0_2 zoos): avg_pool(pool) when H,W > pool ->
 * dw3x3 -> GN -> SiLU -> 1x1 -> GN -> SiLU -> 1x1 + bias, then PER PIXEL clamp(+-30) / T -> softmax over experts, spatial mean, top-k,
 * w / max(sum, 1e-6); weights <= w_min are zeroed (the eval threshold of BatchedExpertComputation moe/utils.py:172-173).
 * scratch: ym_gate_router_scratch_floats(B, H, W, C, R, E, pool) floats. */
int ym_pixel_router(const void* x, int ldx, int B, int H, int W, int C, int pool, const float* dw, const float* gn1_w,
                    const float* gn1_b, int G1, const float* pw1, int R, const float* gn2_w, const float* gn2_b, int G2,
                    const float* pw2, const float* b2, int E, float gn_eps, float temperature, float w_min, int topk, float* scratch,
                    float* w_out, int* idx_out, float* probs_out, void* stream);
long long ym_zero_cost_router_scratch_floats(int B, int C);
int ym_zero_cost_router(const void* x, int ldx, int B, int H, int W, int C, const float* fc, int E, float temperature,
                        const float* cx_w, float cx_b, int topk, float* scratch, float* w_out, int* idx_out, float* probs_out,
                        void* stream);
int ym_fc_gate(const void* v, int ldv, int B, int Cin, const float* w1, int Cr, const float* w2, const float* b2, int Cout,
               float scale, float offset, float* out, void* stream);
long long ym_gated_select_scratch_floats(int B, int topk, int oc);
int ym_gated_select(const void* fo, int ldf, int B, int HW, int E, int oc, int G, float eps, const int* idx, const float* w,
                    int topk, const float* gamma, const float* beta, float* scratch, void* out, int ldo, void* stream);
int ym_ctx_mean3(const void* a, int lda, const void* b, int ldb, const void* c, int ldc, int B, int H, int W, int C, int h2,
                 int w2, int h4, int w4, void* out, int ldo, void* stream);

/* Predictor pre-processing of B same-sized uint8 HWC frames.  Replaces, fused into one pass,
 * LetterBox.apply_image data/augment.py:1792-1822 (cv2.resize INTER_LINEAR + copyMakeBorder BORDER_CONSTANT) and
 * BasePredictor.preprocess engine/predictor.py:166-175 (BGR->RGB, BHWC->BCHW, .half()/.float(), /255).
 *   src uint8 [B][sh][src_pitch] (3 interleaved channels; frame b at src + b*src_stride bytes);
 *   xtab [nw], ytab [nh]: DEVICE tables of {uint32 i0 | i1<<16, uint32 a0 | a1<<16}: the clipped source indices and 11-bit
 *     weights (a0+a1 = 2048) of cv2's fixed-point bilinear kernel, built on the host from LetterBox.get_params
 *     (augment.py:1742-1786); ignored when area2x != 0 (exact 2x downscale in both axes = cv2's INTER_AREA fast path);
 *   the resized nh x nw image lands at (top, left) of the H x W output, everything else is pad_value;
 *   out: chw ? [B][3][H][W] : [B][H][W][3];  out_dtype 0 uint8, 1 fp16 (v/255), 2 fp32 (v/255);  swap_rb reverses the channels. */
int ym_letterbox_u8(const void* src, long long src_stride, int B, int sh, int sw, int src_pitch, const void* xtab,
                    const void* ytab, int area2x, int nw, int nh, int top, int left, int pad_value, int swap_rb, void* out,
                    int out_dtype, int chw, int H, int W, void* stream);

/* ops.scale_boxes + clip_boxes utils/ops.py:119-158,174-201 (DetectionPredictor.construct_result
 * models/yolo/detect/predict.py:107-125), in place on n fp32 rows [ld >= 4]: (b - pad) / gain, then clamp to the original
 * frame unless xywh.  Image of row i = row_img ? row_img[i] : i / rows_per_img.  params_host: HOST fp32 [n_img][5] =
 * (gain, pad_x, pad_y, w0, h0), n_img <= 128 (rides in the kernel's parameter block: no upload, graph-capturable). */
int ym_scale_boxes(float* boxes, int ld, long long n, int rows_per_img, const int* row_img, int n_img,
                   const float* params_host, int padding, int xywh, void* stream);

/* ES_MOE (moe/modules.py:410-741, eval sparse path) on four entry points; the module is router -> per-expert depthwise k x k
 * on the images that retained the expert -> grouped pointwise GEMM (+folded BN, SiLU, routing weight) -> sum + BN + SiLU.
 * ym_esmoe_route: DynamicRoutingLayer.forward/_hard_top_k routers.py:458-496,519-527 + the top-k / dynamic-threshold /
 *   renormalisation of ES_MOE._sparse_forward modules.py:659-684.  w1 fp32 [Cr][C], w2 fp32 [E][Cr].
 *   idx int32 [B][topk] (descending weight, -1 = dropped by the threshold), w fp32 [B][topk], probs [B][E] nullable.
 * ym_esmoe_dwconv: experts.py:284 depthwise for ONE expert over the images that retained it (slot b*topk+j of out).
 * ym_esmoe_pointwise: experts.py:285-296 pointwise + BN + SiLU, scaled by the routing weight, one problem per slot.
 * ym_esmoe_combine: modules.py:690-704 (sum over retained slots) + final BatchNorm + SiLU :496,581. */
long long ym_esmoe_scratch_floats(int B, int HW, int C);
int ym_esmoe_route(const void* x, int ldx, int B, int HW, int C, const float* w1, const float* b1, int Cr, const float* w2,
                   const float* b2, int E, int topk, float dyn_thr, float* scratch, int* idx_out, float* w_out, float* probs_out,
                   void* stream);
int ym_esmoe_dwconv(const void* x, int ldx, const void* w, int B, int H, int W, int C, int ksize, const int* route_idx, int topk,
                    int expert, void* out, int ldo, void* stream);
int ym_esmoe_pointwise(const void* t, int ldt, int P, int HW, int K, const void* w, int Kpad, long long w_expert_stride,
                       const float* bias_all, int N, const int* route_idx, const float* route_w, void* y, int ldy, void* stream);
int ym_esmoe_combine(const void* y, int ldy, const int* route_idx, int topk, const float* scale, const float* shift, void* out,
                     int ldo, int B, int HW, int C, void* stream);

/* Batched NMS / Cluster-Weighted NMS, one CTA per image, no host round trip.
 * mode 0: non_max_suppression utils/nms.py:13-171 (single-label, class-aware) + TorchNMS.nms :245-302: candidates with
 *   best-class conf > conf_thres, class offset cls*max_wh added in fp32 like the reference, greedy IoU > iou_thres,
 *   first max_det survivors; out rows (x1,y1,x2,y2,conf,cls).
 * mode 1: CW-NMS of examples/YOLO-Master-Cross-Platform-Edge-Deployment/cpp/src/common.cpp:56-198 (float64 IoU/offset,
 *   conf >= conf_thres, w = s*exp(-(1-IoU)^2/sigma) over the top-3000 pool, clip to frame_w x frame_h, drop empty);
 *   out rows (x,y,w,h,conf,cls); sigma <= 0 disables the refinement.
 * pred fp32 [B][4+nc][A] (xywh-centre boxes, scores); out fp32 [B][max_det][6]; out_count int32 [B]; out_idx int32
 * [B][max_det] anchor indices; scratch: ym_nms_scratch_bytes(B, A).  Images with more than 16384 candidates set the
 * overflow flag (ym_nms_overflowed) and return count 0 instead of silently truncating. */
long long ym_nms_scratch_bytes(int B, int A);

/* Large-candidate path of mode 0 (same outputs as ym_nms_batched): keys, boxes and suppression flags in global scratch, any number
 * of candidates per image, the first max_nms by score enter the suppression (utils/nms.py:142-146: validation at conf 0.001).
 * utils/nms.py takes it when ym_nms_overflowed() reports more than 16384 candidates in some image. */
long long ym_nms_large_scratch_bytes(int B, int A);
int ym_nms_batched_large(const float* pred, int B, int nc, int A, float conf_thres, float iou_thres, int max_det, int max_nms,
                         float max_wh, float* out, int* out_count, int* out_idx, void* scratch, void* stream);
int ym_nms_batched(const float* pred, int B, int nc, int A, float conf_thres, float iou_thres, int max_det, int max_nms,
                   float max_wh, int mode, float sigma, float frame_w, float frame_h, float* out, int* out_count, int* out_idx,
                   void* scratch, void* stream);
int ym_nms_overflowed(const void* scratch, int B, int A, void* stream);

/* tcgen05 path (TMEM accumulators, swizzled smem operands, single-thread MMA issue).
 * ym_tc_gemm_nt: out[M,N] = act(A[M,K] B[N,K]^T + bias) (+res) — the 1x1 Conv of conv.py:69-89 (a 1x1 conv on NHWC is
 *   exactly this GEMM with A = activation rows, B = packed weights [Cout][K]).
 * ym_moe_dispatch_tc: BatchedExpertComputation.compute_sparse_experts_batched  moe/utils.py:119-209 for 1x1-conv experts:
 *   out[b] = clamp(sum_j fp16(fp16(x[b] W[idx[b,j]]^T) * w[b,j]), +-clamp), routes with w <= w_min dropped (:172-173);
 *   x [B*HW][ldx] fp16, w_all [E][N][ldw] fp16, idx int32 [B,topk], w fp32 [B,topk], topk <= 2, N <= 256.
 *   No gather / scatter copies: every 128-token tile is read once, both routed experts accumulate in TMEM. */
int ym_tc_gemm_nt(const void* a, int lda, const void* b, int ldb, const float* bias, const void* res, int ldr, void* out,
                  int ldo, int M, int N, int K, int act, void* stream);
int ym_moe_dispatch_tc(const void* x, int ldx, int B, int HW, int C, const void* w_all, int ldw, long long w_expert_stride,
                       const int* route_idx, const float* route_w, int topk, int N, float w_min, float clamp, void* out,
                       int ldo, void* stream);
/* Same contract, persistent warp-specialised kernel: x tiles and expert weight tiles arrive by TMA (weights [E*N][ldw]
 * contiguous per expert), both routed experts accumulate in TMEM, output leaves by TMA store.  Shapes: HW % 128 == 0,
 * C % 64 == 0 <= 256, N % 128 == 0 <= 256, top_k <= 2 (ym_moe_dispatch_v2_supported). */
int ym_moe_dispatch_v2_supported(int HW, int C, int N, int topk, int ldx, int ldw, int ldo);
int ym_moe_dispatch_v2(const void* x, int ldx, int B, int HW, int C, const void* w_all, int ldw, int E, const int* route_idx,
                       const float* route_w, int topk, int N, float w_min, float clamp, void* out, int ldo, void* stream);
/* Same contract, CTA-pair kernel (2-CTA cluster, tcgen05.mma.cta_group::2, M = 256): each SM of a TPC pair keeps its own
 * 128-token tile and loads only half of every expert weight tile, halving the L2 -> shared-memory weight traffic per token.
 * Additionally requires HW % 256 == 0 and a device that can co-schedule 2-CTA clusters (ym_moe_dispatch_v3_supported). */
int ym_moe_dispatch_v3_supported(int HW, int C, int N, int topk, int ldx, int ldw, int ldo);
int ym_moe_dispatch_v3(const void* x, int ldx, int B, int HW, int C, const void* w_all, int ldw, int E, const int* route_idx,
                       const float* route_w, int topk, int N, float w_min, float clamp, void* out, int ldo, void* stream);
/* Profiling aid only (tools/profile_dispatch.py): a non-zero mask switches pipeline stages of the v2 kernel off so that
 * their cost can be read off the kernel time; results are then wrong.  Default 0; the package never sets it. */
void ym_set_dispatch_debug(int mask);
int ym_dispatch_debug_mask(void);
void ym_set_dispatch_trace(void* buf);   /* device int64[2][4][256]: clock64 timeline of CTAs 0/1 of the pair kernel, or NULL */

/* TMA + tcgen05 convolution (same contract and weight packing as ym_conv2d_nhwc): the activation k-tiles are loaded by
 * cp.async.bulk.tensor from a 4-D (C,W,H,B) tensor map at the tap-shifted origin (im2col-free, OOB zero fill = padding),
 * MMAs are tcgen05.mma with the accumulator in TMEM.  Supported: k in {1,3} square, stride 1|2, pad k/2, Cin % 16 == 0,
 * Cout % 8 == 0 (ym_conv2d_tc_supported returns 1); everything else goes through ym_conv2d_nhwc. */
int ym_conv2d_tc_supported(int Cin, int Cout, int KH, int KW, int stride, int pad, int ldx);
int ym_set_tc_conv_version(int v);   /* 2 (default): persistent, warp-specialised, TMA-store epilogue; 1: one tile per CTA */
int ym_conv2d_tc(const void* x, int ldx, int B, int H, int W, int Cin, const void* w, int Kpad, const float* bias, int Cout,
                 int KH, int KW, int stride, int pad, void* out, int ldo, int out_f32, const void* res, int ldr, int act,
                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YM_B200_H */
