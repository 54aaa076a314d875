/*
 * nanograd_b200 — C ABI of the sm_100a device backend for nanograd's tensor hot path.
 *
 * The reference (PABannier/nanograd) has no FFI: its device backend is PyOpenCL kernels
 * embedded as strings in nanograd/nn/ops_gpu.py.  Every entry point below replaces one of
 * those OpenCL launchers (or a composition of primitives that the reference runs as many
 * launches); the citation next to each prototype names the reference code it stands in for.
 * The Python binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; ng_last_error() then holds a
 *     human-readable message (the Python bridge raises Exception, matching the reference's
 *     plain-Exception error style, tensor.py:165,201,223).
 *   - device pointers are passed as uint64_t; all tensors are contiguous row-major float32
 *     (reference: nn/buffer.py:17 — GPUBuffer.dtype is always np.float32).
 *   - layouts are the reference's: activations NCHW, filters KCRS (ops_cpu.py:350-351).
 *   - all work is enqueued on one in-order compute stream per process (reference: one
 *     in-order OpenCL queue, tensor.py:30).  Only ng_d2h / ng_sync / ng_event_sync block.
 *   - plain C types only: no torch, no C++ types in signatures.
 */
#ifndef NANOGRAD_B200_H
#define NANOGRAD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ runtime -------- */
/* replaces get_gpu_context_and_queue(), tensor.py:16-30 */
int ng_init(int device_ordinal);
const char* ng_last_error(void);
int ng_device_info(int* sm_count, int* cc_major, int* cc_minor, uint64_t* total_mem_bytes);
int ng_sync(void);
/* number of kernels this library has launched since ng_init (bench.py's gpu_launches) */
int ng_launch_count(uint64_t* out);

/* replaces cl.Buffer(...) in GPUBuffer.__init__, nn/buffer.py:22-23.  Stream-ordered,
 * size-bucketed caching pool: freed blocks are reused by later same-size requests. */
int ng_malloc(uint64_t* out_ptr, uint64_t bytes);
int ng_free(uint64_t ptr);
int ng_mem_stats(uint64_t* live_bytes, uint64_t* pooled_bytes, uint64_t* n_device_allocs);
int ng_pool_release(void);
int ng_memset_zero(uint64_t ptr, uint64_t bytes);
/* replaces cl.enqueue_copy host<->device, tensor.py:171,180 */
int ng_h2d(uint64_t dst, const void* src, uint64_t bytes);
int ng_d2h(void* dst, uint64_t src, uint64_t bytes); /* blocking, like is_blocking=True */
int ng_d2d(uint64_t dst, uint64_t src, uint64_t bytes);
/* input pipeline: copy batch i+1 on a second stream while step i computes.
 * slot in [0,4) names the done-event of one staging buffer; the copy is ordered after everything the
 * compute stream was given before the call (the last readers of a double-buffered destination). */
int ng_h2d_prefetch(int slot, uint64_t dst, const void* pinned_src, uint64_t bytes);
int ng_prefetch_wait(int slot);   /* compute stream waits for the slot's copy */
int ng_prefetch_sync(int slot);   /* host waits for the slot's copy (pinned source reusable) */
/* Batch loader owned by the library (a host thread that never touches the Python interpreter): for every
 * step t it gathers rows plan[offsets[t] .. offsets[t+1]) of the host dataset X (and labels Y, float32,
 * may be NULL) into pinned memory and prefetches them into the caller's double-buffered device tensors
 * x_dev[t & 1] / y_dev[t & 1].  ng_loader_next returns the slot of the next batch once its copy is
 * ordered before the compute stream. */
int ng_loader_create(uint64_t* out_handle, const void* X, const void* Y, int64_t n_rows, uint64_t row_bytes,
                     int batch, int steps, const int64_t* plan, const int64_t* offsets,
                     uint64_t x_dev0, uint64_t x_dev1, uint64_t y_dev0, uint64_t y_dev1);
int ng_loader_next(uint64_t handle, int* out_slot, int64_t* out_rows);
int ng_loader_destroy(uint64_t handle);
/* host-side row gather into (pinned) staging memory: dst[i] = src[idx[i]], rows of row_bytes bytes */
int ng_host_gather_rows(void* dst, const void* src, const int64_t* idx, int64_t n, uint64_t row_bytes,
                        int64_t n_src_rows);
int ng_pinned_alloc(void** out, uint64_t bytes);
int ng_pinned_free(void* p);
/* CUDA-event timers on the compute stream */
/* D2H of a result whose producing kernel was enqueued before `ready_event` was recorded: runs on a read-back stream
 * that waits for that event only (reading a step's loss does not drain the work queued behind it).  Replaces the
 * blocking cl.enqueue_copy of Tensor.cpu(), tensor.py:167-180, for the loss read of a training loop. */
int ng_d2h_after(void* dst, uint64_t src, uint64_t bytes, uint64_t ready_event);
int ng_event_create(uint64_t* out_ev);
int ng_event_record(uint64_t ev);
int ng_event_sync(uint64_t ev);
int ng_event_elapsed_ms(uint64_t ev_start, uint64_t ev_stop, float* out_ms);
int ng_event_destroy(uint64_t ev);
/* writes a scratch buffer larger than the 126 MB L2 (timing hygiene between iterations) */
int ng_flush_l2(void);
/* FFMA-saturating microkernel: measured FP32 SIMT peak (TFLOP/s) for the roofline */
int ng_measure_ffma_peak(int iters, float* out_tflops);
/* per-kernel-class CUDA-event timing on the compute stream (bench.py's roofline objects): total_ms /
 * calls / algorithmic flops and bytes per class.  Contractions: classes 0-3.  Memory-bound classes:
 * STAGE = the channels-last staging passes of the tensor-core operands (with whatever is fused into them),
 * BN = BatchNorm statistics / normalise / backward passes, POOL = pooling forward / backward, OPTIM = the
 * multi-tensor optimizer step. */
enum { NG_PROF_CONV_FPROP = 0, NG_PROF_CONV_DGRAD = 1, NG_PROF_CONV_WGRAD = 2, NG_PROF_GEMM = 3,
       NG_PROF_STAGE = 4, NG_PROF_BN = 5, NG_PROF_POOL = 6, NG_PROF_OPTIM = 7 };
int ng_prof_enable(int on);
int ng_prof_reset(void);
int ng_prof_read(int kernel_class, double* total_ms, int64_t* launches, double* flops, double* bytes);

/* --------------------------------------------------------------- elementwise ------- */
/* unary op codes — replaces unary_op(), ops_gpu.py:79-85,127-131 */
enum {
  NG_U_COPY = 0, NG_U_NEG = 1, NG_U_EXP = 2, NG_U_LOG = 3, NG_U_RELU = 4,
  NG_U_SIGMOID = 5, NG_U_TANH = 6, NG_U_SQRT = 7, NG_U_RECIP = 8
};
int ng_unary(int op, uint64_t out, uint64_t a, int64_t n);
/* binary op codes — replaces element_wise_binary_op(), ops_gpu.py:51-77,114-125.
 * out = f(a, b) with NumPy broadcasting expressed as per-operand element strides
 * (stride 0 on broadcast axes) over the output shape. */
enum {
  NG_B_ADD = 0, NG_B_MUL = 1, NG_B_POW = 2, NG_B_DIV = 3, NG_B_SUB = 4,
  NG_B_EQ = 5,          /* (a == b) ? 1 : 0                 Max/Min backward mask, ops_cpu.py:125 */
  NG_B_RELU_BWD = 6,    /* a * (b >= 0)                     ops_cpu.py:263, ops_gpu.py:530 */
  NG_B_EXP_BWD = 7,     /* a * exp(b)                       ops_cpu.py:228 */
  NG_B_LOG_BWD = 8,     /* a / b                            ops_cpu.py:216 */
  NG_B_SIGMOID_BWD = 9, /* a * s(b) * (1 - s(b))            ops_cpu.py:275 */
  NG_B_TANH_BWD = 10,   /* a * (1 - tanh(b)^2)              ops_cpu.py:287 */
  NG_B_POW_DBASE = 11,  /* b * a^(b-1)                      ops_cpu.py:250 */
  NG_B_POW_DEXP = 12    /* a^b * log(a)                     ops_cpu.py:251 */
};
int ng_binary(int op, uint64_t out, uint64_t a, uint64_t b, int ndim,
              const int64_t* out_shape, const int64_t* a_strides, const int64_t* b_strides);
/* scalar immediates (the reference wraps every Python scalar in a 1-element device tensor,
 * tensor.py:482-483; here the value rides in the launch arguments).  scalar_is_lhs selects
 * f(s, a) instead of f(a, s). */
int ng_binary_scalar(int op, uint64_t out, uint64_t a, float s, int scalar_is_lhs, int64_t n);
/* Fused activations the reference composes from primitives:
 *   NG_ACT_LEAKY_RELU  relu(x) - alpha*relu(-x)          nn.LeakyReLU, module.py:744-772 (param = alpha)
 *   NG_ACT_MISH        x * tanh(log(1 + exp(x)))         nn.Mish,      module.py:844-864
 *   NG_ACT_SWISH       x * sigmoid(beta*x)               nn.Swish,     module.py:818-841 (param = beta)
 * param_ptr (device, 1 float) overrides the immediate `param` when non-zero.  ng_act_bwd writes
 * dx = dy * act'(x); ng_swish_dbeta writes the learnt beta's gradient sum(dy * x^2 * s(1-s)). */
enum { NG_ACT_LEAKY_RELU = 0, NG_ACT_MISH = 1, NG_ACT_SWISH = 2 };
int ng_act_fwd(int kind, uint64_t y, uint64_t x, uint64_t param_ptr, float param, int64_t n);
int ng_act_bwd(int kind, uint64_t dx, uint64_t dy, uint64_t x, uint64_t param_ptr, float param, int64_t n);
int ng_swish_dbeta(uint64_t dbeta1, uint64_t dy, uint64_t x, uint64_t beta_ptr, int64_t n);
int ng_fill(uint64_t out, float value, int64_t n);
/* out[i] += a[i]  (gradient accumulation, tensor.py:243) */
int ng_accumulate(uint64_t out, uint64_t a, int64_t n);

/* ----------------------------------------------------------------- reductions ------ */
/* replaces reduce_op(), ops_gpu.py:87-112,133-155 (serial loop per output there). */
enum { NG_R_SUM = 0, NG_R_MAX = 1, NG_R_MIN = 2 };
int ng_reduce(int op, uint64_t out, uint64_t in, int ndim, const int64_t* shape,
              const int32_t* reduce_axis_mask);
/* out[c] = sum over n, p of in[n, c, p] — the bias-gradient reduction that Add.backward's
 * unbroadcast performs (ops_cpu.py:12-18,174-176; module.py:476) */
int ng_channel_sum(uint64_t out, uint64_t in, int64_t n_outer, int64_t channels, int64_t inner);

/* ----------------------------------------------------------------- shape ops ------- */
/* replaces inner_slice()/gslice, ops_cpu.py:20-33, ops_gpu.py:187-208: N-d slice, zero fill
 * outside the source (this is also how padding works, tensor.py:408-429). */
int ng_slice(uint64_t out, uint64_t in, int ndim, const int64_t* in_shape,
             const int64_t* out_shape, const int64_t* start);
/* replaces perm_axis_op, ops_gpu.py:162-181 */
int ng_permute(uint64_t out, uint64_t in, int ndim, const int64_t* in_shape, const int32_t* perm);
/* replaces one_hot_encoding, ops_gpu.py:234-248 (labels are float32-encoded integers) */
int ng_onehot(uint64_t out, uint64_t labels, int64_t rows, int64_t classes);

/* ------------------------------------------------------------ GEMM / convolution --- */
/* flags for the contraction kernels */
/*   NG_F_TF32    convolutions on tcgen05 tensor cores with TF32 inputs (opt-in, rtol 1e-2)
 *   NG_F_TF32X3  convolutions on tcgen05 with the error-compensated 3xTF32 split: FP32-accurate
 *                (meets the FP32 parity bar rtol 1e-4).  Shapes a tensor-core kernel does not take
 *                run on the FP32 SIMT kernel. */
/*   NG_F_TC_FORCE  take the tensor-core kernel whenever the shape allows; without it NG_F_TF32X3 keeps
 *                problems below ~0.2 GFLOP on the SIMT kernel (staging + extra launches cost more
 *                than they save there). */
/*   NG_F_X_NHWC / NG_F_DY_NHWC  the x (fprop, wgrad) / dy (dgrad, wgrad) argument is already the staged
 *                channels-last copy made by ng_conv2d_stage_nhwc with the same flags: forward and
 *                backward of one layer then share two staging passes instead of running four.  Only
 *                legal when ng_conv2d_tc_plan says that pass takes the tensor-core kernel. */
/*   NG_F_F16X3  FP32-accurate convolutions on tcgen05 with operands staged once per tensor as scaled
 *                (hi, lo) FP16 pairs (11 + 11 significant bits, the precision of 3xTF32) and kind::f16 MMAs
 *                at twice the TF32 rate; taken for stride-1 layers whose channel counts are both multiples
 *                of 64, other layers follow NG_F_TF32X3 when that is set as well.
 *   NG_F_STAGE_F16  (ng_conv2d_stage_nhwc*) make the FP16x3 staged copy: dst holds N*C*HW*4 + 64 bytes. */
/*   NG_F_W_STAGED (FP16x3 engine only) the w argument of ng_conv2d_fprop / ng_conv2d_dgrad is the staged filter copy
 *                made by ng_conv2d_stage_filters_f16 for that pass. */
enum { NG_F_RELU = 1, NG_F_TF32 = 2, NG_F_TF32X3 = 4, NG_F_TC_FORCE = 8, NG_F_X_NHWC = 16, NG_F_DY_NHWC = 32,
       NG_F_F16X3 = 64, NG_F_STAGE_F16 = 128, NG_F_W_STAGED = 256 };
/* C[M,N] = op(A) @ op(B) (+ bias[N]) ; row-major; op = transpose when the flag is set.
 * replaces matmul_op, ops_gpu.py:214-232, and the perm+matmul chain in MatMul.backward,
 * ops_gpu.py:452-459.  batch > 1: equally strided batches of A, B and C.
 * flags: NG_F_RELU, and the math-mode flags below (tensor-core GEMM for unbatched problems). */
int ng_gemm(uint64_t c, uint64_t a, uint64_t b, int64_t M, int64_t N, int64_t K,
            int trans_a, int trans_b, int64_t batch, uint64_t bias, int flags);
/* y[N,K,OH,OW] = conv(x[N,C,H,W], w[K,C,R,S]) (+ bias[K]) ; cross-correlation, one stride for
 * both axes, zero padding pad_t/pad_l (bottom/right padding is implied by OH/OW).
 * replaces conv2d_op, ops_gpu.py:343-374 (ops_cpu.py:345-375) plus the pad2d slice copy
 * (tensor.py:419-429) and the bias Add (module.py:476).  Conv1d = H=R=1 (ops_gpu.py:308-333). */
int ng_conv2d_fprop(uint64_t y, uint64_t x, uint64_t w, uint64_t bias,
                    int N, int C, int H, int W, int K, int R, int S,
                    int stride, int pad_t, int pad_l, int OH, int OW, int flags);
/* dx = dgrad(dy, w): replaces conv_backward_x, ops_gpu.py:769-795 (ops_cpu.py:388-398) */
int ng_conv2d_dgrad(uint64_t dx, uint64_t dy, uint64_t w,
                    int N, int C, int H, int W, int K, int R, int S,
                    int stride, int pad_t, int pad_l, int OH, int OW, int flags);
/* dw = wgrad(dy, x): replaces conv_backward_weight, ops_gpu.py:797-820 (ops_cpu.py:386);
 * deterministic split-K (two passes, no atomics) */
int ng_conv2d_wgrad(uint64_t dw, uint64_t dy, uint64_t x,
                    int N, int C, int H, int W, int K, int R, int S,
                    int stride, int pad_t, int pad_l, int OH, int OW, int flags);

/* Which of fprop / dgrad / wgrad (out[0..2]) will run on the tensor-core kernels for this geometry
 * and these flags (so a caller knows whether staging an NHWC copy pays off), and out[3] = the engine:
 * 0 none, 1 TF32 family (FP32 NHWC staging), 2 FP16x3 (stage with NG_F_STAGE_F16). */
int ng_conv2d_tc_plan(int N, int C, int H, int W, int K, int R, int S, int stride, int OH, int OW,
                      int flags, int* out_fprop_dgrad_wgrad_engine);
/* dst[n][p][c] = src[n][c][p] (p = h*W + w); rounds to TF32 when flags has NG_F_TF32.  With
 * NG_F_STAGE_F16 (C % 64 == 0): dst[n][p][c / 64][hi | lo][c % 64] in FP16, scaled by a power of two,
 * followed by 64 bytes of metadata (float 1/scale, uint32 bits of max|src|). */
int ng_conv2d_stage_nhwc(uint64_t dst, uint64_t src, int64_t N, int64_t C, int64_t HW, int flags);
/* The same copy for dy that also returns the bias gradient sums[c] = sum_{n,p} src[n][c][p]
 * (Add.backward's unbroadcast of the conv bias, ops_cpu.py:12-18, module.py:476) from the one read. */
int ng_conv2d_stage_nhwc_sum(uint64_t dst, uint64_t src, uint64_t sums, int64_t N, int64_t C, int64_t HW, int flags);

/* FP16x3 staging with what the producer of the tensor already knows (no separate magnitude pass, and for
 * BatchNorm neighbours no separate normalise pass either):
 *   ng_conv2d_stage_f16      like NG_F_STAGE_F16 staging; sums (optional) as ng_conv2d_stage_nhwc_sum; absmax_bits
 *                            (optional) = device word with the bits of max|src| or an upper bound of it.
 *   ng_conv2d_stage_f16_bn   staged copy of y = act(BatchNorm2d(x)) recomputed from x and the batch statistics
 *                            (ng_bn2d_act_fwd_stats): y itself is never written.
 *   ng_conv2d_stage_f16_bnbwd staged copy (and channel sums) of the BatchNorm2d input gradient recomputed from
 *                            (dy, x) and the backward sums (ng_bn2d_act_bwd_sums): dx itself is never written.
 * These replace, for the conv -> BatchNorm2d -> ReLU -> conv chains of module.py:356-391 / 476, the second
 * half of the ~30 primitive ops the reference composes per BatchNorm layer. */
int ng_conv2d_stage_f16(uint64_t dst, uint64_t src, uint64_t sums, uint64_t absmax_bits, int64_t N, int64_t C, int64_t HW);
/* Both FP16x3 filter layouts of a KCRS tensor w[K][C][RS] from one pass (the weights of a step serve its forward and
 * its backward): dst_fprop = [hi|lo][rs][k][c], dst_dgrad (optional) = [hi|lo][rs][c][k], each K*C*RS*4 + 64 bytes. */
int ng_conv2d_stage_filters_f16(uint64_t dst_fprop, uint64_t dst_dgrad, uint64_t w, int K, int C, int RS);
int ng_conv2d_stage_f16_bn(uint64_t dst, uint64_t x, uint64_t gamma, uint64_t beta, uint64_t mean, uint64_t invstd,
                           uint64_t absmax_bits, int64_t N, int64_t C, int64_t HW, int act);
int ng_conv2d_stage_f16_bnbwd(uint64_t dst, uint64_t chan_sums, uint64_t dy, uint64_t x, uint64_t gamma, uint64_t beta,
                              uint64_t mean, uint64_t invstd, uint64_t bn_sums, uint64_t bound_bits, int64_t N,
                              int64_t C, int64_t HW, int act);
/* the same when the layer's output went through a 2x2 max pooling (conv -> BatchNorm2d -> ReLU -> MaxPool2d(2)):
 * the pooling backward (Max.backward's tie-splitting mask, ops_cpu.py:121-127) is fused in as well, dy is recomputed
 * per window from the pooled maxima and their gradient (sums from ng_bn2d_act_bwd_sums_pool) */
int ng_conv2d_stage_f16_bnbwd_pool(uint64_t dst, uint64_t chan_sums, uint64_t dpooled, uint64_t pooled, uint64_t x,
                                   uint64_t gamma, uint64_t beta, uint64_t mean, uint64_t invstd, uint64_t bn_sums,
                                   uint64_t bound_bits, int64_t N, int64_t C, int H, int W, int act);

/* -------------------------------------------------------------------- pooling ------ */
/* fused replacements for Tensor._pool2d → Slice, Reshape, Max/Sum(axis=(3,5)), Mul
 * (tensor.py:370-406) and Max.backward's tie-splitting mask (ops_cpu.py:121-127).
 * Window = stride = (ph, pw); the H % ph / W % pw remainder is cropped. */
int ng_maxpool2d_fwd(uint64_t y, uint64_t x, int64_t NC, int H, int W, int ph, int pw);
int ng_maxpool2d_bwd(uint64_t dx, uint64_t dy, uint64_t x, uint64_t y,
                     int64_t NC, int H, int W, int ph, int pw);
int ng_avgpool2d_fwd(uint64_t y, uint64_t x, int64_t NC, int H, int W, int ph, int pw);
int ng_avgpool2d_bwd(uint64_t dx, uint64_t dy, int64_t NC, int H, int W, int ph, int pw);
/* Max pooling of y = act(BatchNorm2d(x)) that was never written (statistics from ng_bn2d_act_fwd_stats): x is the
 * BatchNorm INPUT, every value read is normalised on the fly (C channels, NC = N*C planes); maxima and tie masks
 * are bit-identical to pooling the materialised y.  Removes the write and the re-read of y in
 * conv -> BatchNorm2d -> ReLU -> MaxPool2d chains. */
int ng_maxpool2d_fwd_bn(uint64_t y, uint64_t x, uint64_t gamma, uint64_t beta, uint64_t mean, uint64_t invstd,
                        int64_t NC, int C, int H, int W, int ph, int pw, int act);
int ng_maxpool2d_bwd_bn(uint64_t dx, uint64_t dy, uint64_t x, uint64_t y, uint64_t gamma, uint64_t beta, uint64_t mean,
                        uint64_t invstd, int64_t NC, int C, int H, int W, int ph, int pw, int act);

/* ----------------------------------------------------------------- loss head ------- */
/* fused Tensor.log_softmax (tensor.py:327-331, nine primitive nodes in the reference) */
int ng_logsoftmax_fwd(uint64_t out, uint64_t in, int64_t rows, int64_t cols);
int ng_logsoftmax_bwd(uint64_t dx, uint64_t dy, uint64_t logp, int64_t rows, int64_t cols);
/* fused NLLLoss.forward = -(logp * onehot(labels)).mean() over rows*cols (module.py:684-686) */
int ng_nll_fwd(uint64_t loss1, uint64_t logp, uint64_t labels, int64_t rows, int64_t cols);
int ng_nll_bwd(uint64_t dlogp, uint64_t dloss1, uint64_t labels, int64_t rows, int64_t cols);
/* nn.MSELoss fused: loss1[0] = mean((pred - target)^2) (module.py:689-712; deterministic two-pass sum);
 * ng_mse_bwd writes dloss * 2 (pred - target) / n, negated when for_target != 0. */
int ng_mse_fwd(uint64_t loss1, uint64_t pred, uint64_t target, int64_t n);
int ng_mse_bwd(uint64_t d, uint64_t dloss1, uint64_t pred, uint64_t target, int64_t n, int for_target);

/* device-side accuracy (replaces the per-step D2H of the logits + host argmax, tests/helpers.py:310-313):
 * out[r] = first argmax of row r as float32 ; out1[0] = #{i : a[i] == b[i]} */
int ng_argmax_rows(uint64_t out, uint64_t in, int64_t rows, int64_t cols);
int ng_count_equal(uint64_t out1, uint64_t a, uint64_t b, int64_t n);

/* ------------------------------------------------------------------ batch norm ----- */
/* fused BatchNorm2d.forward/_normalize (module.py:356-391): biased variance normalises,
 * unbiased variance feeds running_var, momentum update in place. */
int ng_bn2d_fwd_train(uint64_t y, uint64_t x, uint64_t gamma, uint64_t beta,
                      uint64_t running_mean, uint64_t running_var,
                      uint64_t save_mean, uint64_t save_invstd,
                      int64_t N, int64_t C, int64_t HW, float eps, float momentum);
int ng_bn2d_fwd_eval(uint64_t y, uint64_t x, uint64_t gamma, uint64_t beta,
                     uint64_t running_mean, uint64_t running_var,
                     int64_t N, int64_t C, int64_t HW, float eps);
int ng_bn2d_bwd(uint64_t dx, uint64_t dgamma, uint64_t dbeta, uint64_t dy, uint64_t x,
                uint64_t gamma, uint64_t save_mean, uint64_t save_invstd,
                int64_t N, int64_t C, int64_t HW);

/* BatchNorm2d followed by an activation in one pass (act: 0 = none, 1 = ReLU): the pair
 * BatchNorm2d -> ReLU of module.py:356-391 + ops_cpu.py:254-263.  The backward recomputes
 * z = (x - mean) * gamma * invstd + beta with the forward's exact expression and masks dy with
 * (z >= 0), the reference's ReLU subgradient (relu'(0) = 1). */
int ng_bn2d_act_fwd_train(uint64_t y, uint64_t x, uint64_t gamma, uint64_t beta,
                          uint64_t running_mean, uint64_t running_var,
                          uint64_t save_mean, uint64_t save_invstd,
                          int64_t N, int64_t C, int64_t HW, float eps, float momentum, int act);
int ng_bn2d_act_fwd_eval(uint64_t y, uint64_t x, uint64_t gamma, uint64_t beta,
                         uint64_t running_mean, uint64_t running_var,
                         int64_t N, int64_t C, int64_t HW, float eps, int act);
int ng_bn2d_act_bwd(uint64_t dx, uint64_t dgamma, uint64_t dbeta, uint64_t dy, uint64_t x,
                    uint64_t gamma, uint64_t beta, uint64_t save_mean, uint64_t save_invstd,
                    int64_t N, int64_t C, int64_t HW, int act);

/* The two halves of the fused layer as separate calls, so that the element-wise half can be fused into the
 * consumer (ng_conv2d_stage_f16_bn / _bnbwd) instead of writing a tensor nobody else reads:
 *   ng_bn2d_act_fwd_stats  statistics + running-stat update; absmax_bits[0] = bits of max|y| of the output the
 *                          apply pass WOULD write (the normalised output is monotone in x per channel, so the
 *                          channel extremes of x, tracked by the statistics pass, give it exactly); snapshot
 *                          (optional, 2C floats) receives copies of gamma and beta as they are at the call, so that
 *                          a normalise pass run after an in-place optimizer step still yields this step's output;
 *   ng_bn2d_act_apply      the normalise (+ activation) pass alone;
 *   ng_bn2d_act_bwd_sums   dgamma, dbeta, sums[2C + 1] = per channel {sum dy, sum dy (x - mean)} + the count,
 *                          bound_bits[0] = bits of an upper bound of max|dx|;
 *   ng_bn2d_act_bwd_dx     the dx pass alone.
 * Single process only (synchronised BatchNorm keeps the one-call forms above). */
int ng_bn2d_act_fwd_stats(uint64_t x, uint64_t gamma, uint64_t beta, uint64_t running_mean, uint64_t running_var,
                          uint64_t save_mean, uint64_t save_invstd, uint64_t absmax_bits, uint64_t snapshot,
                          int64_t N, int64_t C, int64_t HW, float eps, float momentum, int act);
int ng_bn2d_act_apply(uint64_t y, uint64_t x, uint64_t gamma, uint64_t beta, uint64_t save_mean, uint64_t save_invstd,
                      int64_t N, int64_t C, int64_t HW, int act);
int ng_bn2d_act_bwd_sums(uint64_t dgamma, uint64_t dbeta, uint64_t sums, uint64_t bound_bits, uint64_t dy, uint64_t x,
                         uint64_t gamma, uint64_t beta, uint64_t save_mean, uint64_t save_invstd,
                         int64_t N, int64_t C, int64_t HW, int act);
int ng_bn2d_act_bwd_dx(uint64_t dx, uint64_t dy, uint64_t x, uint64_t gamma, uint64_t beta, uint64_t save_mean,
                       uint64_t save_invstd, uint64_t sums, int64_t N, int64_t C, int64_t HW, int act);
/* ng_bn2d_act_bwd_dx that also returns chan_sums[c] = sum over (n, h, w) of dx: the bias gradient of the convolution
 * that produced x (Add.backward's unbroadcast, ops_cpu.py:12-18, module.py:476) without a second pass over dx */
int ng_bn2d_act_bwd_dx_sum(uint64_t dx, uint64_t chan_sums, uint64_t dy, uint64_t x, uint64_t gamma, uint64_t beta,
                           uint64_t save_mean, uint64_t save_invstd, uint64_t sums, int64_t N, int64_t C, int64_t HW,
                           int act);
/* ng_bn2d_act_bwd_sums for a layer whose output went through a 2x2 max pooling, with the pooling backward fused
 * in: the gradient of the pooled tensor (dpooled) and the pooled maxima replace dy, which is never written */
int ng_bn2d_act_bwd_sums_pool(uint64_t dgamma, uint64_t dbeta, uint64_t sums, uint64_t bound_bits, uint64_t dpooled,
                              uint64_t pooled, uint64_t x, uint64_t gamma, uint64_t beta, uint64_t save_mean,
                              uint64_t save_invstd, int64_t N, int64_t C, int H, int W, int act);

/* ------------------------------------------------------------------- optimizers ---- */
/* one multi-tensor launch per step for the whole parameter list.
 * SGD:   m = m*momentum + lr*(g*grad_scale); p -= m                (optimizer.py:52-55)
 * Adam:  m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
 *        p -= lr * (m/bc1) / (sqrt(v/bc2) + eps)                    (optimizer.py:79-89)
 * AdamW: p = p*(1 - lr*wd) - (lr/bc1) * m / (sqrt(v/bc2) + eps)     (optimizer.py:115-128)
 * bc1 = 1 - b1^t, bc2 = 1 - b2^t are computed by the caller.  Parameters are updated in
 * place (the reference rebinds p.data to a fresh buffer; values are identical). */
int ng_multi_sgd(int count, const uint64_t* params, const uint64_t* grads, const uint64_t* moms,
                 const int64_t* numels, float lr, float momentum, float grad_scale);
int ng_multi_adam(int count, const uint64_t* params, const uint64_t* grads,
                  const uint64_t* exp_avg, const uint64_t* exp_avg_sq, const int64_t* numels,
                  float lr, float beta1, float beta2, float eps, float bc1, float bc2,
                  float weight_decay, int adamw, float grad_scale);
/* gather `count` tensors into one flat bucket / scatter back (gradient bucket for allreduce) */
int ng_multi_pack(int count, const uint64_t* srcs, const int64_t* numels, uint64_t dst_flat);
int ng_multi_unpack(int count, const uint64_t* dsts, const int64_t* numels, uint64_t src_flat);

/* --------------------------------------------------------- captured steps (graphs) --- */
/* New functionality.  The reference's training loop (README.md:94-125) dispatches ~30 tiny device
 * operations per README-CNN step from Python; here a step can be recorded once on the compute stream
 * (ng_graph_begin ... the step's calls ... ng_graph_end) and replayed with one launch.  Memory allocated
 * while capturing belongs to the graph (private pool), so replays never alias later allocations.  A
 * captured step must not synchronise, read results back, or copy from pageable host memory. */
int ng_graph_begin(void);
int ng_graph_end(uint64_t* out_handle, uint64_t* out_kernel_nodes, uint64_t* out_private_bytes);
int ng_graph_launch(uint64_t handle);
int ng_graph_destroy(uint64_t handle);
int ng_graph_capturing(int* out_flag);

/* ---------------------------------------------------------- data-parallel (NCCL) --- */
/* New functionality (the reference has no multi-device support, device.py:10-11).  One
 * process per GPU; libnccl is dlopen()ed on first use.
 * ng_nccl_unique_id fills 256 bytes (two ncclUniqueIds: the gradient-bucket communicator on the
 * communication stream and the in-line communicator of the compute stream); rank 0 creates them
 * and shares them with the other ranks (nanograd_b200/dist.py: a file under the run directory).
 * ng_nccl_allreduce_begin launches the sum of buf[0..count) on the communication stream, ordered
 * after the compute-stream work enqueued so far; several chunks may be in flight while the backward
 * pass continues.  ng_nccl_allreduce_end makes the compute stream wait for all of them (before the
 * optimizer kernel).  ng_nccl_allreduce_sum = begin + end.
 * ng_nccl_sync_bn(1): BatchNorm training statistics and the backward's two per-channel sums are
 * exchanged across ranks (allgather of per-rank (mean, M2), Chan's combination; allreduce of the
 * backward sums), so a data-parallel step equals the single-process step on the global batch
 * (module.py:356-391 computes statistics over the whole batch).
 * ng_nccl_host_allreduce: up to 64 host doubles, op 0 = sum, 1 = max (barriers, timings, checksums). */
int ng_nccl_unique_id(void* out_256_bytes);
int ng_nccl_init(const void* id_256_bytes, int world_size, int rank);
int ng_nccl_world(int* world_size, int* rank);
int ng_nccl_allreduce_sum(uint64_t buf, int64_t count);
int ng_nccl_allreduce_begin(uint64_t buf, int64_t count);
int ng_nccl_allreduce_end(void);
int ng_nccl_sync_bn(int on);
int ng_nccl_host_allreduce(double* values, int count, int op);
int ng_nccl_destroy(void);

#ifdef __cplusplus
}
#endif
#endif /* NANOGRAD_B200_H */
