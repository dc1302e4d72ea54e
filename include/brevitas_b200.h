/*
 * brevitas_b200.h — C-ABI of the B200-native fake-quantization hot path.
 *
 * This is the drop-in boundary for the one native plugin of the reference (Giuseppe5/brevitas):
 * `src/brevitas/csrc/autograd_ste_ops.cpp`, whose 12 entry points are registered at
 * csrc/autograd_ste_ops.cpp:258-271 and consumed by `brevitas.function.ops_ste`
 * (src/brevitas/function/ops_ste.py:38-43, 65-67) — plus the fused quantizer entry points that replace the
 * ATen op chains issued by `brevitas.core.quant`, `brevitas.core.scaling` and `brevitas.core.stats`
 * (SURVEY.md §8a).  Every function takes raw DEVICE pointers, element counts, a dtype tag and a
 * `cudaStream_t` (passed as void*), returns an int status, never allocates, never synchronises and
 * keeps no global mutable state, so it is re-entrant, honours the caller's stream and is capturable
 * in CUDA graphs.  There is no CPU implementation behind this interface.
 *
 * All tensors are dense/contiguous.  "T" below is the element type selected by `dtype`.
 * Scalars that the reference carries in 0-dim tensors are passed as floats holding the value the
 * reference tensor holds.  The library applies ATen's (CPU, torch 2.11) rule for a 0-dim operand next to a
 * low-precision tensor: add / sub / compare operands (zero_point, qmin, qmax, scaling_min_val) are rounded
 * to T first; mul / div second operands (int_threshold, a one-element fp32 scale) keep their fp32 value in
 * the fp32 "opmath" and only the result is rounded to T.
 *
 * `scale_dtype`: dtype tag of the scale tensor.  It must equal `dtype`, except that a ONE-element scale may
 * be BVB_F32 while x is bf16/fp16 (fp32 quantizer modules fed low-precision activations).
 */
#ifndef BREVITAS_B200_H_
#define BREVITAS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums ------------------------------------------------------------------------------------------- */
enum { BVB_F32 = 0, BVB_BF16 = 1, BVB_F16 = 2 };                       /* dtype tags                */
enum { BVB_OK = 0, BVB_EINVAL = 1, BVB_EUNSUPPORTED = 2, BVB_ECUDA = 3 };/* status codes              */
/* float_to_int_impl flavours: RoundSte / FloorSte / CeilSte / RoundToZeroSte / DPURoundSte
 * (src/brevitas/core/function_wrapper/ops_ste.py:14-92) */
enum { BVB_ROUND = 0, BVB_FLOOR = 1, BVB_CEIL = 2, BVB_ROUND_TO_ZERO = 3, BVB_DPU_ROUND = 4 };
/* tensor_clamp_impl flavours: TensorClampSte (pass-through gradient, ops/autograd_ste_ops.py:118-120)
 * vs TensorClamp (torch.where autograd => masked gradient, function/ops.py:98-99) */
enum { BVB_CLAMP_STE = 0, BVB_CLAMP_MASKED = 1 };

/* ---- library ----------------------------------------------------------------------------------------- */
int bvb_version(void);
/* thread-local description of the last non-OK status returned on this thread */
const char* bvb_last_error(void);
/* number of SMs of the current device (cached per device), used for grid sizing */
int bvb_sm_count(void);
#ifdef BVB_TUNING_BUILD
/* NOT part of the product library: `make -C brevitas_b200/csrc TUNING=1` builds libbrevitas_b200_tuning.so, the
 * same kernels plus this process-wide launch-geometry override (0 = heuristic) for the sweeps under tools/.
 * libbrevitas_b200.so itself keeps no mutable state. */
void bvb_set_tuning(int rows_threads, int rows_stages, int rows_ctas_per_sm, int stream_threads, int stream_ctas_per_sm);
#endif

/* numerics self-test: the kernels divide by a row-/tensor-constant scale with nvcc's own div.rn.f32 instruction
 * sequence, its loop-invariant reciprocal refinement hoisted (csrc/common.cuh DivBy).  This entry point compares
 * that against the compiler's IEEE division for `count` consecutive numerator bit patterns starting at
 * `first_bits` and writes the number of bitwise mismatches to *mismatches (device pointer, uint64).           */
int bvb_selftest_div(float divisor, uint32_t first_bits, uint64_t count, uint64_t* mismatches, void* stream);
/* bf16 kernels divide a bf16 numerator by a bf16 scale as RN_bf16(a * RN_f32(1/b)); this enumerates ALL 2^16 x 2^16
 * (a, b) pairs of `dtype` (BVB_BF16 / BVB_F16) whose divisor lies in the shortcut's window and compares with
 * RN_T(a / b).  out2 (device, uint64[2]) = { mismatches, pairs checked }.  bf16: 0 mismatches (why the shortcut is
 * used); fp16: > 0 (why it is not).                                                                              */
int bvb_selftest_lowp_div(int dtype, uint64_t* out2, void* stream);

/* host-only introspection (no GPU needed): the constants the bf16 / fp16 kernels' packed fast path derives from an
 * integer range (csrc/common.cuh qdq_vec).  out7 (HOST pointer, uint32[7]) = { ok, lo_is_zero, lo, hi, lo_pre,
 * thr_lo, thr_hi } as 16-bit patterns of `dtype` duplicated in both halves: round(v) < qmin <=> v < thr_lo and
 * round(v) > qmax <=> v > thr_hi for every value v of the dtype.  ok = 0 means the kernels use the literal op
 * sequence for this range.                                                                                       */
int bvb_debug_packed_constants(float zero_point, float qmin, float qmax, int dtype, uint32_t* out7);

/* ---- 1. the 12 STE primitives: forward values of torch.ops.autograd_ste_ops.* -------------------------
 * Backward of every op except abs_binary_sign_grad is the identity on the incoming gradient and
 * needs no kernel (csrc/autograd_ste_ops.cpp:22, 42).  y may alias x (in-place).                       */
int bvb_round_ste_impl(const void* x, void* y, int64_t n, int dtype, void* stream);          /* csrc:14-24   torch.round  */
int bvb_ceil_ste_impl(const void* x, void* y, int64_t n, int dtype, void* stream);           /* csrc:100-110 torch.ceil   */
int bvb_floor_ste_impl(const void* x, void* y, int64_t n, int dtype, void* stream);          /* csrc:112-122 torch.floor  */
int bvb_binary_sign_ste_impl(const void* x, void* y, int64_t n, int dtype, void* stream);    /* csrc:124-137 (x>=0)-(x<0) */
int bvb_ternary_sign_ste_impl(const void* x, void* y, int64_t n, int dtype, void* stream);   /* csrc:140-150 torch.sign   */
int bvb_round_to_zero_ste_impl(const void* x, void* y, int64_t n, int dtype, void* stream);  /* csrc:153-163              */
int bvb_dpu_round_ste_impl(const void* x, void* y, int64_t n, int dtype, void* stream);      /* csrc:166-179              */
int bvb_abs_binary_sign_grad_impl(const void* x, void* y, int64_t n, int dtype, void* stream);/* csrc:182-194 torch.abs   */
/* backward of abs_binary_sign_grad: gx = binary_sign(x) * gy (ops/autograd_ste_ops.py:374-377) */
int bvb_abs_binary_sign_grad_bwd(const void* x, const void* gy, void* gx, int64_t n, int dtype, void* stream);
/* tensor_clamp_ste_impl (csrc:27-44) / tensor_clamp_ste_impl_ (csrc:47-63; Python semantics, in place:
 * pass y == x and inplace_minmax = 1 to get torch.min/torch.max NaN rules of function/ops.py:109-110).
 * min/max are broadcast as element i -> m[(i / inner) % count]  (count = 1: scalar tensor;
 * inner = 1, count = n: same shape; inner = cols, count = rows: per-row). */
int bvb_tensor_clamp_ste_impl(const void* x, const void* min_val, const void* max_val, void* y, int64_t n,
                              int64_t min_inner, int64_t min_count, int64_t max_inner, int64_t max_count,
                              int inplace_minmax, int dtype, void* stream);
/* backward of the DIFFERENTIABLE tensor_clamp (src/brevitas/function/ops.py:76-100, the default clamp of IntQuant and
 * what a learned bit-width reaches its gradient through): gx = gy where no branch replaced the value; gmin_out / gmax_out
 * (nullable, fp32[min_count] / fp32[max_count], zeroed here) = sums of gy over the elements replaced by min / by max */
int bvb_tensor_clamp_bwd(const void* gy, const void* x, const void* min_val, const void* max_val, void* gx,
                         float* gmin_out, float* gmax_out, int64_t n, int64_t min_inner, int64_t min_count,
                         int64_t max_inner, int64_t max_count, int dtype, void* stream);
int bvb_scalar_clamp_ste_impl(const void* x, void* y, int64_t n, double min_val, double max_val, int dtype, void* stream);   /* csrc:66-82 */
int bvb_scalar_clamp_min_ste_impl(const void* x, void* y, int64_t n, double min_val, int dtype, void* stream);               /* csrc:85-97 */

/* ---- 2. IntQuant with a provided scale (src/brevitas/core/quant/int_base.py:64-97) ----------------------
 * y = (clamp(round(x / s + zp), qmin, qmax) - zp) * s, element i uses s[(i / scale_inner) % scale_count].
 * codes_out (nullable, T) receives the integer codes of IntQuant.to_int.                                */
int bvb_int_quant_fwd(const void* x, const void* scale, void* y, void* codes_out, int64_t n,
                      int64_t scale_inner, int64_t scale_count, int scale_dtype,
                      float zero_point, float qmin, float qmax, int round_mode, int dtype, void* stream);
/* gx = ((gy * s) / s) * m  (m = 1 for BVB_CLAMP_STE, clamp mask on the rounded value for BVB_CLAMP_MASKED).
 * If gscale_out (fp32, scale_count entries, zero-filled by the callee) is non-null it receives
 * d(loss)/d(scale) = sum over each scale's region of gy*(q - zp) - m*(gy*s)*((x/s)/s)  (SURVEY.md A.4). */
int bvb_int_quant_bwd(const void* gy, const void* x, const void* scale, void* gx, float* gscale_out, int64_t n,
                      int64_t scale_inner, int64_t scale_count, int scale_dtype,
                      float zero_point, float qmin, float qmax, int round_mode, int clamp_mode,
                      int dtype, void* stream);

/* IntQuant with a TENSOR-valued zero-point (asymmetric quantizers: src/brevitas/core/zero_point.py:38-226, the
 * ShiftedUint8* quantizers of quant/shifted_scaled_int.py): zero_point has the broadcast pattern of scale (one element,
 * or scale_count elements).  Same arithmetic as bvb_int_quant_fwd / bwd; the backward additionally returns
 * d(loss)/d(zero_point) = sum over each region of m*(gy*s) - gy*s (fp32, zero-filled by the callee; pass both
 * gradient outputs or neither).                                                                                  */
int bvb_int_quant_zpt_fwd(const void* x, const void* scale, const void* zero_point, void* y, int64_t n,
                          int64_t scale_inner, int64_t scale_count, int scale_dtype, int zp_dtype,
                          float qmin, float qmax, int round_mode, int dtype, void* stream);
int bvb_int_quant_zpt_bwd(const void* gy, const void* x, const void* scale, const void* zero_point, void* gx,
                          float* gscale_out, float* gzp_out, int64_t n, int64_t scale_inner, int64_t scale_count,
                          int scale_dtype, int zp_dtype, float qmin, float qmax, int round_mode, int clamp_mode,
                          int dtype, void* stream);

/* Integer export: the codes clamp(round(x / scale + zero_point), qmin, qmax) of IntQuant.to_int (int_base.py:64-76) stored
 * in a REAL integer dtype -- what QuantTensor.int() (src/brevitas/quant_tensor/__init__.py:174-187) and the export
 * handlers produce with a round + cast over the dequantized value.  out_kind: BVB_OUT_I8 / BVB_OUT_U8 / BVB_OUT_I32; the
 * range must fit it.  Pass qmin / qmax = the dtype's own limits to convert an already quantized tensor (value / scale
 * + zero_point is integral there, the clamp is a no-op).  1 read of T + 1 write of 1 (4) bytes per element.            */
#define BVB_OUT_I8 0
#define BVB_OUT_U8 1
#define BVB_OUT_I32 2
int bvb_int_quant_to_int(const void* x, const void* scale, void* out, int64_t n, int64_t scale_inner, int64_t scale_count,
                         int scale_dtype, float zero_point, float qmin, float qmax, int round_mode, int out_kind,
                         int dtype, void* stream);

/* QuantReLU fused with its quantizer (FusedActivationQuantProxy: activation_impl = nn.ReLU, then tensor_quant;
 * src/brevitas/proxy/runtime_quant.py:73-84, nn/quant_activation.py:14-31): y = int_quant(relu(x)) in ONE pass, and
 * gx = int_quant_bwd(gy, relu(x)) * [not (x <= 0)] (ATen's threshold_backward) in one pass -- saves the ReLU's
 * read + write forward and its two reads + one write backward.  Same arguments as the two calls above.        */
int bvb_relu_int_quant_fwd(const void* x, const void* scale, void* y, void* codes_out, int64_t n,
                           int64_t scale_inner, int64_t scale_count, int scale_dtype,
                           float zero_point, float qmin, float qmax, int round_mode, int dtype, void* stream);
int bvb_relu_int_quant_bwd(const void* gy, const void* x, const void* scale, void* gx, float* gscale_out, int64_t n,
                           int64_t scale_inner, int64_t scale_count, int scale_dtype,
                           float zero_point, float qmin, float qmax, int round_mode, int clamp_mode,
                           int dtype, void* stream);

/* ---- 3. RescalingIntQuant with abs-max statistics, fused (src/brevitas/core/quant/int.py:156-163 over
 *         core/scaling/runtime.py:19-102, core/stats/stats_op.py:129-141) ----------------------------------
 * Per row of a [rows, cols] view (OverOutputChannelView / OverBatchOverOutputChannelView + AbsMax(dim)):
 *   absmax = max|x|;  thr = clamp_min(absmax, scaling_min_val) (skipped if scaling_min_val <= 0);
 *   s = thr / int_threshold;  y = quant-dequant(x, s).   Single HBM pass: 1 read + 1 write per element.
 * scale_out: [rows] T.  absmax_out: [rows] T, nullable (the un-clamped statistic, e.g. for running_stats). */
int bvb_rows_absmax_int_quant_fwd(const void* x, void* y, void* scale_out, void* absmax_out,
                                  int64_t rows, int64_t cols, float scaling_min_val, float int_threshold,
                                  float zero_point, float qmin, float qmax, int round_mode, int dtype, void* stream);
/* Backward with the gradient flowing through the statistic:
 *   gx = ((gy*s)/s)*m, then gx[row, argmax] += sign(x[argmax]) * (Gs[row] + gscale[row]) / int_threshold,
 *   Gs[row] = sum_j gy*(q - zp) - m*(gy*s)*((x/s)/s); argmax = first index of the row maximum.
 * gscale (nullable, [rows] T) is the gradient arriving on the returned scale.                             */
int bvb_rows_absmax_int_quant_bwd(const void* gy, const void* x, const void* scale, const void* gscale, void* gx,
                                  int64_t rows, int64_t cols, float int_threshold,
                                  float zero_point, float qmin, float qmax, int round_mode, int clamp_mode,
                                  int dtype, void* stream);
/* Host-buffer variant of the two calls above, for tensors that live in HOST memory (pinned for full speed): rows are
 * independent, so the weight is cut into chunks of `chunk_rows` rows and pipelined over three internal streams --
 * H2D of chunk c+1, fwd+bwd kernels of chunk c, D2H of chunk c-1 -- through a ring of device staging slots in
 * `workspace` (>= bvb_host_pipeline_workspace_bytes(...) bytes of device memory).  h_x, h_gy: inputs; h_gx
 * (d loss / d x incl. the path through the abs-max), h_scale ([rows] T): outputs; h_y (the quant-dequantized
 * weight) nullable.  The work is ordered after everything already enqueued on `stream`, and `stream` waits for the
 * last copy: synchronise `stream` before reading the outputs.  Creates / destroys its streams and events per call,
 * therefore NOT capturable in a CUDA graph.  Same arithmetic, same kernels, same results as the device calls.    */
int bvb_host_rows_fakequant_fwd_bwd(const void* h_x, const void* h_gy, void* h_y, void* h_gx, void* h_scale,
                                    int64_t rows, int64_t cols, int64_t chunk_rows, float scaling_min_val,
                                    float int_threshold, float zero_point, float qmin, float qmax, int round_mode,
                                    int clamp_mode, int dtype, void* workspace, int64_t workspace_bytes, void* stream);
int64_t bvb_host_pipeline_workspace_bytes(int64_t rows, int64_t cols, int64_t chunk_rows, int want_y, int dtype);
/* The call above creates and destroys its three streams and eleven events every time.  A caller that repeats it (an
 * optimizer loop over host-resident master weights) creates the set once, passes it to the `_on` form and destroys it at
 * the end; the handle belongs to the caller (one per device and per concurrent caller), the library keeps nothing. */
int bvb_host_pipeline_create(void** handle);
int bvb_host_pipeline_destroy(void* handle);
int bvb_host_rows_fakequant_fwd_bwd_on(void* pipeline, const void* h_x, const void* h_gy, void* h_y, void* h_gx,
                                       void* h_scale, int64_t rows, int64_t cols, int64_t chunk_rows,
                                       float scaling_min_val, float int_threshold, float zero_point, float qmin, float qmax,
                                       int round_mode, int clamp_mode, int dtype, void* workspace, int64_t workspace_bytes,
                                       void* stream);
/* Whole-tensor statistic (OverTensorView + AbsMax(None)): two-phase grid reduction, then quant pass.
 * workspace: >= bvb_workspace_bytes() bytes of device scratch.  absmax_out: 1 element of T; scale_out:
 * 1 element of scale_dtype (fp32 when a 0-dim T threshold is divided by a 0-dim fp32 int_threshold).      */
int bvb_tensor_absmax_int_quant_fwd(const void* x, void* y, void* scale_out, void* absmax_out, int64_t n,
                                    int scale_dtype, float scaling_min_val, float int_threshold,
                                    float zero_point, float qmin, float qmax, int round_mode, int dtype,
                                    void* workspace, void* stream);
/* Backward: torch.max() distributes the statistic's gradient evenly over all tied maxima. */
int bvb_tensor_absmax_int_quant_bwd(const void* gy, const void* x, const void* scale, const void* absmax,
                                    const void* gscale, void* gx, int64_t n, int scale_dtype, float int_threshold,
                                    float zero_point, float qmin, float qmax, int round_mode, int clamp_mode,
                                    int dtype, void* workspace, void* stream);
int64_t bvb_workspace_bytes(void);

/* ---- 4. BinaryQuant / ClampedBinaryQuant (src/brevitas/core/quant/binary.py:60-64, 120-125) -------------
 * y = binary_sign(x) * s;  clamped != 0: y = binary_sign(where-clamp(x, -s, s)) * s (same values for s >= 0;
 * the clamp shapes the gradient).                                                                        */
int bvb_binary_quant_fwd(const void* x, const void* scale, void* y, int64_t n,
                         int64_t scale_inner, int64_t scale_count, int scale_dtype, int clamped, int dtype,
                         void* stream);
/* gx = gy * s (plain) or gy * s * [!(x > s) & !(x < -s)] (clamped);
 * gscale_out (nullable, fp32[scale_count]) = sum gy*sign_b(clamp(x)) (+ gy*s where x > s, - gy*s where x < -s) */
int bvb_binary_quant_bwd(const void* gy, const void* x, const void* scale, void* gx, float* gscale_out, int64_t n,
                         int64_t scale_inner, int64_t scale_count, int scale_dtype, int clamped, int dtype,
                         void* stream);

/* Minimum AND maximum of every row with their positions in ONE read: torch.max(x, dim) + torch.min(x, dim) of AbsMinMax
 * (src/brevitas/core/stats/stats_op.py:144-158) and the torch.min of NegativeMinOrZero (stats_op.py:22-39) that the
 * asymmetric weight quantizers run over the same tensor.  Selection as ATen's CUDA reductions: NaN wins, then the value,
 * then the lowest index; the outputs are the selected elements themselves.  argmin_out / argmax_out (nullable): int64
 * positions inside the row.  workspace: bvb_minmax_workspace_bytes(rows) bytes of device memory.                  */
int64_t bvb_minmax_workspace_bytes(int64_t rows);
int bvb_minmax_rows(const void* x, void* min_out, void* max_out, int64_t* argmin_out, int64_t* argmax_out, int64_t rows,
                    int64_t cols, int dtype, void* workspace, void* stream);

/* ---- 4a. the remaining quantizer flavours, one kernel per direction ------------------------- */
/* General integer quantizer: DecoupledIntQuant.forward (src/brevitas/core/quant/int_base.py:132-182) and
 * IntQuant.forward (int_base.py:64-97) when the range must stay on the device -- a learned bit-width
 * (src/brevitas/core/bit_width/parameter.py:23-98), or a call that may not read anything back (CUDA-graph capture):
 *   c = where-clamp(float_to_int(x / pre_scale + pre_zero_point), min_int, max_int);  y = (c - zero_point) * scale
 * pre_scale / scale: dtype of x, pre_scale_count / scale_count elements each (1, or one shared broadcast pattern
 * scale[(i / scale_inner) % count]).  pre_zero_point, zero_point, min_int, max_int: ONE fp32 element each, in device
 * memory; the bounds are rounded to the tensor dtype like the reference's `.type_as(x)`.                           */
int bvb_general_int_quant_fwd(const void* x, const void* pre_scale, const void* scale, const float* pre_zero_point,
                              const float* zero_point, const float* min_int, const float* max_int, void* y, int64_t n,
                              int64_t scale_inner, int64_t pre_scale_count, int64_t scale_count, int round_mode,
                              int dtype, void* stream);
/* gx = ((gy * scale) * m) / pre_scale  (m = 1 for BVB_CLAMP_STE, else the where-clamp mask).  sums (nullable, fp64,
 * bvb_general_int_quant_sums() elements, zeroed here) = [ d pre_scale (pre_scale_count) | d scale (scale_count) |
 * d min_int | d max_int ]: d scale = sum gy * (c - zero_point), d pre_scale = - sum gx' * ((x / pre_scale) / pre_scale),
 * d min_int / d max_int = the gradient of the lanes clipped low / high (BVB_CLAMP_MASKED only).  same_scale != 0
 * (pre_scale IS scale, IntQuant.forward): the two scale sums are taken as one difference per element and the whole
 * d scale is returned in the second block.                                                                         */
int64_t bvb_general_int_quant_sums(int64_t pre_scale_count, int64_t scale_count);
int bvb_general_int_quant_bwd(const void* gy, const void* x, const void* pre_scale, const void* scale,
                              const float* pre_zero_point, const float* zero_point, const float* min_int,
                              const float* max_int, void* gx, double* sums, int64_t n, int64_t scale_inner,
                              int64_t pre_scale_count, int64_t scale_count, int round_mode, int clamp_mode,
                              int same_scale, int dtype, void* stream);
/* TernaryQuant.forward (src/brevitas/core/quant/ternary.py:58-72): y = float(|x| > threshold * s) * sign(x) * s with ONE
 * scale; fp32 only (the reference's result is fp32 whatever the input, BVB_EUNSUPPORTED otherwise).
 * Backward: gx = (gy * s) * mask; gscale (nullable, one fp64, zeroed here) = sum gy * mask * sign(x).              */
int bvb_ternary_quant_fwd(const void* x, const void* scale, void* y, int64_t n, float threshold, int dtype, void* stream);
int bvb_ternary_quant_bwd(const void* gy, const void* x, const void* scale, void* gx, double* gscale, int64_t n,
                          float threshold, int dtype, void* stream);

/* ---- 4b. batch-norm + ReLU + activation quantizer of a conv block, fused ------------------- */
/* `QuantReLU(BatchNorm2d(conv_out))` -- FusedActivationQuantProxy (src/brevitas/proxy/runtime_quant.py:73-84) behind
 * torch.nn.BatchNorm2d, e.g. brevitas_examples/imagenet_classification/models/mobilenetv1.py:111-115 -- on a
 * channels-last tensor seen as x[rows = N*H*W][channels]: 8 passes over the activation per training step instead of 13
 * (the normalised tensor and the ReLU output are never written).  Training mode (use_running_stats = 0): batch mean and
 * biased variance per channel (fp32 per-thread sums combined in fp64 in a fixed order), running statistics updated in
 * place with `momentum` and the unbiased variance; save_mean / save_invstd (fp32[channels]) are outputs for the
 * backward.  Eval mode (use_running_stats = 1): save_mean / save_invstd are INPUTS (running mean, 1/sqrt(running_var +
 * eps)).  y = quant_dequant(relu(((x - mean) * invstd) * gamma + beta)) with the provided scale: one element, or one per
 * channel (`scale_count` = 1 or channels).  gamma / beta: fp32[channels] or NULL (1 / 0).  round-half-even only.
 * `residual` (nullable, same shape and dtype as x): y = quant_dequant(relu(bn(x) + residual)), the closing block of a
 * ResNet BasicBlock (`relu(bn2(conv2(.)) + identity)`); the backward then also returns `gresidual` (nullable) = the
 * gradient of that input, which equals the gradient of the batch-norm output.
 * Requires channels * sizeof(T) / 16 to divide 256 (BVB_EUNSUPPORTED otherwise: callers use the unfused pair).
 * workspace >= bvb_bn_act_quant_workspace_bytes(channels) bytes of device scratch. */
int64_t bvb_bn_act_quant_workspace_bytes(int64_t channels);
int bvb_bn_act_quant_fwd(const void* x, const void* residual, const float* gamma, const float* beta, float* running_mean, float* running_var,
                         float momentum, float eps, int use_running_stats, const void* scale, int64_t scale_count,
                         int scale_dtype, void* y, float* save_mean, float* save_invstd, int64_t rows, int64_t channels,
                         float zero_point, float qmin, float qmax, int round_mode, int relu, int dtype, void* workspace,
                         void* stream);
/* backward of the above in training mode: gx = d/dx, ggamma / gbeta (fp32[channels]), gscale (nullable, fp32[scale_count])
 * = d/d(scale) of the quantizer; the quantizer part is bvb_relu_int_quant_bwd's arithmetic on the recomputed
 * normalised value (clamp_mode: BVB_CLAMP_STE / BVB_CLAMP_MASKED) */
int bvb_bn_act_quant_bwd(const void* gy, const void* x, const void* residual, void* gresidual, const float* gamma, const float* beta, const float* save_mean,
                         const float* save_invstd, const void* scale, int64_t scale_count, int scale_dtype, void* gx,
                         float* ggamma, float* gbeta, float* gscale, int64_t rows, int64_t channels, float zero_point,
                         float qmin, float qmax, int round_mode, int clamp_mode, int relu, int dtype, void* workspace,
                         void* stream);

/* ---- 5. statistics (src/brevitas/core/stats/stats_op.py) ----------------------------------------------- */
/* AbsMax(stats_reduce_dim=1) on a [rows, cols] view -> out[rows] (T); NaN-propagating (stats_op.py:137-141) */
int bvb_absmax_rows(const void* x, void* out, int64_t rows, int64_t cols, int dtype, void* stream);
/* AbsMax(None) -> out[1] (T) */
int bvb_absmax_tensor(const void* x, void* out, int64_t n, int dtype, void* workspace, void* stream);
/* AbsPercentile (stats_op.py:41-66): k-th smallest (1-indexed) of |x| over each row of a [rows, cols] view
 * (rows = 1: flat).  out[rows] (T); index_out (nullable, int64[rows]) = the SMALLEST index attaining it.
 * Exact radix select; workspace >= bvb_kth_workspace_bytes(rows) bytes of device scratch, contents irrelevant on
 * entry (histograms, a first-index table and, for rows <= 2, a 48 MiB per row buffer the surviving candidates are
 * copied to once few enough are left, so that a high percentile of an fp32 tensor costs two reads instead of four). */
int bvb_abs_kth_value_rows(const void* x, void* out, int64_t* index_out, int64_t rows, int64_t cols, int64_t k,
                           int dtype, void* workspace, void* stream);
/* AbsPercentile of relu(x) -- the statistic of a QuantReLU whose quantizer still collects (nn.ReLU then tensor_quant,
 * src/brevitas/proxy/runtime_quant.py:81-84; core/scaling/standalone.py:230-244) -- without a ReLU pass: keys of
 * max(x, +0).  Same contract and workspace as bvb_abs_kth_value_rows. */
int bvb_relu_abs_kth_value_rows(const void* x, void* out, int64_t* index_out, int64_t rows, int64_t cols, int64_t k,
                                int dtype, void* workspace, void* stream);
/* the SIGNED k-th smallest value, x.kthvalue(k) (NegativePercentileOrZero, PercentileInterval: stats_op.py:69-126),
 * same select on order-preserving keys of x itself (every NaN sorts last, like torch.kthvalue); same workspace */
int bvb_kth_value_rows(const void* x, void* out, int64_t* index_out, int64_t rows, int64_t cols, int64_t k,
                           int dtype, void* workspace, void* stream);
int64_t bvb_kth_workspace_bytes(int64_t rows);
/* _RuntimeStats EMA (stats_wrapper.py:56-65): first != 0: running *= stat; else
 * running = running * one_minus_momentum + (T)(momentum * stat).  The caller passes (float)(1.0 - m) computed
 * in double like the Python expression.  running: fp32 [count]; stat: T [count]. */
int bvb_running_stats_update(float* running, const void* stat, int64_t count, float momentum,
                             float one_minus_momentum, int first, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BREVITAS_B200_H_ */
