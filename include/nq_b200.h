/* nq_b200.h -- C ABI of libnq_b200.so: B200 (sm_100a) kernels for numpy-quant's
 * quantized-inference hot path.
 *
 * The reference (tebartsch/numpy-quant) is pure Python/NumPy and has NO FFI of
 * its own; each entry point below replaces one NumPy routine of the reference
 * (cited as file:line relative to the reference repository) and is what a
 * ctypes binding in numpy_quant/numpy_quantization.py / tensor.py / model.py
 * would call.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - sizes and strides are in ELEMENTS of the pointed-to type unless stated;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     all work is enqueued asynchronously on it, no hidden synchronisation;
 *   - return value 0 = success, non-zero = failure; nq_last_error() gives the
 *     message for the calling thread.  Nothing here aborts the process.
 *   - zero-points: `has_zp` = 0 means symmetric (the reference's `None`).
 *   - integer results are bit-exact w.r.t. the reference under NumPy >= 2
 *     promotion rules (SURVEY.md §8 "Numeric contract").
 */
#ifndef NQ_B200_H
#define NQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NQ_OK 0
#define NQ_ERR_INVALID 1
#define NQ_ERR_CUDA 2
#define NQ_ERR_UNSUPPORTED 3

int nq_version(void);
const char* nq_last_error(void);
/* Fills props_host[0..3] = {SM count, cc major, cc minor, max opt-in smem bytes}. */
int nq_device_info(int* props_host);

/* ---- K1: quantize  (numpy_quantization.py:24-34, tensor.py:227-229) -----------------
 * q = rint(clip(zp + x/scale, lo, hi)), x/scale an IEEE float32 division, the
 * zero-point add in float64, round-half-even.  Output int8 codes (bit_width 2..8). */
int nq_quantize_f32(const float* x, int64_t n, int bit_width, float scale, int has_zp, int64_t zp,
                    int8_t* out, void* stream);

/* Wide symmetric quantize to int64 codes: the 4*bit_width-bit Gemm / Add biases of
 * model.py:383-389 and 405-410 (bit_width up to 32; clip bounds rounded to float32 as
 * np.clip does for float32 data). */
int nq_quantize_f32_i64(const float* x, int64_t n, int bit_width, float scale, int has_zp, int64_t zp,
                        int64_t* out, void* stream);

/* Strided 4-D gather + quantize into a GEMM-ready K-major operand.
 * Logical input x[b0][b1][r][c] with element strides (s0,s1,sr,sc); output
 * out[(b0*d1+b1)][r][c] with row stride ldo (bytes, >= C) and batch stride
 * R*ldo.  Padding columns c in [C, ldo) are zero-filled.  If rowsum != NULL,
 * rowsum[(b0*d1+b1)*R + r] = sum_c q  (exact int32; the `arr.sum(axis)` terms of
 * numpy_quantization.py:52-60). */
int nq_quantize_f32_4d(const float* x, int64_t d0, int64_t d1, int64_t R, int64_t C,
                       int64_t s0, int64_t s1, int64_t sr, int64_t sc,
                       int bit_width, float scale, int has_zp, int64_t zp,
                       int8_t* out, int64_t ldo, int32_t* rowsum, void* stream);

/* ---- K2: dequantize  (numpy_quantization.py:37-41, tensor.py:189-193) ------------------
 * out = float32( float64(q - zp) * float64(scale) );  q is int8 / int32 / int64
 * (elem_bytes = 1, 4, 8). Scalar zero-point. */
int nq_dequantize(const void* q, int elem_bytes, int64_t n, float scale, int has_zp, int64_t zp,
                  float* out, void* stream);

/* Zero-point of a q_matmul accumulator, kept factored (numpy_quantization.py:49-61):
 *   zp[b,m,n] = rowsum_a[b,m]*zp_b + colsum_b[b,n]*zp_a - zp_a*zp_b*K
 * (terms dropped when the corresponding operand is symmetric). */
typedef struct nq_acc_zp {
    int has_zp_a, has_zp_b;
    int64_t zp_a, zp_b;
    int64_t k;                     /* contraction length (a.shape[-1])                */
    const int32_t* rowsum_a;       /* [batch*M]; required iff has_zp_b                */
    const int32_t* colsum_b;       /* [batchB*N]; required iff has_zp_a               */
    int64_t colsum_batch_stride;   /* N, or 0 when B is shared across the batch       */
} nq_acc_zp;

/* Dequantize an int32 accumulator [batch, M, N] (row stride ldacc) with the factored
 * zero-point above; output float32 [batch, M, N] contiguous. */
int nq_dequantize_acc(const int32_t* acc, int64_t batch, int64_t M, int64_t N, int64_t ldacc,
                      float scale, const nq_acc_zp* zp, float* out, void* stream);

/* ---- K3: requantize  (numpy_quantization.py:64-72, tensor.py:195-199, model.py:544-548) --
 * q = clip(rint(zp_out + (1/s_out) * dequantize(acc + bias_q)), lo, hi) -> int8.
 * bias_q (int64[N], may be NULL) is the Gemm bias added in the integer domain
 * (tensor.py:183-187). */
int nq_requantize_acc(const int32_t* acc, int64_t batch, int64_t M, int64_t N, int64_t ldacc,
                      float scale, const nq_acc_zp* zp, const int64_t* bias_q,
                      int out_bits, float out_scale, int has_out_zp, int64_t out_zp,
                      int8_t* out, void* stream);

/* Tail of requantize on an already dequantized float32 tensor:
 * q = clip(rint(zp_out + (1/s_out) * d), lo, hi)  (numpy_quantization.py:68-71). */
int nq_requantize_f32(const float* d, int64_t n, int out_bits, float out_scale, int has_out_zp,
                      int64_t out_zp, int8_t* out, void* stream);

/* rowsum[r] = sum_{c<C} q[r*ld + c]  (int32). Used for rowsum(A) and colsum(B) of
 * K-major operands (numpy_quantization.py:52,55,58-59). */
int nq_rowsum_s8(const int8_t* q, int64_t rows, int64_t C, int64_t ld, int32_t* rowsum, void* stream);

/* ---- K4/K5: integer GEMM on tcgen05 (numpy_quantization.py:44-61, tensor.py:205-210) ----
 * C[b] = A[b] (M x K, K-major, row stride lda bytes) * B[b]^T (N x K, K-major, row
 * stride ldb bytes); exact int32 accumulation on the tensor cores (kind::i8),
 * operands fetched by TMA (lda, ldb, batch strides multiples of 16 bytes, bases
 * 16-byte aligned; K itself is arbitrary -- the tail is zero-filled by TMA).
 * batch strides in elements; stride_b = 0 shares B across the batch.
 * Epilogue modes:
 *   NQ_EPI_RAW      C int32 [batch, M, ldc]
 *   NQ_EPI_DEQUANT  C float32 = dequantize(acc, scale, zp) (+ bias_f32[n]) (+ residual)  (K2 fused;
 *                   model.py:528-538 followed by the bias Add / residual Add of the graph)
 *   NQ_EPI_REQUANT  C int8 = requantize(acc + bias_q[n])                         (K3 fused) */
#define NQ_EPI_RAW 0
#define NQ_EPI_DEQUANT 1
#define NQ_EPI_REQUANT 2
#define NQ_EPI_SOFTMAX_QUANT 4   /* attention scores: softmax(dequant / sm_div) over each row (N <= 224), quantized
                                    with out_scale/out_zp/out_bits; C is the int8 operand [batch, M, ldc] of the
                                    following P.V MatMul, q_rowsum[batch * M] receives the code sums (plain stores) */
#define NQ_EPI_GELU_QUANT 5   /* NQ_EPI_QUANT with the GELU chain (model.py:65-213 Div/Erf/Add/Mul/Mul nodes, erf of
                                 numpy_helper.py:95-112) applied to bias + dequant before the quantization: the
                                 first MLP GEMM writes the second one's int8 operand (row layout only) */
#define NQ_EPI_QUANT 3   /* float result (dequant + bias) quantized for, and scattered into the K-major int8
                            operand of, the NEXT MatMul -- removes the float32 round trip (see q_* fields) */

typedef struct nq_epilogue {
    int mode;
    float scale;                   /* float32(scale_a * scale_b)                       */
    nq_acc_zp zp;
    const float* bias_f32;         /* [N] or NULL (DEQUANT)                            */
    const int64_t* bias_q;         /* [N] or NULL (REQUANT)                            */
    int out_bits;                  /* REQUANT                                          */
    float out_scale;
    int has_out_zp;
    int64_t out_zp;
    const float* residual;         /* [batch, M, N] float32 or NULL (DEQUANT): the graph's residual Add,
                                      out = (bias + dequant) + residual                 */
    int64_t ld_residual;           /* row stride of residual (elements)                */
    int64_t stride_residual;       /* batch stride of residual (elements)              */
    int64_t c_batch_inner;         /* > 1: the batch index is (outer, inner) and C[b] starts at
                                      (b / inner) * stride_c + (b % inner) * stride_c_inner -- lets the
                                      attention context GEMM write [B, S, H, D] directly (the graph's
                                      Transpose(0,2,1,3) of a [B, H, S, D] result) */
    int64_t stride_c_inner;
    /* NQ_EPI_QUANT: code = quantize(bias + dequant; out_scale, out_zp, out_bits) (numpy_quantization.py:24-34
     * applied to the float value the graph would have produced); C is the int8 destination.  With
     * m = mb*q_rows_per_image + ms, n = nh*q_cols_per_head + nd and batch b = bo*c_batch_inner + bi, the
     * code goes to byte C[bo*q_off[0] + bi*q_off[1] + mb*q_off[2] + ms*q_off[3] + nh*q_off[4] + nd*q_off[5]]
     * and, if q_rowsum != NULL, is added to q_rowsum[same decomposition with q_rs[]] (int32; see q_rowsum_count):
     * q_rs[5] == 0 accumulates along n (row sums), q_rs[3] == 0 along m (column sums). */
    int64_t q_rows_per_image, q_cols_per_head;
    int64_t q_off[6], q_rs[6];
    int32_t* q_rowsum;
    int sm_has_div;                /* NQ_EPI_SOFTMAX_QUANT: divide the scores by sm_div first (graph Div node) */
    float sm_div;
    int64_t q_rowsum_count;        /* number of int32 slots behind q_rowsum (NQ_EPI_QUANT / NQ_EPI_GELU_QUANT): the
                                      library zeroes them itself when partial sums have to meet through atomics */
    float gelu_div, gelu_add, gelu_mul;   /* NQ_EPI_GELU_QUANT: constants c1, c2, c3 of the graph's
                                             Div(x, c1) -> Erf -> Add(., c2) -> Mul(x, .) -> Mul(., c3) chain */
    int reverse_tiles;             /* != 0: walk the output tiles from the last to the first (same results): the tail of
                                      an A operand written front to back by the previous kernel is still in L2 */
} nq_epilogue;

/* ---- fused quantized attention (one kernel per layer; replaces, for every image and head, the chain
 * MatMul(Q, K^T) -> dequantize -> Div -> Softmax -> quantize -> MatMul(P, V) -> dequantize -> Transpose(0,2,1,3) ->
 * Reshape -> quantize of the reference's interpreter, model.py:486-565 with tensor.py:139-146 / 205-210).
 * Q  [BH, S, ld_q]  int8 K-major (row = query, contraction D)          -- left operand of the score MatMul
 * Kt [BH, S, ld_k]  int8 K-major (row = key,   contraction D)          -- right operand (K^T), stored transposed
 * Vt [BH, D, ld_v]  int8 K-major (row = head-dim column, contraction S) -- right operand of P.V, stored transposed
 * out [B, S, H*D] int8: the left operand of the output projection; out_rowsum [B*S] int32 or NULL.
 * Scores, probabilities and the context accumulator stay in TMEM / shared memory.  The zero-point terms of both
 * MatMuls (numpy_quantization.py:49-61) are accumulated by the tensor core itself (constant-operand passes), so no
 * row / column sums are passed in; softmax is float glue (1e-5 contract); GIVEN the emitted P codes the context
 * accumulator and the output codes are bit-exact.  Limits: S <= 208, D <= 64, D % 16 == 0, |zq|, |zv| <= 254,
 * lo_p - zp_p in [-254, 254]; anything else returns NQ_ERR_INVALID (callers use NQ_EPI_SOFTMAX_QUANT + NQ_EPI_QUANT). */
typedef struct nq_attention {
    float scale_qk;                /* float32(s_q * s_k)                                  */
    int has_div;                   /* graph Div node in front of the Softmax              */
    float div;
    int has_zq, has_zk;
    int64_t zq, zk;
    int p_bits;                    /* quantization of the probabilities                   */
    float p_scale;
    int has_p_zp;
    int64_t p_zp;
    float scale_pv;                /* float32(s_p * s_v)                                  */
    int has_zv;
    int64_t zv;
    int out_bits;                  /* quantization of the context (next MatMul's operand) */
    float out_scale;
    int has_out_zp;
    int64_t out_zp;
    int8_t* out;
    int32_t* out_rowsum;           /* cleared by the library                              */
    int8_t* p_dump;                /* optional: the emitted P bytes (code - lo_p, unsigned) as [BH, S, ld_p_dump] */
    int64_t ld_p_dump;             /* >= round_up(S, 8), multiple of 8                    */
} nq_attention;

int nq_attention_s8(const int8_t* Q, const int8_t* Kt, const int8_t* Vt, int64_t BH, int64_t H, int64_t S, int64_t D,
                    int64_t ld_q, int64_t ld_k, int64_t ld_v, const nq_attention* a, void* stream);

int nq_qgemm_s8(const int8_t* A, const int8_t* B, void* C,
                int64_t M, int64_t N, int64_t K, int64_t batch,
                int64_t lda, int64_t ldb, int64_t ldc,
                int64_t stride_a, int64_t stride_b, int64_t stride_c,
                const nq_epilogue* ep, void* stream);

/* Implicit-GEMM convolution (replaces the im2col + int matmul of the quantized Conv node: numpy_helper.py:18-92
 * extract_sliding_windows / conv2d, reached from model.py:95-100): the patch matrix is never materialised.
 * X: padded NHWC image codes [n_img][Hp][Wp][Cin] (pad pixels hold the activation zero-point code), Cin % 64 == 0;
 * Wm: filter matrix [O][KH*KW*Cin] in (kh, kw, c) order, row stride ldw bytes (% 16), symmetric codes.
 * out[(n, oh, ow), o] with OH = (Hp - KH) / stride_h + 1, OW = (Wp - KW) / stride_w + 1; the patch rows are read by
 * im2col-mode TMA straight into the MMA's shared-memory operand.  Epilogues: RAW, DEQUANT (+bias), REQUANT. */
int nq_qconv2d_s8(const int8_t* X, const int8_t* Wm, void* C, int64_t n_img, int64_t Hp, int64_t Wp, int64_t Cin,
                  int64_t KH, int64_t KW, int64_t stride_h, int64_t stride_w, int64_t O, int64_t ldw, int64_t ldc,
                  const nq_epilogue* ep, void* stream);

/* Plain CUDA-core integer GEMM with the same arguments (RAW epilogue only): the
 * on-device cross-check for the tensor-core kernel at sizes the CPU oracle cannot reach. */
int nq_qgemm_s8_simt(const int8_t* A, const int8_t* B, int32_t* C,
                     int64_t M, int64_t N, int64_t K, int64_t batch,
                     int64_t lda, int64_t ldb, int64_t ldc,
                     int64_t stride_a, int64_t stride_b, int64_t stride_c, void* stream);

/* NCHW -> padded NHWC relayout feeding nq_qconv2d_s8: x[B,C,H,W] int8 codes (elem_bytes 1) or float32 (elem_bytes 4:
 * quantized on the way with bits / scale / zp exactly like nq_quantize_f32) -> out[B][H+ph0+ph1][W+pw0+pw1][C] int8,
 * pad pixels = pad_code (the activation zero-point code: the quantized 0.0 of the reference's float padding,
 * numpy_helper.py:40-47).  C % 4 == 0. */
int nq_nhwc_pad(const void* x, int elem_bytes, int64_t B, int64_t C, int64_t H, int64_t W, int ph0, int pw0, int ph1, int pw1,
                int pad_code, int bits, float scale, int has_zp, int64_t zp, int8_t* out, void* stream);

/* Batched byte transpose out[b][c][r] = in[b][r][c] (b < batch, r < R, c < C; row strides ld_in / ld_out bytes, batch
 * strides in bytes; bytes r in [R, ld_out) of every output row are written as 0).  Turns the [head][S][D] codes the V
 * projection writes through the row-layout NQ_EPI_QUANT epilogue into the K-major [head][D][S] right operand of the
 * P.V MatMul (the graph's Transpose of V, tensor.py:74 / model.py Transpose node, as a 1-byte-per-element pass). */
int nq_transpose_s8(const int8_t* in, int64_t batch, int64_t R, int64_t C, int64_t ld_in, int64_t stride_in,
                    int8_t* out, int64_t ld_out, int64_t stride_out, void* stream);

/* Conv whose patches tile the image (kernel == stride, no padding -- the ViT patch embedding, reference
 * numpy_helper.py:18-92 reached from model.py:95-100): the patch matrix is a re-indexing of the image, so the input
 * quantizer (numpy_quantization.py:24-34) writes it directly: x[B,C,H,W] float32 ->
 * out[(b, oh, ow)][(c, kh, kw)] int8 (row stride ldo bytes), the K-major left operand of nq_qgemm_s8 against the
 * filters in their natural [O][C*KH*KW] order.  No nq_im2col pass.  KW % 4 == 0, H % KH == 0, W % KW == 0. */
int nq_quantize_patches_f32(const float* x, int64_t B, int64_t C, int64_t H, int64_t W, int64_t KH, int64_t KW,
                            int bit_width, float scale, int has_zp, int64_t zp, int8_t* out, int64_t ldo, void* stream);

/* ---- K6: im2col for Conv (numpy_helper.py:18-92, tensor.py:256-264) ----------------------
 * x[B,C,H,W] (int8 codes or float32; elem_bytes 1 or 4) -> patches
 * out[B*OH*OW][kh][kw][c] with row stride ldo elements; positions that fall in the
 * padding take `pad_value` (the zero-point in the integer domain, 0.0f bits for float).
 * pads = (ph0, pw0, ph1, pw1); OH = ceil((H-kh+ph0+ph1+1)/sh). */
int nq_im2col(const void* x, int elem_bytes, int64_t B, int64_t C, int64_t H, int64_t W,
              int kh, int kw, int ph0, int pw0, int ph1, int pw1, int sh, int sw,
              int32_t pad_value, void* out, int64_t ldo, void* stream);

/* ---- K11: sub-byte storage (no reference counterpart; contract unpack(pack(q)) == q) ------
 * n int8 codes of `bit_width` (2..8) bits <-> little-endian bitstream of
 * ceil(n*bit_width/8) bytes (two's-complement fields). */
int nq_pack_s8(const int8_t* q, int64_t n, int bit_width, uint8_t* packed, void* stream);
int nq_unpack_s8(const uint8_t* packed, int64_t n, int bit_width, int8_t* q, void* stream);

/* ---- K10: calibration statistics (model.py:332-336, tensor.py:232-236) -------------------
 * minmax[2*slot] = min(x), minmax[2*slot+1] = max(x). Slots must be initialised with
 * nq_minmax_init (+inf, -inf). */
int nq_minmax_init(float* minmax, int64_t n_slots, void* stream);
int nq_minmax_f32(const float* x, int64_t n, float* minmax, int64_t slot, void* stream);

/* ---- K7-K9: float32 glue ops of the fake-quant path (tensor.py:47-152, model.py:134-152,
 *      numpy_helper.py:95-112).  IEEE round-to-nearest, no FMA contraction, no fast-math. */
#define NQ_UN_NEG 0
#define NQ_UN_EXP 1
#define NQ_UN_ERF 2      /* A&S 7.1.26 polynomial of numpy_helper.erf, not libm erff */
#define NQ_UN_TANH 3
#define NQ_UN_SIGMOID 4  /* 1 / (1 + exp(-x))                                        */
#define NQ_UN_RELU 5     /* (x > 0) * x   (yields -0.0 for negative x)               */
#define NQ_UN_SQRT 6
#define NQ_UN_INV 7      /* 1 / x                                                    */
#define NQ_UN_COPY 8
int nq_unary_f32(int op, const float* x, int64_t n, float* out, void* stream);

#define NQ_BIN_ADD 0
#define NQ_BIN_MUL 1
#define NQ_BIN_DIV 2
/* out[i] = a[ia(i)] op b[ib(i)] over a contiguous 4-D output of dims d[4]; sa/sb are the
 * element strides of a and b per output dim (0 = broadcast). */
int nq_binary_f32(int op, const float* a, const int64_t* sa_host, const float* b, const int64_t* sb_host,
                  const int64_t* dims_host, float* out, void* stream);

/* GELU as spelled in the ViT graph: ((x / sqrt2) -> erf -> + 1) * x * 0.5, each step
 * rounded to float32 exactly like the five separate ONNX nodes. */
int nq_gelu_erf_f32(const float* x, int64_t n, float div_const, float add_const, float mul_const,
                    float* out, void* stream);

/* LayerNormalization over the last axis (model.py:134-152): two-pass mean / biased
 * variance, d * (1/sqrt(var+eps)) * gamma + beta. x rows have stride ldx. */
int nq_layernorm_f32(const float* x, int64_t rows, int64_t cols, int64_t ldx, const float* gamma,
                     const float* beta, float eps, float* out, void* stream);

/* Softmax over the last axis (tensor.py:139-146): exp(x - max) / sum. */
int nq_softmax_f32(const float* x, int64_t rows, int64_t cols, int64_t ldx, float* out, void* stream);

/* Softmax preceded by the graph's Div(x, c) (attention scores / sqrt(d)): exp(x/c - max) / sum. */
int nq_softmax_div_f32(const float* x, int64_t rows, int64_t cols, int64_t ldx, float div_const, float* out,
                       void* stream);

/* Producer -> quantize fusions (executor `retain=False`): the float op of the graph node(s)
 * and the quantize of the consuming MatMul (model.py:503-513) in one pass; the output is the
 * int8 K-major GEMM operand row (row stride ldo bytes, zero padded) plus optional row sums.
 * Same float32 roundings as running the two kernels back to back. */
int nq_layernorm_quantize_f32(const float* x, int64_t rows, int64_t cols, int64_t ldx, const float* gamma,
                              const float* beta, float eps, int bit_width, float scale, int has_zp, int64_t zp,
                              int8_t* out, int64_t ldo, int32_t* rowsum, int float_glue, void* stream);
/* float_glue bit 0: the normalised value is float glue under the 1e-5 contract (one FMA for gamma/beta, division
 * by the scale through its reciprocal); rounding, clamp and row sums stay exact.  Clear: the same float32 roundings as
 * nq_layernorm_f32 followed by nq_quantize_f32.
 * float_glue bit 1 (with bit 0): walk the rows from the last to the first.  Same results; a producer that wrote x
 * front to back immediately before this call left its last rows in L2 (x is larger than L2 at the batch sizes of
 * BASELINE.json), so the reverse walk reads them from there instead of from HBM. */
int nq_softmax_quantize_f32(const float* x, int64_t rows, int64_t cols, int64_t ldx, int has_div, float div_const,
                            int bit_width, float scale, int has_zp, int64_t zp,
                            int8_t* out, int64_t ldo, int32_t* rowsum, void* stream);
int nq_gelu_quantize_f32(const float* x, int64_t rows, int64_t cols, int64_t ldx, float div_const, float add_const,
                         float mul_const, int bit_width, float scale, int has_zp, int64_t zp,
                         int8_t* out, int64_t ldo, int32_t* rowsum, void* stream);

/* Row reductions over the last axis: op 0 = max, 1 = sum, 2 = mean (tensor.py:124-137). */
int nq_reduce_rows_f32(int op, const float* x, int64_t rows, int64_t cols, float* out, void* stream);

/* Strided 4-D copy into a contiguous tensor (Transpose / Expand / Slice / Concat pieces);
 * elem_bytes 1, 4 or 8. out_offset/ld describe a contiguous destination written at
 * out + out_offset with its own 4-D element strides so. */
int nq_copy_4d(const void* x, int elem_bytes, const int64_t* dims_host, const int64_t* sx_host,
               void* out, const int64_t* so_host, void* stream);

/* cudaMemsetAsync on `stream` (zeroing of row-sum accumulators that kernels add into). */
int nq_memset_async(void* ptr, int value, int64_t bytes, void* stream);

/* Device self-test: counts (into *mismatches_dev, uint64 on the device, caller-zeroed) the inputs among
 * n pseudo-random float pairs for which the kernels' hoisted-reciprocal IEEE division differs from
 * __fdiv_rn; mode 0 = raw bit patterns, 1 = activation / scale ranges.  Must stay 0. */
int nq_selftest_division(int64_t n, int seed, int mode, unsigned long long* mismatches_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NQ_B200_H */
