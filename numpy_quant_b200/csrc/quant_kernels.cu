// K1 quantize, K2 dequantize, K3 requantize, K10 min/max, K11 pack/unpack, row sums.
// All HBM-bound: 16-byte vector accesses, grid-stride over whole waves of the 148 SMs.
#include <algorithm>
#include "common.cuh"

namespace nq {

// ------------------------------------------------------------------ K1 contiguous
// Each thread-iteration: four fully coalesced 16-B loads (512 B per warp each) in flight, four
// coalesced 4-B stores of packed codes.
template <int QMODE>
__global__ void __launch_bounds__(256) quantize_contig_kernel(const float* __restrict__ x, int64_t n, QArgs a,
                                                             int8_t* __restrict__ out) {
    const int64_t n4 = n >> 2;                                  // float4 groups
    const float4* x4 = reinterpret_cast<const float4*>(x);
    int* o32 = reinterpret_cast<int*>(out);
    constexpr int U = 4;
    const int64_t tile = (int64_t)blockDim.x * U;
    const Quantizer qz(a);
    for (int64_t base = (int64_t)blockIdx.x * tile; base < n4; base += (int64_t)gridDim.x * tile) {
        float4 v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int64_t i = base + j * blockDim.x + threadIdx.x;
            v[j] = (i < n4) ? __ldcs(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int64_t i = base + j * blockDim.x + threadIdx.x;
            if (i < n4)
                __stcs(o32 + i, pack4_codes(qz.code<QMODE>(v[j].x),
                                            qz.code<QMODE>(v[j].y),
                                            qz.code<QMODE>(v[j].z),
                                            qz.code<QMODE>(v[j].w)));
        }
    }
    const int64_t t = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // tail (< 4 elements)
    if (t < n) out[t] = (int8_t)qz.code<QMODE>(x[t]);
}

template <int QMODE>
__global__ void __launch_bounds__(256) quantize_scalar_kernel(const float* __restrict__ x, int64_t n, QArgs a,
                                                             int8_t* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const Quantizer qz(a);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (int8_t)qz.code<QMODE>(x[i]);
}

// Non-overlapping patches (Conv with kernel == stride and no padding: the ViT stem): the patch matrix is a pure
// re-indexing of the image, so the quantizer writes it directly -- x[B, C, H, W] float32 -> out[(b, oh, ow)][(c, kh, kw)]
// int8, the K-major A operand of the convolution GEMM (weights in their natural [O][C][KH][KW] order).  One thread
// = one (patch, c, kh) run of KW floats: consecutive threads are consecutive kh / c of one patch, so a warp writes one
// contiguous span of the operand row and reads whole 32-byte sectors of the image.
template <int QMODE>
__global__ void __launch_bounds__(1024) quantize_patches_kernel(const float* __restrict__ x, int64_t n_rows, int C, int H, int W,
                                                               int KH, int KW, QArgs a, int8_t* __restrict__ out, int64_t ldo) {
    // CTA = one row of patches (b, oh); thread = one float4 slot (c, kh, j) of the patch, fixed for the whole launch, so
    // the loops below carry no index arithmetic: along ow the source advances by KW floats and the destination by one
    // operand row.  Consecutive threads read consecutive 16 bytes of a KW-float run, then the next kh / c, and write
    // consecutive 4-byte groups of the operand row; four patches in flight per thread.
    const Quantizer qz(a);
    const int OH = H / KH, OW = W / KW, kw4 = KW >> 2, vec_per_patch = C * KH * kw4;
    for (int t = threadIdx.x; t < vec_per_patch; t += blockDim.x) {
        const int ck = t / kw4, j = t - ck * kw4, c = ck / KH, kh = ck - c * KH;
        for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
            const int64_t b = row / OH;
            const int oh = (int)(row - b * OH);
            const float4* src = reinterpret_cast<const float4*>(x + ((b * C + c) * H + (int64_t)oh * KH + kh) * W) + j;
            int* dst = reinterpret_cast<int*>(out + row * OW * ldo + (int64_t)ck * KW) + j;
            const int64_t dstep = ldo >> 2;
            for (int ow0 = 0; ow0 < OW; ow0 += 4) {
                float4 val[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) val[u] = __ldcs(src + (int64_t)min(ow0 + u, OW - 1) * kw4);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (ow0 + u < OW)
                        dst[(ow0 + u) * dstep] = pack4_codes(qz.code<QMODE>(val[u].x), qz.code<QMODE>(val[u].y),
                                                            qz.code<QMODE>(val[u].z), qz.code<QMODE>(val[u].w));
            }
        }
    }
}

// Batched byte transpose out[b][c][r] = in[b][r][c] (the V operand of the attention: the projection GEMM writes V
// as [head][S][D] through its fast row-layout epilogue, the P.V MatMul wants the contraction axis S contiguous).
// Tile 128 r x 64 c through shared memory: 16-byte coalesced loads, 4 x 4 byte blocks transposed in registers
// (8 PRMT), 4-byte stores that are contiguous along r across a warp.  Word columns are XOR-swizzled with the row
// group so that the column-wise read-back has at most 2-way bank conflicts.  Rows r in [R, ld_out) are written as 0.
__global__ void __launch_bounds__(256) transpose_s8_kernel(const int8_t* __restrict__ in, int R, int Cc, int64_t ld_in,
                                                          int64_t stride_in, int8_t* __restrict__ out, int64_t ld_out,
                                                          int64_t stride_out, int r_tiles, int c_tiles) {
    __shared__ uint32_t tile[128 * 16];
    const int ct = blockIdx.x % c_tiles, rt = (blockIdx.x / c_tiles) % r_tiles;
    const int64_t b = blockIdx.x / (c_tiles * r_tiles);
    const int r0 = rt * 128, c0 = ct * 64;
    const int8_t* src = in + b * stride_in;
    for (int i = threadIdx.x; i < 128 * 4; i += 256) {
        const int r = i >> 2, q4 = i & 3, c = c0 + q4 * 16;
        int4 v = make_int4(0, 0, 0, 0);
        if (r0 + r < R && c < Cc) {
            const int8_t* p = src + (int64_t)(r0 + r) * ld_in + c;
            if (c + 16 <= Cc && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) v = *reinterpret_cast<const int4*>(p);
            else {
                int w[4] = {0, 0, 0, 0};
                for (int k = 0; k < 16 && c + k < Cc; ++k) w[k >> 2] |= (int)(uint8_t)p[k] << ((k & 3) * 8);
                v = make_int4(w[0], w[1], w[2], w[3]);
            }
        }
        const int sw = (r >> 2) & 15;
        uint32_t* row = tile + r * 16;
        row[(q4 * 4 + 0) ^ sw] = (uint32_t)v.x;
        row[(q4 * 4 + 1) ^ sw] = (uint32_t)v.y;
        row[(q4 * 4 + 2) ^ sw] = (uint32_t)v.z;
        row[(q4 * 4 + 3) ^ sw] = (uint32_t)v.w;
    }
    __syncthreads();
    int8_t* dst = out + b * stride_out;
    for (int i = threadIdx.x; i < 32 * 16; i += 256) {
        const int sg = i & 31, dg = i >> 5;                                // lanes: consecutive row groups
        const int sw = sg & 15;
        const uint32_t a0 = tile[(sg * 4 + 0) * 16 + (dg ^ sw)], a1 = tile[(sg * 4 + 1) * 16 + (dg ^ sw)];
        const uint32_t a2 = tile[(sg * 4 + 2) * 16 + (dg ^ sw)], a3 = tile[(sg * 4 + 3) * 16 + (dg ^ sw)];
        const uint32_t t0 = __byte_perm(a0, a1, 0x5140), t1 = __byte_perm(a2, a3, 0x5140);
        const uint32_t t2 = __byte_perm(a0, a1, 0x7362), t3 = __byte_perm(a2, a3, 0x7362);
        const uint32_t o[4] = {__byte_perm(t0, t1, 0x5410), __byte_perm(t0, t1, 0x7632), __byte_perm(t2, t3, 0x5410),
                               __byte_perm(t2, t3, 0x7632)};
        const int r = r0 + sg * 4;
        if (r >= ld_out) continue;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = c0 + dg * 4 + k;
            if (c < Cc) *reinterpret_cast<uint32_t*>(dst + (int64_t)c * ld_out + r) = o[k];
        }
    }
}

// wide symmetric/asymmetric quantize to int64 codes (4*bit_width-bit biases, model.py:383-389, 405-410)
__global__ void quantize_i64_kernel(const float* __restrict__ x, int64_t n, float scale, int has_zp, double zp,
                                    float lo, float hi, double dlo, double dhi, int64_t* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float t = __fdiv_rn(x[i], scale);
        if (has_zp) {
            // int64 zero-point + float32 quotient -> float64 (NEP 50); np.clip keeps the float64 bounds exact
            double u = fmin(fmax(zp + (double)t, dlo), dhi);
            out[i] = __double2ll_rn(u);
        } else {
            t = fminf(fmaxf(t, lo), hi);          // lo/hi already rounded to float32 like np.clip does
            out[i] = __float2ll_rn(t);
        }
    }
}

// ------------------------------------------------------------------ K1 strided 4-D -> K-major operand
// `lpr` lanes (8 / 16 / 32) cooperate on one output row (b, r) and sweep its C (=K) axis, so reads
// are coalesced when sc == 1, short rows (attention heads: C = 64) still fill the warp, and the
// row sum falls out of a dp4a + sub-warp shuffle reduction.
template <int QMODE>
__global__ void __launch_bounds__(256) quantize_rows_kernel(const float* __restrict__ x, uint32_t d1, uint32_t R,
                                                           int64_t C, int64_t s0, int64_t s1, int64_t sr,
                                                           int64_t sc, uint32_t rows_total, QArgs a,
                                                           int8_t* __restrict__ out, int64_t ldo,
                                                           int32_t* __restrict__ rowsum, int lpr) {
    const int lane = threadIdx.x & 31;
    const int sub = lane & (lpr - 1), rsel = lane / lpr, rpw = 32 / lpr;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const Quantizer qz(a);
    for (uint32_t row0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * rpw; row0 < rows_total; row0 += warps * rpw) {
        const uint32_t row = row0 + rsel;
        const bool valid = row < rows_total;
        int sum = 0;
        if (valid) {
            const uint32_t r = row % R, b = row / R;
            const float* src = x + (int64_t)(b / d1) * s0 + (int64_t)(b % d1) * s1 + (int64_t)r * sr;
            int8_t* dst = out + (int64_t)row * ldo;
            if (sc == 1 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (ldo & 3) == 0) {
                const int c4 = (int)(C >> 2);
                // four 16-byte loads per lane in flight before the first use (long rows were one load deep: latency-bound)
                int c = sub;
                for (; c + 3 * lpr < c4; c += 4 * lpr) {
                    float4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(src) + c + u * lpr);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int w = pack4_codes(qz.code<QMODE>(v[u].x), qz.code<QMODE>(v[u].y), qz.code<QMODE>(v[u].z), qz.code<QMODE>(v[u].w));
                        sum = __dp4a(w, 0x01010101, sum);
                        __stcs(reinterpret_cast<int*>(dst) + c + u * lpr, w);
                    }
                }
                for (; c < c4; c += lpr) {
                    const float4 v = __ldcs(reinterpret_cast<const float4*>(src) + c);
                    const int w = pack4_codes(qz.code<QMODE>(v.x), qz.code<QMODE>(v.y), qz.code<QMODE>(v.z), qz.code<QMODE>(v.w));
                    sum = __dp4a(w, 0x01010101, sum);
                    reinterpret_cast<int*>(dst)[c] = w;
                }
                for (int64_t c = ((int64_t)c4 << 2) + sub; c < ldo; c += lpr) {
                    const int q = c < C ? (int)(int8_t)qz.code<QMODE>(src[c]) : 0;
                    sum += q;
                    dst[c] = (int8_t)q;
                }
            } else {
                for (int64_t c = sub; c < ldo; c += lpr) {
                    const int q = c < C ? (int)(int8_t)qz.code<QMODE>(src[c * sc]) : 0;
                    sum += q;
                    dst[c] = (int8_t)q;
                }
            }
        }
        if (rowsum) {
            for (int o = lpr >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (sub == 0 && valid) rowsum[row] = sum;
        }
    }
}

// Transposing variant: the input is contiguous along r (sr == 1), e.g. a [K, N] weight or V
// read as rows = N, cols = K.  32x32 tiles through shared memory keep both sides coalesced.
template <int QMODE>
__global__ void __launch_bounds__(256) quantize_transpose_kernel(const float* __restrict__ x, int64_t d1, int64_t R,
                                                                int64_t C, int64_t s0, int64_t s1, int64_t sc,
                                                                QArgs a, int8_t* __restrict__ out, int64_t ldo,
                                                                int32_t* __restrict__ rowsum) {
    const Quantizer qz(a);
    __shared__ int8_t tile[32][36];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
    const int64_t b = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    const float* src = x + (b / d1) * s0 + (b % d1) * s1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t c = c0 + ty + j * 8, r = r0 + tx;               // lanes along r (unit stride)
        int q = 0;
        if (c < C && r < R) q = (int)(int8_t)qz.code<QMODE>(src[c * sc + r]);
        tile[ty + j * 8][tx] = (int8_t)q;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t r = r0 + ty + j * 8, c = c0 + tx;               // lanes along c for the store
        const int q = tile[tx][ty + j * 8];
        if (r < R && c < ldo) out[(b * R + r) * ldo + c] = (int8_t)q;
        if (rowsum) {
            int s = q;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (tx == 0 && r < R) atomicAdd(rowsum + b * R + r, s);
        }
    }
}

// ------------------------------------------------------------------ K2
template <typename T>
__global__ void __launch_bounds__(256) dequantize_kernel(const T* __restrict__ q, int64_t n, float scale, int64_t zp,
                                                        float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = dequantize_one((int64_t)q[i] - zp, scale);
}

// int8 codes -> float32: one 32-bit load (4 codes) and one 16-byte store per thread-iteration,
// four iterations in flight; both sides fully coalesced.
__global__ void __launch_bounds__(256) dequantize_s8x4_kernel(const int8_t* __restrict__ q, int64_t n4, float scale,
                                                             int zp, float* __restrict__ out) {
    const int* q32 = reinterpret_cast<const int*>(q);
    float4* o4 = reinterpret_cast<float4*>(out);
    const int64_t tile = (int64_t)blockDim.x * 4;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < n4; base += (int64_t)gridDim.x * tile) {
        int w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t i = base + j * blockDim.x + threadIdx.x;
            w[j] = (i < n4) ? __ldcs(q32 + i) : 0;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t i = base + j * blockDim.x + threadIdx.x;
            if (i < n4) {
                float4 f;
                f.x = __fmul_rn((float)((int)(int8_t)(w[j]) - zp), scale);
                f.y = __fmul_rn((float)((int)(int8_t)(w[j] >> 8) - zp), scale);
                f.z = __fmul_rn((float)((int)(int8_t)(w[j] >> 16) - zp), scale);
                f.w = __fmul_rn((float)((int)(int8_t)(w[j] >> 24) - zp), scale);
                __stcs(o4 + i, f);
            }
        }
    }
}

__device__ __forceinline__ int64_t acc_zero_point(const AccZp& z, int64_t b, int64_t m, int64_t n, int64_t M) {
    int64_t v = -z.kterm;
    if (z.use_row) v += (int64_t)z.rowsum_a[b * M + m] * z.zp_b;
    if (z.use_col) v += (int64_t)z.colsum_b[b * z.cs_stride + n] * z.zp_a;
    return v;
}

// acc [batch, M, N] (ld) -> f32 / int8; one thread per 4 consecutive n when aligned.
template <int MODE, bool ASYM_OUT>   // MODE 1 dequant -> f32, 2 requant -> s8
__global__ void __launch_bounds__(256) acc_post_kernel(const int32_t* __restrict__ acc, int64_t batch, int64_t M,
                                                      int64_t N, int64_t ldacc, float scale, AccZp z,
                                                      const int64_t* __restrict__ bias_q, float inv_out_scale,
                                                      double out_zp, float lo, float hi, void* __restrict__ out) {
    const int64_t total = batch * M * N;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t n = i % N;
        const int64_t bm = i / N;
        const int64_t m = bm % M, b = bm / M;
        int64_t a = acc[bm * ldacc + n];
        if (MODE == 2 && bias_q) a += bias_q[n];
        const float d = dequantize_one(a - acc_zero_point(z, b, m, n, M), scale);
        if (MODE == 1) {
            reinterpret_cast<float*>(out)[i] = d;
        } else {
            reinterpret_cast<int8_t*>(out)[i] = (int8_t)requantize_one<ASYM_OUT>(d, inv_out_scale, out_zp, lo, hi);
        }
    }
}

// Vector variant: N % 4 == 0, 16-byte aligned rows.  One thread per 4 consecutive n of a row: one int4 load, the
// row term once, the 4 column terms as one int4 load, float4 / packed-int8 store.  Same arithmetic as above;
// the asymmetric requantize tail rint(zp + t) runs through the quantizer's single-add rounding (|zp| < 2^20:
// exactly round-half-even of zp + t, see common.cuh) instead of float64.
// Lean variant for the common case (no int64 bias, |out_zp| < 2^20): all eight 16-byte loads of a lane (accumulators and
// column sums) are in flight before the first use, the zero-point terms are formed in 64-bit once per row / 4 columns,
// |d| <= 2^24 takes the one-multiply route (exact, see dequantize_one), the requantize tail is the reciprocal multiply
// and the single-add rounding.  Same bits as the general kernel for every int32 input.
template <int MODE>
__global__ void __launch_bounds__(256) acc_post_fast_kernel(const int32_t* __restrict__ acc, int64_t rows, int64_t M, int N4,
                                                           int64_t ldacc, float scale, AccZp z, float inv_out_scale,
                                                           float zpf, float lo, float hi, void* __restrict__ out) {
    const float tlo = lo - zpf, thi = hi - zpf, magic = kMagic + zpf;
    const int lane = threadIdx.x & 31;
    const int segs = (N4 + 127) >> 7;                                     // 128 int4 (512 columns) per warp unit
    const int64_t units = rows * segs, warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < units; u += warps) {
        const int64_t row = u / segs;
        const int c40 = (int)(u - row * segs) << 7;
        const int4* arow = reinterpret_cast<const int4*>(acc + row * ldacc);
        const int4* csrow = z.use_col ? reinterpret_cast<const int4*>(z.colsum_b + (row / M) * z.cs_stride) : nullptr;
        int4 a4[4], cs4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c4 = c40 + k * 32 + lane;
            a4[k] = c4 < N4 ? __ldcs(arow + c4) : make_int4(0, 0, 0, 0);
            cs4[k] = (csrow && c4 < N4) ? __ldg(csrow + c4) : make_int4(0, 0, 0, 0);
        }
        int64_t rowterm = -z.kterm;
        if (z.use_row) rowterm += (int64_t)__ldg(z.rowsum_a + row) * z.zp_b;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c4 = c40 + k * 32 + lane;
            if (c4 >= N4) continue;
            const int av[4] = {a4[k].x, a4[k].y, a4[k].z, a4[k].w}, cv[4] = {cs4[k].x, cs4[k].y, cs4[k].z, cs4[k].w};
            float d[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int64_t dv = (int64_t)av[e] - rowterm - (int64_t)cv[e] * z.zp_a;
                d[e] = (dv >= -16777216 && dv <= 16777216) ? __fmul_rn(__int2float_rn((int)dv), scale)
                                                           : (float)((double)dv * (double)scale);
            }
            const int64_t i = row * N4 + c4;
            if (MODE == 1) {
                __stcs(reinterpret_cast<float4*>(out) + i, make_float4(d[0], d[1], d[2], d[3]));
            } else {
                int c[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    c[e] = __float_as_int(__fadd_rn(fminf(fmaxf(__fmul_rn(inv_out_scale, d[e]), tlo), thi), magic));
                __stcs(reinterpret_cast<int*>(out) + i, pack4_codes(c[0], c[1], c[2], c[3]));
            }
        }
    }
}

template <int MODE, bool F64>   // MODE 1 dequant -> f32, 2 requant -> s8; F64: float64 requantize tail (huge zp)
__global__ void __launch_bounds__(256) acc_post_vec_kernel(const int32_t* __restrict__ acc, int64_t rows, int64_t M,
                                                          int N4, int64_t ldacc, float scale, AccZp z,
                                                          const int64_t* __restrict__ bias_q, float inv_out_scale,
                                                          int has_out_zp, double out_zp, float lo, float hi,
                                                          void* __restrict__ out) {
    // One warp per (row, 128-column segment) unit: the row term is formed once per unit, there is no per-element
    // index division, and 4 independent 16-byte loads per lane are in flight before the first use.
    const float zpf = (float)out_zp, tlo = lo - zpf, thi = hi - zpf, magic = kMagic + zpf;
    const int lane = threadIdx.x & 31;
    const int segs = (N4 + 127) >> 7;                                     // 128 int4 (512 columns) per unit
    const int64_t units = rows * segs, warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < units; u += warps) {
        const int64_t row = u / segs;
        const int c40 = (int)(u - row * segs) << 7;
        const int4* arow = reinterpret_cast<const int4*>(acc + row * ldacc);
        int4 a4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c4 = c40 + k * 32 + lane;
            a4[k] = c4 < N4 ? __ldcs(arow + c4) : make_int4(0, 0, 0, 0);
        }
        int64_t rowterm = -z.kterm;
        if (z.use_row) rowterm += (int64_t)__ldg(z.rowsum_a + row) * z.zp_b;
        const int4* csrow = z.use_col ? reinterpret_cast<const int4*>(z.colsum_b + (row / M) * z.cs_stride) : nullptr;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c4 = c40 + k * 32 + lane;
            if (c4 >= N4) continue;
            const int4 cs = csrow ? __ldg(csrow + c4) : make_int4(0, 0, 0, 0);
            const int av[4] = {a4[k].x, a4[k].y, a4[k].z, a4[k].w}, cv[4] = {cs.x, cs.y, cs.z, cs.w};
            float d[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int64_t a = av[e];
                if (MODE == 2 && bias_q) a += __ldg(bias_q + c4 * 4 + e);
                d[e] = dequantize_one(a - rowterm - (int64_t)cv[e] * z.zp_a, scale);
            }
            const int64_t i = row * N4 + c4;
            if (MODE == 1) {
                __stcs(reinterpret_cast<float4*>(out) + i, make_float4(d[0], d[1], d[2], d[3]));
            } else {
                int c[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (F64) c[e] = has_out_zp ? requantize_one<true>(d[e], inv_out_scale, out_zp, lo, hi)
                                               : requantize_one<false>(d[e], inv_out_scale, 0.0, lo, hi);
                    else c[e] = __float_as_int(__fadd_rn(fminf(fmaxf(__fmul_rn(inv_out_scale, d[e]), tlo), thi), magic));
                }
                __stcs(reinterpret_cast<int*>(out) + i, pack4_codes(c[0], c[1], c[2], c[3]));
            }
        }
    }
}

// requantize tail on an already dequantized tensor: clip(rint(zp + (1/s) * d))
template <bool ASYM>
__global__ void __launch_bounds__(256) requantize_f32_kernel(const float* __restrict__ d, int64_t n, float inv,
                                                            double zp, float lo, float hi, int8_t* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (int8_t)requantize_one<ASYM>(d[i], inv, zp, lo, hi);
}

// ------------------------------------------------------------------ row sums of a K-major s8 operand
__global__ void __launch_bounds__(256) rowsum_s8_kernel(const int8_t* __restrict__ q, int64_t rows, int64_t C,
                                                       int64_t ld, int32_t* __restrict__ rowsum) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < rows; row += warps) {
        const int8_t* src = q + row * ld;
        int sum = 0;
        if ((ld & 15) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0) {
            const int64_t c16 = C >> 4;
            for (int64_t c = lane; c < c16; c += 32) {
                int4 v = *reinterpret_cast<const int4*>(src + c * 16);
                sum = __dp4a(v.x, 0x01010101, sum);
                sum = __dp4a(v.y, 0x01010101, sum);
                sum = __dp4a(v.z, 0x01010101, sum);
                sum = __dp4a(v.w, 0x01010101, sum);
            }
            for (int64_t c = (c16 << 4) + lane; c < C; c += 32) sum += src[c];
        } else {
            for (int64_t c = lane; c < C; c += 32) sum += src[c];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) rowsum[row] = sum;
    }
}

// ------------------------------------------------------------------ K10 min / max
__device__ __forceinline__ void atomic_min_f32(float* addr, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

__global__ void minmax_init_kernel(float* mm, int64_t n_slots) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_slots) {
        mm[2 * i] = __int_as_float(0x7f800000);
        mm[2 * i + 1] = __int_as_float(0xff800000);
    }
}

__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ x, int64_t n, float* mm) {
    float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const int64_t n4 = n >> 2;
        const float4* x4 = reinterpret_cast<const float4*>(x);
        for (int64_t i = tid; i < n4; i += stride) {
            float4 v = __ldcs(x4 + i);
            lo = fminf(fminf(lo, v.x), fminf(v.y, fminf(v.z, v.w)));
            hi = fmaxf(fmaxf(hi, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += stride) {
            lo = fminf(lo, x[i]);
            hi = fmaxf(hi, x[i]);
        }
    } else {
        for (int64_t i = tid; i < n; i += stride) {
            lo = fminf(lo, x[i]);
            hi = fmaxf(hi, x[i]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ float slo[8], shi[8];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        slo[w] = lo;
        shi[w] = hi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
            lo = fminf(lo, slo[i]);
            hi = fmaxf(hi, shi[i]);
        }
        // -0.0 and +0.0 compare equal in NumPy's min/max; normalise so the bit tricks agree
        if (lo == 0.f) lo = 0.f;
        if (hi == 0.f) hi = 0.f;
        atomic_min_f32(mm, lo);
        atomic_max_f32(mm + 1, hi);
    }
}

// ------------------------------------------------------------------ K11 pack / unpack
// Eight codes <-> `bits` bytes, little-endian bitstream of two's-complement fields.
__global__ void __launch_bounds__(256) pack_kernel(const int8_t* __restrict__ q, int64_t n, int bits,
                                                  uint8_t* __restrict__ packed) {
    const int64_t groups = (n + 7) >> 3;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const unsigned mask = (1u << bits) - 1u;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
        unsigned long long word = 0;
        const int64_t base = g << 3;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            unsigned v = (base + j < n) ? ((unsigned)(int)q[base + j] & mask) : 0u;
            word |= (unsigned long long)v << (j * bits);
        }
        const int64_t total_bytes = (n * bits + 7) >> 3;
        for (int j = 0; j < bits; ++j) {
            const int64_t o = g * bits + j;
            if (o < total_bytes) packed[o] = (uint8_t)(word >> (8 * j));
        }
    }
}

__global__ void __launch_bounds__(256) unpack_kernel(const uint8_t* __restrict__ packed, int64_t n, int bits,
                                                    int8_t* __restrict__ q) {
    const int64_t groups = (n + 7) >> 3;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t total_bytes = (n * bits + 7) >> 3;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
        unsigned long long word = 0;
        for (int j = 0; j < bits; ++j) {
            const int64_t o = g * bits + j;
            if (o < total_bytes) word |= (unsigned long long)packed[o] << (8 * j);
        }
        const int64_t base = g << 3;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (base + j < n) {
                int v = (int)((word >> (j * bits)) & ((1u << bits) - 1u));
                v = (v << (32 - bits)) >> (32 - bits);         // sign-extend the field
                q[base + j] = (int8_t)v;
            }
        }
    }
}

// 32 codes per thread: two 16-byte loads -> BITS 32-bit words of the same bitstream (4 groups of 8 codes), all
// shifts compile-time constants.  The (< 32 code) tail runs through the byte-granular kernel above.
template <int BITS>
__global__ void __launch_bounds__(256) pack32_kernel(const int4* __restrict__ q, int64_t chunks, uint32_t* __restrict__ packed) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    constexpr unsigned mask = (1u << BITS) - 1u;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < chunks; t += stride) {
        const int4 a = __ldcs(q + 2 * t), b = __ldcs(q + 2 * t + 1);
        const unsigned w[8] = {(unsigned)a.x, (unsigned)a.y, (unsigned)a.z, (unsigned)a.w,
                               (unsigned)b.x, (unsigned)b.y, (unsigned)b.z, (unsigned)b.w};
        uint32_t out[BITS];
#pragma unroll
        for (int j = 0; j < BITS; ++j) out[j] = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const unsigned v = (w[i >> 2] >> ((i & 3) * 8)) & mask;
            const int bit = i * BITS, j = bit >> 5, sh = bit & 31;
            out[j] |= v << sh;
            if (sh + BITS > 32) out[j + 1] |= v >> (32 - sh);
        }
#pragma unroll
        for (int j = 0; j < BITS; ++j) __stcs(packed + t * BITS + j, out[j]);
    }
}

template <int BITS>
__global__ void __launch_bounds__(256) unpack32_kernel(const uint32_t* __restrict__ packed, int64_t chunks, int4* __restrict__ q) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    constexpr unsigned mask = (1u << BITS) - 1u;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < chunks; t += stride) {
        uint32_t in[BITS + 1];
#pragma unroll
        for (int j = 0; j < BITS; ++j) in[j] = __ldcs(packed + t * BITS + j);
        in[BITS] = 0;
        unsigned w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int bit = i * BITS, j = bit >> 5, sh = bit & 31;
            unsigned v = in[j] >> sh;
            if (sh + BITS > 32) v |= in[j + 1] << (32 - sh);
            int f = (int)(v & mask);
            f = (f << (32 - BITS)) >> (32 - BITS);                          // sign-extend the field
            w[i >> 2] |= ((unsigned)f & 0xffu) << ((i & 3) * 8);
        }
        __stcs(q + 2 * t, make_int4((int)w[0], (int)w[1], (int)w[2], (int)w[3]));
        __stcs(q + 2 * t + 1, make_int4((int)w[4], (int)w[5], (int)w[6], (int)w[7]));
    }
}

#define NQ_DISPATCH_BITS(bits, KERNEL, ...)                  \
    do {                                                     \
        switch (bits) {                                      \
            case 2: KERNEL<2> __VA_ARGS__; break;            \
            case 3: KERNEL<3> __VA_ARGS__; break;            \
            case 4: KERNEL<4> __VA_ARGS__; break;            \
            case 5: KERNEL<5> __VA_ARGS__; break;            \
            case 6: KERNEL<6> __VA_ARGS__; break;            \
            case 7: KERNEL<7> __VA_ARGS__; break;            \
            default: KERNEL<8> __VA_ARGS__; break;           \
        }                                                    \
    } while (0)

}  // namespace nq

using namespace nq;

extern "C" int nq_quantize_f32(const float* x, int64_t n, int bit_width, float scale, int has_zp, int64_t zp,
                               int8_t* out, void* stream) {
    NQ_REQUIRE(bit_width >= 2 && bit_width <= 8, "nq_quantize_f32: bit_width %d outside 2..8", bit_width);
    if (n <= 0) return NQ_OK;
    int qmode;
    const QArgs a = make_qargs(bit_width, scale, has_zp, zp, &qmode);
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 3) == 0);
    if (vec) {
        // small problems (a few passes per CTA at most): one CTA per 16 KB tile -- a grid-stride loop with 3.46 passes
        // per CTA leaves the last pass half empty (4096^2: 22.5 -> 20.5 us); large ones keep the persistent grid
        const int64_t tiles = std::max<int64_t>(1, (n / 4 + 1023) / 1024), cap = (int64_t)sm_count() * 8;
        const int grid = tiles <= 4 * cap ? (int)tiles : stream_grid((n + 15) / 16, 256);
        NQ_DISPATCH_QMODE(qmode, quantize_contig_kernel, <<<grid, 256, 0, s>>>(x, n, a, out));
    } else {
        const int grid = stream_grid(n, 256);
        NQ_DISPATCH_QMODE(qmode, quantize_scalar_kernel, <<<grid, 256, 0, s>>>(x, n, a, out));
    }
    NQ_CHECK_LAUNCH("nq_quantize_f32");
    return NQ_OK;
}

// Grid of the warp-per-unit accumulator kernels (unit = one row x 512 columns): one unit per warp while that is at most
// four waves of resident CTAs (a grid-stride loop with 3.46 passes per warp leaves the last pass half empty at 4096^2),
// the persistent grid for large problems.
static int acc_units_grid(int64_t rows, int64_t n4) {
    const int64_t units = rows * ((n4 + 127) / 128), ctas = (units + 7) / 8, cap = (int64_t)sm_count() * 8;
    return (int)std::max<int64_t>(1, ctas <= 4 * cap ? ctas : cap);
}

extern "C" int nq_transpose_s8(const int8_t* in, int64_t batch, int64_t R, int64_t Cc, int64_t ld_in, int64_t stride_in,
                               int8_t* out, int64_t ld_out, int64_t stride_out, void* stream) {
    NQ_REQUIRE(batch > 0 && R > 0 && Cc > 0, "nq_transpose_s8: empty problem");
    NQ_REQUIRE(ld_in >= Cc && ld_out >= R && ld_out % 4 == 0 && stride_out % 4 == 0 && ((uintptr_t)out & 3) == 0,
               "nq_transpose_s8: ld_in >= C, ld_out >= R, ld_out / stride_out multiples of 4 and a 4-byte aligned output are required");
    NQ_REQUIRE(R < (1ll << 30) && Cc < (1ll << 30), "nq_transpose_s8: extent too large");
    const int64_t r_tiles = (std::max<int64_t>(R, ld_out) + 127) / 128, c_tiles = (Cc + 63) / 64;
    NQ_REQUIRE(batch * r_tiles * c_tiles < (1ll << 31), "nq_transpose_s8: too many tiles");
    transpose_s8_kernel<<<(unsigned)(batch * r_tiles * c_tiles), 256, 0, (cudaStream_t)stream>>>(
        in, (int)R, (int)Cc, ld_in, stride_in, out, ld_out, stride_out, (int)r_tiles, (int)c_tiles);
    NQ_CHECK_LAUNCH("nq_transpose_s8");
    return NQ_OK;
}

extern "C" int nq_quantize_patches_f32(const float* x, int64_t B, int64_t C, int64_t H, int64_t W, int64_t KH, int64_t KW,
                                       int bit_width, float scale, int has_zp, int64_t zp, int8_t* out, int64_t ldo,
                                       void* stream) {
    NQ_REQUIRE(bit_width >= 2 && bit_width <= 8, "nq_quantize_patches_f32: bit_width %d outside 2..8", bit_width);
    NQ_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && KH > 0 && KW > 0, "nq_quantize_patches_f32: empty geometry");
    NQ_REQUIRE(H % KH == 0 && W % KW == 0, "nq_quantize_patches_f32: the patches must tile the image (H %% KH == 0, W %% KW == 0)");
    NQ_REQUIRE(KW % 4 == 0 && ((uintptr_t)x & 15) == 0, "nq_quantize_patches_f32: KW %% 4 == 0 and a 16-byte aligned image are required");
    NQ_REQUIRE(ldo >= C * KH * KW && ldo % 4 == 0 && ((uintptr_t)out & 3) == 0, "nq_quantize_patches_f32: ldo < C*KH*KW or unaligned output");
    NQ_REQUIRE(C * KH < (1ll << 31) && H < (1ll << 31) && W < (1ll << 31), "nq_quantize_patches_f32: extent too large");
    int qmode;
    const QArgs a = make_qargs(bit_width, scale, has_zp, zp, &qmode);
    const int64_t n_rows = B * (H / KH), vec_per_patch = C * KH * (KW / 4);
    const int threads = (int)(vec_per_patch >= 1024 ? 1024 : ((vec_per_patch + 31) / 32) * 32);
    const int64_t cap = (int64_t)sm_count() * (2048 / threads) * 4;
    const int grid = (int)(n_rows < cap ? n_rows : cap);
    cudaStream_t s = (cudaStream_t)stream;
    NQ_DISPATCH_QMODE(qmode, quantize_patches_kernel, <<<grid, threads, 0, s>>>(x, n_rows, (int)C, (int)H, (int)W, (int)KH, (int)KW, a, out, ldo));
    NQ_CHECK_LAUNCH("nq_quantize_patches_f32");
    return NQ_OK;
}

extern "C" int nq_quantize_f32_i64(const float* x, int64_t n, int bit_width, float scale, int has_zp, int64_t zp,
                                   int64_t* out, void* stream) {
    NQ_REQUIRE(bit_width >= 2 && bit_width <= 32, "nq_quantize_f32_i64: bit_width %d outside 2..32", bit_width);
    if (n <= 0) return NQ_OK;
    // np.clip converts the Python-float bounds to the array dtype: float32 when symmetric
    // (2^31-1 rounds to 2^31), float64 when the zero-point add promoted the data.
    // (float64 when the int64 zero-point add promoted the data: the asymmetric branch of the kernel)
    NQ_REQUIRE(!has_zp || (zp > -(1ll << 52) && zp < (1ll << 52)), "nq_quantize_f32_i64: |zero_point| must be < 2^52");
    const double dlo = -ldexp(1.0, bit_width - 1), dhi = ldexp(1.0, bit_width - 1) - 1.0;
    quantize_i64_kernel<<<stream_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, scale, has_zp ? 1 : 0, has_zp ? (double)zp : 0.0,
                                                                              (float)dlo, (float)dhi, dlo, dhi, out);
    NQ_CHECK_LAUNCH("nq_quantize_f32_i64");
    return NQ_OK;
}

extern "C" int nq_quantize_f32_4d(const float* x, int64_t d0, int64_t d1, int64_t R, int64_t C, int64_t s0,
                                  int64_t s1, int64_t sr, int64_t sc, int bit_width, float scale, int has_zp,
                                  int64_t zp, int8_t* out, int64_t ldo, int32_t* rowsum, void* stream) {
    NQ_REQUIRE(bit_width >= 2 && bit_width <= 8, "nq_quantize_f32_4d: bit_width %d outside 2..8", bit_width);
    NQ_REQUIRE(ldo >= C, "nq_quantize_f32_4d: ldo %lld < C %lld", (long long)ldo, (long long)C);
    const int64_t batch = d0 * d1, rows = batch * R;
    if (rows <= 0 || C <= 0) return NQ_OK;
    int qmode;
    const QArgs a = make_qargs(bit_width, scale, has_zp, zp, &qmode);
    cudaStream_t s = (cudaStream_t)stream;
    if (sr == 1 && sc != 1 && batch <= 65535) {
        if (rowsum) {
            cudaError_t e = cudaMemsetAsync(rowsum, 0, sizeof(int32_t) * rows, s);
            if (e != cudaSuccess) return cuda_fail(e, "nq_quantize_f32_4d memset");
        }
        dim3 grid((unsigned)((R + 31) / 32), (unsigned)((ldo + 31) / 32), (unsigned)batch);
        NQ_REQUIRE(grid.y <= 65535, "nq_quantize_f32_4d: C too large for the transposing path");
        NQ_DISPATCH_QMODE(qmode, quantize_transpose_kernel, <<<grid, 256, 0, s>>>(x, d1, R, C, s0, s1, sc, a, out, ldo, rowsum));
    } else {
        NQ_REQUIRE(rows < (1ll << 31) && R < (1ll << 31) && d1 < (1ll << 31), "nq_quantize_f32_4d: too many rows");
        const int64_t c4 = C >> 2;
        const int lpr = (sc == 1 && c4 <= 8) ? 8 : (sc == 1 && c4 <= 16) ? 16 : 32;
        const int grid = stream_grid(rows * lpr, 256);
        NQ_DISPATCH_QMODE(qmode, quantize_rows_kernel, <<<grid, 256, 0, s>>>(x, (uint32_t)d1, (uint32_t)R, C, s0, s1, sr, sc, (uint32_t)rows, a, out, ldo, rowsum, lpr));
    }
    NQ_CHECK_LAUNCH("nq_quantize_f32_4d");
    return NQ_OK;
}

extern "C" int nq_dequantize(const void* q, int elem_bytes, int64_t n, float scale, int has_zp, int64_t zp,
                             float* out, void* stream) {
    if (n <= 0) return NQ_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t z = has_zp ? zp : 0;
    if (elem_bytes == 1) {
        const bool vec = ((reinterpret_cast<uintptr_t>(q) & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                         z > -(1 << 23) && z < (1 << 23);
        const int64_t n4 = vec ? (n >> 2) : 0;
        if (n4) dequantize_s8x4_kernel<<<stream_grid((n4 + 3) / 4, 256), 256, 0, s>>>((const int8_t*)q, n4, scale, (int)z, out);
        const int64_t done = n4 << 2;
        if (done < n)
            dequantize_kernel<int8_t><<<stream_grid(n - done, 256), 256, 0, s>>>((const int8_t*)q + done, n - done, scale, z, out + done);
    } else if (elem_bytes == 4) {
        dequantize_kernel<int32_t><<<stream_grid(n, 256), 256, 0, s>>>((const int32_t*)q, n, scale, z, out);
    } else if (elem_bytes == 8) {
        dequantize_kernel<int64_t><<<stream_grid(n, 256), 256, 0, s>>>((const int64_t*)q, n, scale, z, out);
    } else {
        NQ_REQUIRE(false, "nq_dequantize: elem_bytes %d not in {1,4,8}", elem_bytes);
    }
    NQ_CHECK_LAUNCH("nq_dequantize");
    return NQ_OK;
}

extern "C" int nq_dequantize_acc(const int32_t* acc, int64_t batch, int64_t M, int64_t N, int64_t ldacc,
                                 float scale, const nq_acc_zp* zp, float* out, void* stream) {
    if (batch * M * N <= 0) return NQ_OK;
    if (int rc = check_acc_zp(zp)) return rc;
    const AccZp z = make_acc_zp(zp);
    const bool vec = N % 4 == 0 && ldacc % 4 == 0 && ((uintptr_t)acc % 16 == 0) && ((uintptr_t)out % 16 == 0) && N / 4 < (1ll << 31) &&
                     (!z.use_col || (((uintptr_t)z.colsum_b % 16 == 0) && z.cs_stride % 4 == 0));
    if (vec)
        acc_post_fast_kernel<1><<<acc_units_grid(batch * M, N / 4), 256, 0, (cudaStream_t)stream>>>(
            acc, batch * M, M, (int)(N / 4), ldacc, scale, z, 0.f, 0.f, 0.f, 0.f, out);

    else
        acc_post_kernel<1, false><<<stream_grid(batch * M * N, 256), 256, 0, (cudaStream_t)stream>>>(
            acc, batch, M, N, ldacc, scale, z, nullptr, 0.f, 0.0, 0.f, 0.f, out);
    NQ_CHECK_LAUNCH("nq_dequantize_acc");
    return NQ_OK;
}

extern "C" int nq_requantize_acc(const int32_t* acc, int64_t batch, int64_t M, int64_t N, int64_t ldacc,
                                 float scale, const nq_acc_zp* zp, const int64_t* bias_q, int out_bits,
                                 float out_scale, int has_out_zp, int64_t out_zp, int8_t* out, void* stream) {
    NQ_REQUIRE(out_bits >= 2 && out_bits <= 8, "nq_requantize_acc: out_bits %d outside 2..8", out_bits);
    if (batch * M * N <= 0) return NQ_OK;
    if (int rc = check_acc_zp(zp)) return rc;
    float lo, hi;
    qrange(out_bits, &lo, &hi);
    const float inv = 1.0f / out_scale;            // float32(1) / float32 scale, IEEE division (host)
    const int grid = stream_grid(batch * M * N, 256);
    cudaStream_t s = (cudaStream_t)stream;
    {
        const AccZp z = make_acc_zp(zp);
        const bool vec = N % 4 == 0 && ldacc % 4 == 0 && ((uintptr_t)acc % 16 == 0) && ((uintptr_t)out % 4 == 0) && N / 4 < (1ll << 31) &&
                         (!z.use_col || (((uintptr_t)z.colsum_b % 16 == 0) && z.cs_stride % 4 == 0));
        if (vec) {
            const int g4 = acc_units_grid(batch * M, N / 4);
            const bool f64 = has_out_zp && !(out_zp > -(1 << 20) && out_zp < (1 << 20));
            if (!f64 && !bias_q) {
                acc_post_fast_kernel<2><<<g4, 256, 0, s>>>(acc, batch * M, M, (int)(N / 4), ldacc, scale, z, inv,
                                                           has_out_zp ? (float)out_zp : 0.f, lo, hi, out);
                NQ_CHECK_LAUNCH("nq_requantize_acc");
                return NQ_OK;
            }
            if (f64)
                acc_post_vec_kernel<2, true><<<g4, 256, 0, s>>>(acc, batch * M, M, (int)(N / 4), ldacc, scale, z, bias_q, inv,
                                                                has_out_zp, (double)out_zp, lo, hi, out);
            else
                acc_post_vec_kernel<2, false><<<g4, 256, 0, s>>>(acc, batch * M, M, (int)(N / 4), ldacc, scale, z, bias_q, inv,
                                                                 has_out_zp, has_out_zp ? (double)out_zp : 0.0, lo, hi, out);
            NQ_CHECK_LAUNCH("nq_requantize_acc");
            return NQ_OK;
        }
    }
    if (has_out_zp)
        acc_post_kernel<2, true><<<grid, 256, 0, s>>>(acc, batch, M, N, ldacc, scale, make_acc_zp(zp), bias_q, inv, (double)out_zp, lo, hi, out);
    else
        acc_post_kernel<2, false><<<grid, 256, 0, s>>>(acc, batch, M, N, ldacc, scale, make_acc_zp(zp), bias_q, inv, 0.0, lo, hi, out);
    NQ_CHECK_LAUNCH("nq_requantize_acc");
    return NQ_OK;
}

extern "C" int nq_requantize_f32(const float* d, int64_t n, int out_bits, float out_scale, int has_out_zp,
                                 int64_t out_zp, int8_t* out, void* stream) {
    NQ_REQUIRE(out_bits >= 2 && out_bits <= 8, "nq_requantize_f32: out_bits %d outside 2..8", out_bits);
    if (n <= 0) return NQ_OK;
    float lo, hi;
    qrange(out_bits, &lo, &hi);
    const float inv = 1.0f / out_scale;
    const int grid = stream_grid(n, 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (has_out_zp) requantize_f32_kernel<true><<<grid, 256, 0, s>>>(d, n, inv, (double)out_zp, lo, hi, out);
    else requantize_f32_kernel<false><<<grid, 256, 0, s>>>(d, n, inv, 0.0, lo, hi, out);
    NQ_CHECK_LAUNCH("nq_requantize_f32");
    return NQ_OK;
}

extern "C" int nq_rowsum_s8(const int8_t* q, int64_t rows, int64_t C, int64_t ld, int32_t* rowsum, void* stream) {
    if (rows <= 0) return NQ_OK;
    rowsum_s8_kernel<<<stream_grid(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(q, rows, C, ld, rowsum);
    NQ_CHECK_LAUNCH("nq_rowsum_s8");
    return NQ_OK;
}

extern "C" int nq_minmax_init(float* minmax, int64_t n_slots, void* stream) {
    if (n_slots <= 0) return NQ_OK;
    minmax_init_kernel<<<(unsigned)((n_slots + 255) / 256), 256, 0, (cudaStream_t)stream>>>(minmax, n_slots);
    NQ_CHECK_LAUNCH("nq_minmax_init");
    return NQ_OK;
}

extern "C" int nq_minmax_f32(const float* x, int64_t n, float* minmax, int64_t slot, void* stream) {
    if (n <= 0) return NQ_OK;
    minmax_kernel<<<stream_grid((n + 3) / 4, 256, 4), 256, 0, (cudaStream_t)stream>>>(x, n, minmax + 2 * slot);
    NQ_CHECK_LAUNCH("nq_minmax_f32");
    return NQ_OK;
}

extern "C" int nq_pack_s8(const int8_t* q, int64_t n, int bit_width, uint8_t* packed, void* stream) {
    NQ_REQUIRE(bit_width >= 2 && bit_width <= 8, "nq_pack_s8: bit_width %d outside 2..8", bit_width);
    if (n <= 0) return NQ_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int64_t done = 0;
    if (((uintptr_t)q % 16 == 0) && ((uintptr_t)packed % 4 == 0) && n >= 32) {
        const int64_t chunks = n / 32;
        NQ_DISPATCH_BITS(bit_width, pack32_kernel, <<<stream_grid(chunks, 256), 256, 0, s>>>((const int4*)q, chunks, (uint32_t*)packed));
        done = chunks * 32;
    }
    if (done < n)                                      // tail (or unaligned buffers): byte-granular kernel
        pack_kernel<<<stream_grid((n - done + 7) / 8, 256), 256, 0, s>>>(q + done, n - done, bit_width,
                                                                        packed + done / 8 * bit_width);
    NQ_CHECK_LAUNCH("nq_pack_s8");
    return NQ_OK;
}

extern "C" int nq_unpack_s8(const uint8_t* packed, int64_t n, int bit_width, int8_t* q, void* stream) {
    NQ_REQUIRE(bit_width >= 2 && bit_width <= 8, "nq_unpack_s8: bit_width %d outside 2..8", bit_width);
    if (n <= 0) return NQ_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int64_t done = 0;
    if (((uintptr_t)q % 16 == 0) && ((uintptr_t)packed % 4 == 0) && n >= 32) {
        const int64_t chunks = n / 32;
        NQ_DISPATCH_BITS(bit_width, unpack32_kernel, <<<stream_grid(chunks, 256), 256, 0, s>>>((const uint32_t*)packed, chunks, (int4*)q));
        done = chunks * 32;
    }
    if (done < n)
        unpack_kernel<<<stream_grid((n - done + 7) / 8, 256), 256, 0, s>>>(packed + done / 8 * bit_width, n - done, bit_width,
                                                                          q + done);
    NQ_CHECK_LAUNCH("nq_unpack_s8");
    return NQ_OK;
}
