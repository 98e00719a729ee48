// K4/K5: int8 x int8 -> int32 GEMM on the 5th-gen tensor cores (tcgen05.mma kind::i8),
// replacing the reference's `np.matmul` on int64 arrays (numpy_quantization.py:44-61).
//
// Persistent, warp-specialised, one CTA per SM (or a CTA pair per 256-row tile, cta_group::2, when the main loop
// dominates: K >= 1024):
//   warp 0      TMA producer   (cp.async.bulk.tensor 3-D, 128B swizzle, K-major tiles)
//   warp 1      MMA issuer     (one elected thread, tcgen05.mma, M=128 (256 for a pair), N=BN, K=32)
//   warp 2      TMEM allocator (512 columns: 2 x 256, 4 x 128 or 8 x 64 accumulator buffers)
//   warps 4-19  epilogue, one template instantiation per mode:
//                 RAW / DEQUANT (ragged, general) / REQUANT   tcgen05.ld x16 -> swizzled smem transpose -> 16-byte stores
//                 DEQUANT wide (+bias, +residual)              x32, 4 KB slabs: whole 128-byte lines per row segment
//                 QUANT rows / cols, GELU_QUANT                thread = row, codes straight into the next operand
//                 SOFTMAX_QUANT                                4 warps per lane quarter share a row through smem
//               narrow tiles (BN < 256) are drained by 2 / 4 independent warp groups, one tile each
// smem ring of STAGES x (A 128x128 B + B BNx128 B); mbarrier full/empty per stage and tmem_full/tmem_empty per
// accumulator buffer, so the epilogue of tile i overlaps the main loop of tile i+1; setmaxnreg 56 / 104.
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace nq {

constexpr int NUM_EPI_WARPS = 16;
constexpr int NUM_THREADS = 128 + 32 * NUM_EPI_WARPS;  // 640 threads; registers re-balanced with setmaxnreg (56 / 104)

template <int BN, bool WIDE = false, bool TWO = false> struct Cfg {
    // WIDE (NQ_EPI_DEQUANT on aligned outputs): 32-column staging slabs, so every float32 row segment that a warp
    // loads (residual) or stores is a full 128-byte line; paid for with pipeline stages.
    // TWO: CTA pair (cta_group::2): the pair computes a 256 x BN tile, each CTA stages its 128 rows of A and its half
    // (BN / 2 rows) of B -- per-SM operand traffic per MMA halves, which is what the 1-CTA main loop is bound by.
    static constexpr int A_BYTES = BM * BK;
    static constexpr int B_BYTES = (TWO ? BN / 2 : BN) * BK;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // Narrow tiles are drained by GROUPS independent sets of epilogue warps, each on its own tile, so that several
    // small tiles are in flight per CTA (their per-tile latency chain, not the math, bounds e.g. the P.V GEMM):
    // BN = 256: 1 group of 16 warps, 2 accumulator buffers; 128: 2 groups of 8, 4 buffers; 64: 4 groups of 4, 8.
    static constexpr int GROUPS = 256 / BN;
    static constexpr int NACC = 2 * GROUPS;
    static constexpr int TMEM_COLS = 512;                                  // NACC * BN
    static constexpr int SLAB_WORDS = WIDE ? 32 * 32 : 32 * 16;            // per-warp staging slab (32 rows)
    static constexpr int EPI_BYTES = NUM_EPI_WARPS * (SLAB_WORDS * 4 + 32 * 4);   // staging slabs + row terms
    static constexpr int BAR_BYTES = 256;
    static constexpr int FIT = (232448 - EPI_BYTES - BAR_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = FIT > 6 ? 6 : FIT;                       // 3-6 (full / empty barriers: 2 x 6 slots)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
    static_assert(STAGES >= 3 && SMEM_BYTES <= 232448, "exceeds the 227 KB opt-in shared memory of sm_100");
    static_assert(!TWO || BN == 256, "CTA pairs are built for the 256-column tile only");
};

struct GemmParams {
    int64_t M, N, K, batch;
    int64_t ldc, stride_c;
    int64_t c_inner, stride_c_inner; // C batch offset = (b / c_inner) * stride_c + (b % c_inner) * stride_c_inner
    int a_batched, b_batched;        // 0 -> operand shared across the batch (TMA batch coord 0)
    int mode;
    int fast32;                      // zero-point arithmetic provably fits int32 (host-checked bound)
    int fast24;                      // ... and |acc - zero-point terms| < 2^24: int -> float conversion is exact
    float scale;
    AccZp zp;
    const float* bias_f32;
    const float* residual;           // DEQUANT: out = (bias + dequant) + residual[b, m, n]
    int64_t ldr, stride_r;
    const int64_t* bias_q;
    float inv_out_scale;
    double out_zp;
    int asym_out;
    float lo, hi;
    void* C;
    // QUANT_OUT: float result quantized and scattered as the next GEMM's int8 operand
    QArgs qargs;
    uint32_t q_S, q_D;               // rows per image (m = mb*S + ms), columns per head (n = nh*D + nd)
    int64_t q_off[6];                // byte offset  = bo*[0] + bi*[1] + mb*[2] + ms*[3] + nh*[4] + nd*[5]
    int64_t q_rs[6];                 // row-sum slot = same decomposition
    int32_t* q_rowsum;
    // SOFTMAX epilogue: softmax(dequant / sm_div) over the (single) N tile, quantized with qargs
    int q_rs_exclusive;              // Q8 ROWS: every (row, head) row-sum slot is written by exactly one warp: plain store
    float g_prdiv, g_nl2e, g_add, g_out;   // GELU_QUANT: 0.3275911 / c1, -log2(e) / c1^2, c2, c3 / s_out
    int two_cta;                     // CTA-pair kernel (256-row tiles, cta_group::2)
    int deq_wide;                    // DEQUANT: alignment / extent conditions of the 32-column epilogue hold (host-checked)
    int dbg;                         // NQ_GEMM_DBG bit mask (measurement only, benchmarks/probe_epilogue_parts.py): 1 no epilogue
                                     // math, 2 no TMEM loads, 4 no stores, 8 no MMAs, 16 no epilogue chunks, 32 no TMA loads -- garbage results
    long long* trace;                // NQ_GEMM_TRACE (measurement only): clock64 stamps of CTA 0's first 64 tiles, 8 per tile
    int reverse;                     // tile i of the schedule is output tile total - 1 - i (L2 reuse of the producer's tail)
    int req_rows;                    // REQUANT: 32-bit windows and alignment of the thread-per-row epilogue hold (host-checked)
    int sm_noclamp;                  // SOFTMAX: out_zp >= lo: p / s_out + zp (p in [0, 1]) needs no lower clamp
    float sm_top;                    // SOFTMAX: upper clamp in the magic-sum domain (1.5 * 2^23 + hi), huge when p = 1 fits
    int fast22;                      // SOFTMAX: |acc - zero-point terms| < 2^22 proved on the host (magic int->float route)
    // implicit-GEMM convolution (nq_qconv2d_s8): A rows are output pixels (n, oh, ow) of a padded NHWC image, read
    // by im2col-mode TMA; K = (kh, kw, c) in 64-channel slices
    int conv;
    uint32_t cv_C, cv_KW, cv_OW, cv_OH, cv_sw, cv_sh;
};

// Measurement hooks (benchmarks/probe_epilogue_parts.py, probe_tile_trace.py) exist only in builds with -DNQ_GEMM_DEBUG
// (NQ_EXTRA_NVCC_FLAGS=-DNQ_GEMM_DEBUG python -m numpy_quant_b200.build): the shipped kernel carries neither the
// NQ_GEMM_DBG tests nor the clock stamps in its per-chunk loops.
#ifdef NQ_GEMM_DEBUG
#define NQ_DBG(bit_) ((p.dbg & (bit_)) != 0)
#define NQ_TRACE_KB(li_, kb_, ev_)                                                                                   \
    do {                                                                                                             \
        if (p.trace && blockIdx.x == 0 && (li_) == 5u && (kb_) < 16u) p.trace[512 + (kb_) * 4 + (ev_)] = clock64();     \
    } while (0)
#define NQ_TRACE(li_, ev_)                                                                     \
    do {                                                                                       \
        if (p.trace && blockIdx.x == 0 && (li_) < 64u) p.trace[(li_) * 8 + (ev_)] = clock64(); \
    } while (0)
#else
#define NQ_DBG(bit_) false
#define NQ_TRACE_KB(li_, kb_, ev_) do { } while (0)
#define NQ_TRACE(li_, ev_) do { } while (0)
#endif

// ------------------------------------------------------------------ epilogue math
// dequantize with 32-bit zero-point arithmetic: identical bits to f32(f64(d) * f64(scale)).
// |d| < 2^22 converts through the 1.5*2^23 magic constant (integer add + FADD, full rate)
// instead of I2F; larger values take the conversion / float64 routes.
__device__ __noinline__ float deq_slow(int d, float scale) {       // rare: |d| >= 2^22, kept out of line
    if (d >= -16777216 && d <= 16777216) return __fmul_rn((float)d, scale);
    return (float)((double)d * (double)scale);
}

__device__ __forceinline__ int64_t tile_zp(const AccZp& z, int64_t rowterm, int64_t b, int64_t n) {
    int64_t v = rowterm;
    if (z.use_col) v += (int64_t)__ldg(z.colsum_b + b * z.cs_stride + n) * z.zp_a;
    return v;
}

// GELU chain Div(x, c1) -> Erf -> Add(c2) -> Mul(x, .) -> Mul(., c3) for the fused epilogue: float glue under the
// 1e-5 contract.  Same Abramowitz & Stegun 7.1.26 erf as numpy_helper.py:95-112 with the
// polynomial as fused multiply-adds, 1/(1 + p|u|) and exp(-u^2) on the MUFU: 16 instructions + 2 MUFU per element
// (the standalone kernels keep one IEEE rounding per ONNX node).
__device__ __forceinline__ float gelu_fast(float x, float p_rdiv, float nl2e_rdiv2, float c_add, float c_out) {
    // u = x / c1 enters only as |u| (p_rdiv = 0.3275911 / c1), u^2 (nl2e_rdiv2 = -log2(e) / c1^2) and its sign
    const float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f, a5 = 1.061405429f;
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(__fmaf_rn(p_rdiv, fabsf(x), 1.0f)));
    float y = __fmaf_rn(a5, t, a4);
    y = __fmaf_rn(y, t, a3);
    y = __fmaf_rn(y, t, a2);
    y = __fmaf_rn(y, t, a1);
    y = __fmul_rn(y, t);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(__fmul_rn(x, x), nl2e_rdiv2)));
    y = __fmaf_rn(-y, e, 1.0f);
    const float erf_u = __int_as_float(__float_as_int(y) ^ (__float_as_int(x) & 0x80000000));   // sign(u) * y, c1 > 0
    return __fmul_rn(__fmul_rn(x, __fadd_rn(erf_u, c_add)), c_out);      // c_out = c3 / s_out: the quotient to round
}

// Two elements per instruction: Blackwell's packed FFMA2 (fma.rn.f32x2) halves the issue slots of the float chain --
// every epilogue here is bound by instruction issue.  Same arithmetic as gelu_fast(); the packed multiply / add
// intrinsics may contract into FFMA2 (float glue, 1e-5 contract).  np1..np5 are the NEGATED A&S coefficients.
__device__ __forceinline__ float2 gelu_fast2(float2 x, float p_rdiv, float nl2e_rdiv2, float c_add, float c_out) {
    const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
    const float2 den = __ffma2_rn(make_float2(p_rdiv, p_rdiv), ax, make_float2(1.0f, 1.0f));
    float2 t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
    float2 y = __ffma2_rn(make_float2(-1.061405429f, -1.061405429f), t, make_float2(1.453152027f, 1.453152027f));
    y = __ffma2_rn(y, t, make_float2(-1.421413741f, -1.421413741f));
    y = __ffma2_rn(y, t, make_float2(0.284496736f, 0.284496736f));
    y = __ffma2_rn(y, t, make_float2(-0.254829592f, -0.254829592f));
    y = __fmul2_rn(y, t);                                                 // -poly(t)
    const float2 z = __fmul2_rn(__fmul2_rn(x, x), make_float2(nl2e_rdiv2, nl2e_rdiv2));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(z.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(z.y));
    y = __ffma2_rn(y, e, make_float2(1.0f, 1.0f));                        // 1 - poly(t) * exp(-u^2)
    // (sign(x) erf|u| + c2) * x * c = erf|u| * |x c| + c2 * (x c) for c = c3 / s_out > 0 (host-checked): the sign never
    // has to be copied onto the erf value, and |.| is an operand modifier.  Same FMA inputs up to two cancelling signs.
    const float2 xc = __fmul2_rn(x, make_float2(c_out, c_out));
    const float2 axc = make_float2(fabsf(xc.x), fabsf(xc.y));
    (void)c_add;                                                          // == 1 (host-checked): c2 * (x c) is x c itself
    return __ffma2_rn(y, axc, xc);
}

// Exact re-evaluation of one lane's share of a 32 x 32 step of the wide dequant epilogue (rare: some value left the
// 2^22 window of the magic-constant int->float conversion).  Same slab / row-term layout as the hot loop.
__device__ __noinline__ void deq_wide_fix(const uint32_t* slab, const uint32_t* stg_row, int lane, int ct0, int ct1, int ct2,
                                          int ct3, float b0, float b1, float b2, float b3, bool has_bias, const float* rbase,
                                          int64_t ldr, float* crow, int64_t ldc, int rows_left, float scale) {
    const int cj = lane & 7, rq = lane >> 3;
    const int ct[4] = {ct0, ct1, ct2, ct3};
    const float bs[4] = {b0, b1, b2, b3};
    for (int it = 0; it < 8; ++it) {
        const int rr = it * 4 + rq;
        if (rr >= rows_left) continue;
        const uint32_t* src = slab + rr * 32 + ((cj ^ (rr & 7)) << 2);
        const int rm = (int)stg_row[rr];
        float f[4];
        for (int k = 0; k < 4; ++k) {
            const int d = (int)src[k] + rm - ct[k] - 0x4B400000;          // acc - zero-point terms (int32: host bound)
            f[k] = (d >= -16777216 && d <= 16777216) ? __fmul_rn((float)d, scale) : (float)((double)d * (double)scale);
            if (has_bias) f[k] = __fadd_rn(bs[k], f[k]);
            if (rbase) f[k] = __fadd_rn(f[k], rbase[(int64_t)rr * ldr + k]);
        }
        *reinterpret_cast<float4*>(crow + (int64_t)it * 4 * ldc) = make_float4(f[0], f[1], f[2], f[3]);
    }
}

constexpr int EM_RAW = 0, EM_DEQ_FAST = 1, EM_DEQ_GENERAL = 2, EM_REQUANT = 3, EM_Q8_ROWS = 4, EM_Q8_COLS = 5,
              EM_SOFTMAX_SYM = 6, EM_SOFTMAX_ASYM = 7, EM_Q8_GELU = 8, EM_DEQ_WIDE = 9, EM_DEQ_WIDE_RES = 10, EM_REQ_ROWS = 11;

// requantize (numpy_quantization.py:64-72) of one element on the general 64-bit route: int64 bias add, int64 zero-point
// terms, dequantize, reciprocal multiply, clip(rint(zp + t)).  Out of line: the fast REQUANT epilogue calls it only for
// tiles whose bias / accumulators leave the 32-bit windows.
__device__ __noinline__ int requant_slow(int acc, int64_t rowterm, const AccZp z, int64_t b, int64_t n, const int64_t* bias_q,
                                         float scale, float inv_out_scale, int asym_out, double out_zp, float lo, float hi) {
    int64_t a = (int64_t)acc;
    if (bias_q) a += __ldg(bias_q + n);
    const float d = dequantize_one(a - tile_zp(z, rowterm, b, n), scale);
    return asym_out ? requantize_one<true>(d, inv_out_scale, out_zp, lo, hi) : requantize_one<false>(d, inv_out_scale, 0.0, lo, hi);
}

// Division of n < 2^31 by a divisor fixed for the launch (Granlund / Montgomery round-up method): one multiply-high,
// one add, one shift instead of the ~25-instruction emulated 32-bit divide -- the tile bookkeeping of every role runs
// once per tile per warp and used to be a quarter of the epilogue warps' instruction stream.
struct UDiv {
    uint32_t d, m, l;
};
__device__ __forceinline__ UDiv make_udiv(uint32_t d) {
    UDiv f;
    f.d = d;
    f.l = d > 1u ? 32u - (uint32_t)__clz((int)(d - 1u)) : 0u;             // ceil(log2 d)
    f.m = (uint32_t)(((((uint64_t)1 << f.l) - d) << 32) / d) + 1u;
    return f;
}
__device__ __forceinline__ uint32_t fd_div(uint32_t n, const UDiv& f) { return (__umulhi(f.m, n) + n) >> f.l; }

template <int BN, int EMODE, bool TWO = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
qgemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
    using C = Cfg<BN, (EMODE == EM_DEQ_WIDE || EMODE == EM_DEQ_WIDE_RES), TWO>;
    constexpr int BMT = TWO ? 2 * BM : BM;                                // rows of the (pair's) tile
    const uint32_t cta_rank = TWO ? cluster_ctarank() : 0u;               // 0 = leader (issues the MMAs)
    const uint32_t tile_first = TWO ? (blockIdx.x >> 1) : blockIdx.x, tile_step = TWO ? (gridDim.x >> 1) : gridDim.x;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 128B swizzle atoms need 1024-byte aligned tiles
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();      // swizzle atoms need the 1024-byte alignment requested above
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + C::STAGES * C::A_BYTES;
    uint32_t* epi = reinterpret_cast<uint32_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::EPI_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + C::STAGES;
    uint64_t* tfull_bar = bars + 2 * C::STAGES;
    uint64_t* tempty_bar = bars + 2 * C::STAGES + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 16);
    volatile uint32_t* pace = tmem_slot + 1;                              // tiles issued by the MMA warp (residual prefetcher)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile bookkeeping in 32 bits (host guarantees total_tiles < 2^31): 64-bit divides are long
    // emulated sequences that would otherwise sit in every role's tile loop
    const uint32_t m_tiles = (uint32_t)((p.M + BMT - 1) / BMT), n_tiles = (uint32_t)((p.N + BN - 1) / BN);
    const uint32_t tiles_per_batch = m_tiles * n_tiles, total_tiles = tiles_per_batch * (uint32_t)p.batch;
    const uint32_t k_blocks = (uint32_t)((p.K + BK - 1) / BK);
    const UDiv fd_tpb = make_udiv(tiles_per_batch), fd_nt = make_udiv(n_tiles);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < C::STAGES; ++i) {
            mbar_init(smem_u32(full_bar + i), 1);
            mbar_init(smem_u32(empty_bar + i), 1);
        }
        for (int i = 0; i < C::NACC; ++i) {
            mbar_init(smem_u32(tfull_bar + i), 1);
            // one arrival per epilogue warp of the group (pair: the warps of both CTAs arrive on the leader's barrier)
            mbar_init(smem_u32(tempty_bar + i), (TWO ? 2 : 1) * NUM_EPI_WARPS / C::GROUPS);
        }
        *pace = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (TWO) {                                              // same warp, same slot address in both CTAs
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "n"(C::TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "n"(C::TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if constexpr (TWO) cluster_sync_all();                                // peer barriers initialised, both TMEM halves allocated
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Register re-balancing between the warpgroups.  The pool is what the CTA got at launch (640 threads x 96 =
    // 61440 registers): the producer / MMA / allocator warpgroup shrinks to 56 per thread, the four epilogue
    // warpgroups grow to 104 (128 * 56 + 512 * 104 = 60416 <= 61440; an over-subscribed .inc would block forever).
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    // Both single-issuer roles walk the schedule with the WHOLE warp in uniform control flow and let one elected lane
    // execute only the TMA / tcgen05 instructions: addresses, coordinates and descriptors then live in uniform
    // registers.  Inside an `if (lane == 0)` region the compiler wraps every such instruction in a per-lane
    // register -> uniform-register loop; the MMA role then spent ~1300 cycles issuing the four MMAs (512 tensor
    // cycles) of one K block and took a quarter of its sub-partition's issue slots from the epilogue warps
    // (benchmarks/probe_tile_trace.py).
    if (warp == 0) {
        // ===================== TMA producer =====================
        const bool leader = elect_one_sync();
        int stage = 0;
        uint32_t phase = 0;
        uint32_t pli = 0;
        const uint32_t conv_pix = p.conv ? p.cv_OW * p.cv_OH : 1u;
        for (uint32_t ts = tile_first; ts < total_tiles; ts += tile_step, ++pli) {
            const uint32_t t = p.reverse ? total_tiles - 1u - ts : ts;
            const uint32_t b = fd_div(t, fd_tpb), r = t - b * tiles_per_batch;
            const uint32_t mt = fd_div(r, fd_nt), nt = r - mt * n_tiles;
            if (leader && !NQ_DBG(256)) NQ_TRACE(pli, 0);
            // pair: this CTA stages its own 128 rows of A and its half of the B rows
            const int m0 = (int)mt * BMT + (int)cta_rank * BM;
            const int n0 = (int)nt * BN + (TWO ? (int)cta_rank * (BN / 2) : 0);
            const int ba = p.a_batched ? (int)b : 0, bb = p.b_batched ? (int)b : 0;
            // conv: the base pixel of row m0 is (ow * sw, oh * sh) of image m0 / (OH * OW)
            const uint32_t pix = (uint32_t)m0 % conv_pix, img = (uint32_t)m0 / conv_pix;
            const int w0 = p.conv ? (int)((pix % p.cv_OW) * p.cv_sw) : 0, h0 = p.conv ? (int)((pix / p.cv_OW) * p.cv_sh) : 0;
            for (uint32_t kb = 0; kb < k_blocks; ++kb) {
                mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
                const uint32_t fb = smem_u32(full_bar + stage);
                const uint32_t sa = smem_u32(smem_a + stage * C::A_BYTES), sb = smem_u32(smem_b + stage * C::B_BYTES);
                if (leader) {
                    NQ_TRACE_KB(pli, kb, 2);
                    if constexpr (TWO) {
                        if (cta_rank == 0) mbar_expect_tx(fb, 2 * C::STAGE_BYTES);   // bytes of both CTAs land here
                        tma_load_3d_pair(sa, &tmap_a, (int)(kb * BK), m0, ba, fb);
                        tma_load_3d_pair(sb, &tmap_b, (int)(kb * BK), n0, bb, fb);
                    } else if (p.conv) {
                        // two 64-byte K slices per stage: (filter tap, 64 channels) of 128 output pixels each,
                        // 64B-swizzled sub-tiles
                        const uint32_t nsub = ((uint32_t)p.K - kb * BK) >= (uint32_t)BK ? 2u : 1u;
                        mbar_expect_tx(fb, nsub * (C::STAGE_BYTES / 2));
                        for (uint32_t j = 0; j < nsub; ++j) {
                            const uint32_t k0 = kb * BK + j * 64, tap = k0 / p.cv_C;
                            tma_load_im2col_4d(sa + j * (C::A_BYTES / 2), &tmap_a, (int)(k0 % p.cv_C), w0, h0, (int)img,
                                               (uint16_t)(tap % p.cv_KW), (uint16_t)(tap / p.cv_KW), fb);
                            tma_load_3d(sb + j * (C::B_BYTES / 2), &tmap_b, (int)k0, n0, bb, fb);
                        }
                    } else if (NQ_DBG(32)) {
                        mbar_arrive(fb);                                  // measurement only: no operand traffic at all
                    } else {
                        mbar_expect_tx(fb, C::STAGE_BYTES);
                        tma_load_3d(sa, &tmap_a, (int)(kb * BK), m0, ba, fb);
                        tma_load_3d(sb, &tmap_b, (int)(kb * BK), n0, bb, fb);
                    }
                    NQ_TRACE_KB(pli, kb, 3);
                }
                if (++stage == C::STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (leader && !NQ_DBG(256)) NQ_TRACE(pli, 1);
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (cta_rank == 0) {                                              // pair: the leader CTA issues for both
            const bool leader = elect_one_sync();
            constexpr uint32_t idesc = make_idesc(BN, BMT);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            uint32_t mli = 0;
            // K tail: only the 32-byte steps that hold real data (TMA zero-filled the rest of the box)
            const uint32_t krem_last = (uint32_t)p.K - (k_blocks - 1) * BK;
            const int ksteps_last = krem_last >= (uint32_t)BK ? BK / UMMA_K : (int)((krem_last + UMMA_K - 1) / UMMA_K);
            const bool conv = !TWO && p.conv, no_mma = NQ_DBG(8);
            for (uint32_t t = tile_first; t < total_tiles; t += tile_step, ++mli) {
                mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);     // epilogue drained this buffer
                tc_fence_after();
                if (leader && !NQ_DBG(256)) NQ_TRACE(mli, 2);
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (uint32_t kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(smem_u32(full_bar + stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem_a + stage * C::A_BYTES), sb = smem_u32(smem_b + stage * C::B_BYTES);
                    const int ksteps = (kb + 1 == k_blocks) ? ksteps_last : BK / UMMA_K;
                    if (leader) {
                        NQ_TRACE_KB(mli, kb, 0);
                        if (conv) {
                            // 64B-swizzled sub-tiles (one per 64-channel slice), two 32-byte K steps each
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k) {
                                if (k < ksteps) {
                                    const int j = k >> 1;
                                    mma_i8(d_tmem, make_smem_desc_sw64(sa + j * (C::A_BYTES / 2)) + (uint64_t)((k & 1) * 2),
                                           make_smem_desc_sw64(sb + j * (C::B_BYTES / 2)) + (uint64_t)((k & 1) * 2), idesc,
                                           (kb > 0 || k > 0) ? 1u : 0u);
                                }
                            }
                        } else if (!no_mma && ksteps == BK / UMMA_K) {
                            // full K block (every block but a ragged last one): four unconditional MMAs, nothing else --
                            // the issuing warp's own instruction stream must stay well under the 512 cycles the tensor
                            // pipe needs for them, or the pipe idles (the general loop below costs ~110 instructions)
                            const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sb);
                            if constexpr (TWO) {
                                mma_i8_pair(d_tmem, adesc, bdesc, idesc, kb > 0 ? 1u : 0u);
                                mma_i8_pair(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
                                mma_i8_pair(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
                                mma_i8_pair(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
                            } else {
                                mma_i8(d_tmem, adesc, bdesc, idesc, kb > 0 ? 1u : 0u);
                                mma_i8(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
                                mma_i8(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
                                mma_i8(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
                            }
                        } else if (!no_mma) {
                            const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sb);
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k) {
                                // advance 32 bytes along K inside the swizzle atom: +2 in 16-byte units
                                if (k < ksteps) {
                                    if constexpr (TWO)
                                        mma_i8_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                                    (kb > 0 || k > 0) ? 1u : 0u);
                                    else
                                        mma_i8(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                               (kb > 0 || k > 0) ? 1u : 0u);
                                }
                            }
                        }
                        if constexpr (TWO) tc_commit_pair(smem_u32(empty_bar + stage));   // frees the slot in both CTAs
                        else tc_commit(smem_u32(empty_bar + stage));      // smem slot free once MMAs retire
                        NQ_TRACE_KB(mli, kb, 1);
                    }
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (leader) {
                    if constexpr (TWO) tc_commit_pair(smem_u32(tfull_bar + acc));   // both CTAs' epilogues
                    else tc_commit(smem_u32(tfull_bar + acc));            // accumulator ready
                    if constexpr (EMODE == EM_DEQ_WIDE_RES && !TWO) *pace = mli + 1;
                    if (!NQ_DBG(256)) NQ_TRACE(mli, 3);
                }
                if (++acc == C::NACC) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp == 3) {
        // ===================== residual prefetcher (float32 + residual epilogue only) =====================
        // The epilogue warps can hold only a few residual loads in flight each (registers), so the residual stream's
        // HBM latency was exposed about four times per tile: the output projection took 86 us against 45 us without the
        // residual Add (benchmarks/probe_residual_gemm.py).  This otherwise idle warp pulls the residual block of the
        // tile that is being multiplied into L2 (one bulk prefetch per row), one tile ahead of the epilogue and paced by
        // the accumulator barriers, so that the epilogue's loads are L2 hits.
        // (not for CTA pairs: the K >= 1024 launches they serve are not bound by the residual stream -- MLP-2 measured
        // 113.8 us without and 117.6 us with the prefetcher)
        if constexpr (EMODE == EM_DEQ_WIDE_RES && !TWO) {
            uint32_t li = 0;
            for (uint32_t ts = tile_first; ts < total_tiles; ts += tile_step, ++li) {
                const uint32_t t = p.reverse ? total_tiles - 1u - ts : ts;
                const uint32_t b = fd_div(t, fd_tpb), r = t - b * tiles_per_batch;
                const uint32_t mt = fd_div(r, fd_nt), nt = r - mt * n_tiles;
                const int64_t m0 = (int64_t)mt * BMT + cta_rank * BM, n0 = (int64_t)nt * BN;
                const int64_t cols = (p.N - n0) < BN ? (p.N - n0) : BN;
                const float* base = p.residual + (int64_t)b * p.stride_r + n0;
                // pace: the MMA warp has issued tile li - 1 (a monotonic counter: this warp may lag without harm)
                for (uint32_t spin = 0; li >= 1; ++spin) {
                    if (*pace >= li) break;
                    __nanosleep(200);
                    if (spin > (1u << 24)) __trap();
                }
                for (int rr = lane; rr < BM; rr += 32)
                    if (m0 + rr < p.M) prefetch_l2_bulk(base + (m0 + rr) * p.ldr, (uint32_t)(cols * 4));
            }
        }
        __syncwarp();
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ===================== epilogue (16 warps) =====================
        // The epilogue is a long dependent instruction stream per element, so it is throughput-
        // bound by how many warps each SM sub-partition can interleave: 16 warps (4 per scheduler),
        // each owning a TMEM lane quarter q (hardware rule: warp_id % 4) and every 4th 16-column
        // sub-chunk (h).  Unit of work: 32 rows x 16 columns -> tcgen05.ld x16 -> XOR-swizzled smem
        // slab (2 KB per warp) -> row-contiguous read-back (8 rows x 64 B per store instruction) with the
        // zero-point correction, dequantize, bias and residual applied branch-free on the way out.
        // EMODE is a template parameter so that each instantiation only carries its own path.
        const int q = warp & 3;
        const int h = (warp - 4) >> 2;                                    // 0..3
        const int ew = warp - 4;
        uint32_t* stg = epi + ew * (32 * 16);                             // 32 rows x 16 words
        uint32_t* stg_row = epi + NUM_EPI_WARPS * C::SLAB_WORDS + ew * 32;
        constexpr int HPG = 4 / C::GROUPS;                                // column slots (values of h) per group
        const int grp = h / HPG, hc = h % HPG;                            // tile group, 64-column slot in the tile
        const AccZp z = p.zp;
        const bool c_aligned = ((p.ldc & 3) == 0) && ((p.stride_c & 3) == 0) && ((p.stride_c_inner & 3) == 0) &&
                               ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
        const int cl = lane & 3, rsub = lane >> 2;                       // vector read-back: 4-col chunk, row in group of 8
        const int sc_col = lane & 15, sc_row = lane >> 4;                 // scalar read-back: column, row in pair
        const bool cs_vec = z.use_col && ((reinterpret_cast<uintptr_t>(z.colsum_b) & 15) == 0) && ((z.cs_stride & 3) == 0);
        const bool bias_vec = p.bias_f32 && ((reinterpret_cast<uintptr_t>(p.bias_f32) & 15) == 0);
        const bool res_vec = p.residual && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0) && ((p.ldr & 3) == 0) &&
                             ((p.stride_r & 3) == 0);
        const int zpa = (int)z.zp_a;
        constexpr bool FASTF = (EMODE == EM_DEQ_FAST);                   // float result via the 32-bit fast path
        constexpr bool REQ = (EMODE == EM_REQ_ROWS);                     // Gemm: int bias + requantize, plain [M, N] int8 rows
        constexpr bool Q8 = (EMODE == EM_Q8_ROWS || EMODE == EM_Q8_COLS || EMODE == EM_Q8_GELU || REQ);
        const Quantizer qz(p.qargs);
        const UDiv fd_S = make_udiv(Q8 ? p.q_S : 1u), fd_D = make_udiv(Q8 ? p.q_D : 1u), fd_ci = make_udiv((uint32_t)p.c_inner);
        const int qlo = (int)p.qargs.lo, qhi = (int)p.qargs.hi;          // integer code range of the QUANT epilogues
        constexpr bool SOFTMAX = (EMODE == EM_SOFTMAX_SYM || EMODE == EM_SOFTMAX_ASYM);
        // Per-tile operands of the zero-point correction (this row's rowsum(A), this warp's colsum(B) and bias
        // columns) are fetched ONE TILE AHEAD: with 227 KB of smem there is no L1 to hit, so each of these
        // loads is an L2 round trip that would otherwise sit at the head of every tile's critical path.
        struct TilePre { int rowsum, c0, c1, b0, b1, wide; uint32_t b, m0, n0; };
        auto prefetch_tile = [&](uint32_t tt, TilePre& o) {
            o.rowsum = o.c0 = o.c1 = o.b0 = o.b1 = o.wide = 0;
            o.b = o.m0 = o.n0 = 0;
            if (tt >= total_tiles) return;
            if (p.reverse) tt = total_tiles - 1u - tt;
            const uint32_t pb = fd_div(tt, fd_tpb), pr = tt - pb * tiles_per_batch;
            const uint32_t pmt = fd_div(pr, fd_nt);
            const uint32_t pm0 = pmt * BMT + cta_rank * BM, pn0 = (pr - pmt * n_tiles) * BN;
            o.b = pb; o.m0 = pm0; o.n0 = pn0;
            const int64_t pm = (int64_t)pm0 + q * 32 + lane;
            if (z.use_row && pm < p.M) o.rowsum = ldg_s32(z.rowsum_a + (int64_t)pb * p.M + pm);
            if constexpr (SOFTMAX || Q8) {
                constexpr int W = SOFTMAX ? 56 : 64;                      // columns owned by this warp
                const int64_t c0i = (int64_t)pn0 + hc * W + lane, c1i = c0i + 32;
                const int32_t* pcs = z.use_col ? z.colsum_b + (int64_t)pb * z.cs_stride : nullptr;
                if (lane < W && c0i < p.N) {
                    if (pcs) o.c0 = ldg_s32(pcs + c0i);
                    if (Q8 && !REQ && p.bias_f32) o.b0 = ldg_s32(p.bias_f32 + c0i);
                    if (REQ && p.bias_q) {
                        const int64_t bq = __ldg(p.bias_q + c0i);
                        o.b0 = (int)bq;
                        o.wide |= (bq < -(1ll << 29) || bq > (1ll << 29)) ? 1 : 0;
                    }
                }
                if (lane + 32 < W && c1i < p.N) {
                    if (pcs) o.c1 = ldg_s32(pcs + c1i);
                    if (Q8 && !REQ && p.bias_f32) o.b1 = ldg_s32(p.bias_f32 + c1i);
                    if (REQ && p.bias_q) {
                        const int64_t bq = __ldg(p.bias_q + c1i);
                        o.b1 = (int)bq;
                        o.wide |= (bq < -(1ll << 29) || bq > (1ll << 29)) ? 1 : 0;
                    }
                }
            }
        };
        TilePre nxt;
        prefetch_tile(tile_first + grp * tile_step, nxt);
        // this CTA's (pair's) i-th tile is t = tile_first + i * tile_step, lives in accumulator buffer i % NACC and is
        // drained by group i % GROUPS
        for (uint32_t li = grp, ts = tile_first + grp * tile_step; ts < total_tiles; li += C::GROUPS, ts += C::GROUPS * tile_step) {
            const int acc = (int)(li % C::NACC);
            const uint32_t acc_phase = (li / C::NACC) & 1u;
            const TilePre cur = nxt;                                      // coordinates and operands of this tile
            const int64_t b = cur.b;
            const int64_t m0 = cur.m0, n0 = cur.n0;
            const int64_t mrow0 = m0 + q * 32;
            const int64_t m = mrow0 + lane;                               // this thread's accumulator row
            const bool row_ok = m < p.M;
            const int32_t* cs_b = z.use_col ? z.colsum_b + b * z.cs_stride : nullptr;
            prefetch_tile(ts + C::GROUPS * tile_step, nxt);
            int64_t rowterm = -z.kterm;
            if (z.use_row && row_ok) rowterm += (int64_t)cur.rowsum * z.zp_b;
            if constexpr (Q8) {
                // ---- int8 operand outputs (NQ_EPI_QUANT / NQ_EPI_GELU_QUANT): thread = accumulator row.
                // The float value the graph would have produced (bias + dequant [-> GELU]) is quantized with
                // the consumer's parameters and written straight into the consumer's K-major operand: no
                // smem transpose, the row part of every address is formed once per tile, column terms and
                // bias are staged per warp in shared memory and read back as broadcast LDS.128.
                //   ROWS / GELU: 16 consecutive n are 16 consecutive bytes (one 16-byte store per row)
                //   COLS       : consecutive m are consecutive bytes (transposing scatter, V operand)
                constexpr int CPW = 4;                                    // 16-column chunks per warp, contiguous
                constexpr int WCOLS = CPW * 16;
                const int wcol0 = hc * WCOLS;
                int* ctw = reinterpret_cast<int*>(epi) + ew * 128;        // [64] colsum * zp_a
                float* bsw = reinterpret_cast<float*>(ctw + 64);          // [64] bias
                __syncwarp();
                // REQUANT: the integer bias joins the column term (acc + bias - zero-point, all int32 by the host bound)
                if (lane < WCOLS) {
                    ctw[lane] = cur.c0 * zpa - (REQ ? cur.b0 : 0);
                    bsw[lane] = __int_as_float(cur.b0);
                }
                if (lane + 32 < WCOLS) {
                    ctw[lane + 32] = cur.c1 * zpa - (REQ ? cur.b1 : 0);
                    bsw[lane + 32] = __int_as_float(cur.b1);
                }
                const bool wide_bias = REQ && __any_sync(0xffffffffu, cur.wide != 0);   // some |bias| >= 2^29: general route
                // row part of the scatter offsets (m = mb * S + ms, batch = bo * inner + bi)
                const uint32_t mu = (uint32_t)(row_ok ? m : p.M - 1);
                const uint32_t mb = fd_div(mu, fd_S), ms = mu - mb * p.q_S;
                const uint32_t bu = (uint32_t)b, bo = fd_div(bu, fd_ci), bi = bu - bo * (uint32_t)p.c_inner;
                const int64_t row_off = (int64_t)bo * p.q_off[0] + (int64_t)bi * p.q_off[1] + (int64_t)mb * p.q_off[2] +
                                        (int64_t)ms * p.q_off[3];
                int64_t row_rs = 0;
                if (p.q_rowsum)
                    row_rs = (int64_t)bo * p.q_rs[0] + (int64_t)bi * p.q_rs[1] + (int64_t)mb * p.q_rs[2] + (int64_t)ms * p.q_rs[3];
                __syncwarp();
                if (warp == 4 && lane == 0 && !NQ_DBG(256)) NQ_TRACE(li, 4);
                mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
                tc_fence_after();
                if (warp == 4 && lane == 0 && !NQ_DBG(256)) NQ_TRACE(li, 5);
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
                const bool f24 = !REQ && p.fast24 != 0;                   // host bound: exact int -> float conversion (I2FP)
                const int rm = (f24 ? 0 : 0x4B400000) - (int32_t)rowterm; // else: int -> float magic folded into the row term
                int rs_acc = 0;                                           // ROWS: code sum of the current head
                uint32_t rs_nh = 0xffffffffu;
                auto flush_rowsum = [&]() {
                    if (rs_nh != 0xffffffffu && row_ok) {
                        int32_t* slot = p.q_rowsum + row_rs + (int64_t)rs_nh * p.q_rs[4];
                        if (p.q_rs_exclusive) *slot = rs_acc;
                        else atomicAdd(slot, rs_acc);
                    }
                };
                uint32_t nh = fd_div((uint32_t)(n0 + wcol0), fd_D), nd = (uint32_t)(n0 + wcol0) - nh * p.q_D;
                // Everything the chunk loop needs that does not depend on the chunk is formed here, once per tile, and
                // pinned in registers (the compiler otherwise re-derives thread index, row bound and shared-memory
                // window for every chunk): destination row, row predicate, chunk count, TMEM / shared addresses.
                const int rem_cols = (int)(p.N - (n0 + wcol0));
                int nchunks = rem_cols <= 0 ? 0 : min(CPW, (rem_cols + 15) >> 4);   // warp-uniform (N % 16 == 0)
                if (NQ_DBG(16)) nchunks = 0;
                int8_t* dst_row = reinterpret_cast<int8_t*>(p.C) + row_off;
                int row_ok_i = row_ok ? 1 : 0;
                uint32_t t_col = t_row + (uint32_t)wcol0, ctw_s = smem_u32(ctw);
                const uint32_t q_off4 = (uint32_t)p.q_off[4];             // host: (N / q_D + 1) * q_off[4] < 2^31
                asm volatile("" : "+r"(row_ok_i), "+r"(t_col), "+r"(ctw_s), "+l"(dst_row));
#pragma unroll 1
                for (int i = 0; i < nchunks; ++i) {
                    const int64_t nc = n0 + wcol0 + i * 16;
                    uint32_t v[16];
                    if (!NQ_DBG(2)) tmem_ld_32x32b_x16(t_col, v);
                    else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = (uint32_t)(i + j);
                    }
                    t_col += 16;
                    const uint32_t cs = ctw_s + (uint32_t)i * 64u;        // this step's 16 column terms; + 256: its bias columns
                    int ct[16];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int4 c4 = lds_v4(cs + g * 16);
                        ct[4 * g] = c4.x; ct[4 * g + 1] = c4.y; ct[4 * g + 2] = c4.z; ct[4 * g + 3] = c4.w;
                    }
                    tmem_ld_wait();
                    int x[16];
                    uint32_t bad = 0;
                    float f[16];
                    if (f24) {
                        // |acc - zero-point terms| < 2^24 proved on the host: the plain conversion is exact, no window test
                        const float2 sc2 = make_float2(p.scale, p.scale);
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            float2 t = make_float2(__int2float_rn((int)v[j] + rm - ct[j]), __int2float_rn((int)v[j + 1] + rm - ct[j + 1]));
                            if constexpr (EMODE != EM_Q8_GELU) t = __fmul2_rn(t, sc2);       // GELU: scaled below
                            f[j] = t.x;
                            f[j + 1] = t.y;
                        }
                    } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        x[j] = (int)v[j] + rm - ct[j];
                        bad |= (uint32_t)(x[j] ^ 0x4B000000);
                    }
                    if constexpr (REQ) {
                        if (__builtin_expect(wide_bias || __any_sync(0xffffffffu, (bad & 0xFF800000u) != 0), 0)) {
                            // general 64-bit route for this 16-column step (whole warp: the store below is per thread)
                            uint32_t wq[4] = {0, 0, 0, 0};
#pragma unroll 1
                            for (int j = 0; j < 16; ++j) {
                                const int c = requant_slow((int)v[j], rowterm, z, b, nc + j, p.bias_q, p.scale, p.inv_out_scale,
                                                           p.asym_out, p.out_zp, p.lo, p.hi);
                                wq[j >> 2] |= (uint32_t)(c & 0xff) << ((j & 3) * 8);
                            }
                            if (i > 0) {
                                nd += 16;
                                if (nd >= p.q_D) { nd = 0; ++nh; }
                            }
                            if (row_ok)
                                *reinterpret_cast<int4*>(reinterpret_cast<int8_t*>(p.C) + row_off + (int64_t)nh * p.q_off[4] + nd) =
                                    make_int4((int)wq[0], (int)wq[1], (int)wq[2], (int)wq[3]);
                            continue;
                        }
                    }
                    if (__builtin_expect((bad & 0xFF800000u) != 0, 0)) {  // some |d| >= 2^22: exact slow route
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = deq_slow(x[j] - 0x4B400000, p.scale);
                    } else {
                        // two columns per instruction: a packed add and a packed multiply round each lane exactly like
                        // the scalar pair (the bias add below stays scalar: ptxas would contract a packed multiply +
                        // packed add into one FFMA2, i.e. a single rounding)
                        const float2 nm = make_float2(-12582912.0f, -12582912.0f), sc2 = make_float2(p.scale, p.scale);
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            float2 t = __fadd2_rn(make_float2(__int_as_float(x[j]), __int_as_float(x[j + 1])), nm);
                            if constexpr (EMODE != EM_Q8_GELU) t = __fmul2_rn(t, sc2);       // GELU: scaled below
                            f[j] = t.x;
                            f[j + 1] = t.y;
                        }
                    }
                    }
                    const bool gelu_scaled = (bad & 0xFF800000u) != 0;    // warp-divergent only on the rare slow route
                    // Quantizer tail, two columns per instruction: the quotient t = y / s_out (ROWS / COLS: correctly
                    // rounded -- q0 = y * RN(1 / s), two fused residual corrections, common.cuh; GELU: already folded into
                    // the chain's last constant) is rounded half-to-even by ONE add of 1.5 * 2^23 + zp; the integer
                    // n = zp + rint(t) sits in the low bits of the sum and the saturating pack instruction clamps it to
                    // the int8 range (narrower codes: an integer min / max first).  A quotient below -2^22 would leave
                    // the magic window on the wrong side (n would wrap), hence the one-sided max; above it the bit
                    // pattern only grows, which saturates correctly.
                    const float2 r2 = make_float2(qz.sd.r, qz.sd.r), nb2 = make_float2(-qz.sd.b, -qz.sd.b), mg2 = make_float2(qz.magic, qz.magic);
                    uint32_t w[4];
                    if (NQ_DBG(1)) {
                        w[0] = __float_as_uint(f[0]); w[1] = __float_as_uint(f[5]); w[2] = __float_as_uint(f[10]); w[3] = __float_as_uint(f[15]);
                    } else
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int4 b4i = lds_v4(cs + 256 + g * 16);       // bias columns of this step (bsw = ctw + 64 words)
                        const float4 b4 = make_float4(__int_as_float(b4i.x), __int_as_float(b4i.y), __int_as_float(b4i.z),
                                                      __int_as_float(b4i.w));
                        float2 ta, tb;
                        if constexpr (EMODE == EM_Q8_GELU) {
                            // float glue (1e-5 contract): dequant * scale + bias as one FMA, GELU, and the
                            // quantizer's division folded into the chain's last constant; two elements per FFMA2
                            const float sc = gelu_scaled ? 1.0f : p.scale;
                            const float2 sc2 = make_float2(sc, sc);
                            ta = gelu_fast2(__ffma2_rn(make_float2(f[4 * g], f[4 * g + 1]), sc2, make_float2(b4.x, b4.y)),
                                            p.g_prdiv, p.g_nl2e, p.g_add, p.g_out);
                            tb = gelu_fast2(__ffma2_rn(make_float2(f[4 * g + 2], f[4 * g + 3]), sc2, make_float2(b4.z, b4.w)),
                                            p.g_prdiv, p.g_nl2e, p.g_add, p.g_out);
                        } else if constexpr (REQ) {
                            // requantize: t = (1 / s_out) * dequant, a float32 multiply by the IEEE-rounded reciprocal
                            const float2 inv2 = make_float2(p.inv_out_scale, p.inv_out_scale);
                            ta = __fmul2_rn(inv2, make_float2(f[4 * g], f[4 * g + 1]));
                            tb = __fmul2_rn(inv2, make_float2(f[4 * g + 2], f[4 * g + 3]));
                            (void)b4;
                        } else {
                            // bias adds stay scalar: ptxas contracts a packed multiply followed by a packed add into one
                            // FFMA2 (a single rounding) even when both carry .rn
                            const float2 ya = make_float2(__fadd_rn(b4.x, f[4 * g]), __fadd_rn(b4.y, f[4 * g + 1]));
                            const float2 yb = make_float2(__fadd_rn(b4.z, f[4 * g + 2]), __fadd_rn(b4.w, f[4 * g + 3]));
                            ta = __fmul2_rn(ya, r2);
                            tb = __fmul2_rn(yb, r2);
                            float2 ea = __ffma2_rn(ta, nb2, ya), eb = __ffma2_rn(tb, nb2, yb);
                            ta = __ffma2_rn(ea, r2, ta);
                            tb = __ffma2_rn(eb, r2, tb);
                            ea = __ffma2_rn(ta, nb2, ya);
                            eb = __ffma2_rn(tb, nb2, yb);
                            ta = __ffma2_rn(ea, r2, ta);
                            tb = __ffma2_rn(eb, r2, tb);
                        }
                        if constexpr (EMODE == EM_Q8_GELU) {
                            // GELU >= -0.2403 c1 c3: the quotient cannot leave the magic window downwards (host-checked)
                            ta = __fadd2_rn(ta, mg2);
                            tb = __fadd2_rn(tb, mg2);
                        } else {
                            ta = __fadd2_rn(make_float2(fmaxf(ta.x, -4194304.0f), fmaxf(ta.y, -4194304.0f)), mg2);
                            tb = __fadd2_rn(make_float2(fmaxf(tb.x, -4194304.0f), fmaxf(tb.y, -4194304.0f)), mg2);
                        }
                        int n0q = __float_as_int(ta.x) - 0x4B400000, n1q = __float_as_int(ta.y) - 0x4B400000;
                        int n2q = __float_as_int(tb.x) - 0x4B400000, n3q = __float_as_int(tb.y) - 0x4B400000;
                        if (qlo != -128) {                                 // codes narrower than 8 bits (warp-uniform)
                            n0q = min(max(n0q, qlo), qhi); n1q = min(max(n1q, qlo), qhi);
                            n2q = min(max(n2q, qlo), qhi); n3q = min(max(n3q, qlo), qhi);
                        }
                        uint32_t hi16;
                        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi16) : "r"(n3q), "r"(n2q), "r"(0));
                        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w[g]) : "r"(n1q), "r"(n0q), "r"(hi16));
                    }
                    if (i > 0) {                                          // next 16 columns: step (head, column in head)
                        nd += 16;
                        if (nd >= p.q_D) { nd = 0; ++nh; }
                    }
                    if constexpr (EMODE == EM_Q8_COLS) {
                        // out[.., nd + j, ms]: consecutive lanes are consecutive bytes
                        int8_t* dst = reinterpret_cast<int8_t*>(p.C) + row_off + (int64_t)nh * p.q_off[4] + (int64_t)nd * p.q_off[5];
                        if (row_ok) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) dst[(int64_t)j * p.q_off[5]] = (int8_t)(w[j >> 2] >> ((j & 3) * 8));
                        }
                        if (p.q_rowsum) {
                            // code sums over the rows of each image present in this warp's 32 rows
                            const uint32_t mb_first = __shfl_sync(0xffffffffu, mb, 0), mb_last = __shfl_sync(0xffffffffu, mb, 31);
                            for (uint32_t img = mb_first; img <= mb_last; ++img) {
                                const bool mine_img = row_ok && mb == img;
                                int keep = 0;
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    const int c = mine_img ? (int)(int8_t)(w[j >> 2] >> ((j & 3) * 8)) : 0;
                                    const int tot = __reduce_add_sync(0xffffffffu, c);
                                    keep = (lane == j) ? tot : keep;
                                }
                                if (lane < 16) {
                                    const int64_t slot = (int64_t)bo * p.q_rs[0] + (int64_t)bi * p.q_rs[1] + (int64_t)img * p.q_rs[2] +
                                                         (int64_t)nh * p.q_rs[4] + (int64_t)(nd + lane) * p.q_rs[5];
                                    atomicAdd(p.q_rowsum + slot, keep);
                                }
                            }
                        }
                    } else {
                        if (row_ok_i && !NQ_DBG(4))
                            *reinterpret_cast<int4*>(dst_row + (nh * q_off4 + nd)) = make_int4((int)w[0], (int)w[1], (int)w[2], (int)w[3]);
                        if (p.q_rowsum) {
                            if (nh != rs_nh) {
                                flush_rowsum();
                                rs_acc = 0;
                                rs_nh = nh;
                            }
                            rs_acc = __dp4a((int)w[0], 0x01010101, __dp4a((int)w[1], 0x01010101, __dp4a((int)w[2], 0x01010101, __dp4a((int)w[3], 0x01010101, rs_acc))));
                        }
                    }
                }
                if (EMODE != EM_Q8_COLS && p.q_rowsum) flush_rowsum();
                tc_fence_before();
                __syncwarp();
                if (NQ_DBG(256)) {                                        // alternative trace: release stamps of warps 4..7 / 16..19
                    if (lane == 0 && warp < 8) NQ_TRACE(li, warp);
                    if (lane == 0 && warp >= 16) NQ_TRACE(li, warp - 16);
                } else if (lane == 0 && (warp == 4 || warp == 19)) NQ_TRACE(li, warp == 4 ? 6 : 7);
                if (lane == 0) { if constexpr (TWO) mbar_arrive_cluster(smem_u32(tempty_bar + acc), 0); else mbar_arrive(smem_u32(tempty_bar + acc)); }
                continue;
            }
            if constexpr (EMODE == EM_DEQ_WIDE || EMODE == EM_DEQ_WIDE_RES) {
                // ---- float32 output (dequant + bias + residual), host-checked alignment, N % 32 == 0.
                // 32 rows x 32 columns per step: tcgen05.ld x32 -> XOR-swizzled 4 KB slab (rows of 128 B, 16-byte
                // chunk ^ (row & 7): conflict-free both ways) -> read-back where one warp instruction covers 4 rows
                // x 128 B, so the residual LDG.128 and the result STG.128 move whole 128-byte lines.
                uint32_t* slab = epi + ew * C::SLAB_WORDS;
                const int cj = lane & 7, rq = lane >> 3;                   // read-back: 16-byte chunk, row within 4
                constexpr bool has_res = (EMODE == EM_DEQ_WIDE_RES);
                int4 ct_n = make_int4(0, 0, 0, 0), bs_n = make_int4(0, 0, 0, 0);
                auto fetch32 = [&](int sidx) {
                    const int64_t nc = n0 + sidx * 32;
                    if (sidx >= BN / 32 || nc >= p.N) return;
                    ct_n = cs_b ? ldg_v4(cs_b + nc + cj * 4) : make_int4(0, 0, 0, 0);
                    bs_n = p.bias_f32 ? ldg_v4(p.bias_f32 + nc + cj * 4) : make_int4(0, 0, 0, 0);
                };
                fetch32(hc);
                mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
                tc_fence_after();
                __syncwarp();
                stg_row[lane] = (uint32_t)(0x4B400000 - (int32_t)rowterm);   // row term + int->float magic
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
                const int64_t crow_base = (p.c_inner > 1) ? (b / p.c_inner) * p.stride_c + (b % p.c_inner) * p.stride_c_inner
                                                          : b * p.stride_c;
                const int rows_left = (int)((p.M - mrow0) < 32 ? ((p.M - mrow0) > 0 ? (p.M - mrow0) : 0) : 32);
#pragma unroll 1
                for (int sidx = hc; sidx < BN / 32; sidx += HPG) {
                    const int64_t nc = n0 + sidx * 32;
                    if (nc >= p.N || rows_left == 0) break;               // warp-uniform
                    // residual rows of the first half (16 rows) leave for L2 / HBM before the accumulator round trip:
                    // their latency overlaps the TMEM load and the slab staging; the second half is issued right
                    // after the staging and overlaps the first half's arithmetic
                    const float* rbase = has_res ? p.residual + b * p.stride_r + mrow0 * p.ldr + nc + cj * 4 : nullptr;
                    float4 res0[4], res1[4];
                    auto load_res = [&](int half, float4* dstv) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            // rows past M are clamped to the last valid one (loaded, never stored)
                            const int rr = (half * 4 + k) * 4 + rq;
                            const int rc = rr < rows_left ? rr : rows_left - 1;
                            dstv[k] = __ldcs(reinterpret_cast<const float4*>(rbase + (int64_t)rc * p.ldr));
                        }
                    };
                    if constexpr (has_res) load_res(0, res0);
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_row + (uint32_t)(sidx * 32), v);
                    const int ct0 = ct_n.x * zpa, ct1 = ct_n.y * zpa, ct2 = ct_n.z * zpa, ct3 = ct_n.w * zpa;
                    const float b0 = __int_as_float(bs_n.x), b1 = __int_as_float(bs_n.y);
                    const float b2 = __int_as_float(bs_n.z), b3 = __int_as_float(bs_n.w);
                    fetch32(sidx + HPG);
                    tmem_ld_wait();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4*>(slab + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                            make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    __syncwarp();
                    uint32_t bad_any = 0;
                    float* crow = reinterpret_cast<float*>(p.C) + crow_base + (mrow0 + rq) * p.ldc + nc + cj * 4;
                    if constexpr (has_res) load_res(1, res1);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const float4* res = half ? res1 : res0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int it = half * 4 + k;
                            const int rr = it * 4 + rq;
                            const uint4 val = *reinterpret_cast<const uint4*>(slab + rr * 32 + ((cj ^ (rr & 7)) << 2));
                            const int rm = (int)stg_row[rr];
                            const int x0 = (int)val.x + rm - ct0, x1 = (int)val.y + rm - ct1;
                            const int x2 = (int)val.z + rm - ct2, x3 = (int)val.w + rm - ct3;
                            bad_any |= (uint32_t)(x0 ^ 0x4B000000) | (uint32_t)(x1 ^ 0x4B000000) |
                                       (uint32_t)(x2 ^ 0x4B000000) | (uint32_t)(x3 ^ 0x4B000000);
                            // packed add / packed multiply (each lane rounded like the scalar op); the bias add is
                            // scalar so that ptxas cannot contract multiply + add into a single-rounding FFMA2
                            const float2 nm = make_float2(-12582912.0f, -12582912.0f), sc2 = make_float2(p.scale, p.scale);
                            const float2 fa = __fmul2_rn(__fadd2_rn(make_float2(__int_as_float(x0), __int_as_float(x1)), nm), sc2);
                            const float2 fb = __fmul2_rn(__fadd2_rn(make_float2(__int_as_float(x2), __int_as_float(x3)), nm), sc2);
                            float f0 = fa.x, f1 = fa.y, f2 = fb.x, f3 = fb.y;
                            if (p.bias_f32) {
                                f0 = __fadd_rn(b0, f0); f1 = __fadd_rn(b1, f1); f2 = __fadd_rn(b2, f2); f3 = __fadd_rn(b3, f3);
                            }
                            if constexpr (has_res) {
                                const float2 ra = __fadd2_rn(make_float2(f0, f1), make_float2(res[k].x, res[k].y));
                                const float2 rb = __fadd2_rn(make_float2(f2, f3), make_float2(res[k].z, res[k].w));
                                f0 = ra.x; f1 = ra.y; f2 = rb.x; f3 = rb.y;
                            }
                            if (rr < rows_left)
                                *reinterpret_cast<float4*>(crow + (int64_t)it * 4 * p.ldc) = make_float4(f0, f1, f2, f3);
                        }
                    }
                    // some |acc - zero-point terms| >= 2^22 (outside the magic-constant conversion): this lane's part
                    // of the step is recomputed on the exact general route, out of line and off the hot path
                    if (__builtin_expect((bad_any & 0xFF800000u) != 0, 0))
                        deq_wide_fix(slab, stg_row, lane, ct0, ct1, ct2, ct3, b0, b1, b2, b3, p.bias_f32 != nullptr, rbase,
                                     p.ldr, crow, p.ldc, rows_left, p.scale);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if constexpr (TWO) mbar_arrive_cluster(smem_u32(tempty_bar + acc), 0); else mbar_arrive(smem_u32(tempty_bar + acc)); }
                continue;
            }
            // Column operands (colsum for the zero-point, bias) of the NEXT sub-chunk are always in flight
            // one step ahead (first one: issued before waiting for the accumulator): with 227 KB of
            // smem there is no L1 to hit, so every one of these loads is an L2 round trip.
            int4 ct_n = make_int4(0, 0, 0, 0), bs_n = make_int4(0, 0, 0, 0);
            auto fetch_cols = [&](int sidx) {
                ct_n = make_int4(0, 0, 0, 0);
                bs_n = make_int4(0, 0, 0, 0);
                const int64_t nc = n0 + sidx * 16;
                if (sidx >= BN / 16 || nc >= p.N) return;
                if (c_aligned && nc + 16 <= p.N) {
                    if (cs_b) ct_n = cs_vec ? ldg_v4(cs_b + nc + cl * 4)
                                            : make_int4(ldg_s32(cs_b + nc + cl * 4), ldg_s32(cs_b + nc + cl * 4 + 1),
                                                        ldg_s32(cs_b + nc + cl * 4 + 2), ldg_s32(cs_b + nc + cl * 4 + 3));
                    if (p.bias_f32) {
                        const float* bp = p.bias_f32 + nc + cl * 4;
                        bs_n = bias_vec ? ldg_v4(bp) : make_int4(ldg_s32(bp), ldg_s32(bp + 1), ldg_s32(bp + 2), ldg_s32(bp + 3));
                    }
                } else if (nc + sc_col < p.N) {
                    if (cs_b) ct_n.x = ldg_s32(cs_b + nc + sc_col);
                    if (p.bias_f32) bs_n.x = ldg_s32(p.bias_f32 + nc + sc_col);
                }
            };
            if (FASTF) fetch_cols(hc);
            if (SOFTMAX) {
                // this warp's 56 column terms (colsum * zp_a, fetched one tile ahead) -> its private smem strip,
                // read back as broadcast LDS in the softmax loop
                int* ctw = reinterpret_cast<int*>(epi) + 1024 + ew * 64;
                __syncwarp();
                ctw[lane] = cur.c0 * zpa;
                ctw[32 + lane] = cur.c1 * zpa;
                __syncwarp();
            }
            mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
            tc_fence_after();
            if (FASTF) {
                __syncwarp();
                // row term with the int->float magic constant folded in: x = acc + rowmagic - colterm
                stg_row[lane] = (uint32_t)(0x4B400000 - (int32_t)rowterm);
            }
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
            const int64_t crow_base = (p.c_inner > 1) ? (b / p.c_inner) * p.stride_c + (b % p.c_inner) * p.stride_c_inner
                                                      : b * p.stride_c;
            const int rows_left = (int)((p.M - mrow0) < 32 ? ((p.M - mrow0) > 0 ? (p.M - mrow0) : 0) : 32);
            if (EMODE == EM_SOFTMAX_SYM || EMODE == EM_SOFTMAX_ASYM) {
                // ---- attention scores: dequantize (/ c folded into the scale) -> softmax over the row ->
                // quantize, all in registers.  N <= 224 fits one tile; the 4 warps of a lane quarter split
                // the columns (56 contiguous columns each) and exchange row max / row sum / code sum
                // through shared memory.  Float glue under the 1e-5 contract: the
                // probabilities are e_i * RN(1 / (sum * s_out)) with e_i = ex2(fma(y_i, log2 e, -max*log2 e));
                // the codes are the exact round-half-even of that quotient and the row sums are the exact
                // integer sums of the emitted codes.
                constexpr int QS = (EMODE == EM_SOFTMAX_ASYM) ? 1 : 0;
                constexpr int NSUB = 7;                                    // 7 x 8 columns per warp
                constexpr float kMasked = -1.0e30f;                        // exp() of it is exactly 0, no inf - inf
                const int col0 = h * (NSUB * 8);
                float* red = reinterpret_cast<float*>(epi);                // [2][4][128] max / sum exchange
                int* redq = reinterpret_cast<int*>(epi) + 2048;            // [4][128] code sums
                const int rloc = q * 32 + lane;
                const int ncols_w = (int)(p.N - col0 < NSUB * 8 ? (p.N - col0 > 0 ? p.N - col0 : 0) : NSUB * 8);
                float y[NSUB * 8];
                float lmax = kMasked;
                const bool warp_rows = rows_left > 0;                       // tcgen05.ld is warp-collective: uniform guard
                const int* ctw = reinterpret_cast<const int*>(epi) + 1024 + ew * 64;
                auto pass1 = [&](auto magic_tag) {
                    constexpr bool MAGIC = decltype(magic_tag)::value;
                    // MAGIC: |d| < 2^22 proved on the host -> x is the bit pattern of 1.5*2^23 + d
                    const int rm = (MAGIC ? 0x4B400000 : 0) - (int32_t)rowterm;
                    // accumulator columns in steps of 16 (8 for the odd last group): half as many TMEM round trips
#pragma unroll
                    for (int j2 = 0; j2 < NSUB; j2 += 2) {
                        if (j2 * 8 < ncols_w) {                           // warp-uniform
                            uint32_t a16[16];
                            const bool two = (j2 + 1 < NSUB) && ((j2 + 1) * 8 < ncols_w);
                            if (two) tmem_ld_32x32b_x16(t_row + (uint32_t)(col0 + j2 * 8), a16);
                            else tmem_ld_32x32b_x8(t_row + (uint32_t)(col0 + j2 * 8), a16);
                            // column terms staged by this warp before the accumulator wait (broadcast LDS)
                            int c16[16];
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const int4 c4 = *reinterpret_cast<const int4*>(ctw + j2 * 8 + g * 4);
                                c16[4 * g] = c4.x; c16[4 * g + 1] = c4.y; c16[4 * g + 2] = c4.z; c16[4 * g + 3] = c4.w;
                            }
                            tmem_ld_wait();
                            // columns past N only in the last (ragged) pair of groups: warp-uniform fast path
                            const bool clean = j2 * 8 + 16 <= ncols_w || (j2 + 1 >= NSUB && j2 * 8 + 8 <= ncols_w);
                            if (clean) {
                                // two columns per instruction (packed add / multiply, each lane rounded like the scalar op)
#pragma unroll
                                for (int k = 0; k < 16; k += 2) {
                                    if (j2 * 8 + k >= NSUB * 8) break;    // compile time: the 7th group is 8 wide
                                    const int x0 = (int)a16[k] + rm - c16[k], x1 = (int)a16[k + 1] + rm - c16[k + 1];
                                    const float2 s2 = make_float2(p.scale, p.scale);
                                    const float2 f = MAGIC ? __fmul2_rn(__fadd2_rn(make_float2(__int_as_float(x0), __int_as_float(x1)),
                                                                                   make_float2(-12582912.0f, -12582912.0f)), s2)
                                                           : __fmul2_rn(make_float2(__int2float_rn(x0), __int2float_rn(x1)), s2);
                                    y[j2 * 8 + k] = f.x;
                                    y[j2 * 8 + k + 1] = f.y;
                                    lmax = fmaxf(lmax, fmaxf(f.x, f.y));
                                }
                            } else {
#pragma unroll
                                for (int k = 0; k < 16; ++k) {
                                    if (j2 * 8 + k >= NSUB * 8) break;
                                    const int x = (int)a16[k] + rm - c16[k];
                                    float f = MAGIC ? __fmul_rn(__fadd_rn(__int_as_float(x), -12582912.0f), p.scale)
                                                    : __fmul_rn(__int2float_rn(x), p.scale);
                                    if (j2 * 8 + k >= ncols_w) f = kMasked;  // also covers the unloaded upper half
                                    y[j2 * 8 + k] = f;
                                    lmax = fmaxf(lmax, f);
                                }
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 16; ++k)
                                if (j2 * 8 + k < NSUB * 8) y[j2 * 8 + k] = kMasked;
                        }
                    }
                };
                if (warp_rows) {
                    if (p.fast22) pass1(std::true_type{});
                    else pass1(std::false_type{});
                }
                // the scores now live in registers: hand the accumulator buffer back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if constexpr (TWO) mbar_arrive_cluster(smem_u32(tempty_bar + acc), 0); else mbar_arrive(smem_u32(tempty_bar + acc)); }
                red[h * 128 + rloc] = lmax;
                named_bar_sync(1 + q, 128);
                const float gmax = fmaxf(fmaxf(red[rloc], red[128 + rloc]), fmaxf(red[256 + rloc], red[384 + rloc]));
                float lsum = 0.f;
                if (warp_rows) {
                    const float l2e = 1.44269502162933349609375f;
                    const float m2 = __fmul_rn(gmax, l2e);               // its rounding error scales every e_i alike
                    // four partial sums (columns k & 3) as two packed accumulators: fixed order -> deterministic
                    float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
                    const float2 l2e2 = make_float2(l2e, l2e), nm2 = make_float2(-m2, -m2);
#pragma unroll
                    for (int j = 0; j < NSUB; ++j) {
                        if (j * 8 < ncols_w) {
#pragma unroll
                            for (int k = 0; k < 8; k += 2) {
                                const float2 a = __ffma2_rn(make_float2(y[j * 8 + k], y[j * 8 + k + 1]), l2e2, nm2);
                                float2 e;
                                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(a.x));
                                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(a.y));
                                y[j * 8 + k] = e.x;
                                y[j * 8 + k + 1] = e.y;
                                if ((k & 3) == 0) s01 = __fadd2_rn(s01, e);
                                else s23 = __fadd2_rn(s23, e);
                            }
                        }
                    }
                    lsum = __fadd_rn(__fadd_rn(s01.x, s01.y), __fadd_rn(s23.x, s23.y));
                }
                red[512 + h * 128 + rloc] = lsum;
                named_bar_sync(1 + q, 128);
                // fixed combination order -> deterministic row sum
                const float gsum = __fadd_rn(__fadd_rn(red[512 + rloc], red[640 + rloc]),
                                             __fadd_rn(red[768 + rloc], red[896 + rloc]));
                int qsum = 0;
                if (row_ok && ncols_w > 0) {
                    const float kr = __frcp_rn(__fmul_rn(gsum, p.qargs.scale));   // 1 / (row sum * output scale)
                    int8_t* dst = reinterpret_cast<int8_t*>(p.C) + b * p.stride_c + m * p.ldc + col0;
                    // full groups of 8 columns carry no masking at all; the (at most one) ragged group is separate.
                    // A group that holds a valid column always lies inside the row (ldc = round_up(N, 16)), so every
                    // store is one aligned 8-byte store.
                    const int nfull = ncols_w >> 3, nrem = ncols_w & 7;
                    auto emit_group = [&](int j, auto ragged_tag) {
                        constexpr bool RAGGED = decltype(ragged_tag)::value;
                        int c[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            // probabilities lie in [0, 1]: with zp >= lo (host check) the quotient rounds straight
                            // out of the FMA; only the upper clamp remains, as a min in the magic-sum domain
                            c[k] = p.sm_noclamp ? __float_as_int(fminf(__fmaf_rn(y[j * 8 + k], kr, qz.magic), p.sm_top))
                                                : qz.template code_of_quotient<QS>(__fmul_rn(y[j * 8 + k], kr));
                            if (RAGGED) c[k] = (k < nrem) ? c[k] : 0;
                        }
                        const int w0 = pack4_codes(c[0], c[1], c[2], c[3]), w1 = pack4_codes(c[4], c[5], c[6], c[7]);
                        qsum = __dp4a(w0, 0x01010101, __dp4a(w1, 0x01010101, qsum));
                        *reinterpret_cast<int2*>(dst + j * 8) = make_int2(w0, w1);
                    };
#pragma unroll
                    for (int j = 0; j < NSUB; ++j) {
                        if (j < nfull) emit_group(j, std::false_type{});
                        else if (j == nfull && nrem > 0) emit_group(j, std::true_type{});
                    }
                }
                if (p.q_rowsum) {
                    // code sums of the 4 column groups meet in shared memory: plain store, no memset / atomics
                    redq[h * 128 + rloc] = qsum;
                    named_bar_sync(1 + q, 128);
                    if (h == 0 && row_ok)
                        p.q_rowsum[b * p.M + m] = (redq[rloc] + redq[128 + rloc]) + (redq[256 + rloc] + redq[384 + rloc]);
                }
                // red[] / redq[] reuse by the next tile is ordered by that tile's own barriers (each slot is
                // rewritten only after a barrier that every reader of the previous tile has passed)
                continue;
            }
#pragma unroll 1
            for (int sidx = hc; sidx < BN / 16; sidx += HPG) {
                const int64_t nc = n0 + sidx * 16;
                if (nc >= p.N) break;                                     // warp-uniform
                uint32_t v[16];
                tmem_ld_32x32b_x16(t_row + (uint32_t)(sidx * 16), v);
                const bool vec_chunk = c_aligned && nc + 16 <= p.N;
                int ct[4];
                float bs[4];
                if (FASTF) {
                    ct[0] = ct_n.x * zpa; ct[1] = ct_n.y * zpa; ct[2] = ct_n.z * zpa; ct[3] = ct_n.w * zpa;
                    bs[0] = __int_as_float(bs_n.x); bs[1] = __int_as_float(bs_n.y);
                    bs[2] = __int_as_float(bs_n.z); bs[3] = __int_as_float(bs_n.w);
                    fetch_cols(sidx + HPG);
                }
                tmem_ld_wait();
                if (EMODE == EM_REQUANT) {
                    // int8 codes: 16 bytes per row, written straight from the owning thread
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int64_t n = nc + j;
                        int qv = 0;
                        if (n < p.N) {
                            int64_t a = (int64_t)(int32_t)v[j];
                            if (p.bias_q) a += __ldg(p.bias_q + n);
                            const float d = dequantize_one(a - tile_zp(z, rowterm, b, n), p.scale);
                            qv = p.asym_out ? requantize_one<true>(d, p.inv_out_scale, p.out_zp, p.lo, p.hi)
                                            : requantize_one<false>(d, p.inv_out_scale, 0.0, p.lo, p.hi);
                        }
                        if ((j & 3) == 0) w[j >> 2] = 0;
                        w[j >> 2] |= (uint32_t)(qv & 0xff) << ((j & 3) * 8);
                    }
                    if (row_ok) {
                        int8_t* dst = reinterpret_cast<int8_t*>(p.C) + crow_base + m * p.ldc + nc;
                        if (nc + 16 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                            *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
                        } else {
                            for (int j = 0; j < 16 && nc + j < p.N; ++j) dst[j] = (int8_t)(w[j >> 2] >> ((j & 3) * 8));
                        }
                    }
                    continue;
                }
                if (EMODE == EM_DEQ_GENERAL) {
                    // general path (64-bit zero-point arithmetic) in registers, before the transpose
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int64_t n = nc + j;
                        float d = 0.f;
                        if (n < p.N) {
                            d = dequantize_one((int64_t)(int32_t)v[j] - tile_zp(z, rowterm, b, n), p.scale);
                            if (p.bias_f32) d = __fadd_rn(__ldg(p.bias_f32 + n), d);
                            if (p.residual && row_ok) d = __fadd_rn(d, __ldg(p.residual + b * p.stride_r + m * p.ldr + n));
                        }
                        v[j] = __float_as_uint(d);
                    }
                }
                // transpose through this warp's smem slab: rows of 64 B, 16-byte chunks XOR-swizzled by
                // (row >> 1) & 3 -> row-per-thread writes and 8-rows-per-instruction reads are conflict free
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(stg + lane * 16 + ((j ^ ((lane >> 1) & 3)) << 2)) =
                        make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
                uint32_t* cbase = reinterpret_cast<uint32_t*>(p.C) + crow_base + nc;
                if (vec_chunk) {
                    // each store instruction covers 8 rows x 64 B; this lane owns 4 fixed columns
                    uint32_t* crow = cbase + (mrow0 + rsub) * p.ldc + (cl << 2);
                    const int64_t cstep = 8 * p.ldc;
                    const bool res_on = (EMODE == EM_DEQ_FAST) && p.residual != nullptr;
                    const float* rrow = res_on ? p.residual + b * p.stride_r + (mrow0 + rsub) * p.ldr + nc + (cl << 2) : nullptr;
                    const int64_t rstep = 8 * p.ldr;
                    float4 res[4];
                    if (res_on) {
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const float* rp = rrow + it * rstep;
                            if (it * 8 + rsub >= rows_left) res[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                            else if (res_vec) res[it] = __ldcs(reinterpret_cast<const float4*>(rp));
                            else res[it] = make_float4(__ldg(rp), __ldg(rp + 1), __ldg(rp + 2), __ldg(rp + 3));
                        }
                    }
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int rr = it * 8 + rsub;
                        uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 16 + ((cl ^ ((rr >> 1) & 3)) << 2));
                        if (FASTF) {
                            // x = float bits of (1.5*2^23 + d), valid while |d| < 2^22: checked once per 4
                            const int rm = (int)stg_row[rr];
                            const int x0 = (int)val.x + rm - ct[0], x1 = (int)val.y + rm - ct[1];
                            const int x2 = (int)val.z + rm - ct[2], x3 = (int)val.w + rm - ct[3];
                            const uint32_t bad = ((uint32_t)(x0 ^ 0x4B000000) | (uint32_t)(x1 ^ 0x4B000000) |
                                                  (uint32_t)(x2 ^ 0x4B000000) | (uint32_t)(x3 ^ 0x4B000000)) & 0xFF800000u;
                            float f0, f1, f2, f3;
                            if (__builtin_expect(bad != 0, 0)) {
                                f0 = deq_slow(x0 - 0x4B400000, p.scale); f1 = deq_slow(x1 - 0x4B400000, p.scale);
                                f2 = deq_slow(x2 - 0x4B400000, p.scale); f3 = deq_slow(x3 - 0x4B400000, p.scale);
                            } else {
                                f0 = __fmul_rn(__fadd_rn(__int_as_float(x0), -12582912.0f), p.scale);
                                f1 = __fmul_rn(__fadd_rn(__int_as_float(x1), -12582912.0f), p.scale);
                                f2 = __fmul_rn(__fadd_rn(__int_as_float(x2), -12582912.0f), p.scale);
                                f3 = __fmul_rn(__fadd_rn(__int_as_float(x3), -12582912.0f), p.scale);
                            }
                            if (p.bias_f32) {
                                f0 = __fadd_rn(bs[0], f0); f1 = __fadd_rn(bs[1], f1);
                                f2 = __fadd_rn(bs[2], f2); f3 = __fadd_rn(bs[3], f3);
                            }
                            if (res_on) {
                                f0 = __fadd_rn(f0, res[it].x); f1 = __fadd_rn(f1, res[it].y);
                                f2 = __fadd_rn(f2, res[it].z); f3 = __fadd_rn(f3, res[it].w);
                            }
                            val = make_uint4(__float_as_uint(f0), __float_as_uint(f1), __float_as_uint(f2), __float_as_uint(f3));
                        }
                        if (rr < rows_left) *reinterpret_cast<uint4*>(crow) = val;
                        crow += cstep;
                    }
                } else {
                    // ragged / unaligned: one column per lane (16 columns x 2 rows per instruction)
                    const bool col_ok = nc + sc_col < p.N;
                    uint32_t* crow = cbase + (mrow0 + sc_row) * p.ldc + sc_col;
                    const float* rrow = p.residual ? p.residual + b * p.stride_r + (mrow0 + sc_row) * p.ldr + nc + sc_col : nullptr;
#pragma unroll 4
                    for (int it = 0; it < 16; ++it) {
                        const int rr = it * 2 + sc_row;
                        uint32_t val = stg[rr * 16 + ((((sc_col >> 2) ^ ((rr >> 1) & 3)) << 2) | (sc_col & 3))];
                        if (EMODE == EM_DEQ_FAST) {
                            const int x = (int)val + (int)stg_row[rr] - ct[0];
                            float f = (((uint32_t)(x ^ 0x4B000000) & 0xFF800000u) == 0)
                                          ? __fmul_rn(__fadd_rn(__int_as_float(x), -12582912.0f), p.scale)
                                          : deq_slow(x - 0x4B400000, p.scale);
                            if (p.bias_f32) f = __fadd_rn(bs[0], f);
                            if (rrow && col_ok && rr < rows_left) f = __fadd_rn(f, __ldg(rrow));
                            val = __float_as_uint(f);
                        }
                        if (col_ok && rr < rows_left) *crow = val;
                        crow += 2 * p.ldc;
                        if (rrow) rrow += 2 * p.ldr;
                    }
                }
            }
            // all TMEM reads of this buffer have completed (wait::ld above)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if constexpr (TWO) mbar_arrive_cluster(smem_u32(tempty_bar + acc), 0); else mbar_arrive(smem_u32(tempty_bar + acc)); }
        }
    }

    tc_fence_before();
    if constexpr (TWO) cluster_sync_all();                                // neither CTA leaves while the other still uses it
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if constexpr (TWO)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ CUDA-core cross-check GEMM
__global__ void __launch_bounds__(256) qgemm_simt_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B,
                                                        int32_t* __restrict__ Cm, int64_t M, int64_t N, int64_t K,
                                                        int64_t lda, int64_t ldb, int64_t ldc, int64_t sa, int64_t sb,
                                                        int64_t sc) {
    const int64_t n = (int64_t)blockIdx.x * 16 + (threadIdx.x & 15);
    const int64_t m = (int64_t)blockIdx.y * 16 + (threadIdx.x >> 4);
    const int64_t b = blockIdx.z;
    if (m >= M || n >= N) return;
    const int8_t* a = A + b * sa + m * lda;
    const int8_t* w = B + b * sb + n * ldb;
    int acc = 0;
    for (int64_t k = 0; k < K; ++k) acc += (int)a[k] * (int)w[k];
    Cm[b * sc + m * ldc + n] = acc;
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 3-D map over a K-major int8 operand [batch][rows][K] (row stride ld bytes).
int make_operand_map(CUtensorMap* map, const int8_t* base, int64_t K, int64_t rows, int64_t batch, int64_t ld,
                            int64_t batch_stride, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    NQ_REQUIRE(enc, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)(batch > 1 ? batch_stride : rows * ld)};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NQ_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (CUresult %d; K=%lld rows=%lld batch=%lld ld=%lld)", (int)r,
               (long long)K, (long long)rows, (long long)batch, (long long)ld);
    return NQ_OK;
}

// Implicit-GEMM convolution operands.  A: im2col-mode map over the padded NHWC image [n][Hp][Wp][C]: 128 output
// pixels x 64 channels per load; the bounding box keeps every filter tap of every base pixel inside the (already
// padded) image, so nothing is zero-filled except rows past the last image.  B: filter matrix [O][KH*KW*C] in
// 64-byte K slices.  Both 64B-swizzled (make_smem_desc_sw64).
struct ConvGeom {
    int64_t n_img, Hp, Wp, C, KH, KW, sh, sw, OH, OW;
};

using EncodeIm2colFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_conv_maps(CUtensorMap* ta, CUtensorMap* tb, const int8_t* X, const int8_t* Wm, const ConvGeom& g, int64_t O,
                          int64_t ldw, int bn) {
    static EncodeIm2colFn enc_i2c = nullptr;
    if (!enc_i2c) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            enc_i2c = reinterpret_cast<EncodeIm2colFn>(ptr);
    }
    EncodeTiledFn enc = get_encode_fn();
    NQ_REQUIRE(enc && enc_i2c, "cuTensorMapEncodeTiled / cuTensorMapEncodeIm2col not available from the driver");
    {
        cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.Wp, (cuuint64_t)g.Hp, (cuuint64_t)g.n_img};
        cuuint64_t strides[3] = {(cuuint64_t)g.C, (cuuint64_t)(g.Wp * g.C), (cuuint64_t)(g.Hp * g.Wp * g.C)};
        int lower[2] = {0, 0};                                            // {W, H}: base pixels start at the image corner
        int upper[2] = {-(int)(g.KW - 1), -(int)(g.KH - 1)};              // ... and stop where the last tap still fits
        cuuint32_t estr[4] = {1, (cuuint32_t)g.sw, (cuuint32_t)g.sh, 1};
        CUresult r = enc_i2c(ta, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<int8_t*>(X), dims, strides, lower, upper, 64, BM, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        NQ_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col failed (CUresult %d; C=%lld Wp=%lld Hp=%lld n=%lld)", (int)r,
                   (long long)g.C, (long long)g.Wp, (long long)g.Hp, (long long)g.n_img);
    }
    {
        const int64_t K = g.KH * g.KW * g.C;
        cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)O, 1};
        cuuint64_t strides[2] = {(cuuint64_t)ldw, (cuuint64_t)(O * ldw)};
        cuuint32_t box[3] = {64, (cuuint32_t)bn, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(tb, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(Wm), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        NQ_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (filter matrix) failed (CUresult %d)", (int)r);
    }
    return NQ_OK;
}

template <int BN, int EMODE, bool TWO>
static int launch_qgemm_impl(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
    using C = Cfg<BN, (EMODE == EM_DEQ_WIDE || EMODE == EM_DEQ_WIDE_RES), TWO>;
    static bool configured[64] = {false};                                 // per device (cudaFuncSetAttribute is per device)
    if (int rc = configure_smem_once(configured, qgemm_kernel<BN, EMODE, TWO>, C::SMEM_BYTES, "cudaFuncSetAttribute(qgemm)")) return rc;
    constexpr int BMT = TWO ? 2 * BM : BM;
    const int64_t tiles = ((p.M + BMT - 1) / BMT) * ((p.N + BN - 1) / BN) * p.batch;
    if constexpr (TWO) {
        // one cluster of 2 CTAs (a TPC's SM pair) per 256 x BN tile slot
        const int64_t pairs_max = sm_count() / 2;
        const int pairs = (int)(tiles < pairs_max ? tiles : pairs_max);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * pairs, 1, 1);
        cfg.blockDim = dim3(NUM_THREADS, 1, 1);
        cfg.dynamicSmemBytes = C::SMEM_BYTES;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, qgemm_kernel<BN, EMODE, true>, ta, tb, p);
        if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(qgemm, cluster 2)");
    } else {
        const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
        qgemm_kernel<BN, EMODE, false><<<grid, NUM_THREADS, C::SMEM_BYTES, s>>>(ta, tb, p);
    }
    NQ_CHECK_LAUNCH("nq_qgemm_s8");
    return NQ_OK;
}

template <int BN, int EMODE>
static int launch_qgemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
    if constexpr (BN == 256 && EMODE != EM_SOFTMAX_SYM && EMODE != EM_SOFTMAX_ASYM) {
        if (p.two_cta) return launch_qgemm_impl<BN, EMODE, true>(ta, tb, p, s);
    }
    return launch_qgemm_impl<BN, EMODE, false>(ta, tb, p, s);
}

template <int BN>
static int launch_qgemm_mode(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
    if (p.mode == NQ_EPI_RAW) return launch_qgemm<BN, EM_RAW>(ta, tb, p, s);
    if (p.mode == NQ_EPI_REQUANT) return p.req_rows ? launch_qgemm<BN, EM_REQ_ROWS>(ta, tb, p, s) : launch_qgemm<BN, EM_REQUANT>(ta, tb, p, s);
    if (p.mode == NQ_EPI_SOFTMAX_QUANT) {
        if constexpr (BN == 256)
            return p.asym_out ? launch_qgemm<BN, EM_SOFTMAX_ASYM>(ta, tb, p, s) : launch_qgemm<BN, EM_SOFTMAX_SYM>(ta, tb, p, s);
        else {
            set_error("nq_qgemm_s8: SOFTMAX epilogue is built for the 256-column tile only");
            return NQ_ERR_UNSUPPORTED;
        }
    }
    if (p.mode == NQ_EPI_QUANT)
        return p.q_off[5] == 1 ? launch_qgemm<BN, EM_Q8_ROWS>(ta, tb, p, s) : launch_qgemm<BN, EM_Q8_COLS>(ta, tb, p, s);
    if (p.mode == NQ_EPI_GELU_QUANT) return launch_qgemm<BN, EM_Q8_GELU>(ta, tb, p, s);
    if (p.fast32 && p.deq_wide)
        return p.residual ? launch_qgemm<BN, EM_DEQ_WIDE_RES>(ta, tb, p, s) : launch_qgemm<BN, EM_DEQ_WIDE>(ta, tb, p, s);
    return p.fast32 ? launch_qgemm<BN, EM_DEQ_FAST>(ta, tb, p, s) : launch_qgemm<BN, EM_DEQ_GENERAL>(ta, tb, p, s);
}

}  // namespace nq

using namespace nq;

// nq_qgemm_s8 and nq_qconv2d_s8 (cv != nullptr: A is the padded NHWC image, lda = K is nominal)
static int qgemm_run(const int8_t* A, const int8_t* B, void* Cout, int64_t M, int64_t N, int64_t K, int64_t batch, int64_t lda,
                     int64_t ldb, int64_t ldc, int64_t stride_a, int64_t stride_b, int64_t stride_c, const nq_epilogue* ep,
                     void* stream, const ConvGeom* cv) {
    NQ_REQUIRE(ep, "nq_qgemm_s8: epilogue descriptor is NULL");
    NQ_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0, "nq_qgemm_s8: empty problem M=%lld N=%lld K=%lld batch=%lld",
               (long long)M, (long long)N, (long long)K, (long long)batch);
    NQ_REQUIRE((lda % 16 == 0) && (ldb % 16 == 0) && lda >= K && ldb >= K,
               "nq_qgemm_s8: lda/ldb must be >= K and multiples of 16 bytes (lda=%lld ldb=%lld K=%lld)", (long long)lda,
               (long long)ldb, (long long)K);
    NQ_REQUIRE((stride_a % 16 == 0) && (stride_b % 16 == 0), "nq_qgemm_s8: batch strides must be multiples of 16 bytes");
    NQ_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0), "nq_qgemm_s8: operands must be 16-byte aligned");
    NQ_REQUIRE(ldc >= N, "nq_qgemm_s8: ldc < N");
    NQ_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31) && batch < (1ll << 31), "nq_qgemm_s8: extent too large");
    NQ_REQUIRE(((M + 127) / 128) * ((N + 63) / 64) * batch < (1ll << 31), "nq_qgemm_s8: too many output tiles");
    NQ_REQUIRE(ep->mode >= NQ_EPI_RAW && ep->mode <= NQ_EPI_GELU_QUANT, "nq_qgemm_s8: unknown epilogue mode %d", ep->mode);
    if (ep->mode != NQ_EPI_RAW)
        if (int rc = check_acc_zp(&ep->zp)) return rc;

    GemmParams p{};
    p.M = M; p.N = N; p.K = K; p.batch = batch;
    p.ldc = ldc; p.stride_c = stride_c;
    p.a_batched = (batch > 1 && stride_a != 0);
    p.b_batched = (batch > 1 && stride_b != 0);
    p.mode = ep->mode;
    p.scale = ep->scale;
    p.zp = (ep->mode == NQ_EPI_RAW) ? AccZp{} : make_acc_zp(&ep->zp);
    {
        // |acc| + |rowsum*zp_b - zp_a*zp_b*K| + |colsum*zp_a| with 8-bit operands (|q| <= 128)
        const long double za = (long double)llabs(p.zp.zp_a), zb = (long double)llabs(p.zp.zp_b);
        const long double bound = 16384.0L * K + 128.0L * K * (za + zb) + za * zb * K;
        p.fast32 = (ep->mode == NQ_EPI_DEQUANT || ep->mode == NQ_EPI_QUANT || ep->mode == NQ_EPI_SOFTMAX_QUANT ||
                    ep->mode == NQ_EPI_GELU_QUANT) &&
                   bound < 2147483000.0L;
        static const bool no24 = getenv("NQ_NO_I2F_EPILOGUE") != nullptr;         // A/B switch
        p.fast24 = !no24 && p.fast32 && bound < 16777216.0L;
    }
    if (ep->mode == NQ_EPI_SOFTMAX_QUANT) {
        NQ_REQUIRE(p.fast32, "nq_qgemm_s8: SOFTMAX epilogue needs the 32-bit zero-point bound");
        NQ_REQUIRE(N <= 224, "nq_qgemm_s8: SOFTMAX epilogue handles rows of at most 224 columns (N=%lld)", (long long)N);
        NQ_REQUIRE(ldc % 16 == 0 && ldc >= N, "nq_qgemm_s8: SOFTMAX epilogue writes an int8 operand: ldc %% 16 == 0 required");
        NQ_REQUIRE(ep->out_bits >= 2 && ep->out_bits <= 8, "nq_qgemm_s8: out_bits %d outside 2..8", ep->out_bits);
        int qmode;
        p.qargs = make_qargs(ep->out_bits, ep->out_scale, ep->has_out_zp, ep->out_zp, &qmode);
        NQ_REQUIRE(qmode != 2, "nq_qgemm_s8: SOFTMAX epilogue needs |out_zp| < 2^20");
        p.asym_out = ep->has_out_zp;
        p.q_rowsum = ep->q_rowsum;
        // the Div constant folds into the dequantization scale: exact for a power of two, otherwise one
        // common relative error of 2^-24 on every score of the row (float glue, 1e-5 contract)
        if (ep->sm_has_div) {
            NQ_REQUIRE(ep->sm_div > 0.f && isfinite(ep->sm_div), "nq_qgemm_s8: SOFTMAX divisor must be positive and finite");
            p.scale = ep->scale / ep->sm_div;
        }
        NQ_REQUIRE(p.scale > 1e-30f && p.scale < 1e30f, "nq_qgemm_s8: SOFTMAX epilogue: scale/divisor out of range");
        {
            // exact range of sum_k (a - zp_a)(b - zp_b) over 8-bit codes
            const long double za = (long double)p.zp.zp_a, zb = (long double)p.zp.zp_b;
            const long double ra = fmaxl(fabsl(-128.0L - za), fabsl(127.0L - za));
            const long double rb = fmaxl(fabsl(-128.0L - zb), fabsl(127.0L - zb));
            p.fast22 = (ra * rb * (long double)K) < 4194304.0L;
        }
        {
            // p in [0, 1] (up to a few ulp): p / s_out + zp in [zp, zp + 1 / s_out].  zp >= lo: no lower clamp, the
            // quotient rounds straight out of one FMA with the magic constant; the upper clamp is a min in that
            // domain and vanishes (huge bound) when [.., zp + 1 / s_out], widened by the rounding slack, stays below hi + 0.5
            const double top = (double)p.qargs.zpf + 1.0 / (double)p.qargs.scale * (1.0 + 1e-6);
            p.sm_noclamp = (double)p.qargs.zpf >= (double)p.qargs.lo;
            p.sm_top = (top < (double)p.qargs.hi + 0.49) ? 3.0e38f : 12582912.0f + p.qargs.hi;
        }
    }
    const bool q8 = ep->mode == NQ_EPI_QUANT || ep->mode == NQ_EPI_GELU_QUANT;
    int bn = (ep->mode == NQ_EPI_SOFTMAX_QUANT) ? 256 : (N <= 64) ? 64 : (N <= 128) ? 128 : 256;
    {
        // A/B switch for measurements: narrower tiles for wide N (more A re-reads, two / four tile groups in flight)
        static const int force_bn = getenv("NQ_FORCE_BN") ? atoi(getenv("NQ_FORCE_BN")) : 0;
        if ((force_bn == 64 || force_bn == 128) && ep->mode != NQ_EPI_SOFTMAX_QUANT && force_bn < bn) bn = force_bn;
    }
    if (q8) {
        NQ_REQUIRE(p.fast32, "nq_qgemm_s8: QUANT epilogue needs the 32-bit zero-point bound (K or zero-points too large)");
        NQ_REQUIRE(N % 16 == 0 && ep->q_cols_per_head > 0 && ep->q_cols_per_head % 16 == 0 && ep->q_rows_per_image > 0,
                   "nq_qgemm_s8: QUANT epilogue needs N and cols_per_head multiples of 16");
        NQ_REQUIRE(ep->out_bits >= 2 && ep->out_bits <= 8, "nq_qgemm_s8: out_bits %d outside 2..8", ep->out_bits);
        int qmode;
        p.qargs = make_qargs(ep->out_bits, ep->out_scale, ep->has_out_zp, ep->out_zp, &qmode);
        NQ_REQUIRE(qmode != 2, "nq_qgemm_s8: QUANT epilogue needs |out_zp| < 2^20");
        p.asym_out = ep->has_out_zp;
        p.q_S = (uint32_t)ep->q_rows_per_image;
        p.q_D = (uint32_t)ep->q_cols_per_head;
        for (int i = 0; i < 6; ++i) {
            p.q_off[i] = ep->q_off[i];
            p.q_rs[i] = ep->q_rs[i];
        }
        p.q_rowsum = ep->q_rowsum;
        if (ep->q_off[5] == 1) {
            // ROWS: 16 consecutive n are one aligned 16-byte store
            for (int i = 0; i < 5; ++i)
                NQ_REQUIRE(ep->q_off[i] % 16 == 0, "nq_qgemm_s8: QUANT row layout needs q_off[0..4] multiples of 16 bytes");
            NQ_REQUIRE(((uintptr_t)Cout & 15) == 0, "nq_qgemm_s8: QUANT destination must be 16-byte aligned");
            NQ_REQUIRE(ep->q_off[4] >= 0 && (long double)(N / ep->q_cols_per_head + 1) * (long double)ep->q_off[4] < 2147483648.0L,
                       "nq_qgemm_s8: QUANT row layout: the head part of a destination offset must fit 31 bits");
            NQ_REQUIRE(!ep->q_rowsum || ep->q_rs[5] == 0, "nq_qgemm_s8: QUANT row layout sums codes along n (q_rs[5] == 0)");
            // one warp owns 64 consecutive columns of a row: a (row, head) slot has a single writer when
            // the heads tile that span and the slot does not depend on the batch-inner / other tiles
            const int wcols = 64;
            p.q_rs_exclusive = (wcols % ep->q_cols_per_head == 0) && ep->q_rs[4] != 0 && ep->q_rs[3] != 0 &&
                               (ep->c_batch_inner <= 1 || ep->q_rs[1] != 0);
        } else {
            NQ_REQUIRE(ep->mode == NQ_EPI_QUANT && ep->q_off[3] == 1,
                       "nq_qgemm_s8: QUANT column layout needs q_off[3] == 1 (consecutive m are consecutive bytes)");
            NQ_REQUIRE(!ep->q_rowsum || ep->q_rs[3] == 0, "nq_qgemm_s8: QUANT column layout sums codes along m (q_rs[3] == 0)");
        }
        if (ep->q_rowsum && !p.q_rs_exclusive) {
            // partial sums meet through atomics: the library clears the slots itself (stream-ordered)
            NQ_REQUIRE(ep->q_rowsum_count > 0, "nq_qgemm_s8: q_rowsum_count (number of int32 slots) is required with q_rowsum");
            cudaError_t e = cudaMemsetAsync(ep->q_rowsum, 0, sizeof(int32_t) * (size_t)ep->q_rowsum_count, (cudaStream_t)stream);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(q_rowsum)");
        }
        if (ep->mode == NQ_EPI_GELU_QUANT) {
            NQ_REQUIRE(ep->gelu_div != 0.f && isfinite(ep->gelu_div), "nq_qgemm_s8: GELU divisor must be finite and non-zero");
            NQ_REQUIRE(ep->gelu_div > 0.f, "nq_qgemm_s8: GELU epilogue needs a positive divisor (sign(x / c1) = sign(x))");
            // (erf(x / c1) + 1) * x * c3 >= -0.2403 * c1 * c3: with c2 = 1 and c3 > 0 the quotient by the output scale is
            // bounded below and the epilogue drops the sign copy and the lower clamp; anything else is not a GELU
            NQ_REQUIRE(ep->gelu_add == 1.0f && ep->gelu_mul > 0.f && isfinite(ep->gelu_mul),
                       "nq_qgemm_s8: GELU epilogue needs Add constant 1 and a positive Mul constant");
            NQ_REQUIRE(0.25 * (double)ep->gelu_div * (double)ep->gelu_mul / (double)ep->out_scale < 4.0e6,
                       "nq_qgemm_s8: GELU epilogue: c1 * c3 / out_scale too large for the single-add rounding");
            p.g_prdiv = (float)(0.3275911 / (double)ep->gelu_div);
            p.g_nl2e = (float)(-1.4426950408889634 / ((double)ep->gelu_div * (double)ep->gelu_div));
            p.g_add = ep->gelu_add;
            p.g_out = (float)((double)ep->gelu_mul / (double)ep->out_scale);
        }
    }
    p.bias_f32 = ep->bias_f32;
    p.bias_q = ep->bias_q;
    p.dbg = getenv("NQ_GEMM_DBG") ? atoi(getenv("NQ_GEMM_DBG")) : 0;
    p.reverse = ep->reverse_tiles != 0;
    p.trace = getenv("NQ_GEMM_TRACE") ? (long long*)strtoull(getenv("NQ_GEMM_TRACE"), nullptr, 10) : nullptr;
    p.c_inner = ep->c_batch_inner > 1 ? ep->c_batch_inner : 1;
    p.stride_c_inner = ep->c_batch_inner > 1 ? ep->stride_c_inner : 0;
    NQ_REQUIRE(p.c_inner == 1 || (batch % p.c_inner == 0 && ep->mode != NQ_EPI_REQUANT),
               "nq_qgemm_s8: c_batch_inner must divide batch (and is not available with REQUANT)");
    p.residual = (ep->mode == NQ_EPI_DEQUANT) ? ep->residual : nullptr;
    p.ldr = ep->ld_residual;
    p.stride_r = ep->stride_residual;
    NQ_REQUIRE(!p.residual || p.ldr >= N, "nq_qgemm_s8: ld_residual < N");
    p.C = Cout;
    if (ep->mode == NQ_EPI_DEQUANT) {
        auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
        static const bool no_wide = getenv("NQ_NO_WIDE_EPILOGUE") != nullptr;   // A/B switch for measurements
        p.deq_wide = !no_wide && N % 32 == 0 && ldc % 4 == 0 && stride_c % 4 == 0 && p.stride_c_inner % 4 == 0 && al16(Cout) &&
                     (!p.zp.use_col || (al16(p.zp.colsum_b) && p.zp.cs_stride % 4 == 0)) && (!p.bias_f32 || al16(p.bias_f32)) &&
                     (!p.residual || (al16(p.residual) && p.ldr % 4 == 0 && p.stride_r % 4 == 0));
    }
    if (ep->mode == NQ_EPI_REQUANT) {
        NQ_REQUIRE(ep->out_bits >= 2 && ep->out_bits <= 8, "nq_qgemm_s8: out_bits %d outside 2..8", ep->out_bits);
        p.inv_out_scale = 1.0f / ep->out_scale;
        p.asym_out = ep->has_out_zp;
        p.out_zp = ep->has_out_zp ? (double)ep->out_zp : 0.0;
        p.lo = -ldexpf(1.f, ep->out_bits - 1);
        p.hi = ldexpf(1.f, ep->out_bits - 1) - 1.f;
        // Thread-per-row epilogue (the QUANT epilogues' structure): acc + int bias - zero-point terms in int32 (each of
        // the three below 2^29: host bound here, bias checked on the device per tile), magic-constant conversion, the
        // reciprocal multiply and the single-add rounding of zp + t.  Needs 16-byte rows; everything else -- and any tile
        // whose values leave the windows -- takes the general 64-bit route.
        int qmode;
        p.qargs = make_qargs(ep->out_bits, ep->out_scale, ep->has_out_zp, ep->out_zp, &qmode);
        const long double za = (long double)llabs(p.zp.zp_a), zb = (long double)llabs(p.zp.zp_b);
        const long double bound = 16384.0L * K + 128.0L * K * (za + zb) + za * zb * K;
        static const bool no_fast = getenv("NQ_NO_FAST_REQUANT") != nullptr;      // A/B switch
        // ragged N: the last 16-column step stores into the row's padding, which the caller provides (ldc >= round_up(N, 16))
        p.req_rows = !no_fast && qmode != 2 && bound < 536870000.0L && ldc >= (N + 15) / 16 * 16 && ldc % 16 == 0 && stride_c % 16 == 0 &&
                     ((uintptr_t)Cout % 16 == 0) && p.c_inner == 1 && M < (1ll << 31);
        if (p.req_rows) {
            p.q_S = (uint32_t)M;                                          // one "image" of M rows, one "head" of N columns
            p.q_D = (uint32_t)N;
            const int64_t off[6] = {stride_c, 0, 0, ldc, 0, 1};
            for (int i = 0; i < 6; ++i) {
                p.q_off[i] = off[i];
                p.q_rs[i] = 0;
            }
            p.q_rowsum = nullptr;
        }
    }

    {
        static const bool no_pair = getenv("NQ_NO_2CTA") != nullptr;             // A/B switch for measurements
        // The pair halves the per-SM operand traffic of the main loop; measured (microbench A/B): +10..20 % where the
        // main loop dominates (K >= 1024: 4096^3, 8192^3, the K = 3072 MLP GEMM), -5..10 % for K = 768 tiles whose time
        // is the epilogue (the leader's next MMA has to wait for both CTAs' epilogues).
        static const int pair_min_k = getenv("NQ_2CTA_MINK") ? atoi(getenv("NQ_2CTA_MINK")) : 1024;   // A/B switch
        p.two_cta = !no_pair && !cv && bn == 256 && ep->mode != NQ_EPI_SOFTMAX_QUANT && M >= 256 && K >= pair_min_k;
    }
    CUtensorMap ta, tb;
    if (cv) {
        p.conv = 1;
        p.cv_C = (uint32_t)cv->C; p.cv_KW = (uint32_t)cv->KW;
        p.cv_OW = (uint32_t)cv->OW; p.cv_OH = (uint32_t)cv->OH;
        p.cv_sw = (uint32_t)cv->sw; p.cv_sh = (uint32_t)cv->sh;
        if (int rc = make_conv_maps(&ta, &tb, A, B, *cv, N, ldb, bn)) return rc;
    } else {
        if (int rc = make_operand_map(&ta, A, K, M, p.a_batched ? batch : 1, lda, stride_a, BM)) return rc;
        if (int rc = make_operand_map(&tb, B, K, N, p.b_batched ? batch : 1, ldb, stride_b, p.two_cta ? bn / 2 : bn)) return rc;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (bn == 64) return launch_qgemm_mode<64>(ta, tb, p, s);
    if (bn == 128) return launch_qgemm_mode<128>(ta, tb, p, s);
    return launch_qgemm_mode<256>(ta, tb, p, s);
}

extern "C" int nq_qgemm_s8(const int8_t* A, const int8_t* B, void* Cout, int64_t M, int64_t N, int64_t K,
                           int64_t batch, int64_t lda, int64_t ldb, int64_t ldc, int64_t stride_a, int64_t stride_b,
                           int64_t stride_c, const nq_epilogue* ep, void* stream) {
    return qgemm_run(A, B, Cout, M, N, K, batch, lda, ldb, ldc, stride_a, stride_b, stride_c, ep, stream, nullptr);
}

extern "C" int nq_qconv2d_s8(const int8_t* X, const int8_t* Wm, void* Cout, int64_t n_img, int64_t Hp, int64_t Wp, int64_t Cin,
                             int64_t KH, int64_t KW, int64_t stride_h, int64_t stride_w, int64_t O, int64_t ldw, int64_t ldc,
                             const nq_epilogue* ep, void* stream) {
    NQ_REQUIRE(n_img > 0 && Hp > 0 && Wp > 0 && Cin > 0 && KH > 0 && KW > 0 && O > 0, "nq_qconv2d_s8: empty problem");
    NQ_REQUIRE(Cin % 64 == 0, "nq_qconv2d_s8: channels must be a multiple of 64 (K slices of one filter tap; C=%lld)", (long long)Cin);
    NQ_REQUIRE(KH <= Hp && KW <= Wp && KH <= 128 && KW <= 128, "nq_qconv2d_s8: filter %lldx%lld does not fit the padded image %lldx%lld",
               (long long)KH, (long long)KW, (long long)Hp, (long long)Wp);
    NQ_REQUIRE(stride_h >= 1 && stride_h <= 8 && stride_w >= 1 && stride_w <= 8, "nq_qconv2d_s8: strides must be 1..8 (TMA traversal stride)");
    NQ_REQUIRE(ep && (ep->mode == NQ_EPI_RAW || ep->mode == NQ_EPI_DEQUANT || ep->mode == NQ_EPI_REQUANT),
               "nq_qconv2d_s8: epilogue must be RAW, DEQUANT or REQUANT");
    NQ_REQUIRE(ep->mode == NQ_EPI_RAW || !ep->zp.has_zp_b, "nq_qconv2d_s8: filters must be symmetric (no row sums of the patch matrix exist)");
    ConvGeom g{n_img, Hp, Wp, Cin, KH, KW, stride_h, stride_w, (Hp - KH) / stride_h + 1, (Wp - KW) / stride_w + 1};
    const int64_t M = n_img * g.OH * g.OW, K = KH * KW * Cin;
    NQ_REQUIRE(M < (1ll << 31) && Wp * stride_w < (1ll << 31), "nq_qconv2d_s8: extent too large");
    return qgemm_run(X, Wm, Cout, M, O, K, 1, K, ldw, ldc, 0, 0, 0, ep, stream, &g);
}

extern "C" int nq_qgemm_s8_simt(const int8_t* A, const int8_t* B, int32_t* Cm, int64_t M, int64_t N, int64_t K,
                                int64_t batch, int64_t lda, int64_t ldb, int64_t ldc, int64_t stride_a,
                                int64_t stride_b, int64_t stride_c, void* stream) {
    NQ_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0 && batch <= 65535, "nq_qgemm_s8_simt: bad extents");
    dim3 grid((unsigned)((N + 15) / 16), (unsigned)((M + 15) / 16), (unsigned)batch);
    NQ_REQUIRE(grid.y <= 65535, "nq_qgemm_s8_simt: M too large for the cross-check kernel");
    qgemm_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, B, Cm, M, N, K, lda, ldb, ldc, stride_a, stride_b, stride_c);
    NQ_CHECK_LAUNCH("nq_qgemm_s8_simt");
    return NQ_OK;
}
