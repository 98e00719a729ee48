// K4/K5: int8 x int8 -> int32 GEMM on the 5th-gen tensor cores (tcgen05.mma kind::i8),
// replacing the reference's `np.matmul` on int64 arrays (numpy_quantization.py:44-61).
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor 3-D, 128B swizzle, K-major tiles)
//   warp 1      MMA issuer     (one elected thread, tcgen05.mma cta_group::1, M=128, N=BN, K=32)
//   warp 2      TMEM allocator (2 accumulator buffers of BN int32 columns)
//   warps 4-11  epilogue       (tcgen05.ld 32x32b -> per-warp swizzled smem transpose ->
//                               zero-point correction + dequant(+bias) on the row-contiguous
//                               read-back -> 16-byte coalesced stores; requant -> int8 codes)
// smem ring of STAGES x (A 128x128 B + B BNx128 B); mbarrier full/empty per stage and
// tmem_full/tmem_empty per accumulator buffer, so the epilogue of tile i overlaps the
// main loop of tile i+1.
#include <cuda.h>

#include "common.cuh"

namespace nq {

constexpr int BM = 128;          // rows of A per tile == TMEM lanes
constexpr int BK = 128;          // int8 elements per stage along K == one 128-byte swizzle row
constexpr int UMMA_K = 32;       // K per tcgen05.mma for 8-bit operands
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + 32 * NUM_EPI_WARPS;

template <int BN> struct Cfg {
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int A_BYTES = BM * BK;
    static constexpr int B_BYTES = BN * BK;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int EPI_BYTES = NUM_EPI_WARPS * (32 * 32 * 4 + 32 * 4);   // staging slabs + row terms
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
};

struct GemmParams {
    int64_t M, N, K, batch;
    int64_t ldc, stride_c;
    int64_t c_inner, stride_c_inner; // C batch offset = (b / c_inner) * stride_c + (b % c_inner) * stride_c_inner
    int a_batched, b_batched;        // 0 -> operand shared across the batch (TMA batch coord 0)
    int mode;
    int fast32;                      // zero-point arithmetic provably fits int32 (host-checked bound)
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
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug traps (launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Read-only global loads as volatile asm: the compiler keeps them where they are written (ahead of the
// accumulator wait), so their latency overlaps the main loop instead of being sunk to the first use.
__device__ __forceinline__ int4 ldg_v4(const void* p) {
    int4 r;
    asm volatile("ld.global.nc.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ int ldg_s32(const void* p) {
    int r;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart
// (SBO), LBO unused for swizzled K-major; descriptor version 1 (Blackwell), layout 2.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address      bits [0,14)
    d |= (uint64_t)1 << 16;                         // LBO (16 B units)   bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;               // SBO = 1024 B       bits [32,46)
    d |= (uint64_t)1 << 46;                         // version            bits [46,48)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B       bits [61,64)
    return d;
}

// kind::i8 instruction descriptor: D=s32, A=B=signed int8, both K-major, N>>3, M>>4.
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ------------------------------------------------------------------ epilogue math
// dequantize with 32-bit zero-point arithmetic: identical bits to f32(f64(d) * f64(scale)).
// |d| < 2^22 converts through the 1.5*2^23 magic constant (integer add + FADD, full rate)
// instead of I2F; larger values take the conversion / float64 routes.
__device__ __noinline__ float deq_slow(int d, float scale) {       // rare: |d| >= 2^22, kept out of line
    if (d >= -16777216 && d <= 16777216) return __fmul_rn((float)d, scale);
    return (float)((double)d * (double)scale);
}
__device__ __forceinline__ float deq_fast(int a, int rt, int ct, float scale) {
    const int d = a - rt - ct;
    if (__builtin_expect((unsigned)(d + 0x400000) >= 0x800000u, 0)) return deq_slow(d, scale);
    return __fmul_rn(__fadd_rn(__int_as_float(0x4B400000 + d), -12582912.0f), scale);
}

__device__ __forceinline__ int64_t tile_zp(const AccZp& z, int64_t rowterm, int64_t b, int64_t n) {
    int64_t v = rowterm;
    if (z.use_col) v += (int64_t)__ldg(z.colsum_b + b * z.cs_stride + n) * z.zp_a;
    return v;
}

constexpr int EM_RAW = 0, EM_DEQ_FAST = 1, EM_DEQ_GENERAL = 2, EM_REQUANT = 3;

template <int BN, int EMODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
qgemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    // 128B swizzle atoms need 1024-byte aligned tiles
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps .shared provenance
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + C::STAGES * C::A_BYTES;
    uint32_t* epi = reinterpret_cast<uint32_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::EPI_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + C::STAGES;
    uint64_t* tfull_bar = bars + 2 * C::STAGES;
    uint64_t* tempty_bar = bars + 2 * C::STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile bookkeeping in 32 bits (host guarantees total_tiles < 2^31): 64-bit divides are long
    // emulated sequences that would otherwise sit in every role's tile loop
    const uint32_t m_tiles = (uint32_t)((p.M + BM - 1) / BM), n_tiles = (uint32_t)((p.N + BN - 1) / BN);
    const uint32_t tiles_per_batch = m_tiles * n_tiles, total_tiles = tiles_per_batch * (uint32_t)p.batch;
    const uint32_t k_blocks = (uint32_t)((p.K + BK - 1) / BK);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < C::STAGES; ++i) {
            mbar_init(smem_u32(full_bar + i), 1);
            mbar_init(smem_u32(empty_bar + i), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(tfull_bar + i), 1);
            mbar_init(smem_u32(tempty_bar + i), NUM_EPI_WARPS);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const uint32_t b = t / tiles_per_batch, r = t % tiles_per_batch;
                const int m0 = (int)(r / n_tiles) * BM, n0 = (int)(r % n_tiles) * BN;
                const int ba = p.a_batched ? (int)b : 0, bb = p.b_batched ? (int)b : 0;
                for (uint32_t kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
                    const uint32_t fb = smem_u32(full_bar + stage);
                    mbar_expect_tx(fb, C::STAGE_BYTES);
                    tma_load_3d(smem_u32(smem_a + stage * C::A_BYTES), &tmap_a, (int)(kb * BK), m0, ba, fb);
                    tma_load_3d(smem_u32(smem_b + stage * C::B_BYTES), &tmap_b, (int)(kb * BK), n0, bb, fb);
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                mbar_wait(smem_u32(tempty_bar + acc), acc_phase ^ 1);     // epilogue drained this buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (uint32_t kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(smem_u32(full_bar + stage), phase);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc(smem_u32(smem_a + stage * C::A_BYTES));
                    const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + stage * C::B_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 32 bytes along K inside the swizzle atom: +2 in 16-byte units
                        mma_i8(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                               (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(smem_u32(empty_bar + stage));               // smem slot free once MMAs retire
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit(smem_u32(tfull_bar + acc));                     // accumulator ready
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue (8 warps) =====================
        // warp -> TMEM lane quarter q (hardware rule: warp_id % 4) and column-chunk parity h: two
        // warps share a quarter and take alternate 32-column chunks.  EMODE is a template parameter so
        // that each instantiation carries only its own epilogue (the loop body stays I-cache resident).
        const int q = warp & 3;
        const int h = (warp - 4) >> 2;
        const int ew = warp - 4;
        uint32_t* stg = epi + ew * (32 * 32);                             // 32 rows x 32 words, XOR-swizzled
        uint32_t* stg_row = epi + NUM_EPI_WARPS * 32 * 32 + ew * 32;
        int acc = 0;
        uint32_t acc_phase = 0;
        const AccZp z = p.zp;
        const bool c_aligned = ((p.ldc & 3) == 0) && ((p.stride_c & 3) == 0) && ((p.stride_c_inner & 3) == 0) &&
                               ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
        const int cl = lane & 7, rsub = lane >> 3;                       // read-back: chunk of 4 cols, row in group
        const bool cs_vec = z.use_col && ((reinterpret_cast<uintptr_t>(z.colsum_b) & 15) == 0) && ((z.cs_stride & 3) == 0);
        const bool bias_vec = p.bias_f32 && ((reinterpret_cast<uintptr_t>(p.bias_f32) & 15) == 0);
        const bool res_vec = p.residual && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0) && ((p.ldr & 3) == 0) &&
                             ((p.stride_r & 3) == 0);
        const int zpa = (int)z.zp_a;
        for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int64_t b = t / tiles_per_batch;
            const uint32_t r = t % tiles_per_batch;
            const int64_t m0 = (int64_t)(r / n_tiles) * BM, n0 = (int64_t)(r % n_tiles) * BN;
            const int64_t mrow0 = m0 + q * 32;
            const int64_t m = mrow0 + lane;                               // this thread's accumulator row
            const bool row_ok = m < p.M;
            const int32_t* cs_b = z.use_col ? z.colsum_b + b * z.cs_stride : nullptr;
            int64_t rowterm = -z.kterm;
            if (z.use_row && row_ok) rowterm += (int64_t)ldg_s32(z.rowsum_a + b * p.M + m) * z.zp_b;
            // Column operands of the zero-point correction / bias for the NEXT chunk are always in
            // flight one chunk ahead (first chunk: issued before waiting for the accumulator), so their
            // L2 latency never sits on the critical path (227 KB of smem leaves no L1 to hit).
            int4 ct_n = make_int4(0, 0, 0, 0), bs_n = make_int4(0, 0, 0, 0);
            auto fetch_cols = [&](int c) {
                ct_n = make_int4(0, 0, 0, 0);
                bs_n = make_int4(0, 0, 0, 0);
                const int64_t nc = n0 + c * 32;
                if (c >= BN / 32 || nc >= p.N) return;
                if (c_aligned && nc + 32 <= p.N) {
                    if (cs_b) ct_n = cs_vec ? ldg_v4(cs_b + nc + cl * 4)
                                            : make_int4(ldg_s32(cs_b + nc + cl * 4), ldg_s32(cs_b + nc + cl * 4 + 1),
                                                        ldg_s32(cs_b + nc + cl * 4 + 2), ldg_s32(cs_b + nc + cl * 4 + 3));
                    if (p.bias_f32) {
                        const float* bp = p.bias_f32 + nc + cl * 4;
                        bs_n = bias_vec ? ldg_v4(bp) : make_int4(ldg_s32(bp), ldg_s32(bp + 1), ldg_s32(bp + 2), ldg_s32(bp + 3));
                    }
                } else if (nc + lane < p.N) {
                    if (cs_b) ct_n.x = ldg_s32(cs_b + nc + lane);
                    if (p.bias_f32) bs_n.x = ldg_s32(p.bias_f32 + nc + lane);
                }
            };
            if (EMODE == EM_DEQ_FAST) fetch_cols(h);
            mbar_wait(smem_u32(tfull_bar + acc), acc_phase);
            tc_fence_after();
            if (EMODE == EM_DEQ_FAST) {
                __syncwarp();
                stg_row[lane] = (uint32_t)(int32_t)rowterm;
            }
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
            const int64_t crow_base = (p.c_inner > 1) ? (b / p.c_inner) * p.stride_c + (b % p.c_inner) * p.stride_c_inner
                                                      : b * p.stride_c;
            const int rows_left = (int)((p.M - mrow0) < 32 ? ((p.M - mrow0) > 0 ? (p.M - mrow0) : 0) : 32);
#pragma unroll 1
            for (int c = h; c < BN / 32; c += 2) {
                const int64_t nc = n0 + c * 32;
                if (nc >= p.N) break;                                     // warp-uniform
                uint32_t v[32];
                tmem_ld_32x32b_x32(t_row + (uint32_t)(c * 32), v);
                const bool vec_chunk = c_aligned && nc + 32 <= p.N;
                int ct[4];
                float bs[4];
                float4 res[8];
                if (EMODE == EM_DEQ_FAST) {
                    ct[0] = ct_n.x * zpa; ct[1] = ct_n.y * zpa; ct[2] = ct_n.z * zpa; ct[3] = ct_n.w * zpa;
                    bs[0] = __int_as_float(bs_n.x); bs[1] = __int_as_float(bs_n.y);
                    bs[2] = __int_as_float(bs_n.z); bs[3] = __int_as_float(bs_n.w);
                    fetch_cols(c + 2);
                    if (p.residual && vec_chunk && res_vec) {
                        // residual tile rows for the read-back below; in flight while TMEM drains
                        const float* rrow = p.residual + b * p.stride_r + (mrow0 + rsub) * p.ldr + nc + (cl << 2);
                        const int64_t rstep = 4 * p.ldr;
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            res[it] = (it * 4 + rsub < rows_left) ? __ldcs(reinterpret_cast<const float4*>(rrow))
                                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                            rrow += rstep;
                        }
                    }
                }
                tmem_ld_wait();
                if (EMODE == EM_REQUANT) {
                    // int8 codes: 32 bytes per row, written straight from the owning thread
                    uint32_t w[8];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
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
                        if (nc + 32 <= p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                            reinterpret_cast<uint4*>(dst)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                            reinterpret_cast<uint4*>(dst)[1] = make_uint4(w[4], w[5], w[6], w[7]);
                        } else {
                            for (int j = 0; j < 32 && nc + j < p.N; ++j) dst[j] = (int8_t)(w[j >> 2] >> ((j & 3) * 8));
                        }
                    }
                    continue;
                }
                if (EMODE == EM_DEQ_GENERAL) {
                    // general path (64-bit zero-point arithmetic) in registers, before the transpose
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
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
                // transpose through this warp's smem slab (16-byte chunks XOR-swizzled by row: both the
                // row-per-thread writes and the 4-rows-per-instruction reads are bank-conflict free)
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                        make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
                uint32_t* cbase = reinterpret_cast<uint32_t*>(p.C) + crow_base + nc;
                if (vec_chunk) {
                    // each store instruction covers 4 rows x 128 B; this lane owns 4 fixed columns
                    uint32_t* crow = cbase + (mrow0 + rsub) * p.ldc + (cl << 2);
                    const int64_t cstep = 4 * p.ldc;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rr = it * 4 + rsub;
                        uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 32 + ((cl ^ (rr & 7)) << 2));
                        if (EMODE == EM_DEQ_FAST) {
                            const int rt = (int)stg_row[rr];
                            float f0 = deq_fast((int)val.x, rt, ct[0], p.scale), f1 = deq_fast((int)val.y, rt, ct[1], p.scale);
                            float f2 = deq_fast((int)val.z, rt, ct[2], p.scale), f3 = deq_fast((int)val.w, rt, ct[3], p.scale);
                            if (p.bias_f32) {
                                f0 = __fadd_rn(bs[0], f0); f1 = __fadd_rn(bs[1], f1);
                                f2 = __fadd_rn(bs[2], f2); f3 = __fadd_rn(bs[3], f3);
                            }
                            if (p.residual) {
                                float4 rv;
                                if (res_vec) rv = res[it];
                                else {
                                    const float* rp = p.residual + b * p.stride_r + (mrow0 + rr) * p.ldr + nc + (cl << 2);
                                    rv = (rr < rows_left) ? make_float4(__ldg(rp), __ldg(rp + 1), __ldg(rp + 2), __ldg(rp + 3))
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                                }
                                f0 = __fadd_rn(f0, rv.x); f1 = __fadd_rn(f1, rv.y);
                                f2 = __fadd_rn(f2, rv.z); f3 = __fadd_rn(f3, rv.w);
                            }
                            val = make_uint4(__float_as_uint(f0), __float_as_uint(f1), __float_as_uint(f2), __float_as_uint(f3));
                        }
                        if (rr < rows_left) *reinterpret_cast<uint4*>(crow) = val;
                        crow += cstep;
                    }
                } else {
                    // ragged / unaligned: one column per lane, 32 rows, 128-byte coalesced scalar stores
                    const bool col_ok = nc + lane < p.N;
                    const int lch = lane >> 2, lw = lane & 3;
                    uint32_t* crow = cbase + mrow0 * p.ldc + lane;
                    const float* rrow = p.residual ? p.residual + b * p.stride_r + mrow0 * p.ldr + nc + lane : nullptr;
#pragma unroll 4
                    for (int rr = 0; rr < 32; ++rr) {
                        uint32_t val = stg[rr * 32 + (((lch ^ (rr & 7)) << 2) | lw)];
                        if (EMODE == EM_DEQ_FAST) {
                            float f = deq_fast((int)val, (int)stg_row[rr], ct[0], p.scale);
                            if (p.bias_f32) f = __fadd_rn(bs[0], f);
                            if (rrow && col_ok && rr < rows_left) f = __fadd_rn(f, __ldg(rrow));
                            val = __float_as_uint(f);
                        }
                        if (col_ok && rr < rows_left) *crow = val;
                        crow += p.ldc;
                        if (rrow) rrow += p.ldr;
                    }
                }
            }
            // all TMEM reads of this buffer have completed (wait::ld above)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(tempty_bar + acc));
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
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
static int make_operand_map(CUtensorMap* map, const int8_t* base, int64_t K, int64_t rows, int64_t batch, int64_t ld,
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

template <int BN, int EMODE>
static int launch_qgemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(qgemm_kernel<BN, EMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg<BN>::SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(qgemm)");
        configured = true;
    }
    const int64_t tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN) * p.batch;
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    qgemm_kernel<BN, EMODE><<<grid, NUM_THREADS, Cfg<BN>::SMEM_BYTES, s>>>(ta, tb, p);
    NQ_CHECK_LAUNCH("nq_qgemm_s8");
    return NQ_OK;
}

template <int BN>
static int launch_qgemm_mode(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
    if (p.mode == NQ_EPI_RAW) return launch_qgemm<BN, EM_RAW>(ta, tb, p, s);
    if (p.mode == NQ_EPI_REQUANT) return launch_qgemm<BN, EM_REQUANT>(ta, tb, p, s);
    return p.fast32 ? launch_qgemm<BN, EM_DEQ_FAST>(ta, tb, p, s) : launch_qgemm<BN, EM_DEQ_GENERAL>(ta, tb, p, s);
}

}  // namespace nq

using namespace nq;

extern "C" int nq_qgemm_s8(const int8_t* A, const int8_t* B, void* Cout, int64_t M, int64_t N, int64_t K,
                           int64_t batch, int64_t lda, int64_t ldb, int64_t ldc, int64_t stride_a, int64_t stride_b,
                           int64_t stride_c, const nq_epilogue* ep, void* stream) {
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
    NQ_REQUIRE(ep->mode >= NQ_EPI_RAW && ep->mode <= NQ_EPI_REQUANT, "nq_qgemm_s8: unknown epilogue mode %d", ep->mode);
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
        p.fast32 = (ep->mode == NQ_EPI_DEQUANT) && bound < 2147483000.0L;
    }
    p.bias_f32 = ep->bias_f32;
    p.bias_q = ep->bias_q;
    p.c_inner = ep->c_batch_inner > 1 ? ep->c_batch_inner : 1;
    p.stride_c_inner = ep->c_batch_inner > 1 ? ep->stride_c_inner : 0;
    NQ_REQUIRE(p.c_inner == 1 || (batch % p.c_inner == 0 && ep->mode != NQ_EPI_REQUANT),
               "nq_qgemm_s8: c_batch_inner must divide batch (and is not available with REQUANT)");
    p.residual = (ep->mode == NQ_EPI_DEQUANT) ? ep->residual : nullptr;
    p.ldr = ep->ld_residual;
    p.stride_r = ep->stride_residual;
    NQ_REQUIRE(!p.residual || p.ldr >= N, "nq_qgemm_s8: ld_residual < N");
    p.C = Cout;
    if (ep->mode == NQ_EPI_REQUANT) {
        NQ_REQUIRE(ep->out_bits >= 2 && ep->out_bits <= 8, "nq_qgemm_s8: out_bits %d outside 2..8", ep->out_bits);
        p.inv_out_scale = 1.0f / ep->out_scale;
        p.asym_out = ep->has_out_zp;
        p.out_zp = ep->has_out_zp ? (double)ep->out_zp : 0.0;
        p.lo = -ldexpf(1.f, ep->out_bits - 1);
        p.hi = ldexpf(1.f, ep->out_bits - 1) - 1.f;
    }

    const int bn = (N <= 64) ? 64 : (N <= 128) ? 128 : 256;
    CUtensorMap ta, tb;
    if (int rc = make_operand_map(&ta, A, K, M, p.a_batched ? batch : 1, lda, stride_a, BM)) return rc;
    if (int rc = make_operand_map(&tb, B, K, N, p.b_batched ? batch : 1, ldb, stride_b, bn)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (bn == 64) return launch_qgemm_mode<64>(ta, tb, p, s);
    if (bn == 128) return launch_qgemm_mode<128>(ta, tb, p, s);
    return launch_qgemm_mode<256>(ta, tb, p, s);
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
