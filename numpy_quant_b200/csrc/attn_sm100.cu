// Fused quantized attention: for every (image, head)
//     S = Q . K^T  (int8 x int8 -> int32, tcgen05)      [MatMul, numpy_quantization.py:44-61]
//     P = quantize(softmax(dequantize(S) / c))           [Div, Softmax, quantize of the next MatMul]
//     O = P . V    (int8 x int8 -> int32, tcgen05)       [MatMul]
//     C = quantize(dequantize(O)) scattered as [B, S, H*D] (Transpose(0,2,1,3) + Reshape + quantize of the output
//         projection's left operand)
// in ONE kernel: the scores, the probabilities and the context accumulator never leave the SM -- S and O live in
// TMEM, P goes from registers to a 128-byte-swizzled K-major shared-memory tile that the second MMA reads.
// Same arithmetic as NQ_EPI_SOFTMAX_QUANT followed by NQ_EPI_QUANT (merge heads): identical codes.
//
// Persistent, one CTA (640 threads) per SM, tile = (image*head, 128 query rows):
//   warp 0      TMA producer: Q tile [128 x D], K tile [208 x D], V^T tile [64 x S] per stage (3 stages)
//   warp 1      MMA issuer:   S(i) one tile ahead of the softmax, O(i-1) as soon as P(i-1) is in shared memory
//   warp 2      TMEM allocator (2 x 208 columns for S, 64 for O)
//   warps 4-19  softmax of tile i (as in qgemm_sm100.cu), P(i) -> smem
//   warps 20-23 context epilogue of tile i (dequantize O, quantize, merge heads) -- concurrent with the softmax of
//               tile i+1, so its latencies fill the softmax warps' idle issue slots
// Registers (launch 768 x 80 = 61440): setmaxnreg 48 / 96 / 48.
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace nq {

int make_operand_map(CUtensorMap* map, const int8_t* base, int64_t K, int64_t rows, int64_t batch, int64_t ld,
                     int64_t batch_stride, int box_rows);

namespace attn {

// Waiters that are off the critical path (TMA producer: 3 stages ahead; context warps: a tile behind) back off with
// nanosleep: their poll loops otherwise take ~15 % of the issue slots the softmax warps need.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 1; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done) {
            __nanosleep(256);
            if (spin > (1u << 24)) __trap();                              // ~4 s: protocol bug -> error instead of a hang
        }
    }
}

constexpr int NUM_EPI_WARPS = 16;
constexpr int NUM_CTX_WARPS = 4;                // context epilogue: one warp per TMEM lane quarter
constexpr int NUM_THREADS = 128 + 32 * NUM_EPI_WARPS + 32 * NUM_CTX_WARPS;   // 768
constexpr int BN1 = 208;                       // key columns of the score tile (S <= 208)
constexpr int BN2 = 64;                        // head dim columns of the context tile (D <= 64)
constexpr int STAGES = 3;
constexpr int Q_BYTES = BM * BK;               // 16 KB (box 128 B wide; D < 128 is zero-filled by TMA)
constexpr int K_BYTES = BN1 * BK;              // 26 KB
constexpr int V_BYTES = 2 * BN2 * BK;          // 16 KB: two k-blocks of the key axis
constexpr int STAGE_BYTES = Q_BYTES + K_BYTES + V_BYTES;
constexpr int P_BYTES = 2 * BM * BK;           // 32 KB: P as the K-major A operand of the second MMA (two k-blocks)
constexpr int EPI_WORDS = 4864;                // max / sum / code-sum exchange, per-warp column / row terms (cp.async targets)
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + P_BYTES + EPI_WORDS * 4 + BAR_BYTES;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB opt-in shared memory of sm_100");
constexpr int O_COL = 2 * BN1;                 // TMEM column of the context accumulator (416)

struct Params {
    int64_t BH, S, D, H;
    // scores
    float scale1;                              // s_q * s_k / c  (Div folded)
    const int32_t* rowsum_q;                   // [BH, S]  (used when K^T is asymmetric)
    const int32_t* colsum_k;                   // [BH, S]  (used when Q is asymmetric)
    int zq, zk, use_row1, use_col1;
    int64_t kterm1;                            // zq * zk * D
    int fast22, sm_noclamp;                    // sm_noclamp: zp_p >= lo, so p / s + zp needs no lower clamp
    float sm_top;                              // upper clamp in the magic-sum domain (1.5 * 2^23 + hi), or huge when p = 1 fits
    QArgs qp;                                  // P quantizer
    // context
    float scale2;                              // s_p * s_v
    const int32_t* colsum_v;                   // [BH, D]
    int zp_p, zv, use_row2, use_col2;
    int64_t kterm2;                            // zp_p * zv * S
    QArgs qo;                                  // output quantizer
    int8_t* C;                                 // [B, S, H * D]
    int32_t* o_rowsum;                         // [B * S] or NULL (atomics; caller-zeroed)
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
            const __grid_constant__ CUtensorMap tmap_v, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    uint8_t* smem = smem_raw;
    uint8_t* smem_p = smem + STAGES * STAGE_BYTES;
    uint32_t* epi = reinterpret_cast<uint32_t*>(smem + STAGES * STAGE_BYTES + P_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + P_BYTES + EPI_WORDS * 4);
    uint64_t* full_bar = bars;                 // [STAGES] TMA -> MMA
    uint64_t* empty_bar = bars + STAGES;       // [STAGES] second MMA done with the stage
    uint64_t* sfull_bar = bars + 2 * STAGES;   // [2] scores ready
    uint64_t* sempty_bar = sfull_bar + 2;      // [2] scores drained (16 warps)
    uint64_t* pfull_bar = sempty_bar + 2;      // P tile written (16 warps)
    uint64_t* ofull_bar = pfull_bar + 1;       // context accumulator ready
    uint64_t* oempty_bar = ofull_bar + 1;      // context accumulator drained (4 context warps)
    uint64_t* rs_bar = oempty_bar + 1;         // [2] row sums of P handed to the context warps (one warp per quarter);
                                               // by tile parity: the softmax may complete tile i + 2 before a context warp
                                               // polls for tile i + 1, which a single barrier's phase parity cannot tell apart
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rs_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t m_tiles = (uint32_t)((p.S + BM - 1) / BM);
    const uint32_t total_tiles = m_tiles * (uint32_t)p.BH;
    const int ks1 = (int)((p.D + UMMA_K - 1) / UMMA_K);                   // K steps of Q.K^T (<= 4)
    const int ks2 = (int)((p.S + UMMA_K - 1) / UMMA_K);                   // K steps of P.V (<= 8 over two k-blocks)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_k) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_v) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(smem_u32(full_bar + i), 1);
            mbar_init(smem_u32(empty_bar + i), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(sfull_bar + i), 1);
            mbar_init(smem_u32(sempty_bar + i), NUM_EPI_WARPS);
        }
        mbar_init(smem_u32(pfull_bar), NUM_EPI_WARPS);
        mbar_init(smem_u32(ofull_bar), 1);
        mbar_init(smem_u32(oempty_bar), NUM_CTX_WARPS);
        mbar_init(smem_u32(rs_bar), 4);
        mbar_init(smem_u32(rs_bar + 1), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        if (warp == 0) {
            // ===================== TMA producer =====================
            if (lane == 0) {
                int stage = 0;
                uint32_t phase = 0;
                for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                    const int bh = (int)(t / m_tiles), m0 = (int)(t % m_tiles) * BM;
                    mbar_wait_relaxed(smem_u32(empty_bar + stage), phase ^ 1);
                    const uint32_t fb = smem_u32(full_bar + stage);
                    uint8_t* st = smem + stage * STAGE_BYTES;
                    mbar_expect_tx(fb, STAGE_BYTES);
                    tma_load_3d(smem_u32(st), &tmap_q, 0, m0, bh, fb);
                    tma_load_3d(smem_u32(st + Q_BYTES), &tmap_k, 0, 0, bh, fb);
                    tma_load_3d(smem_u32(st + Q_BYTES + K_BYTES), &tmap_v, 0, 0, bh, fb);
                    tma_load_3d(smem_u32(st + Q_BYTES + K_BYTES + BN2 * BK), &tmap_v, BK, 0, bh, fb);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // ===================== MMA issuer =====================
            if (lane == 0) {
                constexpr uint32_t idesc1 = make_idesc(BN1), idesc2 = make_idesc(BN2);
                const uint64_t pdesc = make_smem_desc(smem_u32(smem_p));
                uint32_t li = 0;                                          // local tile counter
                int prev_stage = -1;
                for (uint32_t t = blockIdx.x;; t += gridDim.x, ++li) {
                    const bool have = t < total_tiles;
                    const int stage = (int)(li % STAGES);
                    if (have) {
                        // ---- scores of tile li (one tile ahead of the softmax warps)
                        const int sb = (int)(li & 1);
                        mbar_wait_backoff(smem_u32(sempty_bar + sb), ((li >> 1) & 1u) ^ 1u);
                        mbar_wait_backoff(smem_u32(full_bar + stage), (li / STAGES) & 1u);
                        tc_fence_after();
                        uint8_t* st = smem + stage * STAGE_BYTES;
                        const uint64_t qd = make_smem_desc(smem_u32(st)), kd = make_smem_desc(smem_u32(st + Q_BYTES));
                        for (int k = 0; k < ks1; ++k)
                            mma_i8(tmem_base + (uint32_t)(sb * BN1), qd + (uint64_t)(k * 2), kd + (uint64_t)(k * 2), idesc1, k > 0 ? 1u : 0u);
                        tc_commit(smem_u32(sfull_bar + sb));
                    }
                    if (li > 0) {
                        // ---- context of tile li - 1: P is in shared memory, the accumulator buffer is free
                        const uint32_t lp = li - 1;
                        mbar_wait_backoff(smem_u32(pfull_bar), lp & 1u);
                        mbar_wait_backoff(smem_u32(oempty_bar), (lp & 1u) ^ 1u);
                        tc_fence_after();
                        uint8_t* st = smem + prev_stage * STAGE_BYTES;
                        const uint64_t vd0 = make_smem_desc(smem_u32(st + Q_BYTES + K_BYTES));
                        const uint64_t vd1 = make_smem_desc(smem_u32(st + Q_BYTES + K_BYTES + BN2 * BK));
                        const uint64_t pd1 = make_smem_desc(smem_u32(smem_p + BM * BK));
                        for (int k = 0; k < ks2; ++k) {
                            const int kk = k & 3;
                            if (k < 4) mma_i8(tmem_base + O_COL, pdesc + (uint64_t)(kk * 2), vd0 + (uint64_t)(kk * 2), idesc2, k > 0 ? 1u : 0u);
                            else mma_i8(tmem_base + O_COL, pd1 + (uint64_t)(kk * 2), vd1 + (uint64_t)(kk * 2), idesc2, 1u);
                        }
                        tc_commit(smem_u32(empty_bar + prev_stage));      // Q / K / V of that tile no longer needed
                        tc_commit(smem_u32(ofull_bar));
                    }
                    prev_stage = stage;
                    if (!have) break;
                }
            }
            __syncwarp();
        }
    } else if (warp >= 4 + NUM_EPI_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        // ===================== context epilogue (4 warps, one per TMEM lane quarter) =====================
        // O(i) = P(i) . V is drained here while the softmax warps already work on tile i + 1.
        const int q = warp & 3;
        const int rloc = q * 32 + lane;
        const int* rsbuf = reinterpret_cast<const int*>(epi) + 4608;      // [2][128] row sums of P (from the softmax warps)
        const Quantizer qzo(p.qo);
        uint32_t li = 0;
        for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x, ++li) {
            const uint32_t bh = t / m_tiles, m0 = (t - bh * m_tiles) * BM;
            const int64_t m = (int64_t)m0 + rloc;
            const bool row_ok = m < p.S;
            const bool warp_rows = (int64_t)m0 + q * 32 < p.S;
            const uint32_t b = bh / (uint32_t)p.H, hh = bh - b * (uint32_t)p.H;
            // column terms of this tile (colsum(V) * zp_p, 64 values, the same for every row): one 16-byte load per
            // lane into the warp's smem strip before the waits, broadcast LDS.128 in the loop
            int* ctv = reinterpret_cast<int*>(epi) + 4096 + q * 64;
            __syncwarp();
            if (lane < 16) {
                int4 c4 = make_int4(0, 0, 0, 0);
                if (p.use_col2 && lane * 4 < p.D) c4 = ldg_v4(p.colsum_v + (int64_t)bh * p.D + lane * 4);
                *reinterpret_cast<int4*>(ctv + lane * 4) = make_int4(c4.x * p.zp_p, c4.y * p.zp_p, c4.z * p.zp_p, c4.w * p.zp_p);
            }
            __syncwarp();
            mbar_wait_relaxed(smem_u32(rs_bar + (li & 1)), (li >> 1) & 1u);
            int rowterm = (int)-p.kterm2;
            if (p.use_row2) rowterm += rsbuf[(li & 1) * 128 + rloc] * p.zv;
            mbar_wait_relaxed(smem_u32(ofull_bar), li & 1u);
            tc_fence_after();
            int8_t* dst = p.C + (((int64_t)b * p.S + m) * p.H + hh) * p.D;
            int rs_out = 0;
            if (warp_rows) {
#pragma unroll 1
                for (int c16 = 0; c16 * 16 < p.D; ++c16) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(O_COL + c16 * 16), v);
                    tmem_ld_wait();
                    int w[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int4 c4 = *reinterpret_cast<const int4*>(ctv + c16 * 16 + g * 4);
                        const int ct[4] = {c4.x, c4.y, c4.z, c4.w};
                        int c[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            // |acc - zero-point terms| <= S * 255 * 255 < 2^24: exact int -> float, one IEEE multiply
                            const int d = (int)v[4 * g + k] - rowterm - ct[k];
                            c[k] = qzo.code<1>(__fmul_rn(__int2float_rn(d), p.scale2));
                        }
                        w[g] = pack4_codes(c[0], c[1], c[2], c[3]);
                    }
                    if (row_ok) {
                        *reinterpret_cast<int4*>(dst + c16 * 16) = make_int4(w[0], w[1], w[2], w[3]);
                        rs_out = __dp4a(w[0], 0x01010101, __dp4a(w[1], 0x01010101, __dp4a(w[2], 0x01010101, __dp4a(w[3], 0x01010101, rs_out))));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(oempty_bar));
            if (p.o_rowsum && row_ok && warp_rows) atomicAdd(p.o_rowsum + (int64_t)b * p.S + m, rs_out);
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        // ===================== softmax (16 warps) =====================
        const int q = warp & 3, h = (warp - 4) >> 2, ew = warp - 4;
        const int rloc = q * 32 + lane;
        float* red = reinterpret_cast<float*>(epi);                        // [2][4][128] max / sum exchange
        int* redq = reinterpret_cast<int*>(epi) + 1024;                    // [4][128] code sums
        int* ctw2 = reinterpret_cast<int*>(epi) + 2048 + ew * 128;         // this warp's 56 score column terms, 2 tile buffers
        int* rowraw = reinterpret_cast<int*>(epi) + 1536 + ew * 32;        // this warp's rowsum(Q) values of the next tile
        int* rsbuf = reinterpret_cast<int*>(epi) + 4608;                   // [2][128] row sums of P for the context warps
        constexpr int NSUB = 7;
        constexpr float kMasked = -1.0e30f;
        const int col0 = h * (NSUB * 8);
        const int ncols_w = (int)(p.S - col0 < NSUB * 8 ? (p.S - col0 > 0 ? p.S - col0 : 0) : NSUB * 8);
        const int nfull = ncols_w >> 3, nrem = ncols_w & 7;
        // Per-tile operands of the zero-point correction (rowsum(Q) of this thread's row, colsum(K) of this warp's 56
        // columns) travel global -> shared memory with cp.async ONE TILE AHEAD: no registers held across the tile, no
        // L2 round trip at the head of a tile's critical path.
        auto stage_next = [&](uint32_t tt, int buf) {
            if (tt >= total_tiles) return;
            const uint32_t bh = tt / m_tiles, m0 = (tt - bh * m_tiles) * BM;
            const int64_t m = (int64_t)m0 + rloc;
            if (p.use_row1 && m < p.S)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(rowraw + lane)), "l"(p.rowsum_q + (int64_t)bh * p.S + m) : "memory");
            if (p.use_col1) {
                const int64_t c0i = col0 + lane, c1i = c0i + 32;
                const int32_t* src = p.colsum_k + (int64_t)bh * p.S;
                if (c0i < p.S)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(ctw2 + buf * 64 + lane)), "l"(src + c0i) : "memory");
                if (lane < 24 && c1i < p.S)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(ctw2 + buf * 64 + 32 + lane)), "l"(src + c1i) : "memory");
            }
        };
        stage_next(blockIdx.x, 0);
        uint32_t li = 0;
        for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x, ++li) {
            const uint32_t bh = t / m_tiles, m0 = (t - bh * m_tiles) * BM;
            const int64_t m = (int64_t)m0 + rloc;
            const bool row_ok = m < p.S;
            const int rows_left = (int)(p.S - (m0 + q * 32));
            const bool warp_rows = rows_left > 0;
            asm volatile("cp.async.wait_all;" ::: "memory");               // this tile's terms (issued one tile ago) have landed
            __syncwarp();
            int* ctw = ctw2 + (li & 1) * 64;
            int rowterm = (int)-p.kterm1;
            if (p.use_row1 && row_ok) rowterm += rowraw[lane] * p.zk;
            if (p.use_col1) {                                              // raw column sums -> column terms, in place
                const int v0 = ctw[lane] * p.zq, v1 = ctw[32 + lane] * p.zq;
                __syncwarp();
                ctw[lane] = v0;
                ctw[32 + lane] = v1;
            } else {
                ctw[lane] = 0;
                ctw[32 + lane] = 0;
            }
            __syncwarp();
            stage_next(t + gridDim.x, (int)((li + 1) & 1));
            const int sb = (int)(li & 1);
            mbar_wait(smem_u32(sfull_bar + sb), (li >> 1) & 1u);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sb * BN1);
            // ---------------- pass 1: scores -> registers, row max
            float y[NSUB * 8];
            float lmax = kMasked;
            auto pass1 = [&](auto magic_tag) {
                constexpr bool MAGIC = decltype(magic_tag)::value;
                const int rm = (MAGIC ? 0x4B400000 : 0) - rowterm;
#pragma unroll
                for (int j2 = 0; j2 < NSUB; j2 += 2) {
                    if (j2 * 8 < ncols_w) {
                        uint32_t a16[16];
                        const bool two = (j2 + 1 < NSUB) && ((j2 + 1) * 8 < ncols_w);
                        if (two) tmem_ld_32x32b_x16(t_row + (uint32_t)(col0 + j2 * 8), a16);
                        else tmem_ld_32x32b_x8(t_row + (uint32_t)(col0 + j2 * 8), a16);
                        int c16[16];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int4 c4 = *reinterpret_cast<const int4*>(ctw + j2 * 8 + g * 4);
                            c16[4 * g] = c4.x; c16[4 * g + 1] = c4.y; c16[4 * g + 2] = c4.z; c16[4 * g + 3] = c4.w;
                        }
                        tmem_ld_wait();
                        const bool clean = j2 * 8 + 16 <= ncols_w || (j2 + 1 >= NSUB && j2 * 8 + 8 <= ncols_w);
                        if (clean) {
                            // two columns per instruction (packed add, packed multiply: each lane IEEE-rounded exactly
                            // like the scalar pair, so the codes stay those of the SOFTMAX_QUANT epilogue)
#pragma unroll
                            for (int k = 0; k < 16; k += 2) {
                                if (j2 * 8 + k >= NSUB * 8) break;
                                const int x0 = (int)a16[k] + rm - c16[k], x1 = (int)a16[k + 1] + rm - c16[k + 1];
                                const float2 s2 = make_float2(p.scale1, p.scale1);
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
                                float f = MAGIC ? __fmul_rn(__fadd_rn(__int_as_float(x), -12582912.0f), p.scale1)
                                                : __fmul_rn(__int2float_rn(x), p.scale1);
                                if (j2 * 8 + k >= ncols_w) f = kMasked;
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
                pass1(std::true_type{});               // host-checked: |score - zero-point terms| < 2^22
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(sempty_bar + sb));        // scores are in registers
            red[h * 128 + rloc] = lmax;
            named_bar_sync(1 + q, 128);
            const float gmax = fmaxf(fmaxf(red[rloc], red[128 + rloc]), fmaxf(red[256 + rloc], red[384 + rloc]));
            // ---------------- pass 2: exp, row sum
            float lsum = 0.f;
            if (warp_rows) {
                const float l2e = 1.44269502162933349609375f;
                const float m2 = __fmul_rn(gmax, l2e);
                // four partial sums (columns k & 3) as two packed accumulators: same additions, same order
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
            const float gsum = __fadd_rn(__fadd_rn(red[512 + rloc], red[640 + rloc]), __fadd_rn(red[768 + rloc], red[896 + rloc]));
            // ---------------- the second MMA of the previous tile has finished reading the P tile (long ago: it was
            // issued before this tile's pass 1); from here on P(i) may overwrite it
            if (li > 0) {
                mbar_wait(smem_u32(ofull_bar), (li - 1) & 1u);
                tc_fence_after();
            }
            // ---------------- pass 3: codes of P straight into shared memory: K-major, 128-byte swizzle
            // (row r, 16-byte chunk ^ (r & 7)), i.e. the layout TMA would have produced for an A operand
            int qsum = 0;
            if (warp_rows) {
                const float kr = __frcp_rn(__fmul_rn(gsum, p.qp.scale));
                const Quantizer qzp(p.qp);
                auto emit_group = [&](int j, auto ragged_tag) {
                    constexpr bool RAGGED = decltype(ragged_tag)::value;
                    int c[8];
                    const float2 kr2 = make_float2(kr, kr), mg2 = make_float2(qzp.magic, qzp.magic);
#pragma unroll
                    for (int k = 0; k < 8; k += 2) {
                        const float2 e = make_float2(y[j * 8 + k], y[j * 8 + k + 1]);
                        if (p.sm_noclamp) {
                            const float2 r = __ffma2_rn(e, kr2, mg2);
                            c[k] = __float_as_int(fminf(r.x, p.sm_top));
                            c[k + 1] = __float_as_int(fminf(r.y, p.sm_top));
                        } else {
                            const float2 t = __fmul2_rn(e, kr2);
                            const float2 r = __fadd2_rn(make_float2(fminf(fmaxf(t.x, qzp.tlo), qzp.thi), fminf(fmaxf(t.y, qzp.tlo), qzp.thi)), mg2);
                            c[k] = __float_as_int(r.x);
                            c[k + 1] = __float_as_int(r.y);
                        }
                        if (RAGGED) {
                            c[k] = (k < nrem) ? c[k] : 0;
                            c[k + 1] = (k + 1 < nrem) ? c[k + 1] : 0;
                        }
                    }
                    int w0 = pack4_codes(c[0], c[1], c[2], c[3]), w1 = pack4_codes(c[4], c[5], c[6], c[7]);
                    if (!row_ok) w0 = w1 = 0;                             // rows past S: zeros (their outputs are never stored)
                    qsum = __dp4a(w0, 0x01010101, __dp4a(w1, 0x01010101, qsum));
                    const int cc0 = col0 + j * 8, kb = cc0 >> 7, cc = cc0 & 127;
                    uint8_t* dst = smem_p + kb * (BM * BK) + rloc * 128 + ((((cc >> 4) ^ (rloc & 7)) << 4) | (cc & 15));
                    *reinterpret_cast<int2*>(dst) = make_int2(w0, w1);
                };
#pragma unroll
                for (int j = 0; j < NSUB; ++j) {
                    if (j < nfull) emit_group(j, std::false_type{});
                    else if (j == nfull && nrem > 0) emit_group(j, std::true_type{});
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(pfull_bar));
            redq[h * 128 + rloc] = qsum;
            named_bar_sync(1 + q, 128);
            if (h == 0) {
                // row sums of P(i) -> the context warps (double-buffered by tile parity)
                rsbuf[(li & 1) * 128 + rloc] = (redq[rloc] + redq[128 + rloc]) + (redq[256 + rloc] + redq[384 + rloc]);
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(rs_bar + (li & 1)));
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

}  // namespace attn
}  // namespace nq

using namespace nq;

extern "C" int nq_attention_s8(const int8_t* Q, const int8_t* Kt, const int8_t* Vt, int64_t BH, int64_t H, int64_t S, int64_t D,
                               int64_t ld_q, int64_t ld_k, int64_t ld_v, const nq_attention* a, void* stream) {
    NQ_REQUIRE(a, "nq_attention_s8: descriptor is NULL");
    NQ_REQUIRE(BH > 0 && H > 0 && BH % H == 0 && S > 0 && D > 0, "nq_attention_s8: bad extents");
    NQ_REQUIRE(S <= attn::BN1 && D <= attn::BN2 && D % 16 == 0, "nq_attention_s8: built for S <= 208, D <= 64, D %% 16 == 0 (S=%lld D=%lld)",
               (long long)S, (long long)D);
    NQ_REQUIRE(ld_q % 16 == 0 && ld_k % 16 == 0 && ld_v % 16 == 0 && ld_q >= D && ld_k >= D && ld_v >= S,
               "nq_attention_s8: operand row strides must be >= the contraction length and multiples of 16");
    NQ_REQUIRE(((uintptr_t)Q % 16 == 0) && ((uintptr_t)Kt % 16 == 0) && ((uintptr_t)Vt % 16 == 0) && ((uintptr_t)a->out % 16 == 0),
               "nq_attention_s8: operands must be 16-byte aligned");
    NQ_REQUIRE(a->p_bits >= 2 && a->p_bits <= 8 && a->out_bits >= 2 && a->out_bits <= 8, "nq_attention_s8: bit widths outside 2..8");
    NQ_REQUIRE(!a->has_zk || a->rowsum_q, "nq_attention_s8: rowsum_q required when K is asymmetric");
    NQ_REQUIRE(!a->has_zq || a->colsum_k, "nq_attention_s8: colsum_k required when Q is asymmetric");
    NQ_REQUIRE(!a->has_p_zp || a->colsum_v, "nq_attention_s8: colsum_v required when P is asymmetric");
    NQ_REQUIRE(BH * ((S + 127) / 128) < (1ll << 31), "nq_attention_s8: too many tiles");
    attn::Params p{};
    p.BH = BH; p.S = S; p.D = D; p.H = H;
    p.scale1 = a->scale_qk;
    if (a->has_div) {
        NQ_REQUIRE(a->div > 0.f && isfinite(a->div), "nq_attention_s8: divisor must be positive and finite");
        p.scale1 = a->scale_qk / a->div;
    }
    NQ_REQUIRE(p.scale1 > 1e-30f && p.scale1 < 1e30f, "nq_attention_s8: scale / divisor out of range");
    p.rowsum_q = a->rowsum_q; p.colsum_k = a->colsum_k;
    p.zq = a->has_zq ? (int)a->zq : 0; p.zk = a->has_zk ? (int)a->zk : 0;
    p.use_row1 = a->has_zk; p.use_col1 = a->has_zq;
    p.kterm1 = (a->has_zq && a->has_zk) ? (int64_t)a->zq * a->zk * D : 0;
    {
        const long double zq = p.zq, zk = p.zk;
        const long double ra = fmaxl(fabsl(-128.0L - zq), fabsl(127.0L - zq)), rb = fmaxl(fabsl(-128.0L - zk), fabsl(127.0L - zk));
        NQ_REQUIRE(ra * rb * (long double)D < 2147483000.0L, "nq_attention_s8: score zero-point terms exceed int32");
        p.fast22 = (ra * rb * (long double)D) < 4194304.0L;
        NQ_REQUIRE(p.fast22, "nq_attention_s8: max|q - zq| * max|k - zk| * D must stay below 2^22 (zero-points outside the int8 range?)");
    }
    int qmode;
    p.qp = make_qargs(a->p_bits, a->p_scale, a->has_p_zp, a->p_zp, &qmode);
    NQ_REQUIRE(qmode != 2, "nq_attention_s8: |p_zp| must be < 2^20");
    {
        // probabilities lie in [0, 1]: with zp >= lo the quotient p / s + zp never needs the lower clamp and rounds
        // straight out of one FMA with the magic constant; the upper clamp is a min in that domain, and vanishes
        // (huge bound) when even p = 1 maps inside the code range
        const double top = (double)p.qp.zpf + 1.0 / (double)p.qp.scale * (1.0 + 1e-6);
        p.sm_noclamp = (double)p.qp.zpf >= (double)p.qp.lo;
        p.sm_top = (top < (double)p.qp.hi + 0.49) ? 3.0e38f : 12582912.0f + p.qp.hi;
    }
    p.scale2 = a->scale_pv;
    p.colsum_v = a->colsum_v;
    p.zp_p = a->has_p_zp ? (int)a->p_zp : 0; p.zv = a->has_zv ? (int)a->zv : 0;
    p.use_row2 = a->has_zv; p.use_col2 = a->has_p_zp;
    p.kterm2 = (a->has_p_zp && a->has_zv) ? (int64_t)a->p_zp * a->zv * S : 0;
    {
        const long double zp = p.zp_p, zv = p.zv;
        const long double ra = fmaxl(fabsl(-128.0L - zp), fabsl(127.0L - zp)), rb = fmaxl(fabsl(-128.0L - zv), fabsl(127.0L - zv));
        NQ_REQUIRE(ra * rb * (long double)S <= 16777216.0L, "nq_attention_s8: context accumulator exceeds 2^24 (exact float window)");
    }
    p.qo = make_qargs(a->out_bits, a->out_scale, a->has_out_zp, a->out_zp, &qmode);
    NQ_REQUIRE(qmode != 2, "nq_attention_s8: |out_zp| must be < 2^20");
    p.C = a->out;
    p.o_rowsum = a->out_rowsum;
    cudaStream_t s = (cudaStream_t)stream;
    if (a->out_rowsum) {
        cudaError_t e = cudaMemsetAsync(a->out_rowsum, 0, sizeof(int32_t) * (size_t)(BH / H * S), s);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(out_rowsum)");
    }
    CUtensorMap tq, tk, tv;
    if (int rc = make_operand_map(&tq, Q, D, S, BH, ld_q, S * ld_q, BM)) return rc;
    if (int rc = make_operand_map(&tk, Kt, D, S, BH, ld_k, S * ld_k, attn::BN1)) return rc;
    if (int rc = make_operand_map(&tv, Vt, S, D, BH, ld_v, D * ld_v, attn::BN2)) return rc;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attn::attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn)");
        configured = true;
    }
    const int64_t tiles = BH * ((S + BM - 1) / BM);
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    attn::attn_kernel<<<grid, attn::NUM_THREADS, attn::SMEM_BYTES, s>>>(tq, tk, tv, p);
    NQ_CHECK_LAUNCH("nq_attention_s8");
    return NQ_OK;
}
