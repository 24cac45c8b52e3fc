// Warp-specialised tcgen05 / TMEM / TMA fused attention forward (sm_100a):  O = softmax((Q*scale) K^T) V, d_qk = 32, d_v in {32, 64}.
//
// Same contract and the same arithmetic as tc_attention.cu (reference: AAttn / Attention, block.py:1708-1722 / :1324-1331); what
// changes is WHO does what, so that the exp unit never waits for a hand-off:
//
//   * one CTA owns TWO 128-row query tiles of one (image, head) and walks the keys once for both: every K / V tile is fetched
//     once per 256 query rows (tc_attention.cu: once per 128) by TMA (one 4-D tensor map per operand: d, token, head, image;
//     rows past N arrive as zeros) into a 6-stage shared-memory ring in the 64B/128B-swizzled canonical UMMA layout;
//   * each query tile has its own issuing thread (warp 8 lane 0: all TMA + tile A's tcgen05.mma; warp 9 lane 0: tile B's).  Per key
//     tile t it issues QK(t+1) as soon as ITS softmax warps have READ S(t) out of tensor memory - one whole softmax phase before
//     they need S(t+1) - then PV(t) when P(t) is complete, so the tensor-core latency is never on the softmax warps' path and the
//     two tiles never wait for each other (they only share the K/V ring, released by one commit-arrive per tile);
//   * warps 0-3 (tile A) and 4-7 (tile B) are softmax warps: thread = query row (shuffle-free max / sum), S row read with
//     tcgen05.ld, exp2 on the MUFU, lazy rescale (O only touched when a row max grows by more than 2^8), packed fp16 P row
//     written back to tensor memory (tcgen05.st) as the A operand of the TS-mode PV MMA.  They never meet at a __syncthreads:
//     every hand-off is an mbarrier (S full / S consumed / P full / PV done), arrived on by one lane per warp.
//
// With 2 CTAs per SM every SM sub-partition holds 4 softmax warps of 4 independent tiles; a warp's non-exp section per key tile
// (tensor-memory load, P store, two barrier polls) is ~300 clk against 4 x 512 clk of MUFU work per round, where tc_attention.cu's
// serial section (S wait -> softmax -> PV wait -> __syncthreads -> MMA issue -> MMA latency) was ~1200 clk and convoyed all
// four CTAs of an SM into the same phase.
//
// Tensor memory per CTA: per query tile S 64 + P 32 + O d_v columns (fp32 / packed fp16 / fp32) -> 256 columns (d_v = 32).
#include <cuda.h>

#include "tc_common.cuh"

namespace ym {

constexpr int A2_BQ = 128, A2_BKV = 64, A2_STAGES = 8, A2_THREADS = 320;   // 8 softmax warps + 2 issuer warps (one per query tile)

__device__ __forceinline__ float a2_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}
// exp2 on the FMA / ALU pipes (same polynomial as tc_attention.cu): round-to-nearest range reduction by the 1.5*2^23 magic add, degree-3
// minimax polynomial of 2^f on [-0.5, 0.5] (max relative error 7.5e-5, a sixth of the fp16 rounding of P), exponent inserted by an
// integer add.  With POLY = n every n-th score of a row takes this path, so the MUFU (16 ex2 / clk / SM) serves (n-1)/n of them.
__device__ __forceinline__ float a2_exp2_poly(float x) {
    x = fmaxf(x, -125.f);
    const float t = x + 12582912.f;
    const float f = x - (t - 12582912.f);
    float p = fmaf(0.05517164617776871f, f, 0.2426111251115799f);
    p = fmaf(p, f, 0.6932609677314758f);
    p = fmaf(p, f, 0.9999280571937561f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ void a2_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void a2_tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

struct A2Bars {
    uint64_t q_full;
    uint64_t kv_full[A2_STAGES], kv_empty[A2_STAGES];
    uint64_t s_full[2], s_free[2], p_full[2], pv_done[2];
    uint32_t tmem_slot;
};

template <int DV, int POLY, int VAR>
__global__ void __launch_bounds__(A2_THREADS, (DV == 32) ? 2 : 1)
tc_attention2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, int N, float scale_log2, __half* __restrict__ out, int ldo, int q_tiles, int var_flags) {
    constexpr int Q_BYTES = A2_BQ * 64, K_BYTES = A2_BKV * 64, V_BYTES = A2_BKV * DV * 2, STAGE_BYTES = K_BYTES + V_BYTES;
    constexpr uint32_t QT_COLS = 64 + 32 + DV;                 // tensor-memory columns of one query tile: S | P | O
    constexpr uint32_t TMEM_COLS = (DV == 32) ? 256 : 512;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (STS, not generic ST)
    unsigned char* sQ = smem;                                  // 2 x [128 rows x 64 B], SW64
    unsigned char* sKV = sQ + 2 * Q_BYTES;                     // [STAGES][K tile | V tile]
    __shared__ A2Bars bars;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * (q_tiles * A2_BQ), h = blockIdx.y, b = blockIdx.z;
    const bool has_b = q_tiles == 2 && q0 + A2_BQ < N;          // the second query tile is in use and holds at least one real row
    const int T = (N + A2_BKV - 1) / A2_BKV;

    if (tid == 0) {
        tc::mbar_init(&bars.q_full, 1);
        for (int s = 0; s < A2_STAGES; ++s) {
            tc::mbar_init(&bars.kv_full[s], 1);
            tc::mbar_init(&bars.kv_empty[s], has_b ? 2 : 1);   // one commit-arrive per query tile that reads the stage
        }
        for (int q = 0; q < 2; ++q) {
            tc::mbar_init(&bars.s_full[q], 1);
            tc::mbar_init(&bars.s_free[q], 4);     // one arrive per softmax warp
            tc::mbar_init(&bars.p_full[q], 4);
            tc::mbar_init(&bars.pv_done[q], 1);
        }
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 8) tc::tmem_alloc(&bars.tmem_slot, TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = bars.tmem_slot;
    pdl_prologue();    // barriers / tensor memory are set up: let the next kernel stage itself, then wait for the qkv conv

    if (warp >= 8) {
        // ====================================== issuers: warp 8 = TMA + tile A's MMAs, warp 9 = tile B's MMAs ======================
        // Each query tile has its OWN issuing thread, so a tile's QK(t+1) is issued the moment ITS softmax warps have read S(t) and
        // its PV(t) the moment ITS P(t) is complete, whatever the sibling tile is doing.  (One thread serving both tiles in a static
        // order locked them one key tile apart: tile A's next S was only issued when tile B finished - ~900 clk of s_full polling per
        // key tile in profiles/r02_attention2_ncu.txt, first capture.)
        const int q = warp - 8;
        const int refill_lag = (var_flags & 2) ? 3 : 1;
        const bool relax = (var_flags & 4) != 0;
        auto iwait = [&](uint64_t* bar, uint32_t parity) {
            if (relax) tc::mbar_wait_sleep(bar, parity, 32);
            else tc::mbar_wait(bar, parity);
        };
        if (lane == 0 && (q == 0 || has_b)) {
            const uint32_t idesc_qk = tc::make_idesc_f16(A2_BQ, A2_BKV, 0);
            const uint32_t idesc_pv = tc::make_idesc_f16(A2_BQ, DV, 1);      // V is MN-major (d_v contiguous)
            auto load_kv = [&](int t) {
                const int st = t % A2_STAGES;
                unsigned char* dK = sKV + st * STAGE_BYTES;
                a2_expect_tx(&bars.kv_full[st], (uint32_t)STAGE_BYTES);
                a2_tma_load_4d(dK, &map_k, 0, t * A2_BKV, h, b, &bars.kv_full[st]);
                a2_tma_load_4d(dK + K_BYTES, &map_v, 0, t * A2_BKV, h, b, &bars.kv_full[st]);
            };
            if (q == 0) {
                a2_expect_tx(&bars.q_full, (uint32_t)(2 * Q_BYTES));
                a2_tma_load_4d(sQ, &map_q, 0, q0, h, b, &bars.q_full);       // one 256-row box: tile A then tile B
                for (int t = 0; t < A2_STAGES && t < T; ++t) load_kv(t);
            }
            const uint32_t t_s = tmem_base + q * QT_COLS, t_p = t_s + 64, t_o = t_p + 32;
            const uint64_t qdesc = tc::make_desc(smem_u32(sQ + q * Q_BYTES), 512, 4);
            auto issue_qk = [&](int t) {
                const uint64_t kdesc = tc::make_desc(smem_u32(sKV + (t % A2_STAGES) * STAGE_BYTES), 512, 4);
                tc::mma_f16_ss(t_s, qdesc, kdesc, idesc_qk, 0u);
                tc::mma_f16_ss(t_s, qdesc + 2, kdesc + 2, idesc_qk, 1u);
                tc::mma_commit(&bars.s_full[q]);
            };
            auto issue_pv = [&](int t) {
                const uint32_t va = smem_u32(sKV + (t % A2_STAGES) * STAGE_BYTES + K_BYTES);
                const uint64_t vdesc = (DV == 32) ? tc::make_desc(va, 512, 4) : tc::make_desc(va, 1024, 2);
                constexpr uint32_t VSTEP = (16 * DV * 2) >> 4;               // 16 keys per MMA k-step, in 16-byte units
#pragma unroll
                for (int k = 0; k < A2_BKV / 16; ++k)
                    tc::mma_f16_ts(t_o, t_p + 8 * k, vdesc + VSTEP * k, idesc_pv, (t | k) ? 1u : 0u);   // 16 keys = 8 columns of P
                tc::mma_commit(&bars.pv_done[q]);
            };

            iwait(&bars.q_full, 0);
            iwait(&bars.kv_full[0], 0);
            issue_qk(0);
            for (int t = 0; t < T; ++t) {
                if (t + 1 < T) {
                    iwait(&bars.kv_full[(t + 1) % A2_STAGES], ((t + 1) / A2_STAGES) & 1);
                    iwait(&bars.s_free[q], t & 1);                   // every softmax warp of this tile holds S(t) in registers
                    tc::fence_after_sync();
                    issue_qk(t + 1);
                }
                iwait(&bars.p_full[q], t & 1);
                tc::fence_after_sync();
                issue_pv(t);
                tc::mma_commit(&bars.kv_empty[t % A2_STAGES]);               // this tile's MMAs on stage t%STAGES have been issued (one arrive per tile)
                // refill the stage of key tile t-lag once BOTH query tiles retired it.  lag = 1 keeps the ring full but parks tile A's issuer
                // behind tile B's PV(t-1); lag = 3 lets the two tiles drift three key tiles apart before either issuer waits on the other.
                if (q == 0 && t >= refill_lag && t - refill_lag + A2_STAGES < T) {
                    iwait(&bars.kv_empty[(t - refill_lag) % A2_STAGES], ((t - refill_lag) / A2_STAGES) & 1);
                    load_kv(t - refill_lag + A2_STAGES);
                }
            }
        }
    } else if (warp < 4 || has_b) {
        // ================================================= softmax warps: thread = query row ======================================
        const int qt = warp >> 2;                                            // 0 = tile A, 1 = tile B
        const int row = tid & 127;
        const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t t_s = tmem_base + qt * QT_COLS + lane_sel, t_p = t_s + 64, t_o = t_s + 96;
        uint64_t* bar_s_full = &bars.s_full[qt];
        uint64_t* bar_s_free = &bars.s_free[qt];
        uint64_t* bar_p_full = &bars.p_full[qt];
        uint64_t* bar_pv = &bars.pv_done[qt];
        float m_used = -INFINITY, l_run = 0.f;

        // Software-pipelined by one hand-off: iteration t first issues the tensor-memory read of S(t) (asynchronous), THEN completes the
        // hand-over of P(t-1) (wait for its store, fence, arrive) while that read is in flight, then consumes S(t).  One read site and one
        // store site per iteration: nothing is carried around the loop in registers (a variant that prefetched S(t+1) at the end of
        // iteration t carried 64 registers across the back edge and spilled them).
        for (int t = 0; t < T; ++t) {
            uint32_t sv[64];                                                  // S row of key tile t
            tc::mbar_wait(bar_s_full, t & 1);                                 // QK(t) was issued when s_free(t-1) arrived: long complete
            tc::fence_after_sync();
            tc::tmem_ld32(t_s, *reinterpret_cast<uint32_t (*)[32]>(&sv[0]));
            tc::tmem_ld32(t_s + 32, *reinterpret_cast<uint32_t (*)[32]>(&sv[32]));
            if (t > 0) {                                                      // P(t-1): its tcgen05.st was issued at the end of iteration t-1
                tc::tmem_st_wait();
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(bar_p_full);
            }
            tc::tmem_ld_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(bar_s_free);                      // S(t) is in registers: QK(t+1) may overwrite it
            const int kv0 = t * A2_BKV;
            if (kv0 + A2_BKV > N) {               // tail tile only: keys >= N do not exist
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (kv0 + i >= N) sv[i] = 0xff800000u;   // -inf
            }
            float mx;
            if (VAR & 1) {                        // four independent chains: 8 dependent FMNMX3 instead of 32 before the first exponential can issue
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int i = 0; i < 16; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) m4[c] = fmaxf(m4[c], __uint_as_float(sv[c * 16 + i]));
                mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            } else {
                mx = -INFINITY;
#pragma unroll
                for (int i = 0; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
            }
            const float m_tile = mx * scale_log2;
            const bool grow = m_tile > m_used + 8.f;         // also true on the first tile (m_used = -inf)
            float alpha = 1.f;
            if (grow) {
                alpha = a2_exp2(m_used - m_tile);            // 0 on the first tile
                m_used = m_tile;
            }
            uint32_t pk[32];
            float ls_row[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float x0 = fmaf(__uint_as_float(sv[2 * i]), scale_log2, -m_used);
                const float x1 = fmaf(__uint_as_float(sv[2 * i + 1]), scale_log2, -m_used);
                const float p0 = (POLY > 0 && (2 * i) % POLY == POLY - 1) ? a2_exp2_poly(x0) : a2_exp2(x0);
                const float p1 = (POLY > 0 && (2 * i + 1) % POLY == POLY - 1) ? a2_exp2_poly(x1) : a2_exp2(x1);
                pk[i] = pack_half2(p0, p1);
                ls_row[i & 3] += p0 + p1;                    // four independent partial sums
            }
            l_run = l_run * alpha + ((ls_row[0] + ls_row[1]) + (ls_row[2] + ls_row[3]));
            // ---- PV(t-1) must have retired before its P operand is overwritten or O is rescaled (issued a whole phase ago)
            if (t > 0) {
                tc::mbar_wait(bar_pv, (t - 1) & 1);
                tc::fence_after_sync();
                if (__any_sync(0xffffffffu, grow)) {         // warp-collective TMEM round trip, lanes that did not grow use 1.0
#pragma unroll
                    for (int c0 = 0; c0 < DV; c0 += 32) {
                        uint32_t o[32];
                        tc::tmem_ld32(t_o + c0, o);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tc::tmem_st32(t_o + c0, o);
                    }
                }
            }
            tc::tmem_st32(t_p, pk);                          // A operand of the TS-mode MMA: lane = query row, column c = keys 2c, 2c+1
        }
        tc::tmem_st_wait();                                  // hand over P(T-1)
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bar_p_full);
        // ---- finalise: O / l -> global
        tc::mbar_wait(bar_pv, (T - 1) & 1);
        tc::fence_after_sync();
        const float inv = 1.f / l_run;
        const int qrow = q0 + qt * A2_BQ + row;
        __half* orow = out + ((long long)b * N + qrow) * ldo + h * DV;
#pragma unroll
        for (int c0 = 0; c0 < DV; c0 += 32) {
            uint32_t o[32];
            tc::tmem_ld32(t_o + c0, o);
            tc::tmem_ld_wait();
            if (qrow < N) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    Half8 hv;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        hv.v[q] = __floats2half2_rn(__uint_as_float(o[c * 8 + 2 * q]) * inv, __uint_as_float(o[c * 8 + 2 * q + 1]) * inv);
                    *reinterpret_cast<Half8*>(orow + c0 + c * 8) = hv;
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

typedef CUresult (*A2EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static A2EncodeFn a2_encode() {
    static A2EncodeFn fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<A2EncodeFn>(q);
    }
    return fn;
}
// (d, token, head, image) view of one operand of the packed qkv buffer; box = d x rows x 1 x 1
static bool a2_map(CUtensorMap* m, const __half* base, int d, int N, int heads, int batch, int ld, int head_stride, int box_rows) {
    cuuint64_t gdim[4] = {(cuuint64_t)d, (cuuint64_t)N, (cuuint64_t)heads, (cuuint64_t)batch};
    cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)head_stride * 2, (cuuint64_t)N * ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)d, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t est[4] = {1, 1, 1, 1};
    return a2_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       d == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace ym

using namespace ym;

// Every n-th softmax exponential of the d_v = 32 kernel on the FMA pipe instead of the MUFU (0 = all on the MUFU; 4).
static int g_attention2_poly = 0;
extern "C" int ym_attention2_poly(void) { return g_attention2_poly; }
extern "C" int ym_set_attention2_poly(int every) {
    const int old = g_attention2_poly;
    if (every == 0 || every == 4) g_attention2_poly = every;
    return old;
}

// Kernel variant bits (measurement knob): bit 0 = row maximum as four independent chains, bit 1 = K / V ring refilled three key tiles
// behind the consumer instead of one, bit 2 = the issuing threads sleep between barrier polls.
static int g_attention2_variant = 0;
extern "C" int ym_set_attention2_variant(int bits) {
    const int old = g_attention2_variant;
    if (bits >= 0 && bits <= 7) g_attention2_variant = bits;
    return old;
}

// Query tiles per CTA: 0 = by wave fit (default), 1 / 2 = forced (tests exercise both code paths on small grids).
static int g_attention2_qtiles = 0;
extern "C" int ym_set_attention2_qtiles(int n) {
    const int old = g_attention2_qtiles;
    if (n >= 0 && n <= 2) g_attention2_qtiles = n;
    return old;
}

extern "C" int ym_attention_fwd_tc2_supported(int heads, int head_stride, int ld) {
    // tensor-map strides are multiples of 16 bytes; a single head needs no head stride
    return a2_encode() != nullptr && ld % 8 == 0 && (heads == 1 || head_stride % 8 == 0);
}

extern "C" int ym_attention_fwd_tc2(const void* qkv, int ld, int batch, int N, int heads, int head_stride, int q_off, int k_off,
                                    int v_off, int d_qk, int d_v, float scale, void* out, int ldo, void* stream) {
    YM_CHECK_ARG(qkv && out, "ym_attention_fwd_tc2: null pointer");
    YM_CHECK_ARG(d_qk == 32, "ym_attention_fwd_tc2: d_qk must be 32 (got %d)", d_qk);
    YM_CHECK_ARG(d_v == 32 || d_v == 64, "ym_attention_fwd_tc2: d_v must be 32 or 64 (got %d)", d_v);
    YM_CHECK_ARG(ld % 8 == 0 && head_stride % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0,
                 "ym_attention_fwd_tc2: offsets/pitch must be multiples of 8 halves");
    YM_CHECK_ARG(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0 && ldo % 8 == 0, "ym_attention_fwd_tc2: alignment");
    YM_CHECK_ARG(N > 0 && heads > 0 && batch >= 0 && batch < 65536, "ym_attention_fwd_tc2: bad sizes");
    YM_CHECK_ARG(a2_encode() != nullptr, "ym_attention_fwd_tc2: cuTensorMapEncodeTiled unavailable");
    if (batch == 0) return YM_OK;
    const __half* base = (const __half*)qkv;
    const int hs = heads == 1 ? 8 : head_stride;          // any legal stride for a size-1 dimension
    CUtensorMap mq, mk, mv;
    if (!a2_map(&mq, base + q_off, 32, N, heads, batch, ld, hs, 2 * A2_BQ) || !a2_map(&mk, base + k_off, 32, N, heads, batch, ld, hs, A2_BKV) ||
        !a2_map(&mv, base + v_off, d_v, N, heads, batch, ld, hs, A2_BKV)) {
        ym_set_error("ym_attention_fwd_tc2: cuTensorMapEncodeTiled failed (N=%d heads=%d batch=%d ld=%d head_stride=%d)", N, heads, batch, ld, hs);
        return YM_ERR_CUDA;
    }
    const float sl2 = scale * 1.4426950408889634f;
    // Two query tiles per CTA keep all eight softmax warps of a CTA busy and halve the K / V traffic; a one-tile CTA leaves half of its
    // threads and tensor memory idle, so it only pays when the two-tile grid would not even put one CTA on every SM (P5: 128 CTAs).
    // Measured at bs32 (profiles/r02_attention_qtiles.json): P4 77 us with two tiles against 97 us with one, P5 11.9 against 11.1 us.
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        sm_count = sms > 0 ? sms : 148;
    }
    const long long ctas2 = (long long)((N + 2 * A2_BQ - 1) / (2 * A2_BQ)) * heads * batch;
    int q_tiles = (d_v == 32 && ctas2 < sm_count) ? 1 : 2;
    if (g_attention2_qtiles) q_tiles = g_attention2_qtiles;
    dim3 grid((N + q_tiles * A2_BQ - 1) / (q_tiles * A2_BQ), heads, batch);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)2 * A2_BQ * 64 + A2_STAGES * (A2_BKV * 64 + A2_BKV * d_v * 2) + 1024;
    cudaError_t e = cudaSuccess;
#define A2_LAUNCH(DV_, POLY_, VAR_)                                                                                                        \
    do {                                                                                                                             \
        e = cudaFuncSetAttribute(tc_attention2_kernel<DV_, POLY_, VAR_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
        if (e == cudaSuccess) e = launch_pdl_prio(ym_kernel_priority() > 0 ? -ym_kernel_priority() : 0, tc_attention2_kernel<DV_, POLY_, VAR_>, grid, A2_THREADS, smem, st, mq, mk, mv, N, sl2, (__half*)out, ldo, q_tiles, lag); \
    } while (0)
    const int poly = ym_attention2_poly();
    const int tree = g_attention2_variant & 1, lag = g_attention2_variant & 6;
    if (d_v == 32) {
        if (poly == 4) { if (tree) A2_LAUNCH(32, 4, 1); else A2_LAUNCH(32, 4, 0); }
        else { if (tree) A2_LAUNCH(32, 0, 1); else A2_LAUNCH(32, 0, 0); }
    } else {
        if (tree) A2_LAUNCH(64, 0, 1); else A2_LAUNCH(64, 0, 0);
    }
#undef A2_LAUNCH
    if (e != cudaSuccess) { ym_set_error("ym_attention_fwd_tc2: smem attr: %s", cudaGetErrorString(e)); return YM_ERR_CUDA; }
    YM_CHECK_LAUNCH("tc_attention2");
    return YM_OK;
}
