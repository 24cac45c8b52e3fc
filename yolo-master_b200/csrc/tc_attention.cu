// tcgen05 / TMEM fused attention forward (sm_100a):  O = softmax((Q*scale) K^T) V  for d_qk = 32, d_v in {32, 64}.
//
// Same contract as attention.cu (see there for the reference mapping, block.py:1708-1722 / :1324-1331); this is the
// Blackwell-native data path:
//   * S = Q K^T and O += P V run on the 5th-gen tensor cores (tcgen05.mma kind::f16, M = 128 query rows), issued by ONE
//     thread; S (128 x 64 fp32) and O (128 x d_v fp32) live in tensor memory.
//   * softmax: thread t of the CTA owns query row t: it reads its whole S row with one tcgen05.ld, so row max / row sum
//     need no shuffles; exp2 on the MUFU; P is written as fp16 into a 128B-swizzled shared-memory tile that is the A
//     operand of the PV MMA.
//   * the softmax row sums are accumulated by the tensor core as well (L += P * ones, a 128x16 accumulator) from the very
//     fp16 P values that multiply V, so numerator and denominator are exactly consistent and no FADD is spent on them.
//   * O / L are only rescaled when a row's running max grows by more than 2^8 (lazy rescaling): most KV tiles never do.
//   * K/V tiles arrive by cp.async into 64B/128B-swizzled canonical UMMA layouts (3-stage ring).
// One CTA = 128 threads, 128 TMEM columns (S 64 + O <= 64) and ~48 KB smem -> 4 CTAs per SM overlap each other's
// MMA / softmax / load phases.
#include "tc_common.cuh"

namespace ym {

constexpr int TA_BQ = 128, TA_BKV = 64, TA_THREADS = 128, TA_STAGES = 3;

__device__ __forceinline__ float ta_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}

// exp2 on the FMA / ALU pipes: round-to-nearest range reduction by the 1.5*2^23 magic add, degree-3 minimax polynomial of 2^f
// on [-0.5, 0.5] (max relative error 7.5e-5, far below the fp16 rounding of P), exponent inserted by an integer add.
// Used for every POLY-th score so that the 16-lane/clk MUFU is no longer the only pipe the softmax waits on.
__device__ __forceinline__ float ta_exp2_poly(float x) {
    x = fmaxf(x, -125.f);
    const float t = x + 12582912.f;
    const float f = x - (t - 12582912.f);
    float p = fmaf(0.05517164617776871f, f, 0.2426111251115799f);
    p = fmaf(p, f, 0.6932609677314758f);
    p = fmaf(p, f, 0.9999280571937561f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// CHUNK: the S row is read from tensor memory in four 16-column chunks; the exponentials of chunk c run while chunk c+1 is in
// flight (tcgen05.ld is asynchronous until tcgen05.wait::ld), so the two 16-scores/clk/SM resources of this kernel - the
// tensor-memory read port and the MUFU - overlap inside every warp instead of alternating.  The running maximum is updated
// chunk by chunk with the same 2^8 slack as the lazy rescale; when a later chunk raises it, the fp16 P values already packed
// for this row are multiplied by the same factor as O / L, so the row stays exactly consistent.
// TSP: P never touches shared memory.  Each thread writes its packed fp16 P row into tensor memory (tcgen05.st, 256 B/clk) and
// O += P V runs as a TS-mode tcgen05.mma whose A operand is read from tensor memory; the row sums are plain fp32 adds in
// registers instead of a second (P x ones) MMA.  Per CTA-tile this removes the 16 KB P store, the 2 x 16 KB of P operand reads
// and the ones operand from the shared-memory pipe: 74 KB -> 24 KB per tile, where 128 B/clk of shared-memory bandwidth
// (580 clk per tile) had been a tighter ceiling than the MUFU and the tensor-memory read port (512 clk each).
template <int DV, int POLY, bool CHUNK, bool TSP>
__global__ void __launch_bounds__(TA_THREADS) tc_attention_kernel(const __half* __restrict__ qkv, int ld, int N, int head_stride,
                                                                 int q_off, int k_off, int v_off, float scale_log2,
                                                                 __half* __restrict__ out, int ldo) {
    constexpr int Q_BYTES = TA_BQ * 64, K_BYTES = TA_BKV * 64, V_BYTES = TA_BKV * DV * 2, P_BYTES = TA_BQ * 128;
    constexpr int ONES_BYTES = 16 * 128;             // B operand of the row-sum MMA: 16 x 64 fp16 ones
    // TMEM columns: S [0,64), O [64,64+DV), L (row sums, 16 cols) [64+DV, 80+DV)
    constexpr uint32_t TMEM_COLS = (DV == 32) ? 128 : 256;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (STS, not generic ST)
    unsigned char* sQ = smem;
    unsigned char* sP = sQ + Q_BYTES;
    unsigned char* sK = sP + P_BYTES;                      // [STAGES][K_BYTES]
    unsigned char* sV = sK + TA_STAGES * K_BYTES;          // [STAGES][V_BYTES]
    unsigned char* sOnes = sV + TA_STAGES * V_BYTES;
    __shared__ uint64_t bar_s, bar_pv;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * TA_BQ, h = blockIdx.y, b = blockIdx.z;
    const __half* base = qkv + (long long)b * N * ld + h * head_stride;
    const __half* gQ = base + q_off;
    const __half* gK = base + k_off;
    const __half* gV = base + v_off;

    if (tid == 0) {
        tc::mbar_init(&bar_s, 1);
        tc::mbar_init(&bar_pv, 1);
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, TMEM_COLS);
    {   // ones tile (layout-agnostic: every element is 1.0)
        const __half2 one2 = __float2half2_rn(1.f);
        uint4 v;
        v.x = v.y = v.z = v.w = *reinterpret_cast<const uint32_t*>(&one2);
        *reinterpret_cast<uint4*>(sOnes + tid * 16) = v;
    }

    // ---- Q tile: 128 rows x 64 B (SW64)
#pragma unroll
    for (int i = 0; i < TA_BQ * 4 / TA_THREADS; ++i) {
        const int idx = tid + i * TA_THREADS;
        const int r = idx >> 2, c = idx & 3;
        const bool ok = q0 + r < N;
        cp_async16(sQ + tc::sw64_offset(r, c), ok ? gQ + (long long)(q0 + r) * ld + c * 8 : gQ, ok ? 16 : 0);
    }
    // per-thread constant parts of the K/V tile copies (2 K chunks and DV/16 V chunks per thread per tile)
    constexpr int VCH = DV / 8, VIT = TA_BKV * VCH / TA_THREADS;
    int k_row[2], v_row[VIT];
    uint32_t k_dst[2], v_dst[VIT];
    long long k_src[2], v_src[VIT];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = tid + i * TA_THREADS;
        k_row[i] = idx >> 2;
        k_dst[i] = tc::sw64_offset(idx >> 2, idx & 3);
        k_src[i] = (long long)(idx >> 2) * ld + (idx & 3) * 8;
    }
#pragma unroll
    for (int i = 0; i < VIT; ++i) {
        const int idx = tid + i * TA_THREADS;
        const int r = idx / VCH, c = idx % VCH;
        v_row[i] = r;
        v_dst[i] = (DV == 32) ? tc::sw64_offset(r, c) : tc::sw128_offset(r, c);
        v_src[i] = (long long)r * ld + c * 8;
    }
    auto load_kv = [&](int t) {
        const int st = t % TA_STAGES, kv0 = t * TA_BKV;
        unsigned char* dK = sK + st * K_BYTES;
        unsigned char* dV = sV + st * V_BYTES;
        const __half* srcK = gK + (long long)kv0 * ld;
        const __half* srcV = gV + (long long)kv0 * ld;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const bool ok = kv0 + k_row[i] < N;
            cp_async16(dK + k_dst[i], ok ? srcK + k_src[i] : gK, ok ? 16 : 0);
        }
#pragma unroll
        for (int i = 0; i < VIT; ++i) {
            const bool ok = kv0 + v_row[i] < N;
            cp_async16(dV + v_dst[i], ok ? srcV + v_src[i] : gV, ok ? 16 : 0);
        }
    };
    const int T = (N + TA_BKV - 1) / TA_BKV;
    load_kv(0);
    cp_async_commit();
    if (T > 1) load_kv(1);
    cp_async_commit();
    cp_async_wait<1>();                       // Q and K_0 / V_0 have landed
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t t_s = tmem_base, t_o = tmem_base + 64, t_l = tmem_base + 64 + DV;
    const uint32_t t_p = tmem_base + 64 + DV;          // TSP: 32 columns of packed fp16 P (takes the place of L)
    float l_run = 0.f;                                 // TSP: running row sum
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;

    const uint32_t idesc_qk = tc::make_idesc_f16(TA_BQ, TA_BKV, 0);
    const uint32_t idesc_pv = tc::make_idesc_f16(TA_BQ, DV, 1);   // V is MN-major (d_v contiguous)
    const uint32_t idesc_l = tc::make_idesc_f16(TA_BQ, 16, 0);
    const uint64_t qdesc = tc::make_desc(smem_u32(sQ), 512, 4);
    const uint64_t pdesc = tc::make_desc(smem_u32(sP), 1024, 2);
    const uint64_t odesc = tc::make_desc(smem_u32(sOnes), 1024, 2);

    auto issue_qk = [&](int t) {
        const uint64_t kdesc = tc::make_desc(smem_u32(sK + (t % TA_STAGES) * K_BYTES), 512, 4);
        tc::mma_f16_ss(t_s, qdesc, kdesc, idesc_qk, 0u);
        tc::mma_f16_ss(t_s, qdesc + 2, kdesc + 2, idesc_qk, 1u);
        tc::mma_commit(&bar_s);
    };
    if (tid == 0) issue_qk(0);

    float m_used = -INFINITY;
    const int row = tid;                      // query row owned by this thread
    unsigned char* prow = sP + row * 128;

    for (int t = 0; t < T; ++t) {
        // ---- softmax of row `row` of S_t
        tc::mbar_wait(&bar_s, t & 1);
        tc::fence_after_sync();
        const int kv0 = t * TA_BKV;
        float alpha = 1.f;
        bool grow = false;
        uint32_t pk[32];
        float ls_row[4] = {0.f, 0.f, 0.f, 0.f};
        if constexpr (CHUNK) {
            uint32_t cur[16];
            tc::tmem_ld16(t_s + lane_sel, cur);
            tc::tmem_ld_wait();
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t nxt[16];
                if (ch < 3) tc::tmem_ld16(t_s + lane_sel + 16 * (ch + 1), nxt);      // lands while this chunk's exps run
                if (kv0 + TA_BKV > N) {           // tail tile only: keys >= N do not exist
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (kv0 + ch * 16 + i >= N) cur[i] = 0xff800000u;   // -inf
                }
                float mx = __uint_as_float(cur[0]);
#pragma unroll
                for (int i = 1; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(cur[i]));
                const float cm = mx * scale_log2;
                if (cm > m_used + 8.f) {          // first chunk of the first tile (m_used = -inf), then rare
                    const float a = ta_exp2(m_used - cm);
                    m_used = cm;
                    alpha *= a;
                    grow = true;
                    const __half2 a2 = __float2half2_rn(a);
#pragma unroll
                    for (int i = 0; i < ch * 8; ++i) {
                        const __half2 v = __hmul2(*reinterpret_cast<const __half2*>(&pk[i]), a2);
                        pk[i] = *reinterpret_cast<const uint32_t*>(&v);
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float p0 = ta_exp2(fmaf(__uint_as_float(cur[2 * i]), scale_log2, -m_used));
                    const float p1 = ta_exp2(fmaf(__uint_as_float(cur[2 * i + 1]), scale_log2, -m_used));
                    pk[ch * 8 + i] = pack_half2(p0, p1);
                }
                if (ch < 3) {
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
                }
            }
        } else {
            uint32_t sr[64];
            {
                uint32_t lo[32], hi[32];
                tc::tmem_ld32(t_s + lane_sel, lo);
                tc::tmem_ld32(t_s + lane_sel + 32, hi);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) { sr[i] = lo[i]; sr[32 + i] = hi[i]; }
            }
            if (kv0 + TA_BKV > N) {               // tail tile only: keys >= N do not exist
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (kv0 + i >= N) sr[i] = 0xff800000u;   // -inf
            }
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(sr[i]));
            const float m_tile = mx * scale_log2;
            grow = m_tile > m_used + 8.f;         // also true on the first tile (m_used = -inf)
            if (grow) {
                alpha = ta_exp2(m_used - m_tile);        // 0 on the first tile
                m_used = m_tile;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float x0 = fmaf(__uint_as_float(sr[2 * i]), scale_log2, -m_used);
                const float x1 = fmaf(__uint_as_float(sr[2 * i + 1]), scale_log2, -m_used);
                const float p0 = (POLY > 0 && (2 * i) % POLY == POLY - 1) ? ta_exp2_poly(x0) : ta_exp2(x0);
                const float p1 = (POLY > 0 && (2 * i + 1) % POLY == POLY - 1) ? ta_exp2_poly(x1) : ta_exp2(x1);
                pk[i] = pack_half2(p0, p1);
                if constexpr (TSP) ls_row[i & 3] += p0 + p1;   // four independent partial sums (no 64-deep dependent chain)
            }
        }
        if constexpr (TSP) {
            // Row sum for the final 1/l.  Whole-row schedule: summed in fp32 before the fp16 rounding of P (the rounding is
            // unbiased, |rel| <= 2^-11 per term, so numerator and denominator agree to ~1e-5); chunked schedule: summed from
            // the packed values, because an in-row rescale may have changed the earlier chunks.
            float ls = (ls_row[0] + ls_row[1]) + (ls_row[2] + ls_row[3]);
            if constexpr (CHUNK) {
                ls = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float2 f = unpack_half2(pk[i]);
                    ls += f.x + f.y;
                }
            }
            l_run = l_run * alpha + ls;
        }
        // ---- PV_{t-1} must be complete before O/L are rescaled, P is overwritten, or its K/V stage is reloaded
        if (t > 0) {
            tc::mbar_wait(&bar_pv, (t - 1) & 1);
            tc::fence_after_sync();
            if (__any_sync(0xffffffffu, grow)) {     // warp-collective TMEM round trip, lanes that did not grow use 1.0
#pragma unroll
                for (int c0 = 0; c0 < DV; c0 += 32) {
                    uint32_t o[32];
                    tc::tmem_ld32(t_o + lane_sel + c0, o);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tc::tmem_st32(t_o + lane_sel + c0, o);
                }
                if constexpr (!TSP) {
                    uint32_t l16[16];
                    tc::tmem_ld16(t_l + lane_sel, l16);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) l16[i] = __float_as_uint(__uint_as_float(l16[i]) * alpha);
                    tc::tmem_st16(t_l + lane_sel, l16);
                }
                tc::tmem_st_wait();
            }
        }
        // ---- prefetch tile t+2 into the ring stage last read by PV_{t-1}
        if (t + 2 < T) load_kv(t + 2);
        cp_async_commit();
        if constexpr (TSP) {
            // ---- P row -> tensor memory (A operand of the TS-mode MMA: lane = query row, column c = keys 2c, 2c+1)
            tc::tmem_st32(t_p + lane_sel, pk);
            tc::tmem_st_wait();
        } else {
            // ---- P row -> swizzled smem (A operand, K-major, 64 keys = 128 B per row)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint4 v = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                *reinterpret_cast<uint4*>(prow + ((c ^ (row & 7)) << 4)) = v;
            }
        }
        cp_async_wait<1>();                   // K_{t+1} / V_{t+1} have landed (only the newest group may be pending)
        tc::fence_proxy_async();
        tc::fence_before_sync();
        __syncthreads();
        // ---- S_{t+1} = Q K_{t+1}^T first (the next softmax waits on it), then O += P V_t and L += P 1
        if (tid == 0) {
            tc::fence_after_sync();
            if (t + 1 < T) issue_qk(t + 1);
            const uint32_t va = smem_u32(sV + (t % TA_STAGES) * V_BYTES);
            const uint64_t vdesc = (DV == 32) ? tc::make_desc(va, 512, 4) : tc::make_desc(va, 1024, 2);
            constexpr uint32_t VSTEP = (16 * DV * 2) >> 4;    // 16 keys per MMA k-step, in 16-byte units
#pragma unroll
            for (int k = 0; k < TA_BKV / 16; ++k) {
                if constexpr (TSP) {
                    tc::mma_f16_ts(t_o, t_p + 8 * k, vdesc + VSTEP * k, idesc_pv, (t | k) ? 1u : 0u);   // 16 keys = 8 columns
                } else {
                    tc::mma_f16_ss(t_o, pdesc + 2 * k, vdesc + VSTEP * k, idesc_pv, (t | k) ? 1u : 0u);
                    tc::mma_f16_ss(t_l, pdesc + 2 * k, odesc + 2 * k, idesc_l, (t | k) ? 1u : 0u);
                }
            }
            tc::mma_commit(&bar_pv);
        }
    }
    cp_async_wait<0>();
    // ---- finalise: O / l -> global
    tc::mbar_wait(&bar_pv, (T - 1) & 1);
    tc::fence_after_sync();
    float inv;
    if constexpr (TSP) {
        inv = 1.f / l_run;
    } else {
        uint32_t l16[16];
        tc::tmem_ld16(t_l + lane_sel, l16);
        tc::tmem_ld_wait();
        inv = 1.f / __uint_as_float(l16[0]);
    }
    const int qrow = q0 + row;
    __half* orow = out + ((long long)b * N + qrow) * ldo + h * DV;
#pragma unroll
    for (int c0 = 0; c0 < DV; c0 += 32) {
        uint32_t o[32];
        tc::tmem_ld32(t_o + lane_sel + c0, o);
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
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace ym

using namespace ym;

// Every POLY-th exponential of the softmax runs on the FMA pipe instead of the MUFU (0 = all on the MUFU; 2/3/4/6 supported).
static int g_attention_poly = 0;
extern "C" int ym_set_attention_poly(int every) {
    const int old = g_attention_poly;
    if (every == 0 || every == 2 || every == 3 || every == 4 || every == 6) g_attention_poly = every;
    return old;
}

// Softmax / PV schedule: 0 = whole-row softmax, P through shared memory (SS-mode MMAs, row sums on the tensor core);
// 1 = chunked softmax (tensor-memory reads overlapped with the exponentials inside each warp), P through shared memory;
// 2 = whole-row softmax, P through tensor memory (TS-mode PV MMA, row sums in registers); 3 = chunked + tensor-memory P.
static int g_attention_chunked = 2;   // measured on B200 (P3 shape, bs32): 1.005 / 1.070 / 0.922 / 1.058 ms for modes 0 / 1 / 2 / 3
extern "C" int ym_set_attention_chunked(int on) {
    const int old = g_attention_chunked;
    if (on >= 0 && on <= 3) g_attention_chunked = on;
    return old;
}

extern "C" int ym_attention_fwd_tc(const void* qkv, int ld, int batch, int N, int heads, int head_stride, int q_off, int k_off,
                                   int v_off, int d_qk, int d_v, float scale, void* out, int ldo, void* stream) {
    YM_CHECK_ARG(qkv && out, "ym_attention_fwd_tc: null pointer");
    YM_CHECK_ARG(d_qk == 32, "ym_attention_fwd_tc: d_qk must be 32 (got %d)", d_qk);
    YM_CHECK_ARG(d_v == 32 || d_v == 64, "ym_attention_fwd_tc: d_v must be 32 or 64 (got %d)", d_v);
    YM_CHECK_ARG(ld % 8 == 0 && head_stride % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0,
                 "ym_attention_fwd_tc: offsets/pitch must be multiples of 8 halves");
    YM_CHECK_ARG(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0 && ldo % 8 == 0, "ym_attention_fwd_tc: alignment");
    YM_CHECK_ARG(N > 0 && heads > 0 && batch >= 0 && batch < 65536, "ym_attention_fwd_tc: bad sizes");
    if (batch == 0) return YM_OK;
    const float sl2 = scale * 1.4426950408889634f;
    dim3 grid((N + TA_BQ - 1) / TA_BQ, heads, batch);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)TA_BQ * 64 + TA_BQ * 128 + TA_STAGES * (TA_BKV * 64 + TA_BKV * d_v * 2) + 16 * 128 + 1024;
    cudaError_t e = cudaSuccess;
#define TA_LAUNCH(DV_, POLY_, CH_, TS_)                                                                                            \
    do {                                                                                                                           \
        e = cudaFuncSetAttribute(tc_attention_kernel<DV_, POLY_, CH_, TS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e == cudaSuccess)                                                                                                      \
            tc_attention_kernel<DV_, POLY_, CH_, TS_><<<grid, TA_THREADS, smem, st>>>((const __half*)qkv, ld, N, head_stride, q_off,   \
                                                                                    k_off, v_off, sl2, (__half*)out, ldo);         \
    } while (0)
#define TA_LAUNCH_DV(DV_)                                                  \
    if (g_attention_chunked == 2) { TA_LAUNCH(DV_, 0, false, true); }      \
    else if (g_attention_chunked == 3) { TA_LAUNCH(DV_, 0, true, true); }  \
    else if (g_attention_chunked == 1) { TA_LAUNCH(DV_, 0, true, false); } \
    else switch (g_attention_poly) {                                       \
        case 2: TA_LAUNCH(DV_, 2, false, false); break;                    \
        case 3: TA_LAUNCH(DV_, 3, false, false); break;                    \
        case 4: TA_LAUNCH(DV_, 4, false, false); break;                    \
        case 6: TA_LAUNCH(DV_, 6, false, false); break;                    \
        default: TA_LAUNCH(DV_, 0, false, false); break;                   \
    }
    if (d_v == 32) { TA_LAUNCH_DV(32) } else { TA_LAUNCH_DV(64) }
#undef TA_LAUNCH_DV
#undef TA_LAUNCH
    if (e != cudaSuccess) { ym_set_error("ym_attention_fwd_tc: smem attr: %s", cudaGetErrorString(e)); return YM_ERR_CUDA; }
    YM_CHECK_LAUNCH("tc_attention");
    return YM_OK;
}
