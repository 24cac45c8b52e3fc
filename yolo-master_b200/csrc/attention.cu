// Fused area / PSA attention forward for NHWC fp16 (sm_100a):  O = softmax((Q*scale)^T K) V
//
// Replaces the reference's materialised N x N attention (block.py:1717-1719 AAttn, :1329-1331 Attention):
// QK^T, online softmax and PV stay on-chip, so per (image, head) the kernel reads Q,K,V once from the
// qkv conv output and writes O once (4*N*d*2 bytes) instead of the 2*N^2*2-byte score matrix.
//
// Layout: `qkv` is the NHWC output of the qkv 1x1 conv, rows = tokens, row pitch `ld` elements.  For head h the
// q/k/v channel runs start at h*head_stride + {q_off,k_off,v_off} (AAttn: [q32|k32|v32] per head, block.py:1713-1716;
// PSA Attention: [q kd|k kd|v hd] per head, block.py:1324-1326).  Output channel = h*DV + d (pitch ldo).
//
// CTA = 8 warps x 16 query rows (BQ=128); KV tiles of 64 double-buffered with cp.async;
// mma.sync.m16n8k16 fp16 -> fp32; softmax in fp32 with exp2.  d_qk is fixed at 32 (every attention on the path).
#include "ym_common.cuh"

namespace ym {

constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 64;
constexpr int ATT_DQK = 32;
constexpr int ATT_THREADS = 256;

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}

template <int DV>
__global__ void __launch_bounds__(ATT_THREADS) attention_fwd_kernel(const __half* __restrict__ qkv, int ld, int N,
                                                                     int head_stride, int q_off, int k_off, int v_off,
                                                                     float scale_log2, __half* __restrict__ out, int ldo) {
    constexpr int SQ = ATT_DQK + 8;  // smem pitch (halves) for Q/K rows
    constexpr int SV = DV + 8;
    __shared__ __align__(16) __half sQ[ATT_BQ * SQ];
    __shared__ __align__(16) __half sK[2][ATT_BKV * SQ];
    __shared__ __align__(16) __half sV[2][ATT_BKV * SV];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = blockIdx.x * ATT_BQ;
    const int h = blockIdx.y;
    const int b = blockIdx.z;
    const __half* base = qkv + (long long)b * N * ld + h * head_stride;
    const __half* gQ = base + q_off;
    const __half* gK = base + k_off;
    const __half* gV = base + v_off;

    // ---- load Q tile (128 rows x 4 chunks)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = tid + i * ATT_THREADS;
        const int r = idx >> 2, c = idx & 3;
        const bool ok = q0 + r < N;
        const __half* src = ok ? gQ + (long long)(q0 + r) * ld + c * 8 : gQ;
        cp_async16(&sQ[r * SQ + c * 8], src, ok ? 16 : 0);
    }
    auto load_kv = [&](int stage, int t) {
        const int kv0 = t * ATT_BKV;
        {
            const int r = tid >> 2, c = tid & 3;
            const bool ok = kv0 + r < N;
            const __half* src = ok ? gK + (long long)(kv0 + r) * ld + c * 8 : gK;
            cp_async16(&sK[stage][r * SQ + c * 8], src, ok ? 16 : 0);
        }
        constexpr int VCH = DV / 8;
#pragma unroll
        for (int i = 0; i < (ATT_BKV * VCH) / ATT_THREADS; ++i) {
            const int idx = tid + i * ATT_THREADS;
            const int r = idx / VCH, c = idx % VCH;
            const bool ok = kv0 + r < N;
            const __half* src = ok ? gV + (long long)(kv0 + r) * ld + c * 8 : gV;
            cp_async16(&sV[stage][r * SV + c * 8], src, ok ? 16 : 0);
        }
    };
    const int T = (N + ATT_BKV - 1) / ATT_BKV;
    load_kv(0, 0);
    cp_async_commit();

    constexpr int ONT = DV / 8;
    float o_acc[ONT][4];
#pragma unroll
    for (int i = 0; i < ONT; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) o_acc[i][q] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    uint32_t qf[2][4];

    for (int t = 0; t < T; ++t) {
        if (t + 1 < T) load_kv((t + 1) & 1, t + 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        if (t == 0) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const __half* pq = &sQ[(warp * 16 + (lane & 15)) * SQ + ks * 16 + (lane >> 4) * 8];
                ldmatrix_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], pq);
            }
        }
        const __half* tK = sK[t & 1];
        const __half* tV = sV[t & 1];
        // ---- S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) s[i][q] = 0.f;
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t b0, b1, b2, b3;
                const __half* pk = tK + (nj * 16 + (lane & 7) + (lane >> 4) * 8) * SQ + ks * 16 + ((lane >> 3) & 1) * 8;
                ldmatrix_x4(b0, b1, b2, b3, pk);
                mma_16816(s[2 * nj], qf[ks], b0, b1);
                mma_16816(s[2 * nj + 1], qf[ks], b2, b3);
            }
        }
        // ---- mask tail keys
        const int kv0 = t * ATT_BKV;
        if (kv0 + ATT_BKV > N) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = kv0 + i * 8 + 2 * (lane & 3);
                if (c >= N) { s[i][0] = -INFINITY; s[i][2] = -INFINITY; }
                if (c + 1 >= N) { s[i][1] = -INFINITY; s[i][3] = -INFINITY; }
            }
        }
        // ---- online softmax (rows g and g+8)
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            mx[0] = fmaxf(mx[0], fmaxf(s[i][0], s[i][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[i][2], s[i][3]));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        }
        float alpha[2], mnew[2], rs[2] = {0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mnew[r] = fmaxf(m_run[r], mx[r] * scale_log2);
            alpha[r] = fast_exp2(m_run[r] - mnew[r]);
            m_run[r] = mnew[r];
        }
        uint32_t pf[4][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float p0 = fast_exp2(fmaf(s[i][0], scale_log2, -mnew[0]));
            const float p1 = fast_exp2(fmaf(s[i][1], scale_log2, -mnew[0]));
            const float p2 = fast_exp2(fmaf(s[i][2], scale_log2, -mnew[1]));
            const float p3 = fast_exp2(fmaf(s[i][3], scale_log2, -mnew[1]));
            rs[0] += p0 + p1;
            rs[1] += p2 + p3;
            pf[i >> 1][(i & 1) * 2 + 0] = pack_half2(p0, p1);
            pf[i >> 1][(i & 1) * 2 + 1] = pack_half2(p2, p3);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * alpha[r] + rs[r];
#pragma unroll
        for (int i = 0; i < ONT; ++i) {
            o_acc[i][0] *= alpha[0];
            o_acc[i][1] *= alpha[0];
            o_acc[i][2] *= alpha[1];
            o_acc[i][3] *= alpha[1];
        }
        // ---- O += P V
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int jn = 0; jn < ONT / 2; ++jn) {
                uint32_t b0, b1, b2, b3;
                const __half* pv = tV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * SV + jn * 16 + (lane >> 4) * 8;
                ldmatrix_x4_trans(b0, b1, b2, b3, pv);
                mma_16816(o_acc[2 * jn], pf[kk], b0, b1);
                mma_16816(o_acc[2 * jn + 1], pf[kk], b2, b3);
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();

    // ---- finalise: row sums across the quad, normalise, store
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
    const int g = lane >> 2, tq = lane & 3;
    const int row0 = q0 + warp * 16 + g;
    __half* ob = out + (long long)b * N * ldo + h * DV;
#pragma unroll
    for (int i = 0; i < ONT; ++i) {
        const int c = i * 8 + 2 * tq;
        if (row0 < N)
            *reinterpret_cast<__half2*>(ob + (long long)row0 * ldo + c) = __floats2half2_rn(o_acc[i][0] * inv0, o_acc[i][1] * inv0);
        if (row0 + 8 < N)
            *reinterpret_cast<__half2*>(ob + (long long)(row0 + 8) * ldo + c) =
                __floats2half2_rn(o_acc[i][2] * inv1, o_acc[i][3] * inv1);
    }
}

}  // namespace ym

using namespace ym;

extern "C" int ym_attention_fwd_tc(const void* qkv, int ld, int batch, int N, int heads, int head_stride, int q_off, int k_off,
                                   int v_off, int d_qk, int d_v, float scale, void* out, int ldo, void* stream);
extern "C" int ym_attention_fwd_tc2(const void* qkv, int ld, int batch, int N, int heads, int head_stride, int q_off, int k_off,
                                    int v_off, int d_qk, int d_v, float scale, void* out, int ldo, void* stream);
extern "C" int ym_attention_fwd_tc2_supported(int heads, int head_stride, int ld);
// 2 = warp-specialised tcgen05 / TMA kernel (tc_attention2.cu, default), 1 = one-tile-per-CTA tcgen05 kernel (tc_attention.cu),
// 0 = mma.sync kernel (this file).  The older kernels stay selectable as A/B baselines of the same contract.
static int g_attention_impl = 2;
extern "C" int ym_set_attention_impl(int impl) {
    const int old = g_attention_impl;
    if (impl >= 0 && impl <= 2) g_attention_impl = impl;
    return old;
}

// Batched multi-head attention over token rows.  batch = images * areas, N = tokens per (image, area).
extern "C" int ym_attention_fwd(const void* qkv, int ld, int batch, int N, int heads, int head_stride, int q_off, int k_off,
                                int v_off, int d_qk, int d_v, float scale, void* out, int ldo, void* stream) {
    YM_CHECK_ARG(qkv && out, "ym_attention_fwd: null pointer");
    YM_CHECK_ARG(d_qk == ATT_DQK, "ym_attention_fwd: d_qk must be 32 (got %d)", d_qk);
    YM_CHECK_ARG(d_v == 32 || d_v == 64, "ym_attention_fwd: d_v must be 32 or 64 (got %d)", d_v);
    YM_CHECK_ARG(ld % 8 == 0 && head_stride % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0,
                 "ym_attention_fwd: offsets/pitch must be multiples of 8 halves");
    YM_CHECK_ARG(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 3) == 0 && ldo % 2 == 0, "ym_attention_fwd: alignment");
    YM_CHECK_ARG(N > 0 && heads > 0 && batch >= 0 && batch < 65536, "ym_attention_fwd: bad sizes");
    if (batch == 0) return YM_OK;
    if (g_attention_impl == 2 && ((uintptr_t)out & 15) == 0 && ldo % 8 == 0 && ym_attention_fwd_tc2_supported(heads, head_stride, ld))
        return ym_attention_fwd_tc2(qkv, ld, batch, N, heads, head_stride, q_off, k_off, v_off, d_qk, d_v, scale, out, ldo, stream);
    if (g_attention_impl >= 1 && ((uintptr_t)out & 15) == 0 && ldo % 8 == 0)
        return ym_attention_fwd_tc(qkv, ld, batch, N, heads, head_stride, q_off, k_off, v_off, d_qk, d_v, scale, out, ldo, stream);
    const float sl2 = scale * 1.4426950408889634f;
    dim3 grid((N + ATT_BQ - 1) / ATT_BQ, heads, batch);
    if (d_v == 32)
        attention_fwd_kernel<32><<<grid, ATT_THREADS, 0, (cudaStream_t)stream>>>(
            (const __half*)qkv, ld, N, head_stride, q_off, k_off, v_off, sl2, (__half*)out, ldo);
    else
        attention_fwd_kernel<64><<<grid, ATT_THREADS, 0, (cudaStream_t)stream>>>(
            (const __half*)qkv, ld, N, head_stride, q_off, k_off, v_off, sl2, (__half*)out, ldo);
    YM_CHECK_LAUNCH("attention_fwd");
    return YM_OK;
}
