// tcgen05 GEMM for 1x1 convolutions / routed expert GEMMs:  out[M,N] = act(A[M,K] * B[N,K]^T + bias) (+ res)
//
// Blackwell-native data path: operand tiles are staged in shared memory in the 128-byte-swizzled K-major layout,
// ONE elected thread issues tcgen05.mma (kind::f16, M=128, N=BN) with the fp32 accumulator in tensor memory,
// completion is tracked with tcgen05.commit -> mbarrier, and the epilogue reads the accumulator back with tcgen05.ld
// (warp w <-> TMEM lanes 32w..32w+31, one output row per thread).
// Grouped mode (MoE dispatch): blockIdx.z = problem, weights selected through the router's index table, up to two
// routed experts accumulated per tile into two TMEM accumulators and combined with the routing weights in the epilogue.
#include "tc_common.cuh"

namespace ym {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;       // one 128-byte swizzle atom of fp16 along K
constexpr int TC_STAGES = 3;
constexpr int TC_THREADS = 128;

struct TcGemmParams {
    const __half* a; int lda;
    const __half* b; int ldb;           // [N][ldb] K-major weights
    const float* bias;
    const __half* res; int ldr;
    __half* out; int ldo;
    int M, N, K, act;
    // grouped dispatch (optional): problem z covers rows [z*rows_per_prob, (z+1)*rows_per_prob)
    const int* route_idx;      // [P*topk] expert ids
    const float* route_w;      // [P*topk] weights
    int topk; int rows_per_prob; long long b_expert_stride; float w_min; float clamp;
};

// load a [ROWS x 64] fp16 tile (row pitch ld) into the swizzled smem tile; rows >= nrows and k >= K are zero-filled
template <int ROWS>
__device__ __forceinline__ void load_tile_sw128(unsigned char* smem_tile, const __half* g, int ld, int row0, int nrows, int k0,
                                                int K, int tid) {
#pragma unroll
    for (int i = 0; i < ROWS * 8 / TC_THREADS; ++i) {
        const int idx = tid + i * TC_THREADS;
        const int r = idx >> 3, c = idx & 7;
        const int k = k0 + c * 8;
        const bool ok = (row0 + r) < nrows && k < K;
        const __half* src = ok ? g + (long long)(row0 + r) * ld + k : g;
        cp_async16(smem_tile + tc::sw128_offset(r, c), src, ok ? 16 : 0);
    }
}

template <int BN, bool GROUPED>
__global__ void __launch_bounds__(TC_THREADS) tc_gemm_kernel(const TcGemmParams p) {
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte aligned operand ring
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (STS, not generic ST)
    constexpr int A_BYTES = TC_BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
    __shared__ uint64_t mma_bar[TC_STAGES];
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NACC = GROUPED ? 2 : 1;
    constexpr uint32_t TMEM_COLS = (BN * NACC <= 32) ? 32 : (BN * NACC <= 64) ? 64 : (BN * NACC <= 128) ? 128 : (BN * NACC <= 256) ? 256 : 512;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) tc::mbar_init(&mma_bar[s], 1);
        tc::mbar_init(&done_bar, 1);
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_slot;

    const int n0 = blockIdx.y * BN;
    int m0 = blockIdx.x * TC_BM;
    int m_end = p.M;
    const __half* bptr[NACC];
    float rw[NACC];
    int nacc_live = 1;
    bptr[0] = p.b;
    rw[0] = 1.f;
    if (GROUPED) {
        const int z = blockIdx.z;
        m0 += z * p.rows_per_prob;
        m_end = (z + 1) * p.rows_per_prob;
        nacc_live = 0;
        for (int j = 0; j < NACC; ++j) {  // routes with weight <= w_min are dropped (moe/utils.py:172-173)
            bptr[j] = p.b;
            rw[j] = 0.f;
            if (j < p.topk) {
                const float w = p.route_w[z * p.topk + j];
                if (w > p.w_min) {
                    bptr[nacc_live] = p.b + (long long)p.route_idx[z * p.topk + j] * p.b_expert_stride;
                    rw[nacc_live] = w;
                    ++nacc_live;
                }
            }
        }
    }
    const int KT = (p.K + TC_BK - 1) / TC_BK;
    const int T = KT * (nacc_live > 0 ? nacc_live : 0);   // pipeline steps: (accumulator, k tile)
    const uint32_t idesc = tc::make_idesc_f16(TC_BM, BN);

    auto issue_load = [&](int step) {
        const int acc = step / KT, kt = step - acc * KT;
        unsigned char* st = smem + (step % TC_STAGES) * STAGE_BYTES;
        load_tile_sw128<TC_BM>(st, p.a, p.lda, m0, m_end, kt * TC_BK, p.K, tid);
        load_tile_sw128<BN>(st + A_BYTES, bptr[acc], p.ldb, n0, p.N, kt * TC_BK, p.K, tid);
    };

#pragma unroll
    for (int s = 0; s < TC_STAGES - 1; ++s) {
        if (s < T) issue_load(s);
        cp_async_commit();
    }
    for (int step = 0; step < T; ++step) {
        cp_async_wait<TC_STAGES - 2>();
        tc::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            const int acc = step / KT, kt = step - acc * KT;
            const uint32_t sa = smem_u32(smem + (step % TC_STAGES) * STAGE_BYTES);
            const uint64_t adesc = tc::make_desc_sw128(sa), bdesc = tc::make_desc_sw128(sa + A_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)   // advance 32 bytes (16 fp16) along K inside the swizzle atom
                tc::mma_f16_ss(tmem_base + acc * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (kt | k) ? 1u : 0u);
            tc::mma_commit(&mma_bar[step % TC_STAGES]);
            if (step == T - 1) tc::mma_commit(&done_bar);
        }
        const int nxt = step + TC_STAGES - 1;
        if (nxt < T) {
            if (nxt >= TC_STAGES) tc::mbar_wait(&mma_bar[nxt % TC_STAGES], ((nxt - TC_STAGES) / TC_STAGES) & 1);
            issue_load(nxt);
        }
        cp_async_commit();
    }
    cp_async_wait<0>();

    // ---------------- epilogue: TMEM -> registers -> global (one output row per thread)
    if (T > 0) {
        tc::mbar_wait(&done_bar, 0);
        tc::fence_after_sync();
    }
    const int m = m0 + warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = 0.f;
        for (int acc = 0; acc < nacc_live; ++acc) {
            uint32_t r[16];
            tc::tmem_ld16(lane_addr + acc * BN + c0, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                float x = __uint_as_float(r[q]);
                if (GROUPED) x = __half2float(__float2half_rn(__half2float(__float2half_rn(x)) * rw[acc]));  // expert out fp16, *w, fp16 (utils.py:202-203)
                v[q] += x;
            }
        }
        const int n = n0 + c0;
        if (m < m_end && n < p.N) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                float x = v[q];
                if (p.bias != nullptr && n + q < p.N) x += p.bias[n + q];
                if (p.act == 1) x = silu_f(x);
                if (GROUPED) x = fminf(fmaxf(x, -p.clamp), p.clamp);
                v[q] = x;
            }
            if (n + 16 <= p.N) {
                if (p.res != nullptr) {
                    const Half8 r0 = *reinterpret_cast<const Half8*>(p.res + (long long)m * p.ldr + n);
                    const Half8 r1 = *reinterpret_cast<const Half8*>(p.res + (long long)m * p.ldr + n + 8);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 f0 = __half22float2(r0.v[q]), f1 = __half22float2(r1.v[q]);
                        v[2 * q] += f0.x; v[2 * q + 1] += f0.y; v[8 + 2 * q] += f1.x; v[8 + 2 * q + 1] += f1.y;
                    }
                }
                Half8 o0, o1;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    o0.v[q] = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
                    o1.v[q] = __floats2half2_rn(v[8 + 2 * q], v[8 + 2 * q + 1]);
                }
                *reinterpret_cast<Half8*>(p.out + (long long)m * p.ldo + n) = o0;
                *reinterpret_cast<Half8*>(p.out + (long long)m * p.ldo + n + 8) = o1;
            } else {
                for (int q = 0; q < 16 && n + q < p.N; ++q) {
                    float x = v[q];
                    if (p.res != nullptr) x += __half2float(p.res[(long long)m * p.ldr + n + q]);
                    p.out[(long long)m * p.ldo + n + q] = __float2half_rn(x);
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, bool GROUPED>
static int launch_tc(const TcGemmParams& p, dim3 grid, cudaStream_t st) {
    const size_t smem = (size_t)TC_STAGES * (TC_BM * 128 + BN * 128) + 1024;
    auto kern = tc_gemm_kernel<BN, GROUPED>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ym_set_error("tc_gemm: smem attr %zu: %s", smem, cudaGetErrorString(e)); return YM_ERR_CUDA; }
    kern<<<grid, TC_THREADS, smem, st>>>(p);
    YM_CHECK_LAUNCH("tc_gemm");
    return YM_OK;
}

}  // namespace ym

using namespace ym;

// out[M,N] = act(A[M,K] B[N,K]^T + bias) (+res): the 1x1 Conv (+folded BN, SiLU, residual) of conv.py:69-89 on tcgen05.
extern "C" int ym_tc_gemm_nt(const void* a, int lda, const void* b, int ldb, const float* bias, const void* res, int ldr,
                             void* out, int ldo, int M, int N, int K, int act, void* stream) {
    YM_CHECK_ARG(a && b && out, "ym_tc_gemm_nt: null pointer");
    YM_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldo % 8 == 0 && N % 8 == 0, "ym_tc_gemm_nt: K, N and pitches must be multiples of 8");
    YM_CHECK_ARG((((uintptr_t)a | (uintptr_t)b | (uintptr_t)out | (uintptr_t)res) & 15) == 0, "ym_tc_gemm_nt: 16-byte alignment");
    YM_CHECK_ARG(res == nullptr || ldr % 8 == 0, "ym_tc_gemm_nt: residual pitch");
    if (M == 0) return YM_OK;
    TcGemmParams p;
    memset(&p, 0, sizeof(p));
    p.a = (const __half*)a; p.lda = lda; p.b = (const __half*)b; p.ldb = ldb; p.bias = bias;
    p.res = (const __half*)res; p.ldr = ldr; p.out = (__half*)out; p.ldo = ldo; p.M = M; p.N = N; p.K = K; p.act = act;
    cudaStream_t st = (cudaStream_t)stream;
    const int mt = (M + TC_BM - 1) / TC_BM;
    if (N <= 64) return launch_tc<64, false>(p, dim3(mt, 1, 1), st);
    if (N <= 128) return launch_tc<128, false>(p, dim3(mt, 1, 1), st);
    return launch_tc<256, false>(p, dim3(mt, (N + 255) / 256, 1), st);
}

// ES-MoE dispatch (BatchedExpertComputation.compute_sparse_experts_batched, moe/utils.py:119-209) with 1x1-conv experts:
// out[b] = clamp(sum_j fp16(fp16(x[b] W[e_bj]^T) * w_bj), +-clamp), routes with w <= w_min dropped, no gather/scatter copies:
// each 128-token tile of image b is read once and multiplied by its (<= 2) routed experts' weights in place.
extern "C" int ym_moe_dispatch_tc(const void* x, int ldx, int B, int HW, int C, const void* w_all, int ldw,
                                  long long w_expert_stride, const int* route_idx, const float* route_w, int topk, int N,
                                  float w_min, float clamp, void* out, int ldo, void* stream) {
    YM_CHECK_ARG(x && w_all && route_idx && route_w && out, "ym_moe_dispatch_tc: null pointer");
    YM_CHECK_ARG(topk >= 1 && topk <= 2, "ym_moe_dispatch_tc: top_k must be 1 or 2 (got %d)", topk);
    YM_CHECK_ARG(C % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0 && ldo % 8 == 0 && N % 16 == 0, "ym_moe_dispatch_tc: multiples of 8 / 16");
    YM_CHECK_ARG(N <= 256, "ym_moe_dispatch_tc: N (%d) must be <= 256 (two fp32 accumulators fill the 512 TMEM columns)", N);
    YM_CHECK_ARG((((uintptr_t)x | (uintptr_t)w_all | (uintptr_t)out) & 15) == 0, "ym_moe_dispatch_tc: 16-byte alignment");
    YM_CHECK_ARG(B >= 0 && B < 65536, "ym_moe_dispatch_tc: batch");
    if (B == 0) return YM_OK;
    TcGemmParams p;
    memset(&p, 0, sizeof(p));
    p.a = (const __half*)x; p.lda = ldx; p.b = (const __half*)w_all; p.ldb = ldw; p.out = (__half*)out; p.ldo = ldo;
    p.M = B * HW; p.N = N; p.K = C; p.act = 0;
    p.route_idx = route_idx; p.route_w = route_w; p.topk = topk; p.rows_per_prob = HW; p.b_expert_stride = w_expert_stride;
    p.w_min = w_min; p.clamp = clamp;
    cudaStream_t st = (cudaStream_t)stream;
    const int mt = (HW + TC_BM - 1) / TC_BM;
    if (N <= 64) return launch_tc<64, true>(p, dim3(mt, 1, B), st);
    if (N <= 128) return launch_tc<128, true>(p, dim3(mt, 1, B), st);
    return launch_tc<256, true>(p, dim3(mt, 1, B), st);
}
