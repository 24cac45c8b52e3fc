// ES-MoE dispatch, CTA-pair version (BatchedExpertComputation.compute_sparse_experts_batched with 1x1-conv experts,
// moe/utils.py:119-209):   out[b] = clamp( sum_j fp16( fp16(x[b] W[e_bj]^T) * w_bj ), +-clamp ),  w <= w_min dropped.
//
// Same contract and numerics as tc_dispatch.cu.  What limits that kernel is shared-memory bandwidth: an SS-mode
// tcgen05.mma of shape 128x128x16 reads 8 KB of operands in its 64 cycles, i.e. the whole 128 B/clk of the SM, so every
// TMA write and every staging byte stalls the tensor pipe (ncu: HMMA pipe busy 72 % of the time at 26 % of its rate).
// This kernel cuts the shared-memory bytes per MMA flop by 2.3x:
//   * two SMs of a TPC form a cluster; each CTA keeps its own 128-token x tile (double-buffered) and loads only HALF of
//     every expert weight tile; one tcgen05.mma.cta_group::2 of shape 256 x N x 16 (N = all output channels, <= 256)
//     multiplies both x tiles by the union of the halves: per CTA 4 KB (x) + 4 KB (W half) per 128 cycles;
//   * the weight traffic L2 -> shared memory per token is halved;
//   * finished 32-channel groups leave during the last expert pass through a 2-slot ring of 16 KB staging tiles and TMA
//     stores, so the (slow) write path drains behind the TMEM reads.
// One accumulator (N <= 256 columns) per routed expert; the two TMEM slots form a ring over expert passes and the epilogue
// keeps the weighted fp16 partial sum of the first expert in registers while the second one is being multiplied.
//   warp 0: TMA producer (both CTAs) | warp 1: MMA issuer (leader CTA only) | warps 2..9: epilogue (both CTAs)
// Barriers: a_full / b_full / t_empty live in the leader (the peer's TMA loads and epilogue arrive there remotely);
//           a_empty / b_empty / t_full are signalled in BOTH CTAs by multicast tcgen05.commit.
#include <cuda.h>

#include "tc_common.cuh"

namespace ym {

constexpr int DQ_THREADS = 320, DQ_BM = 128, DQ_KC = 64, DQ_BSTAGES = 4, DQ_MAXU = 32;

struct Dispatch2Params {
    const int* route_idx; const float* route_w;
    int topk, HW, upi, unit0, units, K, N;   // upi = 256-token units per image; this launch covers units [unit0, unit0 + units)
    float w_min, clamp;
    int debug;
    long long* trace;   // profiling only (ym_set_dispatch_trace): clock64 timeline of CTA 0 / 1, [cta][role][256]
};

__device__ __forceinline__ uint32_t dq_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t dq_clusterid() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;\n" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t dq_nclusters() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;\n" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t dq_mapa(uint32_t addr, uint32_t rank) {
    uint32_t d;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(d) : "r"(addr), "r"(rank));
    return d;
}
__device__ __forceinline__ void dq_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// TMA tile load into THIS CTA's shared memory; the transaction bytes are credited to `bar_cluster_addr` (a shared::cluster
// address, i.e. the leader's barrier).
__device__ __forceinline__ void dq_tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void dq_tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void dq_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool dq_mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void dq_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void dq_mma2_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once all MMAs issued so far have completed
__device__ __forceinline__ void dq_commit2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

#define DQ_TR(role, idx)                                                                                                  \
    do {                                                                                                                \
        if (p.trace && blockIdx.x < 2 && (idx) < 256) p.trace[blockIdx.x * 1024 + (role) * 256 + (idx)] = clock64();   \
    } while (0)

template <int NI>   // MMA N = number of output channels (128 or 256); each CTA supplies NI / 2 weight rows per stage
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DQ_THREADS, 1)
    tc_dispatch2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                        const __grid_constant__ CUtensorMap map_o, const Dispatch2Params p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (STS, not generic ST)
    constexpr int A_CHUNK = DQ_BM * 128, A_TILE = 4 * A_CHUNK, B_SLOT = 128 * 128, B_BYTES = (NI / 2) * 128, NCT = NI / 2;
    const int kchunks = p.K / DQ_KC;                       // <= 4
    unsigned char* sA = smem;                              // [2 tiles][4][A_CHUNK]  128 KB
    unsigned char* sB = sA + 2 * A_TILE;                   // [DQ_BSTAGES][B_SLOT]    64 KB (this CTA's NI/2 output channels x 64 k per stage)
    unsigned char* stg = sB + DQ_BSTAGES * B_SLOT;         // [2 slots][2 halves][128 rows x 64 B] 32 KB output staging ring (32 channels per thread per step)
    __shared__ uint64_t a_full[2], a_empty[2], b_full[DQ_BSTAGES], b_empty[DQ_BSTAGES], t_full[2], t_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ int s_nr[DQ_MAXU], s_e[DQ_MAXU][2];
    __shared__ float s_w[DQ_MAXU][2];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) DQ_TR(3, 0);
    const uint32_t rank = dq_ctarank(), cid = dq_clusterid(), ncl = dq_nclusters();
    const int my_units = ((int)cid < p.units) ? (p.units - 1 - (int)cid) / (int)ncl + 1 : 0;   // units cid, cid + ncl, ... (<= DQ_MAXU)

    if (tid == 0) {
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&a_full[a], 1); tc::mbar_init(&a_empty[a], 1); }
        for (int s = 0; s < DQ_BSTAGES; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&t_full[a], 1); tc::mbar_init(&t_empty[a], 16); }   // 8 epilogue warps x 2 CTAs
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
    }
    tc::fence_before_sync();
    __syncthreads();
    dq_cluster_sync();                                      // both CTAs' barriers are initialised before any remote arrive / TMA credit
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    // barriers / tensor memory are set up: let the next kernel of the stream stage itself, then wait for the producers of x and of
    // the routing table (programmatic dependent launch; both are no-ops for an ordinary launch)
    pdl_prologue();
    // routes of every unit this pair will process, read once (the loops below never touch global memory for them)
    for (int i = tid; i < my_units; i += DQ_THREADS) {
        const int b = (p.unit0 + (int)cid + i * (int)ncl) / p.upi;
        int n = 0;
        for (int j = 0; j < p.topk && j < 2; ++j) {
            const float wj = p.route_w[b * p.topk + j];
            if (wj > p.w_min) { s_e[i][n] = p.route_idx[b * p.topk + j]; s_w[i][n] = wj; ++n; }
        }
        s_nr[i] = n;
    }
    __syncthreads();
    if (tid == 0) DQ_TR(3, 1);
    int tr = 0;                                             // per-role trace cursor

    if (warp == 0) {
        if (lane == 0) {
            // it counts PROCESSED units: x buffer = it & 1, phase of a_full / a_empty = (it >> 1) & 1
            uint32_t it = 0, bidx = 0;
            uint32_t a_full_l[2], b_full_l[DQ_BSTAGES];
            for (int a = 0; a < 2; ++a) a_full_l[a] = dq_mapa(smem_u32(&a_full[a]), 0);
#pragma unroll
            for (int s = 0; s < DQ_BSTAGES; ++s) b_full_l[s] = dq_mapa(smem_u32(&b_full[s]), 0);
            auto next_live = [&](int i) { while (i < my_units && s_nr[i] == 0) ++i; return i; };
            auto load_x = [&](int i, uint32_t n) {
                if (p.debug & 2) return;
                const int u = p.unit0 + (int)cid + i * (int)ncl, b = u / p.upi;
                const int m0 = b * p.HW + (u - b * p.upi) * (2 * DQ_BM) + (int)rank * DQ_BM;
                const uint32_t ab = n & 1;
                tc::mbar_wait(&a_empty[ab], ((n >> 1) & 1) ^ 1);
                DQ_TR(0, tr); ++tr;
                if (rank == 0) dq_expect_tx(&a_full[ab], (uint32_t)(2 * kchunks * A_CHUNK));
                for (int kc = 0; kc < kchunks; ++kc) dq_tma_load_2d(sA + ab * A_TILE + kc * A_CHUNK, &map_x, kc * DQ_KC, m0, a_full_l[ab]);
            };
            int i = next_live(0);
            if (i < my_units) load_x(i, 0);
            while (i < my_units) {
                const int nr = s_nr[i];
                const int nxt = next_live(i + 1);
                bool next_issued = nxt >= my_units;
                int local = 0;
                for (int j = 0; j < nr; ++j)
                    for (int kc = 0; kc < kchunks; ++kc, ++bidx, ++local) {
                        // prefetch the next unit's x tile as soon as the MMAs of the previous unit have released its buffer (polled, so
                        // the weight stream never blocks on it)
                        if (!next_issued && dq_mbar_test(&a_empty[(it + 1) & 1], (((it + 1) >> 1) & 1) ^ 1)) { load_x(nxt, it + 1); next_issued = true; }
                        if (p.debug & 1) continue;
                        const int s = bidx % DQ_BSTAGES;
                        tc::mbar_wait(&b_empty[s], ((bidx / DQ_BSTAGES) & 1) ^ 1);
                        DQ_TR(0, tr); ++tr;
                        if (rank == 0) dq_expect_tx(&b_full[s], (uint32_t)(2 * B_BYTES));
                        // k-chunk order is rotated per cluster: the pairs run in lockstep, and without it all 74 of them would ask the same
                        // few L2 lines of the same expert for the same chunk at the same time
                        const int kr = (kc + (int)cid) % kchunks;
                        dq_tma_load_2d(sB + s * B_SLOT, &map_w, kr * DQ_KC, s_e[i][j] * NI + (int)rank * NCT, b_full_l[s]);
                    }
                if (!next_issued) load_x(nxt, it + 1);
                ++it;
                i = nxt;
            }
            // drain: the leader's last multicast commit must have landed here before this CTA may leave the cluster
            if (it > 0) tc::mbar_wait(&a_empty[(it - 1) & 1], ((it - 1) >> 1) & 1);
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = tc::make_idesc_f16(2 * DQ_BM, NI);
            uint32_t it = 0, bidx = 0, pi = 0;               // pi counts expert passes: accumulator slot = pi & 1
            for (int i = 0; i < my_units; ++i) {
                const int nr = s_nr[i];
                if (nr == 0) continue;
                const uint32_t ab = it & 1;
                if (!(p.debug & 2)) tc::mbar_wait(&a_full[ab], (it >> 1) & 1);
                DQ_TR(1, tr); ++tr;
                tc::fence_after_sync();
                for (int j = 0; j < nr; ++j, ++pi) {
                    const uint32_t slot = pi & 1;
                    tc::mbar_wait(&t_empty[slot], ((pi >> 1) & 1) ^ 1);
                    DQ_TR(1, tr); ++tr;
                    tc::fence_after_sync();
                    const uint32_t tacc = tmem_base + slot * 256;
                    for (int kc = 0; kc < kchunks; ++kc, ++bidx) {
                        const int s = bidx % DQ_BSTAGES;
                        if (!(p.debug & 1)) tc::mbar_wait(&b_full[s], (bidx / DQ_BSTAGES) & 1);
                        DQ_TR(1, tr); ++tr;
                        tc::fence_after_sync();
                        const int kr = (kc + (int)cid) % kchunks;      // the weight chunk in stage s (see the producer)
                        const uint64_t adesc = tc::make_desc(smem_u32(sA + ab * A_TILE + kr * A_CHUNK), 1024, 2);
                        const uint64_t bdesc = tc::make_desc(smem_u32(sB + s * B_SLOT), 1024, 2);
                        if (!(p.debug & 16)) {
#pragma unroll
                            for (int k = 0; k < DQ_KC / 16; ++k) dq_mma2_f16_ss(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) ? 1u : 0u);
                        }
                        dq_commit2(&b_empty[s]);
                    }
                    dq_commit2(&t_full[slot]);
                }
                dq_commit2(&a_empty[ab]);                   // every MMA reading this unit's x chunks (in both CTAs) is done
                ++it;
            }
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;     // TMEM lane quarter; NCT contiguous output channels per thread
        const int r = q * 32 + lane;
        const bool elected = (warp == 2 && lane == 0);
        const __half2 hi2 = __float2half2_rn(p.clamp), lo2 = __float2half2_rn(-p.clamp);
        uint32_t t_empty_l[2];
        for (int a = 0; a < 2; ++a) t_empty_l[a] = dq_mapa(smem_u32(&t_empty[a]), 0);
        uint32_t pi = 0, sidx = 0;
        for (int i = 0; i < my_units; ++i) {
            const int u = p.unit0 + (int)cid + i * (int)ncl, b = u / p.upi;
            const int m0 = b * p.HW + (u - b * p.upi) * (2 * DQ_BM) + (int)rank * DQ_BM;
            const int nr = s_nr[i];
            const float w0 = s_w[i][0], w1 = s_w[i][1];
            // out = sum_j fp16( fp32(fp16(x W_j^T)) * w_j ) accumulated in fp16 like index_add_ on an fp16 tensor (utils.py:200-203)
            __half2 acc[NCT / 2];
#pragma unroll
            for (int c = 0; c < NCT / 2; ++c) acc[c] = __float2half2_rn(0.f);
            // Finished 32-channel groups leave at once through a 2-slot ring of 16 KB staging tiles (64-byte swizzled rows) and TMA
            // stores.  The write path (SM -> L2 -> HBM) is the slowest stage of the last pass whatever issues it (direct 32-byte
            // register stores, burst or interleaved, measured 3-10 % slower): the copy engine at least drains it asynchronously.
            auto stage32 = [&](int ci) {
                unsigned char* slot = stg + (sidx & 1) * (2 * DQ_BM * 64);
                if (elected) asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");   // the store before the previous one has read this slot
                asm volatile("bar.sync 1, 256;\n" ::: "memory");
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pk[4];
#pragma unroll
                    for (int h2 = 0; h2 < 4; ++h2) {
                        const __half2 v = __hmin2(__hmax2(acc[ci * 16 + c * 4 + h2], lo2), hi2);
                        pk[h2] = *reinterpret_cast<const uint32_t*>(&v);
                    }
                    *reinterpret_cast<uint4*>(slot + half * (DQ_BM * 64) + tc::sw64_offset(r, c)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
                tc::fence_proxy_async();
                asm volatile("bar.sync 1, 256;\n" ::: "memory");
                if (elected && !(p.debug & 4)) {
                    dq_tma_store_2d(&map_o, slot, ci * 32, m0);
                    dq_tma_store_2d(&map_o, slot + DQ_BM * 64, NCT + ci * 32, m0);
                }
                if (elected) asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                ++sidx;
            };
            for (int j = 0; j < nr; ++j, ++pi) {
                const uint32_t slot = pi & 1;
                const bool last = (j == nr - 1);
                tc::mbar_wait(&t_full[slot], (pi >> 1) & 1);
                tc::fence_after_sync();
                if (elected) { DQ_TR(2, tr); ++tr; }
                // routing weight as an exact two-term fp16 sum (w = wh + wl up to 2^-22 relative): the weighting then runs as two packed
                // fp16 FMAs on the fp16-rounded expert output instead of unpack / fp32 multiply / repack (3 of the 7 instructions per
                // pair, and the conversions are half-rate).  round16(t*wh + round16(t*wl)) equals the reference's round16(t*w) except
                // when t*w lies within 2^-22 (relative) of a rounding boundary: < 0.1 % of the elements, by one fp16 ulp.
                const float wj = j ? w1 : w0;
                const __half whs = __float2half_rn(wj);
                const __half2 wh = __half2half2(whs), wl = __float2half2_rn(wj - __half2float(whs)), zero2 = __float2half2_rn(0.f);
                const uint32_t tsrc = tmem_base + ((uint32_t)(q * 32) << 16) + slot * 256 + half * NCT;
                auto fold = [&](const uint32_t (&rr)[32], int c0) {
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const __half2 t = __floats2half2_rn(__uint_as_float(rr[2 * c]), __uint_as_float(rr[2 * c + 1]));
                        acc[c0 / 2 + c] = __hadd2(acc[c0 / 2 + c], __hfma2(t, wh, __hfma2(t, wl, zero2)));
                    }
                };
                if (!(p.debug & 8)) {
                    // TMEM reads (64 B/clk per SM) are the floor of this loop: keep one 32-column load in flight while folding the previous
                    uint32_t ra[32], rb[32];
                    tc::tmem_ld32(tsrc, ra);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int c0 = 0; c0 < NCT; c0 += 64) {
                        tc::tmem_ld32(tsrc + c0 + 32, rb);
                        fold(ra, c0);
                        if (last) stage32(c0 / 32);
                        tc::tmem_ld_wait();
                        if (c0 + 64 < NCT) tc::tmem_ld32(tsrc + c0 + 64, ra);
                        else {                              // every column of this slot is in registers: hand it back to the leader's MMA warp
                            tc::fence_before_sync();
                            __syncwarp();
                            if (lane == 0) dq_arrive_cluster(t_empty_l[slot]);
                        }
                        fold(rb, c0 + 32);
                        if (last) stage32(c0 / 32 + 1);
                        if (c0 + 64 < NCT) tc::tmem_ld_wait();
                    }
                } else {
                    tc::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) dq_arrive_cluster(t_empty_l[slot]);
                    if (last) {
#pragma unroll
                        for (int ci = 0; ci < NCT / 32; ++ci) stage32(ci);
                    }
                }
                if (elected) { DQ_TR(2, tr); ++tr; }
            }
            if (nr == 0) {                                  // no live route: the rows are exactly zero
#pragma unroll
                for (int ci = 0; ci < NCT / 32; ++ci) stage32(ci);
            }
            if (elected) { DQ_TR(2, tr); ++tr; }
        }
        if (elected) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    }
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) DQ_TR(3, 2);
    dq_cluster_sync();                                      // the peer may still be reading this CTA's weights / signalling its barriers
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
    if (tid == 0) DQ_TR(3, 3);
}

typedef CUresult (*DqEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static DqEncodeFn dq_encode() {
    static DqEncodeFn fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<DqEncodeFn>(q);
    }
    return fn;
}
static bool dq_map2d(CUtensorMap* m, const void* base, cuuint64_t cols, cuuint64_t rows, cuuint64_t pitch_bytes, cuuint32_t box_c,
                     cuuint32_t box_r, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_c, box_r};
    cuuint32_t est[2] = {1, 1};
    return dq_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// number of CTA pairs that can be co-resident (TPC pairs with both SMs free); 0 when cluster launch is unavailable
template <int NI>
static int dq_max_clusters_t(size_t smem) {
    if (cudaFuncSetAttribute(tc_dispatch2_kernel<NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2, 1, 1);
    cfg.blockDim = dim3(DQ_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, tc_dispatch2_kernel<NI>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n > 0 ? n : 0;
}
static int dq_max_clusters(int N, size_t smem) {
    static int cached[2] = {-1, -1};
    int& c = cached[N == 256];
    if (c < 0) c = N == 256 ? dq_max_clusters_t<256>(smem) : dq_max_clusters_t<128>(smem);
    return c;
}

}  // namespace ym

using namespace ym;

extern "C" int ym_dispatch_debug_mask();                    // tc_dispatch.cu: profiling mask shared with the single-CTA kernel
static long long* g_dispatch_trace = nullptr;
// Profiling aid: device buffer of 2 x 4 x 256 int64 that receives a clock64 timeline of CTAs 0 and 1 of the pair kernel.
extern "C" void ym_set_dispatch_trace(void* buf) { g_dispatch_trace = static_cast<long long*>(buf); }

static constexpr size_t DQ_SMEM = (size_t)8 * DQ_BM * 128 + (size_t)DQ_BSTAGES * 128 * 128 + 2 * DQ_BM * 128 + 1024;

// 1 when the CTA-pair kernel supports the shape: C % 64 == 0 <= 256, N in {128, 256}, top_k <= 2, HW % 256 == 0 (both tiles
// of a pair lie in one image, so they share the routed experts) and a device that can co-schedule 2-CTA clusters.
extern "C" int ym_moe_dispatch_v3_supported(int HW, int C, int N, int topk, int ldx, int ldw, int ldo) {
    return dq_encode() != nullptr && C % 64 == 0 && C <= 256 && (N == 128 || N == 256) && topk >= 1 && topk <= 2 && HW % 256 == 0 &&
           ldx % 8 == 0 && ldw % 8 == 0 && ldo % 8 == 0 && dq_max_clusters(N, DQ_SMEM) > 0;
}

extern "C" int ym_moe_dispatch_v3(const void* x, int ldx, int B, int HW, int C, const void* w_all, int ldw, int E, const int* route_idx,
                                  const float* route_w, int topk, int N, float w_min, float clamp, void* out, int ldo, void* stream) {
    YM_CHECK_ARG(x && w_all && route_idx && route_w && out, "ym_moe_dispatch_v3: null pointer");
    YM_CHECK_ARG(ym_moe_dispatch_v3_supported(HW, C, N, topk, ldx, ldw, ldo), "ym_moe_dispatch_v3: unsupported shape (HW %% 256, C %% 64 <= 256, "
                 "N in {128, 256}, top_k <= 2) or no cluster launch");
    YM_CHECK_ARG((((uintptr_t)x | (uintptr_t)w_all | (uintptr_t)out) & 15) == 0, "ym_moe_dispatch_v3: 16-byte alignment");
    if (B == 0) return YM_OK;
    CUtensorMap mx, mw, mo;
    if (!dq_map2d(&mx, x, (cuuint64_t)C, (cuuint64_t)B * HW, (cuuint64_t)ldx * 2, 64, DQ_BM) ||
        !dq_map2d(&mw, w_all, (cuuint64_t)C, (cuuint64_t)E * N, (cuuint64_t)ldw * 2, 64, (cuuint32_t)(N / 2)) ||
        !dq_map2d(&mo, out, (cuuint64_t)N, (cuuint64_t)B * HW, (cuuint64_t)ldo * 2, 32, DQ_BM, CU_TENSOR_MAP_SWIZZLE_64B)) {
        ym_set_error("ym_moe_dispatch_v3: cuTensorMapEncodeTiled failed");
        return YM_ERR_CUDA;
    }
    Dispatch2Params p;
    p.route_idx = route_idx; p.route_w = route_w; p.topk = topk; p.HW = HW; p.upi = HW / (2 * DQ_BM);
    p.K = C; p.N = N; p.w_min = w_min; p.clamp = clamp;
    p.debug = ym_dispatch_debug_mask(); p.trace = g_dispatch_trace;
    const long long total = (long long)B * p.upi;
    const int maxcl = dq_max_clusters(N, DQ_SMEM);
    for (long long u0 = 0; u0 < total; u0 += (long long)maxcl * DQ_MAXU) {
        const long long n = total - u0 < (long long)maxcl * DQ_MAXU ? total - u0 : (long long)maxcl * DQ_MAXU;
        p.unit0 = (int)u0; p.units = (int)n;
        const int ncl = n < maxcl ? (int)n : maxcl;
        if (N == 256) launch_pdl(tc_dispatch2_kernel<256>, 2 * ncl, DQ_THREADS, DQ_SMEM, (cudaStream_t)stream, mx, mw, mo, p);
        else launch_pdl(tc_dispatch2_kernel<128>, 2 * ncl, DQ_THREADS, DQ_SMEM, (cudaStream_t)stream, mx, mw, mo, p);
        YM_CHECK_LAUNCH("tc_dispatch2");
    }
    return YM_OK;
}
