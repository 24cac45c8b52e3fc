// TMA-staged, im2col-free tcgen05 convolution for NHWC fp16 (sm_100a): Conv(k=1|3, s=1|2, groups=1) + folded-BN bias
// + SiLU (+ residual), conv.py:69-89.
//
// The N-scale layers are HBM-bound on paper but *instruction-issue bound* with per-thread cp.async/ldmatrix/mma.sync
// (ncu: sm__throughput ~50 %, dram ~18 %).  Here the whole main loop is issued by two threads:
//   * producer (warp 0, one lane): one cp.async.bulk.tensor (TMA) per operand per k-tile.  The activation is a 4-D tensor
//     map (C, W, H, B); the k-tile for filter tap (ky,kx) is the box [kc channels] x [TW x TH pixels] at the SHIFTED
//     origin (x0*s + kx - pad, y0*s + ky - pad) with element strides (1, s, s, 1): no im2col buffer, padding comes from
//     TMA out-of-bounds zero fill, and the box lands in shared memory already in the 32/64/128-byte swizzled K-major
//     layout tcgen05 consumes.  1x1/stride-1 convs use a flat 2-D map (C, B*H*W) so tiles never straddle padding.
//   * MMA issuer (warp 1, one lane): tcgen05.mma kind::f16, M = 128 pixels, N = BN output channels, accumulator in TMEM;
//     tcgen05.commit releases the smem stage and finally signals the epilogue.
//   * epilogue (all 4 warps): tcgen05.ld one accumulator row (= one output pixel) per thread, bias + SiLU (+ residual),
//     16-byte vector stores.
// Several CTAs are resident per SM (72 KB smem, <= 128 TMEM columns each) so one CTA's epilogue overlaps another's loads.
#include <cuda.h>

#include "tc_common.cuh"

namespace ym {

constexpr int CV_BM = 128, CV_THREADS = 128, CV_STAGES = 3;

struct TcConvParams {
    const float* bias;
    const __half* res; int ldr;
    void* out; int ldo; int out_f32;
    int B, Ho, Wo, Cout, Cin, KH, KW, stride, pad;
    int kc;            // channels per k-tile (16 / 32 / 64)
    int tiles_x, tiles_y, TW, TH;   // patch mode
    int M;             // flat mode: B*H*W rows
    int act;
    int stages;        // operand ring depth of the persistent kernel (set by launch_conv2)
    int epi_groups;    // persistent kernel: 1 = eight epilogue warps per tile, 2 = two groups of four on alternate tiles
    int dbg;           // timing experiments: bit 0 no TMA store, bit 1 no epilogue arithmetic / staging, bit 2 no tensor-memory read, bit 3 weights loaded for the first tile only
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

template <int BN, bool FLAT>
__global__ void __launch_bounds__(CV_THREADS) tc_conv_kernel(const __grid_constant__ CUtensorMap map_a,
                                                             const __grid_constant__ CUtensorMap map_b, const TcConvParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (STS, not generic ST)
    const int row_bytes = p.kc * 2;                       // 32 / 64 / 128
    const int a_bytes = CV_BM * row_bytes, b_bytes = BN * row_bytes;
    const int stage_bytes = ((a_bytes + b_bytes + 1023) / 1024) * 1024;
    __shared__ uint64_t full_bar[CV_STAGES], empty_bar[CV_STAGES], done_bar;
    __shared__ uint32_t tmem_slot;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < CV_STAGES; ++s) {
            tc::mbar_init(&full_bar[s], 1);
            tc::mbar_init(&empty_bar[s], 1);
        }
        tc::mbar_init(&done_bar, 1);
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    pdl_prologue();    // barriers / tensor memory are set up: let the next kernel stage itself, then wait for our producer grid

    // ---- tile coordinates
    const int n0 = blockIdx.y * BN;
    int b = 0, oy0 = 0, ox0 = 0, m0 = 0;
    if (FLAT) {
        m0 = blockIdx.x * CV_BM;
    } else {
        const int tpi = p.tiles_x * p.tiles_y;
        b = blockIdx.x / tpi;
        const int t = blockIdx.x - b * tpi;
        oy0 = (t / p.tiles_x) * p.TH;
        ox0 = (t % p.tiles_x) * p.TW;
    }
    const int cchunks = (p.Cin + p.kc - 1) / p.kc;
    const int KT = p.KH * p.KW * cchunks;
    const uint32_t layout_type = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
    const uint32_t sbo = 8u * row_bytes;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer
        for (int it = 0; it < KT; ++it) {
            const int s = it % CV_STAGES;
            if (it >= CV_STAGES) tc::mbar_wait(&empty_bar[s], ((it / CV_STAGES) - 1) & 1);
            unsigned char* st = smem + s * stage_bytes;
            const int tap = it / cchunks, c0 = (it - tap * cchunks) * p.kc;
            mbar_expect_tx(&full_bar[s], (uint32_t)(a_bytes + b_bytes));
            if (FLAT) {
                tma_load_2d(st, &map_a, c0, m0, &full_bar[s]);
            } else {
                const int ky = tap / p.KW, kx = tap - ky * p.KW;
                tma_load_4d(st, &map_a, c0, ox0 * p.stride + kx - p.pad, oy0 * p.stride + ky - p.pad, b, &full_bar[s]);
            }
            tma_load_2d(st + a_bytes, &map_b, tap * p.Cin + c0, n0, &full_bar[s]);
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer
        const uint32_t idesc = tc::make_idesc_f16(CV_BM, BN);
        for (int it = 0; it < KT; ++it) {
            const int s = it % CV_STAGES;
            tc::mbar_wait(&full_bar[s], (it / CV_STAGES) & 1);
            tc::fence_after_sync();
            const uint32_t sa = smem_u32(smem + s * stage_bytes);
            const uint64_t adesc = tc::make_desc(sa, sbo, layout_type), bdesc = tc::make_desc(sa + a_bytes, sbo, layout_type);
            for (int k = 0; k < p.kc / 16; ++k) tc::mma_f16_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) ? 1u : 0u);
            tc::mma_commit(&empty_bar[s]);
        }
        tc::mma_commit(&done_bar);
    }
    __syncwarp();

    // ===== epilogue: one output pixel (accumulator row) per thread
    tc::mbar_wait(&done_bar, 0);
    tc::fence_after_sync();
    const int r = warp * 32 + lane;
    long long opix;
    bool rowok;
    if (FLAT) {
        opix = (long long)m0 + r;
        rowok = opix < p.M;
    } else {
        const int ty = r / p.TW, tx = r - ty * p.TW;
        const int oy = oy0 + ty, ox = ox0 + tx;
        rowok = ty < p.TH && oy < p.Ho && ox < p.Wo;
        opix = ((long long)b * p.Ho + oy) * p.Wo + ox;
    }
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    constexpr int NCHUNK = BN < 16 ? 1 : BN / 16;
    // fp32 outputs (the class-logit convs of the Detect towers) in FLAT mode: a thread's 16 floats of a chunk go through a warp-private
    // shared-memory strip (pitch 20 floats: conflict-free float4 accesses) so that one store instruction covers eight rows x 64
    // contiguous bytes.  Row-per-thread float4 stores touched 32 half-used sectors per instruction (59 us for the P3 logits, 44 MB).
    __shared__ __align__(16) float f32_stg[CV_THREADS / 32][32][20];
    const bool staged_f32 = FLAT && p.out_f32 && p.res == nullptr && (p.ldo % 4 == 0);
#pragma unroll 1
    for (int ci = 0; ci < NCHUNK; ++ci) {
        uint32_t rr[16];
        tc::tmem_ld16(lane_addr + ci * 16, rr);
        tc::tmem_ld_wait();
        const int n = n0 + ci * 16;
        if (n >= p.Cout) continue;                  // uniform across the CTA
        if (staged_f32 && n + 16 <= p.Cout) {
            float4* mine = reinterpret_cast<float4*>(&f32_stg[warp][lane][0]);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                float x[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    x[e] = __uint_as_float(rr[4 * q4 + e]);
                    if (p.bias != nullptr) x[e] += p.bias[n + 4 * q4 + e];
                    if (p.act == 1) x[e] = silu_f(x[e]);
                }
                mine[q4] = make_float4(x[0], x[1], x[2], x[3]);
            }
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int rl = it * 8 + (lane >> 2), c4 = lane & 3;
                const long long m = (long long)m0 + warp * 32 + rl;
                if (m < p.M)
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + m * p.ldo + n + 4 * c4) =
                        *reinterpret_cast<const float4*>(&f32_stg[warp][rl][4 * c4]);
            }
            __syncwarp();
            continue;
        }
        if (!rowok) continue;
        float v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            float x = __uint_as_float(rr[q]);
            if (p.bias != nullptr && n + q < p.Cout) x += p.bias[n + q];
            if (p.act == 1) x = silu_f(x);
            v[q] = x;
        }
        if (n + 16 <= p.Cout) {
            if (p.res != nullptr) {
                const Half8 r0 = *reinterpret_cast<const Half8*>(p.res + opix * p.ldr + n);
                const Half8 r1 = *reinterpret_cast<const Half8*>(p.res + opix * p.ldr + n + 8);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f0 = __half22float2(r0.v[q]), f1 = __half22float2(r1.v[q]);
                    v[2 * q] += f0.x; v[2 * q + 1] += f0.y; v[8 + 2 * q] += f1.x; v[8 + 2 * q + 1] += f1.y;
                }
            }
            if (p.out_f32) {
                float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + opix * p.ldo + n);
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
                Half8 o0, o1;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    o0.v[q] = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
                    o1.v[q] = __floats2half2_rn(v[8 + 2 * q], v[8 + 2 * q + 1]);
                }
                __half* dst = reinterpret_cast<__half*>(p.out) + opix * p.ldo + n;
                *reinterpret_cast<Half8*>(dst) = o0;
                *reinterpret_cast<Half8*>(dst + 8) = o1;
            }
        } else {
            for (int q = 0; q < 16 && n + q < p.Cout; ++q) {
                float x = v[q];
                if (p.res != nullptr) x += __half2float(p.res[opix * p.ldr + n + q]);
                if (p.out_f32) reinterpret_cast<float*>(p.out)[opix * p.ldo + n + q] = x;
                else reinterpret_cast<__half*>(p.out)[opix * p.ldo + n + q] = __float2half_rn(x);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// =====================================================================================================================
// v2: persistent, warp-specialised variant (the production path).
//   warp 0 (one lane)  TMA producer  : per k-tile one bulk-tensor load for A (tap-shifted box) and one for B
//   warp 1 (one lane)  MMA issuer    : tcgen05.mma into one of TWO TMEM accumulators (tile i+1 overlaps the epilogue of i)
//   warps 2..5         epilogue      : tcgen05.ld (row per thread) -> bias/SiLU/residual -> fp16 -> 128B-swizzled smem
//                                      staging tile -> ONE TMA store per tile (coalesced, clips image borders / ragged M)
// CTAs are persistent (grid = 2 x #SMs at most) and walk the tile list with a static stride.
// =====================================================================================================================
constexpr int CV2_THREADS = 320, CV2_EPI_THREADS = 256;   // 2 control warps + 8 epilogue warps
// Operand ring depth is chosen per launch (TcConvParams::stages): a k-tile of a low-channel layer is only 4-6 KB (Cin = 16:
// 128 rows x 32 B + weights), and with 3 stages a CTA had ~15 KB in flight - far too little to cover the ~1 us L2/HBM latency
// (profiles/r01_launch_roofline.txt: the P1/P2 convs ran at 0.5-1.2 TB/s).  The ring now takes up to ~64 KB: 3 stages of
// 24 KB at Cin >= 64, 6 at Cin = 32, 12 at Cin = 16.
constexpr int CV2_MIN_STAGES = 3, CV2_MAX_STAGES = 12, CV2_RING_BYTES = 64 * 1024;
static inline int cv2_stages(int stage_bytes) {
    int n = CV2_RING_BYTES / stage_bytes;
    return n < CV2_MIN_STAGES ? CV2_MIN_STAGES : (n > CV2_MAX_STAGES ? CV2_MAX_STAGES : n);
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

// SiLU over a register array, written stage by stage so that every stage is NC independent instructions (the MUFU and
// FMA pipes stay busy instead of waiting on one element's ex2 -> add -> rcp -> mul chain).
template <int NC>
__device__ __forceinline__ void silu_array(float (&v)[NC]) {
    float e[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) e[j] = v[j] * -1.4426950408889634f;
#pragma unroll
    for (int j = 0; j < NC; ++j) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[j]) : "f"(e[j]));
#pragma unroll
    for (int j = 0; j < NC; ++j) e[j] += 1.0f;
#pragma unroll
    for (int j = 0; j < NC; ++j) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(e[j]) : "f"(e[j]));
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] *= e[j];
}

template <int BN, bool FLAT, bool DBG>
__global__ void __launch_bounds__(CV2_THREADS, 2) tc_conv2_kernel(const __grid_constant__ CUtensorMap map_a,
                                                               const __grid_constant__ CUtensorMap map_b,
                                                               const __grid_constant__ CUtensorMap map_o, const TcConvParams p,
                                                               int m_tiles, int n_tiles) {
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment by OFFSET arithmetic: a round trip through uintptr_t loses the shared address space and every staging store
    // below became a generic ST.E.128 with 64-bit address arithmetic and a MEMBAR before the proxy fence (SASS of round-2 call 12)
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int dbg = DBG ? p.dbg : 0;                       // timing probe (tools/conv2_probe.py): compiled out of the production kernels
    const int row_bytes = p.kc * 2;
    const int a_bytes = CV_BM * row_bytes, b_bytes = BN * row_bytes;
    const int stage_bytes = ((a_bytes + b_bytes + 1023) / 1024) * 1024;
    constexpr int OUT_ROW = BN * 2;                        // bytes per staged output row (32 / 64 / 128)
    constexpr int OUT_BYTES = CV_BM * OUT_ROW;
    const int nst = p.stages;
    unsigned char* stg = smem + nst * stage_bytes;         // [2][OUT_BYTES], 1024-aligned
    __shared__ uint64_t full_bar[CV2_MAX_STAGES], empty_bar[CV2_MAX_STAGES], tfull_bar[2], tempty_bar[2], sfree_bar[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float sbias[1024 + 64];       // folded-BN bias, zero padded past Cout
    constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 1024 + 64; i += CV2_THREADS) sbias[i] = (p.bias != nullptr && i < p.Cout) ? p.bias[i] : 0.f;
    if (tid == 0) {
        for (int s = 0; s < nst; ++s) {
            tc::mbar_init(&full_bar[s], 1);
            tc::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            tc::mbar_init(&tfull_bar[a], 1);
            tc::mbar_init(&tempty_bar[a], p.epi_groups == 2 ? 4 : 8);   // one arrive per epilogue warp working on that accumulator
            tc::mbar_init(&sfree_bar[a], 1);          // two-group epilogue: the group's staging buffer has been read by its last store
        }
        tc::fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(&tmem_slot, TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    pdl_prologue();    // barriers / tensor memory are set up: let the next kernel stage itself, then wait for our producer grid

    const int total_tiles = m_tiles * n_tiles;
    const int cchunks = (p.Cin + p.kc - 1) / p.kc;
    const int KT = p.KH * p.KW * cchunks;
    const uint32_t layout_type = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
    const uint32_t sbo = 8u * row_bytes;

    auto tile_coords = [&](int tile, int& n0, int& m0, int& b, int& oy0, int& ox0) {
        const int mt = tile / n_tiles;
        n0 = (tile - mt * n_tiles) * BN;
        m0 = mt * CV_BM;
        b = 0; oy0 = 0; ox0 = 0;
        if (!FLAT) {
            const int tpi = p.tiles_x * p.tiles_y;
            b = mt / tpi;
            const int t = mt - b * tpi;
            oy0 = (t / p.tiles_x) * p.TH;
            ox0 = (t % p.tiles_x) * p.TW;
        }
    };

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer
            int s = 0;
            uint32_t ph = 0;                               // ring position and its phase bit
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int n0, m0, b, oy0, ox0;
                tile_coords(tile, n0, m0, b, oy0, ox0);
                for (int it = 0; it < KT; ++it) {
                    tc::mbar_wait(&empty_bar[s], ph ^ 1);
                    unsigned char* st = smem + s * stage_bytes;
                    const int tap = it / cchunks, c0 = (it - tap * cchunks) * p.kc;
                    const bool load_b = !(dbg & 8) || tile == (int)blockIdx.x;
                    mbar_expect_tx(&full_bar[s], (uint32_t)(a_bytes + (load_b ? b_bytes : 0)));
                    if (FLAT) {
                        tma_load_2d(st, &map_a, c0, m0, &full_bar[s]);
                    } else {
                        const int ky = tap / p.KW, kx = tap - ky * p.KW;
                        tma_load_4d(st, &map_a, c0, ox0 * p.stride + kx - p.pad, oy0 * p.stride + ky - p.pad, b, &full_bar[s]);
                    }
                    if (load_b) tma_load_2d(st + a_bytes, &map_b, tap * p.Cin + c0, n0, &full_bar[s]);
                    if (++s == nst) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer
            const uint32_t idesc = tc::make_idesc_f16(CV_BM, BN);
            uint32_t titer = 0, ph = 0;
            int s = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++titer) {
                const uint32_t acc = titer & 1;
                tc::mbar_wait(&tempty_bar[acc], ((titer >> 1) & 1) ^ 1);       // epilogue has drained this accumulator
                tc::fence_after_sync();
                const uint32_t tacc = tmem_base + acc * BN;
                for (int it = 0; it < KT; ++it) {
                    tc::mbar_wait(&full_bar[s], ph);
                    tc::fence_after_sync();
                    const uint32_t sa = smem_u32(smem + s * stage_bytes);
                    const uint64_t adesc = tc::make_desc(sa, sbo, layout_type), bdesc = tc::make_desc(sa + a_bytes, sbo, layout_type);
                    for (int k = 0; k < p.kc / 16; ++k) tc::mma_f16_ss(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) ? 1u : 0u);
                    tc::mma_commit(&empty_bar[s]);
                    if (++s == nst) { s = 0; ph ^= 1; }
                }
                tc::mma_commit(&tfull_bar[acc]);
            }
        }
    } else {
        // ===== 8 epilogue warps: TMEM lane quarter = warp % 4 (hardware rule).
        // p.epi_groups == 1: all eight work on the same tile, column half = (warp - 2) / 4.
        // p.epi_groups == 2: warps 2-5 and 6-9 are two independent groups that take ALTERNATE tiles (group g owns accumulator g, staging
        //   buffer g and named barrier 1 + g); a thread then handles every column of its row in two passes.  One tile's chain - wait for
        //   the accumulator, tensor-memory read, bias / SiLU (2 MUFU per value: the busiest unit of these kernels), staging, barrier,
        //   store - no longer holds all epilogue warps of the CTA: with the timing probe (profiles/r02_conv2_probe.json) the arithmetic
        //   and the CTA-wide barrier were 29 + 10 us of a 61 us layer whose loads alone take 18 us.
        constexpr int NC = BN / 2;                 // columns per thread and pass: 8 / 16 / 32
        const int ngrp = p.epi_groups, grp = ngrp == 2 ? (warp - 2) >> 2 : 0;
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int npass = ngrp == 2 ? 2 : 1;
        const int r = q * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool elected = ((ngrp == 2 ? (warp & 3) == 2 : warp == 2) && lane == 0);
        uint32_t titer = 0;                        // this group's tile counter
        // (m tile, n tile) of the group's current tile, advanced by carry instead of a division per tile
        const int tstep = ngrp * (int)gridDim.x, step_mt = tstep / n_tiles, step_nt = tstep - step_mt * n_tiles;
        int cur_mt = (int)(blockIdx.x + grp * gridDim.x) / n_tiles, cur_nt = (int)(blockIdx.x + grp * gridDim.x) - cur_mt * n_tiles;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < total_tiles; tile += tstep, ++titer) {
            int n0, m0, b = 0, oy0 = 0, ox0 = 0;
            if (FLAT) {
                n0 = cur_nt * BN;
                m0 = cur_mt * CV_BM;
                cur_nt += step_nt;
                cur_mt += step_mt;
                if (cur_nt >= n_tiles) { cur_nt -= n_tiles; ++cur_mt; }
            } else {
                tile_coords(tile, n0, m0, b, oy0, ox0);
            }
            const uint32_t acc = ngrp == 2 ? (uint32_t)grp : (titer & 1);
            const uint32_t acc_phase = ngrp == 2 ? (titer & 1) : ((titer >> 1) & 1);
            unsigned char* sbuf = stg + acc * OUT_BYTES;
            long long opix;
            bool rowok;
            if (FLAT) {
                opix = (long long)m0 + r;
                rowok = opix < p.M;
            } else {
                const int ty = r / p.TW, tx = r - ty * p.TW;
                const int oy = oy0 + ty, ox = ox0 + tx;
                rowok = oy < p.Ho && ox < p.Wo;
                opix = ((long long)b * p.Ho + oy) * p.Wo + ox;
            }
            bool waited = false;
            for (int pass = 0; pass < npass; ++pass) {
                const int cbase = (ngrp == 2 ? pass : half) * NC;
                const int n = n0 + cbase;
                // residual slice of this thread's row: issued before waiting for the accumulator, lands while the MMAs run
                Half8 rv[NC / 8];
                const bool has_res = p.res != nullptr && rowok && n + NC <= p.Cout;
                if (has_res) {
#pragma unroll
                    for (int c = 0; c < NC / 8; ++c) rv[c] = *reinterpret_cast<const Half8*>(p.res + opix * p.ldr + n + c * 8);
                }
                if (!waited) {
                    tc::mbar_wait(&tfull_bar[acc], acc_phase);
                    tc::fence_after_sync();
                    waited = true;
                }
                uint32_t rr[NC];
                if (!(dbg & 4)) {
                    const uint32_t ta = lane_base + acc * BN + cbase;
                    if constexpr (NC == 32) tc::tmem_ld32(ta, rr);
                    else if constexpr (NC == 16) tc::tmem_ld16(ta, rr);
                    else tc::tmem_ld8(ta, rr);
                    tc::tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < NC; ++j) rr[j] = 0u;
                }
                if (pass == npass - 1) {
                    tc::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&tempty_bar[acc]);   // accumulator drained: the MMAs of the next tile on it may start
                }
                constexpr int CH = NC < 16 ? NC : 16;      // process 8 / 16 columns at a time (register budget: 2 CTAs/SM)
                if (dbg & 2) continue;
                if (ngrp == 2 && pass == 0) {
                    // a group has ONE staging buffer: the TMA store of its previous tile must have read it before anyone writes again.
                    // The store was issued a whole accumulator wait + tensor-memory read ago, so this rarely blocks.
                    if (elected) {
                        tma_store_wait_read0();
                        tc::mbar_arrive(&sfree_bar[grp]);
                    }
                    tc::mbar_wait(&sfree_bar[grp], titer & 1);
                }
#pragma unroll
                for (int c0 = 0; c0 < NC; c0 += CH) {
                    float v[CH];
#pragma unroll
                    for (int j = 0; j < CH; j += 4) {
                        const float4 bb = *reinterpret_cast<const float4*>(&sbias[n + c0 + j]);
                        v[j] = __uint_as_float(rr[c0 + j]) + bb.x;
                        v[j + 1] = __uint_as_float(rr[c0 + j + 1]) + bb.y;
                        v[j + 2] = __uint_as_float(rr[c0 + j + 2]) + bb.z;
                        v[j + 3] = __uint_as_float(rr[c0 + j + 3]) + bb.w;
                    }
                    if (p.act == 1) silu_array<CH>(v);
                    if (has_res) {
#pragma unroll
                        for (int c = 0; c < CH / 8; ++c)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 f = __half22float2(rv[c0 / 8 + c].v[j]);
                                v[c * 8 + 2 * j] += f.x;
                                v[c * 8 + 2 * j + 1] += f.y;
                            }
                    } else if (p.res != nullptr && rowok) {    // ragged channel tail: static indices (a runtime-bounded loop put v[] in local memory)
#pragma unroll
                        for (int j = 0; j < CH; ++j)
                            if (n + c0 + j < p.Cout) v[j] += __half2float(p.res[opix * p.ldr + n + c0 + j]);
                    }
                    // staged row r, 16-byte chunks of this thread's columns, swizzled as the TMA store map expects
#pragma unroll
                    for (int c = 0; c < CH / 8; ++c) {
                        uint4 w;
                        w.x = pack_half2(v[8 * c + 0], v[8 * c + 1]);
                        w.y = pack_half2(v[8 * c + 2], v[8 * c + 3]);
                        w.z = pack_half2(v[8 * c + 4], v[8 * c + 5]);
                        w.w = pack_half2(v[8 * c + 6], v[8 * c + 7]);
                        const int ch = (cbase + c0) / 8 + c;
                        uint32_t off;
                        if (OUT_ROW == 128) off = tc::sw128_offset(r, ch);
                        else if (OUT_ROW == 64) off = tc::sw64_offset(r, ch);
                        else off = (uint32_t)(r * 32 + ((ch ^ ((r >> 2) & 1)) << 4));   // 32-byte swizzle
                        if (!(dbg & 64)) *reinterpret_cast<uint4*>(sbuf + off) = w;
                        else if (w.x == 0x7fc07fc0u) sbuf[off] = 1;   // keeps the arithmetic alive
                    }
                }
            }
            if (dbg & 2) continue;
            if (!(dbg & 16)) tc::fence_proxy_async();       // staged tile -> visible to the TMA store
            if (elected && ngrp == 1) tma_store_wait_read0();   // one group: the previous tile's store has released the OTHER buffer
            if (!(dbg & 32)) {                           // literal barrier ids: a register id makes ptxas reserve all sixteen
                if (ngrp == 1) epi_barrier();
                else if (grp == 0) asm volatile("bar.sync 1, 128;\n" ::: "memory");
                else asm volatile("bar.sync 2, 128;\n" ::: "memory");
            }
            if (elected && !(dbg & 1)) {
                if (FLAT) tma_store_2d(&map_o, sbuf, n0, m0);
                else tma_store_4d(&map_o, sbuf, n0, ox0, oy0, b);
                tma_store_commit();
            }
        }
        if (elected) tma_store_wait_all();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---- host side: tensor-map encoding through the driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static CUtensorMapSwizzle swizzle_for(int row_bytes) {
    return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

static int g_conv2_epi_groups = 2;
static int g_conv2_debug = 0;          // timing experiments only (ym_set_conv2_debug): results are invalid when non-zero

template <int BN, bool FLAT, bool DBG>
static int launch_conv2_t(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const TcConvParams& p_in, int m_tiles,
                        int n_tiles, cudaStream_t st) {
    TcConvParams p = p_in;
    const int row_bytes = p.kc * 2;
    const int stage = ((CV_BM * row_bytes + BN * row_bytes + 1023) / 1024) * 1024;
    p.stages = cv2_stages(stage);
    p.epi_groups = g_conv2_epi_groups;
    p.dbg = g_conv2_debug & 255;
    if ((g_conv2_debug >> 8) & 15) p.stages = (g_conv2_debug >> 8) & 15;
    const size_t smem = (size_t)p.stages * stage + 2 * (size_t)CV_BM * BN * 2 + 1024;
    auto kern = tc_conv2_kernel<BN, FLAT, DBG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ym_set_error("tc_conv2: smem attr %zu: %s", smem, cudaGetErrorString(e)); return YM_ERR_CUDA; }
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const int total = m_tiles * n_tiles;
    const int grid = total < 2 * sms ? total : 2 * sms;
    launch_pdl(kern, grid, CV2_THREADS, smem, st, ma, mb, mo, p, m_tiles, n_tiles);
    YM_CHECK_LAUNCH("tc_conv2");
    return YM_OK;
}

template <int BN, bool FLAT>
static int launch_conv2(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const TcConvParams& p, int m_tiles, int n_tiles,
                        cudaStream_t st) {
    if (g_conv2_debug & 255) return launch_conv2_t<BN, FLAT, true>(ma, mb, mo, p, m_tiles, n_tiles, st);
    return launch_conv2_t<BN, FLAT, false>(ma, mb, mo, p, m_tiles, n_tiles, st);
}

template <int BN, bool FLAT>
static int launch_conv(const CUtensorMap& ma, const CUtensorMap& mb, const TcConvParams& p, dim3 grid, cudaStream_t st) {
    const int row_bytes = p.kc * 2;
    const int stage = ((CV_BM * row_bytes + BN * row_bytes + 1023) / 1024) * 1024;
    const size_t smem = (size_t)CV_STAGES * stage + 1024;
    auto kern = tc_conv_kernel<BN, FLAT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ym_set_error("tc_conv: smem attr %zu: %s", smem, cudaGetErrorString(e)); return YM_ERR_CUDA; }
    launch_pdl(kern, grid, CV_THREADS, smem, st, ma, mb, p);
    YM_CHECK_LAUNCH("tc_conv");
    return YM_OK;
}

}  // namespace ym

using namespace ym;

static int g_tc_conv_version = 2;   // 2 = persistent warp-specialised kernel + TMA store, 1 = one tile per CTA
extern "C" int ym_set_tc_conv_version(int v) {
    const int old = g_tc_conv_version;
    if (v == 1 || v == 2) g_tc_conv_version = v;
    return old;
}

// Returns 1 if ym_conv2d_tc supports this configuration (the Python layer falls back to ym_conv2d_nhwc otherwise).
extern "C" int ym_conv2d_tc_supported(int Cin, int Cout, int KH, int KW, int stride, int pad, int ldx) {
    if (!(KH == KW && (KH == 1 || KH == 3) && (stride == 1 || stride == 2) && pad == KH / 2)) return 0;
    if (Cin % 8 != 0 || Cin < 16) return 0;
    if (Cout % 8 != 0 || ldx % 8 != 0) return 0;
    if (Cout > 1024) return 0;      // the kernels stage the folded bias of all output channels in a 1088-float shared table
    return get_encode() != nullptr;
}

extern "C" int ym_conv2d_tc(const void* x, int ldx, int B, int H, int W, int Cin, const void* w, int Kpad, const float* bias,
                            int Cout, int KH, int KW, int stride, int pad, void* out, int ldo, int out_f32, const void* res,
                            int ldr, int act, void* stream) {
    YM_CHECK_ARG(x && w && out, "ym_conv2d_tc: null pointer");
    YM_CHECK_ARG(ym_conv2d_tc_supported(Cin, Cout, KH, KW, stride, pad, ldx), "ym_conv2d_tc: unsupported configuration "
                 "(Cin=%d Cout=%d k=%d s=%d p=%d); use ym_conv2d_nhwc", Cin, Cout, KH, stride, pad);
    YM_CHECK_ARG((((uintptr_t)x | (uintptr_t)w | (uintptr_t)out | (uintptr_t)res) & 15) == 0, "ym_conv2d_tc: 16-byte alignment");
    YM_CHECK_ARG(ldo % (out_f32 ? 4 : 8) == 0 && (res == nullptr || ldr % 8 == 0) && Kpad % 8 == 0, "ym_conv2d_tc: pitches");
    if (B == 0) return YM_OK;
    EncodeTiledFn enc = get_encode();
    TcConvParams p;
    memset(&p, 0, sizeof(p));
    p.bias = bias; p.res = (const __half*)res; p.ldr = ldr; p.out = out; p.ldo = ldo; p.out_f32 = out_f32;
    p.B = B; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad; p.act = act;
    p.Ho = (H + 2 * pad - KH) / stride + 1;
    p.Wo = (W + 2 * pad - KW) / stride + 1;
    // channels per k-tile: a ragged last chunk (e.g. Cin = 48, 80, 96) is zero-filled by TMA out-of-bounds handling, so the
    // box always spans a full 128-byte swizzle row once Cin > 32 (one bulk load per tap instead of several narrow ones)
    p.kc = Cin > 32 ? 64 : (Cin > 16 ? 32 : 16);
    const int row_bytes = p.kc * 2;
    const bool flat = (KH == 1 && stride == 1);
    int BN;
    if (Cout <= 16) BN = 16; else if (Cout <= 32) BN = 32; else if (Cout <= 64) BN = 64;
    else if (Cout <= 128 || Cout % 128 == 0) BN = 128; else BN = 64;
    const int ntiles = (Cout + BN - 1) / BN;

    CUtensorMap ma, mb;
    CUresult cr;
    if (flat) {
        p.M = B * H * W;
        cuuint64_t gdim[2] = {(cuuint64_t)Cin, (cuuint64_t)p.M};
        cuuint64_t gstr[1] = {(cuuint64_t)ldx * 2};
        cuuint32_t box[2] = {(cuuint32_t)p.kc, (cuuint32_t)CV_BM};
        cuuint32_t est[2] = {1, 1};
        cr = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(x), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 swizzle_for(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        // output patch TH x TW = 128 pixels; prefer the wider tile unless the narrower one wastes fewer pixels
        const int w16 = ((p.Wo + 15) / 16) * 16 * (((p.Ho + 7) / 8) * 8), w8 = ((p.Wo + 7) / 8) * 8 * (((p.Ho + 15) / 16) * 16);
        p.TW = (w8 < w16) ? 8 : 16;
        p.TH = CV_BM / p.TW;
        p.tiles_x = (p.Wo + p.TW - 1) / p.TW;
        p.tiles_y = (p.Ho + p.TH - 1) / p.TH;
        cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t gstr[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)W * ldx * 2, (cuuint64_t)H * W * ldx * 2};
        cuuint32_t box[4] = {(cuuint32_t)p.kc, (cuuint32_t)(p.TW * stride), (cuuint32_t)(p.TH * stride), 1};
        cuuint32_t est[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        cr = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 swizzle_for(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (cr != CUDA_SUCCESS) { ym_set_error("ym_conv2d_tc: cuTensorMapEncodeTiled(activation) failed: %d", (int)cr); return YM_ERR_CUDA; }
    {
        cuuint64_t gdim[2] = {(cuuint64_t)Kpad, (cuuint64_t)Cout};
        cuuint64_t gstr[1] = {(cuuint64_t)Kpad * 2};
        cuuint32_t box[2] = {(cuuint32_t)p.kc, (cuuint32_t)BN};
        cuuint32_t est[2] = {1, 1};
        cr = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 swizzle_for(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (cr != CUDA_SUCCESS) { ym_set_error("ym_conv2d_tc: cuTensorMapEncodeTiled(weights) failed: %d", (int)cr); return YM_ERR_CUDA; }

    cudaStream_t st = (cudaStream_t)stream;
    const int mtiles = flat ? (p.M + CV_BM - 1) / CV_BM : B * p.tiles_x * p.tiles_y;
    if (!out_f32 && Cout % 8 == 0 && g_tc_conv_version == 2) {
        // persistent warp-specialised kernel with TMA-store epilogue (BN <= 64 keeps two CTAs per SM resident)
        const int BN2 = Cout <= 16 ? 16 : (Cout <= 32 ? 32 : 64);
        const int nt2 = (Cout + BN2 - 1) / BN2;
        CUtensorMap mb2, mo;
        {
            cuuint64_t gdim[2] = {(cuuint64_t)Kpad, (cuuint64_t)Cout};
            cuuint64_t gstr[1] = {(cuuint64_t)Kpad * 2};
            cuuint32_t box[2] = {(cuuint32_t)p.kc, (cuuint32_t)BN2};
            cuuint32_t est[2] = {1, 1};
            cr = enc(&mb2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle_for(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) { ym_set_error("ym_conv2d_tc: cuTensorMapEncodeTiled(weights v2) failed: %d", (int)cr); return YM_ERR_CUDA; }
        }
        if (flat) {
            cuuint64_t gdim[2] = {(cuuint64_t)Cout, (cuuint64_t)p.M};
            cuuint64_t gstr[1] = {(cuuint64_t)ldo * 2};
            cuuint32_t box[2] = {(cuuint32_t)BN2, (cuuint32_t)CV_BM};
            cuuint32_t est[2] = {1, 1};
            cr = enc(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, out, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle_for(BN2 * 2), CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            cuuint64_t gdim[4] = {(cuuint64_t)Cout, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)B};
            cuuint64_t gstr[3] = {(cuuint64_t)ldo * 2, (cuuint64_t)p.Wo * ldo * 2, (cuuint64_t)p.Ho * p.Wo * ldo * 2};
            cuuint32_t box[4] = {(cuuint32_t)BN2, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
            cuuint32_t est[4] = {1, 1, 1, 1};
            cr = enc(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, out, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle_for(BN2 * 2), CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (cr != CUDA_SUCCESS) { ym_set_error("ym_conv2d_tc: cuTensorMapEncodeTiled(output) failed: %d", (int)cr); return YM_ERR_CUDA; }
        if (flat) {
            switch (BN2) {
                case 16: return launch_conv2<16, true>(ma, mb2, mo, p, mtiles, nt2, st);
                case 32: return launch_conv2<32, true>(ma, mb2, mo, p, mtiles, nt2, st);
                default: return launch_conv2<64, true>(ma, mb2, mo, p, mtiles, nt2, st);
            }
        } else {
            switch (BN2) {
                case 16: return launch_conv2<16, false>(ma, mb2, mo, p, mtiles, nt2, st);
                case 32: return launch_conv2<32, false>(ma, mb2, mo, p, mtiles, nt2, st);
                default: return launch_conv2<64, false>(ma, mb2, mo, p, mtiles, nt2, st);
            }
        }
    }
    dim3 grid(mtiles, ntiles, 1);
#define YM_LC(BN_)                                                          \
    (flat ? launch_conv<BN_, true>(ma, mb, p, grid, st) : launch_conv<BN_, false>(ma, mb, p, grid, st))
    switch (BN) {
        case 16: return YM_LC(16);
        case 32: return YM_LC(32);
        case 64: return YM_LC(64);
        default: return YM_LC(128);
    }
#undef YM_LC
}

// Epilogue organisation of the persistent tcgen05 conv kernel: 2 = two groups of four warps on alternate tiles (default), 1 = all eight
// warps on every tile.  Bit-identical results.  Returns the previous setting (A/B measurements and tests).
extern "C" int ym_set_conv2_epi_groups(int n) {
    const int old = g_conv2_epi_groups;
    if (n == 1 || n == 2) g_conv2_epi_groups = n;
    return old;
}

// Timing experiments on the persistent kernel (tools/conv2_probe.py): bits 0-3 switch parts of it off (TcConvParams::dbg), bit 4 no proxy fence, bit 5 no epilogue barrier, bit 6 no staging
// writes; bits 8-11 override the operand ring depth.  Outputs are INVALID while non-zero; nothing in the package sets it.  Returns the previous value.
extern "C" int ym_set_conv2_debug(int flags) {
    const int old = g_conv2_debug;
    if (flags >= 0 && flags < 4096) g_conv2_debug = flags;
    return old;
}
