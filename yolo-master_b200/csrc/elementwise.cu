// HBM-bound NHWC fp16 kernels: stem conv (NCHW image -> NHWC16), depthwise k x k, SPPF pooling,
// nearest-upsample + concat.  All accesses are 16-byte vectors over the channel dimension.
#include "ym_common.cuh"

namespace ym {

// ---------------------------------------------------------------------------------------------
// Stem: Conv(3->COUT, k3 s2 p1) + folded BN + SiLU, reading the NCHW image directly (fp16 / fp32 / u8)
// and writing NHWC fp16.  conv.py:69-89 for model.0; fuses the NCHW->NHWC layout change into the load.
// ---------------------------------------------------------------------------------------------
template <typename TIn>
__device__ __forceinline__ float load_px(const TIn* p);
template <>
__device__ __forceinline__ float load_px<__half>(const __half* p) { return __half2float(*p); }
template <>
__device__ __forceinline__ float load_px<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float load_px<unsigned char>(const unsigned char* p) { return (float)(*p) * (1.0f / 255.0f); }

// Weights travel as a __grid_constant__ kernel parameter: every FFMA then takes its weight straight from the constant bank
// (uniform address), so the inner loop is 1 shared-memory load + COUT FFMAs per tap.  (Round-1 history: weights in shared
// memory cost 4 broadcast LDS.128 per tap and made the kernel LSU-bound at 0.63 TB/s, profiles/r01_launch_roofline.txt.)
template <int COUT>
struct StemWeights {
    float w[4 * 9 * COUT];   // [(ci*3+ky)*3+kx][COUT], rows ci >= Cin unused
    float b[COUT];
};

template <typename TIn, int COUT>
__global__ void __launch_bounds__(256) stem_conv_kernel(const TIn* __restrict__ img, int B, int Cin, int H, int W,
                                                        const __grid_constant__ StemWeights<COUT> sw,
                                                        __half* __restrict__ out, int ldo, int Ho, int Wo, int tiles_x) {
    pdl_prologue();
    // CTA = 32 x 8 output pixels; the (65 x 17) x Cin input patch is staged in shared memory one image line per warp pass (lane =
    // column: coalesced, no per-element div / mod - the flat-index loop it replaces spent more instructions on addressing than the
    // 432 FFMAs), even and odd columns in separate planes so that the stride-2 reads of the taps are bank-conflict free.
    constexpr int TW = 32, TH = 8, IW = 2 * TW + 1, IH = 2 * TH + 1;
    __shared__ float se[4][IH][TW + 1 + 3];   // columns rx = 0, 2, ... 64  (taps kx = 0 and kx = 2)
    __shared__ float so[4][IH][TW + 4];       // columns rx = 1, 3, ... 63  (tap kx = 1)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int oy0 = (blockIdx.x / tiles_x) * TH, ox0 = (blockIdx.x % tiles_x) * TW;
    const int iy0 = oy0 * 2 - 1, ix0 = ox0 * 2 - 1;
    for (int line = warp; line < Cin * IH; line += 8) {
        const int ci = line / IH, ry = line - ci * IH;
        const int iy = iy0 + ry;
        const bool row_ok = iy >= 0 && iy < H;
        const TIn* src = img + (((long long)b * Cin + ci) * H + (row_ok ? iy : 0)) * W;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int rx = lane + 32 * j;
            if (rx < IW) {
                const int ix = ix0 + rx;
                const float v = (row_ok && ix >= 0 && ix < W) ? load_px<TIn>(src + ix) : 0.f;
                if (rx & 1) so[ci][ry][rx >> 1] = v;
                else se[ci][ry][rx >> 1] = v;
            }
        }
    }
    __syncthreads();
    const int lx = tid % TW, ly = tid / TW;
    const int ox = ox0 + lx, oy = oy0 + ly;
    if (ox >= Wo || oy >= Ho) return;
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = sw.b[c];
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
        if (ci < Cin) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float v = (kx == 1) ? so[ci][2 * ly + ky][lx] : se[ci][2 * ly + ky][lx + (kx >> 1)];
#pragma unroll
                    for (int c = 0; c < COUT; ++c) acc[c] = fmaf(v, sw.w[((ci * 3 + ky) * 3 + kx) * COUT + c], acc[c]);
                }
            }
        }
    }
    __half* dst = out + (((long long)b * Ho + oy) * Wo + ox) * ldo;
#pragma unroll
    for (int c = 0; c < COUT; c += 8) {
        Half8 h;
#pragma unroll
        for (int q = 0; q < 4; ++q) h.v[q] = __floats2half2_rn(silu_f(acc[c + 2 * q]), silu_f(acc[c + 2 * q + 1]));
        *reinterpret_cast<Half8*>(dst + c) = h;
    }
}

// ---------------------------------------------------------------------------------------------
// Tensor-core stem (Cin <= 3, Cout = 16): the same 3x3 / stride 2 / pad 1 convolution as an implicit GEMM on mma.sync.m16n8k16.
// The FFMA kernel above spends 432 FFMA + im2col addressing per output pixel (FFMA-issue bound at 11 % of HBM,
// profiles/r01_early_layers_ncu.txt); here a pixel costs 6/16 of an HMMA.  K is laid out as (ci, ky) groups of FOUR taps
// (a zero-weight pad and kx = 0, 1, 2), 9 groups padded to 12 -> K = 48 = three k16 steps, so that every A-fragment register is ONE
// aligned 32-bit shared-memory load of two horizontally adjacent fp16 input pixels (the im2col never exists).  Inputs are rounded to
// fp16 as the reference's own fp16 predictor path does (`im.half()`, then `/ 255` for uint8 frames: engine/predictor.py:173-175);
// the folded fp32 weights are split into two fp16 parts (w = hi + lo, lo = the next 11 mantissa bits) and both are multiplied in - the
// first layer keeps the fp32-weight accuracy of the FFMA kernel for six more HMMAs per 16 pixels; accumulation is fp32.
// CTA = 8 warps = 16 output rows x 64 output columns; warp = two output rows, four 16-pixel groups each.
// K slot order inside a (ci, ky) group is {zero, kx = 0, kx = 1, kx = 2} (see the kernel's patch layout).
struct StemTcWeights {
    uint32_t b[2][3][2][32][2];   // B fragments per (hi | lo part, k step, n tile, lane): {b0, b1} as packed half2
    float bias[16];
};

template <typename TIn>
__device__ __forceinline__ __half stem_to_half(TIn v);
template <>
__device__ __forceinline__ __half stem_to_half<__half>(__half v) { return v; }
template <>
__device__ __forceinline__ __half stem_to_half<float>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __half stem_to_half<unsigned char>(unsigned char v) { return __float2half_rn(__fdiv_rn((float)v, 255.f)); }

template <typename TIn, bool VEC>
__global__ void __launch_bounds__(256) stem_conv_tc_kernel(const TIn* __restrict__ img, int B, int Cin, int H, int W,
                                                           const __grid_constant__ StemTcWeights sw, __half* __restrict__ out, int ldo,
                                                           int Ho, int Wo, int tiles_x) {
    pdl_prologue();
    // patch column j <-> image column 2 * ox0 - 2 + j: the four K slots of a (ci, ky) group are {zero weight, kx = 0, kx = 1, kx = 2}, so
    // both fragment pairs of an output pixel - columns (2 lx, 2 lx + 1) and (2 lx + 2, 2 lx + 3) - are aligned 32-bit words AND start on
    // an even image column: fp16 images with even W are staged with 32-bit loads and stores.  Row stride 144 halves = 72 words (8 mod 32).
    constexpr int TW = 64, TH = 16, IW = 2 * TW + 4, IH = 2 * TH + 1, RS = 144;
    __shared__ __align__(16) __half sx[3][IH][RS];
    __shared__ __align__(16) __half sout[8][16][16];                                 // per warp: 16 pixels x 16 channels staging
    __shared__ __align__(8) uint32_t sbf[12 * 32 * 2];                               // B fragments, [(part, ks, nt)][lane][2]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int oy0 = (blockIdx.x / tiles_x) * TH, ox0 = (blockIdx.x % tiles_x) * TW;
    const int iy0 = oy0 * 2 - 1, ix0 = ox0 * 2 - 2;
    // B fragments: kernel parameter -> shared memory once per CTA (3 words per thread), then ONE conflict-free 8-byte read per fragment.
    // Read straight from the parameter bank they are lane-indexed constant loads, which the constant cache serves one address at a time:
    // 12 LDC.64 x 32 lanes per warp were 42 % of the kernel's stall samples (profiles/r02_moe_stem_ncu.txt).
    {
        const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(&sw.b[0][0][0][0][0]);
#pragma unroll
        for (int i = 0; i < 3; ++i) sbf[tid + 256 * i] = wsrc[tid + 256 * i];
    }
    // patch load: the 3 x 33 (channel, row) lines are dealt to the 8 warps; the loads of a batch of lines are all issued before the first
    // shared-memory store, with adds as the only index arithmetic (scalar loads with per-element bounds tests executed 38 M of the
    // kernel's 69 M instructions: profiles/r02_small_stem_ncu.txt)
    constexpr int LINES = 3 * IH, LPW = (LINES + 7) / 8, LB = 5;
    if (VEC) {
        constexpr int PPL = IW / 2, CPL = (PPL + 31) / 32;       // 66 pixel pairs per line, 3 per lane
#pragma unroll 1
        for (int l0 = 0; l0 < LPW; l0 += LB) {
            uint32_t vals[LB][CPL];
#pragma unroll
            for (int li = 0; li < LB; ++li) {
                const int line = warp + 8 * (l0 + li);
                const int ci = line / IH, ry = line - ci * IH;
                const int iy = iy0 + ry;
                const bool row_ok = line < LINES && ci < Cin && iy >= 0 && iy < H;
                const TIn* src = img + (((long long)b * Cin + (row_ok ? ci : 0)) * H + (row_ok ? iy : 0)) * W;
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    const int m = lane + 32 * k, ix = ix0 + 2 * m;          // even: the pair is inside or outside the image as a whole
                    uint32_t v = 0u;
                    if (row_ok && m < PPL && ix >= 0 && ix < W) v = *reinterpret_cast<const uint32_t*>(src + ix);
                    vals[li][k] = v;
                }
            }
#pragma unroll
            for (int li = 0; li < LB; ++li) {
                const int line = warp + 8 * (l0 + li);
                if (line < LINES) {
                    const int ci = line / IH, ry = line - ci * IH;
#pragma unroll
                    for (int k = 0; k < CPL; ++k) {
                        const int m = lane + 32 * k;
                        if (m < PPL) *reinterpret_cast<uint32_t*>(&sx[ci][ry][2 * m]) = vals[li][k];
                    }
                }
            }
        }
    } else {
        constexpr int CPL = (IW + 31) / 32;
#pragma unroll 1
        for (int l0 = 0; l0 < LPW; l0 += LB) {
            __half vals[LB][CPL];
#pragma unroll
            for (int li = 0; li < LB; ++li) {
                const int line = warp + 8 * (l0 + li);
                const int ci = line / IH, ry = line - ci * IH;
                const int iy = iy0 + ry;
                const bool row_ok = line < LINES && ci < Cin && iy >= 0 && iy < H;
                const TIn* src = img + (((long long)b * Cin + (row_ok ? ci : 0)) * H + (row_ok ? iy : 0)) * W;
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    const int rx = lane + 32 * k, ix = ix0 + rx;
                    __half v = __float2half_rn(0.f);
                    if (row_ok && rx < IW && ix >= 0 && ix < W) v = stem_to_half<TIn>(src[ix]);
                    vals[li][k] = v;
                }
            }
#pragma unroll
            for (int li = 0; li < LB; ++li) {
                const int line = warp + 8 * (l0 + li);
                if (line < LINES) {
                    const int ci = line / IH, ry = line - ci * IH;
#pragma unroll
                    for (int k = 0; k < CPL; ++k) {
                        const int rx = lane + 32 * k;
                        if (rx < IW) sx[ci][ry][rx] = vals[li][k];
                    }
                }
            }
        }
    }
    __syncthreads();
    const int g = lane >> 2, t = lane & 3;
    uint32_t bf[2][3][2][2];
#pragma unroll
    for (int part = 0; part < 2; ++part)
#pragma unroll
        for (int ks = 0; ks < 3; ++ks)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const uint2 w2 = *reinterpret_cast<const uint2*>(&sbf[((((part * 3 + ks) * 2 + nt) * 32) + lane) * 2]);
                bf[part][ks][nt][0] = w2.x;
                bf[part][ks][nt][1] = w2.y;
            }
    const float bias0 = sw.bias[2 * t], bias1 = sw.bias[2 * t + 1], bias8 = sw.bias[8 + 2 * t], bias9 = sw.bias[8 + 2 * t + 1];
#pragma unroll 1
    for (int it = 0; it < 2 * (TW / 16); ++it) {
        const int wr = 2 * warp + (it >> 2), grp = it & 3;   // tile row of this warp (two per warp), 16-pixel group
        const int oy = oy0 + wr;
        if (oy >= Ho) break;
        const int lx = grp * 16 + g;                      // output column (tile-local) of fragment row g; row g+8 = lx + 8
        float acc[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[nt][q] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
            // k local 2t, 2t+1 -> tap group 4*ks + (t >> 1), pair (t & 1); k local 2t+8, 2t+9 -> group 4*ks + 2 + (t >> 1), same pair
            uint32_t a[4];
#pragma unroll
            for (int hi = 0; hi < 2; ++hi) {
                const int grpk = 4 * ks + 2 * hi + (t >> 1);          // (ci, ky) group, 9..11 are zero-weight pads
                const int ci = grpk < 9 ? grpk / 3 : 0, ky = grpk < 9 ? grpk - 3 * (grpk / 3) : 0;
                const __half* row = &sx[ci][2 * wr + ky][2 * (t & 1)];
                a[2 * hi + 0] = *reinterpret_cast<const uint32_t*>(row + 2 * lx);            // fragment row g
                a[2 * hi + 1] = *reinterpret_cast<const uint32_t*>(row + 2 * (lx + 8));      // fragment row g + 8
            }
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                mma_16816(acc[nt], a, bf[1][ks][nt][0], bf[1][ks][nt][1]);      // low part first: small terms enter the fp32 sum early
                mma_16816(acc[nt], a, bf[0][ks][nt][0], bf[0][ks][nt][1]);
            }
        }
        // fragments -> per-warp staging tile -> one 16-byte store per lane (16 pixels x 32 B are contiguous in the NHWC output)
        __half2* so = reinterpret_cast<__half2*>(&sout[warp][0][0]);
        so[g * 8 + t] = __floats2half2_rn(silu_f(acc[0][0] + bias0), silu_f(acc[0][1] + bias1));
        so[g * 8 + 4 + t] = __floats2half2_rn(silu_f(acc[1][0] + bias8), silu_f(acc[1][1] + bias9));
        so[(g + 8) * 8 + t] = __floats2half2_rn(silu_f(acc[0][2] + bias0), silu_f(acc[0][3] + bias1));
        so[(g + 8) * 8 + 4 + t] = __floats2half2_rn(silu_f(acc[1][2] + bias8), silu_f(acc[1][3] + bias9));
        __syncwarp();
        const int px = lane >> 1, hf = lane & 1;
        const int ox = ox0 + grp * 16 + px;
        if (ox < Wo)
            *reinterpret_cast<Half8*>(out + (((long long)b * Ho + oy) * Wo + ox) * ldo + hf * 8) =
                *reinterpret_cast<const Half8*>(&sout[warp][px][hf * 8]);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Depthwise k x k (stride 1, pad k/2) + bias (+SiLU) (+add), 8 channels per thread.
// Source channel for output channel c: (c / grp_w) * grp_stride + grp_off + c % grp_w  -- lets `pe(v)` read V in place
// from the head-interleaved qkv tensor (block.py:1688,1731; :1311,1331).  Weights tap-major [k*k][C] fp16.
// ---------------------------------------------------------------------------------------------
template <int KS>
__global__ void __launch_bounds__(256) dwconv_kernel(const __half* __restrict__ x, int ldx, int grp_w, int grp_stride,
                                                     int grp_off, const __half* __restrict__ w, const float* __restrict__ bias,
                                                     int B, int H, int W, int C, int act, const __half* __restrict__ add,
                                                     int ldadd, __half* __restrict__ out, int ldo) {
    pdl_prologue();
    const int chunks = C >> 3;
    const long long total = (long long)B * H * W * chunks;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ch = (int)(idx % chunks);
    const long long pix = idx / chunks;
    const int ox = (int)(pix % W);
    const int oy = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    const int c = ch * 8;
    const int csrc = (c / grp_w) * grp_stride + grp_off + (c % grp_w);
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = bias ? bias[c + q] : 0.f;
    constexpr int R = KS / 2;
    const __half* xb = x + (long long)b * H * W * ldx + csrc;
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
        const int iy = oy + ky - R;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
            const int ix = ox + kx - R;
            if (ix < 0 || ix >= W) continue;
            const Half8 xv = *reinterpret_cast<const Half8*>(xb + ((long long)iy * W + ix) * ldx);
            const Half8 wv = *reinterpret_cast<const Half8*>(w + (ky * KS + kx) * C + c);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 xf = __half22float2(xv.v[q]);
                const float2 wf = __half22float2(wv.v[q]);
                acc[2 * q] = fmaf(xf.x, wf.x, acc[2 * q]);
                acc[2 * q + 1] = fmaf(xf.y, wf.y, acc[2 * q + 1]);
            }
        }
    }
    if (act == 1) {
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = silu_f(acc[q]);
    }
    if (add != nullptr) {
        const Half8 av = *reinterpret_cast<const Half8*>(add + pix * ldadd + c);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 af = __half22float2(av.v[q]);
            acc[2 * q] += af.x;
            acc[2 * q + 1] += af.y;
        }
    }
    Half8 o;
#pragma unroll
    for (int q = 0; q < 4; ++q) o.v[q] = __floats2half2_rn(acc[2 * q], acc[2 * q + 1]);
    *reinterpret_cast<Half8*>(out + pix * ldo + c) = o;
}

// Shared-memory tiled variant (the fast path): a CTA computes TH x TW output pixels x CB channels from a haloed input
// tile staged once in shared memory; each thread produces PX=4 horizontally adjacent pixels x 8 channels with a sliding
// window in registers, so every staged input vector is reused up to min(KS,PX) times per row.  HBM traffic = one read of
// the (haloed) input + one write of the output; the k x k re-reads are served from shared memory.
template <int KS, int CHUNKS>
__global__ void __launch_bounds__(256) dwconv_tiled_kernel(const __half* __restrict__ x, int ldx, int grp_w, int grp_stride,
                                                           int grp_off, const __half* __restrict__ w,
                                                           const float* __restrict__ bias, int H, int W, int C, int act,
                                                           const __half* __restrict__ add, int ldadd, __half* __restrict__ out,
                                                           int ldo, int tiles_x, const int* __restrict__ route_idx, int topk,
                                                           int expert) {
    pdl_prologue();
    constexpr int R = KS / 2, PX = 4, CB = CHUNKS * 8;
    constexpr int GROUPS = 256 / CHUNKS;           // pixel groups per CTA
    constexpr int TW = (CHUNKS == 8) ? 16 : 32;    // tile width
    constexpr int GX = TW / PX;                    // groups per row
    constexpr int TH = GROUPS / GX;                // tile height (8 or 16)
    constexpr int HW_ = TW + 2 * R, HH_ = TH + 2 * R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sx = reinterpret_cast<__half*>(smem_raw);            // [HH_][HW_][CB]
    __half* sw = sx + HH_ * HW_ * CB;                            // [KS*KS][CB]
    const int tid = threadIdx.x;
    const int tile = blockIdx.x, b = blockIdx.y, cb0 = blockIdx.z * CB;
    // routed mode (ES_MOE experts): image b is processed only if `expert` is among its retained routes; the output goes to
    // slot b*topk + j of the per-route scratch tensor
    int ob = b;
    if (route_idx != nullptr) {
        int j = -1;
        for (int q = 0; q < topk; ++q)
            if (route_idx[b * topk + q] == expert) j = q;
        if (j < 0) return;
        ob = b * topk + j;
    }
    const int ty0 = (tile / tiles_x) * TH, tx0 = (tile % tiles_x) * TW;
    const __half* xb = x + (long long)b * H * W * ldx;
    // ---- stage weights + haloed input tile (zero outside the image)
    for (int i = tid; i < KS * KS * CHUNKS; i += 256) {
        const int tap = i / CHUNKS, ch = i % CHUNKS;
        *reinterpret_cast<Half8*>(sw + tap * CB + ch * 8) = *reinterpret_cast<const Half8*>(w + tap * C + cb0 + ch * 8);
    }
    for (int i = tid; i < HH_ * HW_ * CHUNKS; i += 256) {
        const int ch = i % CHUNKS, pp = i / CHUNKS;
        const int hy = pp / HW_, hx = pp % HW_;
        const int iy = ty0 + hy - R, ix = tx0 + hx - R;
        const int c = cb0 + ch * 8;
        const int csrc = (c / grp_w) * grp_stride + grp_off + (c % grp_w);
        const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
        cp_async16(sx + pp * CB + ch * 8, ok ? xb + ((long long)iy * W + ix) * ldx + csrc : xb, ok ? 16 : 0);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    // ---- compute
    const int ch = tid % CHUNKS, grp = tid / CHUNKS;
    const int gy = grp / GX, gx = grp % GX;
    const int c = cb0 + ch * 8;
    float acc[PX][8];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[p][q] = bias ? bias[c + q] : 0.f;
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
        float wr[KS][8];
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
            const Half8 wv = *reinterpret_cast<const Half8*>(sw + (ky * KS + kx) * CB + ch * 8);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 f = __half22float2(wv.v[q]);
                wr[kx][2 * q] = f.x;
                wr[kx][2 * q + 1] = f.y;
            }
        }
        const __half* row = sx + ((gy + ky) * HW_ + gx * PX) * CB + ch * 8;
#pragma unroll
        for (int j = 0; j < PX + KS - 1; ++j) {
            const Half8 xv = *reinterpret_cast<const Half8*>(row + j * CB);
            float xf[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 f = __half22float2(xv.v[q]);
                xf[2 * q] = f.x;
                xf[2 * q + 1] = f.y;
            }
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                const int kx = j - p;
                if (kx >= 0 && kx < KS) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc[p][q] = fmaf(xf[q], wr[kx][q], acc[p][q]);
                }
            }
        }
    }
    // ---- epilogue
    const int oy = ty0 + gy;
    if (oy >= H) return;
#pragma unroll
    for (int p = 0; p < PX; ++p) {
        const int ox = tx0 + gx * PX + p;
        if (ox >= W) continue;
        const long long pix = ((long long)ob * H + oy) * W + ox;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = act == 1 ? silu_f(acc[p][q]) : acc[p][q];
        if (add != nullptr) {
            const Half8 av = *reinterpret_cast<const Half8*>(add + pix * ldadd + c);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 af = __half22float2(av.v[q]);
                v[2 * q] += af.x;
                v[2 * q + 1] += af.y;
            }
        }
        Half8 o;
#pragma unroll
        for (int q = 0; q < 4; ++q) o.v[q] = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
        *reinterpret_cast<Half8*>(out + pix * ldo + c) = o;
    }
}

// ---------------------------------------------------------------------------------------------
// SPPF: three chained MaxPool(k, s1, p=k/2) of y0 == max over (k), (2k-1), (3k-2) windows of y0 (block.py:237-242).
// y0 lives in channel slot 0 of the concat buffer; slots 1..3 are written here.  8 channels per thread.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sppf_pool_kernel(__half* __restrict__ buf, int ld, int B, int H, int W, int C, int k) {
    pdl_prologue();
    const int chunks = C >> 3;
    const long long total = (long long)B * H * W * chunks;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ch = (int)(idx % chunks);
    const long long pix = idx / chunks;
    const int ox = (int)(pix % W);
    const int oy = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    const int r1 = k / 2, r2 = 2 * r1, r3 = 3 * r1;
    __half2 m1[4], m2[4], m3[4];
    const __half2 ninf = __float2half2_rn(-65504.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) m1[q] = m2[q] = m3[q] = ninf;
    const __half* xb = buf + (long long)b * H * W * ld + ch * 8;
    for (int dy = -r3; dy <= r3; ++dy) {
        const int iy = oy + dy;
        if (iy < 0 || iy >= H) continue;
        const int ady = dy < 0 ? -dy : dy;
        for (int dx = -r3; dx <= r3; ++dx) {
            const int ix = ox + dx;
            if (ix < 0 || ix >= W) continue;
            const int adx = dx < 0 ? -dx : dx;
            const int rad = ady > adx ? ady : adx;
            const Half8 v = *reinterpret_cast<const Half8*>(xb + ((long long)iy * W + ix) * ld);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                m3[q] = __hmax2(m3[q], v.v[q]);
                if (rad <= r2) m2[q] = __hmax2(m2[q], v.v[q]);
                if (rad <= r1) m1[q] = __hmax2(m1[q], v.v[q]);
            }
        }
    }
    __half* ob = buf + pix * ld + ch * 8;
    Half8 o;
#pragma unroll
    for (int q = 0; q < 4; ++q) o.v[q] = m1[q];
    *reinterpret_cast<Half8*>(ob + C) = o;
#pragma unroll
    for (int q = 0; q < 4; ++q) o.v[q] = m2[q];
    *reinterpret_cast<Half8*>(ob + 2 * C) = o;
#pragma unroll
    for (int q = 0; q < 4; ++q) o.v[q] = m3[q];
    *reinterpret_cast<Half8*>(ob + 3 * C) = o;
}

// ---------------------------------------------------------------------------------------------
// out[..., 0:Ca] = a (optionally nearest-upsampled by `up`), out[..., Ca:Ca+Cb] = b     (nn.Upsample + Concat)
// ---------------------------------------------------------------------------------------------
// Four 16-byte items per thread, all four loads issued before the first store (one item per thread with 64-bit div / mod moved 1.8 TB/s:
// profiles/r02_launch_roofline.txt, r02g); 32-bit index arithmetic (the host checks the item count fits).
__global__ void __launch_bounds__(256) concat2_kernel(const __half* __restrict__ a, int lda, int Ca, int up,
                                                      const __half* __restrict__ bsrc, int ldb, int Cb,
                                                      __half* __restrict__ out, int ldo, int B, int H, int W) {
    pdl_prologue();
    const unsigned chunks = (unsigned)(Ca + Cb) >> 3, ca_chunks = (unsigned)Ca >> 3;
    const unsigned total = (unsigned)B * H * W * chunks;
    const unsigned base = blockIdx.x * 1024u + threadIdx.x;
    uint4 v[4];
    unsigned pix[4], ch[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned idx = base + j * 256u;
        v[j] = make_uint4(0u, 0u, 0u, 0u);
        pix[j] = idx / chunks;
        ch[j] = idx - pix[j] * chunks;
        if (idx < total) {
            if (ch[j] < ca_chunks) {
                unsigned sp = pix[j];
                if (up > 1) {
                    const unsigned row = pix[j] / (unsigned)W, ox = pix[j] - row * (unsigned)W;       // row = b * H + oy
                    const unsigned b = row / (unsigned)H, oy = row - b * (unsigned)H;
                    sp = (b * (unsigned)(H / up) + oy / (unsigned)up) * (unsigned)(W / up) + ox / (unsigned)up;
                }
                v[j] = *reinterpret_cast<const uint4*>(a + (long long)sp * lda + ch[j] * 8);
            } else {
                v[j] = *reinterpret_cast<const uint4*>(bsrc + (long long)pix[j] * ldb + (ch[j] - ca_chunks) * 8);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (base + j * 256u < total) *reinterpret_cast<uint4*>(out + (long long)pix[j] * ldo + ch[j] * 8) = v[j];
}

}  // namespace ym

namespace ym {

// Separable, chained variant (the fast path for the P5-sized maps SPPF sees): one CTA owns the H x W plane of 8 channels of
// one image in shared memory and applies MaxPool(k) three times, each as a row pass and a column pass (5 + 5 shared-memory
// reads per pixel per stage instead of the 169 global reads per pixel of the direct (3k-2)^2 window above).
__device__ __forceinline__ Half8 hmax8(const Half8& a, const Half8& b) {
    Half8 r;
#pragma unroll
    for (int q = 0; q < 4; ++q) r.v[q] = __hmax2(a.v[q], b.v[q]);
    return r;
}

__global__ void __launch_bounds__(256) sppf_pool_plane_kernel(__half* __restrict__ buf, int ld, int H, int W, int C, int k) {
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char sppf_smem[];
    Half8* cur = reinterpret_cast<Half8*>(sppf_smem);   // [H*W]
    Half8* tmp = cur + H * W;                           // [H*W]
    const int ch = blockIdx.x, b = blockIdx.y, HW = H * W, r = k / 2;
    __half* base = buf + (long long)b * HW * ld + ch * 8;
    for (int p = threadIdx.x; p < HW; p += blockDim.x) cur[p] = *reinterpret_cast<const Half8*>(base + (long long)p * ld);
    __syncthreads();
    for (int stage = 1; stage <= 3; ++stage) {
        for (int p = threadIdx.x; p < HW; p += blockDim.x) {      // row pass
            const int y = p / W, x = p - y * W;
            Half8 m = cur[p];
            for (int dx = -r; dx <= r; ++dx) {
                const int xx = x + dx;
                if (xx >= 0 && xx < W) m = hmax8(m, cur[y * W + xx]);
            }
            tmp[p] = m;
        }
        __syncthreads();
        for (int p = threadIdx.x; p < HW; p += blockDim.x) {      // column pass, result is the next stage's input
            const int y = p / W, x = p - y * W;
            Half8 m = tmp[p];
            for (int dy = -r; dy <= r; ++dy) {
                const int yy = y + dy;
                if (yy >= 0 && yy < H) m = hmax8(m, tmp[yy * W + x]);
            }
            *reinterpret_cast<Half8*>(base + (long long)p * ld + stage * C) = m;
            cur[p] = m;                                           // safe: the row pass of this stage has finished reading cur
        }
        __syncthreads();
    }
}

}  // namespace ym

using namespace ym;

static inline int nblocks(long long total, int bs) { return (int)((total + bs - 1) / bs); }

// 1 = tensor-core stem (mma.sync implicit GEMM, fp16 operands) where Cin <= 3 and Cout = 16; 0 = fp32 FFMA kernel (A/B baseline)
static int g_stem_impl = 1;   // 1 = mma.sync kernel (112 us at bs32 / 640, r02i), 0 = FFMA kernel (176 us)
extern "C" int ym_set_stem_impl(int impl) {
    const int old = g_stem_impl;
    if (impl == 0 || impl == 1) g_stem_impl = impl;
    return old;
}

// in_dtype: 0 = fp16 NCHW, 1 = fp32 NCHW, 2 = uint8 NCHW (scaled by 1/255).  wgt_host / bias_host are HOST pointers (fp32
// [Cin*9][Cout] and [Cout]): they are copied into the kernel's parameter block (constant bank) at launch, i.e. at capture
// time under a CUDA graph.
template <int CO>
static int stem_launch(const void* img, int in_dtype, int B, int Cin, int H, int W, const float* wgt_host, const float* bias_host,
                       void* out, int ldo, int Ho, int Wo, cudaStream_t st) {
    StemWeights<CO> sw;
    memset(&sw, 0, sizeof(sw));
    memcpy(sw.w, wgt_host, sizeof(float) * (size_t)Cin * 9 * CO);
    memcpy(sw.b, bias_host, sizeof(float) * CO);
    if (CO == 16 && Cin <= 3 && g_stem_impl == 1) {
        // tensor-core path: fragment-ordered fp16 weights (K = (ci, ky) groups of four taps, see stem_conv_tc_kernel)
        StemTcWeights tw;
        memset(&tw, 0, sizeof(tw));
        auto wk = [&](int k, int n) -> float {          // weight of GEMM row k (0..47), output channel n
            const int grp = k >> 2, kx = (k & 3) - 1;          // slots of a (ci, ky) group: {zero, kx = 0, kx = 1, kx = 2}
            if (grp >= 9 || kx < 0) return 0.f;
            const int ci = grp / 3, ky = grp % 3;
            if (ci >= Cin) return 0.f;
            return wgt_host[(size_t)((ci * 3 + ky) * 3 + kx) * CO + n];
        };
        auto part_of = [&](float w, int part) -> float {       // part 0: fp16(w), part 1: fp16(w - fp16(w))
            const float hi = __half2float(__float2half_rn(w));
            return part == 0 ? hi : (w - hi);
        };
        for (int part = 0; part < 2; ++part)
            for (int ks = 0; ks < 3; ++ks)
                for (int nt = 0; nt < 2; ++nt)
                    for (int lane = 0; lane < 32; ++lane) {
                        const int g = lane >> 2, t = lane & 3, n = nt * 8 + g;
                        const __half2 b0 = __floats2half2_rn(part_of(wk(16 * ks + 2 * t, n), part), part_of(wk(16 * ks + 2 * t + 1, n), part));
                        const __half2 b1 = __floats2half2_rn(part_of(wk(16 * ks + 2 * t + 8, n), part), part_of(wk(16 * ks + 2 * t + 9, n), part));
                        memcpy(&tw.b[part][ks][nt][lane][0], &b0, 4);
                        memcpy(&tw.b[part][ks][nt][lane][1], &b1, 4);
                    }
        memcpy(tw.bias, bias_host, sizeof(float) * 16);
        const int tx = (Wo + 63) / 64, ty = (Ho + 15) / 16;
        const dim3 gridt(tx * ty, B);
        const bool vec = in_dtype == 0 && W % 2 == 0 && ((uintptr_t)img & 3) == 0;      // every image row then starts on a 4-byte boundary
        if (vec) launch_pdl(stem_conv_tc_kernel<__half, true>, gridt, 256, 0, st, (const __half*)img, B, Cin, H, W, tw, (__half*)out, ldo, Ho, Wo, tx);
        else if (in_dtype == 0) launch_pdl(stem_conv_tc_kernel<__half, false>, gridt, 256, 0, st, (const __half*)img, B, Cin, H, W, tw, (__half*)out, ldo, Ho, Wo, tx);
        else if (in_dtype == 1) launch_pdl(stem_conv_tc_kernel<float, false>, gridt, 256, 0, st, (const float*)img, B, Cin, H, W, tw, (__half*)out, ldo, Ho, Wo, tx);
        else if (in_dtype == 2) launch_pdl(stem_conv_tc_kernel<unsigned char, false>, gridt, 256, 0, st, (const unsigned char*)img, B, Cin, H, W, tw, (__half*)out, ldo, Ho, Wo, tx);
        else { ym_set_error("ym_stem_conv_nchw: bad in_dtype %d", in_dtype); return YM_ERR_ARG; }
        return YM_OK;
    }
    const int tiles_x = (Wo + 31) / 32, tiles_y = (Ho + 7) / 8;
    const dim3 grid(tiles_x * tiles_y, B);
    if (in_dtype == 0) launch_pdl(stem_conv_kernel<__half, CO>, grid, 256, 0, st, (const __half*)img, B, Cin, H, W, sw, (__half*)out, ldo, Ho, Wo, tiles_x);
    else if (in_dtype == 1) launch_pdl(stem_conv_kernel<float, CO>, grid, 256, 0, st, (const float*)img, B, Cin, H, W, sw, (__half*)out, ldo, Ho, Wo, tiles_x);
    else if (in_dtype == 2) launch_pdl(stem_conv_kernel<unsigned char, CO>, grid, 256, 0, st, (const unsigned char*)img, B, Cin, H, W, sw, (__half*)out, ldo, Ho, Wo, tiles_x);
    else { ym_set_error("ym_stem_conv_nchw: bad in_dtype %d", in_dtype); return YM_ERR_ARG; }
    return YM_OK;
}

extern "C" int ym_stem_conv_nchw(const void* img, int in_dtype, int B, int Cin, int H, int W, const float* wgt_host,
                                 const float* bias_host, int Cout, void* out, int ldo, void* stream) {
    YM_CHECK_ARG(img && wgt_host && bias_host && out, "ym_stem_conv_nchw: null pointer");
    YM_CHECK_ARG(Cin >= 1 && Cin <= 4, "ym_stem_conv_nchw: Cin must be 1..4 (got %d)", Cin);
    YM_CHECK_ARG(ldo % 8 == 0 && ((uintptr_t)out & 15) == 0, "ym_stem_conv_nchw: output alignment");
    if (B == 0) return YM_OK;
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    cudaStream_t st = (cudaStream_t)stream;
    YM_CHECK_ARG(B <= 65535, "ym_stem_conv_nchw: batch too large");
    int rc;
    if (Cout == 16) rc = stem_launch<16>(img, in_dtype, B, Cin, H, W, wgt_host, bias_host, out, ldo, Ho, Wo, st);
    else if (Cout == 32) rc = stem_launch<32>(img, in_dtype, B, Cin, H, W, wgt_host, bias_host, out, ldo, Ho, Wo, st);
    else if (Cout == 64) rc = stem_launch<64>(img, in_dtype, B, Cin, H, W, wgt_host, bias_host, out, ldo, Ho, Wo, st);
    else { ym_set_error("ym_stem_conv_nchw: Cout must be 16/32/64 (got %d)", Cout); return YM_ERR_UNSUPPORTED; }
    if (rc) return rc;
    YM_CHECK_LAUNCH("stem_conv");
    return YM_OK;
}

static int dwconv_impl(const void* x, int ldx, int grp_w, int grp_stride, int grp_off, const void* w, const float* bias, int B,
                       int H, int W, int C, int ksize, int act, const void* add, int ldadd, void* out, int ldo,
                       const int* route_idx, int topk, int expert, void* stream);
namespace ym {   // dwconv_tc.cu: depthwise 7x7 as Toeplitz GEMMs on mma.sync
int dwconv7_tc_supported(int C, int ksize, int grp_w, int B, const void* route_idx);
int dwconv7_tc_run(const void* x, int ldx, int grp_w, int grp_stride, int grp_off, const void* w, const float* bias, int B, int H, int W, int C,
                   int act, const void* add, int ldadd, void* out, int ldo, cudaStream_t st);
}

extern "C" int ym_dwconv_nhwc(const void* x, int ldx, int grp_w, int grp_stride, int grp_off, const void* w,
                              const float* bias, int B, int H, int W, int C, int ksize, int act, const void* add, int ldadd,
                              void* out, int ldo, void* stream) {
    return dwconv_impl(x, ldx, grp_w, grp_stride, grp_off, w, bias, B, H, W, C, ksize, act, add, ldadd, out, ldo, nullptr, 0, 0, stream);
}

// ES_MOE expert depthwise stage (experts.py:284): image b is convolved with expert `expert`'s k x k filter only if that
// expert is one of its retained routes; the result lands in slot b*topk + j of out ([B*topk][H][W][C]).
extern "C" int ym_esmoe_dwconv(const void* x, int ldx, const void* w, int B, int H, int W, int C, int ksize, const int* route_idx,
                               int topk, int expert, void* out, int ldo, void* stream) {
    YM_CHECK_ARG(route_idx && topk >= 1, "ym_esmoe_dwconv: routing table required");
    YM_CHECK_ARG(C % 16 == 0, "ym_esmoe_dwconv: C must be a multiple of 16");
    return dwconv_impl(x, ldx, C, C, 0, w, nullptr, B, H, W, C, ksize, 0, nullptr, 0, out, ldo, route_idx, topk, expert, stream);
}

static int dwconv_impl(const void* x, int ldx, int grp_w, int grp_stride, int grp_off, const void* w, const float* bias, int B,
                       int H, int W, int C, int ksize, int act, const void* add, int ldadd, void* out, int ldo,
                       const int* route_idx, int topk, int expert, void* stream) {
    YM_CHECK_ARG(x && w && out, "ym_dwconv_nhwc: null pointer");
    YM_CHECK_ARG(C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && grp_w % 8 == 0 && grp_stride % 8 == 0 && grp_off % 8 == 0,
                 "ym_dwconv_nhwc: channel counts/pitches must be multiples of 8");
    YM_CHECK_ARG(add == nullptr || ldadd % 8 == 0, "ym_dwconv_nhwc: add pitch");
    YM_CHECK_ARG((((uintptr_t)x | (uintptr_t)w | (uintptr_t)out | (uintptr_t)add) & 15) == 0, "ym_dwconv_nhwc: alignment");
    if (B == 0) return YM_OK;
    const long long total = (long long)B * H * W * (C / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (ym::dwconv7_tc_supported(C, ksize, grp_w, B, route_idx)) {
        const int rc = ym::dwconv7_tc_run(x, ldx, grp_w, grp_stride, grp_off, w, bias, B, H, W, C, act, add, ldadd, out, ldo, st);
        if (rc) return rc;
        YM_CHECK_LAUNCH("dwconv7_tc");
        return YM_OK;
    }
    const int chunks = (C % 64 == 0) ? 8 : ((C % 16 == 0) ? 2 : 0);  // source mapping is per 8-channel chunk
    if (chunks && B <= 65535) {
        const int TW = chunks == 8 ? 16 : 32, TH = chunks == 8 ? 8 : 16, CB = chunks * 8, R = ksize / 2;
        const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
        const size_t smem = ((size_t)(TH + 2 * R) * (TW + 2 * R) * CB + (size_t)ksize * ksize * CB) * sizeof(__half);
        dim3 grid(tiles_x * tiles_y, B, C / CB);
#define YM_DWT(KS, CH)                                                                                                        \
    do {                                                                                                                      \
        auto kern = dwconv_tiled_kernel<KS, CH>;                                                                              \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);             \
        launch_pdl(kern, grid, 256, smem, st, (const __half*)x, ldx, grp_w, grp_stride, grp_off, (const __half*)w, bias, H, W, C, act, \
                                      (const __half*)add, ldadd, (__half*)out, ldo, tiles_x, route_idx, topk, expert);        \
    } while (0)
        bool done = true;
        if (chunks == 8) {
            switch (ksize) { case 3: YM_DWT(3, 8); break; case 5: YM_DWT(5, 8); break; case 7: YM_DWT(7, 8); break;
                             case 9: YM_DWT(9, 8); break; default: done = false; }
        } else {
            switch (ksize) { case 3: YM_DWT(3, 2); break; case 5: YM_DWT(5, 2); break; case 7: YM_DWT(7, 2); break;
                             case 9: YM_DWT(9, 2); break; default: done = false; }
        }
#undef YM_DWT
        if (done) { YM_CHECK_LAUNCH("dwconv_tiled"); return YM_OK; }
    }
    YM_CHECK_ARG(route_idx == nullptr, "dwconv: routed mode needs the tiled kernel (C %% 16 == 0, k in 3/5/7/9)");
#define YM_DW(KS)                                                                                                    \
    launch_pdl(dwconv_kernel<KS>, nblocks(total, 256), 256, 0, st, (const __half*)x, ldx, grp_w, grp_stride, grp_off,          \
                                                           (const __half*)w, bias, B, H, W, C, act, (const __half*)add, \
                                                           ldadd, (__half*)out, ldo)
    switch (ksize) {
        case 3: YM_DW(3); break;
        case 5: YM_DW(5); break;
        case 7: YM_DW(7); break;
        case 9: YM_DW(9); break;
        default: ym_set_error("ym_dwconv_nhwc: kernel size %d unsupported (3/5/7/9)", ksize); return YM_ERR_UNSUPPORTED;
    }
#undef YM_DW
    YM_CHECK_LAUNCH("dwconv");
    return YM_OK;
}

extern "C" int ym_sppf_pool_nhwc(void* buf, int ld, int B, int H, int W, int C, int k, void* stream) {
    YM_CHECK_ARG(buf, "ym_sppf_pool_nhwc: null pointer");
    YM_CHECK_ARG(C % 8 == 0 && ld % 8 == 0 && ld >= 4 * C && (k & 1), "ym_sppf_pool_nhwc: bad dims");
    if (B == 0) return YM_OK;
    const size_t plane = (size_t)H * W * 16 * 2;                 // two Half8 planes
    if (plane <= 160 * 1024 && B <= 65535) {
        static size_t attr = 0;
        if (plane > 48 * 1024 && plane > attr) {
            cudaError_t e = cudaFuncSetAttribute(sppf_pool_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane);
            if (e != cudaSuccess) { ym_set_error("ym_sppf_pool_nhwc: smem attr: %s", cudaGetErrorString(e)); return YM_ERR_CUDA; }
            attr = plane;
        }
        launch_pdl(sppf_pool_plane_kernel, dim3(C / 8, B), 256, plane, (cudaStream_t)stream, (__half*)buf, ld, H, W, C, k);
    } else {
        const long long total = (long long)B * H * W * (C / 8);
        launch_pdl(sppf_pool_kernel, nblocks(total, 256), 256, 0, (cudaStream_t)stream, (__half*)buf, ld, B, H, W, C, k);
    }
    YM_CHECK_LAUNCH("sppf_pool");
    return YM_OK;
}

extern "C" int ym_concat2_nhwc(const void* a, int lda, int Ca, int up, const void* b, int ldb, int Cb, void* out, int ldo,
                               int B, int H, int W, void* stream) {
    YM_CHECK_ARG(a && out && (b || Cb == 0), "ym_concat2_nhwc: null pointer");
    YM_CHECK_ARG(Ca % 8 == 0 && Cb % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldo % 8 == 0, "ym_concat2_nhwc: multiples of 8");
    YM_CHECK_ARG(up >= 1 && H % up == 0 && W % up == 0, "ym_concat2_nhwc: bad upsample factor");
    if (B == 0) return YM_OK;
    const long long total = (long long)B * H * W * ((Ca + Cb) / 8);
    YM_CHECK_ARG(total < (1LL << 31), "ym_concat2_nhwc: %lld 16-byte items exceed the 32-bit index range", total);
    launch_pdl(concat2_kernel, nblocks(total, 1024), 256, 0, (cudaStream_t)stream, (const __half*)a, lda, Ca, up, (const __half*)b,
                                                                         ldb, Cb, (__half*)out, ldo, B, H, W);
    YM_CHECK_LAUNCH("concat2");
    return YM_OK;
}
