// Microbenchmarks behind DESIGN.md's attention ceiling: what one SM sustains for
//   (1) tcgen05.ld (tensor-memory reads, 32x32b.x32) with 1 / 4 / 8 / 16 warps,
//   (2) MUFU ex2 with 4 / 8 / 16 warps,
//   (3) both interleaved in the same warps (the softmax inner loop's mix),
// one CTA per SM, clock64 around the loop.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench_tmem.cu -o tools/ubench_tmem.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x)); return y; }

// mode 0: tmem loads only; 1: ex2 only; 2: per iteration one x32 load + 32 ex2 on the PREVIOUS load's values (software pipelined)
template <int MODE>
__global__ void __launch_bounds__(512) k(int iters, long long* cyc, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    uint32_t cur[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) cur[i] = __float_as_uint(-1.f - 0.001f * (float)(threadIdx.x + i));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t nxt[32];
        if (MODE != 1) tmem_ld32(base + 32 * (it & 7), nxt);
        if (MODE != 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += ex2(__uint_as_float(cur[i] + (uint32_t)it));   // the integer add keeps the loop body live
        }
        if (MODE != 1) {
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) cur[i] = (MODE == 2) ? ((nxt[i] & 0x007fffffu) | 0xbf800000u) : (cur[i] ^ nxt[i]);
        }
    }
    const long long t1 = clock64();
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += __uint_as_float(cur[i] & 0x3fffffffu);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(slot), "r"(512u) : "memory");
}

template <int MODE>
static void run(const char* name, int warps, int iters) {
    long long* cyc; float* sink;
    cudaMalloc(&cyc, 148 * sizeof(long long));
    cudaMalloc(&sink, 148 * 512 * sizeof(float));
    k<MODE><<<148, warps * 32>>>(iters, cyc, sink);
    k<MODE><<<148, warps * 32>>>(iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
    const double elems = (double)warps * 32 * 32 * iters;     // 32-bit values loaded / exponentials per SM
    printf("%-22s warps=%2d  cycles=%9.0f  %6.2f values/clk/SM  (%s)\n", name, warps, avg, elems / avg, cudaGetErrorString(e));
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    const int iters = 2000;
    for (int w : {1, 4, 8, 16}) run<0>("tcgen05.ld x32", w, iters);
    for (int w : {4, 8, 16}) run<1>("MUFU ex2", w, iters);
    for (int w : {4, 8, 16}) run<2>("ld x32 + 32 ex2 (pipelined)", w, iters);
    return 0;
}
