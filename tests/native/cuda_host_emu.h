// TEST INFRASTRUCTURE: a minimal CUDA execution model on host threads, so that the .cu translation units written after round 1's GPU
// budget (preproc.cu, gated.cu, nms_large.cu, mix.cu) compile with g++ (-DYM_HOST_EMU -x c++) and their extern "C" entry points -
// argument checks, launch geometry, shared-memory sizing, kernels - run in the GPU-less build container (tests/test_cuda_host_emu.py).
//   * a launch creates blockDim threads once and walks the grid block by block; __syncthreads() is a std::barrier over them;
//   * blockIdx / threadIdx / blockDim / gridDim are thread_local; static __shared__ variables are function statics (blocks run one
//     after the other); dynamic shared memory is a per-launch buffer;
//   * __half is IEEE binary16 via _Float16; the *_rn intrinsics are plain IEEE operations.
// Not covered: warp intrinsics, atomics, TMA / tcgen05 (none of the emulated units use them), timing.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <barrier>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) alignas(n)
#define __shared__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }

// ---- fp16
struct __half {
    _Float16 v;
};
struct alignas(4) __half2 {
    __half x, y;
};
inline float __half2float(__half h) { return (float)h.v; }
inline __half __float2half_rn(float f) { return __half{(_Float16)f}; }
inline __half2 __floats2half2_rn(float a, float b) { return __half2{__float2half_rn(a), __float2half_rn(b)}; }
inline float2 __half22float2(__half2 h) { return {__half2float(h.x), __half2float(h.y)}; }

// ---- intrinsics used by the emulated units
inline float __expf(float v) { return expf(v); }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }

namespace ym_emu {
struct Block {
    std::barrier<>* bar;
    void* smem;
};
inline thread_local Block cur;
inline void sync() { cur.bar->arrive_and_wait(); }
inline void* dyn_smem() { return cur.smem; }

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
    const unsigned nthr = block.x * block.y * block.z;
    std::barrier<> bar(nthr);
    std::vector<unsigned char> smem(smem_bytes + 64);
    std::vector<std::thread> pool;
    pool.reserve(nthr);
    for (unsigned t = 0; t < nthr; ++t)
        pool.emplace_back([&, t] {
            cur = Block{&bar, smem.data()};
            blockDim = block;
            gridDim = grid;
            threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            for (unsigned bz = 0; bz < grid.z; ++bz)
                for (unsigned by = 0; by < grid.y; ++by)
                    for (unsigned bx = 0; bx < grid.x; ++bx) {
                        blockIdx = {bx, by, bz};
                        body();
                        bar.arrive_and_wait();          // the next block reuses the shared memory
                    }
        });
    for (auto& th : pool) th.join();
}
}  // namespace ym_emu

#define __syncthreads() ym_emu::sync()
// launch through a kernel function pointer: kfn<<<grid, block, smem, stream>>>(args...)
#define YM_LAUNCH(kfn, grid, block, smem, stream, ...) ym_emu::launch(dim3(grid), dim3(block), (size_t)(smem), [&] { kfn(__VA_ARGS__); })
#define YM_DYN_SMEM(type, name) type* name = (type*)ym_emu::dyn_smem()
