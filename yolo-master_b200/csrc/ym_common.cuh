// Shared device/host helpers for libym_b200 (sm_100a only).
#pragma once
#ifdef YM_HOST_EMU          // tests/native/cuda_host_emu.h: the CUDA execution model on host threads (GPU-less verification)
#include "cuda_host_emu.h"
#else
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
// kernel launch / dynamic shared memory through macros, so that a translation unit can also be built for the host emulation
#define YM_LAUNCH(kfn, grid, block, smem, stream, ...) kfn<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define YM_DYN_SMEM(type, name) extern __shared__ type name[]
#endif

#define YM_OK 0
#define YM_ERR_ARG 1
#define YM_ERR_CUDA 2
#define YM_ERR_UNSUPPORTED 3

// thread-local last-error string (api.cu)
extern "C" void ym_set_error(const char* fmt, ...);
// programmatic dependent launch switch (api.cu): 1 = kernels that call ym::pdl_prologue() are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the next kernel of the stream is staged while this one drains
extern "C" int ym_pdl_enabled(void);
extern "C" int ym_kernel_priority(void);

#define YM_CHECK_ARG(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            ym_set_error(__VA_ARGS__);          \
            return YM_ERR_ARG;                  \
        }                                       \
    } while (0)

#define YM_CHECK_LAUNCH(name)                                                        \
    do {                                                                             \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            ym_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
            return YM_ERR_CUDA;                                                      \
        }                                                                            \
    } while (0)

namespace ym {

#ifndef YM_HOST_EMU   // inline PTX: real GPU builds only
// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------------------
// A kernel launched through ym::launch_pdl may become resident while its predecessor in the stream is still running.  It must
// therefore call pdl_wait() before it touches global memory (reads of the predecessor's output, and writes - the predecessor may
// still be reading a buffer the allocator hands to this kernel).  pdl_trigger() lets the NEXT kernel of the stream be staged as
// soon as every CTA of this grid has reached it: its CTAs then fill whatever resources this grid's last wave leaves free, run
// their prologue (barrier init, tensor-memory allocation, descriptor prefetch) and park in pdl_wait() until this grid has
// completed and flushed - launch latency and drain / fill bubbles between the ~160 kernels of a forward overlap instead of adding up.
// Without the launch attribute both instructions are no-ops, so eager launches and older call sites stay correct.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
    pdl_trigger();
    pdl_wait();
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte async copy global->shared; src_bytes==0 zero-fills (address must still be valid).
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t& r0, uint32_t& r1, const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3,
                                                  const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(smem_u32(p)));
}

// D(16x8,f32) += A(16x16,f16,row) * B(16x8,f16,col)
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

#endif

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_half2(uint32_t v) {
    __half2 h = *reinterpret_cast<__half2*>(&v);
    return __half22float2(h);
}

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

#ifndef YM_HOST_EMU   // warp shuffles: real GPU builds only
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
#endif

#ifndef YM_HOST_EMU
// kernel<<<grid, block, smem, stream>>>(args...) with the PDL attribute (see pdl_prologue); the kernel MUST call pdl_wait().
// `prio` is the launch priority (cudaLaunchAttributePriority; 0 = default, lower = earlier): launch_pdl uses ym_kernel_priority().
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_prio(int prio, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = ym_pdl_enabled() ? 1 : 0;
    attr[1].id = cudaLaunchAttributePriority;
    attr[1].val.priority = prio;
    cfg.attrs = attr;
    cfg.numAttrs = ym_kernel_priority() != 0 ? 2 : 1;        // feature off: exactly the launch it always was
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    const int kp = ym_kernel_priority();
    return launch_pdl_prio(kp < 0 ? kp : 0, kernel, grid, block, smem, stream, args...);
}
#endif

// Eight fp16 values = one 16-byte memory transaction.  __half2 has user-provided copy operations, so the implicit copy of this struct is
// FOUR member copies: every `*reinterpret_cast<Half8*>(p) = h` / `h = *reinterpret_cast<const Half8*>(p)` in the code base compiled
// to four 32-bit LDG / STG (cuobjdump: 1865 LDG.E, no LDG.E.128 in elementwise.o) - a quarter of the bytes per memory instruction and
// four times the LSU work in every bandwidth-bound kernel.  The explicit copy operations below move the 16 bytes as one uint4.
struct alignas(16) Half8 {
    __half2 v[4];
#ifndef YM_HOST_EMU
    __device__ __forceinline__ Half8() {}
    __device__ __forceinline__ Half8(const Half8& o) { *reinterpret_cast<uint4*>(this) = *reinterpret_cast<const uint4*>(&o); }
    __device__ __forceinline__ Half8& operator=(const Half8& o) {
        *reinterpret_cast<uint4*>(this) = *reinterpret_cast<const uint4*>(&o);
        return *this;
    }
#endif
};

}  // namespace ym
