// Debug event trace (-DTL_TRACE=1 builds only, tools/graph_timeline.py): thread 0 of CTA (0,0) of
// every instrumented kernel appends (tag, %globaltimer) pairs to one host-provided device buffer,
// which gives a cross-kernel timeline of a CUDA-graph replay without a profiler serialising it.
#pragma once

#include <cuda_runtime.h>

#ifndef TL_TRACE
#define TL_TRACE 0
#endif

#if TL_TRACE
namespace tl {
static __device__ unsigned long long *g_trace_buf;
static __device__ unsigned int *g_trace_n;
static __device__ unsigned int g_trace_cap;
__device__ __forceinline__ void trace_stamp(unsigned tag, unsigned thread = 0) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == thread && g_trace_buf != nullptr) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        const unsigned i = atomicAdd(g_trace_n, 1u);
        if (i < g_trace_cap) g_trace_buf[2 * i] = tag, g_trace_buf[2 * i + 1] = t;
    }
}
static void trace_bind(unsigned long long *buf, unsigned int *n, unsigned int cap) {  // one copy per translation unit
    cudaMemcpyToSymbol(g_trace_buf, &buf, sizeof(buf));
    cudaMemcpyToSymbol(g_trace_n, &n, sizeof(n));
    cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap));
}
}  // namespace tl
#define TL_TRACE_STAMP(tag) ::tl::trace_stamp(tag)
#define TL_TRACE_STAMP_T(tag, thread) ::tl::trace_stamp(tag, thread)
#else
#define TL_TRACE_STAMP(tag) do { } while (0)
#define TL_TRACE_STAMP_T(tag, thread) do { } while (0)
#endif
