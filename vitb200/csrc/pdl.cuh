// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start — CTA by
// CTA, as SMs drain — while the kernel in front of it on the stream is still in its tail, run its prologue (barrier init, TMEM
// allocation, descriptor prefetch, index arithmetic) and then block in griddepcontrol.wait until that kernel has completed and
// its writes are visible.  Every kernel launched through launch_pdl() executes pdl_wait() before its first global-memory access
// that could depend on (or race with) an earlier kernel; pdl_launch_dependents() at its start lets ITS successor do the same.
// The ~250 launches of a training step are persistent one-wave kernels, so what this hides is launch latency + prologue per
// boundary.  Measured on the ViT-B/16 step (round 2): 31.70 ms with PDL on every kernel vs 30.94 ms without — inside a CUDA graph the
// kernel-to-kernel latency is already ~1 us and the early CTAs only disturb the tail of the kernel in front — so it is OFF by default
// (VITB200_PDL=1 all kernels, 2 GEMM only, 3 all but the GEMM; griddepcontrol.* are no-ops under plain stream order).
#pragma once
#include <cuda_runtime.h>

namespace vb {

bool pdl_enabled();        // lib.cu
bool pdl_enabled_gemm();

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace vb
