// pdl.cuh — programmatic dependent launch (sm_90+): the kernels of one control step of the closed-loop rollout form a chain on
// one stream (actor layers -> physics -> post-step -> next step's actor ...), each a few microseconds long, so the gaps
// between them are a visible share of the step.  Launched with cudaLaunchAttributeProgrammaticStreamSerialization a kernel may
// be scheduled as soon as every CTA of its predecessor has started (pdl_trigger) and runs its prologue while the predecessor
// drains; pdl_wait() blocks until the predecessor grid has COMPLETED and its writes are visible, so placing it in front of the
// first access to global memory keeps ordinary stream semantics.  Both are no-ops for kernels launched the ordinary way.
#pragma once
#include <cuda_runtime.h>

namespace pdl {
// trigger(): at kernel entry — measured SLOWER in the 3-way pipelined rollout (0.152 vs 0.136 s: dependents that only wait occupy
// the shared memory / TMEM other pipelines' kernels could use) and compiled out; trigger_late(): where a CTA's main work is done
__device__ __forceinline__ void trigger() {
#if defined(SPI_PDL_EARLY_TRIGGER)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void trigger_late() {
#if !defined(SPI_PDL_NO_LATE_TRIGGER)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// host: launch `kernel` on `st`, programmatically serialised behind its predecessor on that stream when `enabled`
template <class... KArgs, class... Args>
inline cudaError_t launch(bool enabled, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
}  // namespace pdl
