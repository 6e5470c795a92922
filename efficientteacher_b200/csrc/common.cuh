// common.cuh -- shared helpers for libetb200 (sm_100a).  Compiled with --fmad=false: every fused
// multiply-add in this library is an explicit fmaf()/__fmaf_rn so the bit-exact kernels (EMA, assigner,
// NMS) reproduce the two-rounding arithmetic of the reference's CPU fp32 path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../include/etb200.h"

void etb_set_error(const char* fmt, ...);
void etb_count_launch();   // every ETB_CHECK_LAUNCH() follows exactly one kernel launch of this library

#define ETB_CHECK_ARG(cond)                                                   \
  do {                                                                        \
    if (!(cond)) {                                                            \
      etb_set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, #cond); \
      return ETB_ERR_INVALID;                                                 \
    }                                                                         \
  } while (0)

#define ETB_CHECK_LAUNCH()                                                               \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      etb_set_error("%s:%d: CUDA launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return ETB_ERR_CUDA;                                                               \
    }                                                                                    \
    etb_count_launch();                                                                  \
  } while (0)

#define ETB_CHECK_CUDA(call)                                                             \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      etb_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return ETB_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------------
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts with
// ETB_PDL_PROLOGUE(): `griddepcontrol.wait` (returns when the preceding kernel of the stream has completed and its memory is
// visible -- so no kernel touches global data before its producer is done) followed by `griddepcontrol.launch_dependents`
// (the NEXT kernel's CTAs may be scheduled as soon as all CTAs of this one have passed this point or exited).  The next
// kernel's launch latency, block scheduling and prologue (smem carve-up, mbarrier init, TMEM allocation, tensor-map
// prefetch: the tcgen05 kernels place the wait after that prologue) thereby overlap this kernel's execution instead of
// following its tail.  Measured on the SSOD step: no gain once the step is replayed as a CUDA graph (the graph's kernel-to-kernel
// latency is already ~1 us), so the attribute is only set with ETB_PDL=1 (eager-launch experiments).  A kernel launched without the attribute (ETB_PDL=0, or a neighbour from another library) sees both instructions
// as no-ops / full stream order, so mixing is safe.  Captured CUDA graphs keep the programmatic edges.
#define ETB_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define ETB_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define ETB_PDL_PROLOGUE() \
  do {                     \
    ETB_PDL_WAIT();        \
    ETB_PDL_TRIGGER();     \
  } while (0)

static inline bool etb_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ETB_PDL");
    v = (e && e[0] == '1') ? 1 : 0;       // opt-in: measured neutral under CUDA-graph replay (33.0 vs 32.8 ms/step, profiles/r2_ablation.md)
  }
  return v != 0;
}

template <typename... KP, typename... A>
static inline void etb_launch(void (*kernel)(KP...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = etb_pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KP>(args)...);    // the caller checks cudaGetLastError() (ETB_CHECK_LAUNCH)
}

static inline int etb_num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32).
// Returns the exclusive prefix; *total receives the block sum.  smem: int[33].
__device__ __forceinline__ int block_excl_scan(int v, int* smem, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // protect smem reuse across calls
  if (lane == 31) smem[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = lane < nw ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    smem[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  *total = smem[32];
  return smem[wid] + inc - v;
}
