// mma_probe.cu -- microbenchmarks of the tcgen05.mma issue path on one CTA per SM (bf16, M = 128 per CTA, K = 16).
//
// Written to find out why the weight-gradient kernels needed 2.3-4x the ideal tensor-pipe cycles per MMA
// (profiles/r2_wgrad_ncu.md).  What it established (profiles/r2_mma_probe.log, B200):
//   * probe2_kernel (operands resident in shared memory, nothing else in the loop): SS-mode MMAs run at the tensor-pipe floor
//     for N >= 128 (64.0 / 128.0 cycles at N = 128 / 256, cta_group::1 and ::2, one or several accumulators, A from shared
//     memory or from TMEM); N = 64 needs 54 cycles instead of 32 (49.5 with A in TMEM).  A dependent accumulator chain
//     costs nothing.
//   * probe_kernel (a producer / consumer pipeline like the real kernels): K-major and MN-major SWIZZLE_128B operands are
//     indistinguishable in every configuration -- the operand mode is not the problem.
//   * the same pipeline is bound by the ONE warp that issues: with a runtime `it % stages` and `it / stages` in the loop
//     (two integer divisions, ~100 dependent instructions) it cannot start more than one 4-MMA K block per ~520-580 cycles,
//     whatever N is and whether or not data is streamed.  Issuing from inside `if (lane == 0)` adds an ELECT / BRA.U.ANY
//     waterfall per UTCHMMA on top (147 vs 130 cycles per MMA here).
// The consequence for csrc/conv_tcgen05.cu: producer and MMA warps run converged with one elected lane, keep stage / phase /
// tile coordinates / descriptors as loop-carried values, and contain no division inside the K loop.
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/mma_probe tools/mma_probe.cu
// run:   tools/mma_probe        (prints one line per configuration: cycles per MMA, median over the SMs)
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 2000000000ll) { printf("probe: mbarrier timeout\n"); __trap(); }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (mn ? ((1u << 15) | (1u << 16)) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// One K block = 64 K elements: K-major A = 128 rows x 128 B (16 KB), B = N rows x 128 B; MN-major: 64-channel groups of
// 64 K rows x 128 B (8 KB each), A = 2 groups, B = N/64 groups.  Same bytes either way.  STAGES blocks are cycled.
// load_bytes > 0: warp 0 keeps a bulk-copy stream of `load_bytes` per K block running into the stage the MMA warp released
// (the real producer/consumer pipeline, global source = L2-resident buffer).
// UNI = false: producer / MMA loops run inside `if (lane == 0)` (every UTCHMMA / bulk copy becomes an ELECT + BRA.U.ANY
// waterfall loop); UNI = true: the whole warp runs the loop and one elected lane issues.
template <int N, bool MN, bool UNI>
__global__ void __launch_bounds__(64, 1) probe_kernel(int kblocks, int stages, uint32_t load_bytes, const uint8_t* __restrict__ src, long long* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = 128 * 128, B_BYTES = N * 128, STAGE = A_BYTES + B_BYTES;
  uint64_t* full = (uint64_t*)(smem + stages * STAGE);
  uint64_t* empty = full + 8;
  uint64_t* done = empty + 8;
  uint32_t* slot = (uint32_t*)(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < stages * STAGE / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 1) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const bool run = UNI ? true : (lane == 0);
  if (warp == 0 && run && load_bytes) {
    const uint8_t* my = src + (size_t)blockIdx.x * (size_t)STAGE * 4;   // 4-stage window per CTA: stays in L2
    for (int it = 0; it < kblocks; ++it) {
      const int s = it % stages;
      mbar_wait(&empty[s], ((it / stages) & 1) ^ 1u);
      if (!UNI || elect_one()) {
        mbar_expect_tx(&full[s], load_bytes);
        uint32_t left = load_bytes, off = 0;
        while (left) {
          const uint32_t n = left > 16384u ? 16384u : left;
          bulk_load(smem + s * STAGE + off, my + (size_t)(it & 3) * STAGE + off, n, &full[s]);
          left -= n; off += n;
        }
      }
      if (UNI) __syncwarp();
    }
  } else if (warp == 1 && run) {
    constexpr uint32_t idesc = idesc_bf16(128, N, MN);
    const uint32_t tm = UNI ? __shfl_sync(0xffffffffu, tmem, 0) : tmem;
    const long long t0 = clock64();
    for (int it = 0; it < kblocks; ++it) {
      const int s = it % stages;
      if (load_bytes) { mbar_wait(&full[s], (it / stages) & 1); tc_fence_after(); }
      const uint32_t sa = smem_u32(smem + s * STAGE);
      if (!UNI || elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint64_t ad, bd;
          if (MN) {
            ad = mnmajor_desc(sa, 64 * 128) + (uint64_t)(k * (2048 >> 4));
            bd = mnmajor_desc(sa + A_BYTES, 64 * 128) + (uint64_t)(k * (2048 >> 4));
          } else {
            ad = kmajor_desc(sa) + (uint64_t)(k * (32 >> 4));
            bd = kmajor_desc(sa + A_BYTES) + (uint64_t)(k * (32 >> 4));
          }
          umma_bf16(tm, ad, bd, idesc, (it | k) != 0 ? 1u : 0u);
        }
        if (load_bytes) umma_commit(&empty[s]);
        if (it == kblocks - 1) umma_commit(done);
      }
      if (UNI) __syncwarp();
    }
    mbar_wait(done, 0);
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}


// ---- second probe: resident operands only; independent accumulators (ACC, round-robin), A from TMEM (TS) and cta_group::2 ----
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// A operand in TMEM (128 lanes x 8 columns per K16 step of bf16)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// N: MMA N (<= 256; in cta_group::2 each CTA supplies N/2 rows of B); ACC accumulators used round-robin; TS: A from TMEM;
// CG: cta_group.  K-major operands, 64-element K block = 4 MMAs.
template <int N, int ACC, bool TS, int CG>
__global__ void __launch_bounds__(64, 1) probe2_kernel(int kblocks, long long* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = 128 * 128, B_BYTES = 256 * 128, STAGE = A_BYTES + B_BYTES, STAGES = 2;
  uint64_t* done = (uint64_t*)(smem + STAGES * STAGE);
  uint32_t* slot = (uint32_t*)(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < STAGES * STAGE / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(done, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 1) { if (CG == 2) tmem_alloc_2sm(slot, 512); else tmem_alloc(slot, 512); }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1 && lane == 0 && rank == 0) {
    constexpr uint32_t idesc = idesc_bf16(128 * CG, N, false);
    const long long t0 = clock64();
    uint32_t j = 0;
    for (int it = 0; it < kblocks; ++it) {
      const uint32_t sa = smem_u32(smem + (it & 1) * STAGE);
#pragma unroll
      for (int k = 0; k < 4; ++k, ++j) {
        const uint32_t acc = tmem + (j % ACC) * (uint32_t)N;
        const uint64_t bd = kmajor_desc(sa + A_BYTES) + (uint64_t)(k * (32 >> 4));
        if (TS) umma_bf16_ts(acc, tmem + 480u + (uint32_t)(k * 8), bd, idesc, j >= ACC ? 1u : 0u);
        else if (CG == 2) umma_bf16_2sm(acc, kmajor_desc(sa) + (uint64_t)(k * (32 >> 4)), bd, idesc, j >= ACC ? 1u : 0u);
        else umma_bf16(acc, kmajor_desc(sa) + (uint64_t)(k * (32 >> 4)), bd, idesc, j >= ACC ? 1u : 0u);
      }
    }
    if (CG == 2) umma_commit_2sm(done); else umma_commit(done);
    mbar_wait(done, 0);
    const long long t1 = clock64();
    out[blockIdx.x / CG] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 1) { tc_fence_after(); if (CG == 2) tmem_dealloc_2sm(tmem, 512); else tmem_dealloc(tmem, 512); }
}

template <int N, int ACC, bool TS, int CG>
static void run2(const char* name, long long* dout, int sms) {
  static_assert(ACC * N + (TS ? 32 : 0) <= 512, "TMEM columns");
  const int kblocks = 4096;
  const size_t smem = (size_t)2 * (128 * 128 + 256 * 128) + 256 + 1024;
  cudaFuncSetAttribute(probe2_kernel<N, ACC, TS, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int ctas = sms / CG * CG, n = ctas / CG;
  for (int rep = 0; rep < 2; ++rep) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, probe2_kernel<N, ACC, TS, CG>, kblocks, dout);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-44s ERROR %s\n", name, cudaGetErrorString(e)); return; }
  std::vector<long long> h(n);
  cudaMemcpy(h.data(), dout, n * sizeof(long long), cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end());
  const double per = (double)h[n / 2] / (kblocks * 4.0), ideal = N / 2.0;   // per SM: 128 rows x N x K16 at 8192 FLOP/clk
  printf("%-30s cta_group::%d M=%3d N=%3d acc=%d  cycles/MMA median %.1f (min %.1f max %.1f)  ideal %.0f  -> %.2f of peak\n", name, CG, 128 * CG, N, ACC, per,
         (double)h[0] / (kblocks * 4.0), (double)h[n - 1] / (kblocks * 4.0), ideal, ideal / per);
}

template <int N, bool MN, bool UNI = false>
static void run(const char* name, int stages, uint32_t load_bytes, const uint8_t* src, long long* dout, int sms) {
  const int kblocks = 4096;
  const size_t smem = (size_t)stages * (128 * 128 + N * 128) + 256 + 1024;
  cudaFuncSetAttribute(probe_kernel<N, MN, UNI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) probe_kernel<N, MN, UNI><<<sms, 64, smem>>>(kblocks, stages, load_bytes, src, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-44s ERROR %s\n", name, cudaGetErrorString(e)); return; }
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), dout, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end());
  const double per = (double)h[sms / 2] / (kblocks * 4.0), ideal = N / 2.0;   // M128 x N x K16 bf16: N/2 cycles at 8192 FLOP/clk/SM
  printf("%-32s %s N=%3d stages=%d load=%6u B/blk  cycles/MMA median %.1f (min %.1f max %.1f)  ideal %.0f  -> %.2f of peak\n", name, UNI ? "converged-warp issue" : "lane-0 issue         ", N, stages, load_bytes,
         per, (double)h[0] / (kblocks * 4.0), (double)h[sms - 1] / (kblocks * 4.0), ideal, ideal / per);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint8_t* src;
  long long* dout;
  const size_t src_bytes = (size_t)sms * 4 * (128 * 128 + 256 * 128);
  cudaMalloc(&src, src_bytes);
  cudaMemset(src, 0, src_bytes);
  cudaMalloc(&dout, sms * sizeof(long long));
  // resident operands: the pure MMA + shared-memory read rate
  run<128, false>("K-major  resident", 3, 0, src, dout, sms);
  run<128, true>("MN-major resident", 3, 0, src, dout, sms);
  run<256, false>("K-major  resident", 3, 0, src, dout, sms);
  run<256, true>("MN-major resident", 3, 0, src, dout, sms);
  run<64, false>("K-major  resident", 3, 0, src, dout, sms);
  run<64, true>("MN-major resident", 3, 0, src, dout, sms);
  // streamed operands: every K block (A + B) arrives through a bulk copy, like the real kernels
  run<128, false>("K-major  A+B streamed", 4, 128 * 128 + 128 * 128, src, dout, sms);
  run<128, true>("MN-major A+B streamed", 4, 128 * 128 + 128 * 128, src, dout, sms);
  run<256, false>("K-major  A+B streamed", 4, 128 * 128 + 256 * 128, src, dout, sms);
  run<256, true>("MN-major A+B streamed", 4, 128 * 128 + 256 * 128, src, dout, sms);
  // half the stream (what a 2-SM pair or a multicast cluster sees per CTA)
  run<128, true>("MN-major half streamed", 4, 128 * 128, src, dout, sms);
  run<256, true>("MN-major half streamed", 4, 128 * 128 + 128 * 128, src, dout, sms);
  run<256, false>("K-major  half streamed", 4, 128 * 128 + 128 * 128, src, dout, sms);
  // the same with converged-warp issue: what the operand feed itself sustains (all SMs pulling from L2)
  run<128, false, true>("K-major  resident", 3, 0, src, dout, sms);
  run<128, true, true>("MN-major resident", 3, 0, src, dout, sms);
  run<256, false, true>("K-major  resident", 3, 0, src, dout, sms);
  run<256, true, true>("MN-major resident", 3, 0, src, dout, sms);
  run<64, true, true>("MN-major resident", 3, 0, src, dout, sms);
  run<128, false, true>("K-major  A+B streamed", 4, 128 * 128 + 128 * 128, src, dout, sms);
  run<128, true, true>("MN-major A+B streamed", 4, 128 * 128 + 128 * 128, src, dout, sms);
  run<256, false, true>("K-major  A+B streamed", 4, 128 * 128 + 256 * 128, src, dout, sms);
  run<256, true, true>("MN-major A+B streamed", 4, 128 * 128 + 256 * 128, src, dout, sms);
  run<256, false, true>("K-major  A+B/2 streamed", 4, 128 * 128 + 128 * 128, src, dout, sms);
  run<256, true, true>("MN-major A+B/2 streamed", 4, 128 * 128 + 128 * 128, src, dout, sms);
  run<256, true, true>("MN-major A streamed", 4, 128 * 128, src, dout, sms);
  run<128, true, true>("MN-major A streamed", 4, 128 * 128, src, dout, sms);
  // is the ~132-cycle floor a dependent-accumulator latency or a per-instruction cost?
  run2<64, 1, false, 1>("SS", dout, sms);
  run2<64, 2, false, 1>("SS", dout, sms);
  run2<64, 4, false, 1>("SS", dout, sms);
  run2<128, 1, false, 1>("SS", dout, sms);
  run2<128, 2, false, 1>("SS", dout, sms);
  run2<128, 4, false, 1>("SS", dout, sms);
  run2<256, 1, false, 1>("SS", dout, sms);
  run2<256, 2, false, 1>("SS", dout, sms);
  // A operand from TMEM
  run2<64, 1, true, 1>("TS (A in TMEM)", dout, sms);
  run2<64, 4, true, 1>("TS (A in TMEM)", dout, sms);
  run2<128, 1, true, 1>("TS (A in TMEM)", dout, sms);
  run2<128, 2, true, 1>("TS (A in TMEM)", dout, sms);
  run2<256, 1, true, 1>("TS (A in TMEM)", dout, sms);
  // cta_group::2
  run2<64, 1, false, 2>("SS 2-SM", dout, sms);
  run2<128, 1, false, 2>("SS 2-SM", dout, sms);
  run2<128, 2, false, 2>("SS 2-SM", dout, sms);
  run2<256, 1, false, 2>("SS 2-SM", dout, sms);
  run2<256, 2, false, 2>("SS 2-SM", dout, sms);
  cudaFree(src);
  cudaFree(dout);
  return 0;
}
