// mma_probe.cu -- microbenchmark: issue rate of tcgen05.mma (cta_group::1, bf16, M=128) with K-major vs MN-major SWIZZLE_128B
// shared-memory operands, with and without a concurrent bulk-copy stream into the same shared memory.
//
// Question it answers (profiles/r2_wgrad_ncu.md): the weight-gradient kernels need 2.3-4x the ideal tensor-pipe cycles per MMA.
// Is that the MN-major operand mode itself, or the operand feed (shared-memory write bandwidth of the TMA stream)?
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
template <int N, bool MN>
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
  if (warp == 0 && lane == 0 && load_bytes) {
    const uint8_t* my = src + (size_t)blockIdx.x * (size_t)STAGE * 4;   // 4-stage window per CTA: stays in L2
    for (int it = 0; it < kblocks; ++it) {
      const int s = it % stages;
      mbar_wait(&empty[s], ((it / stages) & 1) ^ 1u);
      mbar_expect_tx(&full[s], load_bytes);
      uint32_t left = load_bytes, off = 0;
      while (left) {
        const uint32_t n = left > 16384u ? 16384u : left;
        bulk_load(smem + s * STAGE + off, my + (size_t)(it & 3) * STAGE + off, n, &full[s]);
        left -= n; off += n;
      }
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = idesc_bf16(128, N, MN);
    const long long t0 = clock64();
    for (int it = 0; it < kblocks; ++it) {
      const int s = it % stages;
      if (load_bytes) { mbar_wait(&full[s], (it / stages) & 1); tc_fence_after(); }
      const uint32_t sa = smem_u32(smem + s * STAGE);
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
        umma_bf16(tmem, ad, bd, idesc, (it | k) != 0 ? 1u : 0u);
      }
      if (load_bytes) umma_commit(&empty[s]);
    }
    umma_commit(done);
    mbar_wait(done, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, bool MN>
static void run(const char* name, int stages, uint32_t load_bytes, const uint8_t* src, long long* dout, int sms) {
  const int kblocks = 4096;
  const size_t smem = (size_t)stages * (128 * 128 + N * 128) + 256 + 1024;
  cudaFuncSetAttribute(probe_kernel<N, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) probe_kernel<N, MN><<<sms, 64, smem>>>(kblocks, stages, load_bytes, src, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-44s ERROR %s\n", name, cudaGetErrorString(e)); return; }
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), dout, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end());
  const double per = (double)h[sms / 2] / (kblocks * 4.0), ideal = N / 2.0;   // M128 x N x K16 bf16: N/2 cycles at 8192 FLOP/clk/SM
  printf("%-44s N=%3d stages=%d load=%6u B/blk  cycles/MMA median %.1f (min %.1f max %.1f)  ideal %.0f  -> %.2f of peak\n", name, N, stages, load_bytes,
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
  cudaFree(src);
  cudaFree(dout);
  return 0;
}
