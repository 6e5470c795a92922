// conv_tcgen05.cu -- implicit-GEMM convolution on the 5th-gen tensor cores (K1/K5).
//
// Replaces cuDNN conv + BN(eval) + SiLU (+ residual add / Detect bias + permute copy) behind
//   Conv.forward        reference models/backbone/common.py:471-484
//   Bottleneck.forward  reference models/backbone/common.py:534-544   (residual add fused in the epilogue)
//   Detect.forward      reference models/head/yolov5_head.py:55,66     (1x1 conv + bias, scattered to [B,3,ny,nx,85])
//
// GEMM view per CTA:  D[128 pixels, BN couts] += A[128 pixels, 64 ch] * B[BN couts, 64 ch]^T  over taps x channel blocks.
//   * A (activations, NHWC bf16) comes straight from global memory through a 4-D TMA tile
//     {64 ch, TW, TH, 1 image}; the tap shift (kh,kw) is a coordinate offset and the zero padding is TMA's
//     out-of-bounds fill, so no im2col buffer exists.  Stride-2 convs use the tensor map's element strides.
//   * B (weights, [Cout][kh*kw*Cin] bf16, K-major) is a 2-D TMA tile {64, BN}.
//   * both land in 128B-swizzled shared memory = the canonical K-major UMMA layout; one elected thread issues
//     tcgen05.mma (M=128, N=BN, K=16) x4 per stage; the fp32 accumulator lives in TMEM (BN columns).
//   * warp roles: warp0 TMA producer, warp1 TMEM alloc + MMA issue, warps2-5 epilogue (tcgen05.ld -> scale/bias
//     (folded BN) -> SiLU -> +residual -> bf16 NHWC store, written at a channel offset of a wider buffer so
//     torch.cat is free).
//   * every mbarrier wait is bounded (trap after ~2 s) so a descriptor bug cannot hang the GPU.
#include "common.cuh"
#include <cuda.h>

#define CONV_BLOCK_M 128
#define CONV_BLOCK_K 64
#define CONV_THREADS 320     // warp0 TMA, warp1 MMA, warps2-9 epilogue
#define WGRAD_THREADS 192
#define EPI_TILE_BYTES 4096  // one warp's epilogue staging tile: 32 rows x 128 B
#ifndef ETB_WGRAD2_DEFAULT
#define ETB_WGRAD2_DEFAULT -1    // 2-SM weight-gradient kernels: -1 per-shape rule, 0 off, 1 NW = 128, 2 NW = 256 (env ETB_WGRAD2 overrides)
#endif

// ------------------------------------------------------------------------------------------------- PTX wrappers
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
// bounded wait: a wrong descriptor / byte count must fail loudly, never hang the box
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("etb conv: mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}
// One lane of a CONVERGED warp (elect.sync).  The producer / MMA warps run their loops with all 32 lanes (warp-uniform
// control flow, operands in uniform registers) and wrap only the TMA / tcgen05 instructions in `if (elect_one())`: issued from
// inside an `if (lane == 0)` region every UTCHMMA / UTMALDG / UTCBAR is emitted as an ELECT + R2UR.BROADCAST + BRA.U.ANY
// waterfall loop (~50 cycles per instruction on the single issuing thread, measured with tools/mma_probe.cu), which bounds
// every tile with N < 256 by instruction issue instead of by the tensor pipe.
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

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// multicast variant: the box lands at the same CTA-relative smem offset of every CTA in `mask` and signals each one's
// mbarrier at the same offset
__device__ __forceinline__ void tma_load_4d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (8 rows x 128 B = 1024 B)
//   [46,48) version=1 (sm_100) | [61,64) layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (bit 4), a/b format BF16 (bits 7, 10),
// both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------- kernel
struct ConvKArgs {
  int ntaps;              // filter taps visited (kh*kw forward; 1/2/4 per parity class of a stride-2 dgrad)
  signed char tap_dh[12], tap_dw[12];   // input coordinate = tile origin * stride + tap offset (zero padding = TMA OOB fill)
  int kblocks;            // K channels / 64 per tap
  int stride;
  int out_os, out_ph, out_pw;           // output pixel (oh,ow) is stored at (oh*out_os+out_ph, ow*out_os+out_pw) ...
  int out_H, out_W;                     // ... of an out_H x out_W plane (dgrad of stride-2 convs fills one parity lattice)
  int accumulate;         // out_mode 0: y += result (gradient accumulation for tensors with several consumers)
  int TW, TH;             // output tile (TW*TH <= 128 rows)
  int tiles_w, tiles_h;   // per image
  int nimg;               // images (1 in the flat pointwise tiling)
  int Ho, Wo, Cout;
  int y_cstride, y_coffset;
  int res_cstride, res_coffset;
  int act;                // 0 none, 1 SiLU, 2 ReLU
  int epi_staged;         // bf16 epilogues store through the shared-memory staging tile (128 contiguous bytes per pixel row)
  int out_mode;           // 0: bf16 NHWC ; 1: fp32 Detect layout [N, na, Ho, Wo, no] with c = a*no + o
  int det_no, det_hw;     // outputs per anchor, pixels per image (Detect layout)
  const float* scale;     // [Cout] or null (=1)
  const float* bias;      // [Cout] or null (=0)
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  float* y_f32;
};

template <int BN, int STAGES>
struct ConvSmem {
  static constexpr int A_BYTES = CONV_BLOCK_M * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int EPI_OFF = BAR_OFF + 256;             // 8 epilogue warps x one staging tile
  static constexpr int TOTAL = EPI_OFF + 8 * EPI_TILE_BYTES + 1024;   // + slack for the 1024 B alignment
  static constexpr int TMEM_COLS = 2 * BN;                  // double-buffered accumulator
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Epilogue store staging.  tcgen05.ld hands every lane ONE accumulator row, so a direct store instruction touches 32
// different output rows with 16 B each (32 partial-sector transactions).  A warp instead parks a 32-row x 128-byte tile
// (row = lane, 8 pieces of 16 B, XOR-swizzled: conflict-free both ways) in shared memory and reads it back so that store
// instruction i covers rows 4i..4i+3 with 128 contiguous bytes each (4 full lines).
__device__ __forceinline__ void stage_write(uint8_t* tile, int lane, int piece, const uint4& v) {
  *reinterpret_cast<uint4*>(tile + lane * 128 + ((piece ^ (lane & 7)) << 4)) = v;
}
__device__ __forceinline__ uint4 stage_read(const uint8_t* tile, int lane, int i) {
  const int p = 4 * i + (lane >> 3), c = lane & 7;
  return *reinterpret_cast<const uint4*>(tile + p * 128 + ((c ^ (p & 7)) << 4));
}

// Epilogue of one accumulator tile for one warp: columns [chalf*BN/2, (chalf+1)*BN/2) of the BN-wide tile starting at
// output channel n0, for the output pixel `pix` of this thread's TMEM lane.  EPI is a compile-time tail selector so every
// register array is statically indexed:
//   EPI 0: raw bf16 store (+= existing when a.accumulate)       -- dgrad, training forward
//   EPI 1: v*scale+bias (folded BN) -> SiLU/ReLU -> (+residual) -- teacher forward
//   EPI 2: +bias, fp32 scatter into the Detect layout           -- head
// One 32-column chunk; returns false when the chunk starts beyond Cout (warp-uniform).  pre: the 4 x 16 B of the existing
// output this chunk adds to (EPI 0 accumulate), loaded by the caller BEFORE it waited for the accumulator.
template <int BN, int EPI>
__device__ __forceinline__ bool conv_epilogue_chunk(const ConvKArgs& a, uint32_t lane_addr, int n0, int chalf, bool row_ok, size_t pix, int cc,
                                                    const uint4* pre) {
  constexpr int HALF = BN / 2;
{
  const int c0 = chalf * HALF + cc;
  if (n0 + c0 >= a.Cout) return false;  // warp-uniform
  uint32_t v[32];
  tmem_ld32(lane_addr + (uint32_t)c0, v);
  if (!row_ok) return true;
  const int gc0 = n0 + c0;
  const bool full = gc0 + 32 <= a.Cout;
  if (EPI == 2) {
    // Detect train layout: y[img][anchor][oh][ow][o], channel c = anchor*no + o  (yolov5_head.py:66)
    const size_t hw = (size_t)a.det_hw;
    const size_t img_r = pix / hw, pin = pix - img_r * hw;   // pix is the global pixel index in both tilings
    const int na = a.Cout / a.det_no;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int gc = gc0 + j;
      if (gc < a.Cout) {
        const int an = gc / a.det_no, o = gc - an * a.det_no;
        a.y_f32[((img_r * na + an) * hw + pin) * a.det_no + o] = __uint_as_float(v[j]) + (a.bias ? __ldg(a.bias + gc) : 0.0f);
      }
    }
    return true;
  }
  __nv_bfloat16* yp = a.y + pix * a.y_cstride + a.y_coffset + gc0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {         // 8 channels = one 16 B store
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[q * 8 + j]);
    if (EPI == 1) {
      float sc[8], bi[8];
      if (full) {
        const float4 s0 = a.scale ? __ldg(reinterpret_cast<const float4*>(a.scale + gc0) + 2 * q) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 s1 = a.scale ? __ldg(reinterpret_cast<const float4*>(a.scale + gc0) + 2 * q + 1) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 b0 = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + gc0) + 2 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 b1 = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + gc0) + 2 * q + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
        sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
        bi[0] = b0.x; bi[1] = b0.y; bi[2] = b0.z; bi[3] = b0.w; bi[4] = b1.x; bi[5] = b1.y; bi[6] = b1.z; bi[7] = b1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int gc = gc0 + q * 8 + j;
          sc[j] = (a.scale && gc < a.Cout) ? __ldg(a.scale + gc) : 1.0f;
          bi[j] = (a.bias && gc < a.Cout) ? __ldg(a.bias + gc) : 0.0f;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float x = fmaf(f[j], sc[j], bi[j]);
        if (a.act == 1) { const float h = 0.5f * x; float th; asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h)); x = fmaf(h, th, h); }   // SiLU with one MUFU op
        else if (a.act == 2) x = fmaxf(x, 0.0f);
        f[j] = x;
      }
      if (a.residual && full) {
        const uint4 rv = __ldg(reinterpret_cast<const uint4*>(a.residual + pix * a.res_cstride + a.res_coffset + gc0) + q);
        const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 rf = __bfloat1622float2(r2[j]);
          f[2 * j] += rf.x;
          f[2 * j + 1] += rf.y;
        }
      } else if (a.residual) {       // last, partial 32-channel chunk (Cout % 32 != 0: YOLOv5m / custom widths)
        const __nv_bfloat16* rp = a.residual + pix * a.res_cstride + a.res_coffset + gc0 + q * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (gc0 + q * 8 + j < a.Cout) f[j] += __bfloat162float(rp[j]);
      }
    }
    if (full) {
      if (EPI == 0 && a.accumulate) {
        const uint4 pv = pre[q];
        const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&pv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 pf = __bfloat1622float2(p2[j]);
          f[2 * j] += pf.x;
          f[2 * j + 1] += pf.y;
        }
      }
      uint4 ov;
      __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
      for (int j = 0; j < 4; ++j) o2[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
      reinterpret_cast<uint4*>(yp)[q] = ov;
    } else {
      // partial chunk: scalar stores; the fan-in accumulate (dx += dgrad) reads the existing value here (the preload only
      // covers whole 32-channel chunks)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (gc0 + q * 8 + j < a.Cout) {
          float o = f[j];
          if (EPI == 0 && a.accumulate) o += __bfloat162float(yp[q * 8 + j]);
          yp[q * 8 + j] = __float2bfloat16(o);
        }
    }
  }
}
  return true;
}

// EPI 0 accumulate (dx += dgrad: gradient fan-in): the existing output of this thread's pixel row (HALF channels) is
// fetched into registers by epilogue_preload() at the top of the tile, i.e. while the MMAs of the tile are still running,
// so its latency is off the epilogue's critical path (a dependent load per chunk cost +40 % on the pointwise dgrads).
template <int BN, int EPI>
struct EpiPre { uint4 v[EPI == 0 ? BN / 16 : 1]; };

template <int BN, int EPI>
__device__ __forceinline__ void epilogue_preload(const ConvKArgs& a, int n0, int chalf, bool row_ok, size_t pix, EpiPre<BN, EPI>& pre) {
  if (EPI == 0) {
    if (a.accumulate && row_ok) {
      constexpr int HALF = BN / 2;
      const int gc = n0 + chalf * HALF;
      const uint4* yp4 = reinterpret_cast<const uint4*>(a.y + pix * a.y_cstride + a.y_coffset + gc);
#pragma unroll
      for (int i = 0; i < BN / 16; ++i)
        if (gc + i * 8 + 8 <= a.Cout) pre.v[i] = yp4[i];
    }
  }
}

// One 16-byte output piece (8 consecutive channels starting at gc, all < Cout) of this lane's pixel row: EPI 1 applies the
// folded BN scale/bias, the activation and the shortcut; EPI 0 optionally adds the existing output (gradient fan-in).
template <int EPI>
__device__ __forceinline__ uint4 epi_piece(const ConvKArgs& a, const uint32_t* v8, int gc, bool row_ok, size_t pix, const uint4* preq) {
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v8[j]);
  if (EPI == 1) {
    const float4 s0 = a.scale ? __ldg(reinterpret_cast<const float4*>(a.scale + gc)) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 s1 = a.scale ? __ldg(reinterpret_cast<const float4*>(a.scale + gc) + 1) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 b0 = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + gc)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 b1 = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + gc) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float bi[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x = fmaf(f[j], sc[j], bi[j]);
      if (a.act == 1) { const float hh = 0.5f * x; float th; asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(hh)); x = fmaf(hh, th, hh); }
      else if (a.act == 2) x = fmaxf(x, 0.0f);
      f[j] = x;
    }
    if (a.residual && row_ok) {
      const uint4 rv = __ldg(reinterpret_cast<const uint4*>(a.residual + pix * a.res_cstride + a.res_coffset + gc));
      const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 rf = __bfloat1622float2(r2[j]);
        f[2 * j] += rf.x;
        f[2 * j + 1] += rf.y;
      }
    }
  }
  if (EPI == 0 && a.accumulate && row_ok) {
    const uint4 pv = *preq;
    const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&pv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 pf = __bfloat1622float2(p2[j]);
      f[2 * j] += pf.x;
      f[2 * j + 1] += pf.y;
    }
  }
  uint4 ov;
  __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
  for (int j = 0; j < 4; ++j) o2[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
  return ov;
}

// 64 output channels [cc, cc+64) of this warp's 32 pixel rows, stored through the staging tile: every lane converts its own
// row (two tcgen05.ld of 32 columns -> 8 pieces of 8 bf16), parks it, and store instruction i then writes pixels 4i..4i+3 with
// 128 contiguous bytes each.  Same arithmetic as conv_epilogue_chunk's whole-chunk branch; a tail (Cout % 64) takes that path.
template <int BN, int EPI>
__device__ __forceinline__ bool conv_epilogue_chunk64(const ConvKArgs& a, uint32_t lane_addr, int n0, int chalf, bool row_ok, size_t pix, int cc,
                                                      const uint4* pre, uint8_t* tile, int lane) {
  constexpr int HALF = BN / 2;
  const int c0 = chalf * HALF + cc;
  const int gc0 = n0 + c0;
  if (gc0 >= a.Cout) return false;                 // warp-uniform
  if (gc0 + 64 > a.Cout) {
    if (!conv_epilogue_chunk<BN, EPI>(a, lane_addr, n0, chalf, row_ok, pix, cc, pre)) return false;
    return conv_epilogue_chunk<BN, EPI>(a, lane_addr, n0, chalf, row_ok, pix, cc + 32, EPI == 0 ? pre + 4 : pre);
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t v[32];
    tmem_ld32(lane_addr + (uint32_t)(c0 + 32 * h), v);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      stage_write(tile, lane, 4 * h + q, epi_piece<EPI>(a, v + 8 * q, gc0 + 32 * h + 8 * q, row_ok, pix, EPI == 0 ? pre + 4 * h + q : pre));
  }
  __syncwarp();
  const unsigned long long row_off = (unsigned long long)pix * (unsigned long long)a.y_cstride;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + (lane >> 3);
    const unsigned long long off_r = __shfl_sync(0xffffffffu, row_off, r);
    const int ok_r = __shfl_sync(0xffffffffu, (int)row_ok, r);
    const uint4 o = stage_read(tile, lane, i);
    if (ok_r) *reinterpret_cast<uint4*>(a.y + off_r + a.y_coffset + gc0 + (lane & 7) * 8) = o;
  }
  __syncwarp();
  return true;
}

// The 64-wide tile (HALF = 32 channels = 64 B per row and warp): 32 rows x 64 B staged, store instruction i writes rows
// 8i..8i+7 with 64 contiguous bytes each.  Swizzle: 16-byte piece q of row r sits at piece q ^ ((r >> 1) & 3).
template <int BN, int EPI>
__device__ __forceinline__ bool conv_epilogue_chunk32s(const ConvKArgs& a, uint32_t lane_addr, int n0, int chalf, bool row_ok, size_t pix,
                                                       const uint4* pre, uint8_t* tile, int lane) {
  constexpr int HALF = BN / 2;
  const int c0 = chalf * HALF;
  const int gc0 = n0 + c0;
  if (gc0 >= a.Cout) return false;                 // warp-uniform
  if (gc0 + 32 > a.Cout) return conv_epilogue_chunk<BN, EPI>(a, lane_addr, n0, chalf, row_ok, pix, 0, pre);
  uint32_t v[32];
  tmem_ld32(lane_addr + (uint32_t)c0, v);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(tile + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = epi_piece<EPI>(a, v + 8 * q, gc0 + 8 * q, row_ok, pix, EPI == 0 ? pre + q : pre);
  __syncwarp();
  const unsigned long long row_off = (unsigned long long)pix * (unsigned long long)a.y_cstride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + (lane >> 2), c = lane & 3;
    const unsigned long long off_r = __shfl_sync(0xffffffffu, row_off, r);
    const int ok_r = __shfl_sync(0xffffffffu, (int)row_ok, r);
    const uint4 o = *reinterpret_cast<const uint4*>(tile + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
    if (ok_r) *reinterpret_cast<uint4*>(a.y + off_r + a.y_coffset + gc0 + c * 8) = o;
  }
  __syncwarp();
  return true;
}

// Detect layout (EPI 2): y[img][anchor][pixel][o] fp32 with channel c = anchor * no + o (yolov5_head.py:66).  For one anchor the
// outputs of consecutive pixels are contiguous, so the warp parks its 32 rows x 32 channels (+ bias) and then stores ONE pixel
// row per instruction: 32 lanes = 32 consecutive channels = 128 contiguous bytes (two runs at an anchor boundary); the
// channel -> (anchor, o) split is computed once per lane and chunk instead of once per element.
template <int BN>
__device__ __forceinline__ bool conv_epilogue_chunk_det(const ConvKArgs& a, uint32_t lane_addr, int n0, int chalf, bool row_ok, size_t pix, int cc,
                                                        uint8_t* tile, int lane) {
  constexpr int HALF = BN / 2;
  const int c0 = chalf * HALF + cc;
  const int gc0 = n0 + c0;
  if (gc0 >= a.Cout) return false;                 // warp-uniform
  uint32_t v[32];
  tmem_ld32(lane_addr + (uint32_t)c0, v);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float f[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gc = gc0 + 4 * q + j;
      f[j] = __uint_as_float(v[4 * q + j]) + ((a.bias && gc < a.Cout) ? __ldg(a.bias + gc) : 0.0f);
    }
    stage_write(tile, lane, q, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
  }
  __syncwarp();
  const size_t hw = (size_t)a.det_hw;
  const int na = a.Cout / a.det_no;
  const size_t img_r = pix / hw, pin = pix - img_r * hw;          // pix is the global pixel index in both tilings
  const unsigned long long base = (unsigned long long)((img_r * na * hw + pin) * a.det_no);
  const int gc = gc0 + lane;
  const bool c_ok = gc < a.Cout;
  const int an = gc / a.det_no, o = gc - an * a.det_no;
  const unsigned long long coff = (unsigned long long)an * hw * a.det_no + o;
#pragma unroll 4
  for (int r = 0; r < 32; ++r) {
    const unsigned long long base_r = __shfl_sync(0xffffffffu, base, r);
    const int ok_r = __shfl_sync(0xffffffffu, (int)row_ok, r);
    const float val = *reinterpret_cast<const float*>(tile + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
    if (ok_r && c_ok) a.y_f32[base_r + coff] = val;
  }
  __syncwarp();
  return true;
}

template <int BN, int EPI>
__device__ __forceinline__ void conv_epilogue_cols(const ConvKArgs& a, uint32_t lane_addr, int n0, int chalf, bool row_ok, size_t pix,
                                                   const EpiPre<BN, EPI>& pre, uint8_t* tile, int lane) {
  constexpr int HALF = BN / 2;
  if (EPI == 2 && a.epi_staged) {                     // warp-uniform
#pragma unroll 1
    for (int cc = 0; cc < HALF; cc += 32)
      if (!conv_epilogue_chunk_det<BN>(a, lane_addr, n0, chalf, row_ok, pix, cc, tile, lane)) break;
    return;
  }
  if (EPI != 2 && HALF == 32 && a.epi_staged) {
    conv_epilogue_chunk32s<BN, EPI>(a, lane_addr, n0, chalf, row_ok, pix, pre.v, tile, lane);
    return;
  }
  if (EPI != 2 && HALF % 64 == 0 && a.epi_staged) {
    if (EPI == 0) {
#pragma unroll
      for (int cc = 0; cc < HALF; cc += 64)
        if (!conv_epilogue_chunk64<BN, EPI>(a, lane_addr, n0, chalf, row_ok, pix, cc, pre.v + cc / 8, tile, lane)) break;
    } else {
#pragma unroll 1
      for (int cc = 0; cc < HALF; cc += 64)
        if (!conv_epilogue_chunk64<BN, EPI>(a, lane_addr, n0, chalf, row_ok, pix, cc, pre.v, tile, lane)) break;
    }
    return;
  }
  if (EPI == 0) {
#pragma unroll
    for (int cc = 0; cc < HALF; cc += 32)
      if (!conv_epilogue_chunk<BN, EPI>(a, lane_addr, n0, chalf, row_ok, pix, cc, pre.v + cc / 8)) break;
  } else {
#pragma unroll 1
    for (int cc = 0; cc < HALF; cc += 32)
      if (!conv_epilogue_chunk<BN, EPI>(a, lane_addr, n0, chalf, row_ok, pix, cc, pre.v)) break;
  }
}

// Persistent: gridDim.x CTAs walk the tile list (tile = blockIdx.x + i*gridDim.x; N tile fastest so the CTAs running
// concurrently share A tiles in L2).  The smem ring and the two TMEM accumulators run across tile boundaries, so the
// epilogue of tile i (tcgen05.ld -> BN/SiLU -> stores) overlaps the TMA/MMA main loop of tile i+1.
template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_fwd_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const ConvKArgs a) {
  using L = ConvSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + L::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;     // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (a.Cout + BN - 1) / BN;
  const int m_tiles = a.tiles_w * a.tiles_h * a.nimg;
  const int total_tiles = n_tiles * m_tiles;
  const int kiters = a.ntaps * a.kblocks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int q = 0; q < 2; ++q) { mbar_init(&tmem_full[q], 1); mbar_init(&tmem_empty[q], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, L::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ETB_PDL_PROLOGUE();      // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail

  if (warp == 0) {
    // ===== TMA producer: the whole warp runs the loop (uniform), one elected lane issues =====
    const uint32_t a_bytes = (uint32_t)(a.TW * a.TH * 128);
    int s = 0;                     // stage / phase advance incrementally: no divisions inside the K loop
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n0 = (tile % n_tiles) * BN;
      int t = tile / n_tiles;
      const int tw_i = t % a.tiles_w; t /= a.tiles_w;
      const int th_i = t % a.tiles_h; t /= a.tiles_h;
      const int img = t;
      const int w0 = tw_i * a.TW * a.stride, h0 = th_i * a.TH * a.stride;
      int bk = 0;                  // K coordinate of the weight tile
      for (int tp = 0; tp < a.ntaps; ++tp) {
        const int wc = w0 + a.tap_dw[tp], hc = h0 + a.tap_dh[tp];
        for (int kb = 0; kb < a.kblocks; ++kb, bk += CONV_BLOCK_K) {
          mbar_wait(&empty[s], ph ^ 1u);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if (elect_one()) {
            mbar_expect_tx(&full[s], a_bytes + (uint32_t)L::B_BYTES);
            tma_load_4d(&mapA, &full[s], sa, kb * CONV_BLOCK_K, wc, hc, img);
            tma_load_2d(&mapB, &full[s], sb, bk, n0);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-uniform loop, one elected lane issues the MMAs and their commit =====
    constexpr uint32_t idesc = make_idesc_bf16(CONV_BLOCK_M, BN);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    // stage index, phase and the two operand descriptors advance incrementally (descriptor address field = bytes >> 4;
    // the stage ring stays far below the field's 2^14 range, so a plain 64-bit add never carries out of it)
    const uint64_t adesc0 = make_kmajor_sw128_desc(smem_u32(smem));
    const uint64_t bdesc0 = make_kmajor_sw128_desc(smem_u32(smem) + L::A_BYTES);
    uint64_t adesc = adesc0, bdesc = bdesc0;
    int s = 0, lt = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int acc = lt & 1;
      mbar_wait(&tmem_empty[acc], (uint32_t)(((lt >> 1) & 1) ^ 1));   // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_u + (uint32_t)(acc * BN);
      for (int ki = 0; ki < kiters; ++ki) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < CONV_BLOCK_K / 16; ++k)  // +32 B per UMMA_K step inside the 128 B swizzle atom
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (ki | k) != 0 ? 1u : 0u);
          umma_commit(&empty[s]);  // frees the smem stage once these MMAs have read it
          if (ki == kiters - 1) umma_commit(&tmem_full[acc]);  // accumulator complete
        }
        __syncwarp();
        adesc += (uint64_t)(L::STAGE_BYTES >> 4); bdesc += (uint64_t)(L::STAGE_BYTES >> 4);
        if (++s == STAGES) { s = 0; ph ^= 1u; adesc = adesc0; bdesc = bdesc0; }
      }
    }
  } else {
    // ===== epilogue: 8 warps.  A warp may only touch TMEM lanes 32*(warp%4)..+31, so warps w and w+4 share a lane
    // quadrant (= 32 output pixels) and split the BN columns in halves.  EPI selects the fused tail at compile time so
    // every register array is statically indexed (no local memory):
    //   EPI 0: raw bf16 store (+= existing when a.accumulate)       -- dgrad, training forward
    //   EPI 1: v*scale+bias (folded BN) -> SiLU/ReLU -> (+residual) -- teacher forward
    //   EPI 2: +bias, fp32 scatter into the Detect layout           -- head
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int chalf = ew >> 2;
    const int row = 32 * quad + lane;
    const int th = row / a.TW, tw = row - th * a.TW;
    int lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int acc = lt & 1;
      const int n0 = (tile % n_tiles) * BN;
      int t = tile / n_tiles;
      const int tw_i = t % a.tiles_w; t /= a.tiles_w;
      const int th_i = t % a.tiles_h; t /= a.tiles_h;
      const int img = t;
      const int oh = th_i * a.TH + th, ow = tw_i * a.TW + tw;
      const bool row_ok = (row < a.TW * a.TH) && (oh < a.Ho) && (ow < a.Wo);
      const size_t pix = ((size_t)img * a.out_H + (size_t)(oh * a.out_os + a.out_ph)) * a.out_W + (size_t)(ow * a.out_os + a.out_pw);
      EpiPre<BN, EPI> pre;
      epilogue_preload<BN, EPI>(a, n0, chalf, row_ok, pix, pre);
      mbar_wait(&tmem_full[acc], (uint32_t)((lt >> 1) & 1));
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(32 * quad) << 16);
      conv_epilogue_cols<BN, EPI>(a, lane_addr, n0, chalf, row_ok, pix, pre, smem + L::EPI_OFF + ew * EPI_TILE_BYTES, lane);
      // all tcgen05.ld of this warp have completed (wait::ld inside tmem_ld32): hand the accumulator back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, L::TMEM_COLS);
  }
}

// =====================================================================================================================
// 2-SM variant (cta_group::2): a cluster of two CTAs owns a 256-pixel x 256-cout tile.  Each SM stages ONLY its own
// 128-pixel A tile and HALF of the weight tile (128 of the 256 couts) -> 32 KB instead of 48 KB per K block and SM, which is
// what bounds the 1-SM kernel (per-SM TMA->smem ingest).  The leader CTA (rank 0) issues
// tcgen05.mma.cta_group::2 (M=256, N=256, K=16); each CTA's TMEM receives its own 128 rows x 256 columns and runs its own
// epilogue.  Barriers: both producers complete_tx on the LEADER's full barrier (peer bit cleared in the address),
// tcgen05.commit multicasts to both CTAs' empty / tmem_full barriers, both CTAs' epilogue warps arrive on the leader's
// tmem_empty barrier.
// =====================================================================================================================
#define ETB_PEER_BIT_MASK 0xFEFFFFFFu

__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & ETB_PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & ETB_PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
// arrive (no tx) on the LEADER CTA's copy of a barrier, from either CTA of the pair
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & ETB_PEER_BIT_MASK) : "memory");
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
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {   // arrive on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

template <int BN, int STAGES>
struct Conv2Smem {
  static constexpr int A_BYTES = CONV_BLOCK_M * 128;      // this CTA's 128 pixels
  static constexpr int B_BYTES = (BN / 2) * 128;          // this CTA's half (BN/2 couts) of the BN-wide weight tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int EPI_OFF = BAR_OFF + 256;
  static constexpr int TOTAL = EPI_OFF + 8 * EPI_TILE_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BN;                // two BN-column accumulators per CTA
};

template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_fwd2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const ConvKArgs a) {
  using L = Conv2Smem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + L::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;     // [2]  (the leader's copies are the live ones)
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();           // 0 = leader
  const int n_tiles = (a.Cout + BN - 1) / BN;
  const int m_tiles = a.tiles_w * a.tiles_h * a.nimg;
  const int m_pairs = (m_tiles + 1) / 2;
  const int total_tiles = n_tiles * m_pairs;         // pair tiles: 256 pixels x 256 couts
  const int kiters = a.ntaps * a.kblocks;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 2); mbar_init(&empty[s], 1); }
    for (int q = 0; q < 2; ++q) { mbar_init(&tmem_full[q], 1); mbar_init(&tmem_empty[q], 16); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, L::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // both CTAs' barriers + TMEM exist before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ETB_PDL_PROLOGUE();      // prologue above overlapped the previous kernel's tail; no global data touched before this

  if (warp == 0) {
    // ===== TMA producer (both CTAs: own A tile, own half of B; bytes are accounted on the leader's full barrier) =====
    // (warp-uniform loop, one elected lane issues: see elect_one)
    const uint32_t a_bytes = (uint32_t)(a.TW * a.TH * 128);
    int s = 0;
    uint32_t ph = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
      const int n0 = (tile % n_tiles) * BN + (BN / 2) * (int)rank;
      int t = (tile / n_tiles) * 2 + (int)rank;    // this CTA's m tile (may be == m_tiles for the odd tail: OOB -> zeros)
      const int tw_i = t % a.tiles_w; t /= a.tiles_w;
      const int th_i = t % a.tiles_h; t /= a.tiles_h;
      const int img = t;
      const int w0 = tw_i * a.TW * a.stride, h0 = th_i * a.TH * a.stride;
      int bk = 0;
      for (int tp = 0; tp < a.ntaps; ++tp) {
        const int wc = w0 + a.tap_dw[tp], hc = h0 + a.tap_dh[tp];
        for (int kb = 0; kb < a.kblocks; ++kb, bk += CONV_BLOCK_K) {
          mbar_wait(&empty[s], ph ^ 1u);           // my own smem stage is free (commit is multicast to both CTAs)
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(&full[s], 2u * (a_bytes + (uint32_t)L::B_BYTES));
            else mbar_arrive_leader(&full[s]);
            tma_load_4d_2sm(&mapA, &full[s], sa, kb * CONV_BLOCK_K, wc, hc, img);
            tma_load_2d_2sm(&mapB, &full[s], sb, bk, n0);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: leader CTA only, for the pair (warp-uniform loop, one elected lane issues) =====
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint64_t adesc0 = make_kmajor_sw128_desc(smem_u32(smem));
      const uint64_t bdesc0 = make_kmajor_sw128_desc(smem_u32(smem) + L::A_BYTES);
      uint64_t adesc = adesc0, bdesc = bdesc0;
      int s = 0, lt = 0;
      uint32_t ph = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++lt) {
        const int acc = lt & 1;
        mbar_wait(&tmem_empty[acc], (uint32_t)(((lt >> 1) & 1) ^ 1));   // both CTAs' epilogues drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_u + (uint32_t)(acc * BN);
        for (int ki = 0; ki < kiters; ++ki) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < CONV_BLOCK_K / 16; ++k)
              umma_bf16_2sm(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (ki | k) != 0 ? 1u : 0u);
            umma_commit_2sm(&empty[s]);
            if (ki == kiters - 1) umma_commit_2sm(&tmem_full[acc]);
          }
          __syncwarp();
          adesc += (uint64_t)(L::STAGE_BYTES >> 4); bdesc += (uint64_t)(L::STAGE_BYTES >> 4);
          if (++s == STAGES) { s = 0; ph ^= 1u; adesc = adesc0; bdesc = bdesc0; }
        }
      }
    }
  } else {
    // ===== epilogue (both CTAs, own 128 rows) =====
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int chalf = ew >> 2;
    const int row = 32 * quad + lane;
    const int th = row / a.TW, tw = row - th * a.TW;
    int lt = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++lt) {
      const int acc = lt & 1;
      const int n0 = (tile % n_tiles) * BN;
      const int mt = (tile / n_tiles) * 2 + (int)rank;
      int t = mt;
      const int tw_i = t % a.tiles_w; t /= a.tiles_w;
      const int th_i = t % a.tiles_h; t /= a.tiles_h;
      const int img = t;
      const int oh = th_i * a.TH + th, ow = tw_i * a.TW + tw;
      const bool row_ok = (mt < m_tiles) && (row < a.TW * a.TH) && (oh < a.Ho) && (ow < a.Wo);
      const size_t pix = ((size_t)img * a.out_H + (size_t)(oh * a.out_os + a.out_ph)) * a.out_W + (size_t)(ow * a.out_os + a.out_pw);
      EpiPre<BN, EPI> pre;
      epilogue_preload<BN, EPI>(a, n0, chalf, row_ok, pix, pre);
      mbar_wait(&tmem_full[acc], (uint32_t)((lt >> 1) & 1));
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(32 * quad) << 16);
      conv_epilogue_cols<BN, EPI>(a, lane_addr, n0, chalf, row_ok, pix, pre, smem + L::EPI_OFF + ew * EPI_TILE_BYTES, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // the peer may still signal / read until it is done too
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, L::TMEM_COLS);
  }
}

template <int BN, int STAGES, int EPI>
static int launch_conv2_e(const CUtensorMap& mA, const CUtensorMap& mB, const ConvKArgs& ka, dim3 grid, cudaStream_t st) {
  using L = Conv2Smem<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    ETB_CHECK_CUDA(cudaFuncSetAttribute(conv_fwd2_kernel<BN, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(CONV_THREADS); cfg.dynamicSmemBytes = L::TOTAL; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = etb_pdl_enabled() ? 2 : 1;
  ETB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_fwd2_kernel<BN, STAGES, EPI>, mA, mB, ka));
  etb_count_launch();
  return ETB_OK;
}
template <int BN, int STAGES>
static int launch_conv2(const CUtensorMap& mA, const CUtensorMap& mB, const ConvKArgs& ka, dim3 grid, cudaStream_t st) {
  if (ka.out_mode == 1) return launch_conv2_e<BN, STAGES, 2>(mA, mB, ka, grid, st);
  if (ka.scale || ka.bias || ka.act || ka.residual) return launch_conv2_e<BN, STAGES, 1>(mA, mB, ka, grid, st);
  return launch_conv2_e<BN, STAGES, 0>(mA, mB, ka, grid, st);
}

// ------------------------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiled)p;
  }
  return fn;
}

static void pick_tile(int Wo, int Ho, int* TW, int* TH) {
  // maximise useful rows of the 128-row MMA tile: TW*TH <= 128, TW <= 128 (stride-2 boxes need 2*TW <= 256)
  double best = -1.0;
  for (int tw = 1; tw <= 128; ++tw) {
    if (tw > Wo && tw != 1) break;
    int th = 128 / tw;
    if (th > Ho) th = Ho;
    if (th < 1) continue;
    const long tiles = (long)((Wo + tw - 1) / tw) * ((Ho + th - 1) / th);
    const double eff = (double)Wo * Ho / ((double)tiles * 128.0);
    if (eff > best + 1e-9) { best = eff; *TW = tw; *TH = th; }
  }
}

template <int BN, int STAGES, int EPI>
static int launch_conv_e(const CUtensorMap& mA, const CUtensorMap& mB, const ConvKArgs& ka, dim3 grid, cudaStream_t st) {
  using L = ConvSmem<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    ETB_CHECK_CUDA(cudaFuncSetAttribute(conv_fwd_kernel<BN, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  etb_launch(conv_fwd_kernel<BN, STAGES, EPI>, dim3(grid), dim3(CONV_THREADS), L::TOTAL, st, mA, mB, ka);
  ETB_CHECK_LAUNCH();
  return ETB_OK;
}
template <int BN, int STAGES>
static int launch_conv(const CUtensorMap& mA, const CUtensorMap& mB, const ConvKArgs& ka, dim3 grid, cudaStream_t st) {
  if (ka.out_mode == 1) return launch_conv_e<BN, STAGES, 2>(mA, mB, ka, grid, st);
  if (ka.scale || ka.bias || ka.act || ka.residual) return launch_conv_e<BN, STAGES, 1>(mA, mB, ka, grid, st);
  return launch_conv_e<BN, STAGES, 0>(mA, mB, ka, grid, st);
}

// One implicit-GEMM launch: D[pixels, rows_B] = sum_taps A(shifted) * B^T.  `ka` carries the epilogue.
struct GemmGeom {
  const void* a_ptr;          // NHWC bf16 tensor the A tiles come from
  int aN, aH, aW, aC, a_cstride;
  const void* b_ptr;          // [b_rows][ntaps*aC] bf16, K-major
  int b_rows;
  int tile_H, tile_W;         // per-image extent of the output-tile index space
  bool flat;                  // pointwise stride-1: all N*H*W pixels as one row-major [pixels, C] matrix
};

static int launch_gemm(const GemmGeom& g, ConvKArgs ka, cudaStream_t st) {
  PFN_tmapEncodeTiled enc = get_encode();
  if (!enc) {
    etb_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return ETB_ERR_CUDA;
  }
  // aC need not be a multiple of the 64-channel K block: the A box is clipped by TMA (zero fill beyond aC) and the B operand is
  // packed with every tap padded to aCp = ceil64(aC) zero columns (etb_pack_weight Cin_pad / dgrad out_ld)
  ETB_CHECK_ARG(g.aC % 8 == 0 && g.aC > 0 && g.a_cstride >= g.aC && g.a_cstride % 8 == 0);
  const int aCp = (g.aC + CONV_BLOCK_K - 1) / CONV_BLOCK_K * CONV_BLOCK_K;
  ETB_CHECK_ARG((((uintptr_t)g.a_ptr) & 15) == 0 && (((uintptr_t)g.b_ptr) & 15) == 0);
  ETB_CHECK_ARG(ka.ntaps >= 1 && ka.ntaps <= 12);
  {
    // epilogue store staging (conv_epilogue_chunk64): on unless ETB_EPI_STAGE=0; needs 16-byte aligned 64-channel runs
    static int staged = -1;
    if (staged < 0) { const char* e = getenv("ETB_EPI_STAGE"); staged = (e && e[0] == '0') ? 0 : 1; }
    ka.epi_staged = staged && ka.y_cstride % 8 == 0 && ka.y_coffset % 8 == 0 && (((uintptr_t)ka.y) & 15) == 0;
  }
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t box[4], estr[4];
  int nimg;
  if (g.flat) {
    const long npix = (long)g.aN * g.aH * g.aW;
    ETB_CHECK_ARG(npix < (1l << 31));
    ka.TW = 128; ka.TH = 1;
    ka.Ho = 1; ka.Wo = (int)npix;
    ka.tiles_w = (int)((npix + 127) / 128); ka.tiles_h = 1;
    ka.out_os = 1; ka.out_ph = ka.out_pw = 0; ka.out_H = 1; ka.out_W = (int)npix;
    nimg = 1;
    gdim[0] = g.aC; gdim[1] = (cuuint64_t)npix; gdim[2] = 1; gdim[3] = 1;
    gstr[0] = (cuuint64_t)g.a_cstride * 2; gstr[1] = gstr[0] * (cuuint64_t)npix; gstr[2] = gstr[1];
    box[0] = 64; box[1] = 128; box[2] = 1; box[3] = 1;
    estr[0] = estr[1] = estr[2] = estr[3] = 1;
  } else {
    pick_tile(g.tile_W, g.tile_H, &ka.TW, &ka.TH);
    ka.Ho = g.tile_H; ka.Wo = g.tile_W;
    ka.tiles_w = (g.tile_W + ka.TW - 1) / ka.TW; ka.tiles_h = (g.tile_H + ka.TH - 1) / ka.TH;
    nimg = g.aN;
    gdim[0] = g.aC; gdim[1] = g.aW; gdim[2] = g.aH; gdim[3] = g.aN;
    gstr[0] = (cuuint64_t)g.a_cstride * 2; gstr[1] = gstr[0] * g.aW; gstr[2] = gstr[1] * g.aH;
    // with element strides the box is measured in input elements: ceil(box/stride) elements are loaded
    box[0] = 64; box[1] = (cuuint32_t)(ka.TW * ka.stride); box[2] = (cuuint32_t)(ka.TH * ka.stride); box[3] = 1;
    estr[0] = 1; estr[1] = (cuuint32_t)ka.stride; estr[2] = (cuuint32_t)ka.stride; estr[3] = 1;
    ETB_CHECK_ARG(box[1] <= 256 && box[2] <= 256);
  }
  CUtensorMap mA, mB;
  CUresult r = enc(&mA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(g.a_ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    etb_set_error("cuTensorMapEncodeTiled(A) failed: %d", (int)r);
    return ETB_ERR_CUDA;
  }
  const long Ktot = (long)ka.ntaps * aCp;
  const int BN = g.b_rows > 128 ? 256 : (g.b_rows > 64 ? 128 : 64);
  // cta_group::2 (the pair halves each SM's weight ingest): always for the 256-wide tile; for the 128-wide tile on the
  // multi-tap layers only (3x3 128->128 @80: 71 -> 64 us forward, 70 -> 62 us dgrad; the pointwise 128-wide layers are
  // equal or slower: tools/conv_bench.py, profiles/r2_conv_bench_2sm*.json).  ETB_CONV_2SM = bit mask (1: 256-wide,
  // 2: 128-wide) overrides the rule; 0 disables the 2-SM path.
  static int two_sm = -2;
  if (two_sm == -2) { const char* e = getenv("ETB_CONV_2SM"); two_sm = e ? atoi(e) : -1; }
  const long m_tiles_all = (long)ka.tiles_w * ka.tiles_h * nimg;
  const bool want2 = two_sm < 0 ? (BN == 256 || (BN == 128 && ka.ntaps > 1)) : ((BN == 256 && (two_sm & 1)) || (BN == 128 && (two_sm & 2)));
  const bool use2 = want2 && m_tiles_all >= 2;
  cuuint64_t wdim[2] = {(cuuint64_t)Ktot, (cuuint64_t)g.b_rows};
  cuuint64_t wstr[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t wbox[2] = {64, (cuuint32_t)(use2 ? BN / 2 : BN)};
  cuuint32_t westr[2] = {1, 1};
  r = enc(&mB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(g.b_ptr), wdim, wstr, wbox, westr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    etb_set_error("cuTensorMapEncodeTiled(B) failed: %d", (int)r);
    return ETB_ERR_CUDA;
  }
  ka.kblocks = aCp / CONV_BLOCK_K;
  ka.Cout = g.b_rows;
  ka.nimg = nimg;
  if (use2) {
    const long pair_tiles = ((m_tiles_all + 1) / 2) * ((g.b_rows + BN - 1) / BN);
    const long clusters = etb_num_sms() / 2;
    dim3 grid2((unsigned)(2 * (pair_tiles < clusters ? pair_tiles : clusters)), 1);
    return BN == 256 ? launch_conv2<256, 6>(mA, mB, ka, grid2, st) : launch_conv2<128, 8>(mA, mB, ka, grid2, st);
  }
  const long total_tiles = (long)ka.tiles_w * ka.tiles_h * nimg * ((g.b_rows + BN - 1) / BN);
  const long resident = (long)etb_num_sms();   // persistent: one CTA per SM; overlap comes from the 2 TMEM accumulators
  dim3 grid((unsigned)(total_tiles < resident ? total_tiles : resident), 1);
  if (BN == 256) return launch_conv<256, 4>(mA, mB, ka, grid, st);
  if (BN == 128) return launch_conv<128, 6>(mA, mB, ka, grid, st);
  return launch_conv<64, 8>(mA, mB, ka, grid, st);
}

extern "C" size_t etb_conv_workspace_bytes(const EtbConvParams* cp) { (void)cp; return 0; }

extern "C" int etb_conv_fwd(const void* x_bf16, const void* w_bf16, const float* scale, const float* bias,
                            const void* residual_bf16, void* y_bf16, float* y_f32, const EtbConvParams* cp, void* workspace,
                            size_t workspace_bytes, void* stream) {
  (void)workspace; (void)workspace_bytes;
  ETB_CHECK_ARG(x_bf16 && w_bf16 && cp && (y_bf16 || y_f32));
  ETB_CHECK_ARG(cp->N > 0 && cp->H > 0 && cp->W > 0 && cp->Cin > 0 && cp->Cout > 0);
  ETB_CHECK_ARG(cp->kh >= 1 && cp->kw >= 1 && cp->kh * cp->kw <= 12 && (cp->stride == 1 || cp->stride == 2) && cp->pad >= 0);
  const int Ho = (cp->H + 2 * cp->pad - cp->kh) / cp->stride + 1;
  const int Wo = (cp->W + 2 * cp->pad - cp->kw) / cp->stride + 1;
  ETB_CHECK_ARG(Ho > 0 && Wo > 0);
  const bool det = (y_f32 != nullptr);
  if (!det) ETB_CHECK_ARG(cp->y_cstride % 8 == 0 && cp->y_coffset % 8 == 0 && cp->y_cstride >= cp->y_coffset + cp->Cout && (((uintptr_t)y_bf16) & 15) == 0);
  if (residual_bf16) ETB_CHECK_ARG(!det && cp->res_cstride % 8 == 0 && cp->res_coffset % 8 == 0 && cp->Cout % 8 == 0);
  if (det) ETB_CHECK_ARG(cp->det_no > 0 && cp->Cout % cp->det_no == 0);
  ConvKArgs ka;
  memset(&ka, 0, sizeof(ka));
  GemmGeom g;
  g.a_ptr = x_bf16; g.aN = cp->N; g.aH = cp->H; g.aW = cp->W; g.aC = cp->Cin; g.a_cstride = cp->x_cstride;
  g.b_ptr = w_bf16; g.b_rows = cp->Cout;
  g.tile_H = Ho; g.tile_W = Wo;
  g.flat = (cp->kh == 1 && cp->kw == 1 && cp->stride == 1 && cp->pad == 0);
  ka.ntaps = cp->kh * cp->kw;
  for (int kh = 0; kh < cp->kh; ++kh)
    for (int kw = 0; kw < cp->kw; ++kw) {
      ka.tap_dh[kh * cp->kw + kw] = (signed char)(kh - cp->pad);
      ka.tap_dw[kh * cp->kw + kw] = (signed char)(kw - cp->pad);
    }
  ka.stride = cp->stride;
  ka.out_os = 1; ka.out_ph = ka.out_pw = 0; ka.out_H = Ho; ka.out_W = Wo;
  ka.y_cstride = cp->y_cstride; ka.y_coffset = cp->y_coffset;
  ka.res_cstride = cp->res_cstride; ka.res_coffset = cp->res_coffset;
  ka.act = cp->act;
  ka.out_mode = det ? 1 : 0;
  ka.det_no = cp->det_no;
  ka.det_hw = Ho * Wo;
  ka.scale = scale; ka.bias = bias;
  ka.residual = (const __nv_bfloat16*)residual_bf16;
  ka.y = (__nv_bfloat16*)y_bf16;
  ka.y_f32 = y_f32;
  return launch_gemm(g, ka, (cudaStream_t)stream);
}

// ---- data gradient (K2): dx = conv_transpose(dy, W) as implicit GEMMs on the same kernel ------------------------------
// For each output-parity class (ph,pw) of dx (one class when stride==1):  dx[n, s*i+ph, s*j+pw, :] =
//   sum over the taps (kh,kw) with (ph+pad-kh) % s == 0, (pw+pad-kw) % s == 0 of  dy[n, i+dh, j+dw, :] * W[:, :, kh, kw]
//   with dh = (ph+pad-kh)/s, dw = (pw+pad-kw)/s.   B operand: etb_pack_weight_dgrad (same tap order).
static int dgrad_taps(int k, int s, int pad, int ph, int pw, signed char* kh_l, signed char* kw_l, signed char* dh, signed char* dw) {
  int n = 0;
  for (int kh = 0; kh < k; ++kh) {
    if ((ph + pad - kh) % s != 0) continue;
    for (int kw = 0; kw < k; ++kw) {
      if ((pw + pad - kw) % s != 0) continue;
      kh_l[n] = (signed char)kh; kw_l[n] = (signed char)kw;
      // floor division is exact here (remainder checked); C division of negatives truncates toward zero, also exact
      dh[n] = (signed char)((ph + pad - kh) / s); dw[n] = (signed char)((pw + pad - kw) / s);
      ++n;
    }
  }
  return n;
}

extern "C" int64_t etb_dgrad_weight_elems(int32_t Cout, int32_t Cin, int32_t k, int32_t stride) {
  (void)stride;
  const int64_t Coutp = (Cout + CONV_BLOCK_K - 1) / CONV_BLOCK_K * CONV_BLOCK_K;   // every tap padded to the 64-channel K block
  return (int64_t)Cin * k * k * Coutp;   // all parity classes together visit every tap exactly once
}

struct TapTable { signed char v[24]; };
__global__ void __launch_bounds__(256) pack_weight_dgrad_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ o, int Cout, int Coutp, int Cin, int k, int ntaps, TapTable tt) {
  ETB_PDL_PROLOGUE();
  const int64_t total = (int64_t)Cin * ntaps * Coutp;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(e % Coutp);
    const int t = (int)((e / Coutp) % ntaps);
    const int ci = (int)(e / ((int64_t)Coutp * ntaps));
    o[e] = __float2bfloat16(co < Cout ? w[(((int64_t)co * Cin + ci) * k + tt.v[t]) * k + tt.v[12 + t]] : 0.f);
  }
}

// w [Cout,Cin,k,k] fp32 -> for each parity class c (row-major ph,pw) a [Cin][ntaps_c*ceil64(Cout)] bf16 block (zero padded), blocks concatenated
extern "C" int etb_pack_weight_dgrad(const float* w_oihw, void* out_bf16, int32_t Cout, int32_t Cin, int32_t k, int32_t stride,
                                     int32_t pad, void* stream) {
  ETB_CHECK_ARG(w_oihw && out_bf16 && Cout > 0 && Cin > 0 && k >= 1 && k * k <= 12 && (stride == 1 || stride == 2));
  __nv_bfloat16* o = (__nv_bfloat16*)out_bf16;
  for (int ph = 0; ph < stride; ++ph)
    for (int pw = 0; pw < stride; ++pw) {
      TapTable tt;
      signed char dh[12], dw[12];
      const int nt = dgrad_taps(k, stride, pad, ph, pw, tt.v, tt.v + 12, dh, dw);
      if (nt == 0) continue;
      const int Coutp = (Cout + CONV_BLOCK_K - 1) / CONV_BLOCK_K * CONV_BLOCK_K;
      const int64_t total = (int64_t)Cin * nt * Coutp;
      int64_t blocks = (total + 255) / 256;
      if (blocks > 148 * 16) blocks = 148 * 16;
      etb_launch(pack_weight_dgrad_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, w_oihw, o, Cout, Coutp, Cin, k, nt, tt);
      ETB_CHECK_LAUNCH();
      o += total;
    }
  return ETB_OK;
}

// dy [N,Ho,Wo,*] bf16 (channels [dy_coffset.. +Cout) of stride dy_cstride) -> dx [N,H,W,*] bf16 at channel offset.
// cp describes the FORWARD conv (N,H,W,Cin,Cout,k,stride,pad); x_cstride/ y_* name the dy / dx buffers:
//   cp->x_cstride = channel stride of dy, cp->y_cstride/y_coffset = geometry of dx.  cp->act==3 -> dx += (accumulate).
extern "C" int etb_conv_dgrad(const void* dy_bf16, const void* wd_bf16, void* dx_bf16, const EtbConvParams* cp, int32_t accumulate,
                              void* stream) {
  ETB_CHECK_ARG(dy_bf16 && wd_bf16 && dx_bf16 && cp);
  ETB_CHECK_ARG(cp->kh == cp->kw && cp->kh * cp->kw <= 12 && (cp->stride == 1 || cp->stride == 2));
  ETB_CHECK_ARG(cp->Cout % 8 == 0 && cp->y_cstride % 8 == 0 && cp->y_coffset % 8 == 0 && cp->y_cstride >= cp->y_coffset + cp->Cin);
  const int Coutp = (cp->Cout + CONV_BLOCK_K - 1) / CONV_BLOCK_K * CONV_BLOCK_K;   // row pitch of one tap in the packed operand
  const int k = cp->kh, s = cp->stride, pad = cp->pad;
  const int Ho = (cp->H + 2 * pad - k) / s + 1, Wo = (cp->W + 2 * pad - k) / s + 1;
  const __nv_bfloat16* wd = (const __nv_bfloat16*)wd_bf16;
  for (int ph = 0; ph < s; ++ph)
    for (int pw = 0; pw < s; ++pw) {
      ConvKArgs ka;
      memset(&ka, 0, sizeof(ka));
      signed char kh_l[12], kw_l[12];
      const int nt = dgrad_taps(k, s, pad, ph, pw, kh_l, kw_l, ka.tap_dh, ka.tap_dw);
      const int subH = (cp->H - ph + s - 1) / s, subW = (cp->W - pw + s - 1) / s;
      if (subH <= 0 || subW <= 0) continue;
      ETB_CHECK_ARG(nt > 0);   // k >= stride for every conv of the trunk, so every parity class is reached
      GemmGeom g;
      g.a_ptr = dy_bf16; g.aN = cp->N; g.aH = Ho; g.aW = Wo; g.aC = cp->Cout; g.a_cstride = cp->x_cstride;
      g.b_ptr = wd; g.b_rows = cp->Cin;
      g.tile_H = subH; g.tile_W = subW;
      g.flat = (k == 1 && s == 1 && pad == 0);
      ka.ntaps = nt;
      ka.stride = 1;
      ka.out_os = s; ka.out_ph = ph; ka.out_pw = pw; ka.out_H = cp->H; ka.out_W = cp->W;
      ka.y_cstride = cp->y_cstride; ka.y_coffset = cp->y_coffset;
      ka.act = 0; ka.out_mode = 0; ka.accumulate = accumulate;
      ka.y = (__nv_bfloat16*)dx_bf16;
      int rc = launch_gemm(g, ka, (cudaStream_t)stream);
      if (rc != ETB_OK) return rc;
      wd += (size_t)cp->Cin * nt * Coutp;
    }
  return ETB_OK;
}

// =====================================================================================================================
// weight gradient (K2):  dW[co][tap][ci] = sum over pixels  dy[n,oh,ow,co] * x[n, oh*s+kh-p, ow*s+kw-p, ci]
//
// GEMM with the PIXELS as the reduction (K) dimension: D[128 co, BN ci] += A[128 co, kpix]*B[BN ci, kpix]^T where both
// operands are "MN-major" (the channel index is the contiguous one in NHWC).  Each K block is one spatial tile of one
// image, fetched for both operands by 4-D TMA boxes {64 ch, TW, TH, 1} (x with the tap shift / element strides, zero
// padding = OOB fill) into 128B-swizzled smem = the canonical MN-major UMMA layout:
//   64-channel group = kpix rows x 128 B;  8-row K atoms 1024 B apart (SBO);  channel groups one region apart (LBO).
// One CTA owns one (co tile, ci tile, tap) and a contiguous slice of the K blocks (split-K across gridDim.y); every CTA
// stores its partial tile into its own slice of the workspace (plain coalesced stores, no atomics, no memset) and a second
// kernel (wgrad_reduce_kernel / wgrad_reduce_taps_kernel) sums the slices in a fixed order into the parameter layout.
// =====================================================================================================================
struct WgradArgs {
  int ntaps;
  signed char tap_dh[12], tap_dw[12];
  int stride;
  int TW, TH, tiles_w, tiles_h, nimg;
  int kpix;                 // TW*TH, multiple of 16, <= 128
  int co_tiles, ci_tiles;
  int Cout, Cin;
  int flat;
  float* dw;                // split-K partials: slice `blockIdx.y` of [splitk][Cout][ntaps][Cin] fp32
  long dw_split_stride;     // elements per slice
};

// Partial-tile store of one accumulator (this warp's 32 rows x NCOLS fp32 columns) through the staging tile: per 32-column
// chunk each lane parks its row (128 B), then store instruction i writes rows 4i..4i+3 x 128 contiguous bytes.
// dst = this lane's row; rows are row_stride floats apart; row r is valid while co - lane + r < Cout; columns while < ncols_ok.
template <int NCOLS>
__device__ __forceinline__ void wgrad_store_tile(uint32_t lane_addr, float* dst, size_t row_stride, int co, int Cout, int ncols_ok, uint8_t* tile,
                                                 int lane) {
  const int co_base = co - lane;
  float* base = dst - (size_t)lane * row_stride;              // row 0 of this warp's 32
#pragma unroll 1
  for (int c0 = 0; c0 < NCOLS; c0 += 32) {
    if (c0 >= ncols_ok) break;                                // warp-uniform (Cin is a multiple of 8; tails are whole chunks)
    uint32_t v[32];
    tmem_ld32(lane_addr + (uint32_t)c0, v);
#pragma unroll
    for (int q = 0; q < 8; ++q) stage_write(tile, lane, q, make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + (lane >> 3), c = (lane & 7) * 4;
      const uint4 o = stage_read(tile, lane, i);
      if (co_base + r < Cout && c0 + c < ncols_ok) *reinterpret_cast<uint4*>(base + (size_t)r * row_stride + c0 + c) = o;
    }
    __syncwarp();
  }
}

// Tile = (128*MT co) x (BN ci) per CTA, K block = up to KP pixels.  Bigger tiles raise the FLOP per L2 byte
// (128x128: 65, 256x256: 131 flop/B) -- the kernel is L2-bandwidth bound, not MMA bound.
template <int MT, int BN, int KP, int STAGES>
struct WgradSmem {
  static constexpr int GROUP_BYTES = KP * 128;            // one 64-channel group, up to KP K rows
  static constexpr int A_BYTES = 2 * MT * GROUP_BYTES;    // 128*MT co
  static constexpr int B_BYTES = (BN / 64) * GROUP_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
  static constexpr int TMEM_COLS = MT * BN;
};

// MN-major SWIZZLE_128B descriptor: LBO = byte distance between 64-element channel groups, SBO = 1024 (8 K rows)
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N) {   // both operands MN-major (bits 15, 16)
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// CL = 1: 2x2 thread-block cluster (2 co tiles x 2 ci tiles of the same tap and K slice).  Each CTA fetches only ONE
// 64-channel group of dy and ONE of x per K block and TMA-multicasts it to the CTA that shares that operand, halving the
// L2 -> SM operand traffic (65 -> 131 FLOP per L2 byte).  A stage is recycled only when this CTA's MMAs and those of the
// two CTAs it multicasts into have consumed it (empty barrier count 3, multicast tcgen05.commit).
template <int MT, int BN, int KP, int STAGES, int CL>
__global__ void __launch_bounds__(WGRAD_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap mapDy, const __grid_constant__ CUtensorMap mapX, const WgradArgs a) {
  static_assert(CL == 0 || (MT == 1 && BN == 128), "cluster variant: 128x128 tiles");
  using L = WgradSmem<MT, BN, KP, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + L::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int t = blockIdx.x;
  int tap, ci_t, co_t;
  uint32_t crank = 0;
  if (CL) {
    crank = cluster_ctarank();           // == blockIdx.x & 3 for a (4,1,1) cluster
    t >>= 2;
    tap = t % a.ntaps; t /= a.ntaps;
    const int cip = a.ci_tiles >> 1;
    ci_t = (t % cip) * 2 + (int)(crank & 1); t /= cip;
    co_t = t * 2 + (int)(crank >> 1);
  } else {
    tap = t % a.ntaps; t /= a.ntaps;
    ci_t = t % a.ci_tiles; t /= a.ci_tiles;
    co_t = t;
  }
  const int co0 = co_t * 128 * MT, ci0 = ci_t * BN;
  const uint16_t mask_a = (uint16_t)((1u << crank) | (1u << (crank ^ 1u)));   // CTAs with the same co tile (share dy)
  const uint16_t mask_b = (uint16_t)((1u << crank) | (1u << (crank ^ 2u)));   // CTAs with the same ci tile (share x)
  const int total_kb = a.nimg * a.tiles_h * a.tiles_w;
  const int chunk = (total_kb + gridDim.y - 1) / gridDim.y;
  const int kb0 = blockIdx.y * chunk;
  const int kb1 = min(total_kb, kb0 + chunk);
  if (kb0 >= kb1) return;   // uniform for the CTA, before any barrier / allocation
  const int kiters = kb1 - kb0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapDy);
    tma_prefetch_desc(&mapX);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL ? 3 : 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, L::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (CL) cluster_sync_all();            // every CTA's barriers are initialised before any remote complete_tx / arrive
  const uint32_t tmem_base = *tmem_slot;
  ETB_PDL_PROLOGUE();      // prologue above overlapped the previous kernel's tail; no global data touched before this
  const uint32_t box_bytes = (uint32_t)a.kpix * 128u;   // bytes one TMA box writes (all rows, OOB rows zero-filled)

  if (warp == 0) {
    // warp-uniform loop, one elected lane issues (see elect_one)
    // K-block coordinates, stage and phase advance incrementally: the producer is ONE warp of dependent integer code, and two
    // runtime divisions per K block made it (not the tensor pipe, not L2) the bound of every weight-gradient tile shape
    // (ncu source page: the producer never waited on an empty stage, the MMA warp always waited on a full one)
    int tw_i = kb0 % a.tiles_w, th_i = (kb0 / a.tiles_w) % a.tiles_h, img = kb0 / (a.tiles_w * a.tiles_h);
    const int xdw = a.tap_dw[tap], xdh = a.tap_dh[tap];
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < kiters; ++it) {
      const int w0 = tw_i * a.TW, h0 = th_i * a.TH;
      mbar_wait(&empty[s], ph ^ 1u);
      uint8_t* sa = smem + s * L::STAGE_BYTES;
      uint8_t* sb = sa + L::A_BYTES;
      if (elect_one()) {
        mbar_expect_tx(&full[s], box_bytes * (2u * MT + BN / 64));
        if (CL) {
          const int ga = (int)(crank & 1), gb = (int)(crank >> 1);   // my share: one dy group, one x group
          tma_load_4d_mc(&mapDy, &full[s], sa + ga * L::GROUP_BYTES, co0 + 64 * ga, w0, h0, img, mask_a);
          tma_load_4d_mc(&mapX, &full[s], sb + gb * L::GROUP_BYTES, ci0 + 64 * gb, w0 * a.stride + xdw, h0 * a.stride + xdh, img, mask_b);
        } else {
#pragma unroll
          for (int g = 0; g < 2 * MT; ++g) tma_load_4d(&mapDy, &full[s], sa + g * L::GROUP_BYTES, co0 + 64 * g, w0, h0, img);
#pragma unroll
          for (int g = 0; g < BN / 64; ++g)
            tma_load_4d(&mapX, &full[s], sb + g * L::GROUP_BYTES, ci0 + 64 * g, w0 * a.stride + xdw, h0 * a.stride + xdh, img);
        }
      }
      __syncwarp();
      if (++tw_i == a.tiles_w) { tw_i = 0; if (++th_i == a.tiles_h) { th_i = 0; ++img; } }
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16_mn(128, BN);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int ksteps = a.kpix / 16;
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < kiters; ++it) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
      const uint64_t bdesc = make_mnmajor_sw128_desc(sa + L::A_BYTES, L::GROUP_BYTES);
      if (elect_one()) {
        // k outer, m tile inner: consecutive MMAs go to different accumulators
        for (int k = 0; k < ksteps; ++k) {   // 16 K rows = two 1024 B atoms per UMMA_K step
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t adesc = make_mnmajor_sw128_desc(sa + mt * 2 * L::GROUP_BYTES, L::GROUP_BYTES);
            umma_bf16(tmem_u + (uint32_t)(mt * BN), adesc + (uint64_t)(k * (2048 >> 4)), bdesc + (uint64_t)(k * (2048 >> 4)), idesc,
                      (it | k) != 0 ? 1u : 0u);
          }
        }
        if (CL) umma_commit_mc(&empty[s], (uint16_t)(mask_a | mask_b));   // me + the two CTAs that write into my stage
        else umma_commit(&empty[s]);
        if (it == kiters - 1) umma_commit(tmem_full);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
  } else {
    const int row = 32 * (warp & 3) + lane;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    // every MMA has completed (tmem_full) and every TMA box was consumed by one: the stage ring is free -> store staging
    uint8_t* tile = smem + (warp - 2) * EPI_TILE_BYTES;
    const size_t row_stride = (size_t)a.ntaps * a.Cin;        // floats between two co rows of the partial tile
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      const int co = co0 + mt * 128 + row;
      const uint32_t lane_addr = tmem_base + (uint32_t)(mt * BN) + ((uint32_t)(32 * (warp & 3)) << 16);
      // two-stage split-K: every CTA writes its partial tile with plain coalesced stores into its own slice; the
      // slices are summed (and laid out as [Cout,Cin,kh,kw]) by wgrad_reduce_kernel.  No atomics, no memset.
      float* dst = a.dw + (size_t)blockIdx.y * a.dw_split_stride + ((size_t)co * a.ntaps + tap) * a.Cin + ci0;
      wgrad_store_tile<BN>(lane_addr, dst, row_stride, co, a.Cout, a.Cin - ci0, tile, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL) cluster_sync_all();            // no CTA leaves while a peer may still multicast into it / arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, L::TMEM_COLS);
  }
}

static void pick_tile16(int Wo, int Ho, int maxrows, int* TW, int* TH) {
  // K tiles must be a whole number of UMMA_K steps: TW*TH % 16 == 0, <= maxrows; out-of-image rows are zero-filled by TMA
  double best = -1.0;
  for (int tw = 1; tw <= maxrows; ++tw)
    for (int th = 1; th * tw <= maxrows; ++th) {
      if ((tw * th) % 16) continue;
      if (tw > 2 * Wo || th > 2 * Ho) continue;
      const long tiles = (long)((Wo + tw - 1) / tw) * ((Ho + th - 1) / th);
      const double eff = (double)Wo * Ho / ((double)tiles * tw * th) * (tw * th >= maxrows / 2 ? 1.0 : 0.8);
      if (eff > best + 1e-9) { best = eff; *TW = tw; *TH = th; }
    }
}

template <int MT, int BN, int KP, int STAGES, int CL>
static int launch_wgrad(const CUtensorMap& mDy, const CUtensorMap& mX, const WgradArgs& wa, dim3 grid, cudaStream_t st) {
  using L = WgradSmem<MT, BN, KP, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    ETB_CHECK_CUDA(cudaFuncSetAttribute(wgrad_kernel<MT, BN, KP, STAGES, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  if (CL) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(WGRAD_THREADS); cfg.dynamicSmemBytes = L::TOTAL; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = etb_pdl_enabled() ? 2 : 1;
    ETB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, wgrad_kernel<MT, BN, KP, STAGES, CL>, mDy, mX, wa));
    etb_count_launch();
    return ETB_OK;
  }
  etb_launch(wgrad_kernel<MT, BN, KP, STAGES, CL>, dim3(grid), dim3(WGRAD_THREADS), L::TOTAL, st, mDy, mX, wa);
  ETB_CHECK_LAUNCH();
  return ETB_OK;
}

// ---- second stage of the split-K: sum the slices, emit the parameter layout, optionally accumulate ----
// block = 32 lanes (each 4 consecutive ci of one (co, tap): coalesced float4 reads of every slice) x SG slice groups;
// group g sums slices g, g+SG, ... (4 independent loads in flight), the groups are combined through shared memory in a
// fixed order (deterministic).  SG grows with the split count so a 148-way split of a small layer is not one serial chain.
template <int SG>
__global__ void __launch_bounds__(32 * SG) wgrad_reduce_kernel(const float* __restrict__ ws, long slice, int splitk, float* __restrict__ out, int Cout,
                                                               int Cin, int kk, int flags) {
  ETB_PDL_PROLOGUE();
  __shared__ float4 red[SG][32];
  const long n4 = (long)Cout * kk * Cin / 4;
  const int lane = threadIdx.x, sg = threadIdx.y;
  for (long e4 = (long)blockIdx.x * 32 + lane; e4 - lane < n4; e4 += (long)gridDim.x * 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e4 < n4) {
      const float4* p = reinterpret_cast<const float4*>(ws) + e4;
      const size_t st4 = (size_t)slice / 4;
      int sidx = sg;
      for (; sidx + 3 * SG < splitk; sidx += 4 * SG) {
        const float4 v0 = __ldg(p + (size_t)sidx * st4), v1 = __ldg(p + (size_t)(sidx + SG) * st4);
        const float4 v2 = __ldg(p + (size_t)(sidx + 2 * SG) * st4), v3 = __ldg(p + (size_t)(sidx + 3 * SG) * st4);
        acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
        acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
      }
      for (; sidx < splitk; sidx += SG) {
        const float4 v = __ldg(p + (size_t)sidx * st4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    if (SG > 1) {
      red[sg][lane] = acc;
      __syncthreads();
      if (sg == 0) {
#pragma unroll
        for (int g = 1; g < SG; ++g) {
          const float4 v = red[g][lane];
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
    }
    if (sg == 0 && e4 < n4) {
      const long e = e4 * 4;
      const int ci = (int)(e % Cin);
      const long t2 = e / Cin;
      const int t = (int)(t2 % kk), co = (int)(t2 / kk);
      const float vals[4] = {acc.x, acc.y, acc.z, acc.w};
      if (kk == 1 && !(flags & 1) && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {   // pointwise: the GEMM layout IS the parameter layout
        float4* o = reinterpret_cast<float4*>(out + e);
        if (flags & 2) { const float4 q = *o; acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w; }
        *o = acc;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          long dst;
          if (flags & 1) {                       // stem: [Cout][128 slots, k = (c*6+kh)*6+kw] -> [Cout,3,6,6] (same order)
            const int k = ci + j;
            if (k >= 108) continue;
            dst = (long)co * 108 + k;
          } else {
            dst = ((long)co * Cin + ci + j) * kk + t;
          }
          out[dst] = (flags & 2) ? out[dst] + vals[j] : vals[j];
        }
      }
    }
    if (SG > 1) __syncthreads();
  }
}

// kk > 1 (3x3 ...): the GEMM layout [co][tap][ci] has to become the parameter layout [co][ci][tap].  One block = one co
// and 64 consecutive ci: unit (tap, lane) sums its float4 over the slices (coalesced along ci), the [64][kk] patch is
// transposed through shared memory and written (or accumulated) as ONE contiguous run of 64*kk floats -- the naive
// scatter cost 8x sector amplification on both the read-modify-write and the store (60 us per 3x3 layer).
__global__ void __launch_bounds__(256) wgrad_reduce_taps_kernel(const float* __restrict__ ws, long slice, int splitk, float* __restrict__ out, int Cin,
                                                                int kk, int flags) {
  ETB_PDL_PROLOGUE();
  __shared__ float sm[64 * 12];
  const int CW = Cin < 64 ? Cin : 64;            // channels per block (Cin is a multiple of 64, or smaller than 64 and of 8)
  const int L4 = CW >> 2;                        // float4 lanes per tap
  const int chunks = Cin / CW;
  const int co = blockIdx.x / chunks, ci0 = (blockIdx.x - co * chunks) * CW;
  const size_t st4 = (size_t)slice / 4;
  for (int u = threadIdx.x; u < kk * L4; u += 256) {
    const int t = u / L4, lane = u - t * L4;
    const float4* p = reinterpret_cast<const float4*>(ws + ((size_t)co * kk + t) * Cin + ci0) + lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int sidx = 0;
    for (; sidx + 3 < splitk; sidx += 4) {
      const float4 v0 = __ldg(p + (size_t)sidx * st4), v1 = __ldg(p + (size_t)(sidx + 1) * st4);
      const float4 v2 = __ldg(p + (size_t)(sidx + 2) * st4), v3 = __ldg(p + (size_t)(sidx + 3) * st4);
      acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
      acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; sidx < splitk; ++sidx) {
      const float4 v = __ldg(p + (size_t)sidx * st4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    sm[(4 * lane + 0) * kk + t] = acc.x;
    sm[(4 * lane + 1) * kk + t] = acc.y;
    sm[(4 * lane + 2) * kk + t] = acc.z;
    sm[(4 * lane + 3) * kk + t] = acc.w;
  }
  __syncthreads();
  float* o = out + ((size_t)co * Cin + ci0) * kk;
  for (int i = threadIdx.x; i < CW * kk; i += 256) o[i] = (flags & 2) ? o[i] + sm[i] : sm[i];
}

// =====================================================================================================================
// wgrad2: 2-SM (cta_group::2) weight gradient with several accumulators per tile -- the L2-traffic-optimal variant for the
// wide layers (Cout >= 256, Cin >= 128), which carry most of the weight-gradient FLOPs.
//   The 1-SM kernel above is bound by the L2 -> SM operand stream: a 128(co) x 128(ci) x 1-tap CTA ingests 64 KB per 128-pixel
//   K block for 512 MMA cycles = 128 B/clk/SM, ~3x what L2 sustains with all SMs pulling.  Here a cluster of two CTAs owns
//   256 co x 128 ci x NT "virtual columns" (a virtual column = one (ci tile, filter tap) pair; 3 taps of one filter row for
//   the 3x3 layers, up to 4 ci tiles for the pointwise ones): per 64-pixel K block each SM loads ITS 128 co of dy (16 KB, used
//   by all NT accumulators) and its 64-ci HALF of each of the NT x tiles (8 KB each), and the pair issues NT x 4
//   tcgen05.mma.cta_group::2 (M=256, N=128, K=16).  Bytes per MMA cycle and SM: (16 + 8 NT) KB / (256 NT) clk = 53 B/clk at
//   NT = 3, 47 B/clk at NT = 4 -- 2.4-2.7x less than the 1-SM kernel.  TMEM: NT x 128 fp32 columns per CTA (<= 512).
//   Both operands are MN-major straight from the NHWC tensors (LBO = 64-channel group stride), split-K over the pixel blocks
//   with per-slice partial tiles in the workspace exactly like the 1-SM kernel (same reduce kernels).
//   Barrier protocol = conv_fwd2_kernel's: both producers complete_tx on the LEADER's full barrier, tcgen05.commit multicasts
//   to both CTAs' empty / tmem_full barriers.
// =====================================================================================================================
struct Wgrad2Args {
  int ntaps;
  signed char tap_dh[12], tap_dw[12];
  int stride;
  int TW, TH, tiles_w, tiles_h, nimg;
  int kpix;                 // TW*TH, multiple of 16, <= KP
  int NT;                   // virtual columns per cluster tile
  int nvirt;                // virtual columns of the layer
  int groups;               // ceil(nvirt / NT)
  int cit;                  // NW = 256: Cin / 256 (ci tiles per tap)
  int Cout, Cin;
  float* dw;
  long dw_split_stride;
};

template <int KP, int STAGES>
struct Wgrad2Smem {
  static constexpr int GROUP_BYTES = KP * 128;            // one 64-channel group, KP pixel rows
  static constexpr int A_BYTES = 2 * GROUP_BYTES;         // this CTA's 128 co of dy
  static constexpr int B_MAX = 4 * GROUP_BYTES;           // NT virtual columns x this CTA's NW/2 ci = up to four 64-channel groups
  static constexpr int STAGE_BYTES = A_BYTES + B_MAX;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

// virtual column -> (filter tap, first input channel).  NW = 128: a (128-ci tile, tap) pair, taps fastest (a tile of NT = 3
// columns is one filter row of one ci tile); NW = 256: a (tap, 256-ci tile) pair, ci tiles fastest (Cin % 256 == 0).
template <int NW>
__device__ __forceinline__ void wgrad2_vcol(const Wgrad2Args& a, int v, int* tap, int* cbase) {
  if (NW == 128) {
    const int ci_t = v / a.ntaps;
    *tap = v - ci_t * a.ntaps;
    *cbase = ci_t * 128;
  } else {
    const int tp = v / a.cit;
    *tap = tp;
    *cbase = (v - tp * a.cit) * 256;
  }
}

// NW = MMA N = accumulator width.  tools/mma_probe.cu: tcgen05.mma (SS) reaches the tensor-pipe floor for N >= 128 when the
// issuing warp is converged, but the operand feed does not: per 64-pixel K block and SM the NW = 128 tile ingests
// (16 + 8 NT) KB for 256 NT MMA cycles (53 B/clk at NT = 3), the NW = 256 tile (16 + 16 NT) KB for 512 NT cycles (47 B/clk at
// NT = 1, 31 B/clk at NT = 2) -- and the wide MMA halves the instructions per FLOP.
template <int KP, int STAGES, int NW>
__global__ void __launch_bounds__(WGRAD_THREADS, 1)
wgrad2_kernel(const __grid_constant__ CUtensorMap mapDy, const __grid_constant__ CUtensorMap mapX, const Wgrad2Args a) {
  using L = Wgrad2Smem<KP, STAGES>;
  constexpr int GPC = NW / 128;                      // 64-channel x groups per virtual column and CTA
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + L::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();           // 0 = leader
  const int ct = blockIdx.x >> 1;                    // cluster tile: virtual-column group fastest, then the co pair
  const int grp = ct % a.groups;
  const int cop = ct / a.groups;
  const int v0 = grp * a.NT;
  const int nt = min(a.NT, a.nvirt - v0);            // virtual columns of this tile (the last group may be short)
  const int co0 = cop * 256 + (int)rank * 128;       // this CTA's 128 co rows
  const int acc_cols = a.NT * NW;
  const uint32_t tmem_cols = acc_cols <= 128 ? 128u : (acc_cols <= 256 ? 256u : 512u);

  const int total_kb = a.nimg * a.tiles_h * a.tiles_w;
  const int chunk = (total_kb + gridDim.y - 1) / gridDim.y;
  const int kb0 = blockIdx.y * chunk;
  const int kb1 = min(total_kb, kb0 + chunk);
  const int kiters = max(kb1 - kb0, 0);              // uniform for the cluster (host guarantees no empty slice)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapDy);
    tma_prefetch_desc(&mapX);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 2); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ETB_PDL_PROLOGUE();
  const uint32_t box_bytes = (uint32_t)a.kpix * 128u;

  if (warp == 0) {
    // warp-uniform loop, one elected lane issues (see elect_one)
    // per-column x coordinates are fixed for the whole K loop: hoisted into registers (statically indexed, MAXT columns);
    // K-block coordinates, stage and phase advance incrementally (no divisions, no local-memory table reads in the loop)
    constexpr int MAXT = 512 / NW;
    int xc[MAXT], xdw[MAXT], xdh[MAXT];
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      int tap = 0, cbase = 0;
      if (t < nt) wgrad2_vcol<NW>(a, v0 + t, &tap, &cbase);
      xc[t] = cbase + (NW / 2) * (int)rank;
      xdw[t] = a.tap_dw[tap];
      xdh[t] = a.tap_dh[tap];
    }
    int tw_i = kb0 % a.tiles_w, th_i = (kb0 / a.tiles_w) % a.tiles_h, img = kb0 / (a.tiles_w * a.tiles_h);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < kiters; ++it) {
      const int w0 = tw_i * a.TW, h0 = th_i * a.TH;
      const int xw0 = w0 * a.stride, xh0 = h0 * a.stride;
      mbar_wait(&empty[s], ph ^ 1u);
      uint8_t* sa = smem + s * L::STAGE_BYTES;
      uint8_t* sb = sa + L::A_BYTES;
      if (elect_one()) {
        if (rank == 0) mbar_expect_tx(&full[s], 2u * box_bytes * (uint32_t)(2 + nt * GPC));
        else mbar_arrive_leader(&full[s]);
#pragma unroll
        for (int g = 0; g < 2; ++g) tma_load_4d_2sm(&mapDy, &full[s], sa + g * L::GROUP_BYTES, co0 + 64 * g, w0, h0, img);
#pragma unroll
        for (int t = 0; t < MAXT; ++t) {
          if (t < nt) {
#pragma unroll
            for (int g = 0; g < GPC; ++g)
              tma_load_4d_2sm(&mapX, &full[s], sb + (t * GPC + g) * L::GROUP_BYTES, xc[t] + 64 * g, xw0 + xdw[t], xh0 + xdh[t], img);
          }
        }
      }
      __syncwarp();
      if (++tw_i == a.tiles_w) { tw_i = 0; if (++th_i == a.tiles_h) { th_i = 0; ++img; } }
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_mn(256, NW);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const int ksteps = a.kpix / 16;
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < kiters; ++it) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
        const uint64_t adesc = make_mnmajor_sw128_desc(sa, L::GROUP_BYTES);
        if (elect_one()) {
          for (int k = 0; k < ksteps; ++k)
            for (int t = 0; t < nt; ++t) {
              const uint64_t bdesc = make_mnmajor_sw128_desc(sa + L::A_BYTES + t * GPC * L::GROUP_BYTES, L::GROUP_BYTES);
              umma_bf16_2sm(tmem_u + (uint32_t)(t * NW), adesc + (uint64_t)(k * (2048 >> 4)), bdesc + (uint64_t)(k * (2048 >> 4)), idesc,
                            (it | k) != 0 ? 1u : 0u);
            }
          umma_commit_2sm(&empty[s]);
          if (it == kiters - 1) umma_commit_2sm(tmem_full);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (kiters > 0) {
    const int row = 32 * (warp & 3) + lane;
    const int co = co0 + row;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    uint8_t* tile = smem + (warp - 2) * EPI_TILE_BYTES;      // the stage ring is free once tmem_full has fired
    const size_t row_stride = (size_t)a.ntaps * a.Cin;
    for (int t = 0; t < nt; ++t) {
      int tap, ci0;
      wgrad2_vcol<NW>(a, v0 + t, &tap, &ci0);
      const uint32_t lane_addr = tmem_base + (uint32_t)(t * NW) + ((uint32_t)(32 * (warp & 3)) << 16);
      float* dst = a.dw + (size_t)blockIdx.y * a.dw_split_stride + ((size_t)co * a.ntaps + tap) * a.Cin + ci0;
      wgrad_store_tile<NW>(lane_addr, dst, row_stride, co, a.Cout, a.Cin - ci0, tile, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, tmem_cols);
  }
}

// which 2-SM weight-gradient kernel runs on this shape: 0 none (1-SM kernels), 1 the NW = 128 multi-accumulator tile, 2 the
// NW = 256 tile.  ETB_WGRAD2 is read per call so tests and benchmarks can toggle it.
static int wgrad2_mode(const EtbConvParams* cp, int32_t flags) {
  const char* e = getenv("ETB_WGRAD2");
  const int want = e ? atoi(e) : ETB_WGRAD2_DEFAULT;
  const int ntaps = cp->kh * cp->kw;
  if ((flags & 1) || cp->Cout < 256 || ntaps > 12) return 0;
  // default: the 256-wide 2-SM tile where it measured faster than the 1-SM 256x256 / 128x128 tiles (tools/conv_bench.py, batch
  // 32, profiles/r2_conv_bench_final*.json): the stride-2 3x3 layers (170 -> 135, 184 -> 133, 113 -> 94, 112 -> 86 us) and the
  // deepest pointwise layer (2048 -> 1024: 86 -> 75 us); everywhere else the 1-SM tiles are equal or better
  if (want < 0) return (cp->Cin % 256 == 0 && ((cp->stride == 2 && ntaps == 9) || (ntaps == 1 && cp->Cin >= 2048))) ? 2 : 0;
  if (want == 2 && cp->Cin % 256 == 0) return 2;
  if (want == 1 && cp->Cin >= 128 && cp->Cin % 64 == 0 && (ntaps == 1 || ntaps % 3 == 0)) return 1;
  return 0;
}

// tiling + split-K of the 2-SM kernels (shared by the workspace query and the launch)
static void wgrad2_plan(const EtbConvParams* cp, int mode, int* NT_, int* TW, int* TH, int* tiles_w, int* tiles_h, int* nimg, int* groups_,
                        int* co_pairs_, int* nvirt_, int* splitk) {
  constexpr int KP = 64;
  const int Ho = (cp->H + 2 * cp->pad - cp->kh) / cp->stride + 1;
  const int Wo = (cp->W + 2 * cp->pad - cp->kw) / cp->stride + 1;
  const bool flat = (cp->kh == 1 && cp->kw == 1 && cp->stride == 1 && cp->pad == 0);
  if (flat) {
    const long npix = (long)cp->N * cp->H * cp->W;
    *TW = KP; *TH = 1; *tiles_w = (int)((npix + KP - 1) / KP); *tiles_h = 1; *nimg = 1;
  } else {
    pick_tile16(Wo, Ho, KP, TW, TH);
    *tiles_w = (Wo + *TW - 1) / *TW; *tiles_h = (Ho + *TH - 1) / *TH; *nimg = cp->N;
  }
  const int ntaps = cp->kh * cp->kw;
  int nvirt, NT;
  double mma_us;                           // one virtual column of one 64-pixel K block: 4 MMAs
  if (mode == 2) {
    nvirt = ntaps * (cp->Cin / 256);
    NT = nvirt >= 2 ? 2 : 1;
    const char* e = getenv("ETB_WGRAD2_NT");          // tuning override (1 | 2)
    if (e && atoi(e) >= 1 && atoi(e) <= 2) NT = atoi(e);
    mma_us = 0.27;
  } else {
    const int ci_tiles = (cp->Cin + 127) / 128;
    nvirt = ci_tiles * ntaps;
    NT = (ntaps % 3 == 0) ? 3 : (ci_tiles >= 4 ? 4 : ci_tiles);
    mma_us = 0.135;
  }
  const int groups = (nvirt + NT - 1) / NT;
  const int co_pairs = (cp->Cout + 255) / 256;
  const int clusters = groups * co_pairs;
  const int total_kb = *nimg * *tiles_h * *tiles_w;
  // split-K: whole waves of (SMs / 2) clusters; cost per cluster = K blocks * NT columns * 4 MMAs + fixed (launch, TMEM
  // allocation, pipeline fill, the epilogue's partial-tile stores)
  const int slots = etb_num_sms() / 2;
  const double t_kb = mma_us * NT * (double)(*TW * *TH) / 64.0, t_fixed = 7.0 + (mode == 2 ? 1.6 : 1.2) * NT;
  int sk = 1;
  double best = 1e30;
  const int sk_max = total_kb < 512 ? total_kb : 512;
  for (int c = 1; c <= sk_max; ++c) {
    const long cl = (long)clusters * c;
    if (cl > 8L * slots) break;
    const long rounds = (cl + slots - 1) / slots;
    const int per = (total_kb + c - 1) / c;
    if ((long)(c - 1) * per >= total_kb) continue;          // would leave an empty slice
    const double cost = (double)rounds * (per * t_kb + t_fixed) + 0.004 * c;
    if (cost < best) { best = cost; sk = c; }
  }
  *NT_ = NT; *groups_ = groups; *co_pairs_ = co_pairs; *nvirt_ = nvirt; *splitk = sk;
}

template <int NW>
static int launch_wgrad2(const CUtensorMap& mDy, const CUtensorMap& mX, const Wgrad2Args& wa, int clusters, int splitk, cudaStream_t st) {
  using L = Wgrad2Smem<64, 4>;
  static bool attr_set = false;
  if (!attr_set) {
    ETB_CHECK_CUDA(cudaFuncSetAttribute(wgrad2_kernel<64, 4, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * clusters), (unsigned)splitk); cfg.blockDim = dim3(WGRAD_THREADS); cfg.dynamicSmemBytes = L::TOTAL; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = etb_pdl_enabled() ? 2 : 1;
  ETB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, wgrad2_kernel<64, 4, NW>, mDy, mX, wa));
  etb_count_launch();
  return ETB_OK;
}

static void wgrad_plan(const EtbConvParams* cp, int* BN_, int* KP_, int* TW, int* TH, int* tiles_w, int* tiles_h, int* nimg, int* out_tiles,
                       int* splitk, int* MT_ = nullptr) {
  const int Ho = (cp->H + 2 * cp->pad - cp->kh) / cp->stride + 1;
  const int Wo = (cp->W + 2 * cp->pad - cp->kw) / cp->stride + 1;
  int BN = cp->Cin >= 128 ? 128 : 64, KP = 128, MT = 1;
  static int cfg_mode = -1;               // tuning override: ETB_WGRAD_CFG = 0 | 1: 128x256 | 2: 256x128 | 3: 256x256
  if (cfg_mode < 0) { const char* e = getenv("ETB_WGRAD_CFG"); cfg_mode = e ? atoi(e) : 0; }
  if (cfg_mode == 1 && cp->Cin >= 256) { BN = 256; KP = 64; }
  if (cfg_mode == 2 && cp->Cout >= 256 && cp->Cin >= 128) { MT = 2; BN = 128; KP = 64; }
  if (cfg_mode == 3 && cp->Cout >= 256 && cp->Cin >= 256) { MT = 2; BN = 256; KP = 64; }
  // default: 256x256 tiles for the wide 3x3 layers (measured per shape, tools/conv_bench.py batch 32: 3x3 256->256 @40 116 -> 110 us,
  // 3x3 512->512 @20 116 -> 100 us; the pointwise layers and everything narrower stay on 128x128: profiles/r2_conv_bench_wgcfg*.json)
  if (cfg_mode == 0 && cp->kh * cp->kw >= 9 && cp->stride == 1 && cp->Cout >= 256 && cp->Cin >= 256) { MT = 2; BN = 256; KP = 64; }
  if (MT_) *MT_ = MT;
  const bool flat = (cp->kh == 1 && cp->kw == 1 && cp->stride == 1 && cp->pad == 0);
  if (flat) {
    const long npix = (long)cp->N * cp->H * cp->W;
    *TW = KP; *TH = 1; *tiles_w = (int)((npix + KP - 1) / KP); *tiles_h = 1; *nimg = 1;
  } else {
    pick_tile16(Wo, Ho, KP, TW, TH);
    *tiles_w = (Wo + *TW - 1) / *TW; *tiles_h = (Ho + *TH - 1) / *TH; *nimg = cp->N;
  }
  const int ntaps = cp->kh * cp->kw;
  *out_tiles = ((cp->Cout + 128 * MT - 1) / (128 * MT)) * ((cp->Cin + BN - 1) / BN) * ntaps;
  const int total_kb = *nimg * *tiles_h * *tiles_w;
  // split-K: minimise  rounds(out_tiles*sk) * (K blocks per CTA * t_kb + t_fixed)  -- whole waves of 148 CTAs matter more
  // than raw parallelism (measured: 2.2 waves cost 3 rounds).  t_kb ~ 512 MMA cycles per 128-pixel block, t_fixed ~ launch +
  // TMEM alloc + pipeline fill + epilogue of one CTA.
  const int sms = etb_num_sms();
  const double t_kb = 0.27 * (double)(*TW * *TH) / 128.0, t_fixed = 6.0;
  int sk = 1;
  double best = 1e30;
  const int sk_max = total_kb < 512 ? total_kb : 512;
  for (int c = 1; c <= sk_max; ++c) {
    const long ctas = (long)*out_tiles * c;
    if (ctas > 8L * sms) break;
    const long rounds = (ctas + sms - 1) / sms;
    const int per = (total_kb + c - 1) / c;
    if ((long)(c - 1) * per >= total_kb) continue;          // would leave an empty slice
    const double cost = (double)rounds * (per * t_kb + t_fixed) + 0.002 * c;   // tiny bias towards fewer partials to reduce
    if (cost < best) { best = cost; sk = c; }
  }
  *splitk = sk; *BN_ = BN; *KP_ = KP;
}

extern "C" size_t etb_conv_wgrad_workspace_bytes(const EtbConvParams* cp) {
  if (!cp || cp->Cin <= 0 || cp->Cout <= 0) return 0;
  int BN, KP, TW, TH, tw, th, ni, ot, sk;
  wgrad_plan(cp, &BN, &KP, &TW, &TH, &tw, &th, &ni, &ot, &sk);
  const int mode2 = wgrad2_mode(cp, 0);
  if (mode2) {                           // the caller may run either kernel on this shape: size for the larger split
    int NT, g, cpz, nv, sk2;
    wgrad2_plan(cp, mode2, &NT, &TW, &TH, &tw, &th, &ni, &g, &cpz, &nv, &sk2);
    if (sk2 > sk) sk = sk2;
  }
  return (size_t)sk * cp->Cout * cp->kh * cp->kw * cp->Cin * sizeof(float);
}

// x [N,H,W,*] bf16 (cp->x_cstride), dy [N,Ho,Wo,*] bf16 (channel stride cp->y_cstride) -> dw [Cout,Cin,kh,kw] fp32 (the
// nn.Parameter layout).  flags bit0: stem (cp describes the K=128 pointwise GEMM over the im2col buffer, dw is [Cout,3,6,6]);
// bit1: dw += result (dw may be the gradient-arena slice).  workspace: etb_conv_wgrad_workspace_bytes(cp).
extern "C" int etb_conv_wgrad(const void* x_bf16, const void* dy_bf16, float* dw_f32, const EtbConvParams* cp, int32_t flags, void* workspace,
                              size_t workspace_bytes, void* stream) {
  ETB_CHECK_ARG(x_bf16 && dy_bf16 && dw_f32 && cp && workspace);
  ETB_CHECK_ARG(cp->N > 0 && cp->H > 0 && cp->W > 0 && cp->Cin > 0 && cp->Cout > 0 && cp->Cin % 8 == 0 && (cp->Cin % 64 == 0 || cp->Cin < 64));
  ETB_CHECK_ARG(cp->kh * cp->kw <= 12 && (cp->stride == 1 || cp->stride == 2));
  ETB_CHECK_ARG(cp->x_cstride % 8 == 0 && cp->y_cstride % 8 == 0 && (((uintptr_t)x_bf16) & 15) == 0 && (((uintptr_t)dy_bf16) & 15) == 0);
  ETB_CHECK_ARG((((uintptr_t)workspace) & 15) == 0);
  PFN_tmapEncodeTiled enc = get_encode();
  if (!enc) {
    etb_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return ETB_ERR_CUDA;
  }
  const int Ho = (cp->H + 2 * cp->pad - cp->kh) / cp->stride + 1;
  const int Wo = (cp->W + 2 * cp->pad - cp->kw) / cp->stride + 1;
  const int mode2 = wgrad2_mode(cp, flags);
  const bool use2 = mode2 != 0;
  WgradArgs wa;
  memset(&wa, 0, sizeof(wa));
  int BN, KP, out_tiles, splitk, MT;
  wgrad_plan(cp, &BN, &KP, &wa.TW, &wa.TH, &wa.tiles_w, &wa.tiles_h, &wa.nimg, &out_tiles, &splitk, &MT);
  Wgrad2Args w2;
  memset(&w2, 0, sizeof(w2));
  int groups2 = 0, co_pairs2 = 0;
  if (use2) {
    wgrad2_plan(cp, mode2, &w2.NT, &wa.TW, &wa.TH, &wa.tiles_w, &wa.tiles_h, &wa.nimg, &groups2, &co_pairs2, &w2.nvirt, &splitk);
    KP = 64;
  }
  wa.kpix = wa.TW * wa.TH;
  wa.ntaps = cp->kh * cp->kw;
  for (int kh = 0; kh < cp->kh; ++kh)
    for (int kw = 0; kw < cp->kw; ++kw) {
      wa.tap_dh[kh * cp->kw + kw] = (signed char)(kh - cp->pad);
      wa.tap_dw[kh * cp->kw + kw] = (signed char)(kw - cp->pad);
    }
  wa.stride = cp->stride;
  wa.Cout = cp->Cout; wa.Cin = cp->Cin;
  const size_t dw_elems = (size_t)cp->Cout * wa.ntaps * cp->Cin;
  if (workspace_bytes < (size_t)splitk * dw_elems * sizeof(float)) {
    etb_set_error("etb_conv_wgrad: workspace too small (%zu < %zu)", workspace_bytes, (size_t)splitk * dw_elems * sizeof(float));
    return ETB_ERR_NOMEM;
  }
  wa.dw = (float*)workspace;
  wa.dw_split_stride = (long)dw_elems;
  const bool flat = (cp->kh == 1 && cp->kw == 1 && cp->stride == 1 && cp->pad == 0);
  cuuint64_t ddim[4], dstr[3], xdim[4], xstr[3];
  cuuint32_t dbox[4], xbox[4], one[4] = {1, 1, 1, 1}, xes[4];
  if (flat) {
    const long npix = (long)cp->N * cp->H * cp->W;
    ETB_CHECK_ARG(npix < (1l << 31));
    ddim[0] = cp->Cout; ddim[1] = (cuuint64_t)npix; ddim[2] = 1; ddim[3] = 1;
    dstr[0] = (cuuint64_t)cp->y_cstride * 2; dstr[1] = dstr[0] * (cuuint64_t)npix; dstr[2] = dstr[1];
    xdim[0] = cp->Cin; xdim[1] = (cuuint64_t)npix; xdim[2] = 1; xdim[3] = 1;
    xstr[0] = (cuuint64_t)cp->x_cstride * 2; xstr[1] = xstr[0] * (cuuint64_t)npix; xstr[2] = xstr[1];
    dbox[0] = 64; dbox[1] = KP; dbox[2] = 1; dbox[3] = 1;
    xbox[0] = 64; xbox[1] = KP; xbox[2] = 1; xbox[3] = 1;
    xes[0] = xes[1] = xes[2] = xes[3] = 1;
  } else {
    ddim[0] = cp->Cout; ddim[1] = Wo; ddim[2] = Ho; ddim[3] = cp->N;
    dstr[0] = (cuuint64_t)cp->y_cstride * 2; dstr[1] = dstr[0] * Wo; dstr[2] = dstr[1] * Ho;
    xdim[0] = cp->Cin; xdim[1] = cp->W; xdim[2] = cp->H; xdim[3] = cp->N;
    xstr[0] = (cuuint64_t)cp->x_cstride * 2; xstr[1] = xstr[0] * cp->W; xstr[2] = xstr[1] * cp->H;
    dbox[0] = 64; dbox[1] = wa.TW; dbox[2] = wa.TH; dbox[3] = 1;
    xbox[0] = 64; xbox[1] = wa.TW * cp->stride; xbox[2] = wa.TH * cp->stride; xbox[3] = 1;
    xes[0] = 1; xes[1] = cp->stride; xes[2] = cp->stride; xes[3] = 1;
    ETB_CHECK_ARG(xbox[1] <= 256 && xbox[2] <= 256);
  }
  CUtensorMap mDy, mX;
  CUresult r = enc(&mDy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dy_bf16), ddim, dstr, dbox, one, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { etb_set_error("cuTensorMapEncodeTiled(dy) failed: %d", (int)r); return ETB_ERR_CUDA; }
  r = enc(&mX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x_bf16), xdim, xstr, xbox, xes, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { etb_set_error("cuTensorMapEncodeTiled(x) failed: %d", (int)r); return ETB_ERR_CUDA; }
  wa.flat = flat ? 1 : 0;
  wa.co_tiles = (cp->Cout + 128 * MT - 1) / (128 * MT);
  wa.ci_tiles = (cp->Cin + BN - 1) / BN;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)out_tiles, (unsigned)splitk);
  int rc;
  if (use2) {
    w2.ntaps = wa.ntaps;
    memcpy(w2.tap_dh, wa.tap_dh, sizeof(w2.tap_dh));
    memcpy(w2.tap_dw, wa.tap_dw, sizeof(w2.tap_dw));
    w2.stride = wa.stride; w2.TW = wa.TW; w2.TH = wa.TH; w2.tiles_w = wa.tiles_w; w2.tiles_h = wa.tiles_h; w2.nimg = wa.nimg;
    w2.kpix = wa.kpix;
    w2.groups = groups2;
    w2.cit = cp->Cin / 256;
    w2.Cout = cp->Cout; w2.Cin = cp->Cin;
    w2.dw = wa.dw; w2.dw_split_stride = wa.dw_split_stride;
    rc = mode2 == 2 ? launch_wgrad2<256>(mDy, mX, w2, groups2 * co_pairs2, splitk, st)
                    : launch_wgrad2<128>(mDy, mX, w2, groups2 * co_pairs2, splitk, st);
    if (rc != ETB_OK) return rc;
  }
  const bool cluster = (MT == 1 && BN == 128) && (wa.co_tiles % 2 == 0) && (wa.ci_tiles % 2 == 0) && getenv("ETB_WGRAD_CLUSTER");
  if (use2) rc = ETB_OK;
  else if (cluster) rc = launch_wgrad<1, 128, 128, 3, 1>(mDy, mX, wa, grid, st);
  else if (MT == 2 && BN == 256) rc = launch_wgrad<2, 256, 64, 3, 0>(mDy, mX, wa, grid, st);
  else if (MT == 2) rc = launch_wgrad<2, 128, 64, 4, 0>(mDy, mX, wa, grid, st);
  else if (BN == 256) rc = launch_wgrad<1, 256, 64, 4, 0>(mDy, mX, wa, grid, st);
  else if (BN == 128) rc = launch_wgrad<1, 128, 128, 3, 0>(mDy, mX, wa, grid, st);
  else rc = launch_wgrad<1, 64, 128, 4, 0>(mDy, mX, wa, grid, st);
  if (rc != ETB_OK) return rc;
  // rows co >= Cout of the last tile are never written: the reduce only reads [Cout] rows
  const long n4 = (long)dw_elems / 4;
  long blocks = (n4 + 31) / 32;
  const long cap = (long)etb_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (wa.ntaps > 1 && !(flags & 1))
    etb_launch(wgrad_reduce_taps_kernel, dim3((unsigned)(cp->Cout * (cp->Cin < 64 ? 1 : cp->Cin >> 6))), dim3(256), 0, st, (const float*)workspace, (long)dw_elems, splitk, dw_f32, cp->Cin, wa.ntaps,
                                                                                  flags);
  else if (splitk >= 32)
    etb_launch(wgrad_reduce_kernel<16>, dim3((unsigned)blocks), dim3(dim3(32, 16)), 0, st, (const float*)workspace, (long)dw_elems, splitk, dw_f32, cp->Cout, cp->Cin, wa.ntaps, flags);
  else if (splitk >= 6)
    etb_launch(wgrad_reduce_kernel<4>, dim3((unsigned)blocks), dim3(dim3(32, 4)), 0, st, (const float*)workspace, (long)dw_elems, splitk, dw_f32, cp->Cout, cp->Cin, wa.ntaps, flags);
  else
    etb_launch(wgrad_reduce_kernel<1>, dim3((unsigned)blocks), dim3(dim3(32, 1)), 0, st, (const float*)workspace, (long)dw_elems, splitk, dw_f32, cp->Cout, cp->Cin, wa.ntaps, flags);
  ETB_CHECK_LAUNCH();
  return ETB_OK;
}
