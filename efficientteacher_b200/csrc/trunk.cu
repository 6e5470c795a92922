// trunk.cu -- layout / pooling / folding helpers around the tcgen05 convolutions (K3, K4, K18).
//   stem im2col + /255      reference trainer/ssod_trainer.py:694-696, models/backbone/yolov5_backbone.py:56
//   SPPF max pools + concat reference models/backbone/common.py:702-708
//   nearest 2x upsample     reference models/neck/yolov5_neck.py:92,97 (+ Concat common.py:796-797, free by slicing)
//   eval BN folding         reference utils/torch_utils.py:199-219 (fuse_conv_and_bn algebra), bn eps 1e-3
// All elementwise / gather kernels: HBM-bound, 16 B vector accesses along the channel (fastest) dimension.
#include "common.cuh"

static inline unsigned grid_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)etb_num_sms() * 32;
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

// ---- stem im2col (6x6 s2 p2, 3 channels -> K = 128 slots, 108 used) ----
// K order k = (c*6 + kh)*6 + kw, i.e. exactly the [ci][kh][kw] order of the OIHW weight row, so the weight pack is a copy.
// One block = 64 consecutive output pixels of one output row: the 18 (c,kh) input row segments (132 floats each) are
// staged in shared memory with coalesced loads (zero-filled outside the image = the conv padding), then every thread
// assembles 16 B chunks [pixel][8 k] so a warp writes 512 contiguous bytes.  HBM-bound: 12 B/pixel-channel read
// (L2 serves the 3x row overlap), 256 B/pixel written.
#define STEM_TP 64
#define STEM_PITCH 133
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H, int W, float mul) {
  ETB_PDL_PROLOGUE();
  __shared__ float sm[18 * STEM_PITCH];
  const int Ho = H / 2, Wo = W / 2;
  const int tiles_w = (Wo + STEM_TP - 1) / STEM_TP;
  const int tw = blockIdx.x % tiles_w;
  const int oh = (blockIdx.x / tiles_w) % Ho;
  const int n = blockIdx.x / (tiles_w * Ho);
  const int ow0 = tw * STEM_TP;
  const int iw0 = 2 * ow0 - 2, ih0 = 2 * oh - 2;
  for (int i = threadIdx.x; i < 18 * 132; i += 256) {
    const int row = i / 132, col = i - row * 132;
    const int c = row / 6, kh = row - c * 6;
    const int ih = ih0 + kh, iw = iw0 + col;
    float v = 0.f;
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = __fmul_rn(__ldg(x + (((int64_t)n * 3 + c) * H + ih) * W + iw), mul);
    sm[row * STEM_PITCH + col] = v;
  }
  __syncthreads();
  const int g = threadIdx.x & 15;          // the 8-element K group is fixed per thread
  int off[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = g * 8 + j;
    off[j] = k < 108 ? (k / 6) * STEM_PITCH + (k % 6) : -1;
  }
  uint4* yo = reinterpret_cast<uint4*>(y) + (((int64_t)n * Ho + oh) * Wo + ow0) * 16;
#pragma unroll
  for (int q = threadIdx.x; q < STEM_TP * 16; q += 256) {
    const int pp = q >> 4;
    if (ow0 + pp >= Wo) break;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = off[j] >= 0 ? sm[off[j] + 2 * pp] : 0.f;
    uint4 ov;
    __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    yo[q] = ov;
  }
}

extern "C" int etb_stem_im2col(const float* x, void* y_bf16, int32_t N, int32_t H, int32_t W, float mul, void* stream) {
  ETB_CHECK_ARG(x && y_bf16 && N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0);
  const int64_t blocks = (int64_t)N * (H / 2) * ((W / 2 + STEM_TP - 1) / STEM_TP);
  ETB_CHECK_ARG(blocks < (1ll << 31));
  etb_launch(stem_im2col_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, x, (__nv_bfloat16*)y_bf16, N, H, W, mul);
  ETB_CHECK_LAUNCH();
  return ETB_OK;
}

// ---- NCHW fp32 <-> NHWC bf16 (simple gather; used at the edges of the trunk and by the tests) ----
__global__ void __launch_bounds__(256) nchw2nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int C, int H, int W,
                                                        int cs, int co, float mul) {
  ETB_PDL_PROLOGUE();
  const int64_t total = (int64_t)N * H * W * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const int64_t pix = e / C;
    const int w = (int)(pix % W), h = (int)((pix / W) % H), n = (int)(pix / ((int64_t)W * H));
    y[pix * cs + co + c] = __float2bfloat16(__fmul_rn(x[(((int64_t)n * C + c) * H + h) * W + w], mul));
  }
}
__global__ void __launch_bounds__(256) nhwc2nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int N, int C, int H, int W, int cs, int co) {
  ETB_PDL_PROLOGUE();
  const int64_t total = (int64_t)N * H * W * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(e % W), h = (int)((e / W) % H), c = (int)((e / ((int64_t)W * H)) % C), n = (int)(e / ((int64_t)W * H * C));
    y[e] = __bfloat162float(x[(((int64_t)n * H + h) * W + w) * cs + co + c]);
  }
}
extern "C" int etb_nchw_f32_to_nhwc_bf16(const float* x, void* y, int32_t N, int32_t C, int32_t H, int32_t W, int32_t y_cstride,
                                         int32_t y_coffset, float mul, void* stream) {
  ETB_CHECK_ARG(x && y && N > 0 && C > 0 && H > 0 && W > 0 && y_cstride >= y_coffset + C);
  etb_launch(nchw2nhwc_kernel, dim3(grid_for((int64_t)N * C * H * W, 256)), dim3(256), 0, (cudaStream_t)stream, x, (__nv_bfloat16*)y, N, C, H, W, y_cstride, y_coffset, mul);
  ETB_CHECK_LAUNCH();
  return ETB_OK;
}
extern "C" int etb_nhwc_bf16_to_nchw_f32(const void* x, float* y, int32_t N, int32_t C, int32_t H, int32_t W, int32_t x_cstride,
                                         int32_t x_coffset, void* stream) {
  ETB_CHECK_ARG(x && y && N > 0 && C > 0 && H > 0 && W > 0 && x_cstride >= x_coffset + C);
  etb_launch(nhwc2nchw_kernel, dim3(grid_for((int64_t)N * C * H * W, 256)), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, y, N, C, H, W, x_cstride, x_coffset);
  ETB_CHECK_LAUNCH();
  return ETB_OK;
}

// ---- SPPF: three cascaded 5x5 s1 p2 max pools == 5x5, 9x9, 13x13 windows of x (max is idempotent/associative) ----
__device__ __forceinline__ void bf8_max(uint4& acc, const uint4& v) {
  __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&acc);
  const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] = __hmax2(a[j], b[j]);
}
__global__ void __launch_bounds__(256) sppf_pool_kernel(__nv_bfloat16* __restrict__ buf, int N, int H, int W, int C, int cs) {
  ETB_PDL_PROLOGUE();
  const int cg = C / 8;
  const int64_t total = (int64_t)N * H * W * cg;
  const uint32_t ninf2 = 0xFF80FF80u;  // bf16 -inf pair (max-pool padding value)
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(e % cg);
    const int64_t pix = e / cg;
    const int w = (int)(pix % W), h = (int)((pix / W) % H), n = (int)(pix / ((int64_t)W * H));
    uint4 m5 = make_uint4(ninf2, ninf2, ninf2, ninf2), m9 = m5, m13 = m5;
    for (int dy = -6; dy <= 6; ++dy) {
      const int ih = h + dy;
      if (ih < 0 || ih >= H) continue;
      for (int dx = -6; dx <= 6; ++dx) {
        const int iw = w + dx;
        if (iw < 0 || iw >= W) continue;
        const uint4 v = *reinterpret_cast<const uint4*>(buf + (((int64_t)n * H + ih) * W + iw) * cs + g * 8);
        bf8_max(m13, v);
        if (dy >= -4 && dy <= 4 && dx >= -4 && dx <= 4) bf8_max(m9, v);
        if (dy >= -2 && dy <= 2 && dx >= -2 && dx <= 2) bf8_max(m5, v);
      }
    }
    __nv_bfloat16* o = buf + pix * cs + g * 8;
    *reinterpret_cast<uint4*>(o + C) = m5;
    *reinterpret_cast<uint4*>(o + 2 * C) = m9;
    *reinterpret_cast<uint4*>(o + 3 * C) = m13;
  }
}
extern "C" int etb_sppf_pool(void* buf_bf16, int32_t N, int32_t H, int32_t W, int32_t C, int32_t cstride, void* stream) {
  ETB_CHECK_ARG(buf_bf16 && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && cstride >= 4 * C && cstride % 8 == 0);
  etb_launch(sppf_pool_kernel, dim3(grid_for((int64_t)N * H * W * (C / 8), 256)), dim3(256), 0, (cudaStream_t)stream, (__nv_bfloat16*)buf_bf16, N, H, W, C, cstride);
  ETB_CHECK_LAUNCH();
  return ETB_OK;
}

// ---- nearest 2x upsample into a channel slice ----
__global__ void __launch_bounds__(256) upsample2x_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H, int W, int C,
                                                         int xcs, int xco, int ycs, int yco) {
  ETB_PDL_PROLOGUE();
  const int cg = C / 8;
  const int Ho = 2 * H, Wo = 2 * W;
  const int64_t total = (int64_t)N * Ho * Wo * cg;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(e % cg);
    const int64_t pix = e / cg;
    const int w = (int)(pix % Wo), h = (int)((pix / Wo) % Ho), n = (int)(pix / ((int64_t)Wo * Ho));
    const uint4 v = *reinterpret_cast<const uint4*>(x + (((int64_t)n * H + (h >> 1)) * W + (w >> 1)) * xcs + xco + g * 8);
    *reinterpret_cast<uint4*>(y + pix * ycs + yco + g * 8) = v;
  }
}
extern "C" int etb_upsample2x_nhwc(const void* x_bf16, void* y_bf16, int32_t N, int32_t H, int32_t W, int32_t C, int32_t x_cstride,
                                   int32_t x_coffset, int32_t y_cstride, int32_t y_coffset, void* stream) {
  ETB_CHECK_ARG(x_bf16 && y_bf16 && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0);
  ETB_CHECK_ARG(x_cstride % 8 == 0 && x_coffset % 8 == 0 && y_cstride % 8 == 0 && y_coffset % 8 == 0);
  etb_launch(upsample2x_kernel, dim3(grid_for((int64_t)N * 4 * H * W * (C / 8), 256)), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x_bf16, (__nv_bfloat16*)y_bf16, N, H, W, C, x_cstride, x_coffset, y_cstride, y_coffset);
  ETB_CHECK_LAUNCH();
  return ETB_OK;
}

// ---- BN folding + weight packing ----
__global__ void fold_bn_kernel(const float* g, const float* b, const float* m, const float* v, float eps, float* scale, float* bias, int C) {
  ETB_PDL_PROLOGUE();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float s = g[c] / sqrtf(v[c] + eps);
  scale[c] = s;
  bias[c] = b[c] - m[c] * s;
}
extern "C" int etb_fold_bn(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale,
                           float* bias, int32_t C, void* stream) {
  ETB_CHECK_ARG(gamma && beta && mean && var && scale && bias && C > 0);
  etb_launch(fold_bn_kernel, dim3((C + 127) / 128), dim3(128), 0, (cudaStream_t)stream, gamma, beta, mean, var, eps, scale, bias, C);
  ETB_CHECK_LAUNCH();
  return ETB_OK;
}

__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ o, int Cout, int Cin, int kh, int kw, int Cp) {
  ETB_PDL_PROLOGUE();
  const int64_t total = (int64_t)Cout * kh * kw * Cp;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % Cp);
    const int64_t t = e / Cp;
    const int x = (int)(t % kw), y = (int)((t / kw) % kh), oc = (int)(t / ((int64_t)kw * kh));
    o[e] = __float2bfloat16(c < Cin ? w[(((int64_t)oc * Cin + c) * kh + y) * kw + x] : 0.f);
  }
}
extern "C" int etb_pack_weight(const float* w_oihw, void* w_bf16, int32_t Cout, int32_t Cin, int32_t kh, int32_t kw, int32_t Cin_pad, void* stream) {
  ETB_CHECK_ARG(w_oihw && w_bf16 && Cout > 0 && Cin > 0 && kh > 0 && kw > 0 && Cin_pad >= Cin);
  etb_launch(pack_weight_kernel, dim3(grid_for((int64_t)Cout * kh * kw * Cin_pad, 256)), dim3(256), 0, (cudaStream_t)stream, w_oihw, (__nv_bfloat16*)w_bf16, Cout, Cin, kh, kw, Cin_pad);
  ETB_CHECK_LAUNCH();
  return ETB_OK;
}

__global__ void pack_stem_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ o, int Cout) {
  ETB_PDL_PROLOGUE();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Cout * 128) return;
  const int k = e & 127, oc = e >> 7;
  float v = 0.f;
  if (k < 108) v = w[oc * 108 + k];          // K order (c,kh,kw) == the OIHW row
  o[e] = __float2bfloat16(v);
}
extern "C" int etb_pack_stem_weight(const float* w_oihw, void* w_bf16, int32_t Cout, void* stream) {
  ETB_CHECK_ARG(w_oihw && w_bf16 && Cout > 0);
  etb_launch(pack_stem_weight_kernel, dim3((Cout * 128 + 255) / 256), dim3(256), 0, (cudaStream_t)stream, w_oihw, (__nv_bfloat16*)w_bf16, Cout);
  ETB_CHECK_LAUNCH();
  return ETB_OK;
}

// ---- multi-tensor weight packing / BN folding: ONE launch for all ~104 convs of the trunk (replaces ~440 tiny launches/step) ----
// mode 0: fwd  [Cout][kh][kw][Cin or out_ld]   <- w[co][ci][kh][kw]        (dst index e: ci fastest; out_ld > Cin pads every tap)
// mode 1: dgrad class  [Cin][ntaps][out_ld>=Cout] <- w[co][ci][kh_t][kw_t]   (dst: co fastest; row pitch out_ld per tap)
// mode 2: stem [Cout][128] in the etb_stem_im2col K order
// mode 3: mode 1 negated (dgrad operand of a conv that sits behind a GradReverse)
__global__ void __launch_bounds__(256) pack_multi_kernel(const EtbPackDesc* __restrict__ descs, const int2* __restrict__ chunks) {
  ETB_PDL_PROLOGUE();
  const int2 ch = chunks[blockIdx.x];
  const EtbPackDesc d = descs[ch.x];
  const float* __restrict__ w = d.w;
  __nv_bfloat16* __restrict__ o = (__nv_bfloat16*)d.out;
  const int64_t base = (int64_t)ch.y * ETB_PACK_CHUNK;
  const int kk = d.k * d.k;
#pragma unroll 4
  for (int i = threadIdx.x; i < ETB_PACK_CHUNK; i += 256) {
    const int64_t e = base + i;
    if (e >= d.elems) break;
    float v;
    int64_t dst = e;
    if (d.mode == 0) {
      const int ci = (int)(e % d.Cin);
      const int64_t t2 = e / d.Cin;
      const int t = (int)(t2 % kk), co = (int)(t2 / kk);
      v = w[((int64_t)co * d.Cin + ci) * kk + t];
      if (d.out_ld > d.Cin) dst = ((int64_t)co * kk + t) * d.out_ld + ci;   // every tap padded to out_ld = ceil64(Cin) (pad stays zero)
    } else if (d.mode == 1 || d.mode == 3) {
      const int co = (int)(e % d.Cout);
      const int64_t t2 = e / d.Cout;
      const int t = (int)(t2 % d.ntaps), ci = (int)(t2 / d.ntaps);
      v = w[(((int64_t)co * d.Cin + ci) * d.k + d.kh[t]) * d.k + d.kw[t]];
      if (d.mode == 3) v = -v;            // GradReverse in front of the conv: dx = -(W^T dy)
      dst = ((int64_t)ci * d.ntaps + t) * d.out_ld + co;
    } else {
      const int k = (int)(e & 127), oc = (int)(e >> 7);
      v = 0.f;
      if (k < 108) v = w[oc * 108 + k];
    }
    o[dst] = __float2bfloat16(v);
  }
}

extern "C" int etb_pack_multi(const EtbPackDesc* descs_dev, const void* chunks_dev, int32_t n_chunks, void* stream) {
  ETB_CHECK_ARG(descs_dev && chunks_dev && n_chunks >= 0);
  if (n_chunks == 0) return ETB_OK;
  etb_launch(pack_multi_kernel, dim3(n_chunks), dim3(256), 0, (cudaStream_t)stream, descs_dev, (const int2*)chunks_dev);
  ETB_CHECK_LAUNCH();
  return ETB_OK;
}

__global__ void __launch_bounds__(256) fold_multi_kernel(const EtbFoldDesc* __restrict__ descs) {
  ETB_PDL_PROLOGUE();
  const EtbFoldDesc d = descs[blockIdx.x];
  for (int c = threadIdx.x; c < d.C; c += 256) {
    const float s = d.gamma[c] / sqrtf(d.var[c] + d.eps);
    d.scale[c] = s;
    d.bias[c] = d.beta[c] - d.mean[c] * s;
  }
}
extern "C" int etb_fold_bn_multi(const EtbFoldDesc* descs_dev, int32_t n, void* stream) {
  ETB_CHECK_ARG(descs_dev && n >= 0);
  if (n == 0) return ETB_OK;
  etb_launch(fold_multi_kernel, dim3(n), dim3(256), 0, (cudaStream_t)stream, descs_dev);
  ETB_CHECK_LAUNCH();
  return ETB_OK;
}
