// f3: the tiny VAE around the UNet step -- `stream.vae` = diffusers AutoencoderTiny ("madebyollin/taesd",
// live2diff/utils/wrapper.py:468-470) as used by encode_image / encode_depth / decode_image
// (live2diff/pipeline_stream_animation_depth.py:517-542, 565-571) -- plus the uint8 <-> [-1,1] pre/post-processing of
// __call__ (:630; live2diff/image_utils.py:9-30).
//
// Architecture restated from diffusers 0.25.0 models/vae.py (EncoderTiny / DecoderTiny / AutoencoderTinyBlock); state-dict
// keys are diffusers' own (encoder.layers.N..., decoder.layers.N...).  Every 3x3 convolution is a tcgen05 GEMM:
//   * 64 -> 64 stride 1 (the bulk: 3 per residual block): implicit GEMM over the channels-last activation through 4-D TMA
//     halo boxes (gemm_tcgen05.cu), ReLU / residual + ReLU fused into the epilogue
//   * 3 -> 64 / 4 -> 64 stems: a [M,64] im2col straight from the NCHW image / latent with the input transform fused
//     ((x+1)/2 for the encoder, tanh(x/3)*3 for the decoder), then a K = 64 GEMM
//   * stride-2 convs: im2col (stride 2) + GEMM;  Upsample(2, nearest) + conv: one replicate pass, then the implicit GEMM
//   * 64 -> 4 / 64 -> 3 heads: implicit GEMM with the output channels padded to 8, then an NHWC8 -> NCHW gather with the
//     output transform fused (2x-1, clip)
// Activations are channels-last fp16 [N*H*W, 64] end to end.
#include <memory>
#include <string>
#include <vector>

#include "ops.cuh"

namespace l2d {
namespace {

constexpr int TCH = 64;

// conv weight [Cout,Cin,3,3] -> [Cout_pad, Kpad], k = (ky*3+kx)*Cin + cin; rows >= Cout / cols >= 9*Cin are zero
__global__ void tv_repack_kernel(const __half* __restrict__ w, __half* __restrict__ out, int cout, int cin, int cout_pad, int kpad) {
  const size_t total = (size_t)cout_pad * kpad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % kpad), co = (int)(i / kpad);
    __half v = __float2half(0.f);
    if (co < cout && k < 9 * cin) {
      const int tap = k / cin, ci = k - tap * cin;
      v = w[((size_t)co * cin + ci) * 9 + tap];
    }
    out[i] = v;
  }
}
__global__ void tv_pad_vec_kernel(const __half* __restrict__ in, __half* __restrict__ out, int n, int n_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) out[i] = (in && i < n) ? in[i] : __float2half(0.f);
}

// NCHW [N,nch,h,w] (nch = 3 image / 4 latent) -> im2col [N*h*w, 64]: col = tap*nch + c for col < 9*nch, else 0.
// mode 1: x <- (x + 1) / 2 (EncoderTiny.forward);  mode 2: x <- tanh(x / 3) * 3 (DecoderTiny.forward "Clamp").
// The transform applies to in-range pixels only: the conv pads the TRANSFORMED tensor with zeros.
__global__ void __launch_bounds__(256) tv_im2col_nchw_kernel(const __half* __restrict__ x, __half* __restrict__ y, int n_img, int nch,
                                                             int h, int w, int mode) {
  const size_t total = (size_t)n_img * h * w * 8;   // 8 chunks of 8 columns per row
  const int ncol = 9 * nch;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int chunk = (int)(i & 7);
    const size_t pix = i >> 3;
    const int n = (int)(pix / ((size_t)h * w));
    const int rem = (int)(pix % ((size_t)h * w));
    const int oy = rem / w, ox = rem % w;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = chunk * 8 + e;
      float t = 0.f;
      if (col < ncol) {
        const int tap = col / nch, c = col - tap * nch;
        const int iy = oy + tap / 3 - 1, ix = ox + tap % 3 - 1;
        if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
          t = __half2float(x[(((size_t)n * nch + c) * h + iy) * w + ix]);
          // the reference evaluates these on fp16 tensors: one rounding per elementwise op
          if (mode == 1) t = __half2float(__float2half_rn(__half2float(__float2half_rn(t + 1.f)) * 0.5f));
          else if (mode == 2) t = __half2float(__float2half_rn(__half2float(__float2half_rn(tanhf(__half2float(__float2half_rn(t / 3.f))))) * 3.f));
        }
      }
      v[e] = t;
    }
    *reinterpret_cast<uint4*>(y + pix * 64 + chunk * 8) = pack8(v);
  }
}

// nn.Upsample(scale_factor=2, mode="nearest") on a channels-last tensor [N,h,w,C] -> [N,2h,2w,C]
__global__ void __launch_bounds__(256) tv_upsample2x_kernel(const __half* __restrict__ x, __half* __restrict__ y, int n_img, int h,
                                                            int w, int C) {
  const int nch = C >> 3;
  const size_t total = (size_t)n_img * 4 * h * w * nch;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % nch);
    const size_t opix = i / nch;
    const int ox = (int)(opix % (2 * w));
    const int oy = (int)((opix / (2 * w)) % (2 * h));
    const int n = (int)(opix / ((size_t)4 * h * w));
    *reinterpret_cast<uint4*>(y + opix * C + ch * 8) =
        ldg_act(x + (((size_t)n * h + (oy >> 1)) * w + (ox >> 1)) * C + ch * 8);
  }
}

// [M, 8] channels-last (first c_out channels valid) -> NCHW [N,c_out,hw]; mode 0: raw, 1: 2x-1 (DecoderTiny.forward), 2: clip(2x-1,-1,1)
__global__ void __launch_bounds__(256) tv_head_to_nchw_kernel(const __half* __restrict__ x, __half* __restrict__ y, int n_img, int hw,
                                                              int c_out, int mode) {
  const size_t total = (size_t)n_img * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / hw), p = (int)(i % hw);
    float v[8];
    unpack8(ldg_act(x + i * 8), v);
    for (int c = 0; c < c_out; ++c) {
      float t = v[c];
      if (mode >= 1) t = __half2float(__float2half_rn(__half2float(__float2half_rn(t * 2.f)) - 1.f));
      if (mode == 2) t = fminf(fmaxf(t, -1.f), 1.f);
      y[((size_t)n * c_out + c) * hw + p] = __float2half_rn(t);
    }
  }
}

// uint8 [N,H,W,3] -> fp16 NCHW [N,3,H,W] in [-1,1]: VaeImageProcessor.preprocess (x/255 -> 2x-1), computed in fp32
__global__ void __launch_bounds__(256) tv_u8_to_f16_kernel(const uint8_t* __restrict__ x, __half* __restrict__ y, int n_img, int hw) {
  const size_t total = (size_t)n_img * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / hw), p = (int)(i % hw);
#pragma unroll
    for (int c = 0; c < 3; ++c)
      y[((size_t)n * 3 + c) * hw + p] = __float2half_rn((float)x[i * 3 + c] / 255.0f * 2.f - 1.f);
  }
}
// fp16 NCHW [N,3,H,W] -> uint8 [N,H,W,3]: (x/2 + 0.5).clamp(0,1) * 255, round half to even like numpy (image_utils.py:9-30)
__global__ void __launch_bounds__(256) tv_f16_to_u8_kernel(const __half* __restrict__ x, uint8_t* __restrict__ y, int n_img, int hw) {
  const size_t total = (size_t)n_img * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / hw), p = (int)(i % hw);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // denormalize runs on the fp16 tensor (two roundings), the * 255 / round on float32 numpy
      const float d = __half2float(__float2half_rn(__half2float(__float2half_rn(__half2float(x[((size_t)n * 3 + c) * hw + p]) / 2.f)) + 0.5f));
      y[i * 3 + c] = (uint8_t)__float2int_rn(fminf(fmaxf(d, 0.f), 1.f) * 255.0f);
    }
  }
}

struct TvConv {
  __half* w = nullptr;   // [n_pad, kpad]
  __half* b = nullptr;   // [n_pad] or null (bias=False)
  int cin = 0, cout = 0, n_pad = 0, kpad = 0;
};
struct TvBlock {
  TvConv c0, c2, c4;
};

}  // namespace
}  // namespace l2d

using namespace l2d;

struct l2d_taesd {
  std::vector<void*> blocks;   // device allocations
  int64_t bytes = 0;
  int max_n = 0, H = 0, W = 0;
  TvConv enc_in, enc_out, dec_in, dec_out;
  TvConv enc_down[3], dec_up[3];
  std::vector<TvBlock> enc_blocks, dec_blocks;   // 10 + 10
  __half *a = nullptr, *b = nullptr, *c = nullptr, *cols = nullptr, *head = nullptr, *img = nullptr;
  ~l2d_taesd() {
    for (void* p : blocks) cudaFree(p);
  }
  int alloc(void** out, size_t nbytes) {
    nbytes = (std::max<size_t>(nbytes, 16) + 255) & ~size_t(255);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, nbytes);
    if (e != cudaSuccess) return fail(L2D_ERR_CUDA, std::string("cudaMalloc(") + std::to_string(nbytes) + "): " + cudaGetErrorString(e));
    blocks.push_back(p);
    bytes += (int64_t)nbytes;
    *out = p;
    return L2D_OK;
  }
  int halfs(__half** out, size_t n) { return alloc(reinterpret_cast<void**>(out), n * sizeof(__half)); }
};

namespace {

#define TRC(expr)                  \
  do {                             \
    int _rc = (expr);              \
    if (_rc != L2D_OK) return _rc; \
  } while (0)

const l2d_tensor* tv_find(const l2d_tensor* w, int n, const std::string& name) {
  for (int i = 0; i < n; ++i)
    if (w[i].name && name == w[i].name) return &w[i];
  return nullptr;
}

int tv_load_conv(l2d_taesd* t, const l2d_tensor* w, int n, const std::string& p, int cin, int cout, bool bias, TvConv* c) {
  const l2d_tensor* wt = tv_find(w, n, p + ".weight");
  if (!wt || !wt->data) return fail(L2D_ERR_MISSING, "missing weight tensor '" + p + ".weight'");
  int64_t numel = 1;
  for (int d = 0; d < wt->ndim; ++d) numel *= wt->shape[d];
  if (numel != (int64_t)cout * cin * 9) return fail(L2D_ERR_INVALID, "weight '" + p + ".weight' has the wrong size");
  const __half* bsrc = nullptr;
  if (bias) {
    const l2d_tensor* bt = tv_find(w, n, p + ".bias");
    if (!bt || !bt->data) return fail(L2D_ERR_MISSING, "missing weight tensor '" + p + ".bias'");
    bsrc = static_cast<const __half*>(bt->data);
  }
  c->cin = cin;
  c->cout = cout;
  c->n_pad = (cout + 7) / 8 * 8;
  c->kpad = cin % 8 == 0 ? 9 * cin : 64;
  TRC(t->halfs(&c->w, (size_t)c->n_pad * c->kpad));
  tv_repack_kernel<<<64, 256>>>(static_cast<const __half*>(wt->data), c->w, cout, cin, c->n_pad, c->kpad);
  L2D_LAUNCH_CHECK();
  if (bias) {
    TRC(t->halfs(&c->b, c->n_pad));
    tv_pad_vec_kernel<<<1, 128>>>(bsrc, c->b, cout, c->n_pad);
    L2D_LAUNCH_CHECK();
  }
  return L2D_OK;
}

int tv_load_block(l2d_taesd* t, const l2d_tensor* w, int n, const std::string& p, TvBlock* b) {
  TRC(tv_load_conv(t, w, n, p + ".conv.0", TCH, TCH, true, &b->c0));
  TRC(tv_load_conv(t, w, n, p + ".conv.2", TCH, TCH, true, &b->c2));
  TRC(tv_load_conv(t, w, n, p + ".conv.4", TCH, TCH, true, &b->c4));
  return L2D_OK;
}

inline int tv_blocks(size_t work, int per = 256, int cap = 148 * 16) {
  return (int)std::max<size_t>(1, std::min<size_t>((work + per - 1) / per, (size_t)cap));
}

// 3x3 stride-1 conv over channels-last x [n,h,w,64] as an implicit GEMM
int tv_conv(const TvConv& cv, const __half* x, __half* out, int n, int h, int w, int act, const __half* residual, cudaStream_t st) {
  GemmConstWeights cw;
  return conv3x3_launch(x, n, h, w, cv.cin, cv.w, out, cv.n_pad, cv.n_pad, cv.b, nullptr, 0, 1, residual, cv.n_pad, act, st);
}
// GEMM over an im2col matrix already in t->cols
int tv_conv_cols(l2d_taesd* t, const TvConv& cv, __half* out, int m, int act, cudaStream_t st) {
  GemmConstWeights cw;
  return gemm_launch(t->cols, cv.kpad, cv.w, cv.kpad, out, cv.n_pad, m, cv.n_pad, cv.kpad, cv.b, nullptr, 0, 1, nullptr, 0, act, 0, st);
}
// AutoencoderTinyBlock: out = ReLU(conv4(ReLU(conv2(ReLU(conv0(x))))) + x); t1/t2 scratch, out != x
int tv_block(const TvBlock& b, const __half* x, __half* out, __half* t1, __half* t2, int n, int h, int w, cudaStream_t st) {
  TRC(tv_conv(b.c0, x, t1, n, h, w, L2D_ACT_RELU, nullptr, st));
  TRC(tv_conv(b.c2, t1, t2, n, h, w, L2D_ACT_RELU, nullptr, st));
  TRC(tv_conv(b.c4, t2, out, n, h, w, L2D_ACT_RELU_POST, x, st));
  return L2D_OK;
}

}  // namespace

extern "C" int l2d_taesd_create(l2d_taesd** out, const l2d_tensor* weights, int n_weights, int max_batch, int height, int width) {
  L2D_CHECK_ARG(out && weights && n_weights > 0, "null arguments");
  L2D_CHECK_ARG(max_batch >= 1 && max_batch <= 16, "max_batch must be in 1..16");
  L2D_CHECK_ARG(height > 0 && width > 0 && height % 8 == 0 && width % 8 == 0, "image size must be a multiple of 8");
  for (int s = 0; s < 4; ++s)
    L2D_CHECK_ARG(conv3x3_implicit_supported(max_batch, height >> s, width >> s, TCH),
                  "image size not tileable by the implicit-GEMM conv (power-of-two divisors of H, W must tile 128 pixels)");
  std::unique_ptr<l2d_taesd> t(new l2d_taesd());
  t->max_n = max_batch; t->H = height; t->W = width;
  // ---- parameters (diffusers AutoencoderTiny.state_dict() keys) ----
  int i = 0;
  const int enc_nb[4] = {1, 3, 3, 3}, dec_nb[4] = {3, 3, 3, 1};
  for (int s = 0; s < 4; ++s) {
    const std::string p = "encoder.layers." + std::to_string(i++);
    if (s == 0) TRC(tv_load_conv(t.get(), weights, n_weights, p, 3, TCH, true, &t->enc_in));
    else TRC(tv_load_conv(t.get(), weights, n_weights, p, TCH, TCH, false, &t->enc_down[s - 1]));
    for (int b = 0; b < enc_nb[s]; ++b) {
      TvBlock blk;
      TRC(tv_load_block(t.get(), weights, n_weights, "encoder.layers." + std::to_string(i++), &blk));
      t->enc_blocks.push_back(blk);
    }
  }
  TRC(tv_load_conv(t.get(), weights, n_weights, "encoder.layers." + std::to_string(i), TCH, 4, true, &t->enc_out));
  TRC(tv_load_conv(t.get(), weights, n_weights, "decoder.layers.0", 4, TCH, true, &t->dec_in));
  i = 2;
  for (int s = 0; s < 4; ++s) {
    for (int b = 0; b < dec_nb[s]; ++b) {
      TvBlock blk;
      TRC(tv_load_block(t.get(), weights, n_weights, "decoder.layers." + std::to_string(i++), &blk));
      t->dec_blocks.push_back(blk);
    }
    if (s < 3) {
      ++i;   // nn.Upsample slot
      TRC(tv_load_conv(t.get(), weights, n_weights, "decoder.layers." + std::to_string(i++), TCH, TCH, false, &t->dec_up[s]));
    } else {
      TRC(tv_load_conv(t.get(), weights, n_weights, "decoder.layers." + std::to_string(i++), TCH, 3, true, &t->dec_out));
    }
  }
  // ---- workspace ----
  const size_t m0 = (size_t)max_batch * height * width;
  TRC(t->halfs(&t->a, m0 * TCH));
  TRC(t->halfs(&t->b, m0 * TCH));
  TRC(t->halfs(&t->c, m0 * TCH));
  TRC(t->halfs(&t->cols, std::max(m0 * 64, (m0 / 4) * 9 * TCH)));   // stem im2col [M,64] / stride-2 im2col at half resolution
  TRC(t->halfs(&t->head, m0 * 8));
  TRC(t->halfs(&t->img, m0 * 3));
  L2D_CUDA(cudaDeviceSynchronize());
  *out = t.release();
  return L2D_OK;
}

extern "C" void l2d_taesd_destroy(l2d_taesd* t) { delete t; }
extern "C" int64_t l2d_taesd_device_bytes(const l2d_taesd* t) { return t ? t->bytes : 0; }

extern "C" int l2d_taesd_encode(l2d_taesd* t, const void* image_nchw, void* latents, int n, void* stream) {
  L2D_CHECK_ARG(t && image_nchw && latents, "null arguments");
  L2D_CHECK_ARG(n >= 1 && n <= t->max_n, "batch exceeds max_batch");
  cudaStream_t st = (cudaStream_t)stream;
  int h = t->H, w = t->W;
  const size_t m0 = (size_t)n * h * w;
  tv_im2col_nchw_kernel<<<tv_blocks(m0 * 8), 256, 0, st>>>((const __half*)image_nchw, t->cols, n, 3, h, w, 1);
  L2D_LAUNCH_CHECK();
  __half *x = t->a, *y = t->b, *s1 = t->c;
  TRC(tv_conv_cols(t, t->enc_in, x, (int)m0, L2D_ACT_NONE, st));
  // block scratch: the stride-2 im2col buffer is idle while blocks run
  __half* s2 = t->cols;
  int bi = 0;
  const int enc_nb[4] = {1, 3, 3, 3};
  for (int s = 0; s < 4; ++s) {
    if (s > 0) {
      TRC(l2d_im2col3x3(x, t->cols, n, h, w, TCH, 2, 0, 0, stream));
      h >>= 1;
      w >>= 1;
      TRC(tv_conv_cols(t, t->enc_down[s - 1], y, n * h * w, L2D_ACT_NONE, st));
      std::swap(x, y);
    }
    for (int b = 0; b < enc_nb[s]; ++b) {
      TRC(tv_block(t->enc_blocks[bi++], x, y, s1, s2, n, h, w, st));
      std::swap(x, y);
    }
  }
  TRC(tv_conv(t->enc_out, x, t->head, n, h, w, L2D_ACT_NONE, nullptr, st));
  tv_head_to_nchw_kernel<<<tv_blocks((size_t)n * h * w), 256, 0, st>>>(t->head, (__half*)latents, n, h * w, 4, 0);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

extern "C" int l2d_taesd_decode(l2d_taesd* t, const void* latents, void* image_nchw, int n, int clip, void* stream) {
  L2D_CHECK_ARG(t && latents && image_nchw, "null arguments");
  L2D_CHECK_ARG(n >= 1 && n <= t->max_n, "batch exceeds max_batch");
  cudaStream_t st = (cudaStream_t)stream;
  int h = t->H / 8, w = t->W / 8;
  tv_im2col_nchw_kernel<<<tv_blocks((size_t)n * h * w * 8), 256, 0, st>>>((const __half*)latents, t->cols, n, 4, h, w, 2);
  L2D_LAUNCH_CHECK();
  __half *x = t->a, *y = t->b, *s1 = t->c, *s2 = t->cols;
  TRC(tv_conv_cols(t, t->dec_in, x, n * h * w, L2D_ACT_RELU, st));
  int bi = 0;
  const int dec_nb[4] = {3, 3, 3, 1};
  for (int s = 0; s < 4; ++s) {
    for (int b = 0; b < dec_nb[s]; ++b) {
      TRC(tv_block(t->dec_blocks[bi++], x, y, s1, s2, n, h, w, st));
      std::swap(x, y);
    }
    if (s < 3) {
      tv_upsample2x_kernel<<<tv_blocks((size_t)n * 4 * h * w * (TCH / 8)), 256, 0, st>>>(x, s1, n, h, w, TCH);
      L2D_LAUNCH_CHECK();
      h <<= 1;
      w <<= 1;
      TRC(tv_conv(t->dec_up[s], s1, y, n, h, w, L2D_ACT_NONE, nullptr, st));
      std::swap(x, y);
    }
  }
  TRC(tv_conv(t->dec_out, x, t->head, n, h, w, L2D_ACT_NONE, nullptr, st));
  tv_head_to_nchw_kernel<<<tv_blocks((size_t)n * h * w), 256, 0, st>>>(t->head, (__half*)image_nchw, n, h * w, 3, clip ? 2 : 1);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

extern "C" int l2d_image_u8_to_f16(const void* u8_nhwc, void* f16_nchw, int n, int h, int w, void* stream) {
  L2D_CHECK_ARG(u8_nhwc && f16_nchw && n > 0 && h > 0 && w > 0, "bad arguments");
  tv_u8_to_f16_kernel<<<tv_blocks((size_t)n * h * w), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)u8_nhwc, (__half*)f16_nchw, n, h * w);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}

extern "C" int l2d_image_f16_to_u8(const void* f16_nchw, void* u8_nhwc, int n, int h, int w, void* stream) {
  L2D_CHECK_ARG(u8_nhwc && f16_nchw && n > 0 && h > 0 && w > 0, "bad arguments");
  tv_f16_to_u8_kernel<<<tv_blocks((size_t)n * h * w), 256, 0, (cudaStream_t)stream>>>((const __half*)f16_nchw, (uint8_t*)u8_nhwc, n, h * w);
  L2D_LAUNCH_CHECK();
  return L2D_OK;
}
