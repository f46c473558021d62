// Training-step kernels: train-mode BatchNorm (batch statistics over all A*B maps, running-buffer
// update), its fused BN+ReLU backward, the byte movers of the backward graph (zero-stuffing for stride-2 data
// gradients, nearest-upsample forward / backward, gradient accumulation) and the weight-gradient reduction.
//
// The data gradients themselves run through v2x_conv_fwd (a stride-1 correlation of dy with the transposed,
// 180-degree-rotated filter, v2x_b200/transforms.py); everything here is HBM-bound (one read of each operand, one
// write), except conv_wgrad_kernel, which is FMA-bound on the CUDA cores in this first version (fp32 accumulation of
// the hi+lo operands: exact to fp32 -- the tensor-core version with MN-major operands is the next step).
//
// Reference call sites replaced: nn.BatchNorm2d in .train() mode (CP/models/det/backbone/Backbone.py:102-136 as run by
// CoDetModule.py:217-291 after model.train()), and torch.autograd's conv / batch_norm / relu / interpolate backward.
#include "common.cuh"

namespace v2x {

static unsigned grid_cap(long long work_items, int threads, int per_sm) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)sms * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

// ---------------------------------------------------------------------------------------------
// Per-channel reductions over an act tensor [planes][n_pixels][c] (c % 8 == 0, c <= 1024).
// thread = (channel group g = tid % groups, pixel lane = tid / groups); fp64 accumulation (HBM-bound regardless).
// MODE 0: sum[c] += z, sumsq[c] += z^2                                   (batch statistics)
// MODE 1: s1[c] += dyh, s2[c] += dyh * xhat with dyh = dy * [relu mask], xhat = (z - mean) * invstd   (BN backward)
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) channel_reduce_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ z,
                                                             long long n_pixels, int c, int planes,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             const float* __restrict__ mean, const float* __restrict__ invstd,
                                                             int relu, double* __restrict__ out0, double* __restrict__ out1) {
  extern __shared__ double s_red[];     // [2][c]
  double* s0 = s_red;
  double* s1 = s_red + c;
  const int groups = c / 8;
  const int lanes = blockDim.x / groups;          // pixel lanes per block (host guarantees groups <= 128 -> lanes >= 2)
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups;
  for (int i = threadIdx.x; i < c; i += blockDim.x) { s0[i] = 0.0; s1[i] = 0.0; }
  __syncthreads();
  double acc0[8], acc1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { acc0[e] = 0.0; acc1[e] = 0.0; }
  float sc[8], sh[8], mu[8], is[8];
  if (MODE == 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sc[e] = scale[g * 8 + e]; sh[e] = shift[g * 8 + e]; mu[e] = mean[g * 8 + e]; is[e] = invstd[g * 8 + e];
    }
  }
  const long long plane = n_pixels * c;
  if (lane < lanes) {
    for (long long p = (long long)blockIdx.x * lanes + lane; p < n_pixels; p += (long long)gridDim.x * lanes) {
      float v[8];
      act_load8(a + p * c + g * 8, plane, planes, v);
      if (MODE == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { acc0[e] += (double)v[e]; acc1[e] += (double)v[e] * (double)v[e]; }
      } else {
        float zz[8];
        act_load8(z + p * c + g * 8, plane, planes, zz);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float dyh = (!relu || fmaf(zz[e], sc[e], sh[e]) > 0.f) ? v[e] : 0.f;   // same fmaf as the forward apply
          acc0[e] += (double)dyh;
          acc1[e] += (double)dyh * (double)((zz[e] - mu[e]) * is[e]);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    atomicAdd(&s0[g * 8 + e], acc0[e]);
    atomicAdd(&s1[g * 8 + e], acc1[e]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    atomicAdd(out0 + i, s0[i]);
    atomicAdd(out1 + i, s1[i]);
  }
}

// batch statistics -> the affine form the apply kernels use, and the running-buffer update of nn.BatchNorm2d
// (momentum m: running = (1 - m) * running + m * batch; running_var takes the UNBIASED batch variance)
__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean,
                                   float* __restrict__ invstd, int c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const double m = sum[i] / count;
  double var = sumsq[i] / count - m * m;
  if (var < 0.0) var = 0.0;
  const double is = 1.0 / sqrt(var + (double)eps);
  const double g = gamma ? (double)gamma[i] : 1.0, b = beta ? (double)beta[i] : 0.0;
  scale[i] = (float)(g * is);
  shift[i] = (float)(b - m * g * is);
  mean[i] = (float)m;
  invstd[i] = (float)is;
  if (running_mean) running_mean[i] = (float)((1.0 - momentum) * running_mean[i] + momentum * m);
  if (running_var) running_var[i] = (float)((1.0 - momentum) * running_var[i] + momentum * var * (count / fmax(count - 1.0, 1.0)));
}

// y = [relu](z * scale + shift)
__global__ void bn_relu_apply_kernel(const __nv_bfloat16* __restrict__ z, __nv_bfloat16* __restrict__ y, long long n_pixels, int c,
                                     int planes, const float* __restrict__ scale, const float* __restrict__ shift, int relu) {
  const int groups = c / 8;
  const long long total = n_pixels * groups, plane = n_pixels * c;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(gid % groups);
    const long long off = gid * 8;          // == pixel * c + g * 8
    float v[8];
    act_load8(z + off, plane, planes, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float t = fmaf(v[e], __ldg(scale + g * 8 + e), __ldg(shift + g * 8 + e));
      v[e] = relu ? fmaxf(t, 0.f) : t;
    }
    act_store8(y + off, plane, planes, v);
  }
}

// dz = scale * (dyh - s1 / n - xhat * s2 / n)   (training-mode BatchNorm backward with the ReLU mask folded in)
__global__ void bn_relu_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                                         __nv_bfloat16* __restrict__ dz, long long n_pixels, int c, int planes,
                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                         const float* __restrict__ mean, const float* __restrict__ invstd,
                                         const double* __restrict__ s1, const double* __restrict__ s2, int relu) {
  const int groups = c / 8;
  const long long total = n_pixels * groups, plane = n_pixels * c;
  const double inv_n = 1.0 / (double)n_pixels;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(gid % groups);
    const long long off = gid * 8;
    float v[8], zz[8];
    act_load8(dy + off, plane, planes, v);
    act_load8(z + off, plane, planes, zz);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ch = g * 8 + e;
      const float sc = __ldg(scale + ch);
      const float dyh = (!relu || fmaf(zz[e], sc, __ldg(shift + ch)) > 0.f) ? v[e] : 0.f;
      const float xhat = (zz[e] - __ldg(mean + ch)) * __ldg(invstd + ch);
      v[e] = sc * (dyh - (float)(s1[ch] * inv_n) - xhat * (float)(s2[ch] * inv_n));
    }
    act_store8(dz + off, plane, planes, v);
  }
}

// MODE 0: out[n][2i][2j] = in[n][i][j], zero elsewhere (stride-2 data gradient as a stride-1 correlation)
// MODE 1: out[n][y][x] = in[n][y/2][x/2]                (F.interpolate(scale_factor=2), nearest)
// MODE 2: out[n][i][j] = sum of the 2x2 block of in     (its backward)
template <int MODE>
__global__ void resample2_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int n, int h_out, int w_out,
                                 int c, int planes) {
  const int groups = c / 8;
  const int h_in = MODE == 2 ? h_out * 2 : h_out / 2, w_in = MODE == 2 ? w_out * 2 : w_out / 2;
  const long long total = (long long)n * h_out * w_out * groups;
  const long long in_plane = (long long)n * h_in * w_in * c, out_plane = (long long)n * h_out * w_out * c;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(gid % groups);
    const long long pix = gid / groups;
    const int x = (int)(pix % w_out), y = (int)((pix / w_out) % h_out), im = (int)(pix / ((long long)w_out * h_out));
    float v[8];
    if (MODE == 0) {
      if ((x | y) & 1) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
      } else {
        act_load8(in + (((long long)im * h_in + y / 2) * w_in + x / 2) * c + g * 8, in_plane, planes, v);
      }
    } else if (MODE == 1) {
      act_load8(in + (((long long)im * h_in + y / 2) * w_in + x / 2) * c + g * 8, in_plane, planes, v);
    } else {
      float t[8];
      const __nv_bfloat16* b = in + (((long long)im * h_in + 2 * y) * w_in + 2 * x) * c + g * 8;
      act_load8(b, in_plane, planes, v);
      act_load8(b + c, in_plane, planes, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += t[e];
      act_load8(b + (long long)w_in * c, in_plane, planes, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += t[e];
      act_load8(b + (long long)w_in * c + c, in_plane, planes, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += t[e];
    }
    act_store8(out + pix * c + g * 8, out_plane, planes, v);
  }
}

// dst += src (gradient accumulation where a map feeds two consumers, e.g. the encoder skips)
__global__ void act_add_kernel(__nv_bfloat16* __restrict__ dst, const __nv_bfloat16* __restrict__ src, long long n_elems, int planes) {
  const long long total = n_elems / 8;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x) {
    float a[8], b[8];
    act_load8(dst + gid * 8, n_elems, planes, a);
    act_load8(src + gid * 8, n_elems, planes, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] += b[e];
    act_store8(dst + gid * 8, n_elems, planes, a);
  }
}

// ---------------------------------------------------------------------------------------------
// Weight gradient of a 3x3 (pad 1) / 1x1 conv, stride 1 or 2:
//   dw[co][ci_off + ci][tap] += scale * sum_{n, oh, ow} dz[n][oh][ow][co] * x[n][oh * s + kh - pad][ow * s + kw - pad][ci]
// CTA = a 32 x 32 (co, ci) block of the filter over a range of 4 x 16-pixel output tiles; thread (co = tid & 31,
// ci quad = tid >> 5) keeps its 9 x 4 partial sums in registers across all its tiles and flushes them once with fp32
// atomics.  Operands are converted hi + lo -> fp32 when staged in shared memory, so the accumulation is plain fp32.
// ---------------------------------------------------------------------------------------------
constexpr int kWgTH = 4, kWgTW = 16, kWgC = 32;

template <int STRIDE, int TAPS>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const __nv_bfloat16* __restrict__ dz, const __nv_bfloat16* __restrict__ x,
                                                         int n, int h_out, int w_out, int co, int ci, int planes,
                                                         float* __restrict__ dw, int co_log, int ci_log, int ci_off, int ci_total,
                                                         float scale, int tiles_w, int tiles_per_img, int num_tiles, int co_tiles) {
  constexpr int K = TAPS == 9 ? 3 : 1, PAD = TAPS == 9 ? 1 : 0;
  constexpr int HH = (kWgTH - 1) * STRIDE + K, HW = (kWgTW - 1) * STRIDE + K;   // input halo of one output tile
  __shared__ __align__(16) float s_dz[kWgTH * kWgTW][kWgC];
  __shared__ __align__(16) float s_x[HH * HW][kWgC];
  const int co0 = (blockIdx.y % co_tiles) * kWgC, ci0 = (blockIdx.y / co_tiles) * kWgC;
  const int tco = threadIdx.x & 31, tcg = threadIdx.x >> 5;
  const int h_in = h_out * STRIDE, w_in = w_out * STRIDE;
  const long long dz_plane = (long long)n * h_out * w_out * co, x_plane = (long long)n * h_in * w_in * ci;
  float acc[TAPS][4];
#pragma unroll
  for (int t = 0; t < TAPS; ++t)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[t][e] = 0.f;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int im = tile / tiles_per_img, r = tile - im * tiles_per_img;
    const int oh0 = (r / tiles_w) * kWgTH, ow0 = (r % tiles_w) * kWgTW;
    __syncthreads();   // the previous tile's operands are no longer read
    // stage dz: 64 px x 32 co as 4 vectors of 8 per pixel -> 256 vector loads
    {
      const int px = threadIdx.x >> 2, g = threadIdx.x & 3;
      const int oh = oh0 + px / kWgTW, ow = ow0 + px % kWgTW;
      float v[8];
      if (oh < h_out && ow < w_out && co0 + g * 8 < co) {
        act_load8(dz + (((long long)im * h_out + oh) * w_out + ow) * co + co0 + g * 8, dz_plane, planes, v);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) s_dz[px][g * 8 + e] = v[e];
    }
    for (int u = threadIdx.x; u < HH * HW * 4; u += 256) {
      const int px = u >> 2, g = u & 3;
      const int ih = oh0 * STRIDE - PAD + px / HW, iw = ow0 * STRIDE - PAD + px % HW;
      float v[8];
      if (ih >= 0 && ih < h_in && iw >= 0 && iw < w_in && ci0 + g * 8 < ci) {
        act_load8(x + (((long long)im * h_in + ih) * w_in + iw) * ci + ci0 + g * 8, x_plane, planes, v);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) s_x[px][g * 8 + e] = v[e];
    }
    __syncthreads();
#pragma unroll 2
    for (int px = 0; px < kWgTH * kWgTW; ++px) {
      const float d = s_dz[px][tco];
      const int pr = (px / kWgTW) * STRIDE, pc = (px % kWgTW) * STRIDE;
#pragma unroll
      for (int t = 0; t < TAPS; ++t) {
        const float4 xv = *reinterpret_cast<const float4*>(&s_x[(pr + t / K) * HW + pc + t % K][tcg * 4]);
        acc[t][0] = fmaf(d, xv.x, acc[t][0]);
        acc[t][1] = fmaf(d, xv.y, acc[t][1]);
        acc[t][2] = fmaf(d, xv.z, acc[t][2]);
        acc[t][3] = fmaf(d, xv.w, acc[t][3]);
      }
    }
  }
  const int oc = co0 + tco;
  if (oc < co_log) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ic = ci0 + tcg * 4 + e;
      if (ic < ci_log) {
#pragma unroll
        for (int t = 0; t < TAPS; ++t)
          atomicAdd(dw + ((long long)oc * ci_total + ci_off + ic) * TAPS + t, acc[t][e] * scale);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Zero-hidden ConvGRU gates in train mode (CP/utils/convolutional_rnn/functional.py:84-105 with h = 0).
//   a = conv(cat[h, mean], W_ih) + b_ih  (act [n_pixels][3C], channel order [r | z | n], from v2x_conv_fwd)
//   r = sigmoid(a_r + bhh_r), z = sigmoid(a_z + bhh_z), n = tanh(a_n + r * bhh_n), h' = (1 - z) * n
// Units whose agent slot is absent from their scene pass `pass` through (V2VNet.py:104-107 only rewrites present agents).
// Exact expf / tanhf here (the eval epilogue uses the SFU approximations; the training step is held to fp32 formulas).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool unit_absent(const long long* num_agent, int unit, int batch, int agents) {
  if (num_agent == nullptr) return false;
  const int agent = unit / batch, b = unit % batch;
  return agent >= (int)num_agent[(long long)b * agents];
}
__device__ __forceinline__ float sigmoidf_exact(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void gru_gates_fwd_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ bhh,
                                     const __nv_bfloat16* __restrict__ pass, __nv_bfloat16* __restrict__ h, long long n_pixels,
                                     int hw, int C, int planes, const long long* __restrict__ num_agent, int batch, int agents) {
  const int groups = C / 8;
  const long long total = n_pixels * groups, a_plane = n_pixels * 3 * C, h_plane = n_pixels * C;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(gid % groups);
    const long long pix = gid / groups;
    float v[8];
    if (unit_absent(num_agent, (int)(pix / hw), batch, agents)) {
      act_load8(pass + pix * C + g * 8, h_plane, planes, v);
    } else {
      float ar[8], az[8], an[8];
      const __nv_bfloat16* ap = a + pix * 3 * C + g * 8;
      act_load8(ap, a_plane, planes, ar);
      act_load8(ap + C, a_plane, planes, az);
      act_load8(ap + 2 * C, a_plane, planes, an);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = g * 8 + e;
        const float r = sigmoidf_exact(ar[e] + bhh[c]);
        const float z = sigmoidf_exact(az[e] + bhh[C + c]);
        const float nn = tanhf(an[e] + r * bhh[2 * C + c]);
        v[e] = (1.f - z) * nn;
      }
    }
    act_store8(h + pix * C + g * 8, h_plane, planes, v);
  }
}

// da = d(loss)/d(a) for the three gates, dpass = dh for absent units (else 0), dbhn[c] += sum da_n * r (fp64)
__global__ void gru_gates_bwd_kernel(const __nv_bfloat16* __restrict__ dh, const __nv_bfloat16* __restrict__ a,
                                     const float* __restrict__ bhh, __nv_bfloat16* __restrict__ da, __nv_bfloat16* __restrict__ dpass,
                                     long long n_pixels, int hw, int C, int planes, const long long* __restrict__ num_agent,
                                     int batch, int agents, double* __restrict__ dbhn) {
  __shared__ double s_b[1024];
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_b[i] = 0.0;
  __syncthreads();
  const int groups = C / 8;
  const int lanes = blockDim.x / groups;
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups;
  const long long a_plane = n_pixels * 3 * C, h_plane = n_pixels * C;
  double acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.0;
  if (lane < lanes) {
    for (long long pix = (long long)blockIdx.x * lanes + lane; pix < n_pixels; pix += (long long)gridDim.x * lanes) {
      float d[8], vr[8], vz[8], vn[8], zero[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) zero[e] = 0.f;
      act_load8(dh + pix * C + g * 8, h_plane, planes, d);
      __nv_bfloat16* dap = da + pix * 3 * C + g * 8;
      if (unit_absent(num_agent, (int)(pix / hw), batch, agents)) {
        act_store8(dap, a_plane, planes, zero);
        act_store8(dap + C, a_plane, planes, zero);
        act_store8(dap + 2 * C, a_plane, planes, zero);
        act_store8(dpass + pix * C + g * 8, h_plane, planes, d);
        continue;
      }
      const __nv_bfloat16* ap = a + pix * 3 * C + g * 8;
      act_load8(ap, a_plane, planes, vr);
      act_load8(ap + C, a_plane, planes, vz);
      act_load8(ap + 2 * C, a_plane, planes, vn);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = g * 8 + e;
        const float bn = bhh[2 * C + c];
        const float r = sigmoidf_exact(vr[e] + bhh[c]);
        const float z = sigmoidf_exact(vz[e] + bhh[C + c]);
        const float nn = tanhf(vn[e] + r * bn);
        const float dn = d[e] * (1.f - z);
        const float da_z = -d[e] * nn * z * (1.f - z);
        const float da_n = dn * (1.f - nn * nn);
        const float da_r = da_n * bn * r * (1.f - r);
        vr[e] = da_r; vz[e] = da_z; vn[e] = da_n;
        acc[e] += (double)da_n * (double)r;
      }
      act_store8(dap, a_plane, planes, vr);
      act_store8(dap + C, a_plane, planes, vz);
      act_store8(dap + 2 * C, a_plane, planes, vn);
      act_store8(dpass + pix * C + g * 8, h_plane, planes, zero);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) atomicAdd(&s_b[g * 8 + e], acc[e]);
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dbhn + i, s_b[i]);
}

// ---------------------------------------------------------------------------------------------
// Backward of the cross-agent warp + neighbour mean (v2x_warp_mean_fwd; grid_sample backward, DetModelBase.py:167-168):
//   dx[b, j][tap of (i, pixel)] += w_tap / count_i * dmean[b, i][pixel]      for every participating source j of target i
// One warp per (target unit, output pixel), lanes over channels; fp32 atomics into dx [A*B][H][W][C] (zeroed by the caller).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) warp_mean_bwd_kernel(const __nv_bfloat16* __restrict__ dmean, float* __restrict__ dx,
                                                            const double* __restrict__ trans, const long long* __restrict__ num_agent,
                                                            int batch, int agents, int H, int W, int C, int planes,
                                                            int include_self, int only_v2i) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long total_pix = (long long)batch * agents * H * W;
  const long long plane = total_pix * C;
  for (long long wid = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); wid < total_pix;
       wid += (long long)gridDim.x * warps_per_block) {
    const int ow = (int)(wid % W), oh = (int)((wid / W) % H);
    const int map = (int)(wid / ((long long)W * H));
    const int i = map / batch, b = map % batch;
    const int na = (int)num_agent[(long long)b * agents];
    if (i >= na) continue;
    int count = 0;
    for (int j = 0; j < na && j < agents; ++j) {
      if (j == i && !include_self) continue;
      if (only_v2i && i != 0 && j != 0 && j != i) continue;
      ++count;
    }
    if (count == 0) continue;
    const float inv = 1.f / (float)count;
    const float gx = (2.f * ow + 1.f) / W - 1.f, gy = (2.f * oh + 1.f) / H - 1.f;
    for (int j = 0; j < na && j < agents; ++j) {
      if (j == i && !include_self) continue;
      if (only_v2i && i != 0 && j != 0 && j != i) continue;
      float t00, t01, t02, t10, t11, t12;
      if (j == i) {
        t00 = 1.f; t01 = 0.f; t02 = 0.f; t10 = 0.f; t11 = 1.f; t12 = 0.f;
      } else {   // same theta' as warp_mean_kernel (un-flipped domain)
        const double* T = trans + ((((long long)b * agents + j) * agents + i) << 4);
        t00 = (float)T[0]; t01 = -(float)T[1]; t02 = -(float)T[3] * (1.f / 32.f);
        t10 = -(float)T[4]; t11 = (float)T[5]; t12 = (float)T[7] * (1.f / 32.f);
      }
      const float sx = t00 * gx + t01 * gy + t02, sy = t10 * gx + t11 * gy + t12;
      const float ix = ((sx + 1.f) * W - 1.f) * 0.5f, iy = ((sy + 1.f) * H - 1.f) * 0.5f;
      const float fx = floorf(ix), fy = floorf(iy);
      const int x0 = (int)fx, y0 = (int)fy;
      const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
      const long long src_map = (long long)batch * j + b;
      for (int c0 = lane * 8; c0 < C; c0 += 256) {
        float d[8];
        act_load8(dmean + wid * C + c0, plane, planes, d);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
          if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
          const float wgt = ((t & 1) ? wx1 : wx0) * ((t >> 1) ? wy1 : wy0) * inv;
          float* dst = dx + ((src_map * H + yy) * W + xx) * C + c0;
#pragma unroll
          for (int e = 0; e < 8; ++e) atomicAdd(dst + e, wgt * d[e]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Backward of the Mean / Sum / Max fusion reduce (v2x_warp_reduce_fwd; MeanFusion.py:11-12, SumFusion.py:20-21,
// MaxFusion.py:20-21 under loss.backward()): members of target i = {i} U {present j != i allowed by only_v2i};
//   mean / sum : dx[b,j][tap] += w_tap * s * dout[b,i][pixel],  dx[b,i][pixel] += s * dout   (s = 1/count or 1)
//   max        : per channel the gradient goes to the FIRST member (list order: self, then ascending j) that attains the
//                maximum -- torch.max(torch.stack(list), 0) on CPU -- i.e. through that member's four bilinear taps; the
//                member values are recomputed from the saved forward input x
// Absent agent slots keep their own map in the forward (FusionBase.py:41-63), so their gradient passes straight through.
// One warp per (target unit, output pixel), lanes over channels; fp32 atomics into dx [A*B][H][W][C] (zeroed by the caller).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) warp_reduce_bwd_kernel(const __nv_bfloat16* __restrict__ dout,
                                                              const __nv_bfloat16* __restrict__ x, float* __restrict__ dx,
                                                              const float* __restrict__ coef,
                                                              const double* __restrict__ trans,
                                                              const long long* __restrict__ num_agent, int batch, int agents,
                                                              int H, int W, int C, int planes, int mode, int only_v2i) {
  constexpr int kMaxA = 8;
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long total_pix = (long long)batch * agents * H * W;
  const long long plane = total_pix * C;
  for (long long wid = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); wid < total_pix;
       wid += (long long)gridDim.x * warps_per_block) {
    const int ow = (int)(wid % W), oh = (int)((wid / W) % H);
    const int map = (int)(wid / ((long long)W * H));
    const int i = map / batch, b = map % batch;
    const int na = min((int)num_agent[(long long)b * agents], agents);
    if (i >= na) {   // pass-through
      for (int c0 = lane * 8; c0 < C; c0 += 256) {
        float d[8];
        act_load8(dout + wid * C + c0, plane, planes, d);
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(dx + wid * C + c0 + e, d[e]);
      }
      continue;
    }
    // sample positions of the participating neighbours
    const float gx = (2.f * ow + 1.f) / W - 1.f, gy = (2.f * oh + 1.f) / H - 1.f;
    float sx_[kMaxA], sy_[kMaxA];
    unsigned use = 0;
    int count = 1;
#pragma unroll
    for (int j = 0; j < kMaxA; ++j) {
      sx_[j] = 0.f; sy_[j] = 0.f;
      if (j >= na || j == i) continue;
      if (only_v2i && i != 0 && j != 0) continue;
      use |= 1u << j;
      ++count;
      const double* T = trans + ((((long long)b * agents + j) * agents + i) << 4);
      const float t00 = (float)T[0], t01 = -(float)T[1], t02 = -(float)T[3] * (1.f / 32.f);
      const float t10 = -(float)T[4], t11 = (float)T[5], t12 = (float)T[7] * (1.f / 32.f);
      const float sx = t00 * gx + t01 * gy + t02, sy = t10 * gx + t11 * gy + t12;
      sx_[j] = ((sx + 1.f) * W - 1.f) * 0.5f;
      sy_[j] = ((sy + 1.f) * H - 1.f) * 0.5f;
    }
    const float s = mode == 0 ? 1.f / (float)count : 1.f;
    // mode 3 (AgentWiseWeightedFusion.py:27-34): out = sum_k coef[b,i,k] * member_k with constant per-pair coefficients
    const float* cf = mode == 3 ? coef + ((long long)b * agents + i) * agents : nullptr;
    for (int c0 = lane * 8; c0 < C; c0 += 256) {
      float d[8];
      act_load8(dout + wid * C + c0, plane, planes, d);
      int who[8];   // max mode: winning member per channel (-1 = self)
#pragma unroll
      for (int e = 0; e < 8; ++e) who[e] = -1;
      if (mode == 2) {
        float best[8];
        act_load8(x + wid * C + c0, plane, planes, best);
#pragma unroll
        for (int j = 0; j < kMaxA; ++j) {
          if (!((use >> j) & 1u)) continue;
          const float fx = floorf(sx_[j]), fy = floorf(sy_[j]);
          const int x0 = (int)fx, y0 = (int)fy;
          const float wx1 = sx_[j] - fx, wy1 = sy_[j] - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
          float val[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) val[e] = 0.f;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
            if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
            const float wgt = ((t & 1) ? wx1 : wx0) * ((t >> 1) ? wy1 : wy0);
            float f[8];
            act_load8(x + ((((long long)batch * j + b) * H + yy) * W + xx) * C + c0, plane, planes, f);
#pragma unroll
            for (int e = 0; e < 8; ++e) val[e] += wgt * f[e];
          }
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (val[e] > best[e]) { best[e] = val[e]; who[e] = j; }
        }
      }
      // self term
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (mode != 2 || who[e] < 0) atomicAdd(dx + wid * C + c0 + e, (cf ? cf[i] : s) * d[e]);
      // neighbour terms: grid_sample backward through the four taps
#pragma unroll
      for (int j = 0; j < kMaxA; ++j) {
        if (!((use >> j) & 1u)) continue;
        const float fx = floorf(sx_[j]), fy = floorf(sy_[j]);
        const int x0 = (int)fx, y0 = (int)fy;
        const float wx1 = sx_[j] - fx, wy1 = sy_[j] - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
          if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
          const float wgt = ((t & 1) ? wx1 : wx0) * ((t >> 1) ? wy1 : wy0) * (cf ? cf[j] : s);
          float* dst = dx + ((((long long)batch * j + b) * H + yy) * W + xx) * C + c0;
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (mode != 2 || who[e] == j) atomicAdd(dst + e, wgt * d[e]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Backward of the when2com gated fuse (v2x_warp_gated_fwd; When2com.py:199-225 val_mat + :397-412 weighted sum under
// loss.backward()).  With val[b,k,q] = (k == q ? x[b,q] : warp(x[b,q], T[b,q,k])) (warp_flag 1, SURVEY Q8 pairing) or
// x[b,k] (warp_flag 0) and out[b,q] = sum_k coef[b,k,q] * val[b,k,q]:
//   dcoef[b,k,q] = < dout[b,q], val[b,k,q] >                          (sum over pixels and channels; fp32 atomics)
//   dx           += coef[b,k,q] * (grid_sample backward of dout[b,q] through val[b,k,q]'s taps)
// One warp per (target unit, output pixel), lanes over channels.  dx fp32 [A*B][H][W][C] and dcoef fp32 [B][A][A] are
// zeroed by the launcher; dcoef carries the gradient scale of dout.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) warp_gated_bwd_kernel(const __nv_bfloat16* __restrict__ dout,
                                                             const __nv_bfloat16* __restrict__ x, float* __restrict__ dx,
                                                             float* __restrict__ dcoef, const float* __restrict__ coef,
                                                             const double* __restrict__ trans,
                                                             const long long* __restrict__ num_agent, int batch, int agents,
                                                             int H, int W, int C, int planes, int warp_flag, int only_v2i) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long total_pix = (long long)batch * agents * H * W;
  const long long plane = total_pix * C;
  for (long long wid = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); wid < total_pix;
       wid += (long long)gridDim.x * warps_per_block) {
    const int ow = (int)(wid % W), oh = (int)((wid / W) % H);
    const int map = (int)(wid / ((long long)W * H));
    const int q = map / batch, b = map % batch;
    const int na = min((int)num_agent[(long long)b * agents], agents);
    const int nterms = warp_flag ? (q < na ? na : 0) : agents;
    const float gx = (2.f * ow + 1.f) / W - 1.f, gy = (2.f * oh + 1.f) / H - 1.f;
    for (int k = 0; k < nterms; ++k) {
      if (warp_flag && only_v2i && k != q && k != 0 && q != 0) continue;
      const float cf = coef[((long long)b * agents + k) * agents + q];
      const long long src_map = warp_flag ? (long long)batch * q + b : (long long)batch * k + b;
      float wts[4];
      int xs[4], ys[4];
      int ntap = 4;
      if (!warp_flag || k == q) {
        ntap = 1; wts[0] = 1.f; xs[0] = ow; ys[0] = oh;
      } else {
        const double* T = trans + ((((long long)b * agents + q) * agents + k) << 4);
        const float t00 = (float)T[0], t01 = -(float)T[1], t02 = -(float)T[3] * (1.f / 32.f);
        const float t10 = -(float)T[4], t11 = (float)T[5], t12 = (float)T[7] * (1.f / 32.f);
        const float sx = t00 * gx + t01 * gy + t02, sy = t10 * gx + t11 * gy + t12;
        const float ix = ((sx + 1.f) * W - 1.f) * 0.5f, iy = ((sy + 1.f) * H - 1.f) * 0.5f;
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = (int)fx, y0 = (int)fy;
        const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          xs[t] = x0 + (t & 1); ys[t] = y0 + (t >> 1);
          wts[t] = ((t & 1) ? wx1 : wx0) * ((t >> 1) ? wy1 : wy0);
        }
      }
      float dot = 0.f;
      for (int c0 = lane * 8; c0 < C; c0 += 256) {
        float d[8];
        act_load8(dout + wid * C + c0, plane, planes, d);
        for (int t = 0; t < ntap; ++t) {
          if (xs[t] < 0 || xs[t] >= W || ys[t] < 0 || ys[t] >= H) continue;
          const long long off = ((src_map * H + ys[t]) * W + xs[t]) * C + c0;
          float f[8];
          act_load8(x + off, plane, planes, f);
          const float wc = wts[t] * cf;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            dot = fmaf(wts[t] * f[e], d[e], dot);
            if (cf != 0.f) atomicAdd(dx + off + e, wc * d[e]);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      if (lane == 0) atomicAdd(dcoef + ((long long)b * agents + k) * agents + q, dot);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Backward of DiscoNet's per-pixel weighted fuse (v2x_warp_weighted_fwd, coef_mode 1; DiscoNet.py:80-107 under
// loss.backward()): with members k of target i (self + participating present agents), w_k[p] = softmax_k(s[b,i,k,p]) and
// out[b,i,p] = sum_k w_k[p] * member_k[p]:
//   g_k[p]        = < dout[b,i,p,:], member_k[p,:] >            (dot over channels)
//   ds[b,i,k,p]   = w_k[p] * (g_k[p] - sum_m w_m[p] g_m[p])     (softmax backward, per pixel)
//   dx           += w_k[p] * (grid_sample backward of dout[b,i,p] through member k's taps)
// One warp per (target unit, pixel), lanes over channels.  dx fp32 [A*B][H][W][C] and dscores fp32 [B][A][A][HW] are zeroed
// by the launcher; dscores carries the gradient scale of dout.  Absent agent slots pass their gradient through.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) warp_weighted_bwd_kernel(const __nv_bfloat16* __restrict__ dout,
                                                                const __nv_bfloat16* __restrict__ x, float* __restrict__ dx,
                                                                float* __restrict__ dscores, const float* __restrict__ scores,
                                                                const double* __restrict__ trans,
                                                                const long long* __restrict__ num_agent, int batch, int agents,
                                                                int H, int W, int C, int planes, int only_v2i) {
  constexpr int kMaxA = 8;
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int HW = H * W;
  const long long total_pix = (long long)batch * agents * HW;
  const long long plane = total_pix * C;
  for (long long wid = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); wid < total_pix;
       wid += (long long)gridDim.x * warps_per_block) {
    const int p = (int)(wid % HW);
    const int ow = p % W, oh = p / W;
    const int map = (int)(wid / HW);
    const int i = map / batch, b = map % batch;
    const int na = min((int)num_agent[(long long)b * agents], agents);
    if (i >= na) {   // pass-through
      for (int c0 = lane * 8; c0 < C; c0 += 256) {
        float d[8];
        act_load8(dout + wid * C + c0, plane, planes, d);
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(dx + wid * C + c0 + e, d[e]);
      }
      continue;
    }
    const float* sb = scores + ((long long)b * agents + i) * agents * HW;
    float wk[kMaxA], g[kMaxA];
    unsigned use = 0;
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < kMaxA; ++k) {
      wk[k] = 0.f; g[k] = 0.f;
      if (k >= na) continue;
      if (k != i && only_v2i && i != 0 && k != 0) continue;
      use |= 1u << k;
      m = fmaxf(m, sb[(long long)k * HW + p]);
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxA; ++k)
      if ((use >> k) & 1u) { wk[k] = __expf(sb[(long long)k * HW + p] - m); sum += wk[k]; }
    const float inv = 1.f / sum;
    const float gx = (2.f * ow + 1.f) / W - 1.f, gy = (2.f * oh + 1.f) / H - 1.f;
#pragma unroll
    for (int k = 0; k < kMaxA; ++k) {
      if (!((use >> k) & 1u)) continue;
      wk[k] *= inv;
      float wts[4];
      int xs[4], ys[4];
      int ntap = 4;
      if (k == i) {
        ntap = 1; wts[0] = 1.f; xs[0] = ow; ys[0] = oh;
      } else {
        const double* T = trans + ((((long long)b * agents + k) * agents + i) << 4);
        const float t00 = (float)T[0], t01 = -(float)T[1], t02 = -(float)T[3] * (1.f / 32.f);
        const float t10 = -(float)T[4], t11 = (float)T[5], t12 = (float)T[7] * (1.f / 32.f);
        const float sx = t00 * gx + t01 * gy + t02, sy = t10 * gx + t11 * gy + t12;
        const float ix = ((sx + 1.f) * W - 1.f) * 0.5f, iy = ((sy + 1.f) * H - 1.f) * 0.5f;
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = (int)fx, y0 = (int)fy;
        const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          xs[t] = x0 + (t & 1); ys[t] = y0 + (t >> 1);
          wts[t] = ((t & 1) ? wx1 : wx0) * ((t >> 1) ? wy1 : wy0);
        }
      }
      const long long src_map = (long long)batch * k + b;
      float dot = 0.f;
      for (int c0 = lane * 8; c0 < C; c0 += 256) {
        float d[8];
        act_load8(dout + wid * C + c0, plane, planes, d);
        for (int t = 0; t < ntap; ++t) {
          if (xs[t] < 0 || xs[t] >= W || ys[t] < 0 || ys[t] >= H) continue;
          const long long off = ((src_map * H + ys[t]) * W + xs[t]) * C + c0;
          float f[8];
          act_load8(x + off, plane, planes, f);
          const float wc = wts[t] * wk[k];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            dot = fmaf(wts[t] * f[e], d[e], dot);
            atomicAdd(dx + off + e, wc * d[e]);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      g[k] = dot;
    }
    float G = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxA; ++k) G += wk[k] * g[k];
    if (lane == 0) {
      float* db = dscores + ((long long)b * agents + i) * agents * HW;
#pragma unroll
      for (int k = 0; k < kMaxA; ++k)
        if ((use >> k) & 1u) db[(long long)k * HW + p] = wk[k] * (g[k] - G);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Segmentation UNet pieces, backward (CP/models/seg/SegModelBase.py:113,125 under loss.backward()).
// ---------------------------------------------------------------------------------------------
// nn.MaxPool2d(2) backward: the gradient of an output pixel goes to the FIRST maximum of its 2x2 window in scan order
// (ATen keeps a value only if it is strictly greater than the running maximum), the other three get zero.
__global__ void maxpool2_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                    __nv_bfloat16* __restrict__ dx, int n, int h, int w, int c, int planes) {
  const int groups = c / 8;
  const long long total = (long long)n * h * w * groups;
  const long long in_plane = (long long)n * 4 * h * w * c, out_plane = (long long)n * h * w * c;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(gid % groups);
    const long long pix = gid / groups;
    const int ox = (int)(pix % w), oy = (int)((pix / w) % h), im = (int)(pix / ((long long)w * h));
    const long long base = (((long long)im * 2 * h + 2 * oy) * 2 * w + 2 * ox) * c + g * 8;
    const long long off[4] = {0, (long long)c, (long long)2 * w * c, (long long)2 * w * c + c};
    float v[4][8], d[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) act_load8(x + base + off[t], in_plane, planes, v[t]);
    act_load8(dy + pix * c + g * 8, out_plane, planes, d);
    int arg[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int a = 0;
      float m = v[0][e];
#pragma unroll
      for (int t = 1; t < 4; ++t)
        if (v[t][e] > m) { m = v[t][e]; a = t; }
      arg[e] = a;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = arg[e] == t ? d[e] : 0.f;
      act_store8(dx + base + off[t], in_plane, planes, o);
    }
  }
}

// nn.Upsample(scale_factor=2, bilinear, align_corners=True) backward: scatter of dy [n][2h][2w][c] into dx fp32 [n][h][w][c]
// (zeroed by the caller) with the forward's interpolation weights
__global__ void upsample_bilinear2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ dx, int n, int h, int w,
                                              int c, int planes) {
  const int groups = c / 8;
  const int oh_n = 2 * h, ow_n = 2 * w;
  const long long total = (long long)n * oh_n * ow_n * groups;
  const long long plane = (long long)n * oh_n * ow_n * c;
  const float sy = oh_n > 1 ? (float)(h - 1) / (float)(oh_n - 1) : 0.f;
  const float sx = ow_n > 1 ? (float)(w - 1) / (float)(ow_n - 1) : 0.f;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(gid % groups);
    const long long pix = gid / groups;
    const int ox = (int)(pix % ow_n), oy = (int)((pix / ow_n) % oh_n), im = (int)(pix / ((long long)ow_n * oh_n));
    const float fy = oy * sy, fx = ox * sx;
    const int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float wy1 = fy - (float)y0, wx1 = fx - (float)x0, wy0 = 1.f - wy1, wx0 = 1.f - wx1;
    float d[8];
    act_load8(dy + pix * c + g * 8, plane, planes, d);
    float* b = dx + (long long)im * h * w * c + g * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(b + ((long long)y0 * w + x0) * c + e, wy0 * wx0 * d[e]);
      atomicAdd(b + ((long long)y0 * w + x1) * c + e, wy0 * wx1 * d[e]);
      atomicAdd(b + ((long long)y1 * w + x0) * c + e, wy1 * wx0 * d[e]);
      atomicAdd(b + ((long long)y1 * w + x1) * c + e, wy1 * wx1 * d[e]);
    }
  }
}

// out[i] (+)= scale * (float)in[i]   (per-channel double sums -> fp32 parameter gradients)
__global__ void scale_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int count, float scale, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = (accumulate ? out[i] : 0.f) + (float)(in[i] * (double)scale);
}

}  // namespace v2x

using namespace v2x;

#define V2X_CHECK_ACT(c_, planes_)                                                                  \
  V2X_REQUIRE((c_) > 0 && (c_) % 8 == 0 && (c_) <= 1024, "channels must be a multiple of 8, <= 1024"); \
  V2X_REQUIRE((planes_) == 1 || (planes_) == 2, "planes must be 1 or 2")

extern "C" int v2x_bn_stats_fwd(const void* z, int64_t n_pixels, int32_t c, int32_t planes, double* sum, double* sumsq,
                                void* stream) {
  V2X_REQUIRE(z && sum && sumsq && n_pixels > 0, "null/empty");
  V2X_REQUIRE(c > 0 && c % 8 == 0 && c <= 2048 && (planes == 1 || planes == 2), "channels: multiple of 8, <= 2048; planes 1 or 2");
  cudaStream_t s = (cudaStream_t)stream;
  V2X_CUDA_TRY(cudaMemsetAsync(sum, 0, sizeof(double) * c, s));
  V2X_CUDA_TRY(cudaMemsetAsync(sumsq, 0, sizeof(double) * c, s));
  const int lanes = 256 / (c / 8);
  channel_reduce_kernel<0><<<grid_cap((n_pixels + lanes - 1) / lanes * 256, 256, 4), 256, 2 * c * sizeof(double), s>>>(
      reinterpret_cast<const __nv_bfloat16*>(z), nullptr, n_pixels, c, planes, nullptr, nullptr, nullptr, nullptr, 0, sum, sumsq);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_bn_finalize(const double* sum, const double* sumsq, int64_t count, const float* gamma, const float* beta,
                               float eps, float momentum, float* running_mean, float* running_var, float* scale,
                               float* shift, float* mean, float* invstd, int32_t c, void* stream) {
  V2X_REQUIRE(sum && sumsq && scale && shift && mean && invstd && c > 0 && count > 0, "null/empty");
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sum, sumsq, (double)count, gamma, beta, eps, momentum,
                                                                       running_mean, running_var, scale, shift, mean, invstd, c);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_bn_relu_apply_fwd(const void* z, void* y, int64_t n_pixels, int32_t c, int32_t planes, const float* scale,
                                     const float* shift, int32_t relu, void* stream) {
  V2X_REQUIRE(z && y && scale && shift && n_pixels > 0, "null/empty");
  V2X_CHECK_ACT(c, planes);
  bn_relu_apply_kernel<<<grid_cap(n_pixels * (c / 8), 256, 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(z), reinterpret_cast<__nv_bfloat16*>(y), n_pixels, c, planes, scale, shift, relu);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_bn_relu_bwd(const void* dy, const void* z, void* dz, int64_t n_pixels, int32_t c, int32_t planes,
                               const float* scale, const float* shift, const float* mean, const float* invstd, int32_t relu,
                               double* s1, double* s2, void* stream) {
  V2X_REQUIRE(dy && z && dz && scale && shift && mean && invstd && s1 && s2 && n_pixels > 0, "null/empty");
  V2X_CHECK_ACT(c, planes);
  cudaStream_t s = (cudaStream_t)stream;
  V2X_CUDA_TRY(cudaMemsetAsync(s1, 0, sizeof(double) * c, s));
  V2X_CUDA_TRY(cudaMemsetAsync(s2, 0, sizeof(double) * c, s));
  const int lanes = 256 / (c / 8);
  channel_reduce_kernel<1><<<grid_cap((n_pixels + lanes - 1) / lanes * 256, 256, 4), 256, 2 * c * sizeof(double), s>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(z), n_pixels, c, planes, scale, shift,
      mean, invstd, relu, s1, s2);
  bn_relu_bwd_apply_kernel<<<grid_cap(n_pixels * (c / 8), 256, 8), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(z), reinterpret_cast<__nv_bfloat16*>(dz),
      n_pixels, c, planes, scale, shift, mean, invstd, s1, s2, relu);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_resample2(const void* in, void* out, int32_t n, int32_t h_out, int32_t w_out, int32_t c, int32_t planes,
                             int32_t mode, void* stream) {
  V2X_REQUIRE(in && out && n > 0 && h_out > 0 && w_out > 0, "null/empty");
  V2X_CHECK_ACT(c, planes);
  V2X_REQUIRE(mode >= 0 && mode <= 2 && (mode == 2 || (h_out % 2 == 0 && w_out % 2 == 0)), "bad mode / odd output size");
  const long long total = (long long)n * h_out * w_out * (c / 8);
  const unsigned grid = grid_cap(total, 256, 8);
  const __nv_bfloat16* i = reinterpret_cast<const __nv_bfloat16*>(in);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == 0) resample2_kernel<0><<<grid, 256, 0, s>>>(i, o, n, h_out, w_out, c, planes);
  else if (mode == 1) resample2_kernel<1><<<grid, 256, 0, s>>>(i, o, n, h_out, w_out, c, planes);
  else resample2_kernel<2><<<grid, 256, 0, s>>>(i, o, n, h_out, w_out, c, planes);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_act_add(void* dst, const void* src, int64_t n_elems, int32_t planes, void* stream) {
  V2X_REQUIRE(dst && src && n_elems > 0 && n_elems % 8 == 0, "null/empty (elements per plane must be a multiple of 8)");
  V2X_REQUIRE(planes == 1 || planes == 2, "planes must be 1 or 2");
  act_add_kernel<<<grid_cap(n_elems / 8, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<__nv_bfloat16*>(dst), reinterpret_cast<const __nv_bfloat16*>(src), n_elems, planes);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_conv_wgrad(const void* dz, const void* x, int32_t n, int32_t h_out, int32_t w_out, int32_t co, int32_t ci,
                              int32_t planes, int32_t stride, int32_t taps, float* dw, int32_t co_log, int32_t ci_log,
                              int32_t ci_off, int32_t ci_total, float scale, void* stream) {
  V2X_REQUIRE(dz && x && dw && n > 0 && h_out > 0 && w_out > 0, "null/empty");
  V2X_CHECK_ACT(co, planes);
  V2X_CHECK_ACT(ci, planes);
  V2X_REQUIRE((stride == 1 || stride == 2) && (taps == 9 || (taps == 1 && stride == 1)), "3x3 stride 1/2 or 1x1 stride 1");
  V2X_REQUIRE(co_log > 0 && co_log <= co && ci_log > 0 && ci_log <= ci && ci_off >= 0 && ci_off + ci_log <= ci_total,
              "bad logical channel window");
  const int tiles_w = (w_out + kWgTW - 1) / kWgTW, tiles_h = (h_out + kWgTH - 1) / kWgTH;
  const int tiles_per_img = tiles_w * tiles_h, num_tiles = n * tiles_per_img;
  const int co_tiles = (co_log + kWgC - 1) / kWgC;
  const int pairs = co_tiles * ((ci_log + kWgC - 1) / kWgC);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int gx = (4 * sms + pairs - 1) / pairs;          // ~4 CTAs per SM over all (co, ci) blocks
  if (gx > num_tiles) gx = num_tiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, pairs);
  const __nv_bfloat16* a = reinterpret_cast<const __nv_bfloat16*>(dz);
  const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(x);
  cudaStream_t s = (cudaStream_t)stream;
#define V2X_WG(S_, T_)                                                                                              \
  conv_wgrad_kernel<S_, T_><<<grid, 256, 0, s>>>(a, b, n, h_out, w_out, co, ci, planes, dw, co_log, ci_log, ci_off, \
                                                 ci_total, scale, tiles_w, tiles_per_img, num_tiles, co_tiles)
  if (taps == 1) V2X_WG(1, 1);
  else if (stride == 1) V2X_WG(1, 9);
  else V2X_WG(2, 9);
#undef V2X_WG
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_scale_to_f32(const double* in, float* out, int32_t count, float scale, int32_t accumulate, void* stream) {
  V2X_REQUIRE(in && out && count > 0, "null/empty");
  scale_to_f32_kernel<<<(count + 127) / 128, 128, 0, (cudaStream_t)stream>>>(in, out, count, scale, accumulate);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_gru_gates_fwd(const void* a, const float* bhh, const void* pass, void* h, int64_t n_pixels, int32_t hw,
                                 int32_t c, int32_t planes, const int64_t* num_agent, int32_t batch, int32_t agents,
                                 void* stream) {
  V2X_REQUIRE(a && bhh && h && n_pixels > 0 && hw > 0, "null/empty");
  V2X_CHECK_ACT(c, planes);
  V2X_REQUIRE(num_agent == nullptr || (pass && batch > 0 && agents > 0 && n_pixels == (int64_t)batch * agents * hw),
              "num_agent needs pass and n_pixels == batch * agents * hw");
  gru_gates_fwd_kernel<<<grid_cap(n_pixels * (c / 8), 256, 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(a), bhh, reinterpret_cast<const __nv_bfloat16*>(pass),
      reinterpret_cast<__nv_bfloat16*>(h), n_pixels, hw, c, planes, reinterpret_cast<const long long*>(num_agent), batch, agents);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_gru_gates_bwd(const void* dh, const void* a, const float* bhh, void* da, void* dpass, int64_t n_pixels,
                                 int32_t hw, int32_t c, int32_t planes, const int64_t* num_agent, int32_t batch, int32_t agents,
                                 double* dbhn, void* stream) {
  V2X_REQUIRE(dh && a && bhh && da && dpass && dbhn && n_pixels > 0 && hw > 0, "null/empty");
  V2X_CHECK_ACT(c, planes);
  cudaStream_t s = (cudaStream_t)stream;
  V2X_CUDA_TRY(cudaMemsetAsync(dbhn, 0, sizeof(double) * c, s));
  const int lanes = 256 / (c / 8);
  gru_gates_bwd_kernel<<<grid_cap((n_pixels + lanes - 1) / lanes * 256, 256, 4), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(dh), reinterpret_cast<const __nv_bfloat16*>(a), bhh,
      reinterpret_cast<__nv_bfloat16*>(da), reinterpret_cast<__nv_bfloat16*>(dpass), n_pixels, hw, c, planes,
      reinterpret_cast<const long long*>(num_agent), batch, agents, dbhn);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_warp_mean_bwd(const void* dmean, float* dx, const double* trans, const int64_t* num_agent, int32_t batch,
                                 int32_t agents, int32_t h, int32_t w, int32_t c, int32_t planes, int32_t include_self,
                                 int32_t only_v2i, void* stream) {
  V2X_REQUIRE(dmean && dx && trans && num_agent && batch > 0 && agents > 0 && h > 0 && w > 0, "null/empty");
  V2X_CHECK_ACT(c, planes);
  cudaStream_t s = (cudaStream_t)stream;
  const long long total_pix = (long long)batch * agents * h * w;
  V2X_CUDA_TRY(cudaMemsetAsync(dx, 0, sizeof(float) * total_pix * c, s));
  warp_mean_bwd_kernel<<<grid_cap(total_pix * 32, 256, 8), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(dmean), dx, trans, reinterpret_cast<const long long*>(num_agent), batch, agents, h, w,
      c, planes, include_self, only_v2i);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_warp_reduce_bwd(const void* dout, const void* x, float* dx, const float* coef, const double* trans,
                                   const int64_t* num_agent, int32_t batch, int32_t agents, int32_t h, int32_t w, int32_t c,
                                   int32_t planes, int32_t mode, int32_t only_v2i, void* stream) {
  V2X_REQUIRE(dout && dx && trans && num_agent && batch > 0 && agents > 0 && agents <= 8 && h > 0 && w > 0,
              "null/empty (agents <= 8)");
  V2X_REQUIRE(mode >= 0 && mode <= 3, "mode must be 0 (mean), 1 (sum), 2 (max) or 3 (per-pair coefficients)");
  V2X_REQUIRE(mode != 2 || x, "max mode needs the forward input x");
  V2X_REQUIRE(mode != 3 || coef, "mode 3 needs coef [B][A][A]");
  V2X_CHECK_ACT(c, planes);
  cudaStream_t s = (cudaStream_t)stream;
  const long long total_pix = (long long)batch * agents * h * w;
  V2X_CUDA_TRY(cudaMemsetAsync(dx, 0, sizeof(float) * total_pix * c, s));
  warp_reduce_bwd_kernel<<<grid_cap(total_pix * 32, 256, 8), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(dout), reinterpret_cast<const __nv_bfloat16*>(x), dx, coef, trans,
      reinterpret_cast<const long long*>(num_agent), batch, agents, h, w, c, planes, mode, only_v2i);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_warp_gated_bwd(const void* dout, const void* x, float* dx, float* dcoef, const float* coef,
                                  const double* trans, const int64_t* num_agent, int32_t batch, int32_t agents, int32_t h,
                                  int32_t w, int32_t c, int32_t planes, int32_t warp_flag, int32_t only_v2i, void* stream) {
  V2X_REQUIRE(dout && x && dx && dcoef && coef && trans && num_agent, "null pointer");
  V2X_REQUIRE(batch > 0 && agents > 0 && h > 0 && w > 0, "empty geometry");
  V2X_CHECK_ACT(c, planes);
  cudaStream_t s = (cudaStream_t)stream;
  const long long total_pix = (long long)batch * agents * h * w;
  V2X_CUDA_TRY(cudaMemsetAsync(dx, 0, sizeof(float) * total_pix * c, s));
  V2X_CUDA_TRY(cudaMemsetAsync(dcoef, 0, sizeof(float) * (size_t)batch * agents * agents, s));
  warp_gated_bwd_kernel<<<grid_cap(total_pix * 32, 256, 8), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(dout), reinterpret_cast<const __nv_bfloat16*>(x), dx, dcoef, coef, trans,
      reinterpret_cast<const long long*>(num_agent), batch, agents, h, w, c, planes, warp_flag, only_v2i);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_warp_weighted_bwd(const void* dout, const void* x, float* dx, float* dscores, const float* scores,
                                     const double* trans, const int64_t* num_agent, int32_t batch, int32_t agents, int32_t h,
                                     int32_t w, int32_t c, int32_t planes, int32_t only_v2i, void* stream) {
  V2X_REQUIRE(dout && x && dx && dscores && scores && trans && num_agent, "null pointer");
  V2X_REQUIRE(batch > 0 && agents > 0 && agents <= 8 && h > 0 && w > 0, "empty geometry (agents <= 8)");
  V2X_CHECK_ACT(c, planes);
  cudaStream_t s = (cudaStream_t)stream;
  const long long total_pix = (long long)batch * agents * h * w;
  V2X_CUDA_TRY(cudaMemsetAsync(dx, 0, sizeof(float) * total_pix * c, s));
  V2X_CUDA_TRY(cudaMemsetAsync(dscores, 0, sizeof(float) * total_pix * agents, s));
  warp_weighted_bwd_kernel<<<grid_cap(total_pix * 32, 256, 8), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(dout), reinterpret_cast<const __nv_bfloat16*>(x), dx, dscores, scores, trans,
      reinterpret_cast<const long long*>(num_agent), batch, agents, h, w, c, planes, only_v2i);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_maxpool2_bwd(const void* x, const void* dy, void* dx, int32_t n, int32_t h_out, int32_t w_out, int32_t c,
                                int32_t planes, void* stream) {
  V2X_REQUIRE(x && dy && dx && n > 0 && h_out > 0 && w_out > 0, "null/empty");
  V2X_CHECK_ACT(c, planes);
  maxpool2_bwd_kernel<<<grid_cap((long long)n * h_out * w_out * (c / 8), 256, 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<__nv_bfloat16*>(dx), n,
      h_out, w_out, c, planes);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_upsample_bilinear2_bwd(const void* dy, float* dx, int32_t n, int32_t h_in, int32_t w_in, int32_t c,
                                          int32_t planes, void* stream) {
  V2X_REQUIRE(dy && dx && n > 0 && h_in > 0 && w_in > 0, "null/empty");
  V2X_CHECK_ACT(c, planes);
  cudaStream_t s = (cudaStream_t)stream;
  V2X_CUDA_TRY(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)n * h_in * w_in * c, s));
  upsample_bilinear2_bwd_kernel<<<grid_cap((long long)n * 4 * h_in * w_in * (c / 8), 256, 8), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), dx, n, h_in, w_in, c, planes);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}
