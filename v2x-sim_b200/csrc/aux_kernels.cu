// Memory-bound helpers of the hot path: input packing, weight packing (BN fold + bf16 hi/lo split),
// the cross-agent bilinear warp-and-mean, and act -> fp32 NCHW export.
// All are HBM/L2-bound byte movers: coalesced 16-byte accesses, grids sized in multiples of the SM
// count, no tensor cores.  Reference call sites: see include/v2x_b200.h.
#include "common.cuh"
#include "warp_staged.cuh"

namespace v2x {

static int sm_count() {
  // per DEVICE (a process may drive several GPUs; cudaGetDevice follows the caller's torch.cuda.device guard)
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  const int slot = dev & 63;
  if (n[slot] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[slot] = v > 0 ? v : 148;
  }
  return n[slot];
}

// ---------------------------------------------------------------------------------------------
// fp32 NHWC [n_pixels][c] -> bf16 planes [planes][n_pixels][c_pad]; one thread per (pixel, 8-channel group)
// ---------------------------------------------------------------------------------------------
__global__ void pack_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long n_pixels,
                                  int c, int c_pad, int planes) {
  const int groups = c_pad / 8;
  const long long total = n_pixels * groups;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
       gid += (long long)gridDim.x * blockDim.x) {
    const long long pix = gid / groups;
    const int g = (int)(gid - pix * groups);
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c0 = g * 8 + 2 * i;
      const float v0 = c0 < c ? __ldg(x + pix * c + c0) : 0.f;
      const float v1 = c0 + 1 < c ? __ldg(x + pix * c + c0 + 1) : 0.f;
      act_pack2(v0, v1, planes, hi[i], lo[i]);
    }
    __nv_bfloat16* dst = out + pix * c_pad + g * 8;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (planes == 2) *reinterpret_cast<uint4*>(dst + n_pixels * c_pad) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// 13-channel fp32 NHWC -> 16-channel bf16 planes.  A block takes kPackPix consecutive pixels: the 13*kPackPix floats are
// read as coalesced float4s into shared memory (the per-pixel stride of 52 bytes defeats direct vector loads), then
// every thread emits 16-byte output units (pixel, half) -- fully coalesced both ways.
constexpr int kPackPix = 512;
__global__ void __launch_bounds__(256) pack_input13_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                           long long n_pixels, int planes) {
  __shared__ __align__(16) float s[kPackPix * 13];
  const long long pix0 = (long long)blockIdx.x * kPackPix;
  const float4* src = reinterpret_cast<const float4*>(x + pix0 * 13);
#pragma unroll
  for (int k = 0; k < (kPackPix * 13 / 4 + 255) / 256; ++k) {
    const int i = threadIdx.x + 256 * k;
    if (i < kPackPix * 13 / 4) reinterpret_cast<float4*>(s)[i] = __ldg(src + i);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kPackPix * 2 / 256; ++k) {
    const int u = threadIdx.x + 256 * k;
    const int pix = u >> 1, g = u & 1;
    const float* sp = s + pix * 13 + g * 8;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c0 = g * 8 + 2 * i;
      const float v0 = c0 < 13 ? sp[2 * i] : 0.f;
      const float v1 = c0 + 1 < 13 ? sp[2 * i + 1] : 0.f;
      act_pack2(v0, v1, planes, hi[i], lo[i]);
    }
    __nv_bfloat16* dst = out + (pix0 + pix) * 16 + g * 8;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (planes == 2) *reinterpret_cast<uint4*>(dst + n_pixels * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// weight packing: one thread per (co, ci, tap)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int gru_perm(int row, int gates, int cout) {
  if (gates != 3) return row;
  const int C = cout / 3;
  const int g = row / C, c = row % C;
  return (c / 64) * 192 + g * 64 + (c % 64);
}

__global__ void pack_weights_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ mean, const float* __restrict__ var, float eps, int cout,
                                    int cin_total, int taps, int ci_lo, int ci_hi, int cin_pad, int vflip, int gates,
                                    uint16_t* __restrict__ dst, float* __restrict__ dst_bias, int planes,
                                    int cout_pad, int k_total, int row_off, int k_off, int write_bias) {
  const int cin = ci_hi - ci_lo;
  const long long total = (long long)cout * cin * taps;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
       gid += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(gid % taps);
    const int ci = (int)((gid / taps) % cin);
    const int co = (int)(gid / ((long long)taps * cin));
    const float s = gamma ? gamma[co] / sqrtf(var[co] + eps) : 1.f;  // BN(eval) scale, folded into w
    const float v = w[((long long)co * cin_total + ci_lo + ci) * taps + tap] * s;
    int tp = tap;
    if (vflip && taps == 9) tp = (2 - tap / 3) * 3 + tap % 3;
    const int row = row_off + gru_perm(co, gates, cout);
    const long long k = (long long)k_off + (long long)tp * cin_pad + ci;
    uint16_t hi, lo;
    act_split1(v, planes, hi, lo);   // planes = storage format: 1 bf16, 2 fp16 hi/lo, 3 fp16 hi only
    dst[(long long)row * k_total + k] = hi;
    if (planes == 2) dst[((long long)cout_pad + row) * k_total + k] = lo;
    if (write_bias && ci == 0 && tap == 0) {
      float bb = b ? b[co] : 0.f;
      if (gamma) bb = (bb - mean[co]) * s + beta[co];
      dst_bias[row] = bb;
    }
  }
}

__global__ void pack_gru_bias_kernel(const float* __restrict__ b_ih, const float* __restrict__ b_hh, int c,
                                     float* __restrict__ bias, float* __restrict__ bhn) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * c) return;
  const int g = i / c, ch = i % c;
  bias[(ch / 64) * 192 + g * 64 + (ch % 64)] = b_ih[i] + (g < 2 ? b_hh[i] : 0.f);
  if (g == 2) bhn[ch] = b_hh[i];
}

// ---------------------------------------------------------------------------------------------
// Cross-agent warp + mean.  One warp per output pixel; lanes stride over channels in 8-channel
// (16-byte) vectors so every tap of the bilinear gather is a run of fully coalesced 512-byte
// (C = 256) reads; the four taps x (A-1) sources are accumulated in registers and the mean is
// written once.  The maps are 32x32xC (0.5 MB each in bf16) and stay L2-resident.
// ---------------------------------------------------------------------------------------------
constexpr int kWarpMaxAgents = 8;   // sources whose sample positions are precomputed in registers

template <int VEC_PER_LANE, int PLANES>
__global__ void __launch_bounds__(256) warp_mean_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                        const double* __restrict__ trans,
                                                        const long long* __restrict__ num_agent, int batch, int agents,
                                                        int H, int W, int C, int include_self, int only_v2i,
                                                        int unit_offset, int unit_count) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long total_pix = (long long)unit_count * H * W;                 // targets computed by this launch
  const long long plane_stride = (long long)batch * agents * H * W * C;      // of the (complete) source tensor
  const long long out_plane_stride = total_pix * C;
  for (long long wid = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); wid < total_pix;
       wid += (long long)gridDim.x * warps_per_block) {
    const int ow = (int)(wid % W);
    const int oh = (int)((wid / W) % H);
    const int map = (int)(wid / ((long long)W * H)) + unit_offset;  // agent-major: map = batch * i + b
    const int i = map / batch, b = map % batch;
    const int na = (int)num_agent[(long long)b * agents];
    float acc[VEC_PER_LANE][8];
#pragma unroll
    for (int v = 0; v < VEC_PER_LANE; ++v)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[v][e] = 0.f;
    int count = 0;
    if (i < na) {
      const float gx = (2.f * ow + 1.f) / W - 1.f, gy = (2.f * oh + 1.f) / H - 1.f;
      // phase 1: sample positions of every participating source (independent loads of the 4x4 poses issue together,
      // so the pose fetch is one memory round trip per pixel instead of one per source)
      float sx_[kWarpMaxAgents], sy_[kWarpMaxAgents];
      unsigned use = 0;
#pragma unroll
      for (int j = 0; j < kWarpMaxAgents; ++j) {
        sx_[j] = 0.f; sy_[j] = 0.f;
        if (j >= na || j >= agents) continue;
        if (j == i && !include_self) continue;
        if (only_v2i && i != 0 && j != 0 && j != i) continue;  // DetModelBase.py:196-198
        use |= 1u << j;
        float t00, t01, t02, t10, t11, t12;
        if (j == i) {  // FusionBase-style self term: identity
          t00 = 1.f; t01 = 0.f; t02 = 0.f; t10 = 0.f; t11 = 1.f; t12 = 0.f;
        } else {
          const double* T = trans + ((((long long)b * agents + j) * agents + i) << 4);
          // theta' (un-flipped domain): [[T00, -T01, -T03/32], [-T10, T11, +T13/32]]
          t00 = (float)__ldg(T + 0); t01 = -(float)__ldg(T + 1); t02 = -(float)__ldg(T + 3) * (1.f / 32.f);
          t10 = -(float)__ldg(T + 4); t11 = (float)__ldg(T + 5); t12 = (float)__ldg(T + 7) * (1.f / 32.f);
        }
        const float sx = t00 * gx + t01 * gy + t02, sy = t10 * gx + t11 * gy + t12;
        sx_[j] = ((sx + 1.f) * W - 1.f) * 0.5f;
        sy_[j] = ((sy + 1.f) * H - 1.f) * 0.5f;
      }
      count = __popc(use);
      // sources beyond the precomputed window (agents > 8) are not supported by this kernel (checked on the host)
#pragma unroll
      for (int j = 0; j < kWarpMaxAgents; ++j) {
        if (!((use >> j) & 1u)) continue;
        const float ix = sx_[j], iy = sy_[j];
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = (int)fx, y0 = (int)fy;
        const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
        const long long src_map = (long long)batch * j + b;
        // all four taps' loads are issued before any is consumed (one memory round trip per source instead of four)
        float wgt[4];
        bool inb[4];
        const __nv_bfloat16* sp[4];
        bool any = false;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
          inb[t] = xx >= 0 && xx < W && yy >= 0 && yy < H;
          any |= inb[t];
          wgt[t] = ((t & 1) ? wx1 : wx0) * ((t >> 1) ? wy1 : wy0);
          sp[t] = x + ((src_map * H + yy) * W + xx) * C;
        }
        if (!any) continue;   // the whole footprint falls outside the source map (warp-uniform)
#pragma unroll
        for (int v = 0; v < VEC_PER_LANE; ++v) {
          const int c0 = (v * 32 + lane) * 8;
          if (c0 < C) {
            uint4 q[4], ql[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {   // predicated loads: out-of-range taps are zeros padding and cost no traffic
              q[t] = inb[t] ? __ldg(reinterpret_cast<const uint4*>(sp[t] + c0)) : make_uint4(0, 0, 0, 0);
              if (PLANES == 2)
                ql[t] = inb[t] ? __ldg(reinterpret_cast<const uint4*>(sp[t] + plane_stride + c0)) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const uint32_t* h2 = reinterpret_cast<const uint32_t*>(&q[t]);
              const uint32_t* l2 = reinterpret_cast<const uint32_t*>(&ql[t]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = act_unpack2<PLANES>(h2[e], PLANES == 2 ? l2[e] : 0u);
                acc[v][2 * e] += wgt[t] * f.x;
                acc[v][2 * e + 1] += wgt[t] * f.y;
              }
            }
          }
        }
      }
    }
    const float inv = count > 0 ? 1.f / (float)count : 0.f;
    __nv_bfloat16* dp = out + wid * C;
#pragma unroll
    for (int v = 0; v < VEC_PER_LANE; ++v) {
      const int c0 = (v * 32 + lane) * 8;
      if (c0 < C) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          act_pack2<PLANES>(acc[v][2 * e] * inv, acc[v][2 * e + 1] * inv, hi[e], lo[e]);
        }
        *reinterpret_cast<uint4*>(dp + c0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (PLANES == 2) *reinterpret_cast<uint4*>(dp + out_plane_stride + c0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  }
}

// act planes NHWC -> fp32 NCHW (debug / feature export); one thread per output element, w fastest
__global__ void act_to_nchw_kernel(const __nv_bfloat16* __restrict__ act, float* __restrict__ out, int n, int h, int w,
                                   int c, int planes) {
  const long long total = (long long)n * c * h * w;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
       gid += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(gid % w);
    const int y = (int)((gid / w) % h);
    const int ch = (int)((gid / ((long long)w * h)) % c);
    const int im = (int)(gid / ((long long)w * h * c));
    const long long src = (((long long)im * h + y) * w + x) * c + ch;
    out[gid] = act_load1(act + src, total, planes);
  }
}

}  // namespace v2x

using namespace v2x;

static unsigned grid_for(long long work_items, int threads, int per_sm) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

extern "C" int v2x_pack_input(const float* x, void* out, int64_t n_pixels, int32_t c, int32_t c_pad, int32_t planes,
                              void* stream) {
  V2X_REQUIRE(x && out && n_pixels > 0, "null/empty input");
  V2X_REQUIRE(c > 0 && c_pad >= c && c_pad % 8 == 0, "c_pad must be a multiple of 8 and >= c");
  V2X_REQUIRE(planes == 1 || planes == 2, "planes must be 1 or 2");
  const long long total = (long long)n_pixels * (c_pad / 8);
  if (c == 13 && c_pad == 16 && n_pixels % kPackPix == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    // the BEV input of the path (V2VNet.py:51): coalesced float4 reads staged through shared memory
    pack_input13_kernel<<<(unsigned)(n_pixels / kPackPix), 256, 0, (cudaStream_t)stream>>>(
        x, reinterpret_cast<__nv_bfloat16*>(out), n_pixels, planes);
  } else {
    pack_input_kernel<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        x, reinterpret_cast<__nv_bfloat16*>(out), n_pixels, c, c_pad, planes);
  }
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_pack_conv_weights(const float* w, const float* b, const float* bn_gamma, const float* bn_beta,
                                     const float* bn_mean, const float* bn_var, float eps, int32_t cout,
                                     int32_t cin_total, int32_t taps, int32_t ci_lo, int32_t ci_hi, int32_t cin_pad,
                                     int32_t vflip, int32_t gru_gates, void* dst, float* dst_bias, int32_t planes,
                                     int32_t cout_pad, int32_t k_total, int32_t row_off, int32_t k_off,
                                     int32_t write_bias, void* stream) {
  V2X_REQUIRE(w && dst, "null weights/dst");
  V2X_REQUIRE(cout > 0 && taps > 0 && 0 <= ci_lo && ci_lo < ci_hi && ci_hi <= cin_total, "bad channel range");
  V2X_REQUIRE(cin_pad >= ci_hi - ci_lo, "cin_pad too small");
  V2X_REQUIRE(planes >= 1 && planes <= 3, "format must be V2X_FMT_BF16 (1), V2X_FMT_F16X2 (2) or V2X_FMT_F16 (3)");
  V2X_REQUIRE(row_off >= 0 && row_off + cout <= cout_pad, "rows out of range");
  V2X_REQUIRE(k_off >= 0 && k_off + taps * cin_pad <= k_total, "k range out of bounds");
  V2X_REQUIRE(!bn_gamma || (bn_beta && bn_mean && bn_var), "incomplete BN parameters");
  V2X_REQUIRE(gru_gates == 0 || (gru_gates == 3 && cout % 192 == 0), "gru_gates needs cout %% 192 == 0");
  V2X_REQUIRE(!write_bias || dst_bias, "write_bias needs dst_bias");
  const long long total = (long long)cout * (ci_hi - ci_lo) * taps;
  pack_weights_kernel<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      w, b, bn_gamma, bn_beta, bn_mean, bn_var, eps, cout, cin_total, taps, ci_lo, ci_hi, cin_pad, vflip, gru_gates,
      reinterpret_cast<uint16_t*>(dst), dst_bias, planes, cout_pad, k_total, row_off, k_off, write_bias);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_pack_gru_bias(const float* b_ih, const float* b_hh, int32_t c, float* bias, float* bhn,
                                 void* stream) {
  V2X_REQUIRE(b_ih && b_hh && bias && bhn, "null pointer");
  V2X_REQUIRE(c > 0 && c % 64 == 0, "GRU channels must be a multiple of 64");
  pack_gru_bias_kernel<<<(3 * c + 255) / 256, 256, 0, (cudaStream_t)stream>>>(b_ih, b_hh, c, bias, bhn);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_warp_mean_fwd(const void* x, void* out, const double* trans, const int64_t* num_agent,
                                 int32_t batch, int32_t agents, int32_t h, int32_t w, int32_t c, int32_t planes,
                                 int32_t include_self, int32_t only_v2i, int32_t unit_offset, int32_t unit_count,
                                 void* stream) {
  V2X_REQUIRE(x && out && trans && num_agent, "null pointer");
  V2X_REQUIRE(batch > 0 && agents > 0 && h > 0 && w > 0, "empty geometry");
  V2X_REQUIRE(agents <= v2x::kWarpMaxAgents, "at most %d agents per scene", v2x::kWarpMaxAgents);
  V2X_REQUIRE(c > 0 && c % 8 == 0 && c <= 1024, "channels must be a multiple of 8, <= 1024");
  V2X_REQUIRE(planes == 1 || planes == 2, "planes must be 1 or 2");
  if (unit_count <= 0) { unit_offset = 0; unit_count = batch * agents; }
  V2X_REQUIRE(unit_offset >= 0 && unit_offset + unit_count <= batch * agents, "unit range out of bounds");
  const long long total_pix = (long long)unit_count * h * w;
  const int threads = 256;
  const unsigned grid = grid_for(total_pix * 32, threads, 8);
  const __nv_bfloat16* xi = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* xo = reinterpret_cast<__nv_bfloat16*>(out);
  const long long* na = reinterpret_cast<const long long*>(num_agent);
  cudaStream_t s = (cudaStream_t)stream;
  if (v2x::warp_staged_ok(c)) {   // source footprints staged in shared memory (warp_staged.cuh)
    v2x::WarpFuseArgs a{xi, xo, trans, na, nullptr, batch, agents, h, w, c, 0, include_self, only_v2i,
                        unit_offset, unit_count, 0, batch * agents};
    V2X_CUDA_TRY(v2x::launch_warp_fuse_staged<v2x::WF_MEAN>(a, planes, s));
    return V2X_OK;
  }
#define V2X_WARP_MEAN(VEC_, PL_)                                                                              \
  warp_mean_kernel<VEC_, PL_><<<grid, threads, 0, s>>>(xi, xo, trans, na, batch, agents, h, w, c, include_self, \
                                                       only_v2i, unit_offset, unit_count)
  if (c <= 256) { if (planes == 1) V2X_WARP_MEAN(1, 1); else V2X_WARP_MEAN(1, 2); }
  else if (c <= 512) { if (planes == 1) V2X_WARP_MEAN(2, 1); else V2X_WARP_MEAN(2, 2); }
  else { if (planes == 1) V2X_WARP_MEAN(4, 1); else V2X_WARP_MEAN(4, 2); }
#undef V2X_WARP_MEAN
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_act_to_nchw_f32(const void* act, float* out, int32_t n, int32_t h, int32_t w, int32_t c,
                                   int32_t planes, void* stream) {
  V2X_REQUIRE(act && out && n > 0 && h > 0 && w > 0 && c > 0, "null/empty");
  V2X_REQUIRE(planes == 1 || planes == 2, "planes must be 1 or 2");
  const long long total = (long long)n * c * h * w;
  act_to_nchw_kernel<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(act), out, n, h, w, c, planes);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

// =============================================================================================
// when2com / who2com helpers (CP/models/det/When2com.py): KmGenerator MLP, attention scores, gated fuse
// =============================================================================================
namespace v2x {

// y[r][o] = [relu](b[o] + sum_i w[o][i] * x[r][i]).  Block = (row r, group of kLinOut outputs): the input row is staged in
// shared memory ONCE (for in_mode 1 that is the strided NCHW-order gather out of the NHWC act tensor -- the expensive part:
// round 1 redid it for every output, 177 us for the 4096 -> 256 layer of the seg model), then each warp reduces its
// outputs against it with coalesced float4 weight reads.
// in_mode 1 reads x from an act tensor [planes][maps][hw][c] in NCHW-flatten order (i = ch * hw + px), which is
// how `features_map.view(-1, n_feat)` flattens the policy maps (When2com.py:429).
constexpr int kLinOut = 32;      // outputs per block (8 warps x 4)
__global__ void __launch_bounds__(256) linear_kernel(const void* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ b, float* __restrict__ y, int rows, int in_f,
                                                     int out_f, int relu, int in_mode, int hw, int c, int planes, int split,
                                                     int maps) {
  extern __shared__ __align__(16) float s_x[];     // [in_f]
  const int r = blockIdx.y, o0 = blockIdx.x * kLinOut;
  if (in_mode == 0) {
    const float* xr = reinterpret_cast<const float*>(x) + (long long)r * in_f;
    for (int i = threadIdx.x; i < in_f; i += blockDim.x) s_x[i] = __ldg(xr + i);
  } else {
    // virtual row r of `.view(-1, in_f)` over maps flattened in NCHW order: map r / split, channels
    // [(r % split) * c / split, ...) -- split > 1 reproduces the seg when2com quirk (When2Com_UNet.py:207-226, Q9).
    // Gather pixel-major (consecutive threads -> consecutive channels of one pixel: coalesced), scatter into NCHW order.
    const __nv_bfloat16* xa = reinterpret_cast<const __nv_bfloat16*>(x);
    const long long plane_stride = (long long)maps * hw * c;
    const int map = r / split, cs = c / split, ch_base = (r % split) * cs;
    for (int j = threadIdx.x; j < in_f; j += blockDim.x) {
      const int px = j / cs, ch = j - px * cs;
      s_x[ch * hw + px] = act_load1(xa + ((long long)map * hw + px) * c + ch_base + ch, plane_stride, planes);
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = 0; k < kLinOut / 8; ++k) {
    const int o = o0 + warp * (kLinOut / 8) + k;
    if (o >= out_f) break;
    const float* wr = w + (long long)o * in_f;
    float acc = 0.f;
    if ((in_f & 3) == 0) {
      for (int i = lane * 4; i < in_f; i += 128) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + i));
        const float4 xv = *reinterpret_cast<const float4*>(s_x + i);
        acc = fmaf(wv.x, xv.x, acc); acc = fmaf(wv.y, xv.y, acc); acc = fmaf(wv.z, xv.z, acc); acc = fmaf(wv.w, xv.w, acc);
      }
    } else {
      for (int i = lane; i < in_f; i += 32) acc = fmaf(__ldg(wr + i), s_x[i], acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      acc += b ? b[o] : 0.f;
      y[(long long)r * out_f + o] = relu ? fmaxf(acc, 0.f) : acc;
    }
  }
}

// MIMOGeneralDotProductAttention scores (When2com.py:374-412) + the eval-time gating (:259-327, :94-148).
// One block per scene.  keys [A*B][ks], querys [A*B][qs] are agent-major (row = B*i + b).
//   q' = W q + bw;  s[k][j] = key_k . q'_j;  attn[b][k][j] = softmax over k
//   coef (gate_mode): 0 = attn; 1 ("activated") = p * (p > 0.2), p = attn + 0.001 I; 2 ("argmax_test") = one-hot over k of argmax p
__global__ void attn_scores_kernel(const float* __restrict__ keys, const float* __restrict__ querys,
                                   const float* __restrict__ w, const float* __restrict__ bw, float* __restrict__ attn,
                                   float* __restrict__ coef, int batch, int agents, int ks, int qs, int gate_mode) {
  extern __shared__ float sm[];
  float* qp = sm;                    // [agents][ks]
  float* sc = sm + agents * ks;      // [agents][agents]
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  // q'[j][d] = bw[d] + sum_t W[d][t] * q_j[t]: one warp per feature d, lanes over t (coalesced reads of W's row d; round 1
  // walked W with a stride of qs floats per thread and took 100 us for 5 x 1024 features)
  for (int d = warp; d < ks; d += nw) {
    float wv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) wv[u] = (lane + 32 * u) < qs ? __ldg(w + (long long)d * qs + lane + 32 * u) : 0.f;
    for (int j = 0; j < agents; ++j) {
      const float* q = querys + ((long long)batch * j + b) * qs;
      float a = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (lane + 32 * u < qs) a = fmaf(wv[u], __ldg(q + lane + 32 * u), a);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
      if (lane == 0) qp[j * ks + d] = a + bw[d];
    }
  }
  __syncthreads();
  for (int pair = warp; pair < agents * agents; pair += nw) {
    const int k = pair / agents, j = pair - k * agents;
    const float* key = keys + ((long long)batch * k + b) * ks;
    float a = 0.f;
    for (int d = lane; d < ks; d += 32) a = fmaf(key[d], qp[j * ks + d], a);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    if (lane == 0) sc[pair] = a;
  }
  __syncthreads();
  if (threadIdx.x < agents) {
    const int j = threadIdx.x;
    float m = -INFINITY;
    for (int k = 0; k < agents; ++k) m = fmaxf(m, sc[k * agents + j]);
    float sum = 0.f;
    for (int k = 0; k < agents; ++k) sum += expf(sc[k * agents + j] - m);
    int best = 0;
    float bestp = -1.f;
    for (int k = 0; k < agents; ++k) {
      const float pa = expf(sc[k * agents + j] - m) / sum;
      attn[((long long)b * agents + k) * agents + j] = pa;
      const float pp = pa + (k == j ? 0.001f : 0.f);
      if (pp > bestp) { bestp = pp; best = k; }
      if (gate_mode == 0) coef[((long long)b * agents + k) * agents + j] = pa;
      else if (gate_mode == 1) coef[((long long)b * agents + k) * agents + j] = pp > 0.2f ? pp : 0.f;
    }
    if (gate_mode == 2)
      for (int k = 0; k < agents; ++k) coef[((long long)b * agents + k) * agents + j] = k == best ? 1.f : 0.f;
  }
}

// Gated cross-agent fuse of when2com, un-flipped domain (same theta' as warp_mean_kernel):
//   warp_flag 1: out[b,q] = sum_{k<na} coef[b,k,q] * (k == q ? x[b,q] : warp(x[b,q], T[b,q,k]))   (val_mat[b,k,q], Q8)
//                 agents q >= na produce zeros (their val_mat rows/cols are never filled, When2com.py:199-225)
//   warp_flag 0: out[b,q] = sum_{k<A}  coef[b,k,q] * x[b,k]                                        (all agent slots)
template <int VEC_PER_LANE>
__global__ void warp_gated_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                  const double* __restrict__ trans, const long long* __restrict__ num_agent,
                                  const float* __restrict__ coef, int batch, int agents, int H, int W, int C,
                                  int planes, int warp_flag, int only_v2i, int unit_offset, int unit_count,
                                  int x_unit_offset, int x_units) {
  // unit-sharded plans: this launch computes targets [unit_offset, unit_offset + unit_count) of the agent-major units
  // (output indexed locally) from an x tensor holding units [x_unit_offset, x_unit_offset + x_units)
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long total_pix = (long long)unit_count * H * W;
  const long long plane_stride = (long long)x_units * H * W * C;
  const long long out_plane_stride = total_pix * C;
  for (long long wid = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); wid < total_pix;
       wid += (long long)gridDim.x * warps_per_block) {
    const int ow = (int)(wid % W);
    const int oh = (int)((wid / W) % H);
    const int map = (int)(wid / ((long long)W * H)) + unit_offset;
    const int q = map / batch, b = map % batch;
    const int na = (int)num_agent[(long long)b * agents];
    float acc[VEC_PER_LANE][8];
#pragma unroll
    for (int v = 0; v < VEC_PER_LANE; ++v)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[v][e] = 0.f;
    const int nterms = warp_flag ? (q < na ? na : 0) : agents;
    const float gx = (2.f * ow + 1.f) / W - 1.f, gy = (2.f * oh + 1.f) / H - 1.f;
    for (int k = 0; k < nterms; ++k) {
      const float cf = coef[((long long)b * agents + k) * agents + q];
      if (cf == 0.f) continue;
      if (warp_flag && only_v2i && k != q && k != 0 && q != 0) continue;
      const long long src_map = (warp_flag ? (long long)batch * q + b : (long long)batch * k + b) - x_unit_offset;
      float wts[4];
      int xs[4], ys[4];
      int ntap = 4;
      if (!warp_flag || k == q) {
        ntap = 1; wts[0] = 1.f; xs[0] = ow; ys[0] = oh;
      } else {
        const double* T = trans + ((((long long)b * agents + q) * agents + k) << 4);  // q's map into k's frame
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
      for (int t = 0; t < ntap; ++t) {
        if (xs[t] < 0 || xs[t] >= W || ys[t] < 0 || ys[t] >= H) continue;
        const float wgt = wts[t] * cf;
        const __nv_bfloat16* sp = x + ((src_map * H + ys[t]) * W + xs[t]) * C;
#pragma unroll
        for (int v = 0; v < VEC_PER_LANE; ++v) {
          const int c0 = (v * 32 + lane) * 8;
          if (c0 < C) {
            uint4 qv = __ldg(reinterpret_cast<const uint4*>(sp + c0));
            const uint32_t* h2 = reinterpret_cast<const uint32_t*>(&qv);
            uint4 ql = make_uint4(0, 0, 0, 0);
            if (planes == 2) ql = __ldg(reinterpret_cast<const uint4*>(sp + plane_stride + c0));
            const uint32_t* l2 = reinterpret_cast<const uint32_t*>(&ql);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = act_unpack2(h2[e], l2[e], planes);
              acc[v][2 * e] += wgt * f.x;
              acc[v][2 * e + 1] += wgt * f.y;
            }
          }
        }
      }
    }
    __nv_bfloat16* dp = out + wid * C;
#pragma unroll
    for (int v = 0; v < VEC_PER_LANE; ++v) {
      const int c0 = (v * 32 + lane) * 8;
      if (c0 < C) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          act_pack2(acc[v][2 * e], acc[v][2 * e + 1], planes, hi[e], lo[e]);
        }
        *reinterpret_cast<uint4*>(dp + c0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (planes == 2) *reinterpret_cast<uint4*>(dp + out_plane_stride + c0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  }
}

}  // namespace v2x

extern "C" int v2x_linear_fwd(const void* x, const float* w, const float* b, float* y, int32_t rows, int32_t in_f,
                              int32_t out_f, int32_t relu, int32_t in_mode, int32_t hw, int32_t c, int32_t planes,
                              int32_t split, int32_t maps, void* stream) {
  V2X_REQUIRE(x && w && y && rows > 0 && in_f > 0 && out_f > 0, "null/empty");
  if (in_mode == 1 && split <= 0) split = 1;
  V2X_REQUIRE(in_mode == 0 || (in_mode == 1 && hw > 0 && c > 0 && c % split == 0 && hw * (c / split) == in_f &&
                               maps > 0 && rows <= maps * split && (planes == 1 || planes == 2)),
              "bad act input geometry");
  V2X_REQUIRE((size_t)in_f * sizeof(float) <= 48 * 1024, "in_f too large for the staged linear kernel (<= 12288)");
  const dim3 grid((unsigned)((out_f + v2x::kLinOut - 1) / v2x::kLinOut), (unsigned)rows);
  linear_kernel<<<grid, 256, (size_t)in_f * sizeof(float), (cudaStream_t)stream>>>(x, w, b, y, rows, in_f, out_f, relu, in_mode,
                                                                                hw, c, planes, split, maps);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_attn_scores_fwd(const float* keys, const float* querys, const float* w, const float* bw,
                                   float* attn, float* coef, int32_t batch, int32_t agents, int32_t key_size,
                                   int32_t query_size, int32_t gate_mode, void* stream) {
  V2X_REQUIRE(keys && querys && w && bw && attn && coef, "null pointer");
  V2X_REQUIRE(batch > 0 && agents > 0 && agents <= 8 && key_size > 0 && query_size > 0 && query_size <= 128, "bad sizes");
  V2X_REQUIRE(gate_mode >= 0 && gate_mode <= 2, "gate_mode must be 0 (softmax), 1 (activated) or 2 (argmax)");
  const size_t smem = ((size_t)agents * key_size + agents * agents) * sizeof(float);
  V2X_REQUIRE(smem <= 48 * 1024, "key_size too large for the score kernel");
  attn_scores_kernel<<<batch, 1024, smem, (cudaStream_t)stream>>>(keys, querys, w, bw, attn, coef, batch, agents, key_size,
                                                              query_size, gate_mode);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_warp_gated_fwd(const void* x, void* out, const double* trans, const int64_t* num_agent,
                                  const float* coef, int32_t batch, int32_t agents, int32_t h, int32_t w, int32_t c,
                                  int32_t planes, int32_t warp_flag, int32_t only_v2i, int32_t unit_offset,
                                  int32_t unit_count, int32_t x_unit_offset, int32_t x_units, void* stream) {
  V2X_REQUIRE(x && out && trans && num_agent && coef, "null pointer");
  V2X_REQUIRE(batch > 0 && agents > 0 && h > 0 && w > 0, "empty geometry");
  V2X_REQUIRE(c > 0 && c % 8 == 0 && c <= 1024, "channels must be a multiple of 8, <= 1024");
  V2X_REQUIRE(planes == 1 || planes == 2, "planes must be 1 or 2");
  if (unit_count <= 0) { unit_offset = 0; unit_count = batch * agents; }
  if (x_units <= 0) { x_unit_offset = 0; x_units = batch * agents; }
  V2X_REQUIRE(unit_offset >= 0 && unit_offset + unit_count <= batch * agents, "target unit range out of bounds");
  V2X_REQUIRE(x_unit_offset >= 0 && x_unit_offset + x_units <= batch * agents, "source unit range out of bounds");
  // warp_flag 1 reads only the target's own map (val_mat[b,k,q] = q's map warped into k's frame); warp_flag 0
  // reads every agent of the scene
  V2X_REQUIRE(warp_flag ? (x_unit_offset <= unit_offset && unit_offset + unit_count <= x_unit_offset + x_units)
                        : (x_unit_offset == 0 && x_units == batch * agents),
              "x does not hold the units this launch reads");
  const long long total_pix = (long long)unit_count * h * w;
  const int threads = 256;
  const unsigned grid = grid_for(total_pix * 32, threads, 8);
  const __nv_bfloat16* xi = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* xo = reinterpret_cast<__nv_bfloat16*>(out);
  const long long* na = reinterpret_cast<const long long*>(num_agent);
  cudaStream_t s = (cudaStream_t)stream;
  if (warp_flag && v2x::warp_staged_ok(c)) {   // the warped variant: footprints staged in shared memory (warp_staged.cuh)
    v2x::WarpFuseArgs a{xi, xo, trans, na, coef, batch, agents, h, w, c, 0, 1, only_v2i,
                        unit_offset, unit_count, x_unit_offset, x_units};
    V2X_CUDA_TRY(v2x::launch_warp_fuse_staged<v2x::WF_GATED>(a, planes, s));
    return V2X_OK;
  }
  if (c <= 256)
    warp_gated_kernel<1><<<grid, threads, 0, s>>>(xi, xo, trans, na, coef, batch, agents, h, w, c, planes, warp_flag, only_v2i, unit_offset, unit_count,
                                                  x_unit_offset, x_units);
  else if (c <= 512)
    warp_gated_kernel<2><<<grid, threads, 0, s>>>(xi, xo, trans, na, coef, batch, agents, h, w, c, planes, warp_flag, only_v2i, unit_offset, unit_count,
                                                  x_unit_offset, x_units);
  else
    warp_gated_kernel<4><<<grid, threads, 0, s>>>(xi, xo, trans, na, coef, batch, agents, h, w, c, planes, warp_flag, only_v2i, unit_offset, unit_count,
                                                  x_unit_offset, x_units);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

// =============================================================================================
// segmentation UNet helpers (CP/models/seg/SegModelBase.py): NCHW input pack, MaxPool2d(2), bilinear x2 upsample
// All byte movers: one thread per (pixel, 8-channel group), 16-byte accesses.
// =============================================================================================
namespace v2x {

// fp32 NCHW [n][c][h][w] -> act bf16 planes NHWC [planes][n][h][w][c_pad]   (SegModule.py:49 hands the model NCHW)
__global__ void pack_input_nchw_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int n, int c,
                                       int h, int w, int c_pad, int planes) {
  const int groups = c_pad / 8;
  const long long hw = (long long)h * w;
  const long long total = (long long)n * hw * groups;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
       gid += (long long)gridDim.x * blockDim.x) {
    // pixel fastest so that the strided channel reads of a warp are coalesced along w
    const long long pix = gid % hw;
    const int g = (int)((gid / hw) % groups);
    const int im = (int)(gid / (hw * groups));
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c0 = g * 8 + 2 * i;
      const float v0 = c0 < c ? __ldg(x + ((long long)im * c + c0) * hw + pix) : 0.f;
      const float v1 = c0 + 1 < c ? __ldg(x + ((long long)im * c + c0 + 1) * hw + pix) : 0.f;
      act_pack2(v0, v1, planes, hi[i], lo[i]);
    }
    __nv_bfloat16* dst = out + ((long long)im * hw + pix) * c_pad + g * 8;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (planes == 2) *reinterpret_cast<uint4*>(dst + (long long)n * hw * c_pad) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// nn.MaxPool2d(2) (SegModelBase.py:113): act [n][2h][2w][c] -> [n][h][w][c]
__global__ void maxpool2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int n, int h,
                                int w, int c, int planes) {
  const int groups = c / 8;
  const long long total = (long long)n * h * w * groups;
  const long long in_plane = (long long)n * 4 * h * w * c, out_plane = (long long)n * h * w * c;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
       gid += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(gid % groups);
    const long long pix = gid / groups;
    const int ox = (int)(pix % w), oy = (int)((pix / w) % h), im = (int)(pix / ((long long)w * h));
    float m[8], v[8];
    const __nv_bfloat16* base = x + (((long long)im * 2 * h + 2 * oy) * 2 * w + 2 * ox) * c + g * 8;
    act_load8(base, in_plane, planes, m);
    act_load8(base + c, in_plane, planes, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], v[e]);
    act_load8(base + (long long)2 * w * c, in_plane, planes, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], v[e]);
    act_load8(base + (long long)2 * w * c + c, in_plane, planes, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], v[e]);
    act_store8(out + pix * c + g * 8, out_plane, planes, m);
  }
}

// nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True) (SegModelBase.py:125):
// src = dst * (in - 1) / (out - 1); act [n][h][w][c] -> [n][2h][2w][c]
__global__ void upsample_bilinear2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int n,
                                          int h, int w, int c, int planes) {
  const int groups = c / 8;
  const int oh_n = 2 * h, ow_n = 2 * w;
  const long long in_plane = (long long)n * h * w * c, out_plane = (long long)n * oh_n * ow_n * c;
  const float sy = oh_n > 1 ? (float)(h - 1) / (float)(oh_n - 1) : 0.f;
  const float sx = ow_n > 1 ? (float)(w - 1) / (float)(ow_n - 1) : 0.f;
  // 32-bit index arithmetic (the 64-bit divisions of a flat index dominated round 1's version: 1.8 TB/s): blockIdx.y walks
  // the output rows (image * oh_n + oy), the x dimension of the grid the (pixel, channel group) items of one row
  const unsigned row_items = (unsigned)ow_n * (unsigned)groups;
  for (unsigned row = blockIdx.y; row < (unsigned)n * (unsigned)oh_n; row += gridDim.y)
  for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < row_items; it += gridDim.x * blockDim.x) {
    const int g = (int)(it % (unsigned)groups);
    const int ox = (int)(it / (unsigned)groups);
    const int oy = (int)(row % (unsigned)oh_n), im = (int)(row / (unsigned)oh_n);
    const long long pix = ((long long)im * oh_n + oy) * ow_n + ox;
    const float fy = oy * sy, fx = ox * sx;
    const int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float wy1 = fy - (float)y0, wx1 = fx - (float)x0, wy0 = 1.f - wy1, wx0 = 1.f - wx1;
    const __nv_bfloat16* b = x + (long long)im * h * w * c + g * 8;
    float a00[8], a01[8], a10[8], a11[8], r[8];
    act_load8(b + ((long long)y0 * w + x0) * c, in_plane, planes, a00);
    act_load8(b + ((long long)y0 * w + x1) * c, in_plane, planes, a01);
    act_load8(b + ((long long)y1 * w + x0) * c, in_plane, planes, a10);
    act_load8(b + ((long long)y1 * w + x1) * c, in_plane, planes, a11);
#pragma unroll
    for (int e = 0; e < 8; ++e)  // same association as ATen's upsample_bilinear2d: rows first, then columns
      r[e] = wy0 * (wx0 * a00[e] + wx1 * a01[e]) + wy1 * (wx0 * a10[e] + wx1 * a11[e]);
    act_store8(out + pix * c + g * 8, out_plane, planes, r);
  }
}

}  // namespace v2x

extern "C" int v2x_pack_input_nchw(const float* x, void* out, int32_t n, int32_t c, int32_t h, int32_t w, int32_t c_pad,
                                   int32_t planes, void* stream) {
  V2X_REQUIRE(x && out && n > 0 && c > 0 && h > 0 && w > 0, "null/empty input");
  V2X_REQUIRE(c_pad >= c && c_pad % 8 == 0 && (planes == 1 || planes == 2), "bad c_pad / planes");
  const long long total = (long long)n * h * w * (c_pad / 8);
  pack_input_nchw_kernel<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      x, reinterpret_cast<__nv_bfloat16*>(out), n, c, h, w, c_pad, planes);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_maxpool2_fwd(const void* x, void* out, int32_t n, int32_t h_out, int32_t w_out, int32_t c,
                                int32_t planes, void* stream) {
  V2X_REQUIRE(x && out && n > 0 && h_out > 0 && w_out > 0 && c > 0 && c % 8 == 0, "bad geometry");
  V2X_REQUIRE(planes == 1 || planes == 2, "planes must be 1 or 2");
  const long long total = (long long)n * h_out * w_out * (c / 8);
  maxpool2_kernel<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(out), n, h_out, w_out, c, planes);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_upsample_bilinear2_fwd(const void* x, void* out, int32_t n, int32_t h_in, int32_t w_in, int32_t c,
                                          int32_t planes, void* stream) {
  V2X_REQUIRE(x && out && n > 0 && h_in > 0 && w_in > 0 && c > 0 && c % 8 == 0, "bad geometry");
  V2X_REQUIRE(planes == 1 || planes == 2, "planes must be 1 or 2");
  const unsigned row_items = 2u * (unsigned)w_in * (unsigned)(c / 8);
  unsigned rows = (unsigned)n * 2u * (unsigned)h_in;
  if (rows > 65535u) rows = 65535u;
  const dim3 grid((row_items + 255u) / 256u, rows);
  upsample_bilinear2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(out), n, h_in, w_in, c, planes);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}
