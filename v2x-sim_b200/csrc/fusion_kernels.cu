// Cross-agent fuse kernels of the intermediate-fusion baselines (CP/models/det/base/FusionBase.py:23-75 and the
// seg twins CP/models/seg/FusionBase.py:25-84): Mean / Max / Sum / Cat fusion, AgentWiseWeightedFusion and DiscoNet.
// All work in the UN-flipped domain with the same theta' as warp_mean_kernel (aux_kernels.cu); all are L2/HBM-bound
// gathers over 32x32xC maps: one warp per output pixel, lanes over channels in 16-byte vectors, fp32 accumulation.
#include "common.cuh"
#include "warp_staged.cuh"

namespace v2x {

static int fusion_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// sampling position in source j's map for target pixel (ow, oh) of agent i: affine_grid + grid_sample
// (align_corners=False) with theta' = [[T00, -T01, -T03/32], [-T10, T11, +T13/32]], T = trans[b][j][i]
// (DetModelBase.py:158-168 with the H flip folded in).
__device__ __forceinline__ void sample_pos(const double* __restrict__ trans, int b, int j, int i, int agents, int ow,
                                           int oh, int W, int H, float& ix, float& iy) {
  const double* T = trans + ((((long long)b * agents + j) * agents + i) << 4);
  const float t00 = (float)T[0], t01 = -(float)T[1], t02 = -(float)T[3] * (1.f / 32.f);
  const float t10 = -(float)T[4], t11 = (float)T[5], t12 = (float)T[7] * (1.f / 32.f);
  const float gx = (2.f * ow + 1.f) / W - 1.f, gy = (2.f * oh + 1.f) / H - 1.f;
  const float sx = t00 * gx + t01 * gy + t02, sy = t10 * gx + t11 * gy + t12;
  ix = ((sx + 1.f) * W - 1.f) * 0.5f;
  iy = ((sy + 1.f) * H - 1.f) * 0.5f;
}

// Bilinear gather (zeros padding) of one pixel of map `src_map` at (ix, iy): val[v][e] for this lane's channel
// vectors c0 = c_base + (v * 32 + lane) * 8 (< c_hi).  Returns false (val = 0) when the footprint misses the map.
template <int VEC>
__device__ __forceinline__ bool gather_bilinear(const __nv_bfloat16* __restrict__ x, long long plane_stride, int planes,
                                                long long src_map, float ix, float iy, int H, int W, int C, int c_base,
                                                int c_hi, int lane, float (&val)[VEC][8]) {
#pragma unroll
  for (int v = 0; v < VEC; ++v)
#pragma unroll
    for (int e = 0; e < 8; ++e) val[v][e] = 0.f;
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
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
  if (!any) return false;
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int c0 = c_base + (v * 32 + lane) * 8;
    if (c0 < c_hi) {
      uint4 q[4], ql[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        q[t] = inb[t] ? __ldg(reinterpret_cast<const uint4*>(sp[t] + c0)) : make_uint4(0, 0, 0, 0);
        ql[t] = (planes == 2 && inb[t]) ? __ldg(reinterpret_cast<const uint4*>(sp[t] + plane_stride + c0))
                                        : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t* h2 = reinterpret_cast<const uint32_t*>(&q[t]);
        const uint32_t* l2 = reinterpret_cast<const uint32_t*>(&ql[t]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = act_unpack2(h2[e], l2[e], planes);
          val[v][2 * e] += wgt[t] * f.x;
          val[v][2 * e + 1] += wgt[t] * f.y;
        }
      }
    }
  }
  return true;
}

// this lane's channel vectors of pixel `pix` of an act tensor (no interpolation)
template <int VEC>
__device__ __forceinline__ void load_pixel(const __nv_bfloat16* __restrict__ x, long long plane_stride, int planes,
                                           long long pix, int C, int c_base, int c_hi, int lane, float (&val)[VEC][8]) {
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int c0 = c_base + (v * 32 + lane) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) val[v][e] = 0.f;
    if (c0 < c_hi) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(x + pix * C + c0));
      const uint4 ql = planes == 2 ? __ldg(reinterpret_cast<const uint4*>(x + plane_stride + pix * C + c0))
                                   : make_uint4(0, 0, 0, 0);
      const uint32_t* h2 = reinterpret_cast<const uint32_t*>(&q);
      const uint32_t* l2 = reinterpret_cast<const uint32_t*>(&ql);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = act_unpack2(h2[e], l2[e], planes);
        val[v][2 * e] = f.x;
        val[v][2 * e + 1] = f.y;
      }
    }
  }
}

template <int VEC>
__device__ __forceinline__ void store_pixel(__nv_bfloat16* __restrict__ out, long long plane_stride, int planes,
                                            long long pix, int C, int lane, const float (&acc)[VEC][8]) {
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int c0 = (v * 32 + lane) * 8;
    if (c0 < C) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        act_pack2(acc[v][2 * e], acc[v][2 * e + 1], planes, hi[e], lo[e]);
      }
      *reinterpret_cast<uint4*>(out + pix * C + c0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      if (planes == 2)
        *reinterpret_cast<uint4*>(out + plane_stride + pix * C + c0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// FusionBase participation rule: the target itself plus every present agent j != i, restricted to pairs that
// involve agent 0 (the RSU) under only_v2i (DetModelBase.py:194-198).
__device__ __forceinline__ bool participates(int i, int j, int only_v2i) {
  return j == i || !(only_v2i && i != 0 && j != 0);
}

// ---------------------------------------------------------------------------------------------
// out[b,i] = reduce_{j in {i} U neighbours} warp_{j->i}(x[b,j]);  mode 0 mean, 1 sum, 2 max
// (MeanFusion.py:11-12, SumFusion.py:20-21, MaxFusion.py:20-21).  Agents i >= na[b] keep their own map
// (FusionBase.py:41-63 only rewrites the present agents).
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void warp_reduce_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                   const double* __restrict__ trans, const long long* __restrict__ num_agent, int batch,
                                   int agents, int H, int W, int C, int planes, int mode, int only_v2i) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long total_pix = (long long)batch * agents * H * W;
  const long long plane_stride = total_pix * C;
  for (long long wid = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); wid < total_pix;
       wid += (long long)gridDim.x * warps_per_block) {
    const int ow = (int)(wid % W);
    const int oh = (int)((wid / W) % H);
    const int map = (int)(wid / ((long long)W * H));
    const int i = map / batch, b = map % batch;
    const int na = (int)num_agent[(long long)b * agents];
    float acc[VEC][8], val[VEC][8];
    load_pixel<VEC>(x, plane_stride, planes, wid, C, 0, C, lane, acc);   // the self term (identity warp)
    if (i < na) {
      int count = 1;
      for (int j = 0; j < na; ++j) {
        if (j == i || !participates(i, j, only_v2i)) continue;
        ++count;
        float ix, iy;
        sample_pos(trans, b, j, i, agents, ow, oh, W, H, ix, iy);
        gather_bilinear<VEC>(x, plane_stride, planes, (long long)batch * j + b, ix, iy, H, W, C, 0, C, lane, val);
#pragma unroll
        for (int v = 0; v < VEC; ++v)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[v][e] = mode == 2 ? fmaxf(acc[v][e], val[v][e]) : acc[v][e] + val[v][e];
      }
      if (mode == 0) {
        const float inv = 1.f / (float)count;
#pragma unroll
        for (int v = 0; v < VEC; ++v)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[v][e] *= inv;
      }
    }
    store_pixel<VEC>(out, plane_stride, planes, wid, C, lane, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// Pair scores of the learned fusion weights (DiscoNet.py:132-155 PixelWeightedFusionSoftmax,
// AgentWiseWeightedFusion.py:44-76): for target i, list member k and pixel p
//   s[b,i,k,p] = relu(w4 . relu(W3 relu(W2 relu(BN1(conv1_1(cat[tg_i, nb_k])))) ...))
// conv1_1 is linear and 1x1, so it commutes with the bilinear warp: with q = conv1x1(x, [s*Wa ; s*Wb]) (one tensor-core
// launch, BN scale folded, bias in the first half) the first layer is relu(qa_i[p] + warp_{k->i}(qb_k)[p]).
// q: act [planes][A*B][H][W][2*HID], HID = 128.  Layers 2..4 (BN folded on the host, fp32) run here on CUDA cores:
// weights in shared memory, lane o owns hidden unit o of layer 2.
// scores: fp32 [B][A][A][H*W]; entries with k >= na[b] or non-participating pairs are not written.
// ---------------------------------------------------------------------------------------------
constexpr int kHid1 = 128, kHid2 = 32, kHid3 = 8;

__global__ void __launch_bounds__(256) pair_score_kernel(const __nv_bfloat16* __restrict__ q, float* __restrict__ scores,
                                                         const double* __restrict__ trans,
                                                         const long long* __restrict__ num_agent,
                                                         const float* __restrict__ w2, const float* __restrict__ b2,
                                                         const float* __restrict__ w3, const float* __restrict__ b3,
                                                         const float* __restrict__ w4, const float* __restrict__ b4,
                                                         int batch, int agents, int H, int W, int planes, int only_v2i) {
  __shared__ float s_w2t[kHid1 * kHid2];   // [c][o]
  __shared__ float s_w3[kHid3 * kHid2];    // [j][o]
  __shared__ float s_b2[kHid2], s_b3[kHid3], s_w4[kHid3];
  __shared__ __align__(16) float s_v[8][kHid1];
  for (int t = threadIdx.x; t < kHid1 * kHid2; t += blockDim.x) {
    const int o = t / kHid1, c = t % kHid1;
    s_w2t[c * kHid2 + o] = w2[t];
  }
  for (int t = threadIdx.x; t < kHid3 * kHid2; t += blockDim.x) s_w3[t] = w3[t];
  if (threadIdx.x < kHid2) s_b2[threadIdx.x] = b2[threadIdx.x];
  if (threadIdx.x < kHid3) {
    s_b3[threadIdx.x] = b3[threadIdx.x];
    s_w4[threadIdx.x] = w4[threadIdx.x];
  }
  __syncthreads();
  const float bias4 = b4[0];
  const int C = 2 * kHid1;
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int HW = H * W;
  const long long total_pix = (long long)batch * agents * HW;
  const long long plane_stride = total_pix * C;
  float* sv = s_v[wib];
  for (long long wid = (long long)blockIdx.x * warps_per_block + wib; wid < total_pix;
       wid += (long long)gridDim.x * warps_per_block) {
    const int p = (int)(wid % HW);
    const int ow = p % W, oh = p / W;
    const int map = (int)(wid / HW);
    const int i = map / batch, b = map % batch;
    const int na = (int)num_agent[(long long)b * agents];
    if (i >= na) continue;
    // lanes 0..15 hold 8 channels each of the 128-wide halves
    float qa[1][8], qb[1][8];
    load_pixel<1>(q, plane_stride, planes, wid, C, 0, kHid1, lane, qa);
    for (int k = 0; k < na; ++k) {
      if (!participates(i, k, only_v2i)) continue;
      if (k == i) {
        load_pixel<1>(q, plane_stride, planes, wid, C, kHid1, C, lane, qb);
      } else {
        float ix, iy;
        sample_pos(trans, b, k, i, agents, ow, oh, W, H, ix, iy);
        gather_bilinear<1>(q, plane_stride, planes, (long long)batch * k + b, ix, iy, H, W, C, kHid1, C, lane, qb);
      }
      __syncwarp();
      if (lane < kHid1 / 8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) sv[lane * 8 + e] = fmaxf(qa[0][e] + qb[0][e], 0.f);
      }
      __syncwarp();
      float h2 = s_b2[lane];
#pragma unroll 8
      for (int c = 0; c < kHid1; c += 4) {
        const float4 v4 = *reinterpret_cast<const float4*>(sv + c);
        h2 = fmaf(s_w2t[(c + 0) * kHid2 + lane], v4.x, h2);
        h2 = fmaf(s_w2t[(c + 1) * kHid2 + lane], v4.y, h2);
        h2 = fmaf(s_w2t[(c + 2) * kHid2 + lane], v4.z, h2);
        h2 = fmaf(s_w2t[(c + 3) * kHid2 + lane], v4.w, h2);
      }
      h2 = fmaxf(h2, 0.f);
      float s = bias4;
#pragma unroll
      for (int j = 0; j < kHid3; ++j) {
        float t = s_w3[j * kHid2 + lane] * h2;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        s = fmaf(s_w4[j], fmaxf(t + s_b3[j], 0.f), s);
      }
      s = fmaxf(s, 0.f);
      if (lane == 0) scores[(((long long)b * agents + i) * agents + k) * HW + p] = s;
    }
  }
}

// AgentWiseWeightedFusion.py:25-37,73-74: weight[b,i,k] = relu(b5 + sum_p w5f[p] * s[b,i,k,p]) (the 32x32 "valid" conv
// over the H-flipped score map: w5f is the filter with its rows mirrored), softmax over the na list members.
// One block per (b, i); coef [B][A][A] fp32 (zeros for k >= na).
__global__ void __launch_bounds__(256) agent_softmax_kernel(const float* __restrict__ scores, const float* __restrict__ w5f,
                                                            const float* __restrict__ b5,
                                                            const long long* __restrict__ num_agent,
                                                            float* __restrict__ coef, int batch, int agents, int HW) {
  __shared__ float red[8];
  __shared__ float wgt[32];
  const int b = blockIdx.x / agents, i = blockIdx.x % agents;
  const int na = (int)num_agent[(long long)b * agents];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int k = 0; k < agents; ++k) {
    float acc = 0.f;
    if (k < na && i < na) {
      const float* s = scores + (((long long)b * agents + i) * agents + k) * HW;
      for (int p = threadIdx.x; p < HW; p += blockDim.x) acc = fmaf(w5f[p], s[p], acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) red[wib] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = b5[0];
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
      wgt[k] = fmaxf(t, 0.f);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float* c = coef + ((long long)b * agents + i) * agents;
    if (i < na) {
      float m = -INFINITY, sum = 0.f;
      for (int k = 0; k < na; ++k) m = fmaxf(m, wgt[k]);
      for (int k = 0; k < na; ++k) sum += expf(wgt[k] - m);
      for (int k = 0; k < agents; ++k) c[k] = k < na ? expf(wgt[k] - m) / sum : 0.f;
    } else {
      for (int k = 0; k < agents; ++k) c[k] = 0.f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// out[b,i,p] = sum_k c[b,i,k,(p)] * warp_{k->i}(x[b,k])[p]  (k == i: identity)
// coef_mode 0: c = coef[b][i][k] (AgentWiseWeightedFusion.py:27-34)
// coef_mode 1: c = softmax over participating k of scores[b][i][k][p] (DiscoNet.py:88-107: exp / sum of exp)
// Agents i >= na[b] keep their own map.
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void warp_weighted_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                     const double* __restrict__ trans, const long long* __restrict__ num_agent,
                                     const float* __restrict__ coef, int coef_mode, int batch, int agents, int H, int W,
                                     int C, int planes, int only_v2i) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int HW = H * W;
  const long long total_pix = (long long)batch * agents * HW;
  const long long plane_stride = total_pix * C;
  for (long long wid = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); wid < total_pix;
       wid += (long long)gridDim.x * warps_per_block) {
    const int p = (int)(wid % HW);
    const int ow = p % W, oh = p / W;
    const int map = (int)(wid / HW);
    const int i = map / batch, b = map % batch;
    const int na = (int)num_agent[(long long)b * agents];
    float acc[VEC][8], val[VEC][8];
    if (i >= na) {
      load_pixel<VEC>(x, plane_stride, planes, wid, C, 0, C, lane, acc);
      store_pixel<VEC>(out, plane_stride, planes, wid, C, lane, acc);
      continue;
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[v][e] = 0.f;
    const float* cb = coef + ((long long)b * agents + i) * agents * (coef_mode == 1 ? HW : 1);
    float m = 0.f, inv = 1.f;
    if (coef_mode == 1) {
      m = -INFINITY;
      for (int k = 0; k < na; ++k)
        if (participates(i, k, only_v2i)) m = fmaxf(m, cb[(long long)k * HW + p]);
      float sum = 0.f;
      for (int k = 0; k < na; ++k)
        if (participates(i, k, only_v2i)) sum += __expf(cb[(long long)k * HW + p] - m);
      inv = 1.f / sum;
    }
    for (int k = 0; k < na; ++k) {
      if (!participates(i, k, only_v2i)) continue;
      const float c = coef_mode == 1 ? __expf(cb[(long long)k * HW + p] - m) * inv : cb[k];
      if (k == i) {
        load_pixel<VEC>(x, plane_stride, planes, wid, C, 0, C, lane, val);
      } else {
        float ix, iy;
        sample_pos(trans, b, k, i, agents, ow, oh, W, H, ix, iy);
        if (!gather_bilinear<VEC>(x, plane_stride, planes, (long long)batch * k + b, ix, iy, H, W, C, 0, C, lane, val))
          continue;
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[v][e] = fmaf(c, val[v][e], acc[v][e]);
    }
    store_pixel<VEC>(out, plane_stride, planes, wid, C, lane, acc);
  }
}

// out[unit] = x[unit] for the agent slots that are absent in their scene (i >= na[b]); used after a fuse stage that
// rewrote every map (CatFusion's modulation conv) to restore FusionBase's "only present agents are updated".
__global__ void restore_absent_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                      const long long* __restrict__ num_agent, int batch, int agents, long long map_elems,
                                      int planes) {
  const long long vecs = map_elems / 8;
  const long long total = (long long)batch * agents * vecs;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
       gid += (long long)gridDim.x * blockDim.x) {
    const int map = (int)(gid / vecs);
    const int i = map / batch, b = map % batch;
    if (i < (int)num_agent[(long long)b * agents]) continue;
    for (int pl = 0; pl < planes; ++pl) {
      const long long off = (long long)pl * batch * agents * map_elems + gid * 8;
      *reinterpret_cast<uint4*>(out + off) = __ldg(reinterpret_cast<const uint4*>(x + off));
    }
  }
}

}  // namespace v2x

using namespace v2x;

static unsigned fusion_grid(long long work_items, int threads, int per_sm) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)fusion_sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

#define V2X_FUSION_COMMON_CHECKS()                                                            \
  V2X_REQUIRE(batch > 0 && agents > 0 && h > 0 && w > 0, "empty geometry");                   \
  V2X_REQUIRE(planes == 1 || planes == 2, "planes must be 1 or 2")

extern "C" int v2x_warp_reduce_fwd(const void* x, void* out, const double* trans, const int64_t* num_agent,
                                   int32_t batch, int32_t agents, int32_t h, int32_t w, int32_t c, int32_t planes,
                                   int32_t mode, int32_t only_v2i, void* stream) {
  V2X_REQUIRE(x && out && trans && num_agent, "null pointer");
  V2X_FUSION_COMMON_CHECKS();
  V2X_REQUIRE(c > 0 && c % 8 == 0 && c <= 1024, "channels must be a multiple of 8, <= 1024");
  V2X_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0 (mean), 1 (sum) or 2 (max)");
  const long long total_pix = (long long)batch * agents * h * w;
  const unsigned grid = fusion_grid(total_pix * 32, 256, 8);
  const __nv_bfloat16* xi = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* xo = reinterpret_cast<__nv_bfloat16*>(out);
  const long long* na = reinterpret_cast<const long long*>(num_agent);
  cudaStream_t s = (cudaStream_t)stream;
  if (v2x::warp_staged_ok(c)) {   // source footprints staged in shared memory (warp_staged.cuh)
    v2x::WarpFuseArgs a{xi, xo, trans, na, nullptr, batch, agents, h, w, c, mode, 1, only_v2i,
                        0, batch * agents, 0, batch * agents};
    V2X_CUDA_TRY(v2x::launch_warp_fuse_staged<v2x::WF_REDUCE>(a, planes, s));
    return V2X_OK;
  }
  if (c <= 256)
    warp_reduce_kernel<1><<<grid, 256, 0, s>>>(xi, xo, trans, na, batch, agents, h, w, c, planes, mode, only_v2i);
  else if (c <= 512)
    warp_reduce_kernel<2><<<grid, 256, 0, s>>>(xi, xo, trans, na, batch, agents, h, w, c, planes, mode, only_v2i);
  else
    warp_reduce_kernel<4><<<grid, 256, 0, s>>>(xi, xo, trans, na, batch, agents, h, w, c, planes, mode, only_v2i);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_pair_score_fwd(const void* q, float* scores, const double* trans, const int64_t* num_agent,
                                  const float* w2, const float* b2, const float* w3, const float* b3, const float* w4,
                                  const float* b4, int32_t batch, int32_t agents, int32_t h, int32_t w, int32_t planes,
                                  int32_t only_v2i, void* stream) {
  V2X_REQUIRE(q && scores && trans && num_agent && w2 && b2 && w3 && b3 && w4 && b4, "null pointer");
  V2X_FUSION_COMMON_CHECKS();
  const long long total_pix = (long long)batch * agents * h * w;
  const unsigned grid = fusion_grid(total_pix * 32, 256, 4);
  pair_score_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), scores, trans, reinterpret_cast<const long long*>(num_agent), w2, b2,
      w3, b3, w4, b4, batch, agents, h, w, planes, only_v2i);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_agent_softmax_fwd(const float* scores, const float* w5f, const float* b5, const int64_t* num_agent,
                                     float* coef, int32_t batch, int32_t agents, int32_t hw, void* stream) {
  V2X_REQUIRE(scores && w5f && b5 && num_agent && coef, "null pointer");
  V2X_REQUIRE(batch > 0 && agents > 0 && agents <= 32 && hw > 0, "bad geometry (agents <= 32)");
  agent_softmax_kernel<<<batch * agents, 256, 0, (cudaStream_t)stream>>>(
      scores, w5f, b5, reinterpret_cast<const long long*>(num_agent), coef, batch, agents, hw);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_warp_weighted_fwd(const void* x, void* out, const double* trans, const int64_t* num_agent,
                                     const float* coef, int32_t coef_mode, int32_t batch, int32_t agents, int32_t h,
                                     int32_t w, int32_t c, int32_t planes, int32_t only_v2i, void* stream) {
  V2X_REQUIRE(x && out && trans && num_agent && coef, "null pointer");
  V2X_FUSION_COMMON_CHECKS();
  V2X_REQUIRE(c > 0 && c % 8 == 0 && c <= 1024, "channels must be a multiple of 8, <= 1024");
  V2X_REQUIRE(coef_mode == 0 || coef_mode == 1, "coef_mode must be 0 (per pair) or 1 (per-pixel softmax of scores)");
  const long long total_pix = (long long)batch * agents * h * w;
  const unsigned grid = fusion_grid(total_pix * 32, 256, 8);
  const __nv_bfloat16* xi = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* xo = reinterpret_cast<__nv_bfloat16*>(out);
  const long long* na = reinterpret_cast<const long long*>(num_agent);
  cudaStream_t s = (cudaStream_t)stream;
  if (v2x::warp_staged_ok(c)) {   // source footprints staged in shared memory (warp_staged.cuh)
    v2x::WarpFuseArgs a{xi, xo, trans, na, coef, batch, agents, h, w, c, coef_mode, 1, only_v2i,
                        0, batch * agents, 0, batch * agents};
    V2X_CUDA_TRY(v2x::launch_warp_fuse_staged<v2x::WF_WEIGHTED>(a, planes, s));
    return V2X_OK;
  }
  if (c <= 256)
    warp_weighted_kernel<1><<<grid, 256, 0, s>>>(xi, xo, trans, na, coef, coef_mode, batch, agents, h, w, c, planes, only_v2i);
  else if (c <= 512)
    warp_weighted_kernel<2><<<grid, 256, 0, s>>>(xi, xo, trans, na, coef, coef_mode, batch, agents, h, w, c, planes, only_v2i);
  else
    warp_weighted_kernel<4><<<grid, 256, 0, s>>>(xi, xo, trans, na, coef, coef_mode, batch, agents, h, w, c, planes, only_v2i);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_restore_absent_fwd(const void* x, void* out, const int64_t* num_agent, int32_t batch, int32_t agents,
                                      int64_t map_elems, int32_t planes, void* stream) {
  V2X_REQUIRE(x && out && num_agent, "null pointer");
  V2X_REQUIRE(batch > 0 && agents > 0 && map_elems > 0 && map_elems % 8 == 0, "map_elems must be a multiple of 8");
  V2X_REQUIRE(planes == 1 || planes == 2, "planes must be 1 or 2");
  const long long total = (long long)batch * agents * (map_elems / 8);
  restore_absent_kernel<<<fusion_grid(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(out),
      reinterpret_cast<const long long*>(num_agent), batch, agents, map_elems, planes);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}
