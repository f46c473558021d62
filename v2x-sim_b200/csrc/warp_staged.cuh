// warp_staged.cuh -- the cross-agent warp + fuse step with SHARED-MEMORY STAGING of the source maps.
// Status: correct for every fuse rule (tests/test_gpu_ops.py::test_warp_staged_matches_direct_kernels) but measured
// slower than the direct gathers on B200, so it is opt-in (V2X_WARP_STAGED=1; see warp_staged_ok below for the numbers).
//
// Every fuse rule of the path is "for target agent i of scene b: combine, over member agents k, a bilinear resampling of
// a source map under the affine theta'(T[b][.][.])" (DetModelBase.py:139-169 feature_transformation + the model's own
// reduction): V2VNet's neighbour mean (V2VNet.py:85-98), Mean / Sum / Max fusion (FusionBase.py:41-63), when2com's gated
// sum (When2com.py:199-225, 374-412) and the learned per-pair / per-pixel weights of AgentWise / DiscoNet.  The direct
// kernels (aux_kernels.cu, fusion_kernels.cu: one warp per output pixel, four 16-byte gathers per tap and plane) read
// every source pixel about four times through L2 (each source pixel is a tap of ~4 neighbouring outputs): measured
// 8.2 TB/s of L2->SM traffic for V2VNet's 40-map step, two thirds of the LTS cap -- the kernel is L2-bandwidth bound,
// not HBM bound.
//
// Here a CTA owns an 8x8-pixel output tile x 64 channels of one target.  For each member term it computes the bounding box
// of the tile's footprint in the source map (theta' is affine: the box of the four corner samples, +1 for the bilinear
// taps, clipped to the map -- at most 13x13 px for a rigid transform), brings that box into shared memory ONCE with
// coalesced 16-byte cp.async (128 B per pixel and plane), and the 4 taps of all 64 outputs are then served from shared
// memory.  L2->SM bytes per (tile, term) drop from 64 x 4 pixel reads to the box area (~120 px on average over yaw): ~2x
// less.  Terms whose box does not fit (non-rigid matrices) and taps that fall outside the staged box (never, up to float
// rounding; kept for safety) take the direct global path, so the result does not depend on the staging decision.
// 8 lanes (16 B each) cover the 64-channel slice of a pixel and plane, so a warp works on 4 output pixels at a time and
// every shared-memory read is one conflict-free 128-byte wavefront.
#pragma once
#include <cstdlib>

#include "common.cuh"

namespace v2x {

enum { WF_MEAN = 0, WF_REDUCE = 1, WF_GATED = 2, WF_WEIGHTED = 3 };

struct WarpFuseArgs {
  const __nv_bfloat16* x;
  __nv_bfloat16* out;
  const double* trans;            // [B][A][A][4][4]
  const long long* num_agent;     // [B][A]
  const float* coef;              // WF_GATED: [B][A(k)][A(q)]; WF_WEIGHTED: [B][A][A] (mode 0) or scores [B][A][A][HW] (mode 1)
  int batch, agents, H, W, C;
  int mode;                       // WF_REDUCE: 0 mean, 1 sum, 2 max; WF_WEIGHTED: coef_mode
  int include_self, only_v2i;
  int unit_offset, unit_count;    // target units computed by this launch (agent-major), output indexed locally
  int x_unit_offset, x_units;     // units held by x
};

constexpr int kWsTile = 8;        // output tile: 8 x 8 pixels
constexpr int kWsSlice = 64;      // channels per CTA (8 lanes x 16 bytes per plane)
constexpr int kWsCap = 192;       // staged source pixels: 192 x 256 B = 48 KB with two planes (no opt-in needed)
constexpr int kWsThreads = 256;

__device__ __forceinline__ void ws_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ws_cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

struct WsTheta {
  float t00, t01, t02, t10, t11, t12;
};
// theta' of the un-flipped domain: [[T00, -T01, -T03/32], [-T10, T11, +T13/32]]
__device__ __forceinline__ WsTheta ws_theta(const double* __restrict__ T) {
  WsTheta t;
  t.t00 = (float)__ldg(T + 0); t.t01 = -(float)__ldg(T + 1); t.t02 = -(float)__ldg(T + 3) * (1.f / 32.f);
  t.t10 = -(float)__ldg(T + 4); t.t11 = (float)__ldg(T + 5); t.t12 = (float)__ldg(T + 7) * (1.f / 32.f);
  return t;
}
// affine_grid + grid_sample (align_corners=False) sample position of output pixel (ow, oh)
__device__ __forceinline__ void ws_sample(const WsTheta& t, int ow, int oh, int W, int H, float& ix, float& iy) {
  const float gx = (2.f * ow + 1.f) / W - 1.f, gy = (2.f * oh + 1.f) / H - 1.f;
  const float sx = t.t00 * gx + t.t01 * gy + t.t02, sy = t.t10 * gx + t.t11 * gy + t.t12;
  ix = ((sx + 1.f) * W - 1.f) * 0.5f;
  iy = ((sy + 1.f) * H - 1.f) * 0.5f;
}

template <int PLANES>
__device__ __forceinline__ void ws_unpack8(const uint4& q, const uint4& ql, float* f) {
  const uint32_t* h2 = reinterpret_cast<const uint32_t*>(&q);
  const uint32_t* l2 = reinterpret_cast<const uint32_t*>(&ql);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 v = act_unpack2<PLANES>(h2[e], PLANES == 2 ? l2[e] : 0u);
    f[2 * e] = v.x;
    f[2 * e + 1] = v.y;
  }
}

template <int KIND, int PLANES>
__global__ void __launch_bounds__(kWsThreads, 4) warp_fuse_staged_kernel(const WarpFuseArgs a) {
  extern __shared__ __align__(16) uint8_t ws_smem[];   // [kWsCap pixels][PLANES][64 channels]
  constexpr uint32_t PX_BYTES = 128u * PLANES;
  constexpr int UNITS_PER_PX = 8 * PLANES;             // 16-byte units of one staged pixel
  const uint32_t smem0 = smem_u32(ws_smem);

  const int tiles_w = (a.W + kWsTile - 1) / kWsTile, tiles_h = (a.H + kWsTile - 1) / kWsTile;
  const int slices = a.C / kWsSlice;
  int bid = blockIdx.x;
  const int slice = bid % slices; bid /= slices;
  const int tw = bid % tiles_w; bid /= tiles_w;
  const int th = bid % tiles_h; bid /= tiles_h;
  const int unit_local = bid;
  const int unit = unit_local + a.unit_offset;          // agent-major: unit = batch * i + b
  const int i = unit / a.batch, b = unit % a.batch;
  const int na = min((int)a.num_agent[(long long)b * a.agents], a.agents);
  const int c_base = slice * kWsSlice;
  const int tid = threadIdx.x, sub = tid & 7, grp = tid >> 3;
  const int oh0 = th * kWsTile, ow0 = tw * kWsTile;
  const long long HW = (long long)a.H * a.W;
  const long long x_plane = (long long)a.x_units * HW * a.C;
  const long long out_plane = (long long)a.unit_count * HW * a.C;
  const int ch = c_base + sub * 8;

  int pw[2], ph[2];
  bool pv[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int p = grp + 32 * r;
    ph[r] = oh0 + (p >> 3);
    pw[r] = ow0 + (p & 7);
    pv[r] = ph[r] < a.H && pw[r] < a.W;
  }
  float acc[2][8];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[r][e] = 0.f;

  // this thread's 8 channels of pixel (hh, ww) of unit `u_local` of x, straight from global memory
  auto load_direct = [&](long long u_local, int hh, int ww, float* f) {
    const __nv_bfloat16* sp = a.x + ((u_local * a.H + hh) * a.W + ww) * a.C + ch;
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(sp));
    uint4 ql = make_uint4(0, 0, 0, 0);
    if (PLANES == 2) ql = __ldg(reinterpret_cast<const uint4*>(sp + x_plane));
    ws_unpack8<PLANES>(q, ql, f);
  };
  auto store_out = [&](int r) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) act_pack2<PLANES>(acc[r][2 * e], acc[r][2 * e + 1], hi[e], lo[e]);
    __nv_bfloat16* dp = a.out + (((long long)unit_local * a.H + ph[r]) * a.W + pw[r]) * a.C + ch;
    *reinterpret_cast<uint4*>(dp) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (PLANES == 2) *reinterpret_cast<uint4*>(dp + out_plane) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  };
  const long long own_local = (long long)unit - a.x_unit_offset;

  if (i >= na) {
    // absent agent slot: the V2VNet mean and the when2com fuse give zeros (the consumer passes the own map through / the
    // reference never fills those val_mat rows); the FusionBase family keeps the own map (FusionBase.py:41-63)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (!pv[r]) continue;
      if (KIND == WF_REDUCE || KIND == WF_WEIGHTED) load_direct(own_local, ph[r], pw[r], acc[r]);
      store_out(r);
    }
    return;
  }

  int count = 0;
  if (KIND == WF_REDUCE) {   // the self term (identity warp) opens the reduction
    count = 1;
#pragma unroll
    for (int r = 0; r < 2; ++r)
      if (pv[r]) load_direct(own_local, ph[r], pw[r], acc[r]);
  }
  // DiscoNet: per-pixel softmax over the participating members of scores[b][i][k][p] (DiscoNet.py:88-107)
  float sm_m[2] = {0.f, 0.f}, sm_inv[2] = {1.f, 1.f};
  const float* cb = nullptr;
  if (KIND == WF_WEIGHTED) {
    cb = a.coef + ((long long)b * a.agents + i) * a.agents * (a.mode == 1 ? HW : 1);
    if (a.mode == 1) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (!pv[r]) continue;
        const long long p = (long long)ph[r] * a.W + pw[r];
        float m = -INFINITY, sum = 0.f;
        for (int k = 0; k < na; ++k)
          if (k == i || !(a.only_v2i && i != 0 && k != 0)) m = fmaxf(m, __ldg(cb + k * HW + p));
        for (int k = 0; k < na; ++k)
          if (k == i || !(a.only_v2i && i != 0 && k != 0)) sum += __expf(__ldg(cb + k * HW + p) - m);
        sm_m[r] = m;
        sm_inv[r] = 1.f / sum;
      }
    }
  }

  bool smem_in_use = false;   // CTA-uniform: a previous term's box may still be read by other warps
  for (int k = 0; k < na; ++k) {
    const bool ident = k == i;
    const bool v2i_skip = a.only_v2i && !ident && i != 0 && k != 0;   // DetModelBase.py:194-198
    float cf = 1.f;
    if (KIND == WF_MEAN) {
      if ((ident && !a.include_self) || v2i_skip) continue;
    } else if (KIND == WF_REDUCE) {
      if (ident || v2i_skip) continue;
    } else if (KIND == WF_GATED) {
      cf = __ldg(a.coef + ((long long)b * a.agents + k) * a.agents + i);
      if (cf == 0.f || v2i_skip) continue;
    } else {
      if (v2i_skip) continue;
      if (a.mode == 0) cf = __ldg(cb + k);
    }
    ++count;
    // source map and transform: when2com warps the TARGET's own map into member k's frame (val_mat[b,k,q] pairing,
    // SURVEY Q8); every other rule warps member k's map into the target's frame
    const long long src_local = KIND == WF_GATED ? own_local : (long long)a.batch * k + b - a.x_unit_offset;
    float wgt_px[2] = {cf, cf};
    if (KIND == WF_WEIGHTED && a.mode == 1) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
        wgt_px[r] = pv[r] ? __expf(__ldg(cb + k * HW + (long long)ph[r] * a.W + pw[r]) - sm_m[r]) * sm_inv[r] : 0.f;
    }
    if (ident) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (!pv[r]) continue;
        float f[8];
        load_direct(src_local, ph[r], pw[r], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[r][e] = fmaf(wgt_px[r], f[e], acc[r][e]);
      }
      continue;
    }
    const double* T = KIND == WF_GATED ? a.trans + ((((long long)b * a.agents + i) * a.agents + k) << 4)
                                       : a.trans + ((((long long)b * a.agents + k) * a.agents + i) << 4);
    const WsTheta th_ = ws_theta(T);
    // footprint of the tile in the source map (CTA-uniform)
    float ixmin = INFINITY, ixmax = -INFINITY, iymin = INFINITY, iymax = -INFINITY;
#pragma unroll
    for (int cnr = 0; cnr < 4; ++cnr) {
      float cx, cy;
      ws_sample(th_, ow0 + (cnr & 1) * (kWsTile - 1), oh0 + (cnr >> 1) * (kWsTile - 1), a.W, a.H, cx, cy);
      ixmin = fminf(ixmin, cx); ixmax = fmaxf(ixmax, cx);
      iymin = fminf(iymin, cy); iymax = fmaxf(iymax, cy);
    }
    const bool finite = ixmin > -1e6f && ixmax < 1e6f && iymin > -1e6f && iymax < 1e6f;   // false for NaN / inf poses
    int bx0 = 0, bx1 = -1, by0 = 0, by1 = -1, bw = 0;
    bool staged = false;
    if (finite) {
      bx0 = max(0, (int)floorf(ixmin - 0.01f));
      bx1 = min(a.W - 1, (int)floorf(ixmax + 0.01f) + 1);
      by0 = max(0, (int)floorf(iymin - 0.01f));
      by1 = min(a.H - 1, (int)floorf(iymax + 0.01f) + 1);
      if (bx0 > bx1 || by0 > by1) {
        // the whole tile samples outside the source map: zeros padding (a zero still takes part in a max)
        if (KIND == WF_REDUCE && a.mode == 2) {
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[r][e] = fmaxf(acc[r][e], 0.f);
        }
        continue;
      }
      bw = bx1 - bx0 + 1;
      staged = bw * (by1 - by0 + 1) <= kWsCap;
    }
    if (staged) {
      if (smem_in_use) __syncthreads();
      // a warp copies box rows warp, warp + 8, ...; 8 * PLANES lanes cover one pixel (128 B per plane), so a warp moves
      // 32 / (8 * PLANES) adjacent pixels per instruction: no divisions, one 64-bit add per copy
      constexpr int PX_PER_WARP = 32 / UNITS_PER_PX;
      const int lane = tid & 31, wrp = tid >> 5;
      const int lpx = lane / UNITS_PER_PX, lrr = lane % UNITS_PER_PX;
      const int pl = lrr >> 3, cu = lrr & 7;
      const __nv_bfloat16* src0 = a.x + pl * x_plane + ((src_local * a.H + by0) * a.W + bx0) * a.C + c_base + cu * 8;
      const uint32_t dst0 = smem0 + (uint32_t)pl * 128u + (uint32_t)cu * 16u;
      for (int by = wrp; by <= by1 - by0; by += kWsThreads / 32) {
        const __nv_bfloat16* srow = src0 + (long long)by * a.W * a.C;
        const uint32_t drow = dst0 + (uint32_t)(by * bw) * PX_BYTES;
        for (int bx = lpx; bx < bw; bx += PX_PER_WARP) ws_cp_async16(drow + (uint32_t)bx * PX_BYTES, srow + (long long)bx * a.C);
      }
      ws_cp_async_wait_all();
      __syncthreads();
      smem_in_use = true;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (!pv[r]) continue;
      float ix, iy;
      ws_sample(th_, pw[r], ph[r], a.W, a.H, ix, iy);
      const float fx = floorf(ix), fy = floorf(iy);
      const int x0 = (int)fx, y0 = (int)fy;
      const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
      float val[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) val[e] = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
        if (xx < 0 || xx >= a.W || yy < 0 || yy >= a.H) continue;   // zeros padding
        const float wt = ((t & 1) ? wx1 : wx0) * ((t >> 1) ? wy1 : wy0);
        float f[8];
        if (staged && xx >= bx0 && xx <= bx1 && yy >= by0 && yy <= by1) {
          const uint32_t sa = smem0 + (uint32_t)((yy - by0) * bw + (xx - bx0)) * PX_BYTES + (uint32_t)sub * 16u;
          uint4 q, ql = make_uint4(0, 0, 0, 0);
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(sa));
          if (PLANES == 2)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(ql.x), "=r"(ql.y), "=r"(ql.z), "=r"(ql.w) : "r"(sa + 128u));
          ws_unpack8<PLANES>(q, ql, f);
        } else {
          load_direct(src_local, yy, xx, f);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) val[e] = fmaf(wt, f[e], val[e]);
      }
      if (KIND == WF_REDUCE && a.mode == 2) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[r][e] = fmaxf(acc[r][e], val[e]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[r][e] = fmaf(wgt_px[r], val[e], acc[r][e]);
      }
    }
  }
  float scale = 1.f;
  if (KIND == WF_MEAN) scale = count > 0 ? 1.f / (float)count : 0.f;
  if (KIND == WF_REDUCE && a.mode == 0) scale = 1.f / (float)count;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (!pv[r]) continue;
    if (KIND == WF_MEAN || (KIND == WF_REDUCE && a.mode == 0)) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[r][e] *= scale;
    }
    store_out(r);
  }
}

// OPT-IN (V2X_WARP_STAGED=1): measured on B200 (tools/warp_bench.py, profiles/r02_warp_bench.jsonl) the staged kernel is
// bit-compatible with the direct gathers but SLOWER -- 93 vs 76 us for V2VNet's 40-map mean, 96 vs 80 us for the when2com
// gate -- although it moves about half the L2->SM bytes: with four terms per tile, each a dependent
// load -> barrier -> compute -> barrier chain over ~30 KB, a CTA is latency-bound, while the direct kernel keeps 16
// independent 16-byte loads per lane in flight.  The direct kernels therefore stay the default; the environment variable
// is read per launch (launches are captured into graphs) so the tests can A/B both paths.  Takes channel counts that
// split into 64-channel slices.
static inline bool warp_staged_ok(int c) {
  const char* e = getenv("V2X_WARP_STAGED");
  return e && e[0] == '1' && c % kWsSlice == 0;
}

template <int KIND>
static inline cudaError_t launch_warp_fuse_staged(const WarpFuseArgs& a, int planes, cudaStream_t s) {
  const long long tiles = (long long)((a.H + kWsTile - 1) / kWsTile) * ((a.W + kWsTile - 1) / kWsTile);
  const long long blocks = (long long)a.unit_count * tiles * (a.C / kWsSlice);
  const size_t smem = (size_t)kWsCap * 128 * planes;
  // ask for the largest shared-memory carve-out so four 48 KB CTAs are resident per SM (per device, once)
  static bool carved[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !carved[dev]) {
    cudaFuncSetAttribute(warp_fuse_staged_kernel<KIND, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(warp_fuse_staged_kernel<KIND, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    carved[dev] = true;
  }
  if (planes == 2) warp_fuse_staged_kernel<KIND, 2><<<(unsigned)blocks, kWsThreads, smem, s>>>(a);
  else warp_fuse_staged_kernel<KIND, 1><<<(unsigned)blocks, kWsThreads, smem, s>>>(a);
  return cudaGetLastError();
}

}  // namespace v2x
