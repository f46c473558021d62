// Weight gradient of a 3x3 (pad 1, stride 1 / 2) or 1x1 convolution on the sm_100a tensor cores.
//
//   dW[co][ci_off + ci][kh][kw] += scale * sum_{n, oh, ow} dz[n][oh][ow][co] * x[n][oh*s + kh - 1][ow*s + kw - 1][ci]
//
// as a GEMM with  M = output channels (128 TMEM lanes),  N = input channels (64 fp32 TMEM columns per filter tap),
// K = pixels.  Both operands are NHWC act tensors, i.e. for this GEMM the contiguous dimension is M / N, not K: they
// are consumed as **MN-major** UMMA operands (instruction descriptor a_major = b_major = 1).  A TMA box
// {64 channels, 16, 4} of an NHWC map lands in shared memory as 64 rows (= pixels = K) of 128 bytes (= 64 channels =
// MN) with the 128-byte swizzle -- exactly the canonical MN-major SWIZZLE_128B layout ((8,n),(8,k)):((1,LBO),(8,SBO))
// in 16-byte units: 8 K rows per 1024-byte atom (SBO = 1024), the next 64 channels in the next box (LBO = box size).
// So there is no transpose anywhere: dz and the tap-shifted x boxes are fed to tcgen05.mma as they lie in HBM, and the
// shift of a filter tap is just the TMA box coordinate (out-of-range pixels are zero-filled = the conv's padding).
//
// CTA = (128-channel block of co) x (64-channel block of ci) x (filter row kh): three accumulators (kw = 0, 1, 2; 192
// TMEM columns) summed over this CTA's share of the 4x16-pixel K blocks (split-K over blockIdx.x), then reduced into
// the fp32 filter gradient with atomics.  fp16 hi/lo operands take three tensor-core passes per k-step (hi*hi + hi*lo
// + lo*hi), like the forward convs.  Warp roles as in conv_tc_kernel: warp 0 TMA producer, warp 1 MMA issuer, warps
// 2-5 read the accumulators back at the end.
//
// Replaces torch.autograd's conv2d weight-gradient node under loss.backward() (CP/utils/CoDetModule.py:289-291).
#include <mutex>

#include "common.cuh"

namespace v2x {

int encode_map_shared(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                      const cuuint32_t* box, int kc);

constexpr int kWtThreads = 192;
constexpr int kWtTileH = 4, kWtTileW = 16;          // one K block = 64 output pixels
constexpr uint32_t kWtBox = 64u * 128u;             // bytes of one {64 ch, 16, 4} box
constexpr int kWtMaxStages = 4;

struct WgDev {
  int n_maps, h_out, w_out, stride;
  int co_log, ci_log, ci_off, ci_total, taps_total;
  int co_tiles, ci_blocks;
  int tiles_w, tiles_per_img, num_tiles;
  int num_stages;
  int ci_phys;                 // channels of the x tensor (parity offset of the stride-2 view)
  float scale;
  float* dw;
};

// MN-major shared-memory matrix descriptor, SWIZZLE_128B: rows (K) of 128 bytes, 8-row atoms `sbo` bytes apart, 64-element
// MN blocks `lbo` bytes apart
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;      // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;      // SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor with BOTH operands MN-major, M = 128, fp32 accumulate
template <int PLANES>
__host__ __device__ constexpr uint32_t make_idesc_mn_m128(uint32_t n) {
  return make_idesc_m128<PLANES>(n) | (1u << 15) | (1u << 16);
}

template <int PLANES, int TAPS>   // TAPS = filter taps handled by one CTA: 3 (one row of a 3x3 filter) or 1 (1x1 conv)
__global__ void __launch_bounds__(kWtThreads, 1) conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmP,
                                                                      const __grid_constant__ CUtensorMap tmQ, const WgDev p) {
  constexpr uint32_t STAGE = (uint32_t)(PLANES * 2 + TAPS * PLANES) * kWtBox;
  constexpr uint32_t TMEM_COLS = TAPS == 3 ? 256u : 64u;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kWtMaxStages + 1];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[kWtMaxStages]);
  const uint32_t bar_done = smem_u32(&bars[2 * kWtMaxStages]);
  // blockIdx.y = (co tile, ci block, filter row)
  int y = blockIdx.y;
  const int kh = TAPS == 3 ? y % 3 : 0;
  if (TAPS == 3) y /= 3;
  const int co0 = (y % p.co_tiles) * 128, ci0 = (y / p.co_tiles) * 64;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmP);
    prefetch_tmap(&tmQ);
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: per K block the dz boxes (2 channel blocks x planes) and the tap-shifted x boxes =====
    int stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int im = tile / p.tiles_per_img, r = tile - im * p.tiles_per_img;
      const int oh0 = (r / p.tiles_w) * kWtTileH, ow0 = (r % p.tiles_w) * kWtTileW;
      mbar_wait(bar_empty + 8 * stage, phase ^ 1);
      if (elect_one()) {
        const uint32_t full = bar_full + 8 * stage;
        mbar_expect_tx(full, STAGE);
        const uint32_t sa = smem_base + (uint32_t)stage * STAGE;
#pragma unroll
        for (int pl = 0; pl < PLANES; ++pl)
#pragma unroll
          for (int blk = 0; blk < 2; ++blk)
            tma_load_4d(sa + (uint32_t)(pl * 2 + blk) * kWtBox, &tmP, full, co0 + blk * 64, ow0, oh0, pl * p.n_maps + im);
#pragma unroll
        for (int t = 0; t < TAPS; ++t)
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl) {
            const uint32_t dst = sa + (uint32_t)(PLANES * 2 + t * PLANES + pl) * kWtBox;
            const int img = pl * p.n_maps + im;
            if (TAPS == 1) {
              tma_load_4d(dst, &tmQ, full, ci0, ow0, oh0, img);
            } else if (p.stride == 1) {
              tma_load_4d(dst, &tmQ, full, ci0, ow0 + t - 1, oh0 + kh - 1, img);
            } else {
              // input row 2*oh + kh - 1 = 2*(oh + hoff) + hp of the {2C, W/2, 2, H/2, N*planes} view, same along w
              const int hp = kh == 1 ? 0 : 1, hoff = kh == 0 ? -1 : 0;
              const int wp = t == 1 ? 0 : 1, woff = t == 0 ? -1 : 0;
              tma_load_5d(dst, &tmQ, full, wp * p.ci_phys + ci0, ow0 + woff, hp, oh0 + hoff, img);
            }
          }
      }
      __syncwarp();
      if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_mn_m128<PLANES>(64);
    int stage = 0, phase = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(bar_full + 8 * stage, phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_base + (uint32_t)stage * STAGE;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {          // 16 pixels (two 8-row swizzle atoms = 2048 bytes) per k-step
          const uint64_t a_hi = make_desc_mn_sw128(sa + ks * 2048u, kWtBox, 1024u);
          const uint64_t a_lo = make_desc_mn_sw128(sa + 2u * kWtBox + ks * 2048u, kWtBox, 1024u);
#pragma unroll
          for (int t = 0; t < TAPS; ++t) {
            const uint32_t qb = sa + (uint32_t)(PLANES * 2 + t * PLANES) * kWtBox + ks * 2048u;
            const uint64_t b_hi = make_desc_mn_sw128(qb, kWtBox, 1024u);
            const uint32_t d = tmem_base + (uint32_t)t * 64u;
            umma_bf16(d, a_hi, b_hi, idesc, (first && ks == 0) ? 0u : 1u);
            if (PLANES == 2) {
              const uint64_t b_lo = make_desc_mn_sw128(qb + kWtBox, kWtBox, 1024u);
              umma_bf16(d, a_hi, b_lo, idesc, 1u);
              umma_bf16(d, a_lo, b_hi, idesc, 1u);
            }
          }
        }
        umma_commit(bar_empty + 8 * stage);
      }
      __syncwarp();
      first = false;
      if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(bar_done);
    __syncwarp();
  } else {
    // ===== epilogue warps: accumulators -> fp32 atomics (TMEM lane = output channel, column = input channel) =====
    const int quad = warp & 3;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const int co = co0 + quad * 32 + lane;
    const bool any_tile = blockIdx.x < p.num_tiles;
#pragma unroll 1
    for (int t = 0; t < TAPS; ++t) {
#pragma unroll 1
      for (int c16 = 0; c16 < 4; ++c16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * 64 + c16 * 16), v);
        if (any_tile && co < p.co_log) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int ci = ci0 + c16 * 16 + i;
            if (ci < p.ci_log)
              atomicAdd(p.dw + ((long long)co * p.ci_total + p.ci_off + ci) * p.taps_total + (TAPS == 3 ? kh * 3 + t : 0),
                        v[i] * p.scale);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int PLANES, int TAPS>
static int launch_wgrad_tc(const WgDev& d, const CUtensorMap& tp, const CUtensorMap& tq, dim3 grid, size_t smem, cudaStream_t stream) {
  static std::mutex mu;
  static uint64_t done_mask = 0;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (!((done_mask >> (dev & 63)) & 1ull)) {
      cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel<PLANES, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_wgrad_tc_kernel)");
      done_mask |= 1ull << (dev & 63);
    }
  }
  conv_wgrad_tc_kernel<PLANES, TAPS><<<grid, kWtThreads, smem, stream>>>(tp, tq, d);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

}  // namespace v2x

using namespace v2x;

// Same contract as v2x_conv_wgrad (train_kernels.cu), on the tensor cores.  Needs 16 | w_out-tile geometry only through
// TMA zero fill, i.e. any map size; channel counts are multiples of 16 as everywhere on the path.
extern "C" int v2x_conv_wgrad_tc(const void* dz, const void* x, int32_t n, int32_t h_out, int32_t w_out, int32_t co, int32_t ci,
                                 int32_t planes, int32_t stride, int32_t taps, float* dw, int32_t co_log, int32_t ci_log,
                                 int32_t ci_off, int32_t ci_total, float scale, void* stream) {
  V2X_REQUIRE(dz && x && dw && n > 0 && h_out > 0 && w_out > 0, "null/empty");
  V2X_REQUIRE(co > 0 && co % 8 == 0 && ci > 0 && ci % 8 == 0, "channels must be multiples of 8");
  V2X_REQUIRE(planes == 1 || planes == 2, "planes must be 1 or 2");
  V2X_REQUIRE((stride == 1 || stride == 2) && (taps == 9 || (taps == 1 && stride == 1)), "3x3 stride 1/2 or 1x1 stride 1");
  V2X_REQUIRE(co_log > 0 && co_log <= co && ci_log > 0 && ci_log <= ci && ci_off >= 0 && ci_off + ci_log <= ci_total,
              "bad logical channel window");
  V2X_REQUIRE(stride == 1 || ci % 64 == 0, "stride-2 wgrad on the tensor cores needs 64 | ci (the parity view interleaves "
                                           "the two column parities along the channel axis)");
  WgDev d{};
  d.n_maps = n; d.h_out = h_out; d.w_out = w_out; d.stride = stride;
  d.co_log = co_log; d.ci_log = ci_log; d.ci_off = ci_off; d.ci_total = ci_total; d.taps_total = taps;
  d.co_tiles = (co_log + 127) / 128;
  d.ci_blocks = (ci_log + 63) / 64;
  d.tiles_w = (w_out + kWtTileW - 1) / kWtTileW;
  d.tiles_per_img = d.tiles_w * ((h_out + kWtTileH - 1) / kWtTileH);
  d.num_tiles = n * d.tiles_per_img;
  d.scale = scale; d.dw = dw;
  d.ci_phys = ci;
  const int taps_cta = taps == 9 ? 3 : 1;
  const uint32_t stage = (uint32_t)(planes * 2 + taps_cta * planes) * kWtBox;
  int stages = (int)((219u * 1024u) / stage);
  if (stages > kWtMaxStages) stages = kWtMaxStages;
  V2X_REQUIRE(stages >= 2, "wgrad stage does not fit twice in shared memory");
  d.num_stages = stages;
  const size_t smem = (size_t)stages * stage + 1024;

  CUtensorMap tmP, tmQ;
  const int h_in = h_out * stride, w_in = w_out * stride;
  {
    const cuuint64_t C = (cuuint64_t)co, NP = (cuuint64_t)n * planes;
    cuuint64_t dims[4] = {C, (cuuint64_t)w_out, (cuuint64_t)h_out, NP};
    cuuint64_t str[3] = {C * 2, (cuuint64_t)w_out * C * 2, (cuuint64_t)h_out * w_out * C * 2};
    cuuint32_t box[4] = {64, kWtTileW, kWtTileH, 1};
    int rc = encode_map_shared(&tmP, dz, 4, dims, str, box, 64);
    if (rc) return rc;
  }
  {
    const cuuint64_t C = (cuuint64_t)ci, NP = (cuuint64_t)n * planes;
    int rc;
    if (stride == 1) {
      cuuint64_t dims[4] = {C, (cuuint64_t)w_in, (cuuint64_t)h_in, NP};
      cuuint64_t str[3] = {C * 2, (cuuint64_t)w_in * C * 2, (cuuint64_t)h_in * w_in * C * 2};
      cuuint32_t box[4] = {64, kWtTileW, kWtTileH, 1};
      rc = encode_map_shared(&tmQ, x, 4, dims, str, box, 64);
    } else {
      cuuint64_t dims[5] = {2 * C, (cuuint64_t)w_in / 2, 2, (cuuint64_t)h_in / 2, NP};
      cuuint64_t str[4] = {2 * C * 2, (cuuint64_t)w_in * C * 2, 2 * (cuuint64_t)w_in * C * 2, (cuuint64_t)h_in * w_in * C * 2};
      cuuint32_t box[5] = {64, kWtTileW, 1, kWtTileH, 1};
      rc = encode_map_shared(&tmQ, x, 5, dims, str, box, 64);
    }
    if (rc) return rc;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int gy = d.co_tiles * d.ci_blocks * taps_cta;
  int gx = (sms + gy - 1) / gy;            // one resident CTA per SM; every (co, ci, kh) block is split over gx CTAs along K
  if (gx > d.num_tiles) gx = d.num_tiles;
  if (gx < 1) gx = 1;
  const dim3 grid(gx, gy);
  cudaStream_t s = (cudaStream_t)stream;
  if (planes == 1) return taps == 9 ? launch_wgrad_tc<1, 3>(d, tmP, tmQ, grid, smem, s) : launch_wgrad_tc<1, 1>(d, tmP, tmQ, grid, smem, s);
  return taps == 9 ? launch_wgrad_tc<2, 3>(d, tmP, tmQ, grid, smem, s) : launch_wgrad_tc<2, 1>(d, tmP, tmQ, grid, smem, s);
}
