// Tap-packed 3x3 convolution for the narrow-output decoder layers (conv8_1: 96 -> 32 at 256x256, conv7_1: 192 -> 64 at
// 128x128; CP/models/det/backbone/Backbone.py:211-237) on the sm_100a tensor cores.  Output channels are processed in
// groups of 32 (blockIdx.y); the text below describes one group.
//
// Why: a tcgen05.mma of M=128, K=16 re-reads its whole 128-row A sub-tile from shared memory whatever N is, so its cost
// is max(N/2, (128+N)/4) cycles (tools/mma_probe.cu) -- at N = 32 the tensor pipe idles 60% of the time waiting for
// operand reads.  The three horizontal taps of a 3x3 filter read the SAME activations shifted by one pixel, so they are
// packed into the N dimension instead of being issued as separate MMAs:
//
//     Y[q, kw*32 + co] = sum_{kh, ci} x[q + kh*PITCH, ci] * W[co, ci, kh, kw]         (N = 96, K = 3 * Cin)
//     out[p, co]       = Y[p, co] + Y[p + 1, 32 + co] + Y[p + 2, 64 + co]              (epilogue, warp shuffles)
//
// with q, p linear positions in the tile's halo (PITCH = 16 px per halo row).  One k-step now costs 3 MMAs x 56 cycles
// instead of 9 x 40, i.e. 2.1x less tensor-pipe time per output pixel; the price is a tile of 8 x 14 valid output pixels
// per 128 TMEM lanes (the two right-most columns of every 16-lane row only feed their neighbours' sums: 87.5% of the
// lanes do useful work) and two shuffles + adds per output value in the epilogue.
//
//   A operand: ONE TMA box per channel block, the 10 x 16 pixel halo of the tile ({kc, 16, 10, 1} of the NHWC act map,
//              zero-filled outside the map = the conv's padding); filter row kh is read by starting the UMMA
//              descriptor kh halo rows (kh * 16 * kc*2 bytes, swizzle-atom aligned) into the box.
//   B operand: packed weights [planes][96][3 * Cin] (row = kw*32 + co, k = (source, kh, ci)), resident in shared
//              memory for the CTA's lifetime.
//   Persistent, warp-specialised like conv_tc_kernel: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue,
//   two TMEM accumulators (2 x 128 columns), two co-resident CTAs per SM when the operands fit.
#include <mutex>

#include "common.cuh"

namespace v2x {

int encode_map_shared(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                      const cuuint32_t* box, int kc);

constexpr int kP3Threads = 192;
constexpr int kP3TileH = 8, kP3TileW = 14, kP3HaloH = 10, kP3HaloW = 16;
constexpr int kP3MaxStages = 12;
constexpr int kP3MaxCB = 16;   // channel blocks over both sources
constexpr int kP3N = 96;

struct Pack3Dev {
  int n_maps, h_out, w_out, nsrc;
  int cin[2], cblocks[2], num_cb;
  int num_stages;
  int tiles_w, tiles_per_img, m_tiles;
  int relu, debug_mode;
  int groups;                // cout / 32: CTAs with blockIdx.y = g own output channels [32g, 32g + 32)
  void* out0;
  int out_c_total, out_c_off;
  long long out_plane_stride;
  const float* bias;
  uint32_t b_region_bytes, epi_off;
};

__device__ __forceinline__ void p3_st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 p3_ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

constexpr uint32_t kP3StagePlane = 32u * 64u;   // one warp's 32 pixels x 32 bf16 channels

template <int PLANES, int MMAS, int KSTEPS>
__global__ void __launch_bounds__(kP3Threads, PLANES == 1 ? 2 : 1)
    conv_pack3_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                      const __grid_constant__ CUtensorMap tmB, const Pack3Dev p) {
  constexpr int KC = 16 * KSTEPS;
  // operand planes actually read (see conv_tc.cuh): MMAS = 3 both split, 2 activations split only, 1 hi planes only
  constexpr int PA = (PLANES == 2 && MMAS >= 2) ? 2 : 1;
  constexpr int PW = (PLANES == 2 && MMAS == 3) ? 2 : 1;
  constexpr uint32_t ROW = KC * 2u;                                     // bytes of one pixel's channel block
  constexpr uint32_t A_BOX = (uint32_t)(kP3HaloH * kP3HaloW) * ROW;     // 160 halo pixels (multiple of 1 KB)
  constexpr uint32_t B_TILE = (uint32_t)kP3N * ROW;                     // 96 weight rows (multiple of 1 KB)
  constexpr uint32_t STAGE = PA * A_BOX;
  constexpr uint32_t SBO = 8u * ROW;
  constexpr uint32_t LAYOUT = KC == 64 ? 2u : KC == 32 ? 4u : 6u;
  constexpr uint32_t ACC_STRIDE = 128u, TMEM_COLS = 256u;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kP3MaxStages + 5];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[32];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = blockIdx.y;                                           // 32-output-channel group of this CTA
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;    // resident weights, then the halo ring
  const uint32_t ring_base = smem_base + p.b_region_bytes;
  const uint32_t bar_full = smem_u32(&bars[0]);
  const uint32_t bar_empty = smem_u32(&bars[kP3MaxStages]);
  const uint32_t bar_bres = smem_u32(&bars[2 * kP3MaxStages]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * kP3MaxStages + 1]);    // [2]
  const uint32_t bar_tempty = smem_u32(&bars[2 * kP3MaxStages + 3]);   // [2]

  if (threadIdx.x < 32) s_bias[threadIdx.x] = p.bias[grp * kP3N + threadIdx.x];
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA0);
    if (p.nsrc > 1) prefetch_tmap(&tmA1);
    prefetch_tmap(&tmB);
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_bres, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 4);   // one arrive per epilogue warp
    }
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
  const int grid_stride = gridDim.x;
  const int total_cin0 = p.cin[0];
  pdl_launch_dependents();

  if (warp == 0) {
    // ===== TMA producer =====
    const bool no_tma = p.debug_mode == 2;
    if (elect_one()) {
      // resident weights: tile (cb_g, kh) at smem_base + ((cb_g * 3 + kh) * PLANES + pl) * B_TILE
      mbar_expect_tx(bar_bres, (uint32_t)p.num_cb * 3u * PW * B_TILE);
      for (int cbg = 0; cbg < p.num_cb; ++cbg) {
        const int s = cbg < p.cblocks[0] ? 0 : 1;
        const int cb = cbg - s * p.cblocks[0];
        for (int kh = 0; kh < 3; ++kh) {
          const int kcol = (s ? 3 * total_cin0 : 0) + kh * p.cin[s] + cb * KC;
#pragma unroll
          for (int pl = 0; pl < PW; ++pl)
            tma_load_2d(smem_base + (uint32_t)((cbg * 3 + kh) * PW + pl) * B_TILE, &tmB, bar_bres, kcol,
                        (pl * p.groups + grp) * kP3N);
        }
      }
    }
    __syncwarp();
    pdl_wait();   // weights are constants; the activations were written by the previous kernel(s)
    int stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += grid_stride) {
      const int img = tile / p.tiles_per_img;
      const int r = tile - img * p.tiles_per_img;
      const int th = r / p.tiles_w, tw = r - th * p.tiles_w;
      const int oh0 = th * kP3TileH, ow0 = tw * kP3TileW;
      for (int cbg = 0; cbg < p.num_cb; ++cbg) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        if (elect_one()) {
          const uint32_t full = bar_full + 8 * stage;
          if (no_tma) {
            mbar_arrive(full);
          } else {
            const int s = cbg < p.cblocks[0] ? 0 : 1;
            const int cb = cbg - s * p.cblocks[0];
            mbar_expect_tx(full, STAGE);
            const uint32_t sa = ring_base + (uint32_t)stage * STAGE;
#pragma unroll
            for (int pl = 0; pl < PA; ++pl)
              tma_load_4d(sa + pl * A_BOX, s ? &tmA1 : &tmA0, full, cb * KC, ow0 - 1, oh0 - 1, pl * p.n_maps + img);
          }
        }
        __syncwarp();
        if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_m128<PLANES>(kP3N);
    const bool no_mma = p.debug_mode == 1;
    mbar_wait(bar_bres, 0);
    const uint64_t desc_ring = make_smem_desc(ring_base, SBO, LAYOUT);
    const uint64_t desc_b = make_smem_desc(smem_base, SBO, LAYOUT);
    int stage = 0, phase = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += grid_stride, ++it) {
      const int acc_buf = it & 1;
      mbar_wait(bar_tempty + 8 * acc_buf, ((it >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc_buf * ACC_STRIDE;
      for (int cbg = 0; cbg < p.num_cb; ++cbg) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        if (elect_one()) {
          if (no_mma) {
            mbar_arrive(bar_empty + 8 * stage);
          } else {
            const uint64_t da = desc_ring + (uint64_t)((uint32_t)stage * (STAGE >> 4));
            const uint64_t db = desc_b + (uint64_t)((uint32_t)(cbg * 3 * PW) * (B_TILE >> 4));
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              const uint32_t a_off = (uint32_t)(kh * kP3HaloW) * (ROW >> 4);
              const uint32_t b_off = (uint32_t)(kh * PW) * (B_TILE >> 4);
#pragma unroll
              for (int kk = 0; kk < KSTEPS; ++kk) {
                umma_bf16(tmem_d, da + (a_off + 2 * kk), db + (b_off + 2 * kk), idesc, (cbg | kh | kk) == 0 ? 0u : 1u);
                if (PW == 2) umma_bf16(tmem_d, da + (a_off + 2 * kk), db + (b_off + 2 * kk + (B_TILE >> 4)), idesc, 1u);
                if (PA == 2) umma_bf16(tmem_d, da + (a_off + 2 * kk + (A_BOX >> 4)), db + (b_off + 2 * kk), idesc, 1u);
              }
            }
            umma_commit(bar_empty + 8 * stage);
          }
        }
        __syncwarp();
        if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) {
        if (no_mma) mbar_arrive(bar_tfull + 8 * acc_buf);
        else umma_commit(bar_tfull + 8 * acc_buf);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue warps: TMEM lane quadrant (warp & 3); lane = halo position (r, c) = (row >> 4, row & 15) =====
    pdl_wait();
    const int quad = warp & 3;
    const bool no_store = p.debug_mode == 3;
    const bool relu = p.relu != 0;
    const uint32_t stg = smem_base + p.epi_off + (uint32_t)(warp - 2) * (PLANES * kP3StagePlane);
    const uint32_t srow = stg + (uint32_t)lane * 64u;
    const uint32_t ssw = (uint32_t)(lane >> 1) & 3u;
    const int unit = lane & 3;
    __nv_bfloat16* const out = reinterpret_cast<__nv_bfloat16*>(p.out0) + p.out_c_off + grp * 32 + unit * 8;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += grid_stride, ++it) {
      const int img = tile / p.tiles_per_img;
      const int rr = tile - img * p.tiles_per_img;
      const int th = rr / p.tiles_w, tw = rr - th * p.tiles_w;
      const int oh0 = th * kP3TileH, ow0 = tw * kP3TileW;
      const int acc_buf = it & 1;
      mbar_wait(bar_tfull + 8 * acc_buf, (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc_buf * ACC_STRIDE + ((uint32_t)(quad * 32) << 16);
      // all 96 accumulator columns of this lane up front (one TMEM round trip), then hand the buffer straight back to
      // the MMA warp: the shuffles / packing / stores below overlap the next tile's MMAs
      float y0[32], y1[32], y2[32];
      tmem_ld16_async(taddr, y0);
      tmem_ld16_async(taddr + 16, y0 + 16);
      tmem_ld16_async(taddr + 32, y1);
      tmem_ld16_async(taddr + 48, y1 + 16);
      tmem_ld16_async(taddr + 64, y2);
      tmem_ld16_async(taddr + 80, y2 + 16);
      tmem_ld_wait16(y0);
      tmem_ld_wait16(y0 + 16);
      tmem_ld_wait16(y1);
      tmem_ld_wait16(y1 + 16);
      tmem_ld_wait16(y2);
      tmem_ld_wait16(y2 + 16);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc_buf);
#pragma unroll
      for (int c16 = 0; c16 < 2; ++c16) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          // out[p] = Y_kw0[p] + Y_kw1[p + 1] + Y_kw2[p + 2]: the neighbours sit one / two lanes up in the same 16-lane row
          const float a1 = __shfl_down_sync(0xffffffffu, y1[c16 * 16 + i], 1);
          const float a2 = __shfl_down_sync(0xffffffffu, y2[c16 * 16 + i], 2);
          const float v = (y0[c16 * 16 + i] + a1) + a2 + s_bias[c16 * 16 + i];
          y0[c16 * 16 + i] = relu ? fmaxf(v, 0.f) : v;
        }
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) act_pack2<PLANES>(y0[c16 * 16 + 2 * i], y0[c16 * 16 + 2 * i + 1], hi[i], lo[i]);
        const uint32_t u0 = (((uint32_t)(2 * c16)) ^ ssw) << 4, u1 = (((uint32_t)(2 * c16 + 1)) ^ ssw) << 4;
        p3_st_shared_v4(srow + u0, hi[0], hi[1], hi[2], hi[3]);
        p3_st_shared_v4(srow + u1, hi[4], hi[5], hi[6], hi[7]);
        if (PLANES == 2) {
          p3_st_shared_v4(srow + kP3StagePlane + u0, lo[0], lo[1], lo[2], lo[3]);
          p3_st_shared_v4(srow + kP3StagePlane + u1, lo[4], lo[5], lo[6], lo[7]);
        }
      }
      __syncwarp();
      // coalesced write-back: lane -> 16-byte unit (lane & 3) of pixels (lane >> 2) + 8j of this warp's 2 rows x 16 columns
      __nv_bfloat16* const tile_p = out + (((long long)img * p.h_out + oh0) * p.w_out + ow0) * p.out_c_total;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pix = (lane >> 2) + 8 * j;
        const int ph = quad * 2 + (pix >> 4), pw = pix & 15;
        if (pw < kP3TileW && oh0 + ph < p.h_out && ow0 + pw < p.w_out && !no_store) {
          const uint32_t a = stg + (uint32_t)pix * 64u + ((((uint32_t)unit) ^ ((uint32_t)(pix >> 1) & 3u)) << 4);
          __nv_bfloat16* dst = tile_p + ((long long)ph * p.w_out + pw) * p.out_c_total;
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl)
            *reinterpret_cast<uint4*>(dst + pl * p.out_plane_stride) = p3_ld_shared_v4(a + pl * kP3StagePlane);
        }
      }
      __syncwarp();   // the staging buffer is reused by the next tile
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int PLANES, int MMAS, int KSTEPS>
static int launch_pack3_t(const Pack3Dev& d, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, size_t smem,
                          dim3 ctas, cudaStream_t stream) {
  static std::mutex mu;       // function attributes are per device
  static uint64_t done_mask = 0;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (!((done_mask >> (dev & 63)) & 1ull)) {
      cudaError_t attr_err = cudaFuncSetAttribute(conv_pack3_kernel<PLANES, MMAS, KSTEPS>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(conv_pack3_kernel)");
      done_mask |= 1ull << (dev & 63);
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = ctas;
  cfg.blockDim = dim3(kP3Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  V2X_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_pack3_kernel<PLANES, MMAS, KSTEPS>, a0, a1, b, d));
  return V2X_OK;
}

// Host side of v2x_conv_fwd for params with tap_pack != 0 (validated here; called from conv_tcgen05.cu).
int launch_pack3(const v2x_conv_params* p, int debug_mode, cudaStream_t stream) {
  V2X_REQUIRE(p->src[0] && p->weights && p->bias && p->out0, "null src/weights/bias/out0");
  V2X_REQUIRE(p->taps == 9 && p->stride == 1 && p->cout > 0 && p->cout % 32 == 0 && p->cout_pad == 3 * p->cout &&
                  p->block_n == kP3N,
              "tap_pack needs a 3x3 stride-1 conv with 32 | cout, cout_pad == 3 * cout and block_n == 96");
  V2X_REQUIRE(p->epilogue == V2X_EPI_ACT && !p->upsample2x, "tap_pack supports the plain EPI_ACT epilogue only");
  V2X_REQUIRE(p->planes == 1 || p->planes == 2, "planes must be 1 or 2");
  V2X_REQUIRE(p->n_maps > 0 && p->h_out > 0 && p->w_out > 0 && p->h_out % kP3TileH == 0, "tap_pack needs 8 | H");
  V2X_REQUIRE(p->cin[0] > 0 && p->cin[0] % 16 == 0 && (p->src[1] == nullptr || (p->cin[1] > 0 && p->cin[1] % 16 == 0)),
              "channels per source must be multiples of 16");
  V2X_REQUIRE(p->out_c_total >= p->out_c_off + p->cout && p->out_c_total % 8 == 0 && p->out_c_off % 8 == 0,
              "bad output channel window");
  Pack3Dev d{};
  d.n_maps = p->n_maps; d.h_out = p->h_out; d.w_out = p->w_out;
  d.nsrc = p->src[1] ? 2 : 1;
  d.cin[0] = p->cin[0]; d.cin[1] = d.nsrc > 1 ? p->cin[1] : 0;
  int kc = 64;
  for (int s = 0; s < d.nsrc; ++s)
    while (d.cin[s] % kc) kc >>= 1;
  d.cblocks[0] = d.cin[0] / kc; d.cblocks[1] = d.cin[1] / kc;
  d.num_cb = d.cblocks[0] + d.cblocks[1];
  V2X_REQUIRE(d.num_cb <= kP3MaxCB, "too many channel blocks (%d)", d.num_cb);
  d.tiles_w = (p->w_out + kP3TileW - 1) / kP3TileW;
  d.tiles_per_img = d.tiles_w * (p->h_out / kP3TileH);
  d.m_tiles = p->n_maps * d.tiles_per_img;
  d.relu = p->relu; d.debug_mode = debug_mode;
  d.groups = p->cout / 32;
  d.out0 = p->out0; d.out_c_total = p->out_c_total; d.out_c_off = p->out_c_off;
  d.out_plane_stride = (long long)p->n_maps * p->h_out * p->w_out * p->out_c_total;
  d.bias = p->bias;
  const uint32_t row = (uint32_t)kc * 2u;
  const uint32_t a_box = (uint32_t)(kP3HaloH * kP3HaloW) * row, b_tile = (uint32_t)kP3N * row;
  V2X_REQUIRE(p->mmas >= 0 && p->mmas <= 3 && (p->planes == 2 || p->mmas <= 1), "mmas must be 1..3 (0 = default), and > 1 only with planes == 2");
  const int mmas = p->planes == 2 ? (p->mmas == 0 ? 3 : p->mmas) : 1;
  const int pa = mmas >= 2 ? 2 : 1, pw = mmas == 3 ? 2 : 1;
  const uint32_t stage = (uint32_t)pa * a_box;
  d.b_region_bytes = (uint32_t)d.num_cb * 3u * pw * b_tile;
  const uint32_t epi = 4u * (uint32_t)p->planes * kP3StagePlane;
  // two co-resident CTAs per SM (bf16) when weights + >= 3 stages fit in half the shared memory
  int ctas_per_sm = 1;
  uint32_t budget = 219u * 1024u;
  if (p->planes == 1 && d.b_region_bytes + epi + 3u * stage + 1024u <= 106u * 1024u) {
    ctas_per_sm = 2;
    budget = 106u * 1024u;
  }
  V2X_REQUIRE(d.b_region_bytes + epi + 2u * stage + 1024u <= budget, "tap_pack operands do not fit in shared memory");
  int stages = (int)((budget - d.b_region_bytes - epi - 1024u) / stage);
  if (stages > kP3MaxStages) stages = kP3MaxStages;
  d.num_stages = stages;
  d.epi_off = d.b_region_bytes + (uint32_t)stages * stage;
  const size_t smem = (size_t)d.epi_off + epi + 1024;

  CUtensorMap tmA[2], tmB;
  for (int s = 0; s < d.nsrc; ++s) {
    const cuuint64_t C = (cuuint64_t)d.cin[s], NP = (cuuint64_t)p->n_maps * p->planes;
    cuuint64_t dims[4] = {C, (cuuint64_t)p->w_out, (cuuint64_t)p->h_out, NP};
    cuuint64_t str[3] = {C * 2, (cuuint64_t)p->w_out * C * 2, (cuuint64_t)p->h_out * p->w_out * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)kc, kP3HaloW, kP3HaloH, 1};
    int rc = encode_map_shared(&tmA[s], p->src[s], 4, dims, str, box, kc);
    if (rc) return rc;
  }
  if (d.nsrc == 1) tmA[1] = tmA[0];
  {
    const cuuint64_t k_total = 3ull * (cuuint64_t)(d.cin[0] + d.cin[1]);
    cuuint64_t dims[2] = {k_total, (cuuint64_t)pw * d.groups * kP3N};
    cuuint64_t str[1] = {k_total * 2};
    cuuint32_t box[2] = {(cuuint32_t)kc, kP3N};
    int rc = encode_map_shared(&tmB, p->weights, 2, dims, str, box, kc);
    if (rc) return rc;
  }
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  int ctas_x = sms * ctas_per_sm / d.groups;
  if (ctas_x < 1) ctas_x = 1;
  if (ctas_x > d.m_tiles) ctas_x = d.m_tiles;
  const int rounds = (d.m_tiles + ctas_x - 1) / ctas_x;
  ctas_x = (d.m_tiles + rounds - 1) / rounds;
  const dim3 ctas(ctas_x, d.groups);
#define V2X_P3(KS_)                                                                                  \
  if (kc == 16 * KS_)                                                                                \
    return p->planes == 1 ? launch_pack3_t<1, 1, KS_>(d, tmA[0], tmA[1], tmB, smem, ctas, stream)    \
           : mmas == 3    ? launch_pack3_t<2, 3, KS_>(d, tmA[0], tmA[1], tmB, smem, ctas, stream)    \
           : mmas == 2    ? launch_pack3_t<2, 2, KS_>(d, tmA[0], tmA[1], tmB, smem, ctas, stream)    \
                          : launch_pack3_t<2, 1, KS_>(d, tmA[0], tmA[1], tmB, smem, ctas, stream);
  V2X_P3(1) V2X_P3(2) V2X_P3(4)
#undef V2X_P3
  set_error("tap_pack: no instantiation for kc %d", kc);
  return V2X_ERR_UNSUPPORTED;
}

}  // namespace v2x
