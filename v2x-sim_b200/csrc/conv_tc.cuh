// conv_tc.cuh -- the tensor-core convolution kernel family (device code + launch templates); see conv_tcgen05.cu for
// the design notes and the host side.
#pragma once
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "common.cuh"

namespace v2x {

constexpr int kMaxStages = 12;
constexpr int kNumThreads = 192;
constexpr int kSmemLimit = 220 * 1024;  // dynamic smem we opt into; ~5 KB of static smem (k-table, bias, barriers) lives beside it (227 KB per CTA)
constexpr int kTileH = 8;
constexpr int kTileW = 16;

struct ConvDev {
  int n_maps, h_out, w_out, stride, taps, planes;
  int pa, pw;                // planes of the A / B operands actually read: (1,1) one MMA, (2,1) two, (2,2) three per k-step
  int nsrc;
  int cin[2];
  int cblocks[2];
  int kc;
  int num_k;
  int num_stages;
  int tiles_w, tiles_per_img;
  int m_tiles, n_tiles;
  int b_resident;            // whole [BN x K] weight operand kept in smem for the CTA's lifetime
  uint32_t b_region_bytes;   // bytes reserved for it in front of the stage ring
  int kb_per_stage;          // k-blocks (tap x kc channels) per pipeline stage: one mbarrier round trip feeds them all
  int stages_per_tile;       // ceil(num_k / kb_per_stage)
  uint32_t kb_bytes, kb_tx_bytes;  // smem footprint / TMA bytes of one k-block
  int b_stages, b_taps;      // halo mode with streamed weights: B ring depth, taps per B stage (1 or 3)
  uint32_t b_stage_bytes, b_ring_off;
  int halo;                  // 1: 16x8 output tile, one 18x10 halo box per channel block feeds all 9 taps
  int num_b_tiles;           // weight tiles of the resident operand (= taps * sum(cblocks))
  uint32_t epi_off, epi_warp_bytes;  // per-warp output staging buffers (after the operand rings)
  // fused 1x1 tail (V2X_EPI_TAIL_F32_SPLIT): second GEMM on the bf16 ReLU output of this conv, never written to HBM
  int tail_cout, tail_cout_pad;
  const float* tail_bias;
  uint32_t tail_w_off, tail_a_off;   // smem: tail weights [planes][<=64 rows][128 B], A2 tile [planes][128 rows][128 B]
  int ctas_per_sm;           // 2: small-N layers run two co-resident CTAs per SM (two MMA issue streams, eight epilogue warps)
  int debug_mode;            // profiling ablations (v2x_set_debug_mode): 1 = no MMA, 2 = no TMA, 3 = no stores, 4 = no epilogue work
  int cout, cout_pad;
  int epilogue, relu, upsample2x;
  void* out0;
  void* out1;
  int out_c_total, out_c_off, split;
  const float* bias;
  const float* gru_bhn;
  const float* gru_add;      // optional fp32 [pixel][cout] pre-activation term added to the gate accumulators
  int gru_pre_act;           // src[1] = bf16 pre-activations [..][cout]; each N tile accumulates its window through identity weight columns
  const void* passthrough;
  const long long* num_agent;
  int batch, agents, map_offset;
  uint32_t a_tile_bytes, b_tile_bytes, stage_bytes, tx_bytes, sbo, layout_type;
  long long out_plane_stride;  // elements between output planes
  // reference (CUDA-core) kernel only
  const void* src[2];
  const void* weights;
  int k_total;
};

// ---------------------------------------------------------------------------------------------
// Epilogue pieces shared by the tensor-core kernel and the CUDA-core cross-check kernel.
// Each call handles 16 consecutive output channels of one output pixel.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_act16(const ConvDev& p, int n_img, int oh, int ow, int ch0, const float* v) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) act_pack2(v[2 * i], v[2 * i + 1], p.planes, hi[i], lo[i]);
  const int up = p.upsample2x ? 2 : 1;
  const int H = p.h_out * up, W = p.w_out * up;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out0);
  for (int dy = 0; dy < up; ++dy)
    for (int dx = 0; dx < up; ++dx) {
      const long long pix = ((long long)n_img * H + (oh * up + dy)) * W + (ow * up + dx);
      __nv_bfloat16* dst = out + pix * p.out_c_total + p.out_c_off + ch0;
      uint4* d4 = reinterpret_cast<uint4*>(dst);
      d4[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      d4[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      if (p.planes == 2) {
        uint4* l4 = reinterpret_cast<uint4*>(dst + p.out_plane_stride);
        l4[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        l4[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      }
    }
}

__device__ __forceinline__ void epi_act16(const ConvDev& p, int n_img, int oh, int ow, int ch0, float* v,
                                          const float* bias16) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float y = v[i] + bias16[i];
    v[i] = p.relu ? fmaxf(y, 0.f) : y;
  }
  store_act16(p, n_img, oh, ow, ch0, v);
}

__device__ __forceinline__ void epi_f32_split16(const ConvDev& p, int n_img, int oh, int ow, int ch0, const float* v,
                                                const float* bias16) {
  const long long pix = ((long long)n_img * p.h_out + oh) * p.w_out + ow;
  float* o0 = reinterpret_cast<float*>(p.out0) + pix * p.split;
  float* o1 = reinterpret_cast<float*>(p.out1) + pix * (p.cout - p.split) - p.split;
  // split, cout and both row strides are multiples of 4 (checked on the host): 16-byte stores
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int ch = ch0 + 4 * g;
    if (ch < p.cout) {
      float4 y;
      y.x = v[4 * g + 0] + bias16[4 * g + 0];
      y.y = v[4 * g + 1] + bias16[4 * g + 1];
      y.z = v[4 * g + 2] + bias16[4 * g + 2];
      y.w = v[4 * g + 3] + bias16[4 * g + 3];
      float* dst = ch < p.split ? o0 + ch : o1 + ch;
      if (p.debug_mode == 3 && y.x != 12345.678f) continue;  // profiling ablation: no stores
      *reinterpret_cast<float4*>(dst) = y;
    }
  }
}

// fp32 NCHW output (segmentation logits, SegModelBase.py:145-151): out0[((n * cout + ch) * H + oh) * W + ow]
__device__ __forceinline__ void epi_f32_nchw16(const ConvDev& p, int n_img, int oh, int ow, int ch0, const float* v,
                                               const float* bias16) {
  float* o = reinterpret_cast<float*>(p.out0) + (((long long)n_img * p.cout + ch0) * p.h_out + oh) * p.w_out + ow;
  const long long cs = (long long)p.h_out * p.w_out;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (ch0 + i < p.cout) o[i * cs] = v[i] + bias16[i];
}

// GRU gate activations on the SFU: sigmoid(x) = 1 / (1 + 2^(-x log2 e)), tanh(x) = 2 sigmoid(2x) - 1; ex2.approx and
// rcp.approx are accurate to ~2 ulp, i.e. ~1e-6 absolute on gates in (-1, 1) -- far inside the 1e-3 contract -- and an
// order of magnitude fewer instructions than the IEEE division + tanhf they replace (the GRU epilogue was the
// bottleneck of the three ConvGRU launches).
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return fmaf(2.f, __fdividef(1.f, 1.f + __expf(-2.f * x)), -1.f); }

// GRU gates for 16 channels [c0, c0+16) of one pixel.  bias_r16 points at the bias of the r gate of channel
// c0 inside the [r(64) | z(64) | n(64)] block (z at +64, n at +128); bhn16 at b_hh_n of channel c0.
__device__ __forceinline__ void epi_gru16(const ConvDev& p, int n_img, int oh, int ow, int c0, const float* r,
                                          const float* z, const float* nn, const float* bias_r16, const float* bhn16,
                                          const float* add_r16 = nullptr) {
  float h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float ar = add_r16 ? add_r16[i] : 0.f, az = add_r16 ? add_r16[64 + i] : 0.f, an = add_r16 ? add_r16[128 + i] : 0.f;
    const float rr = fast_sigmoid(r[i] + bias_r16[i] + ar);
    const float zz = fast_sigmoid(z[i] + bias_r16[64 + i] + az);
    const float nv = fast_tanh(nn[i] + bias_r16[128 + i] + an + rr * bhn16[i]);
    h[i] = (1.f - zz) * nv;
  }
  store_act16(p, n_img, oh, ow, c0, h);
}

// copy `nch` channels of one pixel from the passthrough tensor (absent agents keep their own map)
__device__ __forceinline__ void copy_passthrough(const ConvDev& p, int n_img, int oh, int ow, int c0, int nch) {
  const long long pix = ((long long)n_img * p.h_out + oh) * p.w_out + ow;
  const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(p.passthrough) + pix * p.out_c_total + p.out_c_off + c0;
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out0) + pix * p.out_c_total + p.out_c_off + c0;
  for (int pl = 0; pl < p.planes; ++pl)
    for (int c = 0; c < nch; c += 8)
      *reinterpret_cast<uint4*>(dst + pl * p.out_plane_stride + c) =
          *reinterpret_cast<const uint4*>(src + pl * p.out_plane_stride + c);
}

__device__ __forceinline__ bool gru_unit_absent(const ConvDev& p, int n_img) {
  if (p.epilogue != V2X_EPI_GRU || p.num_agent == nullptr) return false;
  const int unit = n_img + p.map_offset;  // global agent-major unit (sharded plans hold a slice of the maps)
  const int agent = unit / p.batch, b = unit % p.batch;
  return agent >= (int)p.num_agent[(long long)b * p.agents];
}

// ---------------------------------------------------------------------------------------------
// Lean epilogue pieces of the tensor-core kernel (PLANES compile-time, addresses hoisted per tile)
// ---------------------------------------------------------------------------------------------
// 16 fp32 -> the storage format: bf16 (PLANES == 1, a single cvt.rn.bf16x2.f32 per pair) or fp16 hi + lo (PLANES == 2)
template <int PLANES>
__device__ __forceinline__ void pack16(const float* v, uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int i = 0; i < 8; ++i) act_pack2<PLANES>(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
}

// ---------------------------------------------------------------------------------------------
// Coalesced output through a per-warp shared-memory staging buffer.
//
// TMEM hands every epilogue thread one output PIXEL (a row of the accumulator), so a direct store has the 32 lanes
// of a warp writing 16 bytes each at a stride of one pixel (64..192 bytes): 32 separate L2 requests per instruction,
// which caps the store rate near 16 B/clk/SM and made every write-heavy layer (upsampling decoder layers, heads)
// store-bound (measured with the no-store ablation).  Instead each thread parks a 32-column chunk of its pixel in
// shared memory (XOR-swizzled / padded rows: conflict-free both ways), and the warp writes the chunk back out with
// consecutive lanes on consecutive 16-byte units, i.e. whole 64..128-byte runs per pixel and 512-byte runs where
// pixels are adjacent in memory.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

constexpr uint32_t kStageActPlane = 32u * 64u;    // bf16 chunk: 32 pixels x 32 channels
constexpr uint32_t kStageF32 = 32u * 128u;        // fp32 chunk: 32 pixels x 32 channels, 16-byte units XOR-swizzled by pixel

// park 16 bf16 channels (two 16-byte units: 2*half, 2*half+1) of this lane's pixel
template <int PLANES>
__device__ __forceinline__ void stage_act16(uint32_t stage, int lane, int half, const uint32_t* hi, const uint32_t* lo) {
  const uint32_t row = stage + (uint32_t)lane * 64u;
  const uint32_t sw = (uint32_t)(lane >> 1) & 3u;
  const uint32_t u0 = (((uint32_t)(2 * half)) ^ sw) << 4, u1 = (((uint32_t)(2 * half + 1)) ^ sw) << 4;
  st_shared_v4(row + u0, hi[0], hi[1], hi[2], hi[3]);
  st_shared_v4(row + u1, hi[4], hi[5], hi[6], hi[7]);
  if (PLANES == 2) {
    st_shared_v4(row + kStageActPlane + u0, lo[0], lo[1], lo[2], lo[3]);
    st_shared_v4(row + kStageActPlane + u1, lo[4], lo[5], lo[6], lo[7]);
  }
}

// per-tile output addressing of the staged bf16 path
struct ActOut {
  __nv_bfloat16* tile_p;     // tile origin + channel window of this CTA's N tile
  long long row_stride;      // elements between vertically adjacent output pixels
  long long plane_stride;
  int c_total, up;
  int eo[4];                 // element offsets (from the tile origin) of the four pixels this lane writes back
  uint32_t vmask;            // bit j: pixel j lies inside the map
};

// write the staged chunk back: lane -> unit (lane & 3) of pixels (lane >> 2) + 8j; nunits = valid 16-byte units per pixel
template <int PLANES>
__device__ __forceinline__ void flush_act(uint32_t stage, int lane, int nunits, int ch_off, const ActOut& o) {
  const int c = lane & 3;
  if (c >= nunits) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (!((o.vmask >> j) & 1u)) continue;
    const int pix = (lane >> 2) + 8 * j;
    const uint32_t a = stage + (uint32_t)pix * 64u + ((((uint32_t)c) ^ ((uint32_t)(pix >> 1) & 3u)) << 4);
#pragma unroll
    for (int pl = 0; pl < PLANES; ++pl) {
      const uint4 v = ld_shared_v4(a + pl * kStageActPlane);
      __nv_bfloat16* dst = o.tile_p + pl * o.plane_stride + o.eo[j] + ch_off + c * 8;
      *reinterpret_cast<uint4*>(dst) = v;
      if (o.up == 2) {
        *reinterpret_cast<uint4*>(dst + o.c_total) = v;
        *reinterpret_cast<uint4*>(dst + o.row_stride) = v;
        *reinterpret_cast<uint4*>(dst + o.row_stride + o.c_total) = v;
      }
    }
  }
}

// fp32 NHWC output of one 32-column chunk (channels ch0 .. ch0+31 of [0, cout), split between out0 / out1 like
// V2X_EPI_F32_SPLIT): every lane parks its pixel's (v + bias) in the staging buffer, then lane -> unit (lane & 7) of
// pixels (lane >> 3) + 4j writes it back, 128 contiguous bytes per pixel.
template <bool HALO>
__device__ __forceinline__ void f32_chunk_out(uint32_t stage, int lane, int quad, const float* v0, const float* v1, bool two,
                                              const float* bias32, int ch0, int cout, int split, float* o0, float* o1,
                                              long long tile_pix, int oh0, int ow0, int h_out, int w_out, bool no_store) {
  const float4* b4 = reinterpret_cast<const float4*>(bias32);
  const uint32_t srow = stage + (uint32_t)lane * 128u;
  const uint32_t sw = (uint32_t)lane & 7u;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 b = b4[q];
    st_shared_v4(srow + ((((uint32_t)q) ^ sw) << 4), __float_as_uint(v0[4 * q] + b.x), __float_as_uint(v0[4 * q + 1] + b.y),
                 __float_as_uint(v0[4 * q + 2] + b.z), __float_as_uint(v0[4 * q + 3] + b.w));
  }
  if (two) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = b4[4 + q];
      st_shared_v4(srow + ((((uint32_t)(4 + q)) ^ sw) << 4), __float_as_uint(v1[4 * q] + b.x),
                   __float_as_uint(v1[4 * q + 1] + b.y), __float_as_uint(v1[4 * q + 2] + b.z),
                   __float_as_uint(v1[4 * q + 3] + b.w));
    }
  }
  __syncwarp();
  const int u = lane & 7;
  const int ch = ch0 + 4 * u;
  const int c1 = cout - split;
  if (u < (two ? 8 : 4) && ch < cout && !no_store) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int pix = (lane >> 3) + 4 * j;
      const int prow = quad * 32 + pix;
      const int ph = HALO ? (prow >> 3) : (prow >> 4), pw = HALO ? (prow & 7) : (prow & 15);
      if (oh0 + ph < h_out && ow0 + pw < w_out) {
        const uint4 v = ld_shared_v4(stage + (uint32_t)pix * 128u + ((((uint32_t)u) ^ ((uint32_t)pix & 7u)) << 4));
        const long long pi = tile_pix + (long long)ph * w_out + pw;
        float* dst = ch < split ? o0 + pi * split + ch : o1 + pi * c1 + (ch - split);
        *reinterpret_cast<uint4*>(dst) = v;
      }
    }
  }
  __syncwarp();
}

// Lean variant of f32_chunk_out for the fused heads (V2X_EPI_TAIL_F32_SPLIT): everything that does not change from tile
// to tile is hoisted into per-lane registers once per kernel -- the destination tensor / channel offset / pixel stride of
// this lane's 16-byte unit (so the cls / loc split costs no divergent branch), the unit's four bias values (added on
// the way out, so parking needs no bias loads) and the pixel walk of the write-back -- leaving per store one shared
// load, four adds and one 64-bit add.  (profiles/r01_v9: the generic routine spent 60% of the heads' epilogue time on
// address arithmetic, constant-bank loads and the split branch.)
struct TailOut {
  float* base[2];        // destination of this lane's unit in chunk 0 / 1 (tensor base + channel offset), or nullptr
  long long stride[2];   // floats per pixel of that destination
  float4 bias[2];
};

template <bool HALO>
__device__ __forceinline__ void tail_chunk_out(uint32_t stage, int lane, int quad, const float* v0, const float* v1, bool two,
                                               const TailOut& t, int c32, long long tile_pix, int oh0, int ow0, int h_out,
                                               int w_out, bool full_tile, bool no_store) {
  const uint32_t srow = stage + (uint32_t)lane * 128u;
  const uint32_t sw = (uint32_t)lane & 7u;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    st_shared_v4(srow + ((((uint32_t)q) ^ sw) << 4), __float_as_uint(v0[4 * q]), __float_as_uint(v0[4 * q + 1]),
                 __float_as_uint(v0[4 * q + 2]), __float_as_uint(v0[4 * q + 3]));
  if (two) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      st_shared_v4(srow + ((((uint32_t)(4 + q)) ^ sw) << 4), __float_as_uint(v1[4 * q]), __float_as_uint(v1[4 * q + 1]),
                   __float_as_uint(v1[4 * q + 2]), __float_as_uint(v1[4 * q + 3]));
  }
  __syncwarp();
  float* const base = t.base[c32];
  if (base != nullptr && !no_store) {
    const uint32_t u = (uint32_t)lane & 7u;
    const int l3 = lane >> 3;
    const long long st = t.stride[c32];
    const float4 b = t.bias[c32];
    // pixel (l3 + 4j) of this warp's 32: tile row / column of the write-back walk
    constexpr int ROWS_PER_QUAD = HALO ? 4 : 2;
    float* const p0 = base + (tile_pix + (long long)(quad * ROWS_PER_QUAD) * w_out + l3) * st;
    const long long row_step = (long long)w_out * st;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int dh = HALO ? (j >> 1) : (j >> 2);        // (l3 + 4j) / TILE_W
      const int dw = HALO ? 4 * (j & 1) : 4 * (j & 3);  // (l3 + 4j) % TILE_W - l3
      if (!full_tile) {
        const int ph = quad * ROWS_PER_QUAD + dh, pw = dw + l3;
        if (oh0 + ph >= h_out || ow0 + pw >= w_out) continue;
      }
      const uint32_t pix = (uint32_t)(l3 + 4 * j);
      uint4 v = ld_shared_v4(stage + pix * 128u + ((u ^ (pix & 7u)) << 4));
      v.x = __float_as_uint(__uint_as_float(v.x) + b.x);
      v.y = __float_as_uint(__uint_as_float(v.y) + b.y);
      v.z = __float_as_uint(__uint_as_float(v.z) + b.z);
      v.w = __float_as_uint(__uint_as_float(v.w) + b.w);
      *reinterpret_cast<uint4*>(p0 + dh * row_step + dw * st) = v;
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// Tensor-core kernel: persistent, warp-specialised (192 threads)
//   warp 0 : TMA producer (A boxes every stage; weights either streamed with A or loaded once per CTA
//            when the whole [BN x K] operand fits in shared memory)
//   warp 1 : tcgen05.mma issuer; alternates between two TMEM accumulator buffers
//   warps 2..5 : epilogue (tcgen05.ld -> bias/ReLU/gates -> global); warp w owns TMEM lane quadrant
//            w % 4, so tile row = (w % 4) * 32 + lane
// grid = (ctas_x, n_tiles): a CTA keeps its N tile (blockIdx.y) and walks M tiles
// blockIdx.x, blockIdx.x + gridDim.x, ...  The epilogue of tile i overlaps the main loop of tile
// i+1 through the two accumulator buffers (tmem_full / tmem_empty mbarriers).
//
// The role loops are deliberately lean: a tile of a C=32 layer is only ~700 cycles of tensor work,
// so every role must spend well under that per tile.  Hence: the k-block -> (source, channel offset,
// tap shift) decode is tabulated in shared memory once per CTA; tiles are walked incrementally (no
// divisions); KSTEPS (= kc/16) is a template parameter so the MMA burst is straight-line code; bias
// vectors live in shared memory; and all loops run warp-uniformly with elect_one() only around the
// issuing instructions (running them under `if (lane == 0)` makes the compiler wrap every
// tcgen05.mma / TMA in a divergence waterfall, ~200 cycles per instruction -- measured).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxKBlocks = 160;

struct TileIter {
  int tile, n_img, th, tw;   // current M tile and its decode
  int d_n, d_th, d_tw;       // decode of the grid stride
  int tiles_w, tiles_h;
  __device__ __forceinline__ void init(const ConvDev& p, int first, int stride) {
    tiles_w = p.tiles_w;
    tiles_h = p.tiles_per_img / p.tiles_w;
    tile = first;
    n_img = first / p.tiles_per_img;
    int r = first - n_img * p.tiles_per_img;
    th = r / tiles_w;
    tw = r - th * tiles_w;
    d_n = stride / p.tiles_per_img;
    r = stride - d_n * p.tiles_per_img;
    d_th = r / tiles_w;
    d_tw = r - d_th * tiles_w;
  }
  __device__ __forceinline__ void next(int stride) {
    tile += stride;
    tw += d_tw;
    if (tw >= tiles_w) { tw -= tiles_w; ++th; }
    th += d_th;
    if (th >= tiles_h) { th -= tiles_h; ++n_img; }
    n_img += d_n;
  }
};

// HALO = true (3x3, stride 1, resident weights): the output tile is 16 rows x 8 px and the A operand of ALL nine
// taps comes from ONE TMA box per channel block -- the 18 x 10 px halo of the tile.  Tap (kh, kw) is read by
// shifting the UMMA descriptor's start address by (kh*10 + kw) pixel rows; the 8-row groups of the operand are
// one halo row (10 px) apart, hence SBO = 10 * kc*2 bytes.  This relies on the tensor core applying the swizzle
// to absolute shared-memory address bits (so unaligned starts are fine) -- verified bit-exactly on B200 for the
// 32/64/128-byte swizzles with tools/halo_probe.cu.  9x fewer TMA boxes, 6.4x fewer L2->smem bytes.
constexpr int kHaloH = 18, kHaloW = 10;

template <int BN, int PLANES, int MMAS, int KSTEPS, bool HALO>
__global__ void __launch_bounds__(kNumThreads, BN <= 64 ? 2 : 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                                 const __grid_constant__ CUtensorMap tmA1,
                                                                 const __grid_constant__ CUtensorMap tmB,
                                                                 const __grid_constant__ CUtensorMap tmT,
                                                                 const ConvDev p) {
  constexpr int KC = 16 * KSTEPS;
  // PLANES = storage planes of the act tensors (1: bf16; 2: fp16 hi/lo).  MMAS = tensor-core passes per k-step:
  // 3: hi*hi + hi*w_lo + a_lo*hi (both operands split); 2: hi*hi + a_lo*hi (weights are one fp16 plane);
  // 1: hi*hi only (the lo plane of the input is not even loaded).  The output is always written in PLANES planes.
  constexpr int PA = (PLANES == 2 && MMAS >= 2) ? 2 : 1;
  constexpr int PW = (PLANES == 2 && MMAS == 3) ? 2 : 1;
  constexpr uint32_t ACC_STRIDE = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  constexpr uint32_t TMEM_COLS = BN == 64 ? 256u : 2 * ACC_STRIDE;  // BN == 64: + the fused tail's accumulator (cols 128..191)
  constexpr uint32_t A_TILE = 128u * KC * 2u;                             // bytes
  constexpr uint32_t B_TILE = ((uint32_t)BN * KC * 2u + 1023u) & ~1023u;  // bytes (1 KB aligned)
  constexpr uint32_t ROW = KC * 2u;                                      // bytes of one pixel's channel block
  constexpr uint32_t SBO = 8u * ROW;
  constexpr uint32_t A_HALO = ((uint32_t)(kHaloH * kHaloW) * ROW + 1023u) & ~1023u;
  constexpr uint32_t A_BLOCK = HALO ? A_HALO : A_TILE;                   // smem bytes of one k-block's A operand (per plane)
  constexpr uint32_t LAYOUT = KC == 64 ? 2u : KC == 32 ? 4u : 6u;
  constexpr int TILE_H = HALO ? 16 : kTileH, TILE_W = HALO ? 8 : kTileW;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[4 * kMaxStages + 6];
  __shared__ uint32_t tmem_slot;
  __shared__ int4 ktab[kMaxKBlocks];  // per k-block: {channel coord, dw, dh, src | hp << 1}
  __shared__ __align__(16) float s_bias[BN];
  __shared__ __align__(16) float s_bhn[64];   // GRU: b_hh_n; fused tail: its bias

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BN;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // resident weights first, then the stage ring
  const uint32_t ring_base = smem_base + p.b_region_bytes;
  const uint32_t bar_full = smem_u32(&bars[0]);
  const uint32_t bar_empty = smem_u32(&bars[kMaxStages]);
  const uint32_t bar_bres = smem_u32(&bars[2 * kMaxStages]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * kMaxStages + 1]);   // [2]
  const uint32_t bar_tempty = smem_u32(&bars[2 * kMaxStages + 3]);  // [2]
  const uint32_t bar_bfull = smem_u32(&bars[2 * kMaxStages + 5]);   // weight ring of the halo + streamed-B mode
  const uint32_t bar_bempty = smem_u32(&bars[3 * kMaxStages + 5]);
  const uint32_t bar_tail = smem_u32(&bars[4 * kMaxStages + 5]);    // fused tail GEMM finished
  const bool has_tail = p.epilogue == V2X_EPI_TAIL_F32_SPLIT;
  const uint32_t bring_base = smem_base + p.b_ring_off;
  const bool halo_stream = HALO && !p.b_resident;

  // ---- one-time setup ----
  for (int k = threadIdx.x; k < p.num_k; k += kNumThreads) {
    int4 e;
    if (HALO) {  // k-block = (source, channel block); e = {channel coord, first weight tile, weight tile step per tap, source}
      const int s = k < p.cblocks[0] ? 0 : 1;
      const int cb = k - s * p.cblocks[0];
      e = make_int4(cb * KC, s * 9 * p.cblocks[0] + cb, p.cblocks[s], s);
      // GRU pre-activation window: channels [n0, n0 + BN) of src[1], centre tap only, identity weight tile (bit 1 of e.w)
      if (s == 1 && p.gru_pre_act) e = make_int4(n0 + cb * KC, 9 * p.cblocks[0] + cb, 0, 3);
    } else {
      const int k0 = p.taps * p.cblocks[0];
      const int s = k < k0 ? 0 : 1;
      const int kr = k - s * k0;
      const int tap = kr / p.cblocks[s], cb = kr - tap * p.cblocks[s];
      const int kh = p.taps == 9 ? tap / 3 : 1, kw = p.taps == 9 ? tap - 3 * (tap / 3) : 1;
      if (p.stride == 1) {
        e = make_int4(cb * KC, kw - 1, kh - 1, s);
      } else {  // input row 2*oh + kh - 1 = 2*(oh + hoff) + hp, same along w
        const int hp = kh == 1 ? 0 : 1, hoff = kh == 0 ? -1 : 0;
        const int wp = kw == 1 ? 0 : 1, woff = kw == 0 ? -1 : 0;
        e = make_int4(wp * p.cin[s] + cb * KC, woff, hoff, s | (hp << 1));
      }
    }
    ktab[k] = e;
  }
  for (int i = threadIdx.x; i < BN; i += kNumThreads) s_bias[i] = p.bias[n0 + i];
  if (p.epilogue == V2X_EPI_GRU && threadIdx.x < 64) s_bhn[threadIdx.x] = p.gru_bhn[blockIdx.y * 64 + threadIdx.x];
  if (has_tail && threadIdx.x < 64) s_bhn[threadIdx.x] = threadIdx.x < p.tail_cout_pad ? p.tail_bias[threadIdx.x] : 0.f;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA0);
    if (p.nsrc > 1) prefetch_tmap(&tmA1);
    prefetch_tmap(&tmB);
    if (has_tail) prefetch_tmap(&tmT);
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_bres, 1);
    mbar_init(bar_tail, 1);
    for (int s = 0; s < p.b_stages; ++s) {
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 4);  // one arrive per epilogue warp
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
  const bool is_gru = p.epilogue == V2X_EPI_GRU;
  const int grid_stride = gridDim.x;
  pdl_launch_dependents();   // the next layer's CTAs may take over this SM as soon as this CTA exits (see common.cuh)

  if (warp == 0) {
    // ===== TMA producer =====
    if (p.b_resident && elect_one()) {
      const uint32_t tail_bytes = has_tail ? (uint32_t)PLANES * (uint32_t)p.tail_cout_pad * 128u : 0u;
      mbar_expect_tx(bar_bres, (uint32_t)p.num_b_tiles * PW * (uint32_t)(BN * KC * 2) + tail_bytes);
      if (has_tail) {
#pragma unroll
        for (int pl = 0; pl < PLANES; ++pl)
          tma_load_2d(smem_base + p.tail_w_off + pl * 8192u, &tmT, bar_bres, 0, pl * p.tail_cout_pad);
      }
      for (int k = 0; k < p.num_b_tiles; ++k)
#pragma unroll
        for (int pl = 0; pl < PW; ++pl)
          tma_load_2d(smem_base + (k * PW + pl) * B_TILE, &tmB, bar_bres, k * KC, pl * p.cout_pad + n0);
    }
    __syncwarp();
    pdl_wait();   // weights above are constants; the activations below were written by the previous kernel(s)
    const bool no_tma = p.debug_mode == 2;
    int stage = 0, phase = 0, bstage = 0, bphase = 0;
    TileIter ti;
    ti.init(p, blockIdx.x, grid_stride);
    for (; ti.tile < p.m_tiles; ti.next(grid_stride)) {
      if (is_gru && gru_unit_absent(p, ti.n_img)) continue;
      const int oh0 = ti.th * TILE_H, ow0 = ti.tw * TILE_W;
      int kidx = 0;
      if (halo_stream) {
        // two rings: an A ring of halo tiles (one per channel block) and a weight ring fed tap by tap
        for (int kb = 0; kb < p.num_k; ++kb) {
          const int4 e = ktab[kb];
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          if (elect_one()) {
            const uint32_t full = bar_full + 8 * stage;
            if (no_tma) {
              mbar_arrive(full);
            } else {
              mbar_expect_tx(full, p.kb_tx_bytes);
              const uint32_t sa = ring_base + stage * p.stage_bytes;
#pragma unroll
              for (int pl = 0; pl < PA; ++pl)
                tma_load_4d(sa + pl * A_BLOCK, (e.w & 1) ? &tmA1 : &tmA0, full, e.x, ow0 - 1, oh0 - 1, pl * p.n_maps + ti.n_img);
            }
          }
          __syncwarp();
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
          const bool single = (e.w & 2) != 0;   // identity block: one weight tile (tap step 0), centre tap only
          const int ntaps = single ? 1 : 9, tstep = single ? 1 : p.b_taps;
          for (int t0 = 0; t0 < ntaps; t0 += tstep) {
            mbar_wait(bar_bempty + 8 * bstage, bphase ^ 1);
            if (elect_one()) {
              const uint32_t bfull = bar_bfull + 8 * bstage;
              if (no_tma) {
                mbar_arrive(bfull);
              } else {
                mbar_expect_tx(bfull, (uint32_t)tstep * PW * (uint32_t)(BN * KC * 2));
                uint32_t sb = bring_base + bstage * p.b_stage_bytes;
                for (int t = t0; t < t0 + tstep; ++t)
#pragma unroll
                  for (int pl = 0; pl < PW; ++pl, sb += B_TILE)
                    tma_load_2d(sb, &tmB, bfull, (e.y + t * e.z) * KC, pl * p.cout_pad + n0);
              }
            }
            __syncwarp();
            if (++bstage == p.b_stages) { bstage = 0; bphase ^= 1; }
          }
        }
        continue;
      }
      for (int ks = 0; ks < p.stages_per_tile; ++ks) {
        const int nblk = min(p.kb_per_stage, p.num_k - kidx);
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        const uint32_t full = bar_full + 8 * stage;
        if (elect_one()) {
          if (no_tma) {
            mbar_arrive(full);
          } else {
            mbar_expect_tx(full, nblk * p.kb_tx_bytes);
            uint32_t sa = ring_base + stage * p.stage_bytes;
            for (int j = 0; j < nblk; ++j, sa += p.kb_bytes) {
              const int4 e = ktab[kidx + j];
              const CUtensorMap* tmA = (e.w & 1) ? &tmA1 : &tmA0;
#pragma unroll
              for (int pl = 0; pl < PA; ++pl) {
                const int img = pl * p.n_maps + ti.n_img;
                if (HALO) tma_load_4d(sa + pl * A_BLOCK, (e.w & 1) ? &tmA1 : &tmA0, full, e.x, ow0 - 1, oh0 - 1, img);
                else if (p.stride == 1) tma_load_4d(sa + pl * A_BLOCK, tmA, full, e.x, ow0 + e.y, oh0 + e.z, img);
                else tma_load_5d(sa + pl * A_BLOCK, tmA, full, e.x, ow0 + e.y, e.w >> 1, oh0 + e.z, img);
              }
              if (!HALO && !p.b_resident) {
#pragma unroll
                for (int pl = 0; pl < PW; ++pl)
                  tma_load_2d(sa + PA * A_TILE + pl * B_TILE, &tmB, full, (kidx + j) * KC, pl * p.cout_pad + n0);
              }
            }
          }
        }
        __syncwarp();
        kidx += nblk;
        if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_m128<PLANES>(BN);
    if (p.b_resident) mbar_wait(bar_bres, 0);
    const bool no_mma = p.debug_mode == 1;
    int stage = 0, phase = 0, it = 0, bstage = 0, bphase = 0;
    // descriptors differ only in the 14-bit start-address field (units of 16 bytes)
    const uint64_t desc_ring = make_smem_desc(ring_base, HALO ? kHaloW * ROW : SBO, LAYOUT);
    const uint64_t desc_bres = make_smem_desc(smem_base, SBO, LAYOUT);
    const uint64_t desc_bring = make_smem_desc(bring_base, SBO, LAYOUT);
    const uint32_t stage16 = p.stage_bytes >> 4, kb16 = p.kb_bytes >> 4, bstage16 = p.b_stage_bytes >> 4;
    TileIter ti;
    ti.init(p, blockIdx.x, grid_stride);
    for (; ti.tile < p.m_tiles; ti.next(grid_stride)) {
      if (is_gru && gru_unit_absent(p, ti.n_img)) continue;
      const int acc_buf = it & 1;
      mbar_wait(bar_tempty + 8 * acc_buf, ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc_buf * ACC_STRIDE;
      uint32_t acc = 0;
      int kidx = 0;
      if (halo_stream) {
        // Straight-line issue code: the tap loop is unrolled for both weight-ring groupings so every descriptor offset is
        // an immediate (with runtime tap bounds the warp spent ~12 instructions per MMA on t/3, t%3 and 64-bit address
        // arithmetic and was issue-bound on the N<=128 layers -- profiles/r01_v8).
        for (int kb = 0; kb < p.num_k; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);           // halo tile of this channel block has landed
          const uint64_t da = desc_ring + (uint64_t)(stage * stage16);
          auto issue_taps = [&](auto B_TAPS_C) {
            constexpr int B_TAPS = decltype(B_TAPS_C)::value;
#pragma unroll
            for (int t0 = 0; t0 < 9; t0 += B_TAPS) {
              mbar_wait(bar_bfull + 8 * bstage, bphase);    // weight tiles of the next B_TAPS taps
              tc_fence_after();
              if (elect_one()) {
                if (no_mma) {
                  mbar_arrive(bar_bempty + 8 * bstage);
                } else {
                  const uint64_t db = desc_bring + (uint64_t)(bstage * bstage16);
#pragma unroll
                  for (int tt = 0; tt < B_TAPS; ++tt) {
                    const int t = t0 + tt;
                    const uint32_t a_off = (uint32_t)((((t / 3) * kHaloW + (t % 3)) * ROW) >> 4);
                    const uint32_t b_off = (uint32_t)(tt * PW * (B_TILE >> 4));
#pragma unroll
                    for (int kk = 0; kk < KSTEPS; ++kk) {
                      umma_bf16(tmem_d, da + (a_off + 2 * kk), db + (b_off + 2 * kk), idesc, (t | kk) == 0 ? acc : 1u);
                      if (PW == 2) umma_bf16(tmem_d, da + (a_off + 2 * kk), db + (b_off + 2 * kk + (B_TILE >> 4)), idesc, 1);
                      if (PA == 2) umma_bf16(tmem_d, da + (a_off + 2 * kk + (A_BLOCK >> 4)), db + (b_off + 2 * kk), idesc, 1);
                    }
                  }
                  umma_commit(bar_bempty + 8 * bstage);
                }
              }
              __syncwarp();
              if (++bstage == p.b_stages) { bstage = 0; bphase ^= 1; }
            }
          };
          if (ktab[kb].w & 2) {
            // identity block (GRU pre-activations): centre tap of the halo tile x one weight tile
            mbar_wait(bar_bfull + 8 * bstage, bphase);
            tc_fence_after();
            if (elect_one()) {
              if (no_mma) {
                mbar_arrive(bar_bempty + 8 * bstage);
              } else {
                const uint64_t db = desc_bring + (uint64_t)(bstage * bstage16);
                constexpr uint32_t a_off = (uint32_t)(((kHaloW + 1) * ROW) >> 4);
#pragma unroll
                for (int kk = 0; kk < KSTEPS; ++kk) {
                  umma_bf16(tmem_d, da + (a_off + 2 * kk), db + 2 * kk, idesc, kk == 0 ? acc : 1u);
                  if (PW == 2) umma_bf16(tmem_d, da + (a_off + 2 * kk), db + (2 * kk + (B_TILE >> 4)), idesc, 1);
                  if (PA == 2) umma_bf16(tmem_d, da + (a_off + 2 * kk + (A_BLOCK >> 4)), db + 2 * kk, idesc, 1);
                }
                umma_commit(bar_bempty + 8 * bstage);
              }
            }
            __syncwarp();
            if (++bstage == p.b_stages) { bstage = 0; bphase ^= 1; }
          } else if (p.b_taps == 3) issue_taps(std::integral_constant<int, 3>{});
          else issue_taps(std::integral_constant<int, 1>{});
          acc = 1;
          if (elect_one()) {
            if (no_mma) mbar_arrive(bar_empty + 8 * stage);
            else umma_commit(bar_empty + 8 * stage);         // all nine taps of this halo tile are issued
          }
          __syncwarp();
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
      } else
      for (int ks = 0; ks < p.stages_per_tile; ++ks) {
        const int nblk = min(p.kb_per_stage, p.num_k - kidx);
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        if (elect_one()) {
          if (no_mma) {
            mbar_arrive(bar_empty + 8 * stage);
          } else {
            uint64_t da = desc_ring + (uint64_t)(stage * stage16);
            if (HALO) {
              for (int j = 0; j < nblk; ++j, da += kb16) {
                const int4 e = ktab[kidx + j];
                uint64_t db = desc_bres + (uint64_t)(e.y * PW * (B_TILE >> 4));
                const uint32_t db_step = e.z * PW * (B_TILE >> 4);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap, db += db_step) {
                  const uint64_t dat = da + (uint64_t)((((tap / 3) * kHaloW + (tap % 3)) * ROW) >> 4);
#pragma unroll
                  for (int kk = 0; kk < KSTEPS; ++kk) {
                    umma_bf16(tmem_d, dat + 2 * kk, db + 2 * kk, idesc, acc);
                    acc = 1;
                    if (PW == 2) umma_bf16(tmem_d, dat + 2 * kk, db + 2 * kk + (B_TILE >> 4), idesc, 1);
                    if (PA == 2) umma_bf16(tmem_d, dat + 2 * kk + (A_BLOCK >> 4), db + 2 * kk, idesc, 1);
                  }
                }
              }
            } else {
              uint64_t db = p.b_resident ? desc_bres + (uint64_t)(kidx * PW * (B_TILE >> 4))
                                         : da + (uint64_t)(PA * (A_TILE >> 4));
              const uint32_t db_step = p.b_resident ? PW * (B_TILE >> 4) : kb16;
              for (int j = 0; j < nblk; ++j, da += kb16, db += db_step) {
#pragma unroll
                for (int kk = 0; kk < KSTEPS; ++kk) {
                  umma_bf16(tmem_d, da + 2 * kk, db + 2 * kk, idesc, acc);
                  acc = 1;
                  if (PW == 2) umma_bf16(tmem_d, da + 2 * kk, db + 2 * kk + (B_TILE >> 4), idesc, 1);
                  if (PA == 2) umma_bf16(tmem_d, da + 2 * kk + (A_TILE >> 4), db + 2 * kk, idesc, 1);
                }
              }
            }
            umma_commit(bar_empty + 8 * stage);
          }
        }
        __syncwarp();
        acc = 1;
        kidx += nblk;
        if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) {
        if (no_mma) mbar_arrive(bar_tfull + 8 * acc_buf);
        else umma_commit(bar_tfull + 8 * acc_buf);
      }
      __syncwarp();
      ++it;
    }
  } else {
    // ===== epilogue warps =====
    pdl_wait();   // before the first global store / passthrough read (cheap: nothing to do until the first tile is done)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int r_h = HALO ? (row >> 3) : (row >> 4), r_w = HALO ? (row & 7) : (row & 15);
    const bool dbg_no_store = p.debug_mode == 3;
    const int up = p.upsample2x ? 2 : 1;
    const int Hs = p.h_out * up, Ws = p.w_out * up;
    const uint32_t stage = smem_base + p.epi_off + (uint32_t)(warp - 2) * p.epi_warp_bytes;  // this warp's staging buffer
    // write-back roles (kernel constants): bf16 chunks -> pixels (lane >> 2) + 8j; fp32 chunks -> pixels (lane >> 3) + 4j
    ActOut o;
    o.up = up;
    o.c_total = p.out_c_total;
    o.row_stride = (long long)Ws * p.out_c_total;
    o.plane_stride = p.out_plane_stride;
    int wb_h[4], wb_w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int prow = quad * 32 + (lane >> 2) + 8 * j;
      wb_h[j] = HALO ? (prow >> 3) : (prow >> 4);
      wb_w[j] = HALO ? (prow & 7) : (prow & 15);
      o.eo[j] = (wb_h[j] * up * Ws + wb_w[j] * up) * p.out_c_total;
    }
    TailOut tout;
    if (has_tail) {
      const int c1 = p.tail_cout - p.split;
#pragma unroll
      for (int c32 = 0; c32 < 2; ++c32) {
        const int ch = c32 * 32 + 4 * (lane & 7);
        const bool act = ch < p.tail_cout;
        tout.base[c32] = !act ? nullptr
                              : ch < p.split ? reinterpret_cast<float*>(p.out0) + ch
                                             : reinterpret_cast<float*>(p.out1) + (ch - p.split);
        tout.stride[c32] = ch < p.split ? p.split : c1;
        tout.bias[c32] = act ? *reinterpret_cast<const float4*>(s_bhn + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    int it = 0;
    TileIter ti;
    ti.init(p, blockIdx.x, grid_stride);
    for (; ti.tile < p.m_tiles; ti.next(grid_stride)) {
      const int oh0 = ti.th * TILE_H, ow0 = ti.tw * TILE_W;
      const int oh = oh0 + r_h, ow = ow0 + r_w;
      const bool valid = oh < p.h_out && ow < p.w_out;   // partial tiles at the map border
      if (is_gru && gru_unit_absent(p, ti.n_img)) {
        if (valid) copy_passthrough(p, ti.n_img, oh, ow, blockIdx.y * 64, 64);
        continue;
      }
      const int acc_buf = it & 1;
      mbar_wait(bar_tfull + 8 * acc_buf, (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc_buf * ACC_STRIDE + ((uint32_t)(quad * 32) << 16);
      if (p.debug_mode == 4) {
        // profiling ablation: barrier handshake only (no tcgen05.ld, no math, no stores)
      } else if (p.epilogue == V2X_EPI_F32_NCHW) {
#pragma unroll 1
        for (int c16 = 0; c16 < BN / 16; ++c16) {
          if (n0 + c16 * 16 >= p.cout) break;
          float v0[16];
          tmem_ld16_async(taddr + c16 * 16, v0);
          tmem_ld_wait16(v0);
          if (valid && !dbg_no_store) epi_f32_nchw16(p, ti.n_img, oh, ow, n0 + c16 * 16, v0, s_bias + c16 * 16);
        }
      } else if (p.epilogue == V2X_EPI_F32_SPLIT) {
        // fp32 NHWC heads / gate pre-activations: 32-column chunks staged as [pixel][32 floats]
        const long long tile_pix = ((long long)ti.n_img * p.h_out + oh0) * p.w_out + ow0;
#pragma unroll 1
        for (int c32 = 0; c32 < (BN + 31) / 32; ++c32) {
          const int ch0 = n0 + c32 * 32;
          if (ch0 >= p.cout) break;
          const bool two = (c32 * 32 + 16 < BN) && (ch0 + 16 < p.cout);
          float v0[16], v1[16];
          tmem_ld16_async(taddr + c32 * 32, v0);
          if (two) tmem_ld16_async(taddr + c32 * 32 + 16, v1);
          tmem_ld_wait16(v0);
          if (two) tmem_ld_wait16(v1);
          f32_chunk_out<HALO>(stage, lane, quad, v0, v1, two, s_bias + c32 * 32, ch0, p.cout, p.split,
                              reinterpret_cast<float*>(p.out0), reinterpret_cast<float*>(p.out1), tile_pix, oh0, ow0,
                              p.h_out, p.w_out, dbg_no_store);
        }
      } else if (has_tail) {
        if constexpr (BN == 64) {
          // ---- fused 1x1 tail: relu(acc + bias) -> bf16 -> A2 tile in smem -> second GEMM -> fp32 split output ----
          // phase 1: this warp's 32 pixels x 64 channels into the K-major SWIZZLE_128B A2 tile (row = TMEM lane)
          const uint32_t a2 = smem_base + p.tail_a_off;
          const uint32_t a2row = a2 + (uint32_t)row * 128u;
          const uint32_t rsw = (uint32_t)row & 7u;
#pragma unroll
          for (int c32 = 0; c32 < 2; ++c32) {
            float v0[16], v1[16];
            tmem_ld16_async(taddr + c32 * 32, v0);
            tmem_ld16_async(taddr + c32 * 32 + 16, v1);
            tmem_ld_wait16(v0);
            tmem_ld_wait16(v1);
            const float4* b4 = reinterpret_cast<const float4*>(s_bias + c32 * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b0 = b4[q], b1 = b4[4 + q];
              v0[4 * q + 0] = fmaxf(v0[4 * q + 0] + b0.x, 0.f); v0[4 * q + 1] = fmaxf(v0[4 * q + 1] + b0.y, 0.f);
              v0[4 * q + 2] = fmaxf(v0[4 * q + 2] + b0.z, 0.f); v0[4 * q + 3] = fmaxf(v0[4 * q + 3] + b0.w, 0.f);
              v1[4 * q + 0] = fmaxf(v1[4 * q + 0] + b1.x, 0.f); v1[4 * q + 1] = fmaxf(v1[4 * q + 1] + b1.y, 0.f);
              v1[4 * q + 2] = fmaxf(v1[4 * q + 2] + b1.z, 0.f); v1[4 * q + 3] = fmaxf(v1[4 * q + 3] + b1.w, 0.f);
            }
            uint32_t hi[8], lo[8];
            pack16<PLANES>(v0, hi, lo);
            st_shared_v4(a2row + ((((uint32_t)(4 * c32 + 0)) ^ rsw) << 4), hi[0], hi[1], hi[2], hi[3]);
            st_shared_v4(a2row + ((((uint32_t)(4 * c32 + 1)) ^ rsw) << 4), hi[4], hi[5], hi[6], hi[7]);
            if (PLANES == 2) {
              st_shared_v4(a2row + 16384u + ((((uint32_t)(4 * c32 + 0)) ^ rsw) << 4), lo[0], lo[1], lo[2], lo[3]);
              st_shared_v4(a2row + 16384u + ((((uint32_t)(4 * c32 + 1)) ^ rsw) << 4), lo[4], lo[5], lo[6], lo[7]);
            }
            pack16<PLANES>(v1, hi, lo);
            st_shared_v4(a2row + ((((uint32_t)(4 * c32 + 2)) ^ rsw) << 4), hi[0], hi[1], hi[2], hi[3]);
            st_shared_v4(a2row + ((((uint32_t)(4 * c32 + 3)) ^ rsw) << 4), hi[4], hi[5], hi[6], hi[7]);
            if (PLANES == 2) {
              st_shared_v4(a2row + 16384u + ((((uint32_t)(4 * c32 + 2)) ^ rsw) << 4), lo[0], lo[1], lo[2], lo[3]);
              st_shared_v4(a2row + 16384u + ((((uint32_t)(4 * c32 + 3)) ^ rsw) << 4), lo[4], lo[5], lo[6], lo[7]);
            }
          }
          // the conv accumulator is drained: hand it back to the MMA warp before the tail GEMM
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * acc_buf);
          fence_proxy_async_smem();                                  // A2 (generic-proxy writes) -> visible to the tensor core
          asm volatile("bar.sync 1, 128;" ::: "memory");            // all four epilogue warps have written their rows
          const uint32_t tmem_t = tmem_base + 128u;                  // tail accumulator: columns 128 .. 128+tail_cout_pad
          if (warp == 2) {
            tc_fence_after();
            if (elect_one()) {
              const uint32_t idesc_t = make_idesc_m128<PLANES>((uint32_t)p.tail_cout_pad);
              const uint64_t da2 = make_smem_desc(a2, 1024u, 2u);
              const uint64_t dw2 = make_smem_desc(smem_base + p.tail_w_off, 1024u, 2u);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                umma_bf16(tmem_t, da2 + 2 * kk, dw2 + 2 * kk, idesc_t, kk == 0 ? 0u : 1u);
                if (PLANES == 2) {
                  umma_bf16(tmem_t, da2 + 2 * kk, dw2 + 2 * kk + (8192u >> 4), idesc_t, 1);
                  umma_bf16(tmem_t, da2 + 2 * kk + (16384u >> 4), dw2 + 2 * kk, idesc_t, 1);
                }
              }
              umma_commit(bar_tail);
            }
            __syncwarp();
          }
          mbar_wait(bar_tail, it & 1);
          tc_fence_after();
          // phase 2: tail accumulator + bias -> fp32 NHWC split output; staging reuses this warp's own A2 rows
          const uint32_t tstage = a2 + (uint32_t)(quad * 32) * 128u;
          const uint32_t taddr2 = tmem_t + ((uint32_t)(quad * 32) << 16);
          const long long tile_pix = ((long long)ti.n_img * p.h_out + oh0) * p.w_out + ow0;
          const bool full_tile = oh0 + TILE_H <= p.h_out && ow0 + TILE_W <= p.w_out;
#pragma unroll
          for (int c32 = 0; c32 < 2; ++c32) {
            if (c32 * 32 >= p.tail_cout) break;
            const bool two = c32 * 32 + 16 < p.tail_cout;
            float v0[16], v1[16];
            tmem_ld16_async(taddr2 + c32 * 32, v0);
            if (two) tmem_ld16_async(taddr2 + c32 * 32 + 16, v1);
            tmem_ld_wait16(v0);
            if (two) tmem_ld_wait16(v1);
            tail_chunk_out<HALO>(tstage, lane, quad, v0, v1, two, tout, c32, tile_pix, oh0, ow0, p.h_out, p.w_out,
                                 full_tile, dbg_no_store);
          }
          tc_fence_before();   // orders this tile's tcgen05.ld of the tail accumulator before the next tile's barrier + MMA
          ++it;
          continue;
        }
      } else {
        o.tile_p = reinterpret_cast<__nv_bfloat16*>(p.out0) +
                   (((long long)ti.n_img * Hs + oh0 * up) * Ws + ow0 * up) * p.out_c_total + p.out_c_off +
                   (is_gru ? blockIdx.y * 64 : n0);
        o.vmask = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (oh0 + wb_h[j] < p.h_out && ow0 + wb_w[j] < p.w_out && !dbg_no_store) o.vmask |= 1u << j;
        if (is_gru) {
          if constexpr (BN == 192) {
#pragma unroll 1
            for (int c16 = 0; c16 < 4; ++c16) {
              float r[16], z[16], nn[16];
              // round-invariant half of the pre-activations (W_ih[:, mean] * mean + b): issue the global loads first so
              // their latency overlaps the TMEM reads
              float4 gr[4], gz[4], gn[4];
              const bool has_add = p.gru_add != nullptr && valid;
              if (has_add) {
                const float4* g = reinterpret_cast<const float4*>(
                    p.gru_add + (((long long)ti.n_img * p.h_out + oh) * p.w_out + ow) * p.cout + n0 + c16 * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) { gr[q] = __ldg(g + q); gz[q] = __ldg(g + 16 + q); gn[q] = __ldg(g + 32 + q); }
              }
              tmem_ld16_async(taddr + c16 * 16, r);
              tmem_ld16_async(taddr + 64 + c16 * 16, z);
              tmem_ld16_async(taddr + 128 + c16 * 16, nn);
              tmem_ld_wait16(r);
              tmem_ld_wait16(z);
              tmem_ld_wait16(nn);
              const float* br = s_bias + c16 * 16;
              if (has_add) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  r[4 * q] += gr[q].x; r[4 * q + 1] += gr[q].y; r[4 * q + 2] += gr[q].z; r[4 * q + 3] += gr[q].w;
                  z[4 * q] += gz[q].x; z[4 * q + 1] += gz[q].y; z[4 * q + 2] += gz[q].z; z[4 * q + 3] += gz[q].w;
                  nn[4 * q] += gn[q].x; nn[4 * q + 1] += gn[q].y; nn[4 * q + 2] += gn[q].z; nn[4 * q + 3] += gn[q].w;
                }
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float rr = fast_sigmoid(r[i] + br[i]);
                const float zz = fast_sigmoid(z[i] + br[64 + i]);
                const float nv = fast_tanh(nn[i] + br[128 + i] + rr * s_bhn[c16 * 16 + i]);
                r[i] = (1.f - zz) * nv;
              }
              uint32_t hi[8], lo[8];
              pack16<PLANES>(r, hi, lo);
              stage_act16<PLANES>(stage, lane, c16 & 1, hi, lo);
              if (c16 & 1) {
                __syncwarp();
                flush_act<PLANES>(stage, lane, 4, (c16 - 1) * 16, o);
                __syncwarp();
              }
            }
          }
        } else {
          const bool relu = p.relu != 0;
#pragma unroll 1
          for (int c32 = 0; c32 < (BN + 31) / 32; ++c32) {
            if (n0 + c32 * 32 >= p.cout) break;
            const bool two = (c32 * 32 + 16 < BN) && (n0 + c32 * 32 + 16 < p.cout);
            float v0[16], v1[16];
            tmem_ld16_async(taddr + c32 * 32, v0);
            if (two) tmem_ld16_async(taddr + c32 * 32 + 16, v1);
            tmem_ld_wait16(v0);
            if (two) tmem_ld_wait16(v1);
            const float4* b4 = reinterpret_cast<const float4*>(s_bias + c32 * 32);
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b = b4[q];
              v0[4 * q + 0] += b.x; v0[4 * q + 1] += b.y; v0[4 * q + 2] += b.z; v0[4 * q + 3] += b.w;
            }
            if (relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v0[i] = fmaxf(v0[i], 0.f);
            }
            pack16<PLANES>(v0, hi, lo);
            stage_act16<PLANES>(stage, lane, 0, hi, lo);
            if (two) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 b = b4[4 + q];
                v1[4 * q + 0] += b.x; v1[4 * q + 1] += b.y; v1[4 * q + 2] += b.z; v1[4 * q + 3] += b.w;
              }
              if (relu) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v1[i] = fmaxf(v1[i], 0.f);
              }
              pack16<PLANES>(v1, hi, lo);
              stage_act16<PLANES>(stage, lane, 1, hi, lo);
            }
            __syncwarp();
            flush_act<PLANES>(stage, lane, two ? 4 : 2, c32 * 32, o);
            __syncwarp();
          }
        }
      }
      // all TMEM reads of this warp are complete (tcgen05.wait::ld above): release the buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc_buf);
      ++it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

static inline int sm_count() {
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

template <int BN, int PLANES, int MMAS, int KSTEPS, bool HALO>
static int launch_tc(const ConvDev& d, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                     const CUtensorMap& t, size_t smem, cudaStream_t stream) {
  // function attributes are per device: set them once for every device this process launches on
  static std::mutex mu;
  static uint64_t done_mask = 0;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (!((done_mask >> (dev & 63)) & 1ull)) {
      cudaError_t attr_err = cudaFuncSetAttribute(conv_tc_kernel<BN, PLANES, MMAS, KSTEPS, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
      if (attr_err == cudaSuccess && BN <= 64)
        attr_err = cudaFuncSetAttribute(conv_tc_kernel<BN, PLANES, MMAS, KSTEPS, HALO>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
      done_mask |= 1ull << (dev & 63);
    }
  }
  // persistent grid: one CTA per SM, split between the N tiles (every CTA keeps one N tile)
  int ctas_x = sm_count() * d.ctas_per_sm / d.n_tiles;
  if (ctas_x < 1) ctas_x = 1;
  if (ctas_x > d.m_tiles) ctas_x = d.m_tiles;
  // balance: no CTA should walk more M tiles than ceil(m_tiles / ctas_x); shrink the grid to the
  // smallest one with the same number of rounds (fewer CTAs -> fewer resident-weight loads)
  const int rounds = (d.m_tiles + ctas_x - 1) / ctas_x;
  ctas_x = (d.m_tiles + rounds - 1) / rounds;
  dim3 grid(ctas_x, d.n_tiles);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  V2X_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, PLANES, MMAS, KSTEPS, HALO>, a0, a1, b, t, d));
  return V2X_OK;
}


// One (storage format, MMA passes) variant of the kernel family; each is instantiated in its own translation unit
// (conv_tc_p1.cu, conv_tc_m1.cu, conv_tc_m2.cu, conv_tc_m3.cu) so the build compiles them in parallel.
template <int PLANES, int MMAS>
static int conv_dispatch(const ConvDev& d, int bn, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                         const CUtensorMap& t, size_t smem, cudaStream_t stream) {
#define V2X_LAUNCH(BN_, KS_, HALO_)                                \
  if (bn == BN_ && d.kc == 16 * KS_ && (d.halo != 0) == HALO_)     \
    return launch_tc<BN_, PLANES, MMAS, KS_, HALO_>(d, a0, a1, b, t, smem, stream);
  // kc = 16 only occurs for the 13(16)-channel input layer, kc = 32 for the 32/96-channel layers;
  // halo mode needs resident weights (the small-N layers) or the streamed-weight ring (kc = 64, N >= 64)
  V2X_LAUNCH(32, 1, true) V2X_LAUNCH(32, 2, true) V2X_LAUNCH(32, 4, true)
  V2X_LAUNCH(64, 1, true) V2X_LAUNCH(64, 2, true) V2X_LAUNCH(64, 4, true)
  V2X_LAUNCH(128, 4, true) V2X_LAUNCH(192, 4, true) V2X_LAUNCH(256, 4, true)
  V2X_LAUNCH(32, 1, false) V2X_LAUNCH(32, 2, false) V2X_LAUNCH(32, 4, false)
  V2X_LAUNCH(48, 4, false)
  V2X_LAUNCH(64, 1, false) V2X_LAUNCH(64, 2, false) V2X_LAUNCH(64, 4, false)
  V2X_LAUNCH(128, 2, false) V2X_LAUNCH(128, 4, false)
  V2X_LAUNCH(192, 4, false)
  V2X_LAUNCH(256, 2, false) V2X_LAUNCH(256, 4, false)
#undef V2X_LAUNCH
  set_error("no kernel instantiation for block_n %d with kc %d", bn, d.kc);
  return V2X_ERR_UNSUPPORTED;
}

#define V2X_CONV_DISPATCH_ARGS const ConvDev &d, int bn, const CUtensorMap &a0, const CUtensorMap &a1, const CUtensorMap &b, \
                               const CUtensorMap &t, size_t smem, cudaStream_t stream
int conv_dispatch_p1(V2X_CONV_DISPATCH_ARGS);   // bf16 storage, 1 MMA per k-step
int conv_dispatch_m1(V2X_CONV_DISPATCH_ARGS);   // fp16 hi/lo storage, hi*hi
int conv_dispatch_m2(V2X_CONV_DISPATCH_ARGS);   // fp16 hi/lo storage, + a_lo*w_hi
int conv_dispatch_m3(V2X_CONV_DISPATCH_ARGS);   // fp16 hi/lo storage, + a_hi*w_lo

}  // namespace v2x
