// Fused convolution as an implicit GEMM on the sm_100a tensor cores.
//
//   M = output pixels (one CTA owns an 8x16 pixel patch of one map = 128 GEMM rows = 128 TMEM lanes)
//   N = output channels (block_n columns of fp32 accumulators in TMEM)
//   K = (source, filter tap, input channel); one pipeline stage = one tap x kc channels
//
// A operand: the activation tensor is NHWC bf16, so the im2col rows of one tap are a shifted
//   8x16xkc box of the input -- fetched by ONE TMA tiled load per stage (4-D map {C,W,H,N*planes};
//   negative / out-of-range coordinates are zero-filled by the TMA unit = the conv's zero padding).
//   Stride-2 convs use a 5-D view {2C, W/2, 2, H/2, N*planes} of the same memory so that the
//   parity of the tap selects a dense box (no element strides needed).
// B operand: packed weights [planes][cout_pad][K] (K-major), 2-D TMA loads.
// Both land in shared memory in the canonical K-major swizzled UMMA layout (swizzle width =
// kc * 2 bytes), are consumed by tcgen05.mma (issued by one thread), and the accumulator is read
// back with tcgen05.ld for the fused epilogue (bias+ReLU -> bf16 planes, head split -> fp32, or the
// zero-hidden ConvGRU gates).
//
// Warp roles (128 threads): warp0/lane0 = TMA producer, warp1/lane0 = MMA issuer, warp2 = TMEM
// allocator; all four warps run the epilogue (warp w owns TMEM lanes 32w..32w+31).
// Several CTAs are resident per SM (smem permitting) so one CTA's epilogue overlaps another's
// main loop.
//
// Reference call sites replaced: see include/v2x_b200.h (v2x_conv_fwd).
#include "conv_tc.cuh"

namespace v2x {

// ---------------------------------------------------------------------------------------------
// CUDA-core cross-check kernel: same operands, same epilogues, no TMA / tensor cores.
// One thread = one output pixel x 16 output channels (x3 gates for the GRU).  Test aid only.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ld_act(const ConvDev& p, int s, long long idx, long long plane_stride) {
  const __nv_bfloat16* a = reinterpret_cast<const __nv_bfloat16*>(p.src[s]);
  return act_load1(a + idx, plane_stride, p.planes == 2 && p.pa == 1 ? 3 : p.planes);
}
__device__ __forceinline__ float ld_w(const ConvDev& p, long long row, long long k) {
  const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(p.weights);
  return act_load1(w + row * p.k_total + k, (long long)p.cout_pad * p.k_total, p.planes == 2 && p.pw == 1 ? 3 : p.planes);
}

__global__ void conv_ref_kernel(const ConvDev p) {
  const bool gru = p.epilogue == V2X_EPI_GRU;
  const int chunks = gru ? p.cout / 3 / 16 : p.cout_pad / 16;
  const long long total = (long long)p.n_maps * p.h_out * p.w_out * chunks;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int chunk = (int)(gid % chunks);
  const long long pix = gid / chunks;
  const int ow = (int)(pix % p.w_out);
  const int oh = (int)((pix / p.w_out) % p.h_out);
  const int n_img = (int)(pix / ((long long)p.w_out * p.h_out));
  const int h_in = p.h_out * p.stride, w_in = p.w_out * p.stride;

  if (gru_unit_absent(p, n_img)) {
    copy_passthrough(p, n_img, oh, ow, chunk * 16, 16);
    return;
  }
  const int ngate = gru ? 3 : 1;
  // packed row of gate g, channel chunk: GRU rows are [r(64)|z(64)|n(64)] per 64-channel block
  int prow[3];
  for (int g = 0; g < ngate; ++g) prow[g] = gru ? (chunk / 4) * 192 + g * 64 + (chunk % 4) * 16 : chunk * 16;
  float acc[3][16];
  for (int g = 0; g < 3; ++g)
    for (int i = 0; i < 16; ++i) acc[g][i] = 0.f;
  long long kbase = 0;
  const int nsrc_conv = p.gru_pre_act ? 1 : p.nsrc;
  if (p.gru_pre_act) {  // src[1] = pre-activations [planes][N][H][W][cout] in packed gate order
    const long long pre_plane = (long long)p.n_maps * p.h_out * p.w_out * p.cout;
    const long long base = (((long long)n_img * p.h_out + oh) * p.w_out + ow) * p.cout;
    for (int g = 0; g < ngate; ++g)
      for (int i = 0; i < 16; ++i) acc[g][i] += ld_act(p, 1, base + prow[g] + i, pre_plane);
  }
  for (int s = 0; s < nsrc_conv; ++s) {
    const long long plane_stride = (long long)p.n_maps * h_in * w_in * p.cin[s];
    for (int tap = 0; tap < p.taps; ++tap) {
      const int kh = p.taps == 9 ? tap / 3 : 1, kw = p.taps == 9 ? tap % 3 : 1;
      const int ih = oh * p.stride + kh - 1, iw = ow * p.stride + kw - 1;
      if (ih >= 0 && ih < h_in && iw >= 0 && iw < w_in) {
        const long long abase = (((long long)n_img * h_in + ih) * w_in + iw) * p.cin[s];
        for (int c = 0; c < p.cin[s]; ++c) {
          const float a = ld_act(p, s, abase + c, plane_stride);
          if (a != 0.f) {
            const long long k = kbase + (long long)tap * p.cin[s] + c;
            for (int g = 0; g < ngate; ++g)
              for (int i = 0; i < 16; ++i) acc[g][i] += a * ld_w(p, prow[g] + i, k);
          }
        }
      }
    }
    kbase += (long long)p.taps * p.cin[s];
  }
  if (gru) {
    epi_gru16(p, n_img, oh, ow, chunk * 16, acc[0], acc[1], acc[2], p.bias + prow[0], p.gru_bhn + chunk * 16,
              p.gru_add ? p.gru_add + (((long long)n_img * p.h_out + oh) * p.w_out + ow) * p.cout + prow[0] : nullptr);
  } else {
    const int ch0 = chunk * 16;
    if (ch0 >= p.cout) return;
    if (p.epilogue == V2X_EPI_ACT) epi_act16(p, n_img, oh, ow, ch0, acc[0], p.bias + ch0);
    else if (p.epilogue == V2X_EPI_F32_NCHW) epi_f32_nchw16(p, n_img, oh, ow, ch0, acc[0], p.bias + ch0);
    else epi_f32_split16(p, n_img, oh, ow, ch0, acc[0], p.bias + ch0);
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                      const cuuint32_t* box, int kc) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return V2X_ERR_CUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                                 : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_b, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d, rank %d)", (int)r, rank);
    return V2X_ERR_CUDA;
  }
  return V2X_OK;
}

// shared with conv_pack3.cu
int encode_map_shared(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                      const cuuint32_t* box, int kc) {
  return encode_map(m, base, rank, dims, strides_b, box, kc);
}
int launch_pack3(const v2x_conv_params* p, int debug_mode, cudaStream_t stream);

static int g_debug_mode = 0;

static int fill_dev(const v2x_conv_params* p, ConvDev& d) {
  V2X_REQUIRE(p != nullptr, "null params");
  V2X_REQUIRE(p->src[0] && p->weights && p->bias && p->out0, "null src/weights/bias/out0");
  V2X_REQUIRE(p->stride == 1 || p->stride == 2, "stride must be 1 or 2 (got %d)", p->stride);
  V2X_REQUIRE(p->taps == 9 || (p->taps == 1 && p->stride == 1), "taps must be 9, or 1 with stride 1");
  V2X_REQUIRE(p->planes == 1 || p->planes == 2, "planes must be 1 or 2");
  V2X_REQUIRE(p->mmas >= 0 && p->mmas <= 3 && (p->planes == 2 || p->mmas <= 1), "mmas must be 1..3 (0 = default), and > 1 only with planes == 2");
  V2X_REQUIRE(!(p->gru_pre_act && p->planes == 2), "gru_pre_act is a planes == 1 (bf16) feature");
  V2X_REQUIRE(p->n_maps > 0 && p->h_out > 0 && p->w_out > 0, "empty geometry");
  V2X_REQUIRE(p->stride == 1 || (p->h_out * 2) % 2 == 0, "bad geometry");
  V2X_REQUIRE(p->cin[0] > 0 && p->cin[0] % 16 == 0, "cin[0] must be a positive multiple of 16");
  V2X_REQUIRE(p->src[1] == nullptr || (p->cin[1] > 0 && p->cin[1] % 16 == 0), "cin[1] must be a multiple of 16");
  const int bn = p->block_n;
  V2X_REQUIRE(bn == 32 || bn == 48 || bn == 64 || bn == 128 || bn == 192 || bn == 256, "unsupported block_n %d", bn);
  V2X_REQUIRE(p->cout > 0 && p->cout_pad >= p->cout && p->cout_pad % bn == 0, "cout_pad must be a multiple of block_n");
  V2X_REQUIRE(p->cout % 16 == 0 || p->epilogue == V2X_EPI_F32_SPLIT || p->epilogue == V2X_EPI_F32_NCHW,
              "cout must be a multiple of 16");
  V2X_REQUIRE(!p->gru_pre_act || p->epilogue == V2X_EPI_GRU, "gru_pre_act only with EPI_GRU");
  d = ConvDev{};
  d.n_maps = p->n_maps; d.h_out = p->h_out; d.w_out = p->w_out; d.stride = p->stride; d.taps = p->taps;
  d.planes = p->planes;
  {
    const int mmas = p->planes == 2 ? (p->mmas == 0 ? 3 : p->mmas) : 1;
    d.pa = mmas >= 2 ? 2 : 1;
    d.pw = mmas == 3 ? 2 : 1;
  }
  d.nsrc = p->src[1] ? 2 : 1;
  d.cin[0] = p->cin[0]; d.cin[1] = p->src[1] ? p->cin[1] : 0;
  int kc = 64;
  for (int s = 0; s < d.nsrc; ++s)
    while (d.cin[s] % kc) kc >>= 1;
  d.kc = kc;
  d.num_k = 0; d.k_total = 0;
  d.gru_pre_act = p->gru_pre_act;
  for (int s = 0; s < d.nsrc; ++s) {
    d.cblocks[s] = d.cin[s] / kc;
    d.num_k += p->taps * d.cblocks[s];
    d.k_total += (s == 1 && p->gru_pre_act) ? d.cin[s] : p->taps * d.cin[s];   // identity columns: one "tap"
  }
  // partial tiles are allowed: TMA zero-fills reads beyond the map, the epilogue masks stores beyond it
  d.tiles_w = (p->w_out + kTileW - 1) / kTileW;
  d.tiles_per_img = d.tiles_w * ((p->h_out + kTileH - 1) / kTileH);
  d.m_tiles = p->n_maps * d.tiles_per_img;
  d.n_tiles = p->cout_pad / p->block_n;
  d.cout = p->cout; d.cout_pad = p->cout_pad;
  d.epilogue = p->epilogue; d.relu = p->relu; d.upsample2x = p->upsample2x;
  d.out0 = p->out0; d.out1 = p->out1;
  d.out_c_total = p->out_c_total; d.out_c_off = p->out_c_off; d.split = p->split;
  d.bias = p->bias; d.gru_bhn = p->gru_bhn; d.gru_add = p->gru_add; d.passthrough = p->passthrough;
  d.num_agent = reinterpret_cast<const long long*>(p->num_agent);
  d.batch = p->batch; d.agents = p->agents; d.map_offset = p->map_offset;
  d.debug_mode = g_debug_mode;
  d.a_tile_bytes = 128u * kc * 2u;
  d.b_tile_bytes = ((uint32_t)bn * kc * 2u + 1023u) & ~1023u;
  d.stage_bytes = d.pa * d.a_tile_bytes + d.pw * d.b_tile_bytes;
  d.tx_bytes = d.pa * d.a_tile_bytes + d.pw * (uint32_t)bn * kc * 2u;
  d.sbo = 8u * kc * 2u;
  d.layout_type = kc == 64 ? 2u : kc == 32 ? 4u : 6u;
  const int up = p->upsample2x ? 2 : 1;
  d.src[0] = p->src[0]; d.src[1] = p->src[1]; d.weights = p->weights;
  switch (p->epilogue) {
    case V2X_EPI_ACT:
      V2X_REQUIRE(p->out_c_total >= p->out_c_off + p->cout && p->out_c_total % 8 == 0 && p->out_c_off % 8 == 0,
                  "bad output channel window");
      d.out_plane_stride = (long long)p->n_maps * p->h_out * up * p->w_out * up * p->out_c_total;
      break;
    case V2X_EPI_F32_SPLIT:
      V2X_REQUIRE(p->out1 != nullptr || p->split >= p->cout, "F32_SPLIT needs out1");
      V2X_REQUIRE(p->split > 0 && p->split <= p->cout, "bad split");
      V2X_REQUIRE(p->split % 4 == 0 && p->cout % 4 == 0, "F32_SPLIT needs split %% 4 == 0 and cout %% 4 == 0");
      V2X_REQUIRE(!p->upsample2x, "upsample2x only with EPI_ACT");
      break;
    case V2X_EPI_F32_NCHW:
      V2X_REQUIRE(!p->upsample2x, "upsample2x only with EPI_ACT");
      break;
    case V2X_EPI_TAIL_F32_SPLIT:
      V2X_REQUIRE(bn == 64 && p->cout == 64 && p->cout_pad == 64, "fused tail needs cout == block_n == 64");
      V2X_REQUIRE(p->tail_weights && p->tail_bias, "fused tail needs tail_weights / tail_bias");
      V2X_REQUIRE(p->tail_cout > 0 && p->tail_cout <= p->tail_cout_pad && p->tail_cout_pad <= 64 && p->tail_cout_pad % 16 == 0,
                  "tail_cout_pad must be a multiple of 16, <= 64");
      V2X_REQUIRE(p->out1 != nullptr || p->split >= p->tail_cout, "fused tail needs out1");
      V2X_REQUIRE(p->split > 0 && p->split <= p->tail_cout && p->split % 4 == 0 && p->tail_cout % 4 == 0, "bad tail split");
      V2X_REQUIRE(!p->upsample2x, "upsample2x only with EPI_ACT");
      d.tail_cout = p->tail_cout; d.tail_cout_pad = p->tail_cout_pad; d.tail_bias = p->tail_bias;
      break;
    case V2X_EPI_GRU:
      V2X_REQUIRE(bn == 192 && p->cout % 192 == 0 && p->cout_pad == p->cout, "GRU epilogue needs block_n 192");
      V2X_REQUIRE(p->gru_bhn != nullptr, "GRU epilogue needs gru_bhn");
      V2X_REQUIRE(!p->gru_pre_act || (p->src[1] != nullptr && p->cin[1] == 192 && p->taps == 9 && p->stride == 1 &&
                                      p->cin[0] % 64 == 0),
                  "gru_pre_act needs src[1] (pre-activations), cin[1] == 192, a 3x3 stride-1 conv and cin[0] %% 64 == 0");
      V2X_REQUIRE(!p->upsample2x, "upsample2x only with EPI_ACT");
      V2X_REQUIRE(p->out_c_total >= p->out_c_off + p->cout / 3 && p->out_c_total % 8 == 0 && p->out_c_off % 8 == 0,
                  "bad output channel window");
      V2X_REQUIRE(p->num_agent == nullptr ||
                      (p->batch > 0 && p->agents > 0 && p->map_offset >= 0 && p->map_offset + p->n_maps <= p->batch * p->agents),
                  "num_agent needs map_offset + n_maps <= batch * agents");
      d.out_plane_stride = (long long)p->n_maps * p->h_out * p->w_out * p->out_c_total;
      break;
    default:
      V2X_REQUIRE(false, "unknown epilogue %d", p->epilogue);
  }
  return V2X_OK;
}

}  // namespace v2x

using namespace v2x;

// Shared-memory plan of one launch for a dynamic-smem budget (bytes): resident vs streamed weights, halo mode,
// k-blocks per stage and ring depth.  Fails (V2X_ERR_UNSUPPORTED via V2X_REQUIRE) if not even two stages fit.
static int plan_smem(const v2x_conv_params* p, ConvDev& d, uint32_t budget, size_t* smem_out) {
  const int bn = p->block_n;
  // output staging of the four epilogue warps comes off the top of the budget
  const bool tail = p->epilogue == V2X_EPI_TAIL_F32_SPLIT;
  d.epi_warp_bytes = p->epilogue == V2X_EPI_F32_SPLIT ? kStageF32
                     : (p->epilogue == V2X_EPI_F32_NCHW || tail) ? 0u : (uint32_t)p->planes * kStageActPlane;
  // fused tail: its weights [planes][8 KB] and the A2 tile [planes][16 KB] (which doubles as the fp32 output staging)
  const uint32_t tail_bytes = tail ? (uint32_t)p->planes * (8192u + 16384u) : 0u;
  const uint32_t epi_bytes = 4u * d.epi_warp_bytes + tail_bytes;
  budget -= epi_bytes;
  // Shared-memory plan.  Small weight operands (all of [block_n x K], e.g. the C=32 layers at 256x256)
  // stay resident for the CTA's lifetime so only activations stream; otherwise weights ride in the
  // stage ring next to their A tile.  The ring takes whatever is left, up to kMaxStages deep.
  d.num_b_tiles = d.num_k;
  const uint32_t b_all = (uint32_t)d.num_b_tiles * d.pw * d.b_tile_bytes;
  const uint32_t a_stage = d.pa * d.a_tile_bytes;
  d.b_resident = (b_all <= 96u * 1024u && b_all + 4u * a_stage <= budget) ? 1 : 0;
  // Halo mode: 3x3 stride-1 convs with resident weights read all nine taps out of one 18x10-pixel box.
  const uint32_t a_halo = ((uint32_t)(kHaloH * kHaloW) * d.kc * 2u + 1023u) & ~1023u;
  d.halo = (d.b_resident && p->taps == 9 && p->stride == 1 && (bn == 32 || bn == 64) && p->h_out % 16 == 0 &&
            p->w_out % 8 == 0 && b_all + 3u * d.pa * a_halo <= budget && !getenv("V2X_NO_HALO"))
               ? 1 : 0;
  d.b_stages = 0; d.b_taps = 1; d.b_stage_bytes = 0; d.b_ring_off = 0;
  // Halo mode with STREAMED weights (the big layers): the halo tile of a channel block stays in an A ring while its nine
  // weight tiles flow through a second ring -- activations cross L2->smem once instead of nine times.
  bool halo_stream = false;
  if (!d.b_resident && p->taps == 9 && p->stride == 1 && d.kc == 64 && bn >= 64 && p->h_out % 16 == 0 &&
      p->w_out % 8 == 0 && !getenv("V2X_NO_HALO") && !getenv("V2X_NO_HALO_STREAM")) {
    const uint32_t a_slot = d.pa * a_halo;
    const uint32_t b_tap = d.pw * d.b_tile_bytes;
    for (int taps = 3; taps >= 1 && !halo_stream; taps -= 2) {
      const int a_slots = 2;
      const uint32_t left = budget > a_slots * a_slot ? budget - a_slots * a_slot : 0;
      const int bst = (int)(left / (taps * b_tap));
      if (bst >= (taps == 3 ? 2 : 3)) {
        halo_stream = true;
        d.b_taps = taps;
        d.b_stages = bst > kMaxStages ? kMaxStages : bst;
        d.b_stage_bytes = taps * b_tap;
        d.num_stages = a_slots;
        d.b_ring_off = a_slots * a_slot;
      }
    }
  }
  if (halo_stream) {
    d.halo = 1;
    d.tiles_w = p->w_out / 8;
    d.tiles_per_img = d.tiles_w * (p->h_out / 16);
    d.m_tiles = p->n_maps * d.tiles_per_img;
    d.num_k = d.cblocks[0] + (d.nsrc > 1 ? d.cblocks[1] : 0);
    d.b_region_bytes = 0;
    d.stage_bytes = d.pa * a_halo;
    d.tx_bytes = d.pa * (uint32_t)(kHaloH * kHaloW) * d.kc * 2u;
  } else if (d.halo) {
    d.tiles_w = p->w_out / 8;
    d.tiles_per_img = d.tiles_w * (p->h_out / 16);
    d.m_tiles = p->n_maps * d.tiles_per_img;
    d.num_k = d.cblocks[0] + (d.nsrc > 1 ? d.cblocks[1] : 0);   // k-blocks = channel blocks; 9 taps each
    d.b_region_bytes = b_all;
    d.stage_bytes = d.pa * a_halo;
    d.tx_bytes = d.pa * (uint32_t)(kHaloH * kHaloW) * d.kc * 2u;
  } else if (d.b_resident) {
    d.b_region_bytes = b_all;
    d.stage_bytes = a_stage;
    d.tx_bytes = a_stage;
  } else {
    d.b_region_bytes = 0;
  }
  // One k-block = one filter tap x kc channels (one A box [+ its weight tile]).  A pipeline stage carries
  // several k-blocks so that each mbarrier round trip (a few hundred cycles of single-thread latency on
  // both the producer and the MMA side) feeds >= ~24 KB of operands; tiny stages starve the tensor core.
  d.kb_bytes = d.stage_bytes;
  d.kb_tx_bytes = d.tx_bytes;
  size_t smem_total = 0;
  if (halo_stream) {
    d.kb_per_stage = 1;
    d.stages_per_tile = d.num_k;
    smem_total = (size_t)d.b_ring_off + (size_t)d.b_stages * d.b_stage_bytes + 1024;
  }
  const uint32_t ring = budget - d.b_region_bytes;
  if (!halo_stream) {
  int g = d.halo ? (int)((20u * 1024u + d.kb_bytes - 1) / d.kb_bytes) : (int)((28u * 1024u + d.kb_bytes - 1) / d.kb_bytes);
  if (g > d.num_k) g = d.num_k;
  while (g > 1 && ring / ((uint32_t)g * d.kb_bytes) < 4) --g;      // keep the ring >= 4 deep
  for (int t = g; t >= 1 && t * 2 > g; --t)                         // prefer an even split of the k loop
    if (d.num_k % t == 0) { g = t; break; }
  d.kb_per_stage = g;
  d.stages_per_tile = (d.num_k + g - 1) / g;
  d.stage_bytes = (uint32_t)g * d.kb_bytes;
  int stages = (int)(ring / d.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  V2X_REQUIRE(stages >= 2, "stage of %u bytes does not fit in shared memory", d.stage_bytes);
  d.num_stages = stages;
  smem_total = (size_t)d.b_region_bytes + (size_t)stages * d.stage_bytes + 1024;
  }
  d.epi_off = (uint32_t)(smem_total - 1024);
  d.tail_w_off = d.epi_off + 4u * d.epi_warp_bytes;
  d.tail_a_off = d.tail_w_off + (uint32_t)p->planes * 8192u;
  V2X_REQUIRE(!tail || (d.b_resident && d.epi_off % 1024u == 0), "fused tail needs resident weights");
  *smem_out = smem_total + epi_bytes;
  return V2X_OK;
}

extern "C" int v2x_conv_fwd(const v2x_conv_params* p, void* stream_) {
  V2X_REQUIRE(p != nullptr, "null params");
  if (p->tap_pack) return launch_pack3(p, g_debug_mode, reinterpret_cast<cudaStream_t>(stream_));
  ConvDev d;
  int rc = fill_dev(p, d);
  if (rc) return rc;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int bn = p->block_n;

  // Small-N layers (a tile is only a few hundred cycles of tensor work, so the single MMA-issuing thread and the four
  // epilogue warps of one CTA are the bottleneck -- measured: the MMA warp retires ~210 dependent instructions per tile)
  // run TWO co-resident CTAs per SM when the operands fit in half the shared memory.
  size_t smem = 0;
  d.ctas_per_sm = 1;
  bool planned = false;
  if (bn <= 64 && !getenv("V2X_ONE_CTA")) {
    ConvDev d2 = d;
    d2.ctas_per_sm = 2;
    size_t smem2 = 0;
    const uint32_t budget2 = 106u * 1024u;  // 2 x (106 KB dynamic + ~5 KB static + 1 KB reserved) <= 228 KB per SM
    if (plan_smem(p, d2, budget2, &smem2) == V2X_OK && d2.b_resident && d2.num_stages >= (d2.halo ? 3 : 4)) {
      d = d2;
      smem = smem2;
      planned = true;
    }
  }
  if (!planned) {
    rc = plan_smem(p, d, (uint32_t)kSmemLimit - 1024u, &smem);
    if (rc) return rc;
  }

  CUtensorMap tmA[2], tmB;
  const int h_in = p->h_out * p->stride, w_in = p->w_out * p->stride;
  for (int s = 0; s < d.nsrc; ++s) {
    const cuuint64_t C = (s == 1 && d.gru_pre_act) ? (cuuint64_t)p->cout : (cuuint64_t)d.cin[s];
    const cuuint64_t NP = (cuuint64_t)p->n_maps * p->planes;
    if (p->stride == 1) {
      cuuint64_t dims[4] = {C, (cuuint64_t)w_in, (cuuint64_t)h_in, NP};
      cuuint64_t str[3] = {C * 2, (cuuint64_t)w_in * C * 2, (cuuint64_t)h_in * w_in * C * 2};
      cuuint32_t box[4] = {(cuuint32_t)d.kc, (cuuint32_t)(d.halo ? kHaloW : kTileW), (cuuint32_t)(d.halo ? kHaloH : kTileH), 1};
      rc = encode_map(&tmA[s], p->src[s], 4, dims, str, box, d.kc);
    } else {
      cuuint64_t dims[5] = {2 * C, (cuuint64_t)w_in / 2, 2, (cuuint64_t)h_in / 2, NP};
      cuuint64_t str[4] = {2 * C * 2, (cuuint64_t)w_in * C * 2, 2 * (cuuint64_t)w_in * C * 2,
                           (cuuint64_t)h_in * w_in * C * 2};
      cuuint32_t box[5] = {(cuuint32_t)d.kc, kTileW, 1, kTileH, 1};
      rc = encode_map(&tmA[s], p->src[s], 5, dims, str, box, d.kc);
    }
    if (rc) return rc;
  }
  if (d.nsrc == 1) tmA[1] = tmA[0];
  {
    cuuint64_t dims[2] = {(cuuint64_t)d.k_total, (cuuint64_t)d.pw * p->cout_pad};
    cuuint64_t str[1] = {(cuuint64_t)d.k_total * 2};
    cuuint32_t box[2] = {(cuuint32_t)d.kc, (cuuint32_t)bn};
    rc = encode_map(&tmB, p->weights, 2, dims, str, box, d.kc);
    if (rc) return rc;
  }
  CUtensorMap tmT = tmB;
  if (p->epilogue == V2X_EPI_TAIL_F32_SPLIT) {  // tail weights [planes][tail_cout_pad][64] bf16, one box per plane
    cuuint64_t dims[2] = {64, (cuuint64_t)p->planes * p->tail_cout_pad};
    cuuint64_t str[1] = {64 * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)p->tail_cout_pad};
    rc = encode_map(&tmT, p->tail_weights, 2, dims, str, box, 64);
    if (rc) return rc;
  }
  V2X_REQUIRE(d.num_k <= kMaxKBlocks, "too many k-blocks (%d > %d)", d.num_k, kMaxKBlocks);
  V2X_REQUIRE(!d.gru_pre_act || (d.halo && !d.b_resident), "gru_pre_act needs the halo + streamed-weight mode (16 | H, 8 | W)");
  if (p->planes == 1) return conv_dispatch_p1(d, bn, tmA[0], tmA[1], tmB, tmT, smem, stream);
  if (d.pw == 2) return conv_dispatch_m3(d, bn, tmA[0], tmA[1], tmB, tmT, smem, stream);
  if (d.pa == 2) return conv_dispatch_m2(d, bn, tmA[0], tmA[1], tmB, tmT, smem, stream);
  return conv_dispatch_m1(d, bn, tmA[0], tmA[1], tmB, tmT, smem, stream);
}

// Profiling aid: ablate one pipeline role of v2x_conv_fwd (results are then garbage).
// 0 = normal, 1 = no tcgen05.mma (TMA + epilogue only), 2 = no TMA loads (MMA on stale smem), 3 = no global stores,
// 4 = epilogue warps only do the accumulator-buffer handshake.
extern "C" int v2x_set_debug_mode(int mode) {
  V2X_REQUIRE(mode >= 0 && mode <= 4, "debug mode must be 0..4");
  g_debug_mode = mode;
  return V2X_OK;
}

// Same contract as v2x_conv_fwd, computed on CUDA cores without TMA / tcgen05.  A test aid used by
// tests/ to separate operand-packing bugs from tensor-core-path bugs; not used by the product path.
extern "C" int v2x_conv_fwd_crosscheck(const v2x_conv_params* p, void* stream_) {
  ConvDev d;
  int rc = fill_dev(p, d);
  if (rc) return rc;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const bool gru = p->epilogue == V2X_EPI_GRU;
  const long long chunks = gru ? p->cout / 3 / 16 : p->cout_pad / 16;
  const long long total = (long long)p->n_maps * p->h_out * p->w_out * chunks;
  const int threads = 128;
  conv_ref_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, stream>>>(d);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}
