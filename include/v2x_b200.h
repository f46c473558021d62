/*
 * v2x_b200.h -- C ABI of the sm_100a collaborative-perception hot path.
 *
 * The reference (ai4ce/V2X-Sim -> coperception) has no FFI: its hot path is a chain of ATen
 * library calls made from python nn.Modules.  Each entry point below replaces one such call
 * site (file:line under /root/reference/coperception/coperception/, "CP/"); the python modules
 * in v2x-sim_b200/coperception/ keep the reference's nn.Module surface and call these through
 * ctypes.  INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - `stream` is a cudaStream_t passed as void* (the caller's current stream); no entry point
 *     synchronises, allocates device memory or touches the default stream;
 *   - the caller owns every buffer; nothing is retained after the call returns;
 *   - return 0 on success, <0 on error; v2x_last_error() returns a thread-local message.
 *
 * Activation layout ("act"): NHWC, 16-bit elements, `planes` planes, plane p at element offset
 * p * N*H*W*C.  The plane count fixes the storage format:
 *   planes == 1 (V2X_FMT_BF16) : one bf16 plane -- the throughput mode BASELINE.json's config names;
 *   planes == 2 (V2X_FMT_F16X2): x = hi + lo with hi = fp16(x) (saturating), lo = fp16(x - hi): ~22 mantissa
 *                                bits.  A convolution over such tensors issues `mmas` tensor-core passes per
 *                                k-step (fp32 accumulate): 3 = hi*hi + hi*w_lo + a_lo*w_hi (weights packed
 *                                V2X_FMT_F16X2, "fp16x3"), 2 = hi*hi + a_lo*w_hi (weights one fp16 plane,
 *                                V2X_FMT_F16), 1 = hi*hi only (the lo plane of the input is not read).  The
 *                                per-layer choice that meets the 1e-3 parity contract at the least tensor
 *                                time is the "mixed" precision of v2x_b200/nets.py.
 */
#ifndef V2X_B200_H_
#define V2X_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* storage formats of act tensors (== their plane count) and of packed weights */
#define V2X_FMT_BF16 1   /* one bf16 plane                                      */
#define V2X_FMT_F16X2 2  /* fp16 hi plane + fp16 lo plane                       */
#define V2X_FMT_F16 3    /* one fp16 plane (packed weights of mmas = 1 / 2)     */

#define V2X_OK 0
#define V2X_ERR_ARG -1
#define V2X_ERR_CUDA -2
#define V2X_ERR_UNSUPPORTED -3

/* library / build info; v2x_version() == 100 * major + minor */
int v2x_version(void);
const char* v2x_last_error(void);
/* 1 if the current device is compute capability 10.x (tcgen05/TMA kernels can run) */
int v2x_device_ok(void);

/* ---- epilogue modes of v2x_conv_fwd -------------------------------------------------- */
#define V2X_EPI_ACT 0      /* y = [relu](acc + bias) -> act, optional 2x nearest upsample on store */
#define V2X_EPI_F32_SPLIT 1 /* y = acc + bias -> fp32 NHWC, channels [0,split) to out0, [split,cout) to out1     */
#define V2X_EPI_F32_NCHW 3  /* y = acc + bias -> fp32 NCHW [N][cout][H][W] in out0 (segmentation logits)                       */
#define V2X_EPI_GRU 2      /* zero-hidden ConvGRU gate epilogue, see v2x_conv_params                           */
#define V2X_EPI_TAIL_F32_SPLIT 4 /* t = relu(acc + bias) (act format, kept in shared memory, never stored);
                                    y = tail_weights * t + tail_bias -> fp32 NHWC split like V2X_EPI_F32_SPLIT:
                                    the detection heads' conv3x3+BN+ReLU -> conv1x1 pair in ONE launch
                                    (DetModelBase.py:283-296, 319-329); needs cout == block_n == 64              */

/*
 * One fused convolution launch: implicit GEMM on tcgen05 tensor cores, A tiles (8x16 output
 * pixels x kc channels, one filter tap at a time) and weight tiles staged by TMA, fp32
 * accumulators in TMEM, fused bias(+folded BN)+ReLU / head-split / GRU-gate epilogue.
 *
 * Replaces: F.relu(bn(conv(x)))                       CP/models/det/backbone/Backbone.py:102-136
 *           conv(cat(interpolate(a), b)) (two srcs)   Backbone.py:173-237
 *           Conv3D 1x1x1 + BN3d + ReLU (taps == 1)    Backbone.py:280-300
 *           heads conv3x3+BN+ReLU, conv1x1            CP/models/det/base/DetModelBase.py:283-296,319-329
 *           GRUCell(conv(x,W_ih), conv(0,W_hh))       CP/utils/convolutional_rnn/functional.py:84-105
 */
typedef struct v2x_conv_params {
  /* inputs: up to two sources, concatenated along channels (src0 channels first) */
  const void* src[2];  /* act, [planes][N][Hin][Win][cin[s]]; src[1] may be NULL          */
  int32_t cin[2];      /* channels per source (multiple of 16)                                 */
  int32_t n_maps;      /* N                                                                    */
  int32_t h_out, w_out;/* output spatial size (multiple of 8 / 16); input is stride * output   */
  int32_t stride;      /* 1 or 2                                                               */
  int32_t taps;        /* 9 (3x3, pad 1) or 1 (1x1)                                            */
  int32_t planes;      /* storage format of the inputs and the output: 1 (bf16) or 2 (fp16 hi/lo)   */
  /* packed weights from v2x_pack_conv_weights, bias fp32: [cout_pad][k_total] bf16 (planes == 1),
     [2][cout_pad][k_total] fp16 hi/lo (planes == 2, mmas == 3) or [cout_pad][k_total] fp16
     (planes == 2, mmas == 1 or 2)                                                             */
  const void* weights;
  const float* bias;   /* [cout_pad]                                                           */
  int32_t cout;        /* logical output channels                                              */
  int32_t cout_pad;    /* rows of the packed weights (multiple of block_n)                     */
  int32_t block_n;     /* N tile: 32, 48, 64, 128, 192 (GRU) or 256                            */
  int32_t epilogue;    /* V2X_EPI_*                                                            */
  int32_t relu;        /* EPI_ACT: apply ReLU                                                  */
  int32_t upsample2x;  /* EPI_ACT: store every output pixel to the 2x2 block of a [2H,2W] map  */
  void* out0;          /* EPI_ACT/GRU: act [planes][N][H(*2)][W(*2)][out_c_total]; F32_SPLIT: fp32 */
  void* out1;          /* F32_SPLIT: second fp32 output                                        */
  int32_t out_c_total; /* channel stride of out0 (EPI_ACT/GRU)                                 */
  int32_t out_c_off;   /* channel offset inside out0                                           */
  int32_t split;       /* F32_SPLIT: first `split` channels go to out0 (row stride = split),
                          the rest to out1 (row stride = cout - split)                         */
  /* EPI_GRU: packed rows are [r(64) | z(64) | n(64)] per 64-channel block (block_n == 192),
     bias = [b_ih_r+b_hh_r | b_ih_z+b_hh_z | b_ih_n]; h' = (1 - sigmoid(z)) * tanh(n + sigmoid(r) * bhn[c]).
     Units (maps) whose agent slot is >= num_agent[b] pass `passthrough` through unchanged.   */
  const float* gru_bhn;      /* [cout/3] = b_hh_n                                              */
  const void* passthrough;   /* act, same geometry as the output; may be NULL                  */
  const int64_t* num_agent;  /* [batch][agents] (reference num_agent_tensor) or NULL           */
  int32_t batch, agents;     /* agent-major maps: global unit = batch * agent + b              */
  int32_t map_offset;        /* global unit index of this launch's map 0 (sharded plans), else 0 */
  int32_t gru_pre_act;       /* EPI_GRU: src[1] is an act tensor [planes][N][H][W][cout] of gate pre-activations (packed
                                gate order, e.g. conv(mean, W_ih[:, C:]) + bias from an EPI_ACT launch with relu = 0); cin[1] must
                                be 192 and the packed weights carry 192 extra K columns holding the identity
                                (W[n][taps*cin[0] + j] = (n % 192 == j)), so every N tile accumulates its own window on the
                                tensor core instead of loading it in the epilogue */
  int32_t tap_pack;          /* EPI_ACT, 3x3 stride-1, cout == 32: the three horizontal filter taps are packed into the N
                                dimension (weights [planes][96][3 * sum(cin)], row = kw*32 + co, k = (source, kh, ci);
                                cout_pad == block_n == 96) and summed in the epilogue -- 3 MMAs of N = 96 per k-step instead
                                of 9 of N = 32, whose cost is dominated by re-reading the A operand (csrc/conv_pack3.cu) */
  int32_t mmas;              /* planes == 2 only: tensor-core passes per k-step, 3 (0 means 3), 2 or 1 -- see the file header */
  /* EPI_GRU, optional: fp32 [N*H*W][cout] (packed gate order) added to the gate pre-activations -- the round-invariant
     half conv(mean, W_ih[:, C:]) + bias, computed once per frame by an EPI_F32_SPLIT launch (split == cout) so the three
     GNN rounds only convolve the changing half (V2VNet.py:99: cat([h_i, mean]); the mean never changes). */
  const float* gru_add;
  /* V2X_EPI_TAIL_F32_SPLIT: the fused 1x1 conv. tail_weights = packed [planes][tail_cout_pad][cout] (same format as the acts, K-major, as
     v2x_pack_conv_weights writes a taps == 1 operand), tail_bias fp32 [tail_cout_pad]; channels [0,split) of the tail
     output go to out0 (row stride split), [split,tail_cout) to out1 (row stride tail_cout - split). */
  const void* tail_weights;
  const float* tail_bias;
  int32_t tail_cout, tail_cout_pad;
} v2x_conv_params;

int v2x_conv_fwd(const v2x_conv_params* p, void* stream);
/* profiling aid: ablate one role of v2x_conv_fwd (0 normal, 1 no MMA, 2 no TMA loads, 3 no global stores); outputs are garbage when != 0 */
int v2x_set_debug_mode(int mode);
/* same contract on CUDA cores (no TMA / tcgen05); a test aid to bisect operand-packing vs tensor-core-path bugs */
int v2x_conv_fwd_crosscheck(const v2x_conv_params* p, void* stream);

/*
 * Weight packing (one-time, on device).  Writes a block of the packed operand:
 *   dst[(plane)][row_off + perm(co)][k_off + tap' * cin_pad + (ci - ci_lo)] = fmt(w[co][ci][tap] * s[co])
 *   bias[row_off + perm(co)] = (b[co] - mean[co]) * s[co] + beta[co],  s = gamma / sqrt(var + eps)  (s = 1 without BN)
 * w is OIHW fp32 ([cout][cin_total][taps]); ci in [ci_lo, ci_hi) selects one concat source;
 * cin_pad >= ci_hi - ci_lo (extra k columns must already be zero: memset dst first);
 * vflip != 0 mirrors the filter rows (tap' = (2 - kh) * 3 + kw), used to run the ConvGRU in the
 * un-flipped domain; gru_gates == 3 applies the gate interleave
 * perm(g * C + cb * 64 + c) = cb * 192 + g * 64 + c.
 * Replaces the BN arithmetic of nn.BatchNorm2d (eval) at Backbone.py:102-136 by folding.
 */
int v2x_pack_conv_weights(const float* w, const float* b, const float* bn_gamma, const float* bn_beta,
                          const float* bn_mean, const float* bn_var, float eps, int32_t cout, int32_t cin_total,
                          int32_t taps, int32_t ci_lo, int32_t ci_hi, int32_t cin_pad, int32_t vflip,
                          int32_t gru_gates, void* dst, float* dst_bias, int32_t fmt /* V2X_FMT_* */, int32_t cout_pad,
                          int32_t k_total, int32_t row_off, int32_t k_off, int32_t write_bias, void* stream);

/* bias / b_hh_n vectors of the zero-hidden GRU epilogue (functional.py:95-105 with hidden == 0):
   bias[perm(g*C+c)] = b_ih[g*C+c] + (g < 2 ? b_hh[g*C+c] : 0);  bhn[c] = b_hh[2*C+c]              */
int v2x_pack_gru_bias(const float* b_ih, const float* b_hh, int32_t c, float* bias, float* bhn, void* stream);

/*
 * fp32 NHWC occupancy/feature input -> act with channels zero-padded to c_pad.
 * Replaces x.to(torch.float) + the NCHW view at Backbone.py:100-101 (memory is already NHWC,
 * V2VNet.py:51).
 */
int v2x_pack_input(const float* x, void* out, int64_t n_pixels, int32_t c, int32_t c_pad, int32_t planes,
                   void* stream);

/*
 * Cross-agent warp-and-mean (V2VNet neighbour aggregation), all (scene, target, source) pairs in
 * one launch, in the UN-flipped domain:
 *   out[b,i] = mean_{j != i, j < na[b]} bilinear_sample(x[b,j], theta'(T[b,j,i]))      (include_self == 0)
 *   theta' = [[T00, -T01, -T03/32], [-T10, T11, +T13/32]]  (flip folded in)
 * grid_sample semantics: bilinear, zeros padding, align_corners=False.
 * x, out: act [planes][A*B][H][W][C] agent-major; trans: [B][A][A][4][4] float64 (device);
 * num_agent: [B][A] int64 (device).  Targets i >= na[b] are written as zeros.
 * unit_offset / unit_count select a slice of target units (agent-major index B*i + b) for unit-sharded multi-GPU
 * plans: x is the complete (all-gathered) tensor, out holds unit_count maps; unit_count <= 0 means all.
 * Replaces DetModelBase.feature_transformation / build_neighbors_feature_list + torch.mean(torch.stack)
 * at CP/models/det/base/DetModelBase.py:139-209 and CP/models/det/V2VNet.py:85-98.
 */
int v2x_warp_mean_fwd(const void* x, void* out, const double* trans, const int64_t* num_agent, int32_t batch,
                      int32_t agents, int32_t h, int32_t w, int32_t c, int32_t planes, int32_t include_self,
                      int32_t only_v2i, int32_t unit_offset, int32_t unit_count, void* stream);

/* ---- when2com / who2com (CP/models/det/When2com.py) --------------------------------------- */

/*
 * Small fp32 linear layer y = [relu](x W^T + b), rows x in_f -> rows x out_f (KmGenerator MLP, When2com.py:415-430).
 * in_mode 0: x is fp32 [rows][in_f].  in_mode 1: x is an act [planes][rows][hw][c] read in NCHW-flatten order
 * (i = ch * hw + px), the order `features_map.view(-1, n_feat)` produces (When2com.py:429).
 * split > 1: the maps are viewed as `maps * split` rows of in_f = hw * c / split values (the seg model's
 * `.view(-1, 4096)` over 16384-wide features, When2Com_UNet.py:207-226 -- SURVEY Q9); `rows` of them are computed.
 */
int v2x_linear_fwd(const void* x, const float* w, const float* b, float* y, int32_t rows, int32_t in_f, int32_t out_f,
                   int32_t relu, int32_t in_mode, int32_t hw, int32_t c, int32_t planes, int32_t split, int32_t maps,
                   void* stream);

/*
 * Attention scores of MIMOGeneralDotProductAttention.forward (When2com.py:374-412) and the eval-time gate:
 *   q' = W q + bw (Linear query_size -> key_size); attn[b][k][j] = softmax_k(key[b,k] . q'[b,j])
 *   coef[b][k][j]: gate_mode 0 = attn; 1 "activated" = p * (p > 0.2) with p = attn + 0.001 * I (When2com.py:125-148,273-276);
 *                  2 "argmax_test" = one-hot over k of argmax p (When2com.py:94-123)
 * keys [A*B][key_size], querys [A*B][query_size] fp32, agent-major rows (B*i + b); outputs fp32 [B][A][A].
 */
int v2x_attn_scores_fwd(const float* keys, const float* querys, const float* w, const float* bw, float* attn,
                        float* coef, int32_t batch, int32_t agents, int32_t key_size, int32_t query_size,
                        int32_t gate_mode, void* stream);

/*
 * Gated cross-agent fuse (un-flipped domain, same theta' as v2x_warp_mean_fwd):
 *   warp_flag 1: out[b,q] = sum_{k<na[b]} coef[b,k,q] * (k == q ? x[b,q] : warp(x[b,q], T[b,q,k]))  -- the reference's
 *                val_mat[b,k,q] pairing (When2com.py:206-225,397-412); agents q >= na[b] give zeros
 *   warp_flag 0: out[b,q] = sum_{k<A} coef[b,k,q] * x[b,k]
 * Replaces the [B,A,A,C,H,W] val_mat + broadcast multiply + sum (never materialised here).
 * Unit-sharded plans: targets [unit_offset, unit_offset + unit_count) of the agent-major units are computed (out is
 * indexed locally) from an x tensor holding units [x_unit_offset, x_unit_offset + x_units); 0 counts = all units.
 */
int v2x_warp_gated_fwd(const void* x, void* out, const double* trans, const int64_t* num_agent, const float* coef,
                       int32_t batch, int32_t agents, int32_t h, int32_t w, int32_t c, int32_t planes,
                       int32_t warp_flag, int32_t only_v2i, int32_t unit_offset, int32_t unit_count,
                       int32_t x_unit_offset, int32_t x_units, void* stream);

/* ---- intermediate-fusion baselines (CP/models/det/base/FusionBase.py, CP/models/seg/FusionBase.py) ---------- */

/*
 * out[b,i] = reduce over {x[b,i]} U {warp_{j->i}(x[b,j]) : j != i present} ; mode 0 mean, 1 sum, 2 max.
 * Same geometry / theta' / grid_sample semantics as v2x_warp_mean_fwd; agent slots i >= na[b] keep their own map.
 * Replaces torch.mean/sum/max(torch.stack(neighbor_feat_list)) at MeanFusion.py:11-12, SumFusion.py:20-21,
 * MaxFusion.py:20-21 (and CatFusion.py:23) plus the warps of DetModelBase.build_neighbors_feature_list (:171-209).
 */
int v2x_warp_reduce_fwd(const void* x, void* out, const double* trans, const int64_t* num_agent, int32_t batch,
                        int32_t agents, int32_t h, int32_t w, int32_t c, int32_t planes, int32_t mode,
                        int32_t only_v2i, void* stream);

/*
 * Pair scores of the learned fusion weights: s[b,i,k,p] = relu(conv1_4(relu(bn(conv1_3(relu(bn(conv1_2(
 * relu(bn(conv1_1(cat[x_i, warp_{k->i}(x_k)]))))))))))) at pixel p (DiscoNet.py:132-155, AgentWiseWeightedFusion.py:44-76).
 * q: act [planes][A*B][h][w][256] = conv1x1(x, [Wa ; Wb]) with conv1_1's BN folded (channels [0,128): tg half incl.
 * bias, [128,256): neighbour half, no bias) -- conv1_1 is linear and pointwise, so it commutes with the warp.
 * w2 [32][128], b2 [32], w3 [8][32], b3 [8], w4 [8], b4 [1]: fp32, BN(eval) already folded.
 * scores: fp32 [B][A][A][h*w]; only participating (i, k < na[b]) entries are written.
 */
int v2x_pair_score_fwd(const void* q, float* scores, const double* trans, const int64_t* num_agent, const float* w2,
                       const float* b2, const float* w3, const float* b3, const float* w4, const float* b4,
                       int32_t batch, int32_t agents, int32_t h, int32_t w, int32_t planes, int32_t only_v2i,
                       void* stream);

/*
 * AgentWiseWeightedFusion.py:17-26,73-74: coef[b,i,:] = softmax_k relu(b5 + sum_p w5f[p] * s[b,i,k,p]) over k < na[b]
 * (conv1_5 is a 32x32 "valid" conv over the whole map; w5f = its filter with rows mirrored, because scores live in
 * the un-flipped domain).  coef: fp32 [B][A][A], zeros elsewhere.
 */
int v2x_agent_softmax_fwd(const float* scores, const float* w5f, const float* b5, const int64_t* num_agent,
                          float* coef, int32_t batch, int32_t agents, int32_t hw, void* stream);

/*
 * out[b,i,p] = sum_k c * (k == i ? x[b,i,p] : warp_{k->i}(x[b,k])[p]) over participating k < na[b];
 * coef_mode 0: c = coef[b][i][k] (AgentWiseWeightedFusion.py:27-34);
 * coef_mode 1: coef = scores [B][A][A][h*w] and c = exp(s_k) / sum_k' exp(s_k') per pixel (DiscoNet.py:88-107).
 * Agent slots i >= na[b] keep their own map.
 */
int v2x_warp_weighted_fwd(const void* x, void* out, const double* trans, const int64_t* num_agent, const float* coef,
                          int32_t coef_mode, int32_t batch, int32_t agents, int32_t h, int32_t w, int32_t c,
                          int32_t planes, int32_t only_v2i, void* stream);

/* out[unit] = x[unit] for agent slots absent from their scene (i >= na[b]); map_elems = h*w*c per map.  Restores
 * FusionBase's "only present agents are rewritten" (FusionBase.py:41-63) after a fuse stage that touched every map. */
int v2x_restore_absent_fwd(const void* x, void* out, const int64_t* num_agent, int32_t batch, int32_t agents,
                           int64_t map_elems, int32_t planes, void* stream);

/* ---- input densification on device (CP/datasets/V2XSimDet.py:291-302) ------------------------------------- */

/*
 * Sparse voxel index list -> dense occupancy BEV in the conv path's act layout.  Replaces, per agent and frame,
 *   curr_voxels = np.zeros(dims, bool); curr_voxels[i0, i1, i2] = 1; np.rot90(curr_voxels, 3); .astype(np.float32)
 * (V2XSimDet.py:294-302) + the dense H2D copy + v2x_pack_input.  idx: int32 [capacity][4] rows (map, i0, i1, i2);
 * count: DEVICE int32 scalar = valid rows (<= capacity), read by the kernel so a captured graph replays with new
 * counts; out: act [planes][n_maps][h][w][c_pad] (zeroed here, then 1.0 scattered; rot90_k3 != 0 applies
 * np.rot90(., 3): voxel (i0, i1) -> pixel (i1, h - 1 - i0)); bad_count: DEVICE int32, incremented once per
 * out-of-range row (such rows are dropped; the caller checks it -- numpy raises IndexError there).
 */
int v2x_voxelize_fwd(const int32_t* idx, const int32_t* count, int32_t capacity, void* out, int32_t n_maps, int32_t h,
                     int32_t w, int32_t c, int32_t c_pad, int32_t planes, int32_t rot90_k3, int32_t* bad_count,
                     void* stream);

/* bool / uint8 NHWC occupancy [n_pixels][c] (non-zero = 1.0) -> act, channels zero-padded to c_pad: the
 * `padded_voxel_points` array before its .astype(np.float32) (V2XSimDet.py:299-302), 13 instead of 52 bytes/pixel. */
int v2x_pack_input_u8(const uint8_t* x, void* out, int64_t n_pixels, int32_t c, int32_t c_pad, int32_t planes,
                      void* stream);

/* ---- detection post-processing on device (CP/utils/detection_util.py:256-373, CP/utils/postprocess.py:72-113) -- */

/*
 * Per agent map: score = softmax(cls)[..., 1]; candidates = score > score_thr (0.7 in the reference, postprocess.py:84),
 * visited in descending score order (ties: higher anchor index first); boxes decoded from (loc, anchors) as
 * bev_box_decode_torch + center_to_corner_box2d do (fp32); a candidate is dropped when its polygon IoU (fp64, rounded
 * to fp32 like compute_iou's array) with an already kept box exceeds iou_thr (0.01 at detection_util.py:357-359).
 *   cls [n_maps][n_anchors][2], loc [n_maps][n_anchors][6] fp32 (the model's outputs, device);
 *   anchors fp32 [n_maps][n_anchors][6] (x, y, w, h, sin, cos), or one [n_anchors][6] table when anchors_shared != 0;
 *   cap: candidates kept per map (1024 / 2048 / 4096), highest scores first -- cand_count[m] > cap means overflow and
 *        must be treated as an error by the caller (the reference has no cap);
 *   workspaces: keys_ws u64 [n_maps][cap], boxes_ws f32 [n_maps][cap][8];
 *   outputs (pick order): sel_idx int32 [n_maps][cap] (anchor index = the reference's selected_idx), sel_score,
 *        sel_corners [..][8] (x0,y0,..,x3,y3), sel_count [n_maps], cand_count [n_maps].
 */
int v2x_det_nms_fwd(const float* cls, const float* loc, const float* anchors, int32_t n_maps, int32_t n_anchors,
                    int32_t anchors_shared, float score_thr, float iou_thr, int32_t cap, uint64_t* keys_ws,
                    float* boxes_ws, int32_t* cand_count, int32_t* sel_idx, float* sel_score, float* sel_corners,
                    int32_t* sel_count, void* stream);

/* ---- segmentation UNet pieces (CP/models/seg/SegModelBase.py) ------------------------------ */
/* fp32 NCHW [n][c][h][w] (what SegModule.py:49 hands the model) -> act NHWC, channels zero-padded to c_pad */
int v2x_pack_input_nchw(const float* x, void* out, int32_t n, int32_t c, int32_t h, int32_t w, int32_t c_pad,
                        int32_t planes, void* stream);
/* nn.MaxPool2d(2) (SegModelBase.py:113): act [n][2*h_out][2*w_out][c] -> [n][h_out][w_out][c] */
int v2x_maxpool2_fwd(const void* x, void* out, int32_t n, int32_t h_out, int32_t w_out, int32_t c, int32_t planes,
                     void* stream);
/* nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True) (SegModelBase.py:125): act [n][h][w][c] -> [n][2h][2w][c] */
int v2x_upsample_bilinear2_fwd(const void* x, void* out, int32_t n, int32_t h_in, int32_t w_in, int32_t c,
                               int32_t planes, void* stream);

/* act (NHWC) -> fp32 NCHW, for returning intermediate maps to torch callers */
int v2x_act_to_nchw_f32(const void* act, float* out, int32_t n, int32_t h, int32_t w, int32_t c, int32_t planes,
                        void* stream);

/* =====================================================================================================================
 * Training step: train-mode BatchNorm, its backward, the byte movers of the backward graph and the weight
 * gradient.  Data gradients run through v2x_conv_fwd with transposed / 180-degree-rotated packed weights.
 * Replaces nn.BatchNorm2d in .train() mode (CP/models/det/backbone/Backbone.py:102-136, entered through
 * CP/utils/CoDetModule.py:217-291 after model.train()) and torch.autograd's conv2d / batch_norm / relu / interpolate
 * backward nodes (loss.backward(), CoDetModule.py:289-291).
 * ===================================================================================================================== */
/* per-channel sum / sum of squares of an act [planes][n_pixels][c] over all pixels (fp64); zeroes the outputs first */
int v2x_bn_stats_fwd(const void* z, int64_t n_pixels, int32_t c, int32_t planes, double* sum, double* sumsq, void* stream);
/* batch statistics -> scale = gamma * invstd, shift = beta - mean * scale, mean, invstd (fp32 [c]); updates the running
 * buffers in place like nn.BatchNorm2d (momentum, UNBIASED batch variance); gamma / beta / running_* may be NULL */
int v2x_bn_finalize(const double* sum, const double* sumsq, int64_t count, const float* gamma, const float* beta, float eps,
                    float momentum, float* running_mean, float* running_var, float* scale, float* shift, float* mean,
                    float* invstd, int32_t c, void* stream);
/* y = [relu](z * scale + shift), act -> act */
int v2x_bn_relu_apply_fwd(const void* z, void* y, int64_t n_pixels, int32_t c, int32_t planes, const float* scale,
                          const float* shift, int32_t relu, void* stream);
/* backward of y = relu(BN_train(z)): s1[c] = sum dyh (= d beta), s2[c] = sum dyh * xhat (= d gamma) with dyh = dy * [y > 0],
 * dz = scale * (dyh - s1 / n - xhat * s2 / n); s1 / s2 are fp64 [c] outputs */
int v2x_bn_relu_bwd(const void* dy, const void* z, void* dz, int64_t n_pixels, int32_t c, int32_t planes, const float* scale,
                    const float* shift, const float* mean, const float* invstd, int32_t relu, double* s1, double* s2,
                    void* stream);
/* mode 0: zero-stuffing out[2i][2j] = in[i][j] (stride-2 data gradient as a stride-1 correlation); mode 1: nearest 2x
 * upsample (F.interpolate, Backbone.py:176,195,214,233); mode 2: its backward (2x2 block sums).  (h_out, w_out) = size of `out` */
int v2x_resample2(const void* in, void* out, int32_t n, int32_t h_out, int32_t w_out, int32_t c, int32_t planes, int32_t mode,
                  void* stream);
/* dst += src over act tensors of n_elems elements per plane */
int v2x_act_add(void* dst, const void* src, int64_t n_elems, int32_t planes, void* stream);
/* dw[co][ci_off + ci][tap] += scale * sum_pixels dz[.][co] * x[shifted][ci]  (fp32 OIHW gradient of a 3x3 pad-1 stride-1/2 or
 * 1x1 conv; dz act [planes][n][h_out][w_out][co], x act [planes][n][h_out*stride][w_out*stride][ci]; co_log / ci_log = logical
 * channels (<= the padded co / ci); x is one concat source covering filter input channels [ci_off, ci_off + ci_log) of ci_total) */
int v2x_conv_wgrad(const void* dz, const void* x, int32_t n, int32_t h_out, int32_t w_out, int32_t co, int32_t ci, int32_t planes,
                   int32_t stride, int32_t taps, float* dw, int32_t co_log, int32_t ci_log, int32_t ci_off, int32_t ci_total,
                   float scale, void* stream);
/* out[i] = (accumulate ? out[i] : 0) + scale * in[i]: fp64 per-channel sums -> fp32 parameter gradients */
int v2x_scale_to_f32(const double* in, float* out, int32_t count, float scale, int32_t accumulate, void* stream);

/* zero-hidden ConvGRU gates in train mode (functional.py:84-105 with h = 0): a = conv(cat[h, mean], W_ih) + b_ih as an act
 * [n_pixels][3c] in [r | z | n] channel order, bhh fp32 [3c]; units (pixel / hw) whose agent slot is absent copy `pass` */
int v2x_gru_gates_fwd(const void* a, const float* bhh, const void* pass, void* h, int64_t n_pixels, int32_t hw, int32_t c,
                      int32_t planes, const int64_t* num_agent, int32_t batch, int32_t agents, void* stream);
/* their backward: da (act [n_pixels][3c]), dpass = dh on absent units (else 0), dbhn[c] = sum da_n * r (fp64) */
int v2x_gru_gates_bwd(const void* dh, const void* a, const float* bhh, void* da, void* dpass, int64_t n_pixels, int32_t hw,
                      int32_t c, int32_t planes, const int64_t* num_agent, int32_t batch, int32_t agents, double* dbhn,
                      void* stream);
/* backward of v2x_warp_mean_fwd (grid_sample backward + mean): dx fp32 [A*B][h][w][c] (zeroed here) += scattered dmean */
int v2x_warp_mean_bwd(const void* dmean, float* dx, const double* trans, const int64_t* num_agent, int32_t batch, int32_t agents,
                      int32_t h, int32_t w, int32_t c, int32_t planes, int32_t include_self, int32_t only_v2i, void* stream);

/* backward of v2x_warp_reduce_fwd (Mean / Sum / Max fusion) and of v2x_warp_weighted_fwd with per-pair coefficients
 * (AgentWiseWeightedFusion): dx fp32 [A*B][h][w][c] (zeroed here) += dout scattered through the members' bilinear taps
 * (mode 0 scaled by 1/count; mode 2 routed, per channel, to the first member attaining the maximum, recomputed from the
 * forward input x, which only mode 2 reads; mode 3 scaled by the constant coef[b][i][k], fp32 [B][A][A], else unused);
 * absent agent slots pass their gradient through */
int v2x_warp_reduce_bwd(const void* dout, const void* x, float* dx, const float* coef, const double* trans,
                        const int64_t* num_agent, int32_t batch, int32_t agents, int32_t h, int32_t w, int32_t c,
                        int32_t planes, int32_t mode, int32_t only_v2i, void* stream);

/* backward of v2x_warp_weighted_fwd with coef_mode 1 (DiscoNet's per-pixel softmax of pair scores): dscores fp32
 * [B][A][A][h*w] (zeroed here) = w_k (g_k - sum_m w_m g_m) with g_k = <dout, member_k> over channels, and dx fp32
 * [A*B][h][w][c] (zeroed here) += w_k * dout scattered through member k's taps; x is the forward input */
int v2x_warp_weighted_bwd(const void* dout, const void* x, float* dx, float* dscores, const float* scores, const double* trans,
                          const int64_t* num_agent, int32_t batch, int32_t agents, int32_t h, int32_t w, int32_t c,
                          int32_t planes, int32_t only_v2i, void* stream);

/* backward of v2x_warp_gated_fwd (when2com fuse): dcoef fp32 [B][A][A] (zeroed here) = <dout[b,q], val[b,k,q]> and dx fp32
 * [A*B][h][w][c] (zeroed here) += coef[b,k,q] * dout[b,q] scattered through val[b,k,q]'s taps; x is the forward input */
int v2x_warp_gated_bwd(const void* dout, const void* x, float* dx, float* dcoef, const float* coef, const double* trans,
                       const int64_t* num_agent, int32_t batch, int32_t agents, int32_t h, int32_t w, int32_t c, int32_t planes,
                       int32_t warp_flag, int32_t only_v2i, void* stream);

/* v2x_conv_wgrad on the tensor cores: a GEMM with M = co, N = ci per filter tap, K = pixels whose operands are the NHWC act
 * tensors themselves, consumed as MN-major UMMA operands (csrc/wgrad_tc.cu); same arguments and result */
int v2x_conv_wgrad_tc(const void* dz, const void* x, int32_t n, int32_t h_out, int32_t w_out, int32_t co, int32_t ci, int32_t planes,
                      int32_t stride, int32_t taps, float* dw, int32_t co_log, int32_t ci_log, int32_t ci_off, int32_t ci_total,
                      float scale, void* stream);

/* nn.MaxPool2d(2) backward: x act [n][2h][2w][c] (the forward input), dy act [n][h][w][c] -> dx act like x (gradient to the
 * first maximum of each window, SegModelBase.py:113) */
int v2x_maxpool2_bwd(const void* x, const void* dy, void* dx, int32_t n, int32_t h_out, int32_t w_out, int32_t c, int32_t planes,
                     void* stream);
/* nn.Upsample(x2, bilinear, align_corners=True) backward (SegModelBase.py:125): dy act [n][2h][2w][c] -> dx fp32 [n][h][w][c] */
int v2x_upsample_bilinear2_bwd(const void* dy, float* dx, int32_t n, int32_t h_in, int32_t w_in, int32_t c, int32_t planes,
                               void* stream);

/* ---- multi-GPU exchange over NVLink peer memory (one process per GPU) -------------------------
 *
 * The unit-sharded plans exchange the layer-3 maps x_3 once per forward ("ncclAllGather of x_3 ... issued right after
 * conv3_2", replacing the neighbour reads of CP/models/det/V2VNet.py:85-94 when a scene's agents live on different
 * GPUs).  These entry points do that exchange from the device: every rank owns a peer-visible REGION
 *   [ flag block (V2X_PEER_FLAG_BYTES): ready[8], consumed[8] uint64 step numbers | payload ]
 * and pushes its own maps into its slot of every rank's region with plain stores that travel over NVLink; readiness is
 * published through the flag blocks (st.release.sys / ld.acquire.sys), so there is no host-issued collective and the
 * whole forward is one CUDA graph.  Per step, on one stream:  v2x_peer_begin -> v2x_peer_push -> v2x_peer_wait ->
 * (the fuse kernel reads the region) -> v2x_peer_done.  Every rank must run the same number of steps.
 * v2x_peer_alloc / v2x_peer_open / v2x_peer_close / v2x_peer_free are set-up calls: they DO allocate device memory and
 * synchronise (the exception to the conventions above); handles travel between the processes by any host channel
 * (v2x_b200/sharding.py uses a torch.distributed all-gather of the 64 bytes).  host_* arguments are HOST pointers.
 */
#define V2X_PEER_HANDLE_BYTES 64
#define V2X_PEER_FLAG_BYTES 4096
/* cudaMalloc a zeroed region of `bytes`, return its device pointer and its CUDA IPC handle (64 bytes) */
int v2x_peer_alloc(int64_t bytes, void** host_ptr_out, void* host_handle_out);
/* map another rank's region into this process (cudaIpcOpenMemHandle, peer access enabled lazily) / unmap it */
int v2x_peer_open(const void* host_handle, void** host_ptr_out);
int v2x_peer_close(void* ptr);
int v2x_peer_free(void* ptr);
/* step += 1 (device uint64 `step`), then wait until every other rank has consumed step - 1 (flags of the LOCAL region).
 * A wait that exceeds timeout_ms stores 100 + rank-waited-for into *err (device int32) and gives up instead of hanging. */
int v2x_peer_begin(const void* flags_local, void* step, int32_t rank, int32_t world, int32_t timeout_ms, int32_t* err,
                   void* stream);
/* copy src (act planes [p][elems_per_plane], plane stride src_plane_stride elements) into host_dst[r] (this rank's slot in
 * rank r's payload, plane stride dst_plane_stride) for every r < world, host_dst_planes[r] planes each (0 = skip); when all
 * stores are done, publish ready[rank] = step in every region (host_flag_regions[r] = base of rank r's region).
 * `counter`: device uint32, zero before the first call. */
int v2x_peer_push(const void* src, int64_t elems_per_plane, int64_t src_plane_stride, void* const* host_dst,
                  const int32_t* host_dst_planes, int64_t dst_plane_stride, void* const* host_flag_regions,
                  const void* step, void* counter, int32_t rank, int32_t world, void* stream);
/* wait until ready[r] >= step for every other rank r (error code 200 + r on timeout) */
int v2x_peer_wait(const void* flags_local, const void* step, int32_t rank, int32_t world, int32_t timeout_ms, int32_t* err,
                  void* stream);
/* publish consumed[rank] = step in every region: the payload of this step may be overwritten */
int v2x_peer_done(void* const* host_flag_regions, const void* step, int32_t rank, int32_t world, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* V2X_B200_H_ */
