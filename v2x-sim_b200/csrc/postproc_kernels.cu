// Detection post-processing on device: softmax score -> score filter -> box decode -> corners ->
// greedy rotated-box NMS, per agent map.  Replaces, per agent and frame, the D2H copy of 393 216 x (2 + 6) floats plus
// the python / shapely loop of CP/utils/detection_util.py:256-373 (apply_nms_det) and CP/utils/postprocess.py:72-113
// (non_max_suppression).  Index / integer work is bit-exact against the CPU restatement (oracle/postproc.py); the box
// arithmetic is fp32 like the reference's, the polygon IoU fp64 like shapely's.
#include "common.cuh"

namespace v2x {

// ---------------------------------------------------------------------------------------------
// 1. candidates: score = softmax(cls)[1] (detection_util.py:275); keep score > thr (postprocess.py:84).
//    key = (score bits << 32) | anchor index: positive floats order like their bit patterns, so sorting keys
//    descending orders by score, ties by descending index (= a stable ascending argsort read backwards, :85).
// ---------------------------------------------------------------------------------------------
__global__ void nms_collect_kernel(const float2* __restrict__ cls, long long n_maps, int n_anchors, float score_thr, int cap,
                                   unsigned long long* __restrict__ keys, int* __restrict__ cand_count) {
  const long long total = n_maps * n_anchors;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
       gid += (long long)gridDim.x * blockDim.x) {
    const float2 l = __ldg(cls + gid);
    const float m = fmaxf(l.x, l.y);
    const float e0 = expf(l.x - m), e1 = expf(l.y - m);
    const float s = e1 / (e0 + e1);
    if (s > score_thr) {
      const int map = (int)(gid / n_anchors);
      const int a = (int)(gid - (long long)map * n_anchors);
      const int slot = atomicAdd(cand_count + map, 1);
      if (slot < cap)
        keys[(long long)map * cap + slot] = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned)a;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// convex quad IoU in fp64: Sutherland-Hodgman clip of A by the four half-planes of B (both made counter-clockwise),
// shoelace areas; what shapely's intersection().area / union().area computes for two convex quads (postprocess.py:50).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double signed_area2(const double (&p)[8][2], int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) {
    const int j = i + 1 == n ? 0 : i + 1;
    s += p[i][0] * p[j][1] - p[j][0] * p[i][1];
  }
  return s;
}

__device__ double quad_iou_f64(const float* __restrict__ qa, const float* __restrict__ qb) {
  double a[8][2], b[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a[i][0] = (double)qa[2 * i];
    a[i][1] = (double)qa[2 * i + 1];
    b[i][0] = (double)qb[2 * i];
    b[i][1] = (double)qb[2 * i + 1];
  }
  double sb = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    sb += b[i][0] * b[j][1] - b[j][0] * b[i][1];
  }
  const double area_a = 0.5 * fabs(signed_area2(a, 4));
  const double area_b = 0.5 * fabs(sb);
  if (sb < 0.0) {  // make B counter-clockwise
    double t0 = b[1][0], t1 = b[1][1];
    b[1][0] = b[3][0]; b[1][1] = b[3][1];
    b[3][0] = t0; b[3][1] = t1;
  }
  int n = 4;
  double tmp[8][2];
  for (int e = 0; e < 4 && n > 0; ++e) {
    const double ex0 = b[e][0], ey0 = b[e][1];
    const double dx = b[(e + 1) & 3][0] - ex0, dy = b[(e + 1) & 3][1] - ey0;
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const int j = i + 1 == n ? 0 : i + 1;
      const double ci = dx * (a[i][1] - ey0) - dy * (a[i][0] - ex0);   // >= 0: inside (left of the edge)
      const double cj = dx * (a[j][1] - ey0) - dy * (a[j][0] - ex0);
      if (ci >= 0.0 && m < 8) {
        tmp[m][0] = a[i][0]; tmp[m][1] = a[i][1]; ++m;
      }
      if ((ci >= 0.0) != (cj >= 0.0) && m < 8) {   // a convex quad clipped by 4 half-planes has at most 8 vertices
        const double t = ci / (ci - cj);
        tmp[m][0] = a[i][0] + t * (a[j][0] - a[i][0]);
        tmp[m][1] = a[i][1] + t * (a[j][1] - a[i][1]);
        ++m;
      }
    }
    n = m;
    for (int i = 0; i < n; ++i) {
      a[i][0] = tmp[i][0]; a[i][1] = tmp[i][1];
    }
  }
  if (n < 3) return 0.0;
  const double inter = 0.5 * fabs(signed_area2(a, n));
  const double uni = area_a + area_b - inter;
  return uni > 0.0 ? inter / uni : 0.0;
}

// detection_util.py:376-398 (bev_box_decode_torch) + obj_util.py:270-359 (center_to_corner_box2d, rotation_2d), fp32
__device__ __forceinline__ void decode_corners(const float* __restrict__ enc, const float* __restrict__ an, float* c8) {
  const float xa = an[0], ya = an[1], wa = an[2], ha = an[3], sina = an[4], cosa = an[5];
  const float xp = enc[0], yp = enc[1], wp = enc[2], hp = enc[3], sinp = enc[4], cosp = enc[5];
  const float h = __fdiv_rn(ha, expf(hp));
  const float w = __fdiv_rn(wa, expf(wp));
  const float x = __fsub_rn(xa, __fmul_rn(w, xp));
  const float y = __fsub_rn(ya, __fmul_rn(h, yp));
  const float s = __fadd_rn(__fmul_rn(sina, cosp), __fmul_rn(cosa, sinp));
  const float co = __fsub_rn(__fmul_rn(cosa, cosp), __fmul_rn(sina, sinp));
  // corners of the un-rotated box in the reference's order x0y1, x1y1, x1y0, x0y0 (obj_util.py:300-316), then
  // (cx, cy) -> (cx*cos + cy*sin, -cx*sin + cy*cos) (rotation_2d: corners @ [[cos,-sin],[sin,cos]]), then + centre
  const float hx = __fmul_rn(w, 0.5f), hy = __fmul_rn(h, 0.5f);
  const float cx[4] = {-hx, hx, hx, -hx};
  const float cy[4] = {hy, hy, -hy, -hy};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    c8[2 * i] = __fadd_rn(__fadd_rn(__fmul_rn(cx[i], co), __fmul_rn(cy[i], s)), x);
    c8[2 * i + 1] = __fadd_rn(__fadd_rn(__fmul_rn(cx[i], -s), __fmul_rn(cy[i], co)), y);
  }
}

// ---------------------------------------------------------------------------------------------
// 2. one block per map: sort the candidate keys (bitonic, shared memory), decode their boxes, greedy NMS.
//    The NMS walks the sorted list; for every surviving box all threads test the boxes behind it in parallel
//    (AABB reject in fp32, exact polygon IoU in fp64), so the sequential depth is the number of KEPT boxes.
// ---------------------------------------------------------------------------------------------
template <int CAP>
__global__ void __launch_bounds__(1024) nms_map_kernel(const float* __restrict__ loc, const float* __restrict__ anchors,
                                                       int n_anchors, int anchors_per_map_shared, float iou_thr,
                                                       unsigned long long* __restrict__ keys_ws,
                                                       const int* __restrict__ cand_count, float* __restrict__ boxes_ws,
                                                       int* __restrict__ sel_idx, float* __restrict__ sel_score,
                                                       float* __restrict__ sel_corners, int* __restrict__ sel_count) {
  __shared__ unsigned long long keys[CAP];
  __shared__ unsigned char alive[CAP];
  __shared__ int s_cur, s_nsel;
  const int map = blockIdx.x;
  const int n = min(cand_count[map], CAP);
  unsigned long long* gk = keys_ws + (long long)map * CAP;
  for (int t = threadIdx.x; t < CAP; t += blockDim.x) keys[t] = t < n ? gk[t] : 0ull;
  __syncthreads();
  // bitonic sort, descending
  for (int k = 2; k <= CAP; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < CAP; t += blockDim.x) {
        const int p = t ^ j;
        if (p > t) {
          const unsigned long long a = keys[t], b = keys[p];
          const bool desc = (t & k) == 0;
          if (desc ? a < b : a > b) {
            keys[t] = b;
            keys[p] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  float* boxes = boxes_ws + (long long)map * CAP * 8;
  const float* loc_m = loc + (long long)map * n_anchors * 6;
  const float* an_m = anchors + (anchors_per_map_shared ? 0ll : (long long)map * n_anchors * 6);
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const unsigned a = (unsigned)(keys[t] & 0xffffffffull);
    float c8[8];
    decode_corners(loc_m + (long long)a * 6, an_m + (long long)a * 6, c8);
#pragma unroll
    for (int i = 0; i < 8; ++i) boxes[t * 8 + i] = c8[i];
    alive[t] = 1;
    gk[t] = keys[t];   // sorted keys back to the workspace (diagnostics / tests)
  }
  if (threadIdx.x == 0) {
    s_cur = 0;
    s_nsel = 0;
  }
  __syncthreads();
  int* o_idx = sel_idx + (long long)map * CAP;
  float* o_score = sel_score + (long long)map * CAP;
  float* o_corn = sel_corners + (long long)map * CAP * 8;
  while (true) {
    if (threadIdx.x == 0) {
      int i = s_cur;
      while (i < n && !alive[i]) ++i;
      s_cur = i;
      if (i < n) {
        const int s = s_nsel++;
        o_idx[s] = (int)(keys[i] & 0xffffffffull);
        o_score[s] = __uint_as_float((unsigned)(keys[i] >> 32));
        for (int e = 0; e < 8; ++e) o_corn[s * 8 + e] = boxes[i * 8 + e];
      }
    }
    __syncthreads();
    const int i = s_cur;
    if (i >= n) break;
    float bi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) bi[e] = boxes[i * 8 + e];
    const float ix0 = fminf(fminf(bi[0], bi[2]), fminf(bi[4], bi[6])), ix1 = fmaxf(fmaxf(bi[0], bi[2]), fmaxf(bi[4], bi[6]));
    const float iy0 = fminf(fminf(bi[1], bi[3]), fminf(bi[5], bi[7])), iy1 = fmaxf(fmaxf(bi[1], bi[3]), fmaxf(bi[5], bi[7]));
    for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
      if (!alive[j]) continue;
      const float* bj = boxes + j * 8;
      const float jx0 = fminf(fminf(bj[0], bj[2]), fminf(bj[4], bj[6])), jx1 = fmaxf(fmaxf(bj[0], bj[2]), fmaxf(bj[4], bj[6]));
      const float jy0 = fminf(fminf(bj[1], bj[3]), fminf(bj[5], bj[7])), jy1 = fmaxf(fmaxf(bj[1], bj[3]), fmaxf(bj[5], bj[7]));
      if (jx0 > ix1 || jx1 < ix0 || jy0 > iy1 || jy1 < iy0) continue;   // disjoint boxes: IoU == 0 <= threshold
      if ((float)quad_iou_f64(bi, bj) > iou_thr) alive[j] = 0;            // float32 IoU vs threshold (postprocess.py:52,104)
    }
    __syncthreads();
    if (threadIdx.x == 0) s_cur = i + 1;
    // the next iteration's leading __syncthreads orders this write before any read
  }
  if (threadIdx.x == 0) sel_count[map] = s_nsel;
}

}  // namespace v2x

using namespace v2x;

extern "C" int v2x_det_nms_fwd(const float* cls, const float* loc, const float* anchors, int32_t n_maps,
                               int32_t n_anchors, int32_t anchors_shared, float score_thr, float iou_thr, int32_t cap,
                               uint64_t* keys_ws, float* boxes_ws, int32_t* cand_count, int32_t* sel_idx,
                               float* sel_score, float* sel_corners, int32_t* sel_count, void* stream) {
  V2X_REQUIRE(cls && loc && anchors && keys_ws && boxes_ws && cand_count && sel_idx && sel_score && sel_corners && sel_count,
              "null pointer");
  V2X_REQUIRE(n_maps > 0 && n_anchors > 0, "empty geometry");
  V2X_REQUIRE(cap == 1024 || cap == 2048 || cap == 4096, "cap (candidates kept per map) must be 1024, 2048 or 4096");
  V2X_REQUIRE((reinterpret_cast<uintptr_t>(cls) & 7) == 0, "cls must be 8-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  V2X_CUDA_TRY(cudaMemsetAsync(cand_count, 0, sizeof(int32_t) * n_maps, s));
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const long long total = (long long)n_maps * n_anchors;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  nms_collect_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const float2*>(cls), n_maps, n_anchors, score_thr,
                                                      cap, reinterpret_cast<unsigned long long*>(keys_ws), cand_count);
  V2X_CUDA_TRY(cudaGetLastError());
  unsigned long long* kw = reinterpret_cast<unsigned long long*>(keys_ws);
  if (cap == 1024)
    nms_map_kernel<1024><<<n_maps, 1024, 0, s>>>(loc, anchors, n_anchors, anchors_shared, iou_thr, kw, cand_count, boxes_ws,
                                                sel_idx, sel_score, sel_corners, sel_count);
  else if (cap == 2048)
    nms_map_kernel<2048><<<n_maps, 1024, 0, s>>>(loc, anchors, n_anchors, anchors_shared, iou_thr, kw, cand_count, boxes_ws,
                                                sel_idx, sel_score, sel_corners, sel_count);
  else
    nms_map_kernel<4096><<<n_maps, 1024, 0, s>>>(loc, anchors, n_anchors, anchors_shared, iou_thr, kw, cand_count, boxes_ws,
                                                sel_idx, sel_score, sel_corners, sel_count);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}
