// Probe (profiling/bring-up aid, not product): can tcgen05.mma read a 3x3 tap's A operand straight out of
// a TMA-loaded halo tile by shifting the smem descriptor's start address?
//   halo tile  = [18 rows][10 px][KC ch] bf16, written by ONE cp.async.bulk.tensor.3d with swizzle = KC*2 bytes
//   output tile= 16 rows x 8 px = 128 GEMM rows; row m -> pixel (m / 8, m % 8)
//   tap (kh,kw): A row m = halo[(m/8 + kh)][(m%8 + kw)][:]  => start = base + (kh*10 + kw) * KC*2,
//                8-row groups are one halo row apart => SBO = 10 * KC*2 (not a multiple of the swizzle atom)
// B = identity [N=KC][K=KC], so D[m][n] must equal halo[(m/8+kh)][(m%8+kw)][n].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/halo_probe tools/halo_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../v2x-sim_b200/csrc/common.cuh"

namespace v2x {
void set_error(const char*, ...) {}
int cuda_fail(cudaError_t, const char*) { return -1; }
}  // namespace v2x
using namespace v2x;

constexpr int HH = 18, HW = 10;

template <int KC>
__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmA,
                                                const __grid_constant__ CUtensorMap tmB, int kh, int kw,
                                                int use_base_offset, float* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + 32768;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t ROW = KC * 2;
  constexpr uint32_t LAYOUT = KC == 64 ? 2u : KC == 32 ? 4u : 6u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar[0]), 1);
    mbar_init(smem_u32(&bar[1]), 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_slot), 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 1) {
    if (elect_one()) {
      mbar_expect_tx(smem_u32(&bar[0]), HH * HW * ROW + KC * ROW);
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
          ::"r"(a_base), "l"(&tmA), "r"(smem_u32(&bar[0])), "r"(0), "r"(0), "r"(0) : "memory");
      tma_load_2d(b_base, &tmB, smem_u32(&bar[0]), 0, 0);
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar[0]), 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t start = a_base + (kh * HW + kw) * ROW;
      uint64_t da = make_smem_desc(start, HW * ROW, LAYOUT);
      if (use_base_offset) da |= (uint64_t)((start >> 7) & 7) << 49;
      const uint64_t db = make_smem_desc(b_base, 8 * ROW, LAYOUT);
      constexpr uint32_t idesc = make_idesc_bf16_m128(KC);
      for (int kk = 0; kk < KC / 16; ++kk) umma_bf16(tmem_base, da + 2 * kk, db + 2 * kk, idesc, kk ? 1u : 0u);
      umma_commit(smem_u32(&bar[1]));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar[1]), 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < KC; c += 16) {
    float v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int i = 0; i < 16; ++i) out[row * KC + c + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int KC>
static void run(EncodeTiledFn enc, int pat) {
  const CUtensorMapSwizzle sw = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                                 : CU_TENSOR_MAP_SWIZZLE_32B;
  std::vector<__nv_bfloat16> hA(HH * HW * KC), hB(KC * KC);
  for (int h = 0; h < HH; ++h)
    for (int w = 0; w < HW; ++w)
      for (int c = 0; c < KC; ++c) {
        const int pix = h * HW + w;  // integers <= 255 are exact in bf16; two patterns disambiguate row / channel aliasing
        const float v = pat == 0 ? (float)((pix % 16) * 16 + (c % 16)) : (float)(((pix / 16) % 16) * 16 + ((c / 16) % 4) * 4 + (pix % 4));
        hA[pix * KC + c] = __float2bfloat16(v);
      }
  for (int n = 0; n < KC; ++n)
    for (int k = 0; k < KC; ++k) hB[n * KC + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *dA, *dB;
  float* dOut;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dOut, 128 * KC * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  cuuint32_t es[3] = {1, 1, 1};
  {
    cuuint64_t dims[3] = {(cuuint64_t)KC, HW, HH};
    cuuint64_t str[2] = {(cuuint64_t)KC * 2, (cuuint64_t)HW * KC * 2};
    cuuint32_t box[3] = {(cuuint32_t)KC, HW, HH};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("encode A failed %d\n", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)KC, (cuuint64_t)KC};
    cuuint64_t str[1] = {(cuuint64_t)KC * 2};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)KC};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("encode B failed %d\n", (int)r);
  }
  cudaFuncSetAttribute(probe<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> hOut(128 * KC);
  for (int ubo = 0; ubo < 2; ++ubo) {
    int ok_taps = 0;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        cudaMemset(dOut, 0, 128 * KC * 4);
        probe<KC><<<1, 128, 64 * 1024>>>(tmA, tmB, kh, kw, ubo, dOut);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("KC=%d base_offset=%d tap(%d,%d): CUDA error %s\n", KC, ubo, kh, kw, cudaGetErrorString(e));
          return;
        }
        cudaMemcpy(hOut.data(), dOut, 128 * KC * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first_bad = -1;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < KC; ++n) {
            const float want = __bfloat162float(hA[(((m / 8) + kh) * HW + (m % 8) + kw) * KC + n]);
            if (hOut[m * KC + n] != want) {
              if (first_bad < 0) first_bad = m * KC + n;
              ++bad;
            }
          }
        if (!bad) ++ok_taps;
        else
          printf("  KC=%d base_offset=%d tap(%d,%d): %d mismatches, first at m=%d n=%d got %.4f\n", KC, ubo, kh, kw, bad,
                 first_bad / KC, first_bad % KC, hOut[first_bad]);
      }
    printf("KC=%d (swizzle %dB) pattern %d base_offset=%s : %d / 9 taps exact\n", KC, KC * 2, pat, ubo ? "(start>>7)&7" : "0", ok_taps);
  }
}

int main() {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  EncodeTiledFn enc = (EncodeTiledFn)ptr;
  for (int pat = 0; pat < 2; ++pat) {
    run<64>(enc, pat);
    run<32>(enc, pat);
    run<16>(enc, pat);
  }
  return 0;
}
