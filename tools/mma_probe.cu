// Microbenchmark (profiling aid, not product): cycles per tcgen05.mma (kind::f16, M=128, K=16, SS mode)
// as a function of N, of the number of independent TMEM accumulators the stream alternates between,
// and of the shared-memory swizzle mode.  One CTA per SM, no TMA (operands are whatever is in smem).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu
#include <cstdio>
#include <cstdlib>

#include "../v2x-sim_b200/csrc/common.cuh"

namespace v2x {
void set_error(const char*, ...) {}
int cuda_fail(cudaError_t, const char*) { return -1; }
}  // namespace v2x
using namespace v2x;

struct Result {
  long long cycles;
};

// reps MMAs; accumulator index cycles through n_acc buffers; a_step/b_step advance the operand
// descriptors (bytes) between consecutive MMAs (wrapping every 4) to mimic a k loop.
template <int N>
__global__ void __launch_bounds__(128, 1) probe(int reps, int n_acc, uint32_t layout_type, uint32_t sbo, int a_step,
                                                int b_step, Result* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  constexpr uint32_t idesc = make_idesc_bf16_m128(N);
  constexpr uint32_t ACC = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;
  if (warp == 1) {
    const uint64_t da0 = make_smem_desc(base, sbo, layout_type);
    const uint64_t db0 = make_smem_desc(base + 16384, sbo, layout_type);
    long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {  // pass 0 warms up
      __syncwarp();
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < reps; ++i) {
          const uint32_t d = tmem_base + (uint32_t)(i & (n_acc - 1)) * ACC;
          const uint64_t da = da0 + (uint64_t)(((i & 3) * a_step) >> 4);
          const uint64_t db = db0 + (uint64_t)(((i & 3) * b_step) >> 4);
          umma_bf16(d, da, db, idesc, i >= n_acc ? 1u : 0u);
        }
        umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), pass & 1);
      t1 = clock64();
    }
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out->cycles = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int N>
static void run(const char* tag, int n_acc, uint32_t layout, uint32_t sbo, int a_step, int b_step, Result* dres) {
  const int reps = 512;
  if (n_acc * (N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256) > 512) return;
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe<N><<<148, 128, 64 * 1024>>>(reps, n_acc, layout, sbo, a_step, b_step, dres);
  cudaError_t e = cudaDeviceSynchronize();
  Result r{};
  cudaMemcpy(&r, dres, sizeof(r), cudaMemcpyDeviceToHost);
  printf("%-10s N=%3d acc=%d a_step=%3d b_step=%3d : %7.1f cycles/MMA (math floor %3d)%s\n", tag, N, n_acc, a_step,
         b_step, (double)r.cycles / reps, N / 2, e == cudaSuccess ? "" : "  CUDA ERROR");
}

int main() {
  Result* dres;
  cudaMalloc(&dres, sizeof(Result));
  for (int n_acc : {1, 2, 4}) {
    run<32>("sw128", n_acc, 2, 1024, 32, 32, dres);
    run<64>("sw128", n_acc, 2, 1024, 32, 32, dres);
    run<128>("sw128", n_acc, 2, 1024, 32, 32, dres);
    run<192>("sw128", n_acc, 2, 1024, 32, 32, dres);
    run<256>("sw128", n_acc, 2, 1024, 32, 32, dres);
  }
  run<32>("sw64", 1, 4, 512, 32, 32, dres);
  run<32>("sw32", 1, 6, 256, 0, 0, dres);
  run<64>("sw64", 1, 4, 512, 32, 32, dres);
  // halo-mode addressing of the conv kernel: 8-row groups one halo row (10 px) apart, tap shifts of whole pixel rows
  for (int n_acc : {1, 2}) {
    run<32>("halo-sw64", n_acc, 4, 640, 64, 32, dres);
    run<32>("halo-sw64k", n_acc, 4, 640, 32, 32, dres);
    run<64>("halo-sw64", n_acc, 4, 640, 64, 32, dres);
    run<32>("halo-sw32", n_acc, 6, 320, 32, 0, dres);
    run<64>("halo-sw128", n_acc, 2, 1280, 128, 32, dres);
    run<128>("halo-sw128", n_acc, 2, 1280, 128, 32, dres);
    run<192>("halo-sw128", n_acc, 2, 1280, 128, 32, dres);
  }
  run<32>("sw128-fix", 1, 2, 1024, 0, 0, dres);
  run<192>("sw128-fix", 1, 2, 1024, 0, 0, dres);
  run<256>("sw128-fix", 1, 2, 1024, 0, 0, dres);
  return 0;
}
