// Shared device/host helpers for the sm_100a kernels: error plumbing, mbarrier / TMA / tcgen05
// PTX wrappers, bf16 hi/lo split.  Hand-written for sm_100a only (no other arch is supported).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/v2x_b200.h"

namespace v2x {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define V2X_CUDA_TRY(expr)                                        \
  do {                                                            \
    cudaError_t _e = (expr);                                      \
    if (_e != cudaSuccess) return ::v2x::cuda_fail(_e, #expr);    \
  } while (0)

#define V2X_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::v2x::set_error(__VA_ARGS__);    \
      return V2X_ERR_ARG;               \
    }                                   \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Storage formats of act / packed-weight tensors (include/v2x_b200.h):
//   planes == 1 (V2X_FMT_BF16)  : one bf16 plane
//   planes == 2 (V2X_FMT_F16X2) : x ~= hi + lo, two fp16 planes (about 22 mantissa bits together); hi saturates at the
//                                 largest finite fp16 instead of overflowing to inf
//   V2X_FMT_F16 (3, packed weights only): the hi plane alone
// Elements are addressed as 16-bit units whatever the format (the kernels keep __nv_bfloat16* as "a 16-bit element").
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// two fp32 -> packed fp16x2 (a in the low half), round-to-nearest-even, saturating to +-65504
__device__ __forceinline__ uint32_t f16x2_sat(float a, float b) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
__device__ __forceinline__ float2 f16x2_to_float2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__device__ __forceinline__ float2 bf16x2_to_float2(uint32_t v) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
}

// (a, b) -> packed pair(s) in the storage format of a `planes`-plane tensor; lo is only meaningful for planes == 2
template <int PLANES>
__device__ __forceinline__ void act_pack2(float a, float b, uint32_t& hi, uint32_t& lo) {
  if (PLANES == 2) {
    hi = f16x2_sat(a, b);
    const float2 h = f16x2_to_float2(hi);
    lo = f16x2_sat(a - h.x, b - h.y);
  } else {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = 0u;
  }
}
__device__ __forceinline__ void act_pack2(float a, float b, int planes, uint32_t& hi, uint32_t& lo) {
  if (planes == 2) act_pack2<2>(a, b, hi, lo);
  else act_pack2<1>(a, b, hi, lo);
}
template <int PLANES>
__device__ __forceinline__ float2 act_unpack2(uint32_t hi, uint32_t lo) {
  if (PLANES == 2) {
    float2 f = f16x2_to_float2(hi);
    const float2 g = f16x2_to_float2(lo);
    f.x += g.x; f.y += g.y;
    return f;
  }
  return bf16x2_to_float2(hi);
}
__device__ __forceinline__ float2 act_unpack2(uint32_t hi, uint32_t lo, int planes) {
  return planes == 2 ? act_unpack2<2>(hi, lo) : act_unpack2<1>(hi, lo);
}
// one element of a `planes`-plane tensor (lo plane `plane_stride` elements after the hi plane)
__device__ __forceinline__ float act_load1(const __nv_bfloat16* p, long long plane_stride, int fmt) {
  if (fmt == 2)
    return __half2float(*reinterpret_cast<const __half*>(p)) + __half2float(*reinterpret_cast<const __half*>(p + plane_stride));
  if (fmt == 3) return __half2float(*reinterpret_cast<const __half*>(p));   // the hi plane alone
  return __bfloat162float(*p);
}
// one value -> the 16-bit storage element(s)
__device__ __forceinline__ void act_split1(float x, int fmt, uint16_t& hi, uint16_t& lo) {
  if (fmt == 1) {
    __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi = __bfloat16_as_ushort(h);
    lo = 0;
  } else {
    const uint32_t h = f16x2_sat(x, 0.f);
    hi = (uint16_t)(h & 0xFFFFu);
    const float hf = f16x2_to_float2(h).x;
    lo = (uint16_t)(f16x2_sat(x - hf, 0.f) & 0xFFFFu);
  }
}
// 8 consecutive channels (16 bytes per plane) <-> fp32
__device__ __forceinline__ void act_load8(const __nv_bfloat16* p, long long plane_stride, int planes, float* v) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  uint4 ql = make_uint4(0, 0, 0, 0);
  if (planes == 2) ql = __ldg(reinterpret_cast<const uint4*>(p + plane_stride));
  const uint32_t* h = reinterpret_cast<const uint32_t*>(&q);
  const uint32_t* l = reinterpret_cast<const uint32_t*>(&ql);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = act_unpack2(h[e], l[e], planes);
    v[2 * e] = f.x; v[2 * e + 1] = f.y;
  }
}
__device__ __forceinline__ void act_store8(__nv_bfloat16* p, long long plane_stride, int planes, const float* v) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) act_pack2(v[2 * e], v[2 * e + 1], planes, hi[e], lo[e]);
  *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (planes == 2) *reinterpret_cast<uint4*>(p + plane_stride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a malformed descriptor must not hang the GPU box -- after ~2^31 SM cycles
// (about a second) the kernel traps (reported as a launch failure by the next CUDA call).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFF) == 0 && clock64() - t0 > (1ll << 31)) {
      printf("v2x: mbarrier timeout (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// one lane of a fully converged warp (warp-uniform control flow keeps descriptors in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
      "[%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// launch_dependents: the next kernel in the stream (launched with the programmatic-serialization attribute) may start
// its CTAs as soon as every CTA of this grid has executed this (or exited) -- its prologue (barrier init, TMEM
// allocation, resident-weight loads) and, SM by SM, its first tiles overlap this grid's tail.
// wait: blocks until every prerequisite grid has COMPLETED and its memory is visible; executed before the first
// access to anything an earlier kernel produced.  Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();   // host: V2X_NO_PDL=1 disables the launch attribute (A/B switch)

// ---- tcgen05 / TMEM ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, cta_group::1
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on `bar` when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 16 consecutive fp32 columns: thread t of the warp gets lane (lane_base + t)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// asynchronous variant: issue several loads, then tmem_ld_wait16() on each destination before use
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// waits for all outstanding tcgen05.ld of this thread; the "+r" operands pin the data dependency so the
// compiler cannot read the destination registers before the wait
__device__ __forceinline__ void tmem_ld_wait16(float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// K-major shared-memory matrix descriptor (sm_100 "version 1"), rows of `swizzle_bytes` each,
// 8-row groups `sbo_bytes` apart.  layout_type: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);         // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                             // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;   // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                             // descriptor version 1 (Blackwell)
  d |= (uint64_t)(layout_type & 7) << 61;             // swizzle mode, bits [61,64)
  return d;
}

// kind::f16 instruction descriptor: bf16 A/B (K-major), fp32 accumulate, M = 128
__host__ __device__ constexpr uint32_t make_idesc_bf16_m128(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
// same with fp16 A/B (a_format = b_format = 0)
__host__ __device__ constexpr uint32_t make_idesc_f16_m128(uint32_t n) {
  return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
// operand format follows the storage format: 1 plane = bf16, 2 planes = fp16 hi/lo
template <int PLANES>
__host__ __device__ constexpr uint32_t make_idesc_m128(uint32_t n) {
  return PLANES == 2 ? make_idesc_f16_m128(n) : make_idesc_bf16_m128(n);
}

}  // namespace v2x
