// peer_kernels.cu -- the x_3 exchange of the unit-sharded plans as DEVICE-SIDE PUSHES over NVLink peer memory
// (one process per GPU; SURVEY 8(e): "ncclAllGather of x_3 ... issued right after conv3_2").
//
// Every rank owns a peer-visible region (cudaMalloc + CUDA IPC handle, opened by the other ranks of the node):
//   [flag block: ready[8], consumed[8] (uint64 step numbers, one slot per writing rank) | x3_all act tensor].
// One exchange, all on the launching stream, no host involvement (so the whole forward is ONE CUDA graph):
//   peer_begin   : step += 1; wait until every peer has CONSUMED step - 1 (its fuse kernel finished reading my old maps
//                  out of ITS region -- the region is single-buffered)
//   peer_push    : copy this rank's x_3 units into its slot of EVERY rank's x3_all (own region: all planes; remote
//                  regions: the planes the consumer reads) with 16-byte stores that travel over NVLink; when the last CTA
//                  has finished (fence.sys + device-scope counter), it publishes ready[rank] = step in every region
//   peer_wait    : wait until ready[r] >= step for every rank r: all maps of this step have landed in my region
//   (fuse kernel reads x3_all)
//   peer_done    : publish consumed[rank] = step in every region
// Signals precede waits in every rank's own stream order, so the protocol cannot deadlock as long as every rank runs
// the same number of forwards.  The spin loops give up after `timeout_ms` and raise an error flag instead of hanging
// the GPU (the host checks it: sharding.PeerRegion.check()).
#include <cstring>

#include "common.cuh"

namespace v2x {

constexpr int kMaxPeers = 8;
constexpr int kFlagReady = 0, kFlagConsumed = kMaxPeers;   // uint64 slots inside a region's flag block

struct PeerFlags {
  unsigned long long* region[kMaxPeers];   // flag block of every rank's region (region[rank] = the local one)
};

struct PeerPush {
  const uint4* src;                 // local x_3: [planes][units_local * HWC] in 16-byte units
  long long src_plane_stride;      // 16-byte units between planes of src
  long long units;                  // 16-byte units per plane (units_local * HWC / 8)
  uint4* dst[kMaxPeers];            // this rank's slot (plane 0) in every rank's x3_all
  int dst_planes[kMaxPeers];        // planes pushed to that rank
  long long dst_plane_stride;      // 16-byte units between planes of x3_all
  int world, rank;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// thread r waits for flags[which + r] >= target (r != rank); gives up after timeout_ns and raises *err
__device__ __forceinline__ void spin_until(const unsigned long long* flags, int which, int r, unsigned long long target,
                                           unsigned long long timeout_ns, int* err, int code) {
  const unsigned long long t0 = global_timer_ns();
  while (ld_acquire_sys(flags + which + r) < target) {
    if (global_timer_ns() - t0 > timeout_ns) {
      atomicExch(err, code + r);
      break;
    }
    __nanosleep(64);
  }
}

__global__ void peer_begin_kernel(const unsigned long long* flags, unsigned long long* step, int rank, int world,
                                  unsigned long long timeout_ns, int* err) {
  const unsigned long long s = *step + 1;
  const int r = threadIdx.x;
  if (r < world && r != rank) spin_until(flags, kFlagConsumed, r, s - 1, timeout_ns, err, 100);
  __syncthreads();
  if (threadIdx.x == 0) *step = s;
}

__global__ void __launch_bounds__(256) peer_push_kernel(const PeerPush a, const PeerFlags f, const unsigned long long* step,
                                                        unsigned int* counter) {
  // work items: (destination rank, plane, 16-byte unit); a CTA stays on one destination for a long run of units
  long long total = 0;
  for (int d = 0; d < a.world; ++d) total += (long long)a.dst_planes[d] * a.units;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    long long rem = g;
    int d = 0;
    while (rem >= (long long)a.dst_planes[d] * a.units) { rem -= (long long)a.dst_planes[d] * a.units; ++d; }
    const int pl = (int)(rem / a.units);
    const long long u = rem - (long long)pl * a.units;
    a.dst[d][pl * a.dst_plane_stride + u] = __ldg(a.src + pl * a.src_plane_stride + u);
  }
  // completion: every thread's stores are ordered before its CTA's ticket; the CTA drawing the last ticket publishes
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence_system();
    const unsigned long long s = *step;
    if (threadIdx.x < a.world) st_release_sys(f.region[threadIdx.x] + kFlagReady + a.rank, s);
    if (threadIdx.x == 0) *counter = 0u;   // re-armed for the next step (stream order separates the launches)
  }
}

__global__ void peer_wait_kernel(const unsigned long long* flags, const unsigned long long* step, int rank, int world,
                                 unsigned long long timeout_ns, int* err) {
  const unsigned long long s = *step;
  const int r = threadIdx.x;
  if (r < world && r != rank) spin_until(flags, kFlagReady, r, s, timeout_ns, err, 200);
}

__global__ void peer_done_kernel(const PeerFlags f, const unsigned long long* step, int rank, int world) {
  const unsigned long long s = *step;
  __threadfence_system();
  if (threadIdx.x < world) st_release_sys(f.region[threadIdx.x] + kFlagConsumed + rank, s);
}

static int fill_flags(PeerFlags& f, void* const* host_regions, int world) {
  for (int r = 0; r < kMaxPeers; ++r) f.region[r] = r < world ? reinterpret_cast<unsigned long long*>(host_regions[r]) : nullptr;
  for (int r = 0; r < world; ++r)
    if (f.region[r] == nullptr) return -1;
  return 0;
}

}  // namespace v2x

using namespace v2x;

extern "C" int v2x_peer_alloc(int64_t bytes, void** host_ptr_out, void* host_handle_out) {
  V2X_REQUIRE(bytes > 0 && host_ptr_out && host_handle_out, "null/empty");
  void* p = nullptr;
  V2X_CUDA_TRY(cudaMalloc(&p, (size_t)bytes));
  V2X_CUDA_TRY(cudaMemset(p, 0, (size_t)bytes));
  V2X_CUDA_TRY(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return cuda_fail(e, "cudaIpcGetMemHandle");
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == V2X_PEER_HANDLE_BYTES, "handle size");
  memcpy(host_handle_out, &h, sizeof(h));
  *host_ptr_out = p;
  return V2X_OK;
}

extern "C" int v2x_peer_open(const void* host_handle, void** host_ptr_out) {
  V2X_REQUIRE(host_handle && host_ptr_out, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, host_handle, sizeof(h));
  void* p = nullptr;
  V2X_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *host_ptr_out = p;
  return V2X_OK;
}

extern "C" int v2x_peer_close(void* ptr) {
  V2X_REQUIRE(ptr, "null pointer");
  V2X_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return V2X_OK;
}

extern "C" int v2x_peer_free(void* ptr) {
  V2X_REQUIRE(ptr, "null pointer");
  V2X_CUDA_TRY(cudaFree(ptr));
  return V2X_OK;
}

extern "C" int v2x_peer_begin(const void* flags_local, void* step, int32_t rank, int32_t world, int32_t timeout_ms,
                              int32_t* err, void* stream) {
  V2X_REQUIRE(flags_local && step && err, "null pointer");
  V2X_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "1 <= world <= 8, 0 <= rank < world");
  peer_begin_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(flags_local),
                                                      reinterpret_cast<unsigned long long*>(step), rank, world,
                                                      (unsigned long long)timeout_ms * 1000000ull, err);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_peer_push(const void* src, int64_t elems_per_plane, int64_t src_plane_stride, void* const* host_dst,
                             const int32_t* host_dst_planes, int64_t dst_plane_stride, void* const* host_flag_regions,
                             const void* step, void* counter, int32_t rank, int32_t world, void* stream) {
  V2X_REQUIRE(src && host_dst && host_dst_planes && host_flag_regions && step && counter, "null pointer");
  V2X_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "1 <= world <= 8, 0 <= rank < world");
  V2X_REQUIRE(elems_per_plane > 0 && elems_per_plane % 8 == 0 && src_plane_stride % 8 == 0 && dst_plane_stride % 8 == 0,
              "plane sizes / strides must be multiples of 8 elements (16 bytes)");
  PeerPush a;
  a.src = reinterpret_cast<const uint4*>(src);
  a.src_plane_stride = src_plane_stride / 8;
  a.units = elems_per_plane / 8;
  a.dst_plane_stride = dst_plane_stride / 8;
  a.world = world;
  a.rank = rank;
  long long total = 0;
  for (int r = 0; r < kMaxPeers; ++r) {
    a.dst[r] = r < world ? reinterpret_cast<uint4*>(host_dst[r]) : nullptr;
    a.dst_planes[r] = r < world ? host_dst_planes[r] : 0;
    V2X_REQUIRE(r >= world || (a.dst[r] && a.dst_planes[r] >= 0 && a.dst_planes[r] <= 2), "bad destination %d", r);
    V2X_REQUIRE((reinterpret_cast<uintptr_t>(a.dst[r]) & 15) == 0, "destination %d is not 16-byte aligned", r);
    total += (long long)a.dst_planes[r] * a.units;
  }
  V2X_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0, "src is not 16-byte aligned");
  PeerFlags f;
  V2X_REQUIRE(fill_flags(f, host_flag_regions, world) == 0, "null flag region");
  long long blocks = (total + 256 * 8 - 1) / (256 * 8);   // ~8 stores per thread
  if (blocks > 2 * 148) blocks = 2 * 148;
  if (blocks < 1) blocks = 1;
  peer_push_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, f, reinterpret_cast<const unsigned long long*>(step),
                                                                      reinterpret_cast<unsigned int*>(counter));
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_peer_wait(const void* flags_local, const void* step, int32_t rank, int32_t world, int32_t timeout_ms,
                             int32_t* err, void* stream) {
  V2X_REQUIRE(flags_local && step && err, "null pointer");
  V2X_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "1 <= world <= 8, 0 <= rank < world");
  peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(flags_local),
                                                     reinterpret_cast<const unsigned long long*>(step), rank, world,
                                                     (unsigned long long)timeout_ms * 1000000ull, err);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}

extern "C" int v2x_peer_done(void* const* host_flag_regions, const void* step, int32_t rank, int32_t world, void* stream) {
  V2X_REQUIRE(host_flag_regions && step, "null pointer");
  V2X_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "1 <= world <= 8, 0 <= rank < world");
  PeerFlags f;
  V2X_REQUIRE(fill_flags(f, host_flag_regions, world) == 0, "null flag region");
  peer_done_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, reinterpret_cast<const unsigned long long*>(step), rank, world);
  V2X_CUDA_TRY(cudaGetLastError());
  return V2X_OK;
}
