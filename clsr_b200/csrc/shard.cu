// Row-sharded embedding tables for the configurations whose item table outgrows one GPU's HBM
// (BASELINE configs 4 and 5).  Global row r lives on rank r % world at local
// row r / world (round-robin: the frequency-sorted vocabulary, sequential_reviews.py:114-140, would
// pile every hot id onto rank 0 under a block partition).
//
// The reference has no multi-device path (tf.nn.embedding_lookup on one device,
// sequential_base_model.py:381-437).  Instead of a dedup -> all-to-all(ids) -> gather ->
// all-to-all(rows) -> scatter pipeline, every rank maps its peers' shards into its own address
// space (CUDA IPC over NVLink / NVSwitch peer memory) and ONE kernel gathers straight from the
// owning GPU with 16-byte loads: the transfer is the gather.  The backward pass is the mirror image:
// red.global.add.v4.f32 straight into the owner's gradient shard.  No staging buffers, no id
// exchange, no host synchronisation inside the op; ordering against the owners' own reads/writes is
// the caller's job (one stream-ordered barrier per step, see clsr_b200/sharded.py).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "../../include/clsr_b200.h"
#include "common.cuh"

using namespace clsr;

struct clsr_shard_table {
  int device = 0, rank = 0, world = 1, dim = 0;
  long long n_rows = 0, local_rows = 0;
  float* values = nullptr;
  float* grad = nullptr;
  float* peer_values[CLSR_SHARD_MAX_WORLD] = {};
  float* peer_grad[CLSR_SHARD_MAX_WORLD] = {};
  bool attached = false;
  std::string err;
};

namespace {

thread_local std::string g_shard_error;

int sfail(clsr_shard_table* t, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (t) t->err = buf; else g_shard_error = buf;
  return code;
}

#define SCK(t, call)                                                                                   \
  do {                                                                                                 \
    cudaError_t _c = (call);                                                                           \
    if (_c != cudaSuccess) return sfail(t, CLSR_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(_c)); \
  } while (0)

struct Peers {
  float* p[CLSR_SHARD_MAX_WORLD];
};

// owner / local row of a global id; world is a power of two on every NVSwitch box we target, the
// general division stays for odd test sizes
struct Split {
  int world, shift, mask;
  __device__ __forceinline__ void operator()(int id, int& owner, long long& local) const {
    if (shift >= 0) { owner = id & mask; local = id >> shift; }
    else { owner = id % world; local = id / world; }
  }
};
Split make_split(int world) {
  Split s; s.world = world; s.shift = -1; s.mask = 0;
  for (int b = 0; b < 8; ++b)
    if ((1 << b) == world) { s.shift = b; s.mask = world - 1; }
  return s;
}

// hist[p, :] = concat(item[ih[p]], cate[ch[p]]) with both tables row-sharded over peer memory.
// Same thread mapping as gather_hist_kernel (embed.cuh): one 16-byte vector per thread, the V
// threads of a position read one contiguous row segment per table and write one contiguous output
// row; UNROLL independent remote loads are in flight per thread (NVLink latency is ~2x local HBM).
template <int UNROLL>
__global__ void __launch_bounds__(256)
shard_gather_hist_kernel(const int32_t* __restrict__ ih, const int32_t* __restrict__ ch, Peers item, Peers cate,
                         Split sp, int Di, int Dc, float* __restrict__ out, long long npos) {
  const int VI = Di >> 2, V = (Di + Dc) >> 2;
  const long long nvec = npos * V;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; g0 < nvec; g0 += stride * UNROLL) {
    const float4* src[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      long long g = g0 + u * stride;
      g = g < nvec ? g : nvec - 1;   // clamped: the load is always issued, the store is predicated
      const long long p = g / V;
      const int q = (int)(g - p * V);
      int owner;
      long long local;
      if (q < VI) {
        sp(__ldg(ih + p), owner, local);
        src[u] = reinterpret_cast<const float4*>(item.p[owner] + (size_t)local * Di) + q;
      } else {
        sp(__ldg(ch + p), owner, local);
        src[u] = reinterpret_cast<const float4*>(cate.p[owner] + (size_t)local * Dc) + (q - VI);
      }
    }
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = __ldg(src[u]);   // L1-allocating: the hot (low, frequency-sorted) ids and
                                                            // the padding row are re-read from L1, not over NVLink
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long g = g0 + u * stride;
      if (g < nvec) stg_stream(reinterpret_cast<float4*>(out) + g, v[u]);
    }
  }
}

// grad_item[owner(ih[p])][local(ih[p]), :] += d_hist[p, 0:Di]; same for the category columns.
// d_hist is streamed once with 16-byte loads; every vector becomes one 16-byte reduction that the
// NVSwitch fabric carries to the owning GPU's L2.  Ids are popularity ranks
// (sequential_reviews.py:114-140) and id 0 pads every window, so the `hot` lowest item ids (and
// category 0) are first summed per CTA in shared memory and leave the SM once per CTA instead of
// once per position: without it every rank hammers the same few remote rows.
__global__ void __launch_bounds__(256)
shard_scatter_hist_kernel(const float* __restrict__ dX, const int32_t* __restrict__ ih, const int32_t* __restrict__ ch,
                          Peers gitem, Peers gcate, Split sp, int Di, int Dc, long long npos, int hot) {
  extern __shared__ float hot_acc[];   // [hot][Di] item rows, then [Dc] category row 0
  const int VI = Di >> 2, V = (Di + Dc) >> 2;
  const int nacc = hot * Di + Dc;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) hot_acc[i] = 0.f;
  __syncthreads();
  const long long nvec = npos * V;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < nvec;
       g += (long long)gridDim.x * blockDim.x) {
    const long long p = g / V;
    const int q = (int)(g - p * V);
    const float4 v = ldg_stream(reinterpret_cast<const float4*>(dX) + g);
    int owner;
    long long local;
    if (q < VI) {
      const int id = __ldg(ih + p);
      if (id < hot) {
        float* a = hot_acc + id * Di + q * 4;
        atomicAdd(a + 0, v.x); atomicAdd(a + 1, v.y); atomicAdd(a + 2, v.z); atomicAdd(a + 3, v.w);
      } else {
        sp(id, owner, local);
        red_add_v4(gitem.p[owner] + (size_t)local * Di + q * 4, v);
      }
    } else {
      const int id = __ldg(ch + p);
      if (id == 0) {
        float* a = hot_acc + hot * Di + (q - VI) * 4;
        atomicAdd(a + 0, v.x); atomicAdd(a + 1, v.y); atomicAdd(a + 2, v.z); atomicAdd(a + 3, v.w);
      } else {
        sp(id, owner, local);
        red_add_v4(gcate.p[owner] + (size_t)local * Dc + (q - VI) * 4, v);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x * 4; i < nacc; i += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(hot_acc + i);
    if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
    int owner;
    long long local;
    if (i < hot * Di) {
      const int id = i / Di, c = i - id * Di;
      sp(id, owner, local);
      red_add_v4(gitem.p[owner] + (size_t)local * Di + c, v);
    } else {
      sp(0, owner, local);
      red_add_v4(gcate.p[owner] + (size_t)local * Dc + (i - hot * Di), v);
    }
  }
}

// ---- bulk (TMA) form of the scatter-add ----------------------------------------------------------------
// The element-wise kernel above sends one 16-byte reduction per vector; over NVLink that is packet-rate
// bound (measured 12 G reductions/s per GPU at N = 2 against 67 G/s locally).  Here one warp owns a
// position at a time: lane 0 pulls the position's whole gradient row (D * 4 bytes, contiguous in d_hist)
// into a shared-memory slot with cp.async.bulk, and pushes the item part and the category part to their
// owners with ONE cp.reduce.async.bulk.add.f32 each - the fabric carries a 448-byte reduction instead of
// 28 sixteen-byte ones.  NSLOT positions per warp are in flight.  Hot ids (popularity ranks < hot, category
// 0) are still summed per CTA in shared memory first, cooperatively by the 32 lanes.
// Measured (config-4 rows, 2 GPUs): same 0.53 ms as the element-wise kernel with 28x fewer reduction
// instructions; without the padding-row cache either form takes 1.4 ms.  The remaining cost scales with the
// remote share, not with packet count or hot-row count (A/B in DESIGN.md section 7).
constexpr int SB_WARPS = 8, SB_NSLOT = 4;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sb_mbar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(bar)));
}
__device__ __forceinline__ void sb_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sb_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SB_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SB_WAIT_DONE;\n"
      "bra SB_WAIT_LOOP;\n"
      "SB_WAIT_DONE:\n"
      "}\n" ::"r"(s_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void sb_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(s_u32(bar))
               : "memory");
}
__device__ __forceinline__ void sb_reduce_add(float* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(s_u32(smem_src)),
               "r"(bytes)
               : "memory");
}

__global__ void __launch_bounds__(SB_WARPS * 32)
shard_scatter_bulk_kernel(const float* __restrict__ dX, const int32_t* __restrict__ ih, const int32_t* __restrict__ ch,
                          Peers gitem, Peers gcate, Split sp, int Di, int Dc, long long npos, int hot) {
  extern __shared__ __align__(128) uint8_t smraw[];
  const int D = Di + Dc;
  const int nacc = hot * Di + Dc;
  float* slots = reinterpret_cast<float*>(smraw);                                  // [SB_WARPS][SB_NSLOT][D]
  float* hot_acc = slots + (size_t)SB_WARPS * SB_NSLOT * D;                         // [hot][Di] + [Dc]
  uint64_t* bars = reinterpret_cast<uint64_t*>(hot_acc + ((nacc + 3) & ~3));        // [SB_WARPS][SB_NSLOT]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) hot_acc[i] = 0.f;
  if (lane == 0)
    for (int k = 0; k < SB_NSLOT; ++k) sb_mbar_init(&bars[warp * SB_NSLOT + k]);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  float* myslot = slots + (size_t)warp * SB_NSLOT * D;
  uint64_t* mybar = bars + warp * SB_NSLOT;
  const long long gw = (long long)blockIdx.x * SB_WARPS + warp, GW = (long long)gridDim.x * SB_WARPS;
  uint32_t phase = 0;
  for (long long p0 = gw; p0 < npos; p0 += GW * SB_NSLOT) {
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < SB_NSLOT; ++k) {
        const long long p = p0 + (long long)k * GW;
        if (p < npos) {
          sb_expect_tx(&mybar[k], (uint32_t)D * 4);
          sb_load(myslot + (size_t)k * D, dX + (size_t)p * D, (uint32_t)D * 4, &mybar[k]);
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < SB_NSLOT; ++k) {
      const long long p = p0 + (long long)k * GW;
      if (p >= npos) continue;
      sb_wait(&mybar[k], phase);
      const float* row = myslot + (size_t)k * D;
      const int idi = __ldg(ih + p), idc = __ldg(ch + p);
      int owner;
      long long local;
      if (idi < hot) {
        for (int c = lane * 4; c < Di; c += 128) {
          const float4 v = *reinterpret_cast<const float4*>(row + c);
          float* a = hot_acc + idi * Di + c;
          atomicAdd(a + 0, v.x); atomicAdd(a + 1, v.y); atomicAdd(a + 2, v.z); atomicAdd(a + 3, v.w);
        }
      } else if (lane == 0) {
        sp(idi, owner, local);
        sb_reduce_add(gitem.p[owner] + (size_t)local * Di, row, (uint32_t)Di * 4);
      }
      if (idc == 0) {
        for (int c = lane * 4; c < Dc; c += 128) {
          const float4 v = *reinterpret_cast<const float4*>(row + Di + c);
          float* a = hot_acc + hot * Di + c;
          atomicAdd(a + 0, v.x); atomicAdd(a + 1, v.y); atomicAdd(a + 2, v.z); atomicAdd(a + 3, v.w);
        }
      } else if (lane == 0) {
        sp(idc, owner, local);
        sb_reduce_add(gcate.p[owner] + (size_t)local * Dc, row + Di, (uint32_t)Dc * 4);
      }
    }
    // the slots are rewritten by the next batch: the bulk reductions must have read them, and the lanes'
    // own (generic-proxy) reads must be ordered before the next async-proxy writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (lane == 0) {
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
    phase ^= 1;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncthreads();
  for (int i = threadIdx.x * 4; i < nacc; i += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(hot_acc + i);
    if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
    int owner;
    long long local;
    if (i < hot * Di) {
      const int id = i / Di, c = i - id * Di;
      sp(id, owner, local);
      red_add_v4(gitem.p[owner] + (size_t)local * Di + c, v);
    } else {
      sp(0, owner, local);
      red_add_v4(gcate.p[owner] + (size_t)local * Dc + (i - hot * Di), v);
    }
  }
}

int check_pair(clsr_shard_table* a, clsr_shard_table* b) {
  if (!a || !b) return sfail(a, CLSR_ERR_ARG, "null table");
  if (!a->attached || !b->attached) return sfail(a, CLSR_ERR_STATE, "peer shards not attached (clsr_shard_attach)");
  if (a->world != b->world || a->rank != b->rank || a->device != b->device)
    return sfail(a, CLSR_ERR_ARG, "item and category shards belong to different ranks / devices");
  return 0;
}

}  // namespace

extern "C" {

int clsr_shard_create(int32_t device, int32_t rank, int32_t world, int64_t n_rows, int32_t dim, int32_t with_grad,
                      clsr_shard_table** out) {
  if (!out) return sfail(nullptr, CLSR_ERR_ARG, "null out pointer");
  *out = nullptr;
  if (world < 1 || world > CLSR_SHARD_MAX_WORLD || rank < 0 || rank >= world)
    return sfail(nullptr, CLSR_ERR_ARG, "bad rank %d / world %d (max %d)", rank, world, CLSR_SHARD_MAX_WORLD);
  if (n_rows <= 0 || n_rows > 0x7fffffffLL || dim <= 0 || (dim & 3))
    return sfail(nullptr, CLSR_ERR_ARG, "rows must fit int32 ids and dim must be a multiple of 4 (16-byte rows)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return sfail(nullptr, CLSR_ERR_CUDA, "no CUDA device: clsr_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return sfail(nullptr, CLSR_ERR_ARG, "bad device %d", device);
  clsr_shard_table* t = new clsr_shard_table();
  t->device = device; t->rank = rank; t->world = world; t->dim = dim; t->n_rows = n_rows;
  t->local_rows = (n_rows - rank + world - 1) / world;
  if (t->local_rows < 1) t->local_rows = 1;
  cudaError_t c = cudaSetDevice(device);
  const size_t bytes = (size_t)t->local_rows * dim * sizeof(float);
  if (c == cudaSuccess) c = cudaMalloc(&t->values, bytes);
  if (c == cudaSuccess) c = cudaMemset(t->values, 0, bytes);
  if (c == cudaSuccess && with_grad) c = cudaMalloc(&t->grad, bytes);
  if (c == cudaSuccess && with_grad) c = cudaMemset(t->grad, 0, bytes);
  if (c != cudaSuccess) {
    sfail(nullptr, CLSR_ERR_CUDA, "shard allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(c));
    clsr_shard_destroy(t);
    return CLSR_ERR_CUDA;
  }
  t->peer_values[rank] = t->values;
  t->peer_grad[rank] = t->grad;
  t->attached = world == 1;
  *out = t;
  return CLSR_OK;
}

void clsr_shard_destroy(clsr_shard_table* t) {
  if (!t) return;
  cudaSetDevice(t->device);
  for (int r = 0; r < t->world; ++r) {
    if (r == t->rank) continue;
    if (t->peer_values[r]) cudaIpcCloseMemHandle(t->peer_values[r]);
    if (t->peer_grad[r]) cudaIpcCloseMemHandle(t->peer_grad[r]);
  }
  if (t->values) cudaFree(t->values);
  if (t->grad) cudaFree(t->grad);
  delete t;
}

const char* clsr_shard_last_error(const clsr_shard_table* t) { return t ? t->err.c_str() : g_shard_error.c_str(); }
int64_t clsr_shard_local_rows(const clsr_shard_table* t) { return t ? t->local_rows : 0; }
float* clsr_shard_local_values(clsr_shard_table* t) { return t ? t->values : nullptr; }
float* clsr_shard_local_grad(clsr_shard_table* t) { return t ? t->grad : nullptr; }

int clsr_shard_export(clsr_shard_table* t, void* handles_out) {
  if (!t || !handles_out) return sfail(t, CLSR_ERR_ARG, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  SCK(t, cudaSetDevice(t->device));
  cudaIpcMemHandle_t h[2];
  memset(h, 0, sizeof h);
  SCK(t, cudaIpcGetMemHandle(&h[0], t->values));
  if (t->grad) SCK(t, cudaIpcGetMemHandle(&h[1], t->grad));
  memcpy(handles_out, h, sizeof h);
  return CLSR_OK;
}

int clsr_shard_attach(clsr_shard_table* t, const void* all_handles) {
  if (!t || !all_handles) return sfail(t, CLSR_ERR_ARG, "bad argument");
  if (t->attached) return CLSR_OK;
  SCK(t, cudaSetDevice(t->device));
  const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(all_handles);
  for (int r = 0; r < t->world; ++r) {
    if (r == t->rank) continue;
    void* p = nullptr;
    SCK(t, cudaIpcOpenMemHandle(&p, h[2 * r], cudaIpcMemLazyEnablePeerAccess));
    t->peer_values[r] = static_cast<float*>(p);
    if (t->grad) {
      SCK(t, cudaIpcOpenMemHandle(&p, h[2 * r + 1], cudaIpcMemLazyEnablePeerAccess));
      t->peer_grad[r] = static_cast<float*>(p);
    }
  }
  t->attached = true;
  return CLSR_OK;
}

int clsr_shard_zero_grad(clsr_shard_table* t, void* stream) {
  if (!t || !t->grad) return sfail(t, CLSR_ERR_ARG, "table has no gradient shard");
  SCK(t, cudaSetDevice(t->device));
  SCK(t, cudaMemsetAsync(t->grad, 0, (size_t)t->local_rows * t->dim * sizeof(float), (cudaStream_t)stream));
  return CLSR_OK;
}

int clsr_shard_gather_history(clsr_shard_table* item, clsr_shard_table* cate, const int32_t* ih, const int32_t* ch,
                              int64_t positions, float* out, void* stream) {
  int rc = check_pair(item, cate);
  if (rc) return rc;
  if (!ih || !ch || !out || positions <= 0) return sfail(item, CLSR_ERR_ARG, "bad argument");
  SCK(item, cudaSetDevice(item->device));
  Peers pi, pc;
  for (int r = 0; r < CLSR_SHARD_MAX_WORLD; ++r) { pi.p[r] = item->peer_values[r]; pc.p[r] = cate->peer_values[r]; }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, item->device);
  const long long nvec = positions * ((item->dim + cate->dim) / 4);
  long long want = (nvec + 256LL * 4 - 1) / (256LL * 4);
  const long long cap = (long long)sms * 8;
  const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
  shard_gather_hist_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(ih, ch, pi, pc, make_split(item->world), item->dim,
                                                                     cate->dim, out, positions);
  SCK(item, cudaGetLastError());
  return CLSR_OK;
}

int clsr_shard_scatter_add_history(clsr_shard_table* item, clsr_shard_table* cate, const int32_t* ih,
                                   const int32_t* ch, int64_t positions, const float* d_hist, void* stream) {
  int rc = check_pair(item, cate);
  if (rc) return rc;
  if (!item->grad || !cate->grad) return sfail(item, CLSR_ERR_STATE, "tables were created without gradient shards");
  if (!ih || !ch || !d_hist || positions <= 0) return sfail(item, CLSR_ERR_ARG, "bad argument");
  SCK(item, cudaSetDevice(item->device));
  Peers gi, gc;
  for (int r = 0; r < CLSR_SHARD_MAX_WORLD; ++r) { gi.p[r] = item->peer_grad[r]; gc.p[r] = cate->peer_grad[r]; }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, item->device);
  const long long nvec = positions * ((item->dim + cate->dim) / 4);
  long long want = (nvec + 255) / 256;
  const long long cap = (long long)sms * 4;   // few, long-lived CTAs: each flushes its hot rows once
  const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
  // per-CTA pre-aggregation of the hottest item ids: up to 64 rows / 32 KB of shared memory
  int hot = (32 * 1024) / (item->dim * 4);
  if (hot > 64) hot = 64;
  if ((long long)hot > item->n_rows) hot = (int)item->n_rows;
  static const bool use_red = getenv("CLSR_SHARD_SCATTER_RED") != nullptr;   // element-wise reductions (A/B, fallback)

  if (!use_red) {
    const int D = item->dim + cate->dim;
    const int nacc = hot * item->dim + cate->dim;
    const size_t smem = (size_t)SB_WARPS * SB_NSLOT * D * 4 + (size_t)((nacc + 3) & ~3) * 4 + SB_WARPS * SB_NSLOT * 8;
    if (smem <= 200 * 1024) {
      static bool attr_set[64] = {};
      if (!attr_set[item->device & 63]) {
        SCK(item, cudaFuncSetAttribute(shard_scatter_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[item->device & 63] = true;
      }
      long long wantb = (positions + SB_WARPS * SB_NSLOT - 1) / (SB_WARPS * SB_NSLOT);
      const long long capb = (long long)sms * 4;
      const int gridb = (int)(wantb < capb ? (wantb < 1 ? 1 : wantb) : capb);
      shard_scatter_bulk_kernel<<<gridb, SB_WARPS * 32, smem, (cudaStream_t)stream>>>(d_hist, ih, ch, gi, gc, make_split(item->world),
                                                                                      item->dim, cate->dim, positions, hot);
      SCK(item, cudaGetLastError());
      return CLSR_OK;
    }
  }
  const size_t smem = (size_t)(hot * item->dim + cate->dim) * sizeof(float);
  shard_scatter_hist_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(d_hist, ih, ch, gi, gc, make_split(item->world),
                                                                      item->dim, cate->dim, positions, hot);
  SCK(item, cudaGetLastError());
  return CLSR_OK;
}

}  // extern "C"
