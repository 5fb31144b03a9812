// Embedding-side kernels: history gather (K1+K3), row gathers, id de-duplication through a
// direct-address slot table (K2), sparse-gradient scatter-add (K13), involved-row terms
// (L2 / discrepancy), and the sparse Adam updates (K15).
//
// Reference semantics: tf.nn.embedding_lookup / tf.unique / IndexedSlices gradients
// (sequential_base_model.py:381-437, clsr.py:103-127) and AdamOptimizer's sparse path
// (base_model.py:263-297); see SURVEY.md section 8a rows a2, a18.
#pragma once
#include "common.cuh"

namespace clsr {

// View of a (possibly row-sharded) table: global row r lives on rank r % world at local row r / world
// (round-robin: ids are popularity ranks, sequential_reviews.py:114-140, a block partition would put every hot
// row on rank 0).  p[r] is rank r's shard mapped into this process (CUDA IPC over NVLink / NVSwitch peer
// memory); world is a power of two, so owner = id & mask, local row = id >> shift.  An unsharded table is
// the world = 1 case (mask 0, shift 0).  Gathers read straight from the owning GPU: the transfer IS the gather.
constexpr int kMaxWorld = 16;
struct TabView {
  const float* p[kMaxWorld];
  int shift, mask;
  CLSR_DEVINL const float* row(int id, int dim) const { return p[id & mask] + (size_t)(id >> shift) * dim; }
};

// hist_input[p, :] = concat(item_table[ih[p]], cate_table[ch[p]]); p = s*T + t.
// Index arrays are addressed as base[(p / T) * seq_stride + p % T] so a [B,T] feed can be read at
// every G-th row without a compaction pass.  One thread moves one 16-byte vector; the V = (Di+Dc)/4
// threads of a position read one 128-byte item row + one 32-byte category sector and write one
// contiguous D*4-byte output row, so both sides are fully coalesced.  UNROLL positions per thread
// keep UNROLL independent row loads in flight behind the dependent index loads.
template <int UNROLL>
__global__ void __launch_bounds__(256)
gather_hist_kernel(const int32_t* __restrict__ ih, const int32_t* __restrict__ ch, int seq_stride, int T,
                   TabView item_tab, TabView cate_tab, int Di, int Dc, float* __restrict__ out, long long npos) {
  const int VI = Di >> 2, V = (Di + Dc) >> 2;
  const long long nvec = npos * V;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; g0 < nvec; g0 += stride * UNROLL) {
    const float4* src[UNROLL];
    long long gi[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      long long g = g0 + u * stride;
      gi[u] = g;
      src[u] = nullptr;
      if (g < nvec) {
        long long p = g / V;
        int q = (int)(g - p * V);
        long long s = p / T;
        int t = (int)(p - s * T);
        long long io = s * seq_stride + t;
        if (q < VI) {
          src[u] = reinterpret_cast<const float4*>(item_tab.row(__ldg(ih + io), Di)) + q;
        } else {
          src[u] = reinterpret_cast<const float4*>(cate_tab.row(__ldg(ch + io), Dc)) + (q - VI);
        }
      }
    }
    float4 v[UNROLL];
    // local table: rows are touched once, stream them past L1; sharded table: L1-allocating loads, so that the
    // hot (low, frequency-sorted) ids and the padding row -- a third of all positions, all owned by rank 0 -- are
    // re-read from this SM's L1 instead of crossing NVLink every time
    if (item_tab.mask) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (src[u]) v[u] = __ldg(src[u]);
    } else {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (src[u]) v[u] = ldg_stream(src[u]);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      if (src[u]) stg_stream(reinterpret_cast<float4*>(out) + gi[u], v[u]);
  }
}

// ---- bulk-copy (TMA) form of the history gather -----------------------------------------------------------------
// One thread per position issues two cp.async.bulk copies (item row, category row: global -> shared, completion
// counted in bytes on an mbarrier) straight into the position's slot of a [TILE, D] shared-memory tile, so the tile
// IS the contiguous output block; one elected thread then writes the whole tile back with a single bulk store
// (shared -> global).  Two tiles are in flight per CTA: the row reads of tile k+1 are issued before tile k is
// waited for.  No registers hold data, the SM issues 2 copy instructions per position instead of ~20 load / store /
// address instructions, and several hundred row reads per SM are outstanding at any time.
CLSR_DEVINL uint32_t cvta_s(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
CLSR_DEVINL void g_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cvta_s(bar)), "r"(count));
}
CLSR_DEVINL void g_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cvta_s(bar)), "r"(bytes) : "memory");
}
CLSR_DEVINL void g_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "G_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra G_WAIT_DONE;\n"
      "bra G_WAIT_LOOP;\n"
      "G_WAIT_DONE:\n"
      "}\n" ::"r"(cvta_s(bar)),
      "r"(parity)
      : "memory");
}
CLSR_DEVINL void g_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(cvta_s(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(cvta_s(bar))
               : "memory");
}
CLSR_DEVINL void g_bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(cvta_s(smem_src)), "r"(bytes)
               : "memory");
}

constexpr int kGatherTile = 128;   // positions per tile = threads per CTA

// `hot`: the lowest item ids (ids are popularity ranks, sequential_reviews.py:114-140; id 0 pads every window -- a third
// of all positions) and category 0 are kept in shared memory per CTA and copied from there by the position's thread:
// thousands of bulk reads of one row would otherwise queue on a single L2 slice (or one peer's NVLink port).
__global__ void __launch_bounds__(kGatherTile)
gather_hist_tma_kernel(const int32_t* __restrict__ ih, const int32_t* __restrict__ ch, int seq_stride, int T,
                       TabView item_tab, TabView cate_tab, int Di, int Dc, int hot, float* __restrict__ out,
                       long long npos) {
  extern __shared__ __align__(128) uint8_t gsm[];
  const int D = Di + Dc;
  const uint32_t row_bytes = (uint32_t)D * 4, tile_bytes = row_bytes * kGatherTile;
  float* tile[2] = {reinterpret_cast<float*>(gsm), reinterpret_cast<float*>(gsm + tile_bytes)};
  float* hot_item = reinterpret_cast<float*>(gsm + 2 * (size_t)tile_bytes);   // [hot][Di]
  float* hot_cate = hot_item + (size_t)hot * Di;                               // [Dc]: category 0
  uint64_t* bar = reinterpret_cast<uint64_t*>(hot_cate + ((Dc + 3) & ~3));
  const int tid = threadIdx.x;
  if (tid == 0) {
    g_mbar_init(&bar[0], kGatherTile);   // every thread arrives once per tile, announcing the bytes of its own bulk reads
    g_mbar_init(&bar[1], kGatherTile);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < hot * Di; i += kGatherTile) hot_item[i] = __ldg(item_tab.row(i / Di, Di) + (i % Di));
  for (int i = tid; i < Dc; i += kGatherTile) hot_cate[i] = __ldg(cate_tab.row(0, Dc) + i);
  __syncthreads();
  const long long ntiles = (npos + kGatherTile - 1) / kGatherTile;
  auto issue = [&](long long tl, int buf) {
    const long long p0 = tl * kGatherTile;
    const int n = (int)(npos - p0 < kGatherTile ? npos - p0 : kGatherTile);
    int idi = 0, idc = 0;
    if (tid < n) {
      const long long p = p0 + tid;
      const long long sq = p / T;
      const long long io = sq * seq_stride + (p - sq * T);
      idi = __ldg(ih + io); idc = __ldg(ch + io);
    }
    const bool bi = tid < n && idi >= hot, bc = tid < n && idc != 0;
    float* dst = tile[buf] + (size_t)tid * D;
    if (tid < n && !bi) {
      const float4* srcv = reinterpret_cast<const float4*>(hot_item + (size_t)idi * Di);
      for (int q = 0; q < (Di >> 2); ++q) reinterpret_cast<float4*>(dst)[q] = srcv[q];
    }
    if (tid < n && !bc) {
      const float4* srcv = reinterpret_cast<const float4*>(hot_cate);
      for (int q = 0; q < (Dc >> 2); ++q) reinterpret_cast<float4*>(dst + Di)[q] = srcv[q];
    }
    // generic-proxy writes of the hot rows are ordered before the bulk store by this fence + the arrive below
    if (tid < n && (!bi || !bc)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    g_mbar_expect_tx(&bar[buf], (bi ? (uint32_t)Di * 4 : 0u) + (bc ? (uint32_t)Dc * 4 : 0u));
    if (bi) g_bulk_g2s(dst, item_tab.row(idi, Di), (uint32_t)Di * 4, &bar[buf]);
    if (bc) g_bulk_g2s(dst + Di, cate_tab.row(idc, Dc), (uint32_t)Dc * 4, &bar[buf]);
  };
  int it = 0;
  long long tl = blockIdx.x;
  if (tl < ntiles) issue(tl, 0);
  for (; tl < ntiles; tl += gridDim.x, ++it) {
    const int buf = it & 1;
    const long long nxt = tl + gridDim.x;
    if (nxt < ntiles) {
      // the other buffer was handed to a bulk store one iteration ago: it may be refilled once that store has READ it
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncthreads();
      issue(nxt, buf ^ 1);
    }
    if (tid == 0) {   // only the storing thread waits for the tile
      g_mbar_wait(&bar[buf], (uint32_t)((it >> 1) & 1));
      const long long p0 = tl * kGatherTile;
      const int n = (int)(npos - p0 < kGatherTile ? npos - p0 : kGatherTile);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      g_bulk_s2g(out + (size_t)p0 * D, tile[buf], (uint32_t)n * row_bytes);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// out[r, col0 : col0+dim] = table[idx[r*idx_stride], :]   (targets, users)
__global__ void gather_rows_kernel(const int32_t* __restrict__ idx, int idx_stride, TabView tab, int dim,
                                   float* __restrict__ out, int ldo, int col0, int rows) {
  const int V = dim >> 2;
  long long n = (long long)rows * V;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n;
       g += (long long)gridDim.x * blockDim.x) {
    int r = (int)(g / V), q = (int)(g % V);
    int id = __ldg(idx + (size_t)r * idx_stride);
    float4 v = __ldg(reinterpret_cast<const float4*>(tab.row(id, dim)) + q);
    *reinterpret_cast<float4*>(out + (size_t)r * ldo + col0 + q * 4) = v;
  }
}

// Device-resident feeds: copy what the step reads (one row per sequence, every row's target ids and
// labels) into the engine's compact staging block, validating every id against its table size.  An id
// outside [0, n) is replaced by 0 (the padding row) and reported through err[0] (bit 0 users, 1 items,
// 2 cates, 3 item_history, 4 cate_history) -- a foreign feed must not turn into out-of-bounds reads in the
// gathers or out-of-bounds writes through the slot tables.
struct StageFeed {
  const int32_t *users, *items, *cates, *ih, *ch, *mask;
  const float *tfa, *ttn, *labels;
  int32_t *o_ih, *o_ch, *o_mask, *o_users, *o_items, *o_cates;
  float *o_tfa, *o_ttn, *o_labels;
  int S, G, T, B;
  int seq_g;   // rows between consecutive sequences in the [*,T] source arrays (G for a replicated feed, 1 when
               // the copy engine has already picked every G-th row)
  long long n_items, n_cates, n_users;
  int32_t* err;
};
__global__ void stage_device_feed_kernel(StageFeed a) {
  const long long M = (long long)a.S * a.T;
  const long long stride = (long long)gridDim.x * blockDim.x;
  int bad = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) {
    const long long s = i / a.T;
    const long long src = s * a.seq_g * a.T + (i - s * a.T);
    int it = __ldg(a.ih + src), ct = __ldg(a.ch + src);
    if ((unsigned long long)(long long)it >= (unsigned long long)a.n_items) { bad |= 8; it = 0; }
    if ((unsigned long long)(long long)ct >= (unsigned long long)a.n_cates) { bad |= 16; ct = 0; }
    a.o_ih[i] = it; a.o_ch[i] = ct;
    a.o_mask[i] = __ldg(a.mask + src);
    a.o_tfa[i] = __ldg(a.tfa + src);
    a.o_ttn[i] = __ldg(a.ttn + src);
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.B; i += stride) {
    int it = __ldg(a.items + i), ct = __ldg(a.cates + i);
    if ((unsigned long long)(long long)it >= (unsigned long long)a.n_items) { bad |= 2; it = 0; }
    if ((unsigned long long)(long long)ct >= (unsigned long long)a.n_cates) { bad |= 4; ct = 0; }
    a.o_items[i] = it; a.o_cates[i] = ct;
    a.o_labels[i] = a.labels ? __ldg(a.labels + i) : 0.f;
    if (i < a.S) {
      int u = __ldg(a.users + i * a.G);
      if ((unsigned long long)(long long)u >= (unsigned long long)a.n_users) { bad |= 1; u = 0; }
      a.o_users[i] = u;
    }
  }
  if (bad) atomicOr(a.err, bad);
}

// dst[i] = src[(i / T) * seq_stride + i % T]: contiguous copy of a strided id array (all-gather staging)
__global__ void compact_ids_kernel(const int32_t* __restrict__ src, int T, int seq_stride, long long n,
                                   int32_t* __restrict__ dst) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    long long s = i / T;
    dst[i] = src[s * seq_stride + (i - s * T)];
  }
}

// tf.unique through a direct-address table: slot[id] = compact row (or -1).  The first thread to
// claim an id appends it to uniq[]; later kernels translate ids through slot[].
__global__ void mark_unique_kernel(const int32_t* __restrict__ ids, long long n, int T, int seq_stride,
                                   int32_t* __restrict__ slot, int32_t* __restrict__ uniq,
                                   int32_t* __restrict__ counter) {
  // uniform trip count so the warp stays converged for the ballot; the append counter is bumped once per
  // warp (a per-id atomicAdd on one address serialises ~10^5 atomics per step)
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long nround = (n + stride - 1) / stride * stride;
  const int lane = threadIdx.x & 31;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
    bool claim = false;
    int id = 0;
    if (i < n) {
      const long long s = i / T;
      id = __ldg(ids + s * seq_stride + (i - s * T));
      if (slot[id] == -1) claim = atomicCAS(&slot[id], -1, -2) == -1;
    }
    const unsigned m = __ballot_sync(0xffffffffu, claim);
    if (m) {
      const int leader = __ffs(m) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(counter, __popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (claim) {
        const int c = base + __popc(m & ((1u << lane) - 1u));
        uniq[c] = id;
        slot[id] = c;
      }
    }
  }
}

__global__ void reset_slots_kernel(const int32_t* __restrict__ uniq, const int32_t* __restrict__ counter,
                                   int32_t* __restrict__ slot) {
  int n = *counter;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) slot[uniq[c]] = -1;
}

// Zero the compact gradient rows that this step will use (count is device-resident).
__global__ void zero_compact_kernel(float* __restrict__ g, const int32_t* __restrict__ counter, int dim) {
  long long n = (long long)(*counter) * dim / 4;
  float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(g)[i] = z;
}

// Sparse-gradient scatter-add of d_hist[p, 0:Di | Di:D] into compact rows gi[slot_i[ih[p]]] /
// gc[slot_c[ch[p]]].  d_hist is streamed once with 16-byte loads; each vector becomes one
// red.global.add.v4.f32.  Padded positions (id 0: a large share of every batch and all aimed at
// one row) are first summed per CTA in shared memory and flushed once.  sumsq[0/1] receive the
// squared norm of the un-deduplicated slice values (tf.clip_by_norm on IndexedSlices).
__global__ void __launch_bounds__(256)
scatter_hist_kernel(const float* __restrict__ dX, const int32_t* __restrict__ ih,
                    const int32_t* __restrict__ ch, int seq_stride, int T,
                    const int32_t* __restrict__ slot_i, const int32_t* __restrict__ slot_c,
                    float* __restrict__ gi, float* __restrict__ gc, int Di, int Dc, long long npos,
                    double* __restrict__ sumsq, int hot) {
  // hot: the lowest item ids (popularity ranks; id 0 = padding) and category 0 are summed per CTA in shared memory
  // and leave the SM once per CTA -- thousands of reductions onto one row would serialise in L2
  extern __shared__ float pad_acc[];  // [hot][Di] + [Dc]
  __shared__ float red[2][8];
  const int VI = Di >> 2, V = (Di + Dc) >> 2;
  const int nacc = hot * Di + Dc;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) pad_acc[i] = 0.f;
  __syncthreads();
  const long long nvec = npos * V;
  float ssi = 0.f, ssc = 0.f;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < nvec;
       g += (long long)gridDim.x * blockDim.x) {
    long long p = g / V;
    int q = (int)(g - p * V);
    long long s = p / T;
    long long io = s * seq_stride + (p - s * T);
    float4 v = ldg_stream(reinterpret_cast<const float4*>(dX) + g);
    float sq = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    if (q < VI) {
      const int id = __ldg(ih + io);
      ssi += sq;
      if (id < hot) {
        float* a = pad_acc + id * Di + q * 4;
        atomicAdd(a + 0, v.x); atomicAdd(a + 1, v.y); atomicAdd(a + 2, v.z); atomicAdd(a + 3, v.w);
      } else {
        red_add_v4(gi + (size_t)slot_i[id] * Di + q * 4, v);
      }
    } else {
      const int id = __ldg(ch + io);
      ssc += sq;
      if (id == 0) {
        float* a = pad_acc + hot * Di + (q - VI) * 4;
        atomicAdd(a + 0, v.x); atomicAdd(a + 1, v.y); atomicAdd(a + 2, v.z); atomicAdd(a + 3, v.w);
      } else {
        red_add_v4(gc + (size_t)slot_c[id] * Dc + (q - VI) * 4, v);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x * 4; i < nacc; i += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(pad_acc + i);
    if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
    if (i < hot * Di) {
      const int id = i / Di, c = i - id * Di;
      const int sl = slot_i[id];
      if (sl >= 0) red_add_v4(gi + (size_t)sl * Di + c, v);
    } else {
      const int sl = slot_c[0];
      if (sl >= 0) red_add_v4(gc + (size_t)sl * Dc + (i - hot * Di), v);
    }
  }
  ssi = warp_sum(ssi); ssc = warp_sum(ssc);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = ssi; red[1][w] = ssc; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[threadIdx.x][i];
    atomicAdd(sumsq + threadIdx.x, (double)t);
  }
}

// Scatter-add of dense per-row gradients d[r, col0:col0+dim] into compact rows (targets, users).
__global__ void scatter_rows_kernel(const float* __restrict__ d, int ldd, int col0, int dim,
                                    const int32_t* __restrict__ idx, int idx_stride,
                                    const int32_t* __restrict__ slot, float* __restrict__ g, int rows,
                                    double* __restrict__ sumsq) {
  const int V = dim >> 2;
  long long n = (long long)rows * V;
  float ss = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int r = (int)(i / V), q = (int)(i % V);
    float4 v = *reinterpret_cast<const float4*>(d + (size_t)r * ldd + col0 + q * 4);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    int id = __ldg(idx + (size_t)r * idx_stride);
    red_add_v4(g + (size_t)slot[id] * dim + q * 4, v);
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0 && ss != 0.f) atomicAdd(sumsq, (double)ss);
}

// Involved-row terms, one warp per unique id: the slice TF forms for the `involved_*` lookup
// (embed_l2 * row, plus the discrepancy gradient for the two user tables) is added to the compact
// gradient row; its squared norm joins sumsq; acc[0] += sum(row^2) (regular loss),
// acc[1] += sum((long-short)^2) (discrepancy loss; only when `other` is given and count_disc != 0).
__global__ void involved_kernel(const float* __restrict__ tab, const float* __restrict__ other, int dim,
                                const int32_t* __restrict__ uniq, const int32_t* __restrict__ counter,
                                float* __restrict__ g, float l2, float disc_w, int count_disc,
                                double* __restrict__ sumsq, double* __restrict__ acc) {
  const int n = *counter;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  float ss = 0.f, rr = 0.f, dd = 0.f;
  // d/d(this) of -w*mean((this-other)^2) = -2w/(n*dim)*(this-other), the same form for both tables.
  const float dcoef = other ? -2.0f * disc_w / ((float)n * (float)dim) : 0.f;
  for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < n; c += gridDim.x * wpb) {
    size_t row = (size_t)uniq[c] * dim;
    for (int k = lane; k < dim; k += 32) {
      float x = tab[row + k];
      float v = l2 * x;
      rr += x * x;
      if (other) {
        float df = x - other[row + k];
        v += dcoef * df;
        dd += df * df;
      }
      g[(size_t)c * dim + k] += v;
      ss += v * v;
    }
  }
  ss = warp_sum(ss); rr = warp_sum(rr); dd = warp_sum(dd);
  if (lane == 0) {
    if (ss != 0.f) atomicAdd(sumsq, (double)ss);
    if (rr != 0.f) atomicAdd(acc, (double)rr);
    if (other && count_disc && dd != 0.f) atomicAdd(acc + 1, (double)dd);
  }
}

struct AdamHyper {
  float lr_t, beta1, beta2, eps, clip;  // clip <= 0: no clipping
  const float* lr_dev;                  // when set, the step size is read from device memory (a replayed CUDA graph
                                        // must not bake in a value that changes every step)
};
CLSR_DEVINL float adam_lr(const AdamHyper& hp) { return hp.lr_dev ? __ldg(hp.lr_dev) : hp.lr_t; }
// one-thread kernel: per-step scalars travel in its launch parameters, outside the captured step
__global__ void set_scalar_kernel(float* dst, float v) { *dst = v; }

// TF's non-lazy sparse Adam: m and v decay and the variable moves on EVERY row;
// rows present in the batch add their (clipped) compact gradient.  Pure streaming over
// var / m / v (3 reads + 3 writes of the whole table) plus the 4-byte slot lookup per row.
template <int UN>
__global__ void __launch_bounds__(256)
adam_sweep_kernel(float* __restrict__ var, float* __restrict__ m, float* __restrict__ v,
                  const int32_t* __restrict__ slot, const float* __restrict__ g, int dim, long long rows,
                  AdamHyper hp, const double* __restrict__ sumsq) {
  const float lr_t = adam_lr(hp);
  const int V = dim >> 2;
  float scale = 1.f;
  if (hp.clip > 0.f) {
    float nrm = (float)sqrt(*sumsq);
    scale = hp.clip / fmaxf(nrm, hp.clip);
  }
  const long long nvec = rows * V;
  // each CTA walks contiguous slabs of UN*blockDim vectors: UN x 3 independent 16-byte loads in flight (UN = 2 with
  // 16 CTAs per SM measured best on B200: 0.85 ms for the four tables, 0.89 of the HBM peak)
  // per thread, every warp access still one fully coalesced 512-byte piece
  const long long stride = blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x * UN + threadIdx.x; i0 < nvec;
       i0 += (long long)gridDim.x * blockDim.x * UN) {
    float4 mv[UN], vv[UN], xv[UN];
    int c[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const long long i = i0 + u * stride;
      // loads are issued from a clamped index (no predicated loads); the stores are predicated
      const long long ic = i < nvec ? i : nvec - 1;
      mv[u] = ldg_stream(reinterpret_cast<const float4*>(m) + ic);
      vv[u] = ldg_stream(reinterpret_cast<const float4*>(v) + ic);
      xv[u] = ldg_stream(reinterpret_cast<const float4*>(var) + ic);
      c[u] = __ldg(slot + ic / V);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const long long i = i0 + u * stride;
      if (i >= nvec) continue;
      float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c[u] >= 0) {
        const int q = (int)(i - (i / V) * V);
        gv = *reinterpret_cast<const float4*>(g + (size_t)c[u] * dim + q * 4);
        gv.x *= scale; gv.y *= scale; gv.z *= scale; gv.w *= scale;
      }
#define CLSR_ADAM1(c_)                                                      \
  mv[u].c_ = hp.beta1 * mv[u].c_ + (1.f - hp.beta1) * gv.c_;                \
  vv[u].c_ = hp.beta2 * vv[u].c_ + (1.f - hp.beta2) * gv.c_ * gv.c_;        \
  xv[u].c_ -= lr_t * mv[u].c_ / (sqrtf(vv[u].c_) + hp.eps);
      CLSR_ADAM1(x) CLSR_ADAM1(y) CLSR_ADAM1(z) CLSR_ADAM1(w)
#undef CLSR_ADAM1
      stg_stream(reinterpret_cast<float4*>(m) + i, mv[u]);
      stg_stream(reinterpret_cast<float4*>(v) + i, vv[u]);
      stg_stream(reinterpret_cast<float4*>(var) + i, xv[u]);
    }
  }
}

// LazyAdam (tf.contrib.opt.LazyAdamOptimizer, base_model.py:275-276): only rows in the batch move.
__global__ void adam_lazy_kernel(float* __restrict__ var, float* __restrict__ m, float* __restrict__ v,
                                 const int32_t* __restrict__ uniq, const int32_t* __restrict__ counter,
                                 const float* __restrict__ g, int dim, AdamHyper hp,
                                 const double* __restrict__ sumsq) {
  const float lr_t = adam_lr(hp);
  const int V = dim >> 2;
  float scale = 1.f;
  if (hp.clip > 0.f) {
    float nrm = (float)sqrt(*sumsq);
    scale = hp.clip / fmaxf(nrm, hp.clip);
  }
  const long long nvec = (long long)(*counter) * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    long long c = i / V;
    int q = (int)(i - c * V);
    size_t off = (size_t)uniq[c] * dim + q * 4;
    float4 gv = *reinterpret_cast<const float4*>(g + (size_t)c * dim + q * 4);
    gv.x *= scale; gv.y *= scale; gv.z *= scale; gv.w *= scale;
    float4 mv = *reinterpret_cast<const float4*>(m + off);
    float4 vv = *reinterpret_cast<const float4*>(v + off);
    float4 xv = *reinterpret_cast<const float4*>(var + off);
#define CLSR_ADAM1(c_)                                          \
  mv.c_ = hp.beta1 * mv.c_ + (1.f - hp.beta1) * gv.c_;          \
  vv.c_ = hp.beta2 * vv.c_ + (1.f - hp.beta2) * gv.c_ * gv.c_;  \
  xv.c_ -= lr_t * mv.c_ / (sqrtf(vv.c_) + hp.eps);
    CLSR_ADAM1(x) CLSR_ADAM1(y) CLSR_ADAM1(z) CLSR_ADAM1(w)
    *reinterpret_cast<float4*>(m + off) = mv;
    *reinterpret_cast<float4*>(v + off) = vv;
    *reinterpret_cast<float4*>(var + off) = xv;
  }
}
#undef CLSR_ADAM1

// Zero up to kZeroSegs buffers in one launch (byte counts are multiples of 4; 16-byte stores where the segment allows).
constexpr int kZeroSegs = 24;
struct ZeroMulti { void* p[kZeroSegs]; long long bytes[kZeroSegs]; int n; };
__global__ void zero_multi_kernel(const __grid_constant__ ZeroMulti a) {
  for (int k = 0; k < a.n; ++k) {
    char* p = static_cast<char*>(a.p[k]);
    const long long nb = a.bytes[k];
    if (((reinterpret_cast<uintptr_t>(p) | (uintptr_t)nb) & 15) == 0) {
      uint4* q = reinterpret_cast<uint4*>(p);
      for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nb / 16; i += (long long)gridDim.x * blockDim.x)
        q[i] = make_uint4(0u, 0u, 0u, 0u);
    } else {
      uint32_t* q = reinterpret_cast<uint32_t*>(p);
      for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nb / 4; i += (long long)gridDim.x * blockDim.x)
        q[i] = 0u;
    }
  }
}

// ---- several small per-table launches folded into one (the step is a long chain of short kernels: every launch
// saved is a few microseconds of tail + ramp) ---------------------------------------------------------------------
constexpr int kMaxSeg = 5;
struct GatherRowsSeg { const int32_t* idx; int idx_stride; TabView tab; int dim; float* out; int ldo, col0, rows; };
struct GatherRowsMulti { GatherRowsSeg s[4]; int n; };
__global__ void gather_rows_multi_kernel(const __grid_constant__ GatherRowsMulti a) {
  for (int k = 0; k < a.n; ++k) {
    const GatherRowsSeg& g = a.s[k];
    const int V = g.dim >> 2;
    const long long n = (long long)g.rows * V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const int r = (int)(i / V), q = (int)(i % V);
      const int id = __ldg(g.idx + (size_t)r * g.idx_stride);
      const float4 v = __ldg(reinterpret_cast<const float4*>(g.tab.row(id, g.dim)) + q);
      *reinterpret_cast<float4*>(g.out + (size_t)r * g.ldo + g.col0 + q * 4) = v;
    }
  }
}

struct UniqueSeg { const int32_t* ids; long long n; int T, seq_stride; int32_t* slot; int32_t* uniq; int32_t* counter; };
struct UniqueMulti { UniqueSeg s[kMaxSeg]; int n; int blk0[kMaxSeg + 1]; };
// One 256-id chunk of ONE segment per CTA (blk0: first CTA of every segment), all segments side by side: the kernel is a
// chain of dependent memory operations (id -> slot[id] -> CAS -> append), so it costs one such chain instead of one per
// segment.  Appends are aggregated per CTA: one atomic on the (single-address) counter per CTA, not per warp.
constexpr int kUniqueBlock = 256;
inline void unique_plan(UniqueMulti* a) {
  a->blk0[0] = 0;
  for (int k = 0; k < a->n; ++k) a->blk0[k + 1] = a->blk0[k] + (int)((a->s[k].n + kUniqueBlock - 1) / kUniqueBlock);
}
__global__ void __launch_bounds__(kUniqueBlock) mark_unique_multi_kernel(const __grid_constant__ UniqueMulti a) {
  __shared__ int s_cnt, s_base;
  int k = 0;
  while (k + 1 < a.n && (int)blockIdx.x >= a.blk0[k + 1]) ++k;
  const UniqueSeg& u = a.s[k];
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const long long i = (long long)((int)blockIdx.x - a.blk0[k]) * kUniqueBlock + threadIdx.x;
  bool claim = false;
  int id = 0;
  const bool valid = i < u.n;
  if (valid) {
    const long long sq = i / u.T;
    id = __ldg(u.ids + sq * u.seq_stride + (i - sq * u.T));
  }
  // one lane per distinct id of the warp goes to the slot table: when the kernel starts, every warp finds the hot ids
  // (the padding row above all) still unclaimed, and a compare-and-swap per lane on one address serialises for tens of
  // microseconds
  const unsigned peers = __match_any_sync(0xffffffffu, valid ? id : -1 - lane);
  if (valid && lane == __ffs(peers) - 1 && u.slot[id] == -1) claim = atomicCAS(&u.slot[id], -1, -2) == -1;
  const unsigned m = __ballot_sync(0xffffffffu, claim);
  int wbase = 0;
  if (m) {
    const int leader = __ffs(m) - 1;
    if (lane == leader) wbase = atomicAdd(&s_cnt, __popc(m));
    wbase = __shfl_sync(0xffffffffu, wbase, leader);
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_cnt > 0) s_base = atomicAdd(u.counter, s_cnt);
  __syncthreads();
  if (claim) {
    const int c = s_base + wbase + __popc(m & ((1u << lane) - 1u));
    u.uniq[c] = id;
    u.slot[id] = c;
  }
}

struct CompactSeg { float* g; const int32_t* counter; int dim; const int32_t* uniq; int32_t* slot; };
struct CompactMulti { CompactSeg s[4]; int n; };
__global__ void zero_compact_multi_kernel(const __grid_constant__ CompactMulti a) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < a.n; ++k) {
    const long long n = (long long)(*a.s[k].counter) * a.s[k].dim / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      reinterpret_cast<float4*>(a.s[k].g)[i] = z;
  }
}
__global__ void reset_slots_multi_kernel(const __grid_constant__ CompactMulti a) {
  for (int k = 0; k < a.n; ++k) {
    const int n = *a.s[k].counter;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) a.s[k].slot[a.s[k].uniq[c]] = -1;
  }
}

struct ScatterRowsSeg { const float* d; int ldd, col0, dim; const int32_t* idx; int idx_stride; const int32_t* slot; float* g; int rows; double* sumsq; };
struct ScatterRowsMulti { ScatterRowsSeg s[4]; int n; };
__global__ void scatter_rows_multi_kernel(const __grid_constant__ ScatterRowsMulti a) {
  for (int k = 0; k < a.n; ++k) {
    const ScatterRowsSeg& r = a.s[k];
    const int V = r.dim >> 2;
    const long long n = (long long)r.rows * V;
    float ss = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const int row = (int)(i / V), q = (int)(i % V);
      const float4 v = *reinterpret_cast<const float4*>(r.d + (size_t)row * r.ldd + r.col0 + q * 4);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      const int id = __ldg(r.idx + (size_t)row * r.idx_stride);
      red_add_v4(r.g + (size_t)r.slot[id] * r.dim + q * 4, v);
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0 && ss != 0.f) atomicAdd(r.sumsq, (double)ss);
  }
}

struct InvolvedSeg { const float* tab; const float* other; int dim; const int32_t* uniq; const int32_t* counter; float* g;
                     float l2, disc_w; int count_disc; double* sumsq; double* acc; };
struct InvolvedMulti { InvolvedSeg s[4]; int n; };
__global__ void involved_multi_kernel(const __grid_constant__ InvolvedMulti a) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int k = 0; k < a.n; ++k) {
    const InvolvedSeg& v = a.s[k];
    const int n = *v.counter;
    float ss = 0.f, rr = 0.f, dd = 0.f;
    const float dcoef = v.other ? -2.0f * v.disc_w / ((float)n * (float)v.dim) : 0.f;
    for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < n; c += gridDim.x * wpb) {
      const size_t row = (size_t)v.uniq[c] * v.dim;
      for (int j = lane; j < v.dim; j += 32) {
        const float x = v.tab[row + j];
        float t = v.l2 * x;
        rr += x * x;
        if (v.other) {
          const float df = x - v.other[row + j];
          t += dcoef * df;
          dd += df * df;
        }
        v.g[(size_t)c * v.dim + j] += t;
        ss += t * t;
      }
    }
    ss = warp_sum(ss); rr = warp_sum(rr); dd = warp_sum(dd);
    if (lane == 0) {
      if (ss != 0.f) atomicAdd(v.sumsq, (double)ss);
      if (rr != 0.f) atomicAdd(v.acc, (double)rr);
      if (v.other && v.count_disc && dd != 0.f) atomicAdd(v.acc + 1, (double)dd);
    }
  }
}

// TF's non-lazy Adam over all four tables in one launch (same inner loop as adam_sweep_kernel).
struct SweepSeg { float* var; float* m; float* v; const int32_t* slot; const float* g; int dim; long long rows; const double* sumsq; };
struct SweepMulti { SweepSeg s[4]; int n; };
template <int UN>
__global__ void __launch_bounds__(256)
adam_sweep_multi_kernel(const __grid_constant__ SweepMulti a, AdamHyper hp) {
  const float lr_t = adam_lr(hp);
  for (int k = 0; k < a.n; ++k) {
    const SweepSeg& w = a.s[k];
    const int V = w.dim >> 2;
    float scale = 1.f;
    if (hp.clip > 0.f) {
      const float nrm = (float)sqrt(*w.sumsq);
      scale = hp.clip / fmaxf(nrm, hp.clip);
    }
    const long long nvec = w.rows * V;
    const long long stride = blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x * UN + threadIdx.x; i0 < nvec;
         i0 += (long long)gridDim.x * blockDim.x * UN) {
      float4 mv[UN], vv[UN], xv[UN];
      int c[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const long long i = i0 + u * stride;
        const long long ic = i < nvec ? i : nvec - 1;
        mv[u] = ldg_stream(reinterpret_cast<const float4*>(w.m) + ic);
        vv[u] = ldg_stream(reinterpret_cast<const float4*>(w.v) + ic);
        xv[u] = ldg_stream(reinterpret_cast<const float4*>(w.var) + ic);
        c[u] = __ldg(w.slot + ic / V);
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const long long i = i0 + u * stride;
        if (i >= nvec) continue;
        float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c[u] >= 0) {
          const int q = (int)(i - (i / V) * V);
          gv = *reinterpret_cast<const float4*>(w.g + (size_t)c[u] * w.dim + q * 4);
          gv.x *= scale; gv.y *= scale; gv.z *= scale; gv.w *= scale;
        }
#define CLSR_ADAM1(c_)                                                      \
  mv[u].c_ = hp.beta1 * mv[u].c_ + (1.f - hp.beta1) * gv.c_;                \
  vv[u].c_ = hp.beta2 * vv[u].c_ + (1.f - hp.beta2) * gv.c_ * gv.c_;        \
  xv[u].c_ -= lr_t * mv[u].c_ / (sqrtf(vv[u].c_) + hp.eps);
        CLSR_ADAM1(x) CLSR_ADAM1(y) CLSR_ADAM1(z) CLSR_ADAM1(w)
#undef CLSR_ADAM1
        stg_stream(reinterpret_cast<float4*>(w.m) + i, mv[u]);
        stg_stream(reinterpret_cast<float4*>(w.v) + i, vv[u]);
        stg_stream(reinterpret_cast<float4*>(w.var) + i, xv[u]);
      }
    }
  }
}

// ---- row-sharded tables under the data-parallel step ----------------------------------------------------
// Every rank de-duplicates its own slices exactly as on one GPU (slot table, compact rows), then pushes each
// unique compact row ONCE to the rank that owns it: one 16-byte reduction per vector straight into the
// owner's dense gradient shard over NVLink, plus a `touched` mark.  The owner adds the involved-row terms
// and applies the optimizer to its 1/world of the rows; nothing is all-gathered, no replica repeats
// another's scatter, and the TF-faithful full-table sweep is divided by world.
struct GradView {
  float* g[kMaxWorld];          // rank r's dense gradient shard [local_rows, dim]
  int32_t* touched[kMaxWorld];  // rank r's per-row "present in this step" marks
  int shift, mask;
};

__global__ void push_compact_kernel(const int32_t* __restrict__ uniq, const int32_t* __restrict__ counter,
                                    const float* __restrict__ cg, int dim, GradView gv) {
  const int V = dim >> 2;
  const long long n = (long long)(*counter) * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long c = i / V;
    const int q = (int)(i - c * V);
    const int id = uniq[c];
    const int owner = id & gv.mask;
    const size_t local = (size_t)(id >> gv.shift);
    const float4 v = *reinterpret_cast<const float4*>(cg + (size_t)c * dim + q * 4);
    red_add_v4(gv.g[owner] + local * dim + q * 4, v);
    if (q == 0) gv.touched[owner][local] = 1;
  }
}

// out[0] = number of touched local rows (the owner's share of tf.unique's count)
__global__ void shard_count_touched_kernel(const int32_t* __restrict__ touched, long long rows, int32_t* __restrict__ out) {
  int n = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x)
    n += touched[i] != 0;
  n = __reduce_add_sync(0xffffffffu, n);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(out, n);
}

// involved_kernel on the owner's shard: one warp per touched local row (n_uniq: the GLOBAL unique count, for
// the discrepancy mean).
__global__ void shard_involved_kernel(const float* __restrict__ tab, const float* __restrict__ other, int dim,
                                      const int32_t* __restrict__ touched, long long rows, float* __restrict__ g,
                                      float l2, float disc_w, const int32_t* __restrict__ n_uniq, int count_disc,
                                      double* __restrict__ sumsq, double* __restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  float ss = 0.f, rr = 0.f, dd = 0.f;
  const float dcoef = other ? -2.0f * disc_w / ((float)(*n_uniq) * (float)dim) : 0.f;
  // a warp scans 32 consecutive marks with one coalesced load and then walks the touched rows among them
  for (long long r0 = ((long long)blockIdx.x * wpb + (threadIdx.x >> 5)) * 32; r0 < rows; r0 += (long long)gridDim.x * wpb * 32) {
    const long long rl = r0 + lane;
    unsigned m = __ballot_sync(0xffffffffu, rl < rows && touched[rl] != 0);
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      const size_t row = (size_t)(r0 + b) * dim;
      for (int k = lane; k < dim; k += 32) {
        float x = tab[row + k];
        float v = l2 * x;
        rr += x * x;
        if (other) {
          float df = x - other[row + k];
          v += dcoef * df;
          dd += df * df;
        }
        g[row + k] += v;
        ss += v * v;
      }
    }
  }
  ss = warp_sum(ss); rr = warp_sum(rr); dd = warp_sum(dd);
  if (lane == 0) {
    if (ss != 0.f) atomicAdd(sumsq, (double)ss);
    if (rr != 0.f) atomicAdd(acc, (double)rr);
    if (other && count_disc && dd != 0.f) atomicAdd(acc + 1, (double)dd);
  }
}

// TF's non-lazy Adam over the owner's shard with a dense gradient shard: every row decays and moves, touched
// rows add their clipped gradient; the gradient rows are zeroed again on the way out.
template <int UN>
__global__ void __launch_bounds__(256)
adam_sweep_shard_kernel(float* __restrict__ var, float* __restrict__ m, float* __restrict__ v, float* __restrict__ g,
                        const int32_t* __restrict__ touched, int dim, long long rows, AdamHyper hp,
                        const double* __restrict__ sumsq) {
  const float lr_t = adam_lr(hp);
  const int V = dim >> 2;
  float scale = 1.f;
  if (hp.clip > 0.f) {
    float nrm = (float)sqrt(*sumsq);
    scale = hp.clip / fmaxf(nrm, hp.clip);
  }
  const long long nvec = rows * V;
  const long long stride = blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x * UN + threadIdx.x; i0 < nvec;
       i0 += (long long)gridDim.x * blockDim.x * UN) {
    float4 mv[UN], vv[UN], xv[UN];
    int c[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const long long i = i0 + u * stride;
      const long long ic = i < nvec ? i : nvec - 1;
      mv[u] = ldg_stream(reinterpret_cast<const float4*>(m) + ic);
      vv[u] = ldg_stream(reinterpret_cast<const float4*>(v) + ic);
      xv[u] = ldg_stream(reinterpret_cast<const float4*>(var) + ic);
      c[u] = __ldg(touched + ic / V);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const long long i = i0 + u * stride;
      if (i >= nvec) continue;
      float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c[u]) {
        gv = *reinterpret_cast<const float4*>(g + i * 4);
        *reinterpret_cast<float4*>(g + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        gv.x *= scale; gv.y *= scale; gv.z *= scale; gv.w *= scale;
      }
#define CLSR_ADAM1(c_)                                                      \
  mv[u].c_ = hp.beta1 * mv[u].c_ + (1.f - hp.beta1) * gv.c_;                \
  vv[u].c_ = hp.beta2 * vv[u].c_ + (1.f - hp.beta2) * gv.c_ * gv.c_;        \
  xv[u].c_ -= lr_t * mv[u].c_ / (sqrtf(vv[u].c_) + hp.eps);
      CLSR_ADAM1(x) CLSR_ADAM1(y) CLSR_ADAM1(z) CLSR_ADAM1(w)
#undef CLSR_ADAM1
      stg_stream(reinterpret_cast<float4*>(m) + i, mv[u]);
      stg_stream(reinterpret_cast<float4*>(v) + i, vv[u]);
      stg_stream(reinterpret_cast<float4*>(var) + i, xv[u]);
    }
  }
}

// LazyAdam on the owner's shard (only touched rows move), or -- with hp.lr_t == 0 and `apply` = 0 -- just the
// clean-up of a gradient-only step: one warp per touched row, gradient row zeroed on the way out.
__global__ void adam_lazy_shard_kernel(float* __restrict__ var, float* __restrict__ m, float* __restrict__ v,
                                       float* __restrict__ g, const int32_t* __restrict__ touched, int dim,
                                       long long rows, AdamHyper hp, const double* __restrict__ sumsq, int apply) {
  const float lr_t = adam_lr(hp);
  float scale = 1.f;
  if (hp.clip > 0.f) {
    float nrm = (float)sqrt(*sumsq);
    scale = hp.clip / fmaxf(nrm, hp.clip);
  }
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long r0 = ((long long)blockIdx.x * wpb + (threadIdx.x >> 5)) * 32; r0 < rows; r0 += (long long)gridDim.x * wpb * 32) {
    const long long rl = r0 + lane;
    unsigned msk = __ballot_sync(0xffffffffu, rl < rows && touched[rl] != 0);
    while (msk) {
      const int b = __ffs(msk) - 1;
      msk &= msk - 1;
      const size_t row = (size_t)(r0 + b) * dim;
      for (int k = lane; k < dim; k += 32) {
        const float gv = g[row + k] * scale;
        g[row + k] = 0.f;
        if (apply) {
          const float mv = hp.beta1 * m[row + k] + (1.f - hp.beta1) * gv;
          const float vv = hp.beta2 * v[row + k] + (1.f - hp.beta2) * gv * gv;
          m[row + k] = mv; v[row + k] = vv;
          var[row + k] -= lr_t * mv / (sqrtf(vv) + hp.eps);
        }
      }
    }
  }
}

// ---- one-shot all-reduce over peer memory ------------------------------------------------------------------
// The step's small reductions (BatchNorm sums: 2N doubles per layer and direction; loss partial sums; counts)
// are latency, not bandwidth: an NCCL call costs 10-30 us each and a data-parallel step needs ~20 of them.  Here
// every rank PUSHES its vector into its slot of every peer's buffer (NVLink stores), publishes a sequence flag,
// waits until all peers' flags for this sequence number have arrived in its own memory, and sums the slots in
// rank order -- every rank gets the bit-identical result after one NVLink round trip, inside the kernel that
// consumes it (the BatchNorm finalize kernels below call it in place).  n = 0 is a barrier.  Two phases
// (seq & 1) suffice: a peer can only write phase p again after it has passed the next call, which needs this
// rank's flag, which this rank only publishes after it has read phase p.
constexpr int kPeerSlots = 512;
struct PeerComm {
  double* data[kMaxWorld];                // rank r's buffer: [2][world][kPeerSlots]
  unsigned long long* flag[kMaxWorld];    // rank r's flags:  [2][world]
  int rank, world;
  unsigned long long* seq_ctr;            // this rank's exchange counter, in device memory: every rank issues the
                                          // same sequence of exchanges, so the counters agree without the host
                                          // passing a number per launch (which a replayed CUDA graph would freeze)
};
CLSR_DEVINL void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
CLSR_DEVINL unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
CLSR_DEVINL double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
CLSR_DEVINL void st_volatile_f64(double* p, double v) {
  asm volatile("st.volatile.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// All threads of ONE CTA call this; vals: n doubles in shared memory, replaced by their sum over ranks.
CLSR_DEVINL void peer_allreduce_block(const PeerComm& pc, unsigned long long seq, double* vals, int n) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (pc.seq_ctr) {
    __shared__ unsigned long long s_seq;
    if (tid == 0) { s_seq = *pc.seq_ctr + 1ull; *pc.seq_ctr = s_seq; }
    __syncthreads();
    seq = s_seq;
  }
  const int ph = (int)(seq & 1ull);
  for (int r = 0; r < pc.world; ++r) {
    double* dst = pc.data[r] + ((size_t)ph * pc.world + pc.rank) * kPeerSlots;
    for (int i = tid; i < n; i += nt) st_volatile_f64(dst + i, vals[i]);
  }
  __threadfence_system();
  __syncthreads();
  if (tid < pc.world) {
    st_release_sys(pc.flag[tid] + (size_t)ph * pc.world + pc.rank, seq);
    const unsigned long long* mine = pc.flag[pc.rank] + (size_t)ph * pc.world + tid;
    while (ld_acquire_sys(mine) < seq) __nanosleep(20);
  }
  __syncthreads();
  const double* src = pc.data[pc.rank] + (size_t)ph * pc.world * kPeerSlots;
  for (int i = tid; i < n; i += nt) {
    double s = 0.0;
    for (int r = 0; r < pc.world; ++r) s += ld_volatile_f64(src + (size_t)r * kPeerSlots + i);
    vals[i] = s;
  }
  __syncthreads();
}

// Standalone form for the step's scalars: up to three arrays (doubles, doubles, int32) reduced in one exchange.
__global__ void peer_allreduce_kernel(PeerComm pc, unsigned long long seq, double* a, int na, double* b, int nb,
                                      int32_t* c, int nc) {
  __shared__ double vals[kPeerSlots];
  const int tid = threadIdx.x;
  for (int i = tid; i < na; i += blockDim.x) vals[i] = a[i];
  for (int i = tid; i < nb; i += blockDim.x) vals[na + i] = b[i];
  for (int i = tid; i < nc; i += blockDim.x) vals[na + nb + i] = (double)c[i];
  __syncthreads();
  peer_allreduce_block(pc, seq, vals, na + nb + nc);
  for (int i = tid; i < na; i += blockDim.x) a[i] = vals[i];
  for (int i = tid; i < nb; i += blockDim.x) b[i] = vals[na + i];
  for (int i = tid; i < nc; i += blockDim.x) c[i] = (int32_t)llrint(vals[na + nb + i]);
}

}  // namespace clsr
