// Evaluation metrics on the GPU.
//
// Reference: run_eval / run_weighted_eval (sequential_base_model.py:204-292) collect every prediction of the
// validation / test file in Python lists and hand them to sklearn / pandas loops (deeprec_utils.py:554-810:
// cal_metric -> auc, logloss, mean_mrr, ndcg@k, hit@k, group_auc; cal_weighted_metric -> wauc, the GAUC of the
// README).  Here the predictions stay on the device (clsr_predict_device) and one call computes all of them:
//
//   * auc / wauc: rows are ordered by a 64-bit key (user | order-preserving score bits | label) with an LSD
//     radix sort (8-bit digits, passes whose digit is constant are skipped); every row finds the run of equal
//     scores it lies in by binary search, which gives its tie-averaged rank, and the Mann-Whitney statistic
//     (sum of the positives' ranks) is accumulated per user segment -- exactly sklearn.roc_auc_score's value,
//     ties counting one half.  auc is the same computation with the user bits cleared.
//   * per-impression metrics (groups of 1 + num_ngs consecutive rows): one warp per group, ranks by pairwise
//     counting.  Equal scores are ordered as np.argsort(score)[::-1] orders them on a stable sort (the later
//     row first).
//   * logloss: a reduction in double.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/clsr_b200.h"
#include "common.cuh"

using namespace clsr;

namespace {

typedef unsigned long long u64;

__device__ __forceinline__ uint32_t order_bits(float s) {
  uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void build_keys_kernel(const float* __restrict__ pred, const float* __restrict__ label,
                                  const int32_t* __restrict__ users, long long n, int with_user, u64* __restrict__ keys) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const u64 u = with_user ? (u64)(uint32_t)users[i] : 0ull;
    keys[i] = (u << 33) | ((u64)order_bits(pred[i]) << 1) | (label[i] == 1.0f ? 1ull : 0ull);
  }
}

// ---- LSD radix sort of 64-bit keys: units of work are warps, each owning a contiguous slice ------------------
constexpr int kWarpsPerBlock = 8;

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
radix_hist_kernel(const u64* __restrict__ keys, long long n, int shift, long long per_warp, int nwarps,
                  uint32_t* __restrict__ hist /* [256][nwarps] */) {
  __shared__ uint32_t h[kWarpsPerBlock][256];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarpsPerBlock + w;
  for (int i = lane; i < 256; i += 32) h[w][i] = 0;
  __syncwarp();
  if (gw < nwarps) {
    const long long a = (long long)gw * per_warp, b = a + per_warp < n ? a + per_warp : n;
    for (long long i = a + lane; i < b; i += 32) atomicAdd(&h[w][(keys[i] >> shift) & 255], 1u);
  }
  __syncwarp();
  if (gw < nwarps)
    for (int i = lane; i < 256; i += 32) hist[(size_t)i * nwarps + gw] = h[w][i];
}

// exclusive scan of hist in (digit-major, warp-minor) order; flag[0] = 1 when one digit holds every key
__global__ void __launch_bounds__(1024)
radix_scan_kernel(uint32_t* __restrict__ hist, long long total, long long n, int nwarps, int* __restrict__ constant_digit) {
  __shared__ unsigned long long part[1024];
  __shared__ int is_const;
  const int tid = threadIdx.x;
  if (tid == 0) is_const = 0;
  const long long per = (total + 1023) / 1024;
  const long long a = tid * per, b = a + per < total ? a + per : total;
  unsigned long long s = 0;
  for (long long i = a; i < b; ++i) s += hist[i];
  part[tid] = s;
  __syncthreads();
  // digit totals: digit d occupies [d*nwarps, (d+1)*nwarps)
  if (tid < 256) {
    unsigned long long t = 0;
    for (int w = 0; w < nwarps; ++w) t += hist[(size_t)tid * nwarps + w];
    if ((long long)t == n) is_const = 1;
  }
  __syncthreads();
  if (tid == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < 1024; ++i) { const unsigned long long v = part[i]; part[i] = run; run += v; }
    *constant_digit = is_const;
  }
  __syncthreads();
  if (is_const) return;   // the pass is skipped: offsets are not needed
  unsigned long long run = part[tid];
  for (long long i = a; i < b; ++i) { const uint32_t v = hist[i]; hist[i] = (uint32_t)run; run += v; }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
radix_scatter_kernel(const u64* __restrict__ in, u64* __restrict__ out, long long n, int shift, long long per_warp,
                     int nwarps, const uint32_t* __restrict__ offs, const int* __restrict__ constant_digit) {
  __shared__ uint32_t base[kWarpsPerBlock][256];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarpsPerBlock + w;
  if (gw >= nwarps) return;
  const long long a = (long long)gw * per_warp, b = a + per_warp < n ? a + per_warp : n;
  if (*constant_digit) {   // every key has the same digit: the pass is the identity permutation
    for (long long i = a + lane; i < b; i += 32) out[i] = in[i];
    return;
  }
  for (int i = lane; i < 256; i += 32) base[w][i] = offs[(size_t)i * nwarps + gw];
  __syncwarp();
  for (long long i0 = a; i0 < b; i0 += 32) {
    const long long i = i0 + lane;
    const bool ok = i < b;
    const u64 k = ok ? in[i] : 0ull;
    const int d = ok ? (int)((k >> shift) & 255) : 256 + lane;   // inactive lanes match nobody
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t pos = 0;
    if (ok) pos = base[w][d] + rank;
    __syncwarp();
    if (ok && rank == 0) base[w][d] += __popc(peers);   // the first lane of each digit advances its cursor
    __syncwarp();
    if (ok) out[pos] = k;
  }
}

// ---- rank statistics on the sorted keys ---------------------------------------------------------------------
__device__ __forceinline__ long long lower_bound(const u64* __restrict__ k, long long n, u64 v) {
  long long lo = 0, hi = n;
  while (lo < hi) { const long long mid = (lo + hi) >> 1; if (k[mid] < v) lo = mid + 1; else hi = mid; }
  return lo;
}

// For every positive row: tie-averaged rank inside its user segment -> ranksum[segment start]; npos likewise.
__global__ void rank_accumulate_kernel(const u64* __restrict__ keys, long long n, double* __restrict__ ranksum,
                                       unsigned int* __restrict__ npos) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const u64 k = keys[i];
    if (!(k & 1ull)) continue;
    const u64 tie = k & ~1ull;                       // same user, same score, either label
    const long long a = lower_bound(keys, n, tie), b = lower_bound(keys, n, tie + 2ull);
    const long long seg = lower_bound(keys, n, (k >> 33) << 33);
    const double avg_rank = 0.5 * (double)(a + b + 1) - (double)seg;   // 1-based, tie-averaged, segment-local
    atomicAdd(&ranksum[seg], avg_rank);
    atomicAdd(&npos[seg], 1u);
  }
}

// One thread per segment start: AUC of the segment, weighted by its share of the rows.
// out[0] += weight * auc, out[1] += 1 (segments), flag |= 1 when a segment holds one class only.
__global__ void segment_auc_kernel(const u64* __restrict__ keys, long long n, const double* __restrict__ ranksum,
                                   const unsigned int* __restrict__ npos, double* __restrict__ out, int* __restrict__ flag) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const u64 u = keys[i] >> 33;
    if (i > 0 && (keys[i - 1] >> 33) == u) continue;
    const long long end = u >= 0x7fffffffull ? n : lower_bound(keys, n, (u + 1ull) << 33);
    const double len = (double)(end - i), np = (double)npos[i], nn = len - np;
    if (np == 0.0 || nn == 0.0) { atomicOr(flag, 1); continue; }
    const double auc = (ranksum[i] - np * (np + 1.0) * 0.5) / (np * nn);
    atomicAdd(&out[0], auc * (len / (double)n));
    atomicAdd(&out[1], 1.0);
  }
}

__global__ void logloss_kernel(const float* __restrict__ pred, const float* __restrict__ label, long long n,
                               double* __restrict__ out /* [0] sum, [1] positives */) {
  double s = 0.0, np = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double p = (double)pred[i];
    p = p < 10e-12 ? 10e-12 : (p > 1.0 - 10e-12 ? 1.0 - 10e-12 : p);   // deeprec_utils.py:647-650
    const double y = (double)label[i];
    s -= y * log(p) + (1.0 - y) * log(1.0 - p);
    np += label[i] == 1.0f ? 1.0 : 0.0;
  }
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); np += __shfl_xor_sync(0xffffffffu, np, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], s); atomicAdd(&out[1], np); }
}

// Per-impression metrics, one warp per group of `G` consecutive rows (labels in {0, 1}).
// acc: [0] mrr, [1] group_auc, [2 .. 2+nk) ndcg@k, [2+nk .. 2+2nk) hit@k.
constexpr int kMaxK = 8;
struct KList { int k[kMaxK]; int n; };
__global__ void group_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ label, long long ngroups,
                                     int G, KList ks, double* __restrict__ acc, int* __restrict__ flag) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  double mrr = 0.0, gauc = 0.0, ndcg[kMaxK], hit[kMaxK];
  for (int q = 0; q < kMaxK; ++q) { ndcg[q] = 0.0; hit[q] = 0.0; }
  for (long long g = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); g < ngroups; g += (long long)gridDim.x * wpb) {
    const float* p = pred + g * G;
    const float* y = label + g * G;
    double s_mrr = 0.0, s_pos = 0.0, s_u = 0.0, dcg[kMaxK];
    int hitk[kMaxK];
    for (int q = 0; q < kMaxK; ++q) { dcg[q] = 0.0; hitk[q] = 0; }
    for (int i = lane; i < G; i += 32) {
      if (y[i] != 1.0f) continue;
      const float si = p[i];
      int above = 0, lower = 0, equal_neg = 0;
      for (int j = 0; j < G; ++j) {
        const float sj = p[j];
        above += (sj > si) || (sj == si && j > i);                 // descending order, later row first among ties
        if (y[j] != 1.0f) { lower += sj < si; equal_neg += sj == si; }
      }
      const int rank = above + 1;
      s_mrr += 1.0 / rank;
      s_pos += 1.0;
      s_u += (double)lower + 0.5 * (double)equal_neg;              // Mann-Whitney, ties one half
      for (int q = 0; q < ks.n; ++q)
        if (rank <= (ks.k[q] < G ? ks.k[q] : G)) { dcg[q] += 1.0 / log2((double)rank + 1.0); hitk[q] = 1; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      s_mrr += __shfl_xor_sync(0xffffffffu, s_mrr, o); s_pos += __shfl_xor_sync(0xffffffffu, s_pos, o);
      s_u += __shfl_xor_sync(0xffffffffu, s_u, o);
      for (int q = 0; q < ks.n; ++q) { dcg[q] += __shfl_xor_sync(0xffffffffu, dcg[q], o); hitk[q] |= __shfl_xor_sync(0xffffffffu, hitk[q], o); }
    }
    if (lane == 0) {
      const double nneg = (double)G - s_pos;
      if (s_pos == 0.0 || nneg == 0.0) atomicOr(flag, 2);
      else gauc += s_u / (s_pos * nneg);
      if (s_pos > 0.0) mrr += s_mrr / s_pos;
      for (int q = 0; q < ks.n; ++q) {
        const int kk = ks.k[q] < G ? ks.k[q] : G;
        double ideal = 0.0;
        for (int r = 1; r <= kk && r <= (int)s_pos; ++r) ideal += 1.0 / log2((double)r + 1.0);
        if (ideal > 0.0) ndcg[q] += dcg[q] / ideal;
        hit[q] += hitk[q];
      }
    }
  }
  if (lane == 0) {
    atomicAdd(&acc[0], mrr); atomicAdd(&acc[1], gauc);
    for (int q = 0; q < ks.n; ++q) { atomicAdd(&acc[2 + q], ndcg[q]); atomicAdd(&acc[2 + ks.n + q], hit[q]); }
  }
}

int sort_keys(u64* a, u64* b, long long n, uint32_t* hist, int* cflag, int nwarps, long long per_warp, int bits,
              cudaStream_t st, u64** sorted) {
  const int nblocks = (nwarps + kWarpsPerBlock - 1) / kWarpsPerBlock;
  for (int shift = 0; shift < bits; shift += 8) {
    radix_hist_kernel<<<nblocks, kWarpsPerBlock * 32, 0, st>>>(a, n, shift, per_warp, nwarps, hist);
    radix_scan_kernel<<<1, 1024, 0, st>>>(hist, 256LL * nwarps, n, nwarps, cflag);
    radix_scatter_kernel<<<nblocks, kWarpsPerBlock * 32, 0, st>>>(a, b, n, shift, per_warp, nwarps, hist, cflag);
    u64* t = a; a = b; b = t;
  }
  *sorted = a;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace

extern "C" int clsr_eval_metrics_compute(int32_t device, const float* preds, const float* labels, const int32_t* users,
                                         int64_t n, int32_t group, const int32_t* ks, int32_t nk, clsr_eval_metrics* out,
                                         void* stream) {
  if (!preds || !labels || !out || n <= 0 || nk < 0 || nk > kMaxK || group < 0) return CLSR_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return CLSR_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  memset(out, 0, sizeof *out);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int nwarps = sms * 2 * kWarpsPerBlock;
  const long long per_warp = (n + nwarps - 1) / nwarps;
  // workspace: two key buffers, histogram, per-segment accumulators, scalars
  u64 *ka = nullptr, *kb = nullptr;
  uint32_t* hist = nullptr;
  double* ranksum = nullptr;
  unsigned int* npos = nullptr;
  double* acc = nullptr;   // [0..1] auc, [2..3] wauc, [4..5] logloss, [8..) group metrics
  int* flags = nullptr;    // [0] constant digit, [1] auc flag, [2] wauc flag, [3] group flag
  cudaError_t c = cudaMalloc(&ka, (size_t)n * 8);
  if (c == cudaSuccess) c = cudaMalloc(&kb, (size_t)n * 8);
  if (c == cudaSuccess) c = cudaMalloc(&hist, (size_t)256 * nwarps * 4);
  if (c == cudaSuccess) c = cudaMalloc(&ranksum, (size_t)n * 8);
  if (c == cudaSuccess) c = cudaMalloc(&npos, (size_t)n * 4);
  if (c == cudaSuccess) c = cudaMalloc(&acc, 64 * 8);
  if (c == cudaSuccess) c = cudaMalloc(&flags, 8 * 4);
  int rc = CLSR_OK;
  if (c != cudaSuccess) rc = CLSR_ERR_CUDA;
  const int grid = sms * 8;
  if (rc == CLSR_OK) {
    cudaMemsetAsync(acc, 0, 64 * 8, st);
    cudaMemsetAsync(flags, 0, 8 * 4, st);
    logloss_kernel<<<grid, 256, 0, st>>>(preds, labels, n, acc + 4);
    for (int pass = 0; pass < (users ? 2 : 1) && rc == CLSR_OK; ++pass) {   // pass 0: auc (no user bits), pass 1: wauc
      build_keys_kernel<<<grid, 256, 0, st>>>(preds, labels, users, n, pass, ka);
      u64* sorted = nullptr;
      if (sort_keys(ka, kb, n, hist, flags, nwarps, per_warp, pass ? 64 : 40, st, &sorted)) { rc = CLSR_ERR_CUDA; break; }
      cudaMemsetAsync(ranksum, 0, (size_t)n * 8, st);
      cudaMemsetAsync(npos, 0, (size_t)n * 4, st);
      rank_accumulate_kernel<<<grid, 256, 0, st>>>(sorted, n, ranksum, npos);
      segment_auc_kernel<<<grid, 256, 0, st>>>(sorted, n, ranksum, npos, acc + 2 * pass, flags + 1 + pass);
    }
    KList kl;
    kl.n = nk;
    for (int q = 0; q < kMaxK; ++q) kl.k[q] = q < nk ? ks[q] : 1;
    const long long ngroups = group > 0 ? n / group : 0;
    if (group > 0 && n % group == 0)
      group_metrics_kernel<<<grid, 256, 0, st>>>(preds, labels, ngroups, group, kl, acc + 8, flags + 3);
    double h_acc[64];
    int h_flags[8];
    if (cudaMemcpyAsync(h_acc, acc, sizeof h_acc, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(h_flags, flags, sizeof h_flags, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess)
      rc = CLSR_ERR_CUDA;
    if (rc == CLSR_OK) {
      out->n = n;
      out->n_pos = (int64_t)h_acc[5];
      out->auc = h_acc[0];
      out->n_users = (int64_t)h_acc[3];
      out->wauc = h_acc[2];
      out->logloss = h_acc[4] / (double)n;
      out->n_groups = ngroups;
      out->status = (h_flags[1] ? 1 : 0) | (h_flags[2] ? 2 : 0) | (h_flags[3] ? 4 : 0);
      if (ngroups > 0 && n % group == 0) {
        out->mean_mrr = h_acc[8] / (double)ngroups;
        out->group_auc = h_acc[9] / (double)ngroups;
        for (int q = 0; q < nk; ++q) { out->ndcg[q] = h_acc[10 + q] / (double)ngroups; out->hit[q] = h_acc[10 + nk + q] / (double)ngroups; }
      }
    }
  }
  cudaFree(ka); cudaFree(kb); cudaFree(hist); cudaFree(ranksum); cudaFree(npos); cudaFree(acc); cudaFree(flags);
  return rc;
}
