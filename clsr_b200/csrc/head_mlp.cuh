// The two row-wise _fcn_net blocks of the head (alpha gate, logit: base_model.py:627-708 called from
// clsr.py:262-277) as ONE persistent kernel per direction.
//
// A _fcn_net over B rows is three small layers (K -> n0 -> n1 -> 1) separated by training-mode BatchNorm, i.e. by
// two global reductions.  Launched layer by layer that is 11 (forward) + 15 (backward) kernels per step, each of
// them over only B = 20 480 rows: pure launch / prologue / tail cost (0.59 ms of a 5.4 ms step).  Here every CTA
// owns a fixed slice of the rows for the whole network and keeps its slice of h0 / h1 (forward) or dy / dh
// (backward) in shared memory; the BatchNorm statistics are the only thing that crosses CTAs, through double
// atomics and a grid barrier (all CTAs are co-resident: cooperative launch, one CTA per SM).  Data-parallel runs
// exchange the statistics over NVLink peer memory inside the same kernel (CTA 0, peer_allreduce_block).
// fp32 CUDA-core arithmetic throughout (0.7 GFLOP per network: the tensor cores have nothing to win here).
#pragma once
#include "common.cuh"
#include "embed.cuh"

namespace clsr {

constexpr int kCoopThreads = 512;   // 16 warps: the per-thread FMA chains need them to hide shared-memory latency
constexpr int kCoopMaxN = 128;   // widest hidden layer

struct CoopBn {
  const float *gamma, *beta;
  float *mmean, *mvar;                    // moving statistics (dense parameter block)
  float *scale, *shift, *mean, *rstd;     // forward coefficients (also read by the backward kernel)
  float *al, *be, *ga;                    // backward coefficients: dx = al*dy + be*h + ga
  float *dgamma, *dbeta;
  double *stat_f, *stat_b;                // [2N] each, zeroed at the start of the step
};

struct CoopMlp {
  int rows, K, n0, n1, ld_in;
  const float *in, *w0, *b0, *w1, *b1, *wo, *bo;
  const float *w0T, *w1T;                 // transposed copies ([n0][K], [n1][n0]) from the folded-weight block
  float *h0, *h1, *out;
  const float* dout;
  float* din;
  float *dw0, *dw1, *dwo, *dbo;          // (hidden-layer biases sit in front of a BatchNorm: no gradient)
  CoopBn bn0, bn1;
  double count;                           // rows over all ranks
  float eps, momentum;
  int update_moving;
  float add_scale;                        // 1 on rank 0, else 0: gamma / beta gradients come from global sums
  unsigned long long* bar;                // grid-barrier counter (monotonic, shared by all cooperative launches)
  int world;
  PeerComm pc;
};

CLSR_DEVINL unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Every CTA of the grid arrives exactly once per barrier; the counter only grows, the k-th barrier completes at
// k * gridDim.x arrivals (every cooperative kernel of an engine uses the same grid size).
CLSR_DEVINL void grid_barrier(unsigned long long* bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long n = gridDim.x;
    const unsigned long long old = atomicAdd(bar, 1ull);
    const unsigned long long target = (old / n + 1ull) * n;
    while (ld_acquire_gpu_u64(bar) < target) __nanosleep(40);
    __threadfence();
  }
  __syncthreads();
}
// Asynchronous global -> shared copies (LDGSTS): every thread keeps all of its pieces in flight at once; a staging loop
// of plain loads and stores ran one L2 round trip after the other and was most of these kernels' time.
CLSR_DEVINL uint32_t coop_s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
CLSR_DEVINL void cp_async4(float* sdst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(coop_s32(sdst)), "l"(gsrc) : "memory");
}
CLSR_DEVINL void cp_async16(float* sdst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(coop_s32(sdst)), "l"(gsrc) : "memory");
}
CLSR_DEVINL void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
// n contiguous floats
CLSR_DEVINL void coop_copy(float* sdst, const float* gsrc, int n) {
  const bool v = ((reinterpret_cast<uintptr_t>(gsrc) | (uintptr_t)coop_s32(sdst)) & 15) == 0;
  if (v) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) cp_async16(sdst + 4 * i, gsrc + 4 * i);
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) cp_async4(sdst + i, gsrc + i);
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) cp_async4(sdst + i, gsrc + i);
  }
}
// rows x cols floats, row pitches lds (shared) / ldg (global); columns [cols, lds_fill) of every row are zeroed
CLSR_DEVINL void coop_copy_rows(float* sdst, int lds, const float* gsrc, int ldg, int rows, int cols, int lds_fill) {
  const int n = rows * lds_fill;
  int r = threadIdx.x / lds_fill, k = threadIdx.x - r * lds_fill;
  const int dr = blockDim.x / lds_fill, dk = blockDim.x - dr * lds_fill;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (k < cols) cp_async4(sdst + (size_t)r * lds + k, gsrc + (size_t)r * ldg + k);
    else sdst[(size_t)r * lds + k] = 0.f;
    r += dr; k += dk;
    if (k >= lds_fill) { k -= lds_fill; ++r; }
  }
}

CLSR_DEVINL double ldcg_f64(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// Sum of the statistics over the ranks (CTA 0 exchanges, everybody waits).  One extra grid barrier, world > 1 only.
CLSR_DEVINL void coop_exchange_stats(const CoopMlp& a, double* stat, int n2, double* vals /* shared [kPeerSlots] */) {
  if (a.world <= 1) return;
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < n2; i += blockDim.x) vals[i] = ldcg_f64(stat + i);
    __syncthreads();
    peer_allreduce_block(a.pc, 0ull, vals, n2);
    for (int i = threadIdx.x; i < n2; i += blockDim.x) stat[i] = vals[i];
    __threadfence();
  }
  grid_barrier(a.bar);
}

// scale / shift (+ mean, rstd) of a training-mode BatchNorm from the global sums; CTA 0 publishes them and moves the
// moving statistics.  Same arithmetic as bn_fwd_finalize_kernel.
CLSR_DEVINL void coop_bn_fwd(const CoopMlp& a, const CoopBn& b, int N, float* s_scale, float* s_shift) {
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const double m = ldcg_f64(b.stat_f + n) / a.count;
    double v = ldcg_f64(b.stat_f + N + n) / a.count - m * m;
    if (v < 0.0) v = 0.0;
    const float mean = (float)m, var = (float)v;
    float rstd = rsqrtf(var + a.eps);
    rstd = rstd * (1.5f - 0.5f * (var + a.eps) * rstd * rstd);
    const float sc = rstd * b.gamma[n];
    const float sh = b.beta[n] - mean * sc;
    s_scale[n] = sc;
    s_shift[n] = sh;
    if (blockIdx.x == 0) {
      if (a.update_moving) {
        b.mmean[n] -= (b.mmean[n] - mean) * (1.f - a.momentum);
        b.mvar[n] -= (b.mvar[n] - var) * (1.f - a.momentum);
      }
      b.scale[n] = sc; b.shift[n] = sh; b.mean[n] = mean; b.rstd[n] = rstd;
    }
  }
  __syncthreads();
}
// Backward coefficients from (sum dy, sum dy*xhat); same arithmetic as bn_bwd_finalize_kernel.
CLSR_DEVINL void coop_bn_bwd(const CoopMlp& a, const CoopBn& b, int N, float* s_al, float* s_be, float* s_ga) {
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const double s1 = ldcg_f64(b.stat_b + n), s2 = ldcg_f64(b.stat_b + N + n);
    const float m1 = (float)(s1 / a.count), m2 = (float)(s2 / a.count);
    const float g = b.gamma[n], r = b.rstd[n], mu = b.mean[n];
    const float al = g * r, be = -g * r * r * m2, ga = -g * r * m1 + g * r * r * m2 * mu;
    s_al[n] = al; s_be[n] = be; s_ga[n] = ga;
    if (blockIdx.x == 0) {
      b.al[n] = al; b.be[n] = be; b.ga[n] = ga;
      b.dgamma[n] += a.add_scale * (float)s2;
      b.dbeta[n] += a.add_scale * (float)s1;
    }
  }
  __syncthreads();
}

// C[r, 4cg..4cg+3] for the rows of one pass: A rows in shared memory (row pitch lda), W[k][N] in shared memory.
// Thread (rg, cg) owns rows 4rg..4rg+3 of the pass and columns 4cg..4cg+3.
CLSR_DEVINL void tile4x4(const float* sA, int lda, const float* sW, int N, int K, int rg, int cg, float acc[4][4]) {
  const float* a0 = sA + (size_t)(rg * 4) * lda;
  const float* w = sW + cg * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 wv = *reinterpret_cast<const float4*>(w + (size_t)k * N);
    const float x0 = a0[k], x1 = a0[lda + k], x2 = a0[2 * lda + k], x3 = a0[3 * lda + k];
    acc[0][0] = fmaf(x0, wv.x, acc[0][0]); acc[0][1] = fmaf(x0, wv.y, acc[0][1]);
    acc[0][2] = fmaf(x0, wv.z, acc[0][2]); acc[0][3] = fmaf(x0, wv.w, acc[0][3]);
    acc[1][0] = fmaf(x1, wv.x, acc[1][0]); acc[1][1] = fmaf(x1, wv.y, acc[1][1]);
    acc[1][2] = fmaf(x1, wv.z, acc[1][2]); acc[1][3] = fmaf(x1, wv.w, acc[1][3]);
    acc[2][0] = fmaf(x2, wv.x, acc[2][0]); acc[2][1] = fmaf(x2, wv.y, acc[2][1]);
    acc[2][2] = fmaf(x2, wv.z, acc[2][2]); acc[2][3] = fmaf(x2, wv.w, acc[2][3]);
    acc[3][0] = fmaf(x3, wv.x, acc[3][0]); acc[3][1] = fmaf(x3, wv.y, acc[3][1]);
    acc[3][2] = fmaf(x3, wv.z, acc[3][2]); acc[3][3] = fmaf(x3, wv.w, acc[3][3]);
  }
}

struct CoopFwdSmem {
  int w0, w1, a, h0, h1, vec, total;   // float offsets
  int rc, lda, rpc;
};
// rpc: rows per CTA the buffers are sized for.  Row passes of layer 0 hold (256 / (n0/4)) * 4 rows of the input.
__host__ __device__ inline CoopFwdSmem coop_fwd_smem(int K, int n0, int n1, int rpc) {
  CoopFwdSmem s;
  s.rpc = rpc;
  s.rc = (kCoopThreads / (n0 / 4)) * 4;
  s.lda = K | 1;                       // odd pitch: the four rows a thread reads sit in different banks
  int o = 0;
  s.w0 = o; o += K * n0;
  s.w1 = o; o += n0 * n1;
  s.a = o; o += s.rc * s.lda;
  o = (o + 3) & ~3;
  s.h0 = o; o += (rpc + 4) * n0;       // + 4: a thread's four rows may run past the slice (never stored)
  s.h1 = o; o += (rpc + 4) * n1;
  s.vec = o; o += 8 * kCoopMaxN;
  s.total = o;
  return s;
}

// ---- forward: h0 = in.W0 + b0 | BN | h1 = relu(bn(h0)).W1 + b1 | BN | out = relu(bn(h1)).wo + bo ----
__global__ void __launch_bounds__(kCoopThreads, 1) mlp_fwd_coop_kernel(const CoopMlp a, int rpc_alloc) {
  extern __shared__ __align__(16) float sm[];
  __shared__ double s_stat[2 * kCoopMaxN];
  __shared__ double s_peer[kPeerSlots];
  const CoopFwdSmem L = coop_fwd_smem(a.K, a.n0, a.n1, rpc_alloc);
  float *sW0 = sm + L.w0, *sW1 = sm + L.w1, *sA = sm + L.a, *sH0 = sm + L.h0, *sH1 = sm + L.h1, *sv = sm + L.vec;
  float *sb0 = sv, *sb1 = sv + kCoopMaxN, *ssc = sv + 2 * kCoopMaxN, *ssh = sv + 3 * kCoopMaxN, *swo = sv + 4 * kCoopMaxN;
  const int tid = threadIdx.x, K = a.K, n0 = a.n0, n1 = a.n1;
  const int rpc = (a.rows + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * rpc;
  int nr = a.rows - r0;
  nr = nr < 0 ? 0 : (nr > rpc ? rpc : nr);

  coop_copy(sW0, a.w0, K * n0);
  coop_copy(sW1, a.w1, n0 * n1);
  for (int i = tid; i < n0; i += kCoopThreads) sb0[i] = a.b0[i];
  for (int i = tid; i < n1; i += kCoopThreads) { sb1[i] = a.b1[i]; swo[i] = a.wo[i]; }
  for (int i = tid; i < 2 * kCoopMaxN; i += kCoopThreads) s_stat[i] = 0.0;
  cp_async_wait_all();
  __syncthreads();

  // ---- layer 0 ----
  {
    const int ncg = n0 / 4, nrg = kCoopThreads / ncg, RC = nrg * 4;
    const int rg = tid / ncg, cg = tid - rg * ncg;
    const bool active = rg < nrg;
    float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < nr; c0 += RC) {
      {
        const int live = nr - c0 < RC ? nr - c0 : RC;
        coop_copy_rows(sA, L.lda, a.in + (size_t)(r0 + c0) * a.ld_in, a.ld_in, live, K, K);
        for (int i = live * L.lda + tid; i < RC * L.lda; i += kCoopThreads) sA[i] = 0.f;
      }
      cp_async_wait_all();
      __syncthreads();
      if (active) {
        float acc[4][4];
        tile4x4(sA, L.lda, sW0, n0, K, rg, cg, acc);
        const float4 bv = *reinterpret_cast<const float4*>(sb0 + cg * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = c0 + rg * 4 + i;
          if (r < nr) {
            const float4 v = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
            *reinterpret_cast<float4*>(sH0 + (size_t)r * n0 + cg * 4) = v;
            *reinterpret_cast<float4*>(a.h0 + (size_t)(r0 + r) * n0 + cg * 4) = v;
            cs[0] += v.x; cs[1] += v.y; cs[2] += v.z; cs[3] += v.w;
            cq[0] = fmaf(v.x, v.x, cq[0]); cq[1] = fmaf(v.y, v.y, cq[1]);
            cq[2] = fmaf(v.z, v.z, cq[2]); cq[3] = fmaf(v.w, v.w, cq[3]);
          }
        }
      }
      __syncthreads();
    }
    if (active && nr > 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&s_stat[cg * 4 + j], (double)cs[j]);
        atomicAdd(&s_stat[kCoopMaxN + cg * 4 + j], (double)cq[j]);
      }
    }
    __syncthreads();
    if (nr > 0)
      for (int n = tid; n < n0; n += kCoopThreads) {
        atomicAdd(a.bn0.stat_f + n, s_stat[n]);
        atomicAdd(a.bn0.stat_f + n0 + n, s_stat[kCoopMaxN + n]);
      }
  }
  grid_barrier(a.bar);
  coop_exchange_stats(a, a.bn0.stat_f, 2 * n0, s_peer);
  coop_bn_fwd(a, a.bn0, n0, ssc, ssh);

  // ---- layer 1 (operand: relu(bn(h0)), formed in place) ----
  for (int i = tid; i < 2 * kCoopMaxN; i += kCoopThreads) s_stat[i] = 0.0;
  for (int i = tid; i < nr * n0; i += kCoopThreads) {
    const int c = i % n0;
    sH0[i] = fmaxf(0.f, fmaf(sH0[i], ssc[c], ssh[c]));
  }
  for (int i = nr * n0 + tid; i < (nr + 4) * n0 && i < (L.rpc + 4) * n0; i += kCoopThreads) sH0[i] = 0.f;
  __syncthreads();
  {
    const int ncg = n1 / 4, nrg = kCoopThreads / ncg, RC = nrg * 4;
    const int rg = tid / ncg, cg = tid - rg * ncg;
    const bool active = rg < nrg;
    float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < nr; c0 += RC) {
      if (active && c0 + rg * 4 < nr) {
        float acc[4][4];
        tile4x4(sH0 + (size_t)c0 * n0, n0, sW1, n1, n0, rg, cg, acc);
        const float4 bv = *reinterpret_cast<const float4*>(sb1 + cg * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = c0 + rg * 4 + i;
          if (r < nr) {
            const float4 v = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
            *reinterpret_cast<float4*>(sH1 + (size_t)r * n1 + cg * 4) = v;
            *reinterpret_cast<float4*>(a.h1 + (size_t)(r0 + r) * n1 + cg * 4) = v;
            cs[0] += v.x; cs[1] += v.y; cs[2] += v.z; cs[3] += v.w;
            cq[0] = fmaf(v.x, v.x, cq[0]); cq[1] = fmaf(v.y, v.y, cq[1]);
            cq[2] = fmaf(v.z, v.z, cq[2]); cq[3] = fmaf(v.w, v.w, cq[3]);
          }
        }
      }
    }
    if (active && nr > 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&s_stat[cg * 4 + j], (double)cs[j]);
        atomicAdd(&s_stat[kCoopMaxN + cg * 4 + j], (double)cq[j]);
      }
    }
    __syncthreads();
    if (nr > 0)
      for (int n = tid; n < n1; n += kCoopThreads) {
        atomicAdd(a.bn1.stat_f + n, s_stat[n]);
        atomicAdd(a.bn1.stat_f + n1 + n, s_stat[kCoopMaxN + n]);
      }
  }
  grid_barrier(a.bar);
  coop_exchange_stats(a, a.bn1.stat_f, 2 * n1, s_peer);
  coop_bn_fwd(a, a.bn1, n1, ssc, ssh);

  // ---- output unit: one warp per row ----
  {
    const int lane = tid & 31, wid = tid >> 5;
    const float bo = a.bo[0];
    for (int r = wid; r < nr; r += kCoopThreads / 32) {
      float acc = 0.f;
      for (int n = lane; n < n1; n += 32) acc = fmaf(fmaxf(0.f, fmaf(sH1[(size_t)r * n1 + n], ssc[n], ssh[n])), swo[n], acc);
      acc = warp_sum(acc);
      if (lane == 0) a.out[r0 + r] = acc + bo;
    }
  }
}

struct CoopBwdSmem {
  int r0, r1, r2, dout, vec, total;   // float offsets
  int ldi, kp, rpc;
};
// R0: a0 -> dy0 -> dh0 [rpc+4, n0]; R1: h1 | dy1 -> dh1 [rpc+4, n1] each, later the input rows [rpc, ldi];
// R2: W1^T [n1][n0], later W0^T [n0][kp].
__host__ __device__ inline CoopBwdSmem coop_bwd_smem(int K, int n0, int n1, int rpc) {
  CoopBwdSmem s;
  s.rpc = rpc;
  s.ldi = (K + 3) & ~3;
  s.kp = (K + 3) & ~3;
  int o = 0;
  s.r0 = o; o += (rpc + 4) * n0;
  s.r1 = o;
  const int ab = 2 * (rpc + 4) * n1, c = rpc * s.ldi;
  o += ab > c ? ab : c;
  s.r2 = o;
  const int w1 = n1 * n0, w0 = n0 * s.kp;
  o += w1 > w0 ? w1 : w0;
  s.dout = o; o += (rpc + 4 + 3) & ~3;
  s.vec = o; o += 10 * kCoopMaxN;
  s.total = o;
  return s;
}

CLSR_DEVINL void coop_flush_cols(double* s_stat, double* g1, double* g2, int N, int nr) {
  __syncthreads();
  if (nr > 0)
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      atomicAdd(g1 + n, s_stat[n]);
      atomicAdd(g2 + n, s_stat[kCoopMaxN + n]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * kCoopMaxN; i += blockDim.x) s_stat[i] = 0.0;
  __syncthreads();
}

// ---- backward from dout[rows]: weight / bias / BatchNorm gradients into the dense-gradient block, d(in) ----
__global__ void __launch_bounds__(kCoopThreads, 1) mlp_bwd_coop_kernel(const CoopMlp a, int rpc_alloc) {
  extern __shared__ __align__(16) float sm[];
  __shared__ double s_stat[2 * kCoopMaxN];
  __shared__ double s_peer[kPeerSlots];
  __shared__ float s_red[kCoopMaxN];
  const CoopBwdSmem L = coop_bwd_smem(a.K, a.n0, a.n1, rpc_alloc);
  float *R0 = sm + L.r0, *sH1 = sm + L.r1, *sD1 = sm + L.r1 + (L.rpc + 4) * a.n1, *sIn = sm + L.r1, *sW = sm + L.r2;
  float *sdo = sm + L.dout, *sv = sm + L.vec;
  float *v0 = sv, *v1 = sv + kCoopMaxN, *v2 = sv + 2 * kCoopMaxN, *v3 = sv + 3 * kCoopMaxN, *v4 = sv + 4 * kCoopMaxN;
  float *sal = sv + 5 * kCoopMaxN, *sbe = sv + 6 * kCoopMaxN, *sga = sv + 7 * kCoopMaxN;
  const int tid = threadIdx.x, K = a.K, n0 = a.n0, n1 = a.n1;
  const int rpc = (a.rows + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * rpc;
  int nr = a.rows - r0;
  nr = nr < 0 ? 0 : (nr > rpc ? rpc : nr);
  const int rot = (int)blockIdx.x;   // CTAs start their gradient atomics at different entries
  // 16-byte reductions when the gradient slots allow it (n0, n1 are multiples of 4)
  const bool v4ok = ((reinterpret_cast<uintptr_t>(a.dw0) | reinterpret_cast<uintptr_t>(a.dw1)) & 15) == 0;

  for (int i = tid; i < 2 * kCoopMaxN; i += kCoopThreads) s_stat[i] = 0.0;
  for (int i = tid; i < kCoopMaxN; i += kCoopThreads) s_red[i] = 0.f;
  // layer-1 forward coefficients: v0 scale, v1 shift, v2 rstd, v3 -mean*rstd, v4 wo
  for (int n = tid; n < n1; n += kCoopThreads) {
    v0[n] = a.bn1.scale[n]; v1[n] = a.bn1.shift[n];
    const float rs = a.bn1.rstd[n];
    v2[n] = rs; v3[n] = -a.bn1.mean[n] * rs; v4[n] = a.wo[n];
  }
  // everything phases a and b read from global memory, in flight at once: h1, dout, h0 (raw, into R0), W1^T
  coop_copy(sH1, a.h1 + (size_t)r0 * n1, nr * n1);
  for (int i = nr * n1 + tid; i < (nr + 4) * n1; i += kCoopThreads) { sH1[i] = 0.f; }
  for (int i = tid; i < nr + 4; i += kCoopThreads) sdo[i] = i < nr ? __ldg(a.dout + r0 + i) : 0.f;
  coop_copy(R0, a.h0 + (size_t)r0 * n0, nr * n0);
  coop_copy(sW, a.w1T, n1 * n0);
  cp_async_wait_all();
  __syncthreads();

  // ---- phase a: dy1 = dout * wo where bn(h1) > 0; stat_b1 += (sum dy1, sum dy1 * xhat1); dwo, dbo ----
  {
    const int nrl = kCoopThreads / n1;          // row lanes
    const int rl = tid / n1, c = tid - rl * n1;
    float s1 = 0.f, s2 = 0.f, dw = 0.f;
    if (rl < nrl) {
      const float sc = v0[c], sh = v1[c], rs = v2[c], mr = v3[c], w = v4[c];
      for (int r = rl; r < nr; r += nrl) {
        const float h = sH1[(size_t)r * n1 + c], y = fmaf(h, sc, sh), d = sdo[r];
        const float dy = y > 0.f ? d * w : 0.f;
        sD1[(size_t)r * n1 + c] = dy;
        s1 += dy;
        s2 = fmaf(dy, fmaf(h, rs, mr), s2);
        dw = fmaf(fmaxf(y, 0.f), d, dw);
      }
      atomicAdd(&s_stat[c], (double)s1);
      atomicAdd(&s_stat[kCoopMaxN + c], (double)s2);
      atomicAdd(&s_red[c], dw);
    }
    for (int i = nr * n1 + tid; i < (nr + 4) * n1; i += kCoopThreads) sD1[i] = 0.f;
    __syncthreads();
    if (nr > 0) {
      for (int n = tid; n < n1; n += kCoopThreads) atomicAdd(a.dwo + n, s_red[n]);
      if (tid < 32) {
        float d = 0.f;
        for (int r = tid; r < nr; r += 32) d += sdo[r];
        d = warp_sum(d);
        if (tid == 0) atomicAdd(a.dbo, d);
      }
    }
    coop_flush_cols(s_stat, a.bn1.stat_b, a.bn1.stat_b + n1, n1, nr);
  }
  grid_barrier(a.bar);
  coop_exchange_stats(a, a.bn1.stat_b, 2 * n1, s_peer);
  coop_bn_bwd(a, a.bn1, n1, sal, sbe, sga);

  // ---- phase b: dh1 = BN1 backward; a0 = relu(bn(h0)); dW1 += a0^T.dh1; db1; dy0 = (dh1.W1^T) masked; stat_b0 ----
  for (int i = tid; i < kCoopMaxN; i += kCoopThreads) s_red[i] = 0.f;
  for (int n = tid; n < n0; n += kCoopThreads) {   // layer-0 forward coefficients
    v0[n] = a.bn0.scale[n]; v1[n] = a.bn0.shift[n];
    const float rs = a.bn0.rstd[n];
    v2[n] = rs; v3[n] = -a.bn0.mean[n] * rs;
  }
  __syncthreads();
  {
    const int nrl = kCoopThreads / n1;
    const int rl = tid / n1, c = tid - rl * n1;
    // (no bias gradient: a bias in front of a BatchNorm has none -- the BatchNorm backward makes every column of dh sum to
    //  exactly zero; accumulating it would only add rounding noise to the L2 term the regulariser contributes)
    if (rl < nrl) {
      const float al = sal[c], be = sbe[c], ga = sga[c];
      for (int r = rl; r < nr; r += nrl)
        sD1[(size_t)r * n1 + c] = fmaf(al, sD1[(size_t)r * n1 + c], fmaf(be, sH1[(size_t)r * n1 + c], ga));
    }
    for (int i = tid; i < nr * n0; i += kCoopThreads) {
      const int cc = i % n0;
      R0[i] = fmaxf(0.f, fmaf(R0[i], v0[cc], v1[cc]));
    }
    for (int i = nr * n0 + tid; i < (nr + 4) * n0; i += kCoopThreads) R0[i] = 0.f;
  }
  __syncthreads();
  {
    // dW1[j, c] += sum_r a0[r, j] * dh1[r, c]: 4 x 4 tiles, one thread per tile and round
    const int tj = n0 / 4, tc = n1 / 4, nt = tj * tc;
    for (int t = tid; t < nt && nr > 0; t += kCoopThreads) {
      const int jg = t / tc, cg = t - jg * tc;
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int r = 0; r < nr; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(R0 + (size_t)r * n0 + jg * 4);
        const float4 d = *reinterpret_cast<const float4*>(sD1 + (size_t)r * n1 + cg * 4);
        acc[0][0] = fmaf(x.x, d.x, acc[0][0]); acc[0][1] = fmaf(x.x, d.y, acc[0][1]);
        acc[0][2] = fmaf(x.x, d.z, acc[0][2]); acc[0][3] = fmaf(x.x, d.w, acc[0][3]);
        acc[1][0] = fmaf(x.y, d.x, acc[1][0]); acc[1][1] = fmaf(x.y, d.y, acc[1][1]);
        acc[1][2] = fmaf(x.y, d.z, acc[1][2]); acc[1][3] = fmaf(x.y, d.w, acc[1][3]);
        acc[2][0] = fmaf(x.z, d.x, acc[2][0]); acc[2][1] = fmaf(x.z, d.y, acc[2][1]);
        acc[2][2] = fmaf(x.z, d.z, acc[2][2]); acc[2][3] = fmaf(x.z, d.w, acc[2][3]);
        acc[3][0] = fmaf(x.w, d.x, acc[3][0]); acc[3][1] = fmaf(x.w, d.y, acc[3][1]);
        acc[3][2] = fmaf(x.w, d.z, acc[3][2]); acc[3][3] = fmaf(x.w, d.w, acc[3][3]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float* p = a.dw1 + (size_t)(jg * 4 + i) * n1 + cg * 4;
        if (v4ok) red_add_v4(p, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j) atomicAdd(p + j, acc[i][j]);
        }
      }
    }
  }
  __syncthreads();
  {
    // dy0[r, j] = sum_c dh1[r, c] * W1[j, c], masked by a0 > 0, written over a0; statistics of layer 0
    const int ncg = n0 / 4, nrg = kCoopThreads / ncg, RC = nrg * 4;
    const int rg = tid / ncg, cg = tid - rg * ncg;
    const bool active = rg < nrg;
    float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
    float rs[4], mr[4];
    if (active) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { rs[j] = v2[cg * 4 + j]; mr[j] = v3[cg * 4 + j]; }
    }
    for (int c0 = 0; c0 < nr; c0 += RC) {
      if (active && c0 + rg * 4 < nr) {
        float acc[4][4];
        tile4x4(sD1 + (size_t)c0 * n1, n1, sW, n0, n1, rg, cg, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = c0 + rg * 4 + i;
          if (r < nr) {
            float* p = R0 + (size_t)r * n0 + cg * 4;
            const float4 x = *reinterpret_cast<const float4*>(p);
            const float4 h = __ldg(reinterpret_cast<const float4*>(a.h0 + (size_t)(r0 + r) * n0 + cg * 4));
            float4 d;
            d.x = x.x > 0.f ? acc[i][0] : 0.f; d.y = x.y > 0.f ? acc[i][1] : 0.f;
            d.z = x.z > 0.f ? acc[i][2] : 0.f; d.w = x.w > 0.f ? acc[i][3] : 0.f;
            *reinterpret_cast<float4*>(p) = d;
            cs[0] += d.x; cs[1] += d.y; cs[2] += d.z; cs[3] += d.w;
            cq[0] = fmaf(d.x, fmaf(h.x, rs[0], mr[0]), cq[0]); cq[1] = fmaf(d.y, fmaf(h.y, rs[1], mr[1]), cq[1]);
            cq[2] = fmaf(d.z, fmaf(h.z, rs[2], mr[2]), cq[2]); cq[3] = fmaf(d.w, fmaf(h.w, rs[3], mr[3]), cq[3]);
          }
        }
      }
    }
    if (active && nr > 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&s_stat[cg * 4 + j], (double)cs[j]);
        atomicAdd(&s_stat[kCoopMaxN + cg * 4 + j], (double)cq[j]);
      }
    }
    coop_flush_cols(s_stat, a.bn0.stat_b, a.bn0.stat_b + n0, n0, nr);
  }
  // phase c operands (h1 / dh1 and W1^T are dead): the input rows, zero-padded to the pitch, and W0^T (pitch kp); the
  // copies land while the CTAs wait at the barrier
  coop_copy_rows(sIn, L.ldi, a.in + (size_t)r0 * a.ld_in, a.ld_in, nr, K, L.ldi);
  coop_copy_rows(sW, L.kp, a.w0T, K, n0, K, L.kp);
  grid_barrier(a.bar);
  coop_exchange_stats(a, a.bn0.stat_b, 2 * n0, s_peer);
  coop_bn_bwd(a, a.bn0, n0, sal, sbe, sga);
  cp_async_wait_all();

  // ---- phase c: dh0 = BN0 backward; db0; dW0 += in^T.dh0; din = dh0.W0^T ----
  for (int i = tid; i < kCoopMaxN; i += kCoopThreads) s_red[i] = 0.f;
  __syncthreads();
  {
    const int q4 = n0 / 4, n4 = nr * q4;
    const float4* hsrc = reinterpret_cast<const float4*>(a.h0 + (size_t)r0 * n0);
    for (int i0 = tid; i0 < n4; i0 += 4 * kCoopThreads) {
      float4 h[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kCoopThreads;
        h[u] = __ldg(hsrc + (i < n4 ? i : n4 - 1));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kCoopThreads;
        if (i < n4) {
          const int c = (i % q4) * 4;
          float4 d = *reinterpret_cast<float4*>(R0 + (size_t)i * 4);
          d.x = fmaf(sal[c], d.x, fmaf(sbe[c], h[u].x, sga[c]));
          d.y = fmaf(sal[c + 1], d.y, fmaf(sbe[c + 1], h[u].y, sga[c + 1]));
          d.z = fmaf(sal[c + 2], d.z, fmaf(sbe[c + 2], h[u].z, sga[c + 2]));
          d.w = fmaf(sal[c + 3], d.w, fmaf(sbe[c + 3], h[u].w, sga[c + 3]));
          *reinterpret_cast<float4*>(R0 + (size_t)i * 4) = d;
        }
      }
    }
  }
  __syncthreads();
  {
    // dW0[k, j] += sum_r in[r, k] * dh0[r, j]: 8 x 4 tiles
    const int tk = (K + 7) / 8, tj = n0 / 4, nt = tk * tj;
    for (int t0 = tid; t0 < nt && nr > 0; t0 += kCoopThreads) {
      const int t = (t0 + rot * 7) % nt;
      const int kg = t / tj, jg = t - kg * tj;
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      const bool second = kg * 8 + 4 < L.ldi;
      for (int r = 0; r < nr; ++r) {
        const float4 x0 = *reinterpret_cast<const float4*>(sIn + (size_t)r * L.ldi + kg * 8);
        const float4 x1 = second ? *reinterpret_cast<const float4*>(sIn + (size_t)r * L.ldi + kg * 8 + 4)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 d = *reinterpret_cast<const float4*>(R0 + (size_t)r * n0 + jg * 4);
        const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[i][0] = fmaf(xs[i], d.x, acc[i][0]); acc[i][1] = fmaf(xs[i], d.y, acc[i][1]);
          acc[i][2] = fmaf(xs[i], d.z, acc[i][2]); acc[i][3] = fmaf(xs[i], d.w, acc[i][3]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = kg * 8 + i;
        if (k < K) {
          float* p = a.dw0 + (size_t)k * n0 + jg * 4;
          if (v4ok) red_add_v4(p, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
          else {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(p + j, acc[i][j]);
          }
        }
      }
    }
  }
  {
    // din[r, k] = sum_j dh0[r, j] * W0[k, j]: 4 rows x 4 k per thread, W0^T[j][kp] in shared memory
    const int ncg = L.kp / 4, nrg = kCoopThreads / ncg > 0 ? kCoopThreads / ncg : 1, RC = nrg * 4;
    const int rg = tid / ncg, cg = tid - rg * ncg;
    const bool active = rg < nrg && cg < ncg;
    for (int c0 = 0; c0 < nr; c0 += RC) {
      if (active && c0 + rg * 4 < nr) {
        float acc[4][4];
        tile4x4(R0 + (size_t)c0 * n0, n0, sW, L.kp, n0, rg, cg, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = c0 + rg * 4 + i;
          if (r < nr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int k = cg * 4 + j;
              if (k < K) a.din[(size_t)(r0 + r) * K + k] = acc[i][j];
            }
          }
        }
      }
    }
  }
}

}  // namespace clsr
