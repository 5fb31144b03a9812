// fp32 SIMT linear-layer kernels with fused operand prologues and epilogues.
//
//   gemm_kernel : C[M,N] (+)= op_a(A)[M,K] . W[K,N]  (+ bias / row-broadcast terms),
//                 optional per-column statistics for BatchNorm (forward: sum, sum of squares;
//                 backward: sum dy, sum dy*xhat) accumulated in double.
//   dw_kernel   : dW[K,N] += op_a(A)[M,K]^T . op_b(B)[M,N]   (weight gradients, reduction over
//                 the long M axis), optional column sums of op_b(B) (bias gradients).
//
// The operand "prologues" let one GEMM consume un-materialised activations:
//   BN-normalise+ReLU of a stored pre-activation, the BatchNorm-backward affine form
//   a*dy + b*h + c, the attention feature map [a, a*q] (clsr.py:368-370 after folding the
//   query-only and difference terms into the weights), and the per-target product a*target
//   addressed through the (sequence, position) <- (row, position) group map.
#pragma once
#include "common.cuh"

namespace clsr {

enum AMode : int {
  A_PLAIN = 0,    // A[m,k]
  A_BNRELU = 1,   // max(0, A[m,k]*v0[k] + v1[k])
  A_AFFINE2 = 2,  // v0[k]*A[m,k] + v1[k]*A2[m,k] + v2[k]
  A_CATMUL = 3,   // k<W1 ? A[m,k] : A[m,off+k-W1] * A2[(m/T), k-W1]
  A_MULROW = 4,   // b=m/T,t=m%T,s=b/G : A[(s*T+t), off+k] * A2[b, k]
  A_CAT2ROW = 5,  // k<W1 ? A[(m/G), k] : A2[m, k-W1]
};

struct AOp {
  int mode = A_PLAIN;
  const float* A = nullptr;
  const float* A2 = nullptr;
  int lda = 0, lda2 = 0;
  const float* v0 = nullptr;
  const float* v1 = nullptr;
  const float* v2 = nullptr;
  int T = 1, G = 1, W1 = 0, off = 0;

  CLSR_DEVINL float load(int m, int k) const {
    switch (mode) {
      case A_PLAIN:
        return A[(size_t)m * lda + k];
      case A_BNRELU:
        return fmaxf(0.0f, fmaf(A[(size_t)m * lda + k], v0[k], v1[k]));
      case A_AFFINE2:
        return fmaf(v0[k], A[(size_t)m * lda + k], fmaf(v1[k], A2[(size_t)m * lda2 + k], v2[k]));
      case A_CATMUL: {
        if (k < W1) return A[(size_t)m * lda + k];
        int kk = k - W1;
        return A[(size_t)m * lda + off + kk] * A2[(size_t)(m / T) * lda2 + kk];
      }
      case A_MULROW: {
        int b = m / T, t = m - b * T, s = b / G;
        return A[((size_t)s * T + t) * lda + off + k] * A2[(size_t)b * lda2 + k];
      }
      case A_CAT2ROW: {
        if (k < W1) return A[(size_t)(m / G) * lda + k];
        return A2[(size_t)m * lda2 + (k - W1)];
      }
    }
    return 0.0f;
  }
};

enum EFlags : int {
  E_ACCUM = 1,       // C += result
  E_ROWBIAS = 2,     // + rb[(m / rbT), n]
  E_GROUPADD = 4,    // + ga[((m/T)/G*T + m%T), n]
  E_RELUMASK = 8,    // result *= (hpre*scale + shift > 0)
  E_STAT_XHAT = 16,  // second statistic is v*xhat (BN backward) instead of v*v
};

struct EpiOp {
  float* C = nullptr;
  int ldc = 0;
  int flags = 0;
  const float* bias = nullptr;
  const float* rb = nullptr;
  int ldrb = 0, rbT = 1;
  const float* ga = nullptr;
  int ldga = 0, T = 1, G = 1;
  const float* hpre = nullptr;  // E_RELUMASK / E_STAT_XHAT
  int ldh = 0;
  const float* scale = nullptr;
  const float* shift = nullptr;
  const float* mean = nullptr;
  const float* rstd = nullptr;
  double* stat = nullptr;  // [2][N] (STATS kernels)
};

template <int TM, int TN, int NTX, bool STATS>
__global__ void __launch_bounds__(256)
gemm_kernel(int M, int N, int K, AOp a, const float* __restrict__ W, int ldw, EpiOp ep) {
  constexpr int NTY = 256 / NTX;
  constexpr int BM = TM * NTY;
  constexpr int BN = TN * NTX;
  constexpr int BK = 16;
  static_assert(TM % 4 == 0, "TM must be a multiple of 4");
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ float Ws[BK][BN];
  __shared__ float part[STATS ? 2 : 1][STATS ? NTY : 1][STATS ? BN : 1];  // per row-group column partials
  __shared__ double dacc[STATS ? 2 : 1][STATS ? BN : 1];

  const int tid = threadIdx.x;
  const int tx = tid % NTX, ty = tid / NTX;
  const int n0 = blockIdx.y * BN;
  const int ntiles = (M + BM - 1) / BM;

  if (STATS) {
    for (int c = tid; c < BN; c += 256) { dacc[0][c] = 0.0; dacc[1][c] = 0.0; }
    __syncthreads();
  }

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int m0 = tile * BM;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
      for (int idx = tid; idx < BM * BK; idx += 256) {
        int r = idx / BK, kk = idx % BK;
        int m = m0 + r, k = k0 + kk;
        As[kk][r] = (m < M && k < K) ? a.load(m, k) : 0.f;
      }
      for (int idx = tid; idx < BK * BN; idx += 256) {
        int kk = idx / BN, c = idx % BN;
        int n = n0 + c, k = k0 + kk;
        Ws[kk][c] = (k < K && n < N) ? W[(size_t)k * ldw + n] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float av[TM], bv[TN];
#pragma unroll
        for (int i = 0; i < TM; i += 4) {
          float4 t = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
          av[i] = t.x; av[i + 1] = t.y; av[i + 2] = t.z; av[i + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) bv[j] = Ws[kk][tx + j * NTX];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }

    float s1[TN], s2[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ty * TM + i;
      if (m >= M) continue;
      int garow = 0;
      if (ep.flags & E_GROUPADD) {
        int b = m / ep.T, t = m - b * ep.T;
        garow = (b / ep.G) * ep.T + t;
      }
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx + j * NTX;
        if (n >= N) continue;
        float v = acc[i][j];
        if (ep.bias) v += ep.bias[n];
        if (ep.flags & E_ROWBIAS) v += ep.rb[(size_t)(m / ep.rbT) * ep.ldrb + n];
        if (ep.flags & E_GROUPADD) v += ep.ga[(size_t)garow * ep.ldga + n];
        float hp = 0.f;
        if (ep.flags & (E_RELUMASK | E_STAT_XHAT)) hp = ep.hpre[(size_t)m * ep.ldh + n];
        if (ep.flags & E_RELUMASK) {
          if (!(fmaf(hp, ep.scale[n], ep.shift[n]) > 0.f)) v = 0.f;
        }
        float* cp = ep.C + (size_t)m * ep.ldc + n;
        if (ep.flags & E_ACCUM) v += *cp;
        *cp = v;
        if (STATS) {
          s1[j] += v;
          s2[j] += (ep.flags & E_STAT_XHAT) ? v * ((hp - ep.mean[n]) * ep.rstd[n]) : v * v;
        }
      }
    }
    if (STATS) {
      // plain stores of the per-thread partials, then one thread per column sums the NTY row groups
      // (shared-memory fp32 atomics are CAS loops and serialise badly here)
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        part[0][ty][tx + j * NTX] = s1[j];
        part[1][ty][tx + j * NTX] = s2[j];
      }
      __syncthreads();
      for (int c = tid; c < BN; c += 256) {
        float a1 = 0.f, a2 = 0.f;
#pragma unroll 8
        for (int y = 0; y < NTY; ++y) { a1 += part[0][y][c]; a2 += part[1][y][c]; }
        dacc[0][c] += (double)a1; dacc[1][c] += (double)a2;
      }
      __syncthreads();
    }
  }
  if (STATS) {
    for (int c = tid; c < BN; c += 256) {
      int n = n0 + c;
      if (n < N) {
        atomicAdd(ep.stat + n, dacc[0][c]);
        atomicAdd(ep.stat + N + n, dacc[1][c]);
      }
    }
  }
}

// dW[K,N] += op_a(A)^T . op_b(B) over rows [0,M); colsum[n] += sum_m op_b(B)[m,n].
template <bool COLSUM>
__global__ void __launch_bounds__(256)
dw_kernel(int M, int K, int N, AOp a, AOp b, float* __restrict__ dW, int lddw,
          float* __restrict__ colsum, int rows_per_cta) {
  constexpr int TK = 64, TNN = 64, MR = 32;
  __shared__ __align__(16) float As[MR][TK];
  __shared__ __align__(16) float Bs[MR][TNN];
  const int tid = threadIdx.x;
  const int ti = tid / 16, tj = tid % 16;
  const int k0 = blockIdx.y * TK, n0 = blockIdx.z * TNN;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(M, r0 + rows_per_cta);
  float acc[4][4];
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int mr = r0; mr < r1; mr += MR) {
    for (int idx = tid; idx < MR * TK; idx += 256) {
      int r = idx / TK, kk = idx % TK;
      int m = mr + r, k = k0 + kk;
      As[r][kk] = (m < r1 && k < K) ? a.load(m, k) : 0.f;
    }
    for (int idx = tid; idx < MR * TNN; idx += 256) {
      int r = idx / TNN, c = idx % TNN;
      int m = mr + r, n = n0 + c;
      Bs[r][c] = (m < r1 && n < N) ? b.load(m, n) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < MR; ++r) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[r][ti * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[r][tj * 4]);
      float av[4] = {a4.x, a4.y, a4.z, a4.w};
      float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      if (COLSUM) {
#pragma unroll
        for (int j = 0; j < 4; ++j) cs[j] += bv[j];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int k = k0 + ti * 4 + i;
    if (k >= K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tj * 4 + j;
      if (n < N) atomicAdd(&dW[(size_t)k * lddw + n], acc[i][j]);
    }
  }
  if (COLSUM && blockIdx.y == 0 && ti == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tj * 4 + j;
      if (n < N) atomicAdd(&colsum[n], cs[j]);
    }
  }
}

}  // namespace clsr
