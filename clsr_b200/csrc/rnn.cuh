// Length-masked recurrences of the CLSR graph: two GRUs (short_term_intention, causal2;
// clsr.py:161-168, 230-236) and the Time4LSTM (clsr.py:179-200; rnn_cell_implement.py:129-298).
//
// Everything that does not depend on the recurrent state (x_t . W_x + b for all gates, the time
// gates sigmoid(x.Wk + tanh(.)Tk + b), the o-gate time terms) is hoisted out of the loop into one
// large GEMM over all S*T positions (buffer PX); these kernels only run the serial part
// h . W_h per step with the recurrent weights resident in shared memory.  dynamic_rnn semantics:
// for t >= length the state is copied through and the output is zero.
//
// One CTA owns RNN_NSEQ sequences.  States live transposed in shared memory ([k][seq], row stride
// RNN_LD) so one float4 load feeds four sequences per weight element: 2 LDS per 4 FMA.
#pragma once
#include "common.cuh"

namespace clsr {

constexpr int RNN_NSEQ = 16;
constexpr int RNN_LD = 20;  // padded row stride of the [k][seq] state tiles (16-byte aligned rows)
constexpr int RNN_THREADS = 256;

CLSR_DEVINL void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Software prefetch for the serial loops: the rows a CTA will read `two steps from now` are pulled
// into L2 while the current step computes, so the per-step global loads hit L2 instead of HBM.
// Thread `idx` (0 .. 16*lines-1) touches 128-byte line `idx % lines` of sequence `idx / lines`.
CLSR_DEVINL void prefetch_row_lines(const float* base, size_t ld, int col0, int lines, int idx, int s0, int T,
                                    int tt, const int* slen) {
  const int q = idx / lines, l = idx - q * lines;
  if (q < RNN_NSEQ && tt >= 0 && tt < slen[q]) prefetch_l2(base + ((size_t)(s0 + q) * T + tt) * ld + col0 + l * 32);
}

CLSR_DEVINL int rnn_setup_len(const int* __restrict__ len, int s0, int S, int* slen) {
  __shared__ int s_tmax;
  if (threadIdx.x == 0) s_tmax = 0;
  __syncthreads();
  if (threadIdx.x < RNN_NSEQ) {
    int l = (s0 + (int)threadIdx.x < S) ? len[s0 + threadIdx.x] : 0;
    slen[threadIdx.x] = l;
    atomicMax(&s_tmax, l);
  }
  __syncthreads();
  return s_tmax;
}

__global__ void __launch_bounds__(RNN_THREADS)
gru_fwd_kernel(const float* __restrict__ PX, int ldpx, int colg, int colc,
               const float* __restrict__ h0, const float* __restrict__ Wgh, const float* __restrict__ Wch,
               const int* __restrict__ len, int S, int T, int U, float* __restrict__ gates,
               float* __restrict__ cand, float* __restrict__ hprev, float* __restrict__ rh,
               float* __restrict__ hfinal, int wglob) {
  // wglob: the recurrent weights do not fit shared memory next to the state tiles (wide states, BASELINE
  // configs 4-5): they are read from global memory instead (identical for every CTA and step: L1 / L2 resident)
  extern __shared__ __align__(16) float sm[];
  const float* sWg = wglob ? Wgh : sm;                    // [U][2U]
  const float* sWc = wglob ? Wch : sm + U * 2 * U;        // [U][U]
  float* hT = wglob ? sm : sm + U * 2 * U + U * U;        // [U][LD]
  float* rhT = hT + U * RNN_LD;
  float* uT = rhT + U * RNN_LD;
  __shared__ int slen[RNN_NSEQ];
  const int tid = threadIdx.x;
  const int s0 = blockIdx.x * RNN_NSEQ;
  const int U2 = 2 * U;
  if (!wglob) {
    for (int i = tid; i < U * U2; i += RNN_THREADS) sm[i] = Wgh[i];
    for (int i = tid; i < U * U; i += RNN_THREADS) sm[U * U2 + i] = Wch[i];
  }
  for (int i = tid; i < U * RNN_NSEQ; i += RNN_THREADS) {
    int q = i / U, k = i % U;
    int s = s0 + q;
    hT[k * RNN_LD + q] = (h0 && s < S) ? h0[(size_t)s * U + k] : 0.f;
  }
  const int tmax = rnn_setup_len(len, s0, S, slen);
  const int nt1 = (RNN_NSEQ / 4) * U2, nt2 = (RNN_NSEQ / 4) * U;

  const int pf_lines = (3 * U + 31) / 32 + 1;  // gates + candidate columns are contiguous in PX
  for (int t = 0; t < tmax; ++t) {
    if (tid < RNN_NSEQ * pf_lines) prefetch_row_lines(PX, ldpx, colg, pf_lines, tid, s0, T, t + 2, slen);
    for (int task = tid; task < nt1; task += RNN_THREADS) {
      const int qg = task / U2, j = task - qg * U2;
      float px[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int q = qg * 4 + i;
        px[i] = (t < slen[q]) ? PX[((size_t)(s0 + q) * T + t) * ldpx + colg + j] : 0.f;
      }
#pragma unroll 8
      for (int k = 0; k < U; ++k) {
        float4 h4 = *reinterpret_cast<const float4*>(&hT[k * RNN_LD + qg * 4]);
        float w = sWg[k * U2 + j];
        acc[0] = fmaf(h4.x, w, acc[0]); acc[1] = fmaf(h4.y, w, acc[1]);
        acc[2] = fmaf(h4.z, w, acc[2]); acc[3] = fmaf(h4.w, w, acc[3]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int q = qg * 4 + i;
        if (t < slen[q]) {
          float g = sigmoid_acc(acc[i] + px[i]);
          size_t row = (size_t)(s0 + q) * T + t;
          gates[row * U2 + j] = g;
          if (j < U) {
            float hp = hT[j * RNN_LD + q];
            float v = g * hp;
            rhT[j * RNN_LD + q] = v;
            rh[row * U + j] = v;
            hprev[row * U + j] = hp;
          } else {
            uT[(j - U) * RNN_LD + q] = g;
          }
        }
      }
    }
    __syncthreads();
    for (int task = tid; task < nt2; task += RNN_THREADS) {
      const int qg = task / U, j = task - qg * U;
      float px[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int q = qg * 4 + i;
        px[i] = (t < slen[q]) ? PX[((size_t)(s0 + q) * T + t) * ldpx + colc + j] : 0.f;
      }
#pragma unroll 8
      for (int k = 0; k < U; ++k) {
        float4 h4 = *reinterpret_cast<const float4*>(&rhT[k * RNN_LD + qg * 4]);
        float w = sWc[k * U + j];
        acc[0] = fmaf(h4.x, w, acc[0]); acc[1] = fmaf(h4.y, w, acc[1]);
        acc[2] = fmaf(h4.z, w, acc[2]); acc[3] = fmaf(h4.w, w, acc[3]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int q = qg * 4 + i;
        if (t < slen[q]) {
          float c = tanhf(acc[i] + px[i]);
          size_t row = (size_t)(s0 + q) * T + t;
          cand[row * U + j] = c;
          float u = uT[j * RNN_LD + q], hp = hT[j * RNN_LD + q];
          hT[j * RNN_LD + q] = u * hp + (1.f - u) * c;
        }
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < U * RNN_NSEQ; i += RNN_THREADS) {
    int q = i / U, k = i % U;
    if (s0 + q < S) hfinal[(size_t)(s0 + q) * U + k] = hT[k * RNN_LD + q];
  }
}

// BPTT of the GRU.  Writes the pre-activation gradients of every live step into dPX (columns
// colg..colg+2U, colc..colc+U; dead positions must be pre-zeroed) and the gradient of the
// initial state into dh0.
__global__ void __launch_bounds__(RNN_THREADS)
gru_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ cand,
               const float* __restrict__ hprev, const float* __restrict__ WghT,
               const float* __restrict__ WchT, const float* __restrict__ dh_final,
               const int* __restrict__ len, int S, int T, int U, float* __restrict__ dPX, int ldpx, int colg,
               int colc, float* __restrict__ dh0, int wglob) {
  extern __shared__ __align__(16) float sm[];
  const int U2 = 2 * U;
  const float* sWgT = wglob ? WghT : sm;                  // [2U][U]
  const float* sWcT = wglob ? WchT : sm + U2 * U;         // [U][U]
  float* dhT = wglob ? sm : sm + U2 * U + U * U;          // [U][LD]
  float* dhaT = dhT + U * RNN_LD;      // [U][LD]
  float* dpcT = dhaT + U * RNN_LD;     // [U][LD]
  float* dpgT = dpcT + U * RNN_LD;     // [2U][LD]
  __shared__ int slen[RNN_NSEQ];
  const int tid = threadIdx.x;
  const int s0 = blockIdx.x * RNN_NSEQ;
  if (!wglob) {
    for (int i = tid; i < U2 * U; i += RNN_THREADS) sm[i] = WghT[i];
    for (int i = tid; i < U * U; i += RNN_THREADS) sm[U2 * U + i] = WchT[i];
  }
  for (int i = tid; i < U * RNN_NSEQ; i += RNN_THREADS) {
    int q = i / U, k = i % U;
    int s = s0 + q;
    dhT[k * RNN_LD + q] = (s < S) ? dh_final[(size_t)s * U + k] : 0.f;
  }
  const int tmax = rnn_setup_len(len, s0, S, slen);
  const int ntm = (RNN_NSEQ / 4) * U;

  const int lg = (U2 + 31) / 32 + 1, lc = (U + 31) / 32 + 1;
  for (int t = tmax - 1; t >= 0; --t) {
    {
      int idx = tid;
      if (idx < RNN_NSEQ * lg) prefetch_row_lines(gates, U2, 0, lg, idx, s0, T, t - 2, slen);
      else if ((idx -= RNN_NSEQ * lg) < RNN_NSEQ * lc) prefetch_row_lines(cand, U, 0, lc, idx, s0, T, t - 2, slen);
      else if ((idx -= RNN_NSEQ * lc) < RNN_NSEQ * lc) prefetch_row_lines(hprev, U, 0, lc, idx, s0, T, t - 2, slen);
    }
    for (int e = tid; e < RNN_NSEQ * U; e += RNN_THREADS) {
      const int q = e / U, j = e - q * U;
      const float dhn = dhT[j * RNN_LD + q];
      if (t < slen[q]) {
        size_t row = (size_t)(s0 + q) * T + t;
        float u = gates[row * U2 + U + j], c = cand[row * U + j], hp = hprev[row * U + j];
        float du = dhn * (hp - c), dc = dhn * (1.f - u);
        float dpc = dc * (1.f - c * c), dpu = du * u * (1.f - u);
        dhaT[j * RNN_LD + q] = dhn * u;
        dpcT[j * RNN_LD + q] = dpc;
        dpgT[(U + j) * RNN_LD + q] = dpu;
        dPX[row * ldpx + colc + j] = dpc;
        dPX[row * ldpx + colg + U + j] = dpu;
      } else {
        dhaT[j * RNN_LD + q] = dhn;
        dpcT[j * RNN_LD + q] = 0.f;
        dpgT[(U + j) * RNN_LD + q] = 0.f;
      }
    }
    __syncthreads();
    for (int task = tid; task < ntm; task += RNN_THREADS) {
      const int qg = task / U, k = task - qg * U;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
      for (int j = 0; j < U; ++j) {
        float4 d4 = *reinterpret_cast<const float4*>(&dpcT[j * RNN_LD + qg * 4]);
        float w = sWcT[j * U + k];
        acc[0] = fmaf(d4.x, w, acc[0]); acc[1] = fmaf(d4.y, w, acc[1]);
        acc[2] = fmaf(d4.z, w, acc[2]); acc[3] = fmaf(d4.w, w, acc[3]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int q = qg * 4 + i;
        if (t < slen[q]) {
          size_t row = (size_t)(s0 + q) * T + t;
          float r = gates[row * U2 + k], hp = hprev[row * U + k];
          float dr = acc[i] * hp;
          dhaT[k * RNN_LD + q] += acc[i] * r;
          float dpr = dr * r * (1.f - r);
          dpgT[k * RNN_LD + q] = dpr;
          dPX[row * ldpx + colg + k] = dpr;
        } else {
          dpgT[k * RNN_LD + q] = 0.f;
        }
      }
    }
    __syncthreads();
    for (int task = tid; task < ntm; task += RNN_THREADS) {
      const int qg = task / U, k = task - qg * U;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
      for (int j = 0; j < U2; ++j) {
        float4 d4 = *reinterpret_cast<const float4*>(&dpgT[j * RNN_LD + qg * 4]);
        float w = sWgT[j * U + k];
        acc[0] = fmaf(d4.x, w, acc[0]); acc[1] = fmaf(d4.y, w, acc[1]);
        acc[2] = fmaf(d4.z, w, acc[2]); acc[3] = fmaf(d4.w, w, acc[3]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int q = qg * 4 + i;
        dhT[k * RNN_LD + q] = dhaT[k * RNN_LD + q] + acc[i];
      }
    }
    __syncthreads();
  }
  if (dh0) {
    for (int i = tid; i < U * RNN_NSEQ; i += RNN_THREADS) {
      int q = i / U, k = i % U;
      if (s0 + q < S) dh0[(size_t)(s0 + q) * U + k] = dhT[k * RNN_LD + q];
    }
  }
}

// Time4LSTM forward.  PX columns: colL..colL+4H = x.K_x + b (+ time terms on the o slice),
// colTN / colTL = complete pre-activations of the two time gates.
__global__ void __launch_bounds__(RNN_THREADS)
lstm_fwd_kernel(const float* __restrict__ PX, int ldpx, int colL, int colTN, int colTL,
                const float* __restrict__ Km, const int* __restrict__ len, int S, int T, int H,
                float* __restrict__ gates4, float* __restrict__ cprev, float* __restrict__ mprev,
                float* __restrict__ R, int wglob) {
  extern __shared__ __align__(16) float sm[];
  const int H4 = 4 * H;
  const float* sK = wglob ? Km : sm;           // [H][4H]
  float* mT = wglob ? sm : sm + H * H4;        // [H][LD]
  float* cT = mT + H * RNN_LD;         // [H][LD]
  float* preT = cT + H * RNN_LD;       // [4H][LD]
  __shared__ int slen[RNN_NSEQ];
  const int tid = threadIdx.x;
  const int s0 = blockIdx.x * RNN_NSEQ;
  if (!wglob)
    for (int i = tid; i < H * H4; i += RNN_THREADS) sm[i] = Km[i];
  for (int i = tid; i < H * RNN_LD; i += RNN_THREADS) { mT[i] = 0.f; cT[i] = 0.f; }
  const int tmax = rnn_setup_len(len, s0, S, slen);
  const int nt1 = (RNN_NSEQ / 4) * H4;

  const int pf_lines = (6 * H + 31) / 32 + 1;  // lstm gates + both time gates are contiguous in PX
  for (int t = 0; t < tmax; ++t) {
    if (tid < RNN_NSEQ * pf_lines) prefetch_row_lines(PX, ldpx, colL, pf_lines, tid, s0, T, t + 2, slen);
    for (int task = tid; task < nt1; task += RNN_THREADS) {
      const int qg = task / H4, j = task - qg * H4;
      float px[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int q = qg * 4 + i;
        px[i] = (t < slen[q]) ? PX[((size_t)(s0 + q) * T + t) * ldpx + colL + j] : 0.f;
      }
#pragma unroll 8
      for (int k = 0; k < H; ++k) {
        float4 m4 = *reinterpret_cast<const float4*>(&mT[k * RNN_LD + qg * 4]);
        float w = sK[k * H4 + j];
        acc[0] = fmaf(m4.x, w, acc[0]); acc[1] = fmaf(m4.y, w, acc[1]);
        acc[2] = fmaf(m4.z, w, acc[2]); acc[3] = fmaf(m4.w, w, acc[3]);
      }
      float4 o4 = make_float4(acc[0] + px[0], acc[1] + px[1], acc[2] + px[2], acc[3] + px[3]);
      *reinterpret_cast<float4*>(&preT[j * RNN_LD + qg * 4]) = o4;
    }
    __syncthreads();
    for (int e = tid; e < RNN_NSEQ * H; e += RNN_THREADS) {
      const int q = e / H, j = e - q * H;
      const int s = s0 + q;
      if (s >= S) continue;
      size_t row = (size_t)s * T + t;
      if (t < slen[q]) {
        float gi = sigmoid_acc(preT[j * RNN_LD + q]);
        float gj = tanhf(preT[(H + j) * RNN_LD + q]);
        float gf = sigmoid_acc(preT[(2 * H + j) * RNN_LD + q] + 1.0f);
        float go = sigmoid_acc(preT[(3 * H + j) * RNN_LD + q]);
        float sn = sigmoid_acc(PX[row * ldpx + colTN + j]);
        float sl = sigmoid_acc(PX[row * ldpx + colTL + j]);
        float cp = cT[j * RNN_LD + q], mp = mT[j * RNN_LD + q];
        float cn = gf * sl * cp + gi * sn * gj;
        float mn = go * tanhf(cn);
        gates4[row * H4 + j] = gi; gates4[row * H4 + H + j] = gj;
        gates4[row * H4 + 2 * H + j] = gf; gates4[row * H4 + 3 * H + j] = go;
        cprev[row * H + j] = cp; mprev[row * H + j] = mp;
        R[row * H + j] = mn;
        cT[j * RNN_LD + q] = cn; mT[j * RNN_LD + q] = mn;
      } else {
        R[row * H + j] = 0.f;
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < RNN_NSEQ * H; e += RNN_THREADS) {
    const int q = e / H, j = e - q * H;
    const int s = s0 + q;
    if (s >= S) continue;
    for (int t = tmax; t < T; ++t) R[((size_t)s * T + t) * H + j] = 0.f;
  }
}

__global__ void __launch_bounds__(RNN_THREADS)
lstm_bwd_kernel(const float* __restrict__ PX, int ldpx, int colL, int colTN, int colTL,
                const float* __restrict__ gates4, const float* __restrict__ cprev,
                const float* __restrict__ KmT, const float* __restrict__ dR, const int* __restrict__ len,
                int S, int T, int H, float* __restrict__ dPX, int wglob) {
  extern __shared__ __align__(16) float sm[];
  const int H4 = 4 * H;
  const float* sKT = wglob ? KmT : sm;         // [4H][H]
  float* dmT = wglob ? sm : sm + H4 * H;       // [H][LD]
  float* dcT = dmT + H * RNN_LD;       // [H][LD]
  float* dpT = dcT + H * RNN_LD;       // [4H][LD]
  __shared__ int slen[RNN_NSEQ];
  const int tid = threadIdx.x;
  const int s0 = blockIdx.x * RNN_NSEQ;
  if (!wglob)
    for (int i = tid; i < H4 * H; i += RNN_THREADS) sm[i] = KmT[i];
  for (int i = tid; i < H * RNN_LD; i += RNN_THREADS) { dmT[i] = 0.f; dcT[i] = 0.f; }
  const int tmax = rnn_setup_len(len, s0, S, slen);
  const int ntm = (RNN_NSEQ / 4) * H;

  const int l4 = (H4 + 31) / 32 + 1, l1 = (H + 31) / 32 + 1, l2 = (2 * H + 31) / 32 + 1;
  for (int t = tmax - 1; t >= 0; --t) {
    {
      int idx = tid;
      if (idx < RNN_NSEQ * l4) prefetch_row_lines(gates4, H4, 0, l4, idx, s0, T, t - 2, slen);
      else if ((idx -= RNN_NSEQ * l4) < RNN_NSEQ * l1) prefetch_row_lines(cprev, H, 0, l1, idx, s0, T, t - 2, slen);
      else if ((idx -= RNN_NSEQ * l1) < RNN_NSEQ * l1) prefetch_row_lines(dR, H, 0, l1, idx, s0, T, t - 2, slen);
      else if ((idx -= RNN_NSEQ * l1) < RNN_NSEQ * l2) prefetch_row_lines(PX, ldpx, colTN, l2, idx, s0, T, t - 2, slen);
    }
    for (int e = tid; e < RNN_NSEQ * H; e += RNN_THREADS) {
      const int q = e / H, j = e - q * H;
      float dpi = 0.f, dpj = 0.f, dpf = 0.f, dpo = 0.f;
      if (t < slen[q]) {
        size_t row = (size_t)(s0 + q) * T + t;
        float gi = gates4[row * H4 + j], gj = gates4[row * H4 + H + j];
        float gf = gates4[row * H4 + 2 * H + j], go = gates4[row * H4 + 3 * H + j];
        float cp = cprev[row * H + j];
        float sn = sigmoid_acc(PX[row * ldpx + colTN + j]);
        float sl = sigmoid_acc(PX[row * ldpx + colTL + j]);
        float cn = gf * sl * cp + gi * sn * gj;
        float tc = tanhf(cn);
        float dm = dmT[j * RNN_LD + q] + dR[row * H + j];
        dpo = dm * tc * go * (1.f - go);
        float dcn = dcT[j * RNN_LD + q] + dm * go * (1.f - tc * tc);
        dpf = dcn * sl * cp * gf * (1.f - gf);
        float dsl = dcn * gf * cp * sl * (1.f - sl);
        dpi = dcn * sn * gj * gi * (1.f - gi);
        float dsn = dcn * gi * gj * sn * (1.f - sn);
        dpj = dcn * gi * sn * (1.f - gj * gj);
        dcT[j * RNN_LD + q] = dcn * gf * sl;
        float* dp = dPX + row * ldpx;
        dp[colL + j] = dpi; dp[colL + H + j] = dpj; dp[colL + 2 * H + j] = dpf; dp[colL + 3 * H + j] = dpo;
        dp[colTN + j] = dsn; dp[colTL + j] = dsl;
      }
      dpT[j * RNN_LD + q] = dpi; dpT[(H + j) * RNN_LD + q] = dpj;
      dpT[(2 * H + j) * RNN_LD + q] = dpf; dpT[(3 * H + j) * RNN_LD + q] = dpo;
    }
    __syncthreads();
    for (int task = tid; task < ntm; task += RNN_THREADS) {
      const int qg = task / H, k = task - qg * H;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
      for (int j = 0; j < H4; ++j) {
        float4 d4 = *reinterpret_cast<const float4*>(&dpT[j * RNN_LD + qg * 4]);
        float w = sKT[j * H + k];
        acc[0] = fmaf(d4.x, w, acc[0]); acc[1] = fmaf(d4.y, w, acc[1]);
        acc[2] = fmaf(d4.z, w, acc[2]); acc[3] = fmaf(d4.w, w, acc[3]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int q = qg * 4 + i;
        if (t < slen[q]) dmT[k * RNN_LD + q] = acc[i];
      }
    }
    __syncthreads();
  }
}

// TNL[p, 0:H] = tanh(time_to_now[p]*w1 + b1); TNL[p, H:2H] = tanh(time_from_first[p]*w2 + b2)
// (rnn_cell_implement.py:200-205).  Times are addressed through the (seq, t) stride map.
__global__ void time_feat_kernel(const float* __restrict__ ttn, const float* __restrict__ tfa, int seq_stride,
                                 int T, const float* __restrict__ w1, const float* __restrict__ b1,
                                 const float* __restrict__ w2, const float* __restrict__ b2, int H,
                                 float* __restrict__ TNL, long long npos) {
  const int H2 = 2 * H;
  long long n = npos * H2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    long long p = i / H2;
    int j = (int)(i - p * H2);
    long long s = p / T;
    long long io = s * seq_stride + (p - s * T);
    float v = (j < H) ? tanhf(fmaf(ttn[io], w1[j], b1[j])) : tanhf(fmaf(tfa[io], w2[j - H], b2[j - H]));
    TNL[i] = v;
  }
}

// Gradients of the four time-input vectors: dpre = dTNL*(1-TNL^2); dw += sum dpre*time, db += sum dpre.
// blockDim = (NX >= 2H channels, NY row lanes); each CTA reduces a slab of positions: the NY row lanes
// stride through the slab (coalesced 2H-wide rows), then combine through shared memory.
__global__ void time_feat_bwd_kernel(const float* __restrict__ dTNL, const float* __restrict__ TNL,
                                     const float* __restrict__ ttn, const float* __restrict__ tfa,
                                     int seq_stride, int T, int H, long long npos, int pos_per_cta,
                                     float* __restrict__ dw1, float* __restrict__ db1,
                                     float* __restrict__ dw2, float* __restrict__ db2) {
  extern __shared__ float red[];  // [2][NY][2H]
  const int H2 = 2 * H;
  const int j = threadIdx.x, ty = threadIdx.y, ny = blockDim.y;
  long long p0 = (long long)blockIdx.x * pos_per_cta;
  long long p1 = min(npos, p0 + pos_per_cta);
  float aw = 0.f, ab = 0.f;
  if (j < H2) {
    for (long long p = p0 + ty; p < p1; p += ny) {
      long long s = p / T;
      long long io = s * seq_stride + (p - s * T);
      float y = TNL[p * H2 + j];
      float d = dTNL[p * H2 + j] * (1.f - y * y);
      float tm = (j < H) ? ttn[io] : tfa[io];
      aw = fmaf(d, tm, aw);
      ab += d;
    }
    red[(0 * ny + ty) * H2 + j] = aw;
    red[(1 * ny + ty) * H2 + j] = ab;
  }
  __syncthreads();
  if (j < H2 && ty == 0) {
    float sw = 0.f, sb = 0.f;
    for (int y = 0; y < ny; ++y) { sw += red[(0 * ny + y) * H2 + j]; sb += red[(1 * ny + y) * H2 + j]; }
    if (j < H) { atomicAdd(dw1 + j, sw); atomicAdd(db1 + j, sb); }
    else { atomicAdd(dw2 + j - H, sw); atomicAdd(db2 + j - H, sb); }
  }
}

// 16-byte vectorised forms of the two time-feature kernels (H multiple of 4): one thread = one column quad of
// one position, a quarter of the index arithmetic and of the memory instructions.
__global__ void time_feat_v4_kernel(const float* __restrict__ ttn, const float* __restrict__ tfa, int seq_stride,
                                    int T, const float* __restrict__ w1, const float* __restrict__ b1,
                                    const float* __restrict__ w2, const float* __restrict__ b2, int H,
                                    float* __restrict__ TNL, long long npos) {
  const int Q4 = (2 * H) >> 2;
  const long long n = npos * Q4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / Q4;
    const int j = (int)(i - p * Q4) * 4;
    const long long s = p / T;
    const long long io = s * seq_stride + (p - s * T);
    const bool first = j < H;
    const float tm = first ? ttn[io] : tfa[io];
    const float* w = first ? w1 + j : w2 + (j - H);   // scalar loads: the vectors sit at arbitrary offsets of the
    const float* b = first ? b1 + j : b2 + (j - H);   // dense-variable buffer
    float4 v;
    v.x = tanhf(fmaf(tm, __ldg(w), __ldg(b))); v.y = tanhf(fmaf(tm, __ldg(w + 1), __ldg(b + 1)));
    v.z = tanhf(fmaf(tm, __ldg(w + 2), __ldg(b + 2))); v.w = tanhf(fmaf(tm, __ldg(w + 3), __ldg(b + 3)));
    reinterpret_cast<float4*>(TNL)[i] = v;
  }
}

// blockDim = (2H / 4 column quads, NY row lanes); dynamic shared memory 2 * NY * 2H floats.
__global__ void time_feat_bwd_v4_kernel(const float* __restrict__ dTNL, const float* __restrict__ TNL,
                                        const float* __restrict__ ttn, const float* __restrict__ tfa,
                                        int seq_stride, int T, int H, long long npos, int pos_per_cta,
                                        float* __restrict__ dw1, float* __restrict__ db1,
                                        float* __restrict__ dw2, float* __restrict__ db2) {
  extern __shared__ float red[];  // [2][NY][2H]
  const int H2 = 2 * H, Q4 = H2 >> 2;
  const int q = threadIdx.x, ty = threadIdx.y, ny = blockDim.y;
  const int j = q * 4;
  const long long p0 = (long long)blockIdx.x * pos_per_cta;
  const long long p1 = min(npos, p0 + pos_per_cta);
  float aw[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
  const float* tsrc = j < H ? ttn : tfa;
#pragma unroll 4
  for (long long p = p0 + ty; p < p1; p += ny) {
    const long long s = p / T;
    const float tm = tsrc[s * seq_stride + (p - s * T)];
    const float4 y = ldg_stream(reinterpret_cast<const float4*>(TNL) + p * Q4 + q);
    const float4 g = ldg_stream(reinterpret_cast<const float4*>(dTNL) + p * Q4 + q);
    const float d0 = g.x * (1.f - y.x * y.x), d1 = g.y * (1.f - y.y * y.y);
    const float d2 = g.z * (1.f - y.z * y.z), d3 = g.w * (1.f - y.w * y.w);
    aw[0] = fmaf(d0, tm, aw[0]); aw[1] = fmaf(d1, tm, aw[1]); aw[2] = fmaf(d2, tm, aw[2]); aw[3] = fmaf(d3, tm, aw[3]);
    ab[0] += d0; ab[1] += d1; ab[2] += d2; ab[3] += d3;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red[(0 * ny + ty) * H2 + j + i] = aw[i];
    red[(1 * ny + ty) * H2 + j + i] = ab[i];
  }
  __syncthreads();
  for (int c = ty * Q4 + q; c < H2; c += ny * Q4) {
    float sw = 0.f, sb = 0.f;
    for (int y = 0; y < ny; ++y) { sw += red[(0 * ny + y) * H2 + c]; sb += red[(1 * ny + y) * H2 + c]; }
    if (c < H) { atomicAdd(dw1 + c, sw); atomicAdd(db1 + c, sb); }
    else { atomicAdd(dw2 + c - H, sw); atomicAdd(db2 + c - H, sb); }
  }
}

}  // namespace clsr
