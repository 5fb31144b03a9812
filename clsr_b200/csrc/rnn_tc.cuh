// Tensor-core versions of the serial parts of the three recurrences (rnn.cuh holds the fp32 SIMT
// kernels and the description of what is hoisted out of the loops).
//
// The SIMT kernels are bound by shared-memory bandwidth: every step re-reads the recurrent weights
// (2 shared loads per 4 FMAs).  Here one CTA still owns 16 sequences - exactly the M of one
// mma.sync.m16n8k16 - and warp b owns hidden units 8b .. 8b+7 of EVERY gate:
//   * the recurrent weights live in registers as bf16 hi/lo B fragments for the whole kernel;
//   * the accumulator fragment of thread (g, t) holds sequences g, g+8 x units 8b+2t, 8b+2t+1 for all
//     gates, so the gate non-linearities, the cell / hidden state and (in BPTT) the state gradients
//     are thread-local registers that never touch shared memory;
//   * only the matmul A operand (h, r*h, or the pre-activation gradients) crosses warps: it is split
//     into bf16 hi + lo once by its producer and stored as two [16][K] tiles whose row pitch
//     (K/2 + 4 words) makes every fragment load conflict-free.
// Products are hi*hi + lo*hi + hi*lo with fp32 accumulation (same split as tc_gemm.cuh, ~2^-16
// relative operand error).  tcgen05 does not apply: its smallest M is 64 rows, and the recurrence
// exposes only as many rows per step as there are sequences in a CTA.
#pragma once
#include <cuda_bf16.h>

#include "rnn.cuh"

namespace clsr {
namespace rtc {

constexpr int NSEQ = 16;

CLSR_DEVINL void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// (x, y) -> packed bf16 hi pair and lo pair (low half = x)
CLSR_DEVINL void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 hb = __floats2bfloat162_rn(x, y);
  hi = *reinterpret_cast<uint32_t*>(&hb);
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  __nv_bfloat162 lb = __floats2bfloat162_rn(x - h0, y - h1);
  lo = *reinterpret_cast<uint32_t*>(&lb);
}

// Shared A-operand tile: 16 rows x K bf16 (hi and lo), row pitch in 32-bit words.
template <int K>
struct ATile {
  static constexpr int KP = (K + 15) / 16 * 16;
  static constexpr int NKS = KP / 16;
  static constexpr int PW = KP / 2 + 4;   // == 4 mod 8: the 32 lanes of a fragment load hit 32 banks
  uint32_t hi[NSEQ * PW];
  uint32_t lo[NSEQ * PW];
  CLSR_DEVINL void zero(int tid, int nthreads) {
    for (int i = tid; i < NSEQ * PW; i += nthreads) { hi[i] = 0u; lo[i] = 0u; }
  }
  // store the pair (x, y) at row r, columns c, c+1 (c even)
  CLSR_DEVINL void put(int r, int c, float x, float y) {
    uint32_t h, l;
    split2(x, y, h, l);
    hi[r * PW + (c >> 1)] = h;
    lo[r * PW + (c >> 1)] = l;
  }
  // A fragments of k-step ks for thread (g, t)
  CLSR_DEVINL void frag(int ks, int g, int t, uint32_t (&ah)[4], uint32_t (&al)[4]) const {
    const int w0 = g * PW + ks * 8 + t, w1 = (g + 8) * PW + ks * 8 + t;
    ah[0] = hi[w0]; ah[1] = hi[w1]; ah[2] = hi[w0 + 4]; ah[3] = hi[w1 + 4];
    al[0] = lo[w0]; al[1] = lo[w1]; al[2] = lo[w0 + 4]; al[3] = lo[w1 + 4];
  }
};

// B fragments (bf16 hi / lo) of column n of a row-major fp32 matrix W[K][ldw]; rows >= K are zero.
template <int NKS>
struct BFrag {
  uint32_t hi[NKS][2], lo[NKS][2];
  CLSR_DEVINL void load(const float* __restrict__ W, int ldw, int K, int n, int t) {
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int k = ks * 16 + 2 * t + 8 * half;
        const float w0 = k < K ? W[(size_t)k * ldw + n] : 0.f;
        const float w1 = k + 1 < K ? W[(size_t)(k + 1) * ldw + n] : 0.f;
        split2(w0, w1, hi[ks][half], lo[ks][half]);
      }
  }
};

// c += A . B over all k-steps, three bf16 products per step
template <int K>
CLSR_DEVINL void matmul(float (&c)[4], const ATile<K>& a, const BFrag<ATile<K>::NKS>& b, int g, int t) {
#pragma unroll
  for (int ks = 0; ks < ATile<K>::NKS; ++ks) {
    uint32_t ah[4], al[4];
    a.frag(ks, g, t, ah, al);
    mma_bf16(c, ah, b.hi[ks]);
    mma_bf16(c, al, b.hi[ks]);
    mma_bf16(c, ah, b.lo[ks]);
  }
}

// Gate non-linearities of the tensor-core path: ex2.approx-based exp and a correctly rounded
// reciprocal (absolute error ~1e-7, an order of magnitude below the 2^-16 operand error of the split
// products feeding them).  The fp32 SIMT kernels keep expf / tanhf.
CLSR_DEVINL float sigm(float x) { return __frcp_rn(1.0f + __expf(-x)); }
CLSR_DEVINL float tanh_(float x) { return 1.0f - 2.0f * __frcp_rn(1.0f + __expf(2.0f * x)); }

CLSR_DEVINL float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
CLSR_DEVINL void st2(float* p, float x, float y) { *reinterpret_cast<float2*>(p) = make_float2(x, y); }

// Per-thread geometry: rows (sequences) g and g+8 of the CTA, units col and col+1.
struct Geo {
  int g, t, col;
  int len[2];        // sequence lengths of the two rows (0 for rows past S)
  size_t rowbase[2]; // (clamped sequence index) * T: global row of step 0
  bool valid[2];
};
CLSR_DEVINL Geo make_geo(const int* slen, int s0, int S, int T) {
  Geo G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  G.g = lane >> 2; G.t = lane & 3; G.col = warp * 8 + 2 * G.t;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int q = G.g + 8 * r, s = s0 + q;
    G.valid[r] = s < S;
    G.len[r] = slen[q];
    G.rowbase[r] = (size_t)(s < S ? s : S - 1) * T;   // loads of dead rows stay in bounds, results unused
  }
  return G;
}

// ------------------------------------------------------------------------------------------------
// GRU forward (clsr.py:161-168, 230-236).  Outputs as gru_fwd_kernel.
template <int U>
__global__ void __launch_bounds__(32 * (U / 8))
gru_fwd_tc_kernel(const float* __restrict__ PX, int ldpx, int colg, int colc, const float* __restrict__ h0,
                  const float* __restrict__ Wgh, const float* __restrict__ Wch, const int* __restrict__ len, int S,
                  int T, float* __restrict__ gates, float* __restrict__ cand, float* __restrict__ hprev,
                  float* __restrict__ rh, float* __restrict__ hfinal) {
  constexpr int NT = 32 * (U / 8), U2 = 2 * U;
  __shared__ ATile<U> hA, rhA;
  __shared__ int slen[RNN_NSEQ];
  const int tid = threadIdx.x, s0 = blockIdx.x * NSEQ;
  hA.zero(tid, NT); rhA.zero(tid, NT);
  const int tmax = rnn_setup_len(len, s0, S, slen);
  const Geo G = make_geo(slen, s0, S, T);
  BFrag<ATile<U>::NKS> br, bu, bc;
  br.load(Wgh, U2, U, (tid >> 5) * 8 + G.g, G.t);
  bu.load(Wgh, U2, U, U + (tid >> 5) * 8 + G.g, G.t);
  bc.load(Wch, U, U, (tid >> 5) * 8 + G.g, G.t);
  float h[4];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float2 v = (h0 && G.valid[r]) ? ld2(h0 + (size_t)(s0 + G.g + 8 * r) * U + G.col) : make_float2(0.f, 0.f);
    h[2 * r] = v.x; h[2 * r + 1] = v.y;
    hA.put(G.g + 8 * r, G.col, v.x, v.y);
  }
  __syncthreads();
  float2 pr[2], pu[2], pc[2];
  auto load_px = [&](int t) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float* base = PX + (G.rowbase[r] + t) * ldpx;
      pr[r] = ld2(base + colg + G.col); pu[r] = ld2(base + colg + U + G.col); pc[r] = ld2(base + colc + G.col);
    }
  };
  if (tmax > 0) load_px(0);
  const int pf_lines = (3 * U + 31) / 32 + 1;
  for (int t = 0; t < tmax; ++t) {
    if (tid < NSEQ * pf_lines) prefetch_row_lines(PX, ldpx, colg, pf_lines, tid, s0, T, t + 3, slen);
    float ar[4] = {pr[0].x, pr[0].y, pr[1].x, pr[1].y}, au[4] = {pu[0].x, pu[0].y, pu[1].x, pu[1].y};
    float ac[4] = {pc[0].x, pc[0].y, pc[1].x, pc[1].y};
    if (t + 1 < tmax) load_px(t + 1);
    matmul<U>(ar, hA, br, G.g, G.t);
    matmul<U>(au, hA, bu, G.g, G.t);
    float u[4], rhv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float r = sigm(ar[e]);
      u[e] = sigm(au[e]);
      ar[e] = r;
      rhv[e] = r * h[e];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const bool live = t < G.len[r];
      rhA.put(G.g + 8 * r, G.col, live ? rhv[2 * r] : 0.f, live ? rhv[2 * r + 1] : 0.f);
      if (live) {
        const size_t row = G.rowbase[r] + t;
        st2(gates + row * U2 + G.col, ar[2 * r], ar[2 * r + 1]);
        st2(gates + row * U2 + U + G.col, u[2 * r], u[2 * r + 1]);
        st2(rh + row * U + G.col, rhv[2 * r], rhv[2 * r + 1]);
        st2(hprev + row * U + G.col, h[2 * r], h[2 * r + 1]);
      }
    }
    __syncthreads();
    matmul<U>(ac, rhA, bc, G.g, G.t);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (t < G.len[r]) {
        const float c0 = tanh_(ac[2 * r]), c1 = tanh_(ac[2 * r + 1]);
        st2(cand + (G.rowbase[r] + t) * U + G.col, c0, c1);
        h[2 * r] = u[2 * r] * h[2 * r] + (1.f - u[2 * r]) * c0;
        h[2 * r + 1] = u[2 * r + 1] * h[2 * r + 1] + (1.f - u[2 * r + 1]) * c1;
      }
      hA.put(G.g + 8 * r, G.col, h[2 * r], h[2 * r + 1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r)
    if (G.valid[r]) st2(hfinal + (size_t)(s0 + G.g + 8 * r) * U + G.col, h[2 * r], h[2 * r + 1]);
}

// BPTT of the GRU.  Outputs as gru_bwd_kernel (dead positions of dPX must be pre-zeroed).
template <int U>
__global__ void __launch_bounds__(32 * (U / 8))
gru_bwd_tc_kernel(const float* __restrict__ gates, const float* __restrict__ cand, const float* __restrict__ hprev,
                  const float* __restrict__ WghT, const float* __restrict__ WchT, const float* __restrict__ dh_final,
                  const int* __restrict__ len, int S, int T, float* __restrict__ dPX, int ldpx, int colg, int colc,
                  float* __restrict__ dh0) {
  constexpr int NT = 32 * (U / 8), U2 = 2 * U;
  __shared__ ATile<U> dcA[2];     // d pre-activation of the candidate
  __shared__ ATile<U2> dgA[2];    // d pre-activation of [r, u]
  __shared__ int slen[RNN_NSEQ];
  const int tid = threadIdx.x, s0 = blockIdx.x * NSEQ;
  dcA[0].zero(tid, NT); dcA[1].zero(tid, NT); dgA[0].zero(tid, NT); dgA[1].zero(tid, NT);
  const int tmax = rnn_setup_len(len, s0, S, slen);
  const Geo G = make_geo(slen, s0, S, T);
  BFrag<ATile<U>::NKS> bc;
  BFrag<ATile<U2>::NKS> bg;
  bc.load(WchT, U, U, (tid >> 5) * 8 + G.g, G.t);
  bg.load(WghT, U, U2, (tid >> 5) * 8 + G.g, G.t);
  float dh[4];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float2 v = G.valid[r] ? ld2(dh_final + (size_t)(s0 + G.g + 8 * r) * U + G.col) : make_float2(0.f, 0.f);
    dh[2 * r] = v.x; dh[2 * r + 1] = v.y;
  }
  float2 vr[2], vu[2], vc[2], vh[2];
  auto load_in = [&](int t) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const size_t row = G.rowbase[r] + t;
      vr[r] = ld2(gates + row * U2 + G.col); vu[r] = ld2(gates + row * U2 + U + G.col);
      vc[r] = ld2(cand + row * U + G.col); vh[r] = ld2(hprev + row * U + G.col);
    }
  };
  if (tmax > 0) load_in(tmax - 1);
  const int lg = (U2 + 31) / 32 + 1, lc = (U + 31) / 32 + 1;
  for (int t = tmax - 1; t >= 0; --t) {
    for (int i0 = tid; i0 < NSEQ * (lg + 2 * lc); i0 += NT) {
      int idx = i0;
      if (idx < NSEQ * lg) prefetch_row_lines(gates, U2, 0, lg, idx, s0, T, t - 3, slen);
      else if ((idx -= NSEQ * lg) < NSEQ * lc) prefetch_row_lines(cand, U, 0, lc, idx, s0, T, t - 3, slen);
      else if ((idx -= NSEQ * lc) < NSEQ * lc) prefetch_row_lines(hprev, U, 0, lc, idx, s0, T, t - 3, slen);
    }
    const int par = t & 1;
    const float rg[4] = {vr[0].x, vr[0].y, vr[1].x, vr[1].y}, ug[4] = {vu[0].x, vu[0].y, vu[1].x, vu[1].y};
    const float cg[4] = {vc[0].x, vc[0].y, vc[1].x, vc[1].y}, hp[4] = {vh[0].x, vh[0].y, vh[1].x, vh[1].y};
    if (t > 0) load_in(t - 1);
    float dha[4], dpc[4], dpu[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool live = t < G.len[e >> 1];
      const float dhn = dh[e];
      const float du = dhn * (hp[e] - cg[e]), dc = dhn * (1.f - ug[e]);
      dpc[e] = live ? dc * (1.f - cg[e] * cg[e]) : 0.f;
      dpu[e] = live ? du * ug[e] * (1.f - ug[e]) : 0.f;
      dha[e] = live ? dhn * ug[e] : dhn;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      dcA[par].put(G.g + 8 * r, G.col, dpc[2 * r], dpc[2 * r + 1]);
      dgA[par].put(G.g + 8 * r, U + G.col, dpu[2 * r], dpu[2 * r + 1]);
      if (t < G.len[r]) {
        float* dp = dPX + (G.rowbase[r] + t) * ldpx;
        st2(dp + colc + G.col, dpc[2 * r], dpc[2 * r + 1]);
        st2(dp + colg + U + G.col, dpu[2 * r], dpu[2 * r + 1]);
      }
    }
    __syncthreads();
    float a1[4] = {0.f, 0.f, 0.f, 0.f};
    matmul<U>(a1, dcA[par], bc, G.g, G.t);
    float dpr[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool live = t < G.len[e >> 1];
      const float dr = a1[e] * hp[e];
      dha[e] += live ? a1[e] * rg[e] : 0.f;
      dpr[e] = live ? dr * rg[e] * (1.f - rg[e]) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      dgA[par].put(G.g + 8 * r, G.col, dpr[2 * r], dpr[2 * r + 1]);
      if (t < G.len[r]) st2(dPX + (G.rowbase[r] + t) * ldpx + colg + G.col, dpr[2 * r], dpr[2 * r + 1]);
    }
    __syncthreads();
    float a2[4] = {0.f, 0.f, 0.f, 0.f};
    matmul<U2>(a2, dgA[par], bg, G.g, G.t);
#pragma unroll
    for (int e = 0; e < 4; ++e) dh[e] = dha[e] + a2[e];
  }
  if (dh0) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
      if (G.valid[r]) st2(dh0 + (size_t)(s0 + G.g + 8 * r) * U + G.col, dh[2 * r], dh[2 * r + 1]);
  }
}

// ------------------------------------------------------------------------------------------------
// Time4LSTM forward (rnn_cell_implement.py:129-298).  Outputs as lstm_fwd_kernel.
template <int H>
__global__ void __launch_bounds__(32 * (H / 8))
lstm_fwd_tc_kernel(const float* __restrict__ PX, int ldpx, int colL, int colTN, int colTL, const float* __restrict__ Km,
                   const int* __restrict__ len, int S, int T, float* __restrict__ gates4, float* __restrict__ cprev,
                   float* __restrict__ mprev, float* __restrict__ R) {
  constexpr int NT = 32 * (H / 8), H4 = 4 * H;
  __shared__ ATile<H> mA[2];
  __shared__ int slen[RNN_NSEQ];
  const int tid = threadIdx.x, s0 = blockIdx.x * NSEQ;
  mA[0].zero(tid, NT); mA[1].zero(tid, NT);
  const int tmax = rnn_setup_len(len, s0, S, slen);
  const Geo G = make_geo(slen, s0, S, T);
  BFrag<ATile<H>::NKS> bk[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) bk[q].load(Km, H4, H, q * H + (tid >> 5) * 8 + G.g, G.t);
  float c[4] = {0.f, 0.f, 0.f, 0.f}, m[4] = {0.f, 0.f, 0.f, 0.f};
  float2 pg[4][2], ptn[2], ptl[2];
  auto load_px = [&](int t) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float* base = PX + (G.rowbase[r] + t) * ldpx;
#pragma unroll
      for (int q = 0; q < 4; ++q) pg[q][r] = ld2(base + colL + q * H + G.col);
      ptn[r] = ld2(base + colTN + G.col); ptl[r] = ld2(base + colTL + G.col);
    }
  };
  if (tmax > 0) load_px(0);
  const int pf_lines = (6 * H + 31) / 32 + 1;
  for (int t = 0; t < tmax; ++t) {
    if (tid < NSEQ * pf_lines) prefetch_row_lines(PX, ldpx, colL, pf_lines, tid, s0, T, t + 3, slen);
    float acc[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { acc[q][0] = pg[q][0].x; acc[q][1] = pg[q][0].y; acc[q][2] = pg[q][1].x; acc[q][3] = pg[q][1].y; }
    const float tn[4] = {ptn[0].x, ptn[0].y, ptn[1].x, ptn[1].y}, tl[4] = {ptl[0].x, ptl[0].y, ptl[1].x, ptl[1].y};
    if (t + 1 < tmax) load_px(t + 1);
    const ATile<H>& cur = mA[t & 1];
#pragma unroll
    for (int q = 0; q < 4; ++q) matmul<H>(acc[q], cur, bk[q], G.g, G.t);
    float gi[4], gj[4], gf[4], go[4], cn[4], mn[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      gi[e] = sigm(acc[0][e]);
      gj[e] = tanh_(acc[1][e]);
      gf[e] = sigm(acc[2][e] + 1.0f);
      go[e] = sigm(acc[3][e]);
      const float sn = sigm(tn[e]), sl = sigm(tl[e]);
      cn[e] = gf[e] * sl * c[e] + gi[e] * sn * gj[e];
      mn[e] = go[e] * tanh_(cn[e]);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const size_t row = G.rowbase[r] + t;
      if (t < G.len[r]) {
        st2(gates4 + row * H4 + G.col, gi[2 * r], gi[2 * r + 1]);
        st2(gates4 + row * H4 + H + G.col, gj[2 * r], gj[2 * r + 1]);
        st2(gates4 + row * H4 + 2 * H + G.col, gf[2 * r], gf[2 * r + 1]);
        st2(gates4 + row * H4 + 3 * H + G.col, go[2 * r], go[2 * r + 1]);
        st2(cprev + row * H + G.col, c[2 * r], c[2 * r + 1]);
        st2(mprev + row * H + G.col, m[2 * r], m[2 * r + 1]);
        st2(R + row * H + G.col, mn[2 * r], mn[2 * r + 1]);
        c[2 * r] = cn[2 * r]; c[2 * r + 1] = cn[2 * r + 1];
        m[2 * r] = mn[2 * r]; m[2 * r + 1] = mn[2 * r + 1];
      } else if (G.valid[r]) {
        st2(R + row * H + G.col, 0.f, 0.f);
      }
      mA[(t + 1) & 1].put(G.g + 8 * r, G.col, m[2 * r], m[2 * r + 1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r)
    if (G.valid[r])
      for (int t = tmax; t < T; ++t) st2(R + (G.rowbase[r] + t) * H + G.col, 0.f, 0.f);
}

// BPTT of the Time4LSTM.  Outputs as lstm_bwd_kernel.
template <int H>
__global__ void __launch_bounds__(32 * (H / 8))
lstm_bwd_tc_kernel(const float* __restrict__ PX, int ldpx, int colL, int colTN, int colTL,
                   const float* __restrict__ gates4, const float* __restrict__ cprev, const float* __restrict__ KmT,
                   const float* __restrict__ dR, const int* __restrict__ len, int S, int T, float* __restrict__ dPX) {
  constexpr int NT = 32 * (H / 8), H4 = 4 * H;
  __shared__ ATile<H4> dpA[2];
  __shared__ int slen[RNN_NSEQ];
  const int tid = threadIdx.x, s0 = blockIdx.x * NSEQ;
  dpA[0].zero(tid, NT); dpA[1].zero(tid, NT);
  const int tmax = rnn_setup_len(len, s0, S, slen);
  const Geo G = make_geo(slen, s0, S, T);
  BFrag<ATile<H4>::NKS> bk;
  bk.load(KmT, H, H4, (tid >> 5) * 8 + G.g, G.t);
  float dm[4] = {0.f, 0.f, 0.f, 0.f}, dc[4] = {0.f, 0.f, 0.f, 0.f};
  float2 vg[4][2], vcp[2], vtn[2], vtl[2], vdr[2];
  auto load_in = [&](int t) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const size_t row = G.rowbase[r] + t;
#pragma unroll
      for (int q = 0; q < 4; ++q) vg[q][r] = ld2(gates4 + row * H4 + q * H + G.col);
      vcp[r] = ld2(cprev + row * H + G.col);
      vtn[r] = ld2(PX + row * ldpx + colTN + G.col); vtl[r] = ld2(PX + row * ldpx + colTL + G.col);
      vdr[r] = ld2(dR + row * H + G.col);
    }
  };
  if (tmax > 0) load_in(tmax - 1);
  const int l4 = (H4 + 31) / 32 + 1, l1 = (H + 31) / 32 + 1, l2 = (2 * H + 31) / 32 + 1;
  for (int t = tmax - 1; t >= 0; --t) {
    for (int i0 = tid; i0 < NSEQ * (l4 + 2 * l1 + l2); i0 += NT) {
      int idx = i0;
      if (idx < NSEQ * l4) prefetch_row_lines(gates4, H4, 0, l4, idx, s0, T, t - 3, slen);
      else if ((idx -= NSEQ * l4) < NSEQ * l1) prefetch_row_lines(cprev, H, 0, l1, idx, s0, T, t - 3, slen);
      else if ((idx -= NSEQ * l1) < NSEQ * l1) prefetch_row_lines(dR, H, 0, l1, idx, s0, T, t - 3, slen);
      else if ((idx -= NSEQ * l1) < NSEQ * l2) prefetch_row_lines(PX, ldpx, colTN, l2, idx, s0, T, t - 3, slen);
    }
    const int par = t & 1;
    float gi[4], gj[4], gf[4], go[4], cp[4], tn[4], tl[4], dr[4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      gi[2 * r] = vg[0][r].x; gi[2 * r + 1] = vg[0][r].y; gj[2 * r] = vg[1][r].x; gj[2 * r + 1] = vg[1][r].y;
      gf[2 * r] = vg[2][r].x; gf[2 * r + 1] = vg[2][r].y; go[2 * r] = vg[3][r].x; go[2 * r + 1] = vg[3][r].y;
      cp[2 * r] = vcp[r].x; cp[2 * r + 1] = vcp[r].y; tn[2 * r] = vtn[r].x; tn[2 * r + 1] = vtn[r].y;
      tl[2 * r] = vtl[r].x; tl[2 * r + 1] = vtl[r].y; dr[2 * r] = vdr[r].x; dr[2 * r + 1] = vdr[r].y;
    }
    if (t > 0) load_in(t - 1);
    float dpi[4], dpj[4], dpf[4], dpo[4], dsn[4], dsl[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool live = t < G.len[e >> 1];
      const float sn = sigm(tn[e]), sl = sigm(tl[e]);
      const float cn = gf[e] * sl * cp[e] + gi[e] * sn * gj[e];
      const float tc = tanh_(cn);
      const float dmt = dm[e] + dr[e];
      const float dcn = dc[e] + dmt * go[e] * (1.f - tc * tc);
      dpo[e] = live ? dmt * tc * go[e] * (1.f - go[e]) : 0.f;
      dpf[e] = live ? dcn * sl * cp[e] * gf[e] * (1.f - gf[e]) : 0.f;
      dsl[e] = dcn * gf[e] * cp[e] * sl * (1.f - sl);
      dpi[e] = live ? dcn * sn * gj[e] * gi[e] * (1.f - gi[e]) : 0.f;
      dsn[e] = dcn * gi[e] * gj[e] * sn * (1.f - sn);
      dpj[e] = live ? dcn * gi[e] * sn * (1.f - gj[e] * gj[e]) : 0.f;
      if (live) dc[e] = dcn * gf[e] * sl;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int q = G.g + 8 * r;
      dpA[par].put(q, G.col, dpi[2 * r], dpi[2 * r + 1]);
      dpA[par].put(q, H + G.col, dpj[2 * r], dpj[2 * r + 1]);
      dpA[par].put(q, 2 * H + G.col, dpf[2 * r], dpf[2 * r + 1]);
      dpA[par].put(q, 3 * H + G.col, dpo[2 * r], dpo[2 * r + 1]);
      if (t < G.len[r]) {
        float* dp = dPX + (G.rowbase[r] + t) * ldpx;
        st2(dp + colL + G.col, dpi[2 * r], dpi[2 * r + 1]);
        st2(dp + colL + H + G.col, dpj[2 * r], dpj[2 * r + 1]);
        st2(dp + colL + 2 * H + G.col, dpf[2 * r], dpf[2 * r + 1]);
        st2(dp + colL + 3 * H + G.col, dpo[2 * r], dpo[2 * r + 1]);
        st2(dp + colTN + G.col, dsn[2 * r], dsn[2 * r + 1]);
        st2(dp + colTL + G.col, dsl[2 * r], dsl[2 * r + 1]);
      }
    }
    __syncthreads();
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    matmul<H4>(a, dpA[par], bk, G.g, G.t);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (t < G.len[e >> 1]) dm[e] = a[e];
  }
}

}  // namespace rtc
}  // namespace clsr
