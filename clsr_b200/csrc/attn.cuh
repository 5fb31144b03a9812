// Attention pooling (score unit + masked softmax + weighted sum, clsr.py:371-381,154,221), the
// history proxies hist_mean / hist_recent (clsr.py:157,173-177), and the group-structured
// backward helpers of the factorised first attention layer.
#pragma once
#include "common.cuh"

namespace clsr {

// One warp per row.  score[t] = relu(bn(h1[row,t,:])) . wout + bout for t < len, masked softmax
// over t, att[row,:] = sum_t w[t] * V[seq,t,:].  With hm/hr non-null also the two proxies.
__global__ void __launch_bounds__(128)
pool_fwd_kernel(const float* __restrict__ h1, int A1, const float* __restrict__ scale,
                const float* __restrict__ shift, const float* __restrict__ wout,
                const float* __restrict__ bout, const float* __restrict__ V, int Dv,
                const int* __restrict__ len, int rows, int T, int G, float* __restrict__ w_out,
                float* __restrict__ att, float* __restrict__ hm, float* __restrict__ hr, int recent_k) {
  extern __shared__ __align__(16) float sm[];
  float* s_scale = sm;
  float* s_shift = s_scale + A1;
  float* s_w = s_shift + A1;
  float* s_sc = s_w + A1;  // [warps][T]
  for (int i = threadIdx.x; i < A1; i += blockDim.x) {
    s_scale[i] = scale[i]; s_shift[i] = shift[i]; s_w[i] = wout[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* sc = s_sc + wid * T;
  const float b0 = bout[0];
  const bool vec4 = (A1 & 3) == 0 && (reinterpret_cast<uintptr_t>(h1) & 15) == 0;
  for (int row = blockIdx.x * wpb + wid; row < rows; row += gridDim.x * wpb) {
    const int s = row / G;
    const int L = len[s];
    float mx = -INFINITY;
    for (int t = lane; t < L; t += 32) {
      const float* hp = h1 + ((size_t)row * T + t) * A1;
      float dot = b0;
      if (vec4) {
        // one 16-byte load per four channels (the lane's 4*A1-byte row is contiguous)
        for (int n = 0; n < A1; n += 4) {
          const float4 h = __ldg(reinterpret_cast<const float4*>(hp + n));
          const float4 sc4 = *reinterpret_cast<const float4*>(s_scale + n), sh4 = *reinterpret_cast<const float4*>(s_shift + n);
          const float4 w4 = *reinterpret_cast<const float4*>(s_w + n);
          dot = fmaf(fmaxf(0.f, fmaf(h.x, sc4.x, sh4.x)), w4.x, dot);
          dot = fmaf(fmaxf(0.f, fmaf(h.y, sc4.y, sh4.y)), w4.y, dot);
          dot = fmaf(fmaxf(0.f, fmaf(h.z, sc4.z, sh4.z)), w4.z, dot);
          dot = fmaf(fmaxf(0.f, fmaf(h.w, sc4.w, sh4.w)), w4.w, dot);
        }
      } else {
        for (int n = 0; n < A1; ++n) dot = fmaf(fmaxf(0.f, fmaf(hp[n], s_scale[n], s_shift[n])), s_w[n], dot);
      }
      sc[t] = dot;
      mx = fmaxf(mx, dot);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int t = lane; t < L; t += 32) {
      float e = expf(sc[t] - mx);
      sc[t] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    for (int t = lane; t < T; t += 32) {
      float w = (t < L) ? sc[t] / sum : 0.f;
      sc[t] = w;
      w_out[(size_t)row * T + t] = w;
    }
    __syncwarp();
    const int kk = min(L, recent_k);
    for (int d = lane; d < Dv; d += 32) {
      float acc = 0.f, am = 0.f, ar = 0.f;
      const float* vp = V + (size_t)s * T * Dv + d;
      for (int t = 0; t < L; ++t) {
        float v = vp[(size_t)t * Dv];
        acc = fmaf(sc[t], v, acc);
        am += v;
        if (t >= L - kk) ar += v;
      }
      att[(size_t)row * Dv + d] = acc;
      if (hm) {
        hm[(size_t)s * Dv + d] = am / (float)L;
        hr[(size_t)s * Dv + d] = ar / (float)kk;
      }
    }
    __syncwarp();
  }
}

// Backward of pool_fwd, one warp per group of G rows sharing the values V[seq].
//   dsc[t]  = w[t] * (datt.V[t] - sum_t' w[t'] datt.V[t'])
//   dy1[row,t,n] = dsc[t]*wout[n] where bn(h1)>0 (zero elsewhere and on padded positions)
//   stat += (sum dy1, sum dy1*xhat) per channel;  dwout += sum relu(bn(h1))*dsc;  dbout += sum dsc
//   dV[seq,t,:] (=|+=) sum_g w[g,t] * datt[g,:]  (+ proxy gradients dhm/L, dhr/k when given)
__global__ void __launch_bounds__(128, 8)   // <= 64 registers: 8 CTAs (32 warps) per SM hide the h1 / V load latency
pool_bwd_kernel(const float* __restrict__ datt, const float* __restrict__ w, const float* __restrict__ V,
                int Dv, const float* __restrict__ h1, int A1, const float* __restrict__ scale,
                const float* __restrict__ shift, const float* __restrict__ mean,
                const float* __restrict__ rstd, const float* __restrict__ wout,
                const int* __restrict__ len, int nseq, int T, int G, int vdiv, float* __restrict__ dy1,
                double* __restrict__ stat, float* __restrict__ dwout, float* __restrict__ dbout,
                float* __restrict__ dV, int dV_accum, const float* __restrict__ dhm,
                const float* __restrict__ dhr, int recent_k) {
  // vdiv > 1: every "sequence" index here is a single row (G == 1) whose values / length belong to
  // sequence s / vdiv; dV is then produced by pool_dv_kernel (dV == nullptr here).
  extern __shared__ __align__(16) float sm[];
  const int wpb = blockDim.x >> 5;
  float* s_scale = sm;
  float* s_shift = s_scale + A1;
  float* s_mean = s_shift + A1;
  float* s_rstd = s_mean + A1;
  float* s_w = s_rstd + A1;
  float* s_acc = s_w + A1;            // [3][A1] + 1 : stat1, stat2, dwout, dbout
  float* s_warp = s_acc + ((3 * A1 + 1 + 3) & ~3);  // per warp: dsc[T], wrow[T], dat[Dv] (16-byte aligned when T is even)
  for (int i = threadIdx.x; i < A1; i += blockDim.x) {
    s_scale[i] = scale[i]; s_shift[i] = shift[i]; s_mean[i] = mean[i]; s_rstd[i] = rstd[i]; s_w[i] = wout[i];
  }
  for (int i = threadIdx.x; i < 3 * A1 + 1; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* dsc = s_warp + wid * (2 * T + Dv);
  float* wrow = dsc + T;
  float* dat = wrow + T;
  constexpr int MAXSLOT = 4;  // A1 <= 128
  float st1[MAXSLOT], st2[MAXSLOT], dwo[MAXSLOT], dbo = 0.f;
#pragma unroll
  for (int i = 0; i < MAXSLOT; ++i) { st1[i] = 0.f; st2[i] = 0.f; dwo[i] = 0.f; }
  // quad mapping of the dy1 pass (A1 multiple of 4, at most 32 quads)
  const int nq = A1 >> 2;
  const bool quad = (A1 & 3) == 0 && nq <= 32 && ((reinterpret_cast<uintptr_t>(h1) | reinterpret_cast<uintptr_t>(dy1)) & 15) == 0;
  const int ntl = quad ? 32 / nq : 1;
  const int n4 = quad ? lane % nq : 0, tl = quad ? lane / nq : 0;
  // values V / datt rows addressable with 16-byte vectors (dat sits at a 16-byte aligned shared offset)
  const bool vvec = (Dv & 3) == 0 && (Dv >> 2) <= 32 && (T & 1) == 0 && (A1 & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(V) & 15) == 0;
  float q_sc[4], q_sh[4], q_mn[4], q_rs[4], q_wo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = quad ? n4 * 4 + i : 0;
    q_sc[i] = s_scale[n]; q_sh[i] = s_shift[n]; q_mn[i] = s_mean[n]; q_rs[i] = s_rstd[n]; q_wo[i] = s_w[n];
  }

  for (int s = blockIdx.x * wpb + wid; s < nseq; s += gridDim.x * wpb) {
    const int sv = s / vdiv;
    const int L = len[sv];
    const int kk = min(L, recent_k);
    for (int g = 0; g < G; ++g) {
      const size_t row = (size_t)s * G + g;
      for (int d = lane; d < Dv; d += 32) dat[d] = datt[row * Dv + d];
      __syncwarp();
      float dot = 0.f;
      for (int t = lane; t < T; t += 32) {
        float dwt = 0.f, wt = 0.f;
        if (t < L) {
          const float* vp = V + ((size_t)sv * T + t) * Dv;
          if (vvec) {
            for (int d = 0; d < Dv; d += 4) {
              const float4 x = __ldg(reinterpret_cast<const float4*>(vp + d));
              const float4 a = *reinterpret_cast<const float4*>(dat + d);
              dwt = fmaf(a.x, x.x, fmaf(a.y, x.y, fmaf(a.z, x.z, fmaf(a.w, x.w, dwt))));
            }
          } else {
            for (int d = 0; d < Dv; ++d) dwt = fmaf(dat[d], vp[d], dwt);
          }
          wt = w[row * T + t];
          dot = fmaf(wt, dwt, dot);
        }
        dsc[t] = dwt;
        wrow[t] = wt;
      }
      dot = warp_sum(dot);
      for (int t = lane; t < T; t += 32) {
        float v = wrow[t] * (dsc[t] - dot);
        dsc[t] = v;
        dbo += v;
      }
      __syncwarp();
      if (quad) {
        // lane = (position lane tl, channel quad n4): 16-byte loads / stores, nq * ntl lanes cover ntl
        // consecutive positions (contiguous 4*A1*ntl bytes) per iteration; statistics stay per lane
        if (lane < nq * ntl) {
          const float4* hp4 = reinterpret_cast<const float4*>(h1 + row * T * A1) + n4;
          float4* dp4 = reinterpret_cast<float4*>(dy1 + row * T * A1) + n4;
#pragma unroll 4
          for (int t = tl; t < T; t += ntl) {
            const float4 h = __ldg(hp4 + (size_t)t * nq);
            const float d = dsc[t];
            const float hh[4] = {h.x, h.y, h.z, h.w};
            float vv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float y = fmaf(hh[i], q_sc[i], q_sh[i]);
              const float v = (y > 0.f) ? d * q_wo[i] : 0.f;
              vv[i] = v;
              st1[i] += v;
              st2[i] = fmaf(v, (hh[i] - q_mn[i]) * q_rs[i], st2[i]);
              dwo[i] = fmaf(fmaxf(y, 0.f), d, dwo[i]);
            }
            dp4[(size_t)t * nq] = make_float4(vv[0], vv[1], vv[2], vv[3]);
          }
        }
      } else {
#pragma unroll
      for (int sl = 0; sl < MAXSLOT; ++sl) {
        const int n = lane + 32 * sl;
        if (n < A1) {
          const float scn = s_scale[n], shn = s_shift[n], mn = s_mean[n], rs = s_rstd[n], wo = s_w[n];
          const float* hp = h1 + row * T * A1 + n;
          float* dp = dy1 + row * T * A1 + n;
          for (int t = 0; t < T; ++t) {
            float h = hp[(size_t)t * A1];
            float y = fmaf(h, scn, shn);
            float d = dsc[t];
            float v = (y > 0.f) ? d * wo : 0.f;
            dp[(size_t)t * A1] = v;
            st1[sl] += v;
            st2[sl] = fmaf(v, (h - mn) * rs, st2[sl]);
            dwo[sl] = fmaf(fmaxf(y, 0.f), d, dwo[sl]);
          }
        }
      }
      }
      if (dV && vvec && ((reinterpret_cast<uintptr_t>(dV) & 15) == 0)) {
        // lane = (position lane, column quad): 16-byte read-modify-writes of dV, nqv * ntlv lanes active
        const int nqv = Dv >> 2, ntlv = 32 / nqv;
        if (lane < nqv * ntlv) {
          const int d4 = (lane % nqv) * 4, tl2 = lane / nqv;
          const float4 da = *reinterpret_cast<const float4*>(dat + d4);
          float4 pm = make_float4(0.f, 0.f, 0.f, 0.f), pr = pm;
          if (dhm) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(dhm + (size_t)s * Dv + d4));
            const float inv = 1.f / (float)L;
            pm = make_float4(t4.x * inv, t4.y * inv, t4.z * inv, t4.w * inv);
          }
          if (dhr) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(dhr + (size_t)s * Dv + d4));
            const float inv = 1.f / (float)kk;
            pr = make_float4(t4.x * inv, t4.y * inv, t4.z * inv, t4.w * inv);
          }
          float* op = dV + (size_t)s * T * Dv + d4;
          for (int t = tl2; t < T; t += ntlv) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < L) {
              const float wt = wrow[t];
              v = make_float4(wt * da.x, wt * da.y, wt * da.z, wt * da.w);
              if (g == 0) {
                v.x += pm.x; v.y += pm.y; v.z += pm.z; v.w += pm.w;
                if (t >= L - kk) { v.x += pr.x; v.y += pr.y; v.z += pr.z; v.w += pr.w; }
              }
            }
            float4* o4 = reinterpret_cast<float4*>(op + (size_t)t * Dv);
            if (g == 0 && !dV_accum) *o4 = v;
            else if (t < L) {
              float4 o = *o4;
              o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
              *o4 = o;
            }
          }
        }
      } else
      for (int d = lane; dV && d < Dv; d += 32) {
        const float da = dat[d];
        const float pm = dhm ? dhm[(size_t)s * Dv + d] / (float)L : 0.f;
        const float pr = dhr ? dhr[(size_t)s * Dv + d] / (float)kk : 0.f;
        float* op = dV + (size_t)s * T * Dv + d;
        for (int t = 0; t < T; ++t) {
          float v = 0.f;
          if (t < L) {
            v = wrow[t] * da;
            if (g == 0) { v += pm; if (t >= L - kk) v += pr; }
          }
          if (g == 0 && !dV_accum) op[(size_t)t * Dv] = v;
          else if (t < L) op[(size_t)t * Dv] += v;
        }
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int sl = 0; sl < MAXSLOT; ++sl) {
    const int n = quad ? n4 * 4 + sl : lane + 32 * sl;
    if (n < A1 && (!quad || lane < nq * ntl)) {
      atomicAdd(&s_acc[n], st1[sl]); atomicAdd(&s_acc[A1 + n], st2[sl]); atomicAdd(&s_acc[2 * A1 + n], dwo[sl]);
    }
  }
  dbo = warp_sum(dbo);
  if (lane == 0) atomicAdd(&s_acc[3 * A1], dbo);
  __syncthreads();
  for (int i = threadIdx.x; i < A1; i += blockDim.x) {
    atomicAdd(stat + i, (double)s_acc[i]);
    atomicAdd(stat + A1 + i, (double)s_acc[A1 + i]);
    atomicAdd(dwout + i, s_acc[2 * A1 + i]);
  }
  if (threadIdx.x == 0) atomicAdd(dbout, s_acc[3 * A1]);
}

// dV[s,t,:] = sum_g w[(s*G+g), t] * datt[(s*G+g), :]  for t < len[s], zero beyond (values gradient of
// the short-term attention, summed over the rows of a group).  One thread per (s, t, d).
__global__ void pool_dv_kernel(const float* __restrict__ w, const float* __restrict__ datt,
                               const int* __restrict__ len, int S, int T, int G, int Dv, float* __restrict__ dV) {
  long long n = (long long)S * T * Dv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int d = (int)(i % Dv);
    long long st = i / Dv;
    int t = (int)(st % T), s = (int)(st / T);
    float acc = 0.f;
    if (t < len[s]) {
      for (int g = 0; g < G; ++g) {
        size_t b = (size_t)s * G + g;
        acc = fmaf(w[b * T + t], datt[b * Dv + d], acc);
      }
    }
    dV[i] = acc;
  }
}

// 16-byte form (Dv multiple of 4): one thread per (s, t, column quad).
__global__ void pool_dv_v4_kernel(const float* __restrict__ w, const float* __restrict__ datt,
                                  const int* __restrict__ len, int S, int T, int G, int Dv, float* __restrict__ dV) {
  const int Q4 = Dv >> 2;
  const long long n = (long long)S * T * Q4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % Q4);
    const long long st = i / Q4;
    const int t = (int)(st % T), s = (int)(st / T);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < len[s]) {
      for (int g = 0; g < G; ++g) {
        const size_t b = (size_t)s * G + g;
        const float wt = __ldg(w + b * T + t);
        const float4 d = __ldg(reinterpret_cast<const float4*>(datt + b * Dv) + q);
        acc.x = fmaf(wt, d.x, acc.x); acc.y = fmaf(wt, d.y, acc.y); acc.z = fmaf(wt, d.z, acc.z); acc.w = fmaf(wt, d.w, acc.w);
      }
    }
    reinterpret_cast<float4*>(dV)[i] = acc;
  }
}

// dh0 = al*dy0 + be*h0 + ga (BatchNorm backward of the first attention layer), reduced two ways:
//   dinv[s,t,n] = sum_g dh0[(s*G+g), t, n]      (skipped when dinv == nullptr)
//   dqb[b,n]    = sum_t dh0[b, t, n]
// One CTA per sequence; blockDim = (NX >= A0, NY).
__global__ void h0_reduce_kernel(const float* __restrict__ dy0, const float* __restrict__ h0, int A0,
                                 const float* __restrict__ al, const float* __restrict__ be,
                                 const float* __restrict__ ga, int T, int G, float* __restrict__ dinv,
                                 float* __restrict__ dqb) {
  extern __shared__ float sq[];  // [NY][A0]
  const int s = blockIdx.x;
  const int n = threadIdx.x, ty = threadIdx.y, ny = blockDim.y;
  const bool act = n < A0;
  const float a = act ? al[n] : 0.f, b = act ? be[n] : 0.f, c = act ? ga[n] : 0.f;
  for (int g = 0; g < G; ++g) {
    const size_t row = (size_t)s * G + g;
    float accq = 0.f;
    if (act) {
      for (int t = ty; t < T; t += ny) {
        size_t o = (row * T + t) * A0 + n;
        float v = fmaf(a, dy0[o], fmaf(b, h0[o], c));
        accq += v;
        if (dinv) {
          size_t oi = ((size_t)s * T + t) * A0 + n;
          if (g == 0) dinv[oi] = v; else dinv[oi] += v;
        }
      }
      sq[ty * A0 + n] = accq;
    }
    __syncthreads();
    if (act && ty == 0) {
      float tot = 0.f;
      for (int y = 0; y < ny; ++y) tot += sq[y * A0 + n];
      dqb[row * A0 + n] = tot;
    }
    __syncthreads();
  }
}

// Backward of the per-target product P[b,t,k] = a2[s,t,k] * tgt[b,k]:
//   da2[s,t,k] = sum_g dP[(s*G+g),t,k] * tgt[(s*G+g),k];   dtgt[b,k] += sum_t dP[b,t,k] * a2[s,t,k]
// a2 = as[:, off:off+D].  One CTA per sequence; blockDim = (NX >= D, NY).
__global__ void mulrow_bwd_kernel(const float* __restrict__ dP, int D, const float* __restrict__ as, int lda,
                                  int off, const float* __restrict__ tgt, int ldt, int T, int G,
                                  float* __restrict__ da2, float* __restrict__ dtgt, int lddt) {
  extern __shared__ float sq[];  // [NY][D]
  const int s = blockIdx.x;
  const int k = threadIdx.x, ty = threadIdx.y, ny = blockDim.y;
  const bool act = k < D;
  for (int g = 0; g < G; ++g) {
    const size_t row = (size_t)s * G + g;
    float acct = 0.f;
    if (act) {
      const float tg = tgt[row * ldt + k];
      for (int t = ty; t < T; t += ny) {
        float d = dP[(row * T + t) * D + k];
        size_t m = (size_t)s * T + t;
        acct = fmaf(d, as[m * lda + off + k], acct);
        float v = d * tg;
        if (g == 0) da2[m * D + k] = v; else da2[m * D + k] += v;
      }
      sq[ty * D + k] = acct;
    }
    __syncthreads();
    if (act && ty == 0) {
      float tot = 0.f;
      for (int y = 0; y < ny; ++y) tot += sq[y * D + k];
      dtgt[row * lddt + k] += tot;
    }
    __syncthreads();
  }
}

// 16-byte vectorised forms of h0_reduce_kernel / mulrow_bwd_kernel (A0, D multiples of 4, aligned
// rows).  blockDim = (width / 4, NY); a thread owns one column quad and the positions ty, ty + NY, ...
// (at most TP of them), keeps the group sums of its positions in registers and writes them once, so
// the kernel is a pure stream: every input is read once with independent 16-byte loads and nothing is
// read-modified-written.  Dynamic shared memory: G * NY * (width / 4) float4.
template <int TP>
__global__ void __launch_bounds__(1024)
h0_reduce_v4_kernel(const float* __restrict__ dy0, const float* __restrict__ h0, int A0, const float* __restrict__ al,
                    const float* __restrict__ be, const float* __restrict__ ga, int T, int G,
                    float* __restrict__ dinv, float* __restrict__ dqb) {
  extern __shared__ float4 sq4[];  // [G][NY][NX]
  const int s = blockIdx.x;
  const int n4 = threadIdx.x, ty = threadIdx.y, nx = blockDim.x, ny = blockDim.y;
  const float4 a = reinterpret_cast<const float4*>(al)[n4], b = reinterpret_cast<const float4*>(be)[n4];
  const float4 c = reinterpret_cast<const float4*>(ga)[n4];
  const float4* d4 = reinterpret_cast<const float4*>(dy0);
  const float4* h4 = reinterpret_cast<const float4*>(h0);
  float4 inv[TP];
#pragma unroll
  for (int i = 0; i < TP; ++i) inv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int g = 0; g < G; ++g) {
    const size_t row = (size_t)s * G + g;
    float4 accq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < TP; ++i) {
      const int t = ty + i * ny;
      if (t < T) {
        const size_t o = (row * T + t) * nx + n4;
        const float4 d = ldg_stream(d4 + o), h = ldg_stream(h4 + o);
        float4 v;
        v.x = fmaf(a.x, d.x, fmaf(b.x, h.x, c.x)); v.y = fmaf(a.y, d.y, fmaf(b.y, h.y, c.y));
        v.z = fmaf(a.z, d.z, fmaf(b.z, h.z, c.z)); v.w = fmaf(a.w, d.w, fmaf(b.w, h.w, c.w));
        accq.x += v.x; accq.y += v.y; accq.z += v.z; accq.w += v.w;
        inv[i].x += v.x; inv[i].y += v.y; inv[i].z += v.z; inv[i].w += v.w;
      }
    }
    sq4[((size_t)g * ny + ty) * nx + n4] = accq;
  }
  if (dinv) {
#pragma unroll
    for (int i = 0; i < TP; ++i) {
      const int t = ty + i * ny;
      if (t < T) reinterpret_cast<float4*>(dinv)[((size_t)s * T + t) * nx + n4] = inv[i];
    }
  }
  __syncthreads();
  for (int g = ty; g < G; g += ny) {
    float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int y = 0; y < ny; ++y) {
      const float4 v = sq4[((size_t)g * ny + y) * nx + n4];
      tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w;
    }
    reinterpret_cast<float4*>(dqb)[((size_t)s * G + g) * nx + n4] = tot;
  }
}

template <int TP>
__global__ void __launch_bounds__(1024)
mulrow_bwd_v4_kernel(const float* __restrict__ dP, int D, const float* __restrict__ as, int lda, int off,
                     const float* __restrict__ tgt, int ldt, int T, int G, float* __restrict__ da2,
                     float* __restrict__ dtgt, int lddt) {
  extern __shared__ float4 sq4[];  // [G][NY][NX]
  const int s = blockIdx.x;
  const int n4 = threadIdx.x, ty = threadIdx.y, nx = blockDim.x, ny = blockDim.y;
  const float4* p4 = reinterpret_cast<const float4*>(dP);
  float4 a2v[TP], da[TP];
#pragma unroll
  for (int i = 0; i < TP; ++i) {
    const int t = ty + i * ny;
    da[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    a2v[i] = da[i];
    if (t < T) a2v[i] = __ldg(reinterpret_cast<const float4*>(as + ((size_t)s * T + t) * lda + off) + n4);
  }
  for (int g = 0; g < G; ++g) {
    const size_t row = (size_t)s * G + g;
    const float4 tg = __ldg(reinterpret_cast<const float4*>(tgt + row * ldt) + n4);
    float4 acct = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < TP; ++i) {
      const int t = ty + i * ny;
      if (t < T) {
        const float4 d = ldg_stream(p4 + (row * T + t) * nx + n4);
        acct.x = fmaf(d.x, a2v[i].x, acct.x); acct.y = fmaf(d.y, a2v[i].y, acct.y);
        acct.z = fmaf(d.z, a2v[i].z, acct.z); acct.w = fmaf(d.w, a2v[i].w, acct.w);
        da[i].x = fmaf(d.x, tg.x, da[i].x); da[i].y = fmaf(d.y, tg.y, da[i].y);
        da[i].z = fmaf(d.z, tg.z, da[i].z); da[i].w = fmaf(d.w, tg.w, da[i].w);
      }
    }
    sq4[((size_t)g * ny + ty) * nx + n4] = acct;
  }
#pragma unroll
  for (int i = 0; i < TP; ++i) {
    const int t = ty + i * ny;
    if (t < T) reinterpret_cast<float4*>(da2)[((size_t)s * T + t) * nx + n4] = da[i];
  }
  __syncthreads();
  for (int g = ty; g < G; g += ny) {
    float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int y = 0; y < ny; ++y) {
      const float4 v = sq4[((size_t)g * ny + y) * nx + n4];
      tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w;
    }
    float4* dst = reinterpret_cast<float4*>(dtgt + ((size_t)s * G + g) * lddt) + n4;
    float4 o = *dst;
    o.x += tot.x; o.y += tot.y; o.z += tot.z; o.w += tot.w;
    *dst = o;
  }
}

// Backward of the feature map F = [a (W1 cols), a[:, :W2] * V[s] (W2 cols)]:
//   da[m,k] = dF[m,k] + (k<W2 ? dF[m,W1+k]*V[s,k] : 0) + (add && k>=add_off ? add[m,k-add_off] : 0)
//   dVout[s,k] = sum_t dF[m,W1+k] * a[m,k]     (k < W2; written, not accumulated)
// One CTA per sequence; blockDim = (NX >= W1, NY).
__global__ void catmul_bwd_kernel(const float* __restrict__ dF, int W1, int W2, const float* __restrict__ a,
                                  int lda, const float* __restrict__ Vq, int ldv,
                                  const float* __restrict__ add, int ldadd, int add_off, int T,
                                  float* __restrict__ da, int ldda, float* __restrict__ dVout, int lddv) {
  extern __shared__ float sq[];  // [NY][W2]
  const int s = blockIdx.x;
  const int k = threadIdx.x, ty = threadIdx.y, ny = blockDim.y;
  const int ldf = W1 + W2;
  float acc = 0.f;
  if (k < W1) {
    const float vq = (k < W2) ? Vq[(size_t)s * ldv + k] : 0.f;
    for (int t = ty; t < T; t += ny) {
      size_t m = (size_t)s * T + t;
      float v = dF[m * ldf + k];
      if (k < W2) {
        float d2 = dF[m * ldf + W1 + k];
        v = fmaf(d2, vq, v);
        acc = fmaf(d2, a[m * lda + k], acc);
      }
      if (add && k >= add_off) v += add[m * ldadd + (k - add_off)];
      da[m * ldda + k] = v;
    }
    if (k < W2) sq[ty * W2 + k] = acc;
  }
  __syncthreads();
  if (k < W2 && ty == 0) {
    float tot = 0.f;
    for (int y = 0; y < ny; ++y) tot += sq[y * W2 + k];
    dVout[(size_t)s * lddv + k] = tot;
  }
}

// dq[b, 0:Q] = gradient of the short-term query [sti[s] (U), tgt[b] (D)]:
//   dsti[s,k] += sum_g dq[(s*G+g), k];   dtgt[b,k] += dq[b, U+k]
__global__ void qs_bwd_kernel(const float* __restrict__ dq, int U, int D, int G, int S,
                              float* __restrict__ dsti, float* __restrict__ dtgt) {
  const int Q = U + D;
  int n = S * Q;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int s = i / Q, k = i - s * Q;
    if (k < U) {
      float acc = 0.f;
      for (int g = 0; g < G; ++g) acc += dq[((size_t)s * G + g) * Q + k];
      dsti[(size_t)s * U + k] += acc;
    } else {
      for (int g = 0; g < G; ++g) {
        size_t b = (size_t)s * G + g;
        dtgt[b * D + (k - U)] += dq[b * Q + k];
      }
    }
  }
}

}  // namespace clsr
