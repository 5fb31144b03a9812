// Per-row "head" of the CLSR graph: alpha gate + fusion (clsr.py:239-275), output units of the
// _fcn_net MLPs (base_model.py:688-706), BatchNorm statistic finalisation (base_model.py:673-679),
// the losses (clsr.py:22-82, base_model.py:215-247) with their gradients, dense-variable L2 /
// clip / Adam (base_model.py:281-297), and the small block-copy kernel that builds the folded
// weight matrices and maps their gradients back onto the TF variables.
#pragma once
#include "common.cuh"
#include "embed.cuh"

namespace clsr {

// ---- block copy / linear-combination descriptors ------------------------------------------------
struct BlockOp {
  long long dst, src1, src2;  // float offsets into the dst / src base pointers; src2 < 0: unused
  int rows, cols, ldd, lds1, lds2;
  float c1, c2;
  int transpose;  // dst[c,r] instead of dst[r,c]
  int accumulate;
};

// dst[r,c] (+)= c1*src1[r,c] + c2*src2[r,c]; one CTA per descriptor.
__global__ void blockop_kernel(const BlockOp* __restrict__ ops, float* __restrict__ dst_base,
                               const float* __restrict__ src_base) {
  const BlockOp op = ops[blockIdx.x];
  const int n = op.rows * op.cols;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int r = i / op.cols, c = i - r * op.cols;
    float v = op.c1 * src_base[op.src1 + (long long)r * op.lds1 + c];
    if (op.src2 >= 0) v = fmaf(op.c2, src_base[op.src2 + (long long)r * op.lds2 + c], v);
    long long o = op.transpose ? op.dst + (long long)c * op.ldd + r : op.dst + (long long)r * op.ldd + c;
    if (op.accumulate) dst_base[o] += v; else dst_base[o] = v;
  }
}

// ---- sequence bookkeeping -----------------------------------------------------------------------
// len[s] = sum_t mask[s,t] (clsr.py:150); counts[0] += (len > thr) (contrastive mask, clsr.py:48-52).
__global__ void seq_prep_kernel(const int32_t* __restrict__ mask, int seq_stride, int T, int S, int thr,
                                int32_t* __restrict__ len, int32_t* __restrict__ counts) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
    int l = 0;
    const int32_t* mp = mask + (size_t)s * seq_stride;
    for (int t = 0; t < T; ++t) l += mp[t];
    len[s] = l;
    if (l > thr) atomicAdd(counts, 1);
  }
}

// ---- BatchNorm statistics -----------------------------------------------------------------------
// Training: batch mean / biased variance from (sum, sum of squares); moving statistics move by
// (1 - momentum) toward the batch values.  Inference: moving statistics.
__global__ void bn_fwd_finalize_kernel(const double* __restrict__ stat, int N, double count,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float eps, float momentum, float* __restrict__ mmean,
                                       float* __restrict__ mvar, int train, int update_moving,
                                       float* __restrict__ scale, float* __restrict__ shift,
                                       float* __restrict__ mean_o, float* __restrict__ rstd_o) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float mean, var;
  if (train) {
    double m = stat[n] / count;
    double v = stat[N + n] / count - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m; var = (float)v;
    if (update_moving) {
      mmean[n] -= (mmean[n] - mean) * (1.f - momentum);
      mvar[n] -= (mvar[n] - var) * (1.f - momentum);
    }
  } else {
    mean = mmean[n]; var = mvar[n];
  }
  float rstd = rsqrtf(var + eps);
  rstd = rstd * (1.5f - 0.5f * (var + eps) * rstd * rstd);  // one Newton step: full fp32 accuracy
  float sc = rstd * gamma[n];
  scale[n] = sc;
  shift[n] = beta[n] - mean * sc;
  mean_o[n] = mean;
  rstd_o[n] = rstd;
}

// dx = al*dy + be*h + ga;  dgamma += sum dy*xhat;  dbeta += sum dy.
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ stat, int N, double count,
                                       const float* __restrict__ gamma, const float* __restrict__ mean,
                                       const float* __restrict__ rstd, float* __restrict__ al,
                                       float* __restrict__ be, float* __restrict__ ga,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float add_scale) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double s1 = stat[n], s2 = stat[N + n];
  float m1 = (float)(s1 / count), m2 = (float)(s2 / count);
  float g = gamma[n], r = rstd[n], mu = mean[n];
  al[n] = g * r;
  be[n] = -g * r * r * m2;
  ga[n] = -g * r * m1 + g * r * r * m2 * mu;
  dgamma[n] += add_scale * (float)s2;
  dbeta[n] += add_scale * (float)s1;
}

// Data-parallel forms: the all-reduce of the (sum, sum of squares) / (sum dy, sum dy*xhat) vectors over the ranks
// happens INSIDE the finalize kernel, over NVLink peer memory (peer_allreduce_block, embed.cuh) -- one launch
// per BatchNorm layer and direction instead of an NCCL call plus a finalize launch.  `count` is the global row count.
__global__ void __launch_bounds__(256)
bn_fwd_finalize_peer_kernel(PeerComm pc, unsigned long long seq, double* __restrict__ stat, int N, double count,
                            const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                            float momentum, float* __restrict__ mmean, float* __restrict__ mvar, int update_moving,
                            float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_o,
                            float* __restrict__ rstd_o) {
  __shared__ double vals[kPeerSlots];
  for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) vals[i] = stat[i];
  __syncthreads();
  peer_allreduce_block(pc, seq, vals, 2 * N);
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    stat[n] = vals[n]; stat[N + n] = vals[N + n];
    double m = vals[n] / count;
    double v = vals[N + n] / count - m * m;
    if (v < 0.0) v = 0.0;
    const float mean = (float)m, var = (float)v;
    if (update_moving) {
      mmean[n] -= (mmean[n] - mean) * (1.f - momentum);
      mvar[n] -= (mvar[n] - var) * (1.f - momentum);
    }
    float rstd = rsqrtf(var + eps);
    rstd = rstd * (1.5f - 0.5f * (var + eps) * rstd * rstd);
    const float sc = rstd * gamma[n];
    scale[n] = sc;
    shift[n] = beta[n] - mean * sc;
    mean_o[n] = mean;
    rstd_o[n] = rstd;
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_finalize_peer_kernel(PeerComm pc, unsigned long long seq, double* __restrict__ stat, int N, double count,
                            const float* __restrict__ gamma, const float* __restrict__ mean,
                            const float* __restrict__ rstd, float* __restrict__ al, float* __restrict__ be,
                            float* __restrict__ ga, float* __restrict__ dgamma, float* __restrict__ dbeta,
                            float add_scale) {
  __shared__ double vals[kPeerSlots];
  for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) vals[i] = stat[i];
  __syncthreads();
  peer_allreduce_block(pc, seq, vals, 2 * N);
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const double s1 = vals[n], s2 = vals[N + n];
    stat[n] = s1; stat[N + n] = s2;
    const float m1 = (float)(s1 / count), m2 = (float)(s2 / count);
    const float g = gamma[n], r = rstd[n], mu = mean[n];
    al[n] = g * r;
    be[n] = -g * r * r * m2;
    ga[n] = -g * r * m1 + g * r * r * m2 * mu;
    dgamma[n] += add_scale * (float)s2;
    dbeta[n] += add_scale * (float)s1;
  }
}

// ---- output units -------------------------------------------------------------------------------
// out[m] = relu(bn(h[m,:])) . w + b ; one warp per row.
__global__ void rowdot_kernel(const float* __restrict__ h, int N, const float* __restrict__ scale,
                              const float* __restrict__ shift, const float* __restrict__ w,
                              const float* __restrict__ b, int rows, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int m = blockIdx.x * wpb + (threadIdx.x >> 5); m < rows; m += gridDim.x * wpb) {
    float acc = 0.f;
    for (int n = lane; n < N; n += 32)
      acc = fmaf(fmaxf(0.f, fmaf(h[(size_t)m * N + n], scale[n], shift[n])), w[n], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[m] = acc + b[0];
  }
}

// dy[m,n] = dout[m]*w[n] where bn(h)>0; stat += (sum dy, sum dy*xhat); dw += sum relu(bn(h))*dout;
// db += sum dout.  blockDim.x >= N threads over channels, blockDim.y row lanes.
__global__ void rowdot_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ h, int N,
                                  const float* __restrict__ scale, const float* __restrict__ shift,
                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                  const float* __restrict__ w, int rows, float* __restrict__ dy,
                                  double* __restrict__ stat, float* __restrict__ dw, float* __restrict__ db) {
  const int n = threadIdx.x;
  float s1 = 0.f, s2 = 0.f, sw = 0.f, sb = 0.f;
  if (n < N) {
    const float sc = scale[n], sh = shift[n], mu = mean[n], rs = rstd[n], wn = w[n];
    for (int m = blockIdx.x * blockDim.y + threadIdx.y; m < rows; m += gridDim.x * blockDim.y) {
      float hv = h[(size_t)m * N + n];
      float y = fmaf(hv, sc, sh);
      float d = dout[m];
      float v = (y > 0.f) ? d * wn : 0.f;
      dy[(size_t)m * N + n] = v;
      s1 += v;
      s2 = fmaf(v, (hv - mu) * rs, s2);
      sw = fmaf(fmaxf(y, 0.f), d, sw);
      if (n == 0) sb += d;
    }
    atomicAdd(stat + n, (double)s1);
    atomicAdd(stat + N + n, (double)s2);
    atomicAdd(dw + n, sw);
    if (n == 0) atomicAdd(db, sb);
  }
}

// ---- alpha gate / fusion ------------------------------------------------------------------------
// ca[b,:] = [fs[s], tgt[b], afl[s], afs[b], time_to_now[s, T-1]]  (clsr.py:239-248)
__global__ void concat_alpha_kernel(const float* __restrict__ fs, const float* __restrict__ tgt,
                                    const float* __restrict__ afl, const float* __restrict__ afs,
                                    const float* __restrict__ ttn, int seq_stride, int T, int H, int D, int G,
                                    int B, float* __restrict__ ca, int Hf) {
  // Hf: width of the causal2 final state in front (H, or 0 with predict_long_short = False: fs is then not read)
  const int CA = Hf + 2 * D + H + 1;
  long long n = (long long)B * CA;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int b = (int)(i / CA), k = (int)(i - (long long)b * CA);
    int s = b / G;
    float v;
    if (k < Hf) v = fs[(size_t)s * H + k];
    else if (k < Hf + D) v = tgt[(size_t)b * D + (k - Hf)];
    else if (k < Hf + 2 * D) v = afl[(size_t)s * D + (k - Hf - D)];
    else if (k < Hf + 2 * D + H) v = afs[(size_t)b * H + (k - Hf - 2 * D)];
    else v = ttn[(size_t)s * seq_stride + (T - 1)];
    ca[i] = v;
  }
}

// alpha = sigmoid(alpha_logit); mo[b,:] = [afl[s]*alpha + afs[b]*(1-alpha), tgt[b]]  (clsr.py:264-275)
__global__ void head_mid_kernel(const float* __restrict__ alogit, const float* __restrict__ afl,
                                const float* __restrict__ afs, const float* __restrict__ tgt, int H, int D,
                                int G, int B, float* __restrict__ alpha, float* __restrict__ mo) {
  const int W = H + D;
  long long n = (long long)B * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int b = (int)(i / W), k = (int)(i - (long long)b * W);
    int s = b / G;
    float a = sigmoid_acc(alogit[b]);
    if (k == 0) alpha[b] = a;
    mo[i] = (k < H) ? afl[(size_t)s * D + k] * a + afs[(size_t)b * H + k] * (1.f - a)
                    : tgt[(size_t)b * D + (k - H)];
  }
}

// dalpha_logit[b] = (sum_d dmo[b,d] * (afl[s,d] - afs[b,d])) * alpha*(1-alpha); one warp per row.
__global__ void head_mid_bwd_kernel(const float* __restrict__ dmo, int ldmo, const float* __restrict__ afl,
                                    const float* __restrict__ afs, const float* __restrict__ alpha, int H,
                                    int G, int B, float* __restrict__ dalogit) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < B; b += gridDim.x * wpb) {
    int s = b / G;
    float acc = 0.f;
    for (int d = lane; d < H; d += 32)
      acc = fmaf(dmo[(size_t)b * ldmo + d], afl[(size_t)s * H + d] - afs[(size_t)b * H + d], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      float a = alpha[b];
      dalogit[b] = acc * a * (1.f - a);
    }
  }
}

// ---- losses -------------------------------------------------------------------------------------
// Grouped softmax loss (base_model.py:215-235) over groups of Gs consecutive rows, its gradient,
// and pred = sigmoid(logit).  acc[0] += -(Gs/B) * sum_{label==1} log softmax.
__global__ void softmax_loss_kernel(const float* __restrict__ logit, const float* __restrict__ labels, int Gs,
                                    int B, int Btot, float* __restrict__ dlogit, float* __restrict__ pred,
                                    double* __restrict__ acc) {
  const int ng = B / Gs;
  float part = 0.f;
  const float coef = (float)Gs / (float)Btot;  // Btot: rows of the global (all-rank) batch
  for (int gi = blockIdx.x * blockDim.x + threadIdx.x; gi < ng; gi += gridDim.x * blockDim.x) {
    const float* lp = logit + (size_t)gi * Gs;
    float mx = -INFINITY;
    for (int j = 0; j < Gs; ++j) mx = fmaxf(mx, lp[j]);
    float sum = 0.f;
    for (int j = 0; j < Gs; ++j) sum += expf(lp[j] - mx);
    int npos = 0;
    for (int j = 0; j < Gs; ++j) npos += (labels[(size_t)gi * Gs + j] == 1.0f);
    for (int j = 0; j < Gs; ++j) {
      float sm = expf(lp[j] - mx) / sum;
      bool pos = labels[(size_t)gi * Gs + j] == 1.0f;
      if (pos) part -= coef * logf(sm);
      dlogit[(size_t)gi * Gs + j] = coef * ((float)npos * sm - (pos ? 1.f : 0.f));
      pred[(size_t)gi * Gs + j] = sigmoid_acc(lp[j]);
    }
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0 && part != 0.f) atomicAdd(acc, (double)part);
}

__global__ void sigmoid_kernel(const float* __restrict__ x, int n, float* __restrict__ y) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    y[i] = sigmoid_acc(x[i]);
}

// BPR contrastive loss (clsr.py:53-57), per row: x1 = afl.(hr-hm), x2 = afs.(hm-hr), x3 = hm.(afs-afl),
// x4 = hr.(afl-afs); loss terms softplus(x_i); gs[b,i] = sigmoid(x_i) (the derivative) for the backward.
// One warp per row; acc[1..4] += sum over rows of mask*softplus(x_i).
__global__ void bpr_dots_kernel(const float* __restrict__ afl, const float* __restrict__ afs,
                                const float* __restrict__ hm, const float* __restrict__ hr,
                                const int32_t* __restrict__ len, int thr, int D, int G, int B,
                                float* __restrict__ gs, double* __restrict__ acc) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float l1 = 0.f, l2 = 0.f, l3 = 0.f, l4 = 0.f;
  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < B; b += gridDim.x * wpb) {
    const int s = b / G;
    float x1 = 0.f, x2 = 0.f, x3 = 0.f, x4 = 0.f;
    for (int d = lane; d < D; d += 32) {
      float l = afl[(size_t)s * D + d], sh = afs[(size_t)b * D + d], m = hm[(size_t)s * D + d], r = hr[(size_t)s * D + d];
      x1 = fmaf(l, r - m, x1); x2 = fmaf(sh, m - r, x2); x3 = fmaf(m, sh - l, x3); x4 = fmaf(r, l - sh, x4);
    }
    x1 = warp_sum(x1); x2 = warp_sum(x2); x3 = warp_sum(x3); x4 = warp_sum(x4);
    if (lane == 0) {
      const float cm = (len[s] > thr) ? 1.f : 0.f;
      const float xs[4] = {x1, x2, x3, x4};
      float sp[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sp[i] = fmaxf(xs[i], 0.f) + log1pf(expf(-fabsf(xs[i])));
        gs[(size_t)b * 4 + i] = sigmoid_acc(xs[i]);
      }
      l1 += cm * sp[0]; l2 += cm * sp[1]; l3 += cm * sp[2]; l4 += cm * sp[3];
    }
  }
  if (lane == 0) {
    if (l1 != 0.f) atomicAdd(acc + 1, (double)l1);
    if (l2 != 0.f) atomicAdd(acc + 2, (double)l2);
    if (l3 != 0.f) atomicAdd(acc + 3, (double)l3);
    if (l4 != 0.f) atomicAdd(acc + 4, (double)l4);
  }
}

// Everything that fans out of the head, one thread per (sequence, channel), looping the G rows of
// the group: gradients of the fusion, of concat_all / model_output, and of the triplet contrastive
// loss (clsr.py:58-71) whose value is accumulated in acc[1..4].
//   dfs[s,:], dafl[s,:], dhm[s,:], dhr[s,:] (group sums); dtgt[b,:], dafs[b,:] (per row).
__global__ void head_final_bwd_kernel(const float* __restrict__ dmo, const float* __restrict__ dca,
                                      const float* __restrict__ alpha, const float* __restrict__ afl,
                                      const float* __restrict__ afs, const float* __restrict__ hm,
                                      const float* __restrict__ hr, const int32_t* __restrict__ len,
                                      const int32_t* __restrict__ counts, int thr, float margin, float cw,
                                      const float* __restrict__ bpr_gs, int H, int D, int G, int S,
                                      float* __restrict__ dfs,
                                      float* __restrict__ dtgt, float* __restrict__ dafl,
                                      float* __restrict__ dafs, float* __restrict__ dhm,
                                      float* __restrict__ dhr, double* __restrict__ acc, int Hf) {
  const int CA = Hf + 2 * D + H + 1, W = H + D;   // Hf: see concat_alpha_kernel
  const float den = (float)G * (float)counts[0];
  float l1 = 0.f, l2 = 0.f, l3 = 0.f, l4 = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S * D; i += gridDim.x * blockDim.x) {
    const int s = i / D, d = i - s * D;
    const float cmv = (len[s] > thr) ? 1.f : 0.f;
    const float c = cw * cmv / den;
    const float l = afl[i], m = hm[i], r = hr[i];
    float sfs = 0.f, sl = 0.f, sm = 0.f, sr = 0.f;
    for (int g = 0; g < G; ++g) {
      const size_t b = (size_t)s * G + g;
      const float a = alpha[b];
      const float due = dmo[b * W + d];
      const float sh = afs[b * H + d];
      if (Hf) sfs += dca[b * CA + d];
      dtgt[b * D + d] = dmo[b * W + H + d] + dca[b * CA + Hf + d];
      float dl = dca[b * CA + Hf + D + d] + due * a;
      float ds = dca[b * CA + Hf + 2 * D + d] + due * (1.f - a);
      if (bpr_gs) {  // BPR: gradients through the four inner products, derivative = sigmoid(x_i)
        const float g1 = c * bpr_gs[b * 4 + 0], g2 = c * bpr_gs[b * 4 + 1], g3 = c * bpr_gs[b * 4 + 2],
                    g4 = c * bpr_gs[b * 4 + 3];
        dl += g1 * (r - m) - g3 * m + g4 * r;
        ds += g2 * (m - r) + g3 * m - g4 * r;
        sm += -g1 * l + g2 * sh + g3 * (sh - l);
        sr += g1 * l - g2 * sh + g4 * (l - sh);
        sl += dl;
        dafs[b * H + d] = ds;
        continue;
      }
      const float elm = l - m, elr = l - r, esm = sh - m, esr = sh - r;
      const float dlm = elm * elm, dlr = elr * elr, dsm = esm * esm, dsr = esr * esr;
      const float v1 = dlm - dlr + margin, v2 = dsr - dsm + margin, v3 = dlm - dsm + margin,
                  v4 = dsr - dlr + margin;
      const float a1 = v1 > 0.f ? 1.f : 0.f, a2 = v2 > 0.f ? 1.f : 0.f, a3 = v3 > 0.f ? 1.f : 0.f,
                  a4 = v4 > 0.f ? 1.f : 0.f;
      l1 += cmv * fmaxf(v1, 0.f); l2 += cmv * fmaxf(v2, 0.f);
      l3 += cmv * fmaxf(v3, 0.f); l4 += cmv * fmaxf(v4, 0.f);
      const float glm = c * (a1 + a3), glr = -c * (a1 + a4), gsm = -c * (a2 + a3), gsr = c * (a2 + a4);
      dl += 2.f * (elm * glm + elr * glr);
      ds += 2.f * (esm * gsm + esr * gsr);
      sm -= 2.f * (elm * glm + esm * gsm);
      sr -= 2.f * (elr * glr + esr * gsr);
      sl += dl;
      dafs[b * H + d] = ds;
    }
    dfs[i] = sfs; dafl[i] = sl; dhm[i] = sm; dhr[i] = sr;
  }
  l1 = warp_sum(l1); l2 = warp_sum(l2); l3 = warp_sum(l3); l4 = warp_sum(l4);
  if ((threadIdx.x & 31) == 0 && !bpr_gs) {
    atomicAdd(acc + 1, (double)l1); atomicAdd(acc + 2, (double)l2);
    atomicAdd(acc + 3, (double)l3); atomicAdd(acc + 4, (double)l4);
  }
}

// ---- dense variables: L2, per-variable clip norm, Adam ------------------------------------------
struct DenseVar {
  long long off;
  int n;
  int trainable;
};

// One CTA per variable: g += l2*w; norms[v] = ||g||^2; acc[10] += 0.5*l2*||w||^2.
__global__ void dense_l2_norm_kernel(const DenseVar* __restrict__ vars, const float* __restrict__ w,
                                     float* __restrict__ g, float l2, float* __restrict__ norms,
                                     double* __restrict__ acc, double acc_scale) {
  __shared__ float red[2][32];
  const DenseVar v = vars[blockIdx.x];
  float ss = 0.f, ww = 0.f;
  if (v.trainable) {
    for (int i = threadIdx.x; i < v.n; i += blockDim.x) {
      float x = w[v.off + i];
      float gv = fmaf(l2, x, g[v.off + i]);
      g[v.off + i] = gv;
      ss = fmaf(gv, gv, ss);
      ww = fmaf(x, x, ww);
    }
  }
  ss = warp_sum(ss); ww = warp_sum(ww);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][wid] = ss; red[1][wid] = ww; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += red[0][i]; b += red[1][i]; }
    norms[blockIdx.x] = a;
    if (v.trainable && acc_scale != 0.0) atomicAdd(acc + 10, acc_scale * 0.5 * (double)l2 * (double)b);
  }
}

struct AdamDense {
  float lr_t, beta1, beta2, eps, clip;
  const float* lr_dev;   // see AdamHyper
};

__global__ void dense_adam_kernel(const DenseVar* __restrict__ vars, float* __restrict__ w,
                                  float* __restrict__ m, float* __restrict__ v2,
                                  const float* __restrict__ g, const float* __restrict__ norms, AdamDense hp) {
  const DenseVar v = vars[blockIdx.x];
  if (!v.trainable) return;
  const float lr_t = hp.lr_dev ? __ldg(hp.lr_dev) : hp.lr_t;
  float scale = 1.f;
  if (hp.clip > 0.f) scale = hp.clip / fmaxf(sqrtf(norms[blockIdx.x]), hp.clip);
  for (int i = threadIdx.x; i < v.n; i += blockDim.x) {
    long long o = v.off + i;
    float gv = g[o] * scale;
    float mv = hp.beta1 * m[o] + (1.f - hp.beta1) * gv;
    float vv = hp.beta2 * v2[o] + (1.f - hp.beta2) * gv * gv;
    m[o] = mv; v2[o] = vv;
    w[o] -= lr_t * mv / (sqrtf(vv) + hp.eps);
  }
}

// Scalar outputs of the step (clsr.py:22-34): [loss, data, regular, contrastive, discrepancy].
// acc: [0] data, [1..4] contrastive sums, [5] item rows^2, [6] cate rows^2, [7] user_long rows^2,
//      [8] user_short rows^2, [9] sum (long-short)^2, [10] dense regular term.
// out[5..8] = L2 norms of the four table gradients as the clip sees them (tf.clip_by_norm on the
// IndexedSlices values, base_model.py:289-297); clip_steps[0] += 1 when a shared-history step (G > 1) has one
// of them above max_grad_norm -- the case in which the group-summed slice norm differs from TF's.
__global__ void loss_finalize_kernel(const double* __restrict__ acc, const int32_t* __restrict__ counts,
                                     const int32_t* __restrict__ n_users, int G, int U, float embed_l2,
                                     float cw, float dw, const double* __restrict__ sumsq, float clip, int group,
                                     unsigned long long* __restrict__ clip_steps, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double den = (double)G * (double)counts[0];
  double con = (double)cw * (acc[1] + acc[2] + acc[3] + acc[4]) / den;
  double reg = 0.5 * (double)embed_l2 * (acc[5] + acc[6] + acc[7] + acc[8]) + acc[10];
  double disc = -(double)dw * acc[9] / ((double)n_users[0] * (double)U);
  double data = acc[0];
  out[0] = (float)(data + reg + con + disc);
  out[1] = (float)data; out[2] = (float)reg; out[3] = (float)con; out[4] = (float)disc;
  bool active = false;
  for (int t = 0; t < 4; ++t) {
    const float nrm = (float)sqrt(sumsq[t]);
    out[5 + t] = nrm;
    active = active || (clip > 0.f && nrm > clip);
  }
  if (active && group > 1) clip_steps[0] += 1ull;
}

}  // namespace clsr
