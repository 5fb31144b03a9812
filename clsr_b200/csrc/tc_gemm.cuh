// Tensor-core linear layer for sm_100a: C[M,N] (+)= op_a(A)[M,K] . W[K,N] with the same fused
// operand prologues / epilogues as gemm.cuh, computed by tcgen05.mma with the accumulator in TMEM.
//
// Precision: every fp32 operand x is split into two bf16 terms x = hi + lo (hi = bf16(x),
// lo = bf16(x - hi)); a product is accumulated as hi*hi + lo*hi + hi*lo in fp32 (three
// tcgen05.mma.kind::f16 per 16-wide K block), i.e. ~2^-16 relative operand error.  A single bf16
// pass (2^-8) misses the 1e-3 logit parity bar of this path.
//
// Structure (persistent CTAs, 13 warps):
//   warps 0-7  producers : read 128-row operand tiles from HBM (16-byte loads), apply the prologue,
//                          split, write the hi/lo tiles into shared memory in the UMMA canonical
//                          K-major layout (no swizzle: planes of 8 K-elements, 16 bytes per row)
//   warp  12   MMA       : one elected thread issues tcgen05.mma; tcgen05.commit releases the
//                          operand stage and publishes the accumulator
//   warps 8-11 epilogue  : tcgen05.ld the 128 x N accumulator (one TMEM lane = one row per thread),
//                          stage 32-column slabs through shared memory, apply the epilogue with
//                          coalesced global access, accumulate BatchNorm column statistics
// Two operand stages and two TMEM accumulators are ring-buffered with mbarriers so the producer
// of tile i+1, the MMA of tile i and the epilogue of tile i-1 overlap.  W (K x N, small) is split
// once per CTA and stays resident in shared memory.
#pragma once
#include <cuda_bf16.h>

#include "gemm.cuh"

namespace clsr {
namespace tc {

constexpr int kTileM = 128;
constexpr int kProducers = 256;
constexpr int kThreads = kProducers + 128 + 32;  // 8 producer warps, 4 epilogue warps, 1 MMA warp
constexpr int kEpiCols = 32;

CLSR_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

CLSR_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
CLSR_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
CLSR_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "CLSR_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra CLSR_WAIT_DONE;\n"
      "bra CLSR_WAIT_LOOP;\n"
      "CLSR_WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
CLSR_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
CLSR_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
CLSR_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

CLSR_DEVINL void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
CLSR_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
CLSR_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
CLSR_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
CLSR_DEVINL void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4, [16,30) leading (K-direction) byte offset >> 4,
//   [32,46) stride (M/N-direction, between 8-row groups) byte offset >> 4, [46,48) version = 1.
CLSR_DEVINL uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: D=f32, A=B=bf16, both K-major, M=128.
CLSR_DEVINL uint32_t make_idesc(int npad) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

CLSR_DEVINL void split_store8(const float* x, uint8_t* hi_dst, uint8_t* lo_dst) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * i]), h1 = __float2bfloat16_rn(x[2 * i + 1]);
    __nv_bfloat16 l0 = __float2bfloat16_rn(x[2 * i] - __bfloat162float(h0));
    __nv_bfloat16 l1 = __float2bfloat16_rn(x[2 * i + 1] - __bfloat162float(h1));
    h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  *reinterpret_cast<uint4*>(hi_dst) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_dst) = make_uint4(l[0], l[1], l[2], l[3]);
}

// Operand fetch for one producer task = eight consecutive elements A'[m, k0..k0+7].  issue() only
// starts the 16-byte global loads (so several tasks are in flight per thread), finish() applies the
// prologue.  Every mode used on large-M GEMMs has a vector form; anything else falls back to
// element-wise AOp::load.
struct Raw8 {
  float4 a0, a1, b0, b1;
  int kind;  // 0 zeros, 1 one stream (a), 2 two streams (a, b), 3 scalar fallback
};

CLSR_DEVINL bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

CLSR_DEVINL void issue8(const AOp& a, int m, int k0, int M, int K, Raw8& r) {
  r.kind = 0;
  if (m >= M || k0 >= K) return;
  r.kind = 3;
  if (k0 + 8 > K) return;
  switch (a.mode) {
    case A_PLAIN:
    case A_BNRELU:
      if ((a.lda & 3) == 0 && al16(a.A)) {
        const float4* p = reinterpret_cast<const float4*>(a.A + (size_t)m * a.lda + k0);
        r.a0 = __ldg(p); r.a1 = __ldg(p + 1); r.kind = 1;
      }
      break;
    case A_AFFINE2:
      if ((a.lda & 3) == 0 && (a.lda2 & 3) == 0 && al16(a.A) && al16(a.A2)) {
        const float4* p = reinterpret_cast<const float4*>(a.A + (size_t)m * a.lda + k0);
        const float4* q = reinterpret_cast<const float4*>(a.A2 + (size_t)m * a.lda2 + k0);
        r.a0 = __ldg(p); r.a1 = __ldg(p + 1); r.b0 = __ldg(q); r.b1 = __ldg(q + 1); r.kind = 2;
      }
      break;
    case A_CATMUL:
      if ((a.lda & 3) == 0 && (a.lda2 & 3) == 0 && (a.W1 & 7) == 0 && (a.off & 3) == 0 && al16(a.A) && al16(a.A2)) {
        if (k0 < a.W1) {
          const float4* p = reinterpret_cast<const float4*>(a.A + (size_t)m * a.lda + k0);
          r.a0 = __ldg(p); r.a1 = __ldg(p + 1); r.kind = 1;
        } else {
          int kk = k0 - a.W1;
          const float4* p = reinterpret_cast<const float4*>(a.A + (size_t)m * a.lda + a.off + kk);
          const float4* q = reinterpret_cast<const float4*>(a.A2 + (size_t)(m / a.T) * a.lda2 + kk);
          r.a0 = __ldg(p); r.a1 = __ldg(p + 1); r.b0 = __ldg(q); r.b1 = __ldg(q + 1); r.kind = 2;
        }
      }
      break;
    case A_MULROW:
      if ((a.lda & 3) == 0 && (a.lda2 & 3) == 0 && (a.off & 3) == 0 && al16(a.A) && al16(a.A2)) {
        int b = m / a.T, t = m - b * a.T, sq = b / a.G;
        const float4* p = reinterpret_cast<const float4*>(a.A + ((size_t)sq * a.T + t) * a.lda + a.off + k0);
        const float4* q = reinterpret_cast<const float4*>(a.A2 + (size_t)b * a.lda2 + k0);
        r.a0 = __ldg(p); r.a1 = __ldg(p + 1); r.b0 = __ldg(q); r.b1 = __ldg(q + 1); r.kind = 2;
      }
      break;
    default:
      break;
  }
}

CLSR_DEVINL void finish8(const AOp& a, int m, int k0, int K, const Raw8& r, float* x) {
  if (r.kind == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 0.f;
    return;
  }
  if (r.kind == 3) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = (k0 + i < K) ? a.load(m, k0 + i) : 0.f;
    return;
  }
  const float v[8] = {r.a0.x, r.a0.y, r.a0.z, r.a0.w, r.a1.x, r.a1.y, r.a1.z, r.a1.w};
  if (r.kind == 1) {
    if (a.mode == A_BNRELU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fmaxf(0.f, fmaf(v[i], a.v0[k0 + i], a.v1[k0 + i]));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = v[i];
    }
    return;
  }
  const float h[8] = {r.b0.x, r.b0.y, r.b0.z, r.b0.w, r.b1.x, r.b1.y, r.b1.z, r.b1.w};
  if (a.mode == A_AFFINE2) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fmaf(a.v0[k0 + i], v[i], fmaf(a.v1[k0 + i], h[i], a.v2[k0 + i]));
  } else {  // A_CATMUL second half, A_MULROW: element-wise product of the two streams
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = v[i] * h[i];
  }
}

struct Smem {
  // byte offsets into dynamic shared memory
  int w_hi, w_lo, a_stage0, a_stage_bytes, epi, eop, eop_bytes, bars, total;
};
// eop != 0 reserves two [128 x npad] fp32 tiles for the prefetched epilogue operand.
__host__ __device__ inline Smem smem_layout(int kpad, int npad, int nstages, int eop = 0) {
  Smem s;
  int wbytes = kpad * npad * 2;
  s.w_hi = 0;
  s.w_lo = wbytes;
  s.a_stage0 = 2 * wbytes;
  s.a_stage_bytes = 2 * kTileM * kpad * 2;  // hi + lo
  s.epi = s.a_stage0 + nstages * s.a_stage_bytes;
  s.eop = s.epi + kTileM * (kEpiCols + 1) * 4;
  s.eop_bytes = eop ? kTileM * npad * 4 : 0;
  s.bars = s.eop + 2 * s.eop_bytes;
  s.total = s.bars + 128;
  return s;
}

CLSR_DEVINL void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
CLSR_DEVINL void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// STATS: accumulate per-column statistics into ep.stat (see gemm.cuh).
template <bool STATS>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(int M, int N, int K, int kpad, int npad, int nstages, uint32_t tmem_cols, int eop_kind, AOp a,
               const float* __restrict__ W, int ldw, EpiOp ep) {
  // eop_kind: which per-element epilogue operand the producers prefetch into shared memory with
  // cp.async two tiles ahead (0 none, 1 hpre, 2 group-add rows, 3 old C for accumulation), so the
  // epilogue itself issues no dependent global loads.
  extern __shared__ __align__(128) uint8_t smem[];
  const Smem L = smem_layout(kpad, npad, nstages, eop_kind);
  uint8_t* w_hi = smem + L.w_hi;
  uint8_t* w_lo = smem + L.w_lo;
  float* epi = reinterpret_cast<float*>(smem + L.epi);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* full = bars;          // [2]
  uint64_t* empty = bars + 2;     // [2]
  uint64_t* tfull = bars + 4;     // [2]
  uint64_t* tempty = bars + 6;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  __shared__ float sst[4][2][kEpiCols];  // per epilogue warp: column partial sums of the current slab
  __shared__ double dacc[STATS ? 2 : 1][STATS ? 256 : 1];  // per-CTA column statistics across tiles

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = (M + kTileM - 1) / kTileM;
  const int nkb = kpad >> 4;

  // ---- one-time setup: W split into canonical hi/lo tiles, barriers, TMEM ----
  for (int idx = tid; idx < (kpad >> 3) * npad; idx += kThreads) {
    int c = idx / npad, n = idx - c * npad;  // plane c holds k = 8c..8c+7 for every n
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int k = c * 8 + i;
      x[i] = (k < K && n < N) ? W[(size_t)k * ldw + n] : 0.f;
    }
    size_t off = (size_t)c * npad * 16 + (size_t)n * 16;
    split_store8(x, w_hi + off, w_lo + off);
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], kProducers);
      mbar_init(&empty[i], 1);
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) tmem_alloc(tmem_slot, tmem_cols);
  if (STATS) {
    for (int i = tid; i < 256; i += kThreads) { dacc[0][i] = 0.0; dacc[1][i] = 0.0; }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // =============================== producers ===============================
    const int nchunk = kpad >> 3;
    const int ntask = kTileM * nchunk;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it % nstages;
      const uint32_t ph = (it / nstages) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      uint8_t* a_hi = smem + L.a_stage0 + s * L.a_stage_bytes;
      uint8_t* a_lo = a_hi + kTileM * kpad * 2;
      const int m0 = tile * kTileM;
      if (eop_kind) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);  // the epilogue of tile it-2 has released this tile
        float* dst = reinterpret_cast<float*>(smem + L.eop + acc * L.eop_bytes);
        const int nq = N >> 2;                          // 16-byte pieces per row (N % 4 == 0 checked on the host)
        for (int i = tid; i < kTileM * nq; i += kProducers) {
          const int r = i / nq, q4 = (i - r * nq) * 4;
          const int m = m0 + r;
          if (m >= M) continue;
          const float* src;
          if (eop_kind == 1) src = ep.hpre + (size_t)m * ep.ldh + q4;
          else if (eop_kind == 2) {
            int b = m / ep.T, t = m - b * ep.T;
            src = ep.ga + ((size_t)(b / ep.G) * ep.T + t) * ep.ldga + q4;
          } else src = ep.C + (size_t)m * ep.ldc + q4;
          cp_async16(dst + r * npad + q4, src);
        }
      }
      constexpr int UN = 4;
      for (int task0 = tid; task0 < ntask; task0 += kProducers * UN) {
        // consecutive threads take consecutive 8-element chunks of one row (coalesced 32-byte pieces);
        // UN tasks' loads are issued before the first is consumed
        Raw8 raw[UN];
        int rr[UN], cc[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          int task = task0 + u * kProducers;
          raw[u].kind = -1;
          if (task < ntask) {
            rr[u] = task / nchunk; cc[u] = task - rr[u] * nchunk;
            issue8(a, m0 + rr[u], cc[u] * 8, M, K, raw[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          if (raw[u].kind < 0) continue;
          float x[8];
          finish8(a, m0 + rr[u], cc[u] * 8, K, raw[u], x);
          size_t off = (size_t)cc[u] * kTileM * 16 + (size_t)rr[u] * 16;
          split_store8(x, a_hi + off, a_lo + off);
        }
      }
      if (eop_kind) cp_async_wait_all();
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }
  } else if (warp == 12) {
    // =============================== MMA issue ===============================
    const uint32_t idesc = make_idesc(npad);
    const uint32_t lbo_a = kTileM * 16, lbo_b = (uint32_t)npad * 16, sbo = 128;
    const uint64_t bdesc_hi = make_desc(smem_u32(w_hi), lbo_b, sbo);
    const uint64_t bdesc_lo = make_desc(smem_u32(w_lo), lbo_b, sbo);
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it % nstages;
      const uint32_t ph = (it / nstages) & 1;
      const int acc = it & 1;
      const uint32_t pa = (it >> 1) & 1;
      mbar_wait(&full[s], ph);
      mbar_wait(&tempty[acc], pa ^ 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_base = smem_u32(smem + L.a_stage0 + s * L.a_stage_bytes);
        const uint64_t adesc_hi = make_desc(a_base, lbo_a, sbo);
        const uint64_t adesc_lo = make_desc(a_base + kTileM * kpad * 2, lbo_a, sbo);
        const uint32_t d = tmem_base + (uint32_t)(acc * npad);
        for (int kb = 0; kb < nkb; ++kb) {
          // advance both operands by one 16-wide K block = two planes
          const uint64_t ka = (uint64_t)((2 * lbo_a * kb) >> 4), kbo = (uint64_t)((2 * lbo_b * kb) >> 4);
          umma_bf16(d, adesc_hi + ka, bdesc_hi + kbo, idesc, kb > 0 ? 1u : 0u);
          umma_bf16(d, adesc_lo + ka, bdesc_hi + kbo, idesc, 1u);
          umma_bf16(d, adesc_hi + ka, bdesc_lo + kbo, idesc, 1u);
        }
        umma_commit(&empty[s]);
        umma_commit(&tfull[acc]);
      }
      __syncwarp();
    }
  } else {
    // =============================== epilogue ===============================
    const int et = tid - kProducers;  // 0..127
    const int q = warp - 8;          // TMEM lane quadrant of this warp
    const int row = q * 32 + lane;   // accumulator row owned by this thread
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t pa = (it >> 1) & 1;
      const int m0 = tile * kTileM;
      mbar_wait(&tfull[acc], pa);
      tc_fence_after();
      const float* eop_tile = reinterpret_cast<const float*>(smem + L.eop + acc * L.eop_bytes);
      for (int c0 = 0; c0 < npad; c0 += kEpiCols) {
        float v[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * npad + c0);
        tmem_ld16(taddr, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) epi[row * (kEpiCols + 1) + i] = v[i];
        if (c0 + 16 < npad) {
          tmem_ld16(taddr + 16, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) epi[row * (kEpiCols + 1) + 16 + i] = v[i];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // cooperative finish + store with 16-byte vectors: thread -> (row er0 + 16*j, columns 4*eq..4*eq+3);
        // a warp covers four 128-byte row segments per access, EB accesses are in flight per thread
        const int eq = et & 7, er0 = et >> 3;          // 8 column quads x 16 row lanes
        const int n = c0 + eq * 4;
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        const bool vec_ok = (n + 4 <= N) && ((ep.ldc & 3) == 0) && al16(ep.C) &&
                            (!(ep.flags & (E_RELUMASK | E_STAT_XHAT)) || (((ep.ldh & 3) == 0) && al16(ep.hpre))) &&
                            (!(ep.flags & E_ROWBIAS) || (((ep.ldrb & 3) == 0) && al16(ep.rb))) &&
                            (!(ep.flags & E_GROUPADD) || (((ep.ldga & 3) == 0) && al16(ep.ga)));
        if (n < N) {
          float bias[4], scn[4], shn[4], mun[4], rsn[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ni = n + i < N ? n + i : N - 1;
            bias[i] = ep.bias ? ep.bias[ni] : 0.f;
            scn[i] = (ep.flags & E_RELUMASK) ? ep.scale[ni] : 0.f;
            shn[i] = (ep.flags & E_RELUMASK) ? ep.shift[ni] : 0.f;
            mun[i] = (ep.flags & E_STAT_XHAT) ? ep.mean[ni] : 0.f;
            rsn[i] = (ep.flags & E_STAT_XHAT) ? ep.rstd[ni] : 0.f;
          }
          constexpr int EB = 4;
          for (int j0 = 0; j0 < kTileM / 16; j0 += EB) {
            float4 xs[EB], hp[EB], old[EB];
            bool ok[EB];
#pragma unroll
            for (int u = 0; u < EB; ++u) {
              const int r = er0 + 16 * (j0 + u);
              const int m = m0 + r;
              ok[u] = m < M;
              hp[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              old[u] = hp[u];
              xs[u] = hp[u];
              if (!ok[u]) continue;
              const float* er = epi + r * (kEpiCols + 1) + eq * 4;
              float4 x = make_float4(er[0] + bias[0], er[1] + bias[1], er[2] + bias[2], er[3] + bias[3]);
              if (vec_ok) {
                if (ep.flags & E_ROWBIAS) {
                  float4 t4 = __ldg(reinterpret_cast<const float4*>(ep.rb + (size_t)(m / ep.rbT) * ep.ldrb + n));
                  x.x += t4.x; x.y += t4.y; x.z += t4.z; x.w += t4.w;
                }
                const float4* eo = reinterpret_cast<const float4*>(eop_tile + r * npad + n);
                if (ep.flags & E_GROUPADD) {
                  float4 t4;
                  if (eop_kind == 2) t4 = *eo;
                  else {
                    int b = m / ep.T, t = m - b * ep.T;
                    t4 = __ldg(reinterpret_cast<const float4*>(ep.ga + ((size_t)(b / ep.G) * ep.T + t) * ep.ldga + n));
                  }
                  x.x += t4.x; x.y += t4.y; x.z += t4.z; x.w += t4.w;
                }
                if (ep.flags & (E_RELUMASK | E_STAT_XHAT))
                  hp[u] = (eop_kind == 1) ? *eo : __ldg(reinterpret_cast<const float4*>(ep.hpre + (size_t)m * ep.ldh + n));
                if (ep.flags & E_ACCUM)
                  old[u] = (eop_kind == 3) ? *eo : *reinterpret_cast<const float4*>(ep.C + (size_t)m * ep.ldc + n);
              } else {
                float* xp = &x.x; float* hpp = &hp[u].x; float* op = &old[u].x;
                for (int i = 0; i < 4 && n + i < N; ++i) {
                  if (ep.flags & E_ROWBIAS) xp[i] += ep.rb[(size_t)(m / ep.rbT) * ep.ldrb + n + i];
                  if (ep.flags & E_GROUPADD) {
                    int b = m / ep.T, t = m - b * ep.T;
                    xp[i] += ep.ga[((size_t)(b / ep.G) * ep.T + t) * ep.ldga + n + i];
                  }
                  if (ep.flags & (E_RELUMASK | E_STAT_XHAT)) hpp[i] = ep.hpre[(size_t)m * ep.ldh + n + i];
                  if (ep.flags & E_ACCUM) op[i] = ep.C[(size_t)m * ep.ldc + n + i];
                }
              }
              xs[u] = x;
            }
#pragma unroll
            for (int u = 0; u < EB; ++u) {
              if (!ok[u]) continue;
              const int m = m0 + er0 + 16 * (j0 + u);
              float x[4] = {xs[u].x, xs[u].y, xs[u].z, xs[u].w};
              const float h[4] = {hp[u].x, hp[u].y, hp[u].z, hp[u].w};
              const float o[4] = {old[u].x, old[u].y, old[u].z, old[u].w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                if (ep.flags & E_RELUMASK) {
                  if (!(fmaf(h[i], scn[i], shn[i]) > 0.f)) x[i] = 0.f;
                }
                x[i] += o[i];
                if (STATS && n + i < N) {
                  s1[i] += x[i];
                  s2[i] += (ep.flags & E_STAT_XHAT) ? x[i] * ((h[i] - mun[i]) * rsn[i]) : x[i] * x[i];
                }
              }
              if (vec_ok) {
                *reinterpret_cast<float4*>(ep.C + (size_t)m * ep.ldc + n) = make_float4(x[0], x[1], x[2], x[3]);
              } else {
                for (int i = 0; i < 4 && n + i < N; ++i) ep.C[(size_t)m * ep.ldc + n + i] = x[i];
              }
            }
          }
        }
        if (STATS) {
          // the four row lanes of a warp that share a column quad are 8 lanes apart: two shuffles,
          // then one plain store per column and warp (shared fp32 atomics are CAS loops)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], 8);
            s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], 16);
            s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], 8);
            s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], 16);
          }
          if (lane < 8) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { sst[q][0][lane * 4 + i] = s1[i]; sst[q][1][lane * 4 + i] = s2[i]; }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (STATS && et < kEpiCols && c0 + et < N) {
          dacc[0][c0 + et] += (double)(sst[0][0][et] + sst[1][0][et] + sst[2][0][et] + sst[3][0][et]);
          dacc[1][c0 + et] += (double)(sst[0][1][et] + sst[1][1][et] + sst[2][1][et] + sst[3][1][et]);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
    }
    if (STATS) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int n = et; n < N; n += 128) {
        atomicAdd(ep.stat + n, dacc[0][n]);
        atomicAdd(ep.stat + N + n, dacc[1][n]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}


// ---------------------------------------------------------------------------------------------------
// Weight gradients on tensor cores: dW[K,N] += op_a(A)[M,K]^T . op_b(B)[M,N]   (K <= 120, N <= 240).
//
// The reduction runs over the long M axis.  Each 128-row slab of A and B is produced exactly as in
// tc_gemm_kernel (same physical shared-memory layout: planes of 8 columns, 16 bytes per row), but is
// now read by the UMMA as MN-major operands: "M" of the MMA is the K index of dW (padded to 128
// lanes), "N" is the N index, and the MMA's own K dimension walks the 128 rows 16 at a time.  The
// accumulator stays in TMEM across ALL slabs a CTA processes; there is no per-tile epilogue, only one
// TMEM read + fp32 atomicAdd per CTA at the end.  A constant-one column appended to A at index K
// yields the column sums of op_b(B) (bias gradients) in accumulator row K for free.
struct DwSmem {
  int a_bytes, b_bytes, stage_bytes, bars, total;
};
__host__ __device__ inline DwSmem dw_smem_layout(int npad, int nstages) {
  DwSmem s;
  s.a_bytes = 2 * 16 * kTileM * 16;           // hi + lo, 16 planes (128 MMA lanes) x 128 rows x 16 B
  s.b_bytes = 2 * (npad / 8) * kTileM * 16;   // hi + lo
  s.stage_bytes = s.a_bytes + s.b_bytes;
  s.bars = nstages * s.stage_bytes;
  s.total = s.bars + 128;
  return s;
}

constexpr int kDwThreads = 384 + 32;  // 12 producer warps (warps 0-3 also run the final epilogue) + MMA warp

__global__ void __launch_bounds__(kDwThreads, 1)
tc_dw_kernel(int M, int K, int N, int npad, int nstages, uint32_t tmem_cols, AOp a, AOp b,
             float* __restrict__ dW, int lddw, float* __restrict__ colsum) {
  extern __shared__ __align__(128) uint8_t smem[];
  const DwSmem L = dw_smem_layout(npad, nstages);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* full = bars;       // [2]
  uint64_t* empty = bars + 2;  // [2]
  uint64_t* done = bars + 4;   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = (M + kTileM - 1) / kTileM;
  const int ka = colsum ? K + 1 : K;           // A columns incl. the constant-one column
  const int nchunk_a = (ka + 7) >> 3, nchunk_b = npad >> 3;

  // zero the operand stages once: lanes >= ka of the A operand stay zero for the whole kernel
  for (int i = tid; i < (nstages * L.stage_bytes) / 16; i += kDwThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&full[i], 384); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) tmem_alloc(tmem_slot, tmem_cols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 12) {
    const int ntask_a = kTileM * nchunk_a, ntask = ntask_a + kTileM * nchunk_b;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it % nstages;
      const uint32_t ph = (it / nstages) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      uint8_t* a_hi = smem + s * L.stage_bytes;
      uint8_t* a_lo = a_hi + L.a_bytes / 2;
      uint8_t* b_hi = a_hi + L.a_bytes;
      uint8_t* b_lo = b_hi + L.b_bytes / 2;
      const int m0 = tile * kTileM;
      constexpr int UN = 4;
      for (int task0 = tid; task0 < ntask; task0 += 384 * UN) {
        Raw8 raw[UN];
        int rr[UN], cc[UN];
        bool isb[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          int task = task0 + u * 384;
          raw[u].kind = -1;
          if (task < ntask) {
            isb[u] = task >= ntask_a;
            int tt = isb[u] ? task - ntask_a : task;
            int nc = isb[u] ? nchunk_b : nchunk_a;
            rr[u] = tt / nc; cc[u] = tt - rr[u] * nc;
            if (isb[u]) issue8(b, m0 + rr[u], cc[u] * 8, M, N, raw[u]);
            else issue8(a, m0 + rr[u], cc[u] * 8, M, K, raw[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          if (raw[u].kind < 0) continue;
          float x[8];
          if (isb[u]) finish8(b, m0 + rr[u], cc[u] * 8, N, raw[u], x);
          else {
            finish8(a, m0 + rr[u], cc[u] * 8, K, raw[u], x);
            if (colsum && cc[u] * 8 <= K && K < cc[u] * 8 + 8) x[K - cc[u] * 8] = (m0 + rr[u] < M) ? 1.0f : 0.f;
          }
          size_t off = (size_t)cc[u] * kTileM * 16 + (size_t)rr[u] * 16;
          if (isb[u]) split_store8(x, b_hi + off, b_lo + off);
          else split_store8(x, a_hi + off, a_lo + off);
        }
      }
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }
  } else {
    // MMA issue: D[128 x npad] += A^T-slab . B-slab, both operands MN-major
    const uint32_t idesc = make_idesc(npad) | (1u << 15) | (1u << 16);
    const uint32_t lbo = 128, sbo = kTileM * 16;  // MN-major: LBO = 8-row group along the reduction, SBO = plane
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it % nstages;
      const uint32_t ph = (it / nstages) & 1;
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_base = smem_u32(smem + s * L.stage_bytes);
        const uint32_t b_base = a_base + L.a_bytes;
        const uint64_t a_hi = make_desc(a_base, lbo, sbo), a_lo = make_desc(a_base + L.a_bytes / 2, lbo, sbo);
        const uint64_t b_hi = make_desc(b_base, lbo, sbo), b_lo = make_desc(b_base + L.b_bytes / 2, lbo, sbo);
        for (int rb = 0; rb < kTileM / 16; ++rb) {
          const uint64_t adv = (uint64_t)((rb * 16 * 16) >> 4);  // 16 rows x 16 bytes
          umma_bf16(tmem_base, a_hi + adv, b_hi + adv, idesc, (it > 0 || rb > 0) ? 1u : 0u);
          umma_bf16(tmem_base, a_lo + adv, b_hi + adv, idesc, 1u);
          umma_bf16(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (lane == 0) umma_commit(done);
    __syncwarp();
  }
  if (warp < 4 && ntiles > (int)blockIdx.x) {
    // final epilogue: accumulator lane = dW row
    mbar_wait(done, 0);
    tc_fence_after();
    const int krow = warp * 32 + lane;
    for (int c0 = 0; c0 < npad; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      if (krow < K) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < N) atomicAdd(&dW[(size_t)krow * lddw + c0 + i], v[i]);
      } else if (colsum && krow == K) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < N) atomicAdd(&colsum[c0 + i], v[i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace tc
}  // namespace clsr
