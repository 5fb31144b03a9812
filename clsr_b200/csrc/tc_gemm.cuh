// Tensor-core linear layer for sm_100a: C[M,N] (+)= op_a(A)[M,K] . W[K,N] with the same fused
// operand prologues / epilogues as gemm.cuh, computed by tcgen05.mma with the accumulator in TMEM.
//
// Precision: every fp32 operand x is split into two bf16 terms x = hi + lo (hi = bf16(x),
// lo = bf16(x - hi)); a product is accumulated as hi*hi + lo*hi + hi*lo in fp32 (three
// tcgen05.mma.kind::f16 per 16-wide K block), i.e. ~2^-16 relative operand error.  A single bf16
// pass (2^-8) misses the 1e-3 logit parity bar of this path.
//
// These GEMMs are skinny (K, N <= 256 against M ~ 10^6 rows): they are bound by HBM traffic and, in
// practice, by how many instructions the SM spends per byte moved.  Structure (persistent CTAs,
// 16 warps, one CTA per SM):
//   warps 0-6  producers : 8-element operand pieces are read with one 32-byte load per thread
//                          (LDG.256), run through the prologue, split into bf16 hi/lo and written
//                          into the UMMA canonical K-major layout (no swizzle: planes of 8
//                          K-elements, 16 bytes per row).  Eight consecutive threads take eight
//                          consecutive rows of one plane, so shared-memory stores are
//                          conflict-free; the per-column prologue vectors live in shared memory.
//   warp  7    MMA       : one elected thread issues tcgen05.mma; tcgen05.commit releases the
//                          operand stage and publishes the accumulator.  The same thread prefetches
//                          the per-element epilogue operand (pre-activation for the ReLU mask /
//                          BatchNorm-backward statistic, group-broadcast rows, or old C) of the
//                          tile with cp.async.bulk straight into shared memory (mbarrier tx count).
//   warps 8-15 epilogue  : two warps per TMEM lane quadrant, each owning every other 16-column
//                          chunk.  One thread = one output row: tcgen05.ld, epilogue in registers,
//                          32-byte stores; BatchNorm column statistics by a 16-shuffle transpose
//                          reduction per chunk, accumulated in double per CTA.
// Two operand stages and two TMEM accumulators are ring-buffered with mbarriers so the producer
// of tile i+1, the MMA of tile i and the epilogue of tile i-1 overlap.  W (K x N, small) is split
// once per CTA and stays resident in shared memory.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include <type_traits>

#include "gemm.cuh"

namespace clsr {
namespace tc {

#ifndef CLSR_TC_UN1
#define CLSR_TC_UN1 8   // pieces in flight per producer thread, one-stream operands (32 bytes each)
#endif
#ifndef CLSR_TC_UN2
#define CLSR_TC_UN2 4   // two-stream operands (64 bytes each)
#endif
#ifndef CLSR_TC_PREFETCH
#define CLSR_TC_PREFETCH 3   // tiles of L2 prefetch distance (0 = off)
#endif

constexpr int kTileM = 128;
constexpr int kProducers = 192;                         // warps 0-5
constexpr int kLoadWarp = 6;                            // TMA operand loads / L2 prefetch
constexpr int kMmaWarp = 7;                             // MMA issue + epilogue-operand bulk copies
constexpr int kPlaneBytes = 4096;                       // one operand plane: 128 rows x (16 B hi + 16 B lo)
constexpr int kEpilogue = 256;                          // warps 8-15
constexpr int kThreads = 512;                           // 16 warps x 128 registers fill the register file

CLSR_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

CLSR_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
CLSR_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
CLSR_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
CLSR_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "CLSR_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra CLSR_WAIT_DONE;\n"
      "bra CLSR_WAIT_LOOP;\n"
      "CLSR_WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)   // suspend-time hint: park the warp instead of spinning on issue slots
      : "memory");
}
// 1-D bulk copy global -> shared; completion is counted in bytes on an mbarrier.
CLSR_DEVINL void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 2-D tiled TMA load: box {8 columns, 128 rows} of an fp32 matrix = one operand plane's raw 4 KB
CLSR_DEVINL void tma_load_plane(uint32_t smem_dst, const CUtensorMap* tm, int col, int row, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(tm), "r"(smem_u32(bar)), "r"(col), "r"(row)
      : "memory");
}
// L2 prefetch of a contiguous byte range (no shared-memory destination, no completion tracking)
CLSR_DEVINL void l2_prefetch(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
// 2-D tiled TMA store: one plane (box {8 columns, 128 rows}) of the finished tile from shared to global memory;
// rows past the end of the matrix are clipped by the tensor map.
CLSR_DEVINL void tma_store_plane(const CUtensorMap* tm, uint32_t smem_src, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(smem_src),
               "r"(col), "r"(row)
               : "memory");
}
CLSR_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
CLSR_DEVINL void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
CLSR_DEVINL void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
CLSR_DEVINL void sts4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
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

// Three 16-column loads (accumulator chunk + the two per-thread statistic accumulators) and one wait.
CLSR_DEVINL void tmem_ld16x3(uint32_t t0, uint32_t t1, uint32_t t2, float* a, float* b, float* c) {
  uint32_t r[48];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%48];\n"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%49];\n"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47}, [%50];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
        "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47])
      : "r"(t0), "r"(t1), "r"(t2)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[16 + i]); c[i] = __uint_as_float(r[32 + i]); }
}
CLSR_DEVINL void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
CLSR_DEVINL void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

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

// Operand stage layout: plane p (8 K-elements) = 16 row-octet blocks of 256 bytes; a block holds the UMMA
// core matrix of the bf16 hi parts (8 rows x 16 B) followed by the core matrix of the lo parts.  The raw
// fp32 data of the same 8 rows x 8 elements is also exactly 256 bytes (8 x 32 B), which is what lets a TMA
// box {8 columns, 128 rows} land in the plane and be converted in place.
CLSR_DEVINL uint32_t piece_off(int plane, int r) { return (uint32_t)(plane * kPlaneBytes + (r >> 3) * 256 + (r & 7) * 16); }

// x[0..7] -> 8 bf16 "hi" at dst, 8 bf16 "lo" at dst + 128 (both round-to-nearest), 16-byte shared stores
// (32-bit shared-window addresses: no generic-pointer arithmetic in the producer loops).
CLSR_DEVINL void split_store8(const float* x, uint32_t dst) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 hb = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    const uint32_t hu = *reinterpret_cast<uint32_t*>(&hb);
    const float h0 = __uint_as_float(hu << 16), h1 = __uint_as_float(hu & 0xffff0000u);
    __nv_bfloat162 lb = __floats2bfloat162_rn(x[2 * i] - h0, x[2 * i + 1] - h1);
    h[i] = hu;
    l[i] = *reinterpret_cast<uint32_t*>(&lb);
  }
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
  asm volatile("st.shared.v4.b32 [%0+128], {%1,%2,%3,%4};" ::"r"(dst), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
}
// W (the small resident operand) keeps separate hi / lo arrays
CLSR_DEVINL void split_store8_w(const float* x, uint32_t hi_dst, uint32_t lo_dst) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 hb = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    const uint32_t hu = *reinterpret_cast<uint32_t*>(&hb);
    const float h0 = __uint_as_float(hu << 16), h1 = __uint_as_float(hu & 0xffff0000u);
    __nv_bfloat162 lb = __floats2bfloat162_rn(x[2 * i] - h0, x[2 * i + 1] - h1);
    h[i] = hu;
    l[i] = *reinterpret_cast<uint32_t*>(&lb);
  }
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(hi_dst), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(lo_dst), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
}
CLSR_DEVINL float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

CLSR_DEVINL bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
CLSR_DEVINL bool al32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }

struct F8 {
  float4 lo, hi;
};
// eight consecutive floats: one 32-byte load when the address allows it, else two 16-byte loads
template <bool V256>
CLSR_DEVINL F8 ld8(const float* p) {
  F8 r;
  if (V256) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z),
                   "=f"(r.hi.w)
                 : "l"(p));
  } else {
    r.lo = __ldg(reinterpret_cast<const float4*>(p));
    r.hi = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  return r;
}
CLSR_DEVINL void st8(float* p, const float* x, bool v256) {
  if (v256) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(x[0]), "f"(x[1]), "f"(x[2]),
                 "f"(x[3]), "f"(x[4]), "f"(x[5]), "f"(x[6]), "f"(x[7])
                 : "memory");
  } else {
    *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(x[4], x[5], x[6], x[7]);
  }
}

// ---------------------------------------------------------------------------------------------------
// Operand producers.  A "piece" is eight consecutive elements A'[m, k0..k0+7] = one 16-byte row of one
// canonical plane.  The fast path needs every operand row 16-byte aligned (32 for LDG.256) and the
// piece not to straddle K or a concatenation boundary; everything else goes through AOp::load.
struct Fast {
  bool ok, v256;
};
CLSR_DEVINL Fast fast_eligible(const AOp& a) {
  Fast f;
  f.ok = false; f.v256 = false;
  const bool a4 = (a.lda & 3) == 0 && al16(a.A), a8 = (a.lda & 7) == 0 && al32(a.A);
  const bool b4 = (a.lda2 & 3) == 0 && al16(a.A2), b8 = (a.lda2 & 7) == 0 && al32(a.A2);
  switch (a.mode) {
    case A_PLAIN:
    case A_BNRELU:
      f.ok = a4; f.v256 = a8; break;
    case A_AFFINE2:
      f.ok = a4 && b4; f.v256 = a8 && b8; break;
    case A_CATMUL:
      f.ok = a4 && b4 && (a.W1 & 7) == 0 && (a.off & 3) == 0; f.v256 = a8 && b8 && (a.off & 7) == 0; break;
    case A_MULROW:
      f.ok = a4 && b4 && (a.off & 3) == 0; f.v256 = a8 && b8 && (a.off & 7) == 0; break;
    case A_CAT2ROW:
      f.ok = a4 && b4 && (a.W1 & 7) == 0; f.v256 = a8 && b8; break;
    default: break;
  }
  f.v256 = f.v256 && f.ok;
  return f;
}

// Loads of one piece (row m, columns k0..k0+7).  `second` selects the second half of a concatenated
// operand (thread-constant: every thread keeps one plane for the whole tile).
template <int MODE, bool V256>
CLSR_DEVINL void load_piece(const AOp& a, int m, int k0, bool second, F8& ra, F8& rb) {
  if (MODE == A_PLAIN || MODE == A_BNRELU) {
    ra = ld8<V256>(a.A + (size_t)m * a.lda + k0);
  } else if (MODE == A_AFFINE2) {
    ra = ld8<V256>(a.A + (size_t)m * a.lda + k0);
    rb = ld8<V256>(a.A2 + (size_t)m * a.lda2 + k0);
  } else if (MODE == A_CATMUL) {
    if (!second) {
      ra = ld8<V256>(a.A + (size_t)m * a.lda + k0);
      rb = ra;
    } else {
      const int kk = k0 - a.W1;
      ra = ld8<V256>(a.A + (size_t)m * a.lda + a.off + kk);
      rb = ld8<V256>(a.A2 + (size_t)(m / a.T) * a.lda2 + kk);
    }
  } else if (MODE == A_MULROW) {
    const int b = m / a.T, t = m - b * a.T, sq = b / a.G;
    ra = ld8<V256>(a.A + ((size_t)sq * a.T + t) * a.lda + a.off + k0);
    rb = ld8<V256>(a.A2 + (size_t)b * a.lda2 + k0);
  } else {  // A_CAT2ROW
    if (!second) ra = ld8<V256>(a.A + (size_t)(m / a.G) * a.lda + k0);
    else ra = ld8<V256>(a.A2 + (size_t)m * a.lda2 + (k0 - a.W1));
  }
}

// Fast pieces of one 128-row tile.  Octet o of eight consecutive threads owns plane o % nfull and the
// row octets o / nfull, + per, + 2 per, ... (per = octets per plane): the plane, its prologue vectors
// and the address strides are loop constants, eight consecutive threads write eight consecutive rows
// of one plane (conflict-free 128-byte store) and the four octets of a warp read neighbouring
// 32-byte pieces of the same rows.  Loads are always issued from a clamped row (a predicated load
// would keep its destination registers live across the whole kernel).
template <int MODE, bool V256>
CLSR_DEVINL void produce_fast(const AOp& a, const float* sv, int svld, int m0, int M, int nfull, int ptid, int nprod,
                              uint32_t base) {
  const int noct = nprod >> 3;
  const int per = noct / nfull;
  const int o = ptid >> 3, rl = ptid & 7;
  const int cc = o % nfull, ro0 = o / nfull;
  if (ro0 >= per) return;
  const int k0 = cc * 8;
  const bool full = m0 + kTileM <= M;
  const bool second = (MODE == A_CATMUL || MODE == A_CAT2ROW) ? (k0 >= a.W1) : false;
  float c0[8], c1[8], c2[8];
  if (MODE == A_BNRELU || MODE == A_AFFINE2) {
    const float4 p0 = *reinterpret_cast<const float4*>(sv + k0), p1 = *reinterpret_cast<const float4*>(sv + k0 + 4);
    const float4 q0 = *reinterpret_cast<const float4*>(sv + svld + k0), q1 = *reinterpret_cast<const float4*>(sv + svld + k0 + 4);
    c0[0] = p0.x; c0[1] = p0.y; c0[2] = p0.z; c0[3] = p0.w; c0[4] = p1.x; c0[5] = p1.y; c0[6] = p1.z; c0[7] = p1.w;
    c1[0] = q0.x; c1[1] = q0.y; c1[2] = q0.z; c1[3] = q0.w; c1[4] = q1.x; c1[5] = q1.y; c1[6] = q1.z; c1[7] = q1.w;
  }
  if (MODE == A_AFFINE2) {
    const float4 u0 = *reinterpret_cast<const float4*>(sv + 2 * svld + k0), u1 = *reinterpret_cast<const float4*>(sv + 2 * svld + k0 + 4);
    c2[0] = u0.x; c2[1] = u0.y; c2[2] = u0.z; c2[3] = u0.w; c2[4] = u1.x; c2[5] = u1.y; c2[6] = u1.z; c2[7] = u1.w;
  }
  const int rstep = per * 8;
  const uint32_t sstep = (uint32_t)per * 256;
  uint32_t soff = base + (uint32_t)(cc * kPlaneBytes + ro0 * 256 + rl * 16);
  // pieces in flight per thread: UN1 x 32 bytes (one stream) or UN2 x 64 bytes (two streams)
  constexpr int UN = (MODE == A_AFFINE2 || MODE == A_CATMUL || MODE == A_MULROW) ? CLSR_TC_UN2 : CLSR_TC_UN1;
#pragma unroll 1
  for (int r = ro0 * 8 + rl; r < kTileM; r += rstep * UN, soff += sstep * UN) {
    F8 ra[UN], rb[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      int m = m0 + r + u * rstep;
      m = m < M ? m : M - 1;
      load_piece<MODE, V256>(a, m, k0, second, ra[u], rb[u]);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int rr = r + u * rstep;
      if (rr >= kTileM) continue;
      const float v[8] = {ra[u].lo.x, ra[u].lo.y, ra[u].lo.z, ra[u].lo.w, ra[u].hi.x, ra[u].hi.y, ra[u].hi.z, ra[u].hi.w};
      float x[8];
      if (MODE == A_PLAIN || MODE == A_CAT2ROW) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = v[i];
      } else if (MODE == A_BNRELU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaxf(0.f, fmaf(v[i], c0[i], c1[i]));
      } else {
        const float h[8] = {rb[u].lo.x, rb[u].lo.y, rb[u].lo.z, rb[u].lo.w, rb[u].hi.x, rb[u].hi.y, rb[u].hi.z, rb[u].hi.w};
        if (MODE == A_AFFINE2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = fmaf(c0[i], v[i], fmaf(c1[i], h[i], c2[i]));
        } else if (MODE == A_CATMUL) {
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = second ? v[i] * h[i] : v[i];
        } else {  // A_MULROW
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = v[i] * h[i];
        }
      }
      if (!full && m0 + rr >= M) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = 0.f;
      }
      split_store8(x, soff + u * sstep);
    }
  }
}

// In-place conversion of planes that the TMA delivered as raw fp32 (modes whose rows are plain matrix
// rows: A_PLAIN, A_BNRELU, A_AFFINE2 with the second stream in raw2).  Same plane ownership as
// produce_fast.  The eight threads of an octet read their block's 256 raw bytes, synchronise among
// themselves, then overwrite the block with its hi / lo core matrices.
// A_CATMUL: the loader delivers columns [0, W1) of the rows and, a second time, columns [off, off + K - W1) into the
// planes of the second part; those are multiplied here by the row of the small per-sequence operand (L1 / L2 resident).
template <int MODE>
CLSR_DEVINL void convert_tile(const AOp& a, const float* sv, int svld, int m0, int M, int nfull, int ptid, int nprod,
                              uint32_t base, uint32_t raw2) {
  const int noct = nprod >> 3;
  const int per = noct / nfull;
  const int o = ptid >> 3, rl = ptid & 7;
  const int cc = o % nfull, ro0 = o / nfull;
  if (ro0 >= per) return;
  const int k0 = cc * 8;
  const bool full = m0 + kTileM <= M;
  const unsigned omask = 0xFFu << (8 * ((threadIdx.x & 31) >> 3));
  float c0[8], c1[8], c2[8];
  if (MODE == A_BNRELU || MODE == A_AFFINE2) {
    const float4 p0 = *reinterpret_cast<const float4*>(sv + k0), p1 = *reinterpret_cast<const float4*>(sv + k0 + 4);
    const float4 q0 = *reinterpret_cast<const float4*>(sv + svld + k0), q1 = *reinterpret_cast<const float4*>(sv + svld + k0 + 4);
    c0[0] = p0.x; c0[1] = p0.y; c0[2] = p0.z; c0[3] = p0.w; c0[4] = p1.x; c0[5] = p1.y; c0[6] = p1.z; c0[7] = p1.w;
    c1[0] = q0.x; c1[1] = q0.y; c1[2] = q0.z; c1[3] = q0.w; c1[4] = q1.x; c1[5] = q1.y; c1[6] = q1.z; c1[7] = q1.w;
  }
  if (MODE == A_AFFINE2) {
    const float4 u0 = *reinterpret_cast<const float4*>(sv + 2 * svld + k0), u1 = *reinterpret_cast<const float4*>(sv + 2 * svld + k0 + 4);
    c2[0] = u0.x; c2[1] = u0.y; c2[2] = u0.z; c2[3] = u0.w; c2[4] = u1.x; c2[5] = u1.y; c2[6] = u1.z; c2[7] = u1.w;
  }
  const uint32_t step = (uint32_t)per * 256;
  uint32_t off = (uint32_t)(cc * kPlaneBytes + ro0 * 256);
  constexpr int UN = 2;
  const bool second = MODE == A_CATMUL && k0 >= a.W1;   // thread-constant: one plane per octet
#pragma unroll 1
  for (int ro = ro0; ro < 16; ro += per * UN, off += step * UN) {
    float4 a0[UN], a1[UN], b0[UN], b1[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      // the second block of the last iteration may lie past the plane: read block `ro` again instead
      const uint32_t o2 = off + (ro + u * per < 16 ? u * step : 0) + rl * 32;
      a0[u] = lds4(base + o2); a1[u] = lds4(base + o2 + 16);
      if (MODE == A_AFFINE2) { b0[u] = lds4(raw2 + o2); b1[u] = lds4(raw2 + o2 + 16); }
      if (MODE == A_CATMUL) {
        if (second) {
          int m = m0 + (ro + (ro + u * per < 16 ? u * per : 0)) * 8 + rl;
          m = m < M ? m : M - 1;
          const float4* hp = reinterpret_cast<const float4*>(a.A2 + (size_t)(m / a.T) * a.lda2 + (k0 - a.W1));
          b0[u] = __ldg(hp); b1[u] = __ldg(hp + 1);
        } else {
          b0[u] = make_float4(1.f, 1.f, 1.f, 1.f); b1[u] = b0[u];
        }
      }
    }
    __syncwarp(omask);
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (ro + u * per >= 16) continue;
      const float v[8] = {a0[u].x, a0[u].y, a0[u].z, a0[u].w, a1[u].x, a1[u].y, a1[u].z, a1[u].w};
      float x[8];
      if (MODE == A_PLAIN) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = v[i];
      } else if (MODE == A_BNRELU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaxf(0.f, fmaf(v[i], c0[i], c1[i]));
      } else if (MODE == A_CATMUL) {
        const float h[8] = {b0[u].x, b0[u].y, b0[u].z, b0[u].w, b1[u].x, b1[u].y, b1[u].z, b1[u].w};
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = second ? v[i] * h[i] : v[i];   // (same expression as the register path)
      } else {
        const float h[8] = {b0[u].x, b0[u].y, b0[u].z, b0[u].w, b1[u].x, b1[u].y, b1[u].z, b1[u].w};
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaf(c0[i], v[i], fmaf(c1[i], h[i], c2[i]));
      }
      if (!full && m0 + (ro + u * per) * 8 + rl >= M) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = 0.f;
      }
      split_store8(x, base + off + u * step + rl * 16);
    }
  }
}

// Element-wise pieces: planes [c_lo, c_hi) (K edge, unaligned operands, the constant-one column).
CLSR_DEVINL void produce_slow(const AOp& a, int m0, int M, int K, int c_lo, int c_hi, int one_col, int ptid,
                                          int nprod, uint32_t base) {
  const int ntask = kTileM * (c_hi - c_lo);
#pragma unroll 1
  for (int task = ptid; task < ntask; task += nprod) {
    const int r = task & (kTileM - 1), cc = c_lo + (task >> 7);
    const int m = m0 + r;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = cc * 8 + i;
      x[i] = (m < M && k < K) ? a.load(m, k) : 0.f;
      if (k == one_col && m < M) x[i] = 1.0f;
    }
    split_store8(x, base + piece_off(cc, r));
  }
}

// One operand tile (rows m0..m0+127, columns [0,K) and optionally a constant-one column at one_col).
template <int MODE>
CLSR_DEVINL void produce_fast_v(const AOp& a, bool v256, const float* sv, int svld, int m0, int M, int nfull, int ptid,
                                int nprod, uint32_t base) {
  if (v256) produce_fast<MODE, true>(a, sv, svld, m0, M, nfull, ptid, nprod, base);
  else produce_fast<MODE, false>(a, sv, svld, m0, M, nfull, ptid, nprod, base);
}
// tma: 0 = the producers load the operand themselves; 1 / 2 = the loader warp's TMA has put the raw rows of
// one / two streams into the stage (and raw2) and the producers only convert them.
CLSR_DEVINL void produce_tile(const AOp& a, Fast f, int tma, const float* sv, int svld, int m0, int M, int K, int one_col,
                              int ptid, int nprod, uint32_t base, uint32_t raw2) {
  int nfull = (f.ok || tma) ? (K >> 3) : 0;
  if (nfull > (nprod >> 3)) nfull = 0;   // more planes than thread octets: element-wise path
  const int kall = one_col >= K ? one_col + 1 : K;
  const int nall = (kall + 7) >> 3;
  if (nfull > 0 && tma) {
    if (a.mode == A_PLAIN || a.mode == A_CAT2ROW) convert_tile<A_PLAIN>(a, sv, svld, m0, M, nfull, ptid, nprod, base, raw2);
    else if (a.mode == A_BNRELU) convert_tile<A_BNRELU>(a, sv, svld, m0, M, nfull, ptid, nprod, base, raw2);
    else if (a.mode == A_CATMUL) convert_tile<A_CATMUL>(a, sv, svld, m0, M, nfull, ptid, nprod, base, raw2);
    else convert_tile<A_AFFINE2>(a, sv, svld, m0, M, nfull, ptid, nprod, base, raw2);
  } else if (nfull > 0) switch (a.mode) {
    case A_PLAIN: produce_fast_v<A_PLAIN>(a, f.v256, sv, svld, m0, M, nfull, ptid, nprod, base); break;
    case A_BNRELU: produce_fast_v<A_BNRELU>(a, f.v256, sv, svld, m0, M, nfull, ptid, nprod, base); break;
    case A_AFFINE2: produce_fast_v<A_AFFINE2>(a, f.v256, sv, svld, m0, M, nfull, ptid, nprod, base); break;
    case A_CATMUL: produce_fast_v<A_CATMUL>(a, f.v256, sv, svld, m0, M, nfull, ptid, nprod, base); break;
    case A_MULROW: produce_fast_v<A_MULROW>(a, f.v256, sv, svld, m0, M, nfull, ptid, nprod, base); break;
    default: produce_fast_v<A_CAT2ROW>(a, f.v256, sv, svld, m0, M, nfull, ptid, nprod, base); break;
  }
  if (nall > nfull) produce_slow(a, m0, M, K, nfull, nall, one_col, ptid, nprod, base);
}

// Loader warp: TMA loads of the raw planes of one tile (tma = number of streams), all arriving on `bar`.
// (A_CATMUL: the planes of the second part re-load columns [off, ...) of the same rows)
CLSR_DEVINL void tma_issue_tile(const AOp& a, int tma, const CUtensorMap* t1, const CUtensorMap* t2, int nplanes, int m0,
                                uint32_t base, uint32_t raw2, uint64_t* bar, int lane, int col0 = 0) {
  if (a.mode == A_CAT2ROW) {   // [A | A2] (G == 1): the planes of the second part come from the second matrix
    for (int p = lane; p < nplanes; p += 32) {
      if (p * 8 < a.W1) tma_load_plane(base + p * kPlaneBytes, t1, p * 8, m0, bar);
      else tma_load_plane(base + p * kPlaneBytes, t2, p * 8 - a.W1, m0, bar);
    }
    return;
  }
  const int w1 = a.mode == A_CATMUL ? a.W1 : (1 << 30);
  for (int p = lane; p < nplanes; p += 32) {
    const int col = p * 8 < w1 ? p * 8 : a.off + p * 8 - w1;
    tma_load_plane(base + p * kPlaneBytes, t1, col0 + col, m0, bar);
    if (tma == 2) tma_load_plane(raw2 + p * kPlaneBytes, t2, col0 + p * 8, m0, bar);
  }
}

// L2 prefetch of the HBM-streamed rows of one operand tile, issued by one warp a few tiles ahead of the
// producers so that their register loads see L2 latency instead of HBM latency.
CLSR_DEVINL void prefetch_rows(const float* base, int ld, int m0, int rows, int kfloats, int lane) {
  if (ld == kfloats) {
    if (lane == 0) l2_prefetch(base + (size_t)m0 * ld, (uint32_t)rows * kfloats * 4);
  } else {
    for (int r = lane; r < rows; r += 32) l2_prefetch(base + (size_t)(m0 + r) * ld, (uint32_t)kfloats * 4);
  }
}
CLSR_DEVINL void prefetch_operand(const AOp& a, int m0, int M, int K, int lane) {
  if (m0 >= M || (K & 3)) return;
  const int rows = M - m0 < kTileM ? M - m0 : kTileM;
  const bool a4 = (a.lda & 3) == 0 && (reinterpret_cast<uintptr_t>(a.A) & 15) == 0;
  switch (a.mode) {
    case A_PLAIN:
    case A_BNRELU:
      if (a4) prefetch_rows(a.A, a.lda, m0, rows, K, lane);
      break;
    case A_AFFINE2:
      if (a4) prefetch_rows(a.A, a.lda, m0, rows, K, lane);
      if ((a.lda2 & 3) == 0 && (reinterpret_cast<uintptr_t>(a.A2) & 15) == 0) prefetch_rows(a.A2, a.lda2, m0, rows, K, lane);
      break;
    case A_CATMUL:
      if (a4) prefetch_rows(a.A, a.lda, m0, rows, a.lda, lane);
      break;
    default: break;  // A_MULROW / A_CAT2ROW read small, L2-resident operands
  }
}

// copy the per-column prologue vectors of an operand into shared memory (sv[j*svld + k])
CLSR_DEVINL void stage_vectors(const AOp& a, float* sv, int svld, int K, int tid, int nthreads) {
  const int nv = a.mode == A_BNRELU ? 2 : (a.mode == A_AFFINE2 ? 3 : 0);
  for (int i = tid; i < nv * svld; i += nthreads) {
    const int j = i / svld, k = i - j * svld;
    const float* src = j == 0 ? a.v0 : (j == 1 ? a.v1 : a.v2);
    sv[i] = k < K ? src[k] : 0.f;
  }
}

// Column sums over the 32 lanes of a warp for 16 columns held one row per lane: recursive halving,
// 16 shuffles.  Returns the total of column colmap16(lane) (both lanes of a pair hold it).
CLSR_DEVINL float col_reduce16(const float* v, int lane) {
  float w8[8], w4[4], w2[2];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = b4 ? v[i] : v[i + 8], keep = b4 ? v[i + 8] : v[i];
    w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b3 ? w8[i] : w8[i + 4], keep = b3 ? w8[i + 4] : w8[i];
    w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b2 ? w4[i] : w4[i + 2], keep = b2 ? w4[i + 2] : w4[i];
    w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = b1 ? w2[0] : w2[1], keep = b1 ? w2[1] : w2[0];
  float w1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
  return w1;
}
CLSR_DEVINL int colmap16(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }

struct Smem {
  // byte offsets into dynamic shared memory
  int w_hi, w_lo, a_stage0, a_stage_bytes, raw2, raw2_bytes, eop, eop_bytes, cst, cst_bytes, vec, dstat, bars, total;
};
// eop != 0 reserves two [128 x N] fp32 tiles for the prefetched epilogue operand; tma == 2 reserves one raw
// buffer per stage for the second operand stream.
// tstore != 0 reserves two [N / 8 planes x 4 KB] images of the output tile for the TMA-store epilogue (the store of
// tile i reads one while the epilogue of tile i+1 fills the other).
// kchunks > 1: the contraction runs over kchunks chunks of K columns each (all of W resident, one operand stage per chunk).
__host__ __device__ inline Smem smem_layout(int kpad, int npad, int N, int nstages, int eop, int stats, int tma,
                                            int tstore = 0, int kchunks = 1) {
  Smem s;
  const int wbytes = kchunks * kpad * npad * 2;
  s.w_hi = 0;
  s.w_lo = wbytes;
  s.a_stage0 = 2 * wbytes;
  s.a_stage_bytes = (kpad / 8) * kPlaneBytes;
  s.raw2 = s.a_stage0 + nstages * s.a_stage_bytes;
  s.raw2_bytes = tma == 2 ? s.a_stage_bytes : 0;
  s.eop = s.raw2 + nstages * s.raw2_bytes;
  s.eop_bytes = eop ? ((kTileM * N * 4 + 127) & ~127) : 0;
  s.cst = s.eop + 2 * s.eop_bytes;
  s.cst_bytes = tstore ? (N >> 3) * kPlaneBytes : 0;
  s.vec = s.cst + 2 * s.cst_bytes;
  s.dstat = s.vec + (3 * kpad + 5 * npad) * 4;   // prologue vectors + bias / scale / shift / mean / rstd
  s.dstat = (s.dstat + 15) & ~15;
  s.bars = s.dstat + (stats ? 4 * 2 * npad * 8 : 0);
  s.total = s.bars + 192;
  return s;
}

// STATS: accumulate per-column statistics into ep.stat (see gemm.cuh).
template <bool STATS>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(int M, int N, int K, int kpad, int npad, int nstages, uint32_t tmem_cols, int eop_kind, int tma, int tstore,
               int kchunks, AOp a, const float* __restrict__ W, int ldw, EpiOp ep, const __grid_constant__ CUtensorMap tmA,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmC) {
  // kchunks > 1 (plain operands only): C = sum over kchunks chunks of K columns of A, rows [kc*K, (kc+1)*K) of W; every
  // chunk is one operand stage, the accumulator stays in TMEM across the chunks of a tile (one launch instead of
  // kchunks launches that re-read and re-write C).
  // tstore: the epilogue writes the finished tile into a shared-memory image (same plane layout as the operand
  // stages: 8 columns x 128 rows = 4 KB) and one thread stores it with 2-D TMA (one box per plane) instead of every
  // thread storing 32-byte pieces of its own row (one L1 wavefront per row and instruction).
  // tma: 0 = producers read the operand with register loads; 1 / 2 = the loader warp brings the raw rows of
  // one / two operand streams in with 2-D TMA (one box per plane) and the producers convert in place.
  // eop_kind: which per-element epilogue operand is prefetched into shared memory by bulk copy
  // (0 none, 1 hpre, 2 group-add rows, 3 old C for accumulation); the host only selects one when the
  // operand rows are contiguous (leading dimension == N).
  extern __shared__ __align__(128) uint8_t smem[];
  const Smem L = smem_layout(kpad, npad, N, nstages, eop_kind, STATS ? 1 : 0, tma, tstore, kchunks);
  uint8_t* w_hi = smem + L.w_hi;
  uint8_t* w_lo = smem + L.w_lo;
  float* sv = reinterpret_cast<float*>(smem + L.vec);     // [3][kpad]
  float* sbias = sv + 3 * kpad;                           // [npad] each
  float* sscale = sbias + npad;
  float* sshift = sscale + npad;
  float* smr = sshift + npad;                             // -mean * rstd
  float* srstd = smr + npad;
  double* dstat = reinterpret_cast<double*>(smem + L.dstat);  // [4 quadrants][2][npad]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* full = bars;          // [4] operand stages (up to four)
  uint64_t* empty = bars + 4;     // [4]
  uint64_t* rawfull = bars + 8;   // [4] raw operand planes landed (TMA)
  uint64_t* tfull = bars + 12;    // [2] accumulators
  uint64_t* tempty = bars + 14;   // [2]
  uint64_t* efull = bars + 16;    // [2] epilogue-operand tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = (M + kTileM - 1) / kTileM;
  const int nkb = kpad >> 4;
  const Fast fa = fast_eligible(a);

  // ---- one-time setup: W split into canonical hi/lo tiles, vectors, barriers, TMEM ----
  for (int idx = tid; idx < kchunks * (kpad >> 3) * npad; idx += kThreads) {
    const int ct = idx / npad, n = idx - ct * npad;   // plane ct = (chunk, plane of the chunk) holds 8 consecutive k for every n
    const int kc = ct / (kpad >> 3), c = ct - kc * (kpad >> 3);
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int k = c * 8 + i;
      x[i] = (k < K && n < N) ? W[(size_t)(kc * K + k) * ldw + n] : 0.f;
    }
    const uint32_t off = (uint32_t)(ct * npad * 16 + n * 16);
    split_store8_w(x, smem_u32(w_hi) + off, smem_u32(w_lo) + off);
  }
  // operand stages start as zeros: planes past ceil(K/8) are never written again
  for (int i = tid; i < (nstages * L.a_stage_bytes) / 16; i += kThreads)
    reinterpret_cast<uint4*>(smem + L.a_stage0)[i] = make_uint4(0u, 0u, 0u, 0u);
  stage_vectors(a, sv, kpad, K, tid, kThreads);
  for (int n = tid; n < npad; n += kThreads) {
    const bool in = n < N;
    sbias[n] = (in && ep.bias) ? ep.bias[n] : 0.f;
    sscale[n] = (in && (ep.flags & E_RELUMASK)) ? ep.scale[n] : 0.f;
    sshift[n] = (in && (ep.flags & E_RELUMASK)) ? ep.shift[n] : 0.f;
    const float rs = (in && (ep.flags & E_STAT_XHAT)) ? ep.rstd[n] : 0.f;
    srstd[n] = rs;
    smr[n] = (in && (ep.flags & E_STAT_XHAT)) ? -ep.mean[n] * rs : 0.f;
  }
  if (STATS) {
    for (int i = tid; i < 4 * 2 * npad; i += kThreads) dstat[i] = 0.0;
  }
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&full[i], kProducers);
      mbar_init(&empty[i], 1);
      mbar_init(&rawfull[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], kEpilogue / 32);
      mbar_init(&efull[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, tmem_cols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kLoadWarp) {
    // =============================== producers / converters ===============================
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
      for (int kc = 0; kc < kchunks; ++kc, ++it) {
        const int s = it % nstages;
        const uint32_t ph = (it / nstages) & 1;
        AOp ak = a;
        ak.A = a.A + kc * K;   // (kchunks > 1: plain operand)
        if (tma) mbar_wait(&rawfull[s], ph);      // implies empty[s]: the loader waited for it
        else {
          mbar_wait(&empty[s], ph ^ 1);
          // register-load path: warp 0 prefetches the rows of a later tile into L2 (paced by the pipeline)
          if (CLSR_TC_PREFETCH && warp == 0)
            prefetch_operand(ak, (tile + CLSR_TC_PREFETCH * (int)gridDim.x) * kTileM, M, K, lane);
        }
        const uint32_t a_base = smem_u32(smem + L.a_stage0 + s * L.a_stage_bytes);
        const uint32_t a_raw2 = smem_u32(smem + L.raw2 + s * L.raw2_bytes);
        produce_tile(ak, fa, tma, sv, kpad, tile * kTileM, M, K, -1, tid, kProducers, a_base, a_raw2);
        fence_proxy_async();
        mbar_arrive(&full[s]);
      }
  } else if (warp == kLoadWarp) {
    // =============================== operand loader ===============================
    // (idle on the register-load path: a role that nothing waits for must not wait on the ring barriers,
    //  it could fall two phases behind and alias the parity)
    int it = 0;
    for (int tile = blockIdx.x; tma && tile < ntiles; tile += gridDim.x)
      for (int kc = 0; kc < kchunks; ++kc, ++it) {
        const int s = it % nstages;
        const uint32_t ph = (it / nstages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        const int nplanes = K >> 3;
        if (lane == 0) mbar_expect_tx(&rawfull[s], (uint32_t)(nplanes * kPlaneBytes * tma));
        __syncwarp();
        tma_issue_tile(a, tma, &tmA, &tmA2, nplanes, tile * kTileM, smem_u32(smem + L.a_stage0 + s * L.a_stage_bytes),
                       smem_u32(smem + L.raw2 + s * L.raw2_bytes), &rawfull[s], lane, kc * K);
      }
  } else if (warp == kMmaWarp) {
    // ====================== MMA issue + epilogue-operand prefetch ======================
    const uint32_t idesc = make_idesc(npad);
    // A: planes 4 KB apart, 8-row groups 256 B apart, lo core matrices 128 B behind the hi ones; W: plain planes
    const uint32_t lbo_a = kPlaneBytes, sbo_a = 256, lbo_b = (uint32_t)npad * 16, sbo_b = 128;
    const uint64_t bdesc_hi = make_desc(smem_u32(w_hi), lbo_b, sbo_b);
    const uint64_t bdesc_lo = make_desc(smem_u32(w_lo), lbo_b, sbo_b);
    int it = 0, its = 0;   // tiles, operand stages (tiles x chunks) processed by this CTA
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t pa = (it >> 1) & 1;
      mbar_wait(&tempty[acc], pa ^ 1);   // the epilogue of tile it-2 has released accumulator and operand tile
      if (eop_kind && lane == 0) {
        const int m0 = tile * kTileM;
        const int rows = M - m0 < kTileM ? M - m0 : kTileM;
        float* dst = reinterpret_cast<float*>(smem + L.eop + acc * L.eop_bytes);
        mbar_expect_tx(&efull[acc], (uint32_t)rows * N * 4);
        if (eop_kind == 2) {
          int m = m0;
          while (m < m0 + rows) {
            const int b = m / ep.T, t = m - b * ep.T;
            int run = ep.T - t;
            if (run > m0 + rows - m) run = m0 + rows - m;
            bulk_g2s(dst + (size_t)(m - m0) * N, ep.ga + ((size_t)(b / ep.G) * ep.T + t) * ep.ldga, (uint32_t)run * N * 4,
                     &efull[acc]);
            m += run;
          }
        } else {
          const float* src = eop_kind == 1 ? ep.hpre + (size_t)m0 * ep.ldh : ep.C + (size_t)m0 * ep.ldc;
          bulk_g2s(dst, src, (uint32_t)rows * N * 4, &efull[acc]);
        }
      }
      for (int kc = 0; kc < kchunks; ++kc, ++its) {
        const int s = its % nstages;
        const uint32_t ph = (its / nstages) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_base = smem_u32(smem + L.a_stage0 + s * L.a_stage_bytes);
          const uint64_t adesc_hi = make_desc(a_base, lbo_a, sbo_a);
          const uint64_t adesc_lo = make_desc(a_base + 128, lbo_a, sbo_a);
          const uint32_t d = tmem_base + (uint32_t)(acc * npad);
          for (int kb = 0; kb < nkb; ++kb) {
            // advance both operands by one 16-wide K block = two planes (W: past the blocks of the earlier chunks)
            const uint64_t ka = (uint64_t)((2 * lbo_a * kb) >> 4), kbo = (uint64_t)((2 * lbo_b * (kc * nkb + kb)) >> 4);
            umma_bf16(d, adesc_hi + ka, bdesc_hi + kbo, idesc, (kc > 0 || kb > 0) ? 1u : 0u);
            umma_bf16(d, adesc_lo + ka, bdesc_hi + kbo, idesc, 1u);
            umma_bf16(d, adesc_hi + ka, bdesc_lo + kbo, idesc, 1u);
          }
          umma_commit(&empty[s]);
          if (kc == kchunks - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== epilogue ===============================
    const int et = tid - 256;              // 0..255
    const int e8 = warp - 8;
    const int q = e8 & 3;                  // TMEM lane quadrant (= warp index % 4)
    const int half = e8 >> 2;              // this warp owns 16-column chunks half, half+2, ...
    const int row = q * 32 + lane;         // accumulator row owned by this thread
    const int flags = ep.flags;
    const bool need_h = flags & (E_RELUMASK | E_STAT_XHAT);
    const bool has_bias = ep.bias != nullptr;
    const bool vec_ok = (N & 3) == 0 && (ep.ldc & 3) == 0 && al16(ep.C) &&
                        (!need_h || eop_kind == 1 || ((ep.ldh & 3) == 0 && al16(ep.hpre))) &&
                        (!(flags & E_ROWBIAS) || ((ep.ldrb & 3) == 0 && al16(ep.rb))) &&
                        (!(flags & E_GROUPADD) || eop_kind == 2 || ((ep.ldga & 3) == 0 && al16(ep.ga)));
    const bool st256 = (ep.ldc & 7) == 0 && al32(ep.C);
    const bool ts = tstore != 0 && vec_ok && (N & 7) == 0;
    // The image uses the 32-byte TMA swizzle (16-byte halves of a row exchanged when bit 7 of the address is set, i.e.
    // for rows 4..7 of every 8): eight consecutive lanes then cover all 32 banks and a 16-byte store instruction costs
    // its minimum of four wavefronts.
    const uint32_t cst_row = smem_u32(smem + L.cst) + (uint32_t)row * 32;
    const uint32_t sw0 = ((row >> 2) & 1) * 16, sw1 = 16 - sw0;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    // Per-thread (= per-row) column statistics accumulate in two spare TMEM regions
    // [2 npad, 3 npad) and [3 npad, 4 npad): no shuffles in the tile loop, one reduction per kernel.
    if (STATS) {
      float z[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) z[i] = 0.f;
      for (int c0 = half * 16; c0 < npad; c0 += 32) {
        tmem_st16(tlane + (uint32_t)(2 * npad + c0), z);
        tmem_st16(tlane + (uint32_t)(3 * npad + c0), z);
      }
      tmem_wait_st();
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t pa = (it >> 1) & 1;
      const int m0 = tile * kTileM;
      const int m = m0 + row;
      const bool partial = m0 + kTileM > M;
      const bool ok = m < M;
      const int mc = ok ? m : M - 1;       // clamped: loads of rows past M stay in bounds, results are discarded
      const float* rbrow = (flags & E_ROWBIAS) ? ep.rb + (size_t)(mc / ep.rbT) * ep.ldrb : nullptr;
      const float* garow = nullptr;
      if ((flags & E_GROUPADD) && eop_kind != 2) {
        const int b = mc / ep.T, t = mc - b * ep.T;
        garow = ep.ga + ((size_t)(b / ep.G) * ep.T + t) * ep.ldga;
      }
      const float* hrow = (need_h && eop_kind != 1) ? ep.hpre + (size_t)mc * ep.ldh : nullptr;
      float* crow = ep.C + (size_t)mc * ep.ldc;
      const float* erow = reinterpret_cast<const float*>(smem + L.eop + acc * L.eop_bytes) + (size_t)row * N;
      mbar_wait(&tfull[acc], pa);
      tc_fence_after();
      if (eop_kind) mbar_wait(&efull[acc], pa);

      // one 16-column chunk with every operand 16-byte addressable; FULL: the chunk lies inside [0, N)
      auto chunk_vec = [&](auto full_c, int c0, float* v, float* s1, float* s2) {
        constexpr bool FULL = decltype(full_c)::value;
        const int nq = FULL ? 4 : ((N - c0 + 3) >> 2);   // valid column quads
        float h[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) h[i] = 0.f;
        if (has_bias) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (FULL || j < nq) {
              const float4 t4 = *reinterpret_cast<const float4*>(sbias + c0 + 4 * j);
              v[4 * j] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w;
            }
        }
        if (flags & E_ROWBIAS) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (FULL || j < nq) {
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(rbrow + c0 + 4 * j));
              v[4 * j] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w;
            }
        }
        if (flags & E_GROUPADD) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (FULL || j < nq) {
              const float4 t4 = eop_kind == 2 ? *reinterpret_cast<const float4*>(erow + c0 + 4 * j)
                                              : __ldg(reinterpret_cast<const float4*>(garow + c0 + 4 * j));
              v[4 * j] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w;
            }
        }
        if (need_h) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (FULL || j < nq) {
              const float4 t4 = eop_kind == 1 ? *reinterpret_cast<const float4*>(erow + c0 + 4 * j)
                                              : __ldg(reinterpret_cast<const float4*>(hrow + c0 + 4 * j));
              h[4 * j] = t4.x; h[4 * j + 1] = t4.y; h[4 * j + 2] = t4.z; h[4 * j + 3] = t4.w;
            }
        }
        if (flags & E_RELUMASK) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 sc = *reinterpret_cast<const float4*>(sscale + c0 + 4 * j);
            const float4 sh = *reinterpret_cast<const float4*>(sshift + c0 + 4 * j);
            if (!(fmaf(h[4 * j], sc.x, sh.x) > 0.f)) v[4 * j] = 0.f;
            if (!(fmaf(h[4 * j + 1], sc.y, sh.y) > 0.f)) v[4 * j + 1] = 0.f;
            if (!(fmaf(h[4 * j + 2], sc.z, sh.z) > 0.f)) v[4 * j + 2] = 0.f;
            if (!(fmaf(h[4 * j + 3], sc.w, sh.w) > 0.f)) v[4 * j + 3] = 0.f;
          }
        }
        if (flags & E_ACCUM) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (FULL || j < nq) {
              const float4 t4 = eop_kind == 3 ? *reinterpret_cast<const float4*>(erow + c0 + 4 * j)
                                              : *reinterpret_cast<const float4*>(crow + c0 + 4 * j);
              v[4 * j] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w;
            }
        }
        if (!FULL) {
#pragma unroll
          for (int j = 1; j < 4; ++j)
            if (j >= nq) { v[4 * j] = 0.f; v[4 * j + 1] = 0.f; v[4 * j + 2] = 0.f; v[4 * j + 3] = 0.f; }
        }
        if (partial && !ok) {
#pragma unroll
          for (int i = 0; i < 16; ++i) { v[i] = 0.f; h[i] = 0.f; }
        }
        if (ts) {
          // plane c0 / 8 (and the next one when the chunk holds 16 valid columns), this thread's 32-byte row
          const uint32_t d0 = cst_row + (uint32_t)(acc * L.cst_bytes) + (uint32_t)(c0 >> 3) * kPlaneBytes;
          sts4(d0 + sw0, v[0], v[1], v[2], v[3]);
          sts4(d0 + sw1, v[4], v[5], v[6], v[7]);
          if (FULL || nq > 2) {
            sts4(d0 + kPlaneBytes + sw0, v[8], v[9], v[10], v[11]);
            sts4(d0 + kPlaneBytes + sw1, v[12], v[13], v[14], v[15]);
          }
        } else if (ok) {
          if (FULL && st256) {
            st8(crow + c0, v, true);
            st8(crow + c0 + 8, v + 8, true);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (FULL || j < nq)
                *reinterpret_cast<float4*>(crow + c0 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
        if (STATS) {
          if (flags & E_STAT_XHAT) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 rs = *reinterpret_cast<const float4*>(srstd + c0 + 4 * j);
              const float4 mr = *reinterpret_cast<const float4*>(smr + c0 + 4 * j);
              s2[4 * j] = fmaf(v[4 * j], fmaf(h[4 * j], rs.x, mr.x), s2[4 * j]);
              s2[4 * j + 1] = fmaf(v[4 * j + 1], fmaf(h[4 * j + 1], rs.y, mr.y), s2[4 * j + 1]);
              s2[4 * j + 2] = fmaf(v[4 * j + 2], fmaf(h[4 * j + 2], rs.z, mr.z), s2[4 * j + 2]);
              s2[4 * j + 3] = fmaf(v[4 * j + 3], fmaf(h[4 * j + 3], rs.w, mr.w), s2[4 * j + 3]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) s2[i] = fmaf(v[i], v[i], s2[i]);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) s1[i] += v[i];
        }
      };

      if (ts) {
        // the previous tile's stores must have finished reading the image before it is overwritten
        if (et == 0) bulk_wait_read1();   // (the store of the tile before last used this image)
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      for (int c0 = half * 16; c0 < npad; c0 += 32) {
        float v[16], s1[16], s2[16];
        const uint32_t tcol = tlane + (uint32_t)(acc * npad + c0);
        if (STATS) {
          tmem_wait_st();
          tmem_ld16x3(tcol, tlane + (uint32_t)(2 * npad + c0), tlane + (uint32_t)(3 * npad + c0), v, s1, s2);
        } else {
          tmem_ld16(tcol, v);
        }
        if (vec_ok) {
          if (c0 + 16 <= N) chunk_vec(std::true_type{}, c0, v, s1, s2);
          else chunk_vec(std::false_type{}, c0, v, s1, s2);
        } else {
          // generic element-wise epilogue (odd N / unaligned operands)
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int n = c0 + i;
            float x = 0.f, t = 0.f;
            if (ok && n < N) {
              x = v[i] + sbias[n];
              if (flags & E_ROWBIAS) x += rbrow[n];
              if (flags & E_GROUPADD) x += eop_kind == 2 ? erow[n] : garow[n];
              float hp = 0.f;
              if (need_h) hp = eop_kind == 1 ? erow[n] : hrow[n];
              if (flags & E_RELUMASK) {
                if (!(fmaf(hp, sscale[n], sshift[n]) > 0.f)) x = 0.f;
              }
              if (flags & E_ACCUM) x += eop_kind == 3 ? erow[n] : crow[n];
              crow[n] = x;
              t = (flags & E_STAT_XHAT) ? x * fmaf(hp, srstd[n], smr[n]) : x * x;
            }
            if (STATS) { s1[i] += x; s2[i] += t; }
          }
        }
        if (STATS) {
          tmem_st16(tlane + (uint32_t)(2 * npad + c0), s1);
          tmem_st16(tlane + (uint32_t)(3 * npad + c0), s2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (ts) {
        fence_proxy_async();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) {
          const uint32_t img = smem_u32(smem + L.cst) + (uint32_t)(acc * L.cst_bytes);
          for (int p = 0; p < (N >> 3); ++p) tma_store_plane(&tmC, img + (uint32_t)p * kPlaneBytes, p * 8, m0);
          bulk_commit();
        }
      }
    }
    if (ts && et == 0) bulk_wait_read0();
    if (STATS) {
      // fold the per-row accumulators: 16-shuffle transpose reduction per chunk, quadrants summed in double
      double* myd = dstat + (size_t)q * 2 * npad;
      const int cmap = colmap16(lane);
      tmem_wait_st();
      for (int c0 = half * 16; c0 < npad; c0 += 32) {
        float v[16], s1[16], s2[16];
        tmem_ld16x3(tlane + (uint32_t)c0, tlane + (uint32_t)(2 * npad + c0), tlane + (uint32_t)(3 * npad + c0), v, s1, s2);
        const float r1 = col_reduce16(s1, lane), r2 = col_reduce16(s2, lane);
        if (!(lane & 1) && c0 + cmap < N) {
          myd[c0 + cmap] = (double)r1;
          myd[npad + c0 + cmap] = (double)r2;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int n = et; n < N; n += kEpilogue) {
        double a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) { a1 += dstat[(size_t)qq * 2 * npad + n]; a2 += dstat[(size_t)qq * 2 * npad + npad + n]; }
        atomicAdd(ep.stat + n, a1);
        atomicAdd(ep.stat + N + n, a2);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
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
  int a_bytes, b_bytes, a2_bytes, b2_bytes, stage_bytes, bars, total;
};
// Stage = [A planes (ceil(acols / 8)) | B planes (npad / 8) | raw second stream of A | raw second stream of B].
// The MMA always reads 16 A planes (its 128 lanes); only the first ceil(acols / 8) of them carry operand columns
// (acols = K, + 1 for the constant-one column).  The lanes past them produce accumulator rows nobody reads, and
// accumulator rows are independent, so those planes are not stored: the descriptor simply runs on into whatever
// follows the A planes in shared memory (B planes, raw buffers, the next stage).  That is what lets two stages of
// the wide problems (dWx, dWt, dKm, dWs0t) fit; `total` keeps the 64 KB window of the last stage inside the allocation.
__host__ __device__ inline DwSmem dw_smem_layout(int K, int acols, int N, int npad, int nstages, int tma_a, int tma_b) {
  DwSmem s;
  s.a_bytes = ((acols + 7) >> 3) * kPlaneBytes;
  s.b_bytes = (npad / 8) * kPlaneBytes;
  s.a2_bytes = tma_a == 2 ? (K >> 3) * kPlaneBytes : 0;
  s.b2_bytes = tma_b == 2 ? (N >> 3) * kPlaneBytes : 0;
  s.stage_bytes = s.a_bytes + s.b_bytes + s.a2_bytes + s.b2_bytes;
  s.bars = nstages * s.stage_bytes;
  s.total = s.bars + 128;
  const int window = (nstages - 1) * s.stage_bytes + 16 * kPlaneBytes;
  if (s.total < window) s.total = window;
  return s;
}

// Split of the producer octets between the two operands: every octet owns one plane and every
// floor(octets / planes)-th row octet of it, so the slab time of a group is ceil(16 / floor(octets / planes))
// pieces per thread times the cost of a piece; pick the split with the smallest maximum.
inline int dw_split(int noct, int planes_a, int planes_b, int cost_a, int cost_b) {
  // whole warps only (4 octets): a warp that straddles the two groups would run both kinds of work serially
  int best = -1, best_cost = 1 << 30;
  for (int oa = (planes_a + 3) / 4 * 4; oa <= noct - planes_b; oa += 4) {
    const int pa = oa / planes_a, pb = (noct - oa) / planes_b;
    const int ca = ((16 + pa - 1) / pa) * cost_a, cb = ((16 + pb - 1) / pb) * cost_b;
    const int c = ca > cb ? ca : cb;
    if (c < best_cost) { best_cost = c; best = oa; }
  }
  if (best < 0) best = planes_a;   // no warp-aligned split leaves room for B: fall back to the tightest one
  return best;
}
inline int piece_cost(int mode) {   // relative instruction cost of producing one piece
  switch (mode) {
    case A_PLAIN: return 10;
    case A_BNRELU: return 12;
    case A_AFFINE2: return 18;
    case A_CATMUL: return 18;
    case A_MULROW: return 24;
    default: return 13;
  }
}

constexpr int kDwProducers = 384;              // warps 0-11 (warps 0-3 also run the final epilogue)
constexpr int kDwLoadWarp = 12;                // TMA operand loads / L2 prefetch
constexpr int kDwMmaWarp = 13;
constexpr int kDwThreads = 448;

// Body of the weight-gradient kernel for one problem; `cta` of `nctas` CTAs share the problem's row slabs.
// (The tensor maps are passed by address: they must stay in the kernel's parameter space.)
// transposed: the caller has exchanged the operands (a = the wide one on the MMA's lanes, K = its width; b = the narrow one,
// N = its width, + a constant-one column at index N when colsum is wanted): the accumulator then holds dW^T, lane = column
// of dW.  Used when the original N >= 2 K: the MMA's N shrinks from up to 240 to <= 48 and no lane is wasted.
CLSR_DEVINL void dw_body(uint8_t* smem, float* sva, float* svb, int M, int K, int acols, int N, int npad, int nstages,
                         uint32_t tmem_cols, int tma_a, int tma_b, int octa, int transposed, const AOp& a, const AOp& b,
                         float* __restrict__ dW, int lddw, float* __restrict__ colsum, const CUtensorMap* tmA,
                         const CUtensorMap* tmA2, const CUtensorMap* tmB, const CUtensorMap* tmB2, int cta, int nctas) {
  const DwSmem L = dw_smem_layout(K, acols, N, npad, nstages, tma_a, tma_b);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* full = bars;         // [2]
  uint64_t* empty = bars + 4;    // [4]   (full: bars + 0, [4]; up to four operand stages)
  uint64_t* done = bars + 8;     // [1]
  uint64_t* rawfull = bars + 9;  // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = (M + kTileM - 1) / kTileM;
  const Fast fa = fast_eligible(a), fb = fast_eligible(b);
  const bool any_tma = tma_a || tma_b;

  // zero the operand stages once: lanes / columns past the operand widths stay zero for the whole kernel
  for (int i = tid; i < (nstages * L.stage_bytes) / 16; i += kDwThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  stage_vectors(a, sva, 128, K, tid, kDwThreads);
  stage_vectors(b, svb, 256, N, tid, kDwThreads);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(&full[i], kDwProducers); mbar_init(&empty[i], 1); mbar_init(&rawfull[i], 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kDwMmaWarp) tmem_alloc(tmem_slot, tmem_cols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kDwLoadWarp) {
    // octa (chosen on the host, dw_split): producer octets that work on A; the rest work on B
    int it = 0;
    for (int tile = cta; tile < ntiles; tile += nctas, ++it) {
      const int s = it % nstages;
      const uint32_t ph = (it / nstages) & 1;
      // a group whose operand comes by TMA waits for the raw planes (which implies empty[s]: the loader
      // waited for it); a group that loads its operand itself only needs the stage to be free
      if (tid < octa * 8 ? tma_a != 0 : tma_b != 0) mbar_wait(&rawfull[s], ph);
      else {
        mbar_wait(&empty[s], ph ^ 1);
        // register-load path: the first warp of the group prefetches the rows of a later slab into L2
        if (CLSR_TC_PREFETCH) {
          const int mp = (tile + CLSR_TC_PREFETCH * nctas) * kTileM;
          if (warp == 0 && !tma_a) prefetch_operand(a, mp, M, K, lane);
          if (warp == (octa * 8 + 31) / 32 && tid >= octa * 8 && !tma_b) prefetch_operand(b, mp, M, N, lane);
        }
      }
      const uint32_t a_base = smem_u32(smem + s * L.stage_bytes);
      const uint32_t b_base = a_base + L.a_bytes;
      const uint32_t a_raw2 = b_base + L.b_bytes;
      const uint32_t b_raw2 = a_raw2 + L.a2_bytes;
      const int m0 = tile * kTileM;
      // the producer octets are split between the two operands in proportion to their plane counts, so
      // the loads of A and B are in flight together (one exposed load latency per slab, not two)
      if (tid < octa * 8)
        produce_tile(a, fa, tma_a, sva, 128, m0, M, K, (colsum && !transposed) ? K : -1, tid, octa * 8, a_base, a_raw2);
      else
        produce_tile(b, fb, tma_b, svb, 256, m0, M, N, (colsum && transposed) ? N : -1, tid - octa * 8,
                     kDwProducers - octa * 8, b_base, b_raw2);
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }
  } else if (warp == kDwLoadWarp) {
    // idle unless an operand comes by TMA (see the loader role of tc_gemm_kernel)
    int it = 0;
    for (int tile = cta; any_tma && tile < ntiles; tile += nctas, ++it) {
      const int s = it % nstages;
      const uint32_t ph = (it / nstages) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      const uint32_t a_base = smem_u32(smem + s * L.stage_bytes);
      const uint32_t b_base = a_base + L.a_bytes;
      const uint32_t a_raw2 = b_base + L.b_bytes;
      const uint32_t b_raw2 = a_raw2 + L.a2_bytes;
      if (lane == 0)
        mbar_expect_tx(&rawfull[s], (uint32_t)(((K >> 3) * tma_a + (N >> 3) * tma_b) * kPlaneBytes));
      __syncwarp();
      if (tma_a) tma_issue_tile(a, tma_a, tmA, tmA2, K >> 3, tile * kTileM, a_base, a_raw2, &rawfull[s], lane);
      if (tma_b) tma_issue_tile(b, tma_b, tmB, tmB2, N >> 3, tile * kTileM, b_base, b_raw2, &rawfull[s], lane);
    }
  } else {
    // MMA issue: D[128 x npad] += A^T-slab . B-slab, both operands MN-major
    const uint32_t idesc = make_idesc(npad) | (1u << 15) | (1u << 16);
    const uint32_t lbo = 256, sbo = kPlaneBytes;  // MN-major: LBO = 8-row group along the reduction, SBO = plane
    int it = 0;
    for (int tile = cta; tile < ntiles; tile += nctas, ++it) {
      const int s = it % nstages;
      const uint32_t ph = (it / nstages) & 1;
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_base = smem_u32(smem + s * L.stage_bytes);
        const uint32_t b_base = a_base + L.a_bytes;
        const uint64_t a_hi = make_desc(a_base, lbo, sbo), a_lo = make_desc(a_base + 128, lbo, sbo);
        const uint64_t b_hi = make_desc(b_base, lbo, sbo), b_lo = make_desc(b_base + 128, lbo, sbo);
        for (int rb = 0; rb < kTileM / 16; ++rb) {
          const uint64_t adv = (uint64_t)((rb * 2 * 256) >> 4);  // 16 rows = two 256-byte row-octet blocks
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
  if (warp < 4 && ntiles > (int)cta) {
    // final epilogue: accumulator lane = dW row
    mbar_wait(done, 0);
    tc_fence_after();
    const int krow = warp * 32 + lane;
    for (int c0 = 0; c0 < npad; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      if (transposed) {
        // lane = column krow of dW, accumulator column c = row c of dW; column N = the ones column: column sums
        if (krow < K) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + i;
            if (c < N) atomicAdd(&dW[(size_t)c * lddw + krow], v[i]);
            else if (colsum && c == N) atomicAdd(&colsum[krow], v[i]);
          }
        }
      } else if (krow < K) {
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
  if (warp == kDwMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

__global__ void __launch_bounds__(kDwThreads, 1)
tc_dw_kernel(int M, int K, int acols, int N, int npad, int nstages, uint32_t tmem_cols, int tma_a, int tma_b, int octa, int transposed,
             AOp a, AOp b,
             float* __restrict__ dW, int lddw, float* __restrict__ colsum, const __grid_constant__ CUtensorMap tmA,
             const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmB2) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(16) float sva[3 * 128];
  __shared__ __align__(16) float svb[3 * 256];
  dw_body(smem, sva, svb, M, K, acols, N, npad, nstages, tmem_cols, tma_a, tma_b, octa, transposed, a, b, dW, lddw, colsum, &tmA, &tmA2, &tmB,
          &tmB2, (int)blockIdx.x, (int)gridDim.x);
}

// Several independent weight-gradient problems in ONE launch: the launches over M = S*T rows process only ~11
// slabs per CTA each, so pipeline fill, TMEM allocation and the final atomics dominate them; here the grid is
// partitioned between the problems in proportion to their work and every CTA runs dw_body on its share.
constexpr int kDwGroupMax = 12;
struct DwProblem {
  int M, K, acols, N, npad, nstages, tma_a, tma_b, octa, transposed, lddw, cta0, ncta;
  uint32_t tmem_cols;
  AOp a, b;
  float* dW;
  float* colsum;
  CUtensorMap tmA, tmA2, tmB, tmB2;
};
struct DwGroup {
  int n;
  DwProblem p[kDwGroupMax];
};

__global__ void __launch_bounds__(kDwThreads, 1)
tc_dw_group_kernel(const __grid_constant__ DwGroup g) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(16) float sva[3 * 128];
  __shared__ __align__(16) float svb[3 * 256];
  int i = 0;
  while (i + 1 < g.n && (int)blockIdx.x >= g.p[i].cta0 + g.p[i].ncta) ++i;
  const DwProblem& P = g.p[i];
  dw_body(smem, sva, svb, P.M, P.K, P.acols, P.N, P.npad, P.nstages, P.tmem_cols, P.tma_a, P.tma_b, P.octa, P.transposed, P.a, P.b,
          P.dW, P.lddw,
          P.colsum, &P.tmA, &P.tmA2, &P.tmB, &P.tmB2, (int)blockIdx.x - P.cta0, P.ncta);
}

}  // namespace tc
}  // namespace clsr
