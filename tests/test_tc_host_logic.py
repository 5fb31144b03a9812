"""Host-side logic of the tcgen05 kernels (no GPU): producer split of the weight-gradient kernel and the
shared-memory layouts, compiled with nvcc as plain host code from clsr_b200/csrc/tc_gemm.cuh."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

SRC = r'''
#include <cstdio>
#include "tc_gemm.cuh"
#include "head_mlp.cuh"
using namespace clsr;
int main() {
  // (planes_a, planes_b, mode_a, mode_b) of the model's weight-gradient launches
  const int cases[][4] = {{5, 10, A_MULROW, A_AFFINE2}, {10, 5, A_BNRELU, A_AFFINE2}, {5, 10, A_PLAIN, A_PLAIN},
                          {6, 30, A_PLAIN, A_PLAIN},   {5, 20, A_PLAIN, A_PLAIN},    {15, 10, A_CATMUL, A_AFFINE2},
                          {16, 1, A_PLAIN, A_PLAIN},   {1, 30, A_PLAIN, A_PLAIN}};
  for (auto& c : cases) {
    const int oa = tc::dw_split(tc::kDwProducers / 8, c[0], c[1], tc::piece_cost(c[2]), tc::piece_cost(c[3]));
    printf("split %d %d %d\n", c[0], c[1], oa);
  }
  // shared-memory footprints: (K, N, stages, eop, stats, tma)
  const int g[][6] = {{40, 80, 2, 1, 1, 2}, {80, 40, 2, 0, 1, 1}, {80, 40, 2, 0, 0, 2}, {160, 80, 2, 0, 0, 1}, {40, 240, 2, 0, 0, 1}};
  for (auto& c : g) {
    const int kpad = (c[0] + 15) / 16 * 16, npad = (c[1] + 15) / 16 * 16;
    tc::Smem L = tc::smem_layout(kpad, npad, c[1], c[2], c[3], c[4], c[5]);
    printf("gemm %d %d %d %d\n", c[0], c[1], L.total, L.a_stage_bytes);
  }
  tc::DwSmem D = tc::dw_smem_layout(80, 80, 40, 48, 2, 1, 2);
  printf("dw %d %d %d %d\n", D.a_bytes, D.b_bytes, D.b2_bytes, D.total);
  // the widest weight gradients double-buffer once the unused MMA lanes are not stored: dWx slab (K=40 + ones
  // column, 160 plain columns), dWt (80 x 120), dKm (40 x 160), dWs0t (40 x 80, two-stream B)
  const int w[][5] = {{40, 41, 160, 1, 1}, {80, 80, 120, 1, 1}, {40, 40, 160, 1, 1}, {40, 40, 80, 0, 2}};
  for (auto& c : w) {
    tc::DwSmem E = tc::dw_smem_layout(c[0], c[1], c[2], (c[2] + 15) / 16 * 16, 2, c[3], c[4]);
    printf("dw2 %d %d %d %d\n", c[0], c[2], E.total, (2 - 1) * E.stage_bytes + 16 * 4096);
  }
  // fused _fcn_net kernels (head_mlp.cuh): alpha gate (K = 161) and logit (K = 80) blocks at 139 rows per CTA (20480 / 148)
  const int m[][3] = {{161, 80, 40}, {80, 100, 64}, {121, 80, 40}};
  for (auto& c : m) {
    CoopFwdSmem F = coop_fwd_smem(c[0], c[1], c[2], 139);
    CoopBwdSmem B = coop_bwd_smem(c[0], c[1], c[2], 139);
    printf("coop %d %d %d %d %d %d\n", c[0], c[1], c[2], F.total * 4, B.total * 4, F.rc);
  }
  // three weight-gradient stages where a stage is <= 75 KB (dWgh: 40 x 80 transposed -> 80 lanes + 48 columns)
  tc::DwSmem T3 = tc::dw_smem_layout(80, 80, 40, 48, 3, 1, 1);
  printf("dw3 %d %d\n", T3.total, 2 * T3.stage_bytes + 16 * 4096);
  // chunked contraction (dX: 6 chunks of 80, N = 40) with two stages and the store images
  tc::Smem KX = tc::smem_layout(80, 48, 40, 2, 0, 0, 1, 1, 6);
  printf("kloop %d %d\n", KX.total, KX.a_stage0);
  return 0;
}
'''


@pytest.fixture(scope="module")
def host_out(tmp_path_factory):
    if not shutil.which(NVCC):
        pytest.skip("nvcc not available")
    d = tmp_path_factory.mktemp("tc_host")
    src = d / "t.cu"
    src.write_text(SRC)
    exe = d / "t"
    r = subprocess.run([NVCC, "-std=c++17", "-O1", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-I", os.path.join(ROOT, "clsr_b200", "csrc"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0
    return [ln.split() for ln in out.stdout.splitlines()]


def test_dw_split_is_warp_aligned_and_leaves_room(host_out):
    noct = 48
    for tag, pa, pb, oa in [ln for ln in host_out if ln[0] == "split"]:
        pa, pb, oa = int(pa), int(pb), int(oa)
        assert oa >= pa and noct - oa >= pb, (pa, pb, oa)          # every plane has an octet
        if (pa + 3) // 4 * 4 <= noct - pb:
            assert oa % 4 == 0, (pa, pb, oa)                        # whole warps per operand


def test_operand_stage_is_one_plane_per_eight_columns(host_out):
    for tag, K, N, total, stage in [ln for ln in host_out if ln[0] == "gemm"]:
        K, total, stage = int(K), int(total), int(stage)
        assert stage == (K + 15) // 16 * 16 // 8 * 4096
        assert total <= 227 * 1024, (K, N, total)
    dw = [ln for ln in host_out if ln[0] == "dw"][0]
    a, b, b2, total = map(int, dw[1:])
    assert a == 10 * 4096 and b == 6 * 4096 and b2 == 5 * 4096 and total <= 227 * 1024
    for ln in host_out:
        if ln[0] == "dw2":
            k, n, tot, window = map(int, ln[1:])
            # two stages fit, and the 16-plane window the MMA reads from the last stage stays inside the allocation
            assert tot <= 227 * 1024 - 4608 - 256 and window <= tot, ln


def test_fused_head_and_pipeline_layouts_fit(host_out):
    """Shared-memory budgets the round-2 kernels rely on: the cooperative _fcn_net kernels (one CTA per SM, 512 threads,
    ~8 KB of static shared memory), three weight-gradient stages, the chunked-K GEMM with all of W resident."""
    limit = 227 * 1024
    for ln in host_out:
        if ln[0] == "coop":
            k, n0, n1, fwd, bwd, rc = map(int, ln[1:])
            assert fwd + 8 * 1024 <= limit and bwd + 8 * 1024 <= limit, ln
            assert rc % 4 == 0 and rc >= 4
        if ln[0] == "dw3":
            tot, window = map(int, ln[1:])
            assert tot <= limit - 4608 - 256 and window <= tot, ln
        if ln[0] == "kloop":
            tot, a0 = map(int, ln[1:])
            assert tot <= limit - 1024 and a0 == 2 * 6 * 80 * 48 * 2, ln   # hi + lo copies of all six W chunks
