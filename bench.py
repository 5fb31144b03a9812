#!/usr/bin/env python
"""Benchmark of the CLSR training step (BASELINE.json: user-sequences/sec at seq_len=50,
emb_dim=40; HBM GB/s on the embedding gather).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [--steps K] [--warmup W]      # reference CPU path (restated)

One step = one CLSRModel.train call (forward, losses, backward, sparse-gradient scatter-add,
per-variable clip, Adam incl. the TF non-lazy table sweep, BN moving-stat update) over one
synthetic Taobao-shaped batch: 4096 user-sequences (20480 graph rows), T=50, D=32+8,
4.0M items / 9.4k categories / 1.0M users.  Prints ONE JSON line (rank 0).

Timing: W warm-up steps, then `--windows` (default 5) windows of EXACTLY K steps each, every window
bracketed by barrier + synchronize and timed with CUDA events on the engine's stream with the
per-kernel profiling events OFF; `value` comes from the median window (all windows are listed).  One
further window runs with per-kernel events on and only feeds the per-kernel tables.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1] / configs[2]
    "taobao": dict(T=50, n_items=4_000_000, n_cates=9_400, n_users=1_000_000, time_unit="s", seqs=4096),
    "kuaishou": dict(T=250, n_items=4_000_000, n_cates=9_400, n_users=1_000_000, time_unit="ms", seqs=4096),
    "small": dict(T=50, n_items=64_005, n_cates=2_182, n_users=36_653, time_unit="s", seqs=500),
    # BASELINE.json configs[3] / configs[4]: tables that only fit row-sharded (device-side initialisation)
    "synth50m": dict(T=50, n_items=50_000_000, n_cates=9_400, n_users=1_000_000, time_unit="s", seqs=4096,
                     dims=dict(Di=112, Dc=16, U=128, H=128, A0=80, A1=40, L0=100, L1=64), device_init=True),
    "synth200m": dict(T=200, n_items=200_000_000, n_cates=9_400, n_users=1_000_000, time_unit="s", seqs=4096,
                      dims=dict(Di=224, Dc=32, U=256, H=256, A0=80, A1=40, L0=100, L1=64), device_init=True),
}
DIMS = dict(Di=32, Dc=8, U=40, H=40, A0=80, A1=40, L0=100, L1=64)
G = 5  # 1 + train_num_ngs
TENSOR_PEAK_FALLBACK = 1381.8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--windows", type=int, default=5, help="repeated K-step timed windows; the median is reported")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="taobao", choices=sorted(WORKLOADS))
    ap.add_argument("--seqs", type=int, default=0, help="user-sequences per step per GPU (default: workload's)")
    ap.add_argument("--optimizer", default="adam", choices=["adam", "lazyadam"])
    ap.add_argument("--ref-seqs", type=int, default=500,
                    help="CPU arm: sequences per CPU step (500 = BASELINE config 1, the reference's own batch size)")
    ap.add_argument("--math", default="tc", choices=["tc", "simt"],
                    help="tc: large GEMMs on tcgen05 (split-bf16, fp32 accumulate); simt: fp32 CUDA cores everywhere")
    ap.add_argument("--dp", default="sharded", choices=["sharded", "replicated"],
                    help="N > 1: row-sharded tables + peer-memory reductions (default) or replicated tables + NCCL all-gathers")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the config-3 (T=250) side measurement and the standalone gather")
    ap.add_argument("--profile-out", default="")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return (d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", TENSOR_PEAK_FALLBACK),
                "measured (MEASURED_PEAKS.json)")
    return 6650.0, TENSOR_PEAK_FALLBACK, "fallback (B200_PROFILING.md)"


def make_tables(w, seed, d=DIMS):
    """Truncated-normal(0.01) tables (base_model.py:161-165) as CPU torch tensors."""
    import torch
    g = torch.Generator().manual_seed(seed)
    shapes = {"item_embedding": (w["n_items"], d["Di"]), "cate_embedding": (w["n_cates"], d["Dc"]),
              "user_long_embedding": (w["n_users"], d["U"]), "user_short_embedding": (w["n_users"], d["U"])}
    out = {}
    for k, s in shapes.items():
        t = torch.empty(s, dtype=torch.float32)
        torch.nn.init.trunc_normal_(t, std=0.01, a=-0.02, b=0.02, generator=g)
        out["sequential/embedding/" + k] = t
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            if t0 - 0.2 <= ts <= t1 + 0.2:
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the newest committed ncu capture
    (profiles/rNN_ncu_top_kernels.json; Taobao workload only), or None."""
    t = None
    for r in ("r02", "r01"):
        try:
            with open(os.path.join(ROOT, "profiles", "%s_ncu_top_kernels.json" % r)) as f:
                t = json.load(f)["traffic"]
            break
        except Exception:
            continue
    if t is None:
        return None
    alias = {"adam_sweep": ("adam_sweep (4 tables, one step)",), "scatter_hist": ("clsr::scatter_hist_kernel", "scatter_hist_kernel"),
             "gather_hist": ("clsr::gather_hist_tma_kernel", "gather_hist_tma_kernel", "gather_hist_kernel<4>")}
    for key in (name,) + alias.get(name, ()):
        if key in t:
            return t[key].get("dram_bytes", t[key].get("dram_bytes_per_launch_mean"))
    return None


def kernel_models(S, B, T, w, d=DIMS, shards=1):
    """Algorithmic HBM bytes and FLOPs of one launch of the named kernels (kernel table).
    bytes: every operand read / written once; flops: 2*M*N*K of the contraction (the split-bf16 MMAs issue 3x that)."""
    M, MB = S * T, B * T
    D, U, H, A0, A1 = d["Di"] + d["Dc"], d["U"], d["H"], d["A0"], d["A1"]
    Q, NX = U + D, 3 * U + 9 * H
    rows_state = (w["n_items"] * d["Di"] + w["n_cates"] * d["Dc"] + 2 * w["n_users"] * U) * 4 * 6 + \
                 (w["n_items"] + w["n_cates"] + 2 * w["n_users"]) * 4
    rows_state //= shards   # row-sharded tables: every rank sweeps its 1/world of the rows
    f4 = 4
    m = {
        "gather_hist": (M * (8 + 2 * D * f4), None),                     # T*8 + 2*T*D*e per sequence
        "scatter_hist": (M * (8 + 2 * D * f4), None),                    # read d_hist + ids, RMW-add one row slice
        "adam_sweep": (rows_state, None),
        # hoisted input projections: columns [0, NX-3H) = x.Wx; the last 3H columns = [x | time features].[Wx ; Wt] in one product
        "px": (M * (2 * D + NX - 3 * H) * f4, 2 * M * (NX - 3 * H) * D),
        "px_time": (M * (D + 2 * H + 3 * H) * f4, 2 * M * 3 * H * (D + 2 * H)),
        "al": (M * (D + U) * f4, 2 * M * D * U),
        "h0l": (M * (U + A0) * f4, 2 * M * 2 * U * A0),
        "h1l": (M * (A0 + A1) * f4, 2 * M * A0 * A1),
        "as": (M * (H + Q) * f4, 2 * M * H * Q),
        "invs": (M * (Q + A0) * f4, 2 * M * (Q + U) * A0),
        "h0s": (MB * A0 * f4 + M * (D + A0) * f4, 2 * MB * D * A0),      # write h0s, read a2 + invs once
        "h1s": (MB * (A0 + A1) * f4, 2 * MB * A0 * A1),
        "dy0s": (MB * (2 * A1 + 2 * A0) * f4, 2 * MB * A1 * A0),         # read dy1s,h1s,h0s; write dy0s
        "dP": (MB * (2 * A0 + D) * f4, 2 * MB * A0 * D),
        "dW1s": (MB * (A0 + 2 * A1) * f4, 2 * MB * A0 * A1),
        "dWs0t": (MB * 2 * A0 * f4 + M * D * f4, 2 * MB * D * A0),
        "short_fwd_fused": (M * (D + A0) * f4 + MB * f4, 2 * MB * (D * A0 + A0 * A1)),
        "pool_bwd_short": (MB * 2 * A1 * f4 + M * 2 * D * f4, None),
        "pool_fwd_short": (MB * A1 * f4 + M * D * f4, None),
        "h0_reduce_short": (MB * 2 * A0 * f4, None),
        "dX": (M * (NX + D) * f4, 2 * M * NX * D),                       # one launch, K walked in chunks inside the kernel
        "dW_bptt_group": (M * (NX + D + 2 * H + 2 * U + 3 * H) * f4,
                          2 * M * (D * NX + 2 * H * 3 * H + U * 3 * U + H * 3 * H + H * 4 * H)),
        # recurrences: read the hoisted projections once, write the stored gate / state tensors once
        "rnn_fwd(gru_sti|gru_causal2|time4lstm)": (M * (NX + 5 * U + 5 * H + 7 * H) * f4,
                                                     2 * M * (U * 3 * U + H * 3 * H + H * 4 * H)),
        "rnn_bwd(time4lstm|gru_sti|gru_causal2)": (M * (NX + 4 * U + 4 * H + 6 * H + H + NX) * f4,
                                                     2 * M * (U * 3 * U + H * 3 * H + H * 4 * H)),
    }
    return m


def step_roofline(T, seq_per_s_per_gpu, hbm_gbs, tensor_tflops, d=DIMS):
    """Whole-step bounds of SURVEY.md 8d / BASELINE.md section 3: minimum HBM bytes and
    (group-deduplicated) FLOPs per user-sequence, and the fraction of each bound the measured rate reaches."""
    D, U, H = d["Di"] + d["Dc"], d["U"], d["H"]
    A0, A1, L0, L1, Gr = d["A0"], d["A1"], d["L0"], d["L1"], G
    Q = U + D
    g_min = T * 8 + T * D * 4 + Gr * (8 + D * 4) + (4 + 2 * U * 4) + 3 * T * 4 + 2 * T * D * 4
    long_ = 2 * T * (D * U + 4 * U * A0 + A0 * A1 + A1)
    lstm = 2 * T * (2 * D * H + 4 * H * H + 4 * H * (D + H))
    gru = 2 * T * 3 * H * (D + H)
    short = 2 * T * (H * Q + 4 * Q * A0 + A0 * A1 + A1)
    alpha = 2 * ((2 * H + 2 * D + 1) * A0 + A0 * A1 + A1)
    logit = 2 * ((H + D) * L0 + L0 * L1 + L1)
    flops = 3 * (long_ + lstm + 2 * gru + Gr * (short + alpha + logit))
    hbm_bound = hbm_gbs * 1e9 / g_min
    cmp_bound = tensor_tflops * 1e12 / flops
    return {"hbm_min_bytes_per_seq": g_min, "hbm_bound_seq_per_s_per_gpu": hbm_bound,
            "frac_of_hbm_bound": seq_per_s_per_gpu / hbm_bound, "flops_per_seq": flops,
            "tensor_peak_tflops": tensor_tflops, "compute_bound_seq_per_s_per_gpu": cmp_bound,
            "frac_of_compute_bound": seq_per_s_per_gpu / cmp_bound,
            "note": "the step is latency / instruction bound at D=40 (SURVEY 8d); per-kernel rooflines are in `roofline` and `kernels`"}


def timed_windows(run_step, sync_all, stream, torch, steps, windows):
    """`windows` windows of exactly `steps` steps, each bracketed by barrier + synchronize, CUDA events on `stream`."""
    out = []
    for _ in range(windows):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for i in range(steps):
                run_step(i)
            e1.record(stream)
        sync_all()
        out.append(e0.elapsed_time(e1))
    return out


def measure(a, w, S, optimizer, rank, world, local, dist, steps, warmup, windows, extras):
    """Build an engine for workload `w` and time it; returns a dict of raw measurements."""
    import torch
    from clsr_b200 import params as P, synth
    from clsr_b200.engine import Engine, normalize_feed, TABLE_VARS
    B, T = S * G, w["T"]
    d = w.get("dims", DIMS)
    dev_init = bool(w.get("device_init"))
    eng = Engine(w["n_items"], w["n_cates"], w["n_users"], max_rows=B, seq_len=T, train_group=G,
                 item_dim=d["Di"], cate_dim=d["Dc"], user_dim=d["U"], hidden=d["H"], att_sizes=(d["A0"], d["A1"]),
                 layer_sizes=(d["L0"], d["L1"]), optimizer=optimizer, device=local, math_mode=1 if a.math == "tc" else 0,
                 alloc_tables=not (dev_init and world > 1), max_seqs=S)
    dense = P.init_params(1, 1, 1, Di=d["Di"], Dc=d["Dc"], U=d["U"], H=d["H"], att_sizes=(d["A0"], d["A1"]),
                          layer_sizes=(d["L0"], d["L1"]), seed=42, tables=False)
    eng.set_dense(dense)
    tabs = None
    if not dev_init:
        tabs = make_tables(w, 42, d)
        for t, name in TABLE_VARS.items():
            eng.tables[t].copy_(tabs[name])
    if world > 1:
        eng.comm_init(rank, world, dist, shard=(a.dp == "sharded"))
    if dev_init:   # tables too large to stage through the host: truncated-normal(0.01) drawn on the device, per shard
        gen = torch.Generator(device="cuda").manual_seed(42 + rank)
        for t in TABLE_VARS:
            torch.nn.init.trunc_normal_(eng.tables[t], std=0.01, a=-0.02, b=0.02, generator=gen)
    src = synth.SyntheticSource(w["n_items"], w["n_cates"], w["n_users"], T, seed=42 + 1000 * rank,
                                time_unit=w["time_unit"])
    NB = 4
    host = [normalize_feed(src.batch(S, G - 1)) for _ in range(NB)]
    pinned = []
    for f in host:  # pinned host copies for the end-to-end arm
        pf = {}
        for k, v in f.items():
            t = torch.from_numpy(v).pin_memory()
            pf[k] = t.numpy()
            pf["_keep_" + k] = t
        pinned.append(pf)
    dev = [eng.to_device(f) for f in host]
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    res = {"S": S, "B": B, "T": T}
    # ---- device-resident arm (value): profiling events off ----
    for i in range(warmup):
        eng.train_step(dev[i % NB], group=G, on_device=True, wait=False)
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    t0 = time.time()
    res["window_ms"] = timed_windows(lambda i: eng.train_step(dev[i % NB], group=G, on_device=True, wait=False),
                                     sync_all, stream, torch, steps, windows)
    t1 = time.time()
    res["clocks"] = sampler.stop(t0, t1) if sampler else None
    res["launches_per_step"] = eng.kernel_launches()
    # ---- one more window with per-kernel events on: feeds the per-kernel tables only ----
    eng.set_profiling(True)
    timed_windows(lambda i: eng.train_step(dev[i % NB], group=G, on_device=True, wait=False), sync_all, stream, torch,
                  steps, 1)
    res["prof"] = eng.profile()
    eng.set_profiling(False)
    # ---- end-to-end arm: host (pinned) feed -> C ABI -> losses read back, every step ----
    last = {}

    def e2e_step(i):
        last.update(eng.train_step(pinned[i % NB], group=G, normalized=True))
    for i in range(min(warmup, 3)):
        e2e_step(i)
    res["window_ms_e2e"] = timed_windows(e2e_step, sync_all, stream, torch, steps, windows)
    res["last_losses"] = dict(last)
    res["clip"] = eng.clip_report()
    # ---- standalone history gather (K1+K3 through clsr_gather_history), zipf and uniform ids ----
    if extras and world == 1:
        res["gather"] = {}
        D = d["Di"] + d["Dc"]
        for kind in ("zipf", "uniform"):
            try:
                if kind == "zipf":   # 32 batches' worth of the workload's windows
                    ih = torch.cat([d_["item_history"][::G] for d_ in dev] * 8).contiguous()
                    ch = torch.cat([d_["item_cate_history"][::G] for d_ in dev] * 8).contiguous()
                else:                # every position a uniformly random row: no hot rows for L2 to keep
                    g = torch.Generator(device="cuda").manual_seed(7)
                    n = NB * 8 * S * T
                    ih = torch.randint(1, w["n_items"], (n,), device="cuda", dtype=torch.int32, generator=g)
                    ch = torch.randint(1, w["n_cates"], (n,), device="cuda", dtype=torch.int32, generator=g)
                npos = ih.numel()
                gout = torch.empty(npos, D, device="cuda")
                call = lambda: eng._check(eng.lib.clsr_gather_history(eng.h, ih.data_ptr(), ch.data_ptr(), npos, gout.data_ptr()))
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    for _ in range(3):
                        call()
                    g0.record(stream)
                    for _ in range(10):
                        call()
                    g1.record(stream)
                sync_all()
                gms = g0.elapsed_time(g1) / 10
                gbytes = npos * (8 + 2 * D * 4)
                res["gather"][kind] = {"positions": npos, "ms": gms, "algorithmic_bytes": gbytes,
                                       "achieved": gbytes / 1e9 / (gms / 1e3), "unit": "GB/s", "bound": "hbm", "ids": kind}
                del gout, ih, ch
            except Exception as ex:  # the headline numbers do not depend on this extra measurement
                res["gather"][kind] = {"error": str(ex)}
    eng.close()
    del eng, dev
    torch.cuda.empty_cache()
    return res, dense, tabs


def run_b200(a):
    import torch
    import torch.distributed as dist
    from clsr_b200 import build

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    w = dict(WORKLOADS[a.workload])
    S = a.seqs or w["seqs"]
    res, dense, tabs = measure(a, w, S, a.optimizer, rank, world, local, dist, a.steps, a.warmup, a.windows,
                               extras=not a.no_extra)
    side = None
    if a.workload == "taobao" and world == 1 and not a.no_extra:
        # BASELINE config 3 (Kuaishou window, T=250) measured in the same run, shorter windows
        try:
            w3 = dict(WORKLOADS["kuaishou"])
            r3, _, _ = measure(a, w3, w3["seqs"], a.optimizer, rank, world, local, dist, max(a.steps // 2, 5),
                               min(a.warmup, 3), 3, extras=False)
            ms3, e3 = float(np.median(r3["window_ms"])), float(np.median(r3["window_ms_e2e"]))
            k3 = max(a.steps // 2, 5)
            side = {"kuaishou": {"workload": "seq_len=250 emb_dim=40 batch=4096 user-sequences, 1 GPU",
                                 "value": w3["seqs"] * k3 / (ms3 / 1e3), "ms_per_step": ms3 / k3,
                                 "e2e_value": w3["seqs"] * k3 / (e3 / 1e3), "e2e_ms_per_step": e3 / k3,
                                 "steps": k3, "windows": 3, "unit": "user-sequences/s"}}
        except Exception as ex:
            side = {"kuaishou": {"error": str(ex)}}
    ms_all, e2e_all = res["window_ms"], res["window_ms_e2e"]
    if world > 1:   # every window: max over ranks
        tt = torch.tensor(ms_all + e2e_all, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        v = [float(x) for x in tt]
        ms_all, e2e_all = v[:len(ms_all)], v[len(ms_all):]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms, ms_e2e = float(np.median(ms_all)), float(np.median(e2e_all))
    B, T = res["B"], res["T"]
    prof = res["prof"]
    h2d = 5 * S * T * 4 + S * 4 + 3 * B * 4
    hbm_peak, tensor_peak, peak_src = peaks()
    wd = w.get("dims", DIMS)
    per_kernel = {k: {"ms": v[0] / max(v[1], 1), "calls_per_step": v[1] / a.steps,
                      "share": v[0] / max(sum(x[0] for x in prof.values()), 1e-9)} for k, v in prof.items()}
    step_ms = lambda k: per_kernel[k]["ms"] * per_kernel[k]["calls_per_step"]
    top = max(per_kernel, key=step_ms)
    models = kernel_models(S, B, T, w, w.get("dims", DIMS), world if a.dp == "sharded" else 1)

    def roof(name):
        k = per_kernel.get(name)
        if not k:
            return None
        nbytes, flops = models.get(name, (None, None))
        t_ms = step_ms(name)   # the byte / flop models are per step: all launches recorded under this name together
        r = {"bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None, "traffic": None,
             "kernel": name, "peak_source": peak_src, "ms": t_ms}
        if nbytes is not None:
            gbs = nbytes / 1e9 / (t_ms / 1e3)
            r.update(achieved=gbs, frac=gbs / hbm_peak, algorithmic_bytes=nbytes,
                     traffic=ncu_traffic(name) if a.workload == "taobao" else None)
        if flops is not None:   # tensor-pipe view of the same launch (algorithmic flops, not the 3x issued by the split)
            tf = flops / 1e12 / (t_ms / 1e3)
            r["tensor"] = {"bound": "tensor", "achieved": tf, "peak": tensor_peak, "unit": "TFLOP/s", "frac": tf / tensor_peak,
                           "algorithmic_flops": flops, "issued_flops": 3 * flops}
        if name == "adam_sweep":
            r["note"] = "TF non-lazy Adam sweep, the table launches of one step taken together"
        return r
    main_roof = roof(top)
    if main_roof["achieved"] is None:   # dominant kernel without a byte model: report the top modelled kernel instead
        for k in sorted(per_kernel, key=step_ms, reverse=True):
            if k in models:
                main_roof = roof(k)
                main_roof["note"] = "top kernel by time is %s (no byte model); this is the next one" % top
                break
    gathers = res.get("gather") or {}
    for gk in gathers.values():
        if "achieved" in gk:
            gk.update(peak=hbm_peak, frac=gk["achieved"] / hbm_peak, peak_source=peak_src)
    out = {
        "metric": "user-sequences/sec (CLSR training step, seq_len=%d, emb_dim=%d)" % (T, wd["Di"] + wd["Dc"]),
        "value": S * world * a.steps / (ms / 1e3), "unit": "user-sequences/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 (split-bf16 tcgen05 MMA, fp32 accumulate) + f32" if a.math == "tc" else "f32", "data": "synthetic",
        "config": {"workload": "Taobao-shaped synthetic: seq_len=%d emb_dim=%d (%d+%d) batch=%d user-sequences "
                               "(%d rows) per GPU, %d items / %d cates / %d users, zipf ids, optimizer=%s"
                               % (T, wd["Di"] + wd["Dc"], wd["Di"], wd["Dc"], S, B, w["n_items"], w["n_cates"], w["n_users"], a.optimizer),
                   "l2": "not flushed: each step streams >2 GB of activations and (adam) 5 GB of table state, "
                         "far above the 126 MB L2; 4 distinct batches rotate",
                   "parallelism": "dp%d" % world + ("" if world == 1 else
                                                    (" row-sharded tables (NVLink peer-memory gathers / gradient pushes), peer-memory "
                                                     "all-reduces in the BatchNorm kernels" if a.dp == "sharded" else
                                                     " replicated tables, NCCL all-gathered sparse gradients"))},
        "timing": {"windows": len(ms_all), "window_ms": [round(x, 4) for x in ms_all], "reported": "median window",
                   "window_ms_e2e": [round(x, 4) for x in e2e_all], "per_kernel_events": "off in the timed windows"},
        "e2e": {"value": S * world * a.steps / (ms_e2e / 1e3), "unit": "user-sequences/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 36, "ms_per_step": ms_e2e / a.steps,
                "api": "clsr_train_step (C ABI) with pinned host feed arrays, losses read back every step"},
        "gpu_launches": res["launches_per_step"] * a.steps,
        "launches_per_step": res["launches_per_step"],
        "clocks": res["clocks"],
        "roofline": main_roof,
        "kernels": dict({n: roof(n) for n in ("gather_hist", "scatter_hist", "adam_sweep", "h0s", "h1s", "dy0s", "dP", "dW1s",
                                              "dWs0t", "px", "dX", "short_fwd_fused", "dW_bptt_group",
                                              "rnn_fwd(gru_sti|gru_causal2|time4lstm)", "rnn_bwd(time4lstm|gru_sti|gru_causal2)")
                         if n in per_kernel},
                        **({"gather_hist_large": gathers.get("zipf"), "gather_hist_large_uniform": gathers.get("uniform")}
                           if gathers else {})),
        "top_kernels": sorted(((k, round(step_ms(k), 4)) for k in per_kernel), key=lambda x: -x[1])[:12],
        "small_kernels": {"count_under_50us": sum(1 for k in per_kernel if per_kernel[k]["ms"] < 0.05 for _ in range(int(round(per_kernel[k]["calls_per_step"])))),
                          "sum_ms_under_50us": round(sum(step_ms(k) for k in per_kernel if per_kernel[k]["ms"] < 0.05), 4)},
        "clip": {"table_grad_norms": res["clip"][0], "grouped_steps_with_active_clip": res["clip"][1],
                 "max_grad_norm": 2.0},
        "last_losses": res["last_losses"],
    }
    if side:
        out["configs"] = side
    try:
        out["step_roofline"] = step_roofline(T, S * a.steps / (ms / 1e3), hbm_peak, tensor_peak, wd)
    except Exception as ex:  # informational only
        out["step_roofline"] = {"error": str(ex)}
    if a.profile_out:
        with open(a.profile_out, "w") as f:
            json.dump({"per_kernel": per_kernel, "ms_per_step": ms / a.steps}, f, indent=1)
    if world == 1 and not a.no_cpu_baseline and tabs is not None:
        try:
            out["cpu_baseline"] = cpu_arm(a, w, dense, tabs, steps=8, warmup=2)
        except Exception as ex:   # the GPU numbers above must still be printed
            out["cpu_baseline"] = {"error": "%s: %s" % (type(ex).__name__, ex)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_arm(a, w, dense, tabs, steps, warmup, optimizer=None):
    from clsr_b200 import synth
    from oracle import clsr_oracle as O
    from oracle.cpu_baseline import CpuTrainer, time_steps
    optimizer = optimizer or a.optimizer
    cfg = O.OracleConfig(max_seq_length=w["T"], optimizer=optimizer)
    prm = dict(dense)
    prm.update({k: v.clone() for k, v in tabs.items()})
    tr = CpuTrainer(prm, cfg, threads=os.cpu_count())
    src = synth.SyntheticSource(w["n_items"], w["n_cates"], w["n_users"], w["T"], seed=42, time_unit=w["time_unit"])
    batches = [src.batch(a.ref_seqs, G - 1) for _ in range(4)]
    sec = time_steps(tr, batches, steps, warmup)
    opt_sec = tr.opt_seconds / max(tr.step_no, 1)
    return {"value": a.ref_seqs / sec, "unit": "user-sequences/s", "cores": tr.threads, "kind": "port",
            "sample": "%d timed steps (+%d warm-up) of %d user-sequences (%d rows) -- BASELINE config 1's batch -- of the same "
                      "workload, fp32 PyTorch-CPU restatement of the TF1.15 graph, optimizer=%s; TF1.15 itself is not installable"
                      % (steps, warmup, a.ref_seqs, a.ref_seqs * G, optimizer),
            "sec_per_step": sec, "seqs_per_step": a.ref_seqs,
            "optimizer_sec_per_step": opt_sec,
            "value_without_optimizer": a.ref_seqs / max(sec - opt_sec, 1e-9),
            "same_config": False,
            "note": "the GPU arm runs %d sequences per step; both arms pay the full-table Adam sweep once per step, so the CPU "
                    "arm amortises it over fewer sequences -- `value_without_optimizer` removes that term" % (a.seqs or w["seqs"])}


def run_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from clsr_b200 import params as P
    w = dict(WORKLOADS[a.workload])
    S = a.seqs or w["seqs"]
    dense = P.init_params(1, 1, 1, seed=42, tables=False)
    tabs = make_tables(w, 42)
    wu = max(min(a.warmup, 5), 3)
    cb = cpu_arm(a, w, dense, tabs, steps=a.steps, warmup=wu)
    other = "lazyadam" if a.optimizer == "adam" else "adam"
    try:
        alt = cpu_arm(a, w, dense, tabs, steps=max(a.steps // 4, 3), warmup=2, optimizer=other)
        alt = {"optimizer": other, "value": alt["value"], "sec_per_step": alt["sec_per_step"]}
    except Exception as ex:
        alt = {"optimizer": other, "error": str(ex)}
    world = int(os.environ.get("WORLD_SIZE", 1))
    out = {
        "impl": "reference",
        "metric": "user-sequences/sec (CLSR training step, seq_len=%d, emb_dim=40)" % w["T"],
        "value": cb["value"], "unit": "user-sequences/s", "n_gpus": world, "steps": a.steps, "warmup": wu,
        "ms_per_step": cb["sec_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Taobao-shaped synthetic: seq_len=%d emb_dim=40 (32+8) batch=%d user-sequences per GPU "
                               "(CPU arm: %d sequences per step = BASELINE config 1), %d items / %d cates / %d users, "
                               "optimizer=%s" % (w["T"], S, a.ref_seqs, w["n_items"], w["n_cates"], w["n_users"], a.optimizer),
                   "parallelism": "cpu x%d threads" % cb["cores"]},
        "cpu_baseline": cb,
        "other_optimizer": alt,
        "e2e": {"value": cb["value"], "unit": "user-sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


if __name__ == "__main__":
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    # keep stdout to the one JSON line: NCCL prints its version banner (and warnings) to stdout unless told otherwise
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_%h_%p.log")
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
