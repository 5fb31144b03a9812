#!/usr/bin/env python
"""Benchmark of the CLSR training step (BASELINE.json: user-sequences/sec at seq_len=50,
emb_dim=40; HBM GB/s on the embedding gather).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [--steps K] [--warmup W]      # reference CPU path (restated)

One step = one CLSRModel.train call (forward, losses, backward, sparse-gradient scatter-add,
per-variable clip, Adam incl. the TF non-lazy table sweep, BN moving-stat update) over one
synthetic Taobao-shaped batch: 4096 user-sequences (20480 graph rows), T=50, D=32+8,
4.0M items / 9.4k categories / 1.0M users.  Prints ONE JSON line (rank 0).
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
}
G = 5  # 1 + train_num_ngs


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="taobao", choices=sorted(WORKLOADS))
    ap.add_argument("--seqs", type=int, default=0, help="user-sequences per step per GPU (default: workload's)")
    ap.add_argument("--optimizer", default="adam", choices=["adam", "lazyadam"])
    ap.add_argument("--ref-seqs", type=int, default=256, help="bounded CPU sample: sequences per CPU step")
    ap.add_argument("--math", default="tc", choices=["tc", "simt"],
                    help="tc: large GEMMs on tcgen05 (split-bf16, fp32 accumulate); simt: fp32 CUDA cores everywhere")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default="")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_tables(w, seed):
    """Truncated-normal(0.01) tables (base_model.py:161-165) as CPU torch tensors."""
    import torch
    g = torch.Generator().manual_seed(seed)
    shapes = {"item_embedding": (w["n_items"], 32), "cate_embedding": (w["n_cates"], 8),
              "user_long_embedding": (w["n_users"], 40), "user_short_embedding": (w["n_users"], 40)}
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
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
    (profiles/r01_ncu_top_kernels.json; Taobao workload only), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_ncu_top_kernels.json")) as f:
            t = json.load(f)["traffic"]
    except Exception:
        return None
    key = {"adam_sweep": "adam_sweep (4 tables, one step)", "gather_hist": "gather_hist_kernel<4>",
           "scatter_hist": "scatter_hist_kernel"}.get(name)
    if key not in t:
        return None
    return t[key].get("dram_bytes", t[key].get("dram_bytes_per_launch_mean"))


def algorithmic_bytes(name, S, B, T, w):
    """Algorithmic HBM bytes of one launch of the named kernel (kernel table)."""
    M, MB, D, A0, A1, NX, Q, U = S * T, B * T, 40, 80, 40, 480, 80, 40
    tab = {
        "gather_hist": M * (8 + 2 * D * 4),                    # T*8 + 2*T*D*e per sequence
        "scatter_hist": M * (8 + D * 4 + D * 4),               # read d_hist + ids, RMW-add one row slice
        "adam_sweep": None,
        "h0s": MB * (A0 * 4) + M * (D + A0) * 4,               # write h0s, read a2 + invs once
        "h1s": MB * (A0 + A1) * 4,
        "dy0s": MB * (2 * A1 + 2 * A0) * 4,                    # read dy1s,h1s,h0s; write dy0s
        "dP": MB * (2 * A0 + D) * 4,
        "dW1s": MB * (A0 + 2 * A1) * 4,
        "dWs0t": MB * (2 * A0) * 4 + M * D * 4,
        "pool_bwd_short": MB * (2 * A1) * 4 + M * 2 * D * 4,
        "pool_fwd_short": MB * A1 * 4 + M * D * 4,
        "px": M * (D + NX) * 4,
        "dX": M * (NX + D) * 4,
    }
    return tab.get(name)


def step_roofline(T, seq_per_s_per_gpu, hbm_gbs, tensor_tflops=1388.2):
    """Whole-step bounds of SURVEY.md 8d / BASELINE.md section 3 for the D=40 model: minimum HBM bytes and
    (group-deduplicated) FLOPs per user-sequence, and the fraction of each bound the measured rate reaches."""
    D = U = H = 40
    A0, A1, L0, L1, Gr = 80, 40, 100, 64, G
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


def run_b200(a):
    import torch
    import torch.distributed as dist
    from clsr_b200 import build, params as P, synth
    from clsr_b200.engine import Engine, normalize_feed

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
    B, T = S * G, w["T"]
    eng = Engine(w["n_items"], w["n_cates"], w["n_users"], max_rows=B, seq_len=T, train_group=G,
                 optimizer=a.optimizer, device=local, math_mode=1 if a.math == "tc" else 0)
    dense = P.init_params(1, 1, 1, seed=42, tables=False)
    eng.set_dense(dense)
    tabs = make_tables(w, 42)
    for t, name in __import__("clsr_b200.engine", fromlist=["TABLE_VARS"]).TABLE_VARS.items():
        eng.tables[t].copy_(tabs[name])
    if world > 1:
        eng.comm_init(rank, world, dist)
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

    # ---- device-resident arm (value) ----
    for i in range(a.warmup):
        eng.train_step(dev[i % NB], group=G, on_device=True, wait=False)
    sync_all()
    eng.set_profiling(True)
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(a.steps):
            eng.train_step(dev[(a.warmup + i) % NB], group=G, on_device=True, wait=False)
        e1.record(stream)
    sync_all()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = eng.kernel_launches() * a.steps
    prof = eng.profile()
    eng.set_profiling(False)
    clocks = sampler.stop(t0, t1) if sampler else None
    # ---- end-to-end arm: host (pinned) feed -> C ABI -> losses read back, every step ----
    for i in range(min(a.warmup, 3)):
        eng.train_step(pinned[i % NB], group=G, normalized=True)
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e2.record(stream)
        for i in range(a.steps):
            last = eng.train_step(pinned[i % NB], group=G, normalized=True)
        e3.record(stream)
    sync_all()
    ms_e2e = e2.elapsed_time(e3)
    # ---- standalone history gather (K1+K3 through clsr_gather_history) at a size where launch ramp does not
    # matter: 32 batches' worth of the same zipf windows, one launch, timed with events on the engine stream ----
    gather_big = None
    if world == 1:
        try:
            ih = torch.cat([d_["item_history"][::G] for d_ in dev] * 8).contiguous()
            ch = torch.cat([d_["item_cate_history"][::G] for d_ in dev] * 8).contiguous()
            npos = ih.numel()
            gout = torch.empty(npos, 40, device="cuda")
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                for _ in range(3):
                    eng._check(eng.lib.clsr_gather_history(eng.h, ih.data_ptr(), ch.data_ptr(), npos, gout.data_ptr()))
                g0.record(stream)
                for _ in range(10):
                    eng._check(eng.lib.clsr_gather_history(eng.h, ih.data_ptr(), ch.data_ptr(), npos, gout.data_ptr()))
                g1.record(stream)
            sync_all()
            gms = g0.elapsed_time(g1) / 10
            gbytes = npos * (8 + 2 * 40 * 4)
            gather_big = {"positions": npos, "ms": gms, "algorithmic_bytes": gbytes, "achieved": gbytes / 1e9 / (gms / 1e3),
                          "unit": "GB/s", "bound": "hbm",
                          "note": "clsr_gather_history on 32 batches of zipf windows in one launch (T*8 + 2*T*D*4 bytes per window)"}
            del gout, ih, ch
        except Exception as ex:  # the headline numbers do not depend on this extra measurement
            gather_big = {"error": str(ex)}
    if world > 1:
        tt = torch.tensor([ms, ms_e2e], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(tt[0]), float(tt[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    h2d = 5 * S * T * 4 + S * 4 + 3 * B * 4
    hbm_peak, peak_src = peaks()
    if gather_big and "achieved" in gather_big:
        gather_big.update(peak=hbm_peak, frac=gather_big["achieved"] / hbm_peak, peak_source=peak_src)
    per_kernel = {k: {"ms": v[0] / max(v[1], 1), "calls_per_step": v[1] / a.steps,
                      "share": v[0] / max(sum(x[0] for x in prof.values()), 1e-9)} for k, v in prof.items()}
    top = max(per_kernel, key=lambda k: per_kernel[k]["ms"] * per_kernel[k]["calls_per_step"])
    def roof(name):
        k = per_kernel.get(name)
        if not k:
            return None
        nbytes = algorithmic_bytes(name, S, B, T, w)
        if name == "adam_sweep":
            rows = (w["n_items"] * 32 + w["n_cates"] * 8 + 2 * w["n_users"] * 40) * 4 * 6 + \
                   (w["n_items"] + w["n_cates"] + 2 * w["n_users"]) * 4
            gbs = rows / 1e9 / (k["ms"] * k["calls_per_step"] / 1e3)
            return {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                    "traffic": ncu_traffic(name) if a.workload == "taobao" else None, "kernel": name,
                    "peak_source": peak_src, "algorithmic_bytes": rows, "ms": k["ms"] * k["calls_per_step"],
                    "note": "TF non-lazy Adam sweep, the four table launches of one step taken together"}
        if nbytes is None:
            return {"bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None,
                    "traffic": None, "kernel": name, "peak_source": peak_src}
        gbs = nbytes / 1e9 / (k["ms"] / 1e3)
        return {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                "traffic": ncu_traffic(name) if a.workload == "taobao" else None, "kernel": name,
                "peak_source": peak_src, "algorithmic_bytes": nbytes, "ms": k["ms"]}
    out = {
        "metric": "user-sequences/sec (CLSR training step, seq_len=%d, emb_dim=40)" % T,
        "value": S * world * a.steps / (ms / 1e3), "unit": "user-sequences/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 (split-bf16 tcgen05 MMA, fp32 accumulate) + f32" if a.math == "tc" else "f32", "data": "synthetic",
        "config": {"workload": "Taobao-shaped synthetic: seq_len=%d emb_dim=40 (32+8) batch=%d user-sequences "
                               "(%d rows) per GPU, %d items / %d cates / %d users, zipf ids, optimizer=%s"
                               % (T, S, B, w["n_items"], w["n_cates"], w["n_users"], a.optimizer),
                   "l2": "not flushed: each step streams >2.5 GB of activations and (adam) 5 GB of table state, "
                         "far above the 126 MB L2; 4 distinct batches rotate",
                   "parallelism": "dp%d" % world},
        "e2e": {"value": S * world * a.steps / (ms_e2e / 1e3), "unit": "user-sequences/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 20, "ms_per_step": ms_e2e / a.steps,
                "api": "clsr_train_step (C ABI) with pinned host feed arrays, losses read back every step"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof(top),
        "kernels": dict({n: roof(n) for n in ("gather_hist", "scatter_hist", "adam_sweep", "h0s", "h1s", "dy0s")
                         if n in per_kernel}, **({"gather_hist_large": gather_big} if gather_big else {})),
        "top_kernels": sorted(((k, round(v["ms"] * v["calls_per_step"], 4)) for k, v in per_kernel.items()),
                              key=lambda x: -x[1])[:12],
        "last_losses": last,
    }
    try:
        out["step_roofline"] = step_roofline(T, S * a.steps / (ms / 1e3), hbm_peak)
    except Exception as ex:  # informational only
        out["step_roofline"] = {"error": str(ex)}
    if a.profile_out:
        with open(a.profile_out, "w") as f:
            json.dump({"per_kernel": per_kernel, "ms_per_step": ms / a.steps}, f, indent=1)
    if world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_arm(a, w, dense, tabs, steps=2, warmup=1)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_arm(a, w, dense, tabs, steps, warmup):
    from clsr_b200 import synth
    from oracle import clsr_oracle as O
    from oracle.cpu_baseline import CpuTrainer, time_steps
    cfg = O.OracleConfig(max_seq_length=w["T"], optimizer=a.optimizer)
    prm = dict(dense)
    prm.update(tabs)
    tr = CpuTrainer(prm, cfg, threads=os.cpu_count())
    src = synth.SyntheticSource(w["n_items"], w["n_cates"], w["n_users"], w["T"], seed=42, time_unit=w["time_unit"])
    batches = [src.batch(a.ref_seqs, G - 1) for _ in range(2)]
    sec = time_steps(tr, batches, steps, warmup)
    return {"value": a.ref_seqs / sec, "unit": "user-sequences/s", "cores": tr.threads, "kind": "port",
            "sample": "%d steps of %d user-sequences (%d rows) of the same workload, fp32 PyTorch-CPU restatement "
                      "of the TF1.15 graph incl. the full-table Adam sweep; TF1.15 itself is not installable"
                      % (steps, a.ref_seqs, a.ref_seqs * G),
            "sec_per_step": sec}


def run_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from clsr_b200 import params as P
    w = dict(WORKLOADS[a.workload])
    S = a.seqs or w["seqs"]
    dense = P.init_params(1, 1, 1, seed=42, tables=False)
    tabs = make_tables(w, 42)
    cb = cpu_arm(a, w, dense, tabs, steps=a.steps, warmup=min(a.warmup, 3))
    world = int(os.environ.get("WORLD_SIZE", 1))
    out = {
        "impl": "reference",
        "metric": "user-sequences/sec (CLSR training step, seq_len=%d, emb_dim=40)" % w["T"],
        "value": cb["value"], "unit": "user-sequences/s", "n_gpus": world, "steps": a.steps, "warmup": min(a.warmup, 3),
        "ms_per_step": cb["sec_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Taobao-shaped synthetic: seq_len=%d emb_dim=40 (32+8) batch=%d user-sequences per GPU "
                               "(CPU arm: bounded sample of %d sequences per step), %d items / %d cates / %d users, "
                               "optimizer=%s" % (w["T"], S, a.ref_seqs, w["n_items"], w["n_cates"], w["n_users"], a.optimizer),
                   "parallelism": "cpu x%d threads" % cb["cores"]},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "user-sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


if __name__ == "__main__":
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
