"""Shared helpers of the parity tests: build a small problem, run the CUDA engine and the
CPU oracle on the same feed, and compare named intermediates."""
import numpy as np

from clsr_b200 import params as P
from clsr_b200 import synth

RETAIN = ["logit", "afl", "afs", "hist_mean", "hist_recent", "sti", "fs", "rnn_out", "hist", "target",
          "ul", "us", "alpha"]


def small_problem(S=24, G=5, T=50, seed=3, n_items=3000, n_cates=40, n_users=200, init_scale=8.0,
                  edge_lengths=True, dims=None, variant=None):
    """dims: (item_dim, cate_dim, user_dim = hidden) for the wide configurations (BASELINE configs 4-5);
    variant: graph flags (interest_evolve / predict_long_short / manual_alpha) deciding which variables exist."""
    vkw = {k: v for k, v in (variant or {}).items() if k != "manual_alpha_value"}
    src = synth.SyntheticSource(n_items=n_items, n_cates=n_cates, n_users=n_users, T=T, seed=seed)
    feed = src.batch(S, G - 1) if G > 1 else src.batch(S, 0)
    if G == 1:
        lab = np.zeros((S, 1), np.float32)
        lab[::5] = 1.0
        feed["labels"] = lab
    if dims:
        prm = P.init_params(n_items, n_cates, n_users, Di=dims[0], Dc=dims[1], U=dims[2], H=dims[2], seed=seed, **vkw)
    else:
        prm = P.init_params(n_items, n_cates, n_users, seed=seed, **vkw)
    return feed, scale_params(prm, seed, init_scale)


def scale_params(prm, seed, init_scale=8.0):
    """Scale up the 0.01-std initial weights so every block contributes visibly to the outputs
    (a freshly initialised model is almost linear and would hide errors)."""
    rng = np.random.default_rng(seed + 1)
    for k, v in prm.items():
        if k.endswith("moving_mean"):
            prm[k] = (0.1 * rng.standard_normal(v.shape)).astype(np.float32)
        elif k.endswith("moving_variance"):
            prm[k] = (0.5 + rng.random(v.shape)).astype(np.float32)
        elif k.endswith("gamma"):
            prm[k] = (0.6 + 0.4 * rng.random(v.shape)).astype(np.float32)
        elif k.endswith("beta") or "b_nn_" in k:
            prm[k] = (0.1 * rng.standard_normal(v.shape)).astype(np.float32)
        elif "embedding" in k or "w_nn_" in k or "attention_mat" in k:
            v *= np.float32(init_scale)
    return prm


def set_lengths(feed, lengths, G):
    """Force the history lengths of the first sequences (edge cases: 1, T, <= threshold)."""
    T = feed["mask"].shape[1]
    for s, L in enumerate(lengths):
        rows = slice(s * G, (s + 1) * G)
        live = (np.arange(T) < L)
        for k in ("item_history", "item_cate_history", "mask"):
            if k == "mask":
                feed[k][rows] = live.astype(feed[k].dtype)
            else:
                col = feed[k][rows].copy()
                col[:, L:] = 0
                col[:, :L] = np.where(col[:, :L] == 0, 1 + (np.arange(L) % 7), col[:, :L])
                feed[k][rows] = col
        for k in ("time_diff", "time_from_first_action", "time_to_now"):
            col = feed[k][rows].copy()
            col[:, L:] = 0.0
            col[:, :L] = np.where(col[:, :L] == 0.0, 0.3, col[:, :L])
            feed[k][rows] = col
    return feed


def make_engine(prm, n_items, n_cates, n_users, max_rows, T=50, G=5, **kw):
    from clsr_b200.engine import Engine
    eng = Engine(n_items, n_cates, n_users, max_rows=max_rows, seq_len=T, train_group=G, **kw)
    eng.set_params(prm)
    return eng


def relerr(got, ref):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    scale = max(np.abs(ref).max(), 1e-30)
    return float(np.abs(got - ref).max() / scale)


def relerr_l2(got, ref):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30))


def oracle_config(G, **kw):
    from oracle.clsr_oracle import OracleConfig
    return OracleConfig(train_num_ngs=G - 1, **kw)


def engine_relu_masks(eng, B, S, T, group):
    """ReLU decisions of the engine's last step for the eight hidden layers, shaped for the oracle
    (row-replicated where the engine computed a layer once per sequence)."""
    c = eng.cfg
    A0, A1, L0, L1 = c.att0, c.att1, c.fc0, c.fc1
    spec = [("long", 0, "h0l", (S, T, A0), True), ("long", 1, "h1l", (S, T, A1), True),
            ("short", 0, "h0s", (B, T, A0), False), ("short", 1, "h1s", (B, T, A1), False),
            ("alpha", 0, "ha0", (B, A0), False), ("alpha", 1, "ha1", (B, A1), False),
            ("logit", 0, "hl0", (B, L0), False), ("logit", 1, "hl1", (B, L1), False)]
    masks = {}
    for tag, i, buf, shp, per_seq in spec:
        if tag == "alpha" and getattr(c, "manual_alpha", 0):
            continue   # graph variant without the alpha MLP
        h = eng.debug(buf, shp)
        sc = eng.debug("bn/%s%d/scale" % (tag, i), (shp[-1],))
        sh = eng.debug("bn/%s%d/shift" % (tag, i), (shp[-1],))
        m = (h.astype(np.float64) * sc + sh) > 0     # sign of the exact fused multiply-add the kernels evaluate
        if per_seq and group > 1:
            m = np.repeat(m, group, axis=0)
        masks[(tag, i)] = m.astype(np.float32)
    return masks


def compare_step(eng, feed, prm, G, group, shapes_only=False, metric=None, dtype=None, mask_flips=False):
    """Run one gradient-only step on the engine and the oracle; return {name: relative error}
    (max-norm by default; ``metric=relerr_l2`` for the backward / gradient entries)."""
    bw_err = metric or relerr
    import torch
    from oracle import clsr_oracle as O
    from clsr_b200.engine import STEP_NO_OPTIMIZER, STEP_NO_BN_UPDATE
    B = feed["users"].shape[0]
    T = feed["mask"].shape[1]
    c = eng.cfg
    cfg = oracle_config(G, max_seq_length=T, item_embedding_dim=c.item_dim, cate_embedding_dim=c.cate_dim,
                        user_embedding_dim=c.user_dim, hidden_size=c.hidden)
    dtype = dtype or torch.float64
    S = B // group
    c = eng.cfg
    D, U, H, Q, A0, A1 = c.item_dim + c.cate_dim, c.user_dim, c.hidden, c.user_dim + c.item_dim + c.cate_dim, c.att0, c.att1
    losses = eng.train_step(feed, group=group, flags=STEP_NO_OPTIMIZER | STEP_NO_BN_UPDATE)
    if mask_flips:
        with O.relu_decisions(engine_relu_masks(eng, B, S, T, group)) as rd:
            out, L, dense, slices, ig = O.compute_gradients(prm, feed, cfg, dtype, retain=RETAIN)
        flips = dict(rd.flips)
    else:
        out, L, dense, slices, ig = O.compute_gradients(prm, feed, cfg, dtype, retain=RETAIN)
        flips = {}
    I = {k: (v.detach().numpy() if hasattr(v, "detach") else v) for k, v in out["inter"].items()}
    g = lambda a: a[::group]
    res = {}
    fw = [("X", (S, T, D), g(I["hist"])), ("tgt", (B, D), I["target"]), ("ul", (S, U), g(I["ul"])),
          ("us", (S, U), g(I["us"])), ("sti", (S, U), g(I["sti"])), ("fs", (S, H), g(I["fs"])),
          ("R", (S, T, H), g(I["rnn_out"])), ("al", (S, T, U), g(I["long_a"])),
          ("h0l", (S, T, A0), g(I["long_h0"])), ("h1l", (S, T, A1), g(I["long_h1"])),
          ("wl", (S, T), g(I["long_w"])), ("afl", (S, D), g(I["afl"])), ("hm", (S, D), g(I["hist_mean"])),
          ("hr", (S, D), g(I["hist_recent"])), ("as", (S, T, Q), g(I["short_a"])),
          ("h0s", (B, T, A0), I["short_h0"]), ("h1s", (B, T, A1), I["short_h1"]), ("ws", (B, T), I["short_w"]),
          ("afs", (B, H), I["afs"]), ("ha0", (B, A0), I["alpha_h0"]), ("ha1", (B, A1), I["alpha_h1"]),
          ("alpha", (B,), I["alpha"].reshape(-1)), ("hl0", (B, c.fc0), I["logit_h0"]),
          ("hl1", (B, c.fc1), I["logit_h1"]), ("logit", (B,), I["logit"].reshape(-1))]
    for name, shp, ref in fw:
        res["fwd/" + name] = relerr(eng.debug(name, shp), ref)
    for k in ("loss", "data_loss", "regular_loss", "contrastive_loss", "discrepancy_loss"):
        res["loss/" + k] = abs(losses[k] - float(L[k])) / max(abs(float(L[k])), 1e-12)
    gs = lambda a: a.numpy().reshape(S, group, *a.shape[1:]).sum(1)
    bw = [("dlogit", (B,), ig["logit"].numpy().reshape(-1)), ("dafs", (B, H), ig["afs"].numpy()),
          ("dafl", (S, D), gs(ig["afl"])), ("dhm", (S, D), gs(ig["hist_mean"])),
          ("dhr", (S, D), gs(ig["hist_recent"])), ("dsti", (S, U), gs(ig["sti"])), ("dfs", (S, H), gs(ig["fs"])),
          ("dR", (S, T, H), gs(ig["rnn_out"])), ("dX", (S, T, D), gs(ig["hist"])),
          ("dtgt", (B, D), ig["target"].numpy()), ("dul", (S, U), gs(ig["ul"])), ("dus", (S, U), gs(ig["us"]))]
    for name, shp, ref in bw:
        res["bwd/" + name] = bw_err(eng.debug(name, shp), ref)
    dg = eng.get_dense(3)
    for name, gref in dense.items():
        res["grad/" + name.replace("sequential/", "")] = bw_err(dg[name], gref.numpy().reshape(-1))
    from clsr_b200.engine import TABLE_VARS
    for t, name in TABLE_VARS.items():
        ids, rows = eng.sparse_grad(t)
        tab = name.split("/")[-1]
        idx, val = slices[tab]
        dense_ref = np.zeros(prm[name].shape, np.float64)
        np.add.at(dense_ref, idx.numpy(), val.numpy())
        got = np.zeros(prm[name].shape, np.float64)
        got[ids] = rows
        res["grad/" + tab] = bw_err(got, dense_ref)
        res["uniq/" + tab] = float(len(ids) != len(np.unique(idx.numpy())))
    for (tag, i), n in flips.items():
        res["flips/%s%d" % (tag, i)] = n
    return res, losses
