"""CPU ORACLE for the CLSR training step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path (``clsr_b200`` and
``reco_utils``) never does; it fails loudly when the CUDA library is missing.

What it is: an op-by-op restatement, on PyTorch-CPU tensors (fp64 "truth" or fp32
"reference precision"), of the graph the reference builds with TensorFlow 1.15:

  * reco_utils/recommender/deeprec/models/sequential/sequential_base_model.py:55-74,381-461
  * reco_utils/recommender/deeprec/models/sequential/clsr.py:22-82,103-277,343-381
  * reco_utils/recommender/deeprec/models/base_model.py:118-159,191-247,281-297,627-708
  * reco_utils/recommender/deeprec/models/sequential/rnn_cell_implement.py:129-298 (Time4LSTMCell)

The arithmetic itself lives in TensorFlow 1.15.2 (pinned in prose only, README.md:7;
not vendored, not installable here: Python 3.12, no network).  The TF-internal
semantics restated here (GRUCell gate layout, dynamic_rnn length masking, non-fused
batch_normalization, IndexedSlices concat -> clip_by_norm -> Unique/UnsortedSegmentSum
-> non-lazy sparse Adam) were confirmed against the reference's own shipped artifact
(``epoch_3.meta`` MetaGraphDef) and its parameterisation is
pinned by the shipped checkpoint (names/shapes/values load 1:1, tests/test_oracle.py).

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures with numeric
outputs for this path (tests/__init__.py is empty), and TF1.15 cannot run here, so no
output of this oracle has been compared with an output of the reference itself.

Backward is ``torch.autograd`` over the restated forward; sparse-table gradients are
recovered per lookup site (as TF's IndexedSlices) so the clip/Adam quirks can be
followed exactly.
"""
import math
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch

SC = "sequential/clsr/"
EMB = "sequential/embedding/"


@dataclass
class OracleConfig:
    """Effective hparams (clsr.yaml overridden by examples/00_quick_start/sequential.py:36-68)."""
    max_seq_length: int = 50
    item_embedding_dim: int = 32
    cate_embedding_dim: int = 8
    user_embedding_dim: int = 40
    hidden_size: int = 40
    att_fcn_layer_sizes: List[int] = field(default_factory=lambda: [80, 40])
    layer_sizes: List[int] = field(default_factory=lambda: [100, 64])
    train_num_ngs: int = 4
    embed_l2: float = 1e-6
    layer_l2: float = 1e-6
    contrastive_loss: str = "triplet"
    triplet_margin: float = 1.0
    contrastive_loss_weight: float = 0.1
    discrepancy_loss_weight: float = 0.01
    contrastive_length_threshold: int = 5
    contrastive_recent_k: int = 3
    learning_rate: float = 1e-3
    is_clip_norm: int = 1
    max_grad_norm: float = 2.0
    optimizer: str = "adam"          # "adam" (TF non-lazy sparse Adam) | "lazyadam"
    bn_momentum: float = 0.95        # base_model.py:676
    bn_eps: float = 1e-4             # base_model.py:677
    beta1: float = 0.9
    beta2: float = 0.999
    adam_eps: float = 1e-8
    # graph variants (clsr.py:159-274)
    interest_evolve: bool = True
    predict_long_short: bool = True
    manual_alpha: bool = False
    manual_alpha_value: float = 0.5
    sequential_model: str = "time4lstm"   # | "lstm" | "gru" (clsr.py:179-218)


def _t(x, dtype):
    return torch.as_tensor(np.asarray(x)).to(dtype)


_DIST = None  # set by data_parallel(): object with .all_reduce(t) (autograd-aware sum) and .world


class data_parallel:
    """Context manager for the sharded-batch tests: inside it, BatchNorm statistics and the loss
    normalisers are taken over the batch of ALL ranks (single-device semantics), which is the set of
    exchanges the multi-GPU engine performs."""

    def __init__(self, ctx):
        self.ctx = ctx

    def __enter__(self):
        global _DIST
        _DIST = self.ctx

    def __exit__(self, *a):
        global _DIST
        _DIST = None


def _bn(x, p, prefix, train, cfg, stats):
    """tf.layers.batch_normalization, non-fused (base_model.py:673-679).  Statistics run
    over every axis but the last, padded positions included."""
    gamma, beta = p[prefix + "gamma"], p[prefix + "beta"]
    if train and _DIST is not None:
        flat = x.reshape(-1, x.shape[-1])
        n = flat.shape[0] * _DIST.world
        mean = _DIST.all_reduce(flat.sum(0)) / n
        var = _DIST.all_reduce((flat ** 2).sum(0)) / n - mean ** 2
        stats[prefix] = (mean.detach(), var.detach())
    elif train:
        flat = x.reshape(-1, x.shape[-1])
        mean = flat.mean(0)
        var = ((flat - mean) ** 2).mean(0)          # biased (tf.nn.moments)
        stats[prefix] = (mean.detach(), var.detach())
    else:
        mean, var = p[prefix + "moving_mean"], p[prefix + "moving_variance"]
    inv = torch.rsqrt(var + cfg.bn_eps) * gamma
    return x * inv + (beta - mean * inv)


_RELU_MASKS = None  # set by relu_decisions(): {(tag, layer): 0/1 tensor shaped like that layer's activations}


class relu_decisions:
    """Context manager for the parity tests: inside it, the ReLU of MLP layer (tag, i) is applied as
    ``bn(h) * mask`` with the given 0/1 mask instead of ``relu(bn(h))``.  A unit whose pre-activation lies
    within rounding distance of zero can come out on either side of the kink in two correct
    implementations; its forward value is ~0 either way, but its derivative is 0 or 1.  Feeding the
    decisions of the implementation under test into the oracle separates that effect from real errors.
    ``flips`` collects, per layer, how many decisions differ from the oracle's own."""

    def __init__(self, masks):
        self.masks = masks
        self.flips = {}

    def __enter__(self):
        global _RELU_MASKS
        _RELU_MASKS = self
        return self

    def __exit__(self, *a):
        global _RELU_MASKS
        _RELU_MASKS = None


def _relu(y, tag, i):
    ctx = _RELU_MASKS
    if ctx is None or (tag, i) not in ctx.masks:
        return torch.relu(y)
    m = torch.as_tensor(ctx.masks[(tag, i)]).reshape(y.shape)
    own = y.detach() > 0
    ctx.flips[(tag, i)] = int((own != (m > 0)).sum())
    return y * m.to(y.dtype)


def _fcn_net(x, p, scope, sizes, train, cfg, stats, inter=None, tag=""):
    """BaseModel._fcn_net (base_model.py:627-708): (xW+b -> BN -> ReLU) per hidden layer,
    then a linear output unit."""
    h = x
    for i, _ in enumerate(sizes):
        h = h @ p[scope + "w_nn_layer%d" % i] + p[scope + "b_nn_layer%d" % i]
        if inter is not None:
            inter[tag + "_h%d" % i] = h
        bn = scope + ("batch_normalization/" if i == 0 else "batch_normalization_%d/" % i)
        h = _relu(_bn(h, p, bn, train, cfg, stats), tag, i)
    return h @ p[scope + "w_nn_output"] + p[scope + "b_nn_output"]


def _attention_fcn(query, keys, mask_b, p, scope, train, cfg, stats, inter, tag):
    """CLSRModel._attention_fcn (clsr.py:343-381)."""
    a = keys @ p[scope + "attention_mat"]                      # tensordot over last axis
    q = query.unsqueeze(1).expand(-1, a.shape[1], -1)
    feat = torch.cat([a, q, a - q, a * q], -1)
    sc = _fcn_net(feat, p, scope + "att_fcn/nn_part/", cfg.att_fcn_layer_sizes, train, cfg, stats,
                  inter, tag).squeeze(-1)
    pad = torch.full_like(sc, float(-(2 ** 32) + 1))
    w = torch.softmax(torch.where(mask_b, sc, pad), dim=-1)
    inter[tag + "_a"], inter[tag + "_score"], inter[tag + "_w"] = a, sc, w
    return keys * w.unsqueeze(-1)


def _gru(x, length, h0, p, scope):
    """tf.nn.dynamic_rnn(GRUCell) (clsr.py:161-168, 230-236): gates = sigmoid([x,h]Wg+bg)
    split r,u; c = tanh([x, r*h]Wc+bc); h' = u*h + (1-u)*c; state copied through for
    t >= length.  Returns the final state."""
    wg, bg = p[scope + "gates/kernel"], p[scope + "gates/bias"]
    wc, bc = p[scope + "candidate/kernel"], p[scope + "candidate/bias"]
    h = h0
    units = h0.shape[1]
    for t in range(x.shape[1]):
        xt = x[:, t]
        g = torch.sigmoid(torch.cat([xt, h], 1) @ wg + bg)
        r, u = g[:, :units], g[:, units:]
        c = torch.tanh(torch.cat([xt, r * h], 1) @ wc + bc)
        hn = u * h + (1 - u) * c
        h = torch.where((t < length).unsqueeze(1), hn, h)
    return h


def _time4lstm(x, t_last, t_now, length, p, scope, hidden):
    """dynamic_rnn(Time4LSTMCell) (clsr.py:179-200; rnn_cell_implement.py:129-298).
    ``t_now`` = inputs[:, -1] = time_to_now, ``t_last`` = inputs[:, -2] =
    time_from_first_action.  Outputs are zero for t >= length."""
    g = lambda n: p[scope + n]
    B = x.shape[0]
    c = torch.zeros(B, hidden, dtype=x.dtype)
    m = torch.zeros(B, hidden, dtype=x.dtype)
    outs = []
    for t in range(x.shape[1]):
        xt = x[:, t]
        tn = torch.tanh(t_now[:, t:t + 1] * g("_time_input_w1") + g("_time_input_bias1"))
        tl = torch.tanh(t_last[:, t:t + 1] * g("_time_input_w2") + g("_time_input_bias2"))
        s_now = xt @ g("_time_kernel_w1") + tn @ g("_time_kernel_t1") + g("_time_bias1")
        s_last = xt @ g("_time_kernel_w2") + tl @ g("_time_kernel_t2") + g("_time_bias2")
        lm = torch.cat([xt, m], 1) @ g("kernel") + g("bias")
        i, j, f, o = torch.split(lm, hidden, dim=1)
        o = o + tn @ g("_o_kernel_t1") + tl @ g("_o_kernel_t2")
        cn = torch.sigmoid(f + 1.0) * torch.sigmoid(s_last) * c + \
            torch.sigmoid(i) * torch.sigmoid(s_now) * torch.tanh(j)
        mn = torch.sigmoid(o) * torch.tanh(cn)
        live = (t < length).unsqueeze(1)
        outs.append(torch.where(live, mn, torch.zeros_like(mn)))
        c = torch.where(live, cn, c)
        m = torch.where(live, mn, m)
    return torch.stack(outs, 1)


def _lstm_seq(x, length, p, scope, hidden):
    """dynamic_rnn(tf.nn.rnn_cell.LSTMCell(hidden)) (clsr.py:209-216): gates i, j, f, o of [x, m].kernel + bias,
    forget_bias 1.0, no peepholes; outputs zero / state copied through for t >= length."""
    B = x.shape[0]
    c = torch.zeros(B, hidden, dtype=x.dtype)
    m = torch.zeros(B, hidden, dtype=x.dtype)
    outs = []
    for t in range(x.shape[1]):
        lm = torch.cat([x[:, t], m], 1) @ p[scope + "kernel"] + p[scope + "bias"]
        i, j, f, o = torch.split(lm, hidden, dim=1)
        cn = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
        mn = torch.sigmoid(o) * torch.tanh(cn)
        live = (t < length).unsqueeze(1)
        outs.append(torch.where(live, mn, torch.zeros_like(mn)))
        c = torch.where(live, cn, c)
        m = torch.where(live, mn, m)
    return torch.stack(outs, 1)


def _gru_seq(x, length, p, scope, hidden):
    """dynamic_rnn(tf.nn.rnn_cell.GRUCell(hidden)) outputs (clsr.py:201-208); cell as in _gru."""
    wg, bg = p[scope + "gates/kernel"], p[scope + "gates/bias"]
    wc, bc = p[scope + "candidate/kernel"], p[scope + "candidate/bias"]
    h = torch.zeros(x.shape[0], hidden, dtype=x.dtype)
    outs = []
    for t in range(x.shape[1]):
        xt = x[:, t]
        g = torch.sigmoid(torch.cat([xt, h], 1) @ wg + bg)
        r, u = g[:, :hidden], g[:, hidden:]
        c = torch.tanh(torch.cat([xt, r * h], 1) @ wc + bc)
        hn = u * h + (1 - u) * c
        live = (t < length).unsqueeze(1)
        outs.append(torch.where(live, hn, torch.zeros_like(hn)))
        h = torch.where(live, hn, h)
    return torch.stack(outs, 1)


def forward(params, batch, cfg, train, dtype=torch.float64, leaves=None):
    """Forward pass.  ``params``: {TF variable name: tensor}.  ``batch``: the feed_dict
    contents keyed by placeholder name (sequential_iterator.py:48-70, 517).
    ``leaves``: if a dict, every embedding lookup site is made an autograd leaf and
    stored there (site name -> (table name, flat index tensor, rows tensor))."""
    p = params
    it = lambda k: torch.as_tensor(np.asarray(batch[k]).astype(np.int64))
    users, items, cates = it("users"), it("items"), it("cates")
    ih, ch = it("item_history"), it("item_cate_history")
    mask_i = it("mask")
    tfa, ttn = _t(batch["time_from_first_action"], dtype), _t(batch["time_to_now"], dtype)
    inter, stats = {}, {}

    def lookup(site, table, idx):
        rows = p[EMB + table][idx]
        if leaves is not None:
            rows = rows.detach().clone().requires_grad_(True)
            leaves[site] = (table, idx.reshape(-1), rows)
        return rows

    # sequential_base_model.py:381-437, clsr.py:103-127 (dropout keep=1.0 is identity)
    item_e = lookup("item_target", "item_embedding", items)
    item_h = lookup("item_history", "item_embedding", ih)
    cate_e = lookup("cate_target", "cate_embedding", cates)
    cate_h = lookup("cate_history", "cate_embedding", ch)
    inv_items = torch.unique(torch.cat([ih.reshape(-1), items.reshape(-1)]))
    inv_cates = torch.unique(torch.cat([ch.reshape(-1), cates.reshape(-1)]))
    inv_users = torch.unique(users.reshape(-1))
    embed_params = [lookup("item_involved", "item_embedding", inv_items),
                    lookup("cate_involved", "cate_embedding", inv_cates)]
    target = torch.cat([item_e, cate_e], -1)
    ul = lookup("user_long", "user_long_embedding", users)
    us = lookup("user_short", "user_short_embedding", users)
    inv_ul = lookup("user_long_involved", "user_long_embedding", inv_users)
    inv_us = lookup("user_short_involved", "user_short_embedding", inv_users)
    embed_params += [inv_ul, inv_us]

    # clsr.py:145-150
    hist = torch.cat([item_h, cate_h], 2)
    real_mask = mask_i.to(dtype)
    length = mask_i.sum(1)
    mask_b = mask_i == 1
    H = cfg.hidden_size

    # long term (clsr.py:152-157)
    att_long = _attention_fcn(ul, hist, mask_b, p, SC + "long_term/attention_fcn/", train, cfg, stats,
                              inter, "long")
    afl = att_long.sum(1)
    hist_mean = (hist * real_mask.unsqueeze(-1)).sum(1) / real_mask.sum(1, keepdim=True)

    # short term (clsr.py:159-222)
    if cfg.interest_evolve:   # clsr.py:160-170
        sti = _gru(hist, length, us, p, SC + "short_term/short_term_intention/gru_cell/")
    else:
        sti = us
    position = torch.flip(torch.cumsum(torch.flip(real_mask, [1]), 1), [1])
    recent = ((position >= 1) & (position <= cfg.contrastive_recent_k)).to(dtype)
    hist_recent = (hist * recent.unsqueeze(-1)).sum(1) / recent.sum(1, keepdim=True)
    if cfg.sequential_model == "time4lstm":
        rnn_out = _time4lstm(hist, tfa, ttn, length, p, SC + "short_term/time4lstm/time4lstm_cell/", H)
    elif cfg.sequential_model == "lstm":
        rnn_out = _lstm_seq(hist, length, p, SC + "short_term/simple_lstm/lstm_cell/", H)
    elif cfg.sequential_model == "gru":
        rnn_out = _gru_seq(hist, length, p, SC + "short_term/simple_gru/gru_cell/", H)
    else:
        raise ValueError(cfg.sequential_model)
    sq = torch.cat([sti, target], -1)
    att_short = _attention_fcn(sq, rnn_out, mask_b, p, SC + "short_term/attention_fcn/", train, cfg, stats,
                               inter, "short")
    afs = att_short.sum(1)

    # alpha (clsr.py:225-275)
    fs = None
    if not cfg.manual_alpha:
        if cfg.predict_long_short:
            fs = _gru(hist, length, torch.zeros(hist.shape[0], H, dtype=dtype), p, SC + "causal2/causal2/gru_cell/")
            concat_all = torch.cat([fs, target, afl, afs, ttn[:, -1:]], 1)
        else:
            concat_all = torch.cat([target, afl, afs, ttn[:, -1:]], 1)
        alpha_logit = _fcn_net(concat_all, p, SC + "fcn_alpha/nn_part/", cfg.att_fcn_layer_sizes, train, cfg,
                               stats, inter, "alpha")
        alpha = torch.sigmoid(alpha_logit)
        user_embed = afl * alpha + afs * (1.0 - alpha)
    else:   # clsr.py:272-274
        alpha = torch.full((hist.shape[0], 1), cfg.manual_alpha_value, dtype=dtype)
        user_embed = afl * cfg.manual_alpha_value + afs * (1.0 - cfg.manual_alpha_value)
    model_output = torch.cat([user_embed, target], 1)
    logit = _fcn_net(model_output, p, "sequential/logit_fcn/nn_part/", cfg.layer_sizes, train, cfg, stats,
                     inter, "logit")

    inter.update(hist=hist, target=target, ul=ul, us=us, afl=afl, afs=afs, hist_mean=hist_mean,
                 hist_recent=hist_recent, sti=sti, rnn_out=rnn_out, fs=fs, alpha=alpha,
                 user_embed=user_embed, logit=logit, length=length)
    return dict(logit=logit, pred=torch.sigmoid(logit), alpha=alpha, inter=inter, bn_stats=stats,
                embed_params=embed_params, inv_ul=inv_ul, inv_us=inv_us,
                involved=dict(items=inv_items, cates=inv_cates, users=inv_users))


def losses(out, params, batch, cfg, dtype=torch.float64):
    """CLSRModel._get_loss (clsr.py:22-82) + BaseModel data/regular losses
    (base_model.py:118-130,215-247)."""
    I = out["inter"]
    G = cfg.train_num_ngs + 1
    logits = out["logit"].reshape(-1, G)
    labels = _t(batch["labels"], dtype).reshape(-1, G)
    sm = torch.softmax(logits, -1)
    pos = torch.where(labels == 1, sm, torch.ones_like(sm))
    if _DIST is not None:   # mean over the rows of all ranks
        data_loss = -G * torch.log(pos).sum() / (pos.numel() * _DIST.world)
    else:
        data_loss = -G * torch.log(pos).mean()

    reg = torch.zeros((), dtype=dtype)
    for e in out["embed_params"]:
        reg = reg + cfg.embed_l2 * 0.5 * (e ** 2).sum()
    for name, v in params.items():
        if name.startswith(EMB) or "moving_" in name:
            continue
        reg = reg + cfg.layer_l2 * 0.5 * (v ** 2).sum()

    cm = (I["length"] > cfg.contrastive_length_threshold).to(dtype)
    afl, afs, hm, hr = I["afl"], I["afs"], I["hist_mean"], I["hist_recent"]
    den = cm.sum()
    if _DIST is not None:
        den = _DIST.all_reduce(den.clone()).detach()
    if cfg.contrastive_loss == "bpr":
        sp = torch.nn.functional.softplus
        l1 = (cm * sp((afl * (-hm + hr)).sum(-1))).sum() / den
        l2 = (cm * sp((afs * (-hr + hm)).sum(-1))).sum() / den
        l3 = (cm * sp((hm * (-afl + afs)).sum(-1))).sum() / den
        l4 = (cm * sp((hr * (-afs + afl)).sum(-1))).sum() / den
    else:
        mg = cfg.triplet_margin
        dlm, dlr = (afl - hm) ** 2, (afl - hr) ** 2
        dsm, dsr = (afs - hm) ** 2, (afs - hr) ** 2
        l1 = (cm * torch.clamp(dlm - dlr + mg, min=0).sum(-1)).sum() / den
        l2 = (cm * torch.clamp(dsr - dsm + mg, min=0).sum(-1)).sum() / den
        l3 = (cm * torch.clamp(dlm - dsm + mg, min=0).sum(-1)).sum() / den
        l4 = (cm * torch.clamp(dsr - dlr + mg, min=0).sum(-1)).sum() / den
    con = cfg.contrastive_loss_weight * (l1 + l2 + l3 + l4)
    disc = -cfg.discrepancy_loss_weight * ((out["inv_ul"].reshape(-1) - out["inv_us"].reshape(-1)) ** 2).mean()
    total = data_loss + reg + con + disc
    return dict(loss=total, data_loss=data_loss, regular_loss=reg, contrastive_loss=con,
                discrepancy_loss=disc)


TABLES = ["item_embedding", "cate_embedding", "user_long_embedding", "user_short_embedding"]


def compute_gradients(params, batch, cfg, dtype=torch.float64, retain=()):
    """optimizer.compute_gradients (base_model.py:289).  Returns dense grads by name and,
    per table, the concatenated IndexedSlices (indices, values) as TF forms them."""
    p = {}
    for k, v in params.items():
        tv = _t(v, dtype)
        if not k.startswith(EMB) and "moving_" not in k:
            tv.requires_grad_(True)
        p[k] = tv
    leaves = {}
    out = forward(p, batch, cfg, True, dtype, leaves)
    for k in retain:
        out["inter"][k].retain_grad()
    L = losses(out, p, batch, cfg, dtype)
    L["loss"].backward()
    dense = {k: v.grad.detach() for k, v in p.items() if v.requires_grad}
    slices = {}
    for t in TABLES:
        idx = [ix for (tab, ix, rows) in leaves.values() if tab == t]
        val = [rows.grad.reshape(-1, rows.shape[-1]) for (tab, ix, rows) in leaves.values() if tab == t]
        slices[t] = (torch.cat(idx), torch.cat(val))
    inter_grads = {k: out["inter"][k].grad for k in retain}
    return out, L, dense, slices, inter_grads


def _clip(values, cfg):
    """tf.clip_by_norm (base_model.py:290-296): t*clip/max(||t||2, clip); for
    IndexedSlices the norm runs over the concatenated, not yet deduplicated values."""
    if not cfg.is_clip_norm:
        return values
    n = torch.sqrt((values ** 2).sum())
    return values * cfg.max_grad_norm / torch.maximum(n, torch.tensor(cfg.max_grad_norm, dtype=values.dtype))


def train_step(params, slots, batch, cfg, step, dtype=torch.float64):
    """One CLSRModel.train call (clsr.py:383-408): forward, losses, gradients,
    per-variable clip, Adam (dense ApplyAdam / non-lazy sparse path), BN moving-stat
    update.  ``params`` and ``slots`` ({name: (m, v)}) are numpy dicts updated in place;
    ``step`` is the 1-based Adam step.  Returns (loss dict of floats, aux)."""
    out, L, dense, slices, _ = compute_gradients(params, batch, cfg, dtype)
    lr_t = cfg.learning_rate * math.sqrt(1 - cfg.beta2 ** step) / (1 - cfg.beta1 ** step)
    aux = {"norms": {}}

    def adam(name, g):
        var = _t(params[name], dtype)
        m, v = slots.setdefault(name, (np.zeros_like(params[name]), np.zeros_like(params[name])))
        m, v = _t(m, dtype), _t(v, dtype)
        m = cfg.beta1 * m + (1 - cfg.beta1) * g
        v = cfg.beta2 * v + (1 - cfg.beta2) * g * g
        var = var - lr_t * m / (torch.sqrt(v) + cfg.adam_eps)
        np_dt = params[name].dtype
        params[name] = var.numpy().astype(np_dt)
        slots[name] = (m.numpy().astype(np_dt), v.numpy().astype(np_dt))

    for name, g in dense.items():
        aux["norms"][name] = float(torch.sqrt((g ** 2).sum()))
        adam(name, _clip(g, cfg))
    for t in TABLES:
        idx, val = slices[t]
        aux["norms"][EMB + t] = float(torch.sqrt((val ** 2).sum()))
        val = _clip(val, cfg)
        name = EMB + t
        if cfg.optimizer == "lazyadam":
            uniq, invx = torch.unique(idx, return_inverse=True)
            gs = torch.zeros(len(uniq), val.shape[1], dtype=dtype).index_add_(0, invx, val)
            var = _t(params[name], dtype)
            m, v = slots.setdefault(name, (np.zeros_like(params[name]), np.zeros_like(params[name])))
            m, v = _t(m, dtype), _t(v, dtype)
            m[uniq] = cfg.beta1 * m[uniq] + (1 - cfg.beta1) * gs
            v[uniq] = cfg.beta2 * v[uniq] + (1 - cfg.beta2) * gs * gs
            var[uniq] = var[uniq] - lr_t * m[uniq] / (torch.sqrt(v[uniq]) + cfg.adam_eps)
            np_dt = params[name].dtype
            params[name] = var.numpy().astype(np_dt)
            slots[name] = (m.numpy().astype(np_dt), v.numpy().astype(np_dt))
        else:
            # Unique + UnsortedSegmentSum, then m,v decayed and var updated over ALL rows
            g = torch.zeros(params[name].shape, dtype=dtype).index_add_(0, idx, val)
            adam(name, g)
    # UPDATE_OPS: moving -= (moving - batch) * (1 - momentum)
    for prefix, (mean, var) in out["bn_stats"].items():
        for suffix, b in (("moving_mean", mean), ("moving_variance", var)):
            mv = _t(params[prefix + suffix], dtype)
            mv = mv - (mv - b) * (1 - cfg.bn_momentum)
            params[prefix + suffix] = mv.numpy().astype(params[prefix + suffix].dtype)
    return {k: float(v) for k, v in L.items()}, aux


def predict(params, batch, cfg, dtype=torch.float64):
    """eval / infer path (base_model.py:366-392): BN uses moving statistics."""
    with torch.no_grad():
        p = {k: _t(v, dtype) for k, v in params.items()}
        out = forward(p, batch, cfg, False, dtype)
    return out
