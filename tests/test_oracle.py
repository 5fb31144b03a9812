"""The CPU oracle: structure pinned by the reference's shipped checkpoint, gradients checked against
finite differences, outputs pinned against a committed regression fixture."""
import json
import os

import numpy as np
import pytest
import torch

from clsr_b200 import params as P
from clsr_b200 import synth
from oracle import clsr_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ckpt():
    z = np.load(os.path.join(GOLD, "ckpt_slice.npz"))
    return {k: z[k] for k in z.files if k != "__names__"}, json.loads(str(z["__names__"]))


def test_parameterisation_matches_shipped_checkpoint(ckpt):
    """Every variable the reference's graph saved exists in our inventory with the same name and
    shape (examples/00_quick_start/CLSR/taobao-clsr-debug/model.tar.gz, epoch_3: 85 tensors)."""
    _, names = ckpt
    spec = {n: list(s) for n, s, _, _ in P.dense_spec(40, 40, 40, [80, 40], [100, 64])}
    spec.update({n: list(s) for n, s in P.table_spec(64005, 2182, 36653, 32, 8, 40)})
    assert len(names) == 85
    assert set(spec) == set(names)
    for n, s in names.items():
        assert spec[n] == s, (n, spec[n], s)
    assert sum(int(np.prod(s)) for s in names.values()) == 6589104


def test_oracle_runs_shipped_weights_and_matches_regression_pin(ckpt):
    prm, _ = ckpt
    z = np.load(os.path.join(GOLD, "oracle_regress.npz"))
    feed = {k[5:]: z[k] for k in z.files if k.startswith("feed/")}
    cfg = O.OracleConfig()
    out = O.predict(prm, feed, cfg, torch.float64)
    np.testing.assert_allclose(out["logit"].numpy().reshape(-1), z["eval_logit"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(out["alpha"].numpy().reshape(-1), z["eval_alpha"], rtol=1e-9, atol=1e-12)
    _, L, dense, _, _ = O.compute_gradients(prm, feed, cfg, torch.float64)
    for k, v in L.items():
        assert abs(float(v) - float(z["loss/" + k])) < 1e-9 * max(1.0, abs(float(v)))
    norms = json.loads(str(z["grad_norms_json"]))
    for k, v in dense.items():
        assert abs(float(v.norm()) - norms[k]) <= 1e-7 * max(norms[k], 1e-12) + 1e-15, k
    # fp32 "reference precision" mode stays within 1e-4 of the fp64 truth on logits
    o32 = O.predict(prm, feed, cfg, torch.float32)
    assert np.abs(o32["logit"].numpy().reshape(-1) - z["eval_logit"]).max() < 1e-4 * np.abs(z["eval_logit"]).max()


def test_oracle_gradients_match_finite_differences():
    src = synth.SyntheticSource(n_items=60, n_cates=8, n_users=12, T=6, seed=2)
    feed = src.batch(5, 4)
    cfg = O.OracleConfig(max_seq_length=6, contrastive_length_threshold=2)
    prm = P.init_params(60, 8, 12, seed=1)
    rng = np.random.default_rng(0)
    prm = {k: (v * (10.0 if v.ndim == 2 and "embedding" not in k else 1.0)).astype(np.float64) for k, v in prm.items()}
    for k in prm:
        if k.endswith("beta") or "b_nn" in k:
            prm[k] = 0.1 * rng.standard_normal(prm[k].shape)
    _, L, dense, slices, _ = O.compute_gradients(prm, feed, cfg, torch.float64)

    def loss_at(p):
        with torch.no_grad():
            pt = {k: torch.as_tensor(v) for k, v in p.items()}
            out = O.forward(pt, feed, cfg, True, torch.float64)
            return float(O.losses(out, pt, feed, cfg, torch.float64)["loss"])

    checks = ["sequential/clsr/short_term/time4lstm/time4lstm_cell/kernel",
              "sequential/clsr/short_term/short_term_intention/gru_cell/gates/kernel",
              "sequential/clsr/long_term/attention_fcn/att_fcn/nn_part/w_nn_layer0",
              "sequential/clsr/short_term/attention_fcn/att_fcn/nn_part/batch_normalization/gamma",
              "sequential/clsr/fcn_alpha/nn_part/w_nn_layer0", "sequential/logit_fcn/nn_part/w_nn_layer1",
              "sequential/clsr/short_term/time4lstm/time4lstm_cell/_time_input_w1"]
    eps = 1e-6
    for name in checks:
        g = dense[name].numpy().reshape(-1)
        for idx in rng.choice(g.size, size=3, replace=False):
            p2 = dict(prm)
            a = prm[name].copy().reshape(-1)
            a[idx] += eps
            p2[name] = a.reshape(prm[name].shape)
            up = loss_at(p2)
            a[idx] -= 2 * eps
            p2[name] = a.reshape(prm[name].shape)
            dn = loss_at(p2)
            fd = (up - dn) / (2 * eps)
            assert abs(fd - g[idx]) <= 1e-5 * max(abs(fd), abs(g[idx])) + 1e-8, (name, idx, fd, g[idx])
    # a table row: dense gradient assembled from the IndexedSlices
    idx, val = slices["item_embedding"]
    dense_item = np.zeros((60, 32))
    np.add.at(dense_item, idx.numpy(), val.numpy())
    row = int(feed["items"][0])
    p2 = dict(prm)
    for col in (0, 7):
        t = prm["sequential/embedding/item_embedding"].copy()
        t[row, col] += eps
        p2["sequential/embedding/item_embedding"] = t
        up = loss_at(p2)
        t = t.copy()
        t[row, col] -= 2 * eps
        p2["sequential/embedding/item_embedding"] = t
        dn = loss_at(p2)
        fd = (up - dn) / (2 * eps)
        assert abs(fd - dense_item[row, col]) <= 1e-5 * max(abs(fd), abs(dense_item[row, col])) + 1e-8


def test_oracle_train_step_semantics():
    """Adam bias correction, clip on concatenated slices, non-lazy sweep (untouched rows still move
    once they have momentum), BN moving statistics."""
    src = synth.SyntheticSource(n_items=80, n_cates=8, n_users=12, T=6, seed=4)
    cfg = O.OracleConfig(max_seq_length=6, contrastive_length_threshold=2)
    prm = P.init_params(80, 8, 12, seed=1)
    slots = {}
    f1, f2 = src.batch(5, 4), src.batch(5, 4)
    before = {k: v.copy() for k, v in prm.items()}
    O.train_step(prm, slots, f1, cfg, 1, torch.float64)
    name = "sequential/embedding/item_embedding"
    touched = np.unique(np.concatenate([f1["item_history"].reshape(-1), f1["items"]]))
    moved = np.abs(prm[name] - before[name]).max(1) > 0
    assert set(np.flatnonzero(moved)) == set(touched.tolist())
    # first Adam step moves every touched coordinate with a nonzero gradient by ~lr
    assert np.abs(prm[name] - before[name]).max() <= 1.001e-3
    mid = prm[name].copy()
    O.train_step(prm, slots, f2, cfg, 2, torch.float64)
    t2 = np.unique(np.concatenate([f2["item_history"].reshape(-1), f2["items"]]))
    only_first = np.setdiff1d(touched, t2)
    assert len(only_first) and (np.abs(prm[name][only_first] - mid[only_first]).max(1) > 0).all()
    mm = "sequential/logit_fcn/nn_part/batch_normalization/moving_variance"
    assert not np.allclose(prm[mm], before[mm])
    assert np.array_equal(prm["sequential/embedding/user_embedding"], before["sequential/embedding/user_embedding"])


def test_graph_variants_select_the_reference_inventories():
    """hparams.interest_evolve / predict_long_short / manual_alpha / sequential_model decide which variables the
    reference graph creates (clsr.py:159-274); the parameter inventory follows, and the oracle runs every variant."""
    import torch
    from clsr_b200 import params as P
    from clsr_b200 import synth
    from oracle import clsr_oracle as O
    base = {n for n, *_ in P.dense_spec(40, 40, 40, [80, 40], [100, 64])}
    sti = {n for n in base if "short_term_intention/" in n}
    c2 = {n for n in base if "causal2/" in n}
    fa = {n for n in base if "fcn_alpha/" in n}
    t4 = {n for n in base if "time4lstm/" in n}
    assert len(sti) == 4 and len(c2) == 4 and len(fa) == 14 and len(t4) == 14
    spec = lambda **kw: {n: shp for n, shp, *_ in P.dense_spec(40, 40, 40, [80, 40], [100, 64], **kw)}
    assert set(spec(interest_evolve=False)) == base - sti
    assert set(spec(manual_alpha=True)) == base - c2 - fa
    nl = spec(predict_long_short=False)
    assert set(nl) == base - c2 and nl["sequential/clsr/fcn_alpha/nn_part/w_nn_layer0"] == (40 + 40 + 40 + 1, 80)
    ls = spec(sequential_model="lstm")
    assert set(ls) == (base - t4) | {"sequential/clsr/short_term/simple_lstm/lstm_cell/kernel",
                                     "sequential/clsr/short_term/simple_lstm/lstm_cell/bias"}
    assert ls["sequential/clsr/short_term/simple_lstm/lstm_cell/kernel"] == (80, 160)
    src = synth.SyntheticSource(n_items=300, n_cates=20, n_users=30, T=12, seed=1)
    feed = src.batch(4, 4)
    for kw in (dict(interest_evolve=False), dict(predict_long_short=False), dict(manual_alpha=True, manual_alpha_value=0.25),
               dict(sequential_model="lstm"), dict(sequential_model="gru")):
        vkw = {k: v for k, v in kw.items() if k != "manual_alpha_value"}
        prm = P.init_params(300, 20, 30, seed=2, **vkw)
        cfg = O.OracleConfig(max_seq_length=12, **kw)
        out, L, dense, slices, _ = O.compute_gradients(prm, feed, cfg, torch.float64)
        assert all(torch.isfinite(g).all() for g in dense.values()) and set(dense) == {k for k in prm if "/embedding/" not in k and "moving_" not in k}
        if kw.get("manual_alpha"):
            assert float(out["alpha"].min()) == float(out["alpha"].max()) == 0.25
