"""Parity at the BASELINE.json sizes (the small-problem tests in test_gpu_parity.py cannot reach the
persistent-CTA tile scheduling, the grouped weight-gradient partitioning, the 4 M-row slot tables or
64-bit table offsets): the product path (math_mode=1) through the C ABI against the CPU oracle on
identical feeds.

  * config 2 (Taobao: S=4096 sequences x 5 rows, T=50, 4.0 M items / 9.4 k cates / 1.0 M users) against the
    fp32 "reference precision" oracle (the fp64 one needs ~40 GB at this size);
  * config 3 shape (Kuaishou: T=250) at S=512 against the fp64 oracle;
  * AUC on >= 5000 eval rows: |AUC(engine) - AUC(oracle)| < 1e-3 (north star), logits within 1e-3;
  * history gather / scatter-add bit-exactness on a table with more than 2^31 elements.
Tolerances are stated next to each assert; gathered rows and ids are bit-exact.
"""
import numpy as np
import pytest

import parity_util as PU

pytestmark = pytest.mark.gpu

NI, NC, NU = 4_000_000, 9_400, 1_000_000


def _big_problem(S, T, seed, G=5):
    from clsr_b200 import params as P, synth
    src = synth.SyntheticSource(NI, NC, NU, T, seed=seed)
    feed = src.batch(S, G - 1)
    prm = PU.scale_params(P.init_params(NI, NC, NU, seed=seed), seed)
    return feed, prm


def _step_vs_oracle(S, T, seed, dtype, tol_logit, tol_loss, tol_dense, tol_table):
    import torch
    from oracle import clsr_oracle as O
    from clsr_b200.engine import STEP_NO_OPTIMIZER, STEP_NO_BN_UPDATE, TABLE_VARS
    G = 5
    feed, prm = _big_problem(S, T, seed)
    B = S * G
    eng = PU.make_engine(prm, NI, NC, NU, max_rows=B, T=T, G=G, math_mode=1)
    got = eng.train_step(feed, group=G, flags=STEP_NO_OPTIMIZER | STEP_NO_BN_UPDATE)
    cfg = PU.oracle_config(G, max_seq_length=T)
    # the engine's ReLU decisions are fed into the oracle (see test_step_matches_oracle_tensor_core_path): a unit
    # within rounding distance of the kink may legitimately fall on either side; the flips are counted and printed
    with O.relu_decisions(PU.engine_relu_masks(eng, B, S, T, G)) as rd:
        out, L, dense, slices, _ = O.compute_gradients(prm, feed, cfg, dtype)
    rep = {}
    total_units = B * T * 120 + S * T * 120 + B * (120 + 164)
    nflip = sum(rd.flips.values())
    print("relu decisions differing from the oracle: %d of %d units (%s)" % (nflip, total_units, rd.flips))
    assert nflip < 1e-4 * total_units
    # gathered history rows: bit-exact (compare a strided sample of sequences to keep the copy small)
    X = eng.debug("X", (S, T, 40))
    ref_X = out["inter"]["hist"].detach().numpy().reshape(S, G, T, 40)[:, 0]
    assert np.array_equal(X, ref_X.astype(np.float32)), "gathered history rows must be bit-exact"
    logit = eng.debug("logit", (B,))
    rep["logit"] = PU.relerr(logit, out["logit"].detach().numpy().reshape(-1))
    assert rep["logit"] < tol_logit, rep
    for k in ("loss", "data_loss", "regular_loss", "contrastive_loss", "discrepancy_loss"):
        rep[k] = abs(got[k] - float(L[k])) / max(abs(float(L[k])), 1e-12)
        assert rep[k] < tol_loss, (k, got[k], float(L[k]))
    dg = eng.get_dense(3)
    for name, gref in dense.items():
        if name.endswith("b_nn_output") or "b_nn_layer" in name:
            continue   # softmax shift invariance / a bias in front of BatchNorm: the true gradient is 0, what is left is rounding
        rep["grad/" + name.split("sequential/")[-1]] = PU.relerr_l2(dg[name], gref.numpy().reshape(-1))
    worst = max((v, k) for k, v in rep.items() if k.startswith("grad/"))
    assert worst[0] < tol_dense, worst
    for t, name in TABLE_VARS.items():
        tab = name.split("/")[-1]
        ids, rows = eng.sparse_grad(t)
        idx, val = slices[tab]
        uniq, inv = torch.unique(idx, return_inverse=True)
        assert len(ids) == len(uniq), (tab, len(ids), len(uniq))
        want = torch.zeros(len(uniq), val.shape[1], dtype=torch.float64).index_add_(0, inv, val.double()).numpy()
        order = np.argsort(ids)
        assert np.array_equal(ids[order], uniq.numpy()), tab
        rep["table/" + tab] = PU.relerr_l2(rows[order], want)
        assert rep["table/" + tab] < tol_table, (tab, rep["table/" + tab])
        # size-independent property: column sums of the compact rows = column sums of all slice values
        assert np.allclose(rows.sum(0, dtype=np.float64), val.double().sum(0).numpy(), rtol=5e-3, atol=1e-5), tab
    print("scale parity S=%d T=%d: logit %.2e, worst dense grad %.2e (%s), tables %s"
          % (S, T, rep["logit"], worst[0], worst[1], {k: "%.1e" % v for k, v in rep.items() if k.startswith("table/")}))
    eng.close()


def test_taobao_config_step_matches_fp32_oracle(cuda_lib):
    """BASELINE config 2: the batch bench.py times.  fp32 oracle => its own rounding is part of the gap."""
    import torch
    _step_vs_oracle(S=4096, T=50, seed=101, dtype=torch.float32, tol_logit=1e-3, tol_loss=2e-4,
                    tol_dense=1e-3, tol_table=2e-4)


def test_kuaishou_window_step_matches_fp64_oracle(cuda_lib):
    """BASELINE config 3 shape (T=250) at 512 sequences, fp64 oracle."""
    import torch
    _step_vs_oracle(S=512, T=250, seed=102, dtype=torch.float64, tol_logit=1e-3, tol_loss=2e-4,
                    tol_dense=1e-3, tol_table=2e-4)


def test_auc_matches_oracle(cuda_lib):
    """North star: logits AND AUC within 1e-3 of the reference CPU path on identical batches.
    6000 eval rows (eval feed: one row per line, float users / mask as the reference's eval iterator builds)."""
    import torch
    from sklearn.metrics import roc_auc_score
    from clsr_b200 import params as P, synth
    from oracle import clsr_oracle as O
    n_items, n_cates, n_users, rows, chunk = 50_000, 500, 5_000, 6000, 1500
    src = synth.SyntheticSource(n_items, n_cates, n_users, 50, seed=77)
    prm = PU.scale_params(P.init_params(n_items, n_cates, n_users, seed=77), 77)
    eng = PU.make_engine(prm, n_items, n_cates, n_users, max_rows=chunk, G=5, training=False, math_mode=1)
    preds, refs, logits, rlogits, labels = [], [], [], [], []
    for _ in range(rows // chunk):
        feed = src.batch(chunk, 0)
        feed["users"] = feed["users"].astype(np.float32)
        feed["mask"] = feed["mask"].astype(np.float32)
        p, _ = eng.predict(feed, group=1)
        ref = O.predict(prm, feed, PU.oracle_config(5), torch.float64)
        preds.append(p)
        refs.append(ref["pred"].numpy().reshape(-1))
        logits.append(eng.debug("logit", (chunk,)).copy())
        rlogits.append(ref["logit"].numpy().reshape(-1))
        labels.append(feed["labels"].reshape(-1))
    preds, refs, labels = np.concatenate(preds), np.concatenate(refs), np.concatenate(labels)
    # labels correlated with the model so the AUC is away from 0.5 and sensitive to rank changes
    rng = np.random.default_rng(0)
    labels = (rng.random(rows) < 1.0 / (1.0 + np.exp(-8.0 * (refs - np.median(refs)) / (refs.std() + 1e-12)))).astype(np.float32)
    assert 0 < labels.sum() < rows
    auc_e, auc_o = roc_auc_score(labels, preds), roc_auc_score(labels, refs)
    assert abs(auc_e - auc_o) < 1e-3 * auc_o, (auc_e, auc_o)
    assert PU.relerr(np.concatenate(logits), np.concatenate(rlogits)) < 1e-3
    # rank agreement: almost every pair ordered identically (Kendall-type check through the rank vectors)
    assert np.corrcoef(np.argsort(np.argsort(preds)), np.argsort(np.argsort(refs)))[0, 1] > 0.9999


def test_gather_scatter_on_table_beyond_2_31_elements(cuda_lib):
    """70 M x 32 fp32 item rows = 2.24 G elements (8.96 GB): ids near the end address beyond 2^31 elements /
    2^33 bytes.  Bit-exact gather against torch indexing; scatter-add of integer-valued gradients exact."""
    import torch
    from clsr_b200.engine import Engine, TABLE_ITEM, TABLE_CATE
    n_items, n_cates, T, rows = 70_000_000, 1000, 50, 2048
    assert n_items * 32 > 2 ** 31
    eng = Engine(n_items, n_cates, 1000, max_rows=rows, seq_len=T, training=False)
    g = torch.Generator(device="cuda").manual_seed(1)
    tab = eng.tables[TABLE_ITEM]
    # cheap distinct content: row r = r mod 2^20 + column index / 64
    for a in range(0, n_items, 10_000_000):
        b = min(n_items, a + 10_000_000)
        r = torch.arange(a, b, device="cuda", dtype=torch.int64)
        tab[a:b] = ((r % (1 << 20)).float()[:, None] + torch.arange(32, device="cuda").float()[None, :] / 64.0)
    eng.tables[TABLE_CATE].normal_(generator=g)
    ih = torch.randint(n_items - 3_000_000, n_items, (rows, T), device="cuda", dtype=torch.int32, generator=g)
    ih[:, 45:] = 0
    ih[0, :8] = torch.tensor([n_items - 1, n_items - 2, 2 ** 26, 2 ** 26 + 1, 67_108_863, 67_108_864, 1, 0],
                             dtype=torch.int32, device="cuda")   # 2^26 rows x 32 = 2^31 elements
    ch = torch.randint(0, n_cates, (rows, T), device="cuda", dtype=torch.int32, generator=g)
    out = torch.empty(rows, T, 40, device="cuda")
    eng._check(eng.lib.clsr_gather_history(eng.h, ih.data_ptr(), ch.data_ptr(), rows * T, out.data_ptr()))
    eng.synchronize()
    ref = torch.cat([tab[ih.long()], eng.tables[TABLE_CATE][ch.long()]], -1)
    assert torch.equal(out, ref)
    d = torch.randint(-3, 4, (rows, T, 40), device="cuda", generator=g).float()
    eng._check(eng.lib.clsr_scatter_history_grad(eng.h, ih.data_ptr(), ch.data_ptr(), rows * T, d.data_ptr()))
    ids, rws = eng.sparse_grad(TABLE_ITEM)
    uniq, inv = torch.unique(ih.reshape(-1).long(), return_inverse=True)
    want = torch.zeros(len(uniq), 32, device="cuda").index_add_(0, inv, d[..., :32].reshape(-1, 32)).cpu().numpy()
    order = np.argsort(ids)
    assert np.array_equal(ids[order].astype(np.int64), uniq.cpu().numpy())
    assert np.array_equal(rws[order], want)      # small integers: fp32 sums are exact in any order
    eng.close()


def test_out_of_range_ids_are_refused(cuda_lib):
    """A feed whose ids do not fit the tables is an error, not an out-of-bounds access (host feed: refused before
    the copy; device-resident feed: read as id 0 and reported at the next synchronisation)."""
    from clsr_b200.engine import EngineError
    G, S = 5, 8
    feed, prm = PU.small_problem(S=S, G=G, n_items=300, n_cates=20, n_users=50)
    eng = PU.make_engine(prm, 300, 20, 50, max_rows=S * G, G=G)
    ok = eng.train_step(feed, group=G)
    assert np.isfinite(ok["loss"])
    for key, bad in (("item_history", 300), ("items", -1), ("item_cate_history", 20), ("cates", 1 << 30), ("users", 50)):
        f = {k: v.copy() for k, v in feed.items()}
        f[key].reshape(-1)[3 * G if key == "users" else 7] = bad
        if key in ("users", "item_history", "item_cate_history"):   # keep the group replication intact
            f[key] = np.repeat(f[key][::G], G, axis=0)
            f[key].reshape(f[key].shape[0], -1)[0] = bad
        with pytest.raises(EngineError, match="outside"):
            eng.train_step(f, group=G)
        with pytest.raises(EngineError, match="outside"):
            from clsr_b200.engine import normalize_feed
            eng.train_step(eng.to_device(normalize_feed(f)), group=G, on_device=True)
    again = eng.train_step(feed, group=G)         # the engine is still usable
    assert np.isfinite(again["loss"])
