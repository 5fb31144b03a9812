"""Host-side logic against fixtures produced by the reference's own modules
(tests/golden/make_golden.py): iterator feeds, metrics, hparams."""
import json
import os
import random

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
DATA = os.path.join(GOLD, "dataset")

CLI_OVERRIDES = dict(
    embed_l2=1e-6, layer_l2=1e-6, contrastive_loss="triplet", triplet_margin=1.0, discrepancy_loss_weight=0.01,
    contrastive_loss_weight=0.1, learning_rate=0.001, epochs=100, EARLY_STOP=5, batch_size=16, show_step=500,
    MODEL_DIR="m/", SUMMARIES_DIR="s/", need_sample=True, train_num_ngs=4, max_seq_length=50,
    pairwise_metrics=["mean_mrr", "ndcg@2;4;6", "hit@2;4;6"], weighted_metrics=["wauc"], time_unit="s",
    manual_alpha=False, manual_alpha_value=0.5, interest_evolve=True, predict_long_short=True, is_clip_norm=1,
    contrastive_length_threshold=5, contrastive_recent_k=3, sequential_model="time4lstm")


def make_hparams(**kw):
    from reco_utils.recommender.deeprec.deeprec_utils import prepare_hparams
    yaml_file = os.path.join(os.path.dirname(HERE), "reco_utils/recommender/deeprec/config/clsr.yaml")
    o = dict(CLI_OVERRIDES, user_vocab=os.path.join(DATA, "user_vocab.pkl"),
             item_vocab=os.path.join(DATA, "item_vocab.pkl"), cate_vocab=os.path.join(DATA, "category_vocab.pkl"))
    o.update(kw)
    return prepare_hparams(yaml_file, **o)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "golden_host.npz"), allow_pickle=False)


def test_hparams_match_reference(gold):
    want = json.loads(str(gold["hparams_json"]))
    got = make_hparams().values()
    for k in ("user_vocab", "item_vocab", "cate_vocab"):
        got[k] = os.path.basename(got[k])
    assert set(got) == set(want)
    for k, v in want.items():
        assert got[k] == v, (k, got[k], v)


def test_hparams_errors():
    from reco_utils.recommender.deeprec.deeprec_utils import prepare_hparams
    with pytest.raises(ValueError):
        prepare_hparams(None, model_type="clsr", item_embedding_dim=32)
    hp = make_hparams()
    assert "min_seq_length" in hp and "nope" not in hp
    hp.current_epoch = 3
    assert hp.current_epoch == 3


def _named(fd):
    return {getattr(k, "name", k): v for k, v in fd.items()}


def test_eval_feed_matches_reference(gold):
    from reco_utils.recommender.deeprec.io.sequential_iterator import SASequentialIterator
    it = SASequentialIterator(make_hparams(), None)
    feeds = [_named(fd) for fd in it.load_data_from_file(os.path.join(DATA, "valid_data"), batch_num_ngs=0) if fd]
    assert len(feeds) == int(gold["eval_batches"])
    for k in gold.files:
        if not k.startswith("eval0/"):
            continue
        want, got = gold[k], feeds[0][k[6:]]
        assert got.dtype == want.dtype and got.shape == want.shape, (k, got.dtype, want.dtype, got.shape, want.shape)
        if want.dtype.kind == "f":
            np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7, err_msg=k)
        else:
            assert np.array_equal(got, want), k


def test_train_feed_matches_reference(gold):
    """x5 replication, in-batch negatives (same random.randint stream), attn_labels, padding."""
    from reco_utils.recommender.deeprec.io.sequential_iterator import SASequentialIterator
    it = SASequentialIterator(make_hparams(), None)
    lines = it.parse_file(os.path.join(DATA, "train_data"))[:12]
    cols = [list(c) for c in zip(*lines)]
    random.seed(1234)
    tr = it._convert_data(*cols, 4)
    for k in gold.files:
        if not k.startswith("train/"):
            continue
        want, got = gold[k], tr[k[6:]]
        assert got.dtype == want.dtype and got.shape == want.shape, (k, got.dtype, want.dtype)
        if want.dtype.kind == "f":
            np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7, err_msg=k)
        else:
            assert np.array_equal(got, want), k
    assert it._convert_data(*[c[:4] for c in cols], 4) is None   # < 5 instances: batch dropped


def test_metrics_match_reference(gold):
    from reco_utils.recommender.deeprec.deeprec_utils import cal_metric, cal_weighted_metric
    want = json.loads(str(gold["metrics_json"]))
    y, p, users = gold["metric_y"], gold["metric_p"], gold["metric_users"]
    got = {}
    got.update(cal_metric(list(y), list(p), ["auc", "logloss"]))
    got.update(cal_metric(list(y.reshape(-1, 10)), list(p.reshape(-1, 10)),
                          ["mean_mrr", "ndcg@2;4;6", "hit@2;4;6", "group_auc"]))
    got.update(cal_weighted_metric(list(users), list(p), list(y), ["wauc"]))
    assert set(got) == set(want)
    for k, v in want.items():
        assert abs(float(got[k]) - v) < 1e-9, (k, got[k], v)


def test_detect_group_and_normalize():
    from clsr_b200 import synth
    from clsr_b200.engine import detect_group, normalize_feed
    src = synth.SyntheticSource(n_items=500, n_cates=10, n_users=50, T=20, seed=0)
    f = normalize_feed(src.batch(8, 4))
    assert detect_group(f, 5) == 5
    f["item_history"][3, 2] += 1
    assert detect_group(f, 5) == 1
    g = src.batch(6, 0)
    g["users"] = g["users"].astype(np.float32)
    n = normalize_feed(g, need_labels=False)
    assert n["users"].dtype == np.int32 and n["labels"] is None and n["mask"].dtype == np.int32
