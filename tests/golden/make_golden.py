"""Generate the golden fixtures under tests/golden/ by importing the reference's own Python modules
from /root/reference (possible only in the build container; the fixtures travel, the reference does
not).  TensorFlow is not installable here, so a recording stub stands in for it: the modules used
below touch TF only to declare placeholders / build an HParams object, never for arithmetic.

  golden_host.npz   : SASequentialIterator feeds (eval batch from a file; training batch from
                      _convert_data under random.seed(1234)); cal_metric / cal_weighted_metric /
                      cal_mean_alpha_metric outputs; prepare_hparams(clsr.yaml + CLI overrides).
  ckpt_slice.npz    : every dense variable of the shipped checkpoint
                      examples/00_quick_start/CLSR/taobao-clsr-debug/model.tar.gz (epoch_3) plus the
                      first rows of its five embedding tables.
  oracle_regress.npz: outputs of oracle/clsr_oracle.py itself on a seeded feed with ckpt_slice
                      weights (regression pin of the oracle; NOT a reference output).
Run:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys
import tarfile
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def stub_tensorflow():
    tf = types.ModuleType("tensorflow")
    tf.__version__ = "stub"
    tf.float32, tf.int32, tf.bool = "float32", "int32", "bool"

    class _PH:
        def __init__(self, dtype, shape=None, name=None):
            self.name = name
    tf.placeholder = lambda dtype, shape=None, name=None: _PH(dtype, shape, name)

    class HParams:
        def __init__(self, **kw):
            self.kw = kw
    contrib = types.ModuleType("tensorflow.contrib")
    training = types.ModuleType("tensorflow.contrib.training")
    training.HParams = HParams
    contrib.training = training
    tf.contrib = contrib

    class _G:
        def as_default(self):
            return self
        def __enter__(self):
            return self
        def __exit__(self, *a):
            return False
    tf.Graph = _G
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.contrib"] = contrib
    sys.modules["tensorflow.contrib.training"] = training
    return tf


def main():
    tf = stub_tensorflow()
    from clsr_b200.synth_dataset import write_dataset
    data = os.path.join(HERE, "dataset")
    write_dataset(data, n_users=60, n_items=150, n_cates=12, train_lines=60, valid_users=6, test_users=4, seed=5)
    sys.path.insert(0, REF)
    for m in [k for k in sys.modules if k.startswith("reco_utils")]:
        del sys.modules[m]
    from reco_utils.recommender.deeprec import deeprec_utils as RU
    from reco_utils.recommender.deeprec.io.sequential_iterator import SASequentialIterator as RefIt
    assert RU.__file__.startswith(REF)
    out = {}
    overrides = dict(embed_l2=1e-6, layer_l2=1e-6, contrastive_loss="triplet", triplet_margin=1.0,
                     discrepancy_loss_weight=0.01, contrastive_loss_weight=0.1, learning_rate=0.001, epochs=100,
                     EARLY_STOP=5, batch_size=16, show_step=500, MODEL_DIR="m/", SUMMARIES_DIR="s/",
                     user_vocab=os.path.join(data, "user_vocab.pkl"), item_vocab=os.path.join(data, "item_vocab.pkl"),
                     cate_vocab=os.path.join(data, "category_vocab.pkl"), need_sample=True, train_num_ngs=4,
                     max_seq_length=50, pairwise_metrics=["mean_mrr", "ndcg@2;4;6", "hit@2;4;6"],
                     weighted_metrics=["wauc"], time_unit="s", manual_alpha=False, manual_alpha_value=0.5,
                     interest_evolve=True, predict_long_short=True, is_clip_norm=1,
                     contrastive_length_threshold=5, contrastive_recent_k=3, sequential_model="time4lstm")
    hp = RU.prepare_hparams(os.path.join(REF, "reco_utils/recommender/deeprec/config/clsr.yaml"), **overrides)
    hpv = dict(hp.kw)
    for k in ("user_vocab", "item_vocab", "cate_vocab"):
        hpv[k] = os.path.basename(hpv[k])
    out["hparams_json"] = np.array(json.dumps(hpv, sort_keys=True))

    class HP:
        pass
    h = HP()
    for k, v in hp.kw.items():
        setattr(h, k, v)
    it = RefIt(h, tf.Graph())
    name = lambda fd: {getattr(k, "name", str(k)): v for k, v in fd.items()}
    feeds = [name(fd) for fd in it.load_data_from_file(os.path.join(data, "valid_data"), batch_num_ngs=0) if fd]
    for k, v in feeds[0].items():
        out["eval0/" + k] = np.asarray(v)
    out["eval_batches"] = np.array(len(feeds))
    lines = it.parse_file(os.path.join(data, "train_data"))[:12]
    cols = list(zip(*lines))
    args = [list(cols[0]), list(cols[1]), list(cols[2]), list(cols[3]), list(cols[4]), list(cols[5]), list(cols[6]),
            list(cols[7]), list(cols[8]), list(cols[9])]
    random.seed(1234)
    tr = it._convert_data(*args, 4)
    for k, v in tr.items():
        out["train/" + k] = np.asarray(v)
    # metrics
    rng = np.random.default_rng(3)
    y = (rng.random(400) < 0.3).astype(np.float64)
    y[::10] = 1
    p = np.round(rng.random(400), 3)
    users = np.repeat(np.arange(40), 10)
    gl, gp = y.reshape(-1, 10), p.reshape(-1, 10)
    m1 = RU.cal_metric(list(y), list(p), ["auc", "logloss"])
    m2 = RU.cal_metric(list(gl), list(gp), ["mean_mrr", "ndcg@2;4;6", "hit@2;4;6", "group_auc"])
    m3 = RU.cal_weighted_metric(list(users), list(p), list(y), ["wauc"])
    m4 = RU.cal_mean_alpha_metric(list(rng.random(400)), list(y))
    out["metric_y"], out["metric_p"], out["metric_users"] = y, p, users
    out["metric_alpha_in"] = np.random.default_rng(3).random(0)
    met = {}
    for m in (m1, m2, m3):
        met.update({k: float(v) for k, v in m.items()})
    out["metrics_json"] = np.array(json.dumps(met, sort_keys=True))
    np.savez_compressed(os.path.join(HERE, "golden_host.npz"), **out)

    # shipped checkpoint slice
    for m in [k for k in sys.modules if k.startswith("reco_utils")]:
        del sys.modules[m]
    sys.path.remove(REF)
    from clsr_b200 import tf_bundle as tb
    tmp = tempfile.mkdtemp()
    with tarfile.open(os.path.join(REF, "examples/00_quick_start/CLSR/taobao-clsr-debug/model.tar.gz")) as t:
        t.extractall(tmp)
    ck = tb.read_bundle(os.path.join(tmp, "model", "epoch_3"))
    keep = {"item_embedding": 3000, "cate_embedding": 40, "user_embedding": 200, "user_long_embedding": 200,
            "user_short_embedding": 200}
    sl = {}
    for k, v in ck.items():
        base = k.split("/")[-1]
        sl[k] = v[:keep[base]].copy() if base in keep else v
    sl["__names__"] = np.array(json.dumps({k: list(v.shape) for k, v in ck.items()}, sort_keys=True))
    np.savez_compressed(os.path.join(HERE, "ckpt_slice.npz"), **sl)

    # oracle regression pin
    import torch
    from clsr_b200 import synth
    from oracle import clsr_oracle as O
    src = synth.SyntheticSource(n_items=3000, n_cates=40, n_users=200, T=50, seed=17)
    feed = src.batch(12, 4)
    prm = {k: v for k, v in sl.items() if k != "__names__"}
    cfg = O.OracleConfig()
    o_eval = O.predict(prm, feed, cfg, torch.float64)
    _, L, dense, slices, _ = O.compute_gradients(prm, feed, cfg, torch.float64)
    reg = {"feed/" + k: v for k, v in feed.items()}
    reg["eval_logit"] = o_eval["logit"].numpy().reshape(-1)
    reg["eval_alpha"] = o_eval["alpha"].numpy().reshape(-1)
    for k, v in L.items():
        reg["loss/" + k] = np.array(float(v))
    reg["grad_norms_json"] = np.array(json.dumps({k: float(v.norm()) for k, v in dense.items()}, sort_keys=True))
    np.savez_compressed(os.path.join(HERE, "oracle_regress.npz"), **reg)
    print("wrote fixtures:", [f for f in os.listdir(HERE) if f.endswith(".npz")])


if __name__ == "__main__":
    main()
