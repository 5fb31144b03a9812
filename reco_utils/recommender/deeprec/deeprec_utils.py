"""Configuration and metric helpers of the deeprec models, TensorFlow-free.

Mirrors the public surface of the reference's reco_utils/recommender/deeprec/deeprec_utils.py
that the sequential models use: ``prepare_hparams`` (:514-534) with YAML flattening (:25-39),
required-key checks per model_type (:138-305) and the default table of ``create_hparams``
(:327-511); ``cal_metric`` (:621-699), ``cal_weighted_metric`` (:702-743),
``cal_mean_alpha_metric`` (:812-821), ``load_dict`` (:824-835).  Metrics run on the host, as in
the reference; they are computed with numpy rank statistics instead of sklearn/pandas loops and
return the same keys with the same 4-decimal rounding.
"""
import pickle as pkl

import numpy as np
import yaml


class HParams:
    """Stand-in for tf.contrib.training.HParams: attribute read/write, ``in``, ``values()``."""

    def __init__(self, **kwargs):
        self.__dict__["_hp"] = dict(kwargs)

    def __getattr__(self, name):
        try:
            return self.__dict__["_hp"][name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self.__dict__["_hp"][name] = value

    def __contains__(self, name):
        return name in self.__dict__["_hp"]

    def values(self):
        return dict(self.__dict__["_hp"])

    def __repr__(self):
        return "HParams(%s)" % ", ".join("%s=%r" % kv for kv in sorted(self.__dict__["_hp"].items()))


def flat_config(config):
    """Flatten the two-level YAML (sections data / model / train / info) into one dict."""
    out = {}
    for section in config.values():
        out.update(section)
    return out


def load_yaml(filename):
    try:
        with open(filename, "r") as f:
            return yaml.load(f, yaml.SafeLoader)
    except FileNotFoundError:
        raise
    except Exception:
        raise IOError("load {0} error!".format(filename))


_SEQ_COMMON = ["item_embedding_dim", "cate_embedding_dim", "max_seq_length", "loss", "method",
               "user_vocab", "item_vocab", "cate_vocab"]
_REQUIRED = {
    "clsr": _SEQ_COMMON + ["attention_size", "hidden_size", "att_fcn_layer_sizes", "discrepancy_loss_weight",
                           "contrastive_loss_weight", "is_clip_norm", "contrastive_length_threshold"],
    "sli_rec": _SEQ_COMMON + ["attention_size", "hidden_size", "att_fcn_layer_sizes"],
    "gru4rec": _SEQ_COMMON + ["hidden_size"],
    "asvd": list(_SEQ_COMMON),
    "caser": _SEQ_COMMON + ["user_embedding_dim", "T", "L", "n_v", "n_h", "min_seq_length"],
    "nextitnet": _SEQ_COMMON + ["user_embedding_dim", "dilations", "kernel_size", "min_seq_length"],
}
_ALIASES = {"CLSR": "clsr", "slirec": "sli_rec", "SLI_REC": "sli_rec", "Sli_rec": "sli_rec", "GRU4REC": "gru4rec",
            "GRU4Rec": "gru4rec", "ASVD": "asvd", "a2svd": "asvd", "A2SVD": "asvd", "CASER": "caser",
            "Caser": "caser", "next_it_net": "nextitnet", "NextItNet": "nextitnet", "NEXT_IT_NET": "nextitnet"}


def check_nn_config(f_config):
    """Required keys per model_type and the activation/layer-count consistency check."""
    mt = f_config.get("model_type")
    if mt is None:
        raise KeyError("model_type")
    for p in _REQUIRED.get(_ALIASES.get(mt, mt), []):
        if p not in f_config:
            raise ValueError("Parameters {0} must be set".format(p))
    if mt in ("exDeepFM", "xDeepFM") and f_config.get("data_format") != "ffm":
        raise ValueError("For xDeepFM model, data format must be 'ffm', but your set is {0}".format(
            f_config.get("data_format")))
    if mt in ("dkn", "DKN") and f_config.get("data_format") != "dkn":
        raise ValueError("For dkn model, data format must be 'dkn', but your set is {0}".format(
            f_config.get("data_format")))
    check_type(f_config)


_INT = ["epochs", "batch_size", "show_step", "save_epoch", "item_embedding_dim", "cate_embedding_dim",
        "user_embedding_dim", "max_seq_length", "hidden_size", "T", "L", "n_v", "n_h", "kernel_size",
        "min_seq_length", "attention_size", "train_num_ngs"]
_FLOAT = ["init_value", "learning_rate", "embed_l2", "embed_l1", "layer_l2", "layer_l1", "mu"]
_STR = ["method", "loss", "optimizer", "init_method", "user_vocab", "item_vocab", "cate_vocab"]
_LIST = ["layer_sizes", "activation", "dropout", "att_fcn_layer_sizes", "dilations"]


def check_type(config):
    for names, typ, label in ((_INT, int, "int"), (_FLOAT, float, "float"), (_STR, str, "str"),
                              (_LIST, list, "list")):
        for p in names:
            if p in config and not isinstance(config[p], typ):
                raise TypeError("Parameters {0} must be {1}".format(p, label))


# name -> default (create_hparams, deeprec_utils.py:327-511); names not listed default to None.
_DEFAULTS = dict(
    use_entity=True, use_context=True, cross_activation="identity", user_dropout=False, dropout=[0.0],
    attention_dropout=0.0, load_saved_model=False, fast_CIN_d=0, use_Linear_part=False, use_FM_part=False,
    use_CIN_part=False, use_DNN_part=False, init_method="tnormal", init_value=0.01, embed_l2=0.0, embed_l1=0.0,
    layer_l2=0.0, layer_l1=0.0, cross_l2=0.0, cross_l1=0.0, attn_loss_weight=0.0, contrastive_loss="bpr",
    triplet_margin=1.0, discrepancy_loss_weight=0.0, contrastive_loss_weight=0.0, contrastive_length_threshold=1,
    contrastive_recent_k=3, reg_kg=0.0, learning_rate=0.001, lr_rs=1, lr_kg=0.5, kg_training_interval=5,
    max_grad_norm=2, is_clip_norm=0, vector_alpha=False, manual_alpha=False, manual_alpha_value=0.5,
    interest_evolve=True, predict_long_short=True, dtype=32, optimizer="adam", epochs=10, batch_size=1,
    enable_BN=False, show_step=1, save_model=True, save_epoch=5, write_tfevents=False, train_num_ngs=4,
    need_sample=True, embedding_dropout=0.3, EARLY_STOP=100, min_seq_length=1, counterfactual_recent_k=5,
    use_complex_attention=False, sequential_model="time4lstm", time_unit="s", ncf_layer_sizes=[80, 40],
)
_NONE_DEFAULT = [
    "kg_file", "user_clicks", "FEATURE_COUNT", "FIELD_COUNT", "data_format", "PAIR_NUM", "DNN_FIELD_NUM", "n_user",
    "n_item", "n_user_attr", "n_item_attr", "iterator_type", "SUMMARIES_DIR", "MODEL_DIR", "wordEmb_file",
    "entityEmb_file", "contextEmb_file", "news_feature_file", "user_history_file", "doc_size", "history_size",
    "word_size", "entity_size", "entity_dim", "entity_embedding_method", "transform", "train_ratio", "dim",
    "layer_sizes", "cross_layer_sizes", "cross_layers", "activation", "attention_layer_sizes",
    "attention_activation", "model_type", "method", "load_model_name", "filter_sizes", "num_filters", "mu", "loss",
    "metrics", "item_embedding_dim", "cate_embedding_dim", "user_embedding_dim", "user_vocab", "item_vocab",
    "cate_vocab", "pairwise_metrics", "weighted_metrics", "max_seq_length", "hidden_size", "L", "T", "n_v", "n_h",
    "attention_size", "att_fcn_layer_sizes", "dilations", "kernel_size", "embed_size", "n_layers", "decay",
    "eval_epoch", "top_k",
]


def create_hparams(flags):
    hp = {k: None for k in _NONE_DEFAULT}
    hp.update(_DEFAULTS)
    for k in hp:
        if k in flags:
            hp[k] = flags[k]
    return HParams(**hp)


def prepare_hparams(yaml_file=None, **kwargs):
    """YAML + keyword overrides -> HParams (unknown keys are dropped, as create_hparams does)."""
    config = flat_config(load_yaml(yaml_file)) if yaml_file is not None else {}
    config.update(kwargs)
    check_nn_config(config)
    return create_hparams(config)


def load_dict(filename):
    with open(filename, "rb") as f:
        return pkl.load(f)


# ---- metrics ---------------------------------------------------------------------------------------

def _average_ranks(x):
    """1-based ranks with ties sharing their average rank."""
    order = np.argsort(x, kind="mergesort")
    xs = x[order]
    boundary = np.concatenate(([True], xs[1:] != xs[:-1]))
    start = np.flatnonzero(boundary)
    end = np.concatenate((start[1:], [len(x)]))
    avg = (start + end + 1) / 2.0
    ranks = np.empty(len(x), np.float64)
    ranks[order] = np.repeat(avg, end - start)
    return ranks


def roc_auc(y_true, y_score):
    """Area under the ROC curve = normalised Mann-Whitney U (ties count one half), which is what
    sklearn.metrics.roc_auc_score returns for binary labels."""
    y = np.asarray(y_true, np.float64).reshape(-1)
    s = np.asarray(y_score, np.float64).reshape(-1)
    pos = y == 1
    n_pos, n_neg = int(pos.sum()), int((~pos).sum())
    if n_pos == 0 or n_neg == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    r = _average_ranks(s)
    return (r[pos].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg)


def mrr_score(y_true, y_score):
    order = np.argsort(y_score)[::-1]
    y = np.take(y_true, order)
    return np.sum(y / (np.arange(len(y)) + 1)) / np.sum(y)


def dcg_score(y_true, y_score, k=10):
    k = min(np.shape(y_true)[-1], k)
    order = np.argsort(y_score)[::-1]
    y = np.take(y_true, order[:k])
    return np.sum((2 ** y - 1) / np.log2(np.arange(len(y)) + 2))


def ndcg_score(y_true, y_score, k=10):
    return dcg_score(y_true, y_score, k) / dcg_score(y_true, y_true, k)


def hit_score(y_true, y_score, k=10):
    truth = set(np.where(np.asarray(y_true) == 1)[0].tolist())
    for idx in np.argsort(y_score)[::-1][:k]:
        if int(idx) in truth:
            return 1
    return 0


def _ks(metric, default=(1, 2)):
    parts = metric.split("@")
    return [int(t) for t in parts[1].split(";")] if len(parts) > 1 else list(default)


def cal_metric(labels, preds, metrics):
    """Pointwise metrics on flat lists, pairwise metrics on lists of per-impression groups."""
    res = {}
    if not metrics:
        return res
    for metric in metrics:
        if metric == "auc":
            res["auc"] = round(roc_auc(labels, preds), 4)
        elif metric == "rmse":
            mse = float(np.mean((np.asarray(labels, np.float64) - np.asarray(preds, np.float64)) ** 2))
            res["rmse"] = np.sqrt(round(mse, 4))
        elif metric == "logloss":
            p = np.clip(np.asarray(preds, np.float64).reshape(-1), 10e-12, 1.0 - 10e-12)
            y = np.asarray(labels, np.float64).reshape(-1)
            res["logloss"] = round(float(-np.mean(y * np.log(p) + (1 - y) * np.log(1 - p))), 4)
        elif metric == "acc":
            hard = (np.asarray(preds) >= 0.5).astype(np.float64)
            res["acc"] = round(float(np.mean(hard == np.asarray(labels, np.float64))), 4)
        elif metric == "f1":
            hard = np.asarray(preds) >= 0.5
            y = np.asarray(labels) == 1
            tp = float(np.sum(hard & y))
            den = 2 * tp + float(np.sum(hard & ~y)) + float(np.sum(~hard & y))
            res["f1"] = round(2 * tp / den if den else 0.0, 4)
        elif metric == "mean_mrr":
            res["mean_mrr"] = round(float(np.mean([mrr_score(l, p) for l, p in zip(labels, preds)])), 4)
        elif metric.startswith("ndcg"):
            for k in _ks(metric):
                res["ndcg@{0}".format(k)] = round(
                    float(np.mean([ndcg_score(l, p, k) for l, p in zip(labels, preds)])), 4)
        elif metric.startswith("hit"):
            for k in _ks(metric):
                res["hit@{0}".format(k)] = round(
                    float(np.mean([hit_score(l, p, k) for l, p in zip(labels, preds)])), 4)
        elif metric == "group_auc":
            res["group_auc"] = round(float(np.mean([roc_auc(l, p) for l, p in zip(labels, preds)])), 4)
        else:
            raise ValueError("not define this metric {0}".format(metric))
    return res


def _user_groups(users):
    users = np.asarray(users).reshape(-1)
    order = np.argsort(users, kind="mergesort")
    su = users[order]
    start = np.flatnonzero(np.concatenate(([True], su[1:] != su[:-1])))
    end = np.concatenate((start[1:], [len(su)]))
    return order, start, end


def cal_weighted_metric(users, preds, labels, metrics):
    """Per-user metrics weighted by each user's share of the rows ('wauc' is the GAUC the README
    reports; a user with a single label class raises, exactly as sklearn does in the reference)."""
    res = {}
    if not metrics:
        return res
    preds = np.asarray(preds, np.float64).reshape(-1)
    labels = np.asarray(labels, np.float64).reshape(-1)
    order, start, end = _user_groups(users)
    weight = (end - start) / float(len(preds))

    def per_user(fn):
        return np.array([fn(labels[order[a:b]], preds[order[a:b]]) for a, b in zip(start, end)])

    for metric in metrics:
        if metric == "wauc":
            res["wauc"] = round(float((weight * per_user(roc_auc)).sum()), 4)
        elif metric == "wmrr":
            res["wmrr"] = round(float((weight * per_user(mrr_score)).sum()), 4)
        elif metric.startswith("whit"):
            for k in _ks(metric):
                res["whit@{0}".format(k)] = round(
                    float((weight * per_user(lambda y, p: hit_score(y, p, k))).sum()), 4)
        elif metric.startswith("wndcg"):
            for k in _ks(metric):
                res["wndcg@{0}".format(k)] = round(
                    float((weight * per_user(lambda y, p: ndcg_score(y, p, k))).sum()), 4)
        else:
            raise ValueError("not define this metric {0}".format(metric))
    return res


def cal_mean_alpha_metric(alphas, labels):
    alphas = np.asarray(alphas)
    labels = np.asarray(labels)
    return {"mean_alpha": round(float((alphas * labels).sum() / labels.sum()), 4)}
