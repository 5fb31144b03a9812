"""Evaluation metrics on the GPU: host side of clsr_eval_metrics_compute (include/clsr_b200.h).

Replaces, for the metrics the CLSR configuration uses, the reference's per-epoch host loops
(deeprec_utils.py:554-810 cal_metric / cal_weighted_metric, called by run_eval / run_weighted_eval,
sequential_base_model.py:204-292): predictions stay on the device, one call returns auc, logloss, mean_mrr,
ndcg@k, hit@k, group_auc and wauc with the reference's dict keys and 4-decimal rounding."""
import ctypes as C

import numpy as np

from . import engine as E

GPU_METRICS = ("auc", "logloss")
GPU_PAIRWISE = ("mean_mrr", "group_auc")   # + ndcg@..., hit@...
GPU_WEIGHTED = ("wauc",)


class EvalMetrics(C.Structure):
    _fields_ = [("auc", C.c_double), ("logloss", C.c_double), ("mean_mrr", C.c_double), ("group_auc", C.c_double),
                ("wauc", C.c_double), ("ndcg", C.c_double * 8), ("hit", C.c_double * 8), ("n", C.c_int64),
                ("n_pos", C.c_int64), ("n_groups", C.c_int64), ("n_users", C.c_int64), ("status", C.c_int32)]


def _ks(metric, default=(1, 2)):
    parts = metric.split("@")
    return [int(t) for t in parts[1].split(";")] if len(parts) > 1 else list(default)


def supported(metrics, pairwise_metrics, weighted_metrics):
    ok = all(m in GPU_METRICS for m in (metrics or []))
    ok = ok and all(m in GPU_PAIRWISE or m.startswith(("ndcg", "hit")) for m in (pairwise_metrics or []))
    return ok and all(m in GPU_WEIGHTED for m in (weighted_metrics or []))


def compute(preds, labels, users=None, group=0, ks=(), stream=None):
    """preds: float32 CUDA tensor [n]; labels / users: CUDA tensors or host arrays.  Returns EvalMetrics."""
    import torch
    lib = E.load_library()
    fn = lib.clsr_eval_metrics_compute
    fn.restype = C.c_int
    fn.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32,
                   C.POINTER(EvalMetrics), C.c_void_p]
    dev = preds.device
    preds = preds.reshape(-1).contiguous().float()
    n = preds.numel()
    to_dev = lambda a, dt: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(a).reshape(-1)))).to(dev, dt).contiguous()
    labels = to_dev(labels, torch.float32)
    users = to_dev(users, torch.int32) if users is not None else None
    ks = [int(k) for k in ks][:8]
    karr = (C.c_int32 * max(len(ks), 1))(*ks)
    out = EvalMetrics()
    st = (stream if stream is not None else torch.cuda.current_stream(dev)).cuda_stream
    rc = fn(dev.index or 0, preds.data_ptr(), labels.data_ptr(), users.data_ptr() if users is not None else None, n,
            int(group), karr, len(ks), C.byref(out), C.c_void_p(st))
    if rc != 0:
        raise E.EngineError("clsr_eval_metrics_compute failed (%d)" % rc)
    return out


def metric_dict(preds, labels, users, group, metrics, pairwise_metrics, weighted_metrics):
    """The dict run_eval / run_weighted_eval return, computed on the device (same keys, order and rounding as
    cal_metric / cal_weighted_metric)."""
    ks = []
    for m in pairwise_metrics or []:
        if m.startswith(("ndcg", "hit")):
            ks += [k for k in _ks(m) if k not in ks]
    r = compute(preds, labels, users if weighted_metrics else None, group if pairwise_metrics else 0, ks)
    single = "Only one class present in y_true. ROC AUC score is not defined in that case."
    res = {}
    for m in metrics or []:
        if m == "auc":
            if r.status & 1:
                raise ValueError(single)
            res["auc"] = round(r.auc, 4)
        elif m == "logloss":
            res["logloss"] = round(r.logloss, 4)
    for m in pairwise_metrics or []:
        if m == "mean_mrr":
            res["mean_mrr"] = round(r.mean_mrr, 4)
        elif m.startswith("ndcg"):
            for k in _ks(m):
                res["ndcg@{0}".format(k)] = round(r.ndcg[ks.index(k)], 4)
        elif m.startswith("hit"):
            for k in _ks(m):
                res["hit@{0}".format(k)] = round(r.hit[ks.index(k)], 4)
        elif m == "group_auc":
            if r.status & 4:
                raise ValueError(single)
            res["group_auc"] = round(r.group_auc, 4)
    for m in weighted_metrics or []:
        if m == "wauc":
            if r.status & 2:
                raise ValueError(single)
            res["wauc"] = round(r.wauc, 4)
    return res
