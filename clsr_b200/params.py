"""Parameter inventory of the CLSR graph: TF variable names, shapes, initialisers.

Names and shapes follow what the reference's graph creates (clsr.py:84-101,358-362;
sequential_base_model.py:354-379; base_model.py:644-696; rnn_cell_implement.py:147-198,
227) and what its shipped checkpoint contains, so tensor bundles interchange 1:1.
Initialisers follow the saved graph: truncated-normal(0.01) for tables, attention_mat
and every w_nn_*; glorot-uniform for GRU/Time4LSTM kernels and all Time4LSTM extras
(including the 1-D ones); GRU gate bias 1; BN gamma 1 / moving variance 1.
"""
import math
from collections import OrderedDict

import numpy as np

SC = "sequential/clsr/"
EMB = "sequential/embedding/"
LOGIT = "sequential/logit_fcn/nn_part/"

TABLE_NAMES = ["item_embedding", "cate_embedding", "user_embedding",
               "user_long_embedding", "user_short_embedding"]


def _fcn(spec, scope, in_dim, sizes):
    last = in_dim
    for i, n in enumerate(sizes):
        spec.append((scope + "w_nn_layer%d" % i, (last, n), "tnormal", True))
        spec.append((scope + "b_nn_layer%d" % i, (n,), "zeros", True))
        bn = scope + ("batch_normalization/" if i == 0 else "batch_normalization_%d/" % i)
        spec.append((bn + "gamma", (n,), "ones", True))
        spec.append((bn + "beta", (n,), "zeros", True))
        spec.append((bn + "moving_mean", (n,), "zeros", False))
        spec.append((bn + "moving_variance", (n,), "ones", False))
        last = n
    spec.append((scope + "w_nn_output", (last, 1), "tnormal", True))
    spec.append((scope + "b_nn_output", (1,), "zeros", True))


def dense_spec(D, U, H, att_sizes, layer_sizes, interest_evolve=True, predict_long_short=True, manual_alpha=False,
               sequential_model="time4lstm"):
    """Ordered [(name, shape, init, trainable)] for every non-table variable.  The three flags are the graph
    variants of _build_seq_graph (clsr.py:159-274): without interest_evolve the short_term_intention GRU does not
    exist, with manual_alpha neither the causal2 GRU nor the alpha MLP, without predict_long_short the causal2 GRU
    is absent and the alpha MLP's input loses its final state."""
    assert H == D, "the CLSR graph only type-checks when hidden_size == item_dim + cate_dim"
    s = []
    lt = SC + "long_term/attention_fcn/"
    s.append((lt + "attention_mat", (D, U), "tnormal", True))
    _fcn(s, lt + "att_fcn/nn_part/", 4 * U, att_sizes)
    st = SC + "short_term/"
    Q = U + D
    s.append((st + "attention_fcn/attention_mat", (H, Q), "tnormal", True))
    _fcn(s, st + "attention_fcn/att_fcn/nn_part/", 4 * Q, att_sizes)
    has_gru2 = predict_long_short and not manual_alpha
    for scope, units, present in ((st + "short_term_intention/gru_cell/", U, interest_evolve),
                                  (SC + "causal2/causal2/gru_cell/", H, has_gru2)):
        if not present:
            continue
        s.append((scope + "gates/kernel", (D + units, 2 * units), "glorot", True))
        s.append((scope + "gates/bias", (2 * units,), "ones", True))
        s.append((scope + "candidate/kernel", (D + units, units), "glorot", True))
        s.append((scope + "candidate/bias", (units,), "zeros", True))
    if sequential_model == "lstm":      # tf.nn.rnn_cell.LSTMCell under scope simple_lstm (clsr.py:209-216)
        tl = st + "simple_lstm/lstm_cell/"
        s.append((tl + "kernel", (D + H, 4 * H), "glorot", True))
        s.append((tl + "bias", (4 * H,), "zeros", True))
    elif sequential_model == "gru":     # tf.nn.rnn_cell.GRUCell under scope simple_gru (clsr.py:201-208)
        tl = st + "simple_gru/gru_cell/"
        s.append((tl + "gates/kernel", (D + H, 2 * H), "glorot", True))
        s.append((tl + "gates/bias", (2 * H,), "ones", True))
        s.append((tl + "candidate/kernel", (D + H, H), "glorot", True))
        s.append((tl + "candidate/bias", (H,), "zeros", True))
    else:
        assert sequential_model == "time4lstm", sequential_model
        tl = st + "time4lstm/time4lstm_cell/"
        s.append((tl + "kernel", (D + H, 4 * H), "glorot", True))
        s.append((tl + "bias", (4 * H,), "zeros", True))
        for n in ("_time_input_w1", "_time_input_bias1", "_time_input_w2", "_time_input_bias2",
                  "_time_bias1", "_time_bias2"):
            s.append((tl + n, (H,), "glorot", True))
        for n in ("_time_kernel_w1", "_time_kernel_w2"):
            s.append((tl + n, (D, H), "glorot", True))
        for n in ("_time_kernel_t1", "_time_kernel_t2", "_o_kernel_t1", "_o_kernel_t2"):
            s.append((tl + n, (H, H), "glorot", True))
    if not manual_alpha:
        _fcn(s, SC + "fcn_alpha/nn_part/", (H if has_gru2 else 0) + 2 * D + H + 1, att_sizes)
    _fcn(s, LOGIT, H + D, layer_sizes)
    return s


def table_spec(n_items, n_cates, n_users, Di, Dc, U):
    return [(EMB + "item_embedding", (n_items, Di)), (EMB + "cate_embedding", (n_cates, Dc)),
            (EMB + "user_embedding", (n_users, U)), (EMB + "user_long_embedding", (n_users, U)),
            (EMB + "user_short_embedding", (n_users, U))]


def tnormal(rng, shape, std):
    """tf.truncated_normal_initializer: redraw anything beyond two standard deviations."""
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * std).astype(np.float32)


def init_array(rng, shape, kind, init_value=0.01):
    if kind == "tnormal":
        return tnormal(rng, shape, init_value)
    if kind == "glorot":
        fan_in, fan_out = (shape[0], shape[0]) if len(shape) == 1 else (shape[0], shape[1])
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return rng.uniform(-lim, lim, shape).astype(np.float32)
    if kind == "ones":
        return np.ones(shape, np.float32)
    if kind == "zeros":
        return np.zeros(shape, np.float32)
    raise ValueError(kind)


def init_params(n_items, n_cates, n_users, Di=32, Dc=8, U=40, H=40, att_sizes=(80, 40),
                layer_sizes=(100, 64), seed=None, init_value=0.01, tables=True, **variant):
    """Fresh parameters as an ordered {name: float32 array} (variant: the dense_spec flags)."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shape, kind, _ in dense_spec(Di + Dc, U, H, list(att_sizes), list(layer_sizes), **variant):
        out[name] = init_array(rng, shape, kind, init_value)
    if tables:
        for name, shape in table_spec(n_items, n_cates, n_users, Di, Dc, U):
            out[name] = tnormal(rng, shape, init_value)
    return out
