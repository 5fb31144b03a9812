"""Data feed of the sequential models: text file -> padded numpy batches -> feed_dict.

Same input contract as the reference's reco_utils/recommender/deeprec/io/sequential_iterator.py
(placeholders :47-70,517; line format and time features :90-163; batching :194-302; padding,
x(1+num_ngs) replication, in-batch negative sampling and attn_labels :519-704), but a file is
parsed once into columnar, already padded arrays so that a batch is assembled with a few numpy
gathers instead of the reference's per-row Python loops.
"""
import random

import numpy as np

from reco_utils.recommender.deeprec.deeprec_utils import load_dict
from reco_utils.recommender.deeprec.io.iterator import BaseIterator, Placeholder

__all__ = ["SequentialIterator", "SASequentialIterator"]

GROUP_KEY = "_clsr_b200_group"  # private feed_dict entry: rows per shared-history group
DEVICE_BATCH_KEY = "_clsr_b200_device_batch"  # private feed_dict entry: (device dataset, line indices, num_ngs, seed)


class _Columns:
    """One parsed file, columnar: scalars per line, histories truncated to the last T events and
    left-aligned (sequential_iterator.py:589-610)."""

    def __init__(self, rows, T):
        n = len(rows)
        self.n = n
        self.label = np.array([r[0] for r in rows], np.float32)
        self.user = np.array([r[1] for r in rows], np.int32)
        self.item = np.array([r[2] for r in rows], np.int32)
        self.cate = np.array([r[3] for r in rows], np.int32)
        self.time = np.array([r[6] for r in rows], np.float32)
        self.full_len = np.array([len(r[4]) for r in rows], np.int64)
        self.length = np.minimum(self.full_len, T).astype(np.int32)
        self.ih = np.zeros((n, T), np.int32)
        self.ch = np.zeros((n, T), np.int32)
        self.tdiff = np.zeros((n, T), np.float32)
        self.tfa = np.zeros((n, T), np.float32)
        self.ttn = np.zeros((n, T), np.float32)
        for i, r in enumerate(rows):
            L = int(self.length[i])
            if L == 0:
                continue
            self.ih[i, :L] = r[4][-L:]
            self.ch[i, :L] = r[5][-L:]
            self.tdiff[i, :L] = r[7][-L:]
            self.tfa[i, :L] = r[8][-L:]
            self.ttn[i, :L] = r[9][-L:]
        self.mask = (np.arange(T)[None, :] < self.length[:, None]).astype(np.float32)


class SequentialIterator(BaseIterator):
    def __init__(self, hparams, graph, col_spliter="\t"):
        self.col_spliter = col_spliter
        self.userdict, self.itemdict, self.catedict = (
            load_dict(hparams.user_vocab), load_dict(hparams.item_vocab), load_dict(hparams.cate_vocab))
        self.max_seq_length = hparams.max_seq_length
        self.batch_size = hparams.batch_size
        self.iter_data = dict()
        self._columns = dict()
        # Batches built on the GPU (clsr_build_batch): set by the model around its own fit / eval loops; the
        # generator then yields a token {DEVICE_BATCH_KEY: ...} instead of host arrays.  Off by default, so code
        # that reads the feed arrays itself sees exactly what the reference's iterator yields.
        self.device_engine = None
        self._device_ds = dict()
        self.time_unit = hparams.time_unit
        self.graph = graph
        T = self.max_seq_length
        self.labels = Placeholder("float32", [None, 1], "label")
        self.users = Placeholder("int32", [None], "users")
        self.items = Placeholder("int32", [None], "items")
        self.cates = Placeholder("int32", [None], "cates")
        self.item_history = Placeholder("int32", [None, T], "item_history")
        self.item_cate_history = Placeholder("int32", [None, T], "item_cate_history")
        self.mask = Placeholder("int32", [None, T], "mask")
        self.time = Placeholder("float32", [None], "time")
        self.time_diff = Placeholder("float32", [None, T], "time_diff")
        self.time_from_first_action = Placeholder("float32", [None, T], "time_from_first_action")
        self.time_to_now = Placeholder("float32", [None, T], "time_to_now")

    # -- parsing ---------------------------------------------------------------------------------
    def parse_file(self, input_file):
        with open(input_file, "r") as f:
            return [self.parser_one_line(line) for line in f if line]

    def parser_one_line(self, line):
        """label \\t user \\t item \\t cate \\t time \\t item-hist csv \\t cate-hist csv \\t time-hist csv
        -> (label, user_id, item_id, cate_id, item_hist, cate_hist, time, time_diff,
        time_from_first_action, time_to_now); unknown tokens map to id 0."""
        w = line.strip().split(self.col_spliter)
        label = int(w[0])
        user_id = self.userdict.get(w[1], 0)
        item_id = self.itemdict.get(w[2], 0)
        cate_id = self.catedict.get(w[3], 0)
        now = float(w[4])
        items, cates = self.get_item_cate_history_sequence(w[5].strip().split(","), w[6].strip().split(","), user_id)
        ts = np.asarray(self.get_time_history_sequence(w[7].strip().split(",")), np.float64)
        rng = 3600 * 24 * 1000 if self.time_unit == "ms" else 3600 * 24 / 1000
        nxt = np.append(ts[1:], now)
        time_diff = np.log(np.maximum((nxt - ts) / rng, 0.5))
        time_from_first = np.log(np.maximum((nxt - ts[0]) / rng, 0.5))
        time_to_now = np.log(np.maximum((now - ts) / rng, 0.5))
        return (label, user_id, item_id, cate_id, items, cates, now, time_diff, time_from_first, time_to_now)

    def get_item_cate_history_sequence(self, item_history_words, cate_history_words, user_id):
        return self.get_item_history_sequence(item_history_words), self.get_cate_history_sequence(cate_history_words)

    def get_item_history_sequence(self, item_history_words):
        g = self.itemdict.get
        return [g(x, 0) for x in item_history_words]

    def get_cate_history_sequence(self, cate_history_words):
        g = self.catedict.get
        return [g(x, 0) for x in cate_history_words]

    def get_time_history_sequence(self, time_history_words):
        return [float(x) for x in time_history_words]

    # -- batching --------------------------------------------------------------------------------
    def _load(self, infile):
        if infile not in self.iter_data:
            self.iter_data[infile] = self.parse_file(infile)
            self._columns.pop(infile, None)
        if infile not in self._columns:
            self._columns[infile] = _Columns(self.iter_data[infile], self.max_seq_length)
        return self._columns[infile]

    def load_data_from_file(self, infile, batch_num_ngs=0, min_seq_length=1):
        """Yield feed_dicts of ``batch_size`` lines (the last one may be shorter).  With
        ``batch_num_ngs`` > 0 the lines are visited in a fresh random order and every line is
        expanded to 1 positive + ``batch_num_ngs`` in-batch negatives."""
        cols = self._load(infile)
        order = np.flatnonzero(cols.full_len >= min_seq_length)
        if batch_num_ngs > 0:
            order = order[np.random.permutation(len(order))]
        for a in range(0, len(order), self.batch_size):
            if self.device_engine is not None:
                sel = order[a:a + self.batch_size]
                if batch_num_ngs and len(sel) < 5:   # the reference drops training batches of < 5 instances (:552-554)
                    yield None
                    continue
                yield {DEVICE_BATCH_KEY: (self._dataset_on_device(infile, cols), sel.astype(np.int32), batch_num_ngs,
                                          int(np.random.randint(0, 2 ** 62))), GROUP_KEY: batch_num_ngs + 1}
                continue
            res = self._assemble(cols, order[a:a + self.batch_size], batch_num_ngs)
            batch_input = self.gen_feed_dict(res)
            yield batch_input if batch_input else None

    def _dataset_on_device(self, infile, cols):
        key = (infile, id(self.device_engine))
        if key not in self._device_ds:
            self._device_ds[key] = self.device_engine.create_dataset(
                cols.label, cols.user, cols.item, cols.cate, cols.length, cols.ih, cols.ch, cols.tfa, cols.ttn)
        return self._device_ds[key]

    def _assemble(self, cols, sel, batch_num_ngs):
        n = len(sel)
        G = batch_num_ngs + 1
        pos_item, pos_cate = cols.item[sel], cols.cate[sel]
        res = {}
        if batch_num_ngs:
            if n < 5:
                return None
            neg = np.empty((n, batch_num_ngs), np.int64)
            for i in range(n):  # negatives = other lines' positives, never the line's own item
                k = 0
                while k < batch_num_ngs:
                    j = random.randint(0, n - 1)
                    if pos_item[j] == pos_item[i]:
                        continue
                    neg[i, k] = j
                    k += 1
            src = np.concatenate([np.arange(n)[:, None], neg], 1).reshape(-1)
            rep = np.repeat(sel, G)
            labels = np.tile(np.array([1.0] + [0.0] * batch_num_ngs, np.float32), n)
            items, cates = pos_item[src], pos_cate[src]
            users = cols.user[rep].astype(np.int32)
        else:
            rep = sel
            labels = cols.label[sel]
            items, cates = pos_item, pos_cate
            users = cols.user[sel].astype(np.float32)  # the reference builds eval users as float32 (:694)
        ch = cols.ch[rep]
        mask = cols.mask[rep]
        attn = ((ch == cates[:, None]) * mask).sum(1) / np.maximum(cols.length[rep], 1)
        res["labels"] = labels.astype(np.float32).reshape(-1, 1)
        res["attn_labels"] = attn.astype(np.float32).reshape(-1, 1)
        res["users"] = users
        res["items"] = items.astype(np.int32)
        res["cates"] = cates.astype(np.int32)
        res["item_history"] = cols.ih[rep]
        res["item_cate_history"] = ch
        res["mask"] = mask
        res["time"] = cols.time[rep]
        res["time_diff"] = cols.tdiff[rep]
        res["time_from_first_action"] = cols.tfa[rep]
        res["time_to_now"] = cols.ttn[rep]
        res[GROUP_KEY] = G
        return res

    def _convert_data(self, label_list, user_list, item_list, item_cate_list, item_history_batch,
                      item_cate_history_batch, time_list, time_diff_list, time_from_first_action_list,
                      time_to_now_list, batch_num_ngs):
        """List-based entry point with the reference's signature (:304-316, :519-532)."""
        rows = list(zip(label_list, user_list, item_list, item_cate_list, item_history_batch,
                        item_cate_history_batch, time_list, time_diff_list, time_from_first_action_list,
                        time_to_now_list))
        cols = _Columns(rows, self.max_seq_length)
        return self._assemble(cols, np.arange(cols.n), batch_num_ngs)

    def gen_feed_dict(self, data_dict):
        if not data_dict:
            return dict()
        fd = {
            self.labels: data_dict["labels"], self.users: data_dict["users"], self.items: data_dict["items"],
            self.cates: data_dict["cates"], self.item_history: data_dict["item_history"],
            self.item_cate_history: data_dict["item_cate_history"], self.mask: data_dict["mask"],
            self.time: data_dict["time"], self.time_diff: data_dict["time_diff"],
            self.time_from_first_action: data_dict["time_from_first_action"],
            self.time_to_now: data_dict["time_to_now"],
        }
        if GROUP_KEY in data_dict:
            fd[GROUP_KEY] = data_dict[GROUP_KEY]
        return fd


class SASequentialIterator(SequentialIterator):
    """SequentialIterator + the ``attn_labels`` placeholder (share of the history in the target's
    category); the iterator CLSR uses (examples/00_quick_start/sequential.py:88-89)."""

    def __init__(self, hparams, graph, col_spliter="\t"):
        super(SASequentialIterator, self).__init__(hparams, graph, col_spliter)
        self.attn_labels = Placeholder("float32", [None, 1], "attn_label")

    def gen_feed_dict(self, data_dict):
        fd = super(SASequentialIterator, self).gen_feed_dict(data_dict)
        if fd:
            fd[self.attn_labels] = data_dict["attn_labels"]
        return fd
