"""Synthetic "Taobao-shaped" batches at the feed_dict level.

Mirrors what SASequentialIterator._convert_data produces
(reco_utils/recommender/deeprec/io/sequential_iterator.py:519-649): histories
left-aligned and tail-padded with id 0 / mask 0 / time 0, each line replicated
G = 1 + train_num_ngs times, negatives drawn from other lines' positives in the same
batch, time features log(max(dt / time_range, 0.5)) (:119-150).  Item ids follow a
Zipf law with id = popularity rank (frequency-sorted vocab, sequential_reviews.py:114-140).
"""
import numpy as np


class SyntheticSource:
    def __init__(self, n_items=4_000_000, n_cates=9_400, n_users=1_000_000, T=50, seed=42,
                 time_unit="s", zipf_s=1.0, full_frac=0.3):
        self.n_items, self.n_cates, self.n_users, self.T = n_items, n_cates, n_users, T
        self.rng = np.random.default_rng(seed)
        w = 1.0 / np.arange(1, n_items, dtype=np.float64) ** zipf_s
        self.cdf = np.cumsum(w)
        self.cdf /= self.cdf[-1]
        self.time_range = 3600 * 24 * 1000 if time_unit == "ms" else 3600 * 24 / 1000
        self.full_frac = full_frac

    def cate_of(self, items):
        # fixed item -> category map; id 0 (padding / OOV) keeps category 0
        c = (items.astype(np.int64) * 2654435761 % 4294967296) % (self.n_cates - 1) + 1
        return np.where(items == 0, 0, c).astype(np.int32)

    def sample_items(self, n):
        return (np.searchsorted(self.cdf, self.rng.random(n)) + 1).astype(np.int32)

    def lines(self, S):
        """S file lines: (user, item, cate, len, hist items [S,T], times)."""
        T, rng = self.T, self.rng
        length = rng.integers(1, T + 1, S)
        length[rng.random(S) < self.full_frac] = T
        ih = self.sample_items(S * T).reshape(S, T)
        live = np.arange(T)[None, :] < length[:, None]
        ih = np.where(live, ih, 0).astype(np.int32)
        gaps = rng.exponential(3600.0, (S, T + 1))
        ts = np.cumsum(gaps, 1) + 1.5e9
        if self.time_range > 1e6:
            ts *= 1000.0
        return dict(users=rng.integers(1, self.n_users, S).astype(np.int32),
                    items=self.sample_items(S), length=length, item_history=ih,
                    ts_hist=ts[:, :T], ts_now=ts[np.arange(S), length])

    def batch(self, S, num_ngs=4):
        """One training feed (num_ngs > 0: rows = S*(1+num_ngs)) or eval feed (num_ngs = 0)."""
        L = self.lines(S)
        T, G = self.T, num_ngs + 1
        ih, length = L["item_history"], L["length"]
        live = np.arange(T)[None, :] < length[:, None]
        ch = self.cate_of(ih)
        now = L["ts_now"][:, None]
        first = L["ts_hist"][:, :1]
        ttn = np.log(np.maximum((now - L["ts_hist"]) / self.time_range, 0.5))
        nxt = np.concatenate([L["ts_hist"][:, 1:], now], 1)
        nxt = np.where(np.arange(T)[None, :] == length[:, None] - 1, now, nxt)
        tfa = np.log(np.maximum((nxt - first) / self.time_range, 0.5))
        tdiff = np.log(np.maximum((nxt - L["ts_hist"]) / self.time_range, 0.5))
        z = lambda a: np.where(live, a, 0.0).astype(np.float32)
        items = L["items"]
        if num_ngs:
            neg = self.rng.integers(0, S, (S, num_ngs))
            for _ in range(16):  # reject negatives equal to the positive (:625-626)
                bad = items[neg] == items[:, None]
                if not bad.any():
                    break
                neg[bad] = self.rng.integers(0, S, int(bad.sum()))
            tgt = np.concatenate([items[:, None], items[neg]], 1).reshape(-1)
            labels = np.tile(np.array([1.0] + [0.0] * num_ngs, np.float32), S).reshape(-1, 1)
        else:
            tgt = items
            labels = (self.rng.random(S) < 0.2).astype(np.float32).reshape(-1, 1)
        rep = lambda a: np.ascontiguousarray(np.repeat(a, G, axis=0))
        cates = self.cate_of(tgt)
        hist_c = rep(ch)
        attn = ((hist_c == cates[:, None]) & rep(live)).sum(1) / rep(length)
        return {
            "labels": labels, "attn_labels": attn.astype(np.float32).reshape(-1, 1),
            "users": rep(L["users"]), "items": tgt.astype(np.int32), "cates": cates,
            "item_history": rep(ih), "item_cate_history": hist_c,
            "mask": rep(live.astype(np.int32)),
            "time": rep(L["ts_now"].astype(np.float32)),
            "time_diff": rep(z(tdiff)), "time_from_first_action": rep(z(tfa)), "time_to_now": rep(z(ttn)),
        }
