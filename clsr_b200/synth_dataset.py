"""Write a synthetic dataset directory in the on-disk format the reference's preprocessing
produces (reco_utils/dataset/sequential_reviews.py:77-199): ``train_data`` / ``valid_data`` /
``test_data`` with 8 tab-separated columns

    label  user  item  category  timestamp  item-history csv  category-history csv  time-history csv

(valid/test: each positive line followed by its ``num_ngs`` negative lines, same user and
history), and ``user_vocab.pkl`` / ``item_vocab.pkl`` / ``category_vocab.pkl`` mapping raw tokens
to ids by descending frequency with ``default_*`` at id 0."""
import os
import pickle
import sys
from collections import Counter

import numpy as np


def write_dataset(path, n_users=200, n_items=500, n_cates=20, T=50, train_lines=600, valid_users=40,
                  test_users=40, valid_num_ngs=4, test_num_ngs=9, seed=0, time_unit="s"):
    rng = np.random.default_rng(seed)
    os.makedirs(path, exist_ok=True)
    item_cate = rng.integers(0, n_cates, n_items)
    pop = 1.0 / np.arange(1, n_items + 1)
    pop /= pop.sum()
    scale = 1000.0 if time_unit == "ms" else 1.0

    def history():
        n = int(rng.integers(1, int(T * 1.4)))
        items = rng.choice(n_items, size=n, p=pop)
        ts = 1.5e9 * scale + np.cumsum(rng.exponential(3600.0 * scale, n + 1))
        return items, ts

    def line(label, user, item, items, ts):
        return "\t".join([
            str(label), "u%d" % user, "i%d" % item, "c%d" % item_cate[item], "%d" % ts[-1],
            ",".join("i%d" % i for i in items), ",".join("c%d" % item_cate[i] for i in items),
            ",".join("%d" % t for t in ts[:-1])]) + "\n"

    with open(os.path.join(path, "train_data"), "w") as f:
        for _ in range(train_lines):
            items, ts = history()
            f.write(line(1, int(rng.integers(0, n_users)), int(rng.choice(n_items, p=pop)), items, ts))
    for name, users, ngs in (("valid_data", valid_users, valid_num_ngs), ("test_data", test_users, test_num_ngs)):
        with open(os.path.join(path, name), "w") as f:
            for _ in range(users):
                items, ts = history()
                user = int(rng.integers(0, n_users))
                pos = int(rng.choice(n_items, p=pop))
                f.write(line(1, user, pos, items, ts))
                negs = set()
                while len(negs) < ngs:
                    c = int(rng.choice(n_items, p=pop))
                    if c != pos:
                        negs.add(c)
                for c in negs:
                    f.write(line(0, user, c, items, ts))
    users, its, cats = Counter(), Counter(), Counter()
    with open(os.path.join(path, "train_data")) as f:
        for ln in f:
            a = ln.strip("\n").split("\t")
            users[a[1]] += 1
            its[a[2]] += 1
            cats[a[3]] += 1
            its.update(a[5].split(","))
            cats.update(a[6].split(","))
    for fname, default, cnt in (("user_vocab.pkl", "default_uid", users), ("item_vocab.pkl", "default_mid", its),
                                ("category_vocab.pkl", "default_cat", cats)):
        voc = {default: 0}
        for i, (k, _) in enumerate(sorted(cnt.items(), key=lambda x: x[1], reverse=True)):
            voc[k] = i + 1
        with open(os.path.join(path, fname), "wb") as f:
            pickle.dump(voc, f)
    return path


if __name__ == "__main__":
    print(write_dataset(sys.argv[1] if len(sys.argv) > 1 else "synthetic_taobao"))
