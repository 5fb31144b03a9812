"""Offline preprocessing entry point of the quick-start CLI.

The reference turns the raw Taobao / Kuaishou CSV into train/valid/test text files and vocabulary
pickles (reco_utils/dataset/sequential_reviews.py:27-74); that one-off pandas ETL is outside the
accelerated path.  What matters downstream is its OUTPUT FORMAT, which
``clsr_b200.synth_dataset.write_dataset`` produces for synthetic data and
``SequentialIterator.parser_one_line`` consumes."""
import os

__all__ = ["data_preprocessing"]


def data_preprocessing(reviews_file, meta_file, train_file, valid_file, test_file, user_vocab, item_vocab,
                       cate_vocab, sample_rate=0.01, valid_num_ngs=4, test_num_ngs=9, dataset="taobao",
                       is_history_expanding=True):
    if all(os.path.exists(p) for p in (train_file, valid_file, test_file, user_vocab, item_vocab, cate_vocab)):
        return
    raise NotImplementedError(
        "raw-CSV preprocessing is not part of the B200 CLSR build: run the reference's "
        "data_preprocessing once on the host, or create a synthetic dataset directory with "
        "`python -m clsr_b200.synth_dataset <dir>`; missing: %s" % train_file)
