#coding=utf-8
"""Quick-start driver for the B200 CLSR build: same flags, same flow and same printed output as
the reference's examples/00_quick_start/sequential.py (flags :36-68, hparam assembly :120-154,
main :307-376).  The reference's own script also runs unchanged against this repository (put
clsr_b200/shims on PYTHONPATH for its `import tensorflow`); this re-implementation exists so the
repository is self-contained.

    cd examples/00_quick_start && python sequential.py --dataset taobao [--only_test]
"""
import os
import sys
import time

from absl import app, flags

sys.path.append(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
try:
    import tensorflow as tf
except ImportError:
    sys.path.append(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "clsr_b200", "shims"))
    import tensorflow as tf

from reco_utils.dataset.sequential_reviews import data_preprocessing
from reco_utils.recommender.deeprec.deeprec_utils import prepare_hparams
from reco_utils.recommender.deeprec.io.sequential_iterator import SASequentialIterator
from reco_utils.recommender.deeprec.models.sequential.clsr import CLSRModel

FLAGS = flags.FLAGS
flags.DEFINE_string("dataset", "taobao", "Dataset name.")
flags.DEFINE_integer("gpu_id", 0, "GPU ID.")
flags.DEFINE_integer("val_num_ngs", 4, "Negatives per positive in the validation file.")
flags.DEFINE_integer("test_num_ngs", 99, "Negatives per positive in the test file.")
flags.DEFINE_integer("batch_size", 500, "Batch size (file lines).")
flags.DEFINE_string("save_path", "", "Save path.")
flags.DEFINE_string("contrastive_loss", "triplet", "Contrastive loss: bpr or triplet.")
flags.DEFINE_integer("contrastive_length_threshold", 5, "Minimum sequence length to apply the contrastive loss.")
flags.DEFINE_integer("contrastive_recent_k", 3, "Most recent k embeddings form the short-term proxy.")
flags.DEFINE_string("name", "taobao-clsr-debug", "Experiment name.")
flags.DEFINE_string("model", "CLSR", "Model name.")
flags.DEFINE_boolean("only_test", False, "Only test and do not train.")
flags.DEFINE_boolean("write_prediction_to_file", False, "Whether to write predictions to a file.")
flags.DEFINE_boolean("manual_alpha", False, "Use a predefined alpha for long/short fusion.")
flags.DEFINE_float("manual_alpha_value", 0.5, "Predefined alpha value.")
flags.DEFINE_boolean("interest_evolve", True, "Model interest evolution with a GRU.")
flags.DEFINE_boolean("predict_long_short", True, "Predict whether the next interaction is long- or short-term driven.")
flags.DEFINE_integer("is_clip_norm", 1, "Whether to clip gradient norms.")
flags.DEFINE_string("sequential_model", "time4lstm", "gru, lstm or time4lstm.")
flags.DEFINE_integer("epochs", 100, "Number of epochs.")
flags.DEFINE_integer("early_stop", 5, "Patience for early stop.")
flags.DEFINE_string("data_path", os.path.join("..", "..", "tests", "resources", "deeprec", "sequential"), "Data path.")
flags.DEFINE_integer("train_num_ngs", 4, "Negatives per positive for training.")
flags.DEFINE_float("sample_rate", 1.0, "Fraction of samples for training and testing.")
flags.DEFINE_float("embed_l2", 1e-6, "L2 regulation for embeddings.")
flags.DEFINE_float("layer_l2", 1e-6, "L2 regulation for layers.")
flags.DEFINE_float("attn_loss_weight", 0.001, "Loss weight for supervised attention.")
flags.DEFINE_float("triplet_margin", 1.0, "Margin of the triplet loss.")
flags.DEFINE_float("discrepancy_loss_weight", 0.01, "Weight of the long/short user-embedding discrepancy loss.")
flags.DEFINE_float("contrastive_loss_weight", 0.1, "Weight of the contrastive loss.")
flags.DEFINE_float("learning_rate", 0.001, "Learning rate.")
flags.DEFINE_integer("show_step", 500, "Steps between progress lines.")


def get_model(f, model_path, summary_path, user_vocab, item_vocab, cate_vocab, train_num_ngs):
    if f.model != "CLSR":
        raise NotImplementedError("only --model CLSR is part of the B200 build (got %s)" % f.model)
    if f.dataset == "kuaishou":
        pairwise, T, unit = ["mean_mrr", "ndcg@1;2"], 250, "ms"
    else:
        pairwise, T, unit = ["mean_mrr", "ndcg@2;4;6", "hit@2;4;6"], 50, "s"
    yaml_file = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "reco_utils", "recommender",
                             "deeprec", "config", "clsr.yaml")
    hparams = prepare_hparams(
        yaml_file, embed_l2=f.embed_l2, layer_l2=f.layer_l2, contrastive_loss=f.contrastive_loss,
        triplet_margin=f.triplet_margin, discrepancy_loss_weight=f.discrepancy_loss_weight,
        contrastive_loss_weight=f.contrastive_loss_weight, learning_rate=f.learning_rate, epochs=f.epochs,
        EARLY_STOP=f.early_stop, batch_size=f.batch_size, show_step=f.show_step, MODEL_DIR=model_path,
        SUMMARIES_DIR=summary_path, user_vocab=user_vocab, item_vocab=item_vocab, cate_vocab=cate_vocab,
        need_sample=True, train_num_ngs=train_num_ngs, max_seq_length=T, pairwise_metrics=pairwise,
        weighted_metrics=["wauc"], time_unit=unit, manual_alpha=f.manual_alpha,
        manual_alpha_value=f.manual_alpha_value, interest_evolve=f.interest_evolve,
        predict_long_short=f.predict_long_short, is_clip_norm=f.is_clip_norm,
        contrastive_length_threshold=f.contrastive_length_threshold, contrastive_recent_k=f.contrastive_recent_k,
        sequential_model=f.sequential_model)
    return CLSRModel(hparams, SASequentialIterator, seed=None)


def main(argv):
    f = FLAGS
    print("System version: {}".format(sys.version))
    print("Tensorflow version: {}".format(tf.__version__))
    print("start experiment")
    data_path = os.path.join(f.data_path, f.dataset)
    p = lambda n: os.path.join(data_path, n)
    train_file, valid_file, test_file = p("train_data"), p("valid_data"), p("test_data")
    user_vocab, item_vocab, cate_vocab = p("user_vocab.pkl"), p("item_vocab.pkl"), p("category_vocab.pkl")
    if not os.path.exists(train_file):
        reviews = p("UserBehavior.csv" if f.dataset == "taobao" else "kuaishou.csv")
        data_preprocessing(reviews, p(""), train_file, valid_file, test_file, user_vocab, item_vocab, cate_vocab,
                           sample_rate=f.sample_rate, valid_num_ngs=f.val_num_ngs, test_num_ngs=f.test_num_ngs,
                           dataset=f.dataset)
    save_path = os.path.join(f.save_path, f.model, f.name)
    model_path, summary_path = os.path.join(save_path, "model/"), os.path.join(save_path, "summary/")
    model = get_model(f, model_path, summary_path, user_vocab, item_vocab, cate_vocab, f.train_num_ngs)
    if f.only_test:
        model.load_model(tf.train.latest_checkpoint(model_path))
        print(model.run_weighted_eval(test_file, num_ngs=f.test_num_ngs))
        return
    start = time.time()
    model = model.fit(train_file, valid_file, valid_num_ngs=f.val_num_ngs, eval_metric="wauc")
    print("Time cost for training is {0:.2f} mins".format((time.time() - start) / 60.0))
    model.load_model(tf.train.latest_checkpoint(model_path))
    res = model.run_weighted_eval(test_file, num_ngs=f.test_num_ngs)
    print(f.name)
    print(res)
    if f.write_prediction_to_file:
        model.predict(test_file, p("output.txt"))


if __name__ == "__main__":
    app.run(main)
