"""Model base: construction order, the feed_dict control placeholders, checkpoint I/O.

Reference: reco_utils/recommender/deeprec/models/base_model.py:17-71 (constructor),
:343-392 (train / eval / infer), :394-410 (load_model).  The TensorFlow graph + session of the
reference are replaced by one ``clsr_b200`` engine handle; ``self.sess`` / ``self.graph`` stay as
opaque tokens so user code that passes them around keeps working.
"""
import abc
import json
import os

import numpy as np

from clsr_b200 import tf_bundle
from reco_utils.recommender.deeprec.io.iterator import Placeholder

__all__ = ["BaseModel"]


class _Graph:
    """Token standing in for tf.Graph."""

    def as_default(self):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _Session:
    """Token standing in for tf.Session: owns nothing, ``run`` is not available."""

    def __init__(self, graph):
        self.graph = graph

    def run(self, *a, **k):
        raise NotImplementedError("the B200 build has no graph executor: call model.train/eval/infer")


class Saver:
    """tf.train.Saver look-alike writing/reading TF tensor bundles (``<path>.index`` +
    ``<path>.data-00000-of-00001``) and the text ``checkpoint`` state file."""

    def __init__(self, model, max_to_keep=5):
        self.model, self.max_to_keep, self.saved = model, max_to_keep, []

    def save(self, sess, save_path):
        tensors = self.model._export_variables()
        tf_bundle.write_bundle(save_path, tensors)
        d = os.path.dirname(os.path.abspath(save_path))
        if save_path not in self.saved:
            self.saved.append(save_path)
        self.saved = self.saved[-max(int(self.max_to_keep or 1), 1):]
        tf_bundle.update_checkpoint_state(d, save_path, self.saved)
        return save_path

    def restore(self, sess, save_path):
        self.model._import_variables(tf_bundle.read_bundle(save_path))


class SummaryWriter:
    """Scalar log in place of tf.summary.FileWriter (one JSON line per step)."""

    def __init__(self, logdir, graph=None):
        os.makedirs(logdir, exist_ok=True)
        self.f = open(os.path.join(logdir, "scalars.jsonl"), "a")

    def add_summary(self, summary, step):
        if summary:
            self.f.write(json.dumps(dict(summary, step=int(step))) + "\n")

    def close(self):
        self.f.close()


class BaseModel:
    def __init__(self, hparams, iterator_creator, graph=None, seed=None):
        self.seed = seed
        np.random.seed(seed)
        self.graph = graph if graph is not None else _Graph()
        self.iterator = iterator_creator(hparams, self.graph)
        self.train_num_ngs = hparams.train_num_ngs if "train_num_ngs" in hparams else None
        self.hparams = hparams
        self.layer_params, self.embed_params, self.cross_params = [], [], []
        self.layer_keeps = Placeholder("float32", None, "layer_keeps")
        self.keep_prob_train = None
        self.keep_prob_test = None
        self.is_train_stage = Placeholder("bool", (), "is_training")
        self.group = Placeholder("int32", (), "group")
        self._build_graph()
        self.saver = Saver(self, max_to_keep=self.hparams.epochs)
        self.sess = _Session(self.graph)

    @abc.abstractmethod
    def _build_graph(self):
        pass

    def load_model(self, model_path=None):
        """Restore variables from a checkpoint prefix; IOError if it cannot be read."""
        act_path = self.hparams.load_saved_model
        if model_path is not None:
            act_path = model_path
        try:
            self.saver.restore(self.sess, act_path)
        except Exception:
            raise IOError("Failed to find any matching files for {0}".format(act_path))
