"""Minimal stand-in for ``import tensorflow`` so the reference's quick-start CLI
(examples/00_quick_start/sequential.py:13,312,352,369) runs where TensorFlow 1.15 is not
installable.  Put this directory on PYTHONPATH only when real TensorFlow is absent.  It provides
exactly what the CLI touches: ``tf.__version__`` and ``tf.train.latest_checkpoint``."""
from clsr_b200 import tf_bundle as _tb

__version__ = "1.15.2-clsr_b200-shim"


class _Train:
    @staticmethod
    def latest_checkpoint(checkpoint_dir, latest_filename=None):
        return _tb.latest_checkpoint(checkpoint_dir)


train = _Train()
