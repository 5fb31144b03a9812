"""Abstract iterator surface (reference: reco_utils/recommender/deeprec/io/iterator.py:9-24)."""
import abc


class Placeholder:
    """Key object of a feed_dict; stands where the reference holds a tf.placeholder."""

    def __init__(self, dtype, shape=None, name=None):
        self.dtype, self.shape, self.name = dtype, shape, name

    def __repr__(self):
        return "<Placeholder %s %s %s>" % (self.name, self.dtype, self.shape)


class BaseIterator(object):
    @abc.abstractmethod
    def parser_one_line(self, line):
        pass

    @abc.abstractmethod
    def load_data_from_file(self, infile):
        pass

    @abc.abstractmethod
    def _convert_data(self, labels, features):
        pass

    @abc.abstractmethod
    def gen_feed_dict(self, data_dict):
        pass
