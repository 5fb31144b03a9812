"""Train / eval / predict loops shared by the sequential models.

Keeps the surface and behaviour of the reference's
reco_utils/recommender/deeprec/models/sequential/sequential_base_model.py: ``fit`` with
per-epoch weighted validation, early stop and save-on-improvement (:111-202), ``run_eval``
(:204-237), ``run_weighted_eval`` (:244-292), ``predict`` (:326-347).
"""
import abc
import os

import numpy as np

from reco_utils.recommender.deeprec.deeprec_utils import (cal_mean_alpha_metric, cal_metric,
                                                          cal_weighted_metric)
from reco_utils.recommender.deeprec.io.iterator import Placeholder
from reco_utils.recommender.deeprec.models.base_model import BaseModel, SummaryWriter, _Graph

__all__ = ["SequentialBaseModel"]


class SequentialBaseModel(BaseModel):
    def __init__(self, hparams, iterator_creator, graph=None, seed=None):
        self.hparams = hparams
        self.need_sample = hparams.need_sample
        self.train_num_ngs = hparams.train_num_ngs
        if self.train_num_ngs is None:
            raise ValueError("Please confirm the number of negative samples for each positive instance.")
        self.min_seq_length = hparams.min_seq_length if "min_seq_length" in hparams else 1
        self.hidden_size = hparams.hidden_size if "hidden_size" in hparams else None
        self.graph = _Graph() if not graph else graph
        self.embedding_keeps = Placeholder("float32", None, "embedding_keeps")
        self.embedding_keep_prob_train = None
        self.embedding_keep_prob_test = None
        super().__init__(hparams, iterator_creator, graph=self.graph, seed=seed)

    @abc.abstractmethod
    def _build_seq_graph(self):
        pass

    def _build_graph(self):
        hparams = self.hparams
        self.keep_prob_train = 1 - np.array(hparams.dropout)
        self.keep_prob_test = np.ones_like(hparams.dropout)
        self.embedding_keep_prob_train = 1.0 - hparams.embedding_dropout
        self.embedding_keep_prob_test = 1.0
        self._build_seq_graph()

    # ---- loops ---------------------------------------------------------------------------------
    def batch_train(self, file_iterator, train_sess):
        step, epoch_loss = 0, 0
        for batch_data_input in file_iterator:
            if batch_data_input:
                step_result = self.train(train_sess, batch_data_input)
                step_loss, step_data_loss, summary = step_result[2], step_result[3], step_result[-1]
                if self.hparams.write_tfevents and self.hparams.SUMMARIES_DIR:
                    self.writer.add_summary(summary, step)
                epoch_loss += step_loss
                step += 1
                if step % self.hparams.show_step == 0:
                    print("step {0:d} , total_loss: {1:.4f}, data_loss: {2:.4f}".format(step, step_loss, step_data_loss))
        return epoch_loss

    def fit(self, train_file, valid_file, valid_num_ngs, eval_metric="group_auc"):
        if not self.need_sample and self.train_num_ngs < 1:
            raise ValueError(
                "Please specify a positive integer of negative numbers for training without sampling needed.")
        if valid_num_ngs < 1:
            raise ValueError("Please specify a positive integer of negative numbers for validation.")
        if self.need_sample and self.train_num_ngs < 1:
            self.train_num_ngs = 1
        if self.hparams.write_tfevents and self.hparams.SUMMARIES_DIR:
            self.writer = SummaryWriter(self.hparams.SUMMARIES_DIR, self.sess.graph)
        train_sess = self.sess
        eval_info = list()
        best_metric, self.best_epoch = 0, 0
        for epoch in range(1, self.hparams.epochs + 1):
            self.hparams.current_epoch = epoch
            file_iterator = self.iterator.load_data_from_file(
                train_file, min_seq_length=self.min_seq_length, batch_num_ngs=self.train_num_ngs)
            self.batch_train(file_iterator, train_sess)
            valid_res = self.run_weighted_eval(valid_file, valid_num_ngs)
            print("eval valid at epoch {0}: {1}".format(
                epoch, ",".join(["" + str(k) + ":" + str(v) for k, v in valid_res.items()])))
            eval_info.append((epoch, valid_res))
            progress = False
            early_stop = self.hparams.EARLY_STOP
            if valid_res[eval_metric] > best_metric:
                best_metric = valid_res[eval_metric]
                self.best_epoch = epoch
                progress = True
            elif early_stop > 0 and epoch - self.best_epoch >= early_stop:
                print("early stop at epoch {0}!".format(epoch))
                break
            if self.hparams.save_model and self.hparams.MODEL_DIR:
                if not os.path.exists(self.hparams.MODEL_DIR):
                    os.makedirs(self.hparams.MODEL_DIR)
                if progress:
                    self.saver.save(sess=train_sess, save_path=self.hparams.MODEL_DIR + "epoch_" + str(epoch))
        if self.hparams.write_tfevents and getattr(self, "writer", None) is not None:
            self.writer.close()
        print(eval_info)
        print("best epoch: {0}".format(self.best_epoch))
        return self

    def run_eval(self, filename, num_ngs):
        preds, labels = [], []
        group = num_ngs + 1
        for batch_data_input in self.iterator.load_data_from_file(
                filename, min_seq_length=self.min_seq_length, batch_num_ngs=0):
            if batch_data_input:
                step_pred, step_labels = self.eval(self.sess, batch_data_input)
                preds.append(np.reshape(step_pred, -1))
                labels.append(np.reshape(step_labels, -1))
        preds, labels = np.concatenate(preds), np.concatenate(labels)
        res = cal_metric(labels, preds, self.hparams.metrics)
        res.update(cal_metric(labels.reshape(-1, group), preds.reshape(-1, group), self.hparams.pairwise_metrics))
        return res

    def run_weighted_eval(self, filename, num_ngs, calc_mean_alpha=False, manual_alpha=False):
        users, preds, labels, alphas = [], [], [], []
        group = num_ngs + 1
        for batch_data_input in self.iterator.load_data_from_file(
                filename, min_seq_length=self.min_seq_length, batch_num_ngs=0):
            if batch_data_input:
                if not calc_mean_alpha:
                    step_user, step_pred, step_labels = self.eval_with_user(self.sess, batch_data_input)
                else:
                    step_user, step_pred, step_labels, step_alpha = self.eval_with_user_and_alpha(
                        self.sess, batch_data_input)
                    alphas.append(np.reshape(step_alpha, -1))
                users.append(np.reshape(step_user, -1))
                preds.append(np.reshape(step_pred, -1))
                labels.append(np.reshape(step_labels, -1))
        users, preds, labels = np.concatenate(users), np.concatenate(preds), np.concatenate(labels)
        res = cal_metric(labels, preds, self.hparams.metrics)
        res.update(cal_metric(labels.reshape(-1, group), preds.reshape(-1, group), self.hparams.pairwise_metrics))
        res.update(cal_weighted_metric(users, preds, labels, self.hparams.weighted_metrics))
        if calc_mean_alpha:
            alphas = np.concatenate(alphas)
            if manual_alpha:
                alphas = alphas[0]
            res.update(cal_mean_alpha_metric(alphas, labels))
        return res

    def predict(self, infile_name, outfile_name):
        with open(outfile_name, "w") as wt:
            for batch_data_input in self.iterator.load_data_from_file(infile_name, batch_num_ngs=0):
                if batch_data_input:
                    step_pred = np.reshape(self.infer(self.sess, batch_data_input), -1)
                    wt.write("\n".join(map(str, step_pred)))
                    wt.write("\n")
        return self
