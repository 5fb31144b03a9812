"""CLSRModel on the B200 engine.

Reference: reco_utils/recommender/deeprec/models/sequential/clsr.py.  The whole of
``_build_seq_graph`` (:137-277), ``_attention_fcn`` (:343-381), the losses (:22-82) and the
optimizer step (base_model.py:281-297) execute inside libclsr_b200.so; this class keeps the
reference's constructor, ``train`` / ``eval`` / ``eval_with_user`` / ``eval_with_user_and_alpha`` /
``infer`` signatures and return shapes (:383-408; base_model.py:366-392;
sequential_base_model.py:294-324).
"""
import os
import warnings

import numpy as np

from clsr_b200 import params as P
from clsr_b200.engine import Engine, detect_group, normalize_feed
from reco_utils.recommender.deeprec.deeprec_utils import load_dict
from reco_utils.recommender.deeprec.io.sequential_iterator import DEVICE_BATCH_KEY, GROUP_KEY
from reco_utils.recommender.deeprec.models.sequential.sequential_base_model import SequentialBaseModel

__all__ = ["CLSRModel"]


class CLSRModel(SequentialBaseModel):
    def _build_seq_graph(self):
        hp = self.hparams
        unsupported = []
        if hp.sequential_model not in ("time4lstm", "lstm"): unsupported.append("sequential_model=%s" % hp.sequential_model)
        if hp.loss != "softmax": unsupported.append("loss=%s" % hp.loss)
        if hp.enable_BN is not True: unsupported.append("enable_BN=%s" % hp.enable_BN)
        if list(hp.activation) != ["relu", "relu"]: unsupported.append("activation=%s" % hp.activation)
        if hp.user_dropout or hp.embedding_dropout or any(hp.dropout): unsupported.append("dropout")
        if hp.embed_l1 or hp.layer_l1 or hp.cross_l1 or hp.cross_l2: unsupported.append("l1 / cross regularisers")
        if hp.method != "classification": unsupported.append("method=%s" % hp.method)
        if len(hp.att_fcn_layer_sizes) != 2 or len(hp.layer_sizes) != 2: unsupported.append("MLP depth != 2")
        if unsupported:
            raise NotImplementedError("not available on the B200 CLSR path: " + ", ".join(unsupported))
        self.user_vocab_length = len(load_dict(hp.user_vocab))
        self.item_vocab_length = len(load_dict(hp.item_vocab))
        self.cate_vocab_length = len(load_dict(hp.cate_vocab))
        self.user_embedding_dim = hp.user_embedding_dim
        self.item_embedding_dim = hp.item_embedding_dim
        self.cate_embedding_dim = hp.cate_embedding_dim
        G = self.train_num_ngs + 1
        # Execution knobs that are not reference hparams (hparam if present, else environment):
        #   math_mode  1 (default): large GEMMs on tcgen05 tensor cores (split-bf16, fp32 accumulate) and
        #              the recurrences on mma.sync -- the path bench.py measures; 0: fp32 CUDA cores.
        #   strict_clip  run every training step ungrouped so tf.clip_by_norm sees TF's un-deduplicated
        #              IndexedSlices (base_model.py:289-297) even when the clip is active; default off:
        #              shared-history execution, identical to TF whenever no table norm exceeds
        #              max_grad_norm (the engine counts the steps where one does and warns).
        def knob(name, env, default):
            v = getattr(hp, name, None) if name in hp else None
            return int(os.environ.get(env, default)) if v is None else int(v)
        self.math_mode = knob("math_mode", "CLSR_MATH_MODE", 1)
        self.strict_clip = bool(knob("strict_clip", "CLSR_STRICT_CLIP", 0))
        #   gpu_batches  (default 1) fit / run_eval / run_weighted_eval build their batches on the GPU from a
        #              device-resident copy of the parsed file (padding, x(1+num_ngs) grouping, in-batch negative
        #              sampling: sequential_iterator.py:519-704) instead of assembling host arrays per step.
        self.gpu_batches = bool(knob("gpu_batches", "CLSR_GPU_BATCHES", 1))
        self._clip_warned = False
        self.engine = Engine(
            self.item_vocab_length, self.cate_vocab_length, self.user_vocab_length,
            max_rows=hp.batch_size * G, seq_len=hp.max_seq_length, item_dim=hp.item_embedding_dim,
            cate_dim=hp.cate_embedding_dim, user_dim=hp.user_embedding_dim, hidden=hp.hidden_size,
            att_sizes=tuple(hp.att_fcn_layer_sizes), layer_sizes=tuple(hp.layer_sizes), train_group=G,
            embed_l2=hp.embed_l2, layer_l2=hp.layer_l2, contrastive_loss=hp.contrastive_loss,
            triplet_margin=hp.triplet_margin, contrastive_weight=hp.contrastive_loss_weight,
            discrepancy_weight=hp.discrepancy_loss_weight,
            contrastive_len_threshold=hp.contrastive_length_threshold,
            contrastive_recent_k=hp.contrastive_recent_k, optimizer=hp.optimizer,
            learning_rate=hp.learning_rate, clip_norm=bool(hp.is_clip_norm), max_grad_norm=float(hp.max_grad_norm),
            math_mode=self.math_mode, **self._variant())
        if hp.init_method != "tnormal":
            raise NotImplementedError("init_method=%s (only tnormal) on the B200 CLSR path" % hp.init_method)
        self.engine.set_params(P.init_params(
            self.item_vocab_length, self.cate_vocab_length, self.user_vocab_length, hp.item_embedding_dim,
            hp.cate_embedding_dim, hp.user_embedding_dim, hp.hidden_size, hp.att_fcn_layer_sizes, hp.layer_sizes,
            seed=self.seed, init_value=hp.init_value,
            **{k: v for k, v in self._variant().items() if k != "manual_alpha_value"}))

    def _variant(self):
        """Graph variants of the reference's _build_seq_graph (clsr.py:159-274) selected by hparams."""
        hp = self.hparams
        return dict(interest_evolve=bool(hp.interest_evolve), predict_long_short=bool(hp.predict_long_short),
                    manual_alpha=bool(hp.manual_alpha),
                    manual_alpha_value=float(hp.manual_alpha_value if "manual_alpha_value" in hp else 0.5),
                    sequential_model=str(hp.sequential_model))

    # ---- variables <-> checkpoints ---------------------------------------------------------------
    def _export_variables(self):
        shapes = self.engine.shapes()
        out = {}
        for name, arr in self.engine.get_params().items():
            out[name] = arr.reshape(shapes[name]) if name in shapes else arr
        return out

    def _import_variables(self, tensors):
        self.engine.set_params(tensors, strict=True)

    # ---- feed handling ---------------------------------------------------------------------------
    def _arrays(self, feed_dict, need_labels):
        it = self.iterator
        raw = {"users": feed_dict[it.users], "items": feed_dict[it.items], "cates": feed_dict[it.cates],
               "item_history": feed_dict[it.item_history], "item_cate_history": feed_dict[it.item_cate_history],
               "mask": feed_dict[it.mask], "time_from_first_action": feed_dict[it.time_from_first_action],
               "time_to_now": feed_dict[it.time_to_now], "labels": feed_dict.get(it.labels)}
        feed = normalize_feed(raw, need_labels)
        group = feed_dict.get(GROUP_KEY)
        if group is None:  # foreign feed: share work only if the rows of a group really are identical
            group = detect_group(feed, self.train_num_ngs + 1)
        return feed, int(group)

    def train(self, sess, feed_dict):
        """One optimisation step.  Returns the reference's 8-item fetch list
        [update, update_ops, loss, data_loss, regular_loss, contrastive_loss, discrepancy_loss, summary]."""
        feed_dict[self.layer_keeps] = self.keep_prob_train
        feed_dict[self.embedding_keeps] = self.embedding_keep_prob_train
        feed_dict[self.is_train_stage] = True
        if DEVICE_BATCH_KEY in feed_dict:
            ds, lines, num_ngs, seed = feed_dict[DEVICE_BATCH_KEY]
            self.engine.build_batch(ds, lines, num_ngs, seed)
            group = num_ngs + 1
            out = self.engine.train_step_staged()
        else:
            feed, group = self._arrays(feed_dict, True)
            if self.strict_clip and self.hparams.is_clip_norm:
                group = 1
            out = self.engine.train_step(feed, group=group, normalized=True)
        if group > 1 and self.hparams.is_clip_norm and not self._clip_warned:
            norms = self.engine.last_table_grad_norms
            if max(norms) > float(self.hparams.max_grad_norm):
                self._clip_warned = True
                warnings.warn("a table gradient norm (%.3g) exceeded max_grad_norm=%g in a shared-history step: the "
                              "clip was taken over group-summed slices and differs from TF's; set CLSR_STRICT_CLIP=1 "
                              "(or hparam strict_clip) for TF-exact clipping" % (max(norms), self.hparams.max_grad_norm))
        summary = dict(out) if self.hparams.write_tfevents else None
        return [None, [], out["loss"], out["data_loss"], out["regular_loss"], out["contrastive_loss"],
                out["discrepancy_loss"], summary]

    def batch_train(self, file_iterator, train_sess):
        step = 0
        epoch_loss = 0
        for batch_data_input in file_iterator:
            if batch_data_input:
                (_, _, step_loss, step_data_loss, _, _, _, summary) = self.train(train_sess, batch_data_input)
                if self.hparams.write_tfevents and self.hparams.SUMMARIES_DIR:
                    self.writer.add_summary(summary, step)
                epoch_loss += step_loss
                step += 1
                if step % self.hparams.show_step == 0:
                    print("step {0:d} , total_loss: {1:.4f}, data_loss: {2:.4f}".format(step, step_loss, step_data_loss))
        return epoch_loss

    def _device_loops(self):
        """Context manager: the iterator yields device-batch tokens inside the model's own loops."""
        model = self

        class _Ctx:
            def __enter__(self):
                if model.gpu_batches and not (model.strict_clip and model.hparams.is_clip_norm):
                    model.iterator.device_engine = model.engine

            def __exit__(self, *a):
                model.iterator.device_engine = None
                return False
        return _Ctx()

    def fit(self, train_file, valid_file, valid_num_ngs, eval_metric="group_auc"):
        with self._device_loops():
            return super().fit(train_file, valid_file, valid_num_ngs, eval_metric)

    def _predict(self, feed_dict, with_alpha=False):
        feed_dict[self.layer_keeps] = self.keep_prob_test
        feed_dict[self.embedding_keeps] = self.embedding_keep_prob_test
        feed_dict[self.is_train_stage] = False
        if DEVICE_BATCH_KEY in feed_dict:
            ds, lines, num_ngs, seed = feed_dict[DEVICE_BATCH_KEY]
            self.engine.build_batch(ds, lines, num_ngs, seed)
            p, a, u, y = self.engine.predict_staged(with_alpha=with_alpha)
            self.engine.synchronize()
            feed_dict[self.iterator.labels] = y.cpu().numpy().reshape(-1, 1)
            feed = {"users": u.cpu().numpy() if u is not None else None}
            return feed, p.cpu().numpy().reshape(-1, 1), (a.cpu().numpy().reshape(-1, 1) if with_alpha else None)
        feed, group = self._arrays(feed_dict, False)
        if GROUP_KEY in feed_dict and feed_dict[GROUP_KEY] == 1:
            group = 1
        pred, alpha = self.engine.predict(feed, group=group, normalized=True, with_alpha=with_alpha)
        return feed, pred.reshape(-1, 1), (alpha.reshape(-1, 1) if with_alpha else None)

    def eval(self, sess, feed_dict):
        _, pred, _ = self._predict(feed_dict)
        return [pred, np.asarray(feed_dict[self.iterator.labels])]

    def eval_with_user(self, sess, feed_dict):
        feed, pred, _ = self._predict(feed_dict)
        return [feed["users"], pred, np.asarray(feed_dict[self.iterator.labels])]

    def eval_with_user_and_alpha(self, sess, feed_dict):
        feed, pred, alpha = self._predict(feed_dict, with_alpha=True)
        return [feed["users"], pred, np.asarray(feed_dict[self.iterator.labels]), alpha]

    def infer(self, sess, feed_dict):
        _, pred, _ = self._predict(feed_dict)
        return [pred]

    # ---- evaluation with device-resident predictions and GPU metrics -------------------------------------
    def _eval_on_device(self, filename, num_ngs, weighted, calc_mean_alpha=False, manual_alpha=False):
        """run_eval / run_weighted_eval (sequential_base_model.py:204-292) without the per-batch device -> host
        copies and the host metric loops: predictions of the whole file stay on the GPU and
        clsr_eval_metrics_compute returns auc / logloss / mean_mrr / ndcg@k / hit@k / group_auc / wauc."""
        import torch
        from clsr_b200 import metrics_gpu as MG
        from reco_utils.recommender.deeprec.deeprec_utils import cal_mean_alpha_metric
        preds, alphas, users, labels = [], [], [], []
        it = self.iterator
        was = it.device_engine
        if self.gpu_batches:
            it.device_engine = self.engine
        try:
            batches = list(it.load_data_from_file(filename, min_seq_length=self.min_seq_length, batch_num_ngs=0))
        finally:
            it.device_engine = was
        for feed_dict in batches:
            if not feed_dict:
                continue
            if DEVICE_BATCH_KEY in feed_dict:   # batch built on the device; users / labels never visit the host
                ds, lines, _, seed = feed_dict[DEVICE_BATCH_KEY]
                self.engine.build_batch(ds, lines, 0, seed)
                p, a, u, y = self.engine.predict_staged(with_alpha=calc_mean_alpha)
                preds.append(p)
                if calc_mean_alpha:
                    alphas.append(a)
                users.append(u)
                labels.append(y)
                continue
            feed_dict[self.layer_keeps] = self.keep_prob_test
            feed_dict[self.embedding_keeps] = self.embedding_keep_prob_test
            feed_dict[self.is_train_stage] = False
            feed, _ = self._arrays(feed_dict, False)
            p, a = self.engine.predict_device(feed, group=1, normalized=True, with_alpha=calc_mean_alpha)
            preds.append(p)
            if calc_mean_alpha:
                alphas.append(a)
            users.append(feed["users"].reshape(-1))
            labels.append(np.asarray(feed_dict[it.labels], np.float32).reshape(-1))
        preds = torch.cat(preds)
        if isinstance(users[0], torch.Tensor):
            users, labels = torch.cat(users), torch.cat(labels)
        else:
            users, labels = np.concatenate(users), np.concatenate(labels)
        hp = self.hparams
        res = MG.metric_dict(preds, labels, users, num_ngs + 1, hp.metrics, hp.pairwise_metrics,
                             hp.weighted_metrics if weighted else None)
        if calc_mean_alpha:
            al = torch.cat(alphas).cpu().numpy()
            lab = labels.cpu().numpy() if isinstance(labels, torch.Tensor) else labels
            res.update(cal_mean_alpha_metric(al[0] if manual_alpha else al, lab))
        return res

    def run_eval(self, filename, num_ngs):
        from clsr_b200 import metrics_gpu as MG
        hp = self.hparams
        if MG.supported(hp.metrics, hp.pairwise_metrics, None):
            return self._eval_on_device(filename, num_ngs, weighted=False)
        return super().run_eval(filename, num_ngs)

    def run_weighted_eval(self, filename, num_ngs, calc_mean_alpha=False, manual_alpha=False):
        from clsr_b200 import metrics_gpu as MG
        hp = self.hparams
        if MG.supported(hp.metrics, hp.pairwise_metrics, hp.weighted_metrics):
            return self._eval_on_device(filename, num_ngs, True, calc_mean_alpha, manual_alpha)
        return super().run_weighted_eval(filename, num_ngs, calc_mean_alpha, manual_alpha)
