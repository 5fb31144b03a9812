"""Timed CPU arm -- TEST/BENCH INFRASTRUCTURE (see oracle/clsr_oracle.py header).

Runs the reference's training step (CLSRModel.train, clsr.py:383-408) on host cores as the
restated fp32 PyTorch-CPU graph of oracle/clsr_oracle.py -- TensorFlow 1.15 itself is not
installable here -- at the TF graph's op granularity: unfused ops, one time step per loop
iteration, the x(1+num_ngs)-replicated batch, per-variable clip and the non-lazy sparse Adam
that sweeps every table row.  Unlike ``clsr_oracle.train_step`` the
variables stay resident as fp32 tensors and are updated in place, so a step costs what the
reference's step costs rather than a table-sized dtype conversion.
"""
import math
import time

import numpy as np
import torch

from . import clsr_oracle as O


class CpuTrainer:
    def __init__(self, params, cfg, threads=None):
        if threads:
            torch.set_num_threads(int(threads))
        self.threads = torch.get_num_threads()
        self.cfg = cfg
        self.p = {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))).float()
                  for k, v in params.items()}
        self.m = {}
        self.v = {}
        self.step_no = 0
        self.opt_seconds = 0.0   # time spent in the table optimizer (sparse Adam / lazy Adam)

    def _adam_dense(self, name, g, lr_t):
        c = self.cfg
        m = self.m.setdefault(name, torch.zeros_like(self.p[name]))
        v = self.v.setdefault(name, torch.zeros_like(self.p[name]))
        m.mul_(c.beta1).add_(g, alpha=1 - c.beta1)
        v.mul_(c.beta2).addcmul_(g, g, value=1 - c.beta2)
        self.p[name].addcdiv_(m, v.sqrt().add_(c.adam_eps), value=-lr_t)

    def step(self, batch):
        c = self.cfg
        p = {}
        for k, t in self.p.items():
            if not k.startswith(O.EMB) and "moving_" not in k:
                t = t.detach().requires_grad_(True)
            p[k] = t
        leaves = {}
        out = O.forward(p, batch, c, True, torch.float32, leaves)
        L = O.losses(out, p, batch, c, torch.float32)
        L["loss"].backward()
        self.step_no += 1
        lr_t = c.learning_rate * math.sqrt(1 - c.beta2 ** self.step_no) / (1 - c.beta1 ** self.step_no)
        with torch.no_grad():
            for k, t in p.items():
                if t.requires_grad:
                    self._adam_dense(k, O._clip(t.grad, c), lr_t)
            t_opt = time.perf_counter()
            for tab in O.TABLES:
                name = O.EMB + tab
                idx = torch.cat([ix for (tb, ix, rows) in leaves.values() if tb == tab])
                val = torch.cat([rows.grad.reshape(-1, rows.shape[-1]) for (tb, ix, rows) in leaves.values()
                                 if tb == tab])
                val = O._clip(val, c)
                uniq, inv = torch.unique(idx, return_inverse=True)
                gs = torch.zeros(len(uniq), val.shape[1]).index_add_(0, inv, val)
                var = self.p[name]
                m = self.m.setdefault(name, torch.zeros_like(var))
                v = self.v.setdefault(name, torch.zeros_like(var))
                if c.optimizer == "lazyadam":
                    mu = m[uniq] * c.beta1 + (1 - c.beta1) * gs
                    vu = v[uniq] * c.beta2 + (1 - c.beta2) * gs * gs
                    m[uniq] = mu
                    v[uniq] = vu
                    var[uniq] = var[uniq] - lr_t * mu / (vu.sqrt() + c.adam_eps)
                else:
                    m.mul_(c.beta1)
                    m.index_add_(0, uniq, gs, alpha=1 - c.beta1)
                    v.mul_(c.beta2)
                    v.index_add_(0, uniq, gs * gs, alpha=1 - c.beta2)
                    var.addcdiv_(m, v.sqrt().add_(c.adam_eps), value=-lr_t)
            self.opt_seconds += time.perf_counter() - t_opt
            for prefix, (mean, var_b) in out["bn_stats"].items():
                for suffix, b in (("moving_mean", mean), ("moving_variance", var_b)):
                    mv = self.p[prefix + suffix]
                    mv.sub_((mv - b) * (1 - c.bn_momentum))
        return {k: float(v.detach()) for k, v in L.items()}


def time_steps(trainer, batches, steps, warmup):
    """Seconds per step over ``steps`` timed steps after ``warmup`` untimed ones."""
    n = len(batches)
    for i in range(warmup):
        trainer.step(batches[i % n])
    t0 = time.perf_counter()
    for i in range(steps):
        trainer.step(batches[(warmup + i) % n])
    return (time.perf_counter() - t0) / max(steps, 1)
