"""Independent PyTorch-CPU restatement of the same graph.  TEST INFRASTRUCTURE ONLY.

Used (a) to pin `oracle/icl_oracle.py` (autograd gradients vs the hand-written NumPy BPTT),
(b) as the timed "port" CPU baseline in bench.py (`cpu_baseline.kind == "port"`): the
reference's own TF-1.x CPU path cannot run here (no TensorFlow / Python 2), so the same
per-timestep BasicLSTMCell loop + autograd + TF-style Adam is timed instead.

Follows nn_utils/core.py:271-332 (BiLSTM), :354-440 (batch input), :146-268 (FFW, softmax, CE),
:74-106 (clip + Adam) -- see oracle/icl_oracle.py for the TF semantics spelled out.
"""
import numpy as np
import torch

from . import icl_oracle as O


def to_torch(params, dtype=torch.float64, requires_grad=True):
    return {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=requires_grad) for k, v in params.items()}


def _dir(x, lens, kernel, bias, reverse):
    S, T, E = x.shape
    H = kernel.shape[1] // 4
    h = x.new_zeros(S, H)
    c = x.new_zeros(S, H)
    rows = torch.arange(S)
    outs = []          # (pos, act, h_new)
    tmax = int(lens.max())
    for k in range(tmax):
        act = k < lens
        pos = torch.where(act, (lens - 1 - k) if reverse else torch.full_like(lens, k), torch.zeros_like(lens))
        xt = x[rows, pos]
        z = torch.cat([xt, h], 1) @ kernel + bias
        i, j, f, o = z.chunk(4, 1)
        c_new = c * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)
        h_new = torch.tanh(c_new) * torch.sigmoid(o)
        m = act[:, None]
        c = torch.where(m, c_new, c)
        h = torch.where(m, h_new, h)
        outs.append((pos, act, h_new))
    out = x.new_zeros(S, T, H)
    # functional scatter (keeps autograd happy)
    for pos, act, h_new in outs:
        onehot = torch.zeros(S, T, dtype=x.dtype)
        onehot[rows[act], pos[act]] = 1.0
        out = out + onehot[:, :, None] * h_new[:, None, :]
    return out


def model_loss(tp, cfg, sentences, seq_lengths, head_batches, keep_in=1.0, keep=1.0, masks=None,
               dtype=torch.float64):
    x = torch.tensor(np.asarray(sentences), dtype=dtype)
    lens = torch.tensor(np.asarray(seq_lengths).astype(np.int64))
    if cfg.get("data_norm", False):
        ss = (x * x).sum(-1, keepdim=True)
        x = x / torch.sqrt(torch.clamp(ss, min=1e-12))
    outs = {}
    for d in ("fw", "bw"):
        kn, bn = O.lstm_names(d)
        xin = x
        if masks is not None and ("in_" + d) in masks:
            xin = x / keep_in * torch.tensor(masks["in_" + d], dtype=dtype)
        o = _dir(xin, lens, tp[kn], tp[bn], d == "bw")
        if masks is not None and ("out_" + d) in masks:
            o = o / keep * torch.tensor(masks["out_" + d], dtype=dtype)
        outs[d] = o
    stacked = torch.stack([outs["fw"], outs["bw"]])
    total = 0.0
    probas = []
    for hi, (hc, hb) in enumerate(zip(cfg["heads"], head_batches)):
        plan = O.slot_plan(hc["task"], hc["encoding_scheme"], "b_feats" in hb)
        cols = []
        for name in plan:
            if name in O.DENSE:
                cols.append(torch.tensor(np.asarray(hb[name]), dtype=dtype))
            else:
                idx = torch.tensor(np.asarray(hb[name]).astype(np.int64))
                cols.append(stacked[idx[:, 0], idx[:, 1], idx[:, 2]])
        a = torch.cat(cols, 1)
        names = O.head_names(hc.get("scope", ""), hc["n_layers"])
        for k in range(hc["n_layers"]):
            z = a @ tp[names[k][0]] + tp[names[k][1]]
            act = hc["activation"]
            if act == "sigmoid":
                z = torch.sigmoid(z)
            elif act == "tanh":
                z = torch.tanh(z)
            elif act == "relu":
                z = torch.relu(z)
            elif act == "leaky_relu":
                z = torch.maximum(z, 0.01 * z)
            if masks is not None and "heads" in masks:
                z = z / keep * torch.tensor(masks["heads"][hi][k], dtype=dtype)
            a = z
        logits = a @ tp[names[-1][0]] + tp[names[-1][1]]
        y = torch.tensor(np.asarray(hb["labels"]), dtype=dtype)
        ce = -(y * torch.log_softmax(logits, 1)).sum(1)
        loss = ce.mean() if hc.get("weighted_classes", False) else ce.sum()
        total = total + loss
        probas.append(torch.softmax(logits, 1))
    return total, probas, outs


def grads_via_autograd(params, cfg, sentences, seq_lengths, head_batches, keep_in=1.0, keep=1.0, masks=None):
    tp = to_torch(params)
    loss, probas, outs = model_loss(tp, cfg, sentences, seq_lengths, head_batches, keep_in, keep, masks)
    loss.backward()
    g = {k: (v.grad.numpy() if v.grad is not None else np.zeros(v.shape)) for k, v in tp.items()}
    return float(loss), g, [p.detach().numpy() for p in probas], {k: v.detach().numpy() for k, v in outs.items()}


class TFAdam(object):
    """tf.train.AdamOptimizer update rule (epsilon outside the bias correction) + clip_by_global_norm."""

    def __init__(self, tparams, lr, eps, clip_norm, b1=0.9, b2=0.999):
        self.p, self.lr, self.eps, self.clip, self.b1, self.b2 = tparams, lr, eps, clip_norm, b1, b2
        self.t = 0
        self.m = {k: torch.zeros_like(v) for k, v in tparams.items()}
        self.v = {k: torch.zeros_like(v) for k, v in tparams.items()}

    @torch.no_grad()
    def step(self):
        gs = {k: v.grad for k, v in self.p.items() if v.grad is not None}
        if self.clip is not None:
            gn = torch.sqrt(sum((g.double() ** 2).sum() for g in gs.values()))
            s = self.clip / max(float(gn), self.clip)
            for g in gs.values():
                g.mul_(s)
        self.t += 1
        lr_t = self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for k, g in gs.items():
            self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            self.p[k].sub_(lr_t * self.m[k] / (self.v[k].sqrt() + self.eps))
            self.p[k].grad = None


def fast_dir(x, lens_sorted_desc, kernel, bias, reverse, keep_out_mask=None):
    """Throughput-oriented variant for the CPU baseline: sequences pre-sorted by length (desc) so each
    step works on the active prefix only (what dynamic_rnn's sequence_length skipping buys TF)."""
    S, T, E = x.shape
    H = kernel.shape[1] // 4
    h = x.new_zeros(S, H)
    c = x.new_zeros(S, H)
    lens = lens_sorted_desc
    hs = []
    tmax = int(lens[0])
    rows = torch.arange(S)
    for k in range(tmax):
        n = int((lens > k).sum())
        pos = (lens[:n] - 1 - k) if reverse else torch.full((n,), k, dtype=torch.long)
        xt = x[rows[:n], pos]
        z = torch.cat([xt, h[:n]], 1) @ kernel + bias
        i, j, f, o = z.chunk(4, 1)
        c_new = c[:n] * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)
        h_new = torch.tanh(c_new) * torch.sigmoid(o)
        c = torch.cat([c_new, c[n:]], 0)
        h = torch.cat([h_new, h[n:]], 0)
        hs.append((n, pos, h_new))
    return hs
