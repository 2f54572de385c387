"""CPU oracle for the BiLSTM + mention-span-head hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in NumPy, the arithmetic that the reference executes inside
TensorFlow 1.x for the path `nn_utils/core.py` builds.  It is the *checker* for the
CUDA path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it; nothing under `imagecaptionlearn_py_b200/`
does (a test enforces that).

PARITY STATUS: **unpinned by the reference** -- the reference ships no tests, golden
vectors or fixtures, and TensorFlow 1.x / Python 2 are absent here, so the reference
cannot be executed.  The oracle is pinned instead by (1) an independent PyTorch
autograd restatement (`oracle/torch_check.py`), (2) a mapping onto `torch.nn.LSTM`,
(3) finite differences, all in `tests/test_oracle.py`, and (4) committed known-answer
vectors under `tests/golden/`.

Semantics followed (reference file:line -> function here)
  nn_utils/core.py:121-143  get_widths                -> get_widths
  nn_utils/core.py:289-290  tf.nn.l2_normalize(dim=2) -> l2_normalize
  nn_utils/core.py:305-329  BasicLSTMCell + DropoutWrapper + bidirectional_dynamic_rnn
                                                        -> bilstm_forward / bilstm_backward
  nn_utils/core.py:335-351  gather_nd on (fw,bw)      -> gather_spans
  nn_utils/core.py:354-440  setup_batch_inputs        -> slot_plan / build_batch_input
  nn_utils/core.py:146-199  setup_ffw                 -> head_forward
  nn_utils/core.py:202-232  apply_softmax             -> head_forward
  nn_utils/core.py:235-268  setup_cross_entropy_loss  -> head_forward (sum; "weighted" = as executed: mean)
  nn_utils/core.py:509-511  argmax / accuracy         -> head_forward
  nn_utils/core.py:74-106   clip_by_global_norm + AdamOptimizer -> clip_and_adam
  icl_multitask_lstm.py:248-254 simple_joint loss     -> model_forward over several heads

TF-1.x semantics written out (third-party, un-vendored, version unpinned ~1.3/1.4):
  BasicLSTMCell: z=[x,h]@kernel+bias; i,j,f,o=split(z,4); c'=c*sig(f+1)+sig(i)*tanh(j); h'=tanh(c')*sig(o)
  DropoutWrapper: mask on cell *input* and emitted *output* only; state never dropped
  tf.nn.dropout: x/keep * floor(keep+U[0,1))
  dynamic_rnn(sequence_length): t>=len -> output 0, state copied through
  bidirectional: bw = reverse_sequence(x) -> cell -> reverse_sequence(out)
  AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps)
  clip_by_global_norm: g * clip/max(||g||, clip)
"""
import numpy as np

LSTM_SCOPE = "bidirectional_lstm/bidirectional_rnn"


def lstm_names(direction):
    base = "%s/%s/basic_lstm_cell/" % (LSTM_SCOPE, direction)
    return base + "kernel", base + "bias"


def scoped(scope, name):
    """Name of head variable `name` ('hdn_1/Variable', ...) created under the task scope `scope` (see head_names)."""
    return (scope + "/" + scope + "/" + name) if scope else name


def head_names(scope, n_layers):
    """TF variable names in creation order: hdn_k/Variable (W), hdn_k/Variable_1 (b), softmax/...  Under a task scope
    (icl_multitask_lstm.py:62) the prefix is DOUBLED: core.py:166-172 opens variable_scope("<task>/hdn_k") inside
    variable_scope("<task>"), and TensorFlow nests it -> "<task>/<task>/hdn_k/Variable"."""
    pre = (scope + "/" + scope + "/") if scope else ""
    names = []
    for k in range(1, n_layers + 1):
        names.append((pre + "hdn_%d/Variable" % k, pre + "hdn_%d/Variable_1" % k))
    names.append((pre + "softmax/Variable", pre + "softmax/Variable_1"))
    return names


# ----------------------------------------------------------------------------- helpers
def get_widths(start_width, depth, end_width=None):
    """core.py:121-143 -- note depth d gives d+1 widths when end_width is None."""
    w = [int(start_width)]
    if end_width is None:
        for d in range(1, depth + 1):
            w.append(int(w[d - 1] / 2))
    else:
        for d in range(1, depth):
            w.append(int(max(w[d - 1] - end_width, 0) / 2 + end_width))
        w.append(int(end_width))
    return w


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def l2_normalize(x, eps=1e-12):
    """tf.nn.l2_normalize(x, dim=2): x * rsqrt(max(sum(x^2), eps))."""
    ss = np.sum(x * x, axis=-1, keepdims=True)
    return x / np.sqrt(np.maximum(ss, eps))


def activate(z, act):
    if act == "sigmoid":
        return sigmoid(z)
    if act == "tanh":
        return np.tanh(z)
    if act == "relu":
        return np.maximum(z, 0.0)
    if act == "leaky_relu":
        return np.maximum(z, 0.01 * z)
    return z


def activate_grad(z, a, act):
    if act == "sigmoid":
        return a * (1.0 - a)
    if act == "tanh":
        return 1.0 - a * a
    if act == "relu":
        return (z > 0).astype(z.dtype)
    if act == "leaky_relu":
        return np.where(z > 0, 1.0, 0.01).astype(z.dtype)
    return np.ones_like(z)


# ----------------------------------------------------------------------------- BiLSTM
def _dir_forward(x, lens, kernel, bias, reverse):
    """One direction of dynamic_rnn over already input-dropped x [S,T,E].

    Returns h_all [S,T,H] (pre output-dropout; zero for t>=len) in ORIGINAL token order
    and a cache.  For reverse, step k consumes token len-1-k (reverse_sequence semantics).
    """
    S, T, E = x.shape
    H = kernel.shape[1] // 4
    dt = x.dtype
    h = np.zeros((S, H), dt)
    c = np.zeros((S, H), dt)
    h_all = np.zeros((S, T, H), dt)
    cache = []
    lens = np.asarray(lens).astype(np.int64)
    rows = np.arange(S)
    for k in range(int(lens.max()) if S else 0):
        act = k < lens                       # sequences still running
        pos = np.where(act, (lens - 1 - k) if reverse else k, 0)
        xt = x[rows, pos]
        z = np.concatenate([xt, h], 1) @ kernel + bias
        i, j, f, o = np.split(z, 4, axis=1)
        si, sf, so, tj = sigmoid(i), sigmoid(f + 1.0), sigmoid(o), np.tanh(j)
        c_new = c * sf + si * tj
        tc = np.tanh(c_new)
        h_new = tc * so
        cache.append((act.copy(), pos.copy(), xt, h.copy(), c.copy(), si, sf, so, tj, tc))
        m = act[:, None]
        c = np.where(m, c_new, c)
        h = np.where(m, h_new, h)
        h_all[rows[act], pos[act]] = h_new[act]
    return h_all, cache


def _dir_backward(d_h_all, cache, kernel, E):
    """BPTT for one direction.  d_h_all [S,T,H] = dL/d(pre-output-dropout h).  No dX."""
    S, T, H = d_h_all.shape
    dK = np.zeros_like(kernel)
    db = np.zeros(kernel.shape[1], kernel.dtype)
    dh = np.zeros((S, H), kernel.dtype)
    dc = np.zeros((S, H), kernel.dtype)
    rows = np.arange(S)
    Wh = kernel[E:]
    for (act, pos, xt, h_prev, c_prev, si, sf, so, tj, tc) in reversed(cache):
        m = act[:, None]
        dh_t = np.where(m, dh + d_h_all[rows, pos], 0.0)
        do = dh_t * tc * so * (1 - so)
        dc_t = np.where(m, dc, 0.0) + dh_t * so * (1 - tc * tc)
        di = dc_t * tj * si * (1 - si)
        dj = dc_t * si * (1 - tj * tj)
        df = dc_t * c_prev * sf * (1 - sf)
        dz = np.concatenate([di, dj, df, do], 1)          # already zero on inactive rows
        dK += np.concatenate([xt, h_prev], 1).T @ dz
        db += dz.sum(0)
        dh = np.where(m, dz @ Wh.T, dh)
        dc = np.where(m, dc_t * sf, dc)
    return dK, db


def bilstm_forward(params, x, lens, data_norm=False, keep_in=1.0, keep_out=1.0, masks=None):
    """core.py:271-332.  masks: dict of 0/1 arrays 'in_fw','in_bw' [S,T,E], 'out_fw','out_bw' [S,T,H]
    indexed by ORIGINAL token position (each direction draws its own masks).
    Returns (out_fw, out_bw) as emitted (after output dropout) and a cache."""
    x = np.asarray(x)
    if data_norm:
        x = l2_normalize(x)
    E = x.shape[2]
    outs, caches, hs = {}, {}, {}
    for d in ("fw", "bw"):
        kn, bn = lstm_names(d)
        xin = x
        if masks is not None and ("in_" + d) in masks:
            xin = x / keep_in * masks["in_" + d]
        h_all, cache = _dir_forward(xin, lens, params[kn], params[bn], reverse=(d == "bw"))
        hs[d] = h_all
        caches[d] = cache
        out = h_all
        if masks is not None and ("out_" + d) in masks:
            out = h_all / keep_out * masks["out_" + d]
        outs[d] = out
    return outs["fw"], outs["bw"], dict(caches=caches, E=E, keep_out=keep_out, masks=masks)


def bilstm_backward(params, cache, d_out_fw, d_out_bw):
    grads = {}
    for d, d_out in (("fw", d_out_fw), ("bw", d_out_bw)):
        kn, bn = lstm_names(d)
        masks = cache["masks"]
        if masks is not None and ("out_" + d) in masks:
            d_out = d_out / cache["keep_out"] * masks["out_" + d]
        grads[kn], grads[bn] = _dir_backward(d_out, cache["caches"][d], params[kn], cache["E"])
    return grads


# ----------------------------------------------------------------------------- heads
def slot_plan(task, encoding_scheme, with_b_feats=False):
    """Column blocks of batch_input in order (core.py:377-433).  Entries are index-matrix
    names (gathered LSTM rows) or dense feature names."""
    plan = ["first_i_bw", "last_i_fw"]
    if encoding_scheme == "first_last_sentence":
        plan += ["sent_last_i_fw", "sent_first_i_bw"]
    elif encoding_scheme == "first_last_mention":
        plan += ["first_i_fw", "last_i_bw"]
    if "rel" in task:
        plan += ["first_j_bw", "last_j_fw", "ij_feats"]
        if encoding_scheme == "first_last_sentence" and task == "rel_cross":
            plan += ["sent_last_j_fw", "sent_first_j_bw"]
        elif encoding_scheme == "first_last_mention":
            plan += ["first_j_fw", "last_j_bw"]
    elif task in ("nonvis", "card", "affinity"):
        plan += ["m_feats"]
        if task == "affinity":
            plan += ["box_embeddings"]
            if with_b_feats:
                plan += ["b_feats"]
    return plan


DENSE = ("m_feats", "ij_feats", "box_embeddings", "b_feats")


def gather_spans(out_fw, out_bw, idx):
    """tf.gather_nd((out_fw,out_bw), idx) with idx[b]=[dir,sent,word] (core.py:348-350)."""
    idx = np.asarray(idx).astype(np.int64)
    stacked = np.stack([out_fw, out_bw])
    return stacked[idx[:, 0], idx[:, 1], idx[:, 2]]


def build_batch_input(out_fw, out_bw, batch, task, encoding_scheme):
    plan = slot_plan(task, encoding_scheme, "b_feats" in batch)
    cols = []
    for name in plan:
        if name in DENSE:
            cols.append(np.asarray(batch[name]).astype(out_fw.dtype))
        else:
            cols.append(gather_spans(out_fw, out_bw, batch[name]))
    return np.concatenate(cols, 1), plan


def head_forward(params, scope, batch_input, y, n_layers, act, keep=1.0, masks=None,
                 weighted_classes=False):
    """setup_ffw + apply_softmax + CE + metrics.  masks: list of 0/1 [B,w_k] per hidden layer."""
    names = head_names(scope, n_layers)
    a = batch_input
    layers = []
    for k in range(n_layers):
        W, b = params[names[k][0]], params[names[k][1]]
        z = a @ W + b
        h = activate(z, act)
        out = h
        if masks is not None:
            out = h / keep * masks[k]
        layers.append((a, z, h))
        a = out
    W, b = params[names[-1][0]], params[names[-1][1]]
    logits = a @ W + b                        # "+ epsilon" is float32(4.9e-324) == 0 (core.py:223-229)
    mx = logits.max(1, keepdims=True)
    e = np.exp(logits - mx)
    proba = e / e.sum(1, keepdims=True)
    logp = (logits - mx) - np.log(e.sum(1, keepdims=True))
    res = dict(proba=proba, pred=proba.argmax(1).astype(np.int64), logits=logits)
    if y is not None:
        y = np.asarray(y).astype(logits.dtype)
        ce = -(y * logp).sum(1)
        # weighted_classes as EXECUTED under python2: 1/batch_size == 0 -> weights 1 ->
        # tf.losses SUM_BY_NONZERO_WEIGHTS -> mean CE (core.py:244-267)
        scale = (1.0 / len(ce)) if weighted_classes else 1.0
        res["loss"] = ce.sum() * scale
        res["accuracy"] = np.mean(res["pred"] == y.argmax(1))
        res["_bwd"] = (layers, a, proba, y, scale, names, act, keep, masks)
    return res


def head_backward(params, res, act_grad_override=None, loss_weight=1.0):
    """act_grad_override: optional list (per hidden layer) of arrays replacing activate_grad -- used by the parity
    tests to evaluate a piecewise-linear activation on the same side of its kink as the device did."""
    layers, a_last, proba, y, scale, names, act, keep, masks = res["_bwd"]
    grads = {}
    dlog = (proba * y.sum(1, keepdims=True) - y) * scale * loss_weight     # loss_weight = d joint / d loss of this head
    W = params[names[-1][0]]
    grads[names[-1][0]] = a_last.T @ dlog
    grads[names[-1][1]] = dlog.sum(0, keepdims=True)
    da = dlog @ W.T
    for k in reversed(range(len(layers))):
        a_in, z, h = layers[k]
        if masks is not None:
            da = da / keep * masks[k]
        dz = da * (activate_grad(z, h, act) if act_grad_override is None else act_grad_override[k])
        grads[names[k][0]] = a_in.T @ dz
        grads[names[k][1]] = dz.sum(0, keepdims=True)
        da = dz @ params[names[k][0]].T
    return grads, da                           # da = d batch_input


def scatter_spans(d_batch_input, plan, batch, shape_fw, H):
    """Backward of gather_nd + concat: scatter-add into d_out_fw/d_out_bw."""
    d_fw = np.zeros(shape_fw, d_batch_input.dtype)
    d_bw = np.zeros(shape_fw, d_batch_input.dtype)
    col = 0
    for name in plan:
        if name in DENSE:
            col += np.asarray(batch[name]).shape[1]
            continue
        idx = np.asarray(batch[name]).astype(np.int64)
        blk = d_batch_input[:, col:col + H]
        for dval, tgt in ((0, d_fw), (1, d_bw)):
            sel = idx[:, 0] == dval
            np.add.at(tgt, (idx[sel, 1], idx[sel, 2]), blk[sel])
        col += H
    return d_fw, d_bw


# ----------------------------------------------------------------------------- whole model
def model_forward(params, cfg, sentences, seq_lengths, head_batches, keep_in=1.0, keep=1.0,
                  masks=None, with_labels=True):
    """cfg: dict(H, data_norm, heads=[dict(task, scope, encoding_scheme, n_layers, activation,
    weighted_classes)]).  head_batches[i] is the batch_tensors dict of head i (index matrices address
    rows of the shared `sentences`).  masks: dict(in_fw,in_bw,out_fw,out_bw, heads=[[..],..])."""
    out_fw, out_bw, lcache = bilstm_forward(params, sentences, seq_lengths, cfg.get("data_norm", False),
                                            keep_in, keep, masks)
    results, total = [], 0.0
    for hi, (hc, hb) in enumerate(zip(cfg["heads"], head_batches)):
        bi, plan = build_batch_input(out_fw, out_bw, hb, hc["task"], hc["encoding_scheme"])
        hm = None if masks is None or "heads" not in masks else masks["heads"][hi]
        r = head_forward(params, hc.get("scope", ""), bi, hb["labels"] if with_labels else None,
                         hc["n_layers"], hc["activation"], keep, hm, hc.get("weighted_classes", False))
        r["batch_input"], r["plan"] = bi, plan
        results.append(r)
        if with_labels:
            total = total + r["loss"]
    return dict(heads=results, loss=total, out_fw=out_fw, out_bw=out_bw, _lstm=lcache)


def model_backward(params, cfg, fwd, head_batches, act_grad_override=None, head_weights=None):
    """head_weights: d joint_loss / d loss_t per head (None = the plain sum of the task losses; the `weighted_joint` scheme of
    icl_multitask_lstm.py:248-255 passes the row sums of its trainable mixing matrix)."""
    grads = {}
    H = cfg["H"]
    d_fw = np.zeros_like(fwd["out_fw"])
    d_bw = np.zeros_like(fwd["out_bw"])
    for hi, (hc, hb, r) in enumerate(zip(cfg["heads"], head_batches, fwd["heads"])):
        g, d_bi = head_backward(params, r, None if act_grad_override is None else act_grad_override[hi],
                                1.0 if head_weights is None else head_weights[hi])
        grads.update(g)
        a, b = scatter_spans(d_bi, r["plan"], hb, d_fw.shape, H)
        d_fw += a
        d_bw += b
    grads.update(bilstm_backward(params, fwd["_lstm"], d_fw, d_bw))
    return grads


# ----------------------------------------------------------------------------- optimizer
def global_norm(grads):
    return np.sqrt(sum(float(np.sum(np.square(g.astype(np.float64)))) for g in grads.values()))


def clip_and_adam(params, grads, state, lr, eps, clip_norm, beta1=0.9, beta2=0.999):
    """core.py:94-103.  state: dict(t=int, m={}, v={}) updated in place; params updated in place."""
    if clip_norm is not None:
        gn = global_norm(grads)
        scale = clip_norm / max(gn, clip_norm)
        grads = {k: g * scale for k, g in grads.items()}
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    lr_t = lr * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    for k, g in grads.items():
        m = state.setdefault("m", {}).setdefault(k, np.zeros_like(params[k]))
        v = state.setdefault("v", {}).setdefault(k, np.zeros_like(params[k]))
        m[...] = beta1 * m + (1 - beta1) * g
        v[...] = beta2 * v + (1 - beta2) * g * g
        params[k] -= (lr_t * m / (np.sqrt(v) + eps)).astype(params[k].dtype)
    return grads


# ----------------------------------------------------------------------------- init
def init_params(rng, cfg, E, dtype=np.float64):
    """Glorot-uniform kernels / zero LSTM bias (TF default) and the reference's Xavier rule for the
    heads (core.py:32-36,59-63).  TF's Philox stream is not reproducible; parity tests inject these."""
    H = cfg["H"]
    p = {}
    for d in ("fw", "bw"):
        kn, bn = lstm_names(d)
        lim = np.sqrt(6.0 / (E + H + 4 * H))
        p[kn] = rng.uniform(-lim, lim, (E + H, 4 * H)).astype(dtype)
        p[bn] = np.zeros(4 * H, dtype)
    for hc in cfg["heads"]:
        dims = [hc["in_width"]] + list(hc["widths"]) + [hc["n_classes"]]
        names = head_names(hc.get("scope", ""), len(hc["widths"]))
        for k, (wn, bn) in enumerate(names):
            lw = np.sqrt(6.0 / (dims[k] + dims[k + 1]))
            lb = np.sqrt(6.0 / (1 + dims[k + 1]))
            p[wn] = rng.uniform(-lw, lw, (dims[k], dims[k + 1])).astype(dtype)
            p[bn] = rng.uniform(-lb, lb, (1, dims[k + 1])).astype(dtype)
    return p


def head_in_width(task, encoding_scheme, H, F, box_w=0, Fb=0):
    plan = slot_plan(task, encoding_scheme, Fb > 0)
    w = 0
    for n in plan:
        w += {"m_feats": F, "ij_feats": F, "box_embeddings": box_w, "b_feats": Fb}.get(n, H)
    return w
