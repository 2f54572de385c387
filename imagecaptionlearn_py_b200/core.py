"""Drop-in for the graph layer of the reference: `nn_utils/core.py`.

Same function names, argument meaning and results as the reference, with the TensorFlow graph + `sess.run`
(`nn_utils/core.py:625`) replaced by libicl_b200.so (hand-written sm_100a kernels, C-ABI in include/icl_b200.h):

    setup_bidirectional_lstm(n_hidden, data_norm, n_embedding_width, n_parallel)          core.py:271-332
    setup_core_architecture(task, encoding_scheme, batch_size, ...)                       core.py:443-514
    add_train_op(loss, lrn_rate, adam_epsilon, clip_norm)                                 core.py:74-106
    run_op(sess, op, batch_tensor_list, lstm_input_dropout, dropout, encoding_scheme,
           tasks, scope_names, include_labels)                                            core.py:517-626
    get_pred_scores_mcc(task, encoding_scheme, sess, batch_size, ids, data_dict, ...)     core.py:629-681
    get_collection(name)[0], variable_scope(name), Session(), Saver()                     (the TF symbols the scripts touch)

TF "collections" become entries of a module-level `Graph`; ops are `Op` handles ('train_op', 'loss', 'accuracy',
'predicted_proba', 'pred', optionally '<scope>/...').  `run_op` takes the unchanged `batch_tensors` dicts of
`nn_utils/data.py:load_batch` (float64 or float32, indices as float or int) and returns NumPy values of the
shape/dtype the TF op would: train_op -> None, loss/accuracy -> float32 scalar, predicted_proba -> float32 [B,C],
pred -> int64 [B].

Multi-task (`icl_multitask_lstm.py`): the reference's joint path feeds only the last task's sentences to the shared
LSTM (core.py:558-561 inside the loop at :544), which is a defect; here every task's sentences go through ONE
shared-weight encoder pass (concatenated) and each head indexes its own slice -- the intended semantics.
"""
import contextlib
import ctypes as C
import math
import os

import numpy as np

from . import _cabi
from . import data as nn_data

__all__ = ["set_random_seeds", "get_widths", "setup_bidirectional_lstm", "setup_core_architecture", "add_train_op", "setup_joint_loss",
           "run_op", "get_pred_scores_mcc", "get_collection", "variable_scope", "reset_default_graph", "Session",
           "Saver", "Op"]


class Op(object):
    def __init__(self, kind, scope=""):
        self.kind, self.scope = kind, scope

    def __repr__(self):
        return "<icl op %s%s>" % (self.scope + "/" if self.scope else "", self.kind)


class Graph(object):
    def __init__(self):
        self.lstm = None            # dict(n_hidden, data_norm, E)
        self.heads = []             # list of dicts, in creation order
        self.train = {}             # scope -> dict(lr, eps, clip)
        self.joint = None           # None / "simple_joint" / "weighted_joint" (setup_joint_loss)
        self.scope_stack = []
        self.seed = 20171201


_graph = Graph()


def reset_default_graph():
    global _graph
    _graph = Graph()


def set_random_seeds(seed=20171201):
    """core.py:10-18."""
    import random
    _graph.seed = seed
    np.random.seed(seed)
    random.seed(seed)


@contextlib.contextmanager
def variable_scope(name):
    _graph.scope_stack.append(name)
    try:
        yield
    finally:
        _graph.scope_stack.pop()


def _scope():
    # only the task scope matters to the scripts ('bidirectional_lstm' is handled by name)
    s = [x for x in _graph.scope_stack if x != "bidirectional_lstm"]
    return "/".join(s)


def get_collection(name):
    """tf.get_collection(name) for the names the scripts use: '[<scope>/]loss|accuracy|train_op|predicted_proba|pred'."""
    scope, _, kind = name.rpartition("/")
    if kind in ("loss", "accuracy", "train_op", "predicted_proba", "pred"):
        return [Op(kind, scope)]
    return []


def get_widths(start_width, depth, end_width=None):
    """core.py:121-143 (depth d -> d+1 widths when end_width is None)."""
    w = [int(start_width)]
    if end_width is None:
        for d in range(1, depth + 1):
            w.append(int(w[d - 1] / 2))
    else:
        for d in range(1, depth):
            w.append(int(max(w[d - 1] - end_width, 0) / 2 + end_width))
        w.append(int(end_width))
    return w


def setup_bidirectional_lstm(n_hidden, data_norm=False, n_embedding_width=300, n_parallel=64):
    """core.py:271-332.  n_parallel (TF while-loop parallel_iterations) has no meaning here and is ignored."""
    _graph.lstm = dict(n_hidden=int(n_hidden), data_norm=bool(data_norm), E=int(n_embedding_width))


def setup_core_architecture(task, encoding_scheme, batch_size, start_hidden_width, hidden_depth, weighted_classes,
                            activation, n_classes, n_mention_feats, box_embedding_width=None, n_box_feats=None):
    """core.py:443-514.  The head is registered under the current variable_scope (task name in multitask)."""
    if task not in _cabi.TASKS or encoding_scheme not in _cabi.ENCODINGS:
        raise ValueError("unknown task/encoding_scheme: %r %r" % (task, encoding_scheme))
    _graph.heads.append(dict(task=task, encoding_scheme=encoding_scheme, batch_size=int(batch_size),
                             widths=get_widths(start_hidden_width, hidden_depth), weighted=bool(weighted_classes),
                             activation=activation, n_classes=int(n_classes), F=int(n_mention_feats),
                             box_width=int(box_embedding_width or 0) if task == "affinity" else 0,
                             n_box_feats=int(n_box_feats or 0) if task == "affinity" else 0, scope=_scope()))


def setup_joint_loss(multitask_scheme):
    """icl_multitask_lstm.py:248-255: the joint loss over all heads.  'simple_joint' = sum of the task losses;
    'weighted_joint' = reduce_sum(setup_ffw(stack(losses)[None, :], [n_heads])): a trainable linear map (variables
    'hdn_1/Variable' [n,n] and 'hdn_1/Variable_1' [1,n], Xavier-initialised like every other weight, core.py:22-71) applied to
    the loss vector and summed, so d joint / d loss_t = sum_j W[t, j].  The n*n+n mixing variables live on the host (Session)."""
    if multitask_scheme not in ("simple_joint", "weighted_joint"):
        raise ValueError("unknown joint scheme %r" % (multitask_scheme,))
    _graph.joint = multitask_scheme
    return Op("loss", "")


def add_train_op(loss, lrn_rate, adam_epsilon, clip_norm):
    """core.py:74-106: Adam(lr, eps) with optional clip_by_global_norm over all variables that have a gradient.  Every call
    creates one optimizer = one slot of Adam state on the device (the `alternate` multitask scheme calls this once per task
    under that task's variable_scope, icl_multitask_lstm.py:387-393); learn rate / epsilon / clip of the first call apply."""
    _graph.train[_scope()] = dict(lr=float(lrn_rate), eps=float(adam_epsilon),
                                  clip=-1.0 if clip_norm is None else float(clip_norm))


class Session(object):
    """Stands in for tf.Session: owns the device model (parameters, Adam state, workspaces)."""

    def __init__(self, graph=None, max_seq_len=64, device=None, gemm_mode=_cabi.GEMM_TCGEN05_TF32, dist=None, seed=None):
        self.graph = graph or _graph
        self.max_seq_len = int(max_seq_len)
        self.gemm_mode = gemm_mode
        self.dist = dist                 # None or a torch.distributed process group marker (True = default group)
        self.device = device
        self.handle = None
        self.run_counter = 0
        self.last_train_stats = None
        self._slot = 0
        self._table_ref = None
        self._box_ref = None
        self._side = None
        self.joint_vars = None          # weighted_joint: dict(W, b, mW, vW, mb, vb) on the host
        self.base_seed = self.graph.seed if seed is None else seed
        self._grad_view = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        if self.handle is not None:
            _cabi.lib().icl_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- model construction ------------------------------------------------------------------------------------------
    def _config(self):
        g = self.graph
        if g.lstm is None or not g.heads:
            raise RuntimeError("setup_bidirectional_lstm / setup_core_architecture must be called first")
        cfg = _cabi.Config()
        cfg.embed_width, cfg.lstm_hidden, cfg.data_norm = g.lstm["E"], g.lstm["n_hidden"], int(g.lstm["data_norm"])
        cfg.max_seqs = sum(h["batch_size"] * (2 if h["task"] == "rel_cross" else 1) for h in g.heads)
        cfg.max_seq_len = self.max_seq_len
        cfg.n_heads = len(g.heads)
        for i, h in enumerate(g.heads):
            hc = cfg.heads[i]
            hc.task, hc.encoding = _cabi.TASKS[h["task"]], _cabi.ENCODINGS[h["encoding_scheme"]]
            hc.batch_size, hc.n_classes, hc.n_feats = h["batch_size"], h["n_classes"], h["F"]
            hc.box_width, hc.n_box_feats = h["box_width"], h["n_box_feats"]
            hc.n_hidden = len(h["widths"])
            for k, w in enumerate(h["widths"]):
                hc.widths[k] = w
            hc.activation = _cabi.ACTIVATIONS[h["activation"]]
            hc.weighted_classes = int(h["weighted"])
            hc.scope = h["scope"].encode()
        tr = next(iter(g.train.values())) if g.train else dict(lr=1e-3, eps=1e-8, clip=-1.0)
        cfg.learn_rate, cfg.adam_epsilon, cfg.clip_norm = tr["lr"], tr["eps"], tr["clip"]
        cfg.beta1, cfg.beta2 = 0.9, 0.999
        if self.device is None:
            import os
            self.device = int(os.environ.get("LOCAL_RANK", "0"))
        cfg.device, cfg.gemm_mode = self.device, self.gemm_mode
        return cfg

    @staticmethod
    def _bind_to_gpu_numa_node(device):
        """One process per GPU on a multi-socket host: run this rank (and allocate its pinned staging buffers, first touch) on the
        NUMA node its GPU hangs off, so the host packing and the H2D DMA of the ranks do not all cross one socket's memory
        controllers.  Silently does nothing when the topology is not visible (containers) or ICL_NO_NUMA_BIND is set."""
        if os.environ.get("ICL_NO_NUMA_BIND") or int(os.environ.get("LOCAL_WORLD_SIZE", "1")) < 2:
            return
        try:
            import torch
            p = torch.cuda.get_device_properties(device)
            bdf = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
            node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
            if node < 0:
                return
            cpus = set()
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
        except Exception:
            pass

    def _create(self, state=None):
        cfg = self._config()
        self._bind_to_gpu_numa_node(cfg.device)
        h = C.c_void_p()
        _cabi.check(_cabi.lib().icl_create(C.byref(cfg), C.byref(h)))
        self.handle = h
        self._slot = 0
        self._table_ref = None
        self._box_ref = None
        self._grad_view = None
        if state is None:
            self.initialize()
        else:
            self.load_state(state)
        self._set_loss_weights(None)

    def _world(self):
        if not self.dist:
            return 1
        import torch.distributed as td
        return td.get_world_size()

    def _set_loss_weights(self, base):
        """d joint / d loss_t per head (`base`, None = 1) times the data-parallel factor: a `weighted_classes` head's loss is the
        MEAN over its local batch (core.py:244-267 as executed), so the SUM all-reduce over N ranks must carry 1/N of each rank's
        gradient to equal the gradient of the global-batch mean; summed-CE heads need no factor."""
        n, world = len(self.graph.heads), self._world()
        w = [1.0 if base is None else float(base[i]) for i in range(n)]
        if world > 1:
            w = [w[i] / world if self.graph.heads[i]["weighted"] else w[i] for i in range(n)]
        if base is None and all(x == 1.0 for x in w):
            _cabi.check(_cabi.lib().icl_set_loss_weights(self.handle, None))
        else:
            _cabi.check(_cabi.lib().icl_set_loss_weights(self.handle, (C.c_float * n)(*w)))

    def ensure(self, T=None):
        if T is not None and T > self.max_seq_len:
            state = self.state_dict() if self.handle is not None else None
            slot = self._slot
            self.close()
            self.max_seq_len = int(T)
            self._create(state)
            self.set_optimizer_slot(slot)          # _create starts on slot 0: the `alternate` scheme had already selected one
        elif self.handle is None:
            self._create()

    def param_info(self):
        L = _cabi.lib()
        out = []
        for i in range(L.icl_param_count(self.handle)):
            name, r, c, off = C.c_char_p(), C.c_int32(), C.c_int32(), C.c_int64()
            _cabi.check(L.icl_param_info(self.handle, i, C.byref(name), C.byref(r), C.byref(c), C.byref(off)))
            out.append((name.value.decode(), r.value, c.value, off.value))
        return out

    def initialize(self):
        """global_variables_initializer: glorot-uniform LSTM kernels + zero bias (TF default), the reference's Xavier
        rule for the heads (core.py:32-36,59-63).  TF's Philox stream cannot be reproduced; parity tests inject weights."""
        rng = np.random.RandomState(self.graph.seed % (2 ** 32))
        for name, r, c, _ in self.param_info():
            if name.endswith("basic_lstm_cell/bias"):
                v = np.zeros((r, c), np.float32)
            else:
                lim = math.sqrt(6.0 / (r + c))
                v = rng.uniform(-lim, lim, (r, c)).astype(np.float32)
            self.set_tensor(name, v)

    JOINT_NAMES = ("hdn_1/Variable", "hdn_1/Variable_1")

    def _joint(self):
        """Host-side variables of the weighted_joint loss mixer (created on first use, Xavier like core.py:32-36,59-63)."""
        if self.joint_vars is None:
            n = len(self.graph.heads)
            rng = np.random.RandomState((self.graph.seed + 17) % (2 ** 32))
            lw, lb = math.sqrt(6.0 / (n + n)), math.sqrt(6.0 / (1 + n))
            self.joint_vars = dict(W=rng.uniform(-lw, lw, (n, n)).astype(np.float32), b=rng.uniform(-lb, lb, (1, n)).astype(np.float32),
                                   mW=np.zeros((n, n), np.float32), vW=np.zeros((n, n), np.float32),
                                   mb=np.zeros((1, n), np.float32), vb=np.zeros((1, n), np.float32))
        return self.joint_vars

    def train_weighted_joint(self, batch_tensor_list, keep_in, keep):
        """One synchronous train step of the weighted_joint scheme: device gradients with per-head weights sum_j W[t,j], the
        mixer's own gradients (dW[t,j] = loss_t, db_j = 1) take part in the global-norm clip and get the same TF-Adam step."""
        L = _cabi.lib()
        g = self.graph
        self.ensure(max((bt["sentences"].shape[1] if "sentences" in bt else int(np.max(bt["seq_lengths"])))
                        for bt in batch_tensor_list))
        jv = self._joint()
        n = len(g.heads)
        self._set_loss_weights(jv["W"].sum(1))
        keepalive = []
        b = self.build_batch(batch_tensor_list, True, keepalive)
        self._bind_stream()
        self.run_counter += 1
        seed = (self.base_seed * 1000003 + self.run_counter) & 0xFFFFFFFFFFFFFFFF
        self.last_seed = seed
        outs = (_cabi.HeadOut * _cabi.MAX_HEADS)()
        _cabi.check(L.icl_upload(self.handle, C.byref(b)))
        _cabi.check(L.icl_run_resident(self.handle, _cabi.OP_GRADS, keep_in, keep, seed))
        if self.dist:
            import torch.distributed as td
            self.allreduce_grads()
        _cabi.check(L.icl_fetch(self.handle, outs))
        losses = np.array([outs[i].loss for i in range(n)], np.float32)
        if self.dist:
            import torch
            import torch.distributed as td
            t = torch.tensor(losses, device="cuda:%d" % self.device)
            td.all_reduce(t, op=td.ReduceOp.SUM)
            losses = t.cpu().numpy()
            for i, h in enumerate(g.heads):                      # mean-CE heads: the global-batch mean, not the sum of local means
                if h["weighted"]:
                    losses[i] /= td.get_world_size()
        dW = np.repeat(losses[:, None], n, 1).astype(np.float32)
        db = np.ones((1, n), np.float32)
        gn = C.c_float()
        _cabi.check(L.icl_apply_update_ex(self.handle, float(np.sum(dW.astype(np.float64) ** 2) + n), C.byref(gn)))
        tr = next(iter(g.train.values()))
        scale = tr["clip"] / max(gn.value, tr["clip"]) if tr["clip"] > 0 else 1.0
        t = C.c_int64()
        _cabi.check(L.icl_get_step(self.handle, C.byref(t)))
        lr_t = tr["lr"] * math.sqrt(1.0 - 0.999 ** t.value) / (1.0 - 0.9 ** t.value)
        for name, grad in (("W", dW), ("b", db)):
            gs = (grad * scale).astype(np.float32)
            jv["m" + name] = (0.9 * jv["m" + name] + 0.1 * gs).astype(np.float32)
            jv["v" + name] = (0.999 * jv["v" + name] + 0.001 * gs * gs).astype(np.float32)
            jv[name] = (jv[name] - lr_t * jv["m" + name] / (np.sqrt(jv["v" + name]) + tr["eps"])).astype(np.float32)
        self._set_loss_weights(None)
        self.last_train_stats = [dict(loss=np.float32(outs[i].loss), accuracy=np.float32(outs[i].accuracy)) for i in range(n)]
        return losses

    def get_tensor(self, name, kind=0):
        if name in self.JOINT_NAMES and self.graph.joint == "weighted_joint":
            jv = self._joint()
            key = "W" if name == self.JOINT_NAMES[0] else "b"
            return jv[{0: key, 2: "m" + key, 3: "v" + key}[kind]].copy()
        for n, r, c, _ in self.param_info():
            if n == name:
                out = np.empty((r, c), np.float32)
                _cabi.check(_cabi.lib().icl_get_tensor(self.handle, kind, name.encode(), _cabi.np_ptr(out)))
                return out
        raise KeyError(name)

    def set_tensor(self, name, value, kind=0):
        if name in self.JOINT_NAMES and self.graph.joint == "weighted_joint":
            jv = self._joint()
            key = "W" if name == self.JOINT_NAMES[0] else "b"
            k2 = {0: key, 2: "m" + key, 3: "v" + key}[kind]
            jv[k2] = np.asarray(value, np.float32).reshape(jv[k2].shape).copy()
            return
        v = np.ascontiguousarray(np.asarray(value, dtype=np.float32))
        _cabi.check(_cabi.lib().icl_set_tensor(self.handle, kind, name.encode(), _cabi.np_ptr(v)))

    def set_token_table(self, table):
        """Keep `table` ([n_rows, E] float32: e.g. every caption matrix of the corpus concatenated, data._flat_table) resident
        on the device; batches with 'token_rows' then ship 4 bytes per token instead of an embedding row."""
        if self._table_ref is not table:
            t = np.ascontiguousarray(table, dtype=np.float32)
            _cabi.check(_cabi.lib().icl_set_token_table(self.handle, _cabi.np_ptr(t), int(t.shape[0])))
            self._table_ref = table

    def set_box_table(self, table):
        """Keep the corpus' box features ([n_boxes, box_width] float32) resident on the device; affinity batches with 'box_rows'
        then ship one int32 per mention-box pair instead of a 4096-float row."""
        if self._box_ref is not table:
            t = np.ascontiguousarray(table, dtype=np.float32)
            _cabi.check(_cabi.lib().icl_set_box_table(self.handle, _cabi.np_ptr(t), int(t.shape[0]), int(t.shape[1])))
            self._box_ref = table

    def set_optimizer_slot(self, slot):
        """Select the Adam state (m, v, step) used by the next updates / get_tensor(kind 2, 3) calls."""
        if slot != self._slot:
            _cabi.check(_cabi.lib().icl_set_optimizer_slot(self.handle, int(slot)))
            self._slot = int(slot)

    def state_dict(self):
        """Parameters under their TF variable names + the Adam state of every optimizer slot (slot 0: 'adam_m/<name>',
        'adam_v/<name>', 'adam_step'; slot s > 0: the same keys with an '@s' suffix on the prefix)."""
        L = _cabi.lib()
        st = {}
        info = self.param_info()
        for n, _, _, _ in info:
            st[n] = self.get_tensor(n, 0)
        keep = self._slot
        for s in range(L.icl_optimizer_slots(self.handle)):
            self.set_optimizer_slot(s)
            sfx = "" if s == 0 else "@%d" % s
            for n, _, _, _ in info:
                st["adam_m%s/%s" % (sfx, n)] = self.get_tensor(n, 2)
                st["adam_v%s/%s" % (sfx, n)] = self.get_tensor(n, 3)
            t = C.c_int64()
            _cabi.check(L.icl_get_step(self.handle, C.byref(t)))
            st["adam_step" + sfx] = np.int64(t.value)
        self.set_optimizer_slot(keep)
        if self.graph.joint == "weighted_joint":
            jv = self._joint()
            for name, key in zip(self.JOINT_NAMES, ("W", "b")):
                st[name], st["adam_m/" + name], st["adam_v/" + name] = jv[key].copy(), jv["m" + key].copy(), jv["v" + key].copy()
        return st

    @staticmethod
    def _alias(name):
        """'<s>/<s>/rest' <-> '<s>/rest': checkpoints written before the multitask heads carried TensorFlow's doubled scope."""
        p = name.split("/")
        if len(p) >= 3 and p[0] == p[1]:
            return "/".join(p[1:])
        return None

    def load_state(self, st, strict=True):
        """Loads parameters (+ Adam state of every optimizer slot) by TF variable name.  strict (default): every model parameter
        must be present and every key must be recognised -- a checkpoint whose names do not match raises instead of silently
        leaving heads at their random initialisation.  strict=False restores what matches and returns (missing, unexpected)."""
        info = self.param_info()
        st = dict(st)
        canon = {self._alias(n): n for n, _, _, _ in info if self._alias(n) is not None}
        for k in list(st.keys()):                                # accept the single-scope spelling of older files
            pre, name = (k.split("/", 1)[0] + "/", k.split("/", 1)[1]) if k.startswith(("adam_m", "adam_v")) and "/" in k else ("", k)
            if name in canon and pre + canon[name] not in st:
                st[pre + canon[name]] = st.pop(k)
        used = set()
        if self.graph.joint == "weighted_joint":
            for name in self.JOINT_NAMES:
                for kind, pre in ((0, ""), (2, "adam_m/"), (3, "adam_v/")):
                    if pre + name in st:
                        self.set_tensor(name, st[pre + name], kind)
                        used.add(pre + name)
        missing = []
        for n, r, c, _ in info:
            if n in st:
                a = np.asarray(st[n])
                if a.size != r * c:
                    raise ValueError("load_state: %r has %d elements, the model's tensor is [%d, %d]" % (n, a.size, r, c))
                self.set_tensor(n, a.reshape(r, c), 0)
                used.add(n)
            else:
                missing.append(n)
        keep = self._slot
        slots = sorted(set([0] + [int(k.split("@")[1]) for k in st if k.startswith("adam_step@")]))
        for s in slots:
            sfx = "" if s == 0 else "@%d" % s
            if "adam_step" + sfx not in st:
                continue
            self.set_optimizer_slot(s)
            for n, r, c, _ in info:
                km, kv = "adam_m%s/%s" % (sfx, n), "adam_v%s/%s" % (sfx, n)
                if km in st and kv in st:
                    self.set_tensor(n, np.asarray(st[km]).reshape(r, c), 2)
                    self.set_tensor(n, np.asarray(st[kv]).reshape(r, c), 3)
                    used.update((km, kv))
            _cabi.check(_cabi.lib().icl_set_step(self.handle, int(st["adam_step" + sfx])))
            used.add("adam_step" + sfx)
        self.set_optimizer_slot(keep)
        unexpected = sorted(k for k in st if k not in used)
        if strict and (missing or unexpected):
            raise KeyError("load_state: the checkpoint does not match the graph: %d model tensors missing (%s), %d unrecognised keys "
                           "(%s); pass strict=False to restore the intersection" %
                           (len(missing), ", ".join(missing[:4]), len(unexpected), ", ".join(unexpected[:4])))
        return missing, unexpected

    # -- execution ---------------------------------------------------------------------------------------------------
    def _bind_stream(self):
        if os.environ.get("ICL_NO_TORCH_STREAM"):       # tools run under compute-sanitizer: the library's default stream, no torch
            return
        try:
            import torch
            if torch.cuda.is_available():
                torch.cuda.set_device(self.device)
                _cabi.lib().icl_set_stream(self.handle, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        except ImportError:
            pass

    def grad_tensor(self):
        """torch view of the flat device gradient buffer (for the NCCL all-reduce)."""
        if self._grad_view is None:
            import torch
            p, n = C.c_void_p(), C.c_int64()
            _cabi.check(_cabi.lib().icl_grad_buffer(self.handle, C.byref(p), C.byref(n)))

            class _Arr(object):
                __cuda_array_interface__ = dict(shape=(n.value,), typestr="<f4", data=(p.value, False), version=2)
            self._grad_view = torch.as_tensor(_Arr(), device="cuda:%d" % self.device)
        return self._grad_view

    def _overlap_group(self):
        """A second NCCL communicator capped at a few CTAs (ncclConfig_t.maxCTAs) for the collectives that run WHILE the backward
        pass computes: the BPTT clusters leave 8 of the 148 SMs idle, and an uncapped all-reduce kernel would take SMs from them
        (measured in round 1: 1.598 vs 1.551 ms per step on 2 x B200).  None when the installed torch cannot configure it."""
        if getattr(self, "_ar_group", False) is False:
            self._ar_group = None
            try:
                import torch.distributed as td
                opts = td.ProcessGroupNCCL.Options()
                opts.config.max_ctas = int(os.environ.get("ICL_AR_CTAS", "4"))
                opts.config.min_ctas = 1
                self._ar_group = td.new_group(backend="nccl", pg_options=opts)
            except Exception:
                self._ar_group = None
        return self._ar_group

    def _nvls_setup(self):
        """Puts the flat gradient buffer into symmetric memory (torch.distributed._symmetric_memory: the same allocation on every rank,
        mapped into one NVSwitch multicast object) and hands it to the library (icl_adopt_grad_buffer).  Returns (tensor, handle) or
        None; every rank takes the same decision (a MIN all-reduce of the outcome), so a rank that cannot do it never leaves the
        others waiting in a collective."""
        import torch
        import torch.distributed as td
        ok, t, hdl = 1, None, None
        try:
            if os.environ.get("ICL_AR_NVLS", "1") == "0":
                raise RuntimeError("disabled")
            import torch.distributed._symmetric_memory as symm_mem
            p, n = C.c_void_p(), C.c_int64()
            _cabi.check(_cabi.lib().icl_grad_buffer(self.handle, C.byref(p), C.byref(n)))
            t = symm_mem.empty((n.value + 3) // 4 * 4, dtype=torch.float32, device=torch.device("cuda", self.device))
        except Exception:
            ok = 0
        flag = torch.tensor([ok], device="cuda:%d" % self.device, dtype=torch.int32)
        td.all_reduce(flag, op=td.ReduceOp.MIN)
        if int(flag.item()) == 0:
            return None
        try:
            hdl = symm_mem.rendezvous(t, td.group.WORLD)
            ok = 1 if int(getattr(hdl, "multicast_ptr", 0) or 0) != 0 else 0
        except Exception:
            ok = 0
        flag.fill_(ok)
        td.all_reduce(flag, op=td.ReduceOp.MIN)
        if int(flag.item()) == 0:
            return None
        _cabi.check(_cabi.lib().icl_adopt_grad_buffer(self.handle, C.c_void_p(t.data_ptr()), t.numel()))
        self._grad_view = None
        return (t, hdl)

    def allreduce_grads(self):
        """SUM all-reduce of the flat gradient buffer of the step just enqueued (the loss is a SUM over examples, core.py:267).

        Default on an NVSwitch box: the buffer lives in symmetric / multicast memory and `icl_nvls_allreduce` sums it IN the switch --
        rank r reads slice r of all ranks with multimem.ld_reduce and broadcasts the sum with multimem.st (our kernel, no NCCL, no
        staging), between two cross-rank barriers of the symmetric-memory handle.  Fallbacks: NCCL in three buckets overlapped with
        the backward pass at 8 ranks (`_overlap_group`), else one NCCL all-reduce after the backward pass (ICL_AR_NVLS=0 /
        ICL_AR_OVERLAP=0|1 force)."""
        import torch
        import torch.distributed as td
        L = _cabi.lib()
        if getattr(self, "_nvls", False) is False or getattr(self, "_nvls_handle", None) is not self.handle:
            self._nvls = self._nvls_setup() if td.get_backend() == "nccl" else None
            self._nvls_handle = self.handle
        if self._nvls is not None:
            t, hdl = self._nvls
            hdl.barrier(channel=0)                      # every rank's gradients are complete
            _cabi.check(L.icl_nvls_allreduce(self.handle, C.c_void_p(int(hdl.multicast_ptr)), td.get_rank(), td.get_world_size()))
            hdl.barrier(channel=1)                      # every slice has been written to every rank
            return
        g = self.grad_tensor()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
            self._side2 = torch.cuda.Stream(device=self.device)
            split, bw = C.c_int64(), C.c_int64()
            _cabi.check(L.icl_grad_split(self.handle, C.byref(split)))
            _cabi.check(L.icl_grad_split_lstm(self.handle, C.byref(bw)))
            self._split, self._bw = int(split.value), int(bw.value)
        main = torch.cuda.current_stream()
        # measured (card2048, profiles/r2_allreduce_overlap.md): 8 GPUs 1.453 -> 1.428 ms per step, but 4 GPUs 1.428 -> 1.450 and 2 GPUs
        # 1.403 -> 1.422 (three launches + cross-stream waits cost more than a 9 MB all-reduce between few GPUs): on by default at 8 ranks
        ov = os.environ.get("ICL_AR_OVERLAP")
        overlap = (ov != "0" if ov is not None else td.get_world_size() >= 8) and self._split < g.numel()
        grp = self._overlap_group() if overlap else None
        if grp is None:
            td.all_reduce(g, op=td.ReduceOp.SUM)
            return
        _cabi.check(L.icl_wait_head_grads(self.handle, C.c_void_p(self._side.cuda_stream)))
        with torch.cuda.stream(self._side):
            td.all_reduce(g[self._split:], op=td.ReduceOp.SUM, group=grp)
        _cabi.check(L.icl_wait_fw_lstm_grads(self.handle, C.c_void_p(self._side2.cuda_stream)))
        with torch.cuda.stream(self._side2):
            self._side2.wait_stream(self._side)               # one communicator: its collectives stay in issue order
            td.all_reduce(g[:self._bw], op=td.ReduceOp.SUM, group=grp)
        td.all_reduce(g[self._bw:self._split], op=td.ReduceOp.SUM)
        main.wait_stream(self._side)
        main.wait_stream(self._side2)

    def param_tensor(self):
        """torch view of the flat device parameter buffer (rank-0 broadcast of the initial weights)."""
        import torch
        p, n = C.c_void_p(), C.c_int64()
        _cabi.check(_cabi.lib().icl_param_buffer(self.handle, C.byref(p), C.byref(n)))

        class _Arr(object):
            __cuda_array_interface__ = dict(shape=(n.value,), typestr="<f4", data=(p.value, False), version=2)
        return torch.as_tensor(_Arr(), device="cuda:%d" % self.device)

    def build_batch(self, batch_tensor_list, include_labels, keepalive, head_ids=None):
        """batch_tensor_list[i] feeds head head_ids[i] (default: heads 0..n-1 in creation order); heads that are not fed are
        marked inactive, like TF evaluating only the fetched task's subgraph (icl_multitask_lstm.py:327-334)."""
        g = self.graph
        b = _cabi.Batch()
        head_ids = list(range(len(batch_tensor_list))) if head_ids is None else list(head_ids)
        first = batch_tensor_list[0]
        by_rows = "token_rows" in first
        if by_rows:
            table = first["token_table"]
            if any(bt.get("token_table") is not table for bt in batch_tensor_list):
                raise ValueError("token_rows batches of one call must share one token_table")
            self.set_token_table(table)
        packed = by_rows or "sentences_packed" in first
        sents, lens = [], []
        off = 0
        b.n_heads = len(g.heads)
        for i in range(len(g.heads)):
            b.heads[i].inactive = 0 if i in head_ids else 1
        T = 0
        for i, bt in zip(head_ids, batch_tensor_list):
            hb = b.heads[i]
            hb.sent_offset = off
            ln = np.asarray(bt["seq_lengths"])
            lens.append(ln)
            s = bt["token_rows"] if by_rows else bt["sentences_packed"] if packed else bt["sentences"]
            sents.append(s)
            if not packed:
                T = max(T, s.shape[1])
            off += len(ln)
            task = g.heads[i]["task"]
            hd = g.heads[i]
            Bh, E = hd["batch_size"], g.lstm["E"]

            def want(key, arr, shape):
                # the reference's placeholders have static shapes (core.py:348,476): a wrong batch raises there, and must not
                # become an out-of-bounds host read in icl_upload here
                if tuple(np.shape(arr)) != tuple(shape):
                    raise ValueError("run_op: %r of task %r has shape %r, the graph was built for %r" %
                                     (key, task, tuple(np.shape(arr)), tuple(shape)))
            n_sent = Bh * (2 if task == "rel_cross" else 1)
            if "sentences" in bt and not packed:
                if len(ln) > n_sent or s.ndim != 3 or s.shape[0] != len(ln) or s.shape[2] != E:
                    raise ValueError("run_op: 'sentences' of task %r has shape %r for %d lengths; the graph takes at most [%d, T, %d]" %
                                     (task, tuple(s.shape), len(ln), n_sent, E))
            else:
                if len(ln) > n_sent:
                    raise ValueError("run_op: task %r feeds %d sequences, the graph was built for %d" % (task, len(ln), n_sent))
                n_tok = int(np.sum(ln))
                want("token_rows" if by_rows else "sentences_packed", s, (n_tok,) if by_rows else (n_tok, E))
            idxs = []
            for k, name in enumerate(_cabi.INDEX_ORDER):
                if name in bt:
                    a = _cabi.as_supported(bt[name])
                    want(name, a, (Bh, 3))
                    idxs.append(a)
                else:
                    idxs.append(None)
            dts = set(a.dtype for a in idxs if a is not None)
            if len(dts) > 1:
                idxs = [None if a is None else a.astype(np.float64) for a in idxs]
            for k, a in enumerate(idxs):
                if a is not None:
                    keepalive.append(a)
                    hb.idx[k] = a.ctypes.data
                    hb.idx_dtype = _cabi.dtype_code(a)

            def put(field, key, width):
                if key in bt and bt[key] is not None:
                    a = _cabi.as_supported(bt[key])
                    want(key, a, (Bh, width))
                    keepalive.append(a)
                    setattr(hb, field, a.ctypes.data)
                    setattr(hb, field + "_dtype", _cabi.dtype_code(a))
            put("feats", "ij_feats" if "rel" in task else "m_feats", hd["F"])
            if task == "affinity":
                if "box_rows" in bt:                       # rows of the device-resident box table instead of [B, 4096] floats
                    self.set_box_table(bt["box_table"])
                    a = np.ascontiguousarray(bt["box_rows"], dtype=np.int32)
                    want("box_rows", a, (Bh,))
                    keepalive.append(a)
                    hb.box_rows = a.ctypes.data
                else:
                    put("box", "box_embeddings", hd["box_width"])
                put("bfeats", "b_feats", hd["n_box_feats"])
            if include_labels:
                put("labels", "labels", hd["n_classes"])
        per_head = len(sents) > 1 and not packed
        if len(sents) == 1:
            x = _cabi.as_supported(sents[0])
            ln = _cabi.as_supported(lens[0])
        elif per_head:
            # every task's load_batch built its own padded tensor: the library reads them in place (icl_head_batch.sentences)
            arrs = [_cabi.as_supported(s_) for s_ in sents]
            if len(set(a.dtype for a in arrs)) > 1 or arrs[0].dtype.kind != "f":
                arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in arrs]
            for i, a in zip(head_ids, arrs):
                b.heads[i].sentences, b.heads[i].n_seqs, b.heads[i].padded_T = a.ctypes.data, a.shape[0], a.shape[1]
            keepalive += arrs
            x = arrs[0]
            ln = _cabi.as_supported(np.concatenate(lens, 0))
        else:
            x = _cabi.as_supported(np.concatenate(sents, 0))
            ln = _cabi.as_supported(np.concatenate(lens, 0))
        if by_rows:
            x = np.ascontiguousarray(x, dtype=np.int32)
        elif x.dtype.kind != "f":
            x = x.astype(np.float32)
        keepalive += [x, ln]
        if by_rows:
            b.token_rows, b.sentences, b.sent_packed = x.ctypes.data, None, 1
        elif per_head:
            b.sentences, b.sent_dtype, b.sent_packed = None, _cabi.dtype_code(x), 0
        else:
            b.sentences, b.sent_dtype, b.sent_packed = x.ctypes.data, _cabi.dtype_code(x), int(packed)
        b.seq_lengths, b.len_dtype = ln.ctypes.data, _cabi.dtype_code(ln)
        b.n_seqs = len(ln)
        b.padded_T = T if not packed else int(ln.max()) if len(ln) else 0
        rank = 0
        if self.dist:
            import torch.distributed as td
            rank = td.get_rank()
        b.seq_gid_offset = rank * b.n_seqs
        b.ex_gid_offset = rank * max(h["batch_size"] for h in g.heads)
        return b

    def run(self, op_kind, batch_tensor_list=None, keep_in=1.0, keep=1.0, include_labels=False, head_ids=None):
        if isinstance(op_kind, Op) and op_kind.kind == "init":        # sess.run(tf.global_variables_initializer()), icl_core_lstm.py:110
            self.ensure()
            self.initialize()
            return None
        L = _cabi.lib()
        g = self.graph
        if head_ids is None and len(batch_tensor_list) != len(g.heads):
            raise ValueError("got %d batches for %d heads" % (len(batch_tensor_list), len(g.heads)))
        keepalive = []
        need_T = max((bt["sentences"].shape[1] if "sentences" in bt else int(np.max(bt["seq_lengths"])))
                     for bt in batch_tensor_list)
        self.ensure(need_T)
        b = self.build_batch(batch_tensor_list, include_labels, keepalive, head_ids)
        self._bind_stream()
        outs = (_cabi.HeadOut * _cabi.MAX_HEADS)()
        res = []
        for i, h in enumerate(g.heads):
            pr = np.empty((h["batch_size"], h["n_classes"]), np.float32)
            pd = np.empty((h["batch_size"],), np.int64)
            outs[i].proba = pr.ctypes.data_as(C.POINTER(C.c_float))
            outs[i].pred = pd.ctypes.data_as(C.POINTER(C.c_int64))
            res.append([pr, pd])
        self.run_counter += 1
        seed = (self.base_seed * 1000003 + self.run_counter) & 0xFFFFFFFFFFFFFFFF
        self.last_seed = seed
        if op_kind == _cabi.OP_TRAIN and self.dist:
            import torch.distributed as td
            _cabi.check(L.icl_upload(self.handle, C.byref(b)))
            _cabi.check(L.icl_run_resident(self.handle, _cabi.OP_GRADS, keep_in, keep, seed))
            self.allreduce_grads()
            _cabi.check(L.icl_apply_update(self.handle))
            _cabi.check(L.icl_fetch(self.handle, outs))
        else:
            _cabi.check(L.icl_run(self.handle, op_kind, C.byref(b), keep_in, keep, seed, outs))
        return [dict(proba=r[0], pred=r[1], loss=np.float32(outs[i].loss), accuracy=np.float32(outs[i].accuracy))
                for i, r in enumerate(res)]


    def train_async(self, batch_tensor_list, keep_in, keep, head_ids=None):
        """One pipelined `train_op` step (the reference discards its result, icl_core_lstm.py:138-150): the batch is packed
        and copied on the copy stream into the idle input set while the previous step still computes, the step is enqueued
        and the call returns.  `last_train_stats` holds loss / accuracy per head of the PREVIOUS step (read back
        asynchronously every step)."""
        L = _cabi.lib()
        g = self.graph
        if head_ids is None and len(batch_tensor_list) != len(g.heads):
            raise ValueError("got %d batches for %d heads" % (len(batch_tensor_list), len(g.heads)))
        keepalive = []
        need_T = max((bt["sentences"].shape[1] if "sentences" in bt else int(np.max(bt["seq_lengths"])))
                     for bt in batch_tensor_list)
        self.ensure(need_T)
        b = self.build_batch(batch_tensor_list, True, keepalive, head_ids)
        self._bind_stream()
        prev = (_cabi.HeadOut * _cabi.MAX_HEADS)()
        self.run_counter += 1
        seed = (self.base_seed * 1000003 + self.run_counter) & 0xFFFFFFFFFFFFFFFF
        self.last_seed = seed
        if self.dist:
            import torch.distributed as td
            _cabi.check(L.icl_upload(self.handle, C.byref(b)))
            _cabi.check(L.icl_run_resident(self.handle, _cabi.OP_GRADS, keep_in, keep, seed))
            self.allreduce_grads()
            _cabi.check(L.icl_apply_update(self.handle))
            _cabi.check(L.icl_poll_stats(self.handle, prev))
        else:
            _cabi.check(L.icl_train_async(self.handle, C.byref(b), keep_in, keep, seed, prev))
        self.last_train_stats = [dict(loss=np.float32(prev[i].loss), accuracy=np.float32(prev[i].accuracy))
                                 for i in range(len(g.heads))]


def run_op(sess, op, batch_tensor_list, lstm_input_dropout, dropout, encoding_scheme, tasks, scope_names,
           include_labels=False):
    """core.py:517-626.  `tasks`/`scope_names` must list the heads in the order they were set up (as the
    reference's scripts do); `encoding_scheme` was fixed at setup time and is checked, not re-applied."""
    g = sess.graph
    scopes = [h["scope"] for h in g.heads]
    all_tasks = [h["task"] for h in g.heads]
    if list(tasks) == all_tasks:
        head_ids = None
    else:       # a subset of the heads is fed (multitask predict / alternate training): match by scope name, then by task
        head_ids = []
        for t, sc in zip(tasks, scope_names):
            cand = [i for i, h in enumerate(g.heads) if h["task"] == t and (h["scope"] == sc or sc == "")]
            if not cand:
                raise ValueError("run_op: no head for task %r (scope %r) in the graph's heads %r" % (t, sc, all_tasks))
            head_ids.append(cand[0])
    for h in g.heads:
        if h["encoding_scheme"] != encoding_scheme:
            raise ValueError("encoding_scheme differs from the one the graph was built with")
    kind = {"train_op": _cabi.OP_TRAIN}.get(op.kind, _cabi.OP_PREDICT)
    if kind == _cabi.OP_TRAIN and not include_labels:
        raise ValueError("train_op needs include_labels=True")
    if op.kind == "train_op":
        keys = list(g.train.keys())
        sess.ensure()
        sess.set_optimizer_slot(keys.index(op.scope) if op.scope in keys else 0)
    if op.kind == "train_op" and g.joint == "weighted_joint" and head_ids is None and len(g.heads) > 1:
        sess.train_weighted_joint(batch_tensor_list, float(lstm_input_dropout), float(dropout))
        return None
    if op.kind == "train_op" and not os.environ.get("ICL_SYNC_TRAIN"):
        sess.train_async(batch_tensor_list, float(lstm_input_dropout), float(dropout), head_ids)
        return None                                           # sess.run(train_op) returns None (core.py:625)
    res = sess.run(kind, batch_tensor_list, float(lstm_input_dropout), float(dropout), include_labels, head_ids)
    if op.kind == "train_op":
        return None
    if op.kind == "loss" and op.scope == "" and len(g.heads) > 1 and head_ids is None:
        if g.joint == "weighted_joint":
            jv = sess._joint()
            return np.float32(np.sum(np.array([r["loss"] for r in res], np.float32)[None, :] @ jv["W"] + jv["b"]))
        return np.float32(sum(r["loss"] for r in res))       # simple_joint: sum of the task losses
    if op.scope in scopes:
        i = scopes.index(op.scope)
    else:
        i = head_ids[0] if head_ids else 0
    return res[i][{"predicted_proba": "proba"}.get(op.kind, op.kind)]


def get_pred_scores_mcc(task, encoding_scheme, sess, batch_size, ids, data_dict, n_classes, log=None):
    """core.py:629-681: edge-pad ids to a multiple of B (always >= 1 pad), predict with keep-probs 1.0, keep the
    first B-pad rows of the last batch."""
    pred_scores = dict()
    id_matrix, pad = nn_data.pad_ids_for_predict(list(ids), batch_size)
    scope = _scope()
    for i in range(id_matrix.shape[0]):
        if log is not None:
            log.log_status('info', None, 'Predicting; %d batches complete (%.2f%%)', i, 100.0 * i / id_matrix.shape[0])
        # keep-probabilities are 1.0 here, so every distinct caption of the batch is encoded once (exact; data.load_batch)
        bt = nn_data.load_batch(list(id_matrix[i]), data_dict, task, n_classes, packed=nn_data.default_packing(),
                                dedup=not os.environ.get("ICL_NO_DEDUP"))
        scores = run_op(sess, Op("predicted_proba", scope), [bt], 1.0, 1.0, encoding_scheme, [task], [scope], False)
        n_rows = len(scores) if i < id_matrix.shape[0] - 1 else batch_size - pad
        for j in range(n_rows):
            pred_scores[id_matrix[i][j]] = scores[j].copy()
    return pred_scores, data_dict['labels']


class Saver(object):
    """tf.train.Saver stand-in (icl_core_lstm.py:107,155,393-394): one .npz keyed by the TF variable names,
    plus the Adam moments and step so training can resume."""

    def __init__(self, max_to_keep=100):
        pass

    def save(self, sess, path):
        sess.ensure()
        st = sess.state_dict()
        np.savez(path + ".npz", **st)
        if os.environ.get("ICL_SAVE_TF_BUNDLE"):        # additionally a TensorFlow Saver-V2 bundle (<path>.index / .data-*)
            from . import tf_checkpoint
            tf_checkpoint.write_bundle(path, tf_checkpoint.from_state_dict(st))
        return path

    def restore(self, sess, path):
        """Restores `<path>.npz` (ours) or, if only `<path>.index` exists, a TensorFlow Saver-V2 checkpoint of the reference."""
        sess.ensure()
        if not os.path.exists(path + ".npz") and os.path.exists(path + ".index"):
            from . import tf_checkpoint
            sess.load_state(tf_checkpoint.to_state_dict(tf_checkpoint.read_bundle(path)))
            return
        with np.load(path + ".npz") as z:
            sess.load_state({k: z[k] for k in z.files})


class train(object):
    """`tf.train` as the reference scripts use it: `tf.train.Saver(max_to_keep=100)` (icl_core_lstm.py:107)."""
    Saver = Saver


def global_variables_initializer():
    """`sess.run(tf.global_variables_initializer())` (icl_core_lstm.py:110): creates the device model and initialises it."""
    return Op("init")


def dump_tf_vars():
    """nn_utils/core.py:684-697 prints the graph's trainable variables; here: the heads the graph holds (the variables get their
    TF names when a Session creates the device model)."""
    g = _graph
    print("bidirectional_lstm: %s; heads: %s" % (g.lstm, [(h.get("scope", ""), h.get("task")) for h in g.heads]))


def _debug_mask(sess, stream, n, keep, seed=None):
    """Test hook: the 0/1 dropout mask the kernels derive for elements [0,n) of `stream` (see csrc/icl_kernels.cuh)."""
    out = np.empty(int(n), np.float32)
    _cabi.check(_cabi.lib().icl_debug_mask(sess.handle, C.c_uint64(sess.last_seed if seed is None else seed),
                                           C.c_uint32(stream), 0, int(n), float(keep), _cabi.np_ptr(out)))
    return out


Session.debug_mask = _debug_mask
