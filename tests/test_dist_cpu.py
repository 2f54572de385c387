"""world_size-2 data-parallel logic on CPU (gloo): sharding a batch by example and SUM-all-reducing the per-rank
gradients reproduces the single-process full-batch gradients (the loss is a sum over examples, nn_utils/core.py:265-267),
and dropout masks keyed on global ids make the result independent of the number of ranks.  Gradients come from the
oracle here (test infrastructure); the GPU path runs the same plumbing with NCCL in bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from imagecaptionlearn_py_b200 import dist as D
from oracle import icl_oracle as O
from tests.helpers import tiny_problem


def test_shard_helpers():
    assert [D.shard_range(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    ids = ["im%d#0;mention:%d|b%d" % (i // 6, i % 3, i % 2) for i in range(24)]
    parts = [D.shard_ids(ids, r, 2, group_key=lambda s: s.split("#")[0]) for r in range(2)]
    assert sorted(parts[0] + parts[1]) == sorted(ids)
    assert not ({p.split("#")[0] for p in parts[0]} & {p.split("#")[0] for p in parts[1]})     # images never straddle ranks
    assert D.global_offsets(np.array([5, 7, 3]), 2) == 12


def _slice_problem(p, b, e):
    q = dict(p)
    q["x"], q["lens"] = p["x"][b:e], p["lens"][b:e]
    hb = {}
    for k, v in p["batch"].items():
        v = np.asarray(v)[b:e].copy()
        if v.ndim == 2 and v.shape[1] == 3 and k not in ("labels",) and "feats" not in k:
            v[:, 1] -= b                                     # sentence index is local to the shard
        hb[k] = v
    hb["sentences"], hb["seq_lengths"] = q["x"], q["lens"].astype(np.float64)
    q["batch"] = hb
    q["masks"] = dict(in_fw=p["masks"]["in_fw"][b:e], in_bw=p["masks"]["in_bw"][b:e], out_fw=p["masks"]["out_fw"][b:e],
                      out_bw=p["masks"]["out_bw"][b:e], heads=[[m[b:e] for m in p["masks"]["heads"][0]]])
    return q


def _shim_loss_weight(weighted):
    """What core.Session hands icl_set_loss_weights for a one-head graph under the live process group (the C library is replaced
    by a recorder: no GPU here)."""
    from imagecaptionlearn_py_b200 import _cabi, core

    class Rec(object):
        w = None

        def icl_set_loss_weights(self, handle, arr):
            Rec.w = None if arr is None else [float(arr[i]) for i in range(1)]
            return 0
    g = core.Graph()
    g.lstm = dict(n_hidden=4, data_norm=False, E=5)
    g.heads = [dict(task="card", weighted=weighted)]
    sess = core.Session(graph=g, dist=True)
    real = _cabi.lib
    _cabi.lib = lambda: Rec()
    try:
        sess._set_loss_weights(None)
    finally:
        _cabi.lib = real
    return 1.0 if Rec.w is None else Rec.w[0]


def _worker(rank, world, port, ret, weighted=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    p = tiny_problem(seed=31, S=10, T=6, E=5, H=4, F=3, task="card", act="tanh", dropout=True, weighted=weighted)
    b, e = D.shard_range(p["S"], rank, world)
    q = _slice_problem(p, b, e)
    f = O.model_forward(q["params"], q["cfg"], q["x"], q["lens"], [q["batch"]], q["keep_in"], q["keep"], q["masks"])
    g = O.model_backward(q["params"], q["cfg"], f, [q["batch"]])
    names = sorted(g)
    lw = _shim_loss_weight(weighted)                  # d joint / d loss of the head, as the shim sets it on the device
    flat = torch.from_numpy(np.concatenate([g[n].ravel() for n in names]) * lw)
    D.allreduce_sum_(flat)
    loss = torch.tensor([float(f["loss"]) * lw], dtype=torch.float64)
    td.all_reduce(loss)
    if rank == 0:
        ret["flat"], ret["loss"], ret["names"] = flat.numpy().copy(), float(loss), names
    td.barrier()
    td.destroy_process_group()


def test_two_rank_sum_allreduce_equals_full_batch():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    p = tiny_problem(seed=31, S=10, T=6, E=5, H=4, F=3, task="card", act="tanh", dropout=True)
    f = O.model_forward(p["params"], p["cfg"], p["x"], p["lens"], [p["batch"]], p["keep_in"], p["keep"], p["masks"])
    g = O.model_backward(p["params"], p["cfg"], f, [p["batch"]])
    full = np.concatenate([g[n].ravel() for n in ret["names"]])
    np.testing.assert_allclose(ret["flat"], full, rtol=1e-10, atol=1e-12)
    assert abs(ret["loss"] - float(f["loss"])) < 1e-10


def test_two_rank_weighted_classes_equals_full_batch_mean():
    """weighted_classes as executed = MEAN cross-entropy (core.py:244-267): each rank's gradient is of its local mean, so the shim
    gives that head the loss weight 1/world_size and the SUM all-reduce then equals the gradient of the global-batch mean."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret, True), nprocs=2, join=True)
    p = tiny_problem(seed=31, S=10, T=6, E=5, H=4, F=3, task="card", act="tanh", dropout=True, weighted=True)
    f = O.model_forward(p["params"], p["cfg"], p["x"], p["lens"], [p["batch"]], p["keep_in"], p["keep"], p["masks"])
    g = O.model_backward(p["params"], p["cfg"], f, [p["batch"]])
    full = np.concatenate([g[n].ravel() for n in ret["names"]])
    np.testing.assert_allclose(ret["flat"], full, rtol=1e-10, atol=1e-12)
    assert abs(ret["loss"] - float(f["loss"])) < 1e-10
