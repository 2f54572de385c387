"""Parity at BASELINE.json's full size (configs[1]: 2048 captions per step, H = 300, T padded to 50) through properties that do
not depend on the size, plus one direct comparison with the oracle (the NumPy oracle needs a few seconds for 2048 captions)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _session(B, wl, weights=None):
    from imagecaptionlearn_py_b200 import core
    core.reset_default_graph()
    core.set_random_seeds()
    with core.variable_scope("bidirectional_lstm"):
        core.setup_bidirectional_lstm(wl["H"], wl["data_norm"], n_embedding_width=300)
    core.setup_core_architecture(wl["task"], "first_last_mention", B, wl["start"], wl["depth"], False, "relu", wl["C"], wl["F"])
    core.add_train_op(core.get_collection("loss")[0], 1e-3, 1e-8, 5.0)
    sess = core.Session(max_seq_len=50)
    sess.ensure()
    if weights is None:
        sess.initialize()
    else:
        for k, v in weights.items():
            sess.set_tensor(k, v)
    return sess


def _take(bt, rows):
    out = {}
    for k, v in bt.items():
        out[k] = v[rows] if isinstance(v, np.ndarray) and v.shape[0] == len(bt["labels"]) else v
    return out


def test_card2048_properties_and_oracle():
    import bench
    from imagecaptionlearn_py_b200 import _cabi
    from oracle import icl_oracle as O
    wl = bench.WORKLOADS["card2048"]
    bt = bench.make_batch(wl, 20171201)                   # nonvis/card: sentence i belongs to example i, so rows permute together
    B = wl["B"]
    sess = _session(B, wl)
    weights = {n: sess.get_tensor(n) for n, _, _, _ in sess.param_info()}
    full = sess.run(_cabi.OP_GRADS, [bt], 1.0, 1.0, True)[0]
    g_full = {n: sess.get_tensor(n, 1) for n in weights}

    # (1) permutation equivariance: examples are independent, so a shuffled batch gives the shuffled probabilities BIT-exactly
    # (the rows move to other 128-row tiles and other ranks of the length-sorted order) and the same summed loss
    perm = np.random.default_rng(0).permutation(B)
    btp = _take(bt, perm)
    for k in ("first_i_fw", "first_i_bw", "last_i_fw", "last_i_bw", "sent_last_i_fw", "sent_first_i_bw", "sent_last_j_fw",
              "sent_first_j_bw", "first_j_fw", "first_j_bw", "last_j_fw", "last_j_bw"):
        btp[k] = btp[k].copy()
        btp[k][:, 1] = np.arange(B)                       # example r now sits in sentence row r
    shuf = sess.run(_cabi.OP_PREDICT, [btp], 1.0, 1.0, True)[0]
    assert np.array_equal(shuf["proba"], full["proba"][perm])
    assert np.array_equal(shuf["pred"], full["pred"][perm])
    assert abs(shuf["loss"] - full["loss"]) <= 2e-5 * abs(full["loss"])
    sess.close()

    # (2) linearity over the batch: the loss is a SUM over examples (nn_utils/core.py:267), so the gradient of the batch is the
    # sum of the gradients of its halves (this is also what the data-parallel all-reduce relies on)
    half = _session(B // 2, wl, weights)
    acc = {n: np.zeros_like(v) for n, v in weights.items()}
    loss = 0.0
    for lo in (0, B // 2):
        rows = np.arange(lo, lo + B // 2)
        bh = _take(bt, rows)
        for k in bh:
            if isinstance(bh[k], np.ndarray) and bh[k].ndim == 2 and bh[k].shape[1] == 3 and k.split("_")[0] in ("first", "last", "sent"):
                bh[k] = bh[k].copy()
                bh[k][:, 1] -= lo
        r = half.run(_cabi.OP_GRADS, [bh], 1.0, 1.0, True)[0]
        loss += float(r["loss"])
        for n in acc:
            acc[n] += half.get_tensor(n, 1)
    half.close()
    assert abs(loss - full["loss"]) <= 2e-5 * abs(full["loss"])
    for n in acc:
        scale = max(float(np.max(np.abs(g_full[n]))), 1e-30)
        assert float(np.max(np.abs(acc[n] - g_full[n]))) / scale < 1e-3, n       # TF32 rounding of dZ differs per tile placement

    # (3) the oracle itself on all 2048 captions: loss, probabilities and the LSTM kernel gradients
    widths = [wl["start"] // (2 ** i) for i in range(wl["depth"] + 1)]
    cfg = dict(H=wl["H"], data_norm=wl["data_norm"],
               heads=[dict(task=wl["task"], scope="", encoding_scheme="first_last_mention", n_layers=len(widths), widths=widths,
                           activation="relu", weighted_classes=False, in_width=4 * wl["H"] + wl["F"], n_classes=wl["C"])])
    params = {k: (v.astype(np.float64).reshape(-1) if k.endswith("basic_lstm_cell/bias") else v.astype(np.float64)) for k, v in weights.items()}
    tl = int(bt["seq_lengths"].max())
    f = O.model_forward(params, cfg, bt["sentences"][:, :tl].astype(np.float64), bt["seq_lengths"], [bt])
    assert abs(full["loss"] - f["loss"]) < 1e-3 * abs(f["loss"])
    assert float(np.max(np.abs(full["proba"] - f["heads"][0]["proba"]))) < 1e-3
    agree = np.mean(full["pred"] == f["heads"][0]["pred"])
    margin = np.sort(f["heads"][0]["proba"], 1)
    clear = (margin[:, -1] - margin[:, -2]) > 2e-3
    assert np.array_equal(full["pred"][clear], f["heads"][0]["pred"][clear]) and agree > 0.99
