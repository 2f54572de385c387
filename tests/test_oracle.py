"""Pins the NumPy oracle: autograd restatement, torch.nn.LSTM mapping, finite differences, TF-Adam rule."""
import numpy as np
import pytest
import torch

from oracle import icl_oracle as O
from oracle import torch_check as TC
from tests.helpers import tiny_problem


def test_get_widths():
    assert O.get_widths(512, 2) == [512, 256, 128]          # CLI defaults -> d+1 layers (core.py:136)
    assert O.get_widths(1024, 3) == [1024, 512, 256, 128]   # config/lstm_intra_params.config:96
    assert O.get_widths(100, 3, 10) == [100, 55, 32, 10]


def test_slot_plan_column_order():
    assert O.slot_plan("nonvis", "first_last_mention") == ["first_i_bw", "last_i_fw", "first_i_fw", "last_i_bw", "m_feats"]
    assert O.slot_plan("rel_intra", "first_last_mention") == [
        "first_i_bw", "last_i_fw", "first_i_fw", "last_i_bw", "first_j_bw", "last_j_fw", "ij_feats", "first_j_fw", "last_j_bw"]
    assert O.slot_plan("rel_intra", "first_last_sentence") == [
        "first_i_bw", "last_i_fw", "sent_last_i_fw", "sent_first_i_bw", "first_j_bw", "last_j_fw", "ij_feats"]
    assert O.slot_plan("rel_cross", "first_last_sentence")[-2:] == ["sent_last_j_fw", "sent_first_j_bw"]
    assert O.slot_plan("affinity", "first_last_mention")[-2:] == ["m_feats", "box_embeddings"]


@pytest.mark.parametrize("task,enc,act,dropout,norm,weighted", [
    ("nonvis", "first_last_mention", "relu", False, False, False),
    ("card", "first_last_mention", "tanh", True, True, False),
    ("rel_intra", "first_last_mention", "leaky_relu", True, False, False),
    ("rel_cross", "first_last_sentence", "sigmoid", True, True, True),
    ("affinity", "first_last_sentence", "relu", True, False, False),
])
def test_numpy_backward_matches_autograd(task, enc, act, dropout, norm, weighted):
    p = tiny_problem(seed=3, task=task, enc=enc, act=act, dropout=dropout, data_norm=norm, weighted=weighted,
                     box_w=6 if task == "affinity" else 0)
    fwd = O.model_forward(p["params"], p["cfg"], p["x"], p["lens"], [p["batch"]], p["keep_in"], p["keep"], p["masks"])
    g = O.model_backward(p["params"], p["cfg"], fwd, [p["batch"]])
    loss, tg, probas, outs = TC.grads_via_autograd(p["params"], p["cfg"], p["x"], p["lens"], [p["batch"]],
                                                   p["keep_in"], p["keep"], p["masks"])
    assert abs(loss - fwd["loss"]) < 1e-9 * max(1, abs(loss))
    np.testing.assert_allclose(fwd["out_fw"], outs["fw"], atol=1e-12)
    np.testing.assert_allclose(fwd["out_bw"], outs["bw"], atol=1e-12)
    np.testing.assert_allclose(fwd["heads"][0]["proba"], probas[0], atol=1e-12)
    assert set(g) == set(tg)
    for k in g:
        np.testing.assert_allclose(g[k], tg[k], atol=1e-9, rtol=1e-8, err_msg=k)


def test_bilstm_matches_torch_nn_lstm():
    """TF kernel [E+H,4H] with gate columns i,j,f,o and forget_bias 1.0 mapped onto torch's i,f,g,o rows."""
    p = tiny_problem(seed=5, S=5, T=9, E=4, H=6)
    fw, bw, _ = O.bilstm_forward(p["params"], p["x"], p["lens"])
    E, H = 4, 6
    lstm = torch.nn.LSTM(E, H, batch_first=True, bidirectional=True).double()
    with torch.no_grad():
        for d, suf in (("fw", ""), ("bw", "_reverse")):
            kn, bn = O.lstm_names(d)
            K, b = p["params"][kn], p["params"][bn].copy()
            i, j, f, o = np.split(K, 4, 1)
            bi, bj, bf, bo = np.split(b, 4)
            Kt = np.concatenate([i, f, j, o], 1)
            bt = np.concatenate([bi, bf + 1.0, bj, bo])
            getattr(lstm, "weight_ih_l0" + suf).copy_(torch.tensor(Kt[:E].T))
            getattr(lstm, "weight_hh_l0" + suf).copy_(torch.tensor(Kt[E:].T))
            getattr(lstm, "bias_ih_l0" + suf).copy_(torch.tensor(bt))
            getattr(lstm, "bias_hh_l0" + suf).zero_()
        packed = torch.nn.utils.rnn.pack_padded_sequence(torch.tensor(p["x"]), torch.tensor(p["lens"]),
                                                         batch_first=True, enforce_sorted=False)
        out, _ = lstm(packed)
        out, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=p["T"])
    np.testing.assert_allclose(fw, out[:, :, :H].numpy(), atol=1e-12)
    np.testing.assert_allclose(bw, out[:, :, H:].numpy(), atol=1e-12)


def test_finite_difference_on_lstm_kernel():
    p = tiny_problem(seed=7, S=3, T=4, E=3, H=2, widths=(4,))
    fwd = O.model_forward(p["params"], p["cfg"], p["x"], p["lens"], [p["batch"]])
    g = O.model_backward(p["params"], p["cfg"], fwd, [p["batch"]])
    rng = np.random.default_rng(0)
    for name in (O.lstm_names("fw")[0], O.lstm_names("bw")[0], "hdn_1/Variable"):
        for _ in range(4):
            idx = tuple(rng.integers(0, s) for s in p["params"][name].shape)
            old = p["params"][name][idx]
            eps = 1e-6
            p["params"][name][idx] = old + eps
            lp = O.model_forward(p["params"], p["cfg"], p["x"], p["lens"], [p["batch"]])["loss"]
            p["params"][name][idx] = old - eps
            lm = O.model_forward(p["params"], p["cfg"], p["x"], p["lens"], [p["batch"]])["loss"]
            p["params"][name][idx] = old
            assert abs((lp - lm) / (2 * eps) - g[name][idx]) < 1e-6


def test_outputs_zero_past_length_and_bw_semantics():
    p = tiny_problem(seed=9)
    fw, bw, _ = O.bilstm_forward(p["params"], p["x"], p["lens"])
    for s, L in enumerate(p["lens"]):
        assert np.all(fw[s, L:] == 0) and np.all(bw[s, L:] == 0)
    # bw on a sequence == fw with bw weights on the reversed valid prefix
    s = int(np.argmax(p["lens"]))
    L = p["lens"][s]
    kn, bn = O.lstm_names("bw")
    xr = p["x"][s:s + 1, :L][:, ::-1]
    h_all, _ = O._dir_forward(xr, [L], p["params"][kn], p["params"][bn], reverse=False)
    np.testing.assert_allclose(bw[s, :L], h_all[0, ::-1], atol=1e-13)


def test_tf_adam_and_clip_rule():
    rng = np.random.default_rng(1)
    params = {"w": rng.standard_normal((3, 2))}
    ref = {k: v.copy() for k, v in params.items()}
    state = {}
    m = np.zeros((3, 2)); v = np.zeros((3, 2))
    for t in range(1, 4):
        g = {"w": rng.standard_normal((3, 2)) * 10}
        gn = np.sqrt((g["w"] ** 2).sum())
        gc = g["w"] * 5.0 / max(gn, 5.0)
        m = 0.9 * m + 0.1 * gc
        v = 0.999 * v + 0.001 * gc * gc
        ref["w"] -= 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m / (np.sqrt(v) + 1e-8)
        O.clip_and_adam(params, g, state, 1e-3, 1e-8, 5.0)
        np.testing.assert_allclose(params["w"], ref["w"], atol=1e-15)
    # matches the torch restatement used for the CPU baseline
    tp = {"w": torch.tensor(np.ones((3, 2)), requires_grad=True)}
    opt = TC.TFAdam(tp, 1e-3, 1e-8, 5.0)
    pn = {"w": np.ones((3, 2))}
    st = {}
    for _ in range(3):
        g = rng.standard_normal((3, 2)) * 10
        tp["w"].grad = torch.tensor(g.copy())
        opt.step()
        O.clip_and_adam(pn, {"w": g}, st, 1e-3, 1e-8, 5.0)
    np.testing.assert_allclose(tp["w"].detach().numpy(), pn["w"], atol=1e-14)


def test_weighted_classes_is_mean_ce_as_executed():
    p = tiny_problem(seed=11, weighted=True)
    q = tiny_problem(seed=11, weighted=False)
    lw = O.model_forward(p["params"], p["cfg"], p["x"], p["lens"], [p["batch"]])["loss"]
    ls = O.model_forward(q["params"], q["cfg"], q["x"], q["lens"], [q["batch"]])["loss"]
    assert abs(lw * p["B"] - ls) < 1e-10


def test_multihead_joint_loss_is_sum():
    p = tiny_problem(seed=13, task="nonvis")
    q = tiny_problem(seed=13, task="card")
    cfg = dict(H=p["H"], data_norm=False, heads=[dict(p["cfg"]["heads"][0], scope="nonvis"),
                                                  dict(q["cfg"]["heads"][0], scope="card")])
    rng = np.random.default_rng(0)
    params = O.init_params(rng, cfg, p["E"])
    f = O.model_forward(params, cfg, p["x"], p["lens"], [p["batch"], q["batch"]])
    assert abs(f["loss"] - (f["heads"][0]["loss"] + f["heads"][1]["loss"])) < 1e-12
    g = O.model_backward(params, cfg, f, [p["batch"], q["batch"]])
    _, tg, _, _ = TC.grads_via_autograd(params, cfg, p["x"], p["lens"], [p["batch"], q["batch"]])
    for k in g:
        np.testing.assert_allclose(g[k], tg[k], atol=1e-9, err_msg=k)
