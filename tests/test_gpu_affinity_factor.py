"""Affinity layer 1, factorised (the concat it replaces: nn_utils/core.py:421-433,439).

batch_input of a mention-box pair is [mention encoding | m_feats | box row], so the first hidden layer's pre-activation is
z1 = U[mention of the pair] + V[box of the pair] + b1 with U = [enc, m_feats] W1[0:Dm] once per DISTINCT mention of the batch and
V = box W1[Dm:] once per distinct box.  The library groups the pairs on the host (exact comparison of index rows / feature rows /
box rows), runs layer 1 and its backward per group, and must give what the concatenated formulation gives: the oracle
(oracle/icl_oracle.py restates the concat), at the same tolerances as every other head -- 2e-5 / 5e-5 in the fp32 validation mode,
1e-3 / 2e-3 in product mode -- with and without dropout, for both encodings, an unaligned feature width, host box rows and
rows of the resident box table, and forced on a batch without any repetition (ICL_AFF_FACTOR=2: every pair its own group)."""
import ctypes as C

import numpy as np
import pytest

from oracle import icl_oracle as O
from tests.helpers import tiny_problem
from tests.test_gpu_parity import TOL, device_masks, kink_override, make_session, relerr

pytestmark = pytest.mark.gpu

CASES = {
    "flm-H8": dict(task="affinity", enc="first_last_mention", act="relu", S=60, T=8, E=12, H=8, F=4, widths=(16, 8), box_w=64),
    "fls-H8-tanh": dict(task="affinity", enc="first_last_sentence", act="tanh", S=45, T=8, E=12, H=8, F=4, widths=(16, 8), box_w=64),
    "flm-H300-F33": dict(task="affinity", enc="first_last_mention", act="relu", S=130, T=12, E=300, H=300, F=33, widths=(64, 32), box_w=4096),
}
INDEX_KEYS = ("first_i_fw", "first_i_bw", "last_i_fw", "last_i_bw", "sent_last_i_fw", "sent_first_i_bw")


def paired_problem(seed, n_m, n_b, dropout=False, **case):
    """tiny_problem's affinity batch rewritten as mention x box pairs: pair p = (mention p % n_m, box (p // n_m) % n_b), so that its
    B pairs hold n_m distinct mentions (index rows + m_feats row) and n_b distinct box rows, interleaved."""
    p = tiny_problem(seed=seed, dropout=dropout, **case)
    hb, B = p["batch"], p["B"]
    mi, bi = np.arange(B) % n_m, (np.arange(B) // n_m) % n_b
    if n_m >= B:                  # no mention repeats (a reference-shaped training batch: one caption copy per pair): boxes cycle
        bi = np.arange(B) % n_b
    for k in INDEX_KEYS:
        hb[k] = hb[k][mi].copy()
    hb["m_feats"] = hb["m_feats"][mi].copy()
    hb["box_embeddings"] = hb["box_embeddings"][bi].copy()
    return p, len(set(mi.tolist())), len(set(bi.tolist()))


def factor_stats(sess, head=0):
    from imagecaptionlearn_py_b200 import _cabi
    f, nm, nb = C.c_int32(), C.c_int32(), C.c_int32()
    tot = (C.c_int64 * 3)()
    _cabi.check(_cabi.lib().icl_head_factor_stats(sess.handle, head, C.byref(f), C.byref(nm), C.byref(nb), tot))
    return bool(f.value), nm.value, nb.value, list(tot)


def check_against_oracle(p, mode, dropout, expect):
    from imagecaptionlearn_py_b200 import _cabi
    core, sess = make_session(p, mode)
    r = sess.run(_cabi.OP_GRADS, [dict(p["batch"])], p["keep_in"], p["keep"], True)[0]
    assert factor_stats(sess)[:3] == expect, (factor_stats(sess), expect)
    masks = device_masks(sess, p) if dropout else None
    f = O.model_forward(p["params"], p["cfg"], p["x"], p["lens"], [p["batch"]], p["keep_in"], p["keep"], masks)
    tol = TOL[mode]
    assert relerr(r["proba"], f["heads"][0]["proba"]) < tol["fwd"]
    assert abs(r["loss"] - f["loss"]) < tol["fwd"] * max(1.0, abs(f["loss"]))
    # the concatenated rows a factorised head never builds are still what icl_get_batch_input hands out
    bi = np.empty((p["B"], p["cfg"]["heads"][0]["in_width"]), np.float32)
    _cabi.check(_cabi.lib().icl_get_batch_input(sess.handle, 0, _cabi.np_ptr(bi)))
    assert relerr(bi, f["heads"][0]["_bwd"][0][0][0]) < tol["fwd"]
    over, flips = kink_override(sess, p, f, masks, tol["fwd"])
    if mode == "simt":
        assert flips == 0
    g = O.model_backward(p["params"], p["cfg"], f, [p["batch"]], over)
    worst = {k: relerr(sess.get_tensor(k, 1).reshape(v.shape), v) for k, v in g.items()}
    toy_bias = 5e-3 if (mode == "tf32" and p["H"] < 100) else 0.0
    bad = {k: v for k, v in worst.items() if v > max(tol["grad"], toy_bias if k.endswith(("bias", "Variable_1")) else 0.0)}
    # the softmax bias gradient is sum_b (p_b - y_b): with a handful of distinct boxes and random labels the terms cancel almost
    # completely, so its error is judged on the scale of the terms it adds (each one is right to tol["fwd"]), not of what is left
    k = "softmax/Variable_1"
    if k in bad:
        terms = np.sum(np.abs(f["heads"][0]["proba"] - p["batch"]["labels"]), 0).max()
        if np.max(np.abs(sess.get_tensor(k, 1).reshape(g[k].shape) - g[k])) <= tol["fwd"] * terms:
            del bad[k]
    assert not bad, bad
    sess.close()
    return worst


@pytest.mark.parametrize("mode", ["simt", "tf32"])
@pytest.mark.parametrize("dropout", [False, True], ids=["nodrop", "drop"])
@pytest.mark.parametrize("case", list(CASES), ids=list(CASES))
def test_factorised_layer1_matches_oracle(case, mode, dropout):
    p, n_m, n_b = paired_problem(41, 7, 5, dropout=dropout, **CASES[case])
    check_against_oracle(p, mode, dropout, (True, n_m, n_b))


@pytest.mark.parametrize("mode", ["simt", "tf32"])
@pytest.mark.parametrize("case", ["flm-H8", "flm-H300-F33"])
def test_factorised_box_half_only_matches_oracle(case, mode):
    """The shape of a reference training batch: every pair its own caption copy (no two mentions alike), 9 distinct boxes."""
    p, n_m, n_b = paired_problem(42, 10 ** 6, 9, dropout=True, **CASES[case])
    assert n_m == p["B"]
    check_against_oracle(p, mode, True, (True, n_m, n_b))


@pytest.mark.parametrize("mode", ["simt", "tf32"])
def test_forced_factorisation_without_repeats_matches_oracle(monkeypatch, mode):
    """Every pair its own mention and box group (the worst case of the formulation): forced with ICL_AFF_FACTOR=2; left alone the
    same batch takes the concatenated path (nothing to gain)."""
    p = tiny_problem(seed=43, dropout=True, **CASES["flm-H8"])
    check_against_oracle(p, mode, True, (False, 0, 0))
    monkeypatch.setenv("ICL_AFF_FACTOR", "2")
    check_against_oracle(p, mode, True, (True, p["B"], p["B"]))


def test_factorised_equals_concatenated_on_a_corpus_batch(monkeypatch):
    """A reference-shaped batch (data.load_batch on the synthetic corpus: all mention x box pairs of a few images) with host box rows,
    with rows of the resident box table, and with shared captions (dedup=True, the prediction path): factorised and concatenated
    (ICL_AFF_FACTOR=0) agree to fp32 rounding in the validation mode -- probabilities, predictions, every gradient -- and two
    factorised runs are bit-identical (ordered segment sums).  A reference-shaped batch carries one caption copy per pair
    (nn_utils/data.py:397-403), so only its boxes repeat; with shared captions the mentions collapse as well."""
    from imagecaptionlearn_py_b200 import _cabi, core, synth
    from imagecaptionlearn_py_b200 import data as nn_data
    corpus = synth.make_corpus(3, seed=29, E=12, with_boxes=True, box_width=32)
    dd = synth.make_data_dict(corpus, "affinity", F=8)
    ids = synth.example_ids(dd, "affinity")[:96]
    res = {}
    for factor in ("0", "1"):
        monkeypatch.setenv("ICL_AFF_FACTOR", factor)
        for packed in (False, "rows", "dedup"):
            bt = nn_data.load_batch(ids, dd, "affinity", 2, dedup=True) if packed == "dedup" else nn_data.load_batch(ids, dd, "affinity", 2, packed=packed)
            core.reset_default_graph()
            core.set_random_seeds()
            with core.variable_scope("bidirectional_lstm"):
                core.setup_bidirectional_lstm(8, False, n_embedding_width=12)
            core.setup_core_architecture("affinity", "first_last_mention", 96, 16, 1, False, "relu", 2, 8, box_embedding_width=32)
            core.add_train_op(core.get_collection("loss")[0], 1e-3, 1e-8, 5.0)
            sess = core.Session(max_seq_len=dd["max_seq_len"], gemm_mode=_cabi.GEMM_SIMT_FP32)
            sess.ensure()
            sess.initialize()
            runs = []
            for rep in range(2):
                sess.base_seed, sess.run_counter = 9, 0
                r = sess.run(_cabi.OP_GRADS, [bt], 0.5, 0.5, True)[0]
                runs.append((r["proba"].copy(), r["pred"].copy(), {n: sess.get_tensor(n, 1) for n, _, _, _ in sess.param_info()}))
            on, n_m, n_b, _ = factor_stats(sess)
            assert on == (factor == "1")
            if on:
                assert 0 < n_b and n_b * 2 <= 96 and (n_m < 96 if packed == "dedup" else n_m == 96), (packed, n_m, n_b)
                assert np.array_equal(runs[0][0], runs[1][0])
                for k, v in runs[0][2].items():
                    if "hdn_1" in k:                  # layer 1's own gradients: ordered sums, no atomics
                        assert np.array_equal(runs[1][2][k], v), k
            res[(factor, packed)] = runs[0]
            sess.close()
    for key, (proba, pred, grads) in res.items():
        ref = res[("0", "dedup" if key[1] == "dedup" else False)]      # shared captions draw other input-dropout masks: their own baseline
        assert relerr(proba, ref[0]) < 2e-6, key
        assert np.array_equal(pred, ref[1]), key
        for k, v in ref[2].items():
            assert relerr(grads[k], v) < 2e-5, (key, k, relerr(grads[k], v))
