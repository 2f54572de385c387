"""Golden fixtures: (1) values produced by the reference's own still-importable modules (utils/Score.py, utils/string.py);
(2) known-answer vectors of the oracle (regression anchors -- the TF graph itself cannot run here: parity unpinned)."""
import json
import os

import numpy as np

from oracle import icl_oracle as O
from tests.helpers import tiny_problem

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_score_matches_reference_module_outputs():
    from imagecaptionlearn_py_b200.eval import Score
    from imagecaptionlearn_py_b200.data import kv_str_to_dict
    gold = json.load(open(os.path.join(HERE, "score_golden.json")))
    for c in gold["score"]:
        if "pred" in c:
            s = Score(predicted_count=c["pred"], gold_count=c["gold"], correct_count=c["correct"])
        else:
            s = Score(precision=c["precision"], recall=c["recall"])
        assert (s.p, s.r, s.f1) == (c["p"], c["r"], c["f1"])          # bit-exact
        assert s.to_string() == c["text"] and s.to_latex_string() == c["latex"]
    for c in gold["kv"]:
        assert kv_str_to_dict(c["s"]) == c["d"]


def test_oracle_known_answers():
    kat = np.load(os.path.join(HERE, "oracle_kat.npz"))
    for name, kw in (("nonvis", dict(seed=11, task="nonvis", act="relu", dropout=True)),
                     ("card_sent", dict(seed=12, task="card", enc="first_last_sentence", act="tanh", data_norm=True, S=5, T=6)),
                     ("rel_cross", dict(seed=13, task="rel_cross", enc="first_last_sentence", act="sigmoid", weighted=True, S=8)),
                     ("affinity", dict(seed=14, task="affinity", act="leaky_relu", box_w=6, dropout=True))):
        p = tiny_problem(**kw)
        f = O.model_forward(p["params"], p["cfg"], p["x"], p["lens"], [p["batch"]], p["keep_in"], p["keep"], p["masks"])
        g = O.model_backward(p["params"], p["cfg"], f, [p["batch"]])
        assert abs(float(f["loss"]) - float(kat[name + "/loss"])) < 1e-12 * max(1.0, abs(float(f["loss"])))
        np.testing.assert_allclose(f["heads"][0]["proba"], kat[name + "/proba"], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(f["out_fw"], kat[name + "/out_fw"], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(f["out_bw"], kat[name + "/out_bw"], rtol=1e-12, atol=1e-14)
        for k, v in g.items():
            np.testing.assert_allclose(v, kat[name + "/grad/" + k], rtol=1e-10, atol=1e-13)


def test_evaluate_multiclass_and_scores_file(tmp_path):
    from imagecaptionlearn_py_b200 import eval as ev
    from sklearn import metrics as skm
    rng = np.random.default_rng(0)
    gold, pred = rng.integers(0, 4, 200), rng.integers(0, 4, 200)
    sd = ev.evaluate_multiclass(gold, pred, ["n", "c", "b", "p"])
    # nn_utils/eval.py:121-130 reports sklearn's per-class numbers: ours must agree with them
    p = skm.precision_score(gold, pred, average=None, zero_division=0)
    r = skm.recall_score(gold, pred, average=None, zero_division=0)
    f = skm.f1_score(gold, pred, average=None, zero_division=0)
    for l in range(4):
        s = sd.get_score(l)
        assert abs(s.p - p[l]) < 1e-12 and abs(s.r - r[l]) < 1e-12 and abs(s.f1 - f[l]) < 1e-12
    assert np.array_equal(sd.confusion_matrix, skm.confusion_matrix(gold, pred, labels=range(4)))
    assert abs(sd.accuracy - 100.0 * skm.accuracy_score(gold, pred)) < 1e-12
    path = str(tmp_path / "scores.txt")
    ev.write_scores_file(path, {"a;mention:0": np.array([0.25, 0.75]), "b;mention:1": np.array([0.0, 1.0])})
    lines = open(path).read().strip().split("\n")
    assert lines[0] == "a;mention:0,%s,%s" % (np.log(0.25), np.log(0.75))
    assert lines[1].split(",")[1] == str(np.log(np.nextafter(0, 1))) and lines[1].split(",")[2] == "0.0"
