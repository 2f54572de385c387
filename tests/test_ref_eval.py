"""Pins eval.ScoreDict / evaluate_relations / evaluate_multiclass / rows_to_str to outputs of the
reference's own utils/ScoreDict.py, nn_utils/eval.py and utils/string.py (tests/golden/make_ref_eval.py ran them)."""
import json
import os

import pytest

from imagecaptionlearn_py_b200 import eval as E

REF = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_eval.json")))


def dump(sd, labels):
    return dict(labels=[str(l) for l in labels],
                scores=[[sd.get_score(l).p, sd.get_score(l).r, sd.get_score(l).f1] for l in labels],
                gold=[sd.get_gold_count(l) for l in labels], pred=[sd.get_pred_count(l) for l in labels],
                gold_total=sd.get_gold_count(), pred_total=sd.get_pred_count(), accuracy=sd.get_accuracy(),
                correct=sd.get_correct_count(),
                gold_percent=[sd.get_gold_percent(l) for l in labels], pred_percent=[sd.get_pred_percent(l) for l in labels])


@pytest.mark.parametrize("case", REF["score_dict"], ids=lambda c: "C%d" % len(c["dump"]["labels"]))
def test_score_dict(case):
    C = len(case["dump"]["labels"])
    sd = E.ScoreDict(case["gold"], case["pred"])
    assert dump(sd, list(range(C))) == case["dump"]                      # floats compare exactly: same arithmetic
    assert [int(k) for k in sd.keys] == case["key_order"]
    assert [l.rstrip() for l in sd.confusion_rows()] == [l.rstrip() for l in case["confusion"]]
    n = len(case["gold"])
    merged = E.ScoreDict(case["gold"], case["pred"])
    merged.merge(E.ScoreDict(case["gold"][: n // 3], case["pred"][n // 3: 2 * (n // 3)]))
    assert dump(merged, list(range(C))) == case["merged"]


@pytest.mark.parametrize("case", REF["relations"], ids=lambda c: "n%d" % len(c["gold"]))
def test_evaluate_relations(case):
    gold = {(a, b): v for a, b, v in case["gold"]}
    sd = E.evaluate_relations(case["pairs"], case["pred"], gold, None)
    labels = ["-invalid-", "invalid", "-reverse_sub-", "null", "coref", "subset"]
    assert dump(sd, labels) == case["dump"]
    assert sd.summary_lines == case["printed_head"]


@pytest.mark.parametrize("case", REF["multiclass"], ids=lambda c: "C%d" % len(c["names"]))
def test_evaluate_multiclass(case):
    class Log(object):
        lines = []

        def info(self, msg):
            self.lines.append(msg)
    log = Log()
    log.lines = []
    sd = E.evaluate_multiclass(case["gold"], case["pred"], case["names"], log)
    assert dump(sd, list(range(len(case["names"])))) == case["dump"]
    assert log.lines == case["log"]


@pytest.mark.parametrize("case", REF["rows_to_str"], ids=lambda c: "r%d%s" % (len(c["rows"]), "L" if c["use_latex"] else ""))
def test_rows_to_str(case):
    assert E.rows_to_str(case["rows"], case["has_headers"], case["use_latex"]) == case["text"]
