"""Generates tests/golden/ref_eval.json by running the REAL reference evaluation code from /root/reference:

    utils/ScoreDict.py:13-221   ScoreDict (increment, merge, get_score, counts, percents, accuracy, print_confusion)
    nn_utils/eval.py:10-93      evaluate_relations
    nn_utils/eval.py:95-184     evaluate_multiclass (needs scikit-learn, which is installed)
    utils/string.py:33-128      rows_to_str

imported through tests/golden/ref_import.py (the only change: Python-2 `print x` statements become `print(x)` in memory).
Run in the build container:   python tests/golden/make_ref_eval.py
"""
import contextlib
import io
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


class RecLog(object):
    def __init__(self):
        self.lines = []

    def info(self, *a):
        self.lines.append(a[0] if len(a) == 1 else (a[1] % tuple(a[2:]) if a[0] is None else a[0] % tuple(a[1:])))


def sd_dump(sd, labels):
    return dict(labels=[str(l) for l in labels],
                scores=[[sd.get_score(l).p, sd.get_score(l).r, sd.get_score(l).f1] for l in labels],
                gold=[sd.get_gold_count(l) for l in labels], pred=[sd.get_pred_count(l) for l in labels],
                gold_total=sd.get_gold_count(), pred_total=sd.get_pred_count(), accuracy=sd.get_accuracy(),
                correct=sd.get_correct_count(),
                gold_percent=[sd.get_gold_percent(l) for l in labels], pred_percent=[sd.get_pred_percent(l) for l in labels])


def relation_case(rng, n_docs, labels_p):
    """Synthetic (ij, ji) predictions + gold, including inconsistent link pairs."""
    names = ["null", "coref", "subset_ij", "subset_ji"]
    gold, pairs, pred = {}, [], []
    for d in range(n_docs):
        for a in range(4):
            for b in range(a + 1, 4):
                ij = "doc:%d.jpg;caption_1:0;mention_1:%d;caption_2:0;mention_2:%d" % (d, a, b)
                ji = "doc:%d.jpg;caption_1:0;mention_1:%d;caption_2:0;mention_2:%d" % (d, b, a)
                gold[(ij, ji)] = names[int(rng.choice(4, p=labels_p))]
                if rng.rand() < 0.9:           # some pairs are never predicted
                    pairs += [ij, ji]
                    l = int(rng.choice(4, p=labels_p))
                    if rng.rand() < 0.7:
                        pred += [l, {0: 0, 1: 1, 2: 3, 3: 2}[l]]          # consistent
                    else:
                        pred += [l, int(rng.choice(4))]                   # possibly inconsistent
    return pairs, pred, gold


def main():
    import ref_import
    mods = ref_import.install(tensorflow=types.ModuleType("tensorflow"))
    SD, ev, su = mods["utils.ScoreDict"].ScoreDict, mods["nn_utils.eval"], mods["utils.string"]
    rng = np.random.RandomState(20171201)
    out = dict(source="/root/reference utils/ScoreDict.py, nn_utils/eval.py, utils/string.py (imported; prints rewritten in memory)",
               score_dict=[], relations=[], multiclass=[], rows_to_str=[])
    for C, n in ((2, 50), (4, 200), (12, 500)):
        g = [int(x) for x in rng.randint(0, C, n)]
        p = [int(x) if rng.rand() < 0.6 else int(rng.randint(0, C)) for x in g]
        sd = SD(g, p)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            sd.print_confusion()
        other = SD(g[: n // 3], p[n // 3: 2 * (n // 3)])
        merged = SD(g, p)
        merged.merge(other)
        out["score_dict"].append(dict(gold=g, pred=p, dump=sd_dump(sd, list(range(C))), confusion=buf.getvalue().rstrip("\n").split("\n"),
                                      key_order=[int(k) for k in sd.keys], merged=sd_dump(merged, list(range(C)))))
    for n_docs, lp in ((3, [0.5, 0.3, 0.1, 0.1]), (25, [0.4, 0.2, 0.2, 0.2]), (10, [0.0, 0.0, 0.5, 0.5])):
        pairs, pred, gold = relation_case(rng, n_docs, lp)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            sd = ev.evaluate_relations(pairs, pred, gold, None)
        labels = ["-invalid-", "invalid", "-reverse_sub-", "null", "coref", "subset"]
        out["relations"].append(dict(pairs=pairs, pred=pred, gold=[[k[0], k[1], v] for k, v in gold.items()],
                                     dump=sd_dump(sd, labels), printed_head=buf.getvalue().split("\n")[:10]))
    for C, n, names in ((2, 60, ["v", "n"]), (12, 400, [str(i) for i in range(11)] + ["11+"]), (2, 40, ["0", "1"])):
        g = [int(x) for x in rng.randint(0, C, n)]
        p = [int(x) if rng.rand() < 0.7 else int(rng.randint(0, C)) for x in g]
        g[0], p[0] = C - 1, C - 1                   # the reference indexes bincount(gold)[C-1]: the top class must occur
        log = RecLog()
        sd = ev.evaluate_multiclass(g, p, names, log)
        out["multiclass"].append(dict(gold=g, pred=p, names=names, dump=sd_dump(sd, list(range(C))), log=log.lines))
    for rows, hdr, latex in (([["", "P", "R"], ["a", "1.00%", "22.00%"], ["bcd", "3", "4"]], True, False),
                             ([["gold", "pred"], ["10", "2"], ["7", "15"]], False, False),
                             ([["x", "y"], ["1", "2"]], False, True), ([["h", "c1"], ["r1", "v"], ["r2"]], True, False)):
        out["rows_to_str"].append(dict(rows=rows, has_headers=hdr, use_latex=latex, text=su.rows_to_str(rows, hdr, latex)))
    json.dump(out, open(os.path.join(HERE, "ref_eval.json"), "w"), indent=0)
    ref_import.uninstall()
    print("wrote ref_eval.json")


if __name__ == "__main__":
    main()
