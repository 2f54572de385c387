"""Evaluation helpers of the hot path's callers, restated for Python 3 (host-side; no kernels).

    Score, ScoreDict            utils/Score.py:4-48, utils/ScoreDict.py:13-221
    rows_to_str                 utils/string.py:33-128   (the table layout of the reference's log lines)
    evaluate_relations          nn_utils/eval.py:10-93   (pairwise ij / ji scoring of the relation heads)
    evaluate_multiclass         nn_utils/eval.py:95-184  (per-class P/R/F1, accuracy, confusion matrix)
    write_scores_file           icl_core_lstm.py:240-252 ("<id>,<ln p_0>,...", zero -> nextafter(0,1) before the log)

`utils/Score.py` of the reference imports and runs unchanged under Python 3; tests/golden/score_golden.json was generated
from it (tests/golden/make_golden.py) and pins `Score` here bit-for-bit.  ScoreDict, evaluate_relations and
evaluate_multiclass are pinned the same way by tests/golden/ref_eval.json (tests/golden/make_ref_eval.py runs the
reference's own nn_utils/eval.py and utils/ScoreDict.py, their `print` statements rewritten in memory).
"""
from collections import defaultdict

import numpy as np


class Score(object):
    def __init__(self, precision=0.0, recall=0.0, predicted_count=0, gold_count=0, correct_count=0):
        self.p, self.r = precision, recall
        if predicted_count > 0:
            self.p = float(correct_count) / float(predicted_count)
        if gold_count > 0:
            self.r = float(correct_count) / float(gold_count)
        self.f1 = 0.0
        if self.r > 0 and self.p > 0:
            self.f1 = (2 * self.p * self.r) / (self.p + self.r)

    def to_string(self):
        return "P: %6.2f%% | R: %6.2f%% | F1: %6.2f%%" % (100.0 * self.p, 100.0 * self.r, 100.0 * self.f1)

    def to_latex_string(self):
        return "%6.2f\\%% & %6.2f\\%% & %6.2f\\%% \\\\" % (100.0 * self.p, 100.0 * self.r, 100.0 * self.f1)


class ScoreDict(object):
    def __init__(self, gold_labels=None, pred_labels=None):
        self._gold_counts, self._pred_counts = defaultdict(int), defaultdict(int)
        self._correct_counts, self._confusion = defaultdict(int), defaultdict(int)
        self.keys = set()
        if gold_labels is not None and pred_labels is not None:
            for g, p in zip(gold_labels, pred_labels):
                self.increment(g, p)

    def increment(self, gold_label, pred_label):
        self._gold_counts[gold_label] += 1
        self._pred_counts[pred_label] += 1
        if gold_label == pred_label:
            self._correct_counts[gold_label] += 1
        self.keys.update((gold_label, pred_label))
        self._confusion[(gold_label, pred_label)] += 1

    def merge(self, other):
        for name in ("_gold_counts", "_pred_counts", "_correct_counts", "_confusion"):
            for k, v in getattr(other, name).items():
                getattr(self, name)[k] += v
        self.keys |= other.keys

    def get_score(self, label):
        return Score(predicted_count=self._pred_counts[label], gold_count=self._gold_counts[label],
                     correct_count=self._correct_counts[label])

    def get_gold_count(self, label=None):
        return sum(self._gold_counts.values()) if label is None else self._gold_counts[label]

    def get_pred_count(self, label=None):
        return sum(self._pred_counts.values()) if label is None else self._pred_counts[label]

    def get_gold_percent(self, label):
        return 100.0 * self.get_gold_count(label) / self.get_gold_count()

    def get_pred_percent(self, label):
        return 100.0 * self.get_pred_count(label) / self.get_pred_count()

    def get_accuracy(self):
        n = sum(self._gold_counts.values())
        return 0.0 if n == 0 or sum(self._correct_counts.values()) == 0 else 100.0 * sum(self._correct_counts.values()) / n

    def get_correct_count(self):
        return sum(self._correct_counts.values())

    def confusion_rows(self):
        """utils/ScoreDict.py:161-221 (print_confusion) as a list of text lines: one row per PREDICTED label, one column per
        GOLD label, every cell 'count (share of its gold column)'.  Label order = iteration order of the key set, as there."""
        keys = list(self.keys)
        width = max([len(str(int(v))) for v in self._confusion.values()] or [0]) + 9
        fmt = "%-" + str(width) + "s | " + ("%-" + str(width) + "s") * len(keys)
        col_tot = {g: sum(self._confusion[(g, p)] for p in keys) for g in keys}
        lines = [fmt % tuple([""] + [str(k) for k in keys])]
        for p in keys:
            cells = ["%d (%.1f%%)" % (int(self._confusion[(g, p)]), 100.0 * self._confusion[(g, p)] / col_tot[g] if col_tot[g] > 0 else 0.0)
                     for g in keys]
            lines.append(fmt % tuple([str(p)] + cells))
        return lines

    def print_confusion(self):
        for line in self.confusion_rows():
            print(line)


def rows_to_str(rows, has_headers=False, use_latex=False):
    """utils/string.py:33-128: left-aligned columns one wider than their longest cell; with headers the first column is followed
    by ' | ' and the first row by a dashed rule."""
    n_cols = max([len(r) for r in rows] or [0])
    rows = [list(r) + [""] * (n_cols - len(r)) for r in rows]
    if use_latex:
        body = ["\t" + " & ".join(r) + ("\\\\" if i < len(rows) - 1 else "") for i, r in enumerate(rows)]
        return "\n".join(["\\begin{tabular}{" + "l" * n_cols + "}"] + body + ["\\end{tabular}"])
    w = [max(len(r[c]) for r in rows) for c in range(n_cols)]
    fmt = ("%-" + str(w[0] + 1) + "s | " if has_headers else "") + "".join("%-" + str(w[c] + 1) + "s " for c in range(1 if has_headers else 0, n_cols))
    lines = [fmt % tuple(rows[0])]
    if has_headers:
        lines.append("".join("-" * (w[c] + 2) + ("|-" if c == 0 else "") for c in range(n_cols)))
    lines += [fmt % tuple(r) for r in rows[1:]]
    return "\n".join(lines)


def evaluate_relations(mention_pairs, pred_labels, gold_label_dict, log=None):
    """nn_utils/eval.py:10-93: every gold (ij, ji) pair whose two directed links were both predicted is scored once as
    null / coref / subset; two subset links that disagree with the gold direction count as '-reverse_sub-', any other
    inconsistent combination as 'invalid' (the reference's summary lines ask for the label '-invalid-', which is never
    incremented -- reproduced)."""
    sd = ScoreDict()
    pred = dict(zip(mention_pairs, pred_labels))
    for pair, gold in gold_label_dict.items():
        if pair[0] not in pred or pair[1] not in pred:
            continue
        ij, ji = pred[pair[0]], pred[pair[1]]
        p = "invalid"
        if ij == ji == 0:
            p = "null"
        elif ij == ji == 1:
            p = "coref"
        elif ij + ji == 5:
            if ij == 2:
                p = "subset_ij"
            elif ji == 2:
                p = "subset_ji"
        if gold.startswith("subset_") and p.startswith("subset_"):
            p = "subset" if gold == p else "-reverse_sub-"
            gold = "subset"
        if gold.startswith("subset_"):
            gold = "subset"
        if p.startswith("subset_"):
            p = "subset"
        sd.increment(gold, p)
    labels = ("-invalid-", "-reverse_sub-", "null", "coref", "subset")
    sd.summary_lines = ["%10s: %s" % (l, sd.get_score(l).to_string()) for l in labels] + \
                       ["%10s: pred: %d; gold: %d" % (l, sd.get_pred_count(l), sd.get_gold_count(l)) for l in labels]
    if log is not None:
        for line in sd.summary_lines:
            log.info(line)
    return sd


def evaluate_multiclass(gold_labels, pred_labels, class_names, log=None):
    """nn_utils/eval.py:95-184.  Returns the ScoreDict like the reference and logs the same tables: the gold / predicted counts,
    'Accuracy: <percent>', per-class P / R / F1 (what sklearn's precision_score / recall_score / f1_score with average=None give:
    classes are the sorted union of the labels that occur, so class l's row shows the l-th OCCURRING class's numbers, as there)
    and the confusion matrix over range(len(class_names))."""
    gold_labels = [int(x) for x in gold_labels]
    pred_labels = [int(x) for x in pred_labels]
    sd = ScoreDict(gold_labels, pred_labels)
    C = len(class_names)
    conf = np.zeros((C, C), np.int64)
    for g, p in zip(gold_labels, pred_labels):
        if g < C and p < C:
            conf[g, p] += 1
    gold_counts = np.bincount(gold_labels)
    pred_bins = np.bincount(pred_labels)
    pred_counts = np.pad(pred_bins, C - len(pred_bins), "constant")          # eval.py:113: pads BOTH ends, as written
    counts = [["gold", "pred"]] + [[str(gold_counts[l] if l < len(gold_counts) else 0), str(pred_counts[l])] for l in range(C)]
    present = sorted(set(gold_labels) | set(pred_labels))
    rows = [["", "P", "R", "F1"]]
    for l in range(C):
        if l < len(present):
            s_ = sd.get_score(present[l])
            rows.append([class_names[l], "%.2f%%" % (100.0 * s_.p), "%.2f%%" % (100.0 * s_.r), "%.2f%%" % (100.0 * s_.f1)])
        else:
            rows.append([class_names[l], "0.00%", "0.00%", "0.00%"])
    n_ok = sum(1 for g, p in zip(gold_labels, pred_labels) if g == p)
    acc = 100.0 * (float(n_ok) / len(gold_labels)) if gold_labels else float("nan")
    conf_tbl = [["g| p->"] + list(class_names)] + [[class_names[i]] + [str(int(conf[i][k])) for k in range(C)] for i in range(C)]
    sd.log_lines = ["Counts\n" + rows_to_str(counts), "\nAccuracy: " + str(acc), "\n" + rows_to_str(rows, True),
                    "\n" + rows_to_str(conf_tbl, True, False)]
    if log is not None:
        for line in sd.log_lines:
            log.info(line)
    sd.confusion_matrix, sd.accuracy = conf, acc
    return sd


def write_scores_file(path, pred_scores):
    """icl_core_lstm.py:240-252."""
    with open(path, "w") as f:
        for pid, scores in pred_scores.items():
            line = [pid]
            for s in scores:
                if s == 0:
                    s = np.nextafter(0, 1)
                line.append(str(np.log(s)))
            f.write(",".join(line) + "\n")
