"""Evaluation helpers of the hot path's callers, restated for Python 3 (host-side; no kernels).

    Score, ScoreDict            utils/Score.py:4-48, utils/ScoreDict.py:13-154
    evaluate_multiclass         nn_utils/eval.py:95-184  (per-class P/R/F1, accuracy, confusion matrix)
    write_scores_file           icl_core_lstm.py:240-252 ("<id>,<ln p_0>,...", zero -> nextafter(0,1) before the log)

`utils/Score.py` of the reference imports and runs unchanged under Python 3; tests/golden/score_golden.json was generated
from it (tests/golden/make_golden.py) and pins `Score` here bit-for-bit.
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

    def get_accuracy(self):
        n = sum(self._gold_counts.values())
        return 0.0 if n == 0 else 100.0 * sum(self._correct_counts.values()) / n

    def get_correct_count(self):
        return sum(self._correct_counts.values())


def evaluate_multiclass(gold_labels, pred_labels, class_names, log=None):
    """nn_utils/eval.py:95-184.  Returns the ScoreDict like the reference and logs the same three tables."""
    gold_labels = [int(x) for x in gold_labels]
    pred_labels = [int(x) for x in pred_labels]
    sd = ScoreDict(gold_labels, pred_labels)
    C = len(class_names)
    conf = np.zeros((C, C), np.int64)
    for g, p in zip(gold_labels, pred_labels):
        conf[g, p] += 1
    rows = [["", "P", "R", "F1"]]
    for l in range(C):
        s = sd.get_score(l)
        rows.append([class_names[l], "%.2f%%" % (100.0 * s.p), "%.2f%%" % (100.0 * s.r), "%.2f%%" % (100.0 * s.f1)])
    acc = 100.0 * float(np.trace(conf)) / max(1, len(gold_labels))
    text = "Accuracy: %s\n%s\n%s" % (acc, "\n".join("\t".join(r) for r in rows),
                                     "\n".join("\t".join(str(v) for v in row) for row in conf))
    if log is not None:
        log.info("\n" + text)
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
