"""Drop-in proof: the reference's OWN script body -- `train()` and `predict()` of icl_core_lstm.py (:21-199, :202-252), compiled
from the unmodified source file -- runs against the shim after exactly the import swap of INTEGRATION.md:

    import tensorflow as tf            ->  imagecaptionlearn_py_b200.core          (variable_scope / get_collection / Session / train.Saver /
    from nn_utils import core          ->  imagecaptionlearn_py_b200.core           global_variables_initializer / run_op / get_pred_scores_mcc ...)

Everything else is the reference's code: its parsers and `load_batch` (nn_utils/data.py, float64 host tensors), its evaluation
(nn_utils/eval.py, utils/ScoreDict.py), its Logger, its epoch / save / evaluate loop.  The data is a synthetic corpus in the
reference's on-disk formats (synth.write_dataset).  Python-2-isms of the script that Python 3 cannot execute are bridged WITHOUT
editing it: `n_iter = n_pairs / batch_size` (:119) gets an int subclass whose reflected true division is floor division; the
print statements of nn_utils/eval.py / utils/ScoreDict.py are parenthesised in memory (tests/golden/ref_import.py).

The reference sources are read from /root/reference (build container) or from the unmodified copy tools/stage_reference.sh
stages under baseline/_ref/ (git-ignored; travels to the GPU box).  Skipped when neither exists."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


class Py2Int(int):
    """`n / Py2Int(b)` is Python 2's integer division (a subclass's reflected operator is tried first)."""
    def __rtruediv__(self, other):
        return int(other) // int(self)


class Log(object):
    """utils/Logger.py's interface as the scripts use it (info / log_status / debug / warning); keeps the formatted lines."""
    def __init__(self):
        self.lines = []

    def info(self, *a):
        a = [x for x in a if x is not None]
        self.lines.append(a[0] % tuple(a[1:]) if len(a) > 1 else str(a[0]))

    debug = warning = error = info

    def log_status(self, *a):
        pass


def test_reference_train_and_predict_bodies_run_on_the_shim(tmp_path):
    import ref_import
    if not ref_import.available():
        pytest.skip("reference sources not present (%s): run tools/stage_reference.sh" % ref_import.REF)
    from imagecaptionlearn_py_b200 import core, synth
    assert "box" not in str(tmp_path)
    corpus = synth.make_corpus(70, seed=11, with_boxes=False, vocab=400)
    ev_corpus = synth.make_corpus(12, seed=12, with_boxes=False, vocab=400)
    synth.write_dataset(corpus, str(tmp_path), "synth_train", "nonvis", F=21)
    synth.write_dataset(ev_corpus, str(tmp_path), "synth_dev", "nonvis", F=21)

    def files(root):
        raw, feats = os.path.join(str(tmp_path), "raw"), os.path.join(str(tmp_path), "feats")
        return dict(sent=os.path.join(raw, root + "_captions.txt"), ment=os.path.join(raw, root + "_mentions_nonvis.txt"),
                    feats=os.path.join(feats, root + "_nonvis_neural.feats"), meta=os.path.join(feats, root + "_nonvis_neural_meta.json"))
    tr, ev = files("synth_train"), files("synth_dev")

    core.reset_default_graph()
    core.set_random_seeds()
    mods = ref_import.install(tensorflow=core, nn_core=core)          # the import swap
    try:
        ref_data = mods["nn_utils.data"]
        ref_data.__dict__["__WORD_2_VEC_PATH"] = os.path.join(str(tmp_path), "raw", "synth_train_embeddings.npz")
        ref_data.init_w2v()
        script = ref_import.load_script("icl_core_lstm")
        log = Log()
        B = 8
        model_file = os.path.join(str(tmp_path), "model_nonvis")
        np.random.seed(3)
        # 10 epochs: the reference evaluates (get_pred_scores_mcc + nn_eval.evaluate_multiclass) on every 10th epoch (:157)
        script.train("nonvis", "first_last_mention", "w2v", tr["sent"], tr["ment"], tr["feats"], tr["meta"], 10, Py2Int(B), 32, 64, 1, False,
                     0.9, 0.9, 0.002, 1e-8, 5.0, False, "relu", model_file=model_file, eval_sentence_file=ev["sent"],
                     eval_mention_idx_file=ev["ment"], eval_feature_file=ev["feats"], eval_feature_meta_file=ev["meta"],
                     early_stopping=True, log=log)
        assert os.path.exists(model_file + ".npz")
        saved = [l for l in log.lines if l.startswith("Saving model; Average Loss")]
        assert len(saved) == 10, log.lines[-5:]
        first, last = float(saved[0].split("Loss:")[1].split(";")[0]), float(saved[-1].split("Loss:")[1].split(";")[0])
        assert np.isfinite(first) and np.isfinite(last) and last < first, (first, last)      # it trains
        assert any("New best at current epoch" in l for l in log.lines)                     # the evaluation branch ran

        # predict(): the reference restores through tf.train.import_meta_graph; the graph is already built here, so restore + call
        scores_file = os.path.join(str(tmp_path), "scores_nonvis.csv")
        with core.Session() as sess:
            core.train.Saver().restore(sess, model_file)
            script.predict("nonvis", "first_last_mention", "w2v", sess, B, ev["sent"], ev["ment"], ev["feats"], ev["meta"],
                           scores_file=scores_file, log=log)
            # the same predictions through the repo's own loaders + predict path on the same checkpoint: identical scores
            from imagecaptionlearn_py_b200 import loaders
            dd = loaders.load_sentences(ev["sent"], loaders.Embeddings.from_npz(os.path.join(str(tmp_path), "raw", "synth_train_embeddings.npz")))
            dd.update(loaders.load_mentions(ev["ment"], "nonvis", ev["feats"], ev["meta"], 2))
            ours, _ = core.get_pred_scores_mcc("nonvis", "first_last_mention", sess, B, list(dd["mention_indices"].keys()), dd, 2)
        lines = open(scores_file).read().strip().split("\n")
        assert len(lines) == len(ours)
        for ln in lines:
            f = ln.split(",")
            p = np.exp(np.array(f[1:], np.float64))
            assert abs(p.sum() - 1.0) < 1e-5
            # the script writes str(np.log(score)) of the float32 score (icl_core_lstm.py:246-249): same probabilities -> same text
            assert [str(np.log(x)) for x in ours[f[0]]] == f[1:], f[0]
    finally:
        ref_import.uninstall()
        sys.modules.pop("tensorflow", None)
        core.reset_default_graph()
