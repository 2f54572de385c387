"""File formats + drop-in command lines.  CPU: the synthetic corpus written in the reference's formats and read back
through the loaders gives the same batch_tensors as the in-memory path.  GPU: the icl_core_lstm.py / icl_relation_lstm.py /
icl_affinity_lstm.py command lines train, save, restore, predict and write the reference's scores files."""
import os

import numpy as np
import pytest

from imagecaptionlearn_py_b200 import data as nn_data
from imagecaptionlearn_py_b200 import loaders, synth


def _write(tmp, task, root, n_img=6, F=16):
    corpus = synth.make_corpus(n_img, seed=3, with_boxes=(task == "affinity"), box_width=4096 if task == "affinity" else 64)
    dd = synth.write_dataset(corpus, str(tmp), root, task, F=F)
    return corpus, dd


@pytest.mark.parametrize("task", ["card", "rel_intra", "affinity"])
def test_written_files_load_back_to_the_same_batches(tmp_path, task):
    root = "flickr30k_train"
    corpus, dd = _write(tmp_path, task, root)
    d = str(tmp_path) + "/"
    emb = loaders.Embeddings.from_npz(d + "raw/" + root + "_embeddings.npz")
    ld = loaders.load_sentences(d + "raw/" + root + "_captions.txt", emb)
    C = synth.N_CLASSES[task]
    if task == "card":
        ld.update(loaders.load_mentions(d + "raw/" + root + "_mentions_card.txt", task, d + "feats/" + root + "_card_neural.feats",
                                        d + "feats/" + root + "_card_neural_meta.json", C))
    elif task == "rel_intra":
        ld.update(loaders.load_mentions(d + "raw/" + root + "_mentionPairs_intra.txt", task,
                                        d + "feats/" + root + "_relation_neural_intra.feats",
                                        d + "feats/" + root + "_relation_neural_intra_meta.json", C))
        gold = loaders.load_relation_labels(d + "raw/" + root + "_mentionPair_labels.txt")
        assert gold and all(v in ("null", "coref", "subset_ij", "subset_ji") for v in gold.values())
    else:
        ld.update(loaders.load_mentions(d + "raw/" + root + "_mentions_affinity.txt", task, d + "feats/" + root + "_affinity_neural.feats",
                                        d + "feats/" + root + "_affinity_neural_meta.json", C))
        ld.update(loaders.load_boxes(d + "raw/" + root + "_affinity_labels.txt", d + "feats/flickr30k_boxes/train/"))
    ids = (nn_data.get_valid_mention_box_pairs(ld) if task == "affinity" else list(ld["mention_indices"].keys()))[:40]
    a = nn_data.load_batch(ids, ld, task, C)
    dd["max_seq_len"] = ld["max_seq_len"]
    # the reference's load_sentences leaves the "\n" on the last token of every line, so that token is looked up as UNK
    # (nn_utils/data.py:92-93; pinned in tests/test_ref_fixtures.py): give the in-memory corpus the same last rows
    dd["sentences"] = {k: np.concatenate([m[:-1], corpus["table"][-1:]], 0) for k, m in dd["sentences"].items()}
    dd.pop("_flat", None)
    b = nn_data.load_batch(ids, dd, task, C)
    assert set(a) == set(b)
    for k in a:
        if a[k].dtype.kind in "iu":
            assert np.array_equal(a[k], b[k]), k                        # integers bit-exact
        else:
            np.testing.assert_allclose(a[k], b[k], rtol=1e-5, atol=1e-6, err_msg=k)   # "%g" text round trip of the features


def test_legacy_flag_aliases_parse():
    import argparse
    from imagecaptionlearn_py_b200 import drivers
    p = argparse.ArgumentParser()
    drivers.common_args(p)
    # config/lstm_intra_params.config:96: the sweep lines still use the old flag names
    a = p.parse_args("--data_dir /x --epochs=100 --batch_size=512 --lstm_hidden_width=200 --start_hidden_width=1024 "
                     "--hidden_depth=3 --input_keep_prob=0.5 --other_keep_prob=0.5 --data_norm "
                     "--pair_enc_scheme=first_last_mention".split())
    assert a.lstm_input_dropout == 0.5 and a.dropout == 0.5 and a.encoding_scheme == "first_last_mention"
    assert nn_data.build_model_filename(dict(vars(a), data_root="flickr30k_train"), "nonvis_lstm").endswith(".model")


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["core", "relation", "affinity"])
def test_cli_train_then_predict_writes_scores(tmp_path, which):
    from imagecaptionlearn_py_b200 import drivers
    task = {"core": "card", "relation": "rel_intra", "affinity": "affinity"}[which]
    for root in ("flickr30k_train", "flickr30k_dev"):
        _write(tmp_path, task, root, n_img=5 if which != "relation" else 4)
    common = ["--data_dir", str(tmp_path), "--batch_size", "32", "--lstm_hidden_width", "20", "--start_hidden_width", "32",
              "--hidden_depth", "1", "--epochs", "2", "--eval_every", "1", "--model_file", str(tmp_path / "m.model")]
    if which == "core":
        extra = ["--task", "card", "--data_root", "flickr30k_train", "--eval_data_root", "flickr30k_dev"]
        main = drivers.main_core
    elif which == "relation":
        extra = ["--rel_type", "intra", "--data_root", "flickr30k_train", "--eval_data_root", "flickr30k_dev"]
        main = drivers.main_relation
    else:
        extra = ["--data", "flickr30k", "--split", "train", "--eval_data", "flickr30k", "--eval_split", "dev"]
        main = drivers.main_affinity
    best = main(common + extra + ["--train"])
    assert best is not None and os.path.exists(str(tmp_path / "m.model") + ".npz")
    out = main(common + extra + ["--predict"])
    lines = open(out).read().strip().split("\n")
    assert len(lines) > 30
    C = synth.N_CLASSES[task]
    for ln in lines[:20]:
        f = ln.split(",")
        assert len(f) == C + 1 and abs(sum(np.exp(float(x)) for x in f[1:]) - 1.0) < 1e-4     # "<id>,<ln p_0>,..."


def _write_multitask(tmp, data, split, n_img=4, F=16):
    from imagecaptionlearn_py_b200 import drivers
    corpus = synth.make_corpus(n_img, seed=5, with_boxes=True, box_width=64)
    for task in drivers.MT_TASKS:
        synth.write_dataset(corpus, str(tmp), data + "_" + split, task, F=F, naming="multitask")


def test_multitask_files_load_for_every_task(tmp_path):
    """icl_multitask_lstm.py:153-209: the joint data set's file naming (shared `_relation.feats`, `_mention_box_labels.txt`)."""
    import argparse
    from imagecaptionlearn_py_b200 import drivers
    _write_multitask(tmp_path, "flickr30k", "train")
    p = argparse.ArgumentParser()
    drivers.common_args(p)
    args = p.parse_args(["--data_dir", str(tmp_path)])
    dicts = drivers.mt_load_data(args, "flickr30k", "train")
    ids = drivers.mt_task_ids(dicts)
    assert set(dicts) == set(drivers.MT_TASKS) and all(len(ids[t]) > 0 for t in drivers.MT_TASKS)
    for t in drivers.MT_TASKS:
        bt = nn_data.load_batch(ids[t][:8], dicts[t], t, len(drivers.MT_CLASSES[t]))
        assert bt["labels"].shape == (8, len(drivers.MT_CLASSES[t]))
        assert bt["sentences"].shape[0] == (16 if t == "rel_cross" else 8)


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["simple_joint", "weighted_joint", "alternate"])
def test_multitask_cli_train_then_predict(tmp_path, scheme):
    from imagecaptionlearn_py_b200 import drivers
    for split in ("train", "dev"):
        _write_multitask(tmp_path, "flickr30k", split)
    common = ["--data_dir", str(tmp_path), "--batch_size", "16", "--lstm_hidden_width", "20", "--start_hidden_width", "32",
              "--hidden_depth", "1", "--epochs", "1", "--model_file", str(tmp_path / "mt.model"), "--data", "flickr30k",
              "--split", "train", "--eval_data", "flickr30k", "--eval_split", "dev", "--multitask_scheme", scheme]
    f1 = drivers.main_multitask(common + ["--train"])
    assert set(f1) == set(drivers.MT_TASKS)
    outs = drivers.main_multitask(common + ["--predict"])
    assert len(outs) == 5
    for out, task in zip(outs, drivers.MT_TASKS):
        assert "_mulit_" + scheme + "_lstm.scores" in out
        lines = open(out).read().strip().split("\n")
        C = len(drivers.MT_CLASSES[task])
        f = lines[0].split(",")
        assert len(f) == C + 1 and abs(sum(np.exp(float(x)) for x in f[1:]) - 1.0) < 1e-4


@pytest.mark.gpu
def test_saver_restores_a_tf_saver_v2_bundle(tmp_path, monkeypatch):
    """Saver.save with ICL_SAVE_TF_BUNDLE=1 also writes `<path>.index` / `.data-00000-of-00001` under the TF variable names (Adam
    slots as `<var>/Adam`, `<var>/Adam_1`); Saver.restore falls back to such a bundle when there is no .npz -- the route by which a
    checkpoint of the reference (tf.train.Saver, icl_core_lstm.py:107,155) is loaded."""
    import os
    from imagecaptionlearn_py_b200 import core, tf_checkpoint
    from tests.helpers import tiny_problem
    from tests.test_gpu_parity import make_session
    from imagecaptionlearn_py_b200 import _cabi
    p = tiny_problem(seed=3, task="nonvis", enc="first_last_mention", act="relu", S=16, T=7, E=8, H=4, F=4, widths=(8, 4))
    core_, sess = make_session(p, "simt")
    sess.run(_cabi.OP_TRAIN, [dict(p["batch"])], 1.0, 1.0, True)
    monkeypatch.setenv("ICL_SAVE_TF_BUNDLE", "1")
    path = str(tmp_path / "m.model")
    core.Saver().save(sess, path)
    want = sess.state_dict()
    proba = sess.run(_cabi.OP_PREDICT, [dict(p["batch"])], 1.0, 1.0, True)[0]["proba"].copy()
    sess.close()
    names = tf_checkpoint.read_bundle(path)
    assert "hdn_1/Variable/Adam" in names and "bidirectional_lstm/bidirectional_rnn/fw/basic_lstm_cell/kernel" in names
    os.remove(path + ".npz")
    core_, sess2 = make_session(p, "simt")
    for k in p["params"]:
        sess2.set_tensor(k, np.zeros_like(p["params"][k]).reshape(1, -1) if p["params"][k].ndim == 1 else np.zeros_like(p["params"][k]))
    core.Saver().restore(sess2, path)
    got = sess2.state_dict()
    for k, v in want.items():
        assert np.array_equal(np.asarray(got[k]), np.asarray(v)), k
    assert np.array_equal(sess2.run(_cabi.OP_PREDICT, [dict(p["batch"])], 1.0, 1.0, True)[0]["proba"], proba)
    sess2.close()
