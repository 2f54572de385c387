"""Generates tests/golden/ref_batches.npz + ref_batches.json by running the REAL reference code from /root/reference:

    nn_utils/data.py   load_sentences :77-120, load_mentions :123-195, load_boxes :224-271, load_relation_labels :330-346,
                       load_batch :349-528, build_model_filename :274-327
    utils/data.py      load_sparse_feats :114-220 (called by the above)
    icl_affinity_lstm.py  get_valid_mention_box_pairs :59-74, shuffle_mention_box_pairs :24-56
    icl_relation_lstm.py  get_ij_pairs :213-222, induce_ji_predictions :225-245

on a small synthetic dataset that `imagecaptionlearn_py_b200.synth.write_dataset` emits in the reference's on-disk formats
(deterministic; the test regenerates the same files and checks their hash first).  The reference modules are imported
unmodified (tests/golden/ref_import.py explains the gensim stub and the module aliases); the two helper functions that cannot
run under Python 3 as written get a one-token in-memory patch, recorded in the fixture:
    shuffle_mention_box_pairs: `image_mention_box_pairs.keys()` -> `list(...)`   (np.random.shuffle of a dict view raises)
    induce_ji_predictions:     `pred_scores.keys()`             -> `list(...)`   (the dict grows while it is iterated)

Stored: every integer tensor of load_batch (index matrices, lengths, one-hot labels) verbatim, SHA-256 of the float tensors
(as float32 -- the reference hands TensorFlow float64, which the placeholders convert, nn_utils/core.py:288,348), the
parsers' dictionaries (ids in file order, mention indices, caption ids, max_seq_len, n_mention_feats), and the helpers'
outputs.  Run in the build container:   python tests/golden/make_ref_batches.py
"""
import hashlib
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

TASKS = ("nonvis", "card", "rel_intra", "rel_cross", "affinity")
N_CLASSES = {"nonvis": 2, "card": 12, "rel_intra": 4, "rel_cross": 4, "affinity": 2}
B = 24
DATA_ROOT = "synth_train"


def sha(a, dtype=np.float32):
    return hashlib.sha256(np.ascontiguousarray(np.asarray(a), dtype=dtype).tobytes()).hexdigest()


def write_synthetic(data_dir):
    """The dataset both sides read.  5 images, 4 boxes per image (4096-d, the reference's fixed box width)."""
    from imagecaptionlearn_py_b200 import synth
    corpus = synth.make_corpus(5, seed=20171201, with_boxes=True, n_boxes=4, vocab=300)
    for task in TASKS:
        synth.write_dataset(corpus, data_dir, DATA_ROOT, task, F={"rel_intra": 37, "rel_cross": 37}.get(task, 21))
    return corpus


def dataset_hash(data_dir):
    h = hashlib.sha256()
    for dp, dn, fn in sorted(os.walk(data_dir)):
        dn.sort()
        for f in sorted(fn):
            if f.endswith(".npz"):
                with np.load(os.path.join(dp, f)) as z:
                    for k in sorted(z.files):
                        h.update(k.encode()); h.update(np.ascontiguousarray(z[k]).tobytes())
            else:
                h.update(f.encode()); h.update(open(os.path.join(dp, f), "rb").read())
    return h.hexdigest()


def files_for(data_dir, task):
    raw, feats = os.path.join(data_dir, "raw"), os.path.join(data_dir, "feats")
    if task.startswith("rel"):
        tag = task.replace("rel_", "")
        return dict(sent=os.path.join(raw, DATA_ROOT + "_captions.txt"), ment=os.path.join(raw, DATA_ROOT + "_mentionPairs_%s.txt" % tag),
                    feats=os.path.join(feats, DATA_ROOT + "_relation_neural_%s.feats" % tag),
                    meta=os.path.join(feats, DATA_ROOT + "_relation_neural_%s_meta.json" % tag),
                    gold=os.path.join(raw, DATA_ROOT + "_mentionPair_labels.txt"))
    d = dict(sent=os.path.join(raw, DATA_ROOT + "_captions.txt"), ment=os.path.join(raw, DATA_ROOT + "_mentions_%s.txt" % task),
             feats=os.path.join(feats, DATA_ROOT + "_%s_neural.feats" % task), meta=os.path.join(feats, DATA_ROOT + "_%s_neural_meta.json" % task))
    if task == "affinity":
        data, split = DATA_ROOT.rsplit("_", 1)
        d["labels"] = os.path.join(raw, DATA_ROOT + "_affinity_labels.txt")
        d["bdir"] = os.path.join(feats, data + "_boxes", split)
    return d


def pick_ids(task, all_ids):
    """B ids: file order for the first half, a fixed stride for the rest (affinity keeps image grouping like the reference)."""
    if task == "affinity":
        return list(all_ids[:B])
    idx = list(range(B // 2)) + [(7 * i + 3) % len(all_ids) for i in range(B - B // 2)]
    return [all_ids[i] for i in idx]


def main():
    import ref_import
    mods = ref_import.install(tensorflow=types.ModuleType("tensorflow"))
    ref = mods["nn_utils.data"]
    tmp = tempfile.mkdtemp(prefix="iclref_")
    assert "box" not in tmp          # utils/data.py:208 shifts indices when the PATH contains "box"
    write_synthetic(tmp)
    meta = dict(source="/root/reference nn_utils/data.py, utils/data.py, icl_affinity_lstm.py, icl_relation_lstm.py (imported; see make_ref_batches.py)",
                dataset_sha256=dataset_hash(tmp), B=B, tasks={}, helpers={})
    arrays = {}
    ref.__dict__["__WORD_2_VEC_PATH"] = os.path.join(tmp, "raw", DATA_ROOT + "_embeddings.npz")
    ref.init_w2v()
    for task in TASKS:
        f = files_for(tmp, task)
        C = N_CLASSES[task]
        dd = ref.load_sentences(f["sent"], "w2v")
        dd.update(ref.load_mentions(f["ment"], task, f["feats"], f["meta"], C))
        if task == "affinity":
            dd.update(ref.load_boxes(f["labels"], f["bdir"]))
        t = dict(n_classes=C, max_seq_len=int(dd["max_seq_len"]), n_mention_feats=int(dd["n_mention_feats"]),
                 word_embedding_width=int(dd["word_embedding_width"]), sentence_ids=list(dd["sentences"].keys()),
                 sentences_sha256=sha(np.concatenate([dd["sentences"][k] for k in dd["sentences"]], 0)),
                 sentence_lens=[int(len(dd["sentences"][k])) for k in dd["sentences"]],
                 mention_ids=list(dd["mention_indices"].keys()),
                 mention_indices=[list(map(int, dd["mention_indices"][k])) for k in dd["mention_indices"]],
                 caption_ids=[dd["caption_ids"][k] for k in dd["mention_indices"]],
                 mention_features_sha256=sha(np.stack([dd["mention_features"][k] for k in dd["mention_indices"]])),
                 label_ids_head=list(dd["labels"].keys())[:50],
                 labels_argmax_head=[int(np.argmax(dd["labels"][k])) for k in list(dd["labels"].keys())[:50]], n_labels=len(dd["labels"]))
        if task == "affinity":
            aff = ref_import.load_script("icl_affinity_lstm", patch=lambda s: s.replace(
                "img_ids = image_mention_box_pairs.keys()", "img_ids = list(image_mention_box_pairs.keys())"))
            all_ids = aff.get_valid_mention_box_pairs(dd)
            np.random.seed(20171201)
            shuffled = aff.shuffle_mention_box_pairs(list(all_ids))
            meta["helpers"]["valid_mention_box_pairs_sha256"] = hashlib.sha256("\n".join(all_ids).encode()).hexdigest()
            meta["helpers"]["valid_mention_box_pairs_n"] = len(all_ids)
            meta["helpers"]["shuffled_head"] = shuffled[:40]
            meta["helpers"]["shuffled_sha256"] = hashlib.sha256("\n".join(shuffled).encode()).hexdigest()
            t["box_embedding_width"] = int(dd["box_embedding_width"])
            all_ids = shuffled
        else:
            all_ids = list(dd["mention_indices"].keys())
        if task == "rel_intra":
            rel = ref_import.load_script("icl_relation_lstm", patch=lambda s: s.replace(
                "for ij_pair in pred_scores.keys():", "for ij_pair in list(pred_scores.keys()):"))
            ij = rel.get_ij_pairs(all_ids)
            rng = np.random.RandomState(5)
            scores = {k: rng.rand(4) for k in ij[:20]}
            ind = rel.induce_ji_predictions(dict(scores))
            meta["helpers"]["ij_pairs_sha256"] = hashlib.sha256("\n".join(ij).encode()).hexdigest()
            meta["helpers"]["ij_pairs_n"] = len(ij)
            meta["helpers"]["induce_keys"] = list(ind.keys())
            arrays["induce_in"] = np.stack([scores[k] for k in ij[:20]])
            arrays["induce_out"] = np.stack([ind[k] for k in ind])
            gold = ref.load_relation_labels(f["gold"])
            meta["helpers"]["relation_gold_n"] = len(gold)
            meta["helpers"]["relation_gold_head"] = [[k[0], k[1], v] for k, v in list(gold.items())[:30]]
        ids = pick_ids(task, all_ids)
        bt = ref.load_batch(ids, dd, task, C)
        t["ids"] = ids
        t["batch_keys"] = sorted(bt.keys())
        t["batch_dtypes"] = {k: str(bt[k].dtype) for k in bt}
        t["batch_shapes"] = {k: list(bt[k].shape) for k in bt}
        t["float_sha256"] = {}
        for k, v in bt.items():
            if k in ("sentences", "m_feats", "ij_feats", "box_embeddings", "b_feats"):
                t["float_sha256"][k] = sha(v)
            else:                          # index matrices, seq_lengths, labels: integers stored in float64 arrays
                assert np.all(v == np.round(v))
                arrays["%s/%s" % (task, k)] = v.astype(np.int32)
        meta["tasks"][task] = t
    # build_model_filename (nn_utils/data.py:274-327)
    names = []
    for rel_type, tk in ((None, "nonvis_lstm"), ("intra", "rel_lstm"), (None, "affinity_lstm")):
        for enc in ("first_last_sentence", "first_last_mention"):
            for clip, norm, weighted, early in ((5.0, False, False, False), (None, True, True, True)):
                ad = dict(data_root="flickr30k_train", activation="relu", epochs=100, learn_rate=0.001, batch_size=512,
                          lstm_input_dropout=0.5, dropout=0.5, lstm_hidden_width=200, start_hidden_width=1024, hidden_depth=3,
                          adam_epsilon=1e-08, clip_norm=clip, data_norm=norm, weighted_classes=weighted, early_stopping=early,
                          encoding_scheme=enc, rel_type=rel_type)
                names.append(dict(args=ad, task=tk, name=ref.build_model_filename(ad, tk)))
    meta["model_filenames"] = names
    np.savez_compressed(os.path.join(HERE, "ref_batches.npz"), **arrays)
    json.dump(meta, open(os.path.join(HERE, "ref_batches.json"), "w"), indent=1)
    ref_import.uninstall()
    print("wrote ref_batches.npz (%d arrays) / ref_batches.json; dataset %s" % (len(arrays), meta["dataset_sha256"][:16]))


if __name__ == "__main__":
    main()
