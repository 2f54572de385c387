"""Pins the host side of the path to outputs of the REFERENCE ITSELF.

tests/golden/ref_batches.{npz,json} were produced by importing /root/reference/nn_utils/data.py, utils/data.py,
icl_affinity_lstm.py and icl_relation_lstm.py and running them on a small synthetic dataset (tests/golden/make_ref_batches.py).
Here the same dataset is regenerated (its hash is checked first), parsed with `imagecaptionlearn_py_b200.loaders` and
batched with `imagecaptionlearn_py_b200.data.load_batch`; every integer tensor must be identical and every float tensor
bit-identical after the float64 -> float32 feed conversion.  Nothing here reads /root/reference."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_ref_batches as G  # noqa: E402

from imagecaptionlearn_py_b200 import data as D  # noqa: E402
from imagecaptionlearn_py_b200 import loaders as L  # noqa: E402

META = json.load(open(os.path.join(HERE, "golden", "ref_batches.json")))
ARR = np.load(os.path.join(HERE, "golden", "ref_batches.npz"))


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("iclds"))
    assert "box" not in d
    G.write_synthetic(d)
    if G.dataset_hash(d) != META["dataset_sha256"]:
        pytest.fail("the synthetic dataset generator drifted from the one the reference fixtures were made with; "
                    "re-run tests/golden/make_ref_batches.py in the build container")
    emb = L.Embeddings.from_npz(os.path.join(d, "raw", G.DATA_ROOT + "_embeddings.npz"))
    return d, emb


def load_ours(d, emb, task):
    f = G.files_for(d, task)
    C = G.N_CLASSES[task]
    dd = L.load_sentences(f["sent"], emb)
    dd.update(L.load_mentions(f["ment"], task, f["feats"], f["meta"], C))
    if task == "affinity":
        dd.update(L.load_boxes(f["labels"], f["bdir"]))
    return dd, f, C


@pytest.mark.parametrize("task", G.TASKS)
def test_parsers_match_the_reference(dataset, task):
    d, emb = dataset
    dd, f, C = load_ours(d, emb, task)
    t = META["tasks"][task]
    assert dd["max_seq_len"] == t["max_seq_len"]
    assert dd["n_mention_feats"] == t["n_mention_feats"]
    assert dd["word_embedding_width"] == t["word_embedding_width"]
    assert list(dd["sentences"].keys()) == t["sentence_ids"]
    assert [len(dd["sentences"][k]) for k in dd["sentences"]] == t["sentence_lens"]
    assert G.sha(np.concatenate([dd["sentences"][k] for k in dd["sentences"]], 0)) == t["sentences_sha256"]
    assert list(dd["mention_indices"].keys()) == t["mention_ids"]
    assert [list(map(int, dd["mention_indices"][k])) for k in dd["mention_indices"]] == t["mention_indices"]
    caps = [dd["caption_ids"][k] for k in dd["mention_indices"]]
    assert [list(c) if isinstance(c, tuple) else c for c in caps] == t["caption_ids"]
    assert G.sha(np.stack([dd["mention_features"][k] for k in dd["mention_indices"]])) == t["mention_features_sha256"]
    assert len(dd["labels"]) == t["n_labels"]
    assert list(dd["labels"].keys())[:50] == t["label_ids_head"]
    assert [int(np.argmax(dd["labels"][k])) for k in list(dd["labels"].keys())[:50]] == t["labels_argmax_head"]
    if task == "affinity":
        assert dd["box_embedding_width"] == t["box_embedding_width"]


@pytest.mark.parametrize("task", G.TASKS)
@pytest.mark.parametrize("packed", [False, True, "rows"])
def test_load_batch_matches_the_reference(dataset, task, packed):
    d, emb = dataset
    dd, f, C = load_ours(d, emb, task)
    t = META["tasks"][task]
    bt = D.load_batch(t["ids"], dd, task, C, packed=packed)
    for name in D.INDEX_NAMES + ("seq_lengths", "labels"):
        ref = ARR["%s/%s" % (task, name)]
        got = np.asarray(bt[name])
        assert got.shape == tuple(t["batch_shapes"][name]), name
        assert np.array_equal(got.astype(np.int64), ref.astype(np.int64)), name          # bit-exact integers
    # float tensors: identical after the float64 -> float32 feed conversion
    lens = ARR["%s/seq_lengths" % task]
    if packed is False:
        assert bt["sentences"].shape == tuple(t["batch_shapes"]["sentences"])
        sent = bt["sentences"]
    else:
        rows = bt["token_table"][bt["token_rows"]] if packed == "rows" else bt["sentences_packed"]
        sent = np.zeros(t["batch_shapes"]["sentences"], np.float32)
        pos = 0
        for s, n in enumerate(lens):
            sent[s, :n] = rows[pos:pos + n]
            pos += n
    assert G.sha(sent) == t["float_sha256"]["sentences"]
    fk = "ij_feats" if task.startswith("rel") else "m_feats"
    assert G.sha(bt[fk]) == t["float_sha256"][fk]
    if task == "affinity":
        assert G.sha(bt["box_embeddings"]) == t["float_sha256"]["box_embeddings"]
    # the reference's dict has exactly these keys; ours may add wire-format keys but must not miss any (packed=False: none missing)
    if packed is False:
        assert sorted(k for k in bt.keys()) == t["batch_keys"]


def test_affinity_box_table_path_gives_the_same_rows(dataset):
    """The device-resident box table (loaders.load_all_boxes + load_batch(packed="rows")) must address the very rows the
    reference parses out of the per-image files for every batch."""
    d, emb = dataset
    dd, f, C = load_ours(d, emb, "affinity")
    L.load_all_boxes(dd)
    t = META["tasks"]["affinity"]
    bt = D.load_batch(t["ids"], dd, "affinity", C, packed="rows")
    assert G.sha(bt["box_table"][bt["box_rows"]]) == t["float_sha256"]["box_embeddings"]


def test_enumeration_helpers_match_the_reference(dataset):
    d, emb = dataset
    h = META["helpers"]
    dd, f, C = load_ours(d, emb, "affinity")
    valid = D.get_valid_mention_box_pairs(dd)
    assert len(valid) == h["valid_mention_box_pairs_n"]
    assert hashlib.sha256("\n".join(valid).encode()).hexdigest() == h["valid_mention_box_pairs_sha256"]
    np.random.seed(20171201)
    sh = D.shuffle_mention_box_pairs(list(valid))
    assert sh[:40] == h["shuffled_head"]
    assert hashlib.sha256("\n".join(sh).encode()).hexdigest() == h["shuffled_sha256"]
    dd, f, C = load_ours(d, emb, "rel_intra")
    ij = D.get_ij_pairs(list(dd["mention_indices"].keys()))
    assert len(ij) == h["ij_pairs_n"]
    assert hashlib.sha256("\n".join(ij).encode()).hexdigest() == h["ij_pairs_sha256"]
    scores = {k: ARR["induce_in"][i] for i, k in enumerate(ij[:20])}
    ind = D.induce_ji_predictions(dict(scores))
    assert list(ind.keys()) == h["induce_keys"]
    assert np.array_equal(np.stack([ind[k] for k in ind]), ARR["induce_out"])
    gold = L.load_relation_labels(f["gold"])
    assert len(gold) == h["relation_gold_n"]
    assert [[k[0], k[1], v] for k, v in list(gold.items())[:30]] == h["relation_gold_head"]


def test_model_file_names_match_the_reference():
    for case in META["model_filenames"]:
        assert D.build_model_filename(dict(case["args"]), case["task"]) == case["name"]
