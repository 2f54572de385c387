"""Host batch builder vs the loop oracle (bit-exact integers), enumeration helpers, predict padding."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from imagecaptionlearn_py_b200 import data as D
from imagecaptionlearn_py_b200 import synth
from oracle import load_batch_oracle as LO


@pytest.fixture(scope="module")
def corpus():
    return synth.make_corpus(6, seed=123, E=8, with_boxes=True, box_width=16, n_boxes=3)


@pytest.mark.parametrize("task", ["nonvis", "card", "rel_intra", "rel_cross", "affinity"])
def test_load_batch_bit_exact(corpus, task):
    dd = synth.make_data_dict(corpus, task, F=12)
    ids = synth.example_ids(dd, task)
    rng = np.random.default_rng(0)
    rng.shuffle(ids)
    ids = ids[:17]
    C = synth.N_CLASSES[task]
    got = D.load_batch(ids, dd, task, C)
    ref = LO.load_batch(ids, dd, task, C, box_lookup=lambda b: dd["box_table"][b] if task == "affinity" else None)
    assert set(ref) <= set(got)
    for k in ref:
        assert got[k].shape == ref[k].shape, k
        if k in LO.INDEX_NAMES or k in ("seq_lengths", "labels"):
            assert np.array_equal(got[k].astype(np.int64), ref[k].astype(np.int64)), k
        else:
            assert np.array_equal(got[k].astype(np.float32), ref[k].astype(np.float32)), k
    # packed form carries the same valid tokens
    pk = D.load_batch(ids, dd, task, C, packed=True)
    lens = got["seq_lengths"]
    rows = np.concatenate([got["sentences"][s, :lens[s]] for s in range(len(lens))])
    assert np.array_equal(pk["sentences_packed"], rows)


def test_index_rows_in_range(corpus):
    for task in ("nonvis", "rel_intra", "rel_cross"):
        dd = synth.make_data_dict(corpus, task, F=4)
        ids = synth.example_ids(dd, task)[:23]
        b = D.load_batch(ids, dd, task, synth.N_CLASSES[task])
        S = len(b["seq_lengths"])
        for n in D.INDEX_NAMES:
            m = b[n]
            used = ("_j_" not in n) or task.startswith("rel")
            if not used:
                continue
            assert m[:, 0].min() >= 0 and m[:, 0].max() <= 1
            assert m[:, 1].max() < S
            assert np.all(m[:, 2] < b["seq_lengths"][m[:, 1]]) and m[:, 2].min() >= 0
        if task == "rel_cross":
            assert np.array_equal(b["first_i_fw"][:, 1], 2 * np.arange(len(ids)))
            assert np.array_equal(b["first_j_fw"][:, 1], 2 * np.arange(len(ids)) + 1)


@settings(max_examples=50, deadline=None)
@given(n=st.integers(1, 70), B=st.integers(1, 16))
def test_predict_padding(n, B):
    ids = ["id%d" % i for i in range(n)]
    mat, pad = D.pad_ids_for_predict(ids, B)
    ref, rpad = LO.pad_ids_for_predict(ids, B)
    assert pad == rpad and 1 <= pad <= B and mat.shape == ref.shape
    assert list(mat.ravel()) == list(ref.ravel())
    flat = list(mat.ravel())
    assert flat[:n] == ids and all(x == ids[-1] for x in flat[n:])
    assert (mat.shape[0] - 1) * B + (B - pad) == n         # rows kept by core.py:673-677 == n


def test_ij_ji_induction_is_involution(corpus):
    dd = synth.make_data_dict(corpus, "rel_intra", F=4)
    ids = synth.example_ids(dd, "rel_intra")
    ij = D.get_ij_pairs(ids)
    assert ij == LO.get_ij_pairs(ids) and 0 < len(ij) < len(ids)
    rng = np.random.default_rng(0)
    scores = {p: rng.random(4) for p in ij}
    full = D.induce_ji_predictions(dict(scores))
    ref = LO.induce_ji(scores)
    assert len(full) == 2 * len(ij)
    for k, v in ref.items():
        assert np.array_equal(full[k], v)
    again = D.induce_ji_predictions({k: full[k] for k in ref})
    for p in ij:
        assert np.array_equal(again[p], scores[p])


def test_affinity_pairs_grouped_by_image(corpus):
    dd = synth.make_data_dict(corpus, "affinity", F=4)
    pairs = D.get_valid_mention_box_pairs(dd)
    assert len(pairs) == len(dd["mention_indices"]) * 3
    sh = D.shuffle_mention_box_pairs(pairs, np.random.default_rng(0))
    assert sorted(sh) == sorted(pairs)
    imgs = [p.split("#")[0] for p in sh]
    changes = sum(1 for a, b in zip(imgs, imgs[1:]) if a != b)
    assert changes == len(set(imgs)) - 1


def test_model_filename():
    a = dict(data_root="flickr30k_train", encoding_scheme="first_last_mention", activation="relu", epochs=100,
             learn_rate=0.001, batch_size=512, lstm_input_dropout=0.5, dropout=0.5, lstm_hidden_width=200,
             start_hidden_width=1024, hidden_depth=3, adam_epsilon=1e-08, clip_norm=5.0, data_norm=True,
             weighted_classes=False, early_stopping=True, rel_type="intra")
    assert D.build_model_filename(a, "rel_lstm") == (
        "flickr30k_train_rel_intra_lstm_flm_relu_epch100_lrn0.001_btch512_drp5050_lstm200_hdn1024-3_"
        "admEps1e-08_clip5.0_dataNorm_early.model")


def test_token_rows_mode_addresses_the_same_embedding_rows():
    """load_batch(packed="rows"): 'token_rows' index the concatenated caption matrices ('token_table') and reproduce the packed /
    padded sentence tensors exactly; every other entry of the dict is identical."""
    from imagecaptionlearn_py_b200 import data as nn_data
    from imagecaptionlearn_py_b200 import synth
    corpus = synth.make_corpus(6, seed=11)
    for task in ("card", "rel_cross"):
        dd = synth.make_data_dict(corpus, task, F=8)
        ids = synth.example_ids(dd, task)[:24]
        C = synth.N_CLASSES[task]
        a = nn_data.load_batch(ids, dd, task, C)
        b = nn_data.load_batch(ids, dd, task, C, packed=True)
        c = nn_data.load_batch(ids, dd, task, C, packed="rows")
        assert c["token_rows"].dtype == np.int32 and c["token_table"].dtype == np.float32
        assert np.array_equal(c["token_table"][c["token_rows"]], b["sentences_packed"])
        lens = a["seq_lengths"]
        rebuilt = np.zeros_like(a["sentences"])
        pos = 0
        for s, l in enumerate(lens):
            rebuilt[s, :l] = c["token_table"][c["token_rows"][pos:pos + l]]
            pos += l
        assert np.array_equal(rebuilt, a["sentences"])
        for k in a:
            if k != "sentences":
                assert np.array_equal(a[k], c[k]), k


def test_dedup_batches_address_the_same_tokens():
    """load_batch(dedup=True): distinct captions once, index matrices remapped -- every (sentence, word) an example addresses is
    the same token as without de-duplication, and everything that is not a sentence index is unchanged."""
    from imagecaptionlearn_py_b200 import data as nn_data
    from imagecaptionlearn_py_b200 import synth
    corpus = synth.make_corpus(4, seed=17, with_boxes=True, box_width=8)
    for task in ("nonvis", "rel_intra", "rel_cross", "affinity"):
        dd = synth.make_data_dict(corpus, task, F=8)
        ids = synth.example_ids(dd, task)[:64]
        C = synth.N_CLASSES[task]
        a = nn_data.load_batch(ids, dd, task, C)
        b = nn_data.load_batch(ids, dd, task, C, dedup=True)
        assert b["sentences"].shape[0] < a["sentences"].shape[0] and b["sentences"].shape[0] == len(b["seq_lengths"])
        for k in nn_data.INDEX_NAMES:
            ia, ib = a[k], b[k]
            assert np.array_equal(ia[:, [0, 2]], ib[:, [0, 2]]), k
            assert np.array_equal(a["seq_lengths"][ia[:, 1]], b["seq_lengths"][ib[:, 1]]), k
            assert np.array_equal(a["sentences"][ia[:, 1], ia[:, 2]], b["sentences"][ib[:, 1], ib[:, 2]]), k
            assert np.array_equal(a["sentences"][ia[:, 1]], b["sentences"][ib[:, 1]]), k         # the whole caption matches
        for k in a:
            if k not in nn_data.INDEX_NAMES and k not in ("sentences", "seq_lengths"):
                assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("task", ["nonvis", "card", "rel_intra", "rel_cross", "affinity"])
@pytest.mark.parametrize("packed", [False, "rows"])
def test_per_corpus_tables_give_what_per_batch_stacking_gives(monkeypatch, task, packed):
    """load_batch gathers a batch's labels / mention indices / mention features / caption offsets from per-corpus matrices built
    once per data_dict; without them (dicts too large for the tables: _TABLE_BYTES_MAX) it stacks the batch's rows out of the
    dicts as the first version did.  Same arrays bit for bit, same KeyError for an unknown id, and a data_dict whose dicts are
    replaced gets new tables."""
    corpus = synth.make_corpus(5, seed=3, E=8, with_boxes=(task == "affinity"), box_width=16)
    dd = synth.make_data_dict(corpus, task, F=6)
    ids = synth.example_ids(dd, task)[:37]
    C = {"nonvis": 2, "card": 12, "rel_intra": 4, "rel_cross": 4, "affinity": 2}[task]
    import copy
    dd2 = copy.deepcopy(dd)
    fast = D.load_batch(ids, dd, task, C, packed=packed)
    assert dd["_mention_table"][1] is not None and dd["_label_table"][1] is not None
    monkeypatch.setattr(D, "_TABLE_BYTES_MAX", 0)
    slow = D.load_batch(ids, dd2, task, C, packed=packed)
    assert dd2["_mention_table"][1] is None and dd2["_label_table"][1] is None
    assert set(fast) == set(slow)
    for k, v in slow.items():
        assert fast[k].dtype == v.dtype and np.array_equal(fast[k], v), k
    monkeypatch.undo()
    for d in (dd, dd2):
        with pytest.raises(KeyError):
            D.load_batch(ids[:3] + ["no such mention|no such box" if task == "affinity" else "no such id"], d, task, C, packed=packed)
    # replaced dicts -> rebuilt tables (same object, different contents would be a caller bug the reference has no answer to either)
    key = "ij_feats" if "rel" in task else "m_feats"
    dd["mention_features"] = {m: np.asarray(v) + 1.0 for m, v in dd["mention_features"].items()}
    again = D.load_batch(ids, dd, task, C, packed=packed)
    assert np.array_equal(again[key], fast[key] + 1.0)
