"""Readers for the reference's on-disk formats (host text parsing), Python 3.

    load_sentences        nn_utils/data.py:77-120   "<img>#<cap>\\t<tokens>"           -> data_dict['sentences'][cap_id] = [L, 300]
    load_mentions         nn_utils/data.py:123-195  "<id>\\t<first>,<last>[,...]\\t<label>" + liblinear feats + _meta.json
    load_boxes            nn_utils/data.py:224-271  "<mention|box id>\\t<0|1>" (+ optional box category file)
    load_sparse_feats     utils/data.py:114-220     "<label> <idx>:<val> ... # <id>" (1-based when the path contains "box")
    load_relation_labels  nn_utils/data.py:330-346  "<ij_id> <ji_id> <null|coref|subset_ij|subset_ji>"

Word embeddings: the reference loads the 3.6 GB GoogleNews word2vec binary through gensim (utils/Word2Vec.py:19); neither
is available offline, so an `Embeddings` table is passed in instead: from a GloVe-style text file, an .npz written by
`synth.write_dataset`, or gensim when it is importable.  Unknown words map to the 'UNK' row (Word2Vec.py:37-39).
"""
import json

import numpy as np

from .data import kv_str_to_dict

WORD_EMBEDDING_WIDTH = 300       # nn_utils/data.py:16
BOX_EMBEDDING_WIDTH = 4096       # nn_utils/data.py:17


class Embeddings(object):
    def __init__(self, vocab, matrix, unk=None):
        self.index = {w: i for i, w in enumerate(vocab)}
        self.matrix = np.asarray(matrix, dtype=np.float32)
        self.unk = np.asarray(self.matrix[self.index["UNK"]] if unk is None and "UNK" in self.index else
                              (unk if unk is not None else np.zeros(self.matrix.shape[1])), dtype=np.float32)

    @classmethod
    def from_npz(cls, path):
        z = np.load(path, allow_pickle=False)
        return cls([str(w) for w in z["vocab"]], z["matrix"])

    @classmethod
    def from_text(cls, path):
        """GloVe layout "<word> v0 v1 ..." (nn_utils/data.py:30-47); unknown words get one U(-1,1) vector (:68)."""
        vocab, rows = [], []
        with open(path, "r") as f:
            for line in f:
                p = line.rstrip("\n").split(" ")
                vocab.append(p[0])
                rows.append([float(x) for x in p[1:]])
        return cls(vocab, np.asarray(rows, np.float32), unk=np.random.uniform(-1, 1, len(rows[0])))

    @classmethod
    def from_word2vec_bin(cls, path):
        from gensim.models import KeyedVectors          # utils/Word2Vec.py:1,19
        kv = KeyedVectors.load_word2vec_format(path, binary=True)
        return cls(list(kv.index_to_key), kv.vectors)

    def sentence_matrix(self, words):
        out = np.empty((len(words), self.matrix.shape[1]), np.float32)
        for i, w in enumerate(words):
            j = self.index.get(w)
            out[i] = self.matrix[j] if j is not None else self.unk
        return out


def load_sentences(sentence_file, embeddings):
    """nn_utils/data.py:77-120.  As written, the reference splits the raw line -- `line.split("\\t")[1].split(" ")`, :92-93 --
    without stripping it, so the LAST token of every caption still carries its "\\n", is not found in the embedding model and
    becomes the UNK row (utils/Word2Vec.py:36-39).  Reproduced here token for token (pinned against the reference's own
    load_sentences by tests/test_ref_fixtures.py): a checkpoint trained by the reference saw UNK there."""
    data_dict = dict(sentences={})
    if sentence_file is not None:
        with open(sentence_file, "r") as f:
            for line in f:
                id_split = line.split("\t")
                data_dict["sentences"][id_split[0].strip()] = embeddings.sentence_matrix(id_split[1].split(" "))
    data_dict["max_seq_len"] = max([len(m) for m in data_dict["sentences"].values()] or [-1])
    data_dict["word_embedding_width"] = embeddings.matrix.shape[1]
    return data_dict


def load_sparse_feats(filename, meta_dict=None, n_features=None):
    n_feats = n_features if n_features is not None else meta_dict["max_idx"] + 1
    with open(filename, "r") as f:
        lines = f.readlines()
    x = np.zeros([len(lines), n_feats], np.float32)
    y = np.zeros(len(lines))
    ids = []
    shift = 1 if "box" in filename else 0                       # utils/data.py:208-209
    for i, line in enumerate(lines):
        body, cid = line.split(" # ")
        ids.append(cid.strip())
        parts = body.strip().split(" ")
        y[i] = int(float(parts[0].strip()))
        for tok in parts[1:]:
            k, v = tok.split(":")
            if float(v) != 0.0:
                x[i][int(k.strip()) - shift] = float(v.strip())
    return x, y, ids


def load_mentions(mention_idx_file, task, feats_file, feats_meta_file, n_classes):
    data_dict = dict(caption_ids={}, mention_indices={}, labels={})
    if mention_idx_file is not None:
        with open(mention_idx_file, "r") as f:
            for line in f:
                sp = line.strip().split("\t")
                mid = sp[0].strip()
                if "rel" in task:
                    d = kv_str_to_dict(mid)
                    data_dict["caption_ids"][mid] = (d["doc"] + "#" + d["caption_1"], d["doc"] + "#" + d["caption_2"])
                else:
                    data_dict["caption_ids"][mid] = mid.split(";")[0]
                data_dict["mention_indices"][mid] = [int(i.strip()) for i in sp[1].strip().split(",")]
                if task != "affinity":
                    label = np.zeros([n_classes])
                    label[int(sp[2].strip())] = 1.0
                    data_dict["labels"][mid] = label
    if feats_file is not None and feats_meta_file is not None:
        meta = json.load(open(feats_meta_file, "r"))
        data_dict["n_mention_feats"] = meta["max_idx"] + 1
        X, _, IDs = load_sparse_feats(feats_file, meta)
        data_dict["mention_features"] = {IDs[i]: X[i] for i in range(len(IDs))}
    return data_dict


def load_boxes(mention_box_label_file, box_dir, box_category_file=None):
    data_dict = dict(labels={}, box_categories={}, n_box_feats=None)
    with open(mention_box_label_file, "r") as f:
        for line in f:
            sp = line.strip().split("\t")
            v = np.zeros([2])
            v[int(sp[1].strip())] = 1.0
            data_dict["labels"][sp[0].strip()] = v
    if box_category_file is not None:
        with open(box_category_file, "r") as f:
            for line in f:
                sp = line.strip().split("\t")
                vec = np.array([float(i) for i in sp[1].split(",")])
                data_dict["n_box_feats"] = max(data_dict["n_box_feats"] or 0, len(vec))
                data_dict["box_categories"][sp[0].strip()] = vec
    data_dict["box_dir"] = box_dir
    data_dict["box_embedding_width"] = BOX_EMBEDDING_WIDTH
    return data_dict


def load_relation_labels(filename):
    gold = {}
    with open(filename, "r") as f:
        for line in f:
            sp = line.split(" ")
            gold[(sp[0].strip(), sp[1].strip())] = sp[2].strip()
    return gold


def load_all_boxes(data_dict):
    """Read every per-image box feature file of `box_dir` into data_dict['box_table'] ({box id: vector}) so that the box features
    can stay resident on the device (core.Session.set_box_table): 4096 floats per box, ~20 boxes per image -- Flickr30k-Entities is
    ~10 GB, which the 180 GB of a B200 hold easily; the reference re-parses the text files for every batch instead
    (nn_utils/data.py:506-524).  Opt-in (ICL_BOX_TABLE=1 in the drop-in CLIs) because it needs the same amount of host memory."""
    import os
    from . import data as nn_data
    table = {}
    imgs = sorted(set(k.split("|")[1].split(";")[0] for k in data_dict["labels"].keys()))
    for img in imgs:
        path = os.path.join(data_dict["box_dir"], img.replace(".jpg", ".feats"))
        if os.path.exists(path):
            table.update(nn_data.read_box_feats(path, data_dict["box_embedding_width"]))
    data_dict["box_table"] = table
    return data_dict
