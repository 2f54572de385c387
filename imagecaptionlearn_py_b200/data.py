"""Host-side batch marshalling: the py3 mirror of `nn_utils/data.py` for the hot path.

`load_batch` returns the *same dict* (`batch_tensors`) the reference's `nn_utils/data.py:349-528` returns --
same keys, shapes and integer contents (verified bit-for-bit against a loop restatement in the tests) -- but
is built with vectorised NumPy gathers from a flat token table instead of per-word Python row copies, and in
float32/int32 (what TensorFlow converts the feed to anyway, `nn_utils/core.py:288,294,348`).

Also here: `build_model_filename` (`nn_utils/data.py:274-327`), the predict-time edge padding
(`nn_utils/core.py:656-659`), and the relation / affinity enumeration helpers
(`icl_relation_lstm.py:213-245`, `icl_affinity_lstm.py:24-74`).
"""
import numpy as np

INDEX_NAMES = ("first_i_bw", "first_i_fw", "last_i_fw", "last_i_bw", "sent_last_i_fw", "sent_first_i_bw",
               "first_j_bw", "last_j_fw", "first_j_fw", "last_j_bw", "sent_last_j_fw", "sent_first_j_bw")


def _flat_table(data_dict):
    """Cache: all caption matrices concatenated into one [N,E] float32 array + id -> (offset,len)."""
    ft = data_dict.get("_flat")
    if ft is None or ft[2] != len(data_dict["sentences"]):
        offs, parts, pos = {}, [], 0
        for sid, mat in data_dict["sentences"].items():
            offs[sid] = (pos, len(mat))
            parts.append(np.asarray(mat, dtype=np.float32))
            pos += len(mat)
        E = data_dict["word_embedding_width"]
        flat = np.concatenate(parts, 0) if parts else np.zeros((0, E), np.float32)
        ft = (flat, offs, len(data_dict["sentences"]))
        data_dict["_flat"] = ft
    return ft


def _flat_index(data_dict, offs):
    """Cache: ({caption id: row}, [N,2] int64 (offset, length) of every caption in the flat token table)."""
    fi = data_dict.get("_flat_index")
    if fi is None or fi[0] is not offs:
        ids = list(offs.keys())
        fi = (offs, {c: i for i, c in enumerate(ids)}, np.array([offs[c] for c in ids], dtype=np.int64).reshape(len(ids), 2))
        data_dict["_flat_index"] = fi
    return fi[1], fi[2]


_TABLE_BYTES_MAX = 1 << 30       # per-corpus matrices below are only built when they stay under this (else: per-batch stacking)


def _mention_table(data_dict):
    """Cache: ({mention id: row}, mention_indices [N,2|4] int32, mention_features [N,F] float32) -- a batch then takes its rows with
    two fancy-index gathers instead of stacking B small arrays out of two dicts (np.stack of 2048 feature vectors: 7 ms, more than
    five device steps).  None when the dicts are ragged / do not cover each other / would not fit: load_batch then stacks."""
    feats, mind = data_dict["mention_features"], data_dict["mention_indices"]
    sig = (len(feats), len(mind), id(feats), id(mind))
    mt = data_dict.get("_mention_table")
    if mt is None or mt[0] != sig:
        tab = None
        try:
            ids = list(mind.keys())
            width = len(next(iter(feats.values()))) if feats else 0
            if ids and len(ids) * max(width, 1) * 4 <= _TABLE_BYTES_MAX and all(m in feats for m in ids):
                MI = np.array([mind[m] for m in ids], dtype=np.int32)
                F = np.stack([feats[m] for m in ids]).astype(np.float32)
                if MI.ndim == 2 and F.ndim == 2:
                    tab = ({m: i for i, m in enumerate(ids)}, MI, F)
        except (ValueError, TypeError):
            tab = None
        mt = (sig, tab)
        data_dict["_mention_table"] = mt
    return mt[1]


def _label_table(data_dict, n_classes):
    """Cache: ({example id: row}, labels [N, n_classes] float32), same idea as _mention_table."""
    labels = data_dict["labels"]
    sig = (len(labels), id(labels), n_classes)
    lt = data_dict.get("_label_table")
    if lt is None or lt[0] != sig:
        tab = None
        try:
            if labels and len(labels) * n_classes * 4 <= _TABLE_BYTES_MAX:
                ids = list(labels.keys())
                Y = np.stack([labels[i] for i in ids]).astype(np.float32).reshape(len(ids), n_classes)
                tab = ({e: i for i, e in enumerate(ids)}, Y)
        except (ValueError, TypeError):
            tab = None
        lt = (sig, tab)
        data_dict["_label_table"] = lt
    return lt[1]


def default_packing():
    """What the drop-in drivers ask `load_batch` for: "rows" (device-resident token table, 4 bytes per token on the wire) unless
    ICL_HOST_SENTENCES=1 asks for the reference's padded [S,T,300] host tensor."""
    import os
    return False if os.environ.get("ICL_HOST_SENTENCES") else "rows"


def load_batch(ids, data_dict, task, n_classes, packed=False, dedup=False):
    """Vectorised equivalent of nn_utils/data.py:349-528.

    packed=True returns 'sentences_packed' [sum(len),E] (caption-major valid tokens) instead of materialising the zero-padded
    'sentences' tensor; packed="rows" returns 'token_rows' (int32 [sum(len)], row numbers into 'token_table' = the corpus'
    caption matrices concatenated, which `core.Session` keeps resident on the device).  `core.run_op` understands all three;
    the integer index matrices, lengths, features and labels are identical in every mode.

    dedup=True encodes every distinct caption of the batch ONCE: the reference gives each mention / pair / mention-box example its
    own copy of its caption (nn_utils/data.py:367-403; an affinity batch repeats a caption ~60 times).  The sentence tensors then
    hold the distinct captions in order of first use and the `sent` column of every index matrix points at them.  Exact whenever
    the keep probabilities are 1.0 (prediction): the copies' LSTM outputs are identical.  Not for training: the reference draws
    an independent dropout mask per copy."""
    B = len(ids)
    cross = task == "rel_cross"
    n_seq = 2 * B if cross else B
    T, E = data_dict["max_seq_len"], data_dict["word_embedding_width"]
    flat, offs, _ = _flat_table(data_dict)
    cap_of = data_dict["caption_ids"]

    m_ids, b_ids = list(ids), None
    if task == "affinity":
        split = [s.split("|") for s in ids]
        m_ids = [s[0] for s in split]
        b_ids = [s[1] for s in split]
    if cross:
        sids = [c for m in m_ids for c in cap_of[m]]
    elif task == "rel_intra":
        sids = [cap_of[m][0] for m in m_ids]
    else:
        sids = [cap_of[m] for m in m_ids]
    seq_of = np.arange(n_seq)                                    # index-matrix sentence number -> row of the sentence tensors
    if dedup:
        first = {}
        seq_of = np.array([first.setdefault(s, len(first)) for s in sids], dtype=np.int64)
        sids = list(first.keys())
        n_seq = len(sids)
    crow, OL = _flat_index(data_dict, offs)
    ol = OL[[crow[s] for s in sids]].reshape(n_seq, 2)
    lens = ol[:, 1]
    rows = np.repeat(np.arange(n_seq), lens)
    starts = np.cumsum(lens) - lens
    cols = np.arange(int(lens.sum())) - np.repeat(starts, lens)
    src = np.repeat(ol[:, 0], lens) + cols

    out = {}
    if packed == "rows":
        out["token_rows"] = src.astype(np.int32)
        out["token_table"] = flat
    elif packed:
        out["sentences_packed"] = flat[src]
    else:
        sent = np.zeros([n_seq, T, E], np.float32)
        sent[rows, cols] = flat[src]
        out["sentences"] = sent
    out["seq_lengths"] = lens.astype(np.int32)

    lt, mt = _label_table(data_dict, n_classes), _mention_table(data_dict)
    if lt is not None:
        out["labels"] = lt[1][[lt[0][i] for i in ids]]
    else:
        out["labels"] = np.stack([data_dict["labels"][i] for i in ids]).astype(np.float32).reshape(B, n_classes)
    m_rows = [mt[0][m] for m in m_ids] if mt is not None else None
    mi = mt[1][m_rows] if mt is not None else np.array([data_dict["mention_indices"][m] for m in m_ids], dtype=np.int32)
    ar = np.arange(B, dtype=np.int32)
    si, sj = (2 * ar, 2 * ar + 1) if cross else (ar, ar)
    si, sj = seq_of[si].astype(np.int32), seq_of[sj].astype(np.int32)
    zeros, ones = np.zeros(B, np.int32), np.ones(B, np.int32)

    def rows3(d, s, w):
        return np.stack([d, s, w.astype(np.int32)], 1)

    for name in INDEX_NAMES:
        out[name] = np.zeros([B, 3], np.int32)
    out["first_i_fw"] = rows3(zeros, si, mi[:, 0])
    out["first_i_bw"] = rows3(ones, si, mi[:, 0])
    out["last_i_fw"] = rows3(zeros, si, mi[:, 1])
    out["last_i_bw"] = rows3(ones, si, mi[:, 1])
    if "rel" in task:
        out["first_j_fw"] = rows3(zeros, sj, mi[:, 2])
        out["first_j_bw"] = rows3(ones, sj, mi[:, 2])
        out["last_j_fw"] = rows3(zeros, sj, mi[:, 3])
        out["last_j_bw"] = rows3(ones, sj, mi[:, 3])
    out["sent_last_i_fw"] = rows3(zeros, si, lens[si] - 1)
    out["sent_first_i_bw"] = rows3(ones, si, zeros)
    out["sent_last_j_fw"] = rows3(zeros, sj, lens[sj] - 1)
    out["sent_first_j_bw"] = rows3(ones, sj, zeros)

    feats = mt[2][m_rows] if mt is not None else np.stack([data_dict["mention_features"][m] for m in m_ids]).astype(np.float32)
    out["ij_feats" if "rel" in task else "m_feats"] = feats
    if task == "affinity":
        bm = _box_matrix(data_dict) if packed == "rows" else None
        if bm is not None:                    # device-resident box table: one row number per pair
            out["box_rows"] = np.array([bm[1][b] for b in b_ids], dtype=np.int32)
            out["box_table"] = bm[0]
        else:
            out["box_embeddings"] = _box_rows(b_ids, data_dict)
        if data_dict.get("box_categories"):
            bf = np.zeros([B, data_dict["n_box_feats"]], np.float32)
            for i, b in enumerate(b_ids):
                if b in data_dict["box_categories"]:
                    bf[i] = data_dict["box_categories"][b]
            out["b_feats"] = bf
    return out


def _box_matrix(data_dict):
    """(matrix [n_boxes, W] float32, {box id: row}) when every box of the data set is in memory ('box_table': {id: vector},
    e.g. loaders.load_all_boxes), else None (the per-image files are streamed per batch like the reference does)."""
    table = data_dict.get("box_table")
    if not table:
        return None
    bm = data_dict.get("_box_matrix")
    if bm is None or bm[2] != len(table):
        ids = list(table.keys())
        mat = np.stack([np.asarray(table[b], dtype=np.float32) for b in ids]) if ids else np.zeros((0, data_dict["box_embedding_width"]), np.float32)
        bm = (mat, {b: i for i, b in enumerate(ids)}, len(table))
        data_dict["_box_matrix"] = bm
    return bm


def _box_rows(b_ids, data_dict):
    """Box feature rows: from the in-memory table when present, else the per-image liblinear `.feats` files
    the reference streams (nn_utils/data.py:506-524, utils/data.py:163-217; 1-based indices shifted by -1)."""
    W = data_dict["box_embedding_width"]
    table = data_dict.get("box_table")
    out = np.zeros([len(b_ids), W], np.float32)
    cache = data_dict.setdefault("_box_cache", {})
    for i, b in enumerate(b_ids):
        if table is not None and b in table:
            out[i] = table[b]
            continue
        if b not in cache:
            cache.clear()
            img = b.split(";")[0]
            cache.update(read_box_feats(data_dict["box_dir"] + "/" + img.replace(".jpg", ".feats"), W))
        out[i] = cache[b]
    return out


def read_box_feats(path, width):
    """liblinear line: '<label> <idx>:<val> ... # <id>' with 1-based indices (utils/data.py:163-217)."""
    rows = {}
    with open(path, "r") as f:
        for line in f:
            body, _, cid = line.partition("#")
            vec = np.zeros(width, np.float32)
            for tok in body.split()[1:]:
                k, _, v = tok.partition(":")
                vec[int(k) - 1] = float(v)
            rows[cid.strip()] = vec
    return rows


def pad_ids_for_predict(ids, batch_size):
    """nn_utils/core.py:656-659: pad = B*(n//B + 1) - n (>=1, a whole batch when n % B == 0), mode 'edge'."""
    n = len(ids)
    pad = batch_size * (n // batch_size + 1) - n
    arr = np.pad(np.asarray(list(ids), dtype=object), (0, pad), "edge")
    return arr.reshape([-1, batch_size]), pad


def kv_str_to_dict(s):
    """utils/string.py:132-145: 'k:v;k:v' -> dict (the value is the SECOND ':'-field, anything after it is dropped)."""
    out = {}
    for kv in s.split(";"):
        f = kv.split(":")
        out[f[0]] = f[1]
    return out


def get_ij_pairs(mention_pairs):
    """icl_relation_lstm.py:213-222."""
    out = []
    for p in mention_pairs:
        d = kv_str_to_dict(p)
        if d["caption_1"] == d["caption_2"] and int(d["mention_1"]) < int(d["mention_2"]):
            out.append(p)
    return out


def induce_ji_predictions(pred_scores):
    """icl_relation_lstm.py:225-245: add the mirrored pair with classes 2<->3 swapped (in place)."""
    for ij in list(pred_scores.keys()):
        d = kv_str_to_dict(ij)
        ji = "doc:%s;caption_1:%s;mention_1:%s;caption_2:%s;mention_2:%s" % (
            d["doc"], d["caption_2"], d["mention_2"], d["caption_1"], d["mention_1"])
        s = np.array(pred_scores[ij], dtype=np.float64)
        s[[2, 3]] = s[[3, 2]]
        pred_scores[ji] = s
    return pred_scores


def get_valid_mention_box_pairs(data_dict):
    """icl_affinity_lstm.py:59-74."""
    loaded = set(data_dict["mention_indices"].keys())
    return [mb for mb in data_dict["labels"].keys() if mb.split("|")[0] in loaded]


def shuffle_mention_box_pairs(mention_box_pairs, rng=np.random):
    """icl_affinity_lstm.py:24-56: images in random order, pairs shuffled within an image, grouped by image."""
    by_img = {}
    for mb in mention_box_pairs:
        by_img.setdefault(mb.split("#")[0], []).append(mb)
    imgs = list(by_img.keys())
    rng.shuffle(imgs)
    out = []
    for im in imgs:
        rng.shuffle(by_img[im])
        out.extend(by_img[im])
    return out


def build_model_filename(arg_dict, task):
    """nn_utils/data.py:274-327 -- hyper-parameters encoded in the model file name."""
    name = arg_dict["data_root"] if "data_root" in arg_dict else arg_dict["data"] + "_" + arg_dict["split"]
    if arg_dict.get("rel_type") is not None:
        name += "_" + task.replace("_", "_" + arg_dict["rel_type"] + "_")
    else:
        name += "_" + task
    enc = arg_dict.get("encoding_scheme")
    name += {"first_last_sentence": "_fls", "first_last_mention": "_flm"}.get(enc, "")
    name += "_%s_epch%d_lrn%s_btch%d_drp%d%d_lstm%d_hdn%d-%d_admEps%s" % (
        arg_dict["activation"], int(arg_dict["epochs"]), str(arg_dict["learn_rate"]),
        int(arg_dict["batch_size"]), int(arg_dict["lstm_input_dropout"] * 100), int(arg_dict["dropout"] * 100),
        int(arg_dict["lstm_hidden_width"]), int(arg_dict["start_hidden_width"]), int(arg_dict["hidden_depth"]),
        str(arg_dict["adam_epsilon"]))
    if arg_dict.get("clip_norm") is not None:
        name += "_clip" + str(arg_dict["clip_norm"])
    if arg_dict.get("data_norm"):
        name += "_dataNorm"
    if arg_dict.get("weighted_classes"):
        name += "_weighted"
    if arg_dict.get("early_stopping"):
        name += "_early"
    return name + ".model"
