"""Integer/index oracle for batch marshalling.  TEST INFRASTRUCTURE ONLY.

Slow, loop-for-loop py3 restatement of what `nn_utils/data.py:349-528` (load_batch) computes, plus the
task-level enumeration helpers (`icl_relation_lstm.py:213-245`, `icl_affinity_lstm.py:24-74`) and the
edge-padding of `nn_utils/core.py:656-659`.  The product's vectorised builder
(`imagecaptionlearn_py_b200/data.py`) must match this bit-for-bit on indices, lengths and labels.
Parity unpinned by the reference (it has no fixtures); pinned by property tests in tests/test_data.py.
"""
import numpy as np

INDEX_NAMES = ("first_i_bw", "first_i_fw", "last_i_fw", "last_i_bw", "sent_last_i_fw", "sent_first_i_bw",
               "first_j_bw", "last_j_fw", "first_j_fw", "last_j_bw", "sent_last_j_fw", "sent_first_j_bw")


def load_batch(ids, data_dict, task, n_classes, box_lookup=None):
    B = len(ids)
    n_seq = 2 * B if task == "rel_cross" else B                      # data.py:368-372
    T, E = data_dict["max_seq_len"], data_dict["word_embedding_width"]
    out = {"sentences": np.zeros([n_seq, T, E]), "seq_lengths": np.zeros([n_seq])}
    row = 0
    for i in range(B):                                               # data.py:379-404
        key = ids[i].split("|")[0] if task == "affinity" else ids[i]
        caps = data_dict["caption_ids"][key]
        if task == "rel_cross":
            sids = list(caps)
        elif task == "rel_intra":
            sids = [caps[0]]
        else:
            sids = [caps]
        for sid in sids:
            mat = data_dict["sentences"][sid]
            for w in range(len(mat)):
                out["sentences"][row][w] = mat[w]
            out["seq_lengths"][row] = len(mat)
            row += 1
    out["labels"] = np.zeros([B, n_classes])
    for name in INDEX_NAMES:                                         # data.py:409-415
        out[name] = np.zeros([B, 3])
    feat_key = "ij_feats" if "rel" in task else "m_feats"
    first_m = ids[0].split("|")[0] if task == "affinity" else ids[0]
    out[feat_key] = np.zeros([B, len(data_dict["mention_features"][first_m])])
    if task == "affinity":
        out["box_embeddings"] = np.zeros([B, data_dict["box_embedding_width"]])
        if data_dict.get("box_categories"):
            out["b_feats"] = np.zeros([B, data_dict["n_box_feats"]])
    for i in range(B):                                               # data.py:440-525
        label_id = ids[i]
        m_id, b_id = label_id, None
        if task == "affinity":
            m_id, b_id = label_id.split("|")
        out["labels"][i] = data_dict["labels"][label_id]
        mi = data_dict["mention_indices"][m_id]
        if "rel" in task:
            first_i, last_i, first_j, last_j = mi
        else:
            (first_i, last_i), first_j, last_j = mi, None, None
        si, sj = (2 * i, 2 * i + 1) if task == "rel_cross" else (i, i)
        out["first_i_fw"][i] = [0, si, first_i]
        out["first_i_bw"][i] = [1, si, first_i]
        out["last_i_fw"][i] = [0, si, last_i]
        out["last_i_bw"][i] = [1, si, last_i]
        if first_j is not None:
            out["first_j_fw"][i] = [0, sj, first_j]
            out["first_j_bw"][i] = [1, sj, first_j]
            out["last_j_fw"][i] = [0, sj, last_j]
            out["last_j_bw"][i] = [1, sj, last_j]
        out["sent_last_i_fw"][i] = [0, si, out["seq_lengths"][si] - 1]
        out["sent_first_i_bw"][i] = [1, si, 0]
        out["sent_last_j_fw"][i] = [0, sj, out["seq_lengths"][sj] - 1]
        out["sent_first_j_bw"][i] = [1, sj, 0]
        out[feat_key][i] = data_dict["mention_features"][m_id]
        if task == "affinity":
            if data_dict.get("box_categories") and b_id in data_dict["box_categories"]:
                out["b_feats"][i] = data_dict["box_categories"][b_id]
            out["box_embeddings"][i] = box_lookup(b_id)
    return out


def pad_ids_for_predict(ids, batch_size):
    """core.py:656-659 with python-2 integer division: always pads >=1, a full batch when n % B == 0."""
    n = len(ids)
    pad = batch_size * (n // batch_size + 1) - n
    arr = np.pad(np.asarray(ids, dtype=object), (0, pad), "edge")
    return arr.reshape([-1, batch_size]), pad


def get_ij_pairs(pair_ids):
    """icl_relation_lstm.py:213-222: same caption and mention_1 < mention_2."""
    keep = []
    for pid in pair_ids:
        d = dict(kv.split(":") for kv in pid.split(";"))
        if d["caption_1"] == d["caption_2"] and int(d["mention_1"]) < int(d["mention_2"]):
            keep.append(pid)
    return keep


def induce_ji(pred_scores):
    """icl_relation_lstm.py:225-245: ji id swaps (caption,mention) 1<->2, scores swap classes 2<->3."""
    out = {}
    for pid, s in pred_scores.items():
        d = dict(kv.split(":") for kv in pid.split(";"))
        ji = "doc:%s;caption_1:%s;mention_1:%s;caption_2:%s;mention_2:%s" % (
            d["doc"], d["caption_2"], d["mention_2"], d["caption_1"], d["mention_1"])
        out[ji] = np.array([s[0], s[1], s[3], s[2]])
    return out
