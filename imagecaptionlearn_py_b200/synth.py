"""Synthetic Flickr30k-Entities-shaped corpus ("F30kE-synth").

There is no dataset (and no network) in the build/benchmark environment, so benchmarks and tests run on a
seeded synthetic corpus that has the *shape* the reference's loaders produce (`nn_utils/data.py:77-195,
224-271`): 5 captions per image, caption length clip(1+Poisson(11.3),3,50), 300-d word vectors, mention
spans with inclusive word indices, sparse mention features, 20 boxes per image with fc7-like 4096-d rows.
Everything is drawn from numpy.random.Generator(PCG64(seed)); default seed = the reference's own constant
20171201 (`nn_utils/core.py:10`).

The returned `data_dict`s use exactly the reference's keys, so they can be fed to `data.load_batch`.
"""
import numpy as np

CAPS_PER_IMAGE = 5
N_CLASSES = {"nonvis": 2, "card": 12, "rel_intra": 4, "rel_cross": 4, "affinity": 2}


def make_corpus(n_images, seed=20171201, E=300, T_cap=50, vocab=20000, n_boxes=20, box_width=4096,
                with_boxes=False, dtype=np.float32):
    rng = np.random.Generator(np.random.PCG64(seed))
    table = (rng.standard_normal((vocab + 1, E)) * 0.15).astype(dtype)          # row `vocab` = UNK
    ranks = np.arange(1, vocab + 1, dtype=np.float64)
    zipf = ranks ** -1.1
    zipf /= zipf.sum()
    corpus = dict(E=E, T_cap=T_cap, images=[], sentences={}, word_ids={}, mentions={}, table=table,
                  n_boxes=n_boxes, box_width=box_width, boxes={})
    for im in range(n_images):
        img = "%07d.jpg" % im
        corpus["images"].append(img)
        for c in range(CAPS_PER_IMAGE):
            cap_id = "%s#%d" % (img, c)
            L = int(np.clip(1 + rng.poisson(11.3), 3, T_cap))
            wid = rng.choice(vocab, size=L, p=zipf)
            corpus["word_ids"][cap_id] = wid.astype(np.int32)
            corpus["sentences"][cap_id] = table[wid]
            K = int(np.clip(rng.poisson(3.2), 1, 8))
            spans, pos = [], 0
            for _ in range(K):
                ln = int(np.clip(1 + rng.poisson(1.3), 1, 6))
                gap = int(rng.integers(0, 3))
                first = pos + gap
                last = first + ln - 1
                if last >= L:
                    break
                spans.append((first, last))
                pos = last + 1
            if not spans:
                spans = [(0, min(L - 1, 1))]
            corpus["mentions"][cap_id] = spans
        if with_boxes:
            feats = np.maximum(rng.standard_normal((n_boxes, box_width)) - 0.7, 0).astype(dtype)
            for b in range(n_boxes):
                corpus["boxes"]["%s;box:%d" % (img, b)] = feats[b]
    corpus["rng"] = rng
    return corpus


def _sparse_feats(rng, F, dtype):
    return (rng.random(F) < 0.02).astype(dtype)


def _label(rng, task):
    if task == "nonvis":
        return int(rng.random() < 0.06)
    if task == "card":
        p = np.arange(1, 13, dtype=np.float64) ** -1.2
        return int(rng.choice(12, p=p / p.sum()))
    if task.startswith("rel"):
        return int(rng.choice(4, p=[0.80, 0.14, 0.03, 0.03]))
    return int(rng.random() < 0.08)


def make_data_dict(corpus, task, F=None, max_pairs_per_image=None, dtype=np.float32):
    """Build the reference-shaped data_dict for `task` (keys of nn_utils/data.py:77-195,224-271)."""
    rng = corpus["rng"]
    C = N_CLASSES[task]
    if F is None:
        F = 512 if task.startswith("rel") else 256
    dd = dict(sentences=corpus["sentences"], word_embedding_width=corpus["E"],
              max_seq_len=max(len(m) for m in corpus["sentences"].values()),
              caption_ids={}, mention_indices={}, labels={}, mention_features={}, n_mention_feats=F)

    def onehot(k):
        v = np.zeros([C])
        v[k] = 1.0
        return v

    if task in ("nonvis", "card", "affinity"):
        for cap_id, spans in corpus["mentions"].items():
            for k, (a, b) in enumerate(spans):
                mid = "%s;mention:%d" % (cap_id, k)
                dd["caption_ids"][mid] = cap_id
                dd["mention_indices"][mid] = [a, b]
                dd["mention_features"][mid] = _sparse_feats(rng, F, dtype)
                if task != "affinity":
                    dd["labels"][mid] = onehot(_label(rng, task))
        if task == "affinity":
            dd["box_embedding_width"] = corpus["box_width"]
            dd["box_categories"] = {}
            dd["n_box_feats"] = None
            dd["box_dir"] = None
            dd["box_table"] = corpus["boxes"]
            for mid in dd["mention_indices"]:
                img = mid.split("#")[0]
                for b in range(corpus["n_boxes"]):
                    dd["labels"]["%s|%s;box:%d" % (mid, img, b)] = onehot(_label(rng, task))
    else:
        for img in corpus["images"]:
            ments = [(c, k, s) for c in range(CAPS_PER_IMAGE)
                     for k, s in enumerate(corpus["mentions"]["%s#%d" % (img, c)])]
            n = 0
            for (c1, k1, s1) in ments:
                for (c2, k2, s2) in ments:
                    if (c1, k1) == (c2, k2) or (task == "rel_intra") != (c1 == c2):
                        continue
                    if max_pairs_per_image is not None and n >= max_pairs_per_image:
                        continue
                    pid = "doc:%s;caption_1:%d;mention_1:%d;caption_2:%d;mention_2:%d" % (img, c1, k1, c2, k2)
                    dd["caption_ids"][pid] = ("%s#%d" % (img, c1), "%s#%d" % (img, c2))
                    dd["mention_indices"][pid] = [s1[0], s1[1], s2[0], s2[1]]
                    dd["mention_features"][pid] = _sparse_feats(rng, F, dtype)
                    dd["labels"][pid] = onehot(_label(rng, task))
                    n += 1
    return dd


def example_ids(dd, task):
    if task == "affinity":
        loaded = set(dd["mention_indices"].keys())
        return [k for k in dd["labels"].keys() if k.split("|")[0] in loaded]     # icl_affinity_lstm.py:59-74
    return list(dd["mention_indices"].keys())


def write_dataset(corpus, data_dir, data_root, task, F=None, seed=7, naming="single"):
    """Emit the synthetic corpus in the reference's file formats and directory scheme (icl_core_lstm.py:333-345,
    icl_relation_lstm.py:403-430, icl_affinity_lstm.py:412-433) so the drop-in CLIs run on it unmodified:
    raw/<root>_captions.txt, raw/<root>_mentions_<task>.txt, feats/<root>_<task>_neural.feats + _meta.json
    (relations: raw/<root>_mentionPairs_<rel>.txt, feats/<root>_relation_neural_<rel>.feats, raw/<root>_mentionPair_labels.txt),
    raw/<root>_embeddings.npz (stand-in for the word2vec binary), and for affinity raw/<root>_affinity_labels.txt +
    feats/<data>_boxes/<split>/<img>.feats.
    naming="multitask": the file names icl_multitask_lstm.py:153-209 reads instead -- feats/<root>_<task>.feats (no "_neural"),
    ONE shared feats/<root>_relation.feats for both relation tasks (appended to), raw/<root>_mention_box_labels.txt."""
    import json
    import os
    dd = make_data_dict(corpus, task, F=F)
    F = dd["n_mention_feats"]
    for sub in ("raw", "feats", "scores"):
        os.makedirs(os.path.join(data_dir, sub), exist_ok=True)
    vocab = ["w%d" % i for i in range(len(corpus["table"]) - 1)] + ["UNK"]
    np.savez(os.path.join(data_dir, "raw", data_root + "_embeddings.npz"), vocab=np.array(vocab), matrix=corpus["table"])
    with open(os.path.join(data_dir, "raw", data_root + "_captions.txt"), "w") as f:
        for cap_id, wid in corpus["word_ids"].items():
            f.write("%s\t%s\n" % (cap_id, " ".join("w%d" % w for w in wid)))
    tag = task if not task.startswith("rel") else task.replace("rel_", "")
    ment_name = {"nonvis": "_mentions_nonvis.txt", "card": "_mentions_card.txt", "affinity": "_mentions_affinity.txt"}.get(
        task, "_mentionPairs_%s.txt" % tag)
    with open(os.path.join(data_dir, "raw", data_root + ment_name), "w") as f:
        for mid, idx in dd["mention_indices"].items():
            lab = int(np.argmax(dd["labels"][mid])) if mid in dd["labels"] else 0
            f.write("%s\t%s\t%d\n" % (mid, ",".join(str(i) for i in idx), lab))
    froot = "%s_%s_neural" % (data_root, task) if not task.startswith("rel") else "%s_relation_neural_%s" % (data_root, tag)
    if naming == "multitask":
        froot = "%s_%s" % (data_root, task) if not task.startswith("rel") else "%s_relation" % data_root
    with open(os.path.join(data_dir, "feats", froot + ".feats"), "a" if naming == "multitask" and task.startswith("rel") else "w") as f:
        for mid, v in dd["mention_features"].items():
            nz = np.nonzero(v)[0]
            f.write("0 %s # %s\n" % (" ".join("%d:%g" % (k, v[k]) for k in nz), mid))
    json.dump({"max_idx": F - 1}, open(os.path.join(data_dir, "feats", froot + "_meta.json"), "w"))
    if task == "affinity":
        lab_name = "_mention_box_labels.txt" if naming == "multitask" else "_affinity_labels.txt"
        with open(os.path.join(data_dir, "raw", data_root + lab_name), "w") as f:
            for k, v in dd["labels"].items():
                f.write("%s\t%d\n" % (k, int(np.argmax(v))))
        data, split = data_root.rsplit("_", 1)
        bdir = os.path.join(data_dir, "feats", data + "_boxes", split)
        os.makedirs(bdir, exist_ok=True)
        by_img = {}
        for bid, v in corpus["boxes"].items():
            by_img.setdefault(bid.split(";")[0], []).append((bid, v))
        for img, rows in by_img.items():
            with open(os.path.join(bdir, img.replace(".jpg", ".feats")), "w") as f:
                for bid, v in rows:
                    nz = np.nonzero(v)[0]
                    f.write("0 %s # %s\n" % (" ".join("%d:%g" % (k + 1, v[k]) for k in nz), bid))
    if task.startswith("rel"):
        names = ["null", "coref", "subset_ij", "subset_ji"]
        with open(os.path.join(data_dir, "raw", data_root + "_mentionPair_labels.txt"), "a") as f:
            for pid in dd["mention_indices"]:
                d = dict(kv.split(":")[:2] for kv in pid.split(";"))
                if (d["caption_1"], int(d["mention_1"])) < (d["caption_2"], int(d["mention_2"])):
                    ji = "doc:%s;caption_1:%s;mention_1:%s;caption_2:%s;mention_2:%s" % (
                        d["doc"], d["caption_2"], d["mention_2"], d["caption_1"], d["mention_1"])
                    f.write("%s %s %s\n" % (pid, ji, names[int(np.argmax(dd["labels"][pid]))]))
    return dd
