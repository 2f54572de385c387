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
