"""Shared builders for oracle / parity tests: small seeded problems in the reference's batch_tensors form."""
import numpy as np

from oracle import icl_oracle as O


def tiny_problem(seed=0, S=6, T=7, E=5, H=3, F=4, task="nonvis", enc="first_last_mention", C=None,
                 widths=(8, 4), act="relu", box_w=0, dropout=False, keep_in=0.5, keep=0.5, data_norm=False,
                 weighted=False, dtype=np.float64):
    """One head on one batch.  For rel_cross S must be even (B = S/2 examples)."""
    rng = np.random.default_rng(seed)
    C = C or {"nonvis": 2, "card": 12, "rel_intra": 4, "rel_cross": 4, "affinity": 2}[task]
    lens = rng.integers(1, T + 1, S)
    lens[rng.integers(0, S)] = T
    x = np.zeros((S, T, E), dtype)
    for s in range(S):
        x[s, :lens[s]] = rng.standard_normal((lens[s], E))
    B = S // 2 if task == "rel_cross" else S
    hb = {}
    ar = np.arange(B)
    si, sj = (2 * ar, 2 * ar + 1) if task == "rel_cross" else (ar, ar)

    def span(sent):
        a = np.array([rng.integers(0, lens[s]) for s in sent])
        b = np.array([rng.integers(a[k], lens[s]) for k, s in enumerate(sent)])
        return a, b

    fi, li = span(si)
    fj, lj = span(sj)
    z, o = np.zeros(B, np.int64), np.ones(B, np.int64)
    st = lambda d, s, w: np.stack([d, s, w], 1).astype(np.float64)
    hb.update(first_i_fw=st(z, si, fi), first_i_bw=st(o, si, fi), last_i_fw=st(z, si, li), last_i_bw=st(o, si, li),
              first_j_fw=st(z, sj, fj), first_j_bw=st(o, sj, fj), last_j_fw=st(z, sj, lj), last_j_bw=st(o, sj, lj),
              sent_last_i_fw=st(z, si, lens[si] - 1), sent_first_i_bw=st(o, si, z),
              sent_last_j_fw=st(z, sj, lens[sj] - 1), sent_first_j_bw=st(o, sj, z))
    feats = (rng.random((B, F)) < 0.3).astype(dtype)
    hb["ij_feats" if "rel" in task else "m_feats"] = feats
    if task == "affinity":
        hb["box_embeddings"] = np.maximum(rng.standard_normal((B, box_w)) - 0.3, 0).astype(dtype)
    y = np.zeros((B, C), dtype)
    y[ar, rng.integers(0, C, B)] = 1.0
    hb["labels"] = y
    hb["sentences"], hb["seq_lengths"] = x, lens.astype(np.float64)
    in_w = O.head_in_width(task, enc, H, F, box_w)
    cfg = dict(H=H, data_norm=data_norm,
               heads=[dict(task=task, scope="", encoding_scheme=enc, n_layers=len(widths), widths=list(widths),
                           activation=act, weighted_classes=weighted, in_width=in_w, n_classes=C)])
    params = O.init_params(rng, cfg, E, dtype)
    for d in ("fw", "bw"):
        params[O.lstm_names(d)[1]] = (rng.standard_normal(4 * H) * 0.1).astype(dtype)
    masks = None
    if dropout:
        bern = lambda shape, k: (rng.random(shape) < k).astype(dtype)
        masks = dict(in_fw=bern((S, T, E), keep_in), in_bw=bern((S, T, E), keep_in),
                     out_fw=bern((S, T, H), keep), out_bw=bern((S, T, H), keep),
                     heads=[[bern((B, w), keep) for w in widths]])
    return dict(cfg=cfg, params=params, x=x, lens=lens, batch=hb, masks=masks, keep_in=keep_in if dropout else 1.0,
                keep=keep if dropout else 1.0, B=B, C=C, E=E, H=H, F=F, T=T, S=S)
