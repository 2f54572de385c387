"""Generates the committed fixtures under tests/golden/.  Run in the build container (it reads /root/reference):

    python tests/golden/make_golden.py

1. score_golden.json  -- produced by the REAL reference modules that still import under Python 3
   (/root/reference/utils/Score.py, /root/reference/utils/string.py): Score(p, r, f1, to_string) over a grid of counts
   and kv_str_to_dict over id strings.  These pin imagecaptionlearn_py_b200.eval.Score / data.kv_str_to_dict.
2. oracle_kat.npz     -- known-answer vectors of the NumPy oracle on small seeded problems (the reference's TensorFlow 1.x
   graph cannot run here, so these are regression anchors for the restatement, not reference outputs -- "parity unpinned").
"""
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = "/root/reference"
    Score = load(os.path.join(ref, "utils", "Score.py"), "ref_Score").Score
    ref_str = load(os.path.join(ref, "utils", "string.py"), "ref_string")
    cases = []
    for pred in (0, 1, 3, 10, 57):
        for gold in (0, 2, 8, 57):
            for corr in (0, 1, 6):
                if corr <= pred and corr <= gold or pred == 0 or gold == 0:
                    s = Score(predicted_count=pred, gold_count=gold, correct_count=corr)
                    cases.append(dict(pred=pred, gold=gold, correct=corr, p=s.p, r=s.r, f1=s.f1, text=s.to_string(),
                                      latex=s.to_latex_string()))
    s = Score(precision=0.25, recall=0.5)
    cases.append(dict(precision=0.25, recall=0.5, p=s.p, r=s.r, f1=s.f1, text=s.to_string(), latex=s.to_latex_string()))
    ids = ["doc:1000092795.jpg;caption_1:0;mention_1:1;caption_2:0;mention_2:3",
           "doc:27.jpg;caption_1:4;mention_1:0;caption_2:2;mention_2:7", "a:b;c:d:e"]
    kv = [dict(s=i, d=ref_str.kv_str_to_dict(i)) for i in ids]
    json.dump(dict(source="/root/reference/utils/Score.py + utils/string.py (imported unmodified)", score=cases, kv=kv),
              open(os.path.join(HERE, "score_golden.json"), "w"), indent=1)

    from oracle import icl_oracle as O
    from tests.helpers import tiny_problem
    out = {}
    for name, kw in (("nonvis", dict(seed=11, task="nonvis", act="relu", dropout=True)),
                     ("card_sent", dict(seed=12, task="card", enc="first_last_sentence", act="tanh", data_norm=True, S=5, T=6)),
                     ("rel_cross", dict(seed=13, task="rel_cross", enc="first_last_sentence", act="sigmoid", weighted=True, S=8)),
                     ("affinity", dict(seed=14, task="affinity", act="leaky_relu", box_w=6, dropout=True))):
        p = tiny_problem(**kw)
        f = O.model_forward(p["params"], p["cfg"], p["x"], p["lens"], [p["batch"]], p["keep_in"], p["keep"], p["masks"])
        g = O.model_backward(p["params"], p["cfg"], f, [p["batch"]])
        out[name + "/loss"] = np.float64(f["loss"])
        out[name + "/proba"] = f["heads"][0]["proba"]
        out[name + "/out_fw"] = f["out_fw"]
        out[name + "/out_bw"] = f["out_bw"]
        for k, v in g.items():
            out[name + "/grad/" + k] = v
    np.savez_compressed(os.path.join(HERE, "oracle_kat.npz"), **out)
    print("wrote", len(cases), "score cases,", len(out), "oracle arrays")


if __name__ == "__main__":
    main()
