"""Bring-up: per-tensor gradient error of the smoke problem under the operand-precision knobs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from imagecaptionlearn_py_b200 import _cabi
from oracle import icl_oracle as O
from tests.helpers import tiny_problem
from tests.test_gpu_parity import make_session, device_masks, kink_override, relerr
for seed in (1, 2, 3):
    p = tiny_problem(seed=seed, task="nonvis", enc="first_last_mention", act="relu", S=64, T=16, E=300, H=300, F=32, widths=(128, 64), dropout=True)
    core, sess = make_session(p, "tf32")
    r = sess.run(_cabi.OP_GRADS, [dict(p["batch"])], p["keep_in"], p["keep"], True)[0]
    masks = device_masks(sess, p)
    f = O.model_forward(p["params"], p["cfg"], p["x"], p["lens"], [p["batch"]], p["keep_in"], p["keep"], masks)
    over, flips = kink_override(sess, p, f, masks, 1e-3)
    g = O.model_backward(p["params"], p["cfg"], f, [p["batch"]], over)
    errs = {k.split("/")[-3] + "/" + k.split("/")[-1] if "lstm" in k else k: relerr(sess.get_tensor(k, 1).reshape(v.shape), v) for k, v in g.items()}
    print(seed, os.environ.get("ICL_REC_FP16"), os.environ.get("ICL_K1_FP16"), "flips", flips, {k: "%.1e" % v for k, v in errs.items()})
    sess.close()
