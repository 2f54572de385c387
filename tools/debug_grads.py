"""Bring-up probe: per-tensor gradient error (vs the oracle) of the tf32 and simt modes for the larger parity cases."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagecaptionlearn_py_b200 import _cabi
from oracle import icl_oracle as O
from tests.helpers import tiny_problem
from tests.test_gpu_parity import make_session, relerr, CASES, IDS

for ci in [int(a) for a in sys.argv[1:]] or [5, 6]:
    case = CASES[ci]
    p = tiny_problem(seed=22, dropout=False, **case)
    f = O.model_forward(p["params"], p["cfg"], p["x"], p["lens"], [p["batch"]], 1.0, 1.0, None)
    g = O.model_backward(p["params"], p["cfg"], f, [p["batch"]])
    for mode in ("simt", "tf32"):
        core, sess = make_session(p, mode)
        r = sess.run(_cabi.OP_GRADS, [dict(p["batch"])], 1.0, 1.0, True)[0]
        print("==", IDS[ci], mode, "loss", r["loss"], f["loss"], "proba err", relerr(r["proba"], f["heads"][0]["proba"]))
        for name, ref in g.items():
            got = sess.get_tensor(name, 1).reshape(ref.shape)
            print("   %-70s %.3e   max|ref| %.3e" % (name, relerr(got, ref), np.max(np.abs(ref))))
        sess.close()
