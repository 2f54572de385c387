"""One small train step + one predict of the hot path for compute-sanitizer (memcheck / racecheck / synccheck):

    ICL_NO_TORCH_STREAM=1 compute-sanitizer --tool racecheck python tools/sanitize_run.py [S] [H] [T]

Reduced card workload (default S=256 captions, H=300, T<=20): every kernel of the step runs, including k_rec_fwd16 (flag-published
tiles) and k_bptt_cluster (cluster barriers + DSMEM pulls), at a size the tools finish in minutes."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ICL_NO_TORCH_STREAM", "1")
import numpy as np
from imagecaptionlearn_py_b200 import _cabi
from tests.helpers import tiny_problem
from tests.test_gpu_parity import make_session

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H = int(sys.argv[2]) if len(sys.argv) > 2 else 300
T = int(sys.argv[3]) if len(sys.argv) > 3 else 20
p = tiny_problem(seed=2, task="card", enc="first_last_mention", act="relu", S=S, T=T, E=300, H=H, F=64, widths=(128, 64), dropout=True)
core, sess = make_session(p, "tf32")
for i in range(2):
    r = sess.run(_cabi.OP_TRAIN, [dict(p["batch"])], p["keep_in"], p["keep"], True)[0]
    print("train step", i, "loss", float(r["loss"]))
r = sess.run(_cabi.OP_PREDICT, [dict(p["batch"])], 1.0, 1.0, True)[0]
print("predict loss", float(r["loss"]), "finite", bool(np.all(np.isfinite(r["proba"]))))
n = __import__("ctypes").c_int64()
_cabi.lib().icl_kernel_launches(sess.handle, __import__("ctypes").byref(n))
print("kernel launches", n.value)
sess.close()
