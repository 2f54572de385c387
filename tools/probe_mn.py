"""Bring-up probe (GPU box): tcgen05 TF32 GEMM for every operand major-ness, printing rel. error vs numpy.
MN-major descriptor parameters come from ICL_MN_* env vars (see csrc/gemm_tcgen05.cuh)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagecaptionlearn_py_b200 import _cabi  # noqa: E402
from tests.helpers import tiny_problem  # noqa: E402
from tests.test_gpu_parity import make_session, relerr  # noqa: E402

p = tiny_problem(E=8, H=4)
core, sess = make_session(p, "tf32")
out = []
for (M, N, K) in [(128, 128, 32), (256, 384, 96), (300, 1200, 300)]:
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    Bm = rng.standard_normal((K, N)).astype(np.float32)
    ref = A.astype(np.float64) @ Bm.astype(np.float64)
    for a_mn, b_mn in [(0, 0), (0, 1), (1, 0), (1, 1)]:
        As = np.ascontiguousarray(A.T if a_mn else A)
        Bs = np.ascontiguousarray(Bm if b_mn else Bm.T)
        Cc = np.zeros((M, N), np.float32)
        try:
            _cabi.check(_cabi.lib().icl_gemm(sess.handle, _cabi.GEMM_TCGEN05_TF32, a_mn, b_mn, M, N, K, _cabi.np_ptr(As),
                                             _cabi.np_ptr(Bs), _cabi.np_ptr(Cc), 1))
            e = relerr(Cc, ref)
        except RuntimeError as ex:
            e = str(ex)[:60]
        out.append("%dx%dx%d a_mn=%d b_mn=%d err=%s" % (M, N, K, a_mn, b_mn, e))
print("ENV", {k: v for k, v in os.environ.items() if k.startswith("ICL_MN")})
print("\n".join(out))
