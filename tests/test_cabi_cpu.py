"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every declared symbol; the product
never imports the oracle; creating a model without a GPU fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest


@pytest.fixture(scope="module")
def built(repo_root):
    import __graft_entry__ as ge
    ge.build()
    return True


def test_library_exports_every_header_symbol(built, repo_root):
    from imagecaptionlearn_py_b200 import _cabi
    header = open(os.path.join(repo_root, "include", "icl_b200.h")).read()
    declared = set(re.findall(r"\b(icl_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    L = _cabi.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert L.icl_version() >= 1


def test_struct_layouts_match_header(built):
    """sizeof() of the ctypes mirrors must equal what the compiler laid out (checked via a tiny C probe)."""
    import subprocess, tempfile, textwrap
    from imagecaptionlearn_py_b200 import _cabi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = textwrap.dedent("""
        #include <stdio.h>
        #include "icl_b200.h"
        int main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(icl_head_config), sizeof(icl_config), sizeof(icl_head_batch),
                          sizeof(icl_batch), sizeof(icl_head_out)); return 0;}""")
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), "-o", os.path.join(d, "p"), os.path.join(d, "p.c")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "p")]).split()]
    got = [C.sizeof(_cabi.HeadConfig), C.sizeof(_cabi.Config), C.sizeof(_cabi.HeadBatch), C.sizeof(_cabi.Batch),
           C.sizeof(_cabi.HeadOut)]
    assert got == sizes


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from imagecaptionlearn_py_b200 import core
    core.reset_default_graph()
    with core.variable_scope("bidirectional_lstm"):
        core.setup_bidirectional_lstm(8)
    core.setup_core_architecture("nonvis", "first_last_mention", 4, 8, 1, False, "relu", 2, 4)
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        core.Session().ensure()


def test_product_never_imports_oracle(repo_root):
    pkg = os.path.join(repo_root, "imagecaptionlearn_py_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "/root/reference" not in txt, f


def test_graph_layer_mirrors_reference_names():
    from imagecaptionlearn_py_b200 import core
    for fn in ("setup_bidirectional_lstm", "setup_core_architecture", "add_train_op", "run_op", "get_pred_scores_mcc",
               "get_widths", "set_random_seeds"):
        assert callable(getattr(core, fn))
    assert core.get_widths(512, 2) == [512, 256, 128]
    assert core.get_collection("nonvis/loss")[0].scope == "nonvis"


@pytest.mark.parametrize("src_dtype", ["float32", "float64"])
def test_host_row_packing_is_bit_exact(built, src_dtype):
    """icl_upload's host conversion of the sentence rows (icl_pack_rows = the same routines, no GPU needed): float32 out equals
    NumPy's cast (run_op's feed conversion, core.py:558-561), fp16 out (the half-width wire) equals NumPy's round-to-nearest-even
    float16 cast -- for every destination alignment / length remainder of the 64-byte streaming loop, and for values that hit
    ties, subnormal halves, overflow to inf and signed zeros."""
    import ctypes as C
    from imagecaptionlearn_py_b200 import _cabi
    L = _cabi.lib()
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(4099) * 0.15).astype(np.float32)
    x[:12] = [0.0, -0.0, 1.0, 1.0 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -11, 65504.0, 65520.0, 1e-7, -3e-8, 6.1e-5, 70000.0, -2.0 ** -25]   # ties, limits
    x = x.astype(src_dtype)
    if src_dtype == "float64":
        x[20:40] += rng.standard_normal(20) * 1e-9              # not representable in float32
    code = 0 if src_dtype == "float32" else 1
    for start in (0, 1, 3, 16, 31):
        for n in (0, 1, 15, 63, 64, 65, 300, 4099 - start):
            src = np.ascontiguousarray(x[start:start + n])
            for off in (0, 1, 5):                                 # destination misalignment in elements
                f32 = np.full(n + off + 8, 7.0, np.float32)
                assert L.icl_pack_rows(src.ctypes.data_as(C.c_void_p), code, n, C.c_void_p(f32.ctypes.data + 4 * off), 0) == 0
                assert np.array_equal(f32[off:off + n].view(np.uint32), src.astype(np.float32).view(np.uint32))
                assert np.all(f32[:off] == 7.0) and np.all(f32[off + n:] == 7.0)
                h16 = np.full(n + off + 8, 0x7777, np.uint16)
                assert L.icl_pack_rows(src.ctypes.data_as(C.c_void_p), code, n, C.c_void_p(h16.ctypes.data + 2 * off), 1) == 0
                with np.errstate(over="ignore"):
                    want = src.astype(np.float32).astype(np.float16).view(np.uint16)
                assert np.array_equal(h16[off:off + n], want), (start, n, off)
                assert np.all(h16[:off] == 0x7777) and np.all(h16[off + n:] == 0x7777)
    assert L.icl_pack_rows(None, 0, 4, None, 0) != 0


@pytest.mark.parametrize("width", [4096, 64, 33])
def test_distinct_box_rows_are_found_by_exact_comparison(built, width):
    """icl_upload finds the distinct box rows of an affinity batch (the factorised layer 1 runs once per distinct box)
    through icl_group_rows' routine: a SAMPLED hash picks a row's candidate group, a full comparison decides.  Rows that differ only
    where the hash does not look (32 eight-byte samples of a 16 KB row) must land in different groups; identical rows in the same
    one; groups are numbered in order of first appearance.  Host code only: no GPU."""
    import ctypes as C
    from imagecaptionlearn_py_b200 import _cabi
    L = _cabi.lib()
    rng = np.random.default_rng(width)
    n_distinct, B = 23, 300
    base = np.maximum(rng.standard_normal((n_distinct, width)) - 0.7, 0).astype(np.float32)       # fc7-like: ~3/4 zeros
    pick = rng.integers(0, n_distinct, B)
    rows = base[pick].copy()
    # adversarial twins: copies of an earlier row changed in ONE float that no hash sample covers (bytes 8..11 of a long row)
    twins = [40, 41, 150]
    for t in twins:
        rows[t] = rows[3]
        rows[t, 2] += 1.0 + t
    rows[299] = rows[40]                                                                          # and a true duplicate of a twin
    want, seen = np.empty(B, np.int32), []
    for r in range(B):
        for g, q in enumerate(seen):
            if np.array_equal(rows[q].view(np.uint32), rows[r].view(np.uint32)):
                want[r] = g
                break
        else:
            want[r] = len(seen)
            seen.append(r)
    got, n = np.full(B, -1, np.int32), C.c_int32(-1)
    assert L.icl_group_rows(rows.ctypes.data_as(C.c_void_p), B, width * 4, got.ctypes.data_as(C.c_void_p), C.byref(n)) == 0
    assert n.value == len(seen) and np.array_equal(got, want)
    assert len({got[t] for t in twins} | {got[3]}) == 4 and got[299] == got[40]
    # -0.0 and 0.0 are different bytes: separate groups (an extra group, never a wrong merge)
    z = np.zeros((2, width), np.float32)
    z[1, 0] = -0.0
    assert L.icl_group_rows(z.ctypes.data_as(C.c_void_p), 2, width * 4, got.ctypes.data_as(C.c_void_p), C.byref(n)) == 0 and n.value == 2
    assert L.icl_group_rows(None, 2, 8, got.ctypes.data_as(C.c_void_p), C.byref(n)) != 0
