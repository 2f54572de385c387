"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every declared symbol; the product
never imports the oracle; creating a model without a GPU fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

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
