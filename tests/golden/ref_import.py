"""Imports the REAL reference modules from /root/reference under Python 3 (build container only; the GPU box has no
/root/reference).  Used by the fixture generators tests/golden/make_*.py and by tests that run the reference's own script
bodies against the shim; never by the product.

What it takes to run the Python-2-era sources under 3.12 without editing them on disk:
  * `gensim` is absent: a stub `gensim.models.KeyedVectors` whose `load_word2vec_format(path)` reads the synthetic embedding
    table (.npz written by synth.write_dataset) and answers `word in model` / `model[word]` like gensim's KeyedVectors.
  * `tensorflow` is absent: `sys.modules['tensorflow']` is whatever the caller passes (a module-like object); the data / eval
    modules never touch it.
  * implicit relative imports of Python 2 (`import core as util` in utils/data.py:4, `from Score import Score` in
    utils/ScoreDict.py:2) are satisfied by aliasing utils/core.py and utils/Score.py as top-level modules.
  * `print x` statements (nn_utils/eval.py:87-89,175-177, utils/ScoreDict.py:220, nn_utils/core.py:694) are rewritten to
    `print(x)` IN MEMORY before compilation; nothing else in those files is touched.
"""
import importlib.util
import os
import re
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# /root/reference in the build container; on a GPU box the UNMODIFIED copy that tools/stage_reference.sh puts under baseline/_ref/
# (git-ignored, travels with gpurun like the contract's own `pip install --target baseline/_ref`)
REF = os.environ.get("ICL_REF_DIR") or ("/root/reference" if os.path.isdir("/root/reference/nn_utils") else
                                        os.path.join(os.path.dirname(os.path.dirname(_HERE)), "baseline", "_ref", "ImageCaptionLearn_py"))


def available():
    return os.path.isdir(os.path.join(REF, "nn_utils"))


class _KeyedVectors(object):
    def __init__(self, vocab, matrix):
        self._idx = {w: i for i, w in enumerate(vocab)}
        self._m = matrix

    def __contains__(self, w):
        return w in self._idx

    def __getitem__(self, w):
        return self._m[self._idx[w]]

    @classmethod
    def load_word2vec_format(cls, path, binary=True):
        z = np.load(path, allow_pickle=False)
        return cls([str(w) for w in z["vocab"]], z["matrix"])


def _load_source(name, path, fix_print=False, patch=None):
    src = open(path).read()
    if fix_print:
        src = re.sub(r"(?m)^(\s*)print (.+)$", r"\1print(\2)", src)
    if patch:
        src = patch(src)
    mod = types.ModuleType(name)
    mod.__file__ = path
    sys.modules[name] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def install(tensorflow=None, nn_core=None):
    """Makes `nn_utils.data`, `nn_utils.eval`, `nn_utils.core`, `utils.*` importable from /root/reference.  `nn_core`: a module to
    register as `nn_utils.core` INSTEAD of the reference's (the import swap of INTEGRATION.md: the B200 shim); default = the
    reference's own file (its function bodies need a real TensorFlow, importing it does not).  Returns a dict of the modules."""
    if not available():
        raise RuntimeError("the reference sources are not present (%s)" % REF)
    gensim = types.ModuleType("gensim")
    gm = types.ModuleType("gensim.models")
    gm.KeyedVectors = _KeyedVectors
    gensim.models = gm
    sys.modules["gensim"], sys.modules["gensim.models"] = gensim, gm
    if tensorflow is not None:
        sys.modules["tensorflow"] = tensorflow
    for pkg in ("utils", "nn_utils"):
        p = types.ModuleType(pkg)
        p.__path__ = [os.path.join(REF, pkg)]
        sys.modules[pkg] = p
    out = {}
    out["core"] = _load_source("core", os.path.join(REF, "utils", "core.py"))                 # `import core as util`
    sys.modules["utils.core"] = out["core"]
    sys.modules["utils"].core = out["core"]
    out["Score"] = _load_source("Score", os.path.join(REF, "utils", "Score.py"))              # `from Score import Score`
    sys.modules["utils.Score"] = out["Score"]
    for name in ("string", "Logger", "data", "Word2Vec"):
        out["utils." + name] = _load_source("utils." + name, os.path.join(REF, "utils", name + ".py"))
        setattr(sys.modules["utils"], name, out["utils." + name])
    out["utils.ScoreDict"] = _load_source("utils.ScoreDict", os.path.join(REF, "utils", "ScoreDict.py"), fix_print=True)
    sys.modules["utils"].ScoreDict = out["utils.ScoreDict"]
    out["nn_utils.data"] = _load_source("nn_utils.data", os.path.join(REF, "nn_utils", "data.py"))
    sys.modules["nn_utils"].data = out["nn_utils.data"]
    out["nn_utils.eval"] = _load_source("nn_utils.eval", os.path.join(REF, "nn_utils", "eval.py"), fix_print=True)
    sys.modules["nn_utils"].eval = out["nn_utils.eval"]
    if nn_core is not None:
        sys.modules["nn_utils.core"] = nn_core
        out["nn_utils.core"] = nn_core
    else:
        out["nn_utils.core"] = _load_source("nn_utils.core", os.path.join(REF, "nn_utils", "core.py"), fix_print=True)
    sys.modules["nn_utils"].core = out["nn_utils.core"]
    return out


def uninstall():
    for k in list(sys.modules):
        if k in ("gensim", "gensim.models", "core", "Score", "utils", "nn_utils") or k.startswith(("utils.", "nn_utils.")):
            del sys.modules[k]


def load_script(name, patch=None, extra_modules=None):
    """Compiles one of the reference's root scripts (icl_core_lstm.py, ...) as module `ref_<name>` WITHOUT running its
    `__init__()` entry call.  `patch` may rewrite the source text in memory (Python-2 integer division)."""
    for k, v in (extra_modules or {}).items():
        sys.modules[k] = v

    def strip_entry(src):
        src = re.sub(r"(?m)^__init__\(\)\s*$", "", src)
        return patch(src) if patch else src
    return _load_source("ref_" + name, os.path.join(REF, name + ".py"), patch=strip_entry)
