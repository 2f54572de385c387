"""Host logic of the shim that needs no GPU: feed validation of build_batch (the reference's placeholders have static shapes,
nn_utils/core.py:348,476 -- a wrong batch must raise, not be read out of bounds by the library)."""
import numpy as np
import pytest

from imagecaptionlearn_py_b200 import core
from tests.helpers import tiny_problem


def _session(p, task="nonvis", B=None):
    core.reset_default_graph()
    with core.variable_scope("bidirectional_lstm"):
        core.setup_bidirectional_lstm(p["H"], False, n_embedding_width=p["E"])
    core.setup_core_architecture(task, "first_last_mention", B or p["B"], 8, 1, False, "relu", p["C"], p["F"])
    return core.Session(max_seq_len=p["T"], device=0)


def test_build_batch_accepts_the_reference_dict_and_rejects_wrong_shapes():
    p = tiny_problem(seed=1, S=6, T=7, E=8, H=4, F=4)
    sess = _session(p)
    keep = []
    b = sess.build_batch([dict(p["batch"])], True, keep)               # the unchanged load_batch dict: fine
    assert b.n_seqs == 6 and b.padded_T == 7
    for key, bad in (("first_i_fw", p["batch"]["first_i_fw"][:-1]),    # a short tail batch (static B in the reference's graph)
                     ("m_feats", p["batch"]["m_feats"][:, :-1]),       # wrong feature width
                     ("labels", p["batch"]["labels"][:, :1]),          # wrong class count
                     ("sentences", p["batch"]["sentences"][:, :, :-1])):   # wrong embedding width
        bt = dict(p["batch"])
        bt[key] = bad
        with pytest.raises(ValueError, match=key):
            sess.build_batch([bt], True, [])
    bt = dict(p["batch"])                                              # packed sentences whose row count disagrees with the lengths
    bt.pop("sentences")
    bt["sentences_packed"] = np.zeros((int(p["lens"].sum()) - 1, p["E"]), np.float32)
    with pytest.raises(ValueError, match="sentences_packed"):
        sess.build_batch([bt], True, [])
    sess2 = _session(p, B=p["B"] - 1)                                   # more sequences than the graph was built for
    with pytest.raises(ValueError):
        sess2.build_batch([dict(p["batch"])], True, [])


def test_shim_exposes_the_tensorflow_names_the_reference_scripts_touch():
    """`import tensorflow as tf` -> this module: every tf.* symbol icl_core_lstm.py uses (:89-111,155,393) must
    exist with the reference's call shapes.  (tests/test_ref_script.py runs the script bodies themselves on a GPU.)"""
    tf = core
    core.reset_default_graph()
    with tf.variable_scope("bidirectional_lstm"):
        core.setup_bidirectional_lstm(4, False, n_embedding_width=8)
    core.setup_core_architecture("nonvis", "first_last_mention", 6, 8, 1, False, "relu", 2, 4)
    loss = tf.get_collection("loss")[0]
    core.add_train_op(loss, 1e-3, 1e-8, 5.0)
    assert tf.get_collection("train_op")[0].kind == "train_op" and tf.get_collection("accuracy")[0].kind == "accuracy"
    assert tf.global_variables_initializer().kind == "init"
    assert isinstance(tf.train.Saver(max_to_keep=100), core.Saver)
    core.dump_tf_vars()
    sess = tf.Session()                       # no arguments, like tf.Session(); the device model is created on first use
    assert sess.handle is None and sess.max_seq_len > 0
    core.reset_default_graph()
