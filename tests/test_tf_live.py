"""Optional pin of the oracle against a REAL TensorFlow.  The reference's arithmetic lives in
TensorFlow 1.x, which is not installable in the build container or on the GPU box (no network), so these tests are skipped there and
have never been executed: they exist so that whoever has TensorFlow can close the "parity unpinned" gap in one command
(`pip install tensorflow && python -m pytest tests/test_tf_live.py`).  They build, with `tf.compat.v1`, exactly the ops
`nn_utils/core.py` calls -- `BasicLSTMCell(H, state_is_tuple=True)` x 2 under `bidirectional_dynamic_rnn(..., sequence_length,
time_major=False, parallel_iterations=64)` (`core.py:305-329`) and `clip_by_global_norm` + `AdamOptimizer` (`core.py:94-103`) --
load the oracle's weights into TensorFlow's own variables BY NAME (the names checkpoints use) and compare."""
import numpy as np
import pytest

tf = pytest.importorskip("tensorflow")
tf1 = tf.compat.v1
if not hasattr(tf1.nn, "rnn_cell") or not hasattr(tf1.nn.rnn_cell, "BasicLSTMCell"):
    pytest.skip("this TensorFlow no longer ships tf.compat.v1.nn.rnn_cell.BasicLSTMCell", allow_module_level=True)

from oracle import icl_oracle as O


def test_bilstm_forward_matches_tensorflow():
    S, T, E, H = 5, 7, 6, 4
    rng = np.random.default_rng(3)
    lens = np.array([7, 3, 1, 5, 0], np.int32)
    x = rng.standard_normal((S, T, E)).astype(np.float32)
    for s in range(S):
        x[s, lens[s]:] = 0.0                                   # load_batch leaves the padding at zero (nn_utils/data.py:375)
    params = {}
    for d in ("fw", "bw"):
        kn, bn = O.lstm_names(d)
        params[kn] = (rng.standard_normal((E + H, 4 * H)) * 0.3).astype(np.float32)
        params[bn] = (rng.standard_normal(4 * H) * 0.1).astype(np.float32)
    want_fw, want_bw, _ = O.bilstm_forward(params, x.astype(np.float64), lens)
    g = tf1.Graph()
    with g.as_default():
        with tf1.variable_scope("bidirectional_lstm"):          # core.py:283
            xp = tf1.placeholder(tf.float32, [S, T, E])
            lp = tf1.placeholder(tf.int32, [S])
            cf = tf1.nn.rnn_cell.BasicLSTMCell(H, state_is_tuple=True)
            cb = tf1.nn.rnn_cell.BasicLSTMCell(H, state_is_tuple=True)
            (ofw, obw), _ = tf1.nn.bidirectional_dynamic_rnn(cf, cb, xp, sequence_length=lp, dtype=tf.float32, time_major=False,
                                                             parallel_iterations=64)
        with tf1.Session(graph=g) as sess:
            sess.run(tf1.global_variables_initializer())
            by_name = {v.name.split(":")[0]: v for v in tf1.global_variables()}
            assert set(params) <= set(by_name), (sorted(params), sorted(by_name))       # TF variable names = our tensor names
            for n, v in params.items():
                by_name[n].load(v, sess)
            got_fw, got_bw = sess.run([ofw, obw], {xp: x, lp: lens})
    assert np.max(np.abs(got_fw - want_fw)) < 2e-6 and np.max(np.abs(got_bw - want_bw)) < 2e-6


def test_clip_and_adam_match_tensorflow():
    rng = np.random.default_rng(4)
    w0 = {"a": rng.standard_normal((3, 4)).astype(np.float32), "b": rng.standard_normal(5).astype(np.float32)}
    tgt = {k: rng.standard_normal(v.shape).astype(np.float32) * 4 for k, v in w0.items()}
    lr, eps, clip = 1e-2, 1e-8, 5.0
    params = {k: v.astype(np.float64).copy() for k, v in w0.items()}
    state = {}
    for _ in range(3):                                          # loss = sum((w - target)^2): gradient norm above the clip
        O.clip_and_adam(params, {k: 2.0 * (params[k] - tgt[k]) for k in params}, state, lr, eps, clip)
    g = tf1.Graph()
    with g.as_default():
        vs = {k: tf1.get_variable(k, initializer=tf.constant(v)) for k, v in w0.items()}
        loss = tf.add_n([tf.reduce_sum(tf.square(vs[k] - tgt[k])) for k in vs])
        opt = tf1.train.AdamOptimizer(lr, epsilon=eps)          # core.py:94-103
        gv = opt.compute_gradients(loss)
        clipped, _ = tf.clip_by_global_norm([gr for gr, _ in gv], clip)
        step = opt.apply_gradients(list(zip(clipped, [v for _, v in gv])))
        with tf1.Session(graph=g) as sess:
            sess.run(tf1.global_variables_initializer())
            for _ in range(3):
                sess.run(step)
            got = {k: sess.run(v) for k, v in vs.items()}
    for k in w0:
        assert np.max(np.abs(got[k] - params[k])) < 1e-5, k
