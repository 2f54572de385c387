"""Drop-in task drivers: the train()/predict() loops and command lines of the reference's four LSTM scripts on top of
`core` (hand-written sm_100a kernels) instead of TensorFlow.

    icl_core_lstm.py      --task nonvis|card                 icl_core_lstm.py:21-254,257-406
    icl_relation_lstm.py  --rel_type intra|ordered_intra|cross   icl_relation_lstm.py:21-327,330-495
    icl_affinity_lstm.py                                      icl_affinity_lstm.py:77-343,346-508
    icl_multitask_lstm.py --multitask_scheme simple_joint|weighted_joint|alternate     icl_multitask_lstm.py:28-93,213-476

Same flags, defaults, data_dir/{raw,feats,scores} path scheme, model-file naming, epoch loop (tail batch dropped,
evaluation every 10th epoch, best-average-F1 bookkeeping with the 0.005 slack, early stopping after 10 epochs) and
scores files.  Also accepted: the stale flag names the sweep configs still carry (`--input_keep_prob`,
`--other_keep_prob`, `--pair_enc_scheme`; config/lstm_intra_params.config:27,96-101), and `--embeddings <npz|txt>`
because the word2vec binary + gensim of utils/Word2Vec.py are not available offline.
"""
import argparse
import logging
import os
import sys

import numpy as np

from . import core as nn_util
from . import data as nn_data
from . import eval as nn_eval
from . import loaders

CLASSES_VISUAL = ["v", "n"]
CLASSES_CARD = ["0", "1", "2", "3", "4", "5", "6", "7", "8", "9", "10", "11+"]
CLASSES_REL = ["n", "c", "b", "p"]
CLASSES_AFFINITY = ["0", "1"]


def _log():
    logging.basicConfig(stream=sys.stdout, level=logging.INFO, format="%(asctime)s %(levelname)s %(message)s")
    return logging.getLogger("icl")


def common_args(parser):
    a = parser.add_argument
    a("--epochs", type=int, default=20)
    a("--batch_size", type=int, default=512)
    a("--lstm_hidden_width", type=int, default=200)
    a("--start_hidden_width", type=int, default=512)
    a("--hidden_depth", type=int, default=2)
    a("--weighted_classes", action="store_true")
    a("--learn_rate", type=float, default=0.001)
    a("--adam_epsilon", type=float, default=1e-08)
    a("--clip_norm", type=float, default=5.0)
    a("--data_norm", action="store_true")
    a("--lstm_input_dropout", "--input_keep_prob", dest="lstm_input_dropout", type=float, default=0.5)
    a("--dropout", "--other_keep_prob", dest="dropout", type=float, default=0.5)
    a("--data_dir", required=True)
    a("--train", action="store_true")
    a("--predict", action="store_true")
    a("--activation", choices=["sigmoid", "tanh", "relu", "leaky_relu"], default="relu")
    a("--model_file", type=str)
    a("--model_dir", type=str, default=None, help="where --train puts the auto-named model (reference: a lab path)")
    a("--embedding_type", choices=["w2v", "glove"], default="w2v")
    a("--embeddings", type=str, default=None, help=".npz (vocab, matrix) or GloVe-style text file")
    a("--early_stopping", action="store_true")
    a("--skip_epoch_eval", action="store_true")
    a("--eval_every", type=int, default=10, help="evaluate every N epochs (reference: 10)")
    a("--encoding_scheme", "--pair_enc_scheme", dest="encoding_scheme",
      choices=["first_last_sentence", "first_last_mention"], default="first_last_mention")


def load_embeddings(args, data_dir, data_root):
    path = args.embeddings or os.path.join(data_dir, "raw", data_root + "_embeddings.npz")
    if path.endswith(".npz"):
        return loaders.Embeddings.from_npz(path)
    if path.endswith(".bin"):
        return loaders.Embeddings.from_word2vec_bin(path)
    return loaders.Embeddings.from_text(path)


def build_graph(task, args, n_classes, n_feats, box_w=None, n_box_feats=None):
    nn_util.reset_default_graph()
    nn_util.set_random_seeds()
    with nn_util.variable_scope("bidirectional_lstm"):
        nn_util.setup_bidirectional_lstm(args.lstm_hidden_width, args.data_norm)
    nn_util.setup_core_architecture(task, args.encoding_scheme, args.batch_size, args.start_hidden_width, args.hidden_depth,
                                    args.weighted_classes, args.activation, n_classes, n_feats, box_w, n_box_feats)
    nn_util.add_train_op(nn_util.get_collection("loss")[0], args.learn_rate, args.adam_epsilon, args.clip_norm)


def predict_labels(task, args, sess, ids, data_dict, n_classes, log):
    pred_scores, gold = nn_util.get_pred_scores_mcc(task, args.encoding_scheme, sess, args.batch_size, ids, data_dict, n_classes)
    keys = list(pred_scores.keys())
    return pred_scores, keys, [int(np.argmax(pred_scores[k])) for k in keys], [int(np.argmax(gold[k])) for k in keys]


def train_loop(task, args, data_dict, eval_dict, classes, ids, log, shuffle=None, eval_fn=None, model_file=None):
    """The epoch loop shared by the scripts (icl_core_lstm.py:113-197)."""
    n_classes = len(classes)
    loss_op, acc_op, train_op = (nn_util.get_collection(k)[0] for k in ("loss", "accuracy", "train_op"))
    saver = nn_util.Saver(max_to_keep=100)
    best_avg, best_epoch = -1, -1
    T = max(data_dict["max_seq_len"], eval_dict["max_seq_len"] if eval_dict else 0)
    with nn_util.Session(max_seq_len=T) as sess:
        sess.ensure()
        ids = list(ids)
        for i in range(args.epochs):
            log.info("--- Epoch %d ----", i + 1)
            losses, accs = [], []
            ids = shuffle(ids) if shuffle else list(np.random.permutation(np.asarray(ids, dtype=object)))
            n_iter = len(ids) // args.batch_size                     # the tail batch is dropped (py2 integer division)
            for j in range(n_iter):
                bt = nn_data.load_batch(ids[j * args.batch_size:(j + 1) * args.batch_size], data_dict, task, n_classes,
                                        packed=nn_data.default_packing())
                nn_util.run_op(sess, train_op, [bt], args.lstm_input_dropout, args.dropout, args.encoding_scheme, [task], [""], True)
                if (j + 1) % 100 == 0 or j == n_iter - 1:
                    losses.append(nn_util.run_op(sess, loss_op, [bt], args.lstm_input_dropout, args.dropout,
                                                 args.encoding_scheme, [task], [""], True))
                    accs.append(nn_util.run_op(sess, acc_op, [bt], args.lstm_input_dropout, args.dropout,
                                               args.encoding_scheme, [task], [""], True))
            if losses:
                log.info("Saving model; Average Loss: %.2f; Acc: %.2f%%", sum(losses) / len(losses), 100.0 * sum(accs) / len(accs))
            if model_file:
                saver.save(sess, model_file)
            if (i + 1) % args.eval_every == 0 and eval_dict is not None and not args.skip_epoch_eval:
                avg = eval_fn(sess, eval_dict)
                if avg >= best_avg - 0.005:
                    log.info("Previous best score average F1 of %.2f%% after %d epochs", 100.0 * best_avg, best_epoch)
                    best_avg, best_epoch = avg, i
                    log.info("New best at current epoch (%.2f%%)", 100.0 * best_avg)
                if args.early_stopping and i >= best_epoch + 10:
                    log.info("Stopping early; best scores at %d epochs", best_epoch)
                    break
        if model_file:
            log.info("Saving final model")
            saver.save(sess, model_file)
    return best_avg


def _model_file(args, arg_dict, suffix):
    if args.train and not args.model_file:
        d = args.model_dir or os.path.join(args.data_dir, "models")
        os.makedirs(d, exist_ok=True)
        return os.path.join(d, nn_data.build_model_filename(arg_dict, suffix))
    return args.model_file


# ------------------------------------------------------------------------------------------------ icl_core_lstm.py
def main_core(argv=None):
    log = _log()
    p = argparse.ArgumentParser("ImageCaptionLearn_py: Core Neural Network classification architecture; "
                                "used for nonvis and cardinality prediction")
    common_args(p)
    p.add_argument("--data_root", type=str, required=True)
    p.add_argument("--eval_data_root", type=str)
    p.add_argument("--task", required=True, choices=["nonvis", "card"])
    args = p.parse_args(argv)
    arg_dict = vars(args)
    task = args.task
    classes = CLASSES_VISUAL if task == "nonvis" else CLASSES_CARD
    model_file = _model_file(args, arg_dict, task + "_lstm")

    def files(root):
        d = args.data_dir + "/"
        return (d + "raw/" + root + "_captions.txt", d + "raw/" + root + "_mentions_" + task + ".txt",
                d + "feats/" + root + "_" + task + "_neural.feats", d + "feats/" + root + "_" + task + "_neural_meta.json")

    def load(root):
        emb = load_embeddings(args, args.data_dir, root)
        s, mi, ff, fm = files(root)
        dd = loaders.load_sentences(s, emb)
        dd.update(loaders.load_mentions(mi, task, ff, fm, len(classes)))
        return dd

    data_dict = load(args.data_root)
    if args.train:
        eval_dict = load(args.eval_data_root) if args.eval_data_root else None
        build_graph(task, args, len(classes), data_dict["n_mention_feats"])

        def eval_fn(sess, ed):
            _, _, pred, gold = predict_labels(task, args, sess, list(ed["mention_indices"].keys()), ed, len(classes), log)
            sd = nn_eval.evaluate_multiclass(gold, pred, classes, log)
            return (sd.get_score(0).f1 + sd.get_score(1).f1) / 2.0
        return train_loop(task, args, data_dict, eval_dict, classes, data_dict["mention_indices"].keys(), log, None, eval_fn, model_file)
    if args.predict:
        build_graph(task, args, len(classes), data_dict["n_mention_feats"])
        with nn_util.Session(max_seq_len=data_dict["max_seq_len"]) as sess:
            nn_util.Saver().restore(sess, model_file)
            scores, keys, pred, gold = predict_labels(task, args, sess, list(data_dict["mention_indices"].keys()), data_dict,
                                                      len(classes), log)
            nn_eval.evaluate_multiclass(gold, pred, classes, log)
            out = args.data_dir + "/scores/" + args.data_root + "_" + task + ".scores"
            nn_eval.write_scores_file(out, scores)
            log.info("Wrote scores file %s", out)
            return out


# -------------------------------------------------------------------------------------------- icl_relation_lstm.py
def main_relation(argv=None):
    log = _log()
    p = argparse.ArgumentParser("ImageCaptionLearn_py: Neural Network for Relation Prediction")
    common_args(p)
    p.add_argument("--data_root", type=str, required=True)
    p.add_argument("--eval_data_root", type=str)
    p.add_argument("--rel_type", choices=["intra", "ordered_intra", "cross"], required=True)
    args = p.parse_args(argv)
    arg_dict = vars(args)
    ordered = args.rel_type == "ordered_intra"
    rel = "intra" if ordered else args.rel_type
    task = "rel_" + rel
    model_file = _model_file(args, arg_dict, "relation_lstm")

    def load(root):
        d = args.data_dir + "/"
        emb = load_embeddings(args, args.data_dir, root)
        midx = d + "raw/" + root + "_mentionPairs_" + rel + ("_ij" if ordered else "") + ".txt"
        froot = d + "feats/" + root + "_relation_neural" + ("_intra_ij" if ordered else "_" + rel)
        dd = loaders.load_sentences(d + "raw/" + root + "_captions.txt", emb)
        dd.update(loaders.load_mentions(midx, task, froot + ".feats", froot + "_meta.json", len(CLASSES_REL)))
        lab = d + "raw/" + root + "_mentionPair_labels.txt"
        dd["_gold"] = loaders.load_relation_labels(lab) if os.path.exists(lab) else {}
        return dd

    def scores_for(sess, dd):
        ids = list(dd["mention_indices"].keys())
        scores, _ = nn_util.get_pred_scores_mcc(task, args.encoding_scheme, sess, args.batch_size, ids, dd, len(CLASSES_REL))
        if ordered:
            scores = nn_data.induce_ji_predictions(scores)              # icl_relation_lstm.py:225-245
        return scores

    def eval_fn(sess, ed):
        scores = scores_for(sess, ed)
        keys = list(scores.keys())
        sd = evaluate_relations(keys, [int(np.argmax(scores[k])) for k in keys], ed["_gold"], log)
        return (sd.get_score("coref").f1 + sd.get_score("subset").f1) / 2.0     # icl_relation_lstm.py:188-190

    data_dict = load(args.data_root)
    build_graph(task, args, len(CLASSES_REL), data_dict["n_mention_feats"])
    if args.train:
        eval_dict = load(args.eval_data_root) if args.eval_data_root else None
        return train_loop(task, args, data_dict, eval_dict, CLASSES_REL, data_dict["mention_indices"].keys(), log, None, eval_fn,
                          model_file)
    if args.predict:
        with nn_util.Session(max_seq_len=data_dict["max_seq_len"]) as sess:
            nn_util.Saver().restore(sess, model_file)
            scores = scores_for(sess, data_dict)
            out = args.data_dir + "/scores/" + args.data_root + "_relation_" + rel + ".scores"
            nn_eval.write_scores_file(out, scores)
            return out


evaluate_relations = nn_eval.evaluate_relations          # nn_utils/eval.py:10-93


# -------------------------------------------------------------------------------------------- icl_affinity_lstm.py
def main_affinity(argv=None):
    log = _log()
    p = argparse.ArgumentParser("ImageCaptionLearn_py: Neural Network for Affinity Prediction")
    common_args(p)
    p.add_argument("--data", default="flickr30k")
    p.add_argument("--split", default="train")
    p.add_argument("--eval_data", default="flickr30k")
    p.add_argument("--eval_split", default="dev")
    args = p.parse_args(argv)
    arg_dict = vars(args)
    task = "affinity"
    model_file = _model_file(args, arg_dict, "affinity_lstm")

    def load(data, split):
        d, root = args.data_dir + "/", data + "_" + split
        emb = load_embeddings(args, args.data_dir, root)
        dd = loaders.load_sentences(d + "raw/" + root + "_captions.txt", emb)
        dd.update(loaders.load_mentions(d + "raw/" + root + "_mentions_affinity.txt", task, d + "feats/" + root + "_affinity_neural.feats",
                                        d + "feats/" + root + "_affinity_neural_meta.json", 2))
        dd.update(loaders.load_boxes(d + "raw/" + root + "_affinity_labels.txt", d + "feats/" + data + "_boxes/" + split + "/"))
        if os.environ.get("ICL_BOX_TABLE"):
            loaders.load_all_boxes(dd)                    # box features resident on the device instead of re-read per batch
        return dd, root

    def eval_fn(sess, ed):
        ids = nn_data.get_valid_mention_box_pairs(ed)
        _, _, pred, gold = predict_labels(task, args, sess, ids, ed, 2, log)
        sd = nn_eval.evaluate_multiclass(gold, pred, CLASSES_AFFINITY, log)
        return (sd.get_score(0).f1 + sd.get_score(1).f1) / 2.0

    data_dict, root = load(args.data, args.split)
    build_graph(task, args, 2, data_dict["n_mention_feats"], data_dict["box_embedding_width"], data_dict["n_box_feats"])
    ids = nn_data.get_valid_mention_box_pairs(data_dict)
    if args.train:
        eval_dict = load(args.eval_data, args.eval_split)[0] if args.eval_data else None
        return train_loop(task, args, data_dict, eval_dict, CLASSES_AFFINITY, ids, log, nn_data.shuffle_mention_box_pairs, eval_fn,
                          model_file)
    if args.predict:
        with nn_util.Session(max_seq_len=data_dict["max_seq_len"]) as sess:
            nn_util.Saver().restore(sess, model_file)
            scores, keys, pred, gold = predict_labels(task, args, sess, ids, data_dict, 2, log)
            nn_eval.evaluate_multiclass(gold, pred, CLASSES_AFFINITY, log)
            out = args.data_dir + "/scores/" + root + "_affinity.scores"
            nn_eval.write_scores_file(out, scores)
            return out


# ------------------------------------------------------------------------------------------- icl_multitask_lstm.py
MT_TASKS = ["rel_intra", "rel_cross", "nonvis", "affinity", "card"]                  # icl_multitask_lstm.py:22
MT_CLASSES = {"rel_intra": CLASSES_REL, "rel_cross": CLASSES_REL, "nonvis": CLASSES_VISUAL, "affinity": CLASSES_AFFINITY,
              "card": CLASSES_CARD}


def mt_load_data(args, data, split, label_file=None):
    """icl_multitask_lstm.py:149-209 (file naming of the joint data set; one shared `_relation.feats` for both pair tasks)."""
    d, root = args.data_dir + "/", data + "_" + split
    emb = load_embeddings(args, args.data_dir, root)
    out = {}
    sent = loaders.load_sentences(d + "raw/" + root + "_captions.txt", emb)       # one caption file for all five tasks
    nn_data._flat_table(sent)                                                    # ... and one token table (resident on the device)
    for task in MT_TASKS:
        if task.startswith("rel"):
            midx = d + "raw/" + root + "_mentionPairs_" + task.split("_")[1] + ".txt"
            froot = d + "feats/" + root + "_relation"
        else:
            midx = d + "raw/" + root + "_mentions_" + task + ".txt"
            froot = d + "feats/" + root + "_" + task
        dd = dict(sent)
        dd.update(loaders.load_mentions(midx, task, froot + ".feats", froot + "_meta.json", len(MT_CLASSES[task])))
        if task.startswith("rel"):
            lab = d + "raw/" + root + "_mentionPair_labels.txt"
            dd["gold_label_dict"] = loaders.load_relation_labels(lab) if os.path.exists(lab) else {}
        elif task == "affinity":
            dd.update(loaders.load_boxes(label_file or d + "raw/" + root + "_mention_box_labels.txt",
                                         d + "feats/" + data + "_boxes/" + split + "/"))
        out[task] = dd
    return out


def mt_task_ids(dicts):
    return {t: (nn_data.get_valid_mention_box_pairs(dicts[t]) if t == "affinity" else list(dicts[t]["mention_indices"].keys()))
            for t in MT_TASKS}


def mt_setup(args, dicts, batch_sizes):
    """icl_multitask_lstm.py:28-93: one shared BiLSTM scope, one head per task under a variable_scope named like the task."""
    nn_util.reset_default_graph()
    nn_util.set_random_seeds()
    with nn_util.variable_scope("bidirectional_lstm"):
        nn_util.setup_bidirectional_lstm(args.lstm_hidden_width, args.data_norm)
    for task in MT_TASKS:
        dd = dicts[task]
        with nn_util.variable_scope(task):
            nn_util.setup_core_architecture(task, args.encoding_scheme, batch_sizes[task], args.start_hidden_width, args.hidden_depth,
                                            args.weighted_classes, args.activation, len(MT_CLASSES[task]), dd["n_mention_feats"],
                                            dd.get("box_embedding_width"), dd.get("n_box_feats"))


def mt_evaluate(sess, args, dicts, ids, batch_sizes, log, write_scores=False):
    """icl_multitask_lstm.py:327-358,479-535: per-task prediction through the shared encoder + the task's own head."""
    f1 = {}
    for task in MT_TASKS:
        with nn_util.variable_scope(task):
            scores, gold = nn_util.get_pred_scores_mcc(task, args.encoding_scheme, sess, batch_sizes[task], ids[task], dicts[task],
                                                       len(MT_CLASSES[task]))
        keys = list(scores.keys())
        pred = [int(np.argmax(scores[k])) for k in keys]
        if task.startswith("rel"):
            sd = evaluate_relations(keys, pred, dicts[task]["gold_label_dict"], log)
            f1[task] = (sd.get_score("coref").f1 + sd.get_score("subset").f1) / 2.0
        else:
            sd = nn_eval.evaluate_multiclass([int(np.argmax(gold[k])) for k in keys], pred, MT_CLASSES[task], log)
            f1[task] = float(np.mean([sd.get_score(i).f1 for i in range(len(MT_CLASSES[task]))]))
        if write_scores:
            nn_eval.write_scores_file(dicts[task]["scores_file"], scores)
            log.info("Wrote scores file %s", dicts[task]["scores_file"])
    return f1


def main_multitask(argv=None):
    """icl_multitask_lstm.py:538-760.  The reference's as-written defects (only the last task's sentences
    reach the encoder; the predict call passes its arguments in the wrong positions) are not reproduced: this implements the
    intended semantics -- one shared-weight encoder pass over the concatenation of every task's sentence batch."""
    log = _log()
    p = argparse.ArgumentParser("ImageCaptionLearn_py: Neural Network for multitask learning; shared bidirectional LSTM to hidden "
                                "layers to softmax over labels")
    common_args(p)
    p.add_argument("--data", required=True)
    p.add_argument("--split", required=True)
    p.add_argument("--eval_data", required=True)
    p.add_argument("--eval_split", required=True)
    p.add_argument("--multitask_scheme", choices=["simple_joint", "weighted_joint", "alternate"], default="simple_joint")
    p.add_argument("--mention_box_label_file", type=str)
    p.add_argument("--eval_mention_box_label_file", type=str)
    p.add_argument("--box_category_file", type=str)
    p.add_argument("--eval_box_category_file", type=str)
    args = p.parse_args(argv)
    arg_dict = vars(args)
    scheme = args.multitask_scheme
    model_file = _model_file(args, arg_dict, "multitask_" + scheme + "_lstm")
    dicts = mt_load_data(args, args.data, args.split, args.mention_box_label_file)
    ids = mt_task_ids(dicts)
    batch_sizes = {t: (512 if scheme == "alternate" and args.batch_size == 512 else args.batch_size) for t in MT_TASKS}
    T = max(dd["max_seq_len"] for dd in dicts.values())
    if args.train:
        ev = mt_load_data(args, args.eval_data, args.eval_split, args.eval_mention_box_label_file)
        ev_ids = mt_task_ids(ev)
        T = max([T] + [dd["max_seq_len"] for dd in ev.values()])
        mt_setup(args, dicts, batch_sizes)
        saver = nn_util.Saver(max_to_keep=100)
        B = args.batch_size
        if scheme == "alternate":                                       # one optimizer per task (icl_multitask_lstm.py:387-393)
            train_ops = {}
            for task in MT_TASKS:
                with nn_util.variable_scope(task):
                    nn_util.add_train_op(nn_util.get_collection(task + "/loss")[0], args.learn_rate, args.adam_epsilon, args.clip_norm)
                train_ops[task] = nn_util.get_collection(task + "/train_op")[0]
        else:
            nn_util.add_train_op(nn_util.setup_joint_loss(scheme), args.learn_rate, args.adam_epsilon, args.clip_norm)
            train_op = nn_util.get_collection("train_op")[0]
        f1 = {}
        with nn_util.Session(max_seq_len=T) as sess:
            sess.ensure()
            for i in range(args.epochs):
                log.info("--- Epoch %d ----", i + 1)
                if scheme == "alternate":                               # icl_multitask_lstm.py:405-437
                    batches = []
                    for task in MT_TASKS:
                        t_ids = nn_data.shuffle_mention_box_pairs(ids[task]) if task == "affinity" else list(ids[task])
                        bs = batch_sizes[task]
                        pad = bs * (len(t_ids) // bs + 1) - len(t_ids)
                        mat = np.pad(np.asarray(t_ids, dtype=object), (0, pad), "wrap").reshape(-1, bs)
                        batches.extend((task, list(row)) for row in mat)
                    order = np.random.permutation(len(batches))
                    for j in order:
                        task, row = batches[j]
                        bt = nn_data.load_batch(row, dicts[task], task, len(MT_CLASSES[task]), packed=nn_data.default_packing())
                        nn_util.run_op(sess, train_ops[task], [bt], args.lstm_input_dropout, args.dropout, args.encoding_scheme,
                                       [task], [task], True)
                else:                                                   # icl_multitask_lstm.py:268-323
                    pos, max_samples = {}, 0
                    for task in MT_TASKS:
                        ids[task] = (nn_data.shuffle_mention_box_pairs(ids[task]) if task == "affinity"
                                     else list(np.random.permutation(np.asarray(ids[task], dtype=object))))
                        max_samples = max(max_samples, len(ids[task]))
                        pos[task] = 0
                    for j in range(max_samples // B):
                        bts = []
                        for task in MT_TASKS:
                            t_ids, n, st = ids[task], len(ids[task]), pos[task]
                            if st + B < n:
                                row = t_ids[st:st + B]
                                pos[task] += B
                            else:                                       # wrap-around as written (skips the last id of the list)
                                rem = st + B - n + 1
                                row = list(t_ids[st:n - 1]) + list(t_ids[0:rem])
                                pos[task] = rem
                            bts.append(nn_data.load_batch(row, dicts[task], task, len(MT_CLASSES[task]), packed=nn_data.default_packing()))
                        nn_util.run_op(sess, train_op, bts, args.lstm_input_dropout, args.dropout, args.encoding_scheme, MT_TASKS,
                                       MT_TASKS, True)
                log.info("Saving model")
                saver.save(sess, model_file)
                if not args.skip_epoch_eval and (i + 1) % max(1, args.eval_every if args.eval_every != 10 else 1) == 0:
                    f1 = mt_evaluate(sess, args, ev, ev_ids, batch_sizes, log)
            log.info("Saving final model")
            saver.save(sess, model_file)
        return f1
    if args.predict:
        mt_setup(args, dicts, batch_sizes)
        if scheme != "alternate":
            nn_util.setup_joint_loss(scheme)
        for task in MT_TASKS:
            dicts[task]["scores_file"] = (args.data_dir + "/scores/" + args.data + "_" + args.split + "_" + task + "_mulit_" + scheme +
                                          "_lstm.scores")                # sic (icl_multitask_lstm.py:741-744)
        with nn_util.Session(max_seq_len=T) as sess:
            nn_util.Saver().restore(sess, model_file)
            mt_evaluate(sess, args, dicts, ids, batch_sizes, log, write_scores=True)
        return [dicts[t]["scores_file"] for t in MT_TASKS]
