"""Generates tests/golden/nets/reference_graph.npz: what the REFERENCE'S OWN graph code (rl4rs/nets/{dien,dnn,widedeep,
lstm}.py + nets/utils.py, run from /root/reference through oracle/tf_eager_stub.py) computes on seeded feature rows
with the seeded synthetic weights, plus the variable scopes in the order that code creates them.

    python tests/golden/make_nets_golden.py          # needs /root/reference (build container only)

The layer arithmetic inside the stub is a restatement (TF 1.15 / deepctr 0.9.0 are absent): see the stub's header for
what these vectors pin (the wiring and the variable-scope order) and what they do not (the third-party arithmetic).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from rl4rs_b200 import synth                                   # noqa: E402
from rl4rs_b200.utils import tf_checkpoint as tfc              # noqa: E402
from oracle import tf_eager_stub as stub                       # noqa: E402

SMALL = {"category_hash_size": 600}
NETS_CONFIG = dict(maxlen=64, dense_feature_num=432, category_feature_num=21, class_num=2, seq_num=2, emb_size=128,
                   hidden_units=128, category_hash_size=600, batch_size=8)
CASES = {   # name -> (config['algo'], weight maker, its keyword arguments, W-table -> TF1 variable names)
    "dien_default": ("dien", "make_weights", {}, "dien_variable_names"),
    "dien_stress": ("dien", "make_weights", {"stress": 2.0, "bias_noise": 0.1, "bounded_scores": True}, "dien_variable_names"),
    "dnn": ("dnn", "make_dnn_weights", {"stress": 2.0, "bias_noise": 0.2}, "dnn_variable_names"),
    "widedeep": ("widedeep", "make_widedeep_weights", {"stress": 2.0, "bias_noise": 0.2}, "widedeep_variable_names"),
    "lstm": ("lstm", "make_lstm_weights", {"stress": 2.0, "bias_noise": 0.2}, "lstm_variable_names"),
}


def feature_rows(R=8, seed=11, hash_size=600):
    rs = np.random.RandomState(seed)
    seq = np.zeros((R, 2, 64), np.int64)
    for i in range(R):
        for s in range(2):
            n = rs.randint(0, 65) if i else 64 * s              # row 0: one empty and one full history
            if n:
                seq[i, s, 64 - n:] = rs.randint(1, 284, n)
    return seq, rs.normal(0, 2, (R, 432)), rs.randint(0, hash_size, (R, 21))


def checkpoint_of(case):
    """The {TF1 variable name: array} dict a Saver checkpoint of this simulator holds (the writer side of n1)."""
    algo, maker, kw, names = CASES[case]
    w = getattr(synth, maker)(SMALL, **kw)
    nm = getattr(tfc, names)(SMALL)
    ck = {nm[k]: np.asarray(v) for k, v in w.items() if k in nm}
    if algo == "dnn":                                           # nets/dnn.py builds a sequence embedding that feeds nothing
        ck["embedding_1/embeddings"] = np.zeros((600, 128), np.float32)
    return w, ck


def main():
    seq, dense, cat = feature_rows()
    out = {"seq": seq, "dense": dense, "cat": cat}
    meta = {}
    for case, (algo, _, _, _) in CASES.items():
        _, ck = checkpoint_of(case)
        r = stub.run_reference_graph(algo, NETS_CONFIG, ck, seq.astype(np.float32), dense, cat)
        assert not r["unused"], r["unused"]
        out[case + "_obs"], out[case + "_probs"] = r["obs"], r["probs"]
        meta[case] = {"algo": algo, "variables": [[s, n, list(sh)] for s, n, sh in r["variables"]], "layers": r["layers"]}
        print(case, r["obs"].shape, "variables", len(r["variables"]))
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "nets", "reference_graph.npz"), **out)


if __name__ == "__main__":
    main()
