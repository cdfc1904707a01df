"""CPU: the NumPy oracle of each simulator graph against a second, independently written torch-f64 restatement
(oracle/torch_ref.py).  Not a pin (the deepctr / TF arithmetic has no reference-held vector), but two
independent transcriptions of the published layer definitions must agree to f64 rounding."""
import numpy as np

from rl4rs_b200 import synth
from oracle.dien_np import DienOracle
from oracle.dnn_np import DnnOracle
from oracle import torch_ref

SMALL = {"category_hash_size": 600}


def _rows(R, seed, hs):
    rs = np.random.RandomState(seed)
    seq = np.zeros((R, 2, 64), np.int64)
    for i in range(R):
        for s in range(2):
            n = rs.randint(0, 65)
            if n:
                seq[i, s, 64 - n:] = rs.randint(1, 284, n)
    return seq, rs.normal(0, 2, (R, 432)), rs.randint(0, hs, (R, 21))


def test_dien_numpy_oracle_matches_independent_torch_restatement():
    for kw in ({}, {"stress": 2.0, "bias_noise": 0.1, "bounded_scores": True}):
        w = synth.make_weights(SMALL, **kw)
        seq, dense, cat = _rows(24, 1, 600)
        o_np, p_np = DienOracle(w, np.float64).forward(seq, dense, cat)
        o_t, p_t = torch_ref.dien_forward(w, seq, dense, cat)
        np.testing.assert_allclose(o_np, o_t, rtol=0, atol=1e-11 * max(1.0, np.abs(o_t).max()))
        np.testing.assert_allclose(p_np, p_t, rtol=0, atol=1e-12)
        # and the f32 oracle the GPU tests use sits at f32 rounding of it
        o32, _ = DienOracle(w, np.float32).forward(seq, dense, cat)
        rms = np.sqrt((o_t ** 2).mean(-1, keepdims=True))
        assert (np.abs(o32 - o_t) / np.maximum(np.abs(o_t), rms)).max() < 2e-5


def test_dnn_numpy_oracle_matches_independent_torch_restatement():
    w = synth.make_dnn_weights(SMALL, stress=2.0, bias_noise=0.2)
    _, dense, cat = _rows(64, 2, 600)
    o_np, p_np = DnnOracle(w, np.float64).forward(None, dense, cat)
    o_t, p_t = torch_ref.dnn_forward(w, dense, cat)
    np.testing.assert_allclose(o_np, o_t, rtol=0, atol=1e-12 * max(1.0, np.abs(o_t).max()))
    np.testing.assert_allclose(p_np, p_t, rtol=0, atol=1e-13)
    o32, p32 = DnnOracle(w, np.float32).forward(None, dense, cat)
    assert np.abs(o32 - o_t).max() < 1e-4 * max(1.0, np.abs(o_t).max())


def test_widedeep_numpy_oracle_matches_independent_torch_restatement(tmp_path):
    from oracle.widedeep_np import WideDeepOracle
    from rl4rs_b200.utils import tf_checkpoint as tfc
    w = synth.make_widedeep_weights(SMALL, stress=2.0, bias_noise=0.2)
    seq, dense, cat = _rows(48, 3, 600)
    o_np, p_np = WideDeepOracle(w, np.float64).forward(seq, dense, cat)
    o_t, p_t = torch_ref.widedeep_forward(w, seq, dense, cat)
    assert o_np.shape == (48, 3072)
    np.testing.assert_allclose(o_np, o_t, rtol=0, atol=1e-12 * max(1.0, np.abs(o_t).max()))
    np.testing.assert_allclose(p_np, p_t, rtol=0, atol=1e-13)
    p = tfc.save_widedeep_checkpoint(str(tmp_path / "wd"), w, SMALL)
    got = tfc.load_widedeep_checkpoint(p, SMALL)
    assert set(got) == set(w) and all(np.array_equal(got[k], w[k]) for k in w)
    assert tfc.widedeep_variable_names(SMALL)["fc_w"] == "dense_2/kernel"


def test_dnn_checkpoint_round_trip(tmp_path):
    from rl4rs_b200.utils import tf_checkpoint as tfc
    w = synth.make_dnn_weights(SMALL, bias_noise=0.1)
    names = tfc.dnn_variable_names(SMALL)
    assert names["fc_w"] == "dense_2/kernel" and names["emb_cat"] == "embedding/embeddings" and names["rew_b"] == "simulator_reward/bias"
    extra = dict({names[k]: v for k, v in w.items()}, **{"embedding_1/embeddings": np.zeros((600, 128), np.float32)})
    p = tfc.write_bundle(str(tmp_path / "dnn"), extra)
    got = tfc.load_dnn_checkpoint(p, SMALL)
    assert set(got) == set(w)
    for k in w:
        np.testing.assert_array_equal(got[k], w[k])


def test_lstm_numpy_oracle_matches_independent_torch_restatement(tmp_path):
    from oracle.lstm_np import LstmOracle
    from rl4rs_b200.utils import tf_checkpoint as tfc
    w = synth.make_lstm_weights(SMALL, stress=2.0, bias_noise=0.2)
    seq, dense, cat = _rows(24, 5, 600)
    o_np, p_np = LstmOracle(w, np.float64).forward(seq, dense, cat)
    o_t, p_t = torch_ref.lstm_forward(w, seq, dense, cat)
    assert o_np.shape == (24, 256)
    np.testing.assert_allclose(o_np, o_t, rtol=0, atol=1e-12 * max(1.0, np.abs(o_t).max()))
    np.testing.assert_allclose(p_np, p_t, rtol=0, atol=1e-13)
    o32, _ = LstmOracle(w, np.float32).forward(seq, dense, cat)
    assert np.abs(o32 - o_t).max() < 1e-4 * max(1.0, np.abs(o_t).max())
    p = tfc.save_lstm_checkpoint(str(tmp_path / "lstm"), w, SMALL)
    got = tfc.load_lstm_checkpoint(p, SMALL)
    assert set(got) == set(w) and all(np.array_equal(got[k], w[k]) for k in w)
    names = tfc.lstm_variable_names(SMALL)
    assert names["cgru_rk"] == "gru/recurrent_kernel" and names["sgru1_b"] == "gru_2/bias" and names["emb_seq"] == "embedding_1/embeddings"
