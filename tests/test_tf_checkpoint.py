"""CPU: TF1 Saver checkpoint (tensor bundle) reader / writer and the DIEN variable-name map (SURVEY.md 8f n1;
rl4rs/env/base.py:119-131,148-151, README.md:124-137)."""
import os
import struct

import numpy as np
import pytest

from rl4rs_b200 import synth
from rl4rs_b200.utils import tf_checkpoint as tfc

SMALL = {"category_hash_size": 300}


def test_crc32c_known_answers():
    assert tfc.crc32c(b"123456789") == 0xE3069283                     # the CRC-32C check value
    assert tfc.crc32c(b"") == 0
    assert tfc.mask_crc(tfc.crc32c(b"\x00" * 32)) == ((((0x8A9136AA >> 15) | (0x8A9136AA << 17)) + 0xa282ead8) & 0xFFFFFFFF)


def test_bundle_round_trip(tmp_path):
    rs = np.random.RandomState(0)
    tensors = {"a/kernel": rs.normal(size=(7, 5)).astype(np.float32), "a/bias": np.arange(5, dtype=np.float32),
               "global_step": np.array(1234, dtype=np.int64), "z/w": rs.normal(size=(3, 2, 4)),
               "ids": rs.randint(0, 9, (11,)).astype(np.int32)}
    for i in range(40):                                               # several data blocks + a multi-entry index block
        tensors["layer_%02d/kernel" % i] = rs.normal(size=(i + 1, 3)).astype(np.float32)
    prefix = str(tmp_path / "model")
    tfc.write_bundle(prefix, tensors)
    rd = tfc.TensorBundleReader(prefix)
    assert set(rd.variables()) == set(tensors)
    for k, v in tensors.items():
        got = rd.get_tensor(k, verify=True)
        assert got.dtype == v.dtype and got.shape == v.shape
        np.testing.assert_array_equal(got, v)
    assert rd.variables()["z/w"] == (np.float64, (3, 2, 4))
    # on-disk format facts a TF reader relies on: footer magic, data file = concatenated raw bytes in key order
    idx = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", idx[-8:])[0] == tfc.TABLE_MAGIC
    raw = open(prefix + ".data-00000-of-00001", "rb").read()
    assert len(raw) == sum(v.nbytes for v in tensors.values())
    first = sorted(tensors, key=lambda s: s.encode())[0]
    assert raw[:tensors[first].nbytes] == tensors[first].tobytes()


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "m")
    tfc.write_bundle(prefix, {"w": np.arange(12, dtype=np.float32).reshape(3, 4)})
    raw = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    raw[5] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        tfc.TensorBundleReader(prefix).get_tensor("w")
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[3] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError):
        tfc.TensorBundleReader(prefix)
    with pytest.raises(FileNotFoundError):
        tfc.TensorBundleReader(str(tmp_path / "nothing"))


def test_snappy_blocks_are_decoded():
    # literal "abcdefgh" + copy(offset 8, length 8) + literal "xyz"
    comp = bytes([19]) + bytes([(8 - 1) << 2]) + b"abcdefgh" + bytes([((8 - 1) << 2) | 2, 8, 0]) + bytes([(3 - 1) << 2]) + b"xyz"
    assert tfc._snappy_uncompress(comp) == b"abcdefghabcdefghxyz"


def test_dien_name_map_round_trip_and_resolution(tmp_path):
    """W-table -> Saver checkpoint under the TF1 names of the DIEN graph -> W-table; then the same checkpoint with
    (a) optimizer slots and metric counters added, (b) the deepctr-internal names changed: the resolver keys on the
    top-level layer scopes + shapes, so both still load; an ambiguous / missing variable raises with candidates."""
    w = synth.make_weights(SMALL)
    names = tfc.dien_variable_names(SMALL)
    assert names["emb_cat"] == "embedding/embeddings" and names["emb_seq"] == "embedding_1/embeddings"
    assert names["gru1_wc"] == "dynamic_gru_2/gru_cell/candidate/kernel"
    assert names["augru0_bg"] == "dynamic_gru_1/vec_att_gru_cell/gates/bias"
    assert names["att1_w2"] == "attention_sequence_pooling_layer_1/local_activation_unit/dnn/kernel1"
    assert names["obs_w"] == "simulator_obs/kernel" and names["dense_b2"] == "dense_1/bias"
    assert len(set(names.values())) == len(w) == 38
    p1 = tfc.save_dien_checkpoint(str(tmp_path / "sim"), w, SMALL)
    got = tfc.load_dien_checkpoint(p1, SMALL)
    assert set(got) == set(w)
    for k in w:
        np.testing.assert_array_equal(got[k], w[k])
    # (a) + (b)
    t2 = {}
    for k, nm in names.items():
        nm2 = nm.replace("local_activation_unit/dnn/", "lau/deep/").replace("gru_cell/", "cell/")
        t2[nm2] = w[k]
        t2[nm2 + "/Adam"] = np.zeros_like(w[k])
        t2[nm2 + "/Adam_1"] = np.zeros_like(w[k])
    t2["beta1_power"] = np.float32(0.9)
    t2["training/Adam/iter"] = np.int64(7)
    p2 = tfc.write_bundle(str(tmp_path / "sim2"), t2)
    got2 = tfc.load_dien_checkpoint(p2, SMALL)
    for k in w:
        np.testing.assert_array_equal(got2[k], w[k])
    # explicit override wins; a wrong one is reported
    got3 = tfc.load_dien_checkpoint(p2, SMALL, name_map={"rew_b": "simulator_reward/bias"})
    np.testing.assert_array_equal(got3["rew_b"], w["rew_b"])
    with pytest.raises(KeyError):
        tfc.load_dien_checkpoint(p2, SMALL, name_map={"rew_b": "nope"})
    del t2[names["obs_b"]]
    p3 = tfc.write_bundle(str(tmp_path / "sim3"), t2)
    with pytest.raises(KeyError, match="obs_b"):
        tfc.load_dien_checkpoint(p3, SMALL)


def test_model_file_accepts_saver_prefix(tmp_path):
    """config['model_file'] = Saver prefix goes through get_model exactly like the reference's restore (base.py:148-151)."""
    from rl4rs_b200.env.slate import SlateRecEnv
    w = synth.make_weights(SMALL)
    prefix = tfc.save_dien_checkpoint(str(tmp_path / "simulator_a_dien" / "model"), w, SMALL)
    cfg = dict(SMALL, model_file=prefix, batch_size=4, max_steps=9)
    got = SlateRecEnv.get_model(SlateRecEnv.__new__(SlateRecEnv), cfg)
    for k in w:
        np.testing.assert_array_equal(got[k], w[k])
