"""The one known answer the reference publishes for the simulator arithmetic: tutorial.ipynb cell 10 -- the released
`simulator_a_dien` checkpoint, SlateRecEnv-v0, logged slate [31, 28, 20, 87, 73, 146, 235, 233, 166] -> reward
130.2745725877583 at step 8 (tutorial.ipynb:169-177).  The checkpoint and the dataset are a Drive download, not part of
the reference tree, so these tests SKIP unless the files are present:

    RL4RS_MATERIALS=/path/with   simulator/finetuned/simulator_a_dien/model.{index,data-00000-of-00001}
                                 simulator/rl4rs_dataset_a_shuf.csv        raw_data/item_info.csv

The day they are, this is the pin SURVEY.md section 8c asks for: TensorFlow's own output on a real record, against the
oracle (CPU) and the CUDA path (GPU), through the TF1-checkpoint reader of rl4rs_b200/utils/tf_checkpoint.py."""
import os

import numpy as np
import pytest

KNOWN_SLATE = [31, 28, 20, 87, 73, 146, 235, 233, 166]
KNOWN_REWARD = 130.2745725877583
CONFIG = {"epoch": 10000, "maxlen": 64, "batch_size": 1, "action_size": 284, "class_num": 2, "dense_feature_num": 432,
          "category_feature_num": 21, "category_hash_size": 100000, "seq_num": 2, "emb_size": 128, "is_eval": True,
          "cache_size": 1, "hidden_units": 128, "max_steps": 9, "action_emb_size": 32, "support_rllib_mask": True}


def materials():
    root = os.environ.get("RL4RS_MATERIALS", "")
    files = {"model_file": os.path.join(root, "simulator", "finetuned", "simulator_a_dien", "model"),
             "sample_file": os.path.join(root, "simulator", "rl4rs_dataset_a_shuf.csv"),
             "iteminfo_file": os.path.join(root, "raw_data", "item_info.csv")}
    ok = root and os.path.exists(files["model_file"] + ".index") and all(os.path.exists(files[k]) for k in ("sample_file", "iteminfo_file"))
    if not ok:
        pytest.skip("published simulator checkpoint / dataset not present (set RL4RS_MATERIALS)")
    return files


def records_with_the_known_slate(sample_file, limit=64):
    """Log records whose exposed items are the tutorial's slate (the logged policy replays them: slate.py:149-162)."""
    from rl4rs_b200.utils.datautil import FeatureUtil
    hits = []
    with open(sample_file) as f:
        for line in f:
            line = line.strip()
            if line and FeatureUtil.record_split(line)[3] == KNOWN_SLATE:
                hits.append(line)
                if len(hits) >= limit:
                    break
    if not hits:
        pytest.skip("no record with the tutorial's slate in this sample file")
    return hits


def test_known_answer_oracle():
    """CPU: NumPy oracle + the checkpoint as read without TensorFlow."""
    files = materials()
    from oracle.dien_np import DienOracle
    from oracle.env_np import OracleEnv
    from rl4rs_b200 import synth
    from rl4rs_b200.utils import tf_checkpoint
    from rl4rs_b200.utils.datautil import FeatureUtil
    w = tf_checkpoint.load_dien_checkpoint(files["model_file"], CONFIG)
    cat = synth.Catalog.from_file(files["iteminfo_file"])
    recs = records_with_the_known_slate(files["sample_file"])
    cfg = dict(CONFIG, batch_size=len(recs), cache_size=len(recs))
    env = OracleEnv(cfg, FeatureUtil.parse_log(recs, 64), cat, DienOracle(w, np.float32))
    env.reset()
    for t in range(9):
        a = env.offline_action
        assert (np.asarray(a) == KNOWN_SLATE[t]).all()
        _, reward, done, _ = env.step(a)
    rel = np.abs(np.asarray(reward) - KNOWN_REWARD) / KNOWN_REWARD
    assert rel.min() < 1e-4, ("no record with that slate lands on the tutorial's reward", sorted(np.asarray(reward))[:5])


@pytest.mark.gpu
def test_known_answer_cuda():
    """GPU: the same episode through the product (reference-facing env API -> C-ABI -> kernels)."""
    files = materials()
    from rl4rs_b200 import gymshim
    from rl4rs_b200.env.slate import SlateRecEnv, SlateState
    from rl4rs_b200.utils.datautil import FeatureUtil
    recs = records_with_the_known_slate(files["sample_file"])
    cfg = dict(CONFIG, batch_size=len(recs), cache_size=len(recs), model_file=files["model_file"],
               iteminfo_file=files["iteminfo_file"], log=FeatureUtil.parse_log(recs, 64), output_format="numpy")
    env = gymshim.make("SlateRecEnv-v0", recsim=SlateRecEnv(cfg, state_cls=SlateState))
    env.reset()
    for t in range(9):
        _, reward, done, _ = env.step(env.offline_action)
    rel = np.abs(np.asarray(reward) - KNOWN_REWARD) / KNOWN_REWARD
    assert rel.min() < 1e-4
