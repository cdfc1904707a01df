import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from test_gpu_parity import _synthetic
from test_gpu_dataset import _oracle_dataset
from rl4rs_b200 import dataset
B, epochs = 16, 3
cfg, cat, log, w = _synthetic(B, True)
np.random.seed(11); got = dataset.data_generate_rl4rs_b(dict(cfg, catalog=cat, log=log, weights=w), None, epochs=epochs)
np.random.seed(11); O, A, R, D = _oracle_dataset(cfg, True, False, log, cat, w, epochs)
T = cfg["max_steps"]
g = got["observations"].reshape(epochs, B, T + 1, 266); o = O.reshape(epochs, B, T + 1, 266)
err = np.abs(g[..., :256] - o[..., :256]).max(-1)      # [ep, B, T+1]
print("max err per entry index:", np.round(err.max((0, 1)), 4))
print("max err per epoch:", err.max((1, 2)))
print("tail cols equal:", np.array_equal(g[..., 256:], o[..., 256:]), "actions equal:", np.array_equal(got["actions"], A))
