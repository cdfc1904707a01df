"""GPU: the CUDA simulator forward against the vectors computed by the REFERENCE'S OWN graph code (rl4rs/nets/dien.py +
nets/utils.py run over the layer stand-ins of oracle/tf_eager_stub.py; tests/golden/nets/reference_graph.npz).

The chain CUDA <-> f32 oracle (tests/test_gpu_parity*.py, 1e-4) and f64 oracle <-> these vectors (tests/test_reference_graph.py,
1e-11) already ties the two together; this is the direct comparison.  It was written after the round's GPU budget was
spent and has never run on a GPU, so it is marked xfail(strict=False): it reports XPASS when the kernels land inside the
parity bound on these rows and cannot turn an otherwise green suite red if these eight rows (one with a full 64-step
history under the raw, unbounded attention scores) sit closer to the bound than the 0.46 the measured regimes showed.
The file name sorts it behind every other GPU test."""
import json
import os
import sys

import numpy as np
import pytest

from golden_util import assert_close_rel

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


@pytest.mark.xfail(strict=False, reason="first execution happens on the driver's GPU box; see the module docstring")
@pytest.mark.parametrize("case", ["dien_stress", "dien_default"])
def test_cuda_forward_matches_reference_graph_vectors(case):
    import make_nets_golden as mk
    from test_gpu_parity import _synthetic, make_env
    g = np.load(os.path.join(os.path.dirname(mk.__file__), "nets", "reference_graph.npz"))
    w, _ = mk.checkpoint_of(case)
    cfg, cat, log, _ = _synthetic(8, False, hash_size=600)
    env = make_env(cfg, False, cat, log, w, output_format="numpy")
    obs, probs = env.sim.engine.dien_forward(g["seq"].astype(np.int32), g["dense"].astype(np.float32), g["cat"].astype(np.int32))
    assert_close_rel(obs.cpu().numpy(), g[case + "_obs"], what="cuda vs reference graph %s obs" % case)
    assert_close_rel(probs.cpu().numpy(), g[case + "_probs"], what="cuda vs reference graph %s probs" % case)
