"""Host mirror of rl4rs/env/seqslate.py: SeqSlateState / SeqSlateRecEnv (page-wise slates).

All page logic (window of 9 slots, sequence_id = step//9+1, layer = step%9//3, mask reset at page
boundaries, previous pages as the second sequence, per-page rewards, the violation quirks Q8-Q10)
runs in the CUDA library; see csrc/r4_kernels.cuh (k_act, k_assemble, k_seq_ids, k_reward).
"""
from .slate import SlateRecEnv, SlateState


class SeqSlateState(SlateState):
    seq = True

    def __init__(self, config, records, engine=None):
        super().__init__(config, records, engine)
        self.page_items = config.get("page_items", 9)


class SeqSlateRecEnv(SlateRecEnv):
    seq = True

    def __init__(self, config, state_cls=SeqSlateState):
        super().__init__(config, state_cls)
        self.page_items = config.get("page_items", 9)
