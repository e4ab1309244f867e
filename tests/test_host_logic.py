"""CPU tests of host-side logic that needs no GPU."""
import random
import types

import numpy as np
import torch

from usip_b200.models._common import random_point_dropout


def test_random_point_dropout_matches_reference_statement_sequence():
    """keypoint_detector.py:161-169: one random.uniform, one np.random.choice, the same columns in every tensor."""
    opt = types.SimpleNamespace(random_pc_dropout_lower_limit=0.5, input_pc_num=100, device="cpu")
    a = torch.arange(2 * 3 * 100).float().view(2, 3, 100)
    b = a + 1000
    random.seed(3); np.random.seed(3)
    ra, rb = random_point_dropout(opt, a, b)
    random.seed(3); np.random.seed(3)
    n = round(random.uniform(0.5, 1.0) * 100)
    idx = torch.from_numpy(np.random.choice(100, n, replace=False))
    assert torch.equal(ra, torch.index_select(a, 2, idx)) and torch.equal(rb, torch.index_select(b, 2, idx))
    opt.random_pc_dropout_lower_limit = 1.0                      # disabled: tensors pass through, no RNG draw
    random.seed(5); np.random.seed(5)
    before = (random.random(), np.random.rand())
    random.seed(5); np.random.seed(5)
    out = random_point_dropout(opt, a, b)
    assert out[0] is a and out[1] is b and (random.random(), np.random.rand()) == before
