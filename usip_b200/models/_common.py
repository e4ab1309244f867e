"""Host-side helpers shared by ModelDetector and ModelDescriptor."""
import random

import numpy as np
import torch


def random_point_dropout(opt, *clouds):
    """Training-time point dropout of both model wrappers (keypoint_detector.py:161-169, keypoint_descriptor.py:128-137):
    unless `random_pc_dropout_lower_limit >= 0.99`, keep round(u * input_pc_num) randomly chosen point columns, the same
    ones in every (B, C, N) tensor given.  The random numbers are drawn like the reference does -- one random.uniform,
    then one np.random.choice without replacement -- so seeded runs keep the same points."""
    if opt.random_pc_dropout_lower_limit >= 0.99:
        return clouds
    n_keep = round(random.uniform(opt.random_pc_dropout_lower_limit, 1.0) * opt.input_pc_num)
    columns = torch.from_numpy(np.random.choice(opt.input_pc_num, n_keep, replace=False)).to(opt.device)
    return tuple(torch.index_select(c, 2, columns) for c in clouds)
