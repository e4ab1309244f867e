import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.test_gpu_detector import _setup
np.set_printoptions(precision=4, linewidth=220)
g, d, P, md = _setup("detector_kitti_small.npz", False)
names = ['knnlayer_1.layers_after.1.norm.bias', 'mlp1.norm.bias', 'first_pointnet.layers.0.norm.weight', 'mlp2.conv.weight']
before = {k: v.detach().cpu().numpy().reshape(-1)[:12].copy() for k, v in md.detector.state_dict().items() if k in names}
md.test_model()
md.optimize(epoch=0)
torch.cuda.synchronize()
sd = md.detector.state_dict()
pd = dict(md.detector.named_parameters())
for k in names:
    print(k)
    print('  before   ', before[k])
    print('  ours grad', pd[k].grad.cpu().numpy().reshape(-1)[:12])
    print('  ref  grad', g['grad/' + k][4:16])
    print('  ours after', sd[k].cpu().numpy().reshape(-1)[:12])
    print('  ref  after', g['after/' + k][3:15])
    print('  ours delta', sd[k].cpu().numpy().reshape(-1)[:12] - before[k])
    print('  ref  delta', g['after/' + k][3:15] - before[k])
