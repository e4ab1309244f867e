"""Node generation of the reference's data loaders on the GPU.

Only `FarthestSampler` of data/kitti_detector_loader.py:68-83 is mirrored (the same class is duplicated in the
ModelNet / Oxford loaders); datasets, augmentation and file formats stay out of scope.  The reference runs it per cloud
in DataLoader workers: serial numpy, O(k * Ns) float64 work, ~30 ms for k=512, Ns=5461 -- more than a whole training
step of this repository, hence on the critical path once the step is ~1 ms.
"""
import numpy as np
import torch

from usip_b200 import ops


class FarthestSampler:
    """Drop-in for the reference class: `sample(pts, k)` with pts an (Ns, 3) float array returns the (k, 3) float64
    array the numpy code returns -- bit-identical, including which random number it consumes (one
    np.random.randint(len(pts)) for the first node)."""

    def __init__(self, device=None):
        self.device = torch.device(device if device is not None else "cuda")

    def sample(self, pts, k):
        pts32 = np.ascontiguousarray(pts, dtype=np.float32)
        if pts32.shape[1] != 3:
            raise ValueError("pts must be (Ns, 3)")
        start = np.random.randint(len(pts32))                      # kitti_detector_loader.py:78
        idx = self.sample_indices(torch.from_numpy(pts32).to(self.device)[None], torch.tensor([start], dtype=torch.int32), k)[0]
        return pts32[idx[0].cpu().numpy()].astype(np.float64)

    def sample_indices(self, pts, start, k, want_nodes=False):
        """Batched form for a GPU-resident pipeline: pts (B, Ns, 3) f32 CUDA, start (B,) int -> (idx (B,k) i32,
        nodes (B,3,k) f32 in the detector's node layout or None)."""
        pts = pts.to(self.device, torch.float32).contiguous()
        start = torch.as_tensor(start).to(self.device, torch.int32).contiguous()
        if int(start.min()) < 0 or int(start.max()) >= pts.shape[1]:
            raise IndexError("start index out of range")
        return ops.fps(pts, start, k, want_nodes=want_nodes)
