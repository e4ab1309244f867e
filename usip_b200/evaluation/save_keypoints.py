"""Inference post-processing and wire formats of evaluation/save_keypoints.py: the step right
after the detector.  Mirrors the reference's function names where it has functions; the script body (dataset walking,
checkpoints, plotting) is out of scope.

    nms(keypoints_np, sigmas_np, NMS_radius)         save_keypoints.py:180-216   (GPU, bit-identical selection)
    select_keypoints(...)                            :336-351  nms + sigma-sorted top-k, batched on the device
    write_keypoints_bin / read_keypoints_bin         :389-391  float32 (M', 3) raw file
    read_pointcloud_npy                              data/kitti_detector_loader.py:116-139  (N, 8) .npy reader
"""
import numpy as np
import torch

from usip_b200 import ops


def nms(keypoints_np, sigmas_np, NMS_radius, device="cuda"):
    """Drop-in for the reference function: (M,3) keypoints, (M,) sigmas -> (valid_keypoints, valid_sigmas), emitted in
    ascending sigma (ties by index), every emitted keypoint suppressing the remaining ones within NMS_radius."""
    if NMS_radius < 0.01:
        return keypoints_np, sigmas_np
    kp = torch.from_numpy(np.ascontiguousarray(keypoints_np, np.float32).T.copy()).to(device)[None]
    sg = torch.from_numpy(np.ascontiguousarray(sigmas_np, np.float32)).to(device)[None]
    idx, cnt = ops.nms(kp, sg, NMS_radius)
    keep = idx[0, :int(cnt[0])].cpu().numpy()
    return keypoints_np[keep, :], sigmas_np[keep]


def select_keypoints(keypoints, sigmas, NMS_radius, desired_keypoint_num=None):
    """Batched form for a GPU-resident pipeline: keypoints (B,3,M), sigmas (B,M) CUDA f32 -> list of B float32 (M'_b, 3)
    CUDA tensors: NMS (or pass-through), then the `desired_keypoint_num` smallest sigmas (save_keypoints.py:346-351).
    After the NMS the list is already in ascending sigma; without it a stable sort supplies the order (np.argsort in the
    reference is not stable, so keypoints with EQUAL sigma may come out in another order there)."""
    keypoints = keypoints.contiguous(); sigmas = sigmas.contiguous()
    B, _, M = keypoints.shape
    idx, cnt = ops.nms(keypoints, sigmas, NMS_radius)
    out = []
    cnt_h = cnt.cpu().tolist()
    for b in range(B):
        keep = idx[b, :cnt_h[b]].long()
        if NMS_radius < 0.01:
            keep = keep[torch.argsort(sigmas[b][keep], stable=True)]
        if desired_keypoint_num is not None:
            keep = keep[:min(int(desired_keypoint_num), keep.numel())]
        out.append(keypoints[b][:, keep].t().contiguous())
    return out


def write_keypoints_bin(path, keypoints):
    """float32 (M', 3) row-major raw file, as `output.astype(np.float32).tofile(output_file)` (save_keypoints.py:389-391)."""
    a = keypoints.detach().cpu().numpy() if isinstance(keypoints, torch.Tensor) else np.asarray(keypoints)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError("keypoints must be (M, 3)")
    a.astype(np.float32).tofile(path)


def read_keypoints_bin(path):
    return np.fromfile(path, dtype=np.float32).reshape(-1, 3)


def read_pointcloud_npy(path, surface_normal_len=4):
    """KITTI / Oxford cloud file: (N, 8) = [x, y, z, nx, ny, nz, curvature, reflectance]
    (data/kitti_detector_loader.py:116-139) -> (pc (N,3), sn (N,surface_normal_len)) float32."""
    a = np.load(path)
    if a.ndim != 2 or a.shape[1] < 3 + max(surface_normal_len, 1):
        raise ValueError("expected an (N, >=%d) array, got %r" % (3 + surface_normal_len, a.shape))
    a = a.astype(np.float32)
    sn = a[:, a.shape[1] - 1:] if surface_normal_len == 1 else a[:, 3:3 + surface_normal_len]
    return a[:, 0:3], sn
