"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Import shim that lets the UNMODIFIED reference (lijx10/USIP) run as the checker / the baseline:

  * mode="cpu" (build container and `bench.py --impl reference`): the reference on the host cores, so that the
    oracle restatement (oracle/usip_oracle.py, oracle/usip_oracle.c) can be pinned against it, golden vectors can be
    generated for tests/golden/ (tools/make_golden.py) and the reference's own CPU path can be timed;
  * mode="cuda" (GPU box): the reference's own 1-GPU PyTorch path with its own two CUDA extensions
    (oracle/_ref/{index_max,ball_query}.so) -- the full-size floating-point oracle of tests/test_gpu_vs_reference.py
    and the `reference_gpu` denominator of bench.py.

The reference tree is imported from /root/reference where it is mounted (build container) and otherwise from the
byte-identical staged copy oracle/_ref/py/ (oracle/build_ref.py: stage_py; git-ignored, travels with gpurun), so the
GPU box never reads /root/reference.

What the shim does (follows SURVEY.md Appendix B):
  1. empty stub modules for matplotlib / mpl_toolkits / h5py (imported but unused on the path),
  2. `index_max` / `ball_query` operator modules backed by the reference's own C++ CPU code
     compiled from where it lies (oracle/build_ref.py -> oracle/_ref/), with a restatement of
     ball_query_cuda.cu:22-46 for ball_query (the reference has no CPU ball query),
  3. torch.cuda.device / synchronize / Tensor.get_device patched to CPU no-ops.
"""
import contextlib
import importlib
import os
import sys
import types

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_installed = None          # None | "cpu" | "cuda"


def reference_root():
    from . import build_ref
    return build_ref.reference_py_root()


def build_ref_root():
    """The mounted reference (files that are NOT staged, e.g. evaluation/save_keypoints.py, only exist there)."""
    from . import build_ref
    return build_ref.REFERENCE_ROOT


def reference_available() -> bool:
    return reference_root() is not None


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _ball_query_restatement(node_to_point_dist, radius, K):
    """Restatement of /root/reference/models/ball_query_ext/ball_query_cuda.cu:22-46 (CPU)."""
    from . import usip_oracle as orc
    d = node_to_point_dist.detach().contiguous().cpu().numpy()
    return torch.from_numpy(orc.ball_query_dist(d, float(radius), int(K)))


def install(use_ref_ext: bool = True, mode: str = "cpu", operators: str = "reference"):
    """Make `import models.networks` etc. resolve to the reference (CPU shims, or its own CUDA extensions).
    operators="dropin" (cuda mode): leave `index_max` / `ball_query` to sys.path, with <repo>/usip_b200/dropin in front --
    the literal operator-level recipe of INTEGRATION.md: the unmodified reference on THIS repo's operators."""
    global _installed
    if _installed is not None:
        if _installed != mode:
            raise RuntimeError("ref_shim already installed in %s mode; one mode per process" % _installed)
        return
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree neither mounted nor staged under oracle/_ref/py (run oracle/build_ref.py "
                           "in the build container)")

    # 1. stubs
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    mpl.cm = _stub("matplotlib.cm", jet=None)
    tk = _stub("mpl_toolkits")
    tk.mplot3d = _stub("mpl_toolkits.mplot3d", Axes3D=None)
    _stub("h5py")

    if mode == "cuda":
        if operators == "dropin":
            _finish(root, mode)
            sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "usip_b200", "dropin"))
            for name in ("index_max", "ball_query"):
                sys.modules.pop(name, None)
            return
        # 2'. the reference's own CUDA extensions, compiled from its sources by oracle/build_ref.py
        from . import build_ref
        sys.modules["index_max"] = build_ref._load_so("index_max")
        sys.modules["ball_query"] = build_ref._load_so("ball_query")
        _finish(root, mode)
        return

    # 2. operator modules
    ref_im = None
    if use_ref_ext:
        try:
            from . import build_ref
            ref_im = build_ref.load_reference_index_max()
        except Exception as e:  # pragma: no cover - falls back to the C restatement
            print("[ref_shim] reference index_max ext unavailable (%s); using oracle C restatement" % e)
    from . import usip_oracle as orc

    def _im_cpu(data, index, K):
        if ref_im is not None:
            return ref_im.forward_cpu(data.detach().contiguous().cpu(), index.contiguous().cpu(), int(K))
        return torch.from_numpy(orc.index_max(data.detach().cpu().numpy(), index.cpu().numpy(), int(K)))

    _stub("index_max", forward_cuda_shared_mem=_im_cpu, forward_cuda=_im_cpu, forward_cpu=_im_cpu)
    _stub("ball_query", forward_cuda_shared_mem=_ball_query_restatement)

    # 3. CPU patches
    torch.cuda.device = lambda *a, **k: contextlib.nullcontext()
    torch.cuda.synchronize = lambda *a, **k: None
    torch.Tensor.get_device = lambda self: 0

    _finish(root, mode)


def _finish(root, mode):
    global _installed
    if root in sys.path:
        sys.path.remove(root)
    sys.path.insert(0, root)
    # our own package also has a `models` mirror; make sure the reference wins under this shim
    for name in list(sys.modules):
        if name.split(".")[0] in ("models", "util", "data"):
            del sys.modules[name]
    _installed = mode


def make_opt(**over):
    """Hand-built option object with the fields the hot path reads.
    Defaults are the KITTI detector defaults (kitti/options_detector.py:22-36)."""
    o = types.SimpleNamespace(
        gpu_ids=[-1], device=torch.device("cpu"), scene="outdoor",
        batch_size=2, input_pc_num=1024, surface_normal_len=4, node_num=64, k=1, node_knn_k_1=16,
        activation="relu", normalization="batch", bn_momentum=0.1, bn_momentum_decay_step=None,
        bn_momentum_decay=0.6, lr=0.001, loss_sigma_lower_bound=0.001,
        random_pc_dropout_lower_limit=1.0, keypoint_on_pc_type="point_to_point",
        keypoint_on_pc_alpha=0.01, rot_3d=False, rot_horizontal=True, checkpoints_dir="/tmp",
        # descriptor
        ball_radius=1.0, ball_nsamples=64, descriptor_len=128, sigma_max=3.0, triple_loss_gamma=0.5,
    )
    o.__dict__.update(over)
    return o


def modules(mode="cpu", operators="reference"):
    install(mode=mode, operators=operators)
    networks = importlib.import_module("models.networks")
    losses = importlib.import_module("models.losses")
    layers = importlib.import_module("models.layers")
    som = importlib.import_module("util.som")
    return types.SimpleNamespace(networks=networks, losses=losses, layers=layers, som=som,
                                 keypoint_detector=importlib.import_module("models.keypoint_detector"),
                                 keypoint_descriptor=importlib.import_module("models.keypoint_descriptor"))
