"""The drop-in recipes of INTEGRATION.md, executed literally in a fresh interpreter (CPU: imports only, no compute)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "usip_b200")


def _run(code, cwd="/"):
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, "-c", code], cwd=cwd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_model_level_recipe_resolves_to_this_repo():
    """sys.path.insert(0, '<repo>/usip_b200') then the reference's own import lines (models/networks.py:9-18,
    models/keypoint_detector.py:9-12, kitti/train_detector.py) -- no other path set, cwd elsewhere."""
    out = _run(
        "import sys; sys.path.insert(0, %r)\n"
        "import index_max, ball_query\n"
        "from models import networks, losses, operations, layers\n"
        "from models.keypoint_detector import ModelDetector\n"
        "from models.keypoint_descriptor import ModelDescriptor\n"
        "from util import som\n"
        "from data.kitti_detector_loader import FarthestSampler\n"
        "from evaluation import save_keypoints\n"
        "for m in (index_max, ball_query, networks, losses, operations, layers, som, save_keypoints):\n"
        "    print(m.__file__)\n"
        "assert callable(index_max.forward_cuda_shared_mem) and callable(ball_query.forward_cuda_shared_mem)\n"
        "assert callable(som.query_topk) and hasattr(networks, 'RPN_Detector') and hasattr(networks, 'DescriptorLiteOld')\n"
        % PKG)
    files = out.split()
    assert len(files) == 8 and all(f.startswith(PKG + os.sep) for f in files), files


def test_operator_level_recipe_resolves_only_the_two_extension_names():
    out = _run(
        "import sys; sys.path.insert(0, %r)\n"
        "import index_max, ball_query\n"
        "print(index_max.__file__); print(ball_query.__file__)\n"
        "import importlib.util as u\n"
        "print(u.find_spec('models') is None or %r not in (u.find_spec('models').origin or ''))\n"
        % (os.path.join(PKG, "dropin"), PKG))
    a, b, shadow_free = out.split()
    assert a == os.path.join(PKG, "dropin", "index_max.py") and b == os.path.join(PKG, "dropin", "ball_query.py")
    assert shadow_free == "True"          # the reference's own models/ stays in charge at this level


def test_package_style_import_still_works():
    out = _run("import sys; sys.path.insert(0, %r)\n"
               "from usip_b200 import index_max, ball_query\n"
               "from usip_b200.models.keypoint_detector import ModelDetector\n"
               "from usip_b200.util import som\nprint('ok')\n" % ROOT)
    assert out.strip() == "ok"


@pytest.mark.gpu
def test_model_level_recipe_runs_on_gpu():
    """Same recipe, then one eval forward through the bare-name modules."""
    out = _run(
        "import sys; sys.path.insert(0, %r); sys.path.insert(1, %r)\n"
        "import torch, numpy as np, index_max\n"
        "from models.keypoint_detector import ModelDetector\n"
        "from tests.util_gpu import make_opt\n"
        "md = ModelDetector(make_opt(batch_size=1, input_pc_num=2048, node_num=64))\n"
        "g = torch.Generator().manual_seed(0)\n"
        "pc = torch.rand(1, 3, 2048, generator=g) * 20; sn = torch.rand(1, 4, 2048, generator=g); node = pc[:, :, :64].clone()\n"
        "md.set_input(pc, sn, node, pc, sn, node, torch.eye(3)[None], torch.ones(1), torch.zeros(1, 3, 1))\n"
        "md.test_model(); print(float(md.loss))\n"
        "d = torch.randn(2, 4, 100, device='cuda'); i = torch.randint(0, 8, (2, 100), device='cuda', dtype=torch.int32)\n"
        "print(tuple(index_max.forward_cuda_shared_mem(d, i, 8).shape))\n" % (PKG, ROOT), cwd=ROOT)
    lines = out.strip().splitlines()
    assert float(lines[0]) == float(lines[0]) and lines[1] == "(2, 4, 8)"
