"""TEST INFRASTRUCTURE ONLY.  Recipe that compiles the reference's own two operator extensions
from the sources WHERE THEY LIE under /root/reference (nothing is copied into this repo) into
oracle/_ref/ (git-ignored, travels to the GPU box with gpurun):

    /root/reference/models/index_max_ext/{index_max.cpp,index_max_cuda.cu}   -> oracle/_ref/index_max/index_max.so
    /root/reference/models/ball_query_ext/{ball_query.cpp,ball_query_cuda.cu} -> oracle/_ref/ball_query/ball_query.so

They are used (a) here, on CPU, as the `forward_cpu` oracle that pins oracle/usip_oracle.c and
(b) on the GPU box as the bit-exactness oracle + "kernel to beat" for index_max / ball_query
(tests/test_gpu_ops.py, bench.py --ops).  The reference's own setup.py is NOT run; we call
torch.utils.cpp_extension.load on the two source files directly, arch pinned to sm_100a.
"""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_OUT = os.path.join(_HERE, "_ref")
REFERENCE_ROOT = os.environ.get("USIP_REFERENCE_ROOT", "/root/reference")

_EXTS = {
    "index_max": ["models/index_max_ext/index_max.cpp", "models/index_max_ext/index_max_cuda.cu"],
    "ball_query": ["models/ball_query_ext/ball_query.cpp", "models/ball_query_ext/ball_query_cuda.cu"],
}


def _so_path(name):
    return os.path.join(REF_OUT, name, name + ".so")


def build(name, verbose=False):
    """Compile one reference extension (needs /root/reference; nvcc cross-compiles w/o a GPU)."""
    from torch.utils import cpp_extension
    srcs = [os.path.join(REFERENCE_ROOT, s) for s in _EXTS[name]]
    for s in srcs:
        if not os.path.isfile(s):
            raise FileNotFoundError(s)
    out = os.path.join(REF_OUT, name)
    os.makedirs(out, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    cpp_extension.load(
        name=name, sources=srcs, build_directory=out, verbose=verbose,
        extra_cflags=["-O2", "-w"],
        extra_cuda_cflags=["-O2", "-w", "-gencode", "arch=compute_100a,code=sm_100a"],
        is_python_module=False,
    )
    return _so_path(name)


def _load_so(name):
    path = _so_path(name)
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    import torch  # noqa: F401  (the .so links against libtorch)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_index_max():
    if not os.path.isfile(_so_path("index_max")):
        build("index_max")
    return _load_so("index_max")


def load_reference_ball_query():
    if not os.path.isfile(_so_path("ball_query")):
        build("ball_query")
    return _load_so("ball_query")


def have(name):
    return os.path.isfile(_so_path(name))


# The unmodified Python files of the reference's hot path.  They are STAGED (byte-for-byte, never edited) next to the two
# compiled extensions under the git-ignored oracle/_ref/py/ so that the GPU box -- where /root/reference does not exist --
# can run the reference itself: as the full-size floating-point oracle (tests/test_gpu_vs_reference.py) and as the
# "reference 1-GPU PyTorch path" denominator of bench.py (`reference_gpu`).  Nothing staged is committed.
PY_OUT = os.path.join(REF_OUT, "py")
_PY_FILES = [
    "models/keypoint_detector.py", "models/keypoint_descriptor.py", "models/networks.py", "models/layers.py",
    "models/losses.py", "models/operations.py",
    "util/__init__.py", "util/som.py", "util/potential_field.py", "util/vis_tools.py",
    "data/augmentation.py",
]


def stage_py():
    """Copy the hot-path .py files of the mounted reference into oracle/_ref/py/ (idempotent; returns the directory)."""
    import shutil
    for rel in _PY_FILES:
        src = os.path.join(REFERENCE_ROOT, rel)
        if not os.path.isfile(src):
            raise FileNotFoundError(src)
        dst = os.path.join(PY_OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.isfile(dst) or open(src, "rb").read() != open(dst, "rb").read():
            shutil.copyfile(src, dst)
    return PY_OUT


def have_py():
    return all(os.path.isfile(os.path.join(PY_OUT, rel)) for rel in _PY_FILES)


def reference_py_root():
    """Where the reference's Python tree can be imported from: the mount (build container) or the staged copy."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "models")):
        return REFERENCE_ROOT
    if have_py():
        return PY_OUT
    return None


def build_all(verbose=False):
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "models")):
        return False
    for n in _EXTS:
        if not have(n):
            build(n, verbose=verbose)
    stage_py()
    return True


if __name__ == "__main__":
    ok = build_all(verbose="-v" in sys.argv)
    print("oracle/_ref built" if ok else "reference not mounted; nothing built")
