"""Thin torch-tensor wrappers over the C ABI (include/usip_b200.h).  PyTorch is used for device
memory and streams only; every computation below happens in libusip_b200.so.  All functions launch on
torch's current CUDA stream and never synchronise."""
import ctypes

import torch

from . import _lib
from ._lib import LayerDesc, check

i32 = torch.int32
f32 = torch.float32


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _req(t, dtype, name):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor/variable" % name)       # index_max.cpp:119-121
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s: expected dtype %s, got %s" % (name, dtype, t.dtype))


def tile_rows():
    return _lib.load().usip_layer_tile_rows()


def stat_slots(P, Cout, precision=0, group=0, want_group=False):
    """Rows of the [slots, 2, Cout] BN-statistic partial buffer usip_layer_fwd fills for this layer shape."""
    d = _lib.LayerDesc()
    d.P, d.Cout, d.precision, d.group = int(P), int(Cout), int(precision), int(group)
    d.gmax = 1 if want_group else None          # only tested for NULL-ness by the query
    return _lib.load().usip_layer_stat_slots(ctypes.byref(d))


# ----------------------------------------------------------------------------- reference operators
def index_max(data, index, K):
    """(B,C,N) f32, (B,N) i32 -> (B,C,K) i32.  index_max.forward_cuda_shared_mem semantics."""
    _req(data, f32, "data"); _req(index, i32, "index")
    B, C, N = data.shape
    out = torch.empty((B, C, int(K)), dtype=i32, device=data.device)
    scratch = None
    if K * 12 > 200 * 1024:
        scratch = torch.empty((B * C * K,), dtype=torch.int64, device=data.device)
    with torch.cuda.device(data.device):
        check(_lib.load().usip_index_max_f32(_p(data), _p(index), _p(out), _p(scratch), B, C, N, int(K), _stream()),
              "usip_index_max_f32")
    return out


def ball_query_dist(dist, radius, K):
    _req(dist, f32, "node_to_point_dist")
    B, M, N = dist.shape
    out = torch.empty((B, M, int(K)), dtype=i32, device=dist.device)
    with torch.cuda.device(dist.device):
        check(_lib.load().usip_ball_query_dist_f32(_p(dist), float(radius), _p(out), B, M, N, int(K), _stream()),
              "usip_ball_query_dist_f32")
    return out


_BG_SCRATCH = {}


def _ball_group_scratch(lib, device, B, S, N, M, K):
    """Persistent scratch of the bucket-grid ball query, one per (device, stream, batch): its counter region is cleared once
    here and left cleared by every call (include/usip_b200.h), so the hot path launches no memset."""
    stream = torch.cuda.current_stream()
    key = (device.index, stream.cuda_stream, B)
    ent = _BG_SCRATCH.get(key)
    if ent is None:
        nbytes = int(lib.usip_ball_group_scratch_bytes(B, S, N, M, K))
        buf = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=device)
        check(lib.usip_ball_group_scratch_init(_p(buf), nbytes, B, _stream()), "usip_ball_group_scratch_init")
        _lib.LAUNCHES[0] -= 1                      # a memset, not one of our kernels
        ent = _BG_SCRATCH[key] = (buf, nbytes)
    return ent


def ball_group(xyz, feat, centers, radius, K, want_group=True, rows_ld=0):
    """Fused distance + ball query + gather + decentre.  Returns (idx (B,M,K) i32, group (B,3+S,M,K) f32)."""
    _req(xyz, f32, "xyz"); _req(centers, f32, "centers")
    B, _, N = xyz.shape
    M = centers.shape[2]
    S = 0 if feat is None else feat.shape[1]
    if S:
        _req(feat, f32, "feat")
    lib = _lib.load()
    idx = torch.empty((B, M, int(K)), dtype=i32, device=xyz.device)
    grp = torch.empty((B, 3 + S, M, int(K)), dtype=f32, device=xyz.device) if want_group else None
    rows = torch.empty((B * M * int(K), rows_ld), dtype=f32, device=xyz.device) if rows_ld else None
    with torch.cuda.device(xyz.device):
        scratch, nbytes = _ball_group_scratch(lib, xyz.device, B, S, N, M, int(K))
        check(lib.usip_ball_group_f32(_p(xyz), _p(feat) if S else None, _p(centers), float(radius), _p(idx), _p(grp),
                                      _p(rows), rows_ld, _p(scratch), int(nbytes), B, S, N, M, int(K), _stream()),
              "usip_ball_group_f32")
    if rows_ld:
        return idx, grp, rows
    return idx, grp


def knn_group(xyz, feat, centers, K, want_group=True, rows_ld=0):
    """K nearest points of every centre + gather + decentre (RPN_Detector_KNN front end, networks.py:556-565).
    Returns (idx (B,M,K) i32 ascending, group (B,3+S,M,K) f32 or None[, rows [B*M*K, rows_ld]])."""
    _req(xyz, f32, "xyz"); _req(centers, f32, "centers")
    B, _, N = xyz.shape
    M = centers.shape[2]
    S = 0 if feat is None else feat.shape[1]
    if S:
        _req(feat, f32, "feat")
    idx = torch.empty((B, M, int(K)), dtype=i32, device=xyz.device)
    grp = torch.empty((B, 3 + S, M, int(K)), dtype=f32, device=xyz.device) if want_group else None
    rows = torch.empty((B * M * int(K), rows_ld), dtype=f32, device=xyz.device) if rows_ld else None
    with torch.cuda.device(xyz.device):
        check(_lib.load().usip_knn_group_f32(_p(xyz), _p(feat) if S else None, _p(centers), _p(idx), _p(grp), _p(rows), rows_ld,
                                             B, S, N, M, int(K), _stream()), "usip_knn_group_f32")
    if rows_ld:
        return idx, grp, rows
    return idx, grp


def fps(pts, start, k, want_nodes=True):
    """Farthest point sampling.  pts (B,Ns,3) f32, start (B,) i32 -> (idx (B,k) i32, nodes (B,3,k) f32 or None)."""
    _req(pts, f32, "pts"); _req(start, i32, "start")
    B, Ns, _ = pts.shape
    idx = torch.empty((B, int(k)), dtype=i32, device=pts.device)
    nodes = torch.empty((B, 3, int(k)), dtype=f32, device=pts.device) if want_nodes else None
    with torch.cuda.device(pts.device):
        check(_lib.load().usip_fps_f32(_p(pts), _p(start), _p(idx), _p(nodes), B, Ns, int(k), _stream()), "usip_fps_f32")
    return idx, nodes


def nms(keypoints, sigmas, radius):
    """keypoints (B,3,M) f32, sigmas (B,M) f32 -> (idx (B,M) i32 kept indices in emission order, -1 padded; count (B,) i32)."""
    _req(keypoints, f32, "keypoints"); _req(sigmas, f32, "sigmas")
    B, _, M = keypoints.shape
    idx = torch.empty((B, M), dtype=i32, device=keypoints.device)
    cnt = torch.empty((B,), dtype=i32, device=keypoints.device)
    with torch.cuda.device(keypoints.device):
        check(_lib.load().usip_nms_f32(_p(keypoints), _p(sigmas), float(radius), _p(idx), _p(cnt), B, M, _stream()), "usip_nms_f32")
    return idx, cnt


def knn_gather(src, idx):
    _req(src, f32, "som_node"); _req(idx, i32, "som_node_knn_I")
    B, C, N = src.shape
    M, K = idx.shape[1], idx.shape[2]
    out = torch.empty((B, C, M, K), dtype=f32, device=src.device)
    check(_lib.load().usip_knn_gather_f32(_p(src), _p(idx), _p(out), B, C, N, M, K, _stream()), "usip_knn_gather_f32")
    return out


# ----------------------------------------------------------------------------- grouping
def som_assign(xyz, node, count=None, method="auto"):
    """Nearest node of every point (som.query_topk, k = 1).  method: "auto" | "brute" | "grid" (identical results): the grid
    path bins the nodes into cells and searches the shells around each point (usip_som_assign_grid_f32)."""
    _req(xyz, f32, "x"); _req(node, f32, "node")
    B, _, N = xyz.shape
    M = node.shape[2]
    min_idx = torch.empty((B, N), dtype=i32, device=xyz.device)
    if count is None:
        count = torch.zeros((B, M), dtype=i32, device=xyz.device)
    lib = _lib.load()
    if method == "grid" or (method == "auto" and 64 <= M <= 1024 and N >= 1024):
        nbytes = int(lib.usip_som_assign_grid_scratch_bytes(B, M))
        scratch = torch.empty(((nbytes + 15) // 16, 4), dtype=i32, device=xyz.device)
        check(lib.usip_som_assign_grid_f32(_p(xyz), _p(node), _p(min_idx), _p(count), _p(scratch), nbytes, B, N, M, _stream()),
              "usip_som_assign_grid_f32")
        return min_idx, count
    check(lib.usip_som_assign_f32(_p(xyz), _p(node), _p(min_idx), _p(count), B, N, M, _stream()),
          "usip_som_assign_f32")
    return min_idx, count


def cluster_sort(min_idx, M):
    B, N = min_idx.shape
    dev = min_idx.device
    seg_off = torch.empty((B, M + 1), dtype=i32, device=dev)
    perm = torch.empty((B, N), dtype=i32, device=dev)
    row_seg = torch.empty((B, N), dtype=i32, device=dev)
    chunks = (N + 255) // 256
    scratch = torch.empty((B * chunks * M,), dtype=i32, device=dev)
    check(_lib.load().usip_cluster_sort(_p(min_idx), _p(seg_off), _p(perm), _p(row_seg), _p(scratch), B, N, M,
                                        _stream()), "usip_cluster_sort")
    return seg_off, perm, row_seg


def cluster_mean_decenter(xyz, feat, seg_off, perm, M, ldx=8):
    B, _, N = xyz.shape
    S = 0 if feat is None else feat.shape[1]
    cmean = torch.empty((B, 3, M), dtype=f32, device=xyz.device)
    x_aug = torch.empty((B * N, ldx), dtype=f32, device=xyz.device)
    check(_lib.load().usip_cluster_mean_decenter(_p(xyz), _p(feat) if S else None, _p(seg_off), _p(perm), _p(cmean),
                                                 _p(x_aug), ldx, B, S, N, M, _stream()), "usip_cluster_mean_decenter")
    return cmean, x_aug


def segmax(X, C, seg_off, perm, B, N, M, out=None, arg=None, want_arg=True):
    ldx = X.stride(0)
    if out is None:
        out = torch.empty((B * M, C), dtype=f32, device=X.device)
    if arg is None and want_arg:
        arg = torch.empty((B * M, C), dtype=i32, device=X.device)
    check(_lib.load().usip_segmax(_p(X), ldx, _p(seg_off), _p(perm), _p(out), out.stride(0), _p(arg), B, N, M, C,
                                  _stream()), "usip_segmax")
    return out, arg


def knn_nodes(pts, K):
    _req(pts, f32, "pts")
    B, _, M = pts.shape
    out = torch.empty((B, M, K), dtype=i32, device=pts.device)
    check(_lib.load().usip_knn_nodes(_p(pts), _p(out), B, M, K, _stream()), "usip_knn_nodes")
    return out


# ----------------------------------------------------------------------------- shared-MLP stack
def layer_fwd(X, W, bias, P, Cin, Cout, ldx=None, ldw=None, in_scale=None, in_shift=None, in_relu=False,
              addend=None, add_index=None, add_group=0, Y=None, ldy=None, stat_partial=None,
              gmax=None, gmin=None, garg_max=None, garg_min=None, group=0, precision=0, tc_ws=None, tc_packed=False, w_transposed=False, debug_flags=0, debug_clocks=None,
              pack_entry=None):
    d = LayerDesc()
    d.X = X.data_ptr(); d.ldx = X.stride(0) if ldx is None else ldx
    d.P = P; d.Cin = Cin; d.Cout = Cout
    d.W = W.data_ptr(); d.ldw = (W.stride(0) if ldw is None else ldw)
    d.w_transposed = 1 if w_transposed else 0
    d.bias = None if bias is None else bias.data_ptr()
    d.in_scale = None if in_scale is None else in_scale.data_ptr()
    d.in_shift = None if in_shift is None else in_shift.data_ptr()
    d.in_relu = 1 if in_relu else 0
    d.addend = None if addend is None else addend.data_ptr()
    d.ld_add = 0 if addend is None else addend.stride(0)
    d.add_index = None if add_index is None else add_index.data_ptr()
    d.add_group = add_group
    d.Y = None if Y is None else Y.data_ptr()
    d.ldy = 0 if Y is None else (Y.stride(0) if ldy is None else ldy)
    d.stat_partial = None if stat_partial is None else stat_partial.data_ptr()
    d.gmax = None if gmax is None else gmax.data_ptr()
    d.gmin = None if gmin is None else gmin.data_ptr()
    d.garg_max = None if garg_max is None else garg_max.data_ptr()
    d.garg_min = None if garg_min is None else garg_min.data_ptr()
    d.group = group
    d.precision = precision
    d.debug_flags = debug_flags
    d.debug_clocks = debug_clocks.data_ptr() if debug_clocks is not None else None
    if precision == 1:
        if tc_ws is None:
            tc_ws = torch.empty((2 * Cin * Cout,), dtype=f32, device=X.device)
        d.tc_workspace = tc_ws.data_ptr(); d.tc_workspace_bytes = tc_ws.numel() * 4
        d.tc_weights_packed = 1 if tc_packed else 0
        if pack_entry is not None and pack_entry.desc is None:
            # remembered for engine.prepack_weights(): the pack of this layer can then join the one-launch re-pack of a
            # train step (the copy keeps W / workspace pointers; the tensors are kept alive next to it)
            pack_entry.desc = LayerDesc.from_buffer_copy(d)
            pack_entry.keep = (W, tc_ws)
    check(_lib.load().usip_layer_fwd(ctypes.byref(d), _stream()),
          "usip_layer_fwd_tc" if (precision == 1 and not tc_packed) else "usip_layer_fwd")


def bn_finalize(stat_partial, ntiles, count, C, gamma, beta, eps, momentum, running_mean, running_var,
                scale, shift, save_mean=None, save_invstd=None):
    check(_lib.load().usip_bn_finalize(_p(stat_partial), ntiles, int(count), C, _p(gamma), _p(beta), float(eps),
                                       float(momentum), _p(running_mean), _p(running_var), _p(scale), _p(shift),
                                       _p(save_mean), _p(save_invstd), _stream()), "usip_bn_finalize")


def bn_eval_affine(gamma, beta, running_mean, running_var, eps, scale, shift):
    C = gamma.numel()
    check(_lib.load().usip_bn_eval_affine(_p(gamma), _p(beta), _p(running_mean), _p(running_var), float(eps), C,
                                          _p(scale), _p(shift), _stream()), "usip_bn_eval_affine")


def knn_combine(Z, pts, knn_idx, W, ldw, bias, Y, stat_partial, B, M, K, Cout):
    check(_lib.load().usip_knn_combine(_p(Z), Z.stride(0), _p(pts), _p(knn_idx), _p(W), ldw, _p(bias), _p(Y),
                                       Y.stride(0), _p(stat_partial), B, M, K, Cout, _stream()), "usip_knn_combine")


def group_select(gmax, gmin, scale, shift, out, Q, C):
    check(_lib.load().usip_group_select(_p(gmax), _p(gmin), _p(scale), _p(shift), _p(out), out.stride(0), Q, C,
                                        _stream()), "usip_group_select")


def head_finalize(out4, cluster_mean, lb, B, M):
    kp = torch.empty((B, 3, M), dtype=f32, device=out4.device)
    sig = torch.empty((B, M), dtype=f32, device=out4.device)
    check(_lib.load().usip_head_finalize(_p(out4), out4.stride(0), _p(cluster_mean), float(lb), _p(kp), _p(sig), B, M,
                                         _stream()), "usip_head_finalize")
    return kp, sig


# ----------------------------------------------------------------------------- losses
# databases at least this large go through the cell grid (usip_pairwise_min_grid_f32); smaller ones stay brute force
PAIRWISE_MIN_GRID_FROM = 4096


def pairwise_min(a, b, method="auto"):
    """a (B,3,Ma), b (B,3,Nb) -> (min_d (B,Ma) f32, arg (B,Ma) i32).  method: "auto" | "brute" | "grid" (same result)."""
    _req(a, f32, "a"); _req(b, f32, "b")
    B, _, Ma = a.shape
    Nb = b.shape[2]
    d = torch.empty((B, Ma), dtype=f32, device=a.device)
    arg = torch.empty((B, Ma), dtype=i32, device=a.device)
    lib = _lib.load()
    if method == "grid" or (method == "auto" and Nb >= PAIRWISE_MIN_GRID_FROM):
        nbytes = int(lib.usip_pairwise_min_grid_scratch_bytes(B, Nb))
        scratch = torch.empty(((nbytes + 15) // 16, 4), dtype=i32, device=a.device)
        check(lib.usip_pairwise_min_grid_f32(_p(a), _p(b), _p(d), _p(arg), _p(scratch), nbytes, B, Ma, Nb, _stream()),
              "usip_pairwise_min_grid_f32")
        return d, arg
    packed = torch.empty((B, Ma), dtype=torch.int64, device=a.device)
    check(lib.usip_pairwise_min_f32(_p(a), _p(b), _p(d), _p(arg), _p(packed), B, Ma, Nb, _stream()),
          "usip_pairwise_min_f32")
    if Nb <= 2048:
        _lib.LAUNCHES[0] -= 2          # small databases take the single-kernel path (no init / finish launches)
    return d, arg


def chamfer_prob_reduce(d_sd, i_sd, d_ds, i_ds, sig_src, sig_dst):
    B, M = d_sd.shape
    N = d_ds.shape[1]
    out = torch.empty((3,), dtype=f32, device=d_sd.device)
    check(_lib.load().usip_chamfer_prob_reduce(_p(d_sd), _p(i_sd), _p(d_ds), _p(i_ds), _p(sig_src), _p(sig_dst),
                                               _p(out), B, M, N, _stream()), "usip_chamfer_prob_reduce")
    return out


def transform_points(kp, R, scale, shift):
    B, _, M = kp.shape
    out = torch.empty_like(kp)
    check(_lib.load().usip_transform_points(_p(kp), _p(R), _p(scale), _p(shift), _p(out), B, M, _stream()),
          "usip_transform_points")
    return out


def mean_scale(d, alpha):
    out = torch.empty((1,), dtype=f32, device=d.device)
    check(_lib.load().usip_mean_scale(_p(d), d.numel(), float(alpha), _p(out), _stream()), "usip_mean_scale")
    return out
