"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by usip_b200/, only by tests/,
__graft_entry__.smoke() as the checker, and bench.py's cpu_baseline / --impl reference legs).

CPU restatement (numpy + the plain-C helpers in usip_oracle.c) of the USIP detector / descriptor
hot path.  Every function cites the reference file:line it follows (relative to /root/reference).

Parity pinning: the reference ships no golden vectors for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in the build container by
tools/make_golden.py (reference imported unmodified through oracle/ref_shim.py, its index_max C++
compiled from its own sources by oracle/build_ref.py) and frozen under tests/golden/*.npz.
tests/test_oracle_vs_golden.py re-checks the oracle against those files everywhere (no reference
needed at test time).
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIB = None

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    """gcc -O2 -ffp-contract=off the C restatement into oracle/_build/liboracle.so."""
    src = os.path.join(_HERE, "usip_oracle.c")
    out = os.path.join(_BUILD, "liboracle.so")
    if force or not os.path.isfile(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(_BUILD, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                               "-o", out, src, "-lm"])
    return out


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _fp(a):
    return a.ctypes.data_as(_f)


def _ip(a):
    return a.ctypes.data_as(_i)


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ----------------------------------------------------------------------------- integer / index ops
def index_max(data, index, K):
    """models/index_max_ext/index_max.cpp:73-112."""
    data = _c32(data); index = np.ascontiguousarray(index, dtype=np.int32)
    B, C, N = data.shape
    out = np.zeros((B, C, K), np.int32)
    _lib().orc_index_max(_fp(data), _ip(index), _ip(out), B, C, N, K)
    return out


def ball_query_dist(dist, radius, K):
    """models/ball_query_ext/ball_query_cuda.cu:22-46."""
    dist = _c32(dist)
    B, M, N = dist.shape
    out = np.zeros((B, M, K), np.int32)
    _lib().orc_ball_query_dist(_fp(dist), ctypes.c_float(radius), _ip(out), B, M, N, K)
    return out


def ball_query_xyz(xyz, centers, radius, K):
    """models/networks.py:355-359 (torch.norm distance matrix + ball query), fused."""
    xyz = _c32(xyz); centers = _c32(centers)
    B, _, N = xyz.shape; M = centers.shape[2]
    out = np.zeros((B, M, K), np.int32)
    _lib().orc_ball_query_xyz(_fp(xyz), _fp(centers), ctypes.c_float(radius), _ip(out), B, N, M, K)
    return out


def som_assign(xyz, node):
    """util/som.py:35-39 (k=1): nearest node per point -> min_idx (B,N) int32."""
    xyz = _c32(xyz); node = _c32(node)
    B, _, N = xyz.shape; M = node.shape[2]
    out = np.zeros((B, N), np.int32)
    _lib().orc_som_assign(_fp(xyz), _fp(node), _ip(out), B, N, M)
    return out


def pairwise_min(a, b):
    """models/losses.py:62-66 / 134-141: (min_j ||a_i-b_j||, argmin_j)."""
    a = _c32(a); b = _c32(b)
    B, _, Ma = a.shape; Nb = b.shape[2]
    d = np.zeros((B, Ma), np.float32); arg = np.zeros((B, Ma), np.int32)
    _lib().orc_pairwise_min(_fp(a), _fp(b), _fp(d), _ip(arg), B, Ma, Nb)
    return d, arg


def fps(pts, start, k):
    """FarthestSampler.sample (data/kitti_detector_loader.py:68-83) given the first index: pts (Ns,3) f32 -> idx (k,) i32."""
    pts = np.ascontiguousarray(pts, np.float32)
    out = np.empty(int(k), np.int32)
    _lib().orc_fps(_fp(pts), int(start), _ip(out), int(pts.shape[0]), int(k))
    return out


def nms(kp, sigma, radius):
    """save_keypoints.py:180-216: kp (M,3) f32, sigma (M,) f32 -> original indices of the kept keypoints, emission order."""
    kp = np.ascontiguousarray(kp, np.float32); sigma = np.ascontiguousarray(sigma, np.float32)
    out = np.empty(kp.shape[0], np.int32)
    lib = _lib(); lib.orc_nms.restype = ctypes.c_int
    n = lib.orc_nms(_fp(kp), _fp(sigma), ctypes.c_float(radius), _ip(out), int(kp.shape[0]))
    return out[:n].copy()


def knn(query, db, K):
    """models/layers.py:417-421: topk(K, largest=False, sorted=True) of torch.norm distances."""
    query = _c32(query); db = _c32(db)
    B, _, M = query.shape; N = db.shape[2]
    idx = np.zeros((B, M, K), np.int32); dd = np.zeros((B, M, K), np.float32)
    _lib().orc_knn(_fp(query), _fp(db), _ip(idx), _fp(dd), B, M, N, K)
    return idx, dd


def query_topk(node, x):
    """util/som.py:17-54 (k=1) -> (mask_row_max (B,M) int32, min_idx (B,N) int64).  The one-hot mask
    (B,N,M) itself is not materialised; everything the reference derives from it is reproduced by
    the callers below."""
    min_idx = som_assign(x, node)
    B, M = node.shape[0], node.shape[2]
    row_max = np.zeros((B, M), np.int32)
    for b in range(B):
        row_max[b, np.unique(min_idx[b])] = 1
    return row_max, min_idx.astype(np.int64)


# ----------------------------------------------------------------------------- dense fp building blocks
def conv1x1(x, w, b):
    """nn.Conv1d(k=1) / nn.Conv2d(1x1) (models/layers.py:178,257): x (B,Cin,*S), w (Cout,Cin[,1[,1]])."""
    w2 = w.reshape(w.shape[0], -1)
    sh = x.shape
    xf = x.reshape(sh[0], sh[1], -1)
    y = np.matmul(w2[None], xf) + b.reshape(1, -1, 1)
    return y.reshape((sh[0], w2.shape[0]) + sh[2:])


def batch_norm(x, P, prefix, training, momentum=0.1, eps=1e-5, new_stats=None):
    """F.batch_norm as called at models/layers.py:69-71,119-121.  Training: batch mean / biased var
    for normalisation; running stats updated with the UNBIASED var (torch semantics)."""
    C = x.shape[1]
    axes = (0,) + tuple(range(2, x.ndim))
    g = P[prefix + ".weight"]; be = P[prefix + ".bias"]
    if training:
        xd = x.astype(np.float64)
        mean = xd.mean(axis=axes); var = xd.var(axis=axes)
        n = x.size // C
        if new_stats is not None:
            rm = P[prefix + ".running_mean"]; rv = P[prefix + ".running_var"]
            new_stats[prefix + ".running_mean"] = ((1 - momentum) * rm + momentum * mean).astype(np.float32)
            new_stats[prefix + ".running_var"] = ((1 - momentum) * rv + momentum * var * n / max(n - 1, 1)).astype(np.float32)
        mean = mean.astype(x.dtype); var = var.astype(x.dtype)
    else:
        mean = P[prefix + ".running_mean"].astype(x.dtype); var = P[prefix + ".running_var"].astype(x.dtype)
    shp = (1, C) + (1,) * (x.ndim - 2)
    inv = 1.0 / np.sqrt(var + x.dtype.type(eps))
    return (x - mean.reshape(shp)) * (inv * g.astype(x.dtype)).reshape(shp) + be.astype(x.dtype).reshape(shp)


def layer(x, P, prefix, training, bn=True, relu=True, momentum=0.1, new_stats=None):
    """EquivariantLayer / MyConv2d forward (models/layers.py:289-303, 207-216)."""
    y = conv1x1(x, P[prefix + ".conv.weight"].astype(x.dtype), P[prefix + ".conv.bias"].astype(x.dtype))
    if bn:
        y = batch_norm(y, P, prefix + ".norm", training, momentum, new_stats=new_stats)
    if relu:
        y = np.maximum(y, 0)
    return y


def bn_momentum(opt_momentum, decay_step, decay, epoch):
    """models/layers.py:62-66."""
    if epoch is not None and epoch >= 1 and decay_step is not None and decay_step > 0:
        m = opt_momentum * (decay ** (epoch // decay_step))
        return max(m, 0.01)
    return opt_momentum


def segmented_max(feat, min_idx, M):
    """index_max + gather (models/networks.py:117-120): feat (B,C,N) -> pooled (B,C,M) (empty -> 0),
    plus the arg-max index exactly as index_max returns it."""
    idx = index_max(feat.astype(np.float32), min_idx.astype(np.int32), M)
    pooled = np.take_along_axis(feat, idx.astype(np.int64), axis=2)
    return pooled, idx


def softplus(x):
    """torch.nn.Softplus(beta=1, threshold=20) (models/networks.py:72)."""
    return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, 20))))


def knn_fusion_head(P, coords, feat, K, sigma_lower_bound, kw, dtype=np.float32):
    """GeneralKNNFusionModule on the nodes (models/layers.py:401-440) + head mlp1/2/3 + softplus (models/networks.py:143-154).
    coords (B,3,M): the nodes the kNN runs on / the keypoint offsets are added to; feat (B,C1,M): per-node feature."""
    B, _, M = coords.shape
    knn_i, _ = knn(coords, coords, K)                                       # layers.py:417-421
    ki = knn_i.astype(np.int64).reshape(B, 1, M * K)
    nb_xyz = np.take_along_axis(coords, np.broadcast_to(ki, (B, 3, M * K)), axis=2).reshape(B, 3, M, K)
    nb_feat = np.take_along_axis(feat, np.broadcast_to(ki, (B, feat.shape[1], M * K)), axis=2).reshape(B, -1, M, K)
    g = np.concatenate([nb_xyz - coords[:, :, :, None], nb_feat], axis=1)   # :428-429
    for i in range(3):
        g = layer(g, P, "knnlayer_1.layers_before.%d" % i, **kw)
    gmax = g.max(axis=3, keepdims=True)
    y = np.concatenate([np.broadcast_to(gmax, g.shape), g], axis=1)         # :435 (max first)
    for i in range(2):
        y = layer(y, P, "knnlayer_1.layers_after.%d" % i, **kw)
    knn_feat = y.max(axis=3)                                                # :438
    agg = np.concatenate([feat, knn_feat], axis=1)                          # networks.py:143
    y1 = layer(agg, P, "mlp1", **kw)
    y2 = layer(y1, P, "mlp2", **kw)
    out = layer(y2, P, "mlp3", bn=False, relu=False, **kw)
    keypoints = out[:, 0:3, :] + coords                                     # :151
    sigmas = softplus(out[:, 3, :]) + dtype(sigma_lower_bound)              # :154
    return keypoints, sigmas, knn_i, knn_feat, out


# ----------------------------------------------------------------------------- detector forward
def detector_forward(P, x, sn, node, node_knn_k=16, sigma_lower_bound=1e-3, training=False,
                     momentum=0.1, dtype=np.float32, return_intermediates=False):
    """RPN_Detector.forward (models/networks.py:75-162), k=1.  P: state_dict as numpy arrays.
    x (B,3,N), sn (B,S,N) or None/S=0, node (B,3,M).  Returns dict."""
    x = x.astype(dtype); node = node.astype(dtype)
    B, _, N = x.shape; M = node.shape[2]
    new_stats = {} if training else None
    row_max, min_idx = query_topk(node, x)                                  # networks.py:85-86
    # cluster mean (networks.py:87-97): sum_n mask*x / (count + 1e-5)
    cnt = np.zeros((B, M), np.float64); csum = np.zeros((B, 3, M), np.float64)
    for b in range(B):
        cnt[b] = np.bincount(min_idx[b], minlength=M)
        for c in range(3):
            csum[b, c] = np.bincount(min_idx[b], weights=x[b, c].astype(np.float64), minlength=M)
    cluster_mean = (csum / (cnt[:, None, :] + 1e-5)).astype(dtype)
    centers = np.take_along_axis(cluster_mean, np.broadcast_to(min_idx[:, None, :], (B, 3, N)), axis=2)
    x_dec = x - centers                                                     # networks.py:105-107
    S = 0 if sn is None else sn.shape[1]
    x_aug = np.concatenate([x_dec, sn.astype(dtype)], axis=1) if S >= 1 else x_dec   # :108-114
    kw = dict(training=training, momentum=momentum, new_stats=new_stats)
    h = layer(x_aug, P, "first_pointnet.layers.0", **kw)
    h = layer(h, P, "first_pointnet.layers.1", **kw)
    first = layer(h, P, "first_pointnet.layers.2", bn=False, relu=False, **kw)       # layers.py:531-535
    rm = row_max[:, None, :].astype(dtype)
    pool1, idx1 = segmented_max(first, min_idx, M)
    pool1 = pool1 * rm                                                      # networks.py:117-120
    scattered = np.take_along_axis(pool1, np.broadcast_to(min_idx[:, None, :], (B, pool1.shape[1], N)), axis=2)
    fusion = np.concatenate([first, scattered], axis=1)                     # :123-126
    h = layer(fusion, P, "second_pointnet.layers.0", **kw)
    second = layer(h, P, "second_pointnet.layers.1", bn=False, relu=False, **kw)
    pool2, idx2 = segmented_max(second, min_idx, M)
    pool2 = pool2 * rm                                                      # :130-133
    keypoints, sigmas, knn_i, knn_feat, out = knn_fusion_head(P, cluster_mean, pool2, node_knn_k, sigma_lower_bound, kw, dtype)
    res = dict(node_recomputed=cluster_mean, keypoints=keypoints.astype(np.float32),
               sigmas=sigmas.astype(np.float32), min_idx=min_idx, mask_row_max=row_max,
               new_stats=new_stats)
    if return_intermediates:
        res.update(x_aug=x_aug, first_pn_out=first, idx1=idx1, pool1=pool1, second_pn_out=second, idx2=idx2,
                   pool2=pool2, knn_i=knn_i, knn_feat=knn_feat, head_out=out)
    return res


# ----------------------------------------------------------------------------- ablation detectors
def ablation_forward(P, x, sn, node, mode="knn", node_knn_k=16, sigma_lower_bound=1e-3, training=False, momentum=0.1,
                     dtype=np.float32):
    """RPN_Detector_KNN.forward (models/networks.py:545-608, mode="knn": the 64 points nearest to each node, :556-559) /
    RPN_Detector_Ball.forward (:671-738, mode="ball": ball_query radius 2, k = 64, :681-690).  The nodes are used as given."""
    x = x.astype(dtype); node = node.astype(dtype)
    B, _, N = x.shape; M = node.shape[2]; k = 64
    S = 0 if sn is None else sn.shape[1]
    x_aug = np.concatenate([x, sn.astype(dtype)], axis=1) if S >= 1 else x
    if mode == "knn":
        idx, _ = knn(node, x, k)                                            # any order: only max / BN over k follow
    else:
        idx = ball_query_xyz(x, node, 2.0, k)
    C = x_aug.shape[1]
    gi = np.broadcast_to(idx.astype(np.int64).reshape(B, 1, M * k), (B, C, M * k))
    grp = np.take_along_axis(x_aug, gi, axis=2).reshape(B, C, M, k).copy()
    grp[:, 0:3] -= node[:, :, :, None]                                      # :565 / :692
    new_stats = {} if training else None
    kw = dict(training=training, momentum=momentum, new_stats=new_stats)
    y = layer(grp, P, "conv1", **kw); y = layer(y, P, "conv2", **kw); first = layer(y, P, "conv3", **kw)
    fmax = np.broadcast_to(first.max(axis=3, keepdims=True), first.shape)
    y = layer(np.concatenate([first, fmax], axis=1), P, "conv4", **kw)      # :571 (per-sample first, max last)
    second = layer(y, P, "conv5", **kw)
    second_max = second.max(axis=3)                                         # :572
    keypoints, sigmas, knn_i, knn_feat, out = knn_fusion_head(P, node, second_max, node_knn_k, sigma_lower_bound, kw, dtype)
    return dict(keypoints=keypoints.astype(np.float32), sigmas=sigmas.astype(np.float32), group_idx=idx, new_stats=new_stats)


def ablation_param_shapes(S=4, C1=128, C2=512):
    """RPN_Detector_KNN / RPN_Detector_Ball state_dict layout (models/networks.py:483-543): name -> (shape, has_bn)."""
    h = C1 // 2
    return [("conv1", (h, 3 + S, 1, 1), True), ("conv2", (h, h, 1, 1), True), ("conv3", (h, h, 1, 1), True),
            ("conv4", (C1, C1, 1, 1), True), ("conv5", (C1, C1, 1, 1), True),
            ("knnlayer_1.layers_before.0", (C2 // 2, 3 + C1, 1, 1), True),
            ("knnlayer_1.layers_before.1", (C2 // 2, C2 // 2, 1, 1), True),
            ("knnlayer_1.layers_before.2", (C2 // 2, C2 // 2, 1, 1), True),
            ("knnlayer_1.layers_after.0", (C2, C2, 1, 1), True), ("knnlayer_1.layers_after.1", (C2, C2, 1, 1), True),
            ("mlp1", (512, C1 + C2, 1), True), ("mlp2", (256, 512, 1), True), ("mlp3", (4, 256, 1), False)]


def init_ablation_params(S=4, seed=0, randomize_bn=True):
    return _init_params(ablation_param_shapes(S), seed, randomize_bn)


# ----------------------------------------------------------------------------- losses
def chamfer_prob(src, dst, sig_src, sig_dst):
    """ChamferLoss_Brute.forward, sigma branch (models/losses.py:50-99) -> (loss, pure, weighted)."""
    d_sd, i_sd = pairwise_min(src, dst)
    d_ds, i_ds = pairwise_min(dst, src)
    d_sd = d_sd.astype(np.float64); d_ds = d_ds.astype(np.float64)
    s_sd = (sig_src.astype(np.float64) + np.take_along_axis(sig_dst.astype(np.float64), i_sd.astype(np.int64), 1)) / 2
    s_ds = (sig_dst.astype(np.float64) + np.take_along_axis(sig_src.astype(np.float64), i_ds.astype(np.int64), 1)) / 2
    fwd = (np.log(s_sd) + d_sd / s_sd).mean()
    bwd = (np.log(s_ds) + d_ds / s_ds).mean()
    pure = d_sd.mean() + d_ds.mean()
    w_sd = (1 / s_sd) / (1 / s_sd).mean(); w_ds = (1 / s_ds) / (1 / s_ds).mean()
    weighted = (w_sd * d_sd).mean() + (w_ds * d_ds).mean()
    return np.float32(fwd + bwd), np.float32(pure), np.float32(weighted)


def point_on_surface(kp, pc, sn, g=None):
    """PointOnSurfaceLoss (models/losses.py:146-183): nearest cloud point per keypoint, then the squared cosine between its
    normal (first three channels of sn) and the unit vector from it to the keypoint.  Returns loss (B,M) and, if an
    upstream gradient g (B,M) is given, d(sum g*loss)/d kp (B,3,M) -- the arg-min is piecewise constant."""
    kp = np.asarray(kp, np.float32); pc = np.asarray(pc, np.float32); sn = np.asarray(sn, np.float32)
    _, arg = pairwise_min(kp, pc)
    B, _, M = kp.shape
    p = np.stack([pc[b][:, arg[b]] for b in range(B)]); n = np.stack([sn[b][:3, arg[b]] for b in range(B)])
    d = (kp - p).astype(np.float64); n = n.astype(np.float64)
    r = np.sqrt((d ** 2).sum(1)); e = r + 1e-7
    s = (n * d).sum(1) / e
    loss = (s ** 2).astype(np.float32)
    if g is None:
        return loss
    nd = (n * d).sum(1)
    c = np.where(r > 0, nd / np.maximum(r, 1e-300) / e ** 2, 0.0)
    grad = (2 * s * np.asarray(g, np.float64))[:, None, :] * (n / e[:, None, :] - d * c[:, None, :])
    return loss, grad.astype(np.float32)


def single_side_chamfer(kp, pc):
    """SingleSideChamferLoss_Brute.forward (models/losses.py:125-143) -> (B,M)."""
    d, _ = pairwise_min(kp, pc)
    return d


def transform_keypoints(kp, R, scale, shift):
    """models/keypoint_detector.py:182-184: R @ kp * scale + shift."""
    out = np.matmul(R.astype(np.float32), kp.astype(np.float32))
    out = out * scale.reshape(-1, 1, 1).astype(np.float32)
    return out + shift.astype(np.float32)


def detector_fwd_loss(P, src_pc, src_sn, src_node, dst_pc, dst_sn, dst_node, R, scale, shift,
                      node_knn_k=16, sigma_lower_bound=1e-3, alpha=0.01, training=False, dtype=np.float32):
    """ModelDetector.test_model / the forward half of .optimize (models/keypoint_detector.py:158-241):
    siamese forward on cat(src,dst), transform, chamfer + 2x keypoint-on-pc (point_to_point)."""
    B = src_pc.shape[0]
    f = detector_forward(P, np.concatenate([src_pc, dst_pc]), np.concatenate([src_sn, dst_sn]),
                         np.concatenate([src_node, dst_node]), node_knn_k, sigma_lower_bound, training, dtype=dtype)
    kp_s, kp_d = f["keypoints"][:B], f["keypoints"][B:]
    sg_s, sg_d = f["sigmas"][:B], f["sigmas"][B:]
    kp_t = transform_keypoints(kp_s, R, scale, shift)
    lc, pure, weighted = chamfer_prob(kp_t, kp_d, sg_s, sg_d)
    ls = np.float32(single_side_chamfer(kp_s, src_pc).astype(np.float64).mean() * alpha)
    ld = np.float32(single_side_chamfer(kp_d, dst_pc).astype(np.float64).mean() * alpha)
    return dict(loss=np.float32(lc + ls + ld), loss_chamfer=lc, chamfer_pure=pure, chamfer_weighted=weighted,
                loss_keypoint_on_pc_src=ls, loss_keypoint_on_pc_dst=ld,
                src_keypoints=kp_s, dst_keypoints=kp_d, src_sigmas=sg_s, dst_sigmas=sg_d,
                src_keypoints_transformed=kp_t, node_recomputed=f["node_recomputed"], new_stats=f["new_stats"])


def desc_pair_scan_loss(anc, pos, neg, anc_sigmas, gamma=0.5, sigma_max=3.0):
    """DescPairScanLoss.forward (models/losses.py:200-237) -> (loss (B,M), active_percentage (B))."""
    a = anc.astype(np.float64); p = pos.astype(np.float64); n = neg.astype(np.float64)
    dpos = np.sqrt(((a[:, :, :, None] - p[:, :, None, :]) ** 2).sum(1)).min(2)
    dneg = np.sqrt(((a[:, :, :, None] - n[:, :, None, :]) ** 2).sum(1)).min(2)
    before = dpos - dneg + gamma
    active = (before > 0).mean(1)
    w = np.maximum(sigma_max - anc_sigmas.astype(np.float64), 0)
    w = w / w.mean(1, keepdims=True)
    return (w * np.maximum(before, 0)).astype(np.float32), active.astype(np.float32)


# ----------------------------------------------------------------------------- descriptor forward
def descriptor_forward(P, x, sn, keypoints, radius=1.0, nsamples=64, training=False, momentum=0.1,
                       permute_idx=None, dtype=np.float32):
    """DescriptorLiteOld.forward (models/networks.py:333-385).  `permute_idx` is the host-side
    np.random.permutation(N) of :345 (pass the same one to both sides for seeded parity)."""
    x = x.astype(dtype); keypoints = keypoints.astype(dtype)
    B, _, N = x.shape; M = keypoints.shape[2]; K = nsamples
    if permute_idx is not None:
        x = x[:, :, permute_idx]; sn = None if sn is None else sn[:, :, permute_idx]
    x_aug = x if (sn is None or sn.shape[1] == 0) else np.concatenate([x, sn.astype(dtype)], axis=1)
    idx = ball_query_xyz(x, keypoints, radius, K)                           # :355-359
    C = x_aug.shape[1]
    gi = np.broadcast_to(idx.astype(np.int64).reshape(B, 1, M * K), (B, C, M * K))
    ball = np.take_along_axis(x_aug, gi, axis=2).reshape(B, C, M, K).copy()  # :361-362
    ball[:, 0:3] -= keypoints[:, :, :, None]                                # :373
    new_stats = {} if training else None
    kw = dict(training=training, momentum=momentum, new_stats=new_stats)
    y = layer(ball, P, "conv1", **kw); y = layer(y, P, "conv2", **kw); y_first = layer(y, P, "conv3", **kw)
    ymax = np.broadcast_to(y_first.max(axis=3, keepdims=True), y_first.shape)
    y2 = layer(np.concatenate([y_first, ymax], axis=1), P, "conv4", **kw)   # :380 (per-sample first)
    y2 = layer(y2, P, "conv5", bn=False, relu=False, **kw)
    desc = y2.max(axis=3)
    desc = desc / (np.sqrt((desc.astype(np.float64) ** 2).sum(axis=1, keepdims=True)).astype(dtype) + dtype(1e-5))  # :383
    return dict(descriptor=desc.astype(np.float32), x_features=ball.astype(np.float32), ball_idx=idx,
                new_stats=new_stats)


# ----------------------------------------------------------------------------- synthetic data
def farthest_point_sample(pts, M, rng):
    """Numpy FPS as in data/kitti_detector_loader.py:69-83 (FarthestSampler): start from a random point."""
    N = pts.shape[1]
    far = np.zeros(M, np.int64); dist = np.full(N, 1e10)
    cur = int(rng.integers(N))
    for i in range(M):
        far[i] = cur
        d = ((pts - pts[:, cur:cur + 1]) ** 2).sum(axis=0)
        dist = np.minimum(dist, d)
        cur = int(dist.argmax())
    return pts[:, far]


def synth_pair(B, N, M, S, kind="lidar", seed=1234, node_subset=3):
    """Deterministic synthetic detector batch per SURVEY.md 8(d): S1 'object' U(-1,1)^3 / S2 'lidar'
    x,z~U(-40,40), y~U(-2,2); nodes = FPS of M from a random N/node_subset subset; dst = R*src*1+shift."""
    rng = np.random.default_rng(seed)
    if kind == "lidar":
        src = np.stack([rng.uniform(-40, 40, (B, N)), rng.uniform(-2, 2, (B, N)), rng.uniform(-40, 40, (B, N))], 1)
    else:
        src = rng.uniform(-1, 1, (B, 3, N))
    nrm = rng.normal(size=(B, 3, N)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    sn = nrm if S == 3 else (np.concatenate([nrm, rng.uniform(0, 1, (B, 1, N))], 1) if S == 4 else np.zeros((B, 0, N)))
    R = np.zeros((B, 3, 3)); shift = rng.uniform(-0.5, 0.5, (B, 3, 1)); scale = np.ones((B,))
    for b in range(B):
        a = rng.uniform(0, 2 * np.pi)
        if kind == "lidar":
            R[b] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
        else:
            q = rng.normal(size=4); q /= np.linalg.norm(q); w, x, y, z = q
            R[b] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                    [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                    [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
    dst = np.matmul(R, src) * scale[:, None, None] + shift
    dsn = sn.copy()
    if S >= 3:
        dsn[:, 0:3] = np.matmul(R, sn[:, 0:3])
    src_node = np.zeros((B, 3, M)); dst_node = np.zeros((B, 3, M))
    for b in range(B):
        sub = rng.choice(N, N // node_subset, replace=False)
        src_node[b] = farthest_point_sample(src[b][:, sub], M, rng)
        sub = rng.choice(N, N // node_subset, replace=False)
        dst_node[b] = farthest_point_sample(dst[b][:, sub], M, rng)
    f = np.float32
    return dict(src_pc=src.astype(f), src_sn=sn.astype(f), src_node=src_node.astype(f),
                dst_pc=dst.astype(f), dst_sn=dsn.astype(f), dst_node=dst_node.astype(f),
                R=R.astype(f), scale=scale.astype(f), shift=shift.astype(f))


def detector_param_shapes(S=4, C1=128, C2=512):
    """RPN_Detector state_dict layout: name -> (shape, has_bn)."""
    h = C1 // 2
    spec = [("first_pointnet.layers.0", (h, 3 + S, 1), True), ("first_pointnet.layers.1", (h, h, 1), True),
            ("first_pointnet.layers.2", (h, h, 1), False),
            ("second_pointnet.layers.0", (C1, C1, 1), True), ("second_pointnet.layers.1", (C1, C1, 1), False),
            ("knnlayer_1.layers_before.0", (C2 // 2, 3 + C1, 1, 1), True),
            ("knnlayer_1.layers_before.1", (C2 // 2, C2 // 2, 1, 1), True),
            ("knnlayer_1.layers_before.2", (C2 // 2, C2 // 2, 1, 1), True),
            ("knnlayer_1.layers_after.0", (C2, C2, 1, 1), True), ("knnlayer_1.layers_after.1", (C2, C2, 1, 1), True),
            ("mlp1", (512, C1 + C2, 1), True), ("mlp2", (256, 512, 1), True), ("mlp3", (4, 256, 1), False)]
    return spec


def init_detector_params(S=4, seed=0, C1=128, C2=512, randomize_bn=False):
    """Random init following models/layers.py:196-205,278-287 and networks.py:70-71 (numpy RNG, so
    not bit-identical to torch's init -- used for synthetic benchmarks/tests where both sides get
    the same arrays)."""
    return _init_params(detector_param_shapes(S, C1, C2), seed, randomize_bn)


def _init_params(spec, seed, randomize_bn):
    rng = np.random.default_rng(seed)
    P = {}
    for name, shp, bn in spec:
        fan_in = int(np.prod(shp[1:]))
        std = 1e-4 if name == "mlp3" else np.sqrt(2.0 / fan_in)
        P[name + ".conv.weight"] = rng.normal(0, std, shp).astype(np.float32)
        P[name + ".conv.bias"] = np.zeros(shp[0], np.float32)
        if bn:
            C = shp[0]
            P[name + ".norm.weight"] = (rng.uniform(0.5, 1.5, C) if randomize_bn else np.ones(C)).astype(np.float32)
            P[name + ".norm.bias"] = (rng.normal(0, 0.1, C) if randomize_bn else np.zeros(C)).astype(np.float32)
            P[name + ".norm.running_mean"] = (rng.normal(0, 0.1, C) if randomize_bn else np.zeros(C)).astype(np.float32)
            P[name + ".norm.running_var"] = (rng.uniform(0.5, 1.5, C) if randomize_bn else np.ones(C)).astype(np.float32)
            P[name + ".norm.num_batches_tracked"] = np.zeros((), np.int64)
    if randomize_bn:
        for name, shp, bn in spec:
            P[name + ".conv.bias"] = rng.normal(0, 0.05, shp[0]).astype(np.float32)
    return P


def desc_train_inputs(B, N, M, S, seed):
    """Deterministic inputs of one descriptor train step: an anchor scan, its rigidly moved + jittered positive, keypoints
    near cloud points in both frames, sigmas in [0, 4), a cyclic negative index.  Dense enough (x0.2 in x/z) that balls of
    radius 1 hold 0 .. > K points."""
    d = synth_pair(B, N, 16, S, kind="lidar", seed=seed)
    rng = np.random.default_rng(seed + 1)
    sc = np.array([0.2, 1.0, 0.2], np.float32).reshape(1, 3, 1)
    anc_pc = (d["src_pc"] * sc).astype(np.float32)
    pos_pc = (d["dst_pc"] * sc).astype(np.float32)
    sel = np.stack([rng.choice(N, M, replace=False) for _ in range(B)])
    anc_kp = np.stack([anc_pc[b][:, sel[b]] for b in range(B)]) + rng.normal(0, 0.05, (B, 3, M))
    pos_kp = np.stack([pos_pc[b][:, sel[b]] for b in range(B)]) + rng.normal(0, 0.05, (B, 3, M))
    return dict(anc_pc=anc_pc, anc_sn=d["src_sn"], anc_kp=anc_kp.astype(np.float32),
                anc_sigma=rng.uniform(0, 4, (B, M)).astype(np.float32),
                pos_pc=pos_pc, pos_sn=d["dst_sn"], pos_kp=pos_kp.astype(np.float32),
                pos_sigma=rng.uniform(0, 4, (B, M)).astype(np.float32),
                neg_idx=((np.arange(B) + 1) % B).astype(np.int64))


def ablation_inputs(seed, B=2, N=4096, M=128, S=4):
    """Inputs of the ablation-detector fixtures (regenerated from the seed by the consumers): a LiDAR-like cloud shrunk in
    x/z so that balls of radius 2 hold a few .. more than 64 points, FPS nodes, fixed loss weights."""
    d = synth_pair(B, N, M, S, kind="lidar", seed=seed)
    sc = np.array([0.5, 1.0, 0.5], np.float32).reshape(1, 3, 1)
    rng = np.random.default_rng(seed + 7)
    return dict(pc=(d["src_pc"] * sc).astype(np.float32), sn=d["src_sn"], node=(d["src_node"] * sc).astype(np.float32),
                w_kp=rng.normal(size=(B, 3, M)).astype(np.float32), w_sig=rng.normal(size=(B, M)).astype(np.float32))
