"""-m gpu parity of the fused detector plan (ModelDetector API) against the reference goldens (small shapes,
outputs of the real reference) and against the numpy oracle at larger shapes."""
import numpy as np
import pytest
import torch

from oracle import usip_oracle as orc
from tests.util_gpu import cu, dev, golden, load_params, make_opt, rel_err

pytestmark = pytest.mark.gpu
REL = 1e-4


def _setup(fname, use_tc):
    from usip_b200.models.keypoint_detector import ModelDetector
    g = golden(fname)
    B, N, M, S, Kn, seed = [int(v) for v in g["cfg"]]
    kind, scene = str(g["kind"]), str(g["scene"])
    d = orc.synth_pair(B, N, M, S, kind=kind, seed=seed)
    C1, C2 = (64, 256) if scene == "indoor" else (128, 512)
    P = orc.init_detector_params(S=S, seed=seed, C1=C1, C2=C2, randomize_bn=True)
    P["mlp3.conv.weight"] = (P["mlp3.conv.weight"] * 1000).astype(np.float32)
    opt = make_opt(batch_size=B, input_pc_num=N, node_num=M, surface_normal_len=S, node_knn_k_1=Kn, scene=scene,
                   loss_sigma_lower_bound=float(g["lb"]), keypoint_on_pc_alpha=float(g["alpha"]), use_tensor_cores=use_tc)
    md = ModelDetector(opt)
    load_params(md.detector, P)
    md.set_input(*[torch.from_numpy(d[k]) for k in ("src_pc", "src_sn", "src_node", "dst_pc", "dst_sn", "dst_node",
                                                    "R", "scale", "shift")])
    return g, d, P, md


def _loss_vec(md):
    return np.array([md.loss.item(), md.loss_chamfer.item(), md.chamfer_pure.item(), md.chamfer_weighted.item(),
                     md.loss_keypoint_on_pc_src.item(), md.loss_keypoint_on_pc_dst.item()], np.float32)


@pytest.mark.parametrize("use_tc", [False, True])
@pytest.mark.parametrize("fname", ["detector_kitti_small.npz", "detector_modelnet_small.npz", "detector_lite_small.npz"])
def test_detector_eval_and_train_forward_vs_reference_golden(fname, use_tc):
    g, d, P, md = _setup(fname, use_tc)
    md.test_model()                                                       # eval-mode BN (running stats)
    kp = torch.cat([md.src_keypoints, md.dst_keypoints]).cpu().numpy()
    sig = torch.cat([md.src_sigmas, md.dst_sigmas]).cpu().numpy()
    node = torch.cat([md.src_node_recomputed, md.dst_node_recomputed]).cpu().numpy()
    assert rel_err(node, g["eval_node"]) < 1e-5
    assert rel_err(kp, g["eval_kp"]) < REL, rel_err(kp, g["eval_kp"])
    assert rel_err(sig, g["eval_sig"]) < REL, rel_err(sig, g["eval_sig"])
    lv = _loss_vec(md)
    assert np.all(np.abs(lv - g["eval_loss"]) <= REL * np.abs(g["eval_loss"]) + 1e-6), (lv, g["eval_loss"])

    md.forward_loss(epoch=0, train_bn=True)                               # train-mode BN forward + loss
    kp = torch.cat([md.src_keypoints, md.dst_keypoints]).cpu().numpy()
    sig = torch.cat([md.src_sigmas, md.dst_sigmas]).cpu().numpy()
    assert rel_err(kp, g["train_kp"]) < REL, rel_err(kp, g["train_kp"])
    assert rel_err(sig, g["train_sig"]) < REL, rel_err(sig, g["train_sig"])
    lv = _loss_vec(md)
    assert np.all(np.abs(lv - g["train_loss"]) <= REL * np.abs(g["train_loss"]) + 1e-6), (lv, g["train_loss"])
    # running statistics were updated exactly once (momentum 0.1, unbiased variance)
    sd = md.detector.state_dict()
    for k in sd:
        if k.endswith("running_mean") or k.endswith("running_var"):
            flat = sd[k].cpu().numpy().reshape(-1)
            ref = g["after/" + k]
            assert np.allclose(flat[:24], ref[3:3 + min(24, flat.size)], rtol=2e-4, atol=1e-6), k


@pytest.mark.parametrize("use_tc", [False, True])
def test_detector_vs_oracle_medium(use_tc):
    """B'=4 clouds, N=8192, M=256, Kn=16: fused CUDA plan vs the numpy oracle (fp32), train-mode BN."""
    from usip_b200.models.keypoint_detector import ModelDetector
    B, N, M, S, Kn = 2, 8192, 256, 4, 16
    d = orc.synth_pair(B, N, M, S, kind="lidar", seed=77)
    P = orc.init_detector_params(S=S, seed=5, randomize_bn=True)
    P["mlp3.conv.weight"] = (P["mlp3.conv.weight"] * 1000).astype(np.float32)
    opt = make_opt(batch_size=B, input_pc_num=N, node_num=M, surface_normal_len=S, node_knn_k_1=Kn, use_tensor_cores=use_tc)
    md = ModelDetector(opt)
    load_params(md.detector, P)
    md.set_input(*[torch.from_numpy(d[k]) for k in ("src_pc", "src_sn", "src_node", "dst_pc", "dst_sn", "dst_node",
                                                    "R", "scale", "shift")])
    md.forward_loss(train_bn=True)
    r = orc.detector_fwd_loss(P, d["src_pc"], d["src_sn"], d["src_node"], d["dst_pc"], d["dst_sn"], d["dst_node"],
                              d["R"], d["scale"], d["shift"], node_knn_k=Kn, sigma_lower_bound=opt.loss_sigma_lower_bound,
                              alpha=opt.keypoint_on_pc_alpha, training=True)
    kp = torch.cat([md.src_keypoints, md.dst_keypoints]).cpu().numpy()
    sig = torch.cat([md.src_sigmas, md.dst_sigmas]).cpu().numpy()
    assert rel_err(kp, np.concatenate([r["src_keypoints"], r["dst_keypoints"]])) < REL
    assert rel_err(sig, np.concatenate([r["src_sigmas"], r["dst_sigmas"]])) < REL
    assert abs(md.loss.item() - r["loss"]) <= REL * abs(r["loss"])
    assert abs(md.loss_chamfer.item() - r["loss_chamfer"]) <= REL * abs(r["loss_chamfer"])


@pytest.mark.parametrize("use_tc", [False, True])
@pytest.mark.parametrize("fname", ["detector_kitti_small.npz", "detector_modelnet_small.npz", "detector_lite_small.npz"])
def test_detector_train_step_vs_reference_golden(fname, use_tc):
    """One full ModelDetector.optimize() (train-mode forward, loss, backward, Adam) against the reference's own
    gradients and post-step parameters (tools/make_golden.py ran test_model() then optimize(epoch=0))."""
    g, d, P, md = _setup(fname, use_tc)
    md.test_model()
    md.optimize(epoch=0)
    torch.cuda.synchronize()
    lv = _loss_vec(md)
    assert np.all(np.abs(lv - g["train_loss"]) <= REL * np.abs(g["train_loss"]) + 1e-6), (lv, g["train_loss"])
    worst = 0.0
    for k, p in md.detector.named_parameters():
        ref = g["grad/" + k]
        gr = p.grad.detach().cpu().numpy().reshape(-1).astype(np.float64)
        absmax, norm = float(ref[2]), float(ref[3])
        if k.endswith("conv.bias") and norm < 1e-4:
            assert np.linalg.norm(gr) < 1e-4, k          # bias in front of BN: true gradient is 0 (reference: ~1e-6 fp noise)
            continue
        e_norm = abs(np.linalg.norm(gr) - norm) / max(norm, 1e-12)
        e_el = np.abs(gr[:24] - ref[4:4 + min(24, gr.size)]).max() / max(absmax, 1e-12)
        worst = max(worst, e_norm, e_el)
        assert e_norm < 2e-3 and e_el < 5e-3, (k, e_norm, e_el)
    # parameters after the Adam step: the first Adam update is lr*g/(|g|+eps) ~ lr*sign(g); entries whose reference
    # gradient is above the noise floor must land on the same value, noise-level entries (|g| ~ eps = 1e-8, e.g. dead
    # ReLU channels) may differ by at most 2*lr
    sd = md.detector.state_dict()
    for k, v in sd.items():
        if k.endswith("num_batches_tracked") or k.endswith("running_mean") or k.endswith("running_var"):
            continue
        ref = g["after/" + k]
        gref = g["grad/" + k][4:]
        flat = v.cpu().numpy().reshape(-1)
        n = min(24, flat.size)
        diff = np.abs(flat[:n] - ref[3:3 + n])
        assert diff.max() <= 2.2e-3, (k, diff.max())
        if k.endswith("conv.bias") and float(g["grad/" + k][3]) < 1e-3 * max(float(g["grad/" + k.replace("bias", "weight")][3]), 1e-12):
            continue      # bias feeding a train-mode BN: true gradient 0, reference has rounding noise
        solid = np.abs(gref[:n]) > max(1e-5, 1e-2 * float(g["grad/" + k][2]))     # well above our 5e-3*absmax gradient tolerance
        assert np.all(diff[solid] <= 2e-5 + 2e-4 * np.abs(ref[3:3 + n][solid])), (k, diff, gref[:n])
    print("worst gradient rel err", worst)


def test_descriptor_forward_vs_reference_golden():
    """DescriptorLiteOld (ball query + gather + MLP + max + L2 normalise) vs the reference golden, eval and train BN."""
    from usip_b200.models import networks
    g = golden("descriptor.npz")
    B, N, M, S, K, seed = [int(v) for v in g["cfg"]]
    opt = make_opt(batch_size=B, input_pc_num=N, node_num=M, surface_normal_len=S, ball_radius=1.0, ball_nsamples=K,
                   descriptor_len=128)
    P = {k[len("param/"):]: g[k] for k in g.files if k.startswith("param/")}
    for use_tc in (False, True):
        opt.use_tensor_cores = use_tc
        net = networks.DescriptorLiteOld(opt).to(dev())
        load_params(net, P)
        for mode in ("eval", "train"):
            net.train(mode == "train")
            np.random.seed(seed)                                   # forward draws np.random.permutation(N) (networks.py:345)
            with torch.no_grad():
                desc, feats = net(cu(g["pc"]), cu(g["sn"]), cu(g["kp"]), mode == "train", None)
            assert np.array_equal(feats.cpu().numpy(), g[mode + "_feats"]), (mode, "x_features must be bit-exact")
            e = rel_err(desc.cpu().numpy(), g[mode + "_desc"])
            assert e < REL, (mode, use_tc, e)


@pytest.mark.parametrize("use_tc", [False, True])
def test_descriptor_train_step_vs_reference_golden(use_tc):
    """One full ModelDescriptor.optimize() (train-mode forward of the siamese batch, DescPairScanLoss, backward, Adam)
    against the reference's own loss, descriptors, gradients and post-step parameters (tools/make_golden.py)."""
    from usip_b200.models.keypoint_descriptor import ModelDescriptor
    g = golden("descriptor_train.npz")
    B, N, M, S, K, seed = [int(v) for v in g["cfg"]]
    d = orc.desc_train_inputs(B, N, M, S, seed)
    opt = make_opt(batch_size=B, input_pc_num=N, node_num=M, surface_normal_len=S, ball_radius=1.0, ball_nsamples=K,
                   descriptor_len=128, random_pc_dropout_lower_limit=1.0, use_tensor_cores=use_tc)
    md = ModelDescriptor(opt)
    load_params(md.descriptor, {k[len("param/"):]: g[k] for k in g.files if k.startswith("param/")})
    md.set_input(*[torch.from_numpy(d[k]) for k in ("anc_pc", "anc_sn", "anc_kp", "anc_sigma", "pos_pc", "pos_sn", "pos_kp", "pos_sigma")],
                 torch.from_numpy(d["neg_idx"]))
    np.random.seed(seed)                                       # the forward draws np.random.permutation(N)
    md.optimize(epoch=0)
    torch.cuda.synchronize()
    assert abs(md.loss.item() - float(g["loss"])) <= REL * abs(float(g["loss"])), (md.loss.item(), float(g["loss"]))
    assert abs(md.active_percentage.item() - float(g["active"])) < 1e-6
    assert rel_err(md.anc_descriptors.detach().cpu().numpy(), g["anc_desc"]) < REL
    assert rel_err(md.pos_descriptors.detach().cpu().numpy(), g["pos_desc"]) < REL
    worst = 0.0
    for k, p in md.descriptor.named_parameters():
        ref = g["grad/" + k]
        gr = p.grad.detach().cpu().numpy().reshape(-1).astype(np.float64)
        absmax, norm = float(ref[2]), float(ref[3])
        wnorm = float(g["grad/" + k.replace("bias", "weight")][3]) if k.endswith("conv.bias") else norm
        if k.endswith("conv.bias") and norm < 1e-3 * max(wnorm, 1e-12):
            assert np.linalg.norm(gr) < 1e-3 * max(wnorm, 1e-12), k    # bias in front of BN: true gradient is 0
            continue
        e_norm = abs(np.linalg.norm(gr) - norm) / max(norm, 1e-12)
        e_el = np.abs(gr[:24] - ref[4:4 + min(24, gr.size)]).max() / max(absmax, 1e-12)
        worst = max(worst, e_norm, e_el)
        assert e_norm < 2e-3 and e_el < 5e-3, (k, e_norm, e_el)
    sd = md.descriptor.state_dict()
    for k, v in sd.items():
        if k.endswith("num_batches_tracked") or k.endswith("running_mean") or k.endswith("running_var"):
            continue
        ref = g["after/" + k]
        gref = g["grad/" + k][4:]
        flat = v.cpu().numpy().reshape(-1)
        n = min(24, flat.size)
        diff = np.abs(flat[:n] - ref[3:3 + n])
        assert diff.max() <= 2.2e-3, (k, diff.max())
        if k.endswith("conv.bias") and float(g["grad/" + k][3]) < 1e-3 * max(float(g["grad/" + k.replace("bias", "weight")][3]), 1e-12):
            continue
        solid = np.abs(gref[:n]) > max(1e-5, 1e-2 * float(g["grad/" + k][2]))
        assert np.all(diff[solid] <= 2e-5 + 2e-4 * np.abs(ref[3:3 + n][solid])), (k, diff, gref[:n])
    print("worst gradient rel err", worst)


def test_forward_loss_cuda_graph_matches_eager():
    """ModelDetector.forward_loss(graph=True) replays the captured launch sequence: results must equal the eager run
    bit for bit (all kernels on the fwd+loss path are deterministic), also after the inputs change."""
    g, d, P, md = _setup("detector_kitti_small.npz", True)
    keys = ("src_pc", "src_sn", "src_node", "dst_pc", "dst_sn", "dst_node", "R", "scale", "shift")
    d2 = orc.synth_pair(2, 2048, 64, 4, kind="lidar", seed=999)
    for data in (d, d2, d):
        md.set_input(*[torch.from_numpy(data[k]) for k in keys])
        sd0 = {k: v.clone() for k, v in md.detector.state_dict().items()}
        md.forward_loss(epoch=0, train_bn=True, graph=False)
        eager = (md.loss.item(), torch.cat([md.src_keypoints, md.dst_keypoints]).clone(), torch.cat([md.src_sigmas, md.dst_sigmas]).clone())
        md.detector.load_state_dict(sd0)                       # undo the running-stat update, same starting state
        md.set_input(*[torch.from_numpy(data[k]) for k in keys])
        md.forward_loss(epoch=0, train_bn=True, graph=True)
        assert md.loss.item() == eager[0]
        assert torch.equal(torch.cat([md.src_keypoints, md.dst_keypoints]), eager[1])
        assert torch.equal(torch.cat([md.src_sigmas, md.dst_sigmas]), eager[2])


def test_train_steps_graph_prepack_and_eager_agree():
    """Four optimize() steps three ways -- (a) eager with every layer packing its own weights, (b) eager with the one-launch
    re-pack of all stale weight matrices (engine.prepack_weights / usip_layer_tc_pack_many), (c) the captured CUDA graph.
    The kernels and their order are identical, only who launches the weight packing differs, so the LOSS of every step must
    agree (a forward on weights that are one Adam step old would move it by ~1e-2; run-to-run noise of the atomics is ~1e-6).
    Parameters are not compared element-wise: Adam turns the summation noise of near-zero gradients into +-lr steps."""
    from usip_b200 import engine
    keys = ("src_pc", "src_sn", "src_node", "dst_pc", "dst_sn", "dst_node", "R", "scale", "shift")
    batches = [orc.synth_pair(2, 2048, 64, 4, kind="lidar", seed=700 + i) for i in range(4)]
    losses, packed_jobs = {}, []
    for mode in ("inline", "prepack", "graph"):
        g, d, P, md = _setup("detector_kitti_small.npz", True)
        md.use_cuda_graph = mode == "graph"
        orig = engine.prepack_weights
        if mode == "inline":
            engine.prepack_weights = lambda module: 0
        else:
            def counted(module, _orig=orig, _mode=mode):
                n = _orig(module); packed_jobs.append((_mode, n)); return n
            engine.prepack_weights = counted
        try:
            ls = []
            for data in batches:
                md.set_input(*[torch.from_numpy(data[k]) for k in keys])
                md.optimize(epoch=0)
                ls.append(float(md.loss))
        finally:
            engine.prepack_weights = orig
        losses[mode] = ls
    assert max(n for m, n in packed_jobs if m == "prepack") >= 10     # the batched path really re-packed the layers
    for mode in ("prepack", "graph"):
        for k, (x, y) in enumerate(zip(losses["inline"], losses[mode])):
            assert abs(x - y) <= 5e-4 * abs(x), (mode, k, x, y)
    assert abs(losses["inline"][0] - losses["inline"][-1]) > 1e-2 * abs(losses["inline"][0])   # the steps do move the loss


def test_prepack_weights_fills_every_workspace_with_the_current_weights():
    """After an optimizer step every cached tensor-core workspace is stale; engine.prepack_weights() must rebuild ALL of them
    from the CURRENT weights in one launch.  Checked against a one-layer-at-a-time usip_layer_tc_pack_many into scratch
    buffers (the pack arithmetic itself is the device function the per-layer kernel uses)."""
    import ctypes
    from usip_b200 import _lib, engine, ops
    g, d, P, md = _setup("detector_kitti_small.npz", True)
    md.use_cuda_graph = False
    md.optimize(epoch=0)                                   # registers the layers; Adam then moves the weights
    n = engine.prepack_weights(md.detector)
    assert n >= 10
    lib = _lib.load()
    checked = 0
    for p in md.detector.parameters():
        for ent in p.__dict__.get("_usip_tc", {}).values():
            if ent.desc is None:
                continue
            assert ent.ver == (p._version, _lib.WEIGHT_GEN[0])                       # marked fresh
            d2 = ops.LayerDesc.from_buffer_copy(ent.desc)
            scratch = torch.zeros_like(ent.ws)
            d2.tc_workspace = scratch.data_ptr()
            _lib.check(lib.usip_layer_tc_pack_many(ctypes.byref(d2), 1, ops._stream()), "usip_layer_tc_pack_many")
            assert torch.equal(scratch.view(torch.int32), ent.ws.view(torch.int32))
            assert int((scratch.view(torch.int32) != 0).sum()) > 0
            checked += 1
    assert checked == n
    assert engine.prepack_weights(md.detector) == 0        # nothing is stale any more


def test_prefetch_input_matches_set_input():
    """prefetch_input() stages the next batch on a copy stream; set_input() with the same tensors adopts it, with other
    tensors it copies as usual.  The loss must be the one of the batch passed to set_input either way."""
    g, d, P, md = _setup("detector_kitti_small.npz", True)
    keys = ("src_pc", "src_sn", "src_node", "dst_pc", "dst_sn", "dst_node", "R", "scale", "shift")
    d2 = orc.synth_pair(2, 2048, 64, 4, kind="lidar", seed=999)
    a = [torch.from_numpy(d[k]).pin_memory() for k in keys]
    b = [torch.from_numpy(d2[k]).pin_memory() for k in keys]
    sd0 = {k: v.clone() for k, v in md.detector.state_dict().items()}
    ref = []
    for batch in (a, b):
        md.detector.load_state_dict(sd0)
        md.set_input(*batch); md.forward_loss(epoch=0, train_bn=True, graph=False); ref.append(md.loss.item())
    assert ref[0] != ref[1]
    md.detector.load_state_dict(sd0)
    md.prefetch_input(*a); md.set_input(*a)                    # adopted
    assert md._staged is None
    md.prefetch_input(*b)                                      # staged while step a runs
    md.forward_loss(epoch=0, train_bn=True, graph=False); assert md.loss.item() == ref[0]
    md.detector.load_state_dict(sd0)
    md.set_input(*b); md.forward_loss(epoch=0, train_bn=True, graph=False); assert md.loss.item() == ref[1]
    md.detector.load_state_dict(sd0)
    md.prefetch_input(*b); md.set_input(*a)                    # different tensors: plain copy, the stale stage is dropped later
    md.forward_loss(epoch=0, train_bn=True, graph=False); assert md.loss.item() == ref[0]


@pytest.mark.parametrize("use_tc", [False, True])
def test_detector_ragged_sizes_vs_oracle(use_tc):
    """Sizes that are not multiples of any tile: N=3001 points (what random point dropout produces), M=50 nodes, B'=6 clouds,
    plus one node moved far away so that a cluster is empty.  Train-mode BN, fused CUDA plan vs the numpy oracle."""
    from usip_b200.models.keypoint_detector import ModelDetector
    B, N, M, S, Kn = 3, 3001, 50, 4, 16
    d = orc.synth_pair(B, N, M, S, kind="lidar", seed=31)
    d["src_node"][0, :, 7] = 500.0                      # empty cluster: mean (0,0,0), pooled features 0
    P = orc.init_detector_params(S=S, seed=9, randomize_bn=True)
    P["mlp3.conv.weight"] = (P["mlp3.conv.weight"] * 1000).astype(np.float32)
    opt = make_opt(batch_size=B, input_pc_num=N, node_num=M, surface_normal_len=S, node_knn_k_1=Kn, use_tensor_cores=use_tc)
    md = ModelDetector(opt)
    load_params(md.detector, P)
    keys = ("src_pc", "src_sn", "src_node", "dst_pc", "dst_sn", "dst_node", "R", "scale", "shift")
    md.set_input(*[torch.from_numpy(d[k]) for k in keys])
    md.forward_loss(train_bn=True)
    r = orc.detector_fwd_loss(P, *[d[k] for k in keys], node_knn_k=Kn, sigma_lower_bound=opt.loss_sigma_lower_bound,
                              alpha=opt.keypoint_on_pc_alpha, training=True)
    kp = torch.cat([md.src_keypoints, md.dst_keypoints]).cpu().numpy()
    sig = torch.cat([md.src_sigmas, md.dst_sigmas]).cpu().numpy()
    assert rel_err(kp, np.concatenate([r["src_keypoints"], r["dst_keypoints"]])) < REL
    assert rel_err(sig, np.concatenate([r["src_sigmas"], r["dst_sigmas"]])) < REL
    assert abs(md.loss.item() - r["loss"]) <= REL * abs(r["loss"])
    md.optimize(epoch=0)                                # backward on ragged sizes must run and stay finite
    torch.cuda.synchronize()
    assert all(torch.isfinite(p.grad).all() for p in md.detector.parameters())


def test_desc_pair_scan_loss_and_model_descriptor_api():
    """DescPairScanLoss forward vs the reference golden; ModelDescriptor.test_model / run_model plumbing."""
    from usip_b200.models import losses
    from usip_b200.models.keypoint_descriptor import ModelDescriptor
    g = golden("desc_loss.npz")
    opt = make_opt(triple_loss_gamma=float(g["gamma"]), sigma_max=float(g["sigma_max"]))
    anc = cu(g["anc"]); pos = cu(g["pos"]); idx = torch.from_numpy(g["neg_idx"]).to(dev())
    with torch.no_grad():
        loss, active = losses.DescPairScanLoss(opt)(anc, pos, anc[idx], cu(g["sig"]))
    assert rel_err(loss.cpu().numpy(), g["loss"]) < 1e-5
    assert np.array_equal(active.cpu().numpy(), g["active"])
    # model-level API on the descriptor golden inputs
    gd = golden("descriptor.npz")
    B, N, M, S, K, seed = [int(v) for v in gd["cfg"]]
    opt = make_opt(batch_size=B, input_pc_num=N, node_num=M, surface_normal_len=S, ball_radius=1.0, ball_nsamples=K, descriptor_len=128)
    md = ModelDescriptor(opt)
    load_params(md.descriptor, {k[len("param/"):]: gd[k] for k in gd.files if k.startswith("param/")})
    np.random.seed(seed)
    desc = md.run_model(cu(gd["pc"]), cu(gd["sn"]), cu(gd["kp"]))
    assert rel_err(desc.cpu().numpy(), gd["eval_desc"]) < REL
    sig = torch.rand(B, M)
    md.set_input(torch.from_numpy(gd["pc"]), torch.from_numpy(gd["sn"]), torch.from_numpy(gd["kp"]), sig,
                 torch.from_numpy(gd["pc"]), torch.from_numpy(gd["sn"]), torch.from_numpy(gd["kp"]), sig, torch.tensor([1, 0]))
    md.test_model()
    assert np.isfinite(md.get_current_errors()["O_loss"])
    before = {k: v.clone() for k, v in md.descriptor.state_dict().items()}
    md.optimize(epoch=0)                                        # the train step itself is pinned by the golden test above
    assert np.isfinite(md.get_current_errors()["O_loss"])
    assert any(not torch.equal(v, before[k]) for k, v in md.descriptor.state_dict().items() if k.endswith("conv.weight"))


@pytest.mark.parametrize("name,B,N,M,S,Kn,kind,lb,alpha", [
    ("modelnet", 24, 5000, 512, 3, 32, "object", 1e-4, 1.0),           # BASELINE configs[1]
    ("kitti", 8, 16384, 512, 4, 16, "lidar", 1e-3, 0.01),              # BASELINE configs[2] (the bench workload)
])
def test_full_size_baseline_configs(name, B, N, M, S, Kn, kind, lb, alpha):
    """BASELINE.json's full-size detector configurations through size-independent checks: the node assignment is
    bit-identical to the C oracle, the tcgen05 (3xTF32) and the fp32 SIMT layer paths agree to 1e-4 on keypoints /
    sigmas / loss, CUDA-graph replay equals eager, and one full optimize() step runs and moves the parameters."""
    from usip_b200 import ops
    from usip_b200.models.keypoint_detector import ModelDetector
    d = orc.synth_pair(B, N, M, S, kind=kind, seed=4321)
    P = orc.init_detector_params(S=S, seed=7, randomize_bn=True)
    P["mlp3.conv.weight"] = (P["mlp3.conv.weight"] * 1000).astype(np.float32)
    keys = ("src_pc", "src_sn", "src_node", "dst_pc", "dst_sn", "dst_node", "R", "scale", "shift")
    x = np.concatenate([d["src_pc"], d["dst_pc"]]); node = np.concatenate([d["src_node"], d["dst_node"]])
    min_idx, _ = ops.som_assign(cu(x), cu(node))
    assert np.array_equal(min_idx.cpu().numpy(), orc.som_assign(x, node))
    res = {}
    for use_tc in (True, False):
        opt = make_opt(batch_size=B, input_pc_num=N, node_num=M, surface_normal_len=S, node_knn_k_1=Kn,
                       loss_sigma_lower_bound=lb, keypoint_on_pc_alpha=alpha, use_tensor_cores=use_tc)
        md = ModelDetector(opt)
        load_params(md.detector, P)
        md.set_input(*[torch.from_numpy(d[k]) for k in keys])
        sd0 = {k: v.clone() for k, v in md.detector.state_dict().items()}
        md.forward_loss(epoch=0, train_bn=True, graph=False)
        res[use_tc] = (torch.cat([md.src_keypoints, md.dst_keypoints]).cpu().numpy(),
                       torch.cat([md.src_sigmas, md.dst_sigmas]).cpu().numpy(), md.loss.item())
        assert np.isfinite(res[use_tc][2]) and md.loss_keypoint_on_pc_src.item() >= 0
        if use_tc:
            md.detector.load_state_dict(sd0)
            md.forward_loss(epoch=0, train_bn=True, graph=True)
            assert md.loss.item() == res[True][2]
            assert np.array_equal(torch.cat([md.src_keypoints, md.dst_keypoints]).cpu().numpy(), res[True][0])
            md.detector.load_state_dict(sd0)
            md.optimize(epoch=0)
            torch.cuda.synchronize()
            assert np.isfinite(md.loss.item())
            assert abs(md.loss.item() - res[True][2]) <= 1e-4 * abs(res[True][2])      # same forward, autograd path
            w = md.detector.state_dict()["mlp1.conv.weight"]
            assert not torch.equal(w, sd0["mlp1.conv.weight"]) and bool(torch.isfinite(w).all())
        del md
        torch.cuda.empty_cache()
    assert rel_err(res[True][0], res[False][0]) < REL and rel_err(res[True][1], res[False][1]) < REL
    assert abs(res[True][2] - res[False][2]) <= REL * abs(res[False][2])


@pytest.mark.parametrize("use_tc", [False, True])
@pytest.mark.parametrize("mode", ["knn", "ball"])
def test_ablation_detectors_vs_reference_golden(mode, use_tc):
    """RPN_Detector_KNN / RPN_Detector_Ball (networks.py:482-738) against outputs of the reference's own classes:
    eval and train forward, every parameter gradient of L = sum(w_kp*kp) + sum(w_sig*sigma), running statistics."""
    from usip_b200.models import networks
    g = golden("detector_ablation.npz")
    B, N, M, S, Kn, seed = [int(v) for v in g["cfg"]]
    d = orc.ablation_inputs(seed, B, N, M, S)
    P = orc.init_ablation_params(S=S, seed=seed, randomize_bn=True)
    P["mlp3.conv.weight"] = (P["mlp3.conv.weight"] * 1000).astype(np.float32)
    opt = make_opt(batch_size=B, input_pc_num=N, node_num=M, surface_normal_len=S, node_knn_k_1=Kn, use_tensor_cores=use_tc)
    net = (networks.RPN_Detector_KNN if mode == "knn" else networks.RPN_Detector_Ball)(opt).to(dev())
    load_params(net, P)
    pc, sn, node = cu(d["pc"]), cu(d["sn"]), cu(d["node"])
    net.eval()
    with torch.no_grad():
        node_o, kp, sig, desc = net(pc, sn, node, False, None)
    assert desc is None and torch.equal(node_o, node)
    assert rel_err(kp.cpu().numpy(), g[mode + "/eval_kp"]) < REL
    assert rel_err(sig.cpu().numpy(), g[mode + "/eval_sig"]) < REL
    net.train()
    _, kp, sig, _ = net(pc, sn, node, True, 0)
    loss = (kp * cu(d["w_kp"])).sum() + (sig * cu(d["w_sig"])).sum()
    net.zero_grad()
    loss.backward()
    torch.cuda.synchronize()
    assert rel_err(kp.detach().cpu().numpy(), g[mode + "/train_kp"]) < REL
    assert rel_err(sig.detach().cpu().numpy(), g[mode + "/train_sig"]) < REL
    assert abs(loss.item() - float(g[mode + "/loss"])) <= 2e-4 * abs(float(g[mode + "/loss"]))
    bad, worst = [], 0.0
    for k, p in net.named_parameters():
        ref = g[mode + "/grad/" + k]
        gr = p.grad.detach().cpu().numpy().reshape(-1).astype(np.float64)
        absmax, norm = float(ref[2]), float(ref[3])
        if k.endswith("conv.bias") and (k.replace("conv.bias", "norm.weight") in dict(net.named_parameters())):
            assert np.linalg.norm(gr) <= norm + 1e-12, k                # a bias in front of a train-mode BN: true gradient 0
            continue
        e_norm = abs(np.linalg.norm(gr) - norm) / max(norm, 1e-12)
        e_el = np.abs(gr[:24] - ref[4:4 + min(24, gr.size)]).max() / max(absmax, 1e-12)
        # element tolerance: two max-pools over K = 64 route whole gradients through single arg-max rows (256 groups per
        # channel here): an arg-max decision flipped by a 1e-7 forward difference moves a sum at the percent level
        # (tests/test_gpu_vs_reference.py measures the reference's own fp32-vs-fp64 spread: 1.4e-2 with 8192 groups per
        # channel).  The norm is the tight check.
        worst = max(worst, e_norm)
        if not (e_norm < 3e-3 and e_el < 3e-2):
            bad.append((k, e_norm, e_el))
    print("ablation %s use_tc=%s: worst gradient norm error %.2e" % (mode, use_tc, worst))
    assert not bad, bad
    sd = net.state_dict()
    for k in sd:
        if k.endswith("running_mean") or k.endswith("running_var"):
            a = sd[k].cpu().numpy().reshape(-1)[:24]
            assert np.allclose(a, g[mode + "/after/" + k], rtol=2e-4, atol=1e-6), k


def test_knn_group_vs_oracle():
    """usip_knn_group_f32: the K nearest points of every centre, exact, ascending index; gathered rows / group."""
    from usip_b200 import ops
    rng = np.random.default_rng(3)
    B, N, M, K, S = 3, 5000, 40, 64, 4
    pc = rng.uniform(-5, 5, (B, 3, N)).astype(np.float32)
    pc[:, :, 100:140] = pc[:, :, 60:100]                              # duplicated points: exact distance ties
    sn = rng.normal(size=(B, S, N)).astype(np.float32)
    ctr = np.ascontiguousarray(pc[:, :, :M]) + rng.normal(0, 0.3, (B, 3, M)).astype(np.float32)
    idx, grp, rows = ops.knn_group(cu(pc), cu(sn), cu(ctr), K, want_group=True, rows_ld=8)
    idx = idx.cpu().numpy(); grp = grp.cpu().numpy(); rows = rows.cpu().numpy().reshape(B, M, K, 8)
    ref_i, ref_d = orc.knn(ctr, pc, K)
    d2 = ((pc[:, :, None, :] - ctr[:, :, :, None]) ** 2)
    d2 = (d2[:, 0] + d2[:, 1]) + d2[:, 2]                             # fp32, the reference's summation order
    for b in range(B):
        for m in range(M):
            sel = idx[b, m]
            assert np.all(np.diff(sel) > 0)                           # ascending, distinct
            kth = np.sort(d2[b, m])[K - 1]
            assert np.all(d2[b, m, sel] <= kth) and np.sum(d2[b, m] < kth) == np.sum(d2[b, m, sel] < kth)
    x_aug = np.concatenate([pc, sn], 1)
    want = np.take_along_axis(x_aug, np.broadcast_to(idx.astype(np.int64).reshape(B, 1, M * K), (B, 3 + S, M * K)), 2).reshape(B, 3 + S, M, K)
    want[:, :3] -= ctr[:, :, :, None]
    assert np.array_equal(grp, want)
    assert np.array_equal(rows[..., :7], want.transpose(0, 2, 3, 1)) and np.all(rows[..., 7] == 0)
