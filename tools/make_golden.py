"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (lijx10/USIP at /root/reference, unmodified) on CPU
through oracle/ref_shim.py.  Only runs in the build container (the reference is not shipped); the outputs are
small, committed, and are what pins oracle/ (tests/test_oracle_vs_golden.py) and the CUDA path
(tests/test_gpu_*.py) to the reference.

Inputs are regenerated from seeds by oracle.usip_oracle.{synth_pair,init_detector_params}, so the fixtures only
hold outputs (+ tiny inputs where no generator exists).

    python tools/make_golden.py
"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim, usip_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def load_params(module, P):
    sd = module.state_dict()
    for k in sd:
        assert k in P, k
        sd[k] = torch.from_numpy(np.asarray(P[k])).reshape(sd[k].shape).to(sd[k].dtype)
    module.load_state_dict(sd)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def golden_index_max(ref):
    import index_max as im  # the shimmed module -> reference C++ forward_cpu
    rng = np.random.default_rng(7)
    cases = {}
    # (a) generic, (b) duplicates/ties, (c) empty clusters + values <= -1000, (d) B*K beyond the reference smem cap
    specs = dict(a=(2, 8, 1000, 16), b=(2, 4, 512, 8), c=(2, 4, 300, 32), d=(30, 2, 700, 512))
    for name, (B, C, N, K) in specs.items():
        data = rng.normal(size=(B, C, N)).astype(np.float32)
        index = rng.integers(0, K, size=(B, N)).astype(np.int32)
        if name == "b":
            data = np.round(data * 2) / 2          # many exact ties
        if name == "c":
            index = rng.integers(0, K // 2, size=(B, N)).astype(np.int32)   # upper half empty
            data[:, 0, :] = -2000.0                                        # never above the -1000 floor
            data[:, 1, ::3] = -1000.0
        out = im.forward_cpu(t(data), t(index), K).numpy()
        cases["data_" + name] = data; cases["index_" + name] = index; cases["out_" + name] = out
        cases["K_" + name] = np.int32(K)
    np.savez_compressed(os.path.join(OUT, "index_max.npz"), **cases)


def golden_query_topk(ref):
    out = {}
    for name, (B, N, M, kind, seed) in dict(lidar=(2, 2048, 64, "lidar", 11), obj=(3, 1500, 32, "object", 12)).items():
        d = orc.synth_pair(B, N, M, 4 if kind == "lidar" else 3, kind=kind, seed=seed)
        node = d["src_node"].copy()
        if name == "obj":
            node[:, :, 5] = node[:, :, 4]          # duplicated node -> exact distance ties
            node[:, :, 9] = 100.0                  # far away node -> empty cluster
        mask, row_max, min_idx = ref.som.query_topk(t(node), t(d["src_pc"]), M, 1)
        out["node_" + name] = node; out["seed_" + name] = np.int32(seed)
        out["min_idx_" + name] = min_idx.numpy().astype(np.int32)
        out["row_max_" + name] = row_max.numpy().astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "query_topk.npz"), **out)


def golden_losses(ref):
    torch.manual_seed(1234 + 0)
    B, M = 2, 256                                   # BASELINE.json configs[0]
    src = torch.randn(B, 3, M); dst = torch.randn(B, 3, M)
    ss = torch.rand(B, M) + 0.01; sd = torch.rand(B, M) + 0.01
    opt = ref_shim.make_opt()
    crit = ref.losses.ChamferLoss_Brute(opt)
    a = src.clone().requires_grad_(True); b = dst.clone().requires_grad_(True)
    sa = ss.clone().requires_grad_(True); sb = sd.clone().requires_grad_(True)
    loss, pure, weighted = crit(a, b, sa, sb)
    loss.backward()
    out = dict(src=src.numpy(), dst=dst.numpy(), sig_src=ss.numpy(), sig_dst=sd.numpy(),
               loss=loss.detach().numpy(), pure=pure.numpy(), weighted=weighted.numpy(),
               g_src=a.grad.numpy(), g_dst=b.grad.numpy(), g_sig_src=sa.grad.numpy(), g_sig_dst=sb.grad.numpy())
    # single side
    kp = torch.randn(B, 3, 64); pc = torch.randn(B, 3, 700)
    kpg = kp.clone().requires_grad_(True)
    ssc = ref.losses.SingleSideChamferLoss_Brute(opt)(kpg, pc)
    ssc.mean().backward()
    out.update(kp=kp.numpy(), pc=pc.numpy(), single=ssc.detach().numpy(), g_kp=kpg.grad.numpy())
    # point_to_plane (PointOnSurfaceLoss through KeypointOnPCLoss): value + gradient w.r.t. the keypoints
    sn = torch.nn.functional.normalize(torch.randn(B, 3, 700), dim=1)
    sn4 = torch.cat([sn, torch.rand(B, 1, 700)], 1)              # S=4: only the first three channels are the normal
    kpp = kp.clone().requires_grad_(True)
    pos = ref.losses.KeypointOnPCLoss(opt)(kpp, pc, sn4)
    (pos.mean() * 0.37).backward()
    out.update(sn4=sn4.numpy(), on_surface=pos.detach().numpy(), g_kp_surface=kpp.grad.numpy())
    # no-sigma branch
    l2, p2, w2 = crit(src, dst)
    out.update(nosigma=l2.numpy())
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **out)


def detector_case(ref, name, B, N, M, S, Kn, kind, seed, scene="outdoor", lb=1e-3, alpha=0.01):
    d = orc.synth_pair(B, N, M, S, kind=kind, seed=seed)
    opt = ref_shim.make_opt(batch_size=B, input_pc_num=N, node_num=M, surface_normal_len=S, node_knn_k_1=Kn,
                            scene=scene, loss_sigma_lower_bound=lb, keypoint_on_pc_alpha=alpha)
    torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
    md = ref.keypoint_detector.ModelDetector(opt)
    C1, C2 = (64, 256) if scene == "indoor" else (128, 512)
    P = orc.init_detector_params(S=S, seed=seed, C1=C1, C2=C2, randomize_bn=True)
    P["mlp3.conv.weight"] = (P["mlp3.conv.weight"] * 1000).astype(np.float32)   # make keypoint offsets non-trivial
    load_params(md.detector, P)
    out = dict(cfg=np.array([B, N, M, S, Kn, seed], np.int64), kind=kind, scene=scene,
               lb=np.float32(lb), alpha=np.float32(alpha))
    args = [t(d[k]) for k in ("src_pc", "src_sn", "src_node", "dst_pc", "dst_sn", "dst_node", "R", "scale", "shift")]
    md.set_input(*args)
    # --- eval forward + loss
    md.test_model()
    out.update(eval_kp=torch.cat([md.src_keypoints, md.dst_keypoints]).detach().numpy(),
               eval_sig=torch.cat([md.src_sigmas, md.dst_sigmas]).detach().numpy(),
               eval_node=torch.cat([md.src_node_recomputed, md.dst_node_recomputed]).detach().numpy(),
               eval_loss=np.array([md.loss.item(), md.loss_chamfer.item(), md.chamfer_pure.item(),
                                   md.chamfer_weighted.item(), md.loss_keypoint_on_pc_src.item(),
                                   md.loss_keypoint_on_pc_dst.item()], np.float32))
    # --- one training step (train-mode BN forward, backward, Adam)
    md.optimize(epoch=0)
    out.update(train_kp=torch.cat([md.src_keypoints, md.dst_keypoints]).detach().numpy(),
               train_sig=torch.cat([md.src_sigmas, md.dst_sigmas]).detach().numpy(),
               train_loss=np.array([md.loss.item(), md.loss_chamfer.item(), md.chamfer_pure.item(),
                                    md.chamfer_weighted.item(), md.loss_keypoint_on_pc_src.item(),
                                    md.loss_keypoint_on_pc_dst.item()], np.float32))
    sd = md.detector.state_dict()
    for k, v in sd.items():
        v = v.detach().numpy()
        if k.endswith("num_batches_tracked"):
            out["nbt/" + k] = v
            continue
        flat = v.reshape(-1).astype(np.float64)
        out["after/" + k] = np.concatenate([[flat.mean(), flat.std(), np.abs(flat).max()], flat[:24]]).astype(np.float32)
    grads = {}
    for k, p in md.detector.named_parameters():
        g = p.grad.detach().numpy().reshape(-1).astype(np.float64)
        grads["grad/" + k] = np.concatenate([[g.mean(), g.std(), np.abs(g).max(), np.linalg.norm(g)], g[:24]]).astype(np.float32)
    out.update(grads)
    np.savez_compressed(os.path.join(OUT, "detector_%s.npz" % name), **out)
    print("detector", name, "eval loss", out["eval_loss"], "train loss", out["train_loss"])


def golden_descriptor(ref):
    B, N, M, S, K = 2, 2048, 48, 4, 32
    seed = 21
    rng = np.random.default_rng(seed)
    d = orc.synth_pair(B, N, 16, S, kind="lidar", seed=seed)
    pc = d["src_pc"] * np.array([0.2, 1.0, 0.2], np.float32).reshape(1, 3, 1)      # denser cloud: balls get 0..>K hits
    sel = rng.choice(N, M, replace=False)
    kp = (pc[:, :, sel] + rng.normal(0, 0.1, (B, 3, M))).astype(np.float32)
    kp[:, :, 0] = 500.0                                                             # a keypoint with zero hits
    opt = ref_shim.make_opt(batch_size=B, input_pc_num=N, node_num=M, surface_normal_len=S, ball_radius=1.0,
                            ball_nsamples=K, descriptor_len=128)
    torch.manual_seed(seed)
    net = ref.networks.DescriptorLiteOld(opt)
    out = dict(cfg=np.array([B, N, M, S, K, seed], np.int64), pc=pc, sn=d["src_sn"], kp=kp)
    sdn = {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}
    for k, v in sdn.items():
        out["param/" + k] = v
    for mode in ("eval", "train"):
        net.train(mode == "train")
        np.random.seed(seed)                               # the forward consumes np.random.permutation(N)
        perm = np.random.RandomState(seed).permutation(N)  # same stream, recorded for the consumers
        np.random.seed(seed)
        with torch.no_grad():
            desc, feats = net(t(pc), t(d["src_sn"]), t(kp), mode == "train", None)
        out[mode + "_desc"] = desc.numpy(); out[mode + "_feats"] = feats.numpy()
        out["perm"] = perm.astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "descriptor.npz"), **out)


def golden_descriptor_train(ref):
    """One ModelDescriptor.optimize() of the reference (keypoint_descriptor.py:126-157) on CPU: loss, gradient
    summaries, post-Adam parameter summaries.  Inputs are regenerated from the seed by the consumers."""
    B, N, M, S, K = 3, 2048, 40, 4, 32
    seed = 33
    d = desc_train_inputs(B, N, M, S, seed)
    opt = ref_shim.make_opt(batch_size=B, input_pc_num=N, node_num=M, surface_normal_len=S, ball_radius=1.0,
                            ball_nsamples=K, descriptor_len=128, random_pc_dropout_lower_limit=1.0)
    torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
    md = ref.keypoint_descriptor.ModelDescriptor(opt)
    out = dict(cfg=np.array([B, N, M, S, K, seed], np.int64))
    for k, v in md.descriptor.state_dict().items():
        out["param/" + k] = v.detach().numpy().copy()
    md.set_input(*[t(d[k]) for k in ("anc_pc", "anc_sn", "anc_kp", "anc_sigma", "pos_pc", "pos_sn", "pos_kp", "pos_sigma")],
                 torch.from_numpy(d["neg_idx"]))
    np.random.seed(seed)                                   # the forward consumes np.random.permutation(N)
    md.optimize(epoch=0)
    out["loss"] = np.float32(md.loss.item()); out["active"] = np.float32(md.active_percentage.item())
    out["anc_desc"] = md.anc_descriptors.detach().numpy(); out["pos_desc"] = md.pos_descriptors.detach().numpy()
    for k, p in md.descriptor.named_parameters():
        g = p.grad.detach().numpy().reshape(-1).astype(np.float64)
        out["grad/" + k] = np.concatenate([[g.mean(), g.std(), np.abs(g).max(), np.linalg.norm(g)], g[:24]]).astype(np.float32)
    for k, v in md.descriptor.state_dict().items():
        v = v.detach().numpy()
        if k.endswith("num_batches_tracked"):
            continue
        flat = v.reshape(-1).astype(np.float64)
        out["after/" + k] = np.concatenate([[flat.mean(), flat.std(), np.abs(flat).max()], flat[:24]]).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "descriptor_train.npz"), **out)
    print("descriptor train loss", out["loss"], "active", out["active"])


def desc_train_inputs(B, N, M, S, seed):
    """Synthetic (anchor, positive) scan pair with keypoints + sigmas for the descriptor train step (shared with the
    tests through oracle.usip_oracle.desc_train_inputs)."""
    return orc.desc_train_inputs(B, N, M, S, seed)


def golden_desc_loss(ref):
    torch.manual_seed(5)
    B, C, M = 3, 128, 40
    anc = torch.nn.functional.normalize(torch.randn(B, C, M), dim=1)
    pos = torch.nn.functional.normalize(anc + 0.3 * torch.randn(B, C, M), dim=1)
    neg_idx = torch.tensor([1, 2, 0])
    sig = torch.rand(B, M) * 4
    opt = ref_shim.make_opt()
    loss, act = ref.losses.DescPairScanLoss(opt)(anc, pos, anc[neg_idx], sig)
    np.savez_compressed(os.path.join(OUT, "desc_loss.npz"), anc=anc.numpy(), pos=pos.numpy(), neg_idx=neg_idx.numpy(),
                        sig=sig.numpy(), loss=loss.numpy(), active=act.numpy(),
                        gamma=np.float32(opt.triple_loss_gamma), sigma_max=np.float32(opt.sigma_max))


def golden_fps(ref):
    """FarthestSampler of the reference's KITTI loader (float64 distances, np.argmax) on three small clouds: lidar-like,
    a cloud with duplicated points (exact ties) and k larger than the number of distinct points."""
    import importlib
    loader = importlib.import_module("data.kitti_detector_loader")
    fs = loader.FarthestSampler()
    out = {}
    rng = np.random.default_rng(31)
    clouds = dict(lidar=(rng.uniform(-40, 40, (700, 3)) * np.array([1, 0.05, 1])).astype(np.float32),
                  dup=np.repeat(rng.normal(size=(60, 3)).astype(np.float32), 5, axis=0),
                  few=np.repeat(rng.normal(size=(5, 3)).astype(np.float32), 4, axis=0))
    for name, pts in clouds.items():
        k = dict(lidar=64, dup=48, few=12)[name]
        np.random.seed(77)
        nodes = fs.sample(pts, k)                       # consumes np.random.randint(len(pts)) for the first node
        np.random.seed(77)
        start = np.random.randint(len(pts))
        out["pts_" + name] = pts; out["k_" + name] = np.int32(k); out["start_" + name] = np.int32(start)
        out["nodes_" + name] = nodes
    np.savez_compressed(os.path.join(OUT, "fps.npz"), **out)


def golden_ablation(ref):
    """RPN_Detector_KNN / RPN_Detector_Ball (models/networks.py:482-738): eval and train forward, and the gradients of
    L = sum(w_kp * keypoints) + sum(w_sig * sigmas) (ModelDetector cannot select these networks -- the lines are commented
    out at keypoint_detector.py:23-24 -- so the fixture drives the networks directly)."""
    B, N, M, S, Kn, seed = 2, 4096, 128, 4, 16, 1239          # 256 groups per channel: one flipped arg-max moves a sum by < 1 %
    d = orc.ablation_inputs(seed, B, N, M, S)
    opt = ref_shim.make_opt(batch_size=B, input_pc_num=N, node_num=M, surface_normal_len=S, node_knn_k_1=Kn)
    P = orc.init_ablation_params(S=S, seed=seed, randomize_bn=True)
    P["mlp3.conv.weight"] = (P["mlp3.conv.weight"] * 1000).astype(np.float32)
    out = dict(cfg=np.array([B, N, M, S, Kn, seed], np.int64))
    for mode, cls in (("knn", ref.networks.RPN_Detector_KNN), ("ball", ref.networks.RPN_Detector_Ball)):
        torch.manual_seed(seed)
        net = cls(opt)
        load_params(net, P)
        net.eval()
        with torch.no_grad():
            node_o, kp, sig, _ = net(t(d["pc"]), t(d["sn"]), t(d["node"]), False, None)
        assert torch.equal(node_o, t(d["node"]))
        out[mode + "/eval_kp"] = kp.numpy(); out[mode + "/eval_sig"] = sig.numpy()
        net.train()
        _, kp, sig, _ = net(t(d["pc"]), t(d["sn"]), t(d["node"]), True, 0)
        loss = (kp * t(d["w_kp"])).sum() + (sig * t(d["w_sig"])).sum()
        net.zero_grad()
        loss.backward()
        out[mode + "/train_kp"] = kp.detach().numpy(); out[mode + "/train_sig"] = sig.detach().numpy()
        out[mode + "/loss"] = np.float32(loss.item())
        for k, p_ in net.named_parameters():
            g = p_.grad.detach().numpy().reshape(-1).astype(np.float64)
            out[mode + "/grad/" + k] = np.concatenate([[g.mean(), g.std(), np.abs(g).max(), np.linalg.norm(g)], g[:24]]).astype(np.float32)
        for k, v in net.state_dict().items():
            if k.endswith("running_mean") or k.endswith("running_var"):
                out[mode + "/after/" + k] = v.detach().numpy().reshape(-1)[:24].copy()
        print("ablation", mode, "loss", out[mode + "/loss"])
    np.savez_compressed(os.path.join(OUT, "detector_ablation.npz"), **out)


def _reference_function(relpath, name):
    """The reference's own, unmodified source of one top-level function, executed here (its module cannot be imported:
    evaluation/save_keypoints.py is a script with hard-coded dataset paths and GUI / visdom imports)."""
    import ast
    src = open(os.path.join(ref_shim.build_ref_root(), relpath)).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"np": np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), relpath, "exec"), ns)
    return ns[name]


def golden_nms(ref):
    """evaluation/save_keypoints.py:180-216 nms() + the sigma-sorted top-k that follows it (:346-351)."""
    nms = _reference_function("evaluation/save_keypoints.py", "nms")
    rng = np.random.default_rng(41)
    out = {}
    cases = dict(lidar=(rng.uniform(-40, 40, (512, 3)) * np.array([1, 0.05, 1]), 2.0),
                 dense=(rng.normal(size=(300, 3)) * 1.5, 1.0),
                 ties=(np.round(rng.normal(size=(200, 3)) * 2) / 2, 0.5),       # duplicated positions, exact distance ties
                 off=(rng.normal(size=(64, 3)), 0.0))                            # radius < 0.01: pass-through
    for name, (kp, r) in cases.items():
        kp = kp.astype(np.float32)
        sg = (rng.uniform(0.01, 2.0, kp.shape[0])).astype(np.float32)
        if name == "ties":
            sg = np.round(sg * 4) / 4                                            # equal sigmas too
            sg = sg.astype(np.float32)
        vk, vs = nms(kp.copy(), sg.copy(), r)
        out["kp_" + name] = kp; out["sigma_" + name] = sg; out["radius_" + name] = np.float32(r)
        out["valid_kp_" + name] = vk; out["valid_sigma_" + name] = vs
    np.savez_compressed(os.path.join(OUT, "nms.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    ref = ref_shim.modules()
    if len(sys.argv) > 1:                                  # regenerate selected fixtures only, e.g. `make_golden.py ablation`
        for name in sys.argv[1:]:
            globals()["golden_" + name](ref)
        return
    golden_index_max(ref)
    golden_query_topk(ref)
    golden_losses(ref)
    detector_case(ref, "kitti_small", B=2, N=2048, M=64, S=4, Kn=16, kind="lidar", seed=1237)
    detector_case(ref, "modelnet_small", B=3, N=1000, M=32, S=3, Kn=32, kind="object", seed=1236, lb=1e-4, alpha=1.0)
    detector_case(ref, "lite_small", B=2, N=1024, M=32, S=4, Kn=16, kind="lidar", seed=1238, scene="indoor")
    golden_descriptor(ref)
    golden_desc_loss(ref)
    golden_descriptor_train(ref)
    golden_fps(ref)
    golden_nms(ref)
    golden_ablation(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KB")


if __name__ == "__main__":
    main()
