"""`ModelDescriptor` -- API surface of models/keypoint_descriptor.py:14-318 (outdoor descriptor: DescriptorLiteOld +
DescPairScanLoss) on the fused B200 plan: `forward`, `forward_siamese`, `run_model`, `test_model` and the train step
`optimize()` (backward plan in engine.descriptor_backward, loss backward in losses._DescTripletFn).
`ModelDescriptorIndoor` depends on a network that is
broken in the reference itself (networks.py:447 calls a commented-out function) and is out of scope."""
import os
from collections import OrderedDict

import torch

from usip_b200.optim import FlatAdam

from . import losses, networks
from ._common import random_point_dropout


class ModelDescriptor():
    def __init__(self, opt):
        self.opt = opt
        if opt.gpu_ids[0] < 0:
            raise RuntimeError("usip_b200.ModelDescriptor needs a CUDA device: there is no CPU path")
        self.descriptor = networks.DescriptorLiteOld(opt).to(opt.device)
        self.triplet_criteria = losses.DescPairScanLoss(opt)
        self.old_lr_descriptor = self.opt.lr
        self.optimizer_descriptor = FlatAdam(self.descriptor.parameters(), lr=self.old_lr_descriptor,
                                             betas=(0.9, 0.999), weight_decay=0)        # keypoint_descriptor.py:32-35
        dev = opt.device
        B, N, M = opt.batch_size, opt.input_pc_num, opt.node_num
        self.anc_pc = torch.empty(B, 3, N, device=dev).uniform_(); self.anc_sn = torch.empty(B, 3, N, device=dev).uniform_()
        self.anc_keypoints = torch.empty(B, 3, M, device=dev); self.anc_sigmas = torch.empty(B, M, device=dev)
        self.pos_pc = torch.empty(B, 3, N, device=dev).uniform_(); self.pos_sn = torch.empty(B, 3, N, device=dev).uniform_()
        self.pos_keypoints = torch.empty(B, 3, M, device=dev); self.pos_sigmas = torch.empty(B, M, device=dev)
        self.neg_idx = torch.zeros(B, dtype=torch.long, device=dev)
        z = lambda: torch.tensor([0], dtype=torch.float32, requires_grad=False, device=dev)
        self.test_loss_average = z(); self.test_loss_triplet_average = z(); self.test_active_percentage_average = z()

    def set_input(self, anc_pc, anc_sn, anc_keypoints, anc_sigmas, pos_pc, pos_sn, pos_keypoints, pos_sigmas, neg_idx):
        dev = self.opt.device
        self.anc_pc = anc_pc.float().to(dev); self.anc_sn = anc_sn.float().to(dev)
        self.anc_keypoints = anc_keypoints.float().to(dev); self.anc_sigmas = anc_sigmas.float().to(dev)
        self.pos_pc = pos_pc.float().to(dev); self.pos_sn = pos_sn.float().to(dev)
        self.pos_keypoints = pos_keypoints.float().to(dev); self.pos_sigmas = pos_sigmas.float().to(dev)
        self.neg_idx = neg_idx.long().to(dev)
        torch.cuda.synchronize()

    def forward(self, pc, sn, keypoints, is_train=False, epoch=None):
        with torch.cuda.device(pc.get_device()):
            return self.descriptor(pc, sn, keypoints, is_train, epoch)

    def forward_siamese(self, pc_tuple, sn_tuple, keypoints_tuple, is_train=False, epoch=None):
        n = pc_tuple[0].size()[0]
        descriptors, x_aug_ball = self.descriptor(torch.cat(pc_tuple, dim=0), torch.cat(sn_tuple, dim=0),
                                                  torch.cat(keypoints_tuple, dim=0), is_train, epoch)
        return torch.split(descriptors, n, dim=0), torch.split(x_aug_ball, n, dim=0)

    def _loss(self):
        triplet_loss, active_percentage = self.triplet_criteria(self.anc_descriptors, self.pos_descriptors,
                                                                self.anc_descriptors[self.neg_idx, :, :], self.anc_sigmas)
        self.triplet_loss = torch.mean(triplet_loss)
        self.active_percentage = torch.mean(active_percentage)
        self.loss = self.triplet_loss

    def optimize(self, epoch=None):
        """keypoint_descriptor.py:126-157: optional random point dropout (same RNG streams), train-mode forward of the
        (anchor, positive) batch, triplet loss against the in-batch negatives, backward, Adam."""
        with torch.cuda.device(self.anc_pc.get_device()):
            self.anc_pc, self.anc_sn, self.pos_pc, self.pos_sn = random_point_dropout(
                self.opt, self.anc_pc, self.anc_sn, self.pos_pc, self.pos_sn)
            self.descriptor.train()
            (self.anc_descriptors, self.pos_descriptors), _ = self.forward_siamese(
                (self.anc_pc, self.pos_pc), (self.anc_sn, self.pos_sn), (self.anc_keypoints, self.pos_keypoints),
                is_train=True, epoch=epoch)
            self.optimizer_descriptor.zero_grad()
            self._loss()
            self.loss.backward()
            self.optimizer_descriptor.step()

    def test_model(self):
        self.descriptor.eval()
        with torch.cuda.device(self.anc_pc.device), torch.no_grad():
            (self.anc_descriptors, self.pos_descriptors), _ = self.forward_siamese(
                (self.anc_pc, self.pos_pc), (self.anc_sn, self.pos_sn), (self.anc_keypoints, self.pos_keypoints),
                is_train=False, epoch=None)
            self._loss()

    def freeze_model(self):
        for p in self.descriptor.parameters():
            p.requires_grad = False

    def run_model(self, pc, sn, keypoints):
        self.descriptor.eval()
        with torch.cuda.device(pc.device), torch.no_grad():
            descriptors, _ = self.descriptor(pc, sn, keypoints, False, None)
        return descriptors

    def get_negative_samples(self):
        return (self.anc_pc[self.neg_idx, :, :], self.anc_sn[self.neg_idx, :, :],
                self.anc_keypoints[self.neg_idx, :, :], self.anc_sigmas[self.neg_idx, :])

    def get_current_visuals(self):
        raise NotImplementedError("visdom payloads (keypoint_descriptor.py:241-289) are outside the hot path")

    def get_current_errors(self):
        return OrderedDict([('O_loss', self.loss.item()), ('O_triplet', self.triplet_loss.item()),
                            ('O_active_perc', self.active_percentage.item()), ('E_loss', self.test_loss_average.item()),
                            ('E_triplet', self.test_loss_triplet_average.item()),
                            ('E_active_perc', self.test_active_percentage_average.item())])

    def save_network(self, network, network_label, epoch_label, gpu_id):
        torch.save(network.state_dict(), os.path.join(self.opt.checkpoints_dir, '%s_net_%s.pth' % (epoch_label, network_label)))

    def update_learning_rate(self, ratio):
        lr = max(self.old_lr_descriptor * ratio, 0.00001)
        for g in self.optimizer_descriptor.param_groups:
            g['lr'] = lr
        print('update descriptor learning rate: %f -> %f' % (self.old_lr_descriptor, lr))
        self.old_lr_descriptor = lr
