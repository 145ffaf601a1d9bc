"""Lock-step schedule of the recurrent optimiser.

``DepthPoseNet.forward`` (reference: dro_sfm/networks/depth_pose/DepthPoseNet.py:107-205) builds, at the start of every
outer iteration, one depth-cost closure (poses frozen at their iteration-start value) and one pose-cost closure per
source view (depth frozen at its iteration-start value), and only then runs the depth update block (update.py:155-173)
followed by one pose update block per view (update.py:184-199), each for ``seq_len`` inner steps.  Nothing computed
inside one block is read by another before the iteration ends, so inner step j of all 1 + V blocks can be evaluated
side by side.  ``forward`` below does exactly that -- same sub-modules, same arithmetic, same outputs -- and hands the
1 + V feature-metric cost evaluations of an inner step to ONE kernel launch (``ops.feat_cost_batch``) instead of 1 + V
one-wave launches; autograd then also runs their backward as one launch.

``patch.install()`` grafts this method onto the reference's DepthPoseNet (``install(lockstep=False)`` keeps the
reference's own schedule; the per-call drop-ins of ``networks/cost.py`` serve it).
"""
import torch

from . import cost as _cost


def forward(self, target_image, ref_imgs, intrinsics):
    """Inverse depths and poses of a target frame w.r.t. its source frames; returns what the reference returns:
    training -> (list of n inverse-depth maps, poses [B, V, n, 6]); evaluation -> (last map, last poses [B, V, 6])."""
    from dro_sfm.utils.depth import inv2depth          # the patched tree is importable by construction

    n_views, B = len(ref_imgs), target_image.shape[0]
    # one channels_last conversion (and one gradient buffer) for the stacked encoder output, then batch slices
    fmaps = _cost.split_feature_maps(self.fnet(torch.cat([target_image] + list(ref_imgs), dim=0)), [B] * (1 + n_views))
    fmap, fmaps_ref = fmaps[0], list(fmaps[1:])
    assert target_image.shape[2] / fmap.shape[2] == self.feat_ratio
    scale = 1.0 / self.feat_ratio

    # scale_inv_depth (disp_to_depth, DepthPoseNet.py:38-41) is fused into its producers / consumers when the network
    # normalises its output ('-out' versions): epilogue of the convex up-sampling, prologue of the depth cost
    fused_scale = (self.min_depth, self.max_depth) if self.out_normalize else None

    def upsampled_prediction(inv_lr, mask):
        if fused_scale is not None:
            return _cost.upsample_depth(inv_lr, mask, self.feat_ratio, disp_range=fused_scale)
        return self.scale_inv_depth(self.upsample_depth(inv_lr, mask, ratio=self.feat_ratio))[0]

    poses = [self.pose_head(torch.cat([fmap, f], dim=1)) for f in fmaps_ref]
    inv_depth = self.depth_head(fmap, act_fn=torch.sigmoid)
    inv_depth_predictions = [upsampled_prediction(inv_depth, self.upmask_net(fmap))]
    pose_predictions = [[p.clone() for p in poses]]

    if self.iters > 0:
        hidden_d, inp_d = torch.split(self.cnet_depth(target_image), [self.hdim, self.cdim], dim=1)
        hidden_d, inp_d = torch.tanh(hidden_d), torch.relu(inp_d)
        hidden_p, inp_p = [], []
        for ctx in self.cnet_pose([torch.cat([target_image, r], dim=1) for r in ref_imgs]):
            h, c = torch.split(ctx, [self.hdim, self.cdim], dim=1)
            hidden_p.append(torch.tanh(h))
            inp_p.append(torch.relu(c))

    dblock, pblock = self.update_block_depth, self.update_block_pose
    for _ in range(self.iters):
        inv_depth = inv_depth.detach()
        poses = [p.detach() for p in poses]
        frozen_poses = list(poses)                                             # what the depth cost sees all iteration
        frozen_depth = inv2depth(self.scale_inv_depth(inv_depth)[0])           # what every pose cost sees all iteration
        inv_seq, mask_seq, pose_seq = [], [], [[] for _ in range(n_views)]
        for _ in range(self.seq_len):
            # the 1 + V cost evaluations of this inner step: one launch
            if fused_scale is not None:
                jobs = [(inv_depth, fmap, fmaps_ref, frozen_poses, ("disp",) + fused_scale)]
            else:
                jobs = [(self.scale_inv_depth(inv_depth)[0], fmap, fmaps_ref, frozen_poses, True)]
            jobs += [(frozen_depth, fmap, [fmaps_ref[v]], [poses[v]], False) for v in range(n_views)]
            costs = _cost.cost_batch(jobs, intrinsics, intrinsics, scale)
            # depth block, one inner step (update.py:159-171)
            x = torch.cat([inp_d, dblock.encoder(inv_depth, costs[0])], dim=1)
            hidden_d = dblock.depth_gru(hidden_d, x)
            inv_depth = inv_depth + dblock.depth_head(hidden_d)
            mask_seq.append(.25 * dblock.mask(hidden_d))
            inv_seq.append(inv_depth)
            # pose blocks, one inner step each (update.py:186-196)
            for v in range(n_views):
                x = torch.cat([inp_p[v], pblock.encoder(poses[v], costs[1 + v])], dim=1)
                hidden_p[v] = pblock.pose_gru(hidden_p[v], x)
                poses[v] = poses[v] + pblock.pose_head(hidden_p[v])
                pose_seq[v].append(poses[v])
        keep = range(self.seq_len) if self.inter_sup else [self.seq_len - 1]
        for j in keep:
            inv_depth_predictions.append(upsampled_prediction(inv_seq[j], mask_seq[j]))
            pose_predictions.append([pose_seq[v][j].clone() for v in range(n_views)])

    if not self.training:
        return inv_depth_predictions[-1], torch.stack(pose_predictions[-1], dim=1).view(B, n_views, 6)
    return inv_depth_predictions, torch.stack([torch.stack(per_step, dim=1) for per_step in pose_predictions], dim=2)
