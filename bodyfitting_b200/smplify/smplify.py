"""``SMPLify`` -- multi-view SMPL / SMPL-X fitting, same constructor / call signature and
result dict as the reference's ``smplify/smplify.py:19-254``, running on the B200 kernels.

Extensions over the reference (which supports one frame per call, smplify.py:189-190):
``__call__`` also accepts B frames at once -- ``net_output`` = ([B,10], [B,72]) and
``keypoints`` either a list over frames of per-view OpenPose dict lists, or a packed
[B,Nv,K,3] array in model joint order -- each frame being an independent fit with its
own Adam state, exactly as B separate reference calls.

Not implemented here (SURVEY.md 8f "next" rows): ``use_mask`` (silhouette term).
"""
import os
import pickle

import numpy as np
import torch

from .. import constants as K
from ..engine import FrameBuffers, pack_cameras, pack_keypoints
from ..model import PreparedModel
from ..synthetic import openpose_to_keypoints


class SMPLify(object):
    """Implementation of multiview SMPLify (reference: smplify/smplify.py:19-82)."""

    def __init__(self, smpl_type='smpl', age='adult', step_size=1e-2, batch_size=1, num_iters=600,
                 gender='male', use_mask=False, device=torch.device('cuda'), debug=True,
                 model_data=None, gmm=None, J_regressor_extra=None, data_root='data'):
        if age != 'adult':
            raise NotImplementedError("only age='adult' is supported (kid template: smplify.py:114-115)")
        self.device = torch.device(device)
        self.debug = debug
        self.gender = gender
        self.smpl_type = smpl_type
        self.use_hand_face = (smpl_type == 'smplx')
        self.age = age
        self.use_mask = use_mask
        self.batch_size = batch_size
        self.num_iters = num_iters
        if gmm is None:                                   # smplify.py:46-48 -> prior.py:119-128
            fn = os.path.join(data_root, 'gmm_08.pkl')
            if not os.path.exists(fn):
                raise FileNotFoundError('The path to the mixture prior "%s" does not exist' % fn)
            with open(fn, 'rb') as f:
                gmm = pickle.load(f, encoding='latin1')
        if model_data is None:                            # smplify.py:51,63
            model_data = os.path.join(data_root, smpl_type)
        if smpl_type == 'smpl' and J_regressor_extra is None:
            fn = os.path.join(data_root, 'J_regressor_extra.npy')      # config.py:1, models/smpl.py:62
            J_regressor_extra = np.load(fn) if os.path.exists(fn) else None
        self.model = PreparedModel(smpl_type, model_data, gmm=gmm, J_regressor_extra=J_regressor_extra,
                                   device=self.device)
        self.smpl_faces = self.model.faces.astype(np.int32).reshape(1, -1, 3)
        self.last_trace = None
        self.last_loss_terms = None

    # ------------------------------------------------------------------------------------------
    def _pack_inputs(self, net_output, keypoints):
        init_betas, init_poses = net_output
        init_betas = torch.as_tensor(init_betas, dtype=torch.float32).reshape(-1, 10)
        init_poses = torch.as_tensor(init_poses, dtype=torch.float32)
        init_poses = init_poses.reshape(init_betas.shape[0], -1)
        B = init_poses.shape[0]
        if isinstance(keypoints, (np.ndarray, torch.Tensor)):
            kp = torch.as_tensor(keypoints, dtype=torch.float32)
            if kp.dim() == 3:
                kp = kp[None]
        else:
            frames = keypoints if (len(keypoints) and isinstance(keypoints[0], (list, tuple))) else [keypoints]
            kp = torch.from_numpy(np.stack([openpose_to_keypoints(v, self.smpl_type) for v in frames]))
        assert kp.shape[0] == B, 'keypoints for %d frames, parameters for %d' % (kp.shape[0], B)
        return init_betas, init_poses, kp

    def __call__(self, net_output, c2ws, Ks, keypoints, output_folder=None, use_mask=False, masks=None,
                 use_frames=[0], mask_frames=[0], keyframe=6, imsize=512, use_mesh=False, meshfile=None,
                 displacement=False, return_vertices=True, as_numpy=True):
        if use_mask:
            raise NotImplementedError('silhouette term (smplify/loss.py:85-130) is not part of this build')
        if use_mesh:
            raise NotImplementedError('scan term: use bodyfitting_b200.smplify.smpld (SMPL+D path)')
        m = self.model
        dev = self.device
        N = int(self.num_iters)
        assert N >= 1
        init_betas, init_poses, kp = self._pack_inputs(net_output, keypoints)
        B, Nv = kp.shape[0], kp.shape[1]
        assert kp.shape[2] == m.K_used, 'expected %d keypoints per view, got %d' % (m.K_used, kp.shape[2])
        assert len(c2ws) == Nv and len(Ks) == Nv

        fb = FrameBuffers(m, B, full=False, Nv=Nv, n_trace=N, imsize=imsize)
        # init: body pose / betas / global orient from the network, transl 0, scale 1, rest 0 (smplify.py:103-128)
        theta = m.pack_theta(init_poses[:, :3].to(dev, non_blocking=True), init_poses[:, 3:3 + m.nbody].to(dev, non_blocking=True),
                             init_betas.to(dev, non_blocking=True))
        fb.t['theta'].copy_(theta)
        kp_dev = pack_keypoints(kp.to(dev, non_blocking=True), self.use_hand_face)
        cams = torch.from_numpy(pack_cameras(c2ws, Ks)).to(dev, non_blocking=True)
        fb.bind('kp', kp_dev)
        fb.bind('cams', cams)

        if N > 1:
            fb.call('bf_fit_run', N - 1)
        # the reference returns vertices / joints / full_pose of the LAST forward pass (parameters
        # before the final Adam step) together with the parameters after it (smplify.py:216-226)
        theta_prev = fb.t['theta'].clone()
        fb.struct.iter = N - 1
        fb.call('bf_fit_step')
        out = self._final_outputs(theta_prev, fb.t['theta'], return_vertices)
        self.last_trace = fb.t['trace']
        self.last_loss_terms = fb.t['loss_terms']
        if as_numpy:
            out = {k: (self.cpu(v) if torch.is_tensor(v) else v) for k, v in out.items()}
        return out

    def _final_outputs(self, theta_prev, theta, return_vertices, chunk=4096):
        m = self.model
        B = theta.shape[0]
        cs = K.CONSTANT_SCALE_NO_SCAN
        sp, sn = m.split_theta(theta_prev), m.split_theta(theta)
        verts = torch.empty(B, m.V, 3, device=self.device) if return_vertices else None
        joints = torch.empty(B, m.K_out, 3, device=self.device)
        full_pose = torch.empty(B, 3 * m.J, device=self.device)
        for lo in range(0, B, chunk):
            hi = min(B, lo + chunk)
            ext = dict(theta=theta_prev[lo:hi].contiguous(), joints=joints[lo:hi], full_pose=full_pose[lo:hi])
            if return_vertices:
                ext['verts'] = verts[lo:hi].view(hi - lo, -1)
            full = FrameBuffers(m, hi - lo, full=True, need_backward=False, ext=ext)
            full.call('bf_lbs_forward')
        t, s = sp['transl'][:, None, :], sp['scale'][:, None, :]
        out = {}
        if return_vertices:
            out['vertices'] = (verts + t) * s * cs
        out.update(joints=(joints + t) * s * cs, pose=sn['body_pose'], betas=sn['betas'],
                   global_orient=sn['global_orient'], faces=self.smpl_faces[0],
                   global_transl=sn['transl'] * sn['scale'], scale=sn['scale'], full_pose=full_pose)
        if self.use_hand_face:
            out.update(leye_pose=sn['leye_pose'], reye_pose=sn['reye_pose'],
                       left_hand_pose=sn['left_hand_pose'], right_hand_pose=sn['right_hand_pose'])
        return out

    def cpu(self, tensor):
        return tensor.detach().cpu().squeeze(0).numpy()
