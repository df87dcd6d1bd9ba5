"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np
import torch

from bodyfitting_b200 import synthetic as syn
from oracle import fit_port as fp


def make_port(assets, mt, dtype=torch.float32):
    return fp.FitPort(mt, assets(mt), assets('gmm'), assets('jx'), dtype=dtype)


def gt_param_dict(gt, mt):
    gp = dict(global_orient=gt['global_orient'], body_pose=gt['body_pose'], betas=gt['betas'],
              global_transl=gt['transl'], body_scale=gt['scale'])
    if mt == 'smplx':
        gp.update({k: gt[k] for k in ('leye_pose', 'reye_pose', 'left_hand_pose', 'right_hand_pose')})
    return gp


def make_scene(port, mt, B, nv, seed=0):
    """cameras, GT / init parameters and keypoints (from the oracle's GT joints)."""
    c2ws, Ks = syn.make_cameras(nv, seed=seed)
    gt, init = syn.make_params(mt, B, seed=seed)
    K = 135 if mt == 'smplx' else 25
    ev = port.loss_and_grads(gt_param_dict(gt, mt), c2ws, Ks, np.zeros((B, nv, K, 3), np.float32))
    kp = syn.make_keypoints(ev['joints'][:, :K], c2ws, Ks, seed=seed)
    init_pose = np.concatenate([init['global_orient'], init['body_pose']], 1)
    if mt == 'smplx':
        init_pose = np.concatenate([init_pose, np.zeros((B, 6), np.float32)], 1)
    return dict(c2ws=c2ws, Ks=Ks, gt=gt, init=init, kp=kp, init_pose=init_pose, init_betas=init['betas'])


def perturbed_params(mt, B, seed=0):
    """Generic (non-zero everywhere) parameters to exercise every gradient path."""
    gt, _ = syn.make_params(mt, B, seed=seed)
    return gt_param_dict(gt, mt)


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
