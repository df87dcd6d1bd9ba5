"""Priors of the reference's ``smplify/prior.py`` (same class names / call signatures) on the B200
kernels: ``MaxMixturePrior`` (merged likelihood, :181-196), ``SMPLifyAnglePrior`` (:53-89),
``L2Prior`` (:92-97), ``create_prior`` (:36-50)."""
import os
import pickle
import types

import numpy as np
import torch
import torch.nn as nn

from .. import ops

DEFAULT_DTYPE = torch.float32


def create_prior(prior_type, **kwargs):
    if prior_type == 'gmm':
        return MaxMixturePrior(**kwargs)
    if prior_type == 'l2':
        return L2Prior(**kwargs)
    if prior_type == 'angle':
        return SMPLifyAnglePrior(**kwargs)
    if prior_type == 'none' or prior_type is None:
        return lambda *a, **k: 0.0
    raise ValueError('Prior {}'.format(prior_type) + ' is not implemented')


class SMPLifyAnglePrior(nn.Module):
    def __init__(self, dtype=torch.float32, **kwargs):
        super().__init__()

    def forward(self, pose, with_global_pose=False):
        """pose (B, 69) without / (B, 72) with the global orientation -> (B,4)"""
        body = pose[:, 3:] if with_global_pose else pose
        return ops.angle_prior_op(body)


class L2Prior(nn.Module):
    def __init__(self, dtype=DEFAULT_DTYPE, reduction='sum', **kwargs):
        super().__init__()

    def forward(self, module_input, *args):
        return torch.sum(module_input.pow(2))


class MaxMixturePrior(nn.Module):
    """min_m [0.5 (x - mu_m)^T P_m (x - mu_m) - log nll_w_m]; ``gmm`` may be given directly, otherwise it is
    read from ``<prior_folder>/gmm_{num_gaussians:02d}.pkl`` as the reference does (:119-128)."""

    def __init__(self, prior_folder='prior', num_gaussians=6, dtype=DEFAULT_DTYPE, epsilon=1e-16, use_merged=True,
                 gmm=None, device='cuda', **kwargs):
        super().__init__()
        if not use_merged:
            raise NotImplementedError('only the merged likelihood (the reference default, :104,:228) is implemented')
        if gmm is None:
            fn = os.path.join(prior_folder, 'gmm_{:02d}.pkl'.format(num_gaussians))
            if not os.path.exists(fn):
                raise FileNotFoundError('The path to the mixture prior "{}" does not exist'.format(fn))
            with open(fn, 'rb') as f:
                gmm = pickle.load(f, encoding='latin1')
        self.num_gaussians = len(gmm['weights'])
        self._tables = _GmmTables(gmm, device)
        self.random_var_dim = 69

    def forward(self, pose, betas=None):
        return ops.gmm_pose(pose, self._tables)


class _GmmTables(object):
    """just the GMM part of the model struct (bf_op_gmm_pose reads nothing else)"""

    def __init__(self, gmm, device):
        from .. import _lib
        means = np.asarray(gmm['means']).astype(np.float32)
        covs = np.asarray(gmm['covars']).astype(np.float32)
        prec = np.stack([np.linalg.inv(c) for c in covs]).astype(np.float32)
        sqrdets = np.array([np.sqrt(np.linalg.det(c)) for c in np.asarray(gmm['covars'])])
        const = (2 * np.pi) ** (69 / 2.)
        nllw = np.asarray(np.asarray(gmm['weights']) / (const * (sqrdets / sqrdets.min()))).astype(np.float32)
        psym = np.zeros((means.shape[0], 69, 72), dtype=np.float32)
        psym[:, :, :69] = (prec + np.transpose(prec, (0, 2, 1))) * np.float32(0.5)
        self._keep = [torch.from_numpy(a).to(device) for a in (means, psym, np.log(nllw).astype(np.float32))]
        self.struct = _lib.BfModel()
        self.struct.gmm_mean, self.struct.gmm_psym, self.struct.gmm_logw = [t.data_ptr() for t in self._keep]
        self.struct.n_gmm = means.shape[0]
