"""Runs the all-vertex LBS forward a few times (for an ncu launch list): python tools/dense_fwd_once.py [smpl|smplx] [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from bodyfitting_b200 import synthetic as syn
from bodyfitting_b200.engine import FrameBuffers
from bodyfitting_b200.model import PreparedModel
mt = sys.argv[1] if len(sys.argv) > 1 else 'smpl'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
pm = PreparedModel(mt, syn.make_model(mt, 0), gmm=syn.make_gmm(0),
                   J_regressor_extra=syn.make_J_regressor_extra(seed=0) if mt == 'smpl' else None, device='cuda')
gt, _ = syn.make_params(mt, B, seed=5)
T = lambda a: torch.from_numpy(a)
fb = FrameBuffers(pm, B, full=True)
fb.t['theta'].copy_(pm.pack_theta(T(gt['global_orient']), T(gt['body_pose']), T(gt['betas'])))
fb.bind('djoints', torch.randn(B, pm.K_full, 3, device='cuda'))
fb.t['dverts'].normal_()
for _ in range(3):
    fb.call('bf_lbs_forward'); fb.call('bf_lbs_backward')
torch.cuda.synchronize()
print('done')
