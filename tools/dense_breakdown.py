"""Per-kernel times of the all-vertex LBS operator (BASELINE config 2 and its SMPL-X sibling), L2 flushed between launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from bodyfitting_b200 import synthetic as syn
from bodyfitting_b200.engine import FrameBuffers
from bodyfitting_b200.model import PreparedModel

flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def timed(fn, reps=10):
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)) * 1e3

for mt, B in (('smpl', 1024), ('smplx', 1024), ('smpl', 8192)):
    pm = PreparedModel(mt, syn.make_model(mt, 0), gmm=syn.make_gmm(0),
                       J_regressor_extra=syn.make_J_regressor_extra(seed=0) if mt == 'smpl' else None, device='cuda')
    gt, _ = syn.make_params(mt, B, seed=5)
    T = lambda a: torch.from_numpy(a)
    fb = FrameBuffers(pm, B, full=True)
    fb.t['theta'].copy_(pm.pack_theta(T(gt['global_orient']), T(gt['body_pose']), T(gt['betas'])))
    fb.bind('djoints', torch.randn(B, pm.K_full, 3, device='cuda'))
    fb.t['dverts'].normal_()
    fb.call('bf_lbs_forward'); fb.call('bf_lbs_backward')
    parts = [('pose_fwd', lambda: fb.call('bf_pose_forward')),
             ('skin_fwd', lambda: fb.call('bf_skin_forward', 1)),
             ('joints_fwd', lambda: fb.call('bf_joints_forward', 1)),
             ('joints_bwd', lambda: fb.call('bf_joints_backward', 1, 1)),
             ('skin_bwd_dvp', lambda: fb.call('bf_skin_backward_parts', 1, 1)),
             ('skin_bwd_dA', lambda: fb.call('bf_skin_backward_parts', 1, 2)),
             ('blend_bwd', lambda: fb.call('bf_skin_backward_parts', 1, 4)),
             ('pose_bwd', lambda: fb.call('bf_pose_backward', 0)),
             ('lbs_forward', lambda: fb.call('bf_lbs_forward')),
             ('lbs_backward', lambda: fb.call('bf_lbs_backward'))]
    alg = 4 * (3 * pm.V + 3 * pm.K_out + 3 * pm.J + 10 + 4) * B
    print('%s B=%d alg %.1f MB (%.1f us at 6457 GB/s): ' % (mt, B, alg / 1e6, alg / 6457e3) +
          ' | '.join('%s %.1f' % (n, timed(f)) for n, f in parts) + '  (us)')
