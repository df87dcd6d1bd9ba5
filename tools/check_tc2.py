"""Forward blend GEMM on the active set and on all vertices: max error vs an fp64 matmul, and the kernel time.
Run once with BODYFIT_TC2=0 and once with BODYFIT_TC2=1 (CTA-pair kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from bodyfitting_b200 import synthetic as syn
from bodyfitting_b200.engine import FrameBuffers
from bodyfitting_b200.model import PreparedModel

pm = PreparedModel('smplx', syn.make_model('smplx', 0), gmm=syn.make_gmm(0), device='cuda')
for B, full in ((10000, False), (1000, False), (300, False), (2048, True)):
    gt, _ = syn.make_params('smplx', B, seed=5)
    T = lambda a: torch.from_numpy(a)
    fb = FrameBuffers(pm, B, full=full, Nv=8)
    fb.t['theta'].copy_(pm.pack_theta(T(gt['global_orient']), T(gt['body_pose']), T(gt['betas'])))
    fb.call('bf_pose_forward')
    fb.t['vposed'].fill_(float('nan'))
    fb.call('bf_blend_forward', 1 if full else 0)
    torch.cuda.synchronize()
    vs = pm.struct.full if full else pm.struct.act
    n, ldn = vs.n, vs.ldn
    Bm = pm._dev[('full' if full else 'act') + '_Bm']
    ref = (fb.t['pf_hi'].double() + fb.t['pf_lo'].double()) @ Bm.double()
    out = fb.t['vposed'][:, :3 * n].double()
    err = (out - ref[:, :3 * n]).abs().max().item()
    scale = ref[:, :3 * n].abs().max().item()
    ts = []
    for _ in range(20):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fb.call('bf_blend_forward', 1 if full else 0); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    print('TC2=%s B=%d full=%d  max |err| %.3e (|ref| max %.3f, nan %d)  %.1f us' % (os.environ.get('BODYFIT_TC2', '0'), B, full, err, scale,
          int(torch.isnan(fb.t['vposed'][:, :3 * n]).sum()), 1e3 * float(np.median(ts))), flush=True)
