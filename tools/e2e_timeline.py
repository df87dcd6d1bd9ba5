"""Per-part GPU timeline of the end-to-end path (BODYFIT_E2E_TRACE=1): for every part the times (ms after the first part's
start) at which its inputs are on the device, its fit is done and its results are in host memory; plus the wall clock.
    BODYFIT_PARTS=4 python tools/e2e_timeline.py [B]"""
import json, os, sys, time
os.environ['BODYFIT_E2E_TRACE'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from bodyfitting_b200 import synthetic as syn
from bodyfitting_b200.smplify.smplify import SMPLify

B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
fit = SMPLify(smpl_type='smplx', num_iters=100, gender='neutral', model_data=syn.make_model('smplx', 0), gmm=syn.make_gmm(0))
wl = bench.build_workload(fit.model, 10000, 100, 0, B)
pin = {k: torch.from_numpy(wl[k]).pin_memory() for k in ('kp', 'init_pose', 'init_betas')}
args = ((pin['init_betas'].numpy(), pin['init_pose'].numpy()), list(wl['c2ws']), list(wl['Ks']), pin['kp'].numpy(), None)
for _ in range(3):
    fit(*args, use_frames=list(range(8)), imsize=512)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fit(*args, use_frames=list(range(8)), imsize=512)
    wall = 1e3 * (time.perf_counter() - t0)
    sess = fit.session(B, 8, 512, True)
    print(json.dumps({'B': B, 'wall_ms': round(wall, 2), 'ranges': getattr(sess, 'ranges', None),
                      'per_part_ms[start,inputs_up,fit_done,results_down]': getattr(fit, 'last_e2e_timeline', None)}), flush=True)
