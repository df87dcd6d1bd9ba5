"""Short driver for ncu: builds the bench workload and runs a few fit iterations.
  ncu ... python tools/profile_step.py [--frames 10000] [--iters 3] [--dense]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402

import bench  # noqa: E402
from bodyfitting_b200 import synthetic as syn  # noqa: E402
from bodyfitting_b200.engine import pack_cameras, pack_keypoints  # noqa: E402
from bodyfitting_b200.smplify.smplify import SMPLify  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--frames', type=int, default=10000)
ap.add_argument('--iters', type=int, default=3)
ap.add_argument('--dense', action='store_true', help='also run the SMPL 1024-frame dense LBS operator fwd/bwd')
a = ap.parse_args()
# one batch, one stream, direct launches (no graph): every kernel of the fit appears as its own launch in the ncu list
fit = SMPLify(smpl_type='smplx', num_iters=a.iters, gender='neutral', model_data=syn.make_model('smplx', 0), gmm=syn.make_gmm(0),
              concurrent_parts=1, graph=False)
pm = fit.model
wl = bench.build_workload(pm, a.frames, seed=100)
sess = fit.session(a.frames, 8, 512, True)
T = lambda k: torch.from_numpy(wl[k]).cuda()
sess.load_inputs(T('kp'), torch.from_numpy(pack_cameras(wl['c2ws'], wl['Ks'])).cuda(), T('init_pose'), T('init_betas'))
for _ in range(2):
    sess.run()
torch.cuda.synchronize()
if a.dense:
    print(bench.dense_lbs_bench(0, 6457.1))
print('done')
