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
fit = SMPLify(smpl_type='smplx', num_iters=a.iters, gender='neutral', model_data=syn.make_model('smplx', 0), gmm=syn.make_gmm(0), concurrent_parts=1)
pm = fit.model
wl = bench.build_workload(pm, a.frames, seed=100)
sess = fit.session(a.frames, 8, 512, True)
sess.set_inputs(pack_keypoints(torch.from_numpy(wl['kp']).cuda(), True), torch.from_numpy(pack_cameras(wl['c2ws'], wl['Ks'])).cuda())
poses = torch.from_numpy(wl['init_pose']).cuda()
theta0 = pm.pack_theta(poses[:, :3], poses[:, 3:3 + pm.nbody], torch.from_numpy(wl['init_betas']).cuda())
for _ in range(2):
    sess.run(theta0)
torch.cuda.synchronize()
if a.dense:
    print(bench.dense_lbs_bench(0, 6457.1))
print('done')
