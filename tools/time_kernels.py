"""Per-kernel times of one fit iteration (CUDA events) -- prints one line; env vars select experiments."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import bench
from bodyfitting_b200 import synthetic as syn
from bodyfitting_b200.engine import FitSession, pack_cameras, pack_keypoints
from bodyfitting_b200.model import PreparedModel
F = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
pm = PreparedModel('smplx', syn.make_model('smplx', 0), gmm=syn.make_gmm(0), device='cuda')
wl = bench.build_workload(pm, F, seed=100)
sess = FitSession(pm, F, 8, 100)
sess.set_inputs(pack_keypoints(torch.from_numpy(wl['kp']).cuda(), True), torch.from_numpy(pack_cameras(wl['c2ws'], wl['Ks'])).cuda())
poses = torch.from_numpy(wl['init_pose']).cuda()
theta0 = pm.pack_theta(poses[:, :3], poses[:, 3:3 + pm.nbody], torch.from_numpy(wl['init_betas']).cuda())
sess.run(theta0)
k = bench.kernel_breakdown(pm, sess, F, 6457.1)
print(' | '.join('%s %.4f' % (x['kernel'], x['ms']) for x in k), '| sum %.4f' % sum(x['ms'] * x['launches_per_iteration'] for x in k))
