"""Where the end-to-end gap goes: device-resident (equal-priority streams) vs the priority-stream schedule without copies vs the
full host->host call, 10,000 SMPL-X frames."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from bodyfitting_b200 import synthetic as syn
from bodyfitting_b200.engine import pack_cameras, pack_keypoints
from bodyfitting_b200.smplify.smplify import SMPLify

F = 10000
fit = SMPLify(smpl_type='smplx', num_iters=100, gender='neutral', model_data=syn.make_model('smplx', 0), gmm=syn.make_gmm(0))
pm = fit.model
wl = bench.build_workload(pm, F, seed=100)
args = ((wl['init_betas'], wl['init_pose']), list(wl['c2ws']), list(wl['Ks']), wl['kp'], None)
kp_dev = pack_keypoints(torch.from_numpy(wl['kp']).cuda(), True)
cams = torch.from_numpy(pack_cameras(wl['c2ws'], wl['Ks'])).cuda()
poses = torch.from_numpy(wl['init_pose']).cuda()
theta0 = pm.pack_theta(poses[:, :3], poses[:, 3:3 + pm.nbody], torch.from_numpy(wl['init_betas']).cuda())

def wall(fn, n=5, warm=2):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); a = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - a)
    return 1e3 * float(np.median(ts))

fit(*args, imsize=512)
sess = fit._sess
sess.set_inputs(kp_dev, cams)
print('device-resident, equal-priority streams: %.2f ms' % wall(lambda: sess.run(theta0)))
keep = sess.streams
sess.streams = sess.prio_streams
print('device-resident, priority streams:       %.2f ms' % wall(lambda: sess.run(theta0)))
sess.streams = keep
print('host->device inputs, device results (as_numpy=False): %.2f ms' % wall(lambda: fit(*args, imsize=512, as_numpy=False)))
print('host->host without vertices:             %.2f ms' % wall(lambda: fit(*args, imsize=512, return_vertices=False)))
fit._sess_key = None
print('host->host with vertices (the e2e line): %.2f ms' % wall(lambda: fit(*args, imsize=512)))
t0 = time.perf_counter(); ib, ip, kp = fit._pack_inputs(args[0], args[3]); t1 = time.perf_counter()
print('_pack_inputs (host) %.2f ms' % (1e3 * (t1 - t0)))
