"""Experiment: two half-batches fitted concurrently on two streams vs one full batch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from bodyfitting_b200 import synthetic as syn
from bodyfitting_b200.engine import FitSession, pack_cameras, pack_keypoints
from bodyfitting_b200.smplify.smplify import SMPLify
F = 10000
fit = SMPLify(smpl_type='smplx', num_iters=100, gender='neutral', model_data=syn.make_model('smplx', 0), gmm=syn.make_gmm(0))
pm = fit.model
wl = bench.build_workload(pm, F, seed=100)
kp = pack_keypoints(torch.from_numpy(wl['kp']).cuda(), True)
cams = torch.from_numpy(pack_cameras(wl['c2ws'], wl['Ks'])).cuda()
poses = torch.from_numpy(wl['init_pose']).cuda()
theta0 = pm.pack_theta(poses[:, :3], poses[:, 3:3 + pm.nbody], torch.from_numpy(wl['init_betas']).cuda())
def timeit(fn, reps=4):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
full = FitSession(pm, F, 8, 100); full.set_inputs(kp, cams)
print('one batch of %d: %.1f ms' % (F, timeit(lambda: full.run(theta0))))
ref = full.fb.t['theta'].clone()
for parts in (2, 3, 4):
    n = F // parts
    sess = [FitSession(pm, n, 8, 100) for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    for i, s in enumerate(sess): s.set_inputs(kp[i*n:(i+1)*n].contiguous(), cams)
    th = [theta0[i*n:(i+1)*n].contiguous() for i in range(parts)]
    def run():
        cur = torch.cuda.current_stream()
        for i in range(parts):
            streams[i].wait_stream(cur)
            with torch.cuda.stream(streams[i]):
                sess[i].run(th[i])
        for i in range(parts): cur.wait_stream(streams[i])
    print('%d concurrent parts of %d: %.1f ms' % (parts, n, timeit(run)), 'identical:', bool(torch.equal(torch.cat([s.fb.t['theta'] for s in sess]), ref[:parts*n])))
