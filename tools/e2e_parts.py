"""End-to-end time of SMPLify.__call__ (host numpy in -> host numpy out, 10,000 SMPL-X frames) for several part layouts:
(parts, min_part, lead) -> ms per call (median of 5) and the device-resident time of the same session."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from bodyfitting_b200 import synthetic as syn
from bodyfitting_b200.engine import pack_cameras, pack_keypoints
from bodyfitting_b200.smplify.smplify import SMPLify

F = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
fit = SMPLify(smpl_type='smplx', num_iters=100, gender='neutral', model_data=syn.make_model('smplx', 0), gmm=syn.make_gmm(0))
pm = fit.model
wl = bench.build_workload(pm, F, seed=100)
args = ((wl['init_betas'], wl['init_pose']), list(wl['c2ws']), list(wl['Ks']), wl['kp'], None)
kp_dev = pack_keypoints(torch.from_numpy(wl['kp']).cuda(), True)
cams = torch.from_numpy(pack_cameras(wl['c2ws'], wl['Ks'])).cuda()
poses = torch.from_numpy(wl['init_pose']).cuda()
theta0 = pm.pack_theta(poses[:, :3], poses[:, 3:3 + pm.nbody], torch.from_numpy(wl['init_betas']).cuda())

t0 = time.perf_counter()
pin = torch.empty(wl['kp'].shape, dtype=torch.float32, pin_memory=True)
t1 = time.perf_counter()
src = torch.from_numpy(wl['kp'])
for _ in range(3):
    t2 = time.perf_counter(); pin.copy_(src); t3 = time.perf_counter()
print('pinned alloc %.1f ms; host->pinned staging of %.0f MB: %.1f ms (%.1f GB/s)' % (1e3 * (t1 - t0), src.numel() * 4 / 1e6, 1e3 * (t3 - t2), src.numel() * 4 / (t3 - t2) / 1e9))

configs = [(4, 2048, 0, 0.5), (4, 2048, 0, 1.0), (4, 2048, 0, 1.5), (4, 2048, 0, 2.5), (3, 2048, 0, 1.0), (3, 2048, 0, 2.0), (5, 1024, 0, 1.5), (4, 2048, 0, 0.5)]
for parts, min_part, lead, taper in configs:
    fit.concurrent_parts, fit.concurrent_min_part, fit.concurrent_lead, fit.concurrent_taper = parts, min_part, lead, taper
    fit._sess_key = None
    for _ in range(2):
        fit(*args, imsize=512)
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); a = time.perf_counter()
        fit(*args, imsize=512)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - a)
    sess = fit._sess
    sess.set_inputs(kp_dev, cams)
    td = []
    for _ in range(4):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); sess.run(theta0); e.record(); torch.cuda.synchronize(); td.append(s.elapsed_time(e))
    print('parts %d min_part %d lead %d taper %.1f -> ranges %s | e2e %.2f ms (%.1fk frames/s) | device %.2f ms'
          % (parts, min_part, lead, taper, [hi - lo for lo, hi in getattr(sess, 'ranges', [(0, F)])], 1e3 * float(np.median(ts)),
             F / float(np.median(ts)) / 1e3, float(np.median(td[1:]))), flush=True)
