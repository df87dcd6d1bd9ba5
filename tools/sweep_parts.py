"""Strong-scaling emulation on ONE GPU: fit B frames (what one of N ranks holds of a 10,000-frame sequence) as 1..k concurrent
parts, each one CUDA graph on its own stream, device-resident and end to end.  Usage:
    python tools/sweep_parts.py [B ...]        env: SWEEP_PARTS="1,2,3,4"  SWEEP_E2E=1
Prints one JSON line per (B, parts)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from bodyfitting_b200 import synthetic as syn
from bodyfitting_b200.engine import pack_cameras
from bodyfitting_b200.smplify.smplify import SMPLify

sizes = [int(a) for a in sys.argv[1:]] or [1250, 2500, 5000, 10000]
parts_list = [int(x) for x in os.environ.get('SWEEP_PARTS', '1,2,3,4').split(',')]
N = int(os.environ.get('SWEEP_ITERS', '100'))
model, gmm = syn.make_model('smplx', 0), syn.make_gmm(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for B in sizes:
    wl = None
    for parts in parts_list:
        fit = SMPLify(smpl_type='smplx', num_iters=N, gender='neutral', model_data=model, gmm=gmm, concurrent_parts=parts,
                      concurrent_min_part=max(128, B // (parts * 2)) if parts > 1 else 1 << 30)
        if wl is None:
            wl = bench.build_workload(fit.model, 10000, 100, 0, B)
            pin = {k: torch.from_numpy(wl[k]).pin_memory() for k in ('kp', 'init_pose', 'init_betas')}
        sess = fit.session(B, 8, 512, True)
        cams = torch.from_numpy(pack_cameras(wl['c2ws'], wl['Ks'])).cuda()
        sess.load_inputs(pin['kp'].cuda(), cams, pin['init_pose'].cuda(), pin['init_betas'].cuda())
        for _ in range(3):
            sess.run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); sess.run(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        out = {'B': B, 'parts': len(getattr(sess, 'ranges', [0])), 'graph': os.environ.get('BODYFIT_GRAPH', '1'),
               'ms_per_fit': float(np.median(ts)), 'frames_per_s': B / float(np.median(ts)) * 1e3}
        if os.environ.get('SWEEP_E2E', '1') != '0':
            host_args = ((pin['init_betas'].numpy(), pin['init_pose'].numpy()), list(wl['c2ws']), list(wl['Ks']), pin['kp'].numpy(), None)
            for _ in range(2):
                fit(*host_args, use_frames=list(range(8)), imsize=512)
            ws = []
            for _ in range(5):
                flush.zero_(); torch.cuda.synchronize()
                t0 = time.perf_counter(); fit(*host_args, use_frames=list(range(8)), imsize=512); ws.append(1e3 * (time.perf_counter() - t0))
            out.update(e2e_ms=float(np.median(ws)), e2e_frames_per_s=B / float(np.median(ws)) * 1e3)
        print(json.dumps(out), flush=True)
        del fit, sess
        torch.cuda.empty_cache()
