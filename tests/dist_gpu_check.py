"""Multi-GPU invariance check, launched by torchrun (one rank per GPU):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_gpu_check.py
1. frame-sharded fit (no coupling): gathered parameters == single-GPU fit of all frames, bit for bit;
2. sequence fit with the temporal term, boundary rows by in-kernel NVLink peer stores (sharding.HaloLink), the whole
   coupled fit one CUDA graph per rank: gathered parameters == single-GPU fit of the whole sequence, bit for bit
   (two consecutive runs: the tick epoch carries over);
3. the same with the host-driven fallback (one NCCL send/recv pair per iteration, sharding.exchange_halo)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from bodyfitting_b200 import synthetic as syn                    # noqa: E402
from bodyfitting_b200.sharding import HaloLink, exchange_halo, frame_range, gather_frames   # noqa: E402
from bodyfitting_b200.smplify.smplify import SMPLify            # noqa: E402
from oracle import fit_port as fp                                # noqa: E402
from util import make_scene                                      # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
mt, nv, B, N = 'smplx', 8, 13, 20
model, gmm = syn.make_model(mt, 0), syn.make_gmm(0)
sc = make_scene(fp.FitPort(mt, model, gmm), mt, B, nv, seed=51)
lo, hi = frame_range(B, rank, world)
ok = True
link = HaloLink()
for name, w, extra, reps in (('plain', 0.0, {}, 1), ('temporal, NVLink halo, CUDA graph', 300.0, dict(halo=link), 2),
                             ('temporal, host NCCL halo', 300.0, dict(halo_exchange=exchange_halo), 1)):
    kw = dict(smpl_type=mt, num_iters=N, gender='neutral', model_data=model, gmm=gmm, temporal_weight=w)
    fit = SMPLify(**extra, **kw)
    for rep in range(reps):
        out = fit((sc['init_betas'][lo:hi], sc['init_pose'][lo:hi]), list(sc['c2ws']), list(sc['Ks']), sc['kp'][lo:hi], None,
                  imsize=512, as_numpy=False)
        sess = fit.session(hi - lo, nv, 512, True)
        theta = gather_frames(sess.theta.contiguous(), B)
        verts = gather_frames(out['vertices'].contiguous(), B)
        if rank == 0:
            one = SMPLify(**kw)
            ref = one((sc['init_betas'], sc['init_pose']), list(sc['c2ws']), list(sc['Ks']), sc['kp'], None, imsize=512, as_numpy=False)
            t_ref = one.session(B, nv, 512, True).theta
            same_t = bool(torch.equal(theta, t_ref))
            same_v = bool(torch.equal(verts, ref['vertices']))
            print('%s (run %d, graph %s): %d ranks, sharded == single-GPU  theta: %s  vertices: %s  (max |d theta| %.3e)'
                  % (name, rep, sess.use_graph, world, same_t, same_v, float((theta - t_ref).abs().max())))
            ok = ok and same_t and same_v
        dist.barrier()
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print('DIST CHECK', 'OK' if ok else 'FAILED')
    sys.exit(0 if ok else 1)
