"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Exact closest point on a triangle mesh by brute force (numpy, fp64) -- the referee for
bf_grid_nearest, following the recipe of thirdparty/mesh_grid/test_mesh_grid.py:24-34 (compare
against an exact implementation; trimesh itself is absent offline) -- and the restatement of the
SMPL+D objective: utils/io_utils.py:405-428 (compute_normal_torch), smplify/loss.py:233-242
(point_cloud_loss_mesh_grid), :260-271 (normal_loss_mesh_grid), :273-288 (normal_laplacian_smoothness),
smplify/smplify.py:228-247 (displacement loop).
"""
import numpy as np
import torch


def closest_points_bruteforce(points, verts, faces, chunk=256):
    """points [Q,3], verts [N,3], faces [F,3] -> (closest [Q,3], face [Q], dist2 [Q]) in fp64
    (Voronoi-region classification, Ericson, Real-Time Collision Detection 5.1.5)."""
    P = np.asarray(points, dtype=np.float64)
    V = np.asarray(verts, dtype=np.float64)
    Fc = np.asarray(faces, dtype=np.int64)
    a, b, c = V[Fc[:, 0]], V[Fc[:, 1]], V[Fc[:, 2]]
    ab, ac = b - a, c - a
    outp = np.zeros_like(P); outf = np.zeros(len(P), np.int64); outd = np.zeros(len(P))
    for s in range(0, len(P), chunk):
        p = P[s:s + chunk, None, :]                                   # [q,1,3]
        ap = p - a[None]
        d1 = (ab[None] * ap).sum(-1); d2 = (ac[None] * ap).sum(-1)
        bp = p - b[None]
        d3 = (ab[None] * bp).sum(-1); d4 = (ac[None] * bp).sum(-1)
        cp = p - c[None]
        d5 = (ab[None] * cp).sum(-1); d6 = (ac[None] * cp).sum(-1)
        vc = d1 * d4 - d3 * d2; vb = d5 * d2 - d1 * d6; va = d3 * d6 - d5 * d4
        u = np.zeros_like(d1); v = np.zeros_like(d1)
        done = np.zeros(d1.shape, bool)

        def put(mask, uu, vv):
            m = mask & ~done
            u[m] = uu[m] if isinstance(uu, np.ndarray) else uu
            v[m] = vv[m] if isinstance(vv, np.ndarray) else vv
            done[m] = True
        with np.errstate(divide='ignore', invalid='ignore'):
            put((d1 <= 0) & (d2 <= 0), 0.0, 0.0)
            put((d3 >= 0) & (d4 <= d3), 1.0, 0.0)
            put((vc <= 0) & (d1 >= 0) & (d3 <= 0), d1 / (d1 - d3), 0.0)
            put((d6 >= 0) & (d5 <= d6), 0.0, 1.0)
            put((vb <= 0) & (d2 >= 0) & (d6 <= 0), 0.0, d2 / (d2 - d6))
            w = (d4 - d3) / ((d4 - d3) + (d5 - d6))
            put((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0), 1.0 - w, w)
            den = 1.0 / (va + vb + vc)
            put(np.ones(d1.shape, bool), vb * den, vc * den)
        q = a[None] + u[..., None] * ab[None] + v[..., None] * ac[None]
        dist2 = ((p - q) ** 2).sum(-1)
        j = dist2.argmin(1)
        r = np.arange(len(j))
        outp[s:s + chunk] = q[r, j]; outf[s:s + chunk] = j; outd[s:s + chunk] = dist2[r, j]
    return outp, outf, outd


def compute_normal_torch(vertices, faces):
    """utils/io_utils.py:410-428 (index_add instead of three sparse COO products: same sums)."""
    vertices, faces = vertices.view(-1, 3), faces.view(-1, 3)
    va, vb, vc = vertices[faces[:, 0]], vertices[faces[:, 1]], vertices[faces[:, 2]]
    n = torch.cross(vb - va, vc - va, dim=1)
    n = n / (torch.norm(n, dim=-1, keepdim=True) + 1e-8)
    norm = torch.zeros_like(vertices)
    for j in range(3):
        norm = norm.index_add(0, faces[:, j], n)
    return norm / (torch.norm(norm, dim=-1, keepdim=True) + 1e-8)


def normal_laplacian_smoothness(norms, faces):
    mse = lambda x, y: torch.sum((x - y) ** 2, dim=-1)
    na, nb, nc = norms[faces[:, 0]], norms[faces[:, 1]], norms[faces[:, 2]]
    return torch.mean(mse(na, nb) + mse(nc, na) + mse(nb, nc))


def smpld_loop(body_vertices, body_faces, scan_verts, scan_faces, constant_scale, num_iters, dtype=torch.float64):
    """Displacement loop of smplify/smplify.py:228-247 with the exact brute-force closest point standing in
    for the CUDA-only mesh_grid module.  Returns (disp [V,3], per-iteration [icp, normal, smooth, loss])."""
    bv = torch.as_tensor(body_vertices, dtype=dtype)
    faces = torch.as_tensor(np.asarray(body_faces), dtype=torch.long)
    sv = np.asarray(scan_verts, dtype=np.float64); sf = np.asarray(scan_faces, dtype=np.int64)
    tris = sv[sf]
    face_norms = torch.as_tensor(np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]), dtype=dtype)
    disp = torch.zeros_like(bv, requires_grad=True)
    opt = torch.optim.Adam([disp], lr=5e-2, betas=(0.9, 0.999))
    trace = []
    for _ in range(num_iters):
        deformed = bv + disp
        norms = compute_normal_torch(deformed, faces)
        cp, cf, _ = closest_points_bruteforce(deformed.detach().numpy(), sv, sf)
        icp = torch.norm(deformed.view(-1, 3) - torch.as_tensor(cp, dtype=dtype), p=2)
        norm_loss = torch.mean(1 - torch.sum(face_norms[torch.as_tensor(cf)] * norms, dim=-1))
        smooth = normal_laplacian_smoothness(norms, faces)
        loss = icp + (norm_loss + smooth) * constant_scale * 0.1
        trace.append([float(icp), float(norm_loss), float(smooth), float(loss)])
        opt.zero_grad(); loss.backward(); opt.step()
    return disp.detach().numpy(), np.array(trace)


def inside_bruteforce(points, verts, faces, chunk=512):
    """fp64 referee for MeshGridSearcher.inside_mesh (utils/mesh_grid_searcher.py:86-91): generalised winding
    number (van Oosterom & Strackee solid angles) over ALL faces; returns (sign [+1 inside / -1 outside], winding)."""
    P = np.asarray(points, np.float64); V = np.asarray(verts, np.float64); F = np.asarray(faces, np.int64)
    a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    wn = np.zeros(len(P))
    for s in range(0, len(P), chunk):
        p = P[s:s + chunk, None, :]
        A, B, C = a[None] - p, b[None] - p, c[None] - p
        la, lb, lc = np.linalg.norm(A, axis=-1), np.linalg.norm(B, axis=-1), np.linalg.norm(C, axis=-1)
        num = np.einsum('qfi,qfi->qf', A, np.cross(B, C))
        den = la * lb * lc + (A * B).sum(-1) * lc + (B * C).sum(-1) * la + (C * A).sum(-1) * lb
        wn[s:s + chunk] = (2.0 * np.arctan2(num, den)).sum(1) / (4.0 * np.pi)
    return np.where(np.abs(wn) > 0.5, 1.0, -1.0), wn


def ray_any_bruteforce(origins, dirs, verts, faces, chunk=512):
    """fp64 referee for MeshGridSearcher.intersects_any (utils/mesh_grid_searcher.py:93-99; one-directional ray,
    t >= 0, thirdparty/mesh_grid/mesh_grid_kernel.cu:742-781): Moeller-Trumbore against ALL faces.
    Returns (hit [bool], margin): margin = how far the decisive triangle test is from flipping (small = grazing)."""
    O = np.asarray(origins, np.float64); D = np.asarray(dirs, np.float64)
    V = np.asarray(verts, np.float64); F = np.asarray(faces, np.int64)
    a = V[F[:, 0]]; e1 = V[F[:, 1]] - a; e2 = V[F[:, 2]] - a
    hit = np.zeros(len(O), bool); margin = np.zeros(len(O))
    for s in range(0, len(O), chunk):
        o, d = O[s:s + chunk, None, :], D[s:s + chunk, None, :]
        pv = np.cross(d, e2[None])
        det = (e1[None] * pv).sum(-1)
        with np.errstate(divide='ignore', invalid='ignore'):
            inv = 1.0 / det
            tv = o - a[None]
            u = (tv * pv).sum(-1) * inv
            qv = np.cross(tv, e1[None])
            v = (d * qv).sum(-1) * inv
            t = (e2[None] * qv).sum(-1) * inv
        m = np.minimum(np.minimum(u, v), np.minimum(1.0 - u - v, t))       # > 0: hit with that much slack
        m = np.where(np.isfinite(m), m, -np.inf)
        best = m.max(1)
        hit[s:s + chunk] = best >= 0
        margin[s:s + chunk] = np.abs(best)
    return hit, margin
