"""On-disk formats either side of the fitting path, same function names as the
reference's ``utils/io_utils.py``: ``load_openpose`` (:138-183), ``save_obj_mesh`` (:185-192),
``load_obj_mesh`` (:430-548, geometry only).  Host-side parsing, no arithmetic of the path."""
import json
import re

import numpy as np


def load_openpose(json_name, only_one=True):
    """OpenPose JSON -> {'pose': [25,3], 'hand_left': [21,3], 'hand_right': [21,3], 'face': [70,3]} of the
    person with the largest summed confidence (or the list of all persons); None if nobody was detected.
    Parts whose confidences are all zero are dropped, as in the reference."""
    with open(json_name, 'r') as fid:
        d = json.load(fid)
    people = d.get('people', [])
    if len(people) == 0:
        return None
    data = []
    for label in people:
        entry = {}
        for k, p in label.items():
            if 'keypoints' not in k:
                continue
            p = np.reshape(np.asarray(p, dtype=np.float64), -1)
            if len(p) == 0:
                continue
            dim = re.findall('([2-9]d)', k)
            dim = 2 if len(dim) == 0 else int(dim[-1][0])
            if len(p) % (dim + 1) == 0:
                p = p.reshape(-1, dim + 1)
                if np.abs(p[:, -1]).max() <= 0:
                    continue
            elif len(p) % dim == 0:
                p = p.reshape(-1, dim)
            else:
                p = p[:(len(p) // dim) * dim].reshape(-1, dim)
            entry[k.replace('_keypoints', '').replace('_%dd' % dim, '')] = p
        data.append(entry)
    data = [e for e in data if e]
    if len(data) == 0:
        return None
    if not only_one:
        return data
    scores = [sum(p[:, -1].sum() for p in e.values()) for e in data]
    return data[int(np.argmax(scores))]


def save_obj_mesh(mesh_path, verts, faces):
    with open(mesh_path, 'w') as f:
        for v in verts:
            f.write('v %.4f %.4f %.4f\n' % (v[0], v[1], v[2]))
        for t in faces:
            f.write('f %d %d %d\n' % (t[0] + 1, t[1] + 1, t[2] + 1))


def load_obj_mesh(mesh_file):
    """Wavefront OBJ -> (vertices [N,3] float64, faces [F,3] int, 0-based); polygons are fanned, v/vt/vn
    corner syntax accepted."""
    verts, faces = [], []
    with open(mesh_file, 'r') as f:
        for line in f:
            if line.startswith('v '):
                verts.append([float(x) for x in line.split()[1:4]])
            elif line.startswith('f '):
                idx = [int(tok.split('/')[0]) for tok in line.split()[1:]]
                idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    faces.append([idx[0], idx[k], idx[k + 1]])
    return np.array(verts, dtype=np.float64), np.array(faces, dtype=np.int64)


def compute_normal_torch(vertices, faces):
    """[utils/io_utils.py:410-428] per-vertex unit normals of a triangle mesh (unit face normals summed over the incident faces,
    renormalised with the reference's +1e-8), differentiable w.r.t. ``vertices`` (CUDA kernels, hand-written backward)."""
    from .. import ops
    return ops.vertex_normals(vertices, faces)
