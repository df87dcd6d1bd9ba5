"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Import shim named ``smplx`` so that the reference's own, unmodified files
(/root/reference/models/smpl.py:3-6, smplify/smplify.py:7,80) can be imported in the
authoring container, where the real third-party ``smplx`` package (pinned
smplx==0.1.13, requirements.txt:5) is not installed.  The arithmetic is the
restatement in ``oracle/smplx_port.py`` (parity vs the real package: unpinned).
Model tensors are read from ``<model_path>/<TYPE>_<GENDER>.npz`` files written by
``bodyfitting_b200.synthetic.write_data_dir``.
"""
import os

from oracle.smplx_port import SMPLLayer, SMPLXLayer
from . import lbs  # noqa: F401


def _find(model_path, model_type, gender):
    if os.path.isdir(model_path):
        base = os.path.basename(os.path.normpath(model_path))
        folder = model_path if base == model_type else os.path.join(model_path, model_type)
        for g in (str(gender).upper(), 'NEUTRAL'):
            fn = os.path.join(folder, '%s_%s.npz' % (model_type.upper(), g))
            if os.path.exists(fn):
                return fn
        raise FileNotFoundError('no %s model under %s' % (model_type, model_path))
    return model_path


class SMPL(SMPLLayer):
    def __init__(self, model_path, batch_size=1, gender='neutral', create_transl=True, joint_mapper=None,
                 dtype=None, age='adult', kid_template_path='', **ignored):
        import torch
        super().__init__(_find(model_path, 'smpl', gender), joint_mapper=joint_mapper,
                         create_transl=create_transl, batch_size=batch_size, dtype=dtype or torch.float32,
                         age=age, kid_template_path=kid_template_path)


class SMPLX(SMPLXLayer):
    def __init__(self, model_path, batch_size=1, gender='neutral', create_transl=True, joint_mapper=None,
                 use_face_contour=False, num_pca_comps=6, dtype=None, **ignored):
        import torch
        super().__init__(_find(model_path, 'smplx', gender), joint_mapper=joint_mapper,
                         use_face_contour=use_face_contour, num_pca_comps=num_pca_comps,
                         create_transl=create_transl, batch_size=batch_size, dtype=dtype or torch.float32)


def create(model_path, model_type='smpl', **kwargs):
    kwargs.pop('ext', None)
    if model_type.lower() == 'smpl':
        return SMPL(model_path, **kwargs)
    if model_type.lower() == 'smplx':
        return SMPLX(model_path, **kwargs)
    raise ValueError('unsupported model type %s' % model_type)
