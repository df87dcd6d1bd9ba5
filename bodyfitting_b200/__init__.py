"""B200-native (sm_100a) multi-view SMPLify fitting core.

Drop-in for the fitting path of generalizable-neural-performer/bodyfitting: the
sub-packages mirror the reference's module names (models.smpl, smplify.smplify,
smplify.loss, smplify.prior, utils.mesh_grid_searcher); all arithmetic runs in
hand-written CUDA kernels behind the C ABI in include/bodyfit_b200.h.
"""
__version__ = '0.1.0'
