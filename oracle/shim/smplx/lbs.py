"""ORACLE -- TEST INFRASTRUCTURE ONLY.  ``smplx.lbs`` names used by the reference
(models/smpl.py:6) mapped onto the restatement in oracle/smplx_port.py."""
from oracle.smplx_port import (batch_rodrigues, batch_rigid_transform, blend_shapes,  # noqa: F401
                               find_dynamic_lmk_idx_and_bcoords, lbs, vertices2joints, vertices2landmarks)
