"""eslam_b200 -- Python host-side mirror of the reference's core/ API (Frame, PlaneExtractor, EllipsoidExtractor,
Initializer, Optimizer) over the C-ABI of libeslam_b200.so (include/eslam.h).  The library is CUDA-only: importing
works without a GPU (symbols are checked), creating a handle does not."""
from ._structs import (EDGE_9DOF, EDGE_GRAVITY, EDGE_PROJ, EslamCamera, EslamEdgeBatch, EslamEllipsoidOut, EslamLmStats,  # noqa: F401
                       EslamParams, EslamStageDump, default_camera, default_params)
from .api import (EllipsoidExtractor, EslamError, Frame, Handle, Initializer, Observation, Optimizer, PlaneExtractor, load_library,  # noqa: F401
                  calibrate_measurement, plane_to_world, pose_inverse, pose_mul)
