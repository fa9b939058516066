"""warp3d_b200 -- B200-native (sm_100a) engine for WARP3D's element-block hot path."""
from .model import (Material, Model, State, build_structured_hex8, graded_stretch, new_state,  # noqa: F401
                    slab_domains, smooth_displacement)
