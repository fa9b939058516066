"""warp3d_b200 -- B200-native (sm_100a) engine for WARP3D's element-block hot path."""
from .model import (Material, Model, State, attach_crystals, cyclic_fa_material, cyclic_gp_material, build_hex8_model, build_structured_hex8, graded_stretch,  # noqa: F401
                    make_crystal_type, mm10_hist_size, new_state, slab_domains, smooth_displacement)
from .decks import seb_bend_field, seb_model, test_1_model  # noqa: F401
