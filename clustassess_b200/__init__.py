"""clustassess_b200 -- B200-native (sm_100a) implementation of ClustAssess's flat-disjoint
element-centric-similarity hot path, behind a C ABI (include/clustassess_b200.h).

    from clustassess_b200 import ecs
    ecs.element_consistency(list_of_membership_vectors)

`ecs` mirrors the R API of R/ECS.R; `_capi` is the ctypes binding; `synth` the seeded generators of
SURVEY.md section 8d.  No CPU fallback exists: the CUDA library must be built (python __graft_entry__.py)
and a B200 must be visible.
"""
from . import _capi, synth  # noqa: F401


def load_library():
    """Load libclustassess_b200.so and check that it exports every symbol the header declares."""
    return _capi.lib()


def __getattr__(name):
    if name == "ecs":
        import importlib

        return importlib.import_module(".ecs", __name__)
    raise AttributeError(name)
