"""B200-native hot path of the Frenetix reactive planner.

Pure-host helpers import without a GPU; anything that touches the device goes through
``_capi.Handler`` which raises if libfrx_b200.so or a CUDA device is missing (no CPU fallback).
"""
__version__ = "0.1.0"

__all__ = ["ReactivePlannerB200", "TrajectoryBundle", "TrajectorySample", "CartesianSample", "CurviLinearSample",
           "SamplingHandler", "generate_sampling_matrix", "CoordinateSystem"]


def __getattr__(name):
    if name == "ReactivePlannerB200":
        from .reactive_planner_b200 import ReactivePlannerB200
        return ReactivePlannerB200
    if name in ("TrajectoryBundle", "TrajectorySample", "CartesianSample", "CurviLinearSample"):
        from . import trajectories
        return getattr(trajectories, name)
    if name in ("SamplingHandler", "generate_sampling_matrix"):
        from . import sampling_matrix
        return getattr(sampling_matrix, name)
    if name == "CoordinateSystem":
        from .coordinate_system import CoordinateSystem
        return CoordinateSystem
    raise AttributeError(name)
