"""B200-native surface projection of confocal z-stacks and time-lapse movies.

Drop-in for the projection hot path of kasirershahartau/tissue_image_processing
(tissue_analyzing_tool/surface_projection.py, surface_proj_m.py): same function names and
arguments, CUDA kernels for sm_100a behind a C ABI (include/tsp_b200.h).  No CPU fallback.
"""
from .surface_projection import (build_continues_manifold, large_image_projection,  # noqa: F401
                                 movie_surface_projection, time_point_surface_projection)
from .surface_proj_m import surface_projection_m  # noqa: F401
from .basic_image_manipulations import blur_image, put_channel_axis_first, read_image_in_chunks  # noqa: F401

__all__ = ["time_point_surface_projection", "movie_surface_projection", "large_image_projection",
           "build_continues_manifold", "surface_projection_m", "blur_image",
           "put_channel_axis_first", "read_image_in_chunks"]
