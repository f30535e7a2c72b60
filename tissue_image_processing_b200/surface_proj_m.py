"""B200 mirror of reference surface_proj_m.py (SPM:14-47, 81-100)."""
from __future__ import annotations

import numpy as np

from . import _native
from .basic_image_manipulations import put_channel_axis_first

METHODS = {"max_averages": 0, "max_std": 1}


def surface_projection_m(time_point, axes, reference_channel, min_z, max_z, method, bin_size):
    """Same signature as SPM:14.  The reference calls the misspelt ``put_cannel_axis_first`` and
    raises NameError as shipped; this mirror implements what the function means.  uint16 in,
    uint16 (rows, cols) out; the blurred stack keeps uint16 so each Gaussian pass truncates."""
    import torch
    image, _ = put_channel_axis_first(np.asarray(time_point), axes)
    stack = image[reference_channel][min_z:max_z]
    if stack.dtype != np.uint16:
        raise TypeError("surface_projection_m on the B200 path expects a uint16 stack")
    if stack.ndim != 3:
        raise RuntimeError("sequence argument must have length equal to input rank")
    if method not in METHODS:
        raise TypeError("exceptions must derive from BaseException")      # SPM:27 raises a str
    dev = torch.from_numpy(np.ascontiguousarray(stack)).cuda()
    out = _native.project_m(dev, METHODS[method], int(bin_size))
    return out.cpu().numpy()
