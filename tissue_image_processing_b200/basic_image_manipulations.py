"""The slice of reference basic_image_manipulations.py that the projection path touches,
B200-backed: ``blur_image`` (BIM:373-390), ``put_channel_axis_first`` (BIM:199-231) and the chunked
operator host ``read_image_in_chunks`` (BIM:89-159).

File formats: ``open_image`` is a hook.  By default it imports aicsimageio + Bio-Formats like the reference
(BIM:79-87); where that stack is absent, plain TIFF / BigTIFF files are read by ``tiff_io.TiffImage`` (same surface:
``dims.{T,C,Z,Y,X}``, ``set_scene``, ``get_image_dask_data()`` -> sliceable with ``.compute()``, ``metadata``), every
other format needs the hook.  Tests install an in-memory image object with that surface.
"""
from __future__ import annotations

import numpy as np

from . import _native


def _default_open_image(path):
    try:
        from aicsimageio import AICSImage                   # noqa: WPS433 (optional dependency)
        from aicsimageio.readers import bioformats_reader
    except ImportError as exc:
        from . import tiff_io
        if tiff_io.is_tiff_path(path):
            return tiff_io.open_tiff(path)
        raise ImportError("reading %r needs aicsimageio + Bio-Formats (only TIFF files are read without them): "
                          "install them or set basic_image_manipulations.open_image" % (path,)) from exc
    return AICSImage(path, reader=bioformats_reader.BioformatsReader)


open_image = _default_open_image      # replaceable hook: callable(path) -> image object


def get_image_dimensions(path, series=0):
    """BIM:79-82."""
    img = open_image(path)
    img.set_scene(series)
    return img.dims


def get_image_metadata(path, series=0):
    """BIM:84-87."""
    img = open_image(path)
    img.set_scene(series)
    return img.metadata


def put_channel_axis_first(image, axes):
    """BIM:199-231: (C, [T], [Z], X, Y) - X before Y - and only when C exists and is not first."""
    c = axes.find("C")
    if c <= 0:
        return image, tuple(np.arange(len(axes)))
    order = (c,)
    for name in ("T", "Z"):
        pos = axes.find(name)
        if pos >= 0:
            order += (pos,)
    order += (axes.find("X"), axes.find("Y"))
    return np.transpose(image, axes=order), order


def blur_image(image, std, fp64_accumulate=True):
    """BIM:373-390 on the GPU: ``gaussian_filter(image, std, mode='nearest')`` for 3-D float32 or
    uint16 arrays (the dtypes the projection path blurs).  ``fp64_accumulate`` (default) reproduces
    scipy's float64 line sums bit for bit; False uses fp32 FMA accumulation."""
    import torch
    image = np.asarray(image)
    if image.ndim != 3 or len(std) != 3:
        raise RuntimeError("sequence argument must have length equal to input rank")
    if image.dtype not in (np.float32, np.uint16):
        raise TypeError("blur_image on the B200 path supports float32 and uint16 volumes")
    dev = torch.from_numpy(np.ascontiguousarray(image)).cuda()
    out = _native.gaussian_blur(dev, std, fp64_accumulate=fp64_accumulate)
    return out.cpu().numpy()


def read_image_in_chunks(path, series=0, dx=0, dy=0, dz=0, dc=0, dt=0, apply_function=None, output=None,
                         **apply_function_params):
    """BIM:89-159: walk the image in (t, c, z, y, x) chunks, call ``apply_function(chunk, **params)``
    on each and scatter tuple results into ``output``.  Chunks are independent (no halo), exactly
    like the reference - XY tiles therefore carry their own percentile and edge handling.
    As in the reference (SURVEY trap T10) a non-tuple result re-wraps ``output`` on every chunk."""
    img = open_image(path)
    img.set_scene(series)
    data = img.get_image_dask_data()
    limits = {"t": img.dims.T, "c": img.dims.C, "z": img.dims.Z, "y": img.dims.Y, "x": img.dims.X}
    step = {"t": dt or limits["t"], "c": dc or limits["c"], "z": dz or limits["z"],
            "y": dy or limits["y"], "x": dx or limits["x"]}
    for t in range(0, limits["t"], step["t"]):
        for c in range(0, limits["c"], step["c"]):
            for z in range(0, limits["z"], step["z"]):
                for y in range(0, limits["y"], step["y"]):
                    for x in range(0, limits["x"], step["x"]):
                        lo = (t, c, z, y, x)
                        hi = tuple(min(a + step[k], limits[k]) for a, k in zip(lo, "tczyx"))
                        chunk = data[tuple(slice(a, b) for a, b in zip(lo, hi))]
                        chunk = chunk.compute()
                        if apply_function is None:
                            yield chunk
                            continue
                        result = apply_function(chunk, **apply_function_params)
                        if output is not None:
                            deflate = not isinstance(result, tuple)
                            if deflate:
                                result = [result]
                                output = [output]
                            for i in range(len(result)):
                                shape = output[i].shape
                                sl = tuple(slice(min(a, s), min(b, s)) for a, b, s in zip(lo, hi, shape))
                                output[i][sl] = result[i].reshape(tuple(s.stop - s.start for s in sl))
                            if deflate:
                                result = result[0]
                            yield result
    return
