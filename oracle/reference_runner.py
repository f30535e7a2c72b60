"""Import the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  The reference's I/O dependencies (tifffile, aicsimageio,
scikit-image, matplotlib) are not installed, so empty stand-in modules are registered in
``sys.modules`` before the import; none of them is touched by
``time_point_surface_projection`` with ``bin_size == 1``.  Nothing here is used on the GPU
box: ``/root/reference`` does not exist there and ``available()`` returns False.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TSP_REFERENCE_ROOT", "/root/reference")
_TOOL_DIR = os.path.join(REFERENCE_ROOT, "tissue_analyzing_tool")

_STUBS = {
    "tifffile": ("TiffFile", "imwrite", "TiffWriter"),
    "aicsimageio": ("AICSImage",),
    "aicsimageio.readers": ("czi_reader", "bioformats_reader"),
    "aicsimageio.writers": ("ome_tiff_writer",),
    "skimage": (),
    "skimage.exposure": ("adjust_gamma",),
    "skimage.segmentation": (),
    "skimage.filters": ("difference_of_gaussians", "threshold_local"),
    "skimage.transform": ("resize",),
    "skimage.registration": ("phase_cross_correlation",),
    "skimage.measure": ("block_reduce",),
    "matplotlib": (),
    "matplotlib.pyplot": (),
}


def available():
    return os.path.isfile(os.path.join(_TOOL_DIR, "surface_projection.py"))


def _install_stubs():
    for name, attrs in _STUBS.items():
        if name in sys.modules:
            continue
        try:
            __import__(name)
            continue
        except Exception:
            pass
        mod = types.ModuleType(name)
        for a in attrs:
            setattr(mod, a, None)
        sys.modules[name] = mod
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, mod)


def load_surface_projection():
    """Return the reference ``surface_projection`` module (unmodified source)."""
    if not available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    _install_stubs()
    sys.dont_write_bytecode = True
    if _TOOL_DIR not in sys.path:
        sys.path.insert(0, _TOOL_DIR)
    import surface_projection          # noqa: E402  (the reference module)
    return surface_projection


def load_surface_proj_m(block_reduce):
    """Return the reference ``surface_proj_m`` module with its ``put_cannel_axis_first`` typo
    aliased and ``block_reduce`` supplied (scikit-image is absent), as in SURVEY appendix B."""
    load_surface_projection()
    import surface_proj_m              # noqa: E402
    surface_proj_m.put_cannel_axis_first = surface_proj_m.put_channel_axis_first
    surface_proj_m.block_reduce = block_reduce
    return surface_proj_m
