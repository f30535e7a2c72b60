"""Freeze outputs of the UNMODIFIED reference into tests/golden/ (build container only).

    python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY.  Runs reference surface_projection.py:17-85 and
surface_proj_m.py:14-35 through ``oracle/reference_runner.py`` on the seeded inputs of
``oracle/golden_cases.py`` and stores: projection (float32 - the reference's float64 values are
float32-exact products, asserted below), height map (int32, or float64 for the manifold case
whose reference output can hold x.5 values), and for error cases the exception type name.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

from . import golden_cases, reference_runner
from . import surface_projection_oracle as orc

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    sp = reference_runner.load_surface_projection()
    spm = reference_runner.load_surface_proj_m(orc.block_reduce)
    os.makedirs(OUT_DIR, exist_ok=True)
    arrays, manifest = {}, {"numpy": np.__version__, "cases": {}}
    import scipy
    manifest["scipy"] = scipy.__version__
    for name, build, axes, kw in golden_cases.CASES:
        res = sp.time_point_surface_projection(build(), axes, **kw)
        proj, zmap = res if kw.get("z_map") else (res, None)
        assert proj.dtype == np.float64
        p32 = proj.astype(np.float32)
        assert np.array_equal(p32.astype(np.float64), proj), name
        arrays[name + "/projection"] = p32
        if zmap is not None:
            if np.array_equal(zmap, np.round(zmap)) and zmap.dtype.kind in "iu":
                arrays[name + "/zmap"] = zmap.astype(np.int32)
            else:
                arrays[name + "/zmap"] = np.asarray(zmap, dtype=np.float64)
        manifest["cases"][name] = {"axes": axes, "kwargs": kw, "proj_shape": list(proj.shape),
                                   "zmap_dtype": None if zmap is None else str(zmap.dtype)}
        print("golden", name, proj.shape, flush=True)
    for name, build, axes, kw in golden_cases.ERROR_CASES:
        try:
            sp.time_point_surface_projection(build(), axes, **kw)
        except Exception as exc:                                  # noqa: BLE001
            manifest["cases"][name] = {"axes": axes, "kwargs": kw, "raises": type(exc).__name__}
            print("golden", name, "raises", type(exc).__name__, flush=True)
        else:
            raise AssertionError("reference did not raise for " + name)
    for name, build, axes, kw in golden_cases.SPM_CASES:
        with contextlib.redirect_stdout(io.StringIO()):
            out = spm.surface_projection_m(build(), axes, **kw)
        arrays[name + "/projection"] = out
        manifest["cases"][name] = {"axes": axes, "kwargs": kw, "dtype": str(out.dtype)}
        print("golden", name, out.shape, out.dtype, flush=True)
    np.savez_compressed(os.path.join(OUT_DIR, "reference_outputs.npz"), **arrays)
    with open(os.path.join(OUT_DIR, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("wrote", OUT_DIR)


if __name__ == "__main__":
    sys.exit(main())
