"""Freeze outputs of the UNMODIFIED reference at the sizes the bench quotes (build container only).

    python -m oracle.make_golden_large [case ...]

TEST INFRASTRUCTURE ONLY.  For every case of ``oracle/large_cases.py`` this runs reference
surface_projection.py:17-85 (through ``oracle/reference_runner.py``) on the seeded input and stores, in
``tests/golden/large/<case>.npz``:

  zmap           uint8 (Y,X)      the reference's height map (int64 values < 256)
  zmap_sha256 / proj_sha256       digests of the int64 height map / of the projection as float32 bytes (the
                                  reference's float64 values are float32-exact, asserted) - the ``bitexact`` CUDA mode
                                  must reproduce both digests, which pins it to the reference over the WHOLE frame
  near_tie_bits  packed (Y,X)     oracle top-2 relative score gap <= 1e-4 (the pixels the north-star rule exempts)
  low_gap_index / low_gap_value   the pixels with gap <= 1e-3 and their gaps (float32), for diagnostics
  proj_rows / proj_row_index      the projection rows y % 8 == 0 plus the first / last 16 rows, float32, so that a digest
                                  mismatch can be localised and fast/exact can be compared with reference values
  proj_tile_sum  float64          sums of the projection over 64x64 tiles, per channel
  input_sha256                    digest of the generated input (the numpy generator must reproduce it on the GPU box)

The oracle is run next to the reference to get the score gap, and must agree with it bit for bit (asserted).
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

from . import large_cases, reference_runner
from . import surface_projection_oracle as orc

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "large")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def row_subset(Y):
    rows = set(range(0, Y, 8)) | set(range(min(16, Y))) | set(range(max(0, Y - 16), Y))
    return np.array(sorted(rows), dtype=np.int32)


def tile_sums(proj, tile=64):
    C, Y, X = proj.shape
    ty, tx = -(-Y // tile), -(-X // tile)
    pad = np.zeros((C, ty * tile, tx * tile))
    pad[:, :Y, :X] = proj
    return pad.reshape(C, ty, tile, tx, tile).sum(axis=(2, 4))


def freeze(name, build, kw):
    sp = reference_runner.load_surface_projection()
    t0 = time.time()
    chunk = build()
    t_in = time.time() - t0
    t0 = time.time()
    proj, zmap = sp.time_point_surface_projection(chunk, "TCZYX", **kw)
    t_ref = time.time() - t0
    assert proj.dtype == np.float64 and zmap.dtype == np.int64
    p32 = proj.astype(np.float32)
    assert np.array_equal(p32.astype(np.float64), proj)
    assert zmap.min() >= 0 and zmap.max() < 256
    # the oracle's score for the gap rule; the restatement must agree with the reference here too
    t0 = time.time()
    image = orc.prepare_image(chunk, "TCZYX", kw.get("airyscan", True), kw.get("min_z", 0), kw.get("max_z", 0))
    score = orc.focus_score(image, kw["reference_channel"])
    del image
    assert np.array_equal(kw.get("min_z", 0) + np.argmax(score, axis=0), zmap), "oracle and reference disagree"
    gap = orc.top2_relative_gap(score)
    del score
    t_gap = time.time() - t0
    low = np.flatnonzero(gap.ravel() <= 1e-3)
    rows = row_subset(zmap.shape[0])
    out = dict(zmap=zmap.astype(np.uint8), near_tie_bits=np.packbits(gap <= 1e-4, axis=None),
               low_gap_index=low.astype(np.int32), low_gap_value=gap.ravel()[low].astype(np.float32),
               proj_rows=p32[:, rows], proj_row_index=rows, proj_tile_sum=tile_sums(proj))
    meta = dict(name=name, kwargs=kw, shape=list(chunk.shape), zmap_sha256=sha(zmap), proj_sha256=sha(p32),
                input_sha256=sha(chunk), near_tie_pixels=int((gap <= 1e-4).sum()),
                zmap_min=int(zmap.min()), zmap_max=int(zmap.max()), numpy=np.__version__,
                seconds=dict(input=round(t_in, 1), reference=round(t_ref, 1), oracle_gap=round(t_gap, 1)))
    import scipy
    meta["scipy"] = scipy.__version__
    os.makedirs(OUT_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(OUT_DIR, name + ".npz"), **out)
    with open(os.path.join(OUT_DIR, name + ".json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("golden", name, meta["seconds"], "near ties", meta["near_tie_pixels"], flush=True)


def main(argv):
    want = set(argv)
    for name, build, kw in large_cases.LARGE_CASES:
        if want and name not in want:
            continue
        freeze(name, build, kw)


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
