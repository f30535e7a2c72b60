"""GPU tests of ABI 3: the tsp_params block (the constants the reference hard-codes at SP:28/35/37/55/70-71 as
keyword-only arguments), uint16 outputs converted on the device, the host call's own frame slot, and the float32
rank arithmetic of np.percentile at the voxel count of the bench frame (SURVEY trap T1's own example)."""
import threading

import numpy as np
import pytest

from oracle import synth
from oracle import surface_projection_oracle as orc
from tests.parity import compare_frame

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tsp():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import tissue_image_processing_b200 as pkg
    return pkg


@pytest.fixture(scope="module")
def nat():
    from tissue_image_processing_b200 import _native
    _native.load_library()
    return _native


ORACLE_NAMES = {"percentile": "CLIP_PERCENTILE", "pedestal": "AIRYSCAN_PEDESTAL", "sigma_pre": "SIGMA_PRE",
                "sigma_score": "SIGMA_SCORE", "sigma_mask": "SIGMA_MASK"}

PARAM_CASES = [
    ("percentile_pedestal", dict(percentile=80, pedestal=500), dict(airyscan=True), (14, 96, 320)),
    ("percentile_99_5", dict(percentile=99.5), dict(airyscan=False), (12, 80, 264)),
    ("percentile_100", dict(percentile=100), dict(airyscan=False), (8, 64, 72)),
    ("wide_band", dict(sigma_mask=(3.0, 2.0, 2.0)), dict(airyscan=False, atoh_shift=1), (36, 96, 128)),
    ("flat_band", dict(sigma_mask=(0.0, 2.0, 2.0)), dict(airyscan=False), (10, 64, 96)),
    ("narrow_score", dict(sigma_score=(0.5, 12.0, 12.0), sigma_pre=(0.5, 1.5, 1.5)), dict(airyscan=False), (12, 96, 136)),
    ("everything", dict(percentile=90, pedestal=300, sigma_pre=(1.0, 1.0, 1.0), sigma_score=(1.0, 8.0, 20.0),
                        sigma_mask=(2.0, 1.0, 3.0)), dict(airyscan=True, atoh_shift=-1), (20, 72, 88)),
]


@pytest.mark.parametrize("mode", ["bitexact", "exact", "fast"])
@pytest.mark.parametrize("case", PARAM_CASES, ids=lambda c: c[0])
def test_params_against_oracle_with_the_same_constants(case, mode, tsp, monkeypatch):
    name, params, kw, (Z, Y, X) = case
    img = synth.synth_stack(Z, Y, X, C=2, seed=len(name), airyscan=False)[None]
    if kw.get("airyscan"):
        img = img + np.uint16(params.get("pedestal", 10000) - 150)        # some voxels end up below the pedestal
    for key, value in params.items():
        monkeypatch.setattr(orc, ORACLE_NAMES[key], value)
    okw = dict(reference_channel=0, z_map=True, **kw)
    (want_proj, want_zmap), score = orc.time_point_surface_projection(img, "TCZYX", return_score=True, **okw)
    got_proj, got_zmap = tsp.time_point_surface_projection(img, "TCZYX", mode=mode, **okw, **params)
    stats = compare_frame(got_proj, got_zmap, want_proj, want_zmap, orc.top2_relative_gap(score),
                          exact_zmap=(mode == "bitexact"))
    if mode == "bitexact":
        assert np.array_equal(got_proj, want_proj)
    print(name, mode, stats)


def test_params_validation(tsp):
    img = synth.synth_stack(6, 32, 40, seed=1)[None]
    with pytest.raises(ValueError):
        tsp.time_point_surface_projection(img, "TCZYX", 0, percentile=101)
    with pytest.raises(RuntimeError):
        tsp.time_point_surface_projection(img, "TCZYX", 0, sigma_mask=(1, 2))
    with pytest.raises(Exception):
        tsp.time_point_surface_projection(img, "TCZYX", 0, sigma_score=(0.5, -3, 30))


@pytest.mark.parametrize("q,ped", [(95, 0), (50, 0), (5, 0), (99.9, 0), (0, 0), (100, 0), (95, 700), (37.5, 1234)])
def test_percentile_any_q_matches_numpy(nat, q, ped):
    import torch
    rng = np.random.default_rng(int(q * 10) + ped)
    v = np.clip(rng.normal(1500, 400, size=9_000_001), 0, 65535).astype(np.uint16)
    v[::7] = 0
    st = nat.percentile_nonzero(torch.from_numpy(v).cuda(), q, ped)
    f = v.astype(np.float32) - ped
    f[f < 0] = 0
    nz = f[f > 0]
    assert st["nonzero_count"] == nz.size
    assert st["percentile95"] == np.percentile(nz, q), (q, ped)


def test_percentile_float32_rank_at_the_bench_voxel_count(nat):
    """n = 268 435 456 = 2048 x 2048 x 64: numpy's float32 virtual index is 255 013 680.0, the float64 one would be
    ...682.25 (rank quantum 16-32 at this magnitude).  A value step between the two ranks tells them apart."""
    import torch
    n = 268_435_456
    q = np.float32(95) / np.float32(100)
    vi = np.float32(n - 1) * q
    assert float(vi) == 255013680.0 and abs((n - 1) * 0.95 - 255013682.25) < 1e-3
    d = torch.full((n,), 1000, dtype=torch.int16, device="cuda")
    d[int(vi) + 1:] = 2000                                  # ranks above the float32 index
    st = nat.percentile95_nonzero(d.view(torch.uint16))
    assert st["nonzero_count"] == n
    assert st["percentile95"] == 1000.0                     # float64 ranks would give 2000
    # and against np.percentile itself on random data of that size (one float32 pass on the host)
    g = torch.Generator(device="cuda").manual_seed(5)
    r = torch.randint(1, 4000, (n,), device="cuda", generator=g, dtype=torch.int16)
    want = np.percentile(r.cpu().numpy().astype(np.float32), 95)
    assert nat.percentile95_nonzero(r.view(torch.uint16))["percentile95"] == want


def test_uint16_outputs_are_the_cast_of_the_reference_dtypes(nat):
    """TSP_FRAME_OUT_U16: what BIM:481 / SP:229-231 do on the host (astype('uint16')) happens on the device."""
    img = synth.synth_stack(12, 96, 264, C=2, seed=3)
    p64, z64, _ = nat.project_frame_host(img, 0, airyscan=False, atoh_shift=1, mode="fast")
    p16, z16, st = nat.project_frame_host(img, 0, airyscan=False, atoh_shift=1, mode="fast", out_u16=True)
    assert p16.dtype == np.uint16 and z16.dtype == np.uint16 and st["has_nonzero"]
    assert np.array_equal(p16, p64.astype("uint16")) and np.array_equal(z16, z64.astype("uint16"))
    assert (p64 != np.floor(p64)).any()                     # the cast really truncates something


def test_host_calls_from_several_threads_and_next_to_a_pipeline(tsp, nat):
    """tsp_project_frame_host has a frame slot of its own and takes turns: concurrent blocking calls (ctypes drops
    the GIL) and a slot pipeline running on the same handle do not disturb each other."""
    from tissue_image_processing_b200.movie import FramePipeline
    frames = [synth.synth_stack(10, 72, 264, seed=20 + i)[None] for i in range(4)]
    want = [tsp.time_point_surface_projection(f, "TCZYX", 0, airyscan=False, z_map=True) for f in frames]
    for rep in range(3):                # plain launches, graph capture, graph replay: all the same numbers
        again = [tsp.time_point_surface_projection(f, "TCZYX", 0, airyscan=False, z_map=True) for f in frames]
        for i, ((p0, z0), (p1, z1)) in enumerate(zip(want, again)):
            assert np.array_equal(z0, z1) and np.array_equal(p0, p1), (
                "repeat %d of frame %d differs: %d height-map pixels, max projection difference %g"
                % (rep, i, int((z0 != z1).sum()), float(np.abs(p0 - p1).max())))
    errors, results = [], {}

    def caller(k):
        try:
            for rep in range(6):
                i = (k + rep) % 4
                p, z = tsp.time_point_surface_projection(frames[i], "TCZYX", 0, airyscan=False, z_map=True)
                if not (np.array_equal(p, want[i][0]) and np.array_equal(z, want[i][1])):
                    errors.append((k, rep, i, int((z != want[i][1]).sum()), float(np.abs(p - want[i][0]).max())))
        except Exception as exc:                            # noqa: BLE001
            errors.append(exc)

    def piper():
        try:
            pipe = FramePipeline(slots=3)
            pipe.project_frames(((i, frames[i % 4][0]) for i in range(12)),
                                lambda i, p, z, st: results.__setitem__(i, (p.copy(), z.copy())),
                                reference_channel=0, airyscan=False)
        except Exception as exc:                            # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=caller, args=(k,)) for k in range(3)] + [threading.Thread(target=piper)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for i in range(12):
        assert np.array_equal(results[i][0], want[i % 4][0]) and np.array_equal(results[i][1], want[i % 4][1])
