"""Host-side logic of the drivers and of the frame partition, on CPU (no kernels run here)."""
import os
import sys

import numpy as np
import pytest

from oracle import reference_runner, synth
from oracle import surface_projection_oracle as orc
from tests.fake_image import FakeAICSImage, install

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _toy_operator(chunk, scale=1.0, z_map=True):
    """Cheap stand-in with the operator's return contract: (projection (C,Y,X) float64, zmap (Y,X) int64)."""
    a = chunk.reshape(chunk.shape[1:]).astype(np.float64)
    return (a.max(axis=1) * scale, a[0].argmax(axis=0).astype(np.int64))


@pytest.mark.parametrize("dy,dx", [(0, 0), (16, 16), (10, 24)])
def test_read_image_in_chunks_scatters_tiles_like_the_reference(monkeypatch, dy, dx):
    from tissue_image_processing_b200 import basic_image_manipulations as bim
    rng = np.random.default_rng(0)
    movie = rng.integers(0, 1000, size=(3, 2, 5, 32, 48)).astype(np.uint16)
    install(monkeypatch, {"f": FakeAICSImage([movie])})
    proj = np.zeros((3, 2, 1, 32, 48))
    zmap = np.zeros((3, 1, 1, 32, 48))
    n = sum(1 for _ in bim.read_image_in_chunks("f", dt=1, dy=dy, dx=dx, apply_function=_toy_operator,
                                               output=[proj, zmap], scale=2.0))
    ty = 1 if dy == 0 else -(-32 // dy)
    tx = 1 if dx == 0 else -(-48 // dx)
    assert n == 3 * ty * tx
    assert np.array_equal(proj[:, :, 0], movie.max(axis=2) * 2.0)
    assert np.array_equal(zmap[:, 0, 0], movie[:, 0].argmax(axis=1))
    if reference_runner.available():
        # same walk through the unmodified reference generator (BIM:89-159) with its reader swapped out
        reference_runner.load_surface_projection()
        import basic_image_manipulations as ref_bim
        monkeypatch.setattr(ref_bim, "AICSImage", lambda path, reader=None: FakeAICSImage([movie]))
        monkeypatch.setattr(ref_bim, "bioformats_reader", type("R", (), {"BioformatsReader": None}))
        rproj = np.zeros_like(proj)
        rzmap = np.zeros_like(zmap)
        rn = sum(1 for _ in ref_bim.read_image_in_chunks("f", dt=1, dy=dy, dx=dx, apply_function=_toy_operator,
                                                        output=[rproj, rzmap], scale=2.0))
        assert rn == n and np.array_equal(rproj, proj) and np.array_equal(rzmap, zmap)


def test_put_channel_axis_first_matches_reference_quirk():
    from tissue_image_processing_b200 import put_channel_axis_first
    a = np.zeros((4, 2, 5, 6))                       # Z C Y X
    out, order = put_channel_axis_first(a, "ZCYX")
    assert out.shape == (2, 4, 6, 5) and tuple(order) == (1, 0, 3, 2)      # C, Z, X, Y (X before Y)
    same, order = put_channel_axis_first(a, "CZYX")
    assert same is a and tuple(order) == (0, 1, 2, 3)
    o2, _ = orc.put_channel_axis_first(a, "ZCYX")
    assert o2.shape == out.shape


def test_concatenate_and_save_conventions(tmp_path):
    from tissue_image_processing_b200 import surface_projection as sp
    a = np.full((2, 2, 4, 4), 1000.7)
    b = np.full((1, 1, 4, 4), 70000.0)                 # fewer channels -> padded in front; uint16 wraps
    np.save(tmp_path / "a.npy", a)
    np.save(tmp_path / "b.npy", b)
    out = sp.concatenate_time_points([str(tmp_path / "a.npy"), str(tmp_path / "b.npy")])
    assert out.dtype == np.uint16 and out.shape == (3, 2, 4, 4)
    assert out[0, 0, 0, 0] == 1000 and out[2, 0, 0, 0] == 0 and out[2, 1, 0, 0] == np.float64(70000).astype("uint16")
    seen = {}
    sp.tiff_writer, old = (lambda path, image, axes, metadata: seen.update(path=path, image=image, axes=axes)), sp.tiff_writer
    try:
        sp.save_tiff("x.tif", np.array([[0.0, 0.5], [1.0, 2.0]]), axes="YX", data_type="uint16")
    finally:
        sp.tiff_writer = old
    assert seen["image"].dtype == np.uint16 and seen["image"].max() == 65535 and seen["image"][0, 1] == 16384


def test_frame_partition_is_a_partition():
    from tissue_image_processing_b200.movie import frame_owner
    for world in (1, 2, 3, 8):
        owners = [frame_owner(t, world) for t in range(50)]
        assert set(owners) == set(range(min(world, 50)))
        counts = np.bincount(owners, minlength=world)
        assert counts.max() - counts.min() <= 1


def _gloo_worker(rank, world, port, tmp, shared):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if not shared:
        os.environ["TSP_NO_SHARED_OUTPUTS"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.fake_image import FakeAICSImage, install
    from tissue_image_processing_b200 import basic_image_manipulations as bim
    from tissue_image_processing_b200.movie import FramePipeline
    movie = np.stack([synth.synth_stack(6, 24, 28, C=2, seed=3, t=t) for t in range(5)])      # (T,C,Z,Y,X)
    install(bim, {"m": FakeAICSImage([movie])})
    proj = np.zeros((5, 2, 1, 24, 28))
    zmap = np.zeros((5, 1, 1, 24, 28))
    pipe = FramePipeline(operator=orc.time_point_surface_projection)           # host-logic seam, no GPU here
    pipe.project_movie("m", 0, proj, zmap, reference_channel=0, airyscan=False, atoh_shift=0, min_z=0, max_z=0)
    from tissue_image_processing_b200.movie import SharedFrameCounter
    mine = list(SharedFrameCounter("test").claims(37))                         # every index exactly once across ranks
    np.save(os.path.join(tmp, "claims%d.npy" % rank), np.array(mine, dtype=np.int64))
    np.save(os.path.join(tmp, "proj%d.npy" % rank), proj)
    np.save(os.path.join(tmp, "zmap%d.npy" % rank), zmap)
    # the movie driver itself under two ranks: only rank 0 may touch the output directory (resume files, TIFF, zmap,
    # pickle, the final os.remove loop); the other rank only projects the time points it claims
    from tissue_image_processing_b200 import surface_projection as sp
    m2 = np.stack([synth.synth_stack(6, 24, 28, C=2, seed=4, t=t) for t in range(3)])
    install(bim, {"m1.czi": FakeAICSImage([movie, movie[::-1]]), "m2.czi": FakeAICSImage([m2])})
    written = {}
    sp.tiff_writer = lambda path, image, axes, metadata: written.update({path: (image, axes)})
    out = os.path.join(tmp, "driver")
    os.makedirs(out, exist_ok=True)
    pipe16 = FramePipeline(operator=orc.time_point_surface_projection, out_dtype="uint16")
    sp.movie_surface_projection(["m1.czi", "m2.czi"], 0, [2, 1], 2, out, "max_averages", 1, False, 0, 0, 0, False,
                                frame_pipeline=pipe16)
    np.save(os.path.join(tmp, "wrote%d.npy" % rank), np.array(sorted(os.path.basename(k) for k in written)))
    if rank == 0:
        np.save(os.path.join(tmp, "tif1.npy"), written[os.path.join(out, "position1.tif")][0])
        np.save(os.path.join(tmp, "tif2.npy"), written[os.path.join(out, "position2.tif")][0])
    # the job's output arrays: one mapping shared by the ranks of a single-host job (no assembly step), per-rank
    # arrays + gloo otherwise; nothing may stay behind in /dev/shm
    from tissue_image_processing_b200 import movie as mv
    outs = mv.allocate_outputs([((3, 4), np.uint16), ((2, 2), np.float64)])
    assert outs.shared == (shared and os.path.isdir("/dev/shm")), outs.shared
    a, b = outs.arrays
    assert a.dtype == np.uint16 and b.dtype == np.float64 and not a.any() and not b.any()
    dist.barrier(group=mv.host_group())                   # (everyone has seen the zeros)
    a[rank] = rank + 1                                    # both ranks write their own row ...
    assert mv.finish_outputs([a, b], [rank]) == (rank == 0 or outs.shared)
    if rank == 0:                                         # ... and rank 0 holds both afterwards, either way
        assert a[0, 0] == 1 and a[1, 0] == 2 and a[2, 0] == 0
    mv.host_group() and dist.barrier(group=mv.host_group())
    outs.close()
    # the tiled driver under two ranks: tiles claimed from the shared counter, assembled on rank 0
    big = synth.synth_stack(6, 40, 48, C=2, seed=9)[None]
    install(bim, {os.path.join(tmp, "big.tif"): FakeAICSImage([big])})
    open(os.path.join(tmp, "big.tif"), "a").close()
    written.clear()
    pipe_ref = FramePipeline(operator=orc.time_point_surface_projection)
    sp.large_image_projection(tmp, out, "big.tif", position=1, reference_channel=0, chunk_size=24, airyscan=False,
                              frame_pipeline=pipe_ref)
    if rank == 0:
        np.save(os.path.join(tmp, "big_tif.npy"), written[os.path.join(out, "big_projection.tif")][0])
    dist.barrier(group=mv.host_group())
    if rank == 0:
        left = [f for f in os.listdir("/dev/shm") if f.startswith("tsp_b200_out_")] if os.path.isdir("/dev/shm") else []
        assert left == [], left
    dist.destroy_process_group()


@pytest.mark.parametrize("shared", [True, False], ids=["shared_outputs", "gloo_assembly"])
def test_movie_partition_world_size_2_gloo(tmp_path, shared):
    """Two ranks each project the time points they claim from the shared counter; assembling the arrays is the
    only exchange (none at all when the job's output arrays are one shared mapping).  The operator is the oracle
    here - this checks the host logic, not the kernels."""
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + (7 if shared else 0)) % 2000
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path), shared), nprocs=2, join=True)
    movie = np.stack([synth.synth_stack(6, 24, 28, C=2, seed=3, t=t) for t in range(5)])
    claims = np.concatenate([np.load(tmp_path / ("claims%d.npy" % r)) for r in range(2)])
    assert sorted(claims.tolist()) == list(range(37))
    for rank in range(2):
        proj = np.load(tmp_path / ("proj%d.npy" % rank))
        zmap = np.load(tmp_path / ("zmap%d.npy" % rank))
        for t in range(5):
            want_p, want_z = orc.time_point_surface_projection(movie[t:t + 1], "TCZYX", 0, airyscan=False, z_map=True)
            assert np.array_equal(proj[t, :, 0], want_p), (rank, t)
            assert np.array_equal(zmap[t, 0, 0], want_z), (rank, t)
    # driver: position 1 lives in m1 (5 frames) and m2 (3 frames), position 2 only in m1 (series 1: the frames reversed)
    assert np.load(tmp_path / "wrote0.npy").tolist() == ["position1.tif", "position2.tif"]
    assert np.load(tmp_path / "wrote1.npy").size == 0
    m2 = np.stack([synth.synth_stack(6, 24, 28, C=2, seed=4, t=t) for t in range(3)])
    want1 = [orc.time_point_surface_projection(f[None], "TCZYX", 0, airyscan=False).astype("uint16")
             for f in list(movie) + list(m2)]
    want2 = [orc.time_point_surface_projection(f[None], "TCZYX", 0, airyscan=False).astype("uint16") for f in movie[::-1]]
    assert np.array_equal(np.load(tmp_path / "tif1.npy"), np.stack(want1))
    assert np.array_equal(np.load(tmp_path / "tif2.npy"), np.stack(want2))
    left = sorted(os.listdir(tmp_path / "driver"))
    assert left == ["big_zmap.npy", "stage_locations_position1.pkl", "stage_locations_position2.pkl",
                    "zmap_position1.npy", "zmap_position2.npy"], left
    # tiled driver: four independent 24 x 24 (x 20 / 24) tiles, each with its own percentile and edges (BIM:104-149)
    big = synth.synth_stack(6, 40, 48, C=2, seed=9)[None]
    want_zmap = np.zeros((40, 48))
    want_proj = np.zeros((2, 40, 48))
    for y0 in (0, 24):
        for x0 in (0, 24):
            p, z = orc.time_point_surface_projection(big[:, :, :, y0:y0 + 24, x0:x0 + 24], "TCZYX", 0, airyscan=False,
                                                     z_map=True)
            want_proj[:, y0:y0 + 24, x0:x0 + 24] = p
            want_zmap[y0:y0 + 24, x0:x0 + 24] = z
    assert np.array_equal(np.load(tmp_path / "driver" / "big_zmap.npy")[0], want_zmap)
    want_u16 = np.round(want_proj / want_proj.max() * 65535).astype("uint16")
    assert np.array_equal(np.load(tmp_path / "big_tif.npy"), want_u16)
    import pickle
    with open(tmp_path / "driver" / "stage_locations_position1.pkl", "rb") as f:
        assert len(pickle.load(f)["x"]) == 8
    assert np.load(tmp_path / "driver" / "zmap_position1.npy").shape == (8, 1, 1, 24, 28)


def test_movie_driver_resumes_after_an_interrupted_run(monkeypatch, tmp_path):
    """SP:193-200: a (movie, position) job whose two resume files exist is not projected again.  The files of every
    job but the last are written as soon as the job is done (the last job's would be deleted a moment later by the
    clean-up, they are not written); a run that dies in the final TIFF write restarts at the last job only."""
    from tissue_image_processing_b200 import basic_image_manipulations as bim
    from tissue_image_processing_b200 import surface_projection as sp
    from tissue_image_processing_b200.movie import FramePipeline
    m1 = np.stack([synth.synth_stack(5, 16, 20, C=1, seed=1, t=t) for t in range(3)])
    m2 = np.stack([synth.synth_stack(5, 16, 20, C=1, seed=2, t=t) for t in range(2)])
    install(monkeypatch, {"m1.czi": FakeAICSImage([m1]), "m2.czi": FakeAICSImage([m2])})
    calls = []

    def operator(chunk, **kw):
        calls.append(int(chunk.sum()))
        return _toy_operator(chunk)

    def run(out, writer):
        monkeypatch.setattr(sp, "tiff_writer", writer)
        sp.movie_surface_projection(["m1.czi", "m2.czi"], 0, [2], 1, str(out), "max_averages", 1, False, 0, 0, 0, False,
                                    frame_pipeline=FramePipeline(operator=operator, out_dtype="uint16"))

    clean, broken = tmp_path / "clean", tmp_path / "broken"
    clean.mkdir()
    broken.mkdir()
    kept = {}
    run(clean, lambda path, image, axes, metadata: kept.update(clean=np.array(image)))
    assert len(calls) == 5 and kept["clean"].shape == (5, 1, 16, 20)

    def dies(path, image, axes, metadata):
        raise OSError("disk full")
    with pytest.raises(OSError):
        run(broken, dies)
    assert sorted(os.listdir(broken)) == ["position0_movie0_projection.npy", "position0_movie0_zmap.npy"]
    del calls[:]
    run(broken, lambda path, image, axes, metadata: kept.update(resumed=np.array(image)))
    assert len(calls) == 2                                   # only the two frames of the second movie
    assert np.array_equal(kept["resumed"], kept["clean"])
    assert np.array_equal(np.load(broken / "zmap_position1.npy"), np.load(clean / "zmap_position1.npy"))
    assert sorted(os.listdir(broken)) == sorted(os.listdir(clean)) == ["stage_locations_position1.pkl", "zmap_position1.npy"]


class _FakeNative:
    """Stands in for the C ABI under movie.FramePipeline._run_device: a frame slot computes a toy projection when the
    frame is waited for (so a buffer recycled too early shows up as a wrong result)."""
    MAX_SLOTS = 4
    METHODS = ("max_averages", "max_std", "multi_channel")
    PARAM_KEYS = ()

    def __init__(self):
        self.slots = {}
        self.submitted = []

    def pinned_empty(self, shape, dtype):
        return np.empty(shape, dtype=dtype)

    def bind_host_thread_to_gpu(self, device=None):
        return None

    def frame_submit(self, slot, stack, proj, zmap, **kw):
        assert slot not in self.slots, "slot re-used before it was waited for"
        self.slots[slot] = (stack, proj, zmap)
        self.submitted.append(int(stack[0, 0, 0, 0]))

    def frame_wait(self, slot, device=None):
        stack, proj, zmap = self.slots.pop(slot)
        proj[...] = stack.max(axis=1)
        zmap[...] = stack[0].argmax(axis=0)
        return {"band_index_error": False}


def test_frame_pipeline_feeder_keeps_order_recycles_buffers_and_propagates_errors(monkeypatch):
    """movie.FramePipeline._run_device without a GPU: the feeder thread stages frames ahead (slots + 2 buffers), the
    submitting thread keeps `slots` frames in flight; results arrive in order and intact, errors of the frame source
    and of the sink reach the caller, and nothing is left hanging."""
    import threading
    from tissue_image_processing_b200 import movie
    fake = _FakeNative()
    monkeypatch.setattr(movie, "_native", fake)
    monkeypatch.setattr(movie._Staging, "is_pinned", staticmethod(lambda arr: False))
    pipe = movie.FramePipeline.__new__(movie.FramePipeline)
    pipe.operator, pipe.mode, pipe.out_dtype, pipe.devices, pipe.slots, pipe.copy_threads, pipe.h2d_bytes = (
        None, "fast", "uint16", [0], 2, 3, 0)
    rng = np.random.default_rng(0)
    stacks = [rng.integers(0, 60000, size=(1, 4, 64, 96)).astype(np.uint16) for _ in range(11)]
    for t, st in enumerate(stacks):
        st[0, 0, 0, 0] = t                                   # tag
    got = []
    pipe._run_device(0, ((t, st) for t, st in enumerate(stacks)),
                     dict(reference_channel=0, airyscan=False), lambda k, p, z, s: got.append((k, p.copy(), z.copy())))
    assert [k for k, _, _ in got] == list(range(11)) and fake.submitted == list(range(11))
    for k, p, z in got:
        assert np.array_equal(p, stacks[k].max(axis=1).astype(np.uint16))
        assert np.array_equal(z, stacks[k][0].argmax(axis=0).astype(np.uint16))
    assert pipe.h2d_bytes == sum(st.nbytes for st in stacks) and not fake.slots

    def broken_source():
        yield 0, stacks[0]
        yield 1, stacks[1]
        raise OSError("reader failed")
    with pytest.raises(OSError, match="reader failed"):
        pipe._run_device(0, broken_source(), dict(reference_channel=0, airyscan=False), lambda *a: None)
    assert not fake.slots                                    # frames in flight were waited for

    def bad_sink(k, p, z, s):
        if k == 3:
            raise ValueError("sink failed")
    with pytest.raises(ValueError, match="sink failed"):
        pipe._run_device(0, ((t, st) for t, st in enumerate(stacks)), dict(reference_channel=0, airyscan=False), bad_sink)
    assert not fake.slots
    with pytest.raises(TypeError):                           # a float stack is refused by the feeder, seen by the caller
        pipe._run_device(0, iter([(0, stacks[0].astype(np.float32))]), dict(reference_channel=0, airyscan=False),
                         lambda *a: None)
    assert not [t for t in threading.enumerate() if t.name.startswith("tsp-feeder") and t.is_alive()]


def test_cli_flags_and_dispatch(monkeypatch, tmp_path):
    """SP:329-423: same flags and defaults, same dispatch to the three drivers."""
    import numpy as np
    from tissue_image_processing_b200 import surface_projection as sp
    opts, rest = sp.getOptions([])
    assert rest == []
    assert (opts.input, opts.output, opts.position_number, opts.movie_number, opts.reference_channel) == ("", "", 1, 1, 1)
    assert (opts.chunk_size, opts.method, opts.bin_size, opts.only_position, opts.zmin, opts.zmax) == (
        0, "max_averages", 1, 0, 0, 0)
    assert not (opts.fixed_sample or opts.build_manifold or opts.airyscan or opts.separate_files)
    calls = []
    monkeypatch.setattr(sp, "movie_surface_projection", lambda *a, **k: calls.append(("movie", a, k)))
    monkeypatch.setattr(sp, "large_image_projection", lambda *a, **k: calls.append(("large", a, k)))
    d = str(tmp_path)
    assert sp.main(["-i", d, "-n", "2", "-m", "3", "-r", "0", "-b", "2", "--method", "max_std", "--manifold",
                    "--min-z", "1", "--max-z", "9", "--airyscan"]) == 0
    kind, a, k = calls.pop()
    assert kind == "movie"
    assert a[0] == [os.path.join(d, "m%d.czi" % i) for i in (1, 2, 3)]
    assert a[1:] == (0, [3, 3], 2, d, "max_std", 2, True, 0, 1, 9, True)
    sp.main(["-i", d, "-n", "2", "-m", "3", "-f", "(2, 3)"])
    assert calls.pop()[1][2] == [2, 3]
    sp.main(["-i", d, "-o", d + "/out", "--fixed", "--file", "big.czi", "-c", "2048", "-n", "2"])
    kind, a, k = calls.pop()
    assert kind == "large" and a == (d, d + "/out", "big.czi")
    assert np.array_equal(k["position"], [1, 2]) and k["chunk_size"] == 2048 and k["airyscan"] is False
    open(os.path.join(d, "a.czi"), "w").close()
    sp.main(["-i", d, "--separate-files", "--only-position", "1"])
    kind, a, k = calls.pop()
    assert kind == "movie" and a[0] == [os.path.join(d, "a.czi")] and a[2] == (1,) and k["output_name"] == "a.czi"


def test_rank_to_device_order_interleaves_link_groups():
    """topology.interleaved_order: GPUs grouped by the host-link rate they reach, ranks dealt over the groups in turn
    (the measured 8 x B200 box: four GPUs at 23 GB/s, four at 36 GB/s)."""
    from tissue_image_processing_b200.topology import interleaved_order
    rates = [23.3, 23.3, 23.3, 23.4, 35.6, 35.6, 35.7, 35.5]
    assert interleaved_order(rates) == [4, 0, 5, 1, 6, 2, 7, 3]
    assert interleaved_order([54.0, 53.8, 54.1, 53.9]) == [0, 1, 2, 3]
    assert interleaved_order([50.0]) == [0]
    assert sorted(interleaved_order([10, 30, 20, 30, 10, 20])) == list(range(6))


def test_bench_movie_source_through_the_driver(tmp_path, monkeypatch):
    """bench.py's in-memory movie source (cycling pageable frames) drives movie_surface_projection like an image
    file would; the operator seam stands in for the GPU here."""
    import bench
    from tissue_image_processing_b200 import basic_image_manipulations as bim
    from tissue_image_processing_b200 import surface_projection as sp
    from tissue_image_processing_b200.movie import FramePipeline
    frames = [synth.synth_stack(6, 24, 32, seed=s) for s in range(3)]
    source = bench._CyclicMovie(frames, 7)
    monkeypatch.setattr(bim, "open_image", lambda path: source)
    written = {}
    monkeypatch.setattr(sp, "tiff_writer", lambda path, image, axes, metadata: written.update({path: image}))
    pipe = FramePipeline(operator=orc.time_point_surface_projection, out_dtype="uint16")
    sp.movie_surface_projection(["m.czi"], 0, [1], 1, str(tmp_path), "max_averages", 1, False, 0, 0, 0, False,
                                frame_pipeline=pipe)
    tif = written[os.path.join(str(tmp_path), "position1.tif")]
    assert tif.shape == (7, 1, 24, 32) and tif.dtype == np.uint16
    for t in range(7):
        want = orc.time_point_surface_projection(frames[t % 3][None], "TCZYX", 0, airyscan=False).astype("uint16")
        assert np.array_equal(tif[t], want)


def test_save_tiff_rescale_matches_the_reference_expression():
    """BIM:183-186 in pieces: same values as the one-line numpy expression, for sizes on both sides of the chunked
    path, uint8 and uint16 targets, and the degenerate inputs (all zero: whatever numpy makes of 0/0)."""
    import warnings
    from tissue_image_processing_b200 import surface_projection as sp
    rng = np.random.default_rng(12)
    seen = {}
    old, sp.tiff_writer = sp.tiff_writer, lambda path, image, axes, metadata: seen.update(image=image)
    try:
        for shape in ((2, 64, 64), (3, 1300, 1100)):
            img = (rng.random(shape, dtype=np.float32) * 2900).astype(np.float64)
            img[0, :7] = 0
            for data_type, top in (("uint16", 65535), ("uint8", 255)):
                sp.save_tiff("x.tif", img, axes="CYX", data_type=data_type)
                want = np.round((img / np.max(img)) * top).astype(data_type)
                assert seen["image"].dtype == want.dtype and np.array_equal(seen["image"], want), (shape, data_type)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            zeros = np.zeros((5, 1024, 1024))
            sp.save_tiff("x.tif", zeros, axes="CYX", data_type="uint16")
            assert np.array_equal(seen["image"], np.round((zeros / np.max(zeros)) * 65535).astype("uint16"))
        already = np.arange(12, dtype=np.uint16).reshape(3, 4)
        sp.save_tiff("x.tif", already, axes="YX", data_type="uint16")
        assert seen["image"] is already                               # no conversion when the dtype is the target
    finally:
        sp.tiff_writer = old
