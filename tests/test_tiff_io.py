"""The package's own TIFF reader / writer (tiff_io.py) and the drivers running on real files through it, on CPU.
Pillow (an independent TIFF implementation) is the cross-check in both directions."""
import os
import struct

import numpy as np
import pytest

from oracle import synth
from oracle import surface_projection_oracle as orc

PIL_Image = pytest.importorskip("PIL.Image")


def _pil_pages(path):
    im = PIL_Image.open(path)
    out = []
    for k in range(im.n_frames):
        im.seek(k)
        out.append(np.array(im))
    return np.stack(out)


@pytest.mark.parametrize("axes,shape,order", [("TCZYX", (3, 2, 4, 17, 23), "XYZCT"), ("TCYX", (5, 2, 9, 11), "XYCTZ"),
                                              ("CYX", (2, 8, 8), "XYCZT"), ("ZYX", (6, 5, 7), "XYZCT"),
                                              ("YX", (12, 10), "XYCZT")])
@pytest.mark.parametrize("dtype", ["uint8", "uint16", "float32"])
@pytest.mark.parametrize("bigtiff", [False, True])
def test_written_files_read_back_and_pillow_agrees(tmp_path, axes, shape, order, dtype, bigtiff):
    from tissue_image_processing_b200 import tiff_io
    rng = np.random.default_rng(len(axes))
    a = (rng.random(shape) * 250).astype(dtype)
    path = str(tmp_path / "a.tif")
    tiff_io.write_tiff(path, a, axes, bigtiff=bigtiff)
    with open(path, "rb") as f:
        assert struct.unpack("<2sH", f.read(4)) == (b"II", 43 if bigtiff else 42)
    img = tiff_io.TiffImage(path)
    assert img.dimension_order == order == tiff_io.dimension_order(axes)
    sizes = dict(zip("TCZYX", (1,) * 5), **dict(zip(axes, shape)))
    assert (img.dims.T, img.dims.C, img.dims.Z, img.dims.Y, img.dims.X) == tuple(sizes[k] for k in "TCZYX")
    full = img.get_image_dask_data().compute()
    assert full.dtype == a.dtype and np.array_equal(full, a.reshape(full.shape))
    assert np.array_equal(_pil_pages(path), a.reshape((-1,) + shape[-2:]))          # pages in C order of the array
    img.close() if full.flags.owndata else None


def test_lazy_indexing_matches_numpy_and_frames_are_views(tmp_path):
    from tissue_image_processing_b200 import tiff_io
    a = np.random.default_rng(1).integers(0, 65535, (4, 2, 5, 16, 24), dtype=np.uint16)
    path = str(tmp_path / "m.tif")
    tiff_io.write_tiff(path, a, "TCZYX")
    data = tiff_io.TiffImage(path).get_image_dask_data()
    assert data.shape == a.shape and data[1:3].shape == (2, 2, 5, 16, 24) and data[2, 1].shape == (5, 16, 24)
    for idx in [np.s_[1:2], np.s_[2], np.s_[:, 1], np.s_[1:2, :, :, 3:9, 5:20], np.s_[3, 0, ::2, ::-1, ::3],
                np.s_[..., 4:8, :], np.s_[0:1, :, :, 8:16, :], np.s_[1, :, 2:4], np.s_[0, 1, 4, 5, 6]]:
        assert np.array_equal(np.asarray(data[idx].compute()), a[idx]), idx
    assert np.array_equal(data[1:3][1][:, 2:].compute(), a[1:3][1][:, 2:])       # indexing composes
    frame = data[1:2].compute()                      # what project_movie reads per time point: no copy, read-only
    assert not frame.flags.owndata and not frame.flags.writeable
    tile = data[0:1, :, :, 0:8, 0:12].compute()      # an XY tile is gathered plane by plane
    assert tile.flags.writeable and np.array_equal(tile, a[0:1, :, :, 0:8, 0:12])
    with pytest.raises(TypeError):
        data[[0, 1]]
    with pytest.raises(IndexError):
        data[0, 0, 0, 0, 0, 0]


def test_files_written_by_pillow_strips_imagej_order_and_refusals(tmp_path):
    """Reader against an independent writer: an ImageJ hyperstack description (channels vary fastest), a file with no
    description, and the formats the reader refuses."""
    from tissue_image_processing_b200 import tiff_io
    rng = np.random.default_rng(2)
    T, Z, C, Y, X = 2, 3, 2, 300, 280
    a = rng.integers(0, 65535, (T, Z, C, Y, X), dtype=np.uint16)             # ImageJ order: T, Z, C
    pages = [PIL_Image.fromarray(p) for p in a.reshape(-1, Y, X)]
    path = str(tmp_path / "ij.tif")
    desc = "ImageJ=1.53\nimages=%d\nchannels=%d\nslices=%d\nframes=%d\nhyperstack=true\nspacing=0.5\n" % (T * Z * C, C, Z, T)
    pages[0].save(path, save_all=True, append_images=pages[1:], description=desc)
    img = tiff_io.TiffImage(path)
    assert img.dimension_order == "XYCZT" and img.shape5 == (T, C, Z, Y, X)
    assert img.metadata.images[0].pixels.physical_size_z == 0.5
    got = img.get_image_dask_data()
    assert np.array_equal(got.compute(), a.transpose(0, 2, 1, 3, 4))
    assert np.array_equal(got[1, :, 2, 10:50, 5:9].compute(), a[1, 2, :, 10:50, 5:9])
    # no description: the pages are the planes of one z stack
    plain = str(tmp_path / "plain.tif")
    pages[0].save(plain, save_all=True, append_images=pages[1:5])
    assert tiff_io.TiffImage(plain).shape5 == (1, 1, 5, Y, X)
    packed = str(tmp_path / "lzw.tif")
    pages[0].save(packed, compression="tiff_lzw")
    with pytest.raises(NotImplementedError, match="compressed"):
        tiff_io.TiffImage(packed)
    rgb = str(tmp_path / "rgb.tif")
    PIL_Image.fromarray(rng.integers(0, 255, (8, 8, 3), dtype=np.uint8)).save(rgb)
    with pytest.raises(NotImplementedError, match="samples per pixel"):
        tiff_io.TiffImage(rgb)
    junk = str(tmp_path / "junk.tif")
    with open(junk, "wb") as f:
        f.write(b"not a tiff at all")
    with pytest.raises(ValueError, match="not a TIFF"):
        tiff_io.TiffImage(junk)
    with pytest.raises(IndexError):
        img.set_scene(1)


def test_big_endian_file(tmp_path):
    """A hand-made 'MM' classic TIFF: two 2x3 uint16 pages."""
    from tissue_image_processing_b200 import tiff_io
    planes = [(np.arange(6).reshape(2, 3) + 256).astype(">u2"), (np.arange(6).reshape(2, 3) * 1000).astype(">u2")]

    def ifd(data_at, nxt):
        tags = [(256, 3, 1, 3), (257, 3, 1, 2), (258, 3, 1, 16), (259, 3, 1, 1), (262, 3, 1, 1), (273, 4, 1, data_at),
                (277, 3, 1, 1), (278, 3, 1, 2), (279, 4, 1, 12)]
        body = b"".join(struct.pack(">HHI", t, ty, n) + (struct.pack(">HH", v, 0) if ty == 3 else struct.pack(">I", v))
                        for t, ty, n, v in tags)
        return struct.pack(">H", len(tags)) + body + struct.pack(">I", nxt)

    size = 2 + 12 * 9 + 4
    blob = struct.pack(">2sHI", b"MM", 42, 8 + 24) + planes[0].tobytes() + planes[1].tobytes()
    blob += ifd(8, 8 + 24 + size) + ifd(20, 0)
    path = str(tmp_path / "mm.tif")
    with open(path, "wb") as f:
        f.write(blob)
    img = tiff_io.TiffImage(path)
    out = img.get_image_dask_data().compute()
    assert out.dtype == np.uint16 and out.dtype.isnative and out.shape == (1, 1, 2, 2, 3)
    assert np.array_equal(out[0, 0], np.stack(planes).astype(np.uint16))


def test_page_cut_into_strips_stored_out_of_order(tmp_path):
    """Two pages of 4 x 3 uint8, each cut into two strips of two rows; the strips of page 0 lie in the file in
    reverse order (not adjacent: the page is assembled), those of page 1 in order (adjacent: served as a view)."""
    from tissue_image_processing_b200 import tiff_io
    planes = np.arange(24, dtype=np.uint8).reshape(2, 4, 3) + 7

    def ifd(offsets_at, counts_at, nxt):
        tags = [(256, 3, 1, 3), (257, 3, 1, 4), (258, 3, 1, 8), (259, 3, 1, 1), (262, 3, 1, 1), (273, 4, 2, offsets_at),
                (277, 3, 1, 1), (278, 3, 1, 2), (279, 4, 2, counts_at)]
        body = b"".join(struct.pack("<HHI", t, ty, n) + (struct.pack("<HH", v, 0) if ty == 3 else struct.pack("<I", v))
                        for t, ty, n, v in tags)
        return struct.pack("<H", len(tags)) + body + struct.pack("<I", nxt)

    data_at = 8
    data = planes[0, 2:].tobytes() + planes[0, :2].tobytes() + planes[1].tobytes()           # 6 + 6 + 12 bytes
    tables_at = data_at + len(data)
    tables = struct.pack("<8I", data_at + 6, data_at, 6, 6, data_at + 12, data_at + 18, 6, 6)
    ifd_at = tables_at + len(tables)
    size = 2 + 12 * 9 + 4
    blob = struct.pack("<2sHI", b"II", 42, ifd_at) + data + tables
    blob += ifd(tables_at, tables_at + 8, ifd_at + size) + ifd(tables_at + 16, tables_at + 24, 0)
    path = str(tmp_path / "strips.tif")
    with open(path, "wb") as f:
        f.write(blob)
    img = tiff_io.TiffImage(path)
    assert [len(s) for s in img._strips] == [2, 2] and img._plane_at[0] is None and img._plane_at[1] == data_at + 12
    assert not img._packed
    assert np.array_equal(img.get_image_data()[0, 0], planes)
    assert np.array_equal(_pil_pages(path), planes)
    assert np.array_equal(img.get_image_dask_data()[0, 0, 1, 1:3].compute(), planes[1, 1:3])


def test_metadata_round_trip_and_dimension_order_rules(tmp_path):
    from types import SimpleNamespace as NS
    from tissue_image_processing_b200 import tiff_io
    meta = NS(images=[NS(name='position<3> "a&b"', pixels=NS(physical_size_x=0.25, physical_size_y=0.25,
                                                            physical_size_z=None))])
    path = str(tmp_path / "p.tif")
    tiff_io.hook_writer(path, np.zeros((2, 3, 4, 5), np.uint16), "TCYX", meta)
    im = tiff_io.TiffImage(path).metadata.images[0]
    assert im.pixels.physical_size_x == 0.25 and im.pixels.physical_size_z is None and im.pixels.type == "uint16"
    assert im.pixels.size_t == 2 and im.pixels.size_c == 3 and len(im.pixels.planes) == 6
    assert im.name.startswith("position")
    for bad in ("TCXY", "TTYX", "QYX", "YXC"):
        with pytest.raises(ValueError):
            tiff_io.dimension_order(bad)
    with pytest.raises(ValueError):
        tiff_io.write_tiff(path, np.zeros((2, 3, 4)), "TCYX")
    with pytest.raises(TypeError):
        tiff_io.write_tiff(path, np.zeros((2, 3), dtype=np.complex64))


def test_tiled_driver_on_a_real_tiff_with_the_default_reader_and_writer(tmp_path, monkeypatch):
    """large_image_projection (SP:279-316) from a .tif on disk to ``*_projection.tif`` / ``*_zmap.npy`` with no I/O
    hook installed: without aicsimageio the default ``open_image`` falls back to TiffImage and the default
    writer is tiff_io.  The operator seam runs the oracle - this is the host path, not the kernels."""
    pytest.importorskip("torch")
    from tissue_image_processing_b200 import basic_image_manipulations as bim
    from tissue_image_processing_b200 import surface_projection as sp
    from tissue_image_processing_b200 import tiff_io
    from tissue_image_processing_b200.movie import FramePipeline
    import sys
    for name in ("aicsimageio", "aicsimageio.readers"):       # absent here (other tests may have stubbed them)
        monkeypatch.setitem(sys.modules, name, None)
    monkeypatch.setattr(bim, "open_image", bim._default_open_image)
    monkeypatch.setattr(sp, "tiff_writer", sp._default_tiff_writer)
    big = synth.synth_stack(6, 40, 48, C=2, seed=9)[None]                  # (1, C, Z, Y, X)
    src, out = tmp_path / "in", tmp_path / "out"
    src.mkdir(), out.mkdir()
    tiff_io.write_tiff(str(src / "big.tif"), big, "TCZYX")
    pipe = FramePipeline(operator=orc.time_point_surface_projection)
    sp.large_image_projection(str(src), str(out), "big.tif", position=1, reference_channel=0, chunk_size=24,
                              airyscan=False, frame_pipeline=pipe)
    want_p = np.zeros((2, 40, 48))
    want_z = np.zeros((1, 40, 48))
    for y in (0, 24):
        for x in (0, 24):
            p, z = orc.time_point_surface_projection(big[:, :, :, y:y + 24, x:x + 24], "TCZYX", 0, airyscan=False,
                                                     z_map=True)
            want_p[:, y:y + 24, x:x + 24] = p
            want_z[0, y:y + 24, x:x + 24] = z
    want_u16 = np.round(want_p / want_p.max() * 65535).astype(np.uint16)          # BIM:183-186
    got = tiff_io.TiffImage(str(out / "big_projection.tif"))
    assert got.dimension_order == "XYCZT" and got.shape5 == (1, 2, 1, 40, 48) and got.dtype == np.uint16
    assert np.array_equal(got.get_image_data()[0, :, 0], want_u16)
    assert np.array_equal(_pil_pages(str(out / "big_projection.tif")), want_u16)
    assert np.array_equal(np.load(out / "big_zmap.npy"), want_z)
    with pytest.raises(ImportError, match="aicsimageio"):
        bim.open_image(str(src / "movie.czi"))


def test_movie_driver_writes_a_real_ome_tiff(tmp_path, monkeypatch):
    """movie_surface_projection with the default writer: ``position1.tif`` is a (T,C,Y,X) uint16 TIFF in XYCTZ plane
    order (SP:323) that both readers open, and it equals the truncated oracle projections (BIM:481)."""
    pytest.importorskip("torch")
    from tests.fake_image import FakeAICSImage, install
    from tissue_image_processing_b200 import surface_projection as sp
    from tissue_image_processing_b200 import tiff_io
    from tissue_image_processing_b200.movie import FramePipeline
    movie = np.stack([synth.synth_stack(6, 24, 28, C=2, seed=3, t=t) for t in range(3)])
    install(monkeypatch, {"m1.czi": FakeAICSImage([movie])})
    monkeypatch.setattr(sp, "tiff_writer", sp._default_tiff_writer)
    pipe = FramePipeline(operator=orc.time_point_surface_projection, out_dtype="uint16")
    sp.movie_surface_projection(["m1.czi"], 0, [1], 1, str(tmp_path), "max_averages", 1, False, 0, 0, 0, False,
                                frame_pipeline=pipe)
    want = np.stack([orc.time_point_surface_projection(movie[t:t + 1], "TCZYX", 0, airyscan=False).astype("uint16")
                     for t in range(3)])
    got = tiff_io.TiffImage(os.path.join(str(tmp_path), "position1.tif"))
    assert got.dimension_order == "XYCTZ" and got.shape5 == (3, 2, 1, 24, 28)
    assert got.metadata.images[0].name == "position0" and got.metadata.images[0].pixels.physical_size_x == 0.1
    assert np.array_equal(got.get_image_data()[:, :, 0], want)
    assert np.array_equal(_pil_pages(os.path.join(str(tmp_path), "position1.tif")), want.reshape(-1, 24, 28))


def test_classic_files_switch_to_bigtiff_at_the_offset_limit(tmp_path, monkeypatch):
    from tissue_image_processing_b200 import tiff_io
    a = np.random.default_rng(5).integers(0, 65535, (6, 64, 64), dtype=np.uint16)
    small, large = str(tmp_path / "s.tif"), str(tmp_path / "l.tif")
    tiff_io.write_tiff(small, a, "ZYX")
    monkeypatch.setattr(tiff_io, "_CLASSIC_LIMIT", 40000)            # stands for 4 GiB
    tiff_io.write_tiff(large, a, "ZYX")
    magic = [struct.unpack("<2sH", open(p, "rb").read(4))[1] for p in (small, large)]
    assert magic == [42, 43]
    assert np.array_equal(tiff_io.TiffImage(large).get_image_data()[0, 0], a)
    assert np.array_equal(_pil_pages(large), a)


def test_command_line_runs_on_tiff_movies(tmp_path, monkeypatch):
    """``python -m tissue_image_processing_b200.surface_projection -i DIR -m 2 -r 0`` on a directory that holds
    m1.tif / m2.tif instead of the reference's m1.czi / m2.czi (SP:413): default reader, default writer, the whole
    driver - only the GPU call is replaced by the oracle."""
    pytest.importorskip("torch")
    import sys
    from tissue_image_processing_b200 import basic_image_manipulations as bim
    from tissue_image_processing_b200 import surface_projection as sp
    from tissue_image_processing_b200 import tiff_io
    from tissue_image_processing_b200.movie import FramePipeline
    for name in ("aicsimageio", "aicsimageio.readers"):
        monkeypatch.setitem(sys.modules, name, None)
    monkeypatch.setattr(bim, "open_image", bim._default_open_image)
    monkeypatch.setattr(sp, "tiff_writer", sp._default_tiff_writer)
    monkeypatch.setattr(sp, "_default_pipeline", lambda mode, out_dtype: FramePipeline(
        operator=orc.time_point_surface_projection, out_dtype=out_dtype))
    movies = [np.stack([synth.synth_stack(6, 24, 28, C=1, seed=30 + k, t=t) for t in range(n)]) for k, n in ((0, 3), (1, 2))]
    for k, m in enumerate(movies):
        tiff_io.write_tiff(str(tmp_path / ("m%d.tif" % (k + 1))), m, "TCZYX")
    out = tmp_path / "out"
    out.mkdir()
    assert sp._movie_file(str(tmp_path), 3).endswith("m3.czi")              # nothing there: the reference's name
    assert sp.main(["-i", str(tmp_path), "-o", str(out), "-m", "2", "-r", "0"]) == 0
    got = tiff_io.TiffImage(str(out / "position1.tif"))
    assert got.shape5 == (5, 1, 1, 24, 28)
    frames = list(movies[0]) + list(movies[1])
    want = np.stack([orc.time_point_surface_projection(f[None], "TCZYX", 0, airyscan=False).astype("uint16")
                     for f in frames])
    assert np.array_equal(got.get_image_data()[:, :, 0], want)
    assert np.load(out / "zmap_position1.npy").shape == (5, 1, 1, 24, 28)
    assert sorted(os.listdir(out)) == ["position1.tif", "stage_locations_position1.pkl", "zmap_position1.npy"]


@pytest.mark.parametrize("shape,dtype", [((3, 1, 1, 5, 7), "uint16"), ((0, 4), "float64"), ((), "int64"),
                                         ((5, 1, 1, 1200, 1500), "uint16")])
def test_save_npy_writes_the_bytes_np_save_writes(tmp_path, shape, dtype):
    """The last case (18 MB) is cut into spans written by several threads."""
    from tissue_image_processing_b200 import tiff_io
    a = (np.random.default_rng(3).random(shape) * 60000).astype(dtype)
    tiff_io.save_npy(str(tmp_path / "a.npy"), a, threads=5)
    np.save(tmp_path / "b.npy", a)
    assert (tmp_path / "a.npy").read_bytes() == (tmp_path / "b.npy").read_bytes()
    f = np.asfortranarray(np.arange(12.0).reshape(3, 4))
    tiff_io.save_npy(str(tmp_path / "f.npy"), f)
    assert np.array_equal(np.load(tmp_path / "f.npy"), f)


def test_spans_cover_the_range_once():
    from tissue_image_processing_b200 import tiff_io
    for nbytes in (1, 65535, 65536, 65537, 10 ** 6, (1 << 24) + 13):
        for parts in (1, 2, 3, 8, 64):
            spans = tiff_io._spans(nbytes, parts)
            assert spans[0][0] == 0 and spans[-1][1] == nbytes and len(spans) <= parts
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and all(b > a for a, b in spans)
            assert all(a % 65536 == 0 for a, _ in spans)


def test_threaded_write_and_read_into(tmp_path):
    """An 18 MB movie: the writer's pixel block and ``read_into`` of a frame run on several threads; a frame is one
    run of bytes of the file (preadv into the caller's buffer), an XY tile is gathered plane by plane."""
    from concurrent.futures import ThreadPoolExecutor
    from tissue_image_processing_b200 import tiff_io
    a = np.random.default_rng(4).integers(0, 65535, (2, 1, 9, 1000, 1000), dtype=np.uint16)
    path = str(tmp_path / "m.tif")
    tiff_io.write_tiff(path, a, "TCZYX", threads=4)
    img = tiff_io.TiffImage(path)
    data = img.get_image_dask_data()
    assert np.array_equal(data.compute(), a) and np.array_equal(_pil_pages(path)[11], a[1, 0, 2])
    frame = data[1:2][0]
    assert frame.shape == (1, 9, 1000, 1000) and frame.dtype == np.uint16
    for kw in (dict(), dict(threads=3), dict(pool=ThreadPoolExecutor(4))):
        out = np.zeros(frame.shape, np.uint16)
        assert frame.read_into(out, **kw) is out and np.array_equal(out, a[1])
    tile = data[0, :, :, 100:228, 64:192]
    out = np.zeros(tile.shape, np.uint16)
    tile.read_into(out, threads=3)
    assert np.array_equal(out, a[0, :, :, 100:228, 64:192])
    for bad in (np.zeros((1, 9, 1000, 999), np.uint16), np.zeros(frame.shape, np.int16),
                np.zeros((1, 9, 1000, 2000), np.uint16)[..., ::2]):
        with pytest.raises(ValueError):
            frame.read_into(bad)
    img.close()


def test_pipeline_stages_tiff_frames_by_read_into(tmp_path, monkeypatch):
    """movie.FramePipeline.project_movie on a TIFF movie, the C ABI replaced by the fake of test_drivers_cpu: the
    frames reach the (here: ordinary-memory) staging buffers through ``read_into``, never as arrays; with
    TSP_TIFF_MMAP=1 they travel as read-only views and are copied by the staging threads.  Same results."""
    pytest.importorskip("torch")
    from tests.test_drivers_cpu import _FakeNative
    from tissue_image_processing_b200 import basic_image_manipulations as bim
    from tissue_image_processing_b200 import movie, tiff_io
    a = np.random.default_rng(6).integers(0, 60000, (7, 2, 4, 64, 96), dtype=np.uint16)
    path = str(tmp_path / "m.tif")
    tiff_io.write_tiff(path, a, "TCZYX")
    monkeypatch.setattr(bim, "open_image", tiff_io.TiffImage)
    monkeypatch.setattr(movie._Staging, "is_pinned", staticmethod(lambda arr: False))
    reads = []
    real = tiff_io._LazyPlanes.read_into
    monkeypatch.setattr(tiff_io._LazyPlanes, "read_into", lambda self, out, **kw: (reads.append(out.shape), real(self, out, **kw))[1])
    for env, n_reads in ((None, 7), ("1", 0)):
        fake = _FakeNative()
        monkeypatch.setattr(movie, "_native", fake)
        if env:
            monkeypatch.setenv("TSP_TIFF_MMAP", env)
        pipe = movie.FramePipeline.__new__(movie.FramePipeline)
        pipe.operator, pipe.mode, pipe.out_dtype, pipe.devices, pipe.slots, pipe.copy_threads, pipe.h2d_bytes = (
            None, "fast", "uint16", [0], 2, 3, 0)
        proj = np.zeros((7, 2, 1, 64, 96), np.uint16)
        zmap = np.zeros((7, 1, 1, 64, 96), np.uint16)
        del reads[:]
        owned = pipe.project_movie(path, 0, proj, zmap, reference_channel=0, airyscan=False)
        assert owned == list(range(7)) and len(reads) == n_reads and pipe.h2d_bytes == a.nbytes
        assert np.array_equal(proj[:, :, 0], a.max(axis=2)) and np.array_equal(zmap[:, 0, 0], a[:, 0].argmax(axis=1))


def test_regular_ifd_chains_parse_the_same_with_and_without_the_vectorised_pass(tmp_path, monkeypatch):
    from tissue_image_processing_b200 import tiff_io
    a = np.random.default_rng(8).integers(0, 65535, (3, 2, 4, 6, 5), dtype=np.uint16)
    files = []
    for big in (False, True):
        files.append(str(tmp_path / ("c%d.tif" % big)))
        tiff_io.write_tiff(files[-1], a, "TCZYX", bigtiff=big)
    # big-endian chain of six 2 x 3 pages, IFDs back to back behind the data
    planes = (np.arange(36).reshape(6, 2, 3) * 7).astype(">u2")
    size = 2 + 12 * 9 + 4
    ifd_at = 8 + planes.nbytes
    blob = struct.pack(">2sHI", b"MM", 42, ifd_at) + planes.tobytes()
    for k in range(6):
        tags = [(256, 3, 3), (257, 3, 2), (258, 3, 16), (259, 3, 1), (262, 3, 1), (273, 4, 8 + 12 * k), (277, 3, 1),
                (278, 3, 2), (279, 4, 12)]
        blob += struct.pack(">H", 9) + b"".join(struct.pack(">HHI", t, ty, 1) + (struct.pack(">HH", v, 0) if ty == 3
                                                else struct.pack(">I", v)) for t, ty, v in tags)
        blob += struct.pack(">I", ifd_at + size * (k + 1) if k < 5 else 0)
    files.append(str(tmp_path / "mm.tif"))
    with open(files[-1], "wb") as f:
        f.write(blob)
    calls = []
    real = tiff_io._regular_ifds
    monkeypatch.setattr(tiff_io, "_regular_ifds", lambda *a: (calls.append(1), real(*a))[1])
    fast = [tiff_io._parse_ifds(open(p, "rb").read()) for p in files]
    assert len(calls) == 3                                           # one vectorised pass per file took the rest
    monkeypatch.setattr(tiff_io, "_regular_ifds", lambda buf, bo, big, at, size, n: ([], at + size))
    slow = [tiff_io._parse_ifds(open(p, "rb").read()) for p in files]
    assert fast == slow and [len(f[0]) for f in fast] == [24, 24, 6]
    monkeypatch.setattr(tiff_io, "_regular_ifds", real)
    assert np.array_equal(tiff_io.TiffImage(files[2]).get_image_data()[0, 0], planes.astype(np.uint16))
    # a chain that stops being regular half way (page 4 of 6 has another width): the pass ends there
    broken = bytearray(blob)
    struct.pack_into(">H", broken, ifd_at + size * 4 + 2 + 8, 4)
    pages, _ = tiff_io._parse_ifds(bytes(broken))
    assert [p[256] for p in pages] == [3, 3, 3, 3, 4, 3]


def test_open_tiff_remembers_a_file_until_it_changes(tmp_path):
    from tissue_image_processing_b200 import tiff_io
    path = str(tmp_path / "m.tif")
    tiff_io.write_tiff(path, np.zeros((3, 4, 5), np.uint16), "ZYX")
    first = tiff_io.open_tiff(path)
    assert tiff_io.open_tiff(path) is first and first.shape5 == (1, 1, 3, 4, 5)
    tiff_io.write_tiff(path, np.zeros((7, 4, 5), np.uint16), "ZYX")
    second = tiff_io.open_tiff(path)
    assert second is not first and second.shape5 == (1, 1, 7, 4, 5)
    second.close()
    assert tiff_io.open_tiff(path) is not second                     # a closed image is not handed out again
    for k in range(6):                                               # the cache stays small
        other = str(tmp_path / ("o%d.tif" % k))
        tiff_io.write_tiff(other, np.zeros((2, 2), np.uint8))
        tiff_io.open_tiff(other)
    assert len(tiff_io._open_cache) <= 4


def test_tiled_driver_stages_tiff_tiles_by_read_into(tmp_path, monkeypatch):
    """large_image_projection over a TIFF with the fake C ABI: every tile reaches its staging buffer through
    ``read_into`` (planes gathered from the file mapping by the copy threads), ragged edge tiles included."""
    pytest.importorskip("torch")
    from tests.test_drivers_cpu import _FakeNative
    from tissue_image_processing_b200 import basic_image_manipulations as bim
    from tissue_image_processing_b200 import movie, tiff_io
    from tissue_image_processing_b200 import surface_projection as sp
    a = np.random.default_rng(7).integers(1, 60000, (2, 2, 5, 70, 100), dtype=np.uint16)
    tiff_io.write_tiff(str(tmp_path / "big.tif"), a, "TCZYX")
    monkeypatch.setattr(bim, "open_image", tiff_io.open_tiff)
    monkeypatch.setattr(movie._Staging, "is_pinned", staticmethod(lambda arr: False))
    monkeypatch.setattr(movie, "_native", _FakeNative())
    monkeypatch.setattr(tiff_io, "_BULK_MIN", 1 << 10)               # small tiles still take the threaded gather
    reads = []
    real = tiff_io._LazyPlanes.read_into
    monkeypatch.setattr(tiff_io._LazyPlanes, "read_into", lambda self, out, **kw: (reads.append(out.shape), real(self, out, **kw))[1])
    written = {}
    monkeypatch.setattr(sp, "tiff_writer", lambda path, image, axes, metadata: written.update({path: (image, axes)}))
    pipe = movie.FramePipeline.__new__(movie.FramePipeline)
    pipe.operator, pipe.mode, pipe.out_dtype, pipe.devices, pipe.slots, pipe.copy_threads, pipe.h2d_bytes = (
        None, "fast", "reference", [0], 2, 3, 0)
    sp.large_image_projection(str(tmp_path), str(tmp_path), "big.tif", position=1, chunk_size=48, frame_pipeline=pipe)
    assert sorted(reads) == sorted([(2, 5, 48, 48)] * 4 + [(2, 5, 48, 4)] * 2 + [(2, 5, 22, 48)] * 4 + [(2, 5, 22, 4)] * 2)
    tif, axes = written[str(tmp_path / "big_projection.tif")]
    want = a.max(axis=2).astype(np.float64)
    assert axes == "TCYX" and np.array_equal(tif, np.round(want / want.max() * 65535).astype(np.uint16))
    zmap = np.load(tmp_path / "big_zmap.npy")
    want_z = np.zeros((2, 70, 100))
    for y in (0, 48):
        for x in (0, 48, 96):
            want_z[:, y:y + 48, x:x + 48] = a[:, 0, :, y:y + 48, x:x + 48].argmax(axis=1)   # the fake: argmax of channel 0
    assert np.array_equal(zmap, want_z)


def _two_scene_file(path, same_dims=False):
    """A multi-image OME-TIFF: position 0 = (T=3, C=1, Z=4) in XYZCT order; position 1 = (T=2, C=2, Z=3) in XYCZT
    order (channels fastest) or - ``same_dims``, what the movie driver assumes of the positions of a file (SP:197) -
    (T=3, C=1, Z=4) in XYTZC order (time fastest); stage labels on both; all planes 24 x 28."""
    from tissue_image_processing_b200 import tiff_io
    s0 = np.stack([synth.synth_stack(4, 24, 28, C=1, seed=40, t=t) for t in range(3)])            # (3,1,4,24,28)
    if same_dims:
        s1 = np.stack([synth.synth_stack(4, 24, 28, C=1, seed=41, t=t) for t in range(3)])
        second, order1, (t1, c1, z1) = s1.transpose(1, 2, 0, 3, 4), "XYTZC", (3, 1, 4)
    else:
        s1 = np.stack([synth.synth_stack(3, 24, 28, C=2, seed=41, t=t) for t in range(2)])        # (2,2,3,24,28)
        second, order1, (t1, c1, z1) = s1.transpose(0, 2, 1, 3, 4), "XYCZT", (2, 2, 3)
    pages = np.concatenate([s0.reshape(-1, 24, 28), second.reshape(-1, 24, 28)])
    xml = ('<?xml version="1.0"?><OME xmlns="http://www.openmicroscopy.org/Schemas/OME/2016-06">'
           '<Image ID="Image:0" Name="left"><StageLabel Name="p0" X="10.5" XUnit="um" Y="-3" YUnit="um" Z="1.25"/>'
           '<Pixels ID="Pixels:0" DimensionOrder="XYZCT" Type="uint16" SizeX="28" SizeY="24" SizeZ="4" SizeC="1" SizeT="3" '
           'PhysicalSizeX="0.2" PhysicalSizeY="0.2" PhysicalSizeZ="0.7"><TiffData/></Pixels></Image>'
           '<Image ID="Image:1" Name="right"><StageLabel Name="p1" X="99" Y="7" Z="2"/>'
           '<Pixels ID="Pixels:1" DimensionOrder="%s" Type="uint16" SizeX="28" SizeY="24" SizeZ="%d" SizeC="%d" SizeT="%d">'
           '<TiffData/></Pixels></Image></OME>' % (order1, z1, c1, t1))
    tiff_io.write_tiff(path, pages, "ZYX", description=xml)
    return s0, s1


def test_multi_image_ome_tiff_scenes(tmp_path):
    from tissue_image_processing_b200 import tiff_io
    path = str(tmp_path / "two.tif")
    s0, s1 = _two_scene_file(path)
    img = tiff_io.TiffImage(path)
    assert img.scenes == (0, 1) and img.shape5 == (3, 1, 4, 24, 28) and img.dimension_order == "XYZCT"
    lazy0 = img.get_image_dask_data()
    img.set_scene(1)
    assert img.shape5 == (2, 2, 3, 24, 28) and img.dims.C == 2 and img.dimension_order == "XYCZT"
    lazy1 = img.get_image_dask_data()
    assert np.array_equal(lazy1.compute(), s1) and np.array_equal(lazy1[1, :, 2].compute(), s1[1, :, 2])
    assert lazy0.shape == s0.shape and np.array_equal(lazy0[1:2].compute(), s0[1:2])       # bound to the scene it came from
    out = np.zeros((1, 4, 24, 28), np.uint16)
    lazy0[2:3][0].read_into(out)
    assert np.array_equal(out, s0[2])
    meta = img.metadata
    assert [im.name for im in meta.images] == ["left", "right"]
    assert (meta.images[0].stage_label.x, meta.images[0].stage_label.y, meta.images[0].stage_label.z) == (10.5, -3.0, 1.25)
    assert meta.images[0].stage_label.x_unit == "um" and meta.images[1].stage_label.x_unit is None
    assert meta.images[0].pixels.physical_size_z == 0.7 and meta.images[1].pixels.size_t == 2
    with pytest.raises(IndexError):
        img.set_scene(2)
    # descriptions that do not account for every page fall back to one z stack
    tiff_io.write_tiff(path, np.zeros((5, 4, 4), np.uint8), "ZYX", description=(
        '<OME><Image ID="Image:0"><Pixels DimensionOrder="XYZCT" SizeX="4" SizeY="4" SizeZ="2" SizeC="1" SizeT="2"/>'
        '</Image></OME>'))
    assert tiff_io.TiffImage(path).shape5 == (1, 1, 5, 4, 4)


def test_movie_driver_over_the_positions_of_one_ome_tiff(tmp_path, monkeypatch):
    """movie_surface_projection with two initial positions = the two scenes of one OME-TIFF movie file: one output
    TIFF, height-map file and stage pickle per position (SP:240-276 reads the stage labels of every scene)."""
    pytest.importorskip("torch")
    import pickle
    from tissue_image_processing_b200 import basic_image_manipulations as bim
    from tissue_image_processing_b200 import surface_projection as sp
    from tissue_image_processing_b200 import tiff_io
    from tissue_image_processing_b200.movie import FramePipeline
    path = str(tmp_path / "m1.tif")
    scenes = _two_scene_file(path, same_dims=True)
    monkeypatch.setattr(bim, "open_image", tiff_io.open_tiff)
    monkeypatch.setattr(sp, "tiff_writer", sp._default_tiff_writer)
    out = tmp_path / "out"
    out.mkdir()
    pipe = FramePipeline(operator=orc.time_point_surface_projection, out_dtype="uint16")
    sp.movie_surface_projection([path], 0, [1, 1], 2, str(out), "max_averages", 1, False, 0, 0, 0, False,
                                frame_pipeline=pipe)
    for p, movie in enumerate(scenes):
        got = tiff_io.TiffImage(str(out / ("position%d.tif" % (p + 1))))
        want = np.stack([orc.time_point_surface_projection(movie[t:t + 1], "TCZYX", 0, airyscan=False).astype("uint16")
                         for t in range(len(movie))])
        assert np.array_equal(got.get_image_data()[:, :, 0], want), p
        with open(out / ("stage_locations_position%d.pkl" % (p + 1)), "rb") as f:
            stage = pickle.load(f)
        assert stage["x"] == [(10.5, 99.0)[p]] * len(movie) and stage["physical_size_x"] == (0.2, None)[p]


def test_imagej_big_endian_single_ifd_stack(tmp_path):
    """What Fiji writes for large hyperstacks: big-endian, ONE IFD whose description states ``images=N``, all planes
    back to back behind the first strip.  Frames are byte-swapped on the way into the caller's buffer."""
    from tissue_image_processing_b200 import tiff_io
    T, Z, C, Y, X = 2, 3, 2, 6, 5
    a = np.random.default_rng(9).integers(0, 65535, (T, Z, C, Y, X)).astype(">u2")           # ImageJ order: T, Z, C
    desc = ("ImageJ=1.53t\nimages=%d\nchannels=%d\nslices=%d\nframes=%d\nhyperstack=true\n" % (T * Z * C, C, Z, T)).encode() + b"\0"
    n_tags = 10
    ifd_at, desc_at = 8, 8 + 2 + 12 * n_tags + 4
    data_at = desc_at + len(desc) + (len(desc) & 1)
    tags = [(256, 3, 1, X), (257, 3, 1, Y), (258, 3, 1, 16), (259, 3, 1, 1), (262, 3, 1, 1), (270, 2, len(desc), desc_at),
            (273, 4, 1, data_at), (277, 3, 1, 1), (278, 3, 1, Y), (279, 4, 1, Y * X * 2)]
    body = b"".join(struct.pack(">HHI", t, ty, n) + (struct.pack(">HH", v, 0) if ty == 3 else struct.pack(">I", v))
                    for t, ty, n, v in tags)
    blob = struct.pack(">2sHI", b"MM", 42, ifd_at) + struct.pack(">H", n_tags) + body + struct.pack(">I", 0)
    blob += desc + b"\0" * (len(desc) & 1) + a.tobytes()
    path = str(tmp_path / "fiji.tif")
    with open(path, "wb") as f:
        f.write(blob)
    img = tiff_io.TiffImage(path)
    assert img.shape5 == (T, C, Z, Y, X) and img.dimension_order == "XYCZT" and img.dtype == np.dtype(">u2")
    want = a.astype(np.uint16).transpose(0, 2, 1, 3, 4)
    data = img.get_image_dask_data()
    got = data.compute()
    assert got.dtype.isnative and np.array_equal(got, want)
    out = np.zeros((C, Z, Y, X), np.uint16)
    data[1:2][0].read_into(out, threads=2)
    assert np.array_equal(out, want[1])
    with open(path, "r+b") as f:                          # a truncated file is refused, not read past its end
        f.truncate(len(blob) - 10)
    with pytest.raises(ValueError, match="does not fit"):
        tiff_io.TiffImage(path)


def _cli_rank(rank, world, port, src, out):
    """One rank of ``torchrun ... -m tissue_image_processing_b200.surface_projection``: the command line joins the job
    itself (movie.init_job); only the GPU call is replaced by the oracle."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_WORLD_SIZE=str(world))
    import torch.distributed as dist
    from tissue_image_processing_b200 import surface_projection as sp
    from tissue_image_processing_b200.movie import FramePipeline
    calls = []

    def operator(chunk, **kw):
        calls.append(1)
        return orc.time_point_surface_projection(chunk, **kw)
    sp._default_pipeline = lambda mode, out_dtype: FramePipeline(operator=operator, out_dtype=out_dtype)
    assert sp.main(["-i", src, "-o", out, "-m", "2", "-r", "0"]) == 0
    assert dist.is_initialized() and dist.get_world_size() == world and "gloo" in str(dist.get_backend())
    np.save(os.path.join(out, "calls%d.npy" % rank), np.array(len(calls)))
    dist.barrier()
    dist.destroy_process_group()


def test_command_line_under_a_two_rank_launch_shares_the_frames(tmp_path):
    pytest.importorskip("torch")
    import torch.multiprocessing as mp
    from tissue_image_processing_b200 import tiff_io
    movies = [np.stack([synth.synth_stack(6, 24, 28, C=1, seed=50 + k, t=t) for t in range(n)]) for k, n in ((0, 5), (1, 4))]
    src, out = tmp_path / "in", tmp_path / "out"
    src.mkdir(), out.mkdir()
    for k, m in enumerate(movies):
        tiff_io.write_tiff(str(src / ("m%d.tif" % (k + 1))), m, "TCZYX")
    port = 29500 + (os.getpid() * 11 + 3) % 2000
    mp.spawn(_cli_rank, args=(2, port, str(src), str(out)), nprocs=2, join=True)
    calls = [int(np.load(out / ("calls%d.npy" % r))) for r in range(2)]
    assert sum(calls) == 9, calls                       # every time point projected exactly once across the ranks
    got = tiff_io.TiffImage(str(out / "position1.tif"))
    frames = list(movies[0]) + list(movies[1])
    want = np.stack([orc.time_point_surface_projection(f[None], "TCZYX", 0, airyscan=False).astype("uint16")
                     for f in frames])
    assert np.array_equal(got.get_image_data()[:, :, 0], want)
    assert np.load(out / "zmap_position1.npy").shape == (9, 1, 1, 24, 28)
    assert sorted(f for f in os.listdir(out) if not f.startswith("calls")) == [
        "position1.tif", "stage_locations_position1.pkl", "zmap_position1.npy"]
