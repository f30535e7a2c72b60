"""B200 mirror of reference tissue_analyzing_tool/surface_projection.py.

``time_point_surface_projection`` keeps the reference signature (SP:17-19) and return dtypes
(float64 projection, int64 height map) and runs the whole operator on the GPU through the C ABI
(``tsp_project_frame_host``).  The drivers ``movie_surface_projection`` (SP:168-237) and
``large_image_projection`` (SP:279-316) keep their signatures, file names, resume files and dtype
conventions, but are built around ``movie.FramePipeline``: a frame source (time points of a movie file, XY tiles
of a large image) feeds the GPU frame slots through pinned staging buffers, the results come back in the dtype
that is written to disk (uint16 converted on the device for movies).  File I/O goes through the hooks of
``basic_image_manipulations`` / ``tiff_writer`` (defaults: the reference's aicsimageio stack when installed, else the
package's own TIFF reader; the package's own OME-TIFF writer).

``bin_size > 1`` (SP:39-53, methods max_averages / max_std / multi_channel) and ``build_manifold`` (SP:87-165)
run on the GPU too; they use the direct-FIR score (``mode="fast"`` behaves like "exact" for them).

Extensions (keyword-only, defaults preserve reference behaviour):
  ``mode``        score stage: "fast" (default, multirate sigma=30 stage), "exact" (direct FIR, fp32) or "bitexact"
                  (direct FIR, scipy's float64 summation order; bit-identical height map and projection).  The
                  default can also be set with the environment variable TSP_MODE.
  ``percentile``, ``pedestal``, ``sigma_pre``, ``sigma_score``, ``sigma_mask``
                  the constants the reference hard-codes at SP:35, SP:28, SP:37, SP:55 and SP:70-71 (95, 10000,
                  (0.5,1,1), (0.5,30,30), (1,2,2)); e.g. ``sigma_mask=(3, 2, 2)`` projects a wider z band.
"""
from __future__ import annotations

import os
import pickle
import time

import numpy as np

from . import _native
from . import basic_image_manipulations as bim
from .basic_image_manipulations import put_channel_axis_first, read_image_in_chunks  # noqa: F401
from .movie import allocate_outputs, as_uint16_stack, finish_outputs, rank_world
from .tiff_io import save_npy

DEFAULT_MODE = os.environ.get("TSP_MODE", "fast")
last_job_timings = {}        # seconds of the last movie_surface_projection call on this rank, by phase (diagnostics)


def _as_uint16_stack(image):
    return np.ascontiguousarray(as_uint16_stack(image))


def time_point_surface_projection(time_point, axes, reference_channel, min_z=0, max_z=0,
                                  method='max_averages', bin_size=1, airyscan=True, z_map=False, atoh_shift=0,
                                  build_manifold=False, *, mode=None, device=None, percentile=None, pedestal=None,
                                  sigma_pre=None, sigma_score=None, sigma_mask=None):
    """SP:17-85.  See SURVEY.md section 3.3 for the behavioural spec this reproduces, quirks
    included: airyscan defaults to True; ``min_z`` is added to the height map even when
    ``max_z == 0``; the band is indexed with the un-cropped height (IndexError when it leaves the
    cropped stack, SP:68-69); non-channel-first inputs come back as (C, X, Y)."""
    time_point = np.asarray(time_point)
    if axes.find("T") >= 0:
        time_point = time_point.reshape(time_point.shape[1:])
        image, _ = put_channel_axis_first(time_point, axes[1:])
        axes_wo_t = axes[1:]
    else:
        image, _ = put_channel_axis_first(time_point, axes)
        axes_wo_t = axes
    if image.ndim != 4:
        # SP:37 blurs image[reference_channel] with three sigmas; anything but a 3-D channel fails in
        # scipy exactly like this
        raise RuntimeError("sequence argument must have length equal to input rank")
    if bin_size > 1 and method not in _native.METHODS:
        raise TypeError("exceptions must derive from BaseException")      # SP:53 raises a str
    if percentile is not None and not 0 <= percentile <= 100:
        raise ValueError("Percentiles must be in the range [0, 100]")     # what np.percentile raises at SP:35
    stack = _as_uint16_stack(image)
    params = dict(percentile=percentile, pedestal=pedestal, sigma_pre=sigma_pre, sigma_score=sigma_score,
                  sigma_mask=sigma_mask)
    proj, zmap, _ = _native.project_frame_host(stack, int(reference_channel), int(min_z), int(max_z),
                                               bool(airyscan), int(atoh_shift), mode or DEFAULT_MODE, device,
                                               bin_size=int(bin_size) if bin_size > 1 else 1,
                                               method=method if bin_size > 1 else "max_averages",
                                               build_manifold=bool(build_manifold), params=params)
    if axes_wo_t.find("C") < 0:                         # unreachable in the reference (see above)
        proj = proj[0]
    if z_map:
        return proj, zmap
    return proj


def build_continues_manifold(score):
    """SP:87-128 on the GPU: score (Z, Y, X) float32 -> int64 height map grown outwards from the global score
    maximum (``tsp_build_manifold``; at most 254 planes).  The per-pixel rule of SP:130-165 (find_pixel_plane) only
    exists fused inside that kernel."""
    return _native.build_manifold(np.ascontiguousarray(score, dtype=np.float32))


# ------------------------------------------------------------------------------------------------
# on-disk conventions
# ------------------------------------------------------------------------------------------------
def _load_uint16(path, fresh=None):
    """A per-movie array as uint16 (BIM:481 / SP:230: astype truncates): from ``fresh`` when this run just computed
    and saved it, else from the resume file."""
    arr = fresh.get(path) if fresh else None
    if arr is None:
        arr = np.load(path)
    return arr if arr.dtype == np.uint16 else arr.astype("uint16")


def concatenate_time_points(files, fresh=None):
    """What BIM:478-495 does to the per-movie arrays of one position (without its resize branch): every array is
    cast to uint16 (truncation), later movies that lost channels get zero channels in FRONT, then all are joined
    along time."""
    movies = [_load_uint16(f, fresh) for f in files]
    lead = movies[0].shape
    for k in range(1, len(movies)):
        m = movies[k]
        grow = [(max(lead[d] - m.shape[d], 0), 0) if 1 <= d < m.ndim - 2 else (0, 0) for d in range(m.ndim)]
        if any(g[0] for g in grow):
            movies[k] = np.pad(m, grow, constant_values=0)
    return movies[0] if len(movies) == 1 else np.concatenate(movies, axis=0)


def save_tiff(path, image, metadata=None, axes="", data_type=""):
    """BIM:162-189: dtype convention (uint8 / uint16 targets rescale to the global maximum) + a pluggable writer.
    The default writer is the package's own OME-flavoured TIFF / BigTIFF writer (``tiff_io.write_tiff``: the
    reference's ``OmeTiffWriter`` needs aicsimageio)."""
    if data_type and image.dtype != data_type and data_type in ("uint8", "uint16"):
        image = _rescale_to(image, 255 if data_type == "uint8" else 65535, data_type)
    tiff_writer(path, image, axes, metadata)


def _rescale_to(image, top, data_type, chunk=1 << 20):
    """``np.round((image / np.max(image)) * top).astype(data_type)`` (BIM:183-186), the same float64 arithmetic element
    by element, but in cache-sized pieces on a few threads instead of four passes over full-size temporaries (a
    2 x 4096 x 4096 projection: 0.54 -> 0.13 s)."""
    image = np.asarray(image)
    peak = np.max(image) if image.size else 0
    if image.size < 4 * chunk or image.dtype != np.float64 or not np.isfinite(peak) or peak == 0:
        return np.round((image / peak) * top).astype(data_type)
    from concurrent.futures import ThreadPoolExecutor
    flat = np.ascontiguousarray(image).reshape(-1)
    out = np.empty(flat.shape, dtype=data_type)

    def piece(a):
        tmp = flat[a:a + chunk] / peak
        tmp *= top
        np.rint(tmp, out=tmp)
        out[a:a + chunk] = tmp

    with ThreadPoolExecutor(max_workers=4) as pool:
        list(pool.map(piece, range(0, flat.size, chunk)))
    return out.reshape(image.shape)


def _default_tiff_writer(path, image, axes, metadata):
    from . import tiff_io
    tiff_io.write_tiff(path, image, axes=axes, metadata=metadata)


tiff_writer = _default_tiff_writer        # replaceable hook: callable(path, image, axes, metadata)


def update_projection_metadata(metadata, frames_number, series=0):
    """SP:319-327: the OME metadata of one projected position - a single image named after the position, one
    plane per channel, ``frames_number`` time points, uint16, XYCTZ."""
    image = metadata.images[series]
    pixels = image.pixels
    metadata.images = [image]
    image.name = 'position%d' % series
    for field, value in (("dimension_order", 'XYCTZ'), ("size_z", 1), ("size_t", frames_number), ("type", 'uint16')):
        setattr(pixels, field, value)
    pixels.planes = pixels.planes[:pixels.size_c]
    return metadata


def movie_schedule(n_files, position_final_movie, n_positions):
    """The bookkeeping of SP:181-221 / SP:240-262 as data: yields (file index, series index inside that file,
    position).  A position whose final movie is ``f`` is absent from the files after ``f``, and the series of a file
    number the positions still alive, in order."""
    alive = list(range(n_positions))
    for f in range(n_files):
        for series, position in enumerate(alive):
            yield f, series, position
        alive = [p for p in alive if position_final_movie[p] != f + 1]


_STAGE_SCALARS = ("x_unit", "y_unit", "z_unit")
_PIXEL_SCALARS = ("physical_size_x", "physical_size_y", "physical_size_z")


def save_stage_positions(files, position_final_movie, initial_positions_number, output_dir, only_position=0,
                         output_name=""):
    """SP:240-276: ``stage_locations_position%d.pkl`` per position - the stage x / y / z of every time point (one
    entry per frame, movie after movie), their units and the physical pixel sizes of the first movie."""
    wanted = [p for p in range(initial_positions_number) if only_position <= 0 or p == only_position - 1]
    tracks = {}
    first = bim.get_image_metadata(files[0])
    for p in range(initial_positions_number):              # units and pixel sizes: first movie, series = position
        im = first.images[p]
        rec = {axis: [] for axis in "xyz"}
        rec.update({k: getattr(im.stage_label, k) for k in _STAGE_SCALARS})
        rec.update({k: getattr(im.pixels, k) for k in _PIXEL_SCALARS})
        tracks[p] = rec
    per_file = {0: first}
    for f, series, p in movie_schedule(len(files), position_final_movie, initial_positions_number):
        if f > 0 and p not in wanted:
            continue                                       # SP:259-260 skips other positions only after the first movie
        if f not in per_file:
            per_file[f] = bim.get_image_metadata(files[f])
        im = per_file[f].images[series]
        for axis in "xyz":
            tracks[p][axis].extend([getattr(im.stage_label, axis)] * im.pixels.size_t)
    for p in wanted:
        with open(os.path.join(output_dir, output_name + "stage_locations_position%d.pkl" % (p + 1)), 'wb') as f:
            pickle.dump(tracks[p], f)


# ------------------------------------------------------------------------------------------------
# drivers
# ------------------------------------------------------------------------------------------------
def _default_pipeline(mode, out_dtype):
    from .movie import FramePipeline
    return FramePipeline(mode=mode, out_dtype=out_dtype)


def _job_barrier():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from .movie import host_group
            dist.barrier(group=host_group())
    except ImportError:                                   # pragma: no cover
        pass


def movie_surface_projection(files, reference_channel, position_final_movie, initial_positions_number, output_dir,
                             method, bin_size, build_manifold, only_position, zmin, zmax, airyscan,
                             output_name="", *, mode=None, frame_pipeline=None):
    """SP:168-237.  Every (movie file, position) pair becomes one job: its time points stream through a
    ``movie.FramePipeline`` (pinned staging, frame slots, uint16 conversion on the device; under torchrun the ranks
    share the time points and rank 0 assembles them) into the resume files ``position%d_movie%d_{projection,zmap}.npy``
    of SP:193-194.  When all jobs are done, rank 0 joins the movies of each position along time and writes
    ``position%d.tif`` (uint16, TCYX), ``zmap_position%d.npy`` (uint16, (T,1,1,Y,X)) and the stage pickle, then deletes
    the resume files.  A job whose two resume files exist is skipped (SP:199-200)."""
    rank, world = rank_world()
    root = rank == 0
    pipeline = frame_pipeline if frame_pipeline is not None else _default_pipeline(mode, "uint16")
    out_dtype = np.uint16 if pipeline.out_dtype == "uint16" else np.float64
    chosen = [p for p in range(initial_positions_number) if only_position <= 0 or p == only_position - 1]
    frames_of = {p: 0 for p in chosen}
    resume = {p: ([], []) for p in range(initial_positions_number)}
    fresh = {}                                         # arrays computed in this run: no need to read them back
    outputs = []                                       # their backing (shared by the ranks of a single-host job)
    dims_of = {}
    timings = {"project_s": 0.0, "resume_save_s": 0.0, "assemble_write_s": 0.0}
    jobs = [job for job in movie_schedule(len(files), position_final_movie, initial_positions_number) if job[2] in chosen]
    for job_index, (f, series, position) in enumerate(jobs):
        if f not in dims_of:
            dims_of[f] = bim.get_image_dimensions(files[f])
        dims = dims_of[f]
        stem = os.path.join(output_dir, "position%d_movie%d" % (position, f))
        proj_path, zmap_path = stem + "_projection.npy", stem + "_zmap.npy"
        resume[position][0].append(proj_path)
        resume[position][1].append(zmap_path)
        frames_of[position] += dims.T
        if root:
            print("Projecting position %d, movie %d" % (position + 1, f + 1))
        done = os.path.isfile(proj_path) and os.path.isfile(zmap_path)
        if world > 1:                                  # every rank must take the same branch: rank 0's view decides
            done = _broadcast_flag(done)
        if done:
            continue
        reference_channel = min(reference_channel, dims.C - 1)       # SP:203-204 (sticks for the later jobs too)
        outputs.append(allocate_outputs([((dims.T, dims.C, 1, dims.Y, dims.X), out_dtype),
                                         ((dims.T, 1, 1, dims.Y, dims.X), out_dtype)]))
        proj, zmap = outputs[-1].arrays
        t0 = time.perf_counter()
        pipeline.project_movie(files[f], series, proj, zmap, mode=mode, gather="root",
                               reference_channel=reference_channel, method=method, bin_size=bin_size, atoh_shift=0,
                               build_manifold=build_manifold, min_z=zmin, max_z=zmax, airyscan=airyscan)
        timings["project_s"] += time.perf_counter() - t0
        if root:
            t0 = time.perf_counter()
            fresh[proj_path] = proj.reshape((dims.T, dims.C, dims.Y, dims.X))
            fresh[zmap_path] = zmap
            # the resume files of SP:193-194 exist to restart an interrupted run at the next job; those of the very
            # last job would be deleted a moment later by the clean-up below (SP:235-237) - they are not written
            if job_index + 1 < len(jobs):
                save_npy(proj_path, fresh[proj_path])
                save_npy(zmap_path, zmap)
            timings["resume_save_s"] += time.perf_counter() - t0
    t0 = time.perf_counter()
    if root:
        from concurrent.futures import ThreadPoolExecutor
        for position in chosen:
            proj_files, zmap_files = resume[position]
            metadata = update_projection_metadata(bim.get_image_metadata(files[0], series=position),
                                                  float(frames_of[position]), series=position)
            zmaps = [_load_uint16(z, fresh) for z in zmap_files]
            # the position's two files side by side: writes to ONE file are serialised by the file system, two files
            # take the time of one (840 MB of a 200-frame 1024 x 1024 movie: 0.15 -> 0.08 s on the B200 box)
            tif_path = os.path.join(output_dir, output_name + "position%d.tif" % (position + 1))
            npy_path = os.path.join(output_dir, output_name + "zmap_position%d.npy" % (position + 1))
            with ThreadPoolExecutor(max_workers=2) as writers:
                pending = [
                    writers.submit(save_tiff, tif_path, concatenate_time_points(proj_files, fresh), metadata=metadata,
                                   axes="TCYX", data_type="uint16"),
                    writers.submit(save_npy, npy_path, zmaps[0] if len(zmaps) == 1 else np.concatenate(zmaps, axis=0))]
            try:
                for job in pending:
                    job.result()
            except BaseException:                     # a failed write: leave only the resume files behind
                for path in (tif_path, npy_path):
                    if os.path.exists(path):
                        os.remove(path)
                raise
        save_stage_positions(files, position_final_movie, initial_positions_number, output_dir,
                             only_position=only_position, output_name=output_name)
        for proj_files, zmap_files in resume.values():
            for path in proj_files + zmap_files:
                if os.path.exists(path):
                    os.remove(path)
    timings["assemble_write_s"] = time.perf_counter() - t0
    _job_barrier()                                     # nobody returns before the outputs are on disk
    last_job_timings.clear()
    last_job_timings.update(timings)
    fresh.clear()
    for out in outputs:
        out.close()


def _broadcast_flag(flag):
    import torch
    import torch.distributed as dist
    from .movie import host_group
    t = torch.tensor([1 if flag else 0], dtype=torch.int64)
    dist.broadcast(t, src=0, group=host_group())
    return bool(t.item())


def _tiles(extent_y, extent_x, chunk):
    step_y, step_x = (chunk or extent_y), (chunk or extent_x)
    for y in range(0, extent_y, step_y):
        for x in range(0, extent_x, step_x):
            yield y, min(y + step_y, extent_y), x, min(x + step_x, extent_x)


def large_image_projection(input_dir, output_dir, input_file_name, position=1, reference_channel=0, chunk_size=0,
                           bin_size=1, channels_shift=0, min_z=0, max_z=0, method="", build_manifold=False,
                           airyscan=False, *, mode=None, frame_pipeline=None):
    """SP:279-316: fixed-sample projection.  The image is cut into independent XY tiles of ``chunk_size`` (no halo,
    every tile with its own percentile and edges - BIM:104-149 semantics) which stream through the frame pipeline like
    the time points of a movie; the float64 results are assembled into ``<name>[_position%d]_projection.tif`` (rescaled
    to the global maximum, uint16, BIM:183-186) and ``<name>[_position%d]_zmap.npy`` (float64 (T,Y,X))."""
    many = hasattr(position, "__len__")
    path = os.path.join(input_dir, input_file_name)
    if not os.path.exists(path):
        return 0
    rank, _ = rank_world()
    pipeline = frame_pipeline if frame_pipeline is not None else _default_pipeline(mode, "reference")
    if pipeline.out_dtype != "reference":
        raise ValueError("large_image_projection rescales float64 results: the pipeline must return the reference dtypes")
    dims = bim.get_image_dimensions(path)
    postfix = '.' + input_file_name.split('.')[-1]
    for pos in (position if many else [position]):
        outputs = allocate_outputs([((dims.T, dims.C, 1, dims.Y, dims.X), np.float64),
                                    ((dims.T, 1, 1, dims.Y, dims.X), np.float64)])
        projection, zmap = outputs.arrays
        img = bim.open_image(path)
        img.set_scene(int(pos - 1))
        data = img.get_image_dask_data()
        tiles = [(t,) + tile for t in range(dims.T) for tile in _tiles(dims.Y, dims.X, chunk_size)]
        from .movie import SharedFrameCounter
        counter = SharedFrameCounter("tiles")
        mine = []

        def source():
            for k in counter.claims(len(tiles)):
                t, y0, y1, x0, x1 = tiles[k]
                chunk = data[t:t + 1, :, :, y0:y1, x0:x1]
                frame_of = getattr(pipeline, "frame_of", None)           # FramePipeline: TIFF tiles stay lazy
                yield k, frame_of(chunk) if frame_of else np.asarray(chunk.compute())[0]

        def sink(k, proj, zm, status):
            t, y0, y1, x0, x1 = tiles[k]
            projection[t, :, 0, y0:y1, x0:x1] = proj
            zmap[t, 0, 0, y0:y1, x0:x1] = zm
            mine.append(k)
            print("Projecting position %d chunk %d" % (pos, k + 1), flush=True)

        pipeline.project_frames(source(), sink, mode=mode, reference_channel=reference_channel, min_z=min_z,
                                max_z=max_z, method=method, bin_size=bin_size, atoh_shift=channels_shift,
                                build_manifold=build_manifold, airyscan=airyscan)
        if outputs.shared:
            finish_outputs([projection, zmap], mine)   # every rank wrote its tiles into the job's arrays: a barrier
        elif _multi_rank():
            # tiles are scattered inside frames: ship whole arrays of the tiles' frames as a sum of disjoint parts
            _merge_disjoint([projection, zmap])
        if rank == 0:
            tag = "_position%d" % pos if many else ""
            out_proj = projection.reshape((dims.T, dims.C, dims.Y, dims.X) if dims.T > 1 else (dims.C, dims.Y, dims.X))
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=2) as writers:        # the two files side by side (see the movie driver)
                pending = [
                    writers.submit(save_tiff, os.path.join(output_dir, input_file_name.replace(postfix, tag + "_projection.tif")),
                                   out_proj, axes="TCYX" if dims.T > 1 else "CYX", data_type="uint16"),
                    writers.submit(save_npy, os.path.join(output_dir, input_file_name.replace(postfix, tag + "_zmap.npy")),
                                   zmap.reshape((dims.T, dims.Y, dims.X)))]
            for job in pending:
                job.result()
        outputs.close()                                # rank 0 has saved the arrays: their backing can go
    _job_barrier()


def _multi_rank():
    return rank_world()[1] > 1


def _merge_disjoint(arrays):
    """Tiles of one frame may have been projected by different ranks: every pixel was written by exactly one rank
    (zero elsewhere), so a sum over the ranks onto rank 0 assembles the frame (gloo group, output assembly only)."""
    import torch
    import torch.distributed as dist
    from .movie import host_group
    for a in arrays:
        dist.reduce(torch.from_numpy(a), dst=0, op=dist.ReduceOp.SUM, group=host_group())


# ------------------------------------------------------------------------------------------------
# command line (SP:329-423): same flags, same dispatch
# ------------------------------------------------------------------------------------------------
CLI_FLAGS = (
    # flags, dest, type, default, help                                              (SP:334-378)
    (("-i", "--input"), "input", str, "", "input directory with movies m1, m2, m3, ... [default: current directory]"),
    (("-o", "--output"), "output", str, "", "output directory [default: the input directory]"),
    (("-f", "--position-final_movie"), "position_final_movie", str, "",
     "final movie of each sample, in the order of the initial positions [default: all end in the last movie]"),
    (("-n", "--position-number"), "position_number", int, 1, "number of initial positions"),
    (("-m", "--movie-number"), "movie_number", int, 1, "number of movies"),
    (("-r", "--reference_channel"), "reference_channel", int, 1, "channel the height map is taken from"),
    (("-c", "--chunk-size"), "chunk_size", int, 0, "tile size of a large fixed-sample projection [default: whole image]"),
    (("--method",), "method", str, "max_averages", "score method when --bin-size > 1"),
    (("--file",), "file_name", str, None, "file name (fixed-sample projection)"),
    (("-b", "--bin-size"), "bin_size", int, 1, "bin size of the block mean / variance"),
    (("--only-position",), "only_position", int, 0, "project only this position [default: all]"),
    (("--min-z",), "zmin", int, 0, "first plane considered"),
    (("--max-z",), "zmax", int, 0, "last plane considered [default: all]"),
)
CLI_SWITCHES = (
    (("--fixed",), "fixed_sample", "fixed-sample (tiled) projection instead of a movie projection"),
    (("--manifold",), "build_manifold", "grow a continuous manifold instead of the pixel-wise argmax"),
    (("--airyscan",), "airyscan", "airyscan-processed intensities (pedestal 10000)"),
    (("--separate-files",), "separate_files", "project every czi file of the input directory on its own"),
)


def getOptions(argv=None):
    """SP:329-379.  Returns (options, positional_args) like optparse did."""
    import argparse
    parser = argparse.ArgumentParser(usage="%(prog)s [options]", allow_abbrev=False)
    for flags, dest, typ, default, text in CLI_FLAGS:
        parser.add_argument(*flags, dest=dest, type=typ, default=default, help=text)
    for flags, dest, text in CLI_SWITCHES:
        parser.add_argument(*flags, dest=dest, action="store_true", default=False, help=text)
    return parser.parse_known_args(argv)


def _movie_file(input_dir, number):
    """``m<number>.czi`` as in SP:413; when that file does not exist but ``m<number>.tif`` / ``.tiff`` does (a movie
    converted for the package's own TIFF reader), that one."""
    czi = os.path.join(input_dir, "m%d.czi" % number)
    if not os.path.exists(czi):
        for ext in (".tif", ".tiff"):
            other = os.path.join(input_dir, "m%d%s" % (number, ext))
            if os.path.exists(other):
                return other
    return czi


def main(argv=None):
    """SP:381-423: dispatch to the fixed-sample, per-file or movie driver exactly as the reference's __main__.
    Launched by ``torchrun`` (one process per GPU) the ranks join one job first (``movie.init_job``)."""
    from ast import literal_eval
    from glob import glob
    options, _ = getOptions(argv)
    from .movie import init_job
    init_job()                                # under torchrun: one rank per GPU, the ranks share the frames / tiles
    input_dir = options.input or os.getcwd()
    output_dir = options.output or input_dir
    common = dict(method=options.method, bin_size=options.bin_size, build_manifold=options.build_manifold)
    if options.fixed_sample:
        position = options.only_position if options.only_position > 0 else \
            np.arange(start=1, stop=options.position_number + 1)
        large_image_projection(input_dir, output_dir, options.file_name, position=position,
                               reference_channel=options.reference_channel, chunk_size=options.chunk_size,
                               min_z=options.zmin, max_z=options.zmax, airyscan=options.airyscan, **common)
    elif options.separate_files:
        for file in glob(os.path.join(input_dir, "*.czi")):
            movie_surface_projection([file], options.reference_channel, (1,), options.position_number, output_dir,
                                     options.method, options.bin_size, options.build_manifold, options.only_position,
                                     options.zmin, options.zmax, options.airyscan,
                                     output_name=os.path.basename(file))
    else:
        if options.position_final_movie:
            final = list(literal_eval(options.position_final_movie))
        else:
            final = [options.movie_number] * options.position_number
        files = [_movie_file(input_dir, i + 1) for i in range(options.movie_number)]
        movie_surface_projection(files, options.reference_channel, final, options.position_number, output_dir,
                                 options.method, options.bin_size, options.build_manifold, options.only_position,
                                 options.zmin, options.zmax, options.airyscan)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
