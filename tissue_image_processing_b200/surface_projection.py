"""B200 mirror of reference tissue_analyzing_tool/surface_projection.py.

``time_point_surface_projection`` keeps the reference signature (SP:17-19) and return dtypes
(float64 projection, int64 height map) and runs the whole operator on the GPU through the C ABI
(``tsp_project_frame_host``).  The drivers ``movie_surface_projection`` (SP:168-237) and
``large_image_projection`` (SP:279-316) keep their signatures, resume files and dtype
conventions; file I/O goes through the hooks of ``basic_image_manipulations`` (the reference's
Bio-Formats stack is out of scope).

``bin_size > 1`` (SP:39-53, methods max_averages / max_std / multi_channel) and ``build_manifold`` (SP:87-165)
run on the GPU too; they use the direct-FIR score (``mode="fast"`` behaves like "exact" for them).

Extension (keyword-only, defaults preserve reference behaviour): ``mode`` selects the score
stage - "fast" (default, multirate sigma=30 stage), "exact" (direct FIR, fp32) or "bitexact"
(direct FIR, scipy's float64 summation order; bit-identical height map and projection).
The default can also be set with the environment variable TSP_MODE.
"""
from __future__ import annotations

import os
import pickle

import numpy as np

from . import _native
from . import basic_image_manipulations as bim
from .basic_image_manipulations import put_channel_axis_first, read_image_in_chunks

DEFAULT_MODE = os.environ.get("TSP_MODE", "fast")


def _as_uint16_stack(image):
    if image.dtype == np.uint16:
        return np.ascontiguousarray(image)
    if image.dtype == np.uint8:
        return np.ascontiguousarray(image.astype(np.uint16))
    raise TypeError("the B200 projection path takes uint8/uint16 stacks (got %s)" % image.dtype)


def time_point_surface_projection(time_point, axes, reference_channel, min_z=0, max_z=0,
                                  method='max_averages', bin_size=1, airyscan=True, z_map=False, atoh_shift=0,
                                  build_manifold=False, *, mode=None, device=None):
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
    stack = _as_uint16_stack(image)
    proj, zmap, _ = _native.project_frame_host(stack, int(reference_channel), int(min_z), int(max_z),
                                               bool(airyscan), int(atoh_shift), mode or DEFAULT_MODE, device,
                                               bin_size=int(bin_size) if bin_size > 1 else 1,
                                               method=method if bin_size > 1 else "max_averages",
                                               build_manifold=bool(build_manifold))
    if axes_wo_t.find("C") < 0:                         # unreachable in the reference (see above)
        proj = proj[0]
    if z_map:
        return proj, zmap
    return proj


def build_continues_manifold(score):
    """SP:87-128 on the GPU: score (Z, Y, X) float32 -> int64 height map grown outwards from the global score
    maximum (``tsp_build_manifold``; at most 254 planes)."""
    return _native.build_manifold(np.ascontiguousarray(score, dtype=np.float32))


def find_pixel_plane(score, chozen_z, pixel_row, pixel_col, max_row, max_col, max_plane):
    raise NotImplementedError("find_pixel_plane (SP:130-165) is only available fused inside build_continues_manifold "
                              "on the B200 path")


# ------------------------------------------------------------------------------------------------
# drivers
# ------------------------------------------------------------------------------------------------
def concatenate_time_points(files):
    """BIM:478-495 without the resize branch: load per-movie arrays, truncate to uint16, pad the
    channel axis at the front when a later movie has fewer channels, concatenate along T."""
    imgs = []
    for file in files:
        img = np.load(file).astype("uint16")
        if imgs:
            for dim in range(1, img.ndim - 2):
                missing = imgs[0].shape[dim] - img.shape[dim]
                if missing > 0:
                    pad = [(0, 0)] * img.ndim
                    pad[dim] = (missing, 0)
                    img = np.pad(img, pad_width=pad, constant_values=0)
        imgs.append(img)
    return np.concatenate(imgs, axis=0)


def save_tiff(path, image, metadata=None, axes="", data_type=""):
    """BIM:162-189 dtype convention + pluggable writer.  uint8/uint16 targets rescale to the
    global maximum; the writer hook defaults to tifffile when installed."""
    if data_type and image.dtype != data_type and data_type in ("uint8", "uint16"):
        top = 255 if data_type == "uint8" else 65535
        image = np.round((image / np.max(image)) * top).astype(data_type)
    tiff_writer(path, image, axes, metadata)


def _default_tiff_writer(path, image, axes, metadata):
    try:
        import tifffile
    except ImportError as exc:                           # pragma: no cover - depends on the box
        raise ImportError("no TIFF writer available: set surface_projection.tiff_writer") from exc
    tifffile.imwrite(path, image, metadata={"axes": axes})


tiff_writer = _default_tiff_writer        # replaceable hook: callable(path, image, axes, metadata)


def update_projection_metadata(metadata, frames_number, series=0):
    """SP:319-327."""
    metadata.images = [metadata.images[series]]
    metadata.images[0].name = 'position%d' % series
    metadata.images[0].pixels.dimension_order = 'XYCTZ'
    metadata.images[0].pixels.size_z = 1
    metadata.images[0].pixels.size_t = frames_number
    metadata.images[0].pixels.type = 'uint16'
    metadata.images[0].pixels.planes = metadata.images[0].pixels.planes[:metadata.images[0].pixels.size_c]
    return metadata


def save_stage_positions(files, position_final_movie, initial_positions_number, output_dir, only_position=0,
                         output_name=""):
    """SP:240-276: per-position stage coordinates, one entry per time point, pickled."""
    positions = list(range(initial_positions_number))
    meta = bim.get_image_metadata(files[0])

    def entry(im):
        n = im.pixels.size_t
        return {"x": [im.stage_label.x] * n, "y": [im.stage_label.y] * n, "z": [im.stage_label.z] * n,
                "x_unit": im.stage_label.x_unit, "y_unit": im.stage_label.y_unit,
                "z_unit": im.stage_label.z_unit, "physical_size_x": im.pixels.physical_size_x,
                "physical_size_y": im.pixels.physical_size_y, "physical_size_z": im.pixels.physical_size_z}

    stage_pos = [entry(meta.images[i]) for i in range(initial_positions_number)]
    for position in range(initial_positions_number):
        if position_final_movie[position] == 1:
            positions.remove(position)
    for file_index in range(1, len(files)):
        meta = bim.get_image_metadata(files[file_index])
        done = []
        for position_index, position in enumerate(positions):
            if position_final_movie[position] == file_index + 1:
                done.append(position)
            if only_position > 0 and position != only_position - 1:
                continue
            im = meta.images[position_index]
            for k in "xyz":
                stage_pos[position][k].extend([getattr(im.stage_label, k)] * im.pixels.size_t)
        for p in done:
            positions.remove(p)
    for i in range(initial_positions_number):
        if only_position > 0 and i != only_position - 1:
            continue
        with open(os.path.join(output_dir, output_name + "stage_locations_position%d.pkl" % (i + 1)), 'wb') as f:
            pickle.dump(stage_pos[i], f)


def movie_surface_projection(files, reference_channel, position_final_movie, initial_positions_number, output_dir,
                             method, bin_size, build_manifold, only_position, zmin, zmax, airyscan,
                             output_name="", *, mode=None, frame_pipeline=None):
    """SP:168-237.  Per (file, position): project every time point (frames are independent; with
    ``frame_pipeline`` - see ``movie.FramePipeline`` - they are spread over the GPUs of the box),
    store per-movie .npy resume files, then per position concatenate (uint16 truncation), save the
    OME-TIFF / ``zmap_position%d.npy`` / stage pickle and delete the resume files."""
    positions = list(range(initial_positions_number))
    time_points_number = np.zeros((initial_positions_number, len(files)))
    projection_files = [[] for _ in range(initial_positions_number)]
    zmap_files = [[] for _ in range(initial_positions_number)]
    for file_num, file in enumerate(files):
        remove_positions = []
        dims = bim.get_image_dimensions(file)
        for position_num, position in enumerate(positions):
            if position_final_movie[position] == file_num + 1:
                remove_positions.append(position)
            if only_position > 0 and position != only_position - 1:
                continue
            projection_path = os.path.join(output_dir, "position%d_movie%d_projection.npy" % (position, file_num))
            zmap_path = os.path.join(output_dir, "position%d_movie%d_zmap.npy" % (position, file_num))
            projection_files[position].append(projection_path)
            zmap_files[position].append(zmap_path)
            print("Projecting position %d, movie %d" % (position + 1, file_num + 1))
            time_points_number[position, file_num] = dims.T
            if os.path.isfile(projection_path) and os.path.isfile(zmap_path):
                continue                                                   # resume (SP:199-200)
            current_projection = np.zeros((dims.T, dims.C, 1, dims.Y, dims.X))
            current_zmap = np.zeros((dims.T, 1, 1, dims.Y, dims.X))
            if reference_channel >= dims.C:
                reference_channel = dims.C - 1
            params = dict(axes='TCZYX', reference_channel=reference_channel, z_map=True, method=method,
                          bin_size=bin_size, atoh_shift=0, build_manifold=build_manifold, min_z=zmin,
                          max_z=zmax, airyscan=airyscan)
            if frame_pipeline is not None:
                frame_pipeline.project_movie(file, position_num, current_projection, current_zmap,
                                             mode=mode, **params)
            else:
                projector = read_image_in_chunks(file, series=position_num, dt=1,
                                                 apply_function=_operator_with_mode(mode),
                                                 output=[current_projection, current_zmap], **params)
                for time_point_index, _ in enumerate(projector):
                    print("Projecting timepoint %d" % (time_point_index + 1))
            current_projection = current_projection.reshape((dims.T, dims.C, dims.Y, dims.X))
            np.save(projection_path, current_projection)
            np.save(zmap_path, current_zmap)
        for to_delete in remove_positions:
            positions.remove(to_delete)
    for position in range(initial_positions_number):
        if only_position > 0 and position != only_position - 1:
            continue
        former_metadata = bim.get_image_metadata(files[0], series=position)
        new_metadata = update_projection_metadata(former_metadata, np.sum(time_points_number[position, :]),
                                                  series=position)
        movie_projection = concatenate_time_points(projection_files[position])
        save_tiff(os.path.join(output_dir, output_name + "position%d.tif" % (position + 1)), movie_projection,
                  metadata=new_metadata, axes="TCYX", data_type="uint16")
        movie_zmap = np.concatenate([np.load(f).astype("uint16") for f in zmap_files[position]], axis=0)
        np.save(os.path.join(output_dir, output_name + "zmap_position%d.npy" % (position + 1)), movie_zmap)
    save_stage_positions(files, position_final_movie, initial_positions_number, output_dir,
                         only_position=only_position, output_name=output_name)
    for position_files in projection_files + zmap_files:
        for projection_file in position_files:
            os.remove(projection_file)


def _operator_with_mode(mode):
    if mode is None:
        return time_point_surface_projection

    def op(chunk, **kw):
        return time_point_surface_projection(chunk, mode=mode, **kw)
    return op


def large_image_projection(input_dir, output_dir, input_file_name, position=1, reference_channel=0, chunk_size=0,
                           bin_size=1, channels_shift=0, min_z=0, max_z=0, method="", build_manifold=False,
                           airyscan=False, *, mode=None):
    """SP:279-316: fixed-sample projection in independent (no halo) XY tiles of ``chunk_size``."""
    if not hasattr(position, "__len__"):
        position = [position]
        add_pos = False
    else:
        add_pos = True
    path = os.path.join(input_dir, input_file_name)
    if not os.path.exists(path):
        return 0
    dims = bim.get_image_dimensions(path)
    for pos in position:
        projection = np.zeros((dims.T, dims.C, 1, dims.Y, dims.X))
        zmap = np.zeros((dims.T, 1, 1, dims.Y, dims.X))
        projector = read_image_in_chunks(path, dx=chunk_size, dy=chunk_size, dt=1,
                                         apply_function=_operator_with_mode(mode),
                                         output=[projection, zmap], axes='TCZYX', min_z=min_z, max_z=max_z,
                                         reference_channel=reference_channel, series=int(pos - 1), z_map=True,
                                         method=method, bin_size=bin_size, atoh_shift=channels_shift,
                                         build_manifold=build_manifold, airyscan=airyscan)
        for chunk_num, _ in enumerate(projector):
            print("Projecting position %d chunk %d" % (pos, chunk_num + 1), flush=True)
        if dims.T > 1:
            projection = projection.reshape((dims.T, dims.C, dims.Y, dims.X))
        else:
            projection = projection.reshape((dims.C, dims.Y, dims.X))
        zmap = zmap.reshape((dims.T, dims.Y, dims.X))
        postfix = '.' + input_file_name.split('.')[-1]
        pos_addition = "_position%d" % pos if add_pos else ""
        projection_file_name = os.path.join(output_dir,
                                            input_file_name.replace(postfix, pos_addition + "_projection.tif"))
        zmap_filename = os.path.join(output_dir, input_file_name.replace(postfix, pos_addition + "_zmap.npy"))
        save_tiff(projection_file_name, projection, axes="TCYX" if dims.T > 1 else "CYX", data_type="uint16")
        np.save(zmap_filename, zmap)


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


def main(argv=None):
    """SP:381-423: dispatch to the fixed-sample, per-file or movie driver exactly as the reference's __main__."""
    from ast import literal_eval
    from glob import glob
    options, _ = getOptions(argv)
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
        files = [os.path.join(input_dir, "m%d.czi" % (i + 1)) for i in range(options.movie_number)]
        movie_surface_projection(files, options.reference_channel, final, options.position_number, output_dir,
                                 options.method, options.bin_size, options.build_manifold, options.only_position,
                                 options.zmin, options.zmax, options.airyscan)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
