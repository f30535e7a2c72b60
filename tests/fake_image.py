"""In-memory stand-in for aicsimageio.AICSImage (the surface the projection drivers use): dims.{T,C,Z,Y,X},
set_scene, get_image_dask_data() -> sliceable with .compute(), metadata with stage labels."""
import types

import numpy as np


class _Lazy:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _Lazy(self.arr[idx])

    def compute(self):
        return np.array(self.arr)


class FakeAICSImage:
    def __init__(self, scenes):
        """scenes: list of (T,C,Z,Y,X) arrays, one per position."""
        self.scenes = scenes
        self.scene = 0

    def set_scene(self, i):
        self.scene = int(i)

    @property
    def dims(self):
        T, C, Z, Y, X = self.scenes[self.scene].shape
        return types.SimpleNamespace(T=T, C=C, Z=Z, Y=Y, X=X)

    def get_image_dask_data(self):
        return _Lazy(self.scenes[self.scene])

    @property
    def metadata(self):
        images = []
        for i, s in enumerate(self.scenes):
            T, C = s.shape[:2]
            images.append(types.SimpleNamespace(
                name="scene%d" % i,
                stage_label=types.SimpleNamespace(x=10.0 * i, y=20.0 * i, z=1.0, x_unit="um", y_unit="um", z_unit="um"),
                pixels=types.SimpleNamespace(size_t=T, size_c=C, size_z=s.shape[2], physical_size_x=0.1,
                                             physical_size_y=0.1, physical_size_z=0.5, dimension_order="XYZCT",
                                             type="uint16", planes=list(range(T * C * s.shape[2])))))
        return types.SimpleNamespace(images=images)


def install(monkeypatch_or_module, files):
    """Route basic_image_manipulations.open_image to {path: FakeAICSImage}."""
    from tissue_image_processing_b200 import basic_image_manipulations as bim

    def opener(path):
        return FakeAICSImage(files[path].scenes)
    if hasattr(monkeypatch_or_module, "setattr"):
        monkeypatch_or_module.setattr(bim, "open_image", opener)
    else:
        bim.open_image = opener
