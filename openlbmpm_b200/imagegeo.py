"""Image -> geometry pre-processing of the reference's host classes:
`~/StructureImage/structure.png` (0 = solid, anything else = void) is cropped to the bounding box of its solid
pixels, optionally mirror-tiled, closed with solid side columns and padded with void buffer rows at the outlet
(low row indices) and the inlet (high row indices).
RKD2Q9.py:373-414 (`__processImage`), ShanChenD2Q9.py:514-585 (`__expandImageDomain`, `__processImage`).
Host-side set-up code, never on the timed path."""
import os

import numpy as np


def structure_image_path():
    return os.environ.get("LBM_STRUCTURE_IMAGE", os.path.expanduser("~/StructureImage/structure.png"))


def load_gray(path=None):
    """grey-level image as float64 `[y, x]` (the reference: scipy.misc.imread(file, flatten=True))"""
    path = path or structure_image_path()
    if not os.path.exists(path):
        raise FileNotFoundError("The image file or the directory does not exist: %s" % path)
    if path.endswith(".npy"):
        return np.asarray(np.load(path), dtype=np.float64)
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im.convert("F"), dtype=np.float64)


def crop_to_solid(img):
    """bounding box of the solid (== 0) pixels (RKD2Q9.py:388-399)"""
    ys, xs = np.nonzero(img == 0.0)
    if ys.size == 0:
        raise ValueError("the structure image has no solid (black) pixel")
    return img[ys.min():ys.max() + 1, xs.min():xs.max() + 1].copy()


def expand_image_domain(arr, x_num, y_num):
    """periodic mirror tiling (ShanChenD2Q9.py:514-540): odd tiles are flipped so that the pore space is continuous"""
    rows = []
    for i in range(y_num):
        base = arr if i % 2 == 0 else np.flipud(arr)
        rows.append(np.hstack([base if j % 2 == 0 else np.fliplr(base) for j in range(x_num)]))
    return np.vstack(rows)


def close_and_pad(arr, rows_low, rows_high):
    """solid first / last column, then `rows_low` void rows in front (outlet side) and `rows_high` behind (inlet side)"""
    arr = arr.copy()
    arr[:, 0] = 0.0; arr[:, -1] = 0.0
    w = arr.shape[1]
    return np.vstack([np.full((rows_low, w), 255.0), arr, np.full((rows_high, w), 255.0)])


def to_domain(effective):
    """boolean void mask: everything that is not exactly 0 (RKD2Q9.py:436-439)"""
    return np.ascontiguousarray(effective != 0.0)
