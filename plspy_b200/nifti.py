"""Minimal NIfTI-1 reader / writer (numpy + gzip only), standing in for the `nibabel.load` calls of the reference's
ingest (plspy/io/io.py:43, 69: `nibabel.load(path)`; :207-231: `img.dataobj`).  nibabel is not a dependency of this
package: the ingest path needs exactly three things from an image -- its shape, its voxel array with the header's
intensity scaling applied, and the affine for writing result maps back -- and the single-file NIfTI-1 format
(348-byte header, data at `vox_offset`, x fastest) gives them directly.

Supported: `.nii`, `.nii.gz`, and the two-file `.hdr` / `.img` (+ `.gz`) pair (magic `ni1`; plain Analyze 7.5 headers
without a magic string are read the same way); either byte order; the integer / float datatypes NIfTI defines for
real-valued images.  Not supported (raises `NiftiError`): NIfTI-2, complex and RGB datatypes, header extensions are
skipped.  `dataobj` is a float view only when the header asks for scaling (`scl_slope` not 0 / 1 or `scl_inter` not
0), otherwise the stored dtype is kept -- which is what lets `io.assemble_pinned(dtype=np.float32)` store float32 /
int16 sources exactly in half the bytes.
"""
import gzip
import os
import struct

import numpy as np


class NiftiError(ValueError):
    pass


# NIfTI-1 datatype code -> numpy dtype (nifti1.h DT_*)
_DTYPES = {2: "u1", 4: "i2", 8: "i4", 16: "f4", 64: "f8", 256: "i1", 512: "u2", 768: "u4", 1024: "i8", 1280: "u8"}
_CODES = {np.dtype(v).newbyteorder("=").str[1:]: k for k, v in _DTYPES.items()}


class NiftiImage:
    """What the ingest uses of `nibabel.nifti1.Nifti1Image`: `.dataobj` (array in file axis order x, y, z[, t]),
    `.shape`, `.affine` (4 x 4, from the sform, else from pixdim), `.header` (dict of the fields read)."""

    def __init__(self, dataobj, affine=None, header=None):
        self.dataobj = dataobj
        self.affine = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
        self.header = header or {}

    @property
    def shape(self):
        return self.dataobj.shape

    def get_fdata(self):
        return np.asarray(self.dataobj, dtype=np.float64)


def _open(path):
    return gzip.open(path, "rb") if path.endswith(".gz") else open(path, "rb")


def _strip_gz(path):
    return path[:-3] if path.endswith(".gz") else path


def load(path):
    """Read one image file.  Equivalent of `nibabel.load(path)` for the cases listed in the module docstring."""
    base = _strip_gz(path)
    if base.endswith(".img"):                       # the pair is addressed through its header
        hdr_path = base[:-4] + ".hdr"
        path = hdr_path if os.path.exists(hdr_path) else hdr_path + ".gz"
        base = _strip_gz(path)
    with _open(path) as f:
        raw = f.read(352)
        if len(raw) < 348:
            raise NiftiError(f"{path}: shorter than a NIfTI-1 header")
        bo = "<"
        if struct.unpack("<i", raw[:4])[0] != 348:
            bo = ">"
            if struct.unpack(">i", raw[:4])[0] != 348:
                if struct.unpack("<i", raw[:4])[0] == 540 or struct.unpack(">i", raw[:4])[0] == 540:
                    raise NiftiError(f"{path}: NIfTI-2 is not supported")
                raise NiftiError(f"{path}: not a NIfTI-1 file (sizeof_hdr != 348)")
        dim = struct.unpack(bo + "8h", raw[40:56])
        if not 1 <= dim[0] <= 7:
            raise NiftiError(f"{path}: bad dim[0] = {dim[0]}")
        datatype, bitpix = struct.unpack(bo + "2h", raw[70:74])
        pixdim = struct.unpack(bo + "8f", raw[76:108])
        vox_offset, slope, inter = struct.unpack(bo + "3f", raw[108:120])
        qform_code, sform_code = struct.unpack(bo + "2h", raw[252:256])
        srow = np.array(struct.unpack(bo + "12f", raw[280:328]), dtype=np.float64).reshape(3, 4)
        magic = raw[344:348]
        if datatype not in _DTYPES:
            raise NiftiError(f"{path}: datatype code {datatype} is not supported")
        dt = np.dtype(bo + _DTYPES[datatype])
        if dt.itemsize * 8 != bitpix:
            raise NiftiError(f"{path}: bitpix {bitpix} does not match datatype {datatype}")
        shape = tuple(int(d) for d in dim[1:1 + dim[0]])
        count = int(np.prod(shape, dtype=np.int64))
        single = magic[:3] == b"n+1"
        if single:
            off = int(vox_offset) if vox_offset >= 352 else 352
            f.seek(off)
            buf = f.read(count * dt.itemsize)
    if not single:
        img_path = base[:-4] + ".img"
        if not os.path.exists(img_path):
            img_path += ".gz"
        with _open(img_path) as g:
            g.seek(int(vox_offset) if vox_offset > 0 else 0)
            buf = g.read(count * dt.itemsize)
    if len(buf) != count * dt.itemsize:
        raise NiftiError(f"{path}: expected {count * dt.itemsize} data bytes, found {len(buf)}")
    data = np.frombuffer(buf, dtype=dt).reshape(shape, order="F")           # x varies fastest in the file
    if not dt.isnative:
        data = data.astype(dt.newbyteorder("="))
    slope = slope if np.isfinite(slope) and slope != 0.0 else 1.0          # nifti1.h: slope 0 means "no scaling"
    inter = inter if np.isfinite(inter) else 0.0
    if slope != 1.0 or inter != 0.0:
        data = data.astype(np.float64) * slope + inter
    if sform_code > 0:
        affine = np.vstack([srow, [0.0, 0.0, 0.0, 1.0]])
    else:
        affine = np.diag([pixdim[1] or 1.0, pixdim[2] or 1.0, pixdim[3] or 1.0, 1.0])
    header = {"dim": dim, "datatype": datatype, "bitpix": bitpix, "pixdim": pixdim, "vox_offset": vox_offset,
              "scl_slope": slope, "scl_inter": inter, "qform_code": qform_code, "sform_code": sform_code,
              "magic": magic, "byteorder": bo}
    return NiftiImage(data, affine, header)


def save(path, data, affine=None, pixdim=None):
    """Write `data` (axis order x, y, z[, t]) as a single-file NIfTI-1 image (`.nii` or `.nii.gz`), little-endian, no intensity scaling.  Used for result maps (bootstrap ratios put back into the brain volume,
    io.remap_vectorized_subject_to_4d) and by the tests."""
    a = np.asarray(data)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    key = a.dtype.newbyteorder("=").str[1:]
    if key not in _CODES:
        raise NiftiError(f"dtype {a.dtype} cannot be stored in a NIfTI-1 file")
    if not 1 <= a.ndim <= 7:
        raise NiftiError("NIfTI-1 stores 1 to 7 dimensions")
    a = a.astype(a.dtype.newbyteorder("<"), copy=False)       # the header below is written little-endian
    affine = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
    hdr = bytearray(352)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, a.ndim, *(list(a.shape) + [1] * (7 - a.ndim)))
    struct.pack_into("<2h", hdr, 70, _CODES[key], a.dtype.itemsize * 8)
    pd = [1.0] * 8 if pixdim is None else list(pixdim) + [1.0] * (8 - len(pixdim))
    struct.pack_into("<8f", hdr, 76, *pd)
    struct.pack_into("<3f", hdr, 108, 352.0, 1.0, 0.0)
    struct.pack_into("<2h", hdr, 252, 0, 1)
    struct.pack_into("<12f", hdr, 280, *affine[:3].reshape(-1))
    hdr[344:348] = b"n+1\0"
    with (gzip.open(path, "wb", compresslevel=1) if path.endswith(".gz") else open(path, "wb")) as f:
        f.write(bytes(hdr))
        f.write(a.tobytes(order="F"))
