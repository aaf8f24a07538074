"""Ingest row: NIfTI-1 reader / writer, the mirrored `plspy.io` steps against the unmodified
reference's own functions (baseline/_ref; its nibabel import is stubbed, the functions compared here are pure numpy),
and the one-call `ingest_pinned`."""
import gzip
import os
import struct

import numpy as np
import pytest

import baseline
from plspy_b200 import exceptions, io as gio, nifti


def _vol(rs, shape=(5, 4, 3, 6), dtype=np.float32):
    return (rs.standard_normal(shape) * 50 + 100).astype(dtype)


def test_nifti_roundtrip_dtypes_and_gzip(tmp_path):
    rs = np.random.RandomState(0)
    for dt in (np.float32, np.float64, np.int16, np.uint8, np.int32):
        a = _vol(rs, dtype=dt)
        for ext in (".nii", ".nii.gz"):
            path = str(tmp_path / f"a_{np.dtype(dt).name}{ext}")
            aff = np.diag([2.0, 2.0, 3.0, 1.0]); aff[:3, 3] = [-10, 5, 7]
            nifti.save(path, a, affine=aff)
            img = nifti.load(path)
            assert img.shape == a.shape and img.dataobj.dtype == np.dtype(dt)
            assert np.array_equal(np.asarray(img.dataobj), a)
            assert np.allclose(img.affine, aff)


def test_nifti_header_layout_is_the_standard_one(tmp_path):
    """the writer is only a test double if its bytes are the published layout: check the fixed offsets by hand"""
    a = np.arange(24, dtype=np.int16).reshape(2, 3, 4)
    path = str(tmp_path / "h.nii")
    nifti.save(path, a)
    raw = open(path, "rb").read()
    assert struct.unpack("<i", raw[:4])[0] == 348
    assert struct.unpack("<8h", raw[40:56])[:4] == (3, 2, 3, 4)
    assert struct.unpack("<2h", raw[70:74]) == (4, 16)                 # DT_INT16, 16 bits
    assert struct.unpack("<f", raw[108:112])[0] == 352.0 and raw[344:348] == b"n+1\0"
    assert np.array_equal(np.frombuffer(raw[352:], dtype="<i2"), a.reshape(-1, order="F"))   # x fastest


def test_nifti_big_endian_scaling_and_pair(tmp_path):
    a = np.arange(2 * 3 * 4, dtype=np.int16).reshape(2, 3, 4)
    hdr = bytearray(348)
    struct.pack_into(">i", hdr, 0, 348)
    struct.pack_into(">8h", hdr, 40, 3, 2, 3, 4, 1, 1, 1, 1)
    struct.pack_into(">2h", hdr, 70, 4, 16)
    struct.pack_into(">8f", hdr, 76, 1, 2, 2, 2, 1, 1, 1, 1)
    struct.pack_into(">3f", hdr, 108, 0.0, 0.5, 10.0)                  # vox_offset, scl_slope, scl_inter
    hdr[344:348] = b"ni1\0"
    open(tmp_path / "p.hdr", "wb").write(bytes(hdr))
    with gzip.open(tmp_path / "p.img.gz", "wb") as f:
        f.write(a.astype(">i2").tobytes(order="F"))
    for name in ("p.hdr", "p.img"):
        img = nifti.load(str(tmp_path / name))
        assert img.dataobj.dtype == np.float64
        assert np.array_equal(np.asarray(img.dataobj), a * 0.5 + 10.0)
        assert np.allclose(np.diag(img.affine), [2, 2, 2, 1])
    with pytest.raises(nifti.NiftiError):
        bad = bytearray(hdr); struct.pack_into(">i", bad, 0, 540)
        open(tmp_path / "bad.nii", "wb").write(bytes(bad) + b"\0" * 8)
        nifti.load(str(tmp_path / "bad.nii"))


def test_open_images_in_dir_sorted_and_concat(tmp_path):
    rs = np.random.RandomState(1)
    vols = {f"v{i:02d}.nii": _vol(rs, (4, 3, 2)) for i in (3, 1, 2)}
    for n, v in vols.items():
        nifti.save(str(tmp_path / n), v)
    open(tmp_path / "v00.hdr", "wb").write(b"skip me")
    imgs, names = gio.open_images_in_dir(str(tmp_path))
    assert names == sorted(vols)
    one = gio.read_dir_to_one_image(str(tmp_path))
    assert one.shape == (4, 3, 2, 3)
    assert np.array_equal(one.dataobj[..., 0], vols["v01.nii"])
    mats, shape = gio.extract_matrices_image_list_realign([one])
    assert shape == (3, 4, 3, 2) and np.array_equal(mats[0][2], vols["v03.nii"])


@pytest.mark.skipif(not baseline.reference_available(), reason="baseline/_ref not installed")
def test_ingest_steps_match_the_reference_functions():
    ref = baseline.import_reference().io
    rs = np.random.RandomState(2)
    mats = [np.abs(_vol(rs, (6, 5, 4, 3), np.float64)) * (rs.rand(5, 4, 3) > 0.3) for _ in range(4)]
    for thr in (0.0, 0.15, 0.6):
        m_ref = ref.create_threshold_mask_from_matrices(mats, threshold=thr)
        m = gio.create_threshold_mask_from_matrices(mats, threshold=thr)
        assert np.array_equal(np.asarray(m_ref), m)
    with pytest.raises(exceptions.OutOfRangeError):
        gio.create_threshold_mask_from_matrices(mats, threshold=1.5)
    mask = gio.create_threshold_mask_from_matrices(mats)
    a_ref, a = ref.apply_mask_matrices(mats, mask), gio.apply_mask_matrices(mats, mask)
    assert all(np.array_equal(x, y) for x, y in zip(a_ref, a))
    assert np.array_equal(ref.create_and_apply_mask_list(mats), gio.create_and_apply_mask_list(mats))
    assert np.array_equal(ref.create_binary_mask_from_matrices(mats), gio.create_binary_mask_from_matrices(mats))
    onsets = np.array([[0, 3], [1, 4]])                                   # conditions x onsets
    s_ref = ref.extract_onset_slices_single_subject(mats[0], onsets, 1, 2.0)
    s = gio.extract_onset_slices_single_subject(mats[0], onsets, 1, 2.0)
    assert all(np.array_equal(x, y) for x, y in zip(s_ref, s))
    l_ref = ref.extract_onset_slices_list(mats, [onsets] * 4, 1, 2.0)
    l = gio.extract_onset_slices_list(mats, [onsets] * 4, 1, 2.0)
    assert all(np.array_equal(x, y) for sr, sg in zip(l_ref, l) for x, y in zip(sr, sg))
    assert np.array_equal(ref.concat_assemble_group(l_ref), gio.concat_assemble_group(l))
    vec = a[0]
    assert np.array_equal(ref.remap_vectorized_subject_to_4d(vec, mask, mats[0].shape),
                          gio.remap_vectorized_subject_to_4d(vec, mask, mats[0].shape))


def test_open_onsets_txt(tmp_path):
    np.savetxt(tmp_path / "s2.txt", np.array([[0.0, 4.1], [6.0, 9.9]]))
    np.savetxt(tmp_path / "s1.txt", np.array([[2.0, 4.0], [8.0, 12.0]]))
    on = gio.open_onsets_txt(str(tmp_path), 2.0)
    assert np.array_equal(on[0], [[1, 4], [2, 6]]) and np.array_equal(on[1], [[0, 3], [2, 5]])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_ingest_pinned_equals_the_step_by_step_sequence(tmp_path, dtype):
    """files -> X in one call == the reference's script sequence on the same files"""
    rs = np.random.RandomState(3)
    C, shape = 3, (5, 4, 3)
    paths, subj = [], []
    for g, n in enumerate((3, 2)):
        paths.append([])
        for s in range(n):
            v = np.abs(_vol(rs, shape + (C,)))
            pth = str(tmp_path / f"g{g}_s{s}.nii.gz")
            nifti.save(pth, v)
            paths[-1].append(pth)
            subj.append(v)
    X, sizes, nc, mask, tshape = gio.ingest_pinned(paths, threshold=0.15, dtype=dtype)
    assert sizes == (3, 2) and nc == C and tshape == (C,) + shape
    mats = [np.transpose(v, (3, 0, 1, 2)) for v in subj]
    ref_mask = gio.create_threshold_mask_from_matrices(mats, 0.15)
    assert np.array_equal(mask, ref_mask) and 0 < mask.sum() < mask.size
    groups, k = [], 0
    for n in sizes:
        g = [[m[c][ref_mask] for c in range(C)] for m in mats[k:k + n]]
        groups.append(gio.concat_assemble_group(g))
        k += n
    expect = gio.concat_flatten_all_groups(groups)
    assert X.shape == expect.shape == (sum(sizes) * C, int(mask.sum()))
    assert np.array_equal(X.numpy(), expect.astype(dtype))                # float32 sources: exact in both widths
