"""Assembly of the data matrix X for the GPU path.

The reference builds X in three copying steps -- per-subject condition slices, `concat_assemble_group`
(plspy/io/io.py:654-677: `np.array` over a Python list, condition-major within a group), `concat_flatten_all_groups`
(:680-699: `np.concatenate` + reshape) -- and `PLS(...)` then holds a pageable float64 array.  Here the same rows are
written ONCE, in the same order (group -> condition -> subject), straight into one pinned (page-locked) host buffer,
which is what `Engine` uploads in pipelined voxel ranges while the first kernels already run (engine.py
`_upload_pipelined`).  The image files themselves are read by `plspy_b200.nifti` (a numpy-only NIfTI-1 reader; nibabel
is not needed), and the reference's steps between the files and X -- time-first re-alignment, threshold mask, masking,
onset slices (plspy/io/io.py:10-651) -- are mirrored below under the reference's function names, so an ingest script
written against `plspy.io` runs against `plspy_b200.io` unchanged; `ingest_pinned` is the one-call form that goes from
per-subject image files to the pinned X without the intermediate Python lists of copies.

`dtype=np.float32` stores X in half the bytes: exact when the source images are float32 / int16 (as NIfTI data
are), half the host->device traffic, widened to float64 on the device (`plsb200_widen_f32_f64`); meant for
`analysis="device"`, since the host-side original analysis would have to widen a copy on the CPU.
"""
import os

import numpy as np
import torch

from . import exceptions, nifti


# ---------------------------------------------------------------------------------------------- image files
def open_single_image_in_dir(fpath):
    """plspy/io/io.py:49-72 -- one image file."""
    return nifti.load(f"{fpath}")


def open_images_in_dir(dir_path):
    """plspy/io/io.py:10-46 -- every file of a directory except `.hdr`, sorted by name -> (images, filenames)."""
    filenames = sorted(f.name for f in os.scandir(dir_path) if f.is_file() and not f.name.endswith(".hdr"))
    return [nifti.load(f"{dir_path}/{fl}") for fl in filenames], filenames


def open_images_from_paths_list(fpaths):
    """plspy/io/io.py:75-95."""
    return [open_single_image_in_dir(pth) for pth in fpaths]


def concat_images(images, **kwargs):
    """plspy/io/io.py:98-120 (`nibabel.concat_images`): 3-D volumes of one shape stacked along a new last axis."""
    arrs = [np.asarray(im.dataobj) for im in images]
    if any(a.shape != arrs[0].shape for a in arrs):
        raise ValueError("concat_images: the images do not have one shape")
    return nifti.NiftiImage(np.stack(arrs, axis=-1), images[0].affine, dict(images[0].header))


def read_dir_to_one_image(fpath, **kwargs):
    """plspy/io/io.py:123-155 -- a directory of volumes as one 4-D image."""
    return concat_images(open_images_in_dir(fpath)[0], **kwargs)


def open_multiple_imgs_from_dirs(dir_list, **kwargs):
    """plspy/io/io.py:158-204 -- one concatenated image per directory."""
    return [read_dir_to_one_image(d, **kwargs) for d in dir_list]


def extract_single_matrix(img):
    """plspy/io/io.py:207-231 -- the voxel array of an image (a trailing axis of length 1 dropped)."""
    mat = img.dataobj
    if mat.shape[-1] == 1:
        mat = mat.reshape(mat.shape[:-1])
    return mat


def extract_matrices_from_image_list(img_list):
    """plspy/io/io.py:234-261."""
    return [np.squeeze(extract_single_matrix(img)) for img in img_list]


def realign_axes_time_first(matrix):
    """plspy/io/io.py:264-283 -- (x, y, z, t) -> (t, x, y, z), as a view."""
    return np.transpose(matrix, (3, 0, 1, 2))


def extract_matrices_image_list_realign(img_list):
    """plspy/io/io.py:286-313 -> (time-first matrices, shape of the first)."""
    mats = [realign_axes_time_first(m) for m in extract_matrices_from_image_list(img_list)]
    return mats, mats[0].shape


# ---------------------------------------------------------------------------------------------- masks
def create_threshold_mask_from_matrices(matrices, threshold=0.15):
    """plspy/io/io.py:353-398 -- True where the mean image (over time, then over subjects) exceeds
    `threshold * (max - min) + min`.  The running sum replaces the reference's `np.array(matrices)` copy of the whole
    data set; the means are the same sums divided in the same order."""
    if threshold < 0 or threshold > 1:
        raise exceptions.OutOfRangeError(
            f"threshold must be greater than 0 or less than 1. Value passed in : {threshold}")
    mean_all = None
    for m in matrices:
        tm = np.mean(np.asarray(m), axis=0, dtype=np.float64)
        mean_all = tm if mean_all is None else mean_all + tm
    mean_all = mean_all / len(matrices)
    lo, hi = np.min(mean_all), np.max(mean_all)
    mask = mean_all > (threshold * (hi - lo) + lo)
    return np.asarray(mask)


def create_binary_mask_from_matrices(matrices):
    """plspy/io/io.py:316-350 -- True where no subject has a zero at any time point (`logical_and` over all volumes of
    all subjects, zeros counting as False)."""
    mats = np.array(matrices)
    mats_concat = mats.reshape((-1,) + mats.shape[2:])
    return np.logical_and.reduce(mats_concat, where=(mats_concat != 0), axis=0)


def apply_mask_matrices(matrices, mask):
    """plspy/io/io.py:427-460 -- every matrix flattened to its masked elements (time-major)."""
    return [m[np.broadcast_to(mask, m.shape)] for m in matrices]


def create_and_apply_mask_list(matrices, mask_type="threshold", threshold=0.15):
    """plspy/io/io.py:463-499."""
    if mask_type != "threshold":
        raise exceptions.NotImplementedError(f"Mask type {mask_type} is not implemented.")
    return np.array(apply_mask_matrices(matrices, create_threshold_mask_from_matrices(matrices, threshold=threshold)))


# ---------------------------------------------------------------------------------------------- onsets
def open_onsets_txt(filepath, tr):
    """plspy/io/io.py:502-535 -- one `.txt` per subject (columns = conditions) -> slice indices, conditions first."""
    files = sorted(f.path for f in os.scandir(filepath) if f.is_file() and f.name.endswith(".txt"))
    return [np.rint(np.loadtxt(f, dtype=float) / tr).astype(int).T for f in files]


def extract_onset_slices_single_subject(matrix, onsets, onset_length, tr, return_indiv=True):
    """plspy/io/io.py:538-602 -- per condition, the `onset_length * tr` volumes after every onset, stacked."""
    num_vols = int(np.rint(onset_length * tr))
    out = []
    for i in range(onsets.shape[0]):
        ind = (np.asarray(onsets[i]).reshape(-1, 1) + np.arange(num_vols)).reshape(-1)
        out.append(matrix[ind].reshape(-1, matrix.shape[-3], matrix.shape[-2], matrix.shape[-1]))
    return out if return_indiv else np.array(out)


def extract_onset_slices_list(matrices, onsets, onset_length, tr, use_one=False):
    """plspy/io/io.py:605-651."""
    return [extract_onset_slices_single_subject(matrices[i], onsets[0] if use_one else onsets[i], onset_length, tr)
            for i in range(len(matrices))]


def remap_vectorized_subject_to_4d(vector, mask, original_shape):
    """plspy/io/io.py:701-end -- masked, vectorised (time-major) values back into the (t, x, y, z) volume, zeros
    elsewhere; one boolean assignment instead of the reference's Python loop over voxels."""
    out = np.zeros(original_shape)
    out[:, np.asarray(mask, dtype=bool)] = np.asarray(vector).reshape(original_shape[0], -1)
    return out



def concat_assemble_group(matrices):
    """plspy/io/io.py:654-677 -- list (subjects) of lists (conditions) of arrays -> one group array, condition-major."""
    return np.array([matrices[i][j] for j in range(len(matrices[0])) for i in range(len(matrices))])


def concat_flatten_all_groups(groups_list):
    """plspy/io/io.py:680-699 -- concatenate the groups and flatten every row."""
    full = np.concatenate(groups_list, axis=0)
    return full.reshape(full.shape[0], -1)


def assemble_pinned(groups, dtype=np.float64):
    """X (N x p) in the reference's row order, written once into pinned host memory.

    groups: list (groups) of lists (subjects) of lists (conditions) of arrays; every array is one subject's data for
    one condition (any shape, flattened like `concat_flatten_all_groups` does).  Returns (X, groups_sizes,
    num_conditions): X a pinned torch tensor of `dtype` (float64 or float32) whose `.numpy()` view shares the memory."""
    if dtype not in (np.float64, np.float32):
        raise ValueError("dtype must be numpy float64 or float32")
    if not groups or not groups[0] or not groups[0][0]:
        raise ValueError("groups must be a non-empty list of subjects with a non-empty list of conditions")
    C = len(groups[0][0])
    p = int(np.asarray(groups[0][0][0]).size)
    sizes = tuple(len(g) for g in groups)
    N = sum(sizes) * C
    X = torch.empty((N, p), dtype=torch.float64 if dtype == np.float64 else torch.float32, pin_memory=torch.cuda.is_available())
    Xn = X.numpy()
    row = 0
    for g in groups:
        for c in range(C):
            for subj in g:
                if len(subj) != C:
                    raise ValueError("every subject needs one array per condition")
                a = np.asarray(subj[c])
                if a.size != p:
                    raise ValueError(f"subject array has {a.size} elements, expected {p}")
                Xn[row] = a.reshape(-1)          # converts to `dtype` while copying: the only copy of this row
                row += 1
    return X, sizes, C


def ingest_pinned(group_subject_paths, threshold=0.15, onsets=None, onset_length=None, tr=None, dtype=np.float64):
    """Image files -> pinned X in one call: the reference's script sequence `open_images_from_paths_list` ->
    `extract_matrices_image_list_realign` -> `create_threshold_mask_from_matrices` -> [`extract_onset_slices_list`] ->
    `apply_mask_matrices` -> `concat_assemble_group` -> `concat_flatten_all_groups` (plspy/io/io.py:75-699), with the
    masked rows written straight into the pinned matrix.

    group_subject_paths: list (groups) of lists (subjects) of 4-D image paths.  Without onsets every time point of an
    image is one condition; with `onsets` (per group, per subject: conditions x onsets slice indices, as
    `open_onsets_txt` returns them), `onset_length` and `tr` a condition is the stack of its onset windows.
    Returns (X pinned tensor, groups_sizes, num_conditions, mask, time-first shape of one subject)."""
    mats = [extract_matrices_image_list_realign(open_images_from_paths_list(paths))[0] for paths in group_subject_paths]
    mask = create_threshold_mask_from_matrices([m for g in mats for m in g], threshold=threshold)
    groups = []
    for gi, g in enumerate(mats):
        if onsets is not None:
            g = extract_onset_slices_list(g, onsets[gi], onset_length, tr)
        groups.append([[np.asarray(cond)[np.broadcast_to(mask, np.shape(cond))] for cond in subj] for subj in g])
    X, sizes, C = assemble_pinned(groups, dtype=dtype)
    return X, sizes, C, mask, mats[0][0].shape
