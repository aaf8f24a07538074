"""Assembly of the data matrix X for the GPU path.

The reference builds X in three copying steps -- per-subject condition slices, `concat_assemble_group`
(plspy/io/io.py:654-677: `np.array` over a Python list, condition-major within a group), `concat_flatten_all_groups`
(:680-699: `np.concatenate` + reshape) -- and `PLS(...)` then holds a pageable float64 array.  Here the same rows are
written ONCE, in the same order (group -> condition -> subject), straight into one pinned (page-locked) host buffer,
which is what `Engine` uploads in pipelined voxel ranges while the first kernels already run (engine.py
`_upload_pipelined`).  NIfTI parsing itself (nibabel) is outside the resampling path and not provided.

`dtype=np.float32` stores X in half the bytes: exact when the source images are float32 / int16 (as NIfTI data
are), half the host->device traffic, widened to float64 on the device (`plsb200_widen_f32_f64`); meant for
`analysis="device"`, since the host-side original analysis would have to widen a copy on the CPU.
"""
import numpy as np
import torch


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
