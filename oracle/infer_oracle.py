"""ORACLE (test infrastructure only) -- restatement of the sliding-window inference of the reference.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.

The reference delegates patch extraction and aggregation to torchio==0.19.6 (requirements.txt), which is
not installable here (no network); its published algorithm is restated:

* ``grid_locations``   torchio/data/sampler/grid.py ``GridSampler._get_patches_locations`` with
                       ``patch_overlap = (0, 0, 0)`` (call site ref:src/data_module.py:171-176): per axis
                       ``range(0, size + 1 - patch, patch)``; if the last index is not ``size - patch`` it is
                       appended; the Cartesian product is de-duplicated and sorted lexicographically
                       (``np.unique(..., axis=0)``).
* ``aggregate``        torchio/data/inference/aggregator.py ``GridAggregator.add_batch`` in the default
                       ``overlap_mode='crop'`` with zero overlap: ``output[:, i0:i1, j0:j1, k0:k1] = patch`` in
                       sampler order (call site ref:src/model.py:315-325) -- the later patch wins.
* ``predict_volume``   ref:src/model.py:315-327 ``predict_step`` (metrics / file output stripped).

PARITY PIN: unpinned by the reference (no tests); pinned by the hand-checkable KAT of SURVEY.md section 8d
(160x192x160 with 64^3 patches -> 27 patches at {0,64,96} x {0,64,128} x {0,64,96}).
"""
import numpy as np
import torch


def grid_locations(shape, patch):
    indices = []
    for size, p in zip(shape, patch):
        end = size + 1 - p
        idx = list(range(0, end, p))
        if idx[-1] != size - p:
            idx.append(size - p)
        indices.append(idx)
    ini = np.array(np.meshgrid(*indices)).reshape(len(shape), -1).T
    ini = np.unique(ini, axis=0)
    return [tuple(int(v) for v in row) for row in ini]


def aggregate(out, patches, locations):
    """out (C,D,H,W); patches: iterable of (C,pd,ph,pw) in sampler order."""
    for p, (z, y, x) in zip(patches, locations):
        out[:, z:z + p.shape[1], y:y + p.shape[2], x:x + p.shape[3]] = p
    return out


@torch.no_grad()
def predict_volume(gen, volume, patch=(64, 64, 64), batch=8, out_channels=6):
    """gen: any callable (N,C,pd,ph,pw) -> (N,6,pd,ph,pw); volume (C,D,H,W) torch tensor."""
    locs = grid_locations(tuple(volume.shape[1:]), patch)
    out = torch.zeros((out_channels,) + tuple(volume.shape[1:]), dtype=torch.float32, device=volume.device)
    for i in range(0, len(locs), batch):
        group = locs[i:i + batch]
        x = torch.stack([volume[:, z:z + patch[0], y:y + patch[1], w:w + patch[2]] for (z, y, w) in group])
        y_hat = gen(x).float()
        aggregate(out, list(y_hat), group)
    return out
