"""the whole-scene driver -- what the reference's test_scene.py:66-104 does for one sample of its data
loader, chained over the pieces of this package:

    model.update_sizes(input_dim, input_dim // 2^(levels-1))          test_scene.py:76-77
    output_sdf, output_occs = model(inputs, loss_weights)             :82          (native generator, one C-ABI call)
    remove padding (rows at or beyond sample['orig_dims'])            :88-95
    data_util.save_predictions(output, name, inputs, None, None, [pred], None, world2grid, truncation)   :98
        -> <name>input-mesh.ply, <name>pred-mesh.ply  (data_util.py:248-281: -inf filled dense grids, marching cubes at
           isovalue 0, truncation - 0.1, thresh 10, the reference's binary PLY)

`sample` is one collated batch-1 sample of scene_io.SceneDataset / scene_io.collate (same dict the reference's loader
yields).  Everything numeric runs on the GPU (generator, dense scatter, marching-cubes grid walk); the host does the
order-defining vertex merge and the file write."""
import os
import time

import numpy as np
import torch

from . import mesh


def sparse_to_dense_device(locs, values, dims_zyx, default_val=float('-inf')):
    """data_util.sparse_to_dense_np (:43-54) on the device: locs [N,>=3] (z,y,x), values [N] / [N,1]."""
    d0, d1, d2 = (int(v) for v in dims_zyx)
    dense = torch.full((d0, d1, d2), default_val, dtype=torch.float32, device=values.device)
    li = locs.to(values.device).long()
    dense[li[:, 0], li[:, 1], li[:, 2]] = values.reshape(-1).float()
    return dense


def save_predictions(output_path, names, inputs, target_for_sdf, target_for_occs, output_sdf, output_occs, world2grids,
                     truncation, thresh=1):
    """data_util.save_predictions (:248-286) for the call test_scene.py:98 makes (no targets, no per-level occupancy):
    per sample k the input mesh and the predicted mesh as `.ply`.  inputs = [locs [N,4] (z,y,x,b), feats [N,1]];
    output_sdf[k] = [locs [M,>=3], sdf [M]] or None.  Tensors may live on the host or the GPU."""
    if target_for_sdf is not None or target_for_occs is not None or output_occs is not None:
        raise ValueError('save_predictions: only the test_scene.py form (targets None, output_occs None) is in scope; '
                         'the target / per-level point-cloud dumps belong to train.py')
    os.makedirs(output_path, exist_ok=True)
    dev = torch.device('cuda', torch.cuda.current_device())
    as_t = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(dev)
    in_locs, in_feats = as_t(inputs[0]), as_t(inputs[1])
    hi = in_locs.max(0).values[:3] if in_locs.shape[0] else torch.zeros(3, dtype=torch.long, device=dev)
    first = output_sdf[0]
    if first is not None and len(first[0]):
        hi = torch.maximum(hi, as_t(first[0]).max(0).values[:3])                      # data_util.py:256 (sample 0's prediction)
    dims = [int(v) + 1 for v in hi.tolist()]
    trunc = truncation - 0.1
    written = []
    for k, name in enumerate(names):
        sel = in_locs[:, -1] == k
        dense = sparse_to_dense_device(in_locs[sel][:, :-1], in_feats[sel], dims)
        p = os.path.join(output_path, name + 'input-mesh.ply')
        mesh.marching_cubes(dense, None, 0, trunc, 10, p)
        written.append(p)
        if output_sdf[k] is not None:
            dense = sparse_to_dense_device(as_t(output_sdf[k][0])[:, :3], as_t(output_sdf[k][1]), dims)
            p = os.path.join(output_path, name + 'pred-mesh.ply')
            mesh.marching_cubes(dense, None, 0, trunc, 10, p)
            written.append(p)
    return written


def run_scene(model, sample, output_path=None, truncation=3.0, num_hierarchy_levels=4, loss_weights=None, timings=None):
    """test_scene.py:66-104 for one batch-1 sample.  Returns (inputs, output_sdf) after pad removal: inputs =
    [locs LongTensor [n,4] (host), feats [n,1] (device)], output_sdf = [locs [m,4], sdf [m,1]] (device) -- and, when
    output_path is given, writes the two meshes.  timings (dict) receives device-synchronised stage times in ms."""
    if loss_weights is None:
        loss_weights = np.ones(num_hierarchy_levels + 1, dtype=np.float32)
    dev = next(model.parameters()).device
    tick = time.perf_counter()

    def lap(key):
        nonlocal tick
        if timings is not None:
            torch.cuda.synchronize(dev)
            now = time.perf_counter()
            timings[key] = (now - tick) * 1e3
            tick = now
    inputs = [sample['input'][0], sample['input'][1].to(dev)]
    input_dim = np.array(sample['sdf'].shape[2:])
    model.update_sizes(input_dim, input_dim // (2 ** (num_hierarchy_levels - 1)))
    with torch.no_grad():
        output_sdf, output_occs = model(inputs, loss_weights)
    lap('forward_ms')
    dims = sample['orig_dims'][0]
    d = [int(v) for v in dims]
    if len(output_sdf[0]):
        l = output_sdf[0]
        keep = (l[:, 0] < d[0]) & (l[:, 1] < d[1]) & (l[:, 2] < d[2])
        output_sdf = [l[keep], output_sdf[1][keep]]
    l = inputs[0]
    keep = (l[:, 0] < d[0]) & (l[:, 1] < d[1]) & (l[:, 2] < d[2])
    inputs = [l[keep], inputs[1][keep.to(inputs[1].device)]]
    lap('pad_removal_ms')
    if output_path is not None:
        pred = [None]
        if len(output_sdf[0]):
            pred[0] = [output_sdf[0], output_sdf[1].squeeze(1)]
        save_predictions(output_path, sample['name'], inputs, None, None, pred, None, sample['world2grid'], truncation)
        lap('meshes_ms')
    return inputs, output_sdf
