"""readers for the reference's binary scene formats and the scene -> model-input preparation of its
data loader, so real Matterport scans can be fed to GenModel.  Formats (writer: datagen VoxelGrid.h:120-158,199-220;
reference readers: torch/data_util.py:63-144):

  .sdf   u64 dimx, dimy, dimz | f32 voxelsize | 16 x f32 world2grid (row major) | u64 n | n x 3 u32 (x,y,z) | n x f32 sdf
  .knw   same header | dimz*dimy*dimx u8 "known" flags
  .sdfs  (training chunk, data_util.py:63-109) same header | input block (u64 n | n x 3 u32 | n x f32) | target block
         (same) | u64 dimx*dimy*dimz + that many u8 known flags | 3 hierarchy blocks (same sparse form) at 1/2, 1/4, 1/8

`load_train_file`, `collate` and `SceneDataset` restate data_util.py:63-109 and scene_dataloader.py:13-120 (same return
values, same dict keys); tests/test_scene_io.py runs the reference's own functions next to them.

Coordinates are flipped to (z,y,x) and distances divided by the voxel size (data_util.py:75,78).  `prepare_scene`
restates scene_dataloader.py:79-104: optional height cap, padding of the extent to a multiple of 32, truncation mask.
numpy.fromfile is used instead of struct.unpack (same bytes, ~100x faster on million-voxel scenes)."""
import numpy as np
import torch
import torch.utils.data

_HDR = np.dtype([('dims', '<u8', 3), ('voxelsize', '<f4'), ('world2grid', '<f4', 16)])


def _header(f):
    h = np.fromfile(f, dtype=_HDR, count=1)
    if h.shape[0] != 1:
        raise IOError('truncated header')
    dimx, dimy, dimz = (int(v) for v in h['dims'][0])
    return dimx, dimy, dimz, float(h['voxelsize'][0]), h['world2grid'][0].reshape(4, 4).astype(np.float32)


def load_scene(path):
    """-> ([locs int32 [n,3] (z,y,x), sdf float32 [n] in voxel units], [dimz, dimy, dimx], world2grid 4x4)"""
    with open(path, 'rb') as f:
        dimx, dimy, dimz, vs, w2g = _header(f)
        n = int(np.fromfile(f, dtype='<u8', count=1)[0])
        xyz = np.fromfile(f, dtype='<u4', count=3 * n)
        sdf = np.fromfile(f, dtype='<f4', count=n)
    if xyz.shape[0] != 3 * n or sdf.shape[0] != n:
        raise IOError('%s: truncated payload (%d voxels announced)' % (path, n))
    locs = np.ascontiguousarray(xyz.reshape(n, 3)[:, ::-1]).astype(np.int32)
    return [locs, (sdf / np.float32(vs)).astype(np.float32)], [dimz, dimy, dimx], w2g


def _sparse_block(f, path):
    """u64 n | n x 3 u32 (x,y,z) | n x f32 -> (locs int32 [n,3] (z,y,x), values float32 [n])"""
    n = np.fromfile(f, dtype='<u8', count=1)
    if n.shape[0] != 1:
        raise IOError('%s: truncated block header' % path)
    n = int(n[0])
    xyz = np.fromfile(f, dtype='<u4', count=3 * n)
    val = np.fromfile(f, dtype='<f4', count=n)
    if xyz.shape[0] != 3 * n or val.shape[0] != n:
        raise IOError('%s: truncated payload (%d voxels announced)' % (path, n))
    return np.ascontiguousarray(xyz.reshape(n, 3)[:, ::-1]).astype(np.int32), val


def sparse_to_dense_np(locs, values, dimx, dimy, dimz, default_val):
    """data_util.py:43-54: scatter (z,y,x) rows into a dense [dimz,dimy,dimx(,nf)] grid filled with default_val."""
    nf = 1 if values.ndim == 1 else values.shape[1]
    dense = np.full([dimz, dimy, dimx, nf], default_val, dtype=values.dtype)
    dense[locs[:, 0], locs[:, 1], locs[:, 2], :] = values.reshape(-1, nf)
    return dense if nf > 1 else dense.reshape([dimz, dimy, dimx])


def load_train_file(path):
    """data_util.py:63-109 (.sdfs training chunk) -> ([input_locs int32 [n,3] (z,y,x), input_sdfs float32 [n] in voxels],
    target_sdfs float32 [dimz,dimy,dimx] (-inf = empty), [dimz,dimy,dimx], world2grid, target_known uint8
    [dimz,dimy,dimx], hierarchy = 3 dense float32 grids, coarsest (1/8) first)."""
    with open(path, 'rb') as f:
        dimx, dimy, dimz, vs, w2g = _header(f)
        vs = np.float32(vs)
        in_locs, in_sdf = _sparse_block(f, path)
        in_sdf = (in_sdf / vs).astype(np.float32)
        t_locs, t_sdf = _sparse_block(f, path)
        target = sparse_to_dense_np(t_locs, (t_sdf / vs).astype(np.float32)[:, None], dimx, dimy, dimz, -float('inf'))
        n = np.fromfile(f, dtype='<u8', count=1)
        if n.shape[0] != 1 or int(n[0]) != dimx * dimy * dimz:
            raise IOError('%s: known-grid size does not match the extent' % path)
        known = np.fromfile(f, dtype=np.uint8, count=dimx * dimy * dimz)
        if known.shape[0] != dimx * dimy * dimz:
            raise IOError('%s: truncated known grid' % path)
        known = known.reshape(dimz, dimy, dimx)
        hierarchy, factor = [], 2
        for _ in range(3):
            h_locs, h_val = _sparse_block(f, path)
            hierarchy.append(sparse_to_dense_np(h_locs, (h_val / vs).astype(np.float32)[:, None], dimx // factor,
                                                dimy // factor, dimz // factor, -float('inf')))
            factor *= 2
    hierarchy.reverse()
    return [in_locs, in_sdf], target, [dimz, dimy, dimx], w2g, known, hierarchy


def save_train_file(path, input_locs, input_sdf, target_locs, target_sdf, known, hierarchy, dims_zyx, voxelsize=0.02,
                    world2grid=None):
    """Writer of the .sdfs layout (tests, synthetic chunks).  locs (z,y,x), values in voxel units; hierarchy = 3
    (locs, values) pairs at 1/2, 1/4, 1/8 resolution in file order (finest first)."""
    w2g = np.eye(4, dtype=np.float32) if world2grid is None else np.asarray(world2grid, dtype=np.float32)

    def block(f, locs, vals):
        locs = np.asarray(locs).reshape(-1, 3)
        np.array([locs.shape[0]], dtype='<u8').tofile(f)
        np.ascontiguousarray(locs[:, ::-1]).astype('<u4').tofile(f)
        (np.asarray(vals, dtype=np.float32) * np.float32(voxelsize)).astype('<f4').tofile(f)

    with open(path, 'wb') as f:
        np.array([dims_zyx[2], dims_zyx[1], dims_zyx[0]], dtype='<u8').tofile(f)
        np.array([voxelsize], dtype='<f4').tofile(f)
        w2g.reshape(-1).astype('<f4').tofile(f)
        block(f, input_locs, input_sdf)
        block(f, target_locs, target_sdf)
        known = np.ascontiguousarray(known, dtype=np.uint8)
        np.array([known.size], dtype='<u8').tofile(f)
        known.tofile(f)
        for locs, vals in hierarchy:
            block(f, locs, vals)


def load_scene_known(path):
    with open(path, 'rb') as f:
        dimx, dimy, dimz, _, _ = _header(f)
        k = np.fromfile(f, dtype=np.uint8, count=dimx * dimy * dimz)
    if k.shape[0] != dimx * dimy * dimz:
        raise IOError('%s: truncated known grid' % path)
    return k.reshape(dimz, dimy, dimx)


def save_scene(path, locs_zyx, sdf_voxels, dims_zyx, voxelsize=0.02, world2grid=None):
    """Writer of the same format (tests, synthetic scenes)."""
    w2g = np.eye(4, dtype=np.float32) if world2grid is None else np.asarray(world2grid, dtype=np.float32)
    with open(path, 'wb') as f:
        np.array([dims_zyx[2], dims_zyx[1], dims_zyx[0]], dtype='<u8').tofile(f)
        np.array([voxelsize], dtype='<f4').tofile(f)
        w2g.reshape(-1).astype('<f4').tofile(f)
        np.array([locs_zyx.shape[0]], dtype='<u8').tofile(f)
        np.ascontiguousarray(np.asarray(locs_zyx)[:, ::-1]).astype('<u4').tofile(f)
        (np.asarray(sdf_voxels, dtype=np.float32) * np.float32(voxelsize)).astype('<f4').tofile(f)


def prepare_scene(locs, sdf, dims, truncation=3.0, max_input_height=128, num_hierarchy_levels=4, batch_index=0):
    """scene_dataloader.py:79-104 + collate (:13-36) for one scene: returns (coords LongTensor [n,4] (z,y,x,b),
    feats FloatTensor [n,1], padded dims [3]) ready for GenModel.update_sizes(dims, dims // 8) + forward."""
    dims = np.array(dims, dtype=np.int64)
    if max_input_height > 0 and dims[0] > max_input_height:
        keep = locs[:, 0] < max_input_height
        locs, sdf = locs[keep], sdf[keep]
        dims[0] = max_input_height
    unit = (2 ** (num_hierarchy_levels - 1)) * 4
    dims = (dims + unit - 1) // unit * unit
    keep = np.abs(sdf) < truncation
    locs, sdf = locs[keep], sdf[keep]
    coords = np.concatenate([locs.astype(np.int64), np.full((locs.shape[0], 1), batch_index, dtype=np.int64)], 1)
    return (torch.from_numpy(np.ascontiguousarray(coords)), torch.from_numpy(sdf.astype(np.float32)[:, None].copy()),
            [int(v) for v in dims])


def collate(batch):
    """scene_dataloader.py:13-36: concatenates the samples' sparse inputs with a batch-index column, stacks the dense
    fields.  Same keys and tensor types as the reference's collate."""
    locs, feats = [], []
    for b, x in enumerate(batch):
        l = x['input'][0]
        locs.append(torch.cat([l, torch.full((l.shape[0], 1), b, dtype=torch.long)], 1))
        feats.append(x['input'][1])
    known = torch.stack([x['known'] for x in batch]) if batch[0]['known'] is not None else None
    hierarchy = None
    if batch[0]['hierarchy'] is not None:
        hierarchy = [torch.stack([x['hierarchy'][h] for x in batch]) for h in range(len(batch[0]['hierarchy']))]
    return {'name': [x['name'] for x in batch], 'input': [torch.cat(locs), torch.cat(feats)],
            'sdf': torch.stack([x['sdf'] for x in batch]), 'world2grid': torch.stack([x['world2grid'] for x in batch]),
            'known': known, 'hierarchy': hierarchy, 'orig_dims': torch.stack([x['orig_dims'] for x in batch])}


class SceneDataset(torch.utils.data.Dataset):
    """scene_dataloader.py:39-120.  target_path == '' -> `.sdfs` training chunks (pre-computed hierarchy kept, its
    coarsest levels dropped when num_hierarchy_levels < 4); otherwise (input .sdf, target .sdf + .knw) scene pairs,
    height-capped and padded to a multiple of 4 * 2^(levels-1) like test_scene.py needs them."""
    UP_AXIS = 0

    def __init__(self, files, input_dim, truncation, num_hierarchy_levels, max_input_height, num_overfit=0, target_path=''):
        import os
        assert num_hierarchy_levels <= 4
        self.is_chunks = target_path == ''
        if self.is_chunks:
            self.files = [f for f in files if os.path.isfile(f)]
        else:
            self.files = [(f, os.path.join(target_path, os.path.basename(f))) for f in files
                          if os.path.isfile(f) and os.path.isfile(os.path.join(target_path, os.path.basename(f)))]
        self.input_dim, self.truncation = input_dim, truncation
        self.num_hierarchy_levels, self.max_input_height = num_hierarchy_levels, max_input_height
        if num_overfit > 0:
            self.files = self.files * max(1, num_overfit // len(self.files))

    def __len__(self):
        return len(self.files)

    def __getitem__(self, idx):
        import os
        file = self.files[idx]
        if self.is_chunks:
            name = os.path.splitext(os.path.basename(file))[0]
            inputs, targets, dims, world2grid, known, hierarchy = load_train_file(file)
            if self.num_hierarchy_levels < 4:
                hierarchy = hierarchy[4 - self.num_hierarchy_levels:]
            orig_dims = torch.LongTensor(targets.shape)
        else:
            name = os.path.splitext(os.path.basename(file[0]))[0]
            inputs, dims, world2grid = load_scene(file[0])
            (t_locs, t_sdf), dims, world2grid = load_scene(file[1])
            known = load_scene_known(os.path.splitext(file[1])[0] + '.knw')
            targets = sparse_to_dense_np(t_locs, t_sdf[:, None], dims[2], dims[1], dims[0], -float('inf'))
            hierarchy = None
            orig_dims = torch.LongTensor(targets.shape)
            unit = (2 ** (self.num_hierarchy_levels - 1)) * 4
            mh = self.max_input_height
            max_dim = np.array(targets.shape)
            if mh > 0 and max_dim[self.UP_AXIS] > mh:
                max_dim[self.UP_AXIS] = mh
                keep = inputs[0][:, self.UP_AXIS] < mh
                inputs = [inputs[0][keep], inputs[1][keep]]
            max_dim = (max_dim + unit - 1) // unit * unit
            # as the reference writes it: the slice bound is max_input_height itself, so 0 ("no cap") copies nothing
            h = min(mh, targets.shape[0])
            padded = np.full(tuple(max_dim), -float('inf'), dtype=np.float32)
            padded[:h, :targets.shape[1], :targets.shape[2]] = targets[:mh, :, :]
            targets = padded
            known_pad = np.full(tuple(max_dim), 255, dtype=np.uint8)
            known_pad[:min(mh, known.shape[0]), :known.shape[1], :known.shape[2]] = known[:mh, :, :]
            known = known_pad
        mask = np.abs(inputs[1]) < self.truncation
        inputs = [torch.from_numpy(inputs[0][mask]).long(), torch.from_numpy(inputs[1][mask][:, np.newaxis]).float()]
        if hierarchy is not None:
            hierarchy = [torch.from_numpy(h[np.newaxis, :]) for h in hierarchy]
        return {'name': name, 'input': inputs, 'sdf': torch.from_numpy(targets[np.newaxis, :]),
                'world2grid': torch.from_numpy(world2grid), 'known': torch.from_numpy(known[np.newaxis, :]),
                'hierarchy': hierarchy, 'orig_dims': orig_dims}
