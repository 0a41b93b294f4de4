"""readers for the reference's binary scene formats and the scene -> model-input preparation of its
data loader, so real Matterport scans can be fed to GenModel.  Formats (writer: datagen VoxelGrid.h:120-158,199-220;
reference readers: torch/data_util.py:112-144):

  .sdf   u64 dimx, dimy, dimz | f32 voxelsize | 16 x f32 world2grid (row major) | u64 n | n x 3 u32 (x,y,z) | n x f32 sdf
  .knw   same header | dimz*dimy*dimx u8 "known" flags

Coordinates are flipped to (z,y,x) and distances divided by the voxel size (data_util.py:75,78).  `prepare_scene`
restates scene_dataloader.py:79-104: optional height cap, padding of the extent to a multiple of 32, truncation mask.
numpy.fromfile is used instead of struct.unpack (same bytes, ~100x faster on million-voxel scenes)."""
import numpy as np
import torch

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
