"""Mesh extraction after the forward pass: host-side mirror of the reference's
torch/marching_cubes/marching_cubes.py on top of csrc/mcubes.cu.

    verts, faces = run_marching_cubes(tsdf, isovalue=0, truncation=3, thresh=10)        # == marching_cubes_cpp.run_marching_cubes
    marching_cubes(tsdf, None, isovalue, truncation, thresh, 'pred-mesh.ply')           # == data_util.py:281

The grid walk (all floating-point work) runs on the GPU; the first-come vertex merge that defines the vertex numbering
runs on the host (sgnn_mc_merge_host).  Same vertices (bit for bit), same faces, same order as the reference.
Vertex colours: the reference paints every vertex with the colour of its cell's voxel; with colors=None (how
data_util.py calls it) that is a constant grey 220, which is what this mirror writes.  Per-voxel colours are not
implemented (NotImplementedError)."""
import ctypes as C
import os

import numpy as np
import torch

from ._lib import lib, check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def merge_triangles(tris):
    """Host half: tris float32 [T,3,3] (numpy / CPU tensor) -> (verts float32 [V,3], faces int32 [F,3]) numpy arrays."""
    t = np.ascontiguousarray(tris.numpy() if isinstance(tris, torch.Tensor) else tris, dtype=np.float32).reshape(-1, 9)
    n = t.shape[0]
    verts = np.empty((max(3 * n, 1), 3), dtype=np.float32)
    faces = np.empty((max(n, 1), 3), dtype=np.int32)
    nv, nf = C.c_int64(0), C.c_int64(0)
    check(lib.sgnn_mc_merge_host(t.ctypes.data_as(C.c_void_p), n, verts.ctypes.data_as(C.c_void_p),
                                 faces.ctypes.data_as(C.c_void_p), C.byref(nv), C.byref(nf)), 'sgnn_mc_merge_host')
    return verts[:nv.value].copy(), faces[:nf.value].copy()


def triangle_soup(tsdf, isovalue=0.0, truncation=3.0, thresh=10.0):
    """Device half: dense TSDF [n0,n1,n2] (CUDA fp32; -inf = unobserved) -> triangles [T,3,3] CUDA fp32, reference order."""
    if not tsdf.is_cuda:
        raise RuntimeError('sgnn_b200.mesh: the TSDF must be a CUDA tensor (no CPU fallback)')
    t = tsdf.float().contiguous()
    n0, n1, n2 = (int(v) for v in t.shape)
    total = n0 * n1 * n2
    offs = torch.empty(total + 1, dtype=torch.int32, device=t.device)
    sb = int(lib.sgnn_mc_scratch_bytes(n0, n1, n2))
    scratch = torch.empty(sb, dtype=torch.uint8, device=t.device)
    args = (C.c_void_p(t.data_ptr()), n0, n1, n2, C.c_float(isovalue), C.c_float(truncation), C.c_float(thresh))
    check(lib.sgnn_mc_count(*args, C.c_void_p(offs.data_ptr()), C.c_void_p(scratch.data_ptr()), sb, _stream()), 'sgnn_mc_count')
    n_tri = int(offs[total].item())
    tris = torch.empty((n_tri, 3, 3), dtype=torch.float32, device=t.device)
    if n_tri:
        check(lib.sgnn_mc_emit(*args, C.c_void_p(offs.data_ptr()), C.c_void_p(tris.data_ptr()), _stream()), 'sgnn_mc_emit')
    return tris


def run_marching_cubes(tsdf, colors=None, isovalue=0.0, truncation=3.0, thresh=10.0):
    """-> (vertices FloatTensor [V,3] (x,y,z), vertex colours ByteTensor [V,3], faces IntTensor [F,3]) on the CPU,
    like marching_cubes_cpp.run_marching_cubes (marching_cubes.cpp:480-517)."""
    if colors is not None:
        raise NotImplementedError('per-voxel colours are not implemented; data_util.py passes None (grey 220)')
    verts, faces = merge_triangles(triangle_soup(tsdf, isovalue, truncation, thresh).cpu())
    v = torch.from_numpy(verts)
    return v, torch.full((v.shape[0], 3), 220, dtype=torch.uint8), torch.from_numpy(faces)


def sparse_sdf_to_mesh(locs, sdf, dims_zyx, truncation=3.0, thresh=10.0, output_filename=None):
    """The mesh of a sparse TSDF prediction, as data_util.save_predictions does it (data_util.py:278-281): scatter the
    values into a dense grid filled with -inf (sparse_to_dense_np, :43-54), marching cubes at isovalue 0 with
    truncation - 0.1.  locs [N,>=3] (z,y,x[,b]) and sdf [N] / [N,1] are the generator's outputs (CUDA tensors)."""
    if not sdf.is_cuda:
        raise RuntimeError('sgnn_b200.mesh: the prediction must be CUDA tensors (no CPU fallback)')
    d0, d1, d2 = (int(v) for v in dims_zyx)
    dense = torch.full((d0, d1, d2), float('-inf'), dtype=torch.float32, device=sdf.device)
    li = locs.to(sdf.device).long()
    dense[li[:, 0], li[:, 1], li[:, 2]] = sdf.reshape(-1).float()
    v, c, f = run_marching_cubes(dense, None, 0.0, truncation - 0.1, thresh)
    if output_filename is not None:
        save_mesh(v.numpy(), c.numpy(), f.numpy(), output_filename)
    return v, c, f


def save_mesh(verts, colors, faces, output_file):
    """ASCII .ply / .obj writer (marching_cubes.py:9-26 writes .obj with vertex colours, .ply via plyfile)."""
    verts, colors, faces = np.asarray(verts), np.asarray(colors), np.asarray(faces)
    ext = os.path.splitext(output_file)[1]
    with open(output_file, 'w') as f:
        if ext == '.obj':
            for v, c in zip(verts, colors):
                f.write('v %f %f %f %d %d %d\n' % (v[0], v[1], v[2], c[0], c[1], c[2]))
            f.write('g foo\n')
            for t in faces:
                f.write('f %d %d %d\n' % (t[0] + 1, t[1] + 1, t[2] + 1))
            f.write('g\n')
        else:
            f.write('ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n'
                    'property uchar red\nproperty uchar green\nproperty uchar blue\nelement face %d\n'
                    'property list uchar int vertex_indices\nend_header\n' % (verts.shape[0], faces.shape[0]))
            for v, c in zip(verts, colors):
                f.write('%f %f %f %d %d %d\n' % (v[0], v[1], v[2], c[0], c[1], c[2]))
            for t in faces:
                f.write('3 %d %d %d\n' % (t[0], t[1], t[2]))


def marching_cubes(tsdf, colors, isovalue, truncation, thresh, output_filename):
    """Signature of the reference's marching_cubes.marching_cubes (marching_cubes.py:28-35)."""
    v, c, f = run_marching_cubes(tsdf if tsdf.is_cuda else tsdf.cuda(), colors, isovalue, truncation, thresh)
    save_mesh(v.numpy(), c.numpy(), f.numpy(), output_filename)
