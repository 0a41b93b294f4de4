"""Mesh extraction after the forward pass: host-side mirror of the reference's
torch/marching_cubes/marching_cubes.py on top of csrc/mcubes.cu.

    verts, colors, faces = run_marching_cubes(tsdf, colors, isovalue=0, truncation=3, thresh=10)   # == marching_cubes_cpp.run_marching_cubes
    marching_cubes(tsdf, None, isovalue, truncation, thresh, 'pred-mesh.ply')                      # == data_util.py:281

The grid walk (all floating-point work) runs on the GPU; the first-come vertex merge that defines the vertex numbering
runs on the host (sgnn_mc_merge_host_src).  Same vertices (bit for bit), same faces, same order as the reference.
Vertex colours: the reference paints the three vertices of a triangle with the colour of the voxel of the cell that emitted
it (marching_cubes.cpp:228-231,255-257) and a merged vertex keeps the colour of its first-come copy (:399-431); with
colors=None (how data_util.py calls it) that is a constant grey 220 (marching_cubes.py:29-30).
`.ply` files are written as the reference's save_to_ply does (marching_cubes.cpp:519-560): binary_little_endian, packed
15-byte vertices (3 x float + 3 x uchar), faces as uchar 3 + 3 x int -- byte-identical output (tests/test_mesh.py);
anything else goes through the .obj branch of marching_cubes.py:9-18."""
import ctypes as C
import os

import numpy as np
import torch

from ._lib import lib, check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def merge_triangles(tris, return_source=False):
    """Host half: tris float32 [T,3,3] (numpy / CPU tensor) -> (verts float32 [V,3], faces int32 [F,3]) numpy arrays
    (+ vert_src int32 [V]: the soup vertex, 3*triangle + corner, each merged vertex was created from)."""
    t = np.ascontiguousarray(tris.numpy() if isinstance(tris, torch.Tensor) else tris, dtype=np.float32).reshape(-1, 9)
    n = t.shape[0]
    verts = np.empty((max(3 * n, 1), 3), dtype=np.float32)
    faces = np.empty((max(n, 1), 3), dtype=np.int32)
    src = np.empty(max(3 * n, 1), dtype=np.int32)
    nv, nf = C.c_int64(0), C.c_int64(0)
    check(lib.sgnn_mc_merge_host_src(t.ctypes.data_as(C.c_void_p), n, verts.ctypes.data_as(C.c_void_p),
                                     faces.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p), C.byref(nv),
                                     C.byref(nf)), 'sgnn_mc_merge_host_src')
    if return_source:
        return verts[:nv.value].copy(), faces[:nf.value].copy(), src[:nv.value].copy()
    return verts[:nv.value].copy(), faces[:nf.value].copy()


def triangle_soup(tsdf, isovalue=0.0, truncation=3.0, thresh=10.0, return_cells=False):
    """Device half: dense TSDF [n0,n1,n2] (CUDA fp32; -inf = unobserved) -> triangles [T,3,3] CUDA fp32, reference order
    (+ tri_cell int32 [T]: the linear (z,y,x) index of the cell each triangle was emitted for)."""
    if not tsdf.is_cuda:
        raise RuntimeError('sgnn_b200.mesh: the TSDF must be a CUDA tensor (no CPU fallback)')
    t = tsdf.float().contiguous()
    n0, n1, n2 = (int(v) for v in t.shape)
    total = n0 * n1 * n2
    with torch.cuda.device(t.device):
        offs = torch.empty(total + 1, dtype=torch.int32, device=t.device)
        sb = int(lib.sgnn_mc_scratch_bytes(n0, n1, n2))
        scratch = torch.empty(sb, dtype=torch.uint8, device=t.device)
        args = (C.c_void_p(t.data_ptr()), n0, n1, n2, C.c_float(isovalue), C.c_float(truncation), C.c_float(thresh))
        check(lib.sgnn_mc_count(*args, C.c_void_p(offs.data_ptr()), C.c_void_p(scratch.data_ptr()), sb, _stream()), 'sgnn_mc_count')
        n_tri = int(offs[total].item())
        tris = torch.empty((n_tri, 3, 3), dtype=torch.float32, device=t.device)
        cells = torch.empty(n_tri, dtype=torch.int32, device=t.device)
        if n_tri:
            check(lib.sgnn_mc_emit(*args, C.c_void_p(offs.data_ptr()), C.c_void_p(tris.data_ptr()), _stream()), 'sgnn_mc_emit')
            if return_cells:
                check(lib.sgnn_mc_tri_cells(C.c_void_p(offs.data_ptr()), total, C.c_void_p(cells.data_ptr()), _stream()),
                      'sgnn_mc_tri_cells')
    return (tris, cells) if return_cells else tris


def vertex_colors(colors, tri_cell, vert_src):
    """colors uint8 [n0,n1,n2,3] (per voxel) -> uint8 [V,3]: colour of the cell of the triangle of the first-come copy."""
    c = np.ascontiguousarray(colors.cpu().numpy() if isinstance(colors, torch.Tensor) else colors, dtype=np.uint8).reshape(-1, 3)
    return c[np.asarray(tri_cell, dtype=np.int64)[np.asarray(vert_src, dtype=np.int64) // 3]]


def run_marching_cubes(tsdf, colors=None, isovalue=0.0, truncation=3.0, thresh=10.0):
    """-> (vertices FloatTensor [V,3] (x,y,z), vertex colours ByteTensor [V,3], faces IntTensor [F,3]) on the CPU,
    like marching_cubes_cpp.run_marching_cubes (marching_cubes.cpp:480-517).  colors: None (grey 220) or uint8
    [n0,n1,n2,3] per voxel."""
    if colors is None:
        verts, faces = merge_triangles(triangle_soup(tsdf, isovalue, truncation, thresh).cpu())
        v = torch.from_numpy(verts)
        return v, torch.full((v.shape[0], 3), 220, dtype=torch.uint8), torch.from_numpy(faces)
    if tuple(colors.shape) != tuple(tsdf.shape) + (3,):
        raise ValueError('colors must be uint8 [n0,n1,n2,3] for a TSDF of shape [n0,n1,n2]')
    tris, cells = triangle_soup(tsdf, isovalue, truncation, thresh, return_cells=True)
    verts, faces, src = merge_triangles(tris.cpu(), return_source=True)
    vc = vertex_colors(colors, cells.cpu().numpy(), src)
    return torch.from_numpy(verts), torch.from_numpy(np.ascontiguousarray(vc)), torch.from_numpy(faces)


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


def save_to_ply(output_file, verts, colors, faces):
    """marching_cubes.cpp:519-560 (save_to_ply), byte for byte: the header lines it streams, then numV packed 15-byte
    vertex records (3 x float32 + 3 x uint8), then per face the byte 3 and 3 x int32."""
    verts = np.ascontiguousarray(verts, dtype='<f4').reshape(-1, 3)
    colors = np.ascontiguousarray(colors, dtype=np.uint8).reshape(-1, 3)
    faces = np.ascontiguousarray(faces, dtype='<i4').reshape(-1, 3)
    if colors.shape[0] != verts.shape[0]:
        raise ValueError('one colour per vertex')
    vrec = np.empty(verts.shape[0], dtype=np.dtype({'names': ['p', 'c'], 'formats': [('<f4', 3), ('u1', 3)],
                                                     'offsets': [0, 12], 'itemsize': 15}))
    vrec['p'], vrec['c'] = verts, colors
    frec = np.empty(faces.shape[0], dtype=np.dtype({'names': ['n', 'i'], 'formats': ['u1', ('<i4', 3)],
                                                     'offsets': [0, 1], 'itemsize': 13}))
    frec['n'], frec['i'] = 3, faces
    with open(output_file, 'wb') as f:
        f.write(('ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\n'
                 'property float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nelement face %d\n'
                 'property list uchar int vertex_indices\nend_header\n' % (verts.shape[0], faces.shape[0])).encode('ascii'))
        f.write(vrec.tobytes())
        f.write(frec.tobytes())


def save_mesh(verts, colors, faces, output_file):
    """.ply -> the reference's binary writer (marching_cubes.py:31-32 sends `.ply` to export_marching_cubes ->
    save_to_ply); any other extension -> the .obj text form with vertex colours of marching_cubes.py:9-18 (its plyfile
    branch is unreachable for `.ply` names and drops the vertex colours, so it is not mirrored)."""
    verts, colors, faces = np.asarray(verts), np.asarray(colors), np.asarray(faces)
    if os.path.splitext(output_file)[1] == '.ply':
        save_to_ply(output_file, verts, colors, faces)
        return
    with open(output_file, 'w') as f:
        for v, c in zip(verts, colors):
            f.write('v %f %f %f %d %d %d\n' % (v[0], v[1], v[2], c[0], c[1], c[2]))
        f.write('g foo\n')
        for t in faces:
            f.write('f %d %d %d\n' % (t[0] + 1, t[1] + 1, t[2] + 1))
        f.write('g\n')


def marching_cubes(tsdf, colors, isovalue, truncation, thresh, output_filename):
    """Signature of the reference's marching_cubes.marching_cubes (marching_cubes.py:28-35)."""
    if colors is not None and not isinstance(colors, torch.Tensor):
        colors = torch.as_tensor(colors)
    v, c, f = run_marching_cubes(tsdf if tsdf.is_cuda else tsdf.cuda(), colors, isovalue, truncation, thresh)
    save_mesh(v.numpy(), c.numpy(), f.numpy(), output_filename)
