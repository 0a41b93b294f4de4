"""ctypes wrapper of oracle/mcubes.cpp (ORACLE -- TEST INFRASTRUCTURE ONLY): the CPU restatement of the reference's
run_marching_cubes.  numpy in / numpy out."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libmcubes.so')
_GOLDEN = os.path.join(os.path.dirname(_HERE), 'tests', 'golden', 'mc_ref.npz')
_lib = None
_table = None


def build(force=False):
    src = os.path.join(_HERE, 'mcubes.cpp')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-shared', '-fPIC', '-o', _SO, src])
    return _SO


def tri_table():
    """[256,16] int8, recovered from the reference by probing (tests/golden/make_mc_golden.py)."""
    global _table
    if _table is None:
        _table = np.ascontiguousarray(np.load(_GOLDEN)['tri_table'].astype(np.int8))
    return _table


def marching_cubes(tsdf, isovalue=0.0, truncation=3.0, thresh=10.0):
    """-> (vertices float32 [V,3] in (x,y,z), faces int32 [F,3]); arguments as data_util.py:270 passes them."""
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    t = np.ascontiguousarray(tsdf, dtype=np.float32)
    tab = tri_table()
    _lib.mc_run(t.ctypes.data_as(C.c_void_p), C.c_int(t.shape[0]), C.c_int(t.shape[1]), C.c_int(t.shape[2]),
                C.c_float(isovalue), C.c_float(truncation), C.c_float(thresh), tab.ctypes.data_as(C.c_void_p))
    nv, nf = C.c_int(0), C.c_int(0)
    _lib.mc_counts(C.byref(nv), C.byref(nf))
    verts = np.empty((nv.value, 3), dtype=np.float32)
    faces = np.empty((nf.value, 3), dtype=np.int32)
    _lib.mc_copy(verts.ctypes.data_as(C.c_void_p), faces.ctypes.data_as(C.c_void_p))
    return verts, faces


def triangle_soup(tsdf, isovalue=0.0, truncation=3.0, thresh=10.0):
    """The triangles before the vertex merge, float32 [T,3,3], in the reference's emission order."""
    marching_cubes(tsdf, isovalue, truncation, thresh)
    n = _lib.mc_soup_triangles()
    tris = np.empty((n, 3, 3), dtype=np.float32)
    _lib.mc_soup_copy(tris.ctypes.data_as(C.c_void_p))
    return tris
