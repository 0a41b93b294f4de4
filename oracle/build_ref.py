"""Builds the parts of the REFERENCE that compile from their own few sources, from where they lie under /root/reference,
into oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).  TEST INFRASTRUCTURE ONLY -- nothing in the
product path loads it.

  marching_cubes_cpp.so   /root/reference/torch/marching_cubes/marching_cubes.cpp (+ tables.h, sparsegrid3.h): the torch
                          extension SG-NN's data_util.py:270-284 calls after the forward pass; plain g++
                          against the torch headers, no build system.

SparseConvNet itself (the hot path's arithmetic) is NOT in /root/reference and cannot be built.
No reference source is copied into this repository: the compiler reads the files in place."""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
REF_MC = '/root/reference/torch/marching_cubes'


def mc_path():
    return os.path.join(OUT, 'marching_cubes_cpp.so')


def build_marching_cubes(force=False):
    """Returns the path of the built extension, or None when the reference sources are not on this machine."""
    src = os.path.join(REF_MC, 'marching_cubes.cpp')
    out = mc_path()
    if not os.path.exists(src):
        return out if os.path.exists(out) else None
    if os.path.exists(out) and not force and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    libdirs = ce.library_paths()
    cmd = ['g++', '-O2', '-shared', '-fPIC', '-std=c++17', '-w', '-DTORCH_EXTENSION_NAME=marching_cubes_cpp',
           '-D_GLIBCXX_USE_CXX11_ABI=%d' % int(torch._C._GLIBCXX_USE_CXX11_ABI), '-I' + sysconfig.get_paths()['include']]
    cmd += ['-I' + i for i in ce.include_paths()] + ['-I' + REF_MC, src, '-o', out]
    cmd += ['-L' + l for l in libdirs] + ['-ltorch', '-ltorch_cpu', '-lc10', '-ltorch_python', '-Wl,-rpath,' + libdirs[0]]
    subprocess.check_call(cmd)
    return out


def load_marching_cubes():
    """Imports oracle/_ref/marching_cubes_cpp.so (built earlier) or returns None."""
    p = mc_path()
    if not os.path.exists(p):
        return None
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location('marching_cubes_cpp', p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    print(build_marching_cubes(force='--force' in sys.argv))
