"""ctypes wrapper of oracle/o3.c (ORACLE O3 -- TEST INFRASTRUCTURE ONLY).  CPU torch tensors in/out."""
import ctypes as C
import os
import subprocess
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libo3.so')


def build(force=False):
    src = os.path.join(_HERE, 'o3.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-o', _SO, src, '-lm'])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def conv(x, nbr, weight, n_out, child_mode=False, residual=None, scale=None, shift=None, relu=False):
    K, cin, cout = weight.shape
    x = x.contiguous().float()
    nbr = nbr.contiguous().int()
    weight = weight.contiguous().float()
    out = torch.empty((n_out, cout), dtype=torch.float32)
    if residual is not None:
        residual = residual.contiguous().float()
    lib().o3_conv(_p(x), C.c_int(x.shape[1]), _p(nbr), C.c_int64(nbr.shape[1]), C.c_int(K),
                  C.c_int(1 if child_mode else 0), _p(weight), C.c_int(cin), C.c_int(cout), C.c_int64(n_out),
                  _p(residual), C.c_int(cout), _p(out), C.c_int(cout),
                  _p(scale.contiguous().float() if scale is not None else None),
                  _p(shift.contiguous().float() if shift is not None else None), C.c_int(1 if relu else 0))
    return out


def deconv(x, parent, weight):
    K, cin, cout = weight.shape
    x = x.contiguous().float()
    parent = parent.contiguous().int()
    weight = weight.contiguous().float()
    out = torch.empty((parent.shape[0], cout), dtype=torch.float32)
    lib().o3_deconv(_p(x), C.c_int(x.shape[1]), _p(parent), _p(weight), C.c_int(cin), C.c_int(cout),
                    C.c_int64(parent.shape[0]), _p(out), C.c_int(cout))
    return out


def affine_relu(x, scale, shift, relu=True):
    x = x.contiguous().float()
    y = torch.empty_like(x)
    lib().o3_affine_relu(_p(x), C.c_int(x.shape[1]), _p(y), C.c_int(x.shape[1]), C.c_int64(x.shape[0]),
                         C.c_int(x.shape[1]), _p(scale.contiguous().float() if scale is not None else None),
                         _p(shift.contiguous().float() if shift is not None else None), C.c_int(1 if relu else 0))
    return y


def linear(x, w, b):
    x = x.contiguous().float()
    w = w.contiguous().float()
    cout, cin = w.shape
    y = torch.empty((x.shape[0], cout), dtype=torch.float32)
    lib().o3_linear(_p(x), C.c_int(x.shape[1]), _p(w), _p(b.contiguous().float() if b is not None else None),
                    _p(y), C.c_int(cout), C.c_int64(x.shape[0]), C.c_int(cin), C.c_int(cout))
    return y


def sigmoid_gt_half(x):
    x = x.contiguous().float().view(-1)
    out = torch.empty(x.shape[0], dtype=torch.uint8)
    lib().o3_sigmoid_gt_half(_p(x), C.c_int64(x.shape[0]), _p(out))
    return out.bool()


def dense_conv(x, weight, ksize, stride, pad, scale=None, shift=None, relu=False, transposed=False):
    x = x.contiguous().float()
    weight = weight.contiguous().float()
    nb, cin, d0, d1, d2 = x.shape
    cout = weight.shape[1] if transposed else weight.shape[0]
    if transposed:
        o = [(d - 1) * stride - 2 * pad + ksize for d in (d0, d1, d2)]
    else:
        o = [(d + 2 * pad - ksize) // stride + 1 for d in (d0, d1, d2)]
    out = torch.empty((nb, cout, o[0], o[1], o[2]), dtype=torch.float32)
    lib().o3_dense_conv(C.c_int(1 if transposed else 0), _p(x), C.c_int(cin), C.c_int(nb), C.c_int(d0), C.c_int(d1),
                        C.c_int(d2), _p(weight), C.c_int(cout), C.c_int(ksize), C.c_int(stride), C.c_int(pad),
                        _p(scale.contiguous().float() if scale is not None else None),
                        _p(shift.contiguous().float() if shift is not None else None), C.c_int(1 if relu else 0),
                        _p(out), C.c_int(o[0]), C.c_int(o[1]), C.c_int(o[2]))
    return out
