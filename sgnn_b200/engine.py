"""Thin functional layer over the C ABI: torch tensors in (device memory + current stream are the
only things taken from PyTorch), one libsgnn_b200 call per function.  No CPU / PyTorch fallback:
tensors must live on a CUDA device.
"""
import ctypes as C
import torch

from . import _lib
from ._lib import lib, check, SgnnGrid, SgnnEpilogue, SgnnConvArgs

__all__ = ['Grid', 'build_grid', 'coarsen', 'rulebook_submanifold', 'rulebook_strided',
           'conv', 'tile_plan', 'deconv', 'unpool', 'affine_relu', 'add_rows', 'copy_cols', 'linear',
           'sparse_to_dense', 'dense_to_sparse', 'heads_compact', 'children_coords',
           'concat_skip', 'coords_to_i64', 'grid_lookup', 'fold_bn', 'dense_conv']


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _need_cuda(*ts):
    """Every tensor on a CUDA device, and on the CURRENT one: the C ABI launches on the current device / its current stream,
    so a tensor of another device would hand foreign pointers to kernels of this one.  (GenModel.forward switches the device
    itself; direct users of this layer wrap their calls in `with torch.cuda.device(t.device)`.)"""
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError('sgnn_b200: tensors must be CUDA tensors (no CPU fallback in the product path)')
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError('sgnn_b200: tensor on cuda:%d but the current device is cuda:%d -- wrap the call in '
                               '`with torch.cuda.device(tensor.device)`' % (t.device.index, cur))


# Optional profiler hook (bench.py): an object with .conv(tag, x, nbr, weight, n_out, child_mode) -> ctx or None,
# where ctx has .done().  None in normal operation: zero overhead.
PROFILER = None


def _scratch(nbytes, device):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


class Grid(object):
    """Device-resident active-site set (SgnnGrid + row coordinates)."""

    def __init__(self, nb, dims, device, with_perm):
        self.nb = int(nb)
        self.d = [int(v) for v in dims]
        self.wx = (self.d[2] + 63) // 64
        self.n_words = self.nb * self.d[0] * self.d[1] * self.wx
        self.device = device
        self.mask = torch.empty(max(self.n_words, 1), dtype=torch.int64, device=device)
        self.prefix = torch.empty(self.n_words + 1, dtype=torch.int32, device=device)
        self.row_of_rank = None
        self.with_perm = with_perm
        self.coords = None   # int32 [n,4] in row order
        self.n = 0
        self.c = None

    def _finish(self):
        g = SgnnGrid()
        g.nb, g.d0, g.d1, g.d2 = self.nb, self.d[0], self.d[1], self.d[2]
        g.wx, g.reserved, g.n_words = self.wx, 0, self.n_words
        g.mask = self.mask.data_ptr()
        g.prefix = self.prefix.data_ptr()
        g.row_of_rank = self.row_of_rank.data_ptr() if self.row_of_rank is not None else None
        self.c = g

    def ref(self):
        return C.byref(self.c)


def build_grid(coords, nb, dims, status=None):
    """a1: grid from caller-ordered coords (int64 or int32 CUDA tensor [n,4])."""
    _need_cuda(coords)
    assert coords.dim() == 2 and coords.shape[1] == 4
    coords = coords.contiguous()
    n = coords.shape[0]
    g = Grid(nb, dims, coords.device, True)
    g.n = n
    g.coords = torch.empty((n, 4), dtype=torch.int32, device=coords.device)
    g.row_of_rank = torch.empty(max(n, 1), dtype=torch.int32, device=coords.device)
    g._finish()
    is64 = 1 if coords.dtype == torch.int64 else 0
    if not is64:
        assert coords.dtype == torch.int32
    sb = lib.sgnn_scan_scratch_bytes(g.n_words)
    scr = _scratch(sb, coords.device)
    check(lib.sgnn_grid_build(g.ref(), _ptr(coords), is64, n, _ptr(g.coords), _ptr(status),
                              _ptr(scr), sb, _stream()), 'sgnn_grid_build')
    return g


def coarsen(fine, dims_cap=None):
    """a4: stride-2 coarse site set (raster row order).  One host read of the row count.
    Extent = scn's Convolution output size (S-2)//2+1 per axis (fine cells beyond it are dropped, App. A.5)."""
    d = [((v - 2) // 2 + 1) if v >= 2 else 0 for v in fine.d]
    if dims_cap is not None:
        d = [min(a, int(b)) for a, b in zip(d, dims_cap)]
    g = Grid(fine.nb, d, fine.device, False)
    g._finish()
    sb = lib.sgnn_scan_scratch_bytes(g.n_words)
    scr = _scratch(sb, fine.device)
    check(lib.sgnn_grid_coarsen(fine.ref(), g.ref(), _ptr(scr), sb, _stream()), 'sgnn_grid_coarsen')
    g.n = int(g.prefix[g.n_words].item())
    g.coords = torch.empty((g.n, 4), dtype=torch.int32, device=fine.device)
    if g.n:
        check(lib.sgnn_grid_enumerate(g.ref(), _ptr(g.coords), _stream()), 'sgnn_grid_enumerate')
    return g


def grid_lookup(grid, coords, shift=0):
    _need_cuda(coords)
    coords = coords.contiguous()
    rows = torch.empty(coords.shape[0], dtype=torch.int32, device=coords.device)
    check(lib.sgnn_grid_lookup(grid.ref(), _ptr(coords), coords.shape[0], shift, _ptr(rows), _stream()),
          'sgnn_grid_lookup')
    return rows


def rulebook_submanifold(grid):
    """a2: neighbour table int32 [27, n]."""
    nbr = torch.empty((27, grid.n), dtype=torch.int32, device=grid.device)
    check(lib.sgnn_rulebook_submanifold(grid.ref(), _ptr(grid.coords), grid.n, _ptr(nbr), _stream()),
          'sgnn_rulebook_submanifold')
    return nbr


def rulebook_submanifold_compact(grid):
    """a2, compact form: (slots int32 [27, n] -- only slots[s][i], s < cnt[i], are defined: k << 27 | row --, cnt uint8 [n])."""
    slots = torch.zeros((27, grid.n), dtype=torch.int32, device=grid.device)
    cnt = torch.empty(grid.n, dtype=torch.uint8, device=grid.device)
    check(lib.sgnn_rulebook_submanifold_compact(grid.ref(), _ptr(grid.coords), grid.n, _ptr(slots), _ptr(cnt), _stream()),
          'sgnn_rulebook_submanifold_compact')
    return slots, cnt


def rulebook_strided(fine, coarse):
    """a4: parent int32 [n_fine] (row*8+k) and children int32 [8, n_coarse]."""
    parent = torch.empty(fine.n, dtype=torch.int32, device=fine.device)
    children = torch.empty((8, coarse.n), dtype=torch.int32, device=fine.device)
    check(lib.sgnn_rulebook_strided(coarse.ref(), _ptr(fine.coords), fine.n, _ptr(parent), _ptr(children),
                                    coarse.n, _stream()), 'sgnn_rulebook_strided')
    return parent, children


def coarse_build(fine, coarse):
    """sgnn_grid_coarse_build: coordinates of a raster-ordered coarse set + the strided rulebook against `fine`, one kernel.
    -> (coords int32 [n_coarse, 4], parent int32 [n_fine], children int32 [8, n_coarse])"""
    coords = torch.empty((coarse.n, 4), dtype=torch.int32, device=fine.device)
    parent = torch.empty(fine.n, dtype=torch.int32, device=fine.device)
    children = torch.empty((8, coarse.n), dtype=torch.int32, device=fine.device)
    check(lib.sgnn_grid_coarse_build(fine.ref(), coarse.ref(), fine.n, coarse.n, _ptr(coords), _ptr(parent), _ptr(children),
                                     _stream()), 'sgnn_grid_coarse_build')
    return coords, parent, children


def _epilogue(out, scale=None, shift=None, relu=False):
    e = SgnnEpilogue()
    if out is None:
        e.out, e.ld, e.relu, e.scale, e.shift = None, 0, 0, None, None
        return e
    _need_cuda(out, scale, shift)
    assert out.dtype in (torch.float32, torch.bfloat16) and out.stride(-1) == 1
    e.out = out.data_ptr()
    e.ld = out.stride(0) if out.dim() == 2 else 1
    e.relu = 1 if relu else 0
    e.scale = scale.data_ptr() if scale is not None else None
    e.shift = shift.data_ptr() if shift is not None else None
    return e


def conv(x, nbr, weight, n_out, out_a, child_mode=False, residual=None, scale_a=None, shift_a=None,
         relu_a=False, out_b=None, scale_b=None, shift_b=None, relu_b=False, tc32=False, plan=None, compact=None, flags=0):
    """a3/a4/a9: out[j] = sum_k x[nbr[k][j]] @ W[k]  (+residual, affine, relu; two output slots).
    x / out_* may be column views of wider row-major buffers (stride(1) == 1).
    tc32=True: the tensor-core path for fp32 features (sgnn_conv_forward_tc32: Cout = 16, Cin <= 48; fp32 accuracy,
    not the fixed fmaf order); raises for unsupported shapes.
    compact=(slots, cnt) of rulebook_submanifold_compact: sgnn_conv_forward_compact (nbr may be None)."""
    _need_cuda(x, nbr, weight, out_a, residual, out_b)
    K, cin, cout = weight.shape
    assert weight.is_contiguous() and weight.dtype in (torch.float32, torch.bfloat16)
    assert x.dtype == weight.dtype == out_a.dtype and x.stride(1) == 1 and x.shape[1] == cin
    if compact is not None:
        _need_cuda(*compact)
        nbr = compact[0]
    assert nbr.dtype == torch.int32 and nbr.shape[0] == K and nbr.stride(1) == 1
    a = SgnnConvArgs()
    a.in_ = x.data_ptr()
    a.ld_in = x.stride(0)
    a.dtype = _lib.SGNN_BF16 if x.dtype == torch.bfloat16 else _lib.SGNN_F32
    a.nbr = nbr.data_ptr()
    a.nbr_stride = nbr.stride(0)
    a.K = K
    a.child_mode = 1 if child_mode else 0
    a.weight = weight.data_ptr()
    a.cin, a.cout = cin, cout
    a.n_out = int(n_out)
    a.n_in = int(x.shape[0])
    if residual is not None:
        assert residual.stride(1) == 1
        a.residual = residual.data_ptr()
        a.ld_res = residual.stride(0)
    else:
        a.residual = None
        a.ld_res = 0
    a.a = _epilogue(out_a, scale_a, shift_a, relu_a)
    a.b = _epilogue(out_b, scale_b, shift_b, relu_b)
    a.flags = int(flags)          # SGNN_CONV_ROWLANE (2) / SGNN_CONV_NO_ROWLANE (4): kernel selection A/B, same bits
    ctx = PROFILER.conv(x, nbr, weight, int(n_out), child_mode, residual is not None,
                        out_b is not None) if PROFILER is not None else None
    if compact is not None:
        check(lib.sgnn_conv_forward_compact(C.byref(a), _ptr(compact[0]), _ptr(compact[1]), _stream()),
              'sgnn_conv_forward_compact')
    elif plan is not None and child_mode:
        wb = 64 * 4608
        ws = _scratch(wb, x.device)
        check(lib.sgnn_conv_forward_tc32_urc(C.byref(a), _ptr(plan), C.c_void_p(ws.data_ptr()), wb, _stream()),
              'sgnn_conv_forward_tc32_urc')
    elif plan is not None:
        wb = lib.sgnn_conv_tc32_workspace_bytes(K, cin, 0)
        ws = _scratch(wb, x.device)
        check(lib.sgnn_conv_forward_tc32_ur(C.byref(a), _ptr(plan), C.c_void_p(ws.data_ptr()), wb, _stream()),
              'sgnn_conv_forward_tc32_ur')
    elif tc32:
        wb = lib.sgnn_conv_tc32_workspace_bytes(K, cin, a.child_mode)
        ws = _scratch(wb, x.device)
        check(lib.sgnn_conv_forward_tc32(C.byref(a), C.c_void_p(ws.data_ptr()), wb, _stream()), 'sgnn_conv_forward_tc32')
    else:
        check(lib.sgnn_conv_forward(C.byref(a), _stream()), 'sgnn_conv_forward')
    if ctx is not None:
        ctx.done()
    return out_a


def tile_plan(nbr, n_rows):
    """Unique-row tile plan of a [27, n] neighbour table (sgnn_tile_plan_build); pass it to conv(..., plan=)."""
    _need_cuda(nbr)
    assert nbr.dtype == torch.int32 and nbr.shape[0] == 27 and nbr.stride(1) == 1
    nb = lib.sgnn_tile_plan_bytes(int(n_rows))
    plan = _scratch(nb, nbr.device)
    assert plan.data_ptr() % 256 == 0
    check(lib.sgnn_tile_plan_build(_ptr(nbr), nbr.stride(0), int(n_rows), _ptr(plan), nb, _stream()), 'sgnn_tile_plan_build')
    return plan


def rulebook_submanifold_plan(grid):
    """a2 + tile plan in one kernel (sgnn_rulebook_submanifold_plan): (nbr int32 [27, n], plan) -- the same table and the same
    plan as rulebook_submanifold followed by tile_plan."""
    nbr = torch.empty((27, grid.n), dtype=torch.int32, device=grid.device)
    nb = lib.sgnn_tile_plan_bytes(int(grid.n))
    plan = torch.empty(nb, dtype=torch.uint8, device=grid.device)
    assert plan.data_ptr() % 256 == 0
    check(lib.sgnn_rulebook_submanifold_plan(grid.ref(), _ptr(grid.coords), grid.n, _ptr(nbr), _ptr(plan), nb, _stream()),
          'sgnn_rulebook_submanifold_plan')
    return nbr, plan


def deconv(x, parent, weight, out, scale=None, shift=None, relu=False):
    _need_cuda(x, parent, weight, out)
    K, cin, cout = weight.shape
    assert K == 8 and x.stride(1) == 1
    e = _epilogue(out, scale, shift, relu)
    dt = _lib.SGNN_BF16 if x.dtype == torch.bfloat16 else _lib.SGNN_F32
    assert x.dtype == weight.dtype == out.dtype
    check(lib.sgnn_deconv_forward(_ptr(x), x.stride(0), dt, _ptr(parent), _ptr(weight), cin, cout,
                                  parent.shape[0], C.byref(e), _stream()), 'sgnn_deconv_forward')
    return out


def unpool(x, parent, out, scale=None, shift=None, relu=False):
    _need_cuda(x, parent, out)
    assert x.stride(1) == 1
    e = _epilogue(out, scale, shift, relu)
    check(lib.sgnn_unpool(_ptr(x), x.stride(0), _ptr(parent), x.shape[1], parent.shape[0], C.byref(e),
                          _stream()), 'sgnn_unpool')
    return out


def affine_relu(x, out, scale, shift, relu=True):
    _need_cuda(x, out, scale, shift)
    assert x.stride(1) == 1 and out.stride(1) == 1
    check(lib.sgnn_affine_relu(_ptr(x), x.stride(0), _ptr(out), out.stride(0), x.shape[0], x.shape[1],
                               _ptr(scale), _ptr(shift), 1 if relu else 0, _stream()), 'sgnn_affine_relu')
    return out


def add_rows(a, b, out):
    _need_cuda(a, b, out)
    check(lib.sgnn_add_rows(_ptr(a), a.stride(0), _ptr(b), b.stride(0), _ptr(out), out.stride(0), a.shape[0],
                            a.shape[1], _stream()), 'sgnn_add_rows')
    return out


def copy_cols(src, dst):
    """dst[:, :c] = src (dst is a column view of a wider buffer)."""
    _need_cuda(src, dst)
    check(lib.sgnn_copy_cols(_ptr(src), src.stride(0), _ptr(dst), dst.stride(0), src.shape[0], src.shape[1],
                             _stream()), 'sgnn_copy_cols')
    return dst


def linear(x, weight, bias, out):
    _need_cuda(x, weight, bias, out)
    cout, cin = weight.shape
    assert x.stride(1) == 1 and weight.is_contiguous() and x.shape[1] == cin
    check(lib.sgnn_linear(_ptr(x), x.stride(0), _ptr(weight), _ptr(bias), _ptr(out), out.stride(0), x.shape[0],
                          cin, cout, _stream()), 'sgnn_linear')
    return out


def sparse_to_dense(feats, coords, nb, dims):
    _need_cuda(feats, coords)
    coords = coords.contiguous()
    c = feats.shape[1]
    dense = torch.empty((nb, c, dims[0], dims[1], dims[2]), dtype=torch.float32, device=feats.device)
    check(lib.sgnn_sparse_to_dense(_ptr(feats), feats.stride(0) if feats.shape[0] else c, _ptr(coords),
                                   feats.shape[0], c, _ptr(dense), nb, dims[0], dims[1], dims[2], _stream()),
          'sgnn_sparse_to_dense')
    return dense


def dense_to_sparse(dense_feats, dense_out, ld_feats=None, want_cand=True):
    """a8.  Returns locs int32 [M,4], feats [M, ld_feats] (first c+2 columns filled), cand [cells,2], M.
    One host read of the kept count."""
    _need_cuda(dense_feats, dense_out)
    nb, c, d0, d1, d2 = dense_feats.shape
    assert dense_feats.is_contiguous() and dense_out.is_contiguous() and dense_out.shape[1] == 2
    total = nb * d0 * d1 * d2
    dev = dense_feats.device
    ld = ld_feats or (c + 2)
    locs = torch.empty((total, 4), dtype=torch.int32, device=dev)
    feats = torch.zeros((total, ld), dtype=torch.float32, device=dev)
    cand = torch.empty((total, 2), dtype=torch.float32, device=dev) if want_cand else None
    count = torch.empty(1, dtype=torch.int32, device=dev)
    sb = lib.sgnn_compact_scratch_bytes(total)
    scr = _scratch(sb, dev)
    check(lib.sgnn_dense_to_sparse(_ptr(dense_feats), _ptr(dense_out), nb, c, d0, d1, d2, _ptr(locs), _ptr(feats),
                                   ld, _ptr(cand), _ptr(count), _ptr(scr), sb, _stream()), 'sgnn_dense_to_sparse')
    m = int(count.item())
    return locs[:m], feats[:m], cand, m


def heads_compact(x, w_occ, b_occ, w_sdf, b_sdf, parent_coords, ld_feats=None):
    """a9.  x [8*n_parent, c] post-BNReLU candidate features.  Returns locs [M,4] int32, feats [M, ld]
    (= [x, occ, sdf] in the first c+2 columns), cand [8*n_parent, 2], M (one host read)."""
    _need_cuda(x, w_occ, b_occ, w_sdf, b_sdf, parent_coords)
    parent_coords = parent_coords.contiguous()
    n_parent = parent_coords.shape[0]
    n_cand = 8 * n_parent
    c = x.shape[1]
    assert x.shape[0] == n_cand and x.stride(1) == 1
    dev = x.device
    ld = ld_feats or (c + 2)
    cand = torch.empty((n_cand, 2), dtype=torch.float32, device=dev)
    locs = torch.empty((n_cand, 4), dtype=torch.int32, device=dev)
    feats = torch.zeros((n_cand, ld), dtype=torch.float32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    sb = lib.sgnn_compact_scratch_bytes(n_cand)
    scr = _scratch(sb, dev)
    check(lib.sgnn_heads_compact(_ptr(x), x.stride(0), c, _ptr(w_occ), _ptr(b_occ), _ptr(w_sdf), _ptr(b_sdf),
                                 _ptr(parent_coords), n_parent, _ptr(cand), _ptr(locs), _ptr(feats), ld,
                                 _ptr(count), _ptr(scr), sb, _stream()), 'sgnn_heads_compact')
    m = int(count.item())
    return locs[:m], feats[:m], cand, m


def children_coords(parent_coords):
    _need_cuda(parent_coords)
    parent_coords = parent_coords.contiguous()
    n = parent_coords.shape[0]
    out = torch.empty((8 * n, 4), dtype=torch.int32, device=parent_coords.device)
    check(lib.sgnn_children_coords(_ptr(parent_coords), n, _ptr(out), _stream()), 'sgnn_children_coords')
    return out


def concat_skip(grid, src, coords, dst, col0):
    """a10: dst[:, col0:col0+c] = src[row of coords in grid] or 0."""
    _need_cuda(src, coords, dst)
    coords = coords.contiguous()
    assert dst.stride(1) == 1 and (src.shape[0] == 0 or src.stride(1) == 1)
    check(lib.sgnn_concat_skip(grid.ref(), _ptr(src), src.stride(0) if src.shape[0] else src.shape[1],
                               src.shape[1], _ptr(coords), coords.shape[0], _ptr(dst), dst.stride(0), col0,
                               _stream()), 'sgnn_concat_skip')
    return dst


def coords_to_i64(coords):
    _need_cuda(coords)
    out = torch.empty(coords.shape, dtype=torch.int64, device=coords.device)
    check(lib.sgnn_coords_to_i64(_ptr(coords), coords.shape[0], _ptr(out), _stream()), 'sgnn_coords_to_i64')
    return out


def dense_conv(x0, x1, weight, cout, ksize, stride, pad, scale=None, shift=None, relu=False, transposed=False):
    """a12: nn.Conv3d / nn.ConvTranspose3d (+ folded BatchNorm3d + ReLU) on NCDHW volumes; x1 (optional) is
    concatenated behind x0 along channels without materialising the cat."""
    _need_cuda(x0, x1, weight, scale, shift)
    assert x0.is_contiguous() and x0.dtype == torch.float32 and weight.is_contiguous()
    nb, c0, d0, d1, d2 = x0.shape
    c1 = 0
    if x1 is not None:
        assert x1.is_contiguous() and x1.shape[0] == nb and tuple(x1.shape[2:]) == (d0, d1, d2)
        c1 = x1.shape[1]
    if transposed:
        o = [(d - 1) * stride - 2 * pad + ksize for d in (d0, d1, d2)]
        fn = lib.sgnn_dense_convT3d
    else:
        o = [(d + 2 * pad - ksize) // stride + 1 for d in (d0, d1, d2)]
        fn = lib.sgnn_dense_conv3d
    out = torch.empty((nb, cout, o[0], o[1], o[2]), dtype=torch.float32, device=x0.device)
    check(fn(_ptr(x0), c0, _ptr(x1), c1, nb, d0, d1, d2, _ptr(weight), cout, ksize, stride, pad, _ptr(scale),
             _ptr(shift), 1 if relu else 0, _ptr(out), _stream()), 'sgnn_dense_conv')
    return out


def fold_bn(bn):
    """Eval-mode BatchNormReLU folded to (scale, shift): y = max(fma(x, scale, shift), 0).
    scale = gamma * (running_var + eps)^-1/2, shift = beta - running_mean * scale  (SURVEY App. A.8).
    Folded in fp32 ON THE HOST (one well-defined rounding, shared with the oracle tests) and cached per
    module until a parameter or buffer changes."""
    ts = (bn.weight, bn.bias, bn.running_mean, bn.running_var)
    key = tuple((t._version, t.data_ptr()) for t in ts)
    hit = getattr(bn, '_sgnn_fold', None)        # cached ON the module: dies with it (no global table of strong references)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    w, b, rm, rv = [t.detach().cpu().float() for t in ts]
    inv = (rv + bn.eps).pow(-0.5)
    scale = (inv * w).contiguous()
    shift = (b - rm * scale).contiguous()
    dev = bn.running_var.device
    scale, shift = scale.to(dev), shift.to(dev)
    object.__setattr__(bn, '_sgnn_fold', (key, scale, shift))
    return scale, shift
