"""B200-backed `sparseconvnet` operator surface -- the drop-in boundary of SURVEY §8(b).

Same class names, constructor arguments, parameter names/shapes and error behaviour as the symbols
/root/reference/torch/model.py takes from `import sparseconvnet as scn` (model.py:7, call sites
:31-47,:178-188,:253-257,:296,:380), plus Deconvolution / UnPooling / JoinTable from the north_star
operator list.  Every forward runs hand-written sm_100a kernels through the C ABI
(include/sgnn_b200.h); inputs must be CUDA tensors -- there is no CPU path here (the CPU
restatement lives in oracle/ and is test infrastructure only).

Forward inference only (north_star): BatchNormReLU is evaluated with running statistics and raises
in training mode; no autograd is recorded through the sparse kernels.
"""
import numpy as np
import torch
import torch.nn as nn

from . import engine as E

__all__ = [
    'InputLayer', 'OutputLayer', 'SubmanifoldConvolution', 'Convolution', 'Deconvolution',
    'UnPooling', 'BatchNormReLU', 'BatchNormalization', 'Sequential', 'ConcatTable', 'AddTable',
    'JoinTable', 'Identity', 'SparseToDense', 'FullyConvolutionalNet', 'NetworkInNetwork',
    'SparseConvNetTensor', 'Metadata',
]


def _long3(dimension, x):
    if dimension != 3:
        raise NotImplementedError('sgnn_b200 implements dimension 3 (SG-NN); got %r' % (dimension,))
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().long().view(-1).clone()
    if isinstance(x, (list, tuple, np.ndarray)):
        if len(x) != 3:
            raise ValueError('spatial_size must have 3 entries')
        return torch.LongTensor([int(v) for v in x])
    return torch.LongTensor([int(x)] * 3)


def _key(ssz):
    return tuple(int(v) for v in ssz)


class Metadata(object):
    """Device-side site sets and rulebooks of one tensor chain (role of scn's Metadata<3>)."""

    def __init__(self, dimension=3):
        self.dimension = dimension
        self.grids = {}      # spatial size -> engine.Grid
        self.rb_sub = {}     # spatial size -> nbr [27,n]
        self.rb_str = {}     # (fine size, coarse size) -> (parent, children)
        self.batch_size = 0

    def grid(self, spatial_size):
        try:
            return self.grids[_key(spatial_size)]
        except KeyError:
            raise RuntimeError('no active-site set at spatial size %s in this metadata' % (_key(spatial_size),))

    def getSpatialLocations(self, spatial_size):
        """CPU LongTensor [N,4] (z,y,x,batch) in row order (model.py:380), as scn returns it: the reference's concat_skip
        (model.py:344-353) indexes CPU tensors with it.  The device copy stays internal (self.grid(...).coords)."""
        return E.coords_to_i64(self.grid(spatial_size).coords).cpu()

    def nActive(self, spatial_size):
        return self.grid(spatial_size).n

    def getSubmanifoldRuleBook(self, spatial_size, filter_size=3):
        if int(filter_size) != 3:
            raise NotImplementedError('submanifold filter size %r (SG-NN uses 3)' % (filter_size,))
        k = _key(spatial_size)
        if k not in self.rb_sub:
            self.rb_sub[k] = E.rulebook_submanifold(self.grid(spatial_size))
        return self.rb_sub[k]

    def getRuleBook(self, fine_size, coarse_size, filter_size=2, filter_stride=2):
        if int(filter_size) != 2 or int(filter_stride) != 2:
            raise NotImplementedError('strided filter %r/%r (SG-NN uses 2/2)' % (filter_size, filter_stride))
        k = (_key(fine_size), _key(coarse_size))
        if k not in self.rb_str:
            fine = self.grid(fine_size)
            ck = _key(coarse_size)
            if ck not in self.grids:
                self.grids[ck] = E.coarsen(fine, dims_cap=ck)
            self.rb_str[k] = E.rulebook_strided(fine, self.grids[ck])
        return self.rb_str[k]


class SparseConvNetTensor(object):
    def __init__(self, features=None, metadata=None, spatial_size=None):
        self.features = features
        self.metadata = metadata
        self.spatial_size = spatial_size

    def get_spatial_locations(self, spatial_size=None):
        return self.metadata.getSpatialLocations(self.spatial_size if spatial_size is None else spatial_size)

    def __repr__(self):
        return 'SparseConvNetTensor<<features=%s, spatial_size=%s>>' % (
            tuple(self.features.shape), self.spatial_size.tolist())


# ---------------------------------------------------------------- containers
class Sequential(nn.Sequential):
    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self

    def forward(self, input):
        for m in self._modules.values():
            input = m(input)
        return input


class ConcatTable(nn.Sequential):
    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self

    def forward(self, input):
        return [m(input) for m in self._modules.values()]


class Identity(nn.Module):
    def forward(self, input):
        return input


class AddTable(nn.Module):
    def forward(self, input):
        out = SparseConvNetTensor(None, input[0].metadata, input[0].spatial_size)
        f = input[0].features
        for i in input[1:]:
            f = E.add_rows(f, i.features, torch.empty_like(f))
        out.features = f
        return out


class JoinTable(nn.Module):
    def forward(self, input):
        n = input[0].features.shape[0]
        widths = [i.features.shape[1] for i in input]
        f = torch.empty((n, sum(widths)), dtype=torch.float32, device=input[0].features.device)
        c0 = 0
        for i, w in zip(input, widths):
            E.copy_cols(i.features, f[:, c0:c0 + w])
            c0 += w
        return SparseConvNetTensor(f, input[0].metadata, input[0].spatial_size)


# --------------------------------------------------------------- IO layers
class InputLayer(nn.Module):
    """scn.InputLayer(dimension, spatial_size, mode) -- model.py:31,178,185,253.  mode 0 only (the mode the
    reference uses): coordinates unique, rows keep caller order.  `spatial_size` stays an assignable
    LongTensor(3) (model.py:364-369).  The bitmask grid covers the declared size (or, when update_sizes' quirk
    declares an absurd bound, the data extent rounded up to 64 cells)."""

    def __init__(self, dimension, spatial_size, mode=3):
        nn.Module.__init__(self)
        self.dimension = dimension
        self.spatial_size = _long3(dimension, spatial_size)
        self.mode = mode

    def forward(self, input, device=None):
        coords, feats = input[0], input[1]
        if self.mode != 0:
            raise NotImplementedError('InputLayer mode %r: sgnn_b200 implements mode 0 (model.py:31)' % (self.mode,))
        if not feats.is_cuda:
            raise RuntimeError('sgnn_b200.scn.InputLayer: features must be a CUDA tensor')
        dev = feats.device
        coords = coords.detach()
        if coords.dim() != 2 or coords.shape[1] not in (3, 4):
            raise ValueError('coords must be [N,3] or [N,4]')
        if coords.shape[1] == 3:
            coords = torch.cat([coords, coords.new_zeros((coords.shape[0], 1))], 1)
        if coords.dtype not in (torch.int64, torch.int32):
            coords = coords.long()
        n = coords.shape[0]
        if feats.shape[0] != n:
            raise ValueError('coords/features row mismatch: %d vs %d' % (n, feats.shape[0]))
        if n:
            mx = coords.amax(0).tolist()       # host read (coords are on the CPU in test_scene.py:81)
            mn = int(coords.amin().item())
            if mn < 0:
                raise ValueError('negative coordinate')
        else:
            mx = [0, 0, 0, -1]
        ssz = [int(v) for v in self.spatial_size]
        if n and any(mx[a] >= ssz[a] for a in range(3)):
            raise ValueError('coordinate %s outside spatial_size %s' % (mx[:3], ssz))
        bs = int(input[2]) if len(input) > 2 else mx[3] + 1
        # Grid extent: the declared spatial_size (scn semantics for the strided output sizes) whenever the bitmask
        # stays small; the reference's update_sizes quirk (App. C.2) declares absurd bounds at test time, then the
        # data extent rounded up to a multiple of 64 is used (identical results for <= 6 stride-2 levels).
        if max(bs, 1) * ssz[0] * ssz[1] * ssz[2] <= (1 << 31):
            dims = list(ssz)
        else:
            dims = [max(1, min(ssz[a], (mx[a] + 64) // 64 * 64)) for a in range(3)]
        coords = coords.to(dev).contiguous()
        md = Metadata(self.dimension)
        md.batch_size = bs
        md.grids[_key(self.spatial_size)] = E.build_grid(coords, max(bs, 1), dims)
        return SparseConvNetTensor(feats.float().contiguous(), md, self.spatial_size.clone())


class OutputLayer(nn.Module):
    def __init__(self, dimension):
        nn.Module.__init__(self)
        self.dimension = dimension

    def forward(self, input):
        return input.features.clone()


class SparseToDense(nn.Module):
    def __init__(self, dimension, nPlanes):
        nn.Module.__init__(self)
        self.dimension, self.nPlanes = dimension, nPlanes

    def forward(self, input):
        g = input.metadata.grid(input.spatial_size)
        ssz = [int(v) for v in input.spatial_size]
        f = input.features
        if f.shape[0] == 0:
            f = f.new_zeros((0, self.nPlanes))
        return E.sparse_to_dense(f, g.coords, input.metadata.batch_size, ssz)


# ------------------------------------------------------------ convolutions
class _ConvBase(nn.Module):
    def __init__(self, dimension, nIn, nOut, filter_size, bias, groups=1):
        nn.Module.__init__(self)
        if dimension != 3:
            raise NotImplementedError('dimension 3 only')
        if groups != 1:
            raise NotImplementedError('groups != 1')
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size = int(filter_size)
        self.filter_volume = self.filter_size ** dimension
        std = (2.0 / nIn / self.filter_volume) ** 0.5
        self.weight = nn.Parameter(torch.Tensor(self.filter_volume, nIn, nOut).normal_(0, std))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(nOut).zero_())

    def _w(self):
        w = self.weight.detach()
        if w.dim() == 4:          # later scn releases: [K^3, groups, Cin/g, Cout/g]
            w = w[:, 0]
        return w.contiguous()

    def _add_bias(self, f):
        if hasattr(self, 'bias'):
            f += self.bias.detach()
        return f

    def _check(self, input):
        if input.features.shape[0] and input.features.shape[1] != self.nIn:
            raise RuntimeError('%s: expected %d input planes, got %d' % (
                type(self).__name__, self.nIn, input.features.shape[1]))


class SubmanifoldConvolution(_ConvBase):
    """model.py:32,38,40,179,186,254."""

    def forward(self, input):
        self._check(input)
        n = input.features.shape[0]
        out = SparseConvNetTensor(None, input.metadata, input.spatial_size)
        f = torch.empty((n, self.nOut), dtype=torch.float32, device=input.features.device)
        if n:
            nbr = input.metadata.getSubmanifoldRuleBook(input.spatial_size, self.filter_size)
            E.conv(input.features, nbr, self._w(), n, f)
        out.features = self._add_bias(f)
        return out


class Convolution(_ConvBase):
    """model.py:44 (filter 2, stride 2)."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        _ConvBase.__init__(self, dimension, nIn, nOut, filter_size, bias, groups)
        self.filter_stride = int(filter_stride)

    def forward(self, input):
        self._check(input)
        out_size = (input.spatial_size - self.filter_size) // self.filter_stride + 1
        parent, children = input.metadata.getRuleBook(input.spatial_size, out_size, self.filter_size,
                                                      self.filter_stride)
        n_out = children.shape[1]
        f = torch.empty((n_out, self.nOut), dtype=torch.float32, device=input.features.device)
        if n_out:
            E.conv(input.features, children, self._w(), n_out, f)
        return SparseConvNetTensor(self._add_bias(f), input.metadata, out_size)


class Deconvolution(_ConvBase):
    """the stride-2 rulebook with the two sides swapped; the fine active set must
    already exist in the metadata (U-Net use)."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        _ConvBase.__init__(self, dimension, nIn, nOut, filter_size, bias, groups)
        self.filter_stride = int(filter_stride)

    def forward(self, input):
        self._check(input)
        out_size = (input.spatial_size - 1) * self.filter_stride + self.filter_size
        if _key(out_size) not in input.metadata.grids:
            raise RuntimeError('Deconvolution: fine active set %s not in metadata' % (_key(out_size),))
        parent, children = input.metadata.getRuleBook(out_size, input.spatial_size, self.filter_size,
                                                      self.filter_stride)
        f = torch.empty((parent.shape[0], self.nOut), dtype=torch.float32, device=input.features.device)
        if parent.shape[0]:
            E.deconv(input.features, parent, self._w(), f)
        return SparseConvNetTensor(self._add_bias(f), input.metadata, out_size)


class UnPooling(nn.Module):
    def __init__(self, dimension, pool_size, pool_stride):
        nn.Module.__init__(self)
        self.dimension, self.pool_size, self.pool_stride = dimension, int(pool_size), int(pool_stride)

    def forward(self, input):
        out_size = (input.spatial_size - 1) * self.pool_stride + self.pool_size
        if _key(out_size) not in input.metadata.grids:
            raise RuntimeError('UnPooling: fine active set %s not in metadata' % (_key(out_size),))
        parent, children = input.metadata.getRuleBook(out_size, input.spatial_size, self.pool_size,
                                                      self.pool_stride)
        f = torch.empty((parent.shape[0], input.features.shape[1]), dtype=torch.float32,
                        device=input.features.device)
        if parent.shape[0]:
            E.unpool(input.features, parent, f)
        return SparseConvNetTensor(f, input.metadata, out_size)


class NetworkInNetwork(nn.Module):
    def __init__(self, nIn, nOut, bias):
        nn.Module.__init__(self)
        std = (2.0 / nIn) ** 0.5
        self.nIn, self.nOut = nIn, nOut
        self.weight = nn.Parameter(torch.Tensor(nIn, nOut).normal_(0, std))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(nOut).zero_())

    def forward(self, input):
        x = input.features
        f = torch.empty((x.shape[0], self.nOut), dtype=torch.float32, device=x.device)
        b = self.bias.detach() if hasattr(self, 'bias') else None
        if x.shape[0]:
            E.linear(x, self.weight.detach().t().contiguous(), b, f)
        return SparseConvNetTensor(f, input.metadata, input.spatial_size)


# --------------------------------------------------------------- batchnorm
class BatchNormalization(nn.Module):
    """Eval-mode scn BatchNormalization (eps 1e-4, momentum 0.9 kept for state_dict parity)."""

    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, affine=True, leakiness=1):
        nn.Module.__init__(self)
        self.nPlanes, self.eps, self.momentum = nPlanes, eps, momentum
        self.affine, self.leakiness = affine, leakiness
        if leakiness not in (0, 1):
            raise NotImplementedError('leakiness %r' % (leakiness,))
        self.register_buffer('running_mean', torch.Tensor(nPlanes).fill_(0))
        self.register_buffer('running_var', torch.Tensor(nPlanes).fill_(1))
        if affine:
            self.weight = nn.Parameter(torch.Tensor(nPlanes).fill_(1))
            self.bias = nn.Parameter(torch.Tensor(nPlanes).fill_(0))
        else:
            self.register_buffer('weight', torch.ones(nPlanes))
            self.register_buffer('bias', torch.zeros(nPlanes))

    def forward(self, input):
        if self.training:
            raise NotImplementedError('sgnn_b200 is forward-inference only: call model.eval() '
                                      '(batch statistics / backward are out of scope, SURVEY §2 row 7)')
        x = input.features
        out = SparseConvNetTensor(None, input.metadata, input.spatial_size)
        if x.shape[0] == 0:
            out.features = x
            return out
        if x.shape[1] != self.nPlanes:
            raise RuntimeError('BatchNorm: expected %d planes, got %d' % (self.nPlanes, x.shape[1]))
        scale, shift = E.fold_bn(self)
        out.features = E.affine_relu(x, torch.empty_like(x), scale, shift, relu=(self.leakiness == 0))
        return out


class BatchNormReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9):
        BatchNormalization.__init__(self, nPlanes, eps, momentum, True, 0)


# --------------------------------------------------- network architectures
def FullyConvolutionalNet(dimension, reps, nPlanes, residual_blocks=False, downsample=[2, 2]):
    """scn.FullyConvolutionalNet (model.py:180,255) -- nesting of SURVEY App. A.9 so that state_dict keys
    line up with the reference checkpoints."""

    def block(m, a, b):
        if residual_blocks:
            m.add(ConcatTable()
                  .add(Identity() if a == b else NetworkInNetwork(a, b, False))
                  .add(Sequential()
                       .add(BatchNormReLU(a))
                       .add(SubmanifoldConvolution(dimension, a, b, 3, False))
                       .add(BatchNormReLU(b))
                       .add(SubmanifoldConvolution(dimension, b, b, 3, False)))
                  ).add(AddTable())
        else:
            m.add(Sequential()
                  .add(BatchNormReLU(a))
                  .add(SubmanifoldConvolution(dimension, a, b, 3, False)))

    def U(planes):
        m = Sequential()
        for _ in range(reps):
            block(m, planes[0], planes[0])
        if len(planes) > 1:
            m.add(ConcatTable()
                  .add(Identity())
                  .add(Sequential()
                       .add(BatchNormReLU(planes[0]))
                       .add(Convolution(dimension, planes[0], planes[1], downsample[0], downsample[1], False))
                       .add(U(planes[1:]))
                       .add(UnPooling(dimension, downsample[0], downsample[1]))))
            m.add(JoinTable())
        return m

    return U(list(nPlanes))
