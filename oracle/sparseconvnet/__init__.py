"""ORACLE O2 -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement (torch + numpy, no CUDA) of the 12-symbol `sparseconvnet`
operator surface that /root/reference/torch/model.py imports at model.py:7
and calls at model.py:31-47,178-188,253-257,296,380.

PARITY UNPINNED for the scn arithmetic: the real library
(facebookresearch/SparseConvNet, version unpinned, README.md:11 of the
reference) is neither vendored under /root/reference nor installable offline,
and the reference ships no tests or golden vectors.  This file restates the
library's published algorithm:

  * hash of active sites -> row id, rows keep caller order (InputLayer mode 0)
  * submanifold rulebook: per filter offset k (row-major over (dz,dy,dx), last
    fastest) the list of (in_row, out_row) pairs whose neighbour exists
  * forward: out[rules[k].out] += in[rules[k].in] @ W[k], k ascending,
    W laid out [K^3, Cin, Cout]                (index_select -> mm -> index_add_)
  * Convolution(f, s): out site q covers q*s .. q*s+f-1, offset k = row-major
    index of (p - q*s); Deconvolution/UnPooling reuse that rulebook with the
    roles of the two sides swapped
  * BatchNormReLU: eps 1e-4, momentum 0.9 in the scn sense
  * FullyConvolutionalNet nesting of Appendix A.9

What IS pinned: the graph and glue of the generator.  The unmodified reference
file /root/reference/torch/model.py runs on top of this module (see
tests/golden/make_golden.py), and this module's convolutions are checked
against an independent dense torch.nn.functional.conv3d identity (oracle O1,
oracle/dense_equiv.py).

Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import this.
"""
import numpy as np
import torch
import torch.nn as nn

__all__ = [
    'InputLayer', 'OutputLayer', 'SubmanifoldConvolution', 'Convolution',
    'Deconvolution', 'UnPooling', 'BatchNormReLU', 'BatchNormalization',
    'Sequential', 'ConcatTable', 'AddTable', 'JoinTable', 'Identity',
    'SparseToDense', 'FullyConvolutionalNet', 'NetworkInNetwork',
    'SparseConvNetTensor', 'Metadata',
]


def _to_long3(dimension, x):
    if isinstance(x, torch.Tensor):
        return x.clone().long().view(-1)
    if isinstance(x, (list, tuple, np.ndarray)):
        assert len(x) == dimension
        return torch.LongTensor([int(v) for v in x])
    return torch.LongTensor([int(x)] * dimension)


def _key(ssz):
    return tuple(int(v) for v in ssz)


class _SiteSet(object):
    """Active sites at one spatial size: coords [N,4] (z,y,x,b) in row order
    plus a sorted linear-key index standing in for scn's per-sample hash maps."""

    def __init__(self, coords):
        coords = coords.long().contiguous()
        self.coords = coords
        c = coords.numpy()
        n = c.shape[0]
        if n == 0:
            self.ext = np.array([1, 1, 1, 1], dtype=np.int64)
        else:
            self.ext = c.max(axis=0).astype(np.int64) + 1
        self.n = n
        keys = self.keys_of(c)
        # stable sort; for duplicate coords (mode 0 does not merge) the LATER row
        # owns the hash entry, as a later insert overwrites the map value.
        order = np.argsort(keys, kind='stable')
        self.sorted_keys = keys[order]
        self.order = order

    def keys_of(self, c):
        """Linear key with a one-cell apron so that -1 / +extent never alias."""
        e = self.ext
        return (((c[:, 3] * (e[0] + 2) + (c[:, 0] + 1)) * (e[1] + 2) + (c[:, 1] + 1))
                * (e[2] + 2) + (c[:, 2] + 1))

    def lookup(self, c):
        """rows of coords c ([M,4] int64 numpy) or -1."""
        e = self.ext
        if self.n == 0 or c.shape[0] == 0:
            return np.full(c.shape[0], -1, dtype=np.int64)
        inb = ((c[:, 0] >= -1) & (c[:, 0] <= e[0]) & (c[:, 1] >= -1) & (c[:, 1] <= e[1]) &
               (c[:, 2] >= -1) & (c[:, 2] <= e[2]) & (c[:, 3] >= 0) & (c[:, 3] < e[3]))
        k = self.keys_of(np.where(inb[:, None], c, 0))
        pos = np.searchsorted(self.sorted_keys, k, side='right') - 1
        pos_c = np.clip(pos, 0, self.n - 1)
        hit = inb & (pos >= 0) & (self.sorted_keys[pos_c] == k)
        return np.where(hit, self.order[pos_c], -1)


class Metadata(object):
    """Per-tensor-chain cache of site sets and rulebooks (scn Metadata<3>)."""

    def __init__(self, dimension=3):
        self.dimension = dimension
        self.sites = {}
        self.rb_sub = {}
        self.rb_conv = {}
        self.batch_size = 0

    def setInput(self, spatial_size, coords, batch_size):
        self.sites[_key(spatial_size)] = _SiteSet(coords)
        self.batch_size = batch_size

    def getSpatialLocations(self, spatial_size):
        return self.sites[_key(spatial_size)].coords.clone()

    def nActive(self, spatial_size):
        return self.sites[_key(spatial_size)].n

    # --- SURVEY App. A.3 ------------------------------------------------
    def getSubmanifoldRuleBook(self, spatial_size, filter_size):
        k = (_key(spatial_size), int(filter_size))
        if k in self.rb_sub:
            return self.rb_sub[k]
        ss = self.sites[_key(spatial_size)]
        f = int(filter_size)
        assert f % 2 == 1
        h = f // 2
        c = ss.coords.numpy()
        rules = []
        rows = np.arange(ss.n, dtype=np.int64)
        for dz in range(-h, h + 1):
            for dy in range(-h, h + 1):
                for dx in range(-h, h + 1):
                    q = c.copy()
                    q[:, 0] += dz
                    q[:, 1] += dy
                    q[:, 2] += dx
                    # bounds: neighbour must be a legal coordinate of the extent
                    r = ss.lookup(q)
                    ok = r >= 0
                    rules.append((torch.from_numpy(r[ok]), torch.from_numpy(rows[ok])))
        self.rb_sub[k] = rules
        return rules

    # --- SURVEY App. A.5 ------------------------------------------------
    def getRuleBook(self, in_size, out_size, filter_size, filter_stride):
        k = (_key(in_size), _key(out_size), int(filter_size), int(filter_stride))
        if k in self.rb_conv:
            return self.rb_conv[k]
        f, s = int(filter_size), int(filter_stride)
        ss = self.sites[_key(in_size)]
        c = ss.coords.numpy()
        osz = np.array(_key(out_size), dtype=np.int64)
        cand_in, cand_q, cand_k = [], [], []
        rows = np.arange(ss.n, dtype=np.int64)
        kk = 0
        for oz in range(f):
            for oy in range(f):
                for ox in range(f):
                    o = np.array([oz, oy, ox], dtype=np.int64)
                    d = c[:, :3] - o
                    ok = np.all(d >= 0, axis=1) & np.all(d % s == 0, axis=1)
                    q = d // s
                    ok &= np.all(q < osz, axis=1)
                    qq = np.concatenate([q[ok], c[ok, 3:4]], axis=1)
                    cand_in.append(rows[ok])
                    cand_q.append(qq)
                    cand_k.append(np.full(int(ok.sum()), kk, dtype=np.int64))
                    kk += 1
        cin = np.concatenate(cand_in)
        cq = np.concatenate(cand_q, axis=0)
        ck = np.concatenate(cand_k)
        if _key(out_size) in self.sites:
            oss = self.sites[_key(out_size)]
            orow = oss.lookup(cq)
            keep = orow >= 0
            cin, ck, orow = cin[keep], ck[keep], orow[keep]
        else:
            # new coarse set; row id = first touch walking the INPUT rows in order
            # (upstream: iteration order of the input hash map -- implementation
            # defined; consumers are order independent)
            if cq.shape[0] == 0:
                uq = np.zeros((0, 4), dtype=np.int64)
                orow = np.zeros((0,), dtype=np.int64)
            else:
                touch = np.lexsort((ck, cin))          # by input row, then offset
                cq_t = cq[touch]
                uq_s, first, inv = np.unique(cq_t, axis=0, return_index=True, return_inverse=True)
                rank_of_sorted = np.argsort(np.argsort(first, kind='stable'), kind='stable')
                orow_t = rank_of_sorted[inv.reshape(-1)]
                uq = np.zeros_like(uq_s)
                uq[rank_of_sorted] = uq_s
                orow = np.empty_like(orow_t)
                orow[touch] = orow_t
            self.sites[_key(out_size)] = _SiteSet(torch.from_numpy(uq))
        rules = []
        for k_ in range(f ** 3):
            m = ck == k_
            rules.append((torch.from_numpy(cin[m]), torch.from_numpy(orow[m])))
        self.rb_conv[k] = rules
        return rules


class SparseConvNetTensor(object):
    def __init__(self, features=None, metadata=None, spatial_size=None):
        self.features = features
        self.metadata = metadata
        self.spatial_size = spatial_size

    def get_spatial_locations(self, spatial_size=None):
        if spatial_size is None:
            spatial_size = self.spatial_size
        return self.metadata.getSpatialLocations(spatial_size)

    def cpu(self):
        self.features = self.features.cpu()
        return self

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
            f = f + i.features
        out.features = f
        return out


class JoinTable(nn.Module):
    def forward(self, input):
        out = SparseConvNetTensor(None, input[0].metadata, input[0].spatial_size)
        out.features = torch.cat([i.features for i in input], 1)
        return out


# --------------------------------------------------------------- IO layers
class InputLayer(nn.Module):
    """scn.InputLayer(dimension, spatial_size, mode) -- model.py:31,178,185,253.
    mode 0: coordinates are promised unique; rows keep caller order (App. A.2)."""

    def __init__(self, dimension, spatial_size, mode=3):
        nn.Module.__init__(self)
        self.dimension = dimension
        self.spatial_size = _to_long3(dimension, spatial_size)
        self.mode = mode

    def forward(self, input):
        coords, feats = input[0], input[1]
        coords = coords.detach().cpu().long()
        if coords.shape[1] == self.dimension:
            coords = torch.cat([coords, torch.zeros(coords.shape[0], 1, dtype=torch.long)], 1)
        if self.mode != 0:
            raise NotImplementedError('oracle restates mode 0 only (the mode model.py uses)')
        bs = int(input[2]) if len(input) > 2 else (int(coords[:, 3].max()) + 1 if coords.shape[0] else 0)
        md = Metadata(self.dimension)
        md.setInput(self.spatial_size, coords, bs)
        return SparseConvNetTensor(feats, md, self.spatial_size.clone())


class OutputLayer(nn.Module):
    def __init__(self, dimension):
        nn.Module.__init__(self)
        self.dimension = dimension

    def forward(self, input):
        return input.features.clone()


class SparseToDense(nn.Module):
    def __init__(self, dimension, nPlanes):
        nn.Module.__init__(self)
        self.dimension = dimension
        self.nPlanes = nPlanes

    def forward(self, input):
        ssz = [int(v) for v in input.spatial_size]
        c = input.metadata.getSpatialLocations(input.spatial_size)
        bs = input.metadata.batch_size
        f = input.features
        out = f.new_zeros((bs, f.shape[1], ssz[0], ssz[1], ssz[2]))
        if c.shape[0]:
            out[c[:, 3], :, c[:, 0], c[:, 1], c[:, 2]] = f
        return out


# ------------------------------------------------------------ convolutions
def _rule_forward(rules, in_feats, weight, n_out, swap=False):
    out = in_feats.new_zeros((n_out, weight.shape[2]))
    for k, (ri, ro) in enumerate(rules):
        if swap:
            ri, ro = ro, ri
        if ri.numel() == 0:
            continue
        out.index_add_(0, ro, in_feats.index_select(0, ri).mm(weight[k]))
    return out


class SubmanifoldConvolution(nn.Module):
    """model.py:32,38,40,179,186,254.  W [K^3, Cin, Cout], N(0, sqrt(2/(Cin K^3)))."""

    def __init__(self, dimension, nIn, nOut, filter_size, bias, groups=1):
        nn.Module.__init__(self)
        assert groups == 1
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size = int(filter_size)
        self.filter_volume = self.filter_size ** dimension
        std = (2.0 / nIn / self.filter_volume) ** 0.5
        self.weight = nn.Parameter(torch.Tensor(self.filter_volume, nIn, nOut).normal_(0, std))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(nOut).zero_())

    def forward(self, input):
        assert input.features.numel() == 0 or input.features.shape[1] == self.nIn
        rules = input.metadata.getSubmanifoldRuleBook(input.spatial_size, self.filter_size)
        out = SparseConvNetTensor(None, input.metadata, input.spatial_size)
        f = _rule_forward(rules, input.features, self.weight, input.features.shape[0])
        if hasattr(self, 'bias'):
            f = f + self.bias
        out.features = f
        return out


class Convolution(nn.Module):
    """model.py:44 (filter 2, stride 2)."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        nn.Module.__init__(self)
        assert groups == 1
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size, self.filter_stride = int(filter_size), int(filter_stride)
        self.filter_volume = self.filter_size ** dimension
        std = (2.0 / nIn / self.filter_volume) ** 0.5
        self.weight = nn.Parameter(torch.Tensor(self.filter_volume, nIn, nOut).normal_(0, std))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(nOut).zero_())

    def forward(self, input):
        out_size = (input.spatial_size - self.filter_size) // self.filter_stride + 1
        rules = input.metadata.getRuleBook(input.spatial_size, out_size, self.filter_size, self.filter_stride)
        n_out = input.metadata.nActive(out_size)
        out = SparseConvNetTensor(None, input.metadata, out_size)
        f = _rule_forward(rules, input.features, self.weight, n_out)
        if hasattr(self, 'bias'):
            f = f + self.bias
        out.features = f
        return out


class Deconvolution(nn.Module):
    """App. A.6: the Convolution rulebook fine->coarse with the roles swapped; the
    fine active set must already live in the metadata (U-Net use)."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        nn.Module.__init__(self)
        assert groups == 1
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size, self.filter_stride = int(filter_size), int(filter_stride)
        self.filter_volume = self.filter_size ** dimension
        std = (2.0 / nIn / self.filter_volume) ** 0.5
        self.weight = nn.Parameter(torch.Tensor(self.filter_volume, nIn, nOut).normal_(0, std))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(nOut).zero_())

    def forward(self, input):
        out_size = (input.spatial_size - 1) * self.filter_stride + self.filter_size
        if _key(out_size) not in input.metadata.sites:
            raise RuntimeError('Deconvolution: fine active set %s not in metadata' % (_key(out_size),))
        rules = input.metadata.getRuleBook(out_size, input.spatial_size, self.filter_size, self.filter_stride)
        n_out = input.metadata.nActive(out_size)
        out = SparseConvNetTensor(None, input.metadata, out_size)
        f = _rule_forward(rules, input.features, self.weight, n_out, swap=True)
        if hasattr(self, 'bias'):
            f = f + self.bias
        out.features = f
        return out


class UnPooling(nn.Module):
    def __init__(self, dimension, pool_size, pool_stride):
        nn.Module.__init__(self)
        self.dimension = dimension
        self.pool_size, self.pool_stride = int(pool_size), int(pool_stride)

    def forward(self, input):
        out_size = (input.spatial_size - 1) * self.pool_stride + self.pool_size
        if _key(out_size) not in input.metadata.sites:
            raise RuntimeError('UnPooling: fine active set %s not in metadata' % (_key(out_size),))
        rules = input.metadata.getRuleBook(out_size, input.spatial_size, self.pool_size, self.pool_stride)
        n_out = input.metadata.nActive(out_size)
        f = input.features.new_zeros((n_out, input.features.shape[1]))
        for (r_fine, r_coarse) in rules:
            if r_fine.numel():
                f.index_add_(0, r_fine, input.features.index_select(0, r_coarse))
        out = SparseConvNetTensor(f, input.metadata, out_size)
        return out


class NetworkInNetwork(nn.Module):
    def __init__(self, nIn, nOut, bias):
        nn.Module.__init__(self)
        std = (2.0 / nIn) ** 0.5
        self.weight = nn.Parameter(torch.Tensor(nIn, nOut).normal_(0, std))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(nOut).zero_())

    def forward(self, input):
        out = SparseConvNetTensor(None, input.metadata, input.spatial_size)
        f = input.features.mm(self.weight)
        if hasattr(self, 'bias'):
            f = f + self.bias
        out.features = f
        return out


# --------------------------------------------------------------- batchnorm
class BatchNormalization(nn.Module):
    """App. A.8: eps 1e-4; running = momentum*running + (1-momentum)*batch."""

    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, affine=True, leakiness=1):
        nn.Module.__init__(self)
        self.nPlanes, self.eps, self.momentum = nPlanes, eps, momentum
        self.affine, self.leakiness = affine, leakiness
        self.register_buffer('running_mean', torch.Tensor(nPlanes).fill_(0))
        self.register_buffer('running_var', torch.Tensor(nPlanes).fill_(1))
        if affine:
            self.weight = nn.Parameter(torch.Tensor(nPlanes).fill_(1))
            self.bias = nn.Parameter(torch.Tensor(nPlanes).fill_(0))

    def forward(self, input):
        x = input.features
        out = SparseConvNetTensor(None, input.metadata, input.spatial_size)
        if x.numel() == 0:
            out.features = x
            return out
        assert x.shape[1] == self.nPlanes
        if self.training:
            mean = x.mean(0)
            var_b = ((x - mean) ** 2).mean(0)
            n = x.shape[0]
            with torch.no_grad():
                self.running_mean.mul_(self.momentum).add_(mean.detach() * (1 - self.momentum))
                unb = var_b.detach() * (n / max(n - 1, 1))
                self.running_var.mul_(self.momentum).add_(unb * (1 - self.momentum))
            inv = (var_b + self.eps).pow(-0.5)
        else:
            mean = self.running_mean
            inv = (self.running_var + self.eps).pow(-0.5)
        w = inv * self.weight if self.affine else inv
        b = (-mean * w + self.bias) if self.affine else (-mean * w)
        y = x * w + b
        if self.leakiness != 1:
            y = torch.where(y > 0, y, y * self.leakiness)
        out.features = y
        return out


class BatchNormReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9):
        BatchNormalization.__init__(self, nPlanes, eps, momentum, True, 0)


# --------------------------------------------------- network architectures
def FullyConvolutionalNet(dimension, reps, nPlanes, residual_blocks=False, downsample=[2, 2]):
    """App. A.9 nesting (needed for state_dict key layout)."""

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
