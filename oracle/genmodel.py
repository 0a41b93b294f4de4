"""ORACLE -- TEST INFRASTRUCTURE ONLY: CPU restatement of the SG-NN generator graph.

/root/reference does not exist on the GPU box, so the reference's torch/model.py cannot be imported there.
This file restates its forward pass (model.py:371-416 and the pieces it calls) on top of oracle O2
(oracle/sparseconvnet) with plain CPU torch ops, keeping the reference attribute names so that
`state_dict()` keys match.  It is pinned: tests/test_oracle_genmodel.py (run where /root/reference exists)
checks it output-for-output against the UNMODIFIED reference model.py, and the committed fixtures under
tests/golden/ were produced by the reference file itself (tests/golden/make_golden.py).
The scn arithmetic underneath remains PARITY UNPINNED (see oracle/sparseconvnet/__init__.py).

Used as: checker in tests/, in __graft_entry__.smoke(), and as the timed subject of
bench.py's cpu_baseline / `--impl reference` arm ("port").
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
import sparseconvnet as scn  # noqa: E402  (oracle O2)


def _cbr(conv):
    return nn.Sequential(conv, nn.BatchNorm3d(conv.out_channels), nn.ReLU(True))


def _res(nf):
    inner = scn.Sequential()
    inner.add(scn.BatchNormReLU(nf)).add(scn.SubmanifoldConvolution(3, nf, nf, 3, False))
    inner.add(scn.BatchNormReLU(nf)).add(scn.SubmanifoldConvolution(3, nf, nf, 3, False))
    return scn.ConcatTable().add(scn.Identity()).add(inner)


class _EncLevel(nn.Module):
    # model.py:21-67
    def __init__(self, cin, nf, first, last, size):
        nn.Module.__init__(self)
        self.first, self.last = first, last
        if first:
            self.p0 = scn.InputLayer(3, size, mode=0)
        self.p1 = scn.SubmanifoldConvolution(3, cin, nf, 3, False)
        self.p2 = scn.Sequential().add(_res(nf)).add(scn.AddTable()).add(scn.BatchNormReLU(nf))
        self.p3 = scn.Sequential().add(scn.Convolution(3, nf, nf, 2, 2, False)).add(scn.BatchNormReLU(nf))
        if last:
            self.p4 = scn.SparseToDense(3, nf)


class _Encoder(nn.Module):
    # model.py:69-136
    def __init__(self, nf_in, per_level, nf_out, skip_dense, size):
        nn.Module.__init__(self)
        self.skip_dense = skip_dense
        n = len(per_level)
        self.process_sparse = nn.Sequential(*[
            _EncLevel(nf_in if i == 0 else per_level[i - 1], per_level[i], i == 0, i == n - 1,
                      (np.array(size) // (i + 1)).tolist()) for i in range(n)])
        nf = per_level[-1]
        nf0, nf1 = nf * 3 // 2, nf * 2
        self.encode_dense0 = _cbr(nn.Conv3d(nf, nf0, 4, 2, 1, bias=False))
        self.encode_dense1 = _cbr(nn.Conv3d(nf0, nf1, 4, 2, 1, bias=False))
        self.bottleneck_dense2 = _cbr(nn.Conv3d(nf1, nf1, 1, bias=False))
        nf3 = nf1 * 2 if skip_dense else nf1
        nf4 = nf3 // 2
        self.decode_dense3 = _cbr(nn.ConvTranspose3d(nf3, nf4, 4, 2, 1, bias=False))
        nf4 = nf4 + nf0 if skip_dense else nf4
        nf5 = nf4 // 2
        self.decode_dense4 = _cbr(nn.ConvTranspose3d(nf4, nf5, 4, 2, 1, bias=False))
        self.final = _cbr(nn.Conv3d(nf5, nf_out, 1, bias=False))
        self.occpred = nn.Sequential(nn.Conv3d(nf_out, 1, 1, bias=False))
        self.sdfpred = nn.Sequential(nn.Conv3d(nf_out, 1, 1, bias=False))


class _Refine(nn.Module):
    # model.py:169-190
    def __init__(self, nf_in, nf, size):
        nn.Module.__init__(self)
        self.p0 = scn.InputLayer(3, size, mode=0)
        self.p1 = scn.SubmanifoldConvolution(3, nf_in, nf, 3, False)
        self.p2 = scn.FullyConvolutionalNet(3, reps=1, nPlanes=[nf, nf, nf], residual_blocks=True)
        self.p3 = scn.BatchNormReLU(nf * 3)
        self.p4 = scn.OutputLayer(3)
        self.n0 = scn.InputLayer(3, size, mode=0)
        self.n1 = scn.SubmanifoldConvolution(3, nf * 3, nf, 3, False)
        self.n2 = scn.BatchNormReLU(nf)
        self.n3 = scn.OutputLayer(3)
        self.linear = nn.Linear(nf, 1)
        self.linearsdf = nn.Linear(nf, 1)


class _Surface(nn.Module):
    # model.py:249-258
    def __init__(self, nf_in, nf, size):
        nn.Module.__init__(self)
        self.p0 = scn.InputLayer(3, size, mode=0)
        self.p1 = scn.SubmanifoldConvolution(3, nf_in, nf, 3, False)
        self.p2 = scn.FullyConvolutionalNet(3, reps=1, nPlanes=[nf, nf, nf], residual_blocks=True)
        self.p3 = scn.BatchNormReLU(nf * 3)
        self.p4 = scn.OutputLayer(3)
        self.linear = nn.Linear(nf * 3, 1)


_CHILD = torch.tensor([[z, y, x, 0] for z in (0, 1) for y in (0, 1) for x in (0, 1)], dtype=torch.long)


class OracleGenModel(nn.Module):
    """Defaults of test_scene.py:29-39: pass_occ, pass_feats, both skips."""

    def __init__(self, encoder_dim=8, input_dim=64, input_nf=1, nf_coarse=16, nf=16, num_hierarchy_levels=4):
        nn.Module.__init__(self)
        L = num_hierarchy_levels
        size = [input_dim] * 3 if np.isscalar(input_dim) else [int(v) for v in input_dim]
        per = [int(encoder_dim * (1 + float(k) / (L - 2))) for k in range(L - 1)]       # model.py:286
        self.encoder = _Encoder(input_nf, per, nf_coarse, True, size)
        sizes = [(np.array(size) // (2 ** k)).tolist() for k in range(L - 1)][::-1]       # model.py:290
        per = per + [per[-1]]
        self.refinement = scn.Sequential()
        for h in range(1, L):                                                             # model.py:297-303
            self.refinement.add(_Refine(per[L - h] + 2 + (nf_coarse if h == 1 else nf), nf, sizes[h - 1]))
        self.surfacepred = _Surface(per[0] + 2 + nf, nf, sizes[-1])

    def set_sizes(self, size):
        """Exact (un-quirked) counterpart of update_sizes: only upper bounds of mode-0 InputLayers."""
        size = [int(v) for v in size]
        self.encoder.process_sparse[0].p0.spatial_size = torch.LongTensor(size)
        big = torch.LongTensor([s * 64 for s in size])
        for r in self.refinement:
            r.p0.spatial_size = big.clone()
            r.n0.spatial_size = big.clone()
        self.surfacepred.p0.spatial_size = big.clone()

    # ---- model.py:338-355, as a coordinate join
    @staticmethod
    def _join(skip_locs, skip_feats, locs, feats):
        if skip_locs.shape[0] == 0 or locs.shape[0] == 0:
            return feats
        ext = torch.maximum(skip_locs.max(0).values, locs.max(0).values) + 1

        def key(c):
            return ((c[:, 3] * ext[0] + c[:, 0]) * ext[1] + c[:, 1]) * ext[2] + c[:, 2]
        ks, kt = key(skip_locs), key(locs)
        order = torch.argsort(ks)
        pos = torch.searchsorted(ks[order], kt).clamp(max=ks.shape[0] - 1)
        hit = ks[order][pos] == kt
        add = skip_feats.new_zeros((locs.shape[0], skip_feats.shape[1]))
        add[hit] = skip_feats[order[pos[hit]]]
        return torch.cat([feats, add], 1)

    def forward(self, locs, feats, forced_keep=None):
        """locs LongTensor [N,4] (z,y,x,b), feats [N,1] -> ([locs, sdf], [[cand_locs, cand(occ,sdf)] x 4]).
        forced_keep (tests only, "teacher forcing"): one bool mask per level that REPLACES this oracle's own
        sigmoid(occ) > 0.5 decision for what continues to the next level (the reported logits stay the oracle's own), so
        that a device pass can be checked level by level over the whole input even when a logit within rounding distance
        of the threshold falls on the other side there."""
        enc = self.encoder
        x = [locs, feats]
        skips = []
        for lv in enc.process_sparse:                                                     # model.py:145-150
            if lv.first:
                x = lv.p0(x)
            s = lv.p2(lv.p1(x))
            x = lv.p3(s)
            skips.append(s)
            if lv.last:
                skips.append(x)
                x = lv.p4(x)
        e0 = enc.encode_dense0(x)                                                         # model.py:152-166
        e1 = enc.encode_dense1(e0)
        d0 = enc.decode_dense3(torch.cat([enc.bottleneck_dense2(e1), e1], 1))
        xd = enc.final(enc.decode_dense4(torch.cat([d0, e0], 1)))
        occ, sdf = enc.occpred(xd), enc.sdfpred(xd)
        B, C, D0, D1, D2 = xd.shape
        skips = [(s.metadata.getSpatialLocations(s.spatial_size), s.features) for s in skips]
        # ---- model.py:315-336
        zz, yy, xx = torch.meshgrid(torch.arange(D0), torch.arange(D1), torch.arange(D2), indexing='ij')
        cell = torch.stack([zz, yy, xx], -1).view(-1, 3)
        cand_locs = torch.cat([cell.repeat(B, 1), torch.arange(B).repeat_interleave(cell.shape[0]).view(-1, 1)], 1)
        cand = torch.stack([occ[:, 0].reshape(-1), sdf[:, 0].reshape(-1)], 1)
        keep = torch.sigmoid(cand[:, 0]) > 0.5
        if forced_keep is not None:
            keep = forced_keep[0].reshape(-1).bool()
        f = torch.cat([cand, xd.permute(0, 2, 3, 4, 1).reshape(-1, C)], 1)[keep]
        cur = cand_locs[keep]
        outputs = [[cand_locs, cand]]
        nref = len(self.refinement)
        for h, r in enumerate(self.refinement):                                           # model.py:387-396
            if cur.shape[0] == 0:
                outputs.append([[], []])
                continue
            f = self._join(skips[nref - h][0], skips[nref - h][1], cur, f)
            y = r.p4(r.p3(r.p2(r.p1(r.p0([cur, f])))))
            kids = (cur.unsqueeze(1) * torch.tensor([2, 2, 2, 1]) + _CHILD).view(-1, 4)   # model.py:192-207
            y = r.n3(r.n2(r.n1(r.n0([kids, y.repeat_interleave(8, 0)]))))
            cand = torch.cat([r.linear(y), r.linearsdf(y)], 1)                           # model.py:230-240
            keep = torch.sigmoid(cand[:, 0]) > 0.5
            if forced_keep is not None:
                keep = forced_keep[h + 1].reshape(-1).bool()
            outputs.append([kids, cand])
            cur, f = kids[keep], torch.cat([y[keep], cand[keep]], 1)
        if cur.shape[0] == 0:
            return [cur, []], outputs
        f = self._join(skips[0][0], skips[0][1], cur, f)                                  # model.py:400-402
        s = self.surfacepred
        return [cur, s.linear(s.p4(s.p3(s.p2(s.p1(s.p0([cur, f]))))))], outputs
