"""Host-side mirror of the SG-NN generator (reference torch/model.py:276-416) on the B200 engine.

`GenModel` keeps the reference constructor signature (model.py:277), forward signature and return
structure (model.py:371,415), `update_sizes` (model.py:357) and -- attribute for attribute -- the module
tree, so `state_dict()` keys/shapes equal the reference's and `load_state_dict(checkpoint['state_dict'])`
(test_scene.py:61-62) works unchanged.

Two forward paths over the same parameters:
  * forward_modules(): the reference-shaped composition, every scn.* call replaced by the B200-backed
    module of sgnn_b200.scn and every piece of Python glue (dense_coarse_to_sparse model.py:315,
    concat_skip :338, to_next_level_locs :192, heads+mask :230-247) by one device kernel;
  * forward() = forward_fused(): same arithmetic (bit-identical results), but BatchNormReLU / residual add /
    channel concat are folded into convolution epilogues, the x8 child replication of model.py:202 is never
    materialised (child-mode convolution), and per-resolution grids/rulebooks are shared.
The dense 8^3 U-Net (model.py:89-136,152-166) runs on the engine's fixed-order direct convolutions (dense.cu),
so the whole generator is bit-reproducible; nn.Conv3d / BatchNorm3d modules only hold the parameters.
"""
import numpy as np
import torch
import torch.nn as nn

from . import scn
from . import engine as E

FSIZE0 = 3
FSIZE1 = 2


def _as3(v):
    if isinstance(v, (list, tuple, np.ndarray)):
        return [int(a) for a in v]
    return [int(v)] * 3


class SparseEncoderLayer(nn.Module):
    """model.py:21-67: SMC, residual block, BNReLU, (skip), stride-2 conv + BNReLU, optional densify."""

    def __init__(self, nf_in, nf, input_sparsetensor, return_sparsetensor, max_data_size):
        nn.Module.__init__(self)
        self.nf_in, self.nf = nf_in, nf
        self.input_sparsetensor, self.return_sparsetensor = input_sparsetensor, return_sparsetensor
        self.max_data_size = max_data_size
        if not input_sparsetensor:
            self.p0 = scn.InputLayer(3, max_data_size, mode=0)
        self.p1 = scn.SubmanifoldConvolution(3, nf_in, nf, filter_size=FSIZE0, bias=False)
        res = scn.Sequential()
        res.add(scn.BatchNormReLU(nf)).add(scn.SubmanifoldConvolution(3, nf, nf, FSIZE0, False))
        res.add(scn.BatchNormReLU(nf)).add(scn.SubmanifoldConvolution(3, nf, nf, FSIZE0, False))
        self.p2 = scn.Sequential()
        self.p2.add(scn.ConcatTable().add(scn.Identity()).add(res)).add(scn.AddTable())
        self.p2.add(scn.BatchNormReLU(nf))
        self.p3 = scn.Sequential().add(scn.Convolution(3, nf, nf, FSIZE1, 2, False))
        self.p3.add(scn.BatchNormReLU(nf))
        if not return_sparsetensor:
            self.p4 = scn.SparseToDense(3, nf)

    def forward(self, x):
        if not self.input_sparsetensor:
            x = self.p0(x)
        skip = self.p2(self.p1(x))
        x = self.p3(skip)
        if self.return_sparsetensor:
            return x, [skip]
        return self.p4(x), [skip, x]


def _cbr(conv):
    return nn.Sequential(conv, nn.BatchNorm3d(conv.out_channels), nn.ReLU(True))


class TSDFEncoder(nn.Module):
    """model.py:69-167: sparse pyramid + dense 8^3 U-Net with occupancy / sdf heads."""

    def __init__(self, nf_in, nf_per_level, nf_out, use_skip_sparse, use_skip_dense, input_volume_size):
        nn.Module.__init__(self)
        assert isinstance(nf_per_level, list)
        self.use_skip_sparse, self.use_skip_dense = use_skip_sparse, use_skip_dense
        self.use_bias = False
        sizes = [(np.array(input_volume_size) // (k + 1)).tolist() for k in range(len(nf_per_level))]
        levels = []
        for lv, nf_l in enumerate(nf_per_level):
            cin = nf_in if lv == 0 else nf_per_level[lv - 1]
            levels.append(SparseEncoderLayer(cin, nf_l, lv > 0, lv < len(nf_per_level) - 1, sizes[lv]))
        self.process_sparse = nn.Sequential(*levels)
        nf = nf_per_level[-1]
        nf0, nf1 = nf * 3 // 2, nf * 2
        nf2 = nf1
        b = self.use_bias
        self.encode_dense0 = _cbr(nn.Conv3d(nf, nf0, kernel_size=4, stride=2, padding=1, bias=b))
        self.encode_dense1 = _cbr(nn.Conv3d(nf0, nf1, kernel_size=4, stride=2, padding=1, bias=b))
        self.bottleneck_dense2 = _cbr(nn.Conv3d(nf1, nf2, kernel_size=1, bias=b))
        nf3 = nf2 if not use_skip_dense else nf1 + nf2
        nf4 = nf3 // 2
        self.decode_dense3 = _cbr(nn.ConvTranspose3d(nf3, nf4, kernel_size=4, stride=2, padding=1, bias=b))
        if use_skip_dense:
            nf4 += nf0
        nf5 = nf4 // 2
        self.decode_dense4 = _cbr(nn.ConvTranspose3d(nf4, nf5, kernel_size=4, stride=2, padding=1, bias=b))
        self.final = _cbr(nn.Conv3d(nf5, nf_out, kernel_size=1, bias=b))
        self.occpred = nn.Sequential(nn.Conv3d(nf_out, 1, kernel_size=1, bias=b))
        self.sdfpred = nn.Sequential(nn.Conv3d(nf_out, 1, kernel_size=1, bias=b))

    def dense_unet(self, x):
        """model.py:152-166 on the engine's fixed-order direct convolutions: BatchNorm3d + ReLU
        folded into each convolution, torch.cat of model.py:156,160 read in place from its two sources."""
        def cbr(seq, x0, x1=None, transposed=False):
            conv, bn = seq[0], seq[1]
            s, t = E.fold_bn(bn)
            k, st, pd = conv.kernel_size[0], conv.stride[0], conv.padding[0]
            return E.dense_conv(x0, x1, conv.weight.detach(), conv.out_channels, k, st, pd, s, t, True, transposed)
        x = x.contiguous()
        enc0 = cbr(self.encode_dense0, x)
        enc1 = cbr(self.encode_dense1, enc0)
        bott = cbr(self.bottleneck_dense2, enc1)
        dec0 = cbr(self.decode_dense3, bott, enc1 if self.use_skip_dense else None, True)
        x = cbr(self.decode_dense4, dec0, enc0 if self.use_skip_dense else None, True)
        x = cbr(self.final, x)
        w2 = torch.cat([self.occpred[0].weight.detach(), self.sdfpred[0].weight.detach()], 0).contiguous()
        return x, E.dense_conv(x, None, w2, 2, 1, 1, 0)

    def forward(self, x):
        skips = []
        for layer in self.process_sparse:
            x, ft = layer(x)
            if self.use_skip_sparse:
                skips.extend(ft)
        x, out = self.dense_unet(x)
        return x, out, skips


class Refinement(nn.Module):
    """model.py:169-247: one coarse-to-fine level."""

    def __init__(self, nf_in, nf, pass_occ, pass_feats, max_data_size, truncation=3):
        nn.Module.__init__(self)
        self.pass_occ, self.pass_feats = pass_occ, pass_feats
        self.nf_in, self.nf, self.truncation = nf_in, nf, truncation
        self.p0 = scn.InputLayer(3, max_data_size, mode=0)
        self.p1 = scn.SubmanifoldConvolution(3, nf_in, nf, filter_size=FSIZE0, bias=False)
        self.p2 = scn.FullyConvolutionalNet(3, reps=1, nPlanes=[nf, nf, nf], residual_blocks=True)
        self.p3 = scn.BatchNormReLU(nf * 3)
        self.p4 = scn.OutputLayer(3)
        self.n0 = scn.InputLayer(3, max_data_size, mode=0)
        self.n1 = scn.SubmanifoldConvolution(3, nf * 3, nf, filter_size=FSIZE0, bias=False)
        self.n2 = scn.BatchNormReLU(nf)
        self.n3 = scn.OutputLayer(3)
        self.linear = nn.Linear(nf, 1)
        self.linearsdf = nn.Linear(nf, 1)

    def forward(self, x):
        """x = [locs int32 [M,4] cuda, feats [M,nf_in]] -> ([locs', feats'], [cand_locs, cand(occ,sdf)])."""
        locs = x[0]
        if len(locs) == 0:
            return [[], []], [[], []]
        f = self.p4(self.p3(self.p2(self.p1(self.p0(x)))))
        cand_locs = E.children_coords(locs)
        f8 = f.repeat_interleave(8, dim=0)                      # model.py:202 (reference-shaped path only)
        f = self.n3(self.n2(self.n1(self.n0([cand_locs, f8]))))
        keep_locs, keep_feats, cand, m = E.heads_compact(
            f, self.linear.weight.detach().view(-1), self.linear.bias.detach(),
            self.linearsdf.weight.detach().view(-1), self.linearsdf.bias.detach(), locs)
        if self.pass_feats and self.pass_occ:
            feats = keep_feats
        elif self.pass_feats:
            feats = keep_feats[:, :self.nf].contiguous()
        else:
            feats = keep_feats[:, self.nf:].contiguous()
        return [keep_locs, feats], [cand_locs, cand]


class SurfacePrediction(nn.Module):
    """model.py:249-272."""

    def __init__(self, nf_in, nf, nf_out, max_data_size):
        nn.Module.__init__(self)
        self.p0 = scn.InputLayer(3, max_data_size, mode=0)
        self.p1 = scn.SubmanifoldConvolution(3, nf_in, nf, filter_size=FSIZE0, bias=False)
        self.p2 = scn.FullyConvolutionalNet(3, reps=1, nPlanes=[nf, nf, nf], residual_blocks=True)
        self.p3 = scn.BatchNormReLU(nf * 3)
        self.p4 = scn.OutputLayer(3)
        self.linear = nn.Linear(nf * 3, nf_out)

    def forward(self, x):
        if len(x[0]) == 0:
            return [], []
        f = self.p4(self.p3(self.p2(self.p1(self.p0(x)))))
        out = torch.empty((f.shape[0], self.linear.out_features), dtype=torch.float32, device=f.device)
        return E.linear(f, self.linear.weight.detach(), self.linear.bias.detach(), out)


class GenModel(nn.Module):
    def __init__(self, encoder_dim, input_dim, input_nf, nf_coarse, nf, num_hierarchy_levels, pass_occ,
                 pass_feats, use_skip_sparse, use_skip_dense, truncation=3):
        nn.Module.__init__(self)
        self.truncation, self.pass_occ, self.pass_feats = truncation, pass_occ, pass_feats
        input_dim = _as3(input_dim)
        L = num_hierarchy_levels
        if L > 2:
            self.nf_per_level = [int(encoder_dim * (1 + float(k) / (L - 2))) for k in range(L - 1)]
        else:
            self.nf_per_level = [encoder_dim] * (L - 1)
        self.use_skip_sparse = use_skip_sparse
        self.encoder = TSDFEncoder(input_nf, self.nf_per_level, nf_coarse, use_skip_sparse, use_skip_dense,
                                   input_volume_size=input_dim)
        self.refine_sizes = [(np.array(input_dim) // (2 ** k)).tolist() for k in range(L - 1)][::-1]
        self.nf_per_level.append(self.nf_per_level[-1])
        self.data_dim = 3
        self.refinement = scn.Sequential()
        for h in range(1, L):
            nf_in = self.nf_per_level[L - h] if use_skip_sparse else 0
            nf_in += 2 if pass_occ else 0
            nf_in += (nf_coarse if h == 1 else nf) if pass_feats else 0
            self.refinement.add(Refinement(nf_in, nf, pass_occ, pass_feats, self.refine_sizes[h - 1],
                                           truncation=truncation))
        self.PRED_SURF = True
        nf_in = self.nf_per_level[0] if use_skip_sparse else 0
        nf_in += 2 if pass_occ else 0
        nf_in += nf if pass_feats else 0
        self.surfacepred = SurfacePrediction(nf_in, nf, 1, self.refine_sizes[-1])
        self.return_long = True      # LongTensor coordinates at the boundary, like the reference
        self._native = None
        self._native_ok = None       # cached native.supported(self); re-evaluated when the parameters change or move
        # 'tc32' (default): the wide (Cout = 16) convolutions of the native generator with >= 60000 output rows run on
        # the tensor cores (tcgen05, exact 3-way bf16 split of fp32 features and filters: fp32 accuracy, but not the
        # fmaf order of the FFMA kernels);  'exact': every convolution on the fixed-order FFMA kernels -- bit-reproducible,
        # == oracle/o3.c, and bit-identical to forward_fused / forward_modules.
        self.conv_mode = 'tc32'
        self.tc32_min_rows = 0       # row thresholds of the tensor-core paths (0 = library defaults, SgnnGeneratorW)
        self.ur_min_rows = 0
        self.overlap_max_rows = 0    # levels up to this many rows build their coarse site sets on a side stream (0: default, -1: never)
        self.dense_rules = False     # A/B: dense neighbour table on the encoder's input level (default: compact rulebook; same bits)

    # model.py:357-369.  The sizes are upper bounds of mode-0 InputLayers; the reference doubles
    # refine_max_dim inside the k loop -- mirrored, not "fixed": bounds only grow.
    def update_sizes(self, input_max_dim, refine_max_dim):
        input_max_dim = np.array(_as3(input_max_dim))
        refine_max_dim = np.array(_as3(refine_max_dim))
        for k in range(3):
            self.encoder.process_sparse[0].p0.spatial_size[k] = int(input_max_dim[k])
            for h in range(len(self.refinement)):
                self.refinement[h].p0.spatial_size[k] = int(refine_max_dim[k])
                refine_max_dim = refine_max_dim * 2
                self.refinement[h].n0.spatial_size[k] = int(refine_max_dim[k])
            self.surfacepred.p0.spatial_size[k] = int(refine_max_dim[k])

    # ------------------------------------------------------------------ helpers
    def _locs_out(self, locs):
        if isinstance(locs, list):
            return locs
        return E.coords_to_i64(locs) if self.return_long else locs

    def _skip_join(self, skip, x_sparse):
        """concat_skip (model.py:338-355): skip = (grid, features)."""
        locs, feats = x_sparse
        if skip[1].shape[0] == 0 or len(locs) == 0:
            return x_sparse
        cs = skip[1].shape[1]
        out = torch.empty((feats.shape[0], feats.shape[1] + cs), dtype=torch.float32, device=feats.device)
        E.copy_cols(feats, out[:, :feats.shape[1]])
        E.concat_skip(skip[0], skip[1], locs, out, feats.shape[1])
        return [locs, out]

    # ------------------------------------------------------ reference-shaped path
    def forward_modules(self, x, loss_weights):
        locs_in, feats_in = x[0], x[1]
        dev = feats_in.device
        with torch.no_grad():
            outputs = []
            dense, out, skips = self.encoder([locs_in, feats_in])
            batch_size = dense.shape[0]
            skip_sets = []
            if self.use_skip_sparse:
                for s in skips:
                    skip_sets.append((s.metadata.grid(s.spatial_size), s.features))
            locs, feats, cand, _ = E.dense_to_sparse(dense.contiguous(), out.contiguous())
            if not (self.pass_feats and self.pass_occ):
                feats = feats[:, 2:].contiguous() if self.pass_feats else feats[:, :2].contiguous()
            d = dense.shape[2:]
            cand_locs = _dense_cell_coords(batch_size, d, dev)
            outputs.append([self._locs_out(cand_locs), cand])
            x_sparse = [locs, feats]
            nref = len(self.refinement)
            for h in range(nref):
                if loss_weights[h + 1] > 0:
                    if self.use_skip_sparse:
                        x_sparse = self._skip_join(skip_sets[nref - h], x_sparse)
                    x_sparse, occ = self.refinement[h](x_sparse)
                    outputs.append([self._locs_out(occ[0]), occ[1]])
                else:
                    outputs.append([[], []])
            locs = x_sparse[0]
            if self.PRED_SURF and loss_weights[-1] > 0:
                if self.use_skip_sparse:
                    x_sparse = self._skip_join(skip_sets[0], x_sparse)
                sdf = self.surfacepred(x_sparse)
                if isinstance(sdf, tuple):
                    sdf = []
                return [self._locs_out(locs), sdf], outputs
            return [[], []], outputs

    def forward(self, x, loss_weights):
        """Fast path: the native generator (one C-ABI call, csrc/generator.cu) when the model has the default
        SG-NN structure and every level is requested; otherwise the Python-orchestrated fused path."""
        from . import native
        ok = self._native_ok
        if ok is None or self._native is None or self._native.stale():     # first call, or the parameters changed / moved
            ok = self._native_ok = native.supported(self)
        if ok and all(float(v) > 0 for v in loss_weights):
            return native.forward_native(self, x, loss_weights)
        return self.forward_fused(x, loss_weights)

    def invalidate_native(self):
        """Forget the native generator's cached weight struct and structure check (after replacing sub-modules)."""
        self._native = None
        self._native_ok = None

    def forward_fused(self, x, loss_weights):
        from .fused import forward_fused
        return forward_fused(self, x, loss_weights)


def _dense_cell_coords(nb, d, device):
    """All cells of a dense [nb, d0, d1, d2] grid in (batch-major, raster) order (model.py:319-321)."""
    d0, d1, d2 = int(d[0]), int(d[1]), int(d[2])
    idx = torch.arange(nb * d0 * d1 * d2, device=device, dtype=torch.int32)
    x = idx % d2
    y = (idx // d2) % d1
    z = (idx // (d1 * d2)) % d0
    b = idx // (d0 * d1 * d2)
    return torch.stack([z, y, x, b], 1).contiguous()
