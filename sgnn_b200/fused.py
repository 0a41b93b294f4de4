"""Fused forward of the SG-NN generator on the B200 engine (the hot path that bench.py times).

Arithmetic is IDENTICAL (bit for bit) to GenModel.forward_modules(): the same fmaf chains, the same folded
BatchNorm constants, the same literal sigmoid mask.  What changes is data movement:
  * every BatchNormReLU (51 per pass), residual add and JoinTable concat is folded into the
    epilogue of the convolution (or unpool) that produces its input -- no standalone pass over HBM;
  * the x8 child replication of model.py:202 (192 B per child) is never written: the n1 convolution runs in
    child mode straight off the parent rulebook and parent feature rows;
  * concat_skip writes the encoder channels into spare columns of the level's input buffer (one grid probe per
    row) instead of building two dense int64 indicator volumes (model.py:346-349);
  * one grid + one neighbour table per site set, shared by all convolutions at that resolution.
"""
import torch

from . import engine as E


def _pad4(c):
    return (c + 3) // 4 * 4


def _w(conv):
    w = conv.weight.detach()
    if w.dim() == 4:
        w = w[:, 0]
    return w.contiguous()


class _Level(object):
    """A site set with its submanifold neighbour table."""
    __slots__ = ('grid', 'nbr', 'n')

    def __init__(self, grid):
        self.grid = grid
        self.n = grid.n
        self.nbr = E.rulebook_submanifold(grid) if grid.n else None


def _new(n, c, dev):
    return torch.empty((n, c), dtype=torch.float32, device=dev)


def _res_block(lv, blk, x_raw, x_bn, out_a, a_bn=None, out_b=None, b_bn=None):
    """ConcatTable(Identity, Seq(BNReLU, SMC, BNReLU, SMC)) + AddTable.  x_bn = BNReLU_0(x_raw) was emitted by
    the producer of x_raw.  Result y = SMC(BNReLU(SMC(x_bn))) + x_raw goes to out_a (affine a_bn) / out_b."""
    seq = blk[1]
    dev = x_raw.device
    s2, t2 = E.fold_bn(seq[2])
    mid = _new(lv.n, seq[1].nOut, dev)
    E.conv(x_bn, lv.nbr, _w(seq[1]), lv.n, mid, scale_a=s2, shift_a=t2, relu_a=True)
    sa, ta = a_bn if a_bn is not None else (None, None)
    sb, tb = b_bn if b_bn is not None else (None, None)
    E.conv(mid, lv.nbr, _w(seq[3]), lv.n, out_a, residual=x_raw, scale_a=sa, shift_a=ta, relu_a=a_bn is not None,
           out_b=out_b, scale_b=sb, shift_b=tb, relu_b=b_bn is not None)


def _fcn_bn(fcn, p3, lv0, x_raw, x_bn, dev):
    """FullyConvolutionalNet(reps=1, [c,c,c], residual) followed by BatchNormReLU(3c) (model.py:180-181,
    255-256).  Returns [n,3c].  x_raw is the p1 output, x_bn = BNReLU_{block0.bn0}(x_raw)."""
    s3, t3 = E.fold_bn(p3)
    # walk the nesting: U = Seq(ConcatTable(Id, resSeq), AddTable, ConcatTable(Id, Seq(BN, Conv, U', UnPool)), Join)
    c = fcn[0][1][1].nOut
    m = fcn
    depth = 1
    while len(m) > 2:
        m = m[2][1][2]
        depth += 1
    width = c * depth
    J0 = _new(lv0.n, width, dev)
    if lv0.n == 0:
        return J0

    def run(m, lv, x_raw, x_bn, J, top):
        """Writes U(m)(x) into J[:, :] (J has exactly the channels U(m) emits).  At the top level the p3
        BatchNormReLU is applied on the way out."""
        blk = m[0]
        cc = blk[1][1].nOut
        last = len(m) == 2
        if last:
            _res_block(lv, blk, x_raw, x_bn, J[:, :cc], a_bn=(s3[:cc], t3[:cc]) if top else None)
            return
        down = m[2][1]          # Seq(BNReLU, Convolution, U', UnPooling)
        sd, td = E.fold_bn(down[0])
        y_bn = _new(lv.n, cc, dev)
        _res_block(lv, blk, x_raw, x_bn, J[:, :cc], a_bn=(s3[:cc].contiguous(), t3[:cc].contiguous()) if top else None,
                   out_b=y_bn, b_bn=(sd, td))
        # stride-2 convolution to the coarse set; emits raw + BNReLU of the next block's first BN
        cg = E.coarsen(lv.grid)
        clv = _Level(cg)
        parent, children = E.rulebook_strided(lv.grid, cg)
        sub = down[2]
        cn = down[1].nOut
        z_raw = _new(cg.n, cn, dev)
        z_bn = _new(cg.n, cn, dev)
        sn, tn = E.fold_bn(sub[0][1][0])
        Jc_w = J.shape[1] - cc
        Jc = _new(cg.n, Jc_w, dev)
        if cg.n:
            E.conv(y_bn, children, _w(down[1]), cg.n, z_raw, out_b=z_bn, scale_b=sn, shift_b=tn, relu_b=True)
            run(sub, clv, z_raw, z_bn, Jc, False)
        if top:
            E.unpool(Jc, parent, J[:, cc:], scale=s3[cc:].contiguous(), shift=t3[cc:].contiguous(), relu=True)
        else:
            E.unpool(Jc, parent, J[:, cc:])

    run(fcn, lv0, x_raw, x_bn, J0, True)
    return J0


def _encoder_level(layer, lv, x_in, dev):
    """SparseEncoderLayer (model.py:49-67) fused.  Returns (skip features, coarse grid, coarse features)."""
    c = layer.nf
    blk = layer.p2[0]
    s0, t0 = E.fold_bn(blk[1][0])
    a_raw, a_bn = _new(lv.n, c, dev), _new(lv.n, c, dev)
    E.conv(x_in, lv.nbr, _w(layer.p1), lv.n, a_raw, out_b=a_bn, scale_b=s0, shift_b=t0, relu_b=True)
    skip = _new(lv.n, c, dev)
    _res_block(lv, blk, a_raw, a_bn, skip, a_bn=E.fold_bn(layer.p2[2]))
    cg = E.coarsen(lv.grid)
    parent, children = E.rulebook_strided(lv.grid, cg)
    sc, tc = E.fold_bn(layer.p3[1])
    h = _new(cg.n, c, dev)
    if cg.n:
        E.conv(skip, children, _w(layer.p3[0]), cg.n, h, scale_a=sc, shift_a=tc, relu_a=True)
    return skip, cg, h


def forward_fused(model, x, loss_weights):
    from .model import _dense_cell_coords
    if model.training:
        raise NotImplementedError('sgnn_b200 is forward-inference only: call model.eval()')
    locs_in, feats_in = x[0], x[1]
    if not feats_in.is_cuda:
        raise RuntimeError('sgnn_b200.GenModel: features must be a CUDA tensor (no CPU fallback)')
    dev = feats_in.device
    enc = model.encoder
    with torch.no_grad():
        outputs = []
        # ---------------- encoder: level-0 grid from caller coordinates (a1, a2), declared extent like the native path
        il = enc.process_sparse[0].p0
        locs_dev = locs_in.to(dev).contiguous()
        if locs_dev.dtype not in (torch.int64, torch.int32):
            locs_dev = locs_dev.long()
        nb = int(x[2]) if len(x) > 2 else (int(locs_dev[:, 3].max().item()) + 1 if locs_dev.shape[0] else 1)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        g = E.build_grid(locs_dev, nb, [int(v) for v in il.spatial_size], status=status)
        if int(status.item()):
            raise ValueError('coordinate outside spatial_size %s' % il.spatial_size.tolist())
        feats = feats_in.float().contiguous()
        skips = []
        for li, layer in enumerate(enc.process_sparse):
            lv = _Level(g)
            if lv.n == 0:
                raise RuntimeError('empty input')
            skip, cg, h = _encoder_level(layer, lv, feats, dev)
            skips.append((g, skip))
            g, feats = cg, h
        skips.append((g, feats))                                    # ft3 (model.py:64)
        ssz = il.spatial_size.clone()
        for _ in enc.process_sparse:
            ssz = (ssz - 2) // 2 + 1
        ddims = [int(v) for v in ssz]
        dense = E.sparse_to_dense(feats, g.coords, nb, ddims)       # a7
        xd, outd = enc.dense_unet(dense)                            # a12 (library)
        # ---------------- a8: dense -> sparse
        nref = len(model.refinement)
        both = model.pass_feats and model.pass_occ
        cs0 = skips[nref][1].shape[1] if model.use_skip_sparse else 0
        c0 = xd.shape[1] + 2
        ld0 = _pad4(c0 + cs0) if both else None
        locs, feats, cand, m = E.dense_to_sparse(xd.contiguous(), outd.contiguous(), ld_feats=ld0)
        if not both:
            feats = feats[:, 2:c0].contiguous() if model.pass_feats else feats[:, :2].contiguous()
            c0 = feats.shape[1]
        outputs.append([model._locs_out(_dense_cell_coords(nb, ddims, dev)), cand])
        cur_c = c0          # live channels in feats (feats may be wider: room for the skip join)
        dims = list(ddims)
        for h in range(nref):
            ref = model.refinement[h]
            if not (loss_weights[h + 1] > 0):
                outputs.append([[], []])
                continue
            if m == 0:
                outputs.append([[], []])
                locs, feats = [], []
                continue
            # a10: skip join into spare columns
            if model.use_skip_sparse:
                sg, sf = skips[nref - h]
                if sf.shape[0]:
                    if feats.shape[1] < cur_c + sf.shape[1]:
                        wide = torch.zeros((m, _pad4(cur_c + sf.shape[1])), dtype=torch.float32, device=dev)
                        E.copy_cols(feats[:, :cur_c], wide[:, :cur_c])
                        feats = wide
                    E.concat_skip(sg, sf, locs, feats, cur_c)
                    cur_c += sf.shape[1]
            xin = feats[:, :cur_c]
            grid = E.build_grid(locs, nb, dims)
            lv = _Level(grid)
            nf = ref.nf
            blk0 = ref.p2[0]
            s0, t0 = E.fold_bn(blk0[1][0])
            a_raw, a_bn = _new(m, nf, dev), _new(m, nf, dev)
            E.conv(xin, lv.nbr, _w(ref.p1), m, a_raw, out_b=a_bn, scale_b=s0, shift_b=t0, relu_b=True)
            x48 = _fcn_bn(ref.p2, ref.p3, lv, a_raw, a_bn, dev)
            # a9: children (never materialised) -> n1 (child mode) + n2 -> heads + mask + compaction
            sn, tn = E.fold_bn(ref.n2)
            xc = _new(8 * m, nf, dev)
            E.conv(x48, lv.nbr, _w(ref.n1), 8 * m, xc, child_mode=True, scale_a=sn, shift_a=tn, relu_a=True)
            nxt = h + 1
            cs_next = 0
            if model.use_skip_sparse:
                cs_next = skips[nref - nxt][1].shape[1] if nxt < nref else skips[0][1].shape[1]
            ldn = _pad4(nf + 2 + cs_next) if both else None
            nlocs, nfeats, cand, m2 = E.heads_compact(
                xc, ref.linear.weight.detach().view(-1), ref.linear.bias.detach(),
                ref.linearsdf.weight.detach().view(-1), ref.linearsdf.bias.detach(), locs, ld_feats=ldn)
            outputs.append([model._locs_out(E.children_coords(locs)), cand])
            if both:
                cur_c = nf + 2
            elif model.pass_feats:
                nfeats = nfeats[:, :nf].contiguous()
                cur_c = nf
            else:
                nfeats = nfeats[:, nf:nf + 2].contiguous()
                cur_c = 2
            locs, feats, m = nlocs, nfeats, m2
            dims = [2 * v for v in dims]
        if not (model.PRED_SURF and loss_weights[-1] > 0):
            return [[], []], outputs
        if isinstance(locs, list) or m == 0:
            return [model._locs_out(locs), []], outputs
        # ---------------- a11: surface prediction
        sp_ = model.surfacepred
        if model.use_skip_sparse:
            sg, sf = skips[0]
            if sf.shape[0]:
                if feats.shape[1] < cur_c + sf.shape[1]:
                    wide = torch.zeros((m, _pad4(cur_c + sf.shape[1])), dtype=torch.float32, device=dev)
                    E.copy_cols(feats[:, :cur_c], wide[:, :cur_c])
                    feats = wide
                E.concat_skip(sg, sf, locs, feats, cur_c)
                cur_c += sf.shape[1]
        xin = feats[:, :cur_c]
        grid = E.build_grid(locs, nb, dims)
        lv = _Level(grid)
        nf = sp_.p1.nOut
        s0, t0 = E.fold_bn(sp_.p2[0][1][0])
        a_raw, a_bn = _new(m, nf, dev), _new(m, nf, dev)
        E.conv(xin, lv.nbr, _w(sp_.p1), m, a_raw, out_b=a_bn, scale_b=s0, shift_b=t0, relu_b=True)
        x48 = _fcn_bn(sp_.p2, sp_.p3, lv, a_raw, a_bn, dev)
        sdf = _new(m, sp_.linear.out_features, dev)
        E.linear(x48, sp_.linear.weight.detach(), sp_.linear.bias.detach(), sdf)
        return [model._locs_out(locs), sdf], outputs
