"""Python front of the native generator (csrc/generator.cu): flattens the module tree of GenModel into the
SgnnGeneratorW struct of device pointers (cached until a parameter changes), owns the workspace arena, makes
ONE C-ABI call per forward and wraps the results.  Same results, bit for bit, as fused.forward_fused /
GenModel.forward_modules; only the default SG-NN structure (test_scene.py:29-39) is supported natively."""
import ctypes as C
import itertools

import torch

from . import engine as E
from . import _lib
from ._lib import lib, check


def supported(model):
    """Default SG-NN structure, every parameter on one CUDA device, channel widths inside the kernels' limits (Cin <= 64)."""
    try:
        ok = (len(model.refinement) == 3 and len(model.encoder.process_sparse) == 3 and model.pass_occ and
              model.pass_feats and model.use_skip_sparse and model.encoder.use_skip_dense and model.PRED_SURF)
        devs = set(p.device for p in model.parameters())
        ok = ok and len(devs) == 1 and next(iter(devs)).type == 'cuda'
        widths = [model.surfacepred.p1.nIn] + [r.nf_in for r in model.refinement] + [3 * r.nf for r in model.refinement]
        return bool(ok and max(widths) <= 64)
    except Exception:
        return False


class _Weights(object):
    """Keeps every tensor referenced by the struct alive."""

    def __init__(self, model):
        self.keep = []
        self.tensors = list(itertools.chain(model.parameters(), model.buffers()))
        self.key = self.version_key(model, self.tensors)
        self.w = _lib.SgnnGeneratorW()
        self.dev = next(model.parameters()).device
        w = self.w
        enc = model.encoder
        for l, layer in enumerate(enc.process_sparse):
            e = w.enc[l]
            e.cin, e.c = layer.nf_in, layer.nf
            e.w_in = self.conv(layer.p1)
            self.res(e.res, layer.p2[0])
            self.bn(e.bn_out, layer.p2[2])
            e.w_down = self.conv(layer.p3[0])
            self.bn(e.bn_down, layer.p3[1])
        seqs = [enc.encode_dense0, enc.encode_dense1, enc.bottleneck_dense2, enc.decode_dense3, enc.decode_dense4,
                enc.final]
        cat = [-1, -1, -1, 1, 0, -1]
        for i, seq in enumerate(seqs):
            d = w.dense[i]
            conv = seq[0]
            d.w = self.t(conv.weight)
            self.bn(d.bn, seq[1])
            d.cout, d.ksize, d.stride, d.pad = conv.out_channels, conv.kernel_size[0], conv.stride[0], conv.padding[0]
            d.transposed = 1 if isinstance(conv, torch.nn.ConvTranspose3d) else 0
            d.cat_with = cat[i]
        w.w_heads = self.t(torch.cat([enc.occpred[0].weight.detach().reshape(1, -1),
                                      enc.sdfpred[0].weight.detach().reshape(1, -1)], 0))
        w.nf_coarse = enc.final[0].out_channels
        for h in range(3):
            r, m = w.ref[h], model.refinement[h]
            r.cin, r.c = m.nf_in, m.nf
            r.w_in = self.conv(m.p1)
            self.fcn(r.fcn, m.p2, m.p3)
            r.w_up = self.conv(m.n1)
            self.bn(r.bn_up, m.n2)
            r.w_occ, r.b_occ = self.t(m.linear.weight.view(-1)), self.t(m.linear.bias)
            r.w_sdf, r.b_sdf = self.t(m.linearsdf.weight.view(-1)), self.t(m.linearsdf.bias)
        s, m = w.surf, model.surfacepred
        s.cin, s.c = m.p1.nIn, m.p1.nOut
        s.w_in = self.conv(m.p1)
        self.fcn(s.fcn, m.p2, m.p3)
        s.w_lin, s.b_lin = self.t(m.linear.weight), self.t(m.linear.bias)
        # tensor-core filter banks, prepared once per weight set (sgnn_generator_prepare)
        nb = lib.sgnn_generator_prepared_bytes(C.byref(w))
        self.prepared = torch.empty(max(int(nb), 256), dtype=torch.uint8, device=self.dev)
        w.prepared, w.prepared_bytes = self.prepared.data_ptr(), int(nb)
        with torch.cuda.device(self.dev):
            check(lib.sgnn_generator_prepare(C.byref(w), C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)),
                  'sgnn_generator_prepare')

    @staticmethod
    def version_key(model, tensors=None):
        """(version counter, address) of every parameter and buffer.  `tensors`: the list cached by NativeGenerator -- walking the
        module tree (857 us) or building a state_dict (1.5 ms) on EVERY forward made the host, not the GPU, the bottleneck of
        the pass; over a cached list the key costs ~60 us.  In-place updates (optimizer steps, load_state_dict) bump the
        version, .to()/.cuda()/.half() change the address; replacing sub-MODULES after the first forward is not seen -- call
        GenModel.invalidate_native() after that kind of surgery."""
        if tensors is None:
            tensors = list(itertools.chain(model.parameters(), model.buffers()))
        return tuple((t._version, t.data_ptr()) for t in tensors)

    def t(self, x):
        x = x.detach().float().contiguous()
        self.keep.append(x)
        return x.data_ptr()

    def conv(self, m):
        wt = m.weight.detach()
        if wt.dim() == 4:
            wt = wt[:, 0]
        return self.t(wt)

    def bn(self, dst, m):
        s, t = E.fold_bn(m)
        self.keep += [s, t]
        dst.scale, dst.shift = s.data_ptr(), t.data_ptr()

    def res(self, dst, blk):
        seq = blk[1]
        self.bn(dst.bn0, seq[0])
        dst.w0 = self.conv(seq[1])
        self.bn(dst.bn1, seq[2])
        dst.w1 = self.conv(seq[3])

    def fcn(self, dst, fcn, p3):
        m = fcn
        dst.c = m[0][1][1].nOut
        for lvl in range(3):
            self.res(dst.blk[lvl], m[0])
            if lvl < 2:
                down = m[2][1]
                self.bn(dst.bn_down[lvl], down[0])
                dst.w_down[lvl] = self.conv(down[1])
                m = down[2]
        self.bn(dst.bn_join, p3)


class NativeGenerator(object):
    def __init__(self, model, arena_bytes=1 << 30):
        self.model = model
        self.weights = None
        self.arena = None
        self.arena_bytes = int(arena_bytes)
        self.last = None
        self.profile = False
        self._stale = None
        self.phases = False          # SGNN_GEN_PHASES: per-phase CUDA-event times of the next passes (phase_table())

    def stale(self):
        """True when the parameters changed since the weight struct was built (cheap: cached tensor list)."""
        self._stale = self.weights is None or self.weights.key != _Weights.version_key(self.model, self.weights.tensors)
        return self._stale

    def _prepare(self, dev):
        m = self.model
        stale, self._stale = self._stale, None           # GenModel.forward has just asked: do not compute the key twice
        if stale is None:
            stale = self.stale()
            self._stale = None
        if self.weights is None or self.weights.dev != dev or stale:
            self.weights = _Weights(m)
        if self.arena is None or self.arena.device != dev or self.arena.numel() < self.arena_bytes:
            self.arena = None
            self.arena = torch.empty(self.arena_bytes, dtype=torch.uint8, device=dev)

    def forward(self, locs, feats, want_cand_locs=True, nb=None, cand_parents=False):
        """locs int64/int32 [n,4] CUDA, feats fp32 [n,cin] CUDA -> raw SgnnGeneratorOut-backed tensors (clones)."""
        dev = feats.device
        m = self.model
        cin = m.encoder.process_sparse[0].nf_in
        if locs.dim() != 2 or locs.shape[1] != 4 or feats.dim() != 2 or feats.shape[0] != locs.shape[0]:
            raise ValueError('sgnn_b200: expected locs [n,4] and feats [n,%d], got %s and %s' % (cin, tuple(locs.shape), tuple(feats.shape)))
        if feats.shape[1] != cin:
            raise ValueError('sgnn_b200: the model takes %d input feature channel(s), got %d' % (cin, feats.shape[1]))
        pdev = self.weights.tensors[0].device if self.weights is not None else next(m.parameters()).device
        if pdev != dev:
            raise RuntimeError('sgnn_b200: model parameters on %s, features on %s' % (pdev, dev))
        with torch.cuda.device(dev):
            return self._forward(locs, feats, want_cand_locs, nb, dev, cand_parents)

    def _forward(self, locs, feats, want_cand_locs, nb, dev, cand_parents=False):
        self._prepare(dev)
        m = self.model
        dims = (C.c_int32 * 3)(*[int(v) for v in m.encoder.process_sparse[0].p0.spatial_size])
        if nb is None:
            nb = int(locs[:, 3].max().item()) + 1 if locs.shape[0] else 1
        out = _lib.SgnnGeneratorOut()
        self.weights.w.tc32_min_rows = int(getattr(m, 'tc32_min_rows', 0))
        self.weights.w.ur_min_rows = int(getattr(m, 'ur_min_rows', 0))
        self.weights.w.overlap_max_rows = int(getattr(m, 'overlap_max_rows', 0))
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        for _ in range(8):
            rc = lib.sgnn_generator_forward(C.byref(self.weights.w), C.c_void_p(locs.data_ptr()),
                                            1 if locs.dtype == torch.int64 else 0, C.c_void_p(feats.data_ptr()),
                                            locs.shape[0], nb, dims, C.c_void_p(self.arena.data_ptr()),
                                            self.arena.numel(),
                                            (_lib.GEN_CAND_PARENTS if cand_parents else (_lib.GEN_CAND_LOCS if want_cand_locs else 0)) |
                                            (_lib.GEN_PROFILE if self.profile else 0) |
                                            (_lib.GEN_PHASES if self.phases else 0) |
                                            (_lib.GEN_TC32 if getattr(m, 'conv_mode', 'tc32') == 'tc32' else 0) |
                                            (_lib.GEN_DENSE_RULES if getattr(m, 'dense_rules', False) else 0),
                                            C.byref(out), stream)
            if rc != -6:
                break
            self.arena_bytes = max(int(out.arena_needed), 2 * self.arena_bytes)
            self._prepare(dev)
        check(rc, 'sgnn_generator_forward')
        self.last = out
        return out, nb

    def phase_table(self):
        """[(label, ms)] of the last pass run with self.phases = True."""
        out, name, ms = [], C.create_string_buffer(48), C.c_float(0)
        i = 1
        while lib.sgnn_generator_phase_entry(i, name, C.byref(ms)) == 0:
            out.append((name.value.decode(), float(ms.value)))
            i += 1
        return out

    def view(self, ptr, shape, dtype):
        """Tensor view into the arena (valid until the next forward)."""
        n = 1
        for s in shape:
            n *= s
        if n == 0:
            return torch.empty(shape, dtype=dtype, device=self.arena.device)
        off = ptr - self.arena.data_ptr()
        esz = torch.empty(0, dtype=dtype).element_size()
        return self.arena[off:off + n * esz].view(dtype).view(shape)


def forward_native(model, x, loss_weights):
    from .model import _dense_cell_coords
    if model.training:
        raise NotImplementedError('sgnn_b200 is forward-inference only: call model.eval()')
    locs, feats = x[0], x[1]
    if not feats.is_cuda:
        raise RuntimeError('sgnn_b200.GenModel: features must be a CUDA tensor (no CPU fallback)')
    dev = feats.device
    if model._native is None:
        model._native = NativeGenerator(model)
    g = model._native
    if locs.shape[0] == 0:
        raise RuntimeError('sgnn_b200.GenModel: empty input (scn raises on an empty InputLayer batch too)')
    locs = locs.to(dev).contiguous()
    if locs.dtype not in (torch.int64, torch.int32):
        locs = locs.long()
    feats = feats.float().contiguous()
    ssz = [int(v) for v in model.encoder.process_sparse[0].p0.spatial_size]
    with torch.no_grad(), torch.cuda.device(dev):
        out, nb = g.forward(locs, feats, nb=int(x[2]) if len(x) > 2 else None, cand_parents=True)
        # Every result tensor is written by ONE kernel (sgnn_export): fresh tensors as the reference returns them (the arena is
        # recycled by the next pass), int64 coordinates when model.return_long, the candidate coordinates expanded from the
        # parents on the way.  (Tensor-by-tensor formatting was ~50 framework calls and ~0.9 ms of host time per pass.)
        to64 = 1 if model.return_long else 0
        cdt = torch.int64 if to64 else torch.int32
        dd = list(ssz)
        for _ in range(3):
            dd = [(d - 2) // 2 + 1 for d in dd]
        segs = (_lib.SgnnExportSeg * 16)()
        ns = 0

        def seg(src, dst, n, kind, aux=None):
            nonlocal ns
            sg = segs[ns]
            sg.src, sg.dst, sg.n, sg.kind, sg.to_i64 = src, dst.data_ptr(), n, kind, to64
            if aux is not None:
                sg.aux[0], sg.aux[1], sg.aux[2], sg.aux[3] = aux
            ns += 1
        outputs = []
        n0 = int(out.n_cand[0])
        l0 = torch.empty((n0, 4), dtype=cdt, device=dev)
        c0 = torch.empty((n0, 2), dtype=torch.float32, device=dev)
        seg(None, l0, n0, _lib.EXPORT_DENSE_CELLS, (nb, dd[0], dd[1], dd[2]))
        seg(out.cand[0], c0, 2 * n0, _lib.EXPORT_COPY32)
        outputs.append([l0, c0])
        for h in range(1, 4):
            n = int(out.n_cand[h])
            if n == 0:
                outputs.append([[], []])
                continue
            lh = torch.empty((n, 4), dtype=cdt, device=dev)
            ch = torch.empty((n, 2), dtype=torch.float32, device=dev)
            seg(out.cand_locs[h], lh, n, _lib.EXPORT_CHILDREN)
            seg(out.cand[h], ch, 2 * n, _lib.EXPORT_COPY32)
            outputs.append([lh, ch])
        result = [[], []]
        if out.n_out:
            ol = torch.empty((out.n_out, 4), dtype=cdt, device=dev)
            os_ = torch.empty((out.n_out, 1), dtype=torch.float32, device=dev)
            seg(out.out_locs, ol, int(out.n_out), _lib.EXPORT_COORDS)
            seg(out.out_sdf, os_, int(out.n_out), _lib.EXPORT_COPY32)
            result = [ol, os_]
        check(lib.sgnn_export(segs, ns, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), 'sgnn_export')
        return result, outputs
