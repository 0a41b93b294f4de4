"""-m gpu: the whole generator through the drop-in boundary.

Parity gates (north_star): TSDF head within 1e-3 fp32, occupancy coordinates bit-exact -- against (1) the golden
fixtures the UNMODIFIED reference model.py produced on oracle O2 (tests/golden/make_golden.py), (2) the oracle
generator run live on the CPU, and, at BASELINE.json sizes, through size-independent properties
(fused == module-by-module bit for bit, determinism, row-order / batch-composition invariance).
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from helpers import sorted_rows, compare_generator_outputs

pytestmark = pytest.mark.gpu
ONES = np.ones(5, dtype=np.float32)
TOL_SDF = 1e-3          # north_star tolerance on the TSDF regression head
TOL_LOGIT = 1e-4        # occupancy logits / candidate sdf (tighter than required)


def _model(dims, seed):
    import sgnn_b200
    from sgnn_b200.synth import fill_parameters
    m = sgnn_b200.GenModel(8, list(dims), 1, 16, 16, 4, True, True, 1, 1)
    fill_parameters(m, seed)
    m.conv_mode = 'exact'       # this file pins the bit-reproducible mode (three orchestration paths, same bits);
    return m.cuda().eval()      # the tensor-core mode is covered by tests/test_gpu_tc32.py


def _same_outputs(a, b):
    (la, sa), lva = a
    (lb, sb), lvb = b
    assert torch.equal(la, lb) and torch.equal(sa, sb)
    assert len(lva) == len(lvb)
    for x, y in zip(lva, lvb):
        assert torch.equal(x[0], y[0]) and torch.equal(x[1], y[1])


@pytest.mark.parametrize('name', ['b2_s32', 'ragged', 'b1_s64'])
def test_golden_fixture_from_reference_model(name):
    g = np.load(os.path.join(GOLDEN, 'sgnn_ref_%s.npz' % name))
    m = _model(g['dims'], int(g['param_seed']))
    locs = torch.from_numpy(g['in_locs'].astype(np.int64))          # CPU LongTensor, as test_scene.py:81 leaves it
    feats = torch.from_numpy(g['in_feats']).cuda()
    (out_locs, out_sdf), levels = m([locs, feats], ONES)
    # fixture margin: no oracle logit within 2e-5 of the threshold -> coordinates must match EXACTLY, in order
    assert float(g['margin']) > 2e-5
    for i, l in enumerate(levels):
        assert np.array_equal(l[0].cpu().numpy(), g['cand%d_locs' % i].astype(np.int64)), 'level %d candidates' % i
        assert np.abs(l[1].cpu().numpy() - g['cand%d' % i]).max() <= TOL_LOGIT
        kept = int((torch.sigmoid(l[1][:, 0]) > 0.5).sum())
        assert kept == int(g['kept'][i])
    assert np.array_equal(out_locs.cpu().numpy(), g['out_locs'].astype(np.int64))
    assert np.abs(out_sdf.cpu().numpy() - g['out_sdf']).max() <= TOL_SDF
    # the native one-call generator, the Python-orchestrated fused path and the reference-shaped (module by
    # module) path give the SAME bits
    _same_outputs(((out_locs, out_sdf), levels), m.forward_fused([locs, feats], ONES))
    _same_outputs(((out_locs, out_sdf), levels), m.forward_modules([locs, feats], ONES))


@pytest.mark.parametrize('dims,nb,occ,seed', [((32, 32, 64), 2, 0.07, 21), ((32, 32, 32), 3, 0.1, 22)])
def test_against_live_oracle_generator(dims, nb, occ, seed):
    from genmodel import OracleGenModel
    from sgnn_b200.synth import fill_parameters, synthetic_batch
    locs, feats = synthetic_batch(nb, list(dims), occ, seed0=seed)
    ora = OracleGenModel(input_dim=list(dims))
    fill_parameters(ora, seed)
    ora.eval()
    with torch.no_grad():
        (wl, ws), wlv = ora(locs, feats)
    m = _model(dims, seed)
    got = m([locs.cuda(), feats.cuda()], ONES)
    compare_generator_outputs(((wl, ws), wlv), got, margin=1e-5, tol_logit=TOL_LOGIT, tol_sdf=TOL_SDF, tag='exact %s' % (dims,))


@pytest.mark.parametrize('mode', ['exact', 'tc32'])
def test_baseline_config_against_live_oracle(mode):
    """BASELINE.json configs[1] at FULL size (32 synthetic 64^3 blocks @5 %): the CPU oracle generator, run live, against both
    convolution modes of the native generator -- candidate coordinates at every level, kept coordinates and the TSDF head
    (<= 1e-3), block by block with counted margin flips."""
    from genmodel import OracleGenModel
    from sgnn_b200.synth import fill_parameters, synthetic_batch
    locs, feats = synthetic_batch(32, 64, 0.05)
    ora = OracleGenModel()
    fill_parameters(ora, 0)
    ora.eval()
    with torch.no_grad():
        want = ora(locs, feats)
    m = _model((64, 64, 64), 0)
    m.conv_mode = mode
    got = m([locs.cuda(), feats.cuda(), 32], ONES)
    flips, div = compare_generator_outputs(want, got, margin=1e-5, tol_logit=TOL_LOGIT, tol_sdf=TOL_SDF, tag='configs[1] ' + mode)
    assert len(div) <= 2                      # at most two of the 32 blocks may leave the comparison through legal flips


def test_row_order_and_batch_composition_invariance():
    from sgnn_b200.synth import synthetic_batch
    dims = (32, 32, 32)
    m = _model(dims, 0)
    locs, feats = synthetic_batch(2, list(dims), 0.08)
    base = m([locs.cuda(), feats.cuda()], ONES)
    perm = torch.randperm(locs.shape[0], generator=torch.Generator().manual_seed(0))
    shuf = m([locs[perm].cuda(), feats[perm].cuda()], ONES)
    # encoder rows are permuted, but everything downstream of the dense grid is keyed by coordinates
    a = sorted_rows(base[0][0].cpu().numpy(), base[0][1].cpu().numpy())
    b = sorted_rows(shuf[0][0].cpu().numpy(), shuf[0][1].cpu().numpy())
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # a block gives the same result alone and inside a batch (eval BN is per row; blocks share no rules)
    for bidx in range(2):
        sel = locs[:, 3] == bidx
        l1 = locs[sel].clone()
        l1[:, 3] = 0
        one = m([l1.cuda(), feats[sel].cuda()], ONES)
        in_batch = base[0][0][:, 3].cpu() == bidx
        x = sorted_rows(one[0][0].cpu().numpy()[:, :3], one[0][1].cpu().numpy())
        y = sorted_rows(base[0][0].cpu().numpy()[in_batch.numpy()][:, :3], base[0][1].cpu().numpy()[in_batch.numpy()])
        if x[0].shape == y[0].shape and np.array_equal(x[0], y[0]):        # cuDNN may pick another algo per batch size
            assert np.abs(x[1] - y[1]).max() <= TOL_SDF


def test_baseline_config_properties():
    """BASELINE.json configs[1]: 32 synthetic 64^3 blocks @5 %, full 3-level generator, fp32."""
    from sgnn_b200.synth import synthetic_batch
    m = _model((64, 64, 64), 0)
    locs, feats = synthetic_batch(32, 64, 0.05)
    a = m([locs.cuda(), feats.cuda()], ONES)
    b = m([locs.cuda(), feats.cuda()], ONES)
    _same_outputs(a, b)                                             # deterministic (no atomics on the float path)
    _same_outputs(a, m.forward_fused([locs.cuda(), feats.cuda()], ONES))
    _same_outputs(a, m.forward_modules([locs.cuda(), feats.cuda()], ONES))
    (fl, fs), lv = a
    assert lv[0][0].shape[0] == 32 * 512
    for i in range(1, 4):
        kept_prev = int((torch.sigmoid(lv[i - 1][1][:, 0]) > 0.5).sum())
        assert lv[i][0].shape[0] == 8 * kept_prev                  # 8 children per kept parent (model.py:192-207)
    assert fl.shape[0] == int((torch.sigmoid(lv[3][1][:, 0]) > 0.5).sum()) and fs.shape == (fl.shape[0], 1)
    assert torch.isfinite(fs).all()
    assert int(fl[:, :3].max()) < 64 and int(fl[:, 3].max()) < 32


def test_dropin_sparseconvnet_modules_against_oracle():
    """`import sparseconvnet as scn` bound to the engine: module-level parity with oracle O2."""
    import sys
    from conftest import ROOT
    import sparseconvnet as o2                  # oracle (conftest path)
    import sgnn_b200.scn as scn
    from helpers import random_coords
    rng = np.random.default_rng(3)
    dims = [16, 16, 16]
    c = torch.from_numpy(random_coords(rng, 2, dims, 0.2))
    f = torch.from_numpy(rng.standard_normal((c.shape[0], 4)).astype(np.float32))

    def net(lib):
        torch.manual_seed(0)
        s = lib.Sequential()
        s.add(lib.SubmanifoldConvolution(3, 4, 16, 3, False))
        s.add(lib.FullyConvolutionalNet(3, reps=1, nPlanes=[16, 16, 16], residual_blocks=True))
        s.add(lib.BatchNormReLU(48))
        s.add(lib.Convolution(3, 48, 16, 2, 2, False))
        s.add(lib.Deconvolution(3, 16, 8, 2, 2, False))
        return s
    a, b = net(o2).eval(), net(scn)
    b.load_state_dict(a.state_dict())
    b = b.cuda().eval()
    with torch.no_grad():
        ta = a(o2.InputLayer(3, dims, mode=0)([c, f]))
        tb = b(scn.InputLayer(3, dims, mode=0)([c, f.cuda()]))
        assert torch.equal(tb.metadata.getSpatialLocations(tb.spatial_size).cpu(),
                           ta.metadata.getSpatialLocations(ta.spatial_size))
        assert torch.allclose(scn.OutputLayer(3)(tb).cpu(), o2.OutputLayer(3)(ta), atol=1e-4, rtol=1e-4)
        da = o2.SparseToDense(3, 8)(ta)
        db = scn.SparseToDense(3, 8)(tb)
        assert db.shape == da.shape and torch.allclose(db.cpu(), da, atol=1e-4, rtol=1e-4)
    with pytest.raises(NotImplementedError):
        b.train()(scn.InputLayer(3, dims, mode=0)([c, f.cuda()]))


def test_dropin_under_its_import_name(tmp_path):
    """`import sparseconvnet` with <repo>/sgnn_b200/dropin on sys.path (INTEGRATION.md option 1) -- in a subprocess, because this
    test process already holds the ORACLE under that module name -- runs the module-parity net; outputs against oracle O2."""
    import subprocess
    import sys
    from conftest import ROOT
    import sparseconvnet as o2
    from helpers import random_coords
    rng = np.random.default_rng(5)
    dims = [16, 16, 16]
    c = random_coords(rng, 2, dims, 0.2)
    f = rng.standard_normal((c.shape[0], 4)).astype(np.float32)

    def net(lib):
        torch.manual_seed(0)
        s = lib.Sequential()
        s.add(lib.SubmanifoldConvolution(3, 4, 16, 3, False))
        s.add(lib.FullyConvolutionalNet(3, reps=1, nPlanes=[16, 16, 16], residual_blocks=True))
        s.add(lib.BatchNormReLU(48))
        return s
    a = net(o2).eval()
    torch.save(a.state_dict(), str(tmp_path / 'sd.pt'))
    np.savez(str(tmp_path / 'in.npz'), c=c, f=f)
    script = """
import sys, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
import sparseconvnet as scn
assert 'dropin' in scn.__file__, scn.__file__
d = np.load(%r)
torch.manual_seed(0)
s = scn.Sequential()
s.add(scn.SubmanifoldConvolution(3, 4, 16, 3, False))
s.add(scn.FullyConvolutionalNet(3, reps=1, nPlanes=[16, 16, 16], residual_blocks=True))
s.add(scn.BatchNormReLU(48))
s.load_state_dict(torch.load(%r))
s = s.cuda().eval()
with torch.no_grad():
    t = s(scn.InputLayer(3, [16, 16, 16], mode=0)([torch.from_numpy(d['c']), torch.from_numpy(d['f']).cuda()]))
    locs = t.metadata.getSpatialLocations(t.spatial_size)
    assert not locs.is_cuda and locs.dtype == torch.int64        # scn returns a CPU LongTensor (model.py:344-353 relies on it)
    np.savez(%r, out=scn.OutputLayer(3)(t).cpu().numpy(), locs=locs.numpy())
""" % (ROOT, os.path.join(ROOT, 'sgnn_b200', 'dropin'), str(tmp_path / 'in.npz'), str(tmp_path / 'sd.pt'), str(tmp_path / 'out.npz'))
    r = subprocess.run([sys.executable, '-c', script], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got = np.load(str(tmp_path / 'out.npz'))
    with torch.no_grad():
        ta = a(o2.InputLayer(3, dims, mode=0)([torch.from_numpy(c), torch.from_numpy(f)]))
    assert np.array_equal(got['locs'], ta.metadata.getSpatialLocations(ta.spatial_size).numpy())
    assert np.allclose(got['out'], o2.OutputLayer(3)(ta).numpy(), atol=1e-4, rtol=1e-4)


def test_native_generator_rejects_out_of_range_coordinates():
    from sgnn_b200.synth import synthetic_batch
    from sgnn_b200._lib import SgnnError
    m = _model((32, 32, 32), 0)
    locs, feats = synthetic_batch(1, [32, 32, 32], 0.08)
    bad = locs.clone()
    bad[5, 1] = 40                                   # y beyond the declared 32^3 input size
    with pytest.raises(SgnnError):
        m([bad.cuda(), feats.cuda()], ONES)
    out = m([locs.cuda(), feats.cuda()], ONES)       # the engine stays usable afterwards
    assert out[0][1].shape[0] > 0


def test_whole_scene_shape_batch1_update_sizes():
    """the test_scene.py call pattern -- batch 1, non-cubic scene padded to /32, update_sizes() before the
    forward (test_scene.py:77-82), coordinates left on the CPU -- against the live oracle generator."""
    from genmodel import OracleGenModel
    from sgnn_b200.synth import fill_parameters, synthetic_batch
    dims = (64, 96, 32)
    locs, feats = synthetic_batch(1, list(dims), 0.04, seed0=77)
    ora = OracleGenModel(input_dim=64)
    fill_parameters(ora, 4)
    ora.eval()
    ora.set_sizes(dims)
    with torch.no_grad():
        (wl, ws), wlv = ora(locs, feats)
    import sgnn_b200
    m = sgnn_b200.GenModel(8, 64, 1, 16, 16, 4, True, True, 1, 1)
    m.load_state_dict(ora.state_dict())
    m = m.cuda().eval()
    m.update_sizes(np.array(dims), np.array(dims) // 8)
    (gl, gs), glv = m([locs, feats.cuda()], ONES)
    for i, (w, gg) in enumerate(zip(wlv, glv)):
        assert torch.equal(w[0], gg[0].cpu()), 'level %d' % i
        assert (w[1] - gg[1].cpu()).abs().max() <= TOL_LOGIT
        flips = (torch.sigmoid(w[1][:, 0]) > 0.5) != (torch.sigmoid(gg[1][:, 0].cpu()) > 0.5)
        assert bool((w[1][:, 0][flips].abs() < 1e-5).all())
        if bool(flips.any()):
            return
    assert torch.equal(wl, gl.cpu()) and (ws - gs.cpu()).abs().max() <= TOL_SDF


def test_streaming_runner_matches_direct_calls():
    """Host-buffer pipeline (H2D / compute / D2H on three streams, double-buffered staging) returns, for every step,
    exactly what a direct model call on that batch returns -- including when slots are reused by differently sized
    batches."""
    from sgnn_b200.streaming import StreamingRunner
    from sgnn_b200.synth import synthetic_batch
    dims = (32, 32, 32)
    m = _model(dims, 0)
    batches = [synthetic_batch(nb, list(dims), occ, seed0=s) for nb, occ, s in
               [(2, 0.08, 1), (3, 0.05, 2), (1, 0.12, 3), (2, 0.03, 4), (3, 0.10, 5)]]
    nbs = [2, 3, 1, 2, 3]
    direct = []
    for (l, f), nb in zip(batches, nbs):
        (ol, os_), _ = m([l.cuda(), f.cuda(), nb], ONES)
        direct.append(([], []) if isinstance(ol, list) else (ol.cpu().clone(), os_.cpu().clone()))
    # int16 host coordinates in (8 B per voxel over PCIe), int16 result coordinates out; an int64 batch in between
    pinned = [((l if i == 2 else l.to(torch.int16)).pin_memory(), f.pin_memory()) for i, (l, f) in enumerate(batches)]
    r = StreamingRunner(m, depth=2)
    r.submit(pinned[0][0], pinned[0][1], nbs[0])
    r.submit(pinned[1][0], pinned[1][1], nbs[1])
    with pytest.raises(RuntimeError):                      # more than `depth` batches in flight would overwrite a staging slot
        r.submit(pinned[2][0], pinned[2][1], nbs[2])
    r = StreamingRunner(m, depth=2)
    r.submit(pinned[0][0], pinned[0][1], nbs[0])
    prev = None
    for i in range(len(batches)):
        if i + 1 < len(batches):
            r.submit(pinned[i + 1][0], pinned[i + 1][1], nbs[i + 1])
        t = r.step(ONES)
        if prev is not None:
            _check_host_result(r.result(prev[1]), direct[prev[0]])
        prev = (i, t)
    _check_host_result(r.result(prev[1]), direct[prev[0]])
    assert sum(0 if isinstance(d[0], list) else d[0].shape[0] for d in direct) > 0


def _check_host_result(got, want):
    if isinstance(want[0], list):
        assert isinstance(got[0], list)
    else:
        assert got[0].dtype == torch.int16 and torch.equal(got[0].long(), want[0]) and torch.equal(got[1], want[1])


def test_native_generator_follows_parameter_updates():
    """Weights changed after the first forward (load_state_dict, in-place update) are picked up by the next one: the cached
    weight struct is keyed by (version, address) of every parameter and buffer."""
    import sgnn_b200
    from sgnn_b200.synth import fill_parameters, synthetic_batch
    dims = (32, 32, 32)
    locs, feats = synthetic_batch(2, list(dims), 0.08)
    a = _model(dims, 0)
    out_a = a([locs.cuda(), feats.cuda()], ONES)
    b = _model(dims, 5)
    out_b = b([locs.cuda(), feats.cuda()], ONES)
    assert not torch.equal(out_a[1][0][1], out_b[1][0][1])
    a.load_state_dict(b.state_dict())
    _same_outputs(a([locs.cuda(), feats.cuda()], ONES), out_b)
    with torch.no_grad():
        for p in a.parameters():
            p.mul_(1.0)                                            # versions move, values do not
        a.encoder.occpred[0].weight.neg_()
        b.encoder.occpred[0].weight.neg_()
    _same_outputs(a([locs.cuda(), feats.cuda()], ONES), b([locs.cuda(), feats.cuda()], ONES))


@pytest.mark.parametrize('mode', ['exact', 'tc32'])
def test_baseline_config_teacher_forced_all_blocks(mode):
    """configs[1] at full size once more, with NOTHING excluded: the oracle runs with the device pass's keep decisions
    (helpers.compare_teacher_forced), so all 32 blocks are compared at every level to the end -- coordinates equal, logits within
    1e-4, the oracle's own decisions equal to the device's except within 1e-5 of the threshold (counted), TSDF <= 1e-3."""
    from genmodel import OracleGenModel
    from helpers import compare_teacher_forced
    from sgnn_b200.synth import fill_parameters, synthetic_batch
    locs, feats = synthetic_batch(32, 64, 0.05)
    ora = OracleGenModel()
    fill_parameters(ora, 0)
    ora.eval()
    m = _model((64, 64, 64), 0)
    m.conv_mode = mode
    got = m([locs.cuda(), feats.cuda(), 32], ONES)
    flips = compare_teacher_forced(ora, locs, feats, got, margin=1e-5, tol_logit=TOL_LOGIT, rtol_logit=0.0, tol_sdf=TOL_SDF,
                                   max_flips=4, tag='configs[1] teacher-forced ' + mode)
    assert sum(flips) <= 4
