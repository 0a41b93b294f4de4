"""The oracle generator (oracle/genmodel.py on O2) against (a) the UNMODIFIED reference model.py where the
reference tree exists, (b) the committed golden fixtures that model.py produced (tests/golden/make_golden.py)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, REFERENCE
from genmodel import OracleGenModel
from sgnn_b200.synth import fill_parameters, synthetic_batch


def _load(name):
    return np.load(os.path.join(GOLDEN, 'sgnn_ref_%s.npz' % name))


@pytest.mark.parametrize('name', ['b2_s32', 'ragged', 'b1_s64'])
def test_oracle_reproduces_golden(name):
    g = _load(name)
    m = OracleGenModel(input_dim=[int(v) for v in g['dims']])
    fill_parameters(m, int(g['param_seed']))
    m.eval()
    torch.set_num_threads(8)
    with torch.no_grad():
        out, levels = m(torch.from_numpy(g['in_locs'].astype(np.int64)), torch.from_numpy(g['in_feats']))
    assert np.array_equal(out[0].numpy(), g['out_locs'].astype(np.int64))
    assert np.abs(out[1].numpy() - g['out_sdf']).max() <= 1e-5
    for i, l in enumerate(levels):
        assert np.array_equal(l[0].numpy(), g['cand%d_locs' % i].astype(np.int64))
        assert np.abs(l[1].numpy() - g['cand%d' % i]).max() <= 1e-5


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason='reference tree only exists in the build container')
def test_oracle_equals_unmodified_reference_model():
    sys.path.insert(0, REFERENCE)
    with contextlib.redirect_stdout(io.StringIO()):
        import model as refmodel
        ref = refmodel.GenModel(8, 32, 1, 16, 16, 4, True, True, 1, 1)
    ora = OracleGenModel(input_dim=32)
    assert list(ref.state_dict().keys()) == list(ora.state_dict().keys())
    fill_parameters(ref, 0)
    ora.load_state_dict(ref.state_dict())
    ref.eval(), ora.eval()
    locs, feats = synthetic_batch(2, 32, 0.08)
    with torch.no_grad():
        a = ref([locs.clone(), feats.clone()], np.ones(5, dtype=np.float32))
        b = ora(locs, feats)
    assert torch.equal(a[0][0], b[0][0]) and torch.equal(a[0][1], b[0][1])
    for x, y in zip(a[1], b[1]):
        assert torch.equal(x[0], y[0]) and torch.equal(x[1], y[1])
