"""Host-side mirror of the reference interface: state_dict layout, update_sizes, synthetic inputs, sharding
(world_size-2 gloo), all on CPU."""
import json
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN
import sgnn_b200
from sgnn_b200 import shard
from sgnn_b200.synth import fill_parameters, synthetic_batch, synthetic_block


def test_state_dict_layout_equals_reference():
    layout = json.load(open(os.path.join(GOLDEN, 'state_dict_layout.json')))
    m = sgnn_b200.GenModel(8, 64, 1, 16, 16, 4, True, True, 1, 1)
    mine = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    assert mine == layout
    assert sum(p.numel() for p in m.parameters()) == 643735


def test_oracle_and_product_share_state_dict():
    from genmodel import OracleGenModel
    o = OracleGenModel()
    m = sgnn_b200.GenModel(8, 64, 1, 16, 16, 4, True, True, 1, 1)
    fill_parameters(o, 5)
    m.load_state_dict(o.state_dict())
    for (ka, a), (kb, b) in zip(o.state_dict().items(), m.state_dict().items()):
        assert ka == kb and torch.equal(a, b)


def test_update_sizes_mirrors_reference_bounds():
    m = sgnn_b200.GenModel(8, 64, 1, 16, 16, 4, True, True, 1, 1)
    m.update_sizes(np.array([128, 96, 64]), np.array([128, 96, 64]) // 8)
    assert m.encoder.process_sparse[0].p0.spatial_size.tolist() == [128, 96, 64]
    # model.py:363-369 doubles the refine dims inside the k loop: bounds only grow
    assert m.refinement[0].p0.spatial_size.tolist() == [16, 12 * 8, 8 * 64]
    assert m.surfacepred.p0.spatial_size.tolist() == [128, 96 * 8, 64 * 64]


def test_inference_only_and_no_cpu_path():
    m = sgnn_b200.GenModel(8, 64, 1, 16, 16, 4, True, True, 1, 1).eval()
    locs, feats = synthetic_batch(1, 16, 0.1)
    with pytest.raises(RuntimeError):
        m([locs, feats], np.ones(5))             # CPU features: fail loudly, never fall back
    m.train()
    with pytest.raises((NotImplementedError, RuntimeError)):
        m([locs, feats], np.ones(5))


def test_synthetic_inputs_are_deterministic():
    c, f = synthetic_block(0)
    assert c.shape[0] == 13166 or abs(c.shape[0] - 13107) < 600      # ~5 % of 64^3
    c2, f2 = synthetic_block(0)
    assert np.array_equal(c, c2) and np.array_equal(f, f2)
    assert np.all(np.diff(((c[:, 0] * 64 + c[:, 1]) * 64 + c[:, 2])) > 0)     # raster order, unique
    assert np.abs(f).max() < 3.0
    m1 = fill_parameters(torch.nn.Linear(7, 3), 1).weight.clone()
    m2 = fill_parameters(torch.nn.Linear(7, 3), 1).weight.clone()
    assert torch.equal(m1, m2) and m1.abs().max() > 0


def test_round_robin_shards_partition_the_blocks():
    for world in (1, 2, 4, 8):
        got = sorted(sum([shard.shard_blocks(4096, r, world) for r in range(world)], []))
        assert got == list(range(4096))
        assert shard.shard_blocks(10, 1, 4) == [1, 5, 9]


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(rank)                       # ranks start with DIFFERENT weights
    m = torch.nn.Sequential(torch.nn.Linear(8, 4), torch.nn.BatchNorm1d(4))
    shard.broadcast_parameters(m, src=0)
    _, flat = shard.flatten_state(m)
    mx = shard.max_over_ranks(10.0 + rank, 'cpu')
    sm = shard.sum_over_ranks(len(shard.shard_blocks(7, rank, world)), 'cpu')
    q.put((rank, flat.double().sum().item(), mx, sm))
    dist.destroy_process_group()


def test_gloo_world2_broadcast_and_reductions():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert res[0][1] == res[1][1]                 # identical parameters after the one broadcast
    assert res[0][2] == res[1][2] == 11.0         # max over ranks
    assert res[0][3] == res[1][3] == 7.0          # every block owned exactly once


def test_native_weight_cache_key_sees_updates_and_moves():
    """The native generator caches its weight struct and checks a (version, address) key over a CACHED tensor list on every
    forward (walking the module tree or building a state_dict per call cost more host time than the GPU pass had queued).
    The key must change on in-place updates, load_state_dict and dtype / device style re-allocation; invalidate_native()
    forgets everything."""
    import itertools
    import sgnn_b200
    from sgnn_b200 import native
    from sgnn_b200.synth import fill_parameters
    m = sgnn_b200.GenModel(8, 32, 1, 16, 16, 4, True, True, 1, 1).eval()
    fill_parameters(m, 0)
    tensors = list(itertools.chain(m.parameters(), m.buffers()))
    k0 = native._Weights.version_key(m, tensors)
    assert k0 == native._Weights.version_key(m, tensors) == native._Weights.version_key(m)
    with torch.no_grad():
        m.surfacepred.linear.weight.mul_(2.0)                     # optimizer-style in-place update
    k1 = native._Weights.version_key(m, tensors)
    assert k1 != k0
    other = sgnn_b200.GenModel(8, 32, 1, 16, 16, 4, True, True, 1, 1)
    fill_parameters(other, 7)
    m.load_state_dict(other.state_dict())                         # copies in place: versions move
    k2 = native._Weights.version_key(m, tensors)
    assert k2 != k1
    m.double()                                                    # re-allocates: addresses move (same Parameter objects)
    assert native._Weights.version_key(m, tensors) != k2
    m._native_ok = True
    m.invalidate_native()
    assert m._native is None and m._native_ok is None
    assert native.supported(m) is False                           # CPU / fp64 parameters: not a native-generator model
