"""CPU checks of the arithmetic identities the tensor-core convolution path (csrc/conv_tc32.cu) is built on -- no GPU:
  * the exact 3-way bf16 split of an fp32 value (v = v0 + v1 + v2, each piece representable in bf16) and the size of the
    three dropped partial products;
  * the child-mode restructuring (generative upsampling, model.py:192-207,224-225): for child c and filter offset d the
    neighbour lives in parent offset e = floor((c+d)/2) per axis, so the 27-tap convolution over the 8 children of every
    site equals an 8-tap convolution over PARENT rows with filters pre-summed per (child, parent offset) -- checked
    against oracle O3's child-mode convolution;
  * the enumeration of the 64 (parent offset, child) pairs and the round structure of the warp-specialised kernel."""
import numpy as np
import torch

import o3
from helpers import random_coords, nbr_table


def _split3(v):
    u = v.view(np.uint32)
    h = (u & np.uint32(0xffff0000)).view(np.float32)
    r = v - h
    m = (r.view(np.uint32) & np.uint32(0xffff0000)).view(np.float32)
    lo = r - m
    return h, m, lo


def child_parent_offset(c, d):          # conv_src_row() of csrc/conv.cu
    dz, dy, dx = d // 9 - 1, (d // 3) % 3 - 1, d % 3 - 1
    pz = (((c >> 2) & 1) + dz + 2) // 2 - 1
    py = (((c >> 1) & 1) + dy + 2) // 2 - 1
    px = ((c & 1) + dx + 2) // 2 - 1
    return (pz + 1) * 9 + (py + 1) * 3 + (px + 1)


def child_uses(c, e):                   # child_uses() of csrc/conv_tc32.cu
    ez, ey, ex = e // 9 - 1, (e // 3) % 3 - 1, e % 3 - 1
    cz, cy, cx = (c >> 2) & 1, (c >> 1) & 1, c & 1
    return (ez == 0 or ez == 2 * cz - 1) and (ey == 0 or ey == 2 * cy - 1) and (ex == 0 or ex == 2 * cx - 1)


def test_three_way_bf16_split_is_exact():
    rng = np.random.default_rng(0)
    v = (rng.standard_normal(200000) * np.exp(rng.uniform(-30, 30, 200000))).astype(np.float32)
    v = np.concatenate([v, np.float32([0.0, -0.0, 1.0, -1.0, 3.0e38, 1.0e-30, 1.0 + 2.0 ** -23])])   # (pieces of |v| < 2^-110 go subnormal: the kernel truncates them)
    h, m, lo = _split3(v)
    for piece in (h, m, lo):            # every piece is a bf16 value: its low 16 bits are zero
        assert not np.any(piece.view(np.uint32) & np.uint32(0xffff))
    assert np.array_equal(h.astype(np.float64) + m.astype(np.float64) + lo.astype(np.float64), v.astype(np.float64))
    nz = v != 0
    assert np.all(np.abs(m[nz]) <= np.abs(v[nz]) * 2.0 ** -7) and np.all(np.abs(lo[nz]) <= np.abs(v[nz]) * 2.0 ** -15)


def test_six_partial_products_reach_fp32_accuracy():
    rng = np.random.default_rng(1)
    x = rng.standard_normal(100000).astype(np.float32)
    w = rng.standard_normal(100000).astype(np.float32)
    xs, ws = _split3(x), _split3(w)
    kept = sum(xs[i].astype(np.float64) * ws[j].astype(np.float64) for i in range(3) for j in range(3) if i + j <= 2)
    exact = x.astype(np.float64) * w.astype(np.float64)
    assert np.all(np.abs(kept - exact) <= 2.0 ** -21 * np.abs(exact))       # three dropped terms: < 3 * 2^-23 |x w|
    for i in range(3):                  # every kept product is exact in fp32 (16 significant bits)
        for j in range(3 - i):
            p64 = xs[i].astype(np.float64) * ws[j].astype(np.float64)
            assert np.array_equal(p64.astype(np.float32).astype(np.float64), p64)


def test_child_offsets_collapse_onto_eight_parent_neighbours():
    pairs = [(e, c) for e in range(27) for c in range(8) if child_uses(c, e)]
    assert len(pairs) == 64
    for c in range(8):
        reached = {child_parent_offset(c, d) for d in range(27)}
        assert reached == {e for e in range(27) if child_uses(c, e)} and len(reached) == 8
    # rounds of the warp-specialised kernel: <= 4 children per round, the centre offset (all 8 children) takes two
    per_e = [sum(child_uses(c, e) for c in range(8)) for e in range(27)]
    assert per_e[13] == 8 and sorted(set(per_e)) == [1, 2, 4, 8] and sum(1 for n in per_e if n > 4) == 1


def test_presummed_child_filters_equal_the_27_tap_child_convolution():
    rng = np.random.default_rng(2)
    c = random_coords(rng, 2, (7, 6, 9), 0.4)
    n = c.shape[0]
    x = rng.standard_normal((n, 48)).astype(np.float32)
    w = (rng.standard_normal((27, 48, 16)) * 0.05).astype(np.float32)
    nbr = nbr_table(c)
    want = o3.conv(torch.from_numpy(x), torch.from_numpy(nbr), torch.from_numpy(w), 8 * n, child_mode=True).numpy()
    out = np.zeros((8 * n, 16))
    for e in range(27):
        rows = nbr[e]
        xe = np.where(rows[:, None] >= 0, x[np.maximum(rows, 0)], 0).astype(np.float64)
        for ch in range(8):
            if not child_uses(ch, e):
                continue
            wsum = sum(w[d].astype(np.float64) for d in range(27) if child_parent_offset(ch, d) == e)
            out[np.arange(n) * 8 + ch] += xe @ wsum.astype(np.float32).astype(np.float64)   # rounded once, as the prep kernel does
    bound = o3.conv(torch.from_numpy(np.abs(x)), torch.from_numpy(nbr), torch.from_numpy(np.abs(w)), 8 * n,
                    child_mode=True).numpy()
    assert np.all(np.abs(out - want) <= 4e-6 * bound + 1e-7)
