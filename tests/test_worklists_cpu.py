"""Interior / boundary work lists of the pipelined 3D3V kernel (kernel_fast6d.cu: FastParams::cutg/bsize, decode_interior,
decode_boundary): a Python model of the same mixed-radix enumeration must visit every row of the brick exactly once —
interior rows (no cut direction at its ghost layer) in the first list, all others in the second — for any set of cut
directions, ghost sides and extents.  (The device code itself is exercised by the -m gpu parity tests.)"""
import itertools

import pytest


def lists(ncell, cutg):
    """ncell[1..5], cutg[d] = ghost-layer coordinate or -1 -> (interior rows, boundary rows) as coordinate tuples"""
    dirs = range(1, 6)
    n_int = 1
    for d in dirs:
        n_int *= ncell[d] - (1 if cutg[d] >= 0 else 0)
    bsize = {d: 0 for d in dirs}
    for k in dirs:
        if cutg[k] >= 0:
            sz = 1
            for d in dirs:
                if d != k:
                    sz *= ncell[d] - (1 if (cutg[d] >= 0 and d < k) else 0)
            bsize[k] = sz

    def decode_interior(i):
        c = {}
        for d in dirs:
            cut = cutg[d] >= 0
            r = ncell[d] - (1 if cut else 0)
            q, i = i % r, i // r
            c[d] = q + 1 if (cut and cutg[d] == 0) else q
        return tuple(c[d] for d in dirs)

    def decode_boundary(i):
        dk = 0
        for d in dirs:
            if dk == 0:
                if i < bsize[d]:
                    dk = d
                else:
                    i -= bsize[d]
        c = {}
        for d in dirs:
            cut = cutg[d] >= 0
            if d == dk:
                c[d] = cutg[d]
            else:
                skip = cut and d < dk
                r = ncell[d] - (1 if skip else 0)
                q, i = i % r, i // r
                c[d] = q + 1 if (skip and cutg[d] == 0) else q
        return tuple(c[d] for d in dirs)

    return [decode_interior(i) for i in range(n_int)], [decode_boundary(i) for i in range(sum(bsize.values()))]


CASES = [
    ((8, 8, 8, 2, 2), (-1, 7, -1, -1, -1)),
    ((4, 4, 3, 2, 2), (0, 3, -1, -1, -1)),
    ((4, 4, 3, 2, 2), (3, 0, 0, -1, 1)),
    ((2, 1, 3, 2, 2), (0, 0, -1, -1, -1)),   # a cut direction with a single cell: no interior rows at all
    ((3, 3, 3, 3, 3), (-1, -1, -1, -1, -1)),  # nothing cut: every row is interior
    ((2, 2, 2, 2, 2), (0, 1, 0, 1, 0)),
]


@pytest.mark.parametrize("ncell,cutg", CASES)
def test_lists_partition_the_rows(ncell, cutg):
    nc = dict(zip(range(1, 6), ncell))
    cg = dict(zip(range(1, 6), cutg))
    interior, boundary = lists(nc, cg)
    every = set(itertools.product(*[range(n) for n in ncell]))
    assert len(set(interior)) == len(interior) and len(set(boundary)) == len(boundary)
    assert set(interior) | set(boundary) == every and not (set(interior) & set(boundary))
    for row in interior:
        assert all(cutg[j] < 0 or row[j] != cutg[j] for j in range(5))
    for row in boundary:
        assert any(cutg[j] >= 0 and row[j] == cutg[j] for j in range(5))
