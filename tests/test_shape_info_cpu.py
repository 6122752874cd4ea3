"""Host-side pieces of the C++ shim that need no GPU: internal::MatrixFreeFunctions::ShapeInfo (the face tables of the
phase-space element, matrix_free/shape_info.h:62-231) against a direct restatement of the reference's construction, and
MatrixFree::get_faces_by_cells_boundary_id's lattice logic (internal::face_boundary_id)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r'''
#include <cstdio>
#include "hyperdeal_b200.hpp"
using namespace hyperdeal;
template <int dx, int dv>
void dump(unsigned degree)
{
  internal::MatrixFreeFunctions::ShapeInfo<double> s;
  s.template reinit<dx, dv>(degree);
  std::printf("S %d %d %u %u %u %zu\n", dx, dv, degree, s.dofs_per_cell, s.dofs_per_face, s.face_orientations.size());
  for (auto &row : s.face_to_cell_index_nodal) { for (auto v : row) std::printf("%u ", v); std::printf("\n"); }
  for (auto &row : s.face_orientations) { for (auto v : row) std::printf("%u ", v); std::printf("\n"); }
}
int main()
{
  dump<1, 1>(3); dump<2, 2>(2); dump<3, 3>(1); dump<3, 2>(2); dump<2, 3>(1); dump<3, 3>(3);
  // face_boundary_id: 2D2V lattice 3x2x2x4 cells, x periodic, v Dirichlet; brick = the whole lattice
  hd_mesh_desc d{};
  d.dim_x = 2; d.dim_v = 2;
  const int nc[4] = {3, 2, 2, 4};
  for (int i = 0; i < 4; ++i) { d.n_cells[i] = d.n_cells_global[i] = nc[i]; d.cell_offset[i] = 0; d.side_kind[i][0] = d.side_kind[i][1] = i < 2 ? HD_SIDE_PERIODIC_LOCAL : HD_SIDE_DIRICHLET; }
  for (long long c = 0; c < 48; ++c) { std::printf("B"); for (unsigned f = 0; f < 8; ++f) std::printf(" %d", int(internal::face_boundary_id(d, c, f))); std::printf("\n"); }
  // 1D1V, both Dirichlet: end points are numbered 0 and 1
  hd_mesh_desc e{};
  e.dim_x = 1; e.dim_v = 1;
  for (int i = 0; i < 2; ++i) { e.n_cells[i] = e.n_cells_global[i] = 3; e.cell_offset[i] = 0; e.side_kind[i][0] = e.side_kind[i][1] = HD_SIDE_DIRICHLET; }
  for (long long c = 0; c < 9; ++c) { std::printf("E"); for (unsigned f = 0; f < 4; ++f) std::printf(" %d", int(internal::face_boundary_id(e, c, f))); std::printf("\n"); }
  // multi::brick_grid: how PartitionX / PartitionV ranks cut a Cartesian lattice (slowest direction first)
  {
    CartesianLattice<3> l3; l3.n_cells = {{8, 8, 8}};
    for (int parts : {1, 2, 4, 8, 16, 64, 512}) { auto g = multi::brick_grid<3>(l3, parts); std::printf("G 3 %d %d %d %d\n", parts, g[0], g[1], g[2]); }
    CartesianLattice<2> l2; l2.n_cells = {{4, 6}};
    for (int parts : {1, 2, 3, 4, 6, 12, 24}) { auto g = multi::brick_grid<2>(l2, parts); std::printf("G 2 %d %d %d\n", parts, g[0], g[1]); }
    CartesianLattice<1> l1; l1.n_cells = {{6}};
    for (int parts : {1, 2, 3, 6}) { auto g = multi::brick_grid<1>(l1, parts); std::printf("G 1 %d %d\n", parts, g[0]); }
    bool threw = false;
    try { multi::brick_grid<2>(l2, 5); } catch (const ExcMessage &) { threw = true; }
    std::printf("G throw %d\n", int(threw));
  }
  return 0;
}
'''


def _sub_table(dim, points):
    """fill_face_to_cell_index_nodal (shape_info.h:62-106)"""
    nf = points ** (dim - 1)
    out = [[0] * nf for _ in range(2 * dim)]
    for f in range(2 * dim):
        direction = f // 2
        stride = points if direction < dim - 1 else 1
        shift = points ** direction
        offset = (f % 2) * (points - 1) * shift
        if direction == 0 or direction == dim - 1:
            for i in range(nf):
                out[f][i] = offset + i * stride
        else:
            for j in range(points):
                for i in range(points):
                    out[f][i * points + j] = offset + j * nf + i
    return out


def _expected(dx, dv, degree):
    """ShapeInfo::reinit (shape_info.h:110-160)"""
    n, dim = degree + 1, dx + dv
    tx, tv = _sub_table(dx, n), _sub_table(dv, n)
    rows = []
    for s in range(2 * dim):
        row = []
        if s < 2 * dx:
            for i in range(n ** dv):
                for j in range(n ** (dx - 1)):
                    row.append(tx[s][j] + n ** dx * i)
        else:
            for i in range(n ** (dv - 1)):
                for j in range(n ** dx):
                    row.append(j + n ** dx * tv[s - 2 * dx][i])
        rows.append(row)
    return rows


def _orientation_tables(dx, dv, degree):
    """shape_info.h:163-222"""
    n, dim = degree + 1, dx + dv
    nf = n ** (dim - 1)
    if dx != 3 and dv != 3:
        return []
    t = [[0] * nf for _ in range(16)]
    form = [lambda j, k: k + j * n, lambda j, k: j + k * n, lambda j, k: (n - 1 - k) + (n - 1 - j) * n, lambda j, k: (n - 1 - j) + (n - 1 - k) * n,
            lambda j, k: j + (n - 1 - k) * n, lambda j, k: k + (n - 1 - j) * n, lambda j, k: (n - 1 - j) + k * n, lambda j, k: (n - 1 - k) + j * n]
    if dx == 3:
        c = 0
        for i in range(n ** dv):
            for j in range(n):
                for k in range(n):
                    for o in range(8):
                        t[o][c] = form[o](j, k) + i * n * n
                    c += 1
    else:
        for o in range(8):
            t[o] = list(range(nf))
    if dv == 3:
        c = 0
        for j in range(n):
            for k in range(n):
                for i in range(n ** dx):
                    for o in range(8):
                        t[8 + o][c] = form[o](j, k) * n ** dx + i
                    c += 1
    else:
        for o in range(8):
            t[8 + o] = list(range(nf))
    return t


@pytest.fixture(scope="module")
def dump(tmp_path_factory):
    d = tmp_path_factory.mktemp("shape")
    src = d / "t.cc"
    src.write_text(PROGRAM)
    exe = str(d / "t")
    # (the program calls no library function: the header's inline C-ABI calls are never instantiated)
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "hyperdeal_b200", "cpp"), str(src), "-o", exe,
                        "-Wl,--unresolved-symbols=ignore-all"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return out.stdout.splitlines()


def test_shape_info_matches_the_reference_construction(dump):
    i = 0
    seen = 0
    while i < len(dump) and dump[i].startswith("S"):
        _, dx, dv, degree, per_cell, per_face, n_or = dump[i].split()
        dx, dv, degree, n_or = int(dx), int(dv), int(degree), int(n_or)
        n, dim = degree + 1, dx + dv
        assert int(per_cell) == n ** dim and int(per_face) == n ** (dim - 1)
        rows = [[int(x) for x in dump[i + 1 + f].split()] for f in range(2 * dim)]
        assert rows == _expected(dx, dv, degree), (dx, dv, degree)
        for f, row in enumerate(rows):  # every entry lies on its face, no entry twice
            stride = n ** (f // 2)
            assert len(set(row)) == len(row) and all((v // stride) % n == (f % 2) * (n - 1) for v in row)
        ori = [[int(x) for x in dump[i + 1 + 2 * dim + o].split()] for o in range(n_or)]
        assert ori == _orientation_tables(dx, dv, degree), (dx, dv, degree)
        i += 1 + 2 * dim + n_or
        seen += 1
    assert seen == 6


def test_face_boundary_ids(dump):
    rows = [[int(x) for x in l.split()[1:]] for l in dump if l.startswith("B")]
    assert len(rows) == 48
    internal = 2 ** 32 - 1  # numbers::internal_face_boundary_id as unsigned, printed as int -> -1
    for cell, r in enumerate(rows):
        c = [cell % 3, (cell // 3) % 2, (cell // 6) % 2, cell // 12]
        for f in range(8):
            d, side = f // 2, f % 2
            nc = (3, 2, 2, 4)[d]
            outer = c[d] == (nc - 1 if side else 0)
            expect = 0 if (d >= 2 and outer) else -1
            assert r[f] == expect, (cell, f)
    rows = [[int(x) for x in l.split()[1:]] for l in dump if l.startswith("E")]
    for cell, r in enumerate(rows):
        c = [cell % 3, cell // 3]
        for f in range(4):
            d, side = f // 2, f % 2
            outer = c[d] == (2 if side else 0)
            assert r[f] == (side if outer else -1)


def test_brick_grid_of_the_multi_gpu_shim(dump):
    """hyperdeal::multi::brick_grid: the product of the grid is the number of ranks, every extent divides its direction, the
    slowest direction is cut first (rows of cells along x_0 stay whole as long as possible), impossible requests throw"""
    cells = {3: (8, 8, 8), 2: (4, 6), 1: (6,)}
    seen = 0
    for l in dump:
        if not l.startswith("G ") or l.startswith("G throw"):
            continue
        f = [int(x) for x in l.split()[1:]]
        dim, parts, g = f[0], f[1], f[2:]
        assert len(g) == dim and np_prod(g) == parts
        assert all(c % e == 0 for c, e in zip(cells[dim], g))
        if parts > 1:
            assert g[-1] > 1  # slowest direction first
        if dim == 3 and parts <= 8:
            assert g[0] == 1  # x_0 is cut last
        seen += 1
    assert seen == 7 + 7 + 4
    assert "G throw 1" in dump


def np_prod(v):
    r = 1
    for x in v:
        r *= x
    return r
