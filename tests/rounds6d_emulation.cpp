// TEST HARNESS (tests/test_rounds6d_emulation.py) — not product code, never linked into libhdgpu.so.
// Runs the task bodies of the three-round 3D3V kernel (hyperdeal_b200/csrc/rounds6d_tasks.cuh) on the host, thread by
// thread and round by round, with an emulated shared memory that is filled the way the TMA fills it (128-byte swizzle).
// Rows of cells are walked in upwind order like the kernel's producer does, so the register-carried direction-0 trace,
// the trace addressing in `src` and in a ghost buffer, and the shared-memory maps are all covered.
#define HD_R6_HOST_EMULATION
#include "../hyperdeal_b200/csrc/rounds6d_tasks.cuh"

#include <vector>

#include "../hyperdeal_b200/csrc/basis.hpp"

namespace r6emu
{
  unsigned char *smem = nullptr;
}

namespace
{
  struct Lat
  {
    int       ncell[6], up_delta[6], up_kind[6];
    long long ghost_off[6];
  };
} // namespace

// ghost_mask bit d: the upwind side of direction d is treated as a GHOST side; the harness fills the ghost buffer from the
// periodic neighbour itself (a self-exchange), so the result must equal the periodic one.
extern "C" int
hd_r6_emulate(const double *src, double *dst, const int *ncell, const double *left, const double *right, const double *velocity, double skew,
              int ghost_mask)
{
  try
    {
      using namespace r6;
      hd::Basis1D bs;
      bs.init(3, 4, false);
      bs.set_skew((hd::LD)skew);
      Coef cf;
      Lat  L;
      long long ncells = 1;
      for (int d = 0; d < 6; ++d)
        {
          L.ncell[d] = ncell[d];
          ncells *= ncell[d];
          std::vector<hd::LD> C[4], L0, L1;
          bs.direction_matrices((hd::LD)velocity[d], (hd::LD)((right[d] - left[d]) / ncell[d]), (hd::LD)skew, C, L0, L1);
          bool lo = false, hi = false;
          for (int i = 0; i < 4; ++i)
            {
              lo |= L0[i] != 0;
              hi |= L1[i] != 0;
            }
          L.up_delta[d] = lo ? -1 : (hi ? +1 : 0);
          L.up_kind[d]  = ((ghost_mask >> d) & 1) ? 1 : 0;
          double *Cd = (d & 1) ? cf.B[d / 2] : cf.A[d / 2];
          double *Ld = (d & 1) ? cf.LB[d / 2] : cf.LA[d / 2];
          for (int i = 0; i < 16; ++i)
            Cd[i] = (double)C[0][i];
          for (int i = 0; i < 4; ++i)
            Ld[i] = lo ? (double)L0[i] : (hi ? (double)L1[i] : 0.0);
        }
      // ghost buffer: segments ordered by direction, face cells lexicographic over the other directions, 1024 values each
      std::vector<double> ghost;
      long long           goff = 0;
      for (int d = 0; d < 6; ++d)
        {
          L.ghost_off[d] = goff;
          if (!(L.up_kind[d] == 1 && L.up_delta[d] != 0))
            continue;
          const long long nfc = ncells / ncell[d];
          ghost.resize(goff + nfc * 1024);
          for (long long fc = 0; fc < nfc; ++fc)
            {
              // the periodic neighbour behind the upwind side: the cell at the opposite end
              long long r = fc, cell = 0, m = 1;
              for (int e = 0; e < 6; ++e)
                {
                  int ce;
                  if (e == d)
                    ce = L.up_delta[d] < 0 ? ncell[d] - 1 : 0;
                  else
                    {
                      ce = int(r % ncell[e]);
                      r /= ncell[e];
                    }
                  cell += ce * m;
                  m *= ncell[e];
                }
              const int stride = 1 << (2 * d), layer = L.up_delta[d] < 0 ? 3 : 0;
              for (int i = 0; i < 1024; ++i)
                {
                  const int hi = i / stride, lo = i % stride;
                  ghost[goff + fc * 1024 + i] = src[cell * CELL + hi * 4 * stride + layer * stride + lo];
                }
            }
          goff += nfc * 1024;
        }
      std::vector<unsigned char> sm(2 * U_BYTES);
      r6emu::smem       = sm.data();
      const uint32_t ub = 0, pb = U_BYTES;
      ThreadMap<0>   tm0[128];
      ThreadMap<1>   tm1[128];
      ThreadMap<2>   tm2[128];
      for (int t = 0; t < 128; ++t)
        {
          tm0[t].init(t);
          tm1[t].init(t);
          tm2[t].init(t);
        }
      const bool descend = L.up_delta[0] > 0;
      const long long nrows = ncells / ncell[0];
      std::vector<double> tr(128 * 2 * 4);
      for (long long row = 0; row < nrows; ++row)
        {
          int       c[6];
          long long r = row;
          for (int d = 1; d < 6; ++d)
            {
              c[d] = int(r % ncell[d]);
              r /= ncell[d];
            }
          for (int step = 0; step < ncell[0]; ++step)
            {
              c[0]           = descend ? ncell[0] - 1 - step : step;
              long long cell = 0;
              for (int d = 5; d >= 0; --d)
                cell = cell * ncell[d] + c[d];
              FaceBase fbv[6];
              for (int d = 0; d < 6; ++d)
                fbv[d] = face_base(L, c, d);
              const bool first = step == 0;
              // TMA fill: row rr (16 doubles) -> rr * 128, chunk ch -> ch ^ (rr & 7)
              for (int rr = 0; rr < 256; ++rr)
                for (int ch = 0; ch < 8; ++ch)
                  std::memcpy(sm.data() + ub + rr * 128 + ((ch ^ (rr & 7)) << 4), src + cell * CELL + rr * 16 + ch * 2, 16);
              const double zero[4] = {0, 0, 0, 0};
              for (int t = 0; t < 128; ++t)
                for (int j = 0; j < 2; ++j)
                  {
                    double fa[4] = {0, 0, 0, 0}, fb[4] = {0, 0, 0, 0}, edge[4];
                    if (L.up_delta[0] != 0)
                      {
                        if (first)
                          load_trace<0, 0>(src, ghost.data(), fbv[0].off, fbv[0].ghost, t, j, fa);
                        else
                          for (int b = 0; b < 4; ++b)
                            fa[b] = tr[(t * 2 + j) * 4 + b];
                      }
                    if (L.up_delta[1] != 0)
                      load_trace<0, 1>(src, ghost.data(), fbv[1].off, fbv[1].ghost, t, j, fb);
                    task_round0(cf, ub, pb, tm0[t], j, fa, fb, descend, edge);
                    for (int b = 0; b < 4; ++b)
                      tr[(t * 2 + j) * 4 + b] = edge[b];
                  }
              (void)zero;
              for (int t = 0; t < 128; ++t)
                for (int j = 0; j < 2; ++j)
                  {
                    double fa[4] = {0, 0, 0, 0}, fb[4] = {0, 0, 0, 0};
                    if (L.up_delta[2] != 0)
                      load_trace<1, 0>(src, ghost.data(), fbv[2].off, fbv[2].ghost, t, j, fa);
                    if (L.up_delta[3] != 0)
                      load_trace<1, 1>(src, ghost.data(), fbv[3].off, fbv[3].ghost, t, j, fb);
                    task_round1(cf, ub, pb, tm1[t], j, fa, fb);
                  }
              for (int t = 0; t < 128; ++t)
                for (int j = 0; j < 2; ++j)
                  {
                    double fa[4] = {0, 0, 0, 0}, fb[4] = {0, 0, 0, 0}, q[4][4];
                    if (L.up_delta[4] != 0)
                      load_trace<2, 0>(src, ghost.data(), fbv[4].off, fbv[4].ghost, t, j, fa);
                    if (L.up_delta[5] != 0)
                      load_trace<2, 1>(src, ghost.data(), fbv[5].off, fbv[5].ghost, t, j, fb);
                    task_round2(cf, ub, pb, tm2[t], j, fa, fb, q);
                    const long long g0 = cell * CELL + (t & 15) + 16 * ((t >> 4) + 8 * j);
                    for (int b = 0; b < 4; ++b)
                      for (int a = 0; a < 4; ++a)
                        dst[g0 + 256 * a + 1024 * b] = q[b][a];
                  }
            }
        }
      return 0;
    }
  catch (const std::exception &)
    {
      return -1;
    }
}
