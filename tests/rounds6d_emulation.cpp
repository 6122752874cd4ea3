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
      // the walk of ONE persistent CTA over the whole lattice: rows in lattice order, cells of a row in upwind order
      struct Item
      {
        long long cell;
        bool      first;
        FaceBase  fbv[6];
      };
      std::vector<Item> walk;
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
              c[0] = descend ? ncell[0] - 1 - step : step;
              Item it;
              it.cell = 0;
              for (int d = 5; d >= 0; --d)
                it.cell = it.cell * ncell[d] + c[d];
              it.first = step == 0;
              for (int d = 0; d < 6; ++d)
                it.fbv[d] = face_base(L, c, d);
              walk.push_back(it);
            }
        }
      // per-thread pipeline state of the three rounds, exactly as in r6_compute: (fa, fb) = traces of the task about to run,
      // requested by the previous task's after_traces() callback; round 0: eo = end layer the task after the next needs
      struct TS
      {
        double fa[4], fb[4], eo[4], U[4][4];
      };
      std::vector<TS> st[3];
      for (int R = 0; R < 3; ++R)
        st[R].assign(128, TS{});
      auto request = [&](int R, const Item &it, int t, int j, double(&fa)[4], double(&fb)[4]) {
        const FaceBase &A = it.fbv[2 * R], &B = it.fbv[2 * R + 1];
        if (R == 0)
          {
            if (L.up_delta[0] != 0 && it.first)
              load_trace<0, 0>(src, ghost.data(), A.off, A.ghost, t, j, fa);
            if (L.up_delta[1] != 0)
              load_trace<0, 1>(src, ghost.data(), B.off, B.ghost, t, j, fb);
          }
        else if (R == 1)
          {
            if (L.up_delta[2] != 0)
              load_trace<1, 0>(src, ghost.data(), A.off, A.ghost, t, j, fa);
            if (L.up_delta[3] != 0)
              load_trace<1, 1>(src, ghost.data(), B.off, B.ghost, t, j, fb);
          }
        else
          {
            if (L.up_delta[4] != 0)
              load_trace<2, 0>(src, ghost.data(), A.off, A.ghost, t, j, fa);
            if (L.up_delta[5] != 0)
              load_trace<2, 1>(src, ghost.data(), B.off, B.ghost, t, j, fb);
          }
      };
      for (int R = 0; R < 3; ++R)
        for (int t = 0; t < 128; ++t)
          request(R, walk[0], t, 0, st[R][t].fa, st[R][t].fb);
      // the emulated shared memory holds TWO cell stages (the current cell and the next one: a task loads the u tile of the
      // task after it, which may belong to the next cell) and one P buffer
      std::vector<unsigned char> sm2(3 * U_BYTES);
      r6emu::smem = sm2.data();
      auto fill = [&](size_t k) {
        // TMA fill: row rr (16 doubles) -> rr * 128, chunk ch -> ch ^ (rr & 7)
        const uint32_t st_off = uint32_t(k & 1) * U_BYTES;
        for (int rr = 0; rr < 256; ++rr)
          for (int ch = 0; ch < 8; ++ch)
            std::memcpy(sm2.data() + st_off + rr * 128 + ((ch ^ (rr & 7)) << 4), src + walk[k].cell * CELL + rr * 16 + ch * 2, 16);
      };
      const uint32_t pbuf = 2 * U_BYTES;
      fill(0);
      for (int t = 0; t < 128; ++t)
        {
          load_u<0>(0, tm0[t], 0, st[0][t].U);
          load_u<1>(0, tm1[t], 0, st[1][t].U);
          load_u<2>(0, tm2[t], 0, st[2][t].U);
        }
      for (size_t k = 0; k < walk.size(); ++k)
        {
          const Item &cur = walk[k];
          const uint32_t ucur = uint32_t(k & 1) * U_BYTES, unext = uint32_t((k + 1) & 1) * U_BYTES;
          if (k + 1 < walk.size())
            fill(k + 1);
          for (int R = 0; R < 3; ++R)
            for (int t = 0; t < 128; ++t)
              for (int j = 0; j < 2; ++j)
                {
                  TS & ts    = st[R][t];
                  auto after = [&]() {
                    if (R == 0)
                      for (int b = 0; b < 4; ++b)
                        ts.fa[b] = ts.eo[b];
                    if (j == 0)
                      request(R, cur, t, 1, ts.fa, ts.fb);
                    else if (k + 1 < walk.size())
                      request(R, walk[k + 1], t, 0, ts.fa, ts.fb);
                  };
                  auto after_main = [&]() {
                    const bool     more = j == 0 || k + 1 < walk.size();
                    const uint32_t ub2  = j == 0 ? ucur : unext;
                    if (!more)
                      return;
                    if (R == 0)
                      load_u<0>(ub2, tm0[t], 1 - j, ts.U);
                    else if (R == 1)
                      load_u<1>(ub2, tm1[t], 1 - j, ts.U);
                    else
                      load_u<2>(ub2, tm2[t], 1 - j, ts.U);
                  };
                  if (R == 0)
                    {
                      double edge[4];
                      task_round0(cf, pbuf, tm0[t], j, ts.U, ts.fa, ts.fb, descend, edge, after, after_main);
                      for (int b = 0; b < 4; ++b)
                        ts.eo[b] = edge[b];
                    }
                  else if (R == 1)
                    task_round1(cf, pbuf, tm1[t], j, ts.U, ts.fa, ts.fb, after, after_main, [] {});
                  else
                    {
                      double q[4][4];
                      task_round2(cf, pbuf, tm2[t], j, ts.U, ts.fa, ts.fb, q, after, after_main, [] {});
                      const long long g0 = cur.cell * CELL + (t & 15) + 16 * ((t >> 4) + 8 * j);
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
