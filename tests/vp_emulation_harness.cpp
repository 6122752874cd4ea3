// TEST HARNESS (tests/test_vp_kernel_emulation.py) — not product code, never linked into libhdgpu.so.
// Includes the source of the general-velocity kernel with the host-emulation switch: the per-cell body then runs as ONE
// sequential "thread" per cell with barriers as no-ops, which computes what the barrier-synchronised CUDA kernel computes.
#define HD_VP_HOST_EMULATION
#include <array>
#include "../hyperdeal_b200/csrc/kernel_vp.cu"

// test harness entry point (tests only): every cell of a periodic lattice, one sequential "thread" per cell
extern "C" int
hd_vp_emulate(const double *src, double *dst, const double *a_v, int dim_x, int dim_v, int degree, int n_points, const int *ncell, const double *left,
              const double *right, double skew)
{
  try
    {
      hd::Basis1D b;
      b.init(degree, n_points, false);
      const int dim = dim_x + dim_v, n = b.n, nq = b.nq;
      VpParams  p;
      double    h[HD_MAX_DIM];
      long long ncells = 1, nd = 1, cap = 1;
      for (int d = 0; d < HD_MAX_DIM; ++d)
        {
          p.ncell[d]       = d < dim ? ncell[d] : 1;
          p.cell_offset[d] = 0;
          p.left[d]        = d < dim ? left[d] : 0.0;
          h[d] = p.h[d] = d < dim ? (right[d] - left[d]) / ncell[d] : 1.0;
          if (d < dim)
            {
              ncells *= ncell[d];
              nd *= n;
              cap *= n > nq ? n : nq;
            }
        }
      std::vector<double> coef, basis;
      vp_coefficients(b, dim, h, skew, coef);
      for (auto v : b.nodes)
        basis.push_back((double)v);
      for (auto v : b.xq)
        basis.push_back((double)v);
      for (auto v : b.w)
        basis.push_back((double)v);
      for (auto v : b.S)
        basis.push_back((double)v);
      for (auto v : b.Sinv)
        basis.push_back((double)v);
      p.src = src;
      p.dst = dst;
      p.coef = coef.data();
      p.basis = basis.data();
      p.a_v = a_v;
      p.dim_x = dim_x;
      p.dim_v = dim_v;
      p.n = n;
      p.nq = nq;
      p.nd = nd;
      p.ncells = ncells;
      p.cap = (int)cap;
      p.sol = p.ti_next = nullptr;
      p.fb = p.fa = 0.0;
      p.fused = 0;
      std::vector<double> sm(6 * (size_t)cap + 2 * (size_t)n * n);
      for (long long cell = 0; cell < ncells; ++cell)
        vp_cell<double>(p, sm.data(), cell, 0, 1);
      return 0;
    }
  catch (const std::exception &)
    {
      return -1;
    }
}

// the degree-3 register-tile kernels (kernel_vp_tile.cuh): 1D1V one "thread" per cell; 2D2V the five phases of a warp, run
// lane by lane with the warp's shared-memory block emulated (a __syncwarp separates the phases on the device)
extern "C" int
hd_vp_tile_emulate(const double *src, double *dst, const double *a_v, int dim_x, const int *ncell, const double *left, const double *right, double skew)
{
  try
    {
      hd::Basis1D b;
      b.init(3, 4, false);
      const int dim = 2 * dim_x;
      if (dim != 2 && dim != 4 && dim != 6)
        return -2;
      VpParams  p;
      double    h[HD_MAX_DIM];
      long long ncells = 1;
      for (int d = 0; d < HD_MAX_DIM; ++d)
        {
          p.ncell[d]       = d < dim ? ncell[d] : 1;
          p.cell_offset[d] = 0;
          p.left[d]        = d < dim ? left[d] : 0.0;
          h[d] = p.h[d] = d < dim ? (right[d] - left[d]) / ncell[d] : 1.0;
          if (d < dim)
            ncells *= ncell[d];
        }
      std::vector<double> coef;
      vp_coefficients(b, dim, h, skew, coef);
      VpTileCoef cf;
      vp_tile_coefficients(b, dim, coef, cf);
      p.src = src;
      p.dst = dst;
      p.coef = nullptr;
      p.basis = nullptr;
      p.a_v = a_v;
      p.dim_x = p.dim_v = dim_x;
      p.n = p.nq = 4;
      p.nd = dim == 2 ? 16 : (dim == 4 ? 256 : 4096);
      p.ncells = ncells;
      p.cap = 0;
      p.sol = p.ti_next = nullptr;
      p.fb = p.fa = 0.0;
      p.fused = 0;
      if (dim == 2)
        {
          for (long long cell = 0; cell < ncells; ++cell)
            vpt2_cell<double>(p, cf, cell);
          return 0;
        }
      if (dim == 6)
        {
          // 3D3V: the phases of one CTA (256 threads), thread by thread; __syncthreads separates them on the device
          VpTile6Coef cf6;
          vp_tile6_coefficients(b, coef, cf6);
          std::vector<double> sm6(VPT6_SMEM, std::nan(""));
          std::vector<std::array<std::array<double, 4>, 4>> Us(256);
          for (long long cell = 0; cell < ncells; ++cell)
            {
              Vpt6Cell C;
              vpt6_decode(p, cell, C);
              double *sm = sm6.data();
              auto    U  = [&](int T) -> double(&)[4][4] { return *reinterpret_cast<double(*)[4][4]>(Us[T].data()); };
              struct Tr
              {
                double a[2][4], c[2][4], e[2][4], sv[8][2];
              };
              std::vector<Tr> tr(256);
              for (int T = 0; T < 256; ++T)
                {
                  vpt6_request_traces<double, 2>(p, C, T, tr[T].a);
                  vpt6_phase0<double>(p, cf6, C, sm, T);
                }
              for (int T = 0; T < 256; ++T)
                {
                  vpt6_request_traces<double, 0>(p, C, T, tr[T].c);
                  vpt6_phaseA(cf6, sm, T, tr[T].a);
                }
              for (int T = 0; T < 256; ++T)
                {
                  vpt6_mpart<2, false, false>(sm, 2, T);
                  vpt6_face_s2(cf6, sm, T);
                }
              for (int T = 0; T < 256; ++T)
                {
                  vpt6_request_traces<double, 1>(p, C, T, tr[T].e);
                  vpt6_phaseC(cf6, sm, T, U(T), tr[T].c);
                }
              for (int T = 0; T < 256; ++T)
                vpt6_mpart<1, false, true>(sm, 0, T);
              for (int T = 0; T < 256; ++T)
                vpt6_phaseE(cf6, sm, T, U(T), tr[T].e);
              for (int T = 0; T < 256; ++T)
                vpt6_mpart<2, true, true>(sm, 1, T);
              for (int T = 0; T < 256; ++T)
                vpt6_phaseG(p, cf6, C, sm, T);
              for (int T = 0; T < 256; ++T)
                vpt6_phaseH(p, cf6, C, sm, T);
              for (int T = 0; T < 256; ++T)
                {
                  vpt6_request_sol<double>(p, C, T, tr[T].sv);
                  vpt6_phaseI(cf6, sm, T);
                }
              for (int T = 0; T < 256; ++T)
                vpt6_phaseJ(cf6, sm, T);
              for (int T = 0; T < 256; ++T)
                vpt6_phaseK<double>(p, C, sm, T, tr[T].sv);
            }
          return 0;
        }
      std::vector<double> sm(VPT_WARP, std::nan(""));
      for (long long cell0 = 0; cell0 < ncells; cell0 += 2)
        {
          Vpt4Lane L[32];
          double   R[32][4][4];
          auto     cb   = [&](int lane) { return sm.data() + (lane >> 4) * VPT_CELL; };
          auto     mine = [&](int lane) { return std::min(cell0 + (lane >> 4), ncells - 1); };
          for (int lane = 0; lane < 32; ++lane)
            vpt4_phase0<double>(p, L[lane], cb(lane), lane & 15, mine(lane));
          for (int lane = 0; lane < 32; ++lane)
            vpt4_phase1<double>(p, cf, L[lane], cb(lane), lane & 15);
          for (int lane = 0; lane < 32; ++lane)
            {
              vpt4_phase2a(L[lane], cb(lane), lane & 15);
              vpt4_phase1b_stage(L[lane], cb(lane), lane & 15);
            }
          for (int lane = 0; lane < 32; ++lane)
            vpt4_phase1b(cf, L[lane], cb(lane), lane & 15);
          for (int lane = 0; lane < 32; ++lane)
            vpt4_phase2b(p, cf, L[lane], cb(lane), lane & 15, R[lane]);
          for (int lane = 0; lane < 32; ++lane)
            vpt4_phase2c(L[lane], cb(lane), lane & 15, R[lane]);
          for (int lane = 0; lane < 32; ++lane)
            vpt4_phase3(cf, cb(lane), lane & 15);
          for (int lane = 0; lane < 32; ++lane)
            if (cell0 + (lane >> 4) < ncells)
              vpt4_phase4<double>(p, L[lane], cb(lane), lane & 15);
        }
      return 0;
    }
  catch (const std::exception &)
    {
      return -1;
    }
}

// ---- x-space field solve (poisson_x.cu): operator, mass matrix and gradient bodies, one sequential "thread" per cell
#include "../hyperdeal_b200/csrc/poisson_x.cu"

namespace
{
  struct XsHost
  {
    hd::Basis1D         b;
    std::vector<double> coef, basis, sm;
    XsParams            p;
    XsHost(int dim_x, int degree, int n_points, const int *ncell, const double *h)
    {
      b.init(degree, n_points, false);
      xs_coefficients(b, dim_x, h, coef);
      for (auto v : b.nodes)
        basis.push_back((double)v);
      for (auto v : b.xq)
        basis.push_back((double)v);
      for (auto v : b.w)
        basis.push_back((double)v);
      for (auto v : b.S)
        basis.push_back((double)v);
      for (auto v : b.Sinv)
        basis.push_back((double)v);
      p.coef   = coef.data();
      p.basis  = basis.data();
      p.dim_x  = dim_x;
      p.n      = b.n;
      p.nq     = b.nq;
      p.nd     = 1;
      p.ncells = 1;
      long long cap = 1;
      for (int d = 0; d < 3; ++d)
        {
          p.ncell[d] = d < dim_x ? ncell[d] : 1;
          if (d < dim_x)
            {
              p.nd *= b.n;
              p.ncells *= ncell[d];
              cap *= b.n > b.nq ? b.n : b.nq;
            }
        }
      sm.resize(3 * (size_t)cap);
    }
  };
} // namespace

// which: 0 = K src, 1 = scale * M src, 2 = gradient table of src (dst: [cell][q][d])
extern "C" int
hd_xs_emulate(int which, const double *src, double *dst, int dim_x, int degree, int n_points, const int *ncell, const double *h, double scale)
{
  try
    {
      XsHost x(dim_x, degree, n_points, ncell, h);
      for (long long cell = 0; cell < x.p.ncells; ++cell)
        {
          if (which == 0)
            xs_laplace_cell(x.p, x.sm.data(), src, dst, cell, 0, 1);
          else if (which == 1)
            xs_mass_cell(x.p, x.sm.data(), src, dst, scale, cell, 0, 1);
          else
            xs_gradient_cell(x.p, x.sm.data(), src, dst, cell, 0, 1);
        }
      return 0;
    }
  catch (const std::exception &)
    {
      return -1;
    }
}

// ---- diagnostics (vp_diagnostics.cu)
#include "../hyperdeal_b200/csrc/vp_diagnostics.cu"

extern "C" int
hd_diag_emulate(const double *f, const double *a_v, double *out6, double *energy3, int dim_x, int dim_v, int degree, int n_points, const int *ncell, const double *left,
                const double *right)
{
  try
    {
      hd::Basis1D b;
      b.init(degree, n_points, false);
      std::vector<double> basis;
      for (auto v : b.nodes)
        basis.push_back((double)v);
      for (auto v : b.xq)
        basis.push_back((double)v);
      for (auto v : b.w)
        basis.push_back((double)v);
      for (auto v : b.S)
        basis.push_back((double)v);
      for (auto v : b.Sinv)
        basis.push_back((double)v);
      const int  dim = dim_x + dim_v;
      DiagParams p;
      p.basis = basis.data();
      p.dim_x = dim_x;
      p.dim_v = dim_v;
      p.n     = b.n;
      p.nq    = b.nq;
      p.nd = p.ncells = 1;
      long long cap = 1, ncx = 1;
      for (int d = 0; d < HD_MAX_DIM; ++d)
        {
          p.ncell[d]       = d < dim ? ncell[d] : 1;
          p.cell_offset[d] = 0;
          p.left[d]        = d < dim ? left[d] : 0.0;
          p.h[d]           = d < dim ? (right[d] - left[d]) / ncell[d] : 1.0;
          if (d < dim)
            {
              p.nd *= b.n;
              p.ncells *= ncell[d];
              cap *= b.n > b.nq ? b.n : b.nq;
            }
          if (d < dim_x)
            ncx *= ncell[d];
        }
      p.cap = (int)cap;
      std::vector<double> sm(2 * (size_t)cap + 6);
      for (int k = 0; k < 6; ++k)
        out6[k] = 0.0;
      for (long long cell = 0; cell < p.ncells; ++cell)
        {
          double partial[6];
          diag_cell<double>(p, sm.data(), f, partial, cell, 0, 1);
          for (int k = 0; k < 6; ++k)
            out6[k] += partial[k];
        }
      for (int d = 0; d < 3; ++d)
        energy3[d] = 0.0;
      for (long long cell = 0; cell < ncx; ++cell)
        {
          double partial[3] = {0, 0, 0};
          field_energy_cell(p, a_v, partial, cell);
          for (int d = 0; d < dim_x; ++d)
            energy3[d] += partial[d];
        }
      return 0;
    }
  catch (const std::exception &)
    {
      return -1;
    }
}

// ---- tile kernel with the partial sums in global memory (kernel_tile_global.cu), periodic lattices
#include "../hyperdeal_b200/csrc/kernel_tile_global.cu"

namespace
{
  template <int N>
  int
  tg_emulate_n(const double *src, double *dst, int dim, int degree, const int *ncell, const double *left, const double *right, const double *velocity, double skew,
               bool smem_partials = false)
  {
    hd::Basis1D b;
    b.init(degree, degree + 1, false);
    b.set_skew((hd::LD)skew);
    TgParams<double>  p;
    TgCoef<double, N> cf;
    p.src = src;
    p.dst = dst;
    p.ghost = nullptr;
    p.dim = dim;
    p.ncells = 1;
    for (int d = 0; d < HD_MAX_DIM; ++d)
      {
        p.ncell[d] = d < dim ? ncell[d] : 1;
        p.side_kind[d][0] = p.side_kind[d][1] = HD_SIDE_PERIODIC_LOCAL;
        p.ghost_off[d][0] = p.ghost_off[d][1] = 0;
        p.nb_mask[d] = 0;
        for (int i = 0; i < N * N; ++i)
          cf.C[d][i] = 0.0;
        for (int i = 0; i < N; ++i)
          cf.L0[d][i] = cf.L1[d][i] = 0.0;
        if (d >= dim)
          continue;
        p.ncells *= ncell[d];
        std::vector<hd::LD> C[4], L0, L1;
        b.direction_matrices((hd::LD)velocity[d], (hd::LD)((right[d] - left[d]) / ncell[d]), (hd::LD)skew, C, L0, L1);
        bool any0 = false, any1 = false;
        for (int i = 0; i < N * N; ++i)
          cf.C[d][i] = (double)C[0][i];
        for (int i = 0; i < N; ++i)
          {
            cf.L0[d][i] = (double)L0[i];
            cf.L1[d][i] = (double)L1[i];
            any0 |= L0[i] != 0;
            any1 |= L1[i] != 0;
          }
        p.nb_mask[d] = (any0 ? 1 : 0) | (any1 ? 2 : 0);
      }
    p.sol = p.ti_next = nullptr;
    p.fb = p.fa = 0.0;
    p.fused = 0;
    if (smem_partials)
      {
        // the variant that keeps the partial sums of the rounds in a CTA-private buffer (shared memory on the device)
        long long nd = 1;
        for (int d = 0; d < dim; ++d)
          nd *= N;
        std::vector<double> part((size_t)nd, std::nan(""));
        for (long long cell = 0; cell < p.ncells; ++cell)
          tg_cell<double, N>(p, cf, cell, 0, 1, part.data());
        return 0;
      }
    for (long long cell = 0; cell < p.ncells; ++cell)
      tg_cell<double, N>(p, cf, cell, 0, 1);
    return 0;
  }
} // namespace

extern "C" int
hd_tg_sp_emulate(const double *src, double *dst, int dim, int degree, const int *ncell, const double *left, const double *right, const double *velocity, double skew)
{
  try
    {
      if (degree == 5)
        return tg_emulate_n<6>(src, dst, dim, degree, ncell, left, right, velocity, skew, true);
      if (degree == 3)
        return tg_emulate_n<4>(src, dst, dim, degree, ncell, left, right, velocity, skew, true);
      return -2;
    }
  catch (const std::exception &)
    {
      return -1;
    }
}

extern "C" int
hd_tg_emulate(const double *src, double *dst, int dim, int degree, const int *ncell, const double *left, const double *right, const double *velocity, double skew)
{
  try
    {
      if (degree == 5)
        return tg_emulate_n<6>(src, dst, dim, degree, ncell, left, right, velocity, skew);
      if (degree == 3)
        return tg_emulate_n<4>(src, dst, dim, degree, ncell, left, right, velocity, skew);
      return -2;
    }
  catch (const std::exception &)
    {
      return -1;
    }
}
