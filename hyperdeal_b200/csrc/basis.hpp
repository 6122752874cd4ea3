// 1-D basis data and the collapsed 1-D advection matrices (host, long double).
//
// The reference pulls S, D, S^-1 from dealii::internal::MatrixFreeFunctions::ShapeInfo
// (matrix_free/fe_evaluation_cell.h:93-95, fe_evaluation_cell_inverse.h:96-97,
// operators/advection/advection_operation.h:279-282).  deal.II is an external dependency, so
// the published definitions are recomputed here: FE_DGQ(k) = Lagrange basis on the k+1
// Gauss-Lobatto points of [0,1], QGauss(n_q) / QGaussLobatto(n_q) quadrature.
//
// On a Cartesian cell with a velocity that is constant along each 1-D line, the reference's
// per-cell pipeline (advection_operation.h:221-566: S sweeps -> metric/velocity at quadrature
// points -> D^T sweeps -> face interpolation + upwind flux -> JxW^-1 -> S^-1 sweeps) is, line by
// line, the linear map
//      dst_line = C u_line + L0 * trace_lower_neighbour + L1 * trace_upper_neighbour
// with   V   = S^-1 ( -s D + (1-s) W^-1 D^T W ) S,          l_f = S^-1 W^-1 f_f
//        C   = (a/h) V + sum_f alpha_f l_f e_f^T,            L_f = beta_f l_f
//        alpha_f = -( a n_f/2 + |a|/2 - s a n_f ) / h,       beta_f = -( a n_f - |a| ) / (2h)
// (s = SkewFactor, f_f = collocation basis at xi = f, e_f = unit vector of the end node: the GLL
// basis is interpolatory at the end points, so the own trace is a nodal value).  The transverse
// S / S^-1 pairs cancel.  A Dirichlet face (u+ = -u- + 2g) replaces alpha_f by alpha_f - beta_f
// and adds 2 beta_f l_f g.   Derivation and numerical check: DESIGN.md §3.
#pragma once
#include <cmath>
#include <stdexcept>
#include <vector>

namespace hd
{
  typedef long double LD;

  inline void
  legendre(int n, LD x, LD &p, LD &dp)
  {
    LD p0 = 1, p1 = x;
    if (n == 0)
      {
        p  = 1;
        dp = 0;
        return;
      }
    for (int k = 2; k <= n; ++k)
      {
        const LD p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / LD(k);
        p0          = p1;
        p1          = p2;
      }
    p  = p1;
    dp = n * (x * p1 - p0) / (x * x - 1);
  }

  // QGauss(nq) on [0,1]
  inline void
  gauss_legendre(int nq, std::vector<LD> &x, std::vector<LD> &w)
  {
    x.resize(nq);
    w.resize(nq);
    const LD pi = acosl(-1.0L);
    for (int i = 0; i < nq; ++i)
      {
        LD z = -cosl(pi * (i + 0.75L) / (nq + 0.5L));
        for (int it = 0; it < 100; ++it)
          {
            LD p, dp;
            legendre(nq, z, p, dp);
            const LD dz = p / dp;
            z -= dz;
            if (fabsl(dz) < 1e-19L)
              break;
          }
        LD p, dp;
        legendre(nq, z, p, dp);
        x[i] = (z + 1) / 2;
        w[i] = 1 / ((1 - z * z) * dp * dp);
      }
  }

  // QGaussLobatto(n) on [0,1] (support points of FE_DGQ(n-1))
  inline void
  gauss_lobatto(int n, std::vector<LD> &x, std::vector<LD> &w)
  {
    x.resize(n);
    w.resize(n);
    const int k  = n - 1;
    const LD  pi = acosl(-1.0L);
    std::vector<LD> z(n);
    z[0]     = -1;
    z[n - 1] = 1;
    for (int i = 1; i < n - 1; ++i)
      {
        // roots of P_k'(x): start from Chebyshev-Gauss-Lobatto points
        LD zi = -cosl(pi * i / LD(k));
        for (int it = 0; it < 200; ++it)
          {
            LD p, dp;
            legendre(k, zi, p, dp);
            const LD ddp = (2 * zi * dp - k * (k + 1) * p) / (1 - zi * zi);
            const LD dz  = dp / ddp;
            zi -= dz;
            if (fabsl(dz) < 1e-19L)
              break;
          }
        z[i] = zi;
      }
    for (int i = 0; i < n; ++i)
      {
        LD p, dp;
        if (i == 0 || i == n - 1)
          {
            // P_k(+-1) = (+-1)^k
            p = (i == 0 && (k % 2)) ? -1 : 1;
          }
        else
          legendre(k, z[i], p, dp);
        x[i] = (z[i] + 1) / 2;
        w[i] = 1 / (k * (k + 1) * p * p);
      }
  }

  // L[q*n+i] = l_i(x_q) for the Lagrange basis on `nodes`
  inline std::vector<LD>
  lagrange_eval(const std::vector<LD> &nodes, const std::vector<LD> &x)
  {
    const int       n = nodes.size(), m = x.size();
    std::vector<LD> L(m * n, 1);
    for (int q = 0; q < m; ++q)
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
          if (j != i)
            L[q * n + i] *= (x[q] - nodes[j]) / (nodes[i] - nodes[j]);
    return L;
  }

  // G[q*n+i] = l_i'(x_q)
  inline std::vector<LD>
  lagrange_deriv(const std::vector<LD> &nodes, const std::vector<LD> &x)
  {
    const int       n = nodes.size(), m = x.size();
    std::vector<LD> G(m * n, 0);
    for (int q = 0; q < m; ++q)
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
          {
            if (j == i)
              continue;
            LD term = 1 / (nodes[i] - nodes[j]);
            for (int l = 0; l < n; ++l)
              if (l != i && l != j)
                term *= (x[q] - nodes[l]) / (nodes[i] - nodes[l]);
            G[q * n + i] += term;
          }
    return G;
  }

  inline std::vector<LD>
  matmul(const std::vector<LD> &A, const std::vector<LD> &B, int m, int k, int n)
  {
    std::vector<LD> C(m * n, 0);
    for (int i = 0; i < m; ++i)
      for (int l = 0; l < k; ++l)
        for (int j = 0; j < n; ++j)
          C[i * n + j] += A[i * k + l] * B[l * n + j];
    return C;
  }

  inline std::vector<LD>
  transpose(const std::vector<LD> &A, int m, int n)
  {
    std::vector<LD> T(n * m);
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < n; ++j)
        T[j * m + i] = A[i * n + j];
    return T;
  }

  inline std::vector<LD>
  inverse(std::vector<LD> A, int n)
  {
    std::vector<LD> I(n * n, 0);
    for (int i = 0; i < n; ++i)
      I[i * n + i] = 1;
    for (int c = 0; c < n; ++c)
      {
        int p = c;
        for (int r = c + 1; r < n; ++r)
          if (fabsl(A[r * n + c]) > fabsl(A[p * n + c]))
            p = r;
        if (A[p * n + c] == 0)
          throw std::runtime_error("singular 1-D basis matrix");
        for (int j = 0; j < n; ++j)
          {
            std::swap(A[c * n + j], A[p * n + j]);
            std::swap(I[c * n + j], I[p * n + j]);
          }
        const LD piv = A[c * n + c];
        for (int j = 0; j < n; ++j)
          {
            A[c * n + j] /= piv;
            I[c * n + j] /= piv;
          }
        for (int r = 0; r < n; ++r)
          if (r != c)
            {
              const LD f = A[r * n + c];
              for (int j = 0; j < n; ++j)
                {
                  A[r * n + j] -= f * A[c * n + j];
                  I[r * n + j] -= f * I[c * n + j];
                }
            }
      }
    return I;
  }

  struct Basis1D
  {
    int             n = 0, nq = 0;
    bool            collocation = false;
    std::vector<LD> nodes, xq, w;
    std::vector<LD> w_nodes; // Gauss-Lobatto weights at the nodes (the reference's quadrature index 2)
    std::vector<LD> S;    // nq x n
    std::vector<LD> D;    // nq x nq
    std::vector<LD> Sinv; // n x nq
    std::vector<LD> f0, f1;
    std::vector<LD> V;      // n x n, skew-dependent volume matrix (set by set_skew)
    std::vector<LD> l0, l1; // n, lifting vectors

    void
    init(int degree, int n_points, bool colloc)
    {
      n           = degree + 1;
      nq          = n_points;
      collocation = colloc;
      if (colloc && nq != n)
        throw std::runtime_error("collocation requires n_points == degree + 1");
      gauss_lobatto(n, nodes, w_nodes);
      if (colloc)
        gauss_lobatto(nq, xq, w);
      else
        gauss_legendre(nq, xq, w);
      S = lagrange_eval(nodes, xq);
      D = lagrange_deriv(xq, xq);
      std::vector<LD> zero(1, 0.0L), one(1, 1.0L);
      f0 = lagrange_eval(xq, zero);
      f1 = lagrange_eval(xq, one);
      if (nq == n)
        Sinv = inverse(S, n);
      else
        {
          // n_q != n: W-weighted L2 projection (S^T W S)^-1 S^T W — reproduces the reference's
          // adv_2D_2D_k3_q5 goldens (tests/test_oracle_golden.py)
          std::vector<LD> WS(nq * n);
          for (int q = 0; q < nq; ++q)
            for (int i = 0; i < n; ++i)
              WS[q * n + i] = w[q] * S[q * n + i];
          const auto St  = transpose(S, nq, n);
          const auto M   = matmul(St, WS, n, nq, n);
          const auto Mi  = inverse(M, n);
          const auto WSt = transpose(WS, nq, n); // = S^T W
          Sinv           = matmul(Mi, WSt, n, n, nq);
        }
    }

    void
    set_skew(LD s)
    {
      // B = -s D + (1-s) W^-1 D^T W      (nq x nq)
      std::vector<LD> B(nq * nq);
      for (int q = 0; q < nq; ++q)
        for (int p = 0; p < nq; ++p)
          B[q * nq + p] = -s * D[q * nq + p] + (1 - s) * D[p * nq + q] * w[p] / w[q];
      V = matmul(Sinv, matmul(B, S, nq, nq, n), n, nq, n);
      std::vector<LD> g0(nq), g1(nq);
      for (int q = 0; q < nq; ++q)
        {
          g0[q] = f0[q] / w[q];
          g1[q] = f1[q] / w[q];
        }
      l0 = matmul(Sinv, g0, n, nq, 1);
      l1 = matmul(Sinv, g1, n, nq, 1);
    }

    // Collapsed matrices of one direction: speed a, cell size h.
    // C[variant][n*n] with variant = (lower side is Dirichlet) + 2*(upper side is Dirichlet);
    // L0/L1 = neighbour lifting vectors (zero for the outflow side);
    // G0/G1 = 2*beta_f*l_f, the lifting of the Dirichlet datum g.
    void
    // level: the reference's AdvectionOperationEvaluationLevel (advection_operation.h:37-42) — 0 = all, 1 = cell (no face
    // integrals at all: alpha = beta = 0), 2 = all_without_neighbor_load (the neighbour's trace is not read: beta-part of
    // the interior faces dropped, the own-side face term alpha kept)
    direction_matrices(LD a, LD h, LD s, std::vector<LD> C[4], std::vector<LD> &L0, std::vector<LD> &L1, const int level = 0) const
    {
      LD alpha[2], beta[2];
      for (int f = 0; f < 2; ++f)
        {
          const LD an = a * (f ? 1 : -1);
          alpha[f]    = level == 1 ? LD(0) : -(an / 2 + fabsl(a) / 2 - s * an) / h;
          beta[f]     = level == 1 ? LD(0) : -(an - fabsl(a)) / 2 / h;
        }
      for (int var = 0; var < 4; ++var)
        {
          C[var].assign(n * n, 0);
          for (int i = 0; i < n * n; ++i)
            C[var][i] = (a / h) * V[i];
          for (int f = 0; f < 2; ++f)
            {
              const bool dirichlet = (var >> f) & 1;
              const LD   coef      = dirichlet ? alpha[f] - beta[f] : alpha[f];
              const int  e         = f ? n - 1 : 0;
              const auto &l        = f ? l1 : l0;
              for (int i = 0; i < n; ++i)
                C[var][i * n + e] += coef * l[i];
            }
        }
      L0.resize(n);
      L1.resize(n);
      for (int i = 0; i < n; ++i)
        {
          L0[i] = level == 2 ? LD(0) : beta[0] * l0[i];
          L1[i] = level == 2 ? LD(0) : beta[1] * l1[i];
        }
    }
  };
} // namespace hd
