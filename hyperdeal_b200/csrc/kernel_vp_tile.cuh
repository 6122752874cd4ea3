// Degree-3 general-velocity kernels (n = n_points = 4) for the Vlasov-Poisson lattices of BASELINE.json configs[3]:
// 1D1V and 2D2V.  Included by kernel_vp.cu inside its anonymous namespace (needs VpParams, HD_VP_FN).
//
// Same collapsed form as the generic vp_cell (kernel_vp.cu header), regrouped so that every sweep runs on a 4x4 register
// tile and nothing but 4x4 matrices is ever applied:
//     out = sum_{d in x} [ Ma_d(v_d) (x) Ca_d(x_d)  +  Mabs_d(v_d) (x) Cabs_d(x_d) ] u   (+ neighbour traces)
//         + Sinv_x [ sum_{d in v} ( G_d Ca_d + |G_d| Cabs_d )(v_d)  S_x u ]             (+ traces of S_x u_neighbour)
// Ma_d = Sinv diag(v(q)) S of the cell's v_d-coordinate (4x4, one per cell and direction), G_d = a_v table at the cell's
// x-quadrature points — a SCALAR for a thread that owns one x-quadrature point, so the v-direction matrices of the a- and the
// |a|-part merge into one 4x4 matrix per thread.  ~77 DFMA per DoF (the velocity varies inside the cell, so both upwind
// sides and the quadrature-point products are needed): FP64-bound — 16 B/DoF at the measured HBM peak would allow
// ~410 GDoF/s, 64 DFMA/clk/SM allow ~190 at 100 % pipe utilisation.
//
// 2D2V: a WARP owns two cells (16 lanes each) and alternates between two views of a cell,
//     X role: lane = (v0, v1), tile over (x0, x1)  — 16 contiguous values in global memory (128 B per lane)
//     V role: lane = (x0, x1), tile over (v0, v1)
// exchanging tiles through its private shared-memory buffers (tile stride 18 doubles: the 128-bit X-tile accesses and the
// 64-bit V-tile accesses are both conflict-free).  Only __syncwarp — no block barrier, warps drift freely.
//     phase 0 : coalesced fetch of the cell, the 16 face tiles of the v-neighbours and the x-traces -> BU, BF, TX
//     phase 1 : X: W = S S u -> BW (over BU);  face tiles: S S in place;  P0/Q0 = Ca_0/Cabs_0 u + traces -> BP/BQ;
//               Ma/Mabs entries -> BM
//     phase 2a: V: OX  = Ma_0 P0 + Mabs_0 Q0                       (v0 sweep)
//     phase 1b: X: P1/Q1 (x1 sweep) -> BP/BQ
//     phase 2b: V: OX += Ma_1 P1 + Mabs_1 Q1;  R = sum_d (G_d Ca_d + |G_d| Cabs_d) W + traces;  R -> BP, OX -> BQ
//     phase 3 : X: out = Sinv Sinv R + OX -> buffer 0
//     phase 4 : coalesced store of the cell (or the fused LSRK update)
// 1D1V: one thread per cell, the whole 4x4 cell in registers.
//
// The phase functions are plain host/device code (lane and buffers as arguments): tests/vp_emulation_harness.cpp runs them
// lane by lane on the CPU against the oracle (tests/test_vp_kernel_emulation.py); the product has no CPU path.

#ifdef HD_VP_HOST_EMULATION
#  define HD_VPT_FN inline
#else
#  define HD_VPT_FN __device__ __forceinline__
#endif

struct VpTileCoef
{
  double Ca[4][16], Cabs[4][16]; // [direction][out * 4 + in]
  double La0[4][4], La1[4][4], Labs0[4][4], Labs1[4][4];
  double S[16], Sinv[16]; // S[q * 4 + j], Sinv[i * 4 + q]
  double xq[4];
};

constexpr int VPT_TS   = 18;          // tile stride (doubles)
constexpr int VPT_BUF  = 16 * VPT_TS; // one cell buffer
constexpr int VPT_CELL = 4 * VPT_BUF + 64 + 128; // BU/BW, BF, BP, BQ + BM (Ma_0, Mabs_0, Ma_1, Mabs_1) + TX (traces: 2 sides x 64)
constexpr int VPT_WARP = 2 * VPT_CELL;

// out[b][a] (+)= sum_j M[a * 4 + j] in[b][j]: sweep along the fast tile index
template <bool ADD>
HD_VPT_FN void
vpt_sweep_a(const double *M, const double (&in)[4][4], double (&out)[4][4])
{
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int a = 0; a < 4; ++a)
      {
        double acc = ADD ? out[b][a] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc += M[a * 4 + j] * in[b][j];
        out[b][a] = acc;
      }
}

// out[b][a] (+)= sum_j M[b * 4 + j] in[j][a]: sweep along the slow tile index
template <bool ADD>
HD_VPT_FN void
vpt_sweep_b(const double *M, const double (&in)[4][4], double (&out)[4][4])
{
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int a = 0; a < 4; ++a)
      {
        double acc = ADD ? out[b][a] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc += M[b * 4 + j] * in[j][a];
        out[b][a] = acc;
      }
}

// 16 contiguous values (one X tile) from global memory
template <typename T>
HD_VPT_FN void
vpt_load16(const T *p, double (&U)[4][4])
{
#ifdef HD_VP_HOST_EMULATION
  for (int i = 0; i < 16; ++i)
    U[i >> 2][i & 3] = double(p[i]);
#else
  if (sizeof(T) == 8)
    {
      const double2 *q = reinterpret_cast<const double2 *>(p);
#  pragma unroll
      for (int i = 0; i < 8; ++i)
        {
          const double2 v           = __ldg(q + i);
          U[i >> 1][(i & 1) * 2]     = v.x;
          U[i >> 1][(i & 1) * 2 + 1] = v.y;
        }
    }
  else
    {
      const float4 *q = reinterpret_cast<const float4 *>(p);
#  pragma unroll
      for (int i = 0; i < 4; ++i)
        {
          const float4 v = __ldg(q + i);
          U[i][0]        = v.x;
          U[i][1]        = v.y;
          U[i][2]        = v.z;
          U[i][3]        = v.w;
        }
    }
#endif
}

// X tile <-> shared memory (tile `tile` of buffer `buf`); 16-byte accesses
HD_VPT_FN void
vpt_store_x(double *buf, const int tile, const double (&A)[4][4])
{
  double *q = buf + tile * VPT_TS;
#ifdef HD_VP_HOST_EMULATION
  for (int i = 0; i < 16; ++i)
    q[i] = A[i >> 2][i & 3];
#else
#  pragma unroll
  for (int i = 0; i < 8; ++i)
    reinterpret_cast<double2 *>(q)[i] = make_double2(A[i >> 1][(i & 1) * 2], A[i >> 1][(i & 1) * 2 + 1]);
#endif
}

HD_VPT_FN void
vpt_load_x(const double *buf, const int tile, double (&A)[4][4])
{
  const double *q = buf + tile * VPT_TS;
#ifdef HD_VP_HOST_EMULATION
  for (int i = 0; i < 16; ++i)
    A[i >> 2][i & 3] = q[i];
#else
#  pragma unroll
  for (int i = 0; i < 8; ++i)
    {
      const double2 v            = reinterpret_cast<const double2 *>(q)[i];
      A[i >> 1][(i & 1) * 2]     = v.x;
      A[i >> 1][(i & 1) * 2 + 1] = v.y;
    }
#endif
}

// V tile of lane tx: element (v0, v1) = buf[(v0 + 4 v1) * TS + tx]
HD_VPT_FN void
vpt_load_v(const double *buf, const int tx, double (&A)[4][4])
{
#pragma unroll
  for (int i = 0; i < 16; ++i)
    A[i >> 2][i & 3] = buf[i * VPT_TS + tx];
}

HD_VPT_FN void
vpt_store_v(double *buf, const int tx, const double (&A)[4][4])
{
#pragma unroll
  for (int i = 0; i < 16; ++i)
    buf[i * VPT_TS + tx] = A[i >> 2][i & 3];
}

// periodic neighbours of `cell` along direction d: {lower, upper}
HD_VPT_FN void
vpt_neighbours(const VpParams &p, const long long cell, const int cd, const long long cstr, const int d, long long &lo, long long &hi)
{
  lo = cd == 0 ? cell + (long long)(p.ncell[d] - 1) * cstr : cell - cstr;
  hi = cd == p.ncell[d] - 1 ? cell - (long long)(p.ncell[d] - 1) * cstr : cell + cstr;
}

// two consecutive values from global memory / to global memory (one 16-byte access for double)
template <typename T>
HD_VPT_FN void
vpt_load2(const T *p, double &a, double &b)
{
#ifdef HD_VP_HOST_EMULATION
  a = double(p[0]);
  b = double(p[1]);
#else
  if (sizeof(T) == 8)
    {
      const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
      a               = v.x;
      b               = v.y;
    }
  else
    {
      const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
      a              = v.x;
      b              = v.y;
    }
#endif
}

template <typename T>
HD_VPT_FN void
vpt_store2(T *p, const double a, const double b)
{
#ifdef HD_VP_HOST_EMULATION
  p[0] = T(a);
  p[1] = T(b);
#else
  if (sizeof(T) == 8)
    *reinterpret_cast<double2 *>(p) = make_double2(a, b);
  else
    *reinterpret_cast<float2 *>(p) = make_float2(float(a), float(b));
#endif
}

// two consecutive doubles in shared memory (16-byte aligned: even index)
HD_VPT_FN void
vpt_sm_store2(double *q, const double a, const double b)
{
#ifdef HD_VP_HOST_EMULATION
  q[0] = a;
  q[1] = b;
#else
  *reinterpret_cast<double2 *>(q) = make_double2(a, b);
#endif
}

HD_VPT_FN void
vpt_sm_load2(const double *q, double &a, double &b)
{
#ifdef HD_VP_HOST_EMULATION
  a = q[0];
  b = q[1];
#else
  const double2 v = *reinterpret_cast<const double2 *>(q);
  a               = v.x;
  b               = v.y;
#endif
}

// per-lane state that lives across the phases of the 2D2V kernel
struct Vpt4Lane
{
  double    U[4][4];  // X role: the lane's tile of u
  double    OX[4][4]; // V role: x-direction part of the result
  double    x1[2][2][2]; // this lane's share of the x_1 traces (fetched in phase 0, staged in phase 1b): [side][piece][2 values]
  long long cell;
  int       c[4];
  long long cstr[4];
};

// ---- 2D2V ---------------------------------------------------------------------------------------------------------------
// `cb` = the shared-memory block of this lane's cell (VPT_CELL doubles), t = lane within the cell (0..15).
// Phase 0: the 16 lanes of a cell fetch everything the cell needs from global memory with COALESCED accesses (consecutive
// lanes = consecutive 16-byte pieces) and lay it out tile by tile in shared memory: the cell (16 tiles), the 16 face tiles of
// the four v-neighbours, the x_0 traces (64 values per side, 8 bytes out of every 32) — a lane reading "its" 128-byte tile
// directly costs one L1 wavefront per lane and request, which made the first version of this kernel L1-bound (ncu:
// l1tex data pipe 96 %, 8x sector amplification).
template <typename T>
HD_VPT_FN void
vpt4_phase0(const VpParams &p, Vpt4Lane &L, double *cb, const int t, const long long cell)
{
  double * BU = cb, *BF = cb + VPT_BUF, *TX = cb + 4 * VPT_BUF + 64;
  const T *src = static_cast<const T *>(p.src);
  L.cell       = cell;
  {
    // (32-bit decode: a lattice of 2 KiB cells has far fewer than 2^31 of them; checked at launch)
    unsigned  r = (unsigned)cell;
    long long m = 1;
#pragma unroll
    for (int d = 0; d < 4; ++d)
      {
        const unsigned nc = (unsigned)p.ncell[d], q = r / nc;
        L.c[d]            = int(r - q * nc);
        r                 = q;
        L.cstr[d]         = m;
        m *= p.ncell[d];
      }
  }
  long long nb[4][2]; // [direction][side]
#pragma unroll
  for (int d = 0; d < 4; ++d)
    vpt_neighbours(p, cell, L.c[d], L.cstr[d], d, nb[d][0], nb[d][1]);
  double u[8][2], f[8][2], tx[4][2];
  // the cell: piece k = 16 i + t holds values 2k, 2k + 1
#pragma unroll
  for (int i = 0; i < 8; ++i)
    vpt_load2(src + cell * 256 + 2 * (16 * i + t), u[i][0], u[i][1]);
  // face tiles: tile ft = 2 i + (t >> 3) = (face f = i >> 1, o = ft & 3); face f = (v-direction f >> 1, side f & 1)
#pragma unroll
  for (int i = 0; i < 8; ++i)
    {
      const int ft = 2 * i + (t >> 3), o = ft & 3, dv = i >> 2, side = (i >> 1) & 1;
      const int layer = side ? 0 : 3; // the neighbour's layer that touches the shared face
      const int tvn   = dv == 0 ? layer + 4 * o : o + 4 * layer;
      vpt_load2(src + nb[2 + dv][side] * 256 + 16 * tvn + 2 * (t & 7), f[i][0], f[i][1]);
    }
  // x_0 traces: value k = 16 i + t of a side belongs to tile k >> 2, b = k & 3 (element 3 + 4 b resp. 4 b of that tile)
#pragma unroll
  for (int i = 0; i < 4; ++i)
    {
      const int k = 16 * i + t;
      tx[i][0]    = double(src[nb[0][0] * 256 + 16 * (k >> 2) + 3 + 4 * (k & 3)]);
      tx[i][1]    = double(src[nb[0][1] * 256 + 16 * (k >> 2) + 4 * (k & 3)]);
    }
  // x_1 traces: piece k = 16 i + t (i = 0, 1) = values 2 (k & 1), +1 of the 4 that tile k >> 1 contributes (elements 12..15 of
  // the lower neighbour's tile, 0..3 of the upper one's); kept in registers until phase 1b
#pragma unroll
  for (int side = 0; side < 2; ++side)
#pragma unroll
    for (int i = 0; i < 2; ++i)
      {
        const int k = 16 * i + t;
        vpt_load2(src + nb[1][side] * 256 + 16 * (k >> 1) + (side ? 0 : 12) + 2 * (k & 1), L.x1[side][i][0], L.x1[side][i][1]);
      }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    {
      const int k = 16 * i + t;
      vpt_sm_store2(BU + (k >> 3) * VPT_TS + 2 * (k & 7), u[i][0], u[i][1]);
      vpt_sm_store2(BF + (k >> 3) * VPT_TS + 2 * (k & 7), f[i][0], f[i][1]);
    }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    {
      TX[16 * i + t]      = tx[i][0];
      TX[64 + 16 * i + t] = tx[i][1];
    }
}

template <typename T>
HD_VPT_FN void
vpt4_phase1(const VpParams &p, const VpTileCoef &cf, Vpt4Lane &L, double *cb, const int t)
{
  double *BW = cb, *BF = cb + VPT_BUF, *BP = cb + 2 * VPT_BUF, *BQ = cb + 3 * VPT_BUF, *BM = cb + 4 * VPT_BUF, *TX = BM + 64;
  double  F[4][4];
  vpt_load_x(BW, t, L.U); // (the staged cell; W overwrites it tile by tile below)
  vpt_load_x(BF, t, F);
  double tl[4], th[4];
#pragma unroll
  for (int b = 0; b < 4; ++b)
    {
      tl[b] = TX[4 * t + b];
      th[b] = TX[64 + 4 * t + b];
    }
  // Ma / Mabs of both x-directions: entry (i, j) = t of each (the same for all cells with this v-coordinate)
  {
    const int i = t >> 2, j = t & 3;
#pragma unroll
    for (int d = 0; d < 2; ++d)
      {
        double ma = 0.0, mabs = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          {
            const double g = p.left[2 + d] + p.h[2 + d] * ((L.c[2 + d] + p.cell_offset[2 + d]) + cf.xq[q]);
            const double s = cf.Sinv[i * 4 + q] * cf.S[q * 4 + j];
            ma += s * g;
            mabs += s * fabs(g);
          }
        BM[d * 32 + t]      = ma;
        BM[d * 32 + 16 + t] = mabs;
      }
  }
  double A[4][4], B[4][4];
  // W = S(x1) S(x0) u
  vpt_sweep_a<false>(cf.S, L.U, A);
  vpt_sweep_b<false>(cf.S, A, B);
  vpt_store_x(BW, t, B);
  vpt_sweep_a<false>(cf.S, F, A);
  vpt_sweep_b<false>(cf.S, A, B);
  vpt_store_x(BF, t, B);
  // P0 / Q0: direction 0 = the tile's fast index
  vpt_sweep_a<false>(cf.Ca[0], L.U, A);
  vpt_sweep_a<false>(cf.Cabs[0], L.U, B);
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int a = 0; a < 4; ++a)
      {
        A[b][a] += cf.La0[0][a] * tl[b] + cf.La1[0][a] * th[b];
        B[b][a] += cf.Labs0[0][a] * tl[b] + cf.Labs1[0][a] * th[b];
      }
  vpt_store_x(BP, t, A);
  vpt_store_x(BQ, t, B);
}

HD_VPT_FN void
vpt4_phase2a(Vpt4Lane &L, const double *cb, const int t)
{
  const double *BP = cb + 2 * VPT_BUF, *BQ = cb + 3 * VPT_BUF, *BM = cb + 4 * VPT_BUF;
  double        A[4][4], M[16];
#pragma unroll
  for (int i = 0; i < 16; ++i)
    M[i] = BM[i];
  vpt_load_v(BP, t, A);
  vpt_sweep_a<false>(M, A, L.OX); // v0 = the fast index of a V tile
#pragma unroll
  for (int i = 0; i < 16; ++i)
    M[i] = BM[16 + i];
  vpt_load_v(BQ, t, A);
  vpt_sweep_a<true>(M, A, L.OX);
}

// phase 1b, first half: the x_1 traces fetched in phase 0 go to the (now free) trace buffer; a __syncwarp follows
HD_VPT_FN void
vpt4_phase1b_stage(const Vpt4Lane &L, double *cb, const int t)
{
  double *TX = cb + 4 * VPT_BUF + 64;
#pragma unroll
  for (int side = 0; side < 2; ++side)
#pragma unroll
    for (int i = 0; i < 2; ++i)
      {
        const int k = 16 * i + t;
        vpt_sm_store2(TX + 64 * side + 4 * (k >> 1) + 2 * (k & 1), L.x1[side][i][0], L.x1[side][i][1]);
      }
}

HD_VPT_FN void
vpt4_phase1b(const VpTileCoef &cf, Vpt4Lane &L, double *cb, const int t)
{
  double *BP = cb + 2 * VPT_BUF, *BQ = cb + 3 * VPT_BUF, *TX = cb + 4 * VPT_BUF + 64;
  double  tl[4], th[4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
    {
      tl[a] = TX[4 * t + a];
      th[a] = TX[64 + 4 * t + a];
    }
  double A[4][4], B[4][4];
  vpt_sweep_b<false>(cf.Ca[1], L.U, A);
  vpt_sweep_b<false>(cf.Cabs[1], L.U, B);
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int a = 0; a < 4; ++a)
      {
        A[b][a] += cf.La0[1][b] * tl[a] + cf.La1[1][b] * th[a];
        B[b][a] += cf.Labs0[1][b] * tl[a] + cf.Labs1[1][b] * th[a];
      }
  vpt_store_x(BP, t, A);
  vpt_store_x(BQ, t, B);
}

// V role: finishes the x-part and computes the v-part at this lane's x-quadrature point; returns R (to be stored after a
// __syncwarp: other lanes may still be reading BP / BQ)
HD_VPT_FN void
vpt4_phase2b(const VpParams &p, const VpTileCoef &cf, Vpt4Lane &L, const double *cb, const int t, double (&R)[4][4])
{
  const double *BW = cb, *BF = cb + VPT_BUF, *BP = cb + 2 * VPT_BUF, *BQ = cb + 3 * VPT_BUF, *BM = cb + 4 * VPT_BUF;
  double        A[4][4], M[16];
#pragma unroll
  for (int i = 0; i < 16; ++i)
    M[i] = BM[32 + i];
  vpt_load_v(BP, t, A);
  vpt_sweep_b<true>(M, A, L.OX); // v1 = the slow index of a V tile
#pragma unroll
  for (int i = 0; i < 16; ++i)
    M[i] = BM[48 + i];
  vpt_load_v(BQ, t, A);
  vpt_sweep_b<true>(M, A, L.OX);
  // v-part: G_d at (x-cell, quadrature point t)
  const long long cx = L.c[0] + (long long)p.ncell[0] * L.c[1];
  const double    g0 = p.a_v[(cx * 16 + t) * 2 + 0], g1 = p.a_v[(cx * 16 + t) * 2 + 1];
  const double    a0 = fabs(g0), a1 = fabs(g1);
  vpt_load_v(BW, t, A);
  // direction 2 (v0, fast index): faces f = 0 (lower), 1 (upper), tiles indexed by v1
#pragma unroll
  for (int i = 0; i < 16; ++i)
    M[i] = g0 * cf.Ca[2][i] + a0 * cf.Cabs[2][i];
  vpt_sweep_a<false>(M, A, R);
#pragma unroll
  for (int b = 0; b < 4; ++b)
    {
      const double fl = BF[(0 * 4 + b) * VPT_TS + t], fh = BF[(1 * 4 + b) * VPT_TS + t];
#pragma unroll
      for (int a = 0; a < 4; ++a)
        R[b][a] += (g0 * cf.La0[2][a] + a0 * cf.Labs0[2][a]) * fl + (g0 * cf.La1[2][a] + a0 * cf.Labs1[2][a]) * fh;
    }
  // direction 3 (v1, slow index): faces f = 2, 3, tiles indexed by v0
#pragma unroll
  for (int i = 0; i < 16; ++i)
    M[i] = g1 * cf.Ca[3][i] + a1 * cf.Cabs[3][i];
  vpt_sweep_b<true>(M, A, R);
#pragma unroll
  for (int a = 0; a < 4; ++a)
    {
      const double fl = BF[(2 * 4 + a) * VPT_TS + t], fh = BF[(3 * 4 + a) * VPT_TS + t];
#pragma unroll
      for (int b = 0; b < 4; ++b)
        R[b][a] += (g1 * cf.La0[3][b] + a1 * cf.Labs0[3][b]) * fl + (g1 * cf.La1[3][b] + a1 * cf.Labs1[3][b]) * fh;
    }
}

HD_VPT_FN void
vpt4_phase2c(const Vpt4Lane &L, double *cb, const int t, const double (&R)[4][4])
{
  vpt_store_v(cb + 2 * VPT_BUF, t, R);
  vpt_store_v(cb + 3 * VPT_BUF, t, L.OX);
}

template <typename T>
HD_VPT_FN void
vpt_store_result(const VpParams &p, const long long g, const double (&O)[4][4])
{
  T *dst = static_cast<T *>(p.dst), *sol = static_cast<T *>(p.sol), *tin = static_cast<T *>(p.ti_next);
#ifndef HD_VP_HOST_EMULATION
  if (sizeof(T) == 8)
    {
      // 16-byte accesses (g is a multiple of 16 values)
      if (p.fused)
        {
          double2 *s2 = reinterpret_cast<double2 *>(sol + g), *t2 = reinterpret_cast<double2 *>(tin + g);
#  pragma unroll
          for (int i = 0; i < 8; ++i)
            {
              const double2 s  = s2[i];
              const double  v0 = O[i >> 1][(i & 1) * 2], v1 = O[i >> 1][(i & 1) * 2 + 1];
              s2[i]            = make_double2(s.x + p.fb * v0, s.y + p.fb * v1);
              if (p.fa != 0.0)
                t2[i] = make_double2(s.x + p.fa * v0, s.y + p.fa * v1);
            }
        }
      else
        {
          double2 *d2 = reinterpret_cast<double2 *>(dst + g);
#  pragma unroll
          for (int i = 0; i < 8; ++i)
            d2[i] = make_double2(O[i >> 1][(i & 1) * 2], O[i >> 1][(i & 1) * 2 + 1]);
        }
      return;
    }
#endif
#pragma unroll
  for (int i = 0; i < 16; ++i)
    {
      const double v = O[i >> 2][i & 3];
      if (p.fused)
        {
          const double s = double(sol[g + i]);
          sol[g + i]     = T(s + p.fb * v);
          if (p.fa != 0.0)
            tin[g + i] = T(s + p.fa * v);
        }
      else
        dst[g + i] = T(v);
    }
}

// X role: the result tile goes to shared memory (buffer 0: W is no longer needed) for the coalesced store of phase 4
HD_VPT_FN void
vpt4_phase3(const VpTileCoef &cf, double *cb, const int t)
{
  double A[4][4], B[4][4], O[4][4];
  vpt_load_x(cb + 2 * VPT_BUF, t, A);
  vpt_load_x(cb + 3 * VPT_BUF, t, O);
  vpt_sweep_a<false>(cf.Sinv, A, B);
  vpt_sweep_b<true>(cf.Sinv, B, O);
  vpt_store_x(cb, t, O);
}

// coalesced store (or the fused LSRK update): piece k = 16 i + t holds values 2k, 2k + 1 of the cell
template <typename T>
HD_VPT_FN void
vpt4_phase4(const VpParams &p, const Vpt4Lane &L, const double *cb, const int t)
{
  T *             dst = static_cast<T *>(p.dst), *sol = static_cast<T *>(p.sol), *tin = static_cast<T *>(p.ti_next);
  const long long g0  = L.cell * 256;
  if (p.fused)
    {
      double s[8][2];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        vpt_load2(sol + g0 + 2 * (16 * i + t), s[i][0], s[i][1]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        {
          const int k = 16 * i + t;
          double    v0, v1;
          vpt_sm_load2(cb + (k >> 3) * VPT_TS + 2 * (k & 7), v0, v1);
          vpt_store2(sol + g0 + 2 * k, s[i][0] + p.fb * v0, s[i][1] + p.fb * v1);
          if (p.fa != 0.0)
            vpt_store2(tin + g0 + 2 * k, s[i][0] + p.fa * v0, s[i][1] + p.fa * v1);
        }
    }
  else
    {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        {
          const int k = 16 * i + t;
          double    v0, v1;
          vpt_sm_load2(cb + (k >> 3) * VPT_TS + 2 * (k & 7), v0, v1);
          vpt_store2(dst + g0 + 2 * k, v0, v1);
        }
    }
}

// ---- 1D1V: one thread per cell; tile U[v0][x0] ------------------------------------------------------------------------------
template <typename T>
HD_VPT_FN void
vpt2_cell(const VpParams &p, const VpTileCoef &cf, const long long cell)
{
  const T * src = static_cast<const T *>(p.src);
  const int c1 = int((unsigned)cell / (unsigned)p.ncell[0]), c0 = int((unsigned)cell - (unsigned)c1 * (unsigned)p.ncell[0]);
  long long lo0, hi0, lo1, hi1;
  vpt_neighbours(p, cell, c0, 1, 0, lo0, hi0);
  vpt_neighbours(p, cell, c1, p.ncell[0], 1, lo1, hi1);
  double U[4][4], P[4][4], Q[4][4], O[4][4], M[16], Mb[16];
  vpt_load16(src + cell * 16, U);
  // x-part: direction 0 along the fast index, Ma / Mabs of the cell's v-coordinate along the slow index
  double tl[4], th[4];
#pragma unroll
  for (int b = 0; b < 4; ++b)
    {
      tl[b] = double(src[lo0 * 16 + 3 + 4 * b]);
      th[b] = double(src[hi0 * 16 + 4 * b]);
    }
  vpt_sweep_a<false>(cf.Ca[0], U, P);
  vpt_sweep_a<false>(cf.Cabs[0], U, Q);
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int a = 0; a < 4; ++a)
      {
        P[b][a] += cf.La0[0][a] * tl[b] + cf.La1[0][a] * th[b];
        Q[b][a] += cf.Labs0[0][a] * tl[b] + cf.Labs1[0][a] * th[b];
      }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      {
        double ma = 0.0, mabs = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          {
            const double g = p.left[1] + p.h[1] * ((c1 + p.cell_offset[1]) + cf.xq[q]);
            const double s = cf.Sinv[i * 4 + q] * cf.S[q * 4 + j];
            ma += s * g;
            mabs += s * fabs(g);
          }
        M[i * 4 + j]  = ma;
        Mb[i * 4 + j] = mabs;
      }
  vpt_sweep_b<false>(M, P, O);
  vpt_sweep_b<true>(Mb, Q, O);
  // v-part: W = S(x0) u and the S-transformed end rows of the v-neighbours; G = a_v at the 4 x-quadrature points
  double W[4][4], fl[4], fh[4], g[4], ga[4];
  vpt_sweep_a<false>(cf.S, U, W);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    {
      double sl = 0.0, sh = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        {
          sl += cf.S[q * 4 + j] * double(src[lo1 * 16 + 12 + j]);
          sh += cf.S[q * 4 + j] * double(src[hi1 * 16 + j]);
        }
      fl[q] = sl;
      fh[q] = sh;
      g[q]  = p.a_v[(long long)c0 * 4 + q];
      ga[q] = fabs(g[q]);
    }
  vpt_sweep_b<false>(cf.Ca[1], W, P);
  vpt_sweep_b<false>(cf.Cabs[1], W, Q);
  double R[4][4];
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int a = 0; a < 4; ++a)
      R[b][a] = g[a] * (P[b][a] + cf.La0[1][b] * fl[a] + cf.La1[1][b] * fh[a]) + ga[a] * (Q[b][a] + cf.Labs0[1][b] * fl[a] + cf.Labs1[1][b] * fh[a]);
  vpt_sweep_a<true>(cf.Sinv, R, O);
  vpt_store_result<T>(p, cell * 16, O);
}

#ifndef HD_VP_HOST_EMULATION
// 2D2V: WARPS warps per CTA, two cells per warp; MINB CTAs per SM bound the register allocation (12 warps per SM = 168
// registers, no spills)
template <typename T, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_vp_tile_2d2v(const VpParams p, const __grid_constant__ VpTileCoef cf)
{
  extern __shared__ double sm[];
  const int       lane = threadIdx.x & 31, warp = threadIdx.x >> 5, t = lane & 15;
  const long long cell = ((long long)blockIdx.x * WARPS + warp) * 2 + (lane >> 4);
  double *        cb   = sm + (size_t)warp * VPT_WARP + (lane >> 4) * VPT_CELL;
  // an odd cell count leaves the upper half of the last warp without a cell: it runs along on the last cell (keeps the
  // warp converged for the __syncwarp's) and skips the store
  const bool      live = cell < p.ncells;
  const long long mine = live ? cell : p.ncells - 1;
  Vpt4Lane        L;
  vpt4_phase0<T>(p, L, cb, t, mine);
  __syncwarp();
  vpt4_phase1<T>(p, cf, L, cb, t);
  __syncwarp();
  vpt4_phase2a(L, cb, t);
  vpt4_phase1b_stage(L, cb, t);
  __syncwarp();
  vpt4_phase1b(cf, L, cb, t);
  __syncwarp();
  double R[4][4];
  vpt4_phase2b(p, cf, L, cb, t, R);
  __syncwarp();
  vpt4_phase2c(L, cb, t, R);
  __syncwarp();
  vpt4_phase3(cf, cb, t);
  __syncwarp();
  if (live)
    vpt4_phase4<T>(p, L, cb, t);
}

template <typename T>
__global__ void __launch_bounds__(128) k_vp_tile_1d1v(const VpParams p, const __grid_constant__ VpTileCoef cf)
{
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell < p.ncells)
    vpt2_cell<T>(p, cf, cell);
}
#endif
