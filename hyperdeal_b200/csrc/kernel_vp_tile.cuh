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

// ---- 3D3V ---------------------------------------------------------------------------------------------------------------
// One CTA of 256 threads per cell (4096 values).  The six directions form three pairs, (x0,x1), (x2,v0), (v1,v2), and a
// thread owns the 4x4 tile over one pair — three views of the same shared-memory buffer:
//     P0: thread = tile number tau = (x2,v0,v1,v2),  tile over (x0,x1): 16 contiguous values (stride-18 padding as in 2D2V)
//     P1: thread = (e01 = x0 + 4 x1, w = v1 + 4 v2), tile over (x2,v0)
//     P2: thread = c = x0 + 4 x1 + 16 x2 + 64 v0,    tile over (v1,v2)
// (all three conflict-free on buffers of 256 tiles x 18 doubles).  Four cell buffers B0..B3 (u / S_x u, two temporaries, the
// x-part of the result) and six face buffers (the layers of the v-neighbours, S_x-transformed in shared memory).  Phases,
// separated by __syncthreads (x-direction d pairs C(x_d) with M(v_d); the views are chosen so that u dies early):
//     0  coalesced fetch: cell -> B0, six v-face layers -> BF, Ma/Mabs of the three v-coordinates -> BM
//     A  [P1] T1/T2 = Ca_2/Cabs_2(x2) u + traces -> B1/B2;            face tiles: S(x0) S(x1) in place
//     B  [P2] OUT  = Ma_2(v2) T1 + Mabs_2(v2) T2 -> B3;               face lines: S(x2) in place
//     C  [P0] U <- B0;  T1/T2 = Ca_0/Cabs_0(x0) U + traces;  B0 <- S(x0) S(x1) U
//     D  [P1] OUT += Ma_0(v0) T1 + Mabs_0(v0) T2
//     E  [P0] T1/T2 = Ca_1/Cabs_1(x1) U + traces
//     F  [P2] OUT += Ma_1(v1) T1 + Mabs_1(v1) T2
//     G  [P1] W = S(x2) B0 -> B0;  R = (G_0 Ca_3 + |G_0| Cabs_3)(v0) W + traces -> B1      (G_0 varies along the tile's x2 index)
//     H  [P2] R += (G_1 Ca_4 + |G_1| Cabs_4)(v1) W + (G_2 Ca_5 + |G_2| Cabs_5)(v2) W + traces  (G scalar per thread)
//     I  [P1] R = Sinv(x2) R
//     J  [P0] out = Sinv(x0) Sinv(x1) R + OUT -> B2
//     K  coalesced store of B2 (or the fused LSRK update)
// ~110 DFMA per DoF.  The phase functions take the thread index and the shared-memory block as arguments:
// tests/vp_emulation_harness.cpp runs them thread by thread on the CPU against the oracle.

struct VpTile6Coef
{
  double Ca[6][16], Cabs[6][16]; // [direction][out * 4 + in]
  double La0[6][4], La1[6][4], Labs0[6][4], Labs1[6][4];
  double S[16], Sinv[16];
  double xq[4];
};

constexpr int VPT6_BUF  = 256 * VPT_TS;                 // one cell buffer (doubles)
constexpr int VPT6_FACE = 64 * VPT_TS;                  // one face layer
constexpr int VPT6_BM   = 4 * VPT6_BUF + 6 * VPT6_FACE; // offset of the Ma/Mabs block: [d][2][16]
constexpr int VPT6_SMEM = VPT6_BM + 96;                 // doubles

// tile element (a, b) of thread T in view P -> index into a cell buffer
template <int P>
HD_VPT_FN int
vpt6_addr(const int T, const int a, const int b)
{
  if (P == 0)
    return T * VPT_TS + a + 4 * b;
  if (P == 1)
    return (a + 4 * b + 16 * (T >> 4)) * VPT_TS + (T & 15);
  return ((T >> 4) + 16 * (a + 4 * b)) * VPT_TS + (T & 15);
}

template <int P>
HD_VPT_FN void
vpt6_load(const double *buf, const int T, double (&A)[4][4])
{
  if (P == 0)
    vpt_load_x(buf, T, A);
  else
    {
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int a = 0; a < 4; ++a)
          A[b][a] = buf[vpt6_addr<P>(T, a, b)];
    }
}

template <int P>
HD_VPT_FN void
vpt6_store(double *buf, const int T, const double (&A)[4][4])
{
  if (P == 0)
    vpt_store_x(buf, T, A);
  else
    {
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int a = 0; a < 4; ++a)
          buf[vpt6_addr<P>(T, a, b)] = A[b][a];
    }
}

struct Vpt6Cell // per-CTA (identical in all threads): the cell, its coordinates and neighbours
{
  long long cell;
  int       c[6];
  long long nb[6][2]; // [direction][side]
};

HD_VPT_FN void
vpt6_decode(const VpParams &p, const long long cell, Vpt6Cell &C)
{
  C.cell = cell;
  unsigned  r = (unsigned)cell;
  long long m = 1;
#pragma unroll
  for (int d = 0; d < 6; ++d)
    {
      const unsigned nc = (unsigned)p.ncell[d], q = r / nc;
      C.c[d]            = int(r - q * nc);
      r                 = q;
      vpt_neighbours(p, cell, C.c[d], m, d, C.nb[d][0], C.nb[d][1]);
      m *= p.ncell[d];
    }
}

// phase 0
template <typename T_>
HD_VPT_FN void
vpt6_phase0(const VpParams &p, const VpTile6Coef &cf, const Vpt6Cell &C, double *sm, const int T)
{
  const T_ *src = static_cast<const T_ *>(p.src);
  double *  B0 = sm, *BF = sm + 4 * VPT6_BUF, *BM = sm + VPT6_BM;
  // the cell: piece k = 256 i + T = values 2k, 2k + 1
#pragma unroll
  for (int i = 0; i < 8; ++i)
    {
      const int k = 256 * i + T;
      double    a, b;
      vpt_load2(src + C.cell * 4096 + 2 * k, a, b);
      vpt_sm_store2(B0 + (k >> 3) * VPT_TS + 2 * (k & 7), a, b);
    }
  // the six v-face layers: piece k = 256 i + T (i < 12): face f = k / 512 = (v-direction f >> 1, side f & 1), face tile tf =
  // x2 + 4 o1 + 16 o2 (o1, o2 = the other two v-indices, ascending), part = k & 7
#pragma unroll
  for (int i = 0; i < 12; ++i)
    {
      const int k = 256 * i + T, f = i >> 1, dv = f >> 1, side = f & 1, rem = k & 511, tf = rem >> 3, part = rem & 7;
      const int layer = side ? 0 : 3, x2 = tf & 3, o1 = (tf >> 2) & 3, o2 = tf >> 4;
      const int taun  = dv == 0 ? x2 + 4 * layer + 16 * o1 + 64 * o2 : (dv == 1 ? x2 + 4 * o1 + 16 * layer + 64 * o2 : x2 + 4 * o1 + 16 * o2 + 64 * layer);
      const long long nbc = dv == 0 ? C.nb[3][side] : (dv == 1 ? C.nb[4][side] : C.nb[5][side]);
      double    a, b;
      vpt_load2(src + nbc * 4096 + 16 * taun + 2 * part, a, b);
      vpt_sm_store2(BF + f * VPT6_FACE + tf * VPT_TS + 2 * part, a, b);
    }
  // Ma / Mabs of the three v-coordinates: BM[d][0 | 1][i * 4 + j]
  if (T < 48)
    {
      const int d = T >> 4, i = (T >> 2) & 3, j = T & 3;
      const int cd = d == 0 ? C.c[3] : (d == 1 ? C.c[4] : C.c[5]);
      const double left = d == 0 ? p.left[3] : (d == 1 ? p.left[4] : p.left[5]), h = d == 0 ? p.h[3] : (d == 1 ? p.h[4] : p.h[5]);
      const int off = d == 0 ? p.cell_offset[3] : (d == 1 ? p.cell_offset[4] : p.cell_offset[5]);
      double ma = 0.0, mabs = 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        {
          const double g = left + h * ((cd + off) + cf.xq[q]);
          const double s = cf.Sinv[i * 4 + q] * cf.S[q * 4 + j];
          ma += s * g;
          mabs += s * fabs(g);
        }
      BM[d * 32 + (T & 15)]      = ma;
      BM[d * 32 + 16 + (T & 15)] = mabs;
    }
}

// T1/T2 = Ca_d/Cabs_d U + lifted traces, sweep along the tile's fast (ALONG_A) or slow index; tl/th indexed by the other one
template <bool ALONG_A>
HD_VPT_FN void
vpt6_cpart(const VpTile6Coef &cf, const int d, const double (&U)[4][4], const double (&tl)[4], const double (&th)[4], double (&A)[4][4], double (&B)[4][4])
{
  if (ALONG_A)
    {
      vpt_sweep_a<false>(cf.Ca[d], U, A);
      vpt_sweep_a<false>(cf.Cabs[d], U, B);
    }
  else
    {
      vpt_sweep_b<false>(cf.Ca[d], U, A);
      vpt_sweep_b<false>(cf.Cabs[d], U, B);
    }
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int a = 0; a < 4; ++a)
      {
        const int o = ALONG_A ? a : b, t = ALONG_A ? b : a; // o: index along the sweep direction, t: index of the trace value
        A[b][a] += cf.La0[d][o] * tl[t] + cf.La1[d][o] * th[t];
        B[b][a] += cf.Labs0[d][o] * tl[t] + cf.Labs1[d][o] * th[t];
      }
}

// S(x0) S(x1) on the face tiles (384 tiles: 2 per thread, the second one for T < 128 only)
HD_VPT_FN void
vpt6_face_s01(const VpTile6Coef &cf, double *sm, const int T)
{
  double *BF = sm + 4 * VPT6_BUF;
#pragma unroll 1
  for (int tt = T; tt < 384; tt += 256)
    {
      double *fb = BF + (tt >> 6) * VPT6_FACE;
      double  A[4][4], B[4][4];
      vpt_load_x(fb, tt & 63, A);
      vpt_sweep_a<false>(cf.S, A, B);
      vpt_sweep_b<false>(cf.S, B, A);
      vpt_store_x(fb, tt & 63, A);
    }
}

// S(x2) on the face lines: thread T = (e01, oo): the 4 values x2 = 0..3 at face tile x2 + 4 oo, element e01, of every face
HD_VPT_FN void
vpt6_face_s2(const VpTile6Coef &cf, double *sm, const int T)
{
  double *  BF = sm + 4 * VPT6_BUF;
  const int e01 = T & 15, oo = T >> 4;
#pragma unroll
  for (int f = 0; f < 6; ++f)
    {
      double *q = BF + f * VPT6_FACE + (4 * oo) * VPT_TS + e01;
      double  v[4], o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        v[j] = q[j * VPT_TS];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        o[i] = cf.S[i * 4 + 0] * v[0] + cf.S[i * 4 + 1] * v[1] + cf.S[i * 4 + 2] * v[2] + cf.S[i * 4 + 3] * v[3];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        q[j * VPT_TS] = o[j];
    }
}

// The x-traces (raw values of the x-neighbours, global memory through L2) of a phase are requested one or two phases before
// they are used — with one CTA per SM nothing else hides their latency.  tr[0][.] = lower, tr[1][.] = upper neighbour.
//   direction 2 (phase A, P1 view): layers x2 = 3 / 0 at (e01; v0 = b; w)
//   direction 0 (phase C, P0 view): elements 3 + 4 b / 4 b of the neighbour's tile T
//   direction 1 (phase E, P0 view): elements 12 + a / a
template <typename T_, int D>
HD_VPT_FN void
vpt6_request_traces(const VpParams &p, const Vpt6Cell &C, const int T, double (&tr)[2][4])
{
  const T_ *src = static_cast<const T_ *>(p.src);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    {
      if (D == 2)
        {
          tr[0][i] = double(src[C.nb[2][0] * 4096 + (T & 15) + 48 + 64 * i + 256 * (T >> 4)]);
          tr[1][i] = double(src[C.nb[2][1] * 4096 + (T & 15) + 64 * i + 256 * (T >> 4)]);
        }
      else if (D == 0)
        {
          tr[0][i] = double(src[C.nb[0][0] * 4096 + 16 * T + 3 + 4 * i]);
          tr[1][i] = double(src[C.nb[0][1] * 4096 + 16 * T + 4 * i]);
        }
      else
        {
          tr[0][i] = double(src[C.nb[1][0] * 4096 + 16 * T + 12 + i]);
          tr[1][i] = double(src[C.nb[1][1] * 4096 + 16 * T + i]);
        }
    }
}

HD_VPT_FN void
vpt6_phaseA(const VpTile6Coef &cf, double *sm, const int T, const double (&tr)[2][4])
{
  double *B0 = sm, *B1 = sm + VPT6_BUF, *B2 = sm + 2 * VPT6_BUF;
  double  U[4][4], A[4][4], B[4][4];
  const double(&tl)[4] = tr[0], (&th)[4] = tr[1];
  vpt6_load<1>(B0, T, U);
  vpt6_cpart<true>(cf, 2, U, tl, th, A, B); // x2 = the fast index of a P1 tile
  vpt6_store<1>(B1, T, A);
  vpt6_store<1>(B2, T, B);
  vpt6_face_s01(cf, sm, T);
}

// OUT (+)= Ma_d T1 + Mabs_d T2 in view P, sweep along the fast (ALONG_A) or the slow index
template <int P, bool ALONG_A, bool ADD>
HD_VPT_FN void
vpt6_mpart(double *sm, const int d, const int T)
{
  double *B1 = sm + VPT6_BUF, *B2 = sm + 2 * VPT6_BUF, *B3 = sm + 3 * VPT6_BUF, *BM = sm + VPT6_BM + d * 32;
  double  A[4][4], O[4][4], M[16];
  if (ADD)
    vpt6_load<P>(B3, T, O);
#pragma unroll
  for (int i = 0; i < 16; ++i)
    M[i] = BM[i];
  vpt6_load<P>(B1, T, A);
  if (ALONG_A)
    {
      if (ADD)
        vpt_sweep_a<true>(M, A, O);
      else
        vpt_sweep_a<false>(M, A, O);
    }
  else
    {
      if (ADD)
        vpt_sweep_b<true>(M, A, O);
      else
        vpt_sweep_b<false>(M, A, O);
    }
#pragma unroll
  for (int i = 0; i < 16; ++i)
    M[i] = BM[16 + i];
  vpt6_load<P>(B2, T, A);
  if (ALONG_A)
    vpt_sweep_a<true>(M, A, O);
  else
    vpt_sweep_b<true>(M, A, O);
  vpt6_store<P>(B3, T, O);
}

// phases C and E share the thread's U tile (registers)
HD_VPT_FN void
vpt6_phaseC(const VpTile6Coef &cf, double *sm, const int T, double (&U)[4][4], const double (&tr)[2][4])
{
  double *B0 = sm, *B1 = sm + VPT6_BUF, *B2 = sm + 2 * VPT6_BUF;
  double  A[4][4], B[4][4];
  const double(&tl)[4] = tr[0], (&th)[4] = tr[1];
  vpt6_load<0>(B0, T, U);
  vpt6_cpart<true>(cf, 0, U, tl, th, A, B);
  vpt6_store<0>(B1, T, A);
  vpt6_store<0>(B2, T, B);
  vpt_sweep_a<false>(cf.S, U, A);
  vpt_sweep_b<false>(cf.S, A, B);
  vpt6_store<0>(B0, T, B);
}

HD_VPT_FN void
vpt6_phaseE(const VpTile6Coef &cf, double *sm, const int T, const double (&U)[4][4], const double (&tr)[2][4])
{
  double *B1 = sm + VPT6_BUF, *B2 = sm + 2 * VPT6_BUF;
  double  A[4][4], B[4][4];
  const double(&tl)[4] = tr[0], (&th)[4] = tr[1];
  vpt6_cpart<false>(cf, 1, U, tl, th, A, B);
  vpt6_store<0>(B1, T, A);
  vpt6_store<0>(B2, T, B);
}

HD_VPT_FN void
vpt6_phaseG(const VpParams &p, const VpTile6Coef &cf, const Vpt6Cell &C, double *sm, const int T)
{
  double *  B0 = sm, *B1 = sm + VPT6_BUF, *BF = sm + 4 * VPT6_BUF;
  const int e01 = T & 15, w = T >> 4;
  double    W01[4][4], W[4][4], A[4][4], B[4][4];
  vpt6_load<1>(B0, T, W01);
  vpt_sweep_a<false>(cf.S, W01, W); // S(x2): x2 = the fast index
  vpt6_store<1>(B0, T, W);
  vpt_sweep_b<false>(cf.Ca[3], W, A); // v0 = the slow index
  vpt_sweep_b<false>(cf.Cabs[3], W, B);
  const long long cx = C.c[0] + (long long)p.ncell[0] * (C.c[1] + (long long)p.ncell[1] * C.c[2]);
#pragma unroll
  for (int a = 0; a < 4; ++a)
    {
      const double g  = p.a_v[(cx * 64 + e01 + 16 * a) * 3 + 0], ga = fabs(g);
      const double fl = BF[0 * VPT6_FACE + (a + 4 * w) * VPT_TS + e01], fh = BF[1 * VPT6_FACE + (a + 4 * w) * VPT_TS + e01];
#pragma unroll
      for (int b = 0; b < 4; ++b)
        A[b][a] = g * (A[b][a] + cf.La0[3][b] * fl + cf.La1[3][b] * fh) + ga * (B[b][a] + cf.Labs0[3][b] * fl + cf.Labs1[3][b] * fh);
    }
  vpt6_store<1>(B1, T, A);
}

HD_VPT_FN void
vpt6_phaseH(const VpParams &p, const VpTile6Coef &cf, const Vpt6Cell &C, double *sm, const int T)
{
  double *  B0 = sm, *B1 = sm + VPT6_BUF, *BF = sm + 4 * VPT6_BUF;
  const int e01 = T & 15, x2 = (T >> 4) & 3, v0 = T >> 6;
  double    W[4][4], R[4][4], M[16];
  vpt6_load<2>(B0, T, W);
  vpt6_load<2>(B1, T, R);
  const long long cx = C.c[0] + (long long)p.ncell[0] * (C.c[1] + (long long)p.ncell[1] * C.c[2]);
  const double    g1 = p.a_v[(cx * 64 + (T & 63)) * 3 + 1], g2 = p.a_v[(cx * 64 + (T & 63)) * 3 + 2];
  const double    a1 = fabs(g1), a2 = fabs(g2);
  // v1 = the fast index: faces 2 (lower), 3 (upper) of the v1-neighbours at (x..; v0; v2 = b)
#pragma unroll
  for (int i = 0; i < 16; ++i)
    M[i] = g1 * cf.Ca[4][i] + a1 * cf.Cabs[4][i];
  vpt_sweep_a<true>(M, W, R);
#pragma unroll
  for (int b = 0; b < 4; ++b)
    {
      const double fl = BF[2 * VPT6_FACE + (x2 + 4 * v0 + 16 * b) * VPT_TS + e01], fh = BF[3 * VPT6_FACE + (x2 + 4 * v0 + 16 * b) * VPT_TS + e01];
#pragma unroll
      for (int a = 0; a < 4; ++a)
        R[b][a] += (g1 * cf.La0[4][a] + a1 * cf.Labs0[4][a]) * fl + (g1 * cf.La1[4][a] + a1 * cf.Labs1[4][a]) * fh;
    }
  // v2 = the slow index: faces 4, 5 at (x..; v0; v1 = a)
#pragma unroll
  for (int i = 0; i < 16; ++i)
    M[i] = g2 * cf.Ca[5][i] + a2 * cf.Cabs[5][i];
  vpt_sweep_b<true>(M, W, R);
#pragma unroll
  for (int a = 0; a < 4; ++a)
    {
      const double fl = BF[4 * VPT6_FACE + (x2 + 4 * v0 + 16 * a) * VPT_TS + e01], fh = BF[5 * VPT6_FACE + (x2 + 4 * v0 + 16 * a) * VPT_TS + e01];
#pragma unroll
      for (int b = 0; b < 4; ++b)
        R[b][a] += (g2 * cf.La0[5][b] + a2 * cf.Labs0[5][b]) * fl + (g2 * cf.La1[5][b] + a2 * cf.Labs1[5][b]) * fh;
    }
  vpt6_store<2>(B1, T, R);
}

HD_VPT_FN void
vpt6_phaseI(const VpTile6Coef &cf, double *sm, const int T)
{
  double *B1 = sm + VPT6_BUF;
  double  R[4][4], A[4][4];
  vpt6_load<1>(B1, T, R);
  vpt_sweep_a<false>(cf.Sinv, R, A);
  vpt6_store<1>(B1, T, A);
}

HD_VPT_FN void
vpt6_phaseJ(const VpTile6Coef &cf, double *sm, const int T)
{
  double *B1 = sm + VPT6_BUF, *B2 = sm + 2 * VPT6_BUF, *B3 = sm + 3 * VPT6_BUF;
  double  R[4][4], A[4][4], O[4][4];
  vpt6_load<0>(B1, T, R);
  vpt6_load<0>(B3, T, O);
  vpt_sweep_a<false>(cf.Sinv, R, A);
  vpt_sweep_b<true>(cf.Sinv, A, O);
  vpt6_store<0>(B2, T, O);
}

// fused LSRK update: the thread's `sol` values are requested two phases before the store (phase I), see phase K
template <typename T_>
HD_VPT_FN void
vpt6_request_sol(const VpParams &p, const Vpt6Cell &C, const int T, double (&sv)[8][2])
{
  const T_ *sol = static_cast<const T_ *>(p.sol);
  if (p.fused)
    {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        vpt_load2(sol + C.cell * 4096 + 2 * (256 * i + T), sv[i][0], sv[i][1]);
    }
}

template <typename T_>
HD_VPT_FN void
vpt6_phaseK(const VpParams &p, const Vpt6Cell &C, const double *sm, const int T, const double (&sv)[8][2])
{
  const double *  B2  = sm + 2 * VPT6_BUF;
  T_ *            dst = static_cast<T_ *>(p.dst), *sol = static_cast<T_ *>(p.sol), *tin = static_cast<T_ *>(p.ti_next);
  const long long g0  = C.cell * 4096;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    {
      const int k = 256 * i + T;
      double    v0, v1;
      vpt_sm_load2(B2 + (k >> 3) * VPT_TS + 2 * (k & 7), v0, v1);
      if (p.fused)
        {
          vpt_store2(sol + g0 + 2 * k, sv[i][0] + p.fb * v0, sv[i][1] + p.fb * v1);
          if (p.fa != 0.0)
            vpt_store2(tin + g0 + 2 * k, sv[i][0] + p.fa * v0, sv[i][1] + p.fa * v1);
        }
      else
        vpt_store2(dst + g0 + 2 * k, v0, v1);
    }
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
// 3D3V: one cell per CTA of 256 threads
template <typename T_>
__global__ void __launch_bounds__(256, 1) k_vp_tile_3d3v(const VpParams p, const __grid_constant__ VpTile6Coef cf)
{
  extern __shared__ double sm[];
  const int T = threadIdx.x;
  Vpt6Cell  C;
  vpt6_decode(p, blockIdx.x, C);
  double trA[2][4], trC[2][4], trE[2][4], sv[8][2];
  vpt6_request_traces<T_, 2>(p, C, T, trA);
  vpt6_phase0<T_>(p, cf, C, sm, T);
  __syncthreads();
  vpt6_request_traces<T_, 0>(p, C, T, trC);
  vpt6_phaseA(cf, sm, T, trA);
  __syncthreads();
  vpt6_mpart<2, false, false>(sm, 2, T); // B: OUT = Ma_2(v2) ..., v2 = the slow index of a P2 tile
  vpt6_face_s2(cf, sm, T);
  __syncthreads();
  double U[4][4];
  vpt6_request_traces<T_, 1>(p, C, T, trE);
  vpt6_phaseC(cf, sm, T, U, trC);
  __syncthreads();
  vpt6_mpart<1, false, true>(sm, 0, T); // D: v0 = the slow index of a P1 tile
  __syncthreads();
  vpt6_phaseE(cf, sm, T, U, trE);
  __syncthreads();
  vpt6_mpart<2, true, true>(sm, 1, T); // F: v1 = the fast index of a P2 tile
  __syncthreads();
  vpt6_phaseG(p, cf, C, sm, T);
  __syncthreads();
  vpt6_phaseH(p, cf, C, sm, T);
  __syncthreads();
  vpt6_request_sol<T_>(p, C, T, sv);
  vpt6_phaseI(cf, sm, T);
  __syncthreads();
  vpt6_phaseJ(cf, sm, T);
  __syncthreads();
  vpt6_phaseK<T_>(p, C, sm, T, sv);
}
#endif
