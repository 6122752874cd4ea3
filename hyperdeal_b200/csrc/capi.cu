// C ABI of libhdgpu.so (see include/hyperdeal_b200.h for the contract and the reference
// members each entry point replaces).
#include <cstdlib>
#include <cstring>
#include <new>

#include "hd_internal.h"
#include "rounds6d_tasks.cuh"

namespace hd
{
  static thread_local std::string g_error;

  int
  fail(int code, const std::string &msg)
  {
    g_error = msg;
    return code;
  }
} // namespace hd

// ------------------------------------------------------------------------------------------
// small device kernels: LSRK update, halo pack, interpolate, norms
// ------------------------------------------------------------------------------------------
namespace
{
  // perform_stage update loops, base/time_integrators.templates.h:117-132
  template <typename T>
  __global__ void
  k_stage_update(T *__restrict__ sol, T *__restrict__ ti, const T *__restrict__ K, T b, T a, long long n)
  {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      {
        const T k = K[i], s = sol[i];
        sol[i] = s + b * k;
        if (a != T(0))
          ti[i] = s + a * k;
      }
  }

  struct LatticeParams
  {
    int       dim, n, nq;
    int       ncell[HD_MAX_DIM], cell_offset[HD_MAX_DIM];
    double    left[HD_MAX_DIM], h[HD_MAX_DIM];
    long long nd, ncells;
  };

  // pack loop of export_to_ghosted_array_start, matrix_free/vector_partitioner.h:1443-1460: gather the nodal face layer
  // `side` of direction `dir` of every boundary cell into a contiguous segment.  One thread moves VEC consecutive values
  // (16 bytes whenever the layer is contiguous over at least that much, i.e. for dir >= 1); `out` may be a peer-mapped
  // pointer (the neighbour GPU's ghost segment): the stores then go over NVLink directly.
  template <typename T, int VEC>
  __global__ void __launch_bounds__(128)
    k_halo_pack(const T *__restrict__ src, T *__restrict__ out, LatticeParams lp, int dir, int side, int n_face_cells, int nf, int stride_d, int *started)
  {
    // "this CTA is resident": lets the host gate the launch of the persistent operator kernel behind the start of the
    // pack kernel (a kernel launched after the operator kernel is resident would not get an SM until that one ends)
    if (started && threadIdx.x == 0)
      atomicAdd(started, 1);
    struct alignas(sizeof(T) * VEC) Chunk
    {
      T v[VEC];
    };
    const int layer_off = (side ? lp.n - 1 : 0) * stride_d;
    const int hi_stride = lp.n * stride_d;
    for (int fc = blockIdx.x; fc < n_face_cells; fc += gridDim.x)
      {
        // face cell -> cell (uniform over the CTA)
        long long cell = 0, m = 1;
        int       r    = fc;
        for (int e = 0; e < lp.dim; ++e)
          {
            int ce;
            if (e == dir)
              ce = side ? lp.ncell[e] - 1 : 0;
            else
              {
                ce = r % lp.ncell[e];
                r /= lp.ncell[e];
              }
            cell += ce * m;
            m *= lp.ncell[e];
          }
        const T *s = src + cell * lp.nd + layer_off;
        T *      o = out + (long long)fc * nf;
#pragma unroll 4
        for (int i = threadIdx.x * VEC; i < nf; i += 128 * VEC)
          {
            const int hi = i / stride_d, lo = i - hi * stride_d;
            *reinterpret_cast<Chunk *>(o + i) = *reinterpret_cast<const Chunk *>(s + hi * hi_stride + lo);
          }
      }
  }

  // The built-in fields are products of 1-D factors: value(x) = prod_d builtin_factor(d, x_d), multiplied in the order
  // d = 0, 1, ... (examples/advection/cases/hyperrectangle.h:46-57: sin in direction 0, cos in the others).  The kernels below evaluate the factors once per cell and direction.
  __device__ double
  builtin_factor(int fn_id, int d, double x, double t)
  {
    if (fn_id == HD_FN_HYPERRECTANGLE)
      {
        const double adv[6] = {1.0, 0.15, -0.05, 0.0, 0.0, 0.0};
        const double PI     = 3.14159265358979323846;
        const double arg    = 2.0 * (x - t * adv[d]) * PI;
        return d == 0 ? sin(arg) : cos(arg);
      }
    return 0.0;
  }

  // VectorTools::interpolate, numerics/vector_tools.h:88-137
  template <typename T>
  __global__ void
  k_interpolate(T *__restrict__ vec, LatticeParams lp, const double *__restrict__ nodes, int fn_id, double time, int cells_per_cta)
  {
    // a CTA takes cells_per_cta consecutive cells: 1-D factor tables [cell][direction][node] in shared memory, then one
    // product of dim table entries per nodal value (coalesced stores)
    extern __shared__ double tab[];
    const int       dim = lp.dim, n = lp.n;
    const long long cell0 = (long long)blockIdx.x * cells_per_cta;
    long long       nloc  = lp.ncells - cell0;
    if (nloc > cells_per_cta)
      nloc = cells_per_cta;
    for (int e = threadIdx.x; e < nloc * dim * n; e += blockDim.x)
      {
        const int lc = e / (dim * n), d = (e / n) % dim, id = e % n;
        long long r  = cell0 + lc;
        for (int k = 0; k < d; ++k)
          r /= lp.ncell[k];
        const int    c = int(r % lp.ncell[d]);
        const double x = lp.left[d] + lp.h[d] * ((c + lp.cell_offset[d]) + nodes[id]);
        tab[e]         = builtin_factor(fn_id, d, x, time);
      }
    __syncthreads();
    const long long total = nloc * lp.nd;
    for (long long i = threadIdx.x; i < total; i += blockDim.x)
      {
        const int     lc = int(i / lp.nd);
        long long     o  = i - lc * lp.nd;
        const double *t  = tab + lc * dim * n;
        double        r  = t[o % n];
        o /= n;
        for (int d = 1; d < dim; ++d)
          {
            r *= t[d * n + o % n];
            o /= n;
          }
        vec[cell0 * lp.nd + i] = T(r);
      }
  }

  // VectorTools::norm_and_error, numerics/vector_tools.h:151-220: one CTA per cell, S sweeps in
  // shared memory (double), then sum over the quadrature points.
  template <typename T>
  __global__ void
  k_norm_error(const T *__restrict__ vec, LatticeParams lp, const double *__restrict__ basis, int fn_id, double time, double *__restrict__ out)
  {
    extern __shared__ double sm[];
    const int     dim = lp.dim, n = lp.n, nq = lp.nq;
    const double *xq = basis + n, *w = xq + nq, *S = w + nq;
    int           mx = n > nq ? n : nq;
    long long     cap = 1;
    for (int d = 0; d < dim; ++d)
      cap *= mx;
    double *        A = sm, *B = sm + cap;
    const long long cell = blockIdx.x;
    for (long long i = threadIdx.x; i < lp.nd; i += blockDim.x)
      A[i] = double(vec[cell * lp.nd + i]);
    __syncthreads();
    double *in = A, *outb = B;
    for (int d = 0; d < dim; ++d)
      {
        long long stride = 1;
        for (int e = 0; e < d; ++e)
          stride *= nq;
        long long outer = 1;
        for (int e = d + 1; e < dim; ++e)
          outer *= n;
        const long long total = outer * nq * stride;
        for (long long i = threadIdx.x; i < total; i += blockDim.x)
          {
            const long long lo = i % stride, rest = i / stride, q = rest % nq, o = rest / nq;
            double          acc = 0;
            for (int k = 0; k < n; ++k)
              acc += S[q * n + k] * in[(o * n + k) * stride + lo];
            outb[i] = acc;
          }
        __syncthreads();
        double *t = in;
        in        = outb;
        outb      = t;
      }
    long long nqd = 1;
    for (int d = 0; d < dim; ++d)
      nqd *= nq;
    int       c[HD_MAX_DIM];
    long long r = cell;
    for (int d = 0; d < dim; ++d)
      {
        c[d] = int(r % lp.ncell[d]);
        r /= lp.ncell[d];
      }
    // 1-D factors of the analytic field at the quadrature points of this cell (the sweeps are done: outb is free)
    double *ftab = outb;
    for (int e = threadIdx.x; e < dim * nq; e += blockDim.x)
      {
        const int d = e / nq, q = e % nq;
        ftab[e]     = builtin_factor(fn_id, d, lp.left[d] + lp.h[d] * ((c[d] + lp.cell_offset[d]) + xq[q]), time);
      }
    __syncthreads();
    double s_norm = 0, s_err = 0;
    for (long long i = threadIdx.x; i < nqd; i += blockDim.x)
      {
        long long rr = i;
        double    jxw = 1, f = 1;
        for (int d = 0; d < dim; ++d)
          {
            const int q = int(rr % nq);
            rr /= nq;
            f   = d == 0 ? ftab[q] : f * ftab[d * nq + q];
            jxw *= lp.h[d] * w[q];
          }
        const double u = in[i];
        s_norm += u * u * jxw;
        s_err += (u - f) * (u - f) * jxw;
      }
    // block reduction
    __shared__ double red[2][32];
    for (int off = 16; off > 0; off >>= 1)
      {
        s_norm += __shfl_down_sync(0xffffffffu, s_norm, off);
        s_err += __shfl_down_sync(0xffffffffu, s_err, off);
      }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0)
      {
        red[0][wid] = s_norm;
        red[1][wid] = s_err;
      }
    __syncthreads();
    if (wid == 0)
      {
        const int nw = (blockDim.x + 31) / 32;
        s_norm       = lane < nw ? red[0][lane] : 0.0;
        s_err        = lane < nw ? red[1][lane] : 0.0;
        for (int off = 16; off > 0; off >>= 1)
          {
            s_norm += __shfl_down_sync(0xffffffffu, s_norm, off);
            s_err += __shfl_down_sync(0xffffffffu, s_err, off);
          }
        if (lane == 0)
          {
            atomicAdd(out + 0, s_norm);
            atomicAdd(out + 1, s_err);
          }
      }
  }

  // ---- streaming versions of the two VectorTools kernels for degree 3 (n = 4 nodes per direction) ------------------------
  // k_interpolate4: a thread writes whole lines of 4 nodal values along x_0 (32 bytes, two 16-byte stores; a warp writes
  // 1 KiB contiguously): value = t_0[i_0] * prod_{d>=1} t_d[i_d] with the 1-D factor tables of the cell in shared memory
  // and all index arithmetic on compile-time extents.  Write-only, 8 B/DoF.
  template <typename T, int DIM>
  __global__ void __launch_bounds__(256)
    k_interpolate4(T *__restrict__ vec, LatticeParams lp, const double *__restrict__ nodes, int fn_id, double time, int cells_per_cta)
  {
    extern __shared__ double tab[]; // [cell][direction][node]
    constexpr int   N = 4, LINES = 1 << (2 * (DIM - 1));
    const long long cell0 = (long long)blockIdx.x * cells_per_cta;
    long long       nloc  = lp.ncells - cell0;
    if (nloc > cells_per_cta)
      nloc = cells_per_cta;
    for (int e = threadIdx.x; e < nloc * DIM * N; e += blockDim.x)
      {
        const int lc = e / (DIM * N), d = (e / N) % DIM, id = e % N;
        long long r  = cell0 + lc;
        for (int k = 0; k < d; ++k)
          r /= lp.ncell[k];
        const int    c = int(r % lp.ncell[d]);
        const double x = lp.left[d] + lp.h[d] * ((c + lp.cell_offset[d]) + nodes[id]);
        tab[e]         = builtin_factor(fn_id, d, x, time);
      }
    __syncthreads();
    const int total = int(nloc) * LINES;
    for (int i = threadIdx.x; i < total; i += blockDim.x)
      {
        const int     lc = i / LINES, line = i % LINES;
        const double *t  = tab + lc * DIM * N;
        double        f  = 1.0;
#pragma unroll
        for (int d = 1; d < DIM; ++d)
          f *= t[d * N + ((line >> (2 * (d - 1))) & 3)];
        // same order of multiplications as the generic kernel: ((t0 * t1) * t2) ... — so that the values are bit-identical
        double v[N];
#pragma unroll
        for (int a = 0; a < N; ++a)
          {
            double r = t[a];
#pragma unroll
            for (int d = 1; d < DIM; ++d)
              r *= t[d * N + ((line >> (2 * (d - 1))) & 3)];
            v[a] = r;
          }
        (void)f;
        T *o = vec + (cell0 + lc) * (long long)(LINES * N) + (long long)line * N;
        if (sizeof(T) == 8)
          {
            reinterpret_cast<double2 *>(o)[0] = make_double2(v[0], v[1]);
            reinterpret_cast<double2 *>(o)[1] = make_double2(v[2], v[3]);
          }
        else
          *reinterpret_cast<float4 *>(o) = make_float4(float(v[0]), float(v[1]), float(v[2]), float(v[3]));
      }
  }

  // k_norm_error_3d3v: 3D3V, n = n_q = 4.  The cell sits in shared memory in the swizzled row layout of the three-round
  // operator kernel (rounds6d_tasks.cuh) and the six S sweeps (nodal values -> quadrature points) run as three passes of
  // two directions on a 4x4 register tile per thread, with that kernel's conflict-free thread maps; the third pass keeps
  // its tile in registers and goes straight into the quadrature sums.  A CTA walks over cells and reduces once at the end.
  // Read-only, 8 B/DoF; 24 FMA per value for the sweeps.
  template <typename T>
  __global__ void __launch_bounds__(256)
    k_norm_error_3d3v(const T *__restrict__ vec, LatticeParams lp, const double *__restrict__ basis, int fn_id, double time, double *__restrict__ out)
  {
    __shared__ __align__(1024) unsigned char cellbuf[32768];
    __shared__ double                        ftab[6][4], wtab[6][4], Ssm[16];
    __shared__ double                        red[2][8];
    const int      tid = threadIdx.x, t = tid & 127, j = tid >> 7;
    const uint32_t ub  = (uint32_t)__cvta_generic_to_shared(cellbuf);
    const double * xq = basis + 4, *w = xq + 4, *S = w + 4;
    if (tid < 16)
      Ssm[tid] = S[tid];
    if (tid < 24)
      wtab[tid / 4][tid % 4] = lp.h[tid / 4] * w[tid % 4];
    r6::ThreadMap<0> tm0;
    r6::ThreadMap<1> tm1;
    r6::ThreadMap<2> tm2;
    tm0.init(t);
    tm1.init(t);
    tm2.init(t);
    double s_norm = 0.0, s_err = 0.0;
    // y[b][a] = sum_j S[a][j] x[b][j], then z[b][a] = sum_j S[b][j] y[j][a]
    auto sweep2 = [&](double(&x)[4][4]) {
      double y[4][4];
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int a = 0; a < 4; ++a)
          y[b][a] = Ssm[a * 4 + 0] * x[b][0] + Ssm[a * 4 + 1] * x[b][1] + Ssm[a * 4 + 2] * x[b][2] + Ssm[a * 4 + 3] * x[b][3];
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int a = 0; a < 4; ++a)
          x[b][a] = Ssm[b * 4 + 0] * y[0][a] + Ssm[b * 4 + 1] * y[1][a] + Ssm[b * 4 + 2] * y[2][a] + Ssm[b * 4 + 3] * y[3][a];
    };
    for (long long cell = blockIdx.x; cell < lp.ncells; cell += gridDim.x)
      {
        __syncthreads(); // the previous cell's tables and buffer are no longer read
        if (tid < 24)
          {
            const int d = tid / 4, q = tid % 4;
            long long r = cell;
            for (int k = 0; k < d; ++k)
              r /= lp.ncell[k];
            const int c = int(r % lp.ncell[d]);
            ftab[d][q]  = builtin_factor(fn_id, d, lp.left[d] + lp.h[d] * ((c + lp.cell_offset[d]) + xq[q]), time);
          }
        // coalesced load of the cell into the swizzled layout: chunk g (16 bytes) of row g / 8
#pragma unroll
        for (int m = 0; m < 8; ++m)
          {
            const int      g   = tid + 256 * m;
            const uint32_t row = uint32_t(g >> 3), ch = uint32_t(g & 7);
            const T *      sp  = vec + cell * 4096 + 2 * g;
            r6_sts128(ub + row * 128u + ((ch ^ (row & 7u)) << 4), double(sp[0]), double(sp[1]));
          }
        __syncthreads();
        double x[4][4];
        // pass 0: directions (0,1) — the thread's row
        {
          const uint32_t jo = uint32_t(j) * 16384u;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            {
              const double2 v            = r6_lds128(ub + tm0.x[ch] + jo);
              x[ch >> 1][(ch & 1) * 2]     = v.x;
              x[ch >> 1][(ch & 1) * 2 + 1] = v.y;
            }
          sweep2(x);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            r6_sts128(ub + tm0.x[ch] + jo, x[ch >> 1][(ch & 1) * 2], x[ch >> 1][(ch & 1) * 2 + 1]);
        }
        __syncthreads();
        // pass 1: directions (2,3)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int a = 0; a < 4; ++a)
            x[b][a] = r6_lds64(ub + tm1.elem(a, b, j));
        sweep2(x);
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int a = 0; a < 4; ++a)
            r6_sts64(ub + tm1.elem(a, b, j), x[b][a]);
        __syncthreads();
        // pass 2: directions (4,5), then the quadrature sums at the points (q0,q1 | q2,q3 | a, b)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int a = 0; a < 4; ++a)
            x[b][a] = r6_lds64(ub + tm2.elem(a, b, j));
        sweep2(x);
        const int    q0 = t & 3, q1 = (t >> 2) & 3, E = (t >> 4) + 8 * j, q2 = E & 3, q3 = E >> 2;
        const double wc = ((wtab[0][q0] * wtab[1][q1]) * wtab[2][q2]) * wtab[3][q3];
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int a = 0; a < 4; ++a)
            {
              // same multiplication order as the generic kernel: f = (((((f0 f1) f2) f3) f4) f5), jxw likewise
              const double f   = ((((ftab[0][q0] * ftab[1][q1]) * ftab[2][q2]) * ftab[3][q3]) * ftab[4][a]) * ftab[5][b];
              const double jxw = (wc * wtab[4][a]) * wtab[5][b];
              const double u   = x[b][a];
              s_norm += u * u * jxw;
              s_err += (u - f) * (u - f) * jxw;
            }
      }
    for (int off = 16; off > 0; off >>= 1)
      {
        s_norm += __shfl_down_sync(0xffffffffu, s_norm, off);
        s_err += __shfl_down_sync(0xffffffffu, s_err, off);
      }
    if ((tid & 31) == 0)
      {
        red[0][tid >> 5] = s_norm;
        red[1][tid >> 5] = s_err;
      }
    __syncthreads();
    if (tid == 0)
      {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < 8; ++i)
          {
            a += red[0][i];
            b += red[1][i];
          }
        atomicAdd(out + 0, a);
        atomicAdd(out + 1, b);
      }
  }

  // VectorTools::velocity_space_integration, numerics/vector_tools.h:238-315 with quad_no_v = 2 (Gauss-Lobatto = the nodes):
  // rho[x-cell][x-node] = sum over v-cells and v-nodes of f * JxW_v.  Thread = one x-space value (coalesced over the x-nodes
  // of a cell), blockIdx.y = a range of v-cells; partial sums are accumulated in double and added atomically.
  template <typename T>
  __global__ void __launch_bounds__(256)
    k_velocity_space_integration(const T *__restrict__ f, T *__restrict__ rho, const double *__restrict__ wv, long long n_x_dofs, long long ncx, long long ncv, int ndx, int ndv)
  {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_x_dofs)
      return;
    const long long xc = i / ndx;
    const int       xn = int(i - xc * ndx);
    const long long per = (ncv + gridDim.y - 1) / gridDim.y;
    const long long v0 = blockIdx.y * per, v1 = v0 + per < ncv ? v0 + per : ncv;
    double          acc = 0.0;
    for (long long vc = v0; vc < v1; ++vc)
      {
        const T *p = f + ((vc * ncx + xc) * ndv) * ndx + xn;
#pragma unroll 4
        for (int vn = 0; vn < ndv; ++vn)
          acc += double(p[(long long)vn * ndx]) * wv[vn];
      }
    if (v1 > v0)
      atomicAdd(rho + i, T(acc));
  }

  LatticeParams
  lattice(const hd_mesh *m)
  {
    LatticeParams lp;
    lp.dim = m->dim;
    lp.n   = m->n;
    lp.nq  = m->nq;
    for (int d = 0; d < HD_MAX_DIM; ++d)
      {
        lp.ncell[d]       = d < m->dim ? m->d.n_cells[d] : 1;
        lp.cell_offset[d] = d < m->dim ? m->d.cell_offset[d] : 0;
        lp.left[d]        = m->d.left[d];
        lp.h[d]           = m->h[d];
      }
    lp.nd     = m->nd;
    lp.ncells = m->ncells;
    return lp;
  }

  unsigned
  grid_for(const hd_context *ctx, long long n, int threads)
  {
    long long blocks = (n + threads - 1) / threads;
    long long cap    = (long long)ctx->sm_count * 16;
    if (blocks > cap)
      blocks = cap;
    if (blocks < 1)
      blocks = 1;
    return (unsigned)blocks;
  }
} // namespace

// ------------------------------------------------------------------------------------------
extern "C" {

const char *
hd_last_error(void)
{
  return hd::g_error.c_str();
}

int
hd_version(void)
{
  return 100;
}

// plain device buffers for host layers that do not link the CUDA runtime themselves (e.g. the gradient table of hd_poisson_solve)
int
hd_device_malloc(hd_context *ctx, size_t bytes, void **ptr)
{
  HD_REQUIRE(ctx && ptr, "null argument");
  HD_CUDA(cudaSetDevice(ctx->device));
  HD_CUDA(cudaMalloc(ptr, bytes ? bytes : 16));
  HD_CUDA(cudaMemsetAsync(*ptr, 0, bytes, ctx->stream));
  return HD_OK;
}

int
hd_device_free(hd_context *ctx, void *ptr)
{
  HD_REQUIRE(ctx, "null argument");
  HD_CUDA(cudaSetDevice(ctx->device));
  HD_CUDA(cudaFree(ptr));
  return HD_OK;
}

int
hd_device_count(int *count)
{
  HD_REQUIRE(count, "null argument");
  *count = 0;
  HD_CUDA(cudaGetDeviceCount(count));
  return HD_OK;
}

int
hd_context_create(int device, hd_context **out)
{
  HD_REQUIRE(out, "null argument");
  int count = 0;
  HD_CUDA(cudaGetDeviceCount(&count));
  if (count == 0)
    return hd::fail(HD_ERR_CUDA, "no CUDA device: libhdgpu has no CPU fallback");
  HD_REQUIRE(device >= 0 && device < count, "device index out of range");
  HD_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  HD_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return hd::fail(HD_ERR_CUDA, "libhdgpu is built for sm_100a (Blackwell) only");
  hd_context *ctx = new (std::nothrow) hd_context;
  HD_REQUIRE(ctx, "out of memory");
  ctx->device     = device;
  ctx->sm_count   = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  HD_CUDA(cudaEventCreate(&ctx->ev0));
  HD_CUDA(cudaEventCreate(&ctx->ev1));
  if (const char *e = getenv("HD_PERSIST_L2_MB")) // experiment knob: L2 set-aside for evict_last lines (see HD_L2_HINTS)
    {
      size_t want = (size_t)atoll(e) << 20;
      if (want > (size_t)prop.persistingL2CacheMaxSize)
        want = (size_t)prop.persistingL2CacheMaxSize;
      HD_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
    }
  *out = ctx;
  return HD_OK;
}

int
hd_context_destroy(hd_context *ctx)
{
  if (!ctx)
    return HD_OK;
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  delete ctx;
  return HD_OK;
}

int
hd_context_set_stream(hd_context *ctx, void *stream)
{
  HD_REQUIRE(ctx, "null context");
  ctx->stream = static_cast<cudaStream_t>(stream);
  return HD_OK;
}

int
hd_context_synchronize(hd_context *ctx)
{
  HD_REQUIRE(ctx, "null context");
  HD_CUDA(cudaStreamSynchronize(ctx->stream));
  return HD_OK;
}

int
hd_timer_start(hd_context *ctx)
{
  HD_REQUIRE(ctx, "null context");
  HD_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  return HD_OK;
}

int
hd_timer_stop(hd_context *ctx, double *ms)
{
  HD_REQUIRE(ctx && ms, "null argument");
  HD_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  HD_CUDA(cudaEventSynchronize(ctx->ev1));
  float f = 0;
  HD_CUDA(cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
  *ms = f;
  return HD_OK;
}

// ---- mesh ---------------------------------------------------------------------------------
static int mesh_device_setup(hd_mesh *m, const std::vector<double> &hb);

int
hd_mesh_create(hd_context *ctx, const hd_mesh_desc *desc, hd_mesh **out)
{
  HD_REQUIRE(ctx && desc && out, "null argument");
  const int dim = desc->dim_x + desc->dim_v;
  HD_REQUIRE(desc->dim_x >= 1 && desc->dim_x <= 3 && desc->dim_v >= 1 && desc->dim_v <= 3, "dim_x and dim_v must be in 1..3");
  HD_REQUIRE(desc->degree >= 1 && desc->degree <= 5, "degree must be in 1..5");
  HD_REQUIRE(desc->n_points >= desc->degree + 1 && desc->n_points <= 8, "n_points must be in degree+1..8");
  HD_REQUIRE(desc->number_type == HD_F64 || desc->number_type == HD_F32, "number_type");
  HD_REQUIRE(!desc->collocation || desc->n_points == desc->degree + 1, "collocation requires n_points == degree+1");
  hd_mesh *m = new (std::nothrow) hd_mesh;
  HD_REQUIRE(m, "out of memory");
  m->ctx       = ctx;
  m->d         = *desc;
  m->dim       = dim;
  m->n         = desc->degree + 1;
  m->nq        = desc->n_points;
  m->elem_size = desc->number_type == HD_F64 ? 8 : 4;
  m->nd        = 1;
  m->ncells    = 1;
  for (int d = 0; d < HD_MAX_DIM; ++d)
    {
      m->h[d] = 1.0;
      for (int s = 0; s < 2; ++s)
        m->ghost_off[d][s] = m->ghost_cnt[d][s] = 0;
    }
  for (int d = 0; d < dim; ++d)
    {
      if (!(desc->n_cells[d] >= 1 && desc->n_cells_global[d] >= desc->n_cells[d] && desc->cell_offset[d] >= 0 &&
            desc->cell_offset[d] + desc->n_cells[d] <= desc->n_cells_global[d] && desc->right[d] > desc->left[d]))
        {
          delete m;
          return hd::fail(HD_ERR_INVALID, "inconsistent cell counts / domain in hd_mesh_desc");
        }
      for (int s = 0; s < 2; ++s)
        if (desc->side_kind[d][s] < 0 || desc->side_kind[d][s] > HD_SIDE_DIRICHLET_HOM)
          {
            delete m;
            return hd::fail(HD_ERR_INVALID, "bad side_kind");
          }
      m->nd *= m->n;
      m->ncells *= desc->n_cells[d];
      // same expression as the oracle / the reference's subdivided_hyper_rectangle
      m->h[d] = (desc->right[d] - desc->left[d]) / desc->n_cells_global[d];
    }
  m->nf    = m->nd / m->n;
  m->ndofs = m->nd * m->ncells;
  int64_t off = 0;
  for (int d = 0; d < dim; ++d)
    for (int s = 0; s < 2; ++s)
      {
        m->ghost_off[d][s] = off;
        if (desc->side_kind[d][s] == HD_SIDE_GHOST)
          {
            m->ghost_cnt[d][s] = (m->ncells / desc->n_cells[d]) * m->nf;
            off += m->ghost_cnt[d][s];
            m->has_ghosts = true;
          }
        if (desc->side_kind[d][s] == HD_SIDE_DIRICHLET || desc->side_kind[d][s] == HD_SIDE_DIRICHLET_HOM)
          m->has_dirichlet = true;
      }
  m->ghost_total = off;
  try
    {
      m->basis.init(desc->degree, desc->n_points, desc->collocation != 0);
    }
  catch (const std::exception &e)
    {
      delete m;
      return hd::fail(HD_ERR_INVALID, e.what());
    }
  // device basis block: nodes[n], xq[nq], w[nq], S[nq*n], Sinv[n*nq]
  std::vector<double> hb;
  for (auto v : m->basis.nodes)
    hb.push_back((double)v);
  for (auto v : m->basis.xq)
    hb.push_back((double)v);
  for (auto v : m->basis.w)
    hb.push_back((double)v);
  for (auto v : m->basis.S)
    hb.push_back((double)v);
  for (auto v : m->basis.Sinv)
    hb.push_back((double)v);
  // (device allocations: a failure past this point frees what was allocated — mesh_device_setup returns, the caller destroys)
  const int rc = mesh_device_setup(m, hb);
  if (rc != HD_OK)
    {
      hd_mesh_destroy(m);
      return rc;
    }
  *out = m;
  return HD_OK;
}

static int
mesh_device_setup(hd_mesh *m, const std::vector<double> &hb)
{
  const hd_mesh_desc *desc = &m->d;
  HD_CUDA(cudaSetDevice(m->ctx->device));
  HD_CUDA(cudaMalloc(&m->d_basis, hb.size() * sizeof(double)));
  HD_CUDA(cudaMemcpy(m->d_basis, hb.data(), hb.size() * sizeof(double), cudaMemcpyHostToDevice));
  HD_CUDA(cudaMalloc(&m->d_reduce, 2 * sizeof(double)));
  {
    // Gauss-Lobatto JxW at the nodes of one v-cell, lowest v-direction fastest (velocity_space_integration)
    std::vector<double> wv(1, 1.0);
    for (int d = desc->dim_x; d < m->dim; ++d)
      {
        std::vector<double> nxt;
        for (int i = 0; i < m->n; ++i)
          for (double x : wv)
            nxt.push_back(x * (double)m->basis.w_nodes[i] * m->h[d]);
        wv.swap(nxt);
      }
    HD_CUDA(cudaMalloc(&m->d_wv, wv.size() * sizeof(double)));
    HD_CUDA(cudaMemcpy(m->d_wv, wv.data(), wv.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  return HD_OK;
}

int
hd_mesh_destroy(hd_mesh *m)
{
  if (!m)
    return HD_OK;
  cudaFree(m->d_basis);
  cudaFree(m->d_reduce);
  cudaFree(m->d_wv);
  delete m;
  return HD_OK;
}

int64_t
hd_mesh_n_dofs(const hd_mesh *m)
{
  return m ? m->ndofs : 0;
}
int64_t
hd_mesh_n_cells(const hd_mesh *m)
{
  return m ? m->ncells : 0;
}
int
hd_mesh_dofs_per_cell(const hd_mesh *m)
{
  return m ? (int)m->nd : 0;
}
int64_t
hd_mesh_ghost_size(const hd_mesh *m, int dir, int side)
{
  if (!m || dir < 0 || dir >= m->dim || side < 0 || side > 1)
    return 0;
  return m->ghost_cnt[dir][side];
}
int64_t
hd_halo_offset(const hd_mesh *m, int dir, int side)
{
  if (!m || dir < 0 || dir >= m->dim || side < 0 || side > 1)
    return 0;
  return m->ghost_off[dir][side];
}
int64_t
hd_halo_total(const hd_mesh *m)
{
  return m ? m->ghost_total : 0;
}

int
hd_mesh_basis(const hd_mesh *m, int which, double *out)
{
  HD_REQUIRE(m, "null mesh");
  const std::vector<hd::LD> *v = nullptr;
  switch (which)
    {
      case 0:
        v = &m->basis.nodes;
        break;
      case 1:
        v = &m->basis.xq;
        break;
      case 2:
        v = &m->basis.w;
        break;
      case 3:
        v = &m->basis.S;
        break;
      case 4:
        v = &m->basis.D;
        break;
      case 5:
        v = &m->basis.Sinv;
        break;
      default:
        return hd::fail(HD_ERR_INVALID, "hd_mesh_basis: which must be 0..5");
    }
  if (out)
    for (size_t i = 0; i < v->size(); ++i)
      out[i] = (double)(*v)[i];
  return (int)v->size();
}

// ---- vectors -------------------------------------------------------------------------------
int
hd_vector_alloc(hd_mesh *m, int do_ghosts, void **ptr)
{
  HD_REQUIRE(m && ptr, "null argument");
  HD_CUDA(cudaSetDevice(m->ctx->device));
  const size_t bytes = (size_t)(m->ndofs + (do_ghosts ? m->ghost_total : 0)) * m->elem_size;
  HD_CUDA(cudaMalloc(ptr, bytes ? bytes : 16));
  if (cudaMemsetAsync(*ptr, 0, bytes, m->ctx->stream) != cudaSuccess)
    {
      cudaFree(*ptr);
      *ptr = nullptr;
      return hd::fail(HD_ERR_CUDA, "hd_vector_alloc: cudaMemsetAsync failed");
    }
  {
    std::lock_guard<std::mutex> lock(m->vectors_mutex);
    m->vectors[*ptr] = m->ndofs + (do_ghosts ? m->ghost_total : 0);
  }
  return HD_OK;
}

int
hd_vector_free(hd_mesh *m, void *ptr)
{
  HD_REQUIRE(m, "null mesh");
  {
    std::lock_guard<std::mutex> lock(m->vectors_mutex);
    m->vectors.erase(ptr);
  }
  HD_CUDA(cudaFree(ptr));
  return HD_OK;
}

// element count of an operation on `ptr`: n values requested (n < 0: "the owned range of a phase-space vector");
// a vector this mesh allocated is never accessed beyond its size
static int
checked_count(hd_mesh *m, const void *ptr, int64_t n, int64_t *out, const char *what)
{
  const int64_t have = m->vector_values(ptr);
  if (n < 0)
    n = (have >= 0 && have < m->ndofs) ? have : m->ndofs; // (an x-space vector holds fewer values than the phase-space range)
  if (have >= 0 && n > have)
    return hd::fail(HD_ERR_INVALID, std::string(what) + ": " + std::to_string(n) + " values requested, the vector holds " + std::to_string(have));
  *out = n;
  return HD_OK;
}

int
hd_vector_copy_in(hd_mesh *m, void *ptr, const void *host, int64_t n)
{
  HD_REQUIRE(m && ptr && host && n >= 0, "bad argument");
  int rc = checked_count(m, ptr, n, &n, "hd_vector_copy_in");
  if (rc != HD_OK)
    return rc;
  HD_CUDA(cudaMemcpyAsync(ptr, host, (size_t)n * m->elem_size, cudaMemcpyHostToDevice, m->ctx->stream));
  HD_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return HD_OK;
}

int
hd_vector_copy_out(hd_mesh *m, const void *ptr, void *host, int64_t n)
{
  HD_REQUIRE(m && ptr && host && n >= 0, "bad argument");
  int rc = checked_count(m, ptr, n, &n, "hd_vector_copy_out");
  if (rc != HD_OK)
    return rc;
  HD_CUDA(cudaMemcpyAsync(host, ptr, (size_t)n * m->elem_size, cudaMemcpyDeviceToHost, m->ctx->stream));
  HD_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return HD_OK;
}

int
hd_vector_zero_n(hd_mesh *m, void *ptr, int64_t n)
{
  HD_REQUIRE(m && ptr, "bad argument");
  int rc = checked_count(m, ptr, n, &n, "hd_vector_zero");
  if (rc != HD_OK)
    return rc;
  HD_CUDA(cudaMemsetAsync(ptr, 0, (size_t)n * m->elem_size, m->ctx->stream));
  return HD_OK;
}

int
hd_vector_zero(hd_mesh *m, void *ptr)
{
  return hd_vector_zero_n(m, ptr, -1);
}

int
hd_vector_copy_n(hd_mesh *m, void *dst, const void *src, int64_t n)
{
  HD_REQUIRE(m && dst && src, "bad argument");
  int64_t nd = n, ns = n;
  int     rc = checked_count(m, dst, n, &nd, "hd_vector_copy (dst)");
  if (rc != HD_OK)
    return rc;
  rc = checked_count(m, src, n, &ns, "hd_vector_copy (src)");
  if (rc != HD_OK)
    return rc;
  if (nd != ns)
    return hd::fail(HD_ERR_INVALID, "hd_vector_copy: vectors of different size (" + std::to_string(nd) + " and " + std::to_string(ns) + " values)");
  HD_CUDA(cudaMemcpyAsync(dst, src, (size_t)nd * m->elem_size, cudaMemcpyDeviceToDevice, m->ctx->stream));
  return HD_OK;
}

int
hd_vector_copy(hd_mesh *m, void *dst, const void *src)
{
  return hd_vector_copy_n(m, dst, src, -1);
}

// ---- advection operator ---------------------------------------------------------------------
} // extern "C"
namespace
{
  template <typename T, int N>
  void
  fill_coef(const hd_advection *op, std::vector<unsigned char> &blob)
  {
    const int dim = op->mesh->dim;
    blob.assign(sizeof(DirCoef<T, N>) * dim, 0);
    DirCoef<T, N> *c = reinterpret_cast<DirCoef<T, N> *>(blob.data());
    for (int d = 0; d < dim; ++d)
      {
        for (int v = 0; v < 4; ++v)
          for (int i = 0; i < N * N; ++i)
            c[d].C[v][i] = T(op->hC[d][v][i]);
        for (int i = 0; i < N; ++i)
          {
            c[d].L0[i] = T(op->hL0[d][i]);
            c[d].L1[i] = T(op->hL1[d][i]);
          }
      }
  }
} // namespace
extern "C" {

// collapsed matrices of every direction for op->a, op->skew, op->eval_level: host copies + the device blob of the generic kernel
static int
build_coefficients(hd_advection *op)
{
  hd_mesh *    m    = op->mesh;
  const double skew = op->skew;
  hd::Basis1D &b = m->basis;
  b.set_skew((hd::LD)skew);
  const int           n = m->n;
  std::vector<double> lifts((size_t)m->dim * 2 * n, 0.0);
  for (int d = 0; d < m->dim; ++d)
    {
      std::vector<hd::LD> C[4], L0, L1;
      b.direction_matrices((hd::LD)op->a[d], (hd::LD)m->h[d], (hd::LD)skew, C, L0, L1, op->eval_level);
      for (int v = 0; v < 4; ++v)
        {
          op->hC[d][v].resize(n * n);
          for (int i = 0; i < n * n; ++i)
            op->hC[d][v][i] = (double)C[v][i];
        }
      op->hL0[d].resize(n);
      op->hL1[d].resize(n);
      bool any0 = false, any1 = false;
      for (int i = 0; i < n; ++i)
        {
          op->hL0[d][i] = (double)L0[i];
          op->hL1[d][i] = (double)L1[i];
          any0 |= (L0[i] != 0);
          any1 |= (L1[i] != 0);
          lifts[(size_t)(2 * d + 0) * n + i] = (double)(2 * L0[i]);
          lifts[(size_t)(2 * d + 1) * n + i] = (double)(2 * L1[i]);
        }
      op->nb_mask[d] = (any0 ? 1 : 0) | (any1 ? 2 : 0);
    }
  std::vector<unsigned char> blob;
  const bool                 f64 = m->d.number_type == HD_F64;
  switch (n)
    {
      case 2:
        f64 ? fill_coef<double, 2>(op, blob) : fill_coef<float, 2>(op, blob);
        break;
      case 3:
        f64 ? fill_coef<double, 3>(op, blob) : fill_coef<float, 3>(op, blob);
        break;
      case 4:
        f64 ? fill_coef<double, 4>(op, blob) : fill_coef<float, 4>(op, blob);
        break;
      case 5:
        f64 ? fill_coef<double, 5>(op, blob) : fill_coef<float, 5>(op, blob);
        break;
      case 6:
        f64 ? fill_coef<double, 6>(op, blob) : fill_coef<float, 6>(op, blob);
        break;
      default:
        return hd::fail(HD_ERR_UNSUPPORTED, "degree must be in 1..5");
    }
  // layout: [DirCoef block, padded to 8 bytes][double lifts[dim][2][n]]
  op->coef_bytes = (blob.size() + 7) / 8 * 8;
  blob.resize(op->coef_bytes + lifts.size() * sizeof(double));
  std::memcpy(blob.data() + op->coef_bytes, lifts.data(), lifts.size() * sizeof(double));
  HD_CUDA(cudaSetDevice(m->ctx->device));
  cudaFree(op->d_coef);
  op->d_coef = nullptr;
  HD_CUDA(cudaMalloc(&op->d_coef, blob.size()));
  HD_CUDA(cudaMemcpy(op->d_coef, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  return HD_OK;
}

int
hd_advection_create(hd_mesh *m, double skew, const double *velocity, hd_advection **out)
{
  HD_REQUIRE(m && velocity && out, "null argument");
  hd_advection *op = new (std::nothrow) hd_advection;
  HD_REQUIRE(op, "out of memory");
  op->mesh = m;
  op->skew = skew;
  for (int d = 0; d < HD_MAX_DIM; ++d)
    {
      op->a[d]       = d < m->dim ? velocity[d] : 0.0;
      op->nb_mask[d] = 0;
      for (int s = 0; s < 2; ++s)
        {
          op->d_g[d][s]     = nullptr;
          op->g_count[d][s] = 0;
        }
    }
  const int rc = build_coefficients(op);
  if (rc != HD_OK)
    {
      delete op;
      return rc;
    }
  *out = op;
  return HD_OK;
}

int
hd_advection_set_evaluation_level(hd_advection *op, int level)
{
  HD_REQUIRE(op && level >= HD_EVAL_ALL && level <= HD_EVAL_ALL_WITHOUT_NEIGHBOR_LOAD, "bad argument");
  HD_REQUIRE(!op->mesh->has_dirichlet || level == HD_EVAL_ALL, "partial evaluation levels are profiling variants for periodic / ghosted lattices");
  if (level == op->eval_level)
    return HD_OK;
  HD_CUDA(cudaStreamSynchronize(op->mesh->ctx->stream)); // the coefficient blob may still be in use
  op->eval_level = level;
  return build_coefficients(op);
}

int
hd_advection_destroy(hd_advection *op)
{
  if (!op)
    return HD_OK;
  hd::fast6d_release(op);
  if (op->shadow_op)
    hd_advection_destroy(op->shadow_op);
  if (op->shadow_mesh)
    hd_mesh_destroy(op->shadow_mesh);
  cudaFree(op->d_shadow_ghost);
  cudaFree(op->d_coef);
  cudaFree(op->d_vp_coef);
  cudaFree(op->d_stage_src);
  cudaFree(op->d_stage_dst);
  if (op->s_h2d)
    cudaStreamDestroy(op->s_h2d);
  if (op->s_d2h)
    cudaStreamDestroy(op->s_d2h);
  for (cudaEvent_t e : op->ev_in)
    cudaEventDestroy(e);
  for (cudaEvent_t e : op->ev_done)
    cudaEventDestroy(e);
  for (int d = 0; d < HD_MAX_DIM; ++d)
    for (int s = 0; s < 2; ++s)
      cudaFree(op->d_g[d][s]);
  delete op;
  return HD_OK;
}

int
hd_advection_set_phase_space_velocity(hd_advection *op, const double *a_v_device)
{
  HD_REQUIRE(op, "null argument");
  if (!a_v_device)
    {
      op->d_av = nullptr;
      return HD_OK;
    }
  std::string why;
  if (!hd::vp_supported(op, &why))
    return hd::fail(HD_ERR_UNSUPPORTED, why);
  int rc = hd::vp_upload_coefficients(op);
  if (rc != HD_OK)
    return rc;
  op->d_av = a_v_device;
  return HD_OK;
}

int
hd_advection_set_kernel(hd_advection *op, int which)
{
  HD_REQUIRE(op && which >= 0 && which <= 6, "bad argument");
  if (which == 5 && !hd::tile_global_supported(op))
    return hd::fail(HD_ERR_UNSUPPORTED, "the global-memory tile kernel covers degree 3 and 5 with an even number of directions, without Dirichlet sides");
  if (which == 4 && !hd::tile_row_supported(op))
    return hd::fail(HD_ERR_UNSUPPORTED, "the row-persistent tile kernel covers degree 3 in 3D3V without Dirichlet sides");
  if ((which == 2 || which == 6) && !hd::fast6d_supported(op))
    return hd::fail(HD_ERR_UNSUPPORTED, "the fused 3D3V k=3 kernel does not cover this configuration");
  if (which == 3 && !hd::tile_supported(op))
    return hd::fail(HD_ERR_UNSUPPORTED, "the tile kernel covers degree 3 in 1D1V / 2D2V / 3D3V without Dirichlet sides");
  op->kernel_choice = which;
  return HD_OK;
}

int
hd_advection_set_l2_hints(hd_advection *op, int mask)
{
  HD_REQUIRE(op && mask >= -1 && mask <= 7, "bad argument");
  op->l2_hints = mask;
  return HD_OK;
}

int
hd_advection_set_row_tile(hd_advection *op, const int *tile)
{
  HD_REQUIRE(op && tile, "null argument");
  for (int i = 0; i < 5; ++i)
    {
      HD_REQUIRE(tile[i] >= -1, "bad tile extent");
      op->row_tile[i] = tile[i];
    }
  return HD_OK;
}

const char *
hd_advection_kernel_name(const hd_advection *op)
{
  return op ? op->last_kernel : "none";
}

int64_t
hd_advection_launch_count(const hd_advection *op)
{
  return op ? op->launches : 0;
}

static int apply_impl(hd_advection *op, void *dst, const void *src, const void *ghosts, double time, const FusedUpdate &fu, int part);

// Dirichlet sides on the specialised kernels.  At a Dirichlet face the upwind flux sees u+ = -u- + 2 g
// (advection_operation.h:490-519); in the collapsed form (basis.hpp) that is the INTERIOR formula with the neighbour trace
// -u_face + 2 ghat.  So a lattice with Dirichlet sides is served by any kernel that knows ghost sides: a small kernel writes
// those traces of the inflow sides into an internal ghost buffer (k_dirichlet_source in ghost mode, one CTA per boundary
// face cell), then the operator runs on a shadow description of the same lattice whose inflow Dirichlet sides are
// HD_SIDE_GHOST (outflow sides read nothing: beta = 0).  Single-brick lattices only (a brick with real ghost sides keeps
// the generic kernel + lifting kernel); hd_advection_set_kernel(op, 1) forces that path too.
static bool
dirichlet_shadow(hd_advection *op)
{
  if (op->shadow_state != 0)
    return op->shadow_state > 0;
  op->shadow_state = -1;
  hd_mesh *m = op->mesh;
  if (m->has_ghosts)
    return false;
  hd_mesh_desc d   = m->d;
  bool         any = false;
  for (int k = 0; k < m->dim; ++k)
    for (int s = 0; s < 2; ++s)
      if (d.side_kind[k][s] == HD_SIDE_DIRICHLET || d.side_kind[k][s] == HD_SIDE_DIRICHLET_HOM)
        {
          d.side_kind[k][s] = ((op->nb_mask[k] >> s) & 1) ? HD_SIDE_GHOST : HD_SIDE_PERIODIC_LOCAL;
          any               = true;
        }
  if (!any)
    return false;
  if (hd_mesh_create(m->ctx, &d, &op->shadow_mesh) != HD_OK)
    return false;
  if (hd_advection_create(op->shadow_mesh, op->skew, op->a, &op->shadow_op) != HD_OK)
    return false;
  hd_advection *sh = op->shadow_op;
  // worth it only if the shadow lattice gets one of the specialised kernels
  const bool special = hd::fast6d_supported(sh) || hd::tile_preferred(sh) || (op->shadow_mesh->n == 6 && hd::tile_global_supported(sh));
  if (!special)
    return false;
  const size_t bytes = (size_t)hd_halo_total(op->shadow_mesh) * m->elem_size;
  if (cudaMalloc(&op->d_shadow_ghost, bytes ? bytes : 256) != cudaSuccess)
    {
      cudaGetLastError();
      return false;
    }
  op->shadow_state = 1;
  return true;
}

static int
apply_dirichlet_as_ghosts(hd_advection *op, void *dst, const void *src, double time, const FusedUpdate &fu)
{
  int rc = hd::launch_dirichlet_ghosts(op, op->shadow_mesh, op->d_shadow_ghost, src, time);
  if (rc != HD_OK)
    return rc;
  hd_advection *sh  = op->shadow_op;
  sh->kernel_choice = op->kernel_choice;
  const int64_t before = sh->launches;
  rc = apply_impl(sh, dst, src, op->d_shadow_ghost, time, fu, HD_PART_ALL);
  op->launches += sh->launches - before;
  op->last_kernel = sh->last_kernel;
  return rc;
}

static int
apply_impl(hd_advection *op, void *dst, const void *src, const void *ghosts, double time, const FusedUpdate &fu, int part = HD_PART_ALL)
{
  HD_REQUIRE(part == HD_PART_ALL || part == HD_PART_INTERIOR || part == HD_PART_BOUNDARY, "bad part");
  HD_REQUIRE(op && src && (dst || fu.enabled), "null argument");
  HD_REQUIRE(dst != src, "dst and src must not alias (ECL reads neighbours of src)");
  hd_mesh *m = op->mesh;
  HD_REQUIRE(!m->has_ghosts || ghosts, "mesh has HD_SIDE_GHOST sides but no ghost buffer was passed");
  HD_CUDA(cudaSetDevice(m->ctx->device));
  int  rc;
  if (op->d_av)
    {
      // phase-space velocity field: the general-velocity kernel, no interior/boundary split
      if (part == HD_PART_INTERIOR)
        return HD_OK;
      return hd::launch_vp(op, dst, src, time, fu);
    }
  if (m->has_dirichlet && part == HD_PART_ALL && op->kernel_choice != 1 && dirichlet_shadow(op))
    return apply_dirichlet_as_ghosts(op, dst, src, time, fu);
  bool fast = op->kernel_choice == 2 || op->kernel_choice == 6 || (op->kernel_choice == 0 && hd::fast6d_supported(op));
  if (fast)
    rc = hd::launch_fast6d(op, dst, src, ghosts, time, fu, part);
  else
    {
      // the generic and tile kernels have no interior/boundary split: everything runs in the boundary part
      if (part == HD_PART_INTERIOR)
        return HD_OK;
      const bool tile = op->kernel_choice == 3 || (op->kernel_choice == 0 && hd::tile_preferred(op));
      // degree 5 (n = 6; BASELINE.json configs[2] is 3D3V degree 5 in FP32): the global-memory tile kernel beats the
      // generic one (95 against 79 GDoF/s on 6^3 x 4^3 cells, profiles/r02a_zoo_tg.txt) wherever it applies
      const bool tg = op->kernel_choice == 5 || (op->kernel_choice == 0 && m->n == 6 && hd::tile_global_supported(op));
      if (op->kernel_choice == 4)
        rc = hd::launch_tile_row(op, dst, src, ghosts, time, fu);
      else if (tg)
        rc = hd::launch_tile_global(op, dst, src, ghosts, time, fu);
      else
        rc = tile ? hd::launch_tile(op, dst, src, ghosts, time, fu) : hd::launch_generic(op, dst, src, ghosts, time, fu);
    }
  if (rc != HD_OK)
    return rc;
  if (part == HD_PART_INTERIOR)
    return HD_OK;
  bool any_dirichlet = false;
  for (int d = 0; d < m->dim; ++d)
    for (int s = 0; s < 2; ++s)
      any_dirichlet |= (m->d.side_kind[d][s] == HD_SIDE_DIRICHLET);
  if (any_dirichlet)
    return hd::launch_dirichlet_source(op, dst, time, fu);
  return HD_OK;
}

int
hd_advection_apply(hd_advection *op, void *dst, const void *src, const void *ghosts, double time)
{
  FusedUpdate fu;
  return apply_impl(op, dst, src, ghosts, time, fu);
}

int
hd_advection_apply_part(hd_advection *op, void *dst, const void *src, const void *ghosts, double time, int part)
{
  FusedUpdate fu;
  return apply_impl(op, dst, src, ghosts, time, fu, part);
}

int
hd_advection_apply_overlapped(hd_advection *op, void *dst, const void *src, const void *ghosts, double time, const hd_halo_send *sends, int n_sends,
                              const void *arrival_counters, int target)
{
  HD_REQUIRE(op && dst && src && ghosts && arrival_counters && target > 0, "null argument");
  HD_REQUIRE(dst != src, "dst and src must not alias (ECL reads neighbours of src)");
  hd_mesh *m = op->mesh;
  HD_CUDA(cudaSetDevice(m->ctx->device));
  const bool fast = op->kernel_choice == 2 || op->kernel_choice == 6 || (op->kernel_choice == 0 && hd::fast6d_supported(op));
  if (!fast)
    return hd::fail(HD_ERR_UNSUPPORTED, "hd_advection_apply_overlapped needs the pipelined 3D3V kernel; use hd_advection_apply_part");
  for (int d = 0; d < m->dim; ++d)
    for (int s = 0; s < 2; ++s)
      if (m->d.side_kind[d][s] == HD_SIDE_DIRICHLET)
        return hd::fail(HD_ERR_UNSUPPORTED, "hd_advection_apply_overlapped: Dirichlet sides are not supported");
  FusedUpdate fu;
  return hd::launch_fast6d(op, dst, src, ghosts, time, fu, 3, sends, n_sends, arrival_counters, target);
}

// ---- peer-mapped memory across processes (CUDA IPC) ------------------------------------------------------------
int
hd_ipc_export(hd_context *ctx, const void *ptr, void *handle)
{
  HD_REQUIRE(ctx && ptr && handle, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == HD_IPC_HANDLE_BYTES, "IPC handle size");
  HD_CUDA(cudaSetDevice(ctx->device));
  HD_CUDA(cudaStreamSynchronize(ctx->stream)); // (hd_device_malloc zeroes on the stream)
  cudaIpcMemHandle_t h;
  HD_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(ptr)));
  std::memcpy(handle, &h, sizeof(h));
  return HD_OK;
}

int
hd_ipc_open(hd_context *ctx, const void *handle, void **ptr)
{
  HD_REQUIRE(ctx && handle && ptr, "null argument");
  HD_CUDA(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  HD_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return HD_OK;
}

int
hd_ipc_close(hd_context *ctx, void *ptr)
{
  HD_REQUIRE(ctx, "null context");
  HD_CUDA(cudaSetDevice(ctx->device));
  HD_CUDA(cudaIpcCloseMemHandle(ptr));
  return HD_OK;
}

int
hd_advection_n_ctas(const hd_advection *op)
{
  if (!op)
    return 0;
  const hd_mesh *m = op->mesh;
  if (!hd::fast6d_supported(op))
    return 0;
  long long nrows = m->ncells / m->d.n_cells[0];
  return (int)(nrows < m->ctx->sm_count ? nrows : m->ctx->sm_count);
}

int
hd_advection_n_halo_senders(const hd_advection *op)
{
  if (!op || !hd::fast6d_supported(op))
    return 0;
  return hd::fast6d_halo_senders(op);
}

int
hd_advection_set_halo_senders(hd_advection *op, int n)
{
  HD_REQUIRE(op && n >= 0, "bad argument");
  op->halo_senders = n;
  return HD_OK;
}

int
hd_advection_overlap_status(hd_advection *op, int *timed_out)
{
  HD_REQUIRE(op && timed_out, "null argument");
  HD_CUDA(cudaSetDevice(op->mesh->ctx->device));
  return hd::fast6d_overlap_status(op, timed_out);
}

namespace
{
  // cuStreamWriteValue32 / cuStreamWaitValue32: stream memory operations — no kernel, so they can neither be blocked by
  // nor block a persistent kernel that fills the SMs
  typedef int (*StreamValueFn)(void *, unsigned long long, unsigned int, unsigned int);
  int
  stream_value_fn(const char *name, StreamValueFn *out)
  {
    void *                          f = nullptr;
    cudaDriverEntryPointQueryResult q;
    HD_CUDA(cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &q));
    if (!f || q != cudaDriverEntryPointSuccess)
      return hd::fail(HD_ERR_CUDA, std::string(name) + " is not available in this driver");
    *out = reinterpret_cast<StreamValueFn>(f);
    return HD_OK;
  }
} // namespace

int
hd_stream_write_flag(hd_context *ctx, void *flag_device, int value)
{
  HD_REQUIRE(ctx && flag_device, "null argument");
  HD_CUDA(cudaSetDevice(ctx->device));
  static StreamValueFn fn = nullptr;
  if (!fn)
    {
      const int rc = stream_value_fn("cuStreamWriteValue32", &fn);
      if (rc != HD_OK)
        return rc;
    }
  const int r = fn(ctx->stream, (unsigned long long)(uintptr_t)flag_device, (unsigned int)value, 0u);
  if (r != 0)
    return hd::fail(HD_ERR_CUDA, "cuStreamWriteValue32 failed with code " + std::to_string(r));
  return HD_OK;
}

int
hd_stream_wait_flag(hd_context *ctx, void *flag_device, int value)
{
  HD_REQUIRE(ctx && flag_device, "null argument");
  HD_CUDA(cudaSetDevice(ctx->device));
  static StreamValueFn fn = nullptr;
  if (!fn)
    {
      const int rc = stream_value_fn("cuStreamWaitValue32", &fn);
      if (rc != HD_OK)
        return rc;
    }
  const int r = fn(ctx->stream, (unsigned long long)(uintptr_t)flag_device, (unsigned int)value, 0x0u /* CU_STREAM_WAIT_VALUE_GEQ */);
  if (r != 0)
    return hd::fail(HD_ERR_CUDA, "cuStreamWaitValue32 failed with code " + std::to_string(r));
  return HD_OK;
}

int
hd_advection_apply_host(hd_advection *op, void *dst_host, const void *src_host, double time)
{
  HD_REQUIRE(op && dst_host && src_host, "null argument");
  hd_mesh *m = op->mesh;
  HD_REQUIRE(!m->has_ghosts, "hd_advection_apply_host is for single-GPU meshes");
  HD_CUDA(cudaSetDevice(m->ctx->device));
  const size_t bytes = (size_t)m->ndofs * m->elem_size;
  if (!op->d_stage_src)
    {
      HD_CUDA(cudaMalloc(&op->d_stage_src, bytes));
      HD_CUDA(cudaMalloc(&op->d_stage_dst, bytes));
    }
  FusedUpdate fu;
  // Pipelined variant (pipelined 3D3V kernel): the lattice is cut into its layers along the slowest direction; layer j
  // is computed as soon as it and its upwind neighbour layer are on the device, and travels back while the next
  // layers are still coming in — copy-in, kernel and copy-out run on three streams, PCIe in both directions at once.
  const int  last = m->dim - 1, n5 = m->d.n_cells[last];
  const bool fast = op->kernel_choice == 2 || op->kernel_choice == 6 || (op->kernel_choice == 0 && hd::fast6d_supported(op));
  if (fast && n5 >= 3 && getenv("HD_HOST_SERIAL") == nullptr)
    {
      // units of the pipeline: layers of the slowest direction, cut again along the second slowest one if it is long
      // enough (shorter start-up and drain: a unit waits for itself and its two upwind neighbour units only)
      const int n4 = m->d.n_cells[last - 1] >= 3 ? m->d.n_cells[last - 1] : 1;
      const int nu = n4 * n5;
      if (!op->s_h2d)
        {
          HD_CUDA(cudaStreamCreateWithFlags(&op->s_h2d, cudaStreamNonBlocking));
          HD_CUDA(cudaStreamCreateWithFlags(&op->s_d2h, cudaStreamNonBlocking));
        }
      while ((int)op->ev_in.size() < nu)
        {
          cudaEvent_t a, b;
          HD_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
          HD_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
          op->ev_in.push_back(a);
          op->ev_done.push_back(b);
        }
      const size_t    unit = bytes / nu;
      const long long rows = (m->ncells / m->d.n_cells[0]) / nu;
      auto            upwind = [&](int d) { return (op->nb_mask[d] & 1) ? -1 : ((op->nb_mask[d] & 2) ? +1 : 0); };
      const int       d5 = upwind(last), d4 = n4 > 1 ? upwind(last - 1) : 0;
      const char *    hs = static_cast<const char *>(src_host);
      char *          hd_ = static_cast<char *>(dst_host);
      char *          ds = static_cast<char *>(op->d_stage_src), *dd = static_cast<char *>(op->d_stage_dst);
      // copy-in order: along each of the two directions the upwind neighbour of index 0 (the periodic wrap) goes first
      for (int jj = 0; jj < n5; ++jj)
        for (int ii = 0; ii < n4; ++ii)
          {
            const int    j = d5 < 0 ? (jj + n5 - 1) % n5 : jj, i = d4 < 0 ? (ii + n4 - 1) % n4 : ii;
            const size_t u = (size_t)j * n4 + i;
            HD_CUDA(cudaMemcpyAsync(ds + u * unit, hs + u * unit, unit, cudaMemcpyHostToDevice, op->s_h2d));
            HD_CUDA(cudaEventRecord(op->ev_in[u], op->s_h2d));
          }
      for (int j = 0; j < n5; ++j)
        for (int i = 0; i < n4; ++i)
          {
            const size_t u = (size_t)j * n4 + i;
            HD_CUDA(cudaStreamWaitEvent(m->ctx->stream, op->ev_in[u], 0));
            if (d4 != 0)
              HD_CUDA(cudaStreamWaitEvent(m->ctx->stream, op->ev_in[(size_t)j * n4 + (i + d4 + n4) % n4], 0));
            if (d5 != 0)
              HD_CUDA(cudaStreamWaitEvent(m->ctx->stream, op->ev_in[(size_t)((j + d5 + n5) % n5) * n4 + i], 0));
            int rc = hd::launch_fast6d(op, dd, ds, nullptr, time, fu, HD_PART_ALL, nullptr, 0, nullptr, 0, (long long)u * rows, (long long)(u + 1) * rows);
            if (rc != HD_OK)
              return rc;
            HD_CUDA(cudaEventRecord(op->ev_done[u], m->ctx->stream));
            HD_CUDA(cudaStreamWaitEvent(op->s_d2h, op->ev_done[u], 0));
            HD_CUDA(cudaMemcpyAsync(hd_ + u * unit, dd + u * unit, unit, cudaMemcpyDeviceToHost, op->s_d2h));
          }
      HD_CUDA(cudaStreamSynchronize(op->s_d2h));
      HD_CUDA(cudaStreamSynchronize(m->ctx->stream));
      return HD_OK;
    }
  HD_CUDA(cudaMemcpyAsync(op->d_stage_src, src_host, bytes, cudaMemcpyHostToDevice, m->ctx->stream));
  int         rc = apply_impl(op, op->d_stage_dst, op->d_stage_src, nullptr, time, fu);
  if (rc != HD_OK)
    return rc;
  HD_CUDA(cudaMemcpyAsync(dst_host, op->d_stage_dst, bytes, cudaMemcpyDeviceToHost, m->ctx->stream));
  HD_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return HD_OK;
}

int
hd_advection_set_dirichlet_values(hd_advection *op, int dir, int side, const double *g, int64_t n)
{
  HD_REQUIRE(op && g, "null argument");
  hd_mesh *m = op->mesh;
  HD_REQUIRE(dir >= 0 && dir < m->dim && side >= 0 && side <= 1, "bad (dir, side)");
  HD_REQUIRE(m->d.side_kind[dir][side] == HD_SIDE_DIRICHLET, "side is not an inhomogeneous Dirichlet boundary");
  int64_t nqf = 1;
  for (int e = 0; e < m->dim - 1; ++e)
    nqf *= m->nq;
  const int64_t expect = (m->ncells / m->d.n_cells[dir]) * nqf;
  HD_REQUIRE(n == expect, "wrong number of boundary values");
  HD_CUDA(cudaSetDevice(m->ctx->device));
  if (!op->d_g[dir][side])
    HD_CUDA(cudaMalloc(&op->d_g[dir][side], (size_t)n * sizeof(double)));
  op->g_count[dir][side] = n;
  HD_CUDA(cudaMemcpyAsync(op->d_g[dir][side], g, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, m->ctx->stream));
  HD_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return HD_OK;
}

int
hd_advection_set_dirichlet_builtin(hd_advection *op, int fn_id)
{
  HD_REQUIRE(op && (fn_id == HD_FN_ZERO || fn_id == HD_FN_HYPERRECTANGLE), "bad argument");
  op->dirichlet_fn = fn_id;
  return HD_OK;
}

// ---- halo -----------------------------------------------------------------------------------
int
hd_halo_pack_ex(hd_mesh *m, const void *src, void *send, const int *send_mask, void *const *peer_dst, void *started_counter, int *ctas_launched)
{
  HD_REQUIRE(m && src, "null argument");
  if (ctas_launched)
    *ctas_launched = 0;
  if (!m->has_ghosts)
    return HD_OK;
  HD_CUDA(cudaSetDevice(m->ctx->device));
  const LatticeParams lp = lattice(m);
  for (int d = 0; d < m->dim; ++d)
    for (int s = 0; s < 2; ++s)
      {
        const long long cnt = m->ghost_cnt[d][s];
        if (cnt == 0 || (send_mask && !send_mask[2 * d + s]))
          continue;
        void *out = peer_dst ? peer_dst[2 * d + s] : nullptr;
        if (!out)
          {
            HD_REQUIRE(send, "null send buffer");
            out = static_cast<char *>(send) + (size_t)m->ghost_off[d][s] * m->elem_size;
          }
        // 16-byte chunks if the layer is contiguous over that much and everything is aligned
        long long stride_d = 1;
        for (int e = 0; e < d; ++e)
          stride_d *= m->n;
        const int  vec   = int(16 / m->elem_size);
        const bool wide  = (stride_d % vec == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0) && (m->nd % vec == 0) && (m->nf % vec == 0);
        // 128-thread CTAs, two per SM: small enough to run beside the persistent operator kernel, so that packing
        // (and the NVLink stores of the direct variant) overlaps with the interior cells
        const long long nfc = cnt / m->nf;
        HD_REQUIRE(nfc < (1ll << 31) && m->nf < (1ll << 30), "face too large for the pack kernel");
        unsigned g = (unsigned)(nfc < 2ll * m->ctx->sm_count ? nfc : 2ll * m->ctx->sm_count);
        const int nfi = (int)m->nf, sdi = (int)stride_d, nfci = (int)nfc;
        // An SM only hosts kernels of one shared-memory carve-out at a time: ask for the operator kernel's (maximum shared
        // memory), otherwise pack CTAs and the persistent operator CTAs exclude each other and nothing overlaps.
        bool &carveout_set = m->ctx->pack_carveout_set; // (a per-device attribute: per context, not a process-wide static)
        if (!carveout_set)
          {
            cudaFuncSetAttribute(k_halo_pack<double, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_halo_pack<double, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_halo_pack<float, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(k_halo_pack<float, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            carveout_set = true;
          }
        if (m->d.number_type == HD_F64)
          {
            if (wide)
              k_halo_pack<double, 2><<<g, 128, 0, m->ctx->stream>>>(static_cast<const double *>(src), static_cast<double *>(out), lp, d, s, nfci, nfi, sdi, static_cast<int *>(started_counter));
            else
              k_halo_pack<double, 1><<<g, 128, 0, m->ctx->stream>>>(static_cast<const double *>(src), static_cast<double *>(out), lp, d, s, nfci, nfi, sdi, static_cast<int *>(started_counter));
          }
        else
          {
            if (wide)
              k_halo_pack<float, 4><<<g, 128, 0, m->ctx->stream>>>(static_cast<const float *>(src), static_cast<float *>(out), lp, d, s, nfci, nfi, sdi, static_cast<int *>(started_counter));
            else
              k_halo_pack<float, 1><<<g, 128, 0, m->ctx->stream>>>(static_cast<const float *>(src), static_cast<float *>(out), lp, d, s, nfci, nfi, sdi, static_cast<int *>(started_counter));
          }
        HD_CUDA(cudaGetLastError());
        if (ctas_launched)
          *ctas_launched += (int)g;
      }
  return HD_OK;
}

int
hd_halo_pack(hd_mesh *m, const void *src, void *send)
{
  return hd_halo_pack_ex(m, src, send, nullptr, nullptr, nullptr, nullptr);
}

int
hd_advection_ghost_sides(const hd_advection *op, int *needed)
{
  HD_REQUIRE(op && needed, "null argument");
  const hd_mesh *m = op->mesh;
  for (int d = 0; d < HD_MAX_DIM; ++d)
    for (int s = 0; s < 2; ++s)
      needed[2 * d + s] = (d < m->dim && m->d.side_kind[d][s] == HD_SIDE_GHOST && ((op->nb_mask[d] >> s) & 1)) ? 1 : 0;
  return HD_OK;
}

// ---- LSRK -----------------------------------------------------------------------------------
int
hd_lsrk_create(hd_mesh *m, const char *type, hd_lsrk **out)
{
  HD_REQUIRE(m && type && out, "null argument");
  hd_lsrk *rk = new (std::nothrow) hd_lsrk;
  HD_REQUIRE(rk, "out of memory");
  rk->mesh             = m;
  const std::string t  = type;
  auto &            bi = rk->bi;
  auto &            ai = rk->ai;
  // Kennedy, Carpenter, Lewis (2000) low-storage schemes; base/time_integrators.templates.h:34-86
  if (t == "rk33")
    {
      bi = {0.245170287303492, 0.184896052186740, 0.569933660509768};
      ai = {0.755726351946097, 0.386954477304099};
    }
  else if (t == "rk45")
    {
      bi = {1153189308089. / 22510343858157., 1772645290293. / 4653164025191., -1672844663538. / 4480602732383., 2114624349019. / 3568978502595., 5198255086312. / 14908931495163.};
      ai = {970286171893. / 4311952581923., 6584761158862. / 12103376702013., 2251764453980. / 15575788980749., 26877169314380. / 34165994151039.};
    }
  else if (t == "rk47")
    {
      bi = {0.0941840925477795334, 0.149683694803496998, 0.285204742060440058, -0.122201846148053668, 0.0605151571191401122, 0.345986987898399296, 0.186627171718797670};
      ai = {0.241566650129646868 + bi[0], 0.0423866513027719953 + bi[1], 0.215602732678803776 + bi[2], 0.232328007537583987 + bi[3], 0.256223412574146438 + bi[4], 0.0978694102142697230 + bi[5]};
    }
  else if (t == "rk59")
    {
      bi = {2274579626619. / 23610510767302., 693987741272. / 12394497460941., -347131529483. / 15096185902911., 1144057200723. / 32081666971178., 1562491064753. / 11797114684756., 13113619727965. / 44346030145118., 393957816125. / 7825732611452., 720647959663. / 6565743875477., 3559252274877. / 14424734981077.};
      ai = {1107026461565. / 5417078080134., 38141181049399. / 41724347789894., 493273079041. / 11940823631197., 1851571280403. / 6147804934346., 11782306865191. / 62590030070788., 9452544825720. / 13648368537481., 4435885630781. / 26285702406235., 2357909744247. / 11371140753790.};
    }
  else
    {
      delete rk;
      return hd::fail(HD_ERR_UNSUPPORTED, "LSRK type must be rk33, rk45, rk47 or rk59");
    }
  *out = rk;
  return HD_OK;
}

int
hd_lsrk_destroy(hd_lsrk *rk)
{
  if (!rk)
    return HD_OK;
  cudaFree(rk->d_ti2);
  delete rk;
  return HD_OK;
}

int
hd_lsrk_n_stages(const hd_lsrk *rk)
{
  return rk ? (int)rk->bi.size() : 0;
}

int
hd_lsrk_coefficients(const hd_lsrk *rk, int which, double *out)
{
  HD_REQUIRE(rk && out && (which == 0 || which == 1), "bad argument");
  const auto &v = which == 0 ? rk->bi : rk->ai;
  for (size_t i = 0; i < v.size(); ++i)
    out[i] = v[i];
  return (int)v.size();
}

int
hd_lsrk_stage_update(hd_mesh *m, void *sol, void *ti, const void *K, double b, double a)
{
  HD_REQUIRE(m && sol && K && (ti || a == 0.0), "null argument");
  HD_CUDA(cudaSetDevice(m->ctx->device));
  const unsigned g = grid_for(m->ctx, m->ndofs, 256);
  if (m->d.number_type == HD_F64)
    k_stage_update<double><<<g, 256, 0, m->ctx->stream>>>(static_cast<double *>(sol), static_cast<double *>(ti), static_cast<const double *>(K), b, a, m->ndofs);
  else
    k_stage_update<float><<<g, 256, 0, m->ctx->stream>>>(static_cast<float *>(sol), static_cast<float *>(ti), static_cast<const float *>(K), (float)b, (float)a, m->ndofs);
  HD_CUDA(cudaGetLastError());
  return HD_OK;
}

int
hd_lsrk_step(hd_lsrk *rk, hd_advection *op, void *solution, void *vec_Ki, void *vec_Ti, double t, double dt)
{
  HD_REQUIRE(rk && op && solution && vec_Ki && vec_Ti, "null argument");
  hd_mesh *m = rk->mesh;
  HD_REQUIRE(m == op->mesh, "integrator and operator belong to different meshes");
  HD_REQUIRE(!m->has_ghosts, "hd_lsrk_step is for single-GPU meshes; drive stages with hd_lsrk_stage_update otherwise");
  HD_REQUIRE(!op->d_av, "hd_lsrk_step with a phase-space velocity field: the field changes at every stage; drive the stages with hd_lsrk_stage_update");
  HD_CUDA(cudaSetDevice(m->ctx->device));
  const size_t bytes = (size_t)m->ndofs * m->elem_size;
  // Fused path: the operator's epilogue does the stage update, so K is never stored and each
  // stage streams Ti (read), solution (read+write) and the next Ti (write) once.  The next Ti
  // must not overwrite the Ti the neighbours still read: ping-pong between vec_Ti and vec_Ki
  // (vec_Ki is free because K never materialises).
  // only_Ti_is_ghosted branch (time_integrators.templates.h:142-156): Ti <- solution
  HD_CUDA(cudaMemcpyAsync(vec_Ti, solution, bytes, cudaMemcpyDeviceToDevice, m->ctx->stream));
  void * cur = vec_Ti, *nxt = vec_Ki;
  double sum_prev_b = 0.0;
  const int S = (int)rk->bi.size();
  for (int stage = 0; stage < S; ++stage)
    {
      double c = 0.0;
      if (stage > 0)
        {
          c = sum_prev_b + rk->ai[stage - 1];
          sum_prev_b += rk->bi[stage - 1];
        }
      FusedUpdate fu;
      fu.enabled = 1;
      fu.sol     = solution;
      fu.ti_next = nxt;
      fu.fb      = rk->bi[stage] * dt;
      fu.fa      = (stage == S - 1) ? 0.0 : rk->ai[stage] * dt;
      int rc     = apply_impl(op, nullptr, cur, nullptr, t + c * dt, fu);
      if (rc != HD_OK)
        return rc;
      void *tmp = cur;
      cur       = nxt;
      nxt       = tmp;
    }
  return HD_OK;
}

// stage time t + c_i dt and the update factors of stage i (time_integrators.templates.h:117-132, :174-182)
static void
lsrk_stage_factors(const hd_lsrk *rk, int stage, double dt, double *c, double *fb, double *fa)
{
  const int S = (int)rk->bi.size();
  double    sum_prev_b = 0.0;
  *c                   = 0.0;
  for (int s = 1; s <= stage; ++s)
    {
      *c = sum_prev_b + rk->ai[s - 1];
      sum_prev_b += rk->bi[s - 1];
    }
  *fb = rk->bi[stage] * dt;
  *fa = (stage == S - 1) ? 0.0 : rk->ai[stage] * dt;
}

int
hd_lsrk_stage_fused(hd_lsrk *rk, hd_advection *op, int stage, void *solution, const void *ti_cur, void *ti_next, const void *ghosts, double t, double dt)
{
  HD_REQUIRE(rk && op && solution && ti_cur && ti_next, "null argument");
  HD_REQUIRE(stage >= 0 && stage < (int)rk->bi.size(), "bad stage");
  HD_REQUIRE(ti_cur != ti_next && ti_cur != solution && ti_next != solution, "solution, ti_cur and ti_next must be three different vectors");
  hd_mesh *m = rk->mesh;
  HD_REQUIRE(m == op->mesh, "integrator and operator belong to different meshes");
  // (a phase-space velocity field is allowed here: the caller refreshes it from ti_cur before every stage — the
  // Vlasov-Poisson right-hand side, examples/vlasov_poisson/include/application.h:516-600)
  HD_CUDA(cudaSetDevice(m->ctx->device));
  double      c;
  FusedUpdate fu;
  fu.enabled = 1;
  fu.sol     = solution;
  fu.ti_next = ti_next;
  lsrk_stage_factors(rk, stage, dt, &c, &fu.fb, &fu.fa);
  return apply_impl(op, nullptr, ti_cur, ghosts, t + c * dt, fu);
}

int
hd_lsrk_stage_overlapped(hd_lsrk *rk, hd_advection *op, int stage, void *solution, const void *ti_cur, void *ti_next, const void *ghosts,
                         const hd_halo_send *sends, int n_sends, const void *arrival_counters, int target, double t, double dt)
{
  HD_REQUIRE(rk && op && solution && ti_cur && ti_next && ghosts && arrival_counters && target > 0, "null argument");
  HD_REQUIRE(stage >= 0 && stage < (int)rk->bi.size(), "bad stage");
  HD_REQUIRE(ti_cur != ti_next && ti_cur != solution && ti_next != solution, "solution, ti_cur and ti_next must be three different vectors");
  hd_mesh *m = rk->mesh;
  HD_REQUIRE(m == op->mesh, "integrator and operator belong to different meshes");
  HD_REQUIRE(!op->d_av, "fused stages need a constant velocity");
  HD_CUDA(cudaSetDevice(m->ctx->device));
  const bool fast = op->kernel_choice == 2 || op->kernel_choice == 6 || (op->kernel_choice == 0 && hd::fast6d_supported(op));
  if (!fast)
    return hd::fail(HD_ERR_UNSUPPORTED, "hd_lsrk_stage_overlapped needs one of the 3D3V degree-3 FP64 kernels; use hd_halo_pack + hd_lsrk_stage_fused");
  for (int d = 0; d < m->dim; ++d)
    for (int sd = 0; sd < 2; ++sd)
      if (m->d.side_kind[d][sd] == HD_SIDE_DIRICHLET)
        return hd::fail(HD_ERR_UNSUPPORTED, "hd_lsrk_stage_overlapped: Dirichlet sides are not supported");
  double      c;
  FusedUpdate fu;
  fu.enabled = 1;
  fu.sol     = solution;
  fu.ti_next = ti_next;
  lsrk_stage_factors(rk, stage, dt, &c, &fu.fb, &fu.fa);
  return hd::launch_fast6d(op, nullptr, ti_cur, ghosts, t + c * dt, fu, 3, sends, n_sends, arrival_counters, target);
}

// ---- VectorTools ------------------------------------------------------------------------------
int64_t
hd_mesh_n_dofs_x(const hd_mesh *m)
{
  if (!m)
    return 0;
  int64_t n = 1;
  for (int d = 0; d < m->d.dim_x; ++d)
    n *= (int64_t)m->d.n_cells[d] * m->n;
  return n;
}

int
hd_vector_alloc_x(hd_mesh *m, void **ptr)
{
  HD_REQUIRE(m && ptr, "null argument");
  HD_CUDA(cudaSetDevice(m->ctx->device));
  const size_t bytes = (size_t)hd_mesh_n_dofs_x(m) * m->elem_size;
  HD_CUDA(cudaMalloc(ptr, bytes ? bytes : 16));
  if (cudaMemsetAsync(*ptr, 0, bytes, m->ctx->stream) != cudaSuccess)
    {
      cudaFree(*ptr);
      *ptr = nullptr;
      return hd::fail(HD_ERR_CUDA, "hd_vector_alloc_x: cudaMemsetAsync failed");
    }
  {
    std::lock_guard<std::mutex> lock(m->vectors_mutex);
    m->vectors[*ptr] = hd_mesh_n_dofs_x(m);
  }
  return HD_OK;
}

int
hd_velocity_space_integration(hd_mesh *m, void *dst_x, const void *src)
{
  HD_REQUIRE(m && dst_x && src, "null argument");
  HD_REQUIRE(m->d.dim_v >= 1, "no velocity space");
  HD_CUDA(cudaSetDevice(m->ctx->device));
  const long long nx = hd_mesh_n_dofs_x(m);
  long long       ncx = 1, ncv = 1;
  int             ndx = 1, ndv = 1;
  for (int d = 0; d < m->dim; ++d)
    {
      if (d < m->d.dim_x)
        {
          ncx *= m->d.n_cells[d];
          ndx *= m->n;
        }
      else
        {
          ncv *= m->d.n_cells[d];
          ndv *= m->n;
        }
    }
  HD_CUDA(cudaMemsetAsync(dst_x, 0, (size_t)nx * m->elem_size, m->ctx->stream));
  const unsigned gx   = (unsigned)((nx + 255) / 256);
  long long      want = (8ll * m->ctx->sm_count + gx - 1) / gx; // about 8 CTAs per SM in total
  if (want > ncv)
    want = ncv;
  if (want < 1)
    want = 1;
  const dim3 grid(gx, (unsigned)want);
  if (m->d.number_type == HD_F64)
    k_velocity_space_integration<double><<<grid, 256, 0, m->ctx->stream>>>(static_cast<const double *>(src), static_cast<double *>(dst_x), m->d_wv, nx, ncx, ncv, ndx, ndv);
  else
    k_velocity_space_integration<float><<<grid, 256, 0, m->ctx->stream>>>(static_cast<const float *>(src), static_cast<float *>(dst_x), m->d_wv, nx, ncx, ncv, ndx, ndv);
  HD_CUDA(cudaGetLastError());
  return HD_OK;
}

int
hd_interpolate_builtin(hd_mesh *m, void *vec, int fn_id, double time)
{
  HD_REQUIRE(m && vec, "null argument");
  HD_CUDA(cudaSetDevice(m->ctx->device));
  const LatticeParams lp  = lattice(m);
  if (m->n == 4 && (m->dim == 2 || m->dim == 4 || m->dim == 6))
    {
      // degree 3: the streaming kernel (>= 4096 nodal values per CTA)
      const int       cpb  = m->nd >= 4096 ? 1 : int(4096 / m->nd);
      const size_t    smem = (size_t)cpb * m->dim * 4 * sizeof(double);
      const long long g    = (m->ncells + cpb - 1) / cpb;
#define HD_INTERP4(T, DIM) k_interpolate4<T, DIM><<<(unsigned)g, 256, smem, m->ctx->stream>>>(static_cast<T *>(vec), lp, m->d_basis, fn_id, time, cpb)
      if (m->d.number_type == HD_F64)
        {
          if (m->dim == 2)
            HD_INTERP4(double, 2);
          else if (m->dim == 4)
            HD_INTERP4(double, 4);
          else
            HD_INTERP4(double, 6);
        }
      else
        {
          if (m->dim == 2)
            HD_INTERP4(float, 2);
          else if (m->dim == 4)
            HD_INTERP4(float, 4);
          else
            HD_INTERP4(float, 6);
        }
#undef HD_INTERP4
      HD_CUDA(cudaGetLastError());
      return HD_OK;
    }
  const int           cpb = m->nd >= 1024 ? 1 : int(1024 / m->nd); // >= 1024 nodal values per CTA
  const size_t        smem = (size_t)cpb * m->dim * m->n * sizeof(double);
  const long long     g    = (m->ncells + cpb - 1) / cpb;
  if (m->d.number_type == HD_F64)
    k_interpolate<double><<<(unsigned)g, 256, smem, m->ctx->stream>>>(static_cast<double *>(vec), lp, m->d_basis, fn_id, time, cpb);
  else
    k_interpolate<float><<<(unsigned)g, 256, smem, m->ctx->stream>>>(static_cast<float *>(vec), lp, m->d_basis, fn_id, time, cpb);
  HD_CUDA(cudaGetLastError());
  return HD_OK;
}

int
hd_norm_and_error_builtin(hd_mesh *m, const void *vec, int fn_id, double time, double out[2])
{
  HD_REQUIRE(m && vec && out, "null argument");
  HD_CUDA(cudaSetDevice(m->ctx->device));
  const LatticeParams lp = lattice(m);
  int                 mx = m->n > m->nq ? m->n : m->nq;
  size_t              cap = 1;
  for (int d = 0; d < m->dim; ++d)
    cap *= mx;
  const size_t smem = 2 * cap * sizeof(double);
  if (smem > m->ctx->smem_optin)
    return hd::fail(HD_ERR_UNSUPPORTED, "norm_and_error: cell does not fit into shared memory");
  HD_CUDA(cudaMemsetAsync(m->d_reduce, 0, 2 * sizeof(double), m->ctx->stream));
  if (m->dim == 6 && m->n == 4 && m->nq == 4)
    {
      // 3D3V degree 3 (the benchmark lattice): the register-tile kernel, a few CTAs per SM walking over the cells
      const long long want = (long long)m->ctx->sm_count * 2; // (124 registers x 256 threads: two resident CTAs per SM)
      const unsigned  g    = (unsigned)(m->ncells < want ? m->ncells : want);
      if (m->d.number_type == HD_F64)
        k_norm_error_3d3v<double><<<g, 256, 0, m->ctx->stream>>>(static_cast<const double *>(vec), lp, m->d_basis, fn_id, time, m->d_reduce);
      else
        k_norm_error_3d3v<float><<<g, 256, 0, m->ctx->stream>>>(static_cast<const float *>(vec), lp, m->d_basis, fn_id, time, m->d_reduce);
    }
  else if (m->d.number_type == HD_F64)
    {
      if (smem > 48 * 1024)
        HD_CUDA(cudaFuncSetAttribute(k_norm_error<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_norm_error<double><<<(unsigned)m->ncells, 256, smem, m->ctx->stream>>>(static_cast<const double *>(vec), lp, m->d_basis, fn_id, time, m->d_reduce);
    }
  else
    {
      if (smem > 48 * 1024)
        HD_CUDA(cudaFuncSetAttribute(k_norm_error<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_norm_error<float><<<(unsigned)m->ncells, 256, smem, m->ctx->stream>>>(static_cast<const float *>(vec), lp, m->d_basis, fn_id, time, m->d_reduce);
    }
  HD_CUDA(cudaGetLastError());
  HD_CUDA(cudaMemcpyAsync(out, m->d_reduce, 2 * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
  HD_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return HD_OK;
}

} // extern "C"
