// hd_multi_*: one phase-space lattice cut into bricks on several GPUs of ONE process.
//
// This is the C++-host route to more than one GPU (the Python/torch.distributed route is one process per GPU,
// hyperdeal_b200/partition.py): the reference builds its process grid PartitionX x PartitionV in C++
// (performance/util/driver.h:133-161, examples/advection/advection.cc:82-88) and exchanges ghost faces through MPI-3
// shared-memory windows (matrix_free/vector_partitioner.h:1387-1692).  Here every brick lives on its own device, peer
// access is enabled between all of them (NVLink/NVSwitch: any GPU reaches any other at full bandwidth), and
//   * the pack kernel of the SENDER stores the boundary layers straight into the receiver's ghost buffer (hd_halo_pack_ex
//     with peer pointers) — pack loop + MPI_Isend of export_to_ghosted_array_start in one step;
//   * CUDA events carry the ordering across devices (receiver waits for the senders' pack kernels, senders wait until the
//     receiver's previous use of that ghost buffer is over; two ghost buffers per brick) — no host synchronisation anywhere;
//   * norms are summed on the host (the reference's MPI_Allreduce of two doubles).
// Everything is built on the single-brick C ABI; nothing here launches a kernel of its own.
#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "hd_internal.h"

struct hd_multi
{
  int                             n = 0;
  int                             grid[HD_MAX_DIM];
  hd_mesh_desc                    global;
  std::vector<int>                device;
  std::vector<hd_context *>       ctx;
  std::vector<hd_mesh *>          mesh;
  std::vector<cudaStream_t>       stream;
  std::vector<std::array<int, HD_MAX_DIM>> coords;
  std::vector<void *>             ghost[2]; // two ghost buffers per brick
  std::vector<cudaEvent_t>        ev_packed, ev_done[2];
  std::vector<bool>               done_valid[2];
  long long                       exchanges = 0;
  int                             dim = 0;
  // fused halo (3D3V degree-3 FP64 kernels): arrival counters of every brick, slot 2 * d + side (peer-written), and the
  // number of fused applications so far (the counters are never reset: target = applications * sender CTAs)
  std::vector<int *> counters;
  long long          fused_apps = 0;
};

struct hd_multi_advection
{
  hd_multi *                  mm = nullptr;
  std::vector<hd_advection *> op;
  std::vector<std::array<int, 2 * HD_MAX_DIM>> needed; // per brick: ghost sides its operator reads
  bool                                         fused = false; // the operator kernel packs and sends the halo itself
};

struct hd_multi_lsrk
{
  hd_multi *             mm = nullptr;
  std::vector<hd_lsrk *> rk;
  int                    stages = 0;
};

namespace
{
  int
  brick_of(const hd_multi *mm, const std::array<int, HD_MAX_DIM> &c)
  {
    int r = 0, m = 1;
    for (int d = 0; d < mm->dim; ++d)
      {
        r += c[d] * m;
        m *= mm->grid[d];
      }
    return r;
  }

  // brick behind side (d, side) of brick i, or -1 at a non-periodic domain boundary
  int
  neighbour(const hd_multi *mm, int i, int d, int side)
  {
    auto c = mm->coords[i];
    c[d] += side ? 1 : -1;
    if (c[d] < 0 || c[d] >= mm->grid[d])
      {
        if (mm->global.side_kind[d][0] != HD_SIDE_PERIODIC_LOCAL)
          return -1;
        c[d] = (c[d] + mm->grid[d]) % mm->grid[d];
      }
    return brick_of(mm, c);
  }

  // fill ghost buffer `buf` of every brick with the ghost faces of src[]: senders pack into the receivers' buffers
  int
  exchange(hd_multi *mm, const std::vector<std::array<int, 2 * HD_MAX_DIM>> &needed, void *const *src, int buf)
  {
    const int n = mm->n;
    for (int j = 0; j < n; ++j)
      {
        // brick j's boundary layer (d, s) is the ghost side (d, 1-s) of the neighbour behind (d, s)
        int   mask[2 * HD_MAX_DIM];
        void *peer[2 * HD_MAX_DIM];
        bool  any = false;
        HD_CUDA(cudaSetDevice(mm->device[j]));
        for (int d = 0; d < HD_MAX_DIM; ++d)
          for (int s = 0; s < 2; ++s)
            {
              mask[2 * d + s] = 0;
              peer[2 * d + s] = nullptr;
              if (d >= mm->dim || mm->grid[d] == 1)
                continue;
              const int i = neighbour(mm, j, d, s);
              if (i < 0 || !needed[i][2 * d + (1 - s)])
                continue;
              mask[2 * d + s] = 1;
              peer[2 * d + s] = static_cast<char *>(mm->ghost[buf][i]) + (size_t)hd_halo_offset(mm->mesh[i], d, 1 - s) * mm->mesh[i]->elem_size;
              // the receiver must be done with its previous use of this buffer
              if (mm->done_valid[buf][i])
                HD_CUDA(cudaStreamWaitEvent(mm->stream[j], mm->ev_done[buf][i], 0));
              any = true;
            }
        if (any)
          {
            int rc = hd_halo_pack_ex(mm->mesh[j], src[j], nullptr, mask, peer, nullptr, nullptr);
            if (rc != HD_OK)
              return rc;
          }
        HD_CUDA(cudaEventRecord(mm->ev_packed[j], mm->stream[j]));
      }
    for (int i = 0; i < n; ++i)
      {
        HD_CUDA(cudaSetDevice(mm->device[i]));
        for (int d = 0; d < mm->dim; ++d)
          for (int s = 0; s < 2; ++s)
            {
              if (mm->grid[d] == 1 || !needed[i][2 * d + s])
                continue;
              const int j = neighbour(mm, i, d, s);
              if (j >= 0 && j != i)
                HD_CUDA(cudaStreamWaitEvent(mm->stream[i], mm->ev_packed[j], 0));
            }
      }
    mm->exchanges++;
    return HD_OK;
  }

  int
  mark_done(hd_multi *mm, int buf)
  {
    for (int i = 0; i < mm->n; ++i)
      {
        HD_CUDA(cudaSetDevice(mm->device[i]));
        HD_CUDA(cudaEventRecord(mm->ev_done[buf][i], mm->stream[i]));
        mm->done_valid[buf][i] = true;
      }
    return HD_OK;
  }
  // One operator application (or fused LSRK stage: rk != nullptr) on all bricks with the halo INSIDE the operator kernel
  // (hd_advection_apply_overlapped / hd_lsrk_stage_overlapped): the first CTAs of brick j's kernel pack its boundary
  // layers, store them into the receivers' ghost buffers over NVLink and bump the receivers' arrival counters; the
  // kernel's boundary phase waits for its own counters.  The kernels of all bricks are launched back to back from this
  // thread (asynchronously), so every kernel finds its senders running.  Buffer reuse is ordered by events as in exchange().
  int
  fused_step(hd_multi *mm, hd_multi_advection *mop, hd_multi_lsrk *mrk, int stage, void *const *dst, void *const *src, void *const *solution, void *const *ti_next,
             double t, double dt)
  {
    const int n      = mm->n;
    const int buf    = int(mm->exchanges & 1);
    const int target = int((mm->fused_apps + 1) * hd_advection_n_halo_senders(mop->op[0]));
    for (int j = 0; j < n; ++j)
      {
        hd_halo_send sends[2 * HD_MAX_DIM];
        int          n_sends = 0;
        HD_CUDA(cudaSetDevice(mm->device[j]));
        for (int d = 0; d < mm->dim; ++d)
          for (int s = 0; s < 2; ++s)
            {
              if (mm->grid[d] == 1)
                continue;
              const int i = neighbour(mm, j, d, s);
              if (i < 0 || !mop->needed[i][2 * d + (1 - s)])
                continue;
              sends[n_sends].dir  = d;
              sends[n_sends].side = s;
              sends[n_sends].dst  = static_cast<char *>(mm->ghost[buf][i]) + (size_t)hd_halo_offset(mm->mesh[i], d, 1 - s) * mm->mesh[i]->elem_size;
              sends[n_sends].arrival_counter = mm->counters[i] + (2 * d + (1 - s));
              ++n_sends;
              if (mm->done_valid[buf][i])
                HD_CUDA(cudaStreamWaitEvent(mm->stream[j], mm->ev_done[buf][i], 0));
            }
        int rc;
        if (mrk)
          rc = hd_lsrk_stage_overlapped(mrk->rk[j], mop->op[j], stage, solution[j], src[j], ti_next[j], mm->ghost[buf][j], sends, n_sends, mm->counters[j], target, t, dt);
        else
          rc = hd_advection_apply_overlapped(mop->op[j], dst[j], src[j], mm->ghost[buf][j], t, sends, n_sends, mm->counters[j], target);
        if (rc != HD_OK)
          return rc;
      }
    mm->exchanges++;
    mm->fused_apps++;
    return mark_done(mm, buf);
  }
} // namespace

extern "C" {

static int
multi_init_bricks(hd_multi *mm, const hd_mesh_desc *global)
{
  const int n_gpus = mm->n;
  mm->ctx.assign(n_gpus, nullptr);
  mm->mesh.assign(n_gpus, nullptr);
  mm->stream.assign(n_gpus, nullptr);
  mm->coords.resize(n_gpus);
  mm->ev_packed.assign(n_gpus, nullptr);
  mm->counters.assign(n_gpus, nullptr);
  for (int b = 0; b < 2; ++b)
    {
      mm->ghost[b].assign(n_gpus, nullptr);
      mm->ev_done[b].assign(n_gpus, nullptr);
      mm->done_valid[b].assign(n_gpus, false);
    }
  for (int i = 0; i < n_gpus; ++i)
    {
      int r = i;
      for (int d = 0; d < HD_MAX_DIM; ++d)
        {
          mm->coords[i][d] = d < mm->dim ? r % mm->grid[d] : 0;
          if (d < mm->dim)
            r /= mm->grid[d];
        }
      int rc = hd_context_create(mm->device[i], &mm->ctx[i]);
      if (rc != HD_OK)
        return rc;
      HD_CUDA(cudaSetDevice(mm->device[i]));
      HD_CUDA(cudaStreamCreateWithFlags(&mm->stream[i], cudaStreamNonBlocking));
      hd_context_set_stream(mm->ctx[i], mm->stream[i]);
      hd_mesh_desc d = *global;
      for (int k = 0; k < mm->dim; ++k)
        {
          d.n_cells[k]     = global->n_cells_global[k] / mm->grid[k];
          d.cell_offset[k] = d.n_cells[k] * mm->coords[i][k];
          for (int s = 0; s < 2; ++s)
            {
              const bool domain_side = (s == 0 && mm->coords[i][k] == 0) || (s == 1 && mm->coords[i][k] == mm->grid[k] - 1);
              if (mm->grid[k] == 1)
                d.side_kind[k][s] = global->side_kind[k][s];
              else if (global->side_kind[k][0] == HD_SIDE_PERIODIC_LOCAL || !domain_side)
                d.side_kind[k][s] = HD_SIDE_GHOST;
              else
                d.side_kind[k][s] = global->side_kind[k][s];
            }
        }
      rc = hd_mesh_create(mm->ctx[i], &d, &mm->mesh[i]);
      if (rc != HD_OK)
        return rc;
      const size_t gbytes = (size_t)hd_halo_total(mm->mesh[i]) * mm->mesh[i]->elem_size;
      for (int b = 0; b < 2; ++b)
        {
          HD_CUDA(cudaMalloc(&mm->ghost[b][i], gbytes ? gbytes : 256));
          HD_CUDA(cudaMemset(mm->ghost[b][i], 0, gbytes ? gbytes : 256));
          HD_CUDA(cudaEventCreateWithFlags(&mm->ev_done[b][i], cudaEventDisableTiming));
        }
      HD_CUDA(cudaEventCreateWithFlags(&mm->ev_packed[i], cudaEventDisableTiming));
      HD_CUDA(cudaMalloc(&mm->counters[i], 64 * sizeof(int)));
      HD_CUDA(cudaMemset(mm->counters[i], 0, 64 * sizeof(int)));
    }
  return HD_OK;
}

int
hd_multi_create(int n_gpus, const int *devices, const hd_mesh_desc *global, const int *grid, hd_multi **out)
{
  HD_REQUIRE(n_gpus >= 1 && global && grid && out, "bad argument");
  int count = 0;
  HD_CUDA(cudaGetDeviceCount(&count));
  HD_REQUIRE(n_gpus <= count, "more bricks than CUDA devices in this process");
  hd_multi *mm = new (std::nothrow) hd_multi;
  HD_REQUIRE(mm, "out of memory");
  mm->n      = n_gpus;
  mm->global = *global;
  mm->dim    = global->dim_x + global->dim_v;
  int prod   = 1;
  for (int d = 0; d < HD_MAX_DIM; ++d)
    {
      mm->grid[d] = d < mm->dim ? grid[d] : 1;
      if (mm->grid[d] < 1 || (d < mm->dim && global->n_cells_global[d] % mm->grid[d] != 0))
        {
          delete mm;
          return hd::fail(HD_ERR_INVALID, "hd_multi_create: the brick grid must divide the lattice in every direction");
        }
      prod *= mm->grid[d];
    }
  if (prod != n_gpus)
    {
      delete mm;
      return hd::fail(HD_ERR_INVALID, "hd_multi_create: product of the brick grid != number of GPUs");
    }
  mm->device.resize(n_gpus);
  for (int i = 0; i < n_gpus; ++i)
    mm->device[i] = devices ? devices[i] : i;
  // peer access between all pairs (NVLink/NVSwitch); "already enabled" is fine
  for (int i = 0; i < n_gpus; ++i)
    for (int j = 0; j < n_gpus; ++j)
      if (i != j)
        {
          int can = 0;
          HD_CUDA(cudaDeviceCanAccessPeer(&can, mm->device[i], mm->device[j]));
          if (!can)
            {
              delete mm;
              return hd::fail(HD_ERR_UNSUPPORTED, "hd_multi_create: no peer access between the GPUs");
            }
          HD_CUDA(cudaSetDevice(mm->device[i]));
          cudaError_t e = cudaDeviceEnablePeerAccess(mm->device[j], 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            {
              delete mm;
              return hd::fail(HD_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            }
          cudaGetLastError();
        }
  const int rc = multi_init_bricks(mm, global);
  if (rc != HD_OK)
    {
      hd_multi_destroy(mm); // tolerates a partially built object
      return rc;
    }
  *out = mm;
  return HD_OK;
}

int
hd_multi_destroy(hd_multi *mm)
{
  if (!mm)
    return HD_OK;
  for (int i = 0; i < mm->n; ++i)
    {
      cudaSetDevice(mm->device[i]);
      cudaDeviceSynchronize();
      for (int b = 0; b < 2; ++b)
        {
          cudaFree(mm->ghost[b][i]);
          if (mm->ev_done[b][i])
            cudaEventDestroy(mm->ev_done[b][i]);
        }
      if (mm->ev_packed[i])
        cudaEventDestroy(mm->ev_packed[i]);
      cudaFree(mm->counters[i]);
      hd_mesh_destroy(mm->mesh[i]);
      hd_context_destroy(mm->ctx[i]);
      if (mm->stream[i])
        cudaStreamDestroy(mm->stream[i]);
    }
  delete mm;
  return HD_OK;
}

int
hd_multi_n_gpus(const hd_multi *mm)
{
  return mm ? mm->n : 0;
}

hd_mesh *
hd_multi_mesh(hd_multi *mm, int i)
{
  return (mm && i >= 0 && i < mm->n) ? mm->mesh[i] : nullptr;
}

hd_context *
hd_multi_context(hd_multi *mm, int i)
{
  return (mm && i >= 0 && i < mm->n) ? mm->ctx[i] : nullptr;
}

int64_t
hd_multi_n_dofs(const hd_multi *mm)
{
  int64_t n = 0;
  if (mm)
    for (int i = 0; i < mm->n; ++i)
      n += hd_mesh_n_dofs(mm->mesh[i]);
  return n;
}

int
hd_multi_synchronize(hd_multi *mm)
{
  HD_REQUIRE(mm, "null argument");
  for (int i = 0; i < mm->n; ++i)
    {
      HD_CUDA(cudaSetDevice(mm->device[i]));
      HD_CUDA(cudaStreamSynchronize(mm->stream[i]));
    }
  return HD_OK;
}

int
hd_multi_vector_alloc(hd_multi *mm, void **ptrs)
{
  HD_REQUIRE(mm && ptrs, "null argument");
  for (int i = 0; i < mm->n; ++i)
    {
      int rc = hd_vector_alloc(mm->mesh[i], 0, &ptrs[i]);
      if (rc != HD_OK)
        return rc;
    }
  return HD_OK;
}

int
hd_multi_vector_free(hd_multi *mm, void *const *ptrs)
{
  HD_REQUIRE(mm && ptrs, "null argument");
  for (int i = 0; i < mm->n; ++i)
    hd_vector_free(mm->mesh[i], ptrs[i]);
  return HD_OK;
}

// scatter / gather between the bricks and ONE host vector in the layout of the unpartitioned lattice (global cells
// lexicographic, direction 0 fastest): what a single-GPU run of the same lattice would hold
static int
copy_global(hd_multi *mm, void *const *ptrs, void *host, bool to_device)
{
  const size_t es = mm->mesh[0]->elem_size;
  const size_t nd = (size_t)mm->mesh[0]->nd;
  std::vector<char> tmp;
  for (int i = 0; i < mm->n; ++i)
    {
      hd_mesh *     m  = mm->mesh[i];
      const int64_t nc = m->ncells;
      tmp.resize((size_t)nc * nd * es);
      if (!to_device)
        {
          int rc = hd_vector_copy_out(m, ptrs[i], tmp.data(), nc * (int64_t)nd);
          if (rc != HD_OK)
            return rc;
        }
      for (int64_t lc = 0; lc < nc; ++lc)
        {
          int64_t r = lc, g = 0, mult = 1;
          for (int d = 0; d < mm->dim; ++d)
            {
              const int c = int(r % m->d.n_cells[d]) + m->d.cell_offset[d];
              r /= m->d.n_cells[d];
              g += (int64_t)c * mult;
              mult *= mm->global.n_cells_global[d];
            }
          char *h = static_cast<char *>(host) + (size_t)g * nd * es, *l = tmp.data() + (size_t)lc * nd * es;
          if (to_device)
            std::memcpy(l, h, nd * es);
          else
            std::memcpy(h, l, nd * es);
        }
      if (to_device)
        {
          int rc = hd_vector_copy_in(m, ptrs[i], tmp.data(), nc * (int64_t)nd);
          if (rc != HD_OK)
            return rc;
        }
    }
  return HD_OK;
}

int
hd_multi_vector_copy_in(hd_multi *mm, void *const *ptrs, const void *host_global)
{
  HD_REQUIRE(mm && ptrs && host_global, "null argument");
  return copy_global(mm, ptrs, const_cast<void *>(host_global), true);
}

int
hd_multi_vector_copy_out(hd_multi *mm, void *const *ptrs, void *host_global)
{
  HD_REQUIRE(mm && ptrs && host_global, "null argument");
  return copy_global(mm, ptrs, host_global, false);
}

int
hd_multi_interpolate_builtin(hd_multi *mm, void *const *vec, int fn_id, double time)
{
  HD_REQUIRE(mm && vec, "null argument");
  for (int i = 0; i < mm->n; ++i)
    {
      int rc = hd_interpolate_builtin(mm->mesh[i], vec[i], fn_id, time);
      if (rc != HD_OK)
        return rc;
    }
  return HD_OK;
}

int
hd_multi_norm_and_error_builtin(hd_multi *mm, void *const *vec, int fn_id, double time, double out[2])
{
  HD_REQUIRE(mm && vec && out, "null argument");
  out[0] = out[1] = 0.0;
  for (int i = 0; i < mm->n; ++i)
    {
      double s[2];
      int    rc = hd_norm_and_error_builtin(mm->mesh[i], vec[i], fn_id, time, s);
      if (rc != HD_OK)
        return rc;
      out[0] += s[0]; // (the reference: MPI_Allreduce of the two sums, numerics/vector_tools.h:210-216)
      out[1] += s[1];
    }
  return HD_OK;
}

int
hd_multi_advection_create(hd_multi *mm, double skew_factor, const double *velocity, hd_multi_advection **out)
{
  HD_REQUIRE(mm && velocity && out, "null argument");
  hd_multi_advection *mop = new (std::nothrow) hd_multi_advection;
  HD_REQUIRE(mop, "out of memory");
  mop->mm = mm;
  mop->op.assign(mm->n, nullptr);
  for (int i = 0; i < mm->n; ++i)
    {
      int rc = hd_advection_create(mm->mesh[i], skew_factor, velocity, &mop->op[i]);
      if (rc != HD_OK)
        {
          hd_multi_advection_destroy(mop);
          return rc;
        }
    }
  mop->needed.resize(mm->n);
  for (int i = 0; i < mm->n; ++i)
    hd_advection_ghost_sides(mop->op[i], mop->needed[i].data());
  // the fused halo needs one of the 3D3V degree-3 FP64 kernels, ghost (not Dirichlet) sides, and every brick sending
  // something (a kernel that waits for counters nobody bumps would wait for its time-out); HD_MULTI_FUSED=0 keeps the
  // pack kernels + events of exchange()
  {
    const char *e = getenv("HD_MULTI_FUSED");
    bool ok = (!e || atoi(e) != 0) && mm->n > 1 && hd_advection_n_halo_senders(mop->op[0]) > 0 && !mm->mesh[0]->has_dirichlet;
    for (int i = 0; i < mm->n && ok; ++i)
      {
        bool any = false;
        for (int k = 0; k < 2 * HD_MAX_DIM; ++k)
          any |= mop->needed[i][k] != 0;
        ok = any && !mm->mesh[i]->has_dirichlet;
      }
    mop->fused = ok;
  }
  *out = mop;
  return HD_OK;
}

int
hd_multi_advection_destroy(hd_multi_advection *mop)
{
  if (!mop)
    return HD_OK;
  for (auto *op : mop->op)
    hd_advection_destroy(op);
  delete mop;
  return HD_OK;
}

int
hd_multi_advection_set_dirichlet_builtin(hd_multi_advection *mop, int fn_id)
{
  HD_REQUIRE(mop, "null argument");
  for (auto *op : mop->op)
    {
      int rc = hd_advection_set_dirichlet_builtin(op, fn_id);
      if (rc != HD_OK)
        return rc;
    }
  return HD_OK;
}

const char *
hd_multi_advection_kernel_name(const hd_multi_advection *mop)
{
  return mop ? hd_advection_kernel_name(mop->op[0]) : "none";
}

int
hd_multi_advection_apply(hd_multi_advection *mop, void *const *dst, void *const *src, double time)
{
  HD_REQUIRE(mop && dst && src, "null argument");
  hd_multi *mm  = mop->mm;
  if (mop->fused)
    return fused_step(mm, mop, nullptr, 0, dst, src, nullptr, nullptr, time, 0.0);
  const int buf = int(mm->exchanges & 1);
  int       rc  = exchange(mm, mop->needed, src, buf);
  if (rc != HD_OK)
    return rc;
  for (int i = 0; i < mm->n; ++i)
    if ((rc = hd_advection_apply(mop->op[i], dst[i], src[i], mm->ghost[buf][i], time)) != HD_OK)
      return rc;
  return mark_done(mm, buf);
}

int
hd_multi_lsrk_create(hd_multi *mm, const char *type, hd_multi_lsrk **out)
{
  HD_REQUIRE(mm && type && out, "null argument");
  hd_multi_lsrk *mrk = new (std::nothrow) hd_multi_lsrk;
  HD_REQUIRE(mrk, "out of memory");
  mrk->mm = mm;
  mrk->rk.assign(mm->n, nullptr);
  for (int i = 0; i < mm->n; ++i)
    {
      int rc = hd_lsrk_create(mm->mesh[i], type, &mrk->rk[i]);
      if (rc != HD_OK)
        return rc;
    }
  mrk->stages = hd_lsrk_n_stages(mrk->rk[0]);
  *out        = mrk;
  return HD_OK;
}

int
hd_multi_lsrk_destroy(hd_multi_lsrk *mrk)
{
  if (!mrk)
    return HD_OK;
  for (auto *rk : mrk->rk)
    hd_lsrk_destroy(rk);
  delete mrk;
  return HD_OK;
}

// perform_time_step (time_integrators.templates.h:93-184) on all bricks: per stage one ghost exchange of the current Ti and
// one fused operator + update launch per brick (K never stored; Ti ping-pongs between vec_Ti and vec_Ki as in hd_lsrk_step)
int
hd_multi_lsrk_step(hd_multi_lsrk *mrk, hd_multi_advection *mop, void *const *solution, void *const *vec_Ki, void *const *vec_Ti, double t, double dt)
{
  HD_REQUIRE(mrk && mop && solution && vec_Ki && vec_Ti, "null argument");
  hd_multi *mm = mrk->mm;
  HD_REQUIRE(mm == mop->mm, "integrator and operator belong to different lattices");
  const int n = mm->n;
  int       rc;
  std::vector<void *> cur(n), nxt(n);
  for (int i = 0; i < n; ++i)
    {
      if ((rc = hd_vector_copy(mm->mesh[i], vec_Ti[i], solution[i])) != HD_OK)
        return rc;
      cur[i] = vec_Ti[i];
      nxt[i] = vec_Ki[i];
    }
  for (int stage = 0; stage < mrk->stages; ++stage)
    {
      if (mop->fused)
        {
          if ((rc = fused_step(mm, mop, mrk, stage, nullptr, cur.data(), solution, nxt.data(), t, dt)) != HD_OK)
            return rc;
          cur.swap(nxt);
          continue;
        }
      const int buf = int(mm->exchanges & 1);
      if ((rc = exchange(mm, mop->needed, cur.data(), buf)) != HD_OK)
        return rc;
      for (int i = 0; i < n; ++i)
        if ((rc = hd_lsrk_stage_fused(mrk->rk[i], mop->op[i], stage, solution[i], cur[i], nxt[i], mm->ghost[buf][i], t, dt)) != HD_OK)
          return rc;
      if ((rc = mark_done(mm, buf)) != HD_OK)
        return rc;
      cur.swap(nxt);
    }
  return HD_OK;
}

} // extern "C"
