// Internal declarations shared by the translation units of libhdgpu.so (not installed).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/hyperdeal_b200.h"
#include "basis.hpp"

#define HD_CUDA(call)                                                                              \
  do                                                                                               \
    {                                                                                              \
      cudaError_t e_ = (call);                                                                     \
      if (e_ != cudaSuccess)                                                                       \
        return hd::fail(HD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    }                                                                                              \
  while (0)

#define HD_REQUIRE(cond, msg)                                                                      \
  do                                                                                               \
    {                                                                                              \
      if (!(cond))                                                                                 \
        return hd::fail(HD_ERR_INVALID, std::string(msg) + " (" #cond ")");                        \
    }                                                                                              \
  while (0)

namespace hd
{
  int fail(int code, const std::string &msg);
}

struct hd_context
{
  int          device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t  ev0 = nullptr, ev1 = nullptr;
  int          sm_count = 0;
  size_t       smem_optin = 0;
  bool         pack_carveout_set = false; // k_halo_pack's shared-memory carve-out has been set on this device
};

struct hd_mesh
{
  hd_context * ctx = nullptr;
  hd_mesh_desc d;
  int          dim = 0, n = 0, nq = 0;
  int64_t      nd = 0;     // DoFs per cell  n^dim
  int64_t      nf = 0;     // DoFs per face  n^(dim-1)
  int64_t      ncells = 0; // local cells
  int64_t      ndofs  = 0; // local DoFs
  double       h[HD_MAX_DIM];
  hd::Basis1D  basis;
  size_t       elem_size = 8;
  int64_t      ghost_off[HD_MAX_DIM][2];
  int64_t      ghost_cnt[HD_MAX_DIM][2];
  int64_t      ghost_total = 0;
  bool         has_ghosts  = false;
  bool         has_dirichlet = false;
  // device copies (double) of nodes[n], xq[nq], w[nq], S[nq*n]
  double *d_basis = nullptr;
  double *d_reduce = nullptr; // 2 doubles for norm reductions
  double *d_wv     = nullptr; // n^dim_v doubles: Gauss-Lobatto JxW of one v-cell (velocity_space_integration)
  // vectors handed out by hd_vector_alloc / hd_vector_alloc_x: device pointer -> number of values it holds.  The copy / zero
  // entry points check their element counts against it (pointers the library did not allocate are taken on trust).
  std::map<const void *, int64_t> vectors;
  std::mutex                      vectors_mutex;
  int64_t
  vector_values(const void *p)
  {
    std::lock_guard<std::mutex> lock(vectors_mutex);
    auto                        it = vectors.find(p);
    return it == vectors.end() ? -1 : it->second;
  }
};

// per-direction collapsed matrices as uploaded to the device (T = Number)
template <typename T, int N>
struct DirCoef
{
  T C[4][N * N];
  T L0[N];
  T L1[N];
};

struct hd_advection
{
  hd_mesh *   mesh = nullptr;
  double      skew = 0;
  double      a[HD_MAX_DIM];
  void *      d_coef      = nullptr; // DirCoef<T,N>[dim]
  size_t      coef_bytes  = 0;
  int         nb_mask[HD_MAX_DIM]; // bit0: lower neighbour trace needed, bit1: upper
  int         kernel_choice = 0;
  int         eval_level    = 0; // HD_EVAL_*
  const char *last_kernel   = "none";
  int64_t     launches      = 0;
  // host copies of the collapsed matrices in double (for the specialised kernels)
  std::vector<double> hC[HD_MAX_DIM][4], hL0[HD_MAX_DIM], hL1[HD_MAX_DIM];
  // Dirichlet data
  int     dirichlet_fn = -1;                    // built-in id or -1
  double *d_g[HD_MAX_DIM][2];                   // device g at face quadrature points (double)
  int64_t g_count[HD_MAX_DIM][2];
  // staging for apply_host
  void *  d_stage_src = nullptr, *d_stage_dst = nullptr;
  // apply_host pipeline: copy-in / copy-out streams and per-slab events (slabs = layers of the slowest direction)
  cudaStream_t             s_h2d = nullptr, s_d2h = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_done;
  // phase-space velocity field (kernel_vp.cu): caller-owned a_v table on the device (nullptr = constant velocity) and the
  // speed-independent coefficient block
  const void *d_av      = nullptr;
  void *      d_vp_coef = nullptr;
  std::vector<double> h_vp_coef; // host copy (the tile kernels take their matrices as a kernel parameter)
  // Dirichlet lattices on the fast kernels: the same lattice with its inflow Dirichlet sides declared HD_SIDE_GHOST, an
  // operator on it, and the ghost buffer that receives -u_face + 2 g before every application (capi.cu: apply_dirichlet_as_ghosts)
  hd_mesh *     shadow_mesh  = nullptr;
  hd_advection *shadow_op    = nullptr;
  void *        d_shadow_ghost = nullptr;
  int           shadow_state = 0; // 0 = not tried, 1 = in use, -1 = not applicable
  // fast-kernel private state (tensor maps etc.)
  void *fast_state = nullptr;
  int   row_tile[5] = {-1, -1, -1, -1, -1}; // pipelined kernel: row tile per direction 1..5 (-1 = default, 0 = full extent)
  int   halo_senders = 0; // fused-halo kernel: CTAs that pack and send (0 = default, env HD_HALO_SENDERS or 32)
  int   l2_hints   = -1; // pipelined kernel: L2 residency hints (bit mask), -1 = default (env HD_L2_HINTS or all)
};

struct hd_lsrk
{
  hd_mesh *           mesh = nullptr;
  std::vector<double> bi, ai;
  void *              d_ti2 = nullptr; // second Ti register for the fused path
};

// optional fused LSRK epilogue:  K = (M^-1 A src);  sol += fb*K;  if (fa != 0) ti_next = sol_old + fa*K
struct FusedUpdate
{
  void * sol     = nullptr;
  void * ti_next = nullptr;
  double fb = 0, fa = 0;
  int    enabled = 0;
};

namespace hd
{
  // kernels_generic.cu
  int launch_generic(hd_advection *op, void *dst, const void *src, const void *ghosts, double time, const FusedUpdate &fu);
  bool tile_supported(const hd_advection *op);
  bool tile_preferred(const hd_advection *op);
  bool tile_row_supported(const hd_advection *op);
  int  launch_tile_row(hd_advection *op, void *dst, const void *src, const void *ghosts, double time, const FusedUpdate &fu);
  int  launch_tile(hd_advection *op, void *dst, const void *src, const void *ghosts, double time, const FusedUpdate &fu);
  // kernel_fast6d.cu
  bool fast6d_supported(const hd_advection *op);
  bool fast6d_use_rounds(const hd_advection *op);
  int  fast6d_halo_senders(const hd_advection *op);
  int  launch_fast6d(hd_advection *op, void *dst, const void *src, const void *ghosts, double time, const FusedUpdate &fu, int part,
                     const hd_halo_send *sends = nullptr, int n_sends = 0, const void *halo_flag = nullptr, int halo_target = 0, long long row_begin = 0,
                     long long row_end = -1);
  int  fast6d_overlap_status(hd_advection *op, int *timed_out);
  void fast6d_release(hd_advection *op);
  // kernel_tile_global.cu
  bool tile_global_supported(const hd_advection *op);
  int  launch_tile_global(hd_advection *op, void *dst, const void *src, const void *ghosts, double time, const FusedUpdate &fu);
  // kernel_vp.cu
  bool vp_supported(const hd_advection *op, std::string *why);
  int  vp_upload_coefficients(hd_advection *op);
  int  launch_vp(hd_advection *op, void *dst, const void *src, double time, const FusedUpdate &fu);
  // dirichlet source term (kernels_generic.cu)
  int launch_dirichlet_source(hd_advection *op, void *dst, double time, const FusedUpdate &fu);
  int launch_dirichlet_ghosts(hd_advection *op, const hd_mesh *ghost_mesh, void *ghosts, const void *src, double time);
} // namespace hd
