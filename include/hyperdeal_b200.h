/* hyperdeal_b200.h — C ABI of libhdgpu.so, the B200 (sm_100a) implementation of
 * hyper.deal's advection hot path.
 *
 * This is the drop-in boundary: plain C, opaque handles, plain pointers and sizes,
 * int status codes and a thread-local error string.  The reference has no FFI of
 * its own (it is a header-only C++ template library), so every entry point cites the
 * reference C++ member it stands in for (paths relative to the hyper.deal source
 * tree).  The header-only C++ shim in hyperdeal_b200/cpp/ re-creates those classes
 * (hyperdeal::MatrixFree, advection::AdvectionOperation,
 * LowStorageRungeKuttaIntegrator, VectorTools, ...) on top of these calls, see
 * INTEGRATION.md.
 *
 * Conventions
 *  - every function returns HD_OK (0) or a negative error code; hd_last_error()
 *    gives the message of the last failure on the calling thread.
 *  - "device pointer" arguments are raw CUDA device addresses of this process'
 *    current device; "host" arguments are ordinary host memory.
 *  - all kernels are enqueued on the context's stream (hd_context_set_stream) and
 *    are asynchronous; only the *_host, copy_out, norm and timing calls synchronise.
 *  - DoF vector layout (identical to the reference, matrix_free.templates.h:553-562,
 *    shape_info.h:126-146): cells back to back, local cell id lexicographic with
 *    direction 0 (x_0) fastest and v_last slowest; inside a cell (k+1)^dim nodal
 *    values at the Gauss-Lobatto points, x_0 fastest.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point
 *    fails with HD_ERR_CUDA.
 */
#ifndef HYPERDEAL_B200_H
#define HYPERDEAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HD_MAX_DIM 6

enum
{
  HD_OK              = 0,
  HD_ERR_INVALID     = -1, /* bad argument / unsupported configuration */
  HD_ERR_CUDA        = -2, /* CUDA runtime or driver failure (incl. no device) */
  HD_ERR_UNSUPPORTED = -3, /* valid in the reference but not implemented here */
  HD_ERR_NO_CONVERGENCE = -4 /* an iterative solve missed its tolerance (the reference: dealii::SolverControl::NoConvergence) */
};

/* Number type of the DoF vectors (the reference's template parameter `Number`). */
enum
{
  HD_F64 = 0,
  HD_F32 = 1
};

/* Face neighbour kind on the lower/upper side of the LOCAL brick in one direction. */
enum
{
  HD_SIDE_PERIODIC_LOCAL = 0, /* neighbour is the opposite end of this brick           */
  HD_SIDE_GHOST          = 1, /* neighbour lives on another GPU: read the ghost buffer */
  HD_SIDE_DIRICHLET      = 2, /* domain boundary, u+ = -u- + 2 g  (BoundaryType::DirichletInhomogenous,
                                 operators/advection/boundary_descriptor.h:31-37)     */
  HD_SIDE_DIRICHLET_HOM  = 3  /* domain boundary, u+ = -u-        (DirichletHomogenous) */
};

/* Built-in analytic fields (evaluated on the device). */
enum
{
  HD_FN_ZERO           = 0,
  HD_FN_HYPERRECTANGLE = 1 /* sin(2 pi (x0 - a0 t)) prod cos(2 pi (xd - ad t)), a = (1, .15, -.05, 0, 0, 0);
                              examples/advection/cases/hyperrectangle.h:29-66 */
};

typedef struct hd_context   hd_context;
typedef struct hd_mesh      hd_mesh;
typedef struct hd_advection hd_advection;
typedef struct hd_lsrk      hd_lsrk;
typedef struct hd_poisson   hd_poisson;

/* Description of the Cartesian phase-space lattice owned by this process.
 * Replaces the two dealii::Triangulation/dealii::MatrixFree objects handed to
 * hyperdeal::MatrixFree::reinit (matrix_free/matrix_free.templates.h:862-1235) for the
 * mesh class the hot path is specified on (subdivided_hyper_rectangle,
 * grid/grid_generator.cc:235).  Directions 0..dim_x-1 are x, dim_x..dim_x+dim_v-1 are v. */
typedef struct hd_mesh_desc
{
  int    dim_x, dim_v;
  int    degree;      /* k; FE_DGQ(k) on Gauss-Lobatto nodes                                */
  int    n_points;    /* 1-D quadrature points n_q (k+1 unless over-integration)            */
  int    collocation; /* 1: Gauss-Lobatto quadrature (DoCollocation), requires n_q == k+1   */
  int    number_type; /* HD_F64 / HD_F32                                                    */
  double left[HD_MAX_DIM], right[HD_MAX_DIM]; /* GLOBAL domain                              */
  int    n_cells_global[HD_MAX_DIM];          /* GLOBAL cells per direction                 */
  int    n_cells[HD_MAX_DIM];                 /* cells of the LOCAL brick per direction     */
  int    cell_offset[HD_MAX_DIM];             /* first global cell of the local brick       */
  int    side_kind[HD_MAX_DIM][2];            /* HD_SIDE_* for the lower/upper brick side   */
} hd_mesh_desc;

/* ---- errors ------------------------------------------------------------------------- */
const char *hd_last_error(void);
int         hd_version(void);

/* ---- context: device + stream ------------------------------------------------------- */
/* Stands in for the (comm, comm_sm) pair of hyperdeal::MatrixFree's constructor
 * (matrix_free/matrix_free.h:108): one context per process = one GPU. */
int hd_context_create(int device, hd_context **out);
int hd_context_destroy(hd_context *ctx);
/* `stream` is a cudaStream_t (0 = legacy default stream). */
int hd_context_set_stream(hd_context *ctx, void *stream);
int hd_context_synchronize(hd_context *ctx);
int hd_device_count(int *count);
/* plain zero-initialised device buffers for host layers that do not link the CUDA runtime themselves (e.g. the gradient
 * table hd_poisson_solve fills and hd_advection_set_phase_space_velocity reads) */
int hd_device_malloc(hd_context *ctx, size_t bytes, void **device_ptr);
int hd_device_free(hd_context *ctx, void *device_ptr);

/* ---- mesh / matrix-free data -------------------------------------------------------- */
/* hyperdeal::MatrixFree::reinit (matrix_free.templates.h:862). */
int     hd_mesh_create(hd_context *ctx, const hd_mesh_desc *desc, hd_mesh **out);
int     hd_mesh_destroy(hd_mesh *mesh);
int64_t hd_mesh_n_dofs(const hd_mesh *mesh);        /* owned DoFs of this brick           */
int64_t hd_mesh_n_cells(const hd_mesh *mesh);
int     hd_mesh_dofs_per_cell(const hd_mesh *mesh);
/* number of ghost values behind side (dir, side): n_face_cells * (k+1)^(dim-1), 0 if the
 * side is not HD_SIDE_GHOST  (vector_partitioner.h:916-945). */
int64_t hd_mesh_ghost_size(const hd_mesh *mesh, int dir, int side);
/* 1-D basis data the reference takes from dealii ShapeInfo (fe_evaluation_cell.h:93-95):
 * which: 0 GLL nodes[n], 1 quad points[nq], 2 quad weights[nq], 3 S[nq*n], 4 D[nq*nq],
 * 5 Sinv[n*nq]; returns the number of doubles written (out may be NULL to query). */
int hd_mesh_basis(const hd_mesh *mesh, int which, double *out);

/* ---- DoF vectors ---------------------------------------------------------------------- */
/* hyperdeal::MatrixFree::initialize_dof_vector (matrix_free.templates.h:1369-1413): owned
 * range + (if do_ghosts) ghost-face region, zero-initialised, on the device.  The returned
 * device pointer is what every *_ptr entry point takes; free with hd_vector_free. */
int hd_vector_alloc(hd_mesh *mesh, int do_ghosts, void **device_ptr);
int hd_vector_free(hd_mesh *mesh, void *device_ptr);
int hd_vector_copy_in(hd_mesh *mesh, void *device_ptr, const void *host, int64_t n_values);
int hd_vector_copy_out(hd_mesh *mesh, const void *device_ptr, void *host, int64_t n_values);
int hd_vector_zero(hd_mesh *mesh, void *device_ptr);
/* dst = src for the owned range (device to device). */
int hd_vector_copy(hd_mesh *mesh, void *dst, const void *src);
/* The same with an explicit number of values (n_values < 0: the whole vector — the owned phase-space range, or the
 * x-space range for a vector from hd_vector_alloc_x).  Vectors allocated by this library are never accessed beyond their
 * size: a larger n_values — also in hd_vector_copy_in / hd_vector_copy_out — is HD_ERR_INVALID. */
int hd_vector_zero_n(hd_mesh *mesh, void *device_ptr, int64_t n_values);
int hd_vector_copy_n(hd_mesh *mesh, void *dst, const void *src, int64_t n_values);

/* ---- advection operator ---------------------------------------------------------------- */
/* advection::AdvectionOperation::reinit (operators/advection/advection_operation.h:98) with a
 * ConstantVelocityFieldView (operators/advection/velocity_field_view.h:69) and the SkewFactor of
 * AdvectionOperationParamters (advection_operation_parameters.h:29-41).
 * velocity[dim]: constant transport direction (x components first, then v). */
int hd_advection_create(hd_mesh *mesh, double skew_factor, const double *velocity, hd_advection **out);
int hd_advection_destroy(hd_advection *op);

/* PhaseSpaceVelocityFieldView of the Vlasov-Poisson driver (examples/vlasov_poisson/include/velocity_field_view.h:111-175)
 * instead of the constant transport direction: a_x = v at the v-space quadrature points (taken from the mesh), a_v = the
 * table a_v_device[x-cell][x-quadrature point][dim_v] (doubles on the device; x-cells and quadrature points lexicographic with
 * x_0 fastest — DerivativeContainer's layout, derivative_container.h:157-190; typically grad(phi)).  The table is read at
 * every apply and stays caller-owned; NULL returns to the constant velocity.  Needs dim_x == dim_v and a periodic single-GPU
 * lattice (HD_ERR_UNSUPPORTED otherwise).  Served by a correctness-first kernel (one CTA per cell), parity-checked, not yet tuned. */
int hd_advection_set_phase_space_velocity(hd_advection *op, const double *a_v_device);

/* AdvectionOperation::apply(dst, src, time) (advection_operation.h:137): dst = M^-1 A(src, time)
 * for the owned cells; dst is overwritten (ECL semantics, advection_operation.h:562).
 * src/dst: device pointers with the vector layout above; they must not alias.
 * ghosts: device pointer to the ghost-face values of src filled by the halo exchange
 * (layout: hd_halo_offset), or NULL when no side is HD_SIDE_GHOST. */
int hd_advection_apply(hd_advection *op, void *dst, const void *src, const void *ghosts, double time);
/* The same operator in two parts, so that the ghost-face exchange overlaps with cell work — the reference's
 * overlapping levels (MatrixFree::loop_cell_centric, matrix_free.templates.h:1516-1566: cells without remote faces
 * run between export_to_ghosted_array_start and _finish, the others after):
 *   HD_PART_INTERIOR  cells that read no ghost data (may run while the halo is in flight; `ghosts` is not read)
 *   HD_PART_BOUNDARY  the remaining cells (after the halo has arrived)
 *   HD_PART_ALL       both (what hd_advection_apply does)
 * INTERIOR followed by BOUNDARY writes every owned cell of dst exactly once. */
enum
{
  HD_PART_ALL      = 0,
  HD_PART_INTERIOR = 1,
  HD_PART_BOUNDARY = 2
};
int hd_advection_apply_part(hd_advection *op, void *dst, const void *src, const void *ghosts, double time, int part);
/* Interior cells and boundary layer in ONE launch of the pipelined 3D3V kernel (HD_ERR_UNSUPPORTED for other
 * configurations: use the two parts): the kernel works through the interior cells and starts on the boundary layer once
 * arrival_counters[2*dir+side] >= target for every ghost side the operator reads — export_to_ghosted_array_finish
 * (matrix_free/vector_partitioner.h:1482).  arrival_counters: 2*HD_MAX_DIM 32-bit words in this GPU's memory.
 *  - n_sends == 0: the ghost data is produced by someone else (hd_halo_pack_ex on another stream, a neighbour GPU) who
 *    then writes the counters, e.g. with hd_stream_write_flag.  NB: once this persistent kernel is resident no other
 *    kernel gets an SM until it ends; a pack kernel must therefore have STARTED before this launch (gate the launch
 *    with hd_stream_wait_flag on hd_halo_pack_ex's started_counter).
 *  - n_sends > 0 (fused halo): the first hd_advection_n_halo_senders CTAs of the kernel begin by storing this brick's
 *    boundary layers straight into the neighbour GPUs' ghost segments (sends[i].dst: peer-mapped pointer of the segment
 *    (dir, 1-side) of the neighbour behind side (dir, side)), each adds 1 to sends[i].arrival_counter (peer-mapped
 *    address of that neighbour's arrival_counters[2*dir + (1-side)]) when its share is out, and then joins the others
 *    on the cells — pack loop + MPI_Isend of export_to_ghosted_array_start (:1387-1460) over NVLink inside the operator
 *    kernel.  Use target = applications so far * hd_advection_n_halo_senders (equal bricks on all GPUs).
 * Re-use of a ghost buffer is the caller's business (double-buffer it, hand-shake with hd_stream_write/wait_flag).
 * The wait gives up after 4 s; hd_advection_overlap_status then reports timed_out = 1 (and the result is invalid). */
typedef struct hd_halo_send
{
  int   dir, side;        /* which boundary layer of this brick                         */
  void *dst;              /* where it goes: the neighbour's ghost segment (dir, 1-side)  */
  void *arrival_counter;  /* the neighbour's counter for that ghost side                */
} hd_halo_send;
int hd_advection_apply_overlapped(hd_advection *op, void *dst, const void *src, const void *ghosts, double time, const hd_halo_send *sends, int n_sends,
                                  const void *arrival_counters, int target);
/* CTAs of the pipelined kernel on this mesh (0 if it does not apply). */
int hd_advection_n_ctas(const hd_advection *op);
/* How many of them send the halo in the fused variant, i.e. the increments an arrival counter receives per application
 * (0 if the kernel does not apply); set_halo_senders(n) overrides the default (0 = default: env HD_HALO_SENDERS, else 32). */
int hd_advection_n_halo_senders(const hd_advection *op);
int hd_advection_set_halo_senders(hd_advection *op, int n);
int hd_advection_overlap_status(hd_advection *op, int *timed_out);
/* Stream memory operations on the context's stream (no kernel launch): "*flag = value" (the address may be peer-mapped
 * memory of another GPU) and "wait until *flag >= value".  These carry the halo handshake between GPUs: data-ready
 * flags forward, buffer-consumed flags backward (the role MPI_Isend/Irecv completion + the shared-memory window
 * barriers play in matrix_free/vector_partitioner.h:1482-1592). */
/* Peer-mapped device memory for hosts that run one PROCESS per GPU (the MPI / torch.distributed route): plain device
 * allocations exported through CUDA IPC handles, the counterpart of the MPI-3 shared-memory window of the reference
 * (matrix_free/vector_partitioner.h:552-640).  hd_device_malloc (above) returns zeroed memory; hd_ipc_export writes the 64-byte
 * handle of an allocation made by hd_device_malloc; hd_ipc_open maps another process' allocation into this one
 * (peer access over NVLink); the mapping is closed with hd_ipc_close, the allocation freed with hd_device_free. */
#define HD_IPC_HANDLE_BYTES 64
int hd_ipc_export(hd_context *ctx, const void *ptr, void *handle);
int hd_ipc_open(hd_context *ctx, const void *handle, void **ptr);
int hd_ipc_close(hd_context *ctx, void *ptr);
int hd_stream_write_flag(hd_context *ctx, void *flag_device, int value);
int hd_stream_wait_flag(hd_context *ctx, void *flag_device, int value);
/* needed[2*dir+side] = 1 if the operator reads ghost side (dir, side): with the upwind flux only the inflow side
 * of a direction is read (beta_f = 0 on the outflow side, advection_operation.h:459-467 for a constant velocity), so
 * the exchange can skip the other one.  needed has 2*HD_MAX_DIM entries. */
int hd_advection_ghost_sides(const hd_advection *op, int *needed);
/* Same call on HOST buffers: copies src in, applies, copies dst out (used for the end-to-end
 * timing and by hosts that keep their vectors in host memory). */
int hd_advection_apply_host(hd_advection *op, void *dst_host, const void *src_host, double time);
/* Select the kernel: 0 = automatic (fastest available), 1 = generic kernel, 2 = pipelined 3D3V k=3 FP64 kernel,
 * 3 = tile kernel (degree 3, 1D1V / 2D2V / 3D3V, periodic or ghost sides), 4 = its row-persistent 3D3V variant,
 * 5 = tile kernel with the partial sums in global memory (degree 3 or 5, even number of directions),
 * 6 = three-round 3D3V k=3 FP64 kernel (three light compute warps per SM sub-partition, traces read from L2).
 * With 0 a 3D3V k=3 FP64 mesh runs kernel 6 (HD_FAST_VARIANT=pipe selects 2).
 * HD_ERR_UNSUPPORTED if the kernel does not cover the mesh. */
int hd_advection_set_kernel(hd_advection *op, int which);
/* AdvectionOperationEvaluationLevel (the template parameter of AdvectionOperation::apply, advection_operation.h:37-42,
 * 134-209): profiling variants that attribute the operator's time.  HD_EVAL_CELL = cell integrals only;
 * HD_EVAL_ALL_WITHOUT_NEIGHBOR_LOAD = face integrals with the neighbour's trace not read (taken as zero here; the reference
 * leaves it undefined); HD_EVAL_ALL = the operator (default).  Periodic / ghosted lattices only. */
#define HD_EVAL_ALL 0
#define HD_EVAL_CELL 1
#define HD_EVAL_ALL_WITHOUT_NEIGHBOR_LOAD 2
int hd_advection_set_evaluation_level(hd_advection *op, int level);
/* Pipelined 3D3V kernel: L2 residency hints, a bit mask (1: keep the direction-4 outflow layers in L2 for the downwind
 * neighbour, 2: evict-first on the far face loads, 4: streaming stores; -1 = default = environment HD_L2_HINTS, else 0).
 * A tuning knob, results do not depend on it; only effective in builds with -DHD_HINTS=1 (off by default: measured
 * no gain, and the extra code costs instruction-cache space). */
int hd_advection_set_l2_hints(hd_advection *op, int mask);
/* Pipelined 3D3V kernel: order in which the rows of cells (cells along x_0) are visited — the device counterpart of the
 * cell order of MatrixFree::loop_cell_centric (matrix_free.templates.h:1497-1581: v outer, x inner).  Rows are handed out
 * tile by tile, tile[i] rows along direction i+1 (i = 0..4; -1 = library default, 0 = the full extent, i.e. lattice
 * order), so that upwind face layers are re-read from L2 instead of HBM.  A tuning knob, results do not depend on it. */
int hd_advection_set_row_tile(hd_advection *op, const int *tile);
/* Name of the kernel the last apply launched (for logs and tests). */
const char *hd_advection_kernel_name(const hd_advection *op);
/* number of kernel launches issued by this operator so far */
int64_t hd_advection_launch_count(const hd_advection *op);
/* Inhomogeneous Dirichlet data (BoundaryDescriptor::dirichlet_bc, boundary_descriptor.h:42-76 with
 * MatrixFreeTools::evaluate_scalar_function, matrix_free/tools.h:31): g sampled by the host at the
 * face quadrature points of every boundary face of side (dir, side) for the stage time; values are
 * ordered face-cell major (lexicographic over the other directions' local cells), then the
 * n_q^(dim-1) face quadrature points, lowest direction fastest. */
int hd_advection_set_dirichlet_values(hd_advection *op, int dir, int side, const double *g_host, int64_t n_values);
/* Use a built-in analytic field as Dirichlet data, evaluated on the device at the stage time. */
int hd_advection_set_dirichlet_builtin(hd_advection *op, int fn_id);

/* ---- ghost faces (multi-GPU) ------------------------------------------------------------ */
/* VectorDataExchange::Contiguous::export_to_ghosted_array_start (matrix_free/vector_partitioner.h:1387,
 * pack loop :1443-1460): gather the nodal face layers of `src` that neighbouring bricks need into
 * the contiguous send buffer.  Segment (dir, side) of the SEND buffer holds this brick's own
 * boundary layer on that side; the matching segment of the neighbour's GHOST buffer is
 * (dir, 1-side).  Transport between GPUs (NCCL send/recv or peer copies) is done by the caller
 * between hd_halo_pack and hd_advection_apply. */
int     hd_halo_pack(hd_mesh *mesh, const void *src, void *send_buffer);
/* Selective / direct variant: send_mask[2*dir+side] (NULL = all) selects the boundary layers to pack;
 * peer_dst[2*dir+side] (NULL or entry NULL = the send buffer segment) is a device pointer the layer is written to
 * instead — e.g. the peer-mapped address of the neighbour GPU's ghost segment (dir, 1-side), in which case the pack
 * kernel's stores travel over NVLink and no separate transport step is needed (the caller synchronises the GPUs).
 * started_counter (optional, device int): every CTA adds 1 when it starts; *ctas_launched (optional) returns how many
 * CTAs this call launched, so that "the pack kernels are resident" can be awaited with hd_stream_wait_flag. */
int     hd_halo_pack_ex(hd_mesh *mesh, const void *src, void *send_buffer, const int *send_mask, void *const *peer_dst, void *started_counter,
                        int *ctas_launched);
int64_t hd_halo_offset(const hd_mesh *mesh, int dir, int side); /* offset (values) of a segment  */
int64_t hd_halo_total(const hd_mesh *mesh);                     /* total values of all segments  */

/* ---- low-storage Runge-Kutta ----------------------------------------------------------- */
/* LowStorageRungeKuttaIntegrator (base/time_integrators.h:48, coefficients
 * base/time_integrators.templates.h:34-86). type: "rk33" | "rk45" | "rk47" | "rk59". */
int hd_lsrk_create(hd_mesh *mesh, const char *type, hd_lsrk **out);
int hd_lsrk_destroy(hd_lsrk *rk);
int hd_lsrk_n_stages(const hd_lsrk *rk);
/* coefficients: which 0 = b_i [n_stages], 1 = a_i [n_stages-1] */
int hd_lsrk_coefficients(const hd_lsrk *rk, int which, double *out);
/* One stage update with an externally computed K (perform_stage, time_integrators.templates.h:103-138):
 *   solution += b*K ; if (a != 0) next_Ti = solution_old + a*K        (b, a already multiplied by dt) */
int hd_lsrk_stage_update(hd_mesh *mesh, void *solution, void *next_Ti, const void *K, double b, double a);
/* perform_time_step (time_integrators.templates.h:93-184) with the advection operator as `op`:
 * all stages on the device; uses the fused operator+update kernel when available.  Ki and Ti
 * are the two registers the reference's constructor takes; for single-GPU meshes only. */
int hd_lsrk_step(hd_lsrk *rk, hd_advection *op, void *solution, void *vec_Ki, void *vec_Ti, double t, double dt);

/* One fused LSRK stage on a brick whose ghost faces the CALLER has filled (multi-GPU time stepping: one ghost exchange of the
 * current Ti per stage, then this call): K = op(ti_cur, t + c_stage dt) is never stored,
 *   solution += b_stage dt K ;  ti_next = solution_old + a_stage dt K   (not written in the last stage).
 * solution, ti_cur and ti_next are three different vectors (the neighbours still read ti_cur: ping-pong Ti between the two
 * registers of the reference's constructor, as hd_lsrk_step does).  ghosts = ghost faces of ti_cur, NULL on an unpartitioned mesh. */
int hd_lsrk_stage_fused(hd_lsrk *rk, hd_advection *op, int stage, void *solution, const void *ti_cur, void *ti_next, const void *ghosts, double t, double dt);
/* The same with the ghost exchange INSIDE the kernel (3D3V degree-3 FP64 kernels; arguments as hd_advection_apply_overlapped):
 * sender CTAs store ti_cur's boundary layers into the neighbours' ghost buffers, interior cells run meanwhile. */
int hd_lsrk_stage_overlapped(hd_lsrk *rk, hd_advection *op, int stage, void *solution, const void *ti_cur, void *ti_next, const void *ghosts,
                             const hd_halo_send *sends, int n_sends, const void *arrival_counters, int target, double t, double dt);

/* ---- several GPUs in ONE process (the C++ host's route; one process per GPU goes through hd_halo_pack_ex / *_overlapped) ----
 * The reference builds its process grid PartitionX x PartitionV in C++ (performance/util/driver.h:133-161,
 * examples/advection/advection.cc:82-88) and exchanges ghost faces inside LinearAlgebra::SharedMPI::Vector
 * (matrix_free/vector_partitioner.h:1387-1692).  hd_multi is that for one box: the lattice of global_desc (n_cells_global,
 * left/right, degree..., side_kind = kind of the DOMAIN boundary per direction: HD_SIDE_PERIODIC_LOCAL or a Dirichlet kind;
 * n_cells / cell_offset are ignored) is cut into grid[d] equal bricks per direction, brick i (coordinates: i in mixed radix
 * over grid, direction 0 fastest) lives on devices[i] (NULL: device i), peer access is enabled between all of them, ghost
 * faces travel as direct stores over NVLink into the receiver's ghost buffer — from inside the operator kernel on 3D3V
 * degree-3 FP64 lattices (arrival counters in peer memory, as hd_advection_apply_overlapped), from a pack kernel otherwise
 * (HD_MULTI_FUSED=0 forces the latter) — and the reuse of the two ghost buffers per brick is ordered by CUDA events.
 * Vectors are arrays of n_gpus device pointers (one per brick, the brick's own lattice layout). */
typedef struct hd_multi           hd_multi;
typedef struct hd_multi_advection hd_multi_advection;
typedef struct hd_multi_lsrk      hd_multi_lsrk;
int         hd_multi_create(int n_gpus, const int *devices, const hd_mesh_desc *global_desc, const int *grid, hd_multi **out);
int         hd_multi_destroy(hd_multi *mm);
int         hd_multi_n_gpus(const hd_multi *mm);
hd_mesh *   hd_multi_mesh(hd_multi *mm, int brick);
hd_context *hd_multi_context(hd_multi *mm, int brick);
int64_t     hd_multi_n_dofs(const hd_multi *mm);
int         hd_multi_synchronize(hd_multi *mm);
int         hd_multi_vector_alloc(hd_multi *mm, void **ptrs);
int         hd_multi_vector_free(hd_multi *mm, void *const *ptrs);
/* scatter / gather between the bricks and ONE host vector in the layout of the unpartitioned lattice */
int hd_multi_vector_copy_in(hd_multi *mm, void *const *ptrs, const void *host_global);
int hd_multi_vector_copy_out(hd_multi *mm, void *const *ptrs, void *host_global);
int hd_multi_interpolate_builtin(hd_multi *mm, void *const *vec, int fn_id, double time);
/* out = the two SUMS over all bricks (as hd_norm_and_error_builtin: the caller takes the square roots) */
int hd_multi_norm_and_error_builtin(hd_multi *mm, void *const *vec, int fn_id, double time, double out[2]);
int hd_multi_advection_create(hd_multi *mm, double skew_factor, const double *velocity, hd_multi_advection **out);
int hd_multi_advection_destroy(hd_multi_advection *op);
int hd_multi_advection_set_dirichlet_builtin(hd_multi_advection *op, int fn_id);
const char *hd_multi_advection_kernel_name(const hd_multi_advection *op);
/* ghost exchange + AdvectionOperation::apply on every brick (asynchronous on the bricks' streams) */
int hd_multi_advection_apply(hd_multi_advection *op, void *const *dst, void *const *src, double time);
int hd_multi_lsrk_create(hd_multi *mm, const char *type, hd_multi_lsrk **out);
int hd_multi_lsrk_destroy(hd_multi_lsrk *rk);
/* perform_time_step on all bricks: per stage one ghost exchange of Ti and one fused operator + update launch per brick */
int hd_multi_lsrk_step(hd_multi_lsrk *rk, hd_multi_advection *op, void *const *solution, void *const *vec_Ki, void *const *vec_Ti, double t, double dt);

/* ---- VectorTools ------------------------------------------------------------------------ */
/* VectorTools::interpolate (numerics/vector_tools.h:88): nodal values at the GLL points. */
int hd_interpolate_builtin(hd_mesh *mesh, void *vec, int fn_id, double time);
/* VectorTools::norm_and_error (numerics/vector_tools.h:151): out[0] = sum u_h^2 JxW,
 * out[1] = sum (u_h - f)^2 JxW over the OWNED cells at the quadrature points (the caller
 * all-reduces over GPUs and takes the square roots, vector_tools.h:212-219). */
int hd_norm_and_error_builtin(hd_mesh *mesh, const void *vec, int fn_id, double time, double out[2]);

/* VectorTools::velocity_space_integration (numerics/vector_tools.h:238-315) with quad_no_v = 2, the call of the
 * Vlasov-Poisson driver (examples/vlasov_poisson/include/application.h:520-527): the particle density at the x-space nodes,
 *   rho[x-cell][x-node] = sum over the OWNED v-cells and their Gauss-Lobatto nodes of f * JxW_v .
 * dst_x: device pointer to hd_mesh_n_dofs_x values of the mesh's number type (x-cells lexicographic with x_0 fastest,
 * (k+1)^dim_x nodal values per cell — the layout of the x-space DoF vector, hd_vector_alloc_x); it is overwritten.
 * With a partition of v-space the caller sums dst_x over the GPUs that share the x-brick (vector_tools.h:308-314). */
int64_t hd_mesh_n_dofs_x(const hd_mesh *mesh);
int     hd_vector_alloc_x(hd_mesh *mesh, void **device_ptr);
int     hd_velocity_space_integration(hd_mesh *mesh, void *dst_x, const void *src);

/* ---- x-space field solve of the Vlasov-Poisson right-hand side ----------------------------------------------------------
 * Steps 2-4 of examples/vlasov_poisson/include/application.h:529-583 on a periodic Cartesian x-lattice:
 *   rhs = -M (rho - mean), mean removed again;  K phi = rhs with the symmetric-interior-penalty DG Laplacian of
 *   examples/vlasov_poisson/include/poisson.h:166-250, solved on the device by conjugate gradients preconditioned with the
 *   operator's point-Jacobi diagonal, from the previous potential (the reference: CG preconditioned by a Chebyshev smoother
 *   over the same diagonal, or multigrid, to a relative residual of 1e-7, poisson.h:575-613);  a_v = grad(phi) at the quadrature
 *   points of every x-cell (DerivativeContainer::update, derivative_container.h:157-190) in the layout
 *   hd_advection_set_phase_space_velocity reads: a_v_device[x-cell][x-quadrature point][dim_x] doubles.
 * rho_x: the density of hd_velocity_space_integration (mesh number type).  *iterations (optional) returns the CG steps taken.
 * If the relative residual rel_tol is not reached within max_iterations (or CG breaks down) the call returns
 * HD_ERR_NO_CONVERGENCE and leaves the gradient table untouched — the reference's SolverCG throws in that case. */
int hd_poisson_create(hd_mesh *mesh, hd_poisson **out);
int hd_poisson_destroy(hd_poisson *ps);
int hd_poisson_solve(hd_poisson *ps, const void *rho_x, double *a_v_device, double rel_tol, int max_iterations, int *iterations);
/* CG steps and achieved relative residual |r| / |b| of the last hd_poisson_solve (either pointer may be NULL) */
int hd_poisson_last_solve(const hd_poisson *ps, int *iterations, double *rel_residual);
/* potential of the last solve: device pointer to hd_mesh_n_dofs_x doubles (x-space layout), owned by the solver */
const double *hd_poisson_potential(const hd_poisson *ps);

/* Diagnostics of the Vlasov-Poisson driver (examples/vlasov_poisson/include/diagnostics.h):
 * phase_space_diagnostics (:34-86) at the Gauss points of the OWNED cells: out = {sum f JxW (mass), sum f^2 JxW (the caller
 * all-reduces and takes the square root: L2 norm), sum |v|^2 f JxW (kinetic energy), sum v_d f JxW for d < dim_v (momentum), 0..};
 * compute_electric_energy (:88-143): out[d] = sum_q (a_v[.][q][d])^2 JxW over the x-lattice, d < dim_x, from the gradient
 * table of hd_poisson_solve. */
int hd_phase_space_diagnostics(hd_mesh *mesh, const void *vec, double out[6]);
int hd_field_energy(hd_mesh *mesh, const double *a_v_device, double *out);

/* ---- timing ----------------------------------------------------------------------------- */
/* CUDA-event timing on the context's stream (the device-side counterpart of hyperdeal::Timers,
 * base/timers.h:36): returns milliseconds between the two calls. */
int hd_timer_start(hd_context *ctx);
int hd_timer_stop(hd_context *ctx, double *milliseconds);

#ifdef __cplusplus
}
#endif
#endif
