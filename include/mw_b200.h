/* include/mw_b200.h -- C ABI of libmwb200.so, the B200 (sm_100a) implementation of miniWeatherML's
 * time-stepping hot path.  Plain C: POD structs, raw pointers and sizes, integer status codes.
 *
 * The reference has no FFI for this path: its modules are header-only C++ classes the driver instantiates
 * (experiments/supercell_example/driver.cpp:51-61).  The entry points below are what the host-side C++ module
 * classes of this repo (miniweatherml_b200/host/) call, one per reference method they replace:
 *
 *   mw_dycore_create / _destroy           Dynamics_Euler_Stratified_WenoFV::init, dtor
 *                                         (model/modules/dynamics_euler_stratified_wenofv.h:1197-1683)
 *   mw_dycore_set_background              hy_dens_cells/_theta_cells/_edges/_theta_edges            (DYC:51-54,1325-1328)
 *   mw_dycore_set_immersed                DataManager entry "immersed_proportion"                   (DYC:1313)
 *   mw_dycore_compute_time_step           compute_time_step                                         (DYC:70-77)
 *   mw_dycore_time_step[_host]            time_step(coupler, dt_phys)                               (DYC:81-198)
 *   mw_dycore_init_supercell              init_supercell + convert_dynamics_to_coupler              (DYC:1687-1887,1891)
 *   mw_kessler_step[_host]                Microphysics_Kessler::time_step                  (microphysics_kessler.h:99-162)
 *   mw_surrogate_forward                  custom_modules::Microphysics_Kessler NN part (PON:177-202) + ponni
 *                                         Inference::forward_batch_parallel (external/ponni/src/ponni_Inference.h:177)
 *   mw_sponge_layer                       modules::sponge_layer                                 (sponge_layer.h:8-77)
 *   mw_column_average / mw_nudge_to_column  ColumnNudger::set_column / nudge_to_column       (column_nudging.h:15-106)
 *   mw_perturb_temperature                modules::perturb_temperature (thermal bubble)     (perturb_temperature.h:43-65)
 *   mw_comm_*                             the MPI calls of halo_exchange / sponge / nudging, on NCCL
 *
 * Every function returns MW_OK (0) or a negative mw_status; mw_last_error() gives the message of the last
 * failure on the calling thread (the C++ wrappers turn it into the reference's endrun() -> std::runtime_error).
 * Unless a name ends in _host, all data pointers are DEVICE pointers owned by the caller (the DataManager), laid
 * out exactly like the reference's arrays with nens == 1: [nz][ny][nx], x contiguous.  `stream` is a cudaStream_t
 * passed as void* (NULL = default stream); calls are asynchronous on it unless stated otherwise.
 */
#ifndef MW_B200_H
#define MW_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MW_MAX_TRACERS 50               /* model/core/MultipleFields.h:11 */

typedef enum {
  MW_OK = 0,
  MW_ERR_INVALID = -1,       /* bad argument / unsupported configuration (message says which) */
  MW_ERR_CUDA = -2,          /* a CUDA runtime/driver call failed */
  MW_ERR_NO_DEVICE = -3,     /* no sm_100 device visible: there is NO CPU fallback */
  MW_ERR_NCCL = -4
} mw_status;

enum { MW_BC_PERIODIC = 0, MW_BC_OPEN = 1, MW_BC_WALL = 2 };       /* DYC:46-48 */

/* Grid, decomposition and physics constants: the values the reference keeps in core::Coupler and its options. */
typedef struct {
  int    nx, ny, nz, nens;            /* local (per-rank) interior sizes; nens must be 1                          */
  int    nx_glob, ny_glob;            /* global sizes (CPL:110); sim2d <=> ny_glob == 1                            */
  int    i_beg, j_beg;                /* global index of local cell (0,0)  (CPL:147-153)                           */
  int    nproc_x, nproc_y, px, py;    /* rank grid (CPL:127-145)                                                   */
  double xlen, ylen, zlen;            /* domain size [m]; dx = xlen/nx_glob ... (CPL:316)                          */
  int    num_tracers;                 /* <= MW_MAX_TRACERS                                                         */
  int    idWV;                        /* tracer index of water vapour, -1 if none (DYC:1292)                       */
  int    tracer_positive [MW_MAX_TRACERS];
  int    tracer_adds_mass[MW_MAX_TRACERS];
  double R_d, R_v, cp_d, p0, grav;    /* options set by micro/dycore init (KES:85-94, DYC:1227-1232)               */
  double C0, gamma_d;                 /* DYC:1242-1247                                                             */
  double earthrot, latitude;          /* fcor = 2*earthrot*sin(latitude) (DYC:213)                                 */
  int    bc_x, bc_y, bc_z;            /* each: periodic, open or wall (DYC:46-48); periodic bc_z needs nz >= 5             */
  int    enable_gravity;
  int    use_immersed_boundaries;
} mw_config;

typedef struct mw_dycore mw_dycore;   /* opaque: owns the haloed state buffers, flux scratch, TMA descriptors    */
typedef struct mw_comm   mw_comm;     /* opaque: NCCL communicator + neighbour table + comm stream               */

const char *mw_last_error(void);
int  mw_version(void);
/* 0 if a usable sm_100 device is present (selects nothing), MW_ERR_NO_DEVICE otherwise */
int  mw_device_check(void);
/* fills the derived constants (cv_d, gamma_d, kappa_d, C0) from R_d, cp_d, p0 exactly as DYC:1240-1247 */
int  mw_config_defaults(mw_config *cfg);

/* ---- device selection and memory (what the host-side DataManager allocates with; DM:44-59,126-195,571) ------ */
int  mw_device_set(int ordinal);
int  mw_device_count(int *n);
int  mw_malloc(void **ptr, size_t bytes);
int  mw_free(void *ptr);
int  mw_memset(void *ptr, int byte, size_t bytes, void *stream);
int  mw_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);     /* asynchronous */
int  mw_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);     /* synchronises the stream */
int  mw_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream);
int  mw_fence(void);                                                             /* yakl::fence() */
/* measurement aid (no reference counterpart): DFMA thread-instructions per second this device sustains, the peak of
 * the fp64_pipe roofline in bench.py */
int  mw_probe_fp64_rate(double *dfma_thread_instr_per_s);

/* ---- dycore ------------------------------------------------------------------------------------------- */
int  mw_dycore_create(const mw_config *cfg, mw_dycore **out);
int  mw_dycore_destroy(mw_dycore *h);
int  mw_dycore_get_config(const mw_dycore *h, mw_config *out);
/* host pointers: hy_dens_cells[nz], hy_dens_theta_cells[nz], hy_dens_edges[nz+1], hy_dens_theta_edges[nz+1] */
int  mw_dycore_set_background(mw_dycore *h, const double *hy_dens_cells, const double *hy_dens_theta_cells,
                              const double *hy_dens_edges, const double *hy_dens_theta_edges);
/* read them back (host pointers), e.g. to register "hy_dens_cells" in the DataManager (DYC:1663-1668) */
int  mw_dycore_get_background(mw_dycore *h, double *hy_dens_cells, double *hy_dens_theta_cells,
                              double *hy_dens_edges, double *hy_dens_theta_edges);
/* installs (non-NULL) or removes (NULL) the immersed_proportion mask; use_immersed_boundaries follows it */
int  mw_dycore_set_immersed(mw_dycore *h, const double *immersed_proportion /* device [nz][ny][nx] or NULL */);
double mw_dycore_compute_time_step(const mw_dycore *h);
/* the options compute_tendencies re-reads on every call (DYC:211-225: enable_gravity, grav, latitude, earthrot, C0, gamma_d,
 * bc_z): forward their current coupler values before a step */
int  mw_dycore_update_options(mw_dycore *h, int enable_gravity, double grav, double latitude, double earthrot, double C0,
                              double gamma_d, int bc_z);
/* bc_x / bc_y are read at every halo and edge exchange (DYC:588-589, :846-847): periodic, open or wall.  Open / wall:
 * the halo of a domain boundary repeats the edge cell and the boundary face sees its inner state on both sides, normal
 * velocity zero at a wall (DYC:782-825, :1040-1080).  With ONE rank in a direction the reference applies the face condition
 * to the west / south face only (`else if`, DYC:1051, :1072) and so does this library; with two or more ranks both get it. */
int  mw_dycore_update_lateral_bc(mw_dycore *h, int bc_x, int bc_y);
/* fields: host array of 5+T DEVICE pointers in coupler order (density_dry,uvel,vvel,wvel,temp,tracers...) */
int  mw_dycore_time_step(mw_dycore *h, double *const *fields, double dt_phys, void *stream);
/* same through HOST buffers: H2D of the 5+T fields, the step, D2H of the results; synchronous */
int  mw_dycore_time_step_host(mw_dycore *h, double *const *host_fields, double dt_phys);
/* the schedule mw_dycore_time_step_host walks when it pipelines uploads, kernels and downloads slab by slab (pure host
 * logic, no device needed): ops6 receives n_ops rows of (level, kind, stage, r0, r1, after_upload) -- kind 0 coupler->dycore,
 * 1 stage kernel, 2 tracer finish, 3 dycore->coupler; rows [r0, r1), r1 > ny = both sides of the periodic seam; after_upload =
 * slab upload the operation waits for, -1 = seam phase.  n_ops = 0 when ny is too small for the chain (serial path). */
int  mw_host_pipeline_plan(int ny, int rows_per_slab, int num_tracers, int ncycles, int *ops6, int max_ops, int *n_ops,
                           int *n_slabs);
/* supercell initial condition (hydrostatic GLL-quadrature column + cell averages) straight into coupler fields
 * and into the handle's background profiles */
int  mw_dycore_init_supercell(mw_dycore *h, double *const *fields, void *stream);
/* the other init_data test cases of Dynamics_Euler_Stratified_WenoFV::init (DYC:1338-1653), written into the coupler
 * fields, the handle's background profiles and (building / city) the immersed_proportion mask:
 *   thermal  : rising moist bubble over a constant-theta hydrostatic column, 3-point Gauss-Legendre (DYC:1338-1419)
 *   building : uniform u = 20 m/s flow, one immersed block (DYC:1544-1651); honours enable_gravity
 *   city     : uniform u = 20 m/s flow, a grid of immersed buildings whose heights the caller draws exactly like the
 *              reference (std::mt19937{17}, normal(60,10), row-major [nby][nbx]; DYC:1431-1451, 1506-1512)
 * `immersed` is a DEVICE [nz][ny][nx] array owned by the caller (the "immersed_proportion" DataManager entry); it is
 * filled and installed as the handle's mask (use_immersed_boundaries = true, DYC:1424,1548). */
int  mw_dycore_init_thermal(mw_dycore *h, double *const *fields, void *stream);
int  mw_dycore_init_building(mw_dycore *h, double *const *fields, double *immersed, void *stream);
int  mw_city_layout(double xlen, double ylen, int nx_glob, int *cells_per_building, int *nbuildings_y,
                    int *nbuildings_x);                              /* DYC:1430-1437 */
int  mw_dycore_init_city(mw_dycore *h, double *const *fields, double *immersed, const double *building_heights_host,
                         int nbuildings_y, int nbuildings_x, void *stream);
/* attach a communicator: halos of decomposed directions then go through NCCL send/recv instead of a local wrap */
int  mw_dycore_attach_comm(mw_dycore *h, mw_comm *comm);
/* diagnostics: number of kernels this handle launched since creation */
long long mw_dycore_launch_count(const mw_dycore *h);
/* device time [ms] of the stage kernels / all kernels during the last time_step (CUDA events on the step's stream;
 * valid after the stream has been synchronised); n_stage = stage-kernel launches inside that step */
int  mw_dycore_last_timing(mw_dycore *h, float *stage_kernel_ms, int *n_stage, float *step_ms);
int  mw_dycore_enable_timing(mw_dycore *h, int on);

/* ---- WENO5 building block, exposed for the kernel-level parity tests ----------------------------------- */
/* stencils: device [n][5]; out: device [n][2] (value at the low face, value at the high face)            */
int  mw_weno5_edges(const double *stencils, double *out, long long n, void *stream);

/* ---- Kessler microphysics ------------------------------------------------------------------------------ */
/* device fields [nz][ncol]; precl device [ncol]; rainsplit_out (host int*, may be NULL) forces a sync when given.
 * comm may be NULL (single rank); otherwise the sub-cycle count is min-reduced over ranks.                      */
int  mw_kessler_step(int nz, long long ncol, double dz, double dt, double R_d, double R_v, double cp_d, double p0,
                     double *temp, const double *rho_dry, double *rho_v, double *rho_c, double *rho_r,
                     double *precl, mw_comm *comm, int *rainsplit_out, void *stream);
int  mw_kessler_step_host(int nz, long long ncol, double dz, double dt, double R_d, double R_v, double cp_d,
                          double p0, double *temp, const double *rho_dry, double *rho_v, double *rho_c,
                          double *rho_r, double *precl, int *rainsplit_out);

/* ---- ponni surrogate (5 -> 10 -> LeakyReLU(0.1) -> 4, fp32) ---------------------------------------------- */
/* weights: HOST fp32 W1[5][10], b1[10], W2[10][4], b2[4]; scl_in HOST [5][2], scl_out HOST [4][2];
 * in/out device fp64 [n].  use_tensor_cores != 0 selects the mma path (3xTF32 split), 0 the fp32 FMA path.   */
int  mw_surrogate_forward(long long n, const float *weights, const double *scl_in, const double *scl_out,
                          const double *temp, const double *rho_d, const double *rho_v, const double *rho_c,
                          const double *rho_r, double *o_temp, double *o_rho_v, double *o_rho_c, double *o_rho_r,
                          int use_tensor_cores, void *stream);
/* the bare MLP on normalised fp32 inputs x[5][B] -> y[4][B] (device), for the ponni known-answer tests */
int  mw_mlp_forward(long long B, const float *weights /*host*/, const float *x, float *y, int use_tensor_cores,
                    void *stream);

/* general Dense(nin->nh) + LeakyReLU(negative_slope) + Dense(nh->nout) in fp32 with ponni's operation order (the model family
 * of ponni's own known-answer test, external/ponni/unit/keras_sequential/test_keras_sequential.cpp:11-50); widths <= 256.
 * weights: HOST W1[nin][nh], b1[nh], W2[nh][nout], b2[nout] (Keras "kernel:0" is [in][out]); x device [nin][B], y device [nout][B] */
int  mw_mlp_dense2_forward(long long B, int nin, int nh, int nout, float negative_slope, const float *weights,
                           const float *x, float *y, void *stream);
/* the same network on the tensor cores (tcgen05.mma kind::tf32, 3xTF32 split, TMEM accumulators): nin <= 16, nh <= 256,
 * nout <= 16; agrees with the fp32 path to ponni's test tolerance (1e-6), not bit for bit.  The `use_tensor_cores` flag of
 * mw_mlp_forward / mw_surrogate_forward selects the same kernels for the 5 -> 10 -> 4 network (PON:177-202). */
int  mw_mlp_dense2_forward_tc(long long B, int nin, int nh, int nout, float negative_slope, const float *weights,
                              const float *x, float *y, void *stream);

/* ---- the other calls of the canonical step loop ---------------------------------------------------------- */
/* the surrogate module's per-step diagnostic (PON:258-269: sum(a - b) / size of nfields field pairs), reduced on the device */
int  mw_mean_difference(int nfields, const double *const *a, const double *const *b, long long n, double *mean_host,
                        void *stream);
int  mw_sponge_layer(int nfields, double *const *fields, int nz, int ny, int nx, long long nx_glob_ny_glob,
                     double dz, double zlen, double dt, double time_scale, mw_comm *comm, void *stream);
/* column[5][nz] (device) <- horizontal mean of (density_dry,uvel,vvel,temp,water_vapor) */
int  mw_column_average(const double *const *f5, int nz, int ny, int nx, long long nx_glob_ny_glob, double *column,
                       mw_comm *comm, void *stream);
int  mw_nudge_to_column(double *const *f5, int nz, int ny, int nx, long long nx_glob_ny_glob, double dt,
                        const double *column, mw_comm *comm, void *stream);
int  mw_perturb_temperature(double *temp, int nz, int ny, int nx, int i_beg, int j_beg, double dx, double dy,
                            double dz, double xlen, double ylen, void *stream);

/* ---- experiments/simple_city custom modules ------------------------------------------------------------------ */
/* Horizontal_Sponge::init (horizontal_sponge.h:18-92): column[nfields][nz] (device) <- cell (k,0,0) of each field on
 * rank 0, broadcast to every rank when comm != NULL */
int  mw_extract_column(int nfields, const double *const *fields, int nz, int ny, int nx, double *column,
                       mw_comm *comm, void *stream);
/* Horizontal_Sponge::apply (horizontal_sponge.h:103-193): relax the `sponge_cells` outermost cells of each enabled
 * side towards column[f][k] with weight (cos(pi*d/(sponge_cells-1))+1)/2 * dt/time_scale, sides applied in the
 * reference's order x1, x2, y1, y2; a side is active only on the ranks that own it (px == 0, px == nproc_x-1, ...) */
int  mw_horizontal_sponge_apply(int nfields, double *const *fields, const double *column, int nz, int ny, int nx,
                                int sponge_cells, double time_scale, double dt, int x1, int x2, int y1, int y2,
                                int px, int nproc_x, int py, int nproc_y, void *stream);
/* Time_Averager::accumulate (time_averager.h:34-66): inertia = etime/(etime+dt); avg <- inertia*avg + (1-inertia)*val
 * for nfields fields of n cells (the caller advances etime) */
int  mw_time_average_accumulate(int nfields, double *const *avg, const double *const *val, long long n, double etime,
                                double dt, void *stream);

/* ---- ensemble members ------------------------------------------------------------------------------------------ */
/* The reference's fields are [nz][ny][nx][nens] (CPL:328); the kernels here take one member [nz][ny][nx].  gather copies
 * member iens of each interleaved field into a contiguous member array, scatter writes it back (ncell = nz*ny*nx). */
int  mw_ensemble_gather(int nfields, double *const *member, const double *const *fields, long long ncell, int nens,
                        int iens, void *stream);
int  mw_ensemble_scatter(int nfields, double *const *fields, const double *const *member, long long ncell, int nens,
                         int iens, void *stream);

/* ---- communicator (NCCL over NVLink), one process per GPU -------------------------------------------------- */
int  mw_comm_unique_id(void *id_bytes_128);                       /* rank 0 creates, caller broadcasts            */
int  mw_comm_create(const void *id_bytes_128, int nranks, int rank, mw_comm **out);
int  mw_comm_barrier(mw_comm *c);                                 /* MPI_Barrier: host returns when every rank arrived */
int  mw_comm_destroy(mw_comm *c);

#ifdef __cplusplus
}
#endif
#endif
