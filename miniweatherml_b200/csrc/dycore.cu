// Host side of the dycore C ABI (include/mw_b200.h): owns the device-resident RK registers, the tracer flux
// scratch, the background profiles and the TMA descriptors, and sequences the kernels of one time_step
// (reference: model/modules/dynamics_euler_stratified_wenofv.h:81-198).
#include "dycore_kernels.cuh"
#include "stage_cell.cuh"
#include "comm.cuh"
#include <cmath>
#include <algorithm>
#include <cstring>
#include <vector>
#include <cstdlib>

namespace mw {

thread_local std::string g_last_error;
void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

int device_check_cached() {
  static int cached = 1;
  if (cached == 1) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      set_error("no CUDA device visible (%s); libmwb200 has no CPU fallback", cudaGetErrorString(e));
      cudaGetLastError();
      return MW_ERR_NO_DEVICE;
    }
    int dev = 0, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) {
      set_error("device compute capability %d.x, kernels are built for sm_100a only", major);
      return MW_ERR_NO_DEVICE;
    }
    cached = MW_OK;
  }
  return cached;
}

int encode_tensor_map_f64_4d(CUtensorMap *map, const void *base, const uint64_t dims[4],
                             const uint64_t strides_bytes[3], const uint32_t box[4]) {
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled entry point not available (%s)", cudaGetErrorString(e));
      return MW_ERR_CUDA;
    }
    fn = (encode_fn) p;
  }
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void *>(base), (const cuuint64_t *) dims,
                  (const cuuint64_t *) strides_bytes, (const cuuint32_t *) box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int) r);
    return MW_ERR_CUDA;
  }
  return MW_OK;
}

}  // namespace mw

using namespace mw;

// The stage kernel is the cell kernel of stage_cell.cuh: a 32 x 8 tile of columns per CTA, one thread per cell.
constexpr int TILE_X = 32, TILE_Y = 8;

struct mw_dycore {
  mw_config cfg;
  int N;
  double dx, dy, dz;
  int pitch;
  long long zstride, vstride;
  size_t qbytes;
  double *q[3] = {nullptr, nullptr, nullptr};
  CUtensorMap tmap[3];
  CUtensorMap tmapI[3];                // cell kernel: interior box {32, 8, 1, N} of the same buffers (z windows)
  double *flux_x = nullptr, *flux_y = nullptr, *flux_z = nullptr, *mult = nullptr;
  unsigned char *tflag = nullptr;      // [T][tile rows][tile columns]: the stage scaled a flux of that tile (StageParams::tflag)
  double *bg = nullptr;                // hyc[nz], hytc[nz], hye[nz+1], hyte[nz+1]
  std::vector<double> bg_host;
  bool bg_set = false;
  const double *immersed = nullptr;
  mw_comm *comm = nullptr;
  long long launches = 0;
  int use_tma = 1;
  bool bc_both_faces = false;          // MW_BC_BOTH_FACES=1, see base_params
  // staging for the *_host entry point
  double *dev_fields[NUM_STATE + MW_MAX_TRACERS] = {nullptr};
  bool dev_fields_alloc = false;
  // halo exchange over NCCL (decomposed directions only): 0 = W, 1 = E, 2 = S, 3 = N
  bool dir_active[4] = {false, false, false, false};
  int peer[4] = {0, 0, 0, 0};
  double *hsend[4] = {nullptr}, *hrecv[4] = {nullptr}, *msend[4] = {nullptr}, *mrecv[4] = {nullptr};
  size_t hcount[4] = {0}, mcount[4] = {0};
  // Peer-memory halos (default for decomposed runs, MW_PEER_HALO=0 selects the NCCL exchange): the neighbours' RK
  // registers, FCT-factor arrays and barrier flags mapped through CUDA IPC.  Kernels store the images of their edge cells
  // straight into the neighbour's halo over NVLink; a flag barrier between neighbours replaces every exchange.
  struct Peer {
    double *q[3] = {nullptr, nullptr, nullptr};          // the neighbour's q0, q1, q2 (peer address space)
    double *mult = nullptr;
    unsigned long long *flags = nullptr;                   // the neighbour's flag array: I write slot (d ^ 1)
    int nx = 0, ny = 0, pitch = 0;
    long long zstride = 0, vstride = 0;
  } peer_mem[4];
  bool peer_halo = false;
  unsigned long long *flags = nullptr;                     // my flag array [4]: slot d is written by the neighbour in direction d
  unsigned long long epoch = 0;
  std::vector<void *> ipc_opened;
  // halo exchange overlapped with interior compute: exchanges run on their own stream
  cudaStream_t cs = nullptr, cs2 = nullptr;                // exchange stream; second compute stream for the boundary tiles
  cudaEvent_t ev_ready = nullptr, ev_halo = nullptr, ev_prev = nullptr, ev_bnd = nullptr;
  bool overlap = false, halo_inflight = false;
  // slab-pipelined host step (mw_dycore_time_step_host): upload / compute / download streams and per-slab events
  cudaStream_t s_up = nullptr, s_cmp = nullptr, s_dn = nullptr;
  std::vector<cudaEvent_t> ev_up, ev_done;
  // timing
  bool timing = false;
  std::vector<cudaEvent_t> ev;         // [0]=step begin, [1]=step end, then pairs per stage kernel
  int n_stage_timed = 0;
  bool timing_valid = true;            // false after a slab-pipelined host step (no events recorded)
};

// Open / wall lateral boundary on side d (W, E, S, N) of THIS rank's block: the boundary-condition code, else 0
// (periodic direction, interior rank boundary, 2-D run in y).  DYC:782-825 tests px / py the same way.
static int bc_side(const mw_dycore *h, int d) {
  const mw_config &c = h->cfg;
  if (d >= 2 && c.ny_glob == 1) return 0;
  const int bc = d < 2 ? c.bc_x : c.bc_y;
  if (bc != MW_BC_OPEN && bc != MW_BC_WALL) return 0;
  const int p = d < 2 ? c.px : c.py, np = d < 2 ? c.nproc_x : c.nproc_y;
  return ((d & 1) ? p == np - 1 : p == 0) ? bc : 0;
}

static StageParams base_params(const mw_dycore *h) {
  StageParams P;
  memset(&P, 0, sizeof(P));
  const mw_config &c = h->cfg;
  P.nx = c.nx; P.ny = c.ny; P.nz = c.nz;
  P.pitch = h->pitch; P.zstride = h->zstride; P.vstride = h->vstride;
  P.flux_x = h->flux_x; P.flux_y = h->flux_y; P.flux_z = h->flux_z; P.mult = h->mult;
  P.tflag = h->tflag; P.tf_nbx = (c.nx + TILE_X - 1) / TILE_X; P.tf_nby = (c.ny + TILE_Y - 1) / TILE_Y;
  P.hyc = h->bg; P.hytc = h->bg + c.nz; P.hye = h->bg + 2 * c.nz; P.hyte = h->bg + 3 * c.nz + 1;
  P.ihytc = h->bg + 4 * c.nz + 2; P.pcell = P.ihytc + c.nz; P.ihyte = P.pcell + c.nz; P.pedge = P.ihyte + c.nz + 1;
  P.pser[0] = 1.0;                                         // C(gamma, n)
  for (int n = 1; n < 12; ++n) P.pser[n] = P.pser[n - 1] * (c.gamma_d - (n - 1)) / n;
  P.immersed = h->immersed;
  P.dx = h->dx; P.dy = h->dy; P.dz = h->dz;
  P.rdx = 1.0 / h->dx; P.rdy = 1.0 / h->dy; P.rdz = 1.0 / h->dz;
  P.C0 = c.C0; P.gamma = c.gamma_d; P.grav = c.grav;
  P.fcor = 2 * c.earthrot * sin(c.latitude);
  P.sim2d = (c.ny_glob == 1);
  P.bc_z = c.bc_z;
  P.enable_gravity = c.enable_gravity;
  P.use_immersed = (c.use_immersed_boundaries && h->immersed) ? 1 : 0;
  unsigned pm = 0;
  for (int t = 0; t < c.num_tracers && t < 32; ++t) if (c.tracer_positive[t]) pm |= 1u << t;   // (groups: see group_params)
  P.positive_mask = pm;
  P.use_tma = h->use_tma;
  P.tile_mode = 0;
  for (int d = 0; d < 4; ++d) {
    P.hbc[d] = P.fbc[d] = bc_side(h, d);
    // one rank in the direction: the reference applies the boundary condition to the low FACE only (DYC:1051, :1072
    // `else if`) and the high one keeps the periodic neighbour's outer state; MW_BC_BOTH_FACES=1 gives the
    // two-or-more-ranks behaviour on one rank (tests)
    if ((d & 1) && P.fbc[d] && (d < 2 ? c.nproc_x : c.nproc_y) == 1 && !h->bc_both_faces) P.fbc[d] = MW_FBC_REF1;
    if (P.hbc[d]) P.bc_any = 1;
  }
  if (c.bc_z == MW_BC_PERIODIC) P.bc_any = 1;            // the periodic z boundary also lives in the LBC instantiation of the stage kernel
  // FCT donors across interior rank boundaries (never across the global periodic seam, see StageParams)
  const bool interior[4] = {h->dir_active[0] && c.px > 0, h->dir_active[1] && c.px < c.nproc_x - 1,
                            h->dir_active[2] && c.py > 0, h->dir_active[3] && c.py < c.nproc_y - 1};
  for (int d = 0; d < 4; ++d) {
    StageParams::MultSrc &M = P.msrc[d];
    if (!interior[d]) continue;
    if (h->peer_halo) {                                    // the neighbour's own factor array [T][nz][ny'][nx']
      const mw_dycore::Peer &R = h->peer_mem[d];
      M.base = R.mult;
      M.st_t = (long long) c.nz * R.ny * R.nx; M.st_k = (long long) R.ny * R.nx;
      if (d < 2) { M.st_i = R.nx; M.off = (d == 0) ? R.nx - 1 : 0; }                         // W: its last column, E: its first
      else       { M.st_i = 1;    M.off = (d == 2) ? (long long) (R.ny - 1) * R.nx : 0; }    // S: its last row,    N: its first
    } else {                                               // packed strip received over NCCL: [T][nz][ny] or [T][nz][nx]
      const long long len = (d < 2) ? c.ny : c.nx;
      M.base = h->mrecv[d]; M.st_t = (long long) c.nz * len; M.st_k = len; M.st_i = 1; M.off = 0;
    }
  }
  return P;
}

// Destinations of the edge cells' images for a launch that writes RK register `buf` (StageParams::img): this rank's own
// halo in a direction that is not decomposed (periodic wrap), the neighbour's halo in peer mode, nothing in NCCL mode.
static void set_images(const mw_dycore *h, StageParams &P, int buf) {
  const mw_config &c = h->cfg;
  const bool sim2d = (c.ny_glob == 1);
  for (int d = 0; d < 4; ++d) {
    StageParams::ImgDst &D = P.img[d];
    D.base = nullptr;
    P.idelta[d] = 0;
    if (d == 0) P.ifast = 0;
    const bool decomposed = h->dir_active[d];
    if (d >= 2 && sim2d) continue;
    if (bc_side(h, d)) continue;                           // a domain boundary: the halo holds boundary copies, not images
    int nxr = c.nx, nyr = c.ny;                            // the receiver's block
    if (!decomposed) { D.base = h->q[buf]; D.vstride = h->vstride; D.zstride = h->zstride; D.pitch = h->pitch; }
    else if (h->peer_halo) {
      const mw_dycore::Peer &R = h->peer_mem[d];
      D.base = R.q[buf]; D.vstride = R.vstride; D.zstride = R.zstride; D.pitch = R.pitch; nxr = R.nx; nyr = R.ny;
    } else continue;
    // my cell (j, i) lands at row0 + j, col0 + i of the receiver's haloed array
    D.row0 = HALO; D.col0 = HALO;
    if (d == 0) D.col0 = nxr + HALO;                       // i in [0,3)        -> its east halo columns nx' + 3 + i
    if (d == 1) D.col0 = HALO - c.nx;                      // i in [nx-3, nx)   -> its west halo columns i - nx + 3
    if (d == 2) D.row0 = nyr + HALO;
    if (d == 3) D.row0 = HALO - c.ny;
    if (D.vstride == h->vstride && D.zstride == h->zstride && D.pitch == h->pitch) {
      P.idelta[d] = (D.base - h->q[buf]) + (long long) (D.row0 - HALO) * D.pitch + (D.col0 - HALO);
      P.ifast |= 1 << d;
    }
  }
}

extern "C" const char *mw_last_error(void) { return g_last_error.c_str(); }
extern "C" int mw_version(void) { return 100; }
extern "C" int mw_device_check(void) { return device_check_cached(); }

extern "C" int mw_config_defaults(mw_config *cfg) {
  MW_REQUIRE(cfg, "mw_config_defaults: null config");
  if (cfg->R_d == 0) cfg->R_d = 287.;
  if (cfg->cp_d == 0) cfg->cp_d = 1003.;
  if (cfg->R_v == 0) cfg->R_v = 461.;
  if (cfg->p0 == 0) cfg->p0 = 1.e5;
  if (cfg->grav == 0) cfg->grav = 9.81;
  if (cfg->earthrot == 0) cfg->earthrot = 7.292115e-5;
  const double cv_d = cfg->cp_d - cfg->R_d;                                 // DYC:1240-1247
  cfg->gamma_d = cfg->cp_d / cv_d;
  const double kappa_d = cfg->R_d / cfg->cp_d;
  cfg->C0 = pow(cfg->R_d * pow(cfg->p0, -kappa_d), cfg->gamma_d);
  return MW_OK;
}

extern "C" int mw_dycore_create(const mw_config *cfg, mw_dycore **out) {
  MW_REQUIRE(cfg && out, "mw_dycore_create: null argument");
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_REQUIRE(cfg->nens == 1, "nens = %d: only nens == 1 is implemented", cfg->nens);
  MW_REQUIRE(cfg->nx >= HALO && cfg->nz >= 2 && cfg->ny >= 1, "grid too small (nx=%d ny=%d nz=%d)", cfg->nx, cfg->ny, cfg->nz);
  MW_REQUIRE(cfg->ny_glob == 1 || cfg->ny >= HALO, "ny = %d too small for a 3-D run", cfg->ny);
  MW_REQUIRE(cfg->num_tracers >= 0 && cfg->num_tracers <= MW_MAX_TRACERS, "num_tracers = %d out of range", cfg->num_tracers);
  for (int bc : {cfg->bc_x, cfg->bc_y})
    MW_REQUIRE(bc == MW_BC_PERIODIC || bc == MW_BC_OPEN || bc == MW_BC_WALL, "bc_x / bc_y must be periodic, open or wall");
  MW_REQUIRE(cfg->bc_z == MW_BC_WALL || cfg->bc_z == MW_BC_OPEN || cfg->bc_z == MW_BC_PERIODIC, "bc_z must be periodic, open or wall");
  MW_REQUIRE(cfg->bc_z != MW_BC_PERIODIC || cfg->nz >= 5, "periodic bc_z needs nz >= 5 (nz = %d)", cfg->nz);
  MW_REQUIRE(cfg->C0 > 0 && cfg->gamma_d > 1, "C0/gamma_d not set (call mw_config_defaults)");
  MW_REQUIRE((double) (cfg->nz + 1) * (cfg->ny + 1) * (cfg->nx + 1) < 2147483647.0, "block of %d x %d x %d cells: the face arrays are indexed with 32 bits",
             cfg->nx, cfg->ny, cfg->nz);
  mw_dycore *h = new mw_dycore();
  h->cfg = *cfg;
  h->N = NUM_STATE + cfg->num_tracers;
  h->dx = cfg->xlen / cfg->nx_glob; h->dy = cfg->ylen / cfg->ny_glob; h->dz = cfg->zlen / cfg->nz;
  h->pitch = ((cfg->nx + 2 * HALO + 1) / 2) * 2;
  h->zstride = (long long) (cfg->ny + 2 * HALO) * h->pitch;
  h->vstride = (long long) cfg->nz * h->zstride;
  h->qbytes = (size_t) h->N * h->vstride * sizeof(double);
  const char *e = getenv("MW_NO_TMA");
  h->use_tma = (e && atoi(e) != 0) ? 0 : 1;
  if (cfg->num_tracers > 4) h->use_tma = 0;                // tracer groups (step_groups) run the plain-load instantiation
  e = getenv("MW_BC_BOTH_FACES");
  h->bc_both_faces = e && atoi(e) != 0;
  const int T = cfg->num_tracers > 0 ? cfg->num_tracers : 1;
  const size_t nzl = cfg->nz, nyl = cfg->ny, nxl = cfg->nx;
  cudaError_t ce = cudaSuccess;
  for (int b = 0; b < 3 && ce == cudaSuccess; ++b) {
    ce = cudaMalloc(&h->q[b], h->qbytes);
    if (ce == cudaSuccess) ce = cudaMemset(h->q[b], 0, h->qbytes);
  }
  if (ce == cudaSuccess) ce = cudaMalloc(&h->flux_x, T * nzl * nyl * (nxl + 1) * 8);
  if (ce == cudaSuccess) ce = cudaMalloc(&h->flux_y, T * nzl * (nyl + 1) * nxl * 8);
  if (ce == cudaSuccess) ce = cudaMalloc(&h->flux_z, T * (nzl + 1) * nyl * nxl * 8);
  if (ce == cudaSuccess) ce = cudaMalloc(&h->mult, T * nzl * nyl * nxl * 8);
  const size_t nflag = (size_t) T * ((nyl + TILE_Y - 1) / TILE_Y) * ((nxl + TILE_X - 1) / TILE_X);
  if (ce == cudaSuccess) ce = cudaMalloc(&h->tflag, nflag);
  if (ce == cudaSuccess) ce = cudaMemset(h->tflag, 1, nflag);
  if (ce == cudaSuccess) ce = cudaMalloc(&h->bg, (8 * nzl + 4) * 8);
  if (ce != cudaSuccess) {
    set_error("mw_dycore_create: device allocation failed: %s", cudaGetErrorString(ce));
    mw_dycore_destroy(h);
    return MW_ERR_CUDA;
  }
  for (int b = 0; b < 3 && cfg->num_tracers <= 4; ++b) {
    const uint64_t dims[4] = {(uint64_t) h->pitch, (uint64_t) (cfg->ny + 2 * HALO), (uint64_t) cfg->nz, (uint64_t) h->N};
    const uint64_t str[3] = {(uint64_t) h->pitch * 8, (uint64_t) h->zstride * 8, (uint64_t) h->vstride * 8};
    const uint32_t box[4] = {TILE_X + 2 * HALO, TILE_Y + 2 * HALO, 1, (uint32_t) h->N};
    rc = encode_tensor_map_f64_4d(&h->tmap[b], h->q[b], dims, str, box);
    if (rc != MW_OK) { mw_dycore_destroy(h); return rc; }
    const uint32_t boxi[4] = {TILE_X + 2, TILE_Y, 1, (uint32_t) h->N};   // CellCfg::IW x TY
    rc = encode_tensor_map_f64_4d(&h->tmapI[b], h->q[b], dims, str, boxi);
    if (rc != MW_OK) { mw_dycore_destroy(h); return rc; }
  }
  *out = h;
  return MW_OK;
}

extern "C" int mw_dycore_destroy(mw_dycore *h) {
  if (!h) return MW_OK;
  for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
  cudaFree(h->flags);
  for (int b = 0; b < 3; ++b) cudaFree(h->q[b]);
  cudaFree(h->flux_x); cudaFree(h->flux_y); cudaFree(h->flux_z); cudaFree(h->mult); cudaFree(h->tflag); cudaFree(h->bg);
  if (h->dev_fields_alloc) for (int f = 0; f < h->N; ++f) cudaFree(h->dev_fields[f]);
  for (int d = 0; d < 4; ++d) { cudaFree(h->hsend[d]); cudaFree(h->hrecv[d]); cudaFree(h->msend[d]); cudaFree(h->mrecv[d]); }
  for (auto e : h->ev) cudaEventDestroy(e);
  for (cudaStream_t s : {h->s_up, h->s_cmp, h->s_dn}) if (s) cudaStreamDestroy(s);
  for (auto e : h->ev_up) cudaEventDestroy(e);
  for (auto e : h->ev_done) cudaEventDestroy(e);
  if (h->cs) cudaStreamDestroy(h->cs);
  if (h->cs2) cudaStreamDestroy(h->cs2);
  for (cudaEvent_t e : {h->ev_ready, h->ev_halo, h->ev_prev, h->ev_bnd}) if (e) cudaEventDestroy(e);
  delete h;
  return MW_OK;
}

extern "C" int mw_dycore_set_background(mw_dycore *h, const double *hyc, const double *hytc, const double *hye,
                                        const double *hyte) {
  MW_REQUIRE(h && hyc && hytc && hye && hyte, "mw_dycore_set_background: null argument");
  const int nz = h->cfg.nz;
  h->bg_host.resize(4 * nz + 2);
  memcpy(&h->bg_host[0], hyc, nz * 8);
  memcpy(&h->bg_host[nz], hytc, nz * 8);
  memcpy(&h->bg_host[2 * nz], hye, (nz + 1) * 8);
  memcpy(&h->bg_host[3 * nz + 1], hyte, (nz + 1) * 8);
  // derived tables for the in-kernel equation of state: 1/bg and C0*bg^gamma at cell centres and z edges
  std::vector<double> all(8 * nz + 4);
  memcpy(all.data(), h->bg_host.data(), (4 * nz + 2) * 8);
  double *ihytc = &all[4 * nz + 2], *pcell = ihytc + nz, *ihyte = pcell + nz, *pedge = ihyte + nz + 1;
  for (int k = 0; k < nz; ++k) { ihytc[k] = 1.0 / hytc[k]; pcell[k] = h->cfg.C0 * pow(hytc[k], h->cfg.gamma_d); }
  for (int k = 0; k <= nz; ++k) { ihyte[k] = 1.0 / hyte[k]; pedge[k] = h->cfg.C0 * pow(hyte[k], h->cfg.gamma_d); }
  MW_CUDA_OK(cudaMemcpy(h->bg, all.data(), all.size() * 8, cudaMemcpyHostToDevice));
  h->bg_set = true;
  return MW_OK;
}

extern "C" int mw_dycore_get_background(mw_dycore *h, double *hyc, double *hytc, double *hye, double *hyte) {
  MW_REQUIRE(h && h->bg_set, "mw_dycore_get_background: background not set");
  const int nz = h->cfg.nz;
  if (hyc) memcpy(hyc, &h->bg_host[0], nz * 8);
  if (hytc) memcpy(hytc, &h->bg_host[nz], nz * 8);
  if (hye) memcpy(hye, &h->bg_host[2 * nz], (nz + 1) * 8);
  if (hyte) memcpy(hyte, &h->bg_host[3 * nz + 1], (nz + 1) * 8);
  return MW_OK;
}

extern "C" int mw_dycore_set_immersed(mw_dycore *h, const double *immersed) {
  MW_REQUIRE(h, "mw_dycore_set_immersed: null handle");
  h->immersed = immersed;
  h->cfg.use_immersed_boundaries = immersed ? 1 : 0;      // the "use_immersed_boundaries" option follows the mask (DYC:1424,1548)
  return MW_OK;
}

// The reference re-reads these coupler options at every compute_tendencies call (DYC:211-225); the host module forwards
// their current values before each step so that a set_option() after init() is honoured.
extern "C" int mw_dycore_update_options(mw_dycore *h, int enable_gravity, double grav, double latitude, double earthrot,
                                        double C0, double gamma_d, int bc_z) {
  MW_REQUIRE(h, "mw_dycore_update_options: null handle");
  MW_REQUIRE(bc_z == MW_BC_WALL || bc_z == MW_BC_OPEN || bc_z == MW_BC_PERIODIC, "bc_z must be periodic, open or wall");
  MW_REQUIRE(bc_z != MW_BC_PERIODIC || h->cfg.nz >= 5, "periodic bc_z needs nz >= 5 (nz = %d)", h->cfg.nz);
  MW_REQUIRE(C0 > 0 && gamma_d > 1, "mw_dycore_update_options: C0 = %g, gamma_d = %g", C0, gamma_d);
  mw_config &c = h->cfg;
  const bool eos_changed = (C0 != c.C0 || gamma_d != c.gamma_d);
  c.enable_gravity = enable_gravity ? 1 : 0; c.grav = grav; c.latitude = latitude; c.earthrot = earthrot;
  c.C0 = C0; c.gamma_d = gamma_d; c.bc_z = bc_z;
  if (eos_changed && h->bg_set) {                          // the pressure tables depend on C0 and gamma
    const int nz = c.nz;
    const std::vector<double> bg = h->bg_host;
    return mw_dycore_set_background(h, &bg[0], &bg[nz], &bg[2 * nz], &bg[3 * nz + 1]);
  }
  return MW_OK;
}

extern "C" int mw_dycore_update_lateral_bc(mw_dycore *h, int bc_x, int bc_y) {
  MW_REQUIRE(h, "mw_dycore_update_lateral_bc: null handle");
  for (int bc : {bc_x, bc_y})
    MW_REQUIRE(bc == MW_BC_PERIODIC || bc == MW_BC_OPEN || bc == MW_BC_WALL, "bc_x / bc_y must be periodic, open or wall");
  h->cfg.bc_x = bc_x; h->cfg.bc_y = bc_y;
  return MW_OK;
}

extern "C" double mw_dycore_compute_time_step(const mw_dycore *h) {      // DYC:70-77
  if (!h) return 0.0;
  const double maxwave = 350 + 80, cfl = 0.6;
  return cfl * fmin(fmin(h->dx, h->dy), h->dz) / maxwave;
}

extern "C" int mw_dycore_get_config(const mw_dycore *h, mw_config *out) {
  MW_REQUIRE(h && out, "mw_dycore_get_config: null argument");
  *out = h->cfg;
  return MW_OK;
}

extern "C" long long mw_dycore_launch_count(const mw_dycore *h) { return h ? h->launches : 0; }

extern "C" int mw_dycore_enable_timing(mw_dycore *h, int on) {
  MW_REQUIRE(h, "null handle");
  h->timing = on != 0;
  return MW_OK;
}

extern "C" int mw_dycore_last_timing(mw_dycore *h, float *stage_ms, int *n_stage, float *step_ms) {
  MW_REQUIRE(h && h->timing && h->ev.size() >= 2, "timing not enabled or no step taken");
  MW_REQUIRE(h->timing_valid, "mw_dycore_last_timing: the last step was a slab-pipelined host step, which records no stage events");
  float s = 0, t = 0;
  MW_CUDA_OK(cudaEventElapsedTime(&t, h->ev[0], h->ev[1]));
  for (int i = 0; i < h->n_stage_timed; ++i) {
    float d = 0;
    MW_CUDA_OK(cudaEventElapsedTime(&d, h->ev[2 + 2 * i], h->ev[3 + 2 * i]));
    s += d;
  }
  if (stage_ms) *stage_ms = s;
  if (n_stage) *n_stage = h->n_stage_timed;
  if (step_ms) *step_ms = t;
  return MW_OK;
}

// ---- halo exchange (replaces halo_exchange DYC:574-747 + edge_exchange DYC:830-1003 with ONE width-3 exchange) ----
// Sends go out in the order W,E,S,N and receives are posted in the order E,W,N,S, so that with two ranks in a
// direction (both neighbours are the same peer) the first message sent is the first one received.
static int exchange(mw_dycore *h, double *const *send, double *const *recv, const size_t *count, cudaStream_t st) {
  static const int send_order[4] = {0, 1, 2, 3}, recv_order[4] = {1, 0, 3, 2};
  MW_NCCL_OK(ncclGroupStart());
  for (int n = 0; n < 4; ++n) {
    const int dr = recv_order[n], ds = send_order[n];
    if (h->dir_active[dr]) MW_NCCL_OK(ncclRecv(recv[dr], count[dr], ncclDouble, h->peer[dr], h->comm->comm, st));
    if (h->dir_active[ds]) MW_NCCL_OK(ncclSend(send[ds], count[ds], ncclDouble, h->peer[ds], h->comm->comm, st));
  }
  MW_NCCL_OK(ncclGroupEnd());
  return MW_OK;
}

// Neighbour barrier of the peer-memory mode: thread d tells the neighbour in direction d "everything I launched before this
// on my stream is done and visible" (its stores into your halo included) and waits for the same word from it.  A lost
// neighbour must fail loudly, not hang the GPU: trap after timeout_ns of polling (MW_PEER_TIMEOUT_S, default 300 s -- long
// enough for a neighbour that writes output or initialises while this rank already waits in its first barrier).
__global__ void k_peer_barrier(unsigned long long *mine, unsigned long long *p0, unsigned long long *p1, unsigned long long *p2,
                               unsigned long long *p3, unsigned long long epoch, unsigned long long timeout_ns) {
  unsigned long long *peer[4] = {p0, p1, p2, p3};
  const int d = threadIdx.x;
  if (d >= 4 || !peer[d]) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer[d] + (d ^ 1)), "l"(epoch) : "memory");
  unsigned long long v = 0, t0 = 0;
  for (long long spin = 0;; ++spin) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + d) : "memory");
    if (v >= epoch) break;
    if ((spin & 1023) == 1023) {                           // look at the clock now and then
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > timeout_ns) __trap();
    }
    __nanosleep(100);
  }
}
static int peer_barrier(mw_dycore *h, cudaStream_t st) {
  ++h->epoch;
  unsigned long long *pf[4];
  for (int d = 0; d < 4; ++d) pf[d] = h->dir_active[d] ? h->peer_mem[d].flags : nullptr;
  static const unsigned long long timeout_ns = [] {
    const char *e = getenv("MW_PEER_TIMEOUT_S");
    const double sec = (e && atof(e) > 0) ? atof(e) : 300.0;
    return (unsigned long long) (sec * 1e9);
  }();
  k_peer_barrier<<<1, 32, 0, st>>>(h->flags, pf[0], pf[1], pf[2], pf[3], h->epoch, timeout_ns);
  MW_CUDA_OK(cudaGetLastError());
  h->launches++;
  return MW_OK;
}

static int exchange_halos(mw_dycore *h, double *q, cudaStream_t st) {
  if (!h->dir_active[0] && !h->dir_active[2]) return MW_OK;
  if (h->peer_halo) return peer_barrier(h, st);            // the producing kernels stored the halos themselves
  const mw_config &c = h->cfg;
  HaloParams H;
  H.nx = c.nx; H.ny = c.ny; H.nz = c.nz; H.nvar = h->N; H.pitch = h->pitch; H.zstride = h->zstride; H.vstride = h->vstride;
  H.q = q;
  if (h->dir_active[0]) {
    H.buf[0] = h->hsend[0]; H.buf[1] = h->hsend[1];
    k_halo_x<true><<<dim3((unsigned) ((h->hcount[0] + 255) / 256), 2), 256, 0, st>>>(H);
    h->launches++;
  }
  if (h->dir_active[2]) {
    H.buf[0] = h->hsend[2]; H.buf[1] = h->hsend[3];
    k_halo_y<true><<<dim3((unsigned) ((h->hcount[2] + 255) / 256), 2), 256, 0, st>>>(H);
    h->launches++;
  }
  MW_CUDA_OK(cudaGetLastError());
  int rc = exchange(h, h->hsend, h->hrecv, h->hcount, st);
  if (rc != MW_OK) return rc;
  // a domain boundary's halo keeps its boundary copies: what the periodic neighbour sent is dropped
  if (h->dir_active[0]) {
    H.buf[0] = bc_side(h, 0) ? nullptr : h->hrecv[0]; H.buf[1] = bc_side(h, 1) ? nullptr : h->hrecv[1];
    k_halo_x<false><<<dim3((unsigned) ((h->hcount[0] + 255) / 256), 2), 256, 0, st>>>(H);
    h->launches++;
  }
  if (h->dir_active[2]) {
    H.buf[0] = bc_side(h, 2) ? nullptr : h->hrecv[2]; H.buf[1] = bc_side(h, 3) ? nullptr : h->hrecv[3];
    k_halo_y<false><<<dim3((unsigned) ((h->hcount[2] + 255) / 256), 2), 256, 0, st>>>(H);
    h->launches++;
  }
  MW_CUDA_OK(cudaGetLastError());
  return MW_OK;
}

// boundary cells' FCT factors to the neighbouring ranks, between the stage kernel and the tracer finish
static int exchange_mult(mw_dycore *h, cudaStream_t st) {
  if ((!h->dir_active[0] && !h->dir_active[2]) || h->cfg.num_tracers == 0) return MW_OK;
  if (h->peer_halo) return peer_barrier(h, st);            // the tracer finish reads the neighbours' factor arrays directly
  const mw_config &c = h->cfg;
  MultEdgeParams M;
  M.nx = c.nx; M.ny = c.ny; M.nz = c.nz; M.nt = c.num_tracers; M.mult = h->mult;
  if (h->dir_active[0]) {
    M.out[0] = h->msend[0]; M.out[1] = h->msend[1];
    k_mult_edge_x<<<dim3((unsigned) ((h->mcount[0] + 255) / 256), 2), 256, 0, st>>>(M);
    h->launches++;
  }
  if (h->dir_active[2]) {
    M.out[0] = h->msend[2]; M.out[1] = h->msend[3];
    k_mult_edge_y<<<dim3((unsigned) ((h->mcount[2] + 255) / 256), 2), 256, 0, st>>>(M);
    h->launches++;
  }
  MW_CUDA_OK(cudaGetLastError());
  return exchange(h, h->msend, h->mrecv, h->mcount, st);
}

// tracer finish (FCT-scaled divergence, RK, clip); with decomposed directions the neighbours' FCT factors come first
static int exchange_halos(mw_dycore *h, double *q, cudaStream_t st);
// `Qd2c` (last stage of a step): the tracer finish also writes the coupler fields (k_tracer_update_d2c)
static int finish_stage_local(mw_dycore *h, const StageParams &P, int nt, cudaStream_t st, bool mult_done = false,
                              const ConvertParams *Qd2c = nullptr) {
  if (nt > 0) {
    if (!mult_done) {
      int rc = exchange_mult(h, st);
      if (rc != MW_OK) return rc;
    }
    const long long ncell = (long long) P.nz * P.ny * P.nx;
    const unsigned g = (unsigned) ((ncell + 255) / 256);
    if (Qd2c) {
      switch (nt) {
        case 1: k_tracer_update_d2c<1><<<g, 256, 0, st>>>(P, *Qd2c); break;
        case 2: k_tracer_update_d2c<2><<<g, 256, 0, st>>>(P, *Qd2c); break;
        case 3: k_tracer_update_d2c<3><<<g, 256, 0, st>>>(P, *Qd2c); break;
        case 4: k_tracer_update_d2c<4><<<g, 256, 0, st>>>(P, *Qd2c); break;
      }
    } else {
      switch (nt) {
        case 1: k_tracer_update<1><<<g, 256, 0, st>>>(P); break;
        case 2: k_tracer_update<2><<<g, 256, 0, st>>>(P); break;
        case 3: k_tracer_update<3><<<g, 256, 0, st>>>(P); break;
        case 4: k_tracer_update<4><<<g, 256, 0, st>>>(P); break;
      }
    }
    MW_CUDA_OK(cudaGetLastError());
    h->launches++;
  }
  return MW_OK;
}
// ... followed by the halo exchange of the new state on the same stream (no overlap)
static int finish_stage(mw_dycore *h, const StageParams &P, int nt, cudaStream_t st, const ConvertParams *Qd2c = nullptr) {
  int rc = finish_stage_local(h, P, nt, st, false, Qd2c);
  if (rc != MW_OK) return rc;
  return exchange_halos(h, P.qout, st);
}
// start the width-3 halo exchange of buffer q on the exchange stream once everything queued on `st` has finished
static int exchange_halos_async(mw_dycore *h, double *q, cudaStream_t st) {
  MW_CUDA_OK(cudaEventRecord(h->ev_ready, st));
  MW_CUDA_OK(cudaStreamWaitEvent(h->cs, h->ev_ready, 0));
  int rc = exchange_halos(h, q, h->cs);
  if (rc != MW_OK) return rc;
  MW_CUDA_OK(cudaEventRecord(h->ev_halo, h->cs));
  h->halo_inflight = true;
  return MW_OK;
}
// The cell kernel (stage_cell.cuh) with TMA box loads, or -- MW_NO_TMA=1, tests -- the same kernel filling its planes with plain loads
template <int NT> struct StageKernel {
  using C = CellCfg<NT>;
  static constexpr int TX = TILE_X;
  static void launch(dim3 g, cudaStream_t st, const mw_dycore *h, int b, const StageParams &P0) {
    StageParams P = P0;
    {                                                      // DYC:536-542, once per launch instead of per cell and level
      const double dtI = P.dt_stage, tau = 1.e3 * P.dt_stage;
      P.imm_c = -fmin(1.0, dtI / tau) / dtI;
    }
    if (P.bc_any) {                                        // open / wall lateral boundaries: the instantiation that knows them
      if (h->use_tma) k_stage_cell<NT, true, true><<<g, C::NTHR, C::SMEM, st>>>(h->tmap[b], h->tmapI[b], P);
      else k_stage_cell<NT, false, true><<<g, C::NTHR, C::SMEM, st>>>(h->tmap[b], h->tmapI[b], P);
    } else {
      if (h->use_tma) k_stage_cell<NT, true, false><<<g, C::NTHR, C::SMEM, st>>>(h->tmap[b], h->tmapI[b], P);
      else k_stage_cell<NT, false, false><<<g, C::NTHR, C::SMEM, st>>>(h->tmap[b], h->tmapI[b], P);
    }
  }
  static cudaError_t attr() {
    cudaError_t e = cudaFuncSetAttribute(k_stage_cell<NT, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) C::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_stage_cell<NT, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) C::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_stage_cell<NT, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) C::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_stage_cell<NT, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) C::SMEM);
    return e;
  }
};
template <int NT>
static int launch_stage(mw_dycore *h, const StageParams &P0, int in_buf, cudaStream_t st, bool last,
                        const ConvertParams *Qd2c = nullptr) {
  using K = StageKernel<NT>;
  static bool attr_set = false;
  if (!attr_set) {
    MW_CUDA_OK(K::attr());
    attr_set = true;
  }
  StageParams P = P0;
  constexpr int TX = K::TX;
  const int nbx = (P.nx + TX - 1) / TX, nby = (P.ny + 7) / 8;
  if (h->timing) {
    while (h->ev.size() < (size_t) (4 + 2 * h->n_stage_timed)) { cudaEvent_t e; cudaEventCreate(&e); h->ev.push_back(e); }
    cudaEventRecord(h->ev[2 + 2 * h->n_stage_timed], st);
  }
  if (!h->overlap) {
    K::launch(dim3(nbx, nby), st, h, in_buf, P);
    h->launches++;
  } else {
    // Tiles whose stencil reaches a halo filled by a neighbour rank wait for the exchange; the others ("interior")
    // run while it is in flight.  Both sets are launched as concurrent kernels (second compute stream) so that the
    // small boundary launch fills SMs next to the interior launch instead of adding a wave of its own.
    P.nbx = nbx; P.nby = nby;
    P.tbx_lo = h->dir_active[0] ? 1 : 0; P.tbx_hi = h->dir_active[0] ? std::max((P.nx - HALO) / TX, 0) : nbx;
    P.tby_lo = h->dir_active[2] ? 1 : 0; P.tby_hi = h->dir_active[2] ? std::max((P.ny - HALO) / 8, 0) : nby;
    int n_int = (P.tbx_hi > P.tbx_lo && P.tby_hi > P.tby_lo) ? (P.tbx_hi - P.tbx_lo) * (P.tby_hi - P.tby_lo) : 0;
    if (n_int == 0) { P.tbx_lo = P.tbx_hi = 0; P.tby_lo = nby; P.tby_hi = nby; }   // everything is boundary ("low" rows)
    MW_CUDA_OK(cudaEventRecord(h->ev_prev, st));
    if (n_int > 0) {
      P.tile_mode = 1;
      K::launch(dim3(n_int), st, h, in_buf, P);
      h->launches++;
    }
    MW_CUDA_OK(cudaStreamWaitEvent(h->cs2, h->ev_prev, 0));
    if (h->halo_inflight) { MW_CUDA_OK(cudaStreamWaitEvent(h->cs2, h->ev_halo, 0)); h->halo_inflight = false; }
    P.tile_mode = 2;
    K::launch(dim3(nbx * nby - n_int), h->cs2, h, in_buf, P);
    h->launches++;
    // the boundary cells' FCT factors go to the neighbours as soon as the boundary tiles are done (also overlapped)
    if (NT > 0) { int rc = exchange_mult(h, h->cs2); if (rc != MW_OK) return rc; }
    MW_CUDA_OK(cudaEventRecord(h->ev_bnd, h->cs2));
    MW_CUDA_OK(cudaStreamWaitEvent(st, h->ev_bnd, 0));
  }
  MW_CUDA_OK(cudaGetLastError());
  if (h->timing) { cudaEventRecord(h->ev[3 + 2 * h->n_stage_timed], st); h->n_stage_timed++; }
  if (!h->overlap) return finish_stage(h, P0, NT, st, Qd2c);
  int rc = finish_stage_local(h, P0, NT, st, true, Qd2c);
  if (rc != MW_OK) return rc;
  return last ? MW_OK : exchange_halos_async(h, P0.qout, st);      // the last stage's halos are never read
}
template <int NT>
static int step_impl(mw_dycore *h, double *const *fields, double dt_phys, cudaStream_t st) {
  const mw_config &c = h->cfg;
  const long long ncell = (long long) c.nz * c.ny * c.nx;
  ConvertParams Q;
  memset(&Q, 0, sizeof(Q));
  Q.S = base_params(h);
  for (int f = 0; f < h->N; ++f) Q.fields[f] = fields[f];
  Q.R_d = c.R_d; Q.R_v = c.R_v; Q.idWV = c.idWV;
  unsigned am = 0;
  for (int t = 0; t < c.num_tracers; ++t) if (c.tracer_adds_mass[t]) am |= 1u << t;
  Q.adds_mass_mask = am;

  auto ensure_events = [&](size_t n) {
    while (h->ev.size() < n) { cudaEvent_t e; cudaEventCreate(&e); h->ev.push_back(e); }
  };
  h->n_stage_timed = 0;
  h->timing_valid = true;
  if (h->timing) { ensure_events(2); cudaEventRecord(h->ev[0], st); }

  Q.S.qout = h->q[0];
  set_images(h, Q.S, 0);
  const unsigned cvgrid = (unsigned) ((ncell + 256 * CONV_CPT - 1) / (256 * CONV_CPT));      // CONV_CPT cells per thread
  k_coupler_to_dyn<NT><<<cvgrid, 256, 0, st>>>(Q);
  MW_CUDA_OK(cudaGetLastError());
  h->launches++;
  {
    int rc = h->overlap ? exchange_halos_async(h, h->q[0], st) : exchange_halos(h, h->q[0], st);
    if (rc != MW_OK) return rc;
  }

  double dt_dyn = mw_dycore_compute_time_step(h);
  const int ncycles = (int) ceil(dt_phys / dt_dyn);                        // DYC:104-108
  dt_dyn = dt_phys / ncycles;
  for (int ic = 0; ic < ncycles; ++ic) {
    for (int s = 0; s < 3; ++s) {
      StageParams P = base_params(h);
      int in_buf;
      if (s == 0)      { in_buf = 0; P.qout = h->q[1]; P.rk_a = 0.0;     P.rk_b = 1.0;     P.rk_cdt = dt_dyn;               P.dt_stage = dt_dyn; }
      else if (s == 1) { in_buf = 1; P.qout = h->q[2]; P.rk_a = 3. / 4.; P.rk_b = 1. / 4.; P.rk_cdt = (1. / 4.) * dt_dyn;   P.dt_stage = (1. / 4.) * dt_dyn; }
      else             { in_buf = 2; P.qout = h->q[0]; P.rk_a = 1. / 3.; P.rk_b = 2. / 3.; P.rk_cdt = (2. / 3.) * dt_dyn;   P.dt_stage = (2. / 3.) * dt_dyn; }
      P.qin = h->q[in_buf];
      P.q0 = h->q[0];
      set_images(h, P, (in_buf + 1) % 3);
      const bool last = (ic == ncycles - 1 && s == 2);
      // with tracers the last tracer finish also converts back to the coupler's variables (one launch and one pass less)
      int rc = launch_stage<NT>(h, P, in_buf, st, last, (last && NT > 0) ? &Q : nullptr);
      if (rc != MW_OK) return rc;
    }
  }
  if (NT == 0) {
    Q.S.qin = h->q[0];
    k_dyn_to_coupler<NT><<<cvgrid, 256, 0, st>>>(Q);
    MW_CUDA_OK(cudaGetLastError());
    h->launches++;
  }
  if (h->timing) cudaEventRecord(h->ev[1], st);
  return MW_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// More than four tracers (MultipleFields.h:11 allows 50).  The kernels are instantiated for 0..4 tracers -- the cell
// kernel keeps every variable of a cell in registers -- so a stage is launched once per GROUP of up to four tracers: each
// launch recomputes the state's reconstructions and fluxes (the tracers are advected by the state's mass flux) and
// handles its own tracers; only the last one stores the state.  Plain-load instantiation (the TMA boxes cover contiguous
// variables), run-time-T conversion kernels, no slab pipeline.  Correctness first: a step with T tracers costs about
// ceil(T/4) steps with four.
// ---------------------------------------------------------------------------------------------------------------------
static StageParams group_params(const mw_dycore *h, const StageParams &P, int tr0, int nt, bool last) {
  const mw_config &c = h->cfg;
  StageParams G = P;
  const long long nz = c.nz, ny = c.ny, nx = c.nx;
  G.tr0 = tr0; G.skip_state = last ? 0 : 1;
  G.bc_any = 1;                                            // selects the instantiation that knows tr0 / skip_state
  G.flux_x += (long long) tr0 * nz * ny * (nx + 1);
  G.flux_y += (long long) tr0 * nz * (ny + 1) * nx;
  G.flux_z += (long long) tr0 * (nz + 1) * ny * nx;
  G.mult += (long long) tr0 * nz * ny * nx;
  G.tflag += (long long) tr0 * G.tf_nby * G.tf_nbx;
  unsigned pm = 0;
  for (int t = 0; t < nt; ++t) if (c.tracer_positive[tr0 + t]) pm |= 1u << t;
  G.positive_mask = pm;
  for (int d = 0; d < 4; ++d) if (G.msrc[d].base) G.msrc[d].base += (long long) tr0 * G.msrc[d].st_t;
  return G;
}
template <int NT> static int launch_group_stage(mw_dycore *h, const StageParams &G, int in_buf, cudaStream_t st) {
  using K = StageKernel<NT>;
  static bool attr_set = false;
  if (!attr_set) { MW_CUDA_OK(K::attr()); attr_set = true; }
  K::launch(dim3((G.nx + K::TX - 1) / K::TX, (G.ny + 7) / 8), st, h, in_buf, G);
  MW_CUDA_OK(cudaGetLastError());
  h->launches++;
  return MW_OK;
}
static int step_groups(mw_dycore *h, double *const *fields, double dt_phys, cudaStream_t st) {
  const mw_config &c = h->cfg;
  const int T = c.num_tracers, ngroups = (T + 3) / 4;
  const long long ncell = (long long) c.nz * c.ny * c.nx;
  const unsigned cgrid = (unsigned) ((ncell + 255) / 256);
  MW_REQUIRE(!h->overlap, "more than four tracers: MW_OVERLAP is not supported");
  ConvertParams Q;
  memset(&Q, 0, sizeof(Q));
  Q.S = base_params(h);
  for (int f = 0; f < h->N; ++f) Q.fields[f] = fields[f];
  Q.R_d = c.R_d; Q.R_v = c.R_v; Q.idWV = c.idWV;
  for (int t = 0; t < T; ++t) if (c.tracer_adds_mass[t]) Q.adds_mass_mask64 |= 1ull << t;
  h->n_stage_timed = 0;
  h->timing_valid = true;
  if (h->timing) { while (h->ev.size() < 2) { cudaEvent_t e; cudaEventCreate(&e); h->ev.push_back(e); } cudaEventRecord(h->ev[0], st); }
  Q.S.qout = h->q[0];
  set_images(h, Q.S, 0);
  k_coupler_to_dyn_rt<<<cgrid, 256, 0, st>>>(Q, T);
  MW_CUDA_OK(cudaGetLastError());
  h->launches++;
  int rc = exchange_halos(h, h->q[0], st);
  if (rc != MW_OK) return rc;
  double dt_dyn = mw_dycore_compute_time_step(h);
  const int ncycles = (int) ceil(dt_phys / dt_dyn);                        // DYC:104-108
  dt_dyn = dt_phys / ncycles;
  for (int ic = 0; ic < ncycles; ++ic) {
    for (int s = 0; s < 3; ++s) {
      StageParams P = base_params(h);
      int in_buf;
      if (s == 0)      { in_buf = 0; P.qout = h->q[1]; P.rk_a = 0.0;     P.rk_b = 1.0;     P.rk_cdt = dt_dyn;               P.dt_stage = dt_dyn; }
      else if (s == 1) { in_buf = 1; P.qout = h->q[2]; P.rk_a = 3. / 4.; P.rk_b = 1. / 4.; P.rk_cdt = (1. / 4.) * dt_dyn;   P.dt_stage = (1. / 4.) * dt_dyn; }
      else             { in_buf = 2; P.qout = h->q[0]; P.rk_a = 1. / 3.; P.rk_b = 2. / 3.; P.rk_cdt = (2. / 3.) * dt_dyn;   P.dt_stage = (2. / 3.) * dt_dyn; }
      P.qin = h->q[in_buf];
      P.q0 = h->q[0];
      set_images(h, P, (in_buf + 1) % 3);
      for (int g = 0; g < ngroups; ++g) {
        const int tr0 = 4 * g, nt = std::min(4, T - tr0);
        const StageParams G = group_params(h, P, tr0, nt, g == ngroups - 1);
        switch (nt) {
          case 1: rc = launch_group_stage<1>(h, G, in_buf, st); break;
          case 2: rc = launch_group_stage<2>(h, G, in_buf, st); break;
          case 3: rc = launch_group_stage<3>(h, G, in_buf, st); break;
          default: rc = launch_group_stage<4>(h, G, in_buf, st); break;
        }
        if (rc != MW_OK) return rc;
      }
      rc = exchange_mult(h, st);                           // every tracer's boundary factors (NCCL) / the neighbour barrier (peer memory)
      if (rc != MW_OK) return rc;
      for (int g = 0; g < ngroups; ++g) {
        const int tr0 = 4 * g, nt = std::min(4, T - tr0);
        const StageParams G = group_params(h, P, tr0, nt, true);
        switch (nt) {
          case 1: k_tracer_update<1><<<cgrid, 256, 0, st>>>(G); break;
          case 2: k_tracer_update<2><<<cgrid, 256, 0, st>>>(G); break;
          case 3: k_tracer_update<3><<<cgrid, 256, 0, st>>>(G); break;
          default: k_tracer_update<4><<<cgrid, 256, 0, st>>>(G); break;
        }
        MW_CUDA_OK(cudaGetLastError());
        h->launches++;
      }
      rc = exchange_halos(h, P.qout, st);
      if (rc != MW_OK) return rc;
    }
  }
  Q.S.qin = h->q[0];
  k_dyn_to_coupler_rt<<<cgrid, 256, 0, st>>>(Q, T);
  MW_CUDA_OK(cudaGetLastError());
  h->launches++;
  if (h->timing) cudaEventRecord(h->ev[1], st);
  return MW_OK;
}

extern "C" int mw_dycore_time_step(mw_dycore *h, double *const *fields, double dt_phys, void *stream) {
  MW_REQUIRE(h && fields, "mw_dycore_time_step: null argument");
  MW_REQUIRE(h->bg_set, "mw_dycore_time_step: background profiles not set");
  MW_REQUIRE(dt_phys > 0, "mw_dycore_time_step: dt_phys = %g", dt_phys);
  MW_REQUIRE(h->cfg.nproc_x * h->cfg.nproc_y == 1 || h->comm, "decomposed run without an attached communicator");
  cudaStream_t st = (cudaStream_t) stream;
  switch (h->cfg.num_tracers) {
    case 0: return step_impl<0>(h, fields, dt_phys, st);
    case 1: return step_impl<1>(h, fields, dt_phys, st);
    case 2: return step_impl<2>(h, fields, dt_phys, st);
    case 3: return step_impl<3>(h, fields, dt_phys, st);
    case 4: return step_impl<4>(h, fields, dt_phys, st);
  }
  return step_groups(h, fields, dt_phys, st);              // 5 .. MW_MAX_TRACERS tracers
}

// ---------------------------------------------------------------------------------------------------------------------
// Host-buffer step, slab-pipelined.  The 5+T fields are uploaded in slabs of whole tile rows along y
// (cudaMemcpy2DAsync: nz chunks of rows per field).  The step is a chain of L operations per row --
// coupler->dycore conversion, then per RK stage the stage kernel and the tracer finish, then dycore->coupler -- and
// operation l on a row needs operation l-1 on the rows within 3 cells of it (stencil halo; donor-cell FCT factor).
// So operation l trails operation l-1 by that reach (stage kernels, which work on whole 8-row tile rows, by the next
// multiple of 8): after each slab upload, level l advances its frontier to (uploaded rows) - off[l], off = 0, 8, 9, 16,
// 17, 24, 25, 25 with tracers.  Compute trails the upload by < 1 slab, the download (the other PCIe direction) trails
// compute, and H2D, kernels and D2H overlap instead of adding up.  The rows next to the periodic seam (level l: off[l]
// rows at the bottom, about as many at the top) need both ends of the domain and are
// finished level by level after the last upload.  All kernels go to ONE compute stream in an order in which every
// operation follows the operations it reads from (that also covers the in-place reuse of q[0], the flux / FCT scratch
// and the periodic images); events tie that stream to the two copy streams.  Results are bit-identical to
// mw_dycore_time_step on device-resident fields (tests/test_gpu_dycore.py::test_host_step_pipelined_bit_identical).
// ---------------------------------------------------------------------------------------------------------------------
// The schedule of the pipelined host step as plain data (pure host logic, exported as mw_host_pipeline_plan so that the CPU
// test suite can check coverage and dependency order without a GPU).  kind: 0 coupler->dycore, 1 stage kernel, 2 tracer
// finish, 3 dycore->coupler.  An operation covers rows [r0, r1); r1 > ny means the two sides of the seam, [r0, ny) and
// [0, r1 - ny).  `after_upload` = index of the slab upload it has to wait for, -1 for the seam phase (after all uploads).
namespace {
struct PlanLink { int kind, stage; };
struct PlanOp { int level, kind, stage, r0, r1, after_upload; };
}
static bool host_pipeline_plan(int ny, int rows_per_slab, int num_tracers, int ncycles, std::vector<PlanLink> &chain,
                               std::vector<int> &off, std::vector<int> &cap, std::vector<PlanOp> &ops, int &S) {
  chain.clear(); ops.clear();
  S = std::max(1, ny / rows_per_slab);                                    // the last slab takes the remainder
  chain.push_back({0, 0});
  for (int ic = 0; ic < ncycles; ++ic)
    for (int st = 0; st < 3; ++st) {
      chain.push_back({1, st});
      if (num_tracers > 0) chain.push_back({2, st});
    }
  chain.push_back({3, 0});
  const int L = (int) chain.size();
  // Row offsets of the levels.  off[l]: how far level l trails the upload (and where it starts above the seam);
  // cap[l]: where its sweep ends below the seam.  A stage kernel needs its input 3 rows further (stencil halo) and works
  // on whole tile rows (multiples of 8); the tracer finish needs the stage's FCT factors 1 row further; the conversions
  // are cell-local.  With tracers: off = 0, 8, 9, 16, 17, 24, 25, 25.
  off.assign(L, 0); cap.assign(L, ny);
  for (int l = 1; l < L; ++l) {
    if (chain[l].kind == 1)      { off[l] = ((off[l - 1] + HALO + 7) / 8) * 8; cap[l] = ((cap[l - 1] - HALO) / 8) * 8; }
    else if (chain[l].kind == 2) { off[l] = off[l - 1] + 1;                     cap[l] = cap[l - 1] - 1; }
    else                         { off[l] = off[l - 1];                         cap[l] = cap[l - 1]; }
  }
  if (cap[L - 1] - off[L - 1] < 8) return false;                           // too few rows for this chain
  std::vector<int> hi(off);                                                // level l starts at row off[l]
  for (int s = 0; s < S; ++s) {
    const bool last = (s == S - 1);
    const int u = last ? ny : (s + 1) * rows_per_slab;                     // rows uploaded so far
    for (int l = 0; l < L; ++l) {
      const int nh = last ? cap[l] : std::min(u - off[l], cap[l]);
      if (nh > hi[l]) { ops.push_back({l, chain[l].kind, chain[l].stage, hi[l], nh, s}); hi[l] = nh; }
    }
  }
  for (int l = 1; l < L; ++l) {                                            // the seam region, level by level
    if (chain[l].kind == 1) ops.push_back({l, 1, chain[l].stage, cap[l], ny + off[l], -1});   // stage kernel: both sides in one launch
    else {
      if (ny > cap[l]) ops.push_back({l, chain[l].kind, chain[l].stage, cap[l], ny, -1});
      if (off[l] > 0) ops.push_back({l, chain[l].kind, chain[l].stage, 0, off[l], -1});
    }
  }
  return true;
}

extern "C" int mw_host_pipeline_plan(int ny, int rows_per_slab, int num_tracers, int ncycles, int *ops6, int max_ops,
                                     int *n_ops, int *n_slabs) {
  MW_REQUIRE(ny > 0 && rows_per_slab >= 8 && rows_per_slab % 8 == 0 && num_tracers >= 0 && ncycles >= 1 && n_ops, "mw_host_pipeline_plan: bad argument");
  std::vector<PlanLink> chain; std::vector<int> off, cap; std::vector<PlanOp> ops; int S = 0;
  if (!host_pipeline_plan(ny, rows_per_slab, num_tracers, ncycles, chain, off, cap, ops, S)) { *n_ops = 0; if (n_slabs) *n_slabs = S; return MW_OK; }
  *n_ops = (int) ops.size();
  if (n_slabs) *n_slabs = S;
  for (int i = 0; i < (int) ops.size() && i < max_ops && ops6; ++i) {
    const PlanOp &o = ops[i];
    const int v[6] = {o.level, o.kind, o.stage, o.r0, o.r1, o.after_upload};
    for (int q = 0; q < 6; ++q) ops6[6 * i + q] = v[q];
  }
  return MW_OK;
}

template <int NT>
static int host_step_pipelined(mw_dycore *h, double *const *host_fields, double dt_phys, int rows_per_slab, int ncycles) {
  using K = StageKernel<NT>;
  const mw_config &c = h->cfg;
  static bool attr_set = false;
  if (!attr_set) { MW_CUDA_OK(K::attr()); attr_set = true; }
  const int nbx = (c.nx + K::TX - 1) / K::TX, nby = (c.ny + 7) / 8;
  // ---- the chain of operations and their schedule (host_pipeline_plan above) ----
  double dt_dyn = dt_phys / ncycles;                                       // DYC:104-108
  std::vector<PlanLink> chain;
  std::vector<int> off, cap;
  std::vector<PlanOp> plan;
  int S = 0;
  if (!host_pipeline_plan(c.ny, rows_per_slab, NT, ncycles, chain, off, cap, plan, S)) return 1;   // too few rows: caller falls back
  const int L = (int) chain.size();

  if (!h->s_up) {
    MW_CUDA_OK(cudaStreamCreateWithFlags(&h->s_up, cudaStreamNonBlocking));
    MW_CUDA_OK(cudaStreamCreateWithFlags(&h->s_cmp, cudaStreamNonBlocking));
    MW_CUDA_OK(cudaStreamCreateWithFlags(&h->s_dn, cudaStreamNonBlocking));
  }
  static const bool prof = getenv("MW_HOST_PROF") && atoi(getenv("MW_HOST_PROF")) != 0;   // timeline of the slabs to stderr
  while ((int) h->ev_up.size() < S + 3) {
    cudaEvent_t a, b;
    MW_CUDA_OK(cudaEventCreateWithFlags(&a, prof ? cudaEventDefault : cudaEventDisableTiming));
    MW_CUDA_OK(cudaEventCreateWithFlags(&b, prof ? cudaEventDefault : cudaEventDisableTiming));
    h->ev_up.push_back(a); h->ev_done.push_back(b);
  }
  if (prof) MW_CUDA_OK(cudaEventRecord(h->ev_up[S + 2], h->s_up));                          // t = 0
  const size_t row_bytes = (size_t) c.nx * 8, plane_bytes = (size_t) c.ny * row_bytes;
  auto copy_rows = [&](int j0, int j1, bool up) -> cudaError_t {                            // rows [j0, j1) of every field
    for (int f = 0; f < h->N; ++f) {
      double *d = h->dev_fields[f] + (size_t) j0 * c.nx, *hp = host_fields[f] + (size_t) j0 * c.nx;
      cudaError_t e = up ? cudaMemcpy2DAsync(d, plane_bytes, hp, plane_bytes, (size_t) (j1 - j0) * row_bytes, c.nz, cudaMemcpyHostToDevice, h->s_up)
                         : cudaMemcpy2DAsync(hp, plane_bytes, d, plane_bytes, (size_t) (j1 - j0) * row_bytes, c.nz, cudaMemcpyDeviceToHost, h->s_dn);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  };
  for (int s = 0; s < S; ++s) {
    const int j0 = s * rows_per_slab, j1 = (s == S - 1) ? c.ny : j0 + rows_per_slab;
    MW_CUDA_OK(copy_rows(j0, j1, true));
    MW_CUDA_OK(cudaEventRecord(h->ev_up[s], h->s_up));
  }

  ConvertParams Q;
  memset(&Q, 0, sizeof(Q));
  Q.S = base_params(h);
  for (int f = 0; f < h->N; ++f) Q.fields[f] = h->dev_fields[f];
  Q.R_d = c.R_d; Q.R_v = c.R_v; Q.idWV = c.idWV;
  unsigned am = 0;
  for (int t = 0; t < c.num_tracers; ++t) if (c.tracer_adds_mass[t]) am |= 1u << t;
  Q.adds_mass_mask = am;

  cudaStream_t cs = h->s_cmp;
  int n_done_ev = 0;
  const bool multi = h->comm && (h->dir_active[0] || h->dir_active[2]);
  bool seam = false;                                                       // set for the operations of the seam phase
  // operation `l` of the chain on rows [r0, r1); r1 > ny means the two sides of the seam, rows [r0, ny) and [0, r1 - ny)
  auto run = [&](int l, int r0, int r1) -> int {
    if (r1 <= r0) return MW_OK;
    const int j0 = r0, j1 = std::min(r1, c.ny), nj = j1 - j0;
    const int ta = r0 / 8, tb = r1 > c.ny ? nby + (r1 - c.ny) / 8 : (r1 + 7) / 8;           // stage kernels: whole tile rows
    const unsigned cgrid = (unsigned) (((long long) c.nz * nj * c.nx + 255) / 256);
    const unsigned cvgrid = (unsigned) (((long long) c.nz * nj * c.nx + 256 * CONV_CPT - 1) / (256 * CONV_CPT));
    const PlanLink &lk = chain[l];
    if (lk.kind == 0) {
      ConvertParams q = Q;
      q.S.qout = h->q[0]; q.S.jr_lo = j0; q.S.jr_n = nj;
      set_images(h, q.S, 0);
      k_coupler_to_dyn<NT><<<cvgrid, 256, 0, cs>>>(q);
    } else if (lk.kind == 3) {
      ConvertParams q = Q;
      q.S.qin = h->q[0]; q.S.jr_lo = j0; q.S.jr_n = nj;
      k_dyn_to_coupler<NT><<<cvgrid, 256, 0, cs>>>(q);
      if (n_done_ev >= (int) h->ev_done.size()) {
        cudaEvent_t e;
        MW_CUDA_OK(cudaEventCreateWithFlags(&e, prof ? cudaEventDefault : cudaEventDisableTiming));
        h->ev_done.push_back(e);
      }
      cudaEvent_t e = h->ev_done[n_done_ev++];
      MW_CUDA_OK(cudaEventRecord(e, cs));
      MW_CUDA_OK(cudaStreamWaitEvent(h->s_dn, e, 0));
      MW_CUDA_OK(copy_rows(j0, j1, false));
    } else {
      StageParams P = Q.S;
      int in_buf;
      const int s = lk.stage;
      if (s == 0)      { in_buf = 0; P.qout = h->q[1]; P.rk_a = 0.0;     P.rk_b = 1.0;     P.rk_cdt = dt_dyn;               P.dt_stage = dt_dyn; }
      else if (s == 1) { in_buf = 1; P.qout = h->q[2]; P.rk_a = 3. / 4.; P.rk_b = 1. / 4.; P.rk_cdt = (1. / 4.) * dt_dyn;   P.dt_stage = (1. / 4.) * dt_dyn; }
      else             { in_buf = 2; P.qout = h->q[0]; P.rk_a = 1. / 3.; P.rk_b = 2. / 3.; P.rk_cdt = (2. / 3.) * dt_dyn;   P.dt_stage = (2. / 3.) * dt_dyn; }
      P.qin = h->q[in_buf];
      P.q0 = h->q[0];
      set_images(h, P, (in_buf + 1) % 3);
      if (lk.kind == 1 && tb > nby) {
        // seam: tile rows [ta, nby) and [0, tb - nby) = every tile outside the rectangle of rows [tb - nby, ta) (tile_mode 2)
        P.tile_mode = 2; P.nbx = nbx; P.nby = nby; P.tbx_lo = 0; P.tbx_hi = nbx; P.tby_lo = tb - nby; P.tby_hi = ta;
        K::launch(dim3(nbx * (tb - ta)), cs, h, in_buf, P);
      } else if (lk.kind == 1) {
        P.tile_mode = 1; P.nbx = nbx; P.nby = nby; P.tbx_lo = 0; P.tbx_hi = nbx; P.tby_lo = ta; P.tby_hi = tb;
        K::launch(dim3(nbx * (tb - ta)), cs, h, in_buf, P);
      } else {
        P.jr_lo = j0; P.jr_n = nj;
        k_tracer_update<NT><<<cgrid, 256, 0, cs>>>(P);
      }
    }
    h->launches++;
    // Decomposed directions: what the NEXT operation reads across a rank boundary goes to the neighbours right after this
    // one -- the width-3 halos of the operation's output (after c2d, after the tracer finish, after a stage without
    // tracers) and the boundary cells' FCT factors (after a stage with tracers).  The existing whole-strip exchanges are
    // reused: rows this level has not reached yet travel too and are simply exchanged again when they are final.  Every
    // rank issues the same sequence (equal local sizes are required), so the NCCL calls pair up.
    if (multi && (seam || h->dir_active[0])) {
      int rc = MW_OK;
      if (lk.kind == 0) rc = exchange_halos(h, h->q[0], cs);
      else if (lk.kind == 1 && NT > 0) rc = exchange_mult(h, cs);
      else if (chain[l + 1].kind != 3)                        // the last output before d2c is read by nobody else
        rc = exchange_halos(h, lk.stage == 0 ? h->q[1] : (lk.stage == 1 ? h->q[2] : h->q[0]), cs);
      if (rc != MW_OK) return rc;
    }
    return MW_OK;
  };

  int waited = -1;
  for (const PlanOp &op : plan) {
    if (op.after_upload >= 0) {
      while (waited < op.after_upload) { ++waited; MW_CUDA_OK(cudaStreamWaitEvent(cs, h->ev_up[waited], 0)); }
    } else if (!seam) {
      while (waited < S - 1) { ++waited; MW_CUDA_OK(cudaStreamWaitEvent(cs, h->ev_up[waited], 0)); }
      seam = true;
      // the rows next to a y rank boundary read the neighbour's rows instead of the periodic images: halos of the converted
      // state first (the main phase only exchanged them when x is decomposed), then level by level inside run()
      if (multi && !h->dir_active[0]) { int rc = exchange_halos(h, h->q[0], cs); if (rc != MW_OK) return rc; }
    }
    int rc = run(op.level, op.r0, op.r1);
    if (rc != MW_OK) return rc;
  }
  MW_CUDA_OK(cudaGetLastError());
  if (prof) MW_CUDA_OK(cudaEventRecord(h->ev_up[S + 1], h->s_dn));
  MW_CUDA_OK(cudaStreamSynchronize(h->s_dn));
  MW_CUDA_OK(cudaStreamSynchronize(cs));
  MW_CUDA_OK(cudaStreamSynchronize(h->s_up));
  if (prof) {
    float t = 0;
    fprintf(stderr, "[mw host pipeline] %d slabs of %d rows, chain of %d operations; ms since the first upload was queued\n  uploaded:", S, rows_per_slab, L);
    for (int p2 = 0; p2 < S; ++p2) { cudaEventElapsedTime(&t, h->ev_up[S + 2], h->ev_up[p2]); fprintf(stderr, " %.1f", t); }
    fprintf(stderr, "\n  converted back (download queued behind):");
    for (int s2 = 0; s2 < n_done_ev; ++s2) { cudaEventElapsedTime(&t, h->ev_up[S + 2], h->ev_done[s2]); fprintf(stderr, " %.1f", t); }
    cudaEventElapsedTime(&t, h->ev_up[S + 2], h->ev_up[S + 1]);
    fprintf(stderr, "\n  last download done: %.1f\n", t);
  }
  return MW_OK;
}

extern "C" int mw_dycore_time_step_host(mw_dycore *h, double *const *host_fields, double dt_phys) {
  MW_REQUIRE(h && host_fields, "mw_dycore_time_step_host: null argument");
  MW_REQUIRE(h->bg_set, "mw_dycore_time_step_host: background profiles not set");
  MW_REQUIRE(dt_phys > 0, "mw_dycore_time_step_host: dt_phys = %g", dt_phys);
  const mw_config &c = h->cfg;
  const size_t bytes = (size_t) c.nz * c.ny * c.nx * 8;
  if (!h->dev_fields_alloc) {
    for (int f = 0; f < h->N; ++f) {
      const cudaError_t e = cudaMalloc(&h->dev_fields[f], bytes);
      if (e != cudaSuccess) {                              // all or nothing: a partial set must not survive the failed call
        for (int g = 0; g < f; ++g) { cudaFree(h->dev_fields[g]); h->dev_fields[g] = nullptr; }
        h->dev_fields[f] = nullptr;
        set_error("mw_dycore_time_step_host: device staging buffers (%zu bytes each): %s", bytes, cudaGetErrorString(e));
        return MW_ERR_CUDA;
      }
    }
    h->dev_fields_alloc = true;
  }
  // slab pipeline: 3-D, equal blocks, and enough rows for >= 4 slabs of whole tile rows with most of a wave of tiles
  // (148 SMs, one CTA each) per slab launch.  MW_HOST_SLAB_ROWS overrides (0 = off).
  int rows_per_slab = 0;
  {
    const int nbx = (c.nx + TILE_X - 1) / TILE_X;
    int tile_rows = std::max(1, (148 * 4 / 5) / nbx);      // ~0.8 wave of CTAs per slab launch (r02i sweep: 56 rows at nx = 512)
    rows_per_slab = 8 * tile_rows;
    const char *e = getenv("MW_HOST_SLAB_ROWS");
    if (e) rows_per_slab = (atoi(e) / 8) * 8;
  }
  const int ncycles = (int) ceil(dt_phys / mw_dycore_compute_time_step(h));                 // DYC:104-108
  // decomposed runs: every rank must walk the same schedule, so the blocks must be equal
  const bool equal_blocks = c.nproc_x * c.nproc_y == 1 ||
                            (h->comm && c.nx_glob % c.nproc_x == 0 && c.ny_glob % c.nproc_y == 0 && c.nx == c.nx_glob / c.nproc_x && c.ny == c.ny_glob / c.nproc_y);
  const bool periodic_xy = c.bc_x == MW_BC_PERIODIC && c.bc_y == MW_BC_PERIODIC;     // the slab schedule wraps around the seam
  const bool pipelined = rows_per_slab >= 8 && equal_blocks && periodic_xy && c.num_tracers <= 4 && c.ny_glob > 1 && c.ny / rows_per_slab >= 4;
  if (pipelined) {
    h->timing_valid = false;                               // the slab pipeline records no stage events: mw_dycore_last_timing says so
    MW_CUDA_OK(cudaDeviceSynchronize());                  // the non-blocking streams do not order against earlier default-stream work
    int rc = 1;
    switch (c.num_tracers) {
      case 0: rc = host_step_pipelined<0>(h, host_fields, dt_phys, rows_per_slab, ncycles); break;
      case 1: rc = host_step_pipelined<1>(h, host_fields, dt_phys, rows_per_slab, ncycles); break;
      case 2: rc = host_step_pipelined<2>(h, host_fields, dt_phys, rows_per_slab, ncycles); break;
      case 3: rc = host_step_pipelined<3>(h, host_fields, dt_phys, rows_per_slab, ncycles); break;
      case 4: rc = host_step_pipelined<4>(h, host_fields, dt_phys, rows_per_slab, ncycles); break;
    }
    if (rc != 1) return rc;                               // 1 = grid too small for the chain: unpipelined path below
  }
  for (int f = 0; f < h->N; ++f) MW_CUDA_OK(cudaMemcpyAsync(h->dev_fields[f], host_fields[f], bytes, cudaMemcpyHostToDevice, 0));
  int rc = mw_dycore_time_step(h, h->dev_fields, dt_phys, nullptr);
  if (rc != MW_OK) return rc;
  for (int f = 0; f < h->N; ++f) MW_CUDA_OK(cudaMemcpyAsync(host_fields[f], h->dev_fields[f], bytes, cudaMemcpyDeviceToHost, 0));
  MW_CUDA_OK(cudaStreamSynchronize(0));
  return MW_OK;
}

// Map the neighbours' buffers (CUDA IPC; one process per GPU on one node).  Every rank publishes the handles of its three RK
// registers, its FCT-factor array and its flag array plus its block geometry through one ncclAllGather; each rank opens
// the handles of its (up to four, possibly coinciding) neighbours.
namespace {
struct PeerInfo {
  cudaIpcMemHandle_t q[3], mult, flags;
  int nx, ny, pitch, pad;
  long long zstride, vstride;
};
}
// Returns MW_OK with h->peer_halo set, or MW_OK with it unset when some rank could not map a neighbour (GPUs without peer
// access, IPC unavailable in the container, ...): the decision is collective -- every rank takes part in both NCCL calls
// whatever happened locally, and all of them fall back to the NCCL exchange together.
static int setup_peer_halos(mw_dycore *h) {
  const mw_config &c = h->cfg;
  const int nranks = c.nproc_x * c.nproc_y;
  bool ok = true;
  std::string why;
  auto fail = [&](const char *what, cudaError_t e) { if (ok) { ok = false; why = std::string(what) + ": " + cudaGetErrorString(e); } cudaGetLastError(); };
  PeerInfo mine;
  memset(&mine, 0, sizeof(mine));
  {
    const char *t = getenv("MW_PEER_TEST_FAIL_RANK");      // tests: pretend this rank cannot map its neighbours
    if (t && atoi(t) == h->comm->rank) { ok = false; why = "MW_PEER_TEST_FAIL_RANK"; }
  }
  cudaError_t e = cudaMalloc(&h->flags, 4 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(h->flags, 0, 4 * sizeof(unsigned long long));
  if (e != cudaSuccess) fail("flag array", e);
  for (int b = 0; b < 3 && ok; ++b) if ((e = cudaIpcGetMemHandle(&mine.q[b], h->q[b])) != cudaSuccess) fail("cudaIpcGetMemHandle", e);
  if (ok && (e = cudaIpcGetMemHandle(&mine.mult, h->mult)) != cudaSuccess) fail("cudaIpcGetMemHandle", e);
  if (ok && (e = cudaIpcGetMemHandle(&mine.flags, h->flags)) != cudaSuccess) fail("cudaIpcGetMemHandle", e);
  mine.nx = c.nx; mine.ny = c.ny; mine.pitch = h->pitch; mine.zstride = h->zstride; mine.vstride = h->vstride;
  mine.pad = ok ? 1 : 0;                                   // "my handles are valid"
  char *dsend = nullptr, *drecv = nullptr;
  int *dflag = nullptr;
  MW_CUDA_OK(cudaMalloc(&dsend, sizeof(PeerInfo)));
  MW_CUDA_OK(cudaMalloc(&drecv, sizeof(PeerInfo) * nranks));
  MW_CUDA_OK(cudaMalloc(&dflag, sizeof(int)));
  MW_CUDA_OK(cudaMemcpy(dsend, &mine, sizeof(PeerInfo), cudaMemcpyHostToDevice));
  MW_NCCL_OK(ncclAllGather(dsend, drecv, sizeof(PeerInfo), ncclChar, h->comm->comm, 0));
  std::vector<PeerInfo> all(nranks);
  MW_CUDA_OK(cudaMemcpy(all.data(), drecv, sizeof(PeerInfo) * nranks, cudaMemcpyDeviceToHost));
  for (int d = 0; d < 4 && ok; ++d) {
    if (!h->dir_active[d]) continue;
    const int r = h->peer[d];
    MW_REQUIRE(r != h->comm->rank, "peer halos: rank %d is its own neighbour in an active direction", r);
    int prev = -1;
    for (int q = 0; q < d; ++q) if (h->dir_active[q] && h->peer[q] == r) prev = q;
    mw_dycore::Peer &R = h->peer_mem[d];
    if (prev >= 0) { R = h->peer_mem[prev]; continue; }    // two ranks in a direction: both neighbours are the same process
    const PeerInfo &I = all[r];
    if (!I.pad) { ok = false; why = "a neighbour has no IPC handles"; break; }
    auto open = [&](const cudaIpcMemHandle_t &hd, void **out) {
      if (!ok) return;
      const cudaError_t oe = cudaIpcOpenMemHandle(out, hd, cudaIpcMemLazyEnablePeerAccess);
      if (oe == cudaSuccess) h->ipc_opened.push_back(*out); else fail("cudaIpcOpenMemHandle", oe);
    };
    for (int b = 0; b < 3; ++b) open(I.q[b], (void **) &R.q[b]);
    open(I.mult, (void **) &R.mult);
    open(I.flags, (void **) &R.flags);
    R.nx = I.nx; R.ny = I.ny; R.pitch = I.pitch; R.zstride = I.zstride; R.vstride = I.vstride;
  }
  // everybody or nobody
  int mine_ok = ok ? 1 : 0, all_ok = 0;
  MW_CUDA_OK(cudaMemcpy(dflag, &mine_ok, sizeof(int), cudaMemcpyHostToDevice));
  MW_NCCL_OK(ncclAllReduce(dflag, dflag, 1, ncclInt, ncclMin, h->comm->comm, 0));
  MW_CUDA_OK(cudaMemcpy(&all_ok, dflag, sizeof(int), cudaMemcpyDeviceToHost));
  cudaFree(dsend); cudaFree(drecv); cudaFree(dflag);
  if (!all_ok) {
    for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    h->ipc_opened.clear();
    for (int d = 0; d < 4; ++d) h->peer_mem[d] = mw_dycore::Peer();
    cudaFree(h->flags); h->flags = nullptr;
    cudaGetLastError();
    if (h->comm->rank == 0 || !ok)
      fprintf(stderr, "[mwb200] rank %d: peer-memory halos unavailable (%s); every rank uses the NCCL exchange\n", h->comm->rank,
              ok ? "another rank could not map its neighbours" : why.c_str());
    return MW_OK;
  }
  // x neighbours share my rows, y neighbours my columns (block decomposition, CPL:147-153)
  for (int d = 0; d < 4; ++d) {
    if (!h->dir_active[d]) continue;
    const mw_dycore::Peer &R = h->peer_mem[d];
    MW_REQUIRE(d < 2 ? R.ny == c.ny : R.nx == c.nx, "peer halos: neighbour %d has a %d x %d block, mine is %d x %d", h->peer[d], R.nx, R.ny, c.nx, c.ny);
  }
  h->peer_halo = true;
  return MW_OK;
}

extern "C" int mw_dycore_attach_comm(mw_dycore *h, mw_comm *comm) {
  MW_REQUIRE(h && comm, "mw_dycore_attach_comm: null argument");
  const mw_config &c = h->cfg;
  MW_REQUIRE(comm->nranks == c.nproc_x * c.nproc_y, "communicator has %d ranks, decomposition is %d x %d", comm->nranks,
             c.nproc_x, c.nproc_y);
  MW_REQUIRE(comm->rank == c.py * c.nproc_x + c.px, "rank %d does not sit at (px,py) = (%d,%d)", comm->rank, c.px, c.py);
  h->comm = comm;
  const bool sim2d = (c.ny_glob == 1);
  auto wrap = [](int v, int n) { return (v % n + n) % n; };                 // periodic neighbours, CPL:169-179
  h->peer[0] = c.py * c.nproc_x + wrap(c.px - 1, c.nproc_x);
  h->peer[1] = c.py * c.nproc_x + wrap(c.px + 1, c.nproc_x);
  h->peer[2] = wrap(c.py - 1, c.nproc_y) * c.nproc_x + c.px;
  h->peer[3] = wrap(c.py + 1, c.nproc_y) * c.nproc_x + c.px;
  h->dir_active[0] = h->dir_active[1] = (c.nproc_x > 1);
  h->dir_active[2] = h->dir_active[3] = (c.nproc_y > 1) && !sim2d;
  {
    // Exchange overlapped with the interior tiles (MW_OVERLAP=1) is off by default: a cell-kernel CTA takes a whole SM
    // (254 registers x 256 threads), so the pack / NCCL / unpack kernels and the boundary tiles each wait for a CTA to
    // retire (0.7 ms) and the boundary tiles end up as a tail -- measured 20.3 ms per step against 18.9 ms with the
    // exchange simply following the stage on the same stream (2 GPUs, 512 x 512 x 128 per GPU, r02k).
    const char *e = getenv("MW_OVERLAP");
    h->overlap = (h->dir_active[0] || h->dir_active[2]) && e && atoi(e) != 0;
    if (h->overlap && !h->cs) {
      // high priority: the few boundary CTAs and the pack / NCCL / unpack kernels are dispatched ahead of the queued
      // interior CTAs as SMs free up, so they never form a tail of their own
      int prio_lo = 0, prio_hi = 0;
      MW_CUDA_OK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      MW_CUDA_OK(cudaStreamCreateWithPriority(&h->cs, cudaStreamNonBlocking, prio_hi));
      MW_CUDA_OK(cudaStreamCreateWithPriority(&h->cs2, cudaStreamNonBlocking, prio_hi));
      for (cudaEvent_t *e : {&h->ev_ready, &h->ev_halo, &h->ev_prev, &h->ev_bnd}) MW_CUDA_OK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
  }
  {
    const char *e = getenv("MW_PEER_HALO");
    if (!(e && atoi(e) == 0) && (h->dir_active[0] || h->dir_active[2])) {
      int rc = setup_peer_halos(h);
      if (rc != MW_OK) return rc;
    }
  }
  if (h->peer_halo) { h->overlap = false; return MW_OK; }
  const size_t T = c.num_tracers;
  for (int d = 0; d < 4; ++d) {
    if (!h->dir_active[d]) continue;
    h->hcount[d] = (size_t) h->N * c.nz * HALO * (d < 2 ? c.ny : c.nx);
    h->mcount[d] = T * c.nz * (d < 2 ? c.ny : c.nx);
    MW_CUDA_OK(cudaMalloc(&h->hsend[d], h->hcount[d] * 8));
    MW_CUDA_OK(cudaMalloc(&h->hrecv[d], h->hcount[d] * 8));
    if (T) {
      MW_CUDA_OK(cudaMalloc(&h->msend[d], h->mcount[d] * 8));
      MW_CUDA_OK(cudaMalloc(&h->mrecv[d], h->mcount[d] * 8));
    }
  }
  return MW_OK;
}

extern "C" int mw_weno5_edges(const double *stencils, double *out, long long n, void *stream) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_REQUIRE(stencils && out && n >= 0, "mw_weno5_edges: bad argument");
  if (n == 0) return MW_OK;
  k_weno5_edges<<<(unsigned) ((n + 255) / 256), 256, 0, (cudaStream_t) stream>>>(stencils, out, n);
  MW_CUDA_OK(cudaGetLastError());
  return MW_OK;
}
