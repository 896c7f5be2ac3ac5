// Supercell initial condition: the reference's init_supercell + convert_dynamics_to_coupler
// (model/modules/dynamics_euler_stratified_wenofv.h:1687-1887, 1891-1951; helper formulas :1144-1193).
// The sounding is horizontally uniform, so the 3-D state is one column: the hydrostatic GLL-quadrature column is
// integrated on the host in the reference's operation order (O(nz) work, once) and a kernel broadcasts it.
#include "mw_common.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace mw {
namespace {
const double gll_pts[5] = {-0.5, -0.32732683535398857189914622812342917778, 0.0, 0.32732683535398857189914622812342917778, 0.5};
const double gll_wts[5] = {0.05, 0.27222222222222222222222222222222222222, 0.35555555555555555555555555555555555556,
                           0.27222222222222222222222222222222222222, 0.05};
struct Sounding { double z_0 = 0, z_trop = 12000, z_top, T_0 = 300, T_trop = 213, T_top = 213, p_0 = 100000, R_d, R_v, grav; };

double temperature(const Sounding &s, double z) {                                   // DYC:1144-1153
  if (z <= s.z_trop) { const double lapse = -(s.T_trop - s.T_0) / (s.z_trop - s.z_0); return s.T_0 - lapse * (z - s.z_0); }
  const double lapse = -(s.T_top - s.T_trop) / (s.z_top - s.z_trop);
  return s.T_trop - lapse * (z - s.z_trop);
}
double pressure_dry(const Sounding &s, double z) {                                  // DYC:1157-1177
  double lapse = -(s.T_trop - s.T_0) / (s.z_trop - s.z_0);
  if (z <= s.z_trop) return s.p_0 * pow(temperature(s, z) / s.T_0, s.grav / (s.R_d * lapse));
  const double p_trop = s.p_0 * pow(s.T_trop / s.T_0, s.grav / (s.R_d * lapse));
  lapse = -(s.T_top - s.T_trop) / (s.z_top - s.z_trop);
  if (lapse != 0) return p_trop * pow(temperature(s, z) / s.T_trop, s.grav / (s.R_d * lapse));
  return p_trop * exp(-s.grav * (z - s.z_trop) / (s.R_d * s.T_trop));
}
double relhum(const Sounding &s, double z) { return z <= s.z_trop ? 1.0 - 0.75 * pow(z / s.z_trop, 1.25) : 0.25; }   // DYC:1181
double sat_mix_dry(double press, double T) { return 380 / press * exp(17.27 * (T - 273) / (T - 36)); }                   // DYC:1191
double qv_at(const Sounding &s, double z, double &temp) {
  temp = temperature(s, z);
  const double qvs = sat_mix_dry(pressure_dry(s, z), temp);
  double rh = relhum(s, z);
  if (rh * qvs > 0.014) rh = 0.014 / qvs;
  return fmin(0.014, qvs * rh);
}
}  // namespace

struct BroadcastParams {
  double *fields[5 + MW_MAX_TRACERS];
  const double *col;     // [6][nz]: density_dry, uvel, vvel, wvel, temp, water_vapor
  int nf, idWV, nz;
  long long np;
};
__global__ void __launch_bounds__(256) k_broadcast_column(const BroadcastParams B) {
  const long long n = (long long) B.nz * B.np, c = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int k = (int) (c / B.np);
  for (int f = 0; f < B.nf; ++f) {
    double v = 0.0;
    if (f < 5) v = B.col[f * B.nz + k];
    else if (f - 5 == B.idWV) v = B.col[5 * B.nz + k];
    B.fields[f][c] = v;
  }
}
}  // namespace mw
using namespace mw;

extern "C" int mw_dycore_init_supercell(mw_dycore *h, double *const *fields, void *stream) {
  MW_REQUIRE(h && fields, "mw_dycore_init_supercell: null argument");
  // the handle is opaque here: read its configuration back through the public ABI
  mw_config c;
  int rc = mw_dycore_get_config(h, &c);
  if (rc != MW_OK) return rc;
  const int nz = c.nz, ord = 5;
  const double dz = c.zlen / nz;
  Sounding s; s.z_top = c.zlen; s.R_d = c.R_d; s.R_v = c.R_v; s.grav = c.grav;
  std::vector<double> quad((size_t) nz * 4 * 5), pG((size_t) nz * 5), dG((size_t) nz * 5), dtG((size_t) nz * 5), dvG((size_t) nz * 5);
  for (int k = 0; k < nz; ++k) {                                                     // DYC:1736-1756
    const double cellmid = (k + 0.5) * dz;
    for (int kk = 0; kk < ord - 1; ++kk) {
      const double ord_b = cellmid + gll_pts[kk] * dz, ord_t = cellmid + gll_pts[kk + 1] * dz;
      const double ord_m = 0.5 * (ord_b + ord_t), ord_dz = dz * (gll_pts[kk + 1] - gll_pts[kk]);
      for (int kkk = 0; kkk < ord; ++kkk) {
        double temp;
        const double qv = qv_at(s, ord_m + ord_dz * gll_pts[kkk], temp);
        quad[((size_t) k * 4 + kk) * 5 + kkk] = -(1 + qv) * c.grav / (c.R_d + qv * c.R_v) / temp;
      }
    }
  }
  pG[0] = s.p_0;                                                                     // DYC:1759-1774
  for (int k = 0; k < nz; ++k)
    for (int kk = 0; kk < ord - 1; ++kk) {
      double tot = 0;
      for (int kkk = 0; kkk < ord; ++kkk) tot += quad[((size_t) k * 4 + kk) * 5 + kkk] * gll_wts[kkk];
      tot *= dz * (gll_pts[kk + 1] - gll_pts[kk]);
      pG[k * 5 + kk + 1] = pG[k * 5 + kk] * exp(tot);
      if (kk == ord - 2 && k < nz - 1) pG[(k + 1) * 5] = pG[k * 5 + ord - 1];
    }
  std::vector<double> hyc(nz), hytc(nz), hye(nz + 1), hyte(nz + 1);
  for (int k = 0; k < nz; ++k)                                                       // DYC:1777-1805
    for (int kk = 0; kk < ord; ++kk) {
      double temp;
      const double qv = qv_at(s, (k + 0.5) * dz + gll_pts[kk] * dz, temp);
      const double press = pG[k * 5 + kk];
      const double dens_dry = press / (c.R_d + qv * c.R_v) / temp, dens_vap = qv * dens_dry, dens = dens_dry + dens_vap;
      const double dens_theta = pow(press / c.C0, 1.0 / c.gamma_d);
      dG[k * 5 + kk] = dens; dtG[k * 5 + kk] = dens_theta; dvG[k * 5 + kk] = dens_vap;
      if (kk == 0) { hye[k] = dens; hyte[k] = dens_theta; }
      if (k == nz - 1 && kk == ord - 1) { hye[k + 1] = dens; hyte[k + 1] = dens_theta; }
    }
  for (int k = 0; k < nz; ++k) {                                                     // DYC:1808-1840
    double d = 0, t = 0;
    for (int kk = 0; kk < ord; ++kk) { d += dG[k * 5 + kk] * gll_wts[kk]; t += dtG[k * 5 + kk] * gll_wts[kk]; }
    hyc[k] = d; hytc[k] = t;
  }
  std::vector<double> col((size_t) 6 * nz, 0.0);
  for (int k = 0; k < nz; ++k) {                                                     // DYC:1843-1886, then :1927-1946
    double r = 0, u = 0, t = 0, v = 0;
    for (int kk = 0; kk < ord; ++kk) {
      const double zloc = (k + 0.5) * dz + gll_pts[kk] * dz, dens = dG[k * 5 + kk];
      const double uvel = zloc < 5000.0 ? 30.0 * (zloc / 5000.0) - 15.0 : 30.0 - 15.0;
      for (int jj = 0; jj < ord; ++jj)
        for (int ii = 0; ii < ord; ++ii) {
          const double factor = gll_wts[ii] * gll_wts[jj] * gll_wts[kk];
          r += (dens - dG[k * 5 + kk]) * factor;
          u += dens * uvel * factor;
          t += (dtG[k * 5 + kk] - dtG[k * 5 + kk]) * factor;
          v += dvG[k * 5 + kk] * factor;
        }
    }
    const double rho = r + hyc[k], theta = (t + hytc[k]) / rho;
    const double press = c.C0 * pow(rho * theta, c.gamma_d);
    const double rho_v = (c.idWV >= 0) ? v : 0.0, rho_d = rho - rho_v;
    col[0 * nz + k] = rho_d;
    col[1 * nz + k] = u / rho;
    col[4 * nz + k] = press / (rho_d * c.R_d + rho_v * c.R_v);
    col[5 * nz + k] = rho_v;
  }
  rc = mw_dycore_set_background(h, hyc.data(), hytc.data(), hye.data(), hyte.data());
  if (rc != MW_OK) return rc;
  cudaStream_t st = (cudaStream_t) stream;
  double *dcol = nullptr;
  MW_CUDA_OK(cudaMalloc(&dcol, col.size() * 8));
  MW_CUDA_OK(cudaMemcpyAsync(dcol, col.data(), col.size() * 8, cudaMemcpyHostToDevice, st));
  BroadcastParams B;
  B.nf = 5 + c.num_tracers; B.idWV = c.idWV; B.nz = nz; B.np = (long long) c.ny * c.nx; B.col = dcol;
  for (int f = 0; f < B.nf; ++f) B.fields[f] = fields[f];
  const long long n = (long long) nz * B.np;
  k_broadcast_column<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(B);
  MW_CUDA_OK(cudaGetLastError());
  MW_CUDA_OK(cudaStreamSynchronize(st));
  cudaFree(dcol);
  return MW_OK;
}

// ======================================================================================================================
// The other init_data cases of the reference's init (DYC:1338-1653): thermal, building, city.
// ======================================================================================================================
namespace mw {
namespace {
// hydro_const_theta, DYC:1107-1117
__host__ __device__ inline void hydro_const_theta(double z, double grav, double C0, double cp, double p0, double gamma,
                                                  double rd, double &r, double &t) {
  const double theta0 = 300., exner0 = 1.;
  t = theta0;
  const double exner = exner0 - grav * z / (cp * theta0);
  const double p = p0 * pow(exner, (cp / rd));
  const double rt = pow((p / C0), (1. / gamma));
  r = rt / t;
}
// sample_ellipse_cosine, DYC:1121-1133
__device__ inline double ellipse_cosine(double amp, double x, double y, double z, double x0, double y0, double z0,
                                        double xrad, double yrad, double zrad) {
  const double dist = sqrt(((x - x0) / xrad) * ((x - x0) / xrad) + ((y - y0) / yrad) * ((y - y0) / yrad) +
                           ((z - z0) / zrad) * ((z - z0) / zrad)) * M_PI / 2.;
  if (dist <= M_PI / 2.) { const double c = cos(dist); return amp * (c * c); }
  return 0.;
}

struct ThermalParams {
  double *fields[5 + MW_MAX_TRACERS];
  const double *hyc, *hytc;           // device [nz]
  int nf, idWV, nz, ny, nx, i_beg, j_beg, sim2d;
  int adds_mass[MW_MAX_TRACERS];
  double dx, dy, dz, xlen, ylen, grav, C0, gamma, cp_d, p0, R_d, R_v;
};
// One thread per cell: 27-point Gauss-Legendre average of the thermal() point function (DYC:1360-1391, 1086-1103),
// then convert_dynamics_to_coupler (DYC:1927-1946) in registers.
__global__ void __launch_bounds__(128) k_init_thermal(const ThermalParams P) {
  const long long n = (long long) P.nz * P.ny * P.nx, c = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int i = (int) (c % P.nx), j = (int) ((c / P.nx) % P.ny), k = (int) (c / ((long long) P.nx * P.ny));
  const double qp[3] = {0.112701665379258311482073460022, 0.5, 0.887298334620741688517926539980};
  const double qw[3] = {0.277777777777777777777777777779, 0.444444444444444444444444444444, 0.277777777777777777777777777779};
  double sR = 0, sU = 0, sV = 0, sW = 0, sT = 0, sQ = 0;
  for (int kk = 0; kk < 3; ++kk)
    for (int jj = 0; jj < 3; ++jj)
      for (int ii = 0; ii < 3; ++ii) {
        const double x = (i + P.i_beg + 0.5) * P.dx + (qp[ii] - 0.5) * P.dx;
        double y = (j + P.j_beg + 0.5) * P.dy + (qp[jj] - 0.5) * P.dy;
        if (P.sim2d) y = P.ylen / 2;
        const double z = (k + 0.5) * P.dz + (qp[kk] - 0.5) * P.dz;
        double hr, ht;
        hydro_const_theta(z, P.grav, P.C0, P.cp_d, P.p0, P.gamma, P.R_d, hr, ht);
        const double rho_d = hr;
        const double theta_d = ht + ellipse_cosine(2., x, y, z, P.xlen / 2, P.ylen / 2, 2000., 2000., 2000., 2000.);
        const double p_d = P.C0 * pow(rho_d * theta_d, P.gamma);
        const double temp = p_d / rho_d / P.R_d;
        const double tc = temp - 273.15;
        const double sat_pv = 610.94 * exp(17.625 * tc / (243.04 + tc));                    // DYC:1136-1139
        const double sat_rv = sat_pv / P.R_v / temp;
        const double rho_v = ellipse_cosine(0.8, x, y, z, P.xlen / 2, P.ylen / 2, 2000., 2000., 2000., 2000.) * sat_rv;
        const double p = rho_d * P.R_d * temp + rho_v * P.R_v * temp;
        const double rho = rho_d + rho_v;
        const double theta = pow(p / P.C0, 1. / P.gamma) / rho;
        const double wt = qw[ii] * qw[jj] * qw[kk];
        sR += (rho - hr) * wt;
        sU += rho * 0. * wt; sV += rho * 0. * wt; sW += rho * 0. * wt;
        sT += (rho * theta - hr * ht) * wt;
        sQ += rho_v * wt;
      }
  const double rho = sR + P.hyc[k];
  const double theta = (sT + P.hytc[k]) / rho;
  const double press = P.C0 * pow(rho * theta, P.gamma);
  const double rho_v = P.idWV >= 0 ? sQ : 0.0;
  double rho_d = rho;
  for (int tr = 0; tr < P.nf - 5; ++tr)
    if (P.adds_mass[tr]) rho_d -= (tr == P.idWV ? sQ : 0.0);
  P.fields[0][c] = rho_d;
  P.fields[1][c] = sU / rho;
  P.fields[2][c] = sV / rho;
  P.fields[3][c] = sW / rho;
  P.fields[4][c] = press / (rho_d * P.R_d + rho_v * P.R_v);
  for (int tr = 0; tr < P.nf - 5; ++tr) P.fields[5 + tr][c] = (tr == P.idWV ? sQ : 0.0);
}

struct MaskParams {
  double *immersed;
  const double *heights;              // device [nby][nbx] (city) or null (building)
  int nz, ny, nx, i_beg, j_beg, nx_glob, ny_glob;
  int cells_per_building, buildings_pad, nblocks_x, nblocks_y, nbx;
  double dz;
};
// immersed_proportion of the building (DYC:1599-1608) and city (DYC:1503-1514) cases; zero elsewhere (DYC:1425,1549)
template <int CITY>
__global__ void __launch_bounds__(256) k_immersed_mask(const MaskParams M) {
  const long long n = (long long) M.nz * M.ny * M.nx, c = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int i = (int) (c % M.nx), j = (int) ((c / M.nx) % M.ny), k = (int) (c / ((long long) M.nx * M.ny));
  double v = 0.0;
  if (CITY) {
    const int inorm = (M.i_beg + i) / M.cells_per_building - M.buildings_pad;
    const int jnorm = (M.j_beg + j) / M.cells_per_building - M.buildings_pad;
    if ((inorm >= 0 && inorm < M.nblocks_x * 3 && inorm % 3 < 2) && (jnorm >= 0 && jnorm < M.nblocks_y * 9 && jnorm % 9 < 8))
      if (k <= ceil(M.heights[(size_t) jnorm * M.nbx + inorm] / M.dz)) v = 1.0;
  } else {
    const double x0 = 0.3 * M.nx_glob, y0 = 0.5 * M.ny_glob, xr = 0.05 * M.ny_glob, yr = 0.05 * M.ny_glob;
    if (fabs((double) (M.i_beg + i) - x0) <= xr && fabs((double) (M.j_beg + j) - y0) <= yr && k <= 0.2 * M.nz) v = 1.0;
  }
  M.immersed[c] = v;
}

const double gll9_pts[9] = {-0.50000000000000000000000000000000000000, -0.44987899770573007865617262220916897903,
                            -0.33859313975536887672294271354567122536, -0.18155873191308907935537603435432960651,
                            0.00000000000000000000000000000000000000,  0.18155873191308907935537603435432960651,
                            0.33859313975536887672294271354567122536,  0.44987899770573007865617262220916897903,
                            0.50000000000000000000000000000000000000};
const double gll9_wts[9] = {0.013888888888888888888888888888888888889, 0.082747680780402762523169860014604152919,
                            0.13726935625008086764035280928968636297,  0.17321425548652317255756576606985914397,
                            0.18575963718820861678004535147392290249,  0.17321425548652317255756576606985914397,
                            0.13726935625008086764035280928968636297,  0.082747680780402762523169860014604152919,
                            0.013888888888888888888888888888888888889};

int broadcast_column(const mw_config &c, const std::vector<double> &col, double *const *fields, cudaStream_t st) {
  double *dcol = nullptr;
  MW_CUDA_OK(cudaMalloc(&dcol, col.size() * 8));
  MW_CUDA_OK(cudaMemcpyAsync(dcol, col.data(), col.size() * 8, cudaMemcpyHostToDevice, st));
  BroadcastParams B;
  B.nf = 5 + c.num_tracers; B.idWV = c.idWV; B.nz = c.nz; B.np = (long long) c.ny * c.nx; B.col = dcol;
  for (int f = 0; f < B.nf; ++f) B.fields[f] = fields[f];
  const long long n = (long long) c.nz * B.np;
  k_broadcast_column<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(B);
  MW_CUDA_OK(cudaGetLastError());
  MW_CUDA_OK(cudaStreamSynchronize(st));
  cudaFree(dcol);
  return MW_OK;
}

// Uniform 20 m/s flow of the building and city cases (DYC:1460-1500 == 1565-1597, backgrounds :1516-1541 == 1620-1651):
// horizontally uniform, so the 9^3-point quadrature is evaluated once per level on the host in the reference's
// summation order (kk, jj, ii) and broadcast.  Note the reference samples its [-1/2,1/2] GLL points as
// `(q - 0.5) * dz`, i.e. over [z_k - dz/2 - dz/2, z_k]: reproduced as is.
int init_uniform_flow(mw_dycore *h, const mw_config &c, double *const *fields, cudaStream_t st) {
  const int nz = c.nz;
  const double dz = c.zlen / nz;
  std::vector<double> hyc(nz), hytc(nz), hye(nz + 1), hyte(nz + 1), col((size_t) 6 * nz, 0.0);
  if (c.enable_gravity) {
    for (int k = 0; k < nz; ++k) {
      hyc[k] = 0.; hytc[k] = 0.;
      for (int kk = 0; kk < 9; ++kk) {
        const double z = (k + 0.5) * dz + (gll9_pts[kk] - 0.5) * dz;
        double hr, ht;
        hydro_const_theta(z, c.grav, c.C0, c.cp_d, c.p0, c.gamma_d, c.R_d, hr, ht);
        hyc[k] += hr * gll9_wts[kk]; hytc[k] += hr * ht * gll9_wts[kk];
      }
    }
    for (int k = 0; k < nz + 1; ++k) {
      double hr, ht;
      hydro_const_theta(k * dz, c.grav, c.C0, c.cp_d, c.p0, c.gamma_d, c.R_d, hr, ht);
      hye[k] = hr; hyte[k] = hr * ht;
    }
  } else {
    for (int k = 0; k < nz; ++k) { hyc[k] = 1.15; hytc[k] = 1.15 * 300; }
    for (int k = 0; k < nz + 1; ++k) { hye[k] = 1.15; hyte[k] = 1.15 * 300; }
  }
  for (int k = 0; k < nz; ++k) {
    double sR = 0, sU = 0, sV = 0, sW = 0, sT = 0;
    for (int kk = 0; kk < 9; ++kk) {
      const double z = (k + 0.5) * dz + (gll9_pts[kk] - 0.5) * dz;
      double hr = 1.15, ht = 300;
      if (c.enable_gravity) hydro_const_theta(z, c.grav, c.C0, c.cp_d, c.p0, c.gamma_d, c.R_d, hr, ht);
      const double rho = hr, u = 20, v = 0, w = 0, theta = ht;
      for (int jj = 0; jj < 9; ++jj)
        for (int ii = 0; ii < 9; ++ii) {
          const double wt = gll9_wts[ii] * gll9_wts[jj] * gll9_wts[kk];
          sR += (rho - hr) * wt; sU += rho * u * wt; sV += rho * v * wt; sW += rho * w * wt;
          sT += (rho * theta - hr * ht) * wt;
        }
    }
    const double rho = sR + hyc[k], theta = (sT + hytc[k]) / rho;                      // DYC:1927-1946
    const double press = c.C0 * pow(rho * theta, c.gamma_d);
    const double rho_v = 0.0, rho_d = rho;
    col[0 * nz + k] = rho_d; col[1 * nz + k] = sU / rho; col[2 * nz + k] = sV / rho; col[3 * nz + k] = sW / rho;
    col[4 * nz + k] = press / (rho_d * c.R_d + rho_v * c.R_v);
    col[5 * nz + k] = rho_v;
  }
  int rc = mw_dycore_set_background(h, hyc.data(), hytc.data(), hye.data(), hyte.data());
  if (rc != MW_OK) return rc;
  return broadcast_column(c, col, fields, st);
}
}  // namespace
}  // namespace mw

extern "C" int mw_dycore_init_thermal(mw_dycore *h, double *const *fields, void *stream) {
  MW_REQUIRE(h && fields, "mw_dycore_init_thermal: null argument");
  mw_config c;
  int rc = mw_dycore_get_config(h, &c);
  if (rc != MW_OK) return rc;
  const int nz = c.nz;
  const double dz = c.zlen / nz;
  const double qp[3] = {0.112701665379258311482073460022, 0.5, 0.887298334620741688517926539980};
  const double qw[3] = {0.277777777777777777777777777779, 0.444444444444444444444444444444, 0.277777777777777777777777777779};
  std::vector<double> bg((size_t) 4 * nz + 2);
  double *hyc = bg.data(), *hytc = hyc + nz, *hye = hytc + nz, *hyte = hye + nz + 1;
  for (int k = 0; k < nz; ++k) {                                                      // DYC:1395-1407
    hyc[k] = 0.; hytc[k] = 0.;
    for (int kk = 0; kk < 3; ++kk) {
      double hr, ht;
      hydro_const_theta((k + 0.5) * dz + (qp[kk] - 0.5) * dz, c.grav, c.C0, c.cp_d, c.p0, c.gamma_d, c.R_d, hr, ht);
      hyc[k] += hr * qw[kk]; hytc[k] += hr * ht * qw[kk];
    }
  }
  for (int k = 0; k < nz + 1; ++k) {                                                  // DYC:1410-1418
    double hr, ht;
    hydro_const_theta(k * dz, c.grav, c.C0, c.cp_d, c.p0, c.gamma_d, c.R_d, hr, ht);
    hye[k] = hr; hyte[k] = hr * ht;
  }
  rc = mw_dycore_set_background(h, hyc, hytc, hye, hyte);
  if (rc != MW_OK) return rc;
  cudaStream_t st = (cudaStream_t) stream;
  double *dbg = nullptr;
  MW_CUDA_OK(cudaMalloc(&dbg, (size_t) 2 * nz * 8));
  MW_CUDA_OK(cudaMemcpyAsync(dbg, bg.data(), (size_t) 2 * nz * 8, cudaMemcpyHostToDevice, st));
  ThermalParams P;
  P.nf = 5 + c.num_tracers; P.idWV = c.idWV; P.nz = nz; P.ny = c.ny; P.nx = c.nx; P.i_beg = c.i_beg; P.j_beg = c.j_beg;
  P.sim2d = c.ny_glob == 1;
  for (int f = 0; f < P.nf; ++f) P.fields[f] = fields[f];
  for (int t = 0; t < MW_MAX_TRACERS; ++t) P.adds_mass[t] = c.tracer_adds_mass[t];
  P.hyc = dbg; P.hytc = dbg + nz;
  P.dx = c.xlen / c.nx_glob; P.dy = c.ylen / c.ny_glob; P.dz = dz; P.xlen = c.xlen; P.ylen = c.ylen;
  P.grav = c.grav; P.C0 = c.C0; P.gamma = c.gamma_d; P.cp_d = c.cp_d; P.p0 = c.p0; P.R_d = c.R_d; P.R_v = c.R_v;
  const long long n = (long long) nz * c.ny * c.nx;
  k_init_thermal<<<(unsigned) ((n + 127) / 128), 128, 0, st>>>(P);
  MW_CUDA_OK(cudaGetLastError());
  MW_CUDA_OK(cudaStreamSynchronize(st));
  cudaFree(dbg);
  return MW_OK;
}

extern "C" int mw_dycore_init_building(mw_dycore *h, double *const *fields, double *immersed, void *stream) {
  MW_REQUIRE(h && fields && immersed, "mw_dycore_init_building: null argument");
  mw_config c;
  int rc = mw_dycore_get_config(h, &c);
  if (rc != MW_OK) return rc;
  cudaStream_t st = (cudaStream_t) stream;
  rc = init_uniform_flow(h, c, fields, st);
  if (rc != MW_OK) return rc;
  MaskParams M;
  memset(&M, 0, sizeof(M));
  M.immersed = immersed; M.nz = c.nz; M.ny = c.ny; M.nx = c.nx; M.i_beg = c.i_beg; M.j_beg = c.j_beg;
  M.nx_glob = c.nx_glob; M.ny_glob = c.ny_glob; M.dz = c.zlen / c.nz;
  const long long n = (long long) c.nz * c.ny * c.nx;
  k_immersed_mask<0><<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(M);
  MW_CUDA_OK(cudaGetLastError());
  return mw_dycore_set_immersed(h, immersed);
}

extern "C" int mw_city_layout(double xlen, double ylen, int nx_glob, int *cells_per_building, int *nbuildings_y,
                              int *nbuildings_x) {
  MW_REQUIRE(nx_glob > 0 && xlen > 0 && ylen > 0, "mw_city_layout: bad grid");
  const int building_length = 30, buildings_pad = 20;                                 // DYC:1430-1437
  const double dx = xlen / nx_glob;
  const int cpb = (int) std::round(building_length / dx);
  const int nblocks_x = (static_cast<int>(xlen) / building_length - 2 * buildings_pad) / 3;
  const int nblocks_y = (static_cast<int>(ylen) / building_length - 2 * buildings_pad) / 9;
  if (cells_per_building) *cells_per_building = cpb;
  if (nbuildings_x) *nbuildings_x = nblocks_x * 3;
  if (nbuildings_y) *nbuildings_y = nblocks_y * 9;
  return MW_OK;
}

extern "C" int mw_dycore_init_city(mw_dycore *h, double *const *fields, double *immersed,
                                   const double *building_heights_host, int nby, int nbx, void *stream) {
  MW_REQUIRE(h && fields && immersed, "mw_dycore_init_city: null argument");
  mw_config c;
  int rc = mw_dycore_get_config(h, &c);
  if (rc != MW_OK) return rc;
  int cpb, eby, ebx;
  rc = mw_city_layout(c.xlen, c.ylen, c.nx_glob, &cpb, &eby, &ebx);
  if (rc != MW_OK) return rc;
  MW_REQUIRE(cpb >= 1, "mw_dycore_init_city: dx = %g m gives %d cells per 30 m building (the reference divides by it)",
             c.xlen / c.nx_glob, cpb);
  MW_REQUIRE(nby == eby && nbx == ebx, "mw_dycore_init_city: heights are [%d][%d], the domain needs [%d][%d]", nby, nbx, eby, ebx);
  MW_REQUIRE(nby * nbx == 0 || building_heights_host, "mw_dycore_init_city: null building heights");
  cudaStream_t st = (cudaStream_t) stream;
  rc = init_uniform_flow(h, c, fields, st);
  if (rc != MW_OK) return rc;
  double *dh = nullptr;
  const size_t nh = (size_t) std::max(nby, 0) * std::max(nbx, 0);
  if (nh) {
    MW_CUDA_OK(cudaMalloc(&dh, nh * 8));
    MW_CUDA_OK(cudaMemcpyAsync(dh, building_heights_host, nh * 8, cudaMemcpyHostToDevice, st));
  }
  MaskParams M;
  memset(&M, 0, sizeof(M));
  M.immersed = immersed; M.heights = dh; M.nz = c.nz; M.ny = c.ny; M.nx = c.nx; M.i_beg = c.i_beg; M.j_beg = c.j_beg;
  M.nx_glob = c.nx_glob; M.ny_glob = c.ny_glob; M.dz = c.zlen / c.nz;
  M.cells_per_building = cpb; M.buildings_pad = 20; M.nblocks_x = ebx / 3; M.nblocks_y = eby / 9; M.nbx = nbx;
  const long long n = (long long) c.nz * c.ny * c.nx;
  k_immersed_mask<1><<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(M);
  MW_CUDA_OK(cudaGetLastError());
  MW_CUDA_OK(cudaStreamSynchronize(st));
  if (dh) cudaFree(dh);
  return mw_dycore_set_immersed(h, immersed);
}
