// Supercell initial condition: the reference's init_supercell + convert_dynamics_to_coupler
// (model/modules/dynamics_euler_stratified_wenofv.h:1687-1887, 1891-1951; helper formulas :1144-1193).
// The sounding is horizontally uniform, so the 3-D state is one column: the hydrostatic GLL-quadrature column is
// integrated on the host in the reference's operation order (O(nz) work, once) and a kernel broadcasts it.
#include "mw_common.cuh"
#include <cmath>
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
