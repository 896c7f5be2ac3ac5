"""Host-side supercell initial condition (the reference's Dynamics_Euler_Stratified_WenoFV::init_supercell,
model/modules/dynamics_euler_stratified_wenofv.h:1687-1887, followed by convert_dynamics_to_coupler, :1891-1951).

The supercell sounding is horizontally uniform (u depends on z only), so the whole 3-D initial state is one column:
it is computed here in numpy, in the reference's operation order, and broadcast on the device.  This is host logic
(O(nz) work, once), not the timed path; the thermal bubble is added on the device by mw_perturb_temperature.
"""
import math
import numpy as np

GLL_PTS = np.array([-0.5, -0.32732683535398857189914622812342917778, 0.0, 0.32732683535398857189914622812342917778, 0.5])
GLL_WTS = np.array([0.05, 0.27222222222222222222222222222222222222, 0.35555555555555555555555555555555555556,
                    0.27222222222222222222222222222222222222, 0.05])
ORD = 5


def _temperature(z, z_0, z_trop, z_top, T_0, T_trop, T_top):                          # DYC:1144-1153
    if z <= z_trop:
        lapse = -(T_trop - T_0) / (z_trop - z_0)
        return T_0 - lapse * (z - z_0)
    lapse = -(T_top - T_trop) / (z_top - z_trop)
    return T_trop - lapse * (z - z_trop)


def _pressure_dry(z, z_0, z_trop, z_top, T_0, T_trop, T_top, p_0, R_d, grav):         # DYC:1157-1177
    if z <= z_trop:
        lapse = -(T_trop - T_0) / (z_trop - z_0)
        T = _temperature(z, z_0, z_trop, z_top, T_0, T_trop, T_top)
        return p_0 * math.pow(T / T_0, grav / (R_d * lapse))
    lapse = -(T_trop - T_0) / (z_trop - z_0)
    p_trop = p_0 * math.pow(T_trop / T_0, grav / (R_d * lapse))
    lapse = -(T_top - T_trop) / (z_top - z_trop)
    if lapse != 0:
        T = _temperature(z, z_0, z_trop, z_top, T_0, T_trop, T_top)
        return p_trop * math.pow(T / T_trop, grav / (R_d * lapse))
    return p_trop * math.exp(-grav * (z - z_trop) / (R_d * T_trop))


def _relhum(z, z_0, z_trop):                                                          # DYC:1181-1187
    return 1.0 - 0.75 * math.pow(z / z_trop, 1.25) if z <= z_trop else 0.25


def _sat_mix_dry(press, T):                                                           # DYC:1191-1193
    return 380 / press * math.exp(17.27 * (T - 273) / (T - 36))


def supercell_column(nz, zlen, R_d=287.0, R_v=461.0, cp_d=1003.0, p0=1.0e5, grav=9.81):
    """Returns (bg, col): bg = hy_dens_cells[nz], hy_dens_theta_cells[nz], hy_dens_edges[nz+1], hy_dens_theta_edges[nz+1]
    concatenated; col = dict of coupler-state columns (density_dry, uvel, vvel, wvel, temp, water_vapor), each [nz]."""
    z_0, z_trop, T_0, T_trop, T_top, p_0 = 0.0, 12000.0, 300.0, 213.0, 213.0, 100000.0
    dz = zlen / nz
    ztop = zlen
    gamma = cp_d / (cp_d - R_d)
    kappa = R_d / cp_d
    C0 = math.pow(R_d * math.pow(p0, -kappa), gamma)
    args = (z_0, z_trop, ztop, T_0, T_trop, T_top)

    def qv_at(zloc):
        temp = _temperature(zloc, *args)
        press_dry = _pressure_dry(zloc, *args, p_0, R_d, grav)
        qvs = _sat_mix_dry(press_dry, temp)
        relhum = _relhum(zloc, z_0, z_trop)
        if relhum * qvs > 0.014:
            relhum = 0.014 / qvs
        return min(0.014, qvs * relhum), temp

    # DYC:1736-1756: integrand of d(ln p)/dz at the nested GLL points
    quad = np.zeros((nz, ORD - 1, ORD))
    for k in range(nz):
        cellmid = (k + 0.5) * dz
        for kk in range(ORD - 1):
            ord_b = cellmid + GLL_PTS[kk] * dz
            ord_t = cellmid + GLL_PTS[kk + 1] * dz
            ord_m = 0.5 * (ord_b + ord_t)
            ord_dz = dz * (GLL_PTS[kk + 1] - GLL_PTS[kk])
            for kkk in range(ORD):
                zloc = ord_m + ord_dz * GLL_PTS[kkk]
                qv, temp = qv_at(zloc)
                quad[k, kk, kkk] = -(1 + qv) * grav / (R_d + qv * R_v) / temp
    # DYC:1759-1774: hydrostatic pressure at the GLL points, serial scan upward
    pGLL = np.zeros((nz, ORD))
    pGLL[0, 0] = p_0
    for k in range(nz):
        for kk in range(ORD - 1):
            tot = 0.0
            for kkk in range(ORD):
                tot += quad[k, kk, kkk] * GLL_WTS[kkk]
            tot *= dz * (GLL_PTS[kk + 1] - GLL_PTS[kk])
            pGLL[k, kk + 1] = pGLL[k, kk] * math.exp(tot)
            if kk == ORD - 2 and k < nz - 1:
                pGLL[k + 1, 0] = pGLL[k, ORD - 1]
    # DYC:1777-1805
    dGLL = np.zeros((nz, ORD)); dtGLL = np.zeros((nz, ORD)); dvGLL = np.zeros((nz, ORD))
    hye = np.zeros(nz + 1); hyte = np.zeros(nz + 1)
    for k in range(nz):
        for kk in range(ORD):
            zloc = (k + 0.5) * dz + GLL_PTS[kk] * dz
            qv, temp = qv_at(zloc)
            press = pGLL[k, kk]
            dens_dry = press / (R_d + qv * R_v) / temp
            dens_vap = qv * dens_dry
            dens = dens_dry + dens_vap
            dens_theta = math.pow(press / C0, 1.0 / gamma)
            dGLL[k, kk], dtGLL[k, kk], dvGLL[k, kk] = dens, dens_theta, dens_vap
            if kk == 0:
                hye[k], hyte[k] = dens, dens_theta
            if k == nz - 1 and kk == ORD - 1:
                hye[k + 1], hyte[k + 1] = dens, dens_theta
    # DYC:1808-1840
    hyc = np.zeros(nz); hytc = np.zeros(nz)
    for k in range(nz):
        d = t = 0.0
        for kk in range(ORD):
            d += dGLL[k, kk] * GLL_WTS[kk]
            t += dtGLL[k, kk] * GLL_WTS[kk]
        hyc[k], hytc[k] = d, t
    # DYC:1843-1886: cell averages by 5^3-point quadrature (values depend on the z point only)
    sR = np.zeros(nz); sU = np.zeros(nz); sT = np.zeros(nz); sV = np.zeros(nz)
    for k in range(nz):
        r = u_ = t = v = 0.0
        for kk in range(ORD):
            zloc = (k + 0.5) * dz + GLL_PTS[kk] * dz
            dens = dGLL[k, kk]
            uvel = 30.0 * (zloc / 5000.0) - 15.0 if zloc < 5000.0 else 30.0 - 15.0
            for jj in range(ORD):
                for ii in range(ORD):
                    factor = GLL_WTS[ii] * GLL_WTS[jj] * GLL_WTS[kk]
                    r += (dens - dGLL[k, kk]) * factor
                    u_ += dens * uvel * factor
                    t += (dtGLL[k, kk] - dtGLL[k, kk]) * factor
                    v += dvGLL[k, kk] * factor
        sR[k], sU[k], sT[k], sV[k] = r, u_, t, v
    # DYC:1927-1946 (convert_dynamics_to_coupler)
    rho = sR + hyc
    u = sU / rho
    theta = (sT + hytc) / rho
    press = C0 * np.power(rho * theta, gamma)
    rho_v = sV
    rho_d = rho - rho_v
    temp = press / (rho_d * R_d + rho_v * R_v)
    col = dict(density_dry=rho_d, uvel=u, vvel=np.zeros(nz), wvel=np.zeros(nz), temp=temp, water_vapor=rho_v)
    return np.concatenate([hyc, hytc, hye, hyte]), col
