// d3q19_thermal.inl -- per-cell arithmetic of the thermal double-distribution path, included by
// lbm_strict.cu and lbm_fast.cu after d3q19_mrt.inl.
//
// Fixed by the reference (MPI/Buoyancy_driven_cavity/fortran/3d/bouyancy3d_mpi.F90, "B3"):
//   collision()   B3:656-856  D3Q19 MRT + Coriolis/Boussinesq body force, Guo-style moment-space source
//   macro()       B3:995-1002 u = (sum f e + F/2) / rho
//   collisionT()  B3:1028-1062 D3Q7 MRT for the temperature populations
// MGLC_STRICT build (-fmad=false): operation order and divisions follow the reference expression for
// expression (bit-identical to the oracle and to the machine-evaluated Fortran text).
// Fast build (-fmad=true): same algebra on shared partial sums with constant reciprocals and FMA.

using mglc::ThermalParams;

// macro() with the half-force correction; bit-identical in both builds (explicit _rn operations)
__device__ __forceinline__ void d3q19_macro_forced(const double (&f)[19], double Fx, double Fy, double Fz, double &rho,
                                                   double &u, double &v, double &w) {
    double r = f[0];
    r = __dadd_rn(r, f[1]);  r = __dadd_rn(r, f[2]);  r = __dadd_rn(r, f[3]);  r = __dadd_rn(r, f[4]);
    r = __dadd_rn(r, f[5]);  r = __dadd_rn(r, f[6]);  r = __dadd_rn(r, f[7]);  r = __dadd_rn(r, f[8]);
    r = __dadd_rn(r, f[9]);  r = __dadd_rn(r, f[10]); r = __dadd_rn(r, f[11]); r = __dadd_rn(r, f[12]);
    r = __dadd_rn(r, f[13]); r = __dadd_rn(r, f[14]); r = __dadd_rn(r, f[15]); r = __dadd_rn(r, f[16]);
    r = __dadd_rn(r, f[17]); r = __dadd_rn(r, f[18]);
    const double su = f[1] - f[2] + f[7] - f[8] + f[9] - f[10] + f[11] - f[12] + f[13] - f[14];
    const double sv = f[3] - f[4] + f[7] + f[8] - f[9] - f[10] + f[15] - f[16] + f[17] - f[18];
    const double sw = f[5] - f[6] + f[11] + f[12] - f[13] - f[14] + f[15] + f[16] - f[17] - f[18];
    rho = r;
    u = __ddiv_rn(__dadd_rn(su, __dmul_rn(0.5, Fx)), r);
    v = __ddiv_rn(__dadd_rn(sv, __dmul_rn(0.5, Fy)), r);
    w = __ddiv_rn(__dadd_rn(sw, __dmul_rn(0.5, Fz)), r);
}

// macroT(), B3:1215-1232 (adds only: bit-identical in both builds)
__device__ __forceinline__ double d3q7_temperature(const double (&g)[7]) {
    double t = __dadd_rn(g[0], g[1]);
    t = __dadd_rn(t, g[2]); t = __dadd_rn(t, g[3]); t = __dadd_rn(t, g[4]); t = __dadd_rn(t, g[5]);
    return __dadd_rn(t, g[6]);
}

// body force from the fields the collision sees, B3:743-745 (_rn: identical in both builds, because the
// SAME values feed this step's macro())
__device__ __forceinline__ void thermal_force(double rho, double u, double v, double T, const ThermalParams &p,
                                              double &Fx, double &Fy, double &Fz) {
    Fx = __dmul_rn(__dmul_rn(__dmul_rn(-2.0, rho), v), p.omegaRot);
    Fy = __dmul_rn(__dmul_rn(__dmul_rn(2.0, rho), u), p.omegaRot);
    Fz = __dmul_rn(__dmul_rn(rho, p.gBeta), __dsub_rn(T, p.Tref));
}

#ifdef MGLC_STRICT

__device__ __forceinline__ void d3q19_collide_thermal(const double (&f)[19], double rho, double u, double v, double w,
                                                      double Fx, double Fy, double Fz, const ThermalParams &p,
                                                      double (&fp)[19]) {
    double m[19], meq[19], fs[19], mp[19];
    m[0] = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6]
         + f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] + f[15] + f[16] + f[17] + f[18];
    m[1] = -30.0 * f[0] - 11.0 * (f[1] + f[2] + f[3] + f[4] + f[5] + f[6])
         + 8.0 * (f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] + f[15] + f[16] + f[17] + f[18]);
    m[2] = 12.0 * f[0] - 4.0 * (f[1] + f[2] + f[3] + f[4] + f[5] + f[6])
         + f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] + f[15] + f[16] + f[17] + f[18];
    m[3] = f[1] - f[2] + f[7] - f[8] + f[9] - f[10] + f[11] - f[12] + f[13] - f[14];
    m[4] = -4.0 * (f[1] - f[2]) + f[7] - f[8] + f[9] - f[10] + f[11] - f[12] + f[13] - f[14];
    m[5] = f[3] - f[4] + f[7] + f[8] - f[9] - f[10] + f[15] - f[16] + f[17] - f[18];
    m[6] = -4.0 * (f[3] - f[4]) + f[7] + f[8] - f[9] - f[10] + f[15] - f[16] + f[17] - f[18];
    m[7] = f[5] - f[6] + f[11] + f[12] - f[13] - f[14] + f[15] + f[16] - f[17] - f[18];
    m[8] = -4.0 * (f[5] - f[6]) + f[11] + f[12] - f[13] - f[14] + f[15] + f[16] - f[17] - f[18];
    m[9] = 2.0 * (f[1] + f[2]) - f[3] - f[4] - f[5] - f[6]
         + f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] - 2.0 * (f[15] + f[16] + f[17] + f[18]);
    m[10] = -4.0 * (f[1] + f[2]) + 2.0 * (f[3] + f[4] + f[5] + f[6])
          + f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] - 2.0 * (f[15] + f[16] + f[17] + f[18]);
    m[11] = f[3] + f[4] - f[5] - f[6] + f[7] + f[8] + f[9] + f[10] - (f[11] + f[12] + f[13] + f[14]);
    m[12] = -2.0 * (f[3] + f[4] - f[5] - f[6]) + (f[7] + f[8] + f[9] + f[10]) - (f[11] + f[12] + f[13] + f[14]);
    m[13] = f[7] - f[8] - f[9] + f[10];
    m[14] = f[15] - f[16] - f[17] + f[18];
    m[15] = f[11] - f[12] - f[13] + f[14];
    m[16] = f[7] - f[8] + f[9] - f[10] - f[11] + f[12] - f[13] + f[14];
    m[17] = -f[7] - f[8] + f[9] + f[10] + f[15] - f[16] + f[17] - f[18];
    m[18] = f[11] + f[12] - f[13] - f[14] - f[15] - f[16] + f[17] + f[18];

    meq[0] = rho;
    meq[1] = -11.0 * rho + 19.0 * rho * (u * u + v * v + w * w);
    meq[2] = 3.0 * rho - 11.0 / 2.0 * rho * (u * u + v * v + w * w);
    meq[3] = rho * u;
    meq[4] = -2.0 / 3.0 * meq[3];
    meq[5] = rho * v;
    meq[6] = -2.0 / 3.0 * meq[5];
    meq[7] = rho * w;
    meq[8] = -2.0 / 3.0 * meq[7];
    meq[9] = rho * (2.0 * u * u - v * v - w * w);
    meq[10] = -0.5 * meq[9];
    meq[11] = rho * (v * v - w * w);
    meq[12] = -0.5 * meq[11];                       // WITH rho here (B3:715), unlike the lid driver
    meq[13] = rho * (u * v);
    meq[14] = rho * (v * w);
    meq[15] = rho * (w * u);
    meq[16] = 0.0; meq[17] = 0.0; meq[18] = 0.0;

    fs[0] = 0.0;
    fs[1] = 38.0 * (u * Fx + v * Fy + w * Fz);
    fs[2] = -11.0 * (u * Fx + v * Fy + w * Fz);
    fs[3] = Fx;
    fs[4] = -2.0 / 3.0 * Fx;
    fs[5] = Fy;
    fs[6] = -2.0 / 3.0 * Fy;
    fs[7] = Fz;
    fs[8] = -2.0 / 3.0 * Fz;
    fs[9] = 4.0 * u * Fx - 2.0 * v * Fy - 2.0 * w * Fz;
    fs[10] = -2.0 * u * Fx + v * Fy + w * Fz;
    fs[11] = 2.0 * v * Fy - 2.0 * w * Fz;
    fs[12] = -v * Fy + w * Fz;
    fs[13] = u * Fy + v * Fx;
    fs[14] = v * Fz + w * Fy;
    fs[15] = u * Fz + w * Fx;
    fs[16] = 0.0; fs[17] = 0.0; fs[18] = 0.0;

    const double Snu = p.Snu, Sq = p.Sq;
    const double s[19] = {0.0, Snu, Snu, 0.0, Sq, 0.0, Sq, 0.0, Sq, Snu, Snu, Snu, Snu, Snu, Snu, Snu, Sq, Sq, Sq};
#pragma unroll
    for (int a = 0; a < 19; ++a) mp[a] = m[a] - s[a] * (m[a] - meq[a]) + (1.0 - 0.5 * s[a]) * fs[a];

    fp[0] = mp[0] / 19.0 - 5.0 / 399.0 * mp[1] + mp[2] / 21.0;
    fp[1] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 + (mp[3] - mp[4]) * 0.1 + (mp[9] - mp[10]) / 18.0;
    fp[2] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 - (mp[3] - mp[4]) * 0.1 + (mp[9] - mp[10]) / 18.0;
    fp[3] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 + (mp[5] - mp[6]) * 0.1 - (mp[9] - mp[10]) / 36.0 + (mp[11] - mp[12]) / 12.0;
    fp[4] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 - (mp[5] - mp[6]) * 0.1 - (mp[9] - mp[10]) / 36.0 + (mp[11] - mp[12]) / 12.0;
    fp[5] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 + (mp[7] - mp[8]) * 0.1 - (mp[9] - mp[10]) / 36.0 - (mp[11] - mp[12]) / 12.0;
    fp[6] = mp[0] / 19.0 - 11.0 / 2394.0 * mp[1] - mp[2] / 63.0 - (mp[7] - mp[8]) * 0.1 - (mp[9] - mp[10]) / 36.0 - (mp[11] - mp[12]) / 12.0;
    fp[7] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 + 0.025 * (4.0 * mp[3] + mp[4] + 4.0 * mp[5] + mp[6])
          + mp[9] / 36.0 + mp[10] / 72.0 + mp[11] / 12.0 + mp[12] / 24.0 + mp[13] * 0.25 + (mp[16] - mp[17]) * 0.125;
    fp[8] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 - 0.025 * (4.0 * mp[3] + mp[4] - 4.0 * mp[5] - mp[6])
          + mp[9] / 36.0 + mp[10] / 72.0 + mp[11] / 12.0 + mp[12] / 24.0 - mp[13] * 0.25 - (mp[16] + mp[17]) * 0.125;
    fp[9] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 + 0.025 * (4.0 * mp[3] + mp[4] - 4.0 * mp[5] - mp[6])
          + mp[9] / 36.0 + mp[10] / 72.0 + mp[11] / 12.0 + mp[12] / 24.0 - mp[13] * 0.25 + (mp[16] + mp[17]) * 0.125;
    fp[10] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 - 0.025 * (4.0 * mp[3] + mp[4] + 4.0 * mp[5] + mp[6])
           + mp[9] / 36.0 + mp[10] / 72.0 + mp[11] / 12.0 + mp[12] / 24.0 + mp[13] * 0.25 - (mp[16] - mp[17]) * 0.125;
    fp[11] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 + 0.025 * (4.0 * mp[3] + mp[4] + 4.0 * mp[7] + mp[8])
           + mp[9] / 36.0 + mp[10] / 72.0 - mp[11] / 12.0 - mp[12] / 24.0 + 0.25 * mp[15] - 0.1250 * (mp[16] - mp[18]);
    fp[12] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 - 0.025 * (4.0 * mp[3] + mp[4] - 4.0 * mp[7] - mp[8])
           + mp[9] / 36.0 + mp[10] / 72.0 - mp[11] / 12.0 - mp[12] / 24.0 - 0.25 * mp[15] + 0.125 * (mp[16] + mp[18]);
    fp[13] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 + 0.025 * (4.0 * mp[3] + mp[4] - 4.0 * mp[7] - mp[8])
           + mp[9] / 36.0 + mp[10] / 72.0 - mp[11] / 12.0 - mp[12] / 24.0 - 0.25 * mp[15] - 0.125 * (mp[16] + mp[18]);
    fp[14] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 - 0.025 * (4.0 * mp[3] + mp[4] + 4.0 * mp[7] + mp[8])
           + mp[9] / 36.0 + mp[10] / 72.0 - mp[11] / 12.0 - mp[12] / 24.0 + 0.25 * mp[15] + 0.125 * (mp[16] - mp[18]);
    fp[15] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 + (4.0 * mp[5] + mp[6] + 4.0 * mp[7] + mp[8]) * 0.025
           - (mp[9] + mp[10] * 0.5) / 18.0 + 0.25 * mp[14] + 0.125 * (mp[17] - mp[18]);
    fp[16] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 - (4.0 * mp[5] + mp[6] - 4.0 * mp[7] - mp[8]) * 0.025
           - (mp[9] + mp[10] * 0.5) / 18.0 - 0.25 * mp[14] - 0.125 * (mp[17] + mp[18]);
    fp[17] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 + (4.0 * mp[5] + mp[6] - 4.0 * mp[7] - mp[8]) * 0.025
           - (mp[9] + mp[10] * 0.5) / 18.0 - 0.25 * mp[14] + 0.125 * (mp[17] + mp[18]);
    fp[18] = mp[0] / 19.0 + 4.0 / 1197.0 * mp[1] + mp[2] / 252.0 - (4.0 * mp[5] + mp[6] + 4.0 * mp[7] + mp[8]) * 0.025
           - (mp[9] + mp[10] * 0.5) / 18.0 + 0.25 * mp[14] - 0.125 * (mp[17] - mp[18]);
}

__device__ __forceinline__ void d3q7_collide(const double (&g)[7], double u, double v, double w, double T,
                                             const ThermalParams &p, double (&gp)[7]) {
    double n[7], neq[7], np_[7];
    n[0] = g[0] + g[1] + g[2] + g[3] + g[4] + g[5] + g[6];
    n[1] = g[1] - g[2];
    n[2] = g[3] - g[4];
    n[3] = g[5] - g[6];
    n[4] = -6.0 * g[0] + g[1] + g[2] + g[3] + g[4] + g[5] + g[6];
    n[5] = 2.0 * g[1] + 2.0 * g[2] - g[3] - g[4] - g[5] - g[6];
    n[6] = g[3] + g[4] - g[5] - g[6];
    neq[0] = T; neq[1] = T * u; neq[2] = T * v; neq[3] = T * w; neq[4] = T * p.paraA; neq[5] = 0.0; neq[6] = 0.0;
    const double q[7] = {0.0, p.Qd, p.Qd, p.Qd, p.Qnu, p.Qnu, p.Qnu};
#pragma unroll
    for (int a = 0; a < 7; ++a) np_[a] = n[a] - q[a] * (n[a] - neq[a]);
    gp[0] = np_[0] / 7.0 - np_[4] / 7.0;
    gp[1] = np_[0] / 7.0 + 0.5 * np_[1] + np_[4] / 42.0 + np_[5] / 6.0;
    gp[2] = np_[0] / 7.0 - 0.5 * np_[1] + np_[4] / 42.0 + np_[5] / 6.0;
    gp[3] = np_[0] / 7.0 + 0.5 * np_[2] + np_[4] / 42.0 - np_[5] / 12.0 + 0.25 * np_[6];
    gp[4] = np_[0] / 7.0 - 0.5 * np_[2] + np_[4] / 42.0 - np_[5] / 12.0 + 0.25 * np_[6];
    gp[5] = np_[0] / 7.0 + 0.5 * np_[3] + np_[4] / 42.0 - np_[5] / 12.0 - 0.25 * np_[6];
    gp[6] = np_[0] / 7.0 - 0.5 * np_[3] + np_[4] / 42.0 - np_[5] / 12.0 - 0.25 * np_[6];
}

#else  // ---------------------------------------- fast build -----------------------------------------

// Same restructuring as the lid operator (d3q19_mrt.inl): shared pair sums, post-collision moments
// pre-divided by |row|^2, M^-1 = M^T diag(1/|row|^2); the moment-space source enters each relaxed moment
// as + (1 - s/2) * S.
__device__ __forceinline__ void d3q19_collide_thermal(const double (&f)[19], double rho, double u, double v, double w,
                                                      double Fx, double Fy, double Fz, const ThermalParams &p,
                                                      double (&fp)[19]) {
    const double Snu = p.Snu, Sq = p.Sq;
    const double cnu = 1.0 - 0.5 * Snu, cq = 1.0 - 0.5 * Sq;
    const double a12 = f[1] + f[2], d12 = f[1] - f[2];
    const double a34 = f[3] + f[4], d34 = f[3] - f[4];
    const double a56 = f[5] + f[6], d56 = f[5] - f[6];
    const double sxy = (f[7] + f[8]) + (f[9] + f[10]);
    const double xy_x = (f[7] - f[8]) + (f[9] - f[10]);
    const double xy_y = (f[7] + f[8]) - (f[9] + f[10]);
    const double xy_c = (f[7] - f[8]) - (f[9] - f[10]);
    const double sxz = (f[11] + f[12]) + (f[13] + f[14]);
    const double xz_x = (f[11] - f[12]) + (f[13] - f[14]);
    const double xz_z = (f[11] + f[12]) - (f[13] + f[14]);
    const double xz_c = (f[11] - f[12]) - (f[13] - f[14]);
    const double syz = (f[15] + f[16]) + (f[17] + f[18]);
    const double yz_y = (f[15] - f[16]) + (f[17] - f[18]);
    const double yz_z = (f[15] + f[16]) - (f[17] + f[18]);
    const double yz_c = (f[15] - f[16]) - (f[17] - f[18]);

    const double ax = a12 + a34 + a56, dg = sxy + sxz + syz;
    const double m0 = f[0] + ax + dg;
    const double m1 = -30.0 * f[0] - 11.0 * ax + 8.0 * dg;
    const double m2 = 12.0 * f[0] - 4.0 * ax + dg;
    const double jx_d = xy_x + xz_x, jy_d = xy_y + yz_y, jz_d = xz_z + yz_z;
    const double m3 = d12 + jx_d, m4 = jx_d - 4.0 * d12;
    const double m5 = d34 + jy_d, m6 = jy_d - 4.0 * d34;
    const double m7 = d56 + jz_d, m8 = jz_d - 4.0 * d56;
    const double t9 = sxy + sxz - 2.0 * syz;
    const double m9 = 2.0 * a12 - a34 - a56 + t9;
    const double m10 = -4.0 * a12 + 2.0 * (a34 + a56) + t9;
    const double t11 = sxy - sxz;
    const double m11 = a34 - a56 + t11;
    const double m12 = -2.0 * (a34 - a56) + t11;
    const double m13 = xy_c, m14 = yz_c, m15 = xz_c;
    const double m16 = xy_x - xz_x, m17 = yz_y - xy_y, m18 = xz_z - yz_z;

    const double uu = u * u, vv = v * v, ww = w * w, us2 = uu + vv + ww;
    const double uFx = u * Fx, vFy = v * Fy, wFz = w * Fz, uF = uFx + vFy + wFz;
    const double q0 = m0 * (1.0 / 19.0);
    const double q1 = (m1 - Snu * (m1 - rho * (-11.0 + 19.0 * us2)) + cnu * 38.0 * uF) * (1.0 / 2394.0);
    const double q2 = (m2 - Snu * (m2 - rho * (3.0 - 5.5 * us2)) - cnu * 11.0 * uF) * (1.0 / 252.0);
    const double q3 = (m3 + Fx) * 0.1, q5 = (m5 + Fy) * 0.1, q7 = (m7 + Fz) * 0.1;        // s = 0: + 1 * S
    const double q4 = (m4 - Sq * (m4 + (2.0 / 3.0) * rho * u) - cq * (2.0 / 3.0) * Fx) * (1.0 / 40.0);
    const double q6 = (m6 - Sq * (m6 + (2.0 / 3.0) * rho * v) - cq * (2.0 / 3.0) * Fy) * (1.0 / 40.0);
    const double q8 = (m8 - Sq * (m8 + (2.0 / 3.0) * rho * w) - cq * (2.0 / 3.0) * Fz) * (1.0 / 40.0);
    const double e9 = rho * (2.0 * uu - vv - ww), s9 = 4.0 * uFx - 2.0 * vFy - 2.0 * wFz;
    const double q9 = (m9 - Snu * (m9 - e9) + cnu * s9) * (1.0 / 36.0);
    const double q10 = (m10 - Snu * (m10 + 0.5 * e9) - cnu * 0.5 * s9) * (1.0 / 72.0);
    const double e11 = rho * (vv - ww), s11 = 2.0 * vFy - 2.0 * wFz;
    const double q11 = (m11 - Snu * (m11 - e11) + cnu * s11) * (1.0 / 12.0);
    const double q12 = (m12 - Snu * (m12 + 0.5 * e11) - cnu * 0.5 * s11) * (1.0 / 24.0);
    const double q13 = (m13 - Snu * (m13 - rho * u * v) + cnu * (u * Fy + v * Fx)) * 0.25;
    const double q14 = (m14 - Snu * (m14 - rho * v * w) + cnu * (v * Fz + w * Fy)) * 0.25;
    const double q15 = (m15 - Snu * (m15 - rho * u * w) + cnu * (u * Fz + w * Fx)) * 0.25;
    const double q16 = (m16 - Sq * m16) * 0.125;
    const double q17 = (m17 - Sq * m17) * 0.125;
    const double q18 = (m18 - Sq * m18) * 0.125;

    fp[0] = q0 - 30.0 * q1 + 12.0 * q2;
    const double cax = q0 - 11.0 * q1 - 4.0 * q2;
    const double cdg = q0 + 8.0 * q1 + q2;
    const double ax_x = cax + 2.0 * (q9 - 2.0 * q10);
    const double ax_yz = cax - (q9 - 2.0 * q10);
    const double n11 = q11 - 2.0 * q12;
    const double jx = q3 - 4.0 * q4, jy = q5 - 4.0 * q6, jz = q7 - 4.0 * q8;
    fp[1] = ax_x + jx;
    fp[2] = ax_x - jx;
    fp[3] = ax_yz + n11 + jy;
    fp[4] = ax_yz + n11 - jy;
    fp[5] = ax_yz - n11 + jz;
    fp[6] = ax_yz - n11 - jz;
    const double p9 = q9 + q10, p11 = q11 + q12;
    const double kx = q3 + q4, ky = q5 + q6, kz = q7 + q8;
    const double cxy = cdg + p9 + p11, cxz = cdg + p9 - p11, cyz = cdg - 2.0 * p9;
    fp[7]  = cxy + (kx + q16) + (ky - q17) + q13;
    fp[8]  = cxy - (kx + q16) + (ky - q17) - q13;
    fp[9]  = cxy + (kx + q16) - (ky - q17) - q13;
    fp[10] = cxy - (kx + q16) - (ky - q17) + q13;
    fp[11] = cxz + (kx - q16) + (kz + q18) + q15;
    fp[12] = cxz - (kx - q16) + (kz + q18) - q15;
    fp[13] = cxz + (kx - q16) - (kz + q18) - q15;
    fp[14] = cxz - (kx - q16) - (kz + q18) + q15;
    fp[15] = cyz + (ky + q17) + (kz - q18) + q14;
    fp[16] = cyz - (ky + q17) + (kz - q18) - q14;
    fp[17] = cyz + (ky + q17) - (kz - q18) - q14;
    fp[18] = cyz - (ky + q17) - (kz - q18) + q14;
}

// D3Q7: N rows have squared norms {7, 2, 2, 2, 42, 12, 4}; g = N^T diag(1/|row|^2) n*
__device__ __forceinline__ void d3q7_collide(const double (&g)[7], double u, double v, double w, double T,
                                             const ThermalParams &p, double (&gp)[7]) {
    const double a12 = g[1] + g[2], a34 = g[3] + g[4], a56 = g[5] + g[6];
    const double ax = a12 + a34 + a56;
    const double n0 = g[0] + ax;
    const double n1 = g[1] - g[2], n2 = g[3] - g[4], n3 = g[5] - g[6];
    const double n4 = ax - 6.0 * g[0];
    const double n5 = 2.0 * a12 - a34 - a56;
    const double n6 = a34 - a56;
    const double r0 = n0 * (1.0 / 7.0);
    const double r1 = (n1 - p.Qd * (n1 - T * u)) * 0.5;
    const double r2 = (n2 - p.Qd * (n2 - T * v)) * 0.5;
    const double r3 = (n3 - p.Qd * (n3 - T * w)) * 0.5;
    const double r4 = (n4 - p.Qnu * (n4 - T * p.paraA)) * (1.0 / 42.0);
    const double r5 = (n5 - p.Qnu * n5) * (1.0 / 12.0);
    const double r6 = (n6 - p.Qnu * n6) * 0.25;
    gp[0] = r0 - 6.0 * r4;
    const double c = r0 + r4;
    const double cx = c + 2.0 * r5, cy = c - r5 + r6, cz = c - r5 - r6;
    gp[1] = cx + r1; gp[2] = cx - r1;
    gp[3] = cy + r2; gp[4] = cy - r2;
    gp[5] = cz + r3; gp[6] = cz - r3;
}
#endif
