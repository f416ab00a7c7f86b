// d2q9_thermal.inl -- per-cell arithmetic of the 2-D thermal path (B2 = MPI/Buoyancy_driven_cavity/fortran/2d/mpi_blocked/):
// collision() evolution_f.F90:15-78, collisionT() evolution_g.F90:13-39, macro()/macroT() evolution_f.F90:335-337 /
// evolution_g.F90:171.  Included inside a namespace by thermal2d_kernels.inl (device) and by tests/host_shim/t2d_host.cpp (host,
// CPU-only check against the oracle).  MGLC_STRICT: the reference's operation order, true divisions -- bit-identical to the
// oracle when built without FMA contraction; otherwise the throughput form (shared partial sums, conserved moments passed
// through, constant reciprocals, FMA).

// out: fp[9] and the force Fy = rho*gBeta*(T-Tref) (Fx is identically 0, evolution_f.F90:45)
// acc = true: the OpenACC program's collision() (seq/bouyancy2d_acc.F90:631-695), identical except f_post(0) = m0/9 - m1/9 + m2/9
template <bool ACC>
__device__ __forceinline__ void t2_collide(const double (&f)[9], double rho, double u, double v, double T, double Snu, double Sq,
                                           double gBeta, double Tref, double (&fp)[9], double &Fy_out) {
#ifdef MGLC_STRICT
    double m[9], meq[9], mp[9], fs[9];
    const double s[9] = {0.0, Snu, Snu, 0.0, Sq, 0.0, Sq, Snu, Snu};
    m[0] = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
    m[1] = -4.0 * f[0] - f[1] - f[2] - f[3] - f[4] + 2.0 * (f[5] + f[6] + f[7] + f[8]);
    m[2] = 4.0 * f[0] - 2.0 * (f[1] + f[2] + f[3] + f[4]) + f[5] + f[6] + f[7] + f[8];
    m[3] = f[1] - f[3] + f[5] - f[6] - f[7] + f[8];
    m[4] = -2.0 * f[1] + 2.0 * f[3] + f[5] - f[6] - f[7] + f[8];
    m[5] = f[2] - f[4] + f[5] + f[6] - f[7] - f[8];
    m[6] = -2.0 * f[2] + 2.0 * f[4] + f[5] + f[6] - f[7] - f[8];
    m[7] = f[1] - f[2] + f[3] - f[4];
    m[8] = f[5] - f[6] + f[7] - f[8];
    meq[0] = rho;
    meq[1] = rho * (-2.0 + 3.0 * (u * u + v * v));
    meq[2] = rho * (1.0 - 3.0 * (u * u + v * v));
    meq[3] = rho * u;
    meq[4] = -(rho * u);
    meq[5] = rho * v;
    meq[6] = -(rho * v);
    meq[7] = rho * (u * u - v * v);
    meq[8] = rho * (u * v);
    const double Fx = 0.0;
    const double Fy = rho * gBeta * (T - Tref);
    fs[0] = 0.0;
    fs[1] = (6.0 - 3.0 * s[1]) * (u * Fx + v * Fy);
    fs[2] = -((6.0 - 3.0 * s[2]) * (u * Fx + v * Fy));
    fs[3] = (1.0 - 0.5 * s[3]) * Fx;
    fs[4] = -((1.0 - 0.5 * s[4]) * Fx);
    fs[5] = (1.0 - 0.5 * s[5]) * Fy;
    fs[6] = -((1.0 - 0.5 * s[6]) * Fy);
    fs[7] = (2.0 - s[7]) * (u * Fx - v * Fy);
    fs[8] = (1.0 - 0.5 * s[8]) * (u * Fy + v * Fx);
#pragma unroll
    for (int a = 0; a < 9; ++a) mp[a] = m[a] - s[a] * (m[a] - meq[a]) + fs[a];
    fp[0] = ACC ? mp[0] / 9.0 - mp[1] / 9.0 + mp[2] / 9.0 : (mp[0] - mp[1] + mp[2]) / 9.0;
    fp[1] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 + mp[3] / 6.0 - mp[4] / 6.0 + mp[7] / 4.0;
    fp[2] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 + mp[5] / 6.0 - mp[6] / 6.0 - mp[7] / 4.0;
    fp[3] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 - mp[3] / 6.0 + mp[4] / 6.0 + mp[7] / 4.0;
    fp[4] = mp[0] / 9.0 - mp[1] / 36.0 - mp[2] / 18.0 - mp[5] / 6.0 + mp[6] / 6.0 - mp[7] / 4.0;
    fp[5] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 + mp[3] / 6.0 + mp[4] / 12.0 + mp[5] / 6.0 + mp[6] / 12.0 + mp[8] / 4.0;
    fp[6] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 - mp[3] / 6.0 - mp[4] / 12.0 + mp[5] / 6.0 + mp[6] / 12.0 - mp[8] / 4.0;
    fp[7] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 - mp[3] / 6.0 - mp[4] / 12.0 - mp[5] / 6.0 - mp[6] / 12.0 + mp[8] / 4.0;
    fp[8] = mp[0] / 9.0 + mp[1] / 18.0 + mp[2] / 36.0 + mp[3] / 6.0 + mp[4] / 12.0 - mp[5] / 6.0 - mp[6] / 12.0 - mp[8] / 4.0;
    Fy_out = Fy;
#else
    const double sa = (f[1] + f[3]) + (f[2] + f[4]), sd = (f[5] + f[7]) + (f[6] + f[8]);
    const double m0 = f[0] + (sa + sd);
    const double m1 = 2.0 * sd - sa - 4.0 * f[0], m2 = 4.0 * f[0] - 2.0 * sa + sd;
    const double ax = f[1] - f[3], dx = (f[5] - f[6]) + (f[8] - f[7]);
    const double ay = f[2] - f[4], dy = (f[5] + f[6]) - (f[7] + f[8]);
    const double m3 = ax + dx, m4 = dx - 2.0 * ax, m5 = ay + dy, m6 = dy - 2.0 * ay;
    const double m7 = (f[1] + f[3]) - (f[2] + f[4]), m8 = (f[5] + f[7]) - (f[6] + f[8]);
    // Fy is stored and re-read by the next macro(): same rounding as the reference in both builds
    const double Fy = __dmul_rn(__dmul_rn(rho, gBeta), __dsub_rn(T, Tref));
    const double vF = v * Fy, uF = u * Fy;
    const double uu = u * u, vv = v * v, q3 = 3.0 * (uu + vv);
    const double p1 = m1 - Snu * (m1 - rho * (q3 - 2.0)) + (6.0 - 3.0 * Snu) * vF;
    const double p2 = m2 - Snu * (m2 - rho * (1.0 - q3)) - (6.0 - 3.0 * Snu) * vF;
    const double p4 = m4 - Sq * (m4 + rho * u);
    const double p5 = m5 + Fy;
    const double p6 = m6 - Sq * (m6 + rho * v) - (1.0 - 0.5 * Sq) * Fy;
    const double p7 = m7 - Snu * (m7 - rho * (uu - vv)) - (2.0 - Snu) * vF;
    const double p8 = m8 - Snu * (m8 - rho * (u * v)) + (1.0 - 0.5 * Snu) * uF;
    constexpr double r9 = 1.0 / 9.0, r36 = 1.0 / 36.0, r6 = 1.0 / 6.0, r12 = 1.0 / 12.0;
    fp[0] = (m0 - p1 + p2) * r9;
    const double ca = (4.0 * m0 - p1 - 2.0 * p2) * r36, cd = (4.0 * m0 + 2.0 * p1 + p2) * r36;
    const double hx = (m3 - p4) * r6, hy = (p5 - p6) * r6, h7 = 0.25 * p7;
    fp[1] = ca + hx + h7; fp[3] = ca - hx + h7;
    fp[2] = ca + hy - h7; fp[4] = ca - hy - h7;
    const double gx = (2.0 * m3 + p4) * r12, gy = (2.0 * p5 + p6) * r12, h8 = 0.25 * p8;
    fp[5] = cd + gx + gy + h8; fp[6] = cd - gx + gy - h8;
    fp[7] = cd - gx - gy + h8; fp[8] = cd + gx - gy - h8;
    Fy_out = Fy;
#endif
}

__device__ __forceinline__ void t2_collideT(const double (&g)[5], double u, double v, double T, double Qd, double Qnu, double paraA,
                                            double (&gp)[5]) {
#ifdef MGLC_STRICT
    double n[5], neq[5], np_[5];
    const double q[5] = {0.0, Qd, Qd, Qnu, Qnu};
    n[0] = g[0] + g[1] + g[2] + g[3] + g[4];
    n[1] = g[1] - g[3];
    n[2] = g[2] - g[4];
    n[3] = -4.0 * g[0] + g[1] + g[2] + g[3] + g[4];
    n[4] = g[1] - g[2] + g[3] - g[4];
    neq[0] = T;
    neq[1] = T * u;
    neq[2] = T * v;
    neq[3] = T * paraA;
    neq[4] = 0.0;
#pragma unroll
    for (int a = 0; a < 5; ++a) np_[a] = n[a] - q[a] * (n[a] - neq[a]);
    gp[0] = 0.2 * np_[0] - 0.2 * np_[3];
    gp[1] = 0.2 * np_[0] + 0.5 * np_[1] + 0.05 * np_[3] + 0.25 * np_[4];
    gp[2] = 0.2 * np_[0] + 0.5 * np_[2] + 0.05 * np_[3] - 0.25 * np_[4];
    gp[3] = 0.2 * np_[0] - 0.5 * np_[1] + 0.05 * np_[3] + 0.25 * np_[4];
    gp[4] = 0.2 * np_[0] - 0.5 * np_[2] + 0.05 * np_[3] - 0.25 * np_[4];
#else
    const double sx = g[1] + g[3], sy = g[2] + g[4];
    const double n0 = g[0] + (sx + sy), n1 = g[1] - g[3], n2 = g[2] - g[4], n3 = (sx + sy) - 4.0 * g[0], n4 = sx - sy;
    const double q1 = n1 - Qd * (n1 - T * u), q2 = n2 - Qd * (n2 - T * v);
    const double q3 = n3 - Qnu * (n3 - T * paraA), q4 = n4 - Qnu * n4;
    gp[0] = 0.2 * (n0 - q3);
    const double base = 0.2 * n0 + 0.05 * q3, h4 = 0.25 * q4;
    gp[1] = base + 0.5 * q1 + h4; gp[3] = base - 0.5 * q1 + h4;
    gp[2] = base + 0.5 * q2 - h4; gp[4] = base - 0.5 * q2 - h4;
#endif
}

// macro() + macroT(): ordered adds and IEEE divisions only -- written with explicit _rn operations so that both builds
// produce the reference's bits (Fx is the literal 0 the preceding collision() stored)
__device__ __forceinline__ void t2_macro_cell(const double (&f)[9], const double (&g)[5], double Fx, double Fy, double &rho, double &u,
                                              double &v, double &T) {
    rho = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(f[0], f[1]), f[2]), f[3]), f[4]), f[5]), f[6]), f[7]), f[8]);
    const double jx = __dadd_rn(__dsub_rn(__dsub_rn(__dadd_rn(__dsub_rn(f[1], f[3]), f[5]), f[6]), f[7]), f[8]);
    const double jy = __dsub_rn(__dsub_rn(__dadd_rn(__dadd_rn(__dsub_rn(f[2], f[4]), f[5]), f[6]), f[7]), f[8]);
    u = __ddiv_rn(__dadd_rn(jx, __dmul_rn(0.5, Fx)), rho);
    v = __ddiv_rn(__dadd_rn(jy, __dmul_rn(0.5, Fy)), rho);
    T = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(g[0], g[1]), g[2]), g[3]), g[4]);
}
