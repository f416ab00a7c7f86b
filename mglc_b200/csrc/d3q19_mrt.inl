// d3q19_mrt.inl -- per-cell D3Q19 arithmetic, included by lbm_strict.cu and lbm_fast.cu.
//
// What it computes is fixed by the reference: macro() (L3/macro.f90:13-22) and the MRT collision
// with d'Humieres' moment basis, equilibrium moments and relaxation rates of L3/collision.f90
// (:20-70 forward, :73-91 meq incl. the rho-less meq(12) at :85, :94-116 relaxation, :118-189 inverse).
//
// MGLC_STRICT build (-fmad=false): operation order, parenthesisation and true divisions follow the
// reference expression for expression, so results are bit-identical to the CPU oracle.
// Fast build (-fmad=true): same algebra, shared partial sums, constant reciprocals, FMA.
// (no include guard: tests/host_shim compiles this file twice, once per build flavour)

// rho = sum f (alpha ascending), momentum sums in alpha-ascending order, true divisions: this is
// bit-identical to the reference's accumulation loop in BOTH builds (only adds/subs and 3 divides).
__device__ __forceinline__ void d3q19_macro(const double (&f)[19], double &rho, double &u, double &v, double &w) {
    double r = f[0];
    r = __dadd_rn(r, f[1]);  r = __dadd_rn(r, f[2]);  r = __dadd_rn(r, f[3]);  r = __dadd_rn(r, f[4]);
    r = __dadd_rn(r, f[5]);  r = __dadd_rn(r, f[6]);  r = __dadd_rn(r, f[7]);  r = __dadd_rn(r, f[8]);
    r = __dadd_rn(r, f[9]);  r = __dadd_rn(r, f[10]); r = __dadd_rn(r, f[11]); r = __dadd_rn(r, f[12]);
    r = __dadd_rn(r, f[13]); r = __dadd_rn(r, f[14]); r = __dadd_rn(r, f[15]); r = __dadd_rn(r, f[16]);
    r = __dadd_rn(r, f[17]); r = __dadd_rn(r, f[18]);
    double su = f[1] - f[2] + f[7] - f[8] + f[9] - f[10] + f[11] - f[12] + f[13] - f[14];
    double sv = f[3] - f[4] + f[7] + f[8] - f[9] - f[10] + f[15] - f[16] + f[17] - f[18];
    double sw = f[5] - f[6] + f[11] + f[12] - f[13] - f[14] + f[15] + f[16] - f[17] - f[18];
    rho = r;
    u = __ddiv_rn(su, r);
    v = __ddiv_rn(sv, r);
    w = __ddiv_rn(sw, r);
}

#ifdef MGLC_STRICT
#define DIVC(x, c) ((x) / (c))

__device__ __forceinline__ void d3q19_collide(const double (&f)[19], double rho, double u, double v, double w,
                                              double Snu, double Sq, double (&fp)[19]) {
    double m[19], meq[19], mp[19];
    m[0] = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8] + f[9] + f[10] + f[11] + f[12]
         + f[13] + f[14] + f[15] + f[16] + f[17] + f[18];
    m[1] = -30.0 * f[0] - 11.0 * (f[1] + f[2] + f[3] + f[4] + f[5] + f[6])
         + 8.0 * (f[7] + f[8] + f[9] + f[10] + f[11] + f[12])
         + 8.0 * (f[13] + f[14] + f[15] + f[16] + f[17] + f[18]);
    m[2] = 12.0 * f[0] - 4.0 * (f[1] + f[2] + f[3] + f[4] + f[5] + f[6])
         + f[7] + f[8] + f[9] + f[10] + f[11] + f[12] + f[13] + f[14] + f[15] + f[16] + f[17] + f[18];
    m[3] = f[1] - f[2] + f[7] - f[8] + f[9] - f[10] + f[11] - f[12] + f[13] - f[14];
    m[4] = -4.0 * (f[1] - f[2]) + f[7] - f[8] + f[9] - f[10] + f[11] - f[12] + f[13] - f[14];
    m[5] = f[3] - f[4] + f[7] + f[8] - f[9] - f[10] + f[15] - f[16] + f[17] - f[18];
    m[6] = -4.0 * (f[3] - f[4]) + f[7] + f[8] - f[9] - f[10] + f[15] - f[16] + f[17] - f[18];
    m[7] = f[5] - f[6] + f[11] + f[12] - f[13] - f[14] + f[15] + f[16] - f[17] - f[18];
    m[8] = -4.0 * (f[5] - f[6]) + f[11] + f[12] - f[13] - f[14] + f[15] + f[16] - f[17] - f[18];
    m[9] = 2.0 * (f[1] + f[2]) - f[3] - f[4] - f[5] - f[6] + f[7] + f[8] + f[9] + f[10] + f[11] + f[12]
         + f[13] + f[14] - 2.0 * (f[15] + f[16] + f[17] + f[18]);
    m[10] = -4.0 * (f[1] + f[2]) + 2.0 * (f[3] + f[4] + f[5] + f[6]) + f[7] + f[8] + f[9] + f[10] + f[11]
          + f[12] + f[13] + f[14] - 2.0 * (f[15] + f[16] + f[17] + f[18]);
    m[11] = f[3] + f[4] - f[5] - f[6] + f[7] + f[8] + f[9] + f[10] - f[11] - f[12] - f[13] - f[14];
    m[12] = -2.0 * (f[3] + f[4] - f[5] - f[6]) + f[7] + f[8] + f[9] + f[10] - f[11] - f[12] - f[13] - f[14];
    m[13] = f[7] - f[8] - f[9] + f[10];
    m[14] = f[15] - f[16] - f[17] + f[18];
    m[15] = f[11] - f[12] - f[13] + f[14];
    m[16] = f[7] - f[8] + f[9] - f[10] - f[11] + f[12] - f[13] + f[14];
    m[17] = -f[7] - f[8] + f[9] + f[10] + f[15] - f[16] + f[17] - f[18];
    m[18] = f[11] + f[12] - f[13] - f[14] - f[15] - f[16] + f[17] + f[18];

    meq[0] = rho;
    meq[1] = rho * (-11.0 + 19.0 * (u * u + v * v + w * w));
    meq[2] = rho * (3.0 - 11.0 / 2.0 * (u * u + v * v + w * w));
    meq[3] = rho * u;
    meq[4] = -2.0 / 3.0 * rho * u;
    meq[5] = rho * v;
    meq[6] = -2.0 / 3.0 * rho * v;
    meq[7] = rho * w;
    meq[8] = -2.0 / 3.0 * rho * w;
    meq[9] = rho * (2.0 * u * u - v * v - w * w);
    meq[10] = -1.0 / 2.0 * rho * (2.0 * u * u - v * v - w * w);
    meq[11] = rho * (v * v - w * w);
    meq[12] = -1.0 / 2.0 * (v * v - w * w);          // no rho: reference quirk, L3/collision.f90:85
    meq[13] = rho * u * v;
    meq[14] = rho * v * w;
    meq[15] = rho * u * w;
    meq[16] = 0.0; meq[17] = 0.0; meq[18] = 0.0;

    const double s[19] = {0.0, Snu, Snu, 0.0, Sq, 0.0, Sq, 0.0, Sq, Snu, Snu, Snu, Snu, Snu, Snu, Snu, Sq, Sq, Sq};
#pragma unroll
    for (int a = 0; a < 19; ++a) mp[a] = m[a] - s[a] * (m[a] - meq[a]);

    fp[0] = DIVC(mp[0], 19.0) - 5.0 / 399.0 * mp[1] + DIVC(mp[2], 21.0);
    fp[1] = DIVC(mp[0], 19.0) - 11.0 / 2394.0 * mp[1] - DIVC(mp[2], 63.0) + DIVC(mp[3], 10.0) - DIVC(mp[4], 10.0) + DIVC(mp[9], 18.0) - DIVC(mp[10], 18.0);
    fp[2] = DIVC(mp[0], 19.0) - 11.0 / 2394.0 * mp[1] - DIVC(mp[2], 63.0) - DIVC(mp[3], 10.0) + DIVC(mp[4], 10.0) + DIVC(mp[9], 18.0) - DIVC(mp[10], 18.0);
    fp[3] = DIVC(mp[0], 19.0) - 11.0 / 2394.0 * mp[1] - DIVC(mp[2], 63.0) + DIVC(mp[5], 10.0) - DIVC(mp[6], 10.0) - DIVC(mp[9], 36.0) + DIVC(mp[10], 36.0) + DIVC(mp[11], 12.0) - DIVC(mp[12], 12.0);
    fp[4] = DIVC(mp[0], 19.0) - 11.0 / 2394.0 * mp[1] - DIVC(mp[2], 63.0) - DIVC(mp[5], 10.0) + DIVC(mp[6], 10.0) - DIVC(mp[9], 36.0) + DIVC(mp[10], 36.0) + DIVC(mp[11], 12.0) - DIVC(mp[12], 12.0);
    fp[5] = DIVC(mp[0], 19.0) - 11.0 / 2394.0 * mp[1] - DIVC(mp[2], 63.0) + DIVC(mp[7], 10.0) - DIVC(mp[8], 10.0) - DIVC(mp[9], 36.0) + DIVC(mp[10], 36.0) - DIVC(mp[11], 12.0) + DIVC(mp[12], 12.0);
    fp[6] = DIVC(mp[0], 19.0) - 11.0 / 2394.0 * mp[1] - DIVC(mp[2], 63.0) - DIVC(mp[7], 10.0) + DIVC(mp[8], 10.0) - DIVC(mp[9], 36.0) + DIVC(mp[10], 36.0) - DIVC(mp[11], 12.0) + DIVC(mp[12], 12.0);
    fp[7] = DIVC(mp[0], 19.0) + 4.0 / 1197.0 * mp[1] + DIVC(mp[2], 252.0) + DIVC(mp[3], 10.0) + DIVC(mp[4], 40.0) + DIVC(mp[5], 10.0) + DIVC(mp[6], 40.0) + DIVC(mp[9], 36.0) + DIVC(mp[10], 72.0) + DIVC(mp[11], 12.0) + DIVC(mp[12], 24.0) + DIVC(mp[13], 4.0) + DIVC(mp[16], 8.0) - DIVC(mp[17], 8.0);
    fp[8] = DIVC(mp[0], 19.0) + 4.0 / 1197.0 * mp[1] + DIVC(mp[2], 252.0) - DIVC(mp[3], 10.0) - DIVC(mp[4], 40.0) + DIVC(mp[5], 10.0) + DIVC(mp[6], 40.0) + DIVC(mp[9], 36.0) + DIVC(mp[10], 72.0) + DIVC(mp[11], 12.0) + DIVC(mp[12], 24.0) - DIVC(mp[13], 4.0) - DIVC(mp[16], 8.0) - DIVC(mp[17], 8.0);
    fp[9] = DIVC(mp[0], 19.0) + 4.0 / 1197.0 * mp[1] + DIVC(mp[2], 252.0) + DIVC(mp[3], 10.0) + DIVC(mp[4], 40.0) - DIVC(mp[5], 10.0) - DIVC(mp[6], 40.0) + DIVC(mp[9], 36.0) + DIVC(mp[10], 72.0) + DIVC(mp[11], 12.0) + DIVC(mp[12], 24.0) - DIVC(mp[13], 4.0) + DIVC(mp[16], 8.0) + DIVC(mp[17], 8.0);
    fp[10] = DIVC(mp[0], 19.0) + 4.0 / 1197.0 * mp[1] + DIVC(mp[2], 252.0) - DIVC(mp[3], 10.0) - DIVC(mp[4], 40.0) - DIVC(mp[5], 10.0) - DIVC(mp[6], 40.0) + DIVC(mp[9], 36.0) + DIVC(mp[10], 72.0) + DIVC(mp[11], 12.0) + DIVC(mp[12], 24.0) + DIVC(mp[13], 4.0) - DIVC(mp[16], 8.0) + DIVC(mp[17], 8.0);
    fp[11] = DIVC(mp[0], 19.0) + 4.0 / 1197.0 * mp[1] + DIVC(mp[2], 252.0) + DIVC(mp[3], 10.0) + DIVC(mp[4], 40.0) + DIVC(mp[7], 10.0) + DIVC(mp[8], 40.0) + DIVC(mp[9], 36.0) + DIVC(mp[10], 72.0) - DIVC(mp[11], 12.0) - DIVC(mp[12], 24.0) + DIVC(mp[15], 4.0) - DIVC(mp[16], 8.0) + DIVC(mp[18], 8.0);
    fp[12] = DIVC(mp[0], 19.0) + 4.0 / 1197.0 * mp[1] + DIVC(mp[2], 252.0) - DIVC(mp[3], 10.0) - DIVC(mp[4], 40.0) + DIVC(mp[7], 10.0) + DIVC(mp[8], 40.0) + DIVC(mp[9], 36.0) + DIVC(mp[10], 72.0) - DIVC(mp[11], 12.0) - DIVC(mp[12], 24.0) - DIVC(mp[15], 4.0) + DIVC(mp[16], 8.0) + DIVC(mp[18], 8.0);
    fp[13] = DIVC(mp[0], 19.0) + 4.0 / 1197.0 * mp[1] + DIVC(mp[2], 252.0) + DIVC(mp[3], 10.0) + DIVC(mp[4], 40.0) - DIVC(mp[7], 10.0) - DIVC(mp[8], 40.0) + DIVC(mp[9], 36.0) + DIVC(mp[10], 72.0) - DIVC(mp[11], 12.0) - DIVC(mp[12], 24.0) - DIVC(mp[15], 4.0) - DIVC(mp[16], 8.0) - DIVC(mp[18], 8.0);
    fp[14] = DIVC(mp[0], 19.0) + 4.0 / 1197.0 * mp[1] + DIVC(mp[2], 252.0) - DIVC(mp[3], 10.0) - DIVC(mp[4], 40.0) - DIVC(mp[7], 10.0) - DIVC(mp[8], 40.0) + DIVC(mp[9], 36.0) + DIVC(mp[10], 72.0) - DIVC(mp[11], 12.0) - DIVC(mp[12], 24.0) + DIVC(mp[15], 4.0) + DIVC(mp[16], 8.0) - DIVC(mp[18], 8.0);
    fp[15] = DIVC(mp[0], 19.0) + 4.0 / 1197.0 * mp[1] + DIVC(mp[2], 252.0) + DIVC(mp[5], 10.0) + DIVC(mp[6], 40.0) + DIVC(mp[7], 10.0) + DIVC(mp[8], 40.0) - DIVC(mp[9], 18.0) - DIVC(mp[10], 36.0) + DIVC(mp[14], 4.0) + DIVC(mp[17], 8.0) - DIVC(mp[18], 8.0);
    fp[16] = DIVC(mp[0], 19.0) + 4.0 / 1197.0 * mp[1] + DIVC(mp[2], 252.0) - DIVC(mp[5], 10.0) - DIVC(mp[6], 40.0) + DIVC(mp[7], 10.0) + DIVC(mp[8], 40.0) - DIVC(mp[9], 18.0) - DIVC(mp[10], 36.0) - DIVC(mp[14], 4.0) - DIVC(mp[17], 8.0) - DIVC(mp[18], 8.0);
    fp[17] = DIVC(mp[0], 19.0) + 4.0 / 1197.0 * mp[1] + DIVC(mp[2], 252.0) + DIVC(mp[5], 10.0) + DIVC(mp[6], 40.0) - DIVC(mp[7], 10.0) - DIVC(mp[8], 40.0) - DIVC(mp[9], 18.0) - DIVC(mp[10], 36.0) - DIVC(mp[14], 4.0) + DIVC(mp[17], 8.0) + DIVC(mp[18], 8.0);
    fp[18] = DIVC(mp[0], 19.0) + 4.0 / 1197.0 * mp[1] + DIVC(mp[2], 252.0) - DIVC(mp[5], 10.0) - DIVC(mp[6], 40.0) - DIVC(mp[7], 10.0) - DIVC(mp[8], 40.0) - DIVC(mp[9], 18.0) - DIVC(mp[10], 36.0) + DIVC(mp[14], 4.0) - DIVC(mp[17], 8.0) + DIVC(mp[18], 8.0);
}
#undef DIVC

#else  // ---------------------------------------- fast build -----------------------------------------

// Same operator, restructured for the fp64 pipe (64 lanes/clk/SM): the forward transform shares the
// partial sums every d'Humieres row is built from, relaxation is applied to the non-equilibrium part,
// and the inverse transform (M^-1 = M^T diag(1/|row|^2)) shares the scaled moments between the
// populations that use them.  All divisions by constants are folded into compile-time reciprocals.
__device__ __forceinline__ void d3q19_collide(const double (&f)[19], double rho, double u, double v, double w,
                                              double Snu, double Sq, double (&fp)[19]) {
    // pair sums / differences along each axis and each diagonal plane
    const double a12 = f[1] + f[2], d12 = f[1] - f[2];
    const double a34 = f[3] + f[4], d34 = f[3] - f[4];
    const double a56 = f[5] + f[6], d56 = f[5] - f[6];
    // xy plane: 7(+,+) 8(-,+) 9(+,-) 10(-,-)
    const double sxy = (f[7] + f[8]) + (f[9] + f[10]);
    const double xy_x = (f[7] - f[8]) + (f[9] - f[10]);       // sum ex*f
    const double xy_y = (f[7] + f[8]) - (f[9] + f[10]);       // sum ey*f
    const double xy_c = (f[7] - f[8]) - (f[9] - f[10]);       // sum ex*ey*f
    // xz plane: 11(+,+) 12(-,+) 13(+,-) 14(-,-)
    const double sxz = (f[11] + f[12]) + (f[13] + f[14]);
    const double xz_x = (f[11] - f[12]) + (f[13] - f[14]);
    const double xz_z = (f[11] + f[12]) - (f[13] + f[14]);
    const double xz_c = (f[11] - f[12]) - (f[13] - f[14]);
    // yz plane: 15(+,+) 16(-,+) 17(+,-) 18(-,-)
    const double syz = (f[15] + f[16]) + (f[17] + f[18]);
    const double yz_y = (f[15] - f[16]) + (f[17] - f[18]);
    const double yz_z = (f[15] + f[16]) - (f[17] + f[18]);
    const double yz_c = (f[15] - f[16]) - (f[17] - f[18]);

    const double ax = a12 + a34 + a56;          // axis populations
    const double dg = sxy + sxz + syz;          // diagonal populations
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
    const double m16 = xy_x - xz_x;
    const double m17 = yz_y - xy_y;
    const double m18 = xz_z - yz_z;

    const double uu = u * u, vv = v * v, ww = w * w;
    const double us2 = uu + vv + ww;
    // post-collision moments, each already divided by |row|^2 (M^-1 = M^T diag(1/|row|^2))
    const double q0 = m0 * (1.0 / 19.0);                                                  // s = 0
    const double q1 = (m1 - Snu * (m1 - rho * (-11.0 + 19.0 * us2))) * (1.0 / 2394.0);
    const double q2 = (m2 - Snu * (m2 - rho * (3.0 - 5.5 * us2))) * (1.0 / 252.0);
    const double q3 = m3 * 0.1, q5 = m5 * 0.1, q7 = m7 * 0.1;                             // s = 0
    const double q4 = (m4 - Sq * (m4 + (2.0 / 3.0) * rho * u)) * (1.0 / 40.0);
    const double q6 = (m6 - Sq * (m6 + (2.0 / 3.0) * rho * v)) * (1.0 / 40.0);
    const double q8 = (m8 - Sq * (m8 + (2.0 / 3.0) * rho * w)) * (1.0 / 40.0);
    const double e9 = rho * (2.0 * uu - vv - ww);
    const double q9 = (m9 - Snu * (m9 - e9)) * (1.0 / 36.0);
    const double q10 = (m10 - Snu * (m10 + 0.5 * e9)) * (1.0 / 72.0);
    const double e11 = vv - ww;
    const double q11 = (m11 - Snu * (m11 - rho * e11)) * (1.0 / 12.0);
    const double q12 = (m12 - Snu * (m12 + 0.5 * e11)) * (1.0 / 24.0);                    // quirk: no rho
    const double q13 = (m13 - Snu * (m13 - rho * u * v)) * 0.25;
    const double q14 = (m14 - Snu * (m14 - rho * v * w)) * 0.25;
    const double q15 = (m15 - Snu * (m15 - rho * u * w)) * 0.25;
    const double q16 = (m16 - Sq * m16) * 0.125;
    const double q17 = (m17 - Sq * m17) * 0.125;
    const double q18 = (m18 - Sq * m18) * 0.125;

    // f = M^T q : column alpha of M dotted with q
    fp[0] = q0 - 30.0 * q1 + 12.0 * q2;
    const double cax = q0 - 11.0 * q1 - 4.0 * q2;            // common to the 6 axis populations
    const double cdg = q0 + 8.0 * q1 + q2;                   // common to the 12 diagonal populations
    const double ax_x = cax + 2.0 * (q9 - 2.0 * q10);        // x-axis pair: +2 q9 - 4 q10
    const double ax_yz = cax - (q9 - 2.0 * q10);             // y,z pairs: -q9 + 2 q10
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
    const double cxy = cdg + p9 + p11;                       // xy-plane diagonals
    const double cxz = cdg + p9 - p11;                       // xz-plane diagonals
    const double cyz = cdg - 2.0 * p9;                       // yz-plane diagonals
    // 7(+,+,0) 8(-,+,0) 9(+,-,0) 10(-,-,0): ex*(kx + q16), ey*(ky - q17), ex*ey*q13
    fp[7]  = cxy + (kx + q16) + (ky - q17) + q13;
    fp[8]  = cxy - (kx + q16) + (ky - q17) - q13;
    fp[9]  = cxy + (kx + q16) - (ky - q17) - q13;
    fp[10] = cxy - (kx + q16) - (ky - q17) + q13;
    // 11(+,0,+) 12(-,0,+) 13(+,0,-) 14(-,0,-): ex*(kx - q16), ez*(kz + q18), ex*ez*q15
    fp[11] = cxz + (kx - q16) + (kz + q18) + q15;
    fp[12] = cxz - (kx - q16) + (kz + q18) - q15;
    fp[13] = cxz + (kx - q16) - (kz + q18) - q15;
    fp[14] = cxz - (kx - q16) - (kz + q18) + q15;
    // 15(0,+,+) 16(0,-,+) 17(0,+,-) 18(0,-,-): ey*(ky + q17), ez*(kz - q18), ey*ez*q14
    fp[15] = cyz + (ky + q17) + (kz - q18) + q14;
    fp[16] = cyz - (ky + q17) + (kz - q18) - q14;
    fp[17] = cyz + (ky + q17) - (kz - q18) - q14;
    fp[18] = cyz - (ky + q17) - (kz - q18) + q14;
}
#endif

// The single-relaxation-time alternative the reference keeps as a comment block at the end of collision()'s
// cell loop (L3/collision.f90:191-198): feq as in initial() (L3/initial.f90:63-73), every population relaxed at
// Snu.  Strict build: the reference's expression order (un = u*ex + v*ey + w*ez with the integer components
// promoted, rho*omega*(...) left to right).  Fast build: shared 1 - 1.5 us2 and rho*omega, FMA.
__device__ __forceinline__ void d3q19_collide_bgk(const double (&f)[19], double rho, double u, double v, double w,
                                                  double Snu, double (&fp)[19]) {
    constexpr double ex[19] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
    constexpr double ey[19] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
    constexpr double ez[19] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
    const double us2 = u * u + v * v + w * w;
#ifdef MGLC_STRICT
#pragma unroll
    for (int a = 0; a < 19; ++a) {
        const double om = a == 0 ? 1.0 / 3.0 : (a < 7 ? 1.0 / 18.0 : 1.0 / 36.0);
        const double un = u * ex[a] + v * ey[a] + w * ez[a];
        const double feq = rho * om * (1.0 + 3.0 * un + 4.5 * un * un - 1.5 * us2);
        fp[a] = f[a] - Snu * (f[a] - feq);
    }
#else
    const double base = 1.0 - 1.5 * us2;
    const double r0 = rho * (1.0 / 3.0), r1 = rho * (1.0 / 18.0), r2 = rho * (1.0 / 36.0);
#pragma unroll
    for (int a = 0; a < 19; ++a) {
        const double un = ex[a] * u + ey[a] * v + ez[a] * w;        // folds to +-u +-v at compile time
        const double feq = (a == 0 ? r0 : (a < 7 ? r1 : r2)) * (base + un * (3.0 + 4.5 * un));
        fp[a] = f[a] + Snu * (feq - f[a]);
    }
#endif
}

