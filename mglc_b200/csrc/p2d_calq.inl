// p2d_calq.inl -- calQ of the particle path, shared by particles.cu and the host shim of the CPU test suite
// (tests/host_shim/calq_host.cpp compiles it with gcc and checks it against the oracle bit for bit).
// The caller defines ERR_CALQ / ERR_Q.
// calQ, P4/particle_bounceback.F90:98-141: bisection along link alpha to |dist - radius| < 1e-9 (single literal).
// The reference takes two square roots per halving.  Here a halving whose squared distance D differs from radius^2 by
// more than epsRadius * (2 radius + 2) skips them: with dist <= radius + sqrt(2) on the link of a boundary node,
// |dist - radius| = |D - radius^2| / (dist + radius) then exceeds epsRadius, so the reference's loop test cannot pass
// and the side of the surface is the sign of D - radius^2.  Near the surface (the last one to three halvings) the
// reference's expressions run verbatim.  x0, y0 and q are sums of dyadic fractions (exact in fp64), so the result is
// bit-identical to the reference's for every input.
__device__ inline int calQ_link(double xc, double yc, double rad, double i, double j, double exa, double eya, double &x0, double &y0, double &q) {
    const double epsRadius = (double)1e-9f;
    const double rad2 = rad * rad, far = epsRadius * (2.0 * rad + 2.0) * 1.001;
    q = 0.5;
    double qTemp = 0.5;
    x0 = i + qTemp * exa; y0 = j + qTemp * eya;
    for (;;) {
        const double D = (x0 - xc) * (x0 - xc) + (y0 - yc) * (y0 - yc);
        bool outside;
        if (fabs(D - rad2) > far) outside = D > rad2;
        else {
            const double d = sqrt(D);
            if (!(fabs(d - rad) >= epsRadius)) break;
            if (d > rad) outside = true;
            else if (d < rad) outside = false;
            else return ERR_CALQ;
        }
        qTemp = qTemp / 2.0;
        if (outside) { x0 = x0 + qTemp * exa; y0 = y0 + qTemp * eya; q = q + qTemp; }
        else { x0 = x0 - qTemp * exa; y0 = y0 - qTemp * eya; q = q - qTemp; }
        if (qTemp == 0.0) return ERR_CALQ;
    }
    return (q > 1.0 || q < 0.0) ? ERR_Q : 0;
}

