// Deterministic double sin / cos from IEEE-exact operations only (fdlibm's published reduction and kernels): the same arithmetic as oracle/detmath.h, so the
// rectangle axes are bit-identical to the CPU restatement; within 1 ulp of any libm.  Plain double operations, shared by the device code (lsd_kernels.cuh) and
// the host (lsd_pipeline.cu builds the seed (cos, sin) table with it; tests/host_harness/lsd_rectenum_host.cc checks it against the oracle on the CPU).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define LSD_SC_HD __host__ __device__ __forceinline__
#else
#define LSD_SC_HD inline
#endif

LSD_SC_HD double lsd_ksin(double x, double y, int iy) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double z = x * x, v = z * x;
    const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    if (iy == 0) return x + v * (S1 + z * r);
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}
LSD_SC_HD double lsd_kcos(double x, double y) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double z = x * x;
    const double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    const double ax = fabs(x);
    if (ax < 0.3) return 1.0 - (0.5 * z - (z * r - x * y));
    const double qx = ax > 0.78125 ? 0.28125 : floor(ax * 32.0) / 128.0;
    const double hz = 0.5 * z - qx, a = 1.0 - qx;
    return a - (hz - (z * r - x * y));
}
// valid for |x| < ~1e5 (the detector's angles are below 4 pi)
LSD_SC_HD void lsd_sincos_body(double x, double& s, double& c) {
    const double pio2_1 = 1.57079632673412561417e+00, pio2_1t = 6.07710050650619224932e-11;
    const double pio2_2 = 6.07710050630396597660e-11, pio2_2t = 2.02226624879595063154e-21;
    const double fn = rint(x * 6.36619772367581382433e-01);
    const int n = (int)fn;
    double r = x - fn * pio2_1, w = fn * pio2_1t;
    double y0 = r - w;
    if (y0 == 0.0 || ilogb(x) - ilogb(y0) > 16) {
        const double t = r;
        w = fn * pio2_2;
        r = t - w;
        w = fn * pio2_2t - ((t - r) - w);
        y0 = r - w;
    }
    const double y1 = (r - y0) - w;
    const double ks = lsd_ksin(y0, y1, 1), kc = lsd_kcos(y0, y1);
    switch (n & 3) {
        case 0: s = ks; c = kc; break;
        case 1: s = kc; c = -ks; break;
        case 2: s = -ks; c = -kc; break;
        default: s = -kc; c = ks; break;
    }
}
