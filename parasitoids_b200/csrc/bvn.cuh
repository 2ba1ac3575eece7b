// Bivariate-normal rectangle probabilities (Genz BVU) as device functions.
//
// Replaces scipy.stats.mvn.mvnun (Fortran MVNDST -> BVNMVN -> BVU, A. Genz)
// at its four call sites ParasitoidModel.py:340,356,366,370.  The algorithm is
// restated from the published method (Genz 2004, TVPACK BVU): Gauss-Legendre
// quadrature of the Drezner-Wesolowsky integrand for |rho| < 0.925, and the
// expansion around |rho| = 1 otherwise.  Everything that depends only on the
// covariance (asin(rho), the sines at the quadrature nodes, the reciprocal of
// 1 - sn^2) is precomputed once per covariance into BvnPar, so a lattice point
// costs 2*lg exp() evaluations and no sin().
#pragma once
#include "pkb_platform.cuh"

namespace pkb {

struct BvnPar {
    double sx, sy, rho;
    int lg;     // half-rule size: 3, 6 or 10
    int high;   // |rho| >= 0.925
    int h0;     // support half-width of get_mvn_cdf_values for mu = 0 (filled by k_bvn_setup)
    int pad_;
    double asr4pi;                    // asin(rho) / (4 pi)
    double sn[20], inv[20], w[20];    // low branch: sin at node, 1/(1-sn^2), weight
    double as_, a, ahalf;             // high branch: (1-r)(1+r), sqrt, sqrt/2
    double xs[20], rs[20];            // high branch: (a/2 (1 +- x))^2, sqrt(1 - xs)
};

__device__ __forceinline__ double phid(double z) { return 0.5 * erfc(-z * 0.70710678118654752440); }

__device__ const double kGLw3[3] = {0.1713244923791705, 0.3607615730481384, 0.4679139345726904};
__device__ const double kGLx3[3] = {0.9324695142031522, 0.6612093864662647, 0.2386191860831970};
__device__ const double kGLw6[6] = {0.04717533638651177, 0.1069393259953183, 0.1600783285433464,
                                    0.2031674267230659,  0.2334925365383547, 0.2491470458134029};
__device__ const double kGLx6[6] = {0.9815606342467191, 0.9041172563704750, 0.7699026741943050,
                                    0.5873179542866171, 0.3678314989981802, 0.1252334085114692};
__device__ const double kGLw10[10] = {0.01761400713915212, 0.04060142980038694, 0.06267204833410906,
                                      0.08327674157670475, 0.1019301198172404,  0.1181945319615184,
                                      0.1316886384491766,  0.1420961093183821,  0.1491729864726037,
                                      0.1527533871307259};
__device__ const double kGLx10[10] = {0.9931285991850949, 0.9639719272779138, 0.9122344282513259,
                                      0.8391169718222188, 0.7463319064601508, 0.6360536807265150,
                                      0.5108670019508271, 0.3737060887154196, 0.2277858511416451,
                                      0.07652652113349733};

__device__ __forceinline__ void gl_rule(int lg, const double*& w, const double*& x) {
    if (lg == 3) { w = kGLw3; x = kGLx3; }
    else if (lg == 6) { w = kGLw6; x = kGLx6; }
    else { w = kGLw10; x = kGLx10; }
}

// Fill the covariance-only constants.  S = [[sx^2, rho sx sy], [rho sx sy, sy^2]]
// is formed exactly like Dmat (ParasitoidModel.py:279-280) and then
// standardised like mvnun does, so rho passes through the same roundings.
__device__ inline void bvn_setup_cov(BvnPar& p, double s00, double s11, double s01);
__device__ inline void bvn_setup(BvnPar& p, double sig_x, double sig_y, double rho_in) {
    bvn_setup_cov(p, sig_x * sig_x, sig_y * sig_y, rho_in * sig_x * sig_y);
}
__device__ inline void bvn_setup_cov(BvnPar& p, double s00, double s11, double s01) {
    p.sx = sqrt(s00);
    p.sy = sqrt(s11);
    p.rho = s01 / (p.sx * p.sy);
    const double r = p.rho, ar = fabs(r);
    p.lg = ar < 0.3 ? 3 : (ar < 0.75 ? 6 : 10);
    p.high = ar < 0.925 ? 0 : 1;
    p.h0 = 0;
    p.pad_ = 0;
    const double *w, *x;
    gl_rule(p.lg, w, x);
    for (int i = 0; i < 20; ++i) { p.sn[i] = 0; p.inv[i] = 0; p.w[i] = 0; p.xs[i] = 0; p.rs[i] = 0; }
    p.as_ = p.a = p.ahalf = 0;
    p.asr4pi = 0;
    if (!p.high) {
        const double asr = asin(r);
        p.asr4pi = asr / (4.0 * 3.14159265358979323846);
        for (int i = 0; i < p.lg; ++i) {
            for (int s = 0; s < 2; ++s) {
                const double sg = s ? 1.0 : -1.0;
                const double sn = sin(asr * (1.0 + sg * x[i]) / 2.0);
                p.sn[2 * i + s] = sn;
                p.inv[2 * i + s] = 1.0 / (1.0 - sn * sn);
                p.w[2 * i + s] = w[i];
            }
        }
    } else if (ar < 1.0) {
        p.as_ = (1.0 - r) * (1.0 + r);
        p.a = sqrt(p.as_);
        p.ahalf = p.a / 2.0;
        for (int i = 0; i < p.lg; ++i) {
            for (int s = 0; s < 2; ++s) {
                const double sg = s ? 1.0 : -1.0;
                const double t = p.ahalf * (1.0 + sg * x[i]);
                p.xs[2 * i + s] = t * t;
                p.rs[2 * i + s] = sqrt(1.0 - t * t);
                p.w[2 * i + s] = w[i];
            }
        }
    }
}

// |rho| < 0.925 with the 1-D factors supplied by the caller:
//   hk = h*k, hs = (h^2 + k^2)/2, pp = Phi(-h) * Phi(-k)
__device__ __forceinline__ double bvu_low_core(const BvnPar& p, double hk, double hs, double pp) {
    double acc = 0.0;
    const int n = 2 * p.lg;
    for (int i = 0; i < n; ++i) {
        const double arg = fma(p.sn[i], hk, -hs) * p.inv[i];
        acc = fma(p.w[i], exp(arg), acc);
    }
    return fma(acc, p.asr4pi, pp);
}

// |rho| >= 0.925 (TVPACK BVU, second branch)
__device__ inline double bvu_high(const BvnPar& p, double h, double k) {
    const double r = p.rho;
    const double twopi = 6.283185307179586476925;
    double hk = h * k;
    if (r < 0) { k = -k; hk = -hk; }
    double bvn = 0.0;
    if (fabs(r) < 1.0) {
        const double as_ = p.as_, a = p.a;
        const double bs = (h - k) * (h - k);
        const double c = (4.0 - hk) / 8.0;
        const double d = (12.0 - hk) / 16.0;
        double asr = -(bs / as_ + hk) / 2.0;
        if (asr > -100.0)
            bvn = a * exp(asr) * (1.0 - c * (bs - as_) * (1.0 - d * bs / 5.0) / 3.0 + c * d * as_ * as_ / 5.0);
        if (-hk < 100.0) {
            const double b = sqrt(bs);
            bvn -= exp(-hk / 2.0) * sqrt(twopi) * phid(-b / a) * b * (1.0 - c * bs * (1.0 - d * bs / 5.0) / 3.0);
        }
        const double ah = p.ahalf;
        const int n = 2 * p.lg;
        for (int i = 0; i < n; ++i) {
            const double xs = p.xs[i], rs = p.rs[i];
            asr = -(bs / xs + hk) / 2.0;
            if (asr > -100.0) {
                const double t = 1.0 + rs;
                bvn += ah * p.w[i] * exp(asr) * (exp(-hk * xs / (2.0 * t * t)) / rs - (1.0 + c * xs * (1.0 + d * xs)));
            }
        }
        bvn = -bvn / twopi;
    }
    if (r > 0) {
        bvn += phid(-fmax(h, k));
    } else {
        bvn = -bvn;
        if (k > h) {
            if (h < 0) bvn += phid(k) - phid(h);
            else bvn += phid(-h) - phid(-k);
        }
    }
    return bvn;
}

// P(X > h, Y > k), standard bivariate normal with correlation p.rho
__device__ __forceinline__ double bvu(const BvnPar& p, double h, double k) {
    if (p.high) return bvu_high(p, h, k);
    return bvu_low_core(p, h * k, (h * h + k * k) / 2.0, phid(-h) * phid(-k));
}

// mvnun(low, upp, mu, S)[0] for one rectangle (limits in metres)
__device__ inline double mvn_rect(const BvnPar& p, double xl, double xu, double yl, double yu, double mux, double muy) {
    const double a0 = (xl - mux) / p.sx, a1 = (xu - mux) / p.sx;
    const double b0 = (yl - muy) / p.sy, b1 = (yu - muy) / p.sy;
    return bvu(p, a0, b0) - bvu(p, a1, b0) - bvu(p, a0, b1) + bvu(p, a1, b1);
}

// Probability of the (2h+1)-cell square that get_mvn_cdf_values sums at ring h
// (ParasitoidModel.py:338-339,354-355: low = -h*c - r, upp = (h*c - r) + c).
__device__ inline double square_prob(const BvnPar& p, double cell, int h, double mux, double muy) {
    const double r = cell / 2;
    const double lo = (-h) * cell - r;
    const double up = (h * cell - r) + cell;
    return mvn_rect(p, lo, up, lo, up, mux, muy);
}

// Support half-width by the reference's own running sum (ParasitoidModel.py:338-373): centre cell, then per ring
// the four corners and the four sides in call order, one mvnun per cell, until 1 - val_sum < cdf_eps.  Used only
// where the one-rectangle form above lands within `ring_tol` of cdf_eps (its value agrees with the running sum to
// ~1e-15, so only there could the decision differ); single thread, (2 hmax + 1)^2 rectangles at most.
__device__ inline int ring_halfwidth_ref_order(const BvnPar& p, double cell, double mux, double muy, double cdf_eps, int hmax) {
    const double r = cell / 2;
    double val_sum = mvn_rect(p, -r, -r + cell, -r, -r + cell, mux, muy);
    int h = 0;
    while (1.0 - val_sum >= cdf_eps && h < hmax) {
        ++h;
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) {
                const int ii = a ? h : -h, jj = b ? h : -h;
                const double xl = ii * cell - r, yl = jj * cell - r;
                val_sum += mvn_rect(p, xl, xl + cell, yl, yl + cell, mux, muy);
            }
        for (int a = 0; a < 2; ++a) {
            const int ii = a ? h : -h;
            for (int jj = -h + 1; jj < h; ++jj) {
                const double l0 = ii * cell - r, l1 = jj * cell - r;
                val_sum += mvn_rect(p, l0, l0 + cell, l1, l1 + cell, mux, muy);     // cell (ii, jj)
                val_sum += mvn_rect(p, l1, l1 + cell, l0, l0 + cell, mux, muy);     // cell (jj, ii)
            }
        }
    }
    return h;
}

}  // namespace pkb
