// Shared-memory mixed-radix complex128 FFT (radices 2,3,4,5,7,8), sm_100a.
//
// Replaces the pocketfft/ducc complex FFT the reference reaches through
// scipy.fftpack (CalcSol.py:24,35,65,99) and the Reikna FFT of cuda_lib.py:42-54.
//
// Formulation: decimation-in-frequency, in place, natural-order input ->
// digit-reversed output (fft_dif); the inverse is the exact transpose,
// decimation-in-time, digit-reversed input -> natural-order output (fft_dit_inv).
// Pointwise products are taken in the permuted domain, so the convolution
// pipeline never un-permutes; only the two-real-rows pack/unpack steps look up
// `perm` to pair bin k with bin N-k.
//
// One transform lives in one shared-memory buffer of Npad = roundup(N, 64)
// complex128, addressed through an XOR swizzle so that the power-of-two strides
// of the late stages do not serialise on the 16-byte bank groups.
#pragma once
#include "pkb_platform.cuh"

namespace pkb {

#define PKB_FFT_MAX_STAGES 16

struct FftPlan {
    int N;
    int Npad;
    int nstage;
    int radix[PKB_FFT_MAX_STAGES];
    const cplx* tw;   // device: tw[j] = exp(-2 pi i j / N), j in [0, N)
    const int* perm;  // device: perm[k] = position of frequency k after fft_dif
};

__host__ __device__ __forceinline__ int swz(int i) { return i ^ ((i >> 3) & 7); }

// ---- small DFTs (forward sign: exp(-i...)) -----------------------------------
__device__ __forceinline__ void dft2(cplx& a, cplx& b) {
    cplx t = csub(a, b);
    a = cadd(a, b);
    b = t;
}
// multiply by -i
__device__ __forceinline__ cplx mul_mi(cplx a) { return cmake(a.y, -a.x); }

__device__ __forceinline__ void dft4(cplx& x0, cplx& x1, cplx& x2, cplx& x3) {
    cplx t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = mul_mi(csub(x1, x3));
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    x1 = cadd(t1, t3);
    x3 = csub(t1, t3);
}

template <int R> struct OddTab;
template <> struct OddTab<3> {
    __device__ static __forceinline__ double c(int i) { const double t[2] = {-0.5, -0.5}; return t[i]; }
    __device__ static __forceinline__ double s(int i) {
        const double t[2] = {0.8660254037844386467637, -0.8660254037844386467637};
        return t[i];
    }
};
template <> struct OddTab<5> {
    __device__ static __forceinline__ double c(int i) {
        const double t[4] = {0.3090169943749474241023, -0.8090169943749474241023, -0.8090169943749474241023,
                             0.3090169943749474241023};
        return t[i];
    }
    __device__ static __forceinline__ double s(int i) {
        const double t[4] = {0.9510565162951535721164, 0.5877852522924731291687, -0.5877852522924731291687,
                             -0.9510565162951535721164};
        return t[i];
    }
};
template <> struct OddTab<7> {
    __device__ static __forceinline__ double c(int i) {
        const double t[6] = {0.623489801858733530525,   -0.2225209339563144042889, -0.9009688679024191262361,
                             -0.9009688679024191262361, -0.2225209339563144042889, 0.623489801858733530525};
        return t[i];
    }
    __device__ static __forceinline__ double s(int i) {
        const double t[6] = {0.7818314824680298087084,  0.9749279121818236070181,  0.4338837391175581204758,
                             -0.4338837391175581204758, -0.9749279121818236070181, -0.7818314824680298087084};
        return t[i];
    }
};

// odd-length DFT via the symmetric/antisymmetric pairing x_m +- x_{R-m}
template <int R>
__device__ __forceinline__ void dft_odd(cplx (&v)[R]) {
    constexpr int H = (R - 1) / 2;
    cplx a[H], b[H];
#pragma unroll
    for (int m = 1; m <= H; ++m) {
        a[m - 1] = cadd(v[m], v[R - m]);
        b[m - 1] = csub(v[m], v[R - m]);
    }
    cplx x0 = v[0];
    cplx s0 = x0;
#pragma unroll
    for (int m = 0; m < H; ++m) s0 = cadd(s0, a[m]);
    v[0] = s0;
#pragma unroll
    for (int q = 1; q <= H; ++q) {
        cplx C = x0, S = cmake(0.0, 0.0);
#pragma unroll
        for (int m = 1; m <= H; ++m) {
            const int idx = (q * m) % R - 1;  // compile-time after unrolling
            const double cc = OddTab<R>::c(idx), ss = OddTab<R>::s(idx);
            C.x = fma(cc, a[m - 1].x, C.x);
            C.y = fma(cc, a[m - 1].y, C.y);
            S.x = fma(ss, b[m - 1].x, S.x);
            S.y = fma(ss, b[m - 1].y, S.y);
        }
        v[q] = cmake(C.x + S.y, C.y - S.x);      // C - i S
        v[R - q] = cmake(C.x - S.y, C.y + S.x);  // C + i S
    }
}

template <int R>
__device__ __forceinline__ void dft_fwd(cplx (&v)[R]);
template <> __device__ __forceinline__ void dft_fwd<2>(cplx (&v)[2]) { dft2(v[0], v[1]); }
template <> __device__ __forceinline__ void dft_fwd<3>(cplx (&v)[3]) { dft_odd<3>(v); }
template <> __device__ __forceinline__ void dft_fwd<4>(cplx (&v)[4]) { dft4(v[0], v[1], v[2], v[3]); }
template <> __device__ __forceinline__ void dft_fwd<5>(cplx (&v)[5]) { dft_odd<5>(v); }
template <> __device__ __forceinline__ void dft_fwd<7>(cplx (&v)[7]) { dft_odd<7>(v); }
template <> __device__ __forceinline__ void dft_fwd<8>(cplx (&v)[8]) {
    const double h = 0.7071067811865475244008;
    dft4(v[0], v[2], v[4], v[6]);
    dft4(v[1], v[3], v[5], v[7]);
    // odd outputs times w8^k, k = 0..3
    cplx o1 = cmake(h * (v[3].x + v[3].y), h * (v[3].y - v[3].x));   // * (h, -h)
    cplx o2 = mul_mi(v[5]);                                          // * -i
    cplx o3 = cmake(h * (v[7].y - v[7].x), -h * (v[7].x + v[7].y));  // * (-h, -h)
    cplx e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
    v[0] = cadd(e0, o0);
    v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1);
    v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2);
    v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3);
    v[7] = csub(e3, o3);
}

template <int R>
__device__ __forceinline__ void swap_reim(cplx (&v)[R]) {
#pragma unroll
    for (int q = 0; q < R; ++q) {
        double t = v[q].x;
        v[q].x = v[q].y;
        v[q].y = t;
    }
}

// ---- one in-place stage over a block decomposition of size M -----------------
// INV = false: DIF forward stage (DFT_R then twiddle on outputs)
// INV = true : DIT inverse stage (conj twiddle on inputs then inverse DFT_R)
template <int R, bool INV>
__device__ __forceinline__ void fft_stage(cplx* x, int N, int M, const cplx* __restrict__ tw, int tid, int T) {
    const int Ms = M / R;
    const int nb = N / R;
    const int tscale = N / M;
    for (int j = tid; j < nb; j += T) {
        const int b = j / Ms;
        const int k = j - b * Ms;
        const int base = b * M + k;
        cplx v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = x[swz(base + q * Ms)];
        if (!INV) {
            dft_fwd<R>(v);
            if (k > 0) {
#pragma unroll
                for (int q = 1; q < R; ++q) v[q] = cmul(v[q], __ldg(&tw[k * q * tscale]));
            }
        } else {
            if (k > 0) {
#pragma unroll
                for (int q = 1; q < R; ++q) v[q] = cmulc(v[q], __ldg(&tw[k * q * tscale]));
            }
            swap_reim<R>(v);
            dft_fwd<R>(v);
            swap_reim<R>(v);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) x[swz(base + q * Ms)] = v[q];
    }
}

template <bool INV>
__device__ __forceinline__ void fft_stage_dispatch(cplx* x, int R, int N, int M, const cplx* tw, int tid, int T) {
    switch (R) {
        case 2: fft_stage<2, INV>(x, N, M, tw, tid, T); break;
        case 3: fft_stage<3, INV>(x, N, M, tw, tid, T); break;
        case 4: fft_stage<4, INV>(x, N, M, tw, tid, T); break;
        case 5: fft_stage<5, INV>(x, N, M, tw, tid, T); break;
        case 7: fft_stage<7, INV>(x, N, M, tw, tid, T); break;
        default: fft_stage<8, INV>(x, N, M, tw, tid, T); break;
    }
}

// Forward transform of `nbuf` independent buffers (x + i*stride).  Caller must
// have synchronised after filling the buffers; returns synchronised.
__device__ __forceinline__ void fft_dif(cplx* x, int nbuf, int stride, const FftPlan& p, int tid, int T) {
    int M = p.N;
    for (int s = 0; s < p.nstage; ++s) {
        const int R = p.radix[s];
        for (int i = 0; i < nbuf; ++i) fft_stage_dispatch<false>(x + (size_t)i * stride, R, p.N, M, p.tw, tid, T);
        M /= R;
        __syncthreads();
    }
}

// Unnormalised inverse (result = N * ifft).  Same synchronisation contract.
__device__ __forceinline__ void fft_dit_inv(cplx* x, int nbuf, int stride, const FftPlan& p, int tid, int T) {
    int M = 1;
    for (int s = p.nstage - 1; s >= 0; --s) {
        const int R = p.radix[s];
        M *= R;
        for (int i = 0; i < nbuf; ++i) fft_stage_dispatch<true>(x + (size_t)i * stride, R, p.N, M, p.tw, tid, T);
        __syncthreads();
    }
}

}  // namespace pkb
