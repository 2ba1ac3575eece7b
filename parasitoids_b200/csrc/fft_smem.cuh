// Shared-memory mixed-radix complex128 FFT with register radices up to 21, sm_100a.
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
// A length-N transform is 2-4 passes over shared memory; each pass is a
// radix-R butterfly (R in {2..21}) held entirely in registers.  Composite
// radices (9, 10, 12, 14, 15, 16, 18, 20, 21) are Cooley-Tukey products of the
// base codelets with compile-time twiddles, so fp64 instruction count -- the
// real ceiling of these kernels on B200 (36 TFLOP/s fp64) -- stays close to
// the split-radix count while shared-memory traffic drops to 3 round trips.
//
// One transform lives in one shared-memory buffer of Npad = roundup(N, 64)
// complex128, addressed through an XOR swizzle so that power-of-two strides do
// not serialise on the 16-byte bank groups.
#pragma once
#include "pkb_platform.cuh"
#include "fft_tables.cuh"

namespace pkb {

#define PKB_FFT_MAX_STAGES 16

struct FftPlan {
    int N;
    int Npad;
    int nstage;
    int radix[PKB_FFT_MAX_STAGES];
    const cplx* tw;   // device: tw[j] = exp(-2 pi i j / N), j in [0, N)
    const int* perm;  // device: perm[k] = position of frequency k after fft_dif
    int cols_threads; // k_cols launch: threads per column and last-stage blocks per thread
    int cols_kb;
    // twm: twiddles of the inner stages s = 1 .. nstage-2, stage after stage; stage s
    // holds (R_s - 1) * Ms_s entries laid out [q-1][k] = exp(-2 pi i k q / M_s), so that
    // consecutive threads (consecutive k) read consecutive entries
    const cplx* twm;
    unsigned long long rpack;   // radix[s] in 4-bit fields (register-friendly copy of radix[])
};

__host__ __device__ __forceinline__ int plan_radix(const FftPlan& p, int s) { return (int)((p.rpack >> (4 * s)) & 15ull); }

// floor(j / d) for 0 <= j < 2^16, 1 <= d < 2^16 with one multiply
__device__ __forceinline__ int fast_div(int j, int d, unsigned magic) { return d == 1 ? j : (int)__umulhi((unsigned)j, magic); }
__device__ __forceinline__ unsigned div_magic(int d) { return d == 1 ? 0u : 0xFFFFFFFFu / (unsigned)d + 1u; }

// Shared-memory index map.  Identity: every pass either has consecutive threads on
// consecutive elements or (last stage) a per-thread stride equal to the last
// radix, which the planner keeps odd whenever N has an odd factor -- both are
// conflict-free for 16-byte accesses.
__host__ __device__ __forceinline__ int swz(int i) { return i; }

// ---- compile-time twiddles -----------------------------------------------------
// a * exp(-2 pi i J / R)
template <int R, int J>
__device__ __forceinline__ cplx mul_w(cplx a) {
    constexpr int j = ((J % R) + R) % R;
    constexpr double h = 0.7071067811865475244008;
    if constexpr (j == 0) return a;
    else if constexpr (4 * j == R) return cmake(a.y, -a.x);
    else if constexpr (2 * j == R) return cmake(-a.x, -a.y);
    else if constexpr (4 * j == 3 * R) return cmake(-a.y, a.x);
    else if constexpr (8 * j == R) return cmake(h * (a.x + a.y), h * (a.y - a.x));
    else if constexpr (8 * j == 3 * R) return cmake(h * (a.y - a.x), -h * (a.x + a.y));
    else if constexpr (8 * j == 5 * R) return cmake(-h * (a.x + a.y), h * (a.x - a.y));
    else if constexpr (8 * j == 7 * R) return cmake(h * (a.x - a.y), h * (a.x + a.y));
    else {
        constexpr double c = TwTab<R>::c[j], s = TwTab<R>::s[j];
        return cmake(fma(a.y, s, a.x * c), fma(-a.x, s, a.y * c));
    }
}

// ---- base codelets (forward sign: exp(-i...)) --------------------------------------
template <int R>
__device__ __forceinline__ void dft(cplx (&v)[R]);

template <>
__device__ __forceinline__ void dft<2>(cplx (&v)[2]) {
    const cplx t = csub(v[0], v[1]);
    v[0] = cadd(v[0], v[1]);
    v[1] = t;
}

__device__ __forceinline__ void dft4(cplx& x0, cplx& x1, cplx& x2, cplx& x3) {
    const cplx t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), d = csub(x1, x3);
    const cplx t3 = cmake(d.y, -d.x);   // -i (x1 - x3)
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    x1 = cadd(t1, t3);
    x3 = csub(t1, t3);
}
template <>
__device__ __forceinline__ void dft<4>(cplx (&v)[4]) { dft4(v[0], v[1], v[2], v[3]); }

// odd length via the pairing x_m +- x_{R-m}; (Q, M) recursion keeps every
// table index a compile-time constant
template <int R, int Q, int M>
__device__ __forceinline__ void odd_acc(cplx& C, cplx& S, const cplx (&a)[(R - 1) / 2], const cplx (&b)[(R - 1) / 2]) {
    constexpr double cc = TwTab<R>::c[(Q * M) % R], ss = TwTab<R>::s[(Q * M) % R];
    C.x = fma(cc, a[M - 1].x, C.x);
    C.y = fma(cc, a[M - 1].y, C.y);
    if constexpr (M == 1) {
        S.x = ss * b[0].x;
        S.y = ss * b[0].y;
    } else {
        S.x = fma(ss, b[M - 1].x, S.x);
        S.y = fma(ss, b[M - 1].y, S.y);
    }
    if constexpr (M < (R - 1) / 2) odd_acc<R, Q, M + 1>(C, S, a, b);
}
template <int R, int Q>
__device__ __forceinline__ void odd_out(cplx (&v)[R], cplx x0, const cplx (&a)[(R - 1) / 2], const cplx (&b)[(R - 1) / 2]) {
    cplx C = x0, S = cmake(0.0, 0.0);
    odd_acc<R, Q, 1>(C, S, a, b);
    v[Q] = cmake(C.x + S.y, C.y - S.x);      // C - i S
    v[R - Q] = cmake(C.x - S.y, C.y + S.x);  // C + i S
    if constexpr (Q < (R - 1) / 2) odd_out<R, Q + 1>(v, x0, a, b);
}
template <int R>
__device__ __forceinline__ void dft_odd(cplx (&v)[R]) {
    constexpr int H = (R - 1) / 2;
    cplx a[H], b[H];
#pragma unroll
    for (int m = 1; m <= H; ++m) {
        a[m - 1] = cadd(v[m], v[R - m]);
        b[m - 1] = csub(v[m], v[R - m]);
    }
    const cplx x0 = v[0];
    cplx s0 = x0;
#pragma unroll
    for (int m = 0; m < H; ++m) s0 = cadd(s0, a[m]);
    v[0] = s0;
    odd_out<R, 1>(v, x0, a, b);
}
template <> __device__ __forceinline__ void dft<3>(cplx (&v)[3]) { dft_odd<3>(v); }
template <> __device__ __forceinline__ void dft<5>(cplx (&v)[5]) { dft_odd<5>(v); }
template <> __device__ __forceinline__ void dft<7>(cplx (&v)[7]) { dft_odd<7>(v); }

// ---- composite codelets: R = RA * RB, all indices compile-time -------------------
template <int R, int RA, int I>
__device__ __forceinline__ void ct_twiddle(cplx (&t)[R]) {
    t[I] = mul_w<R, (I / RA) * (I % RA)>(t[I]);
    if constexpr (I + 1 < R) ct_twiddle<R, RA, I + 1>(t);
}
// X[k1 + RA k2] = sum_{n2} w_RB^{n2 k2} [ w_R^{n2 k1} sum_{n1} w_RA^{n1 k1} x[RB n1 + n2] ]
template <int RA, int RB>
__device__ __forceinline__ void dft_ct(cplx (&v)[RA * RB]) {
    constexpr int R = RA * RB;
    cplx t[R];
#pragma unroll
    for (int n2 = 0; n2 < RB; ++n2) {
        cplx u[RA];
#pragma unroll
        for (int n1 = 0; n1 < RA; ++n1) u[n1] = v[RB * n1 + n2];
        dft<RA>(u);
#pragma unroll
        for (int k1 = 0; k1 < RA; ++k1) t[n2 * RA + k1] = u[k1];
    }
    ct_twiddle<R, RA, 0>(t);   // t[n2 * RA + k1] *= w_R^(n2 k1)
#pragma unroll
    for (int k1 = 0; k1 < RA; ++k1) {
        cplx u[RB];
#pragma unroll
        for (int n2 = 0; n2 < RB; ++n2) u[n2] = t[n2 * RA + k1];
        dft<RB>(u);
#pragma unroll
        for (int k2 = 0; k2 < RB; ++k2) v[k1 + RA * k2] = u[k2];
    }
}
template <> __device__ __forceinline__ void dft<6>(cplx (&v)[6]) { dft_ct<2, 3>(v); }
template <> __device__ __forceinline__ void dft<8>(cplx (&v)[8]) { dft_ct<2, 4>(v); }
template <> __device__ __forceinline__ void dft<9>(cplx (&v)[9]) { dft_ct<3, 3>(v); }
template <> __device__ __forceinline__ void dft<10>(cplx (&v)[10]) { dft_ct<2, 5>(v); }
template <> __device__ __forceinline__ void dft<12>(cplx (&v)[12]) { dft_ct<4, 3>(v); }
template <> __device__ __forceinline__ void dft<14>(cplx (&v)[14]) { dft_ct<2, 7>(v); }
template <> __device__ __forceinline__ void dft<15>(cplx (&v)[15]) { dft_ct<3, 5>(v); }
template <> __device__ __forceinline__ void dft<16>(cplx (&v)[16]) { dft_ct<4, 4>(v); }
template <> __device__ __forceinline__ void dft<18>(cplx (&v)[18]) { dft_ct<2, 9>(v); }
template <> __device__ __forceinline__ void dft<20>(cplx (&v)[20]) { dft_ct<4, 5>(v); }
template <> __device__ __forceinline__ void dft<21>(cplx (&v)[21]) { dft_ct<3, 7>(v); }

template <int R>
__device__ __forceinline__ void swap_reim(cplx (&v)[R]) {
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const double t = v[q].x;
        v[q].x = v[q].y;
        v[q].y = t;
    }
}
// unnormalised inverse DFT in registers
template <int R>
__device__ __forceinline__ void idft(cplx (&v)[R]) {
    swap_reim<R>(v);
    dft<R>(v);
    swap_reim<R>(v);
}

// a * w, a * conj(w) with fused multiply-adds
__device__ __forceinline__ cplx cmul_f(cplx a, cplx w) {
    return cmake(fma(-a.y, w.y, a.x * w.x), fma(a.x, w.y, a.y * w.x));
}
__device__ __forceinline__ cplx cmulc_f(cplx a, cplx w) {
    return cmake(fma(a.y, w.y, a.x * w.x), fma(-a.x, w.y, a.y * w.x));
}

// ---- twiddles of the outermost stage (M = N) ------------------------------------
// w_N^(k q), q = 1..R-1, from ONE table load per butterfly: successive products
// w^q = w^(q-1) w.  The outermost stage would otherwise gather R-1 entries of a
// table as large as the transform itself (75 KB at N = 4704); rounding grows to
// ~R ulp on these factors, far below the 1e-10 parity bar.
template <int R, bool CONJ>
__device__ __forceinline__ void twiddle_chain(cplx (&v)[R], cplx w1) {
    cplx w = w1;
#pragma unroll
    for (int q = 1; q < R; ++q) {
        v[q] = CONJ ? cmulc_f(v[q], w) : cmul_f(v[q], w);
        if (q + 1 < R) w = cmul_f(w, w1);
    }
}

// ---- one in-place stage over a block decomposition of size M -----------------
// INV = false: DIF forward stage (DFT_R then twiddle on outputs)
// INV = true : DIT inverse stage (conj twiddle on inputs then inverse DFT_R)
// CHAIN: outermost stage (M == N), twiddles by twiddle_chain
template <int R, bool INV, bool CHAIN>
__device__ __forceinline__ void fft_stage(cplx* x, int N, int M, const cplx* __restrict__ tw, int tid, int T) {
    // tw: CHAIN -> the length-N table (entry k); else this stage's [q-1][k] table
    const int Ms = M / R;
    const int nb = N / R;
    const unsigned magic = div_magic(Ms);
    for (int j = tid; j < nb; j += T) {
        const int b = fast_div(j, Ms, magic);
        const int k = j - b * Ms;
        const int base = b * M + k;
        cplx w1;
        if (CHAIN) w1 = __ldg(&tw[k]);
        cplx v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = x[swz(base + q * Ms)];
        if (!INV) {
            dft<R>(v);
            if (k > 0) {
                if (CHAIN) twiddle_chain<R, false>(v, w1);
                else {
#pragma unroll
                    for (int q = 1; q < R; ++q) v[q] = cmul_f(v[q], __ldg(&tw[(q - 1) * Ms + k]));
                }
            }
        } else {
            if (k > 0) {
                if (CHAIN) twiddle_chain<R, true>(v, w1);
                else {
#pragma unroll
                    for (int q = 1; q < R; ++q) v[q] = cmulc_f(v[q], __ldg(&tw[(q - 1) * Ms + k]));
                }
            }
            idft<R>(v);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) x[swz(base + q * Ms)] = v[q];
    }
}

// First forward stage (M = N) with inputs supplied by `ld(index)` instead of
// shared memory, and last inverse stage with outputs consumed by `st(index, value)`.
template <int R, class Load>
__device__ __forceinline__ void fft_stage_first(cplx* x, int N, const cplx* __restrict__ tw, int tid, int T, Load ld) {
    const int Ms = N / R;
    for (int k = tid; k < Ms; k += T) {
        const cplx w1 = __ldg(&tw[k]);
        cplx v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = ld(k + q * Ms);
        dft<R>(v);
        if (k > 0) twiddle_chain<R, false>(v, w1);
#pragma unroll
        for (int q = 0; q < R; ++q) x[swz(k + q * Ms)] = v[q];
    }
}
template <int R, class Store>
__device__ __forceinline__ void fft_stage_last_inv(const cplx* x, int N, const cplx* __restrict__ tw, int tid, int T, Store st) {
    const int Ms = N / R;
    for (int k = tid; k < Ms; k += T) {
        const cplx w1 = __ldg(&tw[k]);
        cplx v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = x[swz(k + q * Ms)];
        if (k > 0) twiddle_chain<R, true>(v, w1);
        idft<R>(v);
#pragma unroll
        for (int q = 0; q < R; ++q) st(k + q * Ms, v[q]);
    }
}

// radices a plan may use (pkb200.cu: kRadices); larger composite codelets exist
// above but cost registers -- and therefore resident warps -- for no gain in
// fp64 instruction count
#define PKB_RADIX_SWITCH(R, CALL)            \
    switch (R) {                             \
        case 2: { CALL(2); } break;          \
        case 3: { CALL(3); } break;          \
        case 4: { CALL(4); } break;          \
        case 5: { CALL(5); } break;          \
        case 6: { CALL(6); } break;          \
        case 7: { CALL(7); } break;          \
        case 8: { CALL(8); } break;          \
        case 9: { CALL(9); } break;          \
        case 10: { CALL(10); } break;        \
        default: { CALL(12); } break;        \
    }

// tw: the length-N table when M == N (outermost stage), else the stage's own table
template <bool INV>
__device__ __forceinline__ void fft_stage_dispatch(cplx* x, int R, int N, int M, const cplx* tw, int tid, int T) {
    if (M == N) {
#define PKB_CALL_(RR) fft_stage<RR, INV, true>(x, N, M, tw, tid, T)
        PKB_RADIX_SWITCH(R, PKB_CALL_)
#undef PKB_CALL_
    } else {
#define PKB_CALL_(RR) fft_stage<RR, INV, false>(x, N, M, tw, tid, T)
        PKB_RADIX_SWITCH(R, PKB_CALL_)
#undef PKB_CALL_
    }
}
template <class Load>
__device__ __forceinline__ void fft_stage_first_dispatch(cplx* x, int R, int N, const cplx* tw, int tid, int T, Load ld) {
#define PKB_CALL_(RR) fft_stage_first<RR>(x, N, tw, tid, T, ld)
    PKB_RADIX_SWITCH(R, PKB_CALL_)
#undef PKB_CALL_
}
template <class Store>
__device__ __forceinline__ void fft_stage_last_inv_dispatch(const cplx* x, int R, int N, const cplx* tw, int tid, int T, Store st) {
#define PKB_CALL_(RR) fft_stage_last_inv<RR>(x, N, tw, tid, T, st)
    PKB_RADIX_SWITCH(R, PKB_CALL_)
#undef PKB_CALL_
}

// Twiddle table of inner stage s, and the running offset bookkeeping
__device__ __forceinline__ int stage_tw_size(int R, int M) { return (R - 1) * (M / R); }

// One forward stage s (block size M) of a plan: outermost stage uses the chain
// on plan.tw, inner stages their [q-1][k] table at twm + off.
__device__ __forceinline__ void plan_stage_fwd(cplx* x, const FftPlan& p, int R, int M, int off, int tid, int T) {
    fft_stage_dispatch<false>(x, R, p.N, M, M == p.N ? p.tw : p.twm + off, tid, T);
}
__device__ __forceinline__ void plan_stage_inv(cplx* x, const FftPlan& p, int R, int M, int off, int tid, int T) {
    fft_stage_dispatch<true>(x, R, p.N, M, M == p.N ? p.tw : p.twm + off, tid, T);
}

// Forward transform of one buffer filled (and synchronised) by the caller; returns synchronised.
__device__ __forceinline__ void fft_dif(cplx* x, const FftPlan& p, int tid, int T) {
    int M = p.N, off = 0;
    for (int s = 0; s < p.nstage; ++s) {
        const int R = plan_radix(p, s);
        plan_stage_fwd(x, p, R, M, off, tid, T);
        if (s > 0) off += stage_tw_size(R, M);
        M /= R;
        __syncthreads();
    }
}

// Forward transform of one buffer whose inputs come from `ld(index)` (global
// memory): the first stage reads them straight into registers.  Returns synchronised.
template <class Load>
__device__ __forceinline__ void fft_dif_from(cplx* x, const FftPlan& p, int tid, int T, Load ld) {
    const int R0 = plan_radix(p, 0);
    if (p.nstage == 1) {
        for (int i = tid; i < p.N; i += T) x[swz(i)] = ld(i);
        __syncthreads();
        plan_stage_fwd(x, p, R0, p.N, 0, tid, T);
        __syncthreads();
        return;
    }
    fft_stage_first_dispatch(x, R0, p.N, p.tw, tid, T, ld);
    __syncthreads();
    int M = p.N / R0, off = 0;
    for (int s = 1; s < p.nstage; ++s) {
        const int R = plan_radix(p, s);
        plan_stage_fwd(x, p, R, M, off, tid, T);
        off += stage_tw_size(R, M);
        M /= R;
        __syncthreads();
    }
}

// total size of the inner-stage tables up to and including stage `upto`
__device__ __forceinline__ int plan_tw_offset(const FftPlan& p, int upto) {
    int M = p.N / plan_radix(p, 0), off = 0;
    for (int s = 1; s <= upto; ++s) {
        const int R = plan_radix(p, s);
        off += stage_tw_size(R, M);
        M /= R;
    }
    return off;
}

// Unnormalised inverse (result = N * ifft).  Same synchronisation contract.
__device__ __forceinline__ void fft_dit_inv(cplx* x, const FftPlan& p, int tid, int T) {
    int M = 1;
    int off = plan_tw_offset(p, p.nstage - 1);
    for (int s = p.nstage - 1; s >= 0; --s) {
        const int R = plan_radix(p, s);
        M *= R;
        if (s > 0) off -= stage_tw_size(R, M);
        plan_stage_inv(x, p, R, M, off, tid, T);
        __syncthreads();
    }
}

}  // namespace pkb
