// Shared-memory mixed-radix complex128 FFT, register radices 2..12, sm_100a.
//
// Replaces the pocketfft/ducc complex FFT the reference reaches through
// scipy.fftpack (CalcSol.py:24,35,65,99) and the Reikna FFT of cuda_lib.py:42-54.
//
// Formulation: decimation-in-frequency, in place, natural-order input ->
// digit-reversed output (forward); the inverse is the exact transpose,
// decimation-in-time, digit-reversed input -> natural-order output.  Pointwise
// products are taken in the permuted domain, so the convolution pipeline never
// un-permutes; only the two-real-rows pack/unpack steps look up the `pair`
// table to pair bin k with bin N-k.
//
// A length-N transform is 2-5 passes over ONE shared-memory buffer of N
// complex128; each pass is a radix-R butterfly held entirely in registers.
// The two scarce per-SM resources of these kernels are shared-memory bandwidth
// (128 B/clk) and fp64 issue (64 lanes/clk): a pass moves 32 B per point
// through shared memory, so twiddle factors are NOT read as tables -- each
// butterfly reads one base twiddle w = exp(-2 pi i k / M) from a small
// shared-memory table (sum of M_s / R_s entries over the stages, 7 KB at
// N = 4704) and forms w^2 .. w^(R-1) by a depth-4 product tree.
#pragma once
#include "pkb_platform.cuh"
#include "fft_tables.cuh"

namespace pkb {

#define PKB_FFT_MAX_STAGES 8

struct FftPlan {
    int N;
    int nstage;
    int ntw;                    // entries of twb (stages 1 .. nstage-1)
    int threads;                // row kernels: threads per CTA (fewer when more CTAs fit an SM, see pkb200.cu:get_plan)
    int cols_threads, cols_kb;  // k_cols geometry: threads per column, last-stage blocks per thread
    int grid_rows, grid_cols;   // persistent grid sizes (host side): resident CTAs per SM (at most 4) x SM count
    unsigned long long rpack;   // radix of stage s in bits [4s, 4s+4)
    // device tables
    const cplx* tw0;            // stage 0: exp(-2 pi i k / N), k in [0, N / R_0) -- read from global memory (L2): the
                                // stage-0 inputs (forward) come from global memory anyway, and keeping this largest
                                // table out of shared memory is what lets a third 4704-point transform fit an SM
    const cplx* twb;            // stages 1.., one after the other: exp(-2 pi i k / M_s), k in [0, M_s / R_s) -> shared memory
    const int2* pair;           // pair[k] = (pos(k), pos((N - k) mod N)), pos = index of bin k after the forward pass
    const int* perm;            // perm[k] = pos(k)
    // Hermitian pack / unpack loops (row kernels): thread slot s handles the column pair (2c, 2c+1), c = slot[s];
    // spair[s] = (pos(2c), pos(N - 2c), pos(2c + 1), pos(N - 2c - 1)).  The slots are a permutation of the
    // pairs inside windows of 64, chosen on the host so that the 8 threads of a quarter-warp touch 8
    // different 16-byte bank groups: pos(k) of consecutive k differ by N / R_0 (392 at N = 4704, a multiple of
    // 8 elements), which in natural order put 5-6 of every 8 accesses on one bank group.
    const int* slot;
    const int4* spair;
};

__host__ __device__ __forceinline__ int plan_radix(const FftPlan& p, int s) { return (int)((p.rpack >> (4 * s)) & 15ull); }

// floor(j / d) for 0 <= j < 2^16, 1 <= d < 2^16 with one multiply
__device__ __forceinline__ int fast_div(int j, int d, unsigned magic) { return d == 1 ? j : (int)__umulhi((unsigned)j, magic); }
__device__ __forceinline__ unsigned div_magic(int d) { return d == 1 ? 0u : 0xFFFFFFFFu / (unsigned)d + 1u; }

// shared-memory footprint (bytes) of one transform: data + base twiddles of the inner stages
// (never below 1 KB: the row kernels reuse the start of the buffer as reduction scratch)
__host__ __device__ __forceinline__ size_t fft_smem_bytes(const FftPlan& p) {
    const size_t b = (size_t)(p.N + p.ntw) * sizeof(cplx);
    return b < 1024 ? 1024 : b;
}

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

// a * w, a * conj(w), w * w with fused multiply-adds
__device__ __forceinline__ cplx cmul_f(cplx a, cplx w) {
    return cmake(fma(-a.y, w.y, a.x * w.x), fma(a.x, w.y, a.y * w.x));
}
__device__ __forceinline__ cplx cmulc_f(cplx a, cplx w) {
    return cmake(fma(a.y, w.y, a.x * w.x), fma(-a.x, w.y, a.y * w.x));
}
__device__ __forceinline__ cplx csqr_f(cplx w) {
    return cmake(fma(-w.y, w.y, w.x * w.x), (w.x + w.x) * w.y);
}

// v[q] *= w^q (forward) or conj(w)^q (inverse), q = 1 .. R-1, powers by a product
// tree of depth <= 4 (w^q = w^ceil(q/2) * w^floor(q/2)); rounding of the
// derived factors stays at a few ulp, far below the 1e-10 parity bar
template <int R, bool CONJ>
__device__ __forceinline__ void twiddle_apply(cplx (&v)[R], cplx w1) {
    if constexpr (R > 1) {
        cplx w[R];
        w[0] = w1;   // unused
        w[1] = w1;
#pragma unroll
        for (int q = 2; q < R; ++q) w[q] = (q & 1) ? cmul_f(w[q / 2 + 1], w[q / 2]) : csqr_f(w[q / 2]);
#pragma unroll
        for (int q = 1; q < R; ++q) v[q] = CONJ ? cmulc_f(v[q], w[q]) : cmul_f(v[q], w[q]);
    }
}

// ---- one stage over a block decomposition of size M -----------------------------
// INV = false: DIF forward stage (DFT_R then twiddle on outputs)
// INV = true : DIT inverse stage (conj twiddle on inputs then inverse DFT_R)
// twk: this stage's base twiddles exp(-2 pi i k / M), k in [0, M/R)  (shared memory)
// Inputs come from ld(index) and outputs go to st(index, value): shared memory for
// inner stages, global memory for the first forward / last inverse stage.
template <int R, bool INV, bool GTW, class Load, class Store>
__device__ __forceinline__ void fft_stage(int N, int M, const cplx* twk, int tid, int T, Load ld, Store st) {
    const int Ms = M / R;
    const int nb = N / R;
    // Butterfly j of the stage = (block b, offset k).  Consecutive threads normally take consecutive k
    // (contiguous 16-byte elements: conflict free while Ms >= 8).  Once M is odd (all even radices
    // consumed, e.g. M = 49 of 12.8.7.7) consecutive b are M elements apart and M is coprime to the 8
    // 16-byte bank groups, so b runs fastest there instead: a quarter-warp then hits 8 distinct bank
    // groups where the k-fastest order gave 2-way conflicts on every access (ncu: 1.9-2.0 wavefronts
    // per ideal wavefront on that stage), and all of its threads read the same base twiddle.
    const int nblk = N / M;
#ifdef PKB_NO_BFAST
    const bool bfast = false;
#else
    const bool bfast = !GTW && (M & 1) && Ms > 1 && nblk >= 8;
#endif
    const int dv = bfast ? nblk : Ms;
    const unsigned magic = div_magic(dv);
#pragma unroll 1
    for (int j = tid; j < nb; j += T) {
        const int hi = fast_div(j, dv, magic);
        const int lo = j - hi * dv;
        const int b = bfast ? lo : hi;
        const int k = bfast ? hi : lo;
        const int base = b * M + k;
        cplx v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = ld(base + q * Ms);
        if (Ms > 1) {
            const cplx w1 = GTW ? __ldg(&twk[k]) : twk[k];
            if (!INV) {
                dft<R>(v);
                twiddle_apply<R, false>(v, w1);
            } else {
                twiddle_apply<R, true>(v, w1);
                idft<R>(v);
            }
        } else {
            if (!INV) dft<R>(v);
            else idft<R>(v);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) st(base + q * Ms, v[q]);
    }
}

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

// GTW: twk points to global memory (stage 0) instead of shared memory
template <bool INV, bool GTW, class Load, class Store>
__device__ __forceinline__ void fft_stage_dispatch(int R, int N, int M, const cplx* twk, int tid, int T, Load ld, Store st) {
#define PKB_CALL_(RR) fft_stage<RR, INV, GTW>(N, M, twk, tid, T, ld, st)
    PKB_RADIX_SWITCH(R, PKB_CALL_)
#undef PKB_CALL_
}

// shared-memory accessors of the in-place stages
struct SmemLoad {
    const cplx* x;
    __device__ __forceinline__ cplx operator()(int i) const { return x[i]; }
};
struct SmemStore {
    cplx* x;
    __device__ __forceinline__ void operator()(int i, cplx v) const { x[i] = v; }
};

// copy the plan's base twiddles into shared memory (visible after the next barrier)
__device__ __forceinline__ void fft_load_twiddles(cplx* tws, const FftPlan& p, int tid, int T) {
    for (int i = tid; i < p.ntw; i += T) tws[i] = p.twb[i];
}

// offset of stage s's (s >= 1) base table inside twb
__device__ __forceinline__ int fft_tw_offset(const FftPlan& p, int s) {
    int M = p.N / plan_radix(p, 0), off = 0;
    for (int t = 1; t < s; ++t) {
        const int R = plan_radix(p, t);
        off += M / R;
        M /= R;
    }
    return off;
}

// Forward stages s0 .. s1-1 in shared memory (x holds the data on entry; M0 / off0 are the
// block size and table offset of stage s0).  A barrier follows every stage (the one after
// the last stage only if final_sync).
__device__ __forceinline__ void fft_fwd_stages(cplx* x, const cplx* tws, const FftPlan& p, int s0, int s1, int M0, int off0, int tid, int T,
                                               bool final_sync = true) {
    int M = M0, off = off0;
    for (int s = s0; s < s1; ++s) {
        const int R = plan_radix(p, s);
        fft_stage_dispatch<false, false>(R, p.N, M, tws + off, tid, T, SmemLoad{x}, SmemStore{x});
        off += M / R;
        M /= R;
        if (final_sync || s + 1 < s1) __syncthreads();
    }
}
// Inverse stages s1-1 down to s0 in shared memory; M1 = block size AFTER stage s1-1 has been
// undone is M1 * R_{s1-1}..., so pass the block size of stage s1 (1 when s1 == nstage) and the
// table offset of stage s1.  A barrier follows every stage.
__device__ __forceinline__ void fft_inv_stages(cplx* x, const cplx* tws, const FftPlan& p, int s1, int s0, int Mnext, int offnext, int tid, int T) {
    int M = Mnext, off = offnext;
    for (int s = s1 - 1; s >= s0; --s) {
        const int R = plan_radix(p, s);
        M *= R;
        off -= M / R;
        fft_stage_dispatch<true, false>(R, p.N, M, tws + off, tid, T, SmemLoad{x}, SmemStore{x});
        __syncthreads();
    }
}

// Whole forward transform with the first stage fed by ld(index) (global memory).
// tws must be loaded; the caller guarantees x is free.  Returns synchronised.
template <class Load>
__device__ __forceinline__ void fft_forward_from(cplx* x, const cplx* tws, const FftPlan& p, int tid, int T, Load ld, bool final_sync = true) {
    const int R0 = plan_radix(p, 0);
    fft_stage_dispatch<false, true>(R0, p.N, p.N, p.tw0, tid, T, ld, SmemStore{x});
    if (final_sync || p.nstage > 1) __syncthreads();
    fft_fwd_stages(x, tws, p, 1, p.nstage, p.N / R0, 0, tid, T, final_sync);
}

// Whole inverse transform (unnormalised, result = N * ifft) with the last stage
// delivering to st(index, value).  x holds digit-reversed input, synchronised.
template <class Store>
__device__ __forceinline__ void fft_inverse_to(cplx* x, const cplx* tws, const FftPlan& p, int tid, int T, Store st) {
    const int R0 = plan_radix(p, 0);
    fft_inv_stages(x, tws, p, p.nstage, 1, 1, p.ntw, tid, T);
    fft_stage_dispatch<true, true>(R0, p.N, p.N, p.tw0, tid, T, SmemLoad{x}, st);
}

}  // namespace pkb
