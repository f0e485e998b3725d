// tsdr_fft.cu -- hand-written FFT autocorrelation for sm_100a (no cuFFT, no tensor cores).
//
// Replaces calculate_autocorrelation (src/Autocorrelations.jl:23-37):
//     theCorr = ifft(fft(x) .* conj(fft(x)));  10*log10.(abs2.(theCorr[indexMin:indexMax]))
//
// Design (DESIGN.md "K5"): the input is real, so the N-point transforms run as
// M = N/2 point complex transforms of z[j] = x[2j] + i x[2j+1].  M = A*B is split
// four-step style and the whole autocorrelation takes THREE kernels:
//   k_fft_cols   A-point forward FFTs down the columns (stride B) + twiddle W_M^(j2 k1)
//   k_fft_mid    per pair of rows (k1, A-k1): B-point forward FFT, real-input unpack,
//                |X|^2, Hermitian repack, B-point inverse FFT, twiddle -- all in shared memory
//   k_ifft_cols  A-point inverse FFTs down the columns, epilogue 10*log10(r^2) of the lag slice
// Every in-shared-memory FFT is an in-place radix-16/8/4/2 register butterfly network
// (decimation in frequency forward, decimation in time inverse), so no bit-reversal pass
// is ever needed: the forward result stays in digit-reversed order and the matching
// inverse consumes it.  Global traffic: 5 * 8M + 4L bytes.
// Lengths that are not a power of two use the zero-padded transform of size
// N >= 2n and fold the linear lags: r_circ[k] = r_lin[k] + r_lin[n-k].
#include "tsdr_internal.cuh"

#include <cstdlib>
#include <new>
#include <vector>

namespace tsdr {

// ---------------------------------------------------------------- helpers --
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by -i (DIR=+1, forward) or +i (DIR=-1, inverse)
template <int DIR> __device__ __forceinline__ float2 mul_mi(float2 a) {
    return DIR > 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}
template <int DIR> __device__ __forceinline__ float2 mul_c(float2 a, float c, float s) {  // a * (c - i s) fwd, (c + i s) inv
    const float ss = DIR > 0 ? -s : s;
    return make_float2(fmaf(a.x, c, -a.y * ss), fmaf(a.x, ss, a.y * c));
}

// R-point DFT in registers, natural order in and out.  DIR=+1: e^{-2 pi i pq/R}.
template <int DIR> __device__ __forceinline__ void dft2(float2& a, float2& b) {
    const float2 t = csub(a, b); a = cadd(a, b); b = t;
}
template <int DIR> __device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = mul_mi<DIR>(csub(x1, x3));
    x0 = cadd(t0, t2); x2 = csub(t0, t2); x1 = cadd(t1, t3); x3 = csub(t1, t3);
}
template <int DIR> __device__ __forceinline__ void dft8(float2* x) {  // x[0..7]
    float2 e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6], o0 = x[1], o1 = x[3], o2 = x[5], o3 = x[7];
    dft4<DIR>(e0, e1, e2, e3);
    dft4<DIR>(o0, o1, o2, o3);
    const float h = 0.70710678118654752440f;
    o1 = mul_c<DIR>(o1, h, h);
    o2 = mul_mi<DIR>(o2);
    o3 = mul_c<DIR>(o3, -h, h);
    x[0] = cadd(e0, o0); x[4] = csub(e0, o0);
    x[1] = cadd(e1, o1); x[5] = csub(e1, o1);
    x[2] = cadd(e2, o2); x[6] = csub(e2, o2);
    x[3] = cadd(e3, o3); x[7] = csub(e3, o3);
}
template <int DIR> __device__ __forceinline__ void dft16(float2* x) {
    float2 e[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { e[i] = x[2 * i]; o[i] = x[2 * i + 1]; }
    dft8<DIR>(e);
    dft8<DIR>(o);
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
    o[1] = mul_c<DIR>(o[1], c1, s1);
    o[2] = mul_c<DIR>(o[2], h, h);
    o[3] = mul_c<DIR>(o[3], s1, c1);
    o[4] = mul_mi<DIR>(o[4]);
    o[5] = mul_c<DIR>(o[5], -s1, c1);
    o[6] = mul_c<DIR>(o[6], -h, h);
    o[7] = mul_c<DIR>(o[7], -c1, s1);
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = cadd(e[i], o[i]); x[i + 8] = csub(e[i], o[i]); }
}
template <int R, int DIR> __device__ __forceinline__ void dft(float2* x) {
    if (R == 2) dft2<DIR>(x[0], x[1]);
    else if (R == 4) dft4<DIR>(x[0], x[1], x[2], x[3]);
    else if (R == 8) dft8<DIR>(x);
    else dft16<DIR>(x);
}

// shared-memory layouts (element = float2).  Rows: contiguous transform, one pad
// element per 16.  Cols: C interleaved transforms (element (i, col) at i*C + col),
// 8 pad elements per 128 so that radix blocks land in different banks.
struct RowLayout {
    int row_stride;
    __device__ __forceinline__ int operator()(int b, int i) const { return b * row_stride + i + (i >> 4); }
};
struct ColLayout {
    int C;
    __device__ __forceinline__ int operator()(int b, int i) const { const int e = i * C + b; return e + ((e >> 7) << 3); }
};
__host__ __device__ inline int row_padded(int len) { return len + (len >> 4) + 1; }
__host__ __device__ inline int col_padded(int total) { return total + ((total >> 7) << 3) + 8; }

// One in-place stage on sub-transforms of length Lcur (radix R, butterfly stride Lcur/R).
//  FWD (DIF): y = DFT_R(x); y[p] *= W_Lcur^(u p)
//  INV (DIT): x[q] *= conj W_Lcur^(u q); y = IDFT_R(x)
// tw = table of W_len^k, k in [0, len).  COLFAST: batch index varies fastest over threads.
template <int R, int DIR, bool COLFAST, class Layout>
__device__ __forceinline__ void fft_stage(float2* s, const Layout lay, int nbatch, int len, int Lcur,
                                          const float2* __restrict__ tw, int tid, int nthr) {
    const int sub = Lcur / R;
    const int per = len / R;
    const int tws = len / Lcur;
    for (int e = tid; e < nbatch * per; e += nthr) {
        int b, w;
        if (COLFAST) { b = e % nbatch; w = e / nbatch; } else { b = e / per; w = e - b * per; }
        const int blk = w / sub, u = w - blk * sub;
        const int base = blk * Lcur + u;
        float2 x[R];
#pragma unroll
        for (int q = 0; q < R; ++q) x[q] = s[lay(b, base + q * sub)];
        if (DIR < 0 && u != 0) {
#pragma unroll
            for (int q = 1; q < R; ++q) x[q] = cmul(x[q], cconj(__ldg(tw + u * q * tws)));
        }
        dft<R, DIR>(x);
        if (DIR > 0 && u != 0) {
#pragma unroll
            for (int p = 1; p < R; ++p) x[p] = cmul(x[p], __ldg(tw + u * p * tws));
        }
#pragma unroll
        for (int p = 0; p < R; ++p) s[lay(b, base + p * sub)] = x[p];
    }
}

struct Radices { int n; int r[8]; };

template <int DIR, bool COLFAST, class Layout>
__device__ __forceinline__ void fft_inplace(float2* s, const Layout lay, int nbatch, int len, const Radices rad,
                                            const float2* __restrict__ tw, int tid, int nthr) {
    if (DIR > 0) {
        int Lcur = len;
        for (int i = 0; i < rad.n; ++i) {
            const int R = rad.r[i];
            if (R == 16) fft_stage<16, DIR, COLFAST>(s, lay, nbatch, len, Lcur, tw, tid, nthr);
            else if (R == 8) fft_stage<8, DIR, COLFAST>(s, lay, nbatch, len, Lcur, tw, tid, nthr);
            else if (R == 4) fft_stage<4, DIR, COLFAST>(s, lay, nbatch, len, Lcur, tw, tid, nthr);
            else fft_stage<2, DIR, COLFAST>(s, lay, nbatch, len, Lcur, tw, tid, nthr);
            Lcur /= R;
            __syncthreads();
        }
    } else {
        int Lcur = 1;
        for (int i = rad.n - 1; i >= 0; --i) {
            const int R = rad.r[i];
            Lcur *= R;
            if (R == 16) fft_stage<16, DIR, COLFAST>(s, lay, nbatch, len, Lcur, tw, tid, nthr);
            else if (R == 8) fft_stage<8, DIR, COLFAST>(s, lay, nbatch, len, Lcur, tw, tid, nthr);
            else if (R == 4) fft_stage<4, DIR, COLFAST>(s, lay, nbatch, len, Lcur, tw, tid, nthr);
            else fft_stage<2, DIR, COLFAST>(s, lay, nbatch, len, Lcur, tw, tid, nthr);
            __syncthreads();
        }
    }
}

struct FftParams {
    int A, B, C;           // M = A*B, C columns per CTA in the column passes
    int64_t M;
    Radices radA, radB;
    const float2* twA;     // W_A^k
    const float2* twB;     // W_B^k
    const int* revA;       // row position -> k1
    const int* posA;       // k1 -> row position
    const int* revB;       // position -> k2
    const int* posB;       // k2 -> position
    const float2* wlo;     // W_N^t,        t in [0, LO)
    const float2* whi;     // W_N^(t*LO),   t in [0, N/LO)
    int lo_bits;
    const float* x;        // real input, n valid samples (zero beyond)
    int64_t n_valid;
    float2* T;             // workspace M complex
    float2* U;             // workspace M complex
    float inv_scale;       // 1/M
    // epilogue
    float* out;            // lag slice, or raw linear lags when fold != 0
    int64_t m_lo, m_hi;    // 0-based lag range to write
    int log_scale;
    int raw;               // write r (not r^2 / dB): used by the fold path
    // mode 1: FFT-domain integer upsampler (init_resampler / resampler!, src/Resampler.jl:26-62):
    // complex transform of the zero-stuffed real input, spectrum * H, inverse, gain * real part
    int mode;
    int up;                // zero-stuffing factor
    int64_t n_in;          // input samples (M = n_in * up)
    const double2* Hd;     // [M] filter in ComplexF64, natural frequency order
    float gain;            // 2 * upCoeff
    // mode 3: power spectrum of a complex M-point signal (getSpectrum, src/GetSpectrum.jl:21-30): forward transform,
    //         abs2, 10 log10, fftshift -- no inverse.
    // mode 2: the same for a length n_in that is not a power of two, as a chirp-z (Bluestein) convolution on the
    //         M >= 2 n_in - 1 point engine: x[j] conj(b[j]) -> FFT -> * FFT(b) (Hd) -> IFFT; |X[k]|^2 = |y[k]|^2
    //         because the final chirp factor has unit modulus.  b[j] = exp(i pi j^2 / n_in).
    const float2* chirp;   // [n_in] b[j]
    // modes 4 / 5: init_resampler / resampler! for a length N = n_in*up that is NOT a power of two, as two chirp-z
    // transforms on the M >= 2N-1 point engine.  4: zero-stuffed real input -> DFT_N -> * filt (Resampler.jl:51-53)
    // -> W = conj(.);  5: W -> DFT_N -> out = gain * real(.)/N  (IDFT(Y) = conj(DFT(conj Y))/N, Resampler.jl:55-59)
    const double2* filt;   // [N] H of initLPF, natural frequency order
    float2* W;             // [N] conj(spectrum * H) between the two transforms
    int64_t n_tot;         // N
    float inv_n;           // 1/N
};

__device__ __forceinline__ float2 twiddle_n(const FftParams& p, int64_t t) {  // W_N^t, 0 <= t < N
    const float2 lo = __ldg(p.wlo + (t & ((1 << p.lo_bits) - 1)));
    const float2 hi = __ldg(p.whi + (t >> p.lo_bits));
    return cmul(lo, hi);
}

constexpr int kFftThreads = 512;

// ------------------------------------------------------------ k_fft_cols --
// z[j1*B + j2] (the real input reinterpreted as complex pairs) -> A-point forward FFT
// over j1 for C adjacent columns j2 -> * W_M^(j2 k1) -> T[row][j2], row = digit-reversed k1.
__global__ void __launch_bounds__(kFftThreads) k_fft_cols(FftParams p) {
    extern __shared__ float2 sm[];
    const int tid = threadIdx.x;
    const int j2_0 = blockIdx.x * p.C;
    const ColLayout lay{p.C};
    const int total = p.A * p.C;
    // load: consecutive threads read consecutive columns of one row (C*8 contiguous bytes)
    for (int e = tid; e < total; e += kFftThreads) {
        const int j1 = e / p.C, col = e - j1 * p.C;
        const int64_t j = (int64_t)j1 * p.B + j2_0 + col;
        float2 v;
        if (p.mode == 1) {  // containerFFT[1:upCoeff:end] .= in   (Resampler.jl:48)
            const int64_t q = j / p.up;
            v.x = (q * p.up == j && q < p.n_in) ? __ldg(p.x + q) : 0.f;
            v.y = 0.f;
        } else if (p.mode == 3) {
            v = __ldg(reinterpret_cast<const float2*>(p.x) + j);
        } else if (p.mode == 2) {
            v = make_float2(0.f, 0.f);
            if (j < p.n_in) v = cmul(__ldg(reinterpret_cast<const float2*>(p.x) + j), cconj(__ldg(p.chirp + j)));
        } else if (p.mode == 4) {   // containerFFT[1:upCoeff:end] .= in, times conj(b[j])
            v = make_float2(0.f, 0.f);
            if (j < p.n_tot) {
                const int64_t q = j / p.up;
                if (q * p.up == j && q < p.n_in) { const float2 b = __ldg(p.chirp + j); const float xv = __ldg(p.x + q); v = make_float2(xv * b.x, -xv * b.y); }
            }
        } else if (p.mode == 5) {
            v = make_float2(0.f, 0.f);
            if (j < p.n_tot) v = cmul(p.W[j], cconj(__ldg(p.chirp + j)));
        } else if (2 * j + 1 < p.n_valid) v = __ldg(reinterpret_cast<const float2*>(p.x) + j);
        else { v.x = (2 * j < p.n_valid) ? __ldg(p.x + 2 * j) : 0.f; v.y = 0.f; }
        sm[lay(col, j1)] = v;
    }
    __syncthreads();
    fft_inplace<+1, true>(sm, lay, p.C, p.A, p.radA, p.twA, tid, kFftThreads);
    for (int e = tid; e < total; e += kFftThreads) {
        const int row = e / p.C, col = e - row * p.C;
        const int k1 = __ldg(p.revA + row);
        const int j2 = j2_0 + col;
        float2 v = sm[lay(col, row)];
        v = cmul(v, twiddle_n(p, 2 * (int64_t)j2 * k1));  // W_M^(j2 k1) = W_N^(2 j2 k1)
        p.T[(int64_t)row * p.B + j2] = v;
    }
}

// ------------------------------------------------------------- k_fft_mid --
// CTA c handles frequency rows k1 = c and A - c (c = 0 and c = A/2 are self-paired).
__global__ void __launch_bounds__(kFftThreads) k_fft_mid(FftParams p) {
    extern __shared__ float2 sm[];
    const int tid = threadIdx.x;
    const int k1a = blockIdx.x;
    const int k1b = (p.A - k1a) % p.A;
    const bool self = (k1a == k1b);
    const int nrows = self ? 1 : 2;
    const int rowa = __ldg(p.posA + k1a), rowb = __ldg(p.posA + k1b);
    const int rs = row_padded(p.B);
    const RowLayout lay{rs};
    for (int e = tid; e < nrows * p.B; e += kFftThreads) {
        const int r = e / p.B, i = e - r * p.B;
        sm[lay(r, i)] = p.T[(int64_t)(r == 0 ? rowa : rowb) * p.B + i];
    }
    __syncthreads();
    fft_inplace<+1, false>(sm, lay, nrows, p.B, p.radB, p.twB, tid, kFftThreads);

    // real-input unpack, |X|^2, Hermitian repack (pairs k <-> M-k), scaled by 1/M
    const float sc = 0.5f * p.inv_scale;
    if (p.mode == 3) {   // 10*log10.(abs2.(fftshift(fft(ss))))  (GetSpectrum.jl:28)
        for (int e = tid; e < nrows * p.B; e += kFftThreads) {
            const int r = e / p.B, pos = e - r * p.B;
            const int64_t k = (int64_t)(r == 0 ? k1a : k1b) + (int64_t)p.A * __ldg(p.revB + pos);
            const float2 z = sm[lay(r, pos)];
            const float P = z.x * z.x + z.y * z.y;
            int64_t i = k + p.M / 2;
            if (i >= p.M) i -= p.M;
            p.out[i] = p.log_scale ? 10.0f * log10f(P) : P;
        }
        return;
    }
    if (p.mode == 1 || p.mode == 2 || p.mode == 4 || p.mode == 5) {
        // inFFT[n] = inFFT[n] * H[n]: ComplexF32 * ComplexF64 in Float64, rounded to ComplexF32 (Resampler.jl:51-53);
        // the 1/M of the scaled inverse plan is applied here
        for (int e = tid; e < nrows * p.B; e += kFftThreads) {
            const int r = e / p.B, pos = e - r * p.B;
            const int64_t k = (int64_t)(r == 0 ? k1a : k1b) + (int64_t)p.A * __ldg(p.revB + pos);
            const double2 h = __ldg(p.Hd + k);
            const float2 z = sm[lay(r, pos)];
            const double a = (double)z.x, b = (double)z.y;
            sm[lay(r, pos)] = make_float2((float)(a * h.x - b * h.y) * p.inv_scale, (float)(a * h.y + b * h.x) * p.inv_scale);
        }
    } else if (!self) {
        for (int pos = tid; pos < p.B; pos += kFftThreads) {
            const int k2 = __ldg(p.revB + pos);
            const int pos2 = __ldg(p.posB + (p.B - 1 - k2));
            const int64_t k = (int64_t)k1a + (int64_t)p.A * k2;
            const float2 zk = sm[lay(0, pos)], zm = sm[lay(1, pos2)];
            const float2 E = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
            const float2 O = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
            const float2 w = twiddle_n(p, k);
            const float2 wO = cmul(w, O);
            const float2 xp = cadd(E, wO), xm = csub(E, wO);
            const float P = fmaf(xp.x, xp.x, xp.y * xp.y), Pm = fmaf(xm.x, xm.x, xm.y * xm.y);
            const float S = (P + Pm) * sc, D = (P - Pm) * sc;
            // Y[k] = S + i conj(w) D ; Y[M-k] = S + i w D
            sm[lay(0, pos)] = make_float2(S + w.y * D, w.x * D);
            sm[lay(1, pos2)] = make_float2(S - w.y * D, w.x * D);
        }
    } else if (k1a == 0) {
        for (int k2 = tid; k2 <= p.B / 2; k2 += kFftThreads) {
            const int pos = __ldg(p.posB + k2);
            if (k2 == 0) {
                const float2 z = sm[lay(0, pos)];
                const float P0 = (z.x + z.y) * (z.x + z.y), PM = (z.x - z.y) * (z.x - z.y);
                sm[lay(0, pos)] = make_float2((P0 + PM) * sc, (P0 - PM) * sc);
                continue;
            }
            const int pos2 = __ldg(p.posB + (p.B - k2));
            const int64_t k = (int64_t)p.A * k2;
            const float2 zk = sm[lay(0, pos)], zm = sm[lay(0, pos2)];
            const float2 E = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
            const float2 O = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
            const float2 w = twiddle_n(p, k);
            const float2 wO = cmul(w, O);
            const float2 xp = cadd(E, wO), xm = csub(E, wO);
            const float P = fmaf(xp.x, xp.x, xp.y * xp.y), Pm = fmaf(xm.x, xm.x, xm.y * xm.y);
            const float S = (P + Pm) * sc, D = (P - Pm) * sc;
            sm[lay(0, pos)] = make_float2(S + w.y * D, w.x * D);
            if (pos2 != pos) sm[lay(0, pos2)] = make_float2(S - w.y * D, w.x * D);
        }
    } else {  // k1 = A/2: k2 <-> B-1-k2
        for (int k2 = tid; k2 < p.B / 2; k2 += kFftThreads) {
            const int pos = __ldg(p.posB + k2), pos2 = __ldg(p.posB + (p.B - 1 - k2));
            const int64_t k = (int64_t)k1a + (int64_t)p.A * k2;
            const float2 zk = sm[lay(0, pos)], zm = sm[lay(0, pos2)];
            const float2 E = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
            const float2 O = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
            const float2 w = twiddle_n(p, k);
            const float2 wO = cmul(w, O);
            const float2 xp = cadd(E, wO), xm = csub(E, wO);
            const float P = fmaf(xp.x, xp.x, xp.y * xp.y), Pm = fmaf(xm.x, xm.x, xm.y * xm.y);
            const float S = (P + Pm) * sc, D = (P - Pm) * sc;
            sm[lay(0, pos)] = make_float2(S + w.y * D, w.x * D);
            sm[lay(0, pos2)] = make_float2(S - w.y * D, w.x * D);
        }
    }
    __syncthreads();
    fft_inplace<-1, false>(sm, lay, nrows, p.B, p.radB, p.twB, tid, kFftThreads);
    // U[row][j2] = y * W_M^(-j2 k1)
    for (int e = tid; e < nrows * p.B; e += kFftThreads) {
        const int r = e / p.B, j2 = e - r * p.B;
        const int k1 = r == 0 ? k1a : k1b;
        float2 v = sm[lay(r, j2)];
        v = cmul(v, cconj(twiddle_n(p, 2 * (int64_t)j2 * k1)));
        p.U[(int64_t)(r == 0 ? rowa : rowb) * p.B + j2] = v;
    }
}

// ----------------------------------------------------------- k_ifft_cols --
// U[row][j2] -> A-point inverse FFT over rows -> y[j1*B + j2] = r[2j] + i r[2j+1]
// -> epilogue over the requested lags.
__global__ void __launch_bounds__(kFftThreads) k_ifft_cols(FftParams p) {
    extern __shared__ float2 sm[];
    const int tid = threadIdx.x;
    const int j2_0 = blockIdx.x * p.C;
    const ColLayout lay{p.C};
    const int total = p.A * p.C;
    for (int e = tid; e < total; e += kFftThreads) {
        const int row = e / p.C, col = e - row * p.C;
        sm[lay(col, row)] = p.U[(int64_t)row * p.B + j2_0 + col];
    }
    __syncthreads();
    fft_inplace<-1, true>(sm, lay, p.C, p.A, p.radA, p.twA, tid, kFftThreads);
    for (int e = tid; e < total; e += kFftThreads) {
        const int j1 = e / p.C, col = e - j1 * p.C;
        const int64_t j = (int64_t)j1 * p.B + j2_0 + col;
        if (p.mode == 1) { p.out[j] = p.gain * sm[lay(col, j1)].x; continue; }  // out[n] = 2*upCoeff*real(outFFT[n])  (Resampler.jl:57-59)
        if (p.mode == 4) {
            if (j < p.n_tot) {
                const float2 X = cmul(sm[lay(col, j1)], cconj(__ldg(p.chirp + j)));      // DFT_N of the zero-stuffed input
                const double2 h = __ldg(p.filt + j);
                const double a = (double)X.x, b = (double)X.y;                            // ComplexF32 * ComplexF64 -> ComplexF32
                p.W[j] = make_float2((float)(a * h.x - b * h.y), -(float)(a * h.y + b * h.x));
            }
            continue;
        }
        if (p.mode == 5) {
            if (j < p.n_tot) {
                const float2 z = cmul(sm[lay(col, j1)], cconj(__ldg(p.chirp + j)));
                p.out[j] = p.gain * (z.x * p.inv_n);                                      // 2*upCoeff*real(ifft(.))
            }
            continue;
        }
        if (p.mode == 2) {
            if (j < p.n_in) {
                const float2 z = sm[lay(col, j1)];
                const float P = z.x * z.x + z.y * z.y;
                int64_t i = j + p.n_in / 2;
                if (i >= p.n_in) i -= p.n_in;
                p.out[i] = p.log_scale ? 10.0f * log10f(P) : P;
            }
            continue;
        }
        const int64_t m0 = 2 * j;
        if (m0 > p.m_hi || m0 + 1 < p.m_lo) continue;
        const float2 v = sm[lay(col, j1)];
        float a = v.x, b = v.y;
        if (!p.raw) {
            a = a * a; b = b * b;  // abs2 of the (real) correlation
            if (p.log_scale) { a = 10.0f * log10f(a); b = 10.0f * log10f(b); }
        }
        if (m0 >= p.m_lo && m0 <= p.m_hi) p.out[m0 - p.m_lo] = a;
        if (m0 + 1 >= p.m_lo && m0 + 1 <= p.m_hi) p.out[m0 + 1 - p.m_lo] = b;
    }
}

// circular lags of a length-n signal from the linear lags of its zero-padded transform
__global__ void __launch_bounds__(256) k_fold_lags(const float* __restrict__ lin, int64_t n, int64_t m_lo, int64_t count,
                                                   int log_scale, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= count) return;
    const int64_t m = m_lo + i;
    float r = lin[m];
    if (m > 0) r += lin[n - m];
    float v = r * r;
    if (log_scale) v = 10.0f * log10f(v);
    out[i] = v;
}


// ----------------------------------------------------------- k_spec_batch --
// getWelch / getWaterfall (src/GetSpectrum.jl:36-66): the signal is cut into segments of `len`
// samples (a power of two), each gets a forward FFT in shared memory (`rows` segments per pass).
//  mode 0 (waterfall): out[seg][i] = abs2(fftshift(fft(seg)))[i]
//  mode 1 (Welch):     every thread keeps the running sum over the CTA's segments of the bins it
//                      owns (segments added in order), partial[cta][i] -> k_welch_final
struct SpecParams {
    const float2* x;
    int len;
    Radices rad;
    const float2* tw;      // W_len^k
    const int* pos;        // frequency -> position after the in-place DIF
    int64_t nseg;
    int rows;              // segments transformed per pass
    int seg_per_cta;
    int mode;
    float* out;            // waterfall [nseg][len]
    float* partial;        // Welch [gridDim.x][len]
};
constexpr int kSpecThreads = 256;
constexpr int kSpecMaxBins = 32;  // len <= 8192

__global__ void __launch_bounds__(kSpecThreads) k_spec_batch(SpecParams p) {
    extern __shared__ float2 sm[];
    const int tid = threadIdx.x;
    const int rs = row_padded(p.len);
    const RowLayout lay{rs};
    const int64_t s_begin = (int64_t)blockIdx.x * p.seg_per_cta;
    const int64_t s_end = min(p.nseg, s_begin + p.seg_per_cta);
    const int half = p.len / 2;
    float acc[kSpecMaxBins];
#pragma unroll
    for (int q = 0; q < kSpecMaxBins; ++q) acc[q] = 0.f;
    for (int64_t g = s_begin; g < s_end; g += p.rows) {
        const int nb = (int)min((int64_t)p.rows, s_end - g);
        for (int e = tid; e < nb * p.len; e += kSpecThreads) {
            const int b = e / p.len, i = e - b * p.len;
            sm[lay(b, i)] = __ldg(p.x + (g + b) * p.len + i);
        }
        __syncthreads();
        fft_inplace<+1, false>(sm, lay, nb, p.len, p.rad, p.tw, tid, kSpecThreads);
        if (p.mode == 0) {
            for (int e = tid; e < nb * p.len; e += kSpecThreads) {
                const int b = e / p.len, i = e - b * p.len;
                const int k = i >= half ? i - half : i + half;       // fftshift, even length
                const float2 z = sm[lay(b, __ldg(p.pos + k))];
                p.out[(g + b) * p.len + i] = z.x * z.x + z.y * z.y;
            }
        } else {
#pragma unroll
            for (int q = 0; q < kSpecMaxBins; ++q) {
                const int i = q * kSpecThreads + tid;
                if (i < p.len) {
                    const int k = i >= half ? i - half : i + half;
                    const int ps = __ldg(p.pos + k);
                    for (int b = 0; b < nb; ++b) { const float2 z = sm[lay(b, ps)]; acc[q] += z.x * z.x + z.y * z.y; }
                }
            }
        }
        __syncthreads();
    }
    if (p.mode == 1) {
#pragma unroll
        for (int q = 0; q < kSpecMaxBins; ++q) {
            const int i = q * kSpecThreads + tid;
            if (i < p.len) p.partial[(int64_t)blockIdx.x * p.len + i] = acc[q];
        }
    }
}

// S = sum of the CTA partials in segment order; y = 10*log10.(fftshift(S))  (GetSpectrum.jl:49)
__global__ void __launch_bounds__(256) k_welch_final(const float* __restrict__ partial, int nparts, int len, float* __restrict__ y) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= len) return;
    float s = 0.f;
    for (int c = 0; c < nparts; ++c) s += partial[(int64_t)c * len + i];
    y[i] = 10.0f * log10f(s);
}

}  // namespace tsdr

#include "tsdr_fft_fast.cuh"
#include "tsdr_fft3.cuh"

using namespace tsdr;

struct tsdr_autocorr_plan {
    int device;
    cudaStream_t stream;
    bool own_stream;
    size_t n;        // signal length
    size_t N;        // transform length (power of two)
    bool fold;
    FftParams fp;
    size_t smem_cols, smem_mid;
    bool has_fast;
    FastKernels fast;
    bool has_fft3;
    Fft3Kernels f3;
    uint64_t launches;
    void* d_tables;  // one allocation for every table
    float2* d_T; float2* d_U;
    float* d_lin;    // fold path: linear lags 0..n
};

namespace tsdr {

static Radices make_radices(int len) {
    Radices r; r.n = 0;
    int bits = 0; while ((1 << bits) < len) ++bits;
    const int rem = bits % 4;
    if (rem) r.r[r.n++] = 1 << rem;
    for (int i = 0; i < bits / 4; ++i) r.r[r.n++] = 16;
    return r;
}

// position (after the in-place DIF) -> frequency index
static int dif_frequency(int pos, int len, const Radices& rad) {
    int k = 0, weight = 1, span = len;
    for (int i = 0; i < rad.n; ++i) {
        span /= rad.r[i];
        const int digit = (pos / span) % rad.r[i];
        k += digit * weight;
        weight *= rad.r[i];
    }
    return k;
}

static bool is_pow2(size_t v) { return v && !(v & (v - 1)); }

}  // namespace tsdr

extern "C" {

int tsdr_autocorr_plan_destroy(tsdr_autocorr_plan* p) {
    if (!p) return TSDR_OK;
    TSDR_DEVICE(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    cudaFree(p->d_tables); cudaFree(p->d_T); cudaFree(p->d_U); cudaFree(p->d_lin);
    if (p->own_stream && p->stream) cudaStreamDestroy(p->stream);
    delete p;
    return TSDR_OK;
}

int tsdr_autocorr_plan_create(tsdr_autocorr_plan** out, int device, size_t n, void* stream) {
    TSDR_REQUIRE(out, "out is NULL");
    *out = nullptr;
    TSDR_REQUIRE(n >= 2, "autocorrelation needs at least 2 samples");
    TSDR_REQUIRE(n <= ((size_t)1 << 27), "signal too long (max 2^27 samples)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error("no CUDA device available; libtempest_b200 has no CPU fallback"); return TSDR_ERR_CUDA; }
    TSDR_REQUIRE(device >= 0 && device < ndev, "device %d out of range", device);
    tsdr_autocorr_plan* p = new (std::nothrow) tsdr_autocorr_plan();
    if (!p) return TSDR_ERR_NOMEM;
    memset(p, 0, sizeof(*p));
    p->device = device; p->n = n;
    // direct circular transform for powers of two >= 64, zero-pad + fold otherwise
    if (is_pow2(n) && n >= 64) { p->N = n; p->fold = false; }
    else { size_t N = 64; while (N < 2 * n) N <<= 1; p->N = N; p->fold = true; }
    const int64_t M = (int64_t)(p->N / 2);
    int B = M >= ((int64_t)1 << 24) ? 8192 : 4096;
    if ((int64_t)B > M / 2) B = (int)(M / 2);
    const int A = (int)(M / B);
    int C = 8;
    while (C > 1 && (size_t)col_padded(A * C) * sizeof(float2) > 200 * 1024) C >>= 1;
    if (C > B) C = B;
    FftParams& fp = p->fp;
    fp.A = A; fp.B = B; fp.C = C; fp.M = M;
    fp.radA = make_radices(A); fp.radB = make_radices(B);
    fp.inv_scale = 1.0f / (float)M;
    fp.lo_bits = 12;
    const size_t LO = (size_t)1 << fp.lo_bits;
    const size_t HI = (p->N + LO - 1) / LO;
    // host tables (double precision phases rounded once to float)
    std::vector<float2> twA(A), twB(B), wlo(LO), whi(HI);
    std::vector<int> revA(A), posA(A), revB(B), posB(B);
    const double PI2 = 6.283185307179586476925286766559;
    for (int k = 0; k < A; ++k) { twA[k].x = (float)cos(PI2 * k / A); twA[k].y = (float)-sin(PI2 * k / A); }
    for (int k = 0; k < B; ++k) { twB[k].x = (float)cos(PI2 * k / B); twB[k].y = (float)-sin(PI2 * k / B); }
    for (size_t t = 0; t < LO; ++t) { wlo[t].x = (float)cos(PI2 * (double)t / (double)p->N); wlo[t].y = (float)-sin(PI2 * (double)t / (double)p->N); }
    for (size_t t = 0; t < HI; ++t) { const double ph = PI2 * (double)(t * LO) / (double)p->N; whi[t].x = (float)cos(ph); whi[t].y = (float)-sin(ph); }
    for (int i = 0; i < A; ++i) { revA[i] = dif_frequency(i, A, fp.radA); posA[revA[i]] = i; }
    for (int i = 0; i < B; ++i) { revB[i] = dif_frequency(i, B, fp.radB); posB[revB[i]] = i; }

    int rc = TSDR_OK;
    DeviceScope scope;
    if ((rc = scope.enter(device))) { delete p; return rc; }
    if (e == cudaSuccess) {
        if (stream) { p->stream = (cudaStream_t)stream; p->own_stream = false; }
        else { e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking); p->own_stream = true; }
    }
    const size_t bytes = (A + B + LO + HI) * sizeof(float2) + (size_t)(2 * A + 2 * B) * sizeof(int);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_tables, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_T, (size_t)M * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_U, (size_t)M * sizeof(float2));
    if (e == cudaSuccess && p->fold) e = cudaMalloc(&p->d_lin, (n + 2) * sizeof(float));
    if (e == cudaSuccess) {
        char* base = (char*)p->d_tables;
        auto put = [&](const void* src, size_t sz) { void* dst = base; if (e == cudaSuccess) e = cudaMemcpy(dst, src, sz, cudaMemcpyHostToDevice); base += sz; return dst; };
        fp.twA = (const float2*)put(twA.data(), A * sizeof(float2));
        fp.twB = (const float2*)put(twB.data(), B * sizeof(float2));
        fp.wlo = (const float2*)put(wlo.data(), LO * sizeof(float2));
        fp.whi = (const float2*)put(whi.data(), HI * sizeof(float2));
        fp.revA = (const int*)put(revA.data(), A * sizeof(int));
        fp.posA = (const int*)put(posA.data(), A * sizeof(int));
        fp.revB = (const int*)put(revB.data(), B * sizeof(int));
        fp.posB = (const int*)put(posB.data(), B * sizeof(int));
    }
    fp.T = p->d_T; fp.U = p->d_U;
    p->smem_cols = (size_t)col_padded(A * C) * sizeof(float2);
    p->smem_mid = (size_t)2 * row_padded(B) * sizeof(float2);
    if (e == cudaSuccess) e = allow_max_dynamic_smem(k_fft_cols);
    if (e == cudaSuccess) e = allow_max_dynamic_smem(k_ifft_cols);
    if (e == cudaSuccess) e = allow_max_dynamic_smem(k_fft_mid);
    {
        int loga = 0, logb = 0;
        while ((1 << loga) < A) ++loga;
        while ((1 << logb) < B) ++logb;
        int logN = 0;
        while (((size_t)1 << logN) < p->N) ++logN;
        // TSDR_FFT_TWO_LEVEL=1 (read once, here) keeps a plan on the two-level kernels: tools/fft_variants.py times both
        p->has_fft3 = !getenv("TSDR_FFT_TWO_LEVEL") && find_fft3(logN, logb, &p->f3);
        if (p->has_fft3 && e == cudaSuccess) {
            e = allow_max_dynamic_smem(p->f3.p1);
            if (e == cudaSuccess) e = allow_max_dynamic_smem(p->f3.p1_padded);
            if (e == cudaSuccess) e = allow_max_dynamic_smem(p->f3.p5);
            if (e == cudaSuccess) e = allow_max_dynamic_smem(p->f3.p2);
            if (e == cudaSuccess) e = allow_max_dynamic_smem(p->f3.p4);
            if (e == cudaSuccess) e = allow_max_dynamic_smem(p->f3.p3);
        }
        p->has_fast = find_fast(loga, logb, &p->fast);
        if (p->has_fast && e == cudaSuccess) {
            e = allow_max_dynamic_smem(p->fast.cols);
            if (e == cudaSuccess) e = allow_max_dynamic_smem(p->fast.cols_padded);
            if (e == cudaSuccess) e = allow_max_dynamic_smem(p->fast.icols);
            if (e == cudaSuccess) e = allow_max_dynamic_smem(p->fast.mid);
        }
    }
    if (e != cudaSuccess) rc = cuda_fail(e, "tsdr_autocorr_plan_create", __FILE__, __LINE__);
    if (rc != TSDR_OK) { tsdr_autocorr_plan_destroy(p); return rc; }
    *out = p;
    return TSDR_OK;
}

int tsdr_autocorr_plan_exec(tsdr_autocorr_plan* p, const float* x_dev, size_t index_min, size_t index_max,
                            int log_scale, float* out_dev) {
    TSDR_REQUIRE(p && x_dev && out_dev, "NULL argument");
    TSDR_REQUIRE(index_min >= 1 && index_max >= index_min, "need 1 <= indexMin <= indexMax");
    if (index_max > p->n) { set_error("BoundsError: indexMax %zu beyond the %zu-point correlation", index_max, p->n); return TSDR_ERR_BOUNDS; }
    TSDR_REQUIRE((reinterpret_cast<uintptr_t>(x_dev) & 7) == 0, "input must be 8-byte aligned");
    TSDR_DEVICE(p->device);
    FftParams fp = p->fp;
    fp.x = x_dev; fp.n_valid = (int64_t)p->n; fp.log_scale = log_scale;
    if (p->fold) { fp.out = p->d_lin; fp.m_lo = 0; fp.m_hi = (int64_t)p->n; fp.raw = 1; }
    else { fp.out = out_dev; fp.m_lo = (int64_t)index_min - 1; fp.m_hi = (int64_t)index_max - 1; fp.raw = 0; }
    cudaStream_t st = p->stream;
    if (p->has_fft3 && (reinterpret_cast<uintptr_t>(x_dev) & 15) == 0) {
        if (p->n == p->N) p->f3.p1<<<p->f3.grid_p1, p->f3.threads, p->f3.smem_p1, st>>>(fp);
        else p->f3.p1_padded<<<p->f3.grid_p1, p->f3.threads, p->f3.smem_p1, st>>>(fp);
        // P2..P5 as programmatic dependent launches (tsdr_fft3.cuh: pdl_wait); TSDR_FFT_PDL=0 launches them the ordinary way
        const char* pe = getenv("TSDR_FFT_PDL");
        const bool pdl = !(pe && pe[0] == '0');
        auto launch = [&](void (*k)(FftParams), int grid, size_t smem) {
            if (!pdl) { k<<<grid, p->f3.threads, smem, st>>>(fp); return; }
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)p->f3.threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, k, fp);
        };
        launch(p->f3.p2, p->f3.grid_p2, p->f3.smem_p2);
        launch(p->f3.p3, p->f3.grid_p3, p->f3.smem_p3);
        launch(p->f3.p4, p->f3.grid_p2, p->f3.smem_p2);
        launch(p->f3.p5, p->f3.grid_p1, p->f3.smem_p1);
        p->launches += 2;
    } else if (p->has_fast && (reinterpret_cast<uintptr_t>(x_dev) & 15) == 0) {
        const int grid_cols = fp.B >> p->fast.logc;
        if (p->n == p->N) p->fast.cols<<<grid_cols, kFastThreads, p->fast.smem_cols, st>>>(fp);
        else p->fast.cols_padded<<<grid_cols, kFastThreads, p->fast.smem_cols, st>>>(fp);
        p->fast.mid<<<fp.A / 2 + 1, kFastThreads, p->fast.smem_mid, st>>>(fp);
        p->fast.icols<<<grid_cols, kFastThreads, p->fast.smem_cols, st>>>(fp);
    } else {
        k_fft_cols<<<fp.B / fp.C, kFftThreads, p->smem_cols, st>>>(fp);
        k_fft_mid<<<fp.A / 2 + 1, kFftThreads, p->smem_mid, st>>>(fp);
        k_ifft_cols<<<fp.B / fp.C, kFftThreads, p->smem_cols, st>>>(fp);
    }
    p->launches += 3;
    if (p->fold) {
        const int64_t count = (int64_t)(index_max - index_min + 1);
        k_fold_lags<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(p->d_lin, (int64_t)p->n, (int64_t)index_min - 1, count, log_scale, out_dev);
        p->launches += 1;
    }
    TSDR_CUDA(cudaGetLastError());
    return TSDR_OK;
}

int tsdr_autocorr_plan_launch_count(tsdr_autocorr_plan* p, uint64_t* count) {
    TSDR_REQUIRE(p && count, "NULL argument");
    *count = p->launches;
    return TSDR_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ upsampler (R3) --
namespace tsdr {
// host-side radix-2 FFT in double (power-of-two lengths), used once at init to build H
static void host_fft(std::vector<double>& re, std::vector<double>& im, bool inverse) {
    const size_t n = re.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
    }
    const double PI = 3.14159265358979323846264338327950288;
    for (size_t len = 2; len <= n; len <<= 1) {
        const double ang = (inverse ? 2.0 : -2.0) * PI / (double)len;
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const double wr = cos(ang * (double)k), wi = sin(ang * (double)k);
                const size_t a = i + k, b = i + k + len / 2;
                const double tr = re[b] * wr - im[b] * wi, ti = re[b] * wi + im[b] * wr;
                re[b] = re[a] - tr; im[b] = im[a] - ti;
                re[a] += tr; im[a] += ti;
            }
    }
    if (inverse) for (size_t i = 0; i < n; ++i) { re[i] /= (double)n; im[i] /= (double)n; }
}
}  // namespace tsdr

namespace tsdr {
// exp(i pi j^2 / N) with the phase reduced exactly (j^2 mod 2N in 128-bit integers)
static void chirp_phase(size_t j, size_t N, double& c, double& sn) {
    const unsigned long long q = (unsigned long long)(((unsigned __int128)j * j) % (2 * (unsigned __int128)N));
    const long double ph = 3.14159265358979323846264338327950288L * (long double)q / (long double)N;
    c = (double)cosl(ph); sn = (double)sinl(ph);
}
// DFT of any length in double: radix-2 for powers of two, chirp-z (Bluestein) on the next power of two otherwise
static void host_dft_any(std::vector<double>& re, std::vector<double>& im, bool inverse) {
    const size_t N = re.size();
    if (is_pow2(N)) { host_fft(re, im, inverse); return; }
    size_t M = 2; while (M < 2 * N - 1) M <<= 1;
    std::vector<double> ar(M, 0.0), ai(M, 0.0), br(M, 0.0), bi(M, 0.0), cr(N), ci(N);
    const double sgn = inverse ? -1.0 : 1.0;   // inverse: conjugate chirp
    for (size_t j = 0; j < N; ++j) {
        double c, sn; chirp_phase(j, N, c, sn); sn *= sgn;
        cr[j] = c; ci[j] = sn;
        ar[j] = re[j] * c + im[j] * sn; ai[j] = im[j] * c - re[j] * sn;    // x[j] * conj(b[j])
        br[j] = c; bi[j] = sn;
        if (j) { br[M - j] = c; bi[M - j] = sn; }
    }
    host_fft(ar, ai, false); host_fft(br, bi, false);
    for (size_t k = 0; k < M; ++k) { const double r = ar[k] * br[k] - ai[k] * bi[k], i = ar[k] * bi[k] + ai[k] * br[k]; ar[k] = r; ai[k] = i; }
    host_fft(ar, ai, true);
    const double sc = inverse ? 1.0 / (double)N : 1.0;
    for (size_t k = 0; k < N; ++k) {                                       // * conj(b[k])
        re[k] = (ar[k] * cr[k] + ai[k] * ci[k]) * sc;
        im[k] = (ai[k] * cr[k] - ar[k] * ci[k]) * sc;
    }
}
}  // namespace tsdr

struct tsdr_upsampler {
    tsdr_autocorr_plan* plan;  // complex engine: M = buffer_size * up points, or the chirp-z size Mb >= 2M-1 when M is not 2^k
    size_t n_in, M;
    int up;
    bool bluestein;
    size_t Mb;
    double2* d_H;
    float* d_in; float* d_out;
    float2* d_chirp; double2* d_Hb; float2* d_W;   // chirp-z route only
    std::vector<double>* H;    // host copy, interleaved
};

extern "C" {

int tsdr_upsampler_destroy(tsdr_upsampler* u) {
    if (!u) return TSDR_OK;
    DeviceScope scope;
    if (u->plan) { scope.enter(u->plan->device); cudaStreamSynchronize(u->plan->stream); }
    cudaFree(u->d_H); cudaFree(u->d_in); cudaFree(u->d_out);
    cudaFree(u->d_chirp); cudaFree(u->d_Hb); cudaFree(u->d_W);
    tsdr_autocorr_plan_destroy(u->plan);
    delete u->H;
    delete u;
    return TSDR_OK;
}

int tsdr_upsampler_create(size_t buffer_size, int up_coeff, tsdr_upsampler** out) {
    TSDR_REQUIRE(out, "out is NULL");
    *out = nullptr;
    TSDR_REQUIRE(buffer_size >= 1 && up_coeff >= 1, "bufferSize and upCoeff must be positive");
    const size_t M = buffer_size * (size_t)up_coeff;
    TSDR_REQUIRE(M >= 2, "init_resampler: bufferSize*upCoeff must be at least 2");
    // the reference takes any length (FFTW plans any N, src/Resampler.jl:26-40): powers of two in [32, 2^24] go straight
    // through the engine, every other length as two chirp-z transforms on the next power of two >= 2N-1
    const bool direct = is_pow2(M) && M >= 32;
    size_t Mb = M;
    if (!direct) { Mb = 32; while (Mb < 2 * M - 1) Mb <<= 1; }
    if (Mb > ((size_t)1 << 24)) {
        set_error("init_resampler: bufferSize*upCoeff = %zu needs a %zu-point transform; the engine stops at 2^24", M, Mb);
        return TSDR_ERR_UNSUPPORTED;
    }
    TSDR_TIER1_DEVICE(); int rc = TSDR_OK;
    tsdr_upsampler* u = new (std::nothrow) tsdr_upsampler();
    if (!u) return TSDR_ERR_NOMEM;
    memset(u, 0, sizeof(*u));
    u->n_in = buffer_size; u->M = M; u->up = up_coeff; u->bluestein = !direct; u->Mb = Mb;
    if ((rc = tsdr_autocorr_plan_create(&u->plan, current_device(), 2 * Mb, nullptr))) { tsdr_upsampler_destroy(u); return rc; }
    // initLPF (src/Resampler.jl:83-99): brick-wall magnitude, linear phase rounded to integers,
    // ifft, Blackman window, fft, (-1)^k.  (The reference's ifft runs in Float32; here it is
    // evaluated in double and rounded to Float32, a <= 1e-7 relative difference in h.)
    std::vector<double> re(M, 0.0), im(M, 0.0);
    const int64_t bound = round_even((double)M / (double)up_coeff / 2.0);
    const double gd = -((double)M - 1.0) / 2.0;
    for (size_t k = 0; k < M && (int64_t)k < bound; ++k) {
        const double puls = (double)((2.0L * 3.14159265358979323846264338327950288L * (long double)k) / (long double)M);
        const double th = gd * puls;
        re[k] = nearbyint(cos(th)); im[k] = nearbyint(sin(th));
    }
    host_dft_any(re, im, true);
    for (size_t k = 0; k < M; ++k) {
        const double x = (double)k / (double)(M - 1) - 0.5;
        const double w = 0.42 + 0.5 * cos(2.0 * 3.14159265358979323846 * x) + 0.08 * cos(4.0 * 3.14159265358979323846 * x);
        re[k] = (double)(float)re[k] * w; im[k] = (double)(float)im[k] * w;
    }
    host_dft_any(re, im, false);
    u->H = new std::vector<double>(2 * M);
    for (size_t k = 0; k < M; ++k) { const double sg = (k & 1) ? -1.0 : 1.0; (*u->H)[2 * k] = sg * re[k]; (*u->H)[2 * k + 1] = sg * im[k]; }
    cudaError_t e = cudaMalloc(&u->d_H, M * sizeof(double2));
    if (e == cudaSuccess) e = cudaMalloc(&u->d_in, (buffer_size + 4) * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&u->d_out, M * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(u->d_H, u->H->data(), M * sizeof(double2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && u->bluestein) {
        // b[j] = exp(i pi j^2 / N) (Float32 on the device) and FFT_Mb of its wrapped extension (double), as getSpectrum does
        std::vector<float2> chirp(M);
        std::vector<double> br(Mb, 0.0), bi(Mb, 0.0);
        for (size_t j = 0; j < M; ++j) {
            double c, sn; chirp_phase(j, M, c, sn);
            chirp[j] = make_float2((float)c, (float)sn);
            br[j] = c; bi[j] = sn;
            if (j) { br[Mb - j] = c; bi[Mb - j] = sn; }
        }
        host_fft(br, bi, false);
        std::vector<double2> Hb(Mb);
        for (size_t k = 0; k < Mb; ++k) Hb[k] = make_double2(br[k], bi[k]);
        e = cudaMalloc(&u->d_chirp, M * sizeof(float2));
        if (e == cudaSuccess) e = cudaMalloc(&u->d_Hb, Mb * sizeof(double2));
        if (e == cudaSuccess) e = cudaMalloc(&u->d_W, M * sizeof(float2));
        if (e == cudaSuccess) e = cudaMemcpy(u->d_chirp, chirp.data(), M * sizeof(float2), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(u->d_Hb, Hb.data(), Mb * sizeof(double2), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) { tsdr_upsampler_destroy(u); return cuda_fail(e, "tsdr_upsampler_create", __FILE__, __LINE__); }
    *out = u;
    return TSDR_OK;
}

int tsdr_upsampler_get_filter(tsdr_upsampler* u, double* H_interleaved) {
    TSDR_REQUIRE(u && H_interleaved, "NULL argument");
    memcpy(H_interleaved, u->H->data(), 2 * u->M * sizeof(double));
    return TSDR_OK;
}

int tsdr_upsampler_apply_f32(tsdr_upsampler* u, float* out, size_t n_out, const float* in, size_t n_in) {
    TSDR_REQUIRE(u && out && in, "NULL argument");
    // the reference's @assert (src/Resampler.jl:47): input length must match the init size
    TSDR_REQUIRE(n_in == u->n_in, "Size of input %zu should match size used during init %zu", n_in, u->n_in);
    TSDR_REQUIRE(n_out >= u->M, "output holds %zu samples, need bufferSize*upCoeff = %zu", n_out, u->M);
    tsdr_autocorr_plan* p = u->plan;
    TSDR_DEVICE(p->device);
    cudaStream_t st = p->stream;
    TSDR_CUDA(cudaMemcpyAsync(u->d_in, in, n_in * sizeof(float), cudaMemcpyHostToDevice, st));
    FftParams fp = p->fp;
    fp.x = u->d_in; fp.n_in = (int64_t)u->n_in; fp.up = u->up;
    fp.gain = (float)(2 * u->up); fp.out = u->d_out; fp.n_valid = 0;
    const int passes = u->bluestein ? 2 : 1;
    for (int pass = 0; pass < passes; ++pass) {
        if (!u->bluestein) { fp.mode = 1; fp.Hd = u->d_H; }
        else {
            fp.mode = 4 + pass; fp.Hd = u->d_Hb; fp.chirp = u->d_chirp; fp.filt = u->d_H; fp.W = u->d_W;
            fp.n_tot = (int64_t)u->M; fp.inv_n = 1.0f / (float)u->M;
        }
        k_fft_cols<<<fp.B / fp.C, kFftThreads, p->smem_cols, st>>>(fp);
        k_fft_mid<<<fp.A / 2 + 1, kFftThreads, p->smem_mid, st>>>(fp);
        k_ifft_cols<<<fp.B / fp.C, kFftThreads, p->smem_cols, st>>>(fp);
        p->launches += 3;
    }
    TSDR_CUDA(cudaGetLastError());
    TSDR_CUDA(cudaMemcpyAsync(out, u->d_out, u->M * sizeof(float), cudaMemcpyDeviceToHost, st));
    TSDR_CUDA(cudaStreamSynchronize(st));
    return TSDR_OK;
}

int tsdr_autocorr_out_len(size_t len, double Fs, double min_delay, double max_delay, size_t* out_len) {
    TSDR_REQUIRE(out_len, "out_len is NULL");
    const int64_t index_min = 1 + round_even(min_delay * Fs);  // Autocorrelations.jl:24
    const int64_t index_max = round_even(max_delay * Fs);      // Autocorrelations.jl:25
    TSDR_REQUIRE(index_min >= 1 && index_max >= index_min, "empty lag range (indexMin %lld, indexMax %lld)", (long long)index_min, (long long)index_max);
    size_t n = (size_t)(2 * index_max);
    if (len < n) n = len;                                       // :27
    if ((size_t)index_max > n) { *out_len = 0; set_error("BoundsError: signal of %zu samples shorter than indexMax %lld", len, (long long)index_max); return TSDR_ERR_BOUNDS; }
    *out_len = (size_t)(index_max - index_min + 1);
    return TSDR_OK;
}

int tsdr_autocorr_f32(const float* x, size_t len, double Fs, double min_delay, double max_delay,
                      int log_scale, float* out, size_t* out_len) {
    TSDR_REQUIRE(x && out, "NULL buffer");
    size_t L = 0;
    int rc = tsdr_autocorr_out_len(len, Fs, min_delay, max_delay, &L);
    if (out_len) *out_len = L;
    if (rc) return rc;
    const int64_t index_min = 1 + round_even(min_delay * Fs), index_max = round_even(max_delay * Fs);
    size_t n = (size_t)(2 * index_max);
    if (len < n) n = len;
    TSDR_TIER1_DEVICE();
    tsdr_autocorr_plan* plan = nullptr;
    if ((rc = tsdr_autocorr_plan_create(&plan, current_device(), n, nullptr))) return rc;
    void *d_x = nullptr, *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_x, (n + 2) * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, L * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_x, x, n * sizeof(float), cudaMemcpyHostToDevice, plan->stream);
    if (e == cudaSuccess) {
        rc = tsdr_autocorr_plan_exec(plan, (const float*)d_x, (size_t)index_min, (size_t)index_max, log_scale, (float*)d_out);
        if (rc == TSDR_OK) e = cudaMemcpyAsync(out, d_out, L * sizeof(float), cudaMemcpyDeviceToHost, plan->stream);
        if (rc == TSDR_OK && e == cudaSuccess) e = cudaStreamSynchronize(plan->stream);
    }
    if (e != cudaSuccess && rc == TSDR_OK) rc = cuda_fail(e, "tsdr_autocorr_f32", __FILE__, __LINE__);
    cudaFree(d_x); cudaFree(d_out);
    tsdr_autocorr_plan_destroy(plan);
    return rc;
}


// ---------------------------------------------------------------- GetSpectrum.jl --
// getSpectrum(fs, sig; N) (src/GetSpectrum.jl:21-30): y = 10*log10.(abs2.(fftshift(fft(sig[1:N])))) for a
// ComplexF32 signal.  Powers of two >= 32 go straight through the complex engine; every other
// length is evaluated as a chirp-z convolution on the next power of two >= 2N-1.
int tsdr_get_spectrum_f32(const float* sig_iq, size_t N, int log_scale, float* y) {
    TSDR_REQUIRE((sig_iq && y) || N == 0, "NULL argument");
    if (N == 0) return TSDR_OK;
    TSDR_REQUIRE(N <= ((size_t)1 << 23), "getSpectrum: at most 2^23 samples");
    TSDR_TIER1_DEVICE(); int rc = TSDR_OK;
    const bool direct = is_pow2(N) && N >= 32;
    size_t M = N;
    if (!direct) { M = 32; while (M < 2 * N - 1) M <<= 1; }
    tsdr_autocorr_plan* plan = nullptr;
    if ((rc = tsdr_autocorr_plan_create(&plan, current_device(), 2 * M, nullptr))) return rc;
    void *d_x = nullptr, *d_y = nullptr, *d_chirp = nullptr, *d_H = nullptr;
    cudaError_t e = cudaMalloc(&d_x, N * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc(&d_y, N * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_x, sig_iq, N * sizeof(float2), cudaMemcpyHostToDevice, plan->stream);
    FftParams fp = plan->fp;
    fp.x = (const float*)d_x; fp.n_in = (int64_t)N; fp.out = (float*)d_y; fp.log_scale = log_scale; fp.n_valid = 0;
    if (direct) fp.mode = 3;
    else {
        // b[j] = exp(i pi j^2 / N), phase from j^2 mod 2N so that it keeps full precision for large j
        std::vector<float2> chirp(N);
        std::vector<double> re(M, 0.0), im(M, 0.0);
        const double PI = 3.14159265358979323846264338327950288;
        for (size_t j = 0; j < N; ++j) {
            const unsigned long long q = (unsigned long long)(((unsigned __int128)j * j) % (2 * (unsigned __int128)N));
            const double ph = PI * (double)q / (double)N;
            const double c = cos(ph), sn = sin(ph);
            chirp[j] = make_float2((float)c, (float)sn);
            re[j] = c; im[j] = sn;
            if (j) { re[M - j] = c; im[M - j] = sn; }
        }
        host_fft(re, im, false);
        std::vector<double2> H(M);
        for (size_t k = 0; k < M; ++k) H[k] = make_double2(re[k], im[k]);
        if (e == cudaSuccess) e = cudaMalloc(&d_chirp, N * sizeof(float2));
        if (e == cudaSuccess) e = cudaMalloc(&d_H, M * sizeof(double2));
        if (e == cudaSuccess) e = cudaMemcpy(d_chirp, chirp.data(), N * sizeof(float2), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_H, H.data(), M * sizeof(double2), cudaMemcpyHostToDevice);
        fp.mode = 2; fp.chirp = (const float2*)d_chirp; fp.Hd = (const double2*)d_H;
    }
    if (e == cudaSuccess) {
        cudaStream_t st = plan->stream;
        k_fft_cols<<<fp.B / fp.C, kFftThreads, plan->smem_cols, st>>>(fp);
        k_fft_mid<<<fp.A / 2 + 1, kFftThreads, plan->smem_mid, st>>>(fp);
        if (!direct) k_ifft_cols<<<fp.B / fp.C, kFftThreads, plan->smem_cols, st>>>(fp);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(y, d_y, N * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (e != cudaSuccess) rc = cuda_fail(e, "tsdr_get_spectrum_f32", __FILE__, __LINE__);
    cudaFree(d_x); cudaFree(d_y); cudaFree(d_chirp); cudaFree(d_H);
    tsdr_autocorr_plan_destroy(plan);
    return rc;
}

namespace tsdr {
// getWelch (mode 1) / getWaterfall (mode 0): nbSeg = len / sizeFFT segments, the tail is dropped (:39, :55)
static int spec_segments(const float* sig_iq, size_t len, int size_fft, int mode, float* out) {
    TSDR_REQUIRE(size_fft >= 2 && size_fft <= 8192 && is_pow2((size_t)size_fft),
                 "sizeFFT must be a power of two in [2, 8192] (got %d)", size_fft);
    const int64_t nseg = (int64_t)(len / (size_t)size_fft);
    TSDR_REQUIRE((sig_iq && out) || nseg == 0, "NULL argument");
    TSDR_TIER1_DEVICE(); int rc = TSDR_OK;
    SpecParams sp;
    sp.len = size_fft; sp.rad = make_radices(size_fft); sp.nseg = nseg; sp.mode = mode;
    sp.rows = std::max(1, 8192 / size_fft);
    std::vector<float2> tw(size_fft);
    std::vector<int> pos(size_fft);
    const double PI2 = 6.283185307179586476925286766559;
    for (int k = 0; k < size_fft; ++k) { tw[k].x = (float)cos(PI2 * k / size_fft); tw[k].y = (float)-sin(PI2 * k / size_fft); }
    for (int i = 0; i < size_fft; ++i) pos[dif_frequency(i, size_fft, sp.rad)] = i;
    // enough CTAs to fill the GPU, whole passes of `rows` segments each
    int64_t ctas = std::min<int64_t>((nseg + sp.rows - 1) / sp.rows, 148 * 4);
    if (ctas < 1) ctas = 1;
    sp.seg_per_cta = (int)(((nseg + ctas - 1) / ctas + sp.rows - 1) / sp.rows * sp.rows);
    if (sp.seg_per_cta < sp.rows) sp.seg_per_cta = sp.rows;
    ctas = nseg ? (nseg + sp.seg_per_cta - 1) / sp.seg_per_cta : 1;
    const size_t smem = (size_t)sp.rows * row_padded(size_fft) * sizeof(float2);
    const size_t n_used = (size_t)nseg * size_fft;
    const size_t out_floats = mode == 0 ? n_used : (size_t)size_fft;
    void *d_x = nullptr, *d_out = nullptr, *d_part = nullptr, *d_tab = nullptr;
    cudaError_t e = cudaMalloc(&d_x, std::max<size_t>(n_used, 1) * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, std::max<size_t>(out_floats, 1) * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_part, (size_t)ctas * size_fft * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_tab, size_fft * (sizeof(float2) + sizeof(int)));
    if (e == cudaSuccess && n_used) e = cudaMemcpy(d_x, sig_iq, n_used * sizeof(float2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_tab, tw.data(), size_fft * sizeof(float2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy((char*)d_tab + size_fft * sizeof(float2), pos.data(), size_fft * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = allow_max_dynamic_smem(k_spec_batch);
    if (e == cudaSuccess) {
        sp.x = (const float2*)d_x; sp.tw = (const float2*)d_tab; sp.pos = (const int*)((char*)d_tab + size_fft * sizeof(float2));
        sp.out = (float*)d_out; sp.partial = (float*)d_part;
        if (nseg) k_spec_batch<<<(unsigned)ctas, kSpecThreads, smem>>>(sp);
        else e = cudaMemset(d_part, 0, (size_t)size_fft * sizeof(float));   // no segment: S stays zeros -> -Inf dB, as the reference
        if (mode == 1) k_welch_final<<<(size_fft + 255) / 256, 256>>>((const float*)d_part, (int)ctas, size_fft, (float*)d_out);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess && out_floats) e = cudaMemcpy(out, d_out, out_floats * sizeof(float), cudaMemcpyDeviceToHost);
    }
    if (e != cudaSuccess) rc = cuda_fail(e, "spec_segments", __FILE__, __LINE__);
    cudaFree(d_x); cudaFree(d_out); cudaFree(d_part); cudaFree(d_tab);
    return rc;
}
}  // namespace tsdr

// getWelch(fe, sig; sizeFFT) (src/GetSpectrum.jl:36-52): y[sizeFFT] = 10*log10.(fftshift(sum_n abs2.(fft(seg_n))))
int tsdr_get_welch_f32(const float* sig_iq, size_t len, int size_fft, float* y) {
    return spec_segments(sig_iq, len, size_fft, 1, y);
}

// getWaterfall(fe, sig; sizeFFT) (src/GetSpectrum.jl:54-66): sMatrix[:, n] = abs2.(fftshift(fft(seg_n))), column-major
// sizeFFT x nbSeg (Float32 here; the reference widens the same Float32 values into a Float64 matrix)
int tsdr_get_waterfall_f32(const float* sig_iq, size_t len, int size_fft, float* s_matrix) {
    return spec_segments(sig_iq, len, size_fft, 0, s_matrix);
}

}  // extern "C"
