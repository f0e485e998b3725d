// tsdr_fft3.cuh -- three-level autocorrelation FFT for the large sizes (included by tsdr_fft.cu
// after tsdr_fft_fast.cuh).
//
// The two-level kernels read and write 32-byte pieces at 32 KB stride in their column passes
// (1.2 TB/s effective).  Here the M = N/2 complex points are viewed as D[a][b][c]
// (NA x NB x NC, c fastest) and every pass moves >= 128 contiguous bytes:
//   P1  FFT over a for tiles of C1 consecutive (b,c)   * W_M^(ka r)       x  -> T   (natural ka)
//   P2  FFT over b for tiles of C2 consecutive c       * W_M'^(kb c)      T in place (natural kb)
//   P3  rows over c: FFT, real-input unpack, |X|^2, Hermitian repack, inverse FFT, * conj W_M'^(kb c)
//   P4  inverse FFT over kb, * conj W_M^(ka r)                            T in place
//   P5  inverse FFT over ka, epilogue 10 log10(r^2) over the requested lags          T -> out
// Rows/columns are stored in NATURAL frequency order between passes (the digit reversal of the
// in-place DIF/DIT butterflies stays inside shared memory), and P2..P4 work in place, so the
// whole working set is one M-point buffer (64 MB at n = 2^24) that fits the 126 MB L2.
// Twiddles: the long-transform factors W_N^t live in two global tables (lo/hi split); inside the kernels every
// per-element twiddle is the product of two entries of small per-CTA shared-memory tables (index split such as
// ka = 16 kh + kl, filled with a few hundred look-ups per CTA), and the second column of each pair is the first
// times the row's step -- the global look-ups per element were the main stall of the first version.
//
// Frequency index: k = ka + NA (kb + NB kc).  Mirror M-k used by the real-input unpack:
//   ka > 0          : (NA-ka, NB-1-kb, NC-1-kc)   -> row rho = ka NB + kb  pairs with  NA NB + NB - 1 - rho
//   ka = 0, kb > 0  : (0, NB-kb, NC-1-kc)
//   ka = 0, kb = 0  : (0, 0, (NC-kc) mod NC)
#pragma once

namespace tsdr {

// Programmatic dependent launch between the five passes: every pass lets its successor start launching as soon as all
// of its own CTAs are resident (launch_dependents at the top), and the successor blocks at pdl_wait() -- after it has
// built its per-CTA twiddle tables, before its first read of T -- until this grid has completed and flushed.  The
// successor's table prologue (a few hundred dependent look-ups, ~1.5 us) and its launch latency then overlap the
// partially filled last wave of this pass.  Both instructions are no-ops in a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <int LOGNA, int LOGNB, int LOGNC, int LOGC1, int LOGC2, int LOGR, int LOGTAB, int LOGNT>
struct Fft3 {
    static constexpr int NA = 1 << LOGNA, NB = 1 << LOGNB, NC = 1 << LOGNC;
    static constexpr int LOGNBC = LOGNB + LOGNC;
    static constexpr int R = 1 << LOGR;
    // padded row length with stride = 8 (mod 16) elements: two rows then cover all 32 banks in the radix-16 stage
    static constexpr int kRowStride = ((NC + (NC >> 4)) & 15) == 8 ? NC + (NC >> 4) : ((NC + (NC >> 4) + 15) & ~15) + 8;
    static constexpr size_t smem_p1 = (size_t)col_padded_ct<LOGC1>(NA << LOGC1) * sizeof(float2);
    static constexpr size_t smem_p2 = (size_t)col_padded_ct<LOGC2>(NB << LOGC2) * sizeof(float2);
    static constexpr size_t smem_p3 = (size_t)2 * R * kRowStride * sizeof(float2);
    static constexpr int grid_p1 = (1 << LOGNBC) >> LOGC1;
    static constexpr int grid_p2 = NA << (LOGNC - LOGC2);
    static constexpr int n_regular = (NB / 2) * (NA - 1) / R;          // row pairs with ka > 0
    static constexpr int n_special = (NB / 2 + 1 + R - 1) / R;         // rows with ka = 0: t = 0 .. NB/2
    static constexpr int grid_p3 = n_regular + n_special;
};

// W_len^k, k in [0, len), copied from the plan's longer table into shared memory: the butterflies'
// twiddle reads then are LDS instead of global loads in the middle of every dependent chain
template <int LOGLEN, int LOGTAB, int NT>
__device__ __forceinline__ void load_twiddles(float2* tw_s, const float2* __restrict__ tab, int tid) {
    for (int k = tid; k < (1 << LOGLEN); k += NT) tw_s[k] = __ldg(tab + ((size_t)k << (LOGTAB - LOGLEN)));
}

// ------------------------------------------------------------------------------- P1 --
template <class F, bool PADDED>
__global__ void __launch_bounds__(F::NT_v, F::MINB_v) k3_p1(FftParams p) {
    extern __shared__ __align__(16) float2 sm[];
    constexpr int LOGC = F::LOGC1_v, C = 1 << LOGC, HALF = C / 2, NA = F::NA_v;
    __shared__ float2 tw_s[NA];
    __shared__ float2 tw_step[NA];   // W_M^ka: the twiddle of column r + 1 is the twiddle of column r times this
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x << LOGC;
    const ColLayoutCt<LOGC> lay;
    pdl_launch_dependents();
    load_twiddles<F::LOGNA_v, F::LOGTAB_v, F::NT_v>(tw_s, p.twB, tid);
    for (int k = tid; k < NA; k += F::NT_v) tw_step[k] = twiddle_n(p, 2 * (int64_t)k);
    // W_M^(ka r) for the CTA's columns r = r0 + 2 j as the product of two small tables over ka = 16 kh + kl
    __shared__ float2 tw_kh[NA / 16][HALF];
    __shared__ float2 tw_kl[16][HALF];
    for (int e = tid; e < (NA / 16 + 16) * HALF; e += F::NT_v) {
        const int q = e / HALF, jj = e - q * HALF;
        const int64_t r = r0 + 2 * jj;
        if (q < NA / 16) tw_kh[q][jj] = twiddle_n(p, 2 * (int64_t)(16 * q) * r);
        else tw_kl[q - NA / 16][jj] = twiddle_n(p, 2 * (int64_t)(q - NA / 16) * r);
    }
    static_assert(F::NT_v % HALF == 0, "each thread keeps one column pair");
    const int c2 = (tid & (HALF - 1)) * 2;       // this thread's column pair in every loop below
#pragma unroll
    for (int a = tid / HALF; a < NA; a += F::NT_v / HALF) {
        const int64_t j = ((int64_t)a << F::LOGNBC_v) + r0 + c2;
        float4 v;
        if (!PADDED || 2 * j + 3 < p.n_valid) v = __ldg(reinterpret_cast<const float4*>(p.x) + (j >> 1));
        else {
            v.x = 2 * j < p.n_valid ? __ldg(p.x + 2 * j) : 0.f;
            v.y = 2 * j + 1 < p.n_valid ? __ldg(p.x + 2 * j + 1) : 0.f;
            v.z = 2 * j + 2 < p.n_valid ? __ldg(p.x + 2 * j + 2) : 0.f;
            v.w = 0.f;
        }
        *reinterpret_cast<float4*>(&sm[lay(c2, a)]) = v;
    }
    __syncthreads();
    fft_fwd_ct<F::LOGNA_v, 0, true, LOGC, ColLayoutCt<LOGC>, F::LOGNA_v, F::NT_v>(sm, lay, tw_s, tid);
    // two adjacent columns per thread: one table look-up W_M^(ka r); the neighbour's twiddle is that times the
    // row's step W_M^ka from shared memory (the second look-up per pair was the main stall of this loop;
    // four columns per look-up brought nothing more)
    const int r = r0 + c2;
#pragma unroll 4
    for (int rho = tid / HALF; rho < NA; rho += F::NT_v / HALF) {
        const int ka = digit_rev_ct<F::LOGNA_v>(rho);
        const float4 sv = *reinterpret_cast<const float4*>(&sm[lay(c2, rho)]);
        const float2 wa = cmul(tw_kh[ka >> 4][c2 >> 1], tw_kl[ka & 15][c2 >> 1]);                 // W_M^(ka r) = W_N^(2 ka r)
        const float2 a = cmul(make_float2(sv.x, sv.y), wa);
        const float2 b = cmul(make_float2(sv.z, sv.w), cmul(wa, tw_step[ka]));
        reinterpret_cast<float4*>(p.T)[(((int64_t)ka << F::LOGNBC_v) + r) >> 1] = make_float4(a.x, a.y, b.x, b.y);
    }
}

// --------------------------------------------------------------------------- P2 / P4 --
template <class F, int DIR>
__global__ void __launch_bounds__(F::NT_v, F::MINB_v) k3_p24(FftParams p) {
    extern __shared__ __align__(16) float2 sm[];
    constexpr int LOGC = F::LOGC2_v, C = 1 << LOGC, HALF = C / 2, NB = F::NB_v;
    const int tid = threadIdx.x;
    const int ka = blockIdx.x >> (F::LOGNC_v - LOGC);
    const int c0 = (blockIdx.x & ((1 << (F::LOGNC_v - LOGC)) - 1)) << LOGC;
    const int64_t base = ((int64_t)ka << F::LOGNBC_v) + c0;
    const ColLayoutCt<LOGC> lay;
    float4* T4 = reinterpret_cast<float4*>(p.T);
    __shared__ float2 tw_s[NB];
    __shared__ float2 tw_step[NB];   // forward: W_M'^kb, the step from column c to c + 1
                                     // inverse: conj W_M^(ka b NC), the row part of conj W_M^(ka r), r = b NC + c
    __shared__ float2 tw_col[HALF];  // inverse: conj W_M^(ka (c0 + 2 j)), its column part
    pdl_launch_dependents();
    load_twiddles<F::LOGNB_v, F::LOGTAB_v, F::NT_v>(tw_s, p.twB, tid);
    __shared__ float2 tw_kh[NB / 16][HALF];   // forward: W_M'^(kb c) for c = c0 + 2 j, kb = 16 kh + kl, as a product
    __shared__ float2 tw_kl[16][HALF];
    if (DIR > 0) {
        for (int k = tid; k < NB; k += F::NT_v) tw_step[k] = twiddle_n(p, (int64_t)k << (F::LOGNA_v + 1));
        for (int e = tid; e < (NB / 16 + 16) * HALF; e += F::NT_v) {
            const int q = e / HALF, jj = e - q * HALF;
            const int64_t c = c0 + 2 * jj;
            if (q < NB / 16) tw_kh[q][jj] = twiddle_n(p, ((int64_t)(16 * q) * c) << (F::LOGNA_v + 1));
            else tw_kl[q - NB / 16][jj] = twiddle_n(p, ((int64_t)(q - NB / 16) * c) << (F::LOGNA_v + 1));
        }
    } else {
        for (int k = tid; k < NB; k += F::NT_v) tw_step[k] = cconj(twiddle_n(p, (2 * (int64_t)ka * k) << F::LOGNC_v));
        if (tid < HALF) tw_col[tid] = cconj(twiddle_n(p, 2 * (int64_t)ka * (c0 + 2 * tid)));
    }
    const float2 step_inv = cconj(twiddle_n(p, 2 * (int64_t)ka));   // inverse: conj W_M^ka, the same for the whole CTA
    static_assert(F::NT_v % HALF == 0, "each thread keeps one column pair");
    const int c2 = (tid & (HALF - 1)) * 2;       // this thread's column pair in every loop below
    pdl_wait();                                  // T is the previous pass's output
#pragma unroll
    for (int i = tid / HALF; i < NB; i += F::NT_v / HALF) {   // forward: i = b ; inverse: i = kb
        const int pos = DIR > 0 ? i : digit_pos_ct<F::LOGNB_v>(i);
        *reinterpret_cast<float4*>(&sm[lay(c2, pos)]) = T4[(base + ((int64_t)i << F::LOGNC_v) + c2) >> 1];
    }
    __syncthreads();
    if (DIR > 0) fft_fwd_ct<F::LOGNB_v, 0, true, LOGC, ColLayoutCt<LOGC>, F::LOGNB_v, F::NT_v>(sm, lay, tw_s, tid);
    else fft_inv_ct<F::LOGNB_v, CtPlan<F::LOGNB_v>::nst - 1, true, LOGC, ColLayoutCt<LOGC>, F::LOGNB_v, F::NT_v>(sm, lay, tw_s, tid);
#pragma unroll 4
    for (int pos = tid / HALF; pos < NB; pos += F::NT_v / HALF) {
        const float4 sv = *reinterpret_cast<const float4*>(&sm[lay(c2, pos)]);
        float2 a, b;
        int i;
        if (DIR > 0) {   // position pos holds kb; stage-2 twiddle W_M'^(kb c) = W_N^(2 NA kb c)
            i = digit_rev_ct<F::LOGNB_v>(pos);
            const float2 wa = cmul(tw_kh[i >> 4][c2 >> 1], tw_kl[i & 15][c2 >> 1]);
            a = cmul(make_float2(sv.x, sv.y), wa);
            b = cmul(make_float2(sv.z, sv.w), cmul(wa, tw_step[i]));
        } else {         // position pos holds b; undo the stage-1 twiddle W_M^(ka r), r = b NC + c
            i = pos;
            const float2 wa = cmul(tw_step[i], tw_col[c2 >> 1]);
            a = cmul(make_float2(sv.x, sv.y), wa);
            b = cmul(make_float2(sv.z, sv.w), cmul(wa, step_inv));
        }
        T4[(base + ((int64_t)i << F::LOGNC_v) + c2) >> 1] = make_float4(a.x, a.y, b.x, b.y);
    }
}

// ------------------------------------------------------------------------------- P3 --
template <class F>
__global__ void __launch_bounds__(F::NT_v, F::MINB_v) k3_p3(FftParams p) {
    extern __shared__ __align__(16) float2 sm[];
    constexpr int NA = F::NA_v, NB = F::NB_v, NC = F::NC_v, R = F::R_v, LOGNC = F::LOGNC_v, LOGNB = F::LOGNB_v, LOGNA = F::LOGNA_v;
    const int tid = threadIdx.x;
    const bool special = (int)blockIdx.x >= F::n_regular_v;
    const int t0 = (special ? (int)blockIdx.x - F::n_regular_v : (int)blockIdx.x) * R;
    const RowLayoutCt lay{F::kRowStride_v};
    float4* T4 = reinterpret_cast<float4*>(p.T);
    __shared__ float2 tw_s[NC];
    __shared__ float2 tw_c[NC];        // W_N^(kc NA NB): with tw_row, the unpack twiddle W_N^k of k = ka + NA kb + NA NB kc
    __shared__ float2 tw_row[R];       // W_N^(ka + NA kb) of slot s (row A)
    __shared__ float2 tw_rstep[2 * R]; // conj W_M'^kb of smem row srow: the step from column c to c + 1 in the write-back
    __shared__ float2 tw_hi[2 * R][NC / 16];   // conj W_M'^(kb 16 ch) and
    __shared__ float2 tw_lo[2 * R][8];         // conj W_M'^(kb cl), cl even: the write-back twiddle of column 16 ch + cl
    pdl_launch_dependents();
    load_twiddles<LOGNC, F::LOGTAB_v, F::NT_v>(tw_s, p.twB, tid);
    // slot s in [0, R): rows (rowA, rowB); smem row s holds rowA, smem row R + s holds rowB
    auto rowA_of = [&](int s) { return special ? t0 + s : NB + t0 + s; };
    auto rowB_of = [&](int s) { return special ? ((NB - (t0 + s)) & (NB - 1)) : NA * NB - 1 - (t0 + s); };
    auto valid = [&](int s) { return !special || t0 + s <= NB / 2; };
    for (int k = tid; k < NC; k += F::NT_v) tw_c[k] = twiddle_n(p, (int64_t)k << (LOGNA + LOGNB));
    if (tid < R) {
        const int rowA = rowA_of(tid);
        tw_row[tid] = twiddle_n(p, (int64_t)(rowA >> LOGNB) + ((int64_t)(rowA & (NB - 1)) << LOGNA));
    }
    if (tid < 2 * R) {
        const int row = tid < R ? rowA_of(tid) : rowB_of(tid - R);
        tw_rstep[tid] = cconj(twiddle_n(p, (int64_t)(row & (NB - 1)) << (LOGNA + 1)));
    }
    for (int e = tid; e < 2 * R * (NC / 16 + 8); e += F::NT_v) {
        const int srow = e / (NC / 16 + 8), q = e - srow * (NC / 16 + 8);
        const int row = srow < R ? rowA_of(srow) : rowB_of(srow - R);
        const int64_t kb = row & (NB - 1);
        if (q < NC / 16) tw_hi[srow][q] = cconj(twiddle_n(p, (kb * 16 * q) << (LOGNA + 1)));
        else tw_lo[srow][q - NC / 16] = cconj(twiddle_n(p, (kb * 2 * (q - NC / 16)) << (LOGNA + 1)));
    }
    pdl_wait();                                // T is the previous pass's output
    if (!special) {
        // The regular CTAs (all but a handful): every slot is valid and row A never equals row B, and because the
        // thread count is a multiple of the row length each thread keeps ONE column (pair) for the whole kernel --
        // its digit-reversed positions and its column twiddle are computed once instead of per element.
        static_assert(F::NT_v % NC == 0 && F::NT_v % (NC / 2) == 0, "threads per CTA must be a multiple of the row length");
        constexpr int ROWS_PER_IT = F::NT_v / (NC / 2);          // rows covered by one pass of the load / store loops
        const int c2 = (tid & (NC / 2 - 1)) * 2;
        const int rsub = tid / (NC / 2);
        const int rowA0 = NB + t0, rowB0 = NA * NB - 1 - t0;     // rowA_of(s) = rowA0 + s, rowB_of(s) = rowB0 - s
#pragma unroll
        for (int srow = rsub; srow < 2 * R; srow += ROWS_PER_IT) {
            const int row = srow < R ? rowA0 + srow : rowB0 - (srow - R);
            const float4 v = T4[(((int64_t)row << LOGNC) + c2) >> 1];
            sm[lay(srow, c2)] = make_float2(v.x, v.y);
            sm[lay(srow, c2 + 1)] = make_float2(v.z, v.w);
        }
        __syncthreads();
        fft_fwd_ct<LOGNC, 0, false, F::LOGR_v + 1, RowLayoutCt, LOGNC, F::NT_v>(sm, lay, tw_s, tid);
        {
            const float sc = 0.5f * p.inv_scale;
            const int kc = tid & (NC - 1);
            const int posk = digit_pos_ct<LOGNC>(kc), posm = digit_pos_ct<LOGNC>(NC - 1 - kc);
            const float2 wc = tw_c[kc];
            for (int sl = tid >> LOGNC; sl < R; sl += F::NT_v >> LOGNC)
                mid_pair_w(sm[lay(sl, posk)], sm[lay(R + sl, posm)], cmul(tw_row[sl], wc), sc);
        }
        __syncthreads();
        fft_inv_ct<LOGNC, CtPlan<LOGNC>::nst - 1, false, F::LOGR_v + 1, RowLayoutCt, LOGNC, F::NT_v>(sm, lay, tw_s, tid);
#pragma unroll
        for (int srow = rsub; srow < 2 * R; srow += ROWS_PER_IT) {
            const int row = srow < R ? rowA0 + srow : rowB0 - (srow - R);
            const float2 wa = cmul(tw_hi[srow][c2 >> 4], tw_lo[srow][(c2 & 15) >> 1]);
            const float2 a = cmul(sm[lay(srow, c2)], wa);
            const float2 b = cmul(sm[lay(srow, c2 + 1)], cmul(wa, tw_rstep[srow]));
            T4[(((int64_t)row << LOGNC) + c2) >> 1] = make_float4(a.x, a.y, b.x, b.y);
        }
        return;
    }
    // ---- the special CTAs: rows with ka = 0, whose mirrors lie in the same row set
#pragma unroll
    for (int e = tid; e < 2 * R * (NC / 2); e += F::NT_v) {
        const int srow = e / (NC / 2), i2 = (e - srow * (NC / 2)) * 2;
        const int s = srow & (R - 1);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid(s)) {
            const int row = srow < R ? rowA_of(s) : rowB_of(s);
            v = T4[(((int64_t)row << LOGNC) + i2) >> 1];
        }
        sm[lay(srow, i2)] = make_float2(v.x, v.y);
        sm[lay(srow, i2 + 1)] = make_float2(v.z, v.w);
    }
    __syncthreads();
    fft_fwd_ct<LOGNC, 0, false, F::LOGR_v + 1, RowLayoutCt, LOGNC, F::NT_v>(sm, lay, tw_s, tid);
    const float sc = 0.5f * p.inv_scale;
    for (int e = tid; e < R * NC; e += F::NT_v) {
        const int s = e >> LOGNC, kc = e & (NC - 1);
        if (!valid(s)) continue;
        const int rowA = rowA_of(s), rowB = rowB_of(s);
        const float2 w = cmul(tw_row[s], tw_c[kc]);   // W_N^k, k = ka + NA kb + NA NB kc
        const int posk = digit_pos_ct<LOGNC>(kc);
        if (rowA != rowB) {
            mid_pair_w(sm[lay(s, posk)], sm[lay(R + s, digit_pos_ct<LOGNC>(NC - 1 - kc))], w, sc);
        } else if (rowA == 0) {       // (0,0,kc) <-> (0,0,(NC-kc) mod NC)
            if (kc > NC / 2) continue;
            if (kc == 0) {
                const float2 z = sm[lay(s, posk)];
                const float P0 = (z.x + z.y) * (z.x + z.y), PM = (z.x - z.y) * (z.x - z.y);
                sm[lay(s, posk)] = make_float2((P0 + PM) * sc, (P0 - PM) * sc);
            } else {
                const int pos2 = digit_pos_ct<LOGNC>(NC - kc);
                float2 a = sm[lay(s, posk)], b = sm[lay(s, pos2)];
                mid_pair_w(a, b, w, sc);
                sm[lay(s, posk)] = a;
                if (pos2 != posk) sm[lay(s, pos2)] = b;
            }
        } else {                      // row (0, NB/2): kc <-> NC-1-kc inside the row
            if (kc >= NC / 2) continue;
            mid_pair_w(sm[lay(s, posk)], sm[lay(s, digit_pos_ct<LOGNC>(NC - 1 - kc))], w, sc);
        }
    }
    __syncthreads();
    fft_inv_ct<LOGNC, CtPlan<LOGNC>::nst - 1, false, F::LOGR_v + 1, RowLayoutCt, LOGNC, F::NT_v>(sm, lay, tw_s, tid);
    // undo the stage-2 twiddle W_M'^(kb c) and write the rows back in place
    for (int e = tid; e < 2 * R * (NC / 2); e += F::NT_v) {
        const int srow = e / (NC / 2), c2 = (e - srow * (NC / 2)) * 2;
        const int s = srow & (R - 1);
        if (!valid(s)) continue;
        const int rowA = rowA_of(s), rowB = rowB_of(s);
        if (srow >= R && rowA == rowB) continue;
        const int row = srow < R ? rowA : rowB;
        const float2 wa = cmul(tw_hi[srow][c2 >> 4], tw_lo[srow][(c2 & 15) >> 1]);
        const float2 a = cmul(sm[lay(srow, c2)], wa);
        const float2 b = cmul(sm[lay(srow, c2 + 1)], cmul(wa, tw_rstep[srow]));
        T4[(((int64_t)row << LOGNC) + c2) >> 1] = make_float4(a.x, a.y, b.x, b.y);
    }
}

// ------------------------------------------------------------------------------- P5 --
template <class F>
__global__ void __launch_bounds__(F::NT_v, F::MINB_v) k3_p5(FftParams p) {
    extern __shared__ __align__(16) float2 sm[];
    constexpr int LOGC = F::LOGC1_v, C = 1 << LOGC, HALF = C / 2, NA = F::NA_v;
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x << LOGC;
    const ColLayoutCt<LOGC> lay;
    const float4* T4 = reinterpret_cast<const float4*>(p.T);
    __shared__ float2 tw_s[NA];
    pdl_launch_dependents();
    load_twiddles<F::LOGNA_v, F::LOGTAB_v, F::NT_v>(tw_s, p.twB, tid);
    static_assert(F::NT_v % HALF == 0, "each thread keeps one column pair");
    const int c2 = (tid & (HALF - 1)) * 2;
    pdl_wait();                                  // T is the previous pass's output
#pragma unroll
    for (int ka = tid / HALF; ka < NA; ka += F::NT_v / HALF) {
        *reinterpret_cast<float4*>(&sm[lay(c2, digit_pos_ct<F::LOGNA_v>(ka))]) = T4[(((int64_t)ka << F::LOGNBC_v) + r0 + c2) >> 1];
    }
    __syncthreads();
    fft_inv_ct<F::LOGNA_v, CtPlan<F::LOGNA_v>::nst - 1, true, LOGC, ColLayoutCt<LOGC>, F::LOGNA_v, F::NT_v>(sm, lay, tw_s, tid);
    const bool out_aligned = (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
#pragma unroll 4
    for (int a = tid / HALF; a < NA; a += F::NT_v / HALF) {
        const int64_t j = ((int64_t)a << F::LOGNBC_v) + r0 + c2;
        const int64_t m0 = 2 * j;   // y[j] = r[2j] + i r[2j+1]: two adjacent columns = four consecutive lags
        if (m0 > p.m_hi || m0 + 3 < p.m_lo) continue;
        const float4 sv = *reinterpret_cast<const float4*>(&sm[lay(c2, a)]);
        float o[4] = {sv.x, sv.y, sv.z, sv.w};
        if (!p.raw) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                o[t] = o[t] * o[t];  // abs2 of the (real) correlation
                if (p.log_scale) o[t] = db10_fast(o[t]);
            }
        }
        if (out_aligned && m0 >= p.m_lo && m0 + 3 <= p.m_hi && (((m0 - p.m_lo) & 3) == 0)) {
            *reinterpret_cast<float4*>(p.out + (m0 - p.m_lo)) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (m0 + t >= p.m_lo && m0 + t <= p.m_hi) p.out[m0 + t - p.m_lo] = o[t];
        }
    }
}

// ------------------------------------------------------------------------- dispatch --
struct Fft3Kernels {
    void (*p1)(FftParams);
    void (*p1_padded)(FftParams);
    void (*p2)(FftParams);
    void (*p3)(FftParams);
    void (*p4)(FftParams);
    void (*p5)(FftParams);
    int grid_p1, grid_p2, grid_p3, threads;
    size_t smem_p1, smem_p2, smem_p3;
};

// the kernels take the shape through a traits class with plain static members
template <int LOGNA, int LOGNB, int LOGNC, int LOGC1, int LOGC2, int LOGR, int LOGTAB, int LOGNT>
struct Fft3Traits {
    using G = Fft3<LOGNA, LOGNB, LOGNC, LOGC1, LOGC2, LOGR, LOGTAB, LOGNT>;
    static constexpr int LOGNA_v = LOGNA, LOGNB_v = LOGNB, LOGNC_v = LOGNC, LOGC1_v = LOGC1, LOGC2_v = LOGC2, LOGR_v = LOGR,
                         LOGTAB_v = LOGTAB, LOGNBC_v = LOGNB + LOGNC, NA_v = 1 << LOGNA, NB_v = 1 << LOGNB, NC_v = 1 << LOGNC,
                         R_v = 1 << LOGR, kRowStride_v = G::kRowStride, n_regular_v = G::n_regular, NT_v = 1 << LOGNT,
                         MINB_v = (3 * kFastThreads) >> LOGNT;
};

template <int LOGNA, int LOGNB, int LOGNC, int LOGC1, int LOGC2, int LOGR, int LOGTAB, int LOGNT>
static Fft3Kernels make_fft3() {
    using F = Fft3Traits<LOGNA, LOGNB, LOGNC, LOGC1, LOGC2, LOGR, LOGTAB, LOGNT>;
    using G = typename F::G;
    Fft3Kernels k;
    k.p1 = k3_p1<F, false>; k.p1_padded = k3_p1<F, true>;
    k.p2 = k3_p24<F, +1>; k.p4 = k3_p24<F, -1>;
    k.p3 = k3_p3<F>; k.p5 = k3_p5<F>;
    k.threads = F::NT_v;
    k.grid_p1 = G::grid_p1; k.grid_p2 = G::grid_p2; k.grid_p3 = G::grid_p3;
    k.smem_p1 = G::smem_p1; k.smem_p2 = G::smem_p2; k.smem_p3 = G::smem_p3;
    return k;
}

// shapes with a three-level build, keyed by log2 of the transform length N (real samples) and of
// the two-level B table length (the W_B^k table the plan already owns, reused with a stride)
static bool find_fft3(int logN, int logtab, Fft3Kernels* out) {
#define TSDR_FFT3(n, a, b, c, c1, c2, r, tab, nt) if (logN == n && logtab == tab) { *out = make_fft3<a, b, c, c1, c2, r, tab, nt>(); return true; }
    // Shapes measured on B200 (tools/fft_variants.py).  Tiles of 16 columns (128-byte row pieces, 32 KB of shared
    // memory) and 8 row pairs beat the 32-column / 32-pair tiles of the first version by 16-27 %: the per-CTA
    // load -> transform -> store sequence is latency bound, and twice as many half-sized CTAs overlap it better
    // and leave a shorter tail; 8-column tiles (64-byte pieces) lose again.  Against the two-level kernels:
    // 2^23 0.122 vs 0.145 ms, 2^24 0.223 vs 0.306, 2^25 0.450 vs 0.614, 2^26 1.06 vs 1.45; at 2^22 the two-level
    // kernels stay (0.072 ms both).  CTAs of 128 threads on the same tiles (two butterflies per thread and stage,
    // six CTAs per SM) are 10-16 % slower than 256 threads; min-blocks 4 or 5 (64 / 48 registers) do not help either.
    TSDR_FFT3(23, 8, 7, 7, 4, 4, 3, 12, 8)   // M = 2^22: what the GUI's 3e6 / 4e6-sample calls pad to
    TSDR_FFT3(24, 8, 7, 8, 4, 4, 3, 12, 8)   // M = 2^23: the benchmark size
    TSDR_FFT3(25, 8, 8, 8, 4, 4, 3, 13, 8)   // M = 2^24
    TSDR_FFT3(26, 9, 8, 8, 4, 4, 3, 13, 8)   // M = 2^25
#undef TSDR_FFT3
    return false;
}

}  // namespace tsdr
