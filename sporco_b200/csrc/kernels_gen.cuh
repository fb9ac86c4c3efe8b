// kernels_gen.cuh -- any-size row kernels (direct O(N^2) DFT along the row).
//
// The Stockham kernels cover power-of-two lengths.  Every other image size -- the reference
// accepts any (its own tests use 16x17 and 63x63, tests/admm/test_cbpdn.py:124-139, 203-225) --
// takes this path: same data layout, same fusion boundaries and the same arithmetic after the
// transform, with the transform itself evaluated directly from a twiddle table.  It is a
// completeness path for small problems, not a fast one; the column counterpart is k_col with
// N0 = 0 (col_dft_chunk in kernels.cuh).
#pragma once

#include "kernels.cuh"

namespace spcsc {

SPCSC_HD size_t gen_align16(size_t n) { return (n + 15) / 16 * 16; }

// Plan of the mixed-radix path for the row length (computed once per CTA by thread 0)
struct GenRowPlan {
    int rad[16];
    int nrad;
};
template <typename T>
SPCSC_DEV void gen_row_plan(GenRowPlan* pl, int N1, int fast) {
    if (threadIdx.x == 0) pl->nrad = fast ? gen_factor(N1, pl->rad) : 0;
    __syncthreads();
}

// Forward real DFT of TR rows held in shared memory (xs[r*N1 + n]) -> transposed slab store.  `work`: two
// complex buffers of TR*N1 each for the mixed-radix path (null: direct DFT).
template <typename T>
SPCSC_DEV void gen_rows_forward_store(const T* xs, C2<T>* SPCSC_RESTRICT Zt,
                                      const C2<T>* SPCSC_RESTRICT tw, int b, int m, int h0, int nrow,
                                      int N0, int N1, int M, C2<T>* work = nullptr,
                                      const GenRowPlan* pl = nullptr, int TRmax = 0) {
    const int N1f = N1 / 2 + 1;
    C2<T>* out = Zt + (((size_t)b * N1f) * M + m) * N0 + h0;
    const size_t wstride = (size_t)M * N0;
    if (work && pl && pl->nrad > 0) {
        C2<T>* wa = work;
        C2<T>* wb = work + (size_t)TRmax * N1;
        for (int e = threadIdx.x; e < nrow * N1; e += blockDim.x) wa[e] = mk<T>(xs[e], 0);
        __syncthreads();
        const C2<T>* res = gen_fft_batch<T, false>(wa, wb, tw, nrow, N1, pl->rad, pl->nrad);
        for (int e = threadIdx.x; e < nrow * N1f; e += blockDim.x) {
            const int wf = e / nrow, r = e - wf * nrow;
            out[wf * wstride + r] = res[(size_t)r * N1 + wf];
        }
        __syncthreads();
        return;
    }
    for (int e = threadIdx.x; e < nrow * N1f; e += blockDim.x) {
        const int wf = e / nrow, r = e - wf * nrow;
        const T* x = xs + (size_t)r * N1;
        T sr = 0, si = 0;
        int idx = 0;
        for (int n = 0; n < N1; ++n) {
            const C2<T> w = tw[idx];
            sr += x[n] * w.re;
            si += x[n] * w.im;
            idx += wf;
            if (idx >= N1) idx -= N1;
        }
        out[wf * wstride + r] = mk<T>(sr, si);
    }
}

// Gather TR rows of one (b, m) from the slab layout into shared memory (zs[r*N1f + wf]) and
// evaluate the c2r inverse: xs[r*N1 + n] = scale * irfft(zs[r])[n] * N1.
template <typename T>
SPCSC_DEV void gen_rows_inverse(C2<T>* zs, T* xs, const C2<T>* SPCSC_RESTRICT Zt,
                                const C2<T>* SPCSC_RESTRICT tw, int b, int m, int h0, int nrow, int N0,
                                int N1, int M, T scale, C2<T>* work = nullptr,
                                const GenRowPlan* pl = nullptr, int TRmax = 0) {
    const int N1f = N1 / 2 + 1;
    const C2<T>* in = Zt + (((size_t)b * N1f) * M + m) * N0 + h0;
    const size_t wstride = (size_t)M * N0;
    const bool even = (N1 % 2) == 0;
    if (work && pl && pl->nrad > 0) {
        // full Hermitian spectrum (c2r ignores Im X[0] and Im X[N1/2]), complex inverse, real part
        C2<T>* wa = work;
        C2<T>* wb = work + (size_t)TRmax * N1;
        for (int e = threadIdx.x; e < nrow * N1; e += blockDim.x) {
            const int kf = e / nrow, r = e - kf * nrow;
            const int ks = kf < N1f ? kf : N1 - kf;
            C2<T> z = in[ks * wstride + r];
            if (kf >= N1f) z = conj(z);
            if (ks == 0 || (even && ks == N1f - 1)) z.im = 0;
            wa[(size_t)r * N1 + kf] = z;
        }
        __syncthreads();
        const C2<T>* res = gen_fft_batch<T, true>(wa, wb, tw, nrow, N1, pl->rad, pl->nrad);
        for (int e = threadIdx.x; e < nrow * N1; e += blockDim.x) xs[e] = res[e].re * scale;
        __syncthreads();
        return;
    }
    for (int e = threadIdx.x; e < nrow * N1f; e += blockDim.x) {
        const int wf = e / nrow, r = e - wf * nrow;
        zs[(size_t)r * N1f + wf] = in[wf * wstride + r];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nrow * N1; e += blockDim.x) {
        const int r = e / N1, n = e - r * N1;
        const C2<T>* z = zs + (size_t)r * N1f;
        T s = z[0].re;                                      // c2r ignores Im X[0] (and Im X[N1/2])
        int idx = 0;
        for (int k = 1; k < N1f; ++k) {
            idx += n;
            if (idx >= N1) idx -= N1;
            const C2<T> w = tw[idx];                         // exp(-2 pi i n k / N1); need its conjugate
            const T re = z[k].re * w.re + z[k].im * w.im;    // Re(z * conj(w))
            s += (even && k == N1f - 1) ? re : (T)2 * re;
        }
        xs[e] = s * scale;
    }
    __syncthreads();
}

template <typename T>
SPCSC_GLOBAL void k_row_fwd_gen(const T* SPCSC_RESTRICT A, const T* SPCSC_RESTRICT B,
                                const AdmmState<T>* SPCSC_RESTRICT st, C2<T>* SPCSC_RESTRICT Zt,
                                const C2<T>* SPCSC_RESTRICT tw, int N0, int N1, int M, int TR, int fast) {
    if (st && st->stopped) return;
    SPCSC_DYN_SMEM(smem_raw);
    __shared__ GenRowPlan plan;
    gen_row_plan<T>(&plan, N1, fast);
    T* xs = reinterpret_cast<T*>(smem_raw);
    C2<T>* work = fast ? reinterpret_cast<C2<T>*>(smem_raw + gen_align16((size_t)TR * N1 * sizeof(T))) : nullptr;
    const int h0 = blockIdx.x * TR, m = blockIdx.y, b = blockIdx.z;
    const int nrow = (N0 - h0) < TR ? (N0 - h0) : TR;
    const T udiv = (st && B) ? st->udiv : (T)1;
    const size_t base = (((size_t)b * M + m) * N0 + h0) * N1;
    for (int e = threadIdx.x; e < nrow * N1; e += blockDim.x) {
        T v = A[base + e];
        if (B) v -= B[base + e] / udiv;
        xs[e] = v;
    }
    __syncthreads();
    gen_rows_forward_store<T>(xs, Zt, tw, b, m, h0, nrow, N0, N1, M, work, &plan, TR);
}

template <typename T>
SPCSC_GLOBAL void k_row_inv_gen(const C2<T>* SPCSC_RESTRICT Zt, T* SPCSC_RESTRICT X,
                                const C2<T>* SPCSC_RESTRICT tw, int N0, int N1, int M, int TR, T scale, int fast) {
    SPCSC_DYN_SMEM(smem_raw);
    __shared__ GenRowPlan plan;
    gen_row_plan<T>(&plan, N1, fast);
    const int N1f = N1 / 2 + 1;
    C2<T>* zs = reinterpret_cast<C2<T>*>(smem_raw);
    T* xs = reinterpret_cast<T*>(zs + (size_t)TR * N1f);
    C2<T>* work = fast ? reinterpret_cast<C2<T>*>(smem_raw + gen_align16((size_t)TR * N1f * sizeof(C2<T>) + (size_t)TR * N1 * sizeof(T))) : nullptr;
    const int h0 = blockIdx.x * TR, m = blockIdx.y, b = blockIdx.z;
    const int nrow = (N0 - h0) < TR ? (N0 - h0) : TR;
    gen_rows_inverse<T>(zs, xs, Zt, tw, b, m, h0, nrow, N0, N1, M, scale, work, &plan, TR);
    const size_t base = (((size_t)b * M + m) * N0 + h0) * N1;
    for (int e = threadIdx.x; e < nrow * N1; e += blockDim.x) X[base + e] = xs[e];
}

// Inverse rows + relaxation + prox + dual update + residual sums (any size, any Cx <= 4).
template <typename T, int CX>
SPCSC_GLOBAL void k_row_inv_prox_gen(const C2<T>* SPCSC_RESTRICT Zt, T* SPCSC_RESTRICT Y,
                                     T* SPCSC_RESTRICT U, const AdmmState<T>* SPCSC_RESTRICT st,
                                     AdmmParams<T> prm, WeightView<T> wl1, WeightView<T> wl21,
                                     double* SPCSC_RESTRICT acc, const C2<T>* SPCSC_RESTRICT tw, int N0,
                                     int N1, int M, int TR, T scale, int nonneg, int bnd0, int bnd1,
                                     int reg_on_y, int fast) {
    if (st->stopped) return;
    SPCSC_DYN_SMEM(smem_raw);
    __shared__ GenRowPlan plan;
    gen_row_plan<T>(&plan, N1, fast);
    const int N1f = N1 / 2 + 1;
    C2<T>* zs = reinterpret_cast<C2<T>*>(smem_raw);                     // [TR][N1f]
    T* xs = reinterpret_cast<T*>(zs + (size_t)TR * N1f);                // [CX][TR][N1]
    C2<T>* work = fast ? reinterpret_cast<C2<T>*>(smem_raw + gen_align16((size_t)TR * N1f * sizeof(C2<T>) + (size_t)CX * TR * N1 * sizeof(T))) : nullptr;
    const int h0 = blockIdx.x * TR, m = blockIdx.y, k = blockIdx.z;
    const bool ams = m >= prm.ams_m0;   // additive-mask-simulation map: no clipping, not part of RegL1
    const int nrow = (N0 - h0) < TR ? (N0 - h0) : TR;
    for (int c = 0; c < CX; ++c)
        gen_rows_inverse<T>(zs, xs + (size_t)c * TR * N1, Zt, tw, k * CX + c, m, h0, nrow, N0, N1, M,
                            scale, work, &plan, TR);
    const T rho = st->rho, udiv = st->udiv;
    const T lr = prm.lmbda / rho;
    const T mr = prm.joint ? prm.mu / rho : (T)0;
    const T rlx = prm.rlx;
    double sums[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int e = threadIdx.x; e < nrow * N1; e += blockDim.x) {
        const int r = e / N1, n = e - r * N1;
        const int h = h0 + r;
        T wv[CX], ax[CX], ue[CX], xv[CX], yp[CX], w1s[CX];
        T a2 = 0, g2 = 0;
        SPCSC_UNROLL
        for (int c = 0; c < CX; ++c) {
            const size_t off = ((((size_t)(k * CX + c) * M + m) * N0 + h) * N1 + n);
            const T x = xs[((size_t)c * TR + r) * N1 + n];
            const T y = Y[off], u = U[off] / udiv;
            const T w1 = wl1.p[(size_t)k * wl1.sk + (size_t)c * wl1.sc + (size_t)m * wl1.sm +
                              (size_t)h * wl1.s0 + (size_t)n * wl1.s1];
            const T axv = (rlx == (T)1) ? x : rlx * x + ((T)1 - rlx) * y;
            const T w = soft_threshold(axv + u, lr * w1);
            xv[c] = x; yp[c] = y; ax[c] = axv; ue[c] = u; wv[c] = w; w1s[c] = w1;
            a2 += w * w;
            if (!reg_on_y) {
                sums[ACC_L1] += ams ? 0 : (double)fabs(w1 * x);
                g2 += x * x;
            }
        }
        T fac = 1, w21 = 1;
        if (prm.joint) {
            w21 = wl21.p[(size_t)k * wl21.sk + (size_t)m * wl21.sm + (size_t)h * wl21.s0 +
                         (size_t)n * wl21.s1];
            const T a = sqrt(a2);
            const T bq = fmax((T)0, a - mr * w21);
            fac = (a != (T)0) ? bq / a : (T)0;
        }
        SPCSC_UNROLL
        for (int c = 0; c < CX; ++c) {
            const size_t off = ((((size_t)(k * CX + c) * M + m) * N0 + h) * N1 + n);
            T y = prm.joint ? fac * wv[c] : wv[c];
            if (nonneg && !ams && y < (T)0) y = (T)0;
            if (!ams && (h >= bnd0 || n >= bnd1)) y = (T)0;
            const T u = ue[c] + (ax[c] - y);
            const T x = xv[c];
            const T dr = x - y, ds = yp[c] - y;
            sums[ACC_X2] += (double)x * x;
            sums[ACC_Y2] += (double)y * y;
            sums[ACC_U2] += (double)u * u;
            sums[ACC_R2] += (double)dr * dr;
            sums[ACC_S2] += (double)ds * ds;
            if (reg_on_y) {
                sums[ACC_L1] += ams ? 0 : (double)fabs(w1s[c] * y);
                g2 += y * y;
            }
            Y[off] = y;
            U[off] = u;
        }
        if (prm.joint) sums[ACC_L21] += (double)(w21 * sqrt(g2));
    }
    if (prm.need_rsdl || prm.need_obj) {
        double* red = reinterpret_cast<double*>(smem_raw);
        block_accumulate_det<7>(sums, red, reinterpret_cast<unsigned long long*>(acc + ACC_N));
    }
}

// PGM proximal step (see k_row_inv_prox_fwd), any size.
template <typename T>
SPCSC_GLOBAL void k_row_inv_prox_fwd_gen(C2<T>* SPCSC_RESTRICT Vt, T* SPCSC_RESTRICT X, T thr_scale,
                                         WeightView<T> wl1, double* SPCSC_RESTRICT acc,
                                         const C2<T>* SPCSC_RESTRICT tw, int N0, int N1, int M, int Cx,
                                         int TR, T scale, int nonneg, int bnd0, int bnd1, int fast) {
    SPCSC_DYN_SMEM(smem_raw);
    __shared__ GenRowPlan plan;
    gen_row_plan<T>(&plan, N1, fast);
    const int N1f = N1 / 2 + 1;
    C2<T>* zs = reinterpret_cast<C2<T>*>(smem_raw);
    T* xs = reinterpret_cast<T*>(zs + (size_t)TR * N1f);
    C2<T>* work = fast ? reinterpret_cast<C2<T>*>(smem_raw + gen_align16((size_t)TR * N1f * sizeof(C2<T>) + (size_t)TR * N1 * sizeof(T))) : nullptr;
    const int h0 = blockIdx.x * TR, m = blockIdx.y, b = blockIdx.z;
    const int k = b / Cx, c = b - k * Cx;
    const int nrow = (N0 - h0) < TR ? (N0 - h0) : TR;
    gen_rows_inverse<T>(zs, xs, Vt, tw, b, m, h0, nrow, N0, N1, M, scale, work, &plan, TR);
    double sums[1] = {0.0};
    const size_t base = (((size_t)b * M + m) * N0 + h0) * N1;
    for (int e = threadIdx.x; e < nrow * N1; e += blockDim.x) {
        const int r = e / N1, n = e - r * N1;
        const int h = h0 + r;
        const T w1 = wl1.p[(size_t)k * wl1.sk + (size_t)c * wl1.sc + (size_t)m * wl1.sm +
                          (size_t)h * wl1.s0 + (size_t)n * wl1.s1];
        T x = soft_threshold(xs[e], thr_scale * w1);
        if (nonneg && x < (T)0) x = (T)0;
        if (h >= bnd0 || n >= bnd1) x = (T)0;
        sums[0] += (double)fabs(w1 * x);
        xs[e] = x;
        X[base + e] = x;
    }
    __syncthreads();
    gen_rows_forward_store<T>(xs, Vt, tw, b, m, h0, nrow, N0, N1, M, work, &plan, TR);
    double* red = reinterpret_cast<double*>(smem_raw);
    block_accumulate<1>(sums, red, acc + ACC_L1);
}

}  // namespace spcsc
