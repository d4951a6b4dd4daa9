// ransac_math.cuh -- the double-precision arithmetic of RANSAC for the fundamental matrix
// (SURVEY.md section 8, row f2), written so that every rounding happens where the reference's
// does.  Compiles for the device (the product) and for the host (test infrastructure only:
// oracle/ransac_hostcheck.cc checks it against the reference's own functions without a GPU).
//
// What it restates (reference, paths relative to /root/reference/src/mve):
//   math::matrix_svd for square inputs     math/matrix_svd.h:140-760 (Householder
//       bidiagonalisation :252-437, Golub-Kahan iteration :440-641, sign fix and sort
//       :626-641 and :747-759; the M < N case pads with zero rows :729-744)
//   internal::matrix_givens_rotation       math/matrix_qr.h:50-73
//   sfm::fundamental_8_point               sfm/fundamental.cc:78-110
//   sfm::enforce_fundamental_constraints   sfm/fundamental.cc:113-126
//   sfm::sampson_distance                  sfm/fundamental.cc:225-247
//
// Why the care: RANSAC keeps the first sample with the most inliers and an inlier is
// `sampson < threshold^2`, so a result equal to the reference's needs the same doubles.
// The reference is plain IEEE double arithmetic without fused multiply-add (x86-64 SSE2);
// on the device every product and sum therefore goes through __dmul_rn / __dadd_rn /
// __dsub_rn (which the compiler never contracts), and division and sqrt are IEEE in double.
// Sums run in the reference's order.  Terms the reference multiplies by the structural zeros
// and ones of its padded update matrices are skipped: adding an exact zero changes no partial
// sum (only the sign of an exact zero result can differ, which no later step observes).
// The left singular vectors are only accumulated when asked for (the 8-point solve ignores
// them and they feed back into nothing).  The iteration may stop at a fixed point (see
// SquareSvd::changed): the reference would only repeat the same no-op until its limit.
#pragma once

#include <cmath>

#if defined(__CUDACC__)
#define OSFM_HD __host__ __device__ __forceinline__
#else
#define OSFM_HD inline
#endif

namespace osfm {
namespace fmath {

#if defined(__CUDA_ARCH__)
OSFM_HD double mul(double a, double b) { return __dmul_rn(a, b); }
OSFM_HD double add(double a, double b) { return __dadd_rn(a, b); }
OSFM_HD double sub(double a, double b) { return __dsub_rn(a, b); }
#else
OSFM_HD double mul(double a, double b) { return a * b; }
OSFM_HD double add(double a, double b) { return a + b; }
OSFM_HD double sub(double a, double b) { return a - b; }
#endif

// the bit pattern of a double
#if defined(__CUDA_ARCH__)
OSFM_HD unsigned long long bits_of(double a) { return static_cast<unsigned long long>(__double_as_longlong(a)); }
#else
OSFM_HD unsigned long long bits_of(double a) {
    unsigned long long x;
    __builtin_memcpy(&x, &a, 8);
    return x;
}
#endif

constexpr double kSvdEpsilon = 1e-12;     // MATH_SVD_DEFAULT_ZERO_THRESHOLD, matrix_svd.h:30

// Where an N x N matrix lives: in the thread's own memory, or strided through a buffer shared
// by the threads of a CTA (element i of this thread at p[i * stride]).
template <int N>
struct LocalMatrix {
    double a[N * N];
    OSFM_HD double& at(int i) { return a[i]; }
};
struct StridedMatrix {
    double* p;
    int stride;
    OSFM_HD double& at(int i) { return p[i * stride]; }
};

// MATH_EPSILON_EQ(x, 0, eps), math/defines.h:96
OSFM_HD bool near_zero(double x, double eps) { return (sub(0.0, eps) <= x) && (x <= add(0.0, eps)); }

// Householder vector of `in` (matrix_svd.h:146-176); every entry is first divided by nf.
OSFM_HD void householder_vector(const double* in, int len, double* v, double* beta, double eps, double nf)
{
    double sigma = 0.0;
    for (int i = 1; i < len; ++i) {
        double const t = in[i] / nf;
        sigma = add(sigma, mul(t, t));
    }
    v[0] = 1.0;
    for (int i = 1; i < len; ++i) v[i] = in[i] / nf;
    if (near_zero(sigma, eps)) { *beta = 0.0; return; }
    double first = in[0] / nf;
    double const mu = sqrt(add(mul(first, first), sigma));
    if (first < eps) v[0] = sub(first, mu);
    else v[0] = (-sigma) / add(first, mu);
    first = v[0];
    double const f2 = mul(first, first);
    *beta = mul(2.0, f2) / add(sigma, f2);
    for (int i = 0; i < len; ++i) v[i] = v[i] / first;
}

// Givens coefficients (matrix_qr.h:50-73).  Written with selects instead of the reference's
// three-way branch: inside a warp every lane works on its own matrix, and a data-dependent
// branch in a loop body splits the warp for the rest of the loop (measured: 4.6 of 32 lanes
// active in the rotations).  Only the selected quotient is computed, so nothing is evaluated
// that the reference does not evaluate.
OSFM_HD void givens(double alpha, double beta, double* c, double* s, double eps)
{
    bool const tiny = near_zero(beta, eps);
    bool const steep = fabs(beta) > fabs(alpha);
    double const num = steep ? -alpha : -beta;
    double const den = tiny ? 1.0 : (steep ? beta : alpha);
    double const tao = num / den;
    double const r = 1.0 / sqrt(add(1.0, mul(tao, tao)));
    double const rt = mul(r, tao);
    *c = tiny ? 1.0 : (steep ? rt : r);
    *s = tiny ? 0.0 : (steep ? r : rt);
}

// index of the lowest set bit (x != 0)
#if defined(__CUDA_ARCH__)
OSFM_HD int lowest_bit(unsigned x) { return __ffs(static_cast<int>(x)) - 1; }
#else
OSFM_HD int lowest_bit(unsigned x) { return __builtin_ctz(x); }
#endif

// SVD of a square N x N matrix, A = U diag(s) V^T, singular values sorted descending as the
// reference sorts them.  All matrices row-major.  U is produced only when WANT_U.
template <int N, bool WANT_U, class BStore = LocalMatrix<N>>
struct SquareSvd {
    BStore bm;           // A on entry; the bidiagonal / diagonal form afterwards
    LocalMatrix<N> vm;
    LocalMatrix<WANT_U ? N : 1> um;
    double s[N];
    // STOP_AT_FIXED_POINT: an iteration that changes no bit of B, V, U would be repeated
    // unchanged until the reference's iteration limit (the loop's state is those matrices
    // alone), so the loop may end there.  That is the usual end for the 8-point design matrix:
    // its zero singular value is never deflated, two thirds of all samples would run all 81
    // iterations, and 93 % of those sit on a fixed point after about 30.
    unsigned long long delta;      // OR of (old bits ^ new bits) over the stores of the trip
    OSFM_HD bool changed() const { return delta != 0ull; }
    OSFM_HD void note(double before, double after) { delta |= bits_of(before) ^ bits_of(after); }

    OSFM_HD double& B(int r, int c) { return bm.at(r * N + c); }
    OSFM_HD double& V(int r, int c) { return vm.at(r * N + c); }
    OSFM_HD double& U(int r, int c) { return um.at(r * N + c); }

    OSFM_HD void put(double& slot, double value) {
        note(slot, value);
        slot = value;
    }

    // rotate columns i, k of an N x N matrix (matrix_qr.h:75-88).  All loads first: the compiler
    // cannot tell that the stores of one row do not alias the loads of the next, and would
    // otherwise wait for each row's stores before loading the next row.
    template <class M>
    OSFM_HD void rot_columns(M& m, int i, int k, double c, double s_) {
        double t1[N], t2[N];
        for (int j = 0; j < N; ++j) { t1[j] = m.at(j * N + i); t2[j] = m.at(j * N + k); }
        for (int j = 0; j < N; ++j) {
            double const a = sub(mul(c, t1[j]), mul(s_, t2[j]));
            double const b2 = add(mul(s_, t1[j]), mul(c, t2[j]));
            note(t1[j], a);
            note(t2[j], b2);
            m.at(j * N + i) = a;
            m.at(j * N + k) = b2;
        }
    }
    // rotate rows i, k (matrix_qr.h:90-103)
    template <class M>
    OSFM_HD void rot_rows(M& m, int i, int k, double c, double s_) {
        double t1[N], t2[N];
        for (int j = 0; j < N; ++j) { t1[j] = m.at(i * N + j); t2[j] = m.at(k * N + j); }
        for (int j = 0; j < N; ++j) {
            double const a = sub(mul(c, t1[j]), mul(s_, t2[j]));
            double const b2 = add(mul(s_, t1[j]), mul(c, t2[j]));
            note(t1[j], a);
            note(t2[j], b2);
            m.at(i * N + j) = a;
            m.at(k * N + j) = b2;
        }
    }

    OSFM_HD void bidiagonalize(double eps) {
        for (int i = 0; i < N * N; ++i) vm.at(i) = 0.0;
        for (int i = 0; i < N; ++i) V(i, i) = 1.0;
        if (WANT_U) {
            for (int i = 0; i < N * N; ++i) um.at(i) = 0.0;
            for (int i = 0; i < N; ++i) U(i, i) = 1.0;
        }
        double h[N * N], hv[N], line[N], in[N];
        for (int k = 0; k < N - 1; ++k) {
            // ---- from the left: zero column k below the diagonal
            int const len = N - k;
            for (int i = 0; i < len; ++i) in[i] = B(k + i, k);
            double beta;
            householder_vector(in, len, hv, &beta, eps, 1.0);
            for (int i = 0; i < len; ++i)
                for (int j = 0; j < len; ++j)
                    h[i * len + j] = sub(i == j ? 1.0 : 0.0, mul(mul(beta, hv[i]), hv[j]));
            for (int j = 0; j < len; ++j) {
                for (int i = 0; i < len; ++i) line[i] = B(k + i, k + j);
                for (int i = 0; i < len; ++i) {
                    double cur = 0.0;
                    for (int kk = 0; kk < len; ++kk) cur = add(cur, mul(h[i * len + kk], line[kk]));
                    B(k + i, k + j) = cur;
                }
            }
            for (int i = k + 1; i < N; ++i) B(i, k) = 0.0;
            if (WANT_U) {
                for (int i = 0; i < N; ++i) {
                    for (int kk = 0; kk < len; ++kk) line[kk] = U(i, k + kk);
                    for (int j = 0; j < len; ++j) {
                        double cur = 0.0;
                        for (int kk = 0; kk < len; ++kk) cur = add(cur, mul(line[kk], h[kk * len + j]));
                        U(i, k + j) = cur;
                    }
                }
            }
            // ---- from the right: zero row k beyond the superdiagonal
            if (k <= N - 3) {
                double norm = 0.0;
                for (int i = k + 1; i < N; ++i) norm = add(norm, B(k, i));
                if (near_zero(norm, eps)) norm = 1.0;
                int const ilen = N - (k + 1);
                for (int i = 0; i < ilen; ++i) in[i] = B(k, k + 1 + i);
                householder_vector(in, ilen, hv, &beta, eps, norm);
                for (int i = 0; i < ilen; ++i)
                    for (int j = 0; j < ilen; ++j)
                        h[i * ilen + j] = sub(i == j ? 1.0 : 0.0, mul(mul(beta, hv[i]), hv[j]));
                for (int i = 0; i < N - k; ++i) {
                    for (int kk = 0; kk < ilen; ++kk) line[kk] = B(k + i, k + 1 + kk);
                    for (int j = 0; j < ilen; ++j) {
                        double cur = 0.0;
                        for (int kk = 0; kk < ilen; ++kk) cur = add(cur, mul(line[kk], h[kk * ilen + j]));
                        B(k + i, k + 1 + j) = cur;
                    }
                }
                for (int i = k + 2; i < N; ++i) B(k, i) = 0.0;
                for (int i = 0; i < N; ++i) {
                    for (int kk = 0; kk < ilen; ++kk) line[kk] = V(i, k + 1 + kk);
                    for (int j = 0; j < ilen; ++j) {
                        double cur = 0.0;
                        for (int kk = 0; kk < ilen; ++kk) cur = add(cur, mul(line[kk], h[kk * ilen + j]));
                        V(i, k + 1 + j) = cur;
                    }
                }
            }
        }
    }

    // One implicit-shift QR sweep over rows/columns p .. N-q-1 (matrix_svd.h:440-509), cut into
    // its start (the shift) and its rotation steps so that a caller can interleave the steps
    // of many matrices (ransac_kernels.cuh).  State between the pieces: sweep_k, sweep_end,
    // sweep_alpha, sweep_beta.
    int sweep_k, sweep_end;
    double sweep_alpha, sweep_beta;

    OSFM_HD bool sweep_pending() const { return sweep_k < sweep_end; }

    OSFM_HD void sweep_begin(int p, int q) {
        int const len = N - q - p;
        sweep_k = sweep_end = 0;
        if (len < 2) return;      // nothing to rotate
        // the trailing 2 x 2 block of B22 * B22^T
        double c4[4];
        for (int a = 0; a < 2; ++a)
            for (int d = 0; d < 2; ++d) {
                int const ra = p + len - 2 + a, rd = p + len - 2 + d;
                double cur = 0.0;
                for (int kk = 0; kk < N; ++kk) {        // kk < len terms; fixed trip count, no branch
                    int const col = p + (kk < len ? kk : 0);
                    double const term = mul(B(ra, col), B(rd, col));
                    cur = kk < len ? add(cur, term) : cur;
                }
                c4[a * 2 + d] = cur;
            }
        double const tr = add(c4[0], c4[3]);
        double x = add(sub(mul(tr, tr) / 4.0, mul(c4[0], c4[3])), mul(c4[1], c4[2]));
        x = x > 0.0 ? sqrt(x) : 0.0;
        double const eig_1 = sub(tr / 2.0, x), eig_2 = add(tr / 2.0, x);
        double const diff1 = fabs(sub(c4[3], eig_1)), diff2 = fabs(sub(c4[3], eig_2));
        double const mu = diff1 < diff2 ? eig_1 : eig_2;
        sweep_alpha = sub(mul(B(p, p), B(p, p)), mu);
        sweep_beta = mul(B(p, p), B(p, p + 1));
        sweep_k = p;
        sweep_end = N - q - 1;
    }

    OSFM_HD void sweep_rotate(double eps) {
        int const k = sweep_k;
        double c, s_;
        givens(sweep_alpha, sweep_beta, &c, &s_, eps);
        rot_columns(bm, k, k + 1, c, s_);
        rot_columns(vm, k, k + 1, c, s_);
        sweep_alpha = B(k, k);
        sweep_beta = B(k + 1, k);
        givens(sweep_alpha, sweep_beta, &c, &s_, eps);
        rot_rows(bm, k, k + 1, c, s_);
        if (WANT_U) rot_columns(um, k, k + 1, c, s_);
        if (k < sweep_end - 1) {
            sweep_alpha = B(k, k + 1);
            sweep_beta = B(k, k + 2);
        }
        sweep_k = k + 1;
    }

    // a zero on the diagonal: rotate the rest of its row away (matrix_svd.h:511-535)
    OSFM_HD void clear_super_entry(int row, double eps) {
        for (int i = row + 1; i < N; ++i) {
            if (near_zero(B(row, i), eps)) { put(B(row, i), 0.0); break; }
            double norm = add(mul(B(row, i), B(row, i)), mul(B(i, i), B(i, i)));
            norm = mul(sqrt(norm), B(i, i) < 0.0 ? -1.0 : 1.0);
            double const c = B(i, i) / norm;
            double const s_ = B(row, i) / norm;
            rot_rows(bm, row, i, c, s_);
            if (WANT_U) rot_columns(um, row, i, c, s_);
        }
    }

    // The start of one trip of the reference's iteration loop (matrix_svd.h:565-622): the zero
    // tests, the choice of the active block and either the start of a sweep (its rotation
    // steps are then pending) or the clearing of a row.  Returns true when the loop ends here
    // (converged: q == N).
    OSFM_HD bool trip_begin(double eps) {
        delta = 0ull;
        sweep_k = sweep_end = 0;
        // The tests below are the reference's, evaluated without data-dependent branches (see
        // givens): the two zeroing passes, which also record which entries are "zero"
        // (|x| <= eps) as bit masks per row (zr) and per column (zc); the block searches read
        // the masks.  Nothing is modified between the passes and the searches.
        unsigned zr[N], zc[N];
        for (int i = 0; i < N; ++i) zr[i] = zc[i] = 0u;
        for (int r = 0; r < N; ++r)
            for (int c = 0; c < N; ++c) {
                double const x = B(r, c);
                bool const zero = near_zero(x, eps);          // a zeroed entry stays "zero"
                put(B(r, c), zero ? 0.0 : x);
                zr[r] |= (zero ? 1u : 0u) << c;
                zc[c] |= (zero ? 1u : 0u) << r;
            }
        for (int i = 0; i < N - 1; ++i) {
            double const x = B(i, i + 1);
            bool const drop = fabs(x) <= mul(eps, fabs(add(B(i, i), B(i + 1, i + 1))));
            put(B(i, i + 1), drop ? 0.0 : x);
            zr[i] |= (drop ? 1u : 0u) << (i + 1);
            zc[i + 1] |= (drop ? 1u : 0u) << i;
        }
        unsigned const all = (1u << N) - 1u;
        // q: the largest trailing block that is diagonal (every off-diagonal entry zero) and cut
        // off from the rest (the row and the column next to it are zero along the block); every
        // block size is examined, as in the reference, and the last hit counts
        int q = 0;
        bool diagonal = true;
        for (int k = 0; k < N; ++k) {
            int const o = N - k - 1;                        // block B(o.., o..)
            unsigned const beyond = all & ~((2u << o) - 1u);  // bits o+1 .. N-1
            diagonal = diagonal & (((~zr[o]) & beyond) == 0u) & (((~zc[o]) & beyond) == 0u);
            unsigned const along = all & ~((1u << o) - 1u);   // bits o .. N-1
            int const j = o > 0 ? o - 1 : 0;
            bool const enclosed = (((~zr[j]) & along) == 0u) & (((~zc[j]) & along) == 0u);
            bool const hit = diagonal & ((k == N - 1) | enclosed);
            q = hit ? k + 1 : q;
        }
        // z: the block above it whose superdiagonal has no zero (again the last hit counts)
        unsigned super_zero = 0u;                             // bit r: B(r, r+1) is zero
        unsigned diag_zero = 0u;                              // bit r: B(r, r) is zero
        for (int r = 0; r < N; ++r) {
            diag_zero |= ((zr[r] >> r) & 1u) << r;
            if (r < N - 1) super_zero |= ((zr[r] >> (r + 1)) & 1u) << r;
        }
        int z = 0;
        for (int k = 0; k < N; ++k) {
            bool const valid = k < N - q;
            int const o = valid ? N - q - k - 1 : 0;          // superdiagonal rows o .. o+k-1
            unsigned const rows = ((1u << (o + k)) - 1u) & ~((1u << o) - 1u);
            bool const hit = valid & ((super_zero & rows) == 0u);
            z = hit ? k + 1 : z;
        }
        int const p = N - q - z;
        if (q == N) return true;

        // the first zero on the diagonal in rows p .. N-q-2
        unsigned const candidates = diag_zero & ((1u << (N - q - 1)) - 1u) & ~((1u << p) - 1u);
        bool const diagonal_non_zero = candidates == 0u;
        int const nz = diagonal_non_zero ? N - q - 1 : lowest_bit(candidates);
        if (!diagonal_non_zero) put(B(nz, nz), 0.0);
        if (diagonal_non_zero) sweep_begin(p, q);
        else clear_super_entry(nz, eps);
        return false;
    }

    // One whole trip.  Returns true when the loop ends: converged or, with
    // STOP_AT_FIXED_POINT, nothing changed.
    template <bool STOP_AT_FIXED_POINT>
    OSFM_HD bool gk_iteration(double eps) {
        if (trip_begin(eps)) return true;
        while (sweep_pending()) sweep_rotate(eps);
        return STOP_AT_FIXED_POINT && !changed();
    }

    // After the loop: singular values, sign fix, selection sort (matrix_svd.h:624-641, 747-759).
    OSFM_HD void finish(double eps) {
        for (int i = 0; i < N; ++i) s[i] = B(i, i);
        for (int i = 0; i < N; ++i) {
            if (s[i] < eps) {
                s[i] = -s[i];
                if (WANT_U) for (int j = 0; j < N; ++j) U(j, i) = -U(j, i);
            }
        }
        // largest first; stops at the first all-zero tail
        for (int i = 0; i < N; ++i) {
            double largest = 0.0;
            int pos = -1;
            for (int j = i; j < N; ++j)
                if (s[j] > largest) { largest = s[j]; pos = j; }
            if (pos < 0) break;
            if (pos == i) continue;
            double const t = s[i]; s[i] = s[pos]; s[pos] = t;
            for (int r = 0; r < N; ++r) {
                double const tv = V(r, i); V(r, i) = V(r, pos); V(r, pos) = tv;
                if (WANT_U) { double const tu = U(r, i); U(r, i) = U(r, pos); U(r, pos) = tu; }
            }
        }
    }

    // bm holds A on entry.
    template <bool STOP_AT_FIXED_POINT = false>
    OSFM_HD void run(double eps = kSvdEpsilon) {
        bidiagonalize(eps);
        for (int iteration = 0; iteration < N * N; ++iteration)
            if (gk_iteration<STOP_AT_FIXED_POINT>(eps)) break;
        finish(eps);
    }
};

// The 8 x 9 design matrix of eight correspondences, padded to 9 x 9 (fundamental.cc:84-98,
// matrix_svd.h:729-733).  p1 / p2: x0 y0 x1 y1 ... (view 1 / view 2).
template <class M>
OSFM_HD void design_matrix(const double* p1, const double* p2, M& m)
{
    for (int i = 0; i < 8; ++i) {
        double const x1 = p1[2 * i], y1 = p1[2 * i + 1], x2 = p2[2 * i], y2 = p2[2 * i + 1];
        m.at(9 * i + 0) = mul(x2, x1); m.at(9 * i + 1) = mul(x2, y1); m.at(9 * i + 2) = x2;
        m.at(9 * i + 3) = mul(y2, x1); m.at(9 * i + 4) = mul(y2, y1); m.at(9 * i + 5) = y2;
        m.at(9 * i + 6) = x1;          m.at(9 * i + 7) = y1;          m.at(9 * i + 8) = 1.0;
    }
    for (int j = 0; j < 9; ++j) m.at(72 + j) = 0.0;
}

// enforce_fundamental_constraints (fundamental.cc:113-126): F = U * diag(s0, s1, 0) * V^T, each
// product a left-to-right inner product over three terms (math/matrix.h:459-471).
template <bool STOP_AT_FIXED_POINT>
OSFM_HD void enforce_rank2(double* F)
{
    SquareSvd<3, true> svd;
    for (int i = 0; i < 9; ++i) svd.bm.at(i) = F[i];
    svd.template run<STOP_AT_FIXED_POINT>();
    double S[9] = {svd.s[0], 0.0, 0.0, 0.0, svd.s[1], 0.0, 0.0, 0.0, 0.0};
    double us[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double cur = 0.0;
            for (int k = 0; k < 3; ++k) cur = add(cur, mul(svd.U(i, k), S[k * 3 + j]));
            us[i * 3 + j] = cur;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double cur = 0.0;
            for (int k = 0; k < 3; ++k) cur = add(cur, mul(us[i * 3 + k], svd.V(j, k)));
            F[i * 3 + j] = cur;
        }
}

// The fundamental matrix of eight correspondences: the singular vector of the smallest singular
// value of the design matrix, then rank 2 enforced.  F row-major.
template <bool STOP_AT_FIXED_POINT = false>
OSFM_HD void fundamental_from_eight(const double* p1, const double* p2, double* F)
{
    {
        SquareSvd<9, false> svd;
        design_matrix(p1, p2, svd.bm);
        svd.template run<STOP_AT_FIXED_POINT>();
        for (int r = 0; r < 9; ++r) F[r] = svd.V(r, 8);
    }
    enforce_rank2<STOP_AT_FIXED_POINT>(F);
}

// fundamental.cc:225-247
OSFM_HD double sampson_distance(const double* F, double x1, double y1, double x2, double y2)
{
    double const fx0 = add(add(mul(x1, F[0]), mul(y1, F[1])), F[2]);
    double const fx1 = add(add(mul(x1, F[3]), mul(y1, F[4])), F[5]);
    double const fx2 = add(add(mul(x1, F[6]), mul(y1, F[7])), F[8]);
    double e = 0.0;
    e = add(e, mul(x2, fx0));
    e = add(e, mul(y2, fx1));
    e = add(e, mul(1.0, fx2));
    e = mul(e, e);
    double const ft0 = add(add(mul(x2, F[0]), mul(y2, F[3])), F[6]);
    double const ft1 = add(add(mul(x2, F[1]), mul(y2, F[4])), F[7]);
    double sum = 0.0;
    sum = add(sum, mul(fx0, fx0));
    sum = add(sum, mul(fx1, fx1));
    sum = add(sum, mul(ft0, ft0));
    sum = add(sum, mul(ft1, ft1));
    return e / sum;
}

}  // namespace fmath
}  // namespace osfm
