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
// them and they feed back into nothing).
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

constexpr double kSvdEpsilon = 1e-12;     // MATH_SVD_DEFAULT_ZERO_THRESHOLD, matrix_svd.h:30

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

// Givens coefficients (matrix_qr.h:50-73)
OSFM_HD void givens(double alpha, double beta, double* c, double* s, double eps)
{
    if (near_zero(beta, eps)) { *c = 1.0; *s = 0.0; return; }
    if (fabs(beta) > fabs(alpha)) {
        double const tao = (-alpha) / beta;
        *s = 1.0 / sqrt(add(1.0, mul(tao, tao)));
        *c = mul(*s, tao);
    } else {
        double const tao = (-beta) / alpha;
        *c = 1.0 / sqrt(add(1.0, mul(tao, tao)));
        *s = mul(*c, tao);
    }
}

// SVD of a square N x N matrix, A = U diag(s) V^T, singular values sorted descending as the
// reference sorts them.  All matrices row-major.  U is produced only when WANT_U.
template <int N, bool WANT_U>
struct SquareSvd {
    double b[N * N];     // A on entry; the bidiagonal / diagonal form afterwards
    double v[N * N];
    double u[WANT_U ? N * N : 1];
    double s[N];

    OSFM_HD double& B(int r, int c) { return b[r * N + c]; }
    OSFM_HD double& V(int r, int c) { return v[r * N + c]; }
    OSFM_HD double& U(int r, int c) { return u[r * N + c]; }

    // rotate columns i, k of an N x N matrix (matrix_qr.h:75-88)
    OSFM_HD static void rot_columns(double* m, int i, int k, double c, double s_) {
        for (int j = 0; j < N; ++j) {
            double const t1 = m[j * N + i], t2 = m[j * N + k];
            m[j * N + i] = sub(mul(c, t1), mul(s_, t2));
            m[j * N + k] = add(mul(s_, t1), mul(c, t2));
        }
    }
    // rotate rows i, k (matrix_qr.h:90-103)
    OSFM_HD static void rot_rows(double* m, int i, int k, double c, double s_) {
        for (int j = 0; j < N; ++j) {
            double const t1 = m[i * N + j], t2 = m[k * N + j];
            m[i * N + j] = sub(mul(c, t1), mul(s_, t2));
            m[k * N + j] = add(mul(s_, t1), mul(c, t2));
        }
    }

    OSFM_HD void bidiagonalize(double eps) {
        for (int i = 0; i < N * N; ++i) v[i] = 0.0;
        for (int i = 0; i < N; ++i) V(i, i) = 1.0;
        if (WANT_U) {
            for (int i = 0; i < N * N; ++i) u[i] = 0.0;
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

    // one implicit-shift QR sweep over rows/columns p .. N-q-1 (matrix_svd.h:440-509)
    OSFM_HD void gk_step(int p, int q, double eps) {
        int const len = N - q - p;
        if (len < 2) return;      // nothing to rotate
        // the trailing 2 x 2 block of B22 * B22^T
        double c4[4];
        for (int a = 0; a < 2; ++a)
            for (int d = 0; d < 2; ++d) {
                int const ra = p + len - 2 + a, rd = p + len - 2 + d;
                double cur = 0.0;
                for (int kk = 0; kk < len; ++kk) cur = add(cur, mul(B(ra, p + kk), B(rd, p + kk)));
                c4[a * 2 + d] = cur;
            }
        double const tr = add(c4[0], c4[3]);
        double x = add(sub(mul(tr, tr) / 4.0, mul(c4[0], c4[3])), mul(c4[1], c4[2]));
        x = x > 0.0 ? sqrt(x) : 0.0;
        double const eig_1 = sub(tr / 2.0, x), eig_2 = add(tr / 2.0, x);
        double const diff1 = fabs(sub(c4[3], eig_1)), diff2 = fabs(sub(c4[3], eig_2));
        double const mu = diff1 < diff2 ? eig_1 : eig_2;

        double alpha = sub(mul(B(p, p), B(p, p)), mu);
        double beta = mul(B(p, p), B(p, p + 1));
        for (int k = p; k < N - q - 1; ++k) {
            double c, s_;
            givens(alpha, beta, &c, &s_, eps);
            rot_columns(b, k, k + 1, c, s_);
            rot_columns(v, k, k + 1, c, s_);
            alpha = B(k, k);
            beta = B(k + 1, k);
            givens(alpha, beta, &c, &s_, eps);
            rot_rows(b, k, k + 1, c, s_);
            if (WANT_U) rot_columns(u, k, k + 1, c, s_);
            if (k < N - q - 2) {
                alpha = B(k, k + 1);
                beta = B(k, k + 2);
            }
        }
    }

    // a zero on the diagonal: rotate the rest of its row away (matrix_svd.h:511-535)
    OSFM_HD void clear_super_entry(int row, double eps) {
        for (int i = row + 1; i < N; ++i) {
            if (near_zero(B(row, i), eps)) { B(row, i) = 0.0; break; }
            double norm = add(mul(B(row, i), B(row, i)), mul(B(i, i), B(i, i)));
            norm = mul(sqrt(norm), B(i, i) < 0.0 ? -1.0 : 1.0);
            double const c = B(i, i) / norm;
            double const s_ = B(row, i) / norm;
            rot_rows(b, row, i, c, s_);
            if (WANT_U) rot_columns(u, row, i, c, s_);
        }
    }

    // b holds A on entry.
    OSFM_HD void run(double eps = kSvdEpsilon) {
        bidiagonalize(eps);
        for (int iteration = 0; iteration < N * N; ++iteration) {
            for (int i = 0; i < N * N; ++i)
                if (near_zero(b[i], eps)) b[i] = 0.0;
            for (int i = 0; i < N - 1; ++i)
                if (fabs(B(i, i + 1)) <= mul(eps, fabs(add(B(i, i), B(i + 1, i + 1))))) B(i, i + 1) = 0.0;

            // q: the largest trailing block that is diagonal and cut off from the rest
            int q = 0;
            for (int k = 0; k < N; ++k) {
                int const o = N - k - 1;       // block B(o.., o..)
                bool diagonal = true;
                for (int y = 0; y <= k && diagonal; ++y)
                    for (int x2 = 0; x2 <= k; ++x2)
                        if (x2 != y && !near_zero(B(o + y, o + x2), eps)) { diagonal = false; break; }
                if (!diagonal) continue;
                if (k < N - 1) {
                    int const j = o - 1;
                    bool enclosed = true;
                    for (int i = o; i < N; ++i)
                        if (!near_zero(B(j, i), eps) || !near_zero(B(i, j), eps)) { enclosed = false; break; }
                    if (enclosed) q = k + 1;
                } else {
                    q = k + 1;
                }
            }
            // z: the block above it whose superdiagonal has no zero
            int z = 0;
            for (int k = 0; k < N - q; ++k) {
                int const o = N - q - k - 1;
                bool nonzero = true;
                for (int i = 0; i < k; ++i)
                    if (near_zero(B(o + i, o + i + 1), eps)) { nonzero = false; break; }
                if (nonzero) z = k + 1;
            }
            int const p = N - q - z;
            if (q == N) break;

            bool diagonal_non_zero = true;
            int nz = p;
            for (; nz < N - q - 1; ++nz)
                if (near_zero(B(nz, nz), eps)) { diagonal_non_zero = false; B(nz, nz) = 0.0; break; }
            if (diagonal_non_zero) gk_step(p, q, eps);
            else clear_super_entry(nz, eps);
        }

        for (int i = 0; i < N; ++i) s[i] = B(i, i);
        for (int i = 0; i < N; ++i) {
            if (s[i] < eps) {
                s[i] = -s[i];
                if (WANT_U) for (int j = 0; j < N; ++j) U(j, i) = -U(j, i);
            }
        }
        // selection sort, largest first; stops at the first all-zero tail (matrix_svd.h:747-759)
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
};

// The fundamental matrix of eight correspondences: the singular vector of the smallest singular
// value of the 8 x 9 design matrix (padded to 9 x 9), then rank 2 enforced.
// p1 / p2: x0 y0 x1 y1 ... (view 1 / view 2).  F row-major.
OSFM_HD void fundamental_from_eight(const double* p1, const double* p2, double* F)
{
    {
        SquareSvd<9, false> svd;
        for (int i = 0; i < 8; ++i) {
            double const x1 = p1[2 * i], y1 = p1[2 * i + 1], x2 = p2[2 * i], y2 = p2[2 * i + 1];
            double* r = svd.b + 9 * i;
            r[0] = mul(x2, x1); r[1] = mul(x2, y1); r[2] = x2;
            r[3] = mul(y2, x1); r[4] = mul(y2, y1); r[5] = y2;
            r[6] = x1;          r[7] = y1;          r[8] = 1.0;
        }
        for (int j = 0; j < 9; ++j) svd.b[72 + j] = 0.0;
        svd.run();
        for (int r = 0; r < 9; ++r) F[r] = svd.V(r, 8);
    }
    SquareSvd<3, true> svd;
    for (int i = 0; i < 9; ++i) svd.b[i] = F[i];
    svd.run();
    // U * diag(s0, s1, 0) * V^T, each product a left-to-right inner product over three terms
    // (math/matrix.h:459-471)
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
