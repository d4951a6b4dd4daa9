/* ransac_hostcheck.cc -- TEST INFRASTRUCTURE ONLY.
 *
 * orthosfm_b200/csrc/ransac_math.cuh (the product's device arithmetic for RANSAC-F) compiled
 * for the host, so that tests/test_oracle.py can compare it double for double with the
 * reference's sfm::fundamental_8_point + enforce_fundamental_constraints and
 * sfm::sampson_distance (bound in ref_driver.cc) on this CPU-only container.  The product
 * never loads this library. */
#include "ransac_math.cuh"

extern "C" {

void
osfm_hostcheck_fundamental (const double* p1, const double* p2, double* F)
{
    osfm::fmath::fundamental_from_eight(p1, p2, F);
}

double
osfm_hostcheck_sampson (const double* F, const double* m)
{
    return osfm::fmath::sampson_distance(F, m[0], m[1], m[2], m[3]);
}

/* singular values and V of a 9 x 9 / U, s, V of a 3 x 3 matrix */
void
osfm_hostcheck_svd9 (const double* a, double* s, double* v)
{
    osfm::fmath::SquareSvd<9, false> svd;
    for (int i = 0; i < 81; ++i) svd.bm.at(i) = a[i];
    svd.run();
    for (int i = 0; i < 9; ++i) s[i] = svd.s[i];
    for (int i = 0; i < 81; ++i) v[i] = svd.vm.at(i);
}

void
osfm_hostcheck_svd3 (const double* a, double* u, double* s, double* v)
{
    osfm::fmath::SquareSvd<3, true> svd;
    for (int i = 0; i < 9; ++i) svd.bm.at(i) = a[i];
    svd.run();
    for (int i = 0; i < 3; ++i) s[i] = svd.s[i];
    for (int i = 0; i < 9; ++i) { u[i] = svd.um.at(i); v[i] = svd.vm.at(i); }
}


/* The same with the iteration stopped at a fixed point (what the device kernels do), and the
 * 9 x 9 problem taken through the device's stages: bidiagonalise, hand over the bidiagonal
 * (17 numbers) and V, iterate in a strided buffer, finish.  iterations receives the number of
 * trips of the 9 x 9 loop. */
void
osfm_hostcheck_fundamental_staged (const double* p1, const double* p2, double* F, int* iterations)
{
    using namespace osfm::fmath;
    double diag[9], super[8], vv[81];
    {
        SquareSvd<9, false> a;
        design_matrix(p1, p2, a.bm);
        a.bidiagonalize(kSvdEpsilon);
        for (int i = 0; i < 9; ++i) diag[i] = a.B(i, i);
        for (int i = 0; i < 8; ++i) super[i] = a.B(i, i + 1);
        for (int i = 0; i < 81; ++i) vv[i] = a.vm.at(i);
    }
    double buffer[81 * 3];
    SquareSvd<9, false, StridedMatrix> b;
    b.bm.p = buffer + 1;
    b.bm.stride = 3;
    for (int i = 0; i < 81; ++i) b.bm.at(i) = 0.0;
    for (int i = 0; i < 9; ++i) b.B(i, i) = diag[i];
    for (int i = 0; i < 8; ++i) b.B(i, i + 1) = super[i];
    for (int i = 0; i < 81; ++i) b.vm.at(i) = vv[i];
    /* the stepping of ransac_gk_kernel: a trip's start, then its rotation steps one by one */
    int it = 0;
    bool done = false;
    b.sweep_k = b.sweep_end = 0;
    while (!done)
    {
        if (!b.sweep_pending())
        {
            ++it;
            done = b.trip_begin(kSvdEpsilon);
            if (!done && !b.sweep_pending())
                done = !b.changed() || it >= 81;
        }
        if (!done && b.sweep_pending())
        {
            b.sweep_rotate(kSvdEpsilon);
            if (!b.sweep_pending())
                done = !b.changed() || it >= 81;
        }
    }
    *iterations = it;
    b.finish(kSvdEpsilon);
    for (int r = 0; r < 9; ++r) F[r] = b.V(r, 8);
    enforce_rank2<true>(F);
}

/* RANSAC as ransac_kernels.cuh runs it (fit per sample, count, first best), on the host.
 * matches: x1 y1 x2 y2 per match; samples: 8 ascending indices per iteration.  Returns the
 * number of inliers; inliers and F as for the reference binding. */
int
osfm_hostcheck_ransac (const double* matches, int n, const int* samples, int iterations,
    double threshold, int* inliers, double* F)
{
    double const thr2 = threshold * threshold;
    int best = 0;
    for (int it = 0; it < iterations; ++it)
    {
        double p1[16], p2[16], f[9];
        for (int k = 0; k < 8; ++k)
        {
            const double* m = matches + 4 * samples[it * 8 + k];
            p1[2 * k] = m[0]; p1[2 * k + 1] = m[1]; p2[2 * k] = m[2]; p2[2 * k + 1] = m[3];
        }
        osfm::fmath::fundamental_from_eight(p1, p2, f);
        int count = 0;
        for (int i = 0; i < n; ++i)
            count += osfm::fmath::sampson_distance(f, matches[4 * i], matches[4 * i + 1],
                matches[4 * i + 2], matches[4 * i + 3]) < thr2;
        if (count > best)
        {
            best = count;
            for (int k = 0; k < 9; ++k) F[k] = f[k];
        }
    }
    int at = 0;
    if (best > 0)
        for (int i = 0; i < n; ++i)
            if (osfm::fmath::sampson_distance(F, matches[4 * i], matches[4 * i + 1],
                matches[4 * i + 2], matches[4 * i + 3]) < thr2)
                inliers[at++] = i;
    return at;
}

}
