/*
 * oracle_ref.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Dependency-free C/OpenMP restatement of the reference's CPU hot path in the
 * *canonical arithmetic order* (DESIGN.md "Canonical arithmetic"), so that the CUDA path
 * can be compared bit-for-bit, and timed as the CPU baseline on the box's host cores.
 *
 * Follows (paths relative to /root/reference):
 *   orc_renorm_l2      faiss fvec_renorm_L2 at src/kiss-icp/cpp/kiss_icp/core/VoxelHashMap.cpp:474,480
 *   orc_match_top2     faiss IndexFlatIP::search(k=1) at VoxelHashMap.cpp:486-495
 *                      (+ runner-up value for the ratio test / ambiguity report)
 *   orc_kabsch / orc_ransac
 *                      Open3D registration_ransac_based_on_correspondence at
 *                      src/vfm-reg/src/registration_node.py:319-327 (Umeyama w/o scale,
 *                      3 samples with replacement, inlier test d^2 < tau^2, best by
 *                      (count, residual sum, id)); third-party arithmetic, parity unpinned.
 *
 * Must be compiled with -ffp-contract=off (every fused multiply-add below is explicit).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#define ORC_AVX2 1
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* The timed CPU baseline must use the cores the process may run on even when a launcher (torchrun) exported
 * OMP_NUM_THREADS=1 for its workers. */
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

ORC_API int orc_simd(void) {
#ifdef ORC_AVX2
    return 8;
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------
 * Row L2 renormalisation, canonical order: 32 partial sums (partial l takes the float4
 * groups 4l..4l+3, 128+4l.., ... element by element with fused multiply-add = one warp lane
 * issuing 128-bit loads), then an xor-butterfly 16,8,4,2,1.
 * inv = 1/sqrt(s) in float32, rows with s == 0 are left untouched (faiss: `if (nr > 0)`).
 * ---------------------------------------------------------------------------------- */
static float orc_row_sumsq(const float* x, int d) {
    float part[32], tmp[32];
    for (int l = 0; l < 32; ++l) {
        float acc = 0.0f;
        for (int base = 4 * l; base < d; base += 128)
            for (int c = 0; c < 4 && base + c < d; ++c) acc = fmaf(x[base + c], x[base + c], acc);
        part[l] = acc;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        for (int l = 0; l < 32; ++l) tmp[l] = part[l] + part[l ^ off];
        memcpy(part, tmp, sizeof(part));
    }
    return part[0];
}

ORC_API void orc_renorm_l2(const float* x, int64_t n, int d, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float* xi = x + i * (int64_t)d;
        float* oi = out + i * (int64_t)d;
        float s = orc_row_sumsq(xi, d);
        if (s > 0.0f) {
            float inv = 1.0f / sqrtf(s);
            for (int k = 0; k < d; ++k) oi[k] = xi[k] * inv;
        } else {
            for (int k = 0; k < d; ++k) oi[k] = xi[k];
        }
    }
}

/* ------------------------------------------------------------------------------------
 * Inner-product top-2 of every row of a (n x d) against all rows of b (m x d).
 * Canonical value: acc = 0; for k ascending: acc = fmaf(a[k], b[k], acc).
 * Scan j ascending with strict '>' => lowest index wins exact ties (faiss semantics).
 * ---------------------------------------------------------------------------------- */
#define PANEL 16
#define ROWBLK 64

ORC_API void orc_match_top2(const float* a, int64_t n, const float* b, int64_t m, int d,
                            int32_t* idx, float* best, float* second) {
    if (m <= 0) {
        for (int64_t i = 0; i < n; ++i) { idx[i] = -1; best[i] = -INFINITY; if (second) second[i] = -INFINITY; }
        return;
    }
    const int64_t npanel = (m + PANEL - 1) / PANEL;
    /* pack b into k-major panels of 16 columns (zero padded) */
    float* bp = (float*)aligned_alloc(64, (size_t)npanel * d * PANEL * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < npanel; ++p) {
        float* dst = bp + p * (int64_t)d * PANEL;
        for (int k = 0; k < d; ++k)
            for (int c = 0; c < PANEL; ++c) {
                int64_t j = p * PANEL + c;
                dst[k * PANEL + c] = (j < m) ? b[j * (int64_t)d + k] : 0.0f;
            }
    }
    const int64_t nblk = (n + ROWBLK - 1) / ROWBLK;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t rb = 0; rb < nblk; ++rb) {
        const int64_t i0 = rb * ROWBLK;
        const int rows = (int)((n - i0 < ROWBLK) ? (n - i0) : ROWBLK);
        float vb[ROWBLK], vs[ROWBLK];
        int32_t ib[ROWBLK];
        for (int r = 0; r < rows; ++r) { vb[r] = -INFINITY; vs[r] = -INFINITY; ib[r] = 0; }
        for (int64_t p = 0; p < npanel; ++p) {
            const float* pan = bp + p * (int64_t)d * PANEL;
            const int cols = (int)((m - p * PANEL < PANEL) ? (m - p * PANEL) : PANEL);
            for (int r0 = 0; r0 < rows; r0 += 4) {
                float acc[4][PANEL];
                const int rr = (rows - r0 < 4) ? (rows - r0) : 4;
#ifdef ORC_AVX2
                const float* a0 = a + (i0 + r0) * (int64_t)d;
                const float* a1 = (rr > 1) ? a0 + d : a0;
                const float* a2 = (rr > 2) ? a0 + 2 * (int64_t)d : a0;
                const float* a3 = (rr > 3) ? a0 + 3 * (int64_t)d : a0;
                __m256 c00 = _mm256_setzero_ps(), c01 = c00, c10 = c00, c11 = c00;
                __m256 c20 = c00, c21 = c00, c30 = c00, c31 = c00;
                for (int k = 0; k < d; ++k) {
                    __m256 b0 = _mm256_load_ps(pan + k * PANEL);
                    __m256 b1 = _mm256_load_ps(pan + k * PANEL + 8);
                    __m256 x0 = _mm256_broadcast_ss(a0 + k);
                    __m256 x1 = _mm256_broadcast_ss(a1 + k);
                    __m256 x2 = _mm256_broadcast_ss(a2 + k);
                    __m256 x3 = _mm256_broadcast_ss(a3 + k);
                    c00 = _mm256_fmadd_ps(x0, b0, c00); c01 = _mm256_fmadd_ps(x0, b1, c01);
                    c10 = _mm256_fmadd_ps(x1, b0, c10); c11 = _mm256_fmadd_ps(x1, b1, c11);
                    c20 = _mm256_fmadd_ps(x2, b0, c20); c21 = _mm256_fmadd_ps(x2, b1, c21);
                    c30 = _mm256_fmadd_ps(x3, b0, c30); c31 = _mm256_fmadd_ps(x3, b1, c31);
                }
                _mm256_storeu_ps(acc[0], c00); _mm256_storeu_ps(acc[0] + 8, c01);
                _mm256_storeu_ps(acc[1], c10); _mm256_storeu_ps(acc[1] + 8, c11);
                _mm256_storeu_ps(acc[2], c20); _mm256_storeu_ps(acc[2] + 8, c21);
                _mm256_storeu_ps(acc[3], c30); _mm256_storeu_ps(acc[3] + 8, c31);
#else
                for (int r = 0; r < rr; ++r) {
                    const float* ar = a + (i0 + r0 + r) * (int64_t)d;
                    for (int c = 0; c < PANEL; ++c) acc[r][c] = 0.0f;
                    for (int k = 0; k < d; ++k)
                        for (int c = 0; c < PANEL; ++c)
                            acc[r][c] = fmaf(ar[k], pan[k * PANEL + c], acc[r][c]);
                }
#endif
                for (int r = 0; r < rr; ++r) {
                    float b1v = vb[r0 + r], b2v = vs[r0 + r];
                    int32_t bi = ib[r0 + r];
                    for (int c = 0; c < cols; ++c) {
                        float v = acc[r][c];
                        if (v > b1v) { b2v = b1v; b1v = v; bi = (int32_t)(p * PANEL + c); }
                        else if (v > b2v) { b2v = v; }
                    }
                    vb[r0 + r] = b1v; vs[r0 + r] = b2v; ib[r0 + r] = bi;
                }
            }
        }
        for (int r = 0; r < rows; ++r) {
            idx[i0 + r] = ib[r];
            best[i0 + r] = vb[r];
            if (second) second[i0 + r] = vs[r];
        }
    }
    free(bp);
}

/* ------------------------------------------------------------------------------------
 * Counter-based sampler (shared definition with csrc/ransac.cu and oracle/ransac.py).
 * ---------------------------------------------------------------------------------- */
static inline uint32_t orc_sample(uint64_t seed, uint64_t ctr, uint32_t k) {
    uint64_t z = seed * 0x9E3779B97F4A7C15ULL + ctr;
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (uint32_t)(((z >> 32) * (uint64_t)k) >> 32);
}

ORC_API void orc_sample_indices(uint64_t seed, int64_t n_hyp, int32_t n_corr, int32_t* out) {
    for (int64_t i = 0; i < n_hyp * 3; ++i)
        out[i] = n_corr > 0 ? (int32_t)orc_sample(seed, (uint64_t)i, (uint32_t)n_corr) : 0;
}

/* ------------------------------------------------------------------------------------
 * Rigid fit from a 3x3 cross-covariance (row-major S = sum (q-qm)(p-pm)^T), canonical
 * float64 order: M = S^T S; cyclic Jacobi eigen-decomposition (6 sweeps, pairs
 * (0,1),(0,2),(1,2)); v1,v2 = two leading eigenvectors; u_i = S v_i / sigma_i with one
 * Gram-Schmidt step; u3 = u1 x u2, v3 = v1 x v2 (this is R = U diag(1,1,det(U)det(V)) V^T
 * -- SURVEY.md A.3 -- written without the sign test); R = sum u_i v_i^T; t = qm - R pm.
 * Returns 0 when lambda_1 <= 1e-300 or lambda_2 <= 1e-12 lambda_1 (rank-deficient).
 * No multiply-add contraction anywhere in this function.
 * ---------------------------------------------------------------------------------- */
#define ORC_SWEEPS 6

static int orc_fit_from_sigma(const double S[9], const double pm[3], const double qm[3], double rt[12]) {
    double a[3][3], v[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            a[i][j] = (S[0 * 3 + i] * S[0 * 3 + j] + S[1 * 3 + i] * S[1 * 3 + j]) + S[2 * 3 + i] * S[2 * 3 + j];
            v[i][j] = (i == j) ? 1.0 : 0.0;
        }
    static const int PQ[3][3] = {{0, 1, 2}, {0, 2, 1}, {1, 2, 0}};
    for (int sweep = 0; sweep < ORC_SWEEPS; ++sweep) {
        for (int e = 0; e < 3; ++e) {
            const int p = PQ[e][0], q = PQ[e][1], r = PQ[e][2];
            const double apq = a[p][q];
            if (apq == 0.0) continue;
            const double app = a[p][p], aqq = a[q][q];
            const double theta = (aqq - app) / (2.0 * apq);
            double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
            if (theta < 0.0) t = -t;
            const double c = 1.0 / sqrt(t * t + 1.0);
            const double s = t * c;
            a[p][p] = app - t * apq;
            a[q][q] = aqq + t * apq;
            a[p][q] = 0.0; a[q][p] = 0.0;
            const double arp = a[r][p], arq = a[r][q];
            const double nrp = c * arp - s * arq;
            const double nrq = s * arp + c * arq;
            a[r][p] = nrp; a[p][r] = nrp;
            a[r][q] = nrq; a[q][r] = nrq;
            for (int i = 0; i < 3; ++i) {
                const double vip = v[i][p], viq = v[i][q];
                v[i][p] = c * vip - s * viq;
                v[i][q] = s * vip + c * viq;
            }
        }
    }
    /* order eigenvalues descending, ties -> lower index first */
    int i0 = 0, i1 = 1, i2 = 2, tmp;
    double l0 = a[0][0], l1 = a[1][1], l2 = a[2][2], lt;
    if (l1 > l0) { lt = l0; l0 = l1; l1 = lt; tmp = i0; i0 = i1; i1 = tmp; }
    if (l2 > l0) { lt = l0; l0 = l2; l2 = lt; tmp = i0; i0 = i2; i2 = tmp; }
    if (l2 > l1) { lt = l1; l1 = l2; l2 = lt; tmp = i1; i1 = i2; i2 = tmp; }
    (void)i2;
    if (!(l0 > 1e-300) || !(l1 > 1e-12 * l0)) {
        for (int i = 0; i < 12; ++i) rt[i] = 0.0;
        rt[0] = rt[4] = rt[8] = 1.0;
        return 0;
    }
    const double s1 = sqrt(l0), s2 = sqrt(l1);
    double v1[3], v2[3], v3[3], u1[3], u2[3], u3[3];
    for (int i = 0; i < 3; ++i) { v1[i] = v[i][i0]; v2[i] = v[i][i1]; }
    for (int i = 0; i < 3; ++i) {
        u1[i] = ((S[i * 3 + 0] * v1[0] + S[i * 3 + 1] * v1[1]) + S[i * 3 + 2] * v1[2]) / s1;
        u2[i] = ((S[i * 3 + 0] * v2[0] + S[i * 3 + 1] * v2[1]) + S[i * 3 + 2] * v2[2]) / s2;
    }
    double n1 = sqrt((u1[0] * u1[0] + u1[1] * u1[1]) + u1[2] * u1[2]);
    for (int i = 0; i < 3; ++i) u1[i] = u1[i] / n1;
    const double dp = (u1[0] * u2[0] + u1[1] * u2[1]) + u1[2] * u2[2];
    for (int i = 0; i < 3; ++i) u2[i] = u2[i] - dp * u1[i];
    double n2 = sqrt((u2[0] * u2[0] + u2[1] * u2[1]) + u2[2] * u2[2]);
    for (int i = 0; i < 3; ++i) u2[i] = u2[i] / n2;
    u3[0] = u1[1] * u2[2] - u1[2] * u2[1];
    u3[1] = u1[2] * u2[0] - u1[0] * u2[2];
    u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
    v3[0] = v1[1] * v2[2] - v1[2] * v2[1];
    v3[1] = v1[2] * v2[0] - v1[0] * v2[2];
    v3[2] = v1[0] * v2[1] - v1[1] * v2[0];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            rt[i * 3 + j] = (u1[i] * v1[j] + u2[i] * v2[j]) + u3[i] * v3[j];
    for (int i = 0; i < 3; ++i)
        rt[9 + i] = qm[i] - ((rt[i * 3 + 0] * pm[0] + rt[i * 3 + 1] * pm[1]) + rt[i * 3 + 2] * pm[2]);
    return 1;
}

/* 3-point fit: p[3][3], q[3][3] -> rt[12] = R row-major (9) then t (3). */
ORC_API int orc_kabsch3(const double* p, const double* q, double* rt) {
    double pm[3], qm[3], S[9];
    for (int i = 0; i < 3; ++i) {
        pm[i] = ((p[0 + i] + p[3 + i]) + p[6 + i]) / 3.0;
        qm[i] = ((q[0 + i] + q[3 + i]) + q[6 + i]) / 3.0;
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            S[i * 3 + j] = ((q[0 + i] - qm[i]) * (p[0 + j] - pm[j]) + (q[3 + i] - qm[i]) * (p[3 + j] - pm[j])) +
                           (q[6 + i] - qm[i]) * (p[6 + j] - pm[j]);
    return orc_fit_from_sigma(S, pm, qm, rt);
}

/* squared residual of correspondence (p, q) under rt, canonical fma order */
static inline double orc_resid2(const double* rt, const double* p, const double* q) {
    const double ex = fma(rt[0], p[0], fma(rt[1], p[1], fma(rt[2], p[2], rt[9])));
    const double ey = fma(rt[3], p[0], fma(rt[4], p[1], fma(rt[5], p[2], rt[10])));
    const double ez = fma(rt[6], p[0], fma(rt[7], p[1], fma(rt[8], p[2], rt[11])));
    const double dx = ex - q[0], dy = ey - q[1], dz = ez - q[2];
    return fma(dx, dx, fma(dy, dy, dz * dz));
}

static inline int64_t orc_quant(double d2, double scale) {
    double z = d2 * scale + 4503599627370496.0; /* 2^52: round-to-nearest-even integer in the mantissa */
    int64_t b;
    memcpy(&b, &z, 8);
    return b - 0x4330000000000000LL;
}

/*
 * Full solve.  pq: K x 6 doubles (p.xyz, q.xyz) = the gathered correspondences.
 * sample_idx: H x 3 (NULL -> drawn with orc_sample(seed)).  Outputs (all optional except T):
 *   T[16] row-major 4x4, counts[H] (-1 for degenerate hypotheses), sumq[H], mask[K],
 *   stats[4] = {best, inlier count of the returned transform, sumq of the winner, K}.
 * refit != 0: least-squares refit over the winner's inliers (float64, k ascending sums).
 */
ORC_API void orc_ransac(const double* pq, int32_t K, const int32_t* sample_idx, int64_t H, uint64_t seed,
                        double thresh, int refit, double* T, int32_t* counts, int64_t* sumq_out,
                        uint8_t* mask, int64_t* stats) {
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
    if (stats) { stats[0] = -1; stats[1] = 0; stats[2] = 0; stats[3] = K; }
    if (mask) memset(mask, 0, (size_t)(K > 0 ? K : 0));
    if (counts) for (int64_t h = 0; h < H; ++h) counts[h] = -1;
    if (sumq_out) for (int64_t h = 0; h < H; ++h) sumq_out[h] = 0;
    if (K < 3 || H <= 0) return;
    const double tau2 = thresh * thresh;
    const double scale = 1099511627776.0 / tau2; /* 2^40 / tau^2 */
    int32_t* cnt = (int32_t*)malloc((size_t)H * sizeof(int32_t));
    int64_t* sq = (int64_t*)malloc((size_t)H * sizeof(int64_t));
    double* rts = (double*)malloc((size_t)H * 12 * sizeof(double));
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t h = 0; h < H; ++h) {
        int32_t s[3];
        for (int j = 0; j < 3; ++j)
            s[j] = sample_idx ? sample_idx[h * 3 + j] : (int32_t)orc_sample(seed, (uint64_t)(h * 3 + j), (uint32_t)K);
        double p[9], q[9];
        for (int j = 0; j < 3; ++j)
            for (int c = 0; c < 3; ++c) { p[j * 3 + c] = pq[(int64_t)s[j] * 6 + c]; q[j * 3 + c] = pq[(int64_t)s[j] * 6 + 3 + c]; }
        double* rt = rts + h * 12;
        if (!orc_kabsch3(p, q, rt)) { cnt[h] = -1; sq[h] = 0; continue; }
        int32_t c = 0;
        int64_t sum = 0;
        for (int32_t k = 0; k < K; ++k) {
            const double d2 = orc_resid2(rt, pq + (int64_t)k * 6, pq + (int64_t)k * 6 + 3);
            if (d2 < tau2) { ++c; sum += orc_quant(d2, scale); }
        }
        cnt[h] = c; sq[h] = sum;
    }
    int64_t best = -1;
    for (int64_t h = 0; h < H; ++h) {
        if (cnt[h] < 0) continue;
        if (best < 0 || cnt[h] > cnt[best] || (cnt[h] == cnt[best] && sq[h] < sq[best])) best = h;
    }
    if (counts) memcpy(counts, cnt, (size_t)H * sizeof(int32_t));
    if (sumq_out) memcpy(sumq_out, sq, (size_t)H * sizeof(int64_t));
    if (best >= 0) {
        double rt[12];
        memcpy(rt, rts + best * 12, sizeof(rt));
        int64_t n_in = 0;
        double sp[3] = {0, 0, 0}, sqv[3] = {0, 0, 0}, sqp[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int32_t k = 0; k < K; ++k) {
            const double* p = pq + (int64_t)k * 6;
            const double* q = p + 3;
            const int in = orc_resid2(rt, p, q) < tau2;
            if (mask) mask[k] = (uint8_t)in;
            if (in) {
                ++n_in;
                for (int i = 0; i < 3; ++i) { sp[i] += p[i]; sqv[i] += q[i]; }
                for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sqp[i * 3 + j] += q[i] * p[j];
            }
        }
        if (refit && n_in >= 3) {
            double pm[3], qm[3], S[9], rt2[12];
            for (int i = 0; i < 3; ++i) { pm[i] = sp[i] / (double)n_in; qm[i] = sqv[i] / (double)n_in; }
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) S[i * 3 + j] = sqp[i * 3 + j] - (double)n_in * (qm[i] * pm[j]);
            if (orc_fit_from_sigma(S, pm, qm, rt2)) memcpy(rt, rt2, sizeof(rt));
        }
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T[i * 4 + j] = rt[i * 3 + j]; T[i * 4 + 3] = rt[9 + i]; }
        if (stats) { stats[0] = best; stats[1] = n_in; stats[2] = sq[best]; }
    }
    free(cnt); free(sq); free(rts);
}

/* gather helper: src (N x 3), tgt (M x 3) doubles + corr (K x 2) -> pq (K x 6) */
ORC_API void orc_gather_pq(const double* src, const double* tgt, const int32_t* corr, int32_t K, double* pq) {
    for (int32_t k = 0; k < K; ++k) {
        for (int c = 0; c < 3; ++c) {
            pq[(int64_t)k * 6 + c] = src[(int64_t)corr[2 * k] * 3 + c];
            pq[(int64_t)k * 6 + 3 + c] = tgt[(int64_t)corr[2 * k + 1] * 3 + c];
        }
    }
}
