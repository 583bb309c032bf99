/* CPU oracle in plain C -- TEST INFRASTRUCTURE, NOT PRODUCT (see oracle/__init__.py).
 *
 * oracle_knn          restates knn_keops   /root/reference/knn_post_dataloader_utils.py:22-41
 *                     (PARITY UNPINNED at the pykeops boundary; pinned arithmetic below)
 * oracle_knn_inverse  restates create_inverse_python
 *                     /root/reference/cpp_wrappers/cpp_pcf_kernel/test_kernels.py:177-213 and the
 *                     output dtypes of knn_inverse_cuda_forward
 *                     /root/reference/cpp_wrappers/cpp_pcf_kernel/src/knn.cu:104-168
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC   (no FMA contraction: the distance is
 * ((dx*dx)+(dy*dy))+(dz*dz) in fp32, every operation rounded separately).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* out[q*K + i] = index of the i-th smallest (d, idx) reference for query q.  K <= n_ref. */
int oracle_knn(const float *ref, int64_t n_ref, const float *query, int64_t n_query, int K,
               int64_t *out, int threads)
{
    if (K <= 0 || K > n_ref || K > 1024) return 1;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    #pragma omp parallel
    {
        float *bd = (float *)malloc(sizeof(float) * (size_t)K);
        int64_t *bi = (int64_t *)malloc(sizeof(int64_t) * (size_t)K);
        #pragma omp for schedule(dynamic, 64)
        for (int64_t q = 0; q < n_query; ++q) {
            const float qx = query[3 * q], qy = query[3 * q + 1], qz = query[3 * q + 2];
            int filled = 0;
            for (int64_t r = 0; r < n_ref; ++r) {
                const float dx = qx - ref[3 * r], dy = qy - ref[3 * r + 1], dz = qz - ref[3 * r + 2];
                const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
                const float s = xx + yy;
                const float d = s + zz;
                if (filled == K && !(d < bd[K - 1])) continue;   /* strict <: ties keep lower idx */
                int pos = filled < K ? filled : K - 1;
                while (pos > 0 && d < bd[pos - 1]) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
                bd[pos] = d; bi[pos] = r;
                if (filled < K) ++filled;
            }
            memcpy(out + q * K, bi, sizeof(int64_t) * (size_t)K);
        }
        free(bd); free(bi);
    }
    return 0;
}

/* CSR transpose of an [n_out, K] neighbour table over `total` input points.
 * inv_idx[total+1] exclusive prefix of in-degree; segment p lists (n, k) with nei[n,k]==p in
 * ascending n, then k (the order create_inverse_python's nested loops produce).  Entries outside
 * [0,total) (e.g. -1 padding) are skipped, as count_neighbors_kernel does (knn.cu:38). */
int oracle_knn_inverse(const int64_t *nei, int64_t n_out, int K, int64_t total,
                       int32_t *inv_neighbors, uint8_t *inv_k, int32_t *inv_idx)
{
    if (K > 255) return 1;
    int32_t *cur = (int32_t *)calloc((size_t)total + 1, sizeof(int32_t));
    if (!cur) return 2;
    memset(inv_idx, 0, sizeof(int32_t) * ((size_t)total + 1));
    for (int64_t e = 0; e < n_out * K; ++e) {
        int64_t p = nei[e];
        if (p >= 0 && p < total) inv_idx[p + 1]++;
    }
    for (int64_t p = 0; p < total; ++p) inv_idx[p + 1] += inv_idx[p];
    for (int64_t n = 0; n < n_out; ++n)
        for (int k = 0; k < K; ++k) {
            int64_t p = nei[n * K + k];
            if (p < 0 || p >= total) continue;
            int32_t pos = inv_idx[p] + cur[p]++;
            inv_neighbors[pos] = (int32_t)n;
            inv_k[pos] = (uint8_t)k;
        }
    free(cur);
    return 0;
}
