/* CPU ORACLE (test infrastructure, never shipped as the product): plain-C
 * restatement of oracle/align.py, used where the NumPy loops are too slow
 * (full-size parity at 4096 pairs, bench.py's cpu_baseline / reference legs).
 *
 * PARITY UNPINNED: the reference ships no code or golden vectors (SURVEY.md
 * section 0 / 8c); oracle/align.py is the definition and tests/ pin this file
 * against it bit-for-bit.  Reference evidence: README.md:21-22, 44-52.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp (oracle/build.py).
 * -ffp-contract=off keeps every multiply/add individually rounded, matching
 * NumPy; sqrtf and '/' are correctly rounded IEEE operations.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DIR_DIAG 0
#define DIR_UP 1
#define DIR_LEFT 2

/* oracle/align.py:pair_cost — a [Ta,V,Cc], b [Tb,V,Cc], out [Ta,Tb] */
void oracle_pair_cost(const float *a, const float *b, int Ta, int Tb, int V, int Cc, float *out)
{
    for (int i = 0; i < Ta; ++i) {
        const float *ai = a + (size_t)i * V * Cc;
        for (int j = 0; j < Tb; ++j) {
            const float *bj = b + (size_t)j * V * Cc;
            float acc = 0.0f;
            for (int v = 0; v < V; ++v) {
                float dx = ai[v * Cc + 0] - bj[v * Cc + 0];
                float dy = ai[v * Cc + 1] - bj[v * Cc + 1];
                float s = dx * dx;
                float t = dy * dy;
                s = s + t;
                acc = acc + sqrtf(s);
            }
            out[(size_t)i * Tb + j] = acc / (float)V;
        }
    }
}

/* oracle/align.py:dtw_accumulate + dtw_backtrack on a cost matrix.
 * path: [Ta+Tb-1][2] int32 (first *path_len rows valid); returns D[Ta-1][Tb-1]. */
float oracle_dtw(const float *c, int Ta, int Tb, int32_t *path, int32_t *path_len,
                 float *D_out /* may be NULL */, uint8_t *dirs_out /* may be NULL */)
{
    float *D = D_out ? D_out : (float *)malloc(sizeof(float) * (size_t)Ta * Tb);
    uint8_t *dirs = dirs_out ? dirs_out : (uint8_t *)malloc((size_t)Ta * Tb);
    D[0] = c[0];
    dirs[0] = DIR_DIAG;
    for (int j = 1; j < Tb; ++j) {
        D[j] = c[j] + D[j - 1];
        dirs[j] = DIR_LEFT;
    }
    for (int i = 1; i < Ta; ++i) {
        const float *Dp = D + (size_t)(i - 1) * Tb;
        float *Di = D + (size_t)i * Tb;
        const float *ci = c + (size_t)i * Tb;
        uint8_t *di = dirs + (size_t)i * Tb;
        Di[0] = ci[0] + Dp[0];
        di[0] = DIR_UP;
        for (int j = 1; j < Tb; ++j) {
            float best = Dp[j - 1];
            uint8_t d = DIR_DIAG;
            if (Dp[j] < best) { best = Dp[j]; d = DIR_UP; }
            if (Di[j - 1] < best) { best = Di[j - 1]; d = DIR_LEFT; }
            Di[j] = ci[j] + best;
            di[j] = d;
        }
    }
    float total = D[(size_t)Ta * Tb - 1];
    /* backtrack into the tail of path, then shift to the front */
    int maxL = Ta + Tb - 1;
    int pos = maxL;
    int i = Ta - 1, j = Tb - 1;
    for (;;) {
        --pos;
        path[2 * pos + 0] = i;
        path[2 * pos + 1] = j;
        if (i == 0 && j == 0) break;
        uint8_t d = dirs[(size_t)i * Tb + j];
        if (d == DIR_DIAG) { --i; --j; }
        else if (d == DIR_UP) { --i; }
        else { --j; }
    }
    int L = maxL - pos;
    memmove(path, path + 2 * pos, sizeof(int32_t) * 2 * (size_t)L);
    for (int k = 2 * L; k < 2 * maxL; ++k) path[k] = -1;
    *path_len = L;
    if (!D_out) free(D);
    if (!dirs_out) free(dirs);
    return total;
}

/* oracle/align.py:align_phase_ref over a batch: la [N,Ta], lb [N,Tb] phase labels, one more rounded
 * add per cell: c' = c + (la[i] != lb[j] ? penalty : 0).  la == NULL: plain align_ref. */
void oracle_align_phase_batch(const float *a, const float *b, const uint8_t *la, const uint8_t *lb,
                              float penalty, int N, int Ta, int Tb, int V, int Cc, float *cost,
                              int32_t *path, int32_t *path_len, int num_threads)
{
    int maxL = Ta + Tb - 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(num_threads)
    for (int n = 0; n < N; ++n) {
        float *c = (float *)malloc(sizeof(float) * (size_t)Ta * Tb);
        oracle_pair_cost(a + (size_t)n * Ta * V * Cc, b + (size_t)n * Tb * V * Cc, Ta, Tb, V, Cc, c);
        if (la) {
            for (int i = 0; i < Ta; ++i)
                for (int j = 0; j < Tb; ++j) {
                    float pen = la[(size_t)n * Ta + i] != lb[(size_t)n * Tb + j] ? penalty : 0.0f;
                    c[(size_t)i * Tb + j] = c[(size_t)i * Tb + j] + pen;
                }
        }
        cost[n] = oracle_dtw(c, Ta, Tb, path + (size_t)n * maxL * 2, path_len + n, NULL, NULL);
        free(c);
    }
}

/* oracle/align.py:align_ref over a batch. a [N,Ta,V,Cc], b [N,Tb,V,Cc];
 * cost [N], path [N,Ta+Tb-1,2] (-1 padded), path_len [N].  Pairs are
 * independent, so the OpenMP loop does not change any result. */
void oracle_align_batch(const float *a, const float *b, int N, int Ta, int Tb, int V, int Cc,
                        float *cost, int32_t *path, int32_t *path_len, int num_threads)
{
    int maxL = Ta + Tb - 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(num_threads)
    for (int n = 0; n < N; ++n) {
        float *c = (float *)malloc(sizeof(float) * (size_t)Ta * Tb);
        oracle_pair_cost(a + (size_t)n * Ta * V * Cc, b + (size_t)n * Tb * V * Cc, Ta, Tb, V, Cc, c);
        cost[n] = oracle_dtw(c, Ta, Tb, path + (size_t)n * maxL * 2, path_len + n, NULL, NULL);
        free(c);
    }
}
