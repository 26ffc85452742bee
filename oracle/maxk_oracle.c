/*
 * CPU ORACLE (test infrastructure, NOT product code) -- plain C restatement of the
 * MaxK-GNN aggregation hot path of julius-sk/spgemm-gnn, float64 accumulation, OpenMP
 * over rows.  Same contract as oracle/maxk_oracle.py; the two are checked against each
 * other in tests/test_oracle.py.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py may load this library.
 *
 * Parity status: MaxK / CBSR layout pinned by tests/golden (vectors produced by the
 * reference's own Python); SpGEMM forward / SSpMM backward anchored on the reference's own
 * layer code and call sites (tests/golden/make_golden_layers.py, layers_reference.npz);
 * DGL's and the native kernels' own arithmetic cannot be pinned (neither is in
 * /root/reference in runnable form) and partitioning is PARITY UNPINNED -- see the header
 * of maxk_oracle.py and DESIGN.md section 3.
 *
 * Reference anchors (paths relative to /root/reference, SASS offsets as quoted in
 * SURVEY.md section 2.3):
 *   mko_topk_cbsr     utils/models.py:14-20 (selection), K1 maxk_kernel so@0x21110 (layout)
 *   mko_cbsr_scatter  K2 maxk_backward_cuda maxk_cuda_kernels.o@0x4d0
 *   mko_partition     K3 host notes: .warp4 records {row, loc, len, pad}, so@0x24d16
 *   mko_spgemm_fwd    K3 spmm_kernel_opt2_sparse_v3 so@0x24b60, fwd.sass 0b00-0dd0
 *   mko_sspmm_bwd     K4 spmm_kernel_opt2_sparse_backward_v3 so@0x257a0, 0c90-0f00
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define MKO_API __attribute__((visibility("default")))

MKO_API int mko_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* selection order of the build contract: larger first, NaN above +inf, -0 == +0 */
static inline uint32_t order_key(float v) {
    uint32_t b;
    memcpy(&b, &v, 4);
    if (v != v) return 0xFFFFFFFFu;
    if (b == 0x80000000u) return 0x80000000u;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

static inline int64_t col_at(const void* idx, int index_bytes, int64_t pos) {
    return index_bytes == 1 ? (int64_t)((const uint8_t*)idx)[pos] : (int64_t)((const uint16_t*)idx)[pos];
}

/* Exact top-k per row -> CBSR (ascending columns; ties -> lower column). O(D*k) per row
 * by repeated "how many beat me" ranking: rank(c) = #{c': key[c'] > key[c] or (== and c' < c)}. */
MKO_API int mko_topk_cbsr(const float* x, int64_t n, int d, int k, float* sp_data, void* sp_index,
                          int index_bytes) {
    if (k < 1 || k > d) return -1;
    if (index_bytes != 1 && index_bytes != 2) return -2;
    if (index_bytes == 1 && d > 256) return -3;
#pragma omp parallel
    {
        uint32_t* key = (uint32_t*)malloc((size_t)d * sizeof(uint32_t));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            const float* row = x + i * d;
            for (int c = 0; c < d; ++c) key[c] = order_key(row[c]);
            int w = 0;
            for (int c = 0; c < d && w < k; ++c) {
                int rank = 0;
                const uint32_t kc = key[c];
                for (int o = 0; o < d; ++o) rank += (key[o] > kc) || (key[o] == kc && o < c);
                if (rank < k) {
                    sp_data[i * k + w] = row[c];
                    if (index_bytes == 1)
                        ((uint8_t*)sp_index)[i * k + w] = (uint8_t)c;
                    else
                        ((uint16_t*)sp_index)[i * k + w] = (uint16_t)c;
                    ++w;
                }
            }
        }
        free(key);
    }
    return 0;
}

MKO_API int mko_cbsr_scatter(const float* g, const void* sp_index, int index_bytes, float* dense,
                             int64_t n, int k, int d) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float* row = dense + i * d;
        for (int c = 0; c < d; ++c) row[c] = 0.0f;
        for (int t = 0; t < k; ++t) row[col_at(sp_index, index_bytes, i * k + t)] = g[i * k + t];
    }
    return 0;
}

MKO_API int mko_cbsr_gather(const float* dense, const void* sp_index, int index_bytes, float* out,
                            int64_t n, int k, int d) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        for (int t = 0; t < k; ++t)
            out[i * k + t] = dense[i * d + col_at(sp_index, index_bytes, i * k + t)];
    return 0;
}

/* records {row, loc, len, slot}; recs == NULL -> count only. Returns the record count. */
MKO_API int64_t mko_partition(const int32_t* ptr, int64_t n, int max_nz, int32_t* recs,
                              int64_t* num_slots) {
    int64_t p = 0, slot = 0;
    for (int64_t r = 0; r < n; ++r) {
        const int64_t lo = ptr[r], hi = ptr[r + 1];
        const int64_t deg = hi - lo;
        int64_t chunks = (deg + max_nz - 1) / max_nz;
        if (chunks < 1) chunks = 1;
        for (int64_t c = 0; c < chunks; ++c) {
            const int64_t loc = lo + c * max_nz;
            int64_t len = hi - loc;
            if (len > max_nz) len = max_nz;
            if (len < 0) len = 0;
            if (recs) {
                recs[4 * p + 0] = (int32_t)r;
                recs[4 * p + 1] = (int32_t)loc;
                recs[4 * p + 2] = (int32_t)len;
                recs[4 * p + 3] = chunks == 1 ? -1 : (int32_t)slot;
            }
            if (chunks > 1) ++slot;
            ++p;
        }
    }
    if (num_slots) *num_slots = slot;
    return p;
}

/* Y[i, sel[j,t]] += val[e] * data[j,t] over e=(i<-j) in CSR row i. float64 out [n_rows, d]. */
MKO_API int mko_spgemm_fwd(const int32_t* ptr, const int32_t* idx, const float* val,
                           const float* sp_data, const void* sp_index, int index_bytes, double* out,
                           int64_t n_rows, int k, int d) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n_rows; ++i) {
        double* y = out + i * d;
        for (int c = 0; c < d; ++c) y[c] = 0.0;
        for (int64_t e = ptr[i]; e < ptr[i + 1]; ++e) {
            const int64_t nz = idx[e];
            const double v = (double)val[e];
            for (int t = 0; t < k; ++t)
                y[col_at(sp_index, index_bytes, nz * k + t)] += v * (double)sp_data[nz * k + t];
        }
    }
    return 0;
}

/* dXs[j,t] = sum over e=(r<-j) val[e] * dY[r, sel[j,t]]. float64 out [n_src, k].
 * Done as a pull over the transposed structure so that the float64 sums are formed in a
 * fixed order without atomics. */
MKO_API int mko_sspmm_bwd(const int32_t* ptr, const int32_t* idx, const float* val, const float* dy,
                          const void* sp_index, int index_bytes, double* out, int64_t n_rows,
                          int64_t n_src, int k, int d) {
    const int64_t nnz = ptr[n_rows];
    int64_t* tptr = (int64_t*)calloc((size_t)n_src + 1, sizeof(int64_t));
    int32_t* trow = (int32_t*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int32_t));
    float* tval = (float*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(float));
    if (!tptr || !trow || !tval) return -1;
    for (int64_t e = 0; e < nnz; ++e) tptr[idx[e] + 1]++;
    for (int64_t j = 0; j < n_src; ++j) tptr[j + 1] += tptr[j];
    int64_t* fill = (int64_t*)malloc((size_t)(n_src > 0 ? n_src : 1) * sizeof(int64_t));
    memcpy(fill, tptr, (size_t)n_src * sizeof(int64_t));
    for (int64_t r = 0; r < n_rows; ++r)
        for (int64_t e = ptr[r]; e < ptr[r + 1]; ++e) {
            const int64_t q = fill[idx[e]]++;
            trow[q] = (int32_t)r;
            tval[q] = val[e];
        }
    free(fill);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t j = 0; j < n_src; ++j) {
        double* g = out + j * k;
        for (int t = 0; t < k; ++t) g[t] = 0.0;
        for (int64_t q = tptr[j]; q < tptr[j + 1]; ++q) {
            const float* row = dy + (int64_t)trow[q] * d;
            const double v = (double)tval[q];
            for (int t = 0; t < k; ++t)
                g[t] += v * (double)row[col_at(sp_index, index_bytes, j * k + t)];
        }
    }
    free(tptr);
    free(trow);
    free(tval);
    return 0;
}
