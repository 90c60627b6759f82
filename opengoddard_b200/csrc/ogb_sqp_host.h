// ogb_sqp_host.h -- host-side tables of the device SQP (shape, row view of the packed Jacobian pattern, bound
// lists); shared by the CUDA entry points (ogb_sqp.cu) and the serial test harness (tests/emu).
#pragma once
#include <cmath>
#include <string>
#include <vector>

#include "ogb_sqp.h"

struct OgsHostTables {
    OgsShape S;                                   // pointers NOT set (the owner points them at its copies)
    std::vector<int> colptr, prow, rowptr, rcol, rpos, blo, bhi;
    std::vector<double> xl, xu;
};

// colptr [n + 1], prow [nnz]: the packed pattern by variable (ogb_jac_pattern_csc), rows in [0, m]; row m = cost
static inline bool ogs_build_tables(int n, int m, int meq, int nnz, const int* colptr, const int* prow, const double* xl,
                                    const double* xu, double acc, int itermax, OgsHostTables& T, std::string* err) {
    if (n <= 0 || m <= 0 || meq < 0 || meq >= n || meq > m || nnz <= 0 || !colptr || !prow || !xl || !xu) {
        if (err) *err = "ogb_sqp: bad sizes (needs 0 <= meq < n, meq <= m, at least one constraint)";
        return false;
    }
    if (!(acc > 0.0) || itermax < 1) { if (err) *err = "ogb_sqp: acc must be positive and maxiter >= 1"; return false; }
    T.colptr.assign(colptr, colptr + n + 1);
    T.prow.assign(prow, prow + nnz);
    if (T.colptr[0] != 0 || T.colptr[n] != nnz) { if (err) *err = "ogb_sqp: inconsistent column pointers"; return false; }
    const int M = m + 1;
    T.rowptr.assign(M + 1, 0);
    for (int e = 0; e < nnz; ++e) {
        if (prow[e] < 0 || prow[e] >= M) { if (err) *err = "ogb_sqp: pattern row out of range"; return false; }
        T.rowptr[prow[e] + 1] += 1;
    }
    for (int r = 0; r < M; ++r) T.rowptr[r + 1] += T.rowptr[r];
    T.rcol.assign(nnz, 0);
    T.rpos.assign(nnz, 0);
    std::vector<int> fill(T.rowptr.begin(), T.rowptr.end() - 1);
    for (int j = 0; j < n; ++j)
        for (int e = colptr[j]; e < colptr[j + 1]; ++e) {
            const int at = fill[prow[e]]++;
            T.rcol[at] = j;
            T.rpos[at] = e;
        }
    T.xl.assign(xl, xl + n);
    T.xu.assign(xu, xu + n);
    T.blo.clear();
    T.bhi.clear();
    for (int j = 0; j < n; ++j) {
        if (std::isnan(T.xl[j])) T.xl[j] = -INFINITY;
        if (std::isnan(T.xu[j])) T.xu[j] = INFINITY;
        if (std::isfinite(T.xl[j])) T.blo.push_back(j);
        if (std::isfinite(T.xu[j])) T.bhi.push_back(j);
    }
    OgsShape& S = T.S;
    S = OgsShape();
    S.n = n; S.m = m; S.meq = meq; S.nnz = nnz;
    S.nlo = (int)T.blo.size(); S.nhi = (int)T.bhi.size();
    S.acc = acc; S.itermax = itermax;
    ogs_layout(S);
    return true;
}
