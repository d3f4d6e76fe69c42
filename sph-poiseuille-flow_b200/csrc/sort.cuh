// sort.cuh -- device cell sort of the fluid rows (SPH_Poiseuille.m:529-568), see sort.cu.
#pragma once
#include "common.cuh"
#include "ops_csr.cuh"

namespace sphb {

struct SortWork {
    int ncx = 0, key_cap = 0;
    DevBuf keys, cnt, start, fill, perm, scan_tmp;
    void release_all();
};

// every per-particle array of the resident state, permuted together (m:542-552)
struct PermuteArrays {
    static constexpr int kN2 = 5;   // pos, vel, wall_vel, force_prior, force
    static constexpr int kN1 = 5;   // rho, mass, drho_dt, p, Vol
    const double2 *in2[kN2];
    double2 *out2[kN2];
    const double *in1[kN1];
    double *out1[kN1];
    const Mat2 *inB;
    Mat2 *outB;
};

int sort_plan(SortWork &w, int nf, double DL, double h, double y_span_hint);
// perm[k] = old row that moves to new row k, for k < nf
int sort_permutation(SortWork &w, const double2 *pos, int nf, double DL, double h, DevScalars *sc, int *flags,
                     cudaStream_t st, long long *launches);
int sort_apply(const SortWork &w, int nf, int nt, const PermuteArrays &A, cudaStream_t st, long long *launches);

}  // namespace sphb
