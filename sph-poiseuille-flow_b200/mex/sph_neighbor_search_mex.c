/*
 * sph_neighbor_search_mex.c -- drop-in MEX gateway over libsph_b200 (B200, sm_100a).
 *
 * Replaces mexFunction of the reference's mex/sph_neighbor_search_mex.c:185-421 with the same
 * name, the same 5 inputs / 7 outputs and the same error identifiers, so SPH_Poiseuille.m:167,366
 * calls it unchanged:
 *
 *   [pair_i, pair_j, dx, dy, r, W, dW] = sph_neighbor_search_mex(pos, n_fluid, n_total, h, DL)
 *
 * This file holds NO arithmetic: it unpacks prhs[], calls the C ABI (include/sph_b200.h) and packs
 * plhs[].  Build (MATLAB host):  mex -R2018a -O -output sph_neighbor_search_mex \
 *     sph_neighbor_search_mex.c -I<repo>/include -L<repo>/sph-poiseuille-flow_b200 -lsph_b200
 */
#include "mex.h"
#include "sph_b200.h"

static sphb_mat view(const mxArray *a)
{
    sphb_mat m;
    m.p = mxIsDouble(a) ? mxGetDoubles(a) : NULL;
    m.m = mxGetM(a);
    m.n = mxGetN(a);
    return m;
}

static void raise_last(void) { mexErrMsgIdAndTxt(sphb_last_error_id(), "%s", sphb_last_error_msg()); }

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    sphb_pairs *handle = NULL;
    size_t n_pairs = 0;
    int k;

    if (nrhs != 5) mexErrMsgIdAndTxt("SPH:Neighbor:nrhs", "Expected 5 inputs.");      /* nb:216-218 */
    if (nlhs != 7) mexErrMsgIdAndTxt("SPH:Neighbor:nlhs", "Expected 7 outputs.");     /* nb:219-221 */
    if (!mxIsDouble(prhs[0]) || mxGetN(prhs[0]) != 2)                                  /* nb:223-225 */
        mexErrMsgIdAndTxt("SPH:Neighbor:pos", "pos must be a double matrix of size [n_total x 2].");

    if (sphb_neighbor_search(view(prhs[0]), (int)mxGetScalar(prhs[1]), (int)mxGetScalar(prhs[2]),
                             mxGetScalar(prhs[3]), mxGetScalar(prhs[4]), &handle, &n_pairs))
        raise_last();
    for (k = 0; k < 7; ++k) plhs[k] = mxCreateDoubleMatrix((mwSize)n_pairs, 1, mxREAL);   /* nb:394-401 */
    if (sphb_neighbor_fetch(handle, mxGetDoubles(plhs[0]), mxGetDoubles(plhs[1]), mxGetDoubles(plhs[2]),
                            mxGetDoubles(plhs[3]), mxGetDoubles(plhs[4]), mxGetDoubles(plhs[5]),
                            mxGetDoubles(plhs[6])))
        raise_last();
}
