/*
 * sph_physics_shell_mex.c -- drop-in MEX gateway over libsph_b200 (B200, sm_100a).
 *
 * Replaces mexFunction of the reference's mex/sph_physics_mex.c:1745-1772 (built with
 * -output sph_physics_shell_mex, SPH_Poiseuille.m:41) with the same string-dispatched modes, argument
 * positions, output order and error identifiers, so SPH_Poiseuille.m:169-171 and
 * :380-431 call it unchanged.  No arithmetic here: unpack prhs[], call the C ABI, pack plhs[].
 *
 * Build (MATLAB host):  mex -R2018a -O -output sph_physics_shell_mex sph_physics_shell_mex.c \
 *     -I<repo>/include -L<repo>/sph-poiseuille-flow_b200 -lsph_b200
 */
#include <string.h>

#include "mex.h"
#include "sph_b200.h"

static sphb_mat V(const mxArray *a)
{
    sphb_mat m;
    m.p = mxIsDouble(a) ? mxGetDoubles(a) : NULL;
    m.m = mxGetM(a);
    m.n = mxGetN(a);
    return m;
}
static double S(const mxArray *a) { return mxGetScalar(a); }
static int I(const mxArray *a) { return (int)mxGetScalar(a); }
static mwSize rows(int n) { return n > 0 ? (mwSize)n : 0; }
static double *D(mxArray *a) { return a ? mxGetDoubles(a) : NULL; }
static void need(int ok, const char *id, const char *msg)
{
    if (!ok) mexErrMsgIdAndTxt(id, "%s", msg);
}
static void check(int rc)
{
    if (rc) mexErrMsgIdAndTxt(sphb_last_error_id(), "%s", sphb_last_error_msg());
}

/*
 * CHECKED(call): the reference validates its arguments before it allocates any output (e.g. ph:125-162 before
 * ph:164-166), so a garbage n_total is an SPH:Physics:*:count error there, not a giant allocation.  The library's
 * checks run first in validate-only mode (no device work, output pointers untouched); only then does ALLOC create
 * the outputs, and the call runs for real.
 */
#define CHECKED(ALLOC, call)            \
    do {                                \
        int rc__;                       \
        sphb_validate_only(1);          \
        rc__ = (call);                  \
        sphb_validate_only(0);          \
        check(rc__);                    \
        ALLOC;                          \
        check(call);                    \
    } while (0)

/*
 * Outputs are created with the reference's sizes (which depend only on n_total, e.g. ph:164-166)
 * and handed to the library, which performs the reference's argument checks in the reference's
 * order before touching them; on an error MATLAB frees them when mexErrMsgIdAndTxt unwinds.
 */
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    char mode[64];
    const mxArray **a = prhs + 1;   /* a[k] = k-th input after the mode string */
    int nt, k;

    for (k = 0; k < (nlhs > 1 ? nlhs : 1) && k < 9; ++k) plhs[k] = NULL;   /* D(NULL) = NULL during the validation call */

    need(nrhs >= 1, "SPH:Physics:nrhs", "At least mode input is required.");                 /* ph:1749 */
    need(mxIsChar(prhs[0]) && mxGetString(prhs[0], mode, sizeof(mode)) == 0, "SPH:Physics:mode",
         "First input must be mode string.");                                                 /* ph:1750 */

    if (strcmp(mode, "density_correction") == 0) {                                            /* ph:95-374 */
        need(nrhs == 14, "SPH:Physics:density:nrhs", "density_correction expects 13 inputs after mode.");
        need(nlhs == 3, "SPH:Physics:density:nlhs", "density_correction expects 3 outputs.");
        nt = I(a[9]);
        CHECKED(
            {
                plhs[0] = mxCreateDoubleMatrix(rows(nt), 1, mxREAL);
                plhs[1] = mxCreateDoubleMatrix(rows(nt), 1, mxREAL);
                plhs[2] = mxCreateDoubleMatrix(rows(nt), 4, mxREAL);
            },
            sphb_density_correction(V(a[0]), V(a[1]), V(a[2]), V(a[3]), V(a[4]), V(a[5]), V(a[6]), V(a[7]), I(a[8]),
                                          nt, S(a[10]), S(a[11]), S(a[12]), D(plhs[0]), D(plhs[1]), D(plhs[2])));
    } else if (strcmp(mode, "viscous_force") == 0) {                                          /* ph:396-550 */
        need(nrhs == 16 || nrhs == 17, "SPH:Physics:viscous:nrhs", "viscous_force expects 15 inputs after mode.");
        need(nlhs == 1, "SPH:Physics:viscous:nlhs", "viscous_force expects 1 output.");
        nt = I(a[12]);
        CHECKED(
            {
                plhs[0] = mxCreateDoubleMatrix(rows(nt), 2, mxREAL);
            },
            sphb_viscous_force(V(a[0]), V(a[1]), V(a[2]), V(a[3]), V(a[4]), V(a[5]), V(a[6]), V(a[7]), V(a[8]),
                                     S(a[9]), S(a[10]), I(a[11]), nt, V(a[13]), V(a[14]), D(plhs[0])));
    } else if (strcmp(mode, "transport_correction") == 0) {                                   /* ph:569-714 */
        need(nrhs == 13 || nrhs == 14, "SPH:Physics:transport:nrhs",
             "transport_correction expects 12 or 13 inputs after mode.");
        need(nlhs == 1, "SPH:Physics:transport:nlhs", "transport_correction expects 1 output.");
        nt = I(a[11]);
        CHECKED(
            {
                plhs[0] = mxCreateDoubleMatrix(rows(nt), 2, mxREAL);
            },
            sphb_transport_correction(V(a[0]), V(a[1]), V(a[2]), V(a[3]), V(a[4]), V(a[5]), V(a[6]), V(a[7]), V(a[8]),
                                            S(a[9]), I(a[10]), nt, nrhs == 14, nrhs == 14 ? S(a[12]) : 0.0, D(plhs[0])));
    } else if (strcmp(mode, "integration_1st") == 0) {                                        /* ph:736-967 */
        need(nrhs == 22, "SPH:Physics:int1:nrhs", "integration_1st expects 21 inputs after mode.");
        need(nlhs == 5, "SPH:Physics:int1:nlhs", "integration_1st expects 5 outputs.");
        nt = I(a[16]);
        CHECKED(
            {
                plhs[0] = mxCreateDoubleMatrix(rows(nt), 1, mxREAL);
                plhs[1] = mxCreateDoubleMatrix(rows(nt), 1, mxREAL);
                plhs[2] = mxCreateDoubleMatrix(rows(nt), 2, mxREAL);
                plhs[3] = mxCreateDoubleMatrix(rows(nt), 2, mxREAL);
                plhs[4] = mxCreateDoubleMatrix(rows(nt), 1, mxREAL);
            },
            sphb_integration_1st(V(a[0]), V(a[1]), V(a[2]), V(a[3]), V(a[4]), V(a[5]), V(a[6]), V(a[7]), V(a[8]),
                                       V(a[9]), V(a[10]), V(a[11]), V(a[12]), V(a[13]), S(a[14]), I(a[15]), nt, S(a[17]),
                                       S(a[18]), S(a[19]), V(a[20]), D(plhs[0]), D(plhs[1]), D(plhs[2]), D(plhs[3]),
                                       D(plhs[4])));
    } else if (strcmp(mode, "integration_2nd") == 0) {                                        /* ph:987-1119 */
        need(nrhs == 15, "SPH:Physics:int2:nrhs", "integration_2nd expects 15 inputs after mode.");
        need(nlhs == 3, "SPH:Physics:int2:nlhs", "integration_2nd expects 3 outputs.");
        nt = I(a[12]);
        CHECKED(
            {
                plhs[0] = mxCreateDoubleMatrix(rows(nt), 2, mxREAL);
                plhs[1] = mxCreateDoubleMatrix(rows(nt), 1, mxREAL);
                plhs[2] = mxCreateDoubleMatrix(rows(nt), 2, mxREAL);
            },
            sphb_integration_2nd(V(a[0]), V(a[1]), V(a[2]), V(a[3]), V(a[4]), V(a[5]), V(a[6]), V(a[7]), V(a[8]),
                                       V(a[9]), S(a[10]), I(a[11]), nt, V(a[13]), D(plhs[0]), D(plhs[1]), D(plhs[2])));
    } else if (strcmp(mode, "integration_verlet") == 0) {                                     /* ph:1316-1469 */
        need(nrhs == 22, "SPH:Physics:verlet:nrhs", "integration_verlet expects 21 inputs after mode.");
        need(nlhs == 6, "SPH:Physics:verlet:nlhs", "integration_verlet expects 6 outputs.");
        nt = I(a[16]);
        CHECKED(
            {
                plhs[0] = mxCreateDoubleMatrix(rows(nt), 1, mxREAL);
                plhs[1] = mxCreateDoubleMatrix(rows(nt), 1, mxREAL);
                plhs[2] = mxCreateDoubleMatrix(rows(nt), 2, mxREAL);
                plhs[3] = mxCreateDoubleMatrix(rows(nt), 2, mxREAL);
                plhs[4] = mxCreateDoubleMatrix(rows(nt), 1, mxREAL);
                plhs[5] = mxCreateDoubleMatrix(rows(nt), 2, mxREAL);
            },
            sphb_integration_verlet(V(a[0]), V(a[1]), V(a[2]), V(a[3]), V(a[4]), V(a[5]), V(a[6]), V(a[7]), V(a[8]),
                                          V(a[9]), V(a[10]), V(a[11]), V(a[12]), V(a[13]), S(a[14]), I(a[15]), nt,
                                          S(a[17]), S(a[18]), S(a[19]), V(a[20]), D(plhs[0]), D(plhs[1]), D(plhs[2]),
                                          D(plhs[3]), D(plhs[4]), D(plhs[5])));
    } else if (strcmp(mode, "advance_shell_step") == 0) {                                     /* ph:1490-1639 */
        need(nrhs == 24, "SPH:Physics:advance:nrhs", "advance_shell_step expects 23 inputs after mode.");
        need(nlhs == 9, "SPH:Physics:advance:nlhs", "advance_shell_step expects 9 outputs.");
        nt = I(a[15]);
        CHECKED(
            {
                plhs[0] = mxCreateDoubleMatrix(rows(nt), 1, mxREAL);
                plhs[1] = mxCreateDoubleMatrix(rows(nt), 1, mxREAL);
                plhs[2] = mxCreateDoubleMatrix(rows(nt), 2, mxREAL);
                plhs[3] = mxCreateDoubleMatrix(rows(nt), 2, mxREAL);
                plhs[4] = mxCreateDoubleMatrix(rows(nt), 1, mxREAL);
                plhs[5] = mxCreateDoubleMatrix(rows(nt), 2, mxREAL);
                plhs[6] = mxCreateDoubleMatrix(rows(nt), 2, mxREAL);
                plhs[7] = mxCreateDoubleMatrix(rows(nt), 1, mxREAL);
                plhs[8] = mxCreateDoubleMatrix(rows(nt), 4, mxREAL);
            },
            sphb_advance_shell_step(V(a[0]), V(a[1]), V(a[2]), V(a[3]), V(a[4]), V(a[5]), V(a[6]), V(a[7]), V(a[8]),
                                          V(a[9]), V(a[10]), V(a[11]), V(a[12]), S(a[13]), I(a[14]), nt, S(a[16]),
                                          S(a[17]), S(a[18]), S(a[19]), S(a[20]), S(a[21]), S(a[22]), D(plhs[0]),
                                          D(plhs[1]), D(plhs[2]), D(plhs[3]), D(plhs[4]), D(plhs[5]), D(plhs[6]),
                                          D(plhs[7]), D(plhs[8])));
    } else if (strcmp(mode, "wall_shear_monitor") == 0) {                                     /* ph:1653-1743 */
        double tb = 0.0, tt = 0.0;
        need(nrhs == 17, "SPH:Physics:wallshear:nrhs", "wall_shear_monitor expects 16 inputs after mode.");
        need(nlhs == 2, "SPH:Physics:wallshear:nlhs", "wall_shear_monitor expects 2 outputs.");
        check(sphb_wall_shear_monitor(V(a[0]), V(a[1]), V(a[2]), V(a[3]), V(a[4]), V(a[5]), V(a[6]), V(a[7]), V(a[8]),
                                      V(a[9]), V(a[10]), I(a[11]), S(a[12]), S(a[13]), S(a[14]), S(a[15]), &tb, &tt));
        plhs[0] = mxCreateDoubleScalar(tb);
        plhs[1] = mxCreateDoubleScalar(tt);
    } else {
        mexErrMsgIdAndTxt("SPH:Physics:mode", "Unsupported mode.");                            /* ph:1770 */
    }
}
