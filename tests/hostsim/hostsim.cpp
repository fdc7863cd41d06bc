// Host build of the solver sources in motes_b200/csrc/*.cuh with the one-lane SeqCtx.
// TEST INFRASTRUCTURE ONLY: lets the CPU test-suite exercise the exact control flow of the
// CUDA kernels (compiled here by g++) against scipy.  Never loaded by the motes_b200 package.
#include <vector>
#include <cstring>
#include "../../motes_b200/csrc/trf.cuh"

extern "C" int hostsim_fit_batch(const double* profiles, int nfit, int m, const double* alpha0,
                                 double* params, double* cost, int* nfev, int* njev, int* status) {
    std::vector<double> slab(motes::fit_scratch_doubles(m));
    motes::FitScratch ws;
    ws.y = slab.data();
    ws.unit = ws.y + m;
    ws.a = ws.unit + m;
    ws.w = ws.a + 7 * (size_t)m;
    ws.v = ws.w + 36;
    motes::SeqCtx ctx;
    for (int f = 0; f < nfit; ++f) {
        motes::FitResult r;
        motes::trf_fit_moffat(ctx, m, profiles + (size_t)f * m, alpha0[f], ws, &r);
        std::memcpy(params + 6 * (size_t)f, r.x, sizeof(r.x));
        cost[f] = r.cost;
        nfev[f] = r.nfev;
        njev[f] = r.njev;
        status[f] = r.status;
    }
    return 0;
}
