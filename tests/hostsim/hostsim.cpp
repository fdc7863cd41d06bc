// Host build of the solver sources in motes_b200/csrc/*.cuh with the one-lane SeqCtx.
// TEST INFRASTRUCTURE ONLY: lets the CPU test-suite exercise the exact control flow of the
// CUDA kernels (compiled here by g++) against scipy.  Never loaded by the motes_b200 package.
#include <vector>
#include <cstring>
#include "../../motes_b200/csrc/trfb.cuh"

extern "C" int hostsim_fit_batch(const double* profiles, int nfit, int m, const double* alpha0,
                                 double* params, double* cost, int* nfev, int* njev, int* status) {
    std::vector<double> slab(motes::fit_scratch_doubles(m));
    motes::FitScratch ws = motes::carve_fit_scratch(slab.data(), m);
    motes::SeqCtx ctx;
    for (int f = 0; f < nfit; ++f) {
        motes::FitResult r;
        motes::trf_fit_moffat(ctx, m, profiles + (size_t)f * m, 1.0, alpha0[f], ws, &r);
        std::memcpy(params + 6 * (size_t)f, r.x, sizeof(r.x));
        cost[f] = r.cost;
        nfev[f] = r.nfev;
        njev[f] = r.njev;
        status[f] = r.status;
    }
    return 0;
}

// The same fits through the phases the persistent GPU kernel runs (trfb.cuh): start -> evaluate -> propose.
extern "C" int hostsim_fit_batch_phased(const double* profiles, int nfit, int m, const double* alpha0,
                                        double* params, double* cost, int* nfev, int* njev, int* status) {
    std::vector<double> slab(motes::fit_sweep_doubles(m));
    motes::SeqCtx ctx;
    for (int f = 0; f < nfit; ++f) {
        motes::FitResult r;
        motes::trf_fit_moffat_phased(ctx, m, profiles + (size_t)f * m, 1.0, alpha0[f], slab.data(), &r);
        std::memcpy(params + 6 * (size_t)f, r.x, sizeof(r.x));
        cost[f] = r.cost;
        nfev[f] = r.nfev;
        njev[f] = r.njev;
        status[f] = r.status == motes::kStatusRunning ? 0 : r.status;
    }
    return 0;
}

#include "../../motes_b200/csrc/bins.cuh"
#include "../../motes_b200/csrc/limits.cuh"
#include "../../motes_b200/csrc/extract.cuh"

extern "C" int hostsim_snr_bins(const double* data, const double* errs, const uint8_t* qual, int nspat, int ndisp,
                                int row_lo, int row_hi, double snr_limit, double min_cols, int32_t* bins_i,
                                double* bins_snr, uint8_t* gap) {
    std::vector<double> S(ndisp), V(ndisp), tmp_snr(ndisp);
    std::vector<int32_t> tmp_i(2 * (size_t)ndisp);
    std::vector<int> counts(nspat);
    for (int c = 0; c < ndisp; ++c) {
        motes::ColumnStats st = motes::column_stats(data, errs, nspat, ndisp, c, row_lo, row_hi);
        S[c] = st.sum_data;
        V[c] = st.sum_var;
        gap[c] = st.gap;
    }
    // the per-column "every row is good" flags the GPU path computes in qual_columns_kernel
    std::vector<uint8_t> allgood(ndisp);
    for (int c = 0; c < ndisp; ++c) {
        int good = 1;
        for (int r = 0; r < nspat; ++r) good &= (qual[(size_t)r * ndisp + c] == 1);
        allgood[c] = (uint8_t)good;
    }
    unsigned long long cell = 0;
    motes::SeqScope sc;
    // the staged column window and the run-ahead over all-good columns, as bins_kernel sets them up (a short
    // window, so that refills in both scan directions are exercised)
    const int cap = 37;
    std::vector<double> wS(cap), wV(cap), xfer(10);
    std::vector<uint8_t> wG(cap);
    motes::BinWindow win{wS.data(), wV.data(), wG.data(), xfer.data(), cap};
    return motes::snr_bins_scan(sc, qual, S.data(), V.data(), nspat, ndisp, snr_limit, min_cols, counts.data(),
                                bins_i, bins_snr, tmp_i.data(), tmp_snr.data(), allgood.data(), &cell, &win);
}

extern "C" int hostsim_interp_limits(const double* table, int nb, int ndisp, double* out_lo, double* out_hi) {
    std::vector<double> inner(2 * ((size_t)ndisp + 2));
    std::vector<unsigned long long> keys(256);
    uint32_t hist[256];
    unsigned long long cell;
    int icell;
    motes::SeqScope sc;
    motes::interpolate_limits(sc, table, nb, nb, ndisp, out_lo, out_hi, inner.data(), keys.data(), hist, &cell, &icell);
    return 0;
}

extern "C" double hostsim_median(const double* v, int n) {
    std::vector<unsigned long long> keys(n);
    for (int i = 0; i < n; ++i) keys[i] = motes::key_of(v[i]);
    uint32_t hist[256];
    unsigned long long cell;
    motes::SeqScope sc;
    return motes::median_of_keys(sc, keys.data(), n, hist, &cell);
}

#include "../../motes_b200/csrc/sky.cuh"

// Order-0 sky subtraction of one frame with numpy's RandomState (624-word key + position) as input/output.
extern "C" int hostsim_sky_median(double* data, double* errs, int nspat, int ndisp, double* lim_lo, double* lim_hi,
                                  int spatial_floor, uint32_t* mt, int* pos, double* sky, double* sky_err,
                                  int32_t* n_sky, int32_t* n_good, int32_t* flag) {
    motes::SeqScope sc;
    std::vector<double> sorted((size_t)ndisp * nspat), column(nspat);
    std::vector<uint16_t> rank_of((size_t)ndisp * nspat), good_row((size_t)nspat);
    std::vector<unsigned long long> keys(nspat);
    uint32_t hist[256];
    unsigned long long cell;
    int icell;
    for (int c = 0; c < ndisp; ++c) {
        for (int r = 0; r < nspat; ++r) column[r] = data[(size_t)r * ndisp + c];
        motes::sky_prepare_column(sc, column.data(), 1, nspat, lim_lo + c, lim_hi + c, spatial_floor, n_sky + c,
                                  n_good + c, flag + c, sorted.data() + (size_t)c * nspat,
                                  rank_of.data() + (size_t)c * nspat, good_row.data(), keys.data(), hist, &cell, &icell);
        if (flag[c] == motes::kSkyNoPixels) return -4;
    }
    std::vector<uint32_t> bhist((size_t)motes::kBootstrap * motes::boot_row_words(nspat)), ring(motes::kRing), counts(64);
    std::vector<uint16_t> ranks(nspat);
    double meds[100], sq[100];
    int ctl[4];
    std::vector<uint32_t> state(625);
    for (int i = 0; i < 624; ++i) state[i] = mt[i];
    state[624] = (uint32_t)*pos;
    motes::BootShared sh{ring.data(), counts.data(), ctl, bhist.data(), ranks.data(), meds, sq};
    motes::sky_bootstrap_frame(sc, state.data(), ndisp, n_good, flag, sorted.data(), rank_of.data(), (size_t)nspat, sky,
                               sky_err, sh);
    for (int i = 0; i < 624; ++i) mt[i] = state[i];
    *pos = (int)state[624];
    for (int c = 0; c < ndisp; ++c) motes::sky_apply_column(data, errs, nspat, ndisp, c, flag[c], sky[c], sky_err[c]);
    return 0;
}

extern "C" void hostsim_mt_seed(uint32_t* mt, uint32_t seed) { motes::mt_seed(mt, seed); }

// Polynomial-order sky subtraction of one frame (sky.sky_polynomial) on the host simulation.
extern "C" int hostsim_sky_poly(double* data, double* errs, int nspat, int ndisp, double* lim_lo, double* lim_hi,
                                int spatial_floor, int order, uint32_t* mt, int* pos, double* model, double* model_err,
                                int32_t* flag, int32_t* n_good) {
    motes::SeqScope sc;
    std::vector<double> sorted((size_t)ndisp * nspat), column(nspat);
    std::vector<uint16_t> rank_of((size_t)ndisp * nspat), good_row((size_t)ndisp * nspat);
    std::vector<unsigned long long> keys(nspat);
    std::vector<int32_t> n_sky(ndisp);
    uint32_t hist[256];
    unsigned long long cell;
    int icell;
    for (int c = 0; c < ndisp; ++c) {
        for (int r = 0; r < nspat; ++r) column[r] = data[(size_t)r * ndisp + c];
        motes::sky_prepare_column(sc, column.data(), 1, nspat, lim_lo + c, lim_hi + c, spatial_floor, n_sky.data() + c,
                                  n_good + c, flag + c, sorted.data() + (size_t)c * nspat,
                                  rank_of.data() + (size_t)c * nspat, good_row.data() + (size_t)c * nspat, keys.data(),
                                  hist, &cell, &icell);
        if (flag[c] == motes::kSkyNoPixels) return -4;
    }
    std::vector<uint32_t> ring(motes::kRing), counts(16), state(625);
    std::vector<double> q((size_t)nspat * motes::kMaxPoly), rmat(36), colscale(6), coef(600), loc(nspat), scl(nspat), red(40),
        samples(100), y((size_t)100 * nspat);
    std::vector<uint16_t> rows(nspat);
    int ctl[4];
    for (int i = 0; i < 624; ++i) state[i] = mt[i];
    state[624] = (uint32_t)*pos;
    motes::PolyShared sh{ring.data(), counts.data(), ctl, q.data(), rmat.data(), colscale.data(), coef.data(), loc.data(),
                         scl.data(), red.data(), samples.data(), rows.data()};
    motes::sky_bootstrap_poly_frame(sc, state.data(), ndisp, nspat, ndisp, order, spatial_floor, errs, n_good, flag,
                                    sorted.data(), rank_of.data(), good_row.data(), (size_t)nspat, y.data(), model,
                                    model_err, sh);
    for (int i = 0; i < 624; ++i) mt[i] = state[i];
    *pos = (int)state[624];
    for (int c = 0; c < ndisp; ++c) motes::sky_apply_column_poly(data, errs, nspat, ndisp, c, flag[c], model, model_err);
    return 0;
}

// MT19937 jump-ahead polynomials (csrc/mtjump.hpp, host code of the product): the window reached by the
// polynomial jump of 2^log2_stride words against the window reached by running the recurrence that far.
// Returns the number of differing words (word 0: top bit only), or -1 when the table could not be built.
#include "../../motes_b200/csrc/mtjump.hpp"
extern "C" int hostsim_jump_mismatches(uint32_t seed, int log2_stride) {
    const motes::JumpTable& table = motes::jump_table();
    if (!table.ok || log2_stride < 0 || log2_stride > 24) return -1;
    motes::jump_detail::HostMt gen(seed);
    const size_t stride = (size_t)1 << log2_stride;
    std::vector<uint32_t> w(624 + stride + 624);
    for (int i = 0; i < 624; ++i) w[i] = gen.mt[i];
    for (size_t n = 624; n < w.size(); ++n) w[n] = gen.next_raw();
    uint32_t out[624];
    motes::jump_window_host(table.poly(log2_stride), w.data(), out);
    int bad = ((out[0] ^ w[stride]) & 0x80000000u) ? 1 : 0;
    for (int j = 1; j < 624; ++j) bad += out[j] != w[stride + j];
    return bad;
}
