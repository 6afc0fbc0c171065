// host/init_state.h -- bit-exact replay of the reference's initial state on the host.
//
// The reference seeds libc rand() once (src/libfm/libfm.cpp:123-124) and then consumes it, in this order:
//   fm_model::init          v[f][j] = N(0, init_stdev)          K*D draws   (src/fm_core/fm_model.h:92-101)
//   fm.w.init_normal        w[j]    = N(0, init_stdev)          D draws     (libfm.cpp:298 / 307 / 313)
//   fm_learn_vb::init       mu_w'[j]   = 0.1 * N(0,1)           D draws     (src/libfm/src/fm_learn_vb.h:709)
//                           mu_v'[f][j] = 0.1 * N(0,1)          K*D draws   (fm_learn_vb.h:711)
// Each Gaussian is Leva's ratio-of-uniforms method on rand()/(RAND_MAX+1) with a data-dependent number of
// rand() calls (src/util/random.h:150-176), so the state can only be reproduced by replaying the stream on
// the same libc. VB parity (1e-4) is tighter than the reference's own seed-to-seed noise,
// hence the initial state is never generated on the device.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

namespace svbfm_host {

inline double libc_uniform() { return rand() / ((double)RAND_MAX + 1); }

inline double libc_gaussian() {   // Leva 1992, constants as in the reference (random.h:150-164)
    double u, v, x, y, Q;
    do {
        do { u = libc_uniform(); } while (u == 0.0);
        v = 1.7156 * (libc_uniform() - 0.5);
        x = u - 0.449871;
        y = std::fabs(v) + 0.386595;
        Q = x * x + y * (0.19600 * y - 0.25472 * x);
        if (Q < 0.27597) break;
    } while ((Q > 0.27846) || ((v * v) > (-4.0 * u * u * std::log(u))));
    return v / u;
}

inline double libc_gaussian(double mean, double stdev) {   // random.h:166-172
    if (stdev == 0.0 || std::isnan(stdev)) return mean;
    return mean + stdev * libc_gaussian();
}

struct InitialState {
    double w0_mean = 0.0, w0_var = 0.02;              // mu_0' = 0, sigma_0' = .02 (fm_learn_vb.h:695-696); mcmc: w0 = 0
    std::vector<double> w_mean, w_var, v_mean, v_var; // [D], [D], [K*D] row-major [f][j], [K*D]
};

// method: 0 vb, 1 vb_online, 2 mcmc (SVBFM_* of include/svbfm.h). Calls srand(seed) itself.
inline void init_state(long seed, uint32_t D, int K, double init_stdev, int method, InitialState& s) {
    srand((unsigned)seed);
    size_t KD = (size_t)K * D;
    std::vector<double> fm_v(KD), fm_w(D);
    for (size_t i = 0; i < KD; i++) fm_v[i] = libc_gaussian(0.0, init_stdev);
    for (uint32_t j = 0; j < D; j++) fm_w[j] = libc_gaussian(0.0, init_stdev);
    if (method == 2) {   // mcmc works on fm.w / fm.v directly
        s.w0_mean = 0.0; s.w0_var = 0.0;
        s.w_mean = fm_w; s.v_mean = fm_v;
        s.w_var.assign(D, 0.0); s.v_var.assign(KD, 0.0);
        return;
    }
    s.w0_mean = 0.0; s.w0_var = 0.02;
    s.w_mean.resize(D); s.v_mean.resize(KD);
    for (uint32_t j = 0; j < D; j++) s.w_mean[j] = 0.1 * libc_gaussian(0, 1);
    for (size_t i = 0; i < KD; i++) s.v_mean[i] = 0.1 * libc_gaussian(0, 1);
    s.w_var.assign(D, .02);     // fm_learn_vb.h:710
    s.v_var.assign(KD, .02);    // fm_learn_vb.h:712
}

// std::random_shuffle of libstdc++ (what fm_learn_vb_online_simultaneous.h:74 calls): for i in [1,n): swap(a[i], a[rand() % (i+1)])
inline void libc_random_shuffle(uint32_t* a, uint32_t n) {
    for (uint32_t i = 1; i < n; i++) {
        uint32_t j = (uint32_t)(rand() % (i + 1));
        if (i != j) { uint32_t t = a[i]; a[i] = a[j]; a[j] = t; }
    }
}

}  // namespace svbfm_host
