// Special functions for the Student-t decoder of the MIWAE family (reference src/models/VAE.py:3061-3084, SURVEY.md
// section 8f item 4): digamma for positive arguments, needed by d/d(df) of StudentT.log_prob
//     d/d(df) log p = -0.5 log(1 + y^2/df) + 0.5 (df + 1) y^2 / (df^2 (1 + y^2/df))
//                     - [ 0.5/df + 0.5 psi(df/2) - 0.5 psi((df + 1)/2) ]
// (oracle/pcvae_oracle.py: miwae_loss_closed_form_grads).  Host + device so that the host build can be checked against
// scipy.special.digamma (tests/test_abi.py) before any kernel uses it.  Not yet included by a kernel.
#pragma once

#ifdef __CUDACC__
#define PCVAE_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define PCVAE_HD inline
#endif

namespace pcvae {

// psi(x), x > 0: recurrence psi(x) = psi(x + 1) - 1/x up to x >= 6, then the asymptotic series
//   psi(x) ~ ln x - 1/(2x) - 1/(12 x^2) + 1/(120 x^4) - 1/(252 x^6) + 1/(240 x^8)
template <typename T>
PCVAE_HD T digamma_pos(T x) {
    T acc = T(0);
    while (x < T(6)) {
        acc -= T(1) / x;
        x += T(1);
    }
    const T r = T(1) / x, r2 = r * r;
    const T series = r2 * (T(1.0 / 12.0) - r2 * (T(1.0 / 120.0) - r2 * (T(1.0 / 252.0) - r2 * T(1.0 / 240.0))));
    return acc + log(x) - T(0.5) * r - series;
}

// psi((df + 1)/2) - psi(df/2), the combination the Student-t score needs; positive, ~ 1/df for large df
template <typename T>
PCVAE_HD T digamma_half_step(T df) {
    return digamma_pos(T(0.5) * (df + T(1))) - digamma_pos(T(0.5) * df);
}

}  // namespace pcvae
