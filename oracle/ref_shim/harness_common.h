// TEST INFRASTRUCTURE ONLY. Helpers shared by the per-file harnesses that expose the
// reference's own *_api host wrappers (compiled for CPU) behind a C ABI for ctypes.
#pragma once
#include <cstdint>
static inline torch::Tensor T(const void *p, std::vector<int64_t> s, at::ScalarType t = at::ScalarType::Float) {
    return torch::Tensor(const_cast<void *>(p), std::move(s), t);
}
#define REF_RNG_CONTROL(tu)                                                                  \
    extern "C" void ref_##tu##_rng_reset(int64_t n_prior_calls) {                             \
        rng = pcg32{9121};                                                                    \
        for (int64_t k = 0; k < n_prior_calls; ++k) rng.advance();                            \
    }
