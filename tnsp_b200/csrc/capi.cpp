// Host side of the C-ABI: error text, launch counter, and the per-chain host RNG that reproduces
// TAT.random (reference PyTAT/PyTAT.hpp:87-126: one std::mt19937_64 drawing through libstdc++'s
// uniform_int_distribution<int>, uniform_real_distribution<double>, normal_distribution<double>).
#include <atomic>
#include <mutex>
#include <random>
#include <string>
#include <vector>

#include "tnsp_b200.h"

namespace tnsp {
std::atomic<int64_t> g_launches{0};
static std::mutex g_err_mutex;
static std::string g_error;
void set_error(const std::string& what) {
    std::lock_guard<std::mutex> lock(g_err_mutex);
    g_error = what;
}
struct Rng {
    std::vector<std::mt19937_64> engines;
};
}  // namespace tnsp

using namespace tnsp;

extern "C" int tnsp_abi_version(void) { return 1; }
extern "C" const char* tnsp_last_error(void) {
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lock(g_err_mutex);
    copy = g_error;
    return copy.c_str();
}
extern "C" int64_t tnsp_launch_count(void) { return g_launches.load(); }

extern "C" void* tnsp_rng_create_host(int n_chains) {
    auto* r = new Rng();
    r->engines.resize(n_chains);
    return r;
}
extern "C" void tnsp_rng_destroy_host(void* rng) { delete static_cast<Rng*>(rng); }
extern "C" void tnsp_rng_seed_host(void* rng, int chain, uint32_t seed) { static_cast<Rng*>(rng)->engines[chain].seed(seed); }
extern "C" void tnsp_rng_uniform_int_host(void* rng, const int32_t* lo, const int32_t* hi, const uint8_t* active, int32_t* out) {
    auto& e = static_cast<Rng*>(rng)->engines;
    for (size_t i = 0; i < e.size(); ++i) {
        if (active && !active[i]) continue;
        out[i] = std::uniform_int_distribution<int>(lo[i], hi[i])(e[i]);
    }
}
extern "C" void tnsp_rng_uniform_real_host(void* rng, double lo, double hi, const uint8_t* active, double* out) {
    auto& e = static_cast<Rng*>(rng)->engines;
    for (size_t i = 0; i < e.size(); ++i) {
        if (active && !active[i]) continue;
        out[i] = std::uniform_real_distribution<double>(lo, hi)(e[i]);
    }
}
extern "C" void tnsp_rng_normal_host(void* rng, int chain, double mean, double stddev, int64_t n, double* out) {
    auto& e = static_cast<Rng*>(rng)->engines[chain];
    // one distribution object for the whole fill, as the reference's randn_ does (PyTAT.hpp:1131-1134)
    std::normal_distribution<double> d(mean, stddev);
    for (int64_t i = 0; i < n; ++i) out[i] = d(e);
}
