#include "common.hpp"

#include <algorithm>

namespace gaast {

static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }
const std::string& last_error() { return g_last_error; }

uint64_t binomial(unsigned n, unsigned k) {
    if (k > n) return 0;
    k = std::min(k, n - k);
    uint64_t r = 1;
    for (unsigned i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return r;
}

}  // namespace gaast

extern "C" const char* gaast_last_error(void) { return gaast::last_error().c_str(); }
