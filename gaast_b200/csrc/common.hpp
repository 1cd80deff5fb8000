// Shared by the host mirror and the device runtime: error type, last-error
// slot, binomials.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>

#include "gaast_b200.h"

namespace gaast {

struct Error : std::runtime_error {
    gaast_status status;
    Error(gaast_status s, const std::string& m) : std::runtime_error(m), status(s) {}
};

void set_last_error(const std::string& m);
const std::string& last_error();
uint64_t binomial(unsigned n, unsigned k);  // algebra.rs:252-254 (0 when k > n)

}  // namespace gaast
