// harness.h — shared by the integration drivers: deterministic inputs and a fingerprint of
// group elements, so that the *_cpu and *_b200 builds of one driver can be compared bit for
// bit (the reference draws its inputs from std::random_device, LFF/algebra/fields/bigint.tcc:167-179).
#pragma once
#include <chrono>
#include <cstdint>
#include <sstream>
#include <string>
#include <vector>

namespace harness {

inline double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// s_i = seed-dependent non-trivial sequence, full-width in Fr: s_0 = (seed+2)^5 + 7, s_{i+1} = s_i^2 + c
template <typename Fr>
inline std::vector<Fr> scalars(size_t n, long seed)
{
    std::vector<Fr> v(n);
    Fr x = Fr(seed + 2);
    x = x * x * x * x * x + Fr(7);
    const Fr c = Fr(0x9e3779b9L) * Fr(seed + 11);
    for (size_t i = 0; i < n; i++) {
        x = x * x + c;
        v[i] = x;
    }
    return v;
}

// FNV-1a over the affine coordinates (as_bigint limbs) of every point
struct Fingerprint {
    uint64_t h = 1469598103934665603ull;
    void bytes(const void *p, size_t n)
    {
        const unsigned char *b = (const unsigned char *)p;
        for (size_t i = 0; i < n; i++) {
            h ^= b[i];
            h *= 1099511628211ull;
        }
    }
    template <typename G>
    void point(G p)
    {
        std::ostringstream os;
        p.to_affine_coordinates();
        os << p;  // BINARY_OUTPUT serialisation of the normalised point
        const std::string s = os.str();
        bytes(s.data(), s.size());
    }
    std::string hex() const
    {
        char buf[32];
        snprintf(buf, sizeof buf, "%016llx", (unsigned long long)h);
        return buf;
    }
};

}  // namespace harness
