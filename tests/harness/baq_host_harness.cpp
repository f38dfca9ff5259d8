// TEST INFRASTRUCTURE — the arithmetic of lofreq_b200/csrc/baq_core.cuh compiled for the host (g++, -ffp-contract=off), so that
// tests/test_baq_core.py can pin it against the compiled reference (oracle/_ref/libkparef.so) on a machine without a GPU.
// Never linked into the product: liblofreq_b200.so only contains the device instance (baq.cu).
#include <cmath>
#include <cstdint>
#include <vector>
#include "../../lofreq_b200/csrc/baq_core.cuh"

namespace {
struct HostMem {
    int w3;
    std::vector<double> f, b, s;
    double &F(int i, int c) { return f[(size_t)i * w3 + c]; }
    double &B(int p, int c) { return b[(size_t)p * w3 + c]; }
    double &S(int i) { return s[i]; }
};
}  // namespace

extern "C" int lfb_kpa_host_batch(long long n, const uint8_t *ref, const long long *ref_off, const uint8_t *query, const long long *qry_off,
                                  const uint8_t *qual, float d, float e, int bw, int *state, uint8_t *q, long long *n_fix_out)
{
    float q2p[256];
    for (int i = 0; i < 256; ++i) q2p[i] = (float)pow(10, -i / 10.);
    std::vector<lfb::KpaFix> fix(1024);
    unsigned n_fix = 0;
    for (long long r = 0; r < n; ++r) {
        const int l_ref = (int)(ref_off[r + 1] - ref_off[r]), l_query = (int)(qry_off[r + 1] - qry_off[r]);
        int b2 = l_ref > l_query ? l_ref : l_query;
        if (b2 > bw) b2 = bw;
        const int dl = l_ref > l_query ? l_ref - l_query : l_query - l_ref;
        if (b2 < dl) b2 = dl;
        HostMem m;
        m.w3 = (2 * b2 + 1) * 3 + 6;
        m.f.assign((size_t)(l_query + 1) * m.w3, 0.0 / 0.0);      // NaN: a cell read before it is written shows up in the output
        m.b.assign((size_t)2 * m.w3, 0.0 / 0.0);
        m.s.assign(l_query + 2, 0.0 / 0.0);
        lfb::kpa_glocal_core(ref + ref_off[r], l_ref, query + qry_off[r], l_query, qual ? qual + qry_off[r] : nullptr, d, e, bw, q2p, m,
                             state + qry_off[r], q + qry_off[r], qry_off[r], fix.data(), (int)fix.size(), &n_fix);
    }
    // the guard-band bases, decided with this libm (the product does the same on the host)
    for (unsigned i = 0; i < n_fix && i < fix.size(); ++i) {
        const int k = (int)(-4.343 * log(fix[i].x) + .499);
        q[fix[i].base] = (uint8_t)(k > 100 ? 99 : k);
    }
    if (n_fix_out) *n_fix_out = n_fix;
    return 0;
}
