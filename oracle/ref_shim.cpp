// ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin extern "C" door onto the UNMODIFIED reference pricers, compiled
// together with the reference's own sources where they lie under
// /root/reference/src (see oracle/Makefile; nothing is copied into this repo).
// The resulting oracle/_ref/libkwref.so is used (a) to pin oracle/fd1d_oracle.c
// bit-for-bit, (b) to generate tests/golden/*.npz, and (c) as the CPU baseline
// ("kind": "reference") in bench.py -- it runs the reference's own thread pool
// (src/kwThreadPool.cpp:11, hardware_concurrency() workers).
//
// Entry: kw::PricerFactory::create (src/Pricer/kwPricerFactory.h:15-41) then
// kw::Pricer::price (src/Pricer/kwPricer.h:19-21), exactly what
// test/kwPricer_test.cpp:69-76 does.
#include <cstring>
#include <string>
#include <vector>

#include "Pricer/kwPricerFactory.h"
#include "kwThreadPool.h"

static_assert(sizeof(kw::Option) == 56, "kw::Option wire layout");

static void put_err(char* err, size_t errlen, const std::string& s)
{
    if (err && errlen) {
        std::strncpy(err, s.c_str(), errlen - 1);
        err[errlen - 1] = 0;
    }
}

extern "C" {

// mode: "FD1D" | "BS" | "FD1D-BS".  Returns 0 on success (prices[n] filled),
// 1 with the reference's error string otherwise.  n == 0 -> success, nothing written.
int kwref_price(const char* mode, double density, double scale, long long tDim, long long xDim,
                const kw::Option* assets, size_t n, double* prices, char* err, size_t errlen)
{
    kw::Config config;
    config.set("PRICER", std::string(mode));
    config.set("FD1D.DENSITY", density);
    config.set("FD1D.SCALE", scale);
    config.set("FD1D.T_GRID_SIZE", (kw::i64)tDim);
    config.set("FD1D.X_GRID_SIZE", (kw::i64)xDim);

    kw::sPtr<kw::Pricer> pricer;
    if (auto e = kw::PricerFactory::create(config, pricer); !e.empty()) {
        put_err(err, errlen, e);
        return 1;
    }
    std::vector<kw::Option> in(assets, assets + n);
    std::vector<double> out;
    if (auto e = pricer->price(in, out); !e.empty()) {
        put_err(err, errlen, e);
        return 1;
    }
    for (size_t i = 0; i < out.size() && i < n; ++i) prices[i] = out[i];
    return 0;
}

int kwref_pool_size() { return (int)kw::ThreadPool::instance().size(); }

size_t kwref_sizeof_option() { return sizeof(kw::Option); }
}
