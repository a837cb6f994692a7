// test_pricer.cpp -- the reference's pricer tests (test/kwPricer_test.cpp:63-108) re-run
// against the GPU pricers through the same kw::Pricer interface.  gtest is not in the image,
// so a few macros stand in for it.  Exit code 0 = all passed.  Needs a B200.
#include <cmath>
#include <cstdio>
#include <utility>
#include <vector>

#include "kw/kwFd1dGpu.h"

using namespace kw;

static int g_fail = 0;
#define ASSERT_EQ(a, b)                                                            \
    do {                                                                           \
        if (!((a) == (b))) {                                                       \
            std::printf("FAIL %s:%d  %s == %s  (%s)\n", __FILE__, __LINE__, #a, #b, std::string(a).c_str()); \
            return void(++g_fail);                                                 \
        }                                                                          \
    } while (0)
#define EXPECT_NEAR(a, b, tol)                                                     \
    do {                                                                           \
        if (!(std::fabs((a) - (b)) <= (tol))) {                                    \
            std::printf("FAIL %s:%d  |%.12g - %.12g| > %g\n", __FILE__, __LINE__, (double)(a), (double)(b), (double)(tol)); \
            ++g_fail;                                                              \
        }                                                                          \
    } while (0)

static std::vector<std::pair<Option, f64>> testData()
{
    return {
        {{1.0, 100., 0.2, 0.06, 0.02, 90., 0, -1}, 10.627},
        {{1.0, 100., 0.2, 0.06, 0.02, 100., 0, -1}, 5.885},
        {{1.0, 100., 0.2, 0.06, 0.02, 110., 0, -1}, 2.987},
        {{1.0, 100., 0.2, 0.06, 0.02, 90., 0, +1}, 4.668},
        {{1.0, 100., 0.2, 0.06, 0.02, 100., 0, +1}, 9.729},
        {{1.0, 100., 0.2, 0.06, 0.02, 110., 0, +1}, 16.633},
        {{1.0, 100., 0.2, 0.06, 0.08, 90., 1, -1}, 13.988121682},
        {{1.0, 100., 0.2, 0.06, 0.08, 100., 1, -1}, 8.409190396},
        {{1.0, 100., 0.2, 0.06, 0.08, 110., 1, -1}, 4.6592955111},
        {{1.0, 100., 0.2, 0.06, 0.08, 90., 1, +1}, 2.9472036256},
        {{1.0, 100., 0.2, 0.06, 0.08, 100., 1, +1}, 6.8422540642},
        {{1.0, 100., 0.2, 0.06, 0.08, 110., 1, +1}, 12.7940107108},
    };
}

static void runMode(const char* mode)
{
    std::vector<Option> assets;
    for (const auto& t : testData()) assets.push_back(t.first);

    Config config;
    config.set("PRICER", mode);

    sPtr<Pricer> pricer;
    ASSERT_EQ(GpuPricerFactory::create(config, pricer), "");

    std::vector<f64> prices;
    ASSERT_EQ(pricer->price(assets, prices), "");
    const auto data = testData();
    for (size_t i = 0; i < data.size(); ++i) EXPECT_NEAR(data[i].second, prices[i], 1.3e-3);

    // n == 0 leaves `prices` untouched (src/Pricer/kwFd1d.cpp:24-26)
    std::vector<f64> keep = {1., 2.};
    ASSERT_EQ(pricer->price({}, keep), "");
    if (keep.size() != 2) ++g_fail;
    std::printf("[ OK ] kwPricerTest.%s\n", mode);
}

int main()
{
    runMode("FD1D-GPU");     // kwPricerTest.Fd1d
    runMode("FD1D-BS-GPU");  // kwPricerTest.Fd1dBlackScholes
    {
        Config c;
        sPtr<Pricer> p;
        if (GpuPricerFactory::create(c, p) != "PricerFactory: Missing PRICER key") ++g_fail;
        c.set("PRICER", "NOPE");
        if (GpuPricerFactory::create(c, p) != "PricerFactory: Unknown PRICER = NOPE") ++g_fail;
    }
    std::printf(g_fail ? "FAILED (%d)\n" : "PASSED\n", g_fail);
    return g_fail ? 1 : 0;
}
