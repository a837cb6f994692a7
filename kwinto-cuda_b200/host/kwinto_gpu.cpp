// kwinto_gpu.cpp -- command-line driver of the B200 Fd1d pricer.
//
// `price` mirrors the reference's CLI (src/kwinto.cpp:15-96: same options, same output blocks, the
// GPU pricer factory behind Portfolio::price); `bench` restores the historic benchmark command whose
// output the reference keeps under log/ (log/z800_1024_32768.log:3-110, invocation in Makefile:5):
// "Devices Info", "Portfolio", then per precision a "Benchmark for ..." timing block and an
// "Errors for ..." block.  docopt is not in the image, so the arguments are parsed by hand.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "kw/kwPortfolio.h"
#ifdef KW_WITH_REFERENCE
// Built inside the reference tree (make kwinto-gpu-ref: the reference's own pricer sources are compiled into THIS
// driver binary, never into libkwfd1d.so): the --cpu64 arm of the historic bench is the reference's Fd1d_Pricer on
// its thread pool (src/Pricer/kwPricerFactory.h:15-41, mode "FD1D").
#include "Pricer/kwPricerFactory.h"
#endif

using namespace kw;

static const char g_usage[] = R"(
kwinto-gpu - Options Pricing Analytics on NVIDIA B200

Usage:
    kwinto-gpu price [options] <path>
    kwinto-gpu bench [options] <path>

Arguments:
    <path>          CSV file with options data (plain or .zst)

Options:
    --density <num> Density of the x-grid distribution [default: 0.25]
    --scale <num>   Scale of the x-grid distribution [default: 50.0]
    -e <num>        Reject option prices less than <num> from RRMS error stats [default: 0.5]
    -p <name>       Pricer name: FD1D-GPU | FD1D-BS-GPU (FD1D / FD1D-BS are accepted and run on the GPU) [default: FD1D-BS-GPU]
    -t <num>        Use <num> points for t-grid [default: 512]
    -x <num>        Use <num> points for x-grid [default: 512]
    --precision <p> f64 | f32 (fp64 set-up, fp32 march) [default: f64]
    --device <num>  CUDA device [default: 0]
    -v              Show extra details, be verbose

  bench only:
    -b <num>        Options per batch [default: 32768]
    -n <num>        Number of batches [default: 4]
    --gpu32         Benchmark the fp32 march
    --gpu64         Benchmark the fp64 march (default when no arm is given)
    --cpu64         Benchmark the reference's CPU Fd1d pricer (thread pool, fp64); needs the kwinto-gpu-ref build
    --cpu32         Historic fp32 CPU arm (log/z800_1024_32768.log:42-48): the reference's source has no fp32 pricer any more
    --call          Use calls only
    --put           Use puts only

    -h --help       Print this screen
    --version       Print version
)";

struct Args {
    std::string cmd, path;
    std::map<std::string, std::string> opt{{"--density", "0.25"}, {"--scale", "50.0"}, {"-e", "0.5"},
                                           {"-p", "FD1D-BS-GPU"}, {"-t", "512"},      {"-x", "512"},
                                           {"--precision", "f64"}, {"--device", "0"}, {"-b", "32768"},
                                           {"-n", "4"}};
    std::map<std::string, bool> flag{{"-v", false},     {"--gpu32", false}, {"--gpu64", false}, {"--cpu32", false},
                                     {"--cpu64", false}, {"--call", false},  {"--put", false}};
};

static Error parseArgs(int argc, char** argv, Args& a)
{
    for (int i = 1; i < argc; ++i) {
        const std::string s = argv[i];
        if (s == "-h" || s == "--help") {
            std::cout << g_usage;
            std::exit(0);
        }
        if (s == "--version") {
            std::cout << kw_fd1d_version() << std::endl;
            std::exit(0);
        }
        if (a.flag.count(s)) {
            a.flag[s] = true;
        } else if (a.opt.count(s)) {
            if (i + 1 >= argc) return "kwinto-gpu: option " + s + " needs a value";
            a.opt[s] = argv[++i];
        } else if (!s.empty() && s[0] == '-') {
            return "kwinto-gpu: unknown option " + s;
        } else if (a.cmd.empty()) {
            a.cmd = s;
        } else if (a.path.empty()) {
            a.path = s;
        } else {
            return "kwinto-gpu: unexpected argument " + s;
        }
    }
    if (a.cmd != "price" && a.cmd != "bench") return std::string("kwinto-gpu: expected a command (price | bench)\n") + g_usage;
    if (a.path.empty()) return std::string("kwinto-gpu: missing <path>\n") + g_usage;
    return "";
}

static std::string gpuMode(const std::string& p)
{
    if (p == "FD1D") return "FD1D-GPU";
    if (p == "FD1D-BS") return "FD1D-BS-GPU";
    return p;
}

static Error makeConfig(const Args& a, Config& config, f64& tolerance)
{
    config.set("PRICER", gpuMode(a.opt.at("-p")));
    f64 density, scale;
    if (auto error = parseNumber(a.opt.at("--density"), density); !error.empty())
        return "cmdBench: Fail to parse '--density <num>': " + error;
    config.set("FD1D.DENSITY", density);
    if (auto error = parseNumber(a.opt.at("--scale"), scale); !error.empty())
        return "cmdBench: Fail to parse '--scale <num>': " + error;
    config.set("FD1D.SCALE", scale);
    long long t = 0, x = 0, dev = 0;
    if (auto error = parseNumber(a.opt.at("-t"), t); !error.empty()) return "Fail to parse '-t <num>': " + error;
    if (auto error = parseNumber(a.opt.at("-x"), x); !error.empty()) return "Fail to parse '-x <num>': " + error;
    if (auto error = parseNumber(a.opt.at("--device"), dev); !error.empty()) return "Fail to parse '--device <num>': " + error;
    config.set("FD1D.T_GRID_SIZE", (i64)t);
    config.set("FD1D.X_GRID_SIZE", (i64)x);
    config.set("FD1D.GPU.DEVICE", (i64)dev);
    config.set("FD1D.GPU.PRECISION", a.opt.at("--precision"));
    if (auto error = parseNumber(a.opt.at("-e"), tolerance); !error.empty())
        return "cmdBench: Fail to parse '-e <num>': " + error;
    return "";
}

// src/kwinto.cpp:40-92
static Error cmdPrice(const Args& a)
{
    Config config;
    f64 tolerance;
    if (auto err = makeConfig(a, config, tolerance); !err.empty()) return err;

    Portfolio portfolio;
    if (auto err = portfolio.load(a.path); !err.empty()) return "cmdPrice: " + err;

    std::cout << "Portfolio" << std::endl;
    std::cout << "    Assets : " << portfolio.assets().size() << std::endl;
    std::cout << std::endl;

    std::vector<f64> prices;
    if (auto err = portfolio.price(config, prices); !err.empty()) return "cmdPrice: " + err;
    if (auto err = portfolio.printPricesStats(prices, tolerance); !err.empty()) return "cmdPrice: " + err;
    return "";
}

static void printDevices()
{
    const int n = kw_fd1d_device_count();
    for (int d = 0; d < n; ++d) {
        kw_fd1d_device_props p;
        if (kw_fd1d_get_device_props(d, &p) != KW_FD1D_OK) continue;
        std::printf("Devices Info #%d\n", d);
        std::printf("    Name:                 %s\n", p.name);
        std::printf("    Integrated:           %d\n\n", p.integrated);
        std::printf("    Total SM:             %d\n", p.sm_count);
        std::printf("    Clock Rate:           %.1f MHz\n", p.clock_khz / 1000.);
        std::printf("    32-bit Regs (per SM): %d\n", p.regs_per_sm);
        std::printf("    Max Blocks (per SM):  %d\n", p.max_blocks_per_sm);
        std::printf("    Max Threads (per SM): %d\n\n", p.max_threads_per_sm);
        std::printf("    Total Memory:         %llu MB\n", (unsigned long long)(p.total_mem_bytes >> 20));
        std::printf("    Memory Clock Rate:    %.0f MHz\n", p.mem_clock_khz / 1000.);
        std::printf("    Memory Bus Width:     %d bits\n", p.mem_bus_width_bits);
        std::printf("    Peak Bandwidth:       %.3f GB/s\n\n", 2. * p.mem_clock_khz * 1e3 * (p.mem_bus_width_bits / 8.) * 1e-9);
    }
}

// the historic `kwinto bench` (log/z800_1024_32768.log): -n batches of -b options, one price() call each
static Error cmdBench(const Args& a)
{
    Config config;
    f64 tolerance;
    if (auto err = makeConfig(a, config, tolerance); !err.empty()) return err;
    long long batch = 0, count = 0;
    if (auto error = parseNumber(a.opt.at("-b"), batch); !error.empty() || batch <= 0) return "cmdBench: Fail to parse '-b <num>'";
    if (auto error = parseNumber(a.opt.at("-n"), count); !error.empty() || count <= 0) return "cmdBench: Fail to parse '-n <num>'";

    if (a.flag.at("-v")) printDevices();

    Portfolio all;
    if (auto err = all.load(a.path); !err.empty()) return "cmdBench: " + err;
    std::vector<Option> assets;
    std::vector<f64> want;
    for (size_t i = 0; i < all.assets().size(); ++i) {
        const auto& o = all.assets()[i];
        if (a.flag.at("--call") && !a.flag.at("--put") && o.w < 0) continue;
        if (a.flag.at("--put") && !a.flag.at("--call") && o.w > 0) continue;
        assets.push_back(o);
        want.push_back(all.prices()[i]);
    }
    if (assets.empty()) return "cmdBench: no assets left after the --call / --put filter";

    std::cout << "Portfolio" << std::endl;
    std::cout << "    Assets     : " << assets.size() << std::endl;
    std::cout << "    Batch count: " << count << std::endl;
    std::cout << "    Batch size : " << batch << std::endl;
    std::cout << std::endl;

    // the arms of the historic bench, in its order: cpu32, cpu64, gpu32, gpu64 (log/z800_1024_32768.log:42-110)
    std::vector<std::string> arms;
    if (a.flag.at("--cpu32")) arms.push_back("cpu32");
    if (a.flag.at("--cpu64")) arms.push_back("cpu64");
    if (a.flag.at("--gpu32")) arms.push_back("f32");
    if (a.flag.at("--gpu64") || arms.empty()) arms.push_back("f64");

    for (const auto& prec : arms) {
        sPtr<Pricer> pricer;
        std::string label;
        if (prec == "cpu32") {
            std::cout << "Benchmark for Fd1d_Pricer<float>::price\n    not available: the reference's source has no "
                         "single-precision CPU pricer any more (src/Math/kwFd1d.h is f64 only)\n" << std::endl;
            continue;
        } else if (prec == "cpu64") {
#ifdef KW_WITH_REFERENCE
            Config cpu = config;
            cpu.set("PRICER", std::string(a.opt.at("-p") == "FD1D-BS" || a.opt.at("-p") == "FD1D-BS-GPU" ? "FD1D-BS" : "FD1D"));
            if (auto err = PricerFactory::create(cpu, pricer); !err.empty()) return "cmdBench: " + err;
            label = "Fd1d_Pricer<double>::price";
#else
            return "cmdBench: --cpu64 needs the kwinto-gpu-ref build (the reference's pricer sources compiled into the driver: "
                   "make -C kwinto-cuda_b200/host ref)";
#endif
        } else {
            config.set("FD1D.GPU.PRECISION", prec);
            if (auto err = GpuPricerFactory::create(config, pricer); !err.empty()) return "cmdBench: " + err;
            label = std::string("Fd1dGpu_Pricer<") + (prec == "f32" ? "float" : "double") + ">::price";
        }

        // batches wrap around the portfolio when it is smaller than count * batch
        std::vector<Option> in((size_t)batch);
        std::vector<f64> ref((size_t)batch), out;
        std::vector<double> ms;
        f64 absDiffSum1 = 0, absDiffSum2 = 0, relDiffSum1 = 0, relDiffSum2 = 0, mae = 0, mre = 0;
        unsigned long long total = 0;
        {   // warm-up call (buffers, clocks)
            for (long long j = 0; j < batch; ++j) in[j] = assets[j % assets.size()];
            if (auto err = pricer->price(in, out); !err.empty()) return "cmdBench: " + err;
        }
        for (long long b = 0; b < count; ++b) {
            for (long long j = 0; j < batch; ++j) {
                const size_t src = (size_t)((b * batch + j) % (long long)assets.size());
                in[j] = assets[src];
                ref[j] = want[src];
            }
            const auto t0 = std::chrono::steady_clock::now();
            if (auto err = pricer->price(in, out); !err.empty()) return "cmdBench: " + err;
            const auto t1 = std::chrono::steady_clock::now();
            ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
            for (long long j = 0; j < batch; ++j) {
                if (ref[j] < tolerance) continue;
                const double absDiff = std::abs(ref[j] - out[j]);
                const double relDiff = absDiff / ref[j];
                mae = std::max(mae, absDiff);
                mre = std::max(mre, relDiff);
                absDiffSum1 += absDiff;
                absDiffSum2 += absDiff * absDiff;
                relDiffSum1 += relDiff;
                relDiffSum2 += relDiff * relDiff;
                ++total;
            }
        }
        double tot = 0, mn = ms[0], mx = ms[0];
        for (double v : ms) {
            tot += v;
            mn = std::min(mn, v);
            mx = std::max(mx, v);
        }
        const double avg = tot / ms.size();
        double var = 0;
        for (double v : ms) var += (v - avg) * (v - avg);
        const double sd = std::sqrt(var / ms.size());
        std::printf("Benchmark for %s\n", label.c_str());
        std::printf("    funCall : %zu times\n", ms.size());
        std::printf("    totTime : %.3f ms\n", tot);
        std::printf("    avgTime : %.3f ms\n", avg);
        std::printf("    stdTime : %.3f ms\n", sd);
        std::printf("    minTime : %.3f ms\n", mn);
        std::printf("    maxTime : %.3f ms\n", mx);
        std::printf("    options/s : %.0f\n\n", batch / (avg * 1e-3));
        const double am = absDiffSum1 / total, rm = relDiffSum1 / total;
        std::printf("Errors for %s\n", label.c_str());
        std::printf("       RMSE : %.3e\n", std::sqrt(absDiffSum2 / total - am * am));
        std::printf("      RRMSE : %.3e\n", std::sqrt(relDiffSum2 / total - rm * rm));
        std::printf("        MAE : %.3e\n", mae);
        std::printf("        MRE : %.3e\n", mre);
        std::printf("      total : %llu options\n\n", total);
    }
    return "";
}

int main(int argc, char** argv)
{
    Args args;
    if (auto err = parseArgs(argc, argv, args); !err.empty()) {
        std::cerr << err << std::endl;
        return 1;
    }
    std::cout << kw_fd1d_version() << '\n' << std::endl;
    if (args.flag.at("-v")) {
        std::cout << "Command-Line Arguments" << std::endl;
        for (const auto& f : args.flag) std::cout << "    " << f.first << ": " << (f.second ? "true" : "false") << std::endl;
        for (const auto& o : args.opt) std::cout << "    " << o.first << ": \"" << o.second << "\"" << std::endl;
        std::cout << "    <path>: \"" << args.path << "\"" << std::endl;
        std::cout << "    " << args.cmd << ": true" << std::endl;
        std::cout << std::endl;
    }
    const Error error = args.cmd == "price" ? cmdPrice(args) : cmdBench(args);
    if (!error.empty()) {
        std::cerr << error << std::endl;
        return 1;
    }
    return 0;
}
