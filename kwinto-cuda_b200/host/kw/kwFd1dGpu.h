// kwFd1dGpu.h -- kw::Pricer subclasses that put the B200 Fd1d path behind the reference's own
// pricer interface (src/Pricer/kwPricer.h:12-22).  Header-only; links against libkwfd1d.so
// (include/kw_fd1d.h).  Drop-in for kw::Fd1d_Pricer (src/Pricer/kwFd1d.h:14-38) and
// kw::Fd1d_BlackScholes_Pricer (src/Pricer/kwF1d1_BlackScholes.h:10-24).
#pragma once
#include <string>
#include <vector>

#include "kw_fd1d.h"
#include "kwCompat.h"

namespace kw {

static_assert(sizeof(Option) == sizeof(kw_option) && sizeof(Option) == 56, "kw::Option wire layout");

class Fd1dGpu_Pricer : public Pricer {
protected:
    kw_fd1d_handle* m_handle = nullptr;
    virtual int run(const kw_option* a, size_t n, double* out) { return kw_fd1d_price(m_handle, a, n, out); }

public:
    Fd1dGpu_Pricer() = default;
    Fd1dGpu_Pricer(const Fd1dGpu_Pricer&) = delete;
    Fd1dGpu_Pricer& operator=(const Fd1dGpu_Pricer&) = delete;
    ~Fd1dGpu_Pricer() { kw_fd1d_destroy(m_handle); }

    // same keys and defaults as Fd1d_Pricer::init (src/Pricer/kwFd1d.cpp:12-16) + FD1D.GPU.*
    Error init(const Config& config) override
    {
        kw_fd1d_destroy(m_handle);
        m_handle = nullptr;
        kw_fd1d_config c;
        kw_fd1d_config_default(&c);
        c.density = config.get("FD1D.DENSITY", 0.25);
        c.scale = config.get("FD1D.SCALE", 50.);
        c.t_grid_size = config.get("FD1D.T_GRID_SIZE", 512);
        c.x_grid_size = config.get("FD1D.X_GRID_SIZE", 512);
        c.device = (int32_t)config.get("FD1D.GPU.DEVICE", 0);
        c.compress = (int32_t)config.get("FD1D.GPU.COMPRESS", 1);
        c.variant = (int32_t)config.get("FD1D.GPU.VARIANT", 0);
        c.exact = (int32_t)config.get("FD1D.GPU.EXACT", 0);
        c.bs_fused = (int32_t)config.get("FD1D.GPU.BS_FUSED", 0);
        // FD1D.GPU.DEVICES = N: ONE price() call spreads the portfolio over the devices DEVICE ... DEVICE + N - 1
        // (-1: every visible device) -- the handle owns a stream, buffers and pinned staging per device
        c.n_devices = (int32_t)config.get("FD1D.GPU.DEVICES", 0);
        const std::string layout = config.get("FD1D.GPU.LAYOUT", "auto");
        if (layout == "auto")
            c.layout = KW_FD1D_LAYOUT_AUTO;
        else if (layout == "reg")
            c.layout = KW_FD1D_LAYOUT_REG;
        else if (layout == "soa")
            c.layout = KW_FD1D_LAYOUT_SOA;
        else
            return "Fd1dGpu_Pricer::init: unknown FD1D.GPU.LAYOUT = " + layout;
        const std::string prec = config.get("FD1D.GPU.PRECISION", "f64");
        if (prec == "f64")
            c.precision = KW_FD1D_F64;
        else if (prec == "f32")
            c.precision = KW_FD1D_F32;
        else
            return "Fd1dGpu_Pricer::init: unknown FD1D.GPU.PRECISION = " + prec;
        if (kw_fd1d_create(&c, &m_handle) != KW_FD1D_OK) {
            Error e = m_handle ? kw_fd1d_last_error(m_handle) : "Fd1dGpu_Pricer::init: failed";
            kw_fd1d_destroy(m_handle);
            m_handle = nullptr;
            return e;
        }
        return "";
    }

    // Fd1d_Pricer::price (src/Pricer/kwFd1d.cpp:21-160): n == 0 -> "" and prices untouched
    Error price(const std::vector<Option>& assets, std::vector<double>& prices) override
    {
        if (!m_handle) return "Fd1dGpu_Pricer::price: pricer was not initialised";
        if (assets.empty()) return "";
        prices.resize(assets.size());
        const int rc = run(reinterpret_cast<const kw_option*>(assets.data()), assets.size(), prices.data());
        return rc == KW_FD1D_OK ? "" : Error(kw_fd1d_last_error(m_handle));
    }

    kw_fd1d_info info() const
    {
        kw_fd1d_info i{};
        if (m_handle) kw_fd1d_get_info(m_handle, &i);
        return i;
    }
};

class Fd1dGpu_BlackScholes_Pricer : public Fd1dGpu_Pricer {
protected:
    int run(const kw_option* a, size_t n, double* out) override { return kw_fd1d_price_bs(m_handle, a, n, out); }
};

// PricerFactory::create (src/Pricer/kwPricerFactory.h:15-41) with the two GPU modes.  Inside the
// reference, add the same two `else if` branches to its factory instead.
class GpuPricerFactory {
public:
    static Error create(const Config& config, sPtr<Pricer>& pricer)
    {
        const std::string mode = config.get("PRICER", "");
        if (mode.empty()) return "PricerFactory: Missing PRICER key";
        if (mode == "FD1D-GPU")
            pricer = make_sPtr<Fd1dGpu_Pricer>();
        else if (mode == "FD1D-BS-GPU")
            pricer = make_sPtr<Fd1dGpu_BlackScholes_Pricer>();
        else
            return "PricerFactory: Unknown PRICER = " + mode;
        if (auto err = pricer->init(config); !err.empty()) return "PricerFactory: " + err;
        return "";
    }
};

}  // namespace kw
