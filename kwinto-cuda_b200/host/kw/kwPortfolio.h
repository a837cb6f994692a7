// kwPortfolio.h -- the caller side of the Fd1d path: kw::Portfolio of the reference
// (src/Utils/kwPortfolio.h:12-36, src/Utils/kwPortfolio.cpp:13-163) with the GPU pricer factory
// behind price().  Same column names, same parsing (std::from_chars through a comma split), the
// same "Price Statistics" block.  Two additions: load() also accepts the fixtures as shipped
// (*.csv.zst, decoded in memory) and the -e tolerance of the CLI (src/kwinto.cpp:27, parsed there
// but never forwarded -- the reference hard-wires 0.5, src/Utils/kwPortfolio.cpp:110) can be
// passed explicitly; the default keeps the reference's 0.5.
#pragma once
#include <charconv>
#include <cmath>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "kwFd1dGpu.h"
#include "kwZstd.h"

namespace kw {

// kw::split / kw::fromString, src/Core/kwString.h:15-49
inline std::vector<std::string> splitLine(const std::string& src, char delim)
{
    std::vector<std::string> result;
    std::stringstream buf(src);
    for (std::string item; std::getline(buf, item, delim);) result.push_back(item);
    return result;
}
template <class T>
inline Error parseNumber(const std::string& src, T& value)
{
    if (std::from_chars(src.data(), src.data() + src.size(), value).ec != std::errc{})
        return "kw::fromString : Fail to parse '" + src + "'";
    return "";
}

// operator<<(ostream, Option), src/Core/kwAsset.cpp:3-8
inline std::string optionAsString(const Option& o)
{
    std::ostringstream os;
    os << "<Option s=" << o.s << " t=" << o.t << ", k=" << o.k << ", z=" << o.z << ", r=" << o.r << ", q=" << o.q
       << ", " << (o.e ? "amer" : "euro") << ", " << (o.w > 0 ? "call" : "put") << ">";
    return os.str();
}

struct PriceStats {  // what printPricesStats computes (src/Utils/kwPortfolio.cpp:101-163)
    f64 rmse = 0, rrmse = 0, mae = 0, mre = 0;
    Option maeAsset{}, mreAsset{};
    std::uint64_t total = 0;
};

class Portfolio {
    std::vector<Option> m_assets;
    std::vector<f64> m_prices;

    Error parse(std::istream& src)
    {
        int e, k, q, r, s, t, v, w, z;
        {
            std::string header;
            std::getline(src, header);
            if (!header.empty() && header.back() == '\r') header.pop_back();
            int i = 0;
            e = k = q = r = s = t = v = w = z = -1;
            for (const auto& colName : splitLine(header, ',')) {
                if (colName == "exercise")
                    e = i;
                else if (colName == "strike")
                    k = i;
                else if (colName == "dividend_rate")
                    q = i;
                else if (colName == "interest_rate")
                    r = i;
                else if (colName == "spot")
                    s = i;
                else if (colName == "expiry")
                    t = i;
                else if (colName == "price")
                    v = i;
                else if (colName == "parity")
                    w = i;
                else if (colName == "volatility")
                    z = i;
                ++i;
            }
            if (e == -1 || k == -1 || q == -1 || r == -1 || s == -1 || t == -1 || v == -1 || w == -1 || z == -1) {
                std::stringstream error;
                error << "Portfolio::load : Some option data is missing: e=" << e << ", k=" << k << ", q=" << q
                      << ", r=" << r << ", s=" << s << ", t=" << t << ", v=" << v << ", w=" << w << ", z=" << z;
                return error.str();
            }
        }
        const size_t need = (size_t)std::max({e, k, q, r, s, t, v, w, z}) + 1;
        for (std::string line; std::getline(src, line);) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            if (line.empty()) continue;
            auto vals = splitLine(line, ',');
            if (vals.size() < need) return "Portfolio::load : Short row '" + line + "'";
            Option asset{};
            // the reference ignores parse errors (src/Utils/kwPortfolio.cpp:64-69); so do we
            parseNumber(vals[k], asset.k);
            parseNumber(vals[q], asset.q);
            parseNumber(vals[r], asset.r);
            parseNumber(vals[t], asset.t);
            parseNumber(vals[s], asset.s);
            parseNumber(vals[z], asset.z);
            asset.e = vals[e] == "a" ? 1 : 0;
            asset.w = vals[w] == "c" ? +1 : -1;
            m_assets.push_back(asset);
            f64 price = 0;
            parseNumber(vals[v], price);
            m_prices.push_back(price);
        }
        return "";
    }

public:
    const std::vector<Option>& assets() const { return m_assets; }
    const std::vector<f64>& prices() const { return m_prices; }

    // Portfolio::load, src/Utils/kwPortfolio.cpp:13-84; *.zst is decoded in memory first
    Error load(const std::string& srcPath_)
    {
        const auto srcPath = std::filesystem::absolute(srcPath_);
        if (srcPath.extension() == ".zst") {
            std::string text;
            if (auto err = zstdDecodeFile(srcPath.string(), text); !err.empty()) return "Portfolio::load : " + err;
            std::istringstream src(text);
            return parse(src);
        }
        std::ifstream src(srcPath);
        if (!src.is_open()) return "Portfolio::load : Failed to open " + srcPath.string();
        return parse(src);
    }

    // Portfolio::price, src/Utils/kwPortfolio.cpp:87-98, on the GPU factory
    Error price(const Config& config, std::vector<f64>& prices)
    {
        sPtr<Pricer> pricer;
        if (auto err = GpuPricerFactory::create(config, pricer); !err.empty()) return "Portfolio::price : " + err;
        if (auto err = pricer->price(m_assets, prices); !err.empty()) return "Portfolio::price : " + err;
        return "";
    }

    // the arithmetic of printPricesStats, src/Utils/kwPortfolio.cpp:101-146, over [first, first + count)
    PriceStats stats(const std::vector<f64>& prices, f64 tolerance = 0.5, size_t first = 0,
                     size_t count = (size_t)-1) const
    {
        f64 absDiffSum1 = 0, absDiffSum2 = 0, relDiffSum1 = 0, relDiffSum2 = 0;
        PriceStats st;
        const size_t end = std::min(m_prices.size(), count == (size_t)-1 ? m_prices.size() : first + count);
        for (size_t j = first; j < end; j++) {
            const auto& wantPrice = m_prices[j];
            const auto gotPrice = prices[j - first];
            if (wantPrice < tolerance) continue;
            const double absDiff = std::abs(wantPrice - gotPrice);
            const double relDiff = absDiff / wantPrice;
            if (absDiff > st.mae) {
                st.mae = absDiff;
                st.maeAsset = m_assets[j];
            }
            if (relDiff > st.mre) {
                st.mre = relDiff;
                st.mreAsset = m_assets[j];
            }
            absDiffSum1 += absDiff;
            absDiffSum2 += absDiff * absDiff;
            relDiffSum1 += relDiff;
            relDiffSum2 += relDiff * relDiff;
            st.total++;
        }
        const auto absDiffMean = absDiffSum1 / st.total;
        st.rmse = std::sqrt(absDiffSum2 / st.total - absDiffMean * absDiffMean);
        const auto relDiffMean = relDiffSum1 / st.total;
        st.rrmse = std::sqrt(relDiffSum2 / st.total - relDiffMean * relDiffMean);
        return st;
    }

    // Portfolio::printPricesStats, src/Utils/kwPortfolio.cpp:148-160: same block, same formatting
    Error printPricesStats(const std::vector<f64>& prices, f64 tolerance = 0.5, std::ostream& os = std::cout) const
    {
        if (prices.size() < m_prices.size()) return "Portfolio::printPricesStats : fewer prices than assets";
        const PriceStats st = stats(prices, tolerance);
        os << "Price Statistics\n";
        os << std::scientific;
        os << "       RMSE : " << st.rmse << std::endl;
        os << "      RRMSE : " << st.rrmse << std::endl;
        os << "        MAE : " << st.mae << std::endl;
        os << "        MRE : " << st.mre << std::endl;
        os << "  MAE Asset : " << optionAsString(st.maeAsset) << std::endl;
        os << "  MRE Asset : " << optionAsString(st.mreAsset) << std::endl;
        os << std::fixed;
        os << "      total : " << st.total << " options" << std::endl;
        os << std::endl;
        return "";
    }
};

}  // namespace kw
