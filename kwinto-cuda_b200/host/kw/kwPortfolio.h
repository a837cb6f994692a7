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

struct PriceStats {  // the figures of the reference's "Price Statistics" block (src/Utils/kwPortfolio.cpp:148-160)
    f64 rmse = 0, rrmse = 0, mae = 0, mre = 0;
    Option maeAsset{}, mreAsset{};
    std::uint64_t total = 0;
};

// Population spread sqrt(E[x^2] - E[x]^2) and the running maximum of a stream of deviations, with the asset
// that produced the maximum.  (The two-sum form, not Welford's, is what reproduces the reference's printed digits.)
class DeviationStats {
    f64 m_sum = 0, m_sumSq = 0, m_max = 0;
    std::uint64_t m_n = 0;
    Option m_worst{};

public:
    void add(f64 x, const Option& asset)
    {
        if (x > m_max) {
            m_max = x;
            m_worst = asset;
        }
        m_sum += x;
        m_sumSq += x * x;
        ++m_n;
    }
    std::uint64_t count() const { return m_n; }
    f64 max() const { return m_max; }
    const Option& worst() const { return m_worst; }
    f64 spread() const
    {
        const f64 mean = m_sum / m_n;
        return std::sqrt(m_sumSq / m_n - mean * mean);
    }
};

class Portfolio {
    std::vector<Option> m_assets;
    std::vector<f64> m_prices;

    // One CSV column the loader needs: its header name, the one-letter tag of the reference's error message, and
    // what a cell does to the row being built.  The reference ignores number-parse errors
    // (src/Utils/kwPortfolio.cpp:64-69); so do the setters.
    struct Row {
        Option asset{};
        f64 price = 0;
    };
    struct Column {
        const char* name;
        char tag;
        void (*set)(Row&, const std::string&);
        int index = -1;
    };
    static std::vector<Column> columns()
    {
        return {
            {"exercise", 'e', [](Row& r, const std::string& c) { r.asset.e = c == "a" ? 1 : 0; }},
            {"strike", 'k', [](Row& r, const std::string& c) { parseNumber(c, r.asset.k); }},
            {"dividend_rate", 'q', [](Row& r, const std::string& c) { parseNumber(c, r.asset.q); }},
            {"interest_rate", 'r', [](Row& r, const std::string& c) { parseNumber(c, r.asset.r); }},
            {"spot", 's', [](Row& r, const std::string& c) { parseNumber(c, r.asset.s); }},
            {"expiry", 't', [](Row& r, const std::string& c) { parseNumber(c, r.asset.t); }},
            {"price", 'v', [](Row& r, const std::string& c) { parseNumber(c, r.price); }},
            {"parity", 'w', [](Row& r, const std::string& c) { r.asset.w = c == "c" ? +1 : -1; }},
            {"volatility", 'z', [](Row& r, const std::string& c) { parseNumber(c, r.asset.z); }},
        };
    }
    static void chomp(std::string& line)
    {
        if (!line.empty() && line.back() == '\r') line.pop_back();
    }

    // Portfolio::load's parser (src/Utils/kwPortfolio.cpp:19-83): header names -> column indices, then one Option
    // and one reference price per row.  Columns may come in any order; unknown columns are skipped.
    Error parse(std::istream& src)
    {
        auto cols = columns();
        std::string header;
        std::getline(src, header);
        chomp(header);
        const auto names = splitLine(header, ',');
        size_t need = 0;
        bool missing = false;
        for (auto& col : cols) {
            for (size_t i = 0; i < names.size(); ++i)
                if (names[i] == col.name) col.index = (int)i;
            missing = missing || col.index < 0;
            need = std::max(need, (size_t)(col.index + 1));
        }
        if (missing) {
            // the reference's message lists every column tag with the index it found (-1 = absent)
            std::stringstream error;
            error << "Portfolio::load : Some option data is missing: ";
            for (size_t c = 0; c < cols.size(); ++c) error << (c ? ", " : "") << cols[c].tag << "=" << cols[c].index;
            return error.str();
        }
        for (std::string line; std::getline(src, line);) {
            chomp(line);
            if (line.empty()) continue;
            const auto cells = splitLine(line, ',');
            if (cells.size() < need) return "Portfolio::load : Short row '" + line + "'";
            Row row;
            for (const auto& col : cols) col.set(row, cells[(size_t)col.index]);
            m_assets.push_back(row.asset);
            m_prices.push_back(row.price);
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

    // The figures of printPricesStats (src/Utils/kwPortfolio.cpp:101-146) over the assets [first, first + count):
    // absolute and relative deviation of `prices` from the file's reference prices, assets whose reference price is
    // below `tolerance` left out (the reference hard-wires 0.5, :110).
    PriceStats stats(const std::vector<f64>& prices, f64 tolerance = 0.5, size_t first = 0,
                     size_t count = (size_t)-1) const
    {
        const size_t end = std::min(m_prices.size(), count == (size_t)-1 ? m_prices.size() : first + count);
        DeviationStats absolute, relative;
        for (size_t j = first; j < end; j++) {
            const f64 want = m_prices[j];
            if (want < tolerance) continue;
            const f64 dev = std::abs(want - prices[j - first]);
            absolute.add(dev, m_assets[j]);
            relative.add(dev / want, m_assets[j]);
        }
        PriceStats st;
        st.total = absolute.count();
        st.rmse = absolute.spread();
        st.rrmse = relative.spread();
        st.mae = absolute.max();
        st.mre = relative.max();
        st.maeAsset = absolute.worst();
        st.mreAsset = relative.worst();
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
