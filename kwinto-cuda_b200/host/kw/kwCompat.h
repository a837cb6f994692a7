// kwCompat.h -- the slice of the reference's Core/Pricer interface that the Fd1d path needs,
// so the GPU pricer can be built and tested WITHOUT the reference tree.
//
// When compiling inside the reference, define KW_WITH_REFERENCE and the
// reference's own headers are used instead:
//     kw::Option   src/Core/kwAsset.h:12-45     kw::Config  src/Core/kwConfig.h:21-73
//     kw::Error    src/Core/kwTypes.h:18        kw::Pricer  src/Pricer/kwPricer.h:12-22
// The stand-ins below keep the same names, member order and semantics (typed config lookup
// picked by the default's type, Error == std::string with "" meaning success).
#pragma once

#ifdef KW_WITH_REFERENCE
#include "Core/kwAsset.h"
#include "Core/kwConfig.h"
#include "Pricer/kwPricer.h"
#else

#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

namespace kw {

using Error = std::string;
using f64 = double;
using i64 = std::int64_t;
using u8 = std::uint8_t;
using i8 = std::int8_t;
template <class T>
using sPtr = std::shared_ptr<T>;
template <class T, class... A>
sPtr<T> make_sPtr(A&&... a)
{
    return std::make_shared<T>(std::forward<A>(a)...);
}

// aggregate-init order t,k,z,r,q,s,e,w (test/kwPricer_test.cpp:16)
struct Option {
    f64 t, k, z, r, q, s;
    u8 e;  // 0 European, 1 American
    i8 w;  // -1 put, +1 call
};

class Config {
    std::map<std::string, f64> dbl_;
    std::map<std::string, i64> int_;
    std::map<std::string, std::string> str_;

public:
    template <class T>
    void set(const std::string& key, const T& v)
    {
        if constexpr (std::is_floating_point_v<T>)
            dbl_[key] = v;
        else if constexpr (std::is_integral_v<T>)
            int_[key] = v;
        else
            str_[key] = std::string(v);
    }
    template <class T>
    auto get(const std::string& key, const T& dflt) const
    {
        if constexpr (std::is_floating_point_v<T>) {
            auto it = dbl_.find(key);
            return it == dbl_.end() ? f64(dflt) : it->second;
        } else if constexpr (std::is_integral_v<T>) {
            auto it = int_.find(key);
            return it == int_.end() ? i64(dflt) : it->second;
        } else {
            auto it = str_.find(key);
            return it == str_.end() ? std::string(dflt) : it->second;
        }
    }
    void erase(const std::string& key)
    {
        dbl_.erase(key);
        int_.erase(key);
        str_.erase(key);
    }
};

class Pricer {
public:
    virtual Error init(const Config& config) = 0;
    virtual Error price(const std::vector<Option>& assets, std::vector<double>& prices) = 0;
    virtual ~Pricer() = default;
};

}  // namespace kw
#endif
