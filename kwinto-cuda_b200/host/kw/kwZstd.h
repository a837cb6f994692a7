// kwZstd.h -- decode a .zst file into memory through the system's libzstd.so.1, bound at run time
// (dlopen; the image has the library but no zstd headers or CLI).  The reference ships its fixtures
// as test/portfolio_*.csv.zst and expects them to be decoded by hand (`zstdcat`, Makefile check-all)
// because Portfolio::load reads plain CSV only (src/Utils/kwPortfolio.cpp:17); the GPU driver reads
// both.
#pragma once
#include <dlfcn.h>

#include <cstdio>
#include <string>
#include <vector>

namespace kw {

inline std::string zstdDecodeFile(const std::string& path, std::string& out)
{
    std::FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return "kw::zstdDecodeFile : Failed to open " + path;
    std::vector<char> src;
    char buf[1 << 16];
    for (size_t n; (n = std::fread(buf, 1, sizeof buf, f)) > 0;) src.insert(src.end(), buf, buf + n);
    std::fclose(f);

    void* lib = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!lib) return "kw::zstdDecodeFile : libzstd.so.1 not found; decode " + path + " first (zstdcat)";
    // the stable part of the zstd API (zstd.h): streaming decompression
    struct InBuf {
        const void* src;
        size_t size, pos;
    };
    struct OutBuf {
        void* dst;
        size_t size, pos;
    };
    using CreateFn = void* (*)();
    using FreeFn = size_t (*)(void*);
    using StepFn = size_t (*)(void*, OutBuf*, InBuf*);
    using IsErrFn = unsigned (*)(size_t);
    using ErrNameFn = const char* (*)(size_t);
    auto create = (CreateFn)dlsym(lib, "ZSTD_createDStream");
    auto destroy = (FreeFn)dlsym(lib, "ZSTD_freeDStream");
    auto step = (StepFn)dlsym(lib, "ZSTD_decompressStream");
    auto isErr = (IsErrFn)dlsym(lib, "ZSTD_isError");
    auto errName = (ErrNameFn)dlsym(lib, "ZSTD_getErrorName");
    if (!create || !destroy || !step || !isErr || !errName) {
        dlclose(lib);
        return "kw::zstdDecodeFile : libzstd.so.1 lacks the streaming API";
    }
    void* ds = create();
    InBuf in{src.data(), src.size(), 0};
    out.clear();
    std::string err;
    std::vector<char> chunk(1 << 17);
    while (in.pos < in.size) {
        OutBuf ob{chunk.data(), chunk.size(), 0};
        const size_t rc = step(ds, &ob, &in);
        if (isErr(rc)) {
            err = "kw::zstdDecodeFile : " + path + ": " + errName(rc);
            break;
        }
        out.append(chunk.data(), ob.pos);
        if (rc == 0 && in.pos == in.size) break;  // frame complete
    }
    destroy(ds);
    dlclose(lib);
    return err;
}

}  // namespace kw
