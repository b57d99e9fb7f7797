// kin_jit.cpp -- see kin_jit.hpp.
#include "kin_jit.hpp"

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>

// objects made by `ld -r -b binary` from the two header files (kinematics.jl_b200/lib.py: build)
extern "C" {
extern const char _binary_kin_device_math_cuh_start[], _binary_kin_device_math_cuh_end[];
extern const char _binary_kin_gen_skeleton_cuh_start[], _binary_kin_gen_skeleton_cuh_end[];
}

namespace kin {

// development aid: KIN_JIT_SRC_DIR=<csrc directory> reads the two texts from disk instead of the embedded copies
static bool read_override(const char *name, std::string &out) {
    const char *d = std::getenv("KIN_JIT_SRC_DIR");
    if (!d || !*d) return false;
    std::ifstream f(std::string(d) + "/" + name, std::ios::binary);
    if (!f) return false;
    out.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    return true;
}
std::string embedded_device_math() {
    std::string s;
    if (read_override("kin_device_math.cuh", s)) return s;
    return std::string(_binary_kin_device_math_cuh_start, _binary_kin_device_math_cuh_end);
}
std::string embedded_gen_skeleton() {
    std::string s;
    if (read_override("kin_gen_skeleton.cuh", s)) return s;
    return std::string(_binary_kin_gen_skeleton_cuh_start, _binary_kin_gen_skeleton_cuh_end);
}

namespace {

typedef struct _nvrtcProgram *nvrtcProgram;
typedef int nvrtcResult;

struct Nvrtc {
    void *handle = nullptr;
    std::string status = "not loaded";
    nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram *) = nullptr;
    nvrtcResult (*Version)(int *, int *) = nullptr;
    const char *(*GetErrorString)(nvrtcResult) = nullptr;
};

Nvrtc &nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        std::vector<std::string> names;
        if (const char *e = std::getenv("KIN_NVRTC_PATH")) names.push_back(e);
        for (const char *s : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so",
                              "libnvrtc.so.13"})
            names.push_back(s);
        std::string tried;
        for (const std::string &nm : names) {
            n.handle = dlopen(nm.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (n.handle) { n.status = nm; break; }
            tried += nm + " ";
        }
        if (!n.handle) { n.status = "unavailable: none of [" + tried + "] could be loaded (set KIN_NVRTC_PATH)"; return; }
        bool ok = true;
        auto sym = [&](const char *name) { void *p = dlsym(n.handle, name); if (!p) ok = false; return p; };
        n.CreateProgram = (decltype(n.CreateProgram))sym("nvrtcCreateProgram");
        n.CompileProgram = (decltype(n.CompileProgram))sym("nvrtcCompileProgram");
        n.GetCUBINSize = (decltype(n.GetCUBINSize))sym("nvrtcGetCUBINSize");
        n.GetCUBIN = (decltype(n.GetCUBIN))sym("nvrtcGetCUBIN");
        n.GetProgramLogSize = (decltype(n.GetProgramLogSize))sym("nvrtcGetProgramLogSize");
        n.GetProgramLog = (decltype(n.GetProgramLog))sym("nvrtcGetProgramLog");
        n.DestroyProgram = (decltype(n.DestroyProgram))sym("nvrtcDestroyProgram");
        n.Version = (decltype(n.Version))sym("nvrtcVersion");
        n.GetErrorString = (decltype(n.GetErrorString))sym("nvrtcGetErrorString");
        if (!ok) { n.status = "unavailable: " + n.status + " lacks a required NVRTC symbol"; dlclose(n.handle); n.handle = nullptr; return; }
        int major = 0, minor = 0;
        n.Version(&major, &minor);
        n.status = "ok: " + n.status + " (" + std::to_string(major) + "." + std::to_string(minor) + ")";
    });
    return n;
}

// FNV-1a 64 over all inputs, twice with different offsets: a 128-bit file name
std::string hash_hex(const std::string &a) {
    uint64_t h1 = 1469598103934665603ull, h2 = 0x9e3779b97f4a7c15ull;
    for (unsigned char c : a) { h1 = (h1 ^ c) * 1099511628211ull; h2 = (h2 ^ (c + 0x7f)) * 0x100000001b3ull + 0x632be5ab; }
    char b[40];
    std::snprintf(b, sizeof b, "%016llx%016llx", (unsigned long long)h1, (unsigned long long)h2);
    return b;
}

std::string cache_dir() {
    const char *e = std::getenv("KIN_JIT_CACHE_DIR");
    if (e && std::string(e) == "off") return "";
    std::string d = e ? std::string(e) : "/tmp/kin_b200_jit-" + std::to_string((unsigned)getuid());
    mkdir(d.c_str(), 0700);
    return d;
}

}  // namespace

std::string jit_status() { return nvrtc().status; }

bool jit_compile(const std::string &src, const JitHeaders &headers, const std::string &arch, std::vector<char> &cubin,
                 std::string &log, bool *from_cache) {
    if (from_cache) *from_cache = false;
    Nvrtc &n = nvrtc();
    if (!n.handle) { log = n.status; return false; }
    std::string all = n.status + "\n" + arch + "\n" + src;
    for (const auto &hd : headers) all += "\n//@" + hd.first + "\n" + hd.second;
    const std::string dir = cache_dir(), path = dir.empty() ? "" : dir + "/" + hash_hex(all) + ".cubin";
    if (!path.empty()) {
        std::ifstream f(path, std::ios::binary);
        if (f) {
            cubin.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
            if (cubin.size() > 64) { if (from_cache) *from_cache = true; return true; }
        }
    }
    std::vector<const char *> htext, hname;
    for (const auto &hd : headers) { hname.push_back(hd.first.c_str()); htext.push_back(hd.second.c_str()); }
    nvrtcProgram prog = nullptr;
    nvrtcResult r = n.CreateProgram(&prog, src.c_str(), "kin_gen.cu", (int)headers.size(), htext.data(), hname.data());
    if (r != 0) { log = std::string("nvrtcCreateProgram: ") + n.GetErrorString(r); return false; }
    const std::string a = "--gpu-architecture=" + arch;
    const char *opts[] = {a.c_str(), "-std=c++17", "-lineinfo"};
    r = n.CompileProgram(prog, 3, opts);
    size_t ls = 0;
    n.GetProgramLogSize(prog, &ls);
    if (ls > 1) { log.resize(ls); n.GetProgramLog(prog, &log[0]); }
    if (r != 0) {
        log = std::string("nvrtcCompileProgram: ") + n.GetErrorString(r) + "\n" + log;
        n.DestroyProgram(&prog);
        return false;
    }
    size_t cs = 0;
    n.GetCUBINSize(prog, &cs);
    cubin.resize(cs);
    n.GetCUBIN(prog, cubin.data());
    n.DestroyProgram(&prog);
    if (!path.empty()) {
        const std::string tmp = path + "." + std::to_string((long)getpid());
        std::ofstream f(tmp, std::ios::binary);
        if (f) { f.write(cubin.data(), (std::streamsize)cubin.size()); f.close(); std::rename(tmp.c_str(), path.c_str()); }
    }
    return cs > 0;
}

}  // namespace kin
