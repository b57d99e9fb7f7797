// kin_jit.hpp -- run-time compilation of the model-specialised kernels (kin_codegen.hpp) with NVRTC for sm_100a.
// libnvrtc is loaded with dlopen on first use (no link-time dependency); when it cannot be found the library keeps
// using its ahead-of-time (interpreting) kernels and says so in kin_jit_status().
#pragma once
#include <string>
#include <utility>
#include <vector>

namespace kin {

// (include name, text) pairs offered to the compiler as in-memory headers
typedef std::vector<std::pair<std::string, std::string>> JitHeaders;

// Compiles `src` to a cubin for `arch` (e.g. "sm_100a").  Results are cached on disk under $KIN_JIT_CACHE_DIR
// (default /tmp/kin_b200_jit-<uid>; set it to "off" to disable), keyed by a hash of the complete input.
bool jit_compile(const std::string &src, const JitHeaders &headers, const std::string &arch, std::vector<char> &cubin,
                 std::string &log, bool *from_cache);

// "ok: libnvrtc.so.12 (12.9)" or the reason NVRTC is unavailable
std::string jit_status();

// the embedded texts of kin_device_math.cuh and kin_gen_skeleton.cuh (linked into the library as binary objects)
std::string embedded_device_math();
std::string embedded_gen_skeleton();

}  // namespace kin
