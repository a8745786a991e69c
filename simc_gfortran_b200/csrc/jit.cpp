#include "jit.h"
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <sstream>

namespace simc {
namespace {

// ---- NVRTC (nvrtc.h restated as the handful of entry points used) -----------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  void* lib = nullptr;
  int (*Version)(int*, int*) = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string why;
};
// ---- CUDA driver API ---------------------------------------------------------------------------------
struct Driver {
  void* lib = nullptr;
  int (*ModuleLoadData)(void**, const void*) = nullptr;
  int (*ModuleUnload)(void*) = nullptr;
  int (*ModuleGetFunction)(void**, void*, const char*) = nullptr;
  int (*LaunchKernel)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**) = nullptr;
  int (*GetErrorString)(int, const char**) = nullptr;
  std::string why;
};

template <class F>
bool sym(void* lib, const char* name, F& f, std::string& why) {
  f = (F)dlsym(lib, name);
  if (!f) { why = std::string("missing symbol ") + name; return false; }
  return true;
}

Nvrtc& nvrtc() {
  static Nvrtc n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* nm : names) {
      n.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (n.lib) break;
    }
    if (!n.lib) { n.why = "libnvrtc not found (dlopen)"; return; }
    bool ok = sym(n.lib, "nvrtcVersion", n.Version, n.why) && sym(n.lib, "nvrtcCreateProgram", n.CreateProgram, n.why) &&
              sym(n.lib, "nvrtcDestroyProgram", n.DestroyProgram, n.why) && sym(n.lib, "nvrtcCompileProgram", n.CompileProgram, n.why) &&
              sym(n.lib, "nvrtcGetCUBINSize", n.GetCUBINSize, n.why) && sym(n.lib, "nvrtcGetCUBIN", n.GetCUBIN, n.why) &&
              sym(n.lib, "nvrtcGetProgramLogSize", n.GetProgramLogSize, n.why) && sym(n.lib, "nvrtcGetProgramLog", n.GetProgramLog, n.why) &&
              sym(n.lib, "nvrtcGetErrorString", n.GetErrorString, n.why);
    if (!ok) { dlclose(n.lib); n.lib = nullptr; }
  });
  return n;
}

Driver& driver() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    d.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!d.lib) { d.why = "libcuda.so.1 not found (no NVIDIA driver on this machine)"; return; }
    bool ok = sym(d.lib, "cuModuleLoadData", d.ModuleLoadData, d.why) && sym(d.lib, "cuModuleUnload", d.ModuleUnload, d.why) &&
              sym(d.lib, "cuModuleGetFunction", d.ModuleGetFunction, d.why) && sym(d.lib, "cuLaunchKernel", d.LaunchKernel, d.why) &&
              sym(d.lib, "cuGetErrorString", d.GetErrorString, d.why);
    if (!ok) { dlclose(d.lib); d.lib = nullptr; }
  });
  return d;
}

uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ULL) {
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ULL; }
  return h;
}

const char* kOptions[] = {"--gpu-architecture=sm_100a", "--fmad=false", "--std=c++17", "-lineinfo"};
const int kNumOptions = 4;

bool read_file(const std::string& path, std::string& out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::ostringstream ss;
  ss << f.rdbuf();
  out = ss.str();
  return !out.empty();
}

}  // namespace

std::string jit_default_cache_dir() {
  if (const char* e = std::getenv("SIMC_B200_JIT_CACHE")) return e;
  Dl_info info;
  if (dladdr((void*)&jit_default_cache_dir, &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    const size_t k = p.rfind('/');
    if (k != std::string::npos) return p.substr(0, k) + "/jit_cache";
  }
  return "";
}

bool jit_compile_cubin(const std::string& src, const std::string& cache_dir, std::string& cubin, bool* from_cache, std::string& err) {
  if (from_cache) *from_cache = false;
  std::string key_src = src;
  for (int i = 0; i < kNumOptions; ++i) { key_src += '\n'; key_src += kOptions[i]; }
  // two independent 64-bit hashes of the text name the cache file; the NVRTC version is not part of the key, so
  // that a cubin built by the toolkit's nvrtc at build time is found at run time whatever library is present
  char name[64];
  std::snprintf(name, sizeof(name), "%016llx%016llx.cubin", (unsigned long long)fnv1a(key_src),
                (unsigned long long)fnv1a(key_src, 0x9E3779B97F4A7C15ULL));
  const std::string path = cache_dir.empty() ? "" : cache_dir + "/" + name;
  if (!path.empty() && read_file(path, cubin)) {
    if (from_cache) *from_cache = true;
    return true;
  }
  Nvrtc& n = nvrtc();
  if (!n.lib) { err = "map compiler: " + n.why; return false; }
  nvrtcProgram prog = nullptr;
  int rc = n.CreateProgram(&prog, src.c_str(), "simc_b200_maps.cu", 0, nullptr, nullptr);
  if (rc) { err = std::string("nvrtcCreateProgram: ") + n.GetErrorString(rc); return false; }
  rc = n.CompileProgram(prog, kNumOptions, kOptions);
  if (rc) {
    size_t ls = 0;
    n.GetProgramLogSize(prog, &ls);
    std::string log(ls, '\0');
    if (ls) n.GetProgramLog(prog, &log[0]);
    err = std::string("nvrtcCompileProgram: ") + n.GetErrorString(rc) + "\n" + log.substr(0, 4000);
    n.DestroyProgram(&prog);
    return false;
  }
  size_t sz = 0;
  rc = n.GetCUBINSize(prog, &sz);
  if (rc || sz == 0) { err = "nvrtcGetCUBINSize failed"; n.DestroyProgram(&prog); return false; }
  cubin.assign(sz, '\0');
  rc = n.GetCUBIN(prog, &cubin[0]);
  n.DestroyProgram(&prog);
  if (rc) { err = "nvrtcGetCUBIN failed"; return false; }
  if (!path.empty()) {
    mkdir(cache_dir.c_str(), 0755);
    const std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
    std::ofstream f(tmp, std::ios::binary);
    if (f) {
      f.write(cubin.data(), (std::streamsize)cubin.size());
      f.close();
      if (std::rename(tmp.c_str(), path.c_str()) != 0) std::remove(tmp.c_str());
    }
  }
  return true;
}

bool jit_load(const std::string& cubin, const std::vector<std::string>& names, JitModule& out, std::string& err) {
  Driver& d = driver();
  if (!d.lib) { err = "map compiler: " + d.why; return false; }
  void* mod = nullptr;
  int rc = d.ModuleLoadData(&mod, cubin.data());
  if (rc) { err = "cuModuleLoadData: " + jit_error_string(rc); return false; }
  out.module = mod;
  out.fns.clear();
  for (const std::string& nm : names) {
    void* fn = nullptr;
    rc = d.ModuleGetFunction(&fn, mod, nm.c_str());
    if (rc) { err = "cuModuleGetFunction(" + nm + "): " + jit_error_string(rc); d.ModuleUnload(mod); out.module = nullptr; return false; }
    out.fns.push_back(fn);
  }
  return true;
}

void jit_unload(JitModule& m) {
  if (m.module && driver().lib) driver().ModuleUnload(m.module);
  m.module = nullptr;
  m.fns.clear();
}

int jit_launch(void* fn, unsigned grid, unsigned block, void* stream, void** args) {
  Driver& d = driver();
  if (!d.lib) return 999;
  return d.LaunchKernel(fn, grid, 1, 1, block, 1, 1, 0, stream, args, nullptr);
}

std::string jit_error_string(int rc) {
  Driver& d = driver();
  const char* s = nullptr;
  if (d.lib && d.GetErrorString(rc, &s) == 0 && s) return s;
  return "CUDA driver error " + std::to_string(rc);
}

}  // namespace simc
