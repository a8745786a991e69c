// Run-time compilation of generated CUDA source (mapgen.h) with NVRTC, loading through the CUDA driver API.
// libnvrtc and libcuda are opened with dlopen at first use: the library has no link-time dependency on either,
// so it builds and loads on a machine without a GPU.  Compiled cubins are kept in a directory cache keyed by a
// hash of (source, options, NVRTC version); __graft_entry__.build() fills that cache for the shipped optics, so a
// GPU box normally loads a prebuilt cubin and never calls the compiler.
#pragma once
#include <string>
#include <vector>

namespace simc {

struct JitModule {
  void* module = nullptr;                 // CUmodule
  std::vector<void*> fns;                 // CUfunction per requested name
  bool from_cache = false;
};

// Compiles `src` for sm_100a (strict IEEE: --fmad=false) and returns the cubin.  No GPU needed.
// cache_dir may be empty (no cache).  Returns false and fills err on failure.
bool jit_compile_cubin(const std::string& src, const std::string& cache_dir, std::string& cubin, bool* from_cache, std::string& err);
// Loads a cubin into the current context and looks up the kernels.
bool jit_load(const std::string& cubin, const std::vector<std::string>& names, JitModule& out, std::string& err);
void jit_unload(JitModule& m);
// cuLaunchKernel on a runtime-API stream.  Returns a CUresult (0 = success); err text via jit_error_string.
int jit_launch(void* fn, unsigned grid, unsigned block, void* stream, void** args);
std::string jit_error_string(int cu_result);
// <directory of this shared library>/jit_cache, or $SIMC_B200_JIT_CACHE
std::string jit_default_cache_dir();

}  // namespace simc
