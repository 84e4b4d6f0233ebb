// vela_jit.h - NVRTC specialisation of the chain kernel (see vela_jit.cu)
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

namespace vela {

// Compile `src` (the translation unit produced by spec_source() in vela_api.cu) to a cubin for sm_100a.
// Needs libnvrtc but no GPU.  Returns 0 on success; `log` holds the compiler log / the reason otherwise.
int jit_compile_cubin(const std::string& src, std::vector<char>& cubin, std::string& log);

// Compile (cached per process by source text), load and return the `vela_chain_spec` and `vela_prologue_spec` kernels.
int jit_chain_kernel(const std::string& src, cudaKernel_t* out, cudaKernel_t* prologue, std::string& log);

// True when the ptxas statistics in `log` (compiled with -v) report register spills.
bool jit_log_spills(const std::string& log);

}  // namespace vela
