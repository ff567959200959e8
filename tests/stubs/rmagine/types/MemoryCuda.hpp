// STUB of rmagine/types/MemoryCuda.hpp (tests/stubs/README.md): the VRAM_CUDA memory space (cudaMalloc / cudaFree / device-to-device copy) and
// the host<->device assignment helpers the tests use.
#pragma once
#include <cuda_runtime_api.h>

#include "Memory.hpp"

namespace rmagine {

struct VRAM_CUDA {
    template <typename T> static T* alloc(size_t n) { void* p = nullptr; cudaMalloc(&p, sizeof(T) * (n ? n : 1)); return static_cast<T*>(p); }
    template <typename T> static void free(T* p, size_t) { cudaFree(p); }
    template <typename T> static void copy(T* dst, const T* src, size_t n) { cudaMemcpy(dst, src, sizeof(T) * n, cudaMemcpyDeviceToDevice); }
};
template <typename T> void upload(MemoryView<T, VRAM_CUDA>& dst, const MemoryView<T, RAM>& src) { cudaMemcpy(dst.raw(), src.raw(), sizeof(T) * src.size(), cudaMemcpyHostToDevice); }
template <typename T> void download(MemoryView<T, RAM>& dst, const MemoryView<T, VRAM_CUDA>& src) { cudaMemcpy(dst.raw(), src.raw(), sizeof(T) * src.size(), cudaMemcpyDeviceToHost); }

}  // namespace rmagine
