// STUB of rmagine/types/Memory.hpp (tests/stubs/README.md): MemoryView<T, MemT> (non-owning) and Memory<T, MemT> (owning, only ever resized),
// parameterised by a memory-space tag with static alloc / free like rmagine's RAM / VRAM_CUDA.
#pragma once
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <utility>

namespace rmagine {

struct RAM {
    template <typename T> static T* alloc(size_t n) { return static_cast<T*>(std::malloc(sizeof(T) * (n ? n : 1))); }
    template <typename T> static void free(T* p, size_t) { std::free(p); }
    template <typename T> static void copy(T* dst, const T* src, size_t n) { std::memcpy(dst, src, sizeof(T) * n); }
};

template <typename DataT, typename MemT = RAM> class MemoryView {
public:
    using DataType = DataT; using MemType = MemT;
    MemoryView(DataT* mem, size_t n) : m_mem(mem), m_size(n) {}
    static MemoryView<DataT, MemT> Empty() { return MemoryView<DataT, MemT>(nullptr, 0); }
    DataT* raw() { return m_mem; } const DataT* raw() const { return m_mem; }
    size_t size() const { return m_size; } bool empty() const { return m_mem == nullptr; }
    DataT& operator[](size_t i) { return m_mem[i]; } const DataT& operator[](size_t i) const { return m_mem[i]; }        // meaningful for host memory only
    MemoryView<DataT, MemT> operator()(size_t b, size_t e) { return MemoryView<DataT, MemT>(m_mem + b, e - b); }
    const MemoryView<DataT, MemT> operator()(size_t b, size_t e) const { return MemoryView<DataT, MemT>(m_mem + b, e - b); }
protected:
    DataT* m_mem; size_t m_size;
};

template <typename DataT, typename MemT = RAM> class Memory : public MemoryView<DataT, MemT> {
public:
    using Base = MemoryView<DataT, MemT>;
    Memory() : Base(nullptr, 0) {}
    explicit Memory(size_t n) : Base(MemT::template alloc<DataT>(n), n) {}
    Memory(const Memory& o) : Base(MemT::template alloc<DataT>(o.size()), o.size()) { MemT::template copy<DataT>(this->m_mem, o.raw(), o.size()); }
    Memory(Memory&& o) noexcept : Base(o.m_mem, o.m_size) { o.m_mem = nullptr; o.m_size = 0; }
    Memory& operator=(Memory o) { std::swap(this->m_mem, o.m_mem); std::swap(this->m_size, o.m_size); return *this; }
    ~Memory() { if (this->m_mem) MemT::template free<DataT>(this->m_mem, this->m_size); }
    void resize(size_t n)
    {
        DataT* p = MemT::template alloc<DataT>(n);
        if (this->m_mem) { MemT::template copy<DataT>(p, this->m_mem, n < this->m_size ? n : this->m_size); MemT::template free<DataT>(this->m_mem, this->m_size); }
        this->m_mem = p; this->m_size = n;
    }
};

}  // namespace rmagine
