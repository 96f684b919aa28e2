// Minimal stand-in for boost::alignment::aligned_allocator (test infrastructure only).
// Lets the reference's own containers (data_structures/Vector.h:7-14) compile where the
// Boost headers (R package BH) are absent.  Not product code.
#ifndef COGAPS_B200_SHIM_ALIGNED_ALLOCATOR_HPP
#define COGAPS_B200_SHIM_ALIGNED_ALLOCATOR_HPP
#include <cstddef>
#include <cstdlib>
#include <new>
namespace boost { namespace alignment {
template <class T, std::size_t Alignment>
struct aligned_allocator
{
    typedef T value_type;
    typedef T* pointer;
    typedef const T* const_pointer;
    typedef T& reference;
    typedef const T& const_reference;
    typedef std::size_t size_type;
    typedef std::ptrdiff_t difference_type;
    template <class U> struct rebind { typedef aligned_allocator<U, Alignment> other; };
    aligned_allocator() {}
    template <class U> aligned_allocator(const aligned_allocator<U, Alignment>&) {}
    pointer allocate(size_type n, const void* = 0)
    {
        void *p = 0;
        std::size_t a = Alignment < sizeof(void*) ? sizeof(void*) : Alignment;
        if (n == 0) { return 0; }
        if (posix_memalign(&p, a, n * sizeof(T)) != 0) { throw std::bad_alloc(); }
        return static_cast<pointer>(p);
    }
    void deallocate(pointer p, size_type) { std::free(p); }
    size_type max_size() const { return static_cast<size_type>(-1) / sizeof(T); }
    void construct(pointer p, const T &v) { new (static_cast<void*>(p)) T(v); }
    void destroy(pointer p) { p->~T(); }
};
template <class T, class U, std::size_t A>
bool operator==(const aligned_allocator<T, A>&, const aligned_allocator<U, A>&) { return true; }
template <class T, class U, std::size_t A>
bool operator!=(const aligned_allocator<T, A>&, const aligned_allocator<U, A>&) { return false; }
}} // namespace boost::alignment
#endif
