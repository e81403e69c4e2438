// ptrs.h — scope guards for the handles of moshi.h (the reference ships include/moshi/ptrs.h:1-100 with the same two names).
#pragma once
#include <vector>

// deletes the object when the guard goes out of scope
template <typename T>
class own_ptr {
public:
    T *ptr = nullptr;
    own_ptr() = default;
    own_ptr(T *p) : ptr(p) {}
    own_ptr(const own_ptr &) = delete;
    own_ptr &operator=(const own_ptr &) = delete;
    ~own_ptr() { reset(); }
    void reset() { delete ptr; ptr = nullptr; }
    own_ptr &operator=(T *p) { if (p != ptr) { reset(); ptr = p; } return *this; }
    T &operator*() { return *ptr; }
    T *operator->() { return ptr; }
    operator T *() { return ptr; }
    operator const T *() const { return ptr; }
};

// calls the matching unref(T *) overload of moshi.h when the guard goes out of scope
template <typename T>
class unref_ptr {
public:
    T *ptr = nullptr;
    unref_ptr() = default;
    unref_ptr(T *p) : ptr(p) {}
    unref_ptr(const unref_ptr &) = delete;
    unref_ptr &operator=(const unref_ptr &) = delete;
    ~unref_ptr() { reset(); }
    void reset() { if (ptr) { unref(ptr); ptr = nullptr; } }
    unref_ptr &operator=(T *p) { if (p != ptr) { reset(); ptr = p; } return *this; }
    T &operator*() { return *ptr; }
    T *operator->() { return ptr; }
    operator T *() { return ptr; }
    operator const T *() const { return ptr; }
};
