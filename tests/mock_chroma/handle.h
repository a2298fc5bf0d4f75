#ifndef MOCK_HANDLE_H
#define MOCK_HANDLE_H
namespace Chroma {
template <typename T> class Handle {   // lib/handle.h:37-92: reference-counted owner
 public:
  Handle() : p(0), n(0) {}
  Handle(T* q) : p(q), n(q ? new int(1) : 0) {}
  Handle(const Handle& o) : p(o.p), n(o.n) { if (n) ++*n; }
  Handle& operator=(const Handle& o) { if (this != &o) { release(); p = o.p; n = o.n; if (n) ++*n; } return *this; }
  ~Handle() { release(); }
  T& operator*() const { return *p; }
  T* operator->() const { return p; }
 private:
  void release() { if (n && --*n == 0) { delete p; delete n; } p = 0; n = 0; }
  T* p; int* n;
};
}
#endif
