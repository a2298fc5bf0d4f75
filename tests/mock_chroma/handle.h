#ifndef MOCK_HANDLE_H
#define MOCK_HANDLE_H
namespace Chroma {
template <typename T> class Handle {   // lib/handle.h:37-92 (reference counting left out: the test owns the objects)
 public:
  Handle() : p(0) {}
  Handle(T* q) : p(q) {}
  T& operator*() const { return *p; }
  T* operator->() const { return p; }
 private:
  T* p;
};
}
#endif
