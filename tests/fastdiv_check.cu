// Host-side check of the division-free site decode of the batched kernels (chroma_b200/csrc/common.cuh: FastDiv,
// make_fastdiv, fast_div): exact quotients for every divisor / dividend class the kernels can meet (n < 2^28).
// Built with nvcc and run on the CPU by tests/test_fastdiv.py.
#include <cstdio>
#include <cstdlib>
#include "../chroma_b200/csrc/common.cuh"

int main() {
  using namespace b200;
  long checked = 0;
  unsigned long long seed = 88172645463325252ull;
  auto rnd = [&]() { seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17; return seed; };
  const int NMAX = 1 << 28;
  auto check = [&](int n, int d) {
    if (n < 0 || n >= NMAX) return true;
    const FastDiv f = make_fastdiv(d);
    ++checked;
    if (fast_div(n, f) != n / d) { std::printf("FAIL n=%d d=%d got %d want %d\n", n, d, fast_div(n, f), n / d); return false; }
    return true;
  };
  for (int d = 1; d <= 4096; ++d) {                       // every extent / row / slice size of a small lattice, exhaustively at the edges
    for (int n = 0; n < 3 * d + 2; ++n) if (!check(n, d)) return 1;
    for (int m : {7, 1000, 65535, NMAX / d - 1, NMAX / d}) for (int e = -1; e <= 1; ++e) if (!check(m * d + e, d)) return 1;
    if (!check(NMAX - 1, d)) return 1;
  }
  for (int i = 0; i < 200000; ++i) {                      // random divisors up to 2^28 (time-slice and chunk sizes of big lattices)
    const int d = 1 + (int)(rnd() % (unsigned long long)NMAX);
    for (int j = 0; j < 8; ++j) if (!check((int)(rnd() % (unsigned long long)NMAX), d)) return 1;
    for (int m : {1, 2, 3}) for (int e = -1; e <= 1; ++e) if (!check(m * d + e, d)) return 1;
  }
  // the decode of a 64^3 x 128 lattice, every site: idx -> (xh, y, z, t) as mrhs_site does it
  {
    const int Lxh = 32, Ly = 64, Lz = 64, Lt = 128;
    const FastDiv fx = make_fastdiv(Lxh), fy = make_fastdiv(Ly), fz = make_fastdiv(Lz);
    for (int idx = 0; idx < Lxh * Ly * Lz * Lt; idx += 1) {
      const int q = fast_div(idx, fx), xh = idx - q * Lxh;
      const int q2 = fast_div(q, fy), y = q - q2 * Ly;
      const int t = fast_div(q2, fz), z = q2 - t * Lz;
      if (((t * Lz + z) * Ly + y) * Lxh + xh != idx || xh < 0 || xh >= Lxh || y < 0 || y >= Ly || z < 0 || z >= Lz) { std::printf("FAIL decode %d\n", idx); return 1; }
      ++checked;
    }
  }
  std::printf("FASTDIV OK %ld\n", checked);
  return 0;
}
