// Host-only checks of include/myzkp_b200.hpp (no device calls): the scalar negation the C++ mirror uses for the
// verifier equations, printed for comparison with Python integers; and the wrappers must instantiate.
#include <cstdio>
#include <cstring>

#include "myzkp_b200.hpp"

using namespace myzkp_b200;

static void print_le(const Scalar& s) {
  for (int i = 31; i >= 0; i--) printf("%02x", s[i]);
  printf("\n");
}

int main(int argc, char** argv) {
  for (int a = 1; a < argc; a++) {  // each argument: 64 hex digits, big-endian
    Scalar x{};
    if (strlen(argv[a]) != 64) return 2;
    for (int i = 0; i < 32; i++) {
      unsigned v = 0;
      sscanf(argv[a] + 2 * (31 - i), "%2x", &v);
      x[i] = (uint8_t)v;
    }
    print_le(scalar_neg(x));
  }
  (void)&verify_kzg; (void)&batch_open_kzg; (void)&prove_degree_bound; (void)&verify_degree_bound; (void)&powers_2;
  return 0;
}
