// C++ host driving SEVERAL GPUs from one process through the C ABI alone (include/myzkp_b200.hpp -> csrc/multi.cu):
// no torch, no NCCL, one call per commit / open.  usage: abi_multi [log2n] [dev0 dev1 ...]
// With no device list it takes every visible device (twice device 0 on a single-GPU box, i.e. two ranks sharing it).
// Prints the commitment and the opening of a seeded polynomial (the pytest compares them with the oracle) and
// checks them against the single-device path; exits non-zero on any mismatch.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "myzkp_b200.hpp"

using namespace myzkp_b200;

static Scalar scalar_u64(uint64_t v) {
  Scalar s{};
  memcpy(s.data(), &v, 8);
  return s;
}
static std::string hex_be(const uint8_t* le32) {
  char buf[65];
  for (int i = 0; i < 32; i++) snprintf(buf + 2 * i, 3, "%02x", le32[31 - i]);
  return std::string(buf);
}
// splitmix64: the pytest regenerates the same coefficients
static uint64_t sm64(uint64_t& x) {
  uint64_t z = (x += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

int main(int argc, char** argv) {
  try {
    const int log2n = argc > 1 ? atoi(argv[1]) : 16;
    std::vector<int> devices;
    for (int i = 2; i < argc; i++) devices.push_back(atoi(argv[i]));
    if (devices.empty()) {
      const int count = myzkp_device_count();
      if (count < 1) { printf("FAIL no CUDA device\n"); return 1; }
      for (int d = 0; d < count; d++) devices.push_back(d);
      if (count == 1) devices.push_back(0);
    }
    const size_t n = ((size_t)1 << log2n) + 3;  // not a multiple of the rank count
    const Scalar alpha = scalar_u64(0x1234567890abcdefull), u = scalar_u64(0xfedcba9876543ull);
    Polynomial f;
    f.coef.resize(n);
    uint64_t seed = 42;
    for (size_t i = 0; i < n; i++) {
      uint64_t w[4] = {sm64(seed), sm64(seed), sm64(seed), sm64(seed) >> 3};  // < 2^253 < r
      memcpy(f.coef[i].data(), w, 32);
    }
    MultiGpuKZG mk(devices);
    mk.setup(n - 1, alpha);
    if (mk.size() != n || mk.world() != (int)devices.size()) { printf("FAIL setup\n"); return 1; }
    CommitmentKZG c = mk.commit_kzg(f);
    ProofKZG pr = mk.open_kzg(f, u);
    // twice more, timed (both exchange parities, steady state)
    auto t0 = std::chrono::steady_clock::now();
    CommitmentKZG c2 = mk.commit_kzg(f);
    ProofKZG pr2 = mk.open_kzg(f, u);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (!(c2 == c) || pr2.y != pr.y || !(pr2.w == pr.w)) { printf("FAIL second call differs\n"); return 1; }
    // shorter polynomials leave the upper ranks empty
    Polynomial g;
    g.coef.assign(f.coef.begin(), f.coef.begin() + 5);
    CommitmentKZG cg = mk.commit_kzg(g);
    ProofKZG pg = mk.open_kzg(g, u);
    // single-device path (already pinned against the oracle) on the same inputs
    PublicKeyKZG pk(devices[0]);
    setup_kzg(pk, n - 1, alpha);
    CommitmentKZG c1 = commit_kzg(f, pk);
    ProofKZG p1 = open_kzg(f, u, pk);
    if (!(c1 == c) || p1.y != pr.y || !(p1.w == pr.w)) { printf("FAIL multi-device result differs from one device\n"); return 1; }
    if (!(commit_kzg(g, pk) == cg)) { printf("FAIL short commit\n"); return 1; }
    ProofKZG pg1 = open_kzg(g, u, pk);
    if (pg1.y != pg.y || !(pg1.w == pg.w)) { printf("FAIL short open\n"); return 1; }
    bool threw = false;
    Polynomial too_long = f;
    too_long.coef.push_back(scalar_u64(1));
    try { mk.commit_kzg(too_long); } catch (const std::runtime_error&) { threw = true; }
    if (!threw) { printf("FAIL no error for deg > max_d\n"); return 1; }
    printf("world %d\nn %zu\nms_commit_plus_open %.3f\n", mk.world(), n, ms);
    printf("C.x %s\nC.y %s\ny %s\nW.x %s\nW.y %s\n", hex_be(c.xy.data()).c_str(), hex_be(c.xy.data() + 32).c_str(),
           hex_be(pr.y.data()).c_str(), hex_be(pr.w.xy.data()).c_str(), hex_be(pr.w.xy.data() + 32).c_str());
    printf("OK\n");
  } catch (const std::exception& e) {
    printf("FAIL %s\n", e.what());
    return 1;
  }
  return 0;
}
