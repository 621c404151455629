// C++ host using the header-only mirror of the reference surface (include/myzkp_b200.hpp) over the C ABI.
// Reproduces the reference's test_kzg polynomial (kzg.rs:152-175: (x+1)(x+2)(x+3), z = 5) with the fixed
// alpha = 123456789 and compares with the anchor values pinned in SURVEY.md appendix B / tests/test_oracle_kats.py.
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>

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

int main() {
  try {
    PublicKeyKZG pk(0);
    setup_kzg(pk, 3, scalar_u64(123456789));
    if (pk.size() != 4) { printf("FAIL srs size\n"); return 1; }
    Polynomial f;
    for (uint64_t c : {6, 11, 6, 1}) f.coef.push_back(scalar_u64(c));
    CommitmentKZG c = commit_kzg(f, pk);
    ProofKZG pr = open_kzg(f, scalar_u64(5), pk);
    // printed big-endian so the Python side can compare with the decimal anchors
    printf("C.x %s\nC.y %s\ny %s\nW.x %s\nW.y %s\n", hex_be(c.xy.data()).c_str(), hex_be(c.xy.data() + 32).c_str(),
           hex_be(pr.y.data()).c_str(), hex_be(pr.w.xy.data()).c_str(), hex_be(pr.w.xy.data() + 32).c_str());
    // empty polynomial commits to infinity; a too-long polynomial throws where the reference panics
    Polynomial empty;
    if (!commit_kzg(empty, pk).is_point_at_infinity()) { printf("FAIL empty\n"); return 1; }
    Polynomial too_long = f;
    too_long.coef.push_back(scalar_u64(1));
    bool threw = false;
    try { commit_kzg(too_long, pk); } catch (const std::runtime_error&) { threw = true; }
    if (!threw) { printf("FAIL no error for deg > max_d\n"); return 1; }
    // commit_gemini: one batched call, same points as one commit per polynomial
    Polynomial g;
    for (uint64_t v : {7, 0, 5}) g.coef.push_back(scalar_u64(v));
    auto batch = commit_gemini({f, g, empty}, pk);
    if (batch.size() != 3 || batch[0].xy != c.xy || batch[1].xy != commit_kzg(g, pk).xy || !batch[2].is_point_at_infinity()) {
      printf("FAIL commit_gemini batch\n");
      return 1;
    }
    // G2 half of the key: [g2, [alpha]g2]; alpha^0 = 1 gives the generator back, 2 * g2 is a public vector
    {
      auto p2 = powers_2(pk, scalar_u64(2), 2);
      printf("G2.2x0 %s\n", hex_be(p2[1].xy.data()).c_str());
      if (p2.size() != 2 || p2[0].is_point_at_infinity() || p2[1] == p2[0]) { printf("FAIL powers_2\n"); return 1; }
      // pairing product: e(g, g2) e(-g, g2) == 1, and != 1 against 2 g2 (the shape of verify_degree_bound, kzg.rs:136-144)
      G1Point g = pk.powers_1()[0];
      if (!verify_degree_bound(g, g, pk, p2[0], p2[0]) || verify_degree_bound(g, g, pk, p2[0], p2[1])) {
        printf("FAIL pairing product\n");
        return 1;
      }
    }
    // range-sharded commit inside one process: two ranks (own threads), exchange over peer memory
    {
      PublicKeyKZG r0(0), r1(0);
      setup_kzg_range(r0, 0, 2, scalar_u64(123456789));
      setup_kzg_range(r1, 2, 2, scalar_u64(123456789));
      attach_peers({&r0, &r1});
      Polynomial lo, hi;
      lo.coef = {f.coef[0], f.coef[1]};
      hi.coef = {f.coef[2], f.coef[3]};
      CommitmentKZG c0, c1;
      std::string e0, e1;  // an exception must not escape a thread (std::terminate would lose the buffered output)
      std::thread t0([&] { try { c0 = commit_kzg_sharded(lo, r0); } catch (const std::exception& e) { e0 = e.what(); } });
      std::thread t1([&] { try { c1 = commit_kzg_sharded(hi, r1); } catch (const std::exception& e) { e1 = e.what(); } });
      t0.join();
      t1.join();
      if (!e0.empty() || !e1.empty()) { printf("FAIL sharded commit threw: [%s] [%s]\n", e0.c_str(), e1.c_str()); return 1; }
      if (c0.xy != c.xy || c1.xy != c.xy) { printf("FAIL sharded commit\n"); return 1; }
      ProofKZG p0, p1;
      std::thread t2([&] { try { p0 = open_kzg_sharded(lo, scalar_u64(5), r0); } catch (const std::exception& e) { e0 = e.what(); } });
      std::thread t3([&] { try { p1 = open_kzg_sharded(hi, scalar_u64(5), r1); } catch (const std::exception& e) { e1 = e.what(); } });
      t2.join();
      t3.join();
      if (!e0.empty() || !e1.empty()) { printf("FAIL sharded open threw: [%s] [%s]\n", e0.c_str(), e1.c_str()); return 1; }
      if (p0.y != pr.y || p1.y != pr.y || !(p0.w == pr.w) || !(p1.w == pr.w)) { printf("FAIL sharded open\n"); return 1; }
    }
    printf("OK\n");
  } catch (const std::exception& e) {
    printf("FAIL %s\n", e.what());
    return 1;
  }
  return 0;
}
