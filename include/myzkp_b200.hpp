// Header-only C++ mirror of the reference's KZG surface over the C ABI (include/myzkp_b200.h):
// setup_kzg / commit_kzg / open_kzg / commit_gemini (myzkp/src/modules/algebra/kzg.rs:27-72,
// gemini.rs:112-114) with the same argument meaning and error behaviour (errors throw where the
// reference panics).  Scalars and coordinates are 32-byte little-endian canonical values.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "myzkp_b200.h"

namespace myzkp_b200 {

using Scalar = std::array<uint8_t, 32>;  // FqOrder value, canonical, LE
struct G1Point {                         // EllipticCurvePoint<Fq, BN128Curve> (curve.rs:17-22)
  std::array<uint8_t, 64> xy{};          // x || y; all-zero = point at infinity
  bool is_point_at_infinity() const {
    for (uint8_t b : xy) if (b) return false;
    return true;
  }
  bool operator==(const G1Point& o) const { return xy == o.xy; }
};
struct G2Point {                         // EllipticCurvePoint<Fq2, BN128Curve> (bn128.rs:49); all-zero = infinity
  std::array<uint8_t, 128> xy{};         // x.c0 | x.c1 | y.c0 | y.c1, 32 B little-endian canonical each
  bool is_point_at_infinity() const {
    for (uint8_t b : xy) if (b) return false;
    return true;
  }
  bool operator==(const G2Point& o) const { return xy == o.xy; }
};
struct Polynomial {                      // Polynomial<FqOrder> (polynomial.rs:69-74), low -> high degree
  std::vector<Scalar> coef;
};
struct ProofKZG {                        // kzg.rs:15-18
  Scalar y{};
  G1Point w;
};
using CommitmentKZG = G1Point;

class PublicKeyKZG {                     // kzg.rs:8-11; powers_1 resident on the GPU
 public:
  explicit PublicKeyKZG(int device = 0) {
    if (myzkp_ctx_create(&ctx_, device) != MYZKP_OK) throw std::runtime_error("myzkp_b200: no usable CUDA device (no CPU fallback)");
  }
  ~PublicKeyKZG() { myzkp_ctx_destroy(ctx_); }
  PublicKeyKZG(const PublicKeyKZG&) = delete;
  PublicKeyKZG& operator=(const PublicKeyKZG&) = delete;
  myzkp_ctx* ctx() const { return ctx_; }
  size_t size() const { return myzkp_srs_len(ctx_); }
  std::vector<G1Point> powers_1() const {
    std::vector<G1Point> out(size());
    check(myzkp_srs_read_g1(ctx_, 0, out.size(), out.empty() ? nullptr : out[0].xy.data()));
    return out;
  }
  void check(int code) const {
    if (code != MYZKP_OK) throw std::runtime_error(std::string("myzkp_b200: ") + myzkp_last_error(ctx_));
  }

 private:
  myzkp_ctx* ctx_ = nullptr;
};

// setup_kzg (kzg.rs:27-40) with alpha injected; max_d + 1 powers of the standard generator
inline void setup_kzg(PublicKeyKZG& pk, size_t max_d, const Scalar& alpha) {
  pk.check(myzkp_srs_generate_g1(pk.ctx(), alpha.data(), 0, max_d + 1));
}
// powers_2 of the public key: [alpha^i] g2 for i < n, g2 = BN128::generator_g2() when `base` is null
// (n = 2: setup_kzg, kzg.rs:37; n = max_d + 1: setup_kzg_with_full_g2, kzg.rs:47-52)
inline std::vector<G2Point> powers_2(const PublicKeyKZG& pk, const Scalar& alpha, size_t n, const G2Point* base = nullptr) {
  std::vector<G2Point> out(n);
  pk.check(myzkp_srs_generate_g2(pk.ctx(), alpha.data(), base ? base->xy.data() : nullptr, 0, n,
                                 out.empty() ? nullptr : out[0].xy.data()));
  return out;
}
// accumulate_curve_points over G2 (zksnark/utils.rs:83-93): sum_i assignment[i] * g_vec[i]
inline G2Point accumulate_curve_points(const std::vector<G2Point>& g_vec, const std::vector<Scalar>& assignment,
                                       const PublicKeyKZG& pk) {
  const size_t n = g_vec.size() < assignment.size() ? g_vec.size() : assignment.size();  // zip() stops at the shorter
  G2Point out;
  pk.check(myzkp_g2_msm(pk.ctx(), n ? assignment[0].data() : nullptr, n ? g_vec[0].xy.data() : nullptr, n, out.xy.data()));
  return out;
}
// prod_i e(g1[i], g2[i]) == 1 (optimal_ate_pairing, curve/bn128.rs:147-181; one final exponentiation)
inline bool pairing_product_is_one(const std::vector<G1Point>& g1, const std::vector<G2Point>& g2, const PublicKeyKZG& pk) {
  const size_t n = g1.size() < g2.size() ? g1.size() : g2.size();
  int ok = 0;
  pk.check(myzkp_pairing_product_is_one(pk.ctx(), n ? g1[0].xy.data() : nullptr, n ? g2[0].xy.data() : nullptr, n, &ok));
  return ok != 0;
}
// r - x for a canonical scalar x (0 stays 0): the shim's FqOrder negation, done on the wire bytes
inline Scalar scalar_neg(const Scalar& x) {
  static const Scalar r = {0x01, 0x00, 0x00, 0xf0, 0x93, 0xf5, 0xe1, 0x43, 0x91, 0x70, 0xb9, 0x79, 0x48, 0xe8, 0x33, 0x28,
                           0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45, 0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};
  bool zero = true;
  for (uint8_t b : x) zero = zero && b == 0;
  Scalar out{};
  if (zero) return out;
  int borrow = 0;
  for (int i = 0; i < 32; i++) {
    int d = (int)r[i] - (int)x[i] - borrow;
    borrow = d < 0;
    out[i] = (uint8_t)(d + (borrow << 8));
  }
  return out;
}
inline Scalar scalar_one() {
  Scalar s{};
  s[0] = 1;
  return s;
}
// -P through the device group law (a G1 MSM with the scalar r - 1)
inline G1Point g1_neg(const G1Point& p, const PublicKeyKZG& pk) {
  G1Point out;
  const Scalar m1 = scalar_neg(scalar_one());
  pk.check(myzkp_g1_msm(pk.ctx(), m1.data(), p.xy.data(), 1, out.xy.data()));
  return out;
}
// verify_degree_bound (kzg.rs:136-144): e(proof, g2) == e(c, [alpha^(max_d - d)]g2) as a product with -c
inline bool verify_degree_bound(const CommitmentKZG& c, const G1Point& proof, const PublicKeyKZG& pk, const G2Point& g2,
                                const G2Point& g2_pow_maxd_minus_d) {
  return pairing_product_is_one({proof, g1_neg(c, pk)}, {g2, g2_pow_maxd_minus_d}, pk);
}
// verify_kzg (kzg.rs:90-102): e(C, g2) == e(W, [alpha]g2 - [u]g2) * e(g1, g2)^y, evaluated as
// e(C, g2) * e(-W, [alpha - u]g2) * e([-y]g1, g2) == 1; p2 = [g2, [alpha]g2] (kzg.rs:37)
inline bool verify_kzg(const Scalar& u, const CommitmentKZG& c, const ProofKZG& proof, const PublicKeyKZG& pk,
                       const std::vector<G2Point>& p2) {
  const std::vector<G1Point> g1 = {pk.powers_1().at(0)};
  const G2Point g2_alpha_minus_u = accumulate_curve_points({p2.at(1), p2.at(0)}, {scalar_one(), scalar_neg(u)}, pk);
  G1Point g1_minus_y;
  const Scalar my = scalar_neg(proof.y);
  pk.check(myzkp_g1_msm(pk.ctx(), my.data(), g1[0].xy.data(), 1, g1_minus_y.xy.data()));
  return pairing_product_is_one({c, g1_neg(proof.w, pk), g1_minus_y}, {p2.at(0), g2_alpha_minus_u, p2.at(0)}, pk);
}
// batch_open_kzg (kzg.rs:74-88)
struct BatchProofKZG {
  std::vector<Scalar> ys;
  G1Point w;
};
inline BatchProofKZG batch_open_kzg(const Polynomial& f, const std::vector<Scalar>& us, const PublicKeyKZG& pk) {
  BatchProofKZG p;
  p.ys.resize(us.size());
  pk.check(myzkp_kzg_batch_open(pk.ctx(), f.coef.empty() ? nullptr : f.coef[0].data(), f.coef.size(),
                                us.empty() ? nullptr : us[0].data(), us.size(), us.empty() ? nullptr : p.ys[0].data(),
                                p.w.xy.data()));
  return p;
}
// prove_degree_bound (kzg.rs:121-134)
inline G1Point prove_degree_bound(const Polynomial& f, const PublicKeyKZG& pk, size_t d) {
  G1Point out;
  pk.check(myzkp_kzg_prove_degree_bound(pk.ctx(), f.coef.empty() ? nullptr : f.coef[0].data(), f.coef.size(), d, out.xy.data()));
  return out;
}
// commit_kzg (kzg.rs:57-59)
inline CommitmentKZG commit_kzg(const Polynomial& f, const PublicKeyKZG& pk) {
  G1Point c;
  pk.check(myzkp_kzg_commit(pk.ctx(), f.coef.empty() ? nullptr : f.coef[0].data(), f.coef.size(), c.xy.data()));
  return c;
}
// open_kzg (kzg.rs:61-72)
inline ProofKZG open_kzg(const Polynomial& f, const Scalar& u, const PublicKeyKZG& pk) {
  ProofKZG p;
  pk.check(myzkp_kzg_open(pk.ctx(), f.coef.empty() ? nullptr : f.coef[0].data(), f.coef.size(), u.data(), p.y.data(),
                          p.w.xy.data()));
  return p;
}
// commit_gemini (gemini.rs:112-114): one batched call
inline std::vector<CommitmentKZG> commit_gemini(const std::vector<Polynomial>& polys, const PublicKeyKZG& pk) {
  std::vector<const uint8_t*> ptrs;
  std::vector<size_t> lens;
  for (const auto& p : polys) {
    ptrs.push_back(p.coef.empty() ? nullptr : p.coef[0].data());
    lens.push_back(p.coef.size());
  }
  std::vector<CommitmentKZG> out(polys.size());
  if (!polys.empty()) pk.check(myzkp_kzg_commit_batch(pk.ctx(), ptrs.data(), lens.data(), polys.size(), out[0].xy.data()));
  return out;
}

// ---- range-sharded prover: one PublicKeyKZG per GPU, rank g holding powers_1[first, first + count) ----
inline void setup_kzg_range(PublicKeyKZG& pk, size_t first, size_t count, const Scalar& alpha) {
  pk.check(myzkp_srs_generate_g1(pk.ctx(), alpha.data(), first, count));
}
// all ranks live in this process: map every rank's exchange buffer into every other (csrc/peer.cu)
inline void attach_peers(const std::vector<PublicKeyKZG*>& ranks) {
  std::vector<myzkp_ctx*> ctxs;
  for (auto* r : ranks) {
    // size the scratch now: ranks sharing a device must not allocate once a peer may be spinning in an exchange
    r->check(myzkp_ctx_reserve(r->ctx(), myzkp_srs_len(r->ctx())));
    r->check(myzkp_peer_export(r->ctx(), nullptr));
    ctxs.push_back(r->ctx());
  }
  for (size_t g = 0; g < ranks.size(); g++)
    ranks[g]->check(myzkp_peer_attach_local(ranks[g]->ctx(), (int)g, (int)ranks.size(), ctxs.data()));
}
// commit_kzg of the whole polynomial from this rank's coefficient slice; every rank returns the same point.
// Blocks until the peers have called it too: drive each rank from its own host thread.
inline CommitmentKZG commit_kzg_sharded(const Polynomial& local_slice, const PublicKeyKZG& pk) {
  G1Point c;
  pk.check(myzkp_kzg_commit_sharded(pk.ctx(), local_slice.coef.empty() ? nullptr : local_slice.coef[0].data(),
                                    local_slice.coef.size(), c.xy.data()));
  return c;
}

inline ProofKZG open_kzg_sharded(const Polynomial& local_slice, const Scalar& u, const PublicKeyKZG& pk) {
  ProofKZG pr;
  pk.check(myzkp_kzg_open_sharded(pk.ctx(), local_slice.coef.empty() ? nullptr : local_slice.coef[0].data(),
                                  local_slice.coef.size(), u.data(), pr.y.data(), pr.w.xy.data()));
  return pr;
}

// ---- one host thread, several GPUs: the multi-device context (csrc/multi.cu) ----
class MultiGpuKZG {
 public:
  explicit MultiGpuKZG(const std::vector<int>& devices) {
    if (myzkp_mctx_create(&m_, devices.data(), (int)devices.size()) != MYZKP_OK)
      throw std::runtime_error("myzkp_b200: cannot create contexts on the listed devices (there is no CPU fallback)");
  }
  ~MultiGpuKZG() { myzkp_mctx_destroy(m_); }
  MultiGpuKZG(const MultiGpuKZG&) = delete;
  MultiGpuKZG& operator=(const MultiGpuKZG&) = delete;
  void check(int code) const {
    if (code != MYZKP_OK) throw std::runtime_error(std::string("myzkp_b200: ") + myzkp_mctx_last_error(m_));
  }
  void setup(size_t max_d, const Scalar& alpha) { check(myzkp_mctx_srs_generate_g1(m_, alpha.data(), max_d + 1)); }
  size_t size() const { return myzkp_mctx_srs_len(m_); }
  int world() const { return myzkp_mctx_world(m_); }
  CommitmentKZG commit_kzg(const Polynomial& f) const {
    G1Point c;
    check(myzkp_mctx_kzg_commit(m_, f.coef.empty() ? nullptr : f.coef[0].data(), f.coef.size(), c.xy.data()));
    return c;
  }
  ProofKZG open_kzg(const Polynomial& f, const Scalar& u) const {
    ProofKZG pr;
    check(myzkp_mctx_kzg_open(m_, f.coef.empty() ? nullptr : f.coef[0].data(), f.coef.size(), u.data(), pr.y.data(),
                              pr.w.xy.data()));
    return pr;
  }

 private:
  myzkp_mctx* m_ = nullptr;
};

}  // namespace myzkp_b200
