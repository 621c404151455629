/* C restatement of the reference's KZG hot path (BN128), for sizes and timings
 * the Python oracle (oracle/myzkp_oracle.py) is too slow for.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  Never by the product.
 *
 * It keeps the reference's algorithmic structure (citations relative to
 * /root/reference/myzkp/src/modules/algebra/):
 *   - field ops = big-integer op followed by `% modulus`        field.rs:157-183
 *   - inverse   = full extended Euclid with quotient/remainder   field.rs:210-237, utils.rs:52-81
 *   - affine add / double, one inversion each, all special cases curve/curve.rs:56-161
 *   - [k]P      = LSB-first double-and-add                       curve/curve.rs:163-191
 *   - commit    = naive MSM, sum_i [c_i]P_i in index order       polynomial.rs:156-165, kzg.rs:57-59
 *   - eval      = running-power sum                              polynomial.rs:120-128
 *   - quotient  = schoolbook long division that re-copies the
 *                 remainder every iteration (the O(d^2) step)    polynomial.rs:371-405, 85-91
 *   - setup     = [alpha^i]G for i = 0..max_d                    kzg.rs:27-40
 *   - fold      = f[2k] + rho f[2k+1]                            gemini.rs:71-98
 * The reference is single-threaded; `threads` > 1 splits the naive MSM / setup
 * by index ranges (an all-cores variant of the same algorithm, labelled as such
 * wherever it is reported).
 *
 * Parity pinning: checked against the Python oracle and the reference's KATs in
 * tests/test_oracle_c.py (the reference itself cannot be built here: no Rust).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } u256;

static const u256 P_MOD = {{0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}};
static const u256 R_MOD = {{0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}};

/* ---- 256-bit helpers ------------------------------------------------------ */
static int u256_is_zero(const u256* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static int u256_cmp(const u256* a, const u256* b) {
  for (int i = 3; i >= 0; i--) {
    if (a->l[i] < b->l[i]) return -1;
    if (a->l[i] > b->l[i]) return 1;
  }
  return 0;
}
static uint64_t u256_add(u256* r, const u256* a, const u256* b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a->l[i] + b->l[i]; r->l[i] = (uint64_t)c; c >>= 64; }
  return (uint64_t)c;
}
static uint64_t u256_sub(u256* r, const u256* a, const u256* b) {
  uint64_t borrow = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a->l[i] - b->l[i] - borrow;
    r->l[i] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
  return borrow;
}
static int u256_bitlen(const u256* a) {
  for (int i = 3; i >= 0; i--)
    if (a->l[i]) return 64 * i + 64 - __builtin_clzll(a->l[i]);
  return 0;
}
static void u256_shl(u256* r, const u256* a, int s) { /* 0 <= s < 256 */
  int w = s >> 6, b = s & 63;
  u256 t = {{0, 0, 0, 0}};
  for (int i = 3; i >= w; i--) {
    uint64_t v = a->l[i - w] << b;
    if (b && i - w - 1 >= 0) v |= a->l[i - w - 1] >> (64 - b);
    t.l[i] = v;
  }
  *r = t;
}
static void u256_shr1(u256* a) {
  for (int i = 0; i < 3; i++) a->l[i] = (a->l[i] >> 1) | (a->l[i + 1] << 63);
  a->l[3] >>= 1;
}
static void u256_from_le(u256* r, const uint8_t* b) { memcpy(r->l, b, 32); }
static void u256_to_le(uint8_t* b, const u256* a) { memcpy(b, a->l, 32); }

/* q = a / b, rem = a % b for b != 0 (binary long division over the bit-length gap) */
static void u256_divrem(u256* q, u256* rem, const u256* a, const u256* b) {
  u256 r = *a, quo = {{0, 0, 0, 0}};
  int gap = u256_bitlen(a) - u256_bitlen(b);
  if (gap >= 0) {
    u256 d;
    u256_shl(&d, b, gap);
    for (int i = gap; i >= 0; i--) {
      if (u256_cmp(&r, &d) >= 0) {
        u256_sub(&r, &r, &d);
        quo.l[i >> 6] |= 1ULL << (i & 63);
      }
      u256_shr1(&d);
    }
  }
  *q = quo;
  *rem = r;
}
/* low 256 bits of a*b (the callers' results are bounded by the modulus) */
static void u256_mul_lo(u256* r, const u256* a, const u256* b) {
  uint64_t t[4] = {0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; i + j < 4; j++) {
      c += (u128)a->l[i] * b->l[j] + t[i + j];
      t[i + j] = (uint64_t)c;
      c >>= 64;
    }
  }
  memcpy(r->l, t, 32);
}

/* ---- field ops: big-int op, then `% modulus` (field.rs:157-183) ------------ */
/* r = (a * b) % m via the full 512-bit product and a word-wise long division
 * (Knuth D, base 2^64; m has 2 leading zero bits). */
static void fe_mul(u256* r, const u256* a, const u256* b, const u256* m) {
  uint64_t prod[9];
  memset(prod, 0, sizeof prod);
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a->l[i] * b->l[j] + prod[i + j];
      prod[i + j] = (uint64_t)c;
      c >>= 64;
    }
    prod[i + 4] = (uint64_t)c;
  }
  const int s = __builtin_clzll(m->l[3]);
  uint64_t v[4], u[9];
  for (int i = 3; i > 0; i--) v[i] = (m->l[i] << s) | (s ? m->l[i - 1] >> (64 - s) : 0);
  v[0] = m->l[0] << s;
  u[8] = s ? prod[7] >> (64 - s) : 0;
  for (int i = 7; i > 0; i--) u[i] = (prod[i] << s) | (s ? prod[i - 1] >> (64 - s) : 0);
  u[0] = prod[0] << s;
  for (int j = 4; j >= 0; j--) {
    u128 num = ((u128)u[j + 4] << 64) | u[j + 3];
    u128 qhat = num / v[3], rhat = num % v[3];
    while ((qhat >> 64) || (uint64_t)qhat * (u128)v[2] > ((rhat << 64) | u[j + 2])) {
      qhat--;
      rhat += v[3];
      if (rhat >> 64) break;
    }
    /* multiply and subtract */
    u128 carry = 0;
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
      u128 p = (u128)(uint64_t)qhat * v[i] + carry;
      carry = p >> 64;
      u128 d = (u128)u[i + j] - (uint64_t)p - borrow;
      u[i + j] = (uint64_t)d;
      borrow = (uint64_t)(d >> 64) & 1;
    }
    u128 d = (u128)u[j + 4] - (uint64_t)carry - borrow;
    u[j + 4] = (uint64_t)d;
    if ((uint64_t)(d >> 64) & 1) { /* qhat was one too large: add back */
      u128 c = 0;
      for (int i = 0; i < 4; i++) {
        c += (u128)u[i + j] + v[i];
        u[i + j] = (uint64_t)c;
        c >>= 64;
      }
      u[j + 4] += (uint64_t)c;
    }
  }
  for (int i = 0; i < 3; i++) r->l[i] = (u[i] >> s) | (s ? u[i + 1] << (64 - s) : 0);
  r->l[3] = u[3] >> s;
}
static void fe_add(u256* r, const u256* a, const u256* b, const u256* m) {
  u256_add(r, a, b); /* both < m < 2^254: no carry */
  if (u256_cmp(r, m) >= 0) u256_sub(r, r, m);
}
static void fe_sub(u256* r, const u256* a, const u256* b, const u256* m) {
  if (u256_sub(r, a, b)) u256_add(r, r, m);
}
/* inverse by extended Euclid tracking only the coefficient of `a`
 * (utils.rs:52-81 with r0 = m, r1 = a; field.rs:210-237); inverse(0) = 0 */
static void fe_inv(u256* r, const u256* a, const u256* m) {
  u256 r0 = *m, r1 = *a;
  u256 t0 = {{0, 0, 0, 0}}, t1 = {{1, 0, 0, 0}};
  int s0 = 0, s1 = 0; /* signs: 1 = negative */
  while (!u256_is_zero(&r1)) {
    u256 q, rem, qt, nt;
    int ns;
    u256_divrem(&q, &rem, &r0, &r1);
    r0 = r1;
    r1 = rem;
    u256_mul_lo(&qt, &q, &t1); /* |q * t1| */
    /* new_t = t0 - q*t1 */
    if (s0 != s1) { /* opposite signs: magnitudes add, sign of t0 */
      u256_add(&nt, &t0, &qt);
      ns = s0;
    } else if (u256_cmp(&t0, &qt) >= 0) {
      u256_sub(&nt, &t0, &qt);
      ns = s0;
    } else {
      u256_sub(&nt, &qt, &t0);
      ns = !s0;
    }
    t0 = t1; s0 = s1;
    t1 = nt; s1 = ns;
  }
  /* t0 %= m; if negative add m (field.rs:226-229) */
  u256 q, rem;
  u256_divrem(&q, &rem, &t0, m);
  if (s0 && !u256_is_zero(&rem)) u256_sub(&rem, m, &rem);
  *r = rem;
}

/* ---- affine G1 (curve/curve.rs:17-191); inf flag = (None, None) ------------- */
typedef struct { u256 x, y; int inf; } pt_t;

static void pt_double(pt_t* p) { /* curve.rs:72-101 with line_slope :56-70 (a = 0) */
  if (p->inf) return;
  const u256* m = &P_MOD;
  u256 three = {{3, 0, 0, 0}}, two = {{2, 0, 0, 0}}, num, den, s, nx, ny, t;
  fe_mul(&num, &p->x, &p->x, m);
  fe_mul(&num, &num, &three, m);
  fe_mul(&den, &p->y, &two, m);
  fe_inv(&den, &den, m);
  fe_mul(&s, &num, &den, m);
  fe_mul(&nx, &s, &s, m);
  fe_sub(&nx, &nx, &p->x, m);
  fe_sub(&nx, &nx, &p->x, m);
  fe_mul(&t, &s, &nx, m); /* new_y = -s*new_x + s*x - y */
  fe_mul(&ny, &s, &p->x, m);
  fe_sub(&ny, &ny, &t, m);
  fe_sub(&ny, &ny, &p->y, m);
  p->x = nx;
  p->y = ny;
}
static void pt_add_assign(pt_t* a, const pt_t* b) { /* curve.rs:130-161 */
  if (a->inf) { *a = *b; return; }
  if (b->inf) return;
  const u256* m = &P_MOD;
  if (u256_cmp(&a->x, &b->x) == 0) {
    if (u256_cmp(&a->y, &b->y) == 0) { pt_double(a); return; }
    a->inf = 1;
    return;
  }
  u256 num, den, s, nx, ny, t;
  fe_sub(&num, &b->y, &a->y, m);
  fe_sub(&den, &b->x, &a->x, m);
  fe_inv(&den, &den, m);
  fe_mul(&s, &num, &den, m);
  fe_mul(&nx, &s, &s, m);
  fe_sub(&nx, &nx, &a->x, m);
  fe_sub(&nx, &nx, &b->x, m);
  fe_mul(&t, &s, &nx, m); /* new_y = -s*new_x + (s*x1 - y1) */
  fe_mul(&ny, &s, &a->x, m);
  fe_sub(&ny, &ny, &a->y, m);
  fe_sub(&ny, &ny, &t, m);
  a->x = nx;
  a->y = ny;
}
static void pt_mul(pt_t* out, const pt_t* p, const u256* k) { /* curve.rs:168-191 */
  pt_t result, cur = *p;
  memset(&result, 0, sizeof result);
  result.inf = 1;
  u256 bits = *k;
  while (!u256_is_zero(&bits)) {
    if (bits.l[0] & 1) pt_add_assign(&result, &cur);
    pt_double(&cur);
    u256_shr1(&bits);
  }
  *out = result;
}
static void pt_from_bytes(pt_t* p, const uint8_t* b) {
  u256_from_le(&p->x, b);
  u256_from_le(&p->y, b + 32);
  p->inf = u256_is_zero(&p->x) && u256_is_zero(&p->y);
}
static void pt_to_bytes(uint8_t* b, const pt_t* p) {
  if (p->inf) { memset(b, 0, 64); return; }
  u256_to_le(b, &p->x);
  u256_to_le(b + 32, &p->y);
}

/* ---- threaded naive MSM / setup ------------------------------------------- */
typedef struct {
  const uint8_t* coefs; const uint8_t* points; size_t lo, hi; pt_t acc;      /* msm */
  u256 alpha; uint8_t* out;                                                  /* setup */
} job_t;

static void* msm_worker(void* arg) { /* polynomial.rs:156-165 on an index range */
  job_t* j = (job_t*)arg;
  memset(&j->acc, 0, sizeof j->acc);
  j->acc.inf = 1;
  for (size_t i = j->lo; i < j->hi; i++) {
    pt_t p, t;
    u256 k;
    pt_from_bytes(&p, j->points + 64 * i);
    u256_from_le(&k, j->coefs + 32 * i);
    pt_mul(&t, &p, &k);
    pt_add_assign(&j->acc, &t);
  }
  return NULL;
}
static void* setup_worker(void* arg) { /* kzg.rs:31-35 on an index range */
  job_t* j = (job_t*)arg;
  pt_t g;
  memset(&g, 0, sizeof g);
  g.x.l[0] = 1; g.y.l[0] = 2; /* bn128.rs:185-188 */
  u256 ap = {{1, 0, 0, 0}}, e;
  /* alpha^lo by square-and-multiply, then one multiply per step as the reference */
  u256 base = j->alpha;
  for (size_t b = j->lo; b; b >>= 1) {
    if (b & 1) fe_mul(&ap, &ap, &base, &R_MOD);
    fe_mul(&base, &base, &base, &R_MOD);
  }
  for (size_t i = j->lo; i < j->hi; i++) {
    pt_t t;
    e = ap;
    pt_mul(&t, &g, &e);
    pt_to_bytes(j->out + 64 * i, &t);
    fe_mul(&ap, &ap, &j->alpha, &R_MOD);
  }
  return NULL;
}
static void run_jobs(void* (*fn)(void*), job_t* jobs, int threads) {
  if (threads <= 1) { fn(&jobs[0]); return; }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, fn, &jobs[t]);
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  free(th);
}
static void naive_msm(const uint8_t* coefs, const uint8_t* points, size_t n, int threads, pt_t* out) {
  if (threads < 1) threads = 1;
  if ((size_t)threads > n && n) threads = (int)n;
  if (n == 0) threads = 1;
  job_t* jobs = (job_t*)calloc(threads, sizeof(job_t));
  for (int t = 0; t < threads; t++) {
    jobs[t].coefs = coefs; jobs[t].points = points;
    jobs[t].lo = n * t / threads; jobs[t].hi = n * (t + 1) / threads;
  }
  run_jobs(msm_worker, jobs, threads);
  *out = jobs[0].acc;
  for (int t = 1; t < threads; t++) pt_add_assign(out, &jobs[t].acc);
  free(jobs);
}

/* ---- exported API (all 32/64-byte little-endian canonical, as the C ABI) ---- */
int oracle_fe_op(int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  const u256* m = field == 0 ? &P_MOD : &R_MOD;
  u256 x, y, r;
  u256_from_le(&x, a);
  if (b) u256_from_le(&y, b); else memset(&y, 0, sizeof y);
  switch (op) {
    case 0: fe_add(&r, &x, &y, m); break;
    case 1: fe_sub(&r, &x, &y, m); break;
    case 2: fe_mul(&r, &x, &y, m); break;
    case 3: fe_inv(&r, &x, m); break;
    default: return -1;
  }
  u256_to_le(out, &r);
  return 0;
}
int oracle_g1_add(const uint8_t a[64], const uint8_t b[64], uint8_t out[64]) {
  pt_t p, q;
  pt_from_bytes(&p, a);
  pt_from_bytes(&q, b);
  pt_add_assign(&p, &q);
  pt_to_bytes(out, &p);
  return 0;
}
int oracle_g1_mul(const uint8_t pt[64], const uint8_t k[32], uint8_t out[64]) {
  pt_t p, r;
  u256 s;
  pt_from_bytes(&p, pt);
  u256_from_le(&s, k);
  pt_mul(&r, &p, &s);
  pt_to_bytes(out, &r);
  return 0;
}
/* setup_kzg (kzg.rs:27-40): n = max_d + 1 points [alpha^i]G */
int oracle_setup_kzg(const uint8_t alpha[32], size_t n, uint8_t* out_points, int threads) {
  if (threads < 1) threads = 1;
  if ((size_t)threads > n && n) threads = (int)n;
  if (n == 0) return 0;
  job_t* jobs = (job_t*)calloc(threads, sizeof(job_t));
  for (int t = 0; t < threads; t++) {
    u256_from_le(&jobs[t].alpha, alpha);
    jobs[t].out = out_points;
    jobs[t].lo = n * t / threads; jobs[t].hi = n * (t + 1) / threads;
  }
  run_jobs(setup_worker, jobs, threads);
  free(jobs);
  return 0;
}
/* commit_kzg (kzg.rs:57-59) */
int oracle_commit_kzg(const uint8_t* coefs, const uint8_t* points, size_t n, uint8_t out[64], int threads) {
  pt_t acc;
  naive_msm(coefs, points, n, threads, &acc);
  pt_to_bytes(out, &acc);
  return 0;
}
/* Polynomial::eval (polynomial.rs:120-128) */
int oracle_fr_eval(const uint8_t* coefs, size_t n, const uint8_t u[32], uint8_t out_y[32]) {
  u256 result = {{0, 0, 0, 0}}, tp = {{1, 0, 0, 0}}, point, c, t;
  u256_from_le(&point, u);
  for (size_t i = 0; i < n; i++) {
    u256_from_le(&c, coefs + 32 * i);
    fe_mul(&t, &tp, &c, &R_MOD);
    fe_add(&result, &result, &t, &R_MOD);
    fe_mul(&tp, &tp, &point, &R_MOD);
  }
  u256_to_le(out_y, &result);
  return 0;
}
static size_t trim_len(const u256* v, size_t n) { /* polynomial.rs:85-91 */
  while (n && u256_is_zero(&v[n - 1])) n--;
  return n;
}
/* (f - y) / (x - u) by div_rem_ref (polynomial.rs:371-405), including the per-iteration
 * trim copy (:394 -> :90); writes n-1 quotient coefficients (zero padded). */
int oracle_quotient(const uint8_t* coefs, size_t n, const uint8_t u[32], uint8_t out_y[32], uint8_t* out_q) {
  u256 pu, y;
  u256_from_le(&pu, u);
  oracle_fr_eval(coefs, n, u, out_y);
  u256_from_le(&y, out_y);
  if (n >= 2) memset(out_q, 0, 32 * (n - 1));
  if (n == 0) return 0;
  u256* rem = (u256*)malloc(sizeof(u256) * n);
  for (size_t i = 0; i < n; i++) u256_from_le(&rem[i], coefs + 32 * i);
  fe_sub(&rem[0], &rem[0], &y, &R_MOD); /* f - y_poly (polynomial.rs:517-523) */
  size_t len = trim_len(rem, n);
  u256 div0, div1 = {{1, 0, 0, 0}}, zero = {{0, 0, 0, 0}}, lead_inv;
  fe_sub(&div0, &zero, &pu, &R_MOD); /* from_monomials([u]) = [-u, 1] (polynomial.rs:202-212) */
  fe_inv(&lead_inv, &div1, &R_MOD);
  while (len >= 2) {
    u256 lead, t;
    fe_mul(&lead, &rem[len - 1], &lead_inv, &R_MOD);
    size_t deg_diff = len - 2;
    u256_to_le(out_q + 32 * deg_diff, &lead);
    fe_mul(&t, &lead, &div0, &R_MOD);
    fe_sub(&rem[deg_diff], &rem[deg_diff], &t, &R_MOD);
    fe_mul(&t, &lead, &div1, &R_MOD);
    fe_sub(&rem[deg_diff + 1], &rem[deg_diff + 1], &t, &R_MOD);
    size_t nl = trim_len(rem, len);
    u256* copy = (u256*)malloc(sizeof(u256) * (nl ? nl : 1)); /* coef[..end].to_vec() */
    memcpy(copy, rem, sizeof(u256) * nl);
    free(rem);
    rem = copy;
    len = nl;
  }
  free(rem);
  return 0;
}
/* open_kzg (kzg.rs:61-72) */
int oracle_open_kzg(const uint8_t* coefs, size_t n, const uint8_t u[32], const uint8_t* points, uint8_t out_y[32],
                    uint8_t out_w[64], int threads) {
  uint8_t* q = (uint8_t*)malloc(n > 1 ? 32 * (n - 1) : 32);
  oracle_quotient(coefs, n, u, out_y, q);
  pt_t acc;
  naive_msm(q, points, n > 1 ? n - 1 : 0, threads, &acc);
  pt_to_bytes(out_w, &acc);
  free(q);
  return 0;
}
/* one level of split_and_fold (gemini.rs:71-98): out[k] = f[2k] + rho f[2k+1] */
int oracle_fold(const uint8_t* coefs, size_t n_out, const uint8_t rho[32], uint8_t* out) {
  u256 r, e, o, t;
  u256_from_le(&r, rho);
  for (size_t k = 0; k < n_out; k++) {
    u256_from_le(&e, coefs + 64 * k);
    u256_from_le(&o, coefs + 64 * k + 32);
    fe_mul(&t, &o, &r, &R_MOD);
    fe_add(&e, &e, &t, &R_MOD);
    u256_to_le(out + 32 * k, &e);
  }
  return 0;
}
