// One host process driving several GPUs through the C ABI: the multi-device context of SURVEY 8(b)/(e)
// (`myzkp_ctx_create(out, device_ids, n_dev)` there; here a type of its own so the single-device ABI stays as it is).
//
// A myzkp_mctx owns one myzkp_ctx per device plus one persistent host thread per device.  The SRS is
// range-sharded (rank g holds powers_1[lo_g, hi_g), contiguous ceil-split), a commit / open hands every rank
// its coefficient slice, each rank runs the single-GPU pipeline on its own stream and the ranks finish in the
// peer-memory exchange kernel (peer.cu) - no collective library, no torch, nothing but this library on the
// path.  This is what a Rust host (the reference is a single-process Rust program: kzg.rs:57-72) would drive.
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "ctx.cuh"

namespace {

struct Worker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> task;
  bool has_task = false, done = false, stop = false;
  int rc = 0;

  void loop() {
    std::unique_lock<std::mutex> lk(mu);
    while (true) {
      cv.wait(lk, [&] { return has_task || stop; });
      if (stop) return;
      std::function<int()> t = std::move(task);
      has_task = false;
      lk.unlock();
      int r = t();
      lk.lock();
      rc = r;
      done = true;
      cv.notify_all();
    }
  }
  void submit(std::function<int()> t) {
    std::lock_guard<std::mutex> lk(mu);
    task = std::move(t);
    has_task = true;
    done = false;
    cv.notify_all();
  }
  int wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return done; });
    return rc;
  }
};

}  // namespace

struct myzkp_mctx {
  std::vector<myzkp_ctx*> ranks;
  std::vector<Worker*> workers;
  std::vector<size_t> lo, hi;  // SRS range of each rank
  size_t srs_n = 0;
  std::string err;
  bool attached = false;

  // run f(g) on every rank's thread; the first failing rank's code and message win
  int each(const std::function<int(int)>& f) {
    const int G = (int)ranks.size();
    for (int g = 0; g < G; g++) workers[g]->submit([&f, g] { return f(g); });
    int rc = MYZKP_OK;
    for (int g = 0; g < G; g++) {
      int r = workers[g]->wait();
      if (r != MYZKP_OK && rc == MYZKP_OK) {
        rc = r;
        err = "rank " + std::to_string(g) + ": " + myzkp_last_error(ranks[g]);
      }
    }
    return rc;
  }
  void split(size_t n) {
    const size_t G = ranks.size();
    const size_t per = n ? (n + G - 1) / G : 0;
    lo.assign(G, 0);
    hi.assign(G, 0);
    for (size_t g = 0; g < G; g++) {
      lo[g] = g * per < n ? g * per : n;
      hi[g] = lo[g] + per < n ? lo[g] + per : n;
    }
    srs_n = n;
  }
  // after the SRS is in place: size every rank's scratch (so same-device ranks never allocate while a peer spins)
  // and map the exchange buffers
  int finish_srs() {
    int rc = each([this](int g) { return myzkp_ctx_reserve(ranks[g], hi[g] - lo[g]); });
    if (rc != MYZKP_OK) return rc;
    if (!attached) {
      for (size_t g = 0; g < ranks.size(); g++) {
        int r = myzkp_peer_export(ranks[g], nullptr);
        if (r != MYZKP_OK) { err = myzkp_last_error(ranks[g]); return r; }
      }
      for (size_t g = 0; g < ranks.size(); g++) {
        int r = myzkp_peer_attach_local(ranks[g], (int)g, (int)ranks.size(), ranks.data());
        if (r != MYZKP_OK) { err = myzkp_last_error(ranks[g]); return r; }
      }
      attached = true;
    }
    return MYZKP_OK;
  }
  // rank g's slice of an n-coefficient polynomial
  void slice(size_t n, int g, size_t* off, size_t* len) const {
    *off = lo[g] < n ? lo[g] : n;
    const size_t e = hi[g] < n ? hi[g] : n;
    *len = e - *off;
  }
};

extern "C" {

int myzkp_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return count;
}

int myzkp_mctx_create(myzkp_mctx** out, const int* device_ids, int n_dev) {
  if (!out || !device_ids || n_dev < 1 || n_dev > myzkp_ctx::kMaxPeers) return MYZKP_ERR_INVALID_ARG;
  *out = nullptr;
  myzkp_mctx* m = new myzkp_mctx();
  for (int g = 0; g < n_dev; g++) {
    myzkp_ctx* c = nullptr;
    int rc = myzkp_ctx_create(&c, device_ids[g]);
    if (rc != MYZKP_OK) {
      for (myzkp_ctx* x : m->ranks) myzkp_ctx_destroy(x);
      delete m;
      return rc;
    }
    m->ranks.push_back(c);
  }
  for (int g = 0; g < n_dev; g++) {
    Worker* w = new Worker();
    w->th = std::thread([w] { w->loop(); });
    m->workers.push_back(w);
  }
  m->split(0);
  *out = m;
  return MYZKP_OK;
}

int myzkp_mctx_destroy(myzkp_mctx* m) {
  if (!m) return MYZKP_OK;
  for (Worker* w : m->workers) {
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->stop = true;
      w->cv.notify_all();
    }
    w->th.join();
    delete w;
  }
  // detach every rank before any buffer goes away
  for (myzkp_ctx* c : m->ranks) myzkp_peer_detach(c);
  for (myzkp_ctx* c : m->ranks) myzkp_ctx_destroy(c);
  delete m;
  return MYZKP_OK;
}

const char* myzkp_mctx_last_error(const myzkp_mctx* m) { return m ? m->err.c_str() : "null mctx"; }
int myzkp_mctx_world(const myzkp_mctx* m) { return m ? (int)m->ranks.size() : 0; }
myzkp_ctx* myzkp_mctx_rank(myzkp_mctx* m, int g) { return (m && g >= 0 && g < (int)m->ranks.size()) ? m->ranks[g] : nullptr; }
size_t myzkp_mctx_srs_len(const myzkp_mctx* m) { return m ? m->srs_n : 0; }

int myzkp_mctx_srs_generate_g1(myzkp_mctx* m, const uint8_t alpha_le[32], size_t n) {
  if (!m || !alpha_le) return MYZKP_ERR_INVALID_ARG;
  m->split(n);
  int rc = m->each([&](int g) { return myzkp_srs_generate_g1(m->ranks[g], alpha_le, m->lo[g], m->hi[g] - m->lo[g]); });
  if (rc != MYZKP_OK) return rc;
  return m->finish_srs();
}

int myzkp_mctx_srs_load_g1(myzkp_mctx* m, const uint8_t* affine_xy_le, size_t n) {
  if (!m || (!affine_xy_le && n)) return MYZKP_ERR_INVALID_ARG;
  m->split(n);
  int rc = m->each([&](int g) { return myzkp_srs_load_g1(m->ranks[g], affine_xy_le + m->lo[g] * 64, m->hi[g] - m->lo[g]); });
  if (rc != MYZKP_OK) return rc;
  return m->finish_srs();
}

int myzkp_mctx_kzg_commit(myzkp_mctx* m, const uint8_t* coefs_le, size_t n, uint8_t out_c[64]) {
  if (!m || (!coefs_le && n) || !out_c) return MYZKP_ERR_INVALID_ARG;
  if (n > m->srs_n) {
    m->err = "polynomial longer than the SRS (reference panics at polynomial.rs:162)";
    return m->srs_n ? MYZKP_ERR_INVALID_ARG : MYZKP_ERR_NO_SRS;
  }
  const int G = (int)m->ranks.size();
  std::vector<uint8_t> outs((size_t)G * 64);
  int rc = m->each([&](int g) {
    size_t off, len;
    m->slice(n, g, &off, &len);
    return myzkp_kzg_commit_sharded(m->ranks[g], coefs_le + off * 32, len, outs.data() + 64 * g);
  });
  if (rc != MYZKP_OK) return rc;
  memcpy(out_c, outs.data(), 64);  // every rank holds the same commitment
  for (int g = 1; g < G; g++)
    if (memcmp(outs.data() + 64 * g, out_c, 64)) {
      m->err = "ranks disagree on the commitment";
      return MYZKP_ERR_CUDA;
    }
  return MYZKP_OK;
}

int myzkp_mctx_kzg_open(myzkp_mctx* m, const uint8_t* coefs_le, size_t n, const uint8_t u_le[32], uint8_t out_y[32],
                        uint8_t out_w[64]) {
  if (!m || (!coefs_le && n) || !u_le || !out_y || !out_w) return MYZKP_ERR_INVALID_ARG;
  if (n > m->srs_n + 1 || (n > 1 && !m->srs_n)) {
    m->err = "quotient longer than the SRS";
    return m->srs_n ? MYZKP_ERR_INVALID_ARG : MYZKP_ERR_NO_SRS;
  }
  const int G = (int)m->ranks.size();
  // a polynomial one longer than the SRS is legal (its quotient fits): the extra top coefficient rides with the
  // last non-empty rank, whose zero top quotient coefficient needs no SRS point
  std::vector<uint8_t> ys((size_t)G * 32), ws((size_t)G * 64);
  if (n == m->srs_n + 1 && n > 1) {
    m->err = "open of srs_len + 1 coefficients is not supported by the sharded path (use one more SRS power)";
    return MYZKP_ERR_INVALID_ARG;
  }
  int rc = m->each([&](int g) {
    size_t off, len;
    m->slice(n, g, &off, &len);
    return myzkp_kzg_open_sharded(m->ranks[g], coefs_le + off * 32, len, u_le, ys.data() + 32 * g, ws.data() + 64 * g);
  });
  if (rc != MYZKP_OK) return rc;
  memcpy(out_y, ys.data(), 32);
  memcpy(out_w, ws.data(), 64);
  for (int g = 1; g < G; g++)
    if (memcmp(ys.data() + 32 * g, out_y, 32) || memcmp(ws.data() + 64 * g, out_w, 64)) {
      m->err = "ranks disagree on the opening";
      return MYZKP_ERR_CUDA;
    }
  return MYZKP_OK;
}

}  // extern "C"
