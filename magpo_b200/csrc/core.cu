// Library-level entry points: version string and CUDA error reporting.
#include <string.h>

#include <atomic>
#include <vector>

#include "common.cuh"

namespace magpo {

static thread_local char g_err[512] = "";

void set_cuda_error(cudaError_t e, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "%s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e), file, line);
}

int ForkJoin::init() {
  if (s) return MAGPO_OK;
  MAGPO_CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  MAGPO_CUDA_OK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
  MAGPO_CUDA_OK(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
  return MAGPO_OK;
}
void ForkJoin::destroy() {
  if (fork) cudaEventDestroy(fork);
  if (join) cudaEventDestroy(join);
  if (s) cudaStreamDestroy(s);
  s = nullptr; fork = join = nullptr;
}

static thread_local Context* t_ctx = nullptr;
Context& ctx() {
  if (t_ctx) return *t_ctx;
  static thread_local Context scratch;  // magpo_test_* hooks only
  cudaGetDevice(&scratch.device);
  return scratch;
}
CtxScope::CtxScope(Context* c) : prev(t_ctx), ok(false) {
  int dev = -1;
  if (!c || cudaGetDevice(&dev) != cudaSuccess || dev != c->device) return;  // a context is bound to the device it was created on
  t_ctx = c;
  ok = true;
}
CtxScope::~CtxScope() { t_ctx = prev; }

bool once_per_device(int id) {
  constexpr int kMaxDev = 64;
  static std::atomic<bool> done[kMaxDev][ONCE_NUM];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDev || id < 0 || id >= ONCE_NUM) return true;
  return !done[dev][id].exchange(true);
}

void set_error_text(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }

static std::atomic<int64_t> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- profiling
static bool g_prof_on = false;
struct ProfRec { cudaEvent_t a, b; int cat; double work, bytes; };
static std::vector<ProfRec> g_recs;
static size_t g_used = 0;

ProfScope::ProfScope(int cat, cudaStream_t st, double work, double bytes) : idx(-1), s(st) {
  if (!g_prof_on) return;
  if (g_used == g_recs.size()) {
    ProfRec r;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    g_recs.push_back(r);
  }
  idx = (int)g_used++;
  g_recs[idx].cat = cat;
  g_recs[idx].work = work;
  g_recs[idx].bytes = bytes;
  cudaEventRecord(g_recs[idx].a, s);
}
ProfScope::~ProfScope() {
  if (idx >= 0) cudaEventRecord(g_recs[idx].b, s);
}

}  // namespace magpo

extern "C" {
int magpo_context_create(int32_t device, MagpoContext** out) {
  if (!out) return MAGPO_ERR_ARG;
  int cur = -1, n = 0;
  MAGPO_CUDA_OK(cudaGetDevice(&cur));
  MAGPO_CUDA_OK(cudaGetDeviceCount(&n));
  if (device < 0) device = cur;
  if (device >= n) return MAGPO_ERR_ARG;
  MagpoContext* c = new MagpoContext();
  c->device = device;
  // the forked streams belong to the context's device
  int rc = MAGPO_OK;
  if (cudaSetDevice(device) != cudaSuccess) rc = MAGPO_ERR_CUDA;
  if (rc == MAGPO_OK) rc = c->side.init();
  if (rc == MAGPO_OK) rc = c->dec.init();
  if (rc == MAGPO_OK) rc = c->rside.init();
  cudaSetDevice(cur);
  if (rc != MAGPO_OK) {
    c->side.destroy(); c->dec.destroy(); c->rside.destroy();
    delete c;
    return rc;
  }
  *out = c;
  return MAGPO_OK;
}
int magpo_context_destroy(MagpoContext* c) {
  if (!c) return MAGPO_OK;
  int cur = -1;
  cudaGetDevice(&cur);
  cudaSetDevice(c->device);
  c->side.destroy(); c->dec.destroy(); c->rside.destroy();
  cudaSetDevice(cur);
  delete c;
  return MAGPO_OK;
}
int64_t magpo_launch_count(void) { return magpo::g_launches.load(); }
// A host that replays a captured graph of library calls adds the launches of each replay (negative n: undo a capture pass).
int magpo_launch_count_add(int64_t n) {
  magpo::g_launches.fetch_add(n, std::memory_order_relaxed);
  return MAGPO_OK;
}
// on != 0: start recording (drops earlier records); on == 0: stop.
int magpo_prof_enable(int on) {
  magpo::g_prof_on = on != 0;
  if (on) magpo::g_used = 0;
  return MAGPO_OK;
}
// Sums over the recorded scopes of category `cat`: device ms, declared work (flops or bytes), scope count.
int magpo_prof_read(int cat, double* ms, double* work, int64_t* count) {
  using namespace magpo;
  if (!ms || !work || !count || cat < 0 || cat >= PROF_NUM) return MAGPO_ERR_ARG;
  if (cudaDeviceSynchronize() != cudaSuccess) return MAGPO_ERR_CUDA;
  *ms = 0; *work = 0; *count = 0;
  for (size_t i = 0; i < g_used; ++i) {
    if (g_recs[i].cat != cat) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_recs[i].a, g_recs[i].b) != cudaSuccess) continue;
    *ms += t; *work += g_recs[i].work; *count += 1;
  }
  return MAGPO_OK;
}
// Sum of the algorithmic HBM bytes declared by the recorded scopes of category `cat` (GEMM scopes declare flops as work).
int magpo_prof_read_bytes(int cat, double* bytes) {
  using namespace magpo;
  if (!bytes || cat < 0 || cat >= PROF_NUM) return MAGPO_ERR_ARG;
  *bytes = 0;
  for (size_t i = 0; i < g_used; ++i)
    if (g_recs[i].cat == cat) *bytes += g_recs[i].bytes;
  return MAGPO_OK;
}
const char* magpo_version(void) { return "magpo_b200 0.1 (sm_100a)"; }
const char* magpo_last_cuda_error(void) { return magpo::g_err; }
}
