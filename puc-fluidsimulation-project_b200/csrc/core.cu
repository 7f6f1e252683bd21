// core.cu -- library state (stream, error string, launch counter, timers) and
// the Triangle .node/.ele ingest (readNode/readEle, code/StokesColor.py:54-95).
#include <atomic>
#include <cerrno>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "internal.cuh"

namespace fs {
static thread_local std::string g_err;
static thread_local cudaStream_t g_user_stream = nullptr;
static thread_local bool g_use_user_stream = false;
static cudaStream_t g_own_stream = nullptr;
static std::atomic<int64_t> g_launches{0};
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
static int g_sm_count = 0;

void set_last_error(const std::string& s) { g_err = s; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() {
  static const bool v = [] { const char* e = std::getenv("FS_PDL"); return !e || std::atoi(e) != 0; }();
  return v;
}

// keep freed stream-ordered staging buffers in the pool: without this every host-buffer call
// re-maps tens of MB (default release threshold 0 returns the memory at each synchronise)
static void keep_mempool() {
  int dev = 0;
  cudaMemPool_t pool = nullptr;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  cudaGetLastError();
}

cudaStream_t stream() {
  if (g_use_user_stream) return g_user_stream;
  if (!g_own_stream) {
    keep_mempool();
    cudaError_t e = cudaStreamCreateWithFlags(&g_own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess)
      throw Error(FS_ERR_CUDA, std::string("no usable CUDA device (cudaStreamCreate: ") +
                                   cudaGetErrorString(e) + "); libfluidsim has no CPU fallback");
  }
  return g_own_stream;
}

int sm_count() {
  if (!g_sm_count) {
    int dev = 0;
    FS_CUDA(cudaGetDevice(&dev));
    FS_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  return g_sm_count;
}

// ---- text ingest ------------------------------------------------------------
static std::string slurp(const char* path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw Error(FS_ERR_IO, std::string("cannot open ") + path);
  std::ostringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

// whitespace-token cursor over a buffer, line aware (the reference reads one
// line per record with readline().split())
struct Cursor {
  const char* p;
  const char* end;
  bool next_line(const char*& b, const char*& e) {
    if (p >= end) return false;
    b = p;
    while (p < end && *p != '\n') ++p;
    e = p;
    if (p < end) ++p;
    return true;
  }
};

static inline void skip_ws(const char*& b, const char* e) {
  while (b < e && (*b == ' ' || *b == '\t' || *b == '\r')) ++b;
}

static long tok_long(const char*& b, const char* e, const char* what) {
  skip_ws(b, e);
  if (b >= e) throw Error(FS_ERR_IO, std::string("missing field: ") + what);
  char* q = nullptr;
  long v = std::strtol(b, &q, 10);
  if (q == b) throw Error(FS_ERR_IO, std::string("bad integer field: ") + what);
  b = q;
  return v;
}

static double tok_double(const char*& b, const char* e, const char* what) {
  skip_ws(b, e);
  if (b >= e) throw Error(FS_ERR_IO, std::string("missing field: ") + what);
  char* q = nullptr;
  double v = std::strtod(b, &q);   // correctly rounded, same as Python float()
  if (q == b) throw Error(FS_ERR_IO, std::string("bad float field: ") + what);
  b = q;
  return v;
}
}  // namespace fs

using namespace fs;

extern "C" {

int fs_version(void) { return 100; }
const char* fs_last_error(void) { return fs::g_err.c_str(); }

int fs_device_count(int* count) {
  FS_API_BEGIN
  FS_REQUIRE(count, "count is NULL");
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) { *count = 0; cudaGetLastError(); }
  FS_API_END
}

int fs_set_device(int device) {
  FS_API_BEGIN
  FS_CUDA(cudaSetDevice(device));
  g_sm_count = 0;
  if (g_own_stream) { cudaStreamDestroy(g_own_stream); g_own_stream = nullptr; }   // streams belong to a device
  if (g_ev0) { cudaEventDestroy(g_ev0); cudaEventDestroy(g_ev1); g_ev0 = g_ev1 = nullptr; }
  keep_mempool();
  FS_API_END
}

int fs_set_stream(void* s) {
  FS_API_BEGIN
  g_user_stream = (cudaStream_t)s;
  g_use_user_stream = (s != nullptr);
  FS_API_END
}

int fs_sync(void) {
  FS_API_BEGIN
  fs::sync();
  FS_API_END
}

// Order the library stream after everything queued so far on `producer` (the caller's stream that
// wrote a device-pointer argument): event record + stream wait, no host block.
int fs_stream_wait(void* producer) {
  FS_API_BEGIN
  cudaStream_t ps = (cudaStream_t)producer;
  cudaStream_t st = stream();
  if (ps == st) return FS_OK;
  static thread_local cudaEvent_t ev = nullptr;
  if (!ev) FS_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  FS_CUDA(cudaEventRecord(ev, ps));
  FS_CUDA(cudaStreamWaitEvent(st, ev, 0));
  FS_API_END
}

int64_t fs_launch_count(void) { return g_launches.load(); }

int fs_timer_start(void) {
  FS_API_BEGIN
  if (!g_ev0) { FS_CUDA(cudaEventCreate(&g_ev0)); FS_CUDA(cudaEventCreate(&g_ev1)); }
  FS_CUDA(cudaEventRecord(g_ev0, stream()));
  FS_API_END
}

int fs_timer_stop(float* ms) {
  FS_API_BEGIN
  FS_REQUIRE(ms && g_ev0, "timer not started");
  FS_CUDA(cudaEventRecord(g_ev1, stream()));
  FS_CUDA(cudaEventSynchronize(g_ev1));
  FS_CUDA(cudaEventElapsedTime(ms, g_ev0, g_ev1));
  FS_API_END
}

int fs_node_file_count(const char* path, int64_t* n) {
  FS_API_BEGIN
  FS_REQUIRE(path && n, "NULL argument");
  std::ifstream f(path);
  if (!f) throw Error(FS_ERR_IO, std::string("cannot open ") + path);
  long long v = -1;
  f >> v;
  if (!f || v < 0) throw Error(FS_ERR_IO, std::string("bad .node header in ") + path);
  *n = v;
  FS_API_END
}

int fs_read_node(const char* path, double* coords, int32_t* markers, int64_t n) {
  FS_API_BEGIN
  FS_REQUIRE(path && coords && markers, "NULL argument");
  std::string buf = slurp(path);
  Cursor c{buf.data(), buf.data() + buf.size()};
  const char *b, *e;
  if (!c.next_line(b, e)) throw Error(FS_ERR_IO, "empty .node file");
  long hn = tok_long(b, e, "node count");
  FS_REQUIRE(hn == n, "node count does not match the header");
  std::memset(coords, 0, sizeof(double) * 2 * n);
  std::memset(markers, 0, sizeof(int32_t) * n);
  for (int64_t i = 0; i < n; ++i) {
    if (!c.next_line(b, e)) throw Error(FS_ERR_IO, ".node file ends early");
    long id = tok_long(b, e, "node id") - 1;     // 1-based ids index the arrays
    if (id < 0 || id >= n) throw Error(FS_ERR_IO, "node id out of range");
    coords[2 * id] = tok_double(b, e, "x");
    coords[2 * id + 1] = tok_double(b, e, "y");
    long mk = tok_long(b, e, "boundary marker");
    if (mk != 0) markers[id] = (int32_t)mk;
  }
  FS_API_END
}

int fs_ele_file_count(const char* path, int64_t* t, int32_t* npt) {
  FS_API_BEGIN
  FS_REQUIRE(path && t, "NULL argument");
  std::ifstream f(path);
  if (!f) throw Error(FS_ERR_IO, std::string("cannot open ") + path);
  long long v = -1, k = 0;
  f >> v >> k;
  if (!f || v < 0) throw Error(FS_ERR_IO, std::string("bad .ele header in ") + path);
  *t = v;
  if (npt) *npt = (int32_t)k;
  FS_API_END
}

int fs_read_ele(const char* path, int32_t* tris, int64_t t) {
  FS_API_BEGIN
  FS_REQUIRE(path && tris, "NULL argument");
  std::string buf = slurp(path);
  Cursor c{buf.data(), buf.data() + buf.size()};
  const char *b, *e;
  if (!c.next_line(b, e)) throw Error(FS_ERR_IO, "empty .ele file");
  long ht = tok_long(b, e, "triangle count");
  FS_REQUIRE(ht == t, "triangle count does not match the header");
  for (int64_t i = 0; i < t; ++i) {
    if (!c.next_line(b, e)) throw Error(FS_ERR_IO, ".ele file ends early");
    (void)tok_long(b, e, "triangle id");         // ignored: line order is the id
    for (int k = 0; k < 3; ++k) tris[3 * i + k] = (int32_t)(tok_long(b, e, "corner") - 1);
  }
  FS_API_END
}

}  // extern "C"
