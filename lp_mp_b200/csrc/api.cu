// C ABI of the B200-native LP_MP message-passing hot path (include/lpmp_b200.h): handle, problem construction,
// schedule construction (host), kernel launches.  Host logic mirrors LP<FMC> (/root/reference/include/LP_MP.h:200-585)
// for the SRMP simplex configuration; all arithmetic on potentials happens in the kernels.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/lpmp_b200.h"
#include "../../include/lpmp_synthetic.h"
#include "graph_engine.cuh"
#include "kernels_general.cuh"
#include "wavefront_launch.h"

namespace lpmp {

template <typename Real>
cudaError_t launch_wavefront(const CUtensorMap* maps, const GridArgs<Real>& args, int pw, uint32_t L, int hops, int ksel, const WavefrontTuning& tune,
                             int num_sms, cudaStream_t stream) {
  const bool aniso = args.repam == LPMP_REPAM_ANISOTROPIC;
  if (pw == LPMP_PAIRWISE_POTTS)
    return aniso ? launch_family<Real, LPMP_PAIRWISE_POTTS, true>(maps, args, L, hops, ksel, tune, num_sms, stream)
                 : launch_family<Real, LPMP_PAIRWISE_POTTS, false>(maps, args, L, hops, ksel, tune, num_sms, stream);
  if (pw == LPMP_PAIRWISE_TRUNCLIN)
    return aniso ? launch_family<Real, LPMP_PAIRWISE_TRUNCLIN, true>(maps, args, L, hops, ksel, tune, num_sms, stream)
                 : launch_family<Real, LPMP_PAIRWISE_TRUNCLIN, false>(maps, args, L, hops, ksel, tune, num_sms, stream);
  return cudaErrorInvalidValue;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// ---- small helpers ----------------------------------------------------------------------------------------------
// Process-wide cache of released device blocks (exact-size reuse per device).  A solver loop that builds one LP per problem
// instance otherwise pays cudaMalloc / cudaFree for every buffer of every instance -- they synchronise the device and were
// measured at 1-400 ms per large block on B200, more than the solve itself (bench.py e2e arm).  Blocks stay cached up to a
// quarter of the device memory; an allocation that fails flushes the cache and retries.
class BlockCache {
 public:
  static BlockCache& get() { static BlockCache* c = new BlockCache; return *c; }   // leaked on purpose: no cudaFree after context teardown
  void* take(int dev, size_t bytes) {
    std::lock_guard<std::mutex> g(mu_);
    auto it = free_.find({dev, bytes});
    if (it == free_.end()) return nullptr;
    void* p = it->second;
    free_.erase(it);
    cached_[dev & 63] -= bytes;
    return p;
  }
  bool give(int dev, size_t bytes, void* p) {
    std::lock_guard<std::mutex> g(mu_);
    if (limit_[dev & 63] == 0) {
      size_t fr = 0, tot = 0;
      limit_[dev & 63] = cudaMemGetInfo(&fr, &tot) == cudaSuccess ? tot / 4 : (size_t)1 << 30;
    }
    if (cached_[dev & 63] + bytes > limit_[dev & 63]) return false;
    free_.insert({{dev, bytes}, p});
    cached_[dev & 63] += bytes;
    return true;
  }
  void flush(int dev) {
    std::lock_guard<std::mutex> g(mu_);
    for (auto it = free_.begin(); it != free_.end();) {
      if (it->first.first == dev) { cudaFree(it->second); cached_[dev & 63] -= it->first.second; it = free_.erase(it); }
      else ++it;
    }
  }
 private:
  std::mutex mu_;
  std::multimap<std::pair<int, size_t>, void*> free_;
  size_t cached_[64] = {}, limit_[64] = {};
};

// Process-wide pool of the handles' CUDA streams.  A stream is NEVER destroyed while the process lives: host frameworks that were
// handed lpmp_stream() (torch's caching allocator, NCCL process groups) keep the raw handle in their own bookkeeping -- e.g. they
// record an event on every stream a block was used on when the block is freed -- and a destroyed stream there is a crash long after
// lpmp_destroy.  Released streams are drained and reused by the next handle on the same device.
class StreamPool {
 public:
  static StreamPool& get() { static StreamPool* s = new StreamPool; return *s; }   // leaked on purpose, like BlockCache
  cudaError_t take(int dev, cudaStream_t* out) {
    {
      std::lock_guard<std::mutex> g(mu_);
      auto& v = idle_[dev & 63];
      if (!v.empty()) { *out = v.back(); v.pop_back(); return cudaSuccess; }
    }
    return cudaStreamCreateWithFlags(out, cudaStreamNonBlocking);
  }
  void give(int dev, cudaStream_t st) {
    cudaStreamSynchronize(st);
    (void)cudaGetLastError();
    std::lock_guard<std::mutex> g(mu_);
    idle_[dev & 63].push_back(st);
  }
 private:
  std::mutex mu_;
  std::vector<cudaStream_t> idle_[64];
};

template <typename T>
struct DeviceBuffer {
  T* p = nullptr;
  size_t n = 0;
  int dev = 0;
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    cudaGetDevice(&dev);
    const size_t bytes = count * sizeof(T);
    if (void* q = BlockCache::get().take(dev, bytes)) { p = static_cast<T*>(q); return cudaSuccess; }
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), bytes);
    if (e == cudaErrorMemoryAllocation) {   // make room: give the cached blocks back to the driver and try again
      (void)cudaGetLastError();
      BlockCache::get().flush(dev);
      e = cudaMalloc(reinterpret_cast<void**>(&p), bytes);
    }
    if (e != cudaSuccess) { p = nullptr; n = 0; }
    return e;
  }
  void release() {
    if (p) {
      // the block may still be in use by work queued on the owner's stream: wait for the device as cudaFree would
      int cur = 0;
      cudaGetDevice(&cur);
      if (cur != dev) cudaSetDevice(dev);
      cudaDeviceSynchronize();
      if (!BlockCache::get().give(dev, n * sizeof(T), p)) cudaFree(p);
      if (cur != dev) cudaSetDevice(cur);
    }
    p = nullptr; n = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
  ~DeviceBuffer() { release(); }
};

// compact [n][inner] source into records of `dst_stride` elements (and back)
template <typename Dst, typename Src>
__global__ void scatter_strided_kernel(Dst* __restrict__ dst, const Src* __restrict__ src, uint64_t n, uint32_t inner, uint64_t dst_stride) {
  const uint64_t total = n * inner;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (uint64_t)gridDim.x * blockDim.x)
    dst[(k / inner) * dst_stride + (k % inner)] = (Dst)src[k];
}
template <typename Dst, typename Src>
__global__ void gather_strided_kernel(Dst* __restrict__ dst, const Src* __restrict__ src, uint64_t n, uint32_t inner, uint64_t src_stride) {
  const uint64_t total = n * inner;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (uint64_t)gridDim.x * blockDim.x)
    dst[k] = (Dst)src[(k / inner) * src_stride + (k % inner)];
}

// parameters of a plane of grid edges, compact [rows][cols][P], into the parameter records of BOTH endpoints:
// edge at cell (r,c) -> node (r,c) field f_first and node (r,c) + second_off field f_second
template <typename Dst, typename Src>
__global__ void scatter_edge_params_kernel(Dst* __restrict__ par, const Src* __restrict__ src, uint32_t rows, uint32_t cols, uint32_t W, uint32_t P,
                                           uint32_t second_off, int f_first, int f_second, unsigned int* neg_flag) {
  const uint64_t n = (uint64_t)rows * cols * P;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t p = (uint32_t)(k % P);
    const uint64_t cell = k / P;
    const uint64_t node = (cell / cols) * W + (cell % cols);
    if (neg_flag != nullptr && !(src[k] >= Src(0))) *neg_flag = 1u;   // Potts: a negative (or NaN) lambda selects the two-minima update
    const Dst v = (Dst)src[k];
    if (f_first >= 0) par[(node * PAR_EDGES + f_first) * P + p] = v;
    if (f_second >= 0) par[((node + second_off) * PAR_EDGES + f_second) * P + p] = v;
  }
}

// synthetic potentials (include/lpmp_synthetic.h): counters are indices of the FULL grid so that a stripe holds
// exactly the values the unstriped LP would hold
template <typename Real>
__global__ void synth_unary_kernel(Real* __restrict__ rec, uint64_t n_nodes, uint32_t L, uint64_t rec_stride, uint64_t node0, uint64_t seed) {
  const uint64_t total = n_nodes * L;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t node = k / L;
    rec[node * rec_stride + (k % L)] = (Real)lpmp_synth_param(seed, LPMP_SYNTH_STREAM_UNARY, (node0 + node) * L + (k % L), 0.0, 1.0);
  }
}
// vertical == 0: edges (r,c)-(r,c+1); vertical == 1: edges (r,c)-(r+1,c) for r in [-1, H): r = -1 and r = H-1 are the
// stripe-boundary edges, of which only the endpoint inside the stripe is written
template <typename Real>
__global__ void synth_edge_kernel(Real* __restrict__ par, uint32_t H, uint32_t W, uint32_t P, uint32_t row0, uint32_t H_total, int vertical,
                                  uint64_t seed, double lo, double hi, int trunclin) {
  const uint64_t rows = vertical ? (uint64_t)H + 1 : H;
  const uint64_t n = rows * W * P;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t p = (uint32_t)(k % P);
    const uint64_t cell = k / P;
    const uint32_t c = (uint32_t)(cell % W);
    const int64_t r = (int64_t)(cell / W) - (vertical ? 1 : 0);           // row (inside the stripe) of the edge's first endpoint
    const int64_t gr = (int64_t)row0 + r;                                  // ... and in the full grid
    if (gr < 0) continue;
    if (vertical ? gr + 1 >= (int64_t)H_total : c + 1 >= W) continue;       // no such edge
    const uint64_t slot = (uint64_t)gr * W + c;
    const double v = trunclin ? (p == 0 ? lo : hi)
                              : lpmp_synth_param(seed, vertical ? LPMP_SYNTH_STREAM_VERTICAL : LPMP_SYNTH_STREAM_HORIZONTAL, slot * P + p, lo, hi);
    if (r >= 0) par[(((uint64_t)r * W + c) * PAR_EDGES + (vertical ? PAR_D : PAR_R)) * P + p] = (Real)v;
    if (vertical) { if (r + 1 < (int64_t)H) par[(((uint64_t)(r + 1) * W + c) * PAR_EDGES + PAR_U) * P + p] = (Real)v; }
    else par[(((uint64_t)r * W + c + 1) * PAR_EDGES + PAR_L) * P + p] = (Real)v;
  }
}

// dense tables of a plane of grid edges, compact [rows][cols][P], into the padded table plane (pitch W tables per row)
template <typename Dst, typename Src>
__global__ void scatter_tables_kernel(Dst* __restrict__ plane, const Src* __restrict__ src, uint32_t rows, uint32_t cols, uint32_t W, uint32_t P) {
  const uint64_t n = (uint64_t)rows * cols * P;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t cell = k / P;
    plane[((cell / cols) * W + (cell % cols)) * P + (k % P)] = (Dst)src[k];
  }
}
// synthetic dense tables: plane = tabH (vertical == 0, rows [0, H)) or tabV (vertical == 1, plane rows [0, H] hold the
// edges leaving grid rows [-1, H-1]); counters are slot indices of the FULL grid, as in synth_edge_kernel
template <typename Real>
__global__ void synth_tables_kernel(Real* __restrict__ plane, uint32_t H, uint32_t W, uint32_t P, uint32_t row0, uint32_t H_total, int vertical,
                                    uint64_t seed, double lo, double hi) {
  const uint64_t rows = vertical ? (uint64_t)H + 1 : H;
  const uint64_t n = rows * W * P;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t p = (uint32_t)(k % P);
    const uint64_t cell = k / P;
    const uint32_t c = (uint32_t)(cell % W);
    const int64_t gr = (int64_t)row0 + (int64_t)(cell / W) - (vertical ? 1 : 0);   // full-grid row of the edge's first endpoint
    if (gr < 0) continue;
    if (vertical ? gr + 1 >= (int64_t)H_total : c + 1 >= W) continue;               // no such edge
    const uint64_t slot = (uint64_t)gr * W + c;
    plane[k] = (Real)lpmp_synth_param(seed, vertical ? LPMP_SYNTH_STREAM_VERTICAL : LPMP_SYNTH_STREAM_HORIZONTAL, slot * P + p, lo, hi);
  }
}

// CheckPrimalConsistency: number of labels that are unset (0xffffffff) or out of range, added to *count (zeroed by the host)
__global__ void count_bad_labels_kernel(const uint32_t* __restrict__ labels, uint64_t n, uint32_t L, unsigned int* count) {
  unsigned int bad = 0;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) bad += labels[k] >= L ? 1u : 0u;
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(count, bad);
}

static inline int grid_for(uint64_t n, int block, int cap = 148 * 16) {
  uint64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > (uint64_t)cap) g = cap;
  return (int)g;
}

#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      err = std::string(#expr) + ": " + cudaGetErrorString(_e);                              \
      return LPMP_ERR_CUDA;                                                                  \
    }                                                                                        \
  } while (0)

struct ISolver {
  std::string err;
  virtual ~ISolver() {}
  virtual int add_unaries(uint64_t n, uint32_t L, const double* theta) = 0;
  virtual int add_unaries_var(uint64_t n, const uint32_t* L, const double* theta) = 0;
  virtual int dual_size(uint64_t* bytes) = 0;
  virtual int dual_archive(void* archive, uint64_t bytes, bool save) = 0;
  virtual int add_pairwise(int type, uint64_t m, const uint64_t* i, const uint64_t* j, const double* p0, const double* p1) = 0;
  virtual int set_node_ranks(const uint64_t* rank) = 0;
  virtual int set_edge_order(int eo) = 0;
  virtual int set_schedule(int s) = 0;
  virtual int set_tuning(const char* key, int value) = 0;
  virtual int grid_create(uint32_t H, uint32_t W, uint32_t L, int type) = 0;
  virtual int grid_set_unaries(const void* theta, int src) = 0;
  virtual int grid_set_pairwise(const void* hz, const void* vt, int src) = 0;
  virtual int grid_fill_synthetic(uint64_t seed, double lo, double hi) = 0;
  virtual int grid_set_stripe(uint32_t row0, uint32_t H_total) = 0;
  virtual int grid_set_boundary_pairwise(const void* top, const void* bottom, int src) = 0;
  virtual int grid_boundary_ptrs(int which, void** in_buf, uint64_t* bytes) = 0;
  virtual int grid_set_peer_boundary(int which, void* peer_in_buf) = 0;
  virtual int grid_get_boundary_vectors(int which, void* dst_dev) = 0;
  virtual int grid_lower_bound_partial(const void* below_b, double* out) = 0;
  virtual int grid_lower_bound_partial_device(const void* below_b, double* d_out) = 0;
  virtual int grid_evaluate_primal_partial_device(const void* below_b, const uint32_t* below_labels, double* d_out) = 0;
  virtual int grid_push_boundary(int which, void* peer_in_buf, int backward) = 0;
  virtual int grid_get_boundary_labels(int which, uint32_t* dst_dev) = 0;
  virtual int finalize() = 0;
  virtual int begin() = 0;
  virtual int set_reparametrization(int mode) = 0;
  virtual int set_send_mode(int mode) = 0;
  virtual int pass(bool forward, bool primal) = 0;
  virtual int put_in_same_partition(uint64_t m, const uint64_t* i, const uint64_t* j) = 0;
  virtual int partition_pass(uint64_t inner, int overlapping) = 0;
  virtual int lower_bound(double* out) = 0;
  virtual int evaluate_primal(double* out) = 0;
  virtual int check_primal_consistency(int* out) = 0;
  virtual int add_to_constant(double x) = 0;
  virtual int counts(uint64_t* factors, uint64_t* messages) = 0;
  virtual int get_labeling(uint32_t* out) = 0;
  virtual int keep_labeling() = 0;
  virtual int get_kept_labeling(uint32_t* out) = 0;
  virtual int get_unaries(double* th) = 0;
  virtual int get_messages(double* a, double* b) = 0;
  virtual int set_messages(const double* a, const double* b) = 0;
  virtual int set_unaries(const double* th) = 0;
  virtual int stats(lpmp_stats* out) = 0;
  virtual int synchronize() = 0;
  virtual int get_trace(uint64_t* out, uint64_t n) = 0;
  virtual void* stream_ptr() = 0;
};

template <typename Real>
struct Solver : ISolver {
  int device = 0, num_sms = 148;
  cudaStream_t stream = nullptr;
  // problem
  uint64_t N = 0, E = 0, slots = 0;
  uint32_t L = 0, gridH = 0, gridW = 0, row0 = 0, H_total = 0;
  int pw = -1, repam = -1, edge_order = LPMP_EDGE_ORDER_LIFO, schedule = -1, send_mode = LPMP_SEND_SHARED;
  bool finalized = false, labels_valid = false, is_grid = false, striped = false;
  double constant = 0.0;
  WavefrontTuning tune;
  WavefrontShape shape{0, false, false, 0, 0};
  bool potts_neg = false;    // Potts: a negative lambda has been seen
  int tl_reach = 0;          // truncated linear: max #{d >= 1: slope*d < trunc} over all edges seen so far
  bool tl_irregular = false; // a negative slope / trunc was seen: only the all-distances kernel is exact
  // host staging (general graphs)
  std::vector<double> h_theta, h_params;
  std::vector<uint64_t> h_ei, h_ej, h_rank;
  std::vector<uint32_t> h_counts;   // label count per variable of an LP with mixed counts (add_unaries_var), else empty
  // factors in insertion order as runs {kind (0 unary, 1 pairwise), first, count}: the order of the reference's f_ (include/LP_MP.h:214)
  std::vector<std::array<uint64_t, 3>> runs;
  void note_run(uint64_t kind, uint64_t first, uint64_t count) {
    if (count == 0) return;
    if (!runs.empty() && runs.back()[0] == kind && runs.back()[1] + runs.back()[2] == first) runs.back()[2] += count;
    else runs.push_back({kind, first, count});
  }
  bool mixed() const { return !h_counts.empty(); }
  uint32_t Ln(uint64_t i) const { return mixed() ? h_counts[i] : L; }
  GraphEngine<Real, DeviceBuffer> graph;   // general graphs: node-record layout + per-level kernels (graph_engine.cuh)
  // device state.  General graphs: theta[N][L], a[E][L], b[E][L], params[E][P].  Grids: rec[N][5][L], par[N][4][P].
  DeviceBuffer<Real> theta, a, b, params, rec_alloc, par;
  // grids: the node records sit between two GHOST rows (row -1 and row H; zero unless the stripe's uniform-mode passes park the
  // neighbouring stripe's boundary vectors there, later_gather_kernel): `rec` is the view of rows 0 .. H-1
  struct RecView { Real* p = nullptr; size_t n = 0; size_t bytes() const { return n * sizeof(Real); } } rec;
  DeviceBuffer<Real> tab;    // dense grids: table planes tabH [H][W][L*L] then tabV [H+1][W][L*L] (kernels_grid_dense.cuh)
  DeviceBuffer<double> lb_rows;   // dense grids: per-row bound terms produced by the last backward pass
  bool fused_lb_valid = false;    // lb_rows describes the current state (no state change since that pass)
  bool fused_lb_open_below = false;   // ... except for the edges to the stripe below (O(L) grids: added from below_b when the bound is asked for)
  bool levels_built = false; // CSR + level sets uploaded (general graphs; grids on demand)
  DeviceBuffer<Real> handoff, in_top, in_bot;   // band-to-band / stripe-to-stripe column vectors (sentinel protocol)
  Real* peer_top = nullptr;  // upper neighbour's in_bot (delivery target of the backward pass), or null
  Real* peer_bot = nullptr;  // lower neighbour's in_top (delivery target of the forward pass), or null
  DeviceBuffer<uint32_t> d_ei, d_ej, inc_edge, level_nodes, labels, kept_labels;
  bool kept_valid = false;        // kept_labels holds a labeling (keep_labeling)
  DeviceBuffer<uint64_t> inc_off;
  DeviceBuffer<uint8_t> inc_flags;
  DeviceBuffer<double> partial, result;
  DeviceBuffer<unsigned int> bad_count;     // CheckPrimalConsistency: labels that are unset / out of range
  DeviceBuffer<unsigned long long> trace;   // debug timeline of the last wavefront pass (set_tuning wavefront_trace=1)
  int want_trace = 0;        // 1: per-band stamps; 2: also per-step stamps of two neighbouring bands (lean variant)
  DeviceBuffer<unsigned int> control;  // [0] band ticket, [1] sticky abort flag (hand-off poll timed out), [2 + smid] CTA arrivals per SM
  CUtensorMap maps8[3], maps4[3];   // tile width 8 / 4 columns
  CUtensorMap maps_rows[2];         // one-label-per-lane variant: records / parameter records, rows_band_rows(L) / 4 rows x 8 columns
  bool rows_ok = false;             // ... its tensor maps exist
  CUtensorMap maps_split[2];        // edge-split variant: records / parameter records, 1 row x 8 columns
  bool split_ok = false;
  std::vector<uint64_t> level_off;     // host
  uint64_t launches = 0, updates = 0;
  static constexpr int kBoundCtas = 148 * 4;

  explicit Solver(int dev) : device(dev) {}
  ~Solver() override { if (stream) StreamPool::get().give(device, stream); }

  int init() {
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));   // (cudaGetDeviceProperties costs ~2 ms per handle)
    CUDA_TRY(StreamPool::get().take(device, &stream));
    CUDA_TRY(partial.alloc(kBoundCtas));
    CUDA_TRY(result.alloc(1));
    CUDA_TRY(bad_count.alloc(1));
    return LPMP_OK;
  }

  uint32_t P() const { return pairwise_param_count(pw, L); }
  uint64_t REC() const { return (uint64_t)REC_VECS * L; }
  uint64_t PREC() const { return (uint64_t)PAR_EDGES * P(); }
  int bad_state(const char* m) { err = m; return LPMP_ERR_STATE; }
  int need_final() { return finalized ? LPMP_OK : bad_state("finalize() has not been called"); }

  // truncated-linear window of one edge, evaluated in the storage type exactly like the kernels do
  void note_trunclin(double slope_d, double trunc_d) {
    const Real slope = (Real)slope_d, trunc = (Real)trunc_d;
    if (!(slope >= Real(0)) || !(trunc >= Real(0))) { tl_irregular = true; return; }
    int reach = 0;
    while (reach + 1 < (int)L && Real(slope * Real(reach + 1)) < trunc) ++reach;
    tl_reach = std::max(tl_reach, reach);
  }
  template <typename Src>
  void note_trunclin_array(const Src* p, uint64_t n_edges) {
    Src ls = 0, lt = 0;
    bool have = false;
    for (uint64_t e = 0; e < n_edges; ++e) {
      if (have && p[2 * e] == ls && p[2 * e + 1] == lt) continue;
      ls = p[2 * e]; lt = p[2 * e + 1]; have = true;
      note_trunclin((double)ls, (double)lt);
    }
  }
  // automatic choice for the anisotropic passes: edge-split variant where it exists (L = 16, 32), else one label per lane (L = 8;
  // L = 32 in double), else the band kernels (profiles/README.md, round 2)
  bool use_split() const { return tune.split > 0 || (tune.split < 0 && tune.rows <= 0 && tune.lean < 0 && tune.warps_per_cta <= 1); }
  bool use_rows() const { return tune.rows > 0 || (tune.rows < 0 && tune.split <= 0 && tune.lean < 0 && tune.warps_per_cta <= 1); }
  int tl_hops() const {
    if (pw != LPMP_PAIRWISE_TRUNCLIN || tl_irregular) return 0;
    if (tl_reach <= 4) return 1;
    if (tl_reach <= 8 && L >= 16) return 2;
    return 0;
  }

  // ---- construction: general graphs ---------------------------------------------------------------------------
  int add_unaries(uint64_t n, uint32_t L_, const double* th) override {
    if (finalized) return bad_state("add_unaries after finalize");
    if (is_grid) return bad_state("add_unaries on a grid LP");
    if (mixed()) { std::vector<uint32_t> Ls(n, L_); return add_unaries_var(n, Ls.data(), th); }
    if (L_ == 0 || (L != 0 && L != L_)) { err = "one label count per LP (lpmp_add_unaries_var takes mixed counts)"; return LPMP_ERR_UNSUPPORTED; }
    if (!th) { err = "null theta"; return LPMP_ERR_INVALID; }
    L = L_;
    h_theta.insert(h_theta.end(), th, th + n * L);
    note_run(0, N, n);
    N += n;
    return LPMP_OK;
  }
  int add_unaries_var(uint64_t n, const uint32_t* Ls, const double* th) override {
    if (finalized) return bad_state("add_unaries after finalize");
    if (is_grid) return bad_state("add_unaries on a grid LP");
    if (E != 0 && !mixed()) return bad_state("add_unaries_var after pairwise factors of an LP with one label count");
    if (!th || !Ls) { err = "null argument"; return LPMP_ERR_INVALID; }
    for (uint64_t q = 0; q < n; ++q) if (Ls[q] == 0) { err = "variable without labels"; return LPMP_ERR_INVALID; }
    if (!mixed()) h_counts.assign(N, L);   // what has been added so far had one label count
    uint64_t total = 0;
    for (uint64_t q = 0; q < n; ++q) { h_counts.push_back(Ls[q]); L = std::max(L, Ls[q]); total += Ls[q]; }
    h_theta.insert(h_theta.end(), th, th + total);
    note_run(0, N, n);
    N += n;
    return LPMP_OK;
  }
  int add_pairwise(int type, uint64_t m, const uint64_t* i, const uint64_t* j, const double* p0, const double* p1) override {
    if (finalized) return bad_state("add_pairwise after finalize");
    if (is_grid) return bad_state("add_pairwise on a grid LP");
    if (pw != -1 && pw != type) { err = "one pairwise family per LP"; return LPMP_ERR_UNSUPPORTED; }
    for (uint64_t k = 0; k < m; ++k)
      if (i[k] >= N || j[k] >= N || i[k] == j[k]) { err = "edge endpoint out of range"; return LPMP_ERR_INVALID; }
    pw = type;
    h_ei.insert(h_ei.end(), i, i + m);
    h_ej.insert(h_ej.end(), j, j + m);
    if (type == LPMP_PAIRWISE_DENSE) {
      uint64_t total = 0;   // tables of an LP with mixed label counts are packed: L_i x L_j each
      for (uint64_t k = 0; k < m; ++k) total += (uint64_t)Ln(i[k]) * Ln(j[k]);
      h_params.insert(h_params.end(), p0, p0 + total);
    } else if (type == LPMP_PAIRWISE_POTTS) h_params.insert(h_params.end(), p0, p0 + m);
    else for (uint64_t k = 0; k < m; ++k) { h_params.push_back(p0[k]); h_params.push_back(p1[k]); }
    note_run(1, E, m);
    E += m;
    return LPMP_OK;
  }
  int set_node_ranks(const uint64_t* rank) override {
    if (finalized) return bad_state("set_node_ranks after finalize");
    h_rank.assign(rank, rank + N);
    return LPMP_OK;
  }
  int set_edge_order(int eo) override {
    if (finalized) return bad_state("set_edge_order after finalize");
    if (eo != LPMP_EDGE_ORDER_LIFO && eo != LPMP_EDGE_ORDER_FIFO) { err = "bad edge order"; return LPMP_ERR_INVALID; }
    edge_order = eo;
    return LPMP_OK;
  }
  int set_schedule(int s) override {
    if (finalized) return bad_state("set_schedule after finalize");
    if (s != LPMP_SCHEDULE_LEVELS && s != LPMP_SCHEDULE_WAVEFRONT) { err = "bad schedule"; return LPMP_ERR_INVALID; }
    schedule = s;
    return LPMP_OK;
  }
  int set_tuning(const char* key, int value) override {
    const std::string k(key ? key : "");
    if (k == "wavefront_max_ctas") tune.max_ctas = value;
    else if (k == "wavefront_ctas_per_sm") tune.ctas_per_sm = value;
    else if (k == "wavefront_tile_cols") tune.tile_cols = value;
    else if (k == "wavefront_warps_per_cta") tune.warps_per_cta = value;
    else if (k == "wavefront_lean") tune.lean = value;
    else if (k == "wavefront_cluster") tune.cluster = value;
    else if (k == "level_group") tune.level_group = value;
    else if (k == "graph_ctas_per_sm") graph.ctas_per_sm = value;
    else if (k == "graph_fused_bound") graph.fused_bound_enabled = value != 0;
    else if (k == "wavefront_rows") tune.rows = value;
    else if (k == "wavefront_split") tune.split = value;
    else if (k == "wavefront_split_uni") tune.split_uni = value;
    else if (k == "wavefront_fused_bound") tune.fused_bound = value;
    else if (k == "wavefront_trace") want_trace = value;
    else { err = "unknown tuning key " + k; return LPMP_ERR_INVALID; }
    return LPMP_OK;
  }

  // ---- construction: grids --------------------------------------------------------------------------------------
  int grid_create(uint32_t H, uint32_t W, uint32_t L_, int type) override {
    if (N != 0 || finalized) return bad_state("grid_create on a non-empty LP");
    if (H == 0 || W == 0 || L_ == 0 || type < 0 || type > 2) { err = "bad grid shape"; return LPMP_ERR_INVALID; }
    is_grid = true;
    gridH = H; gridW = W; L = L_; pw = type;
    row0 = 0; H_total = H;
    N = (uint64_t)H * W;
    E = (uint64_t)H * (W - 1) + (uint64_t)(H - 1) * W;
    slots = 2 * N;
    note_run(0, 0, N);
    note_run(1, 0, E);
    CUDA_TRY(cudaSetDevice(device));
    const uint64_t ghost = (((uint64_t)W * REC() + 63) / 64) * 64;   // one row of records, rounded to 256 / 512 bytes
    CUDA_TRY(rec_alloc.alloc(N * REC() + 2 * ghost));
    CUDA_TRY(cudaMemsetAsync(rec_alloc.p, 0, rec_alloc.bytes(), stream));
    rec.p = rec_alloc.p + ghost; rec.n = N * REC();
    if (pw == LPMP_PAIRWISE_DENSE) {
      CUDA_TRY(tab.alloc((2 * N + W) * (uint64_t)P()));
      CUDA_TRY(cudaMemsetAsync(tab.p, 0, tab.bytes(), stream));
    } else {
      CUDA_TRY(par.alloc(N * PREC()));
      CUDA_TRY(cudaMemsetAsync(par.p, 0, par.bytes(), stream));
    }
    return LPMP_OK;
  }
  template <typename Src>
  int upload(DeviceBuffer<Src>& tmp, const Src* src, uint64_t n) {
    CUDA_TRY(tmp.alloc(n));
    CUDA_TRY(cudaMemcpyAsync(tmp.p, src, n * sizeof(Src), cudaMemcpyHostToDevice, stream));
    return LPMP_OK;
  }
  template <typename Src>
  int set_unaries_from(const Src* src) {
    DeviceBuffer<Src> tmp;
    if (int r = upload(tmp, src, N * L)) return r;
    scatter_strided_kernel<Real, Src><<<grid_for(N * L, 256), 256, 0, stream>>>(rec.p, tmp.p, N, L, REC());
    ++launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(stream));
    return LPMP_OK;
  }
  int grid_set_unaries(const void* th, int src) override {
    if (!is_grid) return bad_state("not a grid");
    fused_lb_valid = false;
    if (!th) { err = "null theta"; return LPMP_ERR_INVALID; }
    return src == LPMP_F32 ? set_unaries_from(static_cast<const float*>(th)) : set_unaries_from(static_cast<const double*>(th));
  }
  template <typename Src>
  int set_edge_plane(const Src* src, uint32_t rows, uint32_t cols, uint32_t second_off, int f_first, int f_second) {
    const uint32_t Pn = P();
    const uint64_t n = (uint64_t)rows * cols * Pn;
    if (n == 0 || !src) return LPMP_OK;
    if (pw == LPMP_PAIRWISE_TRUNCLIN) note_trunclin_array(src, (uint64_t)rows * cols);
    DeviceBuffer<Src> tmp;
    if (int r = upload(tmp, src, n)) return r;
    // Potts: the sign scan of the lambdas rides on the scatter kernel (a host loop over the 33.5 M edges of C3 costs more than their upload)
    unsigned int* const neg_flag = pw == LPMP_PAIRWISE_POTTS && !potts_neg ? bad_count.p : nullptr;
    if (neg_flag) CUDA_TRY(cudaMemsetAsync(neg_flag, 0, sizeof(unsigned int), stream));
    scatter_edge_params_kernel<Real, Src><<<grid_for(n, 256), 256, 0, stream>>>(par.p, tmp.p, rows, cols, gridW, Pn, second_off, f_first, f_second, neg_flag);
    ++launches;
    CUDA_TRY(cudaGetLastError());
    return finish_edge_plane(neg_flag);
  }
  int finish_edge_plane(unsigned int* neg_flag) {
    unsigned int neg = 0;
    if (neg_flag) CUDA_TRY(cudaMemcpyAsync(&neg, neg_flag, sizeof(neg), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (neg) potts_neg = true;
    return LPMP_OK;
  }
  Real* tabH() const { return tab.p; }
  Real* tabV() const { return tab.p + N * (uint64_t)P(); }   // plane row k = edges leaving grid row k-1
  template <typename Src>
  int set_table_plane(const Src* src, uint32_t rows, uint32_t cols, Real* plane) {
    const uint64_t n = (uint64_t)rows * cols * P();
    if (n == 0 || !src) return LPMP_OK;
    DeviceBuffer<Src> tmp;
    if (int r = upload(tmp, src, n)) return r;
    scatter_tables_kernel<Real, Src><<<grid_for(n, 256), 256, 0, stream>>>(plane, tmp.p, rows, cols, gridW, P());
    ++launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(stream));
    return LPMP_OK;
  }
  int grid_set_pairwise(const void* hz, const void* vt, int src) override {
    if (!is_grid) return bad_state("not a grid");
    if (finalized) return bad_state("grid_set_pairwise after finalize");
    fused_lb_valid = false;
    int r;
    if (pw == LPMP_PAIRWISE_DENSE) {
      Real* const vplane = tabV() + (uint64_t)gridW * P();
      if (src == LPMP_F32) {
        r = set_table_plane(static_cast<const float*>(hz), gridH, gridW - 1, tabH());
        if (r == LPMP_OK) r = set_table_plane(static_cast<const float*>(vt), gridH - 1, gridW, vplane);
      } else {
        r = set_table_plane(static_cast<const double*>(hz), gridH, gridW - 1, tabH());
        if (r == LPMP_OK) r = set_table_plane(static_cast<const double*>(vt), gridH - 1, gridW, vplane);
      }
      return r;
    }
    if (src == LPMP_F32) {
      r = set_edge_plane(static_cast<const float*>(hz), gridH, gridW - 1, 1, PAR_R, PAR_L);
      if (r == LPMP_OK) r = set_edge_plane(static_cast<const float*>(vt), gridH - 1, gridW, gridW, PAR_D, PAR_U);
    } else {
      r = set_edge_plane(static_cast<const double*>(hz), gridH, gridW - 1, 1, PAR_R, PAR_L);
      if (r == LPMP_OK) r = set_edge_plane(static_cast<const double*>(vt), gridH - 1, gridW, gridW, PAR_D, PAR_U);
    }
    return r;
  }
  int grid_fill_synthetic(uint64_t seed, double lo, double hi) override {
    if (!is_grid) return bad_state("not a grid");
    if (finalized) return bad_state("grid_fill_synthetic after finalize");
    fused_lb_valid = false;
    const uint32_t Pn = P();
    if (pw == LPMP_PAIRWISE_DENSE) {
      synth_unary_kernel<Real><<<grid_for(N * L, 256), 256, 0, stream>>>(rec.p, N, L, REC(), (uint64_t)row0 * gridW, seed);
      synth_tables_kernel<Real><<<grid_for(N * Pn, 256, 148 * 32), 256, 0, stream>>>(tabH(), gridH, gridW, Pn, row0, H_total, 0, seed, lo, hi);
      synth_tables_kernel<Real><<<grid_for((N + gridW) * Pn, 256, 148 * 32), 256, 0, stream>>>(tabV(), gridH, gridW, Pn, row0, H_total, 1, seed, lo, hi);
      launches += 3;
      CUDA_TRY(cudaGetLastError());
      return LPMP_OK;
    }
    const int tl = pw == LPMP_PAIRWISE_TRUNCLIN;
    if (tl) note_trunclin(lo, hi);
    else if (!(lo >= 0.0)) potts_neg = true;
    synth_unary_kernel<Real><<<grid_for(N * L, 256), 256, 0, stream>>>(rec.p, N, L, REC(), (uint64_t)row0 * gridW, seed);
    synth_edge_kernel<Real><<<grid_for(N * Pn, 256), 256, 0, stream>>>(par.p, gridH, gridW, Pn, row0, H_total, 0, seed, lo, hi, tl);
    synth_edge_kernel<Real><<<grid_for((N + gridW) * Pn, 256), 256, 0, stream>>>(par.p, gridH, gridW, Pn, row0, H_total, 1, seed, lo, hi, tl);
    launches += 3;
    CUDA_TRY(cudaGetLastError());
    return LPMP_OK;
  }
  int grid_set_stripe(uint32_t r0, uint32_t Ht) override {
    if (!is_grid || finalized) return bad_state("grid_set_stripe: need an unfinalized grid");
    if ((uint64_t)r0 + gridH > Ht) { err = "stripe exceeds the full grid"; return LPMP_ERR_INVALID; }
    row0 = r0; H_total = Ht; striped = true;
    const uint64_t n = (uint64_t)gridW * L;
    if (row0 > 0) { CUDA_TRY(in_top.alloc(n)); CUDA_TRY(cudaMemsetAsync(in_top.p, 0xff, in_top.bytes(), stream)); }
    if (row0 + gridH < H_total) { CUDA_TRY(in_bot.alloc(n)); CUDA_TRY(cudaMemsetAsync(in_bot.p, 0xff, in_bot.bytes(), stream)); }
    CUDA_TRY(cudaStreamSynchronize(stream));
    return LPMP_OK;
  }
  int grid_set_boundary_pairwise(const void* top, const void* bottom, int src) override {
    if (!striped) return bad_state("grid_set_boundary_pairwise: not a stripe");
    if (pw == LPMP_PAIRWISE_DENSE) { err = "stripe-boundary tables: only the synthetic fill is supported for dense grids"; return LPMP_ERR_UNSUPPORTED; }
    int r = LPMP_OK;
    // top[c]: edge (row0-1, c)-(row0, c): field U of the first row; bottom[c]: edge (last row, c)-(below): field D of the last row
    if (src == LPMP_F32) {
      if (top && row0 > 0) r = set_edge_plane(static_cast<const float*>(top), 1, gridW, 0, PAR_U, -1);
      if (r == LPMP_OK && bottom && row0 + gridH < H_total) r = set_bottom_plane(static_cast<const float*>(bottom));
    } else {
      if (top && row0 > 0) r = set_edge_plane(static_cast<const double*>(top), 1, gridW, 0, PAR_U, -1);
      if (r == LPMP_OK && bottom && row0 + gridH < H_total) r = set_bottom_plane(static_cast<const double*>(bottom));
    }
    return r;
  }
  template <typename Src>
  int set_bottom_plane(const Src* bottom) {
    const uint32_t Pn = P();
    const uint64_t n = (uint64_t)gridW * Pn;
    if (pw == LPMP_PAIRWISE_TRUNCLIN) note_trunclin_array(bottom, gridW);
    DeviceBuffer<Src> tmp;
    if (int r = upload(tmp, bottom, n)) return r;
    unsigned int* const neg_flag = pw == LPMP_PAIRWISE_POTTS && !potts_neg ? bad_count.p : nullptr;
    if (neg_flag) CUDA_TRY(cudaMemsetAsync(neg_flag, 0, sizeof(unsigned int), stream));
    scatter_edge_params_kernel<Real, Src><<<grid_for(n, 256), 256, 0, stream>>>(par.p + (uint64_t)(gridH - 1) * gridW * PREC(), tmp.p, 1, gridW, gridW, Pn, 0, PAR_D, -1, neg_flag);
    ++launches;
    CUDA_TRY(cudaGetLastError());
    return finish_edge_plane(neg_flag);
  }
  int grid_boundary_ptrs(int which, void** in_buf, uint64_t* bytes) override {
    if (!striped) return bad_state("grid_boundary_ptrs: not a stripe");
    if (in_buf) *in_buf = which == 0 ? (void*)in_top.p : (void*)in_bot.p;
    if (bytes) *bytes = (uint64_t)gridW * L * sizeof(Real);
    return LPMP_OK;
  }
  int grid_set_peer_boundary(int which, void* peer_in_buf) override {
    if (!striped) return bad_state("grid_set_peer_boundary: not a stripe");
    (which == 0 ? peer_top : peer_bot) = static_cast<Real*>(peer_in_buf);
    return LPMP_OK;
  }
  // own side of the stripe-boundary edges: bU of the first row (which = 0), aD of the last row (which = 1)
  int grid_get_boundary_vectors(int which, void* dst_dev) override {
    if (!is_grid) return bad_state("not a grid");
    const Real* src = which == 0 ? rec.p + (uint64_t)REC_BU * L : rec.p + (uint64_t)(gridH - 1) * gridW * REC() + (uint64_t)REC_AD * L;
    gather_strided_kernel<Real, Real><<<grid_for((uint64_t)gridW * L, 256), 256, 0, stream>>>(static_cast<Real*>(dst_dev), src, gridW, L, REC());
    ++launches;
    CUDA_TRY(cudaGetLastError());
    return LPMP_OK;
  }

  // ---- finalize: SortFactors + weights (include/LP_MP.h:730-797,1230-1449) realised as CSR + level sets ---------
  // edge e of a grid in insertion order (per node: horizontal, then vertical) -> storage slot
  void grid_edge_lists(std::vector<uint64_t>& ei, std::vector<uint64_t>& ej, std::vector<uint64_t>& slot) const {
    for (uint32_t r = 0; r < gridH; ++r)
      for (uint32_t c = 0; c < gridW; ++c) {
        const uint64_t i = (uint64_t)r * gridW + c;
        if (c + 1 < gridW) { ei.push_back(i); ej.push_back(i + 1); slot.push_back(i); }
        if (r + 1 < gridH) { ei.push_back(i); ej.push_back(i + gridW); slot.push_back(N + i); }
      }
  }

  Store<Real> store() const {
    Store<Real> s;
    s.L_ = L;
    if (is_grid) {
      const uint64_t R = REC(), Pn = P();
      s.theta = rec.p; s.theta_stride = R;
      s.a0 = rec.p + (uint64_t)REC_AR * L; s.a1 = rec.p + (uint64_t)REC_AD * L;
      s.b0 = rec.p + R + (uint64_t)REC_BL * L; s.b1 = rec.p + (uint64_t)gridW * R + (uint64_t)REC_BU * L;
      s.split = N; s.estride = R;
      if (pw == LPMP_PAIRWISE_DENSE) { s.p0 = tabH(); s.p1 = tabV() + (uint64_t)gridW * Pn; s.pstride = Pn; }
      else { s.p0 = par.p + (uint64_t)PAR_R * Pn; s.p1 = par.p + (uint64_t)PAR_D * Pn; s.pstride = PREC(); }
    } else {
      s.theta = theta.p; s.theta_stride = L;
      s.a0 = s.a1 = a.p; s.b0 = s.b1 = b.p; s.p0 = s.p1 = params.p;
      s.split = ~uint64_t(0); s.estride = L; s.pstride = P();
    }
    return s;
  }

  int build_levels(const std::vector<uint64_t>& ei, const std::vector<uint64_t>& ej, const std::vector<uint64_t>& slot) {
    std::vector<uint64_t> rank = h_rank;
    if (rank.empty()) { rank.resize(N); std::iota(rank.begin(), rank.end(), uint64_t(0)); }
    {
      std::vector<uint64_t> sorted = rank;
      std::sort(sorted.begin(), sorted.end());
      if (std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end()) { err = "node ranks must be distinct"; return LPMP_ERR_INVALID; }
    }
    const uint64_t m = ei.size();
    // per-node message order: first-variable type, then second-variable type; LIFO / FIFO inside a type
    std::vector<uint64_t> cnt_first(N + 1, 0), cnt_second(N + 1, 0);
    for (uint64_t e = 0; e < m; ++e) { ++cnt_first[ei[e]]; ++cnt_second[ej[e]]; }
    std::vector<uint64_t> off(N + 1, 0);
    for (uint64_t i = 0; i < N; ++i) off[i + 1] = off[i] + cnt_first[i] + cnt_second[i];
    std::vector<uint32_t> inc_e(2 * m);
    std::vector<uint8_t> inc_f(2 * m);
    std::vector<uint64_t> pos_first(N), pos_second(N);
    for (uint64_t i = 0; i < N; ++i) { pos_first[i] = off[i]; pos_second[i] = off[i] + cnt_first[i]; }
    const bool lifo = edge_order == LPMP_EDGE_ORDER_LIFO;
    for (uint64_t e = 0; e < m; ++e) {
      const uint64_t i = ei[e], j = ej[e];
      const uint64_t ki = lifo ? off[i] + cnt_first[i] - 1 - (pos_first[i] - off[i]) : pos_first[i];
      const uint64_t kj = lifo ? off[j] + cnt_first[j] + cnt_second[j] - 1 - (pos_second[j] - off[j] - cnt_first[j]) : pos_second[j];
      ++pos_first[i]; ++pos_second[j];
      inc_e[ki] = (uint32_t)slot[e]; inc_f[ki] = (uint8_t)(0u | ((rank[j] > rank[i] ? 1u : 0u) << 1));
      inc_e[kj] = (uint32_t)slot[e]; inc_f[kj] = (uint8_t)(1u | ((rank[i] > rank[j] ? 1u : 0u) << 1));
    }
    // levels = longest path from the sources of the DAG "lower rank -> higher rank"
    std::vector<uint64_t> order(N);
    std::iota(order.begin(), order.end(), uint64_t(0));
    std::sort(order.begin(), order.end(), [&](uint64_t x, uint64_t y) { return rank[x] < rank[y]; });
    std::vector<uint32_t> level(N, 0);
    std::vector<uint64_t> adj_off(N + 1, 0);
    for (uint64_t e = 0; e < m; ++e) { ++adj_off[ei[e] + 1]; ++adj_off[ej[e] + 1]; }
    for (uint64_t i = 0; i < N; ++i) adj_off[i + 1] += adj_off[i];
    std::vector<uint32_t> adj(2 * m);
    { std::vector<uint64_t> pos(adj_off.begin(), adj_off.end() - 1);
      for (uint64_t e = 0; e < m; ++e) { adj[pos[ei[e]]++] = (uint32_t)ej[e]; adj[pos[ej[e]]++] = (uint32_t)ei[e]; } }
    uint32_t max_level = 0;
    for (uint64_t k = 0; k < N; ++k) {
      const uint64_t i = order[k];
      uint32_t lv = 0;
      for (uint64_t q = adj_off[i]; q < adj_off[i + 1]; ++q) if (rank[adj[q]] < rank[i]) lv = std::max(lv, level[adj[q]] + 1);
      level[i] = lv;
      max_level = std::max(max_level, lv);
    }
    level_off.assign((size_t)max_level + 2, 0);
    for (uint64_t i = 0; i < N; ++i) ++level_off[level[i] + 1];
    for (size_t l = 0; l + 1 < level_off.size(); ++l) level_off[l + 1] += level_off[l];
    std::vector<uint32_t> lnodes(N);
    { std::vector<uint64_t> pos(level_off.begin(), level_off.end() - 1);
      for (uint64_t k = 0; k < N; ++k) { const uint64_t i = order[k]; lnodes[pos[level[i]]++] = (uint32_t)i; } }
    CUDA_TRY(inc_off.alloc(N + 1));
    CUDA_TRY(inc_edge.alloc(std::max<uint64_t>(1, 2 * m)));
    CUDA_TRY(inc_flags.alloc(std::max<uint64_t>(1, 2 * m)));
    CUDA_TRY(level_nodes.alloc(N));
    CUDA_TRY(cudaMemcpyAsync(inc_off.p, off.data(), (N + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
    if (m) CUDA_TRY(cudaMemcpyAsync(inc_edge.p, inc_e.data(), 2 * m * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    if (m) CUDA_TRY(cudaMemcpyAsync(inc_flags.p, inc_f.data(), 2 * m, cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(level_nodes.p, lnodes.data(), N * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    levels_built = true;
    return LPMP_OK;
  }
  int build_grid_levels() {
    std::vector<uint64_t> ei, ej, slot;
    grid_edge_lists(ei, ej, slot);
    return build_levels(ei, ej, slot);
  }

  // TMA descriptors of the record / parameter arrays: tiles of `rows` grid rows x K columns of whole records
  int encode_map(CUtensorMap* m, Real* base, uint32_t inner, uint32_t vecs, uint32_t box_rows, uint32_t box_cols) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { err = "cuTensorMapEncodeTiled is not available in this driver"; return LPMP_ERR_CUDA; }
    const CUtensorMapDataType dt = sizeof(Real) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    CUresult r;
    if (vecs > 1) {   // (label, vector, column, row)
      const cuuint64_t dims[4] = {inner, vecs, gridW, gridH};
      const cuuint64_t strides[3] = {inner * sizeof(Real), (cuuint64_t)inner * vecs * sizeof(Real), (cuuint64_t)gridW * inner * vecs * sizeof(Real)};
      const cuuint32_t box[4] = {inner, vecs, box_cols, box_rows};
      const cuuint32_t es[4] = {1, 1, 1, 1};
      r = fn(m, dt, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {          // (parameter, column, row)
      const cuuint64_t dims[3] = {inner, gridW, gridH};
      const cuuint64_t strides[2] = {inner * sizeof(Real), (cuuint64_t)gridW * inner * sizeof(Real)};
      const cuuint32_t box[3] = {inner, box_cols, box_rows};
      const cuuint32_t es[3] = {1, 1, 1};
      r = fn(m, dt, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r); return LPMP_ERR_CUDA; }
    return LPMP_OK;
  }

  int finalize() override {
    if (finalized) return bad_state("finalize called twice");
    if (N == 0) return bad_state("no factors");
    if (N >= (1ull << 31)) { err = "more than 2^31 nodes"; return LPMP_ERR_UNSUPPORTED; }
    CUDA_TRY(cudaSetDevice(device));
    if (pw == -1) pw = LPMP_PAIRWISE_POTTS;
    if (!is_grid) {
      slots = E;
      if (schedule == LPMP_SCHEDULE_WAVEFRONT) { err = "wavefront schedule needs a grid"; return LPMP_ERR_UNSUPPORTED; }
      schedule = LPMP_SCHEDULE_LEVELS;
      if (int r = graph.build(N, L, h_counts, h_theta, h_ei, h_ej, h_params, pw, h_rank, edge_order == LPMP_EDGE_ORDER_LIFO, stream, err)) return r;
      if (partial.n < graph.partial_slots()) CUDA_TRY(partial.alloc(graph.partial_slots()));
      levels_built = true;
      h_theta.clear(); h_theta.shrink_to_fit(); h_params.clear(); h_params.shrink_to_fit();
    } else {
      const bool dense = pw == LPMP_PAIRWISE_DENSE;
      const bool wf_ok = (dense ? dense_wavefront_supported<Real>(L) : wavefront_shape<Real>(pw, L, shape)) && h_rank.empty();
      if (schedule == -1) schedule = wf_ok ? LPMP_SCHEDULE_WAVEFRONT : LPMP_SCHEDULE_LEVELS;
      if (schedule == LPMP_SCHEDULE_WAVEFRONT && !wf_ok) { err = "no wavefront kernel for this label count / pairwise family"; return LPMP_ERR_UNSUPPORTED; }
      if (schedule == LPMP_SCHEDULE_LEVELS) {
        if (striped) { err = "stripes need the wavefront schedule"; return LPMP_ERR_UNSUPPORTED; }
        if (int r = build_grid_levels()) return r;
      } else if (dense) {
        CUDA_TRY(control.alloc(2 + 1024));   // + one arrival counter per SM (role rotation of the lean kernel's warps)
        CUDA_TRY(cudaMemsetAsync(control.p, 0, control.bytes(), stream));
        CUDA_TRY(handoff.alloc((size_t)gridH * gridW * L));
        CUDA_TRY(cudaMemsetAsync(handoff.p, 0xff, handoff.bytes(), stream));
        CUDA_TRY(lb_rows.alloc(gridH));
      } else {
        const uint32_t bands = (gridH + shape.T - 1) / shape.T;
        CUDA_TRY(control.alloc(2 + 1024));   // + one arrival counter per SM (role rotation of the lean kernel's warps)
        CUDA_TRY(cudaMemsetAsync(control.p, 0, control.bytes(), stream));
        if (want_trace) { CUDA_TRY(trace.alloc(4 * (size_t)bands + (want_trace >= 2 ? 2 * 8192 : 0))); CUDA_TRY(cudaMemsetAsync(trace.p, 0, trace.bytes(), stream)); }
        CUDA_TRY(handoff.alloc((size_t)bands * gridW * L));
        CUDA_TRY(cudaMemsetAsync(handoff.p, 0xff, handoff.bytes(), stream));
        for (int ks = 4; ks <= 8; ks += 4) {
          if (!(ks == 8 ? shape.fits8 : shape.fits4)) continue;
          CUtensorMap* m = ks == 8 ? maps8 : maps4;
          if (int r = encode_map(&m[0], rec.p, L, REC_VECS, (uint32_t)shape.T, (uint32_t)ks)) return r;
          if (int r = encode_map(&m[1], rec.p, L, REC_VECS, (uint32_t)shape.T + 1, (uint32_t)ks)) return r;
          if (int r = encode_map(&m[2], par.p, (uint32_t)PREC(), 1, (uint32_t)shape.T, (uint32_t)ks)) return r;
        }
        if (rows_supported<Real>(pw, L) && rows_band_rows(L) >= shape.T) {   // (the hand-off buffer is sized for bands of shape.T rows)
          const uint32_t rpw = (uint32_t)rows_band_rows(L) / 4;
          if (int r = encode_map(&maps_rows[0], rec.p, L, REC_VECS, rpw, 8)) return r;
          if (int r = encode_map(&maps_rows[1], par.p, (uint32_t)PREC(), 1, rpw, 8)) return r;
          rows_ok = true;
        }
        if (split_supported<Real>(pw, L) && kSplitBandRows >= shape.T) {
          if (int r = encode_map(&maps_split[0], rec.p, L, REC_VECS, 1, 8)) return r;
          if (int r = encode_map(&maps_split[1], par.p, (uint32_t)PREC(), 1, 1, 8)) return r;
          split_ok = true;
          CUDA_TRY(lb_rows.alloc(gridH));
        }
      }
    }
    CUDA_TRY(labels.alloc(N));
    CUDA_TRY(cudaMemsetAsync(labels.p, 0xff, labels.bytes(), stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    finalized = true;
    return LPMP_OK;
  }

  int begin() override { if (int r = need_final()) return r; repam = -1; return LPMP_OK; }
  int set_reparametrization(int mode) override {
    if (mode < 0 || mode > 3) { err = "reparametrization mode unknown"; return LPMP_ERR_INVALID; }
    repam = mode;
    return LPMP_OK;
  }
  int set_send_mode(int mode) override {
    if (mode == LPMP_SEND_ADAPTIVE) {
      err = "adaptive sending needs send_message_to_right_improvement of the marginalisation message (absent from the reference snapshot)";
      return LPMP_ERR_UNSUPPORTED;
    }
    if (mode != LPMP_SEND_SHARED && mode != LPMP_SEND_RESIDUAL) { err = "send mode unknown"; return LPMP_ERR_INVALID; }
    send_mode = mode;
    return LPMP_OK;
  }

  // ---- one pass: LP::ComputePass(it...) (include/LP_MP.h:981-1005) / ComputePassAndPrimal (:1592-1602) -------
  int pass(bool forward, bool primal) override {
    if (int r = need_final()) return r;
    if (repam < 0) return bad_state("no reparametrization mode set");  // include/LP_MP.h:458
    CUDA_TRY(cudaSetDevice(device));
    fused_lb_valid = false;
    // dense-table grids: the row-dataflow kernel covers the anisotropic mode; the uniform modes of the rounding
    // iterations (include/visitors/standard_visitor.hxx:44) run on the level schedule, built on first use
    if (!is_grid) {
      bool fused = false;
      const int n = graph.pass(forward, primal, repam, send_mode == LPMP_SEND_RESIDUAL, labels.p, tune.level_group, num_sms, stream, err, &fused);
      if (n < 0) return LPMP_ERR_UNSUPPORTED;
      fused_lb_valid = fused;
      launches += (uint64_t)n;
      CUDA_TRY(cudaGetLastError());
      updates += ((repam == LPMP_REPAM_ANISOTROPIC || repam == LPMP_REPAM_ANISOTROPIC2) ? 2 : 4) * E;
      if (primal) labels_valid = true;
      return LPMP_OK;
    }
    const bool dense_wf = schedule == LPMP_SCHEDULE_WAVEFRONT && pw == LPMP_PAIRWISE_DENSE;
    // the wavefront kernels implement the shared send with the anisotropic / uniform / damped weights; Anisotropic2 weights and
    // residual sends run on the level schedule of the same node records, built on first use
    const bool off_wavefront = repam == LPMP_REPAM_ANISOTROPIC2 || (send_mode == LPMP_SEND_RESIDUAL && !primal);
    if (((dense_wf && repam != LPMP_REPAM_ANISOTROPIC) || (schedule == LPMP_SCHEDULE_WAVEFRONT && off_wavefront)) && !levels_built) {
      if (striped) { err = "stripes support the anisotropic reparametrization with shared sends only"; return LPMP_ERR_UNSUPPORTED; }
      if (int r = build_grid_levels()) return r;
    }
    if (schedule == LPMP_SCHEDULE_LEVELS || (dense_wf && repam != LPMP_REPAM_ANISOTROPIC) || off_wavefront) {
      // lanes per node: the update of a node is a chain of dependent gathers (incidence -> edge -> vectors / table), so the
      // level kernel is bound by the number of nodes in flight; fewer lanes per node = more nodes per SM ("level_group")
      int G = group_size_for(L);
      if (tune.level_group > 0 && tune.level_group < G && (tune.level_group & (tune.level_group - 1)) == 0) G = tune.level_group;
      else if (tune.level_group == 0 && L == 8) G = 4;   // measured on C4 (profiles/README.md): 8 lanes 0.133 ms / launch, 4 lanes 0.123, 2 lanes 0.188
      int threads = 256;
      size_t per_group = 4 * (size_t)L * sizeof(Real);
      while (threads > G && (threads / G) * per_group > 48 * 1024) threads >>= 1;
      if ((threads / G) * per_group > 48 * 1024) { err = "label count too large for the level kernel"; return LPMP_ERR_UNSUPPORTED; }
      const int groups = threads / G;
      const size_t nlev = level_off.size() - 1;
      for (size_t q = 0; q < nlev; ++q) {
        const size_t lv = forward ? q : nlev - 1 - q;
        const uint64_t cnt = level_off[lv + 1] - level_off[lv];
        if (cnt == 0) continue;
        LevelArgs<Real> A;
        A.st = store();
        A.nodes = level_nodes.p + level_off[lv]; A.num_nodes = (uint32_t)cnt;
        A.inc_off = inc_off.p; A.inc_edge = inc_edge.p; A.inc_flags = inc_flags.p; A.labels = labels.p;
        A.L = L; A.pw = pw; A.repam = repam; A.forward = forward ? 1 : 0; A.primal = primal ? 1 : 0; A.residual = send_mode == LPMP_SEND_RESIDUAL ? 1 : 0;
        const int ctas = (int)std::min<uint64_t>((cnt + groups - 1) / groups, (uint64_t)num_sms * 8);
        level_update_kernel<Real><<<ctas, threads, (size_t)groups * per_group, stream>>>(A, G);
        ++launches;
      }
      CUDA_TRY(cudaGetLastError());
    } else if (dense_wf) {
      DenseGridArgs<Real> A;
      A.H = gridH; A.W = gridW;
      A.forward = forward ? 1 : 0; A.primal = primal ? 1 : 0; A.fifo = edge_order == LPMP_EDGE_ORDER_FIFO;
      const bool has_above = row0 > 0, has_below = row0 + gridH < H_total;
      A.in_halo = forward ? has_above : has_below;
      A.out_halo = forward ? has_below : has_above;
      A.control = control.p; A.rec = rec.p; A.tabH = tabH(); A.tabV = tabV(); A.handoff = handoff.p;
      A.in_boundary = forward ? in_top.p : in_bot.p;
      A.out_boundary = forward ? peer_bot : peer_top;
      A.labels = labels.p;
      A.lb_rows = forward ? nullptr : lb_rows.p;   // the backward pass leaves every edge final: LowerBound comes for free
      CUDA_TRY(cudaMemsetAsync(control.p, 0, sizeof(unsigned int), stream));
      CUDA_TRY(launch_dense_pass<Real>(A, L, tune, num_sms, stream));
      ++launches;
      fused_lb_valid = !forward;
    } else {
      const bool uni_mode = repam == LPMP_REPAM_UNIFORM || repam == LPMP_REPAM_DAMPED_UNIFORM;
      if (striped && !uni_mode && repam != LPMP_REPAM_ANISOTROPIC) { err = "stripes support the anisotropic, uniform and damped-uniform reparametrizations"; return LPMP_ERR_UNSUPPORTED; }
      GridArgs<Real> A;
      A.H = gridH; A.W = gridW;
      A.num_bands = (gridH + shape.T - 1) / shape.T;
      A.forward = forward ? 1 : 0; A.primal = primal ? 1 : 0; A.fifo = edge_order == LPMP_EDGE_ORDER_FIFO; A.repam = repam;
      const bool has_above = row0 > 0, has_below = row0 + gridH < H_total;
      A.in_halo = forward ? has_above : has_below;
      A.out_halo = forward ? has_below : has_above;
      A.tl_reach = tl_reach;
      A.lean_flags = tune.lean > 0 ? tune.lean : 0;
      A.potts_nonneg = pw == LPMP_PAIRWISE_POTTS && !potts_neg;
      A.control = control.p; A.handoff = handoff.p;
      A.in_boundary = forward ? in_top.p : in_bot.p;
      A.out_boundary = forward ? peer_bot : peer_top;
      // uniform modes on a stripe: the last virtual row also receives from the neighbouring stripe's first virtual row, whose
      // in-side vectors (state of the previous pass) are delivered into the OTHER in-buffer before the pass (lpmp_grid_push_boundary)
      A.later_boundary = (uni_mode && A.out_halo) ? (forward ? in_bot.p : in_top.p) : nullptr;
      A.labels = labels.p;
      A.lb_rows = nullptr;
      A.rec = rec.p; A.par = par.p;
      A.trace = trace.p;
      A.trace_steps = (trace.p && want_trace >= 2) ? trace.p + 4 * (size_t)A.num_bands : nullptr;
      A.trace_band = A.num_bands / 2;
      CUDA_TRY(cudaMemsetAsync(control.p, 0, sizeof(unsigned int), stream));
      // one label per lane (kernels_grid_rows.cuh): the anisotropic passes of grids whose pass time is the dependency chain
      const int rows_win = pw == LPMP_PAIRWISE_POTTS ? (potts_neg ? 0 : 1) : (tl_irregular || tl_reach > 8 || (tl_reach > 4 && L < 16) ? -1 : std::max(tl_reach, 1));
      if (split_ok && repam == LPMP_REPAM_ANISOTROPIC && rows_win >= 0 && (pw == LPMP_PAIRWISE_POTTS || L == 16) && use_split()) {
        A.num_bands = (gridH + (uint32_t)kSplitBandRows - 1) / (uint32_t)kSplitBandRows;
        // backward passes accumulate LP::LowerBound on the fly where the pass is bandwidth-bound (more bands than SMs: the three
        // extra group minima per step hide under the memory traffic); latency-bound grids keep the separate 0.1 ms sweep
        // (a latency-bound launch pays ~0.05 us per chain step for the three extra group minima -- measured on C2 -- against a sweep
        // that streams every record once)
        const double sweep_us = (double)N * REC() * sizeof(Real) / 6.5e6, chain_us = 0.05 * ((double)gridH + gridW);
        const bool fuse = !forward && !primal && (tune.fused_bound > 0 || (tune.fused_bound < 0 && (A.num_bands > (uint32_t)num_sms || sweep_us > chain_us)));
        if (fuse) A.lb_rows = lb_rows.p;
        const cudaError_t le = pw == LPMP_PAIRWISE_POTTS ? launch_split<Real, LPMP_PAIRWISE_POTTS>(maps_split, A, L, rows_win, tune, num_sms, stream)
                                                         : launch_split<Real, LPMP_PAIRWISE_TRUNCLIN>(maps_split, A, L, rows_win, tune, num_sms, stream);
        CUDA_TRY(le);
        ++launches;
        updates += 2 * E;
        if (primal) labels_valid = true;
        fused_lb_valid = fuse;
        fused_lb_open_below = fuse && has_below;
        return LPMP_OK;
      }
      // the uniform / damped-uniform modes (the rounding iterations of the default control law, include/visitors/standard_visitor.hxx:44)
      // on the split kernel: rows of whole tiles only
      if (split_ok && uni_mode && (!striped || gridH % (uint32_t)kSplitBandRows == 0) && rows_win >= 0 && (pw == LPMP_PAIRWISE_POTTS || L == 16) && use_split() &&
          tune.split_uni != 0 && gridW % 8 == 0 && gridW >= 16 && split_uni_supported<Real>(pw, L)) {
        A.num_bands = (gridH + (uint32_t)kSplitBandRows - 1) / (uint32_t)kSplitBandRows;
        if (A.later_boundary != nullptr) {
          // stripe (whole bands only): the not yet processed neighbouring stripe's vectors go from the in-buffer into the ghost row
          // next to the stripe's last virtual row, where the tile warp loads them like the row below any other band
          Real* const ghost = forward ? rec.p + (uint64_t)gridH * gridW * REC() + (uint64_t)REC_BU * L : rec.p - (uint64_t)gridW * REC() + (uint64_t)REC_AD * L;
          later_gather_kernel<Real><<<grid_for((uint64_t)gridW * L, 256, 148 * 4), 256, 0, stream>>>(ghost, A.later_boundary, gridW, L, REC(), control.p, !forward);
          ++launches;
          CUDA_TRY(cudaGetLastError());
        }
        const cudaError_t le = pw == LPMP_PAIRWISE_POTTS ? launch_split_uni<Real, LPMP_PAIRWISE_POTTS>(maps_split, A, L, rows_win, tune, num_sms, stream)
                                                         : launch_split_uni<Real, LPMP_PAIRWISE_TRUNCLIN>(maps_split, A, L, rows_win, tune, num_sms, stream);
        CUDA_TRY(le);
        ++launches;
        updates += 4 * E;
        if (primal) labels_valid = true;
        return LPMP_OK;
      }
      if (rows_ok && repam == LPMP_REPAM_ANISOTROPIC && rows_win >= 0 && use_rows()) {
        A.num_bands = (gridH + (uint32_t)rows_band_rows(L) - 1) / (uint32_t)rows_band_rows(L);
        const cudaError_t le = pw == LPMP_PAIRWISE_POTTS ? launch_rows<Real, LPMP_PAIRWISE_POTTS>(maps_rows, A, L, rows_win, tune, num_sms, stream)
                                                         : launch_rows<Real, LPMP_PAIRWISE_TRUNCLIN>(maps_rows, A, L, rows_win, tune, num_sms, stream);
        CUDA_TRY(le);
        ++launches;
        updates += 2 * E;
        if (primal) labels_valid = true;
        return LPMP_OK;
      }
      // tile width: 8 columns unless the bands outnumber the resident warps and the narrow tile doubles residency
      int ksel = shape.default_k();
      if (shape.fits8 && shape.fits4) {
        const uint64_t resident8 = (uint64_t)num_sms * std::max<size_t>(1, (227 * 1024) / (shape.smem8 + 1024));
        if (A.num_bands > resident8) ksel = 4;
        if (tune.tile_cols == 4 || tune.tile_cols == 8) ksel = tune.tile_cols;
      }
      CUDA_TRY(launch_wavefront<Real>(ksel == 8 ? maps8 : maps4, A, pw, L, tl_hops(), ksel, tune, num_sms, stream));
      ++launches;
    }
    updates += ((repam == LPMP_REPAM_ANISOTROPIC || repam == LPMP_REPAM_ANISOTROPIC2) ? 2 : 4) * E;
    if (primal) labels_valid = true;
    return LPMP_OK;
  }

  // LP::put_in_same_partition / compute_partition_pass / compute_overlapping_partition_pass (include/LP_MP.h:463-465,1961-2070)
  int put_in_same_partition(uint64_t m, const uint64_t* i, const uint64_t* j) override {
    if (is_grid) { err = "partition passes run on the general-graph layout: build the grid through add_unaries / add_pairwise_*"; return LPMP_ERR_UNSUPPORTED; }
    for (uint64_t k = 0; k < m; ++k) if (i[k] >= N || j[k] >= N) { err = "variable out of range"; return LPMP_ERR_INVALID; }
    for (uint64_t k = 0; k < m; ++k) graph.partition_graph.push_back({i[k], j[k]});
    graph.partition_valid = false;
    return LPMP_OK;
  }
  int partition_pass(uint64_t inner, int overlapping) override {
    if (int r = need_final()) return r;
    if (is_grid) { err = "partition passes run on the general-graph layout: build the grid through add_unaries / add_pairwise_*"; return LPMP_ERR_UNSUPPORTED; }
    CUDA_TRY(cudaSetDevice(device));
    fused_lb_valid = false;
    int launched = 0;
    const int r = graph.partition_pass(inner, overlapping != 0, tune.level_group, num_sms, stream, err, &launched);
    launches += (uint64_t)launched;
    if (r) return r;
    CUDA_TRY(cudaGetLastError());
    return LPMP_OK;
  }

  template <typename Topo>
  BoundArgs<Real, Topo> bound_args(Topo topo) {
    BoundArgs<Real, Topo> A;
    A.st = store(); A.topo = topo;
    A.num_nodes = N; A.L = L; A.pw = pw; A.partial = partial.p;
    return A;
  }

  // the wavefront kernel bounds its hand-off polls; a timeout leaves a sticky flag and an unusable state
  int check_abort() {
    if (!control.p) return LPMP_OK;
    unsigned int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, control.p + 1, sizeof(flag), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (flag) { err = "wavefront pass aborted: a band waited too long for its predecessor (state is invalid)"; return LPMP_ERR_CUDA; }
    return LPMP_OK;
  }

  int reduce_to_host(int ctas, double add, double* out, const double* src = nullptr, const double* extra = nullptr, int m = 0) {
    final_sum_kernel<<<1, 32, 0, stream>>>(src ? src : partial.p, ctas, add, result.p, extra, m);
    ++launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, result.p, sizeof(double), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return check_abort();
  }

  // the same final sum, left in device memory: no host synchronisation (multi-GPU drivers reduce it with NCCL on this stream)
  int reduce_to(int ctas, double add, double* out, bool to_device, const double* src = nullptr, const double* extra = nullptr, int m = 0) {
    if (!to_device) return reduce_to_host(ctas, add, out, src, extra, m);
    final_sum_kernel<<<1, 32, 0, stream>>>(src ? src : partial.p, ctas, add, out, extra, m);
    ++launches;
    CUDA_TRY(cudaGetLastError());
    return LPMP_OK;
  }

  int bound_launch(const Real* below_b, double add, double* out, bool to_device = false) {
    CUDA_TRY(cudaSetDevice(device));
    if (!is_grid && fused_lb_valid) {   // the last (anisotropic backward) pass left the bound terms: one wide deterministic sum
      double* dst = to_device ? out : result.p;
      wide_sum_kernel<<<1, 1024, 0, stream>>>(graph.fused_partial.p, (uint32_t)graph.fused_used, add, dst);
      ++launches;
      CUDA_TRY(cudaGetLastError());
      if (to_device) return LPMP_OK;
      CUDA_TRY(cudaMemcpyAsync(out, result.p, sizeof(double), cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
      return LPMP_OK;
    }
    if (!is_grid) {
      int used = 0;
      const int n = graph.bound(partial.p, &used, tune.level_group, num_sms, stream, err);
      if (n < 0) return LPMP_ERR_UNSUPPORTED;
      launches += (uint64_t)n;
      CUDA_TRY(cudaGetLastError());
      return reduce_to(used, add, out, to_device);
    }
    const int G = group_size_for(L);
    int threads = 256;
    const size_t per_group = 2 * (size_t)L * sizeof(Real);
    while (threads > G && (threads / G) * per_group > 48 * 1024) threads >>= 1;
    const int groups = threads / G;
    if (fused_lb_valid && (!striped || pw != LPMP_PAIRWISE_DENSE)) {
      // the last backward pass left the terms of every grid row in lb_rows; a stripe adds the edges to the stripe below
      if (!(fused_lb_open_below && below_b)) return reduce_to((int)gridH, add, out, to_device, lb_rows.p);
      const int bctas = (int)std::min<uint64_t>(((uint64_t)gridW + groups - 1) / groups, (uint64_t)kBoundCtas);
      boundary_bound_kernel<Real><<<bctas, threads, (size_t)groups * per_group, stream>>>(store(), N + (uint64_t)(gridH - 1) * gridW, gridW, L, pw, below_b, partial.p, G);
      ++launches;
      CUDA_TRY(cudaGetLastError());
      return reduce_to((int)gridH, add, out, to_device, lb_rows.p, partial.p, bctas);
    }
    const uint64_t total = N + slots;
    const int ctas = (int)std::min<uint64_t>((total + groups - 1) / groups, (uint64_t)kBoundCtas);
    if (is_grid && schedule == LPMP_SCHEDULE_WAVEFRONT && pw != LPMP_PAIRWISE_DENSE) {
      // record-streaming sweep (kernels_grid_bound.cuh): every record and parameter is read once
      GridBoundArgs<Real> B{gridH, gridW, rec.p, par.p, below_b, partial.p};
      int used = 0;
      const cudaError_t le = pw == LPMP_PAIRWISE_POTTS ? launch_grid_bound<Real, LPMP_PAIRWISE_POTTS>(B, L, 0, kBoundCtas, &used, stream)
                                                       : launch_grid_bound<Real, LPMP_PAIRWISE_TRUNCLIN>(B, L, tl_hops(), kBoundCtas, &used, stream);
      CUDA_TRY(le);
      ++launches;
      return reduce_to(used, add, out, to_device);
    }
    GridTopo<Real> topo{gridH, gridW, slots, below_b};
    lower_bound_kernel<Real, GridTopo<Real>><<<ctas, threads, (size_t)groups * per_group, stream>>>(bound_args(topo), G);
    ++launches;
    CUDA_TRY(cudaGetLastError());
    return reduce_to(ctas, add, out, to_device);
  }

  int lower_bound(double* out) override {
    if (int r = need_final()) return r;
    if (striped) { err = "lower_bound of a stripe: use lpmp_grid_lower_bound_partial"; return LPMP_ERR_UNSUPPORTED; }
    return bound_launch(nullptr, constant, out);
  }
  // stripe share of LP::LowerBound: own unaries, horizontal and inner vertical edges, and the edges to the stripe
  // below (their a side is this stripe's aD, their b side arrives in `below_b`)
  int grid_lower_bound_partial(const void* below_b, double* out) override {
    if (int r = need_final()) return r;
    if (!is_grid) return bad_state("not a grid");
    if (row0 + gridH < H_total && !below_b) { err = "the b-side vectors of the stripe below are required"; return LPMP_ERR_INVALID; }
    return bound_launch(row0 + gridH < H_total ? static_cast<const Real*>(below_b) : nullptr, row0 == 0 ? constant : 0.0, out);
  }

  int primal_launch(const Real* below_b, const uint32_t* below_labels, double add, double* out, bool to_device) {
    CUDA_TRY(cudaSetDevice(device));
    if (!is_grid) {
      int used = 0;
      launches += (uint64_t)graph.primal(labels.p, partial.p, &used, stream);
      CUDA_TRY(cudaGetLastError());
      return reduce_to(used, add, out, to_device);
    }
    const uint64_t total = N + slots;
    const int ctas = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)kBoundCtas);
    GridTopo<Real> topo{gridH, gridW, slots, below_b, below_labels};
    primal_cost_kernel<Real, GridTopo<Real>><<<ctas, 256, 0, stream>>>(bound_args(topo), labels.p);
    ++launches;
    CUDA_TRY(cudaGetLastError());
    return reduce_to(ctas, add, out, to_device);
  }
  int evaluate_primal(double* out) override {
    if (int r = need_final()) return r;
    if (!labels_valid) { *out = std::numeric_limits<double>::infinity(); return LPMP_OK; }
    if (striped) { err = "evaluate_primal of a stripe: use lpmp_grid_evaluate_primal_partial_device"; return LPMP_ERR_UNSUPPORTED; }
    return primal_launch(nullptr, nullptr, constant, out, false);
  }
  // stripe share of LP::EvaluatePrimal: own unaries and edges plus the edges to the stripe below (their b-side vectors and the
  // labels of the first row below are passed as device pointers); +inf (as the whole LP's cost) while nothing is labelled
  int grid_evaluate_primal_partial_device(const void* below_b, const uint32_t* below_labels, double* d_out) override {
    if (int r = need_final()) return r;
    if (!is_grid) return bad_state("not a grid");
    const bool has_below = row0 + gridH < H_total;
    if (has_below && (!below_b || !below_labels)) { err = "the b-side vectors and labels of the stripe below are required"; return LPMP_ERR_INVALID; }
    if (!labels_valid) {
      const double inf = std::numeric_limits<double>::infinity();
      CUDA_TRY(cudaMemcpyAsync(d_out, &inf, sizeof(double), cudaMemcpyHostToDevice, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
      return LPMP_OK;
    }
    return primal_launch(has_below ? static_cast<const Real*>(below_b) : nullptr, has_below ? below_labels : nullptr, row0 == 0 ? constant : 0.0, d_out, true);
  }
  int grid_lower_bound_partial_device(const void* below_b, double* d_out) override {
    if (int r = need_final()) return r;
    if (!is_grid) return bad_state("not a grid");
    if (row0 + gridH < H_total && !below_b) { err = "the b-side vectors of the stripe below are required"; return LPMP_ERR_INVALID; }
    return bound_launch(row0 + gridH < H_total ? static_cast<const Real*>(below_b) : nullptr, row0 == 0 ? constant : 0.0, d_out, true);
  }
  // per-pass halo through peer memory: own side of a stripe boundary (0: bU of the first row, 1: aD of the last row) into the
  // neighbour's in-buffer (slot protocol, see push_boundary_kernel); enqueued on the handle's stream.  The pass kernels index an
  // in-buffer by the column's position in processing order, so vectors consumed by a BACKWARD pass are stored right to left
  int grid_push_boundary(int which, void* peer_in_buf, int backward) override {
    if (!is_grid || !finalized) return bad_state("grid_push_boundary: need a finalized grid");
    if (!peer_in_buf) { err = "null destination"; return LPMP_ERR_INVALID; }
    CUDA_TRY(cudaSetDevice(device));
    const Real* src = which == 0 ? rec.p + (uint64_t)REC_BU * L : rec.p + (uint64_t)(gridH - 1) * gridW * REC() + (uint64_t)REC_AD * L;
    push_boundary_kernel<Real><<<grid_for((uint64_t)gridW * L, 256, 148 * 4), 256, 0, stream>>>(static_cast<Real*>(peer_in_buf), src, gridW, L, REC(), control.p, backward != 0);
    ++launches;
    CUDA_TRY(cudaGetLastError());
    return LPMP_OK;
  }
  int grid_get_boundary_labels(int which, uint32_t* dst_dev) override {
    if (!is_grid || !finalized) return bad_state("grid_get_boundary_labels: need a finalized grid");
    CUDA_TRY(cudaMemcpyAsync(dst_dev, labels.p + (which == 0 ? 0 : (uint64_t)(gridH - 1) * gridW), (uint64_t)gridW * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
    return LPMP_OK;
  }

  int check_primal_consistency(int* out) override {
    if (int r = need_final()) return r;
    // pairwise factors hold copies of the unary labels written in the same pass (ComputeRightFromLeftPrimal), so the labeling is
    // consistent exactly when every unary is labelled (a rounding pass labels all of them): count the unset / out-of-range labels
    if (!labels_valid) { *out = 0; return LPMP_OK; }
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaMemsetAsync(bad_count.p, 0, sizeof(unsigned int), stream));
    if (mixed()) count_bad_labels_var_kernel<<<grid_for(N, 256), 256, 0, stream>>>(labels.p, graph.lab_i.p, N, bad_count.p);
    else count_bad_labels_kernel<<<grid_for(N, 256), 256, 0, stream>>>(labels.p, N, L, bad_count.p);
    ++launches;
    CUDA_TRY(cudaGetLastError());
    unsigned int bad = 0;
    CUDA_TRY(cudaMemcpyAsync(&bad, bad_count.p, sizeof(bad), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    *out = bad == 0 ? 1 : 0;
    return LPMP_OK;
  }
  int add_to_constant(double x) override { constant += x; return LPMP_OK; }
  int counts(uint64_t* factors, uint64_t* messages) override {
    if (factors) *factors = N + E;
    if (messages) *messages = 2 * E;
    return LPMP_OK;
  }

  int get_labeling(uint32_t* out) override {
    if (int r = need_final()) return r;
    CUDA_TRY(cudaMemcpyAsync(out, labels.p, N * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return check_abort();
  }
  // Solver::RegisterPrimal keeps the best labeling seen (include/solver.hxx:321-340): a device-side copy, ordered on the handle's
  // stream -- the rounding iterations of a solve never wait for a download of N labels
  int keep_labeling() override {
    if (int r = need_final()) return r;
    if (!labels_valid) return bad_state("keep_labeling: no rounding pass has run");
    CUDA_TRY(cudaSetDevice(device));
    if (kept_labels.n < N) CUDA_TRY(kept_labels.alloc(N));
    CUDA_TRY(cudaMemcpyAsync(kept_labels.p, labels.p, N * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
    kept_valid = true;
    return LPMP_OK;
  }
  int get_kept_labeling(uint32_t* out) override {
    if (!kept_valid) return bad_state("get_kept_labeling: nothing kept");
    CUDA_TRY(cudaMemcpyAsync(out, kept_labels.p, N * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return check_abort();
  }
  int download(std::vector<Real>& tmp, const Real* src, uint64_t n) {
    tmp.resize(n);
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), src, n * sizeof(Real), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return LPMP_OK;
  }
  int get_unaries(double* th) override {
    if (!is_grid && !finalized) { for (size_t k = 0; k < h_theta.size(); ++k) th[k] = (double)(Real)h_theta[k]; return LPMP_OK; }
    std::vector<Real> tmp;
    if (is_grid) {
      if (int r = download(tmp, rec.p, N * REC())) return r;
      for (uint64_t i = 0; i < N; ++i)
        for (uint32_t x = 0; x < L; ++x) th[i * L + x] = (double)tmp[i * REC() + x];
      return LPMP_OK;
    }
    if (int r = download(tmp, graph.vec.p, graph.vec.n)) return r;
    for (uint64_t i = 0, k = 0; i < N; ++i)
      for (uint32_t x = 0; x < Ln(i); ++x) th[k++] = (double)tmp[graph.theta_off[i] + x];
    return LPMP_OK;
  }
  // position of edge e's two vectors inside the host copy of the state (insertion order of the edges)
  void grid_edge_vector_offsets(std::vector<uint64_t>& oa, std::vector<uint64_t>& ob) const {
    const uint64_t R = REC();
    for (uint32_t r = 0; r < gridH; ++r)
      for (uint32_t c = 0; c < gridW; ++c) {
        const uint64_t i = (uint64_t)r * gridW + c;
        if (c + 1 < gridW) { oa.push_back(i * R + (uint64_t)REC_AR * L); ob.push_back((i + 1) * R + (uint64_t)REC_BL * L); }
        if (r + 1 < gridH) { oa.push_back(i * R + (uint64_t)REC_AD * L); ob.push_back((i + gridW) * R + (uint64_t)REC_BU * L); }
      }
  }
  int get_messages(double* pa, double* pb) override {
    if (int r = need_final()) return r;
    std::vector<Real> tmp;
    if (is_grid) {
      std::vector<uint64_t> oa, ob;
      grid_edge_vector_offsets(oa, ob);
      if (int r = download(tmp, rec.p, N * REC())) return r;
      for (uint64_t e = 0; e < E; ++e)
        for (uint32_t x = 0; x < L; ++x) {
          if (pa) pa[e * L + x] = (double)tmp[oa[e] + x];
          if (pb) pb[e * L + x] = (double)tmp[ob[e] + x];
        }
      return LPMP_OK;
    }
    if (int r = download(tmp, graph.vec.p, graph.vec.n)) return r;
    for (uint64_t e = 0, ka = 0, kb = 0; e < E; ++e) {
      for (uint32_t x = 0; x < Ln(h_ei[e]); ++x, ++ka) if (pa) pa[ka] = (double)tmp[graph.a_off[e] + x];
      for (uint32_t x = 0; x < Ln(h_ej[e]); ++x, ++kb) if (pb) pb[kb] = (double)tmp[graph.b_off[e] + x];
    }
    return LPMP_OK;
  }
  int set_messages(const double* pa, const double* pb) override {
    if (int r = need_final()) return r;
    fused_lb_valid = false;
    if (is_grid) {
      std::vector<uint64_t> oa, ob;
      grid_edge_vector_offsets(oa, ob);
      std::vector<Real> tmp;
      if (int r = download(tmp, rec.p, N * REC())) return r;
      for (uint64_t e = 0; e < E; ++e)
        for (uint32_t x = 0; x < L; ++x) {
          if (pa) tmp[oa[e] + x] = (Real)pa[e * L + x];
          if (pb) tmp[ob[e] + x] = (Real)pb[e * L + x];
        }
      CUDA_TRY(cudaMemcpyAsync(rec.p, tmp.data(), tmp.size() * sizeof(Real), cudaMemcpyHostToDevice, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
      return LPMP_OK;
    }
    std::vector<Real> tmp;
    if (int r = download(tmp, graph.vec.p, graph.vec.n)) return r;
    for (uint64_t e = 0, ka = 0, kb = 0; e < E; ++e) {
      for (uint32_t x = 0; x < Ln(h_ei[e]); ++x, ++ka) if (pa) tmp[graph.a_off[e] + x] = (Real)pa[ka];
      for (uint32_t x = 0; x < Ln(h_ej[e]); ++x, ++kb) if (pb) tmp[graph.b_off[e] + x] = (Real)pb[kb];
    }
    CUDA_TRY(cudaMemcpyAsync(graph.vec.p, tmp.data(), tmp.size() * sizeof(Real), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return LPMP_OK;
  }
  int set_unaries(const double* th) override {
    if (int r = need_final()) return r;
    fused_lb_valid = false;
    if (is_grid) return set_unaries_from(th);
    std::vector<Real> tmp;
    if (int r = download(tmp, graph.vec.p, graph.vec.n)) return r;
    for (uint64_t i = 0, k = 0; i < N; ++i)
      for (uint32_t x = 0; x < Ln(i); ++x) tmp[graph.theta_off[i] + x] = (Real)th[k++];
    CUDA_TRY(cudaMemcpyAsync(graph.vec.p, tmp.data(), tmp.size() * sizeof(Real), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return LPMP_OK;
  }
  // dual variables in the reference's archive layout (include/factor_archive.hxx:34-48, include/serialization.hxx:226-425): per factor
  // in insertion order its serialize_dual() members as doubles -- unary: theta; pairwise: msg1 (a), msg2 (b)
  uint64_t packed_len(int which) const {   // 0 theta, 1 a, 2 b
    if (!mixed()) return (which == 0 ? N : E) * (uint64_t)L;
    uint64_t n = 0;
    if (which == 0) for (uint64_t i = 0; i < N; ++i) n += h_counts[i];
    else for (uint64_t e = 0; e < E; ++e) n += h_counts[which == 1 ? h_ei[e] : h_ej[e]];
    return n;
  }
  int dual_size(uint64_t* bytes) override {
    if (int r = need_final()) return r;
    *bytes = sizeof(double) * (packed_len(0) + packed_len(1) + packed_len(2));
    return LPMP_OK;
  }
  int dual_archive(void* archive, uint64_t bytes, bool save) override {
    uint64_t want = 0;
    if (int r = dual_size(&want)) return r;
    if (!archive || bytes != want) { err = "archive size does not match this LP"; return LPMP_ERR_INVALID; }
    std::vector<double> th(packed_len(0)), va(packed_len(1)), vb(packed_len(2));
    if (save) {
      if (int r = get_unaries(th.data())) return r;
      if (int r = get_messages(va.data(), vb.data())) return r;
    }
    // packed positions of the factors (grids: edges in insertion order, the order get_messages uses)
    std::vector<uint64_t> to(N + 1, 0), ao(E + 1, 0), bo(E + 1, 0);
    for (uint64_t i = 0; i < N; ++i) to[i + 1] = to[i] + Ln(i);
    for (uint64_t e = 0; e < E; ++e) {
      ao[e + 1] = ao[e] + (mixed() ? Ln(h_ei[e]) : L);
      bo[e + 1] = bo[e] + (mixed() ? Ln(h_ej[e]) : L);
    }
    double* ar = static_cast<double*>(archive);
    auto move = [&](double* state, uint64_t n) { if (save) std::copy(state, state + n, ar); else std::copy(ar, ar + n, state); ar += n; };
    for (const auto& r : runs)
      for (uint64_t k = r[1]; k < r[1] + r[2]; ++k) {
        if (r[0] == 0) move(th.data() + to[k], to[k + 1] - to[k]);
        else { move(va.data() + ao[k], ao[k + 1] - ao[k]); move(vb.data() + bo[k], bo[k + 1] - bo[k]); }
      }
    if (!save) {
      if (int r = set_unaries(th.data())) return r;
      if (int r = set_messages(va.data(), vb.data())) return r;
    }
    return LPMP_OK;
  }
  int stats(lpmp_stats* out) override {
    std::memset(out, 0, sizeof(*out));
    out->num_nodes = N; out->num_edges = E; out->num_labels = L;
    if (schedule == LPMP_SCHEDULE_WAVEFRONT) out->num_levels = (uint64_t)gridW + gridH - 1;
    else if (!is_grid) out->num_levels = graph.level_off.empty() ? 0 : graph.level_off.size() - 1;
    else out->num_levels = level_off.empty() ? 0 : level_off.size() - 1;
    out->kernel_launches = launches;
    out->message_updates = updates;
    out->device_bytes = graph.device_bytes() + theta.bytes() + a.bytes() + b.bytes() + params.bytes() + rec_alloc.bytes() + par.bytes() + tab.bytes() + labels.bytes() + inc_off.bytes() +
                        inc_edge.bytes() + inc_flags.bytes() + level_nodes.bytes() + d_ei.bytes() + d_ej.bytes() + control.bytes() + handoff.bytes() +
                        in_top.bytes() + in_bot.bytes();
    out->algorithmic_bytes_per_pass = (double)sizeof(Real) * (2.0 * (double)N * L + (double)E * ((double)P() + 5.0 * L));
    return LPMP_OK;
  }
  int synchronize() override { CUDA_TRY(cudaStreamSynchronize(stream)); return check_abort(); }
  int get_trace(uint64_t* out, uint64_t n) override {
    if (!trace.p) return bad_state("no trace: set_tuning wavefront_trace=1 before finalize");
    CUDA_TRY(cudaStreamSynchronize(stream));
    CUDA_TRY(cudaMemcpy(out, trace.p, std::min<uint64_t>(n, trace.n) * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return LPMP_OK;
  }
  void* stream_ptr() override { return (void*)stream; }
};

struct Handle {
  ISolver* s = nullptr;
  ~Handle() { delete s; }
};

static inline ISolver* S(lpmp_solver* h) { return reinterpret_cast<Handle*>(h)->s; }
static inline const ISolver* S(const lpmp_solver* h) { return reinterpret_cast<const Handle*>(h)->s; }

// one-shot single-factor helpers (parity test entry points)
template <typename Real>
static int run_minmarginal(int pw, uint64_t n, uint32_t L1, uint32_t L2, const double* p0, const double* p1, const double* msg1,
                           const double* msg2, int side, double* out, double* lb) {
  if (pw != LPMP_PAIRWISE_DENSE && L1 != L2) return LPMP_ERR_INVALID;
  if (n == 0) return LPMP_OK;
  const uint64_t np0 = pw == LPMP_PAIRWISE_DENSE ? n * L1 * L2 : n;
  auto up = [&](DeviceBuffer<Real>& d, const double* src, uint64_t cnt) -> cudaError_t {
    if (!src) return cudaSuccess;
    std::vector<Real> tmp(cnt);
    for (uint64_t k = 0; k < cnt; ++k) tmp[k] = (Real)src[k];
    cudaError_t e = d.alloc(cnt);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(d.p, tmp.data(), cnt * sizeof(Real), cudaMemcpyHostToDevice);
  };
  DeviceBuffer<Real> dp0, dp1, dm1, dm2, dout, dmm, dlb;
  if (up(dp0, p0, np0) != cudaSuccess || up(dp1, p1, n) != cudaSuccess || up(dm1, msg1, n * L1) != cudaSuccess || up(dm2, msg2, n * L2) != cudaSuccess) return LPMP_ERR_CUDA;
  auto run = [&](int sd, std::vector<Real>& host) -> int {
    const uint32_t Lo = sd == 0 ? L1 : L2;
    if (dout.alloc(n * Lo) != cudaSuccess) return LPMP_ERR_CUDA;
    minmarginal_kernel<Real><<<grid_for(n * Lo, 128, 1 << 20), 128>>>(pw, n, L1, L2, dp0.p, dp1.p, dm1.p, dm2.p, sd, dout.p);
    host.resize(n * Lo);
    if (cudaMemcpy(host.data(), dout.p, host.size() * sizeof(Real), cudaMemcpyDeviceToHost) != cudaSuccess) return LPMP_ERR_CUDA;
    return LPMP_OK;
  };
  std::vector<Real> host;
  if (int r = run(side, host)) return r;
  for (size_t k = 0; k < host.size(); ++k) out[k] = (double)host[k];
  if (lb) {  // PairwiseSimplexFactor::LowerBound = min over min_marginal_1
    if (side != 0) { if (int r = run(0, host)) return r; }
    if (dlb.alloc(n) != cudaSuccess) return LPMP_ERR_CUDA;
    row_min_kernel<Real><<<grid_for(n, 128, 1 << 20), 128>>>(n, L1, dout.p, dlb.p, nullptr);
    std::vector<Real> hl(n);
    if (cudaMemcpy(hl.data(), dlb.p, n * sizeof(Real), cudaMemcpyDeviceToHost) != cudaSuccess) return LPMP_ERR_CUDA;
    for (uint64_t k = 0; k < n; ++k) lb[k] = (double)hl[k];
  }
  return cudaGetLastError() == cudaSuccess ? LPMP_OK : LPMP_ERR_CUDA;
}

// n star-shaped node updates through the kernels of the general-graph passes (kernels_graph.cuh): records [theta | own...] followed by
// the other sides, one descriptor per node, one launch
template <typename Real>
static int run_node_update(uint64_t n, uint32_t L, uint32_t deg, int pw, int repam, int forward, const uint8_t* later, const double* theta, const double* own,
                           const double* other, const double* params, double* theta_out, double* own_out, uint32_t* label) {
  if (n == 0) return LPMP_OK;
  const uint32_t P = pairwise_param_count(pw, L);
  const uint64_t rec = (uint64_t)(1 + deg) * L, Q = n * deg;
  if ((n * rec + Q * L) >= (1ull << 32)) return LPMP_ERR_UNSUPPORTED;
  std::vector<Real> h_vec(n * rec + Q * L), h_tab(std::max<uint64_t>(1, Q * P));
  std::vector<uint32_t> h_inc(n + 1), h_oth(std::max<uint64_t>(1, Q)), h_node(n);
  std::vector<uint8_t> h_flags(std::max<uint64_t>(1, Q));
  std::vector<uint4> h_desc(2 * n);
  for (uint64_t k = 0; k < n; ++k) {
    h_inc[k] = (uint32_t)(k * deg); h_node[k] = (uint32_t)k;
    for (uint32_t x = 0; x < L; ++x) h_vec[k * rec + x] = (Real)theta[k * L + x];
    uint32_t m[4] = {0, 0, 0, 0};
    for (uint32_t s = 0; s < deg; ++s) {
      const uint64_t q = k * deg + s;
      for (uint32_t x = 0; x < L; ++x) { h_vec[k * rec + (1 + s) * L + x] = (Real)own[q * L + x]; h_vec[n * rec + q * L + x] = (Real)other[q * L + x]; }
      h_oth[q] = (uint32_t)(n * rec + q * L);
      h_flags[q] = (uint8_t)((later[q] ? 1u : 0u) << 1);
      if (s < 4) m[s] = h_oth[q] | h_flags[q];
    }
    h_desc[2 * k] = make_uint4((uint32_t)(k * deg), (uint32_t)k, deg, 0u);
    h_desc[2 * k + 1] = make_uint4(m[0], m[1], m[2], m[3]);
  }
  h_inc[n] = (uint32_t)Q;
  for (uint64_t t = 0; t < Q * P; ++t) h_tab[t] = (Real)params[t];
  DeviceBuffer<Real> vec, tab;
  DeviceBuffer<uint32_t> inc, oth, node, labels;
  DeviceBuffer<uint8_t> flags;
  DeviceBuffer<uint4> desc;
  auto up = [&](auto& d, const auto& h) -> bool {
    return d.alloc(h.size()) == cudaSuccess && cudaMemcpy(d.p, h.data(), h.size() * sizeof(h[0]), cudaMemcpyHostToDevice) == cudaSuccess;
  };
  if (!up(vec, h_vec) || !up(tab, h_tab) || !up(inc, h_inc) || !up(oth, h_oth) || !up(node, h_node) || !up(flags, h_flags) || !up(desc, h_desc)) return LPMP_ERR_CUDA;
  if (labels.alloc(n) != cudaSuccess) return LPMP_ERR_CUDA;
  GraphArgs<Real> A;
  std::memset(&A, 0, sizeof(A));
  A.vec = vec.p; A.tab = tab.p; A.inc_off = inc.p; A.oth_off = oth.p; A.flags = flags.p; A.oth_node = node.p; A.node_of = node.p; A.desc = desc.p;
  A.labels = labels.p; A.first = 0; A.count = (uint32_t)n; A.L = L; A.P = P; A.pw = pw; A.repam = repam; A.forward = forward ? 1 : 0; A.primal = 1;
  for (int t = 1; t < 8; ++t) A.wtab[t] = (Real)(1.0 / (double)t);
  const bool fast = pw == LPMP_PAIRWISE_DENSE && deg <= 4 && (L == 4 || L == 8 || (L == 16 && sizeof(Real) == 4));
  if (fast) {
    const int ctas = (int)std::min<uint64_t>((n + 256 / L - 1) / (256 / L), 148 * 12);
    if (L == 4) graph_node_kernel<Real, 4, 4, 0><<<ctas, 256>>>(A);
    else if (L == 8) graph_node_kernel<Real, 8, 4, 0><<<ctas, 256>>>(A);
    else if constexpr (sizeof(Real) == 4) graph_node_kernel<Real, 16, 4, 0><<<ctas, 256>>>(A);
  } else {
    const int G = group_size_for(L);
    int threads = 256;
    const size_t per_group = 3 * (size_t)L * sizeof(Real);
    while (threads > G && (threads / G) * per_group > 48 * 1024) threads >>= 1;
    if ((threads / G) * per_group > 48 * 1024) return LPMP_ERR_UNSUPPORTED;
    const uint64_t groups = threads / G;
    graph_node_generic_kernel<Real, 0><<<(int)std::min<uint64_t>((n + groups - 1) / groups, 148 * 8), threads, (threads / G) * per_group>>>(A, G);
  }
  if (cudaMemcpy(h_vec.data(), vec.p, n * rec * sizeof(Real), cudaMemcpyDeviceToHost) != cudaSuccess) return LPMP_ERR_CUDA;
  for (uint64_t k = 0; k < n; ++k) {
    for (uint32_t x = 0; x < L; ++x) theta_out[k * L + x] = (double)h_vec[k * rec + x];
    for (uint32_t s = 0; s < deg; ++s)
      for (uint32_t x = 0; x < L; ++x) own_out[(k * deg + s) * L + x] = (double)h_vec[k * rec + (1 + s) * L + x];
  }
  if (label && cudaMemcpy(label, labels.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return LPMP_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? LPMP_OK : LPMP_ERR_CUDA;
}

template <typename Real>
static int run_unary_min(uint64_t n, uint32_t L, const double* theta, double* lb, uint32_t* arg) {
  if (n == 0) return LPMP_OK;
  std::vector<Real> tmp(n * L);
  for (uint64_t k = 0; k < n * L; ++k) tmp[k] = (Real)theta[k];
  DeviceBuffer<Real> dth, dmn;
  DeviceBuffer<uint32_t> darg;
  if (dth.alloc(n * L) != cudaSuccess || dmn.alloc(n) != cudaSuccess || darg.alloc(n) != cudaSuccess) return LPMP_ERR_CUDA;
  if (cudaMemcpy(dth.p, tmp.data(), tmp.size() * sizeof(Real), cudaMemcpyHostToDevice) != cudaSuccess) return LPMP_ERR_CUDA;
  row_min_kernel<Real><<<grid_for(n, 128, 1 << 20), 128>>>(n, L, dth.p, dmn.p, darg.p);
  std::vector<Real> hm(n);
  if (cudaMemcpy(hm.data(), dmn.p, n * sizeof(Real), cudaMemcpyDeviceToHost) != cudaSuccess) return LPMP_ERR_CUDA;
  if (lb) for (uint64_t k = 0; k < n; ++k) lb[k] = (double)hm[k];
  if (arg && cudaMemcpy(arg, darg.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return LPMP_ERR_CUDA;
  return LPMP_OK;
}

}  // namespace lpmp

using namespace lpmp;

extern "C" {

int lpmp_create(int dtype, int device, lpmp_solver** out) {
  if (!out || (dtype != LPMP_F32 && dtype != LPMP_F64)) return LPMP_ERR_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    std::fprintf(stderr, "lpmp_create: no CUDA device available (there is no CPU fallback)\n");
    return LPMP_ERR_CUDA;
  }
  if (device < 0 || device >= count) return LPMP_ERR_INVALID;
  Handle* h = new Handle();
  int r;
  if (dtype == LPMP_F32) { auto* s = new Solver<float>(device); h->s = s; r = s->init(); }
  else { auto* s = new Solver<double>(device); h->s = s; r = s->init(); }
  if (r != LPMP_OK) { std::fprintf(stderr, "lpmp_create: %s\n", h->s->err.c_str()); delete h; return r; }
  *out = reinterpret_cast<lpmp_solver*>(h);
  return LPMP_OK;
}
int lpmp_destroy(lpmp_solver* h) { delete reinterpret_cast<Handle*>(h); return LPMP_OK; }
const char* lpmp_last_error(const lpmp_solver* h) { return h ? S(h)->err.c_str() : "null handle"; }

int lpmp_add_unaries(lpmp_solver* h, uint64_t n, uint32_t L, const double* theta) { return S(h)->add_unaries(n, L, theta); }
int lpmp_add_unaries_var(lpmp_solver* h, uint64_t n, const uint32_t* L, const double* theta) { return S(h)->add_unaries_var(n, L, theta); }
int lpmp_dual_size(lpmp_solver* h, uint64_t* bytes) { return S(h)->dual_size(bytes); }
int lpmp_save_dual(lpmp_solver* h, void* archive, uint64_t bytes) { return S(h)->dual_archive(archive, bytes, true); }
int lpmp_load_dual(lpmp_solver* h, const void* archive, uint64_t bytes) { return S(h)->dual_archive(const_cast<void*>(archive), bytes, false); }
int lpmp_add_pairwise_dense(lpmp_solver* h, uint64_t m, const uint64_t* i, const uint64_t* j, const double* t) { return S(h)->add_pairwise(LPMP_PAIRWISE_DENSE, m, i, j, t, nullptr); }
int lpmp_add_pairwise_potts(lpmp_solver* h, uint64_t m, const uint64_t* i, const uint64_t* j, const double* l) { return S(h)->add_pairwise(LPMP_PAIRWISE_POTTS, m, i, j, l, nullptr); }
int lpmp_add_pairwise_trunclin(lpmp_solver* h, uint64_t m, const uint64_t* i, const uint64_t* j, const double* s, const double* t) { return S(h)->add_pairwise(LPMP_PAIRWISE_TRUNCLIN, m, i, j, s, t); }
int lpmp_set_node_ranks(lpmp_solver* h, const uint64_t* rank) { return S(h)->set_node_ranks(rank); }
int lpmp_set_edge_order(lpmp_solver* h, int eo) { return S(h)->set_edge_order(eo); }
int lpmp_set_schedule(lpmp_solver* h, int s) { return S(h)->set_schedule(s); }
int lpmp_set_tuning(lpmp_solver* h, const char* key, int value) { return S(h)->set_tuning(key, value); }

int lpmp_greedy_colouring(uint64_t n, uint64_t m, const uint64_t* i, const uint64_t* j, uint32_t* colour, uint32_t* num_colours) {
  std::vector<uint64_t> off(n + 1, 0);
  for (uint64_t e = 0; e < m; ++e) {
    if (i[e] >= n || j[e] >= n) return LPMP_ERR_INVALID;
    ++off[i[e] + 1]; ++off[j[e] + 1];
  }
  for (uint64_t v = 0; v < n; ++v) off[v + 1] += off[v];
  std::vector<uint64_t> adj(2 * m), pos(off.begin(), off.end() - 1);
  for (uint64_t e = 0; e < m; ++e) { adj[pos[i[e]]++] = j[e]; adj[pos[j[e]]++] = i[e]; }
  uint32_t nc = 0;
  std::vector<uint32_t> used;
  for (uint64_t v = 0; v < n; ++v) {
    used.clear();
    for (uint64_t q = off[v]; q < off[v + 1]; ++q) if (adj[q] < v) used.push_back(colour[adj[q]]);
    std::sort(used.begin(), used.end());
    uint32_t c = 0;
    for (uint32_t x : used) { if (x == c) ++c; else if (x > c) break; }
    colour[v] = c;
    nc = std::max(nc, c + 1);
  }
  if (num_colours) *num_colours = nc;
  return LPMP_OK;
}

int lpmp_grid_create(lpmp_solver* h, uint32_t H, uint32_t W, uint32_t L, int type) { return S(h)->grid_create(H, W, L, type); }
int lpmp_grid_set_unaries(lpmp_solver* h, const void* theta, int src) { return S(h)->grid_set_unaries(theta, src); }
int lpmp_grid_set_pairwise(lpmp_solver* h, const void* hz, const void* vt, int src) { return S(h)->grid_set_pairwise(hz, vt, src); }
int lpmp_grid_fill_synthetic(lpmp_solver* h, uint64_t seed, double lo, double hi) { return S(h)->grid_fill_synthetic(seed, lo, hi); }
int lpmp_grid_set_stripe(lpmp_solver* h, uint32_t row0, uint32_t H_total) { return S(h)->grid_set_stripe(row0, H_total); }
int lpmp_grid_set_boundary_pairwise(lpmp_solver* h, const void* top, const void* bottom, int src) { return S(h)->grid_set_boundary_pairwise(top, bottom, src); }
int lpmp_grid_boundary_ptrs(lpmp_solver* h, int which, void** in_buf, uint64_t* bytes) { return S(h)->grid_boundary_ptrs(which, in_buf, bytes); }
int lpmp_grid_set_peer_boundary(lpmp_solver* h, int which, void* peer_in_buf) { return S(h)->grid_set_peer_boundary(which, peer_in_buf); }
int lpmp_grid_get_boundary_vectors(lpmp_solver* h, int which, void* dst_dev) { return S(h)->grid_get_boundary_vectors(which, dst_dev); }
int lpmp_grid_lower_bound_partial(lpmp_solver* h, const void* below_b, double* out) { return S(h)->grid_lower_bound_partial(below_b, out); }
int lpmp_grid_lower_bound_partial_device(lpmp_solver* h, const void* below_b, double* d_out) { return S(h)->grid_lower_bound_partial_device(below_b, d_out); }
int lpmp_grid_evaluate_primal_partial_device(lpmp_solver* h, const void* below_b, const uint32_t* below_labels, double* d_out) { return S(h)->grid_evaluate_primal_partial_device(below_b, below_labels, d_out); }
int lpmp_grid_push_boundary(lpmp_solver* h, int which, void* peer_in_buf, int backward) { return S(h)->grid_push_boundary(which, peer_in_buf, backward); }
int lpmp_grid_get_boundary_labels(lpmp_solver* h, int which, uint32_t* dst_dev) { return S(h)->grid_get_boundary_labels(which, dst_dev); }

int lpmp_ipc_export(const void* devptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are 64 bytes");
  cudaIpcMemHandle_t hd;
  if (cudaIpcGetMemHandle(&hd, const_cast<void*>(devptr)) != cudaSuccess) { cudaGetLastError(); return LPMP_ERR_CUDA; }
  std::memcpy(handle64, &hd, 64);
  return LPMP_OK;
}
int lpmp_ipc_open(const void* handle64, void** devptr) {
  cudaIpcMemHandle_t hd;
  std::memcpy(&hd, handle64, 64);
  if (cudaIpcOpenMemHandle(devptr, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return LPMP_ERR_CUDA; }
  return LPMP_OK;
}
int lpmp_ipc_close(void* devptr) { return cudaIpcCloseMemHandle(devptr) == cudaSuccess ? LPMP_OK : LPMP_ERR_CUDA; }

int lpmp_finalize(lpmp_solver* h) { return S(h)->finalize(); }
int lpmp_begin(lpmp_solver* h) { return S(h)->begin(); }
int lpmp_end(lpmp_solver* h) { return S(h)->synchronize(); }
int lpmp_set_reparametrization(lpmp_solver* h, int mode) { return S(h)->set_reparametrization(mode); }
int lpmp_set_send_mode(lpmp_solver* h, int mode) { return S(h)->set_send_mode(mode); }
int lpmp_compute_pass(lpmp_solver* h, uint64_t) { if (int r = S(h)->pass(true, false)) return r; return S(h)->pass(false, false); }
int lpmp_compute_forward_pass(lpmp_solver* h) { return S(h)->pass(true, false); }
int lpmp_compute_backward_pass(lpmp_solver* h) { return S(h)->pass(false, false); }
int lpmp_compute_forward_pass_and_primal(lpmp_solver* h, uint64_t) { return S(h)->pass(true, true); }
int lpmp_compute_backward_pass_and_primal(lpmp_solver* h, uint64_t) { return S(h)->pass(false, true); }
int lpmp_put_in_same_partition(lpmp_solver* h, uint64_t m, const uint64_t* i, const uint64_t* j) { return S(h)->put_in_same_partition(m, i, j); }
int lpmp_compute_partition_pass(lpmp_solver* h, uint64_t inner, int overlapping) { return S(h)->partition_pass(inner, overlapping); }
int lpmp_lower_bound(lpmp_solver* h, double* out) { return S(h)->lower_bound(out); }
int lpmp_evaluate_primal(lpmp_solver* h, double* out) { return S(h)->evaluate_primal(out); }
int lpmp_check_primal_consistency(lpmp_solver* h, int* out) { return S(h)->check_primal_consistency(out); }
int lpmp_add_to_constant(lpmp_solver* h, double x) { return S(h)->add_to_constant(x); }
int lpmp_num_factors(const lpmp_solver* h, uint64_t* out) { return const_cast<ISolver*>(S(h))->counts(out, nullptr); }
int lpmp_num_messages(const lpmp_solver* h, uint64_t* out) { return const_cast<ISolver*>(S(h))->counts(nullptr, out); }
int lpmp_get_labeling(lpmp_solver* h, uint32_t* out) { return S(h)->get_labeling(out); }
int lpmp_keep_labeling(lpmp_solver* h) { return S(h)->keep_labeling(); }
int lpmp_get_kept_labeling(lpmp_solver* h, uint32_t* out) { return S(h)->get_kept_labeling(out); }
int lpmp_get_unaries(lpmp_solver* h, double* th) { return S(h)->get_unaries(th); }
int lpmp_get_messages(lpmp_solver* h, double* a, double* b) { return S(h)->get_messages(a, b); }
int lpmp_set_messages(lpmp_solver* h, const double* a, const double* b) { return S(h)->set_messages(a, b); }
int lpmp_set_unaries(lpmp_solver* h, const double* th) { return S(h)->set_unaries(th); }
int lpmp_get_stats(lpmp_solver* h, lpmp_stats* out) { return S(h)->stats(out); }
int lpmp_synchronize(lpmp_solver* h) { return S(h)->synchronize(); }
int lpmp_get_trace(lpmp_solver* h, uint64_t* out, uint64_t n) { return S(h)->get_trace(out, n); }
int lpmp_stream(lpmp_solver* h, void** out) { if (out) *out = S(h)->stream_ptr(); return LPMP_OK; }

int lpmp_minmarginal_dense(int dtype, uint64_t n, uint32_t L1, uint32_t L2, const double* t, const double* msg1, const double* msg2, int side, double* out, double* lb) {
  return dtype == LPMP_F32 ? run_minmarginal<float>(LPMP_PAIRWISE_DENSE, n, L1, L2, t, nullptr, msg1, msg2, side, out, lb)
                           : run_minmarginal<double>(LPMP_PAIRWISE_DENSE, n, L1, L2, t, nullptr, msg1, msg2, side, out, lb);
}
int lpmp_minmarginal_potts(int dtype, uint64_t n, uint32_t L, const double* lambda, const double* msg1, const double* msg2, int side, double* out, double* lb) {
  return dtype == LPMP_F32 ? run_minmarginal<float>(LPMP_PAIRWISE_POTTS, n, L, L, lambda, nullptr, msg1, msg2, side, out, lb)
                           : run_minmarginal<double>(LPMP_PAIRWISE_POTTS, n, L, L, lambda, nullptr, msg1, msg2, side, out, lb);
}
int lpmp_minmarginal_trunclin(int dtype, uint64_t n, uint32_t L, const double* slope, const double* trunc, const double* msg1, const double* msg2, int side, double* out, double* lb) {
  return dtype == LPMP_F32 ? run_minmarginal<float>(LPMP_PAIRWISE_TRUNCLIN, n, L, L, slope, trunc, msg1, msg2, side, out, lb)
                           : run_minmarginal<double>(LPMP_PAIRWISE_TRUNCLIN, n, L, L, slope, trunc, msg1, msg2, side, out, lb);
}
int lpmp_node_update(int dtype, uint64_t n, uint32_t L, uint32_t deg, int pw, int repam, int forward, const uint8_t* later, const double* theta,
                     const double* own, const double* other, const double* params, double* theta_out, double* own_out, uint32_t* label) {
  if (pw < 0 || pw > 2 || repam < 0 || repam > 3 || L == 0 || (dtype != LPMP_F32 && dtype != LPMP_F64)) return LPMP_ERR_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return LPMP_ERR_CUDA;
  return dtype == LPMP_F32 ? run_node_update<float>(n, L, deg, pw, repam, forward, later, theta, own, other, params, theta_out, own_out, label)
                           : run_node_update<double>(n, L, deg, pw, repam, forward, later, theta, own, other, params, theta_out, own_out, label);
}
int lpmp_unary_min(int dtype, uint64_t n, uint32_t L, const double* theta, double* lb, uint32_t* arg) {
  return dtype == LPMP_F32 ? run_unary_min<float>(n, L, theta, lb, arg) : run_unary_min<double>(n, L, theta, lb, arg);
}

}  // extern "C"
