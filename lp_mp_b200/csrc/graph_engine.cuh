// Host side of the general-graph path: builds the node-record layout of kernels_graph.cuh from the factors added through the C ABI
// (LP::SortFactors + the weight / mask computation of /root/reference/include/LP_MP.h:730-797,1230-1449 realised as level sets and
// per-incidence flags), launches the per-level kernels and converts between the layout and the ABI's per-factor arrays.
#pragma once
#include <algorithm>
#include <numeric>
#include <array>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "kernels_graph.cuh"

#ifdef LPMP_BUILD_TIMING
#include <chrono>
#include <cstdio>
#define LPMP_TICK(name) do { auto _t = std::chrono::steady_clock::now(); std::printf("  build %-10s %.3f s\n", name, std::chrono::duration<double>(_t - _tick).count()); _tick = _t; } while (0)
#else
#define LPMP_TICK(name) do { } while (0)
#endif

namespace lpmp {

#define LPMP_G_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(_e); return LPMP_ERR_CUDA; } } while (0)

// construction on the device (one label count per LP): receiver-oriented parameters of incidence q from the edge-ordered host
// arrays (dense: transposed for the second variable), and theta into the node records
template <typename Real>
__global__ void graph_fill_tables_kernel(Real* __restrict__ tab, const double* __restrict__ params, const uint32_t* __restrict__ inc_es, uint64_t Q, uint32_t L,
                                         uint32_t P, int dense) {
  const uint64_t total = Q * P;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t q = k / P;
    const uint32_t r = (uint32_t)(k % P), es = inc_es[q];
    const uint64_t e = es >> 1;
    const uint32_t src = (dense && (es & 1u)) ? (r % L) * L + r / L : r;
    tab[k] = (Real)params[e * P + src];
  }
}
template <typename Real>
__global__ void graph_fill_theta_kernel(Real* __restrict__ vec, const double* __restrict__ theta, const uint32_t* __restrict__ inc_off,
                                        const uint32_t* __restrict__ node_of, uint64_t N, uint32_t L) {
  const uint64_t total = N * L;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = t / L;
    const uint32_t x = (uint32_t)(t % L);
    vec[(k + inc_off[k]) * L + x] = (Real)theta[(uint64_t)node_of[k] * L + x];
  }
}

// host loops over independent nodes, split over a few threads (layout construction of large graphs)
template <typename F>
inline void parallel_blocks(uint64_t n, bool parallel, F f) {
  const unsigned hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
  if (!parallel || n < (1u << 16) || hw == 1) { f(uint64_t(0), n); return; }
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < hw; ++t) pool.emplace_back([=] { f(n * t / hw, n * (t + 1) / hw); });
  for (auto& th : pool) th.join();
}

template <typename Real, template <typename> class Buffer>
struct GraphEngine {
  static constexpr int kDmax = 4;       // fast kernel: at most this many incident edges
  static constexpr int kPartials = 148 * 32;
  // device
  Buffer<Real> vec, tab;
  Buffer<uint32_t> inc_off, oth_off, oth_node, node_of, voff, lab_k, lab_oth, lab_i;
  Buffer<uint4> desc;
  Buffer<uint64_t> toff;
  Buffer<uint8_t> flags;
  // host
  uint64_t N = 0, E = 0;
  uint32_t L = 0, P = 0;
  int pw = 0;
  bool mixed = false;
  std::vector<uint32_t> Lv;                 // label count per node (original index)
  std::vector<uint64_t> theta_off;          // [N] offset of theta_i in vec
  std::vector<uint64_t> a_off, b_off;       // [E] offsets of a_e / b_e in vec
  std::vector<uint64_t> level_off;          // nodes of level l: [level_off[l], level_off[l+1]) in schedule order
  std::vector<uint64_t> level_fast;         // ... of which [level_off[l], level_fast[l]) qualify for the fast kernel
  bool fast_ok = false;
  // the fast part of a level, sorted by (n_in, n_out): ranges of nodes with the same number of earlier / later neighbours
  struct Segment { uint64_t first, count; uint32_t n_in, n_out; };
  std::vector<std::vector<Segment>> segments;
  // staged optimisation (partition passes): host copies of the incidence structure, the partitions and the planned sweeps
  std::vector<uint32_t> k_of, h_inc, h_oth;             // schedule position of node i; inc_off / oth_node by schedule position
  std::vector<uint64_t> rank_order;                     // nodes in rank order
  std::vector<std::array<uint64_t, 2>> partition_graph;
  std::vector<std::vector<uint64_t>> parts;
  bool partition_valid = false;
  struct Plan {
    Buffer<uint32_t> node, base;
    Buffer<uint8_t> code;
    Buffer<Real> w;
    std::vector<uint64_t> level_off;
  };
  std::map<std::pair<int, size_t>, Plan> plans;         // (kind, partition) -> sweep; kinds: 0 forward, 1 backward, 2 push to i+1, 3 push to i-1

  uint32_t Ln(uint64_t i) const { return mixed ? Lv[i] : L; }
  size_t device_bytes() const {
    return vec.bytes() + tab.bytes() + inc_off.bytes() + oth_off.bytes() + oth_node.bytes() + node_of.bytes() + voff.bytes() + lab_k.bytes() + lab_oth.bytes() +
           lab_i.bytes() + toff.bytes() + flags.bytes() + desc.bytes();
  }

  template <typename T>
  cudaError_t put(Buffer<T>& d, const std::vector<T>& h, cudaStream_t stream) {
    cudaError_t e = d.alloc(std::max<size_t>(1, h.size()));
    if (e != cudaSuccess || h.empty()) return e;
    return cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, stream);
  }

  // theta: packed doubles; params: per edge packed (dense: L_i x L_j, x_i major; Potts: 1; truncated linear: 2)
  int build(uint64_t N_, uint32_t L_, const std::vector<uint32_t>& counts, const std::vector<double>& theta, const std::vector<uint64_t>& ei,
            const std::vector<uint64_t>& ej, const std::vector<double>& params, int pw_, const std::vector<uint64_t>& rank_in, bool lifo, cudaStream_t stream,
            std::string& err) {
#ifdef LPMP_BUILD_TIMING
    auto _tick = std::chrono::steady_clock::now();
#endif
    N = N_; E = ei.size(); L = L_; pw = pw_;
    mixed = !counts.empty();
    Lv = counts;
    P = pairwise_param_count(pw, L);
    std::vector<uint64_t> rank = rank_in;
    if (rank.empty()) { rank.resize(N); std::iota(rank.begin(), rank.end(), uint64_t(0)); }
    // nodes in rank order (the default rank is the node index: nothing to sort)
    std::vector<uint64_t> order(N);
    std::iota(order.begin(), order.end(), uint64_t(0));
    if (!rank_in.empty()) {
      std::sort(order.begin(), order.end(), [&](uint64_t x, uint64_t y) { return rank[x] < rank[y]; });
      for (uint64_t t = 1; t < N; ++t) if (rank[order[t]] == rank[order[t - 1]]) { err = "node ranks must be distinct"; return LPMP_ERR_INVALID; }
    }
    LPMP_TICK("init");
    if (2 * E >= (1ull << 32)) { err = "more than 2^31 edges"; return LPMP_ERR_UNSUPPORTED; }
    // per-node message order: first-variable message type, then second-variable type (include/factors_messages.hxx:3402-3419);
    // inside a type LIFO (:2030-2041) or FIFO (:1772-1930)
    std::vector<uint64_t> cnt_first(N, 0), cnt_second(N, 0), off(N + 1, 0);
    for (uint64_t e = 0; e < E; ++e) { ++cnt_first[ei[e]]; ++cnt_second[ej[e]]; }
    for (uint64_t i = 0; i < N; ++i) off[i + 1] = off[i] + cnt_first[i] + cnt_second[i];
    struct Inc { uint32_t edge; uint8_t side, later; };
    std::vector<Inc> inc(2 * E);
    {
      std::vector<uint64_t> pf(N, 0), ps(N, 0);
      for (uint64_t e = 0; e < E; ++e) {
        const uint64_t i = ei[e], j = ej[e];
        const uint64_t ki = off[i] + (lifo ? cnt_first[i] - 1 - pf[i] : pf[i]);
        const uint64_t kj = off[j] + cnt_first[j] + (lifo ? cnt_second[j] - 1 - ps[j] : ps[j]);
        ++pf[i]; ++ps[j];
        inc[ki] = {(uint32_t)e, 0, (uint8_t)(rank[j] > rank[i])};
        inc[kj] = {(uint32_t)e, 1, (uint8_t)(rank[i] > rank[j])};
      }
    }
    LPMP_TICK("incidence");
    // levels = longest path from the sources of the DAG "lower rank -> higher rank" (any linear extension of the relation DAG gives
    // the same state; nodes of one level have no edge between them)
    std::vector<uint32_t> level(N, 0);
    uint32_t max_level = 0;
    for (uint64_t t = 0; t < N; ++t) {
      const uint64_t i = order[t];
      uint32_t lv = 0;
      for (uint64_t q = off[i]; q < off[i + 1]; ++q) {
        const uint64_t o = inc[q].side == 0 ? ej[inc[q].edge] : ei[inc[q].edge];
        if (!inc[q].later) lv = std::max(lv, level[o] + 1);
      }
      level[i] = lv;
      max_level = std::max(max_level, lv);
    }
    LPMP_TICK("levels");
    fast_ok = !mixed && pw == LPMP_PAIRWISE_DENSE && (L == 4 || L == 8 || (L == 16 && sizeof(Real) == 4));
    // schedule order: level by level; inside a level the nodes of the fast kernel first, sorted by their (earlier, later) neighbour
    // counts (the register kernel is launched per range with the instance sized for its receives), each range in rank order
    auto is_fast = [&](uint64_t i) { return fast_ok && off[i + 1] - off[i] <= (uint64_t)kDmax; };
    auto bucket_of = [&](uint64_t i) {   // n_in * (kDmax + 1) + n_out
      uint32_t n_in = 0, n_out = 0;
      for (uint64_t q = off[i]; q < off[i + 1]; ++q) { if (inc[q].later) ++n_out; else ++n_in; }
      return n_in * (uint32_t)(kDmax + 1) + n_out;
    };
    constexpr uint32_t kBuckets = (kDmax + 1) * (kDmax + 1);
    level_off.assign((size_t)max_level + 2, 0);
    level_fast.assign((size_t)max_level + 1, 0);
    for (uint64_t i = 0; i < N; ++i) { ++level_off[level[i] + 1]; if (is_fast(i)) ++level_fast[level[i]]; }
    for (size_t l = 0; l + 1 < level_off.size(); ++l) level_off[l + 1] += level_off[l];
    std::vector<uint32_t> node_of_k(N);
    k_of.assign(N, 0);
    {
      // counting sort of the fast nodes of every level by bucket (stable in rank order)
      std::vector<uint64_t> cnt((size_t)(max_level + 1) * kBuckets, 0), pos_slow(max_level + 1);
      for (uint64_t i = 0; i < N; ++i) if (is_fast(i)) ++cnt[(size_t)level[i] * kBuckets + bucket_of(i)];
      segments.assign((size_t)max_level + 1, {});
      for (size_t l = 0; l <= max_level; ++l) {
        uint64_t p = level_off[l];
        for (uint32_t b = 0; b < kBuckets; ++b) {
          const uint64_t c = cnt[l * kBuckets + b];
          cnt[l * kBuckets + b] = p;   // becomes the write position of the bucket
          if (c) segments[l].push_back({p, c, b / (uint32_t)(kDmax + 1), b % (uint32_t)(kDmax + 1)});
          p += c;
        }
        pos_slow[l] = level_off[l] + level_fast[l]; level_fast[l] += level_off[l];
      }
      for (uint64_t t = 0; t < N; ++t) {
        const uint64_t i = order[t];
        const uint64_t k = is_fast(i) ? cnt[(size_t)level[i] * kBuckets + bucket_of(i)]++ : pos_slow[level[i]]++;
        node_of_k[k] = (uint32_t)i; k_of[i] = (uint32_t)k;
      }
    }
    rank_order = order;
    LPMP_TICK("order");
    // records
    std::vector<uint32_t> h_inc_off(N + 1, 0), h_voff, h_lab_k;
    std::vector<uint64_t> rec_off(N + 1, 0);
    for (uint64_t k = 0; k < N; ++k) {
      const uint64_t i = node_of_k[k], deg = off[i + 1] - off[i];
      h_inc_off[k + 1] = h_inc_off[k] + (uint32_t)deg;
      rec_off[k + 1] = rec_off[k] + (1 + deg) * Ln(i);
    }
    if (rec_off[N] >= (1ull << 32)) { err = "more than 2^32 reals of node records"; return LPMP_ERR_UNSUPPORTED; }
    LPMP_TICK("recoff");
    const uint64_t Q = 2 * E;
    theta_off.resize(N); a_off.resize(E); b_off.resize(E);
    std::vector<Real> h_vec(mixed ? rec_off[N] : 0, Real(0));   // (one label count: filled on the device)
    std::vector<uint8_t> h_flags(Q);
    std::vector<uint32_t> h_oth_node(Q), h_oth_off(Q), h_lab_oth, h_inc_es(mixed ? 0 : Q);
    std::vector<uint64_t> h_toff(Q + 1, 0), theta_src(N + 1, 0), par_src(E + 1, 0);
    for (uint64_t i = 0; i < N; ++i) theta_src[i + 1] = theta_src[i] + Ln(i);
    for (uint64_t e = 0; e < E; ++e)
      par_src[e + 1] = par_src[e] + (pw == LPMP_PAIRWISE_DENSE ? (uint64_t)Ln(ei[e]) * Ln(ej[e]) : (pw == LPMP_PAIRWISE_POTTS ? 1 : 2));
    if (theta_src[N] != theta.size() || par_src[E] != params.size()) { err = "internal: packed sizes do not match the factors"; return LPMP_ERR_INVALID; }
    parallel_blocks(N, true, [&](uint64_t k_begin, uint64_t k_end) {
    for (uint64_t k = k_begin; k < k_end; ++k) {
      const uint64_t i = node_of_k[k];
      theta_off[i] = rec_off[k];
      if (mixed) for (uint32_t x = 0; x < Ln(i); ++x) h_vec[rec_off[k] + x] = (Real)theta[theta_src[i] + x];
      for (uint64_t s = 0; s < off[i + 1] - off[i]; ++s) {
        const Inc& c = inc[off[i] + s];
        (c.side == 0 ? a_off : b_off)[c.edge] = rec_off[k] + (1 + s) * Ln(i);
      }
    }
    });
    // (the running sum of the parameter offsets is only needed -- and only sequential -- with mixed label counts)
    parallel_blocks(N, !mixed, [&](uint64_t k_begin, uint64_t k_end) {
    for (uint64_t k = k_begin; k < k_end; ++k) {
      const uint64_t i = node_of_k[k];
      for (uint64_t s = 0; s < off[i + 1] - off[i]; ++s) {
        const Inc& c = inc[off[i] + s];
        const uint64_t q = h_inc_off[k] + s, o = c.side == 0 ? ej[c.edge] : ei[c.edge];
        h_flags[q] = (uint8_t)(c.side | (c.later << 1));
        if (!mixed) h_inc_es[q] = (c.edge << 1) | c.side;
        h_oth_node[q] = (uint32_t)o;
        h_oth_off[q] = (uint32_t)(c.side == 0 ? b_off : a_off)[c.edge];
        if (mixed) h_toff[q + 1] = h_toff[q] + (pw == LPMP_PAIRWISE_DENSE ? (uint64_t)Ln(i) * Ln(o) : (pw == LPMP_PAIRWISE_POTTS ? 1 : 2));
      }
    }
    });
    LPMP_TICK("vec+meta");
    std::vector<Real> h_tab(mixed ? h_toff[Q] : 0);
    for (uint64_t k = 0; mixed && k < N; ++k) {
      const uint64_t i = node_of_k[k];
      for (uint64_t s = 0; s < off[i + 1] - off[i]; ++s) {
        const Inc& c = inc[off[i] + s];
        const uint64_t q = h_inc_off[k] + s;
        const double* src = params.data() + par_src[c.edge];
        Real* dst = h_tab.data() + h_toff[q];
        if (pw != LPMP_PAIRWISE_DENSE) { for (uint64_t t = 0; t < h_toff[q + 1] - h_toff[q]; ++t) dst[t] = (Real)src[t]; continue; }
        const uint32_t Li = Ln(ei[c.edge]), Lj = Ln(ej[c.edge]);
        if (c.side == 0) for (uint64_t t = 0; t < (uint64_t)Li * Lj; ++t) dst[t] = (Real)src[t];
        else for (uint32_t y = 0; y < Lj; ++y) for (uint32_t x = 0; x < Li; ++x) dst[(uint64_t)y * Li + x] = (Real)src[(uint64_t)x * Lj + y];
      }
    }
    LPMP_TICK("tab");
    LPMP_G_TRY(put(inc_off, h_inc_off, stream));
    LPMP_G_TRY(put(oth_off, h_oth_off, stream));
    LPMP_G_TRY(put(oth_node, h_oth_node, stream));
    LPMP_G_TRY(put(node_of, node_of_k, stream));
    LPMP_G_TRY(put(flags, h_flags, stream));
    if (mixed) {
      LPMP_G_TRY(put(vec, h_vec, stream));
      LPMP_G_TRY(put(tab, h_tab, stream));
    } else {
      Buffer<double> d_src;
      Buffer<uint32_t> d_es;
      LPMP_G_TRY(vec.alloc(std::max<uint64_t>(1, rec_off[N])));
      LPMP_G_TRY(cudaMemsetAsync(vec.p, 0, vec.bytes(), stream));
      LPMP_G_TRY(put(d_src, theta, stream));
      graph_fill_theta_kernel<Real><<<(int)std::min<uint64_t>((N * L + 255) / 256, 148 * 16), 256, 0, stream>>>(vec.p, d_src.p, inc_off.p, node_of.p, N, L);
      LPMP_G_TRY(cudaStreamSynchronize(stream));   // (d_src is reused)
      LPMP_G_TRY(tab.alloc(std::max<uint64_t>(1, Q * P)));
      if (Q) {
        LPMP_G_TRY(put(d_src, params, stream));
        LPMP_G_TRY(put(d_es, h_inc_es, stream));
        graph_fill_tables_kernel<Real><<<(int)std::min<uint64_t>((Q * P + 255) / 256, 148 * 32), 256, 0, stream>>>(tab.p, d_src.p, d_es.p, Q, L, P, pw == LPMP_PAIRWISE_DENSE ? 1 : 0);
      }
      LPMP_G_TRY(cudaGetLastError());
      LPMP_G_TRY(cudaStreamSynchronize(stream));   // the temporaries go out of scope
    }
    if (fast_ok) {   // node descriptors of the fast kernel (the vector offsets are multiples of L >= 4: two low bits carry the flags)
      std::vector<uint4> h_desc(2 * N);
      for (uint64_t k = 0; k < N; ++k) {
        const uint32_t q0 = h_inc_off[k], deg = h_inc_off[k + 1] - q0;
        uint32_t m[4] = {0, 0, 0, 0};
        for (uint32_t s = 0; s < deg && s < 4; ++s) m[s] = h_oth_off[q0 + s] | h_flags[q0 + s];
        h_desc[2 * k] = make_uint4(q0, node_of_k[k], deg, 0u);
        h_desc[2 * k + 1] = make_uint4(m[0], m[1], m[2], m[3]);
      }
      LPMP_G_TRY(put(desc, h_desc, stream));
    }
    h_inc = h_inc_off; h_oth = h_oth_node;
    if (mixed) {
      h_voff.resize(N + 1); h_lab_k.resize(N); h_lab_oth.resize(Q);
      for (uint64_t k = 0; k <= N; ++k) h_voff[k] = (uint32_t)rec_off[k];
      for (uint64_t k = 0; k < N; ++k) h_lab_k[k] = Lv[node_of_k[k]];
      for (uint64_t q = 0; q < Q; ++q) h_lab_oth[q] = Lv[h_oth_node[q]];
      LPMP_G_TRY(put(voff, h_voff, stream));
      LPMP_G_TRY(put(lab_k, h_lab_k, stream));
      LPMP_G_TRY(put(lab_oth, h_lab_oth, stream));
      LPMP_G_TRY(put(toff, h_toff, stream));
      LPMP_G_TRY(put(lab_i, Lv, stream));
    }
    LPMP_TICK("upload");
    LPMP_G_TRY(cudaStreamSynchronize(stream));
    return LPMP_OK;
  }

  // ---- staged optimisation: LP::construct_factor_partition / compute_partition_pass / compute_overlapping_partition_pass
  // (/root/reference/include/LP_MP.h:1716-2051) ------------------------------------------------------------------------------------
  void construct_partition() {
    if (partition_valid) return;
    partition_valid = true;
    plans.clear();
    // the reference's union-find (include/union_find.hxx:22-50): union by size, the root of x's set wins ties; partitions in the
    // order of their roots (:70-92)
    std::vector<uint64_t> id(N), sz(N, 1);
    std::iota(id.begin(), id.end(), uint64_t(0));
    auto find = [&](uint64_t p) { uint64_t root = p; while (root != id[root]) root = id[root]; while (p != root) { const uint64_t nx = id[p]; id[p] = root; p = nx; } return root; };
    for (const auto& pr : partition_graph) {
      const uint64_t i = find(pr[0]), j = find(pr[1]);
      if (i == j) continue;
      if (sz[i] < sz[j]) { id[i] = j; sz[j] += sz[i]; } else { id[j] = i; sz[i] += sz[j]; }
    }
    std::vector<uint64_t> part_of_root(N, ~uint64_t(0));
    uint64_t np = 0;
    for (uint64_t d = 0; d < N; ++d) if (find(d) == d) part_of_root[d] = np++;
    parts.assign(np, {});
    for (uint64_t t = 0; t < N; ++t) parts[part_of_root[find(rank_order[t])]].push_back(rank_order[t]);
  }

  // the sweep over `seq`: per node the receive / send code of every incidence and the send weight (ComputeAnisotropicWeights on a
  // subsequence, include/LP_MP.h:1230-1415: receive from list members before this node and -- unless it is the first of the list --
  // from outside the list; send to list members after it, weight 1 / max(#sends, #other messages)), grouped into the level sets of
  // the dependencies inside the list
  int build_plan(const std::vector<uint64_t>& seq, Plan& pl, cudaStream_t stream, std::string& err) {
    const uint64_t NONE = ~uint64_t(0);
    std::vector<uint64_t> pos(N, NONE);
    for (uint64_t p = 0; p < seq.size(); ++p) pos[seq[p]] = p;
    std::vector<uint32_t> level(seq.size(), 0);
    uint32_t max_level = 0;
    for (uint64_t p = 0; p < seq.size(); ++p) {
      const uint32_t k = k_of[seq[p]];
      uint32_t lv = 0;
      for (uint32_t q = h_inc[k]; q < h_inc[k + 1]; ++q) { const uint64_t po = pos[h_oth[q]]; if (po != NONE && po < p) lv = std::max(lv, level[po] + 1); }
      level[p] = lv;
      max_level = std::max(max_level, lv);
    }
    pl.level_off.assign((size_t)max_level + 2, 0);
    for (uint64_t p = 0; p < seq.size(); ++p) ++pl.level_off[level[p] + 1];
    for (size_t l = 0; l + 1 < pl.level_off.size(); ++l) pl.level_off[l + 1] += pl.level_off[l];
    std::vector<uint64_t> slot(pl.level_off.begin(), pl.level_off.end() - 1);
    std::vector<uint32_t> h_node(seq.size()), h_base(seq.size());
    std::vector<Real> h_w(seq.size());
    std::vector<uint8_t> h_code;
    for (uint64_t p = 0; p < seq.size(); ++p) {
      const uint32_t k = k_of[seq[p]];
      const uint64_t t = slot[level[p]]++;
      h_node[t] = k;
      h_base[t] = (uint32_t)h_code.size();
      uint32_t n_send = 0;
      for (uint32_t q = h_inc[k]; q < h_inc[k + 1]; ++q) {
        const uint64_t po = pos[h_oth[q]];
        const bool recv = po != NONE ? po < p : p > 0, send = po != NONE && po > p;
        n_send += send ? 1u : 0u;
        h_code.push_back((uint8_t)((recv ? 1u : 0u) | (send ? 2u : 0u)));
      }
      const uint32_t deg = h_inc[k + 1] - h_inc[k];
      h_w[t] = n_send ? (Real)(1.0 / (double)std::max(n_send, deg - n_send)) : Real(0);
    }
    LPMP_G_TRY(put(pl.node, h_node, stream));
    LPMP_G_TRY(put(pl.base, h_base, stream));
    LPMP_G_TRY(put(pl.code, h_code, stream));
    LPMP_G_TRY(put(pl.w, h_w, stream));
    LPMP_G_TRY(cudaStreamSynchronize(stream));
    return LPMP_OK;
  }
  int run_plan(int kind, size_t i, int level_group, int num_sms, cudaStream_t stream, std::string& err, int* launched) {
    auto it = plans.find({kind, i});
    if (it == plans.end()) {
      std::vector<uint64_t> seq;
      if (kind == 0) seq = parts[i];
      else if (kind == 1) seq.assign(parts[i].rbegin(), parts[i].rend());
      else { const auto& other = parts[kind == 2 ? i + 1 : i - 1]; seq = parts[i]; seq.insert(seq.end(), other.rbegin(), other.rend()); }
      it = plans.emplace(std::piecewise_construct, std::forward_as_tuple(kind, i), std::forward_as_tuple()).first;
      if (int r = build_plan(seq, it->second, stream, err)) return r;
    }
    const Plan& pl = it->second;
    GraphArgs<Real> A = args(nullptr, nullptr);
    A.plan_node = pl.node.p; A.plan_base = pl.base.p; A.plan_code = pl.code.p; A.plan_w = pl.w.p;
    int G = 0, threads = 0;
    size_t smem = 0;
    if (!generic_shape(level_group, &G, &threads, &smem)) { err = "label count too large for the level kernel"; return LPMP_ERR_UNSUPPORTED; }
    const uint64_t groups = threads / G;
    for (size_t lv = 0; lv + 1 < pl.level_off.size(); ++lv) {
      const uint64_t n = pl.level_off[lv + 1] - pl.level_off[lv];
      if (!n) continue;
      A.first = (uint32_t)pl.level_off[lv]; A.count = (uint32_t)n;
      graph_node_generic_kernel<Real, 2><<<(int)std::min<uint64_t>((n + groups - 1) / groups, (uint64_t)num_sms * 8), threads, smem, stream>>>(A, G);
      ++*launched;
    }
    return LPMP_OK;
  }
  int partition_pass(uint64_t no_passes, bool overlapping, int level_group, int num_sms, cudaStream_t stream, std::string& err, int* launched) {
    construct_partition();
    const size_t np = parts.size();
    auto run = [&](int kind, size_t i) { return run_plan(kind, i, level_group, num_sms, stream, err, launched); };
#define LPMP_RUN(kind, i) do { if (int _r = run(kind, i)) return _r; } while (0)
    if (!overlapping) {   // compute_partition_pass, include/LP_MP.h:1961-1996
      for (size_t i = 0; i < np; ++i) {
        for (uint64_t it = 0; it < no_passes; ++it) { LPMP_RUN(0, i); LPMP_RUN(1, i); }
        if (i + 1 < np) LPMP_RUN(2, i);
      }
      for (size_t ri = 0; ri < np; ++ri) {
        const size_t i = np - ri - 1;
        for (uint64_t it = 0; it < no_passes; ++it) { LPMP_RUN(0, i); LPMP_RUN(1, i); }
        if (i != 0) LPMP_RUN(3, i);
      }
    } else {              // compute_overlapping_partition_pass, :2050-2070: f_forward(i) = push to i+1 of i, f_backward(i) = push to i of i+1
      for (size_t i = 0; i + 1 < np; ++i) {
        for (uint64_t it = 0; it < no_passes; ++it) { LPMP_RUN(2, i); LPMP_RUN(3, i + 1); }
        LPMP_RUN(2, i);
      }
      for (size_t ri = 1; ri < np; ++ri) {
        const size_t i = np - ri - 1;
        for (uint64_t it = 0; it < no_passes; ++it) { LPMP_RUN(3, i + 1); LPMP_RUN(2, i); }
        LPMP_RUN(3, i + 1);
      }
    }
#undef LPMP_RUN
    return LPMP_OK;
  }

  GraphArgs<Real> args(uint32_t* labels, double* partial) const {
    GraphArgs<Real> A;
    A.desc = desc.p; A.vec = vec.p; A.tab = tab.p; A.inc_off = inc_off.p; A.oth_off = oth_off.p; A.flags = flags.p; A.oth_node = oth_node.p; A.node_of = node_of.p;
    A.voff = mixed ? voff.p : nullptr; A.toff = mixed ? toff.p : nullptr; A.lab_k = mixed ? lab_k.p : nullptr; A.lab_oth = mixed ? lab_oth.p : nullptr;
    A.labels = labels; A.partial = partial;
    A.first = 0; A.count = (uint32_t)N; A.L = L; A.P = P; A.pw = pw; A.repam = 0; A.forward = 1; A.primal = 0; A.residual = 0;
    A.plan_node = nullptr; A.plan_base = nullptr; A.plan_code = nullptr; A.plan_w = nullptr;
    A.wtab[0] = Real(0);
    for (int n = 1; n < 8; ++n) A.wtab[n] = (Real)(1.0 / (double)n);
    return A;
  }

  template <int MODE, int R = kDmax, bool LBF = false>
  void launch_fast(const GraphArgs<Real>& A, int ctas, cudaStream_t stream) const {
    if (L == 4) graph_node_kernel<Real, 4, kDmax, MODE, R, LBF><<<ctas, 256, 0, stream>>>(A);
    else if (L == 8) graph_node_kernel<Real, 8, kDmax, MODE, R, LBF><<<ctas, 256, 0, stream>>>(A);
    else if constexpr (sizeof(Real) == 4) graph_node_kernel<Real, 16, kDmax, MODE, R, LBF><<<ctas, 256, 0, stream>>>(A);
  }
  int update_ctas(uint64_t count, int r, int num_sms) const {
    const uint64_t npc = 256 / L;
    const int per_sm = ctas_per_sm > 0 ? ctas_per_sm : (r == 0 ? 32 : (r == 1 ? 24 : 12));
    return (int)std::min<uint64_t>((count + npc - 1) / npc, (uint64_t)num_sms * per_sm);
  }
  // node update of [first, first + count) by the instance with registers for `r` receives per node; `lb`: also the bound terms
  // (one partial sum per CTA at A.partial)
  template <bool LBF>
  void launch_update(GraphArgs<Real>& A, uint64_t first, uint64_t count, int r, int ctas, cudaStream_t stream) const {
    A.first = (uint32_t)first; A.count = (uint32_t)count;
    switch (r) {
      case 0: launch_fast<0, 0, LBF>(A, ctas, stream); break;
      case 1: launch_fast<0, 1, LBF>(A, ctas, stream); break;
      case 2: launch_fast<0, 2, LBF>(A, ctas, stream); break;
      case 3: launch_fast<0, 3, LBF>(A, ctas, stream); break;
      default: launch_fast<0, 4, LBF>(A, ctas, stream); break;
    }
  }
  // LP::LowerBound accumulated by the last anisotropic backward pass: fused_partial[0 .. fused_used)
  mutable Buffer<double> fused_partial;
  mutable int fused_used = 0;
  bool fused_bound_enabled = true;
  bool all_fast() const {
    if (!fast_ok) return false;
    for (size_t lv = 0; lv + 1 < level_off.size(); ++lv) if (level_fast[lv] != level_off[lv + 1]) return false;
    return true;
  }
  // lanes per node / threads per CTA / shared memory of the generic kernel
  bool generic_shape(int level_group, int* G, int* threads, size_t* smem) const {
    int g = group_size_for(L);
    if (level_group > 0 && level_group < g && (level_group & (level_group - 1)) == 0) g = level_group;
    int t = 256;
    const size_t per_group = 3 * (size_t)L * sizeof(Real);
    while (t > g && (t / g) * per_group > 48 * 1024) t >>= 1;
    if ((t / g) * per_group > 48 * 1024) return false;
    *G = g; *threads = t; *smem = (t / g) * per_group;
    return true;
  }

  // one pass = the levels in order (forward) or reverse order (backward); returns the number of kernels launched, < 0 on error
  int ctas_per_sm = 0;   // fast kernel: CTAs per SM of a launch ("graph_ctas_per_sm"; 0 = default)
  int pass(bool forward, bool primal, int repam, bool residual, uint32_t* labels, int level_group, int num_sms, cudaStream_t stream, std::string& err,
           bool* fused) const {
    GraphArgs<Real> A = args(labels, nullptr);
    A.repam = repam; A.forward = forward ? 1 : 0; A.primal = primal ? 1 : 0; A.residual = residual ? 1 : 0;
    int G = 0, threads = 0;
    size_t smem = 0;
    if (!generic_shape(level_group, &G, &threads, &smem)) { err = "label count too large for the level kernel"; return -1; }
    const size_t nlev = level_off.size() - 1;
    const bool aniso = repam == LPMP_REPAM_ANISOTROPIC || repam == LPMP_REPAM_ANISOTROPIC2;
    // the launches of the fast part: consecutive ranges that need the same instance go out as one launch
    struct Launch { size_t lv; uint64_t first, count; int r, ctas; };
    std::vector<Launch> fast_launches;
    size_t slots = 0;
    for (size_t lv = 0; lv < nlev; ++lv) {
      uint64_t first = 0, count = 0;
      int cur = -1;
      auto flush = [&] { if (count) { const int c = update_ctas(count, cur, num_sms); fast_launches.push_back({lv, first, count, cur, c}); slots += (size_t)c; count = 0; } };
      for (const Segment& sg : segments[lv]) {
        const int r = (int)(aniso ? (forward ? sg.n_in : sg.n_out) : sg.n_in + sg.n_out);
        if (r != cur) flush();
        if (!count) first = sg.first;
        cur = r; count += sg.count;
      }
      flush();
    }
    *fused = fused_bound_enabled && !forward && !primal && aniso && all_fast();
    if (*fused) {
      if (fused_partial.n < slots && fused_partial.alloc(slots) != cudaSuccess) { err = "out of device memory (fused bound)"; return -1; }
      fused_used = 0;
    }
    int launched = 0;
    for (size_t t = 0; t < nlev; ++t) {
      const size_t lv = forward ? t : nlev - 1 - t;
      const uint64_t n_fast = level_fast[lv] - level_off[lv], n_slow = level_off[lv + 1] - level_fast[lv];
      if (n_fast) {
        for (const Launch& ln : fast_launches) {
          if (ln.lv != lv) continue;
          if (*fused) {
            A.partial = fused_partial.p + fused_used;
            fused_used += ln.ctas;
            launch_update<true>(A, ln.first, ln.count, ln.r, ln.ctas, stream);
          } else {
            launch_update<false>(A, ln.first, ln.count, ln.r, ln.ctas, stream);
          }
          ++launched;
        }
      }
      if (n_slow) {
        A.first = (uint32_t)level_fast[lv]; A.count = (uint32_t)n_slow;
        const uint64_t groups = threads / G;
        graph_node_generic_kernel<Real, 0><<<(int)std::min<uint64_t>((n_slow + groups - 1) / groups, (uint64_t)num_sms * 8), threads, smem, stream>>>(A, G);
        ++launched;
      }
    }
    return launched;
  }

  // LP::LowerBound terms into partial[0 .. *used): one sweep over all records (the fast kernel when every node qualifies for it)
  int bound(double* partial, int* used, int level_group, int num_sms, cudaStream_t stream, std::string& err) const {
    GraphArgs<Real> A = args(nullptr, partial);
    bool all_fast = fast_ok;
    for (size_t lv = 0; all_fast && lv + 1 < level_off.size(); ++lv) all_fast = level_fast[lv] == level_off[lv + 1];
    if (all_fast) {
      const uint64_t npc = 256 / L;
      const int ctas = (int)std::min<uint64_t>((N + npc - 1) / npc, std::min<uint64_t>(kPartials, (uint64_t)num_sms * (ctas_per_sm > 0 ? ctas_per_sm : 12)));
      launch_fast<1>(A, ctas, stream);
      *used = ctas;
      return 1;
    }
    int G = 0, threads = 0;
    size_t smem = 0;
    if (!generic_shape(level_group, &G, &threads, &smem)) { err = "label count too large for the level kernel"; return -1; }
    const uint64_t groups = threads / G;
    const int ctas = (int)std::min<uint64_t>((N + groups - 1) / groups, (uint64_t)kPartials);
    graph_node_generic_kernel<Real, 1><<<ctas, threads, smem, stream>>>(A, G);
    *used = ctas;
    (void)num_sms;
    return 1;
  }
  static constexpr size_t partial_slots() { return (size_t)kPartials; }

  int primal(const uint32_t* labels, double* partial, int* used, cudaStream_t stream) const {
    GraphArgs<Real> A = args(const_cast<uint32_t*>(labels), partial);
    const int ctas = (int)std::min<uint64_t>((N + 255) / 256, (uint64_t)kPartials);
    graph_primal_kernel<Real><<<ctas, 256, 0, stream>>>(A);
    *used = ctas;
    return 1;
  }
};

}  // namespace lpmp
