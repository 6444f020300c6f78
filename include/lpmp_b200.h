/*
 * lpmp_b200.h -- C ABI of the B200-native LP_MP message-passing hot path.
 *
 * Drop-in boundary: this is the surface a reference-style driver binds instead of LP<FMC>
 * (the LP_TYPE consumed by Solver<LP_TYPE,VISITOR>, /root/reference/include/solver.hxx:29-30).
 * Each entry point cites the reference interface it replaces.  Plain pointers and sizes only;
 * no C++/torch types cross the boundary; errors are integer status codes (0 = ok) with a
 * per-handle message (the reference throws std::runtime_error / asserts).
 *
 * The same header is implemented twice:
 *   - lp_mp_b200/csrc (CUDA, sm_100a)       symbols  lpmp_<name>          -> liblpmp_b200.so
 *   - oracle/lpmp_oracle.cpp (CPU, checker) symbols  lpmp_oracle_<name>   -> oracle/liblpmp_oracle.so
 * so that one test body can drive both.  The product never loads the oracle.
 *
 * Model.  Node i has theta_i in R^L.  Edge e=(i,j) has parameters c_e and two
 * vectors a_e ("msg1", on i's variable) and b_e ("msg2", on j's variable);
 * theta_e(x,y) = c_e(x,y) + a_e(x) + b_e(y).  Messages are implicit: only these reparametrised
 * potentials are stored (include/factors_messages.hxx:432).
 */
#ifndef LPMP_B200_H
#define LPMP_B200_H

#include <stdint.h>

#ifndef LPMP_FN
#define LPMP_FN(name) lpmp_##name
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lpmp_solver lpmp_solver; /* opaque; owns all host + device memory of one LP */

enum { LPMP_OK = 0, LPMP_ERR_INVALID = 1, LPMP_ERR_STATE = 2, LPMP_ERR_CUDA = 3, LPMP_ERR_NOT_DAG = 4, LPMP_ERR_UNSUPPORTED = 5 };

/* storage / arithmetic type of theta, a, b, c on the device.  F64 = parity mode (REAL = double,
 * include/config.hxx:28); F32 = throughput mode (REAL = float is the commented alternative, :23-26). */
enum { LPMP_F32 = 0, LPMP_F64 = 1 };

/* LPReparametrizationMode, include/config.hxx:72 (Mixed: out of scope).  ANISOTROPIC2 = ComputeAnisotropicWeights2
 * (include/LP_MP.h:1084-1154): the receive mask of ANISOTROPIC, send weight 1 / #{messages sent to factors later in the pass}. */
enum { LPMP_REPAM_ANISOTROPIC = 0, LPMP_REPAM_UNIFORM = 1, LPMP_REPAM_DAMPED_UNIFORM = 2, LPMP_REPAM_ANISOTROPIC2 = 3 };

/* how a factor sends, --reparametrizationType of the reference LP (include/LP_MP.h:589,704-722,988-1004):
 * SHARED = FactorContainer::UpdateFactor (every message gets omega * the snapshot), RESIDUAL = update_factor_residual
 * (include/factors_messages.hxx:2270-2279,2957-3007: after the shared send, message k in turn also receives
 * (omega_1 + ... + omega_k) * what is left in the factor).  "adaptive" (:2859-2930) needs send_message_to_right_improvement of the
 * marginalisation message, which lives in the un-mounted sibling repository: lpmp_set_send_mode answers LPMP_ERR_UNSUPPORTED.
 * "partition" / "overlapping_partition": lpmp_compute_partition_pass below. */
enum { LPMP_SEND_SHARED = 0, LPMP_SEND_RESIDUAL = 1, LPMP_SEND_ADAPTIVE = 2 };

/* pairwise factor family: PairwiseSimplexFactor / pairwise_potts_factor (test/potts_factor.cpp:21-71) /
 * truncated linear min(slope*|x-y|, trunc) (new, BASELINE.json config 2) */
enum { LPMP_PAIRWISE_DENSE = 0, LPMP_PAIRWISE_POTTS = 1, LPMP_PAIRWISE_TRUNCLIN = 2 };

/* how the update order of LP::SortFactors (include/LP_MP.h:730-797) is realised on the device */
enum {
  LPMP_SCHEDULE_LEVELS = 0,    /* general graphs: level sets of the factor-relation DAG given by the node ranks */
  LPMP_SCHEDULE_WAVEFRONT = 1, /* grids only: persistent anti-diagonal row-band pipeline (Potts / truncated linear, L in {8,16,32,64}) */
};

/* iteration order of a factor's messages of one type (include/factors_messages.hxx:2005-2077 LIFO
 * intrusive list -- the storage the SRMP configuration selects -- vs :1772-1930 chunked FIFO) */
enum { LPMP_EDGE_ORDER_LIFO = 0, LPMP_EDGE_ORDER_FIFO = 1 };

typedef struct lpmp_stats {
  uint64_t num_nodes, num_edges, num_labels;
  uint64_t num_levels;            /* sequential steps per pass of the active schedule */
  uint64_t kernel_launches;       /* CUDA kernels launched by this handle so far */
  uint64_t message_updates;       /* receives + sends performed so far (SURVEY.md 8d unit of work) */
  uint64_t device_bytes;          /* bytes of device memory owned */
  double   algorithmic_bytes_per_pass; /* s*[2NL + E(P+5L)], BASELINE.md section 4 */
} lpmp_stats;

/* ---- lifetime: LP<FMC>::LP(cmd) / ~LP, include/LP_MP.h:212,596-600 ------------------------------- */
int LPMP_FN(create)(int dtype, int device, lpmp_solver** out);
int LPMP_FN(destroy)(lpmp_solver* h);
const char* LPMP_FN(last_error)(const lpmp_solver* h);

/* ---- problem construction ------------------------------------------------------------------------
 * General graphs.  add_unaries == n x LP::add_factor<UnaryFactor> (include/LP_MP.h:238-253);
 * add_pairwise_* == per edge add_factor<PairwiseFactor> + 2 x add_message + AddFactorRelation(u_i,p),
 * AddFactorRelation(p,u_j) (include/LP_MP.h:266-285,697-702), as mrf.AddUnaryFactor / AddPairwiseFactor do
 * (test/graphical_model.cpp:98-108).  One label count L and one pairwise family per LP.  Inputs are host
 * buffers of doubles and are copied. */
int LPMP_FN(add_unaries)(lpmp_solver* h, uint64_t n, uint32_t L, const double* theta /* [n][L] */);
/* the same with one label count per variable (test/graphical_model.cpp:74-96 builds 2-, 3- and 4-label variables in one model;
 * UnaryPairwiseMessage(4,3) in test/simplex_marginalization.cpp:22-23): theta is packed, L[0] + ... + L[n-1] doubles.  In an LP with
 * mixed label counts every per-variable / per-edge array of this header is packed the same way (theta_i: L_i entries; a_e: L_i(e),
 * b_e: L_j(e); dense table of edge e: L_i(e) x L_j(e), x_i major).  General graphs only (level / colouring schedule). */
int LPMP_FN(add_unaries_var)(lpmp_solver* h, uint64_t n, const uint32_t* L /* [n] */, const double* theta /* packed */);
int LPMP_FN(add_pairwise_dense)(lpmp_solver* h, uint64_t m, const uint64_t* i, const uint64_t* j, const double* tables /* [m][L][L], x_i major */);
int LPMP_FN(add_pairwise_potts)(lpmp_solver* h, uint64_t m, const uint64_t* i, const uint64_t* j, const double* lambda /* [m] */);
int LPMP_FN(add_pairwise_trunclin)(lpmp_solver* h, uint64_t m, const uint64_t* i, const uint64_t* j, const double* slope, const double* trunc /* [m] each */);
/* Orientation of the factor-relation DAG: every edge is directed from its lower-rank to its higher-rank
 * endpoint (default rank = node index, i.e. the reference's u_i -> p -> u_j with i<j).  A proper colouring
 * passed as rank = colour*N + index gives the graph-colouring schedule. */
int LPMP_FN(set_node_ranks)(lpmp_solver* h, const uint64_t* rank /* [N], distinct */);
int LPMP_FN(set_edge_order)(lpmp_solver* h, int edge_order);
int LPMP_FN(set_schedule)(lpmp_solver* h, int schedule);
/* launch-shape knobs of the device schedule (no effect on results): "wavefront_max_ctas" (cap on resident bands),
 * "wavefront_ctas_per_sm", "wavefront_tile_cols" (8 or 4 columns per TMA tile), "wavefront_warps_per_cta" (consecutive
 * bands per CTA with the hand-off between them in shared memory; default 1), "wavefront_lean" (anisotropic passes: 1 = compute
 * warp + two data-mover warps per band, 0 = one warp does everything; default automatic), "wavefront_cluster" (CTAs per
 * thread-block cluster of the lean variant: consecutive bands hand off through distributed shared memory; 1 = none, default 8), "wavefront_trace" (see get_trace),
 * "wavefront_rows" / "wavefront_split" (1 / 0: force / forbid the one-label-per-lane and the one-row-per-warp kernels; default automatic),
 * "wavefront_split_uni" (0: uniform / damped-uniform passes on the one-warp kernel instead of the split kernel's instances),
 * "wavefront_fused_bound" (backward passes accumulate LowerBound: 1 / 0 / automatic).
 * Before finalize. */
int LPMP_FN(set_tuning)(lpmp_solver* h, const char* key, int value);
/* Deterministic greedy colouring in node order (host utility, no handle state): colour[i] = smallest colour
 * not used by an already coloured neighbour.  Returns the number of colours through *num_colours. */
int LPMP_FN(greedy_colouring)(uint64_t n, uint64_t m, const uint64_t* i, const uint64_t* j, uint32_t* colour, uint32_t* num_colours);

/* Grids (4-connected, nodes row-major r*W+c, edges oriented low -> high index, per node the horizontal edge
 * is inserted before the vertical one: SURVEY.md 8d).  Storage is allocated on the device and zeroed.
 * src_dtype is the element type of the HOST buffers (LPMP_F32 / LPMP_F64). */
int LPMP_FN(grid_create)(lpmp_solver* h, uint32_t H, uint32_t W, uint32_t L, int pairwise_type);
int LPMP_FN(grid_set_unaries)(lpmp_solver* h, const void* theta /* [H][W][L] */, int src_dtype);
/* P = L*L (dense), 1 (Potts: lambda), 2 (truncated linear: slope, trunc) parameters per edge */
int LPMP_FN(grid_set_pairwise)(lpmp_solver* h, const void* horizontal /* [H][W-1][P] */, const void* vertical /* [H-1][W][P] */, int src_dtype);
/* Counter-based synthetic potentials generated in place on the device (no host copy): theta ~ U[0,1),
 * pairwise parameters ~ U[p_lo, p_hi) (dense: every table entry; Potts: lambda; truncated linear: p_lo = slope,
 * p_hi = trunc, constant).  The generator is specified in DESIGN.md so that the oracle reproduces it. */
int LPMP_FN(grid_fill_synthetic)(lpmp_solver* h, uint64_t seed, double p_lo, double p_hi);

/* Row-stripe partition (multi-GPU): declare that this grid is rows [row0, row0+H) of a taller grid with
 * H_total rows (call before grid_fill_synthetic / finalize).  A stripe owns its nodes, its horizontal and inner
 * vertical edges and, for the vertical edges crossing a stripe boundary, the message vector on ITS side (bU of its
 * first row, aD of its last row) plus a copy of their parameters.  Stripes run the wavefront schedule: the
 * anisotropic mode, and (Potts / truncated-linear grids) the uniform and damped-uniform modes of the rounding iterations
 * (include/visitors/standard_visitor.hxx:44).
 *
 * A pass consumes the neighbouring stripe's outgoing vectors from an "in" buffer of W*L reals (storage dtype) per
 * boundary: in(0 = top) feeds the forward pass, in(1 = bottom) the backward pass.  Slots carry their own validity (a
 * slot holds the all-ones NaN until the vector is stored; the pass puts that sentinel back after reading), so the
 * same kernel serves both couplings:
 *   - per-pass halo: deposit the neighbour's vectors (lpmp_grid_get_boundary_vectors on the neighbour, then a copy
 *     or ncclSend/Recv into in_buf) before launching the pass;
 *   - exact pipelined: lpmp_grid_set_peer_boundary hands the pass kernel the neighbour's in_buf (peer-mapped
 *     pointer, same node: lpmp_ipc_export / lpmp_ipc_open); the last band of this stripe then stores each
 *     outgoing vector straight into it over NVLink while the neighbour's first band polls -- results are
 *     bit-identical to the unstriped LP.
 * Uniform modes: a node also receives from its two NOT yet processed neighbours, so a forward pass additionally consumes in(1)
 * (the lower stripe's bU vectors as they were BEFORE the pass) and a backward pass in(0) (the upper stripe's aD vectors): every
 * stripe delivers that side with lpmp_grid_push_boundary before launching its pass, in both couplings. */
int LPMP_FN(grid_set_stripe)(lpmp_solver* h, uint32_t row0, uint32_t H_total);
/* parameters of the boundary-crossing vertical edges, [W][P] each (NULL where there is no neighbour) */
int LPMP_FN(grid_set_boundary_pairwise)(lpmp_solver* h, const void* top, const void* bottom, int src_dtype);
int LPMP_FN(grid_boundary_ptrs)(lpmp_solver* h, int which, void** in_buf, uint64_t* bytes);
/* which: 0 = deliver the backward pass's outgoing vectors to the upper neighbour's in(1), 1 = the forward pass's to
 * the lower neighbour's in(0); NULL = do not deliver */
int LPMP_FN(grid_set_peer_boundary)(lpmp_solver* h, int which, void* peer_in_buf);
/* gather this stripe's current side of a boundary (0: bU of the first row, 1: aD of the last row) into a device
 * buffer of W*L reals, enqueued on the handle's stream */
int LPMP_FN(grid_get_boundary_vectors)(lpmp_solver* h, int which, void* dst_dev);
/* stripe share of LP::LowerBound (include/LP_MP.h:1507-1518): own unaries, own edges and the edges to the stripe
 * below, whose b-side vectors (the lower stripe's lpmp_grid_get_boundary_vectors(0)) are passed as a device
 * pointer (NULL for the last stripe).  The LP's bound is the sum over stripes (all-reduce). */
int LPMP_FN(grid_lower_bound_partial)(lpmp_solver* h, const void* below_b, double* out);
/* the same, left in DEVICE memory (d_out: one double) and ordered on the handle's stream -- no host synchronisation, so that a
 * multi-GPU driver can all-reduce it (ncclAllReduce on lpmp_stream) and read the LP's bound with one stream-ordered copy */
int LPMP_FN(grid_lower_bound_partial_device)(lpmp_solver* h, const void* below_b, double* d_out);
/* stripe share of LP::EvaluatePrimal (include/LP_MP.h:1520-1536), device result as above: own unaries and edges plus the edges to
 * the stripe below, given that stripe's b-side vectors ([W][L]) and the labels of its first row ([W], lpmp_grid_get_boundary_labels
 * on the neighbour); +inf while no rounding pass has run */
int LPMP_FN(grid_evaluate_primal_partial_device)(lpmp_solver* h, const void* below_b, const uint32_t* below_labels, double* d_out);
/* labels of the first (which = 0) / last (1) row into a device buffer of W uint32, enqueued on the handle's stream */
int LPMP_FN(grid_get_boundary_labels)(lpmp_solver* h, int which, uint32_t* dst_dev);
/* per-pass halo without a host round trip: store this stripe's side of boundary `which` (0: bU of the first row, 1: aD of the
 * last row) into the neighbour's in-buffer (peer-mapped pointer from lpmp_ipc_open, or a pointer of the same process).  The push
 * obeys the slot protocol -- a slot is only written once the consumer of the previous pass has re-armed it -- so pushes and passes
 * of neighbouring stripes order themselves through the data, with no collective and no host synchronisation.
 * Layout of an in-buffer: slot q holds the vector of the q-th column IN THE PROCESSING ORDER of the pass that consumes it (the
 * pass kernels deliver and poll that way): column q for a forward pass, column W - 1 - q for a backward pass.  `backward` != 0
 * says the neighbour's next pass is a backward pass. */
int LPMP_FN(grid_push_boundary)(lpmp_solver* h, int which, void* peer_in_buf, int backward);
/* cudaIpc wrappers (64-byte handles) so that a host language without CUDA bindings can exchange in_buf pointers */
int LPMP_FN(ipc_export)(const void* devptr, void* handle64);
int LPMP_FN(ipc_open)(const void* handle64, void** devptr);
int LPMP_FN(ipc_close)(void* devptr);

/* finish construction: build schedule + weights (SortFactors / ComputeAnisotropicWeights / ComputeUniformWeights,
 * include/LP_MP.h:730-797,1230-1449) and upload.  Mutating the problem afterwards is an error. */
int LPMP_FN(finalize)(lpmp_solver* h);

/* ---- the LP_TYPE surface used by Solver / visitors (include/solver.hxx:185,233,248,263,269,274,281,306,323,390-392) */
int LPMP_FN(begin)(lpmp_solver* h);                                  /* LP::Begin, include/LP_MP.h:704-728 */
int LPMP_FN(end)(lpmp_solver* h);                                    /* LP::End */
int LPMP_FN(set_reparametrization)(lpmp_solver* h, int mode);        /* LP::set_reparametrization */
int LPMP_FN(set_send_mode)(lpmp_solver* h, int mode);                /* reparametrization_type_, include/LP_MP.h:704-722 (any time after finalize) */
int LPMP_FN(compute_pass)(lpmp_solver* h, uint64_t iteration);       /* LP::ComputePass, include/LP_MP.h:869-887 */
int LPMP_FN(compute_forward_pass)(lpmp_solver* h);                   /* :889-900 */
int LPMP_FN(compute_backward_pass)(lpmp_solver* h);                  /* :902-911 */
int LPMP_FN(compute_forward_pass_and_primal)(lpmp_solver* h, uint64_t iteration);  /* :913-922 */
int LPMP_FN(compute_backward_pass_and_primal)(lpmp_solver* h, uint64_t iteration); /* :924-933 */
/* Staged optimisation (--reparametrizationType partition | overlapping_partition, include/LP_MP.h:463-465,869-887,1716-2051).
 * put_in_same_partition joins the partitions of variables i[k] and j[k] (LP::put_in_same_partition on the two unary factors; only
 * updated factors, i.e. unaries, make up a partition; a variable never mentioned is a partition of its own).  Partitions are ordered
 * as construct_factor_partition orders them (by the root of the reference's union-find, include/union_find.hxx:36-50), the variables
 * of a partition by their rank.
 * compute_partition_pass == LP::compute_partition_pass(inner_iterations) (overlapping == 0: per partition `inner_iterations` forward +
 * backward sweeps over the partition alone, then one sweep pushing into the next partition; then the same from the last partition
 * back) or LP::compute_overlapping_partition_pass (overlapping != 0: the sweeps run over pairs of neighbouring partitions).  A sweep
 * over a subset S follows ComputeAnisotropicWeights on S (include/LP_MP.h:1230-1415): a variable receives on the edges to variables
 * of S before it and -- unless it is the first of S -- on its edges leaving S, sends on the edges to variables of S after it with
 * weight 1 / max(#those, #its other edges).  LP::ComputePass in partition mode is this call; in overlapping_partition mode this call
 * followed by a forward and a backward pass.  General graphs (add_unaries / add_pairwise_*); a grid is partitioned by building it
 * through those calls.
 * NOTE the reference sorts the factors of a partition with a comparator that compares an element with itself (include/LP_MP.h:1776):
 * partitions of at most 16 factors keep their insertion order (== rank order for the default ranks), larger ones come out in an
 * order that is an artefact of std::sort.  This library always uses the rank order. */
int LPMP_FN(put_in_same_partition)(lpmp_solver* h, uint64_t m, const uint64_t* i, const uint64_t* j);
int LPMP_FN(compute_partition_pass)(lpmp_solver* h, uint64_t inner_iterations, int overlapping);
int LPMP_FN(lower_bound)(lpmp_solver* h, double* out);               /* LP::LowerBound, :1507-1518 */
int LPMP_FN(evaluate_primal)(lpmp_solver* h, double* out);           /* LP::EvaluatePrimal, :1520-1536 (+inf if inconsistent) */
int LPMP_FN(check_primal_consistency)(lpmp_solver* h, int* out);     /* LP::CheckPrimalConsistency, :1066-1082 */
int LPMP_FN(add_to_constant)(lpmp_solver* h, double x);              /* LP::add_to_constant, :462 */
int LPMP_FN(num_factors)(const lpmp_solver* h, uint64_t* out);       /* LP::GetNumberOfFactors (N + E) */
int LPMP_FN(num_messages)(const lpmp_solver* h, uint64_t* out);      /* LP::GetNumberOfMessages (2E) */

/* ---- state access (serialize_dual / serialize_primal equivalents, include/serialization.hxx:228-426) -- */
int LPMP_FN(get_labeling)(lpmp_solver* h, uint32_t* out /* [N] */);
/* Solver::RegisterPrimal (include/solver.hxx:321-340) keeps the best labeling seen so far: keep_labeling copies the current one
 * aside ON THE DEVICE (stream-ordered, no host synchronisation -- a rounding iteration never waits for N labels to travel),
 * get_kept_labeling downloads that copy */
int LPMP_FN(keep_labeling)(lpmp_solver* h);
int LPMP_FN(get_kept_labeling)(lpmp_solver* h, uint32_t* out /* [N] */);
int LPMP_FN(get_unaries)(lpmp_solver* h, double* theta /* [N][L] */);
int LPMP_FN(get_messages)(lpmp_solver* h, double* a /* [E][L] */, double* b /* [E][L] */); /* edge insertion order */
int LPMP_FN(set_messages)(lpmp_solver* h, const double* a, const double* b);
int LPMP_FN(set_unaries)(lpmp_solver* h, const double* theta);
/* The dual variables in the byte layout of the reference's archive over ALL factors in insertion order
 * (factor_archive<serialization_functor::dual>, include/factor_archive.hxx:34-48; save_archive / load_archive,
 * include/serialization.hxx:226-425; FactorContainer::serialize_dual, include/factors_messages.hxx:3196-3209): per factor its
 * serialize_dual() members as REAL = double -- unary simplex factor: theta (L_i doubles); pairwise simplex factor: msg1 (L_i
 * doubles) then msg2 (L_j doubles).  Factors appear in the order of the add_* calls (grids: all unaries, then per node the
 * horizontal edge before the vertical one).  save -> lpmp_create -> same construction calls -> finalize -> load resumes the solve
 * with a bit-identical trajectory in LPMP_F64 (checkpoint / warm start; tests/test_dual_archive.py). */
int LPMP_FN(dual_size)(lpmp_solver* h, uint64_t* bytes);
int LPMP_FN(save_dual)(lpmp_solver* h, void* archive, uint64_t bytes);
int LPMP_FN(load_dual)(lpmp_solver* h, const void* archive, uint64_t bytes);
int LPMP_FN(get_stats)(lpmp_solver* h, lpmp_stats* out);
int LPMP_FN(synchronize)(lpmp_solver* h);
/* debug timeline of the last wavefront pass (after set_tuning "wavefront_trace" = 1): per band 4 globaltimer stamps in ns
 * {band claimed, first tile in shared memory, first K steps done, band done}; n = number of uint64 the buffer holds */
int LPMP_FN(get_trace)(lpmp_solver* h, uint64_t* out, uint64_t n);
/* the CUDA stream (cudaStream_t) all work of this handle is enqueued on, for event timing by the caller */
int LPMP_FN(stream)(lpmp_solver* h, void** out);

/* ---- single-factor entry points used by the parity tests (SURVEY.md 8b last row) -----------------------
 * n independent pairwise factors with L1 x L2 labels; side 0: out[x1] = msg1(x1) + min_x2 [c + msg2],
 * side 1: out[x2] = msg2(x2) + min_x1 [c + msg1]   (min_marginal_1/2, test/potts_factor.cpp:27-31;
 * arithmetic of matrix::min1(v)/min2(v), include/vector.hxx:850-859,889-913).  Host buffers of doubles;
 * computed on the device in `dtype`.  lb (optional) receives PairwiseSimplexFactor::LowerBound per factor. */
int LPMP_FN(minmarginal_dense)(int dtype, uint64_t n, uint32_t L1, uint32_t L2, const double* tables, const double* msg1, const double* msg2, int side, double* out, double* lb);
int LPMP_FN(minmarginal_potts)(int dtype, uint64_t n, uint32_t L, const double* lambda, const double* msg1, const double* msg2, int side, double* out, double* lb);
int LPMP_FN(minmarginal_trunclin)(int dtype, uint64_t n, uint32_t L, const double* slope, const double* trunc, const double* msg1, const double* msg2, int side, double* out, double* lb);
/* n independent FactorContainer::UpdateFactor[Primal] calls (include/factors_messages.hxx:2256-2261,2332-2373) on unary factors with
 * `deg` incident pairwise factors each, run by the SAME kernels the passes of general graphs launch (kernels_graph.cuh: the register
 * kernel for dense tables with L in {4, 8, 16} and deg <= 4, the generic kernel otherwise): ReceiveMessages with the receive mask of
 * the mode -> first minimiser -> SendMessages with the mode's weight.  later[k][s] != 0: the other variable of message s comes later
 * in the forward order; `forward` selects the pass direction (a backward pass receives from the later ones).  theta [n][L]; own
 * [n][deg][L] = the node's side of its edges, in its message order; other [n][deg][L] = the other side; params [n][deg][P] in
 * receiver orientation (dense: params[x_own][y_other]).  Outputs: theta_out, own_out (same shapes), label [n]. */
int LPMP_FN(node_update)(int dtype, uint64_t n, uint32_t L, uint32_t deg, int pairwise_type, int repam, int forward, const uint8_t* later,
                         const double* theta, const double* own, const double* other, const double* params, double* theta_out, double* own_out,
                         uint32_t* label);
/* n unary factors: lb[k] = min theta_k (vector::min, include/vector.hxx:289-313), arg[k] = first minimiser */
int LPMP_FN(unary_min)(int dtype, uint64_t n, uint32_t L, const double* theta, double* lb, uint32_t* arg);

#ifdef __cplusplus
}
#endif
#endif /* LPMP_B200_H */
