/* libngpd -- C ABI of the B200-native normal-guided point-cloud denoising hot path.
 *
 * The reference (Ruubje/Normal-Guided-Pointcloud-Denoiser, Pointcloud/Modules) has no FFI of its own:
 * its seam is a Python object API that calls SciPy / torch-CPU / torch_scatter.  Each entry point below
 * replaces one of those calls; the reference line it stands in for is cited next to it.  The Python
 * mirror of the reference classes (package directory `normal-guided-pointcloud-denoiser_b200/`) binds
 * these symbols with ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`;
 *   - positions / normals are fp32, row-major [n,3]; neighbour tables are int32, row-major [m,k];
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous on that
 *     stream unless stated otherwise;
 *   - return value 0 = ok, <0 = error, message in ngpd_last_error() (thread-local);
 *   - no entry point falls back to the CPU: without a CUDA device every compute call fails.
 */
#ifndef NGPD_H
#define NGPD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ngpd_grid ngpd_grid_t;       /* frozen spatial index ("tree") */
typedef struct ngpd_session ngpd_session_t; /* fused, tree-order denoising state */

typedef struct ngpd_grid_info {
    int64_t n;
    double cell_size;
    int32_t dims[3];
    int32_t bricks;
    int64_t occupied_cells;
    int64_t bytes;
    float bbox[6];
    int32_t rebuilds;
} ngpd_grid_info_t;

const char* ngpd_last_error(void);
int ngpd_version(void);

/* ---- spatial index: replaces scipy.spatial.KDTree(graph.pos.cpu()), Selector.py:141 ---------------
 * Copies the n points (the "construction-time positions"; later edits of `pos` do not affect the index,
 * exactly like the SciPy tree).  cell_size <= 0 chooses the cell from the data so that an occupied cell
 * holds about 0.4*k_hint points.  Synchronises `stream`. */
int ngpd_grid_create(const float* pos, int64_t n, float cell_size, int k_hint, void* stream, ngpd_grid_t** out);
int ngpd_grid_destroy(ngpd_grid_t* grid);
int ngpd_grid_info(const ngpd_grid_t* grid, ngpd_grid_info_t* out);
/* perm_out[s] = original index of the s-th point in tree (Morton) order */
int ngpd_grid_order(const ngpd_grid_t* grid, int32_t* perm_out, void* stream);

/* ---- k nearest neighbours: replaces KDTree.query(pos, k), Selector.py:243; torch_cluster.knn_graph,
 * GraphBuilder.py:63 (NGPD_KNN_SKIP_SELF); torch_geometric.nn.pool.knn(x, y, 1), Utils.py:260-261.
 * Rows ascend by (fp64 squared distance of the fp32 coordinates, tree index); slots beyond the tree size
 * hold n, as SciPy does.  d2_out (nullable) receives the fp64 squared distances rounded to fp32. */
#define NGPD_KNN_SKIP_SELF 1        /* query i IS tree point i: leave it out of its own row */
#define NGPD_KNN_QUERY_IS_TREE 2    /* query i is the (possibly moved) tree point i: visit queries in tree order */
#define NGPD_KNN_COHERENT 4         /* queries are already in a spatially coherent order: skip the ordering pass */
#define NGPD_KNN_EXACT_ONLY 8       /* skip the warp-lockstep fast path: every query through the exact shell search (tests) */
int ngpd_knn(const ngpd_grid_t* grid, const float* query, int64_t m, int k, int flags,
             int32_t* idx_out, float* d2_out, void* stream);

/* nearest tree point only, for the metrics of Utils.py:253-295.  d2_out[m] = fp32 squared distance
 * recomputed from the fp32 coordinates as the reference does; idx_out nullable. */
int ngpd_nn_sqdist(const ngpd_grid_t* grid, const float* query, int64_t m, int flags,
                   float* d2_out, int32_t* idx_out, void* stream);

/* ngpd_nn_sqdist with the reduction the metrics' callers apply fused in (Utils.py:253-295: .mean() for Chamfer / sCD, .max() for
 * Hausdorff): acc4_out (device, 4 doubles) = {sum d2, sum d, max d2, rows}, summed in a fixed order.  d2_out, idx_out nullable:
 * a Chamfer evaluation at 10^9 points needs no per-point output at all. */
int ngpd_nn_sqdist_reduce(const ngpd_grid_t* grid, const float* query, int64_t m, int flags,
                          float* d2_out, int32_t* idx_out, double* acc4_out, void* stream);

/* radius selection: replaces scipy KDTree.query_ball_point(pos, radii) behind Selector.getPointsInRangeSelectionVectorized,
 * Selector.py:214-229 (the Yadav-2018 baseline path).  Row q = every tree point with fp64 squared
 * distance <= radii[q]^2, ascending by index (SciPy's order for multi-point queries).  Two passes:
 *   counting: counts_out[m] != NULL, idx_out == NULL;
 *   filling:  offsets[m+1] (exclusive scan of the counts, int32) and idx_out[offsets[m]].
 * flags: NGPD_KNN_QUERY_IS_TREE as in ngpd_knn. */
int ngpd_ball_query(const ngpd_grid_t* grid, const float* query, int64_t m, const float* radii, int flags,
                    int32_t* counts_out /*nullable*/, const int32_t* offsets /*nullable*/, int32_t* idx_out /*nullable*/, void* stream);

/* ---- neighbourhood kernels.  Row r describes centre point `rows ? rows[r] : r` with neighbours
 * idx[offsets ? offsets[r] .. offsets[r+1] : r*k .. r*k+k) (CSR when offsets != NULL). -------------------- */

/* GraphBuilder.getPVTDecompositionWithKNN, GraphBuilder.py:99-111: covariance of the neighbours about their
 * mean, LAPACK-order eigen-decomposition; normals_out [m,3] = eigenvector of the smallest eigenvalue;
 * eigval_out [m,3] ascending and eigvec_out [m,3,3] (columns) are optional (at least one of normals_out /
 * eigvec_out must be given). */
int ngpd_pca_normals(const float* pos, const int32_t* idx, const int32_t* offsets, int64_t m, int k,
                     float* normals_out /*nullable*/, float* eigval_out /*nullable*/, float* eigvec_out /*nullable*/,
                     void* stream);

/* Decompositionor.getBetterFilteredNVT, Decompositionor.py:278-300.  x_thresh = largest fp32 x with
 * acos(x) > rho.  Outputs: eigval [m,3] ascending, eigvec [m,3,3] (columns), optional tensor [m,3,3]
 * and weight count sumw [m]. */
int ngpd_nvt(const float* pos, const float* nrm, const int32_t* idx, const int32_t* offsets, const int32_t* rows,
             int64_t m, int k, float x_thresh, float* eigval_out, float* eigvec_out,
             float* tensor_out /*nullable*/, int32_t* sumw_out /*nullable*/, void* stream);

/* Decompositionor.getNormalFilteredNVT / getNormalFilteredPVT, Decompositionor.py:260-276, 172-211 (Yadav-2018 baseline):
 * neighbour j of centre i counts iff acos(clamp(ni.nj)) <= rho, passed as x_le = smallest fp32 x with acos(x) <= rho.
 * NVT: T = sum w nj nj^T / sum w (nobody passes: ni ni^T).  PVT: covariance of the passing neighbours about their own
 * mean (nobody passes: all neighbours; empty row: the reference's four-sample surrogate).  Outputs as ngpd_nvt. */
int ngpd_nvt_normal(const float* nrm, const int32_t* idx, const int32_t* offsets, const int32_t* rows, int64_t m, int k,
                    float x_le, float* eigval_out, float* eigvec_out, float* tensor_out /*nullable*/, int32_t* sumw_out /*nullable*/,
                    void* stream);
int ngpd_pvt_normal(const float* pos, const float* nrm, const int32_t* idx, const int32_t* offsets, const int32_t* rows,
                    int64_t m, int k, float x_le, float* eigval_out, float* eigvec_out, float* tensor_out /*nullable*/,
                    int32_t* sumw_out /*nullable*/, void* stream);

/* torch.linalg.eigh on a stack of symmetric 3x3 fp32 tensors (lower triangle read), LAPACK ssyevd order
 * of operations and sign convention (Decompositionor.py:300). */
int ngpd_eigh3(const float* tensors /*[m,3,3]*/, int64_t m, float* eigval_out, float* eigvec_out, void* stream);

/* Decomposition.getVUSmoothedNormals, Decompositionor.py:92-106 */
int ngpd_smooth_normals(const float* eigval, const float* eigvec, const float* nrm, int64_t m, float tau, float damp,
                        float* out, void* stream);

/* Decomposition.getClasses, Decompositionor.py:65-69; features_out nullable [m,3] = planarity, linearity, sphericity */
int ngpd_classify(const float* eigval, int64_t m, float scale, uint8_t* labels_out, float* features_out, void* stream);

/* flat_step's cloud-wide scalars, Denoiser.py:106-107: centre = mean of all gathered neighbours, delta = max
 * distance of a gathered neighbour from it.  out4 = {cx, cy, cz, delta} (device). */
int ngpd_center_delta(const float* pos, const int32_t* idx, int64_t total_neighbours, float* out4, void* stream);

/* Denoiser.{flat,edge,feature,corner}_step, Denoiser.py:26-219.  pos_out [m,3] = new position of each row's
 * centre.  edge_vec [n,3] only for NGPD_STEP_EDGE; center_delta (device, from ngpd_center_delta) only for FLAT. */
#define NGPD_STEP_FLAT 0
#define NGPD_STEP_EDGE 1
#define NGPD_STEP_FEATURE 2
#define NGPD_STEP_CORNER 3
int ngpd_update(int kind, const float* pos, const float* nrm, const float* edge_vec, const int32_t* idx,
                const int32_t* offsets, const int32_t* rows, int64_t m, int k, float alpha, float dmax,
                const float* center_delta, float* pos_out, void* stream);

/* sum of |pos[idx] - pos[row]| over all edges (TorchUtils.averageEdgeLength, Utils.py:298): out2 = {sum, count} fp64 */
int ngpd_edge_length_sum(const float* pos, const int32_t* idx, const int32_t* rows, int64_t m, int k, double* out2, void* stream);

/* ---- consistent normal orientation: GraphBuilder.flipNormals, GraphBuilder.py:129-209.
 * Edge cost 1 - |n_u . n_v| on the given graph (edges are taken as undirected; e directed entries), minimum spanning tree
 * (ties ordered by (cost, lower endpoint, higher endpoint)), signs propagated from the top-most point (max z, made to point
 * up): a child is flipped when n_parent . n_child < flip_threshold (the reference: cos(7 pi / 12)).  nrm is updated in
 * place; components not connected to the root keep their normals, as in the reference.  Synchronises.
 * info_out_host (nullable, 3 ints): connected components of the graph, breadth-first levels, Boruvka rounds. */
int ngpd_orient_normals(const float* pos, float* nrm, int64_t n, const int32_t* edge_src, const int32_t* edge_dst, int64_t e,
                        float flip_threshold, int32_t* info_out_host /*nullable*/, void* stream);

/* ---- mesh vertex update: PatchGeneration.Modules.Mesh.updateVertices(n, k), Mesh.py:377-418 (the Vertex_updating
 * notebook's algorithm).  fp64.  v [nv,3], faces [nf,3] int32, face_normals [nf,3] (the target normals),
 * vertex-triangle adjacency as igl.vertex_triangle_adjacency returns it: vta_faces [sum deg] + vta_offsets [nv+1].
 * `iterations` Jacobi sweeps; v_out must not alias v; scratch (nv*3 doubles) is needed for more than one sweep. */
int ngpd_mesh_vertex_update(const double* v, int64_t nv, const int32_t* faces, const double* face_normals,
                            const int32_t* vta_faces, const int32_t* vta_offsets, int iterations,
                            double* scratch /*nullable*/, double* v_out, void* stream);

/* ---- fused session: Processor.denoise / denoiseUntilMinimumError bodies (Processor.py:124-139,158-176)
 * on tree-ordered float4 state.  The index is frozen on `tree_pos` (what Processor.__init__ saw); pos/nrm are
 * the current state in ORIGINAL point order. */
typedef struct ngpd_step_params {
    int32_t k_feature;       /* 16 */
    int32_t k_update;        /* 8  */
    float x_thresh;          /* acos threshold for rho */
    float tau, damp, scale;  /* .3, 3, .2 */
    int32_t strategy[3];     /* NGPD_STEP_* per label; -1 = leave the class alone */
    float alpha[3];
    float dmax;
    int32_t flags;           /* NGPD_STEP_SNAPSHOT_CLASSES or 0 */
    float clamp_radius;      /* > 0 (with ngpd_session_set_original): a new position is kept only while
                                |x_new - x_original| < clamp_radius, PostProcessing.ipynb#c9 "Ours" / "CPSD" */
} ngpd_step_params_t;
/* flags: every class is moved from the SAME snapshot of the positions (the notebook's temp_pos loop, PostProcessing.ipynb#c9)
 * instead of one class after the other in place (Processor.py:127-138, the default) */
#define NGPD_STEP_SNAPSHOT_CLASSES 1

int ngpd_session_create(const float* tree_pos, int64_t n, int k_hint, void* stream, ngpd_session_t** out);
int ngpd_session_destroy(ngpd_session_t* s);
/* allocate now what a step with this neighbourhood size would allocate on first use (optional) */
int ngpd_session_reserve(ngpd_session_t* s, int k_feature, void* stream);
int ngpd_session_set_state(ngpd_session_t* s, const float* pos, const float* nrm, void* stream);
int ngpd_session_get_state(ngpd_session_t* s, float* pos_out, float* nrm_out, uint8_t* labels_out, void* stream);
/* one iteration: kNN(k_feature) -> NVT -> smooth -> NVT -> classify -> class-sequential update; normals <- smoothed */
int ngpd_session_step(ngpd_session_t* s, const ngpd_step_params_t* p, void* stream);
/* k-NN edge lengths of the current positions incl. the zero self edge (Processor.py:120):
 * out_host = {sum of lengths, edge count}; synchronises */
int ngpd_session_mean_edge_length(ngpd_session_t* s, int k, double* out_host, void* stream);
/* per-kernel device time, measured with CUDA events on the launching stream while profiling is on.
 * Categories: 0 kNN, 1 NVT+smoothing, 2 NVT+labels, 3 flat-step scalars, 4 class updates, 5 halo refreshes + cross-rank
 * scalars (slabs; includes the time spent waiting for the peers).
 * get_profile returns and clears the totals (ms_out[6], launches_out[6]); it waits for the recorded events. */
int ngpd_session_set_profiling(ngpd_session_t* s, int on);
int ngpd_session_get_profile(ngpd_session_t* s, double* ms_out, int32_t* launches_out);
/* kNN of the session.  The index is frozen and only the queries move, so every search also stores 2k candidates, the
 * query position and a certified radius per row; the next pass first re-ranks those candidates (tier 0) and searches
 * the grid only for rows that moved too far (3x3x3 streaming tier, 5x5x5 tier, exact shell search).  All tiers give
 * bit-identical rows.  mode 0 = all tiers (default), 1 = exact shell search for every query, 2 = no re-ranking tier.
 * last_fixups = rows the last pass handed to the exact search; knn_stats out3_host = {rows tier 0 handed to the
 * search, rows the 3x3x3 tier handed on, rows the 5x5x5 tier handed to the exact search}.  Both synchronise. */
int ngpd_session_set_knn_mode(ngpd_session_t* s, int mode);
int ngpd_session_last_fixups(ngpd_session_t* s, void* stream);
int ngpd_session_knn_stats(ngpd_session_t* s, int32_t* out3_host, void* stream);
/* number of kernels the last ngpd_session_step launched */
int ngpd_session_launch_count(const ngpd_session_t* s);
/* tree-order views for tests/benchmarks: perm (sorted -> original) */
int ngpd_session_order(const ngpd_session_t* s, int32_t* perm_out, void* stream);

/* The step split into its dependency phases, so that a multi-GPU driver can refresh halo rows in between
 * (ngpd_session_step is exactly: features 0, features 1, then per class [flat scalars 0, 1] + update, commit).
 *   phase_features part 0: kNN(k_feature) + NVT on the current normals + smoothing  -> smoothed normals
 *                  part 1: NVT on the smoothed normals -> labels + crease directions
 *   phase_flat_scalars (only for a class whose strategy is NGPD_STEP_FLAT, Denoiser.py:106-107)
 *                  part 0: {sum x, sum y, sum z, count} of the class' gathered neighbours -> buffer 3 (4 x int64, fixed point)
 *                  part 1: centre = sums/count, then max distance from it -> buffer 4 (4 floats: centre, delta)
 *                  a multi-GPU driver all-reduces buffer 3 (sum) between the parts and buffer 4[3] (max) after
 *   phase_update: class `key` moves, everything else is copied into the other position buffer
 *   phase_commit_normals: graph.n = f_n (Processor.py:139) */
int ngpd_session_phase_features(ngpd_session_t* s, const ngpd_step_params_t* p, int part, void* stream);
int ngpd_session_phase_flat_scalars(ngpd_session_t* s, const ngpd_step_params_t* p, int key, int part, void* stream);
int ngpd_session_phase_update(ngpd_session_t* s, const ngpd_step_params_t* p, int key, void* stream);
int ngpd_session_phase_commit_normals(ngpd_session_t* s);
/* Morton-slab partitioning: owned[s] != 0 marks the tree-order rows this rank computes; the rest are halo copies
 * (read-only for this rank).  NULL clears the mask. */
int ngpd_session_set_owned(ngpd_session_t* s, const uint8_t* owned_tree_order, void* stream);
/* device views of session state in tree order: 0 positions (float4), 1 normals (float4), 2 smoothed normals
 * (float4), 3 flat-step accumulators (4 x int64: sum x, y, z in fixed point + neighbour count -- exact, so the sum over slabs
 * equals the whole cloud's bit for bit), 4 centre+delta (4 floats), 5 labels (u8), 6 neighbour table,
 * 7 hand-over lists of the last kNN pass (int32: n rows each of tiers 0, 1, 2, then the three counters; diagnostics) */
void* ngpd_session_buffer(ngpd_session_t* s, int which);
/* halo traffic: gather / scatter float4 rows of buffer `which` (0..2) listed by tree position */
int ngpd_session_export_rows(ngpd_session_t* s, int which, const int32_t* rows, int64_t m, float* out4, void* stream);
/* the same gather with the halo exchange fused in: row i of the list is stored directly into the receive buffer of the rank
 * that needs it -- rows [seg[p], seg[p+1]) go to peer p, row seg[p] landing at address peer_base[p] (a float4 slot inside
 * peer p's buffer, mapped into this process: CUDA IPC / symmetric memory over NVLink).  seg (world+1 entries) and peer_base
 * (world entries) are device arrays.  The caller orders the peers' reads after the stores (a cross-rank barrier on the stream). */
int ngpd_session_export_rows_peers(ngpd_session_t* s, int which, const int32_t* rows, int64_t m, const int64_t* seg,
                                   const uint64_t* peer_base, int world, void* stream);
int ngpd_session_import_rows(ngpd_session_t* s, int which, const int32_t* rows, int64_t m, const float* in4, void* stream);

/* ---- Morton slabs on several GPUs (new work, the reference is single-process).  One session per rank over its
 * owned rows + halo copies (ngpd_session_set_owned).  Halo VALUES travel by peer stores: every rank owns one block of
 * symmetric memory (ngpd_slab_symm_bytes, zero-filled, the same size on every rank, mapped into all peers -- e.g. torch
 * symmetric memory over NVLink / NVSwitch) and the kernels of a refresh write the rows straight into the peers' blocks,
 * signal with a release store and wait with acquire loads: no NCCL call, no host round trip.  The cloud-wide scalars of
 * flat_step (Denoiser.py:106-107) are reduced through the same block, summed in rank order (identical bits on every rank).
 * Every rank must issue the same sequence of refresh / allreduce / step_slab calls. */
typedef struct ngpd_slab_wiring {
    int32_t world, rank;
    int64_t n_send, n_recv, cap;     /* cap = rows per receive buffer (>= n_recv of every rank) */
    const int32_t* send_rows;        /* device, n_send: tree positions of the owned rows peers hold copies of, grouped by destination rank */
    const int32_t* recv_rows;        /* device, n_recv: tree positions of this rank's halo rows, grouped by source rank, in arrival order */
    const int64_t* send_seg_host;    /* world + 1: rows [seg[q], seg[q+1]) of send_rows go to rank q */
    const int64_t* first_row_host;   /* world: row of rank q's receive buffer where this rank's block starts */
    const uint64_t* symm_base_host;  /* world: address of rank q's symmetric block in this process */
} ngpd_slab_wiring_t;
int64_t ngpd_slab_symm_bytes(int world, int64_t cap);
/* send_rows / recv_rows must stay alive while the wiring is set; NULL wiring clears it */
int ngpd_session_set_slab(ngpd_session_t* s, const ngpd_slab_wiring_t* w, void* stream);
int ngpd_session_slab_refresh(ngpd_session_t* s, int which, void* stream);      /* which: 0 positions, 1 normals, 2 smoothed normals */
int ngpd_session_slab_allreduce(ngpd_session_t* s, int mode, void* stream);     /* 0: buffer 3 <- sum, 1: buffer 4 [3] <- max */
/* ngpd_session_step with the halo refreshes and cross-rank scalars in between, driven from C (one call per iteration) */
int ngpd_session_step_slab(ngpd_session_t* s, const ngpd_step_params_t* p, void* stream);
/* largest (k-th neighbour distance + distance of the query from its tree position) any owned row has shown: the slab's
 * searches equal the whole cloud's iff it is below the halo width.  Synchronises. */
int ngpd_session_halo_need(ngpd_session_t* s, int reset, float* out_host, void* stream);
/* order- and partition-independent digest of the owned rows: out11_host = {hash(id, position bits), hash(id, normal bits),
 * sum x, sum y, sum z (int64, units of 2^-24), sum |n| (same units), label counts 0 / 1 / 2 / other, rows}.  global_ids
 * (nullable device int64 [n]): id of the session's i-th point in the whole cloud.  Synchronises. */
int ngpd_session_checksum(ngpd_session_t* s, const int64_t* global_ids, uint64_t* out11_host, void* stream);
/* positions a clamped run measures its displacement from (ngpd_step_params.clamp_radius); NULL drops them */
int ngpd_session_set_original(ngpd_session_t* s, const float* pos, void* stream);

/* End-to-end entry points with HOST buffers (the path the e2e measurement times): copy this step's positions and
 * normals in, run `iterations` steps, copy positions / normals / labels back.  Synchronous. */
int ngpd_session_run_host(ngpd_session_t* s, const ngpd_step_params_t* p, int iterations, const float* pos_host,
                          const float* nrm_host, float* pos_out_host, float* nrm_out_host, uint8_t* labels_out_host,
                          void* stream);
int ngpd_denoise_host(const float* tree_pos_host, const float* pos_host, const float* nrm_host, int64_t n,
                      const ngpd_step_params_t* p, int iterations, float* pos_out_host, float* nrm_out_host,
                      uint8_t* labels_out_host);

#ifdef __cplusplus
}
#endif
#endif /* NGPD_H */
