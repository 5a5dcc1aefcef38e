/*
 * meshcnn_b200 -- C ABI of the B200 (sm_100a) implementation of MeshCNN's per-edge hot path.
 *
 * The reference (ranahanocka/MeshCNN) is pure Python and has no FFI; the boundary it offers is the Python
 * class surface of models/layers/{mesh_conv,mesh_pool,mesh_union,mesh_unpool,mesh}.py.  The functions below
 * are what a binding for that surface needs; each one cites the reference code it replaces.  The Python
 * mirror (meshcnn_b200/*.py) binds them with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; sizes are element counts
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *   - no allocation, no global mutable state (the per-device cache of configured kernel attributes aside); safe to call
 *     from any host thread and under CUDA-graph capture.  The process-global diagnostics switches (mcnn_debug_*) are NOT
 *     part of this ABI: they are declared in meshcnn_b200_debug.h
 *   - return value: 0 = launched, otherwise a MCNN_E_* code (argument errors are detected on the host;
 *     data-dependent errors of the collapse kernel are written to its per-mesh `status` array)
 *
 * Layouts (see DESIGN.md "Data layout in HBM")
 *   features   x[B][E][C]  fp32, edge-major (the channel vector of one edge is contiguous) -- this is the
 *              channels_last view of the reference's [B, C, E, 1] tensors
 *   topology   gemm[B][E][4] int32 (one-ring neighbour ids, -1 = none), sides[B][E][4] int8,
 *              ec[B] int32 (live edge count of every mesh; rows >= ec[b] are padding)
 *   groups     CSR  (rowptr[B][T+1], cols[B][cap])  pooled edge -> old edges   (clamped MeshUnion rows)
 *              CSC  (colptr[B][E+1], rows[B][cap])  old edge -> pooled edges
 *              occ[B][E] fp32  un-clamped, un-masked column sums
 */
#ifndef MESHCNN_B200_H
#define MESHCNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCNN_OK 0
#define MCNN_E_BADARG 1          /* null pointer / non-positive size / unsupported shape               */
#define MCNN_E_TOO_LARGE 2       /* mesh does not fit the on-chip collapse state and no workspace was given */
#define MCNN_E_CUDA 3            /* a CUDA runtime call failed (see mcnn_last_cuda_error)               */

/* arithmetic of the tensor-core MeshConv kernels (the `_p` entry points).
 *   MCNN_PREC_FP32  every product as three TF32 MMAs (a_hi b_hi + a_hi b_lo + a_lo b_hi): fp32 accuracy (~1e-6 relative) -- what the
 *                   entry points without `_p` do, and the default of the Python layers
 *   MCNN_PREC_TF32  one TF32 MMA per product (operands rounded to 10 mantissa bits, fp32 accumulate): ~5e-4 relative, inside
 *                   the 2e-2 tolerance BASELINE.json states for reduced-precision activations; a third of the tensor work and
 *                   half of the operand traffic.  Opt-in only. */
#define MCNN_PREC_FP32 0
#define MCNN_PREC_TF32 1

/* per-mesh status words written by mcnn_pool_collapse (0 = fine).  The Python mirror turns them into the
 * exception the reference would raise (failure row). */
#define MCNN_ST_OK 0
#define MCNN_ST_QUEUE_EMPTY 1    /* heappop on an empty heap -> IndexError      (mesh_pool.py:49-50)    */
#define MCNN_ST_SHARED_ASSERT 2  /* assert len(shared_items) == 2               (mesh_pool.py:123)      */
#define MCNN_ST_VERTEX_ASSERT 3  /* assert len(vertex) == 1                     (mesh_pool.py:181)      */
#define MCNN_ST_VE_REMOVE 4      /* list.remove(x): x not in list -> ValueError (mesh.py:48)            */
#define MCNN_ST_GROUP_OVERFLOW 5 /* sparse MeshUnion node pool exhausted (implementation limit)         */
#define MCNN_ST_DEGENERATE 6     /* edge with identical endpoints reached merge_vertices (unsupported)  */
#define MCNN_ST_NNZ_OVERFLOW 7   /* groups CSR capacity exceeded (implementation limit)                 */

/* per-pop outcome codes of the optional collapse log */
#define MCNN_POP_SKIP_MASKED 0
#define MCNN_POP_REJ_BOUNDARY 1
#define MCNN_POP_REJ_SIDE0 2
#define MCNN_POP_REJ_SIDE2 3
#define MCNN_POP_REJ_ONE_RING 4
#define MCNN_POP_COLLAPSED 5

int mcnn_version(void);
/* number of kernel launches this library has enqueued since load (monotonic; bench.py reports its delta) */
unsigned long long mcnn_launch_count(void);
const char* mcnn_last_cuda_error(void);

/* ---------------------------------------------------------------------------------------------------------
 * MeshConv  (models/layers/mesh_conv.py)
 * ------------------------------------------------------------------------------------------------------ */

/* Fused pad_gemm + create_GeMM + Conv2d(1x5)  (mesh_conv.py:20-26, :39-70, :72-84).
 *   x      [B][E][Cin]     weight [Cout][Cin][1][5] (nn.Conv2d layout)     bias [Cout] or NULL
 *   out    [B][E][Cout]
 *   wt_ws  [5*Cin][Cout] scratch: the weight re-laid K-major by a first small kernel
 * Rows e >= ec[b] gather edge 0 of mesh b for all five slots; -1 neighbours read zeros. */
int mcnn_meshconv_fwd(const float* x, const int32_t* gemm, const int32_t* ec, const float* weight,
                      const float* bias, float* out, float* wt_ws, int B, int E, int Cin, int Cout, void* stream);

/* Same contract as mcnn_meshconv_fwd, computed on the tensor cores (tcgen05.mma kind::tf32 with the 3xTF32 split, fp32
 * accumulation in TMEM; results agree with the FFMA path to ~1e-6 relative).  Only for shapes where
 * mcnn_meshconv_tc_supported(Cin, Cout) != 0 (Cin % 32 == 0, Cout % 64 == 0).
 *   wt_ws  2*5*Cin*Cout floats of scratch: a first small kernel writes the weight there as the ready-to-use B operand
 *          (TF32 head + remainder, K-major SWIZZLE_128B atoms), which every CTA then fetches with bulk copies */
int mcnn_meshconv_tc_supported(int Cin, int Cout);
/* The two weight images (each 2*5*Cin*Cout floats) used by mcnn_meshconv_fwd_tc* / mcnn_meshconv_bwd_data_tc*, built ahead of
 * time (e.g. on a side stream at the start of a step); those calls then take weight = NULL and the image as wt_ws / wtb_ws. */
int mcnn_meshconv_tc_weight_images(const float* weight, float* fwd_img, float* bwd_img, int Cin, int Cout, void* stream);
int mcnn_meshconv_fwd_tc(const float* x, const int32_t* gemm, const int32_t* ec, const float* weight,
                         const float* bias, float* out, float* wt_ws, int B, int E, int Cin, int Cout, void* stream);
/* the same, additionally writing signs[2][B*E][Cin] (int8: sign(f1 - f3), then sign(f2 - f4), the derivative of the two abs
 * terms of mesh_conv.py:67-68) for mcnn_meshconv_bwd_data_tc_signs; signs may be NULL */
int mcnn_meshconv_fwd_tc_signs(const float* x, const int32_t* gemm, const int32_t* ec, const float* weight,
                               const float* bias, float* out, float* wt_ws, int8_t* signs, int B, int E, int Cin, int Cout,
                               void* stream);
/* the same computation as mcnn_meshconv_fwd_tc_signs by a persistent kernel: the (128-edge tile, 32-wide K block) units
 * are spread evenly over at most one CTA per SM ("stream-K"), so a tile count just above a multiple of the SM count does
 * not pay for a nearly empty last wave.  A tile cut between CTAs is summed in a fixed order (deterministic; it differs from
 * the one-CTA-per-tile kernel by fp32 rounding of that sum only).  sk_ws: mcnn_meshconv_fwd_tc_sk_ws() bytes of device memory,
 * zeroed ONCE by the caller after allocation (the kernel leaves it reusable), not shared by concurrently running launches. */
size_t mcnn_meshconv_fwd_tc_sk_ws(void);
/* the same with the arithmetic selectable (MCNN_PREC_*); likewise for the two gradients below */
int mcnn_meshconv_fwd_tc_sk_p(const float* x, const int32_t* gemm, const int32_t* ec, const float* weight,
                              const float* bias, float* out, float* wt_ws, int8_t* signs, void* sk_ws, int B, int E, int Cin,
                              int Cout, int precision, void* stream);
int mcnn_meshconv_bwd_data_tc_signs_p(const float* dout, const float* x, const int8_t* signs, const int32_t* gemm,
                                      const int32_t* ec, const float* weight, float* dx, float* wtb_ws, int B, int E,
                                      int Cin, int Cout, int precision, void* stream);
int mcnn_meshconv_bwd_weight_tc_p(const float* dout, const float* x, const int32_t* gemm, const int32_t* ec,
                                  float* dweight, float* dbias, float* ws, int B, int E, int Cin, int Cout,
                                  int precision, void* stream);
int mcnn_meshconv_fwd_tc_sk(const float* x, const int32_t* gemm, const int32_t* ec, const float* weight,
                            const float* bias, float* out, float* wt_ws, int8_t* signs, void* sk_ws, int B, int E, int Cin,
                            int Cout, void* stream);


/* dx of the above (autograd of mesh_conv.py:47-69: conv2d dgrad, abs/add backward, index_select scatter).
 *   dout [B][E][Cout]   dx [B][E][Cin]  -- must be ZERO on entry (the kernel accumulates with red.global) */
int mcnn_meshconv_bwd_data(const float* dout, const float* x, const int32_t* gemm, const int32_t* ec,
                           const float* weight, float* dx, int B, int E, int Cin, int Cout, void* stream);

/* dW [Cout][Cin][1][5] and dbias [Cout] (NULL to skip); both must be ZERO on entry. */
int mcnn_meshconv_bwd_weight(const float* dout, const float* x, const int32_t* gemm, const int32_t* ec,
                             float* dweight, float* dbias, int B, int E, int Cin, int Cout, void* stream);

/* Tensor-core versions of the two backward kernels (tcgen05.mma kind::tf32, 3xTF32 split, TMEM accumulators).
 *   bwd_data_tc  : Cin % 32 == 0 and Cout % 32 == 0; dx must be ZERO on entry (red.global.add.v4.f32 scatter);
 *                  wtb_ws 2*5*Cin*Cout floats of scratch for the B-operand image of the weight (head + remainder)
 *   bwd_weight_tc: Cin % 32 == 0 and Cout % 4 == 0; dweight / dbias are OVERWRITTEN (slab partials are written to
 *                  `ws` and reduced by a second kernel: deterministic, no atomics); ws holds
 *                  mcnn_meshconv_bwd_weight_tc_ws(B, E, Cin, Cout) floats */
/* mcnn_meshconv_bwd_data_tc with the signs kept by the forward kernel instead of re-reading four neighbour rows of x per edge
 * (x may then be NULL) */
int mcnn_meshconv_bwd_data_tc_signs(const float* dout, const float* x, const int8_t* signs, const int32_t* gemm,
                                    const int32_t* ec, const float* weight, float* dx, float* wtb_ws, int B, int E, int Cin,
                                    int Cout, void* stream);
int mcnn_meshconv_bwd_data_tc(const float* dout, const float* x, const int32_t* gemm, const int32_t* ec,
                              const float* weight, float* dx, float* wtb_ws, int B, int E, int Cin, int Cout,
                              void* stream);
size_t mcnn_meshconv_bwd_weight_tc_ws(int B, int E, int Cin, int Cout);
int mcnn_meshconv_bwd_weight_tc(const float* dout, const float* x, const int32_t* gemm, const int32_t* ec,
                                float* dweight, float* dbias, float* ws, int B, int E, int Cin, int Cout, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * MeshPool / MeshUnion / Mesh.clean  (models/layers/mesh_pool.py, mesh_union.py, mesh.py:26-71,151-198)
 * ------------------------------------------------------------------------------------------------------ */

/* __build_queue's key: norms[b][e] = sum_c x[b][e][c]^2 in fp32, channels added in ascending order with a
 * separate multiply and add (mesh_pool.py:186).  norms [B][E]. */
int mcnn_pool_norms(const float* x, float* norms, int B, int E, int C, void* stream);

typedef struct {
    /* sizes */
    int32_t B;            /* meshes in the batch                                                        */
    int32_t E_in;         /* edge dimension of the incoming level (rows of gemm_in per mesh)            */
    int32_t T;            /* pool target = edge dimension of the outgoing level                         */
    int32_t V;            /* vertex capacity per mesh                                                   */
    int32_t ve_cap;       /* capacity of ve_idx per mesh (2 * edges at level 0 suffices)                */
    int32_t nnz_cap;      /* capacity of grp_cols / grp_rows per mesh                                   */
    int32_t log_cap;      /* capacity of the optional logs per mesh (>= E_in for pops, unions)          */
    int32_t reserved;
    /* incoming level (read only) */
    const int32_t* gemm_in;   /* [B][E_in][4] */
    const int8_t* sides_in;   /* [B][E_in][4] */
    const int32_t* ec_in;     /* [B]          */
    const float* norms;       /* [B][E_in]    */
    /* state carried across levels (read + overwritten with the cleaned state, Mesh.clean mesh.py:50-71) */
    int32_t* edges;           /* [B][E_cap0][2] vertex ids per edge; row stride given by edges_stride    */
    int32_t edges_stride;     /* rows per mesh in `edges`                                                */
    int32_t reserved2;
    int32_t* ve_ptr;          /* [B][V+1]  CSR of the per-vertex edge lists (a multiset with stale ids)  */
    int32_t* ve_idx;          /* [B][ve_cap] */
    uint8_t* v_mask;          /* [B][V]      */
    double* vs;               /* [B][V][3] or NULL (positions are only needed for export, mesh.py:29-33) */
    /* outgoing level */
    int32_t* gemm_out;        /* [B][T][4] */
    int8_t* sides_out;        /* [B][T][4] */
    int32_t* ec_out;          /* [B]       */
    /* MeshUnion result */
    int32_t* grp_rowptr;      /* [B][T+1]     */
    int32_t* grp_cols;        /* [B][nnz_cap] */
    int32_t* grp_colptr;      /* [B][E_in+1]  */
    int32_t* grp_rows;        /* [B][nnz_cap] */
    float* occ;               /* [B][E_in]    */
    int32_t* status;          /* [B]          */
    /* optional logs (NULL to skip) */
    int32_t* log_npops;       /* [B]             */
    int32_t* log_pops;        /* [B][log_cap]    */
    int8_t* log_outcome;      /* [B][log_cap]    */
    int32_t* log_nunions;     /* [B]             */
    int32_t* log_unions;      /* [B][8*log_cap][2] pairs (src, dst)                                      */
    int32_t* edge_mask_out;   /* [B][E_in] survivor mask (1/0) or NULL                                   */
    /* large meshes: when the collapse state of one mesh does not fit the shared memory of an SM (mcnn_pool_smem_bytes >
     * mcnn_pool_smem_limit) the same kernel runs with 32-bit ids on this global-memory workspace, one slice per mesh */
    void* workspace;          /* [B][workspace_stride] bytes or NULL                                     */
    long long workspace_stride; /* >= mcnn_pool_workspace_bytes(E_in, V, ve_cap, with_vs)                */
    int32_t force_workspace;  /* != 0: use the workspace variant even if the mesh would fit on chip (tests) */
    int32_t reserved3;
    long long* dbg_clocks;    /* [B][24] SM clock at the phase boundaries of each mesh's CTA (diagnostics) or NULL:
                                 0 start, 1 loaded, 2 sorted, 3 collapse done, 4 unions applied, 5 topology written, 6 end;
                                 7 = pops; 8..18 = cycles per collapse stage (only in -DMCNN_POOL_PROF builds)          */
} mcnn_pool_args;

/* __pool_main for every mesh of the batch, one CTA per mesh (mesh_pool.py:41-203 + mesh.py:26-71 +
 * mesh_union.py:9-25): sort by (norm, id), sequential collapse with all the reference's validity tests and
 * side effects, Mesh.clean compaction / re-indexing, MeshUnion rows as CSR + CSC + occurrences. */
int mcnn_pool_collapse(const mcnn_pool_args* args, void* stream);
/* dynamic shared memory the collapse kernel needs for these sizes (bytes); > mcnn_pool_smem_limit() means
 * MCNN_E_TOO_LARGE */
size_t mcnn_pool_smem_bytes(int E_in, int V, int ve_cap, int with_vs);
size_t mcnn_pool_smem_limit(void);
/* bytes of global-memory workspace PER MESH for the large-mesh variant */
size_t mcnn_pool_workspace_bytes(int E_in, int V, int ve_cap, int with_vs);

/* rebuild_features_average (mesh_union.py:27-36):  out[b][r] = mean_{j in row r} x[b][j],  rows >= ec_out = 0
 *   x [B][E_in][C]   out [B][T][C] */
int mcnn_segmean_fwd(const float* x, const int32_t* grp_rowptr, const int32_t* grp_cols, const int32_t* ec_out,
                     float* out, int B, int E_in, int T, int C, int nnz_cap, void* stream);
/* its transpose: dx[b][j] = sum_{r containing j} dout[b][r] / len(r) */
int mcnn_segmean_bwd(const float* dout, const int32_t* grp_rowptr, const int32_t* grp_colptr,
                     const int32_t* grp_rows, const int32_t* ec_in, float* dx, int B, int E_in, int T, int C,
                     int nnz_cap, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * MeshUnpool  (models/layers/mesh_unpool.py:30-41)
 * ------------------------------------------------------------------------------------------------------ */
/* out[b][i] = (sum_{r containing i} f[b][r]) / occ[b][i] for i < ec_hi[b]; 0 for i >= ec_hi[b]
 *   f [B][E_lo][C]   out [B][E_hi][C];  the CSC/occ are the ones of the pool level being undone, whose
 *   E_in was E_grp (E_hi >= E_grp; the extra columns are the reference's zero padding)                    */
int mcnn_unpool_fwd(const float* f, const int32_t* grp_colptr, const int32_t* grp_rows, const float* occ,
                    const int32_t* ec_hi, float* out, int B, int E_lo, int E_hi, int E_grp, int C, int nnz_cap,
                    void* stream);
/* df[b][r] = sum_{i in row r} dout[b][i] / occ[b][i]; rows r >= ec_lo[b] = 0 */
int mcnn_unpool_bwd(const float* dout, const int32_t* grp_rowptr, const int32_t* grp_cols, const float* occ,
                    const int32_t* ec_lo, float* df, int B, int E_lo, int E_hi, int E_grp, int C, int nnz_cap,
                    void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fused normalisation blocks around MeshConv (reference: models/networks.py:149,
 * :171-179 (BatchNorm2d in MResConv, GroupNorm after it), :222-233 / :276-288 (InstanceNorm2d + residual + relu)).
 *     z = pre(x [+ a]);   y = post(gamma * (z - mean) * rstd + beta [+ r])          pre / post = relu or identity
 * x, a, r, y: [nblocks * rows_per_block][C] edge-major; statistics per (block of rows) x (group of C/G channels):
 *     BatchNorm2d: nblocks = 1, rows = B*E, G = C;  GroupNorm: nblocks = B, rows = E, G = groups;  InstanceNorm2d: G = C.
 * mean / rstd: [nblocks][G] (written unless stats_given); part: mcnn_norm_ws() floats of scratch; running_* (NULL or
 * [C]) get the BatchNorm momentum update.  Backward returns dx (== da), dr (NULL to skip), dgamma / dbeta (NULL ok);
 * s1 / s2: [nblocks][G] scratch.
 * ------------------------------------------------------------------------------------------------------ */
size_t mcnn_norm_ws(int nblocks, int rows_per_block, int C, int G);
/* The stages on their own, for statistics that cross processes (SyncBN for MResConv.bn, networks.py:167): the
 * caller reduces the slab partials `part` [nblocks][slabs][2][C] (slabs = mcnn_norm_slabs; [0] = sum z / sum dyh, [1] = sum z^2 /
 * sum dyh*xhat per channel), all-reduces them, and hands mean / rstd (forward: mcnn_norm_fwd with stats_given = 1) or
 * s1 = gamma*A/n, s2 = gamma*B/n per (block, group) (backward) back in. */
int mcnn_norm_slabs(int nblocks, int rows_per_block, int C, int G);
int mcnn_norm_stats(const float* x, const float* a, float* part, int nblocks, int rows_per_block, int C, int G, int pre_relu,
                    void* stream);
int mcnn_norm_bwd_stats(const float* dy, const float* y, const float* x, const float* a, const float* mean, const float* rstd,
                        float* part, int nblocks, int rows_per_block, int C, int G, int pre_relu, int post_relu, void* stream);
int mcnn_norm_bwd_apply(const float* dy, const float* y, const float* x, const float* a, const float* gamma, const float* mean,
                        const float* rstd, const float* s1, const float* s2, float* dx, float* dr, int nblocks,
                        int rows_per_block, int C, int G, int pre_relu, int post_relu, void* stream);
int mcnn_norm_fwd(const float* x, const float* a, const float* r, const float* gamma, const float* beta, float* y,
                  float* mean, float* rstd, float* part, float* running_mean, float* running_var, float momentum,
                  float eps, int nblocks, int rows_per_block, int C, int G, int pre_relu, int post_relu, int stats_given,
                  void* stream);
int mcnn_norm_bwd(const float* dy, const float* y, const float* x, const float* a, const float* gamma, const float* mean,
                  const float* rstd, float* dx, float* dr, float* dgamma, float* dbeta, float* part, float* s1, float* s2,
                  int nblocks, int rows_per_block, int C, int G, int pre_relu, int post_relu, void* stream);
/* mcnn_norm_bwd with dx = (gradient through the block) + dacc (dacc NULL: identical to mcnn_norm_bwd): x also feeds another
 * branch -- the residual of MResConv, networks.py:171-179 -- whose gradient is already known; saves the separate add. */
int mcnn_norm_bwd_acc(const float* dy, const float* y, const float* x, const float* a, const float* gamma, const float* mean,
                      const float* rstd, const float* dacc, float* dx, float* dr, float* dgamma, float* dbeta, float* part,
                      float* s1, float* s2, int nblocks, int rows_per_block, int C, int G, int pre_relu, int post_relu,
                      void* stream);

/* torch.optim.Adam step (no amsgrad) over flat fp32 buffers: parameters, gradients, exp_avg, exp_avg_sq (mesh_classifier.py:38) */
int mcnn_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int step, void* stream);

/* Same update with hyper-parameters and step counter in DEVICE memory, so that the launch can be captured in a CUDA graph
 * and replayed:  state[8] = {lr, beta1, beta2, eps, weight_decay, step, 1 - beta1^step, sqrt(1 - beta2^step)}; the call
 * first advances step (and the two bias corrections), then applies the update. */
int mcnn_adam_flat_state(float* p, const float* g, float* m, float* v, long long n, float* state, void* stream);

/* ----------------------------------------------------------------------------------------------------------
 * Host-side topology builder (no device work): the sequential parts of the reference's mesh
 * preparation, models/layers/mesh_prepare.py:90-113 (remove_non_manifolds) and :116-161 (build_gemm).  All pointers are HOST
 * memory.  faces [nf][3] int64; face_areas [nf] double.
 *   mcnn_host_drop_non_manifold: keep[f] = 1 for faces that survive (non-zero area, no directed edge already used by an
 *     accepted face); returns the number kept.
 *   mcnn_host_build_gemm: edges in order of first appearance; outputs sized for E = 3 nf: edges [3nf][2] int32, gemm / sides
 *     [3nf][4] int64 (-1 = boundary), ve_ptr [nv + 1] / ve_idx [6 nf] int32 (vertex -> incident edges, creation order),
 *     edge_area_sum [3nf] double (face_area / 3 accumulated in face order).  Returns E; -100: an edge with more than two faces.
 * ---------------------------------------------------------------------------------------------------------- */
/*   mcnn_host_edge_features: the five input features of mesh_prepare.py:303-427 up to their arccos -- edge lengths [E], the
 *     clipped cosine of the dihedral angle [E], the clipped cosines of the two opposite angles [2][E], the two height / base
 *     ratios [2][E] -- same operations in the same order as the NumPy formulation (bit-identical; arccos stays in NumPy).
 *     Returns 0; -101 where NumPy's errstate(divide='raise') would raise (the reference's "bad features"). */
int mcnn_host_edge_features(const double* vs, int nv, const int32_t* edges, const int64_t* gemm, int E, double* lengths,
                            double* dih_arg, double* opp_arg, double* ratio);
int mcnn_host_drop_non_manifold(const int64_t* faces, const double* face_areas, int nf, int nv, uint8_t* keep);
int mcnn_host_build_gemm(const int64_t* faces, const double* face_areas, int nf, int nv, int32_t* edges, int64_t* gemm,
                         int64_t* sides, int32_t* ve_ptr, int32_t* ve_idx, double* edge_area_sum);

/* ----------------------------------------------------------------------------------------------------------
 * Classifier head of MeshConvNet (models/networks.py:143-157): AvgPool1d over the edge axis, fc1 + relu, fc2.
 *     pooled[b][c] = mean_e x[b][e][c];  h = relu(W1 pooled + b1);  logits = W2 h + b2
 * x: [B][E][C] edge-major (the whole, zero-padded edge axis is averaged, as AvgPool1d(res[-1]) does); W1: [H][C], W2: [K][H]
 * (torch.nn.Linear layout); pooled [B][C] and h [B][H] are kept for the backward.  Backward: dx (NULL to skip), dW1, db1
 * (NULL ok), dW2, db2 (NULL ok) are OVERWRITTEN; dh: [B][H] scratch.  One forward and two backward launches, deterministic.
 * ---------------------------------------------------------------------------------------------------------- */
int mcnn_head_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, float* pooled, float* h,
                  float* logits, int B, int E, int C, int H, int K, void* stream);
int mcnn_head_bwd(const float* dlogits, const float* pooled, const float* h, const float* w1, const float* w2, float* dh,
                  float* dx, float* dw1, float* db1, float* dw2, float* db2, int B, int E, int C, int H, int K, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MESHCNN_B200_H */
