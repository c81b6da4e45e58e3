/*
 * mgb200 - B200-native multigrid solve phase behind the API of JuliaInv/Multigrid.jl.
 *
 * C ABI of libmgb200.so: plain pointers and sizes, no C++/torch types.  The entry points
 * take exactly what the reference's own FFI style passes (ccall with the raw arrays of a
 * Julia SparseMatrixCSC: colptr / rowval / nzval, 1-based Int64 indices, dense column-major
 * n x nrhs right-hand sides; cf. src/Multigrid/parRelax.jl:61-79, Vanka.jl:436-452,
 * src/ParallelJuliaSolver/parallelJuliaSolver.jl:214-238 of the reference).
 *
 * Storage convention (src/Multigrid/MGdef.jl:75-77, SpMatMul.jl:4-13): the hierarchy holds
 * the ADJOINT matrices  As[l] = A_l^H,  Ps[l] = P_l^T,  Rs[l] = R_l^T  in CSC, so the CSC
 * arrays are the CSR arrays of conj(operator).  Upload converts once: indices - index_base,
 * Int64 -> Int32, values conjugated.  Host arrays are only read during the call (Julia may
 * move them afterwards); the device owns its copies until mgb200_destroy.
 *
 * Levels are numbered 1..levels as in the reference.  All functions return 0 on success and a
 * negative status otherwise (-1 bad argument / state, -2 CUDA error, -3 NCCL error);
 * mgb200_last_error() gives the message of the last failure on the calling thread.
 *
 * One handle drives one GPU on one stream and is not re-entrant.  Multi-GPU runs use one
 * process (or Julia worker) per GPU; see mgb200_dist_* below.
 */
#ifndef MGB200_H
#define MGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mgb200_hierarchy* mgb200_handle;

/* value types (VAL of MGparam{VAL,IND}, src/Multigrid/MGdef.jl:91) */
#define MGB200_FP64 0
#define MGB200_CFP64 1
#define MGB200_FP32 2   /* Float32 / ComplexF32 hierarchies (getMGparam(Float32,...): `singlePrecision`, MGdef.jl:119,151;   */
#define MGB200_CFP32 3  /* MGsetup.jl:31-33,79-82,108-110): every array of the handle, b and x included, is single precision */

/* relaxation kinds */
#define MGB200_RELAX_DIAG 0     /* "Jac" and "SPAI": x += d .* r (MGcycle.jl:122-136)          */
#define MGB200_RELAX_JACGMRES 1 /* "Jac-GMRES": FGMRES_relaxation with D as preconditioner      */

const char* mgb200_last_error(void);
int mgb200_version(void);

/* getMGparam + the device part of MGsetup (MGdef.jl:149-161).  relax_pre/relax_post hold
 * param.relaxPre(l), param.relaxPost(l) for l = 1..levels.  cycle_type is 'V','F','W' or 'K'. */
int mgb200_create(mgb200_handle* h, int val_type, int levels, int nrhs, char cycle_type,
                  int relax_kind, const int64_t* relax_pre, const int64_t* relax_post, int device);

/* clear! (MGdef.jl:179-189) */
int mgb200_destroy(mgb200_handle h);

/* Upload level l < levels of the hierarchy: As[l], Ps[l], Rs[l], relaxPrecs[l].
 *   n   = size(As[l],2) rows of A_l;  nc = rows of A_{l+1}
 *   A:  colptr[n+1],  rowval, nzval (VAL)             CSC of A_l^H   (n x n)
 *   P:  colptr[n+1],  rowval, nzval (real(VAL))       CSC of P_l^T   (nc x n)
 *   R:  colptr[nc+1], rowval, nzval (real(VAL))       CSC of R_l^T   (n x nc)
 *   d:  relaxPrecs[l] (VAL, length n)
 * index_base is 1 for Julia arrays, 0 for C/NumPy arrays. */
int mgb200_upload_level(mgb200_handle h, int level, int64_t n, int64_t nc,
                        const int64_t* a_colptr, const int64_t* a_rowval, const void* a_nzval,
                        const int64_t* p_colptr, const int64_t* p_rowval, const void* p_nzval,
                        const int64_t* r_colptr, const int64_t* r_rowval, const void* r_nzval,
                        const void* d, int index_base);

/* Optional grid hint for the transfer operators of level l, BEFORE mgb200_upload_level(l): the nodes per dimension of
 * the meshes of level l and l+1 (param.Meshes[l].n .+ 1, param.Meshes[l+1].n .+ 1; dim = 1, 2 or 3).  If the uploaded
 * Ps[l] / Rs[l] are exactly the operators those grids imply (full-weighting pair of GeometricTransferOperators.jl on
 * nodal grids, every dimension coarsened; checked row by row at upload, the values are taken from the matrices) and
 * the option "grid_transfers" is on, restriction and prolongation run kernels that read no matrix stream at all
 * (csrc/grid_xfer.cuh); otherwise the hint is ignored.  Results are bit-identical either way. */
int mgb200_set_level_grid(mgb200_handle h, int level, int dim, const int64_t* n_fine_nodes, const int64_t* n_coarse_nodes);

/* defineCoarsestAinv (MGsetup.jl:323-355, default branch): As[end] = A_L^H in CSC; densified,
 * LU-factorised with partial pivoting and prepared for solves on the device. */
int mgb200_upload_coarsest(mgb200_handle h, int64_t n, const int64_t* colptr, const int64_t* rowval,
                           const void* nzval, int index_base);

/* defineCoarsestAinv with coarseSolveType == "GMRES" (MGsetup.jl:333-334, solveCoarsest MGcycle.jl:152-168):
 * As[end] in CSC and d = param.LU = conj(relaxParam ./ diag(As[end])) (VAL, length n).  The coarsest solve is then
 * x = 0; one restart of KrylovMethods.fgmres(10), tol 0.01, right-preconditioned by d.*v (nrhs = 1). */
int mgb200_upload_coarsest_gmres(mgb200_handle h, int64_t n, const int64_t* colptr, const int64_t* rowval,
                                 const void* nzval, const void* d, int index_base);

/* replaceMatrixInHierarchy (MGsetup.jl:226-270) on the device: the fine matrix changes, Ps / Rs stay.  The new
 * As[1] = A^H comes in CSC like in mgb200_upload_level; per level the relaxation weights (getRelaxPrec, MGsetup.jl:142-160:
 * relax_type 0 = "Jac" / "Jac-GMRES", 1 = "SPAI"; relax_param[l] = relaxParam of level l + 1, `levels` entries, NULL = 1.0),
 * the Galerkin product Ps[l]*AT*Rs[l] (numeric only: into the sparsity the first setup produced, csrc/galerkin.cuh) and
 * the coarsest factorisation (defineCoarsestAinv) are redone on the device.  *done = 1: the hierarchy now belongs to
 * the new matrix.  *done = 0: the device path does not apply (the new matrix has another sparsity than the resident
 * one, the hierarchy is row-partitioned, or SPAI on a structurally unsymmetric operator) - nothing was changed and the
 * caller redoes the setup on the host and uploads again. */
int mgb200_replace_matrix(mgb200_handle h, int64_t n, const int64_t* colptr, const int64_t* rowval, const void* nzval,
                          int index_base, int relax_type, const double* relax_param, int* done);

/* Values of a resident matrix back to the host, in the caller's convention (nzval of the stored CSC array of level
 * `level`: which = 0 As[level], 1 Ps[level], 2 Rs[level]; nnz must match) - after mgb200_replace_matrix this is how
 * param.As[2:end] are refreshed on the host when somebody needs them - and relaxPrecs[level]. */
int mgb200_download_values(mgb200_handle h, int level, int which, void* nzval, int64_t nnz);
int mgb200_download_relax_prec(mgb200_handle h, int level, void* d);

/* Mixed precision (getMultigridPreconditioner with VAL != eltype(B), SolveFuncs.jl:52-60): a double-precision handle
 * that holds no hierarchy of its own, only the Krylov matrix (mgb200_set_krylov_matrix, mandatory) and the Krylov
 * vectors; its preconditioner is one cycle of the single-precision hierarchy `inner` (MGB200_FP32 -> an MGB200_FP64
 * outer handle, MGB200_CFP32 -> MGB200_CFP64):  bl .= r (rounded to single); z .= 0; recursiveCycle(param,bl,z,1);
 * z2 .= z (widened).  n and nrhs are taken from `inner`, which must outlive the outer handle.  The Krylov drivers
 * (mgb200_solveCG / solveFGMRES / solveBiCGSTAB and their block forms) run on the outer handle. */
int mgb200_create_mixed(mgb200_handle* outer, mgb200_handle inner);

/* Optional: the matrix the Krylov drivers multiply with when it is not As[1]
 * (solveCG_MG(AT,param,...) takes AT separately, SolveFuncs.jl:77-82).  CSC of A^H.
 * colptr == NULL (or n == 0) releases it: the drivers multiply with As[1] again.  The reference builds Afun from the
 * AT of EVERY call (getAfun, SolveFuncs.jl:65-82), so a front end calls this at the start of every Krylov solve. */
int mgb200_set_krylov_matrix(mgb200_handle h, int64_t n, const int64_t* colptr, const int64_t* rowval,
                             const void* nzval, int index_base);

/* adjustMemoryForNumRHS (MGsetup.jl:166-223) */
int mgb200_adjust_nrhs(mgb200_handle h, int nrhs);

/* setters that mirror mutable MGparam fields */
int mgb200_set_cycle(mgb200_handle h, char cycle_type, const int64_t* relax_pre, const int64_t* relax_post);

/* ---- solve phase, host buffers (column-major n x nrhs, VAL) --------------------------------- */

/* x = recursiveCycle(param,b,x,1) (MGcycle.jl:1-118); x is read and overwritten. */
int mgb200_cycle(mgb200_handle h, const void* b, void* x);

/* The closure getMultigridPreconditioner returns (SolveFuncs.jl:43-63):  z .= 0; recursiveCycle(param,r,z,1); z.
 * Host buffers like mgb200_cycle, but z is only written: the zero initial guess is not copied to the device. */
int mgb200_precondition(mgb200_handle h, const void* r, void* z);

/* solveMG (SolveFuncs.jl:3-39).  resvec must hold max_iter+1 doubles: resvec[0] = res_init,
 * resvec[k] = ||b - A x_k|| after cycle k.  *iter = cycles done. */
int mgb200_solveMG(mgb200_handle h, const void* b, void* x, double tol, int max_iter, int* iter,
                   double* resvec);

/* solveCG_MG (SolveFuncs.jl:103-116): KrylovMethods.cg (nrhs = 1) or blockCG (nrhs > 1) with one
 * cycle as preconditioner.  resvec: max_iter doubles (nrhs = 1) or max_iter*nrhs (block, row k =
 * column relative residuals of iteration k).  flag: 0 converged, -1 max_iter, -2 breakdown, -9 b = 0. */
int mgb200_solveCG(mgb200_handle h, const void* b, void* x, double tol, int max_iter, int* iter,
                   int* flag, double* resvec);

/* solveGMRES_MG (SolveFuncs.jl:120-132): KrylovMethods.fgmres(A,b,inner; flexible) with one cycle
 * as right preconditioner; max_iter counts restarts.  resvec: inner*max_iter doubles,
 * *nres = number of entries written. */
int mgb200_solveFGMRES(mgb200_handle h, const void* b, void* x, int inner, int flexible, double tol,
                       int max_iter, int* iter, int* flag, double* resvec, int* nres);

/* solveBiCGSTAB_MG (SolveFuncs.jl:85-99): KrylovMethods.bicgstb with one cycle as M1 (M2 = identity), nrhs = 1.
 * resvec: max_iter+1 doubles, resvec[0] = initial relative residual.  flag: 0 converged, -1 max_iter,
 * -2 breakdown, -3 converged after the first half of an iteration, -9 b = 0.  *iter = completed iterations,
 * *nprec = cycles applied (the reference's nprec = 2*iter + (flag == -3), SolveFuncs.jl:97). */
int mgb200_solveBiCGSTAB(mgb200_handle h, const void* b, void* x, double tol, int max_iter, int* iter,
                         int* flag, double* resvec, int* nprec);

/* SpMatMul(alpha,AT,x,beta,y) (SpMatMul.jl:4-26) on an uploaded matrix:
 * which = 0: A_l, 1: P_l, 2: R_l.  (alpha,beta) must be one of (1,0),(1,1),(-1,1),(-1,0). */
int mgb200_spmatmul(mgb200_handle h, int level, int which, double alpha, const void* x, double beta,
                    void* y);

/* ---- solve phase, device-resident buffers (RHS-fastest layout x[i*nrhs + j]) ----------------- */

/* Fine-level device buffers of the hierarchy (memCycle[1].b / .x of the reference). */
int mgb200_device_buffers(mgb200_handle h, void** d_b, void** d_x);

/* One cycle with b already in the device buffer.  x_is_zero != 0 is the preconditioner call
 * (z .= 0; recursiveCycle) of getMultigridPreconditioner (SolveFuncs.jl:59).  *d_x_out aliases
 * an internal buffer, as the reference's closure returns memCycle[1].x. */
int mgb200_cycle_device(mgb200_handle h, int x_is_zero, void** d_x_out);

/* Same as the host-buffer solvers with b in the device b buffer and x in the device x buffer. */
int mgb200_solveMG_device(mgb200_handle h, double tol, int max_iter, int* iter, double* resvec);
int mgb200_solveCG_device(mgb200_handle h, double tol, int max_iter, int* iter, int* flag, double* resvec);

int mgb200_synchronize(mgb200_handle h);

/* ---- multi-GPU: one process per GPU, rows partitioned in contiguous ranges ---------------------
 * (z-slabs of the reference's DomainDecomposition layout, src/DomainDecomposition/DDIndices.jl:41-47).
 * Call order: create -> dist_init -> dist_upload_level for the distributed levels (fine ones),
 * upload_level / upload_coarsest with the GLOBAL matrices for the replicated (coarse) levels.
 * Vectors passed to the solve entry points then hold only the owned rows of this rank.
 * On the device a distributed level keeps its vectors as [ghost rows below | owned rows | ghost rows above], so a
 * stencil matrix keeps its row-relative structure (and its stencil dictionary) on every slab.  Ghost rows travel
 * over NVLink peer memory written by the library's own kernels when the ranks can map each other's memory
 * (cudaIpc), otherwise through ncclSend/ncclRecv; Krylov scalars always use ncclAllReduce. */

/* 128-byte NCCL unique id, created on one rank and distributed by the caller (e.g. torch.distributed,
 * Julia Distributed); the same bytes go to every rank's mgb200_dist_init. */
int mgb200_dist_unique_id(char* out128);
int mgb200_dist_init(mgb200_handle h, int rank, int world, const char* unique_id128);

/* Owned rows of level l: like mgb200_upload_level, but every CSC block has one column per OWNED row
 * (A, P: rows row_offsets[rank] .. row_offsets[rank+1]-1 of level l; R: the owned rows of level l+1
 * given by coarse_row_offsets) and its row indices are GLOBAL column indices of the operator.
 * row_offsets / coarse_row_offsets have world+1 entries. */
int mgb200_dist_upload_level(mgb200_handle h, int level, int64_t n_global, const int64_t* row_offsets,
                             int64_t nc_global, const int64_t* coarse_row_offsets,
                             const int64_t* a_colptr, const int64_t* a_rowval, const void* a_nzval,
                             const int64_t* p_colptr, const int64_t* p_rowval, const void* p_nzval,
                             const int64_t* r_colptr, const int64_t* r_rowval, const void* r_nzval,
                             const void* d, int index_base);

/* ---- multi-GPU, single process: ONE handle drives G devices (SURVEY.md 8(b) "Threading": a single ccall from one Julia
 * task is enough; no mpirun / torchrun / Distributed workers on the product path) ------------------------------------------
 * The entry points take GLOBAL arrays exactly as the reference holds them and slice the contiguous row ranges of the
 * z-slab partition themselves; every device runs its part on its own host thread inside the call, ghost planes travel
 * over NVLink peer memory (plain peer access between devices of one process).
 *   mgb200_multi_create         getMGparam + the device part of MGsetup on devices[0 .. n_devices-1] (NULL: 0, 1, ...)
 *   mgb200_multi_upload_level   As[l], Ps[l], Rs[l], relaxPrecs[l] (global CSC arrays as for mgb200_upload_level).
 *                               row_offsets / coarse_row_offsets (n_devices + 1 entries, first 0, last n / nc): the rows of
 *                               level l / l+1 each device owns - for the reference's box partition with NumCells = [1,1,G]
 *                               (getOriginalBoundingBoxCells, DDIndices.jl:41-47) device g owns node planes g*c .. (g+1)*c-1,
 *                               the last one also the final plane.  row_offsets == NULL: the level is replicated on every
 *                               device (the coarse levels below the agglomeration threshold); all levels after the first
 *                               replicated one must be replicated too.
 *   mgb200_multi_upload_coarsest  As[end] on every device
 *   mgb200_multi_solveMG / solveCG / solveFGMRES / precondition  as the single-device calls, with the GLOBAL b and x
 *   mgb200_multi_info           as mgb200_dist_info (of device 0) */
typedef struct mgb200_multi* mgb200_multi_handle;
int mgb200_multi_create(mgb200_multi_handle* mh, int n_devices, const int* devices, int val_type, int levels, int nrhs,
                        char cycle_type, int relax_kind, const int64_t* relax_pre, const int64_t* relax_post);
int mgb200_multi_destroy(mgb200_multi_handle mh);
int mgb200_multi_upload_level(mgb200_multi_handle mh, int level, int64_t n, int64_t nc, const int64_t* row_offsets,
                              const int64_t* coarse_row_offsets, const int64_t* a_colptr, const int64_t* a_rowval,
                              const void* a_nzval, const int64_t* p_colptr, const int64_t* p_rowval, const void* p_nzval,
                              const int64_t* r_colptr, const int64_t* r_rowval, const void* r_nzval, const void* d,
                              int index_base);
int mgb200_multi_set_level_grid(mgb200_multi_handle mh, int level, int dim, const int64_t* n_fine_nodes,
                                const int64_t* n_coarse_nodes);      /* as mgb200_set_level_grid, GLOBAL grids */
int mgb200_multi_upload_coarsest(mgb200_multi_handle mh, int64_t n, const int64_t* colptr, const int64_t* rowval,
                                 const void* nzval, int index_base);
int mgb200_multi_solveMG(mgb200_multi_handle mh, const void* b, void* x, double tol, int max_iter, int* iter, double* resvec);
int mgb200_multi_solveCG(mgb200_multi_handle mh, const void* b, void* x, double tol, int max_iter, int* iter, int* flag,
                         double* resvec);
int mgb200_multi_solveFGMRES(mgb200_multi_handle mh, const void* b, void* x, int inner, int flexible, double tol, int max_iter,
                             int* iter, int* flag, double* resvec, int* nres);
int mgb200_multi_precondition(mgb200_multi_handle mh, const void* r, void* z);
int mgb200_multi_info(mgb200_multi_handle mh, int64_t* out);

/* out[0] = world, out[1] = rank, out[2] = 1 if halo exchange and coarse gather run over NVLink peer memory
 * (CUDA IPC, csrc/p2p.cuh; 0: NCCL send/recv, e.g. when MGB200_P2P=0 or the peers are not IPC reachable),
 * out[3] = number of row-partitioned levels.  Collective on first use (it finalises the distributed setup). */
int mgb200_dist_info(mgb200_handle h, int64_t* out);

/* Host-only planning helper (no GPU): sorted unique ghost ids of a row slab [lo,hi) and the column
 * indices remapped to the [owned | ghost] layout.  ghosts must hold nnz entries. */
int mgb200_host_plan_ghosts(int64_t nnz, const int64_t* cols, int64_t lo, int64_t hi, int64_t* ghosts,
                            int64_t* n_ghost, int64_t* local_cols);

/* Host-only helpers exported for the CPU test-suite (the tiny dense algebra of the Krylov control
 * flow, csrc/smalldense.h): complex numbers are interleaved (re, im), matrices row-major. */
int mgb200_host_pinv_apply(int n, const double* H, const double* xi, double* t);
int mgb200_host_general_pinv(int n, const double* A, double rtol, double* P);
int mgb200_host_hessenberg_lsq(int cols, const double* H, const double* xi, double* y, double* res);
/* block GMRES / block BiCGStab pieces: dense least squares with nb right-hand sides (*res = residual norm);
 * what = 0 upper Cholesky factor of A (out m x m), what = 1 out = A^{-1} B (m x nb); status -5 when A is not
 * positive definite / singular */
int mgb200_host_dense_lsq(int rows, int cols, int nb, const double* A, const double* B, double* Y, double* res);
int mgb200_host_small_factor(int what, int m, int nb, const double* A, const double* B, double* out);
/* rows [out[0], out[1]) of a CSR slab that read no ghost row of their input vector (columns relative to the first
 * owned row: < 0 lower ghost, >= n_in_owned upper ghost): the rows launched beside the halo exchange */
int mgb200_host_interior_rows(int64_t n_rows, const int64_t* rowptr, const int64_t* cols, int64_t n_in_owned,
                              int64_t* out);

/* ---- introspection / measurement -------------------------------------------------------------- */

/* Kernel selection made at upload for matrix `which` of `level` (see mgb200_spmatmul):
 * out[0] = threads per row, out[1] = rows per CTA, out[2] = shared bytes per CTA,
 * out[3] = 1 if the TMA-staged kernel is used, 0 for the row-per-warp fallback,
 * out[4] = nnz, out[5] = max row length. */
int mgb200_kernel_config(mgb200_handle h, int level, int which, int64_t* out);

/* Stencil-dictionary form of matrix `which` of `level` (csrc/pattern.cuh): when the rows of an uploaded CSR
 * matrix are copies of a few (column-offset, value) stencils the device keeps one 16-bit pattern id per row
 * and a small dictionary instead of streaming 12-20 bytes per non-zero; results are bit-identical.
 * out[0] = 1 if in use, out[1] = 1 if offsets are row-relative (no per-row base column), out[2] = patterns,
 * out[3] = dictionary entries, out[4] = 1 if relaxPrecs[level] is folded into the dictionary (which = 0). */
int mgb200_pattern_info(mgb200_handle h, int level, int which, int64_t* out);

/* Runtime options (every MGB200_<KEY> environment variable sets the same default at mgb200_create; results never change
 * by a bit, only the kernels that produce them - tests/test_patterns.py):
 *  "patterns" (1/0: stencil dictionary where the rows deduplicate; set before upload to skip building it),
 *  "graphs" (1/0: replay V/F/W cycles from CUDA graphs), "smem_budget" (bytes per CTA when choosing the rows per CTA of the
 *  CSR-stream kernel at upload), "tma" (1/0: TMA-staged variant of the dictionary kernel), "tma_min_rows",
 *  "box" (1/0: box-stencil kernel on box-structured square operators, csrc/box.cuh), "box_min_rows",
 *  "box_variant" / "box_variant27" / "box_variant_c" (-1 = the library's choice; else the tile
 *  variant for all / 27-point / ComplexF64 levels), "fuse_first" (1/0: first two sweeps from x = 0 in one pass),
 *  "grid_transfers" (1: grid-hinted transfer kernels on levels whose hint was verified, mgb200_set_level_grid; 2: also
 *  the block prolongation; 0: off), "gxp_quad" (1: quad-form prolongation with 64 registers, 2 / 3: 40 / 32; 0: line form),
 *  "mrhs_march" (1/0: marching block kernel for nrhs >= 8),
 *  multi-GPU: "fused_put" (1/0: the producing kernel stores the slab-end rows to the neighbours itself), "overlap_box"
 *  (1/0: the box kernel runs beside the halo exchange of its input and waits for the ghost rows inside the kernel),
 *  "overlap" (1/0: exchange beside the rows that read no ghost; measured slower, off),
 *  "split_test" (rows: single-GPU test hook that forces the split launch sequence of the overlap path). */
int mgb200_set_option(mgb200_handle h, const char* key, int64_t value);

/* Host-only (no GPU): box structure of a row-relative dictionary (csrc/pattern.cuh::detect_box) - every column offset
 * is dz*S2 + dy*S + dx with dx, dy, dz in {-1,0,1}.  Input as mgb200_host_build_patterns.  info[0] = 1 if the matrix
 * has it, info[1] = S (line length, 0: 1-D), info[2] = S2 (plane length, 0: 2-D), info[3] = patterns;
 * mask[p] = presence bits (dz+1)*9 + (dy+1)*3 + (dx+1) of pattern p (caller-allocated, max_patterns). */
int mgb200_host_detect_box(int64_t n_rows, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                           int index_base, int max_patterns, int max_entries, int64_t* info, int32_t* mask);

/* CPU replay of one launch of the box-stencil kernel (csrc/box.cuh) on a matrix given in the upload format: the
 * kernel's own tile plan, copy list and per-thread function run on the host, the stages being host buffers filled where
 * the bulk copies fill shared memory.  Test hook (tests/test_patterns.py); no GPU is used.  x, b, d must carry the 4
 * elements of slack device vectors have.  mode 0: A x, 2: b - A x, 3: x + d.*(b - A x), 4: the first two sweeps from
 * zero in one pass, x1 = 0 + d.*x, y = x1 + d.*(x - A x1) with the right-hand side in `x` and d constant per pattern.  info[0] = 1 if the matrix qualifies (else y is untouched), info[1] = stencil
 * shape (7 or 27), info[2] = patterns, info[3] = rows computed on the constant-coefficient fast path. */
int mgb200_host_box_apply(int64_t n_rows, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                          int index_base, int mode, int rows_per_thread, int base_rows, int ctas, int fold_d,
                          const double* x, const double* b, const double* d, double* y, int64_t* info);

/* Host-only (no GPU): the one-row-per-thread dictionary walk of pat_kernel (csrc/pattern.cuh) on the CPU for a square
 * row-relative matrix given in the upload format - the reference every other kernel form must match bit for bit
 * (tests/test_patterns.py).  mode 0: y = A x, 2: y = b - A x, 3: y = x + d.*(b - A x); fold_d: d taken per pattern.
 * info[0] = 1 if the matrix has a row-relative dictionary (else y is untouched), info[1], info[2] = line / plane length
 * when the offsets have box structure. */
int mgb200_host_pattern_apply(int64_t n_rows, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                              int index_base, int mode, int fold_d, const double* x, const double* b, const double* d,
                              double* y, int64_t* info);

/* Host-only (no GPU): the grid-hinted transfer kernels' per-thread functions (csrc/grid_xfer.cuh, __host__ __device__) run
 * on the CPU for every thread of a launch, for a real Float64 transfer matrix given by its CSC-of-the-transpose arrays
 * as uploaded (kind 1: Ps[l], y += P x with x coarse, y fine; kind 2: Rs[l], y = R x with x fine, y coarse).
 * lines_per_thread != 0: the per-row functions of the kernels; 0: the dictionary walk (the reference they must match bit for bit).
 * info[0] = 1 if the hint matches the matrix (y is then written / updated), else 0 and y is untouched. */
int mgb200_host_grid_transfer(int kind, int dim, const int64_t* n_fine_nodes, const int64_t* n_coarse_nodes, int64_t n_rows,
                              const int64_t* colptr, const int64_t* rowval, const double* nzval, int index_base,
                              int lines_per_thread, const double* x, double* y, int64_t* info);

/* Host-only (no GPU): the window plan of the TMA-staged dictionary kernel for a row-relative matrix, exported for the
 * CPU test-suite.  Input as mgb200_host_build_patterns; `tile` rows per tile, elem_bytes 4, 8 or 16 (copies are rounded
 * outwards to 16 bytes).  info[0] = 1 if a plan exists (row-relative dictionary, few enough windows), info[1] = windows,
 * info[2] = elements per stage, info[3] = stage offset of the centre x[row], info[4] = dictionary entries.
 * win[3*g .. 3*g+2] = (lowest offset, elements copied, first element in the stage buffer) of window g (<= 16 windows);
 * delta[k] / soff[k] = column offset and stage offset of dictionary entry k (caller-allocated, max_entries). */
int mgb200_host_tma_plan(int64_t n_rows, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                         int index_base, int tile, int elem_bytes, int max_patterns, int max_entries, int64_t* info,
                         int32_t* win, int32_t* delta, int32_t* soff);

/* Host-only (no GPU): the row deduplication behind the stencil dictionary, exported for the CPU test-suite.
 * Input: CSR arrays (row pointers colptr[n_rows+1], Int64 columns, Float64 values).  info[0] = 1 if the matrix
 * deduplicates within max_patterns / max_entries, info[1] = row-relative, info[2] = patterns, info[3] = entries.
 * Outputs (caller-allocated): pid[n_rows], c0[n_rows], pat_off[max_patterns+1], delta/val[max_entries]. */
int mgb200_host_build_patterns(int64_t n_rows, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                               int index_base, int max_patterns, int max_entries, int64_t* info, uint16_t* pid,
                               int32_t* c0, int32_t* pat_off, int32_t* delta, double* val);

/* CUDA-event timing of every kernel launch (off by default).  The report is a sequence of
 * records {kind, level, launches, total_ms, algorithmic_bytes, format_bytes}: 6 doubles each
 * (format_bytes = what the device format really streams, e.g. the stencil-dictionary form); returns
 * the number of records written (<= max_records) in *nrec and resets the counters. */
int mgb200_profile_enable(mgb200_handle h, int on);
int mgb200_profile_report(mgb200_handle h, double* records, int max_records, int* nrec);
int64_t mgb200_launch_count(mgb200_handle h);

/* CUDA events on the library's own stream (torch.cuda.Event only sees torch's stream):
 * record event slot idx (0..15); elapsed waits for slot i1 and returns the time from i0 to i1. */
int mgb200_event_record(mgb200_handle h, int idx);
int mgb200_event_elapsed_ms(mgb200_handle h, int i0, int i1, double* ms);

/* Stand-alone SpMV y = A_level x on device-resident buffers (the b and r workspaces of that level), `reps` calls
 * timed with CUDA events on the library's stream: *ms_per_call and the algorithmic bytes of one call
 * (nnz*(sv+4) + 4(n+1) + 2 n sv m).  The r workspace is overwritten. */
int mgb200_bench_spmv(mgb200_handle h, int level, int reps, double* ms_per_call, double* algorithmic_bytes);

/* cudaProfilerStart / cudaProfilerStop, so that `ncu --profile-from-start off` captures exactly
 * the timed region of bench.py. */
int mgb200_profiler_start(void);
int mgb200_profiler_stop(void);

/* kinds used in profile records */
#define MGB200_K_SWEEP 0    /* x' = x + d.*(b - A x)         */
#define MGB200_K_RESID 1    /* r = b - A x                    */
#define MGB200_K_SPMV 2     /* y = A x                        */
#define MGB200_K_RESTRICT 3 /* bc = R r                       */
#define MGB200_K_PROLONG 4  /* x += P xc                      */
#define MGB200_K_DIAG 5     /* x = d.*b (first sweep, x = 0)  */
#define MGB200_K_COARSE 6   /* coarsest dense solve           */
#define MGB200_K_REDUCE 7   /* dots / norms                   */
#define MGB200_K_VECTOR 8   /* axpy-class vector updates      */
#define MGB200_K_COPY 9     /* copies / layout changes        */

#ifdef __cplusplus
}
#endif
#endif /* MGB200_H */
