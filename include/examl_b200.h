/* examl_b200.h -- C ABI of the B200 likelihood engine (libexaml_b200.so).
 *
 * This is the boundary at which ExaML's host code hands the per-site conditional-likelihood
 * work to the GPU.  It sits exactly where the reference crosses from its traversal executors
 * into its SIMD kernels: every entry point below replaces one reference
 * function, takes the same information that function reads from `tree` / `pInfo` /
 * `traversalInfo`, flattened to plain pointers and sizes, and returns what that function
 * leaves behind for its caller.  No torch / CUDA types appear in any signature.
 *
 *   reference function (file:line)                           replaced by
 *   -------------------------------------------------------   ---------------------------
 *   initializePartitions xVector/yVector/wgt/globalScaler      exb_create, exb_set_partition,
 *     allocation (axml.c:1936-2104) + lazy x3 allocation        exb_commit
 *     (newviewGenericSpecial.c:1199-1216)
 *   reading pInfo.{EIGN,EV,EI,tipVector,gammaRates} on every   exb_set_model
 *     call (newviewGenericSpecial.c:939-1045,
 *     evaluateGenericSpecial.c:433-480)
 *   tr->rateHetModel, pInfo.{numberOfCategories,              exb_set_rate_model,
 *     perSiteRates,rateCategory} (-m PSR)                       exb_set_categories
 *   newviewIterative      (newviewGenericSpecial.c:917)        exb_newview
 *   evaluateIterative     (evaluateGenericSpecial.c:403)       exb_evaluate
 *     + MPI_Allreduce     (evaluateGenericSpecial.c:968-975)     (all-reduce inside, NCCL)
 *   makenewzIterative     (makenewzGenericSpecial.c:628)       exb_makenewz_prepare
 *   execCore              (makenewzGenericSpecial.c:849)       exb_core
 *     + MPI_Allreduce     (makenewzGenericSpecial.c:1241-1248)   (all-reduce inside, NCCL)
 *   topLevelMakenewz      (makenewzGenericSpecial.c:1133)      exb_makenewz (single rank: whole NR loop on the device)
 *   evaluatePartialGeneric (evaluatePartialGenericSpecial.c:259) exb_evaluate_partial (batched over sites and rates)
 *   testInsertBIG, non-thorough (searchAlgo.c:682: insertBIG +    exb_score_insertions (all candidates of a pruned subtree
 *     newviewGeneric + evaluateGeneric per candidate branch)        in one call; behind batchedInsertionScanBegin / End)
 *
 * The three public axml.h entry points (evaluateGeneric axml.h:1223, newviewGeneric :1232,
 * makenewzGeneric :1241) are exported by the companion translation unit
 * examl_b200/csrc/axml_backend.c, which is compiled against the reference's own axml.h and
 * calls only the functions declared here.
 *
 * Conventions
 *   - Node numbers are ExaML's: tips 1..mxtips, inner nodes mxtips+1..2*mxtips-2.
 *   - A traversal is the reference's `traversalInfo` array without its 2x256 branch slots:
 *     `exb_entry ti[count]` plus `qz[count*numBranches]`, `rz[count*numBranches]` holding the
 *     raw branch values z = exp(-t) exactly as stored in traversalInfo.qz / .rz
 *     (axml.h:434-442).  The engine applies the reference's clamp `z > zmin ? log z : log zmin`
 *     (newviewGenericSpecial.c:983-984) itself.
 *   - For exb_evaluate / exb_makenewz_prepare, ti[0] is the branch entry (pNumber, qNumber,
 *     qz = branch value), exactly like td[0].ti[0] (evaluateGenericSpecial.c:410-416);
 *     newviews are ti[1..count).  For exb_newview all entries are newviews.
 *   - `executeModel[numPartitions]` is td[0].executeModel (1 = compute); a partition is
 *     processed iff executeModel[m] && width > 0 (newviewGenericSpecial.c:1114).
 *   - Branch slot of partition m is m when numBranches > 1 (-M), else 0
 *     (newviewGenericSpecial.c:972-981).
 *   - All likelihood arithmetic is IEEE double.  CLVs, tip codes, weights, scaler counts and
 *     the sum table stay resident in HBM between calls; only descriptors and O(partitions)
 *     scalars cross the bus.
 *   - Every function returns EXB_OK (0) or a negative error code; exb_last_error() gives the
 *     text.  There is no CPU fallback: without a usable CUDA device exb_create fails.
 *     (The reference's own convention for unrecoverable errors is abort -- axml.c:129-140 --
 *     and the axml backend turns any non-zero return into exactly that.)
 */
#ifndef EXAML_B200_H
#define EXAML_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EXB_OK             0
#define EXB_ERR_ARG       -1   /* bad argument / unsupported configuration */
#define EXB_ERR_CUDA      -2   /* CUDA runtime error                        */
#define EXB_ERR_STATE     -3   /* call sequence error                       */
#define EXB_ERR_NCCL      -4   /* NCCL error / NCCL not loadable            */

#define EXB_TIP_TIP        0   /* axml.h:302-304 */
#define EXB_TIP_INNER      1
#define EXB_INNER_INNER    2

typedef struct exb_engine exb_engine;

/* traversalInfo (axml.h:434-442) minus the branch arrays. */
typedef struct {
  int tipCase;
  int pNumber;
  int qNumber;
  int rNumber;
} exb_entry;

/* ---- life cycle -------------------------------------------------------------------------- */

/* Create an engine on CUDA device `device` for a tree with `mxtips` taxa, `numPartitions`
 * partitions (tr->NumberOfModels) and `numBranches` branch slots (tr->numBranches). */
int  exb_create(exb_engine **out, int device, int mxtips, int numPartitions, int numBranches);
void exb_destroy(exb_engine *e);
const char *exb_last_error(void);

/* Describe partition `model` as this rank sees it (pInfo.states/.width/.wgt/.yResource,
 * axml.h:533-629): `states` is 4 (DNA) or 20 (AA); `width` may be 0; `wgt[width]`;
 * `yResource[mxtips*width]` taxon-major tip codes (byteFile.c:320-346). */
int  exb_set_partition(exb_engine *e, int model, int states, size_t width,
                       const int *wgt, const unsigned char *yResource);

/* Rate heterogeneity model of the run (tr->rateHetModel, axml.h:349-350): EXB_GAMMA (default; four discrete GAMMA
 * categories per site, CLV = 4*states doubles) or EXB_CAT (-m PSR: ONE rate category per site, CLV = states
 * doubles; maxCategories = tr->maxCategories <= 64).  Call before exb_set_partition. */
#define EXB_CAT    0
#define EXB_GAMMA  1
int  exb_set_rate_model(exb_engine *e, int rateHetModel, int maxCategories);

/* Freeze the layout, allocate the device arenas (CLVs for mxtips-2 inner nodes, sum table,
 * scalers) and upload tips and weights. */
int  exb_commit(exb_engine *e);

/* ---- per-call inputs the host may have changed silently ------------------------------------- */

/* Model constants of partition `model`: EIGN[S], EV[S*S], EI[S*S], tipVector[codes*S]
 * (codes = 16 for DNA, 23 for AA), gammaRates[4] (models.c:3234-3457, 3795).  Cheap when
 * unchanged (host-side compare). */
int  exb_set_model(exb_engine *e, int model, const double *EIGN, const double *EV,
                   const double *EI, const double *tipVector, const double *gammaRates);

/* LG4M / LG4X protein models (pInfo.protModels, axml.h:262-263): one eigen-system per GAMMA category and, for LG4X,
 * free category weights -- pInfo.{EIGN,EV,EI,tipVector}_LG4[4] flattened to [4][20], [4][400], [4][400], [4][460],
 * gammaRates[4], weights[4] (newviewGTRGAMMAPROT_LG4 avxLikelihood.c:814, evaluateGenericSpecial.c:1164-1287,
 * makenewzGenericSpecial.c:2489-2578).  GAMMA only. */
int  exb_set_model_lg4(exb_engine *e, int model, const double *EIGN, const double *EV, const double *EI,
                       const double *tipVector, const double *gammaRates, const double *weights);

/* -S, the memory saving option for gappy alignments (tr->saveMemory; SEV, newviewGenericSpecial.c:1160-1216,
 * newviewGTRGAMMA_AVX_GAPPED_SAVE avxLikelihood.c:1806, ...PROT... :3125).  The reference keeps, per inner node, only the
 * sites that are not all-gap in the subtree below it plus one shared "gap column".  The engine applies the idea per
 * (node, tile of <= tile_sites sites): a pair whose sites are all gaps below the node owns no CLV block, is neither
 * computed nor written, and reads as the node's gap column; blocks come from slabs allocated on demand, so the HBM
 * footprint follows the data instead of (taxa - 2) x sites.  Results are those of the dense run.  GAMMA only (under
 * EXB_CAT the arena stays dense).  Before exb_commit.  exb_sev_stats: blocks in use / blocks a dense arena would hold /
 * bytes of slabs allocated (any pointer may be NULL). */
int  exb_set_sev(exb_engine *e, int on);
int  exb_sev_stats(exb_engine *e, size_t *blocks_in_use, size_t *blocks_dense, size_t *bytes_allocated);

/* Arithmetic of the protein (20-state) GAMMA newview.  EXB_AA_FP64 (default): FP64 throughout, on the FP64 tensor pipe
 * (DMMA) -- the reference's precision (newviewGTRGAMMAPROT_AVX avxLikelihood.c:1312).  EXB_AA_TF32X3: the contractions
 * run on the tcgen05 tensor cores as 3xTF32 products with FP32 accumulators in tensor memory, site vectors renormalised
 * by an exact power of two per node; CLVs are still stored as FP64.  Measured (tools/mixed_precision_study.py,
 * tests/test_gpu_parity.py): lnL within 1e-7 relative of the FP64 path -- inside north_star's 1e-6 -- but per-site
 * log-likelihoods differ by up to ~2e-3, so it is opt-in (also: EXB_AA_PRECISION=tf32x3 in the environment).  Applies
 * to GAMMA partitions with a single eigen-system (not LG4M / LG4X, not PSR).  Before exb_commit. */
#define EXB_AA_FP64    0
#define EXB_AA_TF32X3  1
int  exb_set_aa_precision(exb_engine *e, int mode);

/* EXB_CAT only: pInfo.numberOfCategories, pInfo.perSiteRates[numberOfCategories], pInfo.rateCategory[width]
 * (rewritten by optimizeRateCategories, optimizeModel.c:2403).  Cheap when unchanged (host-side compare). */
int  exb_set_categories(exb_engine *e, int model, int numberOfCategories, const double *perSiteRates,
                        const int *rateCategory);

/* ---- the hot path ------------------------------------------------------------------------ */

/* newviewIterative(tr, startIndex): recompute the CLVs listed in ti[0..count). */
int  exb_newview(exb_engine *e, const exb_entry *ti, int count,
                 const double *qz, const double *rz, const unsigned char *executeModel);

/* evaluateIterative + the all-reduce of evaluateGeneric: newviews ti[1..count), then the
 * log-likelihood at branch ti[0].  perPartitionLH[numPartitions]: executed partitions
 * receive their (globally reduced) value, partitions with global width 0 receive 0.0,
 * masked partitions are left untouched (evaluateGenericSpecial.c:845-860). */
int  exb_evaluate(exb_engine *e, const exb_entry *ti, int count,
                  const double *qz, const double *rz, const unsigned char *executeModel,
                  double *perPartitionLH);

/* makenewzIterative: newviews ti[1..count), then the per-site sum table of branch ti[0]. */
int  exb_makenewz_prepare(exb_engine *e, const exb_entry *ti, int count,
                          const double *qz, const double *rz, const unsigned char *executeModel);

/* execCore + the all-reduce of topLevelMakenewz: first and second derivative of lnL with
 * respect to lz = log z at the trial values lz[numBranches] (tr->coreLZ).  Outputs
 * dlnLdlz[numBranches], d2lnLdlz2[numBranches]; with numBranches == 1 all executed partitions
 * accumulate into slot 0 (makenewzGenericSpecial.c:921-946). */
int  exb_core(exb_engine *e, const double *lz, const unsigned char *executeModel,
              double *dlnLdlz, double *d2lnLdlz2);

/* topLevelMakenewz (makenewzGenericSpecial.c:1133-1349) entirely on the device: newviews ti[1..count), the sum
 * table of branch ti[0], then the Newton-Raphson iteration on z for every branch slot -- derivative pass, reduction
 * and the scalar update stay on the GPU (one cooperative kernel with grid barriers); a single copy brings back
 * result[numBranches] (the new z, in [zmin, zmax]), curvatOK[numBranches] (tr->curvatOK) and the number of
 * iterations (iterationsDone, may be NULL).  z0[numBranches] are the start values (= ti[0] branch values), maxiter as in makenewzGeneric,
 * executeModel[] the initial mask.  With several ranks on one node (exb_comm_is_p2p) the all-reduce of every iteration
 * (makenewzGenericSpecial.c:1241-1248) happens inside the same kernel over NVLink peer buffers and every rank returns the
 * same bits; with a communicator but no peer access (several nodes) it returns EXB_ERR_STATE and the caller runs the
 * loop itself with exb_makenewz_prepare + exb_core. */
int  exb_makenewz(exb_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz,
                  const unsigned char *executeModel, const double *z0, int maxiter,
                  double *result, int *curvatOK, int *iterationsDone);

/* evaluatePartialGeneric(tr, i, ki, model)  evaluatePartialGenericSpecial.c:259 -- batched.
 * Under CAT / PSR, optimizeRateCategories (optimizeModel.c:1792-1922) asks, site by site and candidate rate by
 * candidate rate, for the log likelihood of ONE site with ONE explicit rate, recomputed from the tips along the
 * full traversal td[0].  This entry point answers n such questions at once: for pair j the value
 *     wgt[site] * ( log|x_tip . x_inner . exp(EIGN * ki * log z_root)| + scale * log(minlikelihood) )
 * of site sites[j] (index inside partition models[j]) at rate rates[j] is written to out[j].
 * entries[0] is the root branch, its pNumber a tip (evaluatePartialGenericSpecial.c:210); entries[1..count)
 * must be a full traversal (every inner node read is computed earlier in the list).  The engine's CLVs,
 * scalers and categories are neither read nor changed.  Local: no collective is involved. */
int  exb_evaluate_partial(exb_engine *e, const exb_entry *entries, int count, const double *qz, const double *rz,
                          size_t n, const int *models, const int *sites, const double *rates, double *out);

/* Batched scoring of lazy SPR insertions.  The reference scores ONE candidate at a time:
 * testInsertBIG (searchAlgo.c:682) -> insertBIG, non-thorough branch (searchAlgo.c:519-535: both new branches get
 * sqrt(z_qr)) -> newviewGeneric(p) -> evaluateGeneric(p->next->next), i.e. two calls and ~3 node updates per candidate
 * branch, and the CLVs on the path from the pruning point are re-oriented node by node as addTraverseBIG
 * (searchAlgo.c:785) walks outwards.  This entry point executes ONE mixed descriptor for a whole pruned subtree:
 *   ti[0]          the subtree: pNumber = node s = p->back whose CLV (or tip) looks away from the insertion point,
 *                  qz[0..numBranches) = z of the branch p--s (evaluate's diagonal table, calcDiagptable);
 *   ti[1..count)   in order, either an ordinary newview entry (tipCase 0..2) -- used for the CLVs that have to look
 *                  TOWARDS a candidate branch; they may write to, and read from, scratch nodes numbered
 *                  2*mxtips-1 .. 2*mxtips-2+EXB_SCRATCH_NODES that exist only for this purpose, so the tree's own CLVs
 *                  and orientation stay untouched -- or a scoring entry (tipCase | EXB_SCORE): the CLV of a new node
 *                  joining qNumber and rNumber with branch values qz / rz is formed in registers (scaled like any
 *                  newview, avxLikelihood.c:253-311), never stored, and contracted with the subtree's CLV:
 *                      lnl[pNumber][m] = sum_i wgt[i] log(0.25 |sum_c sum_k x_new[i][c][k] x_s[i][c][k] diag[c][k]|)
 *                                        + (scalings) * log(minlikelihood)
 *                  -- by the pulley principle the value evaluateGeneric returns at (p->next->next, r) after that insertion.
 *                  pNumber of a scoring entry = index k of the candidate (0 <= k < ncand).
 * lnl[ncand * numPartitions] (candidate-major) is globally reduced like exb_evaluate's output; masked partitions
 * read 0.  Scratch CLVs and their scaler counts persist between calls, so a long scan may be cut into several calls
 * anywhere.  GAMMA, dense arena (not under EXB_CAT, -S, EXB_AA_TF32X3 or with LG4M / LG4X partitions: EXB_ERR_STATE, the
 * caller then scores the candidates one by one as before). */
#define EXB_SCORE          4
#define EXB_SCRATCH_NODES 64
int  exb_score_insertions(exb_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz,
                          const unsigned char *executeModel, int ncand, double *lnl);
/* 1 when exb_score_insertions is available in the engine's configuration */
int  exb_can_score_insertions(exb_engine *e);
/* Entries one descriptor may hold (any traversal call): at least mxtips + 1 like td[0].ti (axml.c:595), normally
 * 3 * mxtips + 1 so that a whole insertion scan fits one call. */
int  exb_max_entries(exb_engine *e);

/* ---- multi-GPU (one process per GPU) ------------------------------------------------------ */

/* NCCL bootstrap: rank 0 obtains an opaque id (EXB_UNIQUE_ID_BYTES bytes), the launcher
 * distributes it, every rank calls exb_comm_init.  Afterwards exb_evaluate and exb_core
 * sum their outputs over all ranks on the device (ncclAllReduce, double, sum) before the
 * single device-to-host copy. */
#define EXB_UNIQUE_ID_BYTES 128
int  exb_comm_unique_id(void *id_out);
int  exb_comm_init(exb_engine *e, int rank, int nranks, const void *id);
/* 1 when the ranks found each other on one node and mapped each other's receive buffers (CUDA IPC over NVLink):
 * the all-reduce then happens inside the finish kernel -- partial sums stored straight into every peer's memory,
 * summed in rank order on every GPU -- and no NCCL call remains on the per-call path (the default on one node).
 * 0: NCCL all-reduce after the kernel (several nodes, no peer access, or EXB_P2P=0 in the environment). */
int  exb_comm_is_p2p(exb_engine *e);

/* ---- introspection (tests, benchmarks) ----------------------------------------------------- */

/* Copy the CLV of inner node `nodeNumber`, partition `model`, to out[width*4*states]
 * (layout x[i][c][k], newviewGenericSpecial.c:430-682). */
int  exb_read_clv(exb_engine *e, int model, int nodeNumber, double *out);
/* globalScaler[nodeNumber] of partition `model` (newviewGenericSpecial.c:1503-1506). */
int  exb_read_scaler(exb_engine *e, int model, int nodeNumber, unsigned int *out);
/* Copy the sum table of partition `model` to out[width*4*states]. */
int  exb_read_sumtable(exb_engine *e, int model, double *out);
/* Device time in milliseconds of the kernels of the most recent hot-path call (CUDA events
 * on the engine's stream), and the number of kernel launches the engine has made so far. */
int  exb_last_kernel_ms(exb_engine *e, float *ms);
/* The same time split by phase for the most recent traversal call (exb_newview / exb_evaluate / exb_makenewz_prepare):
 * ms3[0] = d-table and tip look-up kernels (makeP, newviewGenericSpecial.c:78), ms3[1] = the traverse kernels,
 * ms3[2] = the finish kernel (scaler chain, per-partition reduction, in-kernel all-reduce).  Needs timing on. */
int  exb_last_phase_ms(exb_engine *e, float *ms3);
/* Switch the per-call CUDA events off (default on); saves two driver calls per hot-path call. */
int  exb_set_timing(exb_engine *e, int on);
long exb_launch_count(exb_engine *e);
/* Tuning knob: sites per work tile (multiple of 32).  Must be set before exb_commit. */
int  exb_set_tile_sites(exb_engine *e, int sites);

#ifdef __cplusplus
}
#endif
#endif /* EXAML_B200_H */
