/* axml_backend.c -- the B200 back-end of ExaML's likelihood API.
 *
 * This translation unit is compiled against the reference's own, unmodified axml.h and REPLACES
 * the five reference objects newviewGenericSpecial.o, evaluateGenericSpecial.o,
 * makenewzGenericSpecial.o, evaluatePartialGenericSpecial.o and avxLikelihood.o (object list:
 * examl/Makefile.AVX.gcc:18).  Every
 * other reference file (axml.c, searchAlgo.c, optimizeModel.c, models.c, byteFile.c, ...) is
 * compiled as it is and links against the four entry points it calls (nm on the reference
 * objects: only these are imported from the replaced files):
 *
 *     void evaluateGeneric(tree *tr, nodeptr p, boolean fullTraversal);                    axml.h:1223
 *     void newviewGeneric (tree *tr, nodeptr p, boolean masked);                           axml.h:1232
 *     void makenewzGeneric(tree *tr, nodeptr p, nodeptr q, double *z0, int maxiter,
 *                          double *result, boolean mask);                                  axml.h:1241
 *     double evaluatePartialGeneric(tree *tr, int i, double ki, int _model);               axml.h:1222
 *
 * plus the internal-but-global helpers of the replaced files, kept for API completeness:
 * computeTraversalInfo, newviewIterative, evaluateIterative, makenewzIterative, execCore,
 * isGap, noGap (axml.h:1224-1252, 1332-1333).
 *
 * What stays on the host (same behaviour as the reference, re-implemented here):
 *   - the traversal descriptor: post-order list of CLVs that are not oriented towards the
 *     branch of interest (reference: newviewGenericSpecial.c:691-811)
 *   - the Newton-Raphson state machine of makenewz with its per-partition convergence masks
 *     (reference: makenewzGenericSpecial.c:1133-1349)
 *   - executeModel / partitionConverged bookkeeping
 * What moves to the GPU (include/examl_b200.h): everything that touches per-site data.  CLVs,
 * tip codes, weights, scaler counts and the sum table never exist on the host; the
 * pInfo.xVector / sumBuffer / globalScaler arrays the reference allocates stay untouched.
 *
 * The two MPI_Allreduce calls of the replaced files (evaluateGenericSpecial.c:968-975,
 * makenewzGenericSpecial.c:1241-1248) become one ncclAllReduce inside exb_evaluate / exb_core.
 *
 * Unsupported modes are refused loudly, the way the reference's MIC back-end does
 * (axml.c:2645-2657): binary data.  -S is accepted but not applied (dense CLVs in HBM, identical results).  Both rate models are
 * implemented: GAMMA (including the LG4M / LG4X protein models) and PSR/CAT (-m PSR, including the per-site
 * rate scan of optimizeRateCategories: evaluatePartialGeneric below).  There is no CPU fallback.
 */
#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "axml.h"
#include "examl_b200.h"

extern int processID;
extern const unsigned int mask32[32];

static exb_engine *g_engine = NULL;
static int g_ranks = 1;
/* Deferred newviewGeneric (below): node numbers whose CLV on the device is out of date.  A node on this list is treated
   by the traversal builder exactly like a node that does not look the right way: it is recomputed, children first, by the
   next call that needs it. */
static unsigned char *g_stale = NULL;            /* [2 * mxtips] */
static int g_defer_newviews = 0;                 /* set in backend_init (EXAML_B200_EAGER_NEWVIEWS=1 switches it off) */
static int g_emit_marks_stale = 0;               /* computeTraversalInfo is building a list that will NOT be executed now */
#define IS_STALE(rec) (g_stale && g_stale[(rec)->number])
static exb_entry *g_entries = NULL;
static double *g_qz = NULL, *g_rz = NULL;
static unsigned char *g_mask = NULL;

static void backend_fail(const char *what)
{
  fprintf(stderr, "\nExaML B200 back-end, rank %d: %s failed: %s\n", processID, what, exb_last_error());
  MPI_Abort(MPI_COMM_WORLD, -1);
  exit(-1);
}

static void refuse(const char *why)
{
  if (processID == 0)
    fprintf(stderr, "\nThe B200 back-end of ExaML does not implement %s.\n"
                    "Supported: DNA and protein data under the GAMMA and PSR models of rate heterogeneity (with or without -M, -S).\n\n", why);
  MPI_Barrier(MPI_COMM_WORLD);
  MPI_Finalize();
  exit(-1);
}

/* Which GPU does this rank drive?  One rank per GPU; launchers export LOCAL_RANK (torchrun), OMPI_COMM_WORLD_LOCAL_RANK,
   SLURM_LOCALID, MPI_LOCALRANKID / PMI_LOCAL_RANK (Hydra), MV2_COMM_WORLD_LOCAL_RANK; otherwise the rank number, modulo
   EXAML_B200_GPUS_PER_NODE when that is given. */
static int local_device(void)
{
  const char *names[] = {"EXAML_B200_DEVICE", "LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID",
                         "MPI_LOCALRANKID", "PMI_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", NULL};
  const char *per_node = getenv("EXAML_B200_GPUS_PER_NODE");
  int i;
  for (i = 0; names[i]; i++) {
    const char *v = getenv(names[i]);
    if (v && *v) return atoi(v);
  }
  /* no launcher variable: ranks are usually placed node by node, so the rank modulo the GPUs of a node */
  if (per_node && atoi(per_node) > 0) return processID % atoi(per_node);
  return processID;
}

/* Bootstrap of the NCCL communicator through the (shimmed or real) MPI: rank 0 creates the id and
   broadcasts it as MPI_BYTE-sized ints. */
static void backend_comm_init(void)
{
  int ranks = 1, i;
  int idbuf[EXB_UNIQUE_ID_BYTES / sizeof(int)];
  MPI_Comm_size(MPI_COMM_WORLD, &ranks);
  g_ranks = ranks;
  if (ranks <= 1) return;
  memset(idbuf, 0, sizeof idbuf);
  if (processID == 0 && exb_comm_unique_id(idbuf) != EXB_OK) backend_fail("exb_comm_unique_id");
  MPI_Bcast(idbuf, (int)(sizeof idbuf / sizeof(int)), MPI_INT, 0, MPI_COMM_WORLD);
  if (exb_comm_init(g_engine, processID, ranks, idbuf) != EXB_OK) backend_fail("exb_comm_init");
  (void)i;
}

/* EXAML_B200_SPR_STATS=1: call statistics at exit */
static long g_calls_eval_full = 0, g_calls_eval = 0, g_calls_newview = 0, g_calls_newview_deferred = 0, g_calls_makenewz = 0, g_entries_total = 0;
static void spr_report(void);

static void backend_init(tree *tr)
{
  int m;
  size_t nb = (size_t)tr->numBranches, slots = (size_t)tr->mxtips + 1;

  /* -S (SEV memory saving, *_GAPPED_SAVE kernels): handed to the engine, which keeps no CLV block for (node, tile)
     pairs that are all-gap below the node (exb_set_sev).  The host-side gap vectors the reference sets up for itself
     (axml.c:2085-2168) stay unused.  Under PSR the engine's arena stays dense. */
  if (tr->saveMemory && tr->rateHetModel == CAT && processID == 0)
    printf("\nB200 back-end: -S under PSR: CLVs stay dense in HBM (results are identical, the memory saving is not applied)\n");
  for (m = 0; m < tr->NumberOfModels; m++) {
    const pInfo *p = &tr->partitionData[m];
    if (p->states != 4 && p->states != 20) refuse("binary or multi-state data");
    if ((p->protModels == LG4M || p->protModels == LG4X) && tr->rateHetModel == CAT)
      refuse("the LG4M / LG4X protein models under PSR");
  }
  if (exb_create(&g_engine, local_device(), tr->mxtips, tr->NumberOfModels, tr->numBranches) != EXB_OK)
    backend_fail("exb_create");
  if (tr->saveMemory && exb_set_sev(g_engine, 1) != EXB_OK) backend_fail("exb_set_sev");
  if (tr->rateHetModel == CAT) {                 /* -m PSR: one rate category per site */
    if (exb_set_rate_model(g_engine, EXB_CAT, tr->maxCategories) != EXB_OK) backend_fail("exb_set_rate_model");
  }
  for (m = 0; m < tr->NumberOfModels; m++) {
    const pInfo *p = &tr->partitionData[m];
    if (exb_set_partition(g_engine, m, p->states, p->width, p->wgt, p->yResource) != EXB_OK)
      backend_fail("exb_set_partition");
  }
  if (exb_commit(g_engine) != EXB_OK) backend_fail("exb_commit");
  g_stale = (unsigned char *)calloc(2 * (size_t)tr->mxtips, 1);
  g_defer_newviews = g_stale && !getenv("EXAML_B200_EAGER_NEWVIEWS");
  exb_set_timing(g_engine, 0);
  if (getenv("EXAML_B200_SPR_STATS")) atexit(spr_report);
  backend_comm_init();
  g_entries = (exb_entry *)malloc(slots * sizeof(exb_entry));
  g_qz = (double *)malloc(slots * nb * sizeof(double));
  g_rz = (double *)malloc(slots * nb * sizeof(double));
  g_mask = (unsigned char *)malloc((size_t)tr->NumberOfModels);
  if (!g_entries || !g_qz || !g_rz || !g_mask) { fprintf(stderr, "out of memory\n"); MPI_Abort(MPI_COMM_WORLD, -1); }
}

/* The host rewrites the model constants without notice (optimizeModel.c:419-462).  But every such change
   invalidates all CLVs, so the host code itself always follows it with a FULL traversal (evaluateGeneric(tr,
   tr->start, TRUE)) before it does anything else with the tree -- a partial traversal after a model change would
   read stale CLVs in the reference too.  The constants (and, under PSR, the per-site categories, rewritten by
   optimizeRateCategories optimizeModel.c:2403 and at start-up / restart) are therefore handed over on full traversals
   and on the first call only: the thousands of small calls of a tree search cost nothing per partition.
   exb_set_model compares and only uploads what changed. */
static int g_constants_stale = 1;       /* model constants / PSR categories must be handed over (again) */
static int g_scan_valid = 0;            /* evaluatePartialGeneric: the batched rate scan below is current */

/* Guard for the caller invariant above (EXAML_B200_CHECK_CONSTANTS=1): a fingerprint of everything that is handed
   over, taken at hand-over time and re-computed on EVERY call.  A host that changed a constant and then issued a
   partial traversal without a full one in between is reported instead of silently computing with the old values. */
static int g_check_constants = -1;
static unsigned long long *g_fingerprint = NULL;     /* [NumberOfModels] */

static unsigned long long fp_mix(unsigned long long h, const void *data, size_t bytes)
{
  const unsigned char *b = (const unsigned char *)data;
  size_t i;
  for (i = 0; i < bytes; i++) { h ^= b[i]; h *= 1099511628211ULL; }     /* FNV-1a */
  return h;
}

static unsigned long long fingerprint_of(const tree *tr, int m)
{
  const pInfo *p = &tr->partitionData[m];
  const size_t S = (size_t)p->states, codes = S == 4 ? 16 : 23;
  unsigned long long h = 1469598103934665603ULL;
  if (p->protModels == LG4M || p->protModels == LG4X) {
    int c;
    for (c = 0; c < 4; c++) {
      h = fp_mix(h, p->EIGN_LG4[c], S * sizeof(double));
      h = fp_mix(h, p->EV_LG4[c], S * S * sizeof(double));
      h = fp_mix(h, p->EI_LG4[c], S * S * sizeof(double));
      h = fp_mix(h, p->tipVector_LG4[c], codes * S * sizeof(double));
    }
    h = fp_mix(h, p->weights, 4 * sizeof(double));
  } else {
    h = fp_mix(h, p->EIGN, S * sizeof(double));
    h = fp_mix(h, p->EV, S * S * sizeof(double));
    h = fp_mix(h, p->EI, S * S * sizeof(double));
    h = fp_mix(h, p->tipVector, codes * S * sizeof(double));
  }
  h = fp_mix(h, p->gammaRates, 4 * sizeof(double));
  if (tr->rateHetModel == CAT) {
    h = fp_mix(h, p->perSiteRates, (size_t)p->numberOfCategories * sizeof(double));
    h = fp_mix(h, p->rateCategory, (size_t)p->width * sizeof(int));
  }
  return h;
}

static void backend_sync(tree *tr)
{
  int m;
  if (!g_engine) backend_init(tr);
  g_scan_valid = 0;                     /* any other likelihood call ends a rate scan */
  for (m = 0; m < tr->NumberOfModels; m++) g_mask[m] = tr->td[0].executeModel[m] ? 1 : 0;
  if (g_check_constants < 0) {
    const char *v = getenv("EXAML_B200_CHECK_CONSTANTS");
    g_check_constants = v && atoi(v) != 0;
    if (g_check_constants) g_fingerprint = (unsigned long long *)calloc((size_t)tr->NumberOfModels, sizeof *g_fingerprint);
  }
  if (!g_constants_stale) {
    if (g_check_constants)
      for (m = 0; m < tr->NumberOfModels; m++)
        if (fingerprint_of(tr, m) != g_fingerprint[m]) {
          fprintf(stderr, "\nB200 back-end: the model constants of partition %d changed without a full traversal "
                          "(EXAML_B200_CHECK_CONSTANTS)\n", m);
          MPI_Abort(MPI_COMM_WORLD, -1);
          exit(-1);
        }
    return;
  }
  for (m = 0; m < tr->NumberOfModels; m++) {
    const pInfo *p = &tr->partitionData[m];
    if (p->protModels == LG4M || p->protModels == LG4X) {
      /* one eigen-system per GAMMA category: flatten pInfo's four pointers each */
      static double eign[4 * 20], ev[4 * 400], ei[4 * 400], tipv[4 * 460];
      int c;
      for (c = 0; c < 4; c++) {
        memcpy(eign + c * 20, p->EIGN_LG4[c], 20 * sizeof(double));
        memcpy(ev + c * 400, p->EV_LG4[c], 400 * sizeof(double));
        memcpy(ei + c * 400, p->EI_LG4[c], 400 * sizeof(double));
        memcpy(tipv + c * 460, p->tipVector_LG4[c], 460 * sizeof(double));
      }
      if (exb_set_model_lg4(g_engine, m, eign, ev, ei, tipv, p->gammaRates, p->weights) != EXB_OK)
        backend_fail("exb_set_model_lg4");
    } else if (exb_set_model(g_engine, m, p->EIGN, p->EV, p->EI, p->tipVector, p->gammaRates) != EXB_OK)
      backend_fail("exb_set_model");
    if (tr->rateHetModel == CAT &&
        exb_set_categories(g_engine, m, p->numberOfCategories, p->perSiteRates, p->rateCategory) != EXB_OK)
      backend_fail("exb_set_categories");
    if (g_check_constants) g_fingerprint[m] = fingerprint_of(tr, m);
  }
  g_constants_stale = 0;
}

/* Flatten td[0].ti[first .. count) into the C-ABI form; returns the number of entries. */
static int pack_descriptor(tree *tr, int first)
{
  const int nb = tr->numBranches;
  int i, b, n = 0;
  for (i = first; i < tr->td[0].count; i++, n++) {
    const traversalInfo *t = &tr->td[0].ti[i];
    g_entries[n].tipCase = t->tipCase;          /* ignored by the engine for the branch entry */
    g_entries[n].pNumber = t->pNumber;
    g_entries[n].qNumber = t->qNumber;
    g_entries[n].rNumber = t->rNumber;
    for (b = 0; b < nb; b++) {
      g_qz[(size_t)n * nb + b] = t->qz[b];
      g_rz[(size_t)n * nb + b] = t->rz[b];
    }
  }
  return n;
}

/* ---------------------------------------------------------------------------------------------- */
/* traversal descriptor                                                                           */
/* ---------------------------------------------------------------------------------------------- */

static void emit_entry(traversalInfo *ti, int *counter, int tipCase, nodeptr p, nodeptr q, nodeptr r, int numBranches)
{
  traversalInfo *t = &ti[*counter];
  int b;
  t->tipCase = tipCase;
  t->pNumber = p->number;
  t->qNumber = q->number;
  t->rNumber = r->number;
  for (b = 0; b < numBranches; b++) {
    t->qz[b] = q->z[b];
    t->rz[b] = r->z[b];
  }
  *counter += 1;
  if (g_stale) g_stale[p->number] = (unsigned char)g_emit_marks_stale;      /* about to be recomputed -- or deferred again */
}

/* Post-order list of the CLVs at and below p that have to be recomputed so that p's CLV looks away
   from p->back.  A child is descended into when it is not oriented towards p (x flag clear) or when
   a full traversal is requested.  A tip child is always reported as q. */
void computeTraversalInfo(nodeptr p, traversalInfo *ti, int *counter, int maxTips, int numBranches, boolean partialTraversal)
{
  nodeptr q, r;
  boolean qTip, rTip;

  if (isTip(p->number, maxTips)) return;
  q = p->next->back;
  r = p->next->next->back;
  qTip = isTip(q->number, maxTips);
  rTip = isTip(r->number, maxTips);

  if (qTip && rTip) {
    if (!p->x) getxnode(p);
    emit_entry(ti, counter, TIP_TIP, p, q, r, numBranches);
    return;
  }
  if (qTip || rTip) {
    nodeptr tip = qTip ? q : r, inner = qTip ? r : q;
    if (!inner->x || !partialTraversal || IS_STALE(inner))
      computeTraversalInfo(inner, ti, counter, maxTips, numBranches, partialTraversal);
    if (!p->x) getxnode(p);
    assert(p->x && inner->x);
    emit_entry(ti, counter, TIP_INNER, p, tip, inner, numBranches);
    return;
  }
  if (!q->x || !partialTraversal || IS_STALE(q)) computeTraversalInfo(q, ti, counter, maxTips, numBranches, partialTraversal);
  if (!r->x || !partialTraversal || IS_STALE(r)) computeTraversalInfo(r, ti, counter, maxTips, numBranches, partialTraversal);
  if (!p->x) getxnode(p);
  assert(p->x && q->x && r->x);
  emit_entry(ti, counter, INNER_INNER, p, q, r, numBranches);
}

boolean isGap(unsigned int *x, int pos) { return (x[pos / 32] & mask32[pos % 32]); }
boolean noGap(unsigned int *x, int pos) { return !(x[pos / 32] & mask32[pos % 32]); }

/* ---------------------------------------------------------------------------------------------- */
/* the *Iterative layer: one engine call each                                                     */
/* ---------------------------------------------------------------------------------------------- */

void newviewIterative(tree *tr, int startIndex)
{
  int n;
  backend_sync(tr);
  n = pack_descriptor(tr, startIndex);
  if (n > 0 && exb_newview(g_engine, g_entries, n, g_qz, g_rz, g_mask) != EXB_OK) backend_fail("exb_newview");
}

/* Unlike the reference's function of the same name this returns the GLOBALLY reduced per-partition
   log-likelihoods: the all-reduce of evaluateGeneric happens on the device, before the single
   device-to-host copy. */
void evaluateIterative(tree *tr)
{
  int n, m;
  backend_sync(tr);
  n = pack_descriptor(tr, 0);
  if (exb_evaluate(g_engine, g_entries, n, g_qz, g_rz, g_mask, tr->perPartitionLH) != EXB_OK) backend_fail("exb_evaluate");
  for (m = 0; m < tr->NumberOfModels; m++)
    if (tr->td[0].executeModel[m] && tr->perPartitionLH[m] != 0.0)
      assert(tr->perPartitionLH[m] < 0.0);       /* evaluateGenericSpecial.c:836 */
}

void makenewzIterative(tree *tr)
{
  int n;
  backend_sync(tr);
  n = pack_descriptor(tr, 0);
  if (exb_makenewz_prepare(g_engine, g_entries, n, g_qz, g_rz, g_mask) != EXB_OK) backend_fail("exb_makenewz_prepare");
}

/* Globally reduced derivatives (the all-reduce of topLevelMakenewz is inside exb_core). */
void execCore(tree *tr, volatile double *_dlnLdlz, volatile double *_d2lnLdlz2)
{
  double d1[NUM_BRANCHES], d2[NUM_BRANCHES];
  int b;
  backend_sync(tr);
  if (exb_core(g_engine, tr->td[0].parameterValues, g_mask, d1, d2) != EXB_OK) backend_fail("exb_core");
  for (b = 0; b < tr->numBranches; b++) {
    _dlnLdlz[b] = d1[b];
    _d2lnLdlz2[b] = d2[b];
  }
}

/* ---------------------------------------------------------------------------------------------- */
/* public entry points                                                                            */
/* ---------------------------------------------------------------------------------------------- */

static void set_all_models(tree *tr, boolean value)
{
  int m;
  for (m = 0; m < tr->NumberOfModels; m++) tr->executeModel[m] = value;
}

static boolean scan_serve_newview(tree *tr, nodeptr p);
static boolean scan_serve_evaluate(tree *tr, nodeptr p);
static void scan_flush(tree *tr);

void evaluateGeneric(tree *tr, nodeptr p, boolean fullTraversal)
{
  nodeptr q = p->back;
  traversalInfo *ti = tr->td[0].ti;
  double sum = 0.0;
  int b, m;

  if (!fullTraversal && scan_serve_evaluate(tr, p)) return;     /* a lazy insertion whose score is already in the table */
  scan_flush(tr);

  ti[0].pNumber = p->number;
  ti[0].qNumber = q->number;
  for (b = 0; b < tr->numBranches; b++) ti[0].qz[b] = q->z[b];
  tr->td[0].count = 1;

  if (fullTraversal) {
    assert(isTip(p->number, tr->mxtips));
    g_constants_stale = 1;
    computeTraversalInfo(q, ti, &tr->td[0].count, tr->mxtips, tr->numBranches, FALSE);
  } else {
    if (!p->x || IS_STALE(p)) computeTraversalInfo(p, ti, &tr->td[0].count, tr->mxtips, tr->numBranches, TRUE);
    if (!q->x || IS_STALE(q)) computeTraversalInfo(q, ti, &tr->td[0].count, tr->mxtips, tr->numBranches, TRUE);
  }
  storeExecuteMaskInTraversalDescriptor(tr);
  tr->td[0].traversalHasChanged = TRUE;

  if (fullTraversal) g_calls_eval_full++; else g_calls_eval++;
  g_entries_total += tr->td[0].count - 1;
  evaluateIterative(tr);            /* newviews + lnL at the branch + all-reduce, on the GPU */

  for (m = 0; m < tr->NumberOfModels; m++) sum += tr->perPartitionLH[m];
  tr->likelihood = sum;
  tr->td[0].traversalHasChanged = FALSE;
}

void newviewGeneric(tree *tr, nodeptr p, boolean masked)
{
  int m;
  if (isTip(p->number, tr->mxtips)) return;
  if (!masked && scan_serve_newview(tr, p)) return;             /* the new node of a lazy insertion the scan has scored */
  scan_flush(tr);

  /* Deferred: newviewGeneric leaves nothing behind that the host reads -- the CLVs are private to the replaced files
      -- except the x flags.  So the flags are updated exactly as always (the same computeTraversalInfo walk),
     but the listed nodes are only MARKED; whichever later call needs one of them finds it on its own walk (IS_STALE) and
     recomputes it, children first, inside its own descriptor.  The thousands of one- and two-entry newview calls of a
     smoothing pass thereby ride along with the makenewz / evaluate calls that follow them, and a CLV that is invalidated
     again before anybody looked at it is never computed.  Only with joint branch lengths: under -M the calls that would
     later pick the marked nodes up run under partition masks (makenewzGeneric(..., mask = TRUE) in update(),
     searchAlgo.c:150-157) and would recompute them for the unconverged partitions only, whereas this call is owed to all. */
  if (g_defer_newviews && !masked && tr->numBranches == 1) {
    g_emit_marks_stale = 1;
    tr->td[0].count = 0;
    computeTraversalInfo(p, tr->td[0].ti, &tr->td[0].count, tr->mxtips, tr->numBranches, TRUE);
    g_emit_marks_stale = 0;
    g_calls_newview_deferred++;
    return;
  }
  tr->td[0].count = 0;
  computeTraversalInfo(p, tr->td[0].ti, &tr->td[0].count, tr->mxtips, tr->numBranches, TRUE);
  tr->td[0].traversalHasChanged = TRUE;

  if (masked)
    for (m = 0; m < tr->NumberOfModels; m++) tr->executeModel[m] = tr->partitionConverged[m] ? FALSE : TRUE;

  g_calls_newview++;
  g_entries_total += tr->td[0].count;
  if (tr->td[0].count > 0) {
    storeExecuteMaskInTraversalDescriptor(tr);
    newviewIterative(tr, 0);
  }
  if (masked) set_all_models(tr, TRUE);
  tr->td[0].traversalHasChanged = FALSE;
}

/* Newton-Raphson on z for the branch in td[0].ti[0]; one branch slot per partition under -M,
   a single joint slot otherwise.  Per slot the iteration is in one of three states:
     looking for negative curvature (curvatOK == FALSE), stepping (curvatOK == TRUE), done. */
static void newton_raphson(tree *tr, double *z0, int maxiterStart, double *result)
{
  const int nb = tr->numBranches;
  double z[NUM_BRANCHES], zprev[NUM_BRANCHES], zstep[NUM_BRANCHES];
  double d1[NUM_BRANCHES], d2[NUM_BRANCHES];
  int budget[NUM_BRANCHES];
  boolean done[NUM_BRANCHES], allDone, sumTableReady = FALSE;
  int i, m;

  for (i = 0; i < nb; i++) {
    z[i] = z0[i];
    budget[i] = maxiterStart;
    done[i] = FALSE;
    tr->curvatOK[i] = TRUE;
  }

  do {
    for (i = 0; i < nb; i++)
      if (!done[i] && tr->curvatOK[i]) {            /* start a new NR step from z */
        tr->curvatOK[i] = FALSE;
        zprev[i] = z[i];
        zstep[i] = (1.0 - zmax) * z[i] + zmin;
      }
    for (i = 0; i < nb; i++)
      if (!done[i] && !tr->curvatOK[i]) {
        if (z[i] < zmin) z[i] = zmin;
        else if (z[i] > zmax) z[i] = zmax;
        tr->coreLZ[i] = log(z[i]);
      }

    /* only slots that still need derivatives are executed */
    if (nb > 1) {
      assert(nb == tr->NumberOfModels);
      for (m = 0; m < tr->NumberOfModels; m++)
        if (tr->executeModel[m]) tr->executeModel[m] = !tr->curvatOK[m];
    } else {
      set_all_models(tr, !tr->curvatOK[0]);
    }
    storeExecuteMaskInTraversalDescriptor(tr);
    storeValuesInTraversalDescriptor(tr, &(tr->coreLZ[0]));

    if (!sumTableReady) {
      makenewzIterative(tr);          /* newviews + sum table, once per branch */
      sumTableReady = TRUE;
    }
    execCore(tr, d1, d2);             /* globally reduced */

    for (i = 0; i < nb; i++)
      if (!done[i] && !tr->curvatOK[i]) {
        if (d2[i] >= 0.0 && z[i] < zmax)
          zprev[i] = z[i] = 0.37 * z[i] + 0.63;     /* wrong side of the maximum: shorten the branch */
        else
          tr->curvatOK[i] = TRUE;
      }

    for (i = 0; i < nb; i++)
      if (tr->curvatOK[i] && !done[i]) {
        if (d2[i] < 0.0) {
          const double step = -d1[i] / d2[i];
          const double cap = 0.25 * zprev[i] + 0.75;
          if (step < 100) {
            z[i] *= exp(step);
            if (z[i] < zmin) z[i] = zmin;
            if (z[i] > cap) z[i] = cap;
          } else
            z[i] = cap;
        }
        if (z[i] > zmax) z[i] = zmax;
        budget[i] -= 1;
        if (fabs(z[i] - zprev[i]) > zstep[i]) {
          if (budget[i] < -20) {                    /* give up: keep the old branch length */
            z[i] = z0[i];
            done[i] = TRUE;
          }
        } else
          done[i] = TRUE;
      }

    allDone = TRUE;
    for (i = 0; i < nb; i++) allDone = allDone && done[i];
  } while (!allDone);

  set_all_models(tr, TRUE);
  for (i = 0; i < nb; i++) result[i] = z[i];
}

void makenewzGeneric(tree *tr, nodeptr p, nodeptr q, double *z0, int maxiter, double *result, boolean mask)
{
  traversalInfo *ti = tr->td[0].ti;
  int i;

  scan_flush(tr);

  ti[0].pNumber = p->number;
  ti[0].qNumber = q->number;
  for (i = 0; i < tr->numBranches; i++) {
    ti[0].qz[i] = z0[i];
    if (mask)
      tr->executeModel[i] = tr->partitionConverged[i] ? FALSE : TRUE;
    else
      assert(tr->executeModel[i]);
  }
  tr->td[0].count = 1;
  if (!p->x || IS_STALE(p)) computeTraversalInfo(p, ti, &tr->td[0].count, tr->mxtips, tr->numBranches, TRUE);
  if (!q->x || IS_STALE(q)) computeTraversalInfo(q, ti, &tr->td[0].count, tr->mxtips, tr->numBranches, TRUE);
  g_calls_makenewz++;
  g_entries_total += tr->td[0].count - 1;

  /* the whole Newton-Raphson loop on the device: single rank, or all ranks on one node (in-kernel all-reduce) */
  if ((g_ranks == 1 || exb_comm_is_p2p(g_engine)) && !getenv("EXAML_B200_HOST_NR")) {
    /* single rank: the whole Newton-Raphson loop runs on the device (exb_makenewz) */
    int cok[NUM_BRANCHES], n, m;
    storeExecuteMaskInTraversalDescriptor(tr);
    backend_sync(tr);
    n = pack_descriptor(tr, 0);
    if (exb_makenewz(g_engine, g_entries, n, g_qz, g_rz, g_mask, z0, maxiter, result, cok, NULL) != EXB_OK)
      backend_fail("exb_makenewz");
    for (i = 0; i < tr->numBranches; i++) tr->curvatOK[i] = cok[i] ? TRUE : FALSE;
    for (m = 0; m < tr->NumberOfModels; m++) tr->executeModel[m] = TRUE;
  } else
    newton_raphson(tr, z0, maxiter, result);

  for (i = 0; i < tr->numBranches; i++) tr->executeModel[i] = TRUE;
}


/* ---------------------------------------------------------------------------------------------- */
/* batched scoring of lazy SPR insertions                                           */
/* ---------------------------------------------------------------------------------------------- */
/* rearrangeBIG (searchAlgo.c:800) prunes the subtree at p and lets addTraverseBIG (searchAlgo.c:785) try it in every
 * branch within the rearrangement radius, one testInsertBIG (searchAlgo.c:682) at a time: hook p in with both halves at
 * sqrt(z) (insertBIG, non-thorough branch, searchAlgo.c:519-535), newviewGeneric(p), evaluateGeneric(p->next->next),
 * unhook.  On a GPU that is two small calls per candidate.  The walk is predictable, though: it depends on the topology
 * only (the lhCutoff test merely prunes it), and the score of a candidate does not depend on the candidates before it.
 * A maintainer therefore brackets the four addTraverseBIG calls of each side of rearrangeBIG with
 *
 *     batchedInsertionScanBegin(tr, p, p1, p2, mintrav, maxtrav);   ...addTraverseBIG...   batchedInsertionScanEnd(tr);
 *
 * (INTEGRATION.md shows the patch; without it nothing below is ever active).  Begin walks the same branches in the same
 * order, builds ONE mixed descriptor -- the CLVs looking from the pruning point towards each candidate go into the
 * engine's scratch nodes level by level, subtrees that do not look at the walk get their ordinary computeTraversalInfo
 * entries, one scoring entry per candidate -- and obtains all log likelihoods with exb_score_insertions.  The unchanged
 * testInsertBIG then runs as always: its newviewGeneric(p) and evaluateGeneric(p->next->next) are answered from the
 * table (all bookkeeping -- bestOfNode, endLH, lhAVG / lhDEC, the cutoff -- stays the reference's code, it just reads
 * tr->likelihood).  The tree's own CLVs and x flags are not re-oriented by the scan, so after End the state is the one
 * right after removeNodeBIG.  Anything that does not match the table (a thorough insertion, a different node) takes the
 * ordinary path; a newview that was answered from the table but is then needed for real is computed first (scan_flush). */
extern int Thorough;                               /* globalVariables.h; set by computeBIGRAPID, searchAlgo.c:1984 */

typedef struct {
  int active;
  nodeptr p;                    /* the node whose subtree p->back is being re-inserted */
  int ncand, cursor;
  nodeptr *cand;                /* [ncand] the record q of candidate branch (q, q->back), in addTraverseBIG order */
  double *lnl;                  /* [ncand][NumberOfModels] */
  int pending;                  /* candidate whose newviewGeneric(p) was answered from the table; -1: none */
  /* descriptor under construction */
  exb_entry *ent; double *qz, *rz; int n, cap, ccap;
  int failed;
} scan_state;
static scan_state g_spr = {0, NULL, 0, 0, NULL, NULL, -1, NULL, NULL, NULL, 0, 0, 0, 0};
/* EXAML_B200_SPR_STATS=1: one line at exit (tests assert that the insertions really came from the table) */
static long g_spr_scans = 0, g_spr_scored = 0, g_spr_served = 0, g_spr_flushed = 0, g_spr_calls = 0;
static void spr_report(void)
{
  if (processID == 0)
    printf("B200 back-end, likelihood calls: %ld evaluateGeneric (full) + %ld (partial), %ld newviewGeneric (+ %ld deferred), "
           "%ld makenewzGeneric; %ld descriptor entries\n", g_calls_eval_full, g_calls_eval, g_calls_newview,
           g_calls_newview_deferred, g_calls_makenewz, g_entries_total);
  if (processID == 0)
    printf("B200 back-end, batched insertion scoring: %ld scans, %ld engine calls, %ld candidates scored, %ld served from the table, "
           "%ld answered newviews computed after all\n", g_spr_scans, g_spr_calls, g_spr_scored, g_spr_served, g_spr_flushed);
}

static void spr_reserve(tree *tr, int more)
{
  const size_t nb = (size_t)tr->numBranches;
  if (g_spr.n + more <= g_spr.cap) return;
  g_spr.cap = 2 * (g_spr.n + more) + 64;
  g_spr.ent = (exb_entry *)realloc(g_spr.ent, (size_t)g_spr.cap * sizeof(exb_entry));
  g_spr.qz = (double *)realloc(g_spr.qz, (size_t)g_spr.cap * nb * sizeof(double));
  g_spr.rz = (double *)realloc(g_spr.rz, (size_t)g_spr.cap * nb * sizeof(double));
  if (!g_spr.ent || !g_spr.qz || !g_spr.rz) { fprintf(stderr, "out of memory\n"); MPI_Abort(MPI_COMM_WORLD, -1); exit(-1); }
}

/* append an entry with children (a, az) and (b, bz); a tip child goes into q (newviewGenericSpecial.c:742-750) */
static void spr_emit(tree *tr, int scoreBit, int pNumber, int a, const double *az, int b, const double *bz)
{
  const int nb = tr->numBranches;
  const boolean at = isTip(a, tr->mxtips), bt = isTip(b, tr->mxtips);
  exb_entry *t;
  int i;
  if (bt && !at) { int x = a; const double *xz = az; a = b; az = bz; b = x; bz = xz; }
  spr_reserve(tr, 1);
  t = &g_spr.ent[g_spr.n];
  t->tipCase = ((at && bt) ? TIP_TIP : ((at || bt) ? TIP_INNER : INNER_INNER)) | scoreBit;
  t->pNumber = pNumber; t->qNumber = a; t->rNumber = b;
  for (i = 0; i < nb; i++) { g_spr.qz[(size_t)g_spr.n * nb + i] = az[i]; g_spr.rz[(size_t)g_spr.n * nb + i] = bz[i]; }
  g_spr.n++;
}

/* make the CLV of rec's node look along rec, the way evaluateGeneric does on demand */
static void spr_orient(tree *tr, nodeptr rec)
{
  int cnt = 0, i, b;
  const int nb = tr->numBranches;
  if (isTip(rec->number, tr->mxtips) || (rec->x && !IS_STALE(rec))) return;
  computeTraversalInfo(rec, tr->td[0].ti, &cnt, tr->mxtips, tr->numBranches, TRUE);   /* td[0] is free between calls */
  spr_reserve(tr, cnt);
  for (i = 0; i < cnt; i++) {
    const traversalInfo *t = &tr->td[0].ti[i];
    exb_entry *e = &g_spr.ent[g_spr.n];
    e->tipCase = t->tipCase; e->pNumber = t->pNumber; e->qNumber = t->qNumber; e->rNumber = t->rNumber;
    for (b = 0; b < nb; b++) { g_spr.qz[(size_t)g_spr.n * nb + b] = t->qz[b]; g_spr.rz[(size_t)g_spr.n * nb + b] = t->rz[b]; }
    g_spr.n++;
  }
}

/* addTraverseBIG's walk (searchAlgo.c:785-799).  up: the record of r's ring that looks back at the pruning point,
   upNumber: the node (real at the top level, scratch below) whose CLV comes from there */
static void spr_walk(tree *tr, nodeptr q, int mintrav, int maxtrav, int depth, nodeptr up, int upNumber)
{
  const int nb = tr->numBranches;
  nodeptr r = q->back, sibRec, sib;
  int uNumber, i;
  if (g_spr.failed) return;
  if (depth >= EXB_SCRATCH_NODES) { g_spr.failed = 1; return; }
  sibRec = (r->next == up) ? r->next->next : r->next;     /* the third record of r's ring */
  sib = sibRec->back;
  spr_orient(tr, sib);
  uNumber = 2 * tr->mxtips - 1 + depth;
  spr_emit(tr, 0, uNumber, upNumber, up->z, sib->number, sibRec->z);
  if (--mintrav <= 0) {
    double zz[NUM_BRANCHES];
    spr_orient(tr, q);
    for (i = 0; i < nb; i++) {                     /* insertBIG searchAlgo.c:522-531 */
      zz[i] = sqrt(q->z[i]);
      if (zz[i] < zmin) zz[i] = zmin;
      if (zz[i] > zmax) zz[i] = zmax;
    }
    if (g_spr.ncand == g_spr.ccap) {
      g_spr.ccap = 2 * g_spr.ccap + 64;
      g_spr.cand = (nodeptr *)realloc(g_spr.cand, (size_t)g_spr.ccap * sizeof(nodeptr));
      if (!g_spr.cand) { fprintf(stderr, "out of memory\n"); MPI_Abort(MPI_COMM_WORLD, -1); exit(-1); }
    }
    spr_emit(tr, EXB_SCORE, g_spr.ncand, q->number, zz, uNumber, zz);
    g_spr.cand[g_spr.ncand++] = q;
  }
  if (!isTip(q->number, tr->mxtips) && --maxtrav > 0) {
    spr_walk(tr, q->next->back, mintrav, maxtrav, depth + 1, q, uNumber);
    spr_walk(tr, q->next->next->back, mintrav, maxtrav, depth + 1, q, uNumber);
  }
}

/* The scan's list is not going to be executed (too deep, nothing to score, engine refused): the tree's own nodes it meant
   to recompute have their x flags set by the walk already -- put them on the stale list so that the next call that needs
   them recomputes them. */
static void spr_abandon(tree *tr)
{
  const int nb = tr->numBranches;
  int i, n = 0, m;
  if (nb == 1) {
    for (i = 1; i < g_spr.n; i++)
      if (!(g_spr.ent[i].tipCase & EXB_SCORE) && g_spr.ent[i].pNumber <= 2 * tr->mxtips - 2)
        g_stale[g_spr.ent[i].pNumber] = 1;
    return;
  }
  /* -M: later calls may run under partition masks, so nothing is left for them: the tree's own nodes of the list are
     recomputed now, for all partitions (their entries are in execution order; they never read a scratch node) */
  for (i = 1; i < g_spr.n; i++)
    if (!(g_spr.ent[i].tipCase & EXB_SCORE) && g_spr.ent[i].pNumber <= 2 * tr->mxtips - 2) {
      g_spr.ent[n] = g_spr.ent[i];
      memmove(g_spr.qz + (size_t)n * nb, g_spr.qz + (size_t)i * nb, (size_t)nb * sizeof(double));
      memmove(g_spr.rz + (size_t)n * nb, g_spr.rz + (size_t)i * nb, (size_t)nb * sizeof(double));
      n++;
    }
  for (m = 0; m < tr->NumberOfModels; m++) g_mask[m] = 1;
  if (n > 0 && exb_newview(g_engine, g_spr.ent, n, g_spr.qz, g_spr.rz, g_mask) != EXB_OK) backend_fail("exb_newview");
  g_spr.n = 0;
}

void batchedInsertionScanBegin(tree *tr, nodeptr p, nodeptr p1, nodeptr p2, int mintrav, int maxtrav)
{
  static int enabled = -1;
  const int nb = tr->numBranches, P = tr->NumberOfModels;
  nodeptr s = p->back, side[2][2];
  int k, m, at, i, cap;
  g_spr.active = 0;
  g_spr.pending = -1;
  if (enabled < 0) enabled = getenv("EXAML_B200_NO_BATCHED_SPR") ? 0 : 1;
  if (!enabled || Thorough) return;
  backend_sync(tr);
  if (!exb_can_score_insertions(g_engine)) return;
  g_spr.n = 0; g_spr.ncand = 0; g_spr.failed = 0;
  /* entry 0: the pruned subtree and the branch p--s (the diagonal table of the evaluation) */
  spr_reserve(tr, 1);
  g_spr.ent[0].tipCase = -1; g_spr.ent[0].pNumber = s->number; g_spr.ent[0].qNumber = s->number; g_spr.ent[0].rNumber = 0;
  for (i = 0; i < nb; i++) { g_spr.qz[i] = p->z[i]; g_spr.rz[i] = 1.0; }
  g_spr.n = 1;
  spr_orient(tr, s);
  side[0][0] = p1; side[0][1] = p2; side[1][0] = p2; side[1][1] = p1;
  for (k = 0; k < 2; k++) {
    nodeptr a = side[k][0], b = side[k][1];
    if (isTip(a->number, tr->mxtips)) continue;
    spr_orient(tr, b);                             /* b's CLV must look at a */
    spr_walk(tr, a->next->back, mintrav, maxtrav, 0, a, b->number);
    spr_walk(tr, a->next->next->back, mintrav, maxtrav, 0, a, b->number);
  }
  if (g_spr.failed || g_spr.ncand == 0) { spr_abandon(tr); return; }
  g_spr.lnl = (double *)realloc(g_spr.lnl, (size_t)g_spr.ncand * P * sizeof(double));
  if (!g_spr.lnl) { fprintf(stderr, "out of memory\n"); MPI_Abort(MPI_COMM_WORLD, -1); exit(-1); }
  for (m = 0; m < P; m++) g_mask[m] = tr->executeModel[m] ? 1 : 0;
  /* a descriptor holds exb_max_entries entries: cut the list anywhere (scratch CLVs persist), candidates renumbered per call */
  cap = exb_max_entries(g_engine) - 1;
  for (at = 1; at < g_spr.n; ) {
    const int n = (g_spr.n - at < cap) ? g_spr.n - at : cap;
    int first = -1, cnt = 0;
    const exb_entry e0 = g_spr.ent[at - 1];
    double q0[NUM_BRANCHES], r0[NUM_BRANCHES];
    /* slot at-1 temporarily becomes entry 0 of this call */
    memcpy(q0, g_spr.qz + (size_t)(at - 1) * nb, (size_t)nb * sizeof(double));
    memcpy(r0, g_spr.rz + (size_t)(at - 1) * nb, (size_t)nb * sizeof(double));
    g_spr.ent[at - 1] = g_spr.ent[0];
    memcpy(g_spr.qz + (size_t)(at - 1) * nb, g_spr.qz, (size_t)nb * sizeof(double));
    for (i = at; i < at + n; i++)
      if (g_spr.ent[i].tipCase & EXB_SCORE) { if (first < 0) first = g_spr.ent[i].pNumber; g_spr.ent[i].pNumber -= first; cnt++; }
    m = exb_score_insertions(g_engine, g_spr.ent + at - 1, n + 1, g_spr.qz + (size_t)(at - 1) * nb, g_spr.rz + (size_t)(at - 1) * nb,
                             g_mask, cnt, cnt ? g_spr.lnl + (size_t)first * P : NULL);
    g_spr_calls++;
    g_spr.ent[at - 1] = e0;
    memcpy(g_spr.qz + (size_t)(at - 1) * nb, q0, (size_t)nb * sizeof(double));
    memcpy(g_spr.rz + (size_t)(at - 1) * nb, r0, (size_t)nb * sizeof(double));
    if (m != EXB_OK) {
      static int told = 0;
      if (!told && processID == 0) printf("\nB200 back-end: batched insertion scoring not used (%s)\n", exb_last_error());
      told = 1;
      /* (entries of earlier pieces were executed; recomputing them once more is harmless) */
      spr_abandon(tr);
      return;                                      /* candidates are scored one by one as before */
    }
    at += n;
  }
  g_spr.p = p; g_spr.cursor = 0; g_spr.active = 1;
  g_spr_scans++; g_spr_scored += g_spr.ncand;
}

void batchedInsertionScanEnd(tree *tr)
{
  scan_flush(tr);
  g_spr.active = 0;
}

static int scan_lookup(nodeptr q)
{
  int i;
  for (i = 0; i < g_spr.ncand; i++) {
    const int k = (g_spr.cursor + i) % g_spr.ncand;
    if (g_spr.cand[k] == q) { g_spr.cursor = k; return k; }
  }
  return -1;
}

/* newviewGeneric(tr, p, FALSE) of insertBIG: p hooked between a candidate's ends with the lazy branch values? */
static boolean scan_serve_newview(tree *tr, nodeptr p)
{
  int k, i;
  nodeptr q, r;
  if (!g_spr.active || p != g_spr.p || Thorough) return FALSE;
  scan_flush(tr);
  q = p->next->back; r = p->next->next->back;
  if (!q || !r) return FALSE;
  k = scan_lookup(q);
  if (k < 0) return FALSE;
  for (i = 0; i < tr->numBranches; i++)            /* both halves equal: the lazy insertion, not the restored original */
    if (p->next->z[i] != p->next->next->z[i]) return FALSE;
  g_spr.pending = k;
  return TRUE;
}

static boolean scan_serve_evaluate(tree *tr, nodeptr pnn)
{
  int m, k = g_spr.pending;
  double sum = 0.0;
  if (!g_spr.active || k < 0) return FALSE;
  if (pnn != g_spr.p->next->next || g_spr.p->next->back != g_spr.cand[k]) return FALSE;
  for (m = 0; m < tr->NumberOfModels; m++) {
    tr->perPartitionLH[m] = g_spr.lnl[(size_t)k * tr->NumberOfModels + m];
    sum += tr->perPartitionLH[m];
  }
  tr->likelihood = sum;
  g_spr_served++;
  g_spr.pending = -1;           /* the caller unhooks p next: its CLV never has to exist */
  return TRUE;
}

/* A newview answered from the table is needed after all: compute it now. */
static void scan_flush(tree *tr)
{
  if (g_spr.pending >= 0) {
    nodeptr p = g_spr.p;
    const int was = g_spr.active;
    g_spr.pending = -1;
    g_spr.active = 0;
    g_spr_flushed++;
    if (p->next->back && p->next->next->back) newviewGeneric(tr, p, FALSE);
    g_spr.active = was;
  }
}

/* ---------------------------------------------------------------------------------------------- */
/* evaluatePartialGeneric: the per-site rate scan of -m PSR                                        */
/* ---------------------------------------------------------------------------------------------- */
/* optimizeRateCategories (optimizeModel.c:2403) walks over the sites one by one (optRateCatPthreads,
 * optimizeModel.c:1792-1922): evaluatePartialGeneric at the site's current rate, then at rate - k*lower_spacing
 * for k = 1, 2, ... while that improves, then at rate + k*upper_spacing likewise.  Each call is a whole
 * traversal for ONE site -- far too little work for a GPU.  But the questions are predictable: the first call
 * of a scan tells us that every site will be asked for its current rate, the first call on each side tells us
 * the spacing, and the caller's stopping rule depends only on the values we return.  So the answers are
 * computed ahead, for all sites of all partitions at once, round by round (round k = step k of every site whose
 * scan is still running), through exb_evaluate_partial; the calls of the caller are then served from that table.
 * A question that is not in the table (it never happened in our runs) is answered by a single-pair launch.
 * Nothing here touches patrat / lhs: the caller's loop still takes every decision itself. */
typedef struct {
  double *rate0, *v0;           /* [width] rate of the site when the scan began, value there */
  size_t *off[2];               /* [width + 1] CSR offsets into val[side] */
  double *val[2];               /* values at rate0 -/+ k * spacing, k = 1 ..  */
} scan_part;
static scan_part *g_scan = NULL;
static int g_scan_models = 0, g_scan_side_ready[2] = {0, 0};
static double g_scan_spacing[2] = {0.0, 0.0};
static int g_scan_count = 0;     /* packed descriptor of the scan (g_entries / g_qz / g_rz) */

static double scan_rate(double r0, int side, int k, double spacing)
{
  return side == 0 ? r0 - k * spacing : r0 + k * spacing;        /* optimizeModel.c:1876, 1892 */
}

static void scan_eval(size_t n, const int *models, const int *sites, const double *rates, double *out)
{
  if (exb_evaluate_partial(g_engine, g_entries, g_scan_count, g_qz, g_rz, n, models, sites, rates, out) != EXB_OK)
    backend_fail("exb_evaluate_partial");
}

static void scan_begin(tree *tr)
{
  int m, s;
  size_t total = 0, at = 0, i;
  int *models, *sites;
  double *rates, *vals;
  backend_sync(tr);                                   /* current model constants; also marks the old scan invalid */
  g_scan_count = pack_descriptor(tr, 0);              /* td[0] still holds the full traversal of evaluateGeneric(tr, start, TRUE) */
  if (!g_scan || g_scan_models != tr->NumberOfModels) {
    g_scan = (scan_part *)calloc((size_t)tr->NumberOfModels, sizeof(scan_part));
    g_scan_models = tr->NumberOfModels;
  }
  for (m = 0; m < tr->NumberOfModels; m++) total += tr->partitionData[m].width;
  models = (int *)malloc((total + 1) * sizeof(int)); sites = (int *)malloc((total + 1) * sizeof(int));
  rates = (double *)malloc((total + 1) * sizeof(double)); vals = (double *)malloc((total + 1) * sizeof(double));
  for (m = 0; m < tr->NumberOfModels; m++) {
    const pInfo *p = &tr->partitionData[m];
    scan_part *sp = &g_scan[m];
    free(sp->rate0); free(sp->v0);
    for (s = 0; s < 2; s++) { free(sp->off[s]); free(sp->val[s]); sp->off[s] = NULL; sp->val[s] = NULL; }
    sp->rate0 = (double *)malloc((p->width + 1) * sizeof(double));
    sp->v0 = (double *)malloc((p->width + 1) * sizeof(double));
    for (i = 0; i < p->width; i++, at++) {
      sp->rate0[i] = p->patrat[i];
      models[at] = m; sites[at] = (int)i; rates[at] = p->patrat[i];
    }
  }
  scan_eval(total, models, sites, rates, vals);
  for (m = 0, at = 0; m < tr->NumberOfModels; m++) {
    memcpy(g_scan[m].v0, vals + at, tr->partitionData[m].width * sizeof(double));
    at += tr->partitionData[m].width;
  }
  free(models); free(sites); free(rates); free(vals);
  g_scan_side_ready[0] = g_scan_side_ready[1] = 0;
  g_scan_spacing[0] = g_scan_spacing[1] = 0.0;
  g_scan_valid = 1;
}

/* Which spacing turns r0 into ki in one step?  The caller's choices are 0.5 / 1.0 on its first invocation and
   0.05/n / 0.1/n afterwards, never below 0.001 (optimizeModel.c:2429-2444). */
static double scan_infer_spacing(int side, double r0, double ki)
{
  int n;
  for (n = 1; n <= 256; n++) {
    double sp = side == 0 ? (n == 1 ? 0.5 : 0.05) / ((double)n) : (n == 1 ? 1.0 : 0.1) / ((double)n);
    if (sp < 0.001) sp = 0.001;
    if (scan_rate(r0, side, 1, sp) == ki) return sp;
  }
  return 0.0;
}

/* All remaining steps of one side, for every site, in rounds. */
static void scan_side(tree *tr, int side, double spacing)
{
  const double epsilon = 0.00001;                     /* optimizeModel.c:1861 */
  size_t total = 0, nact = 0, at, i, j, rounds = 0, rcap = 0;
  int m, *models, *sites, **rsites_m = NULL, **rsites_i = NULL;
  double *rates, *vals, *best, **rvals = NULL;
  size_t *rn = NULL;
  for (m = 0; m < tr->NumberOfModels; m++) total += tr->partitionData[m].width;
  models = (int *)malloc((total + 1) * sizeof(int)); sites = (int *)malloc((total + 1) * sizeof(int));
  rates = (double *)malloc((total + 1) * sizeof(double)); vals = (double *)malloc((total + 1) * sizeof(double));
  best = (double *)malloc((total + 1) * sizeof(double));
  for (m = 0, at = 0; m < tr->NumberOfModels; m++) {
    g_scan[m].off[side] = (size_t *)calloc(tr->partitionData[m].width + 2, sizeof(size_t));
    for (i = 0; i < tr->partitionData[m].width; i++, at++) { models[at] = m; sites[at] = (int)i; best[at] = g_scan[m].v0[i]; }
  }
  nact = total;
  while (nact > 0) {
    const int k = (int)rounds + 1;
    size_t keep = 0;
    /* who takes step k?  left: only while the rate stays above 0.0001 (optimizeModel.c:1875) */
    for (j = 0, at = 0; j < nact; j++) {
      const double r = scan_rate(g_scan[models[j]].rate0[sites[j]], side, k, spacing);
      if (side == 0 && !(r > 0.0001)) continue;
      models[at] = models[j]; sites[at] = sites[j]; best[at] = best[j]; rates[at] = r; at++;
    }
    nact = at;
    if (nact == 0) break;
    scan_eval(nact, models, sites, rates, vals);
    if (rounds == rcap) {
      rcap = rcap ? 2 * rcap : 16;
      rsites_m = (int **)realloc(rsites_m, rcap * sizeof(int *)); rsites_i = (int **)realloc(rsites_i, rcap * sizeof(int *));
      rvals = (double **)realloc(rvals, rcap * sizeof(double *)); rn = (size_t *)realloc(rn, rcap * sizeof(size_t));
    }
    rsites_m[rounds] = (int *)malloc(nact * sizeof(int)); rsites_i[rounds] = (int *)malloc(nact * sizeof(int));
    rvals[rounds] = (double *)malloc(nact * sizeof(double)); rn[rounds] = nact;
    memcpy(rsites_m[rounds], models, nact * sizeof(int)); memcpy(rsites_i[rounds], sites, nact * sizeof(int));
    memcpy(rvals[rounds], vals, nact * sizeof(double));
    for (j = 0; j < nact; j++) {
      g_scan[models[j]].off[side][sites[j] + 1] += 1;            /* counts first, prefix sums below */
      if (vals[j] > best[j] && fabs(best[j] - vals[j]) > epsilon) {   /* the caller goes on (optimizeModel.c:1876-1878, 1892-1893) */
        models[keep] = models[j]; sites[keep] = sites[j]; best[keep] = vals[j]; keep++;
      }
    }
    nact = keep;
    rounds++;
  }
  for (m = 0; m < tr->NumberOfModels; m++) {
    scan_part *sp = &g_scan[m];
    const size_t w = tr->partitionData[m].width;
    for (i = 0; i < w; i++) sp->off[side][i + 1] += sp->off[side][i];
    sp->val[side] = (double *)malloc((sp->off[side][w] + 1) * sizeof(double));
  }
  for (i = 0; i < rounds; i++) {                       /* round i holds step i+1 of its sites */
    for (j = 0; j < rn[i]; j++) {
      scan_part *sp = &g_scan[rsites_m[i][j]];
      sp->val[side][sp->off[side][rsites_i[i][j]] + i] = rvals[i][j];
    }
    free(rsites_m[i]); free(rsites_i[i]); free(rvals[i]);
  }
  free(rsites_m); free(rsites_i); free(rvals); free(rn);
  free(models); free(sites); free(rates); free(vals); free(best);
}

double evaluatePartialGeneric(tree *tr, int i, double ki, int _model)
{
  const scan_part *sp;
  double r0, v;
  int side, k, nsteps;
  assert(tr->rateHetModel == CAT);                    /* evaluatePartialGenericSpecial.c:326 */
  if (!g_scan_valid) scan_begin(tr);
  sp = &g_scan[_model];
  r0 = sp->rate0[i];
  if (ki == r0) return sp->v0[i];
  side = ki < r0 ? 0 : 1;
  if (!g_scan_side_ready[side]) {
    g_scan_spacing[side] = scan_infer_spacing(side, r0, ki);
    if (g_scan_spacing[side] > 0.0) scan_side(tr, side, g_scan_spacing[side]);
    g_scan_side_ready[side] = 1;
  }
  if (g_scan_spacing[side] > 0.0) {
    nsteps = (int)(sp->off[side][i + 1] - sp->off[side][i]);
    for (k = 1; k <= nsteps; k++)
      if (scan_rate(r0, side, k, g_scan_spacing[side]) == ki) return sp->val[side][sp->off[side][i] + k - 1];
  }
  scan_eval(1, &_model, &i, &ki, &v);                 /* not anticipated: answer this one question */
  return v;
}
