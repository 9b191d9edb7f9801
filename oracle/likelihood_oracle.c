/* likelihood_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded restatement of the reference's likelihood hot path, used only
 * as the checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Nothing
 * under examl_b200/ may link, load or call it.
 *
 * Each function follows the reference's readable "FLEX" code (or the SSE3 kernel where the
 * FLEX one does not exist) and cites it:
 *   orc_make_p          newviewGenericSpecial.c:78-132  (makeP) + :983-984 (z clamp)
 *   orc_newview_site*   newviewGenericSpecial.c:430-682 (newviewGAMMA_FLEX), scaler :1503-1506
 *   orc_diag            evaluateGenericSpecial.c:80-119 (calcDiagptable)
 *   evaluate            evaluateGenericSpecial.c:154-228 (evaluateGAMMA_FLEX), scaler undo :830
 *   sum table           makenewzGenericSpecial.c:251-319 (sumGAMMA_FLEX)
 *   core                makenewzGenericSpecial.c:2309-2393 (coreGTRGAMMA; coreGAMMA_FLEX :523
 *                       does not compile in the reference)
 *   CAT / PSR           newviewCAT_FLEX newviewGenericSpecial.c:217-427, evaluateCAT_FLEX
 *                       evaluateGenericSpecial.c:232-330, sumCAT_FLEX makenewzGenericSpecial.c:170-249,
 *                       coreCAT_FLEX makenewzGenericSpecial.c:394-519 (one rate category per site)
 *   orc_score_insertions testInsertBIG searchAlgo.c:682 in its non-thorough form (insertBIG searchAlgo.c:519-535 +
 *                       newviewGeneric + evaluateGeneric), executed candidate by candidate with the functions above:
 *                       the checker of the batched exb_score_insertions
 *   orc_evaluate_partial evaluatePartialGenericSpecial.c:259 (evaluatePartialGeneric) ->
 *                       :190-255 (evaluatePartialCAT_FLEX) + :92-187 (computeVectorCAT_FLEX): one site,
 *                       one explicit rate, the whole traversal recomputed from the tips
 *   partition plumbing  newviewGenericSpecial.c:1114 (executeModel && width>0), :972-981
 *                       (branch slot), evaluateGenericSpecial.c:845-860 (masked partitions),
 *                       makenewzGenericSpecial.c:921-946 (derivative slots)
 *
 * Pinning: tests/test_oracle_golden.py checks this file against values dumped from the
 * reference's own AVX build (oracle/_ref/examl-AVX, built by oracle/Makefile from the
 * unmodified sources) on testData/49 and testData/140: first-evaluation and final lnL,
 * per-partition lnL, the first makenewz result, root scaler counts.  Fixtures and the
 * script that made them: tests/golden/.
 *
 * The API mirrors include/examl_b200.h one to one (orc_ instead of exb_) so the parity tests
 * drive both engines with identical arguments.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_ZMIN 1.0E-15                         /* axml.h:110 */
#define ORC_TWOTOTHE256 115792089237316195423570985008687907853269984665640564039457584007913129639936.0 /* axml.h:113 */
#define ORC_MINLIKELIHOOD (1.0 / ORC_TWOTOTHE256)  /* axml.h:118 */

typedef struct { int tipCase, pNumber, qNumber, rNumber; } orc_entry;

#define ORC_MAX_CATS 64
#define ORC_SCRATCH_NODES 64    /* EXB_SCRATCH_NODES; one more slot holds the candidate's new node */

typedef struct {
  int states, codes;
  int ncat;                     /* CAT: numberOfCategories */
  double catRates[ORC_MAX_CATS];/* CAT: perSiteRates */
  int *rateCategory;            /* CAT: per-site category index [width] */
  size_t width;
  int *wgt;
  unsigned char *tips;          /* [mxtips][width] */
  double *clv;                  /* [(mxtips-2)][width][4][S] */
  double *scratch;              /* [ORC_SCRATCH_NODES + 1][width][4][S]: nodes that exist only for orc_score_insertions (lazy) */
  double *sum;                  /* [width][4][S] */
  unsigned int *scaler;         /* [2*mxtips] (axml.c:1966-1971) */
  double EIGN[20], EV[400], EI[400], tipVector[23 * 20], gammaRates[4];
  /* LG4M / LG4X: one eigen-system per GAMMA category and category weights (axml.h:560-575) */
  int lg4;
  double EIGN4[4][20], EV4[4][400], EI4[4][400], tip4[4][23 * 20], weights[4];
} orc_part;

typedef struct {
  int mxtips, nparts, nbranches;
  int cat;                      /* 0: GAMMA (4 categories per site), 1: CAT/PSR (one category per site) */
  orc_part *parts;
} orc_engine;

/* rate categories a site carries / the rate table they index:
   GAMMA: R = 4, rates = gammaRates; CAT: R = 1, rates = perSiteRates[rateCategory[i]]
   (newviewGenericSpecial.c:958-968, discreteRateCategories axml.c:301) */
static int orc_R(const orc_engine *e) { return e->cat ? 1 : 4; }

int orc_create(orc_engine **out, int mxtips, int nparts, int nbranches)
{
  orc_engine *e = (orc_engine *)calloc(1, sizeof *e);
  e->mxtips = mxtips; e->nparts = nparts; e->nbranches = nbranches;
  e->parts = (orc_part *)calloc((size_t)nparts, sizeof(orc_part));
  *out = e;
  return 0;
}

void orc_destroy(orc_engine *e)
{
  int m;
  if (!e) return;
  for (m = 0; m < e->nparts; m++) {
    free(e->parts[m].wgt); free(e->parts[m].tips); free(e->parts[m].clv);
    free(e->parts[m].sum); free(e->parts[m].scaler); free(e->parts[m].rateCategory); free(e->parts[m].scratch);
  }
  free(e->parts); free(e);
}

int orc_set_partition(orc_engine *e, int model, int states, size_t width, const int *wgt, const unsigned char *y)
{
  orc_part *p = &e->parts[model];
  size_t span = (size_t)orc_R(e) * (size_t)states;
  p->states = states; p->codes = states == 20 ? 23 : 16; p->width = width;
  p->rateCategory = (int *)calloc(width + 1, sizeof(int));
  p->ncat = 1; p->catRates[0] = 1.0;
  p->wgt = (int *)malloc((width + 1) * sizeof(int));
  p->tips = (unsigned char *)malloc((size_t)e->mxtips * width + 1);
  memcpy(p->wgt, wgt, width * sizeof(int));
  memcpy(p->tips, y, (size_t)e->mxtips * width);
  p->clv = (double *)calloc((size_t)(e->mxtips - 2) * width * span + 1, sizeof(double));
  p->sum = (double *)calloc(width * span + 1, sizeof(double));
  p->scaler = (unsigned int *)calloc(2 * (size_t)e->mxtips + ORC_SCRATCH_NODES + 1, sizeof(unsigned int));
  return 0;
}

int orc_set_model(orc_engine *e, int model, const double *EIGN, const double *EV, const double *EI,
                  const double *tipVector, const double *gammaRates)
{
  orc_part *p = &e->parts[model];
  int S = p->states;
  memcpy(p->EIGN, EIGN, S * sizeof(double));
  memcpy(p->EV, EV, S * S * sizeof(double));
  memcpy(p->EI, EI, S * S * sizeof(double));
  memcpy(p->tipVector, tipVector, (size_t)p->codes * S * sizeof(double));
  memcpy(p->gammaRates, gammaRates, 4 * sizeof(double));
  return 0;
}

/* constants of category slot c (LG4: per category; otherwise shared) */
static const double *orc_EIGN(const orc_part *p, int c) { return p->lg4 ? p->EIGN4[c] : p->EIGN; }
static const double *orc_EI(const orc_part *p, int c) { return p->lg4 ? p->EI4[c] : p->EI; }
static const double *orc_EV(const orc_part *p, int c) { return p->lg4 ? p->EV4[c] : p->EV; }
static const double *orc_tipv(const orc_part *p, int c) { return p->lg4 ? p->tip4[c] : p->tipVector; }

/* Number of distinct P matrices of a partition and the rate of each (makeP's rptr / numberOfCategories,
   newviewGenericSpecial.c:958-968). */
static int orc_np(const orc_engine *e, const orc_part *p) { return e->cat ? p->ncat : 4; }
static const double *orc_rates(const orc_engine *e, const orc_part *p) { return e->cat ? p->catRates : p->gammaRates; }
/* index of the P matrix of (site i, category slot c) */
static int orc_pidx(const orc_engine *e, const orc_part *p, size_t i, int c) { return e->cat ? p->rateCategory[i] : c; }

/* P(z): P[n][j][k] = exp(EIGN[k] r_n lz) EI[j][k], P[n][j][0] = 1 for every rate r_n of the partition. */
static void orc_make_p(const orc_engine *e, double z, const orc_part *p, double *P)
{
  int S = p->states, n, j, k, np = orc_np(e, p);
  const double *rates = orc_rates(e, p);
  double lz = (z > ORC_ZMIN) ? log(z) : log(ORC_ZMIN);
  double d[20];
  for (n = 0; n < np; n++) {
    const double *EIGN = orc_EIGN(p, n), *EI = orc_EI(p, n);
    for (k = 1; k < S; k++) d[k] = p->lg4 ? exp(rates[n] * EIGN[k] * lz)        /* makeP_FlexLG4 newview:170-205 */
                                          : exp(rates[n] * (EIGN[k] * lz));
    for (j = 0; j < S; j++) {
      P[n * S * S + j * S] = 1.0;
      for (k = 1; k < S; k++) P[n * S * S + j * S + k] = d[k] * EI[j * S + k];
    }
  }
}

static double *clv_of(const orc_engine *e, const orc_part *p, int number)
{
  if (number > 2 * e->mxtips - 2)
    return p->scratch + (size_t)(number - (2 * e->mxtips - 1)) * p->width * (size_t)orc_R(e) * (size_t)p->states;
  return p->clv + (size_t)(number - e->mxtips - 1) * p->width * (size_t)orc_R(e) * (size_t)p->states;
}
static const unsigned char *tip_of(const orc_engine *e, const orc_part *p, int number)
{
  (void)e;
  return p->tips + (size_t)(number - 1) * p->width;
}

/* One traversal entry for one partition (GAMMA: newviewGAMMA_FLEX :430-682; CAT: newviewCAT_FLEX :217-427). */
static void orc_newview_entry(orc_engine *e, orc_part *p, const orc_entry *t, double qz, double rz)
{
  int S = p->states, R = orc_R(e), span = R * S, c, j, l, np = orc_np(e, p);
  size_t i;
  double *left = (double *)malloc(sizeof(double) * np * S * S), *right = (double *)malloc(sizeof(double) * np * S * S);
  double a[20], b[20], v1[20], v2[20];
  double *x3 = clv_of(e, p, t->pNumber);
  const double *x1 = NULL, *x2 = NULL;
  const unsigned char *tip1 = NULL, *tip2 = NULL;
  unsigned int addScale = 0;

  orc_make_p(e, qz, p, left);
  orc_make_p(e, rz, p, right);

  switch (t->tipCase) {
  case 0: tip1 = tip_of(e, p, t->qNumber); tip2 = tip_of(e, p, t->rNumber); break;
  case 1: tip1 = tip_of(e, p, t->qNumber); x2 = clv_of(e, p, t->rNumber); break;
  default: x1 = clv_of(e, p, t->qNumber); x2 = clv_of(e, p, t->rNumber); break;
  }

  for (i = 0; i < p->width; i++) {
    double *v = x3 + i * span;
    int scale = 1;
    for (c = 0; c < R; c++) {
      const double *L = left + (size_t)orc_pidx(e, p, i, c) * S * S, *Rm = right + (size_t)orc_pidx(e, p, i, c) * S * S;
      /* child vectors in the eigenvector basis; a tip's vector is the same for every category */
      if (tip1) memcpy(v1, orc_tipv(p, c) + (size_t)S * tip1[i], S * sizeof(double));
      else      memcpy(v1, x1 + i * span + c * S, S * sizeof(double));
      if (tip2) memcpy(v2, orc_tipv(p, c) + (size_t)S * tip2[i], S * sizeof(double));
      else      memcpy(v2, x2 + i * span + c * S, S * sizeof(double));
      for (l = 0; l < S; l++) {
        double al = 0.0, ar = 0.0;
        for (j = 0; j < S; j++) {
          al += v1[j] * L[l * S + j];
          ar += v2[j] * Rm[l * S + j];
        }
        a[l] = al; b[l] = ar;
      }
      for (j = 0; j < S; j++) v[c * S + j] = 0.0;
      for (l = 0; l < S; l++) {
        double x1px2 = a[l] * b[l];
        for (j = 0; j < S; j++) v[c * S + j] += x1px2 * orc_EV(p, c)[l * S + j];
      }
    }
    if (t->tipCase != 0) {           /* TIP_TIP is never rescaled (newviewGenericSpecial.c:496-523, 270-313) */
      for (l = 0; scale && l < span; l++) scale = fabs(v[l]) < ORC_MINLIKELIHOOD;
      if (scale) {
        for (l = 0; l < span; l++) v[l] *= ORC_TWOTOTHE256;
        addScale += (unsigned int)p->wgt[i];
      }
    }
  }
  p->scaler[t->pNumber] = p->scaler[t->qNumber] + p->scaler[t->rNumber] + addScale;
  free(left); free(right);
}

static int executes(const orc_engine *e, const unsigned char *mask, int m)
{
  return (!mask || mask[m]) && e->parts[m].width > 0;
}

static void orc_run_newviews(orc_engine *e, const orc_entry *ti, int first, int count,
                             const double *qz, const double *rz, const unsigned char *mask)
{
  int i, m, nb = e->nbranches;
  for (i = first; i < count; i++)
    for (m = 0; m < e->nparts; m++)
      if (executes(e, mask, m)) {
        int slot = nb > 1 ? m : 0;
        orc_newview_entry(e, &e->parts[m], &ti[i], qz[(size_t)i * nb + slot], rz[(size_t)i * nb + slot]);
      }
}

int orc_newview(orc_engine *e, const orc_entry *ti, int count, const double *qz, const double *rz, const unsigned char *mask)
{
  orc_run_newviews(e, ti, 0, count, qz, rz, mask);
  return 0;
}

static int is_tip(const orc_engine *e, int number) { return number <= e->mxtips; }

/* Resolve the two ends of branch ti[0]: if one end is a tip it becomes `tip` and the other x2
 * (evaluateGenericSpecial.c:612-650, makenewzGenericSpecial.c:71-150). */
static void branch_ends(const orc_engine *e, const orc_part *p, const orc_entry *t0,
                        const unsigned char **tipA, const unsigned char **tipB, const double **xa, const double **xb)
{
  int P = t0->pNumber, Q = t0->qNumber;
  *tipA = *tipB = NULL; *xa = *xb = NULL;
  if (is_tip(e, P) && is_tip(e, Q)) { *tipA = tip_of(e, p, P); *tipB = tip_of(e, p, Q); }
  else if (is_tip(e, Q)) { *tipA = tip_of(e, p, Q); *xb = clv_of(e, p, P); }
  else if (is_tip(e, P)) { *tipA = tip_of(e, p, P); *xb = clv_of(e, p, Q); }
  else { *xa = clv_of(e, p, P); *xb = clv_of(e, p, Q); }
}

/* GAMMA: evaluateGAMMA_FLEX eval:154-228 (log(0.25 |term|)); CAT: evaluateCAT_FLEX eval:232-330 (log |term|).
   Log likelihood of partition m at the branch (P, Q) with branch value z, scalers of both ends undone (eval:830). */
static double orc_branch_value(const orc_engine *e, int m, int P, int Q, double z)
{
  const orc_part *p = &e->parts[m];
  int c, k, n, R = orc_R(e);
  int S = p->states, span = R * S, np = orc_np(e, p);
  const double *rates = orc_rates(e, p);
  double diag[ORC_MAX_CATS * 20], lz, sum = 0.0;
  const unsigned char *tipA, *tipB; const double *xa, *xb;
  orc_entry t0;
  size_t i;
  t0.tipCase = -1; t0.pNumber = P; t0.qNumber = Q; t0.rNumber = 0;
  lz = (z < ORC_ZMIN) ? log(ORC_ZMIN) : log(z);
  for (n = 0; n < np; n++) {
    diag[n * S] = 1.0;
    for (k = 1; k < S; k++) diag[n * S + k] = p->lg4 ? exp(rates[n] * orc_EIGN(p, n)[k] * lz)   /* eval:122-145 */
                                                     : exp(rates[n] * (p->EIGN[k] * lz));
  }
  branch_ends(e, p, &t0, &tipA, &tipB, &xa, &xb);
  for (i = 0; i < p->width; i++) {
    double term = 0.0;
    for (c = 0; c < R; c++) {
      const double *dg = diag + (size_t)orc_pidx(e, p, i, c) * S;
      double t = 0.0;
      for (k = 0; k < S; k++) {
        double l = tipA ? orc_tipv(p, c)[(size_t)S * tipA[i] + k] : xa[i * span + c * S + k];
        double r = tipB ? orc_tipv(p, c)[(size_t)S * tipB[i] + k] : xb[i * span + c * S + k];
        t += l * r * dg[k];
      }
      term += p->lg4 ? p->weights[c] * t : t;          /* evaluateGTRGAMMAPROT_LG4 eval:1164-1287 */
    }
    sum += p->wgt[i] * ((e->cat || p->lg4) ? log(fabs(term)) : log(0.25 * fabs(term)));
  }
  sum += (double)(p->scaler[P] + p->scaler[Q]) * log(ORC_MINLIKELIHOOD);
  return sum;
}

int orc_evaluate(orc_engine *e, const orc_entry *ti, int count, const double *qz, const double *rz,
                 const unsigned char *mask, double *perPartitionLH)
{
  int m, nb = e->nbranches;
  orc_run_newviews(e, ti, 1, count, qz, rz, mask);
  for (m = 0; m < e->nparts; m++) {
    if (!executes(e, mask, m)) { if (e->parts[m].width == 0) perPartitionLH[m] = 0.0; continue; }
    perPartitionLH[m] = orc_branch_value(e, m, ti[0].pNumber, ti[0].qNumber, qz[nb > 1 ? m : 0]);
  }
  return 0;
}

/* The mixed descriptor of exb_score_insertions, executed the way the reference scores a lazy insertion: entries without the
   scoring bit are ordinary newviews (scratch nodes allowed); a scoring entry (tipCase | 4, pNumber = candidate) inserts a new
   node between qNumber and rNumber with branch values qz / rz -- insertBIG's non-thorough hook-ups, searchAlgo.c:519-535 --,
   computes its CLV (newviewGeneric(tr, p)) and evaluates the tree across the branch to the pruned subtree ti[0].pNumber with
   branch value ti[0].qz (evaluateGeneric; the reference evaluates at (p->next->next, r), which is the same likelihood by the
   pulley principle).  lnl[ncand][nparts]; masked partitions read 0. */
int orc_score_insertions(orc_engine *e, const orc_entry *ti, int count, const double *qz, const double *rz,
                         const unsigned char *mask, int ncand, double *lnl)
{
  const int nb = e->nbranches, newnode = 2 * e->mxtips - 1 + ORC_SCRATCH_NODES;
  int i, m;
  for (m = 0; m < e->nparts; m++) {
    orc_part *p = &e->parts[m];
    if (!p->scratch)
      p->scratch = (double *)calloc((size_t)(ORC_SCRATCH_NODES + 1) * p->width * (size_t)orc_R(e) * (size_t)p->states + 1, sizeof(double));
  }
  for (i = 0; i < ncand * e->nparts; i++) lnl[i] = 0.0;
  for (i = 1; i < count; i++) {
    if (!(ti[i].tipCase & 4)) { orc_run_newviews(e, ti, i, i + 1, qz, rz, mask); continue; }
    for (m = 0; m < e->nparts; m++) {
      orc_entry t = ti[i];
      int slot = nb > 1 ? m : 0;
      if (!executes(e, mask, m)) continue;
      t.tipCase &= 3; t.pNumber = newnode;
      orc_newview_entry(e, &e->parts[m], &t, qz[(size_t)i * nb + slot], rz[(size_t)i * nb + slot]);
      lnl[(size_t)ti[i].pNumber * e->nparts + m] = orc_branch_value(e, m, newnode, ti[0].pNumber, qz[slot]);
    }
  }
  return 0;
}

int orc_makenewz_prepare(orc_engine *e, const orc_entry *ti, int count, const double *qz, const double *rz,
                         const unsigned char *mask)
{
  int m, c, k, R = orc_R(e);
  orc_run_newviews(e, ti, 1, count, qz, rz, mask);
  for (m = 0; m < e->nparts; m++) {
    orc_part *p = &e->parts[m];
    int S = p->states, span = R * S;
    const unsigned char *tipA, *tipB; const double *xa, *xb;
    size_t i;
    if (!executes(e, mask, m)) continue;
    branch_ends(e, p, &ti[0], &tipA, &tipB, &xa, &xb);
    for (i = 0; i < p->width; i++)
      for (c = 0; c < R; c++)
        for (k = 0; k < S; k++) {
          double l = tipA ? orc_tipv(p, c)[(size_t)S * tipA[i] + k] : xa[i * span + c * S + k];
          double r = tipB ? orc_tipv(p, c)[(size_t)S * tipB[i] + k] : xb[i * span + c * S + k];
          p->sum[i * span + c * S + k] = l * r;
        }
  }
  return 0;
}

/* GAMMA: coreGTRGAMMA mkz:2309-2393; CAT: coreCAT_FLEX mkz:394-519 (the site's rate multiplies the weight,
   not the table: wr1 = r w, wr2 = r^2 w). */
int orc_core(orc_engine *e, const double *lzv, const unsigned char *mask, double *dlnLdlz, double *d2lnLdlz2)
{
  int m, c, k, n, nb = e->nbranches, R = orc_R(e);
  for (m = 0; m < nb; m++) { dlnLdlz[m] = 0.0; d2lnLdlz2[m] = 0.0; }
  for (m = 0; m < e->nparts; m++) {
    orc_part *p = &e->parts[m];
    int S = p->states, span = R * S, slot = nb > 1 ? m : 0, np = orc_np(e, p);
    const double *rates = orc_rates(e, p);
    double d0[ORC_MAX_CATS * 20], d1[ORC_MAX_CATS * 20], d2[ORC_MAX_CATS * 20], lz = lzv[slot], dl = 0.0, d2l = 0.0;
    size_t i;
    if (!executes(e, mask, m)) continue;
    for (n = 0; n < np; n++) {
      double ki = rates[n];
      d0[n * S] = 1.0; d1[n * S] = 0.0; d2[n * S] = 0.0;
      for (k = 1; k < S; k++) {
        if (e->cat) {
          d0[n * S + k] = exp((p->EIGN[k] * lz) * ki);
          d1[n * S + k] = p->EIGN[k];
          d2[n * S + k] = p->EIGN[k] * p->EIGN[k];
        } else {
          const double ek = orc_EIGN(p, n)[k];             /* LG4: coreGTRGAMMAPROT_LG4 mkz:2489-2578 */
          d0[n * S + k] = exp(ek * ki * lz);
          d1[n * S + k] = ek * ki;
          d2[n * S + k] = ek * ek * ki * ki;
        }
      }
    }
    for (i = 0; i < p->width; i++) {
      const double *s = p->sum + i * span;
      double L = 0.0, L1 = 0.0, L2 = 0.0, inv, w1 = p->wgt[i], w2 = p->wgt[i];
      for (c = 0; c < R; c++) {
        int pi = orc_pidx(e, p, i, c);
        double l0 = 0.0, l1 = 0.0, l2 = 0.0;
        for (k = 0; k < S; k++) {
          double t = d0[pi * S + k] * s[c * S + k];
          l0 += t; l1 += t * d1[pi * S + k]; l2 += t * d2[pi * S + k];
        }
        if (p->lg4) { L += p->weights[c] * l0; L1 += p->weights[c] * l1; L2 += p->weights[c] * l2; }
        else { L += l0; L1 += l1; L2 += l2; }
      }
      if (e->cat) { double r = rates[p->rateCategory[i]]; w1 = r * p->wgt[i]; w2 = r * r * p->wgt[i]; }
      inv = 1.0 / fabs(L);
      L1 *= inv; L2 *= inv;
      dl += w1 * L1;
      d2l += w2 * (L2 - L1 * L1);
    }
    dlnLdlz[slot] += dl;
    d2lnLdlz2[slot] += d2l;
  }
  return 0;
}

/* CAT/PSR: switch the engine to one rate category per site (before orc_set_partition) and hand over
   pInfo.numberOfCategories / perSiteRates / rateCategory (axml.h:533-629). */
int orc_set_rate_model(orc_engine *e, int rateHetModel, int maxCategories)
{
  (void)maxCategories;
  e->cat = rateHetModel == 0;          /* CAT = 0, GAMMA = 1 (axml.h:349-350) */
  return 0;
}
int orc_set_categories(orc_engine *e, int model, int ncat, const double *perSiteRates, const int *rateCategory)
{
  orc_part *p = &e->parts[model];
  if (ncat > ORC_MAX_CATS) return -1;
  p->ncat = ncat;
  memcpy(p->catRates, perSiteRates, (size_t)ncat * sizeof(double));
  memcpy(p->rateCategory, rateCategory, p->width * sizeof(int));
  return 0;
}

/* LG4M / LG4X: flattened pInfo.{EIGN,EV,EI,tipVector}_LG4[4] + gammaRates + weights. */
int orc_set_model_lg4(orc_engine *e, int model, const double *EIGN, const double *EV, const double *EI,
                      const double *tipVector, const double *gammaRates, const double *weights)
{
  orc_part *p = &e->parts[model];
  if (p->states != 20) return -1;
  p->lg4 = 1;
  memcpy(p->EIGN4, EIGN, sizeof p->EIGN4); memcpy(p->EV4, EV, sizeof p->EV4); memcpy(p->EI4, EI, sizeof p->EI4);
  memcpy(p->tip4, tipVector, sizeof p->tip4);
  memcpy(p->gammaRates, gammaRates, 4 * sizeof(double)); memcpy(p->weights, weights, 4 * sizeof(double));
  return 0;
}

/* evaluatePartialGeneric (evaluatePartialGenericSpecial.c:259), batched over n (partition, site, rate) triples.
 * Follows evaluatePartialCAT_FLEX (:190-255): entries[0] is the root branch, its pNumber a tip; entries
 * [1..count) are recomputed for this one site with rate ki by computeVectorCAT_FLEX (:92-187), x3 kept
 * per inner node; every entry's branch lengths are clamped to zmin before the logarithm (:220-221), the
 * root branch's is not (:232-234 overwrite the clamp); the result is w * (log|sum| + scale*log(minlikelihood)). */
int orc_evaluate_partial(orc_engine *e, const orc_entry *ti, int count, const double *qz, const double *rz,
                         size_t n, const int *models, const int *sites, const double *rates, double *out)
{
  size_t j;
  if (!e->cat || count < 1 || ti[0].pNumber > e->mxtips) return 1;
  for (j = 0; j < n; j++) {
    const orc_part *p = &e->parts[models[j]];
    const int S = p->states, slot = e->nbranches > 1 ? models[j] : 0;
    const size_t i = (size_t)sites[j];
    const double ki = rates[j];
    double *lVector = (double *)calloc((size_t)S * e->mxtips, sizeof(double)), d1[20], d2[20], x1px2[20], d[20], lz, term;
    const double *x1, *x2;
    int scale = 0, k, a, b;
    for (k = 1; k < count; k++) {
      const orc_entry *t = &ti[k];
      const double zq = qz[(size_t)k * e->nbranches + slot], zr = rz[(size_t)k * e->nbranches + slot];
      const double lz1 = ((zq > ORC_ZMIN) ? log(zq) : log(ORC_ZMIN)) * ki, lz2 = ((zr > ORC_ZMIN) ? log(zr) : log(ORC_ZMIN)) * ki;
      double *x3 = &lVector[S * (t->pNumber - e->mxtips)];
      int sc;
      x1 = t->tipCase == 2 ? &lVector[S * (t->qNumber - e->mxtips)] : &p->tipVector[S * tip_of(e, p, t->qNumber)[i]];
      x2 = t->tipCase == 0 ? &p->tipVector[S * tip_of(e, p, t->rNumber)[i]] : &lVector[S * (t->rNumber - e->mxtips)];
      d1[0] = x1[0]; d2[0] = x2[0];
      for (a = 1; a < S; a++) { d1[a] = x1[a] * exp(p->EIGN[a] * lz1); d2[a] = x2[a] * exp(p->EIGN[a] * lz2); }
      for (a = 0; a < S; a++) {
        double u1 = 0.0, u2 = 0.0;
        for (b = 0; b < S; b++) { u1 += d1[b] * p->EI[a * S + b]; u2 += d2[b] * p->EI[a * S + b]; }
        x1px2[a] = u1 * u2;
      }
      for (a = 0; a < S; a++) x3[a] = 0.0;
      for (a = 0; a < S; a++)
        for (b = 0; b < S; b++) x3[b] += x1px2[a] * p->EV[S * a + b];
      sc = 1;
      for (a = 0; sc && a < S; a++) sc = (x3[a] < ORC_MINLIKELIHOOD) && (x3[a] > -ORC_MINLIKELIHOOD);
      if (sc) { for (a = 0; a < S; a++) x3[a] *= ORC_TWOTOTHE256; scale++; }
    }
    x1 = &p->tipVector[S * tip_of(e, p, ti[0].pNumber)[i]];
    x2 = &lVector[S * (ti[0].qNumber - e->mxtips)];
    lz = log(qz[slot]) * ki;
    d[0] = 1.0;
    for (a = 1; a < S; a++) d[a] = exp(p->EIGN[a] * lz);
    term = 0.0;
    for (a = 0; a < S; a++) term += x1[a] * x2[a] * d[a];
    out[j] = (log(fabs(term)) + scale * log(ORC_MINLIKELIHOOD)) * p->wgt[i];
    free(lVector);
  }
  return 0;
}

int orc_read_clv(orc_engine *e, int model, int number, double *out)
{
  orc_part *p = &e->parts[model];
  memcpy(out, clv_of(e, p, number), p->width * (size_t)orc_R(e) * (size_t)p->states * sizeof(double));
  return 0;
}
int orc_read_scaler(orc_engine *e, int model, int number, unsigned int *out)
{
  *out = e->parts[model].scaler[number];
  return 0;
}
int orc_read_sumtable(orc_engine *e, int model, double *out)
{
  orc_part *p = &e->parts[model];
  memcpy(out, p->sum, p->width * (size_t)orc_R(e) * (size_t)p->states * sizeof(double));
  return 0;
}
