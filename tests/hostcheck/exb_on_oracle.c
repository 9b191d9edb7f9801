/* exb_on_oracle.c -- TEST INFRASTRUCTURE (tests/hostcheck), never part of the product.
 *
 * Implements the C ABI of include/examl_b200.h on top of the CPU oracle (oracle/likelihood_oracle.c) so that the
 * HOST side of the drop-in -- examl_b200/csrc/axml_backend.c: traversal descriptors, x-flag handling, the
 * Newton-Raphson state machine, executeModel masks, the anticipation of the PSR rate scan, the multi-rank
 * plumbing over mpi_tcp.c -- can be exercised by the `-m "not gpu"` tests on a machine without a GPU, linked
 * with the reference's unchanged host objects exactly like bin/examl-b200 is.  The product binary links
 * libexaml_b200.so instead and has no CPU path.
 *
 * Differences from the CUDA engine that the back-end can observe: exb_makenewz (device-resident Newton-Raphson)
 * is not provided -- the harness is run with EXAML_B200_HOST_NR=1, which is the multi-rank form of the loop --
 * and the all-reduce of exb_evaluate / exb_core goes through MPI_Allreduce of the shim instead of NCCL.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mpi.h>
#include "examl_b200.h"

typedef struct orc_engine orc_engine;
int orc_create(orc_engine **out, int mxtips, int nparts, int nbranches);
void orc_destroy(orc_engine *e);
int orc_set_partition(orc_engine *e, int model, int states, size_t width, const int *wgt, const unsigned char *y);
int orc_set_model(orc_engine *e, int model, const double *EIGN, const double *EV, const double *EI, const double *tipVector, const double *gammaRates);
int orc_set_model_lg4(orc_engine *e, int model, const double *EIGN, const double *EV, const double *EI, const double *tipVector, const double *gammaRates, const double *weights);
int orc_set_rate_model(orc_engine *e, int rateHetModel, int maxCategories);
int orc_set_categories(orc_engine *e, int model, int ncat, const double *perSiteRates, const int *rateCategory);
int orc_newview(orc_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz, const unsigned char *mask);
int orc_evaluate(orc_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz, const unsigned char *mask, double *lh);
int orc_makenewz_prepare(orc_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz, const unsigned char *mask);
int orc_core(orc_engine *e, const double *lz, const unsigned char *mask, double *d1, double *d2);
int orc_evaluate_partial(orc_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz, size_t n,
                         const int *models, const int *sites, const double *rates, double *out);
int orc_score_insertions(orc_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz,
                         const unsigned char *mask, int ncand, double *lnl);

struct exb_engine { orc_engine *o; int P, nb, nranks, cat; long calls; };
static char g_err[256] = "";

const char *exb_last_error(void) { return g_err; }

int exb_create(exb_engine **out, int device, int mxtips, int numPartitions, int numBranches)
{
  exb_engine *e = (exb_engine *)calloc(1, sizeof *e);
  (void)device;
  e->P = numPartitions; e->nb = numBranches; e->nranks = 1;
  orc_create(&e->o, mxtips, numPartitions, numBranches);
  *out = e;
  return EXB_OK;
}
void exb_destroy(exb_engine *e) { if (e) { orc_destroy(e->o); free(e); } }
int exb_set_rate_model(exb_engine *e, int rateHetModel, int maxCategories) { e->cat = rateHetModel == EXB_CAT; return orc_set_rate_model(e->o, rateHetModel, maxCategories); }
int exb_set_partition(exb_engine *e, int model, int states, size_t width, const int *wgt, const unsigned char *y)
{ return orc_set_partition(e->o, model, states, width, wgt, y); }
int exb_commit(exb_engine *e) { (void)e; return EXB_OK; }
int exb_set_timing(exb_engine *e, int on) { (void)e; (void)on; return EXB_OK; }
int exb_set_model(exb_engine *e, int model, const double *EIGN, const double *EV, const double *EI, const double *tipVector, const double *gammaRates)
{ return orc_set_model(e->o, model, EIGN, EV, EI, tipVector, gammaRates); }
int exb_set_model_lg4(exb_engine *e, int model, const double *EIGN, const double *EV, const double *EI, const double *tipVector,
                      const double *gammaRates, const double *weights)
{ return orc_set_model_lg4(e->o, model, EIGN, EV, EI, tipVector, gammaRates, weights); }
int exb_set_categories(exb_engine *e, int model, int n, const double *rates, const int *cat) { return orc_set_categories(e->o, model, n, rates, cat); }

int exb_comm_unique_id(void *id_out) { memset(id_out, 0, EXB_UNIQUE_ID_BYTES); return EXB_OK; }
int exb_comm_init(exb_engine *e, int rank, int nranks, const void *id) { (void)rank; (void)id; e->nranks = nranks; return EXB_OK; }
int exb_comm_is_p2p(exb_engine *e) { (void)e; return 0; }
int exb_set_sev(exb_engine *e, int on) { (void)e; (void)on; return EXB_OK; }     /* the CPU checker keeps dense CLVs */

int exb_newview(exb_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz, const unsigned char *mask)
{ e->calls++; return orc_newview(e->o, ti, count, qz, rz, mask); }

/* sum over ranks of the executed partitions' values (evaluateGenericSpecial.c:968-975 / makenewzGenericSpecial.c:1241-1248) */
static void allreduce(exb_engine *e, double *v, int n)
{
  if (e->nranks > 1) MPI_Allreduce(MPI_IN_PLACE, v, n, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
}

int exb_evaluate(exb_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz, const unsigned char *mask, double *lh)
{
  double *tmp = (double *)calloc((size_t)e->P, sizeof(double));
  int m;
  e->calls++;
  orc_evaluate(e->o, ti, count, qz, rz, mask, tmp);
  for (m = 0; m < e->P; m++) if (mask && !mask[m]) tmp[m] = 0.0;
  allreduce(e, tmp, e->P);
  for (m = 0; m < e->P; m++) if (!mask || mask[m]) lh[m] = tmp[m];     /* masked partitions keep their value */
  free(tmp);
  return EXB_OK;
}
int exb_makenewz_prepare(exb_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz, const unsigned char *mask)
{ e->calls++; return orc_makenewz_prepare(e->o, ti, count, qz, rz, mask); }
int exb_core(exb_engine *e, const double *lz, const unsigned char *mask, double *d1, double *d2)
{
  double both[2 * 256];
  int b;
  e->calls++;
  orc_core(e->o, lz, mask, d1, d2);
  for (b = 0; b < e->nb; b++) { both[b] = d1[b]; both[e->nb + b] = d2[b]; }
  allreduce(e, both, 2 * e->nb);
  for (b = 0; b < e->nb; b++) { d1[b] = both[b]; d2[b] = both[e->nb + b]; }
  return EXB_OK;
}
int exb_makenewz(exb_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz, const unsigned char *mask,
                 const double *z0, int maxiter, double *result, int *curvatOK, int *iterationsDone)
{
  (void)e; (void)ti; (void)count; (void)qz; (void)rz; (void)mask; (void)z0; (void)maxiter; (void)result; (void)curvatOK; (void)iterationsDone;
  snprintf(g_err, sizeof g_err, "hostcheck harness: run with EXAML_B200_HOST_NR=1 (no device-resident Newton-Raphson on the oracle)");
  return EXB_ERR_STATE;
}
int exb_evaluate_partial(exb_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz, size_t n,
                         const int *models, const int *sites, const double *rates, double *out)
{
  e->calls++;
  if (orc_evaluate_partial(e->o, ti, count, qz, rz, n, models, sites, rates, out) != 0) {
    snprintf(g_err, sizeof g_err, "orc_evaluate_partial refused the call");
    return EXB_ERR_ARG;
  }
  return EXB_OK;
}

/* batched lazy-SPR scoring: the oracle executes the mixed descriptor candidate by candidate; sum over ranks like exb_evaluate */
int exb_can_score_insertions(exb_engine *e) { return !e->cat; }
int exb_max_entries(exb_engine *e) { (void)e; return 1 << 20; }
int exb_score_insertions(exb_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz,
                         const unsigned char *mask, int ncand, double *lnl)
{
  e->calls++;
  if (e->cat) { snprintf(g_err, sizeof g_err, "hostcheck harness: no batched scoring under PSR (like the CUDA engine)"); return EXB_ERR_STATE; }
  orc_score_insertions(e->o, ti, count, qz, rz, mask, ncand, lnl);
  if (ncand > 0) allreduce(e, lnl, ncand * e->P);
  return EXB_OK;
}
