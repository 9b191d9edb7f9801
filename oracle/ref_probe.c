/* ref_probe.c -- TEST INFRASTRUCTURE.  Linked into the UNMODIFIED reference build
 * (oracle/Makefile, outputs in oracle/_ref/) with
 *      -Wl,--wrap=evaluateGeneric -Wl,--wrap=makenewzGeneric -Wl,--wrap=evaluatePartialGeneric
 * so that the reference's own binary can (a) dump golden states for the parity tests
 * and (b) time its CPU likelihood path for bench.py's reference arm -- without editing or
 * copying a single reference source line.
 *
 * Hooks (all driven by environment variables, inert otherwise):
 *   EXAML_PROBE_DUMP=<prefix>   write <prefix>.first.json after the first evaluateGeneric
 *                               (axml.c:2317 / :2760), <prefix>.mkz.json after the first
 *                               makenewzGeneric, <prefix>.final.json at process exit
 *                               (after one more full evaluateGeneric on the final tree);
 *                               under -m PSR also <prefix>.partial.json: the state during the first
 *                               optimizeRateCategories (optimizeModel.c:2403) and the first
 *                               PARTIAL_REC evaluatePartialGeneric calls of every partition
 *                               (partition, site, rate, returned value).
 *   EXAML_PROBE_BENCH=<K>       after the first evaluateGeneric: W warm-up (EXAML_PROBE_WARMUP,
 *                               default 1) then K timed full-traversal evaluateGeneric calls;
 *                               print one JSON line "PROBE_BENCH {...}" and exit(0).
 *   EXAML_PROBE_NOTIPS=1        omit the alignment bytes / weights from <prefix>.first.json too
 *                               (the mkz / final dumps never carry them).
 *   EXAML_PROBE_EXIT_FIRST=1    exit(0) right after the first evaluateGeneric (and its dump).
 *   EXAML_PROBE_TRACE=<file>    one line per evaluateGeneric call -- "E <p> <p->back> <fullTraversal> <tr->likelihood %.17g>" --
 *                               up to EXAML_PROBE_TRACE_MAX lines (default 5000): the sequence of tree evaluations of a
 *                               search, among them every lazy insertion testInsertBIG scores (searchAlgo.c:735).  Pins the
 *                               batched insertion scan to the reference's own numbers (tests/golden/*.trace.gz).
 *   EXAML_PROBE_TRACE_MKZ=1|2   (debugging aid) also one line per makenewzGeneric call -- "M <p> <q> <mask> <result...>" --;
 *                               2: plus maxiter, the start values, partitionConverged and the derivatives of the branch's
 *                               sum table at the start values as the binary's own execCore computes them.  (This is how the
 *                               -M search on testData/49 was found to leave the reference's path at a branch whose fourth
 *                               partition has derivatives of 1e-20: rounding noise, different in every implementation.)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <math.h>
#include "axml.h"

void __real_evaluateGeneric(tree *tr, nodeptr p, boolean fullTraversal);
void __real_makenewzGeneric(tree *tr, nodeptr p, nodeptr q, double *z0, int maxiter, double *result, boolean mask);
double __real_evaluatePartialGeneric(tree *tr, int i, double ki, int _model);
#ifdef _USE_OMP
void allocateXVectors(tree *tr);
#endif

extern int processID;            /* axml.c: the MPI rank */
static tree *g_tr = NULL;

/* <prefix>.<kind>.json on rank 0, <prefix>.r<rank>.<kind>.json elsewhere (every rank dumps its own slice) */
static void dump_path(char *path, size_t n, const char *prefix, const char *kind)
{
  if (processID == 0) snprintf(path, n, "%s.%s.json", prefix, kind);
  else snprintf(path, n, "%s.r%d.%s.json", prefix, processID, kind);
}
static int g_eval_calls = 0, g_mkz_calls = 0;

/* EXAML_PROBE_TRACE: one line per traced call, up to EXAML_PROBE_TRACE_MAX evaluateGeneric lines */
static FILE *trace_file(int count_it)
{
  static FILE *trace = NULL;
  static long left = -1;
  const char *tf = getenv("EXAML_PROBE_TRACE");
  if (!tf || processID != 0 || left == 0) return NULL;
  if (!trace) { trace = fopen(tf, "w"); left = getenv("EXAML_PROBE_TRACE_MAX") ? atol(getenv("EXAML_PROBE_TRACE_MAX")) : 5000; }
  if (trace && count_it && --left == 0) { FILE *f = trace; trace = NULL; return f; }    /* the caller writes the last line and closes */
  return trace;
}

static double now_s(void)
{
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return tv.tv_sec + 1e-6 * tv.tv_usec;
}

static void put_darr(FILE *f, const char *name, const double *v, size_t n)
{
  size_t i;
  fprintf(f, "\"%s\": [", name);
  for (i = 0; i < n; i++) fprintf(f, "%s%.17g", i ? ", " : "", v[i]);
  fprintf(f, "]");
}

static int tip_codes(int dataType) { return dataType == AA_DATA ? 23 : (dataType == DNA_DATA ? 16 : 3); }

static void dump_state(tree *tr, const char *path, int with_alignment, const char *extra)
{
  FILE *f = fopen(path, "w");
  int m, i, b, nb = tr->numBranches;
  int notips = !with_alignment || getenv("EXAML_PROBE_NOTIPS") != NULL, first_entry = 1;
  if (!f) { perror(path); return; }
  fprintf(f, "{\n\"mxtips\": %d, \"numBranches\": %d, \"NumberOfModels\": %d, \"rateHetModel\": %d, \"maxCategories\": %d,\n",
          tr->mxtips, nb, tr->NumberOfModels, tr->rateHetModel, tr->maxCategories);
  fprintf(f, "\"likelihood\": %.17g,\n", tr->likelihood);
  put_darr(f, "perPartitionLH", tr->perPartitionLH, (size_t)tr->NumberOfModels);
  fprintf(f, ",\n");
  if (extra) fprintf(f, "%s,\n", extra);
  if (!notips && tr->nameList) {
    fprintf(f, "\"names\": [");
    for (i = 1; i <= tr->mxtips; i++) fprintf(f, "%s\"%s\"", i > 1 ? ", " : "", tr->nameList[i]);
    fprintf(f, "],\n");
  }
  fprintf(f, "\"partitions\": [\n");
  for (m = 0; m < tr->NumberOfModels; m++) {
    pInfo *p = &tr->partitionData[m];
    int S = p->states, codes = tip_codes(p->dataType);
    size_t w;
    fprintf(f, "{\"states\": %d, \"dataType\": %d, \"protModels\": %d, \"width\": %zu, \"alpha\": %.17g,\n",
            S, p->dataType, p->protModels, p->width, p->alpha);
    put_darr(f, "EIGN", p->EIGN, (size_t)S); fprintf(f, ",\n");
    put_darr(f, "EV", p->EV, (size_t)S * S); fprintf(f, ",\n");
    put_darr(f, "EI", p->EI, (size_t)S * S); fprintf(f, ",\n");
    put_darr(f, "tipVector", p->tipVector, (size_t)codes * S); fprintf(f, ",\n");
    put_darr(f, "frequencies", p->frequencies, (size_t)S); fprintf(f, ",\n");
    put_darr(f, "gammaRates", p->gammaRates, 4); fprintf(f, ",\n");
    if (p->protModels == LG4M || p->protModels == LG4X) {
      int c;
      const char *names[4] = {"EIGN_LG4", "EV_LG4", "EI_LG4", "tipVector_LG4"};
      double **arrs[4] = {p->EIGN_LG4, p->EV_LG4, p->EI_LG4, p->tipVector_LG4};
      size_t lens[4] = {20, 400, 400, 460};
      int a;
      for (a = 0; a < 4; a++) {
        fprintf(f, "\"%s\": [", names[a]);
        for (c = 0; c < 4; c++) {
          size_t q;
          fprintf(f, "%s[", c ? "," : "");
          for (q = 0; q < lens[a]; q++) fprintf(f, "%s%.17g", q ? "," : "", arrs[a][c][q]);
          fprintf(f, "]");
        }
        fprintf(f, "],\n");
      }
      put_darr(f, "weights", p->weights, 4); fprintf(f, ",\n");
    }
    if (tr->rateHetModel == CAT) {
      fprintf(f, "\"numberOfCategories\": %d,\n", p->numberOfCategories);
      put_darr(f, "perSiteRates", p->perSiteRates, (size_t)p->numberOfCategories); fprintf(f, ",\n");
      fprintf(f, "\"rateCategory\": [");
      for (w = 0; w < p->width; w++) fprintf(f, "%s%d", w ? "," : "", p->rateCategory[w]);
      fprintf(f, "],\n");
    }
    fprintf(f, "\"rootScaler\": %u", p->width ?
            p->globalScaler[tr->td[0].ti[0].pNumber] + p->globalScaler[tr->td[0].ti[0].qNumber] : 0u);
    if (!notips) {
      fprintf(f, ",\n\"wgt\": [");
      for (w = 0; w < p->width; w++) fprintf(f, "%s%d", w ? "," : "", p->wgt[w]);
      fprintf(f, "],\n\"tips\": [");
      for (i = 1; i <= tr->mxtips; i++) {
        fprintf(f, "%s\"", i > 1 ? ",\n" : "");
        for (w = 0; w < p->width; w++) fprintf(f, "%02x", p->yVector[i][w]);
        fprintf(f, "\"");
      }
      fprintf(f, "]");
    }
    fprintf(f, "}%s\n", m + 1 < tr->NumberOfModels ? "," : "");
  }
  fprintf(f, "],\n\"traversal\": {\"count\": %d, \"first\": %d, \"entries\": [\n", tr->td[0].count, first_entry);
  for (i = 0; i < tr->td[0].count; i++) {
    traversalInfo *t = &tr->td[0].ti[i];
    fprintf(f, "{\"tipCase\": %d, \"p\": %d, \"q\": %d, \"r\": %d, \"qz\": [", i ? t->tipCase : -1, t->pNumber, t->qNumber, i ? t->rNumber : 0);
    for (b = 0; b < nb; b++) fprintf(f, "%s%.17g", b ? "," : "", t->qz[b]);
    fprintf(f, "], \"rz\": [");
    for (b = 0; b < nb; b++) fprintf(f, "%s%.17g", b ? "," : "", i ? t->rz[b] : 0.0);
    fprintf(f, "]}%s\n", i + 1 < tr->td[0].count ? "," : "");
  }
  fprintf(f, "]}\n}\n");
  fclose(f);
}

/* ---- evaluatePartialGeneric: record the first calls of every partition, dump before the next evaluateGeneric ---- */
#define PARTIAL_REC 160
#define PARTIAL_MODELS 8
static struct { int model, site; double rate, value; } g_partial[PARTIAL_REC * PARTIAL_MODELS];
static int g_partial_n = 0, g_partial_per_model[PARTIAL_MODELS], g_partial_dumped = 0;

double __wrap_evaluatePartialGeneric(tree *tr, int i, double ki, int _model)
{
  const double v = __real_evaluatePartialGeneric(tr, i, ki, _model);
#ifndef _USE_OMP
  if (!g_partial_dumped && getenv("EXAML_PROBE_DUMP") && _model < PARTIAL_MODELS && g_partial_per_model[_model] < PARTIAL_REC) {
    g_partial_per_model[_model]++;
    g_partial[g_partial_n].model = _model; g_partial[g_partial_n].site = i;
    g_partial[g_partial_n].rate = ki;      g_partial[g_partial_n].value = v;
    g_partial_n++;
  }
#endif
  return v;
}

static void dump_partial(tree *tr)
{
  const char *prefix = getenv("EXAML_PROBE_DUMP");
  char path[4096], *extra = (char *)malloc((size_t)g_partial_n * 96 + 64);
  int off = 0, j;
  off += sprintf(extra + off, "\"partial\": [");
  for (j = 0; j < g_partial_n; j++)
    off += sprintf(extra + off, "%s[%d,%d,%.17g,%.17g]", j ? "," : "", g_partial[j].model, g_partial[j].site, g_partial[j].rate, g_partial[j].value);
  off += sprintf(extra + off, "]");
  dump_path(path, sizeof path, prefix, "partial");
  dump_state(tr, path, 0, extra);      /* td[0] is still the full traversal the calls were made with */
  free(extra);
  g_partial_dumped = 1;
}

static void dump_final(void)
{
  const char *prefix = getenv("EXAML_PROBE_DUMP");
  char path[4096];
  if (!prefix || !g_tr) return;
  __real_evaluateGeneric(g_tr, g_tr->start, TRUE);
  dump_path(path, sizeof path, prefix, "final");
  dump_state(g_tr, path, 0, NULL);
}

void __wrap_evaluateGeneric(tree *tr, nodeptr p, boolean fullTraversal)
{
  const char *prefix = getenv("EXAML_PROBE_DUMP"), *bench = getenv("EXAML_PROBE_BENCH");
#ifdef _USE_OMP
  if (g_eval_calls == 0 && tr->partitionData[0].xVector[0] == NULL)
    allocateXVectors(tr);     /* the reference's -f e path forgets this under OMP (only axml.c:2752 calls it) */
#endif
  if (g_partial_n > 0 && !g_partial_dumped && prefix) dump_partial(tr);
  __real_evaluateGeneric(tr, p, fullTraversal);
  g_eval_calls++;
  {
    FILE *trace = trace_file(1);
    if (trace) {
      fprintf(trace, "E %d %d %d %.17g\n", p->number, p->back->number, fullTraversal ? 1 : 0, tr->likelihood);
      fflush(trace);
    }
  }
  if (g_eval_calls == 1) {
    g_tr = tr;
    if (prefix) {
      char path[4096];
      dump_path(path, sizeof path, prefix, "first");
      dump_state(tr, path, 1, NULL);
      atexit(dump_final);
    }
    if (getenv("EXAML_PROBE_EXIT_FIRST")) { g_tr = NULL; fflush(stdout); exit(0); }
    if (bench) {
      int K = atoi(bench), W = getenv("EXAML_PROBE_WARMUP") ? atoi(getenv("EXAML_PROBE_WARMUP")) : 1, i, m;
      size_t sites = 0;
      double t0, t1;
      for (m = 0; m < tr->NumberOfModels; m++) sites += tr->partitionData[m].width;
      for (i = 0; i < W; i++) __real_evaluateGeneric(tr, p, TRUE);
      t0 = now_s();
      for (i = 0; i < K; i++) __real_evaluateGeneric(tr, p, TRUE);
      t1 = now_s();
      printf("PROBE_BENCH {\"steps\": %d, \"seconds\": %.9g, \"sites\": %zu, \"inner_nodes\": %d, "
             "\"site_updates_per_s\": %.9g, \"likelihood\": %.17g}\n",
             K, t1 - t0, sites, tr->mxtips - 2,
             (double)K * (double)(tr->mxtips - 2) * (double)sites / (t1 - t0), tr->likelihood);
      fflush(stdout);
      g_tr = NULL;   /* no final dump */
      exit(0);
    }
  }
}

void __wrap_makenewzGeneric(tree *tr, nodeptr p, nodeptr q, double *z0, int maxiter, double *result, boolean mask)
{
  const char *prefix = getenv("EXAML_PROBE_DUMP");
  double z0copy[NUM_BRANCHES];
  int b;
  for (b = 0; b < tr->numBranches; b++) z0copy[b] = z0[b];
  __real_makenewzGeneric(tr, p, q, z0, maxiter, result, mask);
  g_mkz_calls++;
  if (getenv("EXAML_PROBE_TRACE_MKZ")) {           /* also the branch optimisations: "M <p> <q> <mask> <result...>" (debugging aid) */
    FILE *trace = trace_file(0);
    if (trace) {
      fprintf(trace, "M %d %d %d", p->number, q->number, mask ? 1 : 0);
      for (b = 0; b < tr->numBranches; b++) fprintf(trace, " %.17g", result[b]);
      if (getenv("EXAML_PROBE_TRACE_MKZ")[0] == '2') {         /* ... and the inputs */
        fprintf(trace, " | maxiter %d z0", maxiter);
        for (b = 0; b < tr->numBranches; b++) fprintf(trace, " %.17g", z0copy[b]);
        fprintf(trace, " conv");
        for (b = 0; b < tr->numBranches; b++) fprintf(trace, " %d", tr->partitionConverged[b] ? 1 : 0);
        {   /* derivatives of the branch's sum table at the start values, by the binary's own execCore */
          volatile double d1[NUM_BRANCHES], d2[NUM_BRANCHES];
          double lz[NUM_BRANCHES];
          for (b = 0; b < tr->numBranches; b++) lz[b] = log(z0copy[b] < zmin ? zmin : z0copy[b]);
          storeExecuteMaskInTraversalDescriptor(tr);
          storeValuesInTraversalDescriptor(tr, lz);
          execCore(tr, d1, d2);
          fprintf(trace, " d1");
          for (b = 0; b < tr->numBranches; b++) fprintf(trace, " %.12g", d1[b]);
          fprintf(trace, " d2");
          for (b = 0; b < tr->numBranches; b++) fprintf(trace, " %.12g", d2[b]);
        }
      }
      fprintf(trace, "\n");
    }
  }
  if (g_mkz_calls == 1 && prefix) {
    char path[4096], extra[16384];
    int off = 0;
    off += snprintf(extra + off, sizeof extra - off, "\"makenewz\": {\"p\": %d, \"q\": %d, \"maxiter\": %d, \"z0\": [", p->number, q->number, maxiter);
    for (b = 0; b < tr->numBranches; b++) off += snprintf(extra + off, sizeof extra - off, "%s%.17g", b ? "," : "", z0copy[b]);
    off += snprintf(extra + off, sizeof extra - off, "], \"result\": [");
    for (b = 0; b < tr->numBranches; b++) off += snprintf(extra + off, sizeof extra - off, "%s%.17g", b ? "," : "", result[b]);
    off += snprintf(extra + off, sizeof extra - off, "]}");
    dump_path(path, sizeof path, prefix, "mkz");
    dump_state(tr, path, 0, extra);
  }
}
