/* TEST INFRASTRUCTURE -- the operator API of the reference restated on plain arrays:
 *   ReactorCvode::react(box) CPU path  ReactorCvode.cpp:1372-1453 (mask == -1 => skip, FC = 0)
 *   ReactorRK64::react(box)            ReactorRK64.cpp:167-327 (mask ignored)
 *   box_flatten / box_unflatten        ReactorUtils.H:76-207 (incl. clean_init_massfrac)
 * Optional pthread fan-out over cells is only a convenience for the CPU-baseline timing (cells are
 * independent; the reference itself is serial per rank). */
#include "oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

void orc_default_opts(orc_opts* o)
{
  memset(o, 0, sizeof(*o));
  o->rtol = 1e-6;  /* ReactorCvode.H:152-153 */
  o->atol = 1e-10;
  o->max_order = 2;
  o->max_nls_iters = 3;
  o->clean_init_massfrac = 0;
  o->rk64_guess = 10;
  o->rk64_min = 5;
  o->rk64_max = 500;
  o->analytical_jacobian = 1;
  o->reactor_type = 2;
}

static double dmin(double a, double b) { return a < b ? a : b; }
static double dmax(double a, double b) { return a > b ? a : b; }

/* one cell given gathered inputs; returns FC */
static double react_cell(const plm_mech_tables* m, const orc_opts* o, int use_rk64, double* rY /*ns in/out*/,
                         const double* rYsrc, double* T, double* rE, double rEsrc, double dt, double time,
                         orc_stats* st)
{
  const int ns = m->ns;
  double y[ORC_MAXN], massfrac[ORC_MAXS];
  double rho = 0.0, rho_old = 0.0;
  const int clean = o->clean_init_massfrac && !use_rk64; /* RK64 ignores it (ReactorRK64.cpp:176) */
  if (clean)
    for (int k = 0; k < ns; k++) rho_old += rY[k];
  for (int k = 0; k < ns; k++) {
    const double v = clean ? dmin(dmax(0.0, rY[k]), rho_old) : rY[k];
    y[k] = v;
    rho += v;
    massfrac[k] = v;
  }
  double rho_inv = 1.0 / rho;
  for (int k = 0; k < ns; k++) massfrac[k] *= rho_inv;
  double temp = *T;
  const double rhoe_init = *rE;
  if (o->reactor_type == 1) orc_get_T_given_eY(m, rhoe_init * rho_inv, massfrac, &temp); /* REY2T, ReactorUtils.H:118-126 */
  else orc_get_T_given_hY(m, rhoe_init * rho_inv, massfrac, &temp);
  y[ns] = temp;
  double fc;
  if (use_rk64) {
    fc = (double)orc_rk64_cell(m, o, y, rhoe_init, rEsrc, rYsrc, time, dt);
    if (st) st->nst = (long)fc;
  } else {
    orc_stats s;
    orc_cvode_cell(m, o, y, rhoe_init, rEsrc, rYsrc, time, time + dt, &s);
    fc = (double)s.nfe;
    if (st) *st = s;
  }
  /* unflatten */
  double rhoY_oldclipped[ORC_MAXS], rho_oldclipped = 0.0;
  if (clean)
    for (int k = 0; k < ns; k++) {
      rhoY_oldclipped[k] = dmin(dmax(0.0, rY[k]), rho_old);
      rho_oldclipped += rhoY_oldclipped[k];
    }
  rho = 0.0;
  for (int k = 0; k < ns; k++) {
    const double v = clean ? rY[k] + (y[k] - rhoY_oldclipped[k]) * rho_old / rho_oldclipped : y[k];
    rY[k] = v;
    rho += v;
    massfrac[k] = v;
  }
  rho_inv = 1.0 / rho;
  for (int k = 0; k < ns; k++) massfrac[k] *= rho_inv;
  temp = y[ns];
  *rE = rhoe_init + dt * rEsrc;
  if (o->reactor_type == 1) orc_get_T_given_eY(m, (*rE) * rho_inv, massfrac, &temp);
  else orc_get_T_given_hY(m, (*rE) * rho_inv, massfrac, &temp);
  *T = temp;
  return fc;
}

static void acc(orc_stats* sum, const orc_stats* s)
{
  sum->nst += s->nst; sum->nfe += s->nfe; sum->nje += s->nje; sum->nsetups += s->nsetups;
  sum->nni += s->nni; sum->ncfn += s->ncfn; sum->netf += s->netf;
  if (s->flag < 0 && sum->flag == 0) sum->flag = s->flag;
}

#define A4(a, i, j, k, n) \
  (a).p[((i) - (a).lo[0]) + ((j) - (a).lo[1]) * (a).jstride + ((k) - (a).lo[2]) * (a).kstride + (long)(n) * (a).nstride]

int orc_react_box(const plm_mech_tables* m, const orc_opts* o, int use_rk64, const int lo[3], const int hi[3],
                  orc_a4 rY, orc_a4 rYsrc, orc_a4 T, orc_a4 rE, orc_a4 rEsrc, orc_a4 FC, const int* mask,
                  long mjs, long mks, const int mlo[3], double dt, double time, orc_stats* sum)
{
  const int ns = m->ns;
  orc_stats total;
  memset(&total, 0, sizeof(total));
  for (int k = lo[2]; k <= hi[2]; k++)
    for (int j = lo[1]; j <= hi[1]; j++)
      for (int i = lo[0]; i <= hi[0]; i++) {
        const int mk = mask ? mask[(i - mlo[0]) + (j - mlo[1]) * mjs + (k - mlo[2]) * mks] : 1;
        if (!use_rk64 && mk == -1) {
          A4(FC, i, j, k, 0) = 0.0;
          continue;
        }
        double ry[ORC_MAXS], src[ORC_MAXS];
        for (int n = 0; n < ns; n++) {
          ry[n] = A4(rY, i, j, k, n);
          src[n] = A4(rYsrc, i, j, k, n);
        }
        orc_stats s;
        memset(&s, 0, sizeof(s));
        const double fc = react_cell(m, o, use_rk64, ry, src, &A4(T, i, j, k, 0), &A4(rE, i, j, k, 0),
                                     A4(rEsrc, i, j, k, 0), dt, time, &s);
        for (int n = 0; n < ns; n++) A4(rY, i, j, k, n) = ry[n];
        A4(FC, i, j, k, 0) = fc;
        acc(&total, &s);
      }
  if (sum) *sum = total;
  return total.flag;
}

/* flat API: rY [ncells][ns+1] incl. T (ReactorCvode.cpp:1462-1763 CPU branch: no flatten, T taken as given) */
int orc_react_flat(const plm_mech_tables* m, const orc_opts* o, int use_rk64, double* rY, double* rYsrc,
                   double* rX, double* rXsrc, double dt, double time, int ncells, double* FC, orc_stats* sum)
{
  const int ns = m->ns, n = ns + 1;
  orc_stats total;
  memset(&total, 0, sizeof(total));
  for (int c = 0; c < ncells; c++) {
    orc_stats s;
    memset(&s, 0, sizeof(s));
    double fc;
    if (use_rk64) {
      fc = (double)orc_rk64_cell_ex(m, o, rY + (long)c * n, rX[c], rXsrc[c], rYsrc + (long)c * ns, time, dt, 0);
      s.nst = (long)fc;
    } else {
      orc_cvode_cell(m, o, rY + (long)c * n, rX[c], rXsrc[c], rYsrc + (long)c * ns, time, time + dt, &s);
      fc = (double)s.nfe;
    }
    rX[c] = rX[c] + dt * rXsrc[c];
    if (FC) FC[c] = fc;
    acc(&total, &s);
  }
  if (sum) *sum = total;
  return total.flag;
}

typedef struct {
  const plm_mech_tables* m; const orc_opts* o; int use_rk64, ncells, c0, c1;
  double *rY, *rYsrc, *T, *rE, *rEsrc, *FC; const int* mask; double dt, time; orc_stats sum;
} job_t;

static void* worker(void* arg)
{
  job_t* jb = (job_t*)arg;
  const int ns = jb->m->ns;
  const long nc = jb->ncells;
  memset(&jb->sum, 0, sizeof(jb->sum));
  for (int c = jb->c0; c < jb->c1; c++) {
    if (!jb->use_rk64 && jb->mask && jb->mask[c] == -1) {
      jb->FC[c] = 0.0;
      continue;
    }
    double ry[ORC_MAXS], src[ORC_MAXS];
    for (int n = 0; n < ns; n++) {
      ry[n] = jb->rY[n * nc + c];
      src[n] = jb->rYsrc[n * nc + c];
    }
    orc_stats s;
    memset(&s, 0, sizeof(s));
    jb->FC[c] = react_cell(jb->m, jb->o, jb->use_rk64, ry, src, jb->T + c, jb->rE + c, jb->rEsrc[c], jb->dt,
                           jb->time, &s);
    for (int n = 0; n < ns; n++) jb->rY[n * nc + c] = ry[n];
    acc(&jb->sum, &s);
  }
  return 0;
}

int orc_react_soa(const char* mech, const orc_opts* o, int use_rk64, int ncells, double* rY, double* rYsrc,
                  double* T, double* rE, double* rEsrc, double* FC, const int* mask, double dt, double time,
                  int nthreads, orc_stats* sum)
{
  const plm_mech_tables* m = orc_mech(mech);
  if (!m) return -100;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > ncells) nthreads = ncells > 0 ? ncells : 1;
  job_t* jobs = (job_t*)calloc(nthreads, sizeof(job_t));
  pthread_t* th = (pthread_t*)calloc(nthreads, sizeof(pthread_t));
  /* interleaved blocks of 64 cells would balance better; contiguous chunks keep it simple */
  for (int t = 0; t < nthreads; t++) {
    job_t jb = {m, o, use_rk64, ncells, (int)((long)ncells * t / nthreads), (int)((long)ncells * (t + 1) / nthreads),
                rY, rYsrc, T, rE, rEsrc, FC, mask, dt, time, {0}};
    jobs[t] = jb;
    if (nthreads > 1) pthread_create(&th[t], 0, worker, &jobs[t]);
    else worker(&jobs[t]);
  }
  orc_stats total;
  memset(&total, 0, sizeof(total));
  for (int t = 0; t < nthreads; t++) {
    if (nthreads > 1) pthread_join(th[t], 0);
    acc(&total, &jobs[t].sum);
  }
  if (sum) *sum = total;
  free(jobs);
  free(th);
  return total.flag;
}
