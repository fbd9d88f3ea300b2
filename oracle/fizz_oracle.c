/* TEST INFRASTRUCTURE ONLY -- see fizz_oracle.h.  Literal scalar restatement of the reference hot path.
 *
 * Arithmetic contract (measured on numpy 2.3.5 / OpenBLAS 0.3.30):
 *   np.dot([a0,a1],[b0,b1])  == fma(a1, b1, a0*b0)
 *   np.linalg.norm([a0,a1])  == sqrt(fma(a1, a1, a0*a0))
 *   everything else is plain IEEE-754 double, one rounding per operation.
 * Build with -ffp-contract=off so the compiler never fuses on its own (see Makefile).
 */
#include "fizz_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n;
    int fixed;
    double radius, mass;
    /* obj.com : the centre-of-mass Point (physics.py:315-321) */
    double px, py, vx, vy, ax, ay;
    /* obj.com.old* (physics.py:398-400) */
    double opx, opy, ovx, ovy, oax, oay;
    /* obj.pos / obj.vel : the snapshot written by finish_update (physics.py:453-460) */
    double sx, sy, svx, svy;
    /* Polygon.__init__ binds obj.pos to the SAME array as com.pos (physics.py:614); the alias lasts until
     * the first reverse_update (:437 rebinds com.pos) or finish_update (:458 rebinds obj.pos).  SURVEY Q4. */
    int snap_alias;
    double offx[FZO_MAXV], offy[FZO_MAXV]; /* |r|cos(phi), |r|sin(phi) */
    double ptx[FZO_MAXV], pty[FZO_MAXV];   /* point.pos of every vertex */
} body;

struct fzo_world {
    int nfree, nfixed;
    body *b; /* free then fixed: the `objects` order of physics.py:270 */
    fzo_params prm;
    double heat;  /* world.heat (physics.py:53,195) */
    long steps;   /* completed update() calls */
    int status;
    long calls_last, calls_total;
    double ops;
    fzo_contact *trace;
    long trace_cap, ncontacts;
};

typedef struct {
    int i, j;
    double nx, ny, mag;
} hit;

/* np.dot / np.linalg.norm of 2-vectors as the host's numpy evaluates them: host-dependent, hence a
 * build-time switch (make FZO_DOT_MODE=0|1|2 -> libfizz_oracle_dot<m>.so); 1 is the golden host's contract. */
#ifndef FZO_DOT_MODE
#define FZO_DOT_MODE 1
#endif
static inline double dot2(double a0, double a1, double b0, double b1) {
#if FZO_DOT_MODE == 1
    return fma(a1, b1, a0 * b0);
#elif FZO_DOT_MODE == 2
    return fma(a0, b0, a1 * b1);
#else
    return a0 * b0 + a1 * b1;
#endif
}
static inline double norm2(double a0, double a1) { return sqrt(dot2(a0, a1, a0, a1)); }
int fzo_dot_mode(void) { return FZO_DOT_MODE; }

fzo_world *fzo_create(int nfree, int nfixed, const int *nverts, const double *com, const double *vel,
                      const double *radius, const double *mass, const double *offx, const double *offy,
                      const double *fixedxy, const fzo_params *prm) {
    fzo_world *w = (fzo_world *)calloc(1, sizeof(*w));
    int nb = nfree + nfixed, k, v, o = 0;
    w->nfree = nfree;
    w->nfixed = nfixed;
    w->prm = *prm;
    w->b = (body *)calloc((size_t)nb, sizeof(body));
    for (k = 0; k < nfree; k++) {
        body *b = &w->b[k];
        if (nverts[k] > FZO_MAXV) { fzo_destroy(w); return NULL; }
        b->n = nverts[k];
        b->radius = radius[k];
        b->mass = mass[k];
        b->px = com[2 * k]; b->py = com[2 * k + 1];
        b->vx = vel[2 * k]; b->vy = vel[2 * k + 1];
        b->ax = 0.0; b->ay = 0.0; /* Point.acc = [0,0] (physics.py:494) */
        b->sx = b->px; b->sy = b->py; b->svx = b->vx; b->svy = b->vy;
        b->snap_alias = 1;
        for (v = 0; v < b->n; v++, o++) {
            b->offx[v] = offx[o];
            b->offy[v] = offy[o];
            /* the input vertex; never read before the first pre_update overwrites it, except by verts() */
            b->ptx[v] = NAN; b->pty[v] = NAN;
        }
    }
    o = 0;
    for (k = nfree; k < nb; k++) {
        body *b = &w->b[k];
        if (nverts[k] > FZO_MAXV) { fzo_destroy(w); return NULL; }
        b->n = nverts[k];
        b->fixed = 1;
        b->radius = radius[k];
        b->mass = INFINITY;
        b->px = b->sx = com[2 * k];
        b->py = b->sy = com[2 * k + 1];
        for (v = 0; v < b->n; v++, o++) {
            b->ptx[v] = fixedxy[2 * o];
            b->pty[v] = fixedxy[2 * o + 1];
        }
    }
    return w;
}

void fzo_destroy(fzo_world *w) {
    if (!w) return;
    free(w->b);
    free(w);
}

/* Obj.pre_update (physics.py:384-427) with Point.move (physics.py:502-529) inlined.  move_angular
 * (:531-557) yields 0 for every shipped shape, so rotate() (:350-363) rebuilds each vertex as
 * |r|cos(phi) + com.x , |r|sin(phi) + com.y (SURVEY Q1,Q2). */
static void pre_update(fzo_world *w, body *b, double dt) {
    const fzo_params *p = &w->prm;
    double halfdt, uvx, uvy, vax, vay, uxx, uxy;
    int v;
    b->opx = b->px; b->opy = b->py; b->ovx = b->vx; b->ovy = b->vy; b->oax = b->ax; b->oay = b->ay;
    halfdt = .5 * dt;               /* :507 */
    uvx = b->ax * halfdt;           /* :508 */
    uvy = b->ay * halfdt;
    vax = b->vx + uvx;              /* :509 */
    vay = b->vy + uvy;
    uxx = vax * dt;                 /* :512 */
    uxy = vay * dt;
    b->px += uxx;                   /* :513 (in place: an aliased obj.pos follows) */
    b->py += uxy;
    if (b->snap_alias) { b->sx = b->px; b->sy = b->py; }
    /* :516  sum([1*global_accel]) / 1   (Obj.mass is still 1 when init_obj runs; com.mass is 1; Q10) */
    b->ax = (0.0 + 1.0 * p->gx) / 1.0;
    b->ay = (0.0 + 1.0 * p->gy) / 1.0;
    b->vx = vax + b->ax * halfdt;   /* :521 */
    b->vy = vay + b->ay * halfdt;
    for (v = 0; v < b->n; v++) {    /* :358-359 */
        b->ptx[v] = b->offx[v] + b->px;
        b->pty[v] = b->offy[v] + b->py;
    }
    w->ops += 12.0 + 2.0 * b->n;
}

/* Obj.reverse_update (physics.py:430-451): restores the COM only; vertices keep the last trial's
 * positions because point.oldpos was saved after the rebuild. */
static void reverse_update(body *b) {
    b->px = b->opx; b->py = b->opy; b->vx = b->ovx; b->vy = b->ovy; b->ax = b->oax; b->ay = b->oay;
    b->snap_alias = 0; /* :437 rebinds com.pos; obj.pos keeps the old array */
}

/* Obj.finish_update (physics.py:453-464) */
static void finish_update(body *b) {
    b->sx = b->px; b->sy = b->py; b->svx = b->vx; b->svy = b->vy;
    b->snap_alias = 0;
}

/* polypoly_collision (physics.py:765-840). Returns 1 and (nx,ny,mag) on overlap, 0 on a separating axis. */
static int polypoly(fzo_world *w, const body *p1, const body *p2, double *onx, double *ony, double *omag) {
    double overlap = INFINITY, fx = NAN, fy = NAN;
    int have = 0, k, v, n1 = p1->n, n2 = p2->n;
    for (k = 0; k < n1 + n2; k++) {
        const body *src = (k < n1) ? p1 : p2;
        int e = (k < n1) ? k : k - n1;
        int e2 = (e + 1 == src->n) ? 0 : e + 1;
        /* :775,:778  normal of edge (p1,p2) = (p2.y - p1.y, p1.x - p2.x) */
        double ax = src->pty[e2] - src->pty[e];
        double ay = src->ptx[e] - src->ptx[e2];
        double nrm = norm2(ax, ay);       /* :790 */
        double nx = ax / nrm, ny = ay / nrm;
        double b1lo, b1hi, b2lo, b2hi, cur, t;
        w->ops += 12.0 + 5.0 * (n1 + n2);
        b1lo = b1hi = dot2(nx, ny, p1->ptx[0], p1->pty[0]); /* :795 */
        b2lo = b2hi = dot2(nx, ny, p2->ptx[0], p2->pty[0]); /* :796 */
        for (v = 0; v < n1; v++) {        /* :799-804 */
            cur = dot2(nx, ny, p1->ptx[v], p1->pty[v]);
            if (cur < b1lo) b1lo = cur;
            if (cur > b1hi) b1hi = cur;
        }
        for (v = 0; v < n2; v++) {        /* :806-811 */
            cur = dot2(nx, ny, p2->ptx[v], p2->pty[v]);
            if (cur < b2lo) b2lo = cur;
            if (cur > b2hi) b2hi = cur;
        }
        if (b1lo > b2lo) {                /* :815-817 */
            t = b1lo; b1lo = b2lo; b2lo = t;
            t = b1hi; b1hi = b2hi; b2hi = t;
        }
        if (b1hi < b2lo) return 0;        /* :820-821 */
        cur = b1hi - b2lo;                /* :824 */
        if (cur < overlap) {              /* :827-830 */
            fx = nx; fy = ny; overlap = cur; have = 1;
        }
    }
    if (!have) return 0; /* all-NaN axes: the reference raises UnboundLocalError at :838 (Q13) */
    *onx = fx * -1.0;    /* :838-839  fix_mag_flag is always truthy (Q7) */
    *ony = fy * -1.0;
    *omag = overlap;
    return 1;
}

/* World.check_collisions (physics.py:262-289). `site`: 0 trial (:124), 1 fallback (:137), 2 resolve (:234). */
static int check_collisions(fzo_world *w, hit *out) {
    int nb = w->nfree + w->nfixed, i, j, n = 0;
    w->calls_last++;
    w->calls_total++;
    for (i = 0; i < nb; i++) {
        const body *a = &w->b[i];
        for (j = i + 1; j < nb; j++) {
            const body *c = &w->b[j];
            double d, r, nx, ny, mag;
            if (a->fixed && c->fixed) continue; /* :280 (its norm has no side effect) */
            d = norm2(a->sx - c->sx, a->sy - c->sy);           /* :277 */
            r = (c->radius > a->radius) ? c->radius : a->radius; /* python max(a, b): b only if b > a */
            w->ops += 8.0;
            if (d > r) continue;                               /* :278 */
            if (polypoly(w, a, c, &nx, &ny, &mag)) {           /* :286-288 */
                out[n].i = i; out[n].j = j; out[n].nx = nx; out[n].ny = ny; out[n].mag = mag;
                if (w->trace) {
                    if (w->ncontacts < w->trace_cap) {
                        fzo_contact *t = &w->trace[w->ncontacts];
                        t->step = (int)w->steps; t->call = (int)w->calls_last; t->i = i; t->j = j;
                        t->nx = nx; t->ny = ny; t->mag = mag;
                    }
                    w->ncontacts++;
                }
                n++;
            }
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------------------------
 * extensions: Coulomb friction and free-free restitution.  The reference has neither (SURVEY
 * section 0, Q9); notes/collisions.tex:16-35 derives the restitution formula.  This is their DEFINITION, written
 * independently of fizz2d_b200/csrc/fizz_ext.cuh from the same model:
 *   t = (-u_y, u_x); the tangential (relative) velocity is reduced by an impulse of at most mu x the normal impulse,
 *   sticking when less is enough; kinetic energy removed -> heat, so K + P + H stays constant.
 * ------------------------------------------------------------------------------------------------------------ */
static void ext_friction_fixed(double mu, double ope, double m1, double ux, double uy, double vdotN,
                               double *vx, double *vy, double *heat) {
    double tx = -uy, ty = ux;
    double vt = dot2(*vx, *vy, tx, ty);
    double lim = mu * (ope * fabs(vdotN));
    double dvt = -vt, vt1;
    if (fabs(vt) > lim) dvt = (vt > 0.0) ? -lim : lim;
    *vx += dvt * tx;
    *vy += dvt * ty;
    vt1 = vt + dvt;
    *heat += (.5 * m1) * (vt * vt - vt1 * vt1);
}

static void ext_restitution_free(double cr, double m1, double mprop1, double mprop2, double vn1, double vn2,
                                 double *v1u, double *v2u, double *heat) {
    double mean = mprop1 * vn1 + mprop2 * vn2;
    double dvn = vn1 - vn2;
    *v1u = mean + (mprop2 * cr) * (vn2 - vn1);
    *v2u = mean + (mprop1 * cr) * (vn1 - vn2);
    *heat += ((.5 * (m1 * mprop2)) * (1.0 - cr * cr)) * (dvn * dvn);
}

static void ext_friction_free(double mu, double cr_eff, double m1, double m2, double mprop2, double ux, double uy,
                              double vn1, double vn2, double *v1x, double *v1y, double *v2x, double *v2y, double *heat) {
    double tx = -uy, ty = ux;
    double vt1 = dot2(*v1x, *v1y, tx, ty), vt2 = dot2(*v2x, *v2y, tx, ty);
    double rel = vt1 - vt2;
    double mred = m1 * mprop2;
    double jmax = mu * (((1.0 + cr_eff) * fabs(vn1 - vn2)) * mred);
    double jt = mred * fabs(rel);
    double sj, dv1, dv2, w1, w2;
    if (jt > jmax) jt = jmax;
    sj = (rel > 0.0) ? -jt : jt;
    dv1 = sj / m1;
    dv2 = -(sj / m2);
    *v1x += dv1 * tx; *v1y += dv1 * ty;
    *v2x += dv2 * tx; *v2y += dv2 * ty;
    w1 = vt1 + dv1; w2 = vt2 + dv2;
    *heat += (.5 * m1) * (vt1 * vt1 - w1 * w1) + (.5 * m2) * (vt2 * vt2 - w2 * w2);
}

int fzo_update(fzo_world *w) {
    const fzo_params *p = &w->prm;
    int nf = w->nfree, k, v, ncol, c;
    int npairs = (nf + w->nfixed) * (nf + w->nfixed - 1) / 2;
    hit *col;
    double dt, max_overlap;
    int first_iter = 1;
    if (w->status) return 1;
    col = (hit *)malloc(sizeof(hit) * (size_t)(npairs > 0 ? npairs : 1));
    w->calls_last = 0;

    dt = p->time_disc;                                  /* :89 */
    max_overlap = p->collision_tol + 1;                 /* :90 */
    ncol = 0;
    while (max_overlap > p->collision_tol && dt > p->time_tol) { /* :99 */
        if (first_iter) {
            first_iter = 0;
        } else {
            dt /= 2;                                    /* :106 */
            for (k = 0; k < nf; k++) reverse_update(&w->b[k]);
        }
        for (k = 0; k < nf; k++) pre_update(w, &w->b[k], dt); /* :116-118 */
        ncol = check_collisions(w, col);                /* :124 */
        max_overlap = 0;                                /* :125-129 */
        if (ncol) {
            max_overlap = fabs(col[0].mag);
            for (c = 1; c < ncol; c++) {
                double m = fabs(col[c].mag);
                if (m > max_overlap) max_overlap = m;
            }
        }
    }
    if (dt < p->time_tol) {                             /* :133-137 */
        for (k = 0; k < nf; k++) reverse_update(&w->b[k]);
        ncol = check_collisions(w, col);
    }

    while (ncol) {                                      /* :140 */
        for (c = 0; c < ncol; c++) {                    /* :142 (list computed before the pass, Q11) */
            body *o1 = &w->b[col[c].i], *o2 = &w->b[col[c].j];
            double ux = col[c].nx, uy = col[c].ny;
            double s = col[c].mag + p->epsilon;         /* :158 */
            double nvx = ux * s, nvy = uy * s;
            if (o1->fixed) { body *t = o1; o1 = o2; o2 = t; } /* :145 (never taken: free bodies come first) */
            if (o2->fixed) {                            /* :171-198 */
                double vdotN, cc, H;
                for (v = 0; v < o1->n; v++) { o1->ptx[v] += nvx; o1->pty[v] += nvy; } /* :174 */
                if (norm2(nvx, nvy) > p->vibrate_tol) { /* :180-181 */
                    o1->px += nvx; o1->py += nvy;
                    if (o1->snap_alias) { o1->sx = o1->px; o1->sy = o1->py; }
                }
                vdotN = dot2(o1->vx, o1->vy, ux, uy);   /* :183 */
                cc = p->one_plus_e * vdotN;             /* :188 */
                o1->vx -= cc * ux;
                o1->vy -= cc * uy;
                H = .5 * o1->mass * p->one_minus_e2 * (vdotN * vdotN); /* :191 */
                w->heat += H;                           /* :195 */
                if (p->friction_mu > 0.0)               /* extension, off by default */
                    ext_friction_fixed(p->friction_mu, p->one_plus_e, o1->mass, ux, uy, vdotN, &o1->vx, &o1->vy, &w->heat);
                finish_update(o1);                      /* :197 */
                w->ops += 20.0 + 2.0 * o1->n;
            } else {                                    /* :201-223 */
                double mtotal = o1->mass + o2->mass;
                double mprop1 = o1->mass / mtotal;
                double mprop2 = 1 - mprop1;
                double vn1, vn2, v1u, v2u, d1, d2;
                o1->px += mprop1 * nvx; o1->py += mprop1 * nvy;          /* :206 */
                if (o1->snap_alias) { o1->sx = o1->px; o1->sy = o1->py; }
                vn1 = dot2(o1->vx, o1->vy, ux, uy);                      /* :207 */
                vn2 = dot2(o2->vx, o2->vy, ux, uy);                      /* :208 */
                v1u = (mprop1 - mprop2) * vn1 + 2 * mprop2 * vn2;        /* :209 */
                v2u = 2 * mprop1 * vn1 + (mprop2 - mprop1) * vn2;        /* :218 (needs nothing computed in between) */
                if (p->ff_restitution >= 0.0)                            /* extension, off by default */
                    ext_restitution_free(p->ff_restitution, o1->mass, mprop1, mprop2, vn1, vn2, &v1u, &v2u, &w->heat);
                d1 = v1u - dot2(o1->vx, o1->vy, ux, uy);                 /* :210 */
                o1->vx += d1 * ux; o1->vy += d1 * uy;
                for (v = 0; v < o1->n; v++) { o1->ptx[v] += mprop1 * nvx; o1->pty[v] += mprop1 * nvy; } /* :212 */
                finish_update(o1);                                       /* :215 */
                o2->px -= mprop2 * nvx; o2->py -= mprop2 * nvy;          /* :217 */
                if (o2->snap_alias) { o2->sx = o2->px; o2->sy = o2->py; }
                d2 = v2u - dot2(o2->vx, o2->vy, ux, uy);                 /* :219 */
                o2->vx += d2 * ux; o2->vy += d2 * uy;
                if (p->friction_mu > 0.0)                                /* extension, off by default */
                    ext_friction_free(p->friction_mu, p->ff_restitution >= 0.0 ? p->ff_restitution : 1.0, o1->mass,
                                      o2->mass, mprop2, ux, uy, vn1, vn2, &o1->vx, &o1->vy, &o2->vx, &o2->vy, &w->heat);
                for (v = 0; v < o2->n; v++) { o2->ptx[v] -= mprop2 * nvx; o2->pty[v] -= mprop2 * nvy; } /* :221 */
                finish_update(o2);                                       /* :223 */
                w->ops += 40.0 + 2.0 * (o1->n + o2->n);
            }
        }
        if (w->calls_last >= p->max_calls) {            /* watchdog (not in the reference) */
            w->status = 1;
            free(col);
            return 1;
        }
        ncol = check_collisions(w, col);                /* :234 */
    }

    for (k = 0; k < nf; k++) finish_update(&w->b[k]);   /* :237-239 */
    /* :243-247 and :253-256 are print-only check_collisions calls: no state change, not restated */
    if (dt != p->time_disc) {                           /* :248-251 */
        for (k = 0; k < nf; k++) {
            pre_update(w, &w->b[k], p->time_disc - dt);
            finish_update(&w->b[k]);
        }
    }
    w->steps++;                                         /* :259 */
    free(col);
    return 0;
}

long fzo_run(fzo_world *w, long nsteps) {
    long t;
    for (t = 0; t < nsteps; t++)
        if (fzo_update(w)) break;
    return w->steps;
}

/* World.energy (physics.py:66-74), Obj.k_energy (:368-369), Obj.p_energy (:371-374) */
void fzo_energy(const fzo_world *w, double *out) {
    double K = 0.0, P = 0.0;
    int k;
    for (k = 0; k < w->nfree; k++) {
        const body *b = &w->b[k];
        double nv = norm2(b->svx, b->svy);
        K += (.5 * b->mass) * (nv * nv);
        P += (w->prm.gy * b->mass) * (w->prm.height - b->sy);
    }
    out[1] = K;
    out[2] = P;
    out[3] = w->heat;
    out[0] = ((0.0 + out[1]) + out[2]) + out[3]; /* builtin sum(E[1:4]) */
}

void fzo_com(const fzo_world *w, double *out) {
    int k;
    for (k = 0; k < w->nfree; k++) {
        out[4 * k + 0] = w->b[k].px; out[4 * k + 1] = w->b[k].py;
        out[4 * k + 2] = w->b[k].vx; out[4 * k + 3] = w->b[k].vy;
    }
}

void fzo_verts(const fzo_world *w, double *out) {
    int k, v, o = 0;
    for (k = 0; k < w->nfree; k++)
        for (v = 0; v < w->b[k].n; v++, o++) {
            out[2 * o] = w->b[k].ptx[v];
            out[2 * o + 1] = w->b[k].pty[v];
        }
}

double fzo_heat(const fzo_world *w) { return w->heat; }
long fzo_steps_done(const fzo_world *w) { return w->steps; }
int fzo_status(const fzo_world *w) { return w->status; }
long fzo_calls_last_step(const fzo_world *w) { return w->calls_last; }
long fzo_total_calls(const fzo_world *w) { return w->calls_total; }
double fzo_total_ops(const fzo_world *w) { return w->ops; }

void fzo_trace(fzo_world *w, fzo_contact *buf, long cap) {
    w->trace = buf;
    w->trace_cap = cap;
    w->ncontacts = 0;
}
long fzo_contact_count(const fzo_world *w) { return w->ncontacts; }

typedef struct {
    fzo_world **worlds;
    long n, nsteps;
    long *next;
    pthread_mutex_t *mu;
} job;

static void *worker(void *arg) {
    job *j = (job *)arg;
    for (;;) {
        long lo, hi, k;
        pthread_mutex_lock(j->mu);
        lo = *j->next;
        *j->next += 16;
        pthread_mutex_unlock(j->mu);
        if (lo >= j->n) break;
        hi = lo + 16 < j->n ? lo + 16 : j->n;
        for (k = lo; k < hi; k++) fzo_run(j->worlds[k], j->nsteps);
    }
    return NULL;
}

void fzo_run_batch(fzo_world **worlds, long n, long nsteps, int nthreads) {
    pthread_t *th;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    long next = 0;
    job j;
    int t;
    if (nthreads < 1) nthreads = 1;
    th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    j.worlds = worlds; j.n = n; j.nsteps = nsteps; j.next = &next; j.mu = &mu;
    for (t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, worker, &j);
    for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
}
