/* orc_ekf.c -- CPU restatement of the reference's EKF: ArucoSlam::addEncoder
 * (src/aruco_slam.cpp:21-74) and the correction / augmentation loop of
 * ArucoSlam::addImage (src/aruco_slam.cpp:88-263), dense row-major doubles in
 * place of Eigen.  TEST INFRASTRUCTURE ONLY (see oracle.h).  PINNED against runs
 * of the reference itself (oracle/_ref = the unmodified aruco_slam.cpp; vectors in
 * tests/golden/slam_*.npz, tests/test_slam_golden.py: mu, Sigma <= 1e-9 per frame,
 * landmark order = the reference's priority-queue order).  dt is an argument instead of
 * ros::Time::now() (:26,31-32); the log prints (:79,:89,:96,:161-171,:283-286)
 * are omitted.  Quirks kept: frame-start snapshot `mu`
 * (:88), float sin/cos in augmentation (:210-211), (I-KG)Sigma form (:204),
 * single-wrap normAngle (:412-421), kl used for both wheels in wkh (:62), the
 * "stationary" branch (:193-198) skips the update (its mu_.topLeftCorner(3, 0) is a
 * 3 x 0 block) and leaves that marker's last_observation_ unset.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct orc_ekf {
    orc_slam_params sp;
    int N;                 /* 3 + 3 n */
    double *mu, *sigma;    /* N, N*N row-major */
    int32_t *ids;          /* landmark k -> aruco id  (aruco_id_map, aruco_slam.h:164) */
    /* last_observed_marker_ (aruco_slam.h:188): id + last_observation_ (NaN = never set) */
    int n_last; int32_t *last_ids; double *last_obs;
};

static void norm_angle(double *a)
{
    const double PI = 3.14159265358979323846, TWO_PI = 2.0 * PI;
    if (*a >= PI) *a -= TWO_PI;
    if (*a < -PI) *a += TWO_PI;
}

orc_ekf *orc_ekf_create(const orc_slam_params *sp)
{
    orc_ekf *e = (orc_ekf *)calloc(1, sizeof(orc_ekf));
    e->sp = *sp; e->N = 3;
    e->mu = (double *)calloc(3, sizeof(double));
    e->sigma = (double *)calloc(9, sizeof(double));
    e->ids = (int32_t *)calloc(1, sizeof(int32_t));
    return e;
}
void orc_ekf_destroy(orc_ekf *e) { if (!e) return; free(e->mu); free(e->sigma); free(e->ids); free(e->last_ids); free(e->last_obs); free(e); }
int orc_ekf_dim(const orc_ekf *e) { return e->N; }
void orc_ekf_get_state(const orc_ekf *e, double *mu, double *sigma, int32_t *ids)
{
    if (mu) memcpy(mu, e->mu, sizeof(double) * (size_t)e->N);
    if (sigma) memcpy(sigma, e->sigma, sizeof(double) * (size_t)e->N * e->N);
    if (ids) memcpy(ids, e->ids, sizeof(int32_t) * (size_t)((e->N - 3) / 3));
}
void orc_ekf_set_state(orc_ekf *e, int N, const double *mu, const double *sigma, const int32_t *ids)
{
    e->N = N;
    e->mu = (double *)realloc(e->mu, sizeof(double) * (size_t)N);
    e->sigma = (double *)realloc(e->sigma, sizeof(double) * (size_t)N * N);
    e->ids = (int32_t *)realloc(e->ids, sizeof(int32_t) * (size_t)((N - 3) / 3 + 1));
    memcpy(e->mu, mu, sizeof(double) * (size_t)N);
    memcpy(e->sigma, sigma, sizeof(double) * (size_t)N * N);
    memcpy(e->ids, ids, sizeof(int32_t) * (size_t)((N - 3) / 3));
}

void orc_ekf_predict(orc_ekf *e, double wl, double wr, double dt)
{
    const orc_slam_params *sp = &e->sp;
    int N = e->N;
    double delta_enl = dt * wl, delta_enr = dt * wr;                 /* :35-36 */
    double delta_sl = sp->kl * delta_enl, delta_sr = sp->kr * delta_enr;
    double l_ = 2 * sp->b;
    double delta_theta = (delta_sr - delta_sl) / l_;
    double delta_s = 0.5 * (delta_sr + delta_sl);
    double tmp_th = e->mu[2] + 0.5 * delta_theta;
    double c = cos(tmp_th), s = sin(tmp_th);
    e->mu[0] += delta_s * c; e->mu[1] += delta_s * s; e->mu[2] += delta_theta;
    norm_angle(&e->mu[2]);
    double Hxi[9] = {1, 0, -delta_s * s, 0, 1, delta_s * c, 0, 0, 1};
    double f = 0.5 * sp->kl * dt;
    double wkh[6] = {c * f, c * f, s * f, s * f, (1 / sp->b) * f, (-1 / sp->b) * f};   /* 3x2 */
    double su[2] = {sp->Q_k * fabs(wl), sp->Q_k * fabs(wr)};
    double Qk[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        Qk[3 * i + j] = wkh[2 * i] * su[0] * wkh[2 * j] + wkh[2 * i + 1] * su[1] * wkh[2 * j + 1];
    /* Sigma = Hx Sigma Hx^T + F Qk F^T with Hx = blkdiag(Hxi, I): dense like :64-73 */
    double *Hx = (double *)calloc((size_t)N * N, sizeof(double));
    for (int i = 0; i < N; i++) Hx[(size_t)i * N + i] = 1;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Hx[(size_t)i * N + j] = Hxi[3 * i + j];
    double *T = (double *)calloc((size_t)N * N, sizeof(double)), *S2 = (double *)calloc((size_t)N * N, sizeof(double));
    for (int i = 0; i < N; i++) for (int k = 0; k < N; k++) { double a = Hx[(size_t)i * N + k]; if (a == 0) continue; for (int j = 0; j < N; j++) T[(size_t)i * N + j] += a * e->sigma[(size_t)k * N + j]; }
    for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) { double s2 = 0; for (int k = 0; k < N; k++) { double a = Hx[(size_t)j * N + k]; if (a != 0) s2 += T[(size_t)i * N + k] * a; } S2[(size_t)i * N + j] = s2; }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) S2[(size_t)i * N + j] += Qk[3 * i + j];
    memcpy(e->sigma, S2, sizeof(double) * (size_t)N * N);
    free(Hx); free(T); free(S2);
}

static void inv3(const double *a, double *t)
{
    double d = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    d = 1. / d;
    t[0] = (a[4] * a[8] - a[5] * a[7]) * d; t[1] = (a[2] * a[7] - a[1] * a[8]) * d; t[2] = (a[1] * a[5] - a[2] * a[4]) * d;
    t[3] = (a[5] * a[6] - a[3] * a[8]) * d; t[4] = (a[0] * a[8] - a[2] * a[6]) * d; t[5] = (a[2] * a[3] - a[0] * a[5]) * d;
    t[6] = (a[3] * a[7] - a[4] * a[6]) * d; t[7] = (a[1] * a[6] - a[0] * a[7]) * d; t[8] = (a[0] * a[4] - a[1] * a[3]) * d;
}

typedef struct { orc_observation o; int seq; } qitem;
/* std::priority_queue<ArucoMarker> (aruco_slam.h:190) with operator< = (a.index > b.index) (:85-88): pops ascending
 * aruco_index_.  The order among EQUAL indices (all new landmarks carry -1) is decided by the heap algorithm; the
 * reference is built with GCC, so this restates libstdc++'s std::push_heap / std::pop_heap (bits/stl_heap.h:
 * __push_heap sift-up, __adjust_heap sift-down to a leaf then sift-up) -- pinned against the reference build
 * (oracle/_ref) in tests/golden/slam_*.npz. */
static int q_less(const qitem *a, const qitem *b) { return a->o.aruco_index > b->o.aruco_index; }
static void heap_push_up(qitem *h, int hole, int top, qitem v)
{
    int parent = (hole - 1) / 2;
    while (hole > top && q_less(&h[parent], &v)) { h[hole] = h[parent]; hole = parent; parent = (hole - 1) / 2; }
    h[hole] = v;
}
static void heap_push(qitem *h, int *n, qitem v) { h[*n] = v; ++*n; heap_push_up(h, *n - 1, 0, v); }
static qitem heap_pop(qitem *h, int *n)
{
    qitem top = h[0];
    if (*n > 1) {
        int len = *n - 1;                    /* the heap that remains; value = the old last element */
        qitem v = h[len];
        int hole = 0, child = 0;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (q_less(&h[child], &h[child - 1])) child--;
            h[hole] = h[child]; hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) { child = 2 * (child + 1); h[hole] = h[child - 1]; hole = child - 1; }
        heap_push_up(h, hole, 0, v);
    }
    --*n;
    return top;
}

void orc_ekf_update(orc_ekf *e, const orc_observation *obs_in, int n, int dense)
{
    if (n <= 0) { e->n_last = 0; return; }
    qitem *heap = (qitem *)malloc(sizeof(qitem) * (size_t)n), *q = (qitem *)malloc(sizeof(qitem) * (size_t)n);
    int nh = 0;
    for (int i = 0; i < n; i++) {
        qitem it;
        it.o = obs_in[i]; it.seq = i;
        int nl = (e->N - 3) / 3, idx = -1;                       /* checkLandmark :423-435 */
        for (int k = 0; k < nl; k++) if (e->ids[k] == obs_in[i].aruco_id) { idx = k; break; }
        it.o.aruco_index = idx;
        heap_push(heap, &nh, it);                                /* obs_.push(ob) :373 */
    }
    for (int i = 0; i < n; i++) q[i] = heap_pop(heap, &nh);      /* obs_.top(); obs_.pop() :94-95 */
    free(heap);
    int N0 = e->N;
    double *mu = (double *)malloc(sizeof(double) * (size_t)N0);   /* snapshot :88 */
    memcpy(mu, e->mu, sizeof(double) * (size_t)N0);
    int32_t *new_last_ids = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    double *new_last_obs = (double *)malloc(sizeof(double) * 3 * (size_t)n);

    for (int qi = 0; qi < n; qi++) {
        orc_observation *ob = &q[qi].o;
        const double *Rk = ob->cov;
        double lastobs[3] = {NAN, NAN, NAN};
        if (ob->aruco_index >= 0) {
            int N = e->N, k = ob->aruco_index, L = 3 + 3 * k;
            double mx = mu[L], my = mu[L + 1], mth = mu[L + 2];
            double x = mu[0], y = mu[1], th = mu[2];
            double s = sin(th), c = cos(th);
            double gdx = mx - x, gdy = my - y, gdt = mth - th;
            norm_angle(&gdt);
            double zhat[3] = {gdx * c + gdy * s, -gdx * s + gdy * c, gdt};
            double z[3] = {ob->x, ob->y, ob->theta};
            double ze[3] = {z[0] - zhat[0], z[1] - zhat[1], z[2] - zhat[2]};
            norm_angle(&ze[2]);
            double Gxm[18] = {-c, -s, -gdx * s + gdy * c, c, s, 0,
                              s, -c, -gdx * c - gdy * s, -s, c, 0,
                              0, 0, -1, 0, 0, 1};
            int cols[6] = {0, 1, 2, L, L + 1, L + 2};
            /* SGt = Sigma Gx^T  (N x 3) */
            double *SGt = (double *)malloc(sizeof(double) * 3 * (size_t)N);
            for (int i = 0; i < N; i++) for (int r = 0; r < 3; r++) {
                double a = 0;
                for (int j = 0; j < 6; j++) a += e->sigma[(size_t)i * N + cols[j]] * Gxm[6 * r + j];
                SGt[3 * i + r] = a;
            }
            double S[9], Si[9];
            for (int r = 0; r < 3; r++) for (int c2 = 0; c2 < 3; c2++) {
                double a = 0;
                for (int j = 0; j < 6; j++) a += Gxm[6 * r + j] * SGt[3 * cols[j] + c2];
                S[3 * r + c2] = a + Rk[3 * r + c2];
            }
            inv3(S, Si);
            double *K = (double *)malloc(sizeof(double) * 3 * (size_t)N);
            for (int i = 0; i < N; i++) for (int c2 = 0; c2 < 3; c2++)
                K[3 * i + c2] = SGt[3 * i] * Si[c2] + SGt[3 * i + 1] * Si[3 + c2] + SGt[3 * i + 2] * Si[6 + c2];
            /* stationary gate :192-198 */
            int stationary = 0;
            for (int l = 0; l < e->n_last; l++) if (e->last_ids[l] == ob->aruco_id) {
                double d0 = e->last_obs[3 * l] - z[0], d1 = e->last_obs[3 * l + 1] - z[1], d2 = e->last_obs[3 * l + 2] - z[2];
                double nn = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
                if (nn < 0.01) stationary = 1;      /* NaN compares false */
                break;
            }
            if (!stationary) {
                lastobs[0] = z[0]; lastobs[1] = z[1]; lastobs[2] = z[2];          /* :202 */
                for (int i = 0; i < N; i++) e->mu[i] += K[3 * i] * ze[0] + K[3 * i + 1] * ze[1] + K[3 * i + 2] * ze[2];   /* :203 */
                if (dense) {                                                     /* :204 as written */
                    double *IKG = (double *)calloc((size_t)N * N, sizeof(double));
                    for (int i = 0; i < N; i++) {
                        IKG[(size_t)i * N + i] = 1;
                        for (int j = 0; j < 6; j++) {
                            double kg = K[3 * i] * Gxm[j] + K[3 * i + 1] * Gxm[6 + j] + K[3 * i + 2] * Gxm[12 + j];
                            IKG[(size_t)i * N + cols[j]] -= kg;
                        }
                    }
                    double *S2 = (double *)calloc((size_t)N * N, sizeof(double));
                    for (int i = 0; i < N; i++) for (int kk = 0; kk < N; kk++) {
                        double a = IKG[(size_t)i * N + kk];
                        if (a == 0) continue;
                        for (int j = 0; j < N; j++) S2[(size_t)i * N + j] += a * e->sigma[(size_t)kk * N + j];
                    }
                    memcpy(e->sigma, S2, sizeof(double) * (size_t)N * N);
                    free(IKG); free(S2);
                } else {                                                         /* Sigma -= K (Gx Sigma) */
                    double *GS = (double *)malloc(sizeof(double) * 3 * (size_t)N);
                    for (int r = 0; r < 3; r++) for (int j = 0; j < N; j++) {
                        double a = 0;
                        for (int jj = 0; jj < 6; jj++) a += Gxm[6 * r + jj] * e->sigma[(size_t)cols[jj] * N + j];
                        GS[(size_t)r * N + j] = a;
                    }
                    for (int i = 0; i < N; i++) for (int j = 0; j < N; j++)
                        e->sigma[(size_t)i * N + j] -= K[3 * i] * GS[j] + K[3 * i + 1] * GS[N + j] + K[3 * i + 2] * GS[2 * N + j];
                    free(GS);
                }
            } else {
                /* mu_.topLeftCorner(3,0) += ... : a 3x0 block, no-op; last_observation_ of this ob stays unset */
            }
            free(SGt); free(K);
        } else {
            /* new landmark :208-260 */
            float sinth = (float)sin(mu[2]), costh = (float)cos(mu[2]);
            int N = e->N;
            double map_x = mu[0] + costh * ob->x - sinth * ob->y;
            double map_y = mu[1] + sinth * ob->x + costh * ob->y;
            double map_th = mu[2] + ob->theta;
            norm_angle(&map_th);
            double dx = map_x - mu[0], dy = map_y - mu[1];
            double Gsk[9] = {-costh, -sinth, -sinth * dx + costh * dy,
                             sinth, -costh, -dx * costh - dy * sinth,
                             0, 0, -1};
            double Gmi[9] = {costh, sinth, 0, -sinth, costh, 0, 0, 0, 1};
            double Ss[9];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Ss[3 * i + j] = e->sigma[(size_t)i * N + j];
            double A[9], B[9];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double a = 0; for (int k2 = 0; k2 < 3; k2++) a += Gsk[3 * i + k2] * Ss[3 * k2 + j]; A[3 * i + j] = a; }
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double a = 0; for (int k2 = 0; k2 < 3; k2++) a += A[3 * i + k2] * Gsk[3 * j + k2]; B[3 * i + j] = a + Rk[3 * i + j]; }
            /* sigma_mm = Gmi * B^T * Gmi^T */
            double C[9], Smm[9];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double a = 0; for (int k2 = 0; k2 < 3; k2++) a += Gmi[3 * i + k2] * B[3 * j + k2]; C[3 * i + j] = a; }
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double a = 0; for (int k2 = 0; k2 < 3; k2++) a += C[3 * i + k2] * Gmi[3 * j + k2]; Smm[3 * i + j] = a; }
            /* sigma_mx = -Gmi Gsk Sigma[0:3,:]  (3 x N) */
            double GG[9];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double a = 0; for (int k2 = 0; k2 < 3; k2++) a += (-Gmi[3 * i + k2]) * Gsk[3 * k2 + j]; GG[3 * i + j] = a; }
            double *Smx = (double *)malloc(sizeof(double) * 3 * (size_t)N);
            for (int i = 0; i < 3; i++) for (int j = 0; j < N; j++) { double a = 0; for (int k2 = 0; k2 < 3; k2++) a += GG[3 * i + k2] * e->sigma[(size_t)k2 * N + j]; Smx[(size_t)i * N + j] = a; }
            int N2 = N + 3;
            double *ns = (double *)calloc((size_t)N2 * N2, sizeof(double));
            for (int i = 0; i < N; i++) memcpy(ns + (size_t)i * N2, e->sigma + (size_t)i * N, sizeof(double) * (size_t)N);
            for (int i = 0; i < 3; i++) for (int j = 0; j < N; j++) { ns[(size_t)(N + i) * N2 + j] = Smx[(size_t)i * N + j]; ns[(size_t)j * N2 + N + i] = Smx[(size_t)i * N + j]; }
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) ns[(size_t)(N + i) * N2 + N + j] = Smm[3 * i + j];
            free(Smx); free(e->sigma); e->sigma = ns;
            e->mu = (double *)realloc(e->mu, sizeof(double) * (size_t)N2);
            e->mu[N] = map_x; e->mu[N + 1] = map_y; e->mu[N + 2] = map_th;
            e->ids = (int32_t *)realloc(e->ids, sizeof(int32_t) * (size_t)((N2 - 3) / 3));
            e->ids[(N2 - 3) / 3 - 1] = ob->aruco_id;                                  /* :256 */
            e->N = N2;
        }
        new_last_ids[qi] = ob->aruco_id;
        memcpy(new_last_obs + 3 * qi, lastobs, sizeof(lastobs));
    }
    free(e->last_ids); free(e->last_obs);
    e->last_ids = new_last_ids; e->last_obs = new_last_obs; e->n_last = n;        /* :263 */
    free(mu); free(q);
}
