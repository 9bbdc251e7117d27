/*
 * CPU restatement (plain C) of the reference's pose parser.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may load the shared
 * object built from this file (oracle/Makefile -> oracle/libppn_oracle.so).  The product never
 * links or calls it.
 *
 * Parity status: PINNED — tests/test_oracle_golden.py checks this file against
 * tests/golden/*.npz, which hold outputs of the reference's own functions run by
 * oracle/make_golden.py (see oracle/ppn_oracle.py for the numpy twin of this file).
 *
 * Reference lines restated (all under /root/reference):
 *   delta = resp*conf ............ rt_test.py:130
 *   cell box ..................... datatest.py:63-71, 80-85
 *   root candidates .............. datatest.py:86-92
 *   greedy IoU NMS ............... datatest.py:134-160
 *   window arg-max ............... datatest.py:100, 113   (numpy argmax: first maximum, NaN is maximal)
 *   track-order walk ............. datatest.py:103-131
 *
 * Build with -ffp-contract=off: every operation below must round once, as numpy's do.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t K, E, inW, inH, W, H, sW, sH;
    int32_t off_h, off_w;            /* datatest.py:115-116 */
    int32_t n_chains;
    const int32_t* chain_off;        /* [n_chains+1] */
    const int32_t* chain_limb;       /* [chain_off[n_chains]]   config.py:75-80 "eis" */
    const int32_t* chain_part;       /* [chain_off[n_chains]]   config.py:75-80 "ts"  */
    float det_thr, nms_thr;
    int32_t min_kp;
} ppn_oracle_geom;

static inline size_t plane(const ppn_oracle_geom* g, int group, int k) {
    return ((size_t)group * g->K + k) * (size_t)(g->H * g->W);
}

/* delta of part k at flat cell c */
static inline float delta_at(const float* img, const ppn_oracle_geom* g, int k, int c) {
    return img[plane(g, 0, k) + c] * img[plane(g, 1, k) + c];
}

/* (ymin, xmin, ymax, xmax) of part k at flat cell c */
static void box_at(const float* img, const ppn_oracle_geom* g, int k, int c, float out[4]) {
    const int h = c / g->W, w = c % g->W;
    const float gridW = (float)(int)((double)g->inW / g->W), gridH = (float)(int)((double)g->inH / g->H);
    float rx = img[plane(g, 2, k) + c] + (float)w;  rx = rx * gridW;
    float ry = img[plane(g, 3, k) + c] + (float)h;  ry = ry * gridH;
    const float rw = (float)g->inW * img[plane(g, 4, k) + c];
    const float rh = (float)g->inH * img[plane(g, 5, k) + c];
    const float hw = rw * 0.5f, hh = rh * 0.5f;
    out[0] = ry - hh; out[1] = rx - hw; out[2] = ry + hh; out[3] = rx + hw;
}

/* numpy argmax over n floats with stride: first maximum; a NaN, once met, wins and stops the scan */
static int argmax_strided(const float* p, int n, size_t stride) {
    float best = p[0];
    int idx = 0;
    if (isnan(best)) return 0;
    for (int a = 1; a < n; ++a) {
        const float v = p[(size_t)a * stride];
        if (!(v <= best)) {
            best = v; idx = a;
            if (isnan(v)) break;
        }
    }
    return idx;
}

/* on-demand form, exactly as the reference takes it: window of limb ei at flat cell c (datatest.py:113) */
int ppn_oracle_window_argmax(const float* img, const ppn_oracle_geom* g, int ei, int c) {
    const size_t HW = (size_t)g->H * g->W, S = (size_t)g->sH * g->sW;
    return argmax_strided(img + (6 * (size_t)g->K + (size_t)ei * S) * HW + c, (int)S, HW);
}

/* dense window arg-max of one image: amax[E, H*W] */
int ppn_oracle_limb_argmax(const float* img, const ppn_oracle_geom* g, int32_t* amax) {
    const int HW = g->H * g->W, S = g->sH * g->sW;
    const float* e = img + (size_t)6 * g->K * HW;
    for (int ei = 0; ei < g->E; ++ei) {
        const float* m = e + (size_t)ei * S * HW;
        /* row-wise sweep keeps the inner loop contiguous; strict '>' keeps the first maximum */
        float* best = (float*)malloc(sizeof(float) * HW);
        uint8_t* done = (uint8_t*)calloc(HW, 1);
        if (!best || !done) { free(best); free(done); return -1; }
        for (int c = 0; c < HW; ++c) { best[c] = m[c]; amax[(size_t)ei * HW + c] = 0; done[c] = isnan(m[c]); }
        for (int a = 1; a < S; ++a) {
            const float* row = m + (size_t)a * HW;
            for (int c = 0; c < HW; ++c) {
                const float v = row[c];
                if (!done[c] && !(v <= best[c])) {
                    best[c] = v; amax[(size_t)ei * HW + c] = a;
                    if (isnan(v)) done[c] = 1;
                }
            }
        }
        free(best); free(done);
    }
    return 0;
}

/* numpy's maximum/minimum propagate NaN (fmaxf/fminf would drop it) */
static inline float np_max(float a, float b) { return (a >= b || a != a) ? a : b; }
static inline float np_min(float a, float b) { return (a <= b || a != a) ? a : b; }

/* order[] = indices sorted by score descending, exact ties -> larger index first */
typedef struct { float s; int32_t i; } scored;
static int by_score_desc(const void* a, const void* b) {
    const scored* x = (const scored*)a; const scored* y = (const scored*)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->i > y->i) ? -1 : (x->i < y->i);
}

/* datatest.py:134-160.  box [n,4]; score may be NULL (input order); limit <= 0 means none.
 * keep[] receives indices into box in visiting order; returns the number kept (or -1). */
int ppn_oracle_nms(const float* box, const float* score, int n, float thr, int limit, int32_t* keep) {
    if (n <= 0) return 0;
    scored* ord = (scored*)malloc(sizeof(scored) * n);
    float* area = (float*)malloc(sizeof(float) * n);
    if (!ord || !area) { free(ord); free(area); return -1; }
    for (int i = 0; i < n; ++i) { ord[i].i = i; ord[i].s = score ? score[i] : 0.0f; }
    if (score) qsort(ord, n, sizeof(scored), by_score_desc);
    for (int i = 0; i < n; ++i) {
        const float* b = box + 4 * (size_t)ord[i].i;
        const float dy = b[2] - b[0], dx = b[3] - b[1];
        area[i] = dy * dx;
    }
    int m = 0;
    int32_t* kept_pos = (int32_t*)malloc(sizeof(int32_t) * n);   /* positions in sorted order */
    if (!kept_pos) { free(ord); free(area); return -1; }
    for (int i = 0; i < n; ++i) {
        const float* b = box + 4 * (size_t)ord[i].i;
        int drop = 0;
        for (int q = 0; q < m && !drop; ++q) {
            const int j = kept_pos[q];
            const float* o = box + 4 * (size_t)ord[j].i;
            const float tly = np_max(b[0], o[0]), tlx = np_max(b[1], o[1]);
            const float bry = np_min(b[2], o[2]), brx = np_min(b[3], o[3]);
            const float dy = bry - tly, dx = brx - tlx;
            const float prod = dy * dx;
            const float inter = prod * ((tly < bry && tlx < brx) ? 1.0f : 0.0f);
            const float sum = area[i] + area[j];
            const float iou = inter / (sum - inter);
            if (iou >= thr) drop = 1;
        }
        if (drop) continue;
        kept_pos[m] = i;
        keep[m] = ord[i].i;
        ++m;
        if (limit > 0 && m >= limit) break;
    }
    free(kept_pos); free(ord); free(area);
    return m;
}

/*
 * Whole path for one image [C,H,W].  Output arrays are sized for H*W roots:
 *   cand_cell[HW], keep_idx[HW], root_cell[HW], part_cell[HW*K], part_score[HW*K], part_box[HW*K*4];
 *   counts[3] = {n_cand, n_keep, n_humans};  amax_out [E*HW] may be NULL.
 */
int ppn_oracle_parse_image(const float* img, const ppn_oracle_geom* g,
                           int32_t* counts, int32_t* cand_cell, int32_t* keep_idx,
                           int32_t* root_cell, int32_t* part_cell, float* part_score, float* part_box,
                           int32_t* amax_out) {
    const int HW = g->H * g->W, K = g->K;
    int n = 0;
    float* cbox = (float*)malloc(sizeof(float) * 4 * HW);
    float* cscore = (float*)malloc(sizeof(float) * HW);
    int32_t* amax = amax_out ? amax_out : (int32_t*)malloc(sizeof(int32_t) * (size_t)g->E * HW);
    if (!cbox || !cscore || !amax) return -1;
    for (int c = 0; c < HW; ++c) {
        const float d = delta_at(img, g, 0, c);
        if (d > g->det_thr) { cand_cell[n] = c; cscore[n] = d; box_at(img, g, 0, c, cbox + 4 * n); ++n; }
    }
    const int m = ppn_oracle_nms(cbox, cscore, n, g->nms_thr, 0, keep_idx);
    if (m < 0) return -1;
    if (ppn_oracle_limb_argmax(img, g, amax)) return -1;
    int nh = 0;
    for (int r = 0; r < m; ++r) {
        const int root = cand_cell[keep_idx[r]];
        int32_t* pc = part_cell + (size_t)nh * K;
        for (int t = 0; t < K; ++t) pc[t] = -1;
        pc[0] = root;
        for (int ch = 0; ch < g->n_chains; ++ch) {
            int ih = root / g->W, iw = root % g->W;
            for (int q = g->chain_off[ch]; q < g->chain_off[ch + 1]; ++q) {
                const int ei = g->chain_limb[q], t = g->chain_part[q];
                const int a = amax[(size_t)ei * HW + ih * g->W + iw];
                const int jh = ih + a / g->sW - g->off_h, jw = iw + a % g->sW - g->off_w;
                if (jh < 0 || jw < 0 || jh >= g->H || jw >= g->W) break;
                if (delta_at(img, g, t, jh * g->W + jw) < g->det_thr) break;
                pc[t] = jh * g->W + jw;
                ih = jh; iw = jw;
            }
        }
        int present = 0;
        for (int t = 1; t < K; ++t) present += (pc[t] >= 0);
        if (g->min_kp > present) continue;
        root_cell[nh] = root;
        for (int t = 0; t < K; ++t) {
            float* bx = part_box + ((size_t)nh * K + t) * 4;
            if (pc[t] >= 0) { part_score[(size_t)nh * K + t] = delta_at(img, g, t, pc[t]); box_at(img, g, t, pc[t], bx); }
            else { part_score[(size_t)nh * K + t] = 0.0f; bx[0] = bx[1] = bx[2] = bx[3] = 0.0f; }
        }
        ++nh;
    }
    counts[0] = n; counts[1] = m; counts[2] = nh;
    free(cbox); free(cscore);
    if (!amax_out) free(amax);
    return 0;
}

/* ---- batch driver: images are independent, so threads take contiguous blocks of them ---- */
typedef struct {
    const float* head; const ppn_oracle_geom* g; int b0, b1;
    int32_t *counts, *cand_cell, *keep_idx, *root_cell, *part_cell; float *part_score, *part_box;
    int rc;
} job;

static void* run_job(void* p) {
    job* j = (job*)p;
    const ppn_oracle_geom* g = j->g;
    const size_t HW = (size_t)g->H * g->W, K = g->K;
    const size_t per = (size_t)(6 * g->K + g->sH * g->sW * g->E) * HW;
    j->rc = 0;
    for (int b = j->b0; b < j->b1; ++b) {
        j->rc |= ppn_oracle_parse_image(j->head + per * b, g, j->counts + 3 * (size_t)b,
                                        j->cand_cell + HW * b, j->keep_idx + HW * b, j->root_cell + HW * b,
                                        j->part_cell + HW * K * b, j->part_score + HW * K * b,
                                        j->part_box + HW * K * 4 * b, NULL);
    }
    return NULL;
}

int ppn_oracle_parse_batch(const float* head, int B, const ppn_oracle_geom* g, int n_threads,
                           int32_t* counts, int32_t* cand_cell, int32_t* keep_idx, int32_t* root_cell,
                           int32_t* part_cell, float* part_score, float* part_box) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > B) n_threads = B > 0 ? B : 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
    job* jobs = (job*)malloc(sizeof(job) * n_threads);
    if (!th || !jobs) return -1;
    const int per = (B + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        job j = { head, g, t * per, (t + 1) * per < B ? (t + 1) * per : B,
                  counts, cand_cell, keep_idx, root_cell, part_cell, part_score, part_box, 0 };
        if (j.b0 > B) j.b0 = B;
        jobs[t] = j;
        pthread_create(&th[t], NULL, run_job, &jobs[t]);
    }
    int rc = 0;
    for (int t = 0; t < n_threads; ++t) { pthread_join(th[t], NULL); rc |= jobs[t].rc; }
    free(th); free(jobs);
    return rc;
}
