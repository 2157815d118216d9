/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's CRF path.
 *
 * Plain C, scalar fp32, compiled with -ffp-contract=off so that every add is
 * the same IEEE operation torch performs.  Each function cites the lines of
 * /root/reference/flair/models/sequence_tagger_model.py it restates.  It is
 * pinned against tests/golden/crf_golden.npz (outputs of the reference's own
 * code, produced by oracle/make_golden.py).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may call it; the product
 * path (kb-ner_b200/) never does.
 *
 * Conventions: emissions emis[B][T][L] fp32, transitions
 * trans[to][from] fp32 (L x L), `pos` (optional, may be NULL) lists for every
 * sentence the original time indices of its kept (non S-X) tokens, klen[b] of
 * them; with pos == NULL the kept tokens are 0..klen[b]-1.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NEG (-1e12f)

static inline const float *row(const float *emis, const int32_t *pos, int b, int i, int T, int L) {
    int t = pos ? pos[(size_t)b * T + i] : i;
    return emis + ((size_t)b * T + t) * L;
}

/* remove-X compaction, sequence_tagger_model.py:2448-2488: keep = mask & (tag != S-X),
 * kept rows are left-packed in order. */
void kbner_oracle_compact(const uint8_t *keep, int B, int T, int32_t *pos, int32_t *klen) {
    for (int b = 0; b < B; ++b) {
        int n = 0;
        for (int t = 0; t < T; ++t)
            if (keep[(size_t)b * T + t]) pos[(size_t)b * T + n++] = t;
        klen[b] = n;
        for (int t = n; t < T; ++t) pos[(size_t)b * T + t] = -1;
    }
}

/* _viterbi_decode, :1248-1304 (+ confidence :1295-1300) and the S-X padding of
 * _obtain_labels, :1198-1208.  tags_out/conf_out are [B][T]; positions >= slen[b]
 * get -1 / 0, un-kept positions < slen[b] get x_idx / 1.0. */
void kbner_oracle_viterbi(const float *emis, const int32_t *pos, const int32_t *klen,
                          const int32_t *slen, const float *trans, int B, int T, int L,
                          int start, int stop, int x_idx, int32_t *tags_out, float *conf_out) {
    float *v = (float *)malloc(sizeof(float) * L);
    float *nv = (float *)malloc(sizeof(float) * L);
    uint8_t *bp = (uint8_t *)malloc((size_t)T * L);
    for (int b = 0; b < B; ++b) {
        int n = klen[b];
        int32_t *tg = tags_out + (size_t)b * T;
        float *cf = conf_out + (size_t)b * T;
        for (int t = 0; t < T; ++t) {
            if (t < slen[b]) { tg[t] = x_idx; cf[t] = 1.0f; }
            else             { tg[t] = -1;    cf[t] = 0.0f; }
        }
        if (n <= 0) continue;
        for (int k = 0; k < L; ++k) v[k] = NEG;
        v[start] = 0.0f;                                             /* :1252-1256 */
        for (int i = 0; i < n; ++i) {
            const float *e = row(emis, pos, b, i, T, L);
            for (int j = 0; j < L; ++j) {
                /* next_tag_var[j][k] = forward_var[k] + transitions[j][k]; torch.max -> first max (:1266-1269) */
                float best = v[0] + trans[j * L + 0];
                int bk = 0;
                for (int k = 1; k < L; ++k) {
                    float c = v[k] + trans[j * L + k];
                    if (c > best) { best = c; bk = k; }
                }
                bp[(size_t)i * L + j] = (uint8_t)bk;
                nv[j] = best + e[j];                                 /* :1270 */
            }
            memcpy(v, nv, sizeof(float) * L);
            /* confidence: softmax(backscore)[argmax backscore] (:1295-1300) */
            float m = v[0];
            for (int k = 1; k < L; ++k) if (v[k] > m) m = v[k];
            float s = 0.0f;
            for (int k = 0; k < L; ++k) s += expf(v[k] - m);
            int t = pos ? pos[(size_t)b * T + i] : i;
            cf[t] = 1.0f / s;
        }
        /* terminal (:1279-1287) */
        float best = 0.0f; int bj = -1;
        for (int k = 0; k < L; ++k) {
            float c = v[k] + trans[stop * L + k];
            if (k == stop || k == start) c = NEG;
            if (bj < 0 || c > best) { best = c; bj = k; }
        }
        /* back-trace (:1289-1304) */
        for (int i = n - 1; i >= 0; --i) {
            int t = pos ? pos[(size_t)b * T + i] : i;
            tg[t] = bj;
            bj = bp[(size_t)i * L + bj];
        }
        /* bj must now be START (assert :1303) -- exported for the tests */
        if (bj != start) tg[0] = -2;
    }
    free(v); free(nv); free(bp);
}

/* _forward_alg :1329-1394 (log Z) and _score_sentence :2544-2591 (gold path score).
 * alpha_out (optional) receives alpha_{i+1} for i < klen[b] as [B][T][L] (compacted index). */
void kbner_oracle_crf_nll(const float *emis, const int32_t *tags, const int32_t *pos,
                          const int32_t *klen, const float *trans, int B, int T, int L,
                          int start, int stop, float *logz, float *gold, float *alpha_out) {
    float *a = (float *)malloc(sizeof(float) * L);
    float *na = (float *)malloc(sizeof(float) * L);
    float *x = (float *)malloc(sizeof(float) * L);
    for (int b = 0; b < B; ++b) {
        int n = klen[b];
        for (int k = 0; k < L; ++k) a[k] = NEG;
        a[start] = 0.0f;                                             /* :1331-1340 */
        for (int i = 0; i < n; ++i) {
            const float *e = row(emis, pos, b, i, T, L);
            for (int j = 0; j < L; ++j) {
                /* tag_var = (emit[j] + trans[j][k]) + forward_var[k]   (:1355-1361) */
                float m = 0.0f;
                for (int k = 0; k < L; ++k) {
                    x[k] = (e[j] + trans[j * L + k]) + a[k];
                    if (k == 0 || x[k] > m) m = x[k];
                }
                float s = 0.0f;
                for (int k = 0; k < L; ++k) s += expf(x[k] - m);     /* :1363-1369 */
                na[j] = m + logf(s);
            }
            memcpy(a, na, sizeof(float) * L);
            if (alpha_out) memcpy(alpha_out + ((size_t)b * T + i) * L, a, sizeof(float) * L);
        }
        /* terminal: forward_var[len] + transitions[STOP]; log_sum_exp_batch (:1381-1392, :64-69) */
        float m = 0.0f;
        for (int k = 0; k < L; ++k) {
            x[k] = a[k] + trans[stop * L + k];
            if (k == 0 || x[k] > m) m = x[k];
        }
        float s = 0.0f;
        for (int k = 0; k < L; ++k) s += expf(x[k] - m);
        logz[b] = m + logf(s);

        /* gold: sum_t e_t[y_t] + A[y_0,START] + sum A[y_t,y_{t-1}] + A[STOP,y_last]  (:2544-2591) */
        float em = 0.0f, tr = 0.0f;
        int prev = start;
        for (int i = 0; i < n; ++i) {
            int t = pos ? pos[(size_t)b * T + i] : i;
            int y = tags[(size_t)b * T + t];
            em += row(emis, pos, b, i, T, L)[y];
            tr += trans[y * L + prev];
            prev = y;
        }
        tr += trans[stop * L + prev];
        gold[b] = tr + em;
    }
    free(a); free(na); free(x);
}

/* Gradient of  sum_b w[b] * (logZ_b - gold_b)  w.r.t. emissions and transitions, by the
 * forward-backward identities (what autograd derives from :1329-1394 / :2544-2591).
 * Double precision internally -- this is the checker for crf_nll_bwd, not a bit-exact port. */
void kbner_oracle_crf_nll_bwd(const float *emis, const int32_t *tags, const int32_t *pos,
                              const int32_t *klen, const float *trans, const float *w,
                              int B, int T, int L, int start, int stop,
                              float *d_emis, float *d_trans) {
    double *al = (double *)malloc(sizeof(double) * (size_t)(T + 1) * L);
    double *be = (double *)malloc(sizeof(double) * (size_t)(T + 1) * L);
    double *dt = (double *)calloc((size_t)L * L, sizeof(double));
    memset(d_emis, 0, sizeof(float) * (size_t)B * T * L);
    for (int b = 0; b < B; ++b) {
        int n = klen[b];
        for (int k = 0; k < L; ++k) al[k] = (k == start) ? 0.0 : -1e12;
        for (int i = 0; i < n; ++i) {
            const float *e = row(emis, pos, b, i, T, L);
            for (int j = 0; j < L; ++j) {
                double m = -INFINITY;
                for (int k = 0; k < L; ++k) {
                    double xv = (double)e[j] + (double)trans[j * L + k] + al[(size_t)i * L + k];
                    if (xv > m) m = xv;
                }
                double s = 0.0;
                for (int k = 0; k < L; ++k)
                    s += exp((double)e[j] + (double)trans[j * L + k] + al[(size_t)i * L + k] - m);
                al[(size_t)(i + 1) * L + j] = m + log(s);
            }
        }
        double m = -INFINITY, s = 0.0;
        for (int k = 0; k < L; ++k) { double xv = al[(size_t)n * L + k] + trans[stop * L + k]; if (xv > m) m = xv; }
        for (int k = 0; k < L; ++k) s += exp(al[(size_t)n * L + k] + trans[stop * L + k] - m);
        double lz = m + log(s);
        for (int k = 0; k < L; ++k) be[(size_t)n * L + k] = (double)trans[stop * L + k];
        for (int i = n - 1; i >= 0; --i) {
            const float *e = row(emis, pos, b, i, T, L);           /* emission of step i feeds alpha_{i+1} */
            for (int k = 0; k < L; ++k) {
                double mm = -INFINITY;
                for (int j = 0; j < L; ++j) {
                    double xv = (double)e[j] + (double)trans[j * L + k] + be[(size_t)(i + 1) * L + j];
                    if (xv > mm) mm = xv;
                }
                double ss = 0.0;
                for (int j = 0; j < L; ++j)
                    ss += exp((double)e[j] + (double)trans[j * L + k] + be[(size_t)(i + 1) * L + j] - mm);
                be[(size_t)i * L + k] = mm + log(ss);
            }
        }
        double wb = (double)w[b];
        int prev = start;
        for (int i = 0; i < n; ++i) {
            int t = pos ? pos[(size_t)b * T + i] : i;
            const float *e = row(emis, pos, b, i, T, L);
            int y = tags[(size_t)b * T + t];
            for (int j = 0; j < L; ++j) {
                double pj = exp(al[(size_t)(i + 1) * L + j] + be[(size_t)(i + 1) * L + j] - lz);
                d_emis[((size_t)b * T + t) * L + j] = (float)(wb * (pj - (j == y ? 1.0 : 0.0)));
                for (int k = 0; k < L; ++k) {
                    double pjk = exp(al[(size_t)i * L + k] + (double)trans[j * L + k] + (double)e[j] +
                                     be[(size_t)(i + 1) * L + j] - lz);
                    dt[j * L + k] += wb * pjk;
                }
            }
            dt[y * L + prev] -= wb;
            prev = y;
        }
        for (int k = 0; k < L; ++k)
            dt[stop * L + k] += wb * exp(al[(size_t)n * L + k] + (double)trans[stop * L + k] - lz);
        dt[stop * L + prev] -= wb;
    }
    for (int i = 0; i < L * L; ++i) d_trans[i] = (float)dt[i];
    free(al); free(be); free(dt);
}
