/*
 * oracle/oracle.c — CPU ORACLE for the sbv hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (libsbv.so) never links or calls it.
 *
 * What it restates
 * ----------------
 * The reference (SmartBFT-Go/consensus) contains NO signature arithmetic: api.Verifier
 * (pkg/api/dependencies.go:54-71) is implemented by the embedding application, and a real
 * application implements it with the Go standard library (crypto/ecdsa, crypto/sha256,
 * encoding/asn1; toolchain pinned by go.mod:3 "go 1.20", CI go 1.21.8).  The Go stdlib is a
 * third-party dependency ABSENT from /root/reference and from this image, so this file
 * restates its published algorithm (FIPS 186-4 §6.4 ECDSA verification with the
 * crypto/ecdsa accept set) on top of OpenSSL 3.0 libcrypto BIGNUM / EC_POINT primitives:
 *
 *   reject if Q is not an on-curve affine point with 0 <= x,y < p
 *   reject if r or s is 0 or >= n                       (no low-S rule: high-S accepts)
 *   e  = leftmost min(len, ceil(log2 n / 8)) digest bytes as a big-endian integer
 *   w  = s^-1 mod n ; u1 = e*w mod n ; u2 = r*w mod n
 *   R  = u1*G + u2*Q ; reject if R = infinity ; accept iff R.x mod n == r
 *
 * PARITY PINNING: the reference holds no ECDSA / SHA-256 golden vector ("parity unpinned" at
 * the reference level).  The oracle is pinned instead against RFC 6979 A.2.5 /
 * A.2.6 known-answer signatures, FIPS 180-4 SHA-256 vectors, OpenSSL's own ECDSA_do_verify,
 * the independent pure-Python restatement in oracle/ecdsa_ref.py and python `cryptography`
 * (tests/test_oracle.py).
 *
 * SHA-256 call sites restated: pkg/types/types.go:64-69, internal/bft/util.go:583-585.
 */
#define OPENSSL_SUPPRESS_DEPRECATED 1
#include <openssl/bn.h>
#include <openssl/ec.h>
#include <openssl/ecdsa.h>
#include <openssl/obj_mac.h>
#include <openssl/sha.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

enum { ORC_P256 = 0, ORC_P384 = 1 };

static int curve_nid(int curve) { return curve == ORC_P256 ? NID_X9_62_prime256v1 : NID_secp384r1; }
static size_t curve_len(int curve) { return curve == ORC_P256 ? 32 : 48; }

typedef struct {
    EC_GROUP *g;
    BIGNUM *n, *p;
    BN_CTX *ctx;
    size_t len;
} orc_ctx;

static int ctx_init(orc_ctx *c, int curve) {
    c->g = EC_GROUP_new_by_curve_name(curve_nid(curve));
    if (!c->g) return -1;
    c->ctx = BN_CTX_new();
    c->n = BN_new();
    c->p = BN_new();
    EC_GROUP_get_order(c->g, c->n, c->ctx);
    EC_GROUP_get_curve(c->g, c->p, NULL, NULL, c->ctx);
    c->len = curve_len(curve);
    return 0;
}
static void ctx_free(orc_ctx *c) {
    BN_free(c->n); BN_free(c->p); BN_CTX_free(c->ctx); EC_GROUP_free(c->g);
}

/* One verification, Go crypto/ecdsa semantics (see header).  Returns 1 accept / 0 reject. */
static int verify_one(orc_ctx *c, const uint8_t *r_be, const uint8_t *s_be, const uint8_t *qx_be,
                      const uint8_t *qy_be, const uint8_t *dig, size_t dlen) {
    int ok = 0;
    BN_CTX_start(c->ctx);
    BIGNUM *r = BN_CTX_get(c->ctx), *s = BN_CTX_get(c->ctx), *x = BN_CTX_get(c->ctx),
           *y = BN_CTX_get(c->ctx), *e = BN_CTX_get(c->ctx), *w = BN_CTX_get(c->ctx),
           *u1 = BN_CTX_get(c->ctx), *u2 = BN_CTX_get(c->ctx), *rx = BN_CTX_get(c->ctx);
    EC_POINT *Q = EC_POINT_new(c->g), *R = EC_POINT_new(c->g);
    BN_bin2bn(r_be, (int)c->len, r);
    BN_bin2bn(s_be, (int)c->len, s);
    BN_bin2bn(qx_be, (int)c->len, x);
    BN_bin2bn(qy_be, (int)c->len, y);
    if (BN_is_zero(r) || BN_is_zero(s) || BN_cmp(r, c->n) >= 0 || BN_cmp(s, c->n) >= 0) goto done;
    if (BN_cmp(x, c->p) >= 0 || BN_cmp(y, c->p) >= 0) goto done;
    /* set_affine_coordinates checks the curve equation; (0,0) is off-curve since b != 0 */
    if (EC_POINT_set_affine_coordinates(c->g, Q, x, y, c->ctx) != 1) goto done;
    if (dlen > c->len) dlen = c->len; /* leftmost bytes; order bit length is a byte multiple */
    BN_bin2bn(dig, (int)dlen, e);
    if (!BN_mod_inverse(w, s, c->n, c->ctx)) goto done;
    BN_mod_mul(u1, e, w, c->n, c->ctx);
    BN_mod_mul(u2, r, w, c->n, c->ctx);
    if (EC_POINT_mul(c->g, R, u1, Q, u2, c->ctx) != 1) goto done;
    if (EC_POINT_is_at_infinity(c->g, R)) goto done;
    if (EC_POINT_get_affine_coordinates(c->g, R, rx, NULL, c->ctx) != 1) goto done;
    BN_nnmod(rx, rx, c->n, c->ctx);
    ok = BN_cmp(rx, r) == 0;
done:
    EC_POINT_free(Q); EC_POINT_free(R);
    BN_CTX_end(c->ctx);
    return ok;
}

/* Strict DER SEQUENCE{INTEGER r, INTEGER s}: minimal lengths, minimal non-negative integers,
 * no trailing bytes (crypto/ecdsa.VerifyASN1 / cryptobyte rules).  Writes r,s left-padded to
 * `len` bytes.  Returns 1 ok, 0 malformed (also when an integer does not fit `len` bytes —
 * such a value is >= n and rejects anyway). */
static int der_int(const uint8_t **pp, const uint8_t *end, uint8_t *out, size_t len) {
    const uint8_t *p = *pp;
    if (end - p < 2 || p[0] != 0x02) return 0;
    size_t l = p[1];
    p += 2;
    if (l & 0x80) return 0; /* integers here are < 128 bytes: long form is non-minimal */
    if (l == 0 || (size_t)(end - p) < l) return 0;
    if (p[0] & 0x80) return 0;                              /* negative */
    if (l > 1 && p[0] == 0x00 && !(p[1] & 0x80)) return 0; /* non-minimal */
    const uint8_t *v = p;
    size_t vl = l;
    if (v[0] == 0x00 && vl > 1) { v++; vl--; }
    if (vl > len) return 0;
    memset(out, 0, len);
    memcpy(out + (len - vl), v, vl);
    *pp = p + l;
    return 1;
}
int orc_der_parse(const uint8_t *sig, size_t siglen, size_t len, uint8_t *r, uint8_t *s) {
    const uint8_t *p = sig, *end = sig + siglen;
    if (siglen < 2 || p[0] != 0x30) return 0;
    size_t l;
    if (p[1] < 0x80) { l = p[1]; p += 2; }
    else if (p[1] == 0x81) { if (siglen < 3 || p[2] < 0x80) return 0; l = p[2]; p += 3; }
    else return 0;
    if ((size_t)(end - p) != l) return 0; /* trailing bytes or truncated */
    if (!der_int(&p, end, r, len)) return 0;
    if (!der_int(&p, end, s, len)) return 0;
    return p == end;
}

typedef struct {
    int curve; size_t lo, hi;
    const uint8_t *r, *s, *qx, *qy, *dig; size_t dlen;
    const uint8_t *sigs; const uint32_t *sig_off; const uint8_t *qxy;
    uint8_t *ok;
} vjob;

static void *verify_worker(void *arg) {
    vjob *j = (vjob *)arg;
    orc_ctx c;
    if (ctx_init(&c, j->curve)) return NULL;
    size_t L = c.len;
    for (size_t i = j->lo; i < j->hi; i++) {
        if (j->sigs) {
            uint8_t r[48], s[48];
            const uint8_t *sg = j->sigs + j->sig_off[i];
            size_t sl = j->sig_off[i + 1] - j->sig_off[i];
            if (!orc_der_parse(sg, sl, L, r, s)) { j->ok[i] = 0; continue; }
            j->ok[i] = (uint8_t)verify_one(&c, r, s, j->qxy + 2 * L * i, j->qxy + 2 * L * i + L,
                                           j->dig + j->dlen * i, j->dlen);
        } else {
            j->ok[i] = (uint8_t)verify_one(&c, j->r + L * i, j->s + L * i, j->qx + L * i,
                                           j->qy + L * i, j->dig + j->dlen * i, j->dlen);
        }
    }
    ctx_free(&c);
    return NULL;
}

static void run_jobs(vjob *tmpl, size_t n, int nthreads, void *(*fn)(void *)) {
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
    vjob *jobs = malloc(sizeof(vjob) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = *tmpl;
        jobs[t].lo = n * t / nthreads;
        jobs[t].hi = n * (t + 1) / nthreads;
        pthread_create(&th[t], NULL, fn, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

/* SoA fixed-width big-endian batch, same layout as sbv_verify_batch (include/sbv.h). */
int orc_verify_batch(int curve, size_t n, const uint8_t *r, const uint8_t *s, const uint8_t *qx,
                     const uint8_t *qy, const uint8_t *digest, size_t dlen, uint8_t *ok,
                     int nthreads) {
    vjob j; memset(&j, 0, sizeof j);
    j.curve = curve; j.r = r; j.s = s; j.qx = qx; j.qy = qy; j.dig = digest; j.dlen = dlen; j.ok = ok;
    run_jobs(&j, n, nthreads, verify_worker);
    return 0;
}

/* DER front end, same layout as sbv_verify_batch_der. */
int orc_verify_batch_der(int curve, size_t n, const uint8_t *sigs, const uint32_t *sig_off,
                         const uint8_t *qxy, const uint8_t *digest, size_t dlen, uint8_t *ok,
                         int nthreads) {
    vjob j; memset(&j, 0, sizeof j);
    j.curve = curve; j.sigs = sigs; j.sig_off = sig_off; j.qxy = qxy; j.dig = digest; j.dlen = dlen; j.ok = ok;
    run_jobs(&j, n, nthreads, verify_worker);
    return 0;
}

/* ---- key generation and deterministic signing (corpus generator; not on the verify path) ---- */
int orc_pubkey(int curve, const uint8_t *d_be, uint8_t *qx, uint8_t *qy) {
    orc_ctx c; if (ctx_init(&c, curve)) return -1;
    BIGNUM *d = BN_bin2bn(d_be, (int)c.len, NULL), *x = BN_new(), *y = BN_new();
    EC_POINT *Q = EC_POINT_new(c.g);
    int rc = -1;
    if (!BN_is_zero(d) && BN_cmp(d, c.n) < 0 && EC_POINT_mul(c.g, Q, d, NULL, NULL, c.ctx) == 1 &&
        EC_POINT_get_affine_coordinates(c.g, Q, x, y, c.ctx) == 1) {
        BN_bn2binpad(x, qx, (int)c.len); BN_bn2binpad(y, qy, (int)c.len); rc = 0;
    }
    EC_POINT_free(Q); BN_free(d); BN_free(x); BN_free(y); ctx_free(&c);
    return rc;
}

/* r = (k*G).x mod n ; s = k^-1 (e + r d) mod n.  rc 0 ok, 1 degenerate (r or s zero / k out of range). */
int orc_sign_batch(int curve, size_t n, const uint8_t *d_be, const uint32_t *key_idx,
                   const uint8_t *digest, size_t dlen, const uint8_t *k_be, uint8_t *r_out,
                   uint8_t *s_out) {
    orc_ctx c; if (ctx_init(&c, curve)) return -1;
    size_t L = c.len; int rc = 0;
    BIGNUM *d = BN_new(), *k = BN_new(), *e = BN_new(), *r = BN_new(), *s = BN_new(), *t = BN_new();
    EC_POINT *R = EC_POINT_new(c.g);
    for (size_t i = 0; i < n; i++) {
        BN_bin2bn(d_be + L * key_idx[i], (int)L, d);
        BN_bin2bn(k_be + L * i, (int)L, k);
        BN_nnmod(k, k, c.n, c.ctx);
        if (BN_is_zero(k)) BN_one(k);
        size_t dl = dlen > L ? L : dlen;
        BN_bin2bn(digest + dlen * i, (int)dl, e);
        EC_POINT_mul(c.g, R, k, NULL, NULL, c.ctx);
        EC_POINT_get_affine_coordinates(c.g, R, r, NULL, c.ctx);
        BN_nnmod(r, r, c.n, c.ctx);
        BN_mod_mul(t, r, d, c.n, c.ctx);
        BN_mod_add(t, t, e, c.n, c.ctx);
        BN_mod_inverse(s, k, c.n, c.ctx);
        BN_mod_mul(s, s, t, c.n, c.ctx);
        if (BN_is_zero(r) || BN_is_zero(s)) rc = 1;
        BN_bn2binpad(r, r_out + L * i, (int)L);
        BN_bn2binpad(s, s_out + L * i, (int)L);
    }
    EC_POINT_free(R); BN_free(d); BN_free(k); BN_free(e); BN_free(r); BN_free(s); BN_free(t);
    ctx_free(&c);
    return rc;
}

/* general a*G + b*Q in affine big-endian; rc 1 if the result is infinity (test-vector crafting) */
int orc_lincomb(int curve, const uint8_t *a_be, const uint8_t *b_be, const uint8_t *qx,
                const uint8_t *qy, uint8_t *ox, uint8_t *oy) {
    orc_ctx c; if (ctx_init(&c, curve)) return -1;
    size_t L = c.len; int rc = -1;
    BIGNUM *a = BN_bin2bn(a_be, (int)L, NULL), *b = BN_bin2bn(b_be, (int)L, NULL);
    BIGNUM *x = BN_bin2bn(qx, (int)L, NULL), *y = BN_bin2bn(qy, (int)L, NULL);
    EC_POINT *Q = EC_POINT_new(c.g), *R = EC_POINT_new(c.g);
    if (EC_POINT_set_affine_coordinates(c.g, Q, x, y, c.ctx) == 1 &&
        EC_POINT_mul(c.g, R, a, Q, b, c.ctx) == 1) {
        if (EC_POINT_is_at_infinity(c.g, R)) rc = 1;
        else if (EC_POINT_get_affine_coordinates(c.g, R, x, y, c.ctx) == 1) {
            BN_bn2binpad(x, ox, (int)L); BN_bn2binpad(y, oy, (int)L); rc = 0;
        }
    }
    EC_POINT_free(Q); EC_POINT_free(R); BN_free(a); BN_free(b); BN_free(x); BN_free(y); ctx_free(&c);
    return rc;
}

/* ---- SHA-256 over a ragged batch: msgs concatenated, off[n+1] (types.go:64-69 restated) ---- */
typedef struct { size_t lo, hi; const uint8_t *msgs; const uint64_t *off; uint8_t *out; } hjob;
static void *hash_worker(void *arg) {
    hjob *j = (hjob *)arg;
    for (size_t i = j->lo; i < j->hi; i++)
        SHA256(j->msgs + j->off[i], j->off[i + 1] - j->off[i], j->out + 32 * i);
    return NULL;
}
int orc_sha256_batch(size_t n, const uint8_t *msgs, const uint64_t *off, uint8_t *out, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    pthread_t th[256]; hjob jobs[256];
    if (nthreads > 256) nthreads = 256;
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (hjob){n * t / nthreads, n * (t + 1) / nthreads, msgs, off, out};
        pthread_create(&th[t], NULL, hash_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}

/* ---- CPU baseline timing: OpenSSL's production verifier (ECDSA_do_verify, ecp_nistz256 for
 * P-256) on pre-built EC_KEYs — the stand-in for "the reference's Go Verifier" (BASELINE.md §2).
 * keys: K affine keys; key_idx[i] selects the key of item i.  Times the verify loop only.
 * Returns seconds (max over threads' common wall window); verdicts in ok. ---- */
typedef struct {
    int curve; size_t lo, hi, K;
    const uint8_t *r, *s, *keys, *dig; const uint32_t *key_idx; size_t dlen; uint8_t *ok;
    pthread_barrier_t *bar; double t0, t1;
} bjob;
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static void *bench_worker(void *arg) {
    bjob *j = (bjob *)arg;
    size_t L = curve_len(j->curve);
    EC_KEY **keys = calloc(j->K, sizeof(EC_KEY *));
    for (size_t k = 0; k < j->K; k++) {
        BIGNUM *x = BN_bin2bn(j->keys + 2 * L * k, (int)L, NULL), *y = BN_bin2bn(j->keys + 2 * L * k + L, (int)L, NULL);
        EC_KEY *ek = EC_KEY_new_by_curve_name(curve_nid(j->curve));
        if (EC_KEY_set_public_key_affine_coordinates(ek, x, y) != 1) { EC_KEY_free(ek); ek = NULL; }
        keys[k] = ek; BN_free(x); BN_free(y);
    }
    ECDSA_SIG *sig = ECDSA_SIG_new();
    pthread_barrier_wait(j->bar);
    j->t0 = now_s();
    for (size_t i = j->lo; i < j->hi; i++) {
        EC_KEY *ek = keys[j->key_idx[i]];
        if (!ek) { j->ok[i] = 0; continue; }
        BIGNUM *r = BN_bin2bn(j->r + L * i, (int)L, NULL), *s = BN_bin2bn(j->s + L * i, (int)L, NULL);
        ECDSA_SIG_set0(sig, r, s); /* frees previous r,s */
        size_t dl = j->dlen > L ? L : j->dlen;
        j->ok[i] = ECDSA_do_verify(j->dig + j->dlen * i, (int)dl, sig, ek) == 1;
    }
    j->t1 = now_s();
    ECDSA_SIG_free(sig);
    for (size_t k = 0; k < j->K; k++) EC_KEY_free(keys[k]);
    free(keys);
    return NULL;
}
double orc_bench_verify(int curve, size_t n, const uint8_t *r, const uint8_t *s, const uint8_t *keys,
                        size_t K, const uint32_t *key_idx, const uint8_t *digest, size_t dlen,
                        uint8_t *ok, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256]; bjob jobs[256]; pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (bjob){curve, n * t / nthreads, n * (t + 1) / nthreads, K, r, s, keys, digest, key_idx, dlen, ok, &bar, 0, 0};
        pthread_create(&th[t], NULL, bench_worker, &jobs[t]);
    }
    double t0 = 1e300, t1 = 0;
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].t0 < t0) t0 = jobs[t].t0;
        if (jobs[t].t1 > t1) t1 = jobs[t].t1;
    }
    pthread_barrier_destroy(&bar);
    return t1 - t0;
}

/* SHA-256 + verify fused CPU baseline (config 3): hash each message then verify as above. */
int orc_version(void) { return 1; }
