/*
 * oracle/densecrf.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Single-threaded plain-C restatement of the dense-CRF mean-field inference that the
 * reference reaches through pydensecrf (reference call sites:
 * PnP_OVSS_0514_updated_segmentation.py:1030-1074, twin _coco.py:1213-1257).
 *
 * pydensecrf (lucasb-eyer/pydensecrf, un-pinned "git clone master" in the reference README:28-30,
 * wrapping Kraehenbuehl & Koltun's densecrf) is NOT vendored in /root/reference, is not
 * installed and cannot be built here (no Eigen, no network).  This file restates its published
 * algorithm (permutohedral lattice of Adams et al. 2010: init / splat / blur / slice; DenseKernel
 * with DIAG_KERNEL + NORMALIZE_SYMMETRIC; Potts compatibility; DenseCRF::inference).
 *
 * PARITY UNPINNED: no pydensecrf binary or golden vector is available to diff against; this
 * restatement is validated by property tests only (tests/test_oracle_crf.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file.  The product (pnp_ovss_b200) never links or calls it.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Hash table: lattice key (d shorts) -> dense vertex index in insertion order.
 * Open addressing, linear probing, hash = (sum-accumulate + key[k]) * 1664525 per coordinate.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int key_size;
    int filled;
    int capacity;
    short *keys;  /* [filled][key_size] */
    int *table;   /* [capacity] -> index or -1 */
} HashTable;

static size_t ht_hash(const HashTable *h, const short *k) {
    size_t r = 0;
    for (int i = 0; i < h->key_size; i++) {
        r += (size_t)(long)k[i];
        r *= 1664525u;
    }
    return r;
}

static void ht_init(HashTable *h, int key_size, int n_elements) {
    h->key_size = key_size;
    h->filled = 0;
    h->capacity = 2 * (n_elements > 8 ? n_elements : 8);
    h->keys = (short *)malloc(sizeof(short) * (size_t)(h->capacity / 2 + 10) * key_size);
    h->table = (int *)malloc(sizeof(int) * (size_t)h->capacity);
    for (int i = 0; i < h->capacity; i++) h->table[i] = -1;
}

static void ht_free(HashTable *h) {
    free(h->keys);
    free(h->table);
}

static void ht_grow(HashTable *h) {
    int old_capacity = h->capacity;
    h->capacity *= 2;
    h->keys = (short *)realloc(h->keys, sizeof(short) * (size_t)(old_capacity + 10) * h->key_size);
    free(h->table);
    h->table = (int *)malloc(sizeof(int) * (size_t)h->capacity);
    for (int i = 0; i < h->capacity; i++) h->table[i] = -1;
    for (int i = 0; i < h->filled; i++) {
        size_t s = ht_hash(h, h->keys + (size_t)i * h->key_size) % (size_t)h->capacity;
        while (h->table[s] >= 0) {
            s++;
            if (s == (size_t)h->capacity) s = 0;
        }
        h->table[s] = i;
    }
}

static int ht_find(HashTable *h, const short *k, int create) {
    if (2 * h->filled >= h->capacity) ht_grow(h);
    size_t s = ht_hash(h, k) % (size_t)h->capacity;
    for (;;) {
        int e = h->table[s];
        if (e == -1) {
            if (create) {
                for (int i = 0; i < h->key_size; i++) h->keys[(size_t)h->filled * h->key_size + i] = k[i];
                h->table[s] = h->filled;
                return h->filled++;
            }
            return -1;
        }
        int good = 1;
        for (int i = 0; i < h->key_size && good; i++)
            if (h->keys[(size_t)e * h->key_size + i] != k[i]) good = 0;
        if (good) return e;
        s++;
        if (s == (size_t)h->capacity) s = 0;
    }
}

/* ------------------------------------------------------------------------------------------
 * Permutohedral lattice
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int N, d, M;
    int *offset;        /* [N][d+1] vertex index (no +1 shift) */
    float *barycentric; /* [N][d+1] */
    int *n1;            /* [d+1][M] neighbour index or -1 */
    int *n2;            /* [d+1][M] */
    short *keys;        /* [M][d] lattice keys, insertion order */
} Lattice;

/* feature: [N][d] (pixel-major, i.e. column k of densecrf's d x N Eigen matrix is feature+k*d) */
Lattice *pl_create(const float *feature, int N, int d) {
    Lattice *L = (Lattice *)calloc(1, sizeof(Lattice));
    L->N = N;
    L->d = d;
    HashTable ht;
    ht_init(&ht, d, N);

    L->offset = (int *)malloc(sizeof(int) * (size_t)N * (d + 1));
    L->barycentric = (float *)malloc(sizeof(float) * (size_t)N * (d + 1));

    float *scale_factor = (float *)malloc(sizeof(float) * d);
    float *elevated = (float *)malloc(sizeof(float) * (d + 1));
    float *rem0 = (float *)malloc(sizeof(float) * (d + 1));
    float *barycentric = (float *)malloc(sizeof(float) * (d + 2));
    short *rank = (short *)malloc(sizeof(short) * (d + 1));
    short *canonical = (short *)malloc(sizeof(short) * (d + 1) * (d + 1));
    short *key = (short *)malloc(sizeof(short) * (d + 1));

    /* canonical simplex */
    for (int i = 0; i <= d; i++) {
        for (int j = 0; j <= d - i; j++) canonical[i * (d + 1) + j] = (short)i;
        for (int j = d - i + 1; j <= d; j++) canonical[i * (d + 1) + j] = (short)(i - (d + 1));
    }

    /* expected std-dev of the filter; diagonal of the elevation matrix E */
    float inv_std_dev = (float)(sqrt(2.0 / 3.0) * (d + 1));
    for (int i = 0; i < d; i++)
        scale_factor[i] = (float)(1.0 / sqrt((double)((i + 2) * (i + 1))) * inv_std_dev);

    for (int k = 0; k < N; k++) {
        const float *f = feature + (size_t)k * d;

        /* elevate: y = E p */
        float sm = 0;
        for (int j = d; j > 0; j--) {
            float cf = f[j - 1] * scale_factor[j - 1];
            elevated[j] = sm - j * cf;
            sm += cf;
        }
        elevated[0] = sm;

        /* closest 0-coloured lattice point by rounding */
        float down_factor = 1.0f / (d + 1);
        float up_factor = (float)(d + 1);
        int sum = 0;
        for (int i = 0; i <= d; i++) {
            int rd2;
            float v = down_factor * elevated[i];
            float up = ceilf(v) * up_factor;
            float down = floorf(v) * up_factor;
            if (up - elevated[i] < elevated[i] - down)
                rd2 = (short)up;
            else
                rd2 = (short)down;
            rem0[i] = (float)rd2;
            sum += rd2 * down_factor; /* int += float: truncating, as upstream */
        }

        /* rank of each coordinate's residual */
        for (int i = 0; i <= d; i++) rank[i] = 0;
        for (int i = 0; i < d; i++) {
            double di = elevated[i] - rem0[i];
            for (int j = i + 1; j <= d; j++)
                if (di < elevated[j] - rem0[j])
                    rank[i]++;
                else
                    rank[j]++;
        }

        /* bring the point back onto the plane if sum != 0 */
        for (int i = 0; i <= d; i++) {
            rank[i] += sum;
            if (rank[i] < 0) {
                rank[i] += d + 1;
                rem0[i] += d + 1;
            } else if (rank[i] > d) {
                rank[i] -= d + 1;
                rem0[i] -= d + 1;
            }
        }

        /* barycentric coordinates */
        for (int i = 0; i <= d + 1; i++) barycentric[i] = 0;
        for (int i = 0; i <= d; i++) {
            float v = (elevated[i] - rem0[i]) * down_factor;
            barycentric[d - rank[i]] += v;
            barycentric[d - rank[i] + 1] -= v;
        }
        barycentric[0] += 1.0f + barycentric[d + 1];

        /* simplex vertices -> hash */
        for (int remainder = 0; remainder <= d; remainder++) {
            for (int i = 0; i < d; i++) key[i] = (short)(rem0[i] + canonical[remainder * (d + 1) + rank[i]]);
            L->offset[(size_t)k * (d + 1) + remainder] = ht_find(&ht, key, 1);
            L->barycentric[(size_t)k * (d + 1) + remainder] = barycentric[remainder];
        }
    }

    L->M = ht.filled;
    int M = L->M;
    L->keys = (short *)malloc(sizeof(short) * (size_t)(M > 0 ? M : 1) * d);
    memcpy(L->keys, ht.keys, sizeof(short) * (size_t)M * d);

    /* blur neighbours along each of the d+1 lattice axes */
    L->n1 = (int *)malloc(sizeof(int) * (size_t)(d + 1) * (M > 0 ? M : 1));
    L->n2 = (int *)malloc(sizeof(int) * (size_t)(d + 1) * (M > 0 ? M : 1));
    short *k1 = (short *)malloc(sizeof(short) * (d + 1));
    short *k2 = (short *)malloc(sizeof(short) * (d + 1));
    for (int j = 0; j <= d; j++) {
        for (int i = 0; i < M; i++) {
            const short *kk = L->keys + (size_t)i * d;
            for (int k = 0; k < d; k++) {
                k1[k] = (short)(kk[k] - 1);
                k2[k] = (short)(kk[k] + 1);
            }
            if (j < d) { /* coordinate d is implicit (coordinates sum to zero) */
                k1[j] = (short)(kk[j] + d);
                k2[j] = (short)(kk[j] - d);
            }
            L->n1[(size_t)j * M + i] = ht_find(&ht, k1, 0);
            L->n2[(size_t)j * M + i] = ht_find(&ht, k2, 0);
        }
    }

    free(k1);
    free(k2);
    free(scale_factor);
    free(elevated);
    free(rem0);
    free(barycentric);
    free(rank);
    free(canonical);
    free(key);
    ht_free(&ht);
    return L;
}

void pl_free(Lattice *L) {
    if (!L) return;
    free(L->offset);
    free(L->barycentric);
    free(L->n1);
    free(L->n2);
    free(L->keys);
    free(L);
}

int pl_M(const Lattice *L) { return L->M; }
int pl_N(const Lattice *L) { return L->N; }
int pl_d(const Lattice *L) { return L->d; }
const int *pl_offset(const Lattice *L) { return L->offset; }
const float *pl_barycentric(const Lattice *L) { return L->barycentric; }
const int *pl_n1(const Lattice *L) { return L->n1; }
const int *pl_n2(const Lattice *L) { return L->n2; }
const short *pl_keys(const Lattice *L) { return L->keys; }

/* out, in: [N][vs] pixel-major.  out may alias in. */
void pl_compute(const Lattice *L, float *out, const float *in, int vs, int reverse) {
    const int N = L->N, d = L->d, M = L->M;
    float *values = (float *)calloc((size_t)(M + 2) * vs, sizeof(float));
    float *new_values = (float *)calloc((size_t)(M + 2) * vs, sizeof(float));

    /* splat */
    for (int i = 0; i < N; i++) {
        for (int j = 0; j <= d; j++) {
            int o = L->offset[(size_t)i * (d + 1) + j] + 1;
            float w = L->barycentric[(size_t)i * (d + 1) + j];
            for (int k = 0; k < vs; k++) values[(size_t)o * vs + k] += w * in[(size_t)i * vs + k];
        }
    }

    /* blur along each axis */
    for (int jj = 0; jj <= d; jj++) {
        int j = reverse ? d - jj : jj;
        for (int i = 0; i < M; i++) {
            float *old_val = values + (size_t)(i + 1) * vs;
            float *new_val = new_values + (size_t)(i + 1) * vs;
            int n1 = L->n1[(size_t)j * M + i] + 1;
            int n2 = L->n2[(size_t)j * M + i] + 1;
            float *n1_val = values + (size_t)n1 * vs;
            float *n2_val = values + (size_t)n2 * vs;
            for (int k = 0; k < vs; k++) new_val[k] = old_val[k] + 0.5f * (n1_val[k] + n2_val[k]);
        }
        float *t = values;
        values = new_values;
        new_values = t;
    }

    /* slice; alpha normalises the blur weights */
    float alpha = 1.0f / (1 + powf(2, -d));
    for (int i = 0; i < N; i++) {
        for (int k = 0; k < vs; k++) out[(size_t)i * vs + k] = 0;
        for (int j = 0; j <= d; j++) {
            int o = L->offset[(size_t)i * (d + 1) + j] + 1;
            float w = L->barycentric[(size_t)i * (d + 1) + j];
            for (int k = 0; k < vs; k++) out[(size_t)i * vs + k] += w * values[(size_t)o * vs + k] * alpha;
        }
    }
    free(values);
    free(new_values);
}

/* ------------------------------------------------------------------------------------------
 * DenseKernel (DIAG_KERNEL, NORMALIZE_SYMMETRIC) + Potts + mean-field inference
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    Lattice *lattice;
    float *norm; /* [N] = 1/sqrt(K 1 + 1e-20) */
    float weight;
} DenseKernel;

static void kernel_init(DenseKernel *K, const float *feature, int N, int d, float weight) {
    K->lattice = pl_create(feature, N, d);
    K->weight = weight;
    K->norm = (float *)malloc(sizeof(float) * N);
    float *ones = (float *)malloc(sizeof(float) * N);
    for (int i = 0; i < N; i++) ones[i] = 1.0f;
    pl_compute(K->lattice, K->norm, ones, 1, 0);
    for (int i = 0; i < N; i++) K->norm[i] = (float)(1.0 / sqrt(K->norm[i] + 1e-20));
    free(ones);
}

static void kernel_free(DenseKernel *K) {
    pl_free(K->lattice);
    free(K->norm);
}

/* out = norm . K (norm . in); in/out [N][C] */
static void kernel_filter(const DenseKernel *K, float *out, const float *in, int C) {
    const int N = K->lattice->N;
    for (int i = 0; i < N; i++)
        for (int k = 0; k < C; k++) out[(size_t)i * C + k] = in[(size_t)i * C + k] * K->norm[i];
    pl_compute(K->lattice, out, out, C, 0);
    for (int i = 0; i < N; i++)
        for (int k = 0; k < C; k++) out[(size_t)i * C + k] = out[(size_t)i * C + k] * K->norm[i];
}

static void exp_and_normalize(float *out, const float *in, int N, int C) {
    for (int i = 0; i < N; i++) {
        const float *b = in + (size_t)i * C;
        float *o = out + (size_t)i * C;
        float mx = b[0];
        for (int k = 1; k < C; k++)
            if (b[k] > mx) mx = b[k];
        float s = 0;
        for (int k = 0; k < C; k++) {
            o[k] = expf(b[k] - mx);
            s += o[k];
        }
        for (int k = 0; k < C; k++) o[k] = o[k] / s;
    }
}

typedef struct {
    int W, H, C, N;
    float *unary; /* [N][C] */
    int n_kernels;
    DenseKernel kernels[8];
} DenseCRF2D;

DenseCRF2D *dcrf_create(int W, int H, int C) {
    DenseCRF2D *c = (DenseCRF2D *)calloc(1, sizeof(DenseCRF2D));
    c->W = W;
    c->H = H;
    c->C = C;
    c->N = W * H;
    c->unary = (float *)calloc((size_t)c->N * C, sizeof(float));
    return c;
}

void dcrf_free(DenseCRF2D *c) {
    if (!c) return;
    for (int k = 0; k < c->n_kernels; k++) kernel_free(&c->kernels[k]);
    free(c->unary);
    free(c);
}

/* unary_cn: float32 [C][N] C-contiguous, exactly what pydensecrf's setUnaryEnergy receives */
void dcrf_set_unary(DenseCRF2D *c, const float *unary_cn) {
    for (int k = 0; k < c->C; k++)
        for (int i = 0; i < c->N; i++) c->unary[(size_t)i * c->C + k] = unary_cn[(size_t)k * c->N + i];
}

int dcrf_add_pairwise_gaussian(DenseCRF2D *c, float sx, float sy, float w) {
    if (c->n_kernels >= 8) return -1;
    float *feature = (float *)malloc(sizeof(float) * (size_t)c->N * 2);
    for (int j = 0; j < c->H; j++)
        for (int i = 0; i < c->W; i++) {
            feature[(size_t)(j * c->W + i) * 2 + 0] = i / sx;
            feature[(size_t)(j * c->W + i) * 2 + 1] = j / sy;
        }
    kernel_init(&c->kernels[c->n_kernels++], feature, c->N, 2, w);
    free(feature);
    return 0;
}

/* rgb: uint8 [H][W][3] */
int dcrf_add_pairwise_bilateral(DenseCRF2D *c, float sx, float sy, float sr, float sg, float sb,
                                const unsigned char *rgb, float w) {
    if (c->n_kernels >= 8) return -1;
    float *feature = (float *)malloc(sizeof(float) * (size_t)c->N * 5);
    for (int j = 0; j < c->H; j++)
        for (int i = 0; i < c->W; i++) {
            size_t p = (size_t)(j * c->W + i);
            feature[p * 5 + 0] = i / sx;
            feature[p * 5 + 1] = j / sy;
            feature[p * 5 + 2] = rgb[p * 3 + 0] / sr;
            feature[p * 5 + 3] = rgb[p * 3 + 1] / sg;
            feature[p * 5 + 4] = rgb[p * 3 + 2] / sb;
        }
    kernel_init(&c->kernels[c->n_kernels++], feature, c->N, 5, w);
    free(feature);
    return 0;
}

int dcrf_kernel_M(const DenseCRF2D *c, int k) { return c->kernels[k].lattice->M; }
const float *dcrf_kernel_norm(const DenseCRF2D *c, int k) { return c->kernels[k].norm; }
const Lattice *dcrf_kernel_lattice(const DenseCRF2D *c, int k) { return c->kernels[k].lattice; }

/* Filter a [C][N] field through kernel k (symmetric-normalised), result [C][N].  Test helper. */
void dcrf_kernel_apply(const DenseCRF2D *c, int k, const float *in_cn, float *out_cn, int C) {
    int N = c->N;
    float *a = (float *)malloc(sizeof(float) * (size_t)N * C);
    float *b = (float *)malloc(sizeof(float) * (size_t)N * C);
    for (int ch = 0; ch < C; ch++)
        for (int i = 0; i < N; i++) a[(size_t)i * C + ch] = in_cn[(size_t)ch * N + i];
    kernel_filter(&c->kernels[k], b, a, C);
    for (int ch = 0; ch < C; ch++)
        for (int i = 0; i < N; i++) out_cn[(size_t)ch * N + i] = b[(size_t)i * C + ch];
    free(a);
    free(b);
}

/* Q_cn: float32 [C][N] (what np.array(d.inference(n)) yields) */
void dcrf_inference(const DenseCRF2D *c, int n_iterations, float *Q_cn) {
    const int N = c->N, C = c->C;
    size_t sz = (size_t)N * C;
    float *Q = (float *)malloc(sizeof(float) * sz);
    float *tmp1 = (float *)malloc(sizeof(float) * sz);
    float *tmp2 = (float *)malloc(sizeof(float) * sz);
    for (size_t i = 0; i < sz; i++) tmp1[i] = -c->unary[i];
    exp_and_normalize(Q, tmp1, N, C);
    for (int it = 0; it < n_iterations; it++) {
        for (size_t i = 0; i < sz; i++) tmp1[i] = -c->unary[i];
        for (int k = 0; k < c->n_kernels; k++) {
            kernel_filter(&c->kernels[k], tmp2, Q, C);
            float w = c->kernels[k].weight;
            for (size_t i = 0; i < sz; i++) tmp2[i] = -w * tmp2[i]; /* Potts */
            for (size_t i = 0; i < sz; i++) tmp1[i] -= tmp2[i];
        }
        exp_and_normalize(Q, tmp1, N, C);
    }
    for (int k = 0; k < C; k++)
        for (int i = 0; i < N; i++) Q_cn[(size_t)k * N + i] = Q[(size_t)i * C + k];
    free(Q);
    free(tmp1);
    free(tmp2);
}
