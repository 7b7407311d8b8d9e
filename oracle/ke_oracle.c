/*
 * ke_oracle.c — CPU restatement of the arithmetic behind kobato-eyes' duplicate-detection
 * hot path.  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py, never by the product path.
 *
 * What each function restates (reference paths relative to /root/reference):
 *
 *   ko_rgb_to_l, ko_resample_*   src/sig/phash.py:21-26  `image.convert("L").resize(size, LANCZOS)`.
 *       The arithmetic lives in Pillow (pinned 11.3.0 in requirements.txt; 12.2.0 installed):
 *       libImaging/Convert.c rgb2l (16.16 fixed point) and libImaging/Resample.c
 *       (precompute_coeffs / normalize_coeffs_8bpc / ImagingResampleHorizontal_8bpc /
 *       ImagingResampleVertical_8bpc: 22-bit fixed-point taps, horizontal pass first, uint8
 *       intermediate).  Pillow is not under /root/reference; this is a restatement of its
 *       published algorithm, pinned byte-for-byte against the installed Pillow by
 *       tests/test_oracle_pinned.py and tests/golden/.
 *   ko_dhash_from_plane          src/sig/phash.py:49-57  (9x8 plane, left<right, MSB first)
 *   ko_phash_from_plane_f64      src/sig/phash.py:33-46  with the DCT in double instead of
 *       cv2.dct's float32 (IPP, closed source).  The PYTHON oracle (oracle/ref_py.py) calls the
 *       real cv2.dct; this double version is what the CUDA kernel is designed to equal exactly,
 *       and the test-suite measures how often the two disagree (near-ties only).
 *   ko_hamming_join              src/dup/scanner.py:227-290 as a set: {i<j : popcount(a^b)<=T
 *       [and some band equal]} (src/sig/phash.py:60-63 for the distance).
 *   ko_ssim_u8                   src/dup/refine.py:44-52 -> skimage.metrics.structural_similarity
 *       (scikit-image 0.25.2, absent here: "parity unpinned" for SSIM values) with
 *       scipy.ndimage.uniform_filter's arithmetic (double running sums, float32 between axes).
 *
 * Single-threaded on purpose: oracle/__init__.py fans calls out over Python threads (ctypes
 * releases the GIL), so the library needs nothing beyond libc/libm.
 * Build: make -C oracle   (gcc -O2 -shared)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KO_PRECISION_BITS (32 - 8 - 2)

/* ---------------------------------------------------------------- luma ---- */

/* Pillow Convert.c: L = (R*19595 + G*38470 + B*7471 + 0x8000) >> 16 */
static inline uint8_t ko_l_of_rgb(uint32_t r, uint32_t g, uint32_t b) {
    return (uint8_t)((r * 19595u + g * 38470u + b * 7471u + 0x8000u) >> 16);
}

/* img: h*w*c interleaved (c = 1 gray, 3 RGB, 4 RGBA/RGBX; alpha ignored like Pillow RGBA->L) */
void ko_rgb_to_l(const uint8_t* img, int h, int w, int c, int64_t row_stride, uint8_t* out) {
    for (int y = 0; y < h; ++y) {
        const uint8_t* row = img + (int64_t)y * row_stride;
        uint8_t* o = out + (int64_t)y * w;
        if (c == 1) {
            memcpy(o, row, (size_t)w);
        } else {
            for (int x = 0; x < w; ++x) o[x] = ko_l_of_rgb(row[x * c], row[x * c + 1], row[x * c + 2]);
        }
    }
}

/* ------------------------------------------------------ resample tables ---- */

static double ko_sinc(double x) {
    if (x == 0.0) return 1.0;
    x *= M_PI;
    return sin(x) / x;
}
static double ko_lanczos3(double x) {
    if (-3.0 <= x && x < 3.0) return ko_sinc(x) * ko_sinc(x / 3.0);
    return 0.0;
}
static double ko_bilinear(double x) {
    if (x < 0.0) x = -x;
    return x < 1.0 ? 1.0 - x : 0.0;
}
static double ko_bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}
static double ko_box(double x) { return (x > -0.5 && x <= 0.5) ? 1.0 : 0.0; }

/* filter ids follow PIL.Image.Resampling: 1 LANCZOS, 2 BILINEAR, 3 BICUBIC, 4 BOX */
static int ko_filter(int id, double (**f)(double), double* support) {
    switch (id) {
        case 1: *f = ko_lanczos3; *support = 3.0; return 0;
        case 2: *f = ko_bilinear; *support = 1.0; return 0;
        case 3: *f = ko_bicubic; *support = 2.0; return 0;
        case 4: *f = ko_box; *support = 0.5; return 0;
    }
    return -1;
}

int ko_resample_ksize(int in_size, int out_size, int filter_id) {
    double (*f)(double);
    double support;
    if (ko_filter(filter_id, &f, &support) || in_size <= 0 || out_size <= 0) return -1;
    double scale = (double)((float)in_size - 0.0f) / out_size;
    double fs = scale < 1.0 ? 1.0 : scale;
    return (int)ceil(support * fs) * 2 + 1;
}

/* kk: out_size*ksize fixed-point taps, bounds: out_size*2 (xmin, count). */
int ko_resample_table(int in_size, int out_size, int filter_id, int32_t* kk, int32_t* bounds, int ksize) {
    double (*f)(double);
    double support;
    if (ko_filter(filter_id, &f, &support)) return -1;
    if (ksize != ko_resample_ksize(in_size, out_size, filter_id)) return -2;
    double scale = (double)((float)in_size - 0.0f) / out_size;
    double fs = scale < 1.0 ? 1.0 : scale;
    support *= fs;
    double* k = (double*)malloc(sizeof(double) * (size_t)ksize);
    if (!k) return -3;
    for (int xx = 0; xx < out_size; ++xx) {
        double center = 0.0 + (xx + 0.5) * scale;
        double ww = 0.0, ss = 1.0 / fs;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        int x;
        for (x = 0; x < xmax; ++x) {
            double w = f((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (x = 0; x < xmax; ++x)
            if (ww != 0.0) k[x] /= ww;
        for (; x < ksize; ++x) k[x] = 0.0;
        for (x = 0; x < ksize; ++x) {
            double v = k[x] * (double)(1 << KO_PRECISION_BITS);
            kk[(int64_t)xx * ksize + x] = v < 0 ? (int32_t)(-0.5 + v) : (int32_t)(0.5 + v);
        }
        bounds[xx * 2] = xmin;
        bounds[xx * 2 + 1] = xmax;
    }
    free(k);
    return 0;
}

static inline uint8_t ko_clip8(int32_t v) {
    v >>= KO_PRECISION_BITS; /* arithmetic shift, like Pillow's lookup index */
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

/* 8bpc single-band resize, Pillow order: horizontal pass (if widths differ) then vertical. */
int ko_resample_u8(const uint8_t* in, int ih, int iw, uint8_t* out, int oh, int ow, int filter_id) {
    const uint8_t* src = in;
    uint8_t* tmp = NULL;
    int rc = 0;
    if (iw != ow) {
        int ks = ko_resample_ksize(iw, ow, filter_id);
        int32_t* kk = (int32_t*)malloc(sizeof(int32_t) * (size_t)ks * ow);
        int32_t* bd = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)ow);
        tmp = (uint8_t*)malloc((size_t)ih * ow);
        rc = ko_resample_table(iw, ow, filter_id, kk, bd, ks);
        for (int y = 0; y < ih && !rc; ++y)
            for (int xx = 0; xx < ow; ++xx) {
                int32_t ss = 1 << (KO_PRECISION_BITS - 1);
                const int32_t* k = kk + (int64_t)xx * ks;
                const uint8_t* p = src + (int64_t)y * iw + bd[xx * 2];
                for (int x = 0; x < bd[xx * 2 + 1]; ++x) ss += (int32_t)p[x] * k[x];
                tmp[(int64_t)y * ow + xx] = ko_clip8(ss);
            }
        free(kk);
        free(bd);
        src = tmp;
    }
    if (!rc && ih != oh) {
        int ks = ko_resample_ksize(ih, oh, filter_id);
        int32_t* kk = (int32_t*)malloc(sizeof(int32_t) * (size_t)ks * oh);
        int32_t* bd = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)oh);
        rc = ko_resample_table(ih, oh, filter_id, kk, bd, ks);
        for (int yy = 0; yy < oh && !rc; ++yy)
            for (int x = 0; x < ow; ++x) {
                int32_t ss = 1 << (KO_PRECISION_BITS - 1);
                const int32_t* k = kk + (int64_t)yy * ks;
                for (int y = 0; y < bd[yy * 2 + 1]; ++y) ss += (int32_t)src[(int64_t)(y + bd[yy * 2]) * ow + x] * k[y];
                out[(int64_t)yy * ow + x] = ko_clip8(ss);
            }
        free(kk);
        free(bd);
    } else if (!rc) {
        memcpy(out, src, (size_t)oh * ow);
    }
    free(tmp);
    return rc;
}

/* ------------------------------------------------------------- hashes ---- */

uint64_t ko_dhash_from_plane(const uint8_t* p /* 8 rows x 9 cols */) {
    uint64_t v = 0;
    for (int r = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c) v = (v << 1) | (uint64_t)(p[r * 9 + c + 1] > p[r * 9 + c]);
    return v;
}

/* Orthonormal DCT-II (what cv2.dct computes), low 8x8 block, all in double.
 * min_margin (nullable) receives min |coef - mean| over the 64 compared coefficients. */
uint64_t ko_phash_from_plane_f64(const uint8_t* p /* 32x32 */, double* min_margin) {
    static double C[8][32];
    static int init = 0;
    if (!init) {
        for (int k = 0; k < 8; ++k)
            for (int n = 0; n < 32; ++n)
                C[k][n] = (k == 0 ? sqrt(1.0 / 32.0) : sqrt(2.0 / 32.0)) * cos(M_PI * (2 * n + 1) * k / 64.0);
        init = 1;
    }
    double t[8][32]; /* t = C * X  (rows 0..7) */
    for (int k = 0; k < 8; ++k)
        for (int x = 0; x < 32; ++x) {
            double s = 0.0;
            for (int n = 0; n < 32; ++n) s += C[k][n] * (double)p[n * 32 + x];
            t[k][x] = s;
        }
    double d[64];
    for (int k = 0; k < 8; ++k)
        for (int l = 0; l < 8; ++l) {
            double s = 0.0;
            for (int n = 0; n < 32; ++n) s += t[k][n] * C[l][n];
            d[k * 8 + l] = s;
        }
    double mean = 0.0;
    for (int i = 1; i < 64; ++i) mean += d[i];
    mean /= 63.0;
    uint64_t v = 0;
    double mm = INFINITY;
    for (int i = 0; i < 64; ++i) {
        v = (v << 1) | (uint64_t)(d[i] > mean);
        double m = fabs(d[i] - mean);
        if (m < mm) mm = m;
    }
    if (min_margin) *min_margin = mm;
    return v;
}

/* Full signature of one decoded image (c in {1,3,4}); planes are optional outputs. */
int ko_signature(const uint8_t* img, int h, int w, int c, int64_t row_stride, uint64_t* ph, uint64_t* dh,
                 double* min_margin, uint8_t* plane32 /* 1024 */, uint8_t* plane9x8 /* 72 */) {
    uint8_t* l = (uint8_t*)malloc((size_t)h * w);
    uint8_t p32[1024], p98[72];
    if (!l) return -3;
    ko_rgb_to_l(img, h, w, c, row_stride, l);
    int rc = ko_resample_u8(l, h, w, p32, 32, 32, 1);
    if (!rc) rc = ko_resample_u8(l, h, w, p98, 8, 9, 1);
    free(l);
    if (rc) return rc;
    if (ph) *ph = ko_phash_from_plane_f64(p32, min_margin);
    if (dh) *dh = ko_dhash_from_plane(p98);
    if (plane32) memcpy(plane32, p32, 1024);
    if (plane9x8) memcpy(plane9x8, p98, 72);
    return 0;
}

int ko_signature_batch(const uint8_t* imgs, int64_t n, int h, int w, int c, int64_t img_stride, int64_t row_stride,
                       uint64_t* ph, uint64_t* dh, double* min_margin) {
    int rc_all = 0;
    for (int64_t i = 0; i < n; ++i) {
        int rc = ko_signature(imgs + i * img_stride, h, w, c, row_stride, ph + i, dh + i,
                              min_margin ? min_margin + i : NULL, NULL, NULL);
        if (rc) rc_all = rc;
    }
    return rc_all;
}

/* --------------------------------------------------------------- join ---- */

static inline int ko_band_equal(uint64_t x, int band_bits, int band_count) {
    uint64_t mask = band_bits >= 64 ? ~0ull : ((1ull << band_bits) - 1);
    for (int b = 0; b < band_count; ++b)
        if (((x >> (b * band_bits)) & mask) == 0) return 1;
    return 0;
}

/* All pairs i<j with popcount(h[i]^h[j]) <= T (and, when require_band, at least one equal
 * band).  Rows i are restricted to [row_begin,row_end) so a caller can sample stripes of a
 * large table.  Returns the total number of hits; the first `capacity` are stored
 * (order unspecified). */
int64_t ko_hamming_join(const uint64_t* h, int64_t n, int threshold, int require_band, int band_bits, int band_count,
                        int64_t row_begin, int64_t row_end, uint32_t* out_i, uint32_t* out_j, uint8_t* out_d,
                        int64_t capacity) {
    int64_t count = 0;
    if (row_end > n) row_end = n;
    for (int64_t i = row_begin; i < row_end; ++i) {
        const uint64_t a = h[i];
        for (int64_t j = i + 1; j < n; ++j) {
            uint64_t x = a ^ h[j];
            int d = __builtin_popcountll(x);
            if (d <= threshold && (!require_band || ko_band_equal(x, band_bits, band_count))) {
                int64_t slot = count++;
                if (slot < capacity) {
                    out_i[slot] = (uint32_t)i;
                    out_j[slot] = (uint32_t)j;
                    out_d[slot] = (uint8_t)d;
                }
            }
        }
    }
    return count;
}

/* --------------------------------------------------------------- SSIM ---- */

/* scipy.ndimage.uniform_filter1d along one axis for the "valid" positions only (the reference
 * crops the border afterwards, so the reflect padding never reaches a kept value): double
 * running sum, result stored as float32. in: rows x cols float, stride-agnostic via (n, step). */
static void ko_uniform1d_valid(const float* in, float* out, int n, int step, int win) {
    /* scipy ni_filters.c NI_UniformFilter1D keeps a running MEAN in double:
     * tmp = sum/size; tmp += (new - old)/size. */
    double tmp = 0.0;
    for (int i = 0; i < win; ++i) tmp += (double)in[(int64_t)i * step];
    tmp /= (double)win;
    out[0] = (float)tmp;
    for (int i = 1; i + win <= n; ++i) {
        tmp += ((double)in[(int64_t)(i + win - 1) * step] - (double)in[(int64_t)(i - 1) * step]) / (double)win;
        out[(int64_t)i * step] = (float)tmp;
    }
}

/* uniform_filter(size=7) on a float32 image, axis 0 then axis 1, valid region only:
 * out is (h-6) x (w-6), row-major. */
static void ko_uniform2d_valid(const float* img, int h, int w, float* out, float* scratch /* (h-6)*w */) {
    const int win = 7, oh = h - 6, ow = w - 6;
    for (int x = 0; x < w; ++x) ko_uniform1d_valid(img + x, scratch + x, h, w, win); /* axis 0 */
    for (int y = 0; y < oh; ++y) {
        /* axis 1 */
        float row[ow > 0 ? ow : 1];
        ko_uniform1d_valid(scratch + (int64_t)y * w, row, w, 1, win);
        memcpy(out + (int64_t)y * ow, row, sizeof(float) * (size_t)ow);
    }
}

/* structural_similarity(a/255, b/255, data_range=1.0): 7x7 uniform window, sample covariance,
 * float32 maps, float64 mean over the interior. Returns NaN when a side is < 7 (the reference
 * raises ValueError there). */
double ko_ssim_u8(const uint8_t* a, const uint8_t* b, int h, int w, int64_t stride_a, int64_t stride_b) {
    if (h < 7 || w < 7) return NAN;
    const int oh = h - 6, ow = w - 6;
    size_t npx = (size_t)h * w, nout = (size_t)oh * ow;
    float* x = (float*)malloc(sizeof(float) * npx * 3);
    float* f = (float*)malloc(sizeof(float) * (nout * 5 + (size_t)oh * w));
    if (!x || !f) {
        free(x);
        free(f);
        return NAN;
    }
    float *fa = x, *fb = x + npx, *prod = x + 2 * npx;
    float *ux = f, *uy = f + nout, *uxx = f + 2 * nout, *uyy = f + 3 * nout, *uxy = f + 4 * nout, *scr = f + 5 * nout;
    for (int y = 0; y < h; ++y)
        for (int i = 0; i < w; ++i) {
            fa[(size_t)y * w + i] = (float)a[y * stride_a + i] / 255.0f;
            fb[(size_t)y * w + i] = (float)b[y * stride_b + i] / 255.0f;
        }
    ko_uniform2d_valid(fa, h, w, ux, scr);
    ko_uniform2d_valid(fb, h, w, uy, scr);
    for (size_t i = 0; i < npx; ++i) prod[i] = fa[i] * fa[i];
    ko_uniform2d_valid(prod, h, w, uxx, scr);
    for (size_t i = 0; i < npx; ++i) prod[i] = fb[i] * fb[i];
    ko_uniform2d_valid(prod, h, w, uyy, scr);
    for (size_t i = 0; i < npx; ++i) prod[i] = fa[i] * fb[i];
    ko_uniform2d_valid(prod, h, w, uxy, scr);
    const float cov_norm = (float)(49.0 / 48.0), C1 = (float)(0.01 * 0.01), C2 = (float)(0.03 * 0.03);
    double acc = 0.0;
    for (size_t i = 0; i < nout; ++i) {
        float vx = cov_norm * (uxx[i] - ux[i] * ux[i]);
        float vy = cov_norm * (uyy[i] - uy[i] * uy[i]);
        float vxy = cov_norm * (uxy[i] - ux[i] * uy[i]);
        float A1 = 2 * ux[i] * uy[i] + C1, A2 = 2 * vxy + C2;
        float B1 = ux[i] * ux[i] + uy[i] * uy[i] + C1, B2 = vx + vy + C2;
        float D = B1 * B2;
        acc += (double)((A1 * A2) / D);
    }
    free(x);
    free(f);
    return acc / (double)nout;
}

/* Exact-arithmetic SSIM (integer window sums, double formula): the value the CUDA kernel
 * is designed to reproduce; used to measure the reference's own float32 rounding noise. */
double ko_ssim_u8_exact(const uint8_t* a, const uint8_t* b, int h, int w, int64_t stride_a, int64_t stride_b) {
    if (h < 7 || w < 7) return NAN;
    const double k1c = 1e-4 * 49.0 * 49.0 * 255.0 * 255.0, k2c = 9e-4 * 48.0 * 49.0 * 255.0 * 255.0;
    double acc = 0.0;
    for (int y = 0; y + 7 <= h; ++y)
        for (int x0 = 0; x0 + 7 <= w; ++x0) {
            int64_t su = 0, sv = 0, suu = 0, svv = 0, suv = 0;
            for (int dy = 0; dy < 7; ++dy)
                for (int dx = 0; dx < 7; ++dx) {
                    int64_t u = a[(y + dy) * stride_a + x0 + dx], v = b[(y + dy) * stride_b + x0 + dx];
                    su += u; sv += v; suu += u * u; svv += v * v; suv += u * v;
                }
            double p = (double)(su * sv), q = (double)(su * su + sv * sv);
            double vxy = (double)(49 * suv - su * sv), vs = (double)(49 * suu - su * su + 49 * svv - sv * sv);
            acc += ((2 * p + k1c) * (2 * vxy + k2c)) / ((q + k1c) * (vs + k2c));
        }
    return acc / ((double)(h - 6) * (w - 6));
}

int ko_ssim_batch(const uint8_t* bank, int h, int w, int64_t img_stride, const int64_t* ia, const int64_t* ib,
                  int64_t n_pairs, double* out, int exact) {
    for (int64_t p = 0; p < n_pairs; ++p) {
        const uint8_t *a = bank + ia[p] * img_stride, *b = bank + ib[p] * img_stride;
        out[p] = exact ? ko_ssim_u8_exact(a, b, h, w, w, w) : ko_ssim_u8(a, b, h, w, w, w);
    }
    return 0;
}

