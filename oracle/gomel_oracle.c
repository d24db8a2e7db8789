/*
 * gomel_oracle.c -- CPU float64 ORACLE (test infrastructure, NOT product code).
 * See gomel_oracle.h for scope, provenance and the parity-pinning statement.
 * Every function cites the reference file:line it restates (paths relative to the reference).
 */
#include "gomel_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ pad / frames / window */

/* mel/impl.go:429-455 == phase/impl.go:424-450 == phase.py:352-377 */
long orc_pad_len(long n, int filter)
{
    long min_target = 15L * filter;
    long pad = 0;
    if (n >= min_target) {
        long rem = (n - min_target) % filter;
        if (rem != 0) pad = filter - rem - 1;
    } else {
        pad = min_target - n - 1;
    }
    return pad > 0 ? pad : 0;
}

/* mel/impl.go:457-479 == phase/impl.go:452-474 */
int orc_is_padded(long original_len, long padded_len, int filter)
{
    long min_target = 15L * filter;
    if (original_len >= min_target) {
        long rem = (original_len - min_target) % filter;
        if (rem != 0) return padded_len == original_len + (filter - rem - 1);
        return padded_len == original_len;
    }
    return padded_len == original_len + (min_target - original_len - 1);
}

/* gossp stft.NumFrames [UPSTREAM]; restated by the reference at phase.py:121 */
long orc_num_frames(long n_padded, int frame_len, int frame_shift)
{
    return (long)((double)(n_padded - frame_len) / (double)frame_shift) + 1;
}

/* go-dsp window.Hann via gossp stft.New [UPSTREAM]; reference restates it as np.hanning (phase.py:122) */
void orc_hann(int n, double *w)
{
    if (n == 1) { w[0] = 1.0; return; }
    for (int i = 0; i < n; i++) w[i] = 0.5 * (1.0 - cos(2.0 * M_PI * (double)i / (double)(n - 1)));
}

/* ------------------------------------------------------------------ FFT (go-dsp definition) */

/* iterative radix-2 DIT, forward kernel e^{-2 pi i k n / N}; twiddles from cos/sin in float64 */
static void fft_pow2(double *re, double *im, int n)
{
    int lg = 0;
    while ((1 << lg) < n) lg++;
    for (int i = 0; i < n; i++) {
        int r = 0;
        for (int b = 0; b < lg; b++) if (i & (1 << b)) r |= 1 << (lg - 1 - b);
        if (r > i) {
            double t = re[i]; re[i] = re[r]; re[r] = t;
            t = im[i]; im[i] = im[r]; im[r] = t;
        }
    }
    double *tw = (double *)malloc(sizeof(double) * (size_t)n);   /* n/2 complex */
    for (int k = 0; k < n / 2; k++) {
        double a = -2.0 * M_PI * (double)k / (double)n;
        tw[2 * k] = cos(a); tw[2 * k + 1] = sin(a);
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int s = 0; s < n; s += len) {
            for (int j = 0; j < half; j++) {
                double wr = tw[2 * j * step], wi = tw[2 * j * step + 1];
                int a = s + j, b = a + half;
                double xr = re[b] * wr - im[b] * wi;
                double xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr; im[b] = im[a] - xi;
                re[a] += xr;        im[a] += xi;
            }
        }
    }
    free(tw);
}

/* go-dsp fft.FFT / fft.IFFT [UPSTREAM]: IFFT(x) = reverse x[1..], forward FFT, divide by N */
void orc_fft(double *re, double *im, int n, int inverse)
{
    if (!inverse) { fft_pow2(re, im, n); return; }
    for (int i = 1, j = n - 1; i < j; i++, j--) {
        double t = re[i]; re[i] = re[j]; re[j] = t;
        t = im[i]; im[i] = im[j]; im[j] = t;
    }
    fft_pow2(re, im, n);
    double inv = (double)n;
    for (int i = 0; i < n; i++) { re[i] /= inv; im[i] /= inv; }
}

/* gossp STFT.STFT [UPSTREAM] at call sites mel/mel.go:50-52, phase/phase.go:45-47:
 * frame i = x[i*hop : i*hop+N] * hann, full complex FFT (fft.FFTReal promotes to complex). */
static void stft_frame(const double *x, long i, int N, int H, const double *w, double *re, double *im)
{
    const double *f = x + i * (long)H;
    for (int j = 0; j < N; j++) { re[j] = f[j] * w[j]; im[j] = 0.0; }
    orc_fft(re, im, N, 0);
}

/* ------------------------------------------------------------------ mel scale helpers */

/* mel/impl.go:298-302 */
static double mel_to_hz(double v) { return 700.0 * (exp(v / 1127.0) - 1.0); }
/* mel/impl.go:304-308 */
static double hz_to_mel(double v) { return 1127.0 * log(1.0 + (v / 700.0)); }

/* the per-band quantities of domel, mel/impl.go:313-323 */
void orc_mel_fwd_tables(int filtersize, int mels, double fmin, double fmax, int *lo, int *hi, double *mod)
{
    double melbin = hz_to_mel(fmax) / (double)mels;
    for (int i = 0; i < mels; i++) {
        double vallo = (double)filtersize * (fmin + mel_to_hz(melbin * (double)i)) / (fmax + fmin);
        double valhi = (double)filtersize * (fmin + mel_to_hz(melbin * (double)(i + 1))) / (fmax + fmin);
        double inlo, modlo = modf(vallo, &inlo);
        double inhi = floor(valhi);
        if (inlo < 0) { inlo = 0; modlo = 0; inhi = 0; }
        lo[i] = (int)inlo; hi[i] = (int)inhi; mod[i] = modlo;
    }
}

/* the per-bin quantities of undomel, mel/impl.go:350-360 */
void orc_mel_inv_tables(int filtersize, int mels, double fmin, double fmax, int *lo, int *hi,
                        double *mod, double *inlo_f, double *inhi_f)
{
    double filterbin = hz_to_mel(fmax) / (double)mels;
    for (int i = 0; i < filtersize; i++) {
        double vallo = hz_to_mel(((double)i * (fmax + fmin) / (double)filtersize) - fmin) / filterbin;
        double valhi = hz_to_mel(((double)(i + 1) * (fmax + fmin) / (double)filtersize) - fmin) / filterbin;
        double inlo, modlo = modf(vallo, &inlo);
        double inhi = floor(valhi);
        if (inlo < 0) { inlo = 0; modlo = 0; inhi = 0; }
        lo[i] = (int)inlo; hi[i] = (int)inhi; mod[i] = modlo;
        if (inlo_f) inlo_f[i] = inlo;
        if (inhi_f) inhi_f[i] = inhi;
    }
}

/* ------------------------------------------------------------------ mel.ToMel */

/* mel/mel.go:46-74.  returns frames, -1 bad args, -2 out buffer too small, -3 index out of range
 * where Go would panic */
long orc_to_mel(const orc_config *c, const double *wav, long n, double *out, long out_cap)
{
    int N = c->resolut, H = c->window, B = N / 2, mels = c->num_mels;
    if (n <= 0 || N <= 0 || H <= 0 || mels <= 0) return -1;
    long np_ = n + orc_pad_len(n, H);                       /* mel/mel.go:48 */
    if (np_ < N) return -1;
    long frames = orc_num_frames(np_, N, H);                /* mel/mel.go:50-52 */
    if (frames * (long)mels * 2 > out_cap) return -2;
    double *x = (double *)calloc((size_t)np_, sizeof(double));
    memcpy(x, wav, sizeof(double) * (size_t)n);
    double *w = (double *)malloc(sizeof(double) * (size_t)N);
    orc_hann(N, w);
    double *re = (double *)malloc(sizeof(double) * (size_t)N), *im = (double *)malloc(sizeof(double) * (size_t)N);
    double *s0 = (double *)malloc(sizeof(double) * (size_t)B), *s1 = (double *)malloc(sizeof(double) * (size_t)B);
    int *lo = (int *)malloc(sizeof(int) * (size_t)mels), *hi = (int *)malloc(sizeof(int) * (size_t)mels);
    double *mod = (double *)malloc(sizeof(double) * (size_t)mels);
    orc_mel_fwd_tables(B, mels, c->mel_fmin, c->mel_fmax, lo, hi, mod);
    long rc = frames;
    for (long i = 0; i < frames && rc >= 0; i++) {
        stft_frame(x, i, N, H, w, re, im);
        for (int j = 0; j < B; j++) {                        /* mel/mel.go:54-66 */
            s0[j] = hypot(re[j], im[j]);                     /* |X[j]|       */
            s1[j] = hypot(re[N - j - 1], im[N - j - 1]);     /* |X[N-1-j]|   */
        }
        for (int m = 0; m < mels; m++) {                     /* domel, mel/impl.go:310-345 */
            for (int l = 0; l < 2; l++) {
                const double *s = l ? s1 : s0;
                double total = 0.0;
                if (lo[m] + 1 == hi[m]) {
                    if (hi[m] >= B && i == frames - 1) { rc = -3; break; }  /* Go: index out of range */
                    /* within a non-final frame Go reads the next frame's bin 0; not reachable at
                     * the benchmark configs -- flag it rather than emulate */
                    if (hi[m] >= B) { rc = -3; break; }
                    total += s[lo[m]] * (1 - mod[m]);
                    total += s[hi[m]] * mod[m];
                } else {
                    if (hi[m] > B) { rc = -3; break; }
                    for (int k = lo[m]; k < hi[m]; k++) total += s[k];
                    total /= (double)(hi[m] - lo[m] + 1);
                }
                /* spectral_normalize, mel/impl.go:410-419 */
                if (total < 1e-5) total = 1e-5;
                out[(i * mels + m) * 2 + l] = log(total);
            }
        }
    }
    free(x); free(w); free(re); free(im); free(s0); free(s1); free(lo); free(hi); free(mod);
    return rc;
}

/* ------------------------------------------------------------------ mel.FromMel + Griffin-Lim */

/* mel/mel.go:142-152: spectral_denormalize (in place) -> undomel -> undospectrum -> ISTFT.
 * n_entries = len(ospectrum) = frames*num_mels. returns ola_len, or <0. */
long orc_from_mel(const orc_config *c, double *mel, long n_entries, const double *init,
                  double *out, long out_cap)
{
    int N = c->resolut, H = c->window, B = N / 2, mels = c->num_mels;
    if (n_entries <= 0 || mels <= 0 || n_entries % mels != 0) return -1;   /* Go: panics (mel/impl.go:366-372) */
    long frames = n_entries / mels;
    long ola = (long)N + (frames - 1) * (long)H;             /* mel/mel.go:79 */
    if (ola > out_cap) return -2;

    for (int l = 0; l < 2; l++)                              /* mel/impl.go:421-427 (mutates caller) */
        for (long i = 0; i < n_entries; i++) mel[2 * i + l] = exp(mel[2 * i + l]);

    int *lo = (int *)malloc(sizeof(int) * (size_t)B), *hi = (int *)malloc(sizeof(int) * (size_t)B);
    double *mod = (double *)malloc(sizeof(double) * (size_t)B);
    double *flo = (double *)malloc(sizeof(double) * (size_t)B), *fhi = (double *)malloc(sizeof(double) * (size_t)B);
    orc_mel_inv_tables(B, mels, c->mel_fmin, c->mel_fmax, lo, hi, mod, flo, fhi);

    /* full complex spectrogram [frames][N], as the reference builds it (mel/impl.go:386-408) */
    double *Sre = (double *)calloc((size_t)frames * (size_t)N, sizeof(double));
    double *Sim = (double *)calloc((size_t)frames * (size_t)N, sizeof(double));
    long rc = ola;
    for (long f = 0; f < frames && rc >= 0; f++) {
        const double *m = mel + f * (long)mels * 2;
        for (int i = 0; i < B; i++) {                        /* undomel, mel/impl.go:347-384 */
            double tot[2];
            for (int l = 0; l < 2; l++) {
                double total = 0.0;
                if (lo[i] == hi[i]) {
                    if (lo[i] >= mels) { rc = -3; break; }
                    total += m[2 * lo[i] + l];
                } else if (lo[i] + 1 == hi[i] && hi[i] < mels) {
                    total += m[2 * lo[i] + l] * (1 - mod[i]);
                    total += m[2 * hi[i] + l] * mod[i];
                } else {
                    if (hi[i] > mels) { rc = -3; break; }
                    for (int k = lo[i]; k < hi[i]; k++) total += m[2 * k + l];
                    total /= fhi[i] - flo[i] + 1;
                }
                tot[l] = total;
            }
            if (rc < 0) break;
            /* undospectrum, mel/impl.go:386-408: Rect(real,0) at [j] and [N-1-j] */
            double r0 = (tot[0] - c->tune_add) / c->tune_mul;
            double r1 = (tot[1] - c->tune_add) / c->tune_mul;
            Sre[f * N + i] = r0;          Sim[f * N + i] = 0.0;
            Sre[f * N + (N - i - 1)] = r1; Sim[f * N + (N - i - 1)] = 0.0;
        }
    }
    free(lo); free(hi); free(mod); free(flo); free(fhi);
    if (rc < 0) { free(Sre); free(Sim); return rc; }

    /* ISTFT, mel/mel.go:76-139 */
    double *w = (double *)malloc(sizeof(double) * (size_t)N);
    orc_hann(N, w);
    double *sig = (double *)malloc(sizeof(double) * (size_t)ola);
    double *nsig = (double *)malloc(sizeof(double) * (size_t)ola);
    memcpy(sig, init, sizeof(double) * (size_t)ola);         /* mel/mel.go:80-83 (injected) */
    double *re = (double *)malloc(sizeof(double) * (size_t)N), *im = (double *)malloc(sizeof(double) * (size_t)N);
    for (int iter = 0; iter < c->gl_iters; iter++) {
        for (long f = 0; f < frames; f++) {                  /* phase update, mel/mel.go:87-109 */
            for (int j = 0; j < N; j++) {
                long pos = f * (long)H + j;
                re[j] = (pos < ola) ? sig[pos] * w[j] : 0.0;
                im[j] = 0.0;
            }
            orc_fft(re, im, N, 0);                           /* fft.FFTReal */
            double *sr = Sre + f * (long)N, *si = Sim + f * (long)N;
            for (int j = 0; j < N; j++) {
                double magnitude = hypot(sr[j], si[j]);      /* cmplx.Abs   */
                double ph = atan2(im[j], re[j]);             /* cmplx.Phase */
                sr[j] = magnitude * cos(ph);                 /* cmplx.Rect  */
                si[j] = magnitude * sin(ph);
            }
            for (int j = 1; j < N / 2; j++) {                /* conj symmetry, mel/mel.go:105-108 */
                sr[N - j] = sr[j]; si[N - j] = -si[j];
            }
        }
        memset(nsig, 0, sizeof(double) * (size_t)ola);       /* mel/mel.go:112 */
        for (long f = 0; f < frames; f++) {                  /* mel/mel.go:115-125 */
            memcpy(re, Sre + f * (long)N, sizeof(double) * (size_t)N);
            memcpy(im, Sim + f * (long)N, sizeof(double) * (size_t)N);
            orc_fft(re, im, N, 1);                           /* fft.IFFT */
            for (int j = 0; j < N; j++) {
                long pos = f * (long)H + j;
                if (pos < ola) nsig[pos] += re[j] * w[j];
            }
        }
        double *t = sig; sig = nsig; nsig = t;               /* mel/mel.go:135 */
    }
    memcpy(out, sig, sizeof(double) * (size_t)ola);
    free(w); free(sig); free(nsig); free(re); free(im); free(Sre); free(Sim);
    return ola;
}

/* ------------------------------------------------------------------ phase.ToPhase / FromPhase */

/* phase/phase.go:41-70 + shrink phase/impl.go:383-391 */
long orc_to_phase(const orc_config *c, const double *wav, long n, double *out, long out_cap)
{
    int N = c->resolut, H = c->window, B = N / 2, nf = c->num_freqs;
    if (n <= 0 || N <= 0 || H <= 0 || nf <= 0) return -1;
    long np_ = n + orc_pad_len(n, H);
    if (np_ < N) return -1;
    long frames = orc_num_frames(np_, N, H);
    int keep = nf < B ? nf : B;                              /* shrink keeps j < omels */
    if (frames * (long)keep * 2 > out_cap) return -2;
    double *x = (double *)calloc((size_t)np_, sizeof(double));
    memcpy(x, wav, sizeof(double) * (size_t)n);
    double *w = (double *)malloc(sizeof(double) * (size_t)N);
    orc_hann(N, w);
    double *re = (double *)malloc(sizeof(double) * (size_t)N), *im = (double *)malloc(sizeof(double) * (size_t)N);
    for (long i = 0; i < frames; i++) {
        stft_frame(x, i, N, H, w, re, im);
        for (int j = 0; j < keep; j++) {
            out[(i * keep + j) * 2 + 0] = im[j + 1];         /* imag(spectrum[i][j+1])        */
            out[(i * keep + j) * 2 + 1] = re[N - j - 1];     /* real(spectrum[i][Resolut-j-1]) */
        }
    }
    free(x); free(w); free(re); free(im);
    return frames;
}

/* phase/phase.go:136-153: grow (phase/impl.go:392-403) -> undospectrum (phase/phase.go:72-91)
 * -> ISTFT (phase/phase.go:93-133) -> VolumeBoost */
long orc_from_phase(const orc_config *c, const double *spec, long n_entries, double *out, long out_cap)
{
    int N = c->resolut, H = c->window, B = N / 2, nf = c->num_freqs;
    if (n_entries <= 0 || nf <= 0 || nf > B || n_entries % nf != 0) return -1;
    long frames = n_entries / nf;
    long ola = (long)N + (frames - 1) * (long)H;
    if (ola > out_cap) return -2;
    double *w = (double *)malloc(sizeof(double) * (size_t)N);
    orc_hann(N, w);
    double *re = (double *)malloc(sizeof(double) * (size_t)N), *im = (double *)malloc(sizeof(double) * (size_t)N);
    double *ws = (double *)calloc((size_t)ola, sizeof(double));
    memset(out, 0, sizeof(double) * (size_t)ola);
    for (long f = 0; f < frames; f++) {
        const double *s = spec + f * (long)nf * 2;
        memset(re, 0, sizeof(double) * (size_t)N);
        memset(im, 0, sizeof(double) * (size_t)N);
        for (int j = 0; j < B; j++) {
            int src = j < nf ? j : nf - 1;                   /* grow: replicate last kept entry */
            double realn1 = s[2 * src + 0], realm0 = s[2 * src + 1];
            re[j + 1] = realm0;     im[j + 1] = realn1;      /* v0 = complex(realm0, realn1)  */
            re[N - j - 1] = realm0; im[N - j - 1] = -realn1; /* v1 = conj, written second     */
        }
        orc_fft(re, im, N, 1);
        for (int j = 0; j < N; j++) {
            long pos = f * (long)H + j;
            if (pos < ola) { out[pos] += re[j] * w[j]; ws[pos] += w[j] * w[j]; }
        }
    }
    double maxws = 0.0;
    for (long i = 0; i < ola; i++) if (ws[i] > maxws) maxws = ws[i];
    double thr = maxws * 0.5;
    for (long i = 0; i < ola; i++) {
        if (ws[i] > thr) out[i] /= ws[i];
        else if (ws[i] > 1e-21) out[i] = out[i] / ws[i] * (ws[i] / thr);
    }
    if (c->volume_boost != 0)                                /* phase/phase.go:146-150 */
        for (long i = 0; i < ola; i++) out[i] *= c->volume_boost;
    free(w); free(re); free(im); free(ws);
    return ola;
}

/* ------------------------------------------------------------------ Image / dumpbuffer */

/* Go `int(f)` on amd64 (CVTTSD2SQ): truncation; NaN / out of range -> 0x8000000000000000 */
static int64_t go_int(double f)
{
    if (f != f || f >= 9223372036854775808.0 || f < -9223372036854775808.0) return INT64_MIN;
    return (int64_t)f;
}

static void dumpbuffer_common(const double *buf, long n_entries, int mels, uint16_t *out, int mel_scan)
{
    long stride = n_entries / mels;
    double mx[2] = { -99999999., -99999999. }, mn[2] = { 9999999., 9999999. };
    for (int l = 0; l < 2; l++)
        for (long x = 0; x < stride; x++)
            for (int y = 0; y < mels; y++) {
                /* mel/impl.go:24 scans buf[stride*y+x]; phase/impl.go:23 scans buf[y+x*mels] */
                long idx = mel_scan ? stride * y + x : y + x * (long)mels;
                double w = buf[2 * idx + l];
                if (w > mx[l]) mx[l] = w;
                if (w < mn[l]) mn[l] = w;
            }
    long o = 0;
    for (long x = 0; x < stride; x++)
        for (int y = 0; y < mels; y++) {
            double v0 = (buf[2 * (y + x * (long)mels) + 0] - mn[0]) / (mx[0] - mn[0]);
            double v1 = (buf[2 * (y + x * (long)mels) + 1] - mn[1]) / (mx[1] - mn[1]);
            out[o++] = (uint16_t)((uint16_t)go_int(255 * v0) | (uint16_t)((uint16_t)go_int(255 * v1) << 8));
        }
}
/* mel/impl.go:16-44 */
void orc_mel_dumpbuffer(const double *buf, long n, int mels, uint16_t *out) { dumpbuffer_common(buf, n, mels, out, 1); }
/* phase/impl.go:15-43 */
void orc_phase_dumpbuffer(const double *buf, long n, int mels, uint16_t *out) { dumpbuffer_common(buf, n, mels, out, 0); }

/* ------------------------------------------------------------------ float16 metadata */

/* x448/float16 Fromfloat32: IEEE round-to-nearest-even f32 -> f16 */
static uint16_t f32_to_f16(float f)
{
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t exp = (x >> 23) & 0xffu, man = x & 0x7fffffu;
    if (exp == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? (0x200u | (man >> 13)) : 0));
    int e = (int)exp - 127 + 15;
    if (e >= 31) return (uint16_t)(sign | 0x7c00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t)sign;
        man |= 0x800000u;
        int shift = 14 - e;
        uint32_t half = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1), mid = 1u << (shift - 1);
        if (rem > mid || (rem == mid && (half & 1))) half++;
        return (uint16_t)(sign | half);
    }
    uint32_t half = ((uint32_t)e << 10) | (man >> 13);
    uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) half++;
    return (uint16_t)(sign | half);
}
static float f16_to_f32(uint16_t h)
{
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ffu, x;
    if (exp == 0) {
        if (man == 0) x = sign;
        else {
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            x = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) x = sign | 0x7f800000u | (man << 13);
    else x = sign | ((exp - 15 + 127) << 23) | (man << 13);
    float f; memcpy(&f, &x, 4); return f;
}
/* mel/impl.go:120-125 packFloat16ToBytes */
uint16_t orc_f16_bits(double f) { return f32_to_f16((float)f); }
/* mel/impl.go:46-50 unpackBytesToFloat64 */
double orc_f16_value(uint16_t bits) { return (double)f16_to_f32(bits); }

/* ------------------------------------------------------------------ mel PNG pixel arithmetic */

/* mel/impl.go:127-193 (everything except os.Create / png.Encode) */
void orc_mel_quantise(const double *buf, long n_entries, int mels, int reverse,
                      double samples_in_mel, double sr, uint8_t *rgba)
{
    long stride = n_entries / mels;
    double mx = -1.79769313486231570814527423731704357e+308, mn = 1.79769313486231570814527423731704357e+308;
    for (long x = 0; x < stride; x++)
        for (int l = 0; l < 2; l++)
            for (int y = 0; y < mels; y++) {
                double w = buf[2 * (stride * y + x) + l];     /* mel/impl.go:143 */
                if (w > mx) mx = w;
                if (w < mn) mn = w;
            }
    uint8_t floats[8];
    uint16_t h[4] = { orc_f16_bits(mx), orc_f16_bits(mn), orc_f16_bits(samples_in_mel), orc_f16_bits(sr) };
    for (int i = 0; i < 4; i++) { floats[2 * i] = (uint8_t)(h[i] & 0xff); floats[2 * i + 1] = (uint8_t)(h[i] >> 8); }
    for (long x = 0; x < stride; x++)
        for (int y = 0; y < mels; y++) {
            double v0 = (buf[2 * (y + x * (long)mels) + 0] - mn) / (mx - mn);
            double v1 = (buf[2 * (y + x * (long)mels) + 1] - mn) / (mx - mn);
            uint8_t R = (uint8_t)go_int(255 * v0), G = (uint8_t)go_int(255 * v1), Bc = 0;
            int meta_start = mels - 8;
            if (x == 0 && y >= meta_start) Bc = floats[y - meta_start];
            int yy = reverse ? mels - y - 1 : y;
            uint8_t *p = rgba + 4 * ((long)yy * stride + x);
            p[0] = R; p[1] = G; p[2] = Bc; p[3] = 255;
        }
}

/* mel/impl.go:52-118 (everything after png.Decode).  8-bit NRGBA with A=255: color.RGBA() gives
 * v*0x101, so r>>8 == v. buf: width*height entries, index x*height + y */
void orc_mel_dequantise(const uint8_t *rgba, int width, int height, int reverse,
                        double *buf, double *samples, double *samplerate)
{
    int mels = height;
    uint8_t floats[8]; int nf = 0;
    long o = 0;
    for (int x = 0; x < width; x++)
        for (int y = 0; y < height; y++) {
            int yy = reverse ? height - y - 1 : y;
            const uint8_t *p = rgba + 4 * ((long)yy * width + x);
            int meta_start = mels - 8;
            if (x == 0 && y >= meta_start && nf < 8) floats[nf++] = p[2];
            buf[2 * o + 0] = (double)p[0] / 255;
            buf[2 * o + 1] = (double)p[1] / 255;
            o++;
        }
    double mx = orc_f16_value((uint16_t)(floats[0] | (floats[1] << 8)));
    double mn = orc_f16_value((uint16_t)(floats[2] | (floats[3] << 8)));
    double sim = orc_f16_value((uint16_t)(floats[4] | (floats[5] << 8)));
    double sr = orc_f16_value((uint16_t)(floats[6] | (floats[7] << 8)));
    if (mx == sim) sim = 0;                                   /* mel/impl.go:105-107 */
    for (long i = 0; i < o; i++) {
        buf[2 * i + 0] = buf[2 * i + 0] * (mx - mn) + mn;
        buf[2 * i + 1] = buf[2 * i + 1] * (mx - mn) + mn;
    }
    *samples = sim * (double)width;
    *samplerate = sr;
}

/* ------------------------------------------------------------------ phase PNG pixel arithmetic */

/* phase/impl.go:168-278 */
void orc_phase_quantise(double *buf, long n_entries, int mels, int reverse, double samples_in_mel,
                        double sr, int ihs_passes, int hdr, uint8_t *out8, uint16_t *out16)
{
    for (int p = 0; p < ihs_passes; p++)
        for (long i = 0; i < n_entries; i++)
            for (int l = 0; l < 2; l++) buf[2 * i + l] = asinh(buf[2 * i + l]);
    long stride = n_entries / mels;
    int maxVal = hdr ? 65535 : 255;
    double mx[2] = { -1.79769313486231570814527423731704357e+308, -1.79769313486231570814527423731704357e+308 };
    double mn[2] = { 1.79769313486231570814527423731704357e+308, 1.79769313486231570814527423731704357e+308 };
    for (long x = 0; x < stride; x++)
        for (int l = 0; l < 2; l++)
            for (int y = 0; y < mels; y++) {
                double w = buf[2 * (y + x * (long)mels) + l];
                if (w > mx[l]) mx[l] = w;
                if (w < mn[l]) mn[l] = w;
            }
    uint16_t h[8] = { orc_f16_bits(mx[0]), orc_f16_bits(mx[1]), orc_f16_bits(0), orc_f16_bits(mn[0]),
                      orc_f16_bits(mn[1]), orc_f16_bits(0), orc_f16_bits(samples_in_mel), orc_f16_bits(sr) };
    uint8_t floats[16];
    for (int i = 0; i < 8; i++) { floats[2 * i] = (uint8_t)(h[i] & 0xff); floats[2 * i + 1] = (uint8_t)(h[i] >> 8); }
    for (long x = 0; x < stride; x++)
        for (int y = 0; y < mels; y++) {
            double v0 = (buf[2 * (y + x * (long)mels) + 0] - mn[0]) / (mx[0] - mn[0]);
            double v1 = (buf[2 * (y + x * (long)mels) + 1] - mn[1]) / (mx[1] - mn[1]);
            double v2 = -v0;
            int meta_start = mels - 16;
            int yy = reverse ? mels - y - 1 : y;
            long o = 4 * ((long)yy * stride + x);
            if (hdr) {
                out16[o + 0] = (uint16_t)go_int((double)maxVal * v0);
                out16[o + 1] = (uint16_t)go_int((double)maxVal * v1);
                out16[o + 2] = (x == 0 && y >= meta_start) ? (uint16_t)floats[y - meta_start]
                                                           : (uint16_t)go_int((double)maxVal * v2);
                out16[o + 3] = 65535;
            } else {
                out8[o + 0] = (uint8_t)go_int((double)maxVal * v0);
                out8[o + 1] = (uint8_t)go_int((double)maxVal * v1);
                out8[o + 2] = (x == 0 && y >= meta_start) ? floats[y - meta_start]
                                                          : (uint8_t)go_int((double)maxVal * v2);
                out8[o + 3] = 255;
            }
        }
}

/* phase/impl.go:51-153 */
void orc_phase_dequantise(const uint8_t *in8, const uint16_t *in16, int width, int height,
                          int reverse, int ihs_passes, int hdr, double *buf,
                          double *samples, double *samplerate)
{
    int mels = height, maxVal = hdr ? 65535 : 255;
    uint8_t floats[16]; int nf = 0;
    long o = 0;
    for (int x = 0; x < width; x++)
        for (int y = 0; y < height; y++) {
            int yy = reverse ? height - y - 1 : y;
            long p = 4 * ((long)yy * width + x);
            int meta_start = mels - 16;
            uint32_t r, g, b;
            if (hdr) { r = in16[p]; g = in16[p + 1]; b = in16[p + 2]; }
            else { r = in8[p] * 0x101u; g = in8[p + 1] * 0x101u; b = in8[p + 2] * 0x101u; }
            if (x == 0 && y >= meta_start && nf < 16) floats[nf++] = hdr ? (uint8_t)(b & 0xff) : (uint8_t)(b >> 8);
            if (hdr) { buf[2 * o] = (double)r / (double)maxVal; buf[2 * o + 1] = (double)g / (double)maxVal; }
            else { buf[2 * o] = (double)(r >> 8) / 255; buf[2 * o + 1] = (double)(g >> 8) / 255; }
            o++;
        }
    double v[8];
    for (int i = 0; i < 8; i++) v[i] = orc_f16_value((uint16_t)(floats[2 * i] | (floats[2 * i + 1] << 8)));
    double mx0 = v[0], mx1 = v[1], mn0 = v[3], mn1 = v[4], sim = v[6], sr = v[7];
    for (long i = 0; i < o; i++) {
        buf[2 * i + 0] = buf[2 * i + 0] * (mx0 - mn0) + mn0;
        buf[2 * i + 1] = buf[2 * i + 1] * (mx1 - mn1) + mn1;
    }
    for (int p = 0; p < ihs_passes; p++)
        for (long i = 0; i < o; i++)
            for (int l = 0; l < 2; l++) buf[2 * i + l] = sinh(buf[2 * i + l]);
    *samples = sim * (double)width;
    *samplerate = sr;
}

/* ------------------------------------------------------------------ zero stuffing */

/* phase/impl.go:476-507 */
void orc_pad_shift(int sr, int *zp, int *zs)
{
    *zp = 0; *zs = 0;
    switch (sr) {
    case 32000: *zp = 2; *zs = 1; break;
    case 24000: *zp = 1; *zs = 1; break;
    case 16000: *zp = 1; *zs = 2; break;
    case 8000:  *zp = 1; *zs = 5; break;
    case 22050: *zp = 1; *zs = 1; break;
    case 11025: *zp = 1; *zs = 3; break;
    default: break;
    }
}
long orc_zero_stuff_len(long n, int zp, int zs)
{
    if (zp == 0) return n;
    return n + ((n + zp - 1) / zp) * zs;
}
/* phase/impl.go:509-529 */
void orc_zero_stuff(const double *audio, long n, int zp, int zs, double *out)
{
    if (zp == 0) { memcpy(out, audio, sizeof(double) * (size_t)n); return; }
    long total = orc_zero_stuff_len(n, zp, zs), o = 0;
    memset(out, 0, sizeof(double) * (size_t)total);
    double boost = (double)(1 + zs);
    for (long i = 0; i < n; i++) {
        out[o++] = audio[i] * boost;
        if ((i + 1) % zp == 0) o += zs;
    }
}

/* ------------------------------------------------------------------ bench helpers */

long orc_from_mel_batch(const orc_config *c, double *mel, long n_entries_per_clip, int n_clips,
                        const double *init, double *out, long ola_len, int threads)
{
    long rc = ola_len;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int i = 0; i < n_clips; i++) {
        long r = orc_from_mel(c, mel + (long)i * n_entries_per_clip * 2, n_entries_per_clip,
                              init + (long)i * ola_len, out + (long)i * ola_len, ola_len);
        if (r < 0) rc = r;
    }
    return rc;
}

long orc_to_mel_batch(const orc_config *c, const double *wav, long n_per_clip, int n_clips,
                      double *out, long out_per_clip, int threads)
{
    long rc = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int i = 0; i < n_clips; i++) {
        long r = orc_to_mel(c, wav + (long)i * n_per_clip, n_per_clip, out + (long)i * out_per_clip, out_per_clip);
        if (r < 0) rc = r; else if (rc >= 0) rc = r;
    }
    return rc;
}
