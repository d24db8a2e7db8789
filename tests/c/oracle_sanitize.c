/* Exercises every entry point of the CPU oracle on small inputs; built with -fsanitize=address,undefined by
 * tests/test_oracle.py::test_oracle_is_clean_under_asan_ubsan (sanitizer build of the oracle).
 * Exit code 0 = ran to completion; the sanitizers abort otherwise. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../oracle/gomel_oracle.h"

static double *vec(long n) { double *p = (double *)malloc((size_t)n * sizeof(double)); if (!p) exit(3); return p; }

int main(void)
{
    unsigned long long s = 99;
    const int geos[2][3] = { { 4096, 1280, 192 }, { 2048, 256, 160 } };
    for (int g = 0; g < 2; g++) {
        orc_config c;
        memset(&c, 0, sizeof c);
        c.resolut = geos[g][0]; c.window = geos[g][1]; c.num_mels = geos[g][2]; c.num_freqs = 768;
        c.mel_fmin = 0; c.mel_fmax = g ? 8000 : 16000; c.tune_mul = 1; c.tune_add = 0; c.volume_boost = 1.5; c.gl_iters = 3;
        const long lens[3] = { 1, 7001, 30011 };
        for (int li = 0; li < 3; li++) {
            const long n = lens[li];
            double *wav = vec(n);
            for (long i = 0; i < n; i++) {
                s = s * 6364136223846793005ULL + 1442695040888963407ULL;
                wav[i] = 0.4 * sin(0.01 * (double)i) + 0.1 * ((double)(s >> 11) / 9007199254740992.0 - 0.5);
            }
            const long np = n + orc_pad_len(n, c.window);
            const long fr = orc_num_frames(np, c.resolut, c.window);
            if (!orc_is_padded(n, np, c.window) || fr < 1) return 4;
            const long ola = c.resolut + (fr - 1) * (long)c.window;
            /* mel */
            const long nm = fr * c.num_mels * 2;
            double *mel = vec(nm), *init = vec(ola), *out = vec(ola);
            if (orc_to_mel(&c, wav, n, mel, nm) != fr) return 5;
            for (long i = 0; i < ola; i++) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; init[i] = (double)(s >> 11) / 9007199254740992.0; }
            uint16_t *img = (uint16_t *)malloc((size_t)(nm / 2) * 2);
            orc_mel_dumpbuffer(mel, nm / 2, c.num_mels, img);
            uint8_t *rgba = (uint8_t *)malloc((size_t)(nm / 2) * 4);
            orc_mel_quantise(mel, nm / 2, c.num_mels, 1, 1280.0, 44100.0, rgba);
            double *mel2 = vec(nm), samples = 0, sr = 0;
            orc_mel_dequantise(rgba, (int)fr, c.num_mels, 1, mel2, &samples, &sr);
            if (orc_from_mel(&c, mel, nm / 2, init, out, ola) != ola) return 6;
            free(mel); free(mel2); free(init); free(out); free(img); free(rgba);
            /* phase (native geometry only, like the reference's NewPhase) */
            if (g == 0) {
                const long ns = fr * c.num_freqs * 2;
                double *spec = vec(ns), *back = vec(ola);
                if (orc_to_phase(&c, wav, n, spec, ns) != fr) return 7;
                if (orc_from_phase(&c, spec, ns / 2, back, ola) != ola) return 8;
                for (int hdr = 0; hdr < 2; hdr++) {
                    double *tmp = vec(ns);
                    memcpy(tmp, spec, (size_t)ns * sizeof(double));
                    uint8_t *o8 = (uint8_t *)malloc((size_t)(ns / 2) * 4);
                    uint16_t *o16 = (uint16_t *)malloc((size_t)(ns / 2) * 8);
                    orc_phase_quantise(tmp, ns / 2, c.num_freqs, 1, 1280.0, 48000.0, hdr ? 0 : 2, hdr, o8, o16);
                    orc_phase_dequantise(o8, o16, (int)fr, c.num_freqs, 1, hdr ? 0 : 2, hdr, tmp, &samples, &sr);
                    free(tmp); free(o8); free(o16);
                }
                uint16_t *pimg = (uint16_t *)malloc((size_t)(ns / 2) * 2);
                orc_phase_dumpbuffer(spec, ns / 2, c.num_freqs, pimg);
                free(pimg); free(spec); free(back);
            }
            free(wav);
        }
    }
    /* helpers */
    int zp = 0, zs = 0;
    orc_pad_shift(22050, &zp, &zs);
    double a[10] = { 1, 2, 3, 4, 5, 6, 7, 8, 9, 10 };
    const long zl = orc_zero_stuff_len(10, 3, 2);
    double *z = vec(zl);
    orc_zero_stuff(a, 10, 3, 2, z);
    free(z);
    if (orc_f16_value(orc_f16_bits(44100.0)) != 44096.0) return 9;
    int lo[192], hi[192]; double mod[192];
    orc_mel_fwd_tables(2048, 192, 0, 16000, lo, hi, mod);
    printf("oracle sanitize ok\n");
    return 0;
}
