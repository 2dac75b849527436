/* sanitize_driver.c -- runs the oracle's entry points under AddressSanitizer / UndefinedBehaviorSanitizer
 * (oracle/Makefile target `sanitize`; tests/test_sanitizers.py).  TEST INFRASTRUCTURE.
 *
 *   sanitize_oracle <scene.bin>
 * scene.bin: int32 n_walls, int32 bands, orc_trace_params (80 bytes), n_walls x 40-byte segments,
 *            n_walls x bands floats (when bands > 1).
 * Prints the histogram checksum and counters so the caller can compare them with the ordinary build. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float ax, ay, bx, by, nx, ny, absorption, scattering, transmission, ior; } seg_t;
typedef struct {
    float source_x, source_y, listener_x, listener_y, listener_radius, speed_of_sound, input_gain;
    int32_t max_bounce_count; uint32_t rng_state_offset; int32_t ray_count, debug_ray_count, sample_rate, impulse_length, bands;
    float time_divisor; uint32_t flags; int64_t ray_begin, ray_end;
} params_t;
typedef struct { uint64_t ray_bounces, nearest_tests, shadow_tests, direct_hits, nee_hits; } ctr_t;

int orc_trace(const seg_t *, int, const float *, const params_t *, int64_t *, void *, int64_t, int64_t *, ctr_t *, int);
void orc_ir_to_float(const int64_t *, int64_t, float *);
void orc_convolve(const float *, int32_t, const float *, int32_t, int32_t, float *, int);
void orc_synthesize_ir(const int64_t *, int32_t, int32_t, int32_t, const float *, float *);
int64_t orc_load_sample(const float *, int64_t, int32_t, int32_t, int32_t, float *);

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t n_walls = 0, bands = 0;
    params_t p;
    if (fread(&n_walls, 4, 1, f) != 1 || fread(&bands, 4, 1, f) != 1 || fread(&p, sizeof p, 1, f) != 1) return 2;
    seg_t *walls = malloc(sizeof(seg_t) * (size_t)(n_walls > 0 ? n_walls : 1));
    if (n_walls > 0 && fread(walls, sizeof(seg_t), (size_t)n_walls, f) != (size_t)n_walls) return 2;
    float *band_abs = NULL;
    if (bands > 1) {
        band_abs = malloc(sizeof(float) * (size_t)n_walls * bands);
        if (fread(band_abs, sizeof(float), (size_t)n_walls * bands, f) != (size_t)n_walls * bands) return 2;
    }
    fclose(f);
    const int64_t words = (int64_t)p.impulse_length * (bands > 1 ? bands : 1);
    int64_t *hist = calloc((size_t)words, sizeof(int64_t));   /* exactly the size the trace may touch */
    const int64_t cap = 1000;                                  /* deliberately smaller than the hit count */
    char *hits = malloc((size_t)cap * 24);
    int64_t n_hits = 0;
    ctr_t c;
    if (orc_trace(walls, n_walls, band_abs, &p, hist, hits, cap, &n_hits, &c, 2) != 0) return 3;
    uint64_t sum = 0;
    for (int64_t i = 0; i < words; i++) sum = sum * 1099511628211ull + (uint64_t)hist[i];
    printf("hist %llu hits %lld bounces %llu nearest %llu shadow %llu\n", (unsigned long long)sum, (long long)n_hits,
           (unsigned long long)c.ray_bounces, (unsigned long long)c.nearest_tests, (unsigned long long)c.shadow_tests);
    /* convolution, synthesis and clip preparation on exact-size buffers */
    float *ir = malloc(sizeof(float) * (size_t)words);
    orc_ir_to_float(hist, words, ir);
    if (bands > 1) {
        float *syn = malloc(sizeof(float) * (size_t)p.impulse_length);
        orc_synthesize_ir(hist, p.impulse_length, bands, 1, NULL, syn);
        memcpy(ir, syn, sizeof(float) * (size_t)p.impulse_length);
        free(syn);
    }
    enum { NX = 777 };
    float x[NX];
    for (int i = 0; i < NX; i++) x[i] = (float)((i * 2654435761u) >> 8) / 16777216.0f - 0.5f;
    float *y = malloc(sizeof(float) * (size_t)(NX + p.impulse_length));
    orc_convolve(x, NX, ir, p.impulse_length, 3, y, 2);
    double e = 0;
    for (int i = 0; i < NX + p.impulse_length; i++) e += (double)y[i] * y[i];
    int64_t n_out = orc_load_sample(x, NX / 3, 3, 44100, 48000, NULL);
    float *mono = malloc(sizeof(float) * (size_t)(n_out > 0 ? n_out : 1));
    orc_load_sample(x, NX / 3, 3, 44100, 48000, mono);
    printf("conv %.9g resampled %lld\n", e, (long long)n_out);
    free(mono); free(y); free(ir); free(hits); free(hist); free(band_abs); free(walls);
    return 0;
}
