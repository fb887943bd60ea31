/*
 * zf_synth.c -- deterministic synthetic stereo PCM for benchmarks and parity tests.
 *
 * Envelope-modulated mixed sinusoids plus triangular noise from a counter-based PRNG, random
 * access by sample index so shards of a long stream can be generated independently per GPU.
 * Host-only utility; it is not part of the encode path.
 *
 *   pan  = 0.5 + 0.45 sin(2pi 0.23 t)
 *   tone = 0.30 sin(2pi 440 t) + 0.15 sin(2pi 1318.5 t + 0.7) + 0.08 sin(2pi 5274 t + 1.9)
 *   st   = 0.05 sin(2pi 659.3 t + 0.3)
 *   env  = (0.05 + 0.95 sin^2(2pi 3.1 t)) (0.25 + 0.75 sin^8(2pi 41 t))
 *   L = round(F env (pan tone + st) + A (u1 - u2)),  R = round(F env ((1 - pan) tone - st) + A (u3 - u4))
 *   F = 2^(b-1), A = 2^(b-9) env, u_i = splitmix64(seed, n, i) / 2^53
 *
 * Phases are reduced with exactly-rounded double operations (multiply, divide, floor) before the
 * libm sin call, so the arguments stay in [0, 2pi) for any stream position.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

static inline uint64_t splitmix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline double uniform01(uint64_t seed, uint64_t n, unsigned i) {
    uint64_t z = seed + (n * 4u + i + 1u) * 0x9E3779B97F4A7C15ull;
    return (double)(splitmix64(z) >> 11) * (1.0 / 9007199254740992.0);
}

static inline double phase_sin(double freq, uint64_t n, double fs, double offset) {
    double cycles = freq * (double)n / fs;
    cycles -= floor(cycles);
    return sin(6.283185307179586476925286766559 * cycles + offset);
}

typedef struct {
    uint8_t *out;
    uint64_t first, count;
    uint32_t sample_rate, bit_depth;
    uint64_t seed;
} synth_job;

static void synth_range(const synth_job *j, uint64_t i0, uint64_t i1) {
    const double fs = (double)j->sample_rate;
    const unsigned b = j->bit_depth;
    const double F = ldexp(1.0, (int)b - 1);
    const double A0 = ldexp(1.0, (int)b - 9);
    const unsigned bytes = b / 8;
    const int64_t lo = -((int64_t)1 << (b - 1)), hi = ((int64_t)1 << (b - 1)) - 1;
    for (uint64_t i = i0; i < i1; i++) {
        const uint64_t n = j->first + i;
        const double pan = 0.5 + 0.45 * phase_sin(0.23, n, fs, 0.0);
        const double tone = 0.30 * phase_sin(440.0, n, fs, 0.0) + 0.15 * phase_sin(1318.5, n, fs, 0.7) +
                            0.08 * phase_sin(5274.0, n, fs, 1.9);
        const double st = 0.05 * phase_sin(659.3, n, fs, 0.3);
        const double e1 = phase_sin(3.1, n, fs, 0.0);
        const double e2 = phase_sin(41.0, n, fs, 0.0);
        const double s2 = e2 * e2, s4 = s2 * s2, s8 = s4 * s4;
        const double env = (0.05 + 0.95 * e1 * e1) * (0.25 + 0.75 * s8);
        const double A = A0 * env;
        const double u1 = uniform01(j->seed, n, 0), u2 = uniform01(j->seed, n, 1);
        const double u3 = uniform01(j->seed, n, 2), u4 = uniform01(j->seed, n, 3);
        int64_t v[2];
        v[0] = llround(F * env * (pan * tone + st) + A * (u1 - u2));
        v[1] = llround(F * env * ((1.0 - pan) * tone - st) + A * (u3 - u4));
        uint8_t *o = j->out + i * 2 * bytes;
        for (int c = 0; c < 2; c++) {
            int64_t s = v[c] < lo ? lo : (v[c] > hi ? hi : v[c]);
            for (unsigned k = 0; k < bytes; k++) *o++ = (uint8_t)((uint64_t)s >> (8 * k));
        }
    }
}

typedef struct {
    const synth_job *job;
    uint64_t i0, i1;
} synth_slice;

static void *synth_thread(void *arg) {
    const synth_slice *s = (const synth_slice *)arg;
    synth_range(s->job, s->i0, s->i1);
    return NULL;
}

/* Writes `count` inter-channel samples (stereo, interleaved little-endian, bit_depth in {16,24,32})
 * starting at stream position `first_sample` into out (count * 2 * bit_depth/8 bytes). */
int zf_synth_pcm(uint8_t *out, uint64_t first_sample, uint64_t count, uint32_t sample_rate, uint32_t bit_depth,
                 uint64_t seed, int n_threads) {
    if ((bit_depth != 16 && bit_depth != 24 && bit_depth != 32) || sample_rate == 0 || !out) return -1;
    synth_job job = {out, first_sample, count, sample_rate, bit_depth, seed};
    if (n_threads < 1) n_threads = 1;
    if (n_threads == 1 || count < 65536) {
        synth_range(&job, 0, count);
        return 0;
    }
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    synth_slice *sl = (synth_slice *)malloc(sizeof(synth_slice) * (size_t)n_threads);
    if (!tid || !sl) { free(tid); free(sl); synth_range(&job, 0, count); return 0; }
    const uint64_t per = (count + (uint64_t)n_threads - 1) / (uint64_t)n_threads;
    int started = 0;
    for (int t = 0; t < n_threads; t++) {
        uint64_t i0 = per * (uint64_t)t, i1 = i0 + per;
        if (i0 >= count) break;
        if (i1 > count) i1 = count;
        sl[t].job = &job; sl[t].i0 = i0; sl[t].i1 = i1;
        if (pthread_create(&tid[started], NULL, synth_thread, &sl[t]) == 0) started++;
        else synth_range(&job, i0, i1);
    }
    for (int t = 0; t < started; t++) pthread_join(tid[t], NULL);
    free(tid);
    free(sl);
    return 0;
}
