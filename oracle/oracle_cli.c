/*
 * oracle_cli.c -- CPU ORACLE CLI (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 * Restates cli.zig:7-27 / wav2flac.zig:10-63: `flac_oracle in.wav out.flac [threads]`,
 * exit 1 on bad usage (cli.zig:17-20), exit 2 on an unsupported WAV format (wav2flac.zig:24-27).
 */
#include <stdio.h>
#include <stdlib.h>

#include "zigflac_oracle.h"

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "error: usage: flac in_file.wav out_file.flac\n");
        return 1;
    }
    int threads = argc > 3 ? atoi(argv[3]) : 1;
    FILE *in = fopen(argv[1], "rb");
    if (!in) { perror(argv[1]); return 3; }
    fseek(in, 0, SEEK_END);
    long len = ftell(in);
    fseek(in, 0, SEEK_SET);
    uint8_t *wav = (uint8_t *)malloc((size_t)len);
    if (!wav || fread(wav, 1, (size_t)len, in) != (size_t)len) { fprintf(stderr, "read failed\n"); return 3; }
    fclose(in);
    uint8_t *flac = NULL;
    size_t flac_len = 0;
    int rc = zo_wav_to_flac(wav, (size_t)len, &flac, &flac_len, threads);
    if (rc == 2) {
        fprintf(stderr, "error: format: flac does not support this wav format\n");
        return 2;
    }
    if (rc) { fprintf(stderr, "error: wav parse/encode failed (%d)\n", rc); return 3; }
    FILE *out = fopen(argv[2], "wb");
    if (!out || fwrite(flac, 1, flac_len, out) != flac_len) { perror(argv[2]); return 3; }
    fclose(out);
    zo_free(flac);
    free(wav);
    return 0;
}
