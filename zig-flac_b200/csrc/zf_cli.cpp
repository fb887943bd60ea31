// zf_cli.cpp -- `flac in_file.wav out_file.flac` with the reference's argv / exit-code contract
// (src/cli.zig:7-27: exit 1 on bad usage; src/cli/wav2flac.zig:24-27: exit 2 on an unsupported format).
// Extra, optional: `--devices 0,1,2,3` shards the stream over several GPUs; `-d in.flac out.wav` decodes (extension: the
// reference has no decoder) and verifies every frame's CRC-16 and the STREAMINFO MD5; `-V in.wav out.flac` encodes, then
// decodes the result on the device and compares it with the WAV's samples (exit 4 on a mismatch).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/zigflac_b200.h"

int main(int argc, char **argv) {
    const char *input = nullptr, *output = nullptr;
    std::vector<int> devices;
    bool decode = false, verify = false;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-d") || !strcmp(argv[i], "--decode")) {
            decode = true;
        } else if (!strcmp(argv[i], "-V") || !strcmp(argv[i], "--verify")) {
            verify = true;
        } else if (!strcmp(argv[i], "--devices") && i + 1 < argc) {
            for (char *tok = strtok(argv[++i], ","); tok; tok = strtok(nullptr, ",")) devices.push_back(atoi(tok));
        } else if (!input) input = argv[i];
        else if (!output) output = argv[i];
    }
    if (!input || !output) {
        fprintf(stderr, "error: usage: flac in_file.wav out_file.flac\n");  // cli.zig:18
        return 1;
    }
    if (decode) {
        const int rc = zf_decode_flac_file(input, output, devices.empty() ? 0 : devices[0], ZF_DECODE_REQUIRE_MD5);
        if (rc) {
            fprintf(stderr, "error: %s (%d) %s\n", zf_strerror(rc), rc, zf_last_cuda_error());
            return 3;
        }
        return 0;
    }
    const int rc = zf_encode_wav_file(input, output, devices.empty() ? nullptr : devices.data(), (int)devices.size());
    if (rc == 2) {
        fprintf(stderr, "error: format: flac does not support this wav format\n");  // wav2flac.zig:25
        return 2;
    }
    if (rc) {
        fprintf(stderr, "error: %s (%d) %s\n", zf_strerror(rc), rc, zf_last_cuda_error());
        return 3;
    }
    if (verify) {
        const int vrc = zf_verify_flac_file(input, output, devices.empty() ? 0 : devices[0]);
        if (vrc) {
            fprintf(stderr, "error: verify: %s (%d) %s\n", zf_strerror(vrc), vrc, zf_last_cuda_error());
            return 4;
        }
    }
    return 0;
}
