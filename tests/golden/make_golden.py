"""Regenerates tests/golden/*.  The reference ships no golden vectors and cannot be run here (Zig
only), so these fixtures pin the ORACLE against regressions, not the oracle against the reference:

  kat_vectors.json      the known-answer vectors derived by hand in SURVEY.md 8-K (frame headers,
                        UTF-8 frame numbers, STREAMINFO, VORBIS_COMMENT, CRC/MD5 check values)
  oracle_digests.json   sha256 of the synthetic PCM and of the oracle's FLAC output for seeded inputs
  synth16_6000.flac     a complete small stream (16-bit stereo 44.1 kHz, 6000 samples) from the oracle

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402
import oracle_lib as O  # noqa: E402
import signals  # noqa: E402
import zigflac_b200 as zf  # noqa: E402


def main():
    digests = {}
    for bits, rate in ((16, 44100), (24, 96000), (32, 192000)):
        n = 4096 * 5 + 1234
        pcm = zf.synth_pcm(n, rate, bits)
        out, sizes = O.encode_pcm(pcm, n, O.config(2, bits), rate)
        digests[f"synth_{bits}_{rate}_{n}"] = {
            "pcm_sha256": hashlib.sha256(pcm.tobytes()).hexdigest(),
            "flac_sha256": hashlib.sha256(out.tobytes()).hexdigest(),
            "frame_sizes": [int(s) for s in sizes],
        }
        for name, L, R in signals.stereo_classes(bits):
            p = O.pcm_bytes_from_int(signals.interleave([L, R]), bits)
            o, s = O.encode_pcm(p, L.size, O.config(2, bits), 44100)
            digests[f"class_{name}_{bits}"] = {"flac_sha256": hashlib.sha256(o.tobytes()).hexdigest(),
                                               "bytes": int(o.size)}
    json.dump(digests, open(os.path.join(HERE, "oracle_digests.json"), "w"), indent=1, sort_keys=True)
    pcm = zf.synth_pcm(6000, 44100, 16)
    rc, flac = O.wav_to_flac(O.make_wav(pcm, 2, 16, 44100))
    assert rc == 0
    open(os.path.join(HERE, "synth16_6000.flac"), "wb").write(flac)
    open(os.path.join(HERE, "synth16_6000.pcm.sha256"), "w").write(hashlib.sha256(pcm.tobytes()).hexdigest() + "\n")
    print("golden fixtures written:", len(digests), "digests,", len(flac), "byte stream")


if __name__ == "__main__":
    main()
