"""Where the time of a whole-file encode goes (development aid): handle creation, page-locking, repeated library calls
with the driver's cache warm, and the CLI in a fresh process.  Prints one line per measurement."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import zigflac_b200 as zf
import oracle_lib


def t(f, n=1):
    t0 = time.perf_counter()
    for _ in range(n):
        r = f()
    return (time.perf_counter() - t0) / n * 1e3, r


ms, _ = t(lambda: zf.device_available())
print(f"first CUDA call (context): {ms:.1f} ms")
for per in (64, 2048):
    ms, e = t(lambda: zf.Encoder(zf.Config.default(2, 24), 96000, max_frames_per_batch=per))
    ms2, _ = t(e.close)
    print(f"Encoder create (max_frames_per_batch {per}): {ms:.1f} ms, close {ms2:.1f} ms")
for mb in (16, 128):
    ms, hb = t(lambda: zf.HostBuffer(mb << 20))
    ms2, _ = t(hb.close)
    print(f"zf_host_alloc {mb} MiB: {ms:.1f} ms, free {ms2:.1f} ms")
for bits, rate, secs in ((16, 44100, 1), (16, 44100, 60), (24, 96000, 600)):
    n = rate * secs
    pcm = zf.synth_pcm(n, rate, bits)
    wav = np.frombuffer(oracle_lib.make_wav(pcm, 2, bits, rate), dtype=np.uint8)
    first, (rc, flac) = t(lambda: zf.wav_to_flac(wav))
    again, _ = t(lambda: zf.wav_to_flac(wav), 3)
    md5 = zf.Md5()
    hms, _ = t(lambda: md5.update(pcm))
    print(f"{bits}-bit {secs} s ({pcm.size / 1e6:.1f} MB): first call {first:.1f} ms, warm {again:.1f} ms, MD5 alone {hms:.1f} ms, status {rc}")
    fin, fout = "/dev/shm/zf_probe.wav", "/dev/shm/zf_probe.flac"
    wav.tofile(fin)
    cms, r = t(lambda: subprocess.run([os.path.join(ROOT, "zig-flac_b200", "flac"), fin, fout], capture_output=True))
    ok = r.returncode == 0 and open(fout, "rb").read() == flac
    print(f"   CLI fresh process: {cms:.1f} ms, exit {r.returncode}, same bytes {ok}")
    for f in (fin, fout):
        if os.path.exists(f):
            os.remove(f)
