import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import zigflac_b200 as zf, oracle_lib
pcm = zf.synth_pcm(96000*600, 96000, 24)
wav = np.frombuffer(oracle_lib.make_wav(pcm, 2, 24, 96000), dtype=np.uint8)
for i in range(3):
    t0=time.perf_counter(); rc, flac = zf.wav_to_flac(wav); print("call", i, round((time.perf_counter()-t0)*1e3,1), "ms", rc, flush=True)
