"""Development probe: per-batch timeline (ZF_TRACE=1) of one end-to-end config-2 call."""
import sys, os
os.environ["ZF_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zigflac_b200 as zf
bits, rate, n = 24, 96000, 57600000
per = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
h_pcm = torch.empty(n * 6, dtype=torch.uint8, pin_memory=True)
zf.synth_pcm(n, rate, bits, out=h_pcm.numpy())
enc = zf.Encoder(zf.Config.default(2, bits), rate, max_frames_per_batch=per)
cap = enc.max_batch_bytes((n + 4095) // 4096)
h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
for _ in range(3):
    sys.stderr.write("---- call ----\n")
    enc.encode_pcm(h_pcm.numpy(), n, 0, out=h_out.numpy())
enc.close()
